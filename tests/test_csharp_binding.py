"""CPU-side contract of the C# host layer.  The build image has no C#/Unity toolchain, so the shipped
sources (planetmodel-sph_b200/csharp/*.cs) are checked against include/sphb200.h by parsing both:

* every SPH_API function has exactly one [DllImport] with the same number of parameters and a compatible type in every
  position (and nothing is imported that the header does not declare);
* every struct the binding mirrors has the header's fields, in order, with the header's element types and array lengths;
* every status code, flag, gravity selector and field id has the header's value;
* the six system classes keep the reference's names, update group, relative order and public constants
  (A/Systems/*.cs: ParticleSmoothingSystem.cs:14-18, KernelSystem.cs:15-17, GravityFieldSystem.cs:14-26,228,
  DensityFieldSystem.cs:8-9, PressureFieldSystem.cs:12-14, VelocitySystem.cs:15-16) and make exactly one native call each.
"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "sphb200.h")
NATIVE = os.path.join(ROOT, "planetmodel-sph_b200", "csharp", "SphB200Native.cs")
SYSTEMS = os.path.join(ROOT, "planetmodel-sph_b200", "csharp", "SphB200Systems.cs")


def strip_comments(src, line_first=False):
    """C header: block comments, then line comments.  C# sources (line_first): the other way round -- they use // only, and one
    of those comments mentions "Systems/*.cs"."""
    if line_first:
        src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", " ", src)


# ---------------------------------------------------------------------------------------------------- header side
def header_functions():
    src = strip_comments(open(HDR).read())
    out = {}
    for ret, name, args in re.findall(r"SPH_API\s+([\w\s\*]+?)\b(sphb200_\w+)\s*\(([^)]*)\)\s*;", src):
        params = []
        for a in args.split(","):
            a = a.strip()
            if a in ("", "void"):
                continue
            a = re.sub(r"\bconst\b", "", a)
            m = re.match(r"^\s*([\w\s]+?)\s*(\**)\s*(\w+)\s*$", a)
            assert m, a
            params.append(re.sub(r"\s+", " ", m.group(1)).strip() + m.group(2))
        out[name] = (re.sub(r"\bconst\b", "", ret).replace(" ", ""), params)
    return out


def header_structs():
    src = strip_comments(open(HDR).read())
    out = {}
    for body, name in re.findall(r"typedef\s+struct\s*\w*\s*\{([^}]*)\}\s*(\w+)\s*;", src):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            ty, rest = decl.split(None, 1)
            for item in rest.split(","):
                m = re.match(r"^\s*(\w+)\s*(?:\[(\d+)\])?\s*$", item)
                assert m, item
                fields.append((m.group(1), ty, int(m.group(2) or 1)))
        out[name] = fields
    return out


def header_enums():
    src = strip_comments(open(HDR).read())
    vals = {}
    for body in re.findall(r"enum\s*\{([^}]*)\}", src):
        nxt = 0
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                k, v = [x.strip() for x in item.split("=")]
                nxt = int(v, 0)
            else:
                k = item
            vals[k] = nxt
            nxt += 1
    return vals


# ---------------------------------------------------------------------------------------------------- C# side
def csharp_imports():
    src = strip_comments(open(NATIVE).read(), True)
    out = {}
    for ret, name, args in re.findall(r"\[DllImport\(Lib\)\]\s*public\s+static\s+extern\s+(\w+\*?)\s+(\w+)\s*\(([^;]*)\)\s*;", src):
        params = []
        for a in re.sub(r"\[MarshalAs\([^\]]*\)\]", "", args).split(","):
            a = a.strip()
            if not a:
                continue
            toks = a.split()
            params.append(" ".join(toks[:-1]))
        assert name not in out, "duplicate DllImport " + name
        out[name] = (ret, params)
    return out


def csharp_structs():
    src = strip_comments(open(NATIVE).read(), True)
    out = {}
    for name, body in re.findall(r"public\s+struct\s+(\w+)\s*\{([^}]*)\}", src):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"^public\s+(fixed\s+)?(\w+)\s+(.*)$", decl, re.S)
            assert m, decl
            for item in m.group(3).split(","):
                mm = re.match(r"^\s*(\w+)\s*(?:\[(\d+)\])?\s*$", item)
                assert mm, item
                fields.append((mm.group(1), m.group(2), int(mm.group(2) or 1)))
        out[name] = fields
    return out


# C parameter type -> the C# spellings that marshal to it on a 64-bit player (blittable, no copies)
COMPAT = {
    "int": {"int"}, "float": {"float"}, "int64_t": {"long"},
    "sph_handle": {"IntPtr"}, "sph_group": {"IntPtr"},
    "sph_handle*": {"out IntPtr"}, "sph_group*": {"out IntPtr"}, "void**": {"out IntPtr"},
    "void*": {"void*"}, "char*": {"string"}, "char**": {"IntPtr*"},
    "int64_t*": {"long*"}, "int32_t*": {"int*"}, "int*": {"int*"}, "uint32_t*": {"uint*"},
    "float*": {"float*"}, "double*": {"double*"},
    "sph_Params*": {"Params*"}, "sph_GridParams*": {"GridParams*"}, "sph_GroupInfo*": {"GroupInfo*"},
    "sph_ParticleInteraction*": {"ParticleInteraction*"},
}
RET = {"int": "int", "char*": "IntPtr"}
ELEM = {"float": "float", "int32_t": "int", "int64_t": "long", "uint32_t": "uint"}


def test_every_export_has_one_dllimport_with_the_same_signature():
    hdr, cs = header_functions(), csharp_imports()
    assert len(hdr) >= 50
    assert sorted(cs) == sorted(hdr), (sorted(set(hdr) - set(cs)), sorted(set(cs) - set(hdr)))
    for name, (ret, params) in hdr.items():
        cret, cparams = cs[name]
        assert RET[ret] == cret, (name, ret, cret)
        assert len(params) == len(cparams), (name, params, cparams)
        for k, (a, b) in enumerate(zip(params, cparams)):
            assert a in COMPAT, (name, a)
            assert b in COMPAT[a], "%s parameter %d: header %r, C# %r" % (name, k, a, b)


def test_python_cpp_and_csharp_bind_the_same_symbols():
    import sphb200
    assert sorted(sphb200.EXPORTS) == sorted(csharp_imports())
    hpp = open(os.path.join(ROOT, "planetmodel-sph_b200", "host_cpp", "sph_systems.hpp")).read()
    assert '#include "sphb200.h"' in hpp or "#include <sphb200.h>" in hpp or "sphb200.h" in hpp
    for used in set(re.findall(r"\b(sphb200_\w+)\s*\(", hpp)):
        assert used in header_functions(), used


def test_mirrored_structs_have_the_header_fields_in_order():
    hs, cs = header_structs(), csharp_structs()
    pairs = {"Params": "sph_Params", "GridParams": "sph_GridParams", "GroupInfo": "sph_GroupInfo",
             "ParticleInteraction": "sph_ParticleInteraction"}
    assert sorted(cs) == sorted(pairs)
    for cname, hname in pairs.items():
        hf, cf = hs[hname], cs[cname]
        assert [f[0] for f in hf] == [f[0] for f in cf], (cname, hf, cf)
        for (n, ht, hl), (_, ct, cl) in zip(hf, cf):
            assert ELEM[ht] == ct and hl == cl, (cname, n, ht, hl, ct, cl)
    # byte sizes implied by the C# field lists == the sizes gcc gives the header (test_abi_exports.py prints the same numbers)
    size = {"float": 4, "int": 4, "uint": 4, "long": 8}
    tot = {c: sum(size[t] * l for _, t, l in f) for c, f in cs.items()}
    assert tot["Params"] == 48 and tot["GridParams"] == 40 and tot["ParticleInteraction"] == 40
    assert tot["GroupInfo"] == 16 + 8 * 8 + 2 * 32 * 8      # no padding: the 4 ints fill 16 bytes before the first long


def test_constants_have_the_header_values():
    en = header_enums()
    src = strip_comments(open(NATIVE).read(), True)
    consts = {}
    for body in re.findall(r"public\s+const\s+int\s+([^;]*);", src):
        for item in body.split(","):
            k, v = [x.strip() for x in item.split("=")]
            consts[k] = int(v)
    want = [k for k in en if k.startswith(("SPH_OK", "SPH_ERR_", "SPH_FLAG_", "SPH_GRAVITY_"))]
    assert len(want) >= 14
    for k in want:
        assert consts.get(k) == en[k], (k, consts.get(k), en[k])
    m = re.search(r"public\s+enum\s+Field\s*\{([^}]*)\}", src)
    fld = {}
    for item in m.group(1).split(","):
        k, v = [x.strip() for x in item.split("=")]
        fld[k.lower()] = int(v)
    for k, v in en.items():
        if k.startswith("SPH_FIELD_") and not k.endswith("COUNT_"):
            assert fld[k[len("SPH_FIELD_"):].replace("_", "").lower()] == v, k


def test_system_classes_keep_the_reference_names_order_and_constants():
    src = strip_comments(open(SYSTEMS).read(), True)
    order = ["ParticleSmoothingSystem", "KernelSystem", "GravityFieldSystem", "DensityFieldSystem", "PressureFieldSystem",
             "VelocitySystem"]
    call = {"ParticleSmoothingSystem": "sphb200_smoothing_update", "KernelSystem": "sphb200_build_neighbors",
            "GravityFieldSystem": "sphb200_gravity", "DensityFieldSystem": "sphb200_density",
            "PressureFieldSystem": "sphb200_pressure", "VelocitySystem": "sphb200_integrate"}
    blocks = re.split(r"(?=\[UpdateInGroup\(typeof\(FixedStepSimulationSystemGroup\)\)\])", src)[1:]
    assert len(blocks) == 6
    for k, (name, blk) in enumerate(zip(order, blocks)):
        assert re.search(r"public\s+class\s+%s\s*:\s*SphSystemBase" % name, blk), name
        if k:
            assert "[UpdateAfter(typeof(%s))]" % order[k - 1] in blk, name
        used = set(re.findall(r"SphB200Native\.(sphb200_\w+)", blk))
        assert used == {call[name]}, (name, used)
    assert re.search(r"TARGET_NEIGHBORS\s*=\s*50\b", src)                       # ParticleSmoothingSystem.cs:18
    assert re.search(r"k_GravConstant\s*=\s*1\.0f", src)                        # GravityFieldSystem.cs:26
    assert re.search(r"k_Theta\s*=\s*0\.7f", src)                               # GravityFieldSystem.cs:228
    assert re.search(r"enum\s+GravityImpl\s*:\s*ushort\s*\{\s*GRAVITY_TREE_CPU\s*,\s*GRAVITY_PARTICLE_CPU\s*\}", src)
    assert "IPhysicsSystem" in src and "GetOutputDependency" in src and "AddInputDependency" in src
    for used in set(re.findall(r"SphB200Native\.(sphb200_\w+)", src)):
        assert used in csharp_imports(), used


def test_csharp_sources_are_lexically_well_formed():
    """No compiler here: at least every bracket closes, no string or comment is left open, and every P/Invoke line ends in ';'."""
    for path in (NATIVE, SYSTEMS):
        raw = re.sub(r"//[^\n]*", " ", open(path).read())        # line comments first: one of them mentions "Systems/*.cs"
        assert raw.count("/*") == raw.count("*/")
        src = re.sub(r"/\*.*?\*/", " ", raw, flags=re.S)
        src = re.sub(r'"(?:\\.|[^"\\\n])*"', '""', src)          # string literals out of the way
        assert src.count('"') % 2 == 0, path
        stack = []
        pairs = {")": "(", "]": "[", "}": "{"}
        for ln, line in enumerate(src.splitlines(), 1):
            for ch in line:
                if ch in "([{":
                    stack.append((ch, ln))
                elif ch in ")]}":
                    assert stack and stack[-1][0] == pairs[ch], "%s:%d unbalanced %r" % (path, ln, ch)
                    stack.pop()
        assert not stack, "%s: unclosed %r" % (path, stack[-1])
    decls = re.findall(r"\[DllImport\(Lib\)\][^;]*;", strip_comments(open(NATIVE).read(), True))
    assert len(decls) == len(csharp_imports())
