"""Multi-GPU driver: one process per GPU, Morton-range ownership of *targets*, NCCL all-gather of the per-step state.

Design (DESIGN.md section "multi-GPU"): after the global Morton sort every rank holds the same sorted SoA arrays
(positions are needed by every rank anyway for gravity: the north-star's "all-gather of positions/masses").  Rank r owns
the contiguous sorted-slot range [r*chunk, (r+1)*chunk) -- a compact spatial domain on the Morton curve -- and computes
density / pressure force / gravity / integration only for those targets.  SPH halos need no separate exchange: the
neighbors of a boundary particle are read from the replicated arrays.  Per step three collectives run over NVLink:
  1. all-gather of (m/rho)P (4 B x N)        after the density pass   -- the "rho,P of halo particles" exchange
  2. all-gather of posh+velm (32 B x N)      after integration        -- positions/masses/h for the next step
  3. all-gather of own-support counts (4 B x N)                       -- input of the next smoothing-length update
All ranks run the O(N) sort / cell-table / LBVH-build kernels redundantly (identical inputs -> bit-identical order).

The partition arithmetic and the slice all-gather are backend-agnostic (tested on CPU with gloo, world_size 2).
"""
import numpy as np


def chunk_size(n, world):
    """Slots per rank (equal chunks; the last ranks may own fewer or zero real particles)."""
    return (int(n) + world - 1) // world


def shard_range(n, rank, world):
    """Owned sorted-slot range [t0, t1) of `rank`."""
    c = chunk_size(n, world)
    t0 = min(rank * c, n)
    t1 = min(t0 + c, n)
    return t0, t1


def padded_capacity(n, world):
    return chunk_size(n, world) * world


def allgather_slices(full, rank, world, group=None):
    """In-place all-gather: `full` is a tensor whose dim 0 is padded_capacity(n, world); every rank contributes its own
    chunk (already written in place) and receives everyone else's."""
    import torch.distributed as td
    c = full.shape[0] // world
    mine = full[rank * c:(rank + 1) * c]
    td.all_gather_into_tensor(full, mine, group=group)
    return full


class _DevArray:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface v3)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class ShardedSimulation:
    """world == 1: a plain Simulation.  world > 1: replicated state, partitioned targets (see module docstring)."""

    def __init__(self, n, device=0, rank=0, world=1, **params):
        from . import Simulation
        self.n = int(n)
        self.rank, self.world = rank, world
        self.cap = padded_capacity(n, world)
        self.sim = Simulation(self.cap, device=device, **params)
        self.device = device
        self._views = {}
        self.stream = None
        self.timing = False
        self._marks = []
        if world > 1:
            # kernels and NCCL collectives must be ordered on ONE stream: a dedicated torch stream whose handle the
            # library adopts (torch's legacy default stream has handle 0, which the C ABI reads as "own stream")
            import torch
            self.stream = torch.cuda.Stream(device=device)
            self.sim.set_stream(self.stream.cuda_stream)

    def upload(self, pos, vel, mass, h):
        self.sim.upload(pos, vel, mass, h)      # every rank uploads the full set (replicated state)
        self._views = {}

    # -- torch views of the library's arrays (float32 words), cached per pointer
    def _view(self, name, words_per_elem):
        import torch
        ptr, nbytes = self.sim.device_ptr(name)
        key = (name, ptr)
        if key not in self._views:
            t = torch.as_tensor(_DevArray(ptr, nbytes), device="cuda:%d" % self.device)
            self._views[key] = t.view(torch.float32).view(-1, words_per_elem)[: self.cap]
        return self._views[key]

    def step(self, dt, impl):
        s = self.sim
        if self.world == 1:
            s.step(dt, impl)
            return
        import torch
        with torch.cuda.stream(self.stream):
            self._step_sharded(dt, impl)

    def _mark(self, name):
        if self.timing:
            import torch
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(self.stream)
            self._marks.append((name, ev))

    def _step_sharded(self, dt, impl):
        from . import GRAVITY_TREE, GRAVITY_PARTICLE
        s = self.sim
        t0, t1 = shard_range(self.n, self.rank, self.world)
        self._marks = []
        self._mark("begin")
        s.set_target_range(t0, t1)
        s.smoothing_update()                    # all N (needs everyone's own-support counts: gathered last step)
        s.prepare_gravity(impl, dt)             # tree gravity: LBVH build overlaps the neighbor pass (auxiliary stream)
        s.build_neighbors()                     # global sort + cell table (redundant), lists + density for [t0,t1)
        self._mark("smoothing_sort_neighbors_density_eos")
        allgather_slices(self._view("cvol", 1), self.rank, self.world)
        self._mark("allgather_cvol")
        if impl == GRAVITY_TREE:
            # the pressure gradient does not depend on gravity: running it first gives the LBVH build on the auxiliary
            # stream (all N on every rank, the critical path once the neighbor pass is 1/world) more time to finish
            s.pressure()
            self._mark("pressure_grad")
        s.gravity(impl, dt)                     # sources: all N (posm / LBVH are global), targets [t0,t1)
        self._mark({GRAVITY_TREE: "gravity_tree", GRAVITY_PARTICLE: "gravity_allpairs"}.get(impl, "gravity_none"))
        if impl != GRAVITY_TREE:
            s.pressure()
            self._mark("pressure_grad")
        s.integrate(dt)
        self._mark("integrate")
        allgather_slices(self._view("posh", 4), self.rank, self.world)
        allgather_slices(self._view("velm", 4), self.rank, self.world)
        allgather_slices(self._view("nown", 1), self.rank, self.world)
        self._mark("allgather_state")

    def enable_timing(self, on=True):
        self.timing = bool(on)
        if self.world == 1:
            self.sim.enable_timing(on)

    def timings(self):
        """[(pass, ms)] of the most recent step (synchronises that step)."""
        if self.world == 1:
            return self.sim.timings()
        if not self._marks:
            return []
        self._marks[-1][1].synchronize()
        return [(self._marks[i][0], self._marks[i - 1][1].elapsed_time(self._marks[i][1])) for i in range(1, len(self._marks))]

    def gather_results(self):
        """Make the per-step result fields complete on every rank (for downloads/diagnostics)."""
        if self.world == 1:
            return
        import torch
        with torch.cuda.stream(self.stream):
            for name, w in (("rho", 1), ("press", 1), ("gradp", 4), ("grav", 4), ("ncount", 1), ("npart", 1), ("napprox", 1)):
                allgather_slices(self._view(name, w), self.rank, self.world)
        self.stream.synchronize()
