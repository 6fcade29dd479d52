// group.cuh -- internal interface between group.cu (host side of the Morton-range decomposition) and kernels_group.cu.
#pragma once
#include "ctx.cuh"

#define SPH_BIN_BITS 18                  // ownership is decided on 2^18 key-prefix bins (never finer than the grid cells)
#define SPH_NBINS (1 << SPH_BIN_BITS)

enum { GR_U32 = 0, GR_U64 = 1, GR_F64 = 2 };
enum { GR_SUM = 0, GR_MIN = 1, GR_MAX = 2 };

int grk_bin_hist(sphb200_ctx* c, const uint32_t* keys, const int32_t* ncount, const int32_t* npart, const int32_t* napprox, int n, uint32_t* hist);
int grk_splitters(sphb200_ctx* c, const uint32_t* hist, int world, int64_t* split, unsigned long long* part, uint32_t* bmask);
int grk_dest(sphb200_ctx* c, const uint32_t* keys, int n, const int64_t* split, int world, uint8_t* dest);
int grk_mig_pack(sphb200_ctx* c, const float4* posh, const float4* velm, const uint32_t* orig, const int32_t* nown, const uint32_t* keys,
                 const uint32_t* perm, int n, uint4* rec);
int grk_mig_keys(sphb200_ctx* c, const uint4* rec, int n, uint32_t* keys);
int grk_halo_lists(sphb200_ctx* c, const uint32_t* keys, int n, const int64_t* split, int world, int me, const uint32_t* bmask, uint32_t* mask, uint32_t* cnt,
                   uint32_t* total, uint32_t* list);
size_t grk_halo_cnt_words(int64_t cap);
int grk_halo_pack(sphb200_ctx* c, const uint4* rec, const uint32_t* idx, const uint32_t* keys, const uint32_t* list, int n, uint4* out);
int grk_assemble_ext(sphb200_ctx* c, const uint4* rec, const uint32_t* idx, const uint32_t* keys_own, const uint4* halo, int low, int nown,
                     int next, uint32_t* keys_ext, size_t ncell);
int grk_sorted_sources(sphb200_ctx* c, const uint4* rec, const uint32_t* idx, const uint32_t* keys, int n, float4* posm, uint32_t* keys_out,
                       float4* tposh, float4* tvelm);
int grk_gather_f32(sphb200_ctx* c, const float* src, const uint32_t* list, int n, float* out);
int grk_boundary(sphb200_ctx* c, const float4* posh, const float4* velm, int n, float4* bnd);
int grk_let_box(sphb200_ctx* c, const float4* posm, int lo, int hi, uint32_t* box);
int grk_let_mask(sphb200_ctx* c, const uint32_t* boxes, int world, int me, uint32_t* mask, uint32_t* cnt);
int grk_let_pack(sphb200_ctx* c, const uint32_t* mask, int world, const uint32_t* soff, uint32_t* cursor, float4* out);
int grk_let_scatter(sphb200_ctx* c, const float4* rec, int64_t nrec);
int grk_body_dest(sphb200_ctx* c, const uint32_t* orig, int n, int64_t chunk, uint8_t* dest);
int grk_result_pack(sphb200_ctx* c, const uint32_t* perm, int n, int own0, float4* rec);
int grk_result_field(sphb200_ctx* c, const float4* rec, int n, int field, int64_t body0, float* out);
int grk_mass_range(sphb200_ctx* c, const float4* velm, int n, uint32_t* mm);
int grk_reduce_ranks(sphb200_ctx* c, const void* scratch, int world, size_t count, int dtype, int op, void* buf);
