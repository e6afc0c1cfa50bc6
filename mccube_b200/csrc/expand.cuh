// Shared description of one Euler/cubature expansion step (used by every kernel that
// needs child coordinates: expand_select, box child keys, fused box ranking).
#pragma once
#include "common.cuh"

namespace mccb {

struct ExpandArgs {
  const float* y0;
  uint64_t n;
  int d, ld, weighted;
  int drift_kind;
  const float* f;
  const float* mean;
  const float* inv_var;
  float dt;
  uint32_t k, k_shift, half_mask;  // k_shift valid when k is a power of two (else 0xFFFFFFFF)
  const float* points;
  float scale;
  const float* lam;
  const uint32_t* idx;
  uint32_t idx_stride;   // idx entries are idx_stride words apart (2: the id column of packed (slot, id) pairs); 0 = 1
  const uint32_t* perm;
  uint32_t out_per_part, in_per_part;
  uint64_t n_out;
  float* y1;
  // n_substeps > 1 (Hadamard, drift kinds 0 / 2 only): child c = ((i*k + j_1)*k + j_2)... ; substep s
  // adds dt_sub[s] * f(y_s) + sign(j_s) * scale_sub[s] to the running state (MCCSolver.step loop,
  // mccube/_solvers.py:132-136).  n_sub == 1 uses dt / scale above.
  int n_sub;
  float dt_sub[MCCB200_MAX_SUBSTEPS], scale_sub[MCCB200_MAX_SUBSTEPS];
  // Multi-GPU scatter form (expand_scatter_peers): output row q of THIS rank is slot
  // sc_order[start + q] of the global result, start / count read on the device from the replicated
  // counts matrix; the child is written straight into the shard of the rank that stores the slot
  // (peer memory over NVLink), no send buffer, no collective.
  const uint64_t* idx64;      // pull form with 64-bit GLOBAL child ids (n_tot * k may exceed 2^32), else nullptr
  const uint32_t* sc_order;   // [n_tot] output slots grouped by owner rank (ascending slot inside a group)
  const uint32_t* sc_counts;  // [G][G] counts[owner][dest]
  float* const* sc_peers;     // [G] base pointer of every rank's output shard
  uint32_t sc_rank, sc_G;
  uint64_t sc_rows_per_rank, sc_child_offset;
  const float* keycols;   // optional [n, 4]: y0[:, c0 .. c0+3] of the 4 vectorised box key columns, packed
  float* keycols_out;     // optional [n_out, 4]: the same columns of y1, written by the expand kernel
  int keycol0;            // c0 of keycols_out
  // Fused index draw (MonteCarloKernel(with_replacement=True), 32-bit randint): the child of output row r is
  // jr.randint(k, (draw_count,), 0, draw_span)[draw_first + r], computed in the kernel -- no index vector in HBM.
  const uint32_t* draw_keys;   // device: randint_hi.xy, randint_lo.xy of the step's DerivedKeys (nullptr = not fused)
  uint32_t draw_count, draw_span, draw_first;
  int draw_partitionable;
  uint32_t* draw_idx_out;      // optional: the drawn ids, [n_out]
  int vpr_shift;               // log2(vpr) when the draw is fused (vpr a power of two, 4 .. 32)
  double* wsum;                // optional (weighted): receives the sum of the child weights written (atomic add)
  int vpr;            // vectors per row = d / VEC
  int rows_per_iter;  // rows covered by one CTA pass = blockDim / vpr
};

int fill_expand_args(ExpandArgs& a, const float* y0, uint64_t n, int d, int weighted, const mccb200_drift* drift,
                     float dt, const mccb200_cubature* cub);

// Loads of a few columns out of every (long) row: only one 64-byte line per row is wanted, so
// ask L2 for a 64-byte fill instead of its default 128-byte one.  Measured on B200 (2^24 rows of
// 256 B, 16 B used per row; experiments/strided_read.cu): 0.175 ms with the hint, 0.30 ms without;
// cudaLimitMaxL2FetchGranularity has no effect.
__device__ __forceinline__ float ld_strided(const float* p) {
  float v;
  asm("ld.global.nc.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_strided4(const float* p) {  // p 16-byte aligned
  float4 v;
  asm("ld.global.nc.L1::no_allocate.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "l"(p));
  return v;
}

// Streaming 128-bit row accesses (read once / written once: no L1 allocation).
__device__ __forceinline__ float4 ld_stream4(const float* p) {  // p 16-byte aligned
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream4(float* p, const float4& v) {  // p 16-byte aligned
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// Drift value f_col(y0[i, :]) of parent i (evaluated on PARENTS only, _term.py:58-63).
__device__ __forceinline__ float drift_value(const ExpandArgs& a, uint64_t i, int col, float y) {
  if (a.drift_kind == 2) return __fmul_rn(__fsub_rn(__ldg(a.mean + col), y), __ldg(a.inv_var + col));
  if (a.drift_kind == 1) return __ldg(a.f + i * a.d + col);
  return 0.0f;
}

// Noise increment of cubature point j in column col: g * (sqrt(dt) * M[j, col]).
__device__ __forceinline__ float noise_value(const ExpandArgs& a, uint32_t j, int col) {
  if (a.points) return __ldg(a.points + (uint64_t)j * a.d + col);
  return hadamard_neg(j, a.half_mask, (uint32_t)col) ? -a.scale : a.scale;
}

// Child coordinate in the reference's fp32 op order: y + (dt * f + noise).
__device__ __forceinline__ float child_value(const ExpandArgs& a, float y, float fv, float noise) {
  return __fadd_rn(y, __fadd_rn(__fmul_rn(a.dt, fv), noise));
}

// Key dimensions / resolution of BoxPartitionKernel.
struct BoxSpec {
  int n_dims, bits;
  int dims[MCCB200_BOX_MAX_DIMS];
};
int make_box_spec(BoxSpec& s, const int32_t* dims_host, int n_dims, int bits, int d);

// Key dims are 4 consecutive, 16-byte aligned columns (the default dims 0..3): one 128-bit load
// per parent instead of four scalar ones.
inline bool box_dims_vec4(const BoxSpec& s, const ExpandArgs& a) {
  if (s.n_dims != 4 || (s.dims[0] & 3) || (a.ld & 3)) return false;
  for (int q = 1; q < 4; ++q)
    if (s.dims[q] != s.dims[0] + q) return false;
  if (((uintptr_t)a.y0) & 15) return false;
  if (a.drift_kind == 1 && ((a.d & 3) || (((uintptr_t)a.f) & 15))) return false;
  return true;
}

// Per-thread constants of the 4 vectorised key dims (analytic drift): loaded once, not per parent.
struct KeyConst4 { float mean[4], ivar[4]; };

__device__ __forceinline__ KeyConst4 key_const4(const ExpandArgs& a, const BoxSpec& s) {
  KeyConst4 kc;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    kc.mean[q] = a.drift_kind == 2 ? __ldg(a.mean + s.dims[0] + q) : 0.0f;
    kc.ivar[q] = a.drift_kind == 2 ? __ldg(a.inv_var + s.dims[0] + q) : 0.0f;
  }
  return kc;
}

// the 4 key columns of parent i: packed hand-over of the previous step's gather (16 contiguous bytes per parent
// instead of a 64-byte DRAM fill per 256-byte row), else from the strided state
__device__ __forceinline__ float4 key_columns4(const ExpandArgs& a, const BoxSpec& s, uint64_t i) {
  return a.keycols ? __ldg(reinterpret_cast<const float4*>(a.keycols) + i) : ld_strided4(a.y0 + i * a.ld + s.dims[0]);
}

// The two candidate child coordinates (- and + noise) of parent i in its 4 vectorised key dims.
__device__ __forceinline__ void key_children4(const ExpandArgs& a, const BoxSpec& s, const KeyConst4& kc, uint64_t i,
                                              const float4 y4, float (&vm)[4], float (&vp)[4]) {
  const float y[4] = {y4.x, y4.y, y4.z, y4.w};
  float f[4] = {0.f, 0.f, 0.f, 0.f};
  if (a.drift_kind == 1) {
    const float4 f4 = ld_strided4(a.f + i * a.d + s.dims[0]);
    f[0] = f4.x; f[1] = f4.y; f[2] = f4.z; f[3] = f4.w;
  } else if (a.drift_kind == 2) {
#pragma unroll
    for (int q = 0; q < 4; ++q) f[q] = __fmul_rn(__fsub_rn(kc.mean[q], y[q]), kc.ivar[q]);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    vm[q] = child_value(a, y[q], f[q], -a.scale);
    vp[q] = child_value(a, y[q], f[q], a.scale);
  }
}

__device__ __forceinline__ void key_children4(const ExpandArgs& a, const BoxSpec& s, uint64_t i, float (&vm)[4],
                                              float (&vp)[4]) {
  key_children4(a, s, key_const4(a, s), i, key_columns4(a, s, i), vm, vp);
}

}  // namespace mccb
