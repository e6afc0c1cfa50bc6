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
  const uint32_t* perm;
  uint32_t out_per_part, in_per_part;
  uint64_t n_out;
  float* y1;
  int vpr;            // vectors per row = d / VEC
  int rows_per_iter;  // rows covered by one CTA pass = blockDim / vpr
};

int fill_expand_args(ExpandArgs& a, const float* y0, uint64_t n, int d, int weighted, const mccb200_drift* drift,
                     float dt, const mccb200_cubature* cub);

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

}  // namespace mccb
