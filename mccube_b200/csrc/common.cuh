// Shared helpers for the mccube_b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mccube_b200.h"

namespace mccb {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);
int sm_count();
int env_int(const char* name, int dflt);
int side_ctas();   // > 0: CTAs per SM for the index-draw / bucketing kernels launched by this thread (see mccb200_set_side_residency)   // experiment knobs (read once per call site)
void count_launches(int n);
void phase_mark(const char* name, cudaStream_t st);  // no-op unless mccb200_profile_enable(1)  // kernels / device copies enqueued by this library (bench.py gpu_launches)

#define MCCB_REQUIRE(cond, ...)                                        \
  do {                                                                 \
    if (!(cond)) return ::mccb::fail(MCCB200_ERR_INVALID_ARGUMENT, __VA_ARGS__); \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------- threefry2x32-20
// Same arithmetic as jax/_src/prng.py `_threefry2x32_lowering` (Random123 Threefry-2x32,
// 20 rounds); host+device so the step-key derivation can run on either side.
struct u32x2 {
  uint32_t x, y;
};

__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t v, int r) {
#ifdef __CUDA_ARCH__
  return __funnelshift_l(v, v, r);
#else
  return (v << r) | (v >> (32 - r));
#endif
}

__host__ __device__ __forceinline__ u32x2 threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += k0;
  x1 += k1;
#define MCCB_TF_R(r) \
  x0 += x1;          \
  x1 = rotl32(x1, r); \
  x1 ^= x0;
  MCCB_TF_R(13) MCCB_TF_R(15) MCCB_TF_R(26) MCCB_TF_R(6)
  x0 += k1; x1 += k2 + 1u;
  MCCB_TF_R(17) MCCB_TF_R(29) MCCB_TF_R(16) MCCB_TF_R(24)
  x0 += k2; x1 += k0 + 2u;
  MCCB_TF_R(13) MCCB_TF_R(15) MCCB_TF_R(26) MCCB_TF_R(6)
  x0 += k0; x1 += k1 + 3u;
  MCCB_TF_R(17) MCCB_TF_R(29) MCCB_TF_R(16) MCCB_TF_R(24)
  x0 += k1; x1 += k2 + 4u;
  MCCB_TF_R(13) MCCB_TF_R(15) MCCB_TF_R(26) MCCB_TF_R(6)
  x0 += k2; x1 += k0 + 5u;
#undef MCCB_TF_R
  return {x0, x1};
}

// jr.split(key, num)[i]  (both jax_threefry_partitionable modes), for small num.
//   original: hash(key, iota(2*num)).reshape(num, 2): flat element e of the hash is word
//             (e >= num) of TF(key; e mod num, e mod num + num).
//   partitionable: TF(key; 0, i).
__host__ __device__ inline u32x2 split_key(u32x2 key, uint32_t num, uint32_t i, bool partitionable) {
  if (partitionable) return threefry2x32(key.x, key.y, 0u, i);
  uint32_t e0 = 2u * i, e1 = 2u * i + 1u;
  u32x2 a = threefry2x32(key.x, key.y, e0 % num, e0 % num + num);
  u32x2 b = threefry2x32(key.x, key.y, e1 % num, e1 % num + num);
  return {e0 >= num ? a.y : a.x, e1 >= num ? b.y : b.x};
}

// jr.fold_in(key, data)
__host__ __device__ inline u32x2 fold_in_key(u32x2 key, uint32_t data) {
  return threefry2x32(key.x, key.y, 0u, data);
}

// random_bits(key, 32, (N,))[i]: one element (original mode wastes half a block).
__device__ __forceinline__ uint32_t random_bits32_at(u32x2 key, uint32_t i, uint32_t n, bool partitionable) {
  if (partitionable) {
    u32x2 o = threefry2x32(key.x, key.y, 0u, i);
    return o.x ^ o.y;
  }
  const uint32_t h = (n + 1u) >> 1;  // padded half length
  if (i < h) {
    uint32_t hi = i + h;
    u32x2 o = threefry2x32(key.x, key.y, i, hi < n ? hi : 0u);  // pad element counts as 0
    return o.x;
  }
  u32x2 o = threefry2x32(key.x, key.y, i - h, i);
  return o.y;
}

// ---------------------------------------------------------------- Hadamard signs
// M[j, c] of mccube/_formulae.py:217-221 as a sign bit: 1 -> -1.
__device__ __forceinline__ uint32_t hadamard_neg(uint32_t j, uint32_t half_mask, uint32_t c) {
  // rows j >= k/2 are the negated block: (j & ~half_mask) != 0  <=>  bit log2(k/2) set.
  uint32_t par = __popc(j & half_mask & c) & 1u;
  uint32_t blk = (j > half_mask) ? 1u : 0u;
  return par ^ blk;
}

}  // namespace mccb
