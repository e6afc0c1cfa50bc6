// Weighted Monte Carlo sampling (sm_100a): the two float branches of jr.choice that
// MonteCarloKernel reaches when `weighting_function` returns probabilities
// (/root/reference/mccube/_kernels/random.py:60-65 -> jax/_src/random.py choice):
//   replace=False : g = -gumbel(key, (P,)) - log(p);  ind = argsort(g)[:count]    (Gumbel top-k)
//   replace=True  : c = cumsum(p); r = c[-1] * (1 - uniform(key, (count,))); ind = searchsorted(c, r)
// The random BITS are jax.random's exactly (threefry, same key chain); the float arithmetic
// (log, cumsum order) cannot be promised bit-identical to XLA on another backend (SURVEY.md
// H3), so parity for these branches is "same indices except where two scores differ by rounding".
#include "common.cuh"
#include "rng.cuh"

namespace mccb {

__device__ __forceinline__ float uniform01_from_bits(uint32_t bits) {
  return __fsub_rn(__uint_as_float((bits >> 9) | 0x3F800000u), 1.0f);
}

__device__ __forceinline__ uint32_t float_sort_key(float v) {
  const uint32_t b = __float_as_uint(v);
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

__global__ void __launch_bounds__(256)
gumbel_keys_kernel(const DerivedKeys* __restrict__ dk, const float* __restrict__ p, int p_stride, uint32_t n,
                   int partitionable, uint32_t* __restrict__ keys) {
  const u32x2 key = dk->step_key;
  const float tiny = 1.17549435e-38f;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float f = uniform01_from_bits(random_bits32_at(key, i, n, partitionable != 0));
    // uniform(minval=tiny, maxval=1): max(tiny, f * (1 - tiny) + tiny)
    const float u = fmaxf(tiny, __fadd_rn(__fmul_rn(f, __fsub_rn(1.0f, tiny)), tiny));
    const float gumbel = -logf(-logf(u));
    const float g = __fsub_rn(-gumbel, logf(p[(uint64_t)i * p_stride]));
    keys[i] = float_sort_key(g);
  }
}

// inclusive scan of p (fp64 accumulation, one fp32 rounding per output): chunk sums -> scan -> apply
constexpr int WS_CHUNK = 4096;

__global__ void __launch_bounds__(256)
wscan_reduce_kernel(const float* __restrict__ p, int p_stride, uint64_t n, double* __restrict__ sums) {
  __shared__ double red[8];
  const uint64_t base = (uint64_t)blockIdx.x * WS_CHUNK;
  double acc = 0.0;
  for (int i = threadIdx.x; i < WS_CHUNK; i += 256) {
    const uint64_t g = base + i;
    if (g < n) acc += (double)p[g * p_stride];
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < 8; ++w) s += red[w];
    sums[blockIdx.x] = s;
  }
}

__global__ void wscan_sums_kernel(double* __restrict__ sums, uint32_t n_sums) {
  if (threadIdx.x || blockIdx.x) return;
  double run = 0.0;
  for (uint32_t i = 0; i < n_sums; ++i) { const double v = sums[i]; sums[i] = run; run += v; }
}

__global__ void __launch_bounds__(256)
wscan_apply_kernel(const float* __restrict__ p, int p_stride, uint64_t n, const double* __restrict__ sums,
                   float* __restrict__ c) {
  __shared__ double wtot[8];
  const uint64_t base = (uint64_t)blockIdx.x * WS_CHUNK + (uint64_t)threadIdx.x * 16;
  double v[16], tsum = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint64_t g = base + i;
    v[i] = g < n ? (double)p[g * p_stride] : 0.0;
    tsum += v[i];
  }
  double x = tsum;
  for (int o = 1; o < 32; o <<= 1) {
    const double y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = x;
  __syncthreads();
  double woff = 0.0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += wtot[w];
  double run = sums[blockIdx.x] + woff + x - tsum;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint64_t g = base + i;
    run += v[i];
    if (g < n) c[g] = (float)run;
  }
}

__global__ void __launch_bounds__(256)
searchsorted_uniform_kernel(const DerivedKeys* __restrict__ dk, const float* __restrict__ c, uint32_t n,
                            uint32_t count, int partitionable, uint32_t* __restrict__ idx) {
  const u32x2 key = dk->step_key;
  const float total = c[n - 1];
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < count; q += gridDim.x * blockDim.x) {
    const float u = uniform01_from_bits(random_bits32_at(key, q, count, partitionable != 0));
    const float r = __fmul_rn(total, __fsub_rn(1.0f, u));
    uint32_t lo = 0, hi = n;              // first i with c[i] >= r  (jnp.searchsorted side='left')
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (c[mid] < r) lo = mid + 1; else hi = mid;
    }
    idx[q] = lo;
  }
}

}  // namespace mccb

using namespace mccb;

extern "C" size_t mccb200_weighted_workspace_bytes(uint64_t n_inputs) {
  const size_t chunks = (size_t)((n_inputs + WS_CHUNK - 1) / WS_CHUNK);
  return align_up(sizeof(DerivedKeys), 256) + align_up((size_t)n_inputs * 4, 256) + align_up(chunks * 8, 256);
}

extern "C" int mccb200_gumbel_keys(const mccb200_rng* rng, const float* p, int p_stride, uint64_t n_inputs,
                                   uint32_t* keys_out, void* workspace, size_t workspace_bytes, void* stream) {
  MCCB_REQUIRE(rng && p && keys_out && workspace && p_stride >= 1, "gumbel_keys: bad argument");
  MCCB_REQUIRE(n_inputs >= 1 && n_inputs < 0xFFFFFFFFull, "gumbel_keys: n_inputs out of range");
  if (workspace_bytes < mccb200_weighted_workspace_bytes(n_inputs))
    return fail(MCCB200_ERR_WORKSPACE_TOO_SMALL, "gumbel_keys: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  DerivedKeys* dk = (DerivedKeys*)workspace;
  launch_derive_keys(*rng, 0, dk, st);
  int grid = (int)min((n_inputs + 255) / 256, (uint64_t)sm_count() * 8);
  gumbel_keys_kernel<<<grid, 256, 0, st>>>(dk, p, p_stride, (uint32_t)n_inputs, rng->partitionable, keys_out);
  count_launches(2);
  return check_launch("gumbel_keys");
}

extern "C" int mccb200_weighted_choice_replace(const mccb200_rng* rng, const float* p, int p_stride,
                                               uint64_t n_inputs, uint64_t count, uint32_t* idx_out,
                                               void* workspace, size_t workspace_bytes, void* stream) {
  MCCB_REQUIRE(rng && p && idx_out && workspace && p_stride >= 1, "weighted_choice_replace: bad argument");
  MCCB_REQUIRE(n_inputs >= 1 && n_inputs < 0xFFFFFFFFull && count < 0xFFFFFFFFull, "weighted_choice_replace: size out of range");
  if (workspace_bytes < mccb200_weighted_workspace_bytes(n_inputs))
    return fail(MCCB200_ERR_WORKSPACE_TOO_SMALL, "weighted_choice_replace: workspace too small");
  if (count == 0) return MCCB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)workspace;
  DerivedKeys* dk = (DerivedKeys*)w;
  float* c = (float*)(w + align_up(sizeof(DerivedKeys), 256));
  double* sums = (double*)(w + align_up(sizeof(DerivedKeys), 256) + align_up((size_t)n_inputs * 4, 256));
  const uint32_t chunks = (uint32_t)((n_inputs + WS_CHUNK - 1) / WS_CHUNK);
  launch_derive_keys(*rng, 0, dk, st);
  wscan_reduce_kernel<<<chunks, 256, 0, st>>>(p, p_stride, n_inputs, sums);
  wscan_sums_kernel<<<1, 32, 0, st>>>(sums, chunks);
  wscan_apply_kernel<<<chunks, 256, 0, st>>>(p, p_stride, n_inputs, sums, c);
  int grid = (int)min((count + 255) / 256, (uint64_t)sm_count() * 8);
  searchsorted_uniform_kernel<<<grid, 256, 0, st>>>(dk, c, (uint32_t)n_inputs, (uint32_t)count, rng->partitionable, idx_out);
  count_launches(5);
  return check_launch("weighted_choice_replace");
}
