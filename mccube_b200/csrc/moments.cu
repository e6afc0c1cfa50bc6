// Moment reductions (sm_100a): weighted column sums and centred second moments in fp64.
//
// Replaces jnp.average (center_of_mass, /root/reference/mccube/_metrics.py:96), and the
// user-side jnp.mean / jnp.cov of /root/reference/README.md:85-86 that feed
// gaussian_wasserstein_metric (_metrics.py:39-44).  One streaming pass over [n, d] each;
// per-warp/CTA partials are flushed with fp64 atomics so the cross-GPU form is a plain
// all-reduce of d + 1 + d*d doubles.
#include "common.cuh"

namespace mccb {

constexpr int MS_THREADS = 256;
constexpr int MS_MAX_ACC = 32;  // d <= 1024

// One warp per row at a time; lane owns columns lane, lane+32, ...
__global__ void __launch_bounds__(MS_THREADS)
moments_sum_kernel(const float* __restrict__ p, uint64_t n, int d, int ld, const float* __restrict__ w,
                   int w_stride, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * MS_THREADS + threadIdx.x) >> 5;
  const uint64_t n_warps = ((uint64_t)gridDim.x * MS_THREADS) >> 5;
  double acc[MS_MAX_ACC];
#pragma unroll
  for (int i = 0; i < MS_MAX_ACC; ++i) acc[i] = 0.0;
  double wsum = 0.0;
  for (uint64_t r = warp; r < n; r += n_warps) {
    const float* row = p + r * ld;
    const float wr = w ? w[r * (uint64_t)w_stride] : 1.0f;
    if (lane == 0) wsum += (double)wr;
#pragma unroll
    for (int i = 0; i < MS_MAX_ACC; ++i) {
      int c = lane + 32 * i;
      if (c < d) acc[i] += (double)wr * (double)row[c];
    }
  }
  __shared__ double s_acc[MS_MAX_ACC * 32 + 1];
  for (int i = threadIdx.x; i < MS_MAX_ACC * 32 + 1; i += MS_THREADS) s_acc[i] = 0.0;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < MS_MAX_ACC; ++i) {
    int c = lane + 32 * i;
    if (c < d) atomicAdd(&s_acc[c], acc[i]);
  }
  if (lane == 0) atomicAdd(&s_acc[MS_MAX_ACC * 32], wsum);
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += MS_THREADS) atomicAdd(out + c, s_acc[c]);
  if (threadIdx.x == 0) atomicAdd(out + d, s_acc[MS_MAX_ACC * 32]);
}

// 64x64 output tile of sum_r w_r (p_r - c)(p_r - c)^T per CTA, rows staged through smem.
constexpr int M2_TILE = 64;
constexpr int M2_ROWS = 32;  // rows per smem stage

__global__ void __launch_bounds__(256)
moments_m2_kernel(const float* __restrict__ p, uint64_t n, int d, int ld, const float* __restrict__ w, int w_stride,
                  const float* __restrict__ centre, uint64_t rows_per_cta, double* __restrict__ out) {
  __shared__ float xi[M2_ROWS][M2_TILE + 1];
  __shared__ float xj[M2_ROWS][M2_TILE + 1];
  const int tiles = (d + M2_TILE - 1) / M2_TILE;
  const int ti = blockIdx.y / tiles, tj = blockIdx.y % tiles;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const uint64_t r0 = (uint64_t)blockIdx.x * rows_per_cta;
  const uint64_t r1 = min(n, r0 + rows_per_cta);
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
  double dacc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) dacc[a][b] = 0.0;

  int flush = 0;
  for (uint64_t rb = r0; rb < r1; rb += M2_ROWS) {
    __syncthreads();
    for (int e = threadIdx.x; e < M2_ROWS * M2_TILE; e += 256) {
      const int rr = e / M2_TILE, cc = e % M2_TILE;
      const uint64_t r = rb + rr;
      float vi = 0.0f, vj = 0.0f;
      if (r < r1) {
        const float wr = w ? w[r * (uint64_t)w_stride] : 1.0f;
        const int ci = ti * M2_TILE + cc, cj = tj * M2_TILE + cc;
        if (ci < d) vi = wr * (p[r * ld + ci] - centre[ci]);
        if (cj < d) vj = p[r * ld + cj] - centre[cj];
      }
      xi[rr][cc] = vi;
      xj[rr][cc] = vj;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < M2_ROWS; ++rr) {
      float a4[4], b4[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) a4[a] = xi[rr][ty * 4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) b4[b] = xj[rr][tx * 4 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a], b4[b], acc[a][b]);
    }
    if (++flush == 32) {  // 1024 rows of fp32 accumulation, then fold into fp64
      flush = 0;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { dacc[a][b] += (double)acc[a][b]; acc[a][b] = 0.0f; }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int ci = ti * M2_TILE + ty * 4 + a, cj = tj * M2_TILE + tx * 4 + b;
      if (ci < d && cj < d) atomicAdd(out + (size_t)ci * d + cj, dacc[a][b] + (double)acc[a][b]);
    }
}

// f = -(y - mean) @ prec, fp32 SIMT tiles: 64 rows x 64 cols per CTA, K in steps of 16.
constexpr int GD_BM = 64, GD_BN = 64, GD_BK = 16;

__global__ void __launch_bounds__(256)
gaussian_full_drift_kernel(const float* __restrict__ y, uint64_t n, int d, int ld, const float* __restrict__ mean,
                           const float* __restrict__ prec, float* __restrict__ f) {
  __shared__ float As[GD_BK][GD_BM + 4];
  __shared__ float Bs[GD_BK][GD_BN + 4];
  const uint64_t row0 = (uint64_t)blockIdx.x * GD_BM;
  const int col0 = blockIdx.y * GD_BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
  for (int k0 = 0; k0 < d; k0 += GD_BK) {
    for (int e = threadIdx.x; e < GD_BM * GD_BK; e += 256) {
      const int rr = e / GD_BK, kk = e % GD_BK;
      const uint64_t r = row0 + rr;
      const int c = k0 + kk;
      As[kk][rr] = (r < n && c < d) ? (y[r * ld + c] - mean[c]) : 0.0f;
    }
    for (int e = threadIdx.x; e < GD_BK * GD_BN; e += 256) {
      const int kk = e / GD_BN, cc = e % GD_BN;
      const int kr = k0 + kk, c = col0 + cc;
      Bs[kk][cc] = (kr < d && c < d) ? prec[(size_t)kr * d + c] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GD_BK; ++kk) {
      float a4[4], b4[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) a4[a] = As[kk][ty * 4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) b4[b] = Bs[kk][tx * 4 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a], b4[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const uint64_t r = row0 + ty * 4 + a;
    if (r >= n) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int c = col0 + tx * 4 + b;
      if (c < d) f[r * d + c] = -acc[a][b];
    }
  }
}

}  // namespace mccb

using namespace mccb;

extern "C" int mccb200_moments_sum(const float* p, uint64_t n, int d, int ld, const float* w, int w_stride,
                                   double* sum_out, void* stream) {
  MCCB_REQUIRE(p && sum_out && d >= 1 && ld >= d, "moments_sum: bad argument");
  MCCB_REQUIRE(d <= MS_MAX_ACC * 32, "moments_sum: d=%d > %d unsupported", d, MS_MAX_ACC * 32);
  if (n == 0) return MCCB200_OK;
  int grid = (int)min((n + 7) / 8, (uint64_t)sm_count() * 8);
  moments_sum_kernel<<<grid, MS_THREADS, 0, (cudaStream_t)stream>>>(p, n, d, ld, w, w_stride, sum_out);
  count_launches(1);
  return check_launch("moments_sum");
}

extern "C" int mccb200_moments_m2(const float* p, uint64_t n, int d, int ld, const float* w, int w_stride,
                                  const float* centre, double* m2_out, void* stream) {
  MCCB_REQUIRE(p && centre && m2_out && d >= 1 && ld >= d, "moments_m2: bad argument");
  if (n == 0) return MCCB200_OK;
  const int tiles = (d + M2_TILE - 1) / M2_TILE;
  uint64_t want_ctas = (uint64_t)sm_count() * 8 / (uint64_t)(tiles * tiles) + 1;
  uint64_t rows_per_cta = (n + want_ctas - 1) / want_ctas;
  rows_per_cta = (rows_per_cta + M2_ROWS - 1) / M2_ROWS * M2_ROWS;
  dim3 grid((unsigned)((n + rows_per_cta - 1) / rows_per_cta), (unsigned)(tiles * tiles));
  moments_m2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, n, d, ld, w, w_stride, centre, rows_per_cta, m2_out);
  count_launches(1);
  return check_launch("moments_m2");
}

extern "C" int mccb200_gaussian_full_drift(const float* y, uint64_t n, int d, int ld, const float* mean,
                                           const float* prec, float* f, void* stream) {
  MCCB_REQUIRE(y && mean && prec && f && d >= 1 && ld >= d, "gaussian_full_drift: bad argument");
  if (n == 0) return MCCB200_OK;
  MCCB_REQUIRE((n + GD_BM - 1) / GD_BM < 0x7FFFFFFFull, "gaussian_full_drift: n too large");
  dim3 grid((unsigned)((n + GD_BM - 1) / GD_BM), (unsigned)((d + GD_BN - 1) / GD_BN));
  gaussian_full_drift_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, n, d, ld, mean, prec, f);
  count_launches(1);
  return check_launch("gaussian_full_drift");
}
