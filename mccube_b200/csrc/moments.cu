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

// Column sums for rows of whole float4s (d % 4 == 0, 16-byte aligned rows): thread = (row slot, column
// quad), 128-bit streaming loads, four rows in flight per thread, fp64 accumulators.  HBM-bound: one read
// of n*d*4 bytes (the warp-per-row form above, kept for ragged d, has too few bytes in flight: 0.5 TB/s).
constexpr int MS_UNROLL = 4;
__global__ void __launch_bounds__(MS_THREADS)
moments_sum_vec4_kernel(const float* __restrict__ p, uint64_t n, int d, int ld, const float* __restrict__ w,
                        int w_stride, double* __restrict__ out) {
  extern __shared__ double s_sum[];                     // [d + 1]
  const int vpr = d >> 2, rows_per_iter = MS_THREADS / vpr;
  const int slot = threadIdx.x / vpr, cq = threadIdx.x - slot * vpr;
  for (int i = threadIdx.x; i <= d; i += MS_THREADS) s_sum[i] = 0.0;
  __syncthreads();
  if (slot < rows_per_iter) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, wsum = 0.0;
    const uint64_t step = (uint64_t)gridDim.x * rows_per_iter;
    for (uint64_t r = (uint64_t)blockIdx.x * rows_per_iter + slot; r < n; r += MS_UNROLL * step) {
      float4 v[MS_UNROLL];
      float wr[MS_UNROLL];
#pragma unroll
      for (int u = 0; u < MS_UNROLL; ++u) {
        const uint64_t ru = r + (uint64_t)u * step;
        const bool ok = ru < n;
        v[u] = ok ? __ldcs(reinterpret_cast<const float4*>(p + ru * ld) + cq) : make_float4(0.f, 0.f, 0.f, 0.f);
        wr[u] = ok ? (w ? __ldg(w + ru * (uint64_t)w_stride) : 1.0f) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < MS_UNROLL; ++u) {
        const double wd = (double)wr[u];
        a0 += wd * (double)v[u].x; a1 += wd * (double)v[u].y; a2 += wd * (double)v[u].z; a3 += wd * (double)v[u].w;
        wsum += wd;
      }
    }
    atomicAdd(&s_sum[cq * 4 + 0], a0); atomicAdd(&s_sum[cq * 4 + 1], a1);
    atomicAdd(&s_sum[cq * 4 + 2], a2); atomicAdd(&s_sum[cq * 4 + 3], a3);
    if (cq == 0) atomicAdd(&s_sum[d], wsum);
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= d; i += MS_THREADS) atomicAdd(out + i, s_sum[i]);
}

// 64x64 output tile of sum_r w_r (p_r - c)(p_r - c)^T per CTA.  At d = 64 this is 32 flop per byte read, far
// above what the FP32 FMA pipe sustains from shared memory (a 4x4 block per thread is shared-memory bound at
// 26 TFLOP/s, an 8x8 block register bound), so the products run on the tensor cores as 3xTF32:
// x = hi + lo with hi = tf32(x), and hi*hi + hi*lo + lo*hi recovers fp32-level products (the dropped lo*lo
// term is 2^-22 relative).  The contraction index is the particle row, so both operands of the
// m16n8k8 `mma.sync` are read straight from one row-major stage in shared memory (pitch 72: conflict-free
// fragment loads); centring (and the row weight) is applied to the fragments.  Rows stream through a
// 4-stage `cp.async` ring of 32 rows (with register prefetch of one stage the kernel was latency bound:
// 8 warps per SM, tensor pipe 26 % busy).  Four warps, a 32x32 quadrant each; fp32 accumulation over at
// most 2048 rows, then one fp64 atomic per entry.  Only tiles on or above the diagonal are computed,
// off-diagonal ones are mirrored.
bool moments_m2_tc_applicable(const float* p, int d, int ld, const float* w, const float* centre);
int moments_m2_tc(const float* p, uint64_t n, int ld, const float* centre, double* m2_out, cudaStream_t st);

constexpr int M2_TILE = 64;
constexpr int M2_ROWS = 32;      // rows per smem stage
constexpr int M2_PITCH = 72;     // floats: bank = (8 * k + column) mod 32 is distinct over a fragment load
constexpr int M2_THREADS = 128;
constexpr int M2_STAGES = 4;
constexpr int M2_FLUSH = 2048 / M2_ROWS;
constexpr int M2_STAGE_FLOATS = M2_ROWS * M2_PITCH;

static size_t m2_smem_bytes(bool two_operands, bool weighted) {
  return ((size_t)M2_STAGES * M2_STAGE_FLOATS * (two_operands ? 2 : 1) + (weighted ? M2_STAGES * M2_ROWS : 0)) * sizeof(float);
}

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  lo = __float_as_uint(x - __uint_as_float(hi));       // exact; the tensor core drops its low 13 bits
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// global -> shared without a register round trip; `bytes` of the 16 (or 4) are read, the rest zero-filled
__device__ __forceinline__ void cp_async16(float* dst, const float* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src, int bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(M2_THREADS)
moments_m2_kernel(const float* __restrict__ p, uint64_t n, int d, int ld, const float* __restrict__ w, int w_stride,
                  const float* __restrict__ centre, uint64_t rows_per_cta, double* __restrict__ out) {
  extern __shared__ __align__(16) float m2_smem[];
  const int tiles = (d + M2_TILE - 1) / M2_TILE;
  int ti = 0, tj = (int)blockIdx.y;                    // blockIdx.y enumerates the pairs ti <= tj row by row
  while (tj >= tiles - ti) { tj -= tiles - ti; ++ti; }
  tj += ti;
  const bool same = ti == tj;
  float* xi = m2_smem;                                              // [STAGES][ROWS][PITCH]
  float* xj = same ? xi : xi + M2_STAGES * M2_STAGE_FLOATS;         // the same rows of the other column tile
  float* ws = xi + M2_STAGES * M2_STAGE_FLOATS * (same ? 1 : 2);    // [STAGES][ROWS] row weights (if any)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;               // mma fragment coordinates
  const int wi = (warp >> 1) * 32, wj = (warp & 1) * 32;   // this warp's quadrant of the tile
  const uint64_t r0 = (uint64_t)blockIdx.x * rows_per_cta;
  const uint64_t r1 = min(n, r0 + rows_per_cta);
  const bool vec = (d % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
  float acc[2][4][4];                                  // [m tile of 16][n tile of 8][fragment]
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[a][b][q] = 0.0f;
  auto flush_acc = [&]() {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ci = ti * M2_TILE + wi + a * 16 + g + (q >> 1) * 8, cj = tj * M2_TILE + wj + b * 8 + 2 * t + (q & 1);
          if (ci < d && cj < d) {
            const double v = (double)acc[a][b][q];
            atomicAdd(out + (size_t)ci * d + cj, v);
            if (!same) atomicAdd(out + (size_t)cj * d + ci, v);      // mirror of an off-diagonal tile
          }
          acc[a][b][q] = 0.0f;
        }
  };
  // centre of the columns this thread's fragments touch (0 past d: those columns are zero-filled)
  float ca[2][2], cb[4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = ti * M2_TILE + wi + a * 16 + g + h * 8;
      ca[a][h] = c < d ? __ldg(centre + c) : 0.0f;
    }
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int c = tj * M2_TILE + wj + b * 8 + g;
    cb[b] = c < d ? __ldg(centre + c) : 0.0f;
  }

  // one stage = 32 rows x 16 column quads per operand: 4 copies per thread and operand
  const int sq = threadIdx.x & 15, sr0 = threadIdx.x >> 4;
  auto issue = [&](uint64_t rb, int slot) {
    if (rb < r1) {
#pragma unroll
      for (int h = 0; h < M2_ROWS / 8; ++h) {
        const int rr = sr0 + h * 8;
        const uint64_t r = rb + rr;
        const bool row_ok = r < r1;
        const float* row = p + (row_ok ? r : r0) * ld;
#pragma unroll
        for (int op = 0; op < 2; ++op) {
          if (op == 1 && same) break;
          const int c0 = (op ? tj : ti) * M2_TILE + sq * 4;
          float* dst = (op ? xj : xi) + slot * M2_STAGE_FLOATS + rr * M2_PITCH + sq * 4;
          if (vec) {
            cp_async16(dst, row + min(c0, d - 4), (row_ok && c0 < d) ? 16 : 0);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) cp_async4(dst + q, row + min(c0 + q, d - 1), (row_ok && c0 + q < d) ? 4 : 0);
          }
        }
      }
      if (w && threadIdx.x < M2_ROWS) {
        const uint64_t r = rb + threadIdx.x;
        cp_async4(ws + slot * M2_ROWS + threadIdx.x, w + (r < r1 ? r : r0) * (uint64_t)w_stride, r < r1 ? 4 : 0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

#pragma unroll
  for (int st = 0; st < M2_STAGES - 1; ++st) issue(r0 + (uint64_t)st * M2_ROWS, st);
  int flush = 0, slot = 0;
  for (uint64_t rb = r0; rb < r1; rb += M2_ROWS) {
    asm volatile("cp.async.wait_group %0;" ::"n"(M2_STAGES - 2) : "memory");
    __syncthreads();                                   // stage `slot` has landed; the slot refilled below is free
    issue(rb + (uint64_t)(M2_STAGES - 1) * M2_ROWS, (slot + M2_STAGES - 1) % M2_STAGES);
    const float* si = xi + slot * M2_STAGE_FLOATS;
    const float* sj = xj + slot * M2_STAGE_FLOATS;
    const float* sw = ws + slot * M2_ROWS;
#pragma unroll
    for (int k0 = 0; k0 < M2_ROWS; k0 += 8) {
      // rows past the end of this CTA's range hold zero-filled data: x - c must not count
      const bool ok0 = rb + k0 + t < r1, ok1 = rb + k0 + t + 4 < r1;
      const float w0 = w ? sw[k0 + t] : 1.0f, w1 = w ? sw[k0 + t + 4] : 1.0f;
      uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int i0 = wi + a * 16 + g;
        const float x0 = ok0 ? w0 * (si[(k0 + t) * M2_PITCH + i0] - ca[a][0]) : 0.0f;
        const float x1 = ok0 ? w0 * (si[(k0 + t) * M2_PITCH + i0 + 8] - ca[a][1]) : 0.0f;
        const float x2 = ok1 ? w1 * (si[(k0 + t + 4) * M2_PITCH + i0] - ca[a][0]) : 0.0f;
        const float x3 = ok1 ? w1 * (si[(k0 + t + 4) * M2_PITCH + i0 + 8] - ca[a][1]) : 0.0f;
        split_tf32(x0, ah[a][0], al[a][0]);
        split_tf32(x1, ah[a][1], al[a][1]);
        split_tf32(x2, ah[a][2], al[a][2]);
        split_tf32(x3, ah[a][3], al[a][3]);
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int j0 = wj + b * 8 + g;
        const float y0 = ok0 ? sj[(k0 + t) * M2_PITCH + j0] - cb[b] : 0.0f;
        const float y1 = ok1 ? sj[(k0 + t + 4) * M2_PITCH + j0] - cb[b] : 0.0f;
        split_tf32(y0, bh[b][0], bl[b][0]);
        split_tf32(y1, bh[b][1], bl[b][1]);
      }
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          mma_tf32(acc[a][b], al[a], bh[b]);           // small terms first
          mma_tf32(acc[a][b], ah[a], bl[b]);
          mma_tf32(acc[a][b], ah[a], bh[b]);
        }
    }
    if (++flush == M2_FLUSH) { flush = 0; flush_acc(); }
    slot = (slot + 1) % M2_STAGES;
  }
  flush_acc();
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
  if (d % 4 == 0 && ld % 4 == 0 && d / 4 <= MS_THREADS && ((uintptr_t)p & 15) == 0) {
    const int rows_per_iter = MS_THREADS / (d / 4);
    const int grid = (int)min((n + rows_per_iter - 1) / rows_per_iter, (uint64_t)sm_count() * 8);
    moments_sum_vec4_kernel<<<grid, MS_THREADS, (size_t)(d + 1) * sizeof(double), (cudaStream_t)stream>>>(
        p, n, d, ld, w, w_stride, sum_out);
    count_launches(1);
    return check_launch("moments_sum");
  }
  int grid = (int)min((n + 7) / 8, (uint64_t)sm_count() * 8);
  moments_sum_kernel<<<grid, MS_THREADS, 0, (cudaStream_t)stream>>>(p, n, d, ld, w, w_stride, sum_out);
  count_launches(1);
  return check_launch("moments_sum");
}

extern "C" int mccb200_moments_m2(const float* p, uint64_t n, int d, int ld, const float* w, int w_stride,
                                  const float* centre, double* m2_out, void* stream) {
  MCCB_REQUIRE(p && centre && m2_out && d >= 1 && ld >= d, "moments_m2: bad argument");
  if (n == 0) return MCCB200_OK;
  if (mccb::moments_m2_tc_applicable(p, d, ld, w, centre))          // d = 64, unweighted: tcgen05 (moments_tc.cu)
    return mccb::moments_m2_tc(p, n, ld, centre, m2_out, (cudaStream_t)stream);
  const int tiles = (d + M2_TILE - 1) / M2_TILE;
  const int pairs = tiles * (tiles + 1) / 2;           // tiles on or above the diagonal
  uint64_t want_ctas = (uint64_t)sm_count() * 8 / (uint64_t)pairs + 1;
  uint64_t rows_per_cta = (n + want_ctas - 1) / want_ctas;
  rows_per_cta = (rows_per_cta + M2_ROWS - 1) / M2_ROWS * M2_ROWS;
  dim3 grid((unsigned)((n + rows_per_cta - 1) / rows_per_cta), (unsigned)pairs);
  const size_t smem = m2_smem_bytes(tiles > 1, w != nullptr);
  // (a per-DEVICE attribute: set on every call)
  cudaFuncSetAttribute(moments_m2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m2_smem_bytes(true, true));
  moments_m2_kernel<<<grid, M2_THREADS, smem, (cudaStream_t)stream>>>(p, n, d, ld, w, w_stride, centre, rows_per_cta, m2_out);
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
