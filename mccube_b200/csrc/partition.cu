// Partitioning keys (sm_100a): stratified norms and integer box keys, on materialised rows
// and on the implicit children of an expansion step.
//
// Replaces /root/reference/mccube/_kernels/stratified.py:36-45 (COM-centre -> norm ->
// argsort; the argsort itself is radix_sort.cu) and provides BoxPartitionKernel's key, a
// kernel the brief names but the reference does not contain: its
// definition is oracle/mccube_ref.py::box_keys.
#include "common.cuh"
#include "expand.cuh"

namespace mccb {

constexpr int BB_THREADS = 256;
constexpr int BB_MAX_CTAS = 1024;

// One thread per row keeps the reference's left-to-right summation order (bit-exact keys), but a thread
// walking its own 256-byte row touches 32 lines per warp load (1.1 TB/s).  So a CTA stages 256 rows x 32
// columns at a time through shared memory with coalesced 128-byte reads, and every thread then sums its
// row of the tile in order (pitch 33: conflict-free both ways).
constexpr int SK_ROWS = 256, SK_COLS = 32;
__global__ void __launch_bounds__(SK_ROWS)
stratified_keys_kernel(const float* __restrict__ p, uint64_t n, int d, int ld, const float* __restrict__ com,
                       uint32_t* __restrict__ keys) {
  __shared__ float tile[SK_ROWS][SK_COLS + 1];
  __shared__ float s_com[SK_COLS];
  const int lane = threadIdx.x & 31, wrow = threadIdx.x >> 5;
  for (uint64_t rb = (uint64_t)blockIdx.x * SK_ROWS; rb < n; rb += (uint64_t)gridDim.x * SK_ROWS) {
    float acc = 0.0f;
    for (int cb = 0; cb < d; cb += SK_COLS) {
      __syncthreads();                                      // the previous chunk has been consumed
      if (threadIdx.x < SK_COLS) s_com[threadIdx.x] = cb + threadIdx.x < d ? __ldg(com + cb + threadIdx.x) : 0.0f;
      float v[SK_ROWS / 8];
#pragma unroll
      for (int q = 0; q < SK_ROWS / 8; ++q) {               // warp w loads rows w, w + 8, ...: 128 contiguous bytes each
        const uint64_t r = rb + wrow + q * 8;
        v[q] = (r < n && cb + lane < d) ? __ldcs(p + r * ld + cb + lane) : 0.0f;
      }
#pragma unroll
      for (int q = 0; q < SK_ROWS / 8; ++q) tile[wrow + q * 8][lane] = v[q];
      __syncthreads();
      const int cols = min(SK_COLS, d - cb);
      for (int c = 0; c < cols; ++c) {
        const float x = __fsub_rn(tile[threadIdx.x][c], s_com[c]);
        acc = __fadd_rn(acc, __fmul_rn(x, x));
      }
    }
    // norm >= 0 (or NaN): the IEEE bit pattern of a non-negative float is order preserving.
    if (rb + threadIdx.x < n) keys[rb + threadIdx.x] = __float_as_uint(__fsqrt_rn(acc));
  }
}

// The same keys for rows that are 16-byte chunks end to end (ld == d, d a multiple of 32): the tile of SKA_ROWS rows is
// one contiguous block of global memory, copied with 16-byte cp.async into shared memory through a three-stage ring
// (loads of tiles i+1, i+2 in flight while tile i is reduced), each row's chunks XOR-swizzled by the row number so that
// a thread reading ITS row chunk by chunk (the sequential order of the sum is part of the contract) meets no bank
// conflict.  No register staging, no second pass through shared memory.
constexpr int SKA_ROWS = 128, SKA_STAGES = 3;

__device__ __forceinline__ void ska_cp16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

__global__ void __launch_bounds__(SKA_ROWS)
stratified_keys_async_kernel(const float* __restrict__ p, uint64_t n, int d, const float* __restrict__ com,
                             uint32_t* __restrict__ keys) {
  extern __shared__ __align__(16) unsigned char ska_smem[];
  const int cpr = d >> 2;                                          // 16-byte chunks per row (a multiple of 8)
  const size_t tile_bytes = (size_t)SKA_ROWS * d * 4;
  float* s_com = reinterpret_cast<float*>(ska_smem + SKA_STAGES * tile_bytes);
  for (int c = threadIdx.x; c < d; c += SKA_ROWS) s_com[c] = __ldg(com + c);
  const uint64_t tiles = (n + SKA_ROWS - 1) / SKA_ROWS;
  const int chunks = SKA_ROWS * cpr;

  auto issue = [&](uint64_t tile, int stage) {
    if (tile < tiles) {
      const uint64_t row0 = tile * SKA_ROWS;
      const float4* src = reinterpret_cast<const float4*>(p + row0 * d);
      unsigned char* dst = ska_smem + stage * tile_bytes;
      const uint64_t valid = min((uint64_t)SKA_ROWS, n - row0) * cpr;
      for (int g = threadIdx.x; g < chunks; g += SKA_ROWS) {
        if ((uint64_t)g >= valid) break;
        const int row = g / cpr, c = g - row * cpr;
        ska_cp16(dst + ((size_t)row * cpr + ((c & ~7) | ((c ^ row) & 7))) * 16, src + g);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  uint64_t tile = blockIdx.x;
  for (int s = 0; s < SKA_STAGES - 1; ++s) issue(tile + (uint64_t)s * gridDim.x, s);
  int stage = 0;
  for (; tile < tiles; tile += gridDim.x) {
    issue(tile + (uint64_t)(SKA_STAGES - 1) * gridDim.x, (stage + SKA_STAGES - 1) % SKA_STAGES);
    asm volatile("cp.async.wait_group %0;" ::"n"(SKA_STAGES - 1) : "memory");
    __syncthreads();                                               // every thread's chunks of this stage have landed
    const uint64_t r = tile * SKA_ROWS + threadIdx.x;
    const float4* rowp = reinterpret_cast<const float4*>(ska_smem + stage * tile_bytes) + (size_t)threadIdx.x * cpr;
    float acc = 0.0f;
    if (r < n) {
      for (int c = 0; c < cpr; ++c) {
        const float4 v = rowp[(c & ~7) | ((c ^ (int)threadIdx.x) & 7)];
        const float4 m = reinterpret_cast<const float4*>(s_com)[c];
        float x = __fsub_rn(v.x, m.x); acc = __fadd_rn(acc, __fmul_rn(x, x));
        x = __fsub_rn(v.y, m.y); acc = __fadd_rn(acc, __fmul_rn(x, x));
        x = __fsub_rn(v.z, m.z); acc = __fadd_rn(acc, __fmul_rn(x, x));
        x = __fsub_rn(v.w, m.w); acc = __fadd_rn(acc, __fmul_rn(x, x));
      }
      keys[r] = __float_as_uint(__fsqrt_rn(acc));
    }
    __syncthreads();                                               // the stage may be refilled
    stage = (stage + 1) % SKA_STAGES;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void block_minmax(float lo, float hi, float* out_lo, float* out_hi) {
  __shared__ float s_lo[BB_THREADS / 32], s_hi[BB_THREADS / 32];
  for (int o = 16; o; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < BB_THREADS / 32; ++w) { lo = fminf(lo, s_lo[w]); hi = fmaxf(hi, s_hi[w]); }
    *out_lo = lo; *out_hi = hi;
  }
}

// partials layout: [cta][dim][2]
// One pass over the rows for all key dims, four rows in flight per thread; VEC4: the dims are 4 consecutive
// 16-byte aligned columns, read as one 128-bit load with a 64-byte L2 fill per row.
template <bool VEC4>
__global__ void __launch_bounds__(BB_THREADS)
box_bounds_rows_kernel(const float* __restrict__ p, uint64_t n, int ld, BoxSpec spec, float* __restrict__ partials) {
  float lo[MCCB200_BOX_MAX_DIMS], hi[MCCB200_BOX_MAX_DIMS];
#pragma unroll
  for (int q = 0; q < MCCB200_BOX_MAX_DIMS; ++q) { lo[q] = INFINITY; hi[q] = -INFINITY; }
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += 4 * stride) {
    if (VEC4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint64_t ru = r + (uint64_t)u * stride;
        v[u] = ld_strided4(p + (ru < n ? ru : r) * ld + spec.dims[0]);   // past the end: repeat row r
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        lo[0] = fminf(lo[0], v[u].x); hi[0] = fmaxf(hi[0], v[u].x);
        lo[1] = fminf(lo[1], v[u].y); hi[1] = fmaxf(hi[1], v[u].y);
        lo[2] = fminf(lo[2], v[u].z); hi[2] = fmaxf(hi[2], v[u].z);
        lo[3] = fminf(lo[3], v[u].w); hi[3] = fmaxf(hi[3], v[u].w);
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint64_t ru = r + (uint64_t)u * stride;
        const float* row = p + (ru < n ? ru : r) * ld;
#pragma unroll
        for (int q = 0; q < MCCB200_BOX_MAX_DIMS; ++q)
          if (q < spec.n_dims) {
            const float x = ld_strided(row + spec.dims[q]);
            lo[q] = fminf(lo[q], x); hi[q] = fmaxf(hi[q], x);
          }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < MCCB200_BOX_MAX_DIMS; ++q)
    if (q < spec.n_dims)
      block_minmax(lo[q], hi[q], partials + ((size_t)blockIdx.x * spec.n_dims + q) * 2,
                   partials + ((size_t)blockIdx.x * spec.n_dims + q) * 2 + 1);
}

template <bool VEC4>
__global__ void __launch_bounds__(BB_THREADS)
box_bounds_children_kernel(ExpandArgs a, BoxSpec spec, float* __restrict__ partials) {
  // one pass: every thread keeps the running min/max of all key dims of its parents
  float lo[MCCB200_BOX_MAX_DIMS], hi[MCCB200_BOX_MAX_DIMS];
#pragma unroll
  for (int q = 0; q < MCCB200_BOX_MAX_DIMS; ++q) { lo[q] = INFINITY; hi[q] = -INFINITY; }
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  if (VEC4) {
    // 4 consecutive aligned key columns and the Hadamard cubature: one 128-bit load per parent (from the
    // packed key columns, or a 64-byte L2 fill from the strided state), four parents in flight per thread
    const KeyConst4 kc = key_const4(a, spec);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += 4 * stride) {
      float vm[4][4], vp[4][4];
      float4 y4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {                                  // all four loads first
        const uint64_t iu = i + (uint64_t)u * stride;
        y4[u] = key_columns4(a, spec, iu < a.n ? iu : i);            // past the end: repeat parent i (min/max unchanged)
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint64_t iu = i + (uint64_t)u * stride;
        key_children4(a, spec, kc, iu < a.n ? iu : i, y4[u], vm[u], vp[u]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          lo[q] = fminf(lo[q], fminf(vm[u][q], vp[u][q]));
          hi[q] = fmaxf(hi[q], fmaxf(vm[u][q], vp[u][q]));
        }
      }
    }
  } else {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
#pragma unroll
      for (int q = 0; q < MCCB200_BOX_MAX_DIMS; ++q) {
        if (q < spec.n_dims) {
          const int c = spec.dims[q];
          const float y = ld_strided(a.y0 + i * a.ld + c);
          const float fv = a.drift_kind == 1 ? ld_strided(a.f + i * a.d + c) : drift_value(a, i, c, y);
          if (a.points) {
            for (uint32_t j = 0; j < a.k; ++j) {
              const float v = child_value(a, y, fv, noise_value(a, j, c));
              lo[q] = fminf(lo[q], v); hi[q] = fmaxf(hi[q], v);
            }
          } else {
            // every Hadamard column takes both signs over the k rows (rows j and j + k/2 negate)
            const float v0 = child_value(a, y, fv, -a.scale), v1 = child_value(a, y, fv, a.scale);
            lo[q] = fminf(lo[q], fminf(v0, v1)); hi[q] = fmaxf(hi[q], fmaxf(v0, v1));
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < MCCB200_BOX_MAX_DIMS; ++q)
    if (q < spec.n_dims)
      block_minmax(lo[q], hi[q], partials + ((size_t)blockIdx.x * spec.n_dims + q) * 2,
                   partials + ((size_t)blockIdx.x * spec.n_dims + q) * 2 + 1);
}

// one warp per key dim: lanes stride over the CTA partials, shuffle-reduce
// raw: write (lo, hi) instead of (lo, cells / range) -- for the cross-device min / max of a global partition
__global__ void box_bounds_final_kernel(const float* __restrict__ partials, int n_ctas, BoxSpec spec,
                                        float* __restrict__ bounds, int raw = 0) {
  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (q >= spec.n_dims) return;
  float lo = INFINITY, hi = -INFINITY;
  for (int b = lane; b < n_ctas; b += 32) {
    lo = fminf(lo, partials[((size_t)b * spec.n_dims + q) * 2]);
    hi = fmaxf(hi, partials[((size_t)b * spec.n_dims + q) * 2 + 1]);
  }
  for (int o = 16; o; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) {
    float rng = __fsub_rn(hi, lo);
    bounds[q] = lo;
    bounds[spec.n_dims + q] = raw ? hi : (rng > 0.0f ? __fdiv_rn((float)(1 << spec.bits), rng) : 0.0f);
  }
}

__device__ __forceinline__ uint32_t box_cell(float x, float lo, float inv_h, int bits) {
  float v = __fmul_rn(__fsub_rn(x, lo), inv_h);
  float c = fminf(fmaxf(floorf(v), 0.0f), (float)((1 << bits) - 1));
  return (uint32_t)c;  // NaN -> 0
}

__global__ void __launch_bounds__(256)
box_keys_rows_kernel(const float* __restrict__ p, uint64_t n, int ld, BoxSpec spec, const float* __restrict__ bounds,
                     uint32_t* __restrict__ keys) {
  float lo[MCCB200_BOX_MAX_DIMS], ih[MCCB200_BOX_MAX_DIMS];
  for (int q = 0; q < spec.n_dims; ++q) { lo[q] = bounds[q]; ih[q] = bounds[spec.n_dims + q]; }
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t key = 0;
    for (int q = 0; q < spec.n_dims; ++q)
      key = (key << spec.bits) | box_cell(p[r * ld + spec.dims[q]], lo[q], ih[q], spec.bits);
    keys[r] = key;
  }
}

__global__ void __launch_bounds__(256)
box_keys_children_kernel(ExpandArgs a, BoxSpec spec, const float* __restrict__ bounds, uint32_t* __restrict__ keys) {
  float lo[MCCB200_BOX_MAX_DIMS], ih[MCCB200_BOX_MAX_DIMS];
  for (int q = 0; q < spec.n_dims; ++q) { lo[q] = bounds[q]; ih[q] = bounds[spec.n_dims + q]; }
  const uint64_t total = a.n * (uint64_t)a.k;
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = c / a.k;
    const uint32_t j = (uint32_t)(c - i * a.k);
    uint32_t key = 0;
    for (int q = 0; q < spec.n_dims; ++q) {
      const int col = spec.dims[q];
      float y = __ldg(a.y0 + i * a.ld + col);
      float x = child_value(a, y, drift_value(a, i, col, y), noise_value(a, j, col));
      key = (key << spec.bits) | box_cell(x, lo[q], ih[q], spec.bits);
    }
    keys[c] = key;
  }
}

int make_box_spec(BoxSpec& s, const int32_t* dims_host, int n_dims, int bits, int d) {
  MCCB_REQUIRE(dims_host && n_dims >= 1 && n_dims <= MCCB200_BOX_MAX_DIMS, "box: n_dims=%d out of range", n_dims);
  MCCB_REQUIRE(bits >= 1 && n_dims * bits <= 32, "box: n_dims*bits=%d exceeds 32", n_dims * bits);
  s.n_dims = n_dims; s.bits = bits;
  for (int q = 0; q < n_dims; ++q) {
    MCCB_REQUIRE(dims_host[q] >= 0 && dims_host[q] < d, "box: dim %d out of range [0, %d)", dims_host[q], d);
    s.dims[q] = dims_host[q];
  }
  return MCCB200_OK;
}

}  // namespace mccb

using namespace mccb;

extern "C" size_t mccb200_box_bounds_workspace_bytes(void) {
  return (size_t)BB_MAX_CTAS * MCCB200_BOX_MAX_DIMS * 2 * sizeof(float);
}

extern "C" int mccb200_stratified_keys(const float* p, uint64_t n, int d, int ld, const float* com, uint32_t* keys,
                                       void* stream) {
  MCCB_REQUIRE(p && com && keys && d >= 1 && ld >= d, "stratified_keys: bad argument");
  if (n == 0) return MCCB200_OK;
  if (ld == d && d % 32 == 0 && d <= 128 && (((uintptr_t)p) & 15) == 0 && env_int("MCCB200_STRATIFIED_ASYNC", 1) != 0) {
    const size_t smem = (size_t)SKA_STAGES * SKA_ROWS * d * 4 + (size_t)d * 4;
    static thread_local int attr_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_dev != dev) {
      cudaFuncSetAttribute(stratified_keys_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr_dev = dev;
    }
    const int per_sm = (int)max((size_t)1, min((size_t)8, (size_t)(220 * 1024) / smem));
    int grid = (int)min((n + SKA_ROWS - 1) / SKA_ROWS, (uint64_t)sm_count() * per_sm);
    stratified_keys_async_kernel<<<grid, SKA_ROWS, smem, (cudaStream_t)stream>>>(p, n, d, com, keys);
    count_launches(1);
    return check_launch("stratified_keys");
  }
  int grid = (int)min((n + SK_ROWS - 1) / SK_ROWS, (uint64_t)sm_count() * 6);
  stratified_keys_kernel<<<grid, SK_ROWS, 0, (cudaStream_t)stream>>>(p, n, d, ld, com, keys);
  count_launches(1);
  return check_launch("stratified_keys");
}

extern "C" int mccb200_box_bounds(const float* p, uint64_t n, int ld, const int32_t* dims_host, int n_dims, int bits,
                                  float* bounds, void* workspace, size_t workspace_bytes, void* stream) {
  BoxSpec s;
  int rc = make_box_spec(s, dims_host, n_dims, bits, ld);
  if (rc) return rc;
  MCCB_REQUIRE(p && bounds && n >= 1, "box_bounds: bad argument");
  if (!workspace || workspace_bytes < mccb200_box_bounds_workspace_bytes())
    return fail(MCCB200_ERR_WORKSPACE_TOO_SMALL, "box_bounds: workspace too small");
  int grid = (int)min((n + BB_THREADS - 1) / BB_THREADS, (uint64_t)min(BB_MAX_CTAS, sm_count() * 4));
  cudaStream_t st = (cudaStream_t)stream;
  bool vec4 = s.n_dims == 4 && (s.dims[0] & 3) == 0 && (ld & 3) == 0 && (((uintptr_t)p) & 15) == 0;
  for (int q = 1; q < 4 && vec4; ++q) vec4 = s.dims[q] == s.dims[0] + q;
  if (vec4) box_bounds_rows_kernel<true><<<grid, BB_THREADS, 0, st>>>(p, n, ld, s, (float*)workspace);
  else box_bounds_rows_kernel<false><<<grid, BB_THREADS, 0, st>>>(p, n, ld, s, (float*)workspace);
  box_bounds_final_kernel<<<1, 32 * MCCB200_BOX_MAX_DIMS, 0, st>>>((const float*)workspace, grid, s, bounds);
  count_launches(2);
  return check_launch("box_bounds");
}

extern "C" int mccb200_box_keys(const float* p, uint64_t n, int ld, const int32_t* dims_host, int n_dims, int bits,
                                const float* bounds, uint32_t* keys, void* stream) {
  BoxSpec s;
  int rc = make_box_spec(s, dims_host, n_dims, bits, ld);
  if (rc) return rc;
  MCCB_REQUIRE(p && bounds && keys, "box_keys: NULL argument");
  if (n == 0) return MCCB200_OK;
  int grid = (int)min((n + 255) / 256, (uint64_t)sm_count() * 8);
  box_keys_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, n, ld, s, bounds, keys);
  count_launches(1);
  return check_launch("box_keys");
}

static int child_bounds_impl(const float* y0, uint64_t n, int d, int ld, const mccb200_drift* drift, float dt,
                             const mccb200_cubature* cub, const int32_t* dims_host, int n_dims, int bits,
                             const float* keycols, float* bounds, void* workspace, size_t workspace_bytes, void* stream,
                             int raw) {
  BoxSpec s;
  int rc = make_box_spec(s, dims_host, n_dims, bits, d);
  if (rc) return rc;
  ExpandArgs a{};
  rc = fill_expand_args(a, y0, n, d, ld != d, drift, dt, cub);
  if (rc) return rc;
  MCCB_REQUIRE(bounds && n >= 1, "box_child_bounds: bad argument");
  if (!workspace || workspace_bytes < mccb200_box_bounds_workspace_bytes())
    return fail(MCCB200_ERR_WORKSPACE_TOO_SMALL, "box_child_bounds: workspace too small");
  // grid-stride kernel: exactly one resident wave (6 CTAs/SM wanted, 4 fit at 58 registers: a grid of 6 per SM
  // ran as 1.5 waves, the second one half empty)
  static int per_sm_vec = 0, per_sm_gen = 0;
  if (!per_sm_vec) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_vec, box_bounds_children_kernel<true>, BB_THREADS, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_gen, box_bounds_children_kernel<false>, BB_THREADS, 0);
    per_sm_vec = max(1, per_sm_vec); per_sm_gen = max(1, per_sm_gen);
  }
  const bool vec_path = !a.points && box_dims_vec4(s, a);
  int grid = (int)min((n + BB_THREADS - 1) / BB_THREADS,
                      (uint64_t)min(BB_MAX_CTAS, sm_count() * (vec_path ? per_sm_vec : per_sm_gen)));
  cudaStream_t st = (cudaStream_t)stream;
  phase_mark("box_child_bounds", st);
  a.keycols = (!a.points && box_dims_vec4(s, a) && keycols && (((uintptr_t)keycols & 15) == 0)) ? keycols : nullptr;
  if (!a.points && box_dims_vec4(s, a)) box_bounds_children_kernel<true><<<grid, BB_THREADS, 0, st>>>(a, s, (float*)workspace);
  else box_bounds_children_kernel<false><<<grid, BB_THREADS, 0, st>>>(a, s, (float*)workspace);
  box_bounds_final_kernel<<<1, 32 * MCCB200_BOX_MAX_DIMS, 0, st>>>((const float*)workspace, grid, s, bounds, raw);
  phase_mark("end", st);
  count_launches(2);
  return check_launch("box_child_bounds");
}

extern "C" int mccb200_box_child_bounds(const float* y0, uint64_t n, int d, int ld, const mccb200_drift* drift,
                                        float dt, const mccb200_cubature* cub, const int32_t* dims_host, int n_dims,
                                        int bits, const float* keycols, float* bounds, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  return child_bounds_impl(y0, n, d, ld, drift, dt, cub, dims_host, n_dims, bits, keycols, bounds, workspace,
                           workspace_bytes, stream, 0);
}

/* minmax[q] = min, minmax[n_dims + q] = max of key dim q over the implicit children: what a global partition over
 * several shards reduces across devices before forming the cell width (float32: 2^bits / (max - min)). */
extern "C" int mccb200_box_child_minmax(const float* y0, uint64_t n, int d, int ld, const mccb200_drift* drift,
                                        float dt, const mccb200_cubature* cub, const int32_t* dims_host, int n_dims,
                                        int bits, const float* keycols, float* minmax, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  return child_bounds_impl(y0, n, d, ld, drift, dt, cub, dims_host, n_dims, bits, keycols, minmax, workspace,
                           workspace_bytes, stream, 1);
}

extern "C" int mccb200_box_child_keys(const float* y0, uint64_t n, int d, int ld, const mccb200_drift* drift,
                                      float dt, const mccb200_cubature* cub, const int32_t* dims_host, int n_dims,
                                      int bits, const float* bounds, uint32_t* keys, void* stream) {
  BoxSpec s;
  int rc = make_box_spec(s, dims_host, n_dims, bits, d);
  if (rc) return rc;
  ExpandArgs a{};
  rc = fill_expand_args(a, y0, n, d, ld != d, drift, dt, cub);
  if (rc) return rc;
  MCCB_REQUIRE(bounds && keys, "box_child_keys: NULL argument");
  if (n == 0) return MCCB200_OK;
  uint64_t total = n * (uint64_t)cub->k;
  int grid = (int)min((total + 255) / 256, (uint64_t)sm_count() * 16);
  phase_mark("box_child_keys", (cudaStream_t)stream);
  box_keys_children_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, s, bounds, keys);
  phase_mark("end", (cudaStream_t)stream);
  count_launches(1);
  return check_launch("box_child_keys");
}
