// Median-split (KD) partitioner on the device (sm_100a).
//
// Replaces the host callback of /root/reference/mccube/_kernels/tree.py:44-58 (sklearn KDTree / BallTree built through
// jax.pure_callback, `idx_array` cut into `leaf_size` chunks).  sklearn's construction (sklearn/neighbors/_binary_tree.pxi
// `_recursive_build`, shared by both trees) is: node = positions [start, end) of idx_array; split dimension = the one
// with the largest max - min over the node's points (first on ties); nth_element around n_mid = (end - start) / 2;
// children [start, start + n_mid) and [start + n_mid, end).  Which points land in which child is therefore fixed
// (up to ties at the median); their ORDER inside a child is std::nth_element's and platform dependent.
//
// Here every level is done for all nodes at once:
//   1. per-node, per-dimension min / max (rows gathered in idx order, a warp per run of positions, flushed with
//      ordered-integer atomics when the run crosses a node boundary)                       -- HBM bound: n * d * 4 bytes
//   2. split dimension per node (spread compared in double, as sklearn's float64 trees do)
//   3. key[p] = order-preserving integer image of y[idx[p], dim(node(p))]
//   4. stable LSD radix sort of the positions by key, then by node id (level bits) -- i.e. a segmented sort
//   5. idx'[i] = idx[pos[i]]
// After `levels` levels the first half of every node holds its n_mid smallest points along the split dimension:
// the same sets as sklearn's nodes at that depth, and the tree can go one level deeper than sklearn's (to leaves of
// exactly leaf_size points when the counts are powers of two) so that every partition is a KD box.
#include "common.cuh"

namespace mccb {

int sort_pairs_impl(const uint32_t*, const uint32_t*, uint64_t, int, uint32_t*, uint32_t*, void*, size_t, cudaStream_t);
size_t sort_pairs_ws(uint64_t);

__device__ __forceinline__ uint32_t kd_order_bits(float v) {       // a < b  <=>  image(a) < image(b)
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float kd_order_float(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

// node_start of the next level: children of [s, e) are [s, s + (e - s) / 2) and [s + (e - s) / 2, e)
__global__ void __launch_bounds__(256)
kd_children_kernel(const uint32_t* __restrict__ cur, uint32_t nodes, uint32_t* __restrict__ nxt) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nodes) {
    const uint32_t s = cur[i], e = cur[i + 1];
    nxt[2 * i] = s;
    nxt[2 * i + 1] = s + (e - s) / 2u;
  }
  if (i == 0) nxt[2 * nodes] = cur[nodes];
}

__global__ void __launch_bounds__(256)
kd_init_kernel(uint32_t n, uint32_t* __restrict__ idx, uint32_t* __restrict__ node_start) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) idx[i] = i;
  if (blockIdx.x == 0 && threadIdx.x == 0) { node_start[0] = 0u; node_start[1] = n; }
}

__global__ void __launch_bounds__(256)
kd_fill_kernel(uint32_t* __restrict__ mm, size_t pairs) {          // (min, max) = (+inf image, -inf image)
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += stride) {
    mm[2 * i] = 0xFFFFFFFFu;
    mm[2 * i + 1] = 0u;
  }
}

__device__ __forceinline__ uint32_t kd_node_of(const uint32_t* __restrict__ node_start, uint32_t nodes, uint32_t p) {
  uint32_t lo = 0, hi = nodes;                                     // largest i with node_start[i] <= p (empty nodes skipped)
  while (lo + 1 < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(node_start + mid) <= p) lo = mid; else hi = mid;
  }
  return lo;
}

constexpr int KD_RUN = 128;        // positions per warp
constexpr int KD_MAXT = 8;         // dimensions per lane: d <= 256

// per-node min / max of every dimension; mm[node][dim] = (min image, max image)
template <int T>
__global__ void __launch_bounds__(256)
kd_minmax_kernel(const float* __restrict__ y, int d, int ld, uint32_t n, const uint32_t* __restrict__ idx,
                 const uint32_t* __restrict__ node_start, uint32_t nodes, uint32_t* __restrict__ mm) {
  const int lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t runs = (n + KD_RUN - 1) / KD_RUN;
  for (uint32_t run = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; run < runs; run += warps) {
    const uint32_t p0 = run * KD_RUN, p1 = min(n, p0 + KD_RUN);
    uint32_t node = kd_node_of(node_start, nodes, p0);
    uint32_t node_end = __ldg(node_start + node + 1);
    float lo[T], hi[T];
#pragma unroll
    for (int t = 0; t < T; ++t) { lo[t] = INFINITY; hi[t] = -INFINITY; }
    auto flush = [&]() {
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int j = lane + 32 * t;
        if (j < d && lo[t] <= hi[t]) {
          atomicMin(mm + ((size_t)node * d + j) * 2, kd_order_bits(lo[t]));
          atomicMax(mm + ((size_t)node * d + j) * 2 + 1, kd_order_bits(hi[t]));
        }
        lo[t] = INFINITY; hi[t] = -INFINITY;
      }
    };
    for (uint32_t p = p0; p < p1; p += 4) {
      uint32_t rows[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) rows[u] = p + u < p1 ? __ldg(idx + p + u) : 0u;
      float v[4][T];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const int j = lane + 32 * t;
          v[u][t] = (p + u < p1 && j < d) ? __ldg(y + (size_t)rows[u] * ld + j) : 0.0f;
        }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (p + u >= p1) break;
        while (p + u >= node_end) {                                // warp-uniform: next (non-empty) node
          flush();
          ++node;
          node_end = __ldg(node_start + node + 1);
        }
#pragma unroll
        for (int t = 0; t < T; ++t) {
          if (lane + 32 * t < d) { lo[t] = fminf(lo[t], v[u][t]); hi[t] = fmaxf(hi[t], v[u][t]); }
        }
      }
    }
    flush();
  }
}

// split dimension = argmax of max - min (first on ties), spreads in double like sklearn's float64 node bounds
__global__ void __launch_bounds__(256)
kd_splitdim_kernel(const uint32_t* __restrict__ mm, int d, uint32_t nodes, uint32_t* __restrict__ dims) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nodes) return;
  double best = -1.0;
  uint32_t arg = 0;
  for (int j = 0; j < d; ++j) {
    const uint32_t a = mm[((size_t)i * d + j) * 2], b = mm[((size_t)i * d + j) * 2 + 1];
    if (a > b) continue;                                           // empty node
    const double spread = (double)kd_order_float(b) - (double)kd_order_float(a);
    if (spread > best) { best = spread; arg = (uint32_t)j; }
  }
  dims[i] = arg;
}

// key of every position along its node's split dimension, and the node id of the position.  node_of holds the node
// of the PREVIOUS level on entry (positions never leave their node): the node of this level is its left or right
// child, one comparison against the child boundary instead of a binary search over all node starts.
__global__ void __launch_bounds__(256)
kd_keys_kernel(const float* __restrict__ y, int ld, uint32_t n, const uint32_t* __restrict__ idx,
               const uint32_t* __restrict__ node_start, int level, const uint32_t* __restrict__ dims,
               uint32_t* __restrict__ keys, uint32_t* __restrict__ node_of) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    uint32_t node = 0;
    if (level > 0) {
      const uint32_t left = 2u * node_of[p];
      node = left + (p >= __ldg(node_start + left + 1) ? 1u : 0u);
    }
    keys[p] = kd_order_bits(__ldg(y + (size_t)__ldg(idx + p) * ld + __ldg(dims + node)));
    node_of[p] = node;
  }
}

__global__ void __launch_bounds__(256)
kd_take_kernel(const uint32_t* __restrict__ table, const uint32_t* __restrict__ at, uint32_t n, uint32_t* __restrict__ out) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = __ldg(table + __ldg(at + i));
}

struct KdPlan { uint32_t max_nodes; size_t off_ns[2], off_mm, off_dims, off_keys, off_node_of, off_pos[2], off_idx[2], off_sort, sort_bytes, total; };

static KdPlan kd_plan(uint64_t n, int d, int levels) {
  KdPlan p{};
  p.max_nodes = 1u << (levels > 0 ? levels - 1 : 0);               // nodes of the deepest level that is split
  size_t cur = 0;
  auto take = [&](size_t bytes) { size_t o = cur; cur += align_up(bytes, 256); return o; };
  for (int i = 0; i < 2; ++i) p.off_ns[i] = take(((size_t)2 * p.max_nodes + 1) * 4);
  p.off_mm = take((size_t)p.max_nodes * d * 8);
  p.off_dims = take((size_t)p.max_nodes * 4);
  p.off_keys = take(n * 4);
  p.off_node_of = take(n * 4);
  for (int i = 0; i < 2; ++i) p.off_pos[i] = take(n * 4);
  for (int i = 0; i < 2; ++i) p.off_idx[i] = take(n * 4);
  p.sort_bytes = sort_pairs_ws(n);
  p.off_sort = take(p.sort_bytes);
  p.total = cur;
  return p;
}

}  // namespace mccb

using namespace mccb;

/* Levels the device partitioner splits for partitions of `leaf_size` points: floor(log2(n / leaf_size)) -- one more
 * than sklearn's tree (n_levels - 1 = floor(log2((n - 1) / leaf_size)) split levels) when n / leaf_size is a power of
 * two, so that the leaves are single partitions. */
extern "C" int mccb200_kd_partition_levels(uint64_t n, uint64_t leaf_size) {
  if (leaf_size == 0 || n < leaf_size) return 0;
  int l = 0;
  while ((n >> (l + 1)) >= leaf_size) ++l;
  return l;
}

extern "C" size_t mccb200_kd_partition_workspace_bytes(uint64_t n, int d, int levels) {
  if (levels < 0 || levels > 30) return 0;
  return kd_plan(n, d, levels).total;
}

/* idx_out[n] = the rows of y ordered as the leaves of a median-split tree of `levels` levels (see the file header;
 * replaces mccube/_kernels/tree.py:44-58).  y: [n, ld] float32 on the device, the first d columns are coordinates. */
extern "C" int mccb200_kd_partition(const float* y, uint64_t n, int d, int ld, int levels, uint32_t* idx_out,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MCCB_REQUIRE(y && idx_out && workspace, "kd_partition: NULL argument");
  MCCB_REQUIRE(n >= 1 && n < 0xFFFFFFFFull && d >= 1 && d <= 32 * KD_MAXT && ld >= d, "kd_partition: n=%llu d=%d ld=%d out of range",
               (unsigned long long)n, d, ld);
  MCCB_REQUIRE(levels >= 0 && levels <= 30 && (n >> levels) >= 1, "kd_partition: levels=%d for n=%llu", levels, (unsigned long long)n);
  const KdPlan p = kd_plan(n, d, levels);
  if (workspace_bytes < p.total) return fail(MCCB200_ERR_WORKSPACE_TOO_SMALL, "kd_partition: workspace %zu < %zu", workspace_bytes, p.total);
  char* w = (char*)workspace;
  uint32_t* ns[2] = {(uint32_t*)(w + p.off_ns[0]), (uint32_t*)(w + p.off_ns[1])};
  uint32_t* mm = (uint32_t*)(w + p.off_mm);
  uint32_t* dims = (uint32_t*)(w + p.off_dims);
  uint32_t* keys = (uint32_t*)(w + p.off_keys);
  uint32_t* node_of = (uint32_t*)(w + p.off_node_of);
  uint32_t* pos[2] = {(uint32_t*)(w + p.off_pos[0]), (uint32_t*)(w + p.off_pos[1])};
  uint32_t* idx[2] = {(uint32_t*)(w + p.off_idx[0]), (uint32_t*)(w + p.off_idx[1])};
  const int sms = sm_count();
  const uint32_t n32 = (uint32_t)n;
  const int grid_n = (int)min(((uint64_t)n + 255) / 256, (uint64_t)sms * 8);
  const int grid_runs = (int)min(((uint64_t)n + KD_RUN * 8 - 1) / (KD_RUN * 8), (uint64_t)sms * 8);
  phase_mark("kd_partition", st);
  uint32_t* cur_idx = levels == 0 ? idx_out : idx[0];
  kd_init_kernel<<<grid_n, 256, 0, st>>>(n32, cur_idx, ns[0]);
  count_launches(1);
  const int T = (d + 31) / 32;
  for (int level = 0; level < levels; ++level) {
    const uint32_t nodes = 1u << level;
    const uint32_t* node_start = ns[level & 1];
    uint32_t* nxt_idx = level == levels - 1 ? idx_out : idx[(level + 1) & 1];
    kd_fill_kernel<<<(int)min(((uint64_t)nodes * d + 255) / 256, (uint64_t)sms * 8), 256, 0, st>>>(mm, (size_t)nodes * d);
#define MCCB_KD_MM(TT) kd_minmax_kernel<TT><<<grid_runs, 256, 0, st>>>(y, d, ld, n32, cur_idx, node_start, nodes, mm)
    switch (T) {
      case 1: MCCB_KD_MM(1); break;
      case 2: MCCB_KD_MM(2); break;
      case 3: MCCB_KD_MM(3); break;
      case 4: MCCB_KD_MM(4); break;
      default: MCCB_KD_MM(KD_MAXT); break;
    }
#undef MCCB_KD_MM
    kd_splitdim_kernel<<<(nodes + 255) / 256, 256, 0, st>>>(mm, d, nodes, dims);
    kd_keys_kernel<<<grid_n, 256, 0, st>>>(y, ld, n32, cur_idx, node_start, level, dims, keys, node_of);
    int rc = sort_pairs_impl(keys, nullptr, n, 32, nullptr, pos[0], w + p.off_sort, p.sort_bytes, st);
    if (rc) return rc;
    const uint32_t* order = pos[0];
    if (level > 0) {                                               // stable second key: the node id (level bits)
      kd_take_kernel<<<grid_n, 256, 0, st>>>(node_of, pos[0], n32, keys);
      rc = sort_pairs_impl(keys, pos[0], n, level, nullptr, pos[1], w + p.off_sort, p.sort_bytes, st);
      if (rc) return rc;
      order = pos[1];
      count_launches(1);
    }
    kd_take_kernel<<<grid_n, 256, 0, st>>>(cur_idx, order, n32, nxt_idx);
    kd_children_kernel<<<(nodes + 255) / 256, 256, 0, st>>>(node_start, nodes, ns[(level + 1) & 1]);
    count_launches(6);
    cur_idx = nxt_idx;
  }
  phase_mark("end", st);
  return check_launch("kd_partition");
}
