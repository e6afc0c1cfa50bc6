// Index bookkeeping of the sharded (multi-GPU) global MonteCarloKernel resample (sm_100a).
//
// The reference has no multi-device code; this supports SURVEY.md section 8e: every rank holds
// n_loc consecutive parents, the global index vector idx[n_tot] (child ids, identical on every
// rank because it is a pure function of (key, t)) says which child lands in which output slot.
//   owner(s) = parent(idx[s]) / n_loc   -- the rank that can materialise slot s without moving data
//   dest(s)  = s / n_loc                -- the rank that stores output slot s
#include <string.h>

#include "common.cuh"

namespace mccb {

// keys[s] = owner rank of slot s; counts[owner * G + dest] += 1
__global__ void __launch_bounds__(256)
shard_owner_keys_kernel(const uint32_t* __restrict__ idx, uint64_t n_tot, uint64_t children_per_rank,
                        uint64_t n_loc, int G, uint32_t* __restrict__ keys, uint32_t* __restrict__ counts) {
  __shared__ uint32_t s_cnt[64 * 64];
  for (int i = threadIdx.x; i < G * G; i += blockDim.x) s_cnt[i] = 0u;
  __syncthreads();
  for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_tot; s += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t owner = (uint32_t)((uint64_t)idx[s] / children_per_rank);
    const uint32_t dest = (uint32_t)(s / n_loc);
    keys[s] = owner;
    atomicAdd(&s_cnt[owner * G + dest], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G * G; i += blockDim.x)
    if (s_cnt[i]) atomicAdd(&counts[i], s_cnt[i]);
}

// out[q] = idx[order[start + q]] - child_offset   (child ids relative to the local parents)
__global__ void __launch_bounds__(256)
shard_send_ids_kernel(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ order, uint64_t start,
                      uint64_t count, uint64_t child_offset, uint32_t* __restrict__ out) {
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < count; q += (uint64_t)gridDim.x * blockDim.x)
    out[q] = (uint32_t)((uint64_t)idx[order[start + q]] - child_offset);
}

// out[idx[r] - slot_offset, :] = in[r, :]
__global__ void __launch_bounds__(256)
scatter_rows_kernel(const float* __restrict__ in, int ld, const uint32_t* __restrict__ slots, uint64_t slot_offset,
                    uint64_t n_rows, float* __restrict__ out) {
  const int lanes = ld < 32 ? ld : 32;   // threads cooperating on one row (strided over columns)
  const int rows_per_cta = blockDim.x / lanes;
  const int rr = threadIdx.x / lanes, c0 = threadIdx.x % lanes;
  if (rr >= rows_per_cta) return;
  for (uint64_t r = (uint64_t)blockIdx.x * rows_per_cta + rr; r < n_rows; r += (uint64_t)gridDim.x * rows_per_cta) {
    const uint64_t dst = (uint64_t)slots[r] - slot_offset;
    for (int c = c0; c < ld; c += lanes) out[dst * ld + c] = in[r * ld + c];
  }
}

// list[cursor++] = s for every slot s a rank fills (idx_slots[s] != 0xFFFFFFFF): warp-aggregated append, any order
__global__ void __launch_bounds__(256)
compact_slots_kernel(const uint32_t* __restrict__ idx_slots, uint64_t n_slots, uint32_t* __restrict__ list,
                     uint32_t capacity, uint32_t* __restrict__ cursor) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t rounds = (n_slots + stride - 1) / stride;
  for (uint64_t it = 0; it < rounds; ++it) {
    const uint64_t s = it * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool mine = s < n_slots && __ldcs(idx_slots + s) != 0xFFFFFFFFu;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, mine);
    if (m == 0u) continue;
    uint32_t base = 0;
    if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(cursor, (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, __ffs(m) - 1);
    if (mine) {
      const uint32_t at = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
      if (at >= capacity) __trap();                        // loud: the caller sized the list too small
      list[at] = (uint32_t)s;
    }
  }
}


// ---------------------------------------------------------------- id exchange: stable partition of a slice by owner
// A rank's slice of the global index vector (child ids, 32 or 64 bit) -> pairs (global slot, LOCAL child id of the
// owner) grouped by owner rank, ascending slot inside a group: the send buffer of the id exchange.  Three small
// kernels: per-CTA histograms over contiguous chunks, one scan, a stable scatter (tile by tile inside the chunk,
// __match_any_sync ranks inside a warp).  8 bytes out per slot; G <= 64.
constexpr int SB_THREADS = 256;
constexpr int SB_MAX_RANKS = 64;

template <typename IdxT>
__device__ __forceinline__ void sb_decode(const IdxT* idx, uint64_t q, uint64_t children_per_rank, uint32_t& owner,
                                          uint32_t& c_loc) {
  const uint64_t c = (uint64_t)idx[q];
  owner = (uint32_t)(c / children_per_rank);
  c_loc = (uint32_t)(c - (uint64_t)owner * children_per_rank);
}

template <typename IdxT>
__global__ void __launch_bounds__(SB_THREADS)
shard_bucket_count_kernel(const IdxT* __restrict__ idx, uint64_t n, uint64_t chunk, uint64_t children_per_rank, int G,
                          uint32_t* __restrict__ hist /*[gridDim.x][G]*/) {
  __shared__ uint32_t s_cnt[SB_MAX_RANKS];
  for (int i = threadIdx.x; i < G; i += blockDim.x) s_cnt[i] = 0u;
  __syncthreads();
  const uint64_t q0 = (uint64_t)blockIdx.x * chunk, q1 = min(n, q0 + chunk);
  for (uint64_t t0 = q0; t0 < q1; t0 += blockDim.x) {          // (whole warps iterate together)
    const uint64_t q = t0 + threadIdx.x;
    const bool have = q < q1;
    uint32_t owner = 0xFFFFFFFFu, c_loc = 0;
    if (have) sb_decode(idx, q, children_per_rank, owner, c_loc);
    const uint32_t peers = __match_any_sync(0xffffffffu, owner);
    if (have && (peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0u) atomicAdd(&s_cnt[owner], (uint32_t)__popc(peers));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G; i += blockDim.x) hist[(uint64_t)blockIdx.x * G + i] = s_cnt[i];
}

// hist[b][g] -> start of CTA b's entries of bucket g in the grouped output; counts[g] = bucket totals
__global__ void __launch_bounds__(SB_MAX_RANKS)
shard_bucket_scan_kernel(uint32_t* __restrict__ hist, int n_ctas, int G, uint32_t* __restrict__ counts) {
  __shared__ uint32_t s_tot[SB_MAX_RANKS];
  const int g = threadIdx.x;
  uint32_t run = 0;
  if (g < G) {
    for (int b = 0; b < n_ctas; ++b) {
      const uint32_t v = hist[(uint64_t)b * G + g];
      hist[(uint64_t)b * G + g] = run;
      run += v;
    }
    s_tot[g] = run;
    counts[g] = run;
  }
  __syncthreads();
  if (g < G) {
    uint32_t base = 0;
    for (int i = 0; i < g; ++i) base += s_tot[i];
    if (base)
      for (int b = 0; b < n_ctas; ++b) hist[(uint64_t)b * G + g] += base;
  }
}

template <typename IdxT>
__global__ void __launch_bounds__(SB_THREADS)
shard_bucket_scatter_kernel(const IdxT* __restrict__ idx, uint64_t n, uint64_t chunk, uint64_t first_slot,
                            uint64_t children_per_rank, int G, const uint32_t* __restrict__ start /*[gridDim.x][G]*/,
                            uint2* __restrict__ pairs) {
  __shared__ uint32_t s_run[SB_MAX_RANKS];                      // next output position of every bucket (this CTA)
  __shared__ uint32_t s_warp[SB_THREADS / 32][SB_MAX_RANKS];    // per-warp bucket counts of the current tile
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < G; i += blockDim.x) s_run[i] = start[(uint64_t)blockIdx.x * G + i];
  const uint64_t q0 = (uint64_t)blockIdx.x * chunk, q1 = min(n, q0 + chunk);
  for (uint64_t t0 = q0; t0 < q1; t0 += blockDim.x) {          // tiles in order: stability across tiles
    for (int i = threadIdx.x; i < (SB_THREADS / 32) * SB_MAX_RANKS; i += blockDim.x) (&s_warp[0][0])[i] = 0u;
    __syncthreads();
    const uint64_t q = t0 + threadIdx.x;
    const bool have = q < q1;
    uint32_t owner = 0xFFFFFFFFu, c_loc = 0, rank_in_warp = 0;
    if (have) sb_decode(idx, q, children_per_rank, owner, c_loc);
    const uint32_t peers = __match_any_sync(0xffffffffu, owner);
    if (have) {
      rank_in_warp = (uint32_t)__popc(peers & ((1u << lane) - 1u));
      if (rank_in_warp == 0u) s_warp[warp][owner] = (uint32_t)__popc(peers);
    }
    __syncthreads();
    if (have) {
      uint32_t before = 0;
      for (int w = 0; w < warp; ++w) before += s_warp[w][owner];
      pairs[s_run[owner] + before + rank_in_warp] = make_uint2((uint32_t)(first_slot + q), c_loc);
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
      uint32_t tot = 0;
      for (int w = 0; w < SB_THREADS / 32; ++w) tot += s_warp[w][g];
      s_run[g] += tot;
    }
    __syncthreads();
  }
}

}  // namespace mccb

using namespace mccb;

/* Appends to `list` (capacity entries) every slot s < n_slots with idx_slots[s] != 0xFFFFFFFF and adds their
 * number to *cursor (device; the caller zeroes it).  Order is arbitrary; a full list traps. */
extern "C" int mccb200_compact_slots(const uint32_t* idx_slots, uint64_t n_slots, uint32_t* list, uint64_t capacity,
                                     uint32_t* cursor, void* stream) {
  MCCB_REQUIRE(idx_slots && list && cursor && capacity < 0xFFFFFFFFull, "compact_slots: bad argument");
  if (n_slots == 0) return MCCB200_OK;
  const int grid = (int)min((n_slots + 255) / 256, (uint64_t)sm_count() * 8);
  compact_slots_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(idx_slots, n_slots, list, (uint32_t)capacity, cursor);
  count_launches(1);
  return check_launch("compact_slots");
}

// ---- peer shards: plain cudaMalloc buffers shared between the ranks of one node through CUDA IPC.
// The importer opens a handle with ITS device current and cudaIpcMemLazyEnablePeerAccess, which maps
// the exporter's memory for kernels of the importing device (NVLink peer loads / stores).
extern "C" int mccb200_peer_alloc(size_t bytes, void** ptr) {
  MCCB_REQUIRE(ptr && bytes > 0, "peer_alloc: bad argument");
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) return fail(MCCB200_ERR_CUDA, "peer_alloc: %s", cudaGetErrorString(e));
  return MCCB200_OK;
}
extern "C" int mccb200_peer_free(void* ptr) {
  if (ptr && cudaFree(ptr) != cudaSuccess) return fail(MCCB200_ERR_CUDA, "peer_free failed");
  return MCCB200_OK;
}
extern "C" int mccb200_peer_export(const void* ptr, unsigned char* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  MCCB_REQUIRE(ptr && handle64, "peer_export: NULL argument");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(ptr));
  if (e != cudaSuccess) return fail(MCCB200_ERR_CUDA, "peer_export: %s", cudaGetErrorString(e));
  memcpy(handle64, &h, 64);
  return MCCB200_OK;
}
extern "C" int mccb200_peer_open(const unsigned char* handle64, void** ptr) {
  MCCB_REQUIRE(ptr && handle64, "peer_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail(MCCB200_ERR_CUDA, "peer_open: %s", cudaGetErrorString(e));
  return MCCB200_OK;
}
extern "C" int mccb200_peer_close(void* ptr) {
  if (ptr && cudaIpcCloseMemHandle(ptr) != cudaSuccess) return fail(MCCB200_ERR_CUDA, "peer_close failed");
  return MCCB200_OK;
}

extern "C" int mccb200_enable_peer_access(int peer_device) {
  int cur = 0;
  if (cudaGetDevice(&cur) != cudaSuccess) return fail(MCCB200_ERR_CUDA, "enable_peer_access: no current device");
  if (cur == peer_device) return MCCB200_OK;
  int can = 0;
  if (cudaDeviceCanAccessPeer(&can, cur, peer_device) != cudaSuccess || !can)
    return fail(MCCB200_ERR_UNSUPPORTED, "enable_peer_access: device %d cannot access device %d", cur, peer_device);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return MCCB200_OK; }
  if (e != cudaSuccess) return fail(MCCB200_ERR_CUDA, "enable_peer_access: %s", cudaGetErrorString(e));
  return MCCB200_OK;
}

extern "C" int mccb200_shard_owner_keys(const uint32_t* idx, uint64_t n_tot, uint64_t n_loc, int k, int n_ranks,
                                        uint32_t* keys, uint32_t* counts, void* stream) {
  MCCB_REQUIRE(idx && keys && counts && n_loc >= 1 && k >= 1, "shard_owner_keys: bad argument");
  MCCB_REQUIRE(n_ranks >= 1 && n_ranks <= 64, "shard_owner_keys: n_ranks=%d out of range [1, 64]", n_ranks);
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(counts, 0, sizeof(uint32_t) * n_ranks * n_ranks, st) != cudaSuccess)
    return fail(MCCB200_ERR_CUDA, "shard_owner_keys: memset failed");
  if (n_tot == 0) return MCCB200_OK;
  int grid = (int)min((n_tot + 255) / 256, (uint64_t)sm_count() * 8);
  shard_owner_keys_kernel<<<grid, 256, 0, st>>>(idx, n_tot, n_loc * (uint64_t)k, n_loc, n_ranks, keys, counts);
  count_launches(2);
  return check_launch("shard_owner_keys");
}

extern "C" int mccb200_shard_send_ids(const uint32_t* idx, const uint32_t* order, uint64_t start, uint64_t count,
                                      uint64_t child_offset, uint32_t* out, void* stream) {
  MCCB_REQUIRE(idx && order && (out || count == 0), "shard_send_ids: NULL argument");
  if (count == 0) return MCCB200_OK;
  int grid = (int)min((count + 255) / 256, (uint64_t)sm_count() * 8);
  shard_send_ids_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(idx, order, start, count, child_offset, out);
  count_launches(1);
  return check_launch("shard_send_ids");
}

extern "C" int mccb200_scatter_rows(const float* in, int ld, const uint32_t* slots, uint64_t slot_offset,
                                    uint64_t n_rows, float* out, void* stream) {
  MCCB_REQUIRE((in && slots && out) || n_rows == 0, "scatter_rows: NULL argument");
  MCCB_REQUIRE(ld >= 1, "scatter_rows: ld=%d", ld);
  if (n_rows == 0) return MCCB200_OK;
  const int lanes = ld < 32 ? ld : 32;
  const int rows_per_cta = 256 / lanes;
  int grid = (int)min((n_rows + rows_per_cta - 1) / rows_per_cta, (uint64_t)sm_count() * 16);
  scatter_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, ld, slots, slot_offset, n_rows, out);
  count_launches(1);
  return check_launch("scatter_rows");
}

static int sb_grid(uint64_t n, uint64_t& chunk) {
  uint64_t ctas = (uint64_t)sm_count() * 4;
  chunk = (n + ctas - 1) / ctas;
  chunk = (chunk + SB_THREADS - 1) / SB_THREADS * SB_THREADS;
  if (chunk == 0) chunk = SB_THREADS;
  return (int)((n + chunk - 1) / chunk);
}

extern "C" size_t mccb200_shard_bucket_ids_workspace_bytes(int n_ranks) {
  if (n_ranks < 1 || n_ranks > SB_MAX_RANKS) return 0;
  return ((size_t)sm_count() * 4 + 1) * (size_t)n_ranks * sizeof(uint32_t);
}

/* The send buffer of the id exchange (`ShardedMonteCarloStep.step_owner_ids`): `idx` = this rank's slice of the
 * global index vector (child ids, uint32 or uint64), output slot of entry q = first_slot + q.  pairs_out[n][2] =
 * (slot, child id relative to the owner's first parent) grouped by owner rank = child / (n_loc * k), ascending slot
 * inside a group; counts_out[n_ranks] = group sizes.  n_loc * k <= 2^32, slots < 2^32. */
extern "C" int mccb200_shard_bucket_ids(const void* idx, int idx_is_64, uint64_t n, uint64_t first_slot, uint64_t n_loc,
                                        int k, int n_ranks, uint32_t* pairs_out, uint32_t* counts_out, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  MCCB_REQUIRE(idx && pairs_out && counts_out && workspace, "shard_bucket_ids: NULL argument");
  MCCB_REQUIRE(n_ranks >= 1 && n_ranks <= SB_MAX_RANKS && n_loc >= 1 && k >= 1, "shard_bucket_ids: bad shape");
  MCCB_REQUIRE(n_loc * (uint64_t)k <= 0x100000000ull, "shard_bucket_ids: n_loc * k exceeds 2^32 local child ids");
  MCCB_REQUIRE(first_slot + n <= 0x100000000ull, "shard_bucket_ids: slots exceed 32 bits");
  MCCB_REQUIRE(workspace_bytes >= mccb200_shard_bucket_ids_workspace_bytes(n_ranks), "shard_bucket_ids: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    if (cudaMemsetAsync(counts_out, 0, sizeof(uint32_t) * n_ranks, st) != cudaSuccess)
      return fail(MCCB200_ERR_CUDA, "shard_bucket_ids: memset failed");
    return MCCB200_OK;
  }
  uint64_t chunk;
  const int grid = sb_grid(n, chunk);
  const uint64_t cpr = n_loc * (uint64_t)k;
  uint32_t* hist = (uint32_t*)workspace;
  if (idx_is_64) {
    shard_bucket_count_kernel<uint64_t><<<grid, SB_THREADS, 0, st>>>((const uint64_t*)idx, n, chunk, cpr, n_ranks, hist);
    shard_bucket_scan_kernel<<<1, SB_MAX_RANKS, 0, st>>>(hist, grid, n_ranks, counts_out);
    shard_bucket_scatter_kernel<uint64_t><<<grid, SB_THREADS, 0, st>>>((const uint64_t*)idx, n, chunk, first_slot, cpr,
                                                                        n_ranks, hist, (uint2*)pairs_out);
  } else {
    shard_bucket_count_kernel<uint32_t><<<grid, SB_THREADS, 0, st>>>((const uint32_t*)idx, n, chunk, cpr, n_ranks, hist);
    shard_bucket_scan_kernel<<<1, SB_MAX_RANKS, 0, st>>>(hist, grid, n_ranks, counts_out);
    shard_bucket_scatter_kernel<uint32_t><<<grid, SB_THREADS, 0, st>>>((const uint32_t*)idx, n, chunk, first_slot, cpr,
                                                                        n_ranks, hist, (uint2*)pairs_out);
  }
  count_launches(3);
  return check_launch("shard_bucket_ids");
}
