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

// The same with 16-byte accesses (ld a multiple of 4, 16-byte aligned bases): ld / 4 threads per row, four rows in
// flight per thread, streaming loads (the rows are read once) -- the scalar form above moved one word per thread and row.
__global__ void __launch_bounds__(256)
scatter_rows_vec4_kernel(const float4* __restrict__ in, int vpr, int rows_per_iter, const uint32_t* __restrict__ slots,
                         uint64_t slot_offset, uint64_t n_rows, float4* __restrict__ out) {
  const int rr = threadIdx.x / vpr, c = threadIdx.x - rr * vpr;
  if (rr >= rows_per_iter) return;
  const uint64_t step = (uint64_t)gridDim.x * rows_per_iter;
  for (uint64_t r0 = (uint64_t)blockIdx.x * rows_per_iter + rr; r0 < n_rows; r0 += 4 * step) {
    uint64_t dst[4];
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t r = r0 + u * step;
      if (r < n_rows) {
        dst[u] = (uint64_t)__ldg(slots + r) - slot_offset;
        v[u] = __ldcs(in + r * vpr + c);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (r0 + u * step < n_rows) __stcs(out + dst[u] * vpr + c, v[u]);
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
// A rank's slice of the global index vector (child ids, 32 or 64 bit) -> (global slot, LOCAL child id of the
// owner), grouped by owner rank and ascending in slot inside a group: the send buffers of the id exchange.
// Every WARP owns a contiguous chunk of the slice: a histogram pass (per-warp counts), one scan over
// [bucket][warp], and a scatter pass in which a warp walks its chunk 32 entries at a time (__match_any_sync gives
// the rank among the same owner inside the warp, the warp's running offsets live in its own shared-memory row), so
// there is no block-wide synchronisation in either pass.  8 bytes out per slot; G <= 64.
constexpr int SB_THREADS = 256;
constexpr int SB_WARPS = SB_THREADS / 32;
constexpr int SB_MAX_RANKS = 64;

struct OwnerOf {                 // owner = child / children_per_rank (a shift when that is a power of two)
  uint64_t cpr;
  int shift;                     // log2(cpr) or -1
  __device__ __forceinline__ void decode(uint64_t c, uint32_t& owner, uint32_t& c_loc) const {
    if (shift >= 0) { owner = (uint32_t)(c >> shift); c_loc = (uint32_t)(c & (cpr - 1ull)); }
    else { owner = (uint32_t)(c / cpr); c_loc = (uint32_t)(c - (uint64_t)owner * cpr); }
  }
};

template <typename IdxT>
__global__ void __launch_bounds__(SB_THREADS)
shard_bucket_count_kernel(const IdxT* __restrict__ idx, uint64_t n, uint64_t chunk, OwnerOf own, int G,
                          uint32_t n_warps, uint32_t* __restrict__ hist /*[G][n_warps]*/) {
  __shared__ uint32_t s_cnt[SB_WARPS][SB_MAX_RANKS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * SB_WARPS + warp;
  for (int g = lane; g < G; g += 32) s_cnt[warp][g] = 0u;
  __syncwarp();
  const uint64_t q0 = (uint64_t)gw * chunk, q1 = min(n, q0 + chunk);
  for (uint64_t t0 = q0; t0 < q1; t0 += 32) {
    const uint64_t q = t0 + lane;
    uint32_t owner = 0xFFFFFFFFu, c_loc;
    if (q < q1) own.decode((uint64_t)__ldcs(idx + q), owner, c_loc);
    const uint32_t peers = __match_any_sync(0xffffffffu, owner);
    if (q < q1 && (peers & ((1u << lane) - 1u)) == 0u) s_cnt[warp][owner] += (uint32_t)__popc(peers);
    __syncwarp();
  }
  if (gw < n_warps)
    for (int g = lane; g < G; g += 32) hist[(uint64_t)g * n_warps + gw] = s_cnt[warp][g];
}

// hist[g][w] (bucket-major) -> exclusive prefix in (bucket, warp) order = where warp w's entries of bucket g start;
// counts[g] = bucket totals.  One CTA, 1024 threads, coalesced passes.
__global__ void __launch_bounds__(1024)
shard_bucket_scan_kernel(uint32_t* __restrict__ hist, uint32_t n_warps, int G, uint32_t region_cap,
                         uint32_t* __restrict__ counts) {
  __shared__ uint32_t wtot[32];
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0u;
  __syncthreads();
  for (int g = 0; g < G; ++g) {
    if (region_cap != 0u) {                                   // fixed-capacity regions: bucket g starts at g * cap
      __syncthreads();
      if (threadIdx.x == 0) carry_s = (uint32_t)g * region_cap;
      __syncthreads();
    }
    const uint32_t bucket_start = carry_s;
    for (uint32_t base = 0; base < n_warps; base += 1024) {
      const uint32_t i = base + threadIdx.x;
      const uint32_t v = i < n_warps ? hist[(uint64_t)g * n_warps + i] : 0u;
      uint32_t x = v;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
      }
      if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = x;
      __syncthreads();
      if (threadIdx.x < 32) {
        const uint32_t t = wtot[threadIdx.x];
        uint32_t sc = t;
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t y = __shfl_up_sync(0xffffffffu, sc, o);
          if (threadIdx.x >= o) sc += y;
        }
        wtot[threadIdx.x] = sc - t;
      }
      __syncthreads();
      const uint32_t carry = carry_s;
      const uint32_t incl = x + wtot[threadIdx.x >> 5];
      if (i < n_warps) hist[(uint64_t)g * n_warps + i] = carry + incl - v;
      __syncthreads();
      if (threadIdx.x == 1023) carry_s = carry + incl;
      __syncthreads();
    }
    if (threadIdx.x == 0) counts[g] = carry_s - bucket_start;
  }
}

template <typename IdxT>
__global__ void __launch_bounds__(SB_THREADS)
shard_bucket_scatter_kernel(const IdxT* __restrict__ idx, uint64_t n, uint64_t chunk, uint64_t first_slot, OwnerOf own,
                            int G, uint32_t n_warps, const uint32_t* __restrict__ start /*[G][n_warps]*/,
                            uint32_t* __restrict__ slots_out, uint32_t* __restrict__ ids_out /* NULL: packed pairs */) {
  __shared__ uint32_t s_run[SB_WARPS][SB_MAX_RANKS];            // next output position of every bucket, per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * SB_WARPS + warp;
  if (gw >= n_warps) return;
  for (int g = lane; g < G; g += 32) s_run[warp][g] = start[(uint64_t)g * n_warps + gw];
  __syncwarp();
  const uint64_t q0 = (uint64_t)gw * chunk, q1 = min(n, q0 + chunk);
  for (uint64_t t0 = q0; t0 < q1; t0 += 32) {
    const uint64_t q = t0 + lane;
    const bool have = q < q1;
    uint32_t owner = 0xFFFFFFFFu, c_loc = 0;
    if (have) own.decode((uint64_t)__ldcs(idx + q), owner, c_loc);
    const uint32_t peers = __match_any_sync(0xffffffffu, owner);
    uint32_t pos = 0;
    if (have) pos = s_run[warp][owner] + (uint32_t)__popc(peers & ((1u << lane) - 1u));
    __syncwarp();
    if (have) {
      if (ids_out == nullptr) {
        reinterpret_cast<uint2*>(slots_out)[pos] = make_uint2((uint32_t)(first_slot + q), c_loc);
      } else {
        slots_out[pos] = (uint32_t)(first_slot + q);
        ids_out[pos] = c_loc;
      }
      if ((peers >> lane) == 1u) s_run[warp][owner] = pos + 1u;   // the highest peer lane advances the bucket
    }
    __syncwarp();
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
  if (ld % 4 == 0 && ld / 4 <= 256 && (((uintptr_t)in | (uintptr_t)out) & 15) == 0) {
    const int vpr = ld / 4, rpi = 256 / vpr;
    int grid = (int)min((n_rows + rpi - 1) / rpi, (uint64_t)sm_count() * 8);
    scatter_rows_vec4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)in, vpr, rpi, slots, slot_offset, n_rows,
                                                                    (float4*)out);
    count_launches(1);
    return check_launch("scatter_rows");
  }
  const int lanes = ld < 32 ? ld : 32;
  const int rows_per_cta = 256 / lanes;
  int grid = (int)min((n_rows + rows_per_cta - 1) / rows_per_cta, (uint64_t)sm_count() * 16);
  scatter_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, ld, slots, slot_offset, n_rows, out);
  count_launches(1);
  return check_launch("scatter_rows");
}

// slots[q] = pairs[q].x, ids[q] = pairs[q].y  (the receive side of the id exchange: one packed all-to-all)
__global__ void __launch_bounds__(256)
unzip_pairs_kernel(const uint2* __restrict__ pairs, uint64_t n, uint32_t* __restrict__ slots, uint32_t* __restrict__ ids) {
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (uint64_t)gridDim.x * blockDim.x) {
    const uint2 v = __ldcs(pairs + q);
    slots[q] = v.x;
    if (ids) ids[q] = v.y;
  }
}

extern "C" int mccb200_unzip_pairs(const uint32_t* pairs, uint64_t n, uint32_t* slots, uint32_t* ids, void* stream) {
  MCCB_REQUIRE((pairs && slots) || n == 0, "unzip_pairs: NULL argument");   // ids may be NULL (slots only)
  if (n == 0) return MCCB200_OK;
  const int grid = (int)min((n + 255) / 256, (uint64_t)mccb::sm_count() * 8);
  unzip_pairs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint2*)pairs, n, slots, ids);
  mccb::count_launches(1);
  return mccb::check_launch("unzip_pairs");
}

static uint32_t sb_warps(uint64_t n, uint64_t& chunk) {
  uint64_t warps = (uint64_t)sm_count() * (side_ctas() > 0 ? side_ctas() : 4) * SB_WARPS;
  chunk = (n + warps - 1) / warps;
  chunk = (chunk + 31) / 32 * 32;
  if (chunk == 0) chunk = 32;
  return (uint32_t)((n + chunk - 1) / chunk);
}

extern "C" size_t mccb200_shard_bucket_ids_workspace_bytes(int n_ranks) {
  if (n_ranks < 1 || n_ranks > SB_MAX_RANKS) return 0;
  return ((size_t)sm_count() * 4 * SB_WARPS + 1) * (size_t)n_ranks * sizeof(uint32_t);
}

extern "C" int mccb200_shard_bucket_ids_regions(const void* idx, int idx_is_64, uint64_t n, uint64_t first_slot,
                                                uint64_t n_loc, int k, int n_ranks, uint64_t region_cap,
                                                uint32_t* slots_out, uint32_t* ids_out, uint32_t* counts_out,
                                                void* workspace, size_t workspace_bytes, void* stream);

/* The send buffers of the id exchange (`ShardedMonteCarloStep.step_owner_ids`): `idx` = this rank's slice of the
 * global index vector (child ids, uint32 or uint64), output slot of entry q = first_slot + q.  slots_out[n] /
 * ids_out[n] = (slot, child id relative to the owner's first parent) grouped by owner rank = child / (n_loc * k),
 * ascending slot inside a group; counts_out[n_ranks] = group sizes.  n_loc * k <= 2^32, slots < 2^32. */
extern "C" int mccb200_shard_bucket_ids(const void* idx, int idx_is_64, uint64_t n, uint64_t first_slot, uint64_t n_loc,
                                        int k, int n_ranks, uint32_t* slots_out, uint32_t* ids_out, uint32_t* counts_out,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  return mccb200_shard_bucket_ids_regions(idx, idx_is_64, n, first_slot, n_loc, k, n_ranks, 0, slots_out, ids_out,
                                          counts_out, workspace, workspace_bytes, stream);
}

/* The same with every owner's group written to its own fixed-capacity region: group g starts at entry g * region_cap
 * (region_cap = 0: groups back to back).  A group larger than region_cap spills into the next region: the caller must
 * check counts_out[g] <= region_cap before using the regions (and fall back to the back-to-back form otherwise). */
extern "C" int mccb200_shard_bucket_ids_regions(const void* idx, int idx_is_64, uint64_t n, uint64_t first_slot,
                                                uint64_t n_loc, int k, int n_ranks, uint64_t region_cap,
                                                uint32_t* slots_out, uint32_t* ids_out, uint32_t* counts_out,
                                                void* workspace, size_t workspace_bytes, void* stream) {
  MCCB_REQUIRE(idx && slots_out && counts_out && workspace, "shard_bucket_ids: NULL argument");
  MCCB_REQUIRE(region_cap * (uint64_t)n_ranks < 0x100000000ull, "shard_bucket_ids: regions exceed 32-bit positions");
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
  const uint32_t n_warps = sb_warps(n, chunk);
  const int grid = (int)((n_warps + SB_WARPS - 1) / SB_WARPS);
  OwnerOf own;
  own.cpr = n_loc * (uint64_t)k;
  own.shift = -1;
  if ((own.cpr & (own.cpr - 1)) == 0) { own.shift = 0; while ((1ull << own.shift) < own.cpr) ++own.shift; }
  uint32_t* hist = (uint32_t*)workspace;
  if (idx_is_64) {
    shard_bucket_count_kernel<uint64_t><<<grid, SB_THREADS, 0, st>>>((const uint64_t*)idx, n, chunk, own, n_ranks, n_warps, hist);
    shard_bucket_scan_kernel<<<1, 1024, 0, st>>>(hist, n_warps, n_ranks, (uint32_t)region_cap, counts_out);
    shard_bucket_scatter_kernel<uint64_t><<<grid, SB_THREADS, 0, st>>>((const uint64_t*)idx, n, chunk, first_slot, own,
                                                                        n_ranks, n_warps, hist, slots_out, ids_out);
  } else {
    shard_bucket_count_kernel<uint32_t><<<grid, SB_THREADS, 0, st>>>((const uint32_t*)idx, n, chunk, own, n_ranks, n_warps, hist);
    shard_bucket_scan_kernel<<<1, 1024, 0, st>>>(hist, n_warps, n_ranks, (uint32_t)region_cap, counts_out);
    shard_bucket_scatter_kernel<uint32_t><<<grid, SB_THREADS, 0, st>>>((const uint32_t*)idx, n, chunk, first_slot, own,
                                                                        n_ranks, n_warps, hist, slots_out, ids_out);
  }
  count_launches(3);
  return check_launch("shard_bucket_ids");
}

/* Device-to-device copy on `stream`, `dst` possibly a peer mapping (mccb200_peer_open): one cudaMemcpyAsync, served by
 * the copy engines over NVLink, so it runs under kernels that occupy every SM. */
extern "C" int mccb200_peer_copy(void* dst, const void* src, size_t bytes, void* stream) {
  MCCB_REQUIRE((dst && src) || bytes == 0, "peer_copy: NULL argument");
  if (bytes == 0) return MCCB200_OK;
  if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) != cudaSuccess) {
    cudaGetLastError();
    return fail(MCCB200_ERR_CUDA, "peer_copy: cudaMemcpyAsync failed");
  }
  count_launches(1);
  return MCCB200_OK;
}
