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
