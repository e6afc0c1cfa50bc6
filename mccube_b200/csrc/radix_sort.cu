// Stable LSD radix sort of (uint32 key, uint32 value) pairs -- hand-written for sm_100a.
//
// Replaces XLA's `lax.sort_key_val(keys, x, is_stable=True)` inside jax.random._shuffle
// (reached from /root/reference/mccube/_kernels/random.py:65) and `jnp.argsort`
// (/root/reference/mccube/_kernels/stratified.py:43).
//
// 8-bit digits.  Per pass three launches:
//   1. rs_histogram : per-tile digit counts -> table[digit][tile]           (reads keys)
//   2. rs_scan_*    : exclusive scan of the digit-major table               (tiny)
//   3. rs_scatter   : per-tile stable ranking with warp match_any + per-warp counters,
//                     re-ordering of the tile in shared memory, then digit-run-coalesced
//                     stores of keys and values.
// Tile order (and therefore stability) is: tile, warp, iteration, lane.
#include "common.cuh"

namespace mccb {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096
constexpr int RS_RADIX = 256;
constexpr int SCAN_CHUNK = 4096;  // table entries per CTA in the scan kernels

__device__ __forceinline__ uint32_t tile_elem(int warp, int it, int lane) {
  return warp * (32 * RS_ITEMS) + it * 32 + lane;
}

__global__ void __launch_bounds__(RS_THREADS)
rs_histogram(const uint32_t* __restrict__ keys, uint64_t count, int shift, uint32_t num_tiles,
             uint32_t* __restrict__ table) {
  __shared__ uint32_t hist[RS_RADIX];
  for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    hist[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)tile * RS_TILE;
    uint32_t kv[RS_ITEMS];
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {   // all loads in flight before the first atomic
      uint64_t g = base + (uint64_t)it * RS_THREADS + threadIdx.x;
      kv[it] = g < count ? __ldcs(keys + g) : 0u;
    }
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
      uint64_t g = base + (uint64_t)it * RS_THREADS + threadIdx.x;
      if (g < count) atomicAdd(&hist[(kv[it] >> shift) & 0xFF], 1u);
    }
    __syncthreads();
    table[(uint64_t)threadIdx.x * num_tiles + tile] = hist[threadIdx.x];
    __syncthreads();
  }
}

// --- exclusive scan of `len` uint32 entries: chunk sums -> scan of sums -> apply.
__global__ void __launch_bounds__(256)
rs_scan_reduce(const uint32_t* __restrict__ table, uint64_t len, uint32_t* __restrict__ sums) {
  __shared__ uint32_t red[8];
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_CHUNK;
  uint32_t acc = 0;
  for (int i = threadIdx.x; i < SCAN_CHUNK; i += 256) {
    uint64_t g = base + i;
    if (g < len) acc += table[g];
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = 0;
    for (int w = 0; w < 8; ++w) s += red[w];
    sums[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(1024)
rs_scan_sums(uint32_t* __restrict__ sums, uint32_t n_sums) {
  // single CTA, sequential over 1024-wide slabs
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_sums; base += 1024) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < n_sums ? sums[i] : 0u;
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t t = warp_tot[threadIdx.x];
      uint32_t s = t;
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
        if (threadIdx.x >= o) s += y;
      }
      warp_tot[threadIdx.x] = s - t;  // exclusive
    }
    __syncthreads();
    uint32_t carry = carry_s;
    uint32_t incl = x + warp_tot[threadIdx.x >> 5];
    if (i < n_sums) sums[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + incl;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
rs_scan_apply(uint32_t* __restrict__ table, uint64_t len, const uint32_t* __restrict__ sums) {
  // exclusive scan inside a chunk of SCAN_CHUNK = 256 threads x 16 consecutive entries
  __shared__ uint32_t warp_tot[8];
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_CHUNK + (uint64_t)threadIdx.x * 16;
  uint32_t v[16];
  uint32_t tsum = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    uint64_t g = base + i;
    v[i] = g < len ? table[g] : 0u;
    tsum += v[i];
  }
  uint32_t x = tsum;
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
  __syncthreads();
  uint32_t woff = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += warp_tot[w];
  uint32_t run = sums[blockIdx.x] + woff + x - tsum;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    uint64_t g = base + i;
    if (g < len) table[g] = run;
    run += v[i];
  }
}

template <bool IOTA_VALUES, bool WRITE_KEYS>
__global__ void __launch_bounds__(RS_THREADS, 3)
rs_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t count, int shift,
           uint32_t num_tiles, const uint32_t* __restrict__ table, uint32_t* __restrict__ keys_out,
           uint32_t* __restrict__ vals_out) {
  __shared__ uint32_t s_keys[RS_TILE];
  __shared__ uint32_t s_vals[RS_TILE];
  __shared__ uint32_t warp_cnt[RS_WARPS][RS_RADIX];
  __shared__ uint32_t lstart[RS_RADIX];
  __shared__ uint32_t goff[RS_RADIX];
  __shared__ uint32_t scan_tmp[RS_WARPS];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;

  for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const uint64_t base = (uint64_t)tile * RS_TILE;
    for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&warp_cnt[0][0])[i] = 0;
    __syncthreads();

    // Ranking without a serial read-modify-write chain: per item one MATCH.ANY, then the group
    // leader issues a shared-memory atomicAdd that returns the running count of its digit in this
    // warp.  Atomics of one warp execute in issue order, so the 16 of them pipeline; the returned
    // values are only collected in the second loop.
    uint32_t key[RS_ITEMS], old[RS_ITEMS];
    uint32_t pl[RS_ITEMS / 2];   // (prefix within group | leader lane << 8), two per word
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
      uint64_t g = base + tile_elem(warp, it, lane);
      key[it] = g < count ? __ldcs(keys_in + g) : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
      uint64_t g = base + tile_elem(warp, it, lane);
      const bool ok = g < count;
      const uint32_t dg = (key[it] >> shift) & 0xFF;
      // peer mask of equal digits from 8 ballots (fixed cost; MATCH.ANY serialises over the ~30
      // distinct digit values a warp of random keys holds)
      uint32_t m = __ballot_sync(0xffffffffu, ok);
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const uint32_t vote = __ballot_sync(0xffffffffu, (dg >> b) & 1u);
        m &= ((dg >> b) & 1u) ? vote : ~vote;
      }
      if (!ok) m = 1u << lane;
      const uint32_t leader = __ffs(m) - 1;
      old[it] = 0u;
      if (ok && leader == (uint32_t)lane) old[it] = atomicAdd(&warp_cnt[warp][dg], (uint32_t)__popc(m));
      const uint32_t v = (uint32_t)__popc(m & lt_mask) | (leader << 8);
      if (it & 1) pl[it >> 1] |= v << 16; else pl[it >> 1] = v;
    }
    uint32_t rank2[RS_ITEMS / 2];
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
      const uint32_t v = (it & 1) ? (pl[it >> 1] >> 16) : (pl[it >> 1] & 0xFFFFu);
      const uint32_t rk = __shfl_sync(0xffffffffu, old[it], v >> 8) + (v & 0xFFu);
      if (it & 1) rank2[it >> 1] |= rk << 16; else rank2[it >> 1] = rk;
    }
    __syncthreads();

    // digit totals + exclusive prefix across warps (thread t owns digit t)
    {
      const int dg = threadIdx.x;
      uint32_t run = 0;
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) {
        uint32_t c = warp_cnt[w][dg];
        warp_cnt[w][dg] = run;
        run += c;
      }
      // exclusive scan of `run` over the 256 digits
      uint32_t x = run;
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      if (lane == 31) scan_tmp[warp] = x;
      __syncthreads();
      uint32_t woff = 0;
      for (int w = 0; w < warp; ++w) woff += scan_tmp[w];
      lstart[dg] = woff + x - run;
      goff[dg] = table[(uint64_t)dg * num_tiles + tile];
    }
    __syncthreads();

#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
      uint64_t g = base + tile_elem(warp, it, lane);
      if (g < count) {
        uint32_t dg = (key[it] >> shift) & 0xFF;
        uint32_t rk = (it & 1) ? (rank2[it >> 1] >> 16) : (rank2[it >> 1] & 0xFFFFu);
        uint32_t pos = lstart[dg] + warp_cnt[warp][dg] + rk;
        s_keys[pos] = key[it];
        s_vals[pos] = IOTA_VALUES ? (uint32_t)g : __ldcs(vals_in + g);
      }
    }
    __syncthreads();

    const uint32_t tile_count = (uint32_t)((count - base) < (uint64_t)RS_TILE ? (count - base) : RS_TILE);
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
      uint32_t q = it * RS_THREADS + threadIdx.x;
      if (q < tile_count) {
        uint32_t k = s_keys[q];
        uint32_t dg = (k >> shift) & 0xFF;
        uint64_t dst = (uint64_t)goff[dg] + (q - lstart[dg]);
        if (WRITE_KEYS) keys_out[dst] = k;
        vals_out[dst] = s_vals[q];
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ host side
struct SortPlan {
  uint32_t num_tiles;
  uint64_t table_len;
  uint32_t n_chunks;
  size_t off_tmp_keys, off_tmp_vals, off_alt_keys, off_table, off_sums, total;
};

static SortPlan make_plan(uint64_t count) {
  SortPlan p;
  p.num_tiles = (uint32_t)((count + RS_TILE - 1) / RS_TILE);
  if (p.num_tiles == 0) p.num_tiles = 1;
  p.table_len = (uint64_t)p.num_tiles * RS_RADIX;
  p.n_chunks = (uint32_t)((p.table_len + SCAN_CHUNK - 1) / SCAN_CHUNK);
  size_t arr = align_up((size_t)count * 4, 256);
  p.off_tmp_keys = 0;
  p.off_tmp_vals = p.off_tmp_keys + arr;
  p.off_alt_keys = p.off_tmp_vals + arr;
  p.off_table = p.off_alt_keys + arr;
  p.off_sums = p.off_table + align_up((size_t)p.table_len * 4, 256);
  p.total = p.off_sums + align_up((size_t)p.n_chunks * 4, 256);
  return p;
}

int sort_pairs_impl(const uint32_t* keys_in, const uint32_t* vals_in, uint64_t count, int key_bits,
                    uint32_t* keys_out, uint32_t* vals_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  MCCB_REQUIRE(count < 0xFFFFFFFFull, "sort_pairs: count %llu >= 2^32-1", (unsigned long long)count);
  MCCB_REQUIRE(key_bits >= 0 && key_bits <= 32, "sort_pairs: key_bits %d out of range", key_bits);
  MCCB_REQUIRE(vals_out != nullptr, "sort_pairs: values_out is NULL");
  if (count == 0) return MCCB200_OK;
  SortPlan p = make_plan(count);
  if (ws_bytes < p.total)
    return fail(MCCB200_ERR_WORKSPACE_TOO_SMALL, "sort_pairs: workspace %zu < %zu", ws_bytes, p.total);
  char* w = (char*)ws;
  uint32_t* tmp_keys = (uint32_t*)(w + p.off_tmp_keys);
  uint32_t* tmp_vals = (uint32_t*)(w + p.off_tmp_vals);
  uint32_t* out_keys = keys_out ? keys_out : (uint32_t*)(w + p.off_alt_keys);
  uint32_t* table = (uint32_t*)(w + p.off_table);
  uint32_t* sums = (uint32_t*)(w + p.off_sums);

  int passes = (key_bits + 7) / 8;
  if (passes == 0) passes = 1;  // still produces a copy (all digits 0 -> identity permutation)
  const int grid = (int)min((uint64_t)p.num_tiles, (uint64_t)sm_count() * 8);

  phase_mark("sort_pairs", st);
  const uint32_t* src_k = keys_in;
  const uint32_t* src_v = vals_in;
  for (int pass = 0; pass < passes; ++pass) {
    const bool to_out = ((passes - 1 - pass) % 2) == 0;
    uint32_t* dst_k = to_out ? out_keys : tmp_keys;
    uint32_t* dst_v = to_out ? vals_out : tmp_vals;
    const int shift = pass * 8;
    const bool last = pass == passes - 1;
    rs_histogram<<<grid, RS_THREADS, 0, st>>>(src_k, count, shift, p.num_tiles, table);
    rs_scan_reduce<<<p.n_chunks, 256, 0, st>>>(table, p.table_len, sums);
    rs_scan_sums<<<1, 1024, 0, st>>>(sums, p.n_chunks);
    rs_scan_apply<<<p.n_chunks, 256, 0, st>>>(table, p.table_len, sums);
    const bool iota = (src_v == nullptr);
    const bool write_keys = !(last && keys_out == nullptr);
    if (iota) {
      if (write_keys) rs_scatter<true, true><<<grid, RS_THREADS, 0, st>>>(src_k, nullptr, count, shift, p.num_tiles, table, dst_k, dst_v);
      else rs_scatter<true, false><<<grid, RS_THREADS, 0, st>>>(src_k, nullptr, count, shift, p.num_tiles, table, dst_k, dst_v);
    } else {
      if (write_keys) rs_scatter<false, true><<<grid, RS_THREADS, 0, st>>>(src_k, src_v, count, shift, p.num_tiles, table, dst_k, dst_v);
      else rs_scatter<false, false><<<grid, RS_THREADS, 0, st>>>(src_k, src_v, count, shift, p.num_tiles, table, dst_k, dst_v);
    }
    src_k = dst_k;
    src_v = dst_v;
  }
  phase_mark("end", st);
  count_launches(5 * passes);
  return check_launch("sort_pairs");
}

size_t sort_pairs_ws(uint64_t count) { return make_plan(count).total; }

}  // namespace mccb

extern "C" size_t mccb200_sort_pairs_workspace_bytes(uint64_t count) { return mccb::sort_pairs_ws(count); }

extern "C" int mccb200_sort_pairs(const uint32_t* keys_in, const uint32_t* values_in, uint64_t count, int key_bits,
                                  uint32_t* keys_out, uint32_t* values_out, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  return mccb::sort_pairs_impl(keys_in, values_in, count, key_bits, keys_out, values_out, workspace,
                               workspace_bytes, (cudaStream_t)stream);
}
