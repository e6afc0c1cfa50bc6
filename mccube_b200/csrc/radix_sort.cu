// Stable LSD radix sort of (uint32 key, uint32 value) pairs -- hand-written for sm_100a.
//
// Replaces XLA's `lax.sort_key_val(keys, x, is_stable=True)` inside jax.random._shuffle
// (reached from /root/reference/mccube/_kernels/random.py:65) and `jnp.argsort`
// (/root/reference/mccube/_kernels/stratified.py:43).
//
// 8-bit digits.  Default form, per pass:
//   1. rs_histogram : per-tile digit counts -> table[digit][tile]           (reads keys)
//   2. rs_scan_*    : exclusive scan of the digit-major table               (tiny)
//   3. rs_scatter   : per-tile stable ranking with ballot peer masks + per-warp counters,
//                     re-ordering of the tile in shared memory, then digit-run-coalesced
//                     stores of keys and values.
// "Onesweep" form (MCCB200_SORT_MODE=1): one kernel reads the keys once and builds the GLOBAL
// histogram of every digit position (order independent), then ONE kernel per pass ranks and
// scatters: a tile's offset inside a digit = sum of that digit's counts in all earlier tiles, found
// by decoupled look-back over per-(tile, digit) status words (aggregate / inclusive prefix, 64-bit so
// flag and value travel together), tiles handed out by an atomic ticket so that every predecessor
// of a running tile is itself running or done.  Bit-identical results, 7 launches instead of 20 per
// sort, but MEASURED SLOWER at 2^25 pairs (3.98 vs 3.69 ms for the 3-round shuffle): ncu shows the
// scatter kernel is issue-bound (154 M warp-instructions per pass, 37 % of them the 8 ballots per
// element that build the peer masks; 2 TB/s of DRAM), so removing the 39 us histogram pass does not
// pay for the ticket + look-back latency inside every tile.  Left switchable for the next round.
// Tile order (and therefore stability) is: tile, warp, iteration, lane.
#include <stdlib.h>

#include "common.cuh"

namespace mccb {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096
constexpr int RS_RADIX = 256;
constexpr int SCAN_CHUNK = 4096;  // table entries per CTA in the scan kernels
// peer masks of equal digits: 8 ballots (0) or one MATCH.ANY (1).  Measured on B200 at 2^25 random keys, 3-round
// shuffle: ballots 3.69 ms, MATCH.ANY 5.20 ms (it serialises over the ~28 distinct digits a warp holds).
#ifndef RS_USE_MATCH
#define RS_USE_MATCH 0
#endif

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
rs_scan_reduce(const uint32_t* __restrict__ table, uint64_t len, uint32_t* __restrict__ sums,
               const uint32_t* __restrict__ guard = nullptr) {
  if (guard && *guard == 0u) return;               // guarded launch: see rng_select.cu
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
rs_scan_sums(uint32_t* __restrict__ sums, uint32_t n_sums, const uint32_t* __restrict__ guard = nullptr) {
  if (guard && *guard == 0u) return;
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
rs_scan_apply(uint32_t* __restrict__ table, uint64_t len, const uint32_t* __restrict__ sums,
              const uint32_t* __restrict__ guard = nullptr) {
  if (guard && *guard == 0u) return;
  // exclusive scan inside a chunk of SCAN_CHUNK = 256 threads x 16 consecutive entries
  __shared__ uint32_t warp_tot[8];
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_CHUNK + (uint64_t)threadIdx.x * 16;
  uint32_t v[16];
  uint32_t tsum = 0;
  const bool vec = base + 16 <= len && (((uintptr_t)table) & 15) == 0;   // base is a multiple of 16 entries
  if (vec) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 t = reinterpret_cast<const uint4*>(table + base)[i];
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      uint64_t g = base + i;
      v[i] = g < len ? table[g] : 0u;
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) tsum += v[i];
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
  if (vec) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 t;
      t.x = run; run += v[4 * i];
      t.y = run; run += v[4 * i + 1];
      t.z = run; run += v[4 * i + 2];
      t.w = run; run += v[4 * i + 3];
      reinterpret_cast<uint4*>(table + base)[i] = t;
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    uint64_t g = base + i;
    if (g < len) table[g] = run;
    run += v[i];
  }
}

// ---- onesweep support
constexpr uint64_t OS_AGG = 1ull << 62, OS_PRE = 2ull << 62, OS_VALUE = (1ull << 56) - 1;   // [63:62] state, [61:56] pass tag

// global digit histograms of all digit positions in one read of the keys: ghist[pass][digit]
__global__ void __launch_bounds__(256)
rs_digit_histograms(const uint32_t* __restrict__ keys, uint64_t count, int passes, uint32_t* __restrict__ ghist) {
  __shared__ uint32_t h[4][RS_RADIX];
  for (int i = threadIdx.x; i < 4 * RS_RADIX; i += 256) (&h[0][0])[i] = 0;
  __syncthreads();
  const bool aligned = (((uintptr_t)keys) & 15) == 0;
  const int lane = threadIdx.x & 31;
  for (uint64_t cb = (uint64_t)blockIdx.x * 1024; cb < count; cb += (uint64_t)gridDim.x * 1024) {   // CTA-uniform trip count
    const uint64_t g0 = cb + (uint64_t)threadIdx.x * 4;
    uint32_t kv[4] = {0u, 0u, 0u, 0u};
    int nk = 0;
    if (aligned && g0 + 3 < count) {
      const uint4 t = __ldcs(reinterpret_cast<const uint4*>(keys + g0));
      kv[0] = t.x; kv[1] = t.y; kv[2] = t.z; kv[3] = t.w; nk = 4;
    } else {
      for (; nk < 4 && g0 + nk < count; ++nk) kv[nk] = keys[g0 + nk];
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool ok = e < nk;
      for (int ps = 0; ps < passes; ++ps) {
        const uint32_t dg = ok ? ((kv[e] >> (8 * ps)) & 0xFFu) : 0xFFFFFFFFu;
        int same;
        __match_all_sync(0xffffffffu, dg, &same);        // skewed keys (box keys, high digits): one add per warp
        if (same) { if (ok && lane == 0) atomicAdd(&h[ps][dg], 32u); }
        else if (ok) atomicAdd(&h[ps][dg], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * RS_RADIX; i += 256)
    if ((&h[0][0])[i]) atomicAdd(&ghist[i], (&h[0][0])[i]);
}

// exclusive scan over the 256 digits of every pass (one warp-scan CTA; 32-bit: count < 2^32)
__global__ void __launch_bounds__(RS_RADIX)
rs_digit_scan(uint32_t* __restrict__ ghist, int passes) {
  __shared__ uint32_t wt[8];
  for (int ps = 0; ps < passes; ++ps) {
    const uint32_t v = ghist[ps * RS_RADIX + threadIdx.x];
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wt[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t off = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) off += wt[w];
    ghist[ps * RS_RADIX + threadIdx.x] = off + x - v;
    __syncthreads();
  }
}

__device__ __forceinline__ uint64_t ld_status(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(uint64_t* p, uint64_t v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// LOOKBACK = false: `table` holds the scanned per-(digit, tile) offsets (earlier form).
// LOOKBACK = true : `table` = digit bases of this pass [256]; status / ticket as described on top.
template <bool IOTA_VALUES, bool WRITE_KEYS, bool LOOKBACK>
__global__ void __launch_bounds__(RS_THREADS, 3)
rs_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t count, int shift,
           uint32_t num_tiles, const uint32_t* __restrict__ table, uint32_t* __restrict__ keys_out,
           uint32_t* __restrict__ vals_out, uint64_t* __restrict__ status, uint32_t* __restrict__ ticket,
           uint32_t pass_tag) {
  __shared__ uint32_t s_ticket;
  __shared__ uint32_t s_keys[RS_TILE];
  __shared__ uint32_t s_vals[RS_TILE];
  __shared__ uint32_t warp_cnt[RS_WARPS][RS_RADIX];
  __shared__ uint32_t lstart[RS_RADIX];
  __shared__ uint32_t goff[RS_RADIX];
  __shared__ uint32_t scan_tmp[RS_WARPS];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;

  for (uint32_t tile_static = blockIdx.x;; tile_static += gridDim.x) {
    uint32_t tile = tile_static;
    if (LOOKBACK) {       // tickets: every predecessor of this tile has been taken by a CTA that is running or done
      if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
      __syncthreads();
      tile = s_ticket;
    }
    if (tile >= num_tiles) break;
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
      uint32_t m;
      if (RS_USE_MATCH) {
        m = __match_any_sync(0xffffffffu, ok ? dg : (0x100u + (uint32_t)lane));
      } else {
        m = __ballot_sync(0xffffffffu, ok);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const uint32_t vote = __ballot_sync(0xffffffffu, (dg >> b) & 1u);
          m &= ((dg >> b) & 1u) ? vote : ~vote;
        }
        if (!ok) m = 1u << lane;
      }
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
    uint32_t my_run = 0;
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
      my_run = run;
      if (!LOOKBACK) {
        goff[dg] = table[(uint64_t)dg * num_tiles + tile];
      } else {
        // publish this tile's digit count right away; the look-back itself runs after the shared-memory
        // re-ordering below, which gives the predecessors time to resolve their own prefixes
        const uint64_t tag = (uint64_t)pass_tag << 56;
        if (tile == 0) st_status(status + dg, OS_PRE | tag | (uint64_t)run);
        else st_status(status + (uint64_t)tile * RS_RADIX + dg, OS_AGG | tag | (uint64_t)run);
      }
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
    if (LOOKBACK) {
      // decoupled look-back: this digit's count in all earlier tiles
      const int dg = threadIdx.x;
      const uint64_t tag = (uint64_t)pass_tag << 56;
      uint64_t excl = 0;
      if (tile != 0) {
        long long t0 = 0;
        uint32_t spin = 0;
        for (int64_t tt = (int64_t)tile - 1;;) {
          const uint64_t wv = ld_status(status + (uint64_t)tt * RS_RADIX + dg);
          if ((wv >> 62) == 0 || ((wv >> 56) & 0x3Full) != pass_tag) {      // predecessor not there yet
            if ((++spin & 1023u) == 0) {
              if (t0 == 0) t0 = clock64();
              else if (clock64() - t0 > 4000000000ll) __trap();              // a protocol bug, not a slow tile
            }
            continue;
          }
          excl += wv & OS_VALUE;
          if ((wv >> 62) == 2) break;
          --tt;
        }
        st_status(status + (uint64_t)tile * RS_RADIX + dg, OS_PRE | tag | (excl + (uint64_t)my_run));
      }
      goff[dg] = table[dg] + (uint32_t)excl;
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
  size_t off_tmp_keys, off_tmp_vals, off_alt_keys, off_table, off_sums, off_ctl, total;
};
constexpr size_t OS_CTL_BYTES = 4 * RS_RADIX * 4 + 256;   // ghist[4][256] + tickets[4] (padded)

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
  p.off_table = p.off_alt_keys + arr;                     // earlier form: uint32 table; onesweep: uint64 status words
  p.off_sums = p.off_table + align_up((size_t)p.table_len * 8, 256);
  p.off_ctl = p.off_sums + align_up((size_t)p.n_chunks * 4, 256);
  p.total = p.off_ctl + OS_CTL_BYTES;
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
  static const bool onesweep = []() { const char* e = getenv("MCCB200_SORT_MODE"); return e && atoi(e) == 1; }();
  if (onesweep) {
    uint64_t* status = (uint64_t*)table;
    uint32_t* ghist = (uint32_t*)(w + p.off_ctl);
    uint32_t* tickets = ghist + 4 * RS_RADIX;
    // status words and control block start from zero (stale tags of an earlier call must not match)
    if (cudaMemsetAsync(status, 0, (size_t)p.table_len * 8, st) != cudaSuccess ||
        cudaMemsetAsync(ghist, 0, OS_CTL_BYTES, st) != cudaSuccess)
      return fail(MCCB200_ERR_CUDA, "sort_pairs: memset failed");
    const int hgrid = (int)min((count + 1023) / 1024, (uint64_t)sm_count() * 8);
    rs_digit_histograms<<<hgrid, 256, 0, st>>>(keys_in, count, passes, ghist);
    rs_digit_scan<<<1, RS_RADIX, 0, st>>>(ghist, passes);
    for (int pass = 0; pass < passes; ++pass) {
      const bool to_out = ((passes - 1 - pass) % 2) == 0;
      uint32_t* dst_k = to_out ? out_keys : tmp_keys;
      uint32_t* dst_v = to_out ? vals_out : tmp_vals;
      const int shift = pass * 8;
      const bool last = pass == passes - 1;
      const bool iota = (src_v == nullptr);
      const bool write_keys = !(last && keys_out == nullptr);
      const uint32_t* dbase = ghist + pass * RS_RADIX;
      uint32_t* tk = tickets + pass;
      const uint32_t tag = (uint32_t)pass + 1u;
#define MCCB_OS(I, WK) rs_scatter<I, WK, true><<<grid, RS_THREADS, 0, st>>>(src_k, src_v, count, shift, p.num_tiles, dbase, dst_k, dst_v, status, tk, tag)
      if (iota) { if (write_keys) MCCB_OS(true, true); else MCCB_OS(true, false); }
      else { if (write_keys) MCCB_OS(false, true); else MCCB_OS(false, false); }
#undef MCCB_OS
      src_k = dst_k;
      src_v = dst_v;
    }
    phase_mark("end", st);
    count_launches(4 + passes);
    return check_launch("sort_pairs");
  }
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
      if (write_keys) rs_scatter<true, true, false><<<grid, RS_THREADS, 0, st>>>(src_k, nullptr, count, shift, p.num_tiles, table, dst_k, dst_v, nullptr, nullptr, 0u);
      else rs_scatter<true, false, false><<<grid, RS_THREADS, 0, st>>>(src_k, nullptr, count, shift, p.num_tiles, table, dst_k, dst_v, nullptr, nullptr, 0u);
    } else {
      if (write_keys) rs_scatter<false, true, false><<<grid, RS_THREADS, 0, st>>>(src_k, src_v, count, shift, p.num_tiles, table, dst_k, dst_v, nullptr, nullptr, 0u);
      else rs_scatter<false, false, false><<<grid, RS_THREADS, 0, st>>>(src_k, src_v, count, shift, p.num_tiles, table, dst_k, dst_v, nullptr, nullptr, 0u);
    }
    src_k = dst_k;
    src_v = dst_v;
  }
  phase_mark("end", st);
  count_launches(5 * passes);
  return check_launch("sort_pairs");
}

size_t sort_pairs_ws(uint64_t count) { return make_plan(count).total; }

// In-place exclusive scan of `len` uint32 entries (the three scan kernels above); `sums` = ceil(len / SCAN_CHUNK) words.
size_t scan_u32_ws(uint64_t len) { return align_up(((len + SCAN_CHUNK - 1) / SCAN_CHUNK + 1) * sizeof(uint32_t), 256); }
// `guard`: nullptr, or a device flag that must be non-zero for the scan to happen.
int scan_u32_exclusive(uint32_t* data, uint64_t len, void* ws, cudaStream_t st, const uint32_t* guard) {
  if (len == 0) return MCCB200_OK;
  const uint32_t n_chunks = (uint32_t)((len + SCAN_CHUNK - 1) / SCAN_CHUNK);
  uint32_t* sums = (uint32_t*)ws;
  rs_scan_reduce<<<n_chunks, 256, 0, st>>>(data, len, sums, guard);
  rs_scan_sums<<<1, 1024, 0, st>>>(sums, n_chunks, guard);
  rs_scan_apply<<<n_chunks, 256, 0, st>>>(data, len, sums, guard);
  count_launches(3);
  return check_launch("scan_u32");
}

// ---- first `top` entries of the stable sort, top << count (the last round of jax.random._shuffle
// feeds permutation(...)[:n]): one stable pass on the MOST significant byte over everything, then
// a full stable LSD sort of a prefix of `u` elements that contains the `top` smallest.  u = top +
// two average bucket sizes: for the uniform 32-bit keys of the shuffle the bucket holding rank top-1
// ends inside that prefix with overwhelming probability; a device-side check traps otherwise (a loud
// failure instead of a wrong permutation).
__global__ void __launch_bounds__(RS_RADIX)
rs_top_check(const uint32_t* __restrict__ table, uint32_t num_tiles, uint64_t count, uint64_t top, uint64_t u) {
  __shared__ uint32_t base[RS_RADIX + 1];
  base[threadIdx.x] = table[(uint64_t)threadIdx.x * num_tiles];   // scanned table: start of digit = offset of its tile 0
  if (threadIdx.x == 0) base[RS_RADIX] = (uint32_t)count;
  __syncthreads();
  // the bucket that holds rank top-1 must end inside the prefix
  if ((uint64_t)base[threadIdx.x] <= top - 1 && top - 1 < (uint64_t)base[threadIdx.x + 1] &&
      (uint64_t)base[threadIdx.x + 1] > u)
    __trap();
}

static uint64_t top_prefix(uint64_t count, uint64_t top) {
  uint64_t u = top + 2 * ((count + RS_RADIX - 1) / RS_RADIX) + 4096;
  return u < count ? u : count;
}

bool sort_pairs_top_worthwhile(uint64_t count, uint64_t top) {
  return count >= (1ull << 20) && top_prefix(count, top) <= count / 4;
}

size_t sort_pairs_top_ws(uint64_t count, uint64_t top) {
  SortPlan p = make_plan(count);
  return p.total + 2 * align_up((size_t)count * 4, 256) + make_plan(top_prefix(count, top)).total;
}

// vals_out must hold top_prefix(count, top) entries; its first `top` are the result.
int sort_pairs_top_impl(const uint32_t* keys_in, const uint32_t* vals_in, uint64_t count, uint64_t top,
                        uint32_t* vals_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  MCCB_REQUIRE(count < 0xFFFFFFFFull && top >= 1 && top <= count, "sort_pairs_top: bad sizes");
  const uint64_t u = top_prefix(count, top);
  SortPlan p = make_plan(count);
  const size_t arr = align_up((size_t)count * 4, 256);
  if (ws_bytes < sort_pairs_top_ws(count, top))
    return fail(MCCB200_ERR_WORKSPACE_TOO_SMALL, "sort_pairs_top: workspace %zu < %zu", ws_bytes, sort_pairs_top_ws(count, top));
  char* w = (char*)ws;
  uint32_t* table = (uint32_t*)(w + p.off_table);
  uint32_t* sums = (uint32_t*)(w + p.off_sums);
  uint32_t* msd_keys = (uint32_t*)(w + p.total);
  uint32_t* msd_vals = (uint32_t*)(w + p.total + arr);
  void* inner_ws = w + p.total + 2 * arr;
  const size_t inner_bytes = ws_bytes - (p.total + 2 * arr);
  const int grid = (int)min((uint64_t)p.num_tiles, (uint64_t)sm_count() * 8);
  phase_mark("sort_pairs", st);
  rs_histogram<<<grid, RS_THREADS, 0, st>>>(keys_in, count, 24, p.num_tiles, table);
  rs_scan_reduce<<<p.n_chunks, 256, 0, st>>>(table, p.table_len, sums);
  rs_scan_sums<<<1, 1024, 0, st>>>(sums, p.n_chunks);
  rs_scan_apply<<<p.n_chunks, 256, 0, st>>>(table, p.table_len, sums);
  rs_top_check<<<1, RS_RADIX, 0, st>>>(table, p.num_tiles, count, top, u);
  if (vals_in) rs_scatter<false, true, false><<<grid, RS_THREADS, 0, st>>>(keys_in, vals_in, count, 24, p.num_tiles, table, msd_keys, msd_vals, nullptr, nullptr, 0u);
  else rs_scatter<true, true, false><<<grid, RS_THREADS, 0, st>>>(keys_in, nullptr, count, 24, p.num_tiles, table, msd_keys, msd_vals, nullptr, nullptr, 0u);
  count_launches(6);
  // stable LSD sort of the prefix by the full key (ties keep the MSD pass's = the input's order)
  return sort_pairs_impl(msd_keys, msd_vals, u, 32, nullptr, vals_out, inner_ws, inner_bytes, st);
}

}  // namespace mccb

extern "C" size_t mccb200_sort_pairs_workspace_bytes(uint64_t count) { return mccb::sort_pairs_ws(count); }

extern "C" int mccb200_sort_pairs(const uint32_t* keys_in, const uint32_t* values_in, uint64_t count, int key_bits,
                                  uint32_t* keys_out, uint32_t* values_out, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  return mccb::sort_pairs_impl(keys_in, values_in, count, key_bits, keys_out, values_out, workspace,
                               workspace_bytes, (cudaStream_t)stream);
}
