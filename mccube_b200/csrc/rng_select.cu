// jax.random._shuffle by rank selection: only permutation[:count] is ever used (sm_100a).
//
// Replaces, for `MonteCarloKernel(with_replacement=False)` -- the reference's default,
// /root/reference/mccube/_kernels/random.py:45,65 -> jr.choice -> permutation -> _shuffle --
// `rounds` x (random_bits + stable sort_key_val of ALL n inputs) when count << n (count = n / k for an MCC step).
//
// _shuffle: x_0 = arange(n); x_r = x_{r-1}[a_r] with a_r = stable argsort of the round's fresh 32-bit keys
// (keys attach to POSITIONS).  Hence  x_R[q] = a_1[a_2[... a_R[q]]]:  a composition of argsorts evaluated at `count`
// points.  a_r[i] = "which position holds the key of rank i in round r" is a rank-SELECT query, and for uniform keys it
// is cheap: bucket the keys by their top B bits (about 16 keys per bucket), count the buckets (pass 1: the keys are
// regenerated from the counter-based threefry, never stored), scan; locate every queried rank in its bucket by binary
// search on the scanned counts; scatter only the members (key, position) of buckets that hold a query (pass 2); and
// one warp per query ranks the bucket's members by (key, position) -- ties keep position order, as the stable sort
// does -- to find the member with exactly r smaller ones.  Rounds are processed from the last to the first, the
// answers of one round being the queries of the one before.  Work per round: two threefry sweeps over n and
// O(count) selection work, against four radix passes over n pairs; bit-identical to the sort (tests: against the
// oracle's jax.random restatement and against the radix path at 2^25 inputs).
#include "common.cuh"
#include "rng.cuh"

namespace mccb {

size_t scan_u32_ws(uint64_t len);
int scan_u32_exclusive(uint32_t* data, uint64_t len, void* ws, cudaStream_t st, const uint32_t* guard);

struct SelShape {
  uint32_t window;   // half-width (buckets) of the interpolation-search window: ~6 sigma of a rank's bucket
  uint32_t n;        // inputs
  uint32_t count;    // outputs wanted
  int bucket_bits;   // B
  int shift;         // 32 - B
  uint32_t nb;       // 2^B
};

static SelShape sel_shape(uint64_t n_inputs, uint64_t count) {
  SelShape s{};
  s.n = (uint32_t)n_inputs;
  s.count = (uint32_t)count;
  int lg = 0;
  while ((2ull << lg) <= n_inputs) ++lg;          // floor(log2 n)
  s.bucket_bits = lg - 3;                         // 8 .. 16 keys per bucket on average
  if (s.bucket_bits < 4) s.bucket_bits = 4;
  if (s.bucket_bits > 28) s.bucket_bits = 28;          // 2^28 buckets: three 1 GiB tables at 2^31 inputs
  s.shift = 32 - s.bucket_bits;
  s.nb = 1u << s.bucket_bits;
  // the bucket of rank i is i * nb / n up to a deviation of sigma <= sqrt(n) / 2 ranks
  const double sigma_buckets = 0.5 * sqrt((double)n_inputs) * (double)s.nb / (double)n_inputs;
  s.window = 2048;
  while ((double)s.window < 6.0 * sigma_buckets) s.window <<= 1;
  return s;
}

struct SelLayout { size_t off_hist[2], off_scan[2], off_cursor, off_bitmap, off_members, off_qs, off_qc, off_qr, off_q0, off_q1, off_flag, total; };

static SelLayout sel_layout(const SelShape& s) {
  SelLayout L{};
  size_t cur = 0;
  auto take = [&](size_t bytes) { size_t o = cur; cur += align_up(bytes, 256); return o; };
  for (int i = 0; i < 2; ++i) {                   // rounds alternate: the next round's counts are built on a side stream
    L.off_hist[i] = take(((size_t)s.nb + 1) * 4);
    L.off_scan[i] = take(scan_u32_ws((uint64_t)s.nb + 1));
  }
  L.off_cursor = take((size_t)s.nb * 4);
  L.off_bitmap = take((size_t)(s.nb + 31) / 32 * 4);
  L.off_members = take((size_t)s.n * 8);          // (key, position) of the members of queried buckets
  L.off_qs = take((size_t)s.count * 4);
  L.off_qc = take((size_t)s.count * 4);
  L.off_qr = take((size_t)s.count * 4);
  L.off_q0 = take((size_t)s.count * 4);
  L.off_q1 = take((size_t)s.count * 4);
  L.off_flag = take(256);
  L.total = cur;
  return L;
}

size_t mc_select_ws(uint64_t n_inputs, uint64_t count) { return sel_layout(sel_shape(n_inputs, count)).total; }

// Selection pays when few outputs are wanted and the bucket table stays in L2: measured against the sort shuffle
// (count = n / 32) 2^25: 1.7 vs 3.1 ms, 2^27: 9.4 vs 11.1, 2^28: 28 vs 22, 2^29: 75 vs 44, 2^31: 291 vs 195 ms -- past
// 2^27 inputs the 2^25+ counters no longer fit the 126 MB L2 and every histogram atomic goes to DRAM.  The results are
// identical at every size (experiments/bench_select.py, shuffle_2p31.py).
bool mc_select_worthwhile(uint64_t n_inputs, uint64_t count) {
  return n_inputs >= (1ull << 16) && n_inputs <= (1ull << 27) && count >= 1 && count * 4 <= n_inputs;
}

// Keys of a round, SEL_U threefry blocks per thread and trip: a block serves the pair of positions (i, i + h) in the
// original mode, one position in the partitionable mode.  keys[] / pos[] hold 2 * SEL_U entries, pos = 0xFFFFFFFF marks
// an unused one.  Batching keeps the dependent memory operations of a trip (bitmap loads, atomics) in flight together.
constexpr int SEL_U = 4;

__device__ __forceinline__ void sel_keys_batch(u32x2 key, uint32_t n, bool part, uint32_t i0, uint32_t stride,
                                               uint32_t (&keys)[2 * SEL_U], uint32_t (&pos)[2 * SEL_U]) {
  const uint32_t h = (n + 1u) >> 1;
#pragma unroll
  for (int u = 0; u < SEL_U; ++u) {
    const uint32_t i = i0 + (uint32_t)u * stride;
    pos[2 * u] = pos[2 * u + 1] = 0xFFFFFFFFu;
    keys[2 * u] = keys[2 * u + 1] = 0u;
    if (part) {
      if (i < n) {
        const u32x2 o = threefry2x32(key.x, key.y, 0u, i);
        keys[2 * u] = o.x ^ o.y; pos[2 * u] = i;
      }
    } else if (i < h) {
      const uint32_t hi = i + h;
      const u32x2 o = threefry2x32(key.x, key.y, i, hi < n ? hi : 0u);
      keys[2 * u] = o.x; pos[2 * u] = i;
      if (hi < n) { keys[2 * u + 1] = o.y; pos[2 * u + 1] = hi; }
    }
  }
}

// Every kernel of a round takes `nbl`, the number of leading buckets the round looks at (keys of later buckets are
// skipped), and `guard`: nullptr, or a device flag that must be non-zero for the kernel to do anything (the re-run of
// a round whose bucket limit turned out too small).
__global__ void __launch_bounds__(256)
sel_zero_kernel(uint32_t* __restrict__ a, size_t na, uint32_t* __restrict__ b, size_t nb_, uint32_t* __restrict__ c, size_t nc,
                const uint32_t* __restrict__ guard) {
  if (guard && *guard == 0u) return;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < na; i += stride) a[i] = 0u;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nb_; i += stride) b[i] = 0u;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += stride) c[i] = 0u;
}

__global__ void __launch_bounds__(256)
sel_hist_kernel(const DerivedKeys* __restrict__ dk, int round, SelShape s, uint32_t nbl, int partitionable,
                uint32_t* __restrict__ hist, const uint32_t* __restrict__ guard) {
  if (guard && *guard == 0u) return;
  const u32x2 key = dk->round_sub[round];
  const bool part = partitionable != 0;
  const uint32_t work = part ? s.n : (s.n + 1u) >> 1;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < work; i0 += stride * SEL_U) {
    uint32_t keys[2 * SEL_U], pos[2 * SEL_U];
    sel_keys_batch(key, s.n, part, i0, stride, keys, pos);
#pragma unroll
    for (int e = 0; e < 2 * SEL_U; ++e) {
      const uint32_t b = keys[e] >> s.shift;
      if (pos[e] != 0xFFFFFFFFu && b < nbl) atomicAdd(hist + b, 1u);
    }
  }
}

// queries (ranks; nullptr = 0 .. count-1) -> bucket, bucket size and the rank inside it; marks the bucket.
// prefix = exclusive scan of the counts over nbl + 1 entries (prefix[nbl] = keys in the first nbl buckets).  A rank
// beyond that raises `overflow` (the round is then re-run over all buckets).
//
// The members of the queried buckets are stored COMPACTLY (about a fifth of the keys: the region then stays in L2
// instead of being 8-byte writes scattered over n * 8 bytes): the query that marks a bucket first reserves the
// bucket's slots -- one atomic per CTA and trip on `region_top` after a block-wide scan of the reservations -- and
// leaves the bucket's base in cursor[bucket]; the scatter pass advances it, the ranking pass reads its end.
__global__ void __launch_bounds__(256)
sel_locate_kernel(const uint32_t* __restrict__ queries, SelShape s, uint32_t nbl, const uint32_t* __restrict__ prefix,
                  uint32_t* __restrict__ qs, uint32_t* __restrict__ qc, uint32_t* __restrict__ qr,
                  uint32_t* __restrict__ bitmap, uint32_t* __restrict__ cursor, uint32_t* __restrict__ region_top,
                  uint32_t* __restrict__ overflow, const uint32_t* __restrict__ guard) {
  if (guard && *guard == 0u) return;
  __shared__ uint32_t warp_tot[8];
  __shared__ uint32_t block_base;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t trips = (s.count + stride - 1) / stride;          // uniform: the block scan needs every thread
  for (uint32_t t = 0; t < trips; ++t) {
    const uint32_t q = t * stride + blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t reserve = 0, lo = 0;
    if (q < s.count) {
      const uint32_t i = queries ? queries[q] : q;
      if (i >= __ldg(prefix + nbl)) {
        if (overflow) *overflow = 1u;
        qc[q] = 0u;
      } else {
        // largest b with prefix[b] <= i.  The keys are uniform: rank i sits near bucket i * nb / n (within a few
        // sigma), so the search starts in a window around that guess and widens only if the window misses.
        const uint32_t guess = min((uint32_t)(((uint64_t)i * s.nb) / s.n), nbl - 1u);
        lo = guess > s.window ? guess - s.window : 0u;
        uint32_t hi = (uint32_t)min((uint64_t)nbl, (uint64_t)guess + s.window);
        if (__ldg(prefix + lo) > i) lo = 0u;
        if (hi < nbl && __ldg(prefix + hi) <= i) hi = nbl;
        while (lo + 1 < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if (__ldg(prefix + mid) <= i) lo = mid; else hi = mid;
        }
        const uint32_t start = __ldg(prefix + lo);
        const uint32_t cnt = __ldg(prefix + lo + 1u) - start;
        qs[q] = lo;
        qc[q] = cnt;
        qr[q] = i - start;
        const uint32_t bit = 1u << (lo & 31);
        if ((atomicOr(bitmap + (lo >> 5), bit) & bit) == 0u) reserve = cnt;   // first to mark it: reserves its slots
      }
    }
    uint32_t x = reserve;                                          // inclusive scan over the CTA
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t run = 0;
      for (int w = 0; w < 8; ++w) { const uint32_t v = warp_tot[w]; warp_tot[w] = run; run += v; }
      block_base = run ? atomicAdd(region_top, run) : 0u;
    }
    __syncthreads();
    if (reserve) cursor[lo] = block_base + warp_tot[wid] + x - reserve;
    __syncthreads();                                               // warp_tot / block_base are rewritten next trip
  }
}

__global__ void __launch_bounds__(256)
sel_scatter_kernel(const DerivedKeys* __restrict__ dk, int round, SelShape s, uint32_t nbl, int partitionable,
                   const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ prefix, uint32_t* __restrict__ cursor,
                   uint2* __restrict__ members, const uint32_t* __restrict__ guard) {
  if (guard && *guard == 0u) return;
  const u32x2 key = dk->round_sub[round];
  const bool part = partitionable != 0;
  const uint32_t work = part ? s.n : (s.n + 1u) >> 1;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < work; i0 += stride * SEL_U) {
    uint32_t keys[2 * SEL_U], pos[2 * SEL_U], slot[2 * SEL_U];
    sel_keys_batch(key, s.n, part, i0, stride, keys, pos);
    bool want[2 * SEL_U];
#pragma unroll
    for (int e = 0; e < 2 * SEL_U; ++e) {             // all bitmap loads first
      const uint32_t b = keys[e] >> s.shift;
      want[e] = pos[e] != 0xFFFFFFFFu && b < nbl && ((__ldg(bitmap + (b >> 5)) >> (b & 31)) & 1u);
    }
#pragma unroll
    for (int e = 0; e < 2 * SEL_U; ++e) {             // then every atomic before the first store
      const uint32_t b = keys[e] >> s.shift;
      slot[e] = want[e] ? atomicAdd(cursor + b, 1u) : 0u;         // the bucket's base was left there by sel_locate_kernel
    }
#pragma unroll
    for (int e = 0; e < 2 * SEL_U; ++e)
      if (want[e]) members[slot[e]] = make_uint2(keys[e], pos[e]);
  }
}

// The member of a query's bucket with exactly qr[q] smaller members in (key, position) order.  A warp takes 32
// consecutive queries (their records loaded coalesced) and resolves two at a time, one per half-warp: a lane holds one
// member, the members travel round the half-warp by shuffles; the next pair's members are loaded before the current
// pair is ranked.  Buckets of more than 16 members (rare: the mean is 8 .. 16) are ranked from memory.
__global__ void __launch_bounds__(256)
sel_answer_kernel(SelShape s, const uint32_t* __restrict__ qs, const uint32_t* __restrict__ qc,
                  const uint32_t* __restrict__ qr, const uint32_t* __restrict__ cursor, const uint2* __restrict__ members,
                  uint32_t* __restrict__ out,
                  const uint32_t* __restrict__ guard) {
  if (guard && *guard == 0u) return;
  const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  auto fetch = [&](uint32_t start, uint32_t cnt) -> uint64_t {
    if ((uint32_t)hl < cnt && cnt <= 16u) {
      const uint2 m = __ldcs(members + start + hl);
      return ((uint64_t)m.x << 32) | m.y;
    }
    return ~0ull;
  };
  for (uint32_t base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u; base < s.count; base += warps * 32u) {
    const uint32_t q = base + lane;
    uint32_t st = 0, c = 0, r = 0;
    if (q < s.count) {
      c = __ldg(qc + q); r = __ldg(qr + q);
      if (c) st = cursor[__ldg(qs + q)] - c;                 // the scatter pass left the END of the bucket's members there
    }
    uint64_t nxt = fetch(__shfl_sync(0xffffffffu, st, half), __shfl_sync(0xffffffffu, c, half));
    for (int it = 0; it < 16; ++it) {
      const int src = 2 * it + half;
      const uint32_t start = __shfl_sync(0xffffffffu, st, src), cnt = __shfl_sync(0xffffffffu, c, src);
      const uint32_t rr = __shfl_sync(0xffffffffu, r, src);
      const uint32_t c_even = __shfl_sync(0xffffffffu, c, 2 * it), c_odd = __shfl_sync(0xffffffffu, c, 2 * it + 1);
      const uint64_t mine = nxt;
      {
        const int nsrc = (2 * it + 2 + half) & 31;
        const uint32_t ns = __shfl_sync(0xffffffffu, st, nsrc), nc = __shfl_sync(0xffffffffu, c, nsrc);
        nxt = it < 15 ? fetch(ns, nc) : ~0ull;
      }
      const uint32_t trips = min(16u, max(c_even <= 16u ? c_even : 0u, c_odd <= 16u ? c_odd : 0u));
      uint32_t smaller = 0;
      for (uint32_t j = 0; j < trips; ++j) {
        const uint64_t other = __shfl_sync(0xffffffffu, mine, (int)j, 16);
        smaller += other < mine ? 1u : 0u;
      }
      if (cnt <= 16u) {
        if ((uint32_t)hl < cnt && smaller == rr) out[base + src] = (uint32_t)mine;
      } else {                                        // a crowded bucket: every lane ranks its members against all
        for (uint32_t m0 = hl; m0 < cnt; m0 += 16) {
          const uint2 mm = members[start + m0];
          const uint64_t me = ((uint64_t)mm.x << 32) | mm.y;
          uint32_t less = 0;
          for (uint32_t j = 0; j < cnt; ++j) {
            const uint2 o = __ldg(members + start + j);
            less += ((((uint64_t)o.x << 32) | o.y) < me) ? 1u : 0u;
          }
          if (less == rr) out[base + src] = mm.y;
        }
      }
    }
  }
}

// Side stream and events of the calling thread's current device (created once, kept for the life of the process).
struct SelSide { cudaStream_t stream = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };

static SelSide* sel_side() {
  static thread_local SelSide sides[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SelSide& sd = sides[dev];
  if (!sd.stream) {
    if (cudaStreamCreateWithFlags(&sd.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&sd.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&sd.join, cudaEventDisableTiming) != cudaSuccess) {
      sd.stream = nullptr;
      return nullptr;
    }
  }
  return &sd;
}

// idx_out[q] = shuffle(key, arange(n))[q] for q < count.  `dk` holds the round keys (derive_keys_kernel).
//
// The round processed first asks for the ranks 0 .. count-1, i.e. for the `count` smallest keys: only the leading
// buckets can hold them, so that round looks at `nbl` buckets chosen to hold count + 8 sigma keys, and is re-run over
// all buckets, on the device (guarded launches, no host round trip), in the 1e-15 case the limit was too small.
//
// The bucket counts of a round do not depend on the queries: while the stream resolves round r (locate, scatter,
// rank: latency-bound), a side stream builds the scanned counts of round r-1 (L2-atomic-bound) in the other buffer.
int mc_select_impl(const DerivedKeys* dk, int rounds, int partitionable, uint64_t n_inputs, uint64_t count,
                   uint32_t* idx_out, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const SelShape s = sel_shape(n_inputs, count);
  const SelLayout L = sel_layout(s);
  if (workspace_bytes < L.total) return fail(MCCB200_ERR_WORKSPACE_TOO_SMALL, "mc_select: workspace %zu < %zu", workspace_bytes, L.total);
  char* w = (char*)workspace;
  uint32_t* hists[2] = {(uint32_t*)(w + L.off_hist[0]), (uint32_t*)(w + L.off_hist[1])};
  void* scans[2] = {w + L.off_scan[0], w + L.off_scan[1]};
  uint32_t* cursor = (uint32_t*)(w + L.off_cursor);
  uint32_t* bitmap = (uint32_t*)(w + L.off_bitmap);
  uint2* members = (uint2*)(w + L.off_members);
  uint32_t* qs = (uint32_t*)(w + L.off_qs);
  uint32_t* qc = (uint32_t*)(w + L.off_qc);
  uint32_t* qr = (uint32_t*)(w + L.off_qr);
  uint32_t* qbuf[2] = {(uint32_t*)(w + L.off_q0), (uint32_t*)(w + L.off_q1)};
  uint32_t* flag = (uint32_t*)(w + L.off_flag);
  uint32_t* region_top = flag + 8;                 // slots of the compact member region handed out so far
  const int sms = sm_count();
  const uint32_t work = partitionable ? s.n : (s.n + 1u) / 2u;
  const int grid_n = (int)min(((uint64_t)work + 256 * SEL_U - 1) / (256 * SEL_U), (uint64_t)sms * 8);
  const int grid_q = (int)min(((uint64_t)s.count + 255) / 256, (uint64_t)sms * 8);
  const int grid_w = (int)min(((uint64_t)s.count + 255) / 256, (uint64_t)sms * 8);
  // buckets that hold the `count` smallest keys, with margin: keys per bucket ~ Binomial(n, 1 / nb)
  const double per_bucket = (double)s.n / (double)s.nb;
  const double want = (double)s.count + 8.0 * sqrt((double)s.count) + 64.0;
  const uint64_t nbl_first64 = (uint64_t)ceil(want / per_bucket) + 1;
  uint32_t nbl_first = nbl_first64 < s.nb ? (uint32_t)nbl_first64 : s.nb;
  // test hook: a limit that is certainly too small, to exercise the guarded re-run
  if (env_int("MCCB200_SELECT_SHORT_LIMIT", 0) != 0) nbl_first = (uint32_t)((double)s.count / per_bucket / 2.0) + 1u;
  SelSide* side = env_int("MCCB200_SELECT_OVERLAP", 1) != 0 ? sel_side() : nullptr;

  // scanned bucket counts of round r over the first nbl buckets, into hists[r & 1]
  auto counts_pass = [&](int r, uint32_t nbl, const uint32_t* guard, cudaStream_t cs) -> int {
    uint32_t* hist = hists[r & 1];
    const int grid_z = (int)min(((uint64_t)nbl + 256) / 256, (uint64_t)sms * 4);
    sel_zero_kernel<<<grid_z, 256, 0, cs>>>(hist, (size_t)nbl + 1, nullptr, 0, nullptr, 0, guard);
    sel_hist_kernel<<<grid_n, 256, 0, cs>>>(dk, r, s, nbl, partitionable, hist, guard);
    count_launches(2);
    return scan_u32_exclusive(hist, (uint64_t)nbl + 1, scans[r & 1], cs, guard);
  };
  // queries -> answers of round r, given its scanned counts
  auto resolve_pass = [&](int r, const uint32_t* queries, uint32_t nbl, uint32_t* out, uint32_t* overflow,
                          const uint32_t* guard) {
    uint32_t* hist = hists[r & 1];
    const int grid_z = (int)min(((uint64_t)nbl + 256) / 256, (uint64_t)sms * 4);
    sel_zero_kernel<<<grid_z, 256, 0, st>>>(region_top, 1, bitmap, ((size_t)nbl + 31) / 32, nullptr, 0, guard);
    sel_locate_kernel<<<grid_q, 256, 0, st>>>(queries, s, nbl, hist, qs, qc, qr, bitmap, cursor, region_top, overflow, guard);
    sel_scatter_kernel<<<grid_n, 256, 0, st>>>(dk, r, s, nbl, partitionable, bitmap, hist, cursor, members, guard);
    sel_answer_kernel<<<grid_w, 256, 0, st>>>(s, qs, qc, qr, cursor, members, out, guard);
    count_launches(4);
  };

  const uint32_t* queries = nullptr;               // round processed first: ranks 0 .. count-1
  const int first = rounds - 1;
  const bool limited = nbl_first < s.nb;
  int rc = counts_pass(first, limited ? nbl_first : s.nb, nullptr, st);
  if (rc) return rc;
  for (int r = first; r >= 0; --r) {
    uint32_t* out = r == 0 ? idx_out : qbuf[r & 1];
    bool forked = false;
    if (r > 0 && side) {                             // hists[(r - 1) & 1] is free: round r + 1 has been enqueued before
      if (cudaEventRecord(side->fork, st) != cudaSuccess || cudaStreamWaitEvent(side->stream, side->fork, 0) != cudaSuccess)
        return fail(MCCB200_ERR_CUDA, "mc_select: fork failed");
      rc = counts_pass(r - 1, s.nb, nullptr, side->stream);
      if (rc) return rc;
      if (cudaEventRecord(side->join, side->stream) != cudaSuccess) return fail(MCCB200_ERR_CUDA, "mc_select: join failed");
      forked = true;
    }
    if (r == first && limited) {
      if (cudaMemsetAsync(flag, 0, 4, st) != cudaSuccess) return fail(MCCB200_ERR_CUDA, "mc_select: memset failed");
      resolve_pass(r, queries, nbl_first, out, flag, nullptr);
      rc = counts_pass(r, s.nb, flag, st);           // does nothing unless the limit was too small
      if (rc) return rc;
      resolve_pass(r, queries, s.nb, out, nullptr, flag);
    } else {
      resolve_pass(r, queries, s.nb, out, nullptr, nullptr);
    }
    if (forked) {
      if (cudaStreamWaitEvent(st, side->join, 0) != cudaSuccess) return fail(MCCB200_ERR_CUDA, "mc_select: join failed");
    } else if (r > 0) {
      rc = counts_pass(r - 1, s.nb, nullptr, st);
      if (rc) return rc;
    }
    queries = out;
  }
  return check_launch("mc_indices(select)");
}

}  // namespace mccb
