// Resampling indices of MonteCarloKernel, bit-exact with jax.random (sm_100a).
//
// Replaces, for /root/reference/mccube/_kernels/random.py:56-65,
//   force_bitcast_convert_type -> jr.fold_in -> split_by_tree -> jr.choice
// i.e. (jax/_src/random.py) choice -> permutation -> _shuffle -> random_bits + sort_key_val
// when replace=False, and choice -> randint when replace=True.  Both values of
// jax_threefry_partitionable are implemented.  The oracle restatement is
// oracle/jax_random.py; SURVEY.md Appendix A states the algorithm.
#include "common.cuh"
#include "rng.cuh"

namespace mccb {

int sort_pairs_impl(const uint32_t*, const uint32_t*, uint64_t, int, uint32_t*, uint32_t*, void*, size_t,
                    cudaStream_t);
size_t sort_pairs_ws(uint64_t);
bool sort_pairs_top_worthwhile(uint64_t count, uint64_t top);
size_t sort_pairs_top_ws(uint64_t count, uint64_t top);
size_t mc_select_ws(uint64_t n_inputs, uint64_t count);
bool mc_select_worthwhile(uint64_t n_inputs, uint64_t count);
int mc_select_impl(const DerivedKeys* dk, int rounds, int partitionable, uint64_t n_inputs, uint64_t count,
                   uint32_t* idx_out, void* workspace, size_t workspace_bytes, cudaStream_t st);
int sort_pairs_top_impl(const uint32_t*, const uint32_t*, uint64_t, uint64_t, uint32_t*, void*, size_t, cudaStream_t);

__global__ void derive_keys_kernel(mccb200_rng rng, int rounds, DerivedKeys* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  u32x2 key = {rng.key[0], rng.key[1]};
  if (rng.key_dev) key = {rng.key_dev[0], rng.key_dev[1]};
  const bool part = rng.partitionable != 0;
  if (rng.fold_time) {
    float t = rng.t_dev ? *rng.t_dev : rng.t;
    key = fold_in_key(key, __float_as_uint(t));
    key = split_key(key, rng.n_leaves, rng.leaf, part);
  }
  out->step_key = key;
  out->randint_hi = split_key(key, 2, 0, part);
  out->randint_lo = split_key(key, 2, 1, part);
  u32x2 k = key;
  for (int r = 0; r < rounds && r < MAX_ROUNDS; ++r) {
    u32x2 nk = split_key(k, 2, 0, part);
    out->round_sub[r] = split_key(k, 2, 1, part);
    k = nk;
  }
}

void launch_derive_keys(const mccb200_rng& rng, int rounds, DerivedKeys* out, cudaStream_t st) {
  derive_keys_kernel<<<1, 32, 0, st>>>(rng, rounds, out);
}

// random_bits(sub, 32, (n,)).  Original mode: one threefry block yields elements i and
// i + ceil(n/2), so each thread produces a pair; partitionable mode: one block per element.
__global__ void __launch_bounds__(256)
random_bits_kernel(const DerivedKeys* __restrict__ dk, int round, uint32_t n, int partitionable,
                   uint32_t* __restrict__ out) {
  const u32x2 key = dk->round_sub[round];
  const uint32_t h = (n + 1u) >> 1;
  const uint32_t stride = gridDim.x * blockDim.x;
  if (partitionable) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      u32x2 o = threefry2x32(key.x, key.y, 0u, i);
      out[i] = o.x ^ o.y;
    }
  } else {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < h; i += stride) {
      uint32_t hi = i + h;
      u32x2 o = threefry2x32(key.x, key.y, i, hi < n ? hi : 0u);
      out[i] = o.x;
      if (hi < n) out[hi] = o.y;
    }
  }
}

// jr.randint(key, (count,), 0, span) for int32 (jax/_src/random.py randint).
// Elements [first, first + slice) of the `count` draws (the whole vector when first = 0, slice = count).
__global__ void __launch_bounds__(256)
randint_kernel(const DerivedKeys* __restrict__ dk, uint32_t count, uint32_t span, int partitionable,
               uint32_t* __restrict__ out, uint32_t first, uint32_t slice) {
  const u32x2 k1 = dk->randint_hi, k2 = dk->randint_lo;
  uint32_t mult = 65536u % span;
  mult = (mult * mult) % span;
  const bool part = partitionable != 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  // mult == 0 (e.g. every power-of-two span >= 2^16: 65536^2 wraps to 0) makes the high draw drop out of
  // ((hi % span) * mult + lo % span) % span: its threefry block and two of the three divisions are skipped, and
  // for a power-of-two span the remaining one is a mask.  Same values, a third of the arithmetic.
  const bool pow2 = (span & (span - 1u)) == 0u;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < slice; q += stride) {
    const uint32_t i = first + q;
    const uint32_t lo = random_bits32_at(k2, i, count, part);
    if (mult == 0u) {
      out[q] = pow2 ? (lo & (span - 1u)) : (lo % span);
    } else {
      const uint32_t hi = random_bits32_at(k1, i, count, part);
      const uint32_t off = (hi % span) * mult + (lo % span);
      out[q] = off % span;
    }
  }
}

// random_bits(key, 64, (n,))[i] (jax/_src/prng.py): original mode hashes iota(2n) and pairs word i with
// word n + i; partitionable mode is one block per element.
__device__ __forceinline__ uint64_t random_bits64_at(u32x2 key, uint32_t i, uint32_t n, bool partitionable) {
  if (partitionable) {
    const u32x2 o = threefry2x32(key.x, key.y, 0u, i);
    return ((uint64_t)o.x << 32) | o.y;
  }
  const uint32_t hi = random_bits32_at(key, i, 2u * n, false);
  const uint32_t lo = random_bits32_at(key, n + i, 2u * n, false);
  return ((uint64_t)hi << 32) | lo;
}

// jr.randint with jax_enable_x64 (int64): elements [first, first + slice) of the `count` draws over [0, span).
__global__ void __launch_bounds__(256)
randint64_kernel(const DerivedKeys* __restrict__ dk, uint32_t count, uint64_t span, int partitionable,
                 uint64_t* __restrict__ out, uint32_t first, uint32_t slice) {
  const u32x2 k1 = dk->randint_hi, k2 = dk->randint_lo;
  uint64_t mult = (1ull << 32) % span;
  mult = (mult * mult) % span;                            // the product wraps in 64 bits, as lax.mul on uint64 does
  const bool part = partitionable != 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  const bool pow2 = (span & (span - 1ull)) == 0ull;       // (see randint_kernel: mult == 0 drops the high draw)
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < slice; q += stride) {
    const uint32_t i = first + q;
    const uint64_t lo = random_bits64_at(k2, i, count, part);
    if (mult == 0ull) {
      out[q] = pow2 ? (lo & (span - 1ull)) : (lo % span);
    } else {
      const uint64_t hi = random_bits64_at(k1, i, count, part);
      const uint64_t off = (hi % span) * mult + (lo % span);
      out[q] = off % span;
    }
  }
}

__global__ void __launch_bounds__(256) iota_kernel(uint32_t count, uint32_t* __restrict__ out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) out[i] = i;
}

// _shuffle for small inputs in ONE launch.  A single CTA derives the keys, then per round draws the 32-bit
// sort keys (kept in registers, up to 16 per thread) and sorts by them entirely in shared memory.  The keys
// are uniform random words, so the sort is a counting pass over their top 14 bits (16384 buckets, at most
// ~1 element each on average: shared-memory atomics hand out a slot inside the bucket) after which every
// element finds its final place as "start of its bucket + members that order before it on (key, position)".
// Positions are unique, so the result does not depend on the order the atomics were served in and equals
// the stable sort_key_val's.  A bucket member is one word: the key's low 18 bits above the 14-bit position.
// The kernel is bound by one SM's issue rate (threefry) and by bank conflicts of the random shared-memory
// accesses (~3.5 cycles per warp access): every phase is written for the fewest such accesses.
constexpr int SMALL_SHUFFLE_MAX = 16384;
constexpr int SS_THREADS = 1024, SS_SLOTS = SMALL_SHUFFLE_MAX / SS_THREADS;
constexpr int SS_BUCKET_BITS = 14, SS_BUCKETS = 1 << SS_BUCKET_BITS, SS_BUCKET_SHIFT = 32 - SS_BUCKET_BITS;
static_assert(SMALL_SHUFFLE_MAX <= (1 << 14) && SS_BUCKET_SHIFT + 14 == 32, "a bucket member packs into 32 bits");

static size_t small_shuffle_smem(uint32_t n) {
  const size_t na = ((size_t)n + 7) & ~(size_t)7;
  return na * 8 + (size_t)SS_BUCKETS * 4;
}

#ifdef MCCB200_SS_TRACE
__device__ long long g_ss_trace[16];
#define SS_TRACE(k) do { if (threadIdx.x == 0 && r == 0) g_ss_trace[k] = clock64(); } while (0)
#else
#define SS_TRACE(k) do { } while (0)
#endif
__global__ void __launch_bounds__(SS_THREADS)
small_shuffle_kernel(mccb200_rng rng, int rounds, uint32_t n, uint32_t count, uint32_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const uint32_t na = (n + 7u) & ~7u;
  uint32_t* cnt = reinterpret_cast<uint32_t*>(s_raw);           // [SS_BUCKETS] counts, then start | count << 16
  uint32_t* el = cnt + SS_BUCKETS;                              // [na] bucket members: key low bits << 14 | position
  uint16_t* xa = reinterpret_cast<uint16_t*>(el + na);          // [na] the sequence being shuffled
  uint16_t* xb = xa + na;                                       // [na] ... and its next version
  __shared__ u32x2 s_sub[MAX_ROUNDS];
  __shared__ uint32_t s_warp_tot[SS_THREADS / 32];
  const bool part = rng.partitionable != 0;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    u32x2 key = {rng.key[0], rng.key[1]};
    if (rng.key_dev) key = {rng.key_dev[0], rng.key_dev[1]};
    if (rng.fold_time) {
      float t = rng.t_dev ? *rng.t_dev : rng.t;
      key = fold_in_key(key, __float_as_uint(t));
      key = split_key(key, rng.n_leaves, rng.leaf, part);
    }
    for (int r = 0; r < rounds; ++r) {
      u32x2 nk = split_key(key, 2, 0, part);
      s_sub[r] = split_key(key, 2, 1, part);
      key = nk;
    }
  }
  for (uint32_t i = threadIdx.x; i < SS_BUCKETS / 4; i += SS_THREADS) reinterpret_cast<uint4*>(cnt)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const uint32_t h = (n + 1u) >> 1;                             // original mode: block b yields elements b and b + h
  for (int r = 0; r < rounds; ++r) {
    const u32x2 sub = s_sub[r];
    SS_TRACE(0);
    uint32_t key[SS_SLOTS], slot[SS_SLOTS];                     // slot s of this thread: element pos_of(s)
    auto pos_of = [&](int s) -> uint32_t {
      return part ? threadIdx.x + (uint32_t)s * SS_THREADS
                  : threadIdx.x + (uint32_t)(s >> 1) * SS_THREADS + ((s & 1) ? h : 0u);
    };
    auto live_at = [&](int s) -> bool {
      const uint32_t b = threadIdx.x + (uint32_t)(s >> 1) * SS_THREADS;
      return part ? pos_of(s) < n : (b < h && pos_of(s) < n);
    };
    if (part) {
#pragma unroll
      for (int s = 0; s < SS_SLOTS; ++s) {
        if ((uint32_t)s * SS_THREADS >= n) { key[s] = 0; continue; }       // CTA-uniform
        const u32x2 o = threefry2x32(sub.x, sub.y, 0u, pos_of(s));
        key[s] = o.x ^ o.y;
      }
    } else {
#pragma unroll
      for (int s = 0; s < SS_SLOTS; s += 2) {
        if ((uint32_t)(s >> 1) * SS_THREADS >= h) { key[s] = key[s + 1] = 0; continue; }   // CTA-uniform
        const uint32_t b = threadIdx.x + (uint32_t)(s >> 1) * SS_THREADS;
        const u32x2 o = threefry2x32(sub.x, sub.y, b, b + h < n ? b + h : 0u);   // the pad element counts as 0
        key[s] = o.x;
        key[s + 1] = o.y;
      }
    }
    SS_TRACE(1);
#pragma unroll
    for (int s = 0; s < SS_SLOTS; ++s)
      if (live_at(s)) slot[s] = atomicAdd(&cnt[key[s] >> SS_BUCKET_SHIFT], 1u);
    __syncthreads();
    SS_TRACE(2);
    {  // exclusive scan of the bucket counts; leaves start | count << 16.  Warp w owns 512 consecutive buckets,
       // read as 4 conflict-free rows of 32 x uint4
      static_assert(SS_BUCKETS == 16 * SS_THREADS, "scan layout");
      uint4* mine = reinterpret_cast<uint4*>(cnt) + warp * 128 + lane;
      uint32_t excl[4], carry = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 v = mine[q * 32];
        const uint32_t sum = v.x + v.y + v.z + v.w;
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, inc, o);
          if (lane >= (uint32_t)o) inc += up;
        }
        excl[q] = carry + inc - sum;
        carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
      }
      if (lane == 0) s_warp_tot[warp] = carry;
      __syncthreads();
      const uint32_t wbase = __reduce_add_sync(0xFFFFFFFFu, lane < warp ? s_warp_tot[lane] : 0u);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 v = mine[q * 32];          // (re-read: conflict-free, cheaper than 16 live registers)
        uint32_t base = wbase + excl[q];
        uint4 w;
        w.x = base | (v.x << 16); base += v.x;
        w.y = base | (v.y << 16); base += v.y;
        w.z = base | (v.z << 16); base += v.z;
        w.w = base | (v.w << 16);
        mine[q * 32] = w;
      }
    }
    __syncthreads();
    SS_TRACE(3);
#pragma unroll
    for (int s = 0; s < SS_SLOTS; ++s)
      if (live_at(s)) {
        const uint32_t e = cnt[key[s] >> SS_BUCKET_SHIFT];           // start | count << 16
        el[(e & 0xFFFFu) + slot[s]] = (key[s] << SS_BUCKET_BITS) | pos_of(s);
        slot[s] = e;
      }
    __syncthreads();
    SS_TRACE(4);
    // final place = start of the bucket + members ordering before this one; the carried sequence moves there
    // directly (round 0 starts from the identity)
#pragma unroll
    for (int s = 0; s < SS_SLOTS; ++s)
      if (live_at(s)) {
        const uint32_t pos = pos_of(s), members = slot[s] >> 16, me = (key[s] << SS_BUCKET_BITS) | pos;
        uint32_t lo = slot[s] & 0xFFFFu;
        if (members > 1) {     // (itself never counts; ~0 is no member's word below any `me`)
          const uint32_t e0 = el[lo], e1 = el[lo + 1], e2 = members > 2 ? el[lo + 2] : ~0u, e3 = members > 3 ? el[lo + 3] : ~0u;
          uint32_t before = (e0 < me) + (e1 < me) + (e2 < me) + (e3 < me);
          for (uint32_t j = 4; j < members; ++j) before += el[lo + j] < me ? 1u : 0u;
          lo += before;
        }
        xb[lo] = r == 0 ? (uint16_t)pos : xa[pos];
      }
    __syncthreads();
    SS_TRACE(5);
    for (uint32_t i = threadIdx.x; i < SS_BUCKETS / 4; i += SS_THREADS) reinterpret_cast<uint4*>(cnt)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    SS_TRACE(6);
    { uint16_t* t = xa; xa = xb; xb = t; }
  }
  for (uint32_t q = threadIdx.x; q < count; q += SS_THREADS) out[q] = xa[q];
}

static int shuffle_rounds(uint64_t size) {
  if (size < 1) size = 1;
  return (int)ceil(3.0 * log((double)size) / log(4294967295.0));
}

struct IdxPlan {
  size_t off_dk, off_keys, off_va, off_vb, off_sort, total;
  int rounds;
};

static IdxPlan make_idx_plan(uint64_t n_inputs, uint64_t count, int replace) {
  IdxPlan p{};
  p.rounds = replace ? 0 : shuffle_rounds(n_inputs);
  p.off_dk = 0;
  size_t cur = align_up(sizeof(DerivedKeys), 256);
  if (!replace && env_int("MCCB200_SHUFFLE_SELECT", 1) != 0 && mc_select_worthwhile(n_inputs, count)) {
    // the selection path lays its own scratch after dk (19 GiB at 2^31 inputs; the sort path would take 66 GiB)
    cur += mc_select_ws(n_inputs, count);
  } else if (!replace) {
    size_t arr = align_up((size_t)n_inputs * 4, 256);
    p.off_keys = cur; cur += arr;
    p.off_va = cur; cur += arr;
    p.off_vb = cur; cur += arr;
    p.off_sort = cur;
    size_t sw = sort_pairs_ws(n_inputs);
    if (count >= 1 && sort_pairs_top_worthwhile(n_inputs, count)) sw = max(sw, sort_pairs_top_ws(n_inputs, count));
    cur += sw;
  }
  p.total = cur;
  return p;
}

}  // namespace mccb

using namespace mccb;

#ifdef MCCB200_SS_TRACE
// Debug hook (not part of the C ABI): clock stamps of the first round of the last small_shuffle_kernel launch.
extern "C" __attribute__((visibility("default"))) void mccb200_debug_ss_trace(long long* host) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host, g_ss_trace, sizeof(long long) * 16);
}
#endif

extern "C" size_t mccb200_mc_indices_workspace_bytes(uint64_t n_inputs, uint64_t count, int replace) {
  return make_idx_plan(n_inputs, count, replace).total;
}

extern "C" int mccb200_mc_indices(const mccb200_rng* rng, uint64_t n_inputs, uint64_t count, int replace,
                                  uint32_t* idx_out, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MCCB_REQUIRE(rng && idx_out, "mc_indices: NULL argument");
  MCCB_REQUIRE(n_inputs >= 1 && n_inputs < 0xFFFFFFFFull, "mc_indices: n_inputs %llu out of range",
               (unsigned long long)n_inputs);
  MCCB_REQUIRE(rng->n_leaves >= 1 && rng->leaf < rng->n_leaves, "mc_indices: bad leaf %u / %u", rng->leaf,
               rng->n_leaves);
  MCCB_REQUIRE(replace || count <= n_inputs, "mc_indices: cannot draw %llu > %llu without replacement",
               (unsigned long long)count, (unsigned long long)n_inputs);
  IdxPlan p = make_idx_plan(n_inputs, count, replace);
  MCCB_REQUIRE(p.rounds <= MAX_ROUNDS, "mc_indices: too many rounds");
  if (workspace_bytes < p.total || workspace == nullptr)
    return fail(MCCB200_ERR_WORKSPACE_TOO_SMALL, "mc_indices: workspace %zu < %zu", workspace_bytes, p.total);
  if (count == 0) return MCCB200_OK;
  if (!replace && p.rounds > 0 && n_inputs <= (uint64_t)SMALL_SHUFFLE_MAX) {
    const size_t smem = small_shuffle_smem((uint32_t)n_inputs);
    // (a per-DEVICE attribute: set on every call, a process-wide flag would leave other devices at 48 KB)
    cudaFuncSetAttribute(small_shuffle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)small_shuffle_smem(SMALL_SHUFFLE_MAX));
    phase_mark("mc_indices", st);
    small_shuffle_kernel<<<1, SS_THREADS, smem, st>>>(*rng, p.rounds, (uint32_t)n_inputs, (uint32_t)count, idx_out);
    phase_mark("end", st);
    count_launches(1);
    return check_launch("mc_indices(small shuffle)");
  }
  char* w = (char*)workspace;
  DerivedKeys* dk = (DerivedKeys*)(w + p.off_dk);
  phase_mark("mc_indices", st);
  derive_keys_kernel<<<1, 32, 0, st>>>(*rng, p.rounds, dk);
  const int sms = sm_count();
  if (replace) {
    int grid = (int)min((uint64_t)(count + 255) / 256, (uint64_t)sms * 8);
    randint_kernel<<<grid, 256, 0, st>>>(dk, (uint32_t)count, (uint32_t)n_inputs, rng->partitionable, idx_out, 0u,
                                         (uint32_t)count);
    phase_mark("end", st);
    count_launches(2);
    return check_launch("mc_indices(randint)");
  }
  if (p.rounds == 0) {  // n_inputs == 1: _shuffle runs zero rounds, the permutation is the identity
    iota_kernel<<<1, 256, 0, st>>>((uint32_t)count, idx_out);
    count_launches(2);
    return check_launch("mc_indices(identity)");
  }
  const bool use_select = env_int("MCCB200_SHUFFLE_SELECT", 1) != 0;   // (0: the radix-sort shuffle, kept for tests)
  if (use_select && mc_select_worthwhile(n_inputs, count)) {
    const size_t off = align_up(sizeof(DerivedKeys), 256);
    int rc = mc_select_impl(dk, p.rounds, rng->partitionable, n_inputs, count, idx_out, w + off, p.total - off, st);
    phase_mark("end", st);
    count_launches(1);
    return rc;
  }
  uint32_t* keys = (uint32_t*)(w + p.off_keys);
  uint32_t* va = (uint32_t*)(w + p.off_va);
  uint32_t* vb = (uint32_t*)(w + p.off_vb);
  void* sort_ws = w + p.off_sort;
  size_t sort_ws_bytes = p.total - p.off_sort;
  const uint32_t n = (uint32_t)n_inputs;
  const uint32_t work = rng->partitionable ? n : (n + 1) / 2;
  int grid = (int)min(((uint64_t)work + 255) / 256, (uint64_t)sms * 8);
  const uint32_t* src = nullptr;  // implicit arange(n)
  uint32_t* dst = va;
  for (int r = 0; r < p.rounds; ++r) {
    random_bits_kernel<<<grid, 256, 0, st>>>(dk, r, n, rng->partitionable, keys);
    // only permutation[:count] is used: the last round needs just the `count` smallest keys in order
    const bool top_only = r == p.rounds - 1 && sort_pairs_top_worthwhile(n, count);
    int rc = top_only ? sort_pairs_top_impl(keys, src, n, count, dst, sort_ws, sort_ws_bytes, st)
                      : sort_pairs_impl(keys, src, n, 32, nullptr, dst, sort_ws, sort_ws_bytes, st);
    if (rc) return rc;
    src = dst;
    dst = (dst == va) ? vb : va;
  }
  if (cudaMemcpyAsync(idx_out, src, (size_t)count * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
    return fail(MCCB200_ERR_CUDA, "mc_indices: copy failed");
  phase_mark("end", st);
  count_launches(2 + p.rounds);
  return check_launch("mc_indices(shuffle)");
}

/* idx_out[q] = jr.choice(key_at(t), n_inputs, (count,), replace=True) [first + q], q in [0, slice): the
 * draws are counter based, so a rank can produce exactly the output slots it stores. */
extern "C" int mccb200_mc_indices_slice(const mccb200_rng* rng, uint64_t n_inputs, uint64_t count, uint64_t first,
                                        uint64_t slice, uint32_t* idx_out, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MCCB_REQUIRE(rng && (idx_out || slice == 0), "mc_indices_slice: NULL argument");
  MCCB_REQUIRE(n_inputs >= 1 && n_inputs < 0xFFFFFFFFull && count < 0xFFFFFFFFull && first + slice <= count,
               "mc_indices_slice: sizes out of range");
  MCCB_REQUIRE(rng->n_leaves >= 1 && rng->leaf < rng->n_leaves, "mc_indices_slice: bad leaf");
  IdxPlan p = make_idx_plan(n_inputs, count, 1);
  if (workspace_bytes < p.total || workspace == nullptr)
    return fail(MCCB200_ERR_WORKSPACE_TOO_SMALL, "mc_indices_slice: workspace %zu < %zu", workspace_bytes, p.total);
  if (slice == 0) return MCCB200_OK;
  DerivedKeys* dk = (DerivedKeys*)((char*)workspace + p.off_dk);
  phase_mark("mc_indices", st);
  derive_keys_kernel<<<1, 32, 0, st>>>(*rng, 0, dk);
  int grid = (int)min((uint64_t)(slice + 255) / 256, (uint64_t)sm_count() * (side_ctas() > 0 ? side_ctas() : 8));
  randint_kernel<<<grid, 256, 0, st>>>(dk, (uint32_t)count, (uint32_t)n_inputs, rng->partitionable, idx_out,
                                       (uint32_t)first, (uint32_t)slice);
  phase_mark("end", st);
  count_launches(2);
  return check_launch("mc_indices_slice");
}

/* The same with jax_enable_x64 semantics (int64 randint: 64-bit draws): n_inputs may exceed 2^32 (config 5:
 * n * k = 2^34 children).  count < 2^31. */
extern "C" int mccb200_mc_indices_slice64(const mccb200_rng* rng, uint64_t n_inputs, uint64_t count, uint64_t first,
                                          uint64_t slice, uint64_t* idx_out, void* workspace, size_t workspace_bytes,
                                          void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MCCB_REQUIRE(rng && (idx_out || slice == 0), "mc_indices_slice64: NULL argument");
  MCCB_REQUIRE(n_inputs >= 1 && count < 0x80000000ull && first + slice <= count, "mc_indices_slice64: sizes out of range");
  MCCB_REQUIRE(rng->n_leaves >= 1 && rng->leaf < rng->n_leaves, "mc_indices_slice64: bad leaf");
  const size_t need = align_up(sizeof(DerivedKeys), 256);
  if (workspace_bytes < need || workspace == nullptr)
    return fail(MCCB200_ERR_WORKSPACE_TOO_SMALL, "mc_indices_slice64: workspace %zu < %zu", workspace_bytes, need);
  if (slice == 0) return MCCB200_OK;
  DerivedKeys* dk = (DerivedKeys*)workspace;
  phase_mark("mc_indices", st);
  derive_keys_kernel<<<1, 32, 0, st>>>(*rng, 0, dk);
  int grid = (int)min((uint64_t)(slice + 255) / 256, (uint64_t)sm_count() * (side_ctas() > 0 ? side_ctas() : 8));
  randint64_kernel<<<grid, 256, 0, st>>>(dk, (uint32_t)count, n_inputs, rng->partitionable, idx_out, (uint32_t)first,
                                         (uint32_t)slice);
  phase_mark("end", st);
  count_launches(2);
  return check_launch("mc_indices_slice64");
}
