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
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < slice; q += stride) {
    const uint32_t i = first + q;
    uint32_t hi = random_bits32_at(k1, i, count, part);
    uint32_t lo = random_bits32_at(k2, i, count, part);
    uint32_t off = (hi % span) * mult + (lo % span);
    out[q] = off % span;
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
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < slice; q += stride) {
    const uint32_t i = first + q;
    const uint64_t hi = random_bits64_at(k1, i, count, part);
    const uint64_t lo = random_bits64_at(k2, i, count, part);
    const uint64_t off = (hi % span) * mult + (lo % span);
    out[q] = off % span;
  }
}

__global__ void __launch_bounds__(256) iota_kernel(uint32_t count, uint32_t* __restrict__ out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) out[i] = i;
}

// _shuffle for small inputs in ONE launch: a single CTA derives the keys, then per round draws
// the 32-bit sort keys and sorts (key << 32 | position) with a shared-memory bitonic network --
// positions are unique, so the unstable network reproduces the stable sort_key_val exactly.
constexpr int SMALL_SHUFFLE_MAX = 8192;

__global__ void __launch_bounds__(1024)
small_shuffle_kernel(mccb200_rng rng, int rounds, uint32_t n, uint32_t npad, uint32_t count, uint32_t* __restrict__ out) {
  extern __shared__ unsigned long long s_comp[];        // [npad]
  uint32_t* xa = reinterpret_cast<uint32_t*>(s_comp + npad);   // [npad]
  uint32_t* xb = xa + npad;                                      // [npad]
  __shared__ u32x2 s_sub[MAX_ROUNDS];
  const bool part = rng.partitionable != 0;
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
  for (uint32_t i = threadIdx.x; i < npad; i += blockDim.x) xa[i] = i;
  __syncthreads();
  for (int r = 0; r < rounds; ++r) {
    const u32x2 sub = s_sub[r];
    for (uint32_t i = threadIdx.x; i < npad; i += blockDim.x)
      s_comp[i] = i < n ? (((unsigned long long)random_bits32_at(sub, i, n, part) << 32) | i) : ~0ull;
    __syncthreads();
    for (uint32_t k = 2; k <= npad; k <<= 1) {
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        for (uint32_t i = threadIdx.x; i < npad; i += blockDim.x) {
          const uint32_t ixj = i ^ j;
          if (ixj > i) {
            const unsigned long long a = s_comp[i], b = s_comp[ixj];
            const bool up = (i & k) == 0;
            if ((a > b) == up) { s_comp[i] = b; s_comp[ixj] = a; }
          }
        }
        __syncthreads();
      }
    }
    for (uint32_t q = threadIdx.x; q < n; q += blockDim.x) xb[q] = xa[(uint32_t)s_comp[q]];
    __syncthreads();
    uint32_t* t = xa; xa = xb; xb = t;
  }
  for (uint32_t q = threadIdx.x; q < count; q += blockDim.x) out[q] = xa[q];
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
  if (!replace) {
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
    uint32_t npad = 2;
    while (npad < n_inputs) npad <<= 1;
    const size_t smem = (size_t)npad * 16;
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(small_shuffle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMALL_SHUFFLE_MAX * 16);
      attr_set = true;
    }
    phase_mark("mc_indices", st);
    small_shuffle_kernel<<<1, 1024, smem, st>>>(*rng, p.rounds, (uint32_t)n_inputs, npad, (uint32_t)count, idx_out);
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
  int grid = (int)min((uint64_t)(slice + 255) / 256, (uint64_t)sm_count() * 8);
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
  int grid = (int)min((uint64_t)(slice + 255) / 256, (uint64_t)sm_count() * 8);
  randint64_kernel<<<grid, 256, 0, st>>>(dk, (uint32_t)count, n_inputs, rng->partitionable, idx_out, (uint32_t)first,
                                         (uint32_t)slice);
  phase_mark("end", st);
  count_launches(2);
  return check_launch("mc_indices_slice64");
}
