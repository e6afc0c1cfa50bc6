// Fused BoxPartitionKernel + PartitioningRecombinationKernel + MonteCarloKernel step on the
// IMPLICIT children of an MCC expansion (sm_100a).
//
// Reference data flow being replaced (/root/reference/mccube/_solvers.py:132-144 ->
// _kernels/base.py:152-177): build all n*k children, partition them (stable sort by key,
// equal chunks), gather the SAME local positions idx_local from every partition (quirk Q4).
// Here nothing of size n*k is ever stored, and no work is done per child:
//
// Child key algebra.  With cell^-_a / cell^+_a the cells of the two possible child
// coordinates in key dim a, key(j) = K0 + (D & FM_j): K0 packs cell^-, D packs
// (cell^+ - cell^-), FM_j is the field mask of the dims where cubature point j has a + sign
// (bit fields never carry).  Let S = {a : cell^+_a != cell^-_a} (the parent's split set).
// Children j, j' of one parent share a key  <=>  their signs agree on S.  So the k children
// fall into at most 2^|S| CLASSES; in the stable order (key, child id) every class occupies a
// run of consecutive ranks, members ordered by j.  Which points form which class depends
// only on (S, cubature, key dims): tables built once per (k, dims, bits) (box_tables_kernel).
//
//   tables     box_tables_kernel     : class tables (data independent, cacheable by the caller)
//   selection  box_selection_kernel  : prefix count over the sorted local positions idx_local
//   pass A     box_count_flat_kernel : per-segment histogram of child keys + packed parent
//              descriptors (K0 | D | S), one shared-memory atomic per class RECORD
//   scan       box_scan_*            : start[seg][key] = same key in earlier segments;
//              key_start[key] = children with a smaller key  ("integer key-sort + scan")
//   pass B     box_rank_flat_kernel  : one warp per segment walks its parents in order and
//              hands every class record its rank run [rank0, rank0 + size); the prefix count
//              tells in two loads whether the run holds a selected position; only then are the
//              hit members resolved and their child ids i*k + j written to
//              idx_out[partition * out_per_part + l]   (4 bytes per selected child)
//   gather     expand_kernel         : the HBM-bound fused expand+select of expand_select.cu
//              materialises exactly those n children: read n rows + write n rows.
//
// Earlier forms of pass B (class-lane with shared counter rows, key sweep, per-batch hash
// table) are in the git history; their measurements are in profiles/README.md.
#include <stdlib.h>

#include "common.cuh"
#include "expand.cuh"

namespace mccb {

constexpr int BP_MAX_KEY_BITS = 12;
constexpr int BP_MAX_K = 1024;
constexpr uint32_t BP_INVALID = 0xFFFFFFFFu;

int run_expand_select(ExpandArgs& a, cudaStream_t st);

// what the class tables depend on
struct TableSpec {
  BoxSpec spec;
  uint32_t k, half_mask;
  uint32_t n_sets;          // 2^n_dims split sets
  uint32_t max_cls;         // classes per split set = min(k, 2^n_dims)
  uint32_t cls_pad_log2;    // log2 of the padded class count (row stride of cls_packed)
};

struct BoxPrkParams {
  ExpandArgs a;
  BoxSpec spec;
  const float* bounds;
  uint32_t n_keys;          // 2^(n_dims*bits)
  uint32_t n_segments;
  uint64_t parents_per_seg;
  uint32_t in_per_part, out_per_part;
  uint32_t in_shift;        // log2(in_per_part) or 0xFFFFFFFF
  uint32_t n_sets, max_cls, cls_pad_log2;
  // class tables, indexed [S][...]
  const uint32_t* cls_count;   // [n_sets]
  const uint32_t* cls_packed;  // [n_sets][cls_pad]   fm | (size-1) << 12 | off << 22, or BP_INVALID
  const uint16_t* members;     // [n_sets][k]         points j sorted by (class, j)
  // selection tables
  const uint32_t* pre;         // [in_per_part + sel_ext + 1]  #selected entries with position < pos (extended)
  const uint2* sp;             // [2 * out_per_part]  (position, l) sorted by position (extended)
  uint32_t sel_ext;
  uint32_t* pdesc;             // [n] per-parent K0 | D << 12 | S << 24 (written by pass A)
  uint32_t* table;             // [n_segments + 1][n_keys]
  const uint32_t* key_start;   // [n_keys] children with a smaller key (rows stay relative to it)
  uint32_t* idx_out;           // [n_out] selected child ids
  int warps_per_cta;
};

__device__ __forceinline__ uint32_t box_cell_u(float x, float lo, float inv_h, int bits) {
  float v = __fmul_rn(__fsub_rn(x, lo), inv_h);
  float c = fminf(fmaxf(floorf(v), 0.0f), (float)((1 << bits) - 1));
  return (uint32_t)c;
}

// sign pattern of cubature point j over the key dims: bit q set <=> + sign in key dim q
__device__ __forceinline__ uint32_t sign_pattern(const TableSpec& t, uint32_t j) {
  uint32_t rho = 0;
  for (int q = 0; q < t.spec.n_dims; ++q)
    if (!hadamard_neg(j, t.half_mask, (uint32_t)t.spec.dims[q])) rho |= 1u << q;
  return rho;
}

__device__ __forceinline__ uint32_t field_mask_of_pattern(const TableSpec& t, uint32_t rho) {
  uint32_t fm = 0;
  const uint32_t field = (1u << t.spec.bits) - 1u;
  for (int q = 0; q < t.spec.n_dims; ++q)
    if ((rho >> q) & 1u) fm |= field << (t.spec.bits * (t.spec.n_dims - 1 - q));
  return fm;
}

// (K0, D, S) of parent i from its key dims only (any key dims; the vectorised form is below)
__device__ __forceinline__ void describe_parent(const BoxPrkParams& p, uint64_t i, uint32_t& K0, uint32_t& D,
                                                uint32_t& S) {
  const ExpandArgs& a = p.a;
  K0 = 0; D = 0; S = 0;
  for (int q = 0; q < p.spec.n_dims; ++q) {
    const int col = p.spec.dims[q];
    const float lo = __ldg(p.bounds + q), ih = __ldg(p.bounds + p.spec.n_dims + q);
    const float y = ld_strided(a.y0 + i * a.ld + col);
    const float fv = a.drift_kind == 1 ? ld_strided(a.f + i * a.d + col) : drift_value(a, i, col, y);
    const uint32_t cm = box_cell_u(child_value(a, y, fv, -a.scale), lo, ih, p.spec.bits);
    const uint32_t cp = box_cell_u(child_value(a, y, fv, a.scale), lo, ih, p.spec.bits);
    K0 = (K0 << p.spec.bits) | cm;
    D = (D << p.spec.bits) | (cp - cm);
    if (cp != cm) S |= 1u << q;
  }
}

// ------------------------------------------------------------------ tables
// One thread per split set S: classes in order of first appearance, members sorted by j.
__global__ void box_tables_kernel(const TableSpec p, uint32_t* cls_count, uint32_t* cls_fm, uint32_t* cls_size,
                                  uint32_t* cls_off, uint32_t* cls_packed, uint16_t* members,
                                  uint16_t* scratch_cid /*[n_sets][n_sets]*/) {
  const uint32_t S = blockIdx.x * blockDim.x + threadIdx.x;
  if (S >= p.n_sets) return;
  const uint32_t k = p.k;
  uint16_t* cid = scratch_cid + (size_t)S * p.n_sets;
  uint32_t* fm = cls_fm + (size_t)S * p.max_cls;
  uint32_t* size = cls_size + (size_t)S * p.max_cls;
  uint32_t* off = cls_off + (size_t)S * p.max_cls;
  uint16_t* mem = members + (size_t)S * k;
  // members per restricted sign pattern sig = rho & S
  for (uint32_t s = 0; s < p.n_sets; ++s) cid[s] = 0u;
  for (uint32_t j = 0; j < k; ++j) cid[sign_pattern(p, j) & S] += 1;
  // class ids in order of increasing key offset D & FM(sig): key dim 0 owns the most significant
  // field, so that is the order of sig with its n_dims bits reversed
  uint32_t count = 0;
  for (uint32_t r = 0; r < p.n_sets; ++r) {
    uint32_t sig = 0;
    for (int q = 0; q < p.spec.n_dims; ++q)
      if ((r >> (p.spec.n_dims - 1 - q)) & 1u) sig |= 1u << q;
    const uint32_t members_of_sig = cid[sig];
    cid[sig] = 0xFFFFu;
    if (members_of_sig == 0u || (sig & ~S) != 0u) continue;
    cid[sig] = (uint16_t)count;
    fm[count] = field_mask_of_pattern(p, sig);
    size[count] = members_of_sig;
    ++count;
  }
  cls_count[S] = count;
  uint32_t run = 0;
  for (uint32_t c = 0; c < count; ++c) { off[c] = run; run += size[c]; size[c] = 0; }
  for (uint32_t j = 0; j < k; ++j) {
    const uint32_t c = cid[sign_pattern(p, j) & S];
    mem[off[c] + size[c]] = (uint16_t)j;
    size[c] += 1;
  }
  const uint32_t cls_pad = 1u << p.cls_pad_log2;
  for (uint32_t c = 0; c < cls_pad; ++c)
    cls_packed[(size_t)S * cls_pad + c] = c < count ? (fm[c] | ((size[c] - 1u) << 12) | (off[c] << 22)) : BP_INVALID;
}

// pre[pos] = #l with idx_local[l] < pos for pos in [0, in_per_part]; sp = (pos, l) sorted by pos.
// Both are EXTENDED by one wrap so that a class run crossing a partition boundary needs no special
// case: pre[in_per_part + x] = out_per_part + pre[x] for x in (0, ext], and
// sp[out_per_part + e] = (sp[e].pos + in_per_part, sp[e].l + out_per_part).
// Single CTA (the selection vector is tiny next to the state).
__global__ void __launch_bounds__(1024)
box_selection_kernel(const uint32_t* __restrict__ idx_local, uint32_t out_per_part, uint32_t in_per_part,
                     uint32_t ext, uint32_t* __restrict__ pre, uint32_t* __restrict__ fill, uint2* __restrict__ sp) {
  __shared__ uint32_t wtot[32];
  __shared__ uint32_t carry_s;
  const uint32_t len = in_per_part + 1;
  for (uint32_t i = threadIdx.x; i < len; i += 1024) { pre[i] = 0u; fill[i] = 0u; }
  __syncthreads();
  for (uint32_t l = threadIdx.x; l < out_per_part; l += 1024) atomicAdd(&pre[idx_local[l]], 1u);
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < len; base += 1024) {  // exclusive scan
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < len ? pre[i] : 0u;
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t t = wtot[threadIdx.x], sc = t;
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, sc, o);
        if (threadIdx.x >= o) sc += y;
      }
      wtot[threadIdx.x] = sc - t;
    }
    __syncthreads();
    const uint32_t carry = carry_s;
    const uint32_t incl = x + wtot[threadIdx.x >> 5];
    if (i < len) pre[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + incl;
    __syncthreads();
  }
  for (uint32_t l = threadIdx.x; l < out_per_part; l += 1024) {
    const uint32_t pos = idx_local[l];
    const uint32_t slot = pre[pos] + atomicAdd(&fill[pos], 1u);  // duplicates: any order (same child)
    sp[slot] = make_uint2(pos, l);
    sp[out_per_part + slot] = make_uint2(pos + in_per_part, l + out_per_part);
  }
  for (uint32_t x = 1 + threadIdx.x; x <= ext; x += 1024) pre[in_per_part + x] = out_per_part + pre[x];
}

// ------------------------------------------------------------------ scan
// table[seg][key] -> exclusive prefix over segments (per key); key_tot[key] = column total.
// CTA = SC_KQ key groups of W keys x SC_CY segment chunks; a thread owns W consecutive keys (one
// 16-byte access when W = 4) of its chunk's rows: independent loads, then a serial add per row.
constexpr int SC_KQ = 4;     // (8 x 128: full 128-byte lines per row, measured no faster)
constexpr int SC_CY = 256;

template <int W>
struct ScanVec { uint32_t v[W]; };

template <int W>
__device__ __forceinline__ ScanVec<W> scan_ld(const uint32_t* p) {
  ScanVec<W> r;
  if (W == 4) { const uint4 t = __ldcg(reinterpret_cast<const uint4*>(p)); r.v[0] = t.x; r.v[1 % W] = t.y; r.v[2 % W] = t.z; r.v[3 % W] = t.w; }
  else r.v[0] = __ldcg(p);
  return r;
}
template <int W>
__device__ __forceinline__ void scan_st(uint32_t* p, const ScanVec<W>& r) {
  if (W == 4) __stcg(reinterpret_cast<uint4*>(p), make_uint4(r.v[0], r.v[1 % W], r.v[2 % W], r.v[3 % W]));
  else __stcg(p, r.v[0]);
}

template <int W>
__global__ void __launch_bounds__(SC_KQ * SC_CY)
box_scan_columns_kernel(uint32_t* __restrict__ table, uint32_t n_segments, uint32_t n_keys,
                        uint32_t* __restrict__ key_tot) {
  __shared__ uint32_t part[W][SC_KQ][SC_CY + 1];
  __shared__ uint32_t wtot[W][SC_KQ][SC_CY / 32];
  const uint32_t kx = threadIdx.x % SC_KQ, cy = threadIdx.x / SC_KQ;
  const uint32_t key = (blockIdx.x * SC_KQ + kx) * W;
  const bool ok = key < n_keys;
  const uint32_t chunk = (n_segments + SC_CY - 1) / SC_CY;
  const uint32_t s0 = min(n_segments, cy * chunk), s1 = min(n_segments, s0 + chunk);
  ScanVec<W> sum;
#pragma unroll
  for (int w = 0; w < W; ++w) sum.v[w] = 0;
  if (ok) {
#pragma unroll 8
    for (uint32_t sgm = s0; sgm < s1; ++sgm) {
      const ScanVec<W> c = scan_ld<W>(table + (uint64_t)sgm * n_keys + key);
#pragma unroll
      for (int w = 0; w < W; ++w) sum.v[w] += c.v[w];
    }
  }
#pragma unroll
  for (int w = 0; w < W; ++w) part[w][kx][cy] = sum.v[w];
  __syncthreads();
  // exclusive scan over the chunks of every (w, kx) column: warp = 32 consecutive chunks
  {
    const int lane = threadIdx.x & 31;
    for (uint32_t col = threadIdx.x >> 5; col < (uint32_t)(W * SC_KQ * (SC_CY / 32)); col += (SC_KQ * SC_CY) / 32) {
      const uint32_t w = col / (SC_KQ * (SC_CY / 32)), rem = col % (SC_KQ * (SC_CY / 32));
      const uint32_t kq = rem / (SC_CY / 32), blk = rem % (SC_CY / 32);
      const uint32_t v = part[w][kq][blk * 32 + lane];
      uint32_t x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      part[w][kq][blk * 32 + lane] = x - v;
      if (lane == 31) wtot[w][kq][blk] = x;
    }
  }
  __syncthreads();
  ScanVec<W> run;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    uint32_t r = part[w][kx][cy];
    for (uint32_t bq = 0; bq < cy / 32; ++bq) r += wtot[w][kx][bq];
    run.v[w] = r;
  }
  if (ok) {
#pragma unroll 8
    for (uint32_t sgm = s0; sgm < s1; ++sgm) {
      uint32_t* ptr = table + (uint64_t)sgm * n_keys + key;
      const ScanVec<W> c = scan_ld<W>(ptr);
      scan_st<W>(ptr, run);
#pragma unroll
      for (int w = 0; w < W; ++w) run.v[w] += c.v[w];
    }
    if (cy == SC_CY - 1) {
#pragma unroll
      for (int w = 0; w < W; ++w) key_tot[key + w] = run.v[w];
    }
  }
}

// single CTA: exclusive scan of key_tot (n_keys <= 4096)
__global__ void __launch_bounds__(1024)
box_scan_keys_kernel(uint32_t* __restrict__ key_tot, uint32_t n_keys) {
  __shared__ uint32_t wtot[32];
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_keys; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < n_keys ? key_tot[i] : 0u;
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t t = wtot[threadIdx.x], sc = t;
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, sc, o);
        if (threadIdx.x >= o) sc += y;
      }
      wtot[threadIdx.x] = sc - t;
    }
    __syncthreads();
    const uint32_t carry = carry_s;
    const uint32_t incl = x + wtot[threadIdx.x >> 5];
    if (i < n_keys) key_tot[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + incl;
    __syncthreads();
  }
}

// ------------------------------------------------------------------ flat record machinery
// A batch = 32 consecutive parents, one per lane, parent `lane` holding ncls class records.
// The records of the batch are numbered parent-major (excl = exclusive scan of ncls) and
// processed 32 per ROUND with every lane busy: record r0 + lane belongs to parent
// `owner` = (#parents whose first record is <= r) - 1.  Inside a round, __match_any_sync on the
// key gives each record its peers; records of one key are in parent order both inside a round
// (lane order) and across rounds (earlier rounds hold earlier parents), and one parent never
// holds a key twice, so "start of the key + sizes of the lower peers" is the stable rank.
// Class sizes are counted in UNITS of k >> cls_pad_log2 (the smallest class): a parent with
// 2^lsz classes has classes of 2^e units, e = cls_pad_log2 - lsz.
constexpr uint32_t FLAT_TOFF_BIAS = 1024;  // >= 32 parents * 32 classes

struct FlatBatch {
  uint32_t desc;   // K0 | D << 12 | S << 24 of this lane's parent
  uint32_t pk;     // (S << cls_pad_log2) + FLAT_TOFF_BIAS - excl  |  e << 12
  uint32_t ncls, excl, R;
};

__device__ __forceinline__ FlatBatch flat_batch(uint32_t desc, uint32_t ncls, int lane, uint32_t cls_pad_log2) {
  FlatBatch b;
  uint32_t incl = ncls;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  b.desc = desc; b.ncls = ncls; b.excl = incl - ncls;
  b.R = __shfl_sync(0xffffffffu, incl, 31);
  const uint32_t e = ncls ? cls_pad_log2 - (uint32_t)(__ffs(ncls) - 1) : 0u;
  b.pk = (((desc >> 24) << cls_pad_log2) + FLAT_TOFF_BIAS - b.excl) | (e << 12);
  return b;
}

struct FlatRecord {
  bool valid;
  int owner;                 // lane holding the record's parent
  uint32_t S, ent, key, e;   // split set, packed class entry, box key, log2(class size in units)
};

// s_tab_biased = s_tab - FLAT_TOFF_BIAS
__device__ __forceinline__ FlatRecord flat_record(const FlatBatch& b, uint32_t r0, int lane,
                                                  const uint32_t* s_tab_biased) {
  FlatRecord rec;
  const uint32_t r = r0 + (uint32_t)lane;
  rec.valid = r < b.R;
  const uint32_t rel = b.excl - r0;                     // wraps (>= 32) when the parent starts earlier
  // (parents without records are the lanes past the end of the segment; their excl == R marks a
  // bit above every valid record of the last round and is never < r0, so they need no guard)
  const uint32_t starts = __reduce_or_sync(0xffffffffu, rel < 32u ? (1u << rel) : 0u);
  const uint32_t before = (uint32_t)__popc(__ballot_sync(0xffffffffu, b.excl < r0));
  const uint32_t le_mask = 0xffffffffu >> (31 - lane);
  rec.owner = (int)(before + (uint32_t)__popc(starts & le_mask)) - 1;   // shuffles take it mod 32
  const uint32_t od = __shfl_sync(0xffffffffu, b.desc, rec.owner);
  const uint32_t opk = __shfl_sync(0xffffffffu, b.pk, rec.owner);
  rec.S = od >> 24;
  rec.ent = rec.valid ? s_tab_biased[(opk & 0xFFFu) + r] : 0u;
  rec.key = rec.valid ? (od & 0xFFFu) + ((od >> 12) & rec.ent & 0xFFFu) : 0xFFFFFFFFu;
  rec.e = rec.valid ? (opk >> 12) : 31u;
  return rec;
}

// Units held by the peers in lower lanes (exu) and by all peers (totu).
__device__ __forceinline__ void peer_units(uint32_t peers, uint32_t e, uint32_t cls_pad_log2, uint32_t lt_mask,
                                           uint32_t& exu, uint32_t& totu) {
  const uint32_t P = peers & lt_mask;
  exu = 0;
#pragma unroll
  for (uint32_t b = 0; b < 6; ++b) {
    if (b <= cls_pad_log2) {                             // warp-uniform
      const uint32_t m = __ballot_sync(0xffffffffu, e == b);
      exu += (uint32_t)__popc(P & m) << b;
    }
  }
  totu = __shfl_sync(0xffffffffu, exu + (1u << (e & 7u)), 31 - __clz(peers));   // the highest peer knows the total
}

// (K0, D, S) from the two candidate child coordinates in 4 vectorised key dims
__device__ __forceinline__ uint32_t describe_parent4(const BoxPrkParams& p, const KeyConst4& kc, uint64_t i, const float4 y4,
                                                     const float (&lo)[4], const float (&ih)[4]) {
  float vm[4], vp[4];
  key_children4(p.a, p.spec, kc, i, y4, vm, vp);
  uint32_t K0 = 0, D = 0, S = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t cm = box_cell_u(vm[q], lo[q], ih[q], p.spec.bits);
    const uint32_t cp = box_cell_u(vp[q], lo[q], ih[q], p.spec.bits);
    K0 = (K0 << p.spec.bits) | cm;
    D = (D << p.spec.bits) | (cp - cm);
    if (cp != cm) S |= 1u << q;
  }
  return K0 | (D << 12) | (S << 24);
}

// ------------------------------------------------------------------ pass A
// CTA per segment, warps interleave its batches; per batch the lanes are grouped by descriptor
// (__match_any_sync) and one lane per group feeds the shared-memory histogram.
template <bool VEC4>
__global__ void __launch_bounds__(256)
box_count_flat_kernel(const BoxPrkParams p) {
  extern __shared__ uint32_t smem[];
  const ExpandArgs& a = p.a;
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t* hist = smem;
  const uint32_t cls_pad = 1u << p.cls_pad_log2;
  uint32_t* s_tab = hist + p.n_keys;
  uint32_t* s_cnt = s_tab + p.n_sets * cls_pad;
  for (uint32_t i = threadIdx.x; i < p.n_keys; i += blockDim.x) hist[i] = 0u;
  for (uint32_t i = threadIdx.x; i < p.n_sets * cls_pad; i += blockDim.x) s_tab[i] = p.cls_packed[i];
  for (uint32_t i = threadIdx.x; i < p.n_sets; i += blockDim.x) s_cnt[i] = p.cls_count[i];
  __syncthreads();
  float lo[4] = {0.f, 0.f, 0.f, 0.f}, ih[4] = {0.f, 0.f, 0.f, 0.f};
  if (VEC4) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { lo[q] = __ldg(p.bounds + q); ih[q] = __ldg(p.bounds + 4 + q); }
  }
  const uint32_t seg = blockIdx.x;
  const uint64_t i0 = (uint64_t)seg * p.parents_per_seg;
  const uint64_t i1 = min(a.n, i0 + p.parents_per_seg);
  KeyConst4 kc{};
  if (VEC4) kc = key_const4(a, p.spec);
  float4 y4_next = make_float4(0.f, 0.f, 0.f, 0.f);    // the key columns of the next batch are in flight during this one
  {
    const uint64_t i = i0 + (threadIdx.x & ~31u) + lane;
    if (VEC4 && i < i1) y4_next = key_columns4(a, p.spec, i);
  }
  for (uint64_t base = i0 + (threadIdx.x & ~31u); base < i1; base += blockDim.x) {
    const uint64_t i = base + lane;
    const bool have = i < i1;
    const float4 y4 = y4_next;
    if (VEC4 && i + blockDim.x < i1) y4_next = key_columns4(a, p.spec, i + blockDim.x);
    uint32_t desc = 0;
    if (have) {
      if (VEC4) desc = describe_parent4(p, kc, i, y4, lo, ih);
      else { uint32_t K0, D, S; describe_parent(p, i, K0, D, S); desc = K0 | (D << 12) | (S << 24); }
      p.pdesc[i] = desc;                                 // pass B re-reads 4 coalesced bytes per parent
    }
    // Parents with the same descriptor (K0, D, S) contribute the same classes: one lane per distinct
    // descriptor of the batch adds (group size x class size) for each of its classes.  On box-ordered
    // input a batch holds a dozen distinct descriptors for ~140 class records.
    const uint32_t peers = __match_any_sync(0xffffffffu, have ? desc : 0xFFFFFFFFu);
    if (have && (peers & lt_mask) == 0u) {
      const uint32_t mult = (uint32_t)__popc(peers);
      const uint32_t S = desc >> 24, K0 = desc & 0xFFFu, D = (desc >> 12) & 0xFFFu;
      const uint32_t ncls = s_cnt[S];
      const uint32_t* trow = s_tab + (S << p.cls_pad_log2);
      // The kernel is bound by instruction issue (84 % active), and this loop -- ~11 trips in a third of the lanes --
      // is 45 % of it.  Measured against its 0.202 ms: class entries preloaded (16 predicated adds 0.209, four per
      // 128-bit load 0.201); distinct descriptors packed into the first lanes and their class records added 32 per
      // round with the record numbering of pass B 0.209.  None kept.
      for (uint32_t c = 0; c < ncls; ++c) {
        const uint32_t ent = trow[c];
        atomicAdd(&hist[K0 + (D & ent & 0xFFFu)], mult * (((ent >> 12) & 0x3FFu) + 1u));
      }
    }
  }
  __syncthreads();
  uint32_t* row = p.table + (uint64_t)seg * p.n_keys;
  for (uint32_t i = threadIdx.x; i < p.n_keys; i += blockDim.x) row[i] = hist[i];
}

// ------------------------------------------------------------------ pass B (flat form)
// One warp per segment walks its batches in order.  Per round: one L2 atomicAdd per distinct
// key on the segment's (relative) counter row hands the key's records their rank run.  The
// atomic of round r+1 is issued before the result of round r is consumed (same-address atomics
// of one warp reach L2 in program order), so its round trip overlaps the selection tests.
// Records whose run holds a selected position (about one in five) are queued in shared memory
// and resolved 32 at a time with every lane busy.
constexpr int FQ_CAP = 64;   // hit-record queue entries per warp (ring)

struct FlatRound {
  uint32_t st, key, exu, e, child0, moff;
  int leader;
  bool valid;
};

// POW2: in_per_part is a power of two (the runtime test made the compiler evaluate the ~20-instruction
// integer modulo for every record and select afterwards: 11 % of the kernel's instructions).
template <bool PRE_SMEM, bool MEM_SMEM, bool POW2>
__global__ void __launch_bounds__(256, 4)
box_rank_flat_kernel(const BoxPrkParams p) {
  extern __shared__ uint32_t smem[];
  const ExpandArgs& a = p.a;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t* cursor = smem;
  uint4* hq = reinterpret_cast<uint4*>(cursor) + (size_t)warp * FQ_CAP;   // (rank0, first, n_hit | moff << 11, child0)
  cursor += (size_t)p.warps_per_cta * FQ_CAP * 4;
  const uint32_t cls_pad = 1u << p.cls_pad_log2;
  uint32_t* s_tab = cursor; cursor += p.n_sets * cls_pad;
  uint32_t* s_cnt = cursor; cursor += p.n_sets;
  for (uint32_t i = threadIdx.x; i < p.n_sets * cls_pad; i += blockDim.x) s_tab[i] = p.cls_packed[i];
  for (uint32_t i = threadIdx.x; i < p.n_sets; i += blockDim.x) s_cnt[i] = p.cls_count[i];
  const uint2* sp = p.sp;
  uint16_t* s_pre16 = nullptr;
  if (PRE_SMEM) {
    cursor += ((cursor - smem) & 1u);                    // 8-byte alignment for sp
    uint2* s_sp = reinterpret_cast<uint2*>(cursor);
    cursor += 4u * p.out_per_part;
    s_pre16 = reinterpret_cast<uint16_t*>(cursor);
    cursor += ((p.in_per_part + p.sel_ext + 4u) & ~3u) / 2u;
    for (uint32_t i = threadIdx.x; i <= p.in_per_part + p.sel_ext; i += blockDim.x) s_pre16[i] = (uint16_t)p.pre[i];
    for (uint32_t i = threadIdx.x; i < 2u * p.out_per_part; i += blockDim.x) s_sp[i] = p.sp[i];
    sp = s_sp;
  }
  auto PRE = [&](uint32_t i) -> uint32_t { return PRE_SMEM ? (uint32_t)s_pre16[i] : __ldg(p.pre + i); };
  const uint16_t* members = p.members;
  if (MEM_SMEM) {
    uint16_t* s_mem = reinterpret_cast<uint16_t*>(cursor);
    for (uint32_t i = threadIdx.x; i < p.n_sets * a.k; i += blockDim.x) s_mem[i] = p.members[i];
    members = s_mem;
  }
  __syncthreads();
  const uint32_t seg = blockIdx.x * p.warps_per_cta + warp;
  if (seg >= p.n_segments) return;
  uint32_t* row = p.table + (uint64_t)seg * p.n_keys;    // same key, earlier segments (relative ranks)
  const uint32_t* s_tab_biased = s_tab - FLAT_TOFF_BIAS;

  const uint64_t i0 = (uint64_t)seg * p.parents_per_seg;
  const uint64_t i1 = min(a.n, i0 + p.parents_per_seg);
  const uint32_t Pm = p.in_per_part, nm = p.out_per_part;
  const int logk = __ffs(a.k) - 1;
  const uint32_t ushift = (uint32_t)logk - p.cls_pad_log2;   // log2(unit)
  uint32_t q_head = 0, q_count = 0;                      // warp-uniform ring state

  // resolve queued hit records [q_head, q_head + cnt)
  auto drain = [&](uint32_t cnt) {
    if ((uint32_t)lane < cnt) {
      const uint4 e4 = hq[(q_head + (uint32_t)lane) & (FQ_CAP - 1)];
      const uint32_t rank0 = e4.x, first = e4.y, n_hit = e4.z & 0x7FFu;
      uint32_t part0, pos0;
      if (POW2) { part0 = rank0 >> p.in_shift; pos0 = rank0 & (Pm - 1u); }
      else { part0 = rank0 / Pm; pos0 = rank0 - part0 * Pm; }
      const uint16_t* mem = members + (e4.z >> 11) - pos0;
      uint32_t* out0 = p.idx_out + (uint64_t)part0 * nm;
      for (uint32_t e = first; e < first + n_hit; ++e) {
        const uint2 pl = sp[e];                          // (position, output slot l), extended past one wrap
        out0[pl.y] = e4.w + mem[pl.x];
      }
    }
    q_head = (q_head + cnt) & (FQ_CAP - 1);
    q_count -= cnt;
    __syncwarp();
  };

  // front half of a round: record, peers, unit prefix, the key's atomic (result left in flight)
  auto front = [&](const FlatBatch& b, uint64_t base, uint32_t r0, FlatRound& o) {
    const FlatRecord rec = flat_record(b, r0, lane, s_tab_biased);
    const uint32_t peers = __match_any_sync(0xffffffffu, rec.key);
    uint32_t exu, totu;
    peer_units(peers, rec.e, p.cls_pad_log2, lt_mask, exu, totu);
    o.leader = __ffs(peers) - 1;
    o.st = 0;
    if (rec.valid && lane == o.leader) o.st = atomicAdd(row + rec.key, totu << ushift);
    o.valid = rec.valid; o.key = rec.key; o.exu = exu; o.e = rec.e;
    o.child0 = ((uint32_t)base + (uint32_t)(rec.owner & 31)) << logk;   // child ids are uint32 (n * k <= 2^32)
    o.moff = (rec.S << logk) + (rec.ent >> 22);
  };
  // back half: rank, selection test, queue the hit records
  auto back = [&](const FlatRound& o) {
    const uint32_t st = __shfl_sync(0xffffffffu, o.st, o.leader);
    uint32_t n_hit = 0, rank0 = 0, first = 0;
    if (o.valid) {
      rank0 = st + __ldg(p.key_start + o.key) + (o.exu << ushift);
      const uint32_t pos0 = POW2 ? (rank0 & (Pm - 1u)) : (rank0 % Pm);
      first = PRE(pos0);
      n_hit = PRE(pos0 + ((1u << o.e) << ushift)) - first;   // size <= k <= Pm: the extended table covers the wrap
    }
    const uint32_t hm = __ballot_sync(0xffffffffu, n_hit != 0u);
    if (hm) {
      if (n_hit) {
        const uint32_t at = (q_head + q_count + (uint32_t)__popc(hm & lt_mask)) & (FQ_CAP - 1);
        hq[at] = make_uint4(rank0, first, n_hit | (o.moff << 11), o.child0);
      }
      q_count += (uint32_t)__popc(hm);
      __syncwarp();
      if (q_count >= 32u) drain(32u);
    }
  };

  FlatRound cur, nxt;
  bool have_cur = false;
  uint32_t ndesc = i0 + lane < i1 ? __ldcs(p.pdesc + i0 + lane) : 0u;
  for (uint64_t base = i0; base < i1; base += 32) {
    const uint32_t desc = ndesc;
    const bool have = base + lane < i1;
    if (base + 32 + lane < i1) ndesc = __ldcs(p.pdesc + base + 32 + lane);   // next batch (coalesced)
    const FlatBatch b = flat_batch(desc, have ? s_cnt[desc >> 24] : 0u, lane, p.cls_pad_log2);
    for (uint32_t r0 = 0; r0 < b.R; r0 += 32) {
      __syncwarp();   // order this round's atomics after the previous round's (same-key runs)
      front(b, base, r0, nxt);
      if (have_cur) back(cur);
      cur = nxt;
      have_cur = true;
    }
  }
  if (have_cur) back(cur);
  if (q_count) drain(q_count);
}

// ------------------------------------------------------------------ pass B + gather, fused (warp-specialised)
// One persistent CTA per SM.  Warps 0 .. RW-1 run the ranking walk of box_rank_flat_kernel unchanged, but a
// selected child is not written out as an id: it becomes a ROW JOB (child id, output row) in the warp's
// shared-memory ring.  The remaining TW warps materialise the jobs: 32 at a time, every lane asks the TMA
// engine for one parent row (cp.async.bulk global -> shared, 256 bytes at d = 64, completion counted on an
// mbarrier), two stages deep, so the bytes in flight cost no registers and no issue slots; when a stage has
// landed the warp applies drift + signed noise in the reference's fp32 op order to 128-bit chunks read from
// shared memory and streams the child rows (and their packed key columns) to their final slots.  The 64 MB
// index vector and the separate HBM-bound gather pass disappear, a parent row selected twice is fetched from
// L2 the second time, and the ranking warps' dependent chains fill the issue slots the copy does not need.
// Ring protocol (shared memory, one ring of 4 x 32 jobs per ranking warp): the producer publishes `tail` after
// writing the jobs and `done` when it has finished; the consumer takes the jobs in blocks of 32 and, once a block
// is materialised, arrives on that quarter's `empty` mbarrier.  A producer entering a quarter for the n-th time
// waits for the (n-1)-th completion of its barrier with mbarrier.try_wait, which suspends the warp in hardware: a
// polling loop here (tried first, volatile loads + nanosleep, which returns after ~20 ns whatever it is asked for)
// took a third of the SM's issue slots from the transform warps.
constexpr int FZ_MAX_THREADS = 1024;
constexpr int FZ_JQ = 128;      // row-job ring entries per ranking warp
constexpr int FZ_STAGES = 2;    // TMA stages per transform warp (32 parent rows each)
constexpr int FZ_MAX_RINGS = 4; // rings served by one transform warp

struct FusedOut {
  float* y1;
  float* keycols_out;        // optional [n_out, 4]
  int keycol0;
  int vpr;                   // 128-bit chunks per row (d / 4 <= 32)
  int lpr_shift;             // log2(lanes per row) = log2(pow2 >= vpr)
  int rank_warps;            // warps 0 .. rank_warps-1 rank, the others materialise rows
  uint32_t row_bytes;        // d * 4
  uint32_t xform_words;      // shared-memory words of one transform warp (stages + barriers)
  uint32_t spin_ns;          // back-off of the ring waits
  float one, neg_one, neg_zero;   // 1, -1, -0 as RUNTIME values: ptxas must not fold the exact-rounding fmas (see f2_fma)
};

__device__ __forceinline__ uint32_t fz_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t ld_vol(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void st_vol(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }
__device__ __forceinline__ void fz_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fz_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fz_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Spins on the phase parity; traps instead of hanging the GPU if a phase never completes.
__device__ __forceinline__ void fz_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && (spin & 1023u) == 1023u) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000ll) __trap();
    }
  }
}
// L2 residency: the counter table (tens of MB, hit by one atomic per record) must stay in L2 while 8 GB of rows
// stream through it; without hints the stream evicts the table and every atomic becomes a DRAM round trip.
__device__ __forceinline__ uint64_t fz_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t fz_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// one row: global -> shared through the TMA engine (addresses and size multiples of 16 bytes)
__device__ __forceinline__ void fz_bulk_row(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void fz_st_stream4(float* p, const float4& v, uint64_t pol) {  // p 16-byte aligned
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
               ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ uint32_t fz_atom_add(uint32_t* p, uint32_t v, uint64_t pol) {
  uint32_t old;
  asm volatile("atom.global.add.L2::cache_hint.u32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(v), "l"(pol) : "memory");
  return old;
}

// Packed fp32 pairs (Blackwell FFMA2).  Every step of the reference's fp32 op order is ONE rounding, written as
// an fma whose extra operand makes it that rounding exactly -- a * b = fma(a, b, -0), a + b = fma(a, 1, b),
// m - y = fma(y, -1, m) -- with the constants 1, -1, -0 passed in as kernel arguments: given literal constants
// ptxas rewrites the fmas into mul.f32x2 / add.f32x2 and then CONTRACTS dt * f + noise into one FFMA2 (seen in
// SASS), which is a different rounding.  With runtime operands every step stays its own FFMA2 and the results are
// bit-identical to __fmul_rn / __fadd_rn / __fsub_rn (checked against the scalar kernels by the parity tests).
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

struct ChildConst {       // per-lane constants of its 128-bit column chunk
  uint64_t mean01, mean23, ivar01, ivar23, dt2, one2, mone2, nzero2;
  uint32_t cmask, top_shift, scale_bits;
};

// One 128-bit chunk of a child row in the reference's fp32 op order: y + (dt * f + noise), f = (mean - y) * inv_var,
// the noise being +-scale with the Hadamard sign of point j in columns col .. col+3 (col % 4 == 0):
// parity(j & col) ^ block ^ {0, j0, j1, j0 ^ j1}, applied to the sign bit of scale.
template <int DRIFT>
__device__ __forceinline__ float4 fz_child4(const ChildConst& c, const float4& y, uint32_t j, uint32_t half_mask) {
  const uint32_t jm = j & half_mask;
  const uint32_t s0 = c.scale_bits ^ (((uint32_t)__popc(jm & c.cmask) + (j >> c.top_shift)) << 31);
  const uint32_t b0 = jm << 31, b1 = (jm >> 1) << 31;
  const uint64_t n01 = f2_pack(__uint_as_float(s0), __uint_as_float(s0 ^ b0));
  const uint64_t n23 = f2_pack(__uint_as_float(s0 ^ b1), __uint_as_float(s0 ^ b0 ^ b1));
  const uint64_t y01 = f2_pack(y.x, y.y), y23 = f2_pack(y.z, y.w);
  uint64_t st01, st23;
  if (DRIFT == 2) {
    const uint64_t f01 = f2_fma(f2_fma(y01, c.mone2, c.mean01), c.ivar01, c.nzero2);   // (mean - y) * inv_var
    const uint64_t f23 = f2_fma(f2_fma(y23, c.mone2, c.mean23), c.ivar23, c.nzero2);
    st01 = f2_fma(f01, c.dt2, c.nzero2);                                              // dt * f
    st23 = f2_fma(f23, c.dt2, c.nzero2);
  } else {
    st01 = st23 = f2_fma(f2_pack(0.f, 0.f), c.dt2, c.nzero2);
  }
  const uint64_t o01 = f2_fma(f2_fma(st01, c.one2, n01), c.one2, y01);                 // y + (dt * f + noise)
  const uint64_t o23 = f2_fma(f2_fma(st23, c.one2, n23), c.one2, y23);
  float4 out;
  f2_unpack(o01, out.x, out.y);
  f2_unpack(o23, out.z, out.w);
  return out;
}

template <int DRIFT, bool D64>
__global__ void __launch_bounds__(FZ_MAX_THREADS, 1)
box_rank_fused_kernel(const BoxPrkParams p, const FusedOut o) {
  extern __shared__ uint32_t smem[];
  const ExpandArgs& a = p.a;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int RW = o.rank_warps, TW = (int)(blockDim.x >> 5) - RW;
  const uint32_t lt_mask = (1u << lane) - 1u;
  // ---- CTA-wide tables (as box_rank_flat_kernel<true, true, true>)
  uint32_t* cursor = smem;
  const uint32_t cls_pad = 1u << p.cls_pad_log2;
  uint32_t* s_tab = cursor; cursor += p.n_sets * cls_pad;
  uint32_t* s_cnt = cursor; cursor += p.n_sets;
  cursor += ((cursor - smem) & 1u);
  uint2* s_sp = reinterpret_cast<uint2*>(cursor); cursor += 4u * p.out_per_part;
  uint16_t* s_pre16 = reinterpret_cast<uint16_t*>(cursor); cursor += ((p.in_per_part + p.sel_ext + 4u) & ~3u) / 2u;
  uint16_t* s_mem = reinterpret_cast<uint16_t*>(cursor); cursor += (p.n_sets * a.k + 1u) / 2u;
  cursor += (4u - ((uint32_t)(cursor - smem) & 3u)) & 3u;          // 16-byte alignment of what follows
  for (uint32_t i = threadIdx.x; i < p.n_sets * cls_pad; i += blockDim.x) s_tab[i] = p.cls_packed[i];
  for (uint32_t i = threadIdx.x; i < p.n_sets; i += blockDim.x) s_cnt[i] = p.cls_count[i];
  for (uint32_t i = threadIdx.x; i <= p.in_per_part + p.sel_ext; i += blockDim.x) s_pre16[i] = (uint16_t)p.pre[i];
  for (uint32_t i = threadIdx.x; i < 2u * p.out_per_part; i += blockDim.x) s_sp[i] = p.sp[i];
  for (uint32_t i = threadIdx.x; i < p.n_sets * a.k; i += blockDim.x) s_mem[i] = p.members[i];
  // ---- per ranking warp: hit-record ring (FQ_CAP x 16 B) | job ring (FZ_JQ x 8 B) | control (4 words)
  constexpr uint32_t RANK_WORDS = FQ_CAP * 4 + FZ_JQ * 2 + 4 + 8;   // ... | control (4 words) | 4 `empty` mbarriers
  auto rank_block = [&](int w) { return cursor + (size_t)w * RANK_WORDS; };
  // ---- per transform warp: FZ_STAGES x 32 rows | FZ_STAGES mbarriers
  uint32_t* xbase = cursor + (size_t)RW * RANK_WORDS;
  if (warp < RW) {
    if (lane < 4) {
      rank_block(warp)[FQ_CAP * 4 + FZ_JQ * 2 + lane] = 0u;
      fz_mbar_init(fz_smem_u32(rank_block(warp) + FQ_CAP * 4 + FZ_JQ * 2 + 4 + 2 * lane), 1u);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (lane == 0) {
    uint32_t* xb = xbase + (size_t)(warp - RW) * o.xform_words;
    for (int s = 0; s < FZ_STAGES; ++s) fz_mbar_init(fz_smem_u32(xb + FZ_STAGES * 8u * o.row_bytes + 2 * s), 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int logk = __ffs(a.k) - 1;

  if (warp >= RW) {
    // ================= transform warps
    const int tw = warp - RW;
    uint32_t* xb = xbase + (size_t)tw * o.xform_words;
    const uint32_t stage_bytes = 32u * o.row_bytes;
    const uint32_t buf0 = fz_smem_u32(xb), bar0 = fz_smem_u32(xb + FZ_STAGES * 8u * o.row_bytes);
    const int lpr_shift = D64 ? 4 : o.lpr_shift;
    const int sub = lane & ((1 << lpr_shift) - 1), rs = lane >> lpr_shift, rpp = 32 >> lpr_shift;
    const int vpr = D64 ? 16 : o.vpr;
    const uint32_t ld = D64 ? 64u : (uint32_t)a.ld;
    const uint32_t col = (uint32_t)sub * 4u;
    ChildConst cc;
    {
      float4 mean = make_float4(0.f, 0.f, 0.f, 0.f), ivar = mean;
      if (DRIFT == 2 && sub < vpr) {
        mean = __ldg(reinterpret_cast<const float4*>(a.mean + col));
        ivar = __ldg(reinterpret_cast<const float4*>(a.inv_var + col));
      }
      cc.mean01 = f2_pack(mean.x, mean.y); cc.mean23 = f2_pack(mean.z, mean.w);
      cc.ivar01 = f2_pack(ivar.x, ivar.y); cc.ivar23 = f2_pack(ivar.z, ivar.w);
      cc.dt2 = f2_pack(a.dt, a.dt); cc.one2 = f2_pack(o.one, o.one); cc.mone2 = f2_pack(o.neg_one, o.neg_one);
      cc.nzero2 = f2_pack(o.neg_zero, o.neg_zero);
      cc.cmask = col & a.half_mask; cc.top_shift = (uint32_t)logk - 1u; cc.scale_bits = __float_as_uint(a.scale);
    }
    const uint64_t pol_stream = fz_policy_evict_first();
    uint32_t fetch[FZ_MAX_RINGS] = {0u, 0u, 0u, 0u};   // jobs handed to the TMA engine, per served ring
    uint32_t st_ring[FZ_STAGES] = {0u, 0u}, st_head[FZ_STAGES] = {0u, 0u}, st_cnt[FZ_STAGES] = {0u, 0u};
    uint32_t phase_bits = 0;                            // parity of each stage's barrier
    int in_flight = 0, next_load = 0, next_use = 0, rr = 0;
    uint32_t idle_spins = 0;
    long long idle_t0 = 0;
    for (;;) {
      // ---- hand the next block of 32 jobs to the TMA engine if a stage is free
      bool issued = false, all_done = true;
      if (in_flight < FZ_STAGES) {
#pragma unroll
        for (int qq = 0; qq < FZ_MAX_RINGS; ++qq) {
          const int q = (rr + qq) & (FZ_MAX_RINGS - 1);
          const int r = tw + q * TW;
          if (r >= RW || issued) continue;
          uint32_t* ctrl = rank_block(r) + FQ_CAP * 4 + FZ_JQ * 2;
          uint32_t done = 0, tail = 0;
          if (lane == 0) { done = ld_vol(ctrl + 2); tail = ld_vol(ctrl); }   // done first: a set flag means tail is final
          done = __shfl_sync(0xffffffffu, done, 0);
          tail = __shfl_sync(0xffffffffu, tail, 0);
          const uint32_t avail = tail - fetch[q];
          if (avail >= 32u || (done && avail)) {
            const uint32_t cnt = avail < 32u ? avail : 32u;
            __threadfence_block();
            const uint2* jq = reinterpret_cast<const uint2*>(rank_block(r) + FQ_CAP * 4);
            const uint32_t bar = bar0 + 8u * (uint32_t)next_load;
#if defined(FZ_EXP) && FZ_EXP == 4   /* timing experiment: the transform warps only consume the jobs (ranking alone, in situ) */
            if (lane == 0) fz_mbar_expect_tx(bar, 0u);
            __syncwarp();
            if (false) {
#else
            if (lane == 0) fz_mbar_expect_tx(bar, cnt * o.row_bytes);
            __syncwarp();
            if ((uint32_t)lane < cnt) {
#endif
              const uint32_t child = jq[(fetch[q] + (uint32_t)lane) & (FZ_JQ - 1)].x;
              fz_bulk_row(buf0 + (uint32_t)next_load * stage_bytes + (uint32_t)lane * o.row_bytes,
                          a.y0 + (uint64_t)(child >> logk) * ld, o.row_bytes, bar, pol_stream);
            }
            st_ring[next_load] = (uint32_t)r; st_head[next_load] = fetch[q]; st_cnt[next_load] = cnt;
            fetch[q] += cnt;
            next_load ^= 1;
            ++in_flight;
            issued = true;
            rr = (q + 1) & (FZ_MAX_RINGS - 1);
          }
          if (!(done && tail == fetch[q])) all_done = false;
        }
      } else {
        all_done = false;
      }
      // ---- materialise the oldest stage when both are busy or nothing new could be issued
      if (in_flight == FZ_STAGES || (!issued && in_flight > 0)) {
        const int s = next_use;
        fz_mbar_wait(bar0 + 8u * (uint32_t)s, (phase_bits >> s) & 1u);
        phase_bits ^= 1u << s;
        uint32_t* rb = rank_block((int)st_ring[s]);
        const uint2* jq = reinterpret_cast<const uint2*>(rb + FQ_CAP * 4);
        const uint32_t head = st_head[s], cnt = st_cnt[s];
        const char* buf = reinterpret_cast<const char*>(xb) + (size_t)s * stage_bytes;
#if defined(FZ_EXP) && FZ_EXP == 4
        if (false) {
#else
        if (sub < vpr) {
#endif
#pragma unroll 4
          for (uint32_t t = (uint32_t)rs; t < cnt; t += (uint32_t)rpp) {
            const uint2 job = jq[(head + t) & (FZ_JQ - 1)];
            const float4 y = *reinterpret_cast<const float4*>(buf + t * o.row_bytes + col * 4u);
            const float4 out = fz_child4<DRIFT>(cc, y, job.x & (a.k - 1u), a.half_mask);
#if defined(FZ_EXP) && FZ_EXP == 1   /* timing experiment: compute, no stores */
            if (out.x == 123.456f) st_stream4(o.y1, out);
#elif defined(FZ_EXP) && FZ_EXP == 3 /* timing experiment: stores into a 16 MB window (L2 resident) */
            st_stream4(o.y1 + ((uint64_t)(job.y & 0xFFFFu) * ld + col), out);
#else
            fz_st_stream4(o.y1 + ((uint64_t)job.y * ld + col), out, pol_stream);
            if (o.keycols_out != nullptr && (int)col == o.keycol0) fz_st_stream4(o.keycols_out + (uint64_t)job.y * 4u, out, pol_stream);
#endif
          }
        }
        __syncwarp();                                       // all reads of the stage are done before it is refilled
        if (lane == 0) fz_mbar_arrive(fz_smem_u32(rb + FQ_CAP * 4 + FZ_JQ * 2 + 4 + 2 * ((head >> 5) & 3u)));   // frees the quarter
        next_use ^= 1;
        --in_flight;
        continue;
      }
      if (issued) continue;
      if (all_done) break;
      __nanosleep(o.spin_ns);
      if ((++idle_spins & 4095u) == 0u) {                     // a protocol bug must trap, not hang the GPU
        if (idle_t0 == 0) idle_t0 = clock64();
        else if (clock64() - idle_t0 > 8000000000ll) __trap();
      }
    }
    return;
  }

  // ================= ranking warps (the walk of box_rank_flat_kernel; see there)
  uint32_t* wb = rank_block(warp);
  uint4* hq = reinterpret_cast<uint4*>(wb);                  // (rank0, first, n_hit | moff << 11, child0)
  uint2* jq = reinterpret_cast<uint2*>(wb + FQ_CAP * 4);
  uint32_t* ctrl = wb + FQ_CAP * 4 + FZ_JQ * 2;
  const uint32_t* s_tab_biased = s_tab - FLAT_TOFF_BIAS;
  const uint32_t Pm = p.in_per_part, nm = p.out_per_part;
  const uint32_t ushift = (uint32_t)logk - p.cls_pad_log2;   // log2(unit)
  uint32_t q_head = 0, q_count = 0;                          // hit-record ring (warp-uniform)
  uint32_t j_tail = 0;                                       // jobs published so far (warp-uniform)
  const uint64_t pol_keep = fz_policy_evict_last();

  const uint32_t empty0 = fz_smem_u32(ctrl + 4);
  uint32_t ready_block = 3;                                  // highest 32-job block of the ring known to be free
  auto wait_space = [&](uint32_t cnt) {                      // until the quarters of jobs [j_tail, j_tail + cnt) are free
    if (cnt == 0u) return;
    const uint32_t last = (j_tail + cnt - 1u) >> 5;
    while (ready_block < last) {
      ++ready_block;                                         // block b re-uses the quarter of block b - 4
      fz_mbar_wait(empty0 + 8u * (ready_block & 3u), ((ready_block >> 2) - 1u) & 1u);
    }
  };
  auto publish = [&](uint32_t cnt) {
    j_tail += cnt;
    __syncwarp();
    if (lane == 0) { __threadfence_block(); st_vol(ctrl, j_tail); }
  };
  // turn cnt (<= 32) queued hit records into row jobs
  auto drain = [&](uint32_t cnt) {
    uint32_t first = 0, n_hit = 0, child0 = 0, slot0 = 0, midx = 0;
    if ((uint32_t)lane < cnt) {
      const uint4 e4 = hq[(q_head + (uint32_t)lane) & (FQ_CAP - 1)];
      slot0 = (e4.x >> p.in_shift) * nm;
      midx = (e4.z >> 11) - (e4.x & (Pm - 1u));              // + position = index into the class's member list
      first = e4.y; n_hit = e4.z & 0x7FFu; child0 = e4.w;
    }
    q_head = (q_head + cnt) & (FQ_CAP - 1);
    q_count -= cnt;
    uint32_t hx = n_hit;                                     // inclusive scan: every lane's ring slots
#pragma unroll
    for (int sft = 1; sft < 32; sft <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, hx, sft);
      if (lane >= sft) hx += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, hx, 31);
    if (total <= 64u) {
      wait_space(total);
      uint32_t at = j_tail + hx - n_hit;
      for (uint32_t it = 0; it < n_hit; ++it, ++at) {
        const uint2 pl = s_sp[first + it];                  // (position, output slot l), extended past one wrap
        jq[at & (FZ_JQ - 1)] = make_uint2(child0 + s_mem[midx + pl.x], slot0 + pl.y);
      }
      publish(total);
    } else {                                                 // (rare: one job per lane and trip)
      for (uint32_t it = 0;; ++it) {
        const bool more = it < n_hit;
        const uint32_t mm = __ballot_sync(0xffffffffu, more);
        if (mm == 0u) break;
        wait_space(32u);
        if (more) {
          const uint2 pl = s_sp[first + it];
          jq[(j_tail + (uint32_t)__popc(mm & lt_mask)) & (FZ_JQ - 1)] = make_uint2(child0 + s_mem[midx + pl.x], slot0 + pl.y);
        }
        publish((uint32_t)__popc(mm));
      }
    }
  };

  const uint32_t total_warps = gridDim.x * (uint32_t)RW;
  for (uint32_t seg = blockIdx.x * (uint32_t)RW + (uint32_t)warp; seg < p.n_segments; seg += total_warps) {
    uint32_t* row = p.table + (uint64_t)seg * p.n_keys;    // same key, earlier segments (relative ranks)
    const uint64_t i0 = (uint64_t)seg * p.parents_per_seg;
    const uint64_t i1 = min(a.n, i0 + p.parents_per_seg);
    auto front = [&](const FlatBatch& b, uint64_t base, uint32_t r0, FlatRound& fr) {
      const FlatRecord rec = flat_record(b, r0, lane, s_tab_biased);
      const uint32_t peers = __match_any_sync(0xffffffffu, rec.key);
      uint32_t exu, totu;
      peer_units(peers, rec.e, p.cls_pad_log2, lt_mask, exu, totu);
      fr.leader = __ffs(peers) - 1;
      fr.st = 0;
      if (rec.valid && lane == fr.leader) fr.st = fz_atom_add(row + rec.key, totu << ushift, pol_keep);
      fr.valid = rec.valid; fr.key = rec.key; fr.exu = exu; fr.e = rec.e;
      fr.child0 = ((uint32_t)base + (uint32_t)(rec.owner & 31)) << logk;   // child ids are uint32 (n * k <= 2^32)
      fr.moff = (rec.S << logk) + (rec.ent >> 22);
    };
    auto back = [&](const FlatRound& fr) {
      const uint32_t st = __shfl_sync(0xffffffffu, fr.st, fr.leader);
      uint32_t n_hit = 0, rank0 = 0, first = 0;
      if (fr.valid) {
        rank0 = st + __ldg(p.key_start + fr.key) + (fr.exu << ushift);
        const uint32_t pos0 = rank0 & (Pm - 1u);
        first = s_pre16[pos0];
        n_hit = (uint32_t)s_pre16[pos0 + ((1u << fr.e) << ushift)] - first;   // size <= k <= Pm: the extended table covers the wrap
      }
      const uint32_t hm = __ballot_sync(0xffffffffu, n_hit != 0u);
      if (hm) {
        if (n_hit) {
          const uint32_t at = (q_head + q_count + (uint32_t)__popc(hm & lt_mask)) & (FQ_CAP - 1);
          hq[at] = make_uint4(rank0, first, n_hit | (fr.moff << 11), fr.child0);
        }
        q_count += (uint32_t)__popc(hm);
        __syncwarp();
        if (q_count >= 32u) drain(32u);
      }
    };
    // The atomic of a round is consumed TWO rounds later (box_rank_flat_kernel waits one): this kernel's own
    // gather keeps HBM busy, which stretches the L2 round trip the ranking chain waits for.
    FlatRound old, cur, nxt;
    int pending = 0;
    uint32_t ndesc = i0 + lane < i1 ? __ldcs(p.pdesc + i0 + lane) : 0u;
    for (uint64_t base = i0; base < i1; base += 32) {
      const uint32_t desc = ndesc;
      const bool have = base + lane < i1;
      if (base + 32 + lane < i1) ndesc = __ldcs(p.pdesc + base + 32 + lane);   // next batch (coalesced)
      const FlatBatch b = flat_batch(desc, have ? s_cnt[desc >> 24] : 0u, lane, p.cls_pad_log2);
      for (uint32_t r0 = 0; r0 < b.R; r0 += 32) {
        __syncwarp();   // order this round's atomics after the previous round's (same-key runs)
        front(b, base, r0, nxt);
        if (pending == 2) { back(old); --pending; }
        old = cur;
        cur = nxt;
        ++pending;
      }
    }
    if (pending == 2) back(old);
    if (pending >= 1) back(cur);
    if (q_count) drain(q_count);
  }
  __syncwarp();
  if (lane == 0) { __threadfence_block(); st_vol(ctrl + 2, 1u); }
}

// ------------------------------------------------------------------ host side: layouts
struct TablesLayout {
  TableSpec t;
  size_t off_cnt, off_packed, off_members, off_fm, off_size, off_off, off_cid, total;
};

static bool make_tables_layout(TablesLayout& L, int k, int n_dims, int bits) {
  if (n_dims < 1 || n_dims > MCCB200_BOX_MAX_DIMS || bits < 1 || n_dims * bits > BP_MAX_KEY_BITS) return false;
  if (k > BP_MAX_K || k < 2 || (k & (k - 1))) return false;
  L.t.k = (uint32_t)k;
  L.t.half_mask = (uint32_t)k / 2u - 1u;
  L.t.n_sets = 1u << n_dims;
  L.t.max_cls = (uint32_t)k < L.t.n_sets ? (uint32_t)k : L.t.n_sets;
  if (L.t.max_cls > 32) return false;   // a batch of 32 parents holds at most FLAT_TOFF_BIAS records
  L.t.cls_pad_log2 = 0;
  while ((1u << L.t.cls_pad_log2) < L.t.max_cls) ++L.t.cls_pad_log2;
  size_t cur = 0;
  auto take = [&](size_t bytes) { size_t o = cur; cur += align_up(bytes, 256); return o; };
  L.off_cnt = take((size_t)L.t.n_sets * 4);
  L.off_packed = take((size_t)L.t.n_sets * (1u << L.t.cls_pad_log2) * 4);
  L.off_members = take((size_t)L.t.n_sets * k * 2);
  L.off_fm = take((size_t)L.t.n_sets * L.t.max_cls * 4);     // scratch of the build
  L.off_size = take((size_t)L.t.n_sets * L.t.max_cls * 4);
  L.off_off = take((size_t)L.t.n_sets * L.t.max_cls * 4);
  L.off_cid = take((size_t)L.t.n_sets * L.t.n_sets * 2);
  L.total = cur;
  return true;
}

// pre is extended by min(in_per_part, BP_MAX_K) positions (a class run holds at most k <= in_per_part children)
struct SelLayout { size_t off_pre, off_fill, off_sp, total; uint32_t ext; };

static bool make_sel_layout(SelLayout& L, uint64_t in_per_part, uint64_t out_per_part) {
  if (in_per_part >= 0x40000000ull || out_per_part >= 0x40000000ull || in_per_part < 1 || out_per_part < 1) return false;
  L.ext = (uint32_t)(in_per_part < (uint64_t)BP_MAX_K ? in_per_part : (uint64_t)BP_MAX_K);
  size_t cur = 0;
  auto take = [&](size_t bytes) { size_t o = cur; cur += align_up(bytes, 256); return o; };
  L.off_pre = take(((size_t)in_per_part + L.ext + 2) * 4);
  L.off_fill = take(((size_t)in_per_part + 2) * 4);
  L.off_sp = take((size_t)out_per_part * 2 * 8);
  L.total = cur;
  return true;
}

static int max_optin_smem() {
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 0;
  return v;
}

// Launch geometry of box_rank_fused_kernel (one CTA per SM): ranking warps, transform warps, shared memory.
struct FusedGeom { int rank_warps, xform_warps; size_t smem, xform_words; };

static bool fused_geometry(uint32_t n_sets, uint32_t cls_pad_log2, uint32_t k, uint64_t in_per_part, uint64_t sel_ext,
                           uint64_t out_per_part, uint32_t row_bytes, FusedGeom& g) {
  size_t words = ((size_t)n_sets << cls_pad_log2) + (size_t)n_sets;
  words += words & 1;
  words += 4 * (size_t)out_per_part;
  words += (((size_t)in_per_part + sel_ext + 4) & ~(size_t)3) / 2;
  words += ((size_t)n_sets * k + 1) / 2;
  words = (words + 3) & ~(size_t)3;
  const size_t rank_words = (size_t)FQ_CAP * 4 + (size_t)FZ_JQ * 2 + 4 + 8;
  g.xform_words = (size_t)FZ_STAGES * 8 * row_bytes + 2 * FZ_STAGES + 2;   // stages (32 rows each) + mbarriers
  g.xform_words = (g.xform_words + 3) & ~(size_t)3;
  const size_t cap = (size_t)max_optin_smem();
  int rw = env_int("MCCB200_FUSED_WARPS", 24), tw = env_int("MCCB200_FUSED_TWARPS", 8);
  if (rw < 1) rw = 1;
  if (tw < 1) tw = 1;
  if (rw + tw > FZ_MAX_THREADS / 32) rw = FZ_MAX_THREADS / 32 - tw;
  while (tw > 1 && (words + (size_t)rw * rank_words + (size_t)tw * g.xform_words) * 4 > cap) --tw;
  if (rw > FZ_MAX_RINGS * tw) rw = FZ_MAX_RINGS * tw;
  if (rw < 1 || (words + (size_t)rw * rank_words + (size_t)tw * g.xform_words) * 4 > cap) return false;
  g.rank_warps = rw;
  g.xform_warps = tw;
  g.smem = (words + (size_t)rw * rank_words + (size_t)tw * g.xform_words) * 4;
  return true;
}

// What the fused kernel needs of the SHAPE (known to the counting and the ranking call alike, so that both cut
// the same segments); pointers and the selection size are checked again at launch (fused_plan).
static bool fused_shape_ok(int k, uint64_t in_per_part, int d, int ld, int drift_kind, int weighted) {
  if (env_int("MCCB200_BOX_FUSED", 1) == 0) return false;
  if (k > 256 || (in_per_part & (in_per_part - 1)) != 0) return false;
  if (weighted || (d & 3) || ld != d || d > 128 || drift_kind == 1) return false;
  const size_t pre_bytes_128 = ((size_t)((in_per_part + BP_MAX_K + 4) & ~3ull)) * 2 + (size_t)128 * 16;
  return pre_bytes_128 <= 28 * 1024;
}

struct PrkPlan {
  TablesLayout tl;
  uint32_t n_keys, n_segments;
  int warps_per_cta, ctas_per_sm, grid_b, pre_in_smem, mem_in_smem;
  uint64_t parents_per_seg;
  size_t smem_a, smem_b;
  size_t off_pdesc, off_keytot, off_table, off_idx, off_tables, off_sel, total;
};

// Workspace: [pdesc | table (+ key totals) | idx_out | class tables | selection tables]; the last
// two regions are only used by the all-in-one mccb200_box_prk_step.
// fused: 1 = segments for the fused rank + gather kernel (one per ranking warp), 0 = one per warp of the 4 resident
// CTAs of box_rank_flat_kernel, -1 = unknown (workspace query: sized for the larger of the two).
static bool make_prk_plan(PrkPlan& pl, uint64_t n, int k, int n_dims, int bits, uint64_t in_per_part,
                          uint64_t out_per_part_max, int fused = 0) {
  if (!make_tables_layout(pl.tl, k, n_dims, bits)) return false;
  if (in_per_part >= 0x40000000ull || in_per_part < (uint64_t)k) return false;  // a class run wraps at most once
  if (out_per_part_max >= 0x40000000ull || out_per_part_max < 1) return false;
  const uint64_t total_children = n * (uint64_t)k;
  if (total_children > 0x100000000ull) return false;
  const TableSpec& t = pl.tl.t;
  pl.n_keys = 1u << (n_dims * bits);
  const size_t row = (size_t)pl.n_keys * 4;
  const size_t tab_bytes = (size_t)t.n_sets * (1u << t.cls_pad_log2) * 4 + (size_t)t.n_sets * 4 + 8;
  const size_t mem_bytes = (size_t)t.n_sets * k * 2 + 8;
  const size_t pre_bytes = ((size_t)((in_per_part + BP_MAX_K + 4) & ~3ull)) * 2 + (size_t)out_per_part_max * 16;
  pl.pre_in_smem = (pre_bytes <= 28 * 1024 && out_per_part_max < 32768) ? 1 : 0;   // 2 * out_per_part fits uint16
  pl.mem_in_smem = mem_bytes <= 16 * 1024 ? 1 : 0;
  pl.warps_per_cta = 8;
  // 4 resident pass-B CTAs per SM.  Measured: 5 / 6 CTAs (48 / 40 registers) leave pass B at 0.75 / 0.83 ms
  // (it is issue-bound, not latency-bound) and push the counter table past L2 (scan 0.09 -> 0.20 / 0.28 ms).
  pl.ctas_per_sm = env_int("MCCB200_RANK_CTAS", 4);
  uint64_t segs = (uint64_t)sm_count() * pl.warps_per_cta * pl.ctas_per_sm;   // one warp per segment
  if (fused != 0 && pl.mem_in_smem) {
    FusedGeom g;
    if (fused_geometry(t.n_sets, t.cls_pad_log2, (uint32_t)k, in_per_part, BP_MAX_K, 128, 256, g)) {
      int spw = env_int("MCCB200_FUSED_SPW", 1);
      if (spw < 1) spw = 1;
      const uint64_t fsegs = (uint64_t)sm_count() * g.rank_warps * spw;
      segs = fused > 0 ? fsegs : (fsegs > segs ? fsegs : segs);
    }
  }
  if (segs > n) segs = n ? n : 1;
  pl.parents_per_seg = (n + segs - 1) / segs;
  pl.n_segments = (uint32_t)((n + pl.parents_per_seg - 1) / pl.parents_per_seg);
  pl.grid_b = (int)((pl.n_segments + pl.warps_per_cta - 1) / pl.warps_per_cta);
  pl.smem_a = row + tab_bytes;
  pl.smem_b = (size_t)pl.warps_per_cta * FQ_CAP * 16 + tab_bytes + (pl.pre_in_smem ? pre_bytes : 0) +
              (pl.mem_in_smem ? mem_bytes : 0);
  const uint64_t n_out_max = total_children / in_per_part * out_per_part_max;
  SelLayout sl;
  if (!make_sel_layout(sl, in_per_part, out_per_part_max)) return false;
  size_t cur = 0;
  auto take = [&](size_t bytes) { size_t o = cur; cur += align_up(bytes, 256); return o; };
  pl.off_pdesc = take((size_t)n * 4);
  pl.off_keytot = take(row);                             // key starts: before the table, so that their offset
  pl.off_table = take((size_t)pl.n_segments * row);      // does not depend on the segment count
  pl.off_idx = take((size_t)n_out_max * 4);
  pl.off_tables = take(pl.tl.total);
  pl.off_sel = take(sl.total);
  pl.total = cur;
  return true;
}

static void bind_tables(BoxPrkParams& p, const TablesLayout& L, const void* tables) {
  const char* w = (const char*)tables;
  p.n_sets = L.t.n_sets; p.max_cls = L.t.max_cls; p.cls_pad_log2 = L.t.cls_pad_log2;
  p.cls_count = (const uint32_t*)(w + L.off_cnt);
  p.cls_packed = (const uint32_t*)(w + L.off_packed);
  p.members = (const uint16_t*)(w + L.off_members);
}

static int launch_tables(const TablesLayout& L, const int32_t* dims_host, int n_dims, int bits, void* tables,
                         cudaStream_t st) {
  TableSpec t = L.t;
  int rc = make_box_spec(t.spec, dims_host, n_dims, bits, 1 << 30);
  if (rc) return rc;
  char* w = (char*)tables;
  box_tables_kernel<<<(t.n_sets + 63) / 64, 64, 0, st>>>(
      t, (uint32_t*)(w + L.off_cnt), (uint32_t*)(w + L.off_fm), (uint32_t*)(w + L.off_size), (uint32_t*)(w + L.off_off),
      (uint32_t*)(w + L.off_packed), (uint16_t*)(w + L.off_members), (uint16_t*)(w + L.off_cid));
  count_launches(1);
  return MCCB200_OK;
}

static int launch_selection(const SelLayout& L, const uint32_t* idx_local, uint64_t out_per_part, uint64_t in_per_part,
                            void* sel, cudaStream_t st) {
  char* w = (char*)sel;
  box_selection_kernel<<<1, 1024, 0, st>>>(idx_local, (uint32_t)out_per_part, (uint32_t)in_per_part, L.ext,
                                           (uint32_t*)(w + L.off_pre), (uint32_t*)(w + L.off_fill), (uint2*)(w + L.off_sp));
  count_launches(1);
  return MCCB200_OK;
}

// common argument checks + parameter block of pass A / pass B
static int setup_params(BoxPrkParams& p, PrkPlan& pl, const float* y0, uint64_t n, int d, const mccb200_drift* drift,
                        float dt, const mccb200_cubature* cub, const int32_t* dims_host, int n_dims, int bits,
                        uint64_t out_per_part, uint64_t in_per_part, void* workspace, size_t workspace_bytes,
                        const char* who) {
  int rc = make_box_spec(p.spec, dims_host, n_dims, bits, d);
  if (rc) return rc;
  rc = fill_expand_args(p.a, y0, n, d, 0, drift, dt, cub);
  if (rc) return rc;
  MCCB_REQUIRE(workspace, "%s: NULL workspace", who);
  if (cub->points != nullptr) return fail(MCCB200_ERR_UNSUPPORTED, "%s: only the Hadamard cubature is fused", who);
  const uint64_t total = n * (uint64_t)cub->k;
  MCCB_REQUIRE(in_per_part >= 1 && total % in_per_part == 0, "%s: in_per_part=%llu does not divide n*k=%llu", who,
               (unsigned long long)in_per_part, (unsigned long long)total);
  MCCB_REQUIRE(out_per_part >= 1, "%s: bad out_per_part", who);
  const int fused_hint = fused_shape_ok(cub->k, in_per_part, p.a.d, p.a.ld, p.a.drift_kind, p.a.weighted) ? 1 : 0;
  if (!make_prk_plan(pl, n, cub->k, n_dims, bits, in_per_part, out_per_part, fused_hint))
    return fail(MCCB200_ERR_UNSUPPORTED, "%s: shape unsupported (key bits %d > %d, k > %d or in_per_part < k)", who,
                n_dims * bits, BP_MAX_KEY_BITS, BP_MAX_K);
  if (workspace_bytes < pl.total)
    return fail(MCCB200_ERR_WORKSPACE_TOO_SMALL, "%s: workspace %zu < %zu", who, workspace_bytes, pl.total);
  char* w = (char*)workspace;
  p.pdesc = (uint32_t*)(w + pl.off_pdesc);
  p.table = (uint32_t*)(w + pl.off_table);
  p.key_start = (const uint32_t*)(w + pl.off_keytot);
  p.idx_out = (uint32_t*)(w + pl.off_idx);
  p.n_keys = pl.n_keys;
  p.n_segments = pl.n_segments;
  p.parents_per_seg = pl.parents_per_seg;
  p.in_per_part = (uint32_t)in_per_part;
  p.out_per_part = (uint32_t)out_per_part;
  p.in_shift = 0xFFFFFFFFu;
  if ((in_per_part & (in_per_part - 1)) == 0) { p.in_shift = 0; while ((1ull << p.in_shift) < in_per_part) ++p.in_shift; }
  p.warps_per_cta = pl.warps_per_cta;
  return MCCB200_OK;
}

static int launch_count(BoxPrkParams& p, const PrkPlan& pl, cudaStream_t st) {
  phase_mark("box_prk.count", st);
  if (box_dims_vec4(p.spec, p.a)) box_count_flat_kernel<true><<<pl.n_segments, 256, pl.smem_a, st>>>(p);
  else box_count_flat_kernel<false><<<pl.n_segments, 256, pl.smem_a, st>>>(p);
  phase_mark("box_prk.scan", st);
  uint32_t* key_tot = const_cast<uint32_t*>(p.key_start);
  if (pl.n_keys % 4 == 0)
    box_scan_columns_kernel<4><<<(pl.n_keys / 4 + SC_KQ - 1) / SC_KQ, SC_KQ * SC_CY, 0, st>>>(p.table, pl.n_segments, pl.n_keys, key_tot);
  else
    box_scan_columns_kernel<1><<<(pl.n_keys + SC_KQ - 1) / SC_KQ, SC_KQ * SC_CY, 0, st>>>(p.table, pl.n_segments, pl.n_keys, key_tot);
  box_scan_keys_kernel<<<1, 1024, 0, st>>>(key_tot, pl.n_keys);
  phase_mark("end", st);
  count_launches(3);
  return check_launch("box_prk_count");
}

static int launch_rank(BoxPrkParams& p, const PrkPlan& pl, cudaStream_t st) {
  phase_mark("box_prk.rank", st);
#define MCCB_FLAT2(PS, MS, P2)                                                                                          \
  do {                                                                                                                  \
    cudaFuncSetAttribute(box_rank_flat_kernel<PS, MS, P2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_b); \
    box_rank_flat_kernel<PS, MS, P2><<<pl.grid_b, pl.warps_per_cta * 32, pl.smem_b, st>>>(p);                          \
  } while (0)
#define MCCB_FLAT(PS, MS)                                      \
  do {                                                         \
    if (p.in_shift != 0xFFFFFFFFu) MCCB_FLAT2(PS, MS, true);   \
    else MCCB_FLAT2(PS, MS, false);                            \
  } while (0)
  if (pl.pre_in_smem && pl.mem_in_smem) MCCB_FLAT(true, true);
  else if (pl.pre_in_smem) MCCB_FLAT(true, false);
  else if (pl.mem_in_smem) MCCB_FLAT(false, true);
  else MCCB_FLAT(false, false);
#undef MCCB_FLAT
#undef MCCB_FLAT2
  count_launches(1);
  return check_launch("box_prk_rank");
}

// Shape / alignment conditions of box_rank_fused_kernel; fills the launch geometry.
struct FusedLaunch { FusedOut o; int warps, grid; size_t smem; };

static bool fused_plan(const BoxPrkParams& p, const PrkPlan& pl, float* y1, float* keycols_out, FusedLaunch& fl) {
  const ExpandArgs& a = p.a;
  if (!fused_shape_ok((int)a.k, p.in_per_part, a.d, a.ld, a.drift_kind, a.weighted)) return false;
  if (!pl.pre_in_smem || !pl.mem_in_smem || p.in_shift == 0xFFFFFFFFu) return false;
  if ((((uintptr_t)a.y0 | (uintptr_t)y1) & 15) != 0) return false;
  if (a.drift_kind == 2 && (((uintptr_t)a.mean | (uintptr_t)a.inv_var) & 15) != 0) return false;
  if (keycols_out && (((uintptr_t)keycols_out & 15) != 0)) return false;
  FusedGeom g;
  if (!fused_geometry(p.n_sets, p.cls_pad_log2, a.k, p.in_per_part, p.sel_ext, p.out_per_part, (uint32_t)a.d * 4u, g))
    return false;
  if ((uint32_t)g.rank_warps > p.n_segments) g.rank_warps = (int)p.n_segments;
  fl.warps = g.rank_warps + g.xform_warps;
  fl.smem = g.smem;
  fl.grid = (int)min((uint64_t)sm_count(), ((uint64_t)p.n_segments + g.rank_warps - 1) / g.rank_warps);
  fl.o.y1 = y1;
  fl.o.keycols_out = keycols_out;
  fl.o.keycol0 = p.a.keycol0;
  fl.o.vpr = a.d / 4;
  fl.o.lpr_shift = 0;
  while ((1 << fl.o.lpr_shift) < fl.o.vpr) ++fl.o.lpr_shift;
  fl.o.rank_warps = g.rank_warps;
  fl.o.row_bytes = (uint32_t)a.d * 4u;
  fl.o.xform_words = (uint32_t)g.xform_words;
  fl.o.spin_ns = (uint32_t)env_int("MCCB200_FUSED_SPIN_NS", 200);
  fl.o.one = 1.0f; fl.o.neg_one = -1.0f; fl.o.neg_zero = -0.0f;
  return true;
}

template <int DRIFT, bool D64>
static void launch_fused_k(const BoxPrkParams& p, const FusedLaunch& fl, cudaStream_t st) {
  cudaFuncSetAttribute(box_rank_fused_kernel<DRIFT, D64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fl.smem);
  box_rank_fused_kernel<DRIFT, D64><<<fl.grid, fl.warps * 32, fl.smem, st>>>(p, fl.o);
}

template <int DRIFT>
static void launch_fused_d(const BoxPrkParams& p, const FusedLaunch& fl, cudaStream_t st) {
  if (p.a.d == 64) launch_fused_k<DRIFT, true>(p, fl, st);
  else launch_fused_k<DRIFT, false>(p, fl, st);
}

static int launch_rank_gather(BoxPrkParams& p, const PrkPlan& pl, float* y1, cudaStream_t st) {
  FusedLaunch fl;
  if (fused_plan(p, pl, y1, p.a.keycols_out, fl)) {
    phase_mark("box_prk.rank_gather", st);
    if (p.a.drift_kind == 2) launch_fused_d<2>(p, fl, st);
    else launch_fused_d<0>(p, fl, st);
    phase_mark("end", st);
    count_launches(1);
    return check_launch("box_prk_rank_gather(fused)");
  }
  int rc0 = launch_rank(p, pl, st);
  if (rc0) return rc0;
  // gather: the HBM-bound fused expand+select materialises the selected children
  ExpandArgs g = p.a;
  g.keycols = nullptr;
  const uint64_t total = p.a.n * (uint64_t)p.a.k;
  g.idx = p.idx_out; g.perm = nullptr; g.out_per_part = 1; g.in_per_part = 1;
  g.n_out = total / p.in_per_part * p.out_per_part; g.y1 = y1;
  int rc = run_expand_select(g, st);               // (counts its own launch, launch_rank counted the ranking kernel)
  if (rc) return rc;
  return check_launch("box_prk_rank_gather");
}

}  // namespace mccb

using namespace mccb;

extern "C" size_t mccb200_box_prk_tables_bytes(int k, int n_dims, int bits) {
  TablesLayout L;
  return make_tables_layout(L, k, n_dims, bits) ? L.total : 0;
}

extern "C" int mccb200_box_prk_tables(int k, const int32_t* dims_host, int n_dims, int bits, void* tables,
                                      size_t tables_bytes, void* stream) {
  TablesLayout L;
  if (!make_tables_layout(L, k, n_dims, bits))
    return fail(MCCB200_ERR_UNSUPPORTED, "box_prk_tables: unsupported (k=%d, n_dims=%d, bits=%d)", k, n_dims, bits);
  MCCB_REQUIRE(tables && tables_bytes >= L.total, "box_prk_tables: buffer %zu < %zu", tables_bytes, L.total);
  int rc = launch_tables(L, dims_host, n_dims, bits, tables, (cudaStream_t)stream);
  if (rc) return rc;
  return check_launch("box_prk_tables");
}

extern "C" size_t mccb200_box_prk_selection_bytes(uint64_t in_per_part, uint64_t out_per_part) {
  SelLayout L;
  return make_sel_layout(L, in_per_part, out_per_part) ? L.total : 0;
}

extern "C" int mccb200_box_prk_selection(const uint32_t* idx_local, uint64_t out_per_part, uint64_t in_per_part,
                                         void* sel, size_t sel_bytes, void* stream) {
  SelLayout L;
  if (!make_sel_layout(L, in_per_part, out_per_part))
    return fail(MCCB200_ERR_UNSUPPORTED, "box_prk_selection: partition sizes out of range");
  MCCB_REQUIRE(idx_local && sel && sel_bytes >= L.total, "box_prk_selection: buffer %zu < %zu", sel_bytes, L.total);
  int rc = launch_selection(L, idx_local, out_per_part, in_per_part, sel, (cudaStream_t)stream);
  if (rc) return rc;
  return check_launch("box_prk_selection");
}

extern "C" size_t mccb200_box_prk_step_workspace_bytes(uint64_t n, int k, int n_dims, int bits,
                                                       uint64_t in_per_part) {
  PrkPlan pl;
  // 0 = unsupported shape.  Sized for out_per_part <= in_per_part.
  if (!make_prk_plan(pl, n, k, n_dims, bits, in_per_part, in_per_part, -1)) return 0;
  return pl.total;
}

extern "C" int mccb200_box_prk_count(const float* y0, uint64_t n, int d, const mccb200_drift* drift, float dt,
                                     const mccb200_cubature* cub, const int32_t* dims_host, int n_dims, int bits,
                                     const float* bounds, uint64_t in_per_part, const void* tables,
                                     const float* keycols, void* workspace, size_t workspace_bytes, void* stream) {
  BoxPrkParams p{};
  PrkPlan pl;
  int rc = setup_params(p, pl, y0, n, d, drift, dt, cub, dims_host, n_dims, bits, 1, in_per_part, workspace,
                        workspace_bytes, "box_prk_count");
  if (rc) return rc;
  MCCB_REQUIRE(bounds && tables, "box_prk_count: NULL argument");
  if (n == 0) return MCCB200_OK;
  p.bounds = bounds;
  bind_tables(p, pl.tl, tables);
  p.a.keycols = (box_dims_vec4(p.spec, p.a) && keycols && (((uintptr_t)keycols & 15) == 0)) ? keycols : nullptr;
  return launch_count(p, pl, (cudaStream_t)stream);
}

extern "C" int mccb200_box_prk_rank_gather(const float* y0, uint64_t n, int d, const mccb200_drift* drift, float dt,
                                           const mccb200_cubature* cub, const int32_t* dims_host, int n_dims, int bits,
                                           uint64_t out_per_part, uint64_t in_per_part, const void* tables,
                                           const void* sel, float* y1, float* keycols_out, void* workspace,
                                           size_t workspace_bytes, void* stream) {
  BoxPrkParams p{};
  PrkPlan pl;
  int rc = setup_params(p, pl, y0, n, d, drift, dt, cub, dims_host, n_dims, bits, out_per_part, in_per_part, workspace,
                        workspace_bytes, "box_prk_rank_gather");
  if (rc) return rc;
  MCCB_REQUIRE(tables && sel && y1, "box_prk_rank_gather: NULL argument");
  if (n == 0) return MCCB200_OK;
  SelLayout sl;
  make_sel_layout(sl, in_per_part, out_per_part);
  bind_tables(p, pl.tl, tables);
  p.pre = (const uint32_t*)((const char*)sel + sl.off_pre);
  p.sp = (const uint2*)((const char*)sel + sl.off_sp);
  p.sel_ext = sl.ext;
  if (keycols_out && box_dims_vec4(p.spec, p.a) && (((uintptr_t)keycols_out | (uintptr_t)y1) & 15) == 0) {
    p.a.keycols_out = keycols_out;
    p.a.keycol0 = p.spec.dims[0];
  } else if (keycols_out) {
    return fail(MCCB200_ERR_UNSUPPORTED, "box_prk_rank_gather: keycols_out needs 4 consecutive 16-byte aligned key columns");
  }
  return launch_rank_gather(p, pl, y1, (cudaStream_t)stream);
}

/* Pass B alone, for particles sharded over several devices with ONE global partition (the reference's):
 * `key_start[key]` is the GLOBAL sorted position of this shard's first child with that key (start of the key
 * over all shards + children of lower-ranked shards with the same key), so ranks, partitions and output slots
 * are global; every selected child's LOCAL id (parent * k + point) is written to `idx_slots[global slot]`,
 * which the caller pre-fills with 0xFFFFFFFF and sizes for all partitions.  No gather. */
extern "C" int mccb200_box_prk_rank_global(const float* y0, uint64_t n, int d, const mccb200_drift* drift, float dt,
                                           const mccb200_cubature* cub, const int32_t* dims_host, int n_dims, int bits,
                                           uint64_t out_per_part, uint64_t in_per_part, const void* tables,
                                           const void* sel, const uint32_t* key_start, uint32_t* idx_slots,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  BoxPrkParams p{};
  PrkPlan pl;
  int rc = setup_params(p, pl, y0, n, d, drift, dt, cub, dims_host, n_dims, bits, out_per_part, in_per_part, workspace,
                        workspace_bytes, "box_prk_rank_global");
  if (rc) return rc;
  MCCB_REQUIRE(tables && sel && key_start && idx_slots, "box_prk_rank_global: NULL argument");
  if (n == 0) return MCCB200_OK;
  SelLayout sl;
  make_sel_layout(sl, in_per_part, out_per_part);
  bind_tables(p, pl.tl, tables);
  p.pre = (const uint32_t*)((const char*)sel + sl.off_pre);
  p.sp = (const uint2*)((const char*)sel + sl.off_sp);
  p.sel_ext = sl.ext;
  p.key_start = key_start;
  p.idx_out = idx_slots;
  return launch_rank(p, pl, (cudaStream_t)stream);
}

/* Byte offset, inside the workspace `mccb200_box_prk_count` filled, of the uint32 array with the LOCAL sorted
 * position of the first child of every key (n_keys entries, ascending; the last key ends at n * k). */
extern "C" size_t mccb200_box_prk_key_start_offset(uint64_t n, int k, int n_dims, int bits, uint64_t in_per_part,
                                                   uint64_t out_per_part, uint32_t* n_keys) {
  PrkPlan pl;
  if (!make_prk_plan(pl, n, (uint32_t)k, n_dims, bits, in_per_part, out_per_part)) return (size_t)-1;
  if (n_keys) *n_keys = pl.n_keys;
  return pl.off_keytot;
}

extern "C" int mccb200_box_prk_step(const float* y0, uint64_t n, int d, const mccb200_drift* drift, float dt,
                                    const mccb200_cubature* cub, const int32_t* dims_host, int n_dims, int bits,
                                    const float* bounds, const uint32_t* idx_local, uint64_t out_per_part,
                                    uint64_t in_per_part, float* y1, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BoxPrkParams p{};
  PrkPlan pl;
  int rc = setup_params(p, pl, y0, n, d, drift, dt, cub, dims_host, n_dims, bits, out_per_part, in_per_part, workspace,
                        workspace_bytes, "box_prk_step");
  if (rc) return rc;
  MCCB_REQUIRE(bounds && idx_local && y1, "box_prk_step: NULL argument");
  if (n == 0) return MCCB200_OK;
  char* w = (char*)workspace;
  SelLayout sl;
  make_sel_layout(sl, in_per_part, out_per_part);
  phase_mark("box_prk.tables", st);
  rc = launch_tables(pl.tl, dims_host, n_dims, bits, w + pl.off_tables, st);
  if (rc) return rc;
  launch_selection(sl, idx_local, out_per_part, in_per_part, w + pl.off_sel, st);
  p.bounds = bounds;
  bind_tables(p, pl.tl, w + pl.off_tables);
  p.pre = (const uint32_t*)(w + pl.off_sel + sl.off_pre);
  p.sp = (const uint2*)(w + pl.off_sel + sl.off_sp);
  p.sel_ext = sl.ext;
  rc = launch_count(p, pl, st);
  if (rc) return rc;
  return launch_rank_gather(p, pl, y1, st);
}
