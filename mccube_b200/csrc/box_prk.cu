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
// only on (S, cubature, key dims): tables built once per call (box_tables_kernel).
//
//   pass A  box_count_kernel : per-segment histogram of child keys: one lane per parent,
//           one shared-memory atomicAdd(hist[key_class], class_size) per CLASS.
//   scan    box_scan_*       : start[seg][key] = #children with a smaller key, or the same
//           key in an earlier segment  (counting sort == "integer key-sort + scan").
//   pass B  box_rank_kernel  : one warp per segment walks its parents IN ORDER with a private
//           shared-memory counter row; lane c owns class c of the current parent:
//           base = counter[key]; counter[key] += size  gives the class its rank run
//           [base, base+size).  A prefix count `pre` over the (sorted) selected positions tells
//           in two loads whether the run contains a selected position; only then are the hit
//           members resolved (members[S][class][t]) and their child ids i*k + j written to
//           idx_out[partition * out_per_part + l]   (4 bytes per selected child).
//   gather  expand_kernel    : the HBM-bound fused expand+select of expand_select.cu
//           materialises exactly those n children: read n rows + write n rows.
#include <stdlib.h>

#include "common.cuh"
#include "expand.cuh"

namespace mccb {

constexpr int BP_MAX_KEY_BITS = 12;
constexpr int BP_MAX_K = 1024;
constexpr int BR_BATCH = 32;  // parents whose key dims one warp loads at once (one per lane)
constexpr int BR_GROUP = 16;  // parents per serial/parallel phase pair (register arrays)
constexpr uint32_t BP_INVALID = 0xFFFFFFFFu;

int run_expand_select(ExpandArgs& a, cudaStream_t st);

struct BoxPrkParams {
  ExpandArgs a;
  BoxSpec spec;
  const float* bounds;
  uint32_t n_keys;          // 2^(n_dims*bits)
  uint32_t n_segments;
  uint64_t parents_per_seg;
  uint32_t in_per_part, out_per_part;
  uint32_t in_shift;        // log2(in_per_part) or 0xFFFFFFFF
  uint32_t n_sets;          // 2^n_dims split sets
  uint32_t max_cls;         // classes per split set (table stride) = min(k, 2^n_dims)
  // class tables, all indexed [S][...]
  const uint32_t* cls_count;   // [n_sets]
  const uint32_t* cls_fm;      // [n_sets][max_cls]   field mask of the class
  const uint32_t* cls_size;    // [n_sets][max_cls]   members
  const uint32_t* cls_off;     // [n_sets][max_cls]   offset of its member list
  const uint32_t* cls_packed;  // [n_sets][cls_pad]   fm | (size-1) << 12 | off << 22, or BP_INVALID
  const uint16_t* members;     // [n_sets][k]         points j sorted by (class, j)
  uint32_t cls_pad_log2;       // log2 of the padded class count (lanes per parent slot in pass B)
  // selection tables
  const uint32_t* pre;         // [in_per_part + 1]  #selected entries with position < pos
  const uint2* sp;             // [out_per_part]     (position, l) sorted by position
  uint32_t* pdesc;             // [n] per-parent K0 | D << 12 | S << 24 (written by pass A)
  uint32_t* table;             // [n_segments + 1][n_keys]
  const uint32_t* key_start;   // [n_keys] children with a smaller key (flat form: rows stay relative)
  uint32_t* idx_out;           // [n_out] selected child ids
  int warps_per_cta;
  int pre_in_smem, tab_in_smem;
  int debug;  // MCCB200_DEBUG_RANK bit mask (timing experiments only; results are wrong when set)
};

__device__ __forceinline__ uint32_t box_cell_u(float x, float lo, float inv_h, int bits) {
  float v = __fmul_rn(__fsub_rn(x, lo), inv_h);
  float c = fminf(fmaxf(floorf(v), 0.0f), (float)((1 << bits) - 1));
  return (uint32_t)c;
}

// sign pattern of cubature point j over the key dims: bit q set <=> + sign in key dim q
__device__ __forceinline__ uint32_t sign_pattern(const BoxPrkParams& p, uint32_t j) {
  uint32_t rho = 0;
  for (int q = 0; q < p.spec.n_dims; ++q)
    if (!hadamard_neg(j, p.a.half_mask, (uint32_t)p.spec.dims[q])) rho |= 1u << q;
  return rho;
}

__device__ __forceinline__ uint32_t field_mask_of_pattern(const BoxPrkParams& p, uint32_t rho) {
  uint32_t fm = 0;
  const uint32_t field = (1u << p.spec.bits) - 1u;
  for (int q = 0; q < p.spec.n_dims; ++q)
    if ((rho >> q) & 1u) fm |= field << (p.spec.bits * (p.spec.n_dims - 1 - q));
  return fm;
}

// (K0, D, S) of parent i from its key dims only
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
__global__ void box_tables_kernel(const BoxPrkParams p, uint32_t* cls_count, uint32_t* cls_fm, uint32_t* cls_size,
                                  uint32_t* cls_off, uint32_t* cls_packed, uint16_t* members,
                                  uint16_t* scratch_cid /*[n_sets][n_sets]*/) {
  const uint32_t S = blockIdx.x * blockDim.x + threadIdx.x;
  if (S >= p.n_sets) return;
  const uint32_t k = p.a.k;
  uint16_t* cid = scratch_cid + (size_t)S * p.n_sets;
  uint32_t* fm = cls_fm + (size_t)S * p.max_cls;
  uint32_t* size = cls_size + (size_t)S * p.max_cls;
  uint32_t* off = cls_off + (size_t)S * p.max_cls;
  uint16_t* mem = members + (size_t)S * k;
  // members per restricted sign pattern sig = rho & S
  for (uint32_t s = 0; s < p.n_sets; ++s) cid[s] = 0u;
  for (uint32_t j = 0; j < k; ++j) cid[sign_pattern(p, j) & S] += 1;
  // class ids in order of increasing key offset D & FM(sig): key dim 0 owns the most significant
  // field, so that is the order of sig with its n_dims bits reversed (pass B's sweep relies on it)
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

// pre[pos] = #l with idx_local[l] < pos (pos in [0, in_per_part]); sp = (pos, l) sorted by pos.
// Single CTA (the selection vector is tiny next to the state).
__global__ void __launch_bounds__(1024)
box_selection_kernel(const uint32_t* __restrict__ idx_local, uint32_t out_per_part, uint32_t in_per_part,
                     uint32_t* __restrict__ pre, uint32_t* __restrict__ fill, uint2* __restrict__ sp) {
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
      uint32_t t = wtot[threadIdx.x], s = t;
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
        if (threadIdx.x >= o) s += y;
      }
      wtot[threadIdx.x] = s - t;
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
  }
}

// ------------------------------------------------------------------ pass A
__global__ void __launch_bounds__(256)
box_count_kernel(const BoxPrkParams p) {
  extern __shared__ uint32_t hist[];
  const ExpandArgs& a = p.a;
  const uint32_t seg = blockIdx.x;
  for (uint32_t i = threadIdx.x; i < p.n_keys; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  const uint64_t i0 = (uint64_t)seg * p.parents_per_seg;
  const uint64_t i1 = min(a.n, i0 + p.parents_per_seg);
  for (uint64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    uint32_t K0, D, S;
    describe_parent(p, i, K0, D, S);
    p.pdesc[i] = K0 | (D << 12) | (S << 24);       // pass B re-reads 4 coalesced bytes per parent
    const uint32_t ncls = __ldg(p.cls_count + S);
    const uint32_t* fm = p.cls_fm + (size_t)S * p.max_cls;
    const uint32_t* sz = p.cls_size + (size_t)S * p.max_cls;
    for (uint32_t c = 0; c < ncls; ++c) atomicAdd(&hist[K0 + (D & __ldg(fm + c))], __ldg(sz + c));
  }
  __syncthreads();
  uint32_t* row = p.table + (uint64_t)seg * p.n_keys;
  for (uint32_t i = threadIdx.x; i < p.n_keys; i += blockDim.x) row[i] = hist[i];
}

// ------------------------------------------------------------------ scan
// table[seg][key] -> exclusive prefix over segments (per key); key_tot[key] = column total.
// CTA = 32 keys x 32 segment-chunks; every table access is a coalesced 128-byte row.
__global__ void __launch_bounds__(1024)
box_scan_columns_kernel(uint32_t* __restrict__ table, uint32_t n_segments, uint32_t n_keys,
                        uint32_t* __restrict__ key_tot) {
  __shared__ uint32_t part[32][33];
  const uint32_t kx = threadIdx.x & 31, cy = threadIdx.x >> 5;
  const uint32_t key = blockIdx.x * 32 + kx;
  const uint32_t chunk = (n_segments + 31) / 32;
  const uint32_t s0 = min(n_segments, cy * chunk), s1 = min(n_segments, s0 + chunk);
  uint32_t sum = 0;
  if (key < n_keys)
    for (uint32_t s = s0; s < s1; ++s) sum += table[(uint64_t)s * n_keys + key];
  part[cy][kx] = sum;
  __syncthreads();
  uint32_t run = 0;
  for (uint32_t c = 0; c < cy; ++c) run += part[c][kx];
  if (key < n_keys) {
    for (uint32_t s = s0; s < s1; ++s) {
      const uint32_t c = table[(uint64_t)s * n_keys + key];
      table[(uint64_t)s * n_keys + key] = run;
      run += c;
    }
    if (cy == 31) key_tot[key] = run;
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
      uint32_t t = wtot[threadIdx.x], s = t;
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
        if (threadIdx.x >= o) s += y;
      }
      wtot[threadIdx.x] = s - t;
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

// table[seg][key] += key_start[key]: every row becomes the absolute start ranks of its segment
__global__ void __launch_bounds__(256)
box_scan_fold_kernel(uint32_t* __restrict__ table, uint32_t n_segments, uint32_t n_keys,
                     const uint32_t* __restrict__ key_start) {
  const uint64_t total = (uint64_t)n_segments * n_keys;
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x)
    table[e] += key_start[e & (n_keys - 1u)];
}

// ------------------------------------------------------------------ pass B (sweep form)
// One warp per segment, LANES = PARENTS (32 consecutive parents per batch).  Every lane walks
// the classes of its own parent in increasing key order; each step the warp agrees on the
// smallest pending key (redux.min), the lanes holding it take consecutive rank runs in lane
// (= parent) order via a warp scan of their class sizes, and one lane advances the segment's
// counter for that key.  Counter rows live in GLOBAL memory (the scanned table itself, L2
// resident): shared memory no longer caps occupancy, and a batch costs one L2 round trip per
// DISTINCT key -- a handful once the state is box-ordered, which every step's output is.
template <bool PRE_SMEM, bool MEM_SMEM>
__global__ void __launch_bounds__(256, 4)
box_rank_sweep_kernel(const BoxPrkParams p) {
  extern __shared__ uint32_t smem[];
  const ExpandArgs& a = p.a;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t* cursor = smem;
  const uint32_t cls_pad = 1u << p.cls_pad_log2;
  uint32_t* s_tab = cursor; cursor += p.n_sets * cls_pad;
  for (uint32_t i = threadIdx.x; i < p.n_sets * cls_pad; i += blockDim.x) s_tab[i] = p.cls_packed[i];
  const uint32_t* pre = p.pre;
  const uint2* sp = p.sp;
  if (PRE_SMEM) {
    uint32_t* s_pre = cursor;
    cursor += (p.in_per_part + 2u) & ~1u;
    uint2* s_sp = reinterpret_cast<uint2*>(cursor);
    cursor += 2u * p.out_per_part;
    for (uint32_t i = threadIdx.x; i <= p.in_per_part; i += blockDim.x) s_pre[i] = p.pre[i];
    for (uint32_t i = threadIdx.x; i < p.out_per_part; i += blockDim.x) s_sp[i] = p.sp[i];
    pre = s_pre; sp = s_sp;
  }
  const uint16_t* members = p.members;
  if (MEM_SMEM) {
    uint16_t* s_mem = reinterpret_cast<uint16_t*>(cursor);
    for (uint32_t i = threadIdx.x; i < p.n_sets * a.k; i += blockDim.x) s_mem[i] = p.members[i];
    members = s_mem;
  }
  __syncthreads();
  const uint32_t seg = blockIdx.x * p.warps_per_cta + warp;
  if (seg >= p.n_segments) return;
  uint32_t* row = p.table + (uint64_t)seg * p.n_keys;   // absolute start ranks of this segment

  const uint64_t i0 = (uint64_t)seg * p.parents_per_seg;
  const uint64_t i1 = min(a.n, i0 + p.parents_per_seg);
  const uint32_t Pm = p.in_per_part, nm = p.out_per_part;

  uint32_t nK0 = 0, nD = 0, nS = 0;
  if (i0 + lane < i1) describe_parent(p, i0 + lane, nK0, nD, nS);
  for (uint64_t base = i0; base < i1; base += 32) {
    const uint32_t K0 = nK0, D = nD, S = nS;
    const bool have = base + lane < i1;
    if (base + 32 + lane < i1) describe_parent(p, base + 32 + lane, nK0, nD, nS);  // next batch
    const uint64_t child0 = (base + lane) * (uint64_t)a.k;
    const uint32_t* trow = s_tab + (S << p.cls_pad_log2);
    uint32_t c = 0;
    uint32_t ent = have ? trow[0] : BP_INVALID;
    uint32_t curkey = ent == BP_INVALID ? 0xFFFFFFFFu : K0 + (D & (ent & 0xFFFu));
    // Every class of one parent has the same size (the classes are the fibres of an affine map
    // over GF(2)): size = k >> lsz, a per-lane constant.  Ballot masks per log-size turn the
    // "sizes of earlier lanes holding the same key" prefix into a few popcounts.
    const uint32_t size = ent == BP_INVALID ? 0u : ((ent >> 12) & 0x3FFu) + 1u;
    const uint32_t lsz = size ? (uint32_t)(__ffs(a.k) - __ffs(size)) : 31u;
    uint32_t szmask[6];
#pragma unroll
    for (int b = 0; b < 6; ++b) szmask[b] = __ballot_sync(0xffffffffu, lsz == (uint32_t)b);

    uint32_t kmin = __reduce_min_sync(0xffffffffu, curkey);
    uint32_t start_next = kmin != 0xFFFFFFFFu ? __ldcg(row + kmin) : 0u;
    while (kmin != 0xFFFFFFFFu) {
      const bool act = curkey == kmin;
      const uint32_t m = __ballot_sync(0xffffffffu, act);
      uint32_t excl = 0, total = 0;
#pragma unroll
      for (int b = 0; b < 6; ++b) {
        const uint32_t mb = m & szmask[b];
        excl += (uint32_t)__popc(mb & lt_mask) * (a.k >> b);
        total += (uint32_t)__popc(mb) * (a.k >> b);
      }
      const uint32_t ent_cur = ent;
      if (act) {                                   // advance first: the next key only needs tables
        ++c;
        ent = c < cls_pad ? trow[c] : BP_INVALID;
        curkey = ent == BP_INVALID ? 0xFFFFFFFFu : K0 + (D & (ent & 0xFFFu));
      }
      const uint32_t knext = __reduce_min_sync(0xffffffffu, curkey);
      const uint32_t start = start_next;
      if (lane == 0) __stcg(row + kmin, start + total);
      if (knext != 0xFFFFFFFFu) start_next = __ldcg(row + knext);  // L2 round trip overlaps the tests below
      if (act) {
        const uint32_t rank0 = start + excl;
        uint32_t part0, pos0;
        if (p.in_shift != 0xFFFFFFFFu) { part0 = rank0 >> p.in_shift; pos0 = rank0 & (Pm - 1u); }
        else { part0 = rank0 / Pm; pos0 = rank0 - part0 * Pm; }
        const uint32_t end = pos0 + size;          // size <= k <= Pm: wraps at most once
        const uint32_t first = pre[pos0];
        const uint32_t n_hit = end <= Pm ? pre[end] - first : (nm - first) + pre[end - Pm];
        if (n_hit) {
          const uint16_t* mem = members + (size_t)S * a.k + (ent_cur >> 22);
          for (uint32_t h = 0; h < n_hit; ++h) {
            uint32_t e = first + h;
            const bool wrap = e >= nm;
            if (wrap) e -= nm;
            const uint2 pl = sp[e];                // (position, output slot l)
            const uint32_t tt = wrap ? pl.x + Pm - pos0 : pl.x - pos0;
            p.idx_out[(uint64_t)(part0 + (wrap ? 1u : 0u)) * nm + pl.y] = (uint32_t)(child0 + mem[tt]);
          }
        }
      }
      kmin = knext;
    }
    __syncwarp();  // lane 0's counter stores are ordered before the next batch's loads
  }
}

// ------------------------------------------------------------------ pass B (hash form)
// One warp per segment, LANES = PARENTS, counter rows in global memory like the sweep form,
// but every phase is record-parallel (lane loops over the classes of its own parent):
//   1. insert : every (lane, class) record inserts its key into a per-warp shared-memory hash
//               table and ORs its lane bit into the key's mask.  A parent contributes at most
//               one record per key, so mask == the set of parents of this batch holding that key.
//   2. leaders: the lowest lane of each mask owns the key: start = row[key]; row[key] += total
//               size of the mask (one L2 read-modify-write per DISTINCT key); start -> table.
//   3. ranks  : rank0 = start + sum of the class sizes of the lower lanes in the mask (popcounts:
//               the class size is a per-lane power of two), then the selection test and hits.
// Result is independent of hash placement.  The table has HT_SLOTS slots; a batch whose records
// do not fit is re-run as sub-batches of fewer parents (earlier parents first).
constexpr int HT_SLOTS = 128;

constexpr int HQ_CAP = 64;   // deferred-hit queue entries per warp

template <bool PRE_SMEM, bool MEM_SMEM>
__global__ void __launch_bounds__(256, 4)
box_rank_hash_kernel(const BoxPrkParams p) {
  extern __shared__ uint32_t smem[];
  const ExpandArgs& a = p.a;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t* cursor = smem;
  uint32_t* tkey = cursor + (size_t)warp * 3 * HT_SLOTS;
  uint32_t* tmask = tkey + HT_SLOTS;
  uint32_t* tstart = tmask + HT_SLOTS;
  cursor += (size_t)p.warps_per_cta * 3 * HT_SLOTS;
  uint4* hq = reinterpret_cast<uint4*>(cursor) + (size_t)warp * HQ_CAP;   // (rank0, owner|c<<5, first, n_hit)
  cursor += (size_t)p.warps_per_cta * HQ_CAP * 4;
  const uint32_t cls_pad = 1u << p.cls_pad_log2;
  uint32_t* s_tab = cursor; cursor += p.n_sets * cls_pad;
  uint32_t* s_cnt = cursor; cursor += p.n_sets;
  for (uint32_t i = threadIdx.x; i < p.n_sets * cls_pad; i += blockDim.x) s_tab[i] = p.cls_packed[i];
  for (uint32_t i = threadIdx.x; i < p.n_sets; i += blockDim.x) s_cnt[i] = p.cls_count[i];
  // shared copy of the selection prefix as uint16 (out_per_part < 65536 when PRE_SMEM)
  const uint2* sp = p.sp;
  uint16_t* s_pre16 = nullptr;
  if (PRE_SMEM) {
    uint2* s_sp = reinterpret_cast<uint2*>(cursor);
    cursor += 2u * p.out_per_part;
    s_pre16 = reinterpret_cast<uint16_t*>(cursor);
    cursor += ((p.in_per_part + 4u) & ~3u) / 2u;
    for (uint32_t i = threadIdx.x; i <= p.in_per_part; i += blockDim.x) s_pre16[i] = (uint16_t)p.pre[i];
    for (uint32_t i = threadIdx.x; i < p.out_per_part; i += blockDim.x) s_sp[i] = p.sp[i];
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
  uint32_t* row = p.table + (uint64_t)seg * p.n_keys;   // absolute start ranks of this segment

  const uint64_t i0 = (uint64_t)seg * p.parents_per_seg;
  const uint64_t i1 = min(a.n, i0 + p.parents_per_seg);
  const uint32_t Pm = p.in_per_part, nm = p.out_per_part;
  const int logk = __ffs(a.k) - 1;
  // sub-batch count that always fits: (32 / n) parents * max_cls records <= HT_SLOTS
  int nsub_safe = 1;
  while ((32 / nsub_safe) * (int)p.max_cls > HT_SLOTS) nsub_safe <<= 1;

  // resolve one deferred hit record (any lane may process any parent's record)
  auto emit_hits = [&](uint32_t rank0, uint32_t S_o, uint32_t ent, uint32_t first, uint32_t n_hit, uint64_t child0) {
    uint32_t part0, pos0;
    if (p.in_shift != 0xFFFFFFFFu) { part0 = rank0 >> p.in_shift; pos0 = rank0 & (Pm - 1u); }
    else { part0 = rank0 / Pm; pos0 = rank0 - part0 * Pm; }
    const uint16_t* mem = members + (size_t)S_o * a.k + (ent >> 22);
    for (uint32_t hh = 0; hh < n_hit; ++hh) {
      uint32_t e = first + hh;
      const bool wrap = e >= nm;
      if (wrap) e -= nm;
      const uint2 pl = sp[e];                      // (position, output slot l)
      const uint32_t tt = wrap ? pl.x + Pm - pos0 : pl.x - pos0;
      p.idx_out[(uint64_t)(part0 + (wrap ? 1u : 0u)) * nm + pl.y] = (uint32_t)(child0 + mem[tt]);
    }
  };

  uint32_t ndesc = i0 + lane < i1 ? __ldcs(p.pdesc + i0 + lane) : 0u;
  for (uint64_t base = i0; base < i1; base += 32) {
    const uint32_t desc = ndesc;
    const bool have = base + lane < i1;
    if (base + 32 + lane < i1) ndesc = __ldcs(p.pdesc + base + 32 + lane);   // next batch (coalesced)
    const uint32_t K0 = desc & 0xFFFu, D = (desc >> 12) & 0xFFFu, S = desc >> 24;
    const uint32_t* trow = s_tab + (S << p.cls_pad_log2);
    const uint32_t ncls = have ? s_cnt[S] : 0u;
    // class size is the same for every class of one parent: k >> lsz, lsz = log2(ncls)
    const uint32_t lsz = ncls ? (uint32_t)(__ffs(ncls) - 1) : 31u;
    const uint32_t size = ncls ? (a.k >> lsz) : 0u;
    uint32_t szmask[6];
#pragma unroll
    for (int b = 0; b < 6; ++b) szmask[b] = __ballot_sync(0xffffffffu, lsz == (uint32_t)b);

    int nsub = 1;
    bool done = false;
    while (!done) {
      done = true;
      const int per = 32 / nsub;
      for (int sub = 0; sub < nsub; ++sub) {
        const bool mine = have && lane >= sub * per && lane < (sub + 1) * per;
        const uint32_t my_ncls = mine ? ncls : 0u;
        const uint32_t maxcls = __reduce_max_sync(0xffffffffu, my_ncls);
        if (maxcls == 0u) continue;
        // ---- clear
#pragma unroll
        for (int i = 0; i < HT_SLOTS / 32; ++i) { tkey[lane + 32 * i] = 0xFFFFFFFFu; tmask[lane + 32 * i] = 0u; }
        __syncwarp();
        // ---- 1. insert: key -> slot, OR the lane bit into the key's mask
        bool overflow = false;
        for (uint32_t c = 0; c < maxcls; ++c) {
          if (c < my_ncls) {
            const uint32_t key = K0 + (D & (trow[c] & 0xFFFu));
            uint32_t h = (key * 0x9E3779B1u) >> 25;
            int probes = 0;
            while (true) {
              const uint32_t old = atomicCAS(&tkey[h], 0xFFFFFFFFu, key);
              if (old == 0xFFFFFFFFu || old == key) break;
              h = (h + 1u) & (HT_SLOTS - 1u);
              if (++probes >= HT_SLOTS) { overflow = true; break; }
            }
            if (!overflow) atomicOr(&tmask[h], 1u << lane);
          }
        }
        if (__any_sync(0xffffffffu, overflow)) {   // nothing outside the table was touched yet
          nsub = nsub_safe > nsub ? nsub_safe : nsub * 2;   // nsub_safe always fits
          done = false;
          break;
        }
        __syncwarp();
        // ---- 2. leaders, slot-parallel: one L2 read-modify-write per distinct key, all in flight
        {
          uint32_t keyv[HT_SLOTS / 32], startv[HT_SLOTS / 32], totv[HT_SLOTS / 32];
#pragma unroll
          for (int i = 0; i < HT_SLOTS / 32; ++i) {
            keyv[i] = tkey[lane + 32 * i];
            const uint32_t m = tmask[lane + 32 * i];
            uint32_t total = 0;
#pragma unroll
            for (int b = 0; b < 6; ++b)
              if (szmask[b]) total += (uint32_t)__popc(m & szmask[b]) << (logk - b);   // warp-uniform skip
            totv[i] = total;
            startv[i] = keyv[i] != 0xFFFFFFFFu ? __ldcg(row + keyv[i]) : 0u;
          }
#pragma unroll
          for (int i = 0; i < HT_SLOTS / 32; ++i) {
            if (keyv[i] != 0xFFFFFFFFu) {
              __stcg(row + keyv[i], startv[i] + totv[i]);
              tstart[lane + 32 * i] = startv[i];
            }
          }
        }
        __syncwarp();
        // ---- 3. ranks + selection test; hits are queued and resolved with full lanes
        uint32_t qn = 0;                            // warp-uniform queue length
        for (uint32_t c = 0; c < maxcls; ++c) {
          uint32_t n_hit = 0, rank0 = 0, first = 0;
          if (c < my_ncls) {
            const uint32_t key = K0 + (D & (trow[c] & 0xFFFu));
            uint32_t h = (key * 0x9E3779B1u) >> 25;
            while (tkey[h] != key) h = (h + 1u) & (HT_SLOTS - 1u);
            const uint32_t m = tmask[h] & lt_mask;
            uint32_t excl = 0;
#pragma unroll
            for (int b = 0; b < 6; ++b)
              if (szmask[b]) excl += (uint32_t)__popc(m & szmask[b]) << (logk - b);      // warp-uniform skip
            rank0 = tstart[h] + excl;
            const uint32_t pos0 = p.in_shift != 0xFFFFFFFFu ? (rank0 & (Pm - 1u)) : (rank0 % Pm);
            const uint32_t end = pos0 + size;      // size <= k <= Pm: wraps at most once
            first = PRE(pos0);
            n_hit = end <= Pm ? PRE(end) - first : (nm - first) + PRE(end - Pm);
          }
          const uint32_t hm = __ballot_sync(0xffffffffu, n_hit != 0u);
          if (hm) {
            const uint32_t at = qn + (uint32_t)__popc(hm & lt_mask);
            if (n_hit) {
              if (at < (uint32_t)HQ_CAP) hq[at] = make_uint4(rank0, (uint32_t)lane | (c << 5), first, n_hit);
              else emit_hits(rank0, S, trow[c], first, n_hit, (base + lane) * (uint64_t)a.k);   // queue full
            }
            qn = min(qn + (uint32_t)__popc(hm), (uint32_t)HQ_CAP);
          }
        }
        __syncwarp();
        for (uint32_t q0 = 0; q0 < qn; q0 += 32) {
          const bool on = q0 + lane < qn;
          const uint4 e = on ? hq[q0 + lane] : make_uint4(0u, 0u, 0u, 0u);
          const uint32_t owner = e.y & 31u, cc = e.y >> 5;
          const uint32_t S_o = __shfl_sync(0xffffffffu, S, owner);
          if (on) emit_hits(e.x, S_o, s_tab[(S_o << p.cls_pad_log2) + cc], e.z, e.w, (base + owner) * (uint64_t)a.k);
        }
        __syncwarp();
      }
    }
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
struct FlatBatch {
  uint32_t desc;   // K0 | D << 12 | S << 24 of this lane's parent
  uint32_t pk;     // excl | log2(ncls) << 16
  uint32_t ncls, excl, R;
};

__device__ __forceinline__ FlatBatch flat_batch(uint32_t desc, uint32_t ncls, int lane) {
  FlatBatch b;
  uint32_t incl = ncls;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  b.desc = desc; b.ncls = ncls; b.excl = incl - ncls;
  b.R = __shfl_sync(0xffffffffu, incl, 31);
  b.pk = b.excl | ((ncls ? (uint32_t)(__ffs(ncls) - 1) : 0u) << 16);
  return b;
}

struct FlatRecord {
  bool valid;
  int owner;        // lane holding the record's parent
  uint32_t S, ent, key, lsz;   // split set, packed class entry, box key, log2(k / class size)
};

__device__ __forceinline__ FlatRecord flat_record(const FlatBatch& b, uint32_t r0, int lane, const uint32_t* s_tab,
                                                  uint32_t cls_pad_log2) {
  FlatRecord rec;
  const uint32_t r = r0 + (uint32_t)lane;
  rec.valid = r < b.R;
  const uint32_t rel = b.excl - r0;                     // wraps (>= 32) when the parent starts earlier
  const uint32_t starts = __reduce_or_sync(0xffffffffu, (b.ncls && rel < 32u) ? (1u << rel) : 0u);
  const uint32_t before = (uint32_t)__popc(__ballot_sync(0xffffffffu, b.ncls && b.excl < r0));
  const uint32_t le_mask = 0xffffffffu >> (31 - lane);
  rec.owner = rec.valid ? (int)(before + (uint32_t)__popc(starts & le_mask)) - 1 : 0;
  const uint32_t od = __shfl_sync(0xffffffffu, b.desc, rec.owner);
  const uint32_t opk = __shfl_sync(0xffffffffu, b.pk, rec.owner);
  rec.S = od >> 24;
  rec.ent = rec.valid ? s_tab[(rec.S << cls_pad_log2) + (r - (opk & 0xFFFFu))] : 0u;
  rec.key = rec.valid ? (od & 0xFFFu) + ((od >> 12) & rec.ent & 0xFFFu) : 0xFFFFFFFFu;
  rec.lsz = rec.valid ? (opk >> 16) : 31u;
  return rec;
}

// Sum of the class sizes k >> lsz over the peers (tot) and over the peers in lower lanes (ex).
__device__ __forceinline__ void peer_sizes(uint32_t peers, uint32_t lsz, int logk, int nb, uint32_t lt_mask,
                                           uint32_t& ex, uint32_t& tot) {
  ex = 0; tot = 0;
  for (int b = 0; b < nb; ++b) {
    const uint32_t m = __ballot_sync(0xffffffffu, lsz == (uint32_t)b);
    if (m) {                                             // warp-uniform
      const uint32_t mb = peers & m;
      ex += (uint32_t)__popc(mb & lt_mask) << (logk - b);
      tot += (uint32_t)__popc(mb) << (logk - b);
    }
  }
}

// (K0, D, S) from the two candidate child coordinates in 4 vectorised key dims
__device__ __forceinline__ uint32_t describe_parent4(const BoxPrkParams& p, uint64_t i, const float (&lo)[4],
                                                     const float (&ih)[4]) {
  float vm[4], vp[4];
  key_children4(p.a, p.spec, i, vm, vp);
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

// ------------------------------------------------------------------ pass A (flat form)
// CTA per segment, warps interleave its batches; one aggregated shared-memory atomic per
// (round, distinct key) instead of one per class record (box-ordered input makes most lanes
// of a batch share their keys, which serialises plain atomics 32-way).
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
  const int logk = __ffs(a.k) - 1;
  const int nb = (int)p.cls_pad_log2 + 1;
  const uint32_t seg = blockIdx.x;
  const uint64_t i0 = (uint64_t)seg * p.parents_per_seg;
  const uint64_t i1 = min(a.n, i0 + p.parents_per_seg);
  for (uint64_t base = i0 + (threadIdx.x & ~31u); base < i1; base += blockDim.x) {
    const uint64_t i = base + lane;
    const bool have = i < i1;
    uint32_t desc = 0;
    if (have) {
      if (VEC4) desc = describe_parent4(p, i, lo, ih);
      else { uint32_t K0, D, S; describe_parent(p, i, K0, D, S); desc = K0 | (D << 12) | (S << 24); }
      p.pdesc[i] = desc;                                 // pass B re-reads 4 coalesced bytes per parent
    }
    const FlatBatch b = flat_batch(desc, have ? s_cnt[desc >> 24] : 0u, lane);
    for (uint32_t r0 = 0; r0 < b.R; r0 += 32) {
      const FlatRecord rec = flat_record(b, r0, lane, s_tab, p.cls_pad_log2);
      const uint32_t peers = __match_any_sync(0xffffffffu, rec.key);
      uint32_t ex, tot;
      peer_sizes(peers, rec.lsz, logk, nb, lt_mask, ex, tot);
      if (rec.valid && (peers & lt_mask) == 0u) atomicAdd(&hist[rec.key], tot);
    }
  }
  __syncthreads();
  uint32_t* row = p.table + (uint64_t)seg * p.n_keys;
  for (uint32_t i = threadIdx.x; i < p.n_keys; i += blockDim.x) row[i] = hist[i];
}

// ------------------------------------------------------------------ pass B (flat form)
// One warp per segment walks its batches in order.  Per round: one L2 atomicAdd per distinct
// key on the segment's (relative) counter row hands the key's records their rank run; the
// selection test and the rare hits follow with every lane holding a record.
template <bool PRE_SMEM, bool MEM_SMEM>
__global__ void __launch_bounds__(256, 4)
box_rank_flat_kernel(const BoxPrkParams p) {
  extern __shared__ uint32_t smem[];
  const ExpandArgs& a = p.a;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t* cursor = smem;
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
    cursor += 2u * p.out_per_part;
    s_pre16 = reinterpret_cast<uint16_t*>(cursor);
    cursor += ((p.in_per_part + 4u) & ~3u) / 2u;
    for (uint32_t i = threadIdx.x; i <= p.in_per_part; i += blockDim.x) s_pre16[i] = (uint16_t)p.pre[i];
    for (uint32_t i = threadIdx.x; i < p.out_per_part; i += blockDim.x) s_sp[i] = p.sp[i];
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

  const uint64_t i0 = (uint64_t)seg * p.parents_per_seg;
  const uint64_t i1 = min(a.n, i0 + p.parents_per_seg);
  const uint32_t Pm = p.in_per_part, nm = p.out_per_part;
  const int logk = __ffs(a.k) - 1;
  const int nb = (int)p.cls_pad_log2 + 1;

  uint32_t ndesc = i0 + lane < i1 ? __ldcs(p.pdesc + i0 + lane) : 0u;
  for (uint64_t base = i0; base < i1; base += 32) {
    const uint32_t desc = ndesc;
    const bool have = base + lane < i1;
    if (base + 32 + lane < i1) ndesc = __ldcs(p.pdesc + base + 32 + lane);   // next batch (coalesced)
    const FlatBatch b = flat_batch(desc, have ? s_cnt[desc >> 24] : 0u, lane);
    for (uint32_t r0 = 0; r0 < b.R; r0 += 32) {
      const FlatRecord rec = flat_record(b, r0, lane, s_tab, p.cls_pad_log2);
      const uint32_t peers = __match_any_sync(0xffffffffu, rec.key);
      uint32_t ex, tot;
      peer_sizes(peers, rec.lsz, logk, nb, lt_mask, ex, tot);
      const int leader = __ffs(peers) - 1;
      uint32_t st = 0;
      if (rec.valid && lane == leader) st = atomicAdd(row + rec.key, tot);
      st = __shfl_sync(0xffffffffu, st, leader);
      if (rec.valid) {
        const uint32_t rank0 = st + __ldg(p.key_start + rec.key) + ex;
        const uint32_t size = a.k >> rec.lsz;
        uint32_t part0, pos0;
        if (p.in_shift != 0xFFFFFFFFu) { part0 = rank0 >> p.in_shift; pos0 = rank0 & (Pm - 1u); }
        else { part0 = rank0 / Pm; pos0 = rank0 - part0 * Pm; }
        const uint32_t end = pos0 + size;                // size <= k <= Pm: wraps at most once
        const uint32_t first = PRE(pos0);
        const uint32_t n_hit = end <= Pm ? PRE(end) - first : (nm - first) + PRE(end - Pm);
        if (n_hit) {
          const uint16_t* mem = members + (size_t)rec.S * a.k + (rec.ent >> 22);
          const uint64_t child0 = (base + (uint64_t)rec.owner) * (uint64_t)a.k;
          for (uint32_t hh = 0; hh < n_hit; ++hh) {
            uint32_t e = first + hh;
            const bool wrap = e >= nm;
            if (wrap) e -= nm;
            const uint2 pl = sp[e];                      // (position, output slot l)
            const uint32_t tt = wrap ? pl.x + Pm - pos0 : pl.x - pos0;
            p.idx_out[(uint64_t)(part0 + (wrap ? 1u : 0u)) * nm + pl.y] = (uint32_t)(child0 + mem[tt]);
          }
        }
      }
      __syncwarp();   // this round's atomics are ordered before the next round's (same-key runs)
    }
  }
}

// ------------------------------------------------------------------ pass B (class-lane form)
// smem layout: [warps][n_keys] counters | packed class table | pre[in_per_part+1] + sp (if PRE_SMEM)
//              | members (if MEM_SMEM)
// Lane c owns class c of the current parent.  Work is split into a minimal SERIAL phase
// (counter read-modify-write per parent, results parked in registers) and an INDEPENDENT
// phase (selection tests of BR_GROUP parents at once, then the rare hits): a lone warp
// pays ~7 cycles per dependent instruction, so everything that can overlap must.
template <bool PRE_SMEM, bool MEM_SMEM>
__global__ void __launch_bounds__(512, 1)
box_rank_kernel(const BoxPrkParams p) {
  extern __shared__ uint32_t smem[];
  const ExpandArgs& a = p.a;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint32_t* counter = smem + (size_t)warp * p.n_keys;
  uint32_t* cursor = smem + (size_t)p.warps_per_cta * p.n_keys;

  const uint32_t cls_pad = 1u << p.cls_pad_log2;
  uint32_t* s_tab = cursor; cursor += p.n_sets * cls_pad;
  for (uint32_t i = threadIdx.x; i < p.n_sets * cls_pad; i += blockDim.x) s_tab[i] = p.cls_packed[i];
  const uint32_t* pre = p.pre;
  const uint2* sp = p.sp;
  if (PRE_SMEM) {
    uint32_t* s_pre = cursor;
    cursor += (p.in_per_part + 2u) & ~1u;  // keep 8-byte alignment for sp
    uint2* s_sp = reinterpret_cast<uint2*>(cursor);
    cursor += 2u * p.out_per_part;
    for (uint32_t i = threadIdx.x; i <= p.in_per_part; i += blockDim.x) s_pre[i] = p.pre[i];
    for (uint32_t i = threadIdx.x; i < p.out_per_part; i += blockDim.x) s_sp[i] = p.sp[i];
    pre = s_pre; sp = s_sp;
  }
  const uint16_t* members = p.members;
  if (MEM_SMEM) {
    uint16_t* s_mem = reinterpret_cast<uint16_t*>(cursor);
    for (uint32_t i = threadIdx.x; i < p.n_sets * a.k; i += blockDim.x) s_mem[i] = p.members[i];
    members = s_mem;
  }

  const uint32_t seg = blockIdx.x * p.warps_per_cta + warp;
  const bool seg_ok = seg < p.n_segments;
  if (seg_ok) {
    // counter[key] = start[seg][key] = (smaller keys, all segments) + (same key, earlier segments)
    const uint32_t* row = p.table + (uint64_t)seg * p.n_keys;   // folded: absolute start ranks
    for (uint32_t i = lane; i < p.n_keys; i += 32) counter[i] = row[i];
  }
  __syncthreads();
  if (!seg_ok) return;

  const uint64_t i0 = (uint64_t)seg * p.parents_per_seg;
  const uint64_t i1 = min(a.n, i0 + p.parents_per_seg);
  const uint32_t Pm = p.in_per_part, nm = p.out_per_part;
  uint32_t nK0 = 0, nD = 0, nS = 0;
  if (i0 + lane < i1) describe_parent(p, i0 + lane, nK0, nD, nS);
  for (uint64_t base = i0; base < i1; base += BR_BATCH) {
    const uint32_t bK0 = nK0, bD = nD, bS = nS;
    if (p.debug & 8) { nK0 = (uint32_t)(base * 2654435761u + lane * 40503u) & (p.n_keys - 1u) & 0x6DBu; nD = 0x249u; nS = 15u; }
    else
    if (base + BR_BATCH + lane < i1) describe_parent(p, base + BR_BATCH + lane, nK0, nD, nS);  // next batch
    const int cnt = (int)min((uint64_t)BR_BATCH, i1 - base);

#pragma unroll 1
    for (int h0 = 0; h0 < cnt; h0 += BR_GROUP) {
      // ---- phase 1: the only serial part.  Per parent: counter[key] -> register, += size.
      uint32_t ent[BR_GROUP], cur[BR_GROUP];
#pragma unroll
      for (int q = 0; q < BR_GROUP; ++q) {
        const int pp = h0 + q;                       // warp-uniform
        const uint32_t K0 = __shfl_sync(0xffffffffu, bK0, pp & 31);
        const uint32_t D = __shfl_sync(0xffffffffu, bD, pp & 31);
        const uint32_t S = __shfl_sync(0xffffffffu, bS, pp & 31);
        uint32_t e = BP_INVALID;
        if (pp < cnt && (uint32_t)lane < cls_pad) e = s_tab[(S << p.cls_pad_log2) + lane];
        ent[q] = e;
        const uint32_t key = K0 + (D & (e & 0xFFFu));
        uint32_t cv = 0;
        if (e != BP_INVALID && !(p.debug & 4)) {     // classes of one parent have distinct keys
          cv = counter[key];
          counter[key] = cv + ((e >> 12) & 0x3FFu) + 1u;
        }
        cur[q] = cv;
        __syncwarp();                                // visible to the next parent's classes
      }
      if (p.debug & 2) continue;
      // ---- phase 2a: selection test of every class run, all independent (ILP)
      uint32_t first[BR_GROUP], nhit[BR_GROUP];
#pragma unroll
      for (int q = 0; q < BR_GROUP; ++q) {
        const uint32_t size = ((ent[q] >> 12) & 0x3FFu) + 1u;
        const uint32_t rank0 = cur[q];
        const uint32_t pos0 = p.in_shift != 0xFFFFFFFFu ? (rank0 & (Pm - 1u)) : (rank0 % Pm);
        const uint32_t end = pos0 + size;            // size <= k <= Pm: wraps at most once
        uint32_t f = 0, nh = 0;
        if (ent[q] != BP_INVALID) {
          f = pre[pos0];
          nh = end <= Pm ? pre[end] - f : (nm - f) + pre[end - Pm];
        }
        first[q] = f; nhit[q] = (p.debug & 1) ? 0u : nh;
      }
      // ---- phase 2b: resolve the (rare) hits: ~1 selected child per parent
#pragma unroll
      for (int q = 0; q < BR_GROUP; ++q) {
        const int pp = h0 + q;
        const uint32_t S = __shfl_sync(0xffffffffu, bS, pp & 31);     // (outside the divergent branch)
        if (nhit[q]) {
          const uint32_t rank0 = cur[q];
          uint32_t part0, pos0;
          if (p.in_shift != 0xFFFFFFFFu) { part0 = rank0 >> p.in_shift; pos0 = rank0 & (Pm - 1u); }
          else { part0 = rank0 / Pm; pos0 = rank0 - part0 * Pm; }
          const uint16_t* mem = members + (size_t)S * a.k + (ent[q] >> 22);
          const uint64_t child0 = (base + pp) * (uint64_t)a.k;
          for (uint32_t h = 0; h < nhit[q]; ++h) {
            uint32_t e = first[q] + h;
            const bool wrap = e >= nm;
            if (wrap) e -= nm;
            const uint2 pl = sp[e];                  // (position, output slot l)
            const uint32_t tt = wrap ? pl.x + Pm - pos0 : pl.x - pos0;
            p.idx_out[(uint64_t)(part0 + (wrap ? 1u : 0u)) * nm + pl.y] = (uint32_t)(child0 + mem[tt]);
          }
        }
      }
    }
  }
}

struct PrkPlan {
  uint32_t n_keys, n_segments, n_sets, max_cls, cls_pad_log2;
  int warps_per_cta, grid_b, pre_in_smem, mem_in_smem, mode;
  uint64_t parents_per_seg;
  size_t smem_b;
  size_t off_pdesc, off_table, off_pre, off_fill, off_sp, off_cnt, off_fm, off_size, off_off, off_packed, off_members, off_cid, off_idx, total;
};

// rank_mode 4 (default): flat kernels (records of a batch spread over all lanes, match_any ranking,
//              one L2 atomic per distinct key and round); rows stay relative, no fold pass.
// rank_mode 3: hash kernel, record-parallel, counter rows in global memory.
// rank_mode 2: sweep kernel, counter rows in global memory, 32 warps per SM.
// rank_mode 1: class-lane kernel, counter rows in shared memory (fewer warps).
static int rank_mode() {
  const char* m = getenv("MCCB200_RANK_MODE");
  return m ? atoi(m) : 4;
}

static bool make_prk_plan(PrkPlan& pl, uint64_t n, int k, int n_dims, int bits, uint64_t in_per_part,
                          uint64_t out_per_part_max) {
  const int B = n_dims * bits;
  if (B > BP_MAX_KEY_BITS || B < 1) return false;
  if (k > BP_MAX_K || k < 2) return false;
  if (in_per_part >= 0x80000000ull || in_per_part < (uint64_t)k) return false;  // a class run wraps at most once
  if (out_per_part_max >= 0x80000000ull) return false;
  const uint64_t total_children = n * (uint64_t)k;
  if (total_children > 0x100000000ull) return false;
  pl.n_keys = 1u << B;
  pl.n_sets = 1u << n_dims;
  pl.max_cls = (uint32_t)k < pl.n_sets ? (uint32_t)k : pl.n_sets;
  if (pl.max_cls > 32) return false;  // pass B gives every class of a parent its own lane
  pl.cls_pad_log2 = 0;
  while ((1u << pl.cls_pad_log2) < pl.max_cls) ++pl.cls_pad_log2;
  const size_t smem_cap = 220 * 1024;
  const size_t row = (size_t)pl.n_keys * 4;
  const size_t pre_bytes = ((size_t)((in_per_part + 2) & ~1ull)) * 4 + (size_t)out_per_part_max * 8;
  const size_t tab_bytes = (size_t)pl.n_sets * (1u << pl.cls_pad_log2) * 4;
  const size_t mem_bytes = (size_t)pl.n_sets * k * 2 + 8;
  pl.pre_in_smem = pre_bytes <= 72 * 1024 ? 1 : 0;
  pl.mem_in_smem = mem_bytes <= 16 * 1024 ? 1 : 0;
  pl.mode = rank_mode();
  if (pl.mode == 2) pl.pre_in_smem = pre_bytes <= 36 * 1024 ? 1 : 0;   // 4 CTAs per SM share 227 KB
  size_t pre_bytes3 = ((size_t)((in_per_part + 4) & ~3ull)) * 2 + (size_t)out_per_part_max * 8;
  if (pl.mode >= 3) pl.pre_in_smem = (pre_bytes3 <= 24 * 1024 && out_per_part_max < 65536) ? 1 : 0;
  size_t shared_part = tab_bytes + (pl.pre_in_smem ? (pl.mode >= 3 ? pre_bytes3 : pre_bytes) : 0) + (pl.mem_in_smem ? mem_bytes : 0);
  if (pl.mode == 3) shared_part += (size_t)8 * 3 * HT_SLOTS * 4 + (size_t)8 * HQ_CAP * 16 + (size_t)pl.n_sets * 4;
  if (pl.mode == 4) shared_part += (size_t)pl.n_sets * 4 + 8;
  int w;
  uint64_t segs;
  const int sms = sm_count();
  if (pl.mode >= 2) {
    w = 8;
    segs = (uint64_t)sms * 32;
  } else {
    w = (int)((smem_cap - shared_part) / row);
    if (w > 16) w = 16;
    if (w < 1) return false;
    segs = (uint64_t)sms * (uint64_t)w;
  }
  pl.warps_per_cta = w;
  if (segs > n) segs = n ? n : 1;
  pl.parents_per_seg = (n + segs - 1) / segs;
  pl.n_segments = (uint32_t)((n + pl.parents_per_seg - 1) / pl.parents_per_seg);
  pl.grid_b = (int)((pl.n_segments + w - 1) / w);
  pl.smem_b = (pl.mode >= 2 ? 0 : (size_t)w * row) + shared_part;
  const uint64_t n_out_max = total_children / in_per_part * out_per_part_max;
  size_t cur = 0;
  auto take = [&](size_t bytes) { size_t o = cur; cur += align_up(bytes, 256); return o; };
  pl.off_pdesc = take((size_t)n * 4);
  pl.off_table = take(((size_t)pl.n_segments + 1) * row);
  pl.off_pre = take(((size_t)in_per_part + 2) * 4);
  pl.off_fill = take(((size_t)in_per_part + 2) * 4);
  pl.off_sp = take((size_t)out_per_part_max * 8);
  pl.off_cnt = take((size_t)pl.n_sets * 4);
  pl.off_fm = take((size_t)pl.n_sets * pl.max_cls * 4);
  pl.off_size = take((size_t)pl.n_sets * pl.max_cls * 4);
  pl.off_off = take((size_t)pl.n_sets * pl.max_cls * 4);
  pl.off_packed = take((size_t)pl.n_sets * (1u << pl.cls_pad_log2) * 4);
  pl.off_members = take((size_t)pl.n_sets * k * 2);
  pl.off_cid = take((size_t)pl.n_sets * pl.n_sets * 2);
  pl.off_idx = take((size_t)n_out_max * 4);
  pl.total = cur;
  return true;
}

}  // namespace mccb

using namespace mccb;

extern "C" size_t mccb200_box_prk_step_workspace_bytes(uint64_t n, int k, int n_dims, int bits,
                                                       uint64_t in_per_part) {
  PrkPlan pl;
  // 0 = unsupported shape.  Sized for out_per_part <= in_per_part.
  if (!make_prk_plan(pl, n, k, n_dims, bits, in_per_part, in_per_part)) return 0;
  return pl.total;
}

extern "C" int mccb200_box_prk_step(const float* y0, uint64_t n, int d, const mccb200_drift* drift, float dt,
                                    const mccb200_cubature* cub, const int32_t* dims_host, int n_dims, int bits,
                                    const float* bounds, const uint32_t* idx_local, uint64_t out_per_part,
                                    uint64_t in_per_part, float* y1, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BoxPrkParams p{};
  int rc = make_box_spec(p.spec, dims_host, n_dims, bits, d);
  if (rc) return rc;
  rc = fill_expand_args(p.a, y0, n, d, 0, drift, dt, cub);
  if (rc) return rc;
  MCCB_REQUIRE(bounds && idx_local && y1 && workspace, "box_prk_step: NULL argument");
  if (cub->points != nullptr)
    return fail(MCCB200_ERR_UNSUPPORTED, "box_prk_step: only the Hadamard cubature is fused");
  const uint64_t total = n * (uint64_t)cub->k;
  MCCB_REQUIRE(in_per_part >= 1 && total % in_per_part == 0, "box_prk_step: in_per_part=%llu does not divide n*k=%llu",
               (unsigned long long)in_per_part, (unsigned long long)total);
  MCCB_REQUIRE(out_per_part >= 1, "box_prk_step: bad out_per_part");
  PrkPlan pl;
  if (!make_prk_plan(pl, n, cub->k, n_dims, bits, in_per_part, out_per_part))
    return fail(MCCB200_ERR_UNSUPPORTED, "box_prk_step: shape unsupported (key bits %d > %d, k > %d or in_per_part < k)",
                n_dims * bits, BP_MAX_KEY_BITS, BP_MAX_K);
  if (workspace_bytes < pl.total)
    return fail(MCCB200_ERR_WORKSPACE_TOO_SMALL, "box_prk_step: workspace %zu < %zu", workspace_bytes, pl.total);
  if (n == 0) return MCCB200_OK;

  char* w = (char*)workspace;
  uint32_t* table = (uint32_t*)(w + pl.off_table);
  p.pdesc = (uint32_t*)(w + pl.off_pdesc);
  uint32_t* pre = (uint32_t*)(w + pl.off_pre);
  uint32_t* fill = (uint32_t*)(w + pl.off_fill);
  uint2* sp = (uint2*)(w + pl.off_sp);
  uint32_t* cls_count = (uint32_t*)(w + pl.off_cnt);
  uint32_t* cls_fm = (uint32_t*)(w + pl.off_fm);
  uint32_t* cls_size = (uint32_t*)(w + pl.off_size);
  uint32_t* cls_off = (uint32_t*)(w + pl.off_off);
  uint32_t* cls_packed = (uint32_t*)(w + pl.off_packed);
  uint16_t* members = (uint16_t*)(w + pl.off_members);
  uint16_t* cid = (uint16_t*)(w + pl.off_cid);
  uint32_t* idx_out = (uint32_t*)(w + pl.off_idx);

  p.bounds = bounds;
  p.n_keys = pl.n_keys;
  p.n_segments = pl.n_segments;
  p.parents_per_seg = pl.parents_per_seg;
  p.in_per_part = (uint32_t)in_per_part;
  p.out_per_part = (uint32_t)out_per_part;
  p.in_shift = 0xFFFFFFFFu;
  if ((in_per_part & (in_per_part - 1)) == 0) { p.in_shift = 0; while ((1ull << p.in_shift) < in_per_part) ++p.in_shift; }
  p.n_sets = pl.n_sets; p.max_cls = pl.max_cls;
  p.cls_count = cls_count; p.cls_fm = cls_fm; p.cls_size = cls_size; p.cls_off = cls_off; p.members = members;
  p.cls_packed = cls_packed; p.cls_pad_log2 = pl.cls_pad_log2;
  p.pre = pre; p.sp = sp; p.table = table; p.idx_out = idx_out;
  p.warps_per_cta = pl.warps_per_cta;
  p.pre_in_smem = pl.pre_in_smem; p.tab_in_smem = 1;
  { const char* dbg = getenv("MCCB200_DEBUG_RANK"); p.debug = dbg ? atoi(dbg) : 0; }

  phase_mark("box_prk.tables", st);
  box_tables_kernel<<<(pl.n_sets + 63) / 64, 64, 0, st>>>(p, cls_count, cls_fm, cls_size, cls_off, cls_packed, members, cid);
  box_selection_kernel<<<1, 1024, 0, st>>>(idx_local, p.out_per_part, p.in_per_part, pre, fill, sp);

  phase_mark("box_prk.count", st);
  uint32_t* key_tot = table + (uint64_t)pl.n_segments * pl.n_keys;
  p.key_start = key_tot;
  if (pl.mode == 4) {
    const size_t smem_a = ((size_t)pl.n_keys + (size_t)pl.n_sets * (1u << pl.cls_pad_log2) + pl.n_sets) * 4;
    if (box_dims_vec4(p.spec, p.a)) box_count_flat_kernel<true><<<pl.n_segments, 256, smem_a, st>>>(p);
    else box_count_flat_kernel<false><<<pl.n_segments, 256, smem_a, st>>>(p);
  } else {
    box_count_kernel<<<pl.n_segments, 256, (size_t)pl.n_keys * 4, st>>>(p);
  }
  phase_mark("box_prk.scan", st);
  box_scan_columns_kernel<<<(pl.n_keys + 31) / 32, 1024, 0, st>>>(table, pl.n_segments, pl.n_keys, key_tot);
  box_scan_keys_kernel<<<1, 1024, 0, st>>>(key_tot, pl.n_keys);
  if (pl.mode != 4)   // the flat rank kernel adds key_start itself
    box_scan_fold_kernel<<<(int)min((uint64_t)sm_count() * 8, ((uint64_t)pl.n_segments * pl.n_keys + 255) / 256), 256, 0, st>>>(
        table, pl.n_segments, pl.n_keys, key_tot);

  phase_mark("box_prk.rank", st);
#define MCCB_RANK(PS, MS)                                                                                     \
  do {                                                                                                        \
    cudaFuncSetAttribute(box_rank_kernel<PS, MS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_b); \
    box_rank_kernel<PS, MS><<<pl.grid_b, pl.warps_per_cta * 32, pl.smem_b, st>>>(p);                          \
  } while (0)
#define MCCB_SWEEP(PS, MS)                                                                                          \
  do {                                                                                                              \
    cudaFuncSetAttribute(box_rank_sweep_kernel<PS, MS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_b); \
    box_rank_sweep_kernel<PS, MS><<<pl.grid_b, pl.warps_per_cta * 32, pl.smem_b, st>>>(p);                          \
  } while (0)
#define MCCB_HASH(PS, MS)                                                                                          \
  do {                                                                                                             \
    cudaFuncSetAttribute(box_rank_hash_kernel<PS, MS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_b); \
    box_rank_hash_kernel<PS, MS><<<pl.grid_b, pl.warps_per_cta * 32, pl.smem_b, st>>>(p);                          \
  } while (0)
#define MCCB_FLAT(PS, MS)                                                                                          \
  do {                                                                                                             \
    cudaFuncSetAttribute(box_rank_flat_kernel<PS, MS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_b); \
    box_rank_flat_kernel<PS, MS><<<pl.grid_b, pl.warps_per_cta * 32, pl.smem_b, st>>>(p);                          \
  } while (0)
  if (pl.mode == 4) {
    if (pl.pre_in_smem && pl.mem_in_smem) MCCB_FLAT(true, true);
    else if (pl.pre_in_smem) MCCB_FLAT(true, false);
    else if (pl.mem_in_smem) MCCB_FLAT(false, true);
    else MCCB_FLAT(false, false);
  } else if (pl.mode == 3) {
    if (pl.pre_in_smem && pl.mem_in_smem) MCCB_HASH(true, true);
    else if (pl.pre_in_smem) MCCB_HASH(true, false);
    else if (pl.mem_in_smem) MCCB_HASH(false, true);
    else MCCB_HASH(false, false);
  } else if (pl.mode == 2) {
    if (pl.pre_in_smem && pl.mem_in_smem) MCCB_SWEEP(true, true);
    else if (pl.pre_in_smem) MCCB_SWEEP(true, false);
    else if (pl.mem_in_smem) MCCB_SWEEP(false, true);
    else MCCB_SWEEP(false, false);
  } else {
    if (pl.pre_in_smem && pl.mem_in_smem) MCCB_RANK(true, true);
    else if (pl.pre_in_smem) MCCB_RANK(true, false);
    else if (pl.mem_in_smem) MCCB_RANK(false, true);
    else MCCB_RANK(false, false);
  }
#undef MCCB_RANK
#undef MCCB_SWEEP
#undef MCCB_HASH
#undef MCCB_FLAT

  // gather: the HBM-bound fused expand+select materialises the selected children
  ExpandArgs g = p.a;
  g.idx = idx_out; g.perm = nullptr; g.out_per_part = 1; g.in_per_part = 1;
  g.n_out = total / in_per_part * out_per_part; g.y1 = y1;
  rc = run_expand_select(g, st);
  if (rc) return rc;
  count_launches(7);
  return check_launch("box_prk_step");
}
