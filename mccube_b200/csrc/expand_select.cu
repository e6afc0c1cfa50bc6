// Fused expand + select: the MCCSolver step without the n*k*d expansion (sm_100a).
//
// Replaces, for /root/reference/mccube/_solvers.py:125-144,
//   Euler.step (y0 + vf_prod) -> MCCTerm.prod broadcast add (_term.py:70-77)
//   -> reshape [n*k, d] (_solvers.py:136) -> kernel gather p[idx] (random.py:65 /
//   base.py:170-177 with the shared local index vector of quirk Q4).
// Only the selected children are ever computed: read n_out parent rows, write n_out rows
// (2 * n_out * d * 4 algorithmic bytes).  HBM-bound; 128-bit loads/stores when d % 4 == 0.
//
// fp32 op order is the reference's, with explicit round-to-nearest intrinsics so ptxas
// cannot contract into FMA:   y1 = y0 + (dt * f + noise).
#include "common.cuh"
#include "expand.cuh"
#include "rng.cuh"

namespace mccb {

template <int VEC> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<1> { using type = float; };

template <int VEC>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[VEC]) {
  using T = typename VecT<VEC>::type;
  T t = __ldg(reinterpret_cast<const T*>(p));
  const float* f = reinterpret_cast<const float*>(&t);
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = f[i];
}

// Streaming (read-once) load: bypass L1 allocation.
template <int VEC>
__device__ __forceinline__ void load_vec_stream(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
  } else if constexpr (VEC == 2) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "l"(p));
  } else {
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v[0]) : "l"(p));
  }
}

template <int VEC>
__device__ __forceinline__ void store_vec_stream(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
  } else if constexpr (VEC == 2) {
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(v[0]), "f"(v[1]) : "memory");
  } else {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" :: "l"(p), "f"(v[0]) : "memory");
  }
}


constexpr int EX_THREADS = 256;
constexpr int EX_UNROLL = 4;

__device__ __forceinline__ uint32_t child_of_row(const ExpandArgs& a, uint64_t r) {
  if (a.perm) {
    uint32_t b = (uint32_t)(r / a.out_per_part);
    uint32_t l = (uint32_t)(r - (uint64_t)b * a.out_per_part);
    return a.perm[(uint64_t)b * a.in_per_part + a.idx[l]];
  }
  return a.idx ? a.idx[a.idx_stride > 1u ? r * a.idx_stride : r] : (uint32_t)r;
}

// jr.randint(key, (count,), 0, span)[i] for int32 (the arithmetic of randint_kernel, rng_indices.cu)
__device__ __forceinline__ uint32_t draw_child(u32x2 k_hi, u32x2 k_lo, uint32_t i, uint32_t count, uint32_t span,
                                               uint32_t mult, bool part) {
  const uint32_t lo = random_bits32_at(k_lo, i, count, part);
  if (mult == 0u) return (span & (span - 1u)) == 0u ? (lo & (span - 1u)) : (lo % span);
  const uint32_t hi = random_bits32_at(k_hi, i, count, part);
  return ((hi % span) * mult + (lo % span)) % span;
}

// MODE 0: plain (the hot path: nothing below costs it an instruction); 1: scatter to peer shards;
// 2: fused sub-steps; 3: pull -- parents are read from the peer shard that owns them; 4: plain with the child ids
// drawn in the kernel (one lane per output row runs the threefry block, the row's lanes fetch the id by shuffle).
template <int VEC, int DRIFT, bool TABLE, int MODE = 0>
__global__ void __launch_bounds__(EX_THREADS)
expand_kernel(const ExpandArgs a) {
  constexpr bool SCATTER = MODE == 1;
  constexpr bool SUBSTEPS = MODE == 2;
  constexpr bool PULL = MODE == 3;
  constexpr bool DRAW = MODE == 4;
  const int row_in_iter = threadIdx.x / a.vpr;
  const int col = (threadIdx.x - row_in_iter * a.vpr) * VEC;
  const bool active = row_in_iter < a.rows_per_iter;
  if (!active) return;
  uint64_t n_out = a.n_out, sc_start = 0;
  if (SCATTER) {   // this rank's block of the owner-grouped slot list, from the replicated counts matrix
    uint64_t run = 0;
    for (uint32_t o = 0; o <= a.sc_rank; ++o) {
      uint64_t per = 0;
      for (uint32_t dd = 0; dd < a.sc_G; ++dd) per += __ldg(a.sc_counts + o * a.sc_G + dd);
      if (o < a.sc_rank) run += per; else n_out = per;
    }
    sc_start = run;
  }

  float mean[VEC], ivar[VEC];
  if (DRIFT == 2) {
    load_vec<VEC>(a.mean + col, mean);
    load_vec<VEC>(a.inv_var + col, ivar);
  }
  // weighted states always take the scalar mapping (run_expand_select): the vector instantiations carry no weight code
  // (the fp64 weight sum cost the VEC = 4 kernels 5 registers = one resident CTA per SM: 1.43 -> 1.58 ms)
  const bool do_w = VEC == 1 && a.weighted && col == 0;
  double wacc = 0.0;
  u32x2 dk_hi{}, dk_lo{};
  uint32_t draw_mult = 0;
  if (DRAW) {
    dk_hi = {__ldg(a.draw_keys), __ldg(a.draw_keys + 1)};
    dk_lo = {__ldg(a.draw_keys + 2), __ldg(a.draw_keys + 3)};
    draw_mult = 65536u % a.draw_span;
    draw_mult = (draw_mult * draw_mult) % a.draw_span;
  }

  const uint64_t rows_per_cta_iter = (uint64_t)a.rows_per_iter * EX_UNROLL;
  for (uint64_t base = (uint64_t)blockIdx.x * rows_per_cta_iter; base < n_out;
       base += (uint64_t)gridDim.x * rows_per_cta_iter) {
    uint32_t c[EX_UNROLL];
    uint64_t r[EX_UNROLL];
    bool ok[EX_UNROLL];
    uint32_t slot[EX_UNROLL];
    uint32_t drawn = 0;
    if (DRAW) {   // a warp holds 32 / vpr rows of every unroll slot: lane l draws for slot l / rpw, row l % rpw
      const int lane = threadIdx.x & 31;
      const int rpw_shift = 5 - a.vpr_shift;
      const int u_l = lane >> rpw_shift, sub_l = lane & ((1 << rpw_shift) - 1);
      if (u_l < EX_UNROLL) {
        const uint64_t row = base + (uint64_t)u_l * a.rows_per_iter + ((threadIdx.x >> 5) << rpw_shift) + sub_l;
        if (row < n_out) {
          drawn = draw_child(dk_hi, dk_lo, a.draw_first + (uint32_t)row, a.draw_count, a.draw_span, draw_mult,
                             a.draw_partitionable != 0);
          if (a.draw_idx_out) a.draw_idx_out[row] = drawn;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < EX_UNROLL; ++u) {
      r[u] = base + (uint64_t)u * a.rows_per_iter + row_in_iter;
      ok[u] = r[u] < n_out;
      if (DRAW) {
        const int lane = threadIdx.x & 31;
        slot[u] = 0u;
        c[u] = __shfl_sync(0xffffffffu, drawn, (u << (5 - a.vpr_shift)) + (lane >> a.vpr_shift));
      } else if (PULL && a.idx64 != nullptr) {                    // 64-bit global child id -> (global parent, point)
        const uint64_t cg = ok[u] ? __ldg(a.idx64 + r[u]) : 0ull;
        slot[u] = (uint32_t)(cg >> a.k_shift);             // parent (n_tot < 2^32), decoded below
        c[u] = (uint32_t)(cg & (uint64_t)(a.k - 1u));
      } else if (SCATTER) {
        slot[u] = ok[u] ? __ldg(a.sc_order + sc_start + r[u]) : 0u;
        c[u] = ok[u] ? (uint32_t)((uint64_t)__ldg(a.idx + slot[u]) - a.sc_child_offset) : 0u;
      } else {
        slot[u] = 0u;
        c[u] = ok[u] ? child_of_row(a, r[u]) : 0u;
      }
    }
    float y[EX_UNROLL][VEC], f[EX_UNROLL][VEC];
    uint32_t pi[EX_UNROLL], pj[EX_UNROLL];
#pragma unroll
    for (int u = 0; u < EX_UNROLL; ++u) {
      if (PULL && a.idx64 != nullptr) {
        pi[u] = slot[u];
        pj[u] = c[u];
      } else if (SUBSTEPS) {                               // k^n_sub children per parent (k a power of two)
        const uint32_t sh = a.k_shift * (uint32_t)a.n_sub;
        pi[u] = sh >= 32u ? 0u : (c[u] >> sh);
        pj[u] = sh >= 32u ? c[u] : (c[u] & ((1u << sh) - 1u));
      } else if (a.k_shift != 0xFFFFFFFFu) {
        pi[u] = c[u] >> a.k_shift;
        pj[u] = c[u] & (a.k - 1u);
      } else {
        pi[u] = c[u] / a.k;
        pj[u] = c[u] - pi[u] * a.k;
      }
      if (ok[u]) {
        if (PULL) {   // global parent id -> (owner rank, row): an NVLink peer load when the owner is another GPU
          const uint32_t owner = (uint32_t)(pi[u] / a.sc_rows_per_rank);
          const uint64_t row = pi[u] - (uint64_t)owner * a.sc_rows_per_rank;
          load_vec_stream<VEC>(a.sc_peers[owner] + row * a.ld + col, y[u]);
        } else {
          load_vec_stream<VEC>(a.y0 + (uint64_t)pi[u] * a.ld + col, y[u]);
          if (DRIFT == 1) load_vec_stream<VEC>(a.f + (uint64_t)pi[u] * a.d + col, f[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < EX_UNROLL; ++u) {
      if (!ok[u]) continue;
      float out[VEC];
      float tab[VEC];
      if (TABLE) load_vec<VEC>(a.points + (uint64_t)pj[u] * a.d + col, tab);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        float fv;
        if (DRIFT == 2) fv = __fmul_rn(__fsub_rn(mean[e], y[u][e]), ivar[e]);
        else if (DRIFT == 1) fv = f[u][e];
        else fv = 0.0f;
        float noise;
        if (TABLE) noise = tab[e];
        else noise = hadamard_neg(pj[u], a.half_mask, (uint32_t)(col + e)) ? -a.scale : a.scale;
        if (SUBSTEPS && !TABLE && DRIFT != 1) {              // fused sub-steps
          float yy = y[u][e];
          for (int sb = 0; sb < a.n_sub; ++sb) {
            const uint32_t j = (pj[u] >> (a.k_shift * (uint32_t)(a.n_sub - 1 - sb))) & (a.k - 1u);
            const float fs = DRIFT == 2 ? __fmul_rn(__fsub_rn(mean[e], yy), ivar[e]) : 0.0f;
            const float ns = hadamard_neg(j, a.half_mask, (uint32_t)(col + e)) ? -a.scale_sub[sb] : a.scale_sub[sb];
            yy = __fadd_rn(yy, __fadd_rn(__fmul_rn(a.dt_sub[sb], fs), ns));
          }
          out[e] = yy;
        } else {
          float step = __fadd_rn(__fmul_rn(a.dt, fv), noise);
          out[e] = __fadd_rn(y[u][e], step);
        }
      }
      if (SCATTER) {   // straight into the destination rank's shard (peer memory)
        const uint32_t dest = (uint32_t)(slot[u] / a.sc_rows_per_rank);
        const uint64_t row = slot[u] - (uint64_t)dest * a.sc_rows_per_rank;
        store_vec_stream<VEC>(a.sc_peers[dest] + row * a.ld + col, out);
        continue;
      }
      store_vec_stream<VEC>(a.y1 + r[u] * a.ld + col, out);
      if (VEC == 4 && a.keycols_out != nullptr && col == a.keycol0)      // packed key columns for the next step's partitioner
        store_vec_stream<VEC>(a.keycols_out + r[u] * 4, out);
      if (do_w) {
        // quirk Q3 (_solvers.py:138-141): weights are flattened from a [k, n] array
        uint64_t cc = c[u];
        uint32_t jn = (uint32_t)(cc / a.n);
        uint64_t in = cc - (uint64_t)jn * a.n;
        float lam = a.lam ? a.lam[jn] : __fdiv_rn(1.0f, (float)a.k);
        const float wc = __fmul_rn(a.y0[in * a.ld + a.d], lam);
        a.y1[r[u] * a.ld + a.d] = wc;
        wacc += (double)wc;
      }
    }
  }
  if (do_w && a.wsum != nullptr) atomicAdd(a.wsum, wacc);   // one add per weight-writing thread (a few thousand)
}

template <int VEC>
__global__ void __launch_bounds__(EX_THREADS)
gather_kernel(const float* __restrict__ in, int ld, int vpr, int rows_per_iter, int tail_col, int tail,
              const uint32_t* __restrict__ idx, uint64_t n_out, float* __restrict__ out) {
  const int row_in_iter = threadIdx.x / vpr;
  const int col = (threadIdx.x - row_in_iter * vpr) * VEC;
  if (row_in_iter >= rows_per_iter) return;
  (void)tail_col; (void)tail;
  for (uint64_t r = (uint64_t)blockIdx.x * rows_per_iter + row_in_iter; r < n_out;
       r += (uint64_t)gridDim.x * rows_per_iter) {
    float v[VEC];
    load_vec_stream<VEC>(in + (uint64_t)idx[r] * ld + col, v);
    store_vec_stream<VEC>(out + r * ld + col, v);
  }
}

// Sum of the weight column (one 4-byte word every `ld` floats: a sector per row) with eight independent loads in
// flight per thread, then y[:, -1] /= sum.  The divide pass re-reads the sectors the sum pass left in L2 when the
// column fits there.  (Round 1: one load in flight per thread, 188 GB/s-equivalent; now bound by sectors per second.)
__global__ void __launch_bounds__(256)
weight_sum_kernel(const float* __restrict__ y, uint64_t n, int ld, double* __restrict__ out) {
  __shared__ double red[8];
  double acc = 0.0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; r + 7 * stride < n; r += 8 * stride) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = ld_strided(y + (r + u * stride) * ld + (ld - 1));
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += (double)v[u];
  }
  for (; r < n; r += stride) acc += (double)ld_strided(y + r * ld + (ld - 1));
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(out, s);
  }
}

__global__ void __launch_bounds__(256)
weight_div_kernel(float* __restrict__ y, uint64_t n, int ld, const double* __restrict__ total) {
  const float s = (float)*total;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; r + 7 * stride < n; r += 8 * stride) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = y[(r + u * stride) * ld + (ld - 1)];
#pragma unroll
    for (int u = 0; u < 8; ++u) y[(r + u * stride) * ld + (ld - 1)] = __fdiv_rn(v[u], s);
  }
  for (; r < n; r += stride) y[r * ld + (ld - 1)] = __fdiv_rn(y[r * ld + (ld - 1)], s);
}

// out = in with the weight column divided by *total: the clone + renormalisation of MCCSolver.step
// (mccube/_solvers.py:146-155: y1 keeps the raw weights, the returned state the normalised ones) in one streaming pass
// over the flat array, 16 bytes per access whatever ld is.
__global__ void __launch_bounds__(256)
normalised_copy_kernel(const float* __restrict__ in, uint64_t n, int ld, const double* __restrict__ total,
                       float* __restrict__ out, int vec_ok) {
  const float s = (float)*total;
  const uint64_t len = n * (uint64_t)ld;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t quads = vec_ok ? len / 4 : 0;
  // column of the first element of a thread's quad, advanced by (4 * stride) mod ld from one quad to the next
  const uint32_t uld = (uint32_t)ld;
  const uint32_t col_step = (uint32_t)((4 * stride) % (uint64_t)ld);
  uint64_t q0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t col = (uint32_t)((4 * q0) % (uint64_t)ld);
  for (; q0 < quads; q0 += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (q0 + u * stride < quads) v[u] = __ldcs(reinterpret_cast<const float4*>(in) + q0 + u * stride);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t q = q0 + u * stride;
      const uint32_t to_w = uld - 1u - col;                        // elements until the weight column
      col += col_step;
      if (col >= uld) col -= uld;
      if (q >= quads) continue;
      if (to_w == 0u) v[u].x = __fdiv_rn(v[u].x, s);
      else if (to_w == 1u) v[u].y = __fdiv_rn(v[u].y, s);
      else if (to_w == 2u) v[u].z = __fdiv_rn(v[u].z, s);
      else if (to_w == 3u) v[u].w = __fdiv_rn(v[u].w, s);
      __stcs(reinterpret_cast<float4*>(out) + q, v[u]);
    }
  }
  for (uint64_t i = quads * 4 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    const float x = in[i];
    out[i] = (i % (uint64_t)ld == (uint64_t)ld - 1u) ? __fdiv_rn(x, s) : x;
  }
}

static int pick_vec(int d, int ld, const void* p0, const void* p1) {
  uintptr_t a = (uintptr_t)p0 | (uintptr_t)p1;
  if (d % 4 == 0 && ld % 4 == 0 && a % 16 == 0) return 4;
  if (d % 2 == 0 && ld % 2 == 0 && a % 8 == 0) return 2;
  return 1;
}

template <int VEC>
static void launch_expand(const ExpandArgs& a, int grid, cudaStream_t st, int pad_smem = 0) {
  const bool table = a.points != nullptr;
  if (a.sc_peers && a.sc_order == nullptr) {   // pull form: Hadamard, drift kinds 0 / 2
    if (a.drift_kind == 2) expand_kernel<VEC, 2, false, 3><<<grid, EX_THREADS, 0, st>>>(a);
    else expand_kernel<VEC, 0, false, 3><<<grid, EX_THREADS, 0, st>>>(a);
    return;
  }
  if (a.sc_peers) {   // scatter form: Hadamard only
    if (a.drift_kind == 2) expand_kernel<VEC, 2, false, 1><<<grid, EX_THREADS, 0, st>>>(a);
    else if (a.drift_kind == 1) expand_kernel<VEC, 1, false, 1><<<grid, EX_THREADS, 0, st>>>(a);
    else expand_kernel<VEC, 0, false, 1><<<grid, EX_THREADS, 0, st>>>(a);
    return;
  }
  if (a.n_sub > 1) {  // fused sub-steps: Hadamard, drift kinds 0 / 2 (checked by the caller)
    if (a.drift_kind == 2) expand_kernel<VEC, 2, false, 2><<<grid, EX_THREADS, 0, st>>>(a);
    else expand_kernel<VEC, 0, false, 2><<<grid, EX_THREADS, 0, st>>>(a);
    return;
  }
  if (a.draw_keys) {  // fused index draw (VEC == 4 checked by the caller)
    if constexpr (VEC == 4) {
#define MCCB_LAUNCH_DRAW(D, T) expand_kernel<4, D, T, 4><<<grid, EX_THREADS, 0, st>>>(a)
      if (a.drift_kind == 2) { if (table) MCCB_LAUNCH_DRAW(2, true); else MCCB_LAUNCH_DRAW(2, false); }
      else if (a.drift_kind == 1) { if (table) MCCB_LAUNCH_DRAW(1, true); else MCCB_LAUNCH_DRAW(1, false); }
      else { if (table) MCCB_LAUNCH_DRAW(0, true); else MCCB_LAUNCH_DRAW(0, false); }
#undef MCCB_LAUNCH_DRAW
    }
    return;
  }
#define MCCB_LAUNCH(D, T) expand_kernel<VEC, D, T><<<grid, EX_THREADS, 0, st>>>(a)
  if (a.drift_kind == 2 && !table && pad_smem > 0) {   // residency cap (co-scheduling experiments)
    cudaFuncSetAttribute(expand_kernel<VEC, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad_smem);
    expand_kernel<VEC, 2, false><<<grid, EX_THREADS, pad_smem, st>>>(a);
    return;
  }
  if (a.drift_kind == 2) { if (table) MCCB_LAUNCH(2, true); else MCCB_LAUNCH(2, false); }
  else if (a.drift_kind == 1) { if (table) MCCB_LAUNCH(1, true); else MCCB_LAUNCH(1, false); }
  else { if (table) MCCB_LAUNCH(0, true); else MCCB_LAUNCH(0, false); }
#undef MCCB_LAUNCH
}

int fill_expand_args(ExpandArgs& a, const float* y0, uint64_t n, int d, int weighted, const mccb200_drift* drift,
                     float dt, const mccb200_cubature* cub) {
  MCCB_REQUIRE(y0 && drift && cub, "expand: NULL argument");
  MCCB_REQUIRE(d >= 1 && d <= 4096, "expand: d=%d out of range", d);
  MCCB_REQUIRE(cub->k >= 1, "expand: k=%d", cub->k);
  MCCB_REQUIRE((uint64_t)n * (uint64_t)cub->k <= 0x100000000ull, "expand: n*k=%llu exceeds 2^32 child ids",
               (unsigned long long)(n * (uint64_t)cub->k));
  MCCB_REQUIRE(drift->kind >= 0 && drift->kind <= 2, "expand: drift kind %d", drift->kind);
  MCCB_REQUIRE(drift->kind != 1 || drift->f, "expand: drift kind 1 needs f");
  MCCB_REQUIRE(drift->kind != 2 || (drift->mean && drift->inv_var), "expand: drift kind 2 needs mean, inv_var");
  a.y0 = y0; a.n = n; a.d = d; a.ld = d + (weighted ? 1 : 0); a.weighted = weighted;
  a.drift_kind = drift->kind; a.f = drift->f; a.mean = drift->mean; a.inv_var = drift->inv_var;
  a.dt = dt;
  a.n_sub = 1;
  a.k = (uint32_t)cub->k;
  const bool pow2 = (a.k & (a.k - 1u)) == 0;
  a.k_shift = 0xFFFFFFFFu;
  if (pow2) { a.k_shift = 0; while ((1u << a.k_shift) < a.k) ++a.k_shift; }
  a.points = cub->points; a.scale = cub->scale; a.lam = cub->lam;
  if (!a.points) {
    MCCB_REQUIRE(pow2 && a.k >= 2, "expand: Hadamard cubature needs k = power of two >= 2 (k=%u)", a.k);
    MCCB_REQUIRE((uint32_t)d <= a.k / 2 || d == 1, "expand: Hadamard cubature needs d <= k/2 (d=%d, k=%u)", d, a.k);
    a.half_mask = a.k / 2 - 1u;
  } else {
    a.half_mask = 0;
  }
  return MCCB200_OK;
}

}  // namespace mccb

using namespace mccb;

namespace mccb { int run_expand_select(ExpandArgs& a, cudaStream_t st); }

static int run_expand(ExpandArgs& a, cudaStream_t st) { return mccb::run_expand_select(a, st); }

static thread_local int g_expand_residency = 0;
extern "C" void mccb200_set_expand_residency(int ctas_per_sm) {
  g_expand_residency = ctas_per_sm < 0 ? 0 : (ctas_per_sm > 8 ? 8 : ctas_per_sm);
}

int mccb::run_expand_select(ExpandArgs& a, cudaStream_t st) {
  if (a.n_out == 0) return MCCB200_OK;
  int vec = a.weighted ? 1 : pick_vec(a.d, a.ld, a.y0, a.y1);
  if (a.drift_kind == 1 && ((uintptr_t)a.f % 16 != 0)) vec = 1;
  if (a.points && ((uintptr_t)a.points % 16 != 0)) vec = 1;
  if (a.drift_kind == 2 && (((uintptr_t)a.mean | (uintptr_t)a.inv_var) % 16 != 0)) vec = 1;
  a.vpr = a.d / vec;
  MCCB_REQUIRE(a.vpr <= EX_THREADS, "expand: d=%d too wide for the row mapping", a.d);
  a.rows_per_iter = EX_THREADS / a.vpr;
  if (a.draw_keys) {
    int sh = 0;
    while ((1 << sh) < a.vpr) ++sh;
    if (vec != 4 || (1 << sh) != a.vpr || a.vpr < 4 || a.vpr > 32)
      return fail(MCCB200_ERR_UNSUPPORTED, "expand_select_draw: d=%d (vectors per row %d) is outside the fused draw's shapes", a.d, a.vpr);
    a.vpr_shift = sh;
  }
  uint64_t rows_per_cta = (uint64_t)a.rows_per_iter * EX_UNROLL;
  uint64_t want = (a.n_out + rows_per_cta - 1) / rows_per_cta;
  int ctas_per_sm = env_int("MCCB200_EXPAND_CTAS", 8);
  int pad_smem = env_int("MCCB200_EXPAND_SMEM_PAD", 0);
  if (g_expand_residency > 0) {      // one persistent wave of `residency` CTAs per SM, enforced by a shared-memory pad
    ctas_per_sm = g_expand_residency;
    pad_smem = (int)((220u * 1024u / (unsigned)g_expand_residency) & ~1023u);
    if (pad_smem > 99 * 1024 && g_expand_residency > 1) pad_smem = 99 * 1024;
    if (g_expand_residency >= 4) pad_smem = 0;
  }
  int grid = (int)min(want, (uint64_t)sm_count() * ctas_per_sm);
  phase_mark(a.draw_keys ? "expand_select_draw" : a.idx ? "expand_select" : "expand_all", st);
  if (vec == 4) launch_expand<4>(a, grid, st, pad_smem);
  else if (vec == 2) launch_expand<2>(a, grid, st);
  else launch_expand<1>(a, grid, st);
  phase_mark("end", st);
  count_launches(1);
  return check_launch("expand");
}

extern "C" int mccb200_expand_select(const float* y0, uint64_t n, int d, int weighted, const mccb200_drift* drift,
                                     float dt, const mccb200_cubature* cub, const uint32_t* idx,
                                     const uint32_t* perm, uint64_t out_per_part, uint64_t in_per_part,
                                     uint64_t n_out, float* y1, void* stream) {
  ExpandArgs a{};
  int rc = fill_expand_args(a, y0, n, d, weighted, drift, dt, cub);
  if (rc) return rc;
  MCCB_REQUIRE(idx && y1, "expand_select: NULL idx / y1");
  if (perm) {
    MCCB_REQUIRE(out_per_part >= 1 && in_per_part >= 1 && n_out % out_per_part == 0,
                 "expand_select: n_out=%llu not a multiple of out_per_part=%llu", (unsigned long long)n_out,
                 (unsigned long long)out_per_part);
    MCCB_REQUIRE(out_per_part < 0x100000000ull && in_per_part < 0x100000000ull, "expand_select: partition too large");
  }
  a.idx = idx; a.perm = perm; a.out_per_part = (uint32_t)out_per_part; a.in_per_part = (uint32_t)in_per_part;
  a.n_out = n_out; a.y1 = y1;
  return run_expand(a, (cudaStream_t)stream);
}

/* MonteCarloKernel(with_replacement=True) + expand in ONE kernel: output row r is child
 * jr.choice(key_at(t), n * k, (n_out,), replace=True)[r] (32-bit randint) of the expanded set, the id drawn inside the
 * kernel instead of being written to and re-read from HBM.  Shapes: unweighted, d a multiple of 4 with d / 4 a power of
 * two in 4 .. 32 (mccb200_expand_select_draw_supported); idx_out (optional) receives the drawn ids.  workspace:
 * mccb200_expand_select_draw_workspace_bytes().  Reference: mccube/_kernels/random.py:45-66 + _solvers.py:117-142. */
extern "C" size_t mccb200_expand_select_draw_workspace_bytes(void) { return 256; }

extern "C" int mccb200_expand_select_draw_supported(int d, int ld, int weighted) {
  if (weighted || d <= 0 || d % 4 != 0 || ld % 4 != 0) return 0;
  const int vpr = d / 4;
  return (vpr & (vpr - 1)) == 0 && vpr >= 4 && vpr <= 32;
}

extern "C" int mccb200_expand_select_draw(const float* y0, uint64_t n, int d, const mccb200_drift* drift, float dt,
                                          const mccb200_cubature* cub, const mccb200_rng* rng, uint64_t n_out, float* y1,
                                          uint32_t* idx_out, void* workspace, size_t workspace_bytes, void* stream) {
  ExpandArgs a{};
  int rc = fill_expand_args(a, y0, n, d, 0, drift, dt, cub);
  if (rc) return rc;
  MCCB_REQUIRE(rng && y1, "expand_select_draw: NULL rng / y1");
  MCCB_REQUIRE(rng->n_leaves >= 1 && rng->leaf < rng->n_leaves, "expand_select_draw: bad leaf");
  const uint64_t span = n * (uint64_t)cub->k;
  MCCB_REQUIRE(span >= 1 && span < 0xFFFFFFFFull && n_out < 0xFFFFFFFFull, "expand_select_draw: sizes out of the 32-bit id range");
  if (!workspace || workspace_bytes < 256)
    return fail(MCCB200_ERR_WORKSPACE_TOO_SMALL, "expand_select_draw: workspace %zu < 256", workspace_bytes);
  if (n_out == 0) return MCCB200_OK;
  DerivedKeys* dk = (DerivedKeys*)workspace;
  launch_derive_keys(*rng, 0, dk, (cudaStream_t)stream);
  count_launches(1);
  a.idx = nullptr; a.perm = nullptr; a.out_per_part = 1; a.in_per_part = 1;
  a.draw_keys = &dk->randint_hi.x;
  a.draw_count = (uint32_t)n_out; a.draw_span = (uint32_t)span; a.draw_first = 0u;
  a.draw_partitionable = rng->partitionable;
  a.draw_idx_out = idx_out;
  a.n_out = n_out; a.y1 = y1;
  return run_expand(a, (cudaStream_t)stream);
}

/* expand_select whose child ids are `idx_stride` 32-bit words apart: idx_stride = 2 reads the id column of packed
 * (slot, id) pairs, as they arrive from the id exchange of the sharded resample (mccb200_shard_bucket_ids). */
extern "C" int mccb200_expand_select_strided(const float* y0, uint64_t n, int d, const mccb200_drift* drift, float dt,
                                             const mccb200_cubature* cub, const uint32_t* idx, int idx_stride,
                                             uint64_t n_out, float* y1, void* stream) {
  ExpandArgs a{};
  int rc = fill_expand_args(a, y0, n, d, 0, drift, dt, cub);
  if (rc) return rc;
  MCCB_REQUIRE(idx && y1 && idx_stride >= 1, "expand_select_strided: bad argument");
  a.idx = idx; a.idx_stride = (uint32_t)idx_stride; a.perm = nullptr; a.out_per_part = 1; a.in_per_part = 1;
  a.n_out = n_out; a.y1 = y1;
  return run_expand(a, (cudaStream_t)stream);
}

static int scatter_peers_impl(const float* y0, uint64_t n_loc, int d, const mccb200_drift* drift, float dt,
                              const mccb200_cubature* cub, const uint32_t* idx, const uint32_t* order,
                              const uint32_t* counts, int rank, int n_ranks, uint64_t rows_per_rank,
                              float* const* peer_out, void* stream, bool local_ids) {
  ExpandArgs a{};
  int rc = fill_expand_args(a, y0, n_loc, d, 0, drift, dt, cub);
  if (rc) return rc;
  MCCB_REQUIRE(idx && order && counts && peer_out, "expand_scatter_peers: NULL argument");
  MCCB_REQUIRE(n_ranks >= 1 && n_ranks <= 64 && rank >= 0 && rank < n_ranks && rows_per_rank >= 1,
               "expand_scatter_peers: rank %d / %d", rank, n_ranks);
  if (cub->points != nullptr) return fail(MCCB200_ERR_UNSUPPORTED, "expand_scatter_peers: Hadamard cubature only");
  MCCB_REQUIRE(rows_per_rank * (uint64_t)n_ranks * (uint64_t)cub->k <= 0x100000000ull,
               "expand_scatter_peers: n_tot * k exceeds 2^32 child ids");
  a.idx = idx; a.perm = nullptr; a.out_per_part = 1; a.in_per_part = 1;
  a.sc_order = order; a.sc_counts = counts; a.sc_peers = peer_out;
  a.sc_rank = (uint32_t)rank; a.sc_G = (uint32_t)n_ranks;
  a.sc_rows_per_rank = rows_per_rank;
  a.sc_child_offset = local_ids ? 0 : (uint64_t)rank * n_loc * (uint64_t)cub->k;
  a.n_out = rows_per_rank + rows_per_rank / 8 + 1024;   // grid sizing only: the true count is read on the device
  a.y1 = const_cast<float*>(y0);                        // alignment probe only (never written in this form)
  return run_expand(a, (cudaStream_t)stream);
}

extern "C" int mccb200_expand_scatter_peers(const float* y0, uint64_t n_loc, int d, const mccb200_drift* drift, float dt,
                                            const mccb200_cubature* cub, const uint32_t* idx, const uint32_t* order,
                                            const uint32_t* counts, int rank, int n_ranks, uint64_t rows_per_rank,
                                            float* const* peer_out, void* stream) {
  return scatter_peers_impl(y0, n_loc, d, drift, dt, cub, idx, order, counts, rank, n_ranks, rows_per_rank, peer_out,
                            stream, false);
}

/* The same kernel for a list of global output slots: `idx_slots[slot]` holds the LOCAL child id (parent * k +
 * point) this rank contributes to `slot`, `slot_list` the slots it contributes to (any order), and
 * `counts[rank * n_ranks + 0..]` their number (summed; other rows of `counts` must be zero). */
extern "C" int mccb200_expand_scatter_slots(const float* y0, uint64_t n_loc, int d, const mccb200_drift* drift, float dt,
                                            const mccb200_cubature* cub, const uint32_t* idx_slots,
                                            const uint32_t* slot_list, const uint32_t* counts, int rank, int n_ranks,
                                            uint64_t rows_per_rank, float* const* peer_out, void* stream) {
  return scatter_peers_impl(y0, n_loc, d, drift, dt, cub, idx_slots, slot_list, counts, rank, n_ranks, rows_per_rank,
                            peer_out, stream, true);
}

extern "C" int mccb200_expand_gather_peers(float* const* peer_in, int n_ranks, uint64_t rows_per_rank, int d,
                                           const mccb200_drift* drift, float dt, const mccb200_cubature* cub,
                                           const uint32_t* idx, const uint64_t* idx64, uint64_t n_out, float* y1,
                                           void* stream) {
  MCCB_REQUIRE(peer_in && (idx || idx64) && y1 && n_ranks >= 1 && n_ranks <= 64 && rows_per_rank >= 1,
               "expand_gather_peers: bad argument");
  MCCB_REQUIRE(rows_per_rank * (uint64_t)n_ranks <= 0xFFFFFFFFull, "expand_gather_peers: more than 2^32 - 1 parents");
  ExpandArgs a{};
  // y0 is only used for alignment checks here; parents come from peer_in[owner].  With 32-bit ids the
  // global child count must fit 2^32; with idx64 only the parent count has to (checked above).
  int rc = fill_expand_args(a, y1, idx64 ? 1 : rows_per_rank * (uint64_t)n_ranks, d, 0, drift, dt, cub);
  if (rc) return rc;
  a.idx64 = idx64;
  if (cub->points != nullptr) return fail(MCCB200_ERR_UNSUPPORTED, "expand_gather_peers: Hadamard cubature only");
  MCCB_REQUIRE(drift->kind != 1, "expand_gather_peers: an array drift lives with its parents (use the scatter form)");
  a.idx = idx; a.perm = nullptr; a.out_per_part = 1; a.in_per_part = 1;
  a.sc_peers = peer_in; a.sc_order = nullptr; a.sc_G = (uint32_t)n_ranks; a.sc_rows_per_rank = rows_per_rank;
  a.n_out = n_out; a.y1 = y1;
  return run_expand(a, (cudaStream_t)stream);
}

extern "C" int mccb200_expand_select_substeps(const float* y0, uint64_t n, int d, const mccb200_drift* drift,
                                              int n_substeps, const float* dts_host, const float* scales_host,
                                              int k, const uint32_t* idx, uint64_t n_out, float* y1, void* stream) {
  MCCB_REQUIRE(n_substeps >= 1 && n_substeps <= MCCB200_MAX_SUBSTEPS && dts_host && scales_host,
               "expand_select_substeps: n_substeps=%d out of range [1, %d]", n_substeps, MCCB200_MAX_SUBSTEPS);
  mccb200_cubature cub{};
  cub.k = k; cub.points = nullptr; cub.scale = scales_host[0]; cub.lam = nullptr;
  ExpandArgs a{};
  int rc = fill_expand_args(a, y0, n, d, 0, drift, dts_host[0], &cub);
  if (rc) return rc;
  MCCB_REQUIRE(drift->kind != 1, "expand_select_substeps: an array drift is only known at the parents (use the generic path)");
  MCCB_REQUIRE(idx && y1, "expand_select_substeps: NULL idx / y1");
  long double total = (long double)n;
  for (int sb = 0; sb < n_substeps; ++sb) total *= (long double)k;
  MCCB_REQUIRE(total <= 4294967296.0L, "expand_select_substeps: n*k^s exceeds 2^32 child ids");
  a.n_sub = n_substeps;
  for (int sb = 0; sb < n_substeps; ++sb) { a.dt_sub[sb] = dts_host[sb]; a.scale_sub[sb] = scales_host[sb]; }
  a.idx = idx; a.perm = nullptr; a.out_per_part = 1; a.in_per_part = 1;
  a.n_out = n_out; a.y1 = y1;
  return run_expand(a, (cudaStream_t)stream);
}

extern "C" int mccb200_expand_all(const float* y0, uint64_t n, int d, int weighted, const mccb200_drift* drift,
                                  float dt, const mccb200_cubature* cub, float* y1, void* stream) {
  ExpandArgs a{};
  int rc = fill_expand_args(a, y0, n, d, weighted, drift, dt, cub);
  if (rc) return rc;
  MCCB_REQUIRE(y1, "expand_all: NULL y1");
  a.idx = nullptr; a.perm = nullptr; a.out_per_part = 1; a.in_per_part = 1;
  a.n_out = n * (uint64_t)cub->k; a.y1 = y1;
  return run_expand(a, (cudaStream_t)stream);
}

extern "C" int mccb200_gather_rows(const float* in, int ld, const uint32_t* idx, uint64_t n_out, float* out,
                                   void* stream) {
  MCCB_REQUIRE(in && idx && out && ld >= 1, "gather_rows: bad argument");
  if (n_out == 0) return MCCB200_OK;
  int vec = pick_vec(ld, ld, in, out);
  int vpr = ld / vec;
  MCCB_REQUIRE(vpr <= EX_THREADS, "gather_rows: ld=%d too wide", ld);
  int rpi = EX_THREADS / vpr;
  int grid = (int)min((n_out + rpi - 1) / rpi, (uint64_t)sm_count() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec == 4) gather_kernel<4><<<grid, EX_THREADS, 0, st>>>(in, ld, vpr, rpi, 0, 0, idx, n_out, out);
  else if (vec == 2) gather_kernel<2><<<grid, EX_THREADS, 0, st>>>(in, ld, vpr, rpi, 0, 0, idx, n_out, out);
  else gather_kernel<1><<<grid, EX_THREADS, 0, st>>>(in, ld, vpr, rpi, 0, 0, idx, n_out, out);
  count_launches(1);
  return check_launch("gather_rows");
}

/* y_out = y_in with the weight column (the last of ld) divided by its sum: MCCSolver.step's renormalised copy of the
 * recombined state (mccube/_solvers.py:146-155) without a separate clone.  `total` (device, optional): the sum of the
 * weights if a producer already has it (mccb200_expand_select_wsum); NULL = computed here into workspace[0]. */
extern "C" int mccb200_normalised_copy(const float* y_in, uint64_t n, int ld, const double* total, float* y_out,
                                       double* workspace, void* stream) {
  MCCB_REQUIRE(y_in && y_out && y_in != y_out && ld >= 2 && (total || workspace), "normalised_copy: bad argument");
  if (n == 0) return MCCB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int launches = 1;
  if (!total) {
    if (cudaMemsetAsync(workspace, 0, sizeof(double), st) != cudaSuccess)
      return fail(MCCB200_ERR_CUDA, "normalised_copy: memset failed");
    int sgrid = (int)min((n + 2047) / 2048, (uint64_t)sm_count() * 8);
    weight_sum_kernel<<<sgrid < 1 ? 1 : sgrid, 256, 0, st>>>(y_in, n, ld, workspace);
    total = workspace;
    launches += 2;
  }
  const int vec_ok = ((((uintptr_t)y_in) | ((uintptr_t)y_out)) & 15) == 0 && ld >= 4;   // ld >= 4: one weight per 16 bytes at most
  const uint64_t quads = n * (uint64_t)ld / 4;
  int grid = (int)min((quads + 1023) / 1024 + 1, (uint64_t)sm_count() * 8);
  normalised_copy_kernel<<<grid, 256, 0, st>>>(y_in, n, ld, total, y_out, vec_ok);
  count_launches(launches);
  return check_launch("normalised_copy");
}

/* mccb200_expand_select for weighted states that also returns the sum of the child weights it wrote (*wsum, a device
 * double the call zeroes first): the renormalisation that follows needs no pass of its own over the weight column. */
extern "C" int mccb200_expand_select_wsum(const float* y0, uint64_t n, int d, const mccb200_drift* drift, float dt,
                                          const mccb200_cubature* cub, const uint32_t* idx, uint64_t n_out, float* y1,
                                          double* wsum, void* stream) {
  ExpandArgs a{};
  int rc = fill_expand_args(a, y0, n, d, 1, drift, dt, cub);
  if (rc) return rc;
  MCCB_REQUIRE(idx && y1 && wsum, "expand_select_wsum: NULL argument");
  if (cudaMemsetAsync(wsum, 0, sizeof(double), (cudaStream_t)stream) != cudaSuccess)
    return fail(MCCB200_ERR_CUDA, "expand_select_wsum: memset failed");
  count_launches(1);
  a.idx = idx; a.perm = nullptr; a.out_per_part = 1; a.in_per_part = 1;
  a.n_out = n_out; a.y1 = y1; a.wsum = wsum;
  return run_expand(a, (cudaStream_t)stream);
}

extern "C" int mccb200_normalise_weights(float* y, uint64_t n, int ld, double* workspace, void* stream) {
  MCCB_REQUIRE(y && workspace && ld >= 2, "normalise_weights: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(workspace, 0, sizeof(double), st) != cudaSuccess)
    return fail(MCCB200_ERR_CUDA, "normalise_weights: memset failed");
  int grid = (int)min((n + 2047) / 2048, (uint64_t)sm_count() * 8);
  if (grid < 1) grid = 1;
  weight_sum_kernel<<<grid, 256, 0, st>>>(y, n, ld, workspace);
  weight_div_kernel<<<grid, 256, 0, st>>>(y, n, ld, workspace);
  count_launches(3);
  return check_launch("normalise_weights");
}
