// Centred second moment  M2 = sum_r (x_r - c)(x_r - c)^T  of [n, 64] fp32 rows on the 5th-generation tensor cores
// (sm_100a): the d = 64 case of mccb200_moments_m2, i.e. the jnp.cov of /root/reference/README.md:85-86 that feeds
// gaussian_wasserstein_metric (mccube/_metrics.py:39-44) at the headline shape.
//
// The contraction index is the particle row, so a row-major stage of rows IS both operands in MN-major form
// (A[m][k] = B[m][k] = x[k][m]): `tcgen05.mma.kind::tf32` with a_major = b_major = MN reads them from shared memory,
// no transpose.  3xTF32 for fp32-level products: x = hi + lo, and
//     x x^T ~= hi hi^T + lo hi^T + (lo hi^T)^T
// -- the three products of the mma.sync kernel collapse into ONE M128 x N64 x K8 instruction per 8 rows:
// A = [hi ; lo] (128 MN rows), B = hi (64 MN rows) gives D[0:64] = hi^T hi and D[64:128] = lo^T hi in one pass,
// and the epilogue adds the transpose of the second block.  A third of the tensor work of three M64 products.
//
// Shared-memory stage (KS = 64 rows): four planes  hi[:, 0:32], hi[:, 32:64], lo[:, 0:32], lo[:, 32:64],  each KS
// rows of 128 bytes in the one layout MN-major TF32 operands may use, SWIZZLE_128B_BASE32B (CUTLASS
// Layout_MN_SW128_32B_Atom: rows of 128 bytes along MN, 4 rows along K per atom, Swizzle<2,5,2> on byte addresses):
// the 32-byte chunk c of row r of a plane sits at chunk c ^ (r & 3); MN blocks of 32 columns are LBO = KS * 128 bytes
// apart (the planes), 4-row K groups SBO = 512 bytes (rows are contiguous).
//
// (A K-major form -- thread = one column x four rows, so that the four scalar loads are one 16-byte chunk of a
// K-major SWIZZLE_128B tile -- gives the same bits at 1.10 ms: with either operand form the stage period follows the
// operand fetch of the eight instructions per stage; profiles/README.md.)
//
// Persistent CTA per SM, warp-specialised:
//   warps 9-20  converters : six groups of two warps on alternate stages; float4 loads straight from global memory
//                            (a group's next stage is in flight while the other five convert theirs: ~80 KB of loads
//                            in flight per SM), centre, split, st.shared into the swizzled planes, fence.proxy.async, arrive
//   warp 8      MMA issuer : one thread, 8 tcgen05.mma (M128 N64 K8) per stage; tcgen05.commit frees the stage;
//                            every M2T_FLUSH_STAGES stages the accumulator (64 TMEM columns) is handed to the epilogue
//                            and the other one is used (fp32 partial sums stay short: 256 rows)
//   warps 0-7   epilogue   : tcgen05.ld 32x32b -> fp64 registers (thread = accumulator row x 32 columns: warp & 3 is
//                            the TMEM lane quarter it may read, warp >> 2 the column half); at the end one fp64
//                            atomic per entry (+ the mirrored one for the lo^T hi block).  With four warps holding
//                            64 doubles each the sums spilled and the epilogue bounded the kernel (1.49 ms).
#include "common.cuh"

namespace mccb {

constexpr int M2T_D = 64;
constexpr int M2T_KS = 64;                        // rows per stage
constexpr int M2T_STAGES = 6;
constexpr int M2T_GROUPS = 6;                     // converter groups
constexpr int M2T_GROUP_WARPS = 2;
constexpr int M2T_FLUSH_STAGES = 4;               // 256 rows per fp32 partial sum
constexpr int M2T_PLANE_BYTES = M2T_KS * 128;     // 8 KB
constexpr int M2T_STAGE_BYTES = 4 * M2T_PLANE_BYTES;   // 32 KB
constexpr int M2T_MMA_WARP = 8;                   // warps 0-7: epilogue, 8: MMA issuer, 9..: converters
constexpr int M2T_CONV_WARP0 = 9;
constexpr int M2T_THREADS = 32 * (M2T_CONV_WARP0 + M2T_GROUPS * M2T_GROUP_WARPS);
constexpr int M2T_SMEM = M2T_STAGES * M2T_STAGE_BYTES + 1024 + 256;
constexpr int M2T_F4_PER_THREAD = M2T_KS * (M2T_D / 4) / (32 * M2T_GROUP_WARPS);   // 16
constexpr int M2T_ROW_STEP = 32 * M2T_GROUP_WARPS / 16;                             // rows between a thread's float4s

__device__ __forceinline__ uint32_t m2t_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void m2t_bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void m2t_bar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void m2t_bar_wait(uint32_t bar, uint32_t parity) {
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
    if (!ok && (spin & 1023u) == 1023u) {            // a protocol bug must trap, not hang the GPU
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000ll) __trap();
    }
  }
}
__device__ __forceinline__ void m2t_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void m2t_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint64_t m2t_pack(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void m2t_unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t m2t_add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t m2t_sub2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ void m2t_sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// MN-major, SWIZZLE_128B_BASE32B: start address, LBO = distance between 32-column MN blocks (planes), SBO = 4-row K groups
__device__ __forceinline__ uint64_t m2t_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(M2T_PLANE_BYTES >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  d |= (uint64_t)1 << 61;                          // SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ void m2t_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(M2T_THREADS, 1)
moments_m2_tc_kernel(const float* __restrict__ p, uint64_t n, int ld, const float* __restrict__ centre,
                     uint64_t stages_per_cta, double* __restrict__ out, int debug) {
  extern __shared__ uint8_t m2t_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(m2t_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + M2T_STAGES * M2T_STAGE_BYTES);
  const uint32_t bar_full = m2t_smem(bars), bar_empty = m2t_smem(bars + M2T_STAGES);
  const uint32_t bar_afull = m2t_smem(bars + 2 * M2T_STAGES), bar_aempty = m2t_smem(bars + 2 * M2T_STAGES + 2);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * M2T_STAGES + 4);
  const uint32_t smem_base = m2t_smem(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint64_t total_stages = (n + M2T_KS - 1) / M2T_KS;
  const uint64_t s0 = (uint64_t)blockIdx.x * stages_per_cta;
  const uint64_t s1 = min(total_stages, s0 + stages_per_cta);
  const uint32_t my_stages = s0 < s1 ? (uint32_t)(s1 - s0) : 0u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < M2T_STAGES; ++s) {
      m2t_bar_init(bar_full + 8 * s, 32 * M2T_GROUP_WARPS);       // every thread of the converting group arrives
      m2t_bar_init(bar_empty + 8 * s, 1);                         // one tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      m2t_bar_init(bar_afull + 8 * a, 1);
      m2t_bar_init(bar_aempty + 8 * a, 256);                      // the eight epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == M2T_MMA_WARP) {   // two accumulators of 64 fp32 columns (128 lanes)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(m2t_smem(s_tmem)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *s_tmem;
  const uint32_t n_flush = (my_stages + M2T_FLUSH_STAGES - 1) / M2T_FLUSH_STAGES;

  if (warp >= M2T_CONV_WARP0) {
    // ===== converters: group g takes stages g, g + GROUPS, ... of this CTA =====
    const int g = (warp - M2T_CONV_WARP0) / M2T_GROUP_WARPS;
    const int t = threadIdx.x - 32 * (M2T_CONV_WARP0 + g * M2T_GROUP_WARPS);   // index inside the group
    // float4 f of a stage: row = f / 16, quad = f % 16; thread t takes f = t + (group size) * i  ->  quad = t % 16 for all i
    const int quad = t & 15;
    const float4 c4 = __ldg(reinterpret_cast<const float4*>(centre) + quad);
    const int plane = quad >> 3, chunk = quad & 7;
    const uint64_t nc01 = m2t_pack(-c4.x, -c4.y), nc23 = m2t_pack(-c4.z, -c4.w);
    float4 cur[M2T_F4_PER_THREAD];
    auto load = [&](uint32_t st, float4 (&v)[M2T_F4_PER_THREAD]) {
      const uint64_t row0 = (s0 + st) * M2T_KS;
#pragma unroll
      for (int i = 0; i < M2T_F4_PER_THREAD; ++i) {
        const uint64_t r = row0 + (uint64_t)((t >> 4) + M2T_ROW_STEP * i);
        v[i] = (r < n && debug != 3) ? __ldcs(reinterpret_cast<const float4*>(p + r * (uint64_t)ld) + quad)
                                     : make_float4(c4.x, c4.y, c4.z, c4.w);      // centred to zero: contributes nothing
      }
    };
    if ((uint32_t)g < my_stages) load((uint32_t)g, cur);
    for (uint32_t st = (uint32_t)g; st < my_stages; st += M2T_GROUPS) {
      const uint32_t slot = st % M2T_STAGES, use = st / M2T_STAGES;
      m2t_bar_wait(bar_empty + 8 * slot, (use & 1u) ^ 1u);
      const uint32_t sb = smem_base + slot * M2T_STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < M2T_F4_PER_THREAD; ++i) {
        const int r = (t >> 4) + M2T_ROW_STEP * i;
        // x = row - centre; the tensor core reads only the top 19 bits of an operand (sign, exponent, 10 mantissa
        // bits), so x itself serves as the TF32 high part and lo = x - (x with the low 13 bits cleared), exact in fp32
        const uint64_t x01 = m2t_add2(m2t_pack(cur[i].x, cur[i].y), nc01), x23 = m2t_add2(m2t_pack(cur[i].z, cur[i].w), nc23);
        const uint64_t lo01 = m2t_sub2(x01, x01 & 0xFFFFE000FFFFE000ull), lo23 = m2t_sub2(x23, x23 & 0xFFFFE000FFFFE000ull);
        float hi[4], lo[4];
        m2t_unpack(x01, hi[0], hi[1]); m2t_unpack(x23, hi[2], hi[3]);
        m2t_unpack(lo01, lo[0], lo[1]); m2t_unpack(lo23, lo[2], lo[3]);
        const uint32_t off = (uint32_t)plane * M2T_PLANE_BYTES + (uint32_t)r * 128u +
                             (uint32_t)((((chunk >> 1) ^ (r & 3)) << 5) | ((chunk & 1) << 4));
        m2t_sts128(sb + off, hi[0], hi[1], hi[2], hi[3]);
        m2t_sts128(sb + 2 * M2T_PLANE_BYTES + off, lo[0], lo[1], lo[2], lo[3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
      m2t_bar_arrive(bar_full + 8 * slot);
      if (st + M2T_GROUPS < my_stages) load(st + M2T_GROUPS, cur);   // in flight while the other groups convert
    }
  } else if (warp == M2T_MMA_WARP) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // c_format F32, a/b TF32, a_major = b_major = MN, N = 64, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(M2T_D >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      for (uint32_t st = 0; st < my_stages; ++st) {
        const uint32_t slot = st % M2T_STAGES, use = st / M2T_STAGES;
        const uint32_t fl = st / M2T_FLUSH_STAGES, acc = fl & 1u, acc_use = fl >> 1;
        if (st % M2T_FLUSH_STAGES == 0) {                          // first stage of a partial sum: the accumulator must be drained
          m2t_bar_wait(bar_aempty + 8 * acc, (acc_use & 1u) ^ 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        m2t_bar_wait(bar_full + 8 * slot, use & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sb = smem_base + slot * M2T_STAGE_BYTES;
#pragma unroll
        for (int kg = 0; kg < M2T_KS / 8; ++kg) {
          if (debug == 2) break;                                   // timing experiment: no tensor work
          const uint64_t desc = m2t_desc(sb + kg * 1024);
          m2t_mma(tmem_base + acc * M2T_D, desc, desc, idesc, (st % M2T_FLUSH_STAGES != 0 || kg != 0) ? 1u : 0u);
        }
        m2t_commit(bar_empty + 8 * slot);                          // the stage may be refilled once these MMAs retire
        if (st % M2T_FLUSH_STAGES == M2T_FLUSH_STAGES - 1 || st + 1 == my_stages) m2t_commit(bar_afull + 8 * acc);
      }
    }
  } else {
    // ===== epilogue: thread = accumulator row (TMEM lane) x 32 columns, fp64 running sums =====
    const int q = warp & 3, ch = warp >> 2;                        // lane quarter this warp may read, column half
    const int row = 32 * q + lane;
    double sum[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) sum[j] = 0.0;
    for (uint32_t fl = 0; fl < n_flush; ++fl) {
      const uint32_t acc = fl & 1u, acc_use = fl >> 1;
      m2t_bar_wait(bar_afull + 8 * acc, acc_use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[16];
        m2t_tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + acc * M2T_D + 32 * ch + 16 * h, v);
        if (debug == 4 && fl != 0) continue;                       // timing experiment: no fp64 accumulation
#pragma unroll
        for (int j = 0; j < 16; ++j) sum[16 * h + j] += (debug == 1) ? 1.0 : (double)__uint_as_float(v[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      m2t_bar_arrive(bar_aempty + 8 * acc);
    }
    if (my_stages > 0) {
      if (row < M2T_D) {                                           // hi^T hi
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(out + (size_t)row * M2T_D + 32 * ch + j, sum[j]);
      } else {                                                     // L = lo^T hi: M2 += L + L^T
        const int i = row - M2T_D;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          atomicAdd(out + (size_t)i * M2T_D + 32 * ch + j, sum[j]);
          atomicAdd(out + (size_t)(32 * ch + j) * M2T_D + i, sum[j]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == M2T_MMA_WARP) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Second form (MCCB200_M2_TC=2): the A operand [hi ; lo] lives in TENSOR MEMORY (tcgen05.mma with A from TMEM), only
// B = hi is staged in shared memory (K-major SWIZZLE_128B).  A converter thread owns one column m of the rows: its 64
// scalar loads are each part of a 128-byte coalesced warp request, its 64 values are row m of A for the stage's 64 k
// values, written with tcgen05.st.32x32b (warp w may write TMEM lanes 32 (w % 4) ..: warps 0,1 of a group hold hi of
// columns 0-31 / 32-63, warps 2,3 the lo parts), and the hi warps also store their values as 16-byte chunks of the
// K-major B tile.  Shared-memory traffic per 64-row stage: 16 KB written + 16 KB read by the eight instructions,
// against 32 + 48 KB in the first form.  TMEM: 2 x 64 accumulator columns + 6 stages x 64 columns of A = 512.
constexpr int M2S_STAGES = 6;
constexpr int M2S_GROUPS = 3;
constexpr int M2S_B_BYTES = 2 * 64 * 128;                              // two K blocks of 64 rows x 128 bytes
constexpr int M2S_CONV_WARP0 = 8;                                      // warps 0-7 epilogue, 8..19 converters, 20 MMA
constexpr int M2S_MMA_WARP = M2S_CONV_WARP0 + 4 * M2S_GROUPS;
constexpr int M2S_THREADS = 32 * (M2S_MMA_WARP + 1);
constexpr int M2S_SMEM = M2S_STAGES * M2S_B_BYTES + 1024 + 256;
constexpr uint32_t M2S_A_COL0 = 2 * M2T_D;                             // first TMEM column of the A stages

__device__ __forceinline__ uint64_t m2s_desc(uint32_t saddr) {         // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void m2s_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void m2s_tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

__global__ void __launch_bounds__(M2S_THREADS, 1)
moments_m2_ts_kernel(const float* __restrict__ p, uint64_t n, int ld, const float* __restrict__ centre,
                     uint64_t stages_per_cta, double* __restrict__ out) {
  extern __shared__ uint8_t m2t_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(m2t_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + M2S_STAGES * M2S_B_BYTES);
  const uint32_t bar_full = m2t_smem(bars), bar_empty = m2t_smem(bars + M2S_STAGES);
  const uint32_t bar_afull = m2t_smem(bars + 2 * M2S_STAGES), bar_aempty = m2t_smem(bars + 2 * M2S_STAGES + 2);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * M2S_STAGES + 4);
  const uint32_t smem_base = m2t_smem(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint64_t total_stages = (n + M2T_KS - 1) / M2T_KS;
  const uint64_t s0 = (uint64_t)blockIdx.x * stages_per_cta;
  const uint64_t s1 = min(total_stages, s0 + stages_per_cta);
  const uint32_t my_stages = s0 < s1 ? (uint32_t)(s1 - s0) : 0u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < M2S_STAGES; ++s) {
      m2t_bar_init(bar_full + 8 * s, 128);                        // the four warps of the converting group
      m2t_bar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      m2t_bar_init(bar_afull + 8 * a, 1);
      m2t_bar_init(bar_aempty + 8 * a, 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == M2S_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(m2t_smem(s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *s_tmem;
  const uint32_t n_flush = (my_stages + M2T_FLUSH_STAGES - 1) / M2T_FLUSH_STAGES;

  if (warp >= M2S_CONV_WARP0 && warp < M2S_MMA_WARP) {
    // ===== converters =====
    const int g = (warp - M2S_CONV_WARP0) >> 2, q = warp & 3;
    const int m = 32 * (q & 1) + lane;                             // column of x = row of A (hi) / row 64 + m (lo)
    const bool is_lo = q >= 2;
    const float cm = __ldg(centre + m);
    const uint64_t ncm2 = m2t_pack(-cm, -cm);
    float cur[M2T_KS];
    auto load = [&](uint32_t st) {
      const uint64_t row0 = (s0 + st) * M2T_KS;
      const float* src = p + row0 * (uint64_t)ld + m;
      if (row0 + M2T_KS <= n) {
#pragma unroll
        for (int r = 0; r < M2T_KS; ++r) cur[r] = __ldcs(src + (size_t)r * ld);
      } else {
#pragma unroll
        for (int r = 0; r < M2T_KS; ++r) cur[r] = row0 + r < n ? __ldcs(src + (size_t)r * ld) : cm;
      }
    };
    if ((uint32_t)g < my_stages) load((uint32_t)g);
    for (uint32_t st = (uint32_t)g; st < my_stages; st += M2S_GROUPS) {
      const uint32_t slot = st % M2S_STAGES, use = st / M2S_STAGES;
      m2t_bar_wait(bar_empty + 8 * slot, (use & 1u) ^ 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sb = smem_base + slot * M2S_B_BYTES + (uint32_t)m * 128u;
      const uint32_t ta = tmem_base + ((uint32_t)(32 * q) << 16) + M2S_A_COL0 + slot * M2T_KS;
#pragma unroll
      for (int h = 0; h < 4; ++h) {                                // 16 k values = four 16-byte chunks per TMEM store
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = 4 * h + i, r = 4 * c;                      // chunk c of the stage: rows 4c .. 4c + 3
          const uint64_t x01 = m2t_add2(m2t_pack(cur[r], cur[r + 1]), ncm2), x23 = m2t_add2(m2t_pack(cur[r + 2], cur[r + 3]), ncm2);
          float a0, a1, a2, a3;
          if (is_lo) {
            m2t_unpack(m2t_sub2(x01, x01 & 0xFFFFE000FFFFE000ull), a0, a1);
            m2t_unpack(m2t_sub2(x23, x23 & 0xFFFFE000FFFFE000ull), a2, a3);
          } else {
            m2t_unpack(x01, a0, a1);
            m2t_unpack(x23, a2, a3);
            m2t_sts128(sb + (uint32_t)(c >> 3) * 8192u + (uint32_t)(((c & 7) ^ (m & 7)) << 4), a0, a1, a2, a3);   // B: chunk c & 7 of K block c >> 3
          }
          v[4 * i] = __float_as_uint(a0); v[4 * i + 1] = __float_as_uint(a1);
          v[4 * i + 2] = __float_as_uint(a2); v[4 * i + 3] = __float_as_uint(a3);
        }
        m2s_tmem_st16(ta + 16 * h, v);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      m2t_bar_arrive(bar_full + 8 * slot);
      if (st + M2S_GROUPS < my_stages) load(st + M2S_GROUPS);
    }
  } else if (warp == M2S_MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(M2T_D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (uint32_t st = 0; st < my_stages; ++st) {
        const uint32_t slot = st % M2S_STAGES, use = st / M2S_STAGES;
        const uint32_t fl = st / M2T_FLUSH_STAGES, acc = fl & 1u, acc_use = fl >> 1;
        if (st % M2T_FLUSH_STAGES == 0) {
          m2t_bar_wait(bar_aempty + 8 * acc, (acc_use & 1u) ^ 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        m2t_bar_wait(bar_full + 8 * slot, use & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sb = smem_base + slot * M2S_B_BYTES;
        const uint32_t ta = tmem_base + M2S_A_COL0 + slot * M2T_KS;
#pragma unroll
        for (int ks = 0; ks < M2T_KS / 8; ++ks)
          m2s_mma_ts(tmem_base + acc * M2T_D, ta + 8 * ks, m2s_desc(sb + (ks >> 2) * 8192 + (ks & 3) * 32), idesc,
                     (st % M2T_FLUSH_STAGES != 0 || ks != 0) ? 1u : 0u);
        m2t_commit(bar_empty + 8 * slot);
        if (st % M2T_FLUSH_STAGES == M2T_FLUSH_STAGES - 1 || st + 1 == my_stages) m2t_commit(bar_afull + 8 * acc);
      }
    }
  } else if (warp < 8) {
    const int q = warp & 3, ch = warp >> 2;
    const int row = 32 * q + lane;
    double sum[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) sum[j] = 0.0;
    for (uint32_t fl = 0; fl < n_flush; ++fl) {
      const uint32_t acc = fl & 1u, acc_use = fl >> 1;
      m2t_bar_wait(bar_afull + 8 * acc, acc_use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[16];
        m2t_tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + acc * M2T_D + 32 * ch + 16 * h, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) sum[16 * h + j] += (double)__uint_as_float(v[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      m2t_bar_arrive(bar_aempty + 8 * acc);
    }
    if (my_stages > 0) {
      if (row < M2T_D) {
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(out + (size_t)row * M2T_D + 32 * ch + j, sum[j]);
      } else {
        const int i = row - M2T_D;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          atomicAdd(out + (size_t)i * M2T_D + 32 * ch + j, sum[j]);
          atomicAdd(out + (size_t)(32 * ch + j) * M2T_D + i, sum[j]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == M2S_MMA_WARP) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// d = 64, unweighted, rows of whole float4s: the tcgen05 kernel applies (mccb200_moments_m2 falls back otherwise)
bool moments_m2_tc_applicable(const float* p, int d, int ld, const float* w, const float* centre) {
  return env_int("MCCB200_M2_TC", 1) != 0 && d == M2T_D && w == nullptr && ld % 4 == 0 &&
         ((((uintptr_t)p) | ((uintptr_t)centre)) & 15) == 0;
}

int moments_m2_tc(const float* p, uint64_t n, int ld, const float* centre, double* m2_out, cudaStream_t st) {
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_dev != dev) {
    if (cudaFuncSetAttribute(moments_m2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, M2T_SMEM) != cudaSuccess)
      return fail(MCCB200_ERR_CUDA, "moments_m2: cannot reserve %d bytes of shared memory", M2T_SMEM);
    attr_dev = dev;
  }
  const uint64_t total_stages = (n + M2T_KS - 1) / M2T_KS;
  uint64_t ctas = min((uint64_t)sm_count(), total_stages);
  uint64_t per = (total_stages + ctas - 1) / ctas;
  per = (per + M2T_FLUSH_STAGES - 1) / M2T_FLUSH_STAGES * M2T_FLUSH_STAGES;
  ctas = (total_stages + per - 1) / per;
  phase_mark("moments_m2.tcgen05", st);
  if (env_int("MCCB200_M2_TC", 1) == 2) {                          // A operand in tensor memory
    static thread_local int attr_dev2 = -1;
    if (attr_dev2 != dev) {
      if (cudaFuncSetAttribute(moments_m2_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, M2S_SMEM) != cudaSuccess)
        return fail(MCCB200_ERR_CUDA, "moments_m2: cannot reserve %d bytes of shared memory", M2S_SMEM);
      attr_dev2 = dev;
    }
    moments_m2_ts_kernel<<<(unsigned)ctas, M2S_THREADS, M2S_SMEM, st>>>(p, n, ld, centre, per, m2_out);
    phase_mark("end", st);
    count_launches(1);
    return check_launch("moments_m2(tcgen05, A in TMEM)");
  }
  moments_m2_tc_kernel<<<(unsigned)ctas, M2T_THREADS, M2T_SMEM, st>>>(p, n, ld, centre, per, m2_out, env_int("MCCB200_M2_TC_DEBUG", 0));
  phase_mark("end", st);
  count_launches(1);
  return check_launch("moments_m2(tcgen05)");
}

}  // namespace mccb
