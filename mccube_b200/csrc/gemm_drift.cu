// Full-covariance Gaussian drift on the 5th-generation tensor cores (sm_100a):
//     F = -(Y - mean) @ prec            Y [n, d] fp32, prec = cov^-1 [d, d] fp32
// the drift of BASELINE config 4 (n = 2^20, d = 512).  In the reference this is
// vmap(grad(multivariate_normal.logpdf)) evaluated on the parents (README.md:72-76 via
// mccube/_term.py:58-63): two batched triangular solves, i.e. a dense n x d x d contraction.
//
// Precision recipe: "3xTF32" -- both operands are split into a TF32-representable high part and
// the fp32 remainder, and three tcgen05.mma.kind::tf32 products are accumulated in fp32 in TMEM:
//     A B  ~=  A_hi B_hi + A_lo B_hi + A_hi B_lo          (A_lo B_lo ~ 2^-22 |A||B| is dropped)
// which keeps the result within ~1e-6 of the fp32/fp64 value (one-pass TF32 would be ~1e-3; the
// parity bar is 1e-5).  prec is split once (mccb200_gaussian_full_drift_tc_prepare), the
// particle tile is centred and split in shared memory by converter warps.
//
// Kernel anatomy (persistent, one CTA per SM, clusters of two CTAs = one CTA-pair UMMA, 14 warps,
// warp-specialised).  A pair tile is 256 particles (128 per CTA) x NT <= 256 output columns; each
// CTA holds its particles and HALF of every prec tile (the tensor cores of the pair read both halves);
// the k-step is 32 fp32 (one 128-byte swizzle row), 64 KB per stage, 3 stages:
//   warp 0      TMA producer : Y tile [128 x 32] + this CTA's half of the prec_hi / prec_lo tiles
//                              [NT/2 x 32] per k-step, 128-byte swizzle, mbarrier complete_tx
//   warps 2-9   converter    : two groups of 4 warps on alternate stages: a = y - mean; hi = rn_tf32(a);
//                              lo = rn_tf32(a - hi), written back to the (swizzled) stage buffers;
//                              fence.proxy.async; group barrier; ONE remote arrive on the leader's barrier
//   warp 1      MMA issuer   : one thread of the LEADER CTA issues 12 tcgen05.mma.cta_group::2
//                              (M256 x N256 x K8, a_negate) per k-step; tcgen05.commit ... multicast frees
//                              the stage in both CTAs and publishes the accumulators
//   warps 10-13 epilogue     : tcgen05.ld 16x256b.x8 -> 8 rows x 32-byte sectors of F straight from
//                              registers; the TMEM accumulator is double-buffered (2 x 256 columns), so
//                              the epilogue of a tile overlaps the main loop of the next
// What bounds it (knock-out runs, clock64 stage tracer, eleven variants: profiles/README.md): 3 * 2 n d^2
// flop on the TF32 pipe would take 1536 cycles per k-step; the stage period follows the shared-memory
// bytes per stage (operand reads of the three products 96 KB + converter 50 KB + TMA fills 48 KB at
// 128 B/cycle).  Measured 2.30 ms at n = 2^20, d = 512 = 718 TFLOP/s on the pipe, 74 % tensor-pipe active.
// TMA moves a box row by row (~2.5-4 cycles per row): 64- and 32-byte rows starved the ring.
// Accuracy: the tensor core truncates (round-toward-zero) at every accumulation, a bias of
// ~2^-24 per K = 8 instruction: measured max error 7.6e-6 (d = 512) / 1.3e-5 (d = 1024) of max|F|.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace mccb {

constexpr int TC_BM = 128;            // rows per accumulator (UMMA M)
constexpr int TC_RB = 1;              // accumulators (row blocks) per CTA tile
constexpr int TC_ACC = 2;             // TMEM accumulator stages (epilogue of tile i overlaps the main loop of tile i+1)
constexpr int TC_BK = 32;             // fp32 per k-step = one swizzle row (8 / 16 / 32 floats <-> 32 / 64 / 128-byte swizzle)
constexpr int TC_UK = 8;              // K of one tcgen05.mma.kind::tf32
constexpr int TC_STAGES = 3;          // 3 x 64 KB
constexpr int TC_MAX_NT = 256;        // UMMA N
constexpr int TC_MAX_D = 1024;
constexpr int TC_CONV_WARPS = 4, TC_EPI_WARPS = 4;
constexpr int TC_CONV_GROUPS = 2;     // converter groups taking alternate stages: the fence / barrier / remote-arrive tail of one
                                      // stage overlaps the conversion of the next
constexpr int TC_THREADS = 32 * (2 + TC_CONV_GROUPS * TC_CONV_WARPS + TC_EPI_WARPS);   // producer, MMA, converters, epilogue
constexpr int TC_CLUSTER = 2;         // CTA pair (cta_group::2): one UMMA spans both SMs, each holds half of every prec tile
constexpr uint32_t TC_ROW_BYTES = TC_BK * 4;                           // 128: TMA moves a box row by row at ~2.5-4 cycles per row,
                                                                       // so short rows starve the ring (measured: 64-byte rows 15-25 B/cycle/SM)
constexpr int TC_CHUNKS = TC_ROW_BYTES / 16;                           // 16-byte chunks per row
constexpr int TC_SWZ_SHIFT = TC_ROW_BYTES == 128 ? 0 : TC_ROW_BYTES == 64 ? 1 : 2;   // row bits that feed the XOR
constexpr uint64_t TC_LAYOUT_TYPE = TC_ROW_BYTES == 128 ? 2 : TC_ROW_BYTES == 64 ? 4 : 6;   // UMMA SWIZZLE_128B/64B/32B
constexpr uint32_t TC_A_BYTES = TC_RB * TC_BM * TC_ROW_BYTES;          // 16 KB
constexpr uint32_t TC_B_BYTES = TC_MAX_NT / 2 * TC_ROW_BYTES;          // 16 KB: each CTA of the pair holds half of the N rows
constexpr uint32_t TC_OFF_AHI = 0, TC_OFF_ALO = TC_A_BYTES, TC_OFF_BHI = 2 * TC_A_BYTES,
                   TC_OFF_BLO = 2 * TC_A_BYTES + TC_B_BYTES;
constexpr uint32_t TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;   // 64 KB
constexpr uint32_t TC_EPI_BYTES = 0;                                   // the epilogue does not stage through shared memory
constexpr uint32_t TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + TC_MAX_D * 4 + 256 + TC_EPI_BYTES + 1024;  // + mean, barriers, align

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Spins on the phase parity; traps instead of hanging the GPU if a phase is never completed.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
    if (!ok && (spin & 1023u) == 1023u) {                 // ~2 s at 2 GHz: a protocol bug, not a slow tile
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000ll) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
// the same box written to the same shared-memory offset of every CTA in cta_mask (each CTA's own
// mbarrier at that offset receives the complete_tx)
__device__ __forceinline__ void tma_load_2d_mcast(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y,
                                                  uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y), "h"(cta_mask)
      : "memory");
}
// pull a tile into L2 only (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int x, int y) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta_rank) {   // same offset in another CTA of the cluster
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// wait that also acquires what remote CTAs released before arriving
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at this offset in every CTA of cta_mask once the pair's MMAs issued so far are done
__device__ __forceinline__ void tc_commit_mcast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::tf32, issued by one thread
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand tile stored as rows of TC_ROW_BYTES with the matching TMA swizzle (16-byte chunk
// c of row r sits at chunk c ^ ((r >> TC_SWZ_SHIFT) & (TC_CHUNKS - 1)): address bits [4,7) ^= bits
// [7,10)); 8-row groups are 8 * TC_ROW_BYTES apart (SBO), descriptor version 1 (sm_100).  Advancing
// along K inside the swizzle row = adding the byte offset to the start address.
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);      // start address, 16-byte units
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(8 * TC_ROW_BYTES >> 4) << 32;  // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= TC_LAYOUT_TYPE << 61;                     // swizzle mode
  return d;
}
// 16 TMEM lanes x 64 columns: thread t holds rows t/4 (v[4i], v[4i+1]) and t/4 + 8 (v[4i+2], v[4i+3]),
// columns 8i + 2(t%4), +1  -- the accumulator-fragment layout: a warp-wide 8-byte store then writes
// 8 rows x one full 32-byte sector
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// round-to-nearest TF32 (10 explicit mantissa bits) of an fp32 value, as fp32 bits
__device__ __forceinline__ float rn_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

struct TcDriftParams {
  uint64_t n;
  int d, nt, n_tiles, k_tiles;
  uint32_t n_m_tiles;
  long long* trace;   // debug: [256 stages][5] clock64 stamps of CTA 0 (producer issue, full seen, converted, MMA start, MMA issued)
  int debug;   // MCCB200_DEBUG_GEMM (timing experiments only, results are wrong): 1 = converter idle, 2 = one product
  const float* mean;
  float* f;
};

__global__ void __cluster_dims__(TC_CLUSTER, 1, 1) __launch_bounds__(TC_THREADS, 1)
gaussian_full_drift_tc_kernel(const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_bhi,
                              const __grid_constant__ CUtensorMap map_blo, const TcDriftParams p) {
  extern __shared__ uint8_t smem_raw[];
  // swizzled tiles want (at least) 512-byte alignment; both CTAs of the pair use identical offsets
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* s_mean = reinterpret_cast<float*>(smem + TC_STAGES * TC_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES + TC_MAX_D * 4);
  // barrier ring: full[s]   (this CTA's TMA landed)                        local, 1 arrival + tx bytes
  //               conv[s]   (both CTAs' particle tiles split)              LEADER's, 1 arrival per CTA
  //               empty[s]  (the pair's MMAs consumed the stage)           each CTA, 1 multicast commit
  //               tmem_full (accumulators ready)                           each CTA, 1 multicast commit
  //               tmem_empty[r] (accumulator r drained in both CTAs)       LEADER's, 1 arrival per CTA
  const uint32_t bar_full = smem_u32(bars), bar_conv = smem_u32(bars + TC_STAGES), bar_empty = smem_u32(bars + 2 * TC_STAGES);
  const uint32_t bar_tfull = smem_u32(bars + 3 * TC_STAGES), bar_tempty = smem_u32(bars + 3 * TC_STAGES + TC_ACC);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 3 * TC_STAGES + 2 * TC_ACC);
  const uint32_t smem_base = smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const uint16_t cta_all = (uint16_t)((1u << TC_CLUSTER) - 1u);
  for (int i = threadIdx.x; i < p.d; i += TC_THREADS) s_mean[i] = p.mean[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, TC_CLUSTER);            // one arrival per CTA (its converter warps sync first)
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < TC_ACC; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, TC_CLUSTER);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: TC_RB accumulators of up to 256 fp32 columns, allocated in both CTAs of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();            // barrier inits visible cluster-wide before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  // The pair walks the same sequence of (pair tile, n-tile, k-step); a CTA whose 256 rows lie past the
  // end recomputes the last tile and stores nothing.
  const uint32_t n_pairs = (p.n_m_tiles + 1u) / 2u, pairs_in_grid = gridDim.x / 2u;
  const uint32_t n_iter = (n_pairs + pairs_in_grid - 1u) / pairs_in_grid;
  const uint32_t pair0 = blockIdx.x / 2u;
  const uint32_t b_rows = (uint32_t)p.nt / TC_CLUSTER;     // this CTA's rows of every prec tile

  const uint32_t stage_tx = TC_A_BYTES + 2u * b_rows * TC_ROW_BYTES;
  if (warp == 0) {
    // ===== TMA producer (each CTA: its 256 particles, its half of the prec tile) =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, g = 0;
      for (uint32_t it = 0; it < n_iter; ++it) {
        const uint32_t mt = min((pair0 + it * pairs_in_grid) * 2u + cta_rank, p.n_m_tiles - 1u);
        for (int nt = 0; nt < p.n_tiles; ++nt) {
          for (int kt = 0; kt < p.k_tiles; ++kt) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
            if (p.trace && blockIdx.x == 0 && g < 256) p.trace[g * 5 + 0] = clock64();
            ++g;
            const uint32_t sb = smem_base + stage * TC_STAGE_BYTES;
            mbar_expect_tx(bar_full + 8 * stage, stage_tx);
            tma_load_2d(sb + TC_OFF_AHI, &map_y, bar_full + 8 * stage, kt * TC_BK, (int)(mt * (TC_RB * TC_BM)));
            tma_load_2d(sb + TC_OFF_BHI, &map_bhi, bar_full + 8 * stage, kt * TC_BK, nt * p.nt + (int)(cta_rank * b_rows));
            tma_load_2d(sb + TC_OFF_BLO, &map_blo, bar_full + 8 * stage, kt * TC_BK, nt * p.nt + (int)(cta_rank * b_rows));
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread of the LEADER CTA drives both SMs' tensor cores =====
    if (lane == 0 && cta_rank == 0) {
      // instruction descriptor: D fp32, A/B tf32, both K-major, A negated, N = nt, M = 256 (128 rows per CTA)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 13) | ((uint32_t)(p.nt >> 3) << 17) |
                             ((uint32_t)(TC_CLUSTER * TC_BM >> 4) << 24);
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0, g = 0;
      for (uint32_t it = 0; it < n_iter; ++it) {
        for (int nt = 0; nt < p.n_tiles; ++nt) {
          for (int kt = 0; kt < p.k_tiles; ++kt) {
            // conv[stage] completes only after every converter thread of BOTH CTAs has seen its
            // full[stage] (all TMA bytes landed) and fenced its own writes
            mbar_wait_cluster(bar_conv + 8 * stage, phase);
            tc_fence_after();
            if (p.trace && blockIdx.x == 0 && g < 256) p.trace[g * 5 + 3] = clock64();
            const uint32_t sb = smem_base + stage * TC_STAGE_BYTES;
            const uint64_t b_hi = tc_smem_desc(sb + TC_OFF_BHI), b_lo = tc_smem_desc(sb + TC_OFF_BLO);
#pragma unroll
            for (int r = 0; r < TC_RB; ++r) {
              if (kt == 0 && r == 0) {                        // accumulator stage drained by both epilogues?
                mbar_wait_cluster(bar_tempty + 8 * as, aphase ^ 1u);
                tc_fence_after();
              }
              const uint32_t d_tmem = tmem_base + (as * TC_RB + (uint32_t)r) * (uint32_t)p.nt;
              const uint64_t a_hi = tc_smem_desc(sb + TC_OFF_AHI + r * TC_BM * TC_ROW_BYTES);
              const uint64_t a_lo = tc_smem_desc(sb + TC_OFF_ALO + r * TC_BM * TC_ROW_BYTES);
              if (!(p.debug & 2)) {
#pragma unroll
                for (int ks = 0; ks < TC_BK / TC_UK; ++ks) {   // descriptor start address advances in 16-byte units
                  const uint64_t ko = (uint64_t)(ks * TC_UK * 4 >> 4);
                  tc_mma_tf32(d_tmem, a_lo + ko, b_hi + ko, idesc, (kt | ks) != 0);
                }
#pragma unroll
                for (int ks = 0; ks < TC_BK / TC_UK; ++ks) {
                  const uint64_t ko = (uint64_t)(ks * TC_UK * 4 >> 4);
                  tc_mma_tf32(d_tmem, a_hi + ko, b_lo + ko, idesc, 1u);
                }
              }
#pragma unroll
              for (int ks = 0; ks < TC_BK / TC_UK; ++ks) {
                const uint64_t ko = (uint64_t)(ks * TC_UK * 4 >> 4);
                tc_mma_tf32(d_tmem, a_hi + ko, b_hi + ko, idesc, (p.debug & 2) ? (uint32_t)((kt | ks) != 0) : 1u);
              }
            }
            tc_commit_mcast(bar_empty + 8 * stage, cta_all);  // both producers: the pair is done with the stage
            if (kt == p.k_tiles - 1) tc_commit_mcast(bar_tfull + 8 * as, cta_all);
            if (p.trace && blockIdx.x == 0 && g < 256) p.trace[g * 5 + 4] = clock64();
            ++g;
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
          }
          if (++as == TC_ACC) { as = 0; aphase ^= 1u; }
        }
      }
    }
  } else if (warp < 2 + TC_CONV_GROUPS * TC_CONV_WARPS) {
    // ===== converter: centre and split the particle tile in place (swizzle aware) =====
    const int group = (warp - 2) / TC_CONV_WARPS;
    const int tc = threadIdx.x - 64 - group * 32 * TC_CONV_WARPS;   // 0 .. 32 * TC_CONV_WARPS - 1 inside the group
    const uint32_t leader_conv = mapa_u32(bar_conv, 0u);
    uint32_t stage = 0, phase = 0, g = 0;
    for (uint32_t it = 0; it < n_iter; ++it) {
      for (int nt = 0; nt < p.n_tiles; ++nt) {
        for (int kt = 0; kt < p.k_tiles; ++kt) {
          if ((int)(g % TC_CONV_GROUPS) != group) {        // the other group's stage
            ++g;
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
            continue;
          }
          mbar_wait(bar_full + 8 * stage, phase);
          if (p.trace && blockIdx.x == 0 && tc == 0 && g < 256) p.trace[g * 5 + 1] = clock64();
          uint8_t* sa = smem + stage * TC_STAGE_BYTES;
          if (!(p.debug & 1)) {
            constexpr int NCH = (int)(TC_A_BYTES / 16) / (32 * TC_CONV_WARPS);   // 16-byte chunks per thread
            float4 v[NCH];
            // A thread's chunks are 32 * TC_CONV_WARPS apart = a multiple of 8 rows, so they all sit in the
            // same logical column group of the swizzled tile: one mean load per stage (shared memory
            // bandwidth is what bounds this kernel).
            static_assert((32 * TC_CONV_WARPS / TC_CHUNKS) % 8 == 0, "chunk stride must keep the swizzle phase");
            const int r0 = tc / TC_CHUNKS, pc0 = tc % TC_CHUNKS;
            const float4 m = *reinterpret_cast<const float4*>(
                s_mean + kt * TC_BK + ((pc0 ^ ((r0 >> TC_SWZ_SHIFT) & (TC_CHUNKS - 1))) << 2));
            // all loads first: the stores below may alias them as far as the compiler knows
#pragma unroll
            for (int i = 0; i < NCH; ++i) v[i] = *reinterpret_cast<const float4*>(sa + TC_OFF_AHI + (tc + 32 * TC_CONV_WARPS * i) * 16);
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
              const int q = tc + 32 * TC_CONV_WARPS * i;
              float4 hi, lo;
              v[i].x -= m.x; v[i].y -= m.y; v[i].z -= m.z; v[i].w -= m.w;
              hi.x = rn_tf32(v[i].x); hi.y = rn_tf32(v[i].y); hi.z = rn_tf32(v[i].z); hi.w = rn_tf32(v[i].w);
              lo.x = rn_tf32(v[i].x - hi.x); lo.y = rn_tf32(v[i].y - hi.y); lo.z = rn_tf32(v[i].z - hi.z); lo.w = rn_tf32(v[i].w - hi.w);
              *reinterpret_cast<float4*>(sa + TC_OFF_AHI + q * 16) = hi;
              *reinterpret_cast<float4*>(sa + TC_OFF_ALO + q * 16) = lo;
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tensor-core (async proxy) reads
          asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(32 * TC_CONV_WARPS) : "memory");   // the converter threads of this group
          if (tc == 0) mbar_arrive_cluster(leader_conv + 8 * stage);      // one remote arrival instead of 128
          if (p.trace && blockIdx.x == 0 && tc == 0 && g < 256) p.trace[g * 5 + 2] = clock64();
          ++g;
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> F =====
    // Shared memory is the scarce resource of this kernel (tensor-core operand reads + converter +
    // TMA fills), so the accumulator goes straight from registers to global memory.  The 32x32b
    // load shape (thread = row) would scatter 16-byte pieces over 32 rows per store (measured 21k
    // cycles per tile); the 16x256b fragment shape lets a warp-wide 8-byte store write 8 rows x one
    // full 32-byte sector.
    const int quad = warp & 3;                             // TMEM lanes 32*quad .. +31 belong to this warp
    const uint32_t leader_tempty = mapa_u32(bar_tempty, 0u);
    uint32_t as = 0, aphase = 0;
    for (uint32_t it = 0; it < n_iter; ++it) {
      const uint32_t mt = (pair0 + it * pairs_in_grid) * 2u + cta_rank;
      const bool live = mt < p.n_m_tiles;
      for (int nt = 0; nt < p.n_tiles; ++nt) {
        mbar_wait(bar_tfull + 8 * as, aphase);
        tc_fence_after();
#pragma unroll 1
        for (int r = 0; r < TC_RB; ++r) {
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {                    // lanes 0-15 / 16-31 of the quadrant
            const uint64_t row_a = (uint64_t)mt * (TC_RB * TC_BM) + (uint64_t)(r * TC_BM + quad * 32 + h * 16 + (lane >> 2));
            float* dst_a = p.f + row_a * (uint64_t)p.d + (uint64_t)(nt * p.nt + 2 * (lane & 3));
            float* dst_b = dst_a + 8 * (uint64_t)p.d;
            const bool ok_a = live && row_a < p.n, ok_b = live && row_a + 8 < p.n;
            for (int c0 = 0; c0 < p.nt; c0 += 64) {
              uint32_t v[32];
              const int ncol = min(64, p.nt - c0);         // nt = 32: half of the 64-column load is another tile's columns
              tmem_ld_16x256b_x8(tmem_base + ((uint32_t)(quad * 32 + h * 16) << 16) + (as * TC_RB + (uint32_t)r) * (uint32_t)p.nt + (uint32_t)c0, v);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (8 * i < ncol) {
                  if (ok_a) __stcs(reinterpret_cast<float2*>(dst_a + c0 + 8 * i), make_float2(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])));
                  if (ok_b) __stcs(reinterpret_cast<float2*>(dst_b + c0 + 8 * i), make_float2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])));
                }
              }
            }
          }
        }
        tc_fence_before();
        asm volatile("bar.sync 8, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");        // all epilogue threads of this CTA
        if (threadIdx.x == 32 * (2 + TC_CONV_GROUPS * TC_CONV_WARPS)) mbar_arrive_cluster(leader_tempty + 8 * as);
        if (++as == TC_ACC) { as = 0; aphase ^= 1u; }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();            // no CTA leaves while the pair's MMAs read its shared memory or peers arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// prec [k][n] -> split[0][n][k] = rn_tf32(prec), split[1][n][k] = rn_tf32(prec - hi)   (B operand, K-major)
__global__ void __launch_bounds__(256)
prec_split_kernel(const float* __restrict__ prec, int d, float* __restrict__ split) {
  const size_t total = (size_t)d * d;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t nn = e / d, kk = e % d;
    const float v = prec[kk * d + nn];
    const float hi = rn_tf32(v);
    split[e] = hi;
    split[total + e] = rn_tf32(v - hi);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// fp32 matrix [rows][cols] with row pitch ld floats; box = [box_rows][TC_BK floats], swizzle = row bytes
static int make_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(MCCB200_ERR_CUDA, "gaussian_full_drift_tc: cuTensorMapEncodeTiled unavailable");
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BK, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   TC_ROW_BYTES == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : TC_ROW_BYTES == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MCCB200_ERR_CUDA, "gaussian_full_drift_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return MCCB200_OK;
}

static long long* g_tc_trace = nullptr;   // debug only (mccb200_debug_gemm_trace)

static bool tc_shape_ok(uint64_t n, int d, int ld, const void* y) {
  return d >= 32 && d <= TC_MAX_D && d % 32 == 0 && ld % 4 == 0 && ((uintptr_t)y & 15) == 0 && n >= 1 &&
         n < 0x7FFFFFFFull;
}

}  // namespace mccb

using namespace mccb;

// Debug hook (not part of the C ABI of include/mccube_b200.h): device buffer of 256 * 5 int64 clock stamps.
extern "C" __attribute__((visibility("default"))) void mccb200_debug_gemm_trace(long long* buf) { g_tc_trace = buf; }

extern "C" int mccb200_gaussian_full_drift_tc_supported(uint64_t n, int d, int ld) {
  return tc_shape_ok(n, d, ld, nullptr) ? 1 : 0;
}

extern "C" size_t mccb200_gaussian_full_drift_tc_prepare_bytes(int d) { return (size_t)2 * d * d * sizeof(float); }

extern "C" int mccb200_gaussian_full_drift_tc_prepare(const float* prec, int d, float* prec_split, void* stream) {
  MCCB_REQUIRE(prec && prec_split && d >= 1, "gaussian_full_drift_tc_prepare: bad argument");
  const size_t total = (size_t)d * d;
  prec_split_kernel<<<(int)min((total + 255) / 256, (size_t)sm_count() * 8), 256, 0, (cudaStream_t)stream>>>(prec, d, prec_split);
  count_launches(1);
  return check_launch("gaussian_full_drift_tc_prepare");
}

extern "C" int mccb200_gaussian_full_drift_tc(const float* y, uint64_t n, int d, int ld, const float* mean,
                                              const float* prec_split, float* f, void* stream) {
  MCCB_REQUIRE(y && mean && prec_split && f, "gaussian_full_drift_tc: NULL argument");
  if (n == 0) return MCCB200_OK;
  if (!tc_shape_ok(n, d, ld, y) || ((uintptr_t)f & 15) || ((uintptr_t)prec_split & 15))
    return fail(MCCB200_ERR_UNSUPPORTED, "gaussian_full_drift_tc: needs d %% 32 == 0, 32 <= d <= %d, 16-byte aligned rows (d=%d ld=%d)",
                TC_MAX_D, d, ld);
  TcDriftParams p{};
  p.n = n; p.d = d;
  p.nt = d % 256 == 0 ? 256 : d % 128 == 0 ? 128 : d % 64 == 0 ? 64 : 32;
  p.n_tiles = d / p.nt;
  p.k_tiles = d / TC_BK;
  p.n_m_tiles = (uint32_t)((n + TC_RB * TC_BM - 1) / (TC_RB * TC_BM));
  p.mean = mean; p.f = f;
  { const char* dbg = getenv("MCCB200_DEBUG_GEMM"); p.debug = dbg ? atoi(dbg) : 0; }
  p.trace = g_tc_trace;
  CUtensorMap map_y, map_bhi, map_blo;
  int rc = make_map(&map_y, y, n, (uint64_t)d, (uint64_t)ld, TC_RB * TC_BM);
  if (rc) return rc;
  rc = make_map(&map_bhi, prec_split, (uint64_t)d, (uint64_t)d, (uint64_t)d, (uint32_t)p.nt / TC_CLUSTER);
  if (rc) return rc;
  rc = make_map(&map_blo, prec_split + (size_t)d * d, (uint64_t)d, (uint64_t)d, (uint64_t)d, (uint32_t)p.nt / TC_CLUSTER);
  if (rc) return rc;
  cudaFuncSetAttribute(gaussian_full_drift_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
  int grid = (int)min((uint32_t)sm_count(), p.n_m_tiles);
  grid = (grid + TC_CLUSTER - 1) / TC_CLUSTER * TC_CLUSTER;   // whole clusters
  if (grid > sm_count()) grid -= TC_CLUSTER;
  if (grid < TC_CLUSTER) grid = TC_CLUSTER;
  phase_mark("gemm_drift_tc", (cudaStream_t)stream);
  gaussian_full_drift_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, (cudaStream_t)stream>>>(map_y, map_bhi, map_blo, p);
  phase_mark("end", (cudaStream_t)stream);
  count_launches(1);
  return check_launch("gaussian_full_drift_tc");
}
