// tcgen05 / TMEM / TMA implicit-GEMM convolution over the split-fp16 NHWC format (sm_100a).
//
//   D[pixel, n] = sum over (tap, channel)  A[pixel @ tap, channel] * W[n, tap, channel]
//
// * M = 128 output pixels (a BH x BW spatial box), N = up to 256 output channels, K = 64 channels of
//   one filter tap per pipeline stage.  Nothing is im2col'ed: the A tile of a tap is ONE TMA box load
//   from the NHWC activation at the tap's spatial offset; out-of-image rows/columns (the conv padding)
//   and channels past the end arrive as zeros from TMA's out-of-bounds fill.  Stride-2 layers view the
//   activation as [H/2][2][W/2][2][C] so a tap is still a dense box.
// * fp16x3: operands are (hi, lo) fp16 pairs; each K=16 slice issues three tcgen05.mma
//   (hi*hi + hi*lo + lo*hi) into one fp32 TMEM accumulator -- fp32-grade products at 1/3 of the fp16
//   tensor rate instead of falling back to CUDA cores.
// * Warp-specialised persistent CTAs: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) and
//   TMEM owner, warps 2-5 = epilogue (tcgen05.ld -> scale/shift/residual/activation -> split NHWC and/or
//   fp32 NCHW).  The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the
//   mainloop of tile i+1.  Small maps use deterministic split-K through an fp32 workspace.
#include <cuda.h>
#include <stdio.h>

#include <algorithm>
#include <stdlib.h>
#include <string.h>

#include "kernels.h"
#include "tc_common.cuh"

namespace accel {

cudaError_t launch_splitk_epilogue(const float* partial, int splits, int npix, int Cout_pad, int Wo, const Epilogue& epi,
                                   cudaStream_t stream);

namespace {

using namespace tc;

constexpr int BM = 128;          // pixels per tile (UMMA M)
constexpr int BK = 64;           // channels per stage: 128 bytes of fp16 = one SWIZZLE_128B row
constexpr int kEpiWarps = 8;      // two warps per TMEM lane quarter; they interleave the 32-column chunks of a tile
constexpr int kDmaWarps = 2;      // TMA epilogue: one thread per chunk set issues the tensor stores / residual loads (warps 10, 11)
constexpr int kThreads = 64 + 32 * kEpiWarps + 32 * kDmaWarps;
constexpr int kThreadsPair = 64 + 32 * kEpiWarps;           // conv_tc2_kernel has no DMA warps
constexpr int kMaxStages = 6;
constexpr int kSmemMaxDynamic = 227 * 1024 - 6 * 1024;   // 227 KB per CTA minus the static barriers/slots/scale-shift stage
// epilogue staging for TMA stores: per chunk set (2) x double buffer (2) x [hi | lo] x 128 pixel rows x 64 bytes
constexpr int kStageOut = 2 * 2 * 2 * 8192;
constexpr int kSmemBudget = kSmemMaxDynamic - 1024;      // minus the 1024-byte alignment slack
constexpr int kSmemEpi1 = 3 * 65536 + 2 * 2 * 8192;      // conv_tc_kernel<..., 8, 1>: three 64 KB stages + 32 KB of staging

struct alignas(64) TcParams {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  CUtensorMap o_hi, o_lo;      // main split output, box = 32 channels x the tile's pixel box (TMA store), when tma_out
  CUtensorMap r_hi, r_lo;      // residual, same box (TMA load into the same staging tile), when tma_out and a residual exists
  Epilogue epi;
  float* partial;
  int BW, BH, BN, stages;
  int tiles_x, tiles_y, n_tiles, splits, kiters, chunks, ntaps;
  int stride2, Cin_pad, Ho, Wo, Cout_pad, tmem_cols;   // Ho: the loop space's height over ALL frames (nb * Hof)
  int nb, Hof;     // batched plans: nb frames stacked along H, Hof output rows each; tiles never straddle two frames and
                   // the A operand's 4-D tensor map (channel, x, y, frame) keeps the conv padding per frame
  int krot;
  int pair;        // 1: conv_tc2_kernel (cta_group::2 pairs)
  int ncat;        // 1: hi*hi and hi*lo issued as ONE MMA over [B_hi ; B_lo] (N = 2*BN, two accumulator halves summed in the epilogue)
  int tma_out;     // 1: the main output leaves through shared memory + TMA tensor stores (3-D map), 2: 5-D map (transposed-conv phase)
  int debug;       // tuning aid (ACCEL_TC_DEBUG): 1 = no epilogue stores, 2 = no residual loads, 4 = no tcgen05.ld, 8 = direct epilogue; TMA epilogue: 16 = no residual loads, 32 = no tensor stores, 64 = no tcgen05.ld, 128 = no MMAs, 256 / 512 / 1024 = no operand / A / B loads
  int vec32;       // every split-NHWC operand of the epilogue is 32-byte aligned with 32-byte row pitch
  // A-slab reuse (stride-1 multi-tap layers whose tile is one full 128-pixel row segment): the taps of one filter row
  // (same dy, dx = dxmin .. dxmax) read the SAME source pixels shifted by whole pixels, so ONE TMA box of
  // slab_w = 128 + dxmax - dxmin pixels x 64 channels serves all ndx of them: the MMA of tap dx starts
  // (dx - dxmin) rows (128 bytes each) into the slab, the descriptor's base-offset field carrying the swizzle phase.
  // A bytes through L2 -> SM drop by ndx (3 for a 3x3); the weights keep their own ring.
  int aslab, ndx, dxmin, slab_w, slab_pl, sa_stages, slab_bo;
  unsigned ring_bytes;   // operand rings in front of the epilogue staging area
  int prefetch;    // > 0: the producer asks L2 for the NEXT item's first `prefetch` A stages (and, bit 8, its residual tile) one
                   // whole tile period ahead (cp.async.bulk.prefetch.tensor): short-K layers keep only 2-3 stages = 128-192 KB
                   // in flight per SM, less than DRAM latency x the SM's L2 port needs (tools/tc_trace.py: 3.8k cycles from
                   // the first TMA of a tile to its arrival)
  int mcast;       // 1: launched as clusters of two CTAs that work on two M tiles of the SAME N tile in lockstep; each CTA loads one
                   // plane of the weight tile (B_hi / B_lo) and multicasts it into both CTAs' rings: 48 instead of 64 KB of L2
                   // reads per CTA and K stage (the L2 output, ~20 TB/s over all SMs, is what bounds the mainloop)
  int epi1;        // TMA epilogue, half-chunk variant with ONE chunk set (conv_tc_kernel<..., 8, 1>): 32 KB of staging, one more
                   // operand stage in the ring
  int epiw16;      // TMA epilogue with 16 epilogue warps (conv_tc_kernel<..., 16>) when the launch has a plain main output only
  int res_ahead;   // TMA epilogue: chunks of L2 prefetch distance for the residual tiles (0: none)
  int kchains;     // 2: the K slices of a tile alternate between the two TMEM accumulator buffers and the epilogue sums
                   // them in fp32 (round-to-nearest): the tensor core's own accumulation truncates, so its error grows
                   // with the number of accumulation steps per accumulator; long-K layers only
  int fused;       // split-K only: the CTA that delivers a tile's LAST partial slab sums the slabs (in split order) and
                   // runs the epilogue itself -- no splitk_epilogue_kernel launch (ACCEL_TC_FUSED_SPLITK)
  unsigned* counters;   // fused: arrivals per output tile, self-resetting (behind the partial slabs)
  int8_t dy[kMaxTaps];
  int8_t dx[kMaxTaps];
};


// Descriptor of a SWIZZLE_128B K-major operand that starts a whole number of 128-byte rows into a 1024-byte swizzle
// atom: bits [49,52) = (start address >> 7) & 7 tell the tensor core the swizzle phase of the first row.
__device__ __forceinline__ uint64_t umma_desc_rows(uint32_t saddr, int mode) {
  if (mode == 0) return tc::umma_desc(saddr);                                                   // address only
  if (mode == 2) return tc::umma_desc(saddr & ~1023u) | ((uint64_t)((saddr >> 7) & 7u) << 49);   // atom address + phase
  return tc::umma_desc(saddr) | ((uint64_t)((saddr >> 7) & 7u) << 49);
}

// In-kernel timeline of CTA 0 (ACCEL_TC_DEBUG bit 2048, tools/bench_layer.py --trace): (tag, item, clock64) records.
//   0 start | 1 producer: first TMA of the item  2 last TMA | 3 MMA warp: accumulators free  4 first stage landed
//   5 last MMA + commit issued | 6 epilogue warp 2: item prologue done, waiting for the accumulator  7 accumulator
//   complete  8 chunk loop done, accumulator released | 9 end
// Compiled in only with -DACCEL_TC_TRACE_BUILD (`python accel_b200/build.py --trace` -> libaccel_b200_trace.so): the extra
// code in the role loops costs a few percent on short-K layers.
__device__ long long g_tc_trace[3 * 2048];
__device__ unsigned g_tc_trace_n;
#ifndef ACCEL_TC_TRACE_BUILD
#define TC_TRACE(tag, item) do { } while (0)
#else
#define TC_TRACE(tag, item)                                                              \
  do {                                                                                   \
    if ((P.debug & 2048) && blockIdx.x == 0 && lane == 0) {                              \
      const unsigned ti_ = atomicAdd(&g_tc_trace_n, 1u);                                 \
      if (ti_ < 2048u) { g_tc_trace[3 * ti_] = (tag); g_tc_trace[3 * ti_ + 1] = (item); g_tc_trace[3 * ti_ + 2] = clock64(); } \
    }                                                                                    \
  } while (0)
#endif

// NCAT: hi*hi and hi*lo issued as ONE MMA over [B_hi ; B_lo] (N = 2*BN <= 256), see the MMA issuer.
// FUSED: in-kernel split-K tail (TcParams::fused); a template parameter so that the default instantiations keep their
// register allocation.
// TWO: two accumulation chains (TcParams::kchains == 2): a template parameter so that the ordinary instantiations do not
// carry the registers of the folded first half, and the two-chain ones do not carry the residual prefetch registers.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_idx() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_count() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// B-operand multicast (TcParams::mcast): one TMA load lands in BOTH CTAs of the cluster, at the same shared-memory offset, and
// completes bytes on each CTA's own mbarrier at the same offset.
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {   // arrives on `bar` in every CTA of the mask
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

// EPIW: epilogue warps.  8 = two per TMEM lane quarter (each thread one pixel row x 32 channels of a chunk).  16 (TMA epilogue
// with a plain main output only, TcParams::epiw): four per quarter -- two chunk sets x two 16-channel halves of a chunk: the
// TMA epilogue is bound by instruction issue / latency with 8 warps (ncu: issue slots 39 % busy, 11.4k warp instructions per
// 128 x 128 tile), so the expand convs, whose mainloop is 4 K stages, spent 80 % of their time in it.
// CSETS = 1 (with EPIW = 8): the half-chunk epilogue with ONE chunk set -- all eight warps work on the same 32-channel chunk,
// 32 KB of staging instead of 64, so that a THIRD 64 KB operand stage fits (229,376 dynamic + 3 KB static = the 227 KB limit
// exactly): the expand convs' 2-stage ring was bound by its commit -> producer -> load -> issuer round trip.
template <bool NCAT, bool FUSED = false, bool TWO = false, int EPIW = 8, int CSETS = 2>
__global__ void __launch_bounds__(64 + 32 * EPIW + 32 * kDmaWarps, 1) conv_tc_kernel(const __grid_constant__ TcParams P) {
  constexpr bool kHalf = EPIW == 16 || CSETS == 1;           // half-chunk TMA epilogue (a thread owns 16 channels of a chunk)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 4];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float epi_sc[2][2][kHalf ? 128 : 256];   // [accumulator][scale | shift][channel of the tile]
  __shared__ __align__(8) uint64_t res_bars[4];              // TMA epilogue: staging buffer [chunk set][buffer] is free / holds its residual
  __shared__ __align__(8) uint64_t stg_bars[4];              // TMA epilogue: all 128 threads of the chunk set wrote their result rows
  __shared__ int s_last;                                     // fused split-K: this CTA delivered the tile's last slab

  // operand ring: [stage][A_hi | A_lo | B_hi | B_lo]
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = BM * 128u, b_bytes = (uint32_t)P.BN * 128u;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  const uint32_t stg0 = smem0 + P.ring_bytes;                       // epilogue staging (1024-byte aligned)
  // A-slab mode: [sa_stages x (slab hi | slab lo)] [stages x (B_hi | B_lo)]
  __shared__ __align__(8) uint64_t abars[8];
  const uint32_t afull0 = smem_u32(&abars[0]), aempty0 = smem_u32(&abars[4]);
  const uint32_t bring0 = smem0 + (uint32_t)P.sa_stages * 2u * (uint32_t)P.slab_pl;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kMaxStages]);
  const uint32_t tfull0 = smem_u32(&bars[2 * kMaxStages]), tempty0 = smem_u32(&bars[2 * kMaxStages + 2]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, P.mcast ? 2 : 1);         // mcast: the slot is rewritten by both CTAs' loads -> both MMA warps free it
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, EPIW);
    }
    for (int a = 0; a < 4; ++a) mbar_init(smem_u32(&res_bars[a]), 1);
    for (int a = 0; a < 4; ++a) mbar_init(smem_u32(&stg_bars[a]), 32 * EPIW / CSETS);   // the threads of one chunk set
    for (int a = 0; a < 8; ++a) mbar_init(smem_u32(&abars[a]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)P.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_trigger();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (P.mcast) cluster_sync_all();                         // the peer's barriers exist before anything is multicast to them
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_wait();                                              // the prologue above touched no global memory
  const uint32_t tmem_base = tmem_base_slot;

  // mcast (TcParams::mcast): clusters of two CTAs walk pair-items q = (two M tiles, one N tile); rank r takes M tile 2*m2 + r
  const bool mc = P.mcast != 0;
  const int rank = mc ? (int)cluster_ctarank() : 0, mstep = mc ? 2 : 1;
  const int it0 = mc ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, itstep = mc ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int tiles = P.tiles_x * P.tiles_y * P.n_tiles;
  const int items = mc ? ((P.tiles_x * P.tiles_y + 1) / 2) * P.n_tiles : tiles * P.splits;
  if (warp == 0) TC_TRACE(0, 0);

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0 && P.aslab) {
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      const uint32_t slab_tx = 2u * (uint32_t)P.slab_w * 128u;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int split = item % P.splits, tile = item / P.splits;
        const int nt = tile % P.n_tiles, mt = tile / P.n_tiles;
        const int x0 = (mt % P.tiles_x) * P.BW, y0 = (mt / P.tiles_x) * P.BH;
        const int kb = (int)(((long long)P.kiters * split) / P.splits);
        const int ke = (int)(((long long)P.kiters * (split + 1)) / P.splits);
        const int fr = y0 / P.Hof, yl = y0 - fr * P.Hof;         // frame of the tile and its first row inside that frame
        for (int it = kb; it < ke; ++it) {
          // K order: filter row j, channel chunk kc, then the ndx taps of the row (they share one slab)
          const int outer = it / P.ndx, i = it - outer * P.ndx;
          const int j = outer / P.chunks, kc = outer - j * P.chunks;
          const int t = j * P.ndx + i;
          if (i == 0 || it == kb) {
            mbar_wait(aempty0 + 8 * as, aph ^ 1);
            const uint32_t fa = afull0 + 8 * as;
            mbar_arrive_expect_tx(fa, slab_tx);
            const uint32_t dst = smem0 + (uint32_t)as * 2u * (uint32_t)P.slab_pl;
            tma_load_4d(dst, &P.a_hi, fa, kc * BK, x0 + P.dxmin, yl + P.dy[t], fr);
            tma_load_4d(dst + (uint32_t)P.slab_pl, &P.a_lo, fa, kc * BK, x0 + P.dxmin, yl + P.dy[t], fr);
            if (++as == P.sa_stages) { as = 0; aph ^= 1; }
          }
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          const uint32_t fb = full0 + 8 * s;
          mbar_arrive_expect_tx(fb, 2 * b_bytes);
          const uint32_t sb = bring0 + (uint32_t)s * 2u * b_bytes;
          const int kcol = t * P.Cin_pad + kc * BK;
          tma_load_2d(sb, &P.b_hi, fb, kcol, nt * P.BN);
          tma_load_2d(sb + b_bytes, &P.b_lo, fb, kcol, nt * P.BN);
          if (++s == P.stages) { s = 0; ph ^= 1; }
        }
      }
    } else if (!P.aslab) {
      // convergent warp, elect.sync around the TMA instructions (see the MMA issuer)
      int s = 0;
      uint32_t ph = 0;
      const bool ldA = !(P.debug & (256 | 512)), ldB = !(P.debug & (256 | 1024));   // timing decomposition only
      const uint32_t tx = (ldA ? 2 * a_bytes : 0u) + (ldB ? 2 * b_bytes : 0u);
      for (int item = it0; item < items; item += itstep) {
        const int split = item % P.splits, tile = item / P.splits;
        const int nt = tile % P.n_tiles, mt = (tile / P.n_tiles) * mstep + rank;
        const int x0 = (mt % P.tiles_x) * P.BW, y0 = (mt / P.tiles_x) * P.BH;
        const int kb = (int)(((long long)P.kiters * split) / P.splits);
        const int ke = (int)(((long long)P.kiters * (split + 1)) / P.splits);
        const int fr = y0 / P.Hof, yl = y0 - fr * P.Hof;         // frame of the tile and its first row inside that frame
        const int rot = P.krot ? (int)(((unsigned)item * 5u) % (unsigned)(ke - kb)) : 0;
        if (P.prefetch && item + itstep < items) {
          const int item2 = item + itstep;
          const int split2 = item2 % P.splits, tile2 = item2 / P.splits;
          const int nt2 = tile2 % P.n_tiles, mt2 = (tile2 / P.n_tiles) * mstep + rank;
          const int x2 = (mt2 % P.tiles_x) * P.BW, y2 = (mt2 / P.tiles_x) * P.BH;
          const int fr2 = y2 / P.Hof, yl2 = y2 - fr2 * P.Hof;
          const int kb2 = (int)(((long long)P.kiters * split2) / P.splits);
          const int ke2 = (int)(((long long)P.kiters * (split2 + 1)) / P.splits);
          const int np = min(ke2 - kb2, P.prefetch & 255);
          if (elect_one()) {
            for (int j = 0; j < np; ++j) {
              const int it2 = kb2 + j;
              const int t2 = it2 / P.chunks, kc2 = it2 - t2 * P.chunks;
              const int dy2 = P.dy[t2], dx2 = P.dx[t2];
              if (P.stride2) {
                tma_prefetch_5d(&P.a_hi, kc2 * BK, dx2 & 1, x2 + (dx2 >> 1), dy2 & 1, y2 + (dy2 >> 1));
                tma_prefetch_5d(&P.a_lo, kc2 * BK, dx2 & 1, x2 + (dx2 >> 1), dy2 & 1, y2 + (dy2 >> 1));
              } else {
                tma_prefetch_4d(&P.a_hi, kc2 * BK, x2 + dx2, yl2 + dy2, fr2);
                tma_prefetch_4d(&P.a_lo, kc2 * BK, x2 + dx2, yl2 + dy2, fr2);
              }
            }
            if ((P.prefetch & 256) && P.tma_out == 1 && P.epi.res_hi != nullptr) {
              for (int c = 0; c < P.BN; c += 32) {
                tma_prefetch_3d(&P.r_hi, nt2 * P.BN + c, x2, y2);
                tma_prefetch_3d(&P.r_lo, nt2 * P.BN + c, x2, y2);
              }
            }
          }
          __syncwarp();
        }
        int it = kb + rot;                                // each CTA walks K from its own offset: neighbours do not
        int t = it / P.chunks, kc = it - t * P.chunks;    // stream the same weight lines at the same moment
        for (int i = kb; i < ke; ++i) {
          if (P.debug & 8192) mbar_wait_spin(empty0 + 8 * s, ph ^ 1);
          else mbar_wait(empty0 + 8 * s, ph ^ 1);
          const uint32_t fb = full0 + 8 * s;
          const uint32_t sa = smem0 + s * stage_bytes;
          const int dy = P.dy[t], dx = P.dx[t];
          const int kcol = t * P.Cin_pad + kc * BK;
          if (P.debug & 4096) TC_TRACE(11, i);               // per stage: ring slot free, loads issued
          if (i == kb) TC_TRACE(1, item);
          if (i == ke - 1) TC_TRACE(2, item);
          if (elect_one()) {
            mbar_arrive_expect_tx(fb, tx);
            if (!ldA) {
            } else if (P.stride2) {
              tma_load_5d(sa, &P.a_hi, fb, kc * BK, dx & 1, x0 + (dx >> 1), dy & 1, y0 + (dy >> 1));
              tma_load_5d(sa + a_bytes, &P.a_lo, fb, kc * BK, dx & 1, x0 + (dx >> 1), dy & 1, y0 + (dy >> 1));
            } else {
              tma_load_4d(sa, &P.a_hi, fb, kc * BK, x0 + dx, yl + dy, fr);
              tma_load_4d(sa + a_bytes, &P.a_lo, fb, kc * BK, x0 + dx, yl + dy, fr);
            }
            if (ldB && mc) {                               // each CTA fetches one plane of the shared weight tile for both
              if (rank == 0) tma_load_2d_mc(sa + 2 * a_bytes, &P.b_hi, fb, kcol, nt * P.BN, (uint16_t)3);
              else tma_load_2d_mc(sa + 2 * a_bytes + b_bytes, &P.b_lo, fb, kcol, nt * P.BN, (uint16_t)3);
            } else if (ldB) {
              tma_load_2d(sa + 2 * a_bytes, &P.b_hi, fb, kcol, nt * P.BN);
              tma_load_2d(sa + 2 * a_bytes + b_bytes, &P.b_lo, fb, kcol, nt * P.BN);
            }
          }
          __syncwarp();
          if (++s == P.stages) { s = 0; ph ^= 1; }
          if (++kc == P.chunks) { kc = 0; ++t; }          // next K stage: channel chunks fastest, then taps ...
          if (++it == ke) { it = kb; t = it / P.chunks; kc = it - t * P.chunks; }   // ... wrapping around at the rotation
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    // The WHOLE warp walks this loop convergently and only the tcgen05 instructions are predicated on elect.sync: every
    // descriptor, TMEM address and barrier address then lives in uniform registers.  Under `if (lane == 0)` the compiler
    // cannot prove them warp-uniform and wraps each UTCHMMA in an ELECT / R2UR.BROADCAST waterfall; measured
    // (tools/mma_probe.cu, profiles/r02_mma_probe.txt) that issue path cost ~180 cycles per MMA in this kernel against the
    // tensor core's 64 (N = 128) / 128 (N = 256) cycles -- the mainloop was ISSUE-bound.
    {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(P.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * P.BN) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t acc_cols = (uint32_t)(NCAT ? 2 * P.BN : P.BN);
      constexpr bool two = TWO;
      const bool no_mma = (P.debug & 128) != 0;
      const uint32_t bn = (uint32_t)P.BN;
      int s = 0, acc = 0;
      uint32_t ph = 0, accph = 0;
      int nas = 0, cur = 0;                               // A-slab ring: next slot to consume, slot in use
      uint32_t aphc = 0;
      for (int item = it0; item < items; item += itstep) {
        const int split = item % P.splits;
        const int kb = (int)(((long long)P.kiters * split) / P.splits);
        const int ke = (int)(((long long)P.kiters * (split + 1)) / P.splits);
        // two chains (TcParams::kchains): the first half of the item's K stages accumulates in TMEM buffer 0, the second half in
        // buffer 1 -- each with the ordinary double-buffer protocol, so the epilogue warps fold buffer 0 into registers while
        // the second half still runs, and the next item's first half starts as soon as its MMAs are issued.
        const int kmid = two ? kb + (ke - kb + 1) / 2 : ke;
        for (int half = 0; half < (two ? 2 : 1); ++half) {
          const int hb = half ? kmid : kb, he = half ? ke : kmid;
          mbar_wait(tempty0 + 8 * acc, accph ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (half == 0) TC_TRACE(3, item);
          const uint32_t d = tmem_base + (uint32_t)acc * acc_cols;
          for (int it = hb; it < he; ++it) {
            uint32_t a_hi, a_lo, b_hi, b_lo;
            int slab_mode = 0;
            bool slab_done = false;
            if (P.aslab) {
              const int outer = it / P.ndx, i = it - outer * P.ndx;
              const int t = (outer / P.chunks) * P.ndx + i;
              if (i == 0 || it == kb) {
                cur = nas;
                mbar_wait(afull0 + 8 * cur, aphc);
                if (++nas == P.sa_stages) { nas = 0; aphc ^= 1; }
              }
              slab_done = i == P.ndx - 1 || it == ke - 1;
              a_hi = smem0 + (uint32_t)cur * 2u * (uint32_t)P.slab_pl + (uint32_t)(P.dx[t] - P.dxmin) * 128u;
              a_lo = a_hi + (uint32_t)P.slab_pl;
              b_hi = bring0 + (uint32_t)s * 2u * b_bytes;
              b_lo = b_hi + b_bytes;
              slab_mode = 1;
            } else {
              a_hi = smem0 + s * stage_bytes;
              a_lo = a_hi + a_bytes;
              b_hi = a_hi + 2 * a_bytes;
              b_lo = b_hi + b_bytes;
            }
            if (P.debug & 8192) mbar_wait_spin(full0 + 8 * s, ph);
            else mbar_wait(full0 + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (it == kb) TC_TRACE(4, item);
            if (P.debug & 4096) TC_TRACE(10, it);             // per stage: operands landed
            const uint64_t ah = slab_mode ? umma_desc_rows(a_hi, P.slab_bo) : umma_desc(a_hi);
            const uint64_t al = slab_mode ? umma_desc_rows(a_lo, P.slab_bo) : umma_desc(a_lo);
            const uint64_t bh = umma_desc(b_hi), bl = umma_desc(b_lo);
            const uint32_t first0 = it > hb ? 1u : 0u;
            if (elect_one()) {
              if (!no_mma) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  const uint64_t adv = (uint64_t)(k * 2);       // 16 fp16 = 32 bytes along K inside the swizzle atom
                  const uint32_t first = k >= 1 ? 1u : first0;
                  if (NCAT) {
                    // B_hi and B_lo are adjacent in the stage with one row pitch: a single N = 2*BN operand.  Columns
                    // [0, BN) collect hi*hi, columns [BN, 2BN) collect hi*lo and -- issued into the upper half alone --
                    // lo*hi: the small cross terms never touch the big accumulator, whose truncating additions are the
                    // dominant rounding error.  The tensor core reads the A_hi slice once and issues two instructions.
                    umma_f16(d, ah + adv, bh + adv, idesc2, first);
                    umma_f16(d + bn, al + adv, bh + adv, idesc, 1u);
                  } else {
                    umma_f16(d, ah + adv, bh + adv, idesc, first);
                    umma_f16(d, ah + adv, bl + adv, idesc, 1u);
                    umma_f16(d, al + adv, bh + adv, idesc, 1u);
                  }
                }
              }
              if (mc) umma_commit_mc(empty0 + 8 * s, (uint16_t)3);   // ... in both CTAs: either may write the other's slot
              else umma_commit(empty0 + 8 * s);               // frees the smem slot when these MMAs retire
              if (slab_done) umma_commit(aempty0 + 8 * cur);  // ... and the A slab after the last tap that reads it
            }
            __syncwarp();
            if (++s == P.stages) { s = 0; ph ^= 1; }
          }
          if (elect_one()) umma_commit(tfull0 + 8 * acc);     // accumulator complete -> epilogue
          __syncwarp();
          if (++acc == 2) { acc = 0; accph ^= 1; }
        }
        TC_TRACE(5, item);
      }
    }
  } else if (warp >= 2 + EPIW) {
    // ===================================== epilogue DMA (TMA epilogue only) ===================
    // One thread per chunk set owns every bulk operation of that set's two staging buffers, so that no epilogue thread ever
    // waits for a store: chunk k (buffer k & 1) is stored when its 128 threads have arrived on stg_bars, and once the
    // store has read the buffer it is handed to chunk k + 2 -- with that chunk's residual tile on the way (res_bars carries
    // the bytes) or, without a residual, by a plain arrival.
    const int cset = warp - (2 + EPIW);
    if (P.tma_out && P.splits == 1 && lane == 0 && cset < CSETS) {
      const Epilogue& E = P.epi;
      const bool has_res = E.res_hi != nullptr && !(P.debug & 16);
      const uint32_t sb0 = stg0 + (uint32_t)cset * 2u * 16384u;
      const uint32_t rbar0 = smem_u32(&res_bars[cset * 2]), sbar0 = smem_u32(&stg_bars[cset * 2]);
      constexpr int cstep = 32 * CSETS;
      const int cpi = P.BN > cset * 32 ? (P.BN - cset * 32 + cstep - 1) / cstep : 0;   // chunks of one item that belong to this set
      const int mine = it0 < items ? (items - it0 + itstep - 1) / itstep : 0;
      const int total = mine * cpi;
      auto coords = [&](int k, int& c0, int& x0, int& y0) {
        const int item = it0 + (k / cpi) * itstep;
        const int tile = item / P.splits, nt = tile % P.n_tiles, mt = (tile / P.n_tiles) * mstep + rank;
        c0 = nt * P.BN + cset * 32 + (k % cpi) * cstep;
        x0 = (mt % P.tiles_x) * P.BW;
        y0 = (mt / P.tiles_x) * P.BH;
      };
      auto fill = [&](int k) {                               // buffer k & 1 becomes chunk k's
        const uint32_t bar = rbar0 + 8u * (uint32_t)(k & 1);
        if (has_res) {
          int c0, x0, y0;
          coords(k, c0, x0, y0);
          const uint32_t dst = sb0 + (uint32_t)(k & 1) * 16384u;
          mbar_arrive_expect_tx(bar, 16384u);
          tma_load_3d(dst, &P.r_hi, bar, c0, x0, y0);
          tma_load_3d(dst + 8192u, &P.r_lo, bar, c0, x0, y0);
        } else {
          mbar_arrive(bar);
        }
      };
      // the residual tile of chunk k + res_ahead is asked into L2 now: a buffer is refilled only one chunk ahead of its
      // use, less than DRAM latency (ncu: 13 % of the kernel's stall samples sat on the wait for the residual bytes)
      const int res_ahead = has_res ? P.res_ahead : 0;
      auto ask = [&](int k) {
        int c0, x0, y0;
        coords(k, c0, x0, y0);
        tma_prefetch_3d(&P.r_hi, c0, x0, y0);
        tma_prefetch_3d(&P.r_lo, c0, x0, y0);
      };
      if (res_ahead > 0)
        for (int k = 2; k < min(total, 2 + res_ahead); ++k) ask(k);
      if (total > 0) fill(0);
      if (total > 1) fill(1);
      for (int k = 0; k < total; ++k) {
        if (res_ahead > 0 && k + 2 + res_ahead < total) ask(k + 2 + res_ahead);
        const uint32_t sb = sb0 + (uint32_t)(k & 1) * 16384u;
        mbar_wait(sbar0 + 8u * (uint32_t)(k & 1), (uint32_t)((k >> 1) & 1));
        if (!(P.debug & 32)) {
          int c0, x0, y0;
          coords(k, c0, x0, y0);
          if (P.tma_out == 2) {
            tma_store_5d(&P.o_hi, sb, c0, E.oox, x0, E.ooy, y0);
            tma_store_5d(&P.o_lo, sb + 8192u, c0, E.oox, x0, E.ooy, y0);
          } else {
            tma_store_3d(&P.o_hi, sb, c0, x0, y0);
            tma_store_3d(&P.o_lo, sb + 8192u, c0, x0, y0);
          }
          bulk_commit();
        }
        if (k + 2 < total) {
          bulk_wait_read0();                                 // the stores of chunk k have read the buffer
          fill(k + 2);
        }
      }
      bulk_wait0();
    }
  } else {
    if constexpr (kHalf) {
      // ===================================== epilogue, TMA both ways, 16 warps ==================
      // Same protocol as the 8-warp version below (buffers, mbarriers, DMA threads); a thread owns one pixel row and 16 of
      // the chunk's 32 channels: warp -> (TMEM lane quarter, chunk set, channel half).  Main split output (+ residual) only.
      const int e = warp - 2;
      const int quarter = warp & 3, cset = CSETS == 2 ? (e >> 2) & 1 : 0, chalf = CSETS == 2 ? e >> 3 : e >> 2;
      const int r = quarter * 32 + lane;
      const int et = threadIdx.x - 64;
      const Epilogue& E = P.epi;
      const bool has_res = E.res_hi != nullptr && !(P.debug & 16);
      const uint32_t sb0 = stg0 + (uint32_t)cset * 2u * 16384u;
      const uint32_t rbar0 = smem_u32(&res_bars[cset * 2]), sbar0 = smem_u32(&stg_bars[cset * 2]);
      const uint32_t acc_cols = (uint32_t)(NCAT ? 2 * P.BN : P.BN);
      const uint32_t row = (uint32_t)r * 64u, sw = ((uint32_t)r >> 1) & 3u;
      const uint32_t p0 = row + (((uint32_t)(2 * chalf) ^ sw) << 4), p1 = row + (((uint32_t)(2 * chalf + 1) ^ sw) << 4);
      int acc = 0, sci = 0, n = 0;
      uint32_t accph = 0;
      float sc_next = 1.f, sh_next = 0.f;
      auto fetch_sc = [&](int item2) {
        const int c = ((item2 / P.splits) % P.n_tiles) * P.BN + et;
        const bool in = et < P.BN && c < E.Cout;
        sc_next = (in && E.scale) ? __ldg(E.scale + c) : 1.f;
        sh_next = (in && E.shift) ? __ldg(E.shift + c) : 0.f;
      };
      if (it0 < items) fetch_sc(it0);
      for (int item = it0; item < items; item += itstep) {
        if (et < P.BN) {
          epi_sc[sci][0][et] = sc_next;
          epi_sc[sci][1][et] = sh_next;
        }
        if (item + itstep < items) fetch_sc(item + itstep);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPIW) : "memory");
        mbar_wait(tfull0 + 8 * acc, accph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * acc_cols + (uint32_t)(chalf * 16);
        for (int cc = cset * 32; cc < P.BN; cc += 32 * CSETS, ++n) {
          const int b = n & 1;
          const uint32_t sb = sb0 + (uint32_t)b * 16384u;
          float v[16];
          if (P.debug & 64) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.f;
          } else {
            tmem_ld16(taddr + cc, v);
            if (NCAT) {
              float v2[16];
              tmem_ld16(taddr + P.BN + cc, v2);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += v2[i];
            }
          }
          const float* sc = &epi_sc[sci][0][cc + chalf * 16];
          const float* sh = &epi_sc[sci][1][cc + chalf * 16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 s4 = *(reinterpret_cast<const float4*>(sc) + q);
            const float4 b4 = *(reinterpret_cast<const float4*>(sh) + q);
            v[4 * q + 0] = fmaf(v[4 * q + 0], s4.x, b4.x);
            v[4 * q + 1] = fmaf(v[4 * q + 1], s4.y, b4.y);
            v[4 * q + 2] = fmaf(v[4 * q + 2], s4.z, b4.z);
            v[4 * q + 3] = fmaf(v[4 * q + 3], s4.w, b4.w);
          }
          mbar_wait(rbar0 + 8u * b, (uint32_t)((n >> 1) & 1));          // buffer b is ours (and holds the residual rows)
          if (has_res) {
            uint32_t h[8], l[8];
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(h[0]), "=r"(h[1]), "=r"(h[2]), "=r"(h[3]) : "r"(sb + p0));
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(h[4]), "=r"(h[5]), "=r"(h[6]), "=r"(h[7]) : "r"(sb + p1));
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(l[0]), "=r"(l[1]), "=r"(l[2]), "=r"(l[3]) : "r"(sb + 8192u + p0));
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(l[4]), "=r"(l[5]), "=r"(l[6]), "=r"(l[7]) : "r"(sb + 8192u + p1));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 fh = unpack_h2(h[i]), fl = unpack_h2(l[i]);
              v[2 * i] += fh.x + fl.x;
              v[2 * i + 1] += fh.y + fl.y;
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = apply_act(v[i], E.act);
          uint32_t wh[8], wl[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) split_pair(v[2 * i], v[2 * i + 1], wh[i], wl[i]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + p0), "r"(wh[0]), "r"(wh[1]), "r"(wh[2]), "r"(wh[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + p1), "r"(wh[4]), "r"(wh[5]), "r"(wh[6]), "r"(wh[7]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + 8192u + p0), "r"(wl[0]), "r"(wl[1]), "r"(wl[2]), "r"(wl[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + 8192u + p1), "r"(wl[4]), "r"(wl[5]), "r"(wl[6]), "r"(wl[7]) : "memory");
          fence_async_smem();
          mbar_arrive(sbar0 + 8u * b);                       // -> the set's DMA thread stores the buffer
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        sci ^= 1;
        if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
        if (++acc == 2) { acc = 0; accph ^= 1; }
      }
    } else if (P.tma_out && P.splits == 1) {
      // ===================================== epilogue, TMA both ways ===========================
      // The lane-per-pixel global accesses of the direct epilogue are 32 separate lines per warp instruction and
      // saturate the L1TEX pipe (profiles/r01_ncu_dcn_col.txt shows the same pattern).  Here global memory is only
      // touched by the TMA unit: per chunk set (4 warps = the tile's 128 pixel rows) and 32-channel chunk, the
      // residual tile lands in a SWIZZLE_64B staging buffer (hi | lo, 16 KB), every thread folds its own 64+64 bytes
      // into its accumulator row and writes the result back IN PLACE, and the set's DMA thread (above) hands the
      // buffer to two tensor stores.  Two buffers per chunk set; the threads of a set synchronise through the
      // buffers' mbarriers only -- no named barrier, nobody waits for a store.
      const int quarter = warp & 3, cset = (warp - 2) >> 2;
      const int r = quarter * 32 + lane;
      const int by = r / P.BW, bx = r - by * P.BW;
      const int et = threadIdx.x - 64;
      const Epilogue& E = P.epi;
      const bool has_res = E.res_hi != nullptr && !(P.debug & 16);
      const uint32_t sb0 = stg0 + (uint32_t)cset * 2u * 16384u;
      const uint32_t rbar0 = smem_u32(&res_bars[cset * 2]), sbar0 = smem_u32(&stg_bars[cset * 2]);
      const uint32_t acc_cols = (uint32_t)(NCAT ? 2 * P.BN : P.BN);
      int acc = 0, sci = 0, n = 0;                          // sci: scale/shift staging slot, alternates per item
      uint32_t accph = 0;
      // this thread's scale / shift of the NEXT item travel in registers while the current item is processed: with one or
      // two K steps per tile the epilogue is the critical path and a load issued at the top of an item would be waited for
      float sc_next = 1.f, sh_next = 0.f;
      auto fetch_sc = [&](int item2) {
        const int c = ((item2 / P.splits) % P.n_tiles) * P.BN + et;
        const bool in = et < P.BN && c < E.Cout;
        sc_next = (in && E.scale) ? __ldg(E.scale + c) : 1.f;
        sh_next = (in && E.shift) ? __ldg(E.shift + c) : 0.f;
      };
      if (it0 < items) fetch_sc(it0);
      for (int item = it0; item < items; item += itstep) {
        const int tile = item / P.splits;
        const int nt = tile % P.n_tiles, mt = (tile / P.n_tiles) * mstep + rank;
        const int x0 = (mt % P.tiles_x) * P.BW, y0 = (mt / P.tiles_x) * P.BH;
        const int x = x0 + bx, y = y0 + by;
        const bool valid = x < P.Wo && y < P.Ho;
        const int pix = (y * E.osy + E.ooy) * E.OWf + x * E.osx + E.oox;
        const int nbase = nt * P.BN;
        if (et < P.BN) {
          epi_sc[sci][0][et] = sc_next;
          epi_sc[sci][1][et] = sh_next;
        }
        if (item + itstep < items) fetch_sc(item + itstep);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 2) TC_TRACE(6, item);
        mbar_wait(tfull0 + 8 * acc, accph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp == 2) TC_TRACE(7, item);
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * acc_cols;
        for (int cc = cset * 32; cc < P.BN; cc += 64, ++n) {
          const int b = n & 1;
          const uint32_t sb = sb0 + (uint32_t)b * 16384u;
          float v[32];
          if (P.debug & 64) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          } else {
          tmem_ld32(taddr + cc, v);
          if (NCAT) {
            float v2[32];
            tmem_ld32(taddr + P.BN + cc, v2);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += v2[i];
          }
          }
          mbar_wait(rbar0 + 8u * b, (uint32_t)((n >> 1) & 1));          // buffer b is ours (and holds the residual rows)
          ResChunk rc;
          if (!has_res) {
            rc = ResChunk{};
          } else {
            const uint32_t row = sb + (uint32_t)r * 64u, sw = ((uint32_t)r >> 1) & 3u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t a = row + (((uint32_t)j ^ sw) << 4);
              uint32_t* dh = &rc.h[j >> 1].v[(j & 1) * 4];
              uint32_t* dl = &rc.l[j >> 1].v[(j & 1) * 4];
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(dh[0]), "=r"(dh[1]), "=r"(dh[2]), "=r"(dh[3]) : "r"(a));
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(dl[0]), "=r"(dl[1]), "=r"(dl[2]), "=r"(dl[3]) : "r"(a + 8192u));
            }
          }
          if (valid) epilogue_chunk32(E, pix, nbase + cc, v, rc, &epi_sc[sci][0][cc], &epi_sc[sci][1][cc], false);
          uint32_t wh[16], wl[16];
          split32_words(v, wh, wl);
          stage_row64(sb, r, wh);
          stage_row64(sb + 8192u, r, wl);
          fence_async_smem();
          mbar_arrive(sbar0 + 8u * b);                       // -> the set's DMA thread stores the buffer
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (warp == 2) TC_TRACE(8, item);
        sci ^= 1;
        if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
        if (++acc == 2) { acc = 0; accph ^= 1; }
      }
    } else {
    // ===================================== epilogue ==========================================
    // Lane = pixel (TMEM lane), 32 consecutive output channels per chunk: 64 contiguous bytes per
    // plane per thread, moved as 256-bit accesses.  Everything the chunk loop would otherwise wait
    // for is fetched while the mainloop of the tile still runs: the tile's per-channel scale/shift go
    // to shared memory, and the residual is kept two chunks ahead in registers (2 x 128 B per thread
    // in flight = 64 KB per SM, what one SM's share of the HBM stream needs).
    const int quarter = warp & 3;                          // TMEM lane quarter this warp may read
    const int cset = (warp - 2) >> 2;                      // which interleaved chunk set (0 / 1)
    const int r = quarter * 32 + lane;                     // tile row = pixel inside the box
    const int by = r / P.BW, bx = r - by * P.BW;
    const int npix = P.Ho * P.Wo;
    const int et = threadIdx.x - 64;                       // 0 .. 255 among the epilogue threads
    const Epilogue& E = P.epi;
    constexpr bool two = TWO;
    const uint32_t acc_cols = (uint32_t)(NCAT ? 2 * P.BN : P.BN);
    int acc = 0, sci = 0;                                  // sci: scale/shift staging slot, alternates per item
    uint32_t accph = 0;
    for (int item = it0; item < items; item += itstep) {
      const int split = item % P.splits, tile = item / P.splits;
      const int nt = tile % P.n_tiles, mt = (tile / P.n_tiles) * mstep + rank;
      const int x = (mt % P.tiles_x) * P.BW + bx, y = (mt / P.tiles_x) * P.BH + by;
      const bool valid = x < P.Wo && y < P.Ho;
      const int pix = (y * E.osy + E.ooy) * E.OWf + x * E.osx + E.oox;
      const int nbase = nt * P.BN;
      const bool use_res = P.splits == 1 && E.res_hi != nullptr && P.vec32 && !(P.debug & 2);
      if (P.splits == 1 && et < P.BN) {                    // this tile's scale / shift -> shared memory
        const int c = nbase + et;
        const bool in = c < E.Cout;
        epi_sc[sci][0][et] = (in && P.epi.scale) ? __ldg(P.epi.scale + c) : 1.f;
        epi_sc[sci][1][et] = (in && P.epi.shift) ? __ldg(P.epi.shift + c) : 0.f;
      }
      ResChunk rc{}, rc1{};
      if (use_res && valid) {
        if (nbase + cset * 32 + 32 <= E.Cout) load_res(E, pix, nbase + cset * 32, rc);
        if (cset * 32 + 64 < P.BN && nbase + cset * 32 + 96 <= E.Cout) load_res(E, pix, nbase + cset * 32 + 64, rc1);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");       // scale/shift visible to all epilogue warps
      if (warp == 2) TC_TRACE(6, item);
      if constexpr (two) {
        // Two chains = two K halves in TMEM buffers 0 and 1 (see the MMA issuer).  Buffer 0 is folded into registers and
        // released while the second half is still being accumulated; buffer 1 is added on top when it completes.  At most
        // two 32-channel chunks per thread (BN <= 128, guaranteed by the plan).
        const uint32_t tq = tmem_base + ((uint32_t)(quarter * 32) << 16);
        float part[2][32];
        mbar_wait(tfull0, accph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp == 2) TC_TRACE(7, item);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int cc = cset * 32 + j * 64;
          if (cc < P.BN && nbase + cc < P.Cout_pad) {
            tmem_ld32(tq + cc, part[j]);
            if (NCAT) {
              float xt[32];
              tmem_ld32(tq + P.BN + cc, xt);
#pragma unroll
              for (int i = 0; i < 32; ++i) part[j][i] += xt[i];
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0);
        mbar_wait(tfull0 + 8, accph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int cc = cset * 32 + j * 64;
          const int n0 = nbase + cc;
          if (cc < P.BN && n0 < P.Cout_pad) {
            float v[32];
            tmem_ld32(tq + acc_cols + cc, v);
            if (NCAT) {
              float xt[32];
              tmem_ld32(tq + acc_cols + P.BN + cc, xt);
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] += xt[i];
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += part[j][i];
            if (valid && !(P.debug & 1)) {
              if (P.splits > 1) {
                float4* dst = reinterpret_cast<float4*>(P.partial + ((size_t)split * npix + (size_t)y * P.Wo + x) * P.Cout_pad + n0);
#pragma unroll
                for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
              } else if (P.vec32 && n0 + 32 <= E.Cout) {
                ResChunk rc{};
                if (use_res) load_res(E, pix, n0, rc);
                epilogue_chunk32(E, pix, n0, v, rc, &epi_sc[sci][0][cc], &epi_sc[sci][1][cc]);
              } else {
#pragma unroll
                for (int g = 0; g < 4; ++g)
                  if (n0 + g * 8 < E.Cout) epilogue_store<8>(E, pix, n0 + g * 8, v + g * 8);
              }
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (warp == 2) TC_TRACE(8, item);
        sci ^= 1;
        if (lane == 0) mbar_arrive(tempty0 + 8);
        accph ^= 1;
      } else {
      ResChunk rc{}, rc1{};
      if (use_res && valid) {
        if (nbase + cset * 32 + 32 <= E.Cout) load_res(E, pix, nbase + cset * 32, rc);
        if (cset * 32 + 64 < P.BN && nbase + cset * 32 + 96 <= E.Cout) load_res(E, pix, nbase + cset * 32 + 64, rc1);
      }
      mbar_wait(tfull0 + 8 * acc, accph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (warp == 2) TC_TRACE(7, item);
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * acc_cols;
      for (int cc = cset * 32; cc < P.BN; cc += 64) {
        const int n0 = nbase + cc;
        if (n0 >= P.Cout_pad) break;
        float v[32];
        if (P.debug & 4) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 1.f;
        } else {
          tmem_ld32(taddr + cc, v);
          if (NCAT) {
            float v2[32];
            tmem_ld32(taddr + P.BN + cc, v2);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += v2[i];
          }
        }
        ResChunk rn{};
        const int nn = n0 + 128;
        if (use_res && valid && cc + 128 < P.BN && nn + 32 <= E.Cout) load_res(E, pix, nn, rn);
        if (valid && !(P.debug & 1)) {
          if (P.splits > 1) {
            float4* dst = reinterpret_cast<float4*>(P.partial + ((size_t)split * npix + (size_t)y * P.Wo + x) * P.Cout_pad + n0);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          } else if (P.vec32 && n0 + 32 <= E.Cout) {
            epilogue_chunk32(E, pix, n0, v, rc, &epi_sc[sci][0][cc], &epi_sc[sci][1][cc]);
          } else {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              if (n0 + g * 8 < E.Cout) epilogue_store<8>(E, pix, n0 + g * 8, v + g * 8);
          }
        }
        rc = rc1;
        rc1 = rn;
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (warp == 2) TC_TRACE(8, item);
      sci ^= 1;
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      if (++acc == 2) { acc = 0; accph ^= 1; }
      }
      if (FUSED) {
        // Deterministic in-kernel split-K tail (the threadFenceReduction pattern): every epilogue thread fences its
        // slab stores, one thread counts the tile's arrivals, and the CTA that arrives last sums all slabs in split
        // order -- the same order and operations as splitk_epilogue_kernel, so the result is bit-identical -- and
        // applies the epilogue.  Nobody waits for anybody: no deadlock with persistent CTAs.  s_last is rewritten
        // only after the next item's `bar.sync 1` pair, so every thread has read it by then.
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0) {
          const unsigned old = atomicAdd(P.counters + tile, 1u);
          const int last = old == (unsigned)(P.splits - 1);
          if (last) P.counters[tile] = 0u;                    // every split has arrived: ready for the next launch
          s_last = last;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (s_last) {
          __threadfence();
          if (valid) {
            const size_t slab = (size_t)npix * P.Cout_pad;
            for (int cc = cset * 32; cc < P.BN; cc += 64) {
              const int n0 = nbase + cc;
              if (n0 >= E.Cout) break;
              float v[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 0.f;
              const float* src = P.partial + ((size_t)y * P.Wo + x) * P.Cout_pad + n0;
              for (int sp = 0; sp < P.splits; ++sp, src += slab) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float4 t = __ldcg(reinterpret_cast<const float4*>(src) + q);
                  v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
                }
              }
#pragma unroll
              for (int g = 0; g < 4; ++g)
                if (n0 + g * 8 < E.Cout) epilogue_store<8>(E, pix, n0 + g * 8, v + g * 8);
            }
          }
        }
      }
    }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (P.mcast) cluster_sync_all();                         // nobody leaves while its partner may still write or signal it
  else __syncthreads();
  if (warp == 0) TC_TRACE(9, 0);
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)P.tmem_cols) : "memory");
  }
}


// ================================================================================================
// CTA-pair variant (tcgen05 cta_group::2): two CTAs of a cluster -- two SMs of one TPC -- share one
// 256-pixel x BN tile.  Each CTA stages its own 128 pixel rows of A and HALF of the BN weight rows;
// the pair's tensor cores read both halves, so per SM the weight traffic (L2 -> shared memory and
// shared memory -> tensor core) is halved, which is what bounds the single-CTA kernel on these
// hi/lo-plane operands.  The leader CTA (cluster rank 0) issues every MMA; TMA loads of both CTAs
// signal the LEADER's `full` barrier, MMA completion is multicast to both CTAs' `empty` / `tfull`
// barriers, and the peer's epilogue warps release the accumulator on the leader's `tempty` barrier.
// ================================================================================================
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait with cluster-scope acquire: the arrival comes from the peer CTA (release.cluster)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma2_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6, %7}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma2_commit(uint32_t bar) {      // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// Pair tile = 256 pixels x BN channels (BN <= 128).  Per CTA and stage: A_hi | A_lo of its own 128 pixels (32 KB) and its
// HALF of the weight rows, B_hi | B_lo (BN/2 rows each) -- 48 KB at BN = 128 against the single-CTA kernel's 64 KB, which is
// what that kernel is bound by (L2 -> shared memory saturates near 73 B/cycle/SM, tools/mma_probe.cu).  Three MMAs per K
// slice, M = 256, N = BN: hi*hi into accumulator columns [0, BN), hi*lo and lo*hi into [BN, 2BN) -- the same separation of
// the small cross terms as the N-concatenated single-CTA issue.  Accumulators are double-buffered
// (4*BN TMEM columns); TWO = two accumulation chains as K halves through the two buffers, exactly as in conv_tc_kernel.
template <bool TWO>
__global__ void __launch_bounds__(kThreadsPair, 1) conv_tc2_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 4];
  __shared__ __align__(8) uint64_t pbars[kMaxStages];           // leader's copy: the PEER's operands of stage s have landed
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float epi_sc[2][2][128];

  // operand ring per CTA: [stage][A_hi | A_lo | B_hi half | B_lo half]
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = BM * 128u, b_bytes = (uint32_t)(P.BN / 2) * 128u;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kMaxStages]);
  const uint32_t tfull0 = smem_u32(&bars[2 * kMaxStages]), tempty0 = smem_u32(&bars[2 * kMaxStages + 2]);
  const uint32_t pfull0 = smem_u32(&pbars[0]);

  // Every CTA's TMA loads complete on its OWN `full` barrier; the peer's relay warp (its idle MMA warp) forwards each
  // completion with ONE remote arrival on the leader's `pfull`.  (Signalling the leader's barrier straight from the peer's
  // TMA loads -- cta_group::2 loads with a remote mbarrier -- measured 1750 cycles per 48 KB stage with nothing else running.)
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
      mbar_init(pfull0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 2 * kEpiWarps);    // (leader's copy) the epilogue warps of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {                                  // one warp of EACH CTA takes part in the pair allocation
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)P.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  pdl_trigger();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                               // barriers of both CTAs are initialised before anyone signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_wait();
  const uint32_t tmem_base = tmem_base_slot;

  const int tiles_m = P.tiles_x * P.tiles_y;
  const int tiles_m2 = (tiles_m + 1) / 2;
  const int items = tiles_m2 * P.n_tiles;           // (the plan never splits K for pair tiles)
  const int cid = (int)cluster_idx(), ncl = (int)cluster_count();
  const uint32_t acc_cols = 2u * (uint32_t)P.BN;

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs; converged warp, elect.sync) ==========
    int s = 0;
    uint32_t ph = 0;
    for (int item = cid; item < items; item += ncl) {
      const int nt = item % P.n_tiles, mt = (item / P.n_tiles) * 2 + (int)rank;       // phantom tile past the end: all OOB
      const int x0 = (mt % P.tiles_x) * P.BW, y0 = (mt / P.tiles_x) * P.BH;
      const int fr = y0 / P.Hof, yl = y0 - fr * P.Hof;
      const int nrow = nt * P.BN + (int)rank * (P.BN / 2);
      int t = 0, kc = 0;
      for (int it = 0; it < P.kiters; ++it) {
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const uint32_t fb = full0 + 8 * s;
        const uint32_t sa = smem0 + s * stage_bytes;
        const int dy = P.dy[t], dx = P.dx[t];
        const int kcol = t * P.Cin_pad + kc * BK;
        if (elect_one()) {
          if (P.debug & 256) {                                           // timing decomposition: no operand loads
            mbar_arrive(fb);
          } else {
            const bool ldA = !(P.debug & 512), ldB = !(P.debug & 1024);   // timing decomposition only
            mbar_arrive_expect_tx(fb, (ldA ? 2 * a_bytes : 0u) + (ldB ? 2 * b_bytes : 0u));
            if (!ldA) {
            } else if (P.stride2) {
              tma_load_5d(sa, &P.a_hi, fb, kc * BK, dx & 1, x0 + (dx >> 1), dy & 1, y0 + (dy >> 1));
              tma_load_5d(sa + a_bytes, &P.a_lo, fb, kc * BK, dx & 1, x0 + (dx >> 1), dy & 1, y0 + (dy >> 1));
            } else {
              tma_load_4d(sa, &P.a_hi, fb, kc * BK, x0 + dx, yl + dy, fr);
              tma_load_4d(sa + a_bytes, &P.a_lo, fb, kc * BK, x0 + dx, yl + dy, fr);
            }
            if (ldB) {
              tma_load_2d(sa + 2 * a_bytes, &P.b_hi, fb, kcol, nrow);
              tma_load_2d(sa + 2 * a_bytes + b_bytes, &P.b_lo, fb, kcol, nrow);
            }
          }
        }
        __syncwarp();
        if (++s == P.stages) { s = 0; ph ^= 1; }
        if (++kc == P.chunks) { kc = 0; ++t; }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only; converged warp) ==================
    if (leader) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(P.BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t bn = (uint32_t)P.BN;
      const bool no_mma = (P.debug & 128) != 0;
      int s = 0, acc = 0;
      uint32_t ph = 0, accph = 0;
      for (int item = cid; item < items; item += ncl) {
        const int kmid = TWO ? (P.kiters + 1) / 2 : P.kiters;
        for (int half = 0; half < (TWO ? 2 : 1); ++half) {
          const int hb = half ? kmid : 0, he = half ? P.kiters : kmid;
          mbar_wait(tempty0 + 8 * acc, accph ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d = tmem_base + (uint32_t)acc * acc_cols;
          for (int it = hb; it < he; ++it) {
            mbar_wait(full0 + 8 * s, ph);
            mbar_wait_cluster(pfull0 + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem0 + s * stage_bytes;
            const uint64_t ah = umma_desc(sa), al = umma_desc(sa + a_bytes);
            const uint64_t bh = umma_desc(sa + 2 * a_bytes), bl = umma_desc(sa + 2 * a_bytes + b_bytes);
            const uint32_t first0 = it > hb ? 1u : 0u;
            if (elect_one()) {
              if (!no_mma) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  const uint64_t adv = (uint64_t)(k * 2);
                  const uint32_t first = k >= 1 ? 1u : first0;
                  umma2_f16(d, ah + adv, bh + adv, idesc, first);          // hi*hi          -> [0, BN)
                  umma2_f16(d + bn, ah + adv, bl + adv, idesc, first);     // hi*lo          -> [BN, 2BN)
                  umma2_f16(d + bn, al + adv, bh + adv, idesc, 1u);        // lo*hi on top
                }
              }
              umma2_commit(empty0 + 8 * s);                   // frees the stage in both CTAs
            }
            __syncwarp();
            if (++s == P.stages) { s = 0; ph ^= 1; }
          }
          if (elect_one()) umma2_commit(tfull0 + 8 * acc);    // accumulator complete -> both CTAs' epilogues
          __syncwarp();
          if (++acc == 2) { acc = 0; accph ^= 1; }
        }
      }
    } else {
      // relay: this CTA's operands of stage s have landed -> one arrival on the leader's pfull[s]
      const uint32_t pf_leader0 = mapa_u32(pfull0, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int item = cid; item < items; item += ncl)
        for (int it = 0; it < P.kiters; ++it) {
          mbar_wait(full0 + 8 * s, ph);
          if (elect_one()) mbar_arrive_cluster(pf_leader0 + 8 * s);
          __syncwarp();
          if (++s == P.stages) { s = 0; ph ^= 1; }
        }
    }
  } else {
    // ===================================== epilogue (each CTA: its own 128 pixel rows) =====================
    const int quarter = warp & 3;
    const int cset = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    const int by = r / P.BW, bx = r - by * P.BW;
    const int et = threadIdx.x - 64;
    const Epilogue& E = P.epi;
    const uint32_t tempty_leader0 = mapa_u32(tempty0, 0);
    const uint32_t tq = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int acc = 0, sci = 0;
    uint32_t accph = 0;
    for (int item = cid; item < items; item += ncl) {
      const int nt = item % P.n_tiles, mt = (item / P.n_tiles) * 2 + (int)rank;
      const int x = (mt % P.tiles_x) * P.BW + bx, y = (mt / P.tiles_x) * P.BH + by;
      const bool valid = mt < tiles_m && x < P.Wo && y < P.Ho;
      const int pix = (y * E.osy + E.ooy) * E.OWf + x * E.osx + E.oox;
      const int nbase = nt * P.BN;
      const bool use_res = E.res_hi != nullptr && P.vec32 && valid;
      if (et < P.BN) {
        const int c = nbase + et;
        const bool in = c < E.Cout;
        epi_sc[sci][0][et] = (in && E.scale) ? __ldg(E.scale + c) : 1.f;
        epi_sc[sci][1][et] = (in && E.shift) ? __ldg(E.shift + c) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float part[2][32];
      if (TWO) {
        mbar_wait(tfull0 + 8 * acc, accph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int cc = cset * 32 + j * 64;
          if (cc < P.BN && nbase + cc < P.Cout_pad) {
            float xt[32];
            tmem_ld32(tq + (uint32_t)acc * acc_cols + cc, part[j]);
            tmem_ld32(tq + (uint32_t)acc * acc_cols + P.BN + cc, xt);
#pragma unroll
            for (int i = 0; i < 32; ++i) part[j][i] += xt[i];
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader0 + 8 * acc);
        if (++acc == 2) { acc = 0; accph ^= 1; }
      }
      mbar_wait(tfull0 + 8 * acc, accph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int cc = cset * 32 + j * 64;
        const int n0 = nbase + cc;
        if (cc < P.BN && n0 < P.Cout_pad) {
          float v[32];
          {
            float xt[32];
            tmem_ld32(tq + (uint32_t)acc * acc_cols + cc, v);
            tmem_ld32(tq + (uint32_t)acc * acc_cols + P.BN + cc, xt);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += xt[i];
          }
          if (TWO) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += part[j][i];
          }
          if (valid && !(P.debug & 1)) {
            if (P.vec32 && n0 + 32 <= E.Cout) {
              ResChunk rc{};
              if (use_res) load_res(E, pix, n0, rc);
              epilogue_chunk32(E, pix, n0, v, rc, &epi_sc[sci][0][cc], &epi_sc[sci][1][cc]);
            } else {
#pragma unroll
              for (int g = 0; g < 4; ++g)
                if (n0 + g * 8 < E.Cout) epilogue_store<8>(E, pix, n0 + g * 8, v + g * 8);
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      sci ^= 1;
      if (lane == 0) mbar_arrive_cluster(tempty_leader0 + 8 * acc);     // the leader's MMA thread owns this wait
      if (++acc == 2) { acc = 0; accph ^= 1; }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                               // nobody leaves while its partner may still signal or read it
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)P.tmem_cols) : "memory");
  }
}

}  // namespace

struct TcPlan {
  TcParams p;
  int grid;
  size_t smem;
  size_t partial_bytes;
  int launches;
};

bool tc_supported(const ConvParams& P) {
  if (P.stride != 1 && P.stride != 2) return false;
  if (P.stride == 2 && ((P.Hin & 1) || (P.Win & 1))) return false;
  if (P.in_ld % 8 || P.Cin_pad % BK || P.Kpad % 8) return false;
  if (((uintptr_t)P.in_hi & 15) || ((uintptr_t)P.in_lo & 15)) return false;
  return true;
}

// Can `C.nb` frames share one launch?  Tiles are BH x BW boxes of the frames stacked along H: no tile may straddle two
// frames, and a stride-2 layer (5-D view, frames folded into H/2) must not reach outside its frame vertically.
bool tc_batchable(const ConvParams& C) {
  if (C.nb <= 1) return true;
  if (!tc_supported(C)) return false;
  int bw = 1;
  while (bw * 2 <= C.Wo && bw * 2 <= BM) bw *= 2;
  const int bh = BM / bw;
  if (C.Ho % bh) return false;
  if (C.stride == 2) {
    for (int t = 0; t < C.ntaps; ++t)
      if (C.dy[t] < 0 || (C.dy[t] >> 1) + C.Ho > C.Hin / 2) return false;
  }
  return true;
}

TcPlan* tc_plan_create(const ConvParams& C, int num_sms, char* err, int errlen) {
  if (!tc_supported(C)) {
    snprintf(err, errlen, "shape not supported by the tcgen05 engine");
    return nullptr;
  }
  TcPlan* plan = new TcPlan();
  TcParams& P = plan->p;
  memset(&P, 0, sizeof(P));
  P.epi = C.epi;
  P.ntaps = C.ntaps;
  memcpy(P.dy, C.dy, sizeof(P.dy));
  memcpy(P.dx, C.dx, sizeof(P.dx));
  P.stride2 = C.stride == 2;
  P.Cin_pad = C.Cin_pad;
  const int nb = C.nb > 1 ? C.nb : 1;
  const int HoT = C.Ho * nb;                    // frames stacked along H: the loop space is simply taller
  P.nb = nb; P.Hof = C.Ho;
  P.Ho = HoT; P.Wo = C.Wo; P.Cout_pad = C.Cout_pad;
  P.chunks = (C.Cin + BK - 1) / BK;            // channel chunks that hold real data (padding chunks are all-zero)
  P.kiters = P.ntaps * P.chunks;

  // spatial box: BW x BH = 128 pixels, BW the largest power of two <= min(Wo, 128)
  int bw = 1;
  while (bw * 2 <= C.Wo && bw * 2 <= BM) bw *= 2;
  P.BW = bw; P.BH = BM / bw;
  P.tiles_x = (C.Wo + P.BW - 1) / P.BW;
  P.tiles_y = (HoT + P.BH - 1) / P.BH;
  // Tile width / split-K from a small cost model (cycles per CTA).  Measured on B200 (tools/bench_layer.py):
  // the fp16x3 mainloop is bound by shared-memory bandwidth -- per 64-channel K step the tensor core reads
  // 12 x (4 KB of A + 32*BN bytes of B) and TMA writes 32 KB + 256*BN bytes, at ~88 B/cycle effective --
  // so wide tiles are cheaper per output channel, but only if the tile count still fills the 148 SMs.
  const int tiles_m = P.tiles_x * P.tiles_y;
  int best_bn = 64, best_splits = 1;
  double best_cost = 1e30;
  const int split_opts[10] = {1, 2, 3, 4, 5, 6, 8, 9, 12, 16};
  // Accuracy first: a 256-wide tile has no room in TMEM for the N-concatenated layout, so its three
  // MMAs per K slice (hi*hi, hi*lo, lo*hi) all add into ONE accumulator and the small cross terms are truncated at the big
  // sum's ulp; measured at 1024x2048 (tools/interval_parity.py) that alone put Accel-101's score error at 0.8-1.5e-3
  // against 2.4-4.0e-4 with every long-K layer on N <= 128.  Layers that walk >= ACCEL_TC_WIDE_KMAX (16) K stages
  // therefore never take BN = 256.
  const int wide_kmax = env_int("ACCEL_TC_WIDE_KMAX", 16);
  for (int bn = 256; bn >= 64; bn >>= 1) {
    if (bn > C.Cout_pad && bn != 64) continue;
    if (bn == 256 && P.kiters >= wide_kmax) continue;
    const int n_tiles = (C.Cout_pad + bn - 1) / bn;
    const int tiles = tiles_m * n_tiles;
    const double per_k = (12.0 * (4096.0 + 32.0 * bn) + 32768.0 + 256.0 * bn) / 88.0;
    for (int si = 0; si < 10; ++si) {
      const int sp = split_opts[si];
      const int kps = (P.kiters + sp - 1) / sp;
      if (sp > 1 && kps < 4) continue;
      const long long work = (long long)tiles * sp;
      const int waves = (int)((work + num_sms - 1) / num_sms);
      double cost = (double)waves * (kps * per_k + 10.0 * bn) + 6000.0;
      if (sp > 1) cost += 8000.0 + 2.0 * sp * (double)HoT * C.Wo * C.Cout_pad * 4.0 / 3000.0;
      if (cost < best_cost) { best_cost = cost; best_bn = bn; best_splits = sp; }
    }
  }
  // Short-K layers with wide outputs (the 1x1 expand convs of the bottleneck blocks and their shortcut convs) are
  // bound by the epilogue, not the mainloop: measured (profiles/r01_layer_tma_epilogue.txt) the TMA-both-ways
  // epilogue at BN = 128 beats the direct epilogue at BN = 256 by 14-28 % there, and loses elsewhere.
  const int tma_mode = env_int("ACCEL_TC_TMA_OUT", -1);          // -1 auto, 0 never, 1 wherever it fits
  const int t_bn = best_bn > 128 ? 128 : best_bn;
  const bool auto_t = tma_mode < 0 && best_splits == 1 && P.kiters <= env_int("ACCEL_TC_TMA_KMAX", 10) && C.epi.Cout % t_bn == 0 &&
                      C.epi.out_hi != nullptr && (!C.epi.out2_hi || env_int("ACCEL_TC_TMA_OUT2", 1) != 0);
  if (auto_t) best_bn = t_bn;
  int bn = env_int("ACCEL_TC_BN", best_bn);
  if (bn != 64 && bn != 128 && bn != 256) bn = best_bn;
  // CTA pairs (conv_tc2_kernel): 256-pixel x 128-channel tiles, each SM stages half of the weight rows.  Pays where the
  // single-CTA mainloop is bound by its 64 KB of operands per K stage, i.e. long-K layers that are not split:
  // ACCEL_TC_PAIR = 0 never, 1 wherever legal, unset = auto (at least ACCEL_TC_PAIR_KMIN K stages, >= 128 output channels,
  // enough pair tiles for the SM pairs).
  {
    const int mode = env_int("ACCEL_TC_PAIR", 0);       // measured slower than the single-CTA kernel: opt-in
    const int pbn = C.Cout_pad >= 128 ? 128 : 64;
    const int pitems = ((tiles_m + 1) / 2) * ((C.Cout_pad + pbn - 1) / pbn);
    const bool legal = C.Cout_pad >= 64 && (P.stride2 ? nb == 1 : true);
    const bool want_pair = mode == 1 || (mode < 0 && best_splits == 1 && !auto_t && P.kiters >= env_int("ACCEL_TC_PAIR_KMIN", 12) &&
                                         C.Cout_pad >= 128 && pitems >= (num_sms / 2) * env_int("ACCEL_TC_PAIR_WAVES", 2));
    P.pair = (legal && want_pair && mode != 0) ? 1 : 0;
    if (P.pair) bn = pbn;
  }
  P.BN = bn;
  P.n_tiles = (C.Cout_pad + bn - 1) / bn;
  int splits = P.pair ? 1 : env_int("ACCEL_TC_SPLITS", bn == best_bn ? best_splits : 1);
  if (splits > P.kiters) splits = P.kiters;
  if (splits < 1) splits = 1;
  const size_t stage_bytes = 2 * (size_t)BM * 128 + (P.pair ? 1 : 2) * (size_t)bn * 128;   // pair: half the weight rows per CTA
  // TMA-both-ways epilogue: needs kStageOut bytes of staging next to the operand ring, so only when at least two
  // ring stages still fit (BN <= 128).
  const bool want_stage = !P.pair && (tma_mode > 0 || (auto_t && bn == best_bn)) && splits == 1 && C.epi.out_hi != nullptr &&
                          (kSmemBudget - kStageOut) / stage_bytes >= 2;
  // Half-size staging (ONE chunk set, TcParams::epi1, ACCEL_TC_EPI1=1) buys a third operand stage: main split output
  // (+ residual) only.  229,376 bytes of dynamic shared memory + the kernel's 3 KB
  // of static = exactly the 227 KB a block may have; the ring is 1024-aligned by the array's own alignment (no slack).
  const Epilogue& E0 = C.epi;
  const int epi1_mode = env_int("ACCEL_TC_EPI1", -1);
  const bool epi1 = want_stage && stage_bytes == 65536 && !E0.out2_hi && !E0.out_nchw && !E0.raw_nchw && !C.ext_outputs &&
                    epi1_mode == 1;      // MEASURED SLOWER (one chunk set halves the epilogue's throughput: res4 expand x5 frames
                                         // 92 -> 129 us, profiles/r02_layer_epi1.txt): opt-in only
  P.epi1 = epi1 ? 1 : 0;
  int stages = epi1 ? 3 : (int)((kSmemBudget - (want_stage ? kStageOut : 0)) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  P.stages = stages;
  P.ring_bytes = (unsigned)(stages * stage_bytes);
  // A-slab reuse (TcParams::aslab): stride 1, one-row tiles (BW = 128), taps on a regular ndy x ndx grid.
  // MEASURED SLOWER and therefore off unless ACCEL_TC_ASLAB=1 (profiles/r02_layer_slab_sweep.txt): a third less operand
  // traffic from L2, yet res4 3x3 35 -> 44 us -- the mainloop is bound by shared-memory bandwidth (tensor-core operand
  // reads + TMA writes), and row-shifted operands cost the tensor core more shared-memory wavefronts than they save.
  P.aslab = 0;
  P.sa_stages = 1; P.slab_pl = 0; P.ndx = 1; P.dxmin = 0; P.slab_w = 0;
  if (!P.stride2 && !P.pair && P.BH == 1 && P.ntaps > 1 && env_int("ACCEL_TC_ASLAB", 0) != 0) {
    int ndx = 1;
    while (ndx < P.ntaps && P.dy[ndx] == P.dy[0]) ++ndx;
    bool grid_ok = ndx >= 2 && P.ntaps % ndx == 0;
    int dxmin = P.dx[0], dxmax = P.dx[0];
    for (int i = 0; grid_ok && i < ndx; ++i) { dxmin = std::min(dxmin, (int)P.dx[i]); dxmax = std::max(dxmax, (int)P.dx[i]); }
    for (int t = 0; grid_ok && t < P.ntaps; ++t)
      grid_ok = P.dy[t] == P.dy[(t / ndx) * ndx] && P.dx[t] == P.dx[t % ndx];
    const int slab_w = P.BW + dxmax - dxmin;
    const size_t slab_pl = ((size_t)slab_w * 128 + 1023) / 1024 * 1024;
    int sa = env_int("ACCEL_TC_ASLAB_SA", 2);
    if (sa < 1 || sa > 4) sa = 2;
    const size_t b_stage = 2 * (size_t)bn * 128;
    const long long left = (long long)kSmemBudget - (want_stage ? kStageOut : 0) - (long long)sa * 2 * slab_pl;
    int sb = left > 0 ? (int)(left / (long long)b_stage) : 0;
    if (sb > kMaxStages) sb = kMaxStages;
    if (grid_ok && slab_w <= 256 && sb >= 3) {
      P.slab_bo = env_int("ACCEL_TC_ASLAB_BO", 0);
      P.aslab = 1; P.ndx = ndx; P.dxmin = dxmin; P.slab_w = slab_w; P.slab_pl = (int)slab_pl; P.sa_stages = sa;
      P.stages = stages = sb;
      P.ring_bytes = (unsigned)(sa * 2 * slab_pl + sb * b_stage);
    }
  }
  plan->smem = P.epi1 ? P.ring_bytes + kStageOut / 2 : P.ring_bytes + 1024 + (want_stage ? kStageOut : 0);
  // hi*hi and hi*lo as one N = 2*BN MMA (see the MMA issuer): needs 4*BN TMEM columns, i.e. BN <= 128.
  // ACCEL_TC_NCAT: -1 auto (on whenever it fits), 0 off, 1 on.
  {
    const int ncat_mode = env_int("ACCEL_TC_NCAT", -1);
    P.ncat = (!P.pair && bn <= 128 && ncat_mode != 0) ? 1 : 0;
  }
  int cols = 32;
  while (cols < ((P.ncat || P.pair) ? 4 : 2) * bn) cols *= 2;   // pair: [hi*hi | cross terms] x two buffers
  P.tmem_cols = cols;

  const int tiles = tiles_m * P.n_tiles;
  P.splits = splits;
  plan->partial_bytes = splits > 1 ? (size_t)splits * HoT * C.Wo * C.Cout_pad * sizeof(float) : 0;
  P.fused = (splits > 1 && !P.pair && env_int("ACCEL_TC_FUSED_SPLITK", 0) != 0) ? 1 : 0;
  if (P.fused) plan->partial_bytes += (size_t)tiles * sizeof(unsigned);          // arrival counters behind the slabs
  int items = tiles * splits;
  plan->grid = items < num_sms ? items : num_sms;
  if (P.pair) {
    items = ((tiles_m + 1) / 2) * P.n_tiles * splits;
    const int ncl = items < num_sms / 2 ? items : num_sms / 2;
    plan->grid = 2 * ncl;
  }
  // B multicast between the two CTAs of a cluster (TcParams::mcast).  ACCEL_TC_MCAST: 0 never, 1 wherever legal,
  // unset = auto (at least ACCEL_TC_MCAST_KMIN K stages per tile).
  {
    const int mode = env_int("ACCEL_TC_MCAST", 0);
    const bool legal = !P.pair && !P.aslab && !P.epi1 && splits == 1 && tiles_m >= 2 && num_sms >= 2;
    const bool want_mc = mode == 1 || (mode < 0 && P.kiters >= env_int("ACCEL_TC_MCAST_KMIN", 8));
    P.mcast = (legal && want_mc) ? 1 : 0;
    if (P.mcast) {
      const int pitems = ((tiles_m + 1) / 2) * P.n_tiles;
      const int ncl = pitems < num_sms / 2 ? pitems : num_sms / 2;
      plan->grid = 2 * ncl;
    }
  }
  plan->launches = (splits > 1 && !P.fused) ? 2 : 1;
  {
    // Long K chains: split the accumulation over both TMEM buffers (TcParams::kchains).  ACCEL_TC_CHAINS: 0/1 never,
    // 2 always, unset = when a split walks at least ACCEL_TC_CHAINS_MIN (6) K stages, i.e. 24 slices of 16.
    const int mode = env_int("ACCEL_TC_CHAINS", -1);
    const int kps = (P.kiters + splits - 1) / splits;
    // ACCEL_TC_CHAINS_MULTI=0: only where every CTA holds a single item (nothing to overlap an epilogue with anyway)
    const bool single_wave = (long long)tiles * splits <= (long long)num_sms;
    const bool want = mode == 2 || (mode < 0 && kps >= env_int("ACCEL_TC_CHAINS_MIN", 6) &&
                                    (single_wave || kps >= env_int("ACCEL_TC_CHAINS_MULTI_MIN", 16) ||
                                     env_int("ACCEL_TC_CHAINS_MULTI", 0) != 0));
    // (the epilogue folds the first half into registers: at most two 32-channel chunks per thread, i.e. BN <= 128, and
    // at least one K stage per half; the TMA epilogue -- short-K layers -- never combines with it, see below)
    P.kchains = (want && mode != 0 && mode != 1 && bn <= 128 && kps >= 2 && P.kiters / splits >= 2) ? 2 : 1;
  }
  {
    auto al32 = [](const void* p) { return ((uintptr_t)p & 31) == 0; };
    const Epilogue& E = C.epi;
    bool v = true;
    if (E.out_hi) v = v && al32(E.out_hi) && al32(E.out_lo) && E.out_ld % 16 == 0;
    if (E.res_hi) v = v && al32(E.res_hi) && al32(E.res_lo) && E.res_ld % 16 == 0;
    if (E.out2_hi) v = v && al32(E.out2_hi) && al32(E.out2_lo) && E.out2_ld % 16 == 0;
    P.vec32 = v ? 1 : 0;
    P.krot = (P.aslab || P.mcast) ? 0 : env_int("ACCEL_TC_KROT", P.BN == 256 ? 1 : 0);
    P.prefetch = (P.aslab || P.pair) ? 0 : env_int("ACCEL_TC_PREFETCH", 0);
    P.res_ahead = env_int("ACCEL_TC_RES_AHEAD", 2);   // measured: 2 chunks -9 % on res2 expand, larger distances lose (profiles/r02_layer_res_ahead.txt)
    P.debug = env_int("ACCEL_TC_DEBUG", 0);
  }

  // tensor maps ------------------------------------------------------------------------------------
  bool ok = true;
  const cuuint64_t e = sizeof(__half);
  if (!P.stride2) {
    // (channel, x, y, frame): out-of-image rows / columns are zero-filled per frame, so the conv padding stays right
    // when several frames share one launch
    cuuint64_t dims[4] = {(cuuint64_t)C.Cin, (cuuint64_t)C.Win, (cuuint64_t)C.Hin, (cuuint64_t)nb};
    cuuint64_t str[3] = {(cuuint64_t)C.in_ld * e, (cuuint64_t)C.in_ld * C.Win * e, (cuuint64_t)C.in_ld * C.Win * C.Hin * e};
    cuuint32_t box[4] = {BK, (cuuint32_t)(P.aslab ? P.slab_w : P.BW), (cuuint32_t)P.BH, 1};
    ok = ok && encode(&P.a_hi, C.in_hi, 4, dims, str, box, err, errlen);
    ok = ok && encode(&P.a_lo, C.in_lo, 4, dims, str, box, err, errlen);
  } else {
    // stride 2: [H/2][2][W/2][2][C]; the frames fold into the outermost dimension (tc_batchable() only admits layers
    // whose taps never leave the frame vertically: the 1x1 / stride-2 convs of the bottleneck nets)
    cuuint64_t dims[5] = {(cuuint64_t)C.Cin, 2, (cuuint64_t)C.Win / 2, 2, (cuuint64_t)C.Hin / 2 * nb};
    cuuint64_t str[4] = {(cuuint64_t)C.in_ld * e, 2 * (cuuint64_t)C.in_ld * e, (cuuint64_t)C.in_ld * C.Win * e,
                         2 * (cuuint64_t)C.in_ld * C.Win * e};
    cuuint32_t box[5] = {BK, 1, (cuuint32_t)P.BW, 1, (cuuint32_t)P.BH};
    ok = ok && encode(&P.a_hi, C.in_hi, 5, dims, str, box, err, errlen);
    ok = ok && encode(&P.a_lo, C.in_lo, 5, dims, str, box, err, errlen);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)C.Kpad, (cuuint64_t)C.Cout_pad};
    cuuint64_t str[1] = {(cuuint64_t)C.Kpad * e};
    cuuint32_t box[2] = {BK, (cuuint32_t)(P.pair ? bn / 2 : bn)};
    ok = ok && encode(&P.b_hi, C.w_hi, 2, dims, str, box, err, errlen);
    ok = ok && encode(&P.b_lo, C.w_lo, 2, dims, str, box, err, errlen);
  }
  {
    // main output as TMA stores: channel chunks of 32 (64 bytes), the tile's BW x BH pixel box
    const Epilogue& E = C.epi;
    const bool can = want_stage && !(P.debug & 8) && P.splits == 1 && P.vec32 && E.out_hi && E.out_ld % 8 == 0 &&
                     E.Cout % P.BN == 0 && (!E.res_hi || (E.osy == 1 && E.res_ld % 8 == 0));
    P.tma_out = 0;
    if (can && E.osy == 1 && E.osx == 1 && E.ooy == 0 && E.oox == 0) {
      cuuint64_t dims[3] = {(cuuint64_t)E.Cout, (cuuint64_t)E.OWf, (cuuint64_t)E.OHf};
      cuuint64_t str[2] = {(cuuint64_t)E.out_ld * e, (cuuint64_t)E.out_ld * E.OWf * e};
      cuuint32_t box[3] = {32, (cuuint32_t)P.BW, (cuuint32_t)P.BH};
      ok = ok && encode(&P.o_hi, E.out_hi, 3, dims, str, box, err, errlen, CU_TENSOR_MAP_SWIZZLE_64B);
      ok = ok && encode(&P.o_lo, E.out_lo, 3, dims, str, box, err, errlen, CU_TENSOR_MAP_SWIZZLE_64B);
      if (E.res_hi) {
        cuuint64_t rstr[2] = {(cuuint64_t)E.res_ld * e, (cuuint64_t)E.res_ld * E.OWf * e};
        ok = ok && encode(&P.r_hi, E.res_hi, 3, dims, rstr, box, err, errlen, CU_TENSOR_MAP_SWIZZLE_64B);
        ok = ok && encode(&P.r_lo, E.res_lo, 3, dims, rstr, box, err, errlen, CU_TENSOR_MAP_SWIZZLE_64B);
      }
      P.tma_out = 1;
    } else if (can && !E.res_hi && E.osy == 2 && E.osx == 2 && E.OWf % 2 == 0 && E.OHf % 2 == 0) {
      cuuint64_t dims[5] = {(cuuint64_t)E.Cout, 2, (cuuint64_t)E.OWf / 2, 2, (cuuint64_t)E.OHf / 2};
      cuuint64_t str[4] = {(cuuint64_t)E.out_ld * e, 2 * (cuuint64_t)E.out_ld * e, (cuuint64_t)E.out_ld * E.OWf * e,
                           2 * (cuuint64_t)E.out_ld * E.OWf * e};
      cuuint32_t box[5] = {32, 1, (cuuint32_t)P.BW, 1, (cuuint32_t)P.BH};
      ok = ok && encode(&P.o_hi, E.out_hi, 5, dims, str, box, err, errlen, CU_TENSOR_MAP_SWIZZLE_64B);
      ok = ok && encode(&P.o_lo, E.out_lo, 5, dims, str, box, err, errlen, CU_TENSOR_MAP_SWIZZLE_64B);
      P.tma_out = 2;
    }
  }
  if (P.tma_out) P.kchains = 1;
  if (P.epi1 && !P.tma_out) {                                  // the output did not qualify for TMA stores after all: the direct
    P.epi1 = 0;                                                //  epilogue needs no staging, the three stages stay
    plan->smem = P.ring_bytes + 1024;
  }
  // measured (profiles/r02_layer_epiw16.txt, five frames batched): res2 shortcut -16 %, res3 expand -9 %, res4 expand -2 %,
  // res5 expand (8 K stages) +4 % -> only where the tile's mainloop is at most ACCEL_TC_EPIW16_KMAX (4) K stages
  {
    const int m = env_int("ACCEL_TC_EPIW16", -1);
    P.epiw16 = (P.tma_out && !P.epi1 && !P.mcast && !P.pair && (m == 1 || (m < 0 && P.kiters <= env_int("ACCEL_TC_EPIW16_KMAX", 4)))) ? 1 : 0;
  }
  if (!ok) {
    delete plan;
    return nullptr;
  }
  if (first_time_on_device(ONCE_CONV_TC)) {
    cudaError_t ce = cudaSuccess;
    const void* fns[8] = {(const void*)conv_tc_kernel<false, false, false>, (const void*)conv_tc_kernel<true, false, false>,
                          (const void*)conv_tc_kernel<false, true, false>,  (const void*)conv_tc_kernel<true, true, false>,
                          (const void*)conv_tc_kernel<false, false, true>,  (const void*)conv_tc_kernel<true, false, true>,
                          (const void*)conv_tc_kernel<false, true, true>,   (const void*)conv_tc_kernel<true, true, true>};
    for (int i = 0; i < 8 && ce == cudaSuccess; ++i)
      ce = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMaxDynamic);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(conv_tc_kernel<false, false, false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMaxDynamic);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(conv_tc_kernel<true, false, false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMaxDynamic);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(conv_tc_kernel<false, false, false, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemEpi1);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(conv_tc_kernel<true, false, false, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemEpi1);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(conv_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMaxDynamic);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(conv_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMaxDynamic);
    if (ce != cudaSuccess) {
      snprintf(err, errlen, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce));
      delete plan;
      return nullptr;
    }
  }
  return plan;
}

// Prints the timeline CTA 0 recorded during the LAST launch with ACCEL_TC_DEBUG bit 2048 (cycles relative to the start tag)
// and clears it.  Tuning aid, called by accel_conv_layer when ACCEL_TC_TRACE is set.
void tc_trace_dump(FILE* f) {
  unsigned n = 0;
  if (cudaMemcpyFromSymbol(&n, g_tc_trace_n, sizeof(n)) != cudaSuccess) return;
  if (n > 2048u) n = 2048u;
  static long long host[3 * 2048];
  if (n && cudaMemcpyFromSymbol(host, g_tc_trace, sizeof(long long) * 3 * n) != cudaSuccess) return;
  long long t0 = 0;
  for (unsigned i = 0; i < n; ++i)
    if (host[3 * i] == 0) t0 = host[3 * i + 2];
  for (unsigned i = 0; i < n; ++i) fprintf(f, "TC_TRACE tag %lld item %lld t %lld\n", host[3 * i], host[3 * i + 1], host[3 * i + 2] - t0);
  n = 0;
  cudaMemcpyToSymbol(g_tc_trace_n, &n, sizeof(n));
}
void tc_trace_reset() {
  unsigned n = 0;
  cudaMemcpyToSymbol(g_tc_trace_n, &n, sizeof(n));
}

void tc_plan_destroy(TcPlan* plan) { delete plan; }
size_t tc_plan_partial_bytes(const TcPlan* plan) { return plan->partial_bytes; }
void tc_plan_set_partial(TcPlan* plan, float* partial) {
  TcParams& P = plan->p;
  P.partial = partial;
  if (P.fused) {
    const size_t tiles = (size_t)P.tiles_x * P.tiles_y * P.n_tiles;
    P.counters = reinterpret_cast<unsigned*>(partial + (size_t)P.splits * P.Ho * P.Wo * P.Cout_pad);
    cudaMemset(P.counters, 0, tiles * sizeof(unsigned));
  }
}
int tc_plan_launches(const TcPlan* plan) { return plan->launches; }

cudaError_t launch_conv_tc_ext(const TcPlan* plan, float* ext_nchw, float* ext_raw, cudaStream_t stream) {
  TcParams P = plan->p;
  if (ext_nchw) P.epi.out_nchw = ext_nchw;
  P.epi.raw_nchw = ext_raw;
  cudaError_t e;
  const dim3 grid(plan->grid), block(kThreads);
#define ACCEL_TC_LAUNCH(N_, F_, T_)                                                                             \
  (P.mcast ? launch_k_cluster(conv_tc_kernel<N_, F_, T_>, grid, block, plan->smem, stream, 2u, P)                \
           : launch_k(conv_tc_kernel<N_, F_, T_>, grid, block, plan->smem, stream, P))
  const bool two = P.kchains == 2;
  const bool w16 = P.epiw16 && P.tma_out && !two && !P.fused && !P.epi.out_nchw && !P.epi.raw_nchw && !P.epi.out2_hi;
  if (P.epi1) {
    if (two || P.fused || P.epi.out_nchw || P.epi.raw_nchw || P.epi.out2_hi) return cudaErrorInvalidValue;   // planned without them
    e = P.ncat ? launch_k(conv_tc_kernel<true, false, false, 8, 1>, grid, block, plan->smem, stream, P)
               : launch_k(conv_tc_kernel<false, false, false, 8, 1>, grid, block, plan->smem, stream, P);
  } else if (w16) {
    const dim3 block16(64 + 32 * 16 + 32 * kDmaWarps);
    e = P.ncat ? launch_k(conv_tc_kernel<true, false, false, 16>, grid, block16, plan->smem, stream, P)
               : launch_k(conv_tc_kernel<false, false, false, 16>, grid, block16, plan->smem, stream, P);
  } else if (P.pair) e = two ? launch_k_cluster(conv_tc2_kernel<true>, dim3(plan->grid), dim3(kThreadsPair), plan->smem, stream, 2u, P)
                      : launch_k_cluster(conv_tc2_kernel<false>, dim3(plan->grid), dim3(kThreadsPair), plan->smem, stream, 2u, P);
  else if (P.fused) e = P.ncat ? (two ? ACCEL_TC_LAUNCH(true, true, true) : ACCEL_TC_LAUNCH(true, true, false))
                               : (two ? ACCEL_TC_LAUNCH(false, true, true) : ACCEL_TC_LAUNCH(false, true, false));
  else e = P.ncat ? (two ? ACCEL_TC_LAUNCH(true, false, true) : ACCEL_TC_LAUNCH(true, false, false))
                  : (two ? ACCEL_TC_LAUNCH(false, false, true) : ACCEL_TC_LAUNCH(false, false, false));
#undef ACCEL_TC_LAUNCH
  if (e != cudaSuccess) return e;
  if (P.splits > 1 && !P.fused) return launch_splitk_epilogue(P.partial, P.splits, P.Ho * P.Wo, P.Cout_pad, P.Wo, P.epi, stream);
  return cudaSuccess;
}

}  // namespace accel
