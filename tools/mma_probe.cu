// tcgen05.mma issue-rate probe (sm_100a): what does ONE K=16 slice of the fp16x3 mainloop cost when nothing but the tensor
// core (and, optionally, a TMA-like stream of bulk copies into the same shared memory) is running?
//
// Every SM runs one CTA (cta_group::1) or every TPC one CTA pair (cta_group::2).  One elected thread issues `iters` K
// stages; a stage is 4 K=16 slices over a SWIZZLE_128B operand stage (like conv_tc_kernel's ring) and a slice is a fixed
// PATTERN of up to 4 MMAs {A offset, B offset, N, D column}.  Operands are random fp16 values (power draw matters: the
// clocks under load are part of the answer).  With `tma_kb` > 0 a second thread keeps `cp.async.bulk` copies of that many
// KB per MMA stage flowing from an L2-resident buffer into a scratch ring, paced by the MMA thread's stage counter.
// Output: cycles per slice (clock64 around the issue loop + final commit wait), averaged / min / max over CTAs.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/mma_probe tools/mma_probe.cu && tools/_build/mma_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

struct Mma { int a_off, b_off, n, dcol; };
struct Pattern {
  const char* name;
  int cg;            // cta_group
  int nmma;
  Mma m[4];
  int stage_bytes;   // operand bytes per stage per CTA (ring stride)
  int stages;
  int tma_kb;        // KB of bulk copies per stage per CTA (0: none)
};
struct Params {
  int nmma, stage_bytes, stages, iters, tma_bytes, scratch_off;
  Mma m[4];
  const uint8_t* src;
};

__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t a) {
  return (uint64_t)((a & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
template <int CG>
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint32_t bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ uint32_t ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

// STYLE 0: `if (lane == 0)` around the whole issue loop (conv_tc_kernel as of round 2: every operand of every MMA goes
//          through an ELECT / R2UR.BROADCAST waterfall because the compiler cannot prove it warp-uniform)
// STYLE 1: the whole warp walks the loop convergently, descriptors live in uniform registers, only the tcgen05
//          instructions are predicated on elect.sync
template <int CG, int NM, int STYLE>
__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ Params P, long long* out, long long* out_tma) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[8];      // 0: done, 1..2: tma ring, 3..6: stage commits (sink)
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int progress;              // MMA stages issued so far
  const uint32_t smem0 = (saddr(smem_raw) + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (smem0 - saddr(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // random small fp16 operands
  {
    const int total = P.stage_bytes * P.stages / 2;
    __half* h = reinterpret_cast<__half*>(base);
    uint32_t x = 1234567u + blockIdx.x * 7919u + threadIdx.x;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      x = x * 1664525u + 1013904223u;
      h[i] = __float2half(((int)(x >> 9) % 2001 - 1000) * 1e-3f);
    }
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(saddr(&bars[i]), i >= 3 ? (1u << 20) - 1u : 1u);
    progress = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const bool leader = CG == 1 || ctarank() == 0;
  const uint32_t done = saddr(&bars[0]);

  if (warp == 1 && leader && (STYLE == 1 || lane == 0)) {
    const int M = CG == 1 ? 128 : 256;
    uint32_t idesc[NM];
#pragma unroll
    for (int j = 0; j < NM; ++j) idesc[j] = (1u << 4) | ((uint32_t)(P.m[j].n >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const long long t0 = clock64();
    int s = 0;
    for (int it = 0; it < P.iters; ++it) {
      const uint32_t sb = smem0 + (uint32_t)s * (uint32_t)P.stage_bytes;
      const bool go = STYLE == 0 || elect_one();
      if (go) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int j = 0; j < NM; ++j)
            umma<CG>(tmem + (uint32_t)P.m[j].dcol, umma_desc(sb + P.m[j].a_off) + (uint64_t)(2 * k),
                     umma_desc(sb + P.m[j].b_off) + (uint64_t)(2 * k), idesc[j], (it > 0 || k > 0) ? 1u : 0u);
        }
        commit<CG>(saddr(&bars[3 + (s & 3)]));        // like the ring's `empty` commit (nobody waits on it)
        progress = it + 1;
      }
      if (STYLE == 1) __syncwarp();
      if (++s == P.stages) s = 0;
    }
    const long long t_issue = clock64();
    if (STYLE == 0 || elect_one()) commit<CG>(done);
    mbar_wait(done, 0);
    const long long t1 = clock64();
    if (lane == 0) {
      out[2 * blockIdx.x] = t1 - t0;
      out[2 * blockIdx.x + 1] = t_issue - t0;
    }
  } else if (warp == 2 && lane == 0 && P.tma_bytes > 0) {
    // bulk-copy stream: tma_bytes per MMA stage, at most 2 stages ahead of the MMA thread's progress counter
    const uint32_t scr = smem0 + (uint32_t)P.scratch_off;
    long long moved = 0;
    const long long tw0 = clock64();
    uint32_t ph[2] = {0, 0};
    int inflight = 0;
    const uint8_t* src = P.src + (size_t)(blockIdx.x % 64) * (512 * 1024);
    for (int it = 0; it < P.iters; ++it) {
      if (leader) while (progress + 3 < it) {}
      const int b = it & 1;
      if (inflight == 2) { mbar_wait(saddr(&bars[1 + b]), ph[b]); ph[b] ^= 1; --inflight; }
      mbar_expect(saddr(&bars[1 + b]), (uint32_t)P.tma_bytes);
      for (int off = 0; off < P.tma_bytes; off += 16384) {
        const int n = P.tma_bytes - off < 16384 ? P.tma_bytes - off : 16384;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         scr + (uint32_t)off),
                     "l"(src + ((size_t)it * 65536 + off) % (448 * 1024)), "r"(n), "r"(saddr(&bars[1 + b]))
                     : "memory");
      }
      ++inflight;
      moved += P.tma_bytes;
    }
    while (inflight > 0) { const int b = (P.iters - inflight) & 1; mbar_wait(saddr(&bars[1 + b]), ph[b]); ph[b] ^= 1; --inflight; }
    out_tma[blockIdx.x] = moved > 0 ? clock64() - tw0 : 0;
  } else if (CG == 2 && warp == 1 && lane == 0 && !leader) {
    mbar_wait(done, 0);                            // multicast commit: the peer sees completion too
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  else __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}


// ---- ring skeleton: the conv kernel's producer <-> MMA-issuer barrier protocol with nothing loaded ----------------------
// warp 0 = producer (waits empty[s], arrives full[s] with 0 bytes), warp 1 = MMA issuer (waits full[s], issues the NCAT
// slice pattern x4 or nothing, commits empty[s]), `pollers` further warps wait on a barrier that only completes at the end
// (what the epilogue warps do during a tile's mainloop).  poll_mode 0: all 32 lanes spin on try_wait; 1: lane 0 only;
// 2: lane 0 with a nanosleep back-off.
struct RingParams { int stages, iters, do_mma, pollers, poll_mode, stage_bytes; };

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(384, 1) ring_probe(const __grid_constant__ RingParams P, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[16];     // 0..5 full, 6..11 empty, 12 done
  __shared__ uint32_t tmem_slot;
  const uint32_t smem0 = (saddr(smem_raw) + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (smem0 - saddr(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int total = P.stage_bytes * P.stages / 2;
    __half* h = reinterpret_cast<__half*>(base);
    uint32_t x = 1234567u + blockIdx.x * 7919u + threadIdx.x;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      x = x * 1664525u + 1013904223u;
      h[i] = __float2half(((int)(x >> 9) % 2001 - 1000) * 1e-3f);
    }
  }
  const uint32_t full0 = saddr(&bars[0]), empty0 = saddr(&bars[6]), done = saddr(&bars[12]);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 13; ++i) mbar_init(saddr(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int A = 16 * 1024;
  if (warp == 0) {
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < P.iters; ++it) {
      mbar_wait(empty0 + 8 * s, ph ^ 1);
      if (elect_one()) mbar_expect(full0 + 8 * s, 0);
      __syncwarp();
      if (++s == P.stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const long long t0 = clock64();
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < P.iters; ++it) {
      mbar_wait(full0 + 8 * s, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sb = smem0 + (uint32_t)s * (uint32_t)P.stage_bytes;
      const uint64_t ah = umma_desc(sb), al = umma_desc(sb + A), bh = umma_desc(sb + 2 * A);
      if (elect_one()) {
        if (P.do_mma) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma<1>(tmem + (k & 1) * 256, ah + 2 * k, bh + 2 * k, idesc2, (it > 0 || k > 1) ? 1u : 0u);
            umma<1>(tmem + (k & 1) * 256 + 128, al + 2 * k, bh + 2 * k, idesc, 1u);
          }
        }
        commit<1>(empty0 + 8 * s);
      }
      __syncwarp();
      if (++s == P.stages) { s = 0; ph ^= 1; }
    }
    if (elect_one()) commit<1>(done);
    __syncwarp();
    mbar_wait(done, 0);
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  } else if (warp < 2 + P.pollers) {
    if (P.poll_mode == 0) {
      mbar_wait(done, 0);
    } else if (P.poll_mode == 1) {
      if (lane == 0) mbar_wait(done, 0);
      __syncwarp();
    } else {
      if (lane == 0) {
        uint32_t ok = 0;
        while (!ok) {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                       : "=r"(ok) : "r"(done), "r"(0u) : "memory");
          if (!ok) __nanosleep(200);
        }
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ---- bulk-copy stream alone: bytes per cycle into one SM's shared memory against the bytes kept in flight --------------
// `depth` slots of `slot_bytes`, each refilled as soon as it lands (nothing consumes the data).  src_mode 0: every CTA
// streams its own 512 KB window of an L2-resident buffer; 1: a 1 GB buffer (DRAM).
__global__ void __launch_bounds__(32, 1) stream_probe(const uint8_t* src, size_t window, int depth, int slot_bytes, int iters,
                                                       long long* out, int chunk) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[8];
  const uint32_t smem0 = (saddr(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(saddr(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint8_t* base = src + (size_t)blockIdx.x * window;
    uint32_t phs = 0;                                   // phase bits of the slots
    size_t off = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters + depth; ++it) {
      const int b = it % depth;
      if (it >= depth) { mbar_wait(saddr(&bars[b]), (phs >> b) & 1u); phs ^= 1u << b; }
      if (it < iters) {
        mbar_expect(saddr(&bars[b]), (uint32_t)slot_bytes);
        for (int o = 0; o < slot_bytes; o += chunk) {
          const int n = slot_bytes - o < chunk ? slot_bytes - o : chunk;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem0 + (uint32_t)(b * slot_bytes + o)),
                       "l"(base + ((off + o) & (window - 1))), "r"(n), "r"(saddr(&bars[b]))
                       : "memory");
        }
        off += slot_bytes;
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const int K = 1024;
  const int A = 16 * K;                       // one 128-row x 64-channel fp16 plane
  // single CTA, BN = 128: stage = [A_hi | A_lo | B_hi | B_lo] = 64 KB
  // pair,      BN = 128: stage = [A_hi | A_lo | B_hi half | B_lo half | dup] per CTA
  const Pattern pats[] = {
      {"cg1 N=256 alone", 1, 1, {{0, 2 * A, 256, 0}}, 4 * A, 2, 0},
      {"cg1 N=128 alone", 1, 1, {{0, 2 * A, 128, 0}}, 4 * A, 2, 0},
      {"cg1 N=64 alone", 1, 1, {{0, 2 * A, 64, 0}}, 4 * A, 2, 0},
      {"cg1 N=32 alone", 1, 1, {{0, 2 * A, 32, 0}}, 4 * A, 2, 0},
      {"cg1 NCAT BN=128 (N=256 hi x [hi;lo], N=128 lo x hi)", 1, 2, {{0, 2 * A, 256, 0}, {A, 2 * A, 128, 128}}, 4 * A, 3, 0},
      {"cg1 NCAT BN=128 + 64 KB/stage bulk copies", 1, 2, {{0, 2 * A, 256, 0}, {A, 2 * A, 128, 128}}, 4 * A, 1, 64},
      {"cg1 3 x N=128 (hh, hl, lh)", 1, 3, {{0, 2 * A, 128, 0}, {0, 3 * A, 128, 128}, {A, 2 * A, 128, 128}}, 4 * A, 3, 0},
      {"cg1 3 x N=256 (BN=256)", 1, 3, {{0, 2 * A, 256, 0}, {0, 4 * A, 256, 0}, {A, 2 * A, 256, 0}}, 6 * A, 2, 0},
      {"cg1 3 x N=256 + 96 KB/stage bulk copies", 1, 3, {{0, 2 * A, 256, 0}, {0, 4 * A, 256, 0}, {A, 2 * A, 256, 0}}, 6 * A, 1, 96},
      {"cg1 NCAT BN=64 (N=128, N=64)", 1, 2, {{0, 2 * A, 128, 0}, {A, 2 * A, 64, 64}}, 3 * A, 3, 0},
      {"cg2 M=256 N=256 alone", 2, 1, {{0, 2 * A, 256, 0}}, 3 * A, 2, 0},
      {"cg2 M=256 N=128 alone", 2, 1, {{0, 2 * A, 128, 0}}, 3 * A, 2, 0},
      {"cg2 M=256 N=64 alone", 2, 1, {{0, 2 * A, 64, 0}}, 3 * A, 2, 0},
      {"cg2 3 x N=128, pair tile 256 x 128 (hh | hl+lh apart)", 2, 3,
       {{0, 2 * A, 128, 0}, {0, 2 * A + A / 2, 128, 128}, {A, 2 * A, 128, 128}}, 3 * A, 3, 0},
      {"cg2 3 x N=128 + 48 KB/stage bulk copies", 2, 3,
       {{0, 2 * A, 128, 0}, {0, 2 * A + A / 2, 128, 128}, {A, 2 * A, 128, 128}}, 3 * A, 1, 48},
      {"cg2 NCAT pair 256 x 128 (N=256 + N=128 on a duplicate half)", 2, 2, {{0, 2 * A, 256, 0}, {A, 3 * A, 128, 64}}, 3 * A + A / 2, 3, 0},
      {"cg2 NCAT pair + 56 KB/stage bulk copies", 2, 2, {{0, 2 * A, 256, 0}, {A, 3 * A, 128, 64}}, 3 * A + A / 2, 1, 56},
      {"cg2 3 x N=256, pair tile 256 x 256", 2, 3, {{0, 2 * A, 256, 0}, {0, 3 * A, 256, 0}, {A, 2 * A, 256, 0}}, 4 * A, 3, 0},
      {"cg2 3 x N=256 + 64 KB/stage bulk copies", 2, 3, {{0, 2 * A, 256, 0}, {0, 3 * A, 256, 0}, {A, 2 * A, 256, 0}}, 4 * A, 1, 64},
      {"cg2 2 x N=256 (256 x 256: hh | hl+lh apart needs 512 cols)", 2, 3,
       {{0, 2 * A, 256, 0}, {0, 3 * A, 256, 256}, {A, 2 * A, 256, 256}}, 4 * A, 3, 0},
  };
  const int npat = sizeof(pats) / sizeof(pats[0]);
  const int iters = 400;
  long long *d_out, *d_tma;
  uint8_t* d_src;
  CK(cudaMalloc(&d_out, sizeof(long long) * 2 * sms));
  CK(cudaMalloc(&d_tma, sizeof(long long) * sms));
  CK(cudaMalloc(&d_src, 64 * 512 * 1024));
  CK(cudaMemset(d_src, 0x11, 64 * 512 * 1024));
  const int smem = 220 * 1024;
  long long* h = (long long*)malloc(sizeof(long long) * 2 * sms);
  long long* ht = (long long*)malloc(sizeof(long long) * sms);
  printf("%d SMs, %d stages of 4 K=16 slices per CTA; cycles are per slice\n", sms, iters);
  printf("%-62s %8s %8s %8s %8s %9s %8s\n", "pattern", "avg", "min", "max", "issue", "wall_us", "MHz");
  for (int pi = 0; pi < 2 * npat; ++pi) {
    const Pattern& pt = pats[pi % npat];
    const int style = pi / npat;
    if (pi == npat) printf("---- STYLE 1: convergent warp, elect.sync around the tcgen05 instructions ----\n");
    void (*kern)(const Params, long long*, long long*) = nullptr;
#define PICK(CGv, NMv) if (pt.cg == CGv && pt.nmma == NMv) kern = style ? probe<CGv, NMv, 1> : probe<CGv, NMv, 0>;
    PICK(1, 1) PICK(1, 2) PICK(1, 3) PICK(2, 1) PICK(2, 2) PICK(2, 3)
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    Params P;
    memset(&P, 0, sizeof(P));
    P.nmma = pt.nmma; P.stage_bytes = pt.stage_bytes; P.stages = pt.stages; P.iters = iters;
    P.tma_bytes = pt.tma_kb * 1024;
    P.scratch_off = ((pt.stage_bytes * pt.stages + 1023) / 1024) * 1024;
    if (P.tma_bytes > 0 && P.scratch_off + P.tma_bytes > smem - 1024) { printf("%s: no room for the scratch slot\n", pt.name); continue; }
    memcpy(P.m, pt.m, sizeof(P.m));
    P.src = d_src;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaMemset(d_tma, 0, sizeof(long long) * sms));
      CK(cudaEventRecord(e0));
      if (pt.cg == 1) {
        kern<<<sms, 128, smem>>>(P, d_out, d_tma);
      } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(sms / 2 * 2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, kern, P, d_out, d_tma));
      }
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    CK(cudaMemcpy(h, d_out, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ht, d_tma, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double sum = 0, isum = 0; long long mn = 1ll << 60, mx = 0; int n = 0;
    for (int b = 0; b < sms / pt.cg * pt.cg; b += pt.cg) { sum += h[2 * b]; isum += h[2 * b + 1]; if (h[2 * b] < mn) mn = h[2 * b]; if (h[2 * b] > mx) mx = h[2 * b]; ++n; }
    const double slices = iters * 4.0;
    const double avg = sum / n / slices;
    printf("%-62s %8.1f %8.1f %8.1f %8.1f %9.1f %8.0f", pt.name, avg, mn / slices, mx / slices, isum / n / slices, best * 1e3,
           (sum / n) / (best * 1e3));
    if (pt.tma_kb) printf("   bulk stream: %.1f cycles per stage = %.1f B/cycle", ht[0] / (double)iters, pt.tma_kb * 1024.0 * iters / ht[0]);
    printf("\n");
  }

  printf("---- ring skeleton (producer <-> issuer ping-pong, no loads): cycles per K stage (4 slices; MMA floor 768) ----\n");
  CK(cudaFuncSetAttribute(ring_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int rcfg[][5] = {  // stages, do_mma, pollers, poll_mode
      {3, 0, 0, 0}, {3, 1, 0, 0}, {3, 1, 8, 0}, {3, 1, 8, 1}, {3, 1, 8, 2}, {3, 1, 10, 0}, {2, 1, 0, 0}, {2, 1, 8, 0}, {1, 1, 0, 0}, {3, 0, 8, 0}, {3, 0, 8, 2},
  };
  for (unsigned i = 0; i < sizeof(rcfg) / sizeof(rcfg[0]); ++i) {
    RingParams R = {rcfg[i][0], 400, rcfg[i][1], rcfg[i][2], rcfg[i][3], 64 * 1024};
    ring_probe<<<sms, 64 + 32 * (R.pollers > 0 ? R.pollers : 0), smem>>>(R, d_out);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d_out, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double sum = 0;
    for (int b = 0; b < sms; ++b) sum += h[b];
    printf("stages %d  mma %d  pollers %2d  poll_mode %d : %8.1f cycles per stage\n", R.stages, R.do_mma, R.pollers, R.poll_mode, sum / sms / R.iters);
  }

  printf("---- bulk-copy stream alone: B/cycle/SM against bytes in flight (all SMs streaming) ----\n");
  CK(cudaFuncSetAttribute(stream_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  uint8_t* d_big = nullptr;
  const size_t big = (size_t)1 << 30;
  CK(cudaMalloc(&d_big, big));
  CK(cudaMemset(d_big, 1, big));
  for (int nsm = sms / 4; nsm <= sms; nsm *= 2) {     // fewer SMs streaming: is the limit per SM or L2-wide?
    stream_probe<<<nsm, 32, smem>>>(d_src, (size_t)256 * 1024, 3, 64 * 1024, 600, d_out, 16384);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d_out, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
    double sum = 0;
    for (int b = 0; b < nsm; ++b) sum += h[b];
    printf("L2    %3d CTAs streaming, 192 KB in flight each: %6.1f B/cycle/SM\n", nsm, 64.0 * 1024.0 * 600 / (sum / nsm));
  }
  for (int mode = 0; mode < 2; ++mode)
    for (int chunk = 16; chunk <= 16; chunk *= 2)
    for (int slot = 64; slot <= 64; slot *= 2)
      for (int depth = 1; depth <= 3 && depth * slot <= 208; ++depth) {
        const int its = 600;
        const size_t window = mode ? (size_t)4 << 20 : (size_t)256 * 1024;   // powers of two (masked in the kernel)
        stream_probe<<<sms, 32, smem>>>(mode ? d_big : d_src, window, depth, slot * 1024, its, d_out, chunk * 1024);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d_out, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
        double sum = 0;
        for (int b = 0; b < sms; ++b) sum += h[b];
        printf("%s  %2d KB copies, slot %2d KB x depth %d = %3d KB in flight: %6.1f B/cycle/SM\n", mode ? "DRAM" : "L2  ", chunk, slot, depth, slot * depth,
               (double)slot * 1024.0 * its / (sum / sms));
      }
  return 0;
}
