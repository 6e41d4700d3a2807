// 7x7 / stride 2 / pad 3 stems on the tensor cores (R101/R50 `conv1`, R18/34 `conv0`, FlowNet `flow_conv1`
// with its 2x2 average pool and /255 folded into the patch load).  Cin is 3 (one frame) or 6 (frame pair).
//
// Too few input channels for the NHWC/TMA implicit GEMM of conv_tc.cu, so the A operand is built in shared
// memory by software producers: a CTA tile is 4 x 32 output pixels (UMMA M = 128); the fp32 input patch
// (Cin x 13 x 72) is staged once, then each producer thread owns one output pixel and writes, per (channel,
// filter row), the 8 consecutive input columns it needs as one 16-byte chunk of fp16 hi / lo -- i.e.
// K index = (c*7 + ky)*8 + (kx+1), with a zero weight in slot 0 -- straight into the SWIZZLE_128B K-major
// layout tcgen05.mma reads.  Weights (64 x K, hi/lo) are TMA-loaded once per CTA.  Warp roles: 2 producer
// groups of 4 warps that alternate tiles (one group's global patch load overlaps the other's build), one MMA
// issuer (fp16x3 into a double-buffered TMEM accumulator), 4 epilogue warps.
#include <string.h>

#include "kernels.h"
#include "tc_common.cuh"

namespace accel {

namespace {

using namespace tc;

constexpr int ST_BW = 32, ST_BH = 4;
constexpr int ST_PH = (ST_BH - 1) * 2 + 7;      // 13 input rows
constexpr int ST_PW = 72;                       // 70 input columns used (one leading pad column), 16-byte rows
constexpr int kProdWarps = 8;
constexpr int kMmaWarp = 8;
constexpr int kFirstEpiWarp = 10;               // warps 10..13 -> TMEM lane quarters 2,3,0,1
constexpr int kThreads = 32 * 14;
constexpr int kMaxStages = 4;                  // total; each producer group owns half of them (its own ring)
constexpr int kSmemMaxDynamic = 227 * 1024 - 1024;

struct alignas(64) StemTcParams {
  CUtensorMap b_hi, b_lo;
  Epilogue epi;
  const float* src0;
  const float* src1;
  int Hs, Ws;            // source frame size
  int Hp, Wp;            // size the convolution sees (pooled for FlowNet)
  int pool, Cin;
  float in_scale[6], in_shift[6];
  int Ho, Wo, tiles_x, tiles_y;
  int nblocks;           // K blocks of 64 (8 chunks)
  int nchunks;           // real chunks = Cin * 7
  int stages;            // per producer group (ring)
  int debug;             // ACCEL_STEM_DEBUG (timing decomposition only): 1 no epilogue stores, 2 no patch loads, 4 no A build
};

__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) stem_tc_kernel(const __grid_constant__ StemTcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 5];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = (uint32_t)P.nblocks * 16384u;            // per block: hi 8 KB + lo 8 KB
  const uint32_t stage0 = smem0 + b_bytes;                          // per stage: A hi 16 KB + A lo 16 KB
  const uint32_t patch0 = stage0 + 2u * (uint32_t)P.stages * 32768u;
  const uint32_t patch_floats = (uint32_t)P.Cin * ST_PH * ST_PW;
  uint8_t* gen0 = smem_raw + (smem0 - smem_u32(smem_raw));          // generic pointer to smem0

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kMaxStages]);
  const uint32_t tfull0 = smem_u32(&bars[2 * kMaxStages]), tempty0 = smem_u32(&bars[2 * kMaxStages + 2]);
  const uint32_t bfull = smem_u32(&bars[2 * kMaxStages + 4]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2 * P.stages; ++s) {  // ring g uses barriers [g*stages, (g+1)*stages)
      mbar_init(full0 + 8 * s, 4);            // the 4 warps of the producer group that owns the ring
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 4);
    }
    mbar_init(bfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_trigger();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_wait();                                                       // everything above touched no global memory
  const uint32_t tmem_base = tmem_base_slot;
  const int ntiles = P.tiles_x * P.tiles_y;

  if (warp < kProdWarps) {
    // ===================================== A-operand producers =====================================
    const int grp = warp >> 2;
    const int r = (warp & 3) * 32 + lane;              // tile row = output pixel
    const int ty = r >> 5, tx = r & 31;
    float* patch = reinterpret_cast<float*>(gen0 + (patch0 - smem0)) + (size_t)grp * patch_floats;
    const int gtid = threadIdx.x & 127;
    const int nvec = P.Cin * ST_PH * (ST_PW / 4);
    for (int k = grp; blockIdx.x + k * (int)gridDim.x < ntiles; k += 2) {
      const int tile = blockIdx.x + k * gridDim.x;
      const int oy0 = (tile / P.tiles_x) * ST_BH, ox0 = (tile % P.tiles_x) * ST_BW;
      const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 4;  // patch column 0 = input column ix0 (a multiple of 4)
      // ---- stage the fp32 patch (16-byte vectors: each one is entirely inside or outside the image) ----
      for (int i = gtid; i < nvec; i += 128) {
        const int c = i / (ST_PH * (ST_PW / 4));
        const int rem = i - c * (ST_PH * (ST_PW / 4));
        const int py = rem / (ST_PW / 4), pv = rem - py * (ST_PW / 4);
        const int iy = iy0 + py, ix = ix0 + pv * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!(P.debug & 2) && iy >= 0 && iy < P.Hp && ix >= 0 && ix < P.Wp) {
          const float* src = (c < 3 ? P.src0 : P.src1) + (size_t)(c < 3 ? c : c - 3) * P.Hs * P.Ws;
          if (P.pool) {
            const float4* q0 = reinterpret_cast<const float4*>(src + (size_t)(2 * iy) * P.Ws + 2 * ix);
            const float4* q1 = reinterpret_cast<const float4*>(src + (size_t)(2 * iy + 1) * P.Ws + 2 * ix);
            const float4 a0 = __ldg(q0), a1 = __ldg(q0 + 1), b0 = __ldg(q1), b1 = __ldg(q1 + 1);
            v.x = ((a0.x + a0.y) + (b0.x + b0.y)) * 0.25f;
            v.y = ((a0.z + a0.w) + (b0.z + b0.w)) * 0.25f;
            v.z = ((a1.x + a1.y) + (b1.x + b1.y)) * 0.25f;
            v.w = ((a1.z + a1.w) + (b1.z + b1.w)) * 0.25f;
          } else {
            v = __ldg(reinterpret_cast<const float4*>(src + (size_t)iy * P.Ws + ix));
          }
          const float sc = P.in_scale[c], sh = P.in_shift[c];
          v.x = fmaf(v.x, sc, sh); v.y = fmaf(v.y, sc, sh); v.z = fmaf(v.z, sc, sh); v.w = fmaf(v.w, sc, sh);
        }
        reinterpret_cast<float4*>(patch)[i] = v;
      }
      named_sync(1 + grp, 128);
      // ---- build the K blocks of this tile ----
      int c = 0, ky = 0;
      for (int b = 0; b < P.nblocks; ++b) {
        // Each group is the only producer of its own ring, so the parity protocol never aliases.
        const int q = (k >> 1) * P.nblocks + b;
        const int s = grp * P.stages + q % P.stages;
        const uint32_t ph = (uint32_t)(q / P.stages) & 1u;
        mbar_wait(empty0 + 8 * s, ph ^ 1u);
        uint8_t* a_hi = gen0 + (stage0 - smem0) + (size_t)s * 32768 + (size_t)r * 128;
        uint8_t* a_lo = a_hi + 16384;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 hv = make_uint4(0u, 0u, 0u, 0u), lv = hv;
          if (b * 8 + j < P.nchunks && !(P.debug & 4)) {
            const float2* src = reinterpret_cast<const float2*>(patch + (c * ST_PH + ty * 2 + ky) * ST_PW + tx * 2);
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = src[e];
              split_pair(f.x, f.y, hw[e], lw[e]);
            }
            hv = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            lv = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            if (++ky == 7) { ky = 0; ++c; }
          }
          const int off = ((j ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(a_hi + off) = hv;
          *reinterpret_cast<uint4*>(a_lo + off) = lv;
        }
        if (!(P.debug & 8)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
        __syncwarp();
        if (lane == 0) mbar_arrive(full0 + 8 * s);
      }
      named_sync(1 + grp, 128);                      // the patch buffer is rewritten next
    }
  } else if (warp == kMmaWarp) {
    // ===================================== MMA issuer =======================================
    // converged warp, elect.sync around the tcgen05 / TMA instructions: operands in uniform registers (under
    // `if (lane == 0)` every UTCHMMA sat in an ELECT / R2UR waterfall, ~180 cycles per MMA against the N = 64 MMA's 62 --
    // this loop, not the producers, was the 77 us hand-off floor of profiles/r01_stem_decomposition.txt)
    {
      if (elect_one()) {
        mbar_arrive_expect_tx(bfull, b_bytes);
        for (int b = 0; b < P.nblocks; ++b) {
          tma_load_2d(smem0 + b * 16384, &P.b_hi, bfull, b * 64, 0);
          tma_load_2d(smem0 + b * 16384 + 8192, &P.b_lo, bfull, b * 64, 0);
        }
      }
      __syncwarp();
      mbar_wait(bfull, 0);
      const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const int kreal = P.nchunks * 8;
      int acc = 0;
      uint32_t accph = 0;
      for (int k = 0; blockIdx.x + k * (int)gridDim.x < ntiles; ++k) {
        mbar_wait(tempty0 + 8 * acc, accph ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem_base + (uint32_t)(acc * 64);
        for (int b = 0; b < P.nblocks; ++b) {
          const int q = (k >> 1) * P.nblocks + b;
          const int s = (k & 1) * P.stages + q % P.stages;
          mbar_wait(full0 + 8 * s, (uint32_t)(q / P.stages) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = stage0 + s * 32768;
          const uint64_t ah = umma_desc(sa), al = umma_desc(sa + 16384);
          const uint64_t bh = umma_desc(smem0 + b * 16384), bl = umma_desc(smem0 + b * 16384 + 8192);
          const int slices = min(4, (kreal - b * 64 + 15) / 16);
          const uint32_t first0 = b > 0 ? 1u : 0u;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (ks < slices) {
                const uint64_t adv = (uint64_t)(ks * 2);
                umma_f16(d, ah + adv, bh + adv, idesc, ks > 0 ? 1u : first0);
                umma_f16(d, ah + adv, bl + adv, idesc, 1u);
                umma_f16(d, al + adv, bh + adv, idesc, 1u);
              }
            }
            umma_commit(empty0 + 8 * s);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(tfull0 + 8 * acc);
        __syncwarp();
        if (++acc == 2) { acc = 0; accph ^= 1; }
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ===================================== epilogue ==========================================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int ty = r >> 5, tx = r & 31;
    const bool vec = (P.epi.out_ld % 16 == 0) && ((reinterpret_cast<uintptr_t>(P.epi.out_hi) & 31) == 0) &&
                     ((reinterpret_cast<uintptr_t>(P.epi.out_lo) & 31) == 0) && P.epi.out2_hi == nullptr &&
                     P.epi.res_hi == nullptr;
    int acc = 0;
    uint32_t accph = 0;
    for (int k = 0; blockIdx.x + k * (int)gridDim.x < ntiles; ++k) {
      const int tile = blockIdx.x + k * gridDim.x;
      const int y = (tile / P.tiles_x) * ST_BH + ty, x = (tile % P.tiles_x) * ST_BW + tx;
      const bool valid = y < P.Ho && x < P.Wo;
      const int pix = y * P.Wo + x;
      mbar_wait(tfull0 + 8 * acc, accph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 64);
#pragma unroll
      for (int cc = 0; cc < 64; cc += 32) {
        float v[32];
        tmem_ld32(taddr + cc, v);
        if (valid && !(P.debug & 1)) {
          if (vec) {
            ResChunk none{};
            epilogue_chunk32(P.epi, pix, cc, v, none);
          } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) epilogue_store<8>(P.epi, pix, cc + g * 8, v + g * 8);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      if (++acc == 2) { acc = 0; accph ^= 1; }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

}  // namespace

struct StemTcPlan {
  StemTcParams p;
  int grid;
  size_t smem;
};

int stem_tc_kpad(int cin) { return ((cin * 7 * 8 + 63) / 64) * 64; }

// Packs fp32 (64, Cin, 7, 7) stem weights into K-major rows of stem_tc_kpad(Cin) fp16 (hi and lo planes):
// K index = (c*7 + ky)*8 + kx + 1, slot 0 of every 8 is the zero pad; rows are pre-scaled by a power of two.
void stem_tc_pack_weights(const float* w, int cin, std::vector<__half>& hi, std::vector<__half>& lo,
                          std::vector<float>& prescale) {
  const int kpad = stem_tc_kpad(cin);
  hi.assign((size_t)64 * kpad, __float2half_rn(0.f));
  lo.assign((size_t)64 * kpad, __float2half_rn(0.f));
  prescale.assign(64, 1.f);
  for (int n = 0; n < 64; ++n) {
    float m = 0.f;
    for (int i = 0; i < cin * 49; ++i) m = fmaxf(m, fabsf(w[(size_t)n * cin * 49 + i]));
    float mult = 1.f;
    if (m > 0.f && isfinite(m)) {
      const int ex = ilogbf(m);
      mult = ldexpf(1.f, -ex);
      prescale[n] = ldexpf(1.f, ex);
    }
    for (int c = 0; c < cin; ++c)
      for (int ky = 0; ky < 7; ++ky)
        for (int kx = 0; kx < 7; ++kx) {
          const float v = w[(((size_t)n * cin + c) * 7 + ky) * 7 + kx] * mult;
          const size_t k = (size_t)n * kpad + (c * 7 + ky) * 8 + kx + 1;
          const __half h = __float2half_rn(v);
          hi[k] = h;
          lo[k] = __float2half_rn(v - __half2float(h));
        }
  }
}

bool stem_tc_supported(const StemParams& S) {
  const int Wp = S.pool ? S.Ws / 2 : S.Ws;
  return (S.Cin == 3 || S.Cin == 6) && Wp % 4 == 0 && S.Ws % 4 == 0 && (!S.pool || S.Ws % 8 == 0);
}

StemTcPlan* stem_tc_plan_create(const StemParams& S, const __half* w_hi, const __half* w_lo, int num_sms, char* err,
                                int errlen) {
  if (!stem_tc_supported(S)) {
    snprintf(err, errlen, "stem shape not supported by the tcgen05 stem");
    return nullptr;
  }
  StemTcPlan* plan = new StemTcPlan();
  StemTcParams& P = plan->p;
  memset(&P, 0, sizeof(P));
  P.epi = S.epi;
  P.Hs = S.Hs; P.Ws = S.Ws;
  P.pool = S.pool; P.Cin = S.Cin;
  P.Hp = S.pool ? S.Hs / 2 : S.Hs;
  P.Wp = S.pool ? S.Ws / 2 : S.Ws;
  memcpy(P.in_scale, S.in_scale, sizeof(P.in_scale));
  memcpy(P.in_shift, S.in_shift, sizeof(P.in_shift));
  P.Ho = S.Ho; P.Wo = S.Wo;
  P.tiles_x = (S.Wo + ST_BW - 1) / ST_BW;
  P.tiles_y = (S.Ho + ST_BH - 1) / ST_BH;
  const int kpad = stem_tc_kpad(S.Cin);
  P.nblocks = kpad / 64;
  P.nchunks = S.Cin * 7;
  const size_t fixed = (size_t)P.nblocks * 16384 + 2 * (size_t)S.Cin * ST_PH * ST_PW * sizeof(float) + 1024;
  int stages = (int)((kSmemMaxDynamic - fixed) / 32768);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) {
    snprintf(err, errlen, "stem does not fit in shared memory");
    delete plan;
    return nullptr;
  }
  P.stages = stages / 2;
  P.debug = env_int("ACCEL_STEM_DEBUG", 0);
  plan->smem = fixed + 2 * (size_t)P.stages * 32768;
  const int ntiles = P.tiles_x * P.tiles_y;
  plan->grid = ntiles < num_sms ? ntiles : num_sms;
  cuuint64_t dims[2] = {(cuuint64_t)kpad, 64};
  cuuint64_t str[1] = {(cuuint64_t)kpad * sizeof(__half)};
  cuuint32_t box[2] = {64, 64};
  if (!encode(&P.b_hi, w_hi, 2, dims, str, box, err, errlen) || !encode(&P.b_lo, w_lo, 2, dims, str, box, err, errlen)) {
    delete plan;
    return nullptr;
  }
  if (first_time_on_device(ONCE_STEM_TC)) {
    cudaError_t ce = cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMaxDynamic);
    if (ce != cudaSuccess) {
      snprintf(err, errlen, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce));
      delete plan;
      return nullptr;
    }
  }
  return plan;
}

void stem_tc_plan_destroy(StemTcPlan* plan) { delete plan; }

cudaError_t launch_stem_tc(const StemTcPlan* plan, const float* src0, const float* src1, cudaStream_t stream) {
  StemTcParams P = plan->p;
  P.src0 = src0;
  P.src1 = src1;
  return launch_k(stem_tc_kernel, dim3(plan->grid), dim3(kThreads), plan->smem, stream, P);
  return cudaGetLastError();
}

}  // namespace accel
