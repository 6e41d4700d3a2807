// Shared building blocks of the tcgen05 kernels (conv_tc.cu, stem_tc.cu): mbarrier / TMA / UMMA / TMEM
// wrappers, the 32-channel epilogue over the split-fp16 NHWC format, and the host-side tensor-map encoder.
#pragma once
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace accel {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// One lane of a converged warp (elect.sync): the tcgen05 / TMA issue instructions are predicated on it while the
// surrounding loop stays warp-uniform, so their operands live in uniform registers.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// Bounded wait: a protocol bug must surface as a launch failure, not as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  // not unrolled: ptxas otherwise replicates the try_wait 64 times at every call site (2 KB of code each; the conv kernel
  // grew to 15 K instructions and stalled on instruction fetch -- `no_instructions` 28 % of its samples)
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  // (no printf here by default: the call's ABI frame made ptxas spill registers that are live across every wait of the
  // conv kernels' hot loops -- 340 bytes of spill loads per thread; build with -DACCEL_MBAR_DIAG to get the message back)
#ifdef ACCEL_MBAR_DIAG
  printf("accel_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
#endif
  __trap();
}

// Polling wait (mbarrier.test_wait in a tight loop) for the single-purpose role warps: try_wait parks the warp and its
// wake-up costs a few hundred cycles per hand-off, which a short operand ring cannot hide.
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// L2 prefetch of one box of a tensor map (no shared memory, no barrier): a later tma_load of the same box is an L2 hit
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_5d(const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2),
               "r"(c3), "r"(c4)
               : "memory");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
      "[%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float v[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- TMA tensor stores (shared -> global), bulk-group completion -----------------------------------------
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(map),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One pixel row of a 32-channel fp16 chunk (64 bytes = four 16-byte pieces) into a SWIZZLE_64B staging tile:
// piece j of row r lives at chunk j ^ ((r >> 1) & 3) -- address bits [4,5] xor bits [7,8] -- which also spreads
// the 32 lanes of a warp over all banks.  `tile` is 512-byte aligned, rows are 64 bytes apart.
__device__ __forceinline__ void stage_row64(uint32_t tile, int r, const uint32_t w[16]) {
  const uint32_t row = tile + (uint32_t)r * 64u;
  const uint32_t sw = ((uint32_t)r >> 1) & 3u;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((uint32_t)j ^ sw) << 4)), "r"(w[4 * j]),
                 "r"(w[4 * j + 1]), "r"(w[4 * j + 2]), "r"(w[4 * j + 3])
                 : "memory");
}

// 32 activated fp32 values -> packed hi / lo fp16 pairs (16 words each)
__device__ __forceinline__ void split32_words(const float v[32], uint32_t hi[16], uint32_t lo[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) split_pair(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
}

// ---- 256-bit global accesses (one full 32-byte sector per thread) ----------------------------------
struct alignas(32) U8 {
  uint32_t v[8];
};

__device__ __forceinline__ U8 ld256(const void* p) {
  U8 r;
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}

__device__ __forceinline__ void st256(void* p, const U8& r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]),
               "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7])
               : "memory");
}

// residual of one 32-channel chunk of one pixel: 64 bytes of hi + 64 bytes of lo
struct ResChunk {
  U8 h[2];
  U8 l[2];
};

__device__ __forceinline__ void load_res(const Epilogue& e, int pix, int c0, ResChunk& rc) {
  const __half* ph = e.res_hi + (size_t)pix * e.res_ld + c0;
  const __half* pl = e.res_lo + (size_t)pix * e.res_ld + c0;
  rc.h[0] = ld256(ph);
  rc.h[1] = ld256(ph + 16);
  rc.l[0] = ld256(pl);
  rc.l[1] = ld256(pl + 16);
}

__device__ __forceinline__ float2 unpack_h2(uint32_t u) {
  __half2 h = *reinterpret_cast<__half2*>(&u);
  return __half22float2(h);
}

// split 32 fp32 values into hi/lo fp16 planes and store them as 2 + 2 sectors
__device__ __forceinline__ void store_split32(__half* hi, __half* lo, const float v[32]) {
  U8 a[2], b[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) split_pair(v[2 * i], v[2 * i + 1], a[i >> 3].v[i & 7], b[i >> 3].v[i & 7]);
  st256(hi, a[0]);
  st256(hi + 16, a[1]);
  st256(lo, b[0]);
  st256(lo + 16, b[1]);
}

// Epilogue of 32 consecutive channels [c0, c0+32) (all < Cout) of full-map pixel `pix`.
// sc32 / sh32: when non-null, the chunk's 32 scale / shift values staged in shared memory by the caller
// (they replace e.scale / e.shift).
// store_main = false: the caller writes the main split output itself (TMA store from shared memory); `v`
// holds the activated values on return either way.
__device__ __forceinline__ void epilogue_chunk32(const Epilogue& e, int pix, int c0, float (&v)[32], const ResChunk& rc,
                                                 const float* sc32 = nullptr, const float* sh32 = nullptr,
                                                 bool store_main = true) {
  if (e.raw_nchw) {                         // the linear part alone (key frame `fc6`: W*F for the commuted L head)
    size_t base, plane;
    if (nchw_base(e, pix, base, plane)) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 s = sc32 ? *(reinterpret_cast<const float4*>(sc32) + q)
                              : (e.scale ? __ldg(reinterpret_cast<const float4*>(e.scale + c0) + q) : make_float4(1.f, 1.f, 1.f, 1.f));
        float* rp = e.raw_nchw + base + (size_t)(c0 + 4 * q) * plane;
        rp[0] = v[4 * q + 0] * s.x; rp[plane] = v[4 * q + 1] * s.y; rp[2 * plane] = v[4 * q + 2] * s.z; rp[3 * plane] = v[4 * q + 3] * s.w;
      }
    }
  }
  if (sc32) {                               // (no run-time test inside the loops: each one costs a branch per four channels)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 s = *(reinterpret_cast<const float4*>(sc32) + q);
      const float4 b = *(reinterpret_cast<const float4*>(sh32) + q);
      v[4 * q + 0] = fmaf(v[4 * q + 0], s.x, b.x);
      v[4 * q + 1] = fmaf(v[4 * q + 1], s.y, b.y);
      v[4 * q + 2] = fmaf(v[4 * q + 2], s.z, b.z);
      v[4 * q + 3] = fmaf(v[4 * q + 3], s.w, b.w);
    }
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 s = e.scale ? __ldg(reinterpret_cast<const float4*>(e.scale + c0) + q) : make_float4(1.f, 1.f, 1.f, 1.f);
      const float4 b = e.shift ? __ldg(reinterpret_cast<const float4*>(e.shift + c0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      v[4 * q + 0] = fmaf(v[4 * q + 0], s.x, b.x);
      v[4 * q + 1] = fmaf(v[4 * q + 1], s.y, b.y);
      v[4 * q + 2] = fmaf(v[4 * q + 2], s.z, b.z);
      v[4 * q + 3] = fmaf(v[4 * q + 3], s.w, b.w);
    }
  }
  if (e.res_hi) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 h = unpack_h2(rc.h[i >> 3].v[i & 7]);
      const float2 l = unpack_h2(rc.l[i >> 3].v[i & 7]);
      v[2 * i] += h.x + l.x;
      v[2 * i + 1] += h.y + l.y;
    }
  }
  apply_act_n(v, e.act);
  if (e.out_hi && store_main) store_split32(e.out_hi + (size_t)pix * e.out_ld + c0, e.out_lo + (size_t)pix * e.out_ld + c0, v);
  if (e.out_nchw) {
    size_t base, plane;
    if (nchw_base(e, pix, base, plane)) {
#pragma unroll
      for (int i = 0; i < 32; ++i) e.out_nchw[base + (size_t)(c0 + i) * plane] = v[i];
    }
  }
  if (e.out2_hi) {
    float v2[32];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 s = __ldg(reinterpret_cast<const float4*>(e.scale2 + c0) + q);
      const float4 b = __ldg(reinterpret_cast<const float4*>(e.shift2 + c0) + q);
      v2[4 * q + 0] = apply_act(fmaf(v[4 * q + 0], s.x, b.x), e.act2);
      v2[4 * q + 1] = apply_act(fmaf(v[4 * q + 1], s.y, b.y), e.act2);
      v2[4 * q + 2] = apply_act(fmaf(v[4 * q + 2], s.z, b.z), e.act2);
      v2[4 * q + 3] = apply_act(fmaf(v[4 * q + 3], s.w, b.w), e.act2);
    }
    store_split32(e.out2_hi + (size_t)pix * e.out2_ld + c0, e.out2_lo + (size_t)pix * e.out2_ld + c0, v2);
  }
}

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

inline bool encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
            const cuuint32_t* box, char* err, int errlen, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B,
            CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT16) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled is unavailable");
    return false;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, dtype, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu] box [%u %u %u]", (int)r, rank,
             (unsigned long long)dims[0], (unsigned long long)dims[1], rank > 2 ? (unsigned long long)dims[2] : 0ull, box[0],
             box[1], rank > 2 ? box[2] : 0u);
    return false;
  }
  return true;
}

inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}


}  // namespace tc
}  // namespace accel
