// Minimal producer/consumer ring with the same synchronisation as warp_kernel_fused (warp_staged.cu): one producer lane
// issues cp.async.bulk into a shared-memory slot and signals a "full" mbarrier through complete_tx; consumer warps wait on
// it, read the slot, and release it through an "empty" mbarrier the producer waits on before overwriting the slot.
// Run under `compute-sanitizer --tool racecheck` to see what the tool makes of a ring that is correct by construction
// (results are checked below): profiles/r02_sanitizer_racecheck.txt quotes the outcome next to the product kernel's.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/_build/racecheck_probe tools/racecheck_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int STAGES = 2, SLOT = 256, CONS = 64, ITERS = 16;

__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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

__global__ void ring_probe(const float* __restrict__ src, float* __restrict__ dst) {
  __shared__ __align__(128) float ring[STAGES][SLOT];
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(saddr(&full[s]), 1); mbar_init(saddr(&empty[s]), CONS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid >= CONS) {                                            // producer warp
    if (tid == CONS) {
      int s = 0; uint32_t ph = 0;
      for (int i = 0; i < ITERS; ++i) {
        mbar_wait(saddr(&empty[s]), ph ^ 1);
        mbar_expect(saddr(&full[s]), SLOT * 4);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         saddr(&ring[s][0])), "l"(src + (size_t)i * SLOT), "r"(SLOT * 4), "r"(saddr(&full[s])) : "memory");
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
    return;
  }
  int s = 0; uint32_t ph = 0;
  for (int i = 0; i < ITERS; ++i) {
    mbar_wait(saddr(&full[s]), ph);
    float acc = 0.f;
    for (int k = 0; k < SLOT / CONS; ++k) acc += ring[s][(tid + k * CONS + 17) % SLOT];
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(saddr(&empty[s]));
    dst[(size_t)i * CONS + tid] = acc;
    if (++s == STAGES) { s = 0; ph ^= 1; }
  }
}

int main() {
  const int n = ITERS * SLOT;
  float *h = new float[n], *src, *dst, out[ITERS * CONS];
  for (int i = 0; i < n; ++i) h[i] = (float)((i * 2654435761u) >> 20);
  cudaMalloc(&src, n * 4); cudaMalloc(&dst, sizeof(out));
  cudaMemcpy(src, h, n * 4, cudaMemcpyHostToDevice);
  ring_probe<<<4, CONS + 32>>>(src, dst);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 2; }
  cudaMemcpy(out, dst, sizeof(out), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < ITERS; ++i)
    for (int t = 0; t < CONS; ++t) {
      float acc = 0.f;
      for (int k = 0; k < SLOT / CONS; ++k) acc += h[i * SLOT + (t + k * CONS + 17) % SLOT];
      bad += acc != out[i * CONS + t];
    }
  printf("ring_probe: %d mismatches of %d\n", bad, ITERS * CONS);
  return bad != 0;
}
