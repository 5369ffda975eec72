// Microbenchmark: back-to-back tcgen05.mma kind::tf32 dispatch cost on B200 (one CTA per SM, one issuing thread).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_umma tools/ubench_umma.cu
//   ./tools/ubench_umma
// Prints cycles per MMA for M=128, K=8 and several N, with A taken from TMEM (.ts form) or from shared memory,
// accumulating into one accumulator or alternating between two.  Used to size the shapelet cross-term tiles
// (DESIGN.md §3): the dispatch floor is 128*N/256 cycles; anything above it is per-instruction overhead.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}
// mode: 0 = A from TMEM, 1 = A from smem ; nacc accumulators used round-robin ; kind 0 tf32, 1 bf16
template <int mode, int kind, int nacc>
__global__ void __launch_bounds__(128, 1) ubench(int N, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<float*>(base)[i] = 0.001f * (i & 255);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (threadIdx.x < 32 && elect_one()) {
    const uint32_t idesc = kind ? idesc_bf16(128, N) : idesc_tf32(128, N);
    const uint64_t bdesc = desc_sw128(smem_u32(base));
    const uint64_t adesc = desc_sw128(smem_u32(base + 32 * 1024));
    const uint32_t a_tmem = tmem + 480;     // 8 columns of A (garbage values are fine for timing)
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tmem + (uint32_t)((i % nacc) * (N <= 128 ? 128 : 256));
      if (mode == 0) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bdesc),
                     "r"(idesc), "r"(i >= nacc ? 1 : 0) : "memory");
      } else if (kind == 0) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc),
                     "r"(idesc), "r"(i >= nacc ? 1 : 0) : "memory");
      } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc),
                     "r"(idesc), "r"(i >= nacc ? 1 : 0) : "memory");
      }
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  const int iters = 2048;
  printf("%5s %6s %5s %5s | %10s %10s | floor\n", "kind", "A", "N", "nacc", "issue/mma", "total/mma");
  for (int kind = 0; kind < 2; ++kind)
    for (int mode = 0; mode < 2; ++mode) {
      if (kind == 1 && mode == 0) continue;
      for (int N : {16, 32, 64, 80, 96, 128, 160, 240, 256})
        for (int nacc : {1, 2}) {
          if (nacc * (N <= 128 ? 128 : 256) > 448) continue;
          auto launch = [&](auto kern) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024); kern<<<148, 128, 66 * 1024>>>(N, iters, out); };
          if (kind == 0 && mode == 0 && nacc == 1) launch(ubench<0, 0, 1>);
          if (kind == 0 && mode == 0 && nacc == 2) launch(ubench<0, 0, 2>);
          if (kind == 0 && mode == 1 && nacc == 1) launch(ubench<1, 0, 1>);
          if (kind == 0 && mode == 1 && nacc == 2) launch(ubench<1, 0, 2>);
          if (kind == 1 && mode == 1 && nacc == 1) launch(ubench<1, 1, 1>);
          if (kind == 1 && mode == 1 && nacc == 2) launch(ubench<1, 1, 2>);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
          long long h[2];
          cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
          printf("%5s %6s %5d %5d | %10.1f %10.1f | %d\n", kind ? "bf16" : "tf32", mode ? "smem" : "tmem", N, nacc,
                 (double)h[0] / iters, (double)h[1] / iters, 128 * N / 256);
        }
    }
  return 0;
}
