// Microbenchmark: cost of the MMA issuer's per-stage sequence of the shapelet tcgen05 kernel, in isolation:
//   [mbarrier try_wait on a completed phase] + tcgen05.fence + 12 x tcgen05.mma (3xTF32, N=80, A in TMEM) + 2 commits
// against the 12 x 40 = 480-cycle execution floor.  Variants drop one ingredient at a time.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_umma_stage tools/ubench_umma_stage.cu
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(200000u) : "memory");
}

// flags: 1 = wait on a completed barrier each stage, 2 = tcgen05.fence each stage, 4 = two commits each stage
__global__ void __launch_bounds__(128, 1) ubench(int stages, int flags, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_done, bar_ready, bar_sink[8];
  __shared__ uint32_t tslot;
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < (80 * 1024) / 4; i += blockDim.x) reinterpret_cast<float*>(base)[i] = 0.001f * (i & 255);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_done)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_ready)));
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_sink[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_ready)) : "memory");   // phase 0 complete for ever
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
    const uint32_t idesc = idesc_tf32(128, 80);
    long long t0 = clock64();
    for (int it = 0; it < stages; ++it) {
      const int s = it & 3;
      if (flags & 1) wait(&bar_ready, 0);
      if (flags & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sb_hi = smem_u32(base + s * 20480), sb_lo = sb_hi + 10240;
      const uint32_t a_hi = tmem + 256 + s * 64, a_lo = a_hi + 32;
#pragma unroll
      for (int k8 = 0; k8 < 4; ++k8) {
        mma(tmem, a_lo + k8 * 8, desc_sw128(sb_hi + k8 * 32), idesc, (it | k8) != 0);
        mma(tmem, a_hi + k8 * 8, desc_sw128(sb_lo + k8 * 32), idesc, 1);
        mma(tmem, a_hi + k8 * 8, desc_sw128(sb_hi + k8 * 32), idesc, 1);
      }
      if (flags & 4) { commit(&bar_sink[s]); commit(&bar_sink[4 + s]); }
    }
    long long t1 = clock64();
    commit(&bar_done);
    wait(&bar_done, 0);
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
  cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, 84 * 1024);
  const int stages = 1024;
  printf("per-stage cost of the MMA issuer (12 MMAs 128x80x8 tf32, floor 480 cycles)\n%-34s | %10s %10s\n", "variant", "issue/stage", "total/stage");
  const char* names[] = {"12 MMAs only", "+ wait(done barrier)", "+ tcgen05.fence", "+ wait + fence", "+ 2 commits", "+ wait + commits",
                         "+ fence + commits", "+ wait + fence + 2 commits (kernel)"};
  for (int flags = 0; flags < 8; ++flags) {
    ubench<<<148, 128, 84 * 1024>>>(stages, flags, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[2];
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%-34s | %10.1f %10.1f\n", names[flags], (double)h[0] / stages, (double)h[1] / stages);
  }
  return 0;
}
