// Microbenchmark: tcgen05.mma (kind::tf32, M=128, N=80, K=8, A in TMEM) dispatch rate while other warps of the CTA
// hammer tensor memory with tcgen05.st (the A-stage producers) and / or tcgen05.ld (the epilogue).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_umma_contention tools/ubench_umma_contention.cu
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
__device__ volatile int g_sink;

// warps: 0 = MMA issuer; 1..nst = tcgen05.st loops; next nld = tcgen05.ld loops.  gap = ALU instructions between ops.
__global__ void __launch_bounds__(416, 1) ubench(int N, int iters, int nst, int nld, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  __shared__ volatile int stop;
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < (32 * 1024) / 4; i += blockDim.x) reinterpret_cast<float*>(base)[i] = 0.001f * (i & 255);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = idesc_tf32(128, N);
      const uint64_t bdesc = desc_sw128(smem_u32(base));
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t a_tmem = tmem + 256 + (i & 3) * 64 + ((i >> 2) & 3) * 8;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(a_tmem), "l"(bdesc),
                     "r"(idesc), "r"(i ? 1 : 0) : "memory");
      }
      long long t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
      long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
      stop = 1;
    }
    __syncwarp();
  } else if (warp <= nst) {
    // store loop: 64 columns (hi + lo) of this warp's lane quarter per iteration, like one A stage
    const uint32_t lane_base = tmem + ((uint32_t)(((warp - 1) & 3) * 32) << 16) + 256 + 64 * (((warp - 1) >> 2) & 1) + 128;
    uint32_t v = threadIdx.x;
    long long cnt = 0;
    while (!stop) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
                   ::"r"(lane_base), "r"(v) : "memory");
      asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
                   ::"r"(lane_base + 32), "r"(v) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      ++cnt; ++v;
    }
    if (blockIdx.x == 0 && threadIdx.x == 32) out[2] = cnt;
  } else if (warp <= nst + nld) {
    const uint32_t lane_base = tmem + ((uint32_t)(((warp - 1 - nst) & 3) * 32) << 16) + 128;   // columns 128.. (not the live accumulator)
    uint32_t acc = 0;
    long long cnt = 0;
    while (!stop) {
      uint32_t r[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                     "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(lane_base + (uint32_t)((cnt & 3) * 16)) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) acc += r[j];
      ++cnt;
    }
    if (acc == 0x12345678u) g_sink = 1;
    if (blockIdx.x == 0 && threadIdx.x == 32 * (nst + 1)) out[3] = cnt;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 64);
  cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  const int iters = 4096, N = 80;
  printf("tcgen05.mma tf32 128x%dx8, A in TMEM: cycles per MMA (floor %d) under concurrent TMEM traffic\n", N, 128 * N / 256);
  printf("%8s %8s | %10s | %14s %14s\n", "st warps", "ld warps", "cyc/mma", "st iters/kcyc", "ld iters/kcyc");
  for (int nst : {0, 4, 8})
    for (int nld : {0, 4, 8}) {
      if (1 + nst + nld > 13) continue;
      cudaMemset(out, 0, 64);
      ubench<<<148, 32 * (1 + nst + nld), 40 * 1024>>>(N, iters, nst, nld, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[4];
      cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
      printf("%8d %8d | %10.1f | %14.2f %14.2f\n", nst, nld, (double)h[1] / iters, 1000.0 * h[2] / h[1], 1000.0 * h[3] / h[1]);
    }
  return 0;
}
