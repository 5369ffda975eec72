// EXPERIMENT (negative result, kept for the record — see DESIGN.md §3.2):
// RESULT ON B200: with either operand flagged MN-major in the instruction descriptor and a SWIZZLE_NONE shared-memory
// descriptor, tcgen05.mma kind::tf32 contributes exactly ZERO to the accumulator (modes 1, 2, 3; also with the
// canonical non-overlapping layout, modes 11 / 27), while the K-major form of the same harness computes (mode 0).
// A zero-copy Hankel operand for the BACKWARD contraction (K = window index) is therefore not available this way.
//
// Known-answer test: tcgen05.mma kind::tf32 with BOTH operands MN-major, no swizzle, read straight from raw rows:
//   A[u][r] = x[4 r + u]            (Hankel matrix of a series, row stride 4 samples = one 16-byte chunk)
//   B[(s,j)][r] = c[s][4 r + j]     (4 shifted views of each coefficient row)
//   D[u][(s,j)] = sum_r A[u][r] B[(s,j)][r] = sum_r x[4r+u] c[s][4r+j]
// Canonical MN-major INTERLEAVE layout (cute/atom/mma_traits_sm100.hpp): 16-byte chunk (n, k) lives at
// base + n*SBO + (k%8)*16 B + (k/8)*LBO.  The Hankel operand is that layout with SBO = 16 B (overlapping core
// matrices), the coefficient operand with SBO = row pitch.  Integer-valued inputs make the tf32 product exact.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/test_hankel_umma tools/test_hankel_umma.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int kXLen = 2048, kRows = 8, kPitch = 1024, kSteps = 6;   // 6 k-steps of 8 rows r

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_mn_none(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// D=f32, A=B=tf32, both MN-major (bits 15, 16), N>>3 at 17, M>>4 at 24
__device__ __forceinline__ uint32_t idesc_tf32_mn(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint64_t desc_k_none(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__global__ void __launch_bounds__(128, 1) hankel_kernel(const float* __restrict__ xg, const float* __restrict__ cg, float* __restrict__ out, int mode) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  float* xs = smem;                 // [kXLen]
  float* cs = smem + kXLen;         // [kRows][kPitch]
  float* acan = cs + kRows * kPitch;                 // [kSteps][32 chunks n][8 k][4]  canonical, SBO = 128 B
  float* bcan = acan + kSteps * 32 * 32;             // [kSteps][8 chunks n][8 k][4]
  for (int i = threadIdx.x; i < kXLen; i += blockDim.x) xs[i] = xg[i];
  for (int i = threadIdx.x; i < kRows * kPitch; i += blockDim.x) cs[i] = cg[i];
  __syncthreads();
  for (int i = threadIdx.x; i < kSteps * 32 * 32; i += blockDim.x) {
    const int t = i & 3, k = (i >> 2) & 7, n = (i >> 5) & 31, ks = i >> 10;
    acan[i] = xs[4 * (8 * ks + k) + 4 * n + t];
  }
  for (int i = threadIdx.x; i < kSteps * 8 * 32; i += blockDim.x) {
    const int t = i & 3, k = (i >> 2) & 7, n = (i >> 5) & 7, ks = i >> 8;
    bcan[i] = cs[n * kPitch + 4 * (8 * ks + k) + t];
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  {   // prefill the accumulator with 7.0 so that "the MMA wrote nothing" is distinguishable from "it wrote zeros"
    const uint32_t seven = __float_as_uint(7.0f);
    const uint32_t taddr0 = tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
                 ::"r"(taddr0), "r"(seven) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (threadIdx.x < 32 && elect_one()) {
    uint32_t idesc = idesc_tf32_mn(128, 32);
    if (!(mode & 1)) idesc &= ~(1u << 15);
    if (!(mode & 2)) idesc &= ~(1u << 16);
    for (int ks = 0; ks < kSteps; ++ks) {
      uint64_t adesc = desc_mn_none(smem_u32(xs + 32 * ks), 128, 16);              // 8 rows r = 32 samples per step
      uint64_t bdesc = desc_mn_none(smem_u32(cs + 32 * ks), 128, kPitch * 4);
      if (mode & 8) {
        adesc = desc_mn_none(smem_u32(acan + ks * 1024), (mode & 16) ? 4096 : 128, 128);
        bdesc = desc_mn_none(smem_u32(bcan + ks * 256), (mode & 16) ? 1024 : 128, 128);
      }
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((mode & 4) ? 1 : ks) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  __syncwarp();
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t v[32];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int c = 0; c < 32; ++c) out[(warp * 32 + lane) * 32 + c] = __uint_as_float(v[c]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main(int argc, char** argv) {
  float *hx = (float*)malloc(kXLen * 4), *hc = (float*)malloc(kRows * kPitch * 4), *ho = (float*)malloc(128 * 32 * 4);
  srand(7);
  for (int i = 0; i < kXLen; ++i) hx[i] = (float)(rand() % 9 - 4);
  for (int i = 0; i < kRows * kPitch; ++i) hc[i] = (float)(rand() % 5 - 2);
  float *dx, *dc, *dout;
  cudaMalloc(&dx, kXLen * 4); cudaMalloc(&dc, kRows * kPitch * 4); cudaMalloc(&dout, 128 * 32 * 4);
  cudaMemcpy(dx, hx, kXLen * 4, cudaMemcpyHostToDevice); cudaMemcpy(dc, hc, kRows * kPitch * 4, cudaMemcpyHostToDevice);
  const size_t smem = (kXLen + kRows * kPitch + kSteps * 32 * 32 + kSteps * 8 * 32) * 4 + 1024;
  cudaFuncSetAttribute(hankel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int mode = argc > 1 ? atoi(argv[1]) : 3;
  hankel_kernel<<<1, 128, smem>>>(dx, dc, dout, mode);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(ho, dout, 128 * 32 * 4, cudaMemcpyDeviceToHost);
  int bad = 0; double maxerr = 0;
  for (int u = 0; u < 128; ++u)
    for (int n = 0; n < 32; ++n) {
      const int s = n / 4, j = n % 4;
      double ref = 0;
      for (int r = 0; r < 8 * kSteps; ++r) ref += (double)hc[s * kPitch + 4 * r + j] * hx[4 * r + u];
      const double err = fabs(ref - ho[u * 32 + n]);
      if (err > maxerr) maxerr = err;
      if (err > 1e-3) { if (bad < 5) printf("mismatch u=%d s=%d j=%d: got %g want %g\n", u, s, j, ho[u * 32 + n], ref); ++bad; }
    }
  printf("mode %d sample outputs: %g %g %g %g | %g %g\n", mode, ho[0], ho[1], ho[2], ho[33], ho[64 * 32], ho[127 * 32 + 31]);
  printf("zero-copy Hankel MN-major UMMA: %d / %d mismatches, max abs err %g -> %s\n", bad, 128 * 32, maxerr, bad ? "FAIL" : "PASS");
  return bad != 0;
}
