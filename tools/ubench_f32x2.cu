// Microbenchmark: issue/pipe cost of the L1 inner-loop instruction mixes on sm_100a, including the packed
// fp32x2 forms (add.f32x2 / fma.rn.f32x2).  Every variant processes the same 32 (window, lag) elements per
// "step" from registers; the result is reported as SM cycles per element per SMSP-warp (ideal for a plain
// two-instruction mix = 2.0 at one issue per cycle).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_f32x2 tools/ubench_f32x2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ float fset_gt(float x, float w) {
  float d; asm("set.gt.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(x), "f"(w)); return d;
}
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
  float d; asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}

constexpr int LT = 8;
constexpr float BIG = 1.2676506e30f;   // 2^100

// MODE 0: FSETP + @P FADD (current kernel)        1: FSET + FFMA2          2: FFMA.SAT + FFMA2
//      3: half FSET / half FFMA.SAT + FFMA2        4: FFMA.SAT + FFMA (scalar, FMA pipe only)
//      5: pure FFMA2 stream                        6: pure FFMA stream (same flop count as 5)
//      7: fwd  FADD + FADD|.|                      8: fwd  FADD2 (sub) + 2 FADD|.|
template <int MODE>
__global__ void __launch_bounds__(256) k(const float* __restrict__ xin, const float* __restrict__ cin, float* out, int iters) {
  __shared__ __align__(16) float xs[2048 + 64];
  __shared__ __align__(16) float cs[2048 + 64];
  for (int i = threadIdx.x; i < 2048 + 64; i += blockDim.x) { xs[i] = xin[i]; cs[i] = cin[i]; }
  __syncthreads();
  float w[LT], wb[LT], acc[LT], acc_b[LT];
#pragma unroll
  for (int i = 0; i < LT; ++i) { w[i] = xin[threadIdx.x + i] * 0.5f; wb[i] = -w[i] * BIG; acc[i] = 0.f; acc_b[i] = 0.f; }
  const int off = (threadIdx.x >> 5) * 4;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 4
    for (int t = 0; t < 2048; t += 4) {
      float xv[LT + 4];
      const float4 a0 = *reinterpret_cast<const float4*>(xs + t + off);
      const float4 a1 = *reinterpret_cast<const float4*>(xs + t + off + 4);
      const float4 a2 = *reinterpret_cast<const float4*>(xs + t + off + 8);
      xv[0] = a0.x; xv[1] = a0.y; xv[2] = a0.z; xv[3] = a0.w; xv[4] = a1.x; xv[5] = a1.y; xv[6] = a1.z; xv[7] = a1.w;
      xv[8] = a2.x; xv[9] = a2.y; xv[10] = a2.z; xv[11] = a2.w;
      const float4 c4 = *reinterpret_cast<const float4*>(cs + t + off);
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < LT; ++i) {
          if (xv[i] > w[i]) acc[i] += c4.x;
          if (xv[i + 1] > w[i]) acc[i] += c4.y;
          if (xv[i + 2] > w[i]) acc[i] += c4.z;
          if (xv[i + 3] > w[i]) acc[i] += c4.w;
        }
      } else if (MODE >= 1 && MODE <= 3) {
        const uint64_t c01 = pack2(c4.x, c4.y), c23 = pack2(c4.z, c4.w);
#pragma unroll
        for (int i = 0; i < LT; ++i) {
          float i0, i1, i2, i3;
          const bool alu = MODE == 1 || (MODE == 3 && (i & 1));
          if (alu) {
            i0 = fset_gt(xv[i], w[i]); i1 = fset_gt(xv[i + 1], w[i]);
            i2 = fset_gt(xv[i + 2], w[i]); i3 = fset_gt(xv[i + 3], w[i]);
          } else {
            i0 = fma_sat(xv[i], BIG, wb[i]); i1 = fma_sat(xv[i + 1], BIG, wb[i]);
            i2 = fma_sat(xv[i + 2], BIG, wb[i]); i3 = fma_sat(xv[i + 3], BIG, wb[i]);
          }
          uint64_t ac = pack2(acc[i], acc_b[i]);
          ac = fma2(c01, pack2(i0, i1), ac);
          ac = fma2(c23, pack2(i2, i3), ac);
          unpack2(ac, acc[i], acc_b[i]);
        }
      } else if (MODE == 4) {
#pragma unroll
        for (int i = 0; i < LT; ++i) {
          acc[i] = fmaf(c4.x, fma_sat(xv[i], BIG, wb[i]), acc[i]);
          acc[i] = fmaf(c4.y, fma_sat(xv[i + 1], BIG, wb[i]), acc[i]);
          acc[i] = fmaf(c4.z, fma_sat(xv[i + 2], BIG, wb[i]), acc[i]);
          acc[i] = fmaf(c4.w, fma_sat(xv[i + 3], BIG, wb[i]), acc[i]);
        }
      } else if (MODE >= 9) {
        // mixed: lags i < NF take the FMA-pipe indicator (FFMA.SAT + FFMA), the rest FSETP + predicated FADD
        constexpr int NF = MODE - 8;
#pragma unroll
        for (int i = 0; i < LT; ++i) {
          if (i < NF) {
            acc[i] = fmaf(c4.x, fma_sat(xv[i], BIG, wb[i]), acc[i]);
            acc[i] = fmaf(c4.y, fma_sat(xv[i + 1], BIG, wb[i]), acc[i]);
            acc[i] = fmaf(c4.z, fma_sat(xv[i + 2], BIG, wb[i]), acc[i]);
            acc[i] = fmaf(c4.w, fma_sat(xv[i + 3], BIG, wb[i]), acc[i]);
          } else {
            if (xv[i] > w[i]) acc[i] += c4.x;
            if (xv[i + 1] > w[i]) acc[i] += c4.y;
            if (xv[i + 2] > w[i]) acc[i] += c4.z;
            if (xv[i + 3] > w[i]) acc[i] += c4.w;
          }
        }
      } else if (MODE == 5) {
        const uint64_t c01 = pack2(c4.x, c4.y), c23 = pack2(c4.z, c4.w);
#pragma unroll
        for (int i = 0; i < LT; ++i) {
          uint64_t ac = pack2(acc[i], acc_b[i]);
          ac = fma2(c01, pack2(xv[i], xv[i + 1]), ac);
          ac = fma2(c23, pack2(xv[i + 2], xv[i + 3]), ac);
          ac = fma2(c23, pack2(xv[i], xv[i + 1]), ac);
          ac = fma2(c01, pack2(xv[i + 2], xv[i + 3]), ac);
          unpack2(ac, acc[i], acc_b[i]);
        }
      } else if (MODE == 6) {
#pragma unroll
        for (int i = 0; i < LT; ++i) {
          acc[i] = fmaf(c4.x, xv[i], acc[i]); acc_b[i] = fmaf(c4.y, xv[i + 1], acc_b[i]);
          acc[i] = fmaf(c4.z, xv[i + 2], acc[i]); acc_b[i] = fmaf(c4.w, xv[i + 3], acc_b[i]);
          acc[i] = fmaf(c4.z, xv[i], acc[i]); acc_b[i] = fmaf(c4.w, xv[i + 1], acc_b[i]);
          acc[i] = fmaf(c4.x, xv[i + 2], acc[i]); acc_b[i] = fmaf(c4.y, xv[i + 3], acc_b[i]);
        }
      } else if (MODE == 7) {
#pragma unroll
        for (int i = 0; i < LT; ++i) {
          acc[i] += fabsf(xv[i] - w[i]); acc_b[i] += fabsf(xv[i + 1] - w[i]);
          acc[i] += fabsf(xv[i + 2] - w[i]); acc_b[i] += fabsf(xv[i + 3] - w[i]);
        }
      } else if (MODE == 8) {
#pragma unroll
        for (int i = 0; i < LT; ++i) {
          const uint64_t nw = pack2(wb[i], wb[i]);
          float d0, d1, d2, d3;
          unpack2(add2(pack2(xv[i], xv[i + 1]), nw), d0, d1);
          unpack2(add2(pack2(xv[i + 2], xv[i + 3]), nw), d2, d3);
          acc[i] += fabsf(d0); acc_b[i] += fabsf(d1); acc[i] += fabsf(d2); acc_b[i] += fabsf(d3);
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LT; ++i) s += acc[i] + acc_b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// warp -> scheduler mapping probe: the current L1-backward mix at block sizes that are not a multiple of 4 warps
void run_blocks(const float* x, const float* c, float* out) {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int iters = 40;
  for (int threads = 128; threads <= 256; threads += 32)
    for (int cps = 2; cps <= 5; ++cps) {
      if (cps * threads > 1280) continue;   // 48 registers x 1280 threads fit an SM
      const int grid = sms * cps;
      k<0><<<grid, threads>>>(x, c, out, 2);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      k<0><<<grid, threads>>>(x, c, out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      const double warps_sm = cps * threads / 32.0;
      const double cyc = ms * 1e-3 * khz * 1e3 / ((double)iters * 512 * 32);   // cycles per element-step of one warp
      printf("block %3d x %d CTAs/SM = %4.1f warps/SM: %.3f ms, %.3f cycles/element if warps spread evenly (x%.2f if max-loaded scheduler paces: ceil(w/4)=%d)\n",
             threads, cps, warps_sm, ms, cyc / (warps_sm / 4.0), (warps_sm / 4.0) / ((int)((warps_sm + 3) / 4)), (int)((warps_sm + 3) / 4));
    }
}

template <int MODE>
void run(const char* name, const float* x, const float* c, float* out, int ctas_per_sm) {
  int dev = 0, sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int iters = 40;
  const int grid = sms * ctas_per_sm;
  k<MODE><<<grid, 256>>>(x, c, out, 2);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(x, c, out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double elems_per_warp = (double)iters * 512 * 32 * ((MODE == 5 || MODE == 6) ? 2 : 1);
  const double warps_per_smsp = ctas_per_sm * 8 / 4.0;
  const double cycles = ms * 1e-3 * khz * 1e3;
  printf("%-44s %8.3f ms  %.3f cycles/element/warp-slot (SMSP)  [%d CTAs/SM, clock attr %d kHz] %s\n", name, ms,
         cycles / (elems_per_warp * warps_per_smsp), ctas_per_sm, khz, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float *x, *c, *out;
  cudaMalloc(&x, 4096 * 4); cudaMalloc(&c, 4096 * 4); cudaMalloc(&out, 148 * 8 * 256 * 4);
  float h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = (float)((i * 2654435761u) >> 8 & 0xffff) / 65536.f - 0.5f;
  cudaMemcpy(x, h, sizeof h, cudaMemcpyHostToDevice);
  cudaMemcpy(c, h, sizeof h, cudaMemcpyHostToDevice);
  run_blocks(x, c, out);
  for (int cps = 2; cps <= 4; cps += 2) {
    run<0>("0 FSETP + @P FADD (current L1 bwd)", x, c, out, cps);
    run<1>("1 FSET + FFMA2", x, c, out, cps);
    run<2>("2 FFMA.SAT + FFMA2", x, c, out, cps);
    run<3>("3 half FSET / half FFMA.SAT, + FFMA2", x, c, out, cps);
    run<4>("4 FFMA.SAT + FFMA (scalar)", x, c, out, cps);
    run<5>("5 FFMA2 stream (per fp32 lane-op)", x, c, out, cps);
    run<6>("6 FFMA stream (per fp32 lane-op)", x, c, out, cps);
    run<7>("7 FADD + FADD|.| (current L1 fwd)", x, c, out, cps);
    run<8>("8 FADD2 + 2 FADD|.|", x, c, out, cps);
    run<9>("9  mixed: 1 of 8 lags on the FMA pipe", x, c, out, cps);
    run<10>("10 mixed: 2 of 8 lags on the FMA pipe", x, c, out, cps);
    run<11>("11 mixed: 3 of 8 lags on the FMA pipe", x, c, out, cps);
    run<12>("12 mixed: 4 of 8 lags on the FMA pipe", x, c, out, cps);
  }
  return 0;
}
