// Shapelet diversity regulariser, forward and backward (ShapeBottleneckModel.loss, Shapelet.py:217-230):
//   div_g = mean over (channel m, shapelet a, shapelet b) of [a != b] * exp(-|| w_b - w_a + 1e-6 ||_2)
// (the 1e-6 is nn.PairwiseDistance's eps, added to the difference; the mean runs over all M*K*K entries).
// The reference builds the [M,K,K,L] difference tensor with eager broadcasting: ~130 tiny kernels per training
// step for the four length groups (1.4 ms of a 24.8 ms step at config 2) and M*K^2*L*4 bytes of temporaries
// (250 GB at K=1000, L=500).  Here: one launch each way per group, nothing materialised beyond [M,K,K].
//   forward   CTA = (channel, 16x16 block of shapelet pairs), difference norms accumulated over L from
//             shared-memory tiles; writes coef[m,a,b] = exp(-d)/d (what the backward needs) and one partial sum
//             per CTA (summed by the caller in a fixed order: deterministic, no atomics)
//   backward  CTA = (channel, shapelet c):
//             dW[c,m,l] = -g/(M K K) * sum_{a != c} [ coef[m,a,c] (w_c - w_a + eps)[l] - coef[m,c,a] (w_a - w_c + eps)[l] ]
#include "ign_common.cuh"

#include <math.h>

namespace ign {
namespace {

constexpr int kTile = 16;      // shapelets per pair-block side
constexpr int kChunk = 32;     // lags per shared-memory tile
constexpr float kPairEps = 1e-6f;

__global__ void __launch_bounds__(kTile * kTile) diversity_fwd_kernel(const float* __restrict__ W, float* __restrict__ coef,
                                                                      float* __restrict__ partial, int K, int M, int L) {
  __shared__ float wa[kTile][kChunk + 1], wb[kTile][kChunk + 1];
  __shared__ float red[kTile * kTile / 32];
  const int m = blockIdx.x, a0 = blockIdx.y * kTile, b0 = blockIdx.z * kTile;
  const int ta = threadIdx.x / kTile, tb = threadIdx.x % kTile;
  const int a = a0 + ta, b = b0 + tb;
  float acc = 0.f;
  for (int l0 = 0; l0 < L; l0 += kChunk) {
    __syncthreads();
    for (int i = threadIdx.x; i < kTile * kChunk; i += blockDim.x) {
      const int r = i / kChunk, c = i - r * kChunk;
      const int l = l0 + c;
      wa[r][c] = (a0 + r < K && l < L) ? __ldg(W + ((size_t)(a0 + r) * M + m) * L + l) : 0.f;
      wb[r][c] = (b0 + r < K && l < L) ? __ldg(W + ((size_t)(b0 + r) * M + m) * L + l) : 0.f;
    }
    __syncthreads();
    const int n = min(kChunk, L - l0);
    for (int c = 0; c < n; ++c) {
      const float v = wb[tb][c] - wa[ta][c] + kPairEps;
      acc = fmaf(v, v, acc);
    }
  }
  float e = 0.f;
  if (a < K && b < K) {
    float cf = 0.f;
    if (a != b) {
      const float d = sqrtf(acc);
      e = expf(-d);
      cf = e / fmaxf(d, 1e-30f);
    }
    coef[((size_t)m * K + a) * K + b] = cf;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < kTile * kTile / 32; ++i) s += red[i];      // fixed order
    partial[((size_t)m * gridDim.y + blockIdx.y) * gridDim.z + blockIdx.z] = s;
  }
}

__global__ void __launch_bounds__(128) diversity_bwd_kernel(const float* __restrict__ W, const float* __restrict__ coef,
                                                            const float* __restrict__ gout, float* __restrict__ dW,
                                                            int K, int M, int L) {
  extern __shared__ float cf[];                   // [2][K]: coef[m,a,c] and coef[m,c,a] for all a
  const int m = blockIdx.x, c = blockIdx.y;
  for (int a = threadIdx.x; a < K; a += blockDim.x) {
    cf[a] = coef[((size_t)m * K + a) * K + c];
    cf[K + a] = coef[((size_t)m * K + c) * K + a];
  }
  __syncthreads();
  const float scale = -__ldg(gout) / ((float)M * (float)K * (float)K);
  const float* wc = W + ((size_t)c * M + m) * L;
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const float x = wc[l];
    float g = 0.f;
    for (int a = 0; a < K; ++a) {                 // coef is 0 on the diagonal
      const float y = __ldg(W + ((size_t)a * M + m) * L + l);
      g = fmaf(cf[a], x - y + kPairEps, g);
      g = fmaf(-cf[K + a], y - x + kPairEps, g);
    }
    dW[((size_t)c * M + m) * L + l] = scale * g;
  }
}

}  // namespace

int diversity_blocks(int K) { return ceil_div(K, kTile); }

int launch_diversity_fwd(const float* W, float* coef, float* partial, int K, int M, int L, cudaStream_t st) {
  const int nb = diversity_blocks(K);
  diversity_fwd_kernel<<<dim3(M, nb, nb), kTile * kTile, 0, st>>>(W, coef, partial, K, M, L);
  IGN_CUDA(cudaGetLastError());
  return IGN_OK;
}

int launch_diversity_bwd(const float* W, const float* coef, const float* gout, float* dW, int K, int M, int L,
                         cudaStream_t st) {
  const size_t smem = (size_t)2 * K * sizeof(float);
  if (smem > 48 * 1024) { set_error("diversity_backward: K=%d shapelets per channel exceed the kernel's staging (6144)", K); return IGN_ERR_UNSUPPORTED; }
  diversity_bwd_kernel<<<dim3(M, K), 128, smem, st>>>(W, coef, gout, dW, K, M, L);
  IGN_CUDA(cudaGetLastError());
  return IGN_OK;
}

}  // namespace ign
