// Instance norm + layout change, and the sliding-window prefix-sum pass.  Both HBM-bound.
//
//   instnorm_kernel : x[B,T,M] (channels last) -> xn[B,M,Tp] (time contiguous, 16-byte rows)
//                     reference: Shapelet.py:186-187  (x-mean_T)/(std_T(unbiased)+1e-8)
//   prefix_kernel   : fp64 exclusive prefix sums of xn and xn^2 per series, from which every window's
//                     sum / squared norm is one subtraction (norm terms of cosine/pearson/sql2,
//                     Shapelet.py:11-19,28,64-66)
#include "ign_common.cuh"

namespace ign {

namespace {

constexpr int kNormThreads = 256;
constexpr int kNormWarps = kNormThreads / 32;

// One CTA per (32-channel tile, sample).  x is read three times (mean, variance, normalise); the tile is
// T*128 bytes, so passes 2 and 3 hit L2/L1 and DRAM sees one read + one write.
__global__ void __launch_bounds__(kNormThreads) instnorm_kernel(const float* __restrict__ x,
                                                                float* __restrict__ xn,
                                                                float* __restrict__ mean_out,
                                                                float* __restrict__ rstd_out, int T, int M,
                                                                int Tp) {
  __shared__ float part[kNormWarps][32];
  __shared__ float s_mean[32], s_den[32];
  __shared__ float tile[kNormWarps][32][33];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m0 = blockIdx.x * 32, b = blockIdx.y;
  const int m = m0 + lane;
  const bool mv = m < M;
  const float* xb = x + (size_t)b * T * M;

  // pass 1: mean
  float acc = 0.f;
  for (int t = warp; t < T; t += kNormWarps) acc += mv ? __ldg(xb + (size_t)t * M + m) : 0.f;
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kNormWarps; ++w) s += part[w][lane];
    s_mean[lane] = s / (float)T;
  }
  __syncthreads();
  const float mu = s_mean[lane];
  // pass 2: unbiased variance about the mean (two-pass, like torch.std)
  acc = 0.f;
  for (int t = warp; t < T; t += kNormWarps) {
    float v = mv ? __ldg(xb + (size_t)t * M + m) - mu : 0.f;
    acc = fmaf(v, v, acc);
  }
  __syncthreads();
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kNormWarps; ++w) s += part[w][lane];
    float sd = sqrtf(s / (float)(T - 1));   // T==1 -> NaN, as torch.std
    s_den[lane] = sd + 1e-8f;
    if (mv) {
      if (mean_out) mean_out[(size_t)b * M + m] = mu;
      if (rstd_out) rstd_out[(size_t)b * M + m] = 1.f / (sd + 1e-8f);
    }
  }
  __syncthreads();
  // pass 3: normalise + transpose through a padded smem tile; rows of xn are written 128 B at a time
  const int ntile = (Tp + 31) / 32;
  for (int tt = warp; tt < ntile; tt += kNormWarps) {
    const int t0 = tt * 32;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      int t = t0 + i;
      float v = (mv && t < T) ? __ldg(xb + (size_t)t * M + m) : 0.f;
      tile[warp][i][lane] = v;
    }
    __syncwarp();
    const int t = t0 + lane;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      int mm = m0 + i;
      if (mm < M && t < Tp) {
        float v = tile[warp][lane][i];
        float o = (t < T) ? (v - s_mean[i]) / s_den[i] : 0.f;
        xn[((size_t)b * M + mm) * Tp + t] = o;
      }
    }
    __syncwarp();
  }
}

__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// One warp per series row; each lane owns 4 consecutive samples per 128-sample chunk.
__global__ void __launch_bounds__(256) prefix_kernel(const float* __restrict__ xn,
                                                     double* __restrict__ pre1,
                                                     double* __restrict__ pre2, int rows, int T, int Tp) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = xn + (size_t)row * Tp;
  double* o1 = pre1 + (size_t)row * (T + 1);
  double* o2 = pre2 + (size_t)row * (T + 1);
  double c1 = 0.0, c2 = 0.0;
  if (lane == 0) { o1[0] = 0.0; o2[0] = 0.0; }
  for (int base = 0; base < T; base += 128) {
    const int j = base + lane * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < Tp) v = *reinterpret_cast<const float4*>(xr + j);   // Tp % 4 == 0, pad columns are zero
    double a0 = v.x, a1 = a0 + (double)v.y, a2 = a1 + (double)v.z, a3 = a2 + (double)v.w;
    double q0 = (double)v.x * v.x, q1 = q0 + (double)v.y * v.y, q2 = q1 + (double)v.z * v.z,
           q3 = q2 + (double)v.w * v.w;
    double s1 = warp_incl_scan(a3, lane), s2 = warp_incl_scan(q3, lane);
    double e1 = c1 + s1 - a3, e2 = c2 + s2 - q3;   // exclusive offset of this lane
    if (j + 0 < T) { o1[j + 1] = e1 + a0; o2[j + 1] = e2 + q0; }
    if (j + 1 < T) { o1[j + 2] = e1 + a1; o2[j + 2] = e2 + q1; }
    if (j + 2 < T) { o1[j + 3] = e1 + a2; o2[j + 3] = e2 + q2; }
    if (j + 3 < T) { o1[j + 4] = e1 + a3; o2[j + 4] = e2 + q3; }
    c1 += __shfl_sync(0xffffffffu, s1, 31);
    c2 += __shfl_sync(0xffffffffu, s2, 31);
  }
}

}  // namespace

int launch_instnorm(const float* x, float* xn, float* mean, float* rstd, int B, int T, int M,
                    cudaStream_t st) {
  dim3 grid(ceil_div(M, 32), B);
  instnorm_kernel<<<grid, kNormThreads, 0, st>>>(x, xn, mean, rstd, T, M, padded_len(T));
  IGN_CUDA(cudaGetLastError());
  return IGN_OK;
}

int launch_prefix(const float* xn, double* pre1, double* pre2, int B, int M, int T, cudaStream_t st) {
  const int rows = B * M;
  prefix_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(xn, pre1, pre2, rows, T, padded_len(T));
  IGN_CUDA(cudaGetLastError());
  return IGN_OK;
}

}  // namespace ign
