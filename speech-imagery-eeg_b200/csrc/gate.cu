// InterpGN Gini gate + expert mixture, forward and backward (InterpGN.py:44-52).
//   q = softmax(s), eta = (C*sum q^2 - 1)/(C-1), optional hard gate eta=1 where eta>gating_value,
//   out = eta*s + (1-eta)*z.      eta is NOT detached in the reference: its gradient flows into s.
// [B,C] with C = 3..39: one warp per sample, launch-latency bound — fusing the ~10 eager kernels of the
// reference into one launch each way is the whole optimisation.
#include "ign_common.cuh"

#include <math.h>

namespace ign {
namespace {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// softmax statistics of one row: returns max and 1/sum exp, and gini = sum q^2
__device__ __forceinline__ void row_stats(const float* s, int C, int lane, float& mx, float& inv, float& gini) {
  mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, s[c]);
  mx = wmax(mx);
  float z = 0.f;
  for (int c = lane; c < C; c += 32) z += expf(s[c] - mx);
  z = wsum(z);
  inv = 1.f / z;
  float g2 = 0.f;
  for (int c = lane; c < C; c += 32) { float q = expf(s[c] - mx) * inv; g2 = fmaf(q, q, g2); }
  gini = wsum(g2);
}

__global__ void __launch_bounds__(128) gate_fwd_kernel(const float* __restrict__ s, const float* __restrict__ z,
                                                       float* __restrict__ out, float* __restrict__ eta_out,
                                                       int B, int C, int use_gate, float gv) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* sr = s + (size_t)b * C;
  const float* zr = z + (size_t)b * C;
  float mx, inv, gini;
  row_stats(sr, C, lane, mx, inv, gini);
  float eta = ((float)C * gini - 1.f) / (float)(C - 1);
  if (use_gate && eta > gv) eta = 1.f;
  for (int c = lane; c < C; c += 32) out[(size_t)b * C + c] = eta * sr[c] + (1.f - eta) * zr[c];
  if (lane == 0) eta_out[b] = eta;
}

__global__ void __launch_bounds__(128) gate_bwd_kernel(const float* __restrict__ s, const float* __restrict__ z,
                                                       const float* __restrict__ go, const float* __restrict__ ge,
                                                       float* __restrict__ gs, float* __restrict__ gz, int B,
                                                       int C, int use_gate, float gv) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* sr = s + (size_t)b * C;
  const float* zr = z + (size_t)b * C;
  const float* gr = go + (size_t)b * C;
  float mx, inv, gini;
  row_stats(sr, C, lane, mx, inv, gini);
  const float eta_raw = ((float)C * gini - 1.f) / (float)(C - 1);
  const bool fired = use_gate && eta_raw > gv;
  const float eta = fired ? 1.f : eta_raw;
  float de = 0.f;   // dLoss/d eta
  for (int c = lane; c < C; c += 32) de = fmaf(gr[c], sr[c] - zr[c], de);
  de = wsum(de);
  if (ge) de += ge[b];
  if (fired) de = 0.f;   // the hard gate replaces eta by the constant 1
  const float coef = de * 2.f * (float)C / (float)(C - 1);
  for (int c = lane; c < C; c += 32) {
    const float q = expf(sr[c] - mx) * inv;
    gs[(size_t)b * C + c] = eta * gr[c] + coef * q * (q - gini);
    gz[(size_t)b * C + c] = (1.f - eta) * gr[c];
  }
}

}  // namespace

int launch_gate_fwd(const float* s, const float* z, float* out, float* eta, int B, int C, int use_gate,
                    float gv, cudaStream_t st) {
  gate_fwd_kernel<<<ceil_div(B, 4), 128, 0, st>>>(s, z, out, eta, B, C, use_gate, gv);
  IGN_CUDA(cudaGetLastError());
  return IGN_OK;
}

int launch_gate_bwd(const float* s, const float* z, const float* go, const float* ge, float* gs,
                    float* gz, int B, int C, int use_gate, float gv, cudaStream_t st) {
  gate_bwd_kernel<<<ceil_div(B, 4), 128, 0, st>>>(s, z, go, ge, gs, gz, B, C, use_gate, gv);
  IGN_CUDA(cudaGetLastError());
  return IGN_OK;
}

}  // namespace ign
