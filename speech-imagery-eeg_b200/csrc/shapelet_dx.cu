// Gradient of the pooled shapelet outputs with respect to the (normalised) input series, dL/dxn.
//
// Not on the training hot path: in the reference's loop the raw batch never requires grad
// (experiment_classification.py:315); this is what autograd hands a user who asks for input gradients (saliency maps,
// gradcheck) through Shapelet.forward — defined for the L1 distance (Shapelet.py:74), cosine (:64-66) and pearson
// (:11-19, :67-69).  The memory_efficient squared-L2 Function returns zeros for its input (Shapelet.py:40) and so does
// this library (the launcher skips the group).
//
// With a_t = dLoss/dd_t (what the pooling backward leaves in the coefficient workspace, times the mode's norm factor):
//   L1      dxn[tau] += (1/L) sum_k sum_l a_k[t] sign(xn[tau] - w_k[l])                      tau = t s + l
//   cosine  dxn[tau] += sum_k ( -inw_k sum_l coef_k[t] w_k[l]  +  xn[tau] sum_l beta_k[t] ),
//           coef = a inx (as stored),  beta = coef (1 - d) inx              d d/dx_l = -inx inw w_l + (1-d) inx^2 x_l
//   pearson dxn[tau] += sum_k ( -sum_l coef_k[t] (w_k[l] - mean w_k)  +  xn[tau] sum_l beta_k[t]  -  sum_l beta_k[t] mu[t] ),
//           coef = a / D (as stored),  beta = coef (1 - d) ||w-mean|| / ||x_w-mu||
// i.e. one full correlation of every coefficient row with its shapelet (E multiply-adds, like the forward) plus, for
// the cross-term modes, two sliding-window sums that come from one prefix scan per row.
//
// One CTA per (sample, channel) series row; the K shapelets are walked one at a time (coefficient row zero-padded by
// the shapelet length on both sides in shared memory, shapelet in shared memory); a thread owns four consecutive
// output samples and slides along the lags with two aligned 16-byte loads of the padded row and one of the shapelet per
// four lags (16 multiply-adds).  Strides > 1 (seq_len >= 3000) take a scalar loop.  The row of dxn is owned by its CTA:
// the result is ADDED to what is there (the caller zeroes dxn once and calls once per length group), no atomics.
#include "ign_common.cuh"

namespace ign {
namespace {

constexpr int kDxThreads = 256;

struct DxArgs {
  const float* xn; const float* W; const float* coef; const float* dstore; const float* st0; const float* st1;
  float* dxn;
  int B, M, T, Tp, K, L, s, Tw, Ts, SP;
  int PADL;      // front padding of the coefficient row in shared memory: L - 1 rounded up to 4
  int CPW;       // pitch of the padded row
};

__device__ __forceinline__ float warp_sum_dx(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// exclusive prefix of n values held in `buf` (shared), in place; n <= 16 * blockDim.x
__device__ void block_exclusive_scan(float* buf, int n, float* warp_tot) {
  const int per = (n + kDxThreads - 1) / kDxThreads;
  const int beg = min(n, (int)threadIdx.x * per), end = min(n, beg + per);
  float loc = 0.f;
  for (int i = beg; i < end; ++i) loc += buf[i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float inc = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const float v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  float base = 0.f;
  for (int w = 0; w < warp; ++w) base += warp_tot[w];
  float run = base + inc - loc;
  for (int i = beg; i < end; ++i) { const float v = buf[i]; buf[i] = run; run += v; }
  __syncthreads();
}

template <int DIST>
__global__ void __launch_bounds__(kDxThreads) shapelet_dx_kernel(const DxArgs a) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float s_warp[kDxThreads / 32];
  __shared__ float s_stat[2];
  const int m = blockIdx.x, b = blockIdx.y;
  float* xs = sm;                                   // [Tp]
  float* wsm = xs + a.Tp;                           // [L rounded to 4]: the shapelet (centred for pearson)
  float* cp = wsm + round_up(a.L, 4);               // [CPW]: zero | coef[0..Tw) | zero
  float* pb = cp + a.CPW;                           // [Tw + 1] prefix of beta           (cross-term modes)
  float* pm = pb + round_up(a.Tw + 1, 4);           // [Tw + 1] prefix of beta * mu      (pearson)
  const size_t row = (size_t)b * a.M + m;
  const float* xr = a.xn + row * a.Tp;
  for (int i = threadIdx.x; i < a.Tp; i += kDxThreads) xs[i] = xr[i];
  for (int i = threadIdx.x; i < a.CPW; i += kDxThreads) cp[i] = 0.f;
  float* out = a.dxn + row * a.Tp;
  const float invL = 1.f / (float)a.L;
  const int nq = (a.T + 3) / 4;                     // output quads

  for (int k = 0; k < a.K; ++k) {
    __syncthreads();                                // the previous shapelet's rows are no longer read
    // ---- stage the shapelet and its statistic
    const float* wk = a.W + ((size_t)k * a.M + m) * a.L;
    float s1 = 0.f, s2 = 0.f;
    for (int l = threadIdx.x; l < a.L; l += kDxThreads) { const float w = wk[l]; s1 += w; s2 = fmaf(w, w, s2); }
    s1 = warp_sum_dx(s1); s2 = warp_sum_dx(s2);
    if ((threadIdx.x & 31) == 0) { s_warp[threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < kDxThreads / 32; ++w) t += s_warp[w]; s_stat[0] = t / (float)a.L; }
    __syncthreads();
    const float wmean = DIST == IGN_DIST_PEARSON ? s_stat[0] : 0.f;
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = s2;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < kDxThreads / 32; ++w) t += s_warp[w]; s_stat[1] = t; }
    __syncthreads();
    // cosine: 1 / max(||w||, 1e-8); pearson: ||w - mean|| = sqrt(sum w^2 - L mean^2)
    const float wss = s_stat[1];
    const float inw = DIST == IGN_DIST_COSINE ? 1.f / fmaxf(sqrtf(wss), 1e-8f) : 0.f;
    const float wcn = DIST == IGN_DIST_PEARSON ? sqrtf(fmaxf(wss - (float)a.L * wmean * wmean, 0.f)) : 0.f;
    for (int l = threadIdx.x; l < round_up(a.L, 4); l += kDxThreads) wsm[l] = l < a.L ? wk[l] - wmean : 0.f;
    // ---- stage the coefficient row (padded) and the sliding-sum terms
    const size_t crow = (row * a.K + k) * a.Ts;
    for (int t = threadIdx.x; t < a.Tw; t += kDxThreads) {
      const float c = a.coef[crow + t];
      cp[a.PADL + t] = c;
      if (DIST != IGN_DIST_L1) {
        const float d = a.dstore[crow + t];
        const float s0 = a.st0[row * a.SP + t];
        float beta;
        if (DIST == IGN_DIST_COSINE) beta = c * (1.f - d) * s0;                       // s0 = 1 / max(||x_w||, 1e-8)
        else beta = s0 > 0.f ? c * (1.f - d) * wcn / s0 : 0.f;                       // s0 = ||x_w - mu||
        pb[t] = beta;
        if (DIST == IGN_DIST_PEARSON) pm[t] = beta * a.st1[row * a.SP + t];
      }
    }
    if (DIST != IGN_DIST_L1 && threadIdx.x == 0) { pb[a.Tw] = 0.f; if (DIST == IGN_DIST_PEARSON) pm[a.Tw] = 0.f; }
    __syncthreads();
    if (DIST != IGN_DIST_L1) {
      block_exclusive_scan(pb, a.Tw + 1, s_warp);                                     // pb[t] = sum_{u<t} beta_u
      if (DIST == IGN_DIST_PEARSON) block_exclusive_scan(pm, a.Tw + 1, s_warp);
    }
    // ---- correlation: out[tau] = sum_l f(coef[(tau - l) / s], w[l], x[tau])
    if (a.s == 1) {
      for (int q = threadIdx.x; q < nq; q += kDxThreads) {
        const int tau0 = 4 * q;
        const float4 x4 = *reinterpret_cast<const float4*>(xs + tau0);
        const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int l0 = 0; l0 < a.L; l0 += 4) {
          // padded-row indices tau0 + i - (l0 + dl) + PADL for i, dl in 0..3: the two aligned float4s around `base`
          const int base = tau0 + a.PADL - l0;
          const float4 lo = *reinterpret_cast<const float4*>(cp + base - 4);
          const float4 hi = *reinterpret_cast<const float4*>(cp + base);
          const float4 w4 = *reinterpret_cast<const float4*>(wsm + l0);
          const float cw[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};      // cw[4 + j] = cp[base + j]
          const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int dl = 0; dl < 4; ++dl) {
            if (l0 + dl >= a.L) break;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float c = cw[4 + i - dl];
              if (DIST == IGN_DIST_L1) {
                const float d = xv[i] - wv[dl];
                acc[i] += d > 0.f ? c : (d < 0.f ? -c : 0.f);                         // sign(0) = 0, as torch's abs backward
              } else {
                acc[i] = fmaf(c, wv[dl], acc[i]);
              }
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int tau = tau0 + i;
          if (tau >= a.T) break;
          float g;
          if (DIST == IGN_DIST_L1) {
            g = acc[i] * invL;
          } else {
            // windows that cover tau: t in [max(0, tau - L + 1), min(tau, Tw - 1)]
            const int tl = max(0, tau - a.L + 1), th = min(tau, a.Tw - 1) + 1;
            const float sb = th > tl ? pb[th] - pb[tl] : 0.f;
            if (DIST == IGN_DIST_COSINE) g = -inw * acc[i] + xv[i] * sb;
            else g = -acc[i] + xv[i] * sb - (th > tl ? pm[th] - pm[tl] : 0.f);
          }
          out[tau] += g;
        }
      }
    } else {
      for (int tau = threadIdx.x; tau < a.T; tau += kDxThreads) {
        const float x = xs[tau];
        float acc = 0.f, sb = 0.f, smu = 0.f;
        for (int l = tau % a.s; l < a.L && l <= tau; l += a.s) {
          const int t = (tau - l) / a.s;
          if (t >= a.Tw) continue;
          const float c = cp[a.PADL + t];
          if (DIST == IGN_DIST_L1) {
            const float d = x - wsm[l];
            acc += d > 0.f ? c : (d < 0.f ? -c : 0.f);
          } else {
            acc = fmaf(c, wsm[l], acc);
            sb += pb[t + 1] - pb[t];
            if (DIST == IGN_DIST_PEARSON) smu += pm[t + 1] - pm[t];
          }
        }
        float g;
        if (DIST == IGN_DIST_L1) g = acc * invL;
        else if (DIST == IGN_DIST_COSINE) g = -inw * acc + x * sb;
        else g = -acc + x * sb - smu;
        out[tau] += g;
      }
    }
  }
}

}  // namespace

size_t shapelet_dx_smem(const ign_shapelet_desc& d) {
  const int Tw = num_windows(d.T, d.L, d.stride);
  const int PADL = round_up(d.L - 1, 4) + 4;
  const int CPW = round_up(PADL + Tw + d.L + 8, 4);
  return ((size_t)d.Tp + round_up(d.L, 4) + CPW + 2 * (size_t)round_up(Tw + 1, 4)) * sizeof(float);
}

// coef: the coefficient rows [B,M,K,Ts] the pooling backward wrote (workspace of ign_shapelet_backward)
int launch_shapelet_dx(const ign_shapelet_desc& d, const float* xn, const float* st0, const float* st1, const float* W,
                       const float* coef, const float* dstore, float* dxn, cudaStream_t st) {
  if (d.dist == IGN_DIST_SQL2) return IGN_OK;      // the reference's ShapeletDistanceFunc returns zeros for its input
  DxArgs a;
  a.xn = xn; a.W = W; a.coef = coef; a.dstore = dstore; a.st0 = st0; a.st1 = st1; a.dxn = dxn;
  a.B = d.B; a.M = d.M; a.T = d.T; a.Tp = d.Tp; a.K = d.K; a.L = d.L; a.s = d.stride;
  a.Tw = num_windows(d.T, d.L, d.stride); a.Ts = round_up(a.Tw, 4); a.SP = stats_pitch(d.T, d.L, d.stride);
  a.PADL = round_up(d.L - 1, 4) + 4;               // base - 4 >= 0 for every lag block
  a.CPW = round_up(a.PADL + a.Tw + d.L + 8, 4);
  const size_t smem = shapelet_dx_smem(d);
  if (smem > (size_t)max_optin_smem() - 1024) { set_error("shapelet_backward_input: series of %d samples do not fit shared memory", d.T); return IGN_ERR_UNSUPPORTED; }
  if (d.B > 65535) { set_error("shapelet_backward_input: batch too large"); return IGN_ERR_INVALID; }
  const dim3 grid(d.M, d.B);
#define IGN_DX_LAUNCH(DV)                                                                                         \
  { IGN_CUDA(cudaFuncSetAttribute(shapelet_dx_kernel<DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    shapelet_dx_kernel<DV><<<grid, kDxThreads, smem, st>>>(a); }
  if (d.dist == IGN_DIST_L1) IGN_DX_LAUNCH(IGN_DIST_L1)
  else if (d.dist == IGN_DIST_COSINE) IGN_DX_LAUNCH(IGN_DIST_COSINE)
  else IGN_DX_LAUNCH(IGN_DIST_PEARSON)
#undef IGN_DX_LAUNCH
  IGN_CUDA(cudaGetLastError());
  return IGN_OK;
}

}  // namespace ign
