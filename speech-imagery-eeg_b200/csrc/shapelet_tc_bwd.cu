// tcgen05 / TMEM backward contraction for the cross-term distances (cosine, pearson, squared-L2): sm_100a only.
//
//   G[k,l] = sum_b sum_t a[b,k,t] * x[b,m,t+l]        (a = the per-window coefficients written by pool_bwd_kernel)
//
// is the transpose problem of the forward (shapelet_tc.cu).  With P = 8 shifts, window t = 8 r + j:
//   Gsh[u,(k,j)] = sum_{(b,r)} x[b, 8 r + u] * a[b,k,8 r + j],   u in [0, L+7),       G[k,l] = sum_j Gsh[l+j,(k,j)]
// i.e. a GEMM with  M' = u (the TMEM lanes; up to 4 tiles of 128 per item),  N' = 8 K (padded to 16),  and the
// contraction over the window-group rows (b, r) — the whole batch chunk accumulates into ONE resident set of
// accumulators, so there is no epilogue per tile: only one small drain per (channel, shapelet block, batch chunk).
// Shapelets longer than 505 lags split the lag axis into blocks of 504 lags (512 lanes u, the 8 extra ones feed the
// shift sum of the block's last lags); each (lag block) is its own item with its own accumulators.
//   A'[u][(b,r)]     = x[b, 8 r + u]   transposed Hankel rows: gathered by the producer warps (one LDS.32 per element,
//                      conflict-free: consecutive lanes = consecutive u) straight into tensor memory (tcgen05.st), hi and
//                      lo = x - trunc_tf32(x) columns, exactly as in the forward
//   B'[(k,j)][(b,r)] = a[b,k,8 r + j]  the coefficient rows re-tiled (8 x 32 transposes) by the builder warps into the
//                      128B-swizzled K-major stage image (hi | lo), then fence.proxy.async
// Precision modes as in the forward: 3xTF32 (lo*hi + hi*lo + hi*hi) or single-pass TF32.
//
// Roles (448 threads, one persistent CTA per SM walking a contiguous range of (channel, shapelet block, chunk) items):
//   warps 0-7   A' producers (two per TMEM lane quarter, alternating stages); warps 0-3 also drain the accumulators
//   warps 8-11  B' builders
//   warp 12     one elected thread issues tcgen05.mma / tcgen05.commit
//   warp 13     series-row loader (cp.async, double buffered, mbarrier hand-off)
// The drain writes per-chunk partial sums part[chunk][k][m][l] in a fixed order; shapelet_bwd_finalize (shapelet_simt.cu)
// combines them exactly as for the FP32 engine: bit-reproducible, no float atomics.
#include "ign_common.cuh"

#include <math.h>

namespace ign {
namespace {

#include "tc_ptx.cuh"

constexpr int kBShifts = 8;            // P
constexpr int kBRows = 128;            // UMMA M
constexpr int kBKBlock = 32;           // window-group rows per stage (one 128-byte swizzle row of fp32)
constexpr int kBProdWarps = 8, kBBuildWarps = 8;
constexpr int kBMmaWarp = kBProdWarps + kBBuildWarps, kBRowWarp = kBMmaWarp + 1;
constexpr int kBThreads = (kBRowWarp + 1) * 32;
constexpr int kBAStages = 4, kBBStages = 4;
constexpr int kBAStageCols = 64;       // 32 hi + 32 lo
constexpr int kBACol0 = 256;           // accumulators live in columns [0,256), A' stages behind them
constexpr int kBChunk = 32;            // samples per item
constexpr int kBMaxMT = 4;             // M' tiles (of 128 lanes u) per item: 4 accumulators of N' <= 64 columns
constexpr int kBLagStep = kBMaxMT * kBRows - kBShifts;   // 504 lags per lag block

// role-level wait accounting (debug builds only: -DIGN_TC_PROFILE); slots: 0 prod wait row, 1 prod wait emptyA,
// 2 build wait emptyB, 3 mma wait fullB, 4 mma wait fullA, 5 mma wait accempty, 6 drain wait accfull, 7 loader wait
// rowempty, 8 prod total, 9 build total, 10 mma total, 11 loader total
#ifdef IGN_TC_PROFILE
__device__ unsigned long long g_bwd_prof[16];
#define BP_CLK() clock64()
#define BP_ADD(slot, t0) (bp_local[slot] += (unsigned long long)(clock64() - (t0)))
#define BP_DECL() unsigned long long bp_local[16] = {0}
#define BP_FLUSH() do { for (int i_ = 0; i_ < 16; ++i_) if (bp_local[i_]) atomicAdd(&g_bwd_prof[i_], bp_local[i_]); } while (0)
#else
#define BP_CLK() 0ll
#define BP_ADD(slot, t0) ((void)(t0))
#define BP_DECL() ((void)0)
#define BP_FLUSH() ((void)0)
#endif

__device__ __forceinline__ bool bp_any(long long) { return true; }

struct BwdTcGeo {
  int B, M, T, Tp, K, L;
  int Tw, Ts;          // windows, coefficient row pitch
  int RI, NKB;         // window groups (of 8) per sample, 32-row k-blocks per sample
  int MT;              // 128-lane tiles over u = l + j of ONE lag block (<= kBMaxMT; the last block may use fewer)
  int nlb;             // lag blocks: ceil(L / 504)
  int KG, nkb, N;      // shapelets per N' tile, shapelet blocks, N' = 8*KG rounded up to 16
  int XR;              // floats per series row in shared memory (zero padded)
  int nchunk, nitems;
  int split;
};

struct BwdTcArgs {
  const float* xn; const float* coef; float* part;
};

struct ItemCoord { int m, kblk, chunk, lb, mt; };   // mt: M' tiles of this item's lag block
__device__ __forceinline__ ItemCoord item_coord(const BwdTcGeo& g, int w) {
  ItemCoord c;
  int mk = w / g.nchunk;
  c.chunk = w - mk * g.nchunk;
  const int mkl = mk;
  mk = mkl / g.nlb;
  c.lb = mkl - mk * g.nlb;
  c.m = mk / g.nkb;
  c.kblk = mk - c.m * g.nkb;
  c.mt = min(g.MT, (g.L + kBShifts - 1 - c.lb * kBLagStep + kBRows - 1) / kBRows);
  return c;
}

__global__ void __launch_bounds__(kBThreads, 1) shapelet_bwd_tc_kernel(const BwdTcGeo g, const BwdTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  BP_DECL();
  const long long t_entry = BP_CLK();
  const int wbeg = (int)(((long long)g.nitems * blockIdx.x) / gridDim.x);
  const int wend = (int)(((long long)g.nitems * (blockIdx.x + 1)) / gridDim.x);

  const int b_bytes = g.N * 128;
  const int stage_bytes = b_bytes * (g.split ? 2 : 1);
  uint8_t* stage0 = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ptr = stage0 + (size_t)kBBStages * stage_bytes;
  float* xrow = reinterpret_cast<float*>(ptr);                      // [2][XR]
  ptr += (size_t)2 * g.XR * sizeof(float);
  const int TU = kBRows * g.MT + 8;                                 // pitch of one drained shift row
  float* tk = reinterpret_cast<float*>(ptr);                        // [8][TU] one shapelet's Gsh, transposed
  ptr += (size_t)kBShifts * TU * sizeof(float);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ptr);
  uint64_t* fullA = bars; uint64_t* emptyA = fullA + kBAStages;
  uint64_t* fullB = emptyA + kBAStages; uint64_t* emptyB = fullB + kBBStages;
  uint64_t* rowfull = emptyB + kBBStages; uint64_t* rowempty = rowfull + 2;
  uint64_t* accfull = rowempty + 2; uint64_t* accempty = accfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kBAStages; ++s) { mbar_init(&fullA[s], 4); mbar_init(&emptyA[s], 1); }
    for (int s = 0; s < kBBStages; ++s) { mbar_init(&fullB[s], kBBuildWarps); mbar_init(&emptyB[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&rowfull[i], 1); mbar_init(&rowempty[i], kBProdWarps); }
    mbar_init(accfull, 1); mbar_init(accempty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // dummy shapelet rows (N' padding) of the B' ring must hold finite values: zero the ring once
  for (int i = threadIdx.x; i < kBBStages * stage_bytes / 16; i += blockDim.x)
    reinterpret_cast<float4*>(stage0)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp == kBMmaWarp) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kBProdWarps) {
    // =================================================================== A' PRODUCERS (+ drain by warps 0-3)
    const int quarter = warp & 3, grp = warp >> 2;
    const int up = quarter * 32 + lane;                              // lane within the M' tile
    const uint32_t a_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + kBACol0;
    uint32_t ia = 0;                                                 // global A' stage counter
    int sidx = 0; uint32_t sph = 0;
    int nrow = 0;                                                    // samples consumed (row buffer parity)
    int item_no = 0;
    for (int w = wbeg; w < wend; ++w, ++item_no) {
      const ItemCoord ic = item_coord(g, w);
      const int b0 = ic.chunk * kBChunk, b1 = min(g.B, b0 + kBChunk);
      for (int b = b0; b < b1; ++b, ++nrow) {
        const int buf = nrow & 1;
        long long t0 = BP_CLK();
        mbar_wait(&rowfull[buf], (nrow >> 1) & 1);
        BP_ADD(0, t0);
        const float* xs = xrow + (size_t)buf * g.XR;
        for (int kb = 0; kb < g.NKB; ++kb) {
          for (int mt = 0; mt < ic.mt; ++mt, ++ia) {
            const int s = sidx;
            const uint32_t ph = sph;
            if (++sidx == kBAStages) { sidx = 0; sph ^= 1; }
            if ((int)(ia & 1) != grp) continue;
            // this lane's row of A': x[8 (32 kb + c) + 128 mt + up], c = 0..31 (zero beyond the series)
            const float* src = xs + 8 * (kBKBlock * kb) + kBRows * mt + up + ic.lb * kBLagStep;
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float v = src[8 * c];
              hi[c] = __float_as_uint(v);
              lo[c] = __float_as_uint(v - __uint_as_float(__float_as_uint(v) & 0xffffe000u));
            }
            t0 = BP_CLK();
            mbar_wait(&emptyA[s], ph ^ 1);
            BP_ADD(1, t0);
            tc_fence_after();
            tmem_st32(a_lane + s * kBAStageCols, hi);
            if (g.split) tmem_st32(a_lane + s * kBAStageCols + 32, lo);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&fullA[s]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&rowempty[buf]);
      }
      // ---- drain (warps 0-3): G[k,l] = sum_j Gsh[l+j,(k,j)], one shapelet at a time through shared memory
      if (warp < 4) {
        long long t1 = BP_CLK();
        mbar_wait(accfull, item_no & 1);
        BP_ADD(6, t1);
        tc_fence_after();
        const int k0 = ic.kblk * g.KG;
        const int et = threadIdx.x;                                  // 0..127
        for (int kl = 0; kl < g.KG; ++kl) {
          const int k = k0 + kl;
          if (k >= g.K) break;                                       // uniform
          for (int mt = 0; mt < ic.mt; ++mt) {
            uint32_t v[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * g.N + kl * kBShifts))
                         : "memory");
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < kBShifts; ++j) tk[j * TU + kBRows * mt + up] = __uint_as_float(v[j]);
          }
          bar_sync(3, 128);
          const int l0 = ic.lb * kBLagStep, nl = min(kBLagStep, g.L - l0);            // this block's lags [l0, l0 + nl)
          for (int l = et; l < nl; l += 128) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < kBShifts; ++j) s += tk[j * TU + l + j];                // fixed order
            a.part[(((size_t)ic.chunk * g.K + k) * g.M + ic.m) * g.L + l0 + l] = s;
          }
          bar_sync(3, 128);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(accempty);
      }
    }
  } else if (warp < kBMmaWarp) {
    // =================================================================== B' BUILDERS
    // Coefficient rows come straight from global memory (L2 / HBM latency): the loads of tile i+1 are issued before
    // tile i is written, and before waiting for its stage to be free.
    const int bt = threadIdx.x - kBProdWarps * 32;                   // 0..127
    int sb = 0; uint32_t phb = 0;
    // flat walk over (item, sample, k-block)
    int cw = wbeg, cb = 0, ckb = 0, cb1 = 0;
    ItemCoord cic{0, 0, 0};
    auto enter_item = [&]() {
      if (cw < wend) { cic = item_coord(g, cw); cb = cic.chunk * kBChunk; cb1 = min(g.B, cb + kBChunk); ckb = 0; }
    };
    auto fetch = [&](float4 (&v)[2][2]) {                            // this thread's (up to) two tasks of the current tile
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        v[q][0] = make_float4(0.f, 0.f, 0.f, 0.f); v[q][1] = v[q][0];
        const int task = bt + q * kBBuildWarps * 32;
        if (cw >= wend || task >= g.KG * kBKBlock) continue;
        const int kl = task >> 5, r = task & 31;
        const int k = cic.kblk * g.KG + kl;
        const int t = kBShifts * (kBKBlock * ckb + r);
        if (k < g.K) {
          const float* src = a.coef + (((size_t)cb * g.M + cic.m) * g.K + k) * g.Ts + t;
          if (t < g.Ts) v[q][0] = __ldg(reinterpret_cast<const float4*>(src));
          if (t + 4 < g.Ts) v[q][1] = __ldg(reinterpret_cast<const float4*>(src) + 1);
        }
      }
    };
    auto advance = [&]() {
      if (++ckb == g.NKB) { ckb = 0; if (++cb == cb1) { ++cw; enter_item(); } }
    };
    enter_item();
    long long todo = 0;                                              // tiles this CTA builds
    for (int w2 = wbeg; w2 < wend; ++w2) {
      const ItemCoord c2 = item_coord(g, w2);
      todo += (long long)(min(g.B, c2.chunk * kBChunk + kBChunk) - c2.chunk * kBChunk) * g.NKB;
    }
    auto emit = [&](const float4 (&v)[2][2]) {                       // write one tile from registers, publish it
      long long t0 = BP_CLK();
      mbar_wait(&emptyB[sb], phb ^ 1);
      BP_ADD(2, t0);
      uint8_t* st = stage0 + (size_t)sb * stage_bytes;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int task = bt + q * kBBuildWarps * 32;
        if (task >= g.KG * kBKBlock) continue;
        const int kl = task >> 5, r = task & 31;
        const float vv[8] = {v[q][0].x, v[q][0].y, v[q][0].z, v[q][0].w, v[q][1].x, v[q][1].y, v[q][1].z, v[q][1].w};
#pragma unroll
        for (int j = 0; j < kBShifts; ++j) {
          const int n = kl * kBShifts + j;
          const uint32_t off = sw128_off(n, r >> 2) + (uint32_t)((r & 3) * 4);
          *reinterpret_cast<float*>(st + off) = vv[j];
          if (g.split) *reinterpret_cast<float*>(st + b_bytes + off) = vv[j] - __uint_as_float(__float_as_uint(vv[j]) & 0xffffe000u);
        }
      }
      fence_proxy_async_smem();                                      // generic-proxy writes -> visible to the MMA's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&fullB[sb]);
      if (++sb == kBBStages) { sb = 0; phb ^= 1; }
    };
    // three register sets, unrolled by three: the loads of tiles i+1 and i+2 are in flight while tile i is written
    // (the coefficients stream from HBM: one tile of prefetch distance left the latency exposed, builders 85 % busy at
    // L=100), and no register copy between the sets ever waits for a load
    float4 va[2][2], vb[2][2], vc[2][2];
    fetch(va);
    advance();
    fetch(vb);
    while (true) {
      if (cw < wend) advance();
      fetch(vc);
      emit(va);
      if (--todo == 0) break;
      if (cw < wend) advance();
      fetch(va);
      emit(vb);
      if (--todo == 0) break;
      if (cw < wend) advance();
      fetch(vb);
      emit(vc);
      if (--todo == 0) break;
    }
  } else if (warp == kBMmaWarp) {
    // =================================================================== MMA ISSUER
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_tf32(kBRows, g.N);
      const uint32_t bdesc0 = (smem_u32(stage0) & 0x3FFFFu) >> 4;
      const uint32_t bstage16 = (uint32_t)stage_bytes >> 4, bimg16 = (uint32_t)b_bytes >> 4;
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      int item_no = 0;
      for (int w = wbeg; w < wend; ++w, ++item_no) {
        const ItemCoord ic = item_coord(g, w);
        const int b0 = ic.chunk * kBChunk, b1 = min(g.B, b0 + kBChunk);
        long long t0 = BP_CLK();
        mbar_wait(accempty, (item_no & 1) ^ 1);                      // the previous item's accumulators are drained
        BP_ADD(5, t0);
        tc_fence_after();
        for (int b = b0; b < b1; ++b) {
          for (int kb = 0; kb < g.NKB; ++kb) {
            t0 = BP_CLK();
            mbar_wait(&fullB[sb], phb);
            BP_ADD(3, t0);
            const uint32_t bd_hi = bdesc0 + (uint32_t)sb * bstage16, bd_lo = bd_hi + bimg16;
            const bool first = (b == b0) && (kb == 0);
            for (int mt = 0; mt < ic.mt; ++mt) {
              t0 = BP_CLK();
              mbar_wait(&fullA[sa], pha);
              BP_ADD(4, t0);
              tc_fence_after();
              const uint32_t d_tmem = tmem_base + (uint32_t)(mt * g.N);
              const uint32_t a_hi = tmem_base + kBACol0 + sa * kBAStageCols, a_lo = a_hi + 32;
#pragma unroll
              for (int k8 = 0; k8 < kBKBlock / 8; ++k8) {
                if (g.split) {
                  umma_tf32_ts_lo(d_tmem, a_lo + k8 * 8, bd_hi + 2 * k8, idesc, !(first && k8 == 0));
                  umma_tf32_ts_lo(d_tmem, a_hi + k8 * 8, bd_lo + 2 * k8, idesc, 1);
                  umma_tf32_ts_lo(d_tmem, a_hi + k8 * 8, bd_hi + 2 * k8, idesc, 1);
                } else {
                  umma_tf32_ts_lo(d_tmem, a_hi + k8 * 8, bd_hi + 2 * k8, idesc, !(first && k8 == 0));
                }
              }
              umma_commit(&emptyA[sa]);
              if (++sa == kBAStages) { sa = 0; pha ^= 1; }
            }
            umma_commit(&emptyB[sb]);
            if (++sb == kBBStages) { sb = 0; phb ^= 1; }
          }
        }
        umma_commit(accfull);
      }
    }
    __syncwarp();
  } else {
    // =================================================================== SERIES-ROW LOADER
    for (int i = lane; i < 2 * g.XR; i += 32) xrow[i] = 0.f;          // pads behind Tp stay zero for ever
    __syncwarp();
    const int chunks = g.Tp / 4;
    int nrow = 0;
    for (int w = wbeg; w < wend; ++w) {
      const ItemCoord ic = item_coord(g, w);
      const int b0 = ic.chunk * kBChunk, b1 = min(g.B, b0 + kBChunk);
      for (int b = b0; b < b1; ++b, ++nrow) {
        const int buf = nrow & 1;
        long long t0 = BP_CLK();
        if (nrow >= 2) mbar_wait(&rowempty[buf], ((nrow >> 1) - 1) & 1);
        BP_ADD(7, t0);
        const float* src = a.xn + ((size_t)b * g.M + ic.m) * g.Tp;
        float* dst = xrow + (size_t)buf * g.XR;
        for (int c = lane; c < chunks; c += 32) cp_async16(dst + c * 4, src + c * 4);
        cp_async_commit();
        cp_async_wait_all();
        __threadfence_block();
        __syncwarp();
        if (lane == 0) mbar_arrive(&rowfull[buf]);
      }
    }
  }

  if (lane == 0 || (warp == kBMmaWarp)) {
    if (warp == 0) { BP_ADD(8, t_entry); BP_FLUSH(); }
    else if (warp == kBProdWarps && lane == 0) { BP_ADD(9, t_entry); BP_FLUSH(); }
    else if (warp == kBMmaWarp) { if (bp_any(BP_CLK())) { BP_ADD(10, t_entry); BP_FLUSH(); } }
    else if (warp == kBRowWarp) { BP_ADD(11, t_entry); BP_FLUSH(); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kBMmaWarp) tmem_dealloc(tmem_base, 512);
}

void bwd_tc_geo(const ign_shapelet_desc& d, BwdTcGeo& g) {
  g.B = d.B; g.M = d.M; g.T = d.T; g.Tp = d.Tp; g.K = d.K; g.L = d.L;
  g.Tw = num_windows(d.T, d.L, 1); g.Ts = round_up(g.Tw, 4);
  g.RI = ceil_div(g.Tw, kBShifts); g.NKB = ceil_div(g.RI, kBKBlock);
  g.nlb = ceil_div(d.L, kBLagStep);
  g.MT = min(kBMaxMT, ceil_div(d.L + kBShifts - 1, kBRows));
  g.nkb = ceil_div(d.K, 8); g.KG = ceil_div(d.K, g.nkb); g.N = round_up(kBShifts * g.KG, 16);
  g.XR = round_up(max(d.Tp, kBShifts * kBKBlock * g.NKB + (g.nlb - 1) * kBLagStep + kBRows * g.MT) + 8, 4);
  g.nchunk = ceil_div(d.B, kBChunk);
  g.nitems = d.M * g.nkb * g.nlb * g.nchunk;
  g.split = d.precision == IGN_PREC_3XTF32 ? 1 : 0;
}

size_t bwd_tc_smem(const BwdTcGeo& g) {
  const size_t stage = (size_t)(g.N * 128) * (g.split ? 2 : 1);
  return kBBStages * stage + (size_t)2 * g.XR * 4 + (size_t)kBShifts * (kBRows * g.MT + 8) * 4 +
         (2 * kBAStages + 2 * kBBStages + 6) * 8 + 16 + 1024;
}

}  // namespace

int bwd_tc_profile_read(unsigned long long* host16, int reset) {
#ifdef IGN_TC_PROFILE
  IGN_CUDA(cudaMemcpyFromSymbol(host16, g_bwd_prof, sizeof(unsigned long long) * 16));
  if (reset) { unsigned long long z[16] = {0}; IGN_CUDA(cudaMemcpyToSymbol(g_bwd_prof, z, sizeof(z))); }
  return IGN_OK;
#else
  (void)host16; (void)reset;
  set_error("library built without -DIGN_TC_PROFILE");
  return IGN_ERR_UNSUPPORTED;
#endif
}

// the tensor-core contraction covers: cross-term distances, unit stride, the tcgen05 operand modes; any shapelet
// length (lag blocks of 504) as long as one series row and its zero padding fit shared memory twice
bool shapelet_bwd_tc_supported(const ign_shapelet_desc& d) {
  if (d.dist == IGN_DIST_L1 || d.stride != 1) return false;
  if (d.precision != IGN_PREC_3XTF32 && d.precision != IGN_PREC_TF32) return false;
  if (num_windows(d.T, d.L, 1) <= 0) return false;
  BwdTcGeo g;
  bwd_tc_geo(d, g);
  if (g.MT * g.N > kBACol0) return false;
  return bwd_tc_smem(g) <= (size_t)max_optin_smem();
}

int shapelet_bwd_tc_chunks(const ign_shapelet_desc& d) { return ceil_div(d.B, kBChunk); }

// coef [B,M,K,Ts] from pool_bwd_kernel -> part [nchunk,K,M,L] (nchunk = shapelet_bwd_tc_chunks)
int launch_shapelet_bwd_tc(const ign_shapelet_desc& d, const float* xn, const float* coef, float* part, cudaStream_t st) {
  BwdTcGeo g;
  bwd_tc_geo(d, g);
  const size_t smem = max(bwd_tc_smem(g), (size_t)118 * 1024);       // > half the SM: one CTA per SM (it owns all of TMEM)
  IGN_CUDA(cudaFuncSetAttribute(shapelet_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BwdTcArgs a{xn, coef, part};
  const int grid = min(sm_count(), g.nitems);
  shapelet_bwd_tc_kernel<<<grid, kBThreads, smem, st>>>(g, a);
  IGN_CUDA(cudaGetLastError());
  return IGN_OK;
}

}  // namespace ign
