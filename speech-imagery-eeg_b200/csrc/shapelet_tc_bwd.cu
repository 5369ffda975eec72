// tcgen05 / TMEM backward contraction for the cross-term distances (cosine, pearson, squared-L2): sm_100a only.
//
//   G[k,l] = sum_b sum_t a[b,k,t] * x[b,m,t+l]        (a = the per-window coefficients written by pool_bwd_kernel)
//
// is the transpose problem of the forward (shapelet_tc.cu).  With P = 16 shifts, window t = P r + j:
//   Gsh[u,(k,j)] = sum_{(b,r)} x[b, P r + u] * a[b,k,P r + j],   u in [0, L+P-1),     G[k,l] = sum_j Gsh[l+j,(k,j)]
// i.e. a GEMM with  M' = u (the TMEM lanes; up to 256/N' tiles of 128 per item),  N' = P K (padded to 16),  and the
// contraction over the window-group rows (b, r) — the whole batch chunk accumulates into ONE resident set of
// accumulators, so there is no epilogue per tile: only one small drain per (channel, shapelet block, batch chunk).
// Longer shapelets split the lag axis into blocks of (128 MT - P) lags (the P extra lanes u feed the shift sum of the
// block's last lags); each lag block is its own item with its own accumulators.
// (P = 8 was the first version: twice the A' tiles, B' stages and MMA issue slots per sample at N' = 48 for K = 5.)
//   A'[u][(b,r)]     = x[b, 8 r + u]   transposed Hankel rows: gathered by the producer warps (one LDS.32 per element,
//                      conflict-free: consecutive lanes = consecutive u) straight into tensor memory (tcgen05.st), hi and
//                      lo = x - trunc_tf32(x) columns, exactly as in the forward
//   B'[(k,j)][(b,r)] = a[b,k,8 r + j]  the coefficient rows re-tiled (8 x 32 transposes) by the builder warps into the
//                      128B-swizzled K-major stage image (hi | lo), then fence.proxy.async
// Precision modes as in the forward: 3xTF32 (lo*hi + hi*lo + hi*hi) or single-pass TF32.
//
// Roles (704 threads, one persistent CTA per SM walking a contiguous range of (channel, shapelet block, chunk) items):
//   warps 0-11  A' producers: three groups of four warps (one warp per TMEM lane quarter), group g takes the stages
//               with (stage number % 3) == g — a stage is a latency chain (32 gathers, split, two tcgen05.st, wait::st,
//               arrive: ~1600 cycles for ~110 instructions), so three of them are kept in flight; warps 0-3 also drain
//               the accumulators
//   warps 12-19 B' builders: two groups of four warps that build ALTERNATE stages — a stage is a short latency chain
//               (wait for the slot, 16-32 scattered stores, proxy fence, arrive), so two of them have to be in flight for
//               the MMA thread not to wait for B' (it waited 45 % of the time at L = 100 with all eight warps on one stage)
//   warp 20     one elected thread issues tcgen05.mma / tcgen05.commit
//   warp 21     series-row loader: one elected thread streams the rows with 1-D bulk TMA copies into a ring of four
//               row buffers (completion counted on the row's mbarrier), so the DRAM latency of up to three rows is in
//               flight while one is consumed (a cp.async + wait per row exposed ~1 us per sample: more than the 960 MMA
//               cycles a sample needs at L = 100)
// The drain writes per-chunk partial sums part[chunk][k][m][l] in a fixed order; shapelet_bwd_finalize (shapelet_simt.cu)
// combines them exactly as for the FP32 engine: bit-reproducible, no float atomics.
#include "ign_common.cuh"

#include <math.h>

namespace ign {
namespace {

#include "tc_ptx.cuh"

constexpr int kBRows = 128;            // UMMA M
constexpr int kBKBlock = 32;           // window-group rows per stage (one 128-byte swizzle row of fp32)
#ifndef IGN_BWD_PROD_GROUPS
#define IGN_BWD_PROD_GROUPS 3
#endif
constexpr int kBProdGroups = IGN_BWD_PROD_GROUPS;                   // A' stages in production at once
constexpr int kBProdWarps = 4 * kBProdGroups, kBBuildWarps = 8, kBBuildGroups = 2;
constexpr int kBMmaWarp = kBProdWarps + kBBuildWarps, kBRowWarp = kBMmaWarp + 1;
constexpr int kBThreads = (kBRowWarp + 1) * 32;
constexpr int kBAStages = 4, kBBStages = 4;
constexpr int kBAStageCols = 64;       // 32 hi + 32 lo
constexpr int kBACol0 = 256;           // accumulators live in columns [0,256), A' stages behind them
constexpr int kBChunk = 32;            // samples per item
constexpr int kBRowBufs = 4;           // series-row ring
constexpr int kBMaxMT = 4;             // M' tiles (of 128 lanes u) per item, at most; MT * N' <= 256 accumulator columns

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
  int s;               // window stride: residue r of the series / shapelet is its own unit-stride contraction (one item
                       // dimension): G[k, q s + r] = sum_t a[k,t] x_r[t + q],  x_r[j] = x[j s + r],  q < Lr(r) = ceil((L - r) / s)
  int L0;              // lags of the longest residue: ceil(L / s)
  int Tw, Ts;          // windows, coefficient row pitch
  int RI, NKB;         // window groups (of 8) per sample, 32-row k-blocks per sample
  int P;               // shifts: 16 or 8
  int MT;              // 128-lane tiles over u = l + j of ONE lag block (the last block may use fewer)
  int nlb, lagstep;    // lag blocks and their step in lags: 128 MT - P
  int KG, nkb, N;      // shapelets per N' tile, shapelet blocks, N' = 8*KG rounded up to 16
  int XR;              // floats per series row in shared memory (zero padded)
  int nchunk, nitems;
  int split;
};

struct BwdTcArgs {
  const float* xn; const float* coef; float* part;
};

struct ItemCoord { int m, kblk, chunk, lb, mt, res, Lr; };   // mt: M' tiles of this item's lag block; res / Lr: residue, its lags
__device__ __forceinline__ ItemCoord item_coord(const BwdTcGeo& g, int w) {
  ItemCoord c;
  int mk = w / g.nchunk;
  c.chunk = w - mk * g.nchunk;
  const int mkl = mk;
  mk = mkl / g.nlb;
  c.lb = mkl - mk * g.nlb;
  const int mkr = mk;
  mk = mkr / g.s;
  c.res = mkr - mk * g.s;
  c.m = mk / g.nkb;
  c.kblk = mk - c.m * g.nkb;
  c.Lr = (g.L - c.res + g.s - 1) / g.s;
  // lag blocks past this residue's lags (the shorter residues of a strided group) get no tiles: nothing to do
  c.mt = max(0, min(g.MT, (c.Lr + g.P - 1 - c.lb * g.lagstep + kBRows - 1) / kBRows));
  if (c.lb * g.lagstep >= c.Lr) c.mt = 0;
  return c;
}

// KQ: shapelets per block the builder threads hold in registers (5 covers the reference's default K = 5 without the
// register pressure of the general 8)
constexpr int kBShifts = 16;           // P
template <int KQ>
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
  float* xrow = reinterpret_cast<float*>(ptr);                      // [kBRowBufs][XR]
  ptr += (size_t)kBRowBufs * g.XR * sizeof(float);
  const int TU = kBRows * g.MT + 8;                                 // pitch of one drained shift row
  float* tk = reinterpret_cast<float*>(ptr);                        // [8][TU] one shapelet's Gsh, transposed
  ptr += (size_t)kBShifts * TU * sizeof(float);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ptr);
  uint64_t* fullA = bars; uint64_t* emptyA = fullA + kBAStages;
  uint64_t* fullB = emptyA + kBAStages; uint64_t* emptyB = fullB + kBBStages;
  uint64_t* rowfull = emptyB + kBBStages; uint64_t* rowempty = rowfull + kBRowBufs;
  uint64_t* accfull = rowempty + kBRowBufs; uint64_t* accempty = accfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kBAStages; ++s) { mbar_init(&fullA[s], 4); mbar_init(&emptyA[s], 1); }
    for (int s = 0; s < kBBStages; ++s) { mbar_init(&fullB[s], kBBuildWarps / kBBuildGroups); mbar_init(&emptyB[s], 1); }
    for (int i = 0; i < kBRowBufs; ++i) { mbar_init(&rowfull[i], 1); mbar_init(&rowempty[i], kBProdWarps); }
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
        const int buf = nrow % kBRowBufs;
        long long t0 = BP_CLK();
        mbar_wait(&rowfull[buf], (nrow / kBRowBufs) & 1);
        BP_ADD(0, t0);
        const float* xs = xrow + (size_t)buf * g.XR;
        for (int kb = 0; kb < g.NKB; ++kb) {
          for (int mt = 0; mt < ic.mt; ++mt, ++ia) {
            const int s = sidx;
            const uint32_t ph = sph;
            if (++sidx == kBAStages) { sidx = 0; sph ^= 1; }
            if ((int)(ia % kBProdGroups) != grp) continue;
            // this lane's row of A': x[P (32 kb + c) + 128 mt + up], c = 0..31 (zero beyond the series)
            const float* src = xs + kBShifts * (kBKBlock * kb) + kBRows * mt + up + ic.lb * g.lagstep;
            float xv[32];                                            // the gather does not depend on the stage being free
#pragma unroll
            for (int c = 0; c < 32; ++c) xv[c] = src[kBShifts * c];
            t0 = BP_CLK();
            mbar_wait(&emptyA[s], ph ^ 1);
            BP_ADD(1, t0);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {                            // two halves of 16 columns: 64 live registers, not 96
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const float v = xv[16 * h + c];
                hi[c] = __float_as_uint(v);
                lo[c] = __float_as_uint(v - __uint_as_float(__float_as_uint(v) & 0xffffe000u));
              }
              tmem_st16(a_lane + s * kBAStageCols + 16 * h, hi);
              if (g.split) tmem_st16(a_lane + s * kBAStageCols + 32 + 16 * h, lo);
            }
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
#pragma unroll
            for (int h = 0; h < kBShifts / 8; ++h) {
              uint32_t v[8];
              asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                           : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                           : "r"(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * g.N + kl * kBShifts + 8 * h))
                           : "memory");
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 8; ++j) tk[(8 * h + j) * TU + kBRows * mt + up] = __uint_as_float(v[j]);
            }
          }
          bar_sync(3, 128);
          const int l0 = ic.lb * g.lagstep, nl = min(g.lagstep, ic.Lr - l0);          // this block's lags [l0, l0 + nl) of the residue
          for (int l = et; l < nl; l += 128) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < kBShifts; ++j) s += tk[j * TU + l + j];                // fixed order
            a.part[(((size_t)ic.chunk * g.K + k) * g.M + ic.m) * g.L + (size_t)(l0 + l) * g.s + ic.res] = s;
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
    // Group `bgrp` (four warps) builds the tiles with (tile number & 1) == bgrp.  Work split inside a tile: thread
    // (quad = warp in group, r = lane) owns, for EVERY shapelet kl of the block, the four shifts j = 4 quad .. 4 quad + 3
    // of window group r — one 16-byte load and four (hi, lo) pairs of scattered 4-byte stores per shapelet, the same
    // count for every thread.  All address arithmetic is hoisted: four swizzled row offsets per thread for the whole
    // launch, one source pointer per tile (+ kl * Ts per shapelet), shapelet offsets as immediates — the first version
    // recomputed them per element and spent ~490 instructions per warp and tile (ncu: 16 % of the builders' samples
    // issuing, 39 % in fixed-latency dependencies), which made the builders, not the tensor pipe, pace L <= 200.
    // The coefficients stream from HBM: the loads of the group's next tile are in flight (registers) while the current
    // one is written, and the tile after that is pulled into L2 (prefetch.global.L2) at the same time.
    static_assert(kBShifts == 16 && kBBuildWarps / kBBuildGroups == 4, "one quad of the 16 shifts per builder warp");
    const int bgrp = (warp - kBProdWarps) / (kBBuildWarps / kBBuildGroups);
    const int jq = (warp - kBProdWarps) % (kBBuildWarps / kBBuildGroups);   // which quad of shifts
    uint32_t rowoff[4];                                              // byte offset of (row 4 jq + jj, column lane) in a stage image
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) rowoff[jj] = sw128_off(4 * jq + jj, lane >> 2) + (uint32_t)((lane & 3) * 4);
    struct Cursor { int cw, cb, cb1, ckb, kvalid; const float* base; };   // base: coefficient row of (cb, m, first shapelet)
    Cursor cur{wbeg, 0, 0, 0, 0, nullptr};
    auto enter_item = [&](Cursor& c) {
      if (c.cw < wend) {
        const ItemCoord ic = item_coord(g, c.cw);
        c.cb = ic.chunk * kBChunk; c.cb1 = min(g.B, c.cb + kBChunk); c.ckb = 0;
        c.kvalid = min(g.KG, g.K - ic.kblk * g.KG);
        c.base = a.coef + (((size_t)c.cb * g.M + ic.m) * g.K + (size_t)ic.kblk * g.KG) * g.Ts;
      }
    };
    const size_t sample_pitch = (size_t)g.M * g.K * g.Ts;            // coefficient rows of the next sample, same (m, k)
    auto advance1 = [&](Cursor& c) {
      if (c.cw >= wend) return;
      if (++c.ckb == g.NKB) {
        c.ckb = 0;
        if (++c.cb == c.cb1) { ++c.cw; enter_item(c); } else c.base += sample_pitch;
      }
    };
    // this thread's source of shapelet 0 in the cursor's tile, or nullptr (past the end / pad windows)
    auto src_of = [&](const Cursor& c) -> const float* {
      const int t = kBShifts * (kBKBlock * c.ckb + lane) + 4 * jq;
      return (c.cw < wend && t < g.Ts) ? c.base + t : nullptr;
    };
    auto fetch = [&](const Cursor& c, float4 (&v)[KQ]) {
      const float* src = src_of(c);
#pragma unroll
      for (int q = 0; q < KQ; ++q) {
        v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src && q < c.kvalid) v[q] = __ldg(reinterpret_cast<const float4*>(src + (size_t)q * g.Ts));
      }
    };
    auto prefetch_l2 = [&](const Cursor& c) {
      const float* src = src_of(c);
#pragma unroll
      for (int q = 0; q < KQ; ++q)
        if (src && q < c.kvalid) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (size_t)q * g.Ts));
    };
    enter_item(cur);
    long long total = 0;                                             // tiles this CTA builds
    for (int w2 = wbeg; w2 < wend; ++w2) {
      const ItemCoord c2 = item_coord(g, w2);
      total += (long long)(min(g.B, c2.chunk * kBChunk + kBChunk) - c2.chunk * kBChunk) * g.NKB;
    }
    long long todo = (total - bgrp + 1) / 2;                         // tiles bgrp, bgrp + 2, ...
    int sb = bgrp; uint32_t phb = 0;                                 // kBBStages is even: the group's stages keep its parity
    static_assert(kBBStages % kBBuildGroups == 0, "a builder group must keep its stage parity");
    auto emit = [&](const float4 (&v)[KQ]) {                         // write this thread's share of one tile, publish it
      long long t0 = BP_CLK();
      mbar_wait(&emptyB[sb], phb ^ 1);
      BP_ADD(2, t0);
      uint8_t* hi = stage0 + (size_t)sb * stage_bytes;
      uint8_t* lo = hi + b_bytes;
#pragma unroll
      for (int q = 0; q < KQ; ++q) {                                 // pad shapelets (q >= KG) are never read back: rows of N'
        if (q >= g.KG) break;                                        // beyond 16 KG do not exist
        const float vv[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const uint32_t off = rowoff[jj] + (uint32_t)(q * kBShifts * 128);   // row kl * 16 + 4 jq + jj: (n & 7) does not depend on kl
          *reinterpret_cast<float*>(hi + off) = vv[jj];
          if (g.split) *reinterpret_cast<float*>(lo + off) = vv[jj] - __uint_as_float(__float_as_uint(vv[jj]) & 0xffffe000u);
        }
      }
      fence_proxy_async_smem();                                      // generic-proxy writes -> visible to the MMA's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&fullB[sb]);
      sb += kBBuildGroups;
      if (sb >= kBBStages) { sb -= kBBStages; phb ^= 1; }
    };
    if (todo > 0) {
      if (bgrp) advance1(cur);                                       // group 1 starts at tile 1
      if constexpr (KQ <= 5) {
        float4 va[KQ], vb[KQ];
        fetch(cur, va);
        advance1(cur); advance1(cur);
        while (true) {
          fetch(cur, vb);                                            // the group's next tile -> registers
          { Cursor nx = cur; advance1(nx); advance1(nx); prefetch_l2(nx); }   // the one after -> L2
          emit(va);
          if (--todo == 0) break;
          advance1(cur); advance1(cur);
          fetch(cur, va);
          { Cursor nx = cur; advance1(nx); advance1(nx); prefetch_l2(nx); }
          emit(vb);
          if (--todo == 0) break;
          advance1(cur); advance1(cur);
        }
      } else {                                                       // 6..8 shapelets per block: one register set (80 registers per
        float4 va[KQ];                                               // thread), the next two tiles are pulled into L2 instead
        { Cursor nx = cur; advance1(nx); advance1(nx); prefetch_l2(nx); }
        while (true) {
          fetch(cur, va);
          advance1(cur); advance1(cur);
          { Cursor nx = cur; advance1(nx); advance1(nx); prefetch_l2(nx); }
          emit(va);
          if (--todo == 0) break;
        }
      }
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
    for (int i = lane; i < kBRowBufs * g.XR; i += 32) xrow[i] = 0.f;  // pads behind Tp stay zero for ever
    fence_proxy_async_smem();                                        // generic zero fill before the async-proxy row copies
    __syncwarp();
    if (g.s == 1) {
      if (elect_one()) {
        const uint32_t row_bytes = (uint32_t)g.Tp * sizeof(float);   // Tp % 4 == 0: a multiple of 16, rows 16-byte aligned
        int nrow = 0;
        for (int w = wbeg; w < wend; ++w) {
          const ItemCoord ic = item_coord(g, w);
          const int b0 = ic.chunk * kBChunk, b1 = min(g.B, b0 + kBChunk);
          for (int b = b0; b < b1; ++b, ++nrow) {
            const int buf = nrow % kBRowBufs;
            long long t0 = BP_CLK();
            if (nrow >= kBRowBufs) mbar_wait(&rowempty[buf], ((nrow / kBRowBufs) - 1) & 1);   // row nrow - kBRowBufs is consumed
            BP_ADD(7, t0);
            mbar_arrive_expect_tx(&rowfull[buf], row_bytes);
            tma_bulk_g2s(xrow + (size_t)buf * g.XR, a.xn + ((size_t)b * g.M + ic.m) * g.Tp, row_bytes, &rowfull[buf]);
          }
        }
      }
      __syncwarp();
    } else {
      // strided group: the item's residue row x_r[j] = x[j s + r] is gathered with 4-byte cp.async (the row is re-read
      // from L2 by the s residue items); rows nrow-1 and nrow-2 stay in flight while row nrow is issued
      int nrow = 0, done = 0;                                        // rows issued / rows published
      for (int w = wbeg; w < wend; ++w) {
        const ItemCoord ic = item_coord(g, w);
        const int b0 = ic.chunk * kBChunk, b1 = min(g.B, b0 + kBChunk);
        const int nq = (g.T - ic.res + g.s - 1) / g.s;              // samples of this residue
        for (int b = b0; b < b1; ++b, ++nrow) {
          const int buf = nrow % kBRowBufs;
          if (nrow >= kBRowBufs) mbar_wait(&rowempty[buf], ((nrow / kBRowBufs) - 1) & 1);
          const float* src = a.xn + ((size_t)b * g.M + ic.m) * g.Tp + ic.res;
          float* dst = xrow + (size_t)buf * g.XR;
          for (int j = lane; j < nq; j += 32)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst + j)), "l"(src + (size_t)j * g.s) : "memory");
          for (int j = nq + lane; j < g.XR && j < nq + 64; j += 32) dst[j] = 0.f;   // the slots a longer residue row left behind
          cp_async_commit();
          if (nrow - done >= 2) {                                    // publish the oldest row in flight
            asm volatile("cp.async.wait_group 2;" ::: "memory");
            __threadfence_block();
            __syncwarp();
            if (lane == 0) mbar_arrive(&rowfull[done % kBRowBufs]);
            ++done;
          }
        }
      }
      cp_async_wait_all();
      __threadfence_block();
      __syncwarp();
      for (; done < nrow; ++done)
        if (lane == 0) mbar_arrive(&rowfull[done % kBRowBufs]);
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
  g.s = d.stride; g.L0 = ceil_div(d.L, d.stride);
  g.Tw = num_windows(d.T, d.L, d.stride); g.Ts = round_up(g.Tw, 4);
  g.nkb = ceil_div(d.K, 8); g.KG = ceil_div(d.K, g.nkb);
  // 16 shifts when the whole shapelet then fits one lag block (half the A' tiles, B' stages and MMA issue slots per
  // sample), else 8 shifts with up to four 128-lane tiles per block
  // (an 8-shift variant, N' = 48 with four 128-lane tiles per block, was the first version: at config 2, K = 5 it needs
  //  16 A' tiles per sample at L = 500 against 10 with two lag blocks of 16 shifts, and twice the stages everywhere)
  g.P = kBShifts;
  g.N = kBShifts * g.KG;                                             // <= 128: at least two accumulator tiles
  const int mtmax = max(1, min(kBMaxMT, kBACol0 / g.N));
  g.RI = ceil_div(g.Tw, g.P); g.NKB = ceil_div(g.RI, kBKBlock);
  g.lagstep = mtmax * kBRows - g.P;
  g.nlb = ceil_div(g.L0, g.lagstep);
  g.MT = min(mtmax, ceil_div(g.L0 + g.P - 1, kBRows));
  // one (residue) row: its samples, or the furthest A' gather (window groups + lag lanes), whichever is longer
  g.XR = round_up(max(g.s == 1 ? d.Tp : ceil_div(d.T, g.s), g.P * kBKBlock * g.NKB + (g.nlb - 1) * g.lagstep + kBRows * g.MT) + 8, 4);
  g.nchunk = ceil_div(d.B, kBChunk);
  g.nitems = d.M * g.nkb * g.s * g.nlb * g.nchunk;
  g.split = d.precision == IGN_PREC_3XTF32 ? 1 : 0;
}

size_t bwd_tc_smem(const BwdTcGeo& g) {
  const size_t stage = (size_t)(g.N * 128) * (g.split ? 2 : 1);
  return kBBStages * stage + (size_t)kBRowBufs * g.XR * 4 + (size_t)g.P * (kBRows * g.MT + 8) * 4 +
         (2 * kBAStages + 2 * kBBStages + 2 * kBRowBufs + 2) * 8 + 16 + 1024;
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

// the tensor-core contraction covers: cross-term distances in the tcgen05 operand modes, any stride (one item per
// residue: s unit-stride contractions) and any shapelet length (lag blocks), as long as the row ring fits shared memory
bool shapelet_bwd_tc_supported(const ign_shapelet_desc& d) {
  if (d.dist == IGN_DIST_L1) return false;
  if (d.precision != IGN_PREC_3XTF32 && d.precision != IGN_PREC_TF32) return false;
  if (num_windows(d.T, d.L, d.stride) <= 0) return false;
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
  auto kern = g.KG <= 5 ? shapelet_bwd_tc_kernel<5> : shapelet_bwd_tc_kernel<8>;
  IGN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BwdTcArgs a{xn, coef, part};
  const int grid = min(sm_count(), g.nitems);
  kern<<<grid, kBThreads, smem, st>>>(g, a);
  IGN_CUDA(cudaGetLastError());
  return IGN_OK;
}

}  // namespace ign
