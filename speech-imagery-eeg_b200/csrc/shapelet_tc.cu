// tcgen05 / TMEM forward for the cross-term distances (cosine, pearson, squared-L2): sm_100a only.
//
// The window-by-shapelet cross term  cross[b,t,k] = sum_l x[b,m,t+l] * w[k,m,l]  (Shapelet.py:28,64-69) is a
// dense contraction with a Hankel (sliding-window) left operand and only K (=5) output columns per channel.
// Two re-arrangements make it a tensor-core GEMM without an im2col blow-up:
//
//  * shifted shapelets: with P = 16 shifts, out[(b,i), (k,j)] = sum_u x[b,m,16 i + u] * Wsh[(k,j), u],
//    Wsh[(k,j), u] = w[k, u-j] (zero outside the shapelet), u in [0, L+15].  N becomes 16*K (80 for K=5) and
//    the left operand becomes a plain strided matrix: row (b,i) is the series segment starting at 16 i.
//    Window t = 16 i + j lives in row i, column 16 k + j, so one TMEM row holds 16 consecutive windows of
//    each shapelet — exactly the 64-byte runs the distance store wants.
//  * the overlapping rows are never materialised in HBM, L2 or even shared memory: each producer thread owns
//    one accumulator row, reads its 32-sample segment of the raw series row (bank-skewed layout, conflict-free
//    LDS.128) and writes it straight into TENSOR MEMORY with tcgen05.st; the MMA takes A from TMEM
//    (tcgen05.mma [d], [a_tmem], b_desc), which dispatches at the 128*N/256-cycle floor (40 cycles for N = 80,
//    profiles/r1d_ubench_umma_dispatch.txt; A from shared memory costs 52).
//
// Precision: tcgen05 has no fp32 kind.  kind::tf32 reads fp32 bits and drops the low 13 mantissa bits, so
// the "hi" operand is the raw fp32 value and lo = x - trunc_tf32(x) is formed by the producers.
//   IGN_PREC_3XTF32: hi*hi + hi*lo + lo*hi  (fp32-equivalent: error ~2^-21).  Three N-column MMAs per k-step, or —
//                    for long tiles, which the issuing thread paces — two: A_hi x [B_hi | B_lo] (2N columns, the lo
//                    image simply follows the hi image in the stage) and A_lo x B_hi; the epilogue adds the halves
//   IGN_PREC_TF32  : hi*hi only             (own, looser tolerance)
//
//  * the shifted-shapelet operand depends only on (channel, k-block): a small pre-pass writes it to a workspace
//    already in the 128B-swizzled K-major tile image (hi | lo), and a dedicated warp streams one tile per
//    stage with a single 1-D bulk TMA copy (cp.async.bulk + mbarrier complete_tx).
//
// Strided windows (seq_len >= 3000, Shapelet.py:162): sum_l x[t s + l] w[l] = sum_r sum_q x_r[t + q] w_r[q] with the
// residues x_r[j] = x[j s + r], w_r[q] = w[q s + r] — s unit-stride cross terms that accumulate into the SAME tile, so
// their k-blocks simply concatenate (NKB = s * NKBr).  The series rows are streamed one RESIDUE at a time (a ring of
// (tile, residue) units, each RB rows of T/s samples gathered with 4-byte cp.async), so shared memory holds a few KB
// per unit instead of s whole rows per sample — which is what lets a tile keep its 4-8 samples at T = 4000.
//
// Persistent kernel: one CTA per SM walks a contiguous range of the (channel, shapelet block, sample tile) list,
// so there is no wave-quantisation tail and the TMEM allocation, barriers and the pipeline state live for the
// whole launch.
// TMEM map (512 columns): accumulators at 0 and 128 (double buffered), A stages at 256 + 64 s (hi 32 | lo 32).
// Two rings: A stages in tensor memory (4), B stages in shared memory (as deep as fits: the bulk-copy latency of a
// 20 KB tile is ~2000 cycles against 480 MMA cycles per stage); when the ring holds every k-block of a (channel,
// shapelet block) run the B tiles are loaded once per run and stay resident.
// Roles (608 threads):
//   warps 0-7   producers: tcgen05.st the A stage (hi = raw fp32, lo = x - trunc_tf32(x)), arrive on fullA[stage].
//               Warp w owns TMEM lanes 32 (w%4)..; the two warps of a lane quarter take alternate stages, so the
//               LDS -> split -> tcgen05.st -> wait latency of one stage hides under the other warp's stage.
//   warps 8-15  epilogue : tcgen05.ld the accumulator (two warps per lane quarter, alternate shapelets), cross ->
//               distance with the window norm terms, coalesced store of d, per-row minimum; the arg-min of a
//               (sample, shapelet) pair is formed with redux.sync inside the warp and published as one packed
//               64-bit word (ordered distance | window index: first index wins ties) to the warp's own cell — no
//               atomics; one rotating warp per tile merges the four lane-quarter cells and writes p / d_min /
//               arg-min, the other seven only bar.arrive and move on.
//   warp 16     one elected thread issues tcgen05.mma (M=128, N=16*KG, K=8) and tcgen05.commit
//   warp 17     one elected thread streams the B tiles (bulk TMA)
//   warp 18     loads the next tiles' series rows (cp.async into the bank-skewed layout) and hands them to the
//               producers through mbarriers — the producers never meet a CTA-wide barrier
// Pipelines: fullA/emptyA (producers <-> MMA), fullB/emptyB (TMA <-> MMA), TMEM full/empty per accumulator
// (MMA <-> epilogue).
#include "ign_common.cuh"

#include <math.h>

namespace ign {
namespace {

constexpr int kShifts = 16;            // P
constexpr int kRows = 128;             // UMMA M
constexpr int kKBlock = 32;            // fp32 elements per 128-byte swizzle row
constexpr int kProdWarps = 8, kEpiWarps = 8;
constexpr int kProducerThreads = kProdWarps * 32;
constexpr int kEpilogueThreads = kEpiWarps * 32;
constexpr int kMmaWarp = kProdWarps + kEpiWarps, kTmaWarp = kMmaWarp + 1, kRowWarp = kMmaWarp + 2;
constexpr int kThreadsTC = (kRowWarp + 1) * 32;
constexpr int kAStages = 4;            // A ring (tensor memory)
constexpr int kMaxBStages = 12;        // B ring (shared memory): as deep as fits, resident when it holds all k-blocks
constexpr int kAccCols = 128;          // TMEM columns reserved per accumulator (N <= 128)
constexpr int kAStageCols = 64;        // TMEM columns per A stage: 32 hi + 32 lo
constexpr int kMaxAcc = 4;             // accumulators in flight (2, or 3-4 when N is small and the tiles are short)
constexpr int kMaxRB = 8;              // samples per tile (bounds the arg-min cell arrays)
constexpr int kMaxRowBufs = 4;         // ring of series-row units: 2 for unit stride (one unit per tile), 4 for strided groups
constexpr unsigned long long kSlotEmpty = 0xffffffffffffffffull;

struct TcGeo {
  int B, M, T, Tp, K, L;
  int Tw, Ts;          // windows, dstore pitch
  int RI, RB;          // window groups (of 16) per sample IN ONE TILE (<= 128), samples per 128-row tile
  int RItot, nseg;     // window groups per sample in total; tiles per sample (1 unless a sample has more than 2048 windows:
                       // then RB = 1, tile `seg` covers groups [128 seg, 128 seg + 128) and the arg-min is merged across the
                       // sample's tiles with a 64-bit atomicMin on (ordered distance | window index), decoded by tc_finish_kernel)
  int KG, nkb;         // shapelets per N tile, number of shapelet blocks
  int N;               // 16*KG
  int s;               // window stride; residue r of series and shapelet is its own group of k-blocks
  int NKBr;            // 32-wide k-blocks per residue: ceil((ceil(L/s)+15)/32)
  int NKB;             // k-blocks per tile: s * NKBr
  int nrb;             // series-row units in the ring
  int XR;              // floats per series row in smem
  int nacc, accp;      // accumulators in flight and their TMEM column pitch
  int stack;           // 3xTF32 with [B_hi | B_lo] stacked along N: two MMAs per k-step (N and 2N) instead of three
  int nast, acol0;     // A ring depth in use (<= kAStages) and its first TMEM column
  int ncb;             // arg-min cell generations / finaliser barrier ids in flight: a power of two >= nacc + 2
  int nbs, resident;   // B ring depth; 1 = the ring holds every k-block of a (channel, shapelet block) run
  int tpm;             // sample tiles per (channel, shapelet block): ceil(B / RB)
  int ntiles;          // M * nkb * tpm
  int dist, pool, split;
  float eps;
};

struct TcArgs {
  const float* xn; const float* st0; const float* thr;   // st0: window statistics [B,M,SP]
  float* p; float* dmin; int* argmin; float* dstore;
  int SP;
  unsigned long long* packed;   // [B,K,M] cross-tile arg-min cells (nseg > 1 only)
  const uint8_t* btiles;   // [M][nkb][NKB][stage_bytes] pre-swizzled B stage images
  const float* wstat;      // [M][nkb][8] shapelet statistic of the distance mode (pre-pass)
};

#include "tc_ptx.cuh"   // mbarrier / TMA / tcgen05 wrappers (inside namespace ign { namespace {)

// Optional role-level cycle accounting (debug builds only: -DIGN_TC_PROFILE).
#ifdef IGN_TC_PROFILE
__device__ unsigned long long g_tc_prof[16];
#define TC_CLK() clock64()
#define TC_ADD(slot, t0) (tc_prof_local[slot] += (unsigned long long)(clock64() - (t0)))   // flushed once at CTA end
#define TC_DECL() unsigned long long tc_prof_local[16] = {0}
#define TC_FLUSH() do { for (int i_ = 0; i_ < 16; ++i_) if (tc_prof_local[i_]) atomicAdd(&g_tc_prof[i_], tc_prof_local[i_]); } while (0)
// event trace of CTA 0: [role lane 0..11][local tile 0..31][slot 0..3] = clock64()
__device__ long long g_tc_trace[12 * 32 * 4];
#define TC_TRACE(role, tile, slot) do { if (blockIdx.x == 0 && (tile) < 32) g_tc_trace[((role) * 32 + (tile)) * 4 + (slot)] = clock64(); } while (0)
#else
#define TC_TRACE(role, tile, slot) ((void)0)
#define TC_CLK() 0ll
#define TC_ADD(slot, t0) ((void)(t0))
#define TC_DECL() ((void)0)
#define TC_FLUSH() ((void)0)
#endif

// ---------------------------------------------------------------- B-tile pre-pass
// One CTA per (channel m, shapelet block, k-block kb): the stage image the MMA reads, i.e. rows n = 16 u + j
// (shapelet u, shift j), 32 fp32 columns, 128-byte rows, 16-byte chunks XOR-swizzled with (n & 7); the hi image
// (raw fp32, truncated to tf32 by the tensor core) is followed by the lo image when the 3xTF32 split is on.
//   B[n][c] = w'[u][32 kb + c - j]   (0 outside the shapelet; w' = w - mean for pearson)
// The kb == 0 CTAs also write the per-shapelet statistic the epilogue needs (same arithmetic as the FP32 engine):
//   SQL2 ||w||^2, COSINE 1/max(||w||,1e-8), PEARSON ||w-mean||.
__global__ void __launch_bounds__(128) tc_build_b_kernel(const float* __restrict__ W, uint8_t* __restrict__ out,
                                                         float* __restrict__ wstat, int M, int K, int L, int KG,
                                                         int nkb, int NKB, int N, int split, int dist, int stride,
                                                         int NKBr) {
  __shared__ float s_mean[16];
  const int m = blockIdx.x, kblk = blockIdx.y, kb = blockIdx.z;
  const int k0 = kblk * KG;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool centre = dist == IGN_DIST_PEARSON;
  for (int u = warp; u < KG; u += 4) {
    float mean = 0.f, stat = 0.f;
    if (k0 + u < K && (centre || kb == 0)) {
      const float* src = W + ((size_t)(k0 + u) * M + m) * L;
      float s1 = 0.f, s2 = 0.f;
      for (int l = lane; l < L; l += 32) { const float w = __ldg(src + l); s1 += w; s2 = fmaf(w, w, s2); }
#pragma unroll
      for (int o = 16; o; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
      if (centre) {
        mean = s1 / (float)L;
        float c2 = 0.f;
        for (int l = lane; l < L; l += 32) { const float w = __ldg(src + l) - mean; c2 = fmaf(w, w, c2); }
#pragma unroll
        for (int o = 16; o; o >>= 1) c2 += __shfl_xor_sync(0xffffffffu, c2, o);
        stat = sqrtf(c2);
      } else if (dist == IGN_DIST_COSINE) {
        stat = 1.f / fmaxf(sqrtf(s2), 1e-8f);
      } else {
        stat = s2;
      }
    }
    if (lane == 0) {
      s_mean[u] = mean;
      if (kb == 0) wstat[((size_t)m * nkb + kblk) * 8 + u] = stat;
    }
  }
  __syncthreads();
  const int b_bytes = N * 128;
  uint8_t* dst = out + (((size_t)m * nkb + kblk) * NKB + kb) * (size_t)(b_bytes * (split ? 2 : 1));
  const int res = kb / NKBr, kbr = kb - res * NKBr;                  // k-block kb = residue res, its kbr-th block
  for (int task = threadIdx.x; task < N * 8; task += blockDim.x) {
    const int n = task >> 3, c = task & 7;
    const int u = n >> 4, j = n & 15;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int q = kbr * kKBlock + 4 * c + e - j;                     // lag inside the residue
      const int l = q * stride + res;
      v[e] = (k0 + u < K && q >= 0 && l < L) ? __ldg(W + ((size_t)(k0 + u) * M + m) * L + l) - s_mean[u] : 0.f;
    }
    const float4 hi = make_float4(v[0], v[1], v[2], v[3]);
    const uint32_t off = sw128_off(n, c);
    *reinterpret_cast<float4*>(dst + off) = hi;
    if (split) *reinterpret_cast<float4*>(dst + b_bytes + off) = tf32_lo(hi);
  }
}

// ---------------------------------------------------------------- kernel helpers
__device__ __forceinline__ void bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// order-preserving map float -> uint32 (smaller float <=> smaller key), total for all non-NaN values
__device__ __forceinline__ uint32_t ordered_key(float d) {
  const uint32_t u = __float_as_uint(d);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_val(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct TileCoord { int m, kblk, b0, seg; };
// GEN = false is the unit-stride, one-tile-per-sample case (every BASELINE config below T = 2276): stride, tile segment
// and residue loops are compile-time constants there — carrying them as run-time values cost the epilogue-bound
// L = 100 group 8 % (registers are at the 96 cap of a 19-warp CTA).
template <bool GEN>
__device__ __forceinline__ TileCoord tile_coord(const TcGeo& g, int w) {
  TileCoord c;
  const int mk = w / g.tpm;
  const int idx = w - mk * g.tpm;
  if (GEN && g.nseg > 1) { c.b0 = idx / g.nseg; c.seg = idx - c.b0 * g.nseg; }
  else { c.b0 = idx * g.RB; c.seg = 0; }
  c.m = mk / g.nkb;
  c.kblk = mk - c.m * g.nkb;
  return c;
}

// ---------------------------------------------------------------- kernel
// Which (tile, k-block) steps load / wait for / release a B stage.  Resident mode: stage = k-block, loaded on the
// first tile of a (channel, shapelet block) run and released on its last; streaming mode: a plain ring.
struct RunFlags { bool first, last; };
__device__ __forceinline__ RunFlags run_flags(const TcGeo& g, int w, int wbeg, int wend) {
  const int mk = w / g.tpm;
  RunFlags f;
  f.first = (w == wbeg) || ((w - 1) / g.tpm != mk);
  f.last = (w == wend - 1) || ((w + 1) / g.tpm != mk);
  return f;
}

// Merge the four lane-quarter cells of every (sample, shapelet) pair of tile `w` (local index n) and write the pooled
// outputs.  Executed by one warp, after all eight epilogue warps have published the tile (named barrier).
template <bool GEN>
__device__ __forceinline__ void finalize_tile(const TcGeo& g, const TcArgs& a, const unsigned long long* cells,
                                              const float* celld, int n, int w, int lane) {
  bar_sync(2 + (n & (g.ncb - 1)), kEpilogueThreads);
  const TileCoord tc = tile_coord<GEN>(g, w);
  const int npair = g.RB * g.KG;
  const int nb = min(g.RB, g.B - tc.b0);
  const unsigned long long* cbuf = cells + (size_t)(n & (g.ncb - 1)) * 4 * npair;
  const float* dbuf = celld + (size_t)(n & (g.ncb - 1)) * 4 * npair * 16;
  for (int pair = lane; pair < npair; pair += 32) {
    const int rbl = pair / g.KG, kl = pair - rbl * g.KG;
    const int k = tc.kblk * g.KG + kl;
    if (rbl >= nb || k >= g.K) continue;
    unsigned long long pk = cbuf[pair];
    int qw = 0;
#pragma unroll
    for (int q = 1; q < 4; ++q) {        // (ordered min | first window of the row): the earlier row wins ties
      const unsigned long long o = cbuf[q * npair + pair];
      if (o < pk) { pk = o; qw = q; }
    }
    const float dmn = ordered_val((uint32_t)(pk >> 32));
    const float4* dc = reinterpret_cast<const float4*>(dbuf + (size_t)(qw * npair + pair) * 16);
    int jj = 15;
#pragma unroll
    for (int j = 3; j >= 0; --j) {
      const float4 d4 = dc[j];
      if (d4.w == dmn) jj = 4 * j + 3;
      if (d4.z == dmn) jj = 4 * j + 2;
      if (d4.y == dmn) jj = 4 * j + 1;
      if (d4.x == dmn) jj = 4 * j;
    }
    const int imn = (int)(uint32_t)pk + jj;
    const size_t o = ((size_t)(tc.b0 + rbl) * g.K + k) * g.M + tc.m;
    if (GEN && g.nseg > 1) {     // the sample spans several tiles: smallest (distance, index) over them wins, order-independent
      atomicMin(a.packed + o, (pk & 0xffffffff00000000ull) | (unsigned long long)(unsigned)imn);
      continue;
    }
    float pv;
    if (g.pool == IGN_POOL_RBF_MAX) { const float ed = g.eps * dmn; pv = expf(-(ed * ed)); }
    else pv = 1.f / (1.f + expf(-(a.thr[(size_t)k * g.M + tc.m] - dmn)));
    a.p[o] = pv; a.dmin[o] = dmn;
    if (a.argmin) a.argmin[o] = imn;
  }
}

// nseg > 1: decode the merged cells into the pooled outputs
__global__ void __launch_bounds__(256) tc_finish_kernel(const unsigned long long* __restrict__ packed, const float* __restrict__ thr,
                                                       float* __restrict__ p, float* __restrict__ dmin, int* __restrict__ argmin,
                                                       int n, int K, int M, int pool, float eps) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n) return;
  const unsigned long long pk = packed[o];
  const float dmn = ordered_val((uint32_t)(pk >> 32));
  float pv;
  if (pool == IGN_POOL_RBF_MAX) { const float ed = eps * dmn; pv = expf(-(ed * ed)); }
  else { const int km = o % (K * M); pv = 1.f / (1.f + expf(-(thr[km] - dmn))); }
  p[o] = pv; dmin[o] = dmn;
  if (argmin) argmin[o] = (int)(uint32_t)pk;
}

template <int DIST, bool STACK, bool GEN>
__global__ void __launch_bounds__(kThreadsTC, 1) shapelet_fwd_tc_kernel(const TcGeo g, const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int g_s = GEN ? g.s : 1, g_nrb = GEN ? g.nrb : 2, g_NKBr = GEN ? g.NKBr : g.NKB;
  const long long t_entry = TC_CLK();
  TC_DECL();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // this CTA's contiguous tile range (balanced to within one tile)
  const int wbeg = (int)(((long long)g.ntiles * blockIdx.x) / gridDim.x);
  const int wend = (int)(((long long)g.ntiles * (blockIdx.x + 1)) / gridDim.x);

  // ---- shared memory carve-up (stage tiles first: they need 1024-byte alignment) ----
  const int b_bytes = g.N * 128;
  const int stage_bytes = b_bytes * (g.split ? 2 : 1);
  uint8_t* stage0 = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ptr = stage0 + (size_t)g.nbs * stage_bytes;
  float* xbuf = reinterpret_cast<float*>(ptr);                      // [nrb][RB][XR] series-row units (ring)
  ptr += (size_t)g_nrb * g.RB * g.XR * sizeof(float);
  float* celld = reinterpret_cast<float*>(ptr);                     // [ncb][4][RB][KG][16] winner rows' distances
  ptr += (size_t)g.ncb * 4 * g.RB * g.KG * 16 * sizeof(float);
  unsigned long long* cells = reinterpret_cast<unsigned long long*>(ptr);   // [ncb][4][RB][KG] ordered min | first window of the row
  ptr += (size_t)g.ncb * 4 * g.RB * g.KG * sizeof(unsigned long long);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ptr);
  uint64_t* fullA = bars; uint64_t* emptyA = fullA + kAStages;
  uint64_t* fullB = emptyA + kAStages; uint64_t* emptyB = fullB + kMaxBStages;
  uint64_t* tfull = emptyB + kMaxBStages; uint64_t* tempty = tfull + kMaxAcc;
  uint64_t* rowfull = tempty + kMaxAcc; uint64_t* rowempty = rowfull + kMaxRowBufs;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rowempty + kMaxRowBufs);

  const uint32_t tmem_cols = 512;                                   // nacc accumulators in columns [0,256) + kAStages A stages

  if (threadIdx.x == 0) {
    for (int s = 0; s < kAStages; ++s) { mbar_init(&fullA[s], 4); mbar_init(&emptyA[s], 1); }   // 4 producer warps
    for (int s = 0; s < kMaxBStages; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    for (int i = 0; i < kMaxAcc; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEpiWarps); }
    for (int i = 0; i < kMaxRowBufs; ++i) { mbar_init(&rowfull[i], 1); mbar_init(&rowempty[i], kProdWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) TC_ADD(10, t_entry);       // barrier init + TMEM alloc

  if (warp < kProdWarps) {
    // =================================================================== PRODUCERS
    const int p = threadIdx.x;                                      // 0..255
    (void)p;
    const int row = (warp & 3) * 32 + lane;                         // accumulator row = TMEM lane
    const int grp = warp >> 2;                                      // takes the stages with (it & 1) == grp
    // this thread's row: samples 16 i + 32 kb + (0..31) of series bl
    const int a_bl = row / g.RI, a_i = row - a_bl * g.RI;
    int a_off[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int idx = a_i * kShifts + 4 * c;
      a_off[c] = a_bl < g.RB ? a_bl * g.XR + idx + 4 * (idx >> 5) : -1;
    }
    const uint32_t a_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)g.acol0;
    uint32_t it = 0;                                                 // global stage counter
    int sidx = 0; uint32_t sph = 0;                                  // A ring slot / phase of stage `it`
    int rbuf = 0; uint32_t rph = 0;                                  // row-unit ring slot / phase
    for (int w = wbeg; w < wend; ++w) {
      const TileCoord tc = tile_coord<GEN>(g, w);
      const int nb = min(g.RB, g.B - tc.b0);
      const bool a_live = a_off[0] >= 0 && a_bl < nb && (!GEN || tc.seg * kRows + a_i < g.RItot);
      const int seg_off = GEN ? tc.seg * (kRows * kShifts / 32) * 36 : 0;      // 2048 samples further along the bank-skewed row
      for (int res = 0; res < g_s; ++res) {                          // one row unit per residue (one per tile at unit stride)
      const int buf = rbuf;
      long long tr0 = TC_CLK();
      mbar_wait(&rowfull[buf], rph);                                 // this unit's series rows are in shared memory
      if (++rbuf == g_nrb) { rbuf = 0; rph ^= 1; }
      if (lane == 0 && (warp & 3) == 0 && res == 0) TC_TRACE(9 + grp, w - wbeg, 0);
      if (p == 0) TC_ADD(3, tr0);
      const float* xb = xbuf + (size_t)buf * g.RB * g.XR;
      for (int kb = 0; kb < g_NKBr; ++kb, ++it) {
        const int s = sidx;
        const uint32_t ph = sph;
        if (++sidx == g.nast) { sidx = 0; sph ^= 1; }
        if ((int)(it & 1) != grp) continue;
        // A stage -> tensor memory: 32 columns of raw fp32 (the MMA truncates to tf32 = hi) and 32 of lo.
        // The shared-memory reads do not depend on the stage being free: issue them before the wait.
        float4 v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a_live) v[c] = *reinterpret_cast<const float4*>(xb + a_off[c] + seg_off + kb * 36);   // 32 samples + 4 pad
        }
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 l4 = tf32_lo(v[c]);
          hi[4 * c] = __float_as_uint(v[c].x); hi[4 * c + 1] = __float_as_uint(v[c].y);
          hi[4 * c + 2] = __float_as_uint(v[c].z); hi[4 * c + 3] = __float_as_uint(v[c].w);
          lo[4 * c] = __float_as_uint(l4.x); lo[4 * c + 1] = __float_as_uint(l4.y);
          lo[4 * c + 2] = __float_as_uint(l4.z); lo[4 * c + 3] = __float_as_uint(l4.w);
        }
        long long tp0 = TC_CLK();
        mbar_wait(&emptyA[s], ph ^ 1);                               // passes immediately on the first lap
        tc_fence_after();
        if (p == 0) TC_ADD(0, tp0);
        tp0 = TC_CLK();
        tmem_st32(a_lane + s * kAStageCols, hi);
        if (g.split) tmem_st32(a_lane + s * kAStageCols + 32, lo);
        tmem_st_wait();                                              // A columns written
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&fullA[s]);
        if (lane == 0 && (warp & 3) == 0) { if (kb < 2) TC_TRACE(9 + grp, w - wbeg, 1); TC_TRACE(9 + grp, w - wbeg, 2); }
        if (p == 0) TC_ADD(1, tp0);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&rowempty[buf]);                    // this warp no longer reads the unit's rows
      }   // residues
    }
  } else if (warp < kMmaWarp) {
    // =================================================================== EPILOGUE
    const int ew = warp - kProdWarps;                               // 0..7
    const int quarter = ew & 3;                                     // TMEM lane quarter
    // The two warps of a quarter take alternate shapelets (kl = half, half+2, ..) and swap parity every tile, so an
    // odd shapelet count (3 + 2 for K = 5) balances out over two tiles — the accumulator ring absorbs the rest.
    const int e = quarter * 32 + lane;                              // accumulator row
    const int et = threadIdx.x - kProducerThreads;                  // 0..255
    const int bl = e / g.RI, i = e - bl * g.RI;
    const float invL = 1.f / (float)g.L;
    const int npair = g.RB * g.KG;
    const int cells_per_buf = 4 * npair;
    for (int q = et; q < g.ncb * cells_per_buf; q += kEpilogueThreads) cells[q] = kSlotEmpty;
    bar_sync(2, kEpilogueThreads);
    // the samples this warp's 32 rows belong to (row -> sample is the same for every tile, so a cell
    // (quarter, sample, shapelet) is either rewritten by exactly one warp on every tile or stays empty for ever)
    const int bl_lo = __shfl_sync(0xffffffffu, bl, 0);
    const int bl_hi = min(__shfl_sync(0xffffffffu, bl, 31), g.RB - 1);
    int n = 0;                                                      // local tile counter
    for (int w = wbeg; w < wend; ++w, ++n) {
      const int acc = n % g.nacc;
      const int half = (ew >> 2) ^ (n & 1);
      const TileCoord tc = tile_coord<GEN>(g, w);
      const int m = tc.m, k0 = tc.kblk * g.KG;
      const int nb = min(g.RB, g.B - tc.b0);
      const int gi = GEN ? tc.seg * kRows + i : i;                  // this row's window group inside its sample
      const bool row_live = bl < nb && (!GEN || gi < g.RItot);      // (bl < RB is implied: nb <= RB)
      const int t0 = gi * kShifts;
      const int nst = min(4, max(0, (g.Ts - t0) / 4));              // float4 stores per distance row segment (Ts % 4 == 0)
      const int b = tc.b0 + bl;
      unsigned long long* cbuf = cells + (size_t)(n & (g.ncb - 1)) * cells_per_buf;
      float* dbuf = celld + (size_t)(n & (g.ncb - 1)) * cells_per_buf * 16;
      // this row's 16 window norm terms (fp32, from the window-statistics pass) and the shapelet statistics:
      // loads issued before the accumulator wait so their latency hides under the MMAs of this tile
      float4 xs4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) xs4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_live && t0 < a.SP) {
        const float4* sp = reinterpret_cast<const float4*>(a.st0 + ((size_t)b * g.M + m) * a.SP + t0);
#pragma unroll
        for (int j = 0; j < 4; ++j) xs4[j] = __ldg(sp + j);
      }
      const float* wsp = a.wstat + ((size_t)m * g.nkb + tc.kblk) * 8;
      float wst_next = half < g.KG ? __ldg(wsp + half) : 0.f;
      long long te0 = TC_CLK();
      mbar_wait(&tfull[acc], (n / g.nacc) & 1);
      if (lane == 0) TC_TRACE(ew, n, 0);
      if (et == 0) TC_ADD(7, te0);
      te0 = TC_CLK();
      tc_fence_after();
      const float xst[16] = {xs4[0].x, xs4[0].y, xs4[0].z, xs4[0].w, xs4[1].x, xs4[1].y, xs4[1].z, xs4[1].w,
                             xs4[2].x, xs4[2].y, xs4[2].z, xs4[2].w, xs4[3].x, xs4[3].y, xs4[3].z, xs4[3].w};
      const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * g.accp);
      float* drow = (a.dstore && row_live) ? a.dstore + (((size_t)b * g.M + m) * g.K + k0) * g.Ts + t0 : nullptr;
      uint32_t v[16], v2[16];
      if (half < g.KG) { tmem_ld16(trow + half * 16, v); if (STACK) tmem_ld16(trow + g.N + half * 16, v2); }
#pragma unroll 1
      for (int kl = half; kl < g.KG; kl += 2) {
        tmem_ld_wait();
        float raw[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) raw[j] = __uint_as_float(v[j]);
        if (STACK) {
#pragma unroll
          for (int j = 0; j < 16; ++j) raw[j] += __uint_as_float(v2[j]);      // + hi*lo term of the stacked form
        }
        if (kl + 2 < g.KG) {                                         // next shapelet's columns load under this one's math
          tmem_ld16(trow + (kl + 2) * 16, v);
          if (STACK) tmem_ld16(trow + g.N + (kl + 2) * 16, v2);
        }
        const int k = k0 + kl;
        const float wst = wst_next;
        if (kl + 2 < g.KG) wst_next = __ldg(wsp + kl + 2);
        float dv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (DIST == IGN_DIST_SQL2) dv[j] = fmaxf((xst[j] + wst - 2.f * raw[j]) * invL, 0.f);
          else if (DIST == IGN_DIST_COSINE) dv[j] = 1.f - raw[j] * xst[j] * wst;
          else dv[j] = 1.f - __fdividef(raw[j], xst[j] * wst + 1e-8f);   // norms hoisted; no IEEE slow paths on zeros
        }
        const bool kvalid = k < g.K;
        // pad windows (t >= T') got NaN / +inf from the statistics pass: fminf skips them, no per-window masks
        float best = fminf(fminf(fminf(dv[0], dv[1]), fminf(dv[2], dv[3])), fminf(fminf(dv[4], dv[5]), fminf(dv[6], dv[7])));
        best = fminf(best, fminf(fminf(fminf(dv[8], dv[9]), fminf(dv[10], dv[11])), fminf(fminf(dv[12], dv[13]), fminf(dv[14], dv[15]))));
        if (drow && kvalid) {
          float4* dg = reinterpret_cast<float4*>(drow + (size_t)kl * g.Ts);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < nst) dg[j] = make_float4(dv[4 * j], dv[4 * j + 1], dv[4 * j + 2], dv[4 * j + 3]);
        }
        // arg-min of each (sample, shapelet) pair over this warp's rows: redux.min on the ordered key; the first
        // row holding the minimum publishes (key | its first window) and its 16 distances to the warp's own cell —
        // plain stores, no atomics, no intra-warp round trip; the window inside the row is found by the finaliser
        const uint32_t key = (row_live && kvalid) ? ordered_key(best) : 0xffffffffu;
        for (int sb = bl_lo; sb <= bl_hi; ++sb) {
          const bool mine = bl == sb;
          const uint32_t wmin = __reduce_min_sync(0xffffffffu, mine ? key : 0xffffffffu);
          const unsigned hit = __ballot_sync(0xffffffffu, mine && key == wmin);
          if (lane == __ffs(hit) - 1) {
            const int cell = (quarter * g.RB + sb) * g.KG + kl;
            float4* dc = reinterpret_cast<float4*>(dbuf + (size_t)cell * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) dc[j] = make_float4(dv[4 * j], dv[4 * j + 1], dv[4 * j + 2], dv[4 * j + 3]);
            cbuf[cell] = ((unsigned long long)wmin << 32) | (unsigned)t0;
          }
        }
      }
      __threadfence_block();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);                      // this warp's share of the accumulator is drained
      if (lane == 0) TC_TRACE(ew, n, 1);
      if (et == 0) TC_ADD(8, te0);
      // One rotating warp (tile & 7) finalises each tile, ONE TILE LATE: at the end of tile n the finaliser of tile
      // n-1 syncs on that tile's barrier, which the other seven warps only arrived at — so nobody ever waits for a
      // slower warp of the same tile, and there is no serial chain through the finalisers.
      // Barrier ids and cell buffers rotate over ncb >= nacc + 2 tiles: a warp can reach the barrier (or rewrite the
      // cells) of tile n-1+ncb only after the MMAs of that tile, which need every warp's tempty arrival of tile
      // n-1+ncb-nacc >= n+1, which the finaliser of tile n-1 issues after it has left the barrier of tile n-1.
      if ((n & (kEpiWarps - 1)) != ew) bar_arrive(2 + (n & (g.ncb - 1)), kEpilogueThreads);
      if (n > 0 && ((n - 1) & (kEpiWarps - 1)) == ew) {
        te0 = TC_CLK();
        finalize_tile<GEN>(g, a, cells, celld, n - 1, w - 1, lane);
        if (lane == 0) TC_TRACE(ew, n, 2);
        if (lane == 0) TC_ADD(9, te0);
      }
    }
    if (n > 0 && ((n - 1) & (kEpiWarps - 1)) == ew) finalize_tile<GEN>(g, a, cells, celld, n - 1, wend - 1, lane);
    if (lane == 0) TC_FLUSH();
  } else if (warp == kMmaWarp) {
    // =================================================================== MMA ISSUER (one elected thread)
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_tf32(kRows, g.N);
      const uint32_t bdesc0 = (smem_u32(stage0) & 0x3FFFFu) >> 4;    // descriptor low word of B stage 0, hi image
      const uint32_t bstage16 = (uint32_t)stage_bytes >> 4, bimg16 = (uint32_t)b_bytes >> 4;
      const uint32_t a_tmem0 = tmem_base + (uint32_t)g.acol0;
      const uint32_t idesc2 = umma_idesc_tf32(kRows, 2 * g.N);        // stacked [B_hi | B_lo] form
      // ring positions and phases are carried incrementally: no divisions in the issue loop
      int sa = 0, sbr = 0, acc = 0;
      uint32_t pha = 0, phb = 0, phacc = 0;
      int n = 0, run = -1;
      for (int w = wbeg; w < wend; ++w, ++n) {
        const RunFlags rf = run_flags(g, w, wbeg, wend);
        if (rf.first) ++run;
        long long tm0 = TC_CLK();
        mbar_wait(&tempty[acc], phacc ^ 1);                          // passes immediately on the first lap
        TC_TRACE(8, n, 0);
        TC_ADD(4, tm0);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * g.accp);
        for (int kb = 0; kb < g.NKB; ++kb) {
          int sbi;
          tm0 = TC_CLK();
          if (g.resident) {
            sbi = kb;
            if (rf.first) mbar_wait(&fullB[sbi], run & 1);
          } else {
            sbi = sbr;
            mbar_wait(&fullB[sbi], phb);
          }
          TC_ADD(2, tm0);
          tm0 = TC_CLK();
          mbar_wait(&fullA[sa], pha);
          if (kb == 0) TC_TRACE(8, n, 1);
          if (kb == g.NKB - 1) TC_TRACE(8, n, 2);
          TC_ADD(5, tm0);
          tm0 = TC_CLK();
          tc_fence_after();
          // descriptor low words: (shared address >> 4); all stage images sit below 256 KB, so no 14-bit wrap
          const uint32_t bd_hi = bdesc0 + (uint32_t)sbi * bstage16;
          const uint32_t bd_lo = bd_hi + bimg16;
          const uint32_t a_hi = a_tmem0 + sa * kAStageCols, a_lo = a_hi + 32;
          if (STACK) {
            // the lo image follows the hi image in the stage (80-row multiples of the 1024-byte swizzle atom), so a
            // 2N-row descriptor over the hi image IS [B_hi | B_lo]: columns [0,N) collect hi*hi + lo*hi, [N,2N) hi*lo
#pragma unroll
            for (int k8 = 0; k8 < kKBlock / 8; ++k8) {
              umma_tf32_ts_lo(d_tmem, a_hi + k8 * 8, bd_hi + 2 * k8, idesc2, (kb | k8) != 0);
              umma_tf32_ts_lo(d_tmem, a_lo + k8 * 8, bd_hi + 2 * k8, idesc, 1);
            }
          } else if (g.split) {
#pragma unroll
            for (int k8 = 0; k8 < kKBlock / 8; ++k8) {               // small terms first; 8 tf32 = 32 bytes = 2 units
              umma_tf32_ts_lo(d_tmem, a_lo + k8 * 8, bd_hi + 2 * k8, idesc, (kb | k8) != 0);
              umma_tf32_ts_lo(d_tmem, a_hi + k8 * 8, bd_lo + 2 * k8, idesc, 1);
              umma_tf32_ts_lo(d_tmem, a_hi + k8 * 8, bd_hi + 2 * k8, idesc, 1);
            }
          } else {
#pragma unroll
            for (int k8 = 0; k8 < kKBlock / 8; ++k8)
              umma_tf32_ts_lo(d_tmem, a_hi + k8 * 8, bd_hi + 2 * k8, idesc, (kb | k8) != 0);
          }
          umma_commit(&emptyA[sa]);                                  // A stage reusable once these MMAs retire
          if (!g.resident || rf.last) umma_commit(&emptyB[sbi]);
          TC_ADD(6, tm0);
          if (++sa == g.nast) { sa = 0; pha ^= 1; }
          if (++sbr == g.nbs) { sbr = 0; phb ^= 1; }
        }
        umma_commit(&tfull[acc]);                                    // accumulator complete
        TC_TRACE(8, n, 3);
        if (++acc == g.nacc) { acc = 0; phacc ^= 1; }
      }
      TC_ADD(14, t_entry);
      TC_FLUSH();
    }
    __syncwarp();
  } else if (warp == kTmaWarp) {
    // =================================================================== B-TILE LOADER (one elected thread)
    if (elect_one()) {
      uint32_t itb = 0;
      int run = -1;
      for (int w = wbeg; w < wend; ++w) {
        const RunFlags rf = run_flags(g, w, wbeg, wend);
        if (rf.first) ++run;
        if (g.resident && !rf.first) continue;
        const TileCoord tc = tile_coord<GEN>(g, w);
        const uint8_t* bsrc = a.btiles + ((size_t)tc.m * g.nkb + tc.kblk) * g.NKB * (size_t)stage_bytes;
        for (int kb = 0; kb < g.NKB; ++kb, ++itb) {
          int sbi; uint32_t par;
          if (g.resident) { sbi = kb; par = (uint32_t)(run & 1) ^ 1u; }
          else { sbi = itb % g.nbs; par = ((itb / g.nbs) & 1) ^ 1u; }
          mbar_wait(&emptyB[sbi], par);                              // the MMAs that read this stage have retired
          mbar_arrive_expect_tx(&fullB[sbi], (uint32_t)stage_bytes);
          tma_bulk_g2s(stage0 + (size_t)sbi * stage_bytes, bsrc + (size_t)kb * stage_bytes, (uint32_t)stage_bytes, &fullB[sbi]);
        }
      }
    }
    __syncwarp();
  }

  if (warp == kRowWarp) {
    // =================================================================== SERIES-ROW LOADER (one warp)
    // Rows are stored bank-skewed (4 floats of padding after every 32) so that the 128 row owners, whose segments
    // start 16 samples apart, read with conflict-free LDS.128.  Ring of nrb (tile, residue) units; the unit issued two
    // steps ago is published while the newer ones are still in flight (cp.async groups).
    for (int i = lane; i < g_nrb * g.RB * g.XR; i += 32) xbuf[i] = 0.f;
    __syncwarp();
    const int chunks = g.Tp / 4;
    const int lag = g_s == 1 ? 0 : 2;                                // units kept in flight behind the one being issued
    int u = 0, done = 0;                                             // units issued / published
    for (int w = wbeg; w < wend; ++w) {
      const TileCoord tc = tile_coord<GEN>(g, w);
      for (int res = 0; res < g_s; ++res, ++u) {
        const int buf = u % g_nrb;
        if (u >= g_nrb) mbar_wait(&rowempty[buf], ((u / g_nrb) - 1) & 1);   // every producer warp is done with unit u - nrb
        for (int bl = 0; bl < g.RB; ++bl) {
          if (tc.b0 + bl >= g.B) break;
          float* dst = xbuf + ((size_t)buf * g.RB + bl) * g.XR;
          const float* src = a.xn + ((size_t)(tc.b0 + bl) * g.M + tc.m) * g.Tp;
          if (g_s == 1) {
            for (int c = lane; c < chunks; c += 32) cp_async16(dst + c * 4 + 4 * (c >> 3), src + c * 4);
          } else {
            const int nq = (g.T - res + g.s - 1) / g.s;             // x_r[j] = x[j s + res]
            for (int j = lane; j < nq; j += 32)
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst + j + 4 * (j >> 5))), "l"(src + (size_t)j * g.s + res) : "memory");
            if (lane == 0) dst[nq + 4 * (nq >> 5)] = 0.f;          // the sample a one-longer residue row left behind
          }
        }
        cp_async_commit();
        if (u - done >= lag) {
          if (lag == 0) cp_async_wait_all();
          else asm volatile("cp.async.wait_group 2;" ::: "memory");
          __threadfence_block();
          __syncwarp();
          if (lane == 0) mbar_arrive(&rowfull[done % g_nrb]);
          ++done;
        }
      }
    }
    cp_async_wait_all();
    __threadfence_block();
    __syncwarp();
    for (; done < u; ++done)
      if (lane == 0) mbar_arrive(&rowfull[done % g_nrb]);
  }

  if (threadIdx.x == 0) TC_ADD(12, t_entry);         // producer thread 0 finished its loop
  if (threadIdx.x == kProducerThreads) TC_ADD(13, t_entry);   // epilogue thread 0 finished
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, tmem_cols);
  if (threadIdx.x == 0) { TC_ADD(15, t_entry); TC_FLUSH(); }  // CTA lifetime
}

size_t tc_smem_fixed(const TcGeo& g) {   // everything except the B ring
  return (size_t)g.nrb * g.RB * g.XR * 4 + (size_t)g.ncb * 4 * g.RB * g.KG * (8 + 64) +
         (2 * kAStages + 2 * kMaxBStages + 2 * kMaxAcc + 2 * kMaxRowBufs) * 8 + 16 + 1024;
}
size_t tc_stage_bytes(const TcGeo& g) { return (size_t)(g.N * 128) * (g.split ? 2 : 1); }

// B ring depth: resident (every k-block of a run stays in shared memory) when it fits, else as deep as fits
bool tc_plan_ring(TcGeo& g) {
  const size_t cap = (size_t)max_optin_smem();
  const size_t fixed = tc_smem_fixed(g), stage = tc_stage_bytes(g);
  if (fixed + 2 * stage > cap) return false;
  const int fit = (int)min((size_t)kMaxBStages, (cap - fixed) / stage);
  g.resident = g.NKB <= fit ? 1 : 0;
  g.nbs = g.resident ? g.NKB : fit;
  return true;
}

}  // namespace

int tc_trace_read(long long* host, int n) {
#ifdef IGN_TC_PROFILE
  IGN_CUDA(cudaMemcpyFromSymbol(host, g_tc_trace, sizeof(long long) * (size_t)min(n, 12 * 32 * 4)));
  return IGN_OK;
#else
  (void)host; (void)n;
  set_error("library built without -DIGN_TC_PROFILE");
  return IGN_ERR_UNSUPPORTED;
#endif
}

int tc_profile_read(unsigned long long* host16, int reset) {
#ifdef IGN_TC_PROFILE
  IGN_CUDA(cudaMemcpyFromSymbol(host16, g_tc_prof, sizeof(unsigned long long) * 16));
  if (reset) { unsigned long long z[16] = {0}; IGN_CUDA(cudaMemcpyToSymbol(g_tc_prof, z, sizeof(z))); }
  return IGN_OK;
#else
  (void)host16; (void)reset;
  set_error("library built without -DIGN_TC_PROFILE");
  return IGN_ERR_UNSUPPORTED;
#endif
}

static void tc_geo(const ign_shapelet_desc& d, TcGeo& g) {
  g.B = d.B; g.M = d.M; g.T = d.T; g.Tp = d.Tp; g.K = d.K; g.L = d.L;
  g.s = d.stride; g.nrb = g.s == 1 ? 2 : kMaxRowBufs;
  g.Tw = num_windows(d.T, d.L, d.stride); g.Ts = round_up(g.Tw, 4);
  g.RItot = ceil_div(g.Tw, kShifts);
  g.nseg = ceil_div(g.RItot, kRows);
  g.RI = min(g.RItot, kRows);
  g.RB = g.nseg > 1 ? 1 : max(1, min(min(kRows / g.RI, d.B), kMaxRB));
  g.nkb = ceil_div(d.K, 8); g.KG = ceil_div(d.K, g.nkb); g.N = 16 * g.KG;   // N <= 128 = kAccCols
  g.NKBr = ceil_div(ceil_div(d.L, g.s) + kShifts - 1, kKBlock);
  g.NKB = g.s * g.NKBr;
  {  // bank-skewed series rows (one residue per row unit): 36 floats per 32 samples
    const int span = max(g.s == 1 ? d.Tp : ceil_div(d.T, g.s), (g.RItot - 1) * kShifts + g.NKBr * kKBlock) + 32;
    g.XR = round_up(span + 4 * (span / 32) + 8, 4);
  }
  g.tpm = g.nseg > 1 ? d.B * g.nseg : ceil_div(d.B, g.RB);
  g.ntiles = d.M * g.nkb * g.tpm;
  g.dist = d.dist; g.pool = d.pool; g.eps = d.eps;
  g.split = d.precision == IGN_PREC_3XTF32 ? 1 : 0;
  // Short tiles (few k-blocks) finish their MMAs faster than the epilogue drains an accumulator: keep three
  // accumulators in flight when they fit the 256 TMEM columns reserved for them.
  g.nacc = (g.NKB <= 7 && g.N <= 80) ? 3 : 2;
  g.accp = g.nacc == 3 ? 80 : kAccCols;
  g.ncb = g.nacc == 3 ? 8 : 4;
  g.stack = 0; g.nast = kAStages; g.acol0 = 256;
  // Long tiles are paced by the MMA-issuing thread: with [B_hi | B_lo] stacked along N the 3xTF32 product takes two
  // MMAs per k-step (2N and N columns) instead of three.  Two 2N-column accumulators + three A stages fill TMEM.
  if (g.split && g.nacc == 2 && 4 * g.N + 3 * kAStageCols <= 512) {
    g.stack = 1; g.accp = 2 * g.N; g.acol0 = 4 * g.N; g.nast = min(kAStages, (512 - g.acol0) / kAStageCols);
  }
  g.nbs = 2; g.resident = 0;
}

static size_t tc_btile_bytes(const ign_shapelet_desc& d, const TcGeo& g) {
  return (size_t)d.M * g.nkb * g.NKB * (size_t)(g.N * 128) * (g.split ? 2 : 1);   // a multiple of 128
}

static size_t tc_wstat_bytes(const ign_shapelet_desc& d, const TcGeo& g) {
  return ((size_t)d.M * g.nkb * 8 * sizeof(float) + 127) / 128 * 128;
}

// bytes of workspace the tcgen05 forward needs: pre-swizzled B tiles, the per-shapelet statistics and (series with more
// than 2048 windows only) the cross-tile arg-min cells
size_t shapelet_fwd_tc_workspace(const ign_shapelet_desc& d) {
  TcGeo g;
  tc_geo(d, g);
  return tc_btile_bytes(d, g) + tc_wstat_bytes(d, g) + (g.nseg > 1 ? (size_t)d.B * d.K * d.M * sizeof(unsigned long long) : 0);
}

bool shapelet_fwd_tc_supported(const ign_shapelet_desc& d) {
  if (d.dist == IGN_DIST_L1) return false;
  if (d.precision != IGN_PREC_3XTF32 && d.precision != IGN_PREC_TF32) return false;
  const int Tw = num_windows(d.T, d.L, d.stride);
  if (Tw <= 0 || ceil_div(Tw, kShifts) > 16 * kRows) return false;   // at most 16 tiles (32768 windows) per sample
  TcGeo g;
  tc_geo(d, g);
  return tc_plan_ring(g);
}

int launch_shapelet_fwd_tc(const ign_shapelet_desc& d, const float* xn, const float* st0,
                           const float* W, const float* thr, float* p, float* dmin, int* argmin, float* dstore,
                           void* ws, size_t ws_bytes, cudaStream_t st) {
  TcGeo g;
  tc_geo(d, g);
  const size_t need = shapelet_fwd_tc_workspace(d);
  if (!ws || ws_bytes < need) { set_error("shapelet_forward(tcgen05): workspace %zu < %zu bytes (ign_shapelet_forward_workspace)", ws_bytes, need); return IGN_ERR_INVALID; }
  if (((uintptr_t)ws & 127) != 0) { set_error("shapelet_forward(tcgen05): workspace must be 128-byte aligned"); return IGN_ERR_INVALID; }
  if (!tc_plan_ring(g)) { set_error("shapelet_forward(tcgen05): L=%d K=%d does not fit shared memory", d.L, d.K); return IGN_ERR_UNSUPPORTED; }
  uint8_t* btiles = reinterpret_cast<uint8_t*>(ws);
  float* wstat = reinterpret_cast<float*>(btiles + tc_btile_bytes(d, g));
  unsigned long long* packed = g.nseg > 1 ? reinterpret_cast<unsigned long long*>(btiles + tc_btile_bytes(d, g) + tc_wstat_bytes(d, g)) : nullptr;
  if (packed) IGN_CUDA(cudaMemsetAsync(packed, 0xff, (size_t)d.B * d.K * d.M * sizeof(unsigned long long), st));
  // 1. shifted-shapelet operand + shapelet statistics, once per launch, already in the swizzled tile image
  tc_build_b_kernel<<<dim3(d.M, g.nkb, g.NKB), 128, 0, st>>>(W, btiles, wstat, d.M, d.K, d.L, g.KG, g.nkb, g.NKB, g.N,
                                                            g.split, d.dist, g.s, g.NKBr);
  IGN_CUDA(cudaGetLastError());
  // 2. main kernel: persistent, one CTA per SM (the kernel owns all 512 TMEM columns of its SM)
  const size_t smem = max(tc_smem_fixed(g) + g.nbs * tc_stage_bytes(g), (size_t)118 * 1024);   // > half the SM: one CTA per SM
  TcArgs a{xn, st0, thr, p, dmin, argmin, dstore, stats_pitch(d.T, d.L, d.stride), packed, btiles, wstat};
  const int grid = min(sm_count(), g.ntiles);
  const bool gen = g.s != 1 || g.nseg != 1;
#define IGN_TC_PICK(GEN)                                                                                   \
  (g.stack ? (d.dist == IGN_DIST_SQL2 ? shapelet_fwd_tc_kernel<IGN_DIST_SQL2, true, GEN>                    \
              : d.dist == IGN_DIST_COSINE ? shapelet_fwd_tc_kernel<IGN_DIST_COSINE, true, GEN>              \
                                          : shapelet_fwd_tc_kernel<IGN_DIST_PEARSON, true, GEN>)            \
           : (d.dist == IGN_DIST_SQL2 ? shapelet_fwd_tc_kernel<IGN_DIST_SQL2, false, GEN>                   \
              : d.dist == IGN_DIST_COSINE ? shapelet_fwd_tc_kernel<IGN_DIST_COSINE, false, GEN>             \
                                          : shapelet_fwd_tc_kernel<IGN_DIST_PEARSON, false, GEN>))
  auto kern = gen ? IGN_TC_PICK(true) : IGN_TC_PICK(false);
#undef IGN_TC_PICK
  IGN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kThreadsTC, smem, st>>>(g, a);
  IGN_CUDA(cudaGetLastError());
  if (packed) {
    const int n = d.B * d.K * d.M;
    tc_finish_kernel<<<ceil_div(n, 256), 256, 0, st>>>(packed, thr, p, dmin, argmin, n, d.K, d.M, d.pool, d.eps);
    IGN_CUDA(cudaGetLastError());
  }
  return IGN_OK;
}

}  // namespace ign
