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
//    (tcgen05.mma [d], [a_tmem], b_desc).  Shared-memory bandwidth — the limit of an smem-A design at N = 80,
//    measured at ~1220 cycles per k-block against 480 MMA cycles — is left to the B tiles alone.
//
// Precision: tcgen05 has no fp32 kind.  kind::tf32 reads fp32 bits and drops the low 13 mantissa bits, so
// the "hi" operand is the raw fp32 value and lo = x - trunc_tf32(x) is formed by the producers.
//   IGN_PREC_3XTF32: hi*hi + hi*lo + lo*hi  (3 MMAs per k-step, fp32-equivalent: error ~2^-21)
//   IGN_PREC_TF32  : hi*hi only             (own, looser tolerance)
//
//  * the shifted-shapelet operand depends only on (channel, k-block): a small pre-pass writes it to a workspace
//    already in the 128B-swizzled K-major tile image (hi | lo), and the main kernel streams one tile per
//    stage with a single 1-D bulk TMA copy (cp.async.bulk + mbarrier complete_tx) — no producer instructions.
// TMEM map (512 columns): accumulators at 0 and 128 (double buffered), A stages at 256 + 64 s (hi 32 | lo 32).
// Roles (one CTA per SM, 288 threads):
//   warps 0-3  producers: cp.async the next series rows, tcgen05.st the A stage, arrive on full[stage];
//              thread 0 also arms full[stage] with the byte count and issues the bulk TMA copy of the B tile
//   warps 4-7  epilogue : tcgen05.ld the accumulator (TMEM lanes 32*(w%4)..), cross -> distance with the
//                         prefix-sum window norms, coalesced store of d, per-row arg-min candidates
//   warp  8    one elected thread issues tcgen05.mma (M=128, N=16*KG, K=8) and tcgen05.commit
// Pipelines: smem full/empty per stage (producers <-> MMA), TMEM full/empty per accumulator (MMA <-> epilogue).
#include "ign_common.cuh"

#include <math.h>

namespace ign {
namespace {

constexpr int kShifts = 16;            // P
constexpr int kRows = 128;             // UMMA M
constexpr int kKBlock = 32;            // fp32 elements per 128-byte swizzle row
constexpr int kProducerThreads = 128;
constexpr int kEpilogueThreads = 128;
constexpr int kThreadsTC = kProducerThreads + kEpilogueThreads + 32;
constexpr int kWshPad = 16;            // zero floats left of each shifted shapelet row
constexpr int kMaxStages = 4;
constexpr int kAccCols = 128;          // TMEM columns reserved per accumulator (N <= 128)
constexpr int kAStageCols = 64;        // TMEM columns per A stage: 32 hi + 32 lo

struct TcGeo {
  int B, M, T, Tp, K, L;
  int Tw, Ts;          // windows, dstore pitch
  int RI, RB;          // window groups (of 16) per sample, samples per 128-row tile
  int KG, nkb;         // shapelets per N tile, number of shapelet blocks
  int N;               // 16*KG
  int NKB;             // 32-wide k-blocks: ceil((L+15)/32)
  int XR;              // floats per series row in smem
  int nstage;
  int bpc;             // samples per CTA chunk
  int dist, pool, split;
  float eps;
};

struct TcArgs {
  const float* xn; const float* st0; const float* W; const float* thr;   // st0: window statistics [B,M,SP]
  float* p; float* dmin; int* argmin; float* dstore;
  int SP;
  const uint8_t* btiles;   // [M][nkb][NKB][stage_bytes] pre-swizzled B stage images
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  unsigned spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 24)) __trap();   // a pipeline bug must fail fast, never hang the GPU
  }
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk TMA: global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::tf32, issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// this thread's 32 consecutive 32-bit columns of its TMEM lane
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One lane of a fully converged warp.  Code under this predicate is provably single-threaded, so ptxas emits
// the uniform-datapath instructions (UTCHMMA, UTCBAR, UBLKCP) straight-line; under `lane == 0` it wraps every
// one of them in an ELECT / BRA.U.ANY loop, measured at 64 cycles per MMA issue against a 40-cycle dispatch
// floor (profiles/r1d_ubench_umma_dispatch.txt).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// K-major, 128-byte-swizzled shared-memory matrix descriptor (sm_100 format): start>>4 | SBO(1024 B)>>4 at
// bit 32 | version 1 at bit 46 | layout SWIZZLE_128B (2) at bit 61.  LBO is unused for swizzled K-major.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32 instruction descriptor: D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), both K-major, N>>3 at 17, M>>4 at 24
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of 16-byte chunk c (0..7) of row r inside a 128B-swizzled K-major tile (1024-byte aligned base)
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ float4 tf32_lo(float4 v) {   // v - trunc_tf32(v), exact in fp32
  float4 o;
  o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
  o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
  o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
  o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
  return o;
}

// Optional role-level cycle accounting (debug builds only: -DIGN_TC_PROFILE).
#ifdef IGN_TC_PROFILE
__device__ unsigned long long g_tc_prof[16];
#define TC_CLK() clock64()
#define TC_ADD(slot, t0) (tc_prof_local[slot] += (unsigned long long)(clock64() - (t0)))   // flushed once at CTA end
#define TC_DECL() unsigned long long tc_prof_local[16] = {0}
#define TC_FLUSH() do { for (int i_ = 0; i_ < 16; ++i_) if (tc_prof_local[i_]) atomicAdd(&g_tc_prof[i_], tc_prof_local[i_]); } while (0)
#else
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
__global__ void __launch_bounds__(128) tc_build_b_kernel(const float* __restrict__ W, uint8_t* __restrict__ out, int M,
                                                         int K, int L, int KG, int nkb, int NKB, int N, int split,
                                                         int centre) {
  __shared__ float s_mean[16];
  const int m = blockIdx.x, kblk = blockIdx.y, kb = blockIdx.z;
  const int k0 = kblk * KG;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int u = warp; u < KG; u += 4) {
    float mean = 0.f;
    if (centre && k0 + u < K) {
      const float* src = W + ((size_t)(k0 + u) * M + m) * L;
      float s1 = 0.f;
      for (int l = lane; l < L; l += 32) s1 += __ldg(src + l);
#pragma unroll
      for (int o = 16; o; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      mean = s1 / (float)L;
    }
    if (lane == 0) s_mean[u] = mean;
  }
  __syncthreads();
  const int b_bytes = N * 128;
  uint8_t* dst = out + (((size_t)m * nkb + kblk) * NKB + kb) * (size_t)(b_bytes * (split ? 2 : 1));
  for (int task = threadIdx.x; task < N * 8; task += blockDim.x) {
    const int n = task >> 3, c = task & 7;
    const int u = n >> 4, j = n & 15;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int l = kb * kKBlock + 4 * c + e - j;
      v[e] = (k0 + u < K && l >= 0 && l < L) ? __ldg(W + ((size_t)(k0 + u) * M + m) * L + l) - s_mean[u] : 0.f;
    }
    const float4 hi = make_float4(v[0], v[1], v[2], v[3]);
    const uint32_t off = sw128_off(n, c);
    *reinterpret_cast<float4*>(dst + off) = hi;
    if (split) *reinterpret_cast<float4*>(dst + b_bytes + off) = tf32_lo(hi);
  }
}

// ---------------------------------------------------------------- kernel
__global__ void __launch_bounds__(kThreadsTC, 1) shapelet_fwd_tc_kernel(const TcGeo g, const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const long long t_entry = TC_CLK();
  TC_DECL();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x, k0 = blockIdx.y * g.KG;
  const int bbeg = blockIdx.z * g.bpc, bend = min(g.B, bbeg + g.bpc);
  const int ntile = (bend - bbeg + g.RB - 1) / g.RB;

  // ---- shared memory carve-up (stage tiles first: they need 1024-byte alignment) ----
  const int b_bytes = g.N * 128;
  const int stage_bytes = b_bytes * (g.split ? 2 : 1);
  uint8_t* stage0 = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ptr = stage0 + (size_t)g.nstage * stage_bytes;
  float* xbuf = reinterpret_cast<float*>(ptr);                      // [2][RB][XR]
  ptr += (size_t)2 * g.RB * g.XR * sizeof(float);
  float* cand_d = reinterpret_cast<float*>(ptr);                    // [128][KG]
  ptr += (size_t)kRows * g.KG * sizeof(float);
  int* cand_i = reinterpret_cast<int*>(ptr);
  ptr += (size_t)kRows * g.KG * sizeof(int);
  float* wstat = reinterpret_cast<float*>(ptr);                     // [KG] written and read by the epilogue warps only
  ptr += 16 * sizeof(float);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ptr);                // full[3], empty[3], tfull[2], tempty[2]
  uint64_t* full = bars; uint64_t* empty = bars + kMaxStages;
  uint64_t* tfull = bars + 2 * kMaxStages; uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);

  const int acc_cols = kAccCols;                                    // column pitch of one accumulator
  const uint32_t tmem_cols = 512;                                   // 2 accumulators + kMaxStages A stages
  const uint32_t a_col0 = 2 * kAccCols;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full[s], kProducerThreads / 32 + 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEpilogueThreads / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) TC_ADD(10, t_entry);       // barrier init + TMEM alloc

  if (warp < 4) {
    // =================================================================== PRODUCERS
    const int p = threadIdx.x;                                      // 0..127
    for (int i = p; i < 2 * g.RB * g.XR; i += kProducerThreads) xbuf[i] = 0.f;
    bar_sync(1, kProducerThreads);
    // first tile's series rows.  Rows are stored bank-skewed (4 floats of padding after every 32) so that the
    // 128 producer threads, whose segments start 16 samples apart, read with conflict-free LDS.128.
    auto prefetch_rows = [&](int tile, int buf) {
      const int b0 = bbeg + tile * g.RB;
      const int chunks = g.Tp / 4;
      for (int i = p; i < g.RB * chunks; i += kProducerThreads) {
        const int bl = i / chunks, c = i - bl * chunks;
        if (b0 + bl < bend)
          cp_async16(xbuf + ((size_t)buf * g.RB + bl) * g.XR + c * 4 + 4 * (c >> 3),
                     a.xn + ((size_t)(b0 + bl) * g.M + m) * g.Tp + c * 4);
      }
      cp_async_commit();
    };
    if (ntile > 0) prefetch_rows(0, 0);
    if (p == 0) TC_ADD(11, t_entry);                 // producer prologue done
    // Per-thread constants.  A: this thread owns accumulator row p = bl*RI + i, i.e. samples 16 i + 32 kb + (0..31)
    // of series bl.  B task u covers shapelet u, shift j = p >> 3, 16-byte chunk c = p & 7.
    const int a_bl = p / g.RI, a_i = p - a_bl * g.RI;
    int a_off[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int idx = a_i * kShifts + 4 * c;
      a_off[c] = a_bl < g.RB ? a_bl * g.XR + idx + 4 * (idx >> 5) : -1;
    }
    const uint8_t* bsrc = a.btiles + ((size_t)m * g.nkb + blockIdx.y) * g.NKB * (size_t)stage_bytes;
    const uint32_t a_lane = tmem_base + ((uint32_t)(warp * 32) << 16) + a_col0;
    uint32_t it = 0;                                                 // global stage counter
    for (int tile = 0; tile < ntile; ++tile) {
      const int buf = tile & 1;
      const int nb = min(g.RB, bend - (bbeg + tile * g.RB));
      long long tr0 = TC_CLK();
      cp_async_wait_all();
      bar_sync(1, kProducerThreads);      // rows of this tile landed; everyone is done with the other buffer
      if (p == 0) TC_ADD(3, tr0);
      if (tile + 1 < ntile) prefetch_rows(tile + 1, buf ^ 1);
      const float* xb = xbuf + (size_t)buf * g.RB * g.XR;
      const bool a_live = a_off[0] >= 0 && a_bl < nb;
      for (int kb = 0; kb < g.NKB; ++kb, ++it) {
        const int s = it % g.nstage;
        const uint32_t ph = (it / g.nstage) & 1;
        long long tp0 = TC_CLK();
        mbar_wait(&empty[s], ph ^ 1);                                // passes immediately on the first lap
        tc_fence_after();
        if (p == 0) {   // B stage: one bulk TMA copy of the pre-swizzled tile image, counted in bytes on full[s]
          mbar_arrive_expect_tx(&full[s], (uint32_t)stage_bytes);
          tma_bulk_g2s(stage0 + (size_t)s * stage_bytes, bsrc + (size_t)kb * stage_bytes, (uint32_t)stage_bytes, &full[s]);
          TC_ADD(0, tp0);
        }
        tp0 = TC_CLK();
        // A stage -> tensor memory: 32 columns of raw fp32 (the MMA truncates to tf32 = hi) and 32 of lo
        {
          uint32_t hi[32], lo[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a_live) v = *reinterpret_cast<const float4*>(xb + a_off[c] + kb * 36);   // 32 samples + 4 pad
            const float4 l4 = tf32_lo(v);
            hi[4 * c] = __float_as_uint(v.x); hi[4 * c + 1] = __float_as_uint(v.y);
            hi[4 * c + 2] = __float_as_uint(v.z); hi[4 * c + 3] = __float_as_uint(v.w);
            lo[4 * c] = __float_as_uint(l4.x); lo[4 * c + 1] = __float_as_uint(l4.y);
            lo[4 * c + 2] = __float_as_uint(l4.z); lo[4 * c + 3] = __float_as_uint(l4.w);
          }
          tmem_st32(a_lane + s * kAStageCols, hi);
          if (g.split) tmem_st32(a_lane + s * kAStageCols + 32, lo);
        }
        if (p == 0) TC_ADD(1, tp0);
        tp0 = TC_CLK();
        tmem_st_wait();                                              // A columns written
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
        if (p == 0) TC_ADD(2, tp0);
      }
    }
  } else if (warp < 8) {
    // =================================================================== EPILOGUE
    const int e = threadIdx.x - kProducerThreads;                   // 0..127 = accumulator row
    const int ew = warp - 4;                                        // TMEM lane quarter
    const int bl = e / g.RI, i = e - bl * g.RI;
    const float invL = 1.f / (float)g.L;
    for (int kl = ew; kl < g.KG; kl += 4) {       // shapelet statistics (same arithmetic as the FP32 engine)
      const int k = k0 + kl;
      float stat = 0.f;
      if (k < g.K) {
        const float* src = a.W + ((size_t)k * g.M + m) * g.L;
        float s1 = 0.f, s2 = 0.f;
        for (int l = lane; l < g.L; l += 32) { const float w = __ldg(src + l); s1 += w; s2 = fmaf(w, w, s2); }
#pragma unroll
        for (int o = 16; o; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
        if (g.dist == IGN_DIST_PEARSON) {
          const float mean = s1 / (float)g.L;
          float c2 = 0.f;
          for (int l = lane; l < g.L; l += 32) { const float w = __ldg(src + l) - mean; c2 = fmaf(w, w, c2); }
#pragma unroll
          for (int o = 16; o; o >>= 1) c2 += __shfl_xor_sync(0xffffffffu, c2, o);
          stat = sqrtf(c2);
        } else if (g.dist == IGN_DIST_COSINE) {
          stat = 1.f / fmaxf(sqrtf(s2), 1e-8f);
        } else {
          stat = s2;
        }
      }
      if (lane == 0) wstat[kl] = stat;
    }
    bar_sync(2, kEpilogueThreads);
    for (int tile = 0; tile < ntile; ++tile) {
      const int acc = tile & 1;
      const int b0 = bbeg + tile * g.RB;
      const int nb = min(g.RB, bend - b0);
      const bool row_live = bl < g.RB && bl < nb;
      const int b = b0 + bl;
      const int t0 = i * kShifts;
      // this row's 16 window norm terms (fp32, from the window-statistics pass): four 16-byte loads issued
      // before the accumulator wait so their latency hides under the MMAs of this tile
      float4 xs4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) xs4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_live && t0 < a.SP) {
        const float4* sp = reinterpret_cast<const float4*>(a.st0 + ((size_t)b * g.M + m) * a.SP + t0);
#pragma unroll
        for (int j = 0; j < 4; ++j) xs4[j] = __ldg(sp + j);
      }
      long long te0 = TC_CLK();
      mbar_wait(&tfull[acc], (tile >> 1) & 1);
      if (e == 0) TC_ADD(7, te0);
      te0 = TC_CLK();
      tc_fence_after();
      const float xst[16] = {xs4[0].x, xs4[0].y, xs4[0].z, xs4[0].w, xs4[1].x, xs4[1].y, xs4[1].z, xs4[1].w,
                             xs4[2].x, xs4[2].y, xs4[2].z, xs4[2].w, xs4[3].x, xs4[3].y, xs4[3].z, xs4[3].w};
      const uint32_t trow = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * acc_cols);
      for (int kl = 0; kl < g.KG; ++kl) {
        uint32_t v[16];
        tmem_ld16(trow + kl * 16, v);
        tmem_ld_wait();
        const int k = k0 + kl;
        float best = INFINITY; int bi = 0x7fffffff;
        float dv[16];
        const float wst = wstat[kl];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float raw = __uint_as_float(v[j]);
          float d;
          if (g.dist == IGN_DIST_SQL2) d = fmaxf((xst[j] + wst - 2.f * raw) * invL, 0.f);
          else if (g.dist == IGN_DIST_COSINE) d = 1.f - raw * xst[j] * wst;
          else d = 1.f - __fdividef(raw, xst[j] * wst + 1e-8f);          // norms hoisted; no IEEE slow paths on zeros
          const bool valid = row_live && k < g.K && (t0 + j) < g.Tw;
          dv[j] = valid ? d : 0.f;
          if (valid && d < best) { best = d; bi = t0 + j; }
        }
        if (a.dstore && row_live && k < g.K) {
          float* dg = a.dstore + (((size_t)b * g.M + m) * g.K + k) * g.Ts + t0;
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            if (t0 + j < g.Ts) *reinterpret_cast<float4*>(dg + j) = make_float4(dv[j], dv[j + 1], dv[j + 2], dv[j + 3]);
        }
        cand_d[e * g.KG + kl] = best;
        cand_i[e * g.KG + kl] = bi;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);                      // this warp's quarter of the accumulator is drained
      if (e == 0) TC_ADD(8, te0);
      te0 = TC_CLK();
      bar_sync(2, kEpilogueThreads);
      // per (sample, shapelet) arg-min over the RI rows of that sample, first index on ties: one warp per pair
      for (int pair = ew; pair < g.RB * g.KG; pair += 4) {
        const int rbl = pair / g.KG, kl = pair - rbl * g.KG;
        const int k = k0 + kl;
        if (rbl >= nb || k >= g.K) continue;
        float dmn = INFINITY; int imn = 0x7fffffff;
        for (int ii = lane; ii < g.RI; ii += 32) {
          const float d = cand_d[(rbl * g.RI + ii) * g.KG + kl];
          if (d < dmn) { dmn = d; imn = cand_i[(rbl * g.RI + ii) * g.KG + kl]; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
          const float od = __shfl_xor_sync(0xffffffffu, dmn, o);
          const int oi = __shfl_xor_sync(0xffffffffu, imn, o);
          if (od < dmn || (od == dmn && oi < imn)) { dmn = od; imn = oi; }
        }
        if (lane == 0) {
          const size_t o = ((size_t)(b0 + rbl) * g.K + k) * g.M + m;
          float pv;
          if (g.pool == IGN_POOL_RBF_MAX) { const float ed = g.eps * dmn; pv = expf(-(ed * ed)); }
          else pv = 1.f / (1.f + expf(-(a.thr[(size_t)k * g.M + m] - dmn)));
          a.p[o] = pv; a.dmin[o] = dmn;
          if (a.argmin) a.argmin[o] = imn;
        }
      }
      bar_sync(2, kEpilogueThreads);                                 // cand arrays free for the next tile
      if (e == 0) TC_ADD(9, te0);
    }
  } else {
    // =================================================================== MMA ISSUER (one elected thread)
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_tf32(kRows, g.N);
      uint32_t it = 0;
      for (int tile = 0; tile < ntile; ++tile) {
        const int acc = tile & 1;
        long long tm0 = TC_CLK();
        mbar_wait(&tempty[acc], ((tile >> 1) & 1) ^ 1);              // passes immediately for the first two tiles
        TC_ADD(4, tm0);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_cols);
        for (int kb = 0; kb < g.NKB; ++kb, ++it) {
          const int s = it % g.nstage;
          tm0 = TC_CLK();
          mbar_wait(&full[s], (it / g.nstage) & 1);
          TC_ADD(5, tm0);
          tm0 = TC_CLK();
          tc_fence_after();
          const uint32_t sb_hi = smem_u32(stage0 + (size_t)s * stage_bytes);
          const uint32_t sb_lo = sb_hi + b_bytes;
          const uint32_t a_hi = tmem_base + a_col0 + s * kAStageCols, a_lo = a_hi + 32;
#pragma unroll
          for (int k8 = 0; k8 < kKBlock / 8; ++k8) {
            const uint32_t koff = k8 * 32;                           // 8 tf32 = 32 bytes inside the swizzle atom
            if (g.split) {                                           // small terms first
              umma_tf32_ts(d_tmem, a_lo + k8 * 8, umma_desc_sw128(sb_hi + koff), idesc, (kb | k8) != 0);
              umma_tf32_ts(d_tmem, a_hi + k8 * 8, umma_desc_sw128(sb_lo + koff), idesc, 1);
              umma_tf32_ts(d_tmem, a_hi + k8 * 8, umma_desc_sw128(sb_hi + koff), idesc, 1);
            } else {
              umma_tf32_ts(d_tmem, a_hi + k8 * 8, umma_desc_sw128(sb_hi + koff), idesc, (kb | k8) != 0);
            }
          }
          umma_commit(&empty[s]);                                    // stage reusable once these MMAs retire
          TC_ADD(6, tm0);
        }
        umma_commit(&tfull[acc]);                                    // accumulator complete
      }
    }
    __syncwarp();
  }

  if (threadIdx.x == 0) TC_ADD(12, t_entry);         // producer thread 0 finished its loop
  if (threadIdx.x == kProducerThreads) TC_ADD(13, t_entry);   // epilogue thread 0 finished
  if (threadIdx.x == 256) TC_ADD(14, t_entry);       // MMA thread finished
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, tmem_cols);
  if (threadIdx.x == 0) TC_ADD(15, t_entry);         // CTA lifetime
  if (threadIdx.x == 0 || threadIdx.x == kProducerThreads || threadIdx.x == 256) TC_FLUSH();
}

size_t tc_smem_bytes(const TcGeo& g, int nstage) {
  const size_t stage = (size_t)(g.N * 128) * (g.split ? 2 : 1);
  return nstage * stage + (size_t)2 * g.RB * g.XR * 4 + (size_t)kRows * g.KG * 8 + 64 +
         (2 * kMaxStages + 4) * 8 + 16 + 1024;
}

}  // namespace

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
  g.Tw = num_windows(d.T, d.L, 1); g.Ts = round_up(g.Tw, 4);
  g.RI = ceil_div(g.Tw, kShifts); g.RB = max(1, min(kRows / g.RI, d.B));
  g.nkb = ceil_div(d.K, 8); g.KG = ceil_div(d.K, g.nkb); g.N = 16 * g.KG;   // N <= 128 = kAccCols
  g.NKB = ceil_div(d.L + kShifts - 1, kKBlock);
  {  // bank-skewed series rows: 36 floats per 32 samples
    const int span = max(d.Tp, (g.RI - 1) * kShifts + g.NKB * kKBlock) + 32;
    g.XR = round_up(span + 4 * (span / 32) + 8, 4);
  }
  g.dist = d.dist; g.pool = d.pool; g.eps = d.eps;
  g.split = d.precision == IGN_PREC_3XTF32 ? 1 : 0;
}

// bytes of the pre-swizzled B-tile workspace the tcgen05 forward needs
size_t shapelet_fwd_tc_workspace(const ign_shapelet_desc& d) {
  TcGeo g;
  tc_geo(d, g);
  return (size_t)d.M * g.nkb * g.NKB * (size_t)(g.N * 128) * (g.split ? 2 : 1);
}

bool shapelet_fwd_tc_supported(const ign_shapelet_desc& d) {
  if (d.dist == IGN_DIST_L1 || d.stride != 1) return false;
  if (d.precision != IGN_PREC_3XTF32 && d.precision != IGN_PREC_TF32) return false;
  const int Tw = num_windows(d.T, d.L, 1);
  if (Tw <= 0 || ceil_div(Tw, kShifts) > kRows) return false;     // one sample must fit a 128-row tile
  return true;
}

int launch_shapelet_fwd_tc(const ign_shapelet_desc& d, const float* xn, const float* st0,
                           const float* W, const float* thr, float* p, float* dmin, int* argmin, float* dstore,
                           void* ws, size_t ws_bytes, cudaStream_t st) {
  TcGeo g;
  tc_geo(d, g);
  const size_t need = shapelet_fwd_tc_workspace(d);
  if (!ws || ws_bytes < need) { set_error("shapelet_forward(tcgen05): workspace %zu < %zu bytes (ign_shapelet_forward_workspace)", ws_bytes, need); return IGN_ERR_INVALID; }
  if (((uintptr_t)ws & 127) != 0) { set_error("shapelet_forward(tcgen05): workspace must be 128-byte aligned"); return IGN_ERR_INVALID; }
  const size_t cap = (size_t)max_optin_smem();
  g.nstage = kMaxStages;
  while (g.nstage > 1 && tc_smem_bytes(g, g.nstage) > cap) --g.nstage;
  if (tc_smem_bytes(g, g.nstage) > cap) { set_error("shapelet_forward(tcgen05): L=%d K=%d does not fit shared memory", d.L, d.K); return IGN_ERR_UNSUPPORTED; }
  // 1. shifted-shapelet operand, once per launch, already in the swizzled tile image
  tc_build_b_kernel<<<dim3(d.M, g.nkb, g.NKB), 128, 0, st>>>(W, reinterpret_cast<uint8_t*>(ws), d.M, d.K, d.L, g.KG,
                                                            g.nkb, g.NKB, g.N, g.split, d.dist == IGN_DIST_PEARSON);
  IGN_CUDA(cudaGetLastError());
  // 2. main kernel, one CTA per SM: spread (channel, shapelet block) over batch chunks to ~4 waves
  const int per_chunk = d.M * g.nkb;
  int nchunk = max(1, ceil_div(4 * sm_count(), per_chunk));
  nchunk = min(nchunk, ceil_div(d.B, g.RB));
  g.bpc = round_up(ceil_div(d.B, nchunk), g.RB);
  // the kernel allocates all 512 TMEM columns: keep it to one CTA per SM by asking for more than half the smem
  const size_t smem = max(tc_smem_bytes(g, g.nstage), (size_t)118 * 1024);
  IGN_CUDA(cudaFuncSetAttribute(shapelet_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TcArgs a{xn, st0, W, thr, p, dmin, argmin, dstore, stats_pitch(d.T, d.L, 1), reinterpret_cast<const uint8_t*>(ws)};
  dim3 grid(d.M, g.nkb, ceil_div(d.B, g.bpc));
  shapelet_fwd_tc_kernel<<<grid, kThreadsTC, smem, st>>>(g, a);
  IGN_CUDA(cudaGetLastError());
  return IGN_OK;
}

}  // namespace ign
