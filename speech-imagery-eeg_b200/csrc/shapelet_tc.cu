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
//  * the overlapping rows are never materialised in HBM or L2: producer warps build the 128B-swizzled
//    K-major A tile straight from the raw series row in shared memory (one LDS.128 + two STS.128 per 4
//    elements), so DRAM/L2 sees each series row once.
//
// Precision: tcgen05 has no fp32 kind.  kind::tf32 reads fp32 bits and drops the low 13 mantissa bits, so
// the "hi" operand is the raw fp32 value and lo = x - trunc_tf32(x) is formed by the producers.
//   IGN_PREC_3XTF32: hi*hi + hi*lo + lo*hi  (3 MMAs per k-step, fp32-equivalent: error ~2^-21)
//   IGN_PREC_TF32  : hi*hi only             (own, looser tolerance)
//
// Roles (one CTA per SM, 288 threads):
//   warps 0-3  producers: cp.async the next series rows, build A/B stage tiles, arrive on full[stage]
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
constexpr int kMaxStages = 3;

struct TcGeo {
  int B, M, T, Tp, K, L;
  int Tw, Ts;          // windows, dstore pitch
  int RI, RB;          // window groups (of 16) per sample, samples per 128-row tile
  int KG, nkb;         // shapelets per N tile, number of shapelet blocks
  int N;               // 16*KG
  int NKB;             // 32-wide k-blocks: ceil((L+15)/32)
  int XR;              // floats per series row in smem
  int WR;              // floats per shifted shapelet row in smem
  int nstage;
  int bpc;             // samples per CTA chunk
  int dist, pool, split;
  float eps;
};

struct TcArgs {
  const float* xn; const float* st0; const float* W; const float* thr;   // st0: window statistics [B,M,SP]
  float* p; float* dmin; int* argmin; float* dstore;
  int SP;
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
#define TC_ADD(slot, t0) atomicAdd(&g_tc_prof[slot], (unsigned long long)(clock64() - (t0)))
#else
#define TC_CLK() 0ll
#define TC_ADD(slot, t0) ((void)(t0))
#endif

// ---------------------------------------------------------------- kernel
__global__ void __launch_bounds__(kThreadsTC, 1) shapelet_fwd_tc_kernel(const TcGeo g, const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x, k0 = blockIdx.y * g.KG;
  const int bbeg = blockIdx.z * g.bpc, bend = min(g.B, bbeg + g.bpc);
  const int ntile = (bend - bbeg + g.RB - 1) / g.RB;

  // ---- shared memory carve-up (stage tiles first: they need 1024-byte alignment) ----
  const int a_bytes = kRows * 128, b_bytes = g.N * 128;
  const int stage_bytes = (a_bytes + b_bytes) * (g.split ? 2 : 1);
  uint8_t* stage0 = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ptr = stage0 + (size_t)g.nstage * stage_bytes;
  float* wsh = reinterpret_cast<float*>(ptr);                       // [4][KG][WR]
  ptr += (size_t)4 * g.KG * g.WR * sizeof(float);
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

  const int acc_cols = g.N <= 128 ? 128 : 256;                      // column pitch of one accumulator
  const uint32_t tmem_cols = 2 * acc_cols;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full[s], kProducerThreads / 32); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEpilogueThreads / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // =================================================================== PRODUCERS
    const int p = threadIdx.x;                                      // 0..127
    // shifted, zero-padded shapelet rows: wsh[s][k][kWshPad + v] = w'[k][v - s]
    for (int i = p; i < 4 * g.KG * g.WR; i += kProducerThreads) wsh[i] = 0.f;
    for (int i = p; i < 2 * g.RB * g.XR; i += kProducerThreads) xbuf[i] = 0.f;
    bar_sync(1, kProducerThreads);
    for (int kl = warp; kl < g.KG; kl += 4) {
      const int k = k0 + kl;
      if (k < g.K) {
        const float* src = a.W + ((size_t)k * g.M + m) * g.L;
        float mean = 0.f;
        if (g.dist == IGN_DIST_PEARSON) {
          float s1 = 0.f;
          for (int l = lane; l < g.L; l += 32) s1 += __ldg(src + l);
#pragma unroll
          for (int o = 16; o; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          mean = s1 / (float)g.L;
        }
        for (int l = lane; l < g.L; l += 32) {
          const float w = __ldg(src + l) - mean;
#pragma unroll
          for (int s = 0; s < 4; ++s) wsh[((size_t)s * g.KG + kl) * g.WR + kWshPad + l + s] = w;
        }
      }
    }
    // first tile's series rows
    auto prefetch_rows = [&](int tile, int buf) {
      const int b0 = bbeg + tile * g.RB;
      const int chunks = g.Tp / 4;
      for (int i = p; i < g.RB * chunks; i += kProducerThreads) {
        const int bl = i / chunks, c = i - bl * chunks;
        if (b0 + bl < bend)
          cp_async16(xbuf + ((size_t)buf * g.RB + bl) * g.XR + c * 4,
                     a.xn + ((size_t)(b0 + bl) * g.M + m) * g.Tp + c * 4);
      }
      cp_async_commit();
    };
    if (ntile > 0) prefetch_rows(0, 0);
    // Per-thread constants: A task q covers row r = 16 q + (p >> 3), chunk c = p & 7; B task u covers shapelet
    // u, shift j = p >> 3, chunk c.  Swizzled destinations differ only by q * 2048 / u * 2048 bytes.
    const int pr = p >> 3, pc = p & 7;
    const uint32_t dst0 = (uint32_t)(pr * 128 + ((pc ^ (pr & 7)) << 4));
    const int b_src0 = (pr & 3) * g.KG * g.WR + kWshPad + 4 * (pc - (pr >> 2));
    int a_src[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int r = q * 16 + pr;
      const int bl = r / g.RI, i = r - bl * g.RI;
      a_src[q] = bl < g.RB ? bl * g.XR + i * kShifts + pc * 4 : -1;
    }
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
      const int live_lim = nb * g.XR;                                // a_src below this offset belongs to a live sample
      for (int kb = 0; kb < g.NKB; ++kb, ++it) {
        const int s = it % g.nstage;
        const uint32_t ph = (it / g.nstage) & 1;
        long long tp0 = TC_CLK();
        mbar_wait(&empty[s], ph ^ 1);                                // passes immediately on the first lap
        if (p == 0) TC_ADD(0, tp0);
        tp0 = TC_CLK();
        uint8_t* sa_hi = stage0 + (size_t)s * stage_bytes + dst0;
        uint8_t* sb_hi = sa_hi + a_bytes;
        const int lo_off = a_bytes + b_bytes;
        // A tile: row r = bl*RI + i holds x[bl][16 i + 32 kb + (0..31)]
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a_src[q] >= 0 && a_src[q] < live_lim) v = *reinterpret_cast<const float4*>(xb + a_src[q] + kb * kKBlock);
          *reinterpret_cast<float4*>(sa_hi + q * 2048) = v;
          if (g.split) *reinterpret_cast<float4*>(sa_hi + q * 2048 + lo_off) = tf32_lo(v);
        }
        // B tile: row n = 16 u + j holds w'[u][32 kb + (0..31) - j]
        for (int u = 0; u < g.KG; ++u) {
          const float4 v = *reinterpret_cast<const float4*>(wsh + b_src0 + u * g.WR + kb * kKBlock);
          *reinterpret_cast<float4*>(sb_hi + u * 2048) = v;
          if (g.split) *reinterpret_cast<float4*>(sb_hi + u * 2048 + lo_off) = tf32_lo(v);
        }
        if (p == 0) TC_ADD(1, tp0);
        tp0 = TC_CLK();
        fence_proxy_async();                                         // generic-proxy writes -> async proxy (UMMA)
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
          stat = c2;
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
          else d = 1.f - __fdividef(raw, sqrtf(xst[j] * wst) + 1e-8f);   // no IEEE slow path on zero numerators
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
    // =================================================================== MMA ISSUER (one thread)
    if (lane == 0) {
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
          const uint32_t sa_hi = smem_u32(stage0 + (size_t)s * stage_bytes);
          const uint32_t sb_hi = sa_hi + a_bytes;
          const uint32_t sa_lo = sb_hi + b_bytes;
          const uint32_t sb_lo = sa_lo + a_bytes;
#pragma unroll
          for (int k8 = 0; k8 < kKBlock / 8; ++k8) {
            const uint32_t koff = k8 * 32;                           // 8 tf32 = 32 bytes inside the swizzle atom
            if (g.split) {                                           // small terms first
              umma_tf32(d_tmem, umma_desc_sw128(sa_lo + koff), umma_desc_sw128(sb_hi + koff), idesc, (kb | k8) != 0);
              umma_tf32(d_tmem, umma_desc_sw128(sa_hi + koff), umma_desc_sw128(sb_lo + koff), idesc, 1);
              umma_tf32(d_tmem, umma_desc_sw128(sa_hi + koff), umma_desc_sw128(sb_hi + koff), idesc, 1);
            } else {
              umma_tf32(d_tmem, umma_desc_sw128(sa_hi + koff), umma_desc_sw128(sb_hi + koff), idesc, (kb | k8) != 0);
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

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, tmem_cols);
}

size_t tc_smem_bytes(const TcGeo& g, int nstage) {
  const size_t stage = (size_t)(kRows * 128 + g.N * 128) * (g.split ? 2 : 1);
  return nstage * stage + (size_t)4 * g.KG * g.WR * 4 + (size_t)2 * g.RB * g.XR * 4 + (size_t)kRows * g.KG * 8 + 64 +
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

bool shapelet_fwd_tc_supported(const ign_shapelet_desc& d) {
  if (d.dist == IGN_DIST_L1 || d.stride != 1) return false;
  if (d.precision != IGN_PREC_3XTF32 && d.precision != IGN_PREC_TF32) return false;
  const int Tw = num_windows(d.T, d.L, 1);
  if (Tw <= 0 || ceil_div(Tw, kShifts) > kRows) return false;     // one sample must fit a 128-row tile
  return true;
}

int launch_shapelet_fwd_tc(const ign_shapelet_desc& d, const float* xn, const float* st0,
                           const float* W, const float* thr, float* p, float* dmin, int* argmin, float* dstore,
                           cudaStream_t st) {
  TcGeo g;
  g.B = d.B; g.M = d.M; g.T = d.T; g.Tp = d.Tp; g.K = d.K; g.L = d.L;
  g.Tw = num_windows(d.T, d.L, 1); g.Ts = round_up(g.Tw, 4);
  g.RI = ceil_div(g.Tw, kShifts); g.RB = max(1, min(kRows / g.RI, d.B));
  g.nkb = ceil_div(d.K, 8); g.KG = ceil_div(d.K, g.nkb); g.N = 16 * g.KG;
  g.NKB = ceil_div(d.L + kShifts - 1, kKBlock);
  g.XR = round_up(max(d.Tp, (g.RI - 1) * kShifts + g.NKB * kKBlock) + 8, 4);
  g.WR = round_up(kWshPad + g.NKB * kKBlock + 16, 4);
  g.dist = d.dist; g.pool = d.pool; g.eps = d.eps;
  g.split = d.precision == IGN_PREC_3XTF32 ? 1 : 0;
  const size_t cap = (size_t)max_optin_smem();
  g.nstage = kMaxStages;
  while (g.nstage > 1 && tc_smem_bytes(g, g.nstage) > cap) --g.nstage;
  if (tc_smem_bytes(g, g.nstage) > cap) { set_error("shapelet_forward(tcgen05): L=%d K=%d does not fit shared memory", d.L, d.K); return IGN_ERR_UNSUPPORTED; }
  // one CTA per SM: spread (channel, shapelet block) over batch chunks to ~4 waves
  const int per_chunk = d.M * g.nkb;
  int nchunk = max(1, ceil_div(4 * sm_count(), per_chunk));
  nchunk = min(nchunk, ceil_div(d.B, g.RB));
  g.bpc = round_up(ceil_div(d.B, nchunk), g.RB);
  const size_t smem = tc_smem_bytes(g, g.nstage);
  IGN_CUDA(cudaFuncSetAttribute(shapelet_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TcArgs a{xn, st0, W, thr, p, dmin, argmin, dstore, stats_pitch(d.T, d.L, 1)};
  dim3 grid(d.M, g.nkb, ceil_div(d.B, g.bpc));
  shapelet_fwd_tc_kernel<<<grid, kThreadsTC, smem, st>>>(g, a);
  IGN_CUDA(cudaGetLastError());
  return IGN_OK;
}

}  // namespace ign
