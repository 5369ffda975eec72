// PTX wrappers shared by the tcgen05 kernels (forward: shapelet_tc.cu, backward: shapelet_tc_bwd.cu): mbarrier, bulk TMA,
// tcgen05 alloc / st / ld / mma / commit, cp.async, named barriers, swizzle helpers.  sm_100a only.
// Include INSIDE `namespace ign { namespace {` (every translation unit gets its own internal-linkage copy).
#pragma once

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // try_wait with a suspend-time hint: the waiting thread is parked by the hardware (woken when the phase
  // completes) instead of polling — with 18 warps per SM the polling loops otherwise eat the issue slots of the
  // warps doing the work (measured: 73 M polling instructions per launch in the producer loop alone).
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  unsigned spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(200000u)
        : "memory");
    if (!done && ++spins > (1u << 14)) __trap();   // a pipeline bug must fail fast, never hang the GPU
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
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads, TMA)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
// D[tmem] (+)= A[tmem] * B[smem], kind::tf32, issued by one thread.  The shared-memory descriptor is given as its low
// word only (start address >> 4; the high word — SBO,
// version, swizzle mode — is one constant register): the issuing thread then spends one 32-bit add per MMA on
// descriptors instead of a shift/mask/mov sequence (-5 % on every geometry: its instruction stream paces the kernel).
constexpr uint32_t kDescHiSw128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // bits 32.. of the descriptor below
__device__ __forceinline__ void umma_tf32_ts_lo(uint32_t d_tmem, uint32_t a_tmem, uint32_t bdesc_lo, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 bd, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(bdesc_lo), "r"(idesc), "r"(accumulate), "r"(kDescHiSw128)
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
// 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
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
// (Built as low word = start>>4, high word = kDescHiSw128, see umma_tf32_ts_lo.)
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

