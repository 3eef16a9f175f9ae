// Inline-PTX helpers shared by the tcgen05 kernels of this library (sm_100a only): mbarriers, TMA loads (plain and
// cluster-multicast), tcgen05.mma / commit / ld, UMMA shared-memory descriptors for K-major 64B-swizzled fp16 tiles,
// cluster barriers and DSMEM stores.  Same forms as decode.cu (where they were first proven on B200).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace icrl_tc {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol mistake must trap, never hang the device.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  for (unsigned spins = 0; spins < (1u << 27); ++spins) {
    unsigned ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
// {k, row, part} box: the hi and lo' tiles of an operand arrive with ONE instruction.
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mcast(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar,
                                                  unsigned short mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
// shared -> global tensor store of one {c0, c1, c2} box (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, unsigned src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the shared-memory source of every committed store has been read (the global writes may still be in flight)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_commit_mcast(unsigned bar, unsigned short mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_mma(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc,
                                       unsigned accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc),
      "r"(accumulate) : "memory");
}
// K-major, 64B-swizzle operand descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 = 1 [16,30) (unused
// for swizzled K-major), SBO>>4 = 32 [32,46) (8 rows x 64 B between row groups), version = 1 [46,48), layout
// SWIZZLE_64B = 4 [61,64).
__device__ __forceinline__ unsigned long long smem_desc_sw64(unsigned addr) {
  return (unsigned long long)((addr & 0x3FFFF) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
// c_format F32 [4,6) = 1, a/b_format F16 = 0, K-major, n_dim = N>>3 [17,23), m_dim = M>>4 [24,29)
constexpr unsigned idesc_f16_m128(int n) { return (1u << 4) | ((unsigned)(n >> 3) << 17) | ((unsigned)(128 >> 4) << 24); }

// TMEM -> registers.  The loads and their wait live in ONE asm statement so that no consumer of the destination
// registers can be scheduled before tcgen05.wait::ld.
__device__ __forceinline__ void tmem_ld8x2(unsigned ta, unsigned tb, float* a, float* b) {
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%16];\n"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8, %9, %10, %11, %12, %13, %14, %15}, [%17];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(ta), "r"(tb)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[8 + i]); }
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
// arrive without memory ordering: for callers that have fenced what the peers need themselves (a release arrive would also
// wait for every other outstanding global store of the thread)
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// raw special-function forms (the same MUFU instructions __expf / __fdividef end in, without their range fix-ups)
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr float L2E = 1.4426950408889634f;          // log2(e)
constexpr float SAT = 20.f * L2E;                   // exponent clamp: sigmoid(+-20) and tanh(+-10) are saturated to the last float bit
__device__ __forceinline__ float clamp_sat(float y) { return fminf(fmaxf(y, -SAT), SAT); }
// tanh(x) = 1 - 2 / (1 + e^(2x))
__device__ __forceinline__ float tanh_lean(float x) { return fmaf(-2.f, rcp_ftz(1.f + ex2_ftz(clamp_sat(x * (2.f * L2E)))), 1.f); }

// fp32 -> {hi, lo'} fp16 pair of four values: x = hi + lo'/2048 (22 mantissa bits; the scaling keeps lo' normal)
__device__ __forceinline__ void split4_f16(const float4 v, uint2& hi, uint2& lo) {
  const __half a0 = __float2half_rn(v.x), a1 = __float2half_rn(v.y), a2 = __float2half_rn(v.z), a3 = __float2half_rn(v.w);
  __half2 h2[2] = {__halves2half2(a0, a1), __halves2half2(a2, a3)};
  __half2 l2[2] = {__halves2half2(__float2half_rn((v.x - __half2float(a0)) * 2048.f), __float2half_rn((v.y - __half2float(a1)) * 2048.f)),
                   __halves2half2(__float2half_rn((v.z - __half2float(a2)) * 2048.f), __float2half_rn((v.w - __half2float(a3)) * 2048.f))};
  hi = *reinterpret_cast<const uint2*>(h2);
  lo = *reinterpret_cast<const uint2*>(l2);
}

}  // namespace icrl_tc
