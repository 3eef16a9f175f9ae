// Weight-gradient contraction of the value chain on tcgen05:
//   dW_hh[m][n] = sum_t dgates[t][m] * h_prev[t][n],   M = 2048, N = 512, K = T (number of serial steps, up to ~10^6).
// The operands arrive time-major ([T][M], [T][N] fp32), so two pre-passes put them into the K-major fp16-split form
// the MMA wants:
//   1. column |max| of dgates over t  -> a power-of-two scale per gate row m (gradients are ~1e-4 .. 1e-12 and
//      would fall into fp16 subnormals / underflow; h is in (-1, 1) and needs none)
//   2. transpose + split: x*s = hi + lo'/2048 (fp16 pair, 22 mantissa bits) -> [2][rows][Tp] with Tp = T rounded to 64
// GEMM: 128x128 tiles, split-K over gridDim.z, 3-stage TMA ring of 64-wide K blocks (hi and lo' of an operand in ONE
// 3-D box), 2 MMAs per K step (hi * [hi | lo'] -> [main | correction], lo'*hi -> correction accumulator).  Because the tensor core
// truncates on every accumulate (gemm_tc.cu header), a TMEM accumulator only ever sums 512 K elements: after 8 K blocks
// the epilogue warps add main + correction/2048 into fp32 REGISTERS (round to nearest) while the MMAs continue in the
// second accumulator set.  Each split writes its partial tile; a last kernel sums the splits and undoes the row scale.
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 64, UMMA_K = 16, STAGES = 3;
constexpr int OP_TILE = 2 * BM * BK * 2;                   // hi + lo' of one operand: 32 KB
constexpr int STAGE_BYTES = 2 * OP_TILE;                   // 64 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int FLUSH_KB = 8;                                // K blocks per TMEM accumulation (512 K elements)
constexpr int TMEM_COLS = 512;                             // 2 sets x (main 128 + correction 128)
constexpr int THREADS = 192;
constexpr int CLN = 4;                                      // CTAs of a cluster: the 4 N tiles of one (M tile, split) share A
constexpr int A_SLICE_ROWS = BM / CLN, A_SLICE = A_SLICE_ROWS * BK * 2;    // 32 rows = 4 KB per part fetched per CTA
constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f;

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
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_3d_mcast(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned bar,
                                                  unsigned short mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mcast(unsigned bar, unsigned short mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
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
__device__ __forceinline__ void tc_mma(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc,
                                       unsigned accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc),
      "r"(accumulate) : "memory");
}
// K-major, 128B swizzle (see gemm_tc.cu)
__device__ __forceinline__ unsigned long long smem_desc(unsigned addr) {
  return (unsigned long long)((addr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// F32 accumulate, fp16 x fp16, K-major, N = 128, M = 128
constexpr unsigned IDESC = (1u << 4) | ((unsigned)(BN >> 3) << 17) | ((unsigned)(BM >> 4) << 24);
constexpr unsigned IDESC2 = (1u << 4) | ((unsigned)(2 * BN >> 3) << 17) | ((unsigned)(BM >> 4) << 24);     // N = 256

__device__ __forceinline__ void tmem_ld32x2(unsigned ta, unsigned tb, float* a, float* b) {
  unsigned r[64];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%64];\n"
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, "
      "%47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%65];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
        "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
        "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
        "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(ta), "r"(tb)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[32 + i]); }
}

// ------------------------------------------------------------------------------------------------ pre-passes
__global__ void col_absmax_kernel(long long T, int C, const float* __restrict__ X, int ldx, unsigned* __restrict__ mx) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const long long per = (T + gridDim.y - 1) / gridDim.y;
  const long long t0 = (long long)blockIdx.y * per, t1 = min(T, t0 + per);
  float m = 0.f;
  for (long long t = t0; t < t1; ++t) m = fmaxf(m, fabsf(X[(size_t)t * ldx + c]));
  atomicMax(mx + c, __float_as_uint(m));               // non-negative floats order like their bit patterns
}

// out_hi/out_lo [C][Tp] fp16 <- X [T][C] fp32 (row stride ldx); scale[c] = 2^-e with max|X[:,c]| in [2^(e-1), 2^e)
// (null mx: no scaling).  Columns t in [T, Tp) are zero.
__global__ void transpose_split_kernel(long long T, long long Tp, int C, const float* __restrict__ X, int ldx,
                                       const unsigned* __restrict__ mx, __half* __restrict__ out_hi,
                                       __half* __restrict__ out_lo, float* __restrict__ inv_scale) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32;
  const long long t0 = (long long)blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;          // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const long long t = t0 + r;
    const int c = c0 + tx;
    tile[r][tx] = (t < T && c < C) ? X[(size_t)t * ldx + c] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r;
    const long long t = t0 + tx;
    if (c < C && t < Tp) {
      float s = 1.f;
      if (mx) {
        const float m = __uint_as_float(mx[c]);
        int e = 0;
        if (m > 0.f) frexpf(m, &e);
        s = exp2f((float)-e);
        if (blockIdx.y == 0 && tx == 0) inv_scale[c] = exp2f((float)e);
      }
      const float v = tile[tx][r] * s;
      const __half h = __float2half_rn(v);
      out_hi[(size_t)c * Tp + t] = h;
      out_lo[(size_t)c * Tp + t] = __float2half_rn((v - __half2float(h)) * LO_SCALE);
    }
  }
}

// The same pass on 64 x 64 tiles with 16-byte accesses on both sides (float4 rows in, 8 halves along t out): the 32 x 32
// version above moved 4 + 2 bytes per access and ran at a third of the copy bandwidth (7.5 ms per step at B = 4096, more
// than the contraction it feeds).  Needs C % 4 == 0, ldx % 4 == 0 and a 16-byte aligned X.
__global__ void __launch_bounds__(256) transpose_split64_kernel(long long T, long long Tp, int C, const float* __restrict__ X,
                                                                int ldx, const unsigned* __restrict__ mx,
                                                                __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                                                                float* __restrict__ inv_scale) {
  __shared__ float tile[64][65];
  const int c0 = blockIdx.y * 64;
  const long long t0 = (long long)blockIdx.x * 64;
  const int tid = threadIdx.x;
  {
    const int q = tid & 15, r0 = tid >> 4;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int r = r0 + 16 * p;
      const long long t = t0 + r;
      const int c = c0 + 4 * q;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < T && c < C) v = __ldg(reinterpret_cast<const float4*>(X + (size_t)t * ldx + c));
      tile[r][4 * q] = v.x; tile[r][4 * q + 1] = v.y; tile[r][4 * q + 2] = v.z; tile[r][4 * q + 3] = v.w;
    }
  }
  __syncthreads();
  const int tch = tid & 7;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int cc = (tid >> 3) + 32 * pass, c = c0 + cc;
    if (c < C) {
      float s = 1.f;
      if (mx) {
        const float m = __uint_as_float(mx[c]);
        int e = 0;
        if (m > 0.f) frexpf(m, &e);
        s = exp2f((float)-e);
        if (blockIdx.x == 0 && tch == 0) inv_scale[c] = exp2f((float)e);
      }
      __half hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float v = tile[tch * 8 + i][cc] * s;
        hi[i] = __float2half_rn(v);
        lo[i] = __float2half_rn((v - __half2float(hi[i])) * LO_SCALE);
      }
      const size_t o = (size_t)c * Tp + t0 + tch * 8;
      *reinterpret_cast<uint4*>(out_hi + o) = *reinterpret_cast<const uint4*>(hi);
      *reinterpret_cast<uint4*>(out_lo + o) = *reinterpret_cast<const uint4*>(lo);
    }
  }
}

// ------------------------------------------------------------------------------------------------ GEMM
// map_a: 3-D {Tp, rows, 2 parts}, box {64, 32, 1}; map_b: box {64, 128, 2}.  partial [gridDim.z][M][N].
// The 4 CTAs of a cluster (the N tiles of one M tile and split) consume the same A tile: each fetches a quarter of its
// rows and TMA-multicasts it to all four, so A crosses L2 -> SM once per cluster instead of once per CTA and the four
// stay in lockstep (a ring stage is refilled only after all four have consumed it).  (Also sharing the B tile between two
// M tiles -- clusters of 4 x 2 -- was measured and is slower: the 128 CTAs of the value chain's contraction are 16 clusters
// of 8, of which a B200 holds 15 at a time, so the last cluster runs as a second wave.)
__global__ void __cluster_dims__(CLN, 1, 1) __launch_bounds__(THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int M, int N,
             int kb_total, float* __restrict__ partial) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + STAGES * STAGE_BYTES);
  // bars: full[3], empty[3], acc_full[2], acc_empty[2]
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 2 * STAGES + 4);
  const unsigned bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[STAGES]);
  const unsigned bar_acc_full = smem_u32(&bars[2 * STAGES]), bar_acc_empty = smem_u32(&bars[2 * STAGES + 2]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // this split's K blocks
  const int per = (kb_total + gridDim.z - 1) / gridDim.z;
  const int kb0 = blockIdx.z * per, kb1 = min(kb_total, kb0 + per);
  const int nkb = max(kb1 - kb0, 0);
  const int nchunk = (nkb + FLUSH_KB - 1) / FLUSH_KB;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, CLN); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_acc_full + 8 * s, 1); mbar_init(bar_acc_empty + 8 * s, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = *tmem_slot;
  const unsigned rank = cluster_rank();
  cluster_sync_all();                                       // every CTA's barriers exist before a peer multicasts into it

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        mbar_wait(bar_empty + 8 * s, ((unsigned)(i / STAGES) & 1u) ^ 1u);      // all 4 CTAs have consumed the stage
        const unsigned full = bar_full + 8 * s;
        mbar_expect_tx(full, STAGE_BYTES);
        const unsigned base = smem_u32(smem + s * STAGE_BYTES);
        const unsigned a_dst = base + rank * A_SLICE;
        const int a_row = m0 + (int)rank * A_SLICE_ROWS;
        tma_load_3d_mcast(a_dst, &map_a, (kb0 + i) * BK, a_row, 0, full, (unsigned short)0xF);
        tma_load_3d_mcast(a_dst + OP_TILE / 2, &map_a, (kb0 + i) * BK, a_row, 1, full, (unsigned short)0xF);
        tma_load_3d(base + OP_TILE, &map_b, (kb0 + i) * BK, n0, 0, full);
      }
    }
  } else if (warp == 1) {
    for (int c = 0; c < nchunk; ++c) {
      const int set = c & 1;
      if (c >= 2) mbar_wait(bar_acc_empty + 8 * set, (unsigned)((c >> 1) - 1) & 1u);     // the flush of chunk c-2 drained this set
      tc_fence_after();
      const unsigned acc_main = tmem_base + (unsigned)set * 2 * BN, acc_corr = acc_main + BN;
      const int i0 = c * FLUSH_KB, i1 = min(nkb, i0 + FLUSH_KB);
      for (int i = i0; i < i1; ++i) {
        const int s = i % STAGES;
        mbar_wait(bar_full + 8 * s, (unsigned)(i / STAGES) & 1u);
        tc_fence_after();
        const unsigned base = smem_u32(smem + s * STAGE_BYTES);
        const unsigned long long dA0 = smem_desc(base), dA1 = smem_desc(base + OP_TILE / 2);
        const unsigned long long dB0 = smem_desc(base + OP_TILE);       // 256 rows: hi, then lo' (adjacent in the stage)
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const unsigned first = (i == i0 && k == 0) ? 0u : 1u;
            // [main | correction] = A_hi x [B_hi | B_lo'] in ONE N = 256 instruction (A_hi is read from shared memory once
            // instead of twice: operand reads + TMA fill bound this kernel, chain_tc.cu BwdCfg), then correction += A_lo' x B_hi
            tc_mma(acc_main, dA0 + 2 * k, dB0 + 2 * k, IDESC2, first);
            tc_mma(acc_corr, dA1 + 2 * k, dB0 + 2 * k, IDESC, 1u);
          }
          tc_commit_mcast(bar_empty + 8 * s, (unsigned short)0xF);
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(bar_acc_full + 8 * set);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;                                   // TMEM lane quarter of this warp
    const int row = m0 + 32 * q + lane;
    float acc[BN];
#pragma unroll
    for (int i = 0; i < BN; ++i) acc[i] = 0.f;
    for (int c = 0; c < nchunk; ++c) {
      const int set = c & 1;
      mbar_wait(bar_acc_full + 8 * set, (unsigned)(c >> 1) & 1u);
      tc_fence_after();
      const unsigned tl = tmem_base + ((unsigned)(32 * q) << 16) + (unsigned)set * 2 * BN;
#pragma unroll
      for (int g = 0; g < BN / 32; ++g) {
        float mv[32], cv[32];
        tmem_ld32x2(tl + 32 * g, tl + BN + 32 * g, mv, cv);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[32 * g + i] += fmaf(cv[i], LO_INV, mv[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * set);
    }
    if (row < M) {
      float* dst = partial + ((size_t)blockIdx.z * M + row) * N + n0;
#pragma unroll
      for (int i = 0; i < BN; i += 4)
        if (n0 + i + 3 < N) *reinterpret_cast<float4*>(dst + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                       // nobody leaves while a peer may still multicast into it
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// C[m][n] = inv_scale[m] * sum_s partial[s][m][n]
__global__ void wgrad_reduce_kernel(int M, int N, int S, const float* __restrict__ partial, const float* __restrict__ inv_scale,
                                    float* __restrict__ C, int ldc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int m = (int)(i / N), n = (int)(i % N);
  float s = 0.f;
  for (int k = 0; k < S; ++k) s += partial[((size_t)k * M + m) * N + n];
  C[(size_t)m * ldc + n] = s * inv_scale[m];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(CUtensorMap* map, const void* ptr, long long Tp, int rows, int box_rows, int box_parts) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult qres;
    void* q = nullptr;
    ICRL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qres));
    ICRL_REQUIRE(q && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled unavailable");
    fn = reinterpret_cast<EncodeTiledFn>(q);
  }
  const cuuint64_t dims[3] = {(cuuint64_t)Tp, (cuuint64_t)rows, 2};
  const cuuint64_t strides[2] = {(cuuint64_t)Tp * 2, (cuuint64_t)rows * Tp * 2};
  const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, (cuuint32_t)box_parts};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    icrl_set_error("cuTensorMapEncodeTiled (wgrad) failed with CUresult %d (Tp %lld, rows %d)", (int)r, Tp, rows);
    return ICRL_ERR_CUDA;
  }
  return ICRL_OK;
}

inline long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

}  // namespace

// workspace bytes for icrl_wgrad_tc_impl
size_t icrl_wgrad_tc_ws_bytes_impl(int M, int N, long long T, int splits) {
  const long long Tp = round_up(T, 64);
  return (size_t)4 * Tp * (M + N)            // hi + lo' of both operands, fp16
         + (size_t)splits * M * N * 4        // partial tiles
         + (size_t)M * 8 + 1024;             // column maxima + inverse scales (+ alignment)
}

// C [M][ldc] = A^T B with A [T][lda] (M columns used, scaled per column), B [T][ldb] (N columns used).
// The B operand's pre-pass alone (transpose + split into the workspace of the same M, N, T, splits): B is often final long
// before A (the value chain's h stash exists after the forward), so a caller can run this on another stream beside
// whatever produces A and then pass b_packed = 1 to the contraction.
int icrl_wgrad_tc_pack_b_impl(cudaStream_t st, int M, int N, long long T, const float* B, int ldb, void* ws, size_t ws_bytes,
                              int splits) {
  ICRL_REQUIRE(M % BM == 0 && N % (BN * CLN) == 0 && T > 0 && splits >= 1, "wgrad_tc needs M a multiple of 128, N of 512");
  ICRL_REQUIRE(ws && ws_bytes >= icrl_wgrad_tc_ws_bytes_impl(M, N, T, splits), "wgrad_tc workspace too small");
  ICRL_REQUIRE(ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0, "wgrad_tc_pack_b needs 16-byte aligned rows");
  const long long Tp = round_up(T, 64);
  __half* b_pk = reinterpret_cast<__half*>(ws) + (size_t)2 * M * Tp;
  dim3 gb((unsigned)(Tp / 64), N / 64);
  transpose_split64_kernel<<<gb, 256, 0, st>>>(T, Tp, N, B, ldb, nullptr, b_pk, b_pk + (size_t)N * Tp, nullptr);
  ICRL_LAUNCH_CHECK();
  return ICRL_OK;
}

// colmax (nullable): M words holding the bit patterns of max_t |A[t][m]| when the caller has them already (the gate-table
// scatter reads all of A and returns them: icrl_scatter_add_stream); otherwise a pre-pass over A computes them.
// b_packed: the B operand already lies in the workspace (icrl_wgrad_tc_pack_b_impl).
int icrl_wgrad_tc_impl(cudaStream_t st, int M, int N, long long T, const float* A, int lda, const float* B, int ldb,
                       float* C, int ldc, void* ws, size_t ws_bytes, int splits, const unsigned* colmax, int b_packed) {
  ICRL_REQUIRE(M % BM == 0 && N % (BN * CLN) == 0 && T > 0 && splits >= 1, "wgrad_tc needs M a multiple of 128, N of 512");
  ICRL_REQUIRE(ws && ws_bytes >= icrl_wgrad_tc_ws_bytes_impl(M, N, T, splits), "wgrad_tc workspace too small");
  static bool attr_set = false;
  if (!attr_set) {
    ICRL_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  const long long Tp = round_up(T, 64);
  __half* a_pk = reinterpret_cast<__half*>(ws);                          // [2][M][Tp]
  __half* b_pk = a_pk + (size_t)2 * M * Tp;                             // [2][N][Tp]
  float* partial = reinterpret_cast<float*>(b_pk + (size_t)2 * N * Tp);  // [splits][M][N]
  unsigned* mx = reinterpret_cast<unsigned*>(partial + (size_t)splits * M * N);
  float* inv_scale = reinterpret_cast<float*>(mx + M);      // (before mx may be redirected to the caller's maxima)
  if (colmax) {
    mx = const_cast<unsigned*>(colmax);
  } else {
    ICRL_CUDA(cudaMemsetAsync(mx, 0, (size_t)M * sizeof(unsigned), st));
    dim3 grid(icrl_cdiv(M, 128), (unsigned)min((long long)592, (T + 255) / 256));
    col_absmax_kernel<<<grid, 128, 0, st>>>(T, M, A, lda, mx);
    ICRL_LAUNCH_CHECK();
  }
  const bool vec = lda % 4 == 0 && ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0;
  if (vec) {
    dim3 ga((unsigned)(Tp / 64), M / 64), gb((unsigned)(Tp / 64), N / 64);
    transpose_split64_kernel<<<ga, 256, 0, st>>>(T, Tp, M, A, lda, mx, a_pk, a_pk + (size_t)M * Tp, inv_scale);
    ICRL_LAUNCH_CHECK();
    if (!b_packed) {
      transpose_split64_kernel<<<gb, 256, 0, st>>>(T, Tp, N, B, ldb, nullptr, b_pk, b_pk + (size_t)N * Tp, nullptr);
      ICRL_LAUNCH_CHECK();
    }
  } else {
    dim3 blk(32, 8);
    dim3 ga(icrl_cdiv(M, 32), (unsigned)(Tp / 32)), gb(icrl_cdiv(N, 32), (unsigned)(Tp / 32));
    transpose_split_kernel<<<ga, blk, 0, st>>>(T, Tp, M, A, lda, mx, a_pk, a_pk + (size_t)M * Tp, inv_scale);
    ICRL_LAUNCH_CHECK();
    if (!b_packed) {
      transpose_split_kernel<<<gb, blk, 0, st>>>(T, Tp, N, B, ldb, nullptr, b_pk, b_pk + (size_t)N * Tp, nullptr);
      ICRL_LAUNCH_CHECK();
    }
  }
  CUtensorMap ma, mb;
  int rc;
  if ((rc = make_map(&ma, a_pk, Tp, M, A_SLICE_ROWS, 1))) return rc;
  if ((rc = make_map(&mb, b_pk, Tp, N, BN, 2))) return rc;
  dim3 grid(N / BN, M / BM, splits);
  wgrad_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(ma, mb, M, N, (int)(Tp / BK), partial);
  ICRL_LAUNCH_CHECK();
  wgrad_reduce_kernel<<<icrl_cdiv((long long)M * N, 256), 256, 0, st>>>(M, N, splits, partial, inv_scale, C, ldc);
  ICRL_LAUNCH_CHECK();
  return ICRL_OK;
}
