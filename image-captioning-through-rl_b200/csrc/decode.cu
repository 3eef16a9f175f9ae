// Persistent fused policy decode kernel for sm_100a (north_star subsystem 1): ONE launch runs every
// timestep of the rollout -- LSTM gate GEMM on tcgen05, cell / hidden update, vocab projection on tcgen05,
// softmax, numpy-semantics inverse-CDF sampling (or greedy argmax / forced tokens) and log-prob --
// replacing PolicyNetwork.forward re-runs + F.softmax + np.random.choice + gather/log of the reference
// rollout body (models.py:71-84, 286; trainers.py:441-458; greedy trainers.py:57-70).
//
// Decomposition.  Rows of the batch are independent in the policy, so a CLUSTER of 8 CTAs owns 128 rows for
// the whole rollout and never talks to another cluster (no grid-wide sync).  Inside a cluster CTA r owns
// hidden units [64r, 64r+64) (256 gate columns: i,f,g,o of those units, W_hh rows permuted at pack time) and
// vocab columns [128r, 128r+128) (V padded to 1024, padded logits are -inf and can never be sampled).
//
// Precision.  Token ids must match the fp32 reference bit for bit (single-pass bf16/TF32 flips
// tokens), so every fp32 operand is split into TWO fp16 parts, x = hi + lo'/2048 with hi = fp16(x),
// lo' = fp16((x - hi) * 2048) (22 mantissa bits; the scaling keeps lo' out of the fp16 subnormal range), and
// three products are accumulated in f32 in tensor memory: hi*hi into a "main" accumulator, hi*lo' + lo'*hi
// into a "correction" accumulator that the epilogue scales by 2^-11.  Keeping the correction terms out of the
// full-magnitude accumulator matters: the tensor core truncates on every accumulate (gemm_tc.cu header,
// scripts/tc_accuracy_probe.py).  The vocab GEMM further splits the main term over two accumulators by
// K-block parity.  Operand-split error 7e-8 of max|C| (emulated), i.e. below fp32 rounding of the dot itself.
//
// Per CTA: warp 0 = TMA producer (2-stage ring of 64-wide K blocks: A hi/lo 2 x 16 KB, B hi/lo 2 x 32 KB),
// warp 1 = tcgen05.mma issuer (M=128, N=256 gates / N=128 vocab, 12 MMAs per stage), warps 2..9 = epilogue
// (TMEM lane quarter = warp % 4, two warps per quarter each taking half of the columns).
//
// Per cell step j:   G-MMA -> G-epilogue (gate table gather by token + activations + c,h update; writes the
// fp32 stash for backward and the fp16 split of h for the next GEMMs) -> cluster barrier (all 512 h columns
// of the 128 rows now exist) -> V-MMA -> V-epilogue (logits stash, then softmax / sampling across the 8 CTAs
// through DSMEM: each thread st.async's its 64-column partial {max,argmax} / sum / f64 cdf total / count to
// all 8 CTAs, completion on mbarriers) -- the V-epilogue's math overlaps the next step's G-MMA, which only
// needs h.
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

constexpr int H = ICRL_H;
constexpr int CL = 8;                       // CTAs per cluster
constexpr int BM = 128;                     // rows per cluster
constexpr int GN = 4 * H / CL;              // 256 gate columns per CTA
constexpr int UN = H / CL;                  // 64 hidden units per CTA
constexpr int VN = 128;                     // vocab columns per CTA
constexpr int VPAD = CL * VN;               // 1024
constexpr int BK = 64, KB = H / BK, UMMA_K = 16, STAGES = 2;
constexpr int A_TILE = BM * BK * 2;         // 16 KB
constexpr int BG_TILE = GN * BK * 2;        // 32 KB
constexpr int BV_TILE = VN * BK * 2;        // 16 KB
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * BG_TILE;      // 96 KB
constexpr int XCHG_SLOTS = 2 * CL;          // partials per row: 8 CTAs x 2 column halves
constexpr int XCHG_BYTES = XCHG_SLOTS * BM * 8;            // 16 KB per exchange buffer
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * EPI_WARPS;               // 320
constexpr int NBARS = 2 * STAGES + 2 + 2;   // full[2], empty[2], acc_full, acc_empty, xchg[2]
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * XCHG_BYTES + 256 + 1024;
constexpr int TMEM_COLS = 512;
constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f;

struct DecodeArgs {
  int B, V, p0, S, greedy;
  const float* table;        // [V][2048]  W_ih E[v] + b_ih + b_hh
  const float* b_v;          // [V]
  const double* uniforms;    // [S][B] or null
  const long long* forced;   // [B][S] or null
  int* tokcm;                // [(p0+S)][B]
  long long* tokens_out;     // [B][S]
  float* logp;               // [B][S]
  float* Hs;                 // [(n_cell+1)][B][512]
  float* Cs;                 // [(n_cell+1)][B][512]
  float* Gs;                 // [n_cell][B][2048] activated i,f,g,o   (nullable: inference)
  float* logits;             // [S][B][V]                             (nullable: inference)
  float* last_logits;        // [B][V] logits of the last step        (nullable)
  __half* hparts;            // [2 buffers][2 parts][B][512]
};

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
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
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
// K-major, 128B-swizzle operand descriptor (see gemm_tc.cu)
__device__ __forceinline__ unsigned long long smem_desc(unsigned addr) {
  return (unsigned long long)((addr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// c_format F32 [4,6) = 1, a/b_format F16 = 0, K-major, n_dim = N>>3 [17,23), m_dim = M>>4 [24,29)
constexpr unsigned idesc_f16(int n) { return (1u << 4) | ((unsigned)(n >> 3) << 17) | ((unsigned)(BM >> 4) << 24); }

// TMEM -> registers.  The loads and their wait live in ONE asm statement so that no consumer of the
// destination registers can be scheduled before tcgen05.wait::ld.
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
__device__ __forceinline__ void tmem_ld8x3(unsigned ta, unsigned tb, unsigned tc, float* a, float* b, float* c) {
  unsigned r[24];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%24];\n"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8, %9, %10, %11, %12, %13, %14, %15}, [%25];\n"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16, %17, %18, %19, %20, %21, %22, %23}, [%26];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23])
      : "r"(ta), "r"(tb), "r"(tc)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[8 + i]); c[i] = __uint_as_float(r[16 + i]); }
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_b64(unsigned raddr, unsigned long long v, unsigned rbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(raddr), "l"(v), "r"(rbar)
               : "memory");
}

// Sends this thread's partial to slot [slot][row] of the exchange buffer of every CTA of the cluster.
__device__ __forceinline__ void xchg_send(unsigned buf, unsigned bar, int slot, int row, unsigned long long v) {
  const unsigned a = buf + (unsigned)(slot * BM + row) * 8u;
#pragma unroll
  for (unsigned dst = 0; dst < (unsigned)CL; ++dst) st_async_b64(mapa(a, dst), v, mapa(bar, dst));
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1)
policy_decode_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_whh,
                     const __grid_constant__ CUtensorMap map_wv, const DecodeArgs p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* xbuf = smem + STAGES * STAGE_BYTES;                       // [2][XCHG_SLOTS][BM] 8-byte slots
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(xbuf + 2 * XCHG_BYTES);
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + NBARS);
  const unsigned bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[STAGES]);
  const unsigned bar_acc_full = smem_u32(&bars[2 * STAGES]), bar_acc_empty = smem_u32(&bars[2 * STAGES + 1]);
  const unsigned bar_x = smem_u32(&bars[2 * STAGES + 2]);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int m0 = (blockIdx.x / CL) * BM;
  const int B = p.B, V = p.V;
  const int n_cell = p.p0 - 1 + p.S;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_acc_full, 1);
    mbar_init(bar_acc_empty, EPI_WARPS);
    mbar_init(bar_x, 1);
    mbar_init(bar_x + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_h) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_whh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wv) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = *tmem_slot;
  cluster_arrive();                      // every CTA's mbarriers are initialised before anyone sends to them
  cluster_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    unsigned it = 0;
    for (int j = 0; j < n_cell; ++j) {
      if (lane == 0) {
        const int arow = ((j & 1) * 2) * B + m0;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(bar_empty + 8 * s, ((it / STAGES) & 1u) ^ 1u);
          const unsigned full = bar_full + 8 * s;
          mbar_expect_tx(full, 2 * A_TILE + 2 * BG_TILE);
          const unsigned base = smem_u32(smem + s * STAGE_BYTES);
          tma_load_2d(base, &map_h, kb * BK, arow, full);
          tma_load_2d(base + A_TILE, &map_h, kb * BK, arow + B, full);
          tma_load_2d(base + 2 * A_TILE, &map_whh, kb * BK, (int)rank * GN, full);
          tma_load_2d(base + 2 * A_TILE + BG_TILE, &map_whh, kb * BK, 4 * H + (int)rank * GN, full);
        }
      }
      __syncwarp();
      cluster_arrive();
      cluster_wait();                    // h_j of all 8 CTAs is in global memory
      if (lane == 0) {
        fence_proxy_async();
        if (j >= p.p0 - 1) {
          const int arow = (((j + 1) & 1) * 2) * B + m0;
          for (int kb = 0; kb < KB; ++kb, ++it) {
            const int s = it % STAGES;
            mbar_wait(bar_empty + 8 * s, ((it / STAGES) & 1u) ^ 1u);
            const unsigned full = bar_full + 8 * s;
            mbar_expect_tx(full, 2 * A_TILE + 2 * BV_TILE);
            const unsigned base = smem_u32(smem + s * STAGE_BYTES);
            tma_load_2d(base, &map_h, kb * BK, arow, full);
            tma_load_2d(base + A_TILE, &map_h, kb * BK, arow + B, full);
            tma_load_2d(base + 2 * A_TILE, &map_wv, kb * BK, (int)rank * VN, full);
            tma_load_2d(base + 2 * A_TILE + BG_TILE, &map_wv, kb * BK, VPAD + (int)rank * VN, full);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    unsigned it = 0, nphase = 0;
    for (int j = 0; j < n_cell; ++j) {
      if (lane == 0) {
        if (nphase > 0) mbar_wait(bar_acc_empty, (nphase - 1) & 1u);       // previous epilogue drained TMEM
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(bar_full + 8 * s, (it / STAGES) & 1u);
          tc_fence_after();
          const unsigned base = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
          for (int pair = 0; pair < 3; ++pair) {                           // (hi,hi) (hi,lo') (lo',hi)
            const unsigned a = base + (pair == 2 ? A_TILE : 0);
            const unsigned b = base + 2 * A_TILE + (pair == 1 ? BG_TILE : 0);
            const unsigned acc = pair == 0 ? 0u : (unsigned)GN;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const unsigned accumulate = pair == 0 ? ((kb | k) != 0) : !(kb == 0 && pair == 1 && k == 0);
              tc_mma(tmem_base + acc, smem_desc(a + k * UMMA_K * 2), smem_desc(b + k * UMMA_K * 2), idesc_f16(GN), accumulate);
            }
          }
          tc_commit(bar_empty + 8 * s);
        }
        tc_commit(bar_acc_full);
        ++nphase;
      }
      __syncwarp();
      cluster_arrive();
      cluster_wait();
      if (lane == 0 && j >= p.p0 - 1) {
        mbar_wait(bar_acc_empty, (nphase - 1) & 1u);
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(bar_full + 8 * s, (it / STAGES) & 1u);
          tc_fence_after();
          const unsigned base = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
          for (int pair = 0; pair < 3; ++pair) {
            const unsigned a = base + (pair == 2 ? A_TILE : 0);
            const unsigned b = base + 2 * A_TILE + (pair == 1 ? BG_TILE : 0);
            const unsigned acc = pair == 0 ? (unsigned)(kb & 1) * VN : 2u * VN;   // main even/odd K blocks, correction
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const unsigned accumulate = pair == 0 ? (kb >= 2 || k != 0) : !(kb == 0 && pair == 1 && k == 0);
              tc_mma(tmem_base + acc, smem_desc(a + k * UMMA_K * 2), smem_desc(b + k * UMMA_K * 2), idesc_f16(VN), accumulate);
            }
          }
          tc_commit(bar_empty + 8 * s);
        }
        tc_commit(bar_acc_full);
        ++nphase;
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;
    const int q = warp & 3;                    // TMEM lane quarter this warp may read
    const int ch = ew >> 2;                    // column half
    const int rl = 32 * q + lane;              // row inside the cluster tile
    const int row = m0 + rl;
    const bool valid = row < B;
    const unsigned tq = tmem_base + ((unsigned)(32 * q) << 16);
    const int slot = (int)rank * 2 + ch;
    unsigned ephase = 0;                       // accumulator phases consumed
    unsigned xph0 = 0, xph1 = 0;               // exchange barrier phases
    int tok = valid ? p.tokcm[row] : 0;        // token consumed by cell 0 (prefix column 0)

    for (int j = 0; j < n_cell; ++j) {
      const size_t BH = (size_t)B * H;
      // ---------------- gate epilogue: cell update for this thread's 32 hidden units
      mbar_wait(bar_acc_full, ephase & 1u);
      ++ephase;
      tc_fence_after();
      const float* trow = p.table + (size_t)tok * (4 * H) + rank * UN + 32 * ch;
      const size_t hoff = (size_t)row * H + rank * UN + 32 * ch;
      __half* hp_hi = p.hparts + ((size_t)(((j + 1) & 1) * 2) * B + row) * H + rank * UN + 32 * ch;
      __half* hp_lo = hp_hi + BH;
#pragma unroll 1
      for (int c8 = 0; c8 < 4; ++c8) {
        float acc[4][8], cor[4][8];
#pragma unroll
        for (int g = 0; g < 4; ++g)
          tmem_ld8x2(tq + (unsigned)(g * UN + 32 * ch + 8 * c8), tq + (unsigned)(GN + g * UN + 32 * ch + 8 * c8), acc[g], cor[g]);
        if (valid) {
          float tb[4][8], cp[8];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 t0 = *reinterpret_cast<const float4*>(trow + g * H + 8 * c8);
            const float4 t1 = *reinterpret_cast<const float4*>(trow + g * H + 8 * c8 + 4);
            tb[g][0] = t0.x; tb[g][1] = t0.y; tb[g][2] = t0.z; tb[g][3] = t0.w;
            tb[g][4] = t1.x; tb[g][5] = t1.y; tb[g][6] = t1.z; tb[g][7] = t1.w;
          }
          {
            const float4 c0 = *reinterpret_cast<const float4*>(p.Cs + (size_t)j * BH + hoff + 8 * c8);
            const float4 c1 = *reinterpret_cast<const float4*>(p.Cs + (size_t)j * BH + hoff + 8 * c8 + 4);
            cp[0] = c0.x; cp[1] = c0.y; cp[2] = c0.z; cp[3] = c0.w; cp[4] = c1.x; cp[5] = c1.y; cp[6] = c1.z; cp[7] = c1.w;
          }
          float gi[8], gf[8], gg[8], go[8], cn[8], hn[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            gi[i] = sigmoidf_acc(fmaf(cor[0][i], LO_INV, acc[0][i]) + tb[0][i]);
            gf[i] = sigmoidf_acc(fmaf(cor[1][i], LO_INV, acc[1][i]) + tb[1][i]);
            gg[i] = tanhf(fmaf(cor[2][i], LO_INV, acc[2][i]) + tb[2][i]);
            go[i] = sigmoidf_acc(fmaf(cor[3][i], LO_INV, acc[3][i]) + tb[3][i]);
            cn[i] = gf[i] * cp[i] + gi[i] * gg[i];
            hn[i] = go[i] * tanhf(cn[i]);
          }
          if (p.Gs) {
            float* gs = p.Gs + ((size_t)j * B + row) * (4 * H) + rank * UN + 32 * ch + 8 * c8;
            *reinterpret_cast<float4*>(gs) = make_float4(gi[0], gi[1], gi[2], gi[3]);
            *reinterpret_cast<float4*>(gs + 4) = make_float4(gi[4], gi[5], gi[6], gi[7]);
            *reinterpret_cast<float4*>(gs + H) = make_float4(gf[0], gf[1], gf[2], gf[3]);
            *reinterpret_cast<float4*>(gs + H + 4) = make_float4(gf[4], gf[5], gf[6], gf[7]);
            *reinterpret_cast<float4*>(gs + 2 * H) = make_float4(gg[0], gg[1], gg[2], gg[3]);
            *reinterpret_cast<float4*>(gs + 2 * H + 4) = make_float4(gg[4], gg[5], gg[6], gg[7]);
            *reinterpret_cast<float4*>(gs + 3 * H) = make_float4(go[0], go[1], go[2], go[3]);
            *reinterpret_cast<float4*>(gs + 3 * H + 4) = make_float4(go[4], go[5], go[6], go[7]);
          }
          float* cs = p.Cs + (size_t)(j + 1) * BH + hoff + 8 * c8;
          float* hs = p.Hs + (size_t)(j + 1) * BH + hoff + 8 * c8;
          *reinterpret_cast<float4*>(cs) = make_float4(cn[0], cn[1], cn[2], cn[3]);
          *reinterpret_cast<float4*>(cs + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
          *reinterpret_cast<float4*>(hs) = make_float4(hn[0], hn[1], hn[2], hn[3]);
          *reinterpret_cast<float4*>(hs + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
          __half2 hi2[4], lo2[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const __half h0 = __float2half_rn(hn[2 * i]), h1 = __float2half_rn(hn[2 * i + 1]);
            hi2[i] = __halves2half2(h0, h1);
            lo2[i] = __halves2half2(__float2half_rn((hn[2 * i] - __half2float(h0)) * LO_SCALE),
                                    __float2half_rn((hn[2 * i + 1] - __half2float(h1)) * LO_SCALE));
          }
          *reinterpret_cast<uint4*>(hp_hi + 8 * c8) = *reinterpret_cast<const uint4*>(hi2);
          *reinterpret_cast<uint4*>(hp_lo + 8 * c8) = *reinterpret_cast<const uint4*>(lo2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty);
      __threadfence();
      fence_proxy_async();                     // generic-proxy writes of the h split -> visible to the peers' TMA loads
      cluster_arrive();
      cluster_wait();

      const int s = j - (p.p0 - 1);
      if (s < 0) {                             // teacher-forced prefix cell: next token comes from the caption
        tok = valid ? p.tokcm[(size_t)(j + 1) * B + row] : 0;
        continue;
      }
      // ---------------- vocab epilogue, part 1: logits of this thread's 64 columns into registers
      mbar_wait(bar_acc_full, ephase & 1u);
      ++ephase;
      tc_fence_after();
      float x[64];
      const int v0 = (int)rank * VN + 64 * ch;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        float m0v[8], m1v[8], cv[8];
        tmem_ld8x3(tq + (unsigned)(64 * ch + 8 * c8), tq + (unsigned)(VN + 64 * ch + 8 * c8),
                   tq + (unsigned)(2 * VN + 64 * ch + 8 * c8), m0v, m1v, cv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int v = v0 + 8 * c8 + i;
          x[8 * c8 + i] = v < V ? fmaf(cv[i], LO_INV, m0v[i] + m1v[i]) + __ldg(p.b_v + v) : -INFINITY;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty);          // the next gate GEMM may overwrite TMEM now
      if (valid) {
        float* lg = p.logits ? p.logits + ((size_t)s * B + row) * V + v0 : nullptr;
        float* ll = (p.last_logits && s == p.S - 1) ? p.last_logits + (size_t)row * V + v0 : nullptr;
#pragma unroll
        for (int c4 = 0; c4 < 16; ++c4) {
          if (v0 + 4 * c4 + 3 < V) {                       // V % 4 == 0 is required by the host wrapper
            const float4 o = make_float4(x[4 * c4], x[4 * c4 + 1], x[4 * c4 + 2], x[4 * c4 + 3]);
            if (lg) *reinterpret_cast<float4*>(lg + 4 * c4) = o;
            if (ll) *reinterpret_cast<float4*>(ll + 4 * c4) = o;
          }
        }
      }
      // ---------------- part 2: softmax + sampling across the cluster (trainers.py:444-458)
      // E1: row max and first argmax
      float mx = -INFINITY;
      int amax = VPAD;
#pragma unroll
      for (int i = 0; i < 64; ++i)
        if (x[i] > mx) { mx = x[i]; amax = v0 + i; }
      if (ew == 0 && lane == 0) mbar_expect_tx(bar_x, XCHG_BYTES);
      xchg_send(smem_u32(xbuf), bar_x, slot, rl, ((unsigned long long)(unsigned)amax << 32) | __float_as_uint(mx));
      mbar_wait(bar_x, xph0 & 1u);
      ++xph0;
      {
        const unsigned long long* xs = reinterpret_cast<const unsigned long long*>(xbuf) + rl;
        mx = -INFINITY;
        amax = VPAD;
#pragma unroll
        for (int k = 0; k < XCHG_SLOTS; ++k) {            // slots are in ascending column order: strict > keeps the first
          const unsigned long long w = xs[k * BM];
          const float m = __uint_as_float((unsigned)w);
          if (m > mx) { mx = m; amax = (int)(w >> 32); }
        }
      }
      // E2: sum of exp(x - max) in f32 (F.softmax)
      float lsum = 0.f;
#pragma unroll
      for (int i = 0; i < 64; ++i) { x[i] = expf(x[i] - mx); lsum += x[i]; }
      if (ew == 0 && lane == 0) mbar_expect_tx(bar_x + 8, XCHG_BYTES);
      xchg_send(smem_u32(xbuf + XCHG_BYTES), bar_x + 8, slot, rl, (unsigned long long)__float_as_uint(lsum));
      mbar_wait(bar_x + 8, xph1 & 1u);
      ++xph1;
      float tot = 0.f;
      {
        const unsigned long long* xs = reinterpret_cast<const unsigned long long*>(xbuf + XCHG_BYTES) + rl;
#pragma unroll
        for (int k = 0; k < XCHG_SLOTS; ++k) tot += __uint_as_float((unsigned)xs[k * BM]);
      }
      const float inv = 1.0f / tot;
      int a;
      if (p.forced) {
        a = valid ? (int)p.forced[(size_t)row * p.S + s] : 0;
      } else if (p.greedy) {
        a = min(amax, V - 1);
      } else {
        // E3: float64 cdf (np.random.choice: cdf = cumsum(float64(p)); cdf /= cdf[-1]; searchsorted(u, 'right'))
        double loc = 0.0;
#pragma unroll
        for (int i = 0; i < 64; ++i) loc += (double)(x[i] * inv);
        if (ew == 0 && lane == 0) mbar_expect_tx(bar_x, XCHG_BYTES);
        xchg_send(smem_u32(xbuf), bar_x, slot, rl, (unsigned long long)__double_as_longlong(loc));
        mbar_wait(bar_x, xph0 & 1u);
        ++xph0;
        double pre = 0.0, total = 0.0;
        {
          const unsigned long long* xs = reinterpret_cast<const unsigned long long*>(xbuf) + rl;
#pragma unroll
          for (int k = 0; k < XCHG_SLOTS; ++k) {
            const double d = __longlong_as_double((long long)xs[k * BM]);
            if (k < slot) pre += d;
            total += d;
          }
        }
        const double u = valid ? p.uniforms[(size_t)s * B + row] : 0.0;
        double run = pre;
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          run += (double)(x[i] * inv);
          cnt += (v0 + i < V && run / total <= u) ? 1 : 0;
        }
        // E4: count of cdf entries <= u
        if (ew == 0 && lane == 0) mbar_expect_tx(bar_x + 8, XCHG_BYTES);
        xchg_send(smem_u32(xbuf + XCHG_BYTES), bar_x + 8, slot, rl, (unsigned long long)(unsigned)cnt);
        mbar_wait(bar_x + 8, xph1 & 1u);
        ++xph1;
        int total_cnt = 0;
        {
          const unsigned long long* xs = reinterpret_cast<const unsigned long long*>(xbuf + XCHG_BYTES) + rl;
#pragma unroll
          for (int k = 0; k < XCHG_SLOTS; ++k) total_cnt += (int)(unsigned)xs[k * BM];
        }
        a = min(total_cnt, V - 1);
      }
      if (valid) {
        if (a >= v0 && a < v0 + 64) {                      // the thread that holds column a: log p[a] (trainers.py:458)
          float ea = 0.f;
#pragma unroll
          for (int i = 0; i < 64; ++i) ea = (a - v0 == i) ? x[i] : ea;
          p.logp[(size_t)row * p.S + s] = logf(ea * inv);
        }
        if (slot == 0) {
          p.tokens_out[(size_t)row * p.S + s] = a;
          p.tokcm[(size_t)(p.p0 + s) * B + row] = a;
        }
      }
      tok = a;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_arrive();                          // nobody leaves while a peer may still write into its shared memory
  cluster_wait();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// fp32 -> {hi, lo'} fp16 split (see header).  dst_hi/dst_lo index = src index.
__global__ void split_f16x2_kernel(long long n, const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn((v - __half2float(h)) * LO_SCALE);
  }
}

// W_hh [2048][512] -> whh_pk [2][2048][512] with row (rank*256 + g*64 + jj) = W_hh row (g*512 + rank*64 + jj);
// W_v [V][512] -> wv_pk [2][1024][512], rows >= V zero.
__global__ void pack_decode_weights_kernel(int V, const float* __restrict__ W_hh, const float* __restrict__ W_v,
                                           __half* __restrict__ whh_pk, __half* __restrict__ wv_pk) {
  const long long n_hh = (long long)4 * H * H, n_v = (long long)VPAD * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_hh + n_v; i += (long long)gridDim.x * blockDim.x) {
    float v;
    __half *hi, *lo;
    if (i < n_hh) {
      const int prow = (int)(i / H), k = (int)(i % H);
      const int r = prow / GN, g = (prow % GN) / UN, jj = prow % UN;
      v = W_hh[(size_t)(g * H + r * UN + jj) * H + k];
      hi = whh_pk + i;
      lo = whh_pk + n_hh + i;
    } else {
      const long long e = i - n_hh;
      const int vrow = (int)(e / H);
      v = vrow < V ? W_v[e] : 0.f;
      hi = wv_pk + e;
      lo = wv_pk + n_v + e;
    }
    const __half h = __float2half_rn(v);
    *hi = h;
    *lo = __float2half_rn((v - __half2float(h)) * LO_SCALE);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map_f16(CUtensorMap* map, const void* ptr, long long rows, int box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult qres;
    void* q = nullptr;
    ICRL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qres));
    ICRL_REQUIRE(q && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled unavailable");
    fn = reinterpret_cast<EncodeTiledFn>(q);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)H * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    icrl_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %lld)", (int)r, rows);
    return ICRL_ERR_CUDA;
  }
  return ICRL_OK;
}

}  // namespace

size_t icrl_decode_weight_halves_impl() { return (size_t)2 * (4 * H + VPAD) * H; }

int icrl_pack_decode_weights_impl(cudaStream_t st, int V, const float* W_hh, const float* W_v, void* packed) {
  ICRL_REQUIRE(V > 0 && V <= VPAD && V % 4 == 0, "fused decode kernel needs V <= 1024 and V % 4 == 0");
  __half* whh_pk = reinterpret_cast<__half*>(packed);
  __half* wv_pk = whh_pk + (size_t)2 * 4 * H * H;
  pack_decode_weights_kernel<<<148 * 4, 256, 0, st>>>(V, W_hh, W_v, whh_pk, wv_pk);
  ICRL_LAUNCH_CHECK();
  return ICRL_OK;
}

int icrl_split_f16x2_impl(cudaStream_t st, long long n, const float* x, void* hi, void* lo) {
  const int blocks = (int)min((long long)148 * 8, (n + 255) / 256);
  split_f16x2_kernel<<<blocks, 256, 0, st>>>(n, x, (__half*)hi, (__half*)lo);
  ICRL_LAUNCH_CHECK();
  return ICRL_OK;
}

// Hs[0] = h0, Cs[0] = c0 must be set by the caller; hparts [2][2][B][512] fp16 scratch.
int icrl_policy_decode_impl(cudaStream_t st, int B, int V, int p0, int S, int greedy, const float* table,
                            const void* packed, const float* b_v, const double* uniforms, const long long* forced,
                            int* tokcm, long long* tokens_out, float* logp, float* Hs, float* Cs, float* Gs,
                            float* logits, float* last_logits, void* hparts) {
  ICRL_REQUIRE(B > 0 && p0 >= 1 && S >= 1, "bad rollout shape");
  ICRL_REQUIRE(V > 0 && V <= VPAD && V % 4 == 0, "fused decode kernel needs V <= 1024 and V % 4 == 0");
  ICRL_REQUIRE(greedy || uniforms || forced, "sampling needs uniforms");
  static bool attr_set = false;
  if (!attr_set) {
    ICRL_CUDA(cudaFuncSetAttribute(policy_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  const __half* whh_pk = reinterpret_cast<const __half*>(packed);
  const __half* wv_pk = whh_pk + (size_t)2 * 4 * H * H;
  // initial h split into buffer 0
  {
    int rc = icrl_split_f16x2_impl(st, (long long)B * H, Hs, hparts, reinterpret_cast<__half*>(hparts) + (size_t)B * H);
    if (rc) return rc;
  }
  CUtensorMap mh, mw, mv;
  int rc;
  if ((rc = make_map_f16(&mh, hparts, (long long)4 * B, BM))) return rc;
  if ((rc = make_map_f16(&mw, whh_pk, (long long)2 * 4 * H, GN))) return rc;
  if ((rc = make_map_f16(&mv, wv_pk, (long long)2 * VPAD, VN))) return rc;
  DecodeArgs a;
  a.B = B; a.V = V; a.p0 = p0; a.S = S; a.greedy = greedy;
  a.table = table; a.b_v = b_v; a.uniforms = (greedy || forced) ? nullptr : uniforms; a.forced = forced;
  a.tokcm = tokcm; a.tokens_out = tokens_out; a.logp = logp; a.Hs = Hs; a.Cs = Cs; a.Gs = Gs; a.logits = logits;
  a.last_logits = last_logits; a.hparts = reinterpret_cast<__half*>(hparts);
  const int n_mtiles = icrl_cdiv(B, BM);
  policy_decode_kernel<<<dim3(CL * n_mtiles), dim3(THREADS), SMEM_BYTES, st>>>(mh, mw, mv, a);
  ICRL_LAUNCH_CHECK();
  return ICRL_OK;
}
