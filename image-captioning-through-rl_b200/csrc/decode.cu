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
// Per CTA (384 threads): warp 0 = TMA producer (4-stage ring of 32-wide K blocks: A hi/lo 2 x 8 KB, B hi/lo
// 2 x 16 KB), warp 1 = tcgen05.mma issuer (M=128, N=256 gates / N=128 vocab, 6 MMAs per stage), warps 4..11 =
// epilogue (TMEM lane quarter = warp % 4, two warps per quarter each taking half of the columns).  Warpgroup 0
// gives its registers to the two epilogue warpgroups (setmaxnreg 40 / 232).
//
// Global-memory access of the epilogues.  tcgen05.ld hands a thread one accumulator ROW, and a warp-wide
// float4 access with one row per lane touches 32 cache lines (32 L1 tag cycles): the first version of the gate
// epilogue spent 42 K of its 100 K cycles per step there.  So the gate epilogue stages through shared memory
// (the TMA ring is idle between the gate GEMM and the cluster barrier): cp.async gathers the gate-table rows
// and c_{t-1} as [32 rows][32 units] tiles with 8 lanes per row (4 lines per access), the cell update runs
// thread-per-row on the swizzled tile in place, and the stash goes out the same coalesced way.  The logits
// stash is staged through the second exchange buffer 16 columns at a time.
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
constexpr int BK = 32, KB = H / BK, UMMA_K = 16, STAGES = 4;   // 64-byte rows (SWIZZLE_64B); 4 stages cover the L2->SMEM latency
constexpr int A_TILE = BM * BK * 2;         // 8 KB
constexpr int BG_TILE = GN * BK * 2;        // 16 KB
constexpr int BV_TILE = VN * BK * 2;        // 8 KB
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * BG_TILE;      // 48 KB
constexpr int XCHG_SLOTS = 2 * CL;          // partials per row: 8 CTAs x 2 column halves
constexpr int XCHG_BYTES = XCHG_SLOTS * BM * 8;            // 16 KB per exchange buffer
constexpr int EPI_WARPS = 8;
constexpr int EPI_WARP0 = 4;                // warps 0..3 = warpgroup 0 (TMA, MMA, 2 idle); 4..11 = epilogue warpgroups
constexpr int THREADS = 32 * (EPI_WARP0 + EPI_WARPS);      // 384
constexpr int GSTAGE_WARP = 6 * 4096;       // gate-epilogue staging per warp: 6 arrays of [32 rows][32 units] f32
constexpr int NBARS = 2 * STAGES + 2 + 2;   // full[2], empty[2], acc_full, acc_empty, xchg[2]
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * XCHG_BYTES + 256 + 1024;
constexpr int TMEM_COLS = 512;
// The A operand (fp16 split of h for the cluster's 128 rows) is the same in all 8 CTAs: each CTA fetches a 16-row
// slice and TMA-multicasts it to the whole cluster (the stage's "empty" barrier then counts the MMA commits of all
// 8 CTAs).  L2->SMEM bytes per stage: gate 48 -> 34 KB, vocab 32 -> 18 KB.
#ifndef ICRL_DECODE_MCAST
#define ICRL_DECODE_MCAST 1
#endif
constexpr bool MCAST = ICRL_DECODE_MCAST != 0;
constexpr int A_SLICE_ROWS = BM / CL;                       // 16
constexpr int A_SLICE = A_SLICE_ROWS * BK * 2;              // 1 KB
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
  long long* prof;           // optional [n_cell][16] clock64 stamps of CTA 0's first epilogue warp (icrl_decode_set_profile)
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
// {k, row, part} box: the hi and lo' tiles of an operand arrive with ONE instruction (a single thread issues every
// TMA of the CTA; with four 2-D loads per 32-wide K block the GEMM phases were bound by TMA issue).
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
__device__ __forceinline__ unsigned long long smem_desc(unsigned addr) {
  return (unsigned long long)((addr & 0x3FFFF) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
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
  // map_h: MCAST ? 2-D [4B rows][512] box {32,16} : 3-D {512, B, 4 = buffer*2 + part} box {32,128,2}
  // map_whh / map_wv: 3-D {512, rows, 2 parts} box {32, 256 | 128, 2}
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
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, MCAST ? CL : 1); }
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

  if (warp == 2 || warp == 3) {
    // idle warps of warpgroup 0: only keep the cluster barrier counts complete
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    for (int j = 0; j < n_cell; ++j) { cluster_arrive(); cluster_wait(); }
  } else if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
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
          if (MCAST) {
            tma_load_2d_mcast(base + rank * A_SLICE, &map_h, kb * BK, arow + (int)rank * A_SLICE_ROWS, full, 0xFF);
            tma_load_2d_mcast(base + A_TILE + rank * A_SLICE, &map_h, kb * BK, arow + B + (int)rank * A_SLICE_ROWS, full, 0xFF);
          } else {
            tma_load_3d(base, &map_h, kb * BK, m0, 2 * (j & 1), full);
          }
          tma_load_3d(base + 2 * A_TILE, &map_whh, kb * BK, (int)rank * GN, 0, full);
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
            if (MCAST) {
              tma_load_2d_mcast(base + rank * A_SLICE, &map_h, kb * BK, arow + (int)rank * A_SLICE_ROWS, full, 0xFF);
              tma_load_2d_mcast(base + A_TILE + rank * A_SLICE, &map_h, kb * BK, arow + B + (int)rank * A_SLICE_ROWS, full, 0xFF);
            } else {
              tma_load_3d(base, &map_h, kb * BK, m0, 2 * ((j + 1) & 1), full);
            }
            tma_load_3d(base + 2 * A_TILE, &map_wv, kb * BK, (int)rank * VN, 0, full);   // hi at +0, lo' at +BV_TILE
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs the loops so that every address / descriptor stays warp-uniform (uniform registers);
    // elect.sync picks the lane that issues.  With the loops inside `if (lane == 0)` the compiler had to move
    // five operands per MMA from vector to uniform registers in an elect loop: ~130 cycles per MMA instead of 64.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    unsigned it = 0, nphase = 0;
    const bool profw = p.prof != nullptr && blockIdx.x == 0;
    for (int j = 0; j < n_cell; ++j) {
      {
        if (nphase > 0) mbar_wait(bar_acc_empty, (nphase - 1) & 1u);       // previous epilogue drained TMEM
        tc_fence_after();
        const long long tg0 = profw ? clock64() : 0;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(bar_full + 8 * s, (it / STAGES) & 1u);
          tc_fence_after();
          const unsigned base = smem_u32(smem + s * STAGE_BYTES);
          const unsigned long long dA0 = smem_desc(base), dA1 = smem_desc(base + A_TILE);
          const unsigned long long dB0 = smem_desc(base + 2 * A_TILE), dB1 = smem_desc(base + 2 * A_TILE + BG_TILE);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {                        // +2 = 32 bytes (16 fp16 along K) in 16-byte units
              tc_mma(tmem_base, dA0 + 2 * k, dB0 + 2 * k, idesc_f16(GN), (kb | k) != 0);        // hi  * hi  -> main
              tc_mma(tmem_base + GN, dA0 + 2 * k, dB1 + 2 * k, idesc_f16(GN), (kb | k) != 0);   // hi  * lo' -> correction
              tc_mma(tmem_base + GN, dA1 + 2 * k, dB0 + 2 * k, idesc_f16(GN), 1u);              // lo' * hi  -> correction
            }
            if (MCAST) tc_commit_mcast(bar_empty + 8 * s, 0xFF); else tc_commit(bar_empty + 8 * s);
          }
          __syncwarp();
        }
        if (elect_one()) tc_commit(bar_acc_full);
        ++nphase;
        if (profw && lane == 0) p.prof[(size_t)j * 16 + 14] = clock64() - tg0;   // gate-GEMM phase (issue span)
      }
      __syncwarp();
      cluster_arrive();
      cluster_wait();
      if (j >= p.p0 - 1) {
        mbar_wait(bar_acc_empty, (nphase - 1) & 1u);
        tc_fence_after();
        const long long tv0 = profw ? clock64() : 0;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(bar_full + 8 * s, (it / STAGES) & 1u);
          tc_fence_after();
          const unsigned base = smem_u32(smem + s * STAGE_BYTES);
          const unsigned long long dA0 = smem_desc(base), dA1 = smem_desc(base + A_TILE);
          const unsigned long long dB0 = smem_desc(base + 2 * A_TILE), dB1 = smem_desc(base + 2 * A_TILE + BV_TILE);
          const unsigned main_acc = tmem_base + (unsigned)(kb & 1) * VN;   // main term: even / odd K blocks
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              tc_mma(main_acc, dA0 + 2 * k, dB0 + 2 * k, idesc_f16(VN), (kb >= 2 || k != 0) ? 1u : 0u);
              tc_mma(tmem_base + 2 * VN, dA0 + 2 * k, dB1 + 2 * k, idesc_f16(VN), (kb | k) != 0);
              tc_mma(tmem_base + 2 * VN, dA1 + 2 * k, dB0 + 2 * k, idesc_f16(VN), 1u);
            }
            if (MCAST) tc_commit_mcast(bar_empty + 8 * s, 0xFF); else tc_commit(bar_empty + 8 * s);
          }
          __syncwarp();
        }
        if (elect_one()) tc_commit(bar_acc_full);
        ++nphase;
        if (profw && lane == 0) p.prof[(size_t)j * 16 + 15] = clock64() - tv0;   // vocab-GEMM phase (issue span)
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int ew = warp - EPI_WARP0;
    const int q = warp & 3;                    // TMEM lane quarter this warp may read
    const int ch = ew >> 2;                    // column half
    const int rl = 32 * q + lane;              // row inside the cluster tile
    const int row = m0 + rl;
    const bool valid = row < B;
    const unsigned tq = tmem_base + ((unsigned)(32 * q) << 16);
    const int slot = (int)rank * 2 + ch;
    unsigned rxb[CL], rbar[CL];                // this CTA's exchange buffer 0 / barrier 0 as seen in every peer
#pragma unroll
    for (unsigned d = 0; d < (unsigned)CL; ++d) { rxb[d] = mapa(smem_u32(xbuf), d); rbar[d] = mapa(bar_x, d); }
    const unsigned xoff = (unsigned)(slot * BM + rl) * 8u;
    auto xsend = [&](int which, unsigned long long v) {
#pragma unroll
      for (int d = 0; d < CL; ++d) st_async_b64(rxb[d] + which * XCHG_BYTES + xoff, v, rbar[d] + which * 8);
    };
    unsigned char* gst = smem + ew * GSTAGE_WARP;          // gate-epilogue staging of this warp (inside the TMA ring)
    float* lst = reinterpret_cast<float*>(xbuf + XCHG_BYTES + ew * 2048);   // logits staging [32 rows][16 cols]
    const int ucol0 = (int)rank * UN + 32 * ch;
    unsigned ephase = 0;                       // accumulator phases consumed
    unsigned xph0 = 0, xph1 = 0;               // exchange barrier phases
    int tok = valid ? p.tokcm[row] : 0;        // token consumed by cell 0 (prefix column 0)

    const bool prof = p.prof != nullptr && blockIdx.x == 0 && ew == 0 && lane == 0;
#define STAMP(k) do { if (prof) p.prof[(size_t)j * 16 + (k)] = clock64(); } while (0)
    for (int j = 0; j < n_cell; ++j) {
      const size_t BH = (size_t)B * H;
      STAMP(0);
      // ---------------- gate epilogue: cell update of [32 rows] x [32 hidden units] per warp
      mbar_wait(bar_acc_full, ephase & 1u);
      ++ephase;
      tc_fence_after();
      STAMP(1);
      // (L) gate-table rows of the consumed tokens (arrays 0..3 = i,f,g,o) and c_{j-1} (array 4) -> staging,
      //     8 lanes per row; element (row, 16-byte chunk c) lives at row*128 + ((c ^ (row & 7)) << 4).
      {
        const int c4 = lane & 7;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + (lane >> 3);
          const int tokr = __shfl_sync(0xffffffffu, tok, r);
          const int growr = m0 + 32 * q + r;
          if (growr < B) {
            const unsigned dst = smem_u32(gst) + (unsigned)(r * 128 + ((c4 ^ (r & 7)) << 4));
            const float* tsrc = p.table + (size_t)tokr * (4 * H) + ucol0 + c4 * 4;
#pragma unroll
            for (int a = 0; a < 4; ++a)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + a * 4096), "l"(tsrc + a * H) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 4 * 4096),
                         "l"(p.Cs + (size_t)j * BH + (size_t)growr * H + ucol0 + c4 * 4) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        __syncwarp();
      }
      STAMP(11);
      // (C) cell update, one row per lane, in place on the staging tile
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        float acc[4][8], cor[4][8];
#pragma unroll
        for (int g = 0; g < 4; ++g)
          tmem_ld8x2(tq + (unsigned)(g * UN + 32 * ch + 8 * c8), tq + (unsigned)(GN + g * UN + 32 * ch + 8 * c8), acc[g], cor[g]);
        unsigned char* e0 = gst + lane * 128 + (((2 * c8) ^ (lane & 7)) << 4);
        unsigned char* e1 = gst + lane * 128 + (((2 * c8 + 1) ^ (lane & 7)) << 4);
        float tin[5][8];
#pragma unroll
        for (int a = 0; a < 5; ++a) {
          const float4 u0 = *reinterpret_cast<const float4*>(e0 + a * 4096);
          const float4 u1 = *reinterpret_cast<const float4*>(e1 + a * 4096);
          tin[a][0] = u0.x; tin[a][1] = u0.y; tin[a][2] = u0.z; tin[a][3] = u0.w;
          tin[a][4] = u1.x; tin[a][5] = u1.y; tin[a][6] = u1.z; tin[a][7] = u1.w;
        }
        float gi[8], gf[8], gg[8], go[8], cn[8], hn[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          gi[i] = sigmoidf_sfu(fmaf(cor[0][i], LO_INV, acc[0][i]) + tin[0][i]);
          gf[i] = sigmoidf_sfu(fmaf(cor[1][i], LO_INV, acc[1][i]) + tin[1][i]);
          gg[i] = tanhf_sfu(fmaf(cor[2][i], LO_INV, acc[2][i]) + tin[2][i]);
          go[i] = sigmoidf_sfu(fmaf(cor[3][i], LO_INV, acc[3][i]) + tin[3][i]);
          cn[i] = gf[i] * tin[4][i] + gi[i] * gg[i];
          hn[i] = go[i] * tanhf_sfu(cn[i]);
        }
        *reinterpret_cast<float4*>(e0) = make_float4(gi[0], gi[1], gi[2], gi[3]);
        *reinterpret_cast<float4*>(e1) = make_float4(gi[4], gi[5], gi[6], gi[7]);
        *reinterpret_cast<float4*>(e0 + 4096) = make_float4(gf[0], gf[1], gf[2], gf[3]);
        *reinterpret_cast<float4*>(e1 + 4096) = make_float4(gf[4], gf[5], gf[6], gf[7]);
        *reinterpret_cast<float4*>(e0 + 2 * 4096) = make_float4(gg[0], gg[1], gg[2], gg[3]);
        *reinterpret_cast<float4*>(e1 + 2 * 4096) = make_float4(gg[4], gg[5], gg[6], gg[7]);
        *reinterpret_cast<float4*>(e0 + 3 * 4096) = make_float4(go[0], go[1], go[2], go[3]);
        *reinterpret_cast<float4*>(e1 + 3 * 4096) = make_float4(go[4], go[5], go[6], go[7]);
        *reinterpret_cast<float4*>(e0 + 4 * 4096) = make_float4(cn[0], cn[1], cn[2], cn[3]);
        *reinterpret_cast<float4*>(e1 + 4 * 4096) = make_float4(cn[4], cn[5], cn[6], cn[7]);
        *reinterpret_cast<float4*>(e0 + 5 * 4096) = make_float4(hn[0], hn[1], hn[2], hn[3]);
        *reinterpret_cast<float4*>(e1 + 5 * 4096) = make_float4(hn[4], hn[5], hn[6], hn[7]);
      }
      __syncwarp();
      STAMP(12);
      // (S) stash for backward: activated gates, c_j, h_j -- 8 lanes per row again
      {
        const int c4 = lane & 7;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + (lane >> 3);
          const int growr = m0 + 32 * q + r;
          if (growr < B) {
            const unsigned char* e = gst + r * 128 + ((c4 ^ (r & 7)) << 4);
            if (p.Gs) {
              float* gs = p.Gs + ((size_t)j * B + growr) * (4 * H) + ucol0 + c4 * 4;
#pragma unroll
              for (int a = 0; a < 4; ++a) *reinterpret_cast<float4*>(gs + a * H) = *reinterpret_cast<const float4*>(e + a * 4096);
            }
            const size_t o = (size_t)(j + 1) * BH + (size_t)growr * H + ucol0 + c4 * 4;
            *reinterpret_cast<float4*>(p.Cs + o) = *reinterpret_cast<const float4*>(e + 4 * 4096);
            const float4 h4 = *reinterpret_cast<const float4*>(e + 5 * 4096);
            *reinterpret_cast<float4*>(p.Hs + o) = h4;
            // fp16 split of h for the next GEMMs
            const __half a0 = __float2half_rn(h4.x), a1 = __float2half_rn(h4.y), a2 = __float2half_rn(h4.z), a3 = __float2half_rn(h4.w);
            __half2 hi2[2] = {__halves2half2(a0, a1), __halves2half2(a2, a3)};
            __half2 lo2[2] = {__halves2half2(__float2half_rn((h4.x - __half2float(a0)) * LO_SCALE), __float2half_rn((h4.y - __half2float(a1)) * LO_SCALE)),
                              __halves2half2(__float2half_rn((h4.z - __half2float(a2)) * LO_SCALE), __float2half_rn((h4.w - __half2float(a3)) * LO_SCALE))};
            __half* hp = p.hparts + ((size_t)(((j + 1) & 1) * 2) * B + growr) * H + ucol0 + c4 * 4;
            *reinterpret_cast<uint2*>(hp) = *reinterpret_cast<const uint2*>(hi2);
            *reinterpret_cast<uint2*>(hp + BH) = *reinterpret_cast<const uint2*>(lo2);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty);
      STAMP(2);
      __threadfence();
      fence_proxy_async();                     // generic-proxy writes of the h split -> visible to the peers' TMA loads
      cluster_arrive();
      cluster_wait();
      STAMP(3);

      const int s = j - (p.p0 - 1);
      if (s < 0) {                             // teacher-forced prefix cell: next token comes from the caption
        tok = valid ? p.tokcm[(size_t)(j + 1) * B + row] : 0;
        continue;
      }
      // ---------------- vocab epilogue, part 1: logits of this thread's 64 columns into registers
      mbar_wait(bar_acc_full, ephase & 1u);
      ++ephase;
      tc_fence_after();
      STAMP(4);
      float x[64];
      const int v0 = (int)rank * VN + 64 * ch;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        float m0v[8], m1v[8], cv[8];
        tmem_ld8x3(tq + (unsigned)(64 * ch + 8 * c8), tq + (unsigned)(VN + 64 * ch + 8 * c8),
                   tq + (unsigned)(2 * VN + 64 * ch + 8 * c8), m0v, m1v, cv);
        float bz[8];
#pragma unroll
        for (int h4 = 0; h4 < 2; ++h4) {                  // V % 4 == 0: a float4 of bias is all valid or all padding
          const int v = v0 + 8 * c8 + 4 * h4;
          const float4 b4 = v < V ? __ldg(reinterpret_cast<const float4*>(p.b_v + v)) : make_float4(0.f, 0.f, 0.f, 0.f);
          bz[4 * h4] = b4.x; bz[4 * h4 + 1] = b4.y; bz[4 * h4 + 2] = b4.z; bz[4 * h4 + 3] = b4.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int v = v0 + 8 * c8 + i;
          x[8 * c8 + i] = v < V ? fmaf(cv[i], LO_INV, m0v[i] + m1v[i]) + bz[i] : -INFINITY;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty);          // the next gate GEMM may overwrite TMEM now
      STAMP(13);
      if (p.logits || (p.last_logits && s == p.S - 1)) {
        // logits stash, 16 columns per pass: one row per lane into staging, 4 lanes per row out (8 lines per access)
        float* lgw = p.logits ? p.logits + (size_t)s * B * V : nullptr;
        float* llw = (p.last_logits && s == p.S - 1) ? p.last_logits : nullptr;
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<float4*>(lst + lane * 16 + ((c ^ ((lane >> 1) & 3)) << 2)) =
                make_float4(x[16 * ps + 4 * c], x[16 * ps + 4 * c + 1], x[16 * ps + 4 * c + 2], x[16 * ps + 4 * c + 3]);
          __syncwarp();
          const int c = lane & 3;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int r = it * 8 + (lane >> 2);
            const int growr = m0 + 32 * q + r;
            const int v = v0 + 16 * ps + 4 * c;
            if (growr < B && v + 3 < V) {                  // V % 4 == 0 is required by the host wrapper
              const float4 o = *reinterpret_cast<const float4*>(lst + r * 16 + ((c ^ ((r >> 1) & 3)) << 2));
              if (lgw) *reinterpret_cast<float4*>(lgw + (size_t)growr * V + v) = o;
              if (llw) *reinterpret_cast<float4*>(llw + (size_t)growr * V + v) = o;
            }
          }
          __syncwarp();
        }
      }
      STAMP(5);
      // ---------------- part 2: softmax + sampling across the cluster (trainers.py:444-458)
      // E1: row max and first argmax
      float mx = -INFINITY;
      int amax = VPAD;
#pragma unroll
      for (int i = 0; i < 64; ++i)
        if (x[i] > mx) { mx = x[i]; amax = v0 + i; }
      if (ew == 0 && lane == 0) mbar_expect_tx(bar_x, XCHG_BYTES);
      xsend(0, ((unsigned long long)(unsigned)amax << 32) | __float_as_uint(mx));
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
      STAMP(6);
      // E2: sum of exp(x - max) in f32 (F.softmax)
      float lsum = 0.f;
#pragma unroll
      for (int i = 0; i < 64; ++i) { x[i] = expf(x[i] - mx); lsum += x[i]; }
      if (ew == 0 && lane == 0) mbar_expect_tx(bar_x + 8, XCHG_BYTES);
      xsend(1, (unsigned long long)__float_as_uint(lsum));
      mbar_wait(bar_x + 8, xph1 & 1u);
      ++xph1;
      float tot = 0.f;
      {
        const unsigned long long* xs = reinterpret_cast<const unsigned long long*>(xbuf + XCHG_BYTES) + rl;
#pragma unroll
        for (int k = 0; k < XCHG_SLOTS; ++k) tot += __uint_as_float((unsigned)xs[k * BM]);
      }
      const float inv = 1.0f / tot;
      STAMP(7);
      int a;
      if (p.forced) {
        a = valid ? (int)p.forced[(size_t)row * p.S + s] : 0;
      } else if (p.greedy) {
        a = min(amax, V - 1);
      } else {
        // E3: the cdf of np.random.choice (cdf = cumsum(float64(p)); cdf /= cdf[-1]; searchsorted(u, 'right')),
        // carried in 62-bit fixed point: fx(p) = floor(p * 2^62) is exact for every f32 probability >= 2^-38 and
        // integer prefix sums are exact and independent of how the row is cut into partials, where the f64
        // cumsum rounds at 2^-53 per add.  (FP64 adds run at ~1/32 rate on this part: the f64 form of this
        // phase cost 21 K cycles per step, this one 3 K.)
        unsigned long long loc = 0;
#pragma unroll
        for (int i = 0; i < 64; ++i) loc += __float2ull_rz(x[i] * inv * 0x1p62f);
        if (ew == 0 && lane == 0) mbar_expect_tx(bar_x, XCHG_BYTES);
        xsend(0, loc);
        mbar_wait(bar_x, xph0 & 1u);
        ++xph0;
        unsigned long long pre = 0, total = 0;
        {
          const unsigned long long* xs = reinterpret_cast<const unsigned long long*>(xbuf) + rl;
#pragma unroll
          for (int k = 0; k < XCHG_SLOTS; ++k) {
            const unsigned long long d = xs[k * BM];
            if (k < slot) pre += d;
            total += d;
          }
        }
        STAMP(8);
        // threshold: cdf_k / total <= u  <=>  cdf_k <= floor(u * total), evaluated exactly (u = mu * 2^(eb-1075))
        unsigned long long thr;
        {
          const unsigned long long ub = (unsigned long long)__double_as_longlong(valid ? p.uniforms[(size_t)s * B + row] : 0.0);
          const int eb = (int)((ub >> 52) & 0x7FF);
          const unsigned long long mu = (ub & ((1ull << 52) - 1)) | (eb ? (1ull << 52) : 0ull);
          const unsigned long long hi = __umul64hi(mu, total), lo = mu * total;
          const int sh = 1075 - (eb ? eb : 1);                 // >= 53 because u < 1
          thr = sh >= 128 ? 0ull : (sh >= 64 ? (hi >> (sh - 64)) : ((hi << (64 - sh)) | (lo >> sh)));
        }
        unsigned long long run = pre;
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          run += __float2ull_rz(x[i] * inv * 0x1p62f);
          cnt += (v0 + i < V && run <= thr) ? 1 : 0;
        }
        STAMP(9);
        // E4: count of cdf entries <= u
        if (ew == 0 && lane == 0) mbar_expect_tx(bar_x + 8, XCHG_BYTES);
        xsend(1, (unsigned long long)(unsigned)cnt);
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
      STAMP(10);
    }
#undef STAMP
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

int make_map_f16_3d(CUtensorMap* map, const void* ptr, long long rows, int parts, int box_rows);

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
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    icrl_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %lld)", (int)r, rows);
    return ICRL_ERR_CUDA;
  }
  return ICRL_OK;
}

int make_map_f16_3d(CUtensorMap* map, const void* ptr, long long rows, int parts, int box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult qres;
    void* q = nullptr;
    ICRL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qres));
    ICRL_REQUIRE(q && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled unavailable");
    fn = reinterpret_cast<EncodeTiledFn>(q);
  }
  const cuuint64_t dims[3] = {(cuuint64_t)H, (cuuint64_t)rows, (cuuint64_t)parts};
  const cuuint64_t strides[2] = {(cuuint64_t)H * 2, (cuuint64_t)rows * H * 2};
  const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 2};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    icrl_set_error("cuTensorMapEncodeTiled (3-D) failed with CUresult %d (rows %lld)", (int)r, rows);
    return ICRL_ERR_CUDA;
  }
  return ICRL_OK;
}

}  // namespace

static long long* g_decode_prof = nullptr;
void icrl_decode_set_profile_impl(long long* buf) { g_decode_prof = buf; }

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
  if (MCAST) rc = make_map_f16(&mh, hparts, (long long)4 * B, A_SLICE_ROWS);
  else rc = make_map_f16_3d(&mh, hparts, B, 4, BM);
  if (rc) return rc;
  if ((rc = make_map_f16_3d(&mw, whh_pk, 4 * H, 2, GN))) return rc;
  if ((rc = make_map_f16_3d(&mv, wv_pk, VPAD, 2, VN))) return rc;
  DecodeArgs a;
  a.B = B; a.V = V; a.p0 = p0; a.S = S; a.greedy = greedy;
  a.table = table; a.b_v = b_v; a.uniforms = (greedy || forced) ? nullptr : uniforms; a.forced = forced;
  a.tokcm = tokcm; a.tokens_out = tokens_out; a.logp = logp; a.Hs = Hs; a.Cs = Cs; a.Gs = Gs; a.logits = logits;
  a.last_logits = last_logits; a.hparts = reinterpret_cast<__half*>(hparts); a.prof = g_decode_prof;
  const int n_mtiles = icrl_cdiv(B, BM);
  policy_decode_kernel<<<dim3(CL * n_mtiles), dim3(THREADS), SMEM_BYTES, st>>>(mh, mw, mv, a);
  ICRL_LAUNCH_CHECK();
  return ICRL_OK;
}
