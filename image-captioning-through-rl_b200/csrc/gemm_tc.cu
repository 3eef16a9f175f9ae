// Tensor-core GEMM for the policy decode step on sm_100a: C[M,N] (f32) = A[M,K] * B[N,K]^T (+ bias),
// computed as a split-precision bf16 product with f32 accumulation in tensor memory.  Each fp32 operand is
// split into THREE bf16 parts x = x0 + x1 + x2 (|x_i| ~ 2^-8i |x|) and the six products with i + j <= 2 are
// accumulated:   C = A0 B0' + A0 B1' + A1 B0' + A0 B2' + A1 B1' + A2 B0'     (dropped terms ~2^-24).
// Measured on the reference: single-pass bf16 / TF32 flip sampled tokens against
// the fp32 path; the 2-part split has 3e-6 logit error (measured here: 4.9e-6 of max|C|), which at 78 K
// samples per step is ~1 flipped token per step; the 3-part split is fp32-grade (3e-7) and keeps the
// bit-exact-token contract while running on the tcgen05 pipe.
//
// Structure (one CTA per 128x128 output tile, 192 threads, warp-specialised):
//   warp 0      TMA producer: per 64-wide K block loads the three A parts and three B parts (128x64 bf16
//               each, 128B-swizzled) into a 2-stage ring (96 KB per stage), completion on `full` mbarriers
//   warp 1      tcgen05.mma issuer (one lane): 6 operand pairs x 4 K-steps of 16 per stage, M=128 N=128;
//               tcgen05.commit releases the stage / signals the epilogue
//   warps 2-5   epilogue: tcgen05.ld 32 lanes x 32 columns at a time -> + bias -> global (f32)
//
// Accumulator layout.  The tensor core truncates (toward zero) when it adds an MMA's products into the f32
// accumulator: measured on B200 (scripts/tc_accuracy_probe.py) the relative bias is ~ -1.5e-8 per MMA
// instruction that touches an accumulator, i.e. -4.7e-7 after 32 (K = 512, bf16-exact operands) but -3.5e-6
// when all 192 MMAs of the 6-term split share one accumulator.  So the A0 B0' term -- the only one of full
// magnitude -- goes to two accumulators by K-block parity (16 MMAs each at K = 512) and the five correction
// terms (2^-8 and 2^-16 of the magnitude, so their truncation is 2^-8 smaller too) go to a third; the epilogue
// adds the three in f32 with round-to-nearest.
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 2, UMMA_K = 16, PARTS = 3;
constexpr int A_TILE = BM * BK * 2, B_TILE = BN * BK * 2;
constexpr int STAGE_BYTES = PARTS * (A_TILE + B_TILE);            // 98,304
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;     // + alignment slack + barriers
constexpr int TMEM_COLS = 512;   // three 128-column accumulators (main even/odd K blocks, correction); allocation must be a power of two
constexpr int THREADS = 192;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a descriptor / byte-count mistake must fail loudly (trap), never hang the device.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  for (unsigned spins = 0; spins < (1u << 26); ++spins) {
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
// K-major, 128B-swizzle shared-memory operand descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14),
// LBO>>4 = 1 [16,30) (unused for swizzled K-major), SBO>>4 = 64 [32,46) (8 rows x 128 B between row groups),
// version = 1 [46,48), layout SWIZZLE_128B = 2 [61,64).
__device__ __forceinline__ unsigned long long smem_desc(unsigned addr) {
  return (unsigned long long)((addr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor: c_format F32 = 1 [4,6), a/b_format BF16 = 1 [7,10)/[10,13), K-major A and B,
// n_dim = N>>3 [17,23), m_dim = M>>4 [24,29).
constexpr unsigned IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(BN >> 3) << 17) | ((unsigned)(BM >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                   const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b0,
                   const __grid_constant__ CUtensorMap map_b1, const __grid_constant__ CUtensorMap map_b2, int M,
                   int N, int K, float* __restrict__ C, int ldc, const float* __restrict__ bias) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + STAGES * STAGE_BYTES);
  // bars[0..STAGES) full, [STAGES..2*STAGES) empty, [2*STAGES] tmem_full; then the TMEM base address
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 2 * STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int KB = K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[STAGES + s]), 1); }
    mbar_init(smem_u32(&bars[2 * STAGES]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b2) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        const unsigned ph = (unsigned)(kb / STAGES) & 1u;
        mbar_wait(smem_u32(&bars[STAGES + s]), ph ^ 1u);          // slot free (fresh barrier: passes at once)
        const unsigned full = smem_u32(&bars[s]);
        mbar_expect_tx(full, STAGE_BYTES);
        const unsigned base = smem_u32(smem + s * STAGE_BYTES);
        tma_load_2d(base, &map_a0, kb * BK, m0, full);
        tma_load_2d(base + A_TILE, &map_a1, kb * BK, m0, full);
        tma_load_2d(base + 2 * A_TILE, &map_a2, kb * BK, m0, full);
        tma_load_2d(base + 3 * A_TILE, &map_b0, kb * BK, n0, full);
        tma_load_2d(base + 3 * A_TILE + B_TILE, &map_b1, kb * BK, n0, full);
        tma_load_2d(base + 3 * A_TILE + 2 * B_TILE, &map_b2, kb * BK, n0, full);
      }
    }
  } else if (warp == 1) {
    // the whole warp runs the loop (descriptors stay in uniform registers); elect.sync picks the issuing lane
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % STAGES;
      const unsigned ph = (unsigned)(kb / STAGES) & 1u;
      mbar_wait(smem_u32(&bars[s]), ph);
      tc_fence_after();
      const unsigned base = smem_u32(smem + s * STAGE_BYTES);
      unsigned long long da[PARTS], db[PARTS];
#pragma unroll
      for (int i = 0; i < PARTS; ++i) { da[i] = smem_desc(base + i * A_TILE); db[i] = smem_desc(base + 3 * A_TILE + i * B_TILE); }
      if (elect_one()) {
        // pairs (i,j), i + j <= 2: (0,0) main -> accumulator kb & 1; (0,1) (1,0) (0,2) (1,1) (2,0) corrections -> accumulator 2
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {          // +2 = 32 bytes along K in 16-byte descriptor units
          tc_mma(tmem_base + (unsigned)(kb & 1) * BN, da[0] + 2 * k, db[0] + 2 * k, IDESC, (kb >= 2 || k != 0) ? 1u : 0u);
          tc_mma(tmem_base + 2 * BN, da[0] + 2 * k, db[1] + 2 * k, IDESC, (kb | k) != 0 ? 1u : 0u);
          tc_mma(tmem_base + 2 * BN, da[1] + 2 * k, db[0] + 2 * k, IDESC, 1u);
          tc_mma(tmem_base + 2 * BN, da[0] + 2 * k, db[2] + 2 * k, IDESC, 1u);
          tc_mma(tmem_base + 2 * BN, da[1] + 2 * k, db[1] + 2 * k, IDESC, 1u);
          tc_mma(tmem_base + 2 * BN, da[2] + 2 * k, db[0] + 2 * k, IDESC, 1u);
        }
        tc_commit(smem_u32(&bars[STAGES + s]));                   // stage reusable once these MMAs retire
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(smem_u32(&bars[2 * STAGES]));      // accumulator complete
    __syncwarp();
  } else {
    const int q = warp & 3;                                       // TMEM lane quarter this warp may read
    const int row = m0 + 32 * q + lane;
    mbar_wait(smem_u32(&bars[2 * STAGES]), 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      unsigned v[32], v1[32], v2[32];
      const unsigned tl = tmem_base + ((unsigned)(32 * q) << 16) + (unsigned)(c * 32);
      tmem_ld32(tl, v);
      tmem_ld32(tl + 2 * BN, v2);
      if (KB > 1) {
        tmem_ld32(tl + BN, v1);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v1[j]));
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
      const int col0 = n0 + c * 32;
      if (row < M && col0 < N) {
        float* dst = C + (size_t)row * ldc + col0;
        if (col0 + 32 <= N && (ldc & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            o.x = __uint_as_float(v[j + 0]) + (bias ? bias[col0 + j + 0] : 0.f);
            o.y = __uint_as_float(v[j + 1]) + (bias ? bias[col0 + j + 1] : 0.f);
            o.z = __uint_as_float(v[j + 2]) + (bias ? bias[col0 + j + 2] : 0.f);
            o.w = __uint_as_float(v[j + 3]) + (bias ? bias[col0 + j + 3] : 0.f);
            *reinterpret_cast<float4*>(dst + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < N) dst[j] = __uint_as_float(v[j]) + (bias ? bias[col0 + j] : 0.f);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// x -> three bf16 parts, parts[0][i] + parts[1][i] + parts[2][i] ~= x[i] (exact to ~2^-24)
__global__ void split_bf16x3_kernel(long long n, const float* __restrict__ x, __nv_bfloat16* __restrict__ parts) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const __nv_bfloat16 p0 = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(p0);
    const __nv_bfloat16 p1 = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(p1);
    parts[i] = p0;
    parts[n + i] = p1;
    parts[2 * n + i] = __float2bfloat16_rn(r2);
  }
}

// W [rows][cols] fp32 -> parts [3][cols][Kp] bf16 of W^T, columns k >= rows zero (operand of C = A W for A [M][Kp])
__global__ void pack_transposed_bf16x3_kernel(int rows, int cols, int Kp, const float* __restrict__ W,
                                              __nv_bfloat16* __restrict__ parts) {
  const long long n = (long long)cols * Kp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / Kp), k = (int)(i % Kp);
    const float v = k < rows ? W[(size_t)k * cols + c] : 0.f;
    const __nv_bfloat16 p0 = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(p0);
    const __nv_bfloat16 p1 = __float2bfloat16_rn(r1);
    parts[i] = p0;
    parts[n + i] = p1;
    parts[2 * n + i] = __float2bfloat16_rn(r1 - __bfloat162float(p1));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(CUtensorMap* map, const void* ptr, int rows, int K, int box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult qres;
    void* p = nullptr;
    ICRL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    ICRL_REQUIRE(p && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled unavailable");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    icrl_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %d, K %d)", (int)r, rows, K);
    return ICRL_ERR_CUDA;
  }
  return ICRL_OK;
}

}  // namespace

int icrl_split_bf16x3_impl(cudaStream_t st, long long n, const float* x, void* parts) {
  const int blocks = (int)min((long long)148 * 8, (n + 255) / 256);
  split_bf16x3_kernel<<<blocks, 256, 0, st>>>(n, x, (__nv_bfloat16*)parts);
  ICRL_LAUNCH_CHECK();
  return ICRL_OK;
}

int icrl_pack_transposed_bf16x3_impl(cudaStream_t st, int rows, int cols, int Kp, const float* W, void* parts) {
  ICRL_REQUIRE(rows > 0 && cols > 0 && Kp >= rows, "bad shape");
  pack_transposed_bf16x3_kernel<<<148 * 4, 256, 0, st>>>(rows, cols, Kp, W, (__nv_bfloat16*)parts);
  ICRL_LAUNCH_CHECK();
  return ICRL_OK;
}

// a_parts: [3][M][K] bf16, b_parts: [3][N][K] bf16 (icrl_split_bf16x3 layout)
int icrl_gemm_bf16x3_impl(cudaStream_t st, int M, int N, int K, const void* a_parts, const void* b_parts, float* C,
                          int ldc, const float* bias) {
  ICRL_REQUIRE(M > 0 && N > 0 && K > 0 && K % BK == 0, "K must be a multiple of 64");
  ICRL_REQUIRE(((uintptr_t)a_parts | (uintptr_t)b_parts) % 16 == 0 && ((size_t)M * K) % 8 == 0 && ((size_t)N * K) % 8 == 0,
               "operands must be 16B aligned");
  static bool attr_set = false;
  if (!attr_set) {
    ICRL_CUDA(cudaFuncSetAttribute(gemm_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap ma[3], mb[3];
  const __nv_bfloat16* ap = (const __nv_bfloat16*)a_parts;
  const __nv_bfloat16* bp = (const __nv_bfloat16*)b_parts;
  for (int i = 0; i < 3; ++i) {
    int rc;
    if ((rc = make_map(&ma[i], ap + (size_t)i * M * K, M, K, BM)) || (rc = make_map(&mb[i], bp + (size_t)i * N * K, N, K, BN)))
      return rc;
  }
  dim3 grid(icrl_cdiv(N, BN), icrl_cdiv(M, BM));
  gemm_bf16x3_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(ma[0], ma[1], ma[2], mb[0], mb[1], mb[2], M, N, K, C, ldc, bias);
  ICRL_LAUNCH_CHECK();
  return ICRL_OK;
}
