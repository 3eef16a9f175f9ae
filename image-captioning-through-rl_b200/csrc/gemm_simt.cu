// FP32 SIMT GEMM used for every dense contraction on the path whose result feeds a
// bit-exact-token or 1e-5 parity contract (see DESIGN.md "precision").  Plain CUDA
// cores, fp32 FMA accumulation: C = beta*C + bias + op(A) * op(B).
//
//   transA = 0 : A is [M][K] row-major (K contiguous)      A(m,k) = A[m*lda + k]
//   transA = 1 : A is [K][M] row-major (M contiguous)      A(m,k) = A[k*lda + m]
//   transB = 0 : B is [K][N] row-major (N contiguous)      B(k,n) = B[k*ldb + n]
//   transB = 1 : B is [N][K] row-major (K contiguous)      B(k,n) = B[n*ldb + k]   (C = A * B^T)
//
// Tile 128x128x16, 256 threads, 8x8 outputs per thread held as 2x2 blocks of 4x4 so the
// shared-memory reads are 128-bit.  Contractions with a huge K and few output tiles (the
// weight-gradient GEMMs, K = number of serial steps) are split along K over gridDim.z into a
// workspace and reduced by a second kernel in a fixed order (deterministic, no atomics).
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

template <bool TA, bool TB, bool VEC>
__global__ void __launch_bounds__(256)
gemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
            float* __restrict__ C, int ldc, const float* __restrict__ bias, float beta, int k_per_split,
            float* __restrict__ ws) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- A tile -> As[k][m]
    if (!TA) {                      // K contiguous: each thread two float4 along k
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = (tid >> 2) + r * 64, kq = (tid & 3) * 4;
        const int m = m0 + row, k = k0 + kq;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (m < M) {
          if (VEC) {
            if (k < kend) {         // K % 4 == 0 and k_per_split % 4 == 0 guaranteed by the host
              const float4 t = *reinterpret_cast<const float4*>(A + (size_t)m * lda + k);
              v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) if (k + q < kend) v[q] = A[(size_t)m * lda + k + q];
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) As[kq + q][row] = v[q];
      }
    } else {                        // M contiguous: each thread two float4 along m
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int kk = (tid >> 5) + r * 8, mq = (tid & 31) * 4;
        const int k = k0 + kk, m = m0 + mq;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < kend) {
          if (VEC) {
            if (m < M) t = *reinterpret_cast<const float4*>(A + (size_t)k * lda + m);
          } else {
            const float* p = A + (size_t)k * lda + m;
            if (m + 0 < M) t.x = p[0];
            if (m + 1 < M) t.y = p[1];
            if (m + 2 < M) t.z = p[2];
            if (m + 3 < M) t.w = p[3];
          }
        }
        *reinterpret_cast<float4*>(&As[kk][mq]) = t;
      }
    }
    // ---- B tile -> Bs[k][n]
    if (TB) {                       // K contiguous
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = (tid >> 2) + r * 64, kq = (tid & 3) * 4;
        const int n = n0 + row, k = k0 + kq;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (n < N) {
          if (VEC) {
            if (k < kend) {
              const float4 t = *reinterpret_cast<const float4*>(B + (size_t)n * ldb + k);
              v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) if (k + q < kend) v[q] = B[(size_t)n * ldb + k + q];
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) Bs[kq + q][row] = v[q];
      }
    } else {                        // N contiguous
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int kk = (tid >> 5) + r * 8, nq = (tid & 31) * 4;
        const int k = k0 + kk, n = n0 + nq;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < kend) {
          if (VEC) {
            if (n < N) t = *reinterpret_cast<const float4*>(B + (size_t)k * ldb + n);
          } else {
            const float* p = B + (size_t)k * ldb + n;
            if (n + 0 < N) t.x = p[0];
            if (n + 1 < N) t.y = p[1];
            if (n + 2 < N) t.z = p[2];
            if (n + 3 < N) t.w = p[3];
          }
        }
        *reinterpret_cast<float4*>(&Bs[kk][nq]) = t;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      if (split) {
        ws[((size_t)blockIdx.z * M + m) * N + n] = acc[i][j];
      } else {
        float v = acc[i][j];
        if (bias) v += bias[n];
        if (beta != 0.f) v += beta * C[(size_t)m * ldc + n];
        C[(size_t)m * ldc + n] = v;
      }
    }
  }
}

__global__ void splitk_reduce_kernel(int M, int N, int splits, const float* __restrict__ ws, float* __restrict__ C,
                                     int ldc, const float* __restrict__ bias, float beta) {
  const size_t total = (size_t)M * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / N), n = (int)(i % N);
    float v = 0.f;
    for (int z = 0; z < splits; ++z) v += ws[(size_t)z * total + i];
    if (bias) v += bias[n];
    if (beta != 0.f) v += beta * C[(size_t)m * ldc + n];
    C[(size_t)m * ldc + n] = v;
  }
}

template <bool TA, bool TB>
void launch(bool vec, dim3 grid, cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B,
            int ldb, float* C, int ldc, const float* bias, float beta, int kps, float* ws) {
  if (vec)
    gemm_kernel<TA, TB, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, beta, kps, ws);
  else
    gemm_kernel<TA, TB, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, beta, kps, ws);
}

}  // namespace

// Returns the number of kernels launched through *launches (may be null).
int icrl_gemm_f32_impl(cudaStream_t st, int transA, int transB, int M, int N, int K, const float* A, int lda,
                       const float* B, int ldb, float* C, int ldc, const float* bias, float beta, float* ws,
                       size_t ws_bytes, int* launches) {
  ICRL_REQUIRE(M > 0 && N > 0 && K > 0, "empty GEMM");
  ICRL_REQUIRE(A && B && C, "null operand");
  const int contigA = transA ? M : K, contigB = transB ? K : N;
  const bool vec = (contigA % 4 == 0) && (contigB % 4 == 0) && (lda % 4 == 0) && (ldb % 4 == 0) &&
                   ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0);
  const int tiles = icrl_cdiv(M, BM) * icrl_cdiv(N, BN);
  int splits = 1;
  if (ws && tiles < 148 && K >= 512) {
    // few output tiles: split K so that ~2 waves of CTAs exist; very long K (weight gradients over the serial steps) in
    // chunks >= 1024, everything else (gate-table products with K = vocabulary, K = 2048 BPTT products of small local
    // batches, K = batch rows) in chunks >= 256
    splits = min(icrl_cdiv(2 * 148, tiles), K >= 65536 ? K / 1024 : K / 256);
    while (splits > 1 && (size_t)splits * M * N * sizeof(float) > ws_bytes) --splits;
    if (splits < 1) splits = 1;
  }
  int kps = icrl_cdiv(K, splits);
  kps = icrl_cdiv(kps, BK) * BK;               // multiple of 16 (keeps float4 k-loads whole)
  splits = icrl_cdiv(K, kps);
  dim3 grid(icrl_cdiv(N, BN), icrl_cdiv(M, BM), splits);
  if (!transA && !transB) launch<false, false>(vec, grid, st, M, N, K, A, lda, B, ldb, C, ldc, bias, beta, kps, ws);
  else if (!transA && transB) launch<false, true>(vec, grid, st, M, N, K, A, lda, B, ldb, C, ldc, bias, beta, kps, ws);
  else if (transA && !transB) launch<true, false>(vec, grid, st, M, N, K, A, lda, B, ldb, C, ldc, bias, beta, kps, ws);
  else launch<true, true>(vec, grid, st, M, N, K, A, lda, B, ldb, C, ldc, bias, beta, kps, ws);
  ICRL_LAUNCH_CHECK();
  int n = 1;
  if (splits > 1) {
    const size_t total = (size_t)M * N;
    splitk_reduce_kernel<<<(int)min((size_t)1184, (total + 255) / 256), 256, 0, st>>>(M, N, splits, ws, C, ldc, bias, beta);
    ICRL_LAUNCH_CHECK();
    n = 2;
  }
  if (launches) *launches += n;
  return ICRL_OK;
}
