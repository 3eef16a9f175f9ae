// FP32 issue-rate microbenchmark for the register-resident batched GEMV of the chain kernels (round-2 question:
// is the forward GEMV phase, ~1.4 cycles per instruction and scheduler, limited by the FFMA operand pattern or by the
// number of warps?).  Every variant does the same work per thread as one chunk of the forward kernel:
// NG = 4 gate rows x 16 weights per lane against NB hidden vectors (values from shared memory), ITER times.
//   variant 0  current loop order: for k-group, for v: load h[v]; for g: 4 dependent FMAs   (h operand reused across g)
//   variant 1  weight-stationary:  for k, for g: w = w[g][k]; for v: acc[g][v] += w * h[v][k] (w operand reused across v)
//   variant 2  variant 0 with fma.rn.f32x2 on pairs of shards (v, v+1)
//   variant 3  no reuse at all (batch-1 pattern: 3 distinct registers per FMA)
// Reported: cycles per call, FMA per clock and SM, at 256 threads per CTA (2 warps per scheduler) and, for 8 shards, 512.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/ffma_micro scripts/ffma_micro.cu && ./scripts/ffma_micro
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NG = 4, KW = 16, H = 512;

template <int NB, int VARIANT, int TH>
__global__ void __launch_bounds__(TH, 1) gemv_kernel(const float* __restrict__ wsrc, const float* __restrict__ hsrc, float* out, long long* cycles, int iters) {
  extern __shared__ __align__(16) float hb[];          // [NB][H]
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NB * H; i += blockDim.x) hb[i] = hsrc[i];
  float w[NG][KW];
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int k = 0; k < KW; ++k) w[g][k] = wsrc[(g * KW + k) * 32 + lane];
  __syncthreads();
  float acc[NG][NB];
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int v = 0; v < NB; ++v) acc[g][v] = 0.f;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    asm volatile("" ::: "memory");                  // the hidden vectors change every step in the real kernel: reload them
    if constexpr (VARIANT == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int v = 0; v < NB; ++v) {
          const float4 hv = *reinterpret_cast<const float4*>(&hb[v * H + 128 * j + 4 * lane]);
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            acc[g][v] = fmaf(w[g][4 * j + 0], hv.x, acc[g][v]);
            acc[g][v] = fmaf(w[g][4 * j + 1], hv.y, acc[g][v]);
            acc[g][v] = fmaf(w[g][4 * j + 2], hv.z, acc[g][v]);
            acc[g][v] = fmaf(w[g][4 * j + 3], hv.w, acc[g][v]);
          }
        }
    } else if constexpr (VARIANT == 3) {               // same count, but neighbouring FMAs never share an operand register
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int v = 0; v < NB; ++v) {
          const float4 hv = *reinterpret_cast<const float4*>(&hb[v * H + 128 * j + 4 * lane]);
          const float ha[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int g = 0; g < NG; ++g) acc[g][v] = fmaf(w[g][4 * j + ((k + g) & 3)], ha[(k + g) & 3], acc[g][v]);
        }
    } else if constexpr (VARIANT == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int v0 = 0; v0 < NB; v0 += 4) {             // four shards' h values in registers at a time
          float4 hv[4];
#pragma unroll
          for (int v = 0; v < 4; ++v) hv[v] = *reinterpret_cast<const float4*>(&hb[(v0 + v) * H + 128 * j + 4 * lane]);
#pragma unroll
          for (int g = 0; g < NG; ++g) {
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[g][v0 + v] = fmaf(w[g][4 * j + 0], hv[v].x, acc[g][v0 + v]);
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[g][v0 + v] = fmaf(w[g][4 * j + 1], hv[v].y, acc[g][v0 + v]);
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[g][v0 + v] = fmaf(w[g][4 * j + 2], hv[v].z, acc[g][v0 + v]);
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[g][v0 + v] = fmaf(w[g][4 * j + 3], hv[v].w, acc[g][v0 + v]);
          }
        }
    } else {                                           // VARIANT == 2: packed pairs of shards
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int v = 0; v < NB; v += 2) {
          const float4 h0 = *reinterpret_cast<const float4*>(&hb[v * H + 128 * j + 4 * lane]);
          const float4 h1 = *reinterpret_cast<const float4*>(&hb[(v + 1) * H + 128 * j + 4 * lane]);
          const float ha[4] = {h0.x, h0.y, h0.z, h0.w}, hc[4] = {h1.x, h1.y, h1.z, h1.w};
#pragma unroll
          for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              unsigned long long a, b, c, d;
              asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(w[g][4 * j + k]), "f"(w[g][4 * j + k]));
              asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(ha[k]), "f"(hc[k]));
              asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc[g][v]), "f"(acc[g][v + 1]));
              asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
              asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[g][v]), "=f"(acc[g][v + 1]) : "l"(d));
            }
        }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int v = 0; v < NB; ++v) s += acc[g][v];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int NB, int VARIANT, int TH>
static void run(const char* name, const float* w, const float* h, float* out, long long* cyc) {
  const int iters = 2000, sms = 148, threads = TH;
  const size_t smem = (size_t)NB * H * sizeof(float);
  cudaFuncSetAttribute(gemv_kernel<NB, VARIANT, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  gemv_kernel<NB, VARIANT, TH><<<sms, threads, smem>>>(w, h, out, cyc, 10);
  gemv_kernel<NB, VARIANT, TH><<<sms, threads, smem>>>(w, h, out, cyc, iters);
  long long c = 0;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const cudaError_t e = cudaGetLastError();
  const double fma_per_thread = (double)iters * NG * KW * NB;
  printf("%-34s NB=%2d threads=%3d: %8.1f cycles per call, %6.1f FMA/clk/SM%s\n", name, NB, threads, (double)c / iters,
         fma_per_thread * threads / (double)c, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  float *w, *h, *out;
  long long* cyc;
  cudaMalloc(&w, NG * KW * 32 * sizeof(float));
  cudaMalloc(&h, 16 * H * sizeof(float));
  cudaMalloc(&out, 148 * 512 * sizeof(float));
  cudaMalloc(&cyc, sizeof(long long));
  cudaMemset(w, 0, NG * KW * 32 * sizeof(float));
  cudaMemset(h, 0, 16 * H * sizeof(float));
  run<8, 0, 256>("h reused across gates (current)", w, h, out, cyc);
  run<8, 1, 256>("weight reused across shards", w, h, out, cyc);
  run<8, 2, 256>("fma.rn.f32x2 on shard pairs", w, h, out, cyc);
  run<8, 3, 256>("no operand reuse", w, h, out, cyc);
  run<16, 0, 256>("h reused across gates (current)", w, h, out, cyc);
  run<16, 1, 256>("weight reused across shards", w, h, out, cyc);
  run<16, 2, 256>("fma.rn.f32x2 on shard pairs", w, h, out, cyc);
  run<8, 0, 512>("h reused across gates (current)", w, h, out, cyc);       // 4 warps per scheduler, <= 128 registers
  run<8, 1, 512>("weight reused across shards", w, h, out, cyc);
  run<8, 2, 512>("fma.rn.f32x2 on shard pairs", w, h, out, cyc);
  return 0;
}
