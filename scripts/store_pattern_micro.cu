// Microbenchmark: how fast can one CTA per SM push fp32 rows towards L2 / HBM with
//   (A) coalesced stores: 8 lanes x 16 B = one 128-byte row chunk, 4 rows per warp instruction (what the chain kernels'
//       staged store passes do), and
//   (B) thread-per-row stores: every lane writes ONE full 32-byte sector of its own row with st.global.v8.f32
//       (what an epilogue that keeps its results in registers could do without a shared-memory staging pass)?
// Rows are 8 KB apart (a [position][2048] f32 array), 120 CTAs x 256 threads, each CTA writes `iters` x 128 KB.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/store_pattern_micro.cu -o /tmp/store_micro
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 1) pat_a(float* out, int iters, long long* cyc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // 128 rows x 256 floats per CTA-iteration: warp w owns rows 16 w .. 16 w + 15; 8 lanes per 32-float chunk
    float* base = out + ((size_t)blockIdx.x * iters + it) * 128 * 2048;
    for (int c = 0; c < 8; ++c)            // 8 column chunks of 32 floats
      for (int r4 = 0; r4 < 4; ++r4) {
        const int row = 16 * warp + 4 * r4 + (lane >> 3);
        const float4 v = make_float4(it, c, row, lane);
        *reinterpret_cast<float4*>(base + (size_t)row * 2048 + 32 * c + 4 * (lane & 7)) = v;
      }
  }
  __threadfence();
  if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

__global__ void __launch_bounds__(256, 1) pat_b(float* out, int iters, long long* cyc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // the same 128 rows x 256 floats: thread t owns row (t & 127) and 16 of the 32 sectors of its 1 KB
    float* base = out + ((size_t)blockIdx.x * iters + it) * 128 * 2048;
    const int row = 32 * (warp & 3) + lane, half = warp >> 2;
    for (int c = 0; c < 16; ++c) {
      float* d = base + (size_t)row * 2048 + 8 * (16 * half + c);
      const float a = it, b = c;
      asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(d), "f"(a), "f"(b), "f"(a), "f"(b), "f"(a), "f"(b),
                   "f"(a), "f"(b) : "memory");
    }
  }
  __threadfence();
  if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

int main() {
  const int ctas = 120, iters = 400;
  float* out;
  long long* cyc;
  cudaMalloc(&out, (size_t)ctas * iters * 128 * 2048 * sizeof(float));     // 12.6 GB
  cudaMalloc(&cyc, ctas * sizeof(long long));
  long long h[ctas];
  for (int rep = 0; rep < 2; ++rep)
    for (int pat = 0; pat < 2; ++pat) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      if (pat == 0) pat_a<<<ctas, 256>>>(out, iters, cyc); else pat_b<<<ctas, 256>>>(out, iters, cyc);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bytes = (double)iters * 128 * 256 * 4;
      printf("pattern %c: %.3f ms, %.1f B/clk/SM (max CTA cycles %lld), %.2f TB/s aggregate  [%s]\n", pat ? 'B' : 'A', ms,
             bytes / mx, mx, ctas * bytes / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
