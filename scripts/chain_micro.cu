// Second round of latency microbenchmarks for the serial-chain kernels.
//   pp   : one-way store->poll latency between two CTAs through L2 for every {publish, poll} instruction pair
//   gemv : cycles per step of a register(+smem)-resident batch-1 GEMV slice, scalar FFMA vs packed fma.rn.f32x2,
//          for the per-SM weight volumes of the 64-CTA (L2 exchange) and 16-CTA-cluster (DSMEM exchange) designs
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o chain_micro chain_micro.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

typedef unsigned long long u64;

template <int ST> __device__ __forceinline__ void publish(u64* p, u64 w) {
  if (ST == 0) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
  if (ST == 1) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
  if (ST == 2) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
  if (ST == 3) { u64 old; asm volatile("atom.relaxed.gpu.global.exch.b64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(w) : "memory"); }
  if (ST == 4) asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
  if (ST == 5) { asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory"); asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
  if (ST == 6) asm volatile("st.global.wt.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
  if (ST == 7) asm volatile("st.global.cg.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
template <int LD> __device__ __forceinline__ u64 poll(const u64* p) {
  u64 v = 0;
  if (LD == 0) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  if (LD == 1) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  if (LD == 2) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  if (LD == 3) asm volatile("atom.relaxed.gpu.global.or.b64 %0, [%1], 0;" : "=l"(v) : "l"(p) : "memory");
  if (LD == 4) asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  if (LD == 5) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// CTA `a` and CTA `b` of the grid bounce a tagged word; everybody else exits.
template <int ST, int LD>
__global__ void pp_kernel(u64* xchg, int steps, int a, int b, int* fail, int* smids) {
  if (threadIdx.x != 0) return;
  int me;
  if ((int)blockIdx.x == a) me = 0; else if ((int)blockIdx.x == b) me = 1; else return;
  unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); smids[me] = (int)sm;
  for (int t = 0; t < steps; ++t) {
    if ((t & 1) == me) {
      publish<ST>(xchg + 64 * me, ((u64)(t + 1) << 32) | 1u);
    } else {
      unsigned spins = 0;
      while (true) {
        const u64 v = poll<LD>(xchg + 64 * (1 - me));
        if ((unsigned)(v >> 32) == (unsigned)(t + 1)) break;
        if (++spins > (1u << 22)) { *fail = 1; return; }
      }
    }
  }
}

// pipelined poll: NP loads in flight, re-issued round robin; LD=0 semantics
template <int NP>
__global__ void pp_pipe_kernel(u64* xchg, int steps, int a, int b, int* fail) {
  if (threadIdx.x != 0) return;
  int me;
  if ((int)blockIdx.x == a) me = 0; else if ((int)blockIdx.x == b) me = 1; else return;
  for (int t = 0; t < steps; ++t) {
    if ((t & 1) == me) {
      publish<0>(xchg + 64 * me, ((u64)(t + 1) << 32) | 1u);
    } else {
      const u64* p = xchg + 64 * (1 - me);
      u64 v[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) { v[i] = poll<0>(p); if (i + 1 < NP) { const long long c0 = clock64(); while (clock64() - c0 < 600 / NP) {} } }
      unsigned spins = 0;
      bool done = false;
      while (!done) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          if ((unsigned)(v[i] >> 32) == (unsigned)(t + 1)) { done = true; break; }
          v[i] = poll<0>(p);
        }
        if (++spins > (1u << 22)) { *fail = 1; return; }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ gemv
__device__ __forceinline__ u64 pack2(float a, float b) {
  u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}

// Each warp owns RR rows with weights in registers and RS rows with weights in shared memory; lanes split
// K = 512 (16 per lane per row).  Per step: read h (4 x float4 per lane), FMAs, transposed butterfly
// reduction, lane r writes row r's sum, __syncthreads.  F2 = packed fma.rn.f32x2.
template <int WARPS, int RR, int RS, bool F2>
__global__ void __launch_bounds__(WARPS * 32, 1) gemv_kernel(int steps, float* out, long long* cyc, const float* wsrc) {
  extern __shared__ __align__(16) float smem[];
  float* h = smem;                                  // [2][512]
  float* ws = smem + 1024;                          // [WARPS][RS][4][32 lanes][4]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int R = RR + RS;
  for (int i = threadIdx.x; i < 1024; i += WARPS * 32) h[i] = 0.001f * (i & 511);
  for (int i = threadIdx.x; i < WARPS * RS * 512; i += WARPS * 32) ws[i] = 1e-3f * (i % 37);
  float w[RR > 0 ? RR : 1][16];
  u64 w2[RR > 0 ? RR : 1][8];
#pragma unroll
  for (int r = 0; r < RR; ++r)
#pragma unroll
    for (int j = 0; j < 16; ++j) w[r][j] = wsrc[((warp * 8 + r) * 16 + j) * 32 + lane];
#pragma unroll
  for (int r = 0; r < RR; ++r)
#pragma unroll
    for (int j = 0; j < 8; ++j) w2[r][j] = pack2(w[r][2 * j], w[r][2 * j + 1]);
  __syncthreads();
  float s = 0.f;
  const long long c0 = clock64();
  for (int t = 0; t < steps; ++t) {
    const float* hb = h + (t & 1) * 512;
    float acc[R];
    if (!F2) {
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 hv = *reinterpret_cast<const float4*>(&hb[128 * j + 4 * lane]);
#pragma unroll
        for (int r = 0; r < RR; ++r) {
          acc[r] = fmaf(w[r][4 * j], hv.x, acc[r]); acc[r] = fmaf(w[r][4 * j + 1], hv.y, acc[r]);
          acc[r] = fmaf(w[r][4 * j + 2], hv.z, acc[r]); acc[r] = fmaf(w[r][4 * j + 3], hv.w, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < RS; ++r) {
          const float4 wv = *reinterpret_cast<const float4*>(&ws[(((warp * RS + r) * 4 + j) * 32 + lane) * 4]);
          acc[RR + r] = fmaf(wv.x, hv.x, acc[RR + r]); acc[RR + r] = fmaf(wv.y, hv.y, acc[RR + r]);
          acc[RR + r] = fmaf(wv.z, hv.z, acc[RR + r]); acc[RR + r] = fmaf(wv.w, hv.w, acc[RR + r]);
        }
      }
    } else {
      u64 a2[R];
#pragma unroll
      for (int r = 0; r < R; ++r) a2[r] = 0ull;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const ulonglong2 hv = *reinterpret_cast<const ulonglong2*>(&hb[128 * j + 4 * lane]);
#pragma unroll
        for (int r = 0; r < RR; ++r) { a2[r] = fma2(w2[r][2 * j], hv.x, a2[r]); a2[r] = fma2(w2[r][2 * j + 1], hv.y, a2[r]); }
#pragma unroll
        for (int r = 0; r < RS; ++r) {
          const ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(&ws[(((warp * RS + r) * 4 + j) * 32 + lane) * 4]);
          a2[RR + r] = fma2(wv.x, hv.x, a2[RR + r]); a2[RR + r] = fma2(wv.y, hv.y, a2[RR + r]);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) { float x, y; unpack2(a2[r], x, y); acc[r] = x + y; }
    }
    // transposed butterfly: after it, lane l holds the full sum of row (l % R') for the power-of-two padded R
    float v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = r < R ? acc[r] : 0.f;
    // 8 -> 4
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const bool up = lane & 16;
      const float send = up ? v[r] : v[r + 4];
      const float keep = up ? v[r + 4] : v[r];
      v[r] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const bool up = lane & 8;
      const float send = up ? v[r] : v[r + 2];
      const float keep = up ? v[r + 2] : v[r];
      v[r] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
      const bool up = lane & 4;
      const float send = up ? v[0] : v[1];
      const float keep = up ? v[1] : v[0];
      v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    s += v[0];
    if ((lane & 3) == 0) h[((t + 1) & 1) * 512 + ((warp * 8 + (lane >> 2)) * 7 + t) % 512] = v[0] * 1e-6f;
    __syncthreads();
  }
  const long long c1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
  out[blockIdx.x * WARPS * 32 + threadIdx.x] = s;
}

template <int WARPS, int RR, int RS, bool F2>
static void run_gemv(const char* name, float* out, long long* cyc) {
  const int steps = 20000;
  const int smem = (1024 + WARPS * RS * 512) * 4;
  CK(cudaFuncSetAttribute(gemv_kernel<WARPS, RR, RS, F2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  gemv_kernel<WARPS, RR, RS, F2><<<16, WARPS * 32, smem>>>(steps, out, cyc, out + 16 * 1024);
  CK(cudaDeviceSynchronize());
  long long c;
  CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, gemv_kernel<WARPS, RR, RS, F2>));
  printf("gemv %-34s warps=%2d rows/warp=%d(reg)+%d(smem) f32x2=%d regs=%3d spill=%zu : %7.1f cyc/step  (%d MACs/SM/step)\n", name,
         WARPS, RR, RS, (int)F2, fa.numRegs, (size_t)fa.localSizeBytes, (double)c / steps, WARPS * (RR + RS) * 512);
}

template <int ST, int LD>
static void run_pp(const char* sn, const char* ln, u64* xchg, int* fail, int* smids, int a, int b, double ghz) {
  const int steps = 20000;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaMemset(xchg, 0, 4096));
  CK(cudaMemset(fail, 0, 4));
  pp_kernel<ST, LD><<<148, 32>>>(xchg, 200, a, b, fail, smids);
  CK(cudaMemset(xchg, 0, 4096));
  CK(cudaEventRecord(e0));
  pp_kernel<ST, LD><<<148, 32>>>(xchg, steps, a, b, fail, smids);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  int hf, sm[2];
  CK(cudaMemcpy(&hf, fail, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(sm, smids, 8, cudaMemcpyDeviceToHost));
  const double ns = ms * 1e6 / steps;
  printf("pp st=%-18s ld=%-16s ctas=(%3d,%3d) sm=(%3d,%3d): %7.1f ns = %5.0f cyc one-way %s\n", sn, ln, a, b, sm[0], sm[1], ns, ns * ghz, hf ? "FAILED" : "");
}

template <int NP>
static void run_pp_pipe(u64* xchg, int* fail, int a, int b, double ghz) {
  const int steps = 20000;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaMemset(xchg, 0, 4096));
  CK(cudaMemset(fail, 0, 4));
  CK(cudaEventRecord(e0));
  pp_pipe_kernel<NP><<<148, 32>>>(xchg, steps, a, b, fail);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  int hf; CK(cudaMemcpy(&hf, fail, 4, cudaMemcpyDeviceToHost));
  const double ns = ms * 1e6 / steps;
  printf("pp pipelined polls in flight=%d ctas=(%3d,%3d): %7.1f ns = %5.0f cyc one-way %s\n", NP, a, b, ns, ns * ghz, hf ? "FAILED" : "");
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const double ghz = prop.clockRate * 1e-6;
  printf("device %s, %d SMs, %.3f GHz\n", prop.name, prop.multiProcessorCount, ghz);
  u64* xchg; int* fail; int* smids; float* out; long long* cyc;
  CK(cudaMalloc(&xchg, 4096)); CK(cudaMalloc(&fail, 4)); CK(cudaMalloc(&smids, 8));
  CK(cudaMalloc(&out, (16 * 1024 + 16 * 8 * 16 * 32) * 4)); CK(cudaMemset(out, 0, (16 * 1024 + 16 * 8 * 16 * 32) * 4)); CK(cudaMalloc(&cyc, 16 * 8));

#define PP(ST, SN, LD, LN) run_pp<ST, LD>(SN, LN, xchg, fail, smids, 0, 1, ghz)
  PP(0, "st.volatile", 0, "ld.volatile");
  PP(0, "st.volatile", 1, "ld.relaxed.gpu");
  PP(0, "st.volatile", 2, "ld.acquire.gpu");
  PP(0, "st.volatile", 3, "atom.or 0");
  PP(0, "st.volatile", 4, "ld.cv");
  PP(0, "st.volatile", 5, "ld.cg");
  PP(1, "st.relaxed.gpu", 1, "ld.relaxed.gpu");
  PP(2, "st.release.gpu", 1, "ld.relaxed.gpu");
  PP(3, "atom.exch", 0, "ld.volatile");
  PP(3, "atom.exch", 3, "atom.or 0");
  PP(4, "red.max", 0, "ld.volatile");
  PP(4, "red.max", 3, "atom.or 0");
  PP(5, "st.volatile+fence", 0, "ld.volatile");
  PP(6, "st.wt", 0, "ld.volatile");
  PP(7, "st.cg", 0, "ld.volatile");
  PP(7, "st.cg", 5, "ld.cg");
  const int others[] = {2, 17, 37, 74, 75, 111, 147};
  for (int o : others) run_pp<0, 0>("st.volatile", "ld.volatile", xchg, fail, smids, 0, o, ghz);
  for (int o : others) run_pp<4, 3>("red.max", "atom.or 0", xchg, fail, smids, 0, o, ghz);
  run_pp_pipe<2>(xchg, fail, 0, 1, ghz);
  run_pp_pipe<4>(xchg, fail, 0, 1, ghz);
  run_pp_pipe<8>(xchg, fail, 0, 1, ghz);

  run_gemv<8, 4, 0, false>("64-CTA design (today)", out, cyc);
  run_gemv<8, 4, 0, true>("64-CTA design, f32x2", out, cyc);
  run_gemv<16, 4, 0, false>("16 warps x 4 rows", out, cyc);
  run_gemv<16, 4, 0, true>("16 warps x 4 rows, f32x2", out, cyc);
  run_gemv<16, 6, 2, false>("cluster-16 LSTM slice", out, cyc);
  run_gemv<16, 6, 2, true>("cluster-16 LSTM slice, f32x2", out, cyc);
  run_gemv<16, 5, 3, false>("cluster-16 LSTM slice (5+3)", out, cyc);
  run_gemv<16, 5, 3, true>("cluster-16 LSTM slice (5+3), f32x2", out, cyc);
  run_gemv<16, 6, 0, false>("cluster-16 GRU slice", out, cyc);
  run_gemv<16, 6, 0, true>("cluster-16 GRU slice, f32x2", out, cyc);
  run_gemv<12, 8, 0, false>("12 warps x 8 rows regs", out, cyc);
  run_gemv<12, 8, 0, true>("12 warps x 8 rows regs, f32x2", out, cyc);
  run_gemv<8, 8, 0, false>("32-CTA design: 8 warps x 8 rows", out, cyc);
  run_gemv<8, 8, 0, true>("32-CTA design: 8 warps x 8 rows, f32x2", out, cyc);
  printf("done\n");
  return 0;
}
