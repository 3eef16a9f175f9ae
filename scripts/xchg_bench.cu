// Latency microbenchmarks behind the serial-chain kernel design (DESIGN.md "chain latency budget").
// Measures, per serial step, on a B200:
//   l2  : all-to-all exchange of N floats between G CTAs through L2 with {value,tag} 64-bit words
//   dsm : the same exchange inside one thread-block cluster through DSMEM st.async + mbarrier
//   math: the LSTM pointwise chain (3 sigmoid + 2 tanh, accurate expf/tanhf)
//   shfl: 5-level butterfly reduction of 4 values
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o xchg_bench xchg_bench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void ld2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
}
__device__ __forceinline__ void ld2_relaxed(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
}
__device__ __forceinline__ void st1(unsigned long long* p, unsigned long long w) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ void st1_relaxed(unsigned long long* p, unsigned long long w) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}

// ---------------------------------------------------------------- L2 tagged all-to-all
// N = total floats exchanged per step; each CTA has `threads` threads; every thread fetches N/(2*threads)
// pairs; producers: the first N/G threads of each CTA store one word each.
template <int MODE>   // 0 volatile, 1 relaxed.gpu
__global__ void l2_xchg_kernel(unsigned long long* xchg, int N, int steps, int* fail, float* sink) {
  extern __shared__ float sm[];
  const int G = gridDim.x, per_cta = N / G;
  const int pairs = N / 2;
  float acc = 0.f;
  for (int t = 0; t < steps; ++t) {
    unsigned long long* buf = xchg + (size_t)(t & 1) * N;
    if (threadIdx.x < per_cta) {
      const unsigned long long w = ((unsigned long long)(t + 1) << 32) | (unsigned)__float_as_uint(acc + 1.f);
      if (MODE == 0) st1(buf + blockIdx.x * per_cta + threadIdx.x, w); else st1_relaxed(buf + blockIdx.x * per_cta + threadIdx.x, w);
    }
    for (int e = threadIdx.x; e < pairs; e += blockDim.x) {
      unsigned long long a, b;
      unsigned spins = 0;
      while (true) {
        if (MODE == 0) ld2(buf + 2 * e, a, b); else ld2_relaxed(buf + 2 * e, a, b);
        if ((unsigned)(a >> 32) == (unsigned)(t + 1) && (unsigned)(b >> 32) == (unsigned)(t + 1)) break;
        if (++spins > (1u << 22)) { *fail = 1; return; }
      }
      sm[2 * e] = __uint_as_float((unsigned)a);
      sm[2 * e + 1] = __uint_as_float((unsigned)b);
    }
    __syncthreads();
    acc = sm[(threadIdx.x * 7) % N] * 0.5f;
  }
  if (acc == 123.f) sink[0] = acc;
}

// per-warp fetch variant: every warp pulls the whole vector itself (no smem, no __syncthreads)
__global__ void l2_xchg_warp_kernel(unsigned long long* xchg, int N, int steps, int* fail, float* sink) {
  const int G = gridDim.x, per_cta = N / G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc = 0.f;
  for (int t = 0; t < steps; ++t) {
    unsigned long long* buf = xchg + (size_t)(t & 1) * N;
    if (lane == 0 && warp < per_cta) st1(buf + blockIdx.x * per_cta + warp, ((unsigned long long)(t + 1) << 32) | (unsigned)__float_as_uint(acc + 1.f));
    float s = 0.f;
    for (int e = lane; e < N / 2; e += 32) {
      unsigned long long a, b;
      unsigned spins = 0;
      while (true) {
        ld2(buf + 2 * e, a, b);
        if ((unsigned)(a >> 32) == (unsigned)(t + 1) && (unsigned)(b >> 32) == (unsigned)(t + 1)) break;
        if (++spins > (1u << 22)) { *fail = 1; return; }
      }
      s += __uint_as_float((unsigned)a) + __uint_as_float((unsigned)b);
    }
    acc = s * 0.25f;
  }
  if (acc == 123.f) sink[0] = acc;
}

// ---------------------------------------------------------------- DSMEM cluster all-to-all
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_f32(unsigned raddr, float v, unsigned rbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(__float_as_uint(v)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void st_async_v4(unsigned raddr, float4 v, unsigned rbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
               "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\nbra WAIT_LOOP;\nDONE:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// cluster of C CTAs; N floats per step in total, each CTA produces N/C of them (from its first N/C/4 lanes
// as float4) and broadcasts to all C CTAs.  VEC=1: v4 stores, VEC=0: scalar stores.
template <int VEC>
__global__ void dsm_xchg_kernel(int N, int steps, float* sink) {
  extern __shared__ __align__(16) float smd[];          // [2][N] data
  __shared__ __align__(8) unsigned long long bars[2];
  unsigned C, rank;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(C));
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int per_cta = N / C;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  float acc = 0.f;
  for (int t = 0; t < steps; ++t) {
    const int b = t & 1;
    const unsigned bar = smem_u32(&bars[b]);
    if (threadIdx.x == 0) mbar_expect_tx(bar, N * 4);
    // produce + broadcast
    if (VEC) {
      const int nvec = per_cta / 4;                    // float4 per CTA
      for (int i = threadIdx.x; i < nvec * (int)C; i += blockDim.x) {
        const int dst = i / nvec, q = i % nvec;
        const float4 v = make_float4(acc + 1.f, acc + 2.f, acc + 3.f, acc + 4.f);
        st_async_v4(mapa(smem_u32(&smd[b * N + rank * per_cta + q * 4]), dst), v, mapa(bar, dst));
      }
    } else {
      for (int i = threadIdx.x; i < per_cta * (int)C; i += blockDim.x) {
        const int dst = i / per_cta, q = i % per_cta;
        st_async_f32(mapa(smem_u32(&smd[b * N + rank * per_cta + q]), dst), acc + 1.f, mapa(bar, dst));
      }
    }
    mbar_wait(bar, (t >> 1) & 1);
    acc = smd[b * N + (threadIdx.x * 7) % N] * 0.5f;
    __syncthreads();     // all reads of this buffer done before it is re-armed two steps later (conservative)
  }
  if (acc == 123.f) sink[0] = acc;
  cluster_sync_all();
}

// ---------------------------------------------------------------- math / shuffle chains
__global__ void math_kernel(int steps, float* out) {
  float h = 0.1f * threadIdx.x, c = 0.f;
  for (int t = 0; t < steps; ++t) {
    const float i = 1.f / (1.f + expf(-(h + 0.1f)));
    const float f = 1.f / (1.f + expf(-(h - 0.2f)));
    const float g = tanhf(h + 0.3f);
    const float o = 1.f / (1.f + expf(-(h * 0.5f)));
    c = f * c + i * g;
    h = o * tanhf(c);
  }
  out[threadIdx.x] = h;
}
__global__ void math_fast_kernel(int steps, float* out) {
  float h = 0.1f * threadIdx.x, c = 0.f;
  for (int t = 0; t < steps; ++t) {
    const float i = __fdividef(1.f, 1.f + __expf(-(h + 0.1f)));
    const float f = __fdividef(1.f, 1.f + __expf(-(h - 0.2f)));
    const float g = 2.f * __fdividef(1.f, 1.f + __expf(-2.f * (h + 0.3f))) - 1.f;
    const float o = __fdividef(1.f, 1.f + __expf(-(h * 0.5f)));
    c = f * c + i * g;
    h = o * (2.f * __fdividef(1.f, 1.f + __expf(-2.f * c)) - 1.f);
  }
  out[threadIdx.x] = h;
}
__global__ void shfl_kernel(int steps, float* out) {
  float a0 = threadIdx.x, a1 = 1.f, a2 = 2.f, a3 = 3.f;
  for (int t = 0; t < steps; ++t) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
    }
    a0 = a0 * 1e-3f + a3; a1 = a1 * 1e-3f; a2 = a2 * 1e-3f; a3 = a3 * 1e-3f + 1.f;
  }
  out[threadIdx.x] = a0 + a1 + a2 + a3;
}
// GEMV from registers: 4 rows x 16 k per lane (the chain kernel's inner loop) + smem broadcast reads
__global__ void gemv_kernel(int steps, float* out) {
  __shared__ __align__(16) float h[512];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) h[i] = 0.001f * i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float w[4][16];
  for (int g = 0; g < 4; ++g) for (int j = 0; j < 16; ++j) w[g][j] = 1e-3f * (g + j + lane);
  float s = 0.f;
  for (int t = 0; t < steps; ++t) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 hv = *reinterpret_cast<const float4*>(&h[128 * j + 4 * lane]);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        acc[g] = fmaf(w[g][4 * j], hv.x, acc[g]); acc[g] = fmaf(w[g][4 * j + 1], hv.y, acc[g]);
        acc[g] = fmaf(w[g][4 * j + 2], hv.z, acc[g]); acc[g] = fmaf(w[g][4 * j + 3], hv.w, acc[g]);
      }
    }
    s += acc[0] + acc[1] + acc[2] + acc[3];
    h[(threadIdx.x + t) & 511] = s * 1e-6f;     // keep the loop dependent on the previous step
    __syncthreads();
  }
  out[threadIdx.x] = s;
}


// ---------------------------------------------------------------- second round: floors and polling variants
// one-way latency floors: two CTAs bounce a tagged word (L2) / two CTAs of a cluster bounce via st.async
__global__ void pingpong_l2_kernel(unsigned long long* xchg, int steps, int* fail) {
  if (threadIdx.x != 0) return;
  const int me = blockIdx.x;
  for (int t = 0; t < steps; ++t) {
    if ((t & 1) == me) {
      st1(xchg + 64 * me, ((unsigned long long)(t + 1) << 32) | 1u);
    } else {
      unsigned long long a, b; unsigned spins = 0;
      while (true) { ld2(xchg + 64 * (1 - me), a, b); if ((unsigned)(a >> 32) == (unsigned)(t + 1)) break; if (++spins > (1u << 22)) { *fail = 1; return; } }
    }
  }
}
__global__ void __cluster_dims__(2, 1, 1) pingpong_dsm_kernel(int steps, float* sink) {
  __shared__ __align__(16) float slot[4];
  __shared__ __align__(8) unsigned long long bar;
  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) mbar_init(smem_u32(&bar), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x == 0) {
    const unsigned rslot = mapa(smem_u32(&slot[0]), 1 - rank), rbar = mapa(smem_u32(&bar), 1 - rank);
    unsigned phase = 0;
    for (int t = 0; t < steps; ++t) {
      if ((t & 1) == (int)rank) {
        st_async_f32(rslot, (float)t, rbar);
      } else {
        mbar_expect_tx(smem_u32(&bar), 4);
        mbar_wait(smem_u32(&bar), phase);
        phase ^= 1;
      }
    }
    if (slot[0] == -5.f) sink[0] = 1.f;
  }
  cluster_sync_all();
}

// L2 all-to-all where only POLLERS threads per CTA poll (all loads issued before any check; only the
// failing ones are re-polled), PIPE=1 keeps a second, time-staggered poll in flight.
template <int POLLERS, int PIPE>
__global__ void l2_xchg2_kernel(unsigned long long* xchg, int N, int steps, int* fail, float* sink) {
  extern __shared__ float sm[];
  const int G = gridDim.x, per_cta = N / G;
  constexpr int MAXP = 8;                    // pairs per poller thread (N=512: 256/POLLERS)
  const int ppt = (N / 2) / POLLERS;
  float acc = 0.f;
  for (int t = 0; t < steps; ++t) {
    unsigned long long* buf = xchg + (size_t)(t & 1) * N;
    const unsigned tag = (unsigned)(t + 1);
    if (threadIdx.x < per_cta) st1(buf + blockIdx.x * per_cta + threadIdx.x, ((unsigned long long)tag << 32) | (unsigned)__float_as_uint(acc + 1.f));
    if (threadIdx.x < POLLERS) {
      unsigned long long a[MAXP], b[MAXP];
      unsigned pending = (1u << ppt) - 1u, spins = 0;
      while (pending) {
#pragma unroll
        for (int i = 0; i < MAXP; ++i)
          if (i < ppt && (pending >> i & 1)) ld2(buf + 2 * (i * POLLERS + threadIdx.x), a[i], b[i]);
        if (PIPE) { const long long c0 = clock64(); while (clock64() - c0 < 150) {} }
#pragma unroll
        for (int i = 0; i < MAXP; ++i)
          if (i < ppt && (pending >> i & 1) && (unsigned)(a[i] >> 32) == tag && (unsigned)(b[i] >> 32) == tag) {
            pending &= ~(1u << i);
            sm[2 * (i * POLLERS + threadIdx.x)] = __uint_as_float((unsigned)a[i]);
            sm[2 * (i * POLLERS + threadIdx.x) + 1] = __uint_as_float((unsigned)b[i]);
          }
        if (++spins > (1u << 22)) { *fail = 1; return; }
      }
    }
    __syncthreads();
    acc = sm[(threadIdx.x * 7) % N] * 0.5f;
  }
  if (acc == 123.f) sink[0] = acc;
}

__global__ void math_mid_kernel(int steps, float* out) {
  float h = 0.1f * threadIdx.x, c = 0.f;
  for (int t = 0; t < steps; ++t) {
    const float i = __frcp_rn(1.f + expf(-(h + 0.1f)));
    const float f = __frcp_rn(1.f + expf(-(h - 0.2f)));
    const float g = 1.f - 2.f * __frcp_rn(1.f + expf(2.f * (h + 0.3f)));
    const float o = __frcp_rn(1.f + expf(-(h * 0.5f)));
    c = f * c + i * g;
    h = o * (1.f - 2.f * __frcp_rn(1.f + expf(2.f * c)));
  }
  out[threadIdx.x] = h;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const double ghz = prop.clockRate * 1e-6;
  printf("device %s, %d SMs, %.3f GHz\n", prop.name, prop.multiProcessorCount, ghz);
  unsigned long long* xchg;
  int* fail;
  float* sink;
  CK(cudaMalloc(&xchg, 2 * 4096 * sizeof(unsigned long long)));
  CK(cudaMalloc(&fail, 4));
  CK(cudaMalloc(&sink, 4096 * 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int steps = 20000;

  // ---- L2 exchange
  const int Ns[] = {512, 2048};
  const int Gs[] = {16, 32, 64, 128};
  for (int mode = 0; mode < 3; ++mode)
    for (int N : Ns)
      for (int G : Gs) {
        for (int threads : {128, 256}) {
          if (mode == 2 && threads != 256) continue;
          CK(cudaMemset(xchg, 0, 2 * 4096 * sizeof(unsigned long long)));
          CK(cudaMemset(fail, 0, 4));
          int n = N, st = steps;
          void* args[] = {&xchg, &n, &st, &fail, &sink};
          const void* fn = mode == 0 ? (const void*)l2_xchg_kernel<0> : mode == 1 ? (const void*)l2_xchg_kernel<1> : (const void*)l2_xchg_warp_kernel;
          if (mode == 2 && N / G > threads / 32) continue;
          for (int rep = 0; rep < 2; ++rep) {
            CK(cudaMemset(xchg, 0, 2 * 4096 * sizeof(unsigned long long)));
            CK(cudaEventRecord(e0));
            CK(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(threads), args, mode == 2 ? 0 : N * 4, 0));
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
          }
          int hf = 0;
          CK(cudaMemcpy(&hf, fail, 4, cudaMemcpyDeviceToHost));
          const double ns = time_ms(e0, e1) * 1e6 / steps;
          printf("l2 mode=%s N=%4d G=%3d threads=%3d : %8.1f ns/step = %6.0f cyc %s\n", mode == 0 ? "volatile" : mode == 1 ? "relaxed " : "perwarp ",
                 N, G, threads, ns, ns * ghz, hf ? "FAILED(watchdog)" : "");
        }
      }

  // ---- DSMEM exchange
  for (int vec = 0; vec < 2; ++vec)
    for (int C : {8, 16})
      for (int N : Ns) {
        const void* fn = vec ? (const void*)dsm_xchg_kernel<1> : (const void*)dsm_xchg_kernel<0>;
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 2048 * 4));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(C); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 2 * N * 4; cfg.stream = 0;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = N, st = steps;
        cudaError_t err = cudaSuccess;
        for (int rep = 0; rep < 2 && err == cudaSuccess; ++rep) {
          CK(cudaEventRecord(e0));
          if (vec) err = cudaLaunchKernelEx(&cfg, dsm_xchg_kernel<1>, n, st, sink); else err = cudaLaunchKernelEx(&cfg, dsm_xchg_kernel<0>, n, st, sink);
          CK(cudaEventRecord(e1));
          cudaError_t e2 = cudaDeviceSynchronize();
          if (err == cudaSuccess) err = e2;
        }
        if (err != cudaSuccess) { printf("dsm vec=%d C=%2d N=%4d : launch failed: %s\n", vec, C, N, cudaGetErrorString(err)); cudaGetLastError(); continue; }
        const double ns = time_ms(e0, e1) * 1e6 / steps;
        printf("dsm vec=%d C=%2d N=%4d : %8.1f ns/step = %6.0f cyc\n", vec, C, N, ns, ns * ghz);
      }

  // ---- math / shuffle / gemv chains
  struct { const char* name; void (*fn)(int, float*); int threads; } ks[] = {
      {"lstm pointwise (expf/tanhf)", math_kernel, 32}, {"lstm pointwise (fast intrinsics)", math_fast_kernel, 32},
      {"butterfly reduce x4", shfl_kernel, 32}, {"gemv 4x512 regs, 8 warps + bar", gemv_kernel, 256}};
  for (auto& k : ks) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      k.fn<<<1, k.threads>>>(steps, sink);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
    }
    const double ns = time_ms(e0, e1) * 1e6 / steps;
    printf("%-36s : %8.1f ns/step = %6.0f cyc\n", k.name, ns, ns * ghz);
  }
  // ---- second round
  {
    int st = steps;
    void* args[] = {&xchg, &st, &fail};
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaMemset(xchg, 0, 2 * 4096 * sizeof(unsigned long long)));
      CK(cudaEventRecord(e0));
      CK(cudaLaunchCooperativeKernel((const void*)pingpong_l2_kernel, dim3(2), dim3(32), args, 0, 0));
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
    }
    double ns = time_ms(e0, e1) * 1e6 / steps;
    printf("pingpong L2 (one-way store->poll)    : %8.1f ns = %6.0f cyc\n", ns, ns * ghz);
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      pingpong_dsm_kernel<<<2, 32>>>(steps, sink);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
    }
    ns = time_ms(e0, e1) * 1e6 / steps;
    printf("pingpong DSMEM (st.async->mbar wait) : %8.1f ns = %6.0f cyc\n", ns, ns * ghz);
  }
  {
    struct V { const char* name; const void* fn; } vs[] = {
        {"pollers=256 pipe=0", (const void*)l2_xchg2_kernel<256, 0>}, {"pollers=256 pipe=1", (const void*)l2_xchg2_kernel<256, 1>},
        {"pollers=128 pipe=0", (const void*)l2_xchg2_kernel<128, 0>}, {"pollers= 64 pipe=0", (const void*)l2_xchg2_kernel<64, 0>},
        {"pollers= 32 pipe=0", (const void*)l2_xchg2_kernel<32, 0>},  {"pollers= 64 pipe=1", (const void*)l2_xchg2_kernel<64, 1>}};
    for (auto& v : vs)
      for (int G : {32, 64}) {
        int n = 512, st = steps;
        void* args[] = {&xchg, &n, &st, &fail, &sink};
        CK(cudaMemset(fail, 0, 4));
        for (int rep = 0; rep < 2; ++rep) {
          CK(cudaMemset(xchg, 0, 2 * 4096 * sizeof(unsigned long long)));
          CK(cudaEventRecord(e0));
          CK(cudaLaunchCooperativeKernel(v.fn, dim3(G), dim3(256), args, 512 * 4, 0));
          CK(cudaEventRecord(e1));
          CK(cudaDeviceSynchronize());
        }
        int hf = 0;
        CK(cudaMemcpy(&hf, fail, 4, cudaMemcpyDeviceToHost));
        const double ns = time_ms(e0, e1) * 1e6 / steps;
        printf("l2v2 %s N=512 G=%3d : %8.1f ns/step = %6.0f cyc %s\n", v.name, G, ns, ns * ghz, hf ? "FAILED" : "");
      }
  }
  {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      math_mid_kernel<<<1, 32>>>(steps, sink);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
    }
    const double ns = time_ms(e0, e1) * 1e6 / steps;
    printf("%-36s : %8.1f ns/step = %6.0f cyc\n", "lstm pointwise (expf + frcp_rn)", ns, ns * ghz);
  }
  return 0;
}
