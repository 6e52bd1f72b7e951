// Micro-benchmarks that size the DAS kernel design on B200 (sm_100a):
//  (1) FP32 issue: FFMA / FADD scalar vs packed FFMA2 / FADD2 (fma.rn.f32x2, add.rn.f32x2)
//  (2) shared-memory wavefronts of LDS.32/.64/.128 under lane-broadcast patterns
//  (3) overlap of packed FP32 with LDS.128
// Output: one line per test, "name warps/SM cycles ops/clk/SM".
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float add1(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }

#define NACC 16
#define ITER 2048

// mode 0: FFMA, 1: FFMA2, 2: FADD2, 3: FFMA2+FADD2 (DAS shape), 4: FFMA+FADD, 5: FADD
template <int MODE>
__global__ void fp_kernel(float* out, long long* cyc, float seed) {
  float a[NACC]; u64 p[NACC];
  for (int i = 0; i < NACC; i++) { a[i] = seed + i + threadIdx.x; p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] * 0.5f); }
  float f = seed * 0.999f, g = seed * 1e-3f;
  u64 f2 = ((u64)__float_as_uint(f) << 32) | __float_as_uint(f);
  u64 g2 = ((u64)__float_as_uint(g) << 32) | __float_as_uint(g);
  __syncthreads();
  long long t0 = clk();
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      if (MODE == 0) a[i] = fma1(a[i], f, g);
      if (MODE == 1) p[i] = fma2(p[i], f2, g2);
      if (MODE == 2) p[i] = add2(p[i], g2);
      if (MODE == 3) p[i] = add2(p[i], fma2(f2, g2, p[(i + 1) % NACC]));
      if (MODE == 4) a[i] = add1(a[i], fma1(f, g, a[(i + 1) % NACC]));
      if (MODE == 5) a[i] = add1(a[i], g);
    }
  }
  long long t1 = clk();
  float s = 0; for (int i = 0; i < NACC; i++) s += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  __shared__ long long smin, smax;
  if (threadIdx.x == 0) { smin = t0; smax = t1; }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { atomicMin(&smin, t0); atomicMax(&smax, t1); }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = smax - smin;
}

// LDS patterns. VEC = 1,2,4 words. pattern: chunk index of lane l (in units of VEC words).
// 0: l   1: l/2   2: l/4   3: 0   4: l/4+(l&1)   5: l/8  6: (l/4)*2 (stride-2 chunks, 4-lane groups) 7: l/4 + 3*(l&3) (use-case like: 4 dirs with offsets)
__device__ __forceinline__ int pat(int P, int l) {
  switch (P) { case 0: return l; case 1: return l / 2; case 2: return l / 4; case 3: return 0; case 4: return l / 4 + (l & 1);
    case 5: return l / 8; case 6: return (l / 4) * 2; default: return l / 4 + 3 * (l & 3); }
}
#define LITER 2048
template <int VEC, int NF>
__global__ void lds_kernel(float* out, long long* cyc, int P, float seed) {
  extern __shared__ float4 smem4[];
  float* sm = (float*)smem4;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
  __syncthreads();
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (w & 3) * 4096 + pat(P, lane) * VEC * 4;
  float acc = 0, acc2 = 0;
  u64 p[8]; for (int i = 0; i < 8; i++) p[i] = (u64)(threadIdx.x + i) * 0x3f80000000010000ull;
  float f = seed * 0.999f, g = seed * 1e-3f;
  u64 f2 = ((u64)__float_as_uint(f) << 32) | __float_as_uint(f);
  u64 g2 = ((u64)__float_as_uint(g) << 32) | __float_as_uint(g);
  long long t0 = clk();
  for (int it = 0; it < LITER; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      unsigned addr = base + ((it * 8 + u) & 15) * 64 * VEC;   // warp-uniform slide, keeps pattern alignment (multiple of 16B*VEC/...)
      float x, y, z, ww;
      if (VEC == 4) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(ww) : "r"(addr));
      if (VEC == 2) asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(addr));
      if (VEC == 1) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr));
      if (VEC == 4) { acc = fmaf(x, y, acc); acc2 = fmaf(z, ww, acc2); }
      if (VEC == 2) acc = fmaf(x, y, acc);
      if (VEC == 1) acc += x;
      #pragma unroll
      for (int k = 0; k < NF; k++) p[(u + k) & 7] = fma2(p[(u + k) & 7], f2, g2);
    }
  }
  long long t1 = clk();
  for (int i = 0; i < 8; i++) acc += __uint_as_float((unsigned)p[i]);
  acc += acc2;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __shared__ long long smin, smax;
  if (threadIdx.x == 0) { smin = t0; smax = t1; }
  __syncthreads();
  if (lane == 0) { atomicMin(&smin, t0); atomicMax(&smax, t1); }
  __syncthreads();
  if (lane == 0) cyc[blockIdx.x * (blockDim.x / 32) + w] = smax - smin;
}

static long long maxcyc(long long* d, int n) {
  std::vector<long long> h(n); CK(cudaMemcpy(h.data(), d, n * sizeof(long long), cudaMemcpyDeviceToHost));
  std::sort(h.begin(), h.end()); return h[n / 2];  // median warp
}

int main() {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  int sms = pr.multiProcessorCount;
  printf("device %s sms %d clock %d kHz smem/SM %zu\n", pr.name, sms, pr.clockRate, pr.sharedMemPerMultiprocessor);
  float* out; long long* cyc; CK(cudaMalloc(&out, sms * 1024 * 4)); CK(cudaMalloc(&cyc, sms * 32 * 8));
  const char* fpn[] = {"FFMA", "FFMA2", "FADD2", "FFMA2+FADD2", "FFMA+FADD", "FADD"};
  for (int warps : {4, 8, 12, 16}) {
    for (int m = 0; m < 6; m++) {
      for (int rep = 0; rep < 2; rep++) {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventRecord(e0));
        switch (m) {
          case 0: fp_kernel<0><<<sms, warps * 32>>>(out, cyc, 1.0f); break;
          case 1: fp_kernel<1><<<sms, warps * 32>>>(out, cyc, 1.0f); break;
          case 2: fp_kernel<2><<<sms, warps * 32>>>(out, cyc, 1.0f); break;
          case 3: fp_kernel<3><<<sms, warps * 32>>>(out, cyc, 1.0f); break;
          case 4: fp_kernel<4><<<sms, warps * 32>>>(out, cyc, 1.0f); break;
          case 5: fp_kernel<5><<<sms, warps * 32>>>(out, cyc, 1.0f); break;
        }
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep == 0) continue;
        long long c = maxcyc(cyc, sms * warps);
        double instr = (double)ITER * NACC * ((m == 3 || m == 4) ? 2 : 1);   // warp-instructions per warp
        double laneops = instr * 32 * ((m >= 1 && m <= 3) ? 2 : 1);         // fp32 lane-ops per warp
        printf("FP %-12s warps/SM %2d cycles %9lld  warp-instr/clk/SM %.3f  lane-ops/clk/SM %.1f  ms %.3f  MHz_eff %.0f\n", fpn[m], warps, c,
               instr * warps / c, laneops * warps / c, ms, c / (ms * 1e3));
      }
    }
  }
  const char* pn[] = {"distinct", "pairs(l/2)", "quads(l/4)", "all-same", "l/4+(l&1)", "octs(l/8)", "quads-stride2", "l/4+3*(l&3)"};
  for (int vec : {1, 2, 4}) {
    for (int P = 0; P < 8; P++) {
      for (int warps : {8, 16}) {
        for (int rep = 0; rep < 2; rep++) {
          if (vec == 4) lds_kernel<4, 0><<<sms, warps * 32, 32768>>>(out, cyc, P, 1.0f);
          if (vec == 2) lds_kernel<2, 0><<<sms, warps * 32, 32768>>>(out, cyc, P, 1.0f);
          if (vec == 1) lds_kernel<1, 0><<<sms, warps * 32, 32768>>>(out, cyc, P, 1.0f);
          CK(cudaDeviceSynchronize());
        }
        long long c = maxcyc(cyc, sms * warps);
        double n = (double)LITER * 8 * warps;
        printf("LDS.%-3d %-14s warps/SM %2d cycles %9lld  clk/warp-LDS/SM %.3f  bytes/clk/SM(lane) %.1f\n", vec * 32, pn[P], warps, c, c / n, n * 32 * vec * 4 / c);
      }
    }
  }
  // overlap: LDS.128 distinct + nf FFMA2 per load
  for (int P : {0, 2, 4}) for (int nf : {2, 4, 6, 8, 12, 16}) for (int warps : {8, 12, 16}) {
    for (int rep = 0; rep < 2; rep++) {
      switch (nf) {
        case 2: lds_kernel<4, 2><<<sms, warps * 32, 32768>>>(out, cyc, P, 1.0f); break;
        case 4: lds_kernel<4, 4><<<sms, warps * 32, 32768>>>(out, cyc, P, 1.0f); break;
        case 6: lds_kernel<4, 6><<<sms, warps * 32, 32768>>>(out, cyc, P, 1.0f); break;
        case 8: lds_kernel<4, 8><<<sms, warps * 32, 32768>>>(out, cyc, P, 1.0f); break;
        case 12: lds_kernel<4, 12><<<sms, warps * 32, 32768>>>(out, cyc, P, 1.0f); break;
        case 16: lds_kernel<4, 16><<<sms, warps * 32, 32768>>>(out, cyc, P, 1.0f); break;
      }
      CK(cudaDeviceSynchronize());
    }
    long long c = maxcyc(cyc, sms * warps);
    double n = (double)LITER * 8 * warps;
    printf("MIX LDS.128 %-10s + %2d FFMA2  warps/SM %2d  clk/iter/SM %.3f  (ideal fp %.2f, lds-distinct 4.0)  lane-ops/clk/SM %.1f\n", pn[P], nf, warps, c / n, nf * 2 / 4.0, n * nf * 64 / c);
  }
  printf("done\n");
  return 0;
}
