// Register-file operand bandwidth of FFMA2 on sm_100a: does an FFMA2 with TWO fresh 64-bit register operands
// (window pair + accumulator pair) still issue every 2 cycles, and does it depend on which registers they are?
// The window registers are filled by ld.shared.v2.b64 (LDS.128 -> aligned register quads), as in das_tile.
//   mode 0: acc[k] += g*w[k+1]; acc[k] += f*w[k]        (the two-FMA body, delta 0)
//   mode 1: acc[k] += g*w[k+2]; acc[k] += f*w[k]        (both window operands of an accumulator have the same index parity)
//   mode 2: acc[k] += g*w[1];   acc[k] += f*w[0]        (window operands fixed -> reuse cache)
//   mode 3: mode 0 with delta 1
// Prints cycles per FFMA2 per scheduler (4 warps per scheduler) and lets cuobjdump show the register numbers.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }
#define ITER 4096
template <int MODE>
__global__ void __launch_bounds__(512, 1) rf_kernel(float *out, long long *cyc, float seed) {
    __shared__ __align__(16) u64 sm[512 * 12 / 8 + 64];
    for (int i = threadIdx.x; i < 512 * 12 / 8 + 64; i += blockDim.x) sm[i] = 0x3f8000003f800000ull + i;
    __syncthreads();
    u64 acc[8], w[12];
    for (int k = 0; k < 8; k++) acc[k] = 0;
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16;
    float f = seed * 0.25f, g = 1.0f - f;
    u64 f2, g2;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(f2) : "f"(f));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(g2) : "f"(g));
    unsigned ix = threadIdx.x, iy = blockIdx.x;
    long long t0 = clk();
    for (int it = 0; it < ITER; it++) {
        if ((it & 63) == 0) {
#pragma unroll
            for (int m = 0; m < 6; m++)
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(w[2 * m]), "=l"(w[2 * m + 1]) : "r"(base + 512 * m + ((it >> 6) & 3) * 16));
        }
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int j = MODE == 0 ? k + 1 : MODE == 1 ? k + 2 : (MODE == 2 || MODE >= 4) ? 1 : k + 2;
                acc[k] = fma2(g2, w[j], acc[k]);
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int j = MODE == 0 ? k : MODE == 1 ? k : (MODE == 2 || MODE >= 4) ? 0 : k + 1;
                acc[k] = fma2(f2, w[j], acc[k]);
                // MODE 4 / 5: one / two independent integer instructions per FFMA2 -- do they issue in its shadow?
                if (MODE >= 4) asm volatile("lop3.b32 %0, %0, %1, 0x55555555, 0x96;" : "+r"(ix) : "r"(iy));
                if (MODE >= 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(iy) : "r"(ix));
            }
        }
    }
    long long t1 = clk();
    float s = __uint_as_float(ix);
    for (int k = 0; k < 8; k++) s += __uint_as_float((unsigned)acc[k]) + __uint_as_float((unsigned)(acc[k] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 16 + threadIdx.x / 32] = t1 - t0;
}
template <int MODE>
void run(const char *name, float *out, long long *cyc, int sms) {
    rf_kernel<MODE><<<sms, 512>>>(out, cyc, 1.0f);
    CK(cudaDeviceSynchronize());
    rf_kernel<MODE><<<sms, 512>>>(out, cyc, 1.0f);
    CK(cudaDeviceSynchronize());
    long long h[16];
    CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < 16; i++) mx = h[i] > mx ? h[i] : mx;
    // 4 warps per scheduler, 64 FFMA2 per iteration per warp
    printf("%-40s cycles %lld  cycles/FFMA2/scheduler %.3f\n", name, mx, (double)mx / (4.0 * 64 * ITER));
}
int main() {
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    float *out; long long *cyc;
    CK(cudaMalloc(&out, pr.multiProcessorCount * 512 * 4));
    CK(cudaMalloc(&cyc, pr.multiProcessorCount * 16 * 8));
    run<0>("two-FMA body delta 0 (w[k+1], w[k])", out, cyc, pr.multiProcessorCount);
    run<3>("two-FMA body delta 1 (w[k+2], w[k+1])", out, cyc, pr.multiProcessorCount);
    run<1>("same-parity window operands (w[k+2], w[k])", out, cyc, pr.multiProcessorCount);
    run<2>("fixed window operands (reuse)", out, cyc, pr.multiProcessorCount);
    run<4>("fixed operands + 0.5 LOP3 per FFMA2", out, cyc, pr.multiProcessorCount);
    run<5>("fixed operands + 0.5 LOP3 + 0.5 IADD per FFMA2", out, cyc, pr.multiProcessorCount);
    return 0;
}
