// coresident.cu -- does a small CTA (128 threads, <= 32 registers) get an SM slot next to a resident 512-thread CTA that
// holds 120 registers per thread and ~200 KB of shared memory?  Kernel A spins ~2 ms on every SM; kernel B is launched on
// a second stream 200 us later and stamps %globaltimer.  Prints when B ran relative to A, per register budget of A.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <int REGS>
__global__ void __launch_bounds__(512, 1) big(unsigned long long *t, unsigned long long spin_ns, float *sink) {
    extern __shared__ float sm[];
    float acc[REGS];
#pragma unroll
    for (int i = 0; i < REGS; i++) acc[i] = threadIdx.x * 0.001f + i;
    const unsigned long long t0 = gtime();
    while (gtime() - t0 < spin_ns) {
#pragma unroll
        for (int i = 0; i < REGS; i++) acc[i] = acc[i] * 1.0001f + sm[(threadIdx.x + i) & 1023];
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < REGS; i++) s += acc[i];
    sink[blockIdx.x * 512 + threadIdx.x] = s;
    if (threadIdx.x == 0) { t[2 * blockIdx.x] = t0; t[2 * blockIdx.x + 1] = gtime(); }
}

__global__ void __launch_bounds__(128) small(unsigned long long *t, float *sink) {
    float acc[18];
#pragma unroll
    for (int i = 0; i < 18; i++) acc[i] = threadIdx.x * 0.5f + i;
    const unsigned long long t0 = gtime();
    while (gtime() - t0 < 20000) {
#pragma unroll
        for (int i = 0; i < 18; i++) acc[i] = acc[i] * 1.0001f + 0.5f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 18; i++) s += acc[i];
    if (s == 12345.f) sink[threadIdx.x] = s;
    if (threadIdx.x == 0) { t[2 * blockIdx.x] = t0; t[2 * blockIdx.x + 1] = gtime(); }
}

template <int REGS>
static void run(int sms, int smem, bool carve) {
    unsigned long long *ta, *tb; float *sink;
    CK(cudaMalloc(&ta, sms * 16)); CK(cudaMalloc(&tb, sms * 4 * 16)); CK(cudaMalloc(&sink, sms * 512 * 4));
    cudaStream_t a, b; CK(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking));
    CK(cudaFuncSetAttribute(big<REGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (carve) CK(cudaFuncSetAttribute(small, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, big<REGS>));
    cudaFuncAttributes fb; CK(cudaFuncGetAttributes(&fb, small));
    for (int rep = 0; rep < 2; rep++) {
        big<REGS><<<sms, 512, smem, a>>>(ta, 2000000ull, sink);
        small<<<sms * 4, 128, 0, b>>>(tb, sink);
        CK(cudaDeviceSynchronize());
    }
    unsigned long long ha[2], *hb = (unsigned long long *)malloc(sms * 4 * 16);
    CK(cudaMemcpy(ha, ta, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hb, tb, sms * 4 * 16, cudaMemcpyDeviceToHost));
    unsigned long long first = ~0ull, last = 0;
    for (int i = 0; i < sms * 4; i++) { if (hb[2 * i] < first) first = hb[2 * i]; if (hb[2 * i + 1] > last) last = hb[2 * i + 1]; }
    printf("big: %d regs, %d B smem, carveout hint %d; small: %d regs.  big ran 0 .. %.0f us; small CTAs ran %.0f .. %.0f us -> %s\n", fa.numRegs, smem,
           (int)carve, fb.numRegs, (ha[1] - ha[0]) / 1e3, ((double)first - (double)ha[0]) / 1e3, ((double)last - (double)ha[0]) / 1e3,
           last < ha[1] ? "CO-RESIDENT" : "after big");
    cudaFree(ta); cudaFree(tb); cudaFree(sink); free(hb);
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    printf("%s, %d SMs, regs/SM %d, smem/SM %zu\n", pr.name, sms, pr.regsPerMultiprocessor, pr.sharedMemPerMultiprocessor);
    run<60>(sms, 202116, true);
    run<64>(sms, 202116, true);
    run<66>(sms, 202116, true);
    run<68>(sms, 202116, true);
    run<70>(sms, 202116, true);
    run<72>(sms, 202116, true);
    run<74>(sms, 202116, true);
    return 0;
}
