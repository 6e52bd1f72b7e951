// das_generic.cu -- delay-and-sum for an arbitrary list of directions (one CTA per direction x frame).
//
// This is the general path: any direction list, any channel mask, any frame length.  It serves the
// dynamic-steering (MISO) calls -- T tracked targets, Particle::das + Particle::beam
// (src/dsp/particle.cpp:51-103) -- and the full-grid power map (MIMOWorker::update,
// src/dsp/mimo.cpp:97-151) when the grid does not fit the register-tiled kernel in das_tile.cu.
//
// Arithmetic per (direction, channel, sample) is the reference's delay() triple with every rounding
// pinned (src/dsp/delay.cpp:24): out = out + fma(f, s[i] - s[i+1], s[i+1]), channels in mask order.
#include "bflk_internal.h"

namespace bflk {

constexpr int kGenericThreads = 256;

__global__ void __launch_bounds__(kGenericThreads) das_generic_kernel(GenericArgs a) {
    // [N + 2] delayed sum of this direction | [usable] fraction | [usable] element offset of the channel's first tap.
    // The per-channel table (mask index -> offset, fraction) is resolved once into shared memory: the channel loop
    // then has no dependent global loads (index -> offset -> sample) and its sample loads pipeline across channels.
    extern __shared__ __align__(8) unsigned char s_raw[];
    __shared__ float s_red[kGenericThreads / 32];
    const int d = blockIdx.x;
    const int b = blockIdx.y;
    const int N = a.frame_len;
    long long *s_addr = reinterpret_cast<long long *>(s_raw);
    float *s_frac = reinterpret_cast<float *>(s_raw + sizeof(long long) * a.usable);   // FIR mode: the phase index, as int bits
    float *s_out = s_frac + a.usable;
    const int32_t *off = a.off + (size_t)d * a.C;
    const float *frac = a.frac + (size_t)d * a.C;
    const float *base = a.stream + (size_t)b * a.frame_stride;
    for (int s = threadIdx.x; s < a.usable; s += kGenericThreads) {
        const int c = a.index[s];
        s_addr[s] = (long long)c * a.row_stride + off[c];
        if (a.fir)   // delay.cpp:30-31: float get_filter = fraction * 100.0f + 0.5f; int delay_int = (int) get_filter;
            s_frac[s] = __int_as_float(min(a.fir_phases - 1, (int)__fadd_rn(__fmul_rn(frac[c], (float)(a.fir_phases - 1)), 0.5f)));
        else
            s_frac[s] = frac[c];
    }
    __syncthreads();

    for (int i0 = 0; i0 < N; i0 += kGenericThreads) {
        const int i = i0 + threadIdx.x;
        float acc = 0.0f;
        if (i < N) {
            const float *sig0 = base + i;
            if (a.fir) {
                // 8-tap fractional-delay FIR of the reference's USE_FILTER build (delay.cpp:33-37):
                // out[n] += coeffs[phase][i] * signal[n + i], i = 0..taps-1 in order; product and sum rounded separately
                // (that branch only exists in builds without AVX2 / FMA, delay.cpp:8 -- pinned against such a build of the
                // reference file, oracle/_ref/libref_fir.so)
                for (int s = 0; s < a.usable; s++) {
                    const float *sig = sig0 + s_addr[s];
                    const float *cf = a.fir + (size_t)__float_as_int(s_frac[s]) * a.fir_taps;
#pragma unroll 8
                    for (int t = 0; t < a.fir_taps; t++) acc = __fadd_rn(acc, __fmul_rn(__ldg(cf + t), __ldg(sig + t)));
                }
            } else
#pragma unroll 8
            for (int s = 0; s < a.usable; s++) {
                const float *sig = sig0 + s_addr[s];
                const float f = s_frac[s];
                const float cur = __ldg(sig), nxt = __ldg(sig + 1);
                acc = __fadd_rn(acc, __fmaf_rn(f, __fsub_rn(cur, nxt), nxt));
            }
            s_out[i] = acc;
            if (a.audio) a.audio[((size_t)b * a.n_dir + d) * N + i] = acc;
        }
    }
    if (!a.power) return;
    __syncthreads();
    // 3-tap high-pass + mean square, mimo.cpp:131-137 / particle.cpp:68-79
    float p = 0.0f;
    for (int i = 1 + threadIdx.x; i < N - 1; i += kGenericThreads) {
        float ma = __fsub_rn(__fmul_rn(s_out[i], 0.5f), __fmul_rn(0.25f, __fadd_rn(s_out[i + 1], s_out[i - 1])));
        p = __fmaf_rn(ma, ma, p);
    }
    for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = p;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < kGenericThreads / 32; w++) t += s_red[w];
        a.power[(size_t)b * a.n_dir + d] = __fdiv_rn(t, a.norm);
    }
}

cudaError_t launch_das_generic(const GenericArgs &a, cudaStream_t st) {
    if (a.n_dir <= 0 || a.n_frames <= 0) return cudaSuccess;
    size_t smem = (size_t)(a.frame_len + 2) * sizeof(float) + (size_t)a.usable * (sizeof(long long) + sizeof(float));
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(das_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    for (int b0 = 0; b0 < a.n_frames; b0 += 65535) {   // grid.y is limited to 65535
        GenericArgs s = a;
        s.n_frames = a.n_frames - b0 < 65535 ? a.n_frames - b0 : 65535;
        s.stream = a.stream + (size_t)b0 * a.frame_stride;
        if (a.power) s.power = a.power + (size_t)b0 * a.n_dir;
        if (a.audio) s.audio = a.audio + (size_t)b0 * a.n_dir * a.frame_len;
        dim3 grid(a.n_dir, s.n_frames);
        das_generic_kernel<<<grid, kGenericThreads, smem, st>>>(s);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace bflk
