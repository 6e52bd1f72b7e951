// tile_bench.cu -- issue / shared-memory micro-benchmark of the generated channel loops (tools/gen_tile_asm.py).
//
// Runs the PTX stage functions of das_tile on synthetic shared-memory rows and table entries (random but valid window
// offsets and deltas), 16 warps per SM, no TMA, no epilogue: what remains is exactly the instruction stream of the channel
// loop.  Prints, per variant, the fraction of the FFMA2 issue peak (64 FFMA2 per (warp, channel) x 2 port cycles over the
// elapsed cycles of the SM's four schedulers).  Used to choose between tilings before building the tables / pack / epilogue
// around them.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tile_bench tile_bench.cu [-DWIDE_INC=\"...\"]
#include <cuda_runtime.h>

#include <algorithm>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

typedef unsigned long long u64;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

#ifndef EXACT_INC
#define EXACT_INC "../../beamforming-lk_b200/csrc/das_tile_asm.inc"
#endif
#include EXACT_INC
#include "../../beamforming-lk_b200/csrc/das_tile_fast_asm.inc"
#ifndef WIDE_INC
#define WIDE_INC "wide_default.inc"   // python ../gen_tile_asm.py --wide > wide_default.inc
#endif
#include WIDE_INC
#include "k16.inc"   // python ../gen_tile_asm.py --k16 > k16.inc
#include "k6.inc"    // python ../gen_tile_asm.py --k6 > k6.inc

enum Variant { FAST_SINGLE = 0, FAST_DUAL = 1, WIDE = 2, WIDE_EXACT = 3, EXACT_SINGLE = 4, EXACT_DUAL = 5, K16_SINGLE = 6, K16_DUAL = 7, K6_DUAL = 8, K6_SINGLE = 9 };

struct Args {
    const char *entries;   // [n_sets][warps][cc][ent]
    int n_sets, cc, ent_bytes, row_bytes, rows_bytes, iters, lane_stride;
    float *out;
    long long *cyc;
};

template <int V, int NCH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) bench_kernel(Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t smem = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // rows: a.rows_bytes of floats; entries after them
    float *rows = reinterpret_cast<float *>(smem_raw);
    for (int i = threadIdx.x; i < a.rows_bytes / 4; i += blockDim.x) rows[i] = 1e-3f * (float)((i * 2654435761u) >> 20);
    const int set_bytes = WARPS * a.cc * a.ent_bytes;
    unsigned char *ents = smem_raw + a.rows_bytes;
    for (int i = threadIdx.x; i < a.n_sets * set_bytes / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(ents)[i] = reinterpret_cast<const uint32_t *>(a.entries)[i];
    __syncthreads();
    constexpr int NACC = (V == K16_SINGLE || V == K16_DUAL) ? 64 : (V == K6_DUAL || V == K6_SINGLE) ? 24 : 32;
    u64 acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = 0ull;
    const uint32_t rows_s = smem, ents_s = smem + a.rows_bytes;
    long long t0, t1;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0));
    int set = 0;
    for (int it = 0; it < a.iters; it++) {
        const uint32_t tiles_s = ents_s + set * set_bytes + warp * a.cc * a.ent_bytes;
        const uint32_t row = rows_s + a.lane_stride * lane;
        if constexpr (V == FAST_SINGLE) tile_stage_fast<NCH>(*reinterpret_cast<u64(*)[4][8]>(acc), tiles_s, row, tiles_s + a.cc * a.ent_bytes);
        if constexpr (V == FAST_DUAL) tile_stage_fast_dual<NCH>(*reinterpret_cast<u64(*)[4][8]>(acc), tiles_s, row, tiles_s + a.cc * a.ent_bytes);
        if constexpr (V == WIDE) tile_stage_wide<NCH>(*reinterpret_cast<u64(*)[2][16]>(acc), tiles_s, row, tiles_s + a.cc * a.ent_bytes);
        if constexpr (V == K16_SINGLE) tile_stage_k16<NCH>(*reinterpret_cast<u64(*)[4][16]>(acc), tiles_s, row, tiles_s + a.cc * a.ent_bytes);
        if constexpr (V == K16_DUAL) tile_stage_k16_dual<NCH>(*reinterpret_cast<u64(*)[4][16]>(acc), tiles_s, row, tiles_s + a.cc * a.ent_bytes);
        if constexpr (V == K6_DUAL) tile_stage_k6_dual(*reinterpret_cast<u64(*)[4][6]>(acc), tiles_s, row, tiles_s + a.cc * a.ent_bytes);
        if constexpr (V == K6_SINGLE) tile_stage_k6(*reinterpret_cast<u64(*)[4][6]>(acc), tiles_s, row, tiles_s + a.cc * a.ent_bytes);
        if constexpr (V == WIDE_EXACT) tile_stage_wide_exact<NCH>(*reinterpret_cast<u64(*)[2][16]>(acc), tiles_s, row, tiles_s + a.cc * a.ent_bytes);
        if constexpr (V == EXACT_SINGLE || V == EXACT_DUAL) {
            uint32_t e0, e1;
            float f0, f1, f2, f3;
            uint32_t ent = tiles_s, r = row;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e0), "=r"(e1) : "r"(ent));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f0), "=f"(f1), "=f"(f2), "=f"(f3) : "r"(ent + 16));
#pragma unroll 1
            for (int c = 0; c < a.cc; c++) {
                ent += 32;
                if constexpr (V == EXACT_DUAL) tile_channel_step_dual<NCH>(*reinterpret_cast<u64(*)[4][8]>(acc), e0, e1, f0, f1, f2, f3, r, ent);
                else tile_channel_step<NCH>(*reinterpret_cast<u64(*)[4][8]>(acc), e0, e1, f0, f1, f2, f3, r, ent);
                r += a.row_bytes;
            }
        }
        if (++set == a.n_sets) set = 0;
    }
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1));
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)(acc[i] >> 32));
    a.out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    __shared__ long long smin, smax;
    if (threadIdx.x == 0) { smin = t0; smax = t1; }
    __syncthreads();
    if (lane == 0) { atomicMin(&smin, t0); atomicMax(&smax, t1); }
    __syncthreads();
    if (threadIdx.x == 0) a.cyc[blockIdx.x] = smax - smin;
}

static int padded4(int c) { return c + (c >> 2); }
static int padded8(int c) { return c + (c >> 3); }

// span distribution of cfg3 (oracle LUT, SURVEY 8d geometry): direction pairs along the short axis 0..3, 2x2 tiles 0..9
static const double kPairSpan[4] = {0.246, 0.377, 0.297, 0.080};
static const double kTileSpanFine[2] = {0.65, 0.35};

template <int V, int NCH, int WARPS>
static void run(const char *name, int sms, int max_q, double same_window_share, const double *span_p, int n_span) {
    constexpr bool k16 = V == K16_SINGLE || V == K16_DUAL;
    constexpr bool k6 = V == K6_DUAL || V == K6_SINGLE;
    const int cc = (V == WIDE || V == WIDE_EXACT || k16) ? 4 : 8;
    const int lane_chunks = (V == WIDE || V == WIDE_EXACT || k16) ? 8 : (k6 ? 3 : 4);
    const int row_chunks = max_q + lane_chunks * 31 + NCH + 1;
    const int copy_bytes = k6 ? 16 * row_chunks : 16 * ((lane_chunks == 8 ? padded8(row_chunks - 1) : padded4(row_chunks - 1)) + 1);
    const int row_bytes = 2 * copy_bytes;
    const int ent_bytes = k6 ? 48 : V == FAST_SINGLE ? 64 : V == FAST_DUAL ? 80 : (V == WIDE || V == WIDE_EXACT) ? 48 : V == K16_SINGLE ? 80 : V == K16_DUAL ? 112 : 32;
    const int n_sets = 8;
    std::mt19937 rng(12345);
    auto pick_span = [&]() {
        double u = std::uniform_real_distribution<double>(0, 1)(rng), s = 0;
        for (int i = 0; i < n_span; i++) { s += span_p[i]; if (u < s) return i; }
        return n_span - 1;
    };
    std::vector<unsigned char> ents((size_t)n_sets * WARPS * cc * ent_bytes, 0);
    for (int set = 0; set < n_sets; set++)
        for (int w = 0; w < WARPS; w++)
            for (int c = 0; c < cc; c++) {
                unsigned char *e = &ents[(((size_t)set * WARPS + w) * cc + c) * ent_bytes];
                auto window = [&](int &odd, int &q) { odd = rng() & 1; q = rng() % (max_q + 1); };
                float fr[4];
                for (int k = 0; k < 4; k++) fr[k] = std::uniform_real_distribution<float>(0, 1)(rng);
                if (k6) {
                    uint32_t o2[2] = {0, 0}, dl = 0;
                    float g[4];
                    for (int k = 0; k < 4; k++) g[k] = 1.0f - fr[k];
                    for (int w2 = 0; w2 < (V == K6_DUAL ? 2 : 1); w2++) {
                        int odd, q; window(odd, q);
                        o2[w2] = c * row_bytes + odd * copy_bytes + 16 * q;
                        const int sp = pick_span();
                        if (V == K6_DUAL) dl |= ((rng() & 1) ? (uint32_t)sp : ((uint32_t)sp << 6)) << (12 * w2);
                        else {
                            const int zero_slot = rng() & 3;
                            for (int k = 0; k < 4; k++) dl |= (uint32_t)(k == zero_slot ? 0 : (sp ? rng() % (sp + 1) : 0)) << (6 * k);
                            if (sp) dl = (dl & ~(63u << (6 * ((zero_slot + 1) & 3)))) | ((uint32_t)sp << (6 * ((zero_slot + 1) & 3)));
                        }
                    }
                    if (V == K6_DUAL && std::uniform_real_distribution<double>(0, 1)(rng) < same_window_share) dl |= 1u << 28;
                    memcpy(e, o2, 8); memcpy(e + 8, &dl, 4); memcpy(e + 16, fr, 16); memcpy(e + 32, g, 16);
                } else if (k16) {
                    uint32_t oa[8], ob[8], dl = 0;
                    float g[4];
                    for (int k = 0; k < 4; k++) g[k] = 1.0f - fr[k];
                    for (int w2 = 0; w2 < (V == K16_DUAL ? 2 : 1); w2++) {
                        int odd, q; window(odd, q);
                        const int r = q & 7;
                        for (int k = 0; k < 8; k++) (w2 ? ob : oa)[k] = c * row_bytes + odd * copy_bytes + 16 * padded8(q) + ((k > 0 && r + k >= 8) ? 16 : 0);
                        const int sp = pick_span();
                        if (V == K16_DUAL) dl |= ((rng() & 1) ? (uint32_t)sp : ((uint32_t)sp << 6)) << (12 * w2);
                        else {
                            const int zero_slot = rng() & 3;
                            for (int k = 0; k < 4; k++) dl |= (uint32_t)(k == zero_slot ? 0 : (sp ? rng() % (sp + 1) : 0)) << (6 * k);
                            if (sp) dl = (dl & ~(63u << (6 * ((zero_slot + 1) & 3)))) | ((uint32_t)sp << (6 * ((zero_slot + 1) & 3)));
                        }
                    }
                    if (V == K16_DUAL && std::uniform_real_distribution<double>(0, 1)(rng) < same_window_share) dl |= 1u << 28;
                    if (V == K16_DUAL) { memcpy(e, oa, 32); memcpy(e + 32, ob, 32); memcpy(e + 64, fr, 16); memcpy(e + 80, g, 16); memcpy(e + 96, &dl, 4); }
                    else { memcpy(e, oa, 32); memcpy(e + 32, fr, 16); memcpy(e + 48, g, 16); memcpy(e + 64, &dl, 4); }
                } else if (V == WIDE || V == WIDE_EXACT) {
                    int odd, q; window(odd, q);
                    const int r = q & 7, sp = pick_span();
                    uint32_t o[8];
                    for (int k = 0; k < 8; k++) o[k] = c * row_bytes + odd * copy_bytes + 16 * padded8(q) + ((k > 0 && r + k >= 8) ? 16 : 0);
                    uint32_t dl = (rng() & 1) ? (uint32_t)sp : ((uint32_t)sp << 6);
                    memcpy(e, o, 32); memcpy(e + 32, &dl, 4); memcpy(e + 36, fr, 8);
                } else if (V == FAST_SINGLE) {
                    int odd, q; window(odd, q);
                    const int r = q & 3;
                    uint32_t o[4];
                    for (int k = 0; k < 4; k++) o[k] = c * row_bytes + odd * copy_bytes + 16 * padded4(q) + ((k > 0 && r >= 4 - k) ? 16 : 0);
                    const int sp = pick_span();
                    uint32_t dl = 0;
                    const int zero_slot = rng() & 3;
                    for (int k = 0; k < 4; k++) dl |= (uint32_t)(k == zero_slot ? 0 : (sp ? rng() % (sp + 1) : 0)) << (6 * k);
                    if (sp) dl = (dl & ~(63u << (6 * ((zero_slot + 1) & 3)))) | ((uint32_t)sp << (6 * ((zero_slot + 1) & 3)));
                    if (sp >= 2 * NCH - 10) dl |= 1u << 29;   // a delta reaches the window's last chunk
                    float g[4];
                    for (int k = 0; k < 4; k++) g[k] = 1.0f - fr[k];
                    memcpy(e, o, 16); memcpy(e + 16, &dl, 4); memcpy(e + 32, fr, 16); memcpy(e + 48, g, 16);
                } else if (V == FAST_DUAL) {
                    uint32_t oa[4], ob[4], dl = 0;
                    for (int w2 = 0; w2 < 2; w2++) {
                        int odd, q; window(odd, q);
                        const int r = q & 3;
                        for (int k = 0; k < 4; k++) (w2 ? ob : oa)[k] = c * row_bytes + odd * copy_bytes + 16 * padded4(q) + ((k > 0 && r >= 4 - k) ? 16 : 0);
                        const int sp = pick_span();
                        dl |= ((rng() & 1) ? (uint32_t)sp : ((uint32_t)sp << 6)) << (12 * w2);
                        if (sp >= 2 * NCH - 10) dl |= 1u << (29 + w2);
                    }
                    if (std::uniform_real_distribution<double>(0, 1)(rng) < same_window_share) {
                        dl |= 1u << 28;
                        if (dl & (1u << 30)) dl = (dl & ~(1u << 30)) | (1u << 29);   // all four deltas refer to window A
                    }
                    float g[4];
                    for (int k = 0; k < 4; k++) g[k] = 1.0f - fr[k];
                    memcpy(e, oa, 16); memcpy(e + 16, ob, 16); memcpy(e + 32, fr, 16); memcpy(e + 48, g, 16); memcpy(e + 64, &dl, 4);
                } else {   // exact narrow: win_off (two 16-bit halves in dual mode), deltas | pad phase, span, -, frac[4]
                    uint32_t win = 0, dl = 0;
                    for (int w2 = 0; w2 < (V == EXACT_DUAL ? 2 : 1); w2++) {
                        int odd, q; window(odd, q);
                        win |= (uint32_t)(odd * copy_bytes + 16 * padded4(q)) << (16 * w2);
                        dl |= (uint32_t)(q & 3) << (24 + 2 * w2);
                        const int sp = pick_span();
                        if (V == EXACT_DUAL) dl |= ((rng() & 1) ? (uint32_t)sp : ((uint32_t)sp << 6)) << (12 * w2);
                        else for (int k = 0; k < 4; k++) dl |= (uint32_t)(k == 0 ? 0 : (k == 1 ? sp : (sp ? rng() % (sp + 1) : 0))) << (6 * k);
                    }
                    if (V == EXACT_DUAL && std::uniform_real_distribution<double>(0, 1)(rng) < same_window_share) dl |= 1u << 28;
                    memcpy(e, &win, 4); memcpy(e + 4, &dl, 4); memcpy(e + 16, fr, 16);
                }
            }
    char *d_ents; float *d_out; long long *d_cyc;
    CK(cudaMalloc(&d_ents, ents.size() + 256)); CK(cudaMemset(d_ents, 0, ents.size() + 256));
    CK(cudaMemcpy(d_ents, ents.data(), ents.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_out, (size_t)sms * WARPS * 32 * 4)); CK(cudaMalloc(&d_cyc, sms * 8));
    Args a{};
    a.entries = d_ents; a.n_sets = n_sets; a.cc = cc; a.ent_bytes = ent_bytes; a.row_bytes = row_bytes;
    a.rows_bytes = cc * row_bytes; a.iters = 4096 / cc; a.lane_stride = lane_chunks == 8 ? 144 : (k6 ? 48 : 80); a.out = d_out; a.cyc = d_cyc;
    const size_t smem = (size_t)a.rows_bytes + ents.size() + 256;
    if (smem > 227 * 1024) { printf("%-28s does not fit shared memory (%zu B)\n", name, smem); return; }
    CK(cudaFuncSetAttribute(bench_kernel<V, NCH, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int rep = 0; rep < 2; rep++) { bench_kernel<V, NCH, WARPS><<<sms, WARPS * 32, smem>>>(a); CK(cudaDeviceSynchronize()); }
    std::vector<long long> cyc(sms);
    CK(cudaMemcpy(cyc.data(), d_cyc, sms * 8, cudaMemcpyDeviceToHost));
    std::sort(cyc.begin(), cyc.end());
    const double c = (double)cyc[sms / 2], steps = (double)a.iters * cc;
    const double frac = steps * (k16 ? 128 : (k6 ? 48 : 64)) * 2 * (WARPS / 4.0) / c;
    printf("%-28s nch %2d warps %2d row %5d B  cycles/channel-step/warp %.1f  fraction of FFMA2 issue peak %.3f\n", name, NCH, WARPS,
           row_bytes, c / steps, frac);
    cudaFree(d_ents); cudaFree(d_out); cudaFree(d_cyc);
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    printf("device %s, %d SMs\n", pr.name, sms);
    const double span1[2] = {0.65, 0.35}, span2[3] = {0.55, 0.45, 0.0}, span3[4] = {0.15, 0.35, 0.32, 0.18};
    (void)span2; (void)kTileSpanFine;
    // fine grids (cfg1 / cfg5): one window per 2x2 tile
    run<FAST_SINGLE, 5, 16>("fast single (cfg1-like)", sms, 14, 0, span1, 2);
    run<FAST_SINGLE, 6, 16>("fast single (span 0..3)", sms, 49, 0, span3, 4);
    run<FAST_SINGLE, 6, 16>("fast single (cfg5 spans)", sms, 49, 0, span2, 3);
    run<EXACT_SINGLE, 5, 16>("exact single (cfg1-like)", sms, 14, 0, span1, 2);
    run<EXACT_SINGLE, 6, 16>("exact single (cfg5-like)", sms, 49, 0, span3, 4);
    // coarse grid (cfg3): two windows per 2x2 tile today vs one direction pair x 16 sample pairs per lane
    run<FAST_DUAL, 6, 16>("fast dual (cfg3-like)", sms, 49, 0.5, kPairSpan, 4);
    run<K6_DUAL, 5, 20>("k6 dual 20 warps (cfg3-like)", sms, 49, 0.5, kPairSpan, 4);
    run<K6_DUAL, 5, 16>("k6 dual 16 warps (cfg3-like)", sms, 49, 0.5, kPairSpan, 4);
    run<K6_DUAL, 5, 20>("k6 dual 20 warps, 1 window", sms, 49, 1.0, kPairSpan, 4);
    run<K6_SINGLE, 5, 20>("k6 single 20 warps (span 0..1)", sms, 49, 0, span1, 2);
    run<FAST_DUAL, 6, 16>("fast dual, always 1 window", sms, 49, 1.0, kPairSpan, 4);
    run<FAST_DUAL, 6, 16>("fast dual, always 2 windows", sms, 49, 0.0, kPairSpan, 4);
    run<EXACT_DUAL, 6, 16>("exact dual (cfg3-like)", sms, 49, 0.5, kPairSpan, 4);
    run<WIDE, 10, 16>("wide pair (cfg3-like)", sms, 49, 0, kPairSpan, 4);
    run<WIDE_EXACT, 10, 16>("wide pair exact (cfg3-like)", sms, 49, 0, kPairSpan, 4);
    run<K16_DUAL, 10, 8>("k16 2x2 dual (cfg3-like)", sms, 49, 0.5, kPairSpan, 4);
    run<K16_DUAL, 10, 10>("k16 2x2 dual (cfg3-like)", sms, 49, 0.5, kPairSpan, 4);
    run<K16_SINGLE, 9, 10>("k16 2x2 single (cfg1-like)", sms, 14, 0, span1, 2);
    run<K16_SINGLE, 10, 10>("k16 2x2 single (cfg5-like)", sms, 49, 0, span3, 4);
    run<K16_SINGLE, 9, 8>("k16 2x2 single (cfg1-like)", sms, 14, 0, span1, 2);
    run<K16_SINGLE, 10, 8>("k16 2x2 single (cfg5-like)", sms, 49, 0, span3, 4);
    run<WIDE, 9, 16>("wide pair (cfg1-like)", sms, 14, 0, span1, 2);
    run<WIDE_EXACT, 9, 16>("wide pair exact (cfg1-like)", sms, 14, 0, span1, 2);
    run<WIDE, 10, 16>("wide pair (cfg5-like)", sms, 49, 0, span3, 4);
    printf("done\n");
    return 0;
}
