// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
// Thin extern "C" harness around the UNMODIFIED reference sources, compiled where they lie:
//   /root/reference/src/dsp/delay.cpp   (delay(), AVX2 variant, delay.cpp:16-26)
//   /root/reference/src/fpga/streams.hpp (Streams ring buffers, streams.hpp:54-183)
// It restates only the loop nest of MIMOWorker::update (src/dsp/mimo.cpp:97-151) and
// Particle::beam/das (src/dsp/particle.cpp:51-103) around the real delay(); built by
// oracle/Makefile into oracle/_ref/libref.so.  Used to pin oracle.c and as the "reference" CPU
// baseline in bench.py.  Never part of the product.
#include "delay.h"
#include "streams.hpp"

#include <cstring>
#include <thread>
#include <vector>

extern "C" {

int ref_n_samples() { return N_SAMPLES; }

// 1 when this object holds the FIR variant of delay() (built without AVX2, delay.cpp:8,28), and the reference's
// coefficient table (src/dsp/filter.h:10-112) as that build sees it
#if !defined(__AVX2__) && USE_FILTER
int ref_is_fir() { return 1; }
const float *ref_filter_coeffs() { return &filter_coeffs[0][0]; }
#else
int ref_is_fir() { return 0; }
const float *ref_filter_coeffs() { return nullptr; }
#endif

void ref_delay(float *out, const float *signal, float fraction) { delay(out, signal, fraction); }

// Push n_frames blocks of N_SAMPLES floats through one real Streams ring (write_stream + forward,
// like Pipeline::producer, pipeline.cpp:243-255 / :294-296) and return what a worker sees:
// read_stream(0, window) (mimo.cpp:100-103) and get_signal(0, probe_offset)[0..N_SAMPLES].
int ref_streams_window(const float *frames, int n_frames, float *window, int probe_offset, float *probe) {
    Streams streams;
    if (!streams.create_stream(0)) return -1;
    for (int f = 0; f < n_frames; f++) {
        streams.write_stream(0, const_cast<float *>(frames + (size_t)f * N_SAMPLES));
        streams.forward();
    }
    streams.read_stream(0, window);
    const float *p = streams.get_signal(0, probe_offset);
    std::memcpy(probe, p, sizeof(float) * (N_SAMPLES + 1));
    return (int)(N_ITEMS_BUFFER);
}

static void mimo_range(const float *window, int C, int W, int n, const int *index, int usable,
                       const int *offsets, const float *fractions, int d0, int d1, float *power, float *das_out) {
    std::vector<float> outv(n + 8);
    float *out = outv.data();
    for (int m = d0; m < d1; m++) {
        std::memset(out, 0, sizeof(float) * n);
        int count = 0;
        for (int s = 0; s < usable; s++) {
            int i = index[s];
            float fraction = fractions[(size_t)m * C + i];
            int offset = offsets[(size_t)m * C + i];
            const float *sig = &window[(size_t)i * W + offset];
            for (int b = 0; b < n; b += N_SAMPLES) delay(&out[b], sig + b, fraction);
            count++;
        }
        if (das_out) std::memcpy(&das_out[(size_t)m * n], out, sizeof(float) * n);
        float p = 0.0f;
        for (int i = 1; i < n - 1; i++) {
            float MA = out[i] * 0.5f - 0.25f * (out[i + 1] + out[i - 1]);
            p += MA * MA;
        }
        p /= static_cast<float>(n * count);
        if (power) power[m] = p;
    }
}

// MIMOWorker::update loop nest (mimo.cpp:121-150) over directions [0, D) split across n_threads
// std::threads (1 == the reference's real per-worker concurrency, worker.h:90).  n must be a
// multiple of N_SAMPLES.
void ref_mimo_update(const float *window, int C, int W, int n, const int *index, int usable,
                     const int *offsets, const float *fractions, int D, float *power, float *das_out, int n_threads) {
    if (n_threads <= 1) { mimo_range(window, C, W, n, index, usable, offsets, fractions, 0, D, power, das_out); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) {
        int d0 = (int)((long long)D * t / n_threads), d1 = (int)((long long)D * (t + 1) / n_threads);
        th.emplace_back(mimo_range, window, C, W, n, index, usable, offsets, fractions, d0, d1, power, das_out);
    }
    for (auto &t : th) t.join();
}
}
