// cuda_workers.h -- drop-in Worker subclasses that run the reference's delay-and-sum workers on a B200.
//
// Header-only; compile it INSIDE the reference tree (it includes the reference's own worker.h) and link
// libbflk.so.  The classes keep the constructor signatures, virtual overrides, threading and ownership
// contract of the workers they replace:
//   CudaMIMOWorker  <->  MIMOWorker  (src/dsp/mimo.h:36, src/dsp/mimo.cpp:7-13,61-156)
//   CudaMISOWorker  <->  MISOWorker  (src/dsp/miso.h:18, src/dsp/miso.cpp:5-55; tracker heuristics excluded)
// Registration is one line each in AWProcessingUnit::start (src/aw_processing_unit/aw_processing_unit.cpp:73-78),
// see INTEGRATION.md.  All compute goes through the C ABI in include/bflk.h; there is no CPU fallback:
// if the library cannot create a handle the constructor reports it on std::cerr (the reference's error
// convention, src/fpga/pipeline.cpp:31-35) and update() leaves powerdB untouched.
//
// Destruction.  The reference's ~Worker is NOT virtual (src/dsp/worker.h:111-114) and AWProcessingUnit deletes its
// workers through Worker* (aw_processing_unit.cpp:51-53): only ~Worker runs -- it clears `looping` and joins
// thread_loop; the derived destructor and the derived members' destructors never do.  The adapters are correct under
// that: the GPU handle is owned by the worker THREAD (run() = loop() then release()), so it is destroyed before the
// join in ~Worker returns; the derived destructor only covers the case where the object is destroyed as what it is.
// What a delete through Worker* still leaks is host memory of the derived members (two std::vectors), exactly like
// MIMOWorker's own powerdB / offsetDelays do in the reference.
//
// Numerics.  The adapters select the kernel whose delayed sums are bit-identical to delay() (bflk_set_kernel 2): a live
// worker processes one frame per 5.24 ms and has no use for the 13 % the two-FMA form saves in batch throughput.
// -DBFLK_WORKER_AUTOMATIC_KERNEL keeps the library's automatic choice (two-FMA form, power within 1e-4) and lets a thread-block
// cluster split the channels of the frame (bflk_set_channel_split): the lowest latency, not bit-identical sums.
//
// Build modes: by default the reference headers are included; with -DBFLK_STANDIN_HEADERS the minimal
// stand-ins under tests/standin are used instead (Eigen / OpenCV are not installed in the build image).
#pragma once

#include <chrono>
#include <cstring>
#include <iostream>
#include <thread>
#include <vector>

#include "bflk.h"

#ifdef BFLK_STANDIN_HEADERS
#include "reference_standin.h"
#else
#include "worker.h"
#endif

namespace bflk_host {

// Antenna::points is a 3 x n column-major matrix (src/geometry/antenna.h:85); the C ABI wants [n][3].
inline std::vector<float> antenna_xyz(const Antenna &antenna, int n) {
    std::vector<float> xyz(3 * (size_t)n);
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) xyz[3 * i + k] = antenna.points(k, i);
    return xyz;
}

inline bflk_handle *make_handle(const Antenna &antenna, int n_elements, const char *who) {
    bflk_config cfg;
    bflk_default_config(&cfg);  // N_SAMPLES 256, window 1024, 48828 Hz, 340 m/s: the reference's constants
    cfg.n_channels = n_elements;
    bflk_handle *h = nullptr;
    if (bflk_create(&cfg, &h) != BFLK_OK) {
        std::cerr << who << ": " << bflk_last_error(nullptr) << std::endl;
        return nullptr;
    }
    std::vector<float> xyz = antenna_xyz(antenna, n_elements);
    if (bflk_set_geometry(h, xyz.data(), n_elements) != BFLK_OK ||
        (antenna.usable > 0 && bflk_set_channel_mask(h, antenna.index, antenna.usable) != BFLK_OK)) {
        std::cerr << who << ": " << bflk_last_error(h) << std::endl;
        bflk_destroy(h);
        return nullptr;
    }
#ifndef BFLK_WORKER_AUTOMATIC_KERNEL
    bflk_set_kernel(h, 2);  // delayed sums bit-identical to delay() (src/dsp/delay.cpp:16-26)
#else
    bflk_set_channel_split(h, 1);  // a live worker has one frame at a time: latency shape (cfg3 frame 207 -> 89 us)
#endif
    return h;
}

// The snapshot buffer: signals[l] = ring of antenna.index[l] in the reference (mimo.cpp:100-103); the ABI takes the
// snapshot by physical channel, so every stream of the array is copied once: window[c][1024].  Page-locked once, so the
// upload is a single asynchronous DMA instead of a staged pageable copy.
struct Snapshot {
    std::vector<float> window;
    bool pinned = false;
    void take(Streams *streams, int n_elements) {
        if (window.empty()) {
            window.resize((size_t)n_elements * N_ITEMS_BUFFER);
            pinned = bflk_pin_host(window.data(), window.size() * sizeof(float)) == BFLK_OK;
        }
        for (int c = 0; c < n_elements; c++) streams->read_stream(c, &window[(size_t)c * N_ITEMS_BUFFER]);
    }
    void release() {
        if (pinned) bflk_unpin_host(window.data());
        pinned = false;
    }
};

}  // namespace bflk_host

class CudaMIMOWorker : public Worker {
public:
    // how many map peaks become Targets, and how strong (relative to the map's maximum) a peak must be
    static constexpr int kMaxTargets = 8;
    static constexpr float kMinRelPower = 0.5f;

    CudaMIMOWorker(Pipeline *pipeline, Antenna &antenna, bool *running, int rows, int columns, float fov)
        : Worker(pipeline, antenna, running), columns(columns), rows(rows), fov(fov) {
        maxIndex = rows * columns;
        powerdB = std::vector<float>(maxIndex, 0.0);
        handle = bflk_host::make_handle(antenna, ELEMENTS, "CudaMIMOWorker");
        if (handle && bflk_set_grid_fov(handle, rows, columns, fov) != BFLK_OK) {  // computeDelayLUT()
            std::cerr << "CudaMIMOWorker: " << bflk_last_error(handle) << std::endl;
            bflk_destroy(handle);
            handle = nullptr;
        }
        thread_loop = std::thread(&CudaMIMOWorker::run, this);
    }

    ~CudaMIMOWorker() {
        // only reached when the object is destroyed as a CudaMIMOWorker; ~Worker (which always runs) joins the thread,
        // and the thread has released the handle by then -- see the header comment
        looping = false;
        if (thread_loop.joinable()) thread_loop.join();
        release();
        thread_loop = std::thread([] {});  // ~Worker joins unconditionally (worker.h:111-114)
    }

    worker_t get_type() override { return worker_t::MIMO; }

    // direct entry points for tests and offline replay (the live path is loop() -> update())
    void update_once() {
        std::lock_guard<std::mutex> g(handle_lock);
        update();
    }
    const std::vector<float> &power() const { return powerdB; }

protected:
    void reset() override {}
    void setup() override {}

    // MIMOWorker::update (mimo.cpp:97-156) + the Targets TargetHandler polls (worker.h:136-142, target_handler.cpp:29-36)
    void update() override {
        if (!handle) return;
        snap.take(streams, ELEMENTS);
        if (bflk_power_map(handle, snap.window.data(), powerdB.data()) != BFLK_OK) {
            std::cerr << "CudaMIMOWorker: " << bflk_last_error(handle) << std::endl;
            return;
        }
        bflk_target found[kMaxTargets];
        int32_t n = 0;
        if (bflk_targets(handle, nullptr, kMaxTargets, kMinRelPower, found, &n) != BFLK_OK) return;   // the map is still on the device
        const auto now = std::chrono::high_resolution_clock::now();
        std::vector<Target> next;
        next.reserve(n);
        for (int i = 0; i < n; i++) {
            Target t(Spherical(found[i].theta, found[i].phi), found[i].power, found[i].probability, now);
            for (const Target &old : tracking)
                if (old == t) t.start = old.start;   // "time when target was first found" (worker.h:42-43)
            next.push_back(t);
        }
        tracking.swap(next);
    }

    void populateHeatmap(cv::Mat *heatmap) override {
        std::lock_guard<std::mutex> g(handle_lock);
        if (!handle) return;
        std::vector<uint8_t> heat(maxIndex);
        int32_t arg = 0;
        float maxV = 0.f;
        if (bflk_heatmap(handle, powerdB.data(), maxIndex, heat.data(), &arg, &maxV) != BFLK_OK) return;
        float alpha = 0.2;
        prevPower = maxV * alpha + (1 - alpha) * prevPower;  // mimo.cpp:73-74
        int i = 0;
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < columns; c++) heatmap->at<uchar>(r, c) = heat[i++];
    }

private:
    void run() {
        loop();      // Worker::loop (worker.h:212-224): update() under Worker::lock until looping is cleared
        release();   // the worker thread owns the GPU handle: gone before ~Worker's join returns
    }
    void release() {
        std::lock_guard<std::mutex> g(handle_lock);
        if (handle) bflk_destroy(handle);
        handle = nullptr;
        snap.release();
    }

    int maxIndex;
    const int columns;
    const int rows;
    const float fov;
    float prevPower = 1.0;
    bflk_handle *handle = nullptr;
    std::mutex handle_lock;   // update_once() / populateHeatmap() from other threads vs release() at the end of run()
    std::vector<float> powerdB;
    bflk_host::Snapshot snap;
};

class CudaMISOWorker : public Worker {
public:
    CudaMISOWorker(Pipeline *pipeline, Antenna &antenna, bool *running, double fov) : Worker(pipeline, antenna, running), fov(fov) {
        handle = bflk_host::make_handle(antenna, ELEMENTS, "CudaMISOWorker");
        thread_loop = std::thread(&CudaMISOWorker::run, this);
    }

    ~CudaMISOWorker() {
        looping = false;
        if (thread_loop.joinable()) thread_loop.join();
        release();
        thread_loop = std::thread([] {});
    }

    worker_t get_type() override { return worker_t::MISO; }

    // MISOWorker::steer -> startTracking(direction) (miso.cpp:14-19); here the direction is steered directly,
    // the gradient tracker that refines it stays on the host side of the reference
    void steer(Spherical direction) override {
        theta = direction.theta;
        phi = direction.phi;
    }

    void update_once() {
        std::lock_guard<std::mutex> g(handle_lock);
        update();
    }
    const float *audio() const { return data; }   // what AudioWrapper plays (audio_wrapper.cpp:125-143)
    double beam_power() const { return power; }

protected:
    void reset() override {}
    void setup() override {}

    void update() override {
        if (!handle) return;
        snap.take(streams, ELEMENTS);
        float p = 0.f;
        // beamformer.steer(directionCurrent); beamformer.das(&data[0]) (miso.cpp:42-46) + beam() power
        if (bflk_miso(handle, &theta, &phi, 1, snap.window.data(), data, &p) != BFLK_OK)
            std::cerr << "CudaMISOWorker: " << bflk_last_error(handle) << std::endl;
        power = p;
    }

    void populateHeatmap(cv::Mat *) override {}

private:
    void run() {
        loop();
        release();
    }
    void release() {
        std::lock_guard<std::mutex> g(handle_lock);
        if (handle) bflk_destroy(handle);
        handle = nullptr;
        snap.release();
    }

    double fov;
    double theta = 0.0, phi = 0.0;
    float data[N_SAMPLES] = {0};
    double power = 0.0;
    bflk_handle *handle = nullptr;
    std::mutex handle_lock;
    bflk_host::Snapshot snap;
};
