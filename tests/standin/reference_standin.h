// reference_standin.h -- TEST INFRASTRUCTURE ONLY.
// Minimal stand-ins for the reference declarations beamforming-lk_b200/host/cuda_workers.h touches, so the
// adapter can be compiled and exercised in an image without Eigen / OpenCV.  Shapes follow
// src/dsp/worker.h:66-233, src/fpga/streams.hpp:54-139, src/fpga/pipeline.h:40-107,
// src/geometry/antenna.h:80-103, src/geometry/geometry.h (Spherical); nothing here is copied code: the
// stand-ins only reproduce names, member order and call contracts.
#pragma once
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#define N_SAMPLES 256
#define N_ITEMS_BUFFER 1024
#define ELEMENTS 64

typedef unsigned char uchar;
namespace cv {
struct Mat {
    int rows = 0, cols = 0;
    std::vector<uchar> store;
    Mat(int r, int c) : rows(r), cols(c), store((size_t)r * c, 0) {}
    template <typename T> T &at(int r, int c) { return reinterpret_cast<T &>(store[(size_t)r * cols + c]); }
};
}  // namespace cv

struct Spherical {
    double theta = 0, phi = 0, radius = 1;
    Spherical() {}
    Spherical(double t, double p) : theta(t), phi(p) {}
};

struct PointsStandin {  // Eigen::MatrixXf(3, n) look-alike: operator()(row, col)
    int n = 0;
    std::vector<float> v;
    float operator()(int r, int c) const { return v[(size_t)3 * c + r]; }
};

struct Antenna {
    PointsStandin points;
    int id = 0;
    int usable = 0;
    int *index = nullptr;
    float *power_correction_mask = nullptr;
};

// plain ring buffers with the reference's window semantics (oldest block first after forward())
class Streams {
public:
    std::vector<std::vector<float>> rings;
    unsigned position = 0;  // in floats here
    void create(int n) { rings.assign(n, std::vector<float>(2 * N_ITEMS_BUFFER, 0.f)); }
    void write_stream(unsigned i, const float *data) {
        std::memcpy(&rings[i][position], data, N_SAMPLES * sizeof(float));
        std::memcpy(&rings[i][(position + N_ITEMS_BUFFER) % (2 * N_ITEMS_BUFFER)], data, N_SAMPLES * sizeof(float));
    }
    void forward() { position = (position + N_SAMPLES) % N_ITEMS_BUFFER; }
    float *get_signal(unsigned i, int offset) { return &rings[i][position + offset]; }
    void read_stream(unsigned i, float *data, unsigned offset = 0) {
        std::memcpy(data, &rings[i][position + offset], N_ITEMS_BUFFER * sizeof(float));
    }
};

class Pipeline {
public:
    Streams streams;
    std::atomic<int> modified{0};
    std::atomic<bool> running{true};
    std::mutex m;
    std::condition_variable cv_;
    Streams *getStreams() { return &streams; }
    bool isRunning() { return running; }
    int mostRecent() { return modified; }
    void barrier() {
        std::unique_lock<std::mutex> lk(m);
        int seen = modified;
        cv_.wait(lk, [&] { return modified != seen || !running; });
    }
    void release_barrier() {
        { std::lock_guard<std::mutex> lk(m); modified++; }
        cv_.notify_all();
    }
    void stop() { running = false; cv_.notify_all(); }
};

// src/dsp/worker.h:32-61
struct Target {
    Spherical direction;
    float power;
    float probability;
    std::chrono::time_point<std::chrono::high_resolution_clock> start;
    bool operator==(const Target &other) const {
        return fabs(direction.phi - other.direction.phi) < 1e-2 && fabs(direction.theta - other.direction.theta) < 1e-2;
    }
    Target(Spherical direction, float power, float probability, std::chrono::time_point<std::chrono::high_resolution_clock> start)
        : direction(direction), power(power), probability(probability), start(start) {}
};

enum worker_t { GENERIC, PSO, MIMO, MISO, SOUND, GRADIENT };

class Worker {
public:
    bool *running;
    bool looping;
    std::thread thread_loop;
    Pipeline *pipeline;
    Worker(Pipeline *pipeline, Antenna &antenna, bool *running) : running(running), looping(true), pipeline(pipeline), antenna(antenna) {
        streams = pipeline->getStreams();
    }
    ~Worker() {   // NOT virtual, like the reference (src/dsp/worker.h:111-114): AWProcessingUnit deletes through Worker*
        looping = false;
        thread_loop.join();
    }
    std::vector<Target> getTargets() const {   // worker.h:136-142 (called without the lock by TargetHandler)
        std::vector<Target> r_targets;
        for (size_t i = 0; i < tracking.size(); ++i) r_targets.insert(r_targets.end(), tracking[i]);
        return r_targets;
    }
    virtual worker_t get_type() { return worker_t::GENERIC; }
    void draw(cv::Mat *heatmap) {
        lock.lock();
        populateHeatmap(heatmap);
        lock.unlock();
    }
    virtual void steer(Spherical) {}

protected:
    Spherical direction;
    Streams *streams;
    Antenna &antenna;
    std::vector<Target> tracking;
    virtual void update() {}
    virtual void reset() {}
    virtual void populateHeatmap(cv::Mat *) {}
    virtual void setup() {}
    void loop() {
        setup();
        while (looping && pipeline->isRunning()) {
            pipeline->barrier();
            if (!looping || !pipeline->isRunning()) break;
            lock.lock();
            reset();
            update();
            lock.unlock();
        }
    }

private:
    std::mutex lock;
};
