// adapter_main.cpp -- TEST INFRASTRUCTURE ONLY: drives CudaMIMOWorker / CudaMISOWorker the way
// AWProcessingUnit does (frames pushed through the rings, barrier released, draw() from another thread)
// and prints what the workers produced so tests/test_adapter.py can compare with the oracle.
// usage: adapter_main <stream.f32 (64 x T floats, channel-major)> <T> <rows> <cols> <fov> <theta> <phi>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cuda_workers.h"

int main(int argc, char **argv) {
    if (argc < 8) return 2;
    const int T = atoi(argv[2]), rows = atoi(argv[3]), cols = atoi(argv[4]);
    const float fov = atof(argv[5]);
    const double theta = atof(argv[6]), phi = atof(argv[7]);
    std::vector<float> stream((size_t)ELEMENTS * T);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(stream.data(), sizeof(float), stream.size(), f) != stream.size()) return 3;
    fclose(f);

    Pipeline pipeline;
    pipeline.streams.create(ELEMENTS);
    Antenna antenna;
    antenna.points.n = ELEMENTS;
    antenna.points.v.resize(3 * ELEMENTS);
    {   // create_antenna(Position(0,0,0), COLUMNS, ROWS, DISTANCE), src/geometry/antenna.cpp:60-76
        const float distance = 0.02f, half = distance / 2;
        int i = 0;
        for (int r = 0; r < 8; r++)
            for (int c = 0; c < 8; c++, i++) {
                antenna.points.v[3 * i + 0] = static_cast<float>(c) * distance - 8 * half + half;
                antenna.points.v[3 * i + 1] = static_cast<float>(r) * distance - 8 * half + half;
                antenna.points.v[3 * i + 2] = 0.f;
            }
    }
    bool running = true;
    // four frames fill the rings (what AWProcessingUnit::calibrate waits for, aw_processing_unit.cpp:105-107)
    for (int b = 0; b < T / N_SAMPLES; b++) {
        for (int c = 0; c < ELEMENTS; c++) pipeline.streams.write_stream(c, &stream[(size_t)c * T + b * N_SAMPLES]);
        pipeline.streams.forward();
    }
    {
        // created and destroyed the way AWProcessingUnit does: new in start() (aw_processing_unit.cpp:67-95), delete through
        // Worker* in the destructor (aw_processing_unit.cpp:51-53) -- ~Worker is not virtual
        CudaMIMOWorker *mimo = new CudaMIMOWorker(&pipeline, antenna, &running, rows, cols, fov);
        CudaMISOWorker *miso = new CudaMISOWorker(&pipeline, antenna, &running, fov);
        std::vector<Worker *> workers = {mimo, miso};
        mimo->update_once();
        cv::Mat heat(rows, cols);
        mimo->draw(&heat);
        printf("type %d %d\n", (int)mimo->get_type(), (int)miso->get_type());
        printf("power");
        for (float p : mimo->power()) printf(" %.9g", p);
        printf("\nheat");
        for (uchar h : heat.store) printf(" %d", (int)h);
        printf("\n");
        // AWProcessingUnit::targets() = workers[0]->getTargets() (aw_processing_unit.cpp:267-269), polled by TargetHandler
        const std::vector<Target> first = workers[0]->getTargets();
        printf("targets");
        for (const Target &t : first) printf(" %.17g %.17g %.9g %.9g", t.direction.theta, t.direction.phi, t.power, t.probability);
        printf("\n");
        mimo->update_once();
        const std::vector<Target> second = workers[0]->getTargets();
        int kept = second.size() == first.size();
        for (size_t i = 0; kept && i < second.size(); i++) kept = second[i] == first[i] && second[i].start == first[i].start;
        printf("start_kept %d\n", kept);
        miso->steer(Spherical(theta, phi));
        miso->update_once();
        printf("beam %.9g\naudio", miso->beam_power());
        for (int i = 0; i < N_SAMPLES; i++) printf(" %.9g", miso->audio()[i]);
        printf("\n");
        pipeline.release_barrier();  // one live frame through both Worker::loop() threads
        std::this_thread::sleep_for(std::chrono::milliseconds(50));
        pipeline.stop();             // AWProcessingUnit::~AWProcessingUnit disconnects before deleting workers
        for (Worker *job : workers) delete job;
        printf("deleted_through_base 1\n");
    }
    pipeline.stop();
    return 0;
}
