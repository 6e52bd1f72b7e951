// steer.cuh -- the device side of steering_vector_spherical, shared by the table kernel (tables.cu) and the fused MISO
// kernel (das_miso.cu) so both produce the same bits.  Every rounding is pinned (__f*_rn): see tables.cu.
#pragma once
#include "bflk_internal.h"

namespace bflk {

// z'' of steer() (src/geometry/antenna.cpp:99-107) for one element
__device__ __forceinline__ float steer_z(const DirTrig t, float px, float py, float pz) {
    // rotateZ(float(phi)) * p, k = 0,1,2 from a zero accumulator (documented evaluation order)
    float xr = __fmaf_rn(0.0f, pz, __fmaf_rn(-t.sz, py, __fmaf_rn(t.cz, px, 0.0f)));
    float yr = __fmaf_rn(0.0f, pz, __fmaf_rn(t.cz, py, __fmaf_rn(t.sz, px, 0.0f)));
    float zr = __fmaf_rn(1.0f, pz, __fmaf_rn(0.0f, py, __fmaf_rn(0.0f, px, 0.0f)));
    // row Z of rotateY(-float(theta)): (-sin, 0, cos)
    return __fmaf_rn(t.cy, zr, __fmaf_rn(0.0f, yr, __fmaf_rn(-t.sy, xr, 0.0f)));
}


// compute_delays() scale (antenna.cpp:90-92): row(Z) * float(SAMPLE_RATE / PROPAGATION_SPEED)
__device__ __forceinline__ float steer_delay(const DirTrig t, const float *__restrict__ xyz, int c, float k_scale) {
    return __fmul_rn(steer_z(t, xyz[3 * c + 0], xyz[3 * c + 1], xyz[3 * c + 2]), k_scale);
}

}  // namespace bflk
