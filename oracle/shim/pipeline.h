/* Shim for building the reference's src/dsp/delay.cpp on its own: delay.h includes "pipeline.h"
 * only to obtain N_SAMPLES, which src/fpga/streams.hpp defines.  The real pipeline.h drags in Eigen
 * (not installed here).  TEST INFRASTRUCTURE ONLY. */
#pragma once
#include "streams.hpp"
