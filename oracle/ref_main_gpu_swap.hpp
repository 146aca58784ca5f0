// TEST INFRASTRUCTURE. Force-included (-include) when the reference's UNMODIFIED main.cpp is compiled into
// oracle/_ref/coalref_main_gpu: it performs, without touching the source, the one-line change of INTEGRATION.md --
// the sampler type at main.cpp:281 becomes the reference-side binding include/GpuImportanceSampler.hpp.
#pragma once
#include "ImportanceSampler.hpp"        // the reference's class first, so that its include guard is set
#include "GpuImportanceSampler.hpp"
#define ImportanceSampler GpuImportanceSampler
