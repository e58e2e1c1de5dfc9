// runtime.cuh -- error reporting, launch accounting and small host helpers shared by the C-ABI
// translation units.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "common.cuh"

namespace mirb200 {

void set_error(const std::string& msg);
void clear_error();
// Returns MIR_B200_OK or records the CUDA error string and returns MIR_B200_ECUDA / ENODEVICE.
int check_cuda(cudaError_t e, const char* what);
// Makes sure a CUDA device is usable; selects `device` if >= 0.  No CPU fallback: callers bail out.
int require_device(int device);
int sm_count();
// thread-local, grow-only pinned host buffer (valid until the same thread asks for a larger one)
void* pinned_scratch(size_t bytes);

#define MIRB200_CUDA(expr)                                              \
    do { int _rc = ::mirb200::check_cuda((expr), #expr); if (_rc) return _rc; } while (0)

}  // namespace mirb200
