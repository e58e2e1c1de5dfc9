// nccl_dl.h -- run-time bound NCCL (see nccl_dl.cpp)
#pragma once
#include <stddef.h>
#include "../../include/mir_optim_b200.h"

namespace mirb200 {
int nccl_available();                                                         // MIR_B200_OK or MIR_B200_ENCCL (+ last_error)
int nccl_allreduce_sum(void* buf, size_t count, bool is_double, void* comm, void* stream);   // in place
}  // namespace mirb200
