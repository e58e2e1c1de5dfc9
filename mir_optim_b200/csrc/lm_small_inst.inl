// lm_small_inst.inl -- model dispatch; included by lm_small_inst_d.cu / lm_small_inst_s.cu with REAL defined.
#include <initializer_list>
#ifndef MIRB200_GAUSS_LANES
#define MIRB200_GAUSS_LANES 8
#endif
#include "lm_small_launch.cuh"

namespace mirb200 {

template <>
int launch_small_model<REAL>(unsigned model, size_t n, const Num<REAL>::Settings& st, const SmallBatchArgs& args, cudaStream_t stream)
{
    using T = REAL;
    auto bad_n = [&](int want) {
        set_error("mir_optim_b200: model expects n = " + std::to_string(want) + ", got " + std::to_string(n));
        return (int)MIR_B200_EINVAL;
    };
    // m beyond the largest rows-per-lane instantiation of the lane-group kernel (4 x 32 rows): the thread-per-problem
    // kernel takes any m, whatever the batch size (a single problem is left to the caller: the legacy entry point
    // sends it to the row-parallel large-problem engine instead of one thread)
    // (warm-started batches, MIR_MODEL_WARM_START, go to the kernels below: lane-group for m <= 128, general otherwise)
    const bool warm = (args.flags & MIR_MODEL_WARM_START) != 0;
    if ((use_thread_per_problem(args.batch) && !(warm && args.m <= 128)) || (args.m > 128 && args.batch > 1 && !warm)) {
        switch (model) {
        case MIR_MODEL_EXPDECAY2:  if (n != 2) return bad_n(2); return launch_tpp<ModelExpDecay2<T, true>, T>(st, args, stream);
        case MIR_MODEL_EXPTAU3:    if (n != 3) return bad_n(3); return launch_tpp<ModelExpTau3<T, true>, T>(st, args, stream);
        case MIR_MODEL_EXPDECAY3:  if (n != 3) return bad_n(3); return launch_tpp<ModelExpDecay3<T, true>, T>(st, args, stream);
        case MIR_MODEL_GAUSS4:     if (n != 4) return bad_n(4); return launch_tpp<ModelGauss4<T, true>, T>(st, args, stream);
        case MIR_MODEL_SUMEXP:
            if (n == 4) return launch_tpp<ModelSumExp<T, 4, true>, T>(st, args, stream);
            // n = 8: the per-thread state (36-entry packed J^T J + factor, 8 KB of Jacobian) spills and the slab
            // traffic dominates -- measured 4x slower than the lane-group kernel on B200, which therefore keeps it
            break;
        default: break;      // data-free two-row models: the group kernel below
        }
    }
    switch (model) {
    case MIR_MODEL_LINEAR2:    if (n != 2) return bad_n(2); return launch_small<ModelLinear2<T>, T, 32, 1>(st, args, stream);
    case MIR_MODEL_ROSENBROCK: if (n != 2) return bad_n(2); return launch_small<ModelRosenbrock<T>, T, 32, 1>(st, args, stream);
    case MIR_MODEL_SQRTCIRCLE: if (n != 2) return bad_n(2); return launch_small<ModelSqrtCircle<T>, T, 32, 1>(st, args, stream);
    case MIR_MODEL_EXPDECAY2:  if (n != 2) return bad_n(2); return launch_small<ModelExpDecay2<T>, T, 32, 1, 4>(st, args, stream);
    case MIR_MODEL_EXPTAU3:    if (n != 3) return bad_n(3); return launch_small<ModelExpTau3<T>, T, 32, 1, 4>(st, args, stream);
    case MIR_MODEL_EXPDECAY3:  if (n != 3) return bad_n(3); return launch_small<ModelExpDecay3<T>, T, 32, 1, 4>(st, args, stream);
    case MIR_MODEL_GAUSS4:
        if (n != 4) return bad_n(4);
        // m <= 64: an 8-lane group per problem (4 problems per warp) measured fastest on B200 (the n-sized
        // serial part is replicated 8x instead of 32x); larger m: one full warp per problem.
        if (args.m <= 64) return launch_small<ModelGauss4<T>, T, MIRB200_GAUSS_LANES, 64 / MIRB200_GAUSS_LANES>(st, args, stream);
        return launch_small<ModelGauss4<T>, T, 32, 4>(st, args, stream);
    case MIR_MODEL_SUMEXP:
        if (n == 4) return launch_small<ModelSumExp<T, 4>, T, 32, 2, 4>(st, args, stream);
        if (n == 8) {
            // four problems per warp (lm_mux.cuh); MIRB200_N8_KERNEL=warp keeps the one-warp-per-problem kernel (experiments)
            static const bool warpKernel = [] { const char* e = std::getenv("MIRB200_N8_KERNEL"); return e && !std::strcmp(e, "warp"); }();
            if (args.m <= (unsigned)MUX_MMAX && !warpKernel) return launch_mux<ModelSumExp<T, 8, true>, T>(st, args, stream);
            return launch_small<ModelSumExp<T, 8>, T, 32, 2, 4>(st, args, stream);
        }
        set_error("mir_optim_b200: batched SUMEXP is instantiated for n = 4 and n = 8");
        return MIR_B200_EUNSUPPORTED;
    default:
        set_error("mir_optim_b200: model id not available in the batched small-problem path");
        return MIR_B200_EUNSUPPORTED;
    }
}

}  // namespace mirb200
