// lm_small_launch.cuh -- host-side launcher for lm_small_kernel: picks the rows-per-lane
// instantiation from m, sizes a persistent grid (a multiple of the SM count) and enqueues.
#pragma once
#include "lm_small.cuh"
#include "runtime.cuh"

namespace mirb200 {

template <class Model, class T, int LANES, int R, bool FD>
int launch_small_fd(const typename Num<T>::Settings& st, const SmallBatchArgs& args, cudaStream_t stream)
{
    auto kern = lm_small_kernel<Model, T, LANES, R, FD>;
    int blocksPerSM = 0;
    MIRB200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, kern, 128, 0));
    if (blocksPerSM < 1) blocksPerSM = 1;
    const unsigned long long groupsPerBlock = 128 / LANES;
    unsigned long long blocksWanted = (args.batch + groupsPerBlock - 1) / groupsPerBlock;
    unsigned long long grid = (unsigned long long)sm_count() * blocksPerSM;     // persistent: one resident wave
    if (blocksWanted < grid) grid = blocksWanted ? blocksWanted : 1;
    kern<<<(unsigned)grid, 128, 0, stream>>>(st, args);
    count_launch();
    return check_cuda(cudaGetLastError(), "lm_small_kernel launch");
}

template <class Model, class T, int LANES, int R>
int launch_small_one(const typename Num<T>::Settings& st, const SmallBatchArgs& args, cudaStream_t stream)
{
    return (args.flags & MIR_MODEL_FD_JACOBIAN) ? launch_small_fd<Model, T, LANES, R, true>(st, args, stream)
                                                : launch_small_fd<Model, T, LANES, R, false>(st, args, stream);
}

// Rs...: the rows-per-lane instantiations compiled for this model, ascending.
template <class Model, class T, int LANES, int... Rs>
int launch_small(const typename Num<T>::Settings& st, const SmallBatchArgs& args, cudaStream_t stream)
{
    int rc = MIR_B200_EUNSUPPORTED;
    bool done = false;
    // pick the first R with R*LANES >= m
    (void)std::initializer_list<int>{
        ((!done && (unsigned)(Rs * LANES) >= args.m) ? (done = true, rc = launch_small_one<Model, T, LANES, Rs>(st, args, stream), 0) : 0)...};
    if (!done) set_error("mir_optim_b200: m is larger than the batched small-problem kernels support for this model");
    return rc;
}

// Entry used by the C ABI (lm_batched.cu); one explicit instantiation per precision lives in
// lm_small_inst_{d,s}.cu so the two precisions compile in parallel.
template <class T>
int launch_small_model(unsigned model, size_t n, const typename Num<T>::Settings& st, const SmallBatchArgs& args, cudaStream_t stream);

}  // namespace mirb200
