// common.cuh -- shared device/host plumbing for the B200 LM engine (sm_100a only).
#pragma once

#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#endif

#include "../../include/mir_optim_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "mir_optim_b200 targets sm_100a (B200) only"
#endif

namespace mirb200 {

// ----------------------------------------------------------------------------------------
// Per-precision constants.  LAPACK's dlamch/slamch values are spelled out because posvx's
// control flow (equilibration threshold, refinement stop test) depends on them.
// ----------------------------------------------------------------------------------------
template <class T> struct Num;
template <> struct Num<double> {
    using Settings = mir_least_squares_settings_d;
    using Result = mir_least_squares_result_d;
    using QPSettings = mir_box_qp_settings_d;
#ifdef __CUDACC_RTC__
    __device__ static double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
#else
    __host__ __device__ static constexpr double inf() { return __builtin_huge_val(); }
#endif
    __host__ __device__ static constexpr double lapack_eps() { return 0x1p-53; }      // dlamch('Epsilon')
    __host__ __device__ static constexpr double safmin() { return 0x1p-1022; }         // dlamch('Safe minimum')
    __host__ __device__ static constexpr double small_() { return 0x1p-970; }          // safmin / dlamch('Precision')
    __host__ __device__ static constexpr double large_() { return 0x1p970; }
    __host__ __device__ static constexpr double sqrt_max() { return 1.3407807929942596e154; }   // sqrt(double.max)
    __host__ __device__ static constexpr double sqrt_min_normal() { return 0x1p-511; }
};
template <> struct Num<float> {
    using Settings = mir_least_squares_settings_s;
    using Result = mir_least_squares_result_s;
    using QPSettings = mir_box_qp_settings_s;
#ifdef __CUDACC_RTC__
    __device__ static float inf() { return __int_as_float(0x7f800000); }
#else
    __host__ __device__ static constexpr float inf() { return __builtin_huge_valf(); }
#endif
    __host__ __device__ static constexpr float lapack_eps() { return 0x1p-24f; }
    __host__ __device__ static constexpr float safmin() { return 0x1p-126f; }
    __host__ __device__ static constexpr float small_() { return 0x1p-103f; }
    __host__ __device__ static constexpr float large_() { return 0x1p103f; }
    __host__ __device__ static constexpr float sqrt_max() { return 1.8446743e19f; }             // sqrt(float.max)
    __host__ __device__ static constexpr float sqrt_min_normal() { return 0x1p-63f; }
};

// Explicitly rounded add (never contracted into an FMA, never re-associated): used where the
// reference relies on two separately rounded operations, e.g. (dx + x) - x at LS:1096-1097.
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float  add_rn(float a, float b)   { return __fadd_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  mul_rn(float a, float b)   { return __fmul_rn(a, b); }

__device__ __forceinline__ double sub_rn(double a, double b) { return __dadd_rn(a, -b); }
__device__ __forceinline__ float  sub_rn(float a, float b)   { return __fadd_rn(a, -b); }
__device__ __forceinline__ double t_abs(double a)  { return fabs(a); }
__device__ __forceinline__ float  t_abs(float a)   { return fabsf(a); }
__device__ __forceinline__ double t_min(double a, double b) { return fmin(a, b); }   // IEEE minNum, as D's fmin
__device__ __forceinline__ float  t_min(float a, float b)   { return fminf(a, b); }
// a - b * c in ONE rounding, spelled out: the compiler's own contraction of `a -= b * c` depends on the shape of the
// surrounding code (measured: the same source line fused in one unrolled copy and split into DMUL + DADD in another),
// which breaks "same operations, same order => same bits" between two restatements of one algorithm.
__device__ __forceinline__ double fnma(double b, double c, double a) { return fma(-b, c, a); }
__device__ __forceinline__ float  fnma(float b, float c, float a)   { return fmaf(-b, c, a); }
__device__ __forceinline__ double t_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float  t_max(float a, float b)   { return fmaxf(a, b); }

// ----------------------------------------------------------------------------------------
// Group (sub-warp) collectives.  A "group" is LANES consecutive lanes of one warp working on
// one problem; every lane keeps a bit-identical replica of the problem's small state.  The
// xor-butterfly delivers the same bits to every lane because each stage adds the same two
// operands (in either order) on both partners.
// ----------------------------------------------------------------------------------------
template <int LANES> __device__ __forceinline__ unsigned group_mask()
{
    if constexpr (LANES == 32) {
        return 0xffffffffu;
    } else {
        const unsigned lane = threadIdx.x & 31u;
        return ((1u << LANES) - 1u) << (lane & ~(unsigned)(LANES - 1));
    }
}

__device__ __forceinline__ double shfl_xor(unsigned mask, double v, int off, int width) { return __shfl_xor_sync(mask, v, off, width); }
__device__ __forceinline__ float  shfl_xor(unsigned mask, float v, int off, int width)  { return __shfl_xor_sync(mask, v, off, width); }

template <int LANES, class T> __device__ __forceinline__ T group_sum(unsigned mask, T v)
{
#pragma unroll
    for (int off = LANES / 2; off > 0; off >>= 1) v += shfl_xor(mask, v, off, LANES);
    return v;
}

// all-reduce K values at once (stage-major so the K shuffles of one stage are independent)
template <int LANES, int K, class T> __device__ __forceinline__ void group_sum_array(unsigned mask, T (&v)[K])
{
#pragma unroll
    for (int off = LANES / 2; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] += shfl_xor(mask, v[k], off, LANES);
    }
}

// packed lower-triangular index, i >= j
__host__ __device__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }
__host__ __device__ constexpr int trisym(int i, int j) { return i >= j ? tri(i, j) : tri(j, i); }

// Launch accounting (mir_b200_kernel_launches)
void count_launch(long long n = 1);

}  // namespace mirb200
