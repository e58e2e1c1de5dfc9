// repro_math.cuh -- bit-reproducible exp for the device residual functors, plus out-of-line
// wrappers for the IEEE division / square root sequences.
//
// Why a private exp: with the finite-difference Jacobian (least_squares.d:1016-1050) residual
// rounding differences are amplified by 1/(2*jacobianEpsilon) ~ 3e7, so a 1-ulp difference between
// CUDA's exp and the host libm's exp moves the first step by ~1e-8 relative -- far outside the 1e-10
// parity bar.  exp_repro uses only operations that round identically on the host and on the GPU
// (fma, mul, rint, exact power-of-two scaling), so a host callback written with the same sequence
// (oracle/repro_math.h, an independent restatement) returns the same bits for every input.
// Algorithm: k = rint(x*log2(e)); r = x - k*ln2 (two-constant Cody-Waite with fma); degree-13
// (double) / degree-7 (float) Taylor polynomial in Horner form with fma; result = p * 2^k.
// Max error about 1 ulp.
//
// Why out-of-line (__noinline__) helpers: the warp-per-problem LM kernel is instruction-fetch
// bound when everything is inlined (ncu: stall_no_instruction dominates); one copy of each
// 15-40 instruction sequence keeps the hot loop inside the instruction cache.
#pragma once
#include "common.cuh"

namespace mirb200 {

// (static: one private copy per translation unit, so the host-side stubs never collide at link time)

// Polynomial / reduction constants live in __constant__ memory: a DFMA takes a constant-bank operand directly,
// while a 64-bit immediate costs two extra moves per use (ncu, round 1: 38 % of the thread-per-problem row loop
// was constant materialisation).  Same bits as the literals they replace.
static __constant__ double kExpD[17] = {
    0x1.71547652b82fep+0, 0x1.62e42feep-1, 0x1.a39ef35793c76p-33,                     // log2(e), ln2 hi, ln2 lo
    0x1.6124613a86d09p-33, 0x1.1eed8eff8d898p-29, 0x1.ae64567f544e4p-26, 0x1.27e4fb7789f5cp-22, 0x1.71de3a556c734p-19,
    0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-13, 0x1.6c16c16c16c17p-10, 0x1.1111111111111p-7, 0x1.5555555555555p-5,
    0x1.5555555555555p-3, 0.5, 1.0, 1.0};
static __constant__ float kExpS[11] = {
    0x1.715476p+0f, 0x1.62e4p-1f, 0x1.7f7d1cp-20f,
    0x1.a01a02p-13f, 0x1.6c16c2p-10f, 0x1.111112p-7f, 0x1.555556p-5f, 0x1.555556p-3f, 0.5f, 1.0f, 1.0f};

// Branch-free on purpose: the range checks are selects at the end and the scaling is two exact power-of-two
// multiplications (p * 2^(k/2) is exact, the second product rounds once -- the value ldexp returns, also into the
// subnormals and into overflow).  exp_repro_many evaluates K independent arguments with the Horner steps
// INTERLEAVED in source order, so the K dependent-FMA chains overlap in the pipe (ncu, round 1: with early returns
// every exp was its own branch region, and even in one basic block ptxas emitted the chains one after the other;
// each of the 16 dependent DFMAs then waited out the full pipe latency).
template <int K>
static __device__ __forceinline__ void exp_repro_many(const double* __restrict__ x, double* __restrict__ e)
{
    // No clamping of the argument: for x outside [-745.2, 709.78] (or NaN) the polynomial path produces garbage
    // (the float -> int conversion saturates, nothing traps) that the final selects replace.
    double k[K], r[K], p[K];
#pragma unroll
    for (int j = 0; j < K; ++j) k[j] = rint(__dmul_rn(x[j], kExpD[0]));
#pragma unroll
    for (int j = 0; j < K; ++j) r[j] = fma(-k[j], kExpD[1], x[j]);
#pragma unroll
    for (int j = 0; j < K; ++j) r[j] = fma(-k[j], kExpD[2], r[j]);
#pragma unroll
    for (int j = 0; j < K; ++j) p[j] = kExpD[3];
#pragma unroll
    for (int i = 4; i < 17; ++i) {
        const double c = kExpD[i];
#pragma unroll
        for (int j = 0; j < K; ++j) p[j] = fma(p[j], r[j], c);
    }
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const int ki = __double2int_rn(k[j]), h = ki >> 1;                  // saturating conversion: defined for any k
        const double s1 = __hiloint2double((int)((unsigned)(h + 1023) << 20), 0);
        const double s2 = __hiloint2double((int)((unsigned)(ki - h + 1023) << 20), 0);
        double v = __dmul_rn(__dmul_rn(p[j], s1), s2);
        v = (x[j] > 709.782712893384) ? Num<double>::inf() : v;
        e[j] = (x[j] > -745.2) ? v : ((x[j] == x[j]) ? 0.0 : x[j]);
    }
}

// Warp-convergent variant (all 32 lanes must call it together): same value for every input, but the scaling p * 2^k is
// an integer add on the exponent field when every argument of the warp is in [-700, 700] -- the result is then a normal
// number and both exact power-of-two products of the general path reduce to that add.  Anything else (huge arguments,
// NaN) sends the whole warp through the general path.  20 instead of 33 instructions per exponential.
template <int K>
static __device__ __forceinline__ void exp_repro_many_conv(const double* __restrict__ x, double* __restrict__ e, unsigned mask = 0xffffffffu)
{
    bool ok = true;
#pragma unroll
    for (int j = 0; j < K; ++j) ok = ok && (fabs(x[j]) < 700.0);
    if (__all_sync(mask, ok)) {
        double k[K], r[K], p[K];
#pragma unroll
        for (int j = 0; j < K; ++j) k[j] = rint(__dmul_rn(x[j], kExpD[0]));
#pragma unroll
        for (int j = 0; j < K; ++j) r[j] = fma(-k[j], kExpD[1], x[j]);
#pragma unroll
        for (int j = 0; j < K; ++j) r[j] = fma(-k[j], kExpD[2], r[j]);
#pragma unroll
        for (int j = 0; j < K; ++j) p[j] = kExpD[3];
#pragma unroll
        for (int i = 4; i < 17; ++i) {
            const double c = kExpD[i];
#pragma unroll
            for (int j = 0; j < K; ++j) p[j] = fma(p[j], r[j], c);
        }
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int ki = __double2int_rn(k[j]);
            e[j] = __hiloint2double(__double2hiint(p[j]) + (ki << 20), __double2loint(p[j]));
        }
    } else {
        exp_repro_many<K>(x, e);
    }
}
template <int K>
static __device__ __forceinline__ void exp_repro_many_conv(const float* __restrict__ x, float* __restrict__ e, unsigned mask = 0xffffffffu);

template <int K>
static __device__ __forceinline__ void exp_repro_many(const float* __restrict__ x, float* __restrict__ e)
{
    float k[K], r[K], p[K];
#pragma unroll
    for (int j = 0; j < K; ++j) k[j] = rintf(__fmul_rn(x[j], kExpS[0]));
#pragma unroll
    for (int j = 0; j < K; ++j) r[j] = fmaf(-k[j], kExpS[1], x[j]);
#pragma unroll
    for (int j = 0; j < K; ++j) r[j] = fmaf(-k[j], kExpS[2], r[j]);
#pragma unroll
    for (int j = 0; j < K; ++j) p[j] = kExpS[3];
#pragma unroll
    for (int i = 4; i < 11; ++i) {
        const float c = kExpS[i];
#pragma unroll
        for (int j = 0; j < K; ++j) p[j] = fmaf(p[j], r[j], c);
    }
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const int ki = __float2int_rn(k[j]), h = ki >> 1;
        const float s1 = __int_as_float((int)((unsigned)(h + 127) << 23));
        const float s2 = __int_as_float((int)((unsigned)(ki - h + 127) << 23));
        float v = __fmul_rn(__fmul_rn(p[j], s1), s2);
        v = (x[j] > 88.72284f) ? Num<float>::inf() : v;
        e[j] = (x[j] > -104.0f) ? v : ((x[j] == x[j]) ? 0.0f : x[j]);
    }
}

template <int K>
static __device__ __forceinline__ void exp_repro_many_conv(const float* __restrict__ x, float* __restrict__ e, unsigned mask)
{
    bool ok = true;
#pragma unroll
    for (int j = 0; j < K; ++j) ok = ok && (fabsf(x[j]) < 80.0f);
    if (__all_sync(mask, ok)) {
        float k[K], r[K], p[K];
#pragma unroll
        for (int j = 0; j < K; ++j) k[j] = rintf(__fmul_rn(x[j], kExpS[0]));
#pragma unroll
        for (int j = 0; j < K; ++j) r[j] = fmaf(-k[j], kExpS[1], x[j]);
#pragma unroll
        for (int j = 0; j < K; ++j) r[j] = fmaf(-k[j], kExpS[2], r[j]);
#pragma unroll
        for (int j = 0; j < K; ++j) p[j] = kExpS[3];
#pragma unroll
        for (int i = 4; i < 11; ++i) {
            const float c = kExpS[i];
#pragma unroll
            for (int j = 0; j < K; ++j) p[j] = fmaf(p[j], r[j], c);
        }
#pragma unroll
        for (int j = 0; j < K; ++j) e[j] = __int_as_float(__float_as_int(p[j]) + (__float2int_rn(k[j]) << 23));
    } else {
        exp_repro_many<K>(x, e);
    }
}

static __device__ __forceinline__ double exp_repro_inl(double x) { double e; exp_repro_many<1>(&x, &e); return e; }
static __device__ __forceinline__ float  exp_repro_inl(float x)  { float e;  exp_repro_many<1>(&x, &e); return e; }

// Throughput variant for kernels with plenty of warps per SM (the single-large-problem evaluators): early returns and
// immediate constants cost fewer FP64-pipe slots than the branch-free form, and other warps hide the dependent chain
// (measured on B200: large_eval_kernel 0.40 ms with this form, 0.54 ms with the branch-free one).  Same value for
// every input: both scalings are exact / correctly rounded.
static __device__ __noinline__ double exp_repro_tp(double x)
{
    if (!(x > -745.2)) return (x == x) ? 0.0 : x;
    if (x > 709.782712893384) return Num<double>::inf();
    const double k = rint(__dmul_rn(x, 0x1.71547652b82fep+0));
    double r = fma(-k, 0x1.62e42feep-1, x);
    r = fma(-k, 0x1.a39ef35793c76p-33, r);
    double p = 0x1.6124613a86d09p-33;
    p = fma(p, r, 0x1.1eed8eff8d898p-29);
    p = fma(p, r, 0x1.ae64567f544e4p-26);
    p = fma(p, r, 0x1.27e4fb7789f5cp-22);
    p = fma(p, r, 0x1.71de3a556c734p-19);
    p = fma(p, r, 0x1.a01a01a01a01ap-16);
    p = fma(p, r, 0x1.a01a01a01a01ap-13);
    p = fma(p, r, 0x1.6c16c16c16c17p-10);
    p = fma(p, r, 0x1.1111111111111p-7);
    p = fma(p, r, 0x1.5555555555555p-5);
    p = fma(p, r, 0x1.5555555555555p-3);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int ki = (int)k;
    if (ki >= -1000 && ki <= 1000) return __dmul_rn(p, __longlong_as_double((long long)(ki + 1023) << 52));
    return ldexp(p, ki);
}
static __device__ __noinline__ float exp_repro_tp(float x)
{
    if (!(x > -104.0f)) return (x == x) ? 0.0f : x;
    if (x > 88.72284f) return Num<float>::inf();
    const float k = rintf(__fmul_rn(x, 0x1.715476p+0f));
    float r = fmaf(-k, 0x1.62e4p-1f, x);
    r = fmaf(-k, 0x1.7f7d1cp-20f, r);
    float p = 0x1.a01a02p-13f;
    p = fmaf(p, r, 0x1.6c16c2p-10f);
    p = fmaf(p, r, 0x1.111112p-7f);
    p = fmaf(p, r, 0x1.555556p-5f);
    p = fmaf(p, r, 0x1.555556p-3f);
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    const int ki = (int)k;
    if (ki >= -120 && ki <= 120) return __fmul_rn(p, __int_as_float((ki + 127) << 23));
    return ldexpf(p, ki);
}

static __device__ __noinline__ double exp_repro(double x) { return exp_repro_inl(x); }
static __device__ __noinline__ float  exp_repro(float x)  { return exp_repro_inl(x); }
// INL = true: inlined, so independent rows interleave their polynomial chains (thread-per-problem kernel);
// INL = false: one out-of-line copy (lane-group kernel, instruction-cache bound when inlined).
template <bool INL, class T> __device__ __forceinline__ T exp_sel(T x) { if constexpr (INL) return exp_repro_inl(x); else return exp_repro(x); }

static __device__ __noinline__ double rcp_ni(double a) { return 1.0 / a; }
static __device__ __noinline__ float  rcp_ni(float a)  { return 1.0f / a; }
static __device__ __noinline__ double div_ni(double a, double b) { return a / b; }
static __device__ __noinline__ float  div_ni(float a, float b)   { return a / b; }
static __device__ __noinline__ double sqrt_ni(double a) { return sqrt(a); }
static __device__ __noinline__ float  sqrt_ni(float a)  { return sqrtf(a); }

// Pivot step of the Cholesky factorisation: l = sqrt(a) and r = 1 / l (potf2 scales the column by the reciprocal).  The
// IEEE sqrt followed by the IEEE reciprocal is a chain of ~160 cycles and ~65 instructions; this one starts from the
// hardware's reciprocal-square-root seed (MUFU.RSQ64H, ~2^-22), runs two Newton steps and corrects l and r once more:
// both are within an ulp of the rounded values (the restatement is compared with LAPACK by tolerance, not bit for bit).
static __device__ __forceinline__ void mux_sqrt_rcp(double a, double& l, double& r)
{
    // (the seed instruction flushes subnormals: tiny pivots are moved up by 2^200 first, an exact scaling; NaN, +-inf
    //  and pivots <= 0 produce values nobody uses -- the caller has recorded the breakdown, or LAPACK would return NaN too)
    const bool tiny = a < 0x1p-900;
    const double as = tiny ? a * 0x1p200 : a;
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(as));
    double t = as * y, e = fma(-t, y, 1.0);
    y = fma(0.5 * y, e, y);
    t = as * y; e = fma(-t, y, 1.0);
    y = fma(0.5 * y, e, y);
    l = as * y;
    l = fma(fma(-l, l, as), 0.5 * y, l);
    r = fma(fma(-l, y, 1.0), y, y);
    l = tiny ? l * 0x1p-100 : l;
    r = tiny ? r * 0x1p100 : r;
}
static __device__ __forceinline__ void mux_sqrt_rcp(float a, float& l, float& r)
{
    const bool tiny = a < 0x1p-100f;
    const float as = tiny ? a * 0x1p40f : a;
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(as));
    const float t = as * y, e = fmaf(-t, y, 1.0f);
    y = fmaf(0.5f * y, e, y);
    l = as * y;
    l = fmaf(fmaf(-l, l, as), 0.5f * y, l);
    r = fmaf(fmaf(-l, y, 1.0f), y, y);
    l = tiny ? l * 0x1p-20f : l;
    r = tiny ? r * 0x1p20f : r;
}


}  // namespace mirb200
