/*
 * repro_math.h -- TEST INFRASTRUCTURE (oracle side).
 *
 * Host restatement of the bit-reproducible exp used by the device residual functors
 * (mir_optim_b200/csrc/repro_math.cuh documents why it exists).  Written independently with
 * <cmath> fma/rint/ldexp, which round exactly like their CUDA counterparts; compiled with
 * -ffp-contract=off so nothing else is fused.  k = rint(x*log2 e); r = x - k*ln2 (Cody-Waite,
 * two constants); Taylor polynomial (degree 13 double / 7 float) by Horner with fma; p * 2^k.
 */
#pragma once
#include <cmath>
#include <limits>

namespace oracle_math {

inline double exp_repro(double x)
{
    if (!(x > -745.2)) return (x == x) ? 0.0 : x;
    if (x > 709.782712893384) return std::numeric_limits<double>::infinity();
    const double k = std::rint(x * 0x1.71547652b82fep+0);
    double r = std::fma(-k, 0x1.62e42feep-1, x);
    r = std::fma(-k, 0x1.a39ef35793c76p-33, r);
    static const double c[14] = {1.0, 1.0, 0.5, 0x1.5555555555555p-3, 0x1.5555555555555p-5, 0x1.1111111111111p-7,
                                 0x1.6c16c16c16c17p-10, 0x1.a01a01a01a01ap-13, 0x1.a01a01a01a01ap-16, 0x1.71de3a556c734p-19,
                                 0x1.27e4fb7789f5cp-22, 0x1.ae64567f544e4p-26, 0x1.1eed8eff8d898p-29, 0x1.6124613a86d09p-33};
    double p = c[13];
    for (int i = 12; i >= 0; --i) p = std::fma(p, r, c[i]);
    return std::ldexp(p, (int)k);            /* exact scaling (correctly rounded into the subnormals) */
}

inline float exp_repro(float x)
{
    if (!(x > -104.0f)) return (x == x) ? 0.0f : x;
    if (x > 88.72284f) return std::numeric_limits<float>::infinity();
    const float k = std::rint(x * 0x1.715476p+0f);
    float r = std::fma(-k, 0x1.62e4p-1f, x);
    r = std::fma(-k, 0x1.7f7d1cp-20f, r);
    static const float c[8] = {1.0f, 1.0f, 0.5f, 0x1.555556p-3f, 0x1.555556p-5f, 0x1.111112p-7f, 0x1.6c16c2p-10f, 0x1.a01a02p-13f};
    float p = c[7];
    for (int i = 6; i >= 0; --i) p = std::fma(p, r, c[i]);
    return std::ldexp(p, (int)k);
}

}  // namespace oracle_math
