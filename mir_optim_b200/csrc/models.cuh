// models.cuh -- device residual functors ("device functors" of the north star).
//
// A model supplies, for one residual row i with abscissa t_i and observation y_i,
//     residual(pre, p, row, t, y)   = r_i(p)
//     jacobian(pre, p, row, t, Jr)  = d r_i / d p_k, k < N     (row-major row of J, LS:154)
// `prepare(p)` hoists per-parameter-vector work (reciprocals) out of the row loop.  These are
// the GPU counterparts of the reference's LeastSquaresFunction / LeastSquaresJacobian
// callbacks (least_squares.d:73-80); the first five are the reference's own unit-test
// problems (least_squares.d:217-434).
//
// Every operation is an explicitly rounded IEEE operation (mul_rn/add_rn/sub_rn, correctly
// rounded division and sqrt, exp_repro): no FMA contraction, no fast-math.  A host callback
// that performs the same operations in the same order (oracle/models_oracle.cpp, compiled with
// -ffp-contract=off) therefore returns bit-identical residuals and Jacobian rows, which is what
// lets finite-difference runs be compared with the CPU reference at 1e-10 (see repro_math.cuh).
#pragma once
#include "common.cuh"
#include "repro_math.cuh"

namespace mirb200 {

template <class T> struct NoPre {};

// r = (p0, 2 - p1)                                            least_squares.d:230-241
template <class T> struct ModelLinear2 {
    static constexpr int N = 2, NE = 0; static constexpr bool kHasData = false;
    using Pre = NoPre<T>;
    __device__ static Pre prepare(const T (&)[N]) { return {}; }
    __device__ static T residual(const Pre&, const T (&p)[N], int row, T, T) { return row == 0 ? p[0] : sub_rn((T)2, p[1]); }
    __device__ static void jacobian(const Pre&, const T (&)[N], int row, T, T (&J)[N]) {
        J[0] = row == 0 ? (T)1 : (T)0; J[1] = row == 0 ? (T)0 : (T)-1;
    }
    __device__ static void residual_jacobian(const Pre& q, const T (&p)[N], int row, T t, T y, T& r, T (&J)[N]) {
        r = residual(q, p, row, t, y); jacobian(q, p, row, t, J);
    }
};

// Rosenbrock: r = (10 (p1 - p0^2), 1 - p0)                     least_squares.d:261-265, 295-301
template <class T> struct ModelRosenbrock {
    static constexpr int N = 2, NE = 0; static constexpr bool kHasData = false;
    using Pre = NoPre<T>;
    __device__ static Pre prepare(const T (&)[N]) { return {}; }
    __device__ static T residual(const Pre&, const T (&p)[N], int row, T, T) {
        return row == 0 ? mul_rn((T)10, sub_rn(p[1], mul_rn(p[0], p[0]))) : sub_rn((T)1, p[0]);
    }
    __device__ static void jacobian(const Pre&, const T (&p)[N], int row, T, T (&J)[N]) {
        J[0] = row == 0 ? mul_rn((T)-20, p[0]) : (T)-1; J[1] = row == 0 ? (T)10 : (T)0;
    }
    __device__ static void residual_jacobian(const Pre& q, const T (&p)[N], int row, T t, T y, T& r, T (&J)[N]) {
        r = residual(q, p, row, t, y); jacobian(q, p, row, t, J);
    }
};

// r = sqrt(1 - (p0^2 + p1^2)), m = 1 < n = 2                   least_squares.d:427-430
template <class T> struct ModelSqrtCircle {
    static constexpr int N = 2, NE = 0; static constexpr bool kHasData = false;
    using Pre = NoPre<T>;
    __device__ static Pre prepare(const T (&)[N]) { return {}; }
    __device__ static T residual(const Pre&, const T (&p)[N], int, T, T) {
        return sqrt_ni(sub_rn((T)1, add_rn(mul_rn(p[0], p[0]), mul_rn(p[1], p[1]))));
    }
    __device__ static void jacobian(const Pre&, const T (&p)[N], int, T, T (&J)[N]) {
        const T s = sqrt_ni(sub_rn((T)1, add_rn(mul_rn(p[0], p[0]), mul_rn(p[1], p[1]))));
        J[0] = div_ni(-p[0], s); J[1] = div_ni(-p[1], s);
    }
    __device__ static void residual_jacobian(const Pre& q, const T (&p)[N], int row, T t, T y, T& r, T (&J)[N]) {
        r = residual(q, p, row, t, y); jacobian(q, p, row, t, J);
    }
};

// The exponential models are written in three pieces so that a caller can batch the exps of several rows / points
// into one interleaved exp_repro_many call:  exp_args (the NE exp arguments of a row), then finish_r / finish_j /
// finish_rj (residual and/or Jacobian row from the NE exp values).  residual(), jacobian() and residual_jacobian()
// are defined through the same pieces, so every path performs the same operations in the same order.
template <class M, class T, bool INL> struct ExpModelBase {
    // true: the model provides fd_shared(), a finite-difference Jacobian that reuses the sub-expressions its 2n
    // evaluations have in common (lm_mux.cuh); false: the kernels evaluate the residual 2n times
    static constexpr bool kFDShared = false;
    template <class P, int NN> __device__ static T residual(const P& q, const T (&p)[NN], int, T t, T y) {
        T a[M::NE], e[M::NE], r;
        M::exp_args(q, p, t, a);
#pragma unroll
        for (int k = 0; k < M::NE; ++k) e[k] = exp_sel<INL>(a[k]);
        M::finish_r(q, p, t, y, e, r);
        return r;
    }
    template <class P, int NN> __device__ static void jacobian(const P& q, const T (&p)[NN], int, T t, T (&J)[NN]) {
        T a[M::NE], e[M::NE];
        M::exp_args(q, p, t, a);
#pragma unroll
        for (int k = 0; k < M::NE; ++k) e[k] = exp_sel<INL>(a[k]);
        M::finish_j(q, p, t, e, J);
    }
    template <class P, int NN> __device__ static void residual_jacobian(const P& q, const T (&p)[NN], int, T t, T y, T& r, T (&J)[NN]) {
        T a[M::NE], e[M::NE];
        M::exp_args(q, p, t, a);
#pragma unroll
        for (int k = 0; k < M::NE; ++k) e[k] = exp_sel<INL>(a[k]);
        M::finish_r(q, p, t, y, e, r);
        M::finish_j(q, p, t, e, J);
    }
    template <class P, int NN> __device__ static void finish_rj(const P& q, const T (&p)[NN], T t, T y, const T* e, T& r, T (&J)[NN]) {
        M::finish_r(q, p, t, y, e, r);
        M::finish_j(q, p, t, e, J);
    }
};

// r_i = p0 exp(-t_i p1) - y_i                                  least_squares.d:347, 360
template <class T, bool INL = false> struct ModelExpDecay2 : ExpModelBase<ModelExpDecay2<T, INL>, T, INL> {
    static constexpr int N = 2, NE = 1; static constexpr bool kHasData = true;
    using Pre = NoPre<T>;
    __device__ static Pre prepare(const T (&)[N]) { return {}; }
    __device__ static void exp_args(const Pre&, const T (&p)[N], T t, T* a) { a[0] = mul_rn(-t, p[1]); }
    __device__ static void finish_r(const Pre&, const T (&p)[N], T, T y, const T* e, T& r) { r = sub_rn(mul_rn(p[0], e[0]), y); }
    __device__ static void finish_j(const Pre&, const T (&p)[N], T t, const T* e, T (&J)[N]) {
        J[0] = e[0]; J[1] = mul_rn(-mul_rn(p[0], t), e[0]);
    }
};

// r_i = p0 exp(-t_i / p1) + p2 - y_i                           least_squares.d:378, 390
template <class T, bool INL = false> struct ModelExpTau3 : ExpModelBase<ModelExpTau3<T, INL>, T, INL> {
    static constexpr int N = 3, NE = 1; static constexpr bool kHasData = true;
    using Pre = NoPre<T>;
    __device__ static Pre prepare(const T (&)[N]) { return {}; }
    __device__ static void exp_args(const Pre&, const T (&p)[N], T t, T* a) { a[0] = div_ni(-t, p[1]); }
    __device__ static void finish_r(const Pre&, const T (&p)[N], T, T y, const T* e, T& r) { r = sub_rn(add_rn(mul_rn(p[0], e[0]), p[2]), y); }
    __device__ static void finish_j(const Pre&, const T (&p)[N], T t, const T* e, T (&J)[N]) {
        J[0] = e[0]; J[1] = div_ni(mul_rn(mul_rn(p[0], e[0]), t), mul_rn(p[1], p[1])); J[2] = (T)1;
    }
};

// r_i = p0 exp(-p1 t_i) + p2 - y_i                             BASELINE configs[0]
template <class T, bool INL = false> struct ModelExpDecay3 : ExpModelBase<ModelExpDecay3<T, INL>, T, INL> {
    static constexpr int N = 3, NE = 1; static constexpr bool kHasData = true;
    using Pre = NoPre<T>;
    __device__ static Pre prepare(const T (&)[N]) { return {}; }
    __device__ static void exp_args(const Pre&, const T (&p)[N], T t, T* a) { a[0] = mul_rn(-p[1], t); }
    __device__ static void finish_r(const Pre&, const T (&p)[N], T, T y, const T* e, T& r) { r = sub_rn(add_rn(mul_rn(p[0], e[0]), p[2]), y); }
    __device__ static void finish_j(const Pre&, const T (&p)[N], T t, const T* e, T (&J)[N]) {
        J[0] = e[0]; J[1] = mul_rn(-mul_rn(p[0], t), e[0]); J[2] = (T)1;
    }
};

// Gaussian peak on a baseline: r_i = A exp(-(t_i - mu)^2 / (2 sigma^2)) + c - y_i, p = (A, mu, sigma, c)
//                                                             BASELINE configs[1]
template <class T> struct Gauss4Pre { T is; };
template <class T, bool INL = false> struct ModelGauss4 : ExpModelBase<ModelGauss4<T, INL>, T, INL> {
    static constexpr int N = 4, NE = 1; static constexpr bool kHasData = true;
    using Pre = Gauss4Pre<T>;
    __device__ static Pre prepare(const T (&p)[N]) { return {rcp_ni(p[2])}; }
    __device__ static void exp_args(const Pre& q, const T (&p)[N], T t, T* a) {
        const T z = mul_rn(sub_rn(t, p[1]), q.is);
        a[0] = mul_rn((T)-0.5, mul_rn(z, z));
    }
    __device__ static void finish_r(const Pre&, const T (&p)[N], T, T y, const T* e, T& r) { r = sub_rn(add_rn(mul_rn(p[0], e[0]), p[3]), y); }
    __device__ static void finish_j(const Pre& q, const T (&p)[N], T t, const T* e, T (&J)[N]) {
        const T z = mul_rn(sub_rn(t, p[1]), q.is);
        const T zz = mul_rn(z, z);
        const T ae = mul_rn(p[0], e[0]);
        J[0] = e[0]; J[1] = mul_rn(mul_rn(ae, z), q.is); J[2] = mul_rn(mul_rn(ae, zz), q.is); J[3] = (T)1;
    }
};

// Sum of exponentials: r_i = sum_k p[2k] exp(-p[2k+1] t_i) - y_i          BASELINE configs[2]
template <class T, int N_, bool INL = false> struct ModelSumExp : ExpModelBase<ModelSumExp<T, N_, INL>, T, INL> {
    static constexpr int N = N_, NE = N_ / 2; static constexpr bool kHasData = true;
    static_assert(N_ % 2 == 0, "sum-of-exponentials has (amplitude, rate) pairs");
    using Pre = NoPre<T>;
    __device__ static Pre prepare(const T (&)[N]) { return {}; }
    __device__ static void exp_args(const Pre&, const T (&p)[N], T t, T* a) {
#pragma unroll
        for (int k = 0; k < N; k += 2) a[k / 2] = mul_rn(-p[k + 1], t);
    }
    __device__ static void finish_r(const Pre&, const T (&p)[N], T, T y, const T* e, T& r) {
        T acc = (T)0;
#pragma unroll
        for (int k = 0; k < N; k += 2) acc = add_rn(acc, mul_rn(p[k], e[k / 2]));
        r = sub_rn(acc, y);
    }
    __device__ static void finish_j(const Pre&, const T (&p)[N], T t, const T* e, T (&J)[N]) {
#pragma unroll
        for (int k = 0; k < N; k += 2) { J[k] = e[k / 2]; J[k + 1] = mul_rn(-mul_rn(p[k], t), e[k / 2]); }
    }

    // ---- finite-difference Jacobian with shared exponentials (lm_mux.cuh) -------------------------------------------
    // The central difference of least_squares.d:1018-1049 evaluates f at x + h e_j and x - h e_j for every parameter j.
    // Perturbing the amplitude p[2k] changes no exponential, perturbing the rate p[2k+1] changes one: the others have the
    // same argument mul_rn(-p[2k'+1], t), hence the same bits, in all 2n evaluations.  They are computed once (the
    // "base" item) and every perturbed residual is re-summed from them with the operations of finish_r in the same
    // order, so each Jacobian entry is bit-identical to the one obtained from 2n full evaluations -- with 3 NE instead
    // of 4 NE^2 exponentials per row.  Work is cut into items of KB = R NE exponentials (R = 4 rows per lane):
    //   item 0                 base: the NE exps of the R rows at p
    //   items 1 .. NE / CPI    CPI = NE / 2 components each: exp(-(b_k + h) t) and exp(-(b_k - h) t) for the R rows;
    //                          delivers the columns 2k and 2k+1 of those components
    static constexpr bool kFDShared = (NE == 2 || NE == 4);
    static constexpr int CPI = NE / 2;
    template <int R> struct FDState { T eb[R][NE]; };
    template <int R>
    __device__ static void fds_args(int it, const T (&p)[N], const T (&tk)[R], const T* xp, const T* xm, T* ea) {
        if (it == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int k = 0; k < NE; ++k) ea[r * NE + k] = mul_rn(-p[2 * k + 1], tk[r]);
        } else {
#pragma unroll
            for (int c = 0; c < CPI; ++c)
#pragma unroll
                for (int r = 0; r < R; ++r) { ea[c * 2 * R + r] = mul_rn(-xp[c], tk[r]); ea[c * 2 * R + R + r] = mul_rn(-xm[c], tk[r]); }
        }
    }
    // xpr / xmr / rtr [c]: x + h, x - h and 1 / (2h) of the rate p[2k+1] of component k = (it - 1) CPI + c; xpa / xma / rta:
    // of its amplitude p[2k] (rt == 0 marks a column that the reference sets to zero, LS:1045-1047).
    // put(j, r, v): J[row r of this lane][j] = v.
    template <int R, class PUT>
    __device__ static void fds_deliver(int it, const T (&p)[N], const T (&yo)[R], const T* ee, FDState<R>& st,
                                       const T* xpr, const T* xmr, const T* rtr, const T* xpa, const T* xma, const T* rta, PUT put) {
        (void)xpr; (void)xmr;
        if (it == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int k = 0; k < NE; ++k) st.eb[r][k] = ee[r * NE + k];
        } else {
#pragma unroll
            for (int c = 0; c < CPI; ++c) {
                const int kc = (it - 1) * CPI + c;
                auto resum = [&](int r, T amp, T enew, bool newAmp, bool newExp) -> T {
                    T acc = (T)0;
#pragma unroll
                    for (int k = 0; k < NE; ++k) {
                        const T a = (newAmp && k == kc) ? amp : p[2 * k];
                        const T e = (newExp && k == kc) ? enew : st.eb[r][k];
                        acc = add_rn(acc, mul_rn(a, e));
                    }
                    return sub_rn(acc, yo[r]);
                };
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const T fpr = resum(r, (T)0, ee[c * 2 * R + r], false, true), fmr = resum(r, (T)0, ee[c * 2 * R + R + r], false, true);
                    put(2 * kc + 1, r, (rtr[c] != (T)0) ? (fpr - fmr) * rtr[c] : (T)0);
                    const T fpa = resum(r, xpa[c], (T)0, true, false), fma_ = resum(r, xma[c], (T)0, true, false);
                    put(2 * kc, r, (rta[c] != (T)0) ? (fpa - fma_) * rta[c] : (T)0);
                }
            }
        }
    }
};

}  // namespace mirb200
