// boxqp_small.cuh -- register-resident BOXCQP and LAPACK-?posvx('E','L') restatement for small
// compile-time N (<= 8).  Everything is fully unrolled so the packed matrices live in registers;
// in the warp-per-problem LM kernel every lane of the group executes this redundantly on
// bit-identical inputs (no communication, no divergence).
//
//   posvx_small  <- LAPACK 3.x dposvx/sposvx with FACT='E', UPLO='L' as called at
//                   boxcqp.d:194-205 and :310-321 (third party; steps restated from the LAPACK
//                   sources: ?poequ, ?laqsy, ?potrf (OpenBLAS potf2 dot form), ?potrs, ?porfs).
//                   Only the outputs the reference consumes are produced: x and info
//                   (0 = ok, k>0 = Cholesky breakdown at pivot k).  The reciprocal-condition
//                   estimate only feeds `info = n+1`, which boxcqp.d:212/323 accepts, so it is
//                   not computed.
//   boxqp_small  <- solveBoxQP!T full overload, boxcqp.d:122-379, unconstrainedSolution=false.
//
// The active-set sub-systems (boxcqp.d:282-321) are solved "in place": a fixed variable keeps its
// row/column as an identity row instead of being compacted away.  Every operation that touches a
// free entry then sees exactly the operands of the compacted system plus exact zeros, so the
// result is bit-identical to the compacted solve while all indexing stays static.
#pragma once
#include "common.cuh"

namespace mirb200 {

template <int N> struct FullMask { static constexpr unsigned value = (N >= 32) ? 0xffffffffu : ((1u << N) - 1u); };

// The system matrix is presented as JJ (packed lower, undamped) + lambda on the diagonal
// (least_squares.d:1078-1079 adds lambda to the diagonal before the solve), optionally with
// pinned rows (MASKED) and symmetric diagonal scaling s (equilibration).
template <class T, int N, bool MASKED>
struct SysView {
    const T (&JJ)[N * (N + 1) / 2];
    T lambda;
    unsigned free;
    __device__ __forceinline__ bool is_free(int i) const { return !MASKED || ((free >> i) & 1u); }
    // unscaled entry of the (masked) system matrix, i >= j
    __device__ __forceinline__ T at(int i, int j) const {
        T v = JJ[tri(i, j)];
        if (i == j) v = v + lambda;
        if (MASKED) {
            const bool fi = (free >> i) & 1u, fj = (free >> j) & 1u;
            if (!(fi && fj)) v = (i == j) ? (T)1 : (T)0;
        }
        return v;
    }
};

// Returns LAPACK info (0 ok, k = 1-based pivot where the Cholesky factorisation broke down).
template <class T, int N, bool MASKED>
__device__ __forceinline__ int posvx_small(const T (&JJ)[N * (N + 1) / 2], T lambda, unsigned free,
                                           const T (&b_in)[N], T (&x)[N])
{
    constexpr int NP = N * (N + 1) / 2;
    const SysView<T, N, MASKED> A{JJ, lambda, free};
    T s[N];
    T b[N];
    T f[NP];      // Cholesky factor, packed lower
    T rinv[N];    // 1 / f_ii

    // ---- ?poequ + ?laqsy: decide on equilibration (free rows only) ----
    T smin = Num<T>::inf(), amax = -Num<T>::inf();
#pragma unroll
    for (int i = 0; i < N; ++i) if (A.is_free(i)) { T d = A.at(i, i); smin = t_min(smin, d); amax = t_max(amax, d); }
    bool equil = false;
    if (smin > (T)0) {
        const T scond = t_sqrt(smin) / t_sqrt(amax);
        equil = !(scond >= (T)0.1 && amax >= Num<T>::small_() && amax <= Num<T>::large_());
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        s[i] = (T)1;
        if (equil && A.is_free(i)) s[i] = (T)1 / t_sqrt(A.at(i, i));
        b[i] = equil ? s[i] * b_in[i] : b_in[i];                  // dposvx: B := diag(S) B
        if (MASKED && !A.is_free(i)) b[i] = (T)0;
    }
    // equilibrated entry (dlaqsy: A(i,j) = cj * s(i) * A(i,j)); exact identity when s == 1
    auto a = [&](int i, int j) -> T { return equil ? (s[j] * s[i]) * A.at(i, j) : A.at(i, j); };

    // ---- ?potrf, lower, dot form (OpenBLAS potf2_L) ----
#pragma unroll
    for (int j = 0; j < N; ++j) {
        T ajj = a(j, j);
#pragma unroll
        for (int k = 0; k < j; ++k) ajj -= f[tri(j, k)] * f[tri(j, k)];
        if (!(ajj > (T)0)) return j + 1;                           // ajj <= 0 or NaN
        ajj = t_sqrt(ajj);
        f[tri(j, j)] = ajj;
        const T r = (T)1 / ajj;
        rinv[j] = r;
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            T v = a(i, j);
#pragma unroll
            for (int k = 0; k < j; ++k) v -= f[tri(i, k)] * f[tri(j, k)];
            f[tri(i, j)] = v * r;
        }
    }

    // L L^T solve (?potrs)
    auto solve = [&](T (&v)[N]) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            T acc = v[i];
#pragma unroll
            for (int k = 0; k < i; ++k) acc -= f[tri(i, k)] * v[k];
            v[i] = acc * rinv[i];
        }
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
            T acc = v[i];
#pragma unroll
            for (int k = i + 1; k < N; ++k) acc -= f[tri(k, i)] * v[k];
            v[i] = acc * rinv[i];
        }
    };

#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = b[i];
    solve(x);

    // ---- ?porfs: iterative refinement driven by the componentwise backward error ----
    int nfree = N;
    if (MASKED) nfree = __popc(free & FullMask<N>::value);
    const T eps = Num<T>::lapack_eps();
    const T safe1 = (T)(nfree + 1) * Num<T>::safmin();
    const T safe2 = safe1 / eps;
    T lstres = (T)3;
#pragma unroll 1
    for (int count = 1;; ++count) {
        T r[N], w[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { r[i] = b[i]; w[i] = t_abs(b[i]); }
#pragma unroll
        for (int i = 0; i < N; ++i) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const T aij = (i >= j) ? a(i, j) : a(j, i);
                r[i] -= aij * x[j];
                w[i] += t_abs(aij) * t_abs(x[j]);
            }
        }
        T berr = (T)0;
#pragma unroll
        for (int i = 0; i < N; ++i) if (A.is_free(i)) {
            const T q = (w[i] > safe2) ? t_abs(r[i]) / w[i] : (t_abs(r[i]) + safe1) / (w[i] + safe1);
            berr = t_max(berr, q);
        }
        if (berr > eps && (T)2 * berr <= lstres && count <= 5) {
            solve(r);
#pragma unroll
            for (int i = 0; i < N; ++i) x[i] += r[i];
            lstres = berr;
            continue;
        }
        break;
    }

    if (equil) {
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] *= s[i];
    }
    return 0;
}

// Kahan-Babuska-Neumaier accumulator (mir.math.sum Summation.kbn, used at boxcqp.d:284).
template <class T> struct KBN {
    T s, c;
    __device__ __forceinline__ explicit KBN(T v) : s(v), c((T)0) {}
    __device__ __forceinline__ void put(T v) {
        const T t = add_rn(s, v);
        if (t_abs(s) >= t_abs(v)) c = add_rn(c, add_rn(add_rn(s, -t), v));
        else                      c = add_rn(c, add_rn(add_rn(v, -t), s));
        s = t;
    }
    __device__ __forceinline__ T sum() const { return add_rn(s, c); }
};

struct QPCounters { unsigned solves; unsigned iterations; };

// solveBoxQP, boxcqp.d:122-379 with P = JJ + lambda I (lower triangle only is read).
// Returns mir_box_qp_status.
template <class T, int N>
__device__ __forceinline__ int boxqp_small(const typename Num<T>::QPSettings& st, const T (&JJ)[N * (N + 1) / 2], T lambda,
                                           const T (&q)[N], const T (&l)[N], const T (&u)[N], T (&x)[N], QPCounters& cnt)
{
    constexpr unsigned FULL = FullMask<N>::value;
    auto P = [&](int i, int j) -> T {          // symmetric read through the lower triangle
        T v = JJ[trisym(i, j)];
        return (i == j) ? v + lambda : v;
    };

    {   // unconstrained minimiser, boxcqp.d:168-214
        T b[N];
#pragma unroll
        for (int i = 0; i < N; ++i) b[i] = -q[i];
        ++cnt.solves;
        if (posvx_small<T, N, false>(JJ, lambda, FULL, b, x) != 0) return mir_qp_numericError;
    }
    bool inside = true;                        // boxcqp.d:216-219
#pragma unroll
    for (int i = 0; i < N; ++i) inside = inside && (l[i] <= x[i] && x[i] <= u[i]);
    if (inside) return mir_qp_solved;

    unsigned maxIterations = st.maxIterations ? st.maxIterations : (unsigned)N * 10u + 100u;   // boxcqp.d:224-226
    T la[N], mu[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { la[i] = (T)0; mu[i] = (T)0; }

#pragma unroll 1
    for (unsigned step = 0; step < maxIterations; ++step) {                                    // boxcqp.d:234
        ++cnt.iterations;
        unsigned lo = 0, up = 0;               // flags: bit set in `lo` = at lower bound, in `up` = at upper bound
#pragma unroll
        for (int i = 0; i < N; ++i) {          // boxcqp.d:239-263
            const T xl = x[i] - l[i];
            const T ux = u[i] - x[i];
            if (xl < (T)0 || (xl < st.relTolerance + st.absTolerance * t_abs(l[i]) && la[i] >= (T)0)) {
                lo |= 1u << i; x[i] = l[i]; mu[i] = (T)0;
            } else if (ux < (T)0 || (ux < st.relTolerance + st.absTolerance * t_abs(u[i]) && mu[i] >= (T)0)) {
                up |= 1u << i; x[i] = u[i]; la[i] = (T)0;
            } else {
                mu[i] = (T)0; la[i] = (T)0;
            }
        }
        const unsigned fixed = lo | up;
        const unsigned free = FULL & ~fixed;
        if (free == FULL) break;               // boxcqp.d:265-266 -> falls out with maxIterations

        if (free) {                            // reduced system, boxcqp.d:282-329
            T b[N], sx[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                KBN<T> sum(q[i]);
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    if ((fixed >> j) & 1u) {
                        const T bound = ((lo >> j) & 1u) ? l[j] : u[j];
                        sum.put(mul_rn(P(i, j), bound));
                    }
                }
                b[i] = -sum.sum();
            }
            ++cnt.solves;
            if (posvx_small<T, N, true>(JJ, lambda, free, b, sx) != 0) return mir_qp_numericError;
#pragma unroll
            for (int i = 0; i < N; ++i) if ((free >> i) & 1u) x[i] = sx[i];
        }

#pragma unroll
        for (int i = 0; i < N; ++i) if ((fixed >> i) & 1u) {      // multipliers, boxcqp.d:333-337
            T d1 = (T)0, d2 = (T)0;
#pragma unroll
            for (int j = 0; j < i; ++j) d1 += P(i, j) * x[j];
#pragma unroll
            for (int j = i; j < N; ++j) d2 += P(j, i) * x[j];
            const T val = d1 + d2 + q[i];
            if ((lo >> i) & 1u) la[i] = val; else mu[i] = -val;
        }

        bool again = false;                    // boxcqp.d:339-347
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if ((lo >> i) & 1u)      again = again || !(la[i] >= (T)0);
            else if ((up >> i) & 1u) again = again || !(mu[i] >= (T)0);
            else                     again = again || !(x[i] >= l[i] && x[i] <= u[i]);
        }
        if (again) continue;

#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = t_max(t_min(x[i], u[i]), l[i]);   // applyBounds, boxcqp.d:349, 404-410
        return mir_qp_solved;
    }
    return mir_qp_maxIterations;               // boxcqp.d:378
}

}  // namespace mirb200
