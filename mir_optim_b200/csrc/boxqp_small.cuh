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
#include "repro_math.cuh"

namespace mirb200 {

template <int N> struct FullMask { static constexpr unsigned value = (N >= 32) ? 0xffffffffu : ((1u << N) - 1u); };

// The system matrix is JJ (packed lower, undamped) + lambda on the diagonal
// (least_squares.d:1078-1079 adds lambda to the diagonal before the solve); rows/columns whose
// bit is clear in `free` are pinned to the identity.  Returns LAPACK info (0 ok, k = 1-based
// pivot where the Cholesky factorisation broke down).  One instantiation serves both the
// unconstrained solve (free = all ones) and the active-set solves, to keep the code small.
template <class T, int N>
__device__ __forceinline__ int posvx_small(const T (&JJ)[N * (N + 1) / 2], T lambda, unsigned free,
                                           const T (&b_in)[N], T (&x)[N], bool* equed_out = nullptr)
{
    constexpr int NP = N * (N + 1) / 2;
    T a[NP];      // (equilibrated) system matrix, packed lower
    T f[NP];      // Cholesky factor, packed lower
    T b[N], s[N], rinv[N];

    // ---- masked system + ?poequ over the free rows ----
    T smin = Num<T>::inf(), amax = -Num<T>::inf();
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const bool fi = (free >> i) & 1u;
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const bool fj = (free >> j) & 1u;
            T v = JJ[tri(i, j)];
            if (i == j) v = v + lambda;
            if (!(fi && fj)) v = (i == j) ? (T)1 : (T)0;
            a[tri(i, j)] = v;
        }
        b[i] = fi ? b_in[i] : (T)0;
        if (fi) { smin = t_min(smin, a[tri(i, i)]); amax = t_max(amax, a[tri(i, i)]); }
    }
    // ---- ?laqsy: equilibrate when badly scaled ----
    bool equil = false;
    if (smin > (T)0) {
        // dlaqsy: no scaling iff scond = sqrt(smin)/sqrt(amax) >= 0.1 (and amax in range).  The two square
        // roots and the division are only evaluated when smin/amax is within 2 % of the 0.01 boundary.
        bool wellScaled;
        if (smin >= (T)0.0102 * amax) wellScaled = true;
        else if (smin <= (T)0.0098 * amax) wellScaled = false;
        else wellScaled = div_ni(sqrt_ni(smin), sqrt_ni(amax)) >= (T)0.1;
        equil = !(wellScaled && amax >= Num<T>::small_() && amax <= Num<T>::large_());
    }
    if (equed_out) *equed_out = equil;          // (unit tests only: the ?laqsy decision, LAPACK's EQUED)
    if (equil) {
#pragma unroll
        for (int i = 0; i < N; ++i) s[i] = ((free >> i) & 1u) ? rcp_ni(sqrt_ni(a[tri(i, i)])) : (T)1;
#pragma unroll
        for (int i = 0; i < N; ++i) {
#pragma unroll
            for (int j = 0; j <= i; ++j) a[tri(i, j)] = (s[j] * s[i]) * a[tri(i, j)];   // dlaqsy: cj * s(i) * A(i,j)
            b[i] = s[i] * b[i];                                                        // dposvx: B := diag(S) B
        }
    }

    // ---- ?potrf, lower, dot form (OpenBLAS potf2_L) ----
#pragma unroll
    for (int j = 0; j < N; ++j) {
        T ajj = a[tri(j, j)];
#pragma unroll
        for (int k = 0; k < j; ++k) ajj -= f[tri(j, k)] * f[tri(j, k)];
        if (ajj <= (T)0) return j + 1;          // breakdown; a NaN pivot passes through, as in OpenBLAS' potrf (potf2: `ajj <= 0`)
        ajj = sqrt_ni(ajj);
        f[tri(j, j)] = ajj;
        const T r = rcp_ni(ajj);
        rinv[j] = r;
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            T v = a[tri(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) v -= f[tri(i, k)] * f[tri(j, k)];
            f[tri(i, j)] = v * r;
        }
    }

    // ---- ?potrs, then ?porfs: iterative refinement driven by the componentwise backward error.
    // One loop: pass 0 solves for b, later passes solve for the residual and correct x.
    const int nfree = __popc(free & FullMask<N>::value);
    const T eps = Num<T>::lapack_eps();
    const T safe1 = (T)(nfree + 1) * Num<T>::safmin();
    const T safe2 = safe1 * ((T)1 / Num<T>::lapack_eps());       // safe1 / eps, exact: 1/eps is a power of two (constant-folded)
    T lstres = (T)3;
    T v[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { v[i] = b[i]; x[i] = (T)0; }
#pragma unroll 1
    for (int count = 0;; ++count) {
        // L L^T v = v
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
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] += v[i];                  // pass 0: x = 0 + A^-1 b (exact add)

        // residual r = b - A x and |b| + |A||x|
        T w[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { v[i] = b[i]; w[i] = t_abs(b[i]); }
#pragma unroll
        for (int i = 0; i < N; ++i) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const T aij = a[trisym(i, j)];
                v[i] -= aij * x[j];
                w[i] += t_abs(aij) * t_abs(x[j]);
            }
        }
        // berr = max_i num_i / den_i (dporfs).  The maximising row is found by cross-multiplication, so
        // one division per sweep instead of N (the maximum itself is the same quotient LAPACK forms).
        T bn = (T)0, bd = (T)1;
#pragma unroll
        for (int i = 0; i < N; ++i) if ((free >> i) & 1u) {
            const bool big = w[i] > safe2;
            const T num = big ? t_abs(v[i]) : t_abs(v[i]) + safe1;
            const T den = big ? w[i] : w[i] + safe1;
            if (num * bd > bn * den) { bn = num; bd = den; }
        }
        const T berr = div_ni(bn, bd);
        if (!(berr > eps && (T)2 * berr <= lstres && count < 5)) break;   // dporfs: at most ITMAX = 5 corrections
        lstres = berr;
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
// Returns mir_box_qp_status.  The loop is arranged so that the unconstrained solve
// (boxcqp.d:168-214) and the active-set solves (boxcqp.d:310-321) share one posvx call site.
template <class T, int N>
__device__ __forceinline__ int boxqp_small(const typename Num<T>::QPSettings& st, const T (&JJ)[N * (N + 1) / 2], T lambda,
                                           const T (&q)[N], const T (&l)[N], const T (&u)[N], T (&x)[N], QPCounters& cnt)
{
    constexpr unsigned FULL = FullMask<N>::value;
    auto P = [&](int i, int j) -> T {          // symmetric read through the lower triangle
        T v = JJ[trisym(i, j)];
        return (i == j) ? v + lambda : v;
    };
    const unsigned maxIterations = st.maxIterations ? st.maxIterations : (unsigned)N * 10u + 100u;   // boxcqp.d:224-226

    T b[N], la[N], mu[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { b[i] = -q[i]; la[i] = (T)0; mu[i] = (T)0; }                        // boxcqp.d:191, 231-232
    unsigned free = FULL, lo = 0, up = 0;
    bool first = true;
    unsigned step = 0;

#pragma unroll 1
    for (;;) {
        if (free) {
            T sx[N];
            ++cnt.solves;
            if (posvx_small<T, N>(JJ, lambda, free, b, sx) != 0) return mir_qp_numericError;         // boxcqp.d:212, 323
#pragma unroll
            for (int i = 0; i < N; ++i) if ((free >> i) & 1u) x[i] = sx[i];                          // boxcqp.d:327-329
        }
        if (first) {
            first = false;
            bool inside = true;                                                                      // boxcqp.d:216-219
#pragma unroll
            for (int i = 0; i < N; ++i) inside = inside && (l[i] <= x[i] && x[i] <= u[i]);
            if (inside) return mir_qp_solved;
        } else {
            const unsigned fixed = lo | up;
            bool again = false;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if ((fixed >> i) & 1u) {                                                             // multipliers, boxcqp.d:333-337
                    T d1 = (T)0, d2 = (T)0;
#pragma unroll
                    for (int j = 0; j < i; ++j) d1 += P(i, j) * x[j];
#pragma unroll
                    for (int j = i; j < N; ++j) d2 += P(j, i) * x[j];
                    const T val = d1 + d2 + q[i];
                    if ((lo >> i) & 1u) { la[i] = val; again = again || !(val >= (T)0); }            // boxcqp.d:343
                    else                { mu[i] = -val; again = again || !(-val >= (T)0); }          // boxcqp.d:344
                } else {
                    again = again || !(x[i] >= l[i] && x[i] <= u[i]);                                // boxcqp.d:345
                }
            }
            if (!again) {
#pragma unroll
                for (int i = 0; i < N; ++i) x[i] = t_max(t_min(x[i], u[i]), l[i]);                   // applyBounds, boxcqp.d:349
                return mir_qp_solved;
            }
            ++step;
        }
        if (step >= maxIterations) return mir_qp_maxIterations;                                      // boxcqp.d:378
        ++cnt.iterations;

        lo = 0; up = 0;                        // flags: bit in `lo` = at lower bound, in `up` = at upper bound
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
        free = FULL & ~fixed;
        if (free == FULL) return mir_qp_maxIterations;   // boxcqp.d:265-266: `break` falls out to :378

#pragma unroll
        for (int i = 0; i < N; ++i) {          // reduced right-hand side, boxcqp.d:282-305
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
    }
}

}  // namespace mirb200
