// boxqp_cta.cuh -- CTA-cooperative BOXCQP and ?posvx('E','L') restatement for runtime n (<= 128),
// matrices in shared memory.  Used by the batched BoxQP kernel (one CTA per QP, BASELINE
// configs[4]) and by the control kernel of the large single-problem LM path (n = 128).
//
//   cta_posvx  <- LAPACK dposvx/sposvx FACT='E', UPLO='L' as called at boxcqp.d:194-205, 310-321:
//                 ?poequ/?laqsy equilibration decision, factorisation, solve, ?porfs refinement
//                 (<= 5 sweeps driven by the componentwise backward error), un-scaling.
//                 The factorisation is the square-root-free LDL^T form of Cholesky (same pivots
//                 d_j > 0 test as ?potrf's breakdown test); OpenBLAS' blocked potrf cannot be
//                 matched bit for bit anyway, the parity bar is the stated tolerance.
//   cta_boxqp  <- solveBoxQP!T full overload, boxcqp.d:122-379 (unconstrainedSolution = false).
//                 Active-set sub-systems are compacted through the ascending free-index list,
//                 exactly like boxcqp.d:269-305.
#pragma once
#include "boxqp_small.cuh"   // KBN
#include "common.cuh"

namespace mirb200 {

template <class T, int NT> __device__ __forceinline__ T cta_reduce_max(T v, T* red)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = t_max(v, __shfl_xor_sync(0xffffffffu, v, off));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    T r = red[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) r = t_max(r, red[w]);
    return r;
}
template <class T, int NT> __device__ __forceinline__ T cta_reduce_min(T v, T* red) { return -cta_reduce_max<T, NT>(-v, red); }
template <int NT> __device__ __forceinline__ bool cta_any(bool p) { return __syncthreads_or(p ? 1 : 0) != 0; }

// Shared-memory scratch of one CTA-level QP solve.  nmax = capacity, ldf = nmax + 8 (bank spread).
template <class T> struct CtaQPScratch {
    T* F;        // nmax * ldf   factor (lower, LDL^T: unscaled Schur columns)
    T* sc;       // equilibration scale
    T* dinv;     // 1 / d_j
    T* b;        // right-hand side (scaled)
    T* sx;       // sub-system solution
    T* r;        // refinement residual
    T* la;       // multipliers of lower bounds
    T* mu;       // multipliers of upper bounds
    T* red;      // 32 reduction slots
    T* blk;      // blocked factorisation / solve staging: panel nmaxPad x 9, diagonal tile 8 x 9, 8 broadcast values
    int* idx;    // free-index list (ascending)
    signed char* flag;   // -1 lower, 0 free, +1 upper (boxcqp.d:153-158)
    int ldf;
    __host__ __device__ static size_t bytes(int nmax) {
        return sizeof(T) * ((size_t)nmax * (nmax + 8) + 7 * (size_t)nmax + 32 + blk_elems(nmax)) + sizeof(int) * nmax + ((nmax + 15) & ~15);
    }
    __host__ __device__ static size_t blk_elems(int nmax) { return (size_t)((nmax + 7) & ~7) * 9 + 72 + 8;
    }
    __device__ void carve(void* base, int nmax) {
        ldf = nmax + 8;
        T* p = static_cast<T*>(base);
        F = p; p += (size_t)nmax * ldf;
        sc = p; p += nmax; dinv = p; p += nmax; b = p; p += nmax; sx = p; p += nmax; r = p; p += nmax;
        la = p; p += nmax; mu = p; p += nmax; red = p; p += 32; blk = p; p += blk_elems(nmax);
        idx = reinterpret_cast<int*>(p);
        flag = reinterpret_cast<signed char*>(idx + nmax);
    }
};

// L D L^T solve with warp 0; v in shared memory, overwritten by the solution.  Other warps wait.
template <class T, int NT>
__device__ __forceinline__ void cta_ldl_solve(int s, const T* F, int ldf, const T* dinv, T* v)
{
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int j = 0; j < s; ++j) {                       // forward, unit lower L' = F[i][j] * dinv[j]
            __syncwarp();
            const T t = v[j] * dinv[j];
            for (int i = j + 1 + lane; i < s; i += 32) v[i] -= F[i * ldf + j] * t;
        }
        __syncwarp();
        for (int i = lane; i < s; i += 32) v[i] *= dinv[i]; // D^-1
        for (int j = s - 1; j > 0; --j) {                   // backward with L'^T
            __syncwarp();
            const T xj = v[j];
            for (int i = lane; i < j; i += 32) v[i] -= F[j * ldf + i] * dinv[i] * xj;
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// Blocked variants (8 x 8 tiles).  They perform, for every matrix / vector element, the SAME floating-point operations
// in the SAME order as the column-by-column loops they replace (an element (i,k) is updated with pivots p = 0, 1, ...
// ascending; the blocking only groups eight pivots between barriers), so the results are bit-identical -- but the
// trailing matrix lives in registers (one lower-triangular 8 x 8 tile per thread), and a factorisation needs 3 CTA
// barriers per eight columns instead of one per column with all operands in shared memory.  ncu, round 1: the
// column loop made large_ctl_mid_kernel (n = 128) 0.31-0.47 ms per LM pass, 4 warps mostly waiting at barriers.
// ---------------------------------------------------------------------------------------------------------------
template <int NT> __host__ __device__ constexpr bool cta_blocked_ok(int s) { return ((s + 7) / 8) * ((s + 7) / 8 + 1) / 2 <= NT; }

// LDL^T (square-root free), lower triangle of the matrix in F on entry, factor (unscaled Schur columns) in F and the
// reciprocal pivots in dinv on exit.  Returns 0 or the 1-based index of the first pivot <= 0 (uniform over the CTA).
template <class T, int NT>
__device__ int cta_ldl_factor_blocked(int s, T* F, int ldf, T* dinv, T* blk)
{
    const int tid = threadIdx.x;
    const int nb = (s + 7) >> 3;
    T* Pn = blk;                                   // panel: row i at Pn[i * 9 + p], p < 8
    T* Dg = blk + (size_t)nb * 8 * 9;              // diagonal tile: Dg[k * 9 + p]
    __shared__ int s_info;
    int ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= tid) ++ti;
    const int tj = tid - ti * (ti + 1) / 2;
    const bool owner = ti < nb;                    // this thread owns tile (ti, tj), ti >= tj
    T a[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int i = 8 * ti + r, k = 8 * tj + c;
            a[r][c] = (owner && i < s && k <= i) ? F[i * ldf + k] : ((i == k) ? (T)1 : (T)0);      // identity padding
        }
    if (tid == 0) s_info = 0;
    __syncthreads();

    for (int jt = 0; jt < nb; ++jt) {
        // 1. the diagonal tile: plain right-looking LDL^T in registers
        if (owner && ti == jt && tj == jt) {
            int info = 0;
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const T d = a[p][p];
                if (info == 0 && 8 * jt + p < s && d <= (T)0) info = 8 * jt + p + 1;     // a NaN pivot passes, as in OpenBLAS' potrf
                const T inv = (T)1 / d;
                if (8 * jt + p < s) dinv[8 * jt + p] = inv;
                Dg[64 + 8 + p] = inv;                 // (Dg has 72 + 8 slots: the 8 reciprocal pivots of this block)
#pragma unroll
                for (int i = p + 1; i < 8; ++i) {
                    const T li = a[i][p] * inv;
#pragma unroll
                    for (int k = p + 1; k <= i; ++k) a[i][k] -= li * a[k][p];
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) {
                    Dg[r * 9 + c] = a[r][c];
                    if (8 * jt + r < s) F[(8 * jt + r) * ldf + 8 * jt + c] = a[r][c];
                }
            if (info) s_info = info;
        }
        __syncthreads();
        if (s_info) return s_info;
        // 2. the panel below it: columns of the block, left to right
        if (owner && tj == jt && ti > jt) {
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const T inv = Dg[64 + 8 + p];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const T li = a[r][p] * inv;
#pragma unroll
                    for (int k = p + 1; k < 8; ++k) a[r][k] -= li * Dg[k * 9 + p];
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    Pn[(8 * ti + r) * 9 + c] = a[r][c];
                    if (8 * ti + r < s) F[(8 * ti + r) * ldf + 8 * jt + c] = a[r][c];
                }
        }
        __syncthreads();
        // 3. the trailing tiles: eight rank-1 updates from the panel, operands in registers
        if (owner && tj > jt) {
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const T inv = Dg[64 + 8 + p];
                T li[8], ck[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) { li[r] = Pn[(8 * ti + r) * 9 + p] * inv; ck[r] = Pn[(8 * tj + r) * 9 + p]; }
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) a[r][c] -= li[r] * ck[c];
            }
        }
        __syncthreads();
    }
    return 0;
}

// L D L^T solve, blocked by eight columns: the 8 x 8 diagonal part runs in one thread, the rest is row-parallel.
// Same operations per element and the same order as cta_ldl_solve.
template <class T, int NT>
__device__ __forceinline__ void cta_ldl_solve_blocked(int s, const T* F, int ldf, const T* dinv, T* v, T* blk)
{
    const int tid = threadIdx.x;
    const int nb = (s + 7) >> 3;
    T* bc = blk;                                   // 8 broadcast values
    __syncthreads();
    for (int jb = 0; jb < nb; ++jb) {              // forward: t_j = v_j dinv_j, v_i -= F_ij t_j, j ascending
        const int j0 = 8 * jb;
        if (tid == 0) {
            T vv[8], tt[8];
#pragma unroll
            for (int p = 0; p < 8; ++p) vv[p] = (j0 + p < s) ? v[j0 + p] : (T)0;
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                tt[p] = (j0 + p < s) ? vv[p] * dinv[j0 + p] : (T)0;
#pragma unroll
                for (int i = p + 1; i < 8; ++i) if (j0 + i < s) vv[i] -= F[(j0 + i) * ldf + j0 + p] * tt[p];
            }
#pragma unroll
            for (int p = 0; p < 8; ++p) { bc[p] = tt[p]; if (j0 + p < s) v[j0 + p] = vv[p]; }
        }
        __syncthreads();
        for (int i = j0 + 8 + tid; i < s; i += NT) {
            T vi = v[i];
#pragma unroll
            for (int p = 0; p < 8; ++p) vi -= F[i * ldf + j0 + p] * bc[p];
            v[i] = vi;
        }
        __syncthreads();
    }
    for (int i = tid; i < s; i += NT) v[i] *= dinv[i];     // D^-1
    __syncthreads();
    for (int jb = nb - 1; jb >= 0; --jb) {         // backward: v_i -= F_ji dinv_i x_j, j descending
        const int j0 = 8 * jb;
        if (tid == 0) {
            T vv[8];
#pragma unroll
            for (int p = 0; p < 8; ++p) vv[p] = (j0 + p < s) ? v[j0 + p] : (T)0;
#pragma unroll
            for (int p = 7; p >= 1; --p) {
                if (j0 + p < s) {
                    const T xj = vv[p];
#pragma unroll
                    for (int i = 0; i < p; ++i) vv[i] -= F[(j0 + p) * ldf + j0 + i] * dinv[j0 + i] * xj;
                }
            }
#pragma unroll
            for (int p = 0; p < 8; ++p) { bc[p] = vv[p]; if (j0 + p < s) v[j0 + p] = vv[p]; }
        }
        __syncthreads();
        for (int i = tid; i < j0; i += NT) {
            T vi = v[i];
            const T di = dinv[i];
#pragma unroll
            for (int p = 7; p >= 0; --p) if (j0 + p < s) vi -= F[(j0 + p) * ldf + i] * di * bc[p];
            v[i] = vi;
        }
        __syncthreads();
    }
}

// A(a, c) for a >= c returns the (unscaled) entry of the s x s system.  b: rhs (overwritten by its
// scaled copy), x: solution.  Returns LAPACK info (0, or k>0 = breakdown at pivot k), uniform over the CTA.
// BLK: use the blocked register-tiled factorisation / solves.  They win when ONE CTA works alone and latency is all
// that matters (the LM control kernel, n = 128: 0.47 -> 0.22 ms); the batched BoxQP kernel, with two CTAs per SM
// competing for issue slots, is faster with the all-threads column loop (measured 1.35 M vs 1.17 M QP/s at n = 64).
template <class T, int NT, bool BLK, class AGet>
__device__ int cta_posvx(int s, AGet A, CtaQPScratch<T>& w, T* b, T* x, int* equed_out = nullptr)
{
    const int tid = threadIdx.x;
    T* F = w.F; const int ldf = w.ldf;

    // ?poequ / ?laqsy
    T mn = Num<T>::inf(), mx = -Num<T>::inf();
    for (int a = tid; a < s; a += NT) { const T d = A(a, a); mn = t_min(mn, d); mx = t_max(mx, d); }
    const T smin = cta_reduce_min<T, NT>(mn, w.red);
    const T amax = cta_reduce_max<T, NT>(mx, w.red);
    bool equil = false;
    if (smin > (T)0) {
        const T scond = sqrt_ni(smin) / sqrt_ni(amax);
        equil = !(scond >= (T)0.1 && amax >= Num<T>::small_() && amax <= Num<T>::large_());
    }
    if (equed_out && tid == 0) *equed_out = equil ? 1 : 0;      // (unit tests only: the ?laqsy decision, LAPACK's EQUED)
    for (int a = tid; a < s; a += NT) {
        const T sa = equil ? (T)1 / sqrt_ni(A(a, a)) : (T)1;
        w.sc[a] = sa;
        b[a] = equil ? sa * b[a] : b[a];
    }
    __syncthreads();
    for (int e = tid; e < s * s; e += NT) {
        const int a = e / s, c = e - a * s;
        if (c <= a) F[a * ldf + c] = equil ? (w.sc[c] * w.sc[a]) * A(a, c) : A(a, c);
    }
    __syncthreads();

    // factorisation: right-looking, square-root free.  Column j keeps its unscaled Schur values.
    const bool blocked = BLK && cta_blocked_ok<NT>(s);
    if (blocked) {
        const int info = cta_ldl_factor_blocked<T, NT>(s, F, ldf, w.dinv, w.blk);
        if (info) return info;
    }
    constexpr int TK = 8, TI = NT / TK;
    const int tx = tid % TK, ty = tid / TK;
    for (int j = 0; j < (blocked ? 0 : s); ++j) {
        const T d = F[j * ldf + j];
        if (d <= (T)0) return j + 1;     // uniform (every thread reads the same d).  A NaN pivot passes, as in OpenBLAS' potrf
        const T inv = (T)1 / d;
        for (int i = j + 1 + ty; i < s; i += TI) {
            const T li = F[i * ldf + j] * inv;
            for (int k = j + 1 + tx; k <= i; k += TK) F[i * ldf + k] -= li * F[k * ldf + j];
        }
        if (tid == 0) w.dinv[j] = inv;
        __syncthreads();
    }

    for (int a = tid; a < s; a += NT) x[a] = b[a];
    if (blocked) cta_ldl_solve_blocked<T, NT>(s, F, ldf, w.dinv, x, w.blk);
    else cta_ldl_solve<T, NT>(s, F, ldf, w.dinv, x);

    // ?porfs
    const T eps = Num<T>::lapack_eps();
    const T safe1 = (T)(s + 1) * Num<T>::safmin();
    const T safe2 = safe1 / eps;
    T lstres = (T)3;
    for (int count = 1;; ++count) {
        T q = (T)0;
        for (int a = tid; a < s; a += NT) {
            T ra = b[a], wa = t_abs(b[a]);
            for (int c = 0; c < s; ++c) {
                T aij = (a >= c) ? A(a, c) : A(c, a);
                if (equil) aij = (w.sc[a] * w.sc[c]) * aij;
                ra -= aij * x[c];
                wa += t_abs(aij) * t_abs(x[c]);
            }
            w.r[a] = ra;
            q = t_max(q, (wa > safe2) ? t_abs(ra) / wa : (t_abs(ra) + safe1) / (wa + safe1));
        }
        const T berr = cta_reduce_max<T, NT>(q, w.red);
        if (berr > eps && (T)2 * berr <= lstres && count <= 5) {
            if (blocked) cta_ldl_solve_blocked<T, NT>(s, F, ldf, w.dinv, w.r, w.blk);
            else cta_ldl_solve<T, NT>(s, F, ldf, w.dinv, w.r);
            for (int a = tid; a < s; a += NT) x[a] += w.r[a];
            __syncthreads();
            lstres = berr;
            continue;
        }
        break;
    }
    if (equil) for (int a = tid; a < s; a += NT) x[a] *= w.sc[a];
    __syncthreads();
    return 0;
}

// P(i, j) for i >= j: lower triangle of the QP matrix.  q, l, u, x: length n (shared or global).
// Returns mir_box_qp_status (uniform).  iterations: BOXCQP main-loop count, solves: posvx calls.
template <class T, int NT, bool BLK, class PGet>
__device__ int cta_boxqp(const typename Num<T>::QPSettings& st, int n, PGet P, const T* q, const T* l, const T* u, T* x,
                         CtaQPScratch<T>& w, unsigned& iterations, unsigned& solves)
{
    const int tid = threadIdx.x;
    iterations = 0; solves = 0;
    if (n == 0) return mir_qp_solved;                                          // boxcqp.d:162-163

    for (int i = tid; i < n; i += NT) w.b[i] = -q[i];                          // boxcqp.d:191
    __syncthreads();
    ++solves;
    if (cta_posvx<T, NT, BLK>(n, P, w, w.b, x) != 0) return mir_qp_numericError;    // boxcqp.d:194-213

    bool out = false;                                                          // boxcqp.d:216-219
    for (int i = tid; i < n; i += NT) out = out || !(l[i] <= x[i] && x[i] <= u[i]);
    if (!cta_any<NT>(out)) return mir_qp_solved;

    const unsigned maxIterations = st.maxIterations ? st.maxIterations : (unsigned)n * 10u + 100u;   // boxcqp.d:224-226
    for (int i = tid; i < n; i += NT) { w.la[i] = (T)0; w.mu[i] = (T)0; }
    __shared__ int s_free;

    for (unsigned step = 0; step < maxIterations; ++step) {                    // boxcqp.d:234
        ++iterations;
        __syncthreads();
        for (int i = tid; i < n; i += NT) {                                    // boxcqp.d:239-263
            const T xl = x[i] - l[i];
            const T ux = u[i] - x[i];
            if (xl < (T)0 || (xl < st.relTolerance + st.absTolerance * t_abs(l[i]) && w.la[i] >= (T)0)) {
                w.flag[i] = -1; x[i] = l[i]; w.mu[i] = (T)0;
            } else if (ux < (T)0 || (ux < st.relTolerance + st.absTolerance * t_abs(u[i]) && w.mu[i] >= (T)0)) {
                w.flag[i] = 1; x[i] = u[i]; w.la[i] = (T)0;
            } else {
                w.flag[i] = 0; w.mu[i] = (T)0; w.la[i] = (T)0;
            }
        }
        __syncthreads();
        if (tid == 0) {                                                        // ascending free list
            int s = 0;
            for (int i = 0; i < n; ++i) if (w.flag[i] == 0) w.idx[s++] = i;
            s_free = s;
        }
        __syncthreads();
        const int s = s_free;
        if (s == n) break;                                                     // boxcqp.d:265-266 -> maxIterations

        if (s > 0) {
            for (int a = tid; a < s; a += NT) {                                // boxcqp.d:282-305
                const int i = w.idx[a];
                KBN<T> sum(q[i]);
                for (int j = 0; j < n; ++j) {
                    const int f = w.flag[j];
                    if (f) sum.put(mul_rn((i >= j) ? P(i, j) : P(j, i), f < 0 ? l[j] : u[j]));
                }
                w.b[a] = -sum.sum();
            }
            __syncthreads();
            ++solves;
            const int* idx = w.idx;
            auto Asub = [&](int a, int c) -> T { return P(idx[a], idx[c]); };  // idx ascending: a >= c => idx[a] >= idx[c]
            if (cta_posvx<T, NT, BLK>(s, Asub, w, w.b, w.sx) != 0) return mir_qp_numericError;   // boxcqp.d:310-324
            for (int a = tid; a < s; a += NT) x[w.idx[a]] = w.sx[a];           // boxcqp.d:327-329
            __syncthreads();
        }

        for (int i = tid; i < n; i += NT) if (w.flag[i]) {                     // boxcqp.d:333-337
            T d1 = (T)0, d2 = (T)0;
            for (int j = 0; j < i; ++j) d1 += P(i, j) * x[j];
            for (int j = i; j < n; ++j) d2 += P(j, i) * x[j];
            const T val = d1 + d2 + q[i];
            if (w.flag[i] < 0) w.la[i] = val; else w.mu[i] = -val;
        }
        __syncthreads();
        bool again = false;                                                    // boxcqp.d:339-347
        for (int i = tid; i < n; i += NT) {
            const int f = w.flag[i];
            if (f < 0)      again = again || !(w.la[i] >= (T)0);
            else if (f > 0) again = again || !(w.mu[i] >= (T)0);
            else            again = again || !(x[i] >= l[i] && x[i] <= u[i]);
        }
        if (cta_any<NT>(again)) continue;

        for (int i = tid; i < n; i += NT) x[i] = t_max(t_min(x[i], u[i]), l[i]);   // applyBounds, boxcqp.d:349
        __syncthreads();
        return mir_qp_solved;
    }
    return mir_qp_maxIterations;                                               // boxcqp.d:378
}

}  // namespace mirb200
