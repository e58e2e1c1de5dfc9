// lm_mux.cuh -- batched Levenberg-Marquardt for problems with up to 8 parameters and up to 128 residuals
// (BASELINE configs[2]: 8-parameter sums of exponentials, finite-difference Jacobian), four problems per warp.
// Follows optimizeLeastSquaresImplGeneric!T, least_squares.d:877-1176, and solveBoxQP, boxcqp.d:122-379 (cites inline).
//
// Why a third batched kernel.  At n = 8 the n-sized part of a pass (the 8 x 8 ?posvx + BOXCQP + lambda control) is as
// long as the m-sized part.  One warp per problem (lm_small.cuh) runs it 32 times redundantly and needs the packed
// system, its equilibrated copy and its factor (3 x 36 values) in every lane: 255 registers, 2 KB of spills, 145 KB of
// code, 0.22 M fits/s (ncu, round 1: 2.9 G local-memory instructions per 65k fits, top stall no_instruction).  One thread
// per problem (lm_tpp.cuh) has no room for an 8 x 8 system plus an 8-column Jacobian slab either.  Here each phase gets
// the lane layout that suits it:
//
//   * ROW phases (residual evaluation, finite-difference / analytic Jacobian, Broyden update LS:999-1006, J^T J and
//     J^T y, LS:1052, 1065) are WARP-cooperative: all 32 lanes work on ONE of the warp's four problems at a time, lane L
//     owning rows L, L+32, L+64, L+96.  A warp does exactly the row work its problems ask for -- a problem that needs a
//     fresh finite-difference Jacobian (16 evaluations) does not hold three others that only need a trial evaluation
//     in lock step, as sub-warp groups would.  Current / trial residuals and the observations of the four problems
//     live in registers (3 x 4 x 4 values per lane); the Jacobian lives in shared memory as [slot][parameter][row]
//     (8 KB per problem in double), conflict-free for both the column writes of the finite differences and the row
//     reads of the update.
//   * N-SIZED phases run in four 8-lane GROUPS at once, group g on problem g, lane i owning parameter i: row i of the
//     system matrix, x_i, its bounds, its multipliers.  The ?posvx restatement is a right-looking Cholesky with one
//     broadcast per pivot / column element (__shfl_sync inside the group): per matrix element the same operations in
//     the same order as the dot form of boxqp_small.cuh (element (i,k) takes its updates with pivot j ascending), 8
//     values of the matrix per lane instead of 108, no spills, and the square roots / reciprocals of the eight pivots
//     are the only serial chain.  BOXCQP's per-variable logic (flags, multipliers, KBN right-hand side, BQ:239-347) is
//     naturally lane-parallel; its any / all tests are group votes.
//
// Equivalences (bit-exact w.r.t. this file's own arithmetic, as in lm_small.cuh): J^T J is rebuilt only when J changed;
// a trial point equal to x skips the evaluation (fCalls still counts it); the provably inert lambda-overflow tail is
// replayed as a scalar recurrence (tail_is_inert, lm_small.cuh -- restated here for the distributed layout).
#pragma once
#include "lm_small.cuh"

namespace mirb200 {

constexpr int MUX_SLOTS = 4;            // problems per warp
constexpr int MUX_G = 8;                // lanes per problem in the n-sized phases: n <= 8
constexpr int MUX_R = 4;                // rows per lane in the row phases (row = lane + 32 k): m <= 128
constexpr int MUX_MMAX = 32 * MUX_R;
constexpr int MUX_WARPS = 2;            // warps per CTA (warps never talk to each other)
constexpr int MUX_RED = 11;             // values per round of the cross-lane reduction (n = 8: 36 + 8 = 4 x 11)
enum { MUX_JAC_NONE = 0, MUX_JAC_BROYDEN = 1, MUX_JAC_FRESH = 2, MUX_EVAL = 4, MUX_EVAL_INIT = 8 };

template <class T> struct MuxWarpSmem {
    T J[MUX_SLOTS][MUX_G][MUX_MMAX];    // Jacobian of each slot, [parameter][row]
    T JJ[MUX_SLOTS][MUX_G * MUX_G];     // J^T J, full symmetric storage, undamped
    T Jy[MUX_SLOTS][MUX_G];             // J^T y
    T red[MUX_RED][33];                 // reduction scratch, one padded row per value
};

// ---- group (8-lane) collectives; every lane of the group receives the same bits -------------------------------------
template <class T> __device__ __forceinline__ T gshfl(unsigned gmask, T v, int src) { return __shfl_sync(gmask, v, src, MUX_G); }
template <class T> __device__ __forceinline__ T gmax8(unsigned gmask, T v)
{
#pragma unroll
    for (int off = MUX_G / 2; off > 0; off >>= 1) v = t_max(v, __shfl_xor_sync(gmask, v, off, MUX_G));
    return v;
}
template <class T> __device__ __forceinline__ T gmin8(unsigned gmask, T v)
{
#pragma unroll
    for (int off = MUX_G / 2; off > 0; off >>= 1) v = t_min(v, __shfl_xor_sync(gmask, v, off, MUX_G));
    return v;
}
template <class T> __device__ __forceinline__ T gsum8(unsigned gmask, T v)
{
#pragma unroll
    for (int off = MUX_G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(gmask, v, off, MUX_G);
    return v;
}

// register arrays indexed by a run-time slot: selects / predicated moves instead of local memory
template <class T> __device__ __forceinline__ T slot_get(const T (&a)[MUX_SLOTS][MUX_R], int s, int k)
{
    T r = a[0][k];
    r = (s == 1) ? a[1][k] : r; r = (s == 2) ? a[2][k] : r; r = (s == 3) ? a[3][k] : r;
    return r;
}
template <class T> __device__ __forceinline__ void slot_put(T (&a)[MUX_SLOTS][MUX_R], int s, int k, T v)
{
    a[0][k] = (s == 0) ? v : a[0][k]; a[1][k] = (s == 1) ? v : a[1][k];
    a[2][k] = (s == 2) ? v : a[2][k]; a[3][k] = (s == 3) ? v : a[3][k];
}

__host__ __device__ constexpr int untri_row(int v) { int i = 0; while ((i + 1) * (i + 2) / 2 <= v) ++i; return i; }

// ---------------------------------------------------------------------------------------------------------------------
// LAPACK ?posvx(FACT='E', UPLO='L') as called at boxcqp.d:194-205 / 310-321, distributed over the 8 lanes of a group:
// lane gl owns row gl of the system.  Steps as in posvx_small (boxqp_small.cuh): ?poequ / ?laqsy over the free rows,
// Cholesky, ?potrs, ?porfs refinement (<= 5 sweeps on the componentwise backward error), un-scaling.  Rows / columns
// whose bit is clear in `free` (fixed variables of the active set, and lanes >= n) are pinned to the identity, so the
// free entries see exactly the operands of the compacted system plus exact zeros.
//   Prow   row gl of P = J^T J + lambda I (all 8 columns, unmasked), Pdiag = Prow[gl]
// Returns LAPACK info (0, or k > 0: factorisation broke down at pivot k), uniform over the group.
// ---------------------------------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ int posvx_dist(unsigned gmask, int gl, const T (&Prow)[MUX_G], T Pdiag, unsigned free, T b_in, T& x_out)
{
    constexpr int G = MUX_G;
    const bool fi = (free >> gl) & 1u;
    T a[G], a0[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const bool fj = (free >> j) & 1u;
        a[j] = (fi && fj) ? Prow[j] : ((j == gl) ? (T)1 : (T)0);
    }
    T b = fi ? b_in : (T)0;
    const T d = fi ? Pdiag : (T)1;
    // ?poequ over the free rows, ?laqsy decision
    const T smin = gmin8(gmask, fi ? d : Num<T>::inf());
    const T amax = gmax8(gmask, fi ? d : -Num<T>::inf());
    bool equil = false;
    if (smin > (T)0) {
        bool wellScaled;                       // scond = sqrt(smin) / sqrt(amax) >= 0.1; the roots only near the boundary
        if (smin >= (T)0.0102 * amax) wellScaled = true;
        else if (smin <= (T)0.0098 * amax) wellScaled = false;
        else wellScaled = div_ni(sqrt_ni(smin), sqrt_ni(amax)) >= (T)0.1;
        equil = !(wellScaled && amax >= Num<T>::small_() && amax <= Num<T>::large_());
    }
    T s = (T)1;
    if (equil) {
        s = fi ? rcp_ni(sqrt_ni(d)) : (T)1;
#pragma unroll
        for (int j = 0; j < G; ++j) { const T sj = gshfl(gmask, s, j); a[j] = (sj * s) * a[j]; }   // dlaqsy: cj * s(i) * A(i,j)
        b = s * b;                                                                                 // dposvx: B := diag(S) B
    }
#pragma unroll
    for (int j = 0; j < G; ++j) a0[j] = a[j];

    // ?potrf, lower, right-looking: a[] becomes row gl of L (columns <= gl), ft[k] = L(k, gl) (row gl of L^T)
    T ft[G];
    T rinv = (T)0;
#pragma unroll
    for (int j = 0; j < G; ++j) ft[j] = (T)0;
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const T ajj = gshfl(gmask, a[j], j);
        if (ajj <= (T)0) return j + 1;         // breakdown (a NaN pivot passes through, as in OpenBLAS' potf2: `ajj <= 0`)
        const T ljj = sqrt_ni(ajj);
        const T r = rcp_ni(ljj);
        const T fij = a[j] * r;                // L(gl, j) for gl > j
        if (gl == j) { a[j] = ljj; rinv = r; } else a[j] = fij;
#pragma unroll
        for (int k = j + 1; k < G; ++k) {
            const T fk = gshfl(gmask, fij, k); // L(k, j)
            a[k] = fma(-fij, fk, a[k]);        // a(gl, k) -= L(gl, j) L(k, j): the part k <= gl is the matrix, the rest is never read
            if (gl == j) ft[k] = fk;
        }
    }

    auto solve = [&](T v) -> T {               // L L^T z = v, lane gl in: v_gl, out: z_gl
        T acc = v;
#pragma unroll
        for (int k = 0; k < G; ++k) {
            const T cand = acc * rinv;
            const T vk = gshfl(gmask, cand, k);
            if (gl == k) acc = cand; else if (gl > k) acc = fma(-a[k], vk, acc);
        }
#pragma unroll
        for (int k = G - 1; k >= 0; --k) {
            const T cand = acc * rinv;
            const T xk = gshfl(gmask, cand, k);
            if (gl == k) acc = cand; else if (gl < k) acc = fma(-ft[k], xk, acc);
        }
        return acc;
    };

    // ?potrs, then ?porfs: pass 0 solves for b, later passes solve for the residual and correct x
    const int nfree = __popc(free & 0xffu);
    const T eps = Num<T>::lapack_eps();
    const T safe1 = (T)(nfree + 1) * Num<T>::safmin();
    const T safe2 = safe1 * ((T)1 / Num<T>::lapack_eps());
    T lstres = (T)3, x = (T)0, v = b;
#pragma unroll 1
    for (int count = 0;; ++count) {
        v = solve(v);
        x += v;
        T r = b, w = t_abs(b);
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const T xj = gshfl(gmask, x, j);
            r = fma(-a0[j], xj, r);
            w = fma(t_abs(a0[j]), t_abs(xj), w);
        }
        T qv = (T)0;                           // dporfs: berr = max_i |r_i| / (|b| + |A||x|)_i with the safe1 / safe2 guard
        if (fi) {
            const bool big = w > safe2;
            qv = div_ni(big ? t_abs(r) : t_abs(r) + safe1, big ? w : w + safe1);
        }
        const T berr = gmax8(gmask, qv);
        v = r;
        if (!(berr > eps && (T)2 * berr <= lstres && count < 5)) break;      // at most ITMAX = 5 corrections
        lstres = berr;
    }
    x_out = equil ? x * s : x;
    return 0;
}

// solveBoxQP, boxcqp.d:122-379 (unconstrainedSolution = false) with P = J^T J + lambda I, lane gl owning variable gl.
// JJrow: row gl of the undamped J^T J.  q, l, u, x: this lane's entries (lanes >= n: q = 0, l = -inf, u = +inf).
// Returns mir_box_qp_status, uniform over the group.
template <class T, int N>
__device__ __forceinline__ int boxqp_dist(unsigned gmask, int gl, int gshift, const typename Num<T>::QPSettings& st,
                                          const T (&JJrow)[MUX_G], T lambda, T q, T l, T u, T& x, QPCounters& cnt)
{
    constexpr int G = MUX_G;
    constexpr unsigned FULL = (1u << N) - 1u;
    const bool valid = gl < N;
    T Prow[G];
    T Pdiag = (T)0;
#pragma unroll
    for (int j = 0; j < G; ++j) {                                                        // LS:1078-1079: lambda on the diagonal
        Prow[j] = (j == gl) ? JJrow[j] + lambda : JJrow[j];
        Pdiag = (j == gl) ? Prow[j] : Pdiag;
    }
    const unsigned maxIterations = st.maxIterations ? st.maxIterations : (unsigned)N * 10u + 100u;   // BQ:224-226
    T b = -q, la = (T)0, mu = (T)0;                                                      // BQ:191, 231-232
    unsigned free = FULL, lo = 0, up = 0;
    bool first = true;
    unsigned step = 0;
    x = (T)0;
#pragma unroll 1
    for (;;) {
        if (free) {
            T sx;
            ++cnt.solves;
            if (posvx_dist<T>(gmask, gl, Prow, Pdiag, free, b, sx) != 0) return mir_qp_numericError;   // BQ:212, 323
            if ((free >> gl) & 1u) x = sx;                                               // BQ:327-329
        }
        if (first) {
            first = false;                                                               // BQ:216-219
            if (__all_sync(gmask, !valid || (l <= x && x <= u))) return mir_qp_solved;
        } else {
            T d1 = (T)0, d2 = (T)0;                                                      // multipliers, BQ:333-337
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const T xj = gshfl(gmask, x, j);
                if (j < gl) d1 = fma(Prow[j], xj, d1); else d2 = fma(Prow[j], xj, d2);
            }
            const T val = d1 + d2 + q;
            bool bad;
            if ((lo >> gl) & 1u)      { la = val;  bad = !(val >= (T)0); }               // BQ:343
            else if ((up >> gl) & 1u) { mu = -val; bad = !(-val >= (T)0); }              // BQ:344
            else bad = valid && !(x >= l && x <= u);                                     // BQ:345
            if (!__any_sync(gmask, bad)) {
                x = t_max(t_min(x, u), l);                                               // applyBounds, BQ:349
                return mir_qp_solved;
            }
            ++step;
        }
        if (step >= maxIterations) return mir_qp_maxIterations;                          // BQ:378
        ++cnt.iterations;

        int mine = 0;                                                                    // BQ:239-263
        if (valid) {
            const T xl = x - l, ux = u - x;
            if (xl < (T)0 || (xl < st.relTolerance + st.absTolerance * t_abs(l) && la >= (T)0)) { mine = 1; x = l; mu = (T)0; }
            else if (ux < (T)0 || (ux < st.relTolerance + st.absTolerance * t_abs(u) && mu >= (T)0)) { mine = 2; x = u; la = (T)0; }
            else { mu = (T)0; la = (T)0; }
        }
        lo = (__ballot_sync(gmask, mine == 1) >> gshift) & FULL;
        up = (__ballot_sync(gmask, mine == 2) >> gshift) & FULL;
        const unsigned fixed = lo | up;
        free = FULL & ~fixed;
        if (free == FULL) return mir_qp_maxIterations;                                   // BQ:265-266: `break` falls out to :378

        const T bound = (mine == 1) ? l : u;                                             // reduced right-hand side, BQ:282-305
        KBN<T> sum(q);
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const T bj = gshfl(gmask, bound, j);
            if ((fixed >> j) & 1u) sum.put(mul_rn(Prow[j], bj));
        }
        b = -sum.sum();
    }
}

template <class Model, class T, bool FD>
__global__ void __launch_bounds__(MUX_WARPS * 32, 1)
lm_mux_kernel(const typename Num<T>::Settings st, const SmallBatchArgs args)
{
    constexpr int N = Model::N, G = MUX_G, S = MUX_SLOTS, R = MUX_R;
    constexpr int NP = N * (N + 1) / 2, K = NP + N;
    constexpr int NE = Model::NE;
    static_assert(N <= G, "lm_mux_kernel: at most 8 parameters");
    using Result = typename Num<T>::Result;
    constexpr unsigned FULLW = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char mux_smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int grp = lane >> 3, gl = lane & 7, gshift = grp * 8;
    const unsigned gmask = 0xffu << gshift;
    MuxWarpSmem<T>& sm = reinterpret_cast<MuxWarpSmem<T>*>(mux_smem_raw)[wid];
    const bool valid = gl < N;
    const int m = (int)args.m;
    const bool gridPerProblem = (args.flags & MIR_MODEL_GRID_PER_PROBLEM) != 0;
    const bool tailShortcut = (args.flags & MIR_MODEL_NO_TAIL_SHORTCUT) == 0;
    const T* __restrict__ tptr = static_cast<const T*>(args.t);
    const T* __restrict__ yptr = static_cast<const T*>(args.y);

    for (int e = lane; e < S * G * G; e += 32) (&sm.JJ[0][0])[e] = (T)0;     // rows / columns >= n stay zero for good
    for (int e = lane; e < S * G; e += 32) (&sm.Jy[0][0])[e] = (T)0;
    __syncwarp();

    // ---- warp-role state: my rows (lane + 32 k) of the four problems of this warp
    T tts[R];
#pragma unroll
    for (int k = 0; k < R; ++k) { const int row = lane + 32 * k; tts[k] = (Model::kHasData && !gridPerProblem && row < m) ? tptr[row] : (T)0; }
    T yv[S][R], mb[S][R], yo[S][R];           // y (current residuals), mBuffer (trial residuals), observations
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int k = 0; k < R; ++k) { yv[s][k] = (T)0; mb[s][k] = (T)0; yo[s][k] = (T)0; }

    // ---- group-role state: problem `grp` of this warp, parameter gl.  Scalars are bit-identical in the 8 lanes.
    bool active = false, retired = false, init = false;
    unsigned long long prob = 0;
    T x = (T)0, xt = (T)0, dX = (T)0, Jy = (T)0, lo = -Num<T>::inf(), up = Num<T>::inf();
    T lambda = (T)0, mu = (T)1, residual = Num<T>::inf(), deltaX_dot = (T)0, nd = (T)0, trial = (T)0;
    unsigned age = 0, maxAge = 1, iterations = 0, fCalls = 0, gCalls = 0;
    int status = mir_ls_numericError;
    bool needJacobian = false, fConverged = false;
    unsigned sPasses = 0, sAccepted = 0, sFresh = 0, sBroyden = 0, sEvals = 0, sSolves = 0, sQPIt = 0, sProblems = 0;

    // Residuals of MY rows of one problem at parameter vector p: out[k] = r_row(p), returns the lane's partial ||r||^2.
    // The exps of two rows (2 NE independent chains) go through one interleaved exp_repro_many call.
    auto eval_rows = [&](const T (&p)[N], const T (&tk)[R], const T (&yk)[R], T (&out)[R]) -> T {
        const typename Model::Pre pre = Model::prepare(p);
        T part = (T)0;
#pragma unroll
        for (int k = 0; k < R; k += 2) {
            T r0, r1;
            if constexpr (NE > 0) {
                T ea[2 * NE], ee[2 * NE];
                Model::exp_args(pre, p, tk[k], ea); Model::exp_args(pre, p, tk[k + 1], ea + NE);
                exp_repro_many<2 * NE>(ea, ee);
                Model::finish_r(pre, p, tk[k], yk[k], ee, r0); Model::finish_r(pre, p, tk[k + 1], yk[k + 1], ee + NE, r1);
            } else {
                r0 = Model::residual(pre, p, lane + 32 * k, tk[k], yk[k]); r1 = Model::residual(pre, p, lane + 32 * (k + 1), tk[k + 1], yk[k + 1]);
            }
            r0 = (lane + 32 * k < m) ? r0 : (T)0; r1 = (lane + 32 * (k + 1) < m) ? r1 : (T)0;
            out[k] = r0; out[k + 1] = r1;
            part += r0 * r0; part += r1 * r1;
        }
        return part;
    };

    for (;;) {
        // =============================================================== phase 0 (groups): refill, guards LS:974-995, Jacobian decision LS:996-1015
        int jacMode = MUX_JAC_NONE;
        bool doEval = false, evalInit = false, finished = false, skipRest = false, accepted = false;
        T fd_xp = (T)0, fd_xm = (T)0, fd_rt = (T)0;
        if (!active && !retired) {
            unsigned int idx = 0, staged = 1;
            if (gl == 0) {
                idx = atomicAdd(args.counter, 1u);
                if (idx < args.batch) staged = wait_staged(args.ready, idx, args.spin_limit) ? 1u : 0u;
            }
            idx = gshfl(gmask, idx, 0); staged = gshfl(gmask, staged, 0);
            if (idx >= args.batch) retired = true;
            else {
                ++sProblems;
                if (!staged) {                     // inputs never arrived: the host discards this launch (flag ready[1]); do not touch x
                    if (gl == 0) {
                        Result bad;
                        bad.status = mir_ls_numericError; bad.iterations = 0; bad.fCalls = 0; bad.gCalls = 0; bad.residual = Num<T>::inf(); bad.lambda = (T)0;
                        static_cast<Result*>(args.results)[idx] = bad;
                    }
                } else {
                    prob = idx;
                    x = valid ? static_cast<const T*>(args.x)[prob * N + gl] : (T)0;
                    lo = valid ? static_cast<const T*>(args.l)[prob * args.bound_stride + gl] : -Num<T>::inf();
                    up = valid ? static_cast<const T*>(args.u)[prob * args.bound_stride + gl] : Num<T>::inf();
                    // validation, LS:930-943 (first failure wins)
                    const bool finite = __all_sync(gmask, !valid || (-Num<T>::inf() < x && x < Num<T>::inf()));
                    const bool inb = __all_sync(gmask, !valid || ((lo <= x) && (x <= up)));
                    int vs = 0;
                    if (m == 0 || !finite) vs = mir_ls_badGuess;
                    else if (!inb) vs = mir_ls_badBounds;
                    else if (!((T)0 <= st.minStepQuality && st.minStepQuality < (T)1)) vs = mir_ls_badMinStepQuality;
                    else if (!((T)0 <= st.goodStepQuality && st.goodStepQuality <= (T)1)) vs = mir_ls_badGoodStepQuality;
                    else if (!(st.minStepQuality < st.goodStepQuality)) vs = mir_ls_badStepQuality;
                    else if (!((T)1 <= st.lambdaIncrease && st.lambdaIncrease <= Num<T>::sqrt_max())) vs = mir_ls_badLambdaParams;
                    else if (!(Num<T>::sqrt_min_normal() <= st.lambdaDecrease && st.lambdaDecrease <= (T)1)) vs = mir_ls_badLambdaParams;
                    if (vs) {
                        if (gl == 0) {
                            Result ret;
                            ret.status = vs; ret.iterations = 0; ret.fCalls = 0; ret.gCalls = 0; ret.residual = Num<T>::inf(); ret.lambda = (T)0;
                            static_cast<Result*>(args.results)[prob] = ret;                  // x is left untouched
                        }
                    } else {
                        active = true; init = true;
                        xt = x; dX = (T)0; Jy = (T)0;
                        maxAge = st.maxAge ? st.maxAge : (FD ? 2u * N : 3u);                                 // LS:945
                        iterations = 0; fCalls = 0; gCalls = 0; status = mir_ls_maxIterations;               // LS:959-971
                        residual = Num<T>::inf(); lambda = (T)0; mu = (T)1; deltaX_dot = (T)0;
                        age = maxAge; needJacobian = false; fConverged = false;
                    }
                }
            }
        }
        if (active) {
            if (init) { doEval = true; evalInit = true; }                                                    // initial residual, LS:953-956
            else {
                ++sPasses;
                if (fConverged) { status = mir_ls_fConverged; finished = true; }                             // LS:974-978
                else if (!(lambda <= st.maxLambda)) { status = mir_ls_furtherImprovement; finished = true; } // LS:979-983
                else {
                    if (mu > (T)16 && age) { needJacobian = true; age = maxAge; mu = (T)1; }                 // LS:984-989
                    if (__any_sync(gmask, valid && !(x <= x))) { status = mir_ls_numericError; finished = true; }   // LS:990-995
                    else {
                        bool inert = false;
                        if (!needJacobian && age == 0 && tailShortcut) {
                            // tail_is_inert (lm_small.cuh), distributed: lane gl holds x_gl, (J^T y)_gl and row gl of J^T J
                            const T q2 = gsum8(gmask, Jy * Jy);
                            const T xmin = gmin8(gmask, valid ? t_abs(x) : Num<T>::inf());
                            if (xmin > (T)0 && st.maxStep > (T)0 && sqrt_ni(q2) < lambda * (xmin * (Num<T>::lapack_eps() * (T)0.125))) {
                                const bool strict = __all_sync(gmask, !valid || ((lo < x) && (x < up)));
                                const bool inside = __all_sync(gmask, !valid || ((lo <= x) && (x <= up)));
                                if (strict) inert = true;
                                else if (inside) {                                                           // tail_bounds_certificate
                                    T row = (T)0;
#pragma unroll
                                    for (int j = 0; j < G; ++j) row += t_abs(sm.JJ[grp][gl * G + j]);
                                    const T nu = gmax8(gmask, row), qinf = gmax8(gmask, t_abs(Jy));
                                    const T thr = ((T)8 * nu) * (qinf / lambda), dmax = xmin * (Num<T>::lapack_eps() * (T)0.25);
                                    const T ql = lo - x, qu = up - x;
                                    const bool onL = ql == (T)0, onU = qu == (T)0;
                                    const bool farL = (-ql - dmax) >= (T)2 * (st.qpSettings.relTolerance + st.qpSettings.absTolerance * t_abs(ql));
                                    const bool farU = (qu - dmax) >= (T)2 * (st.qpSettings.relTolerance + st.qpSettings.absTolerance * t_abs(qu));
                                    bool ok;
                                    if (onL && onU) ok = false;
                                    else if (onL || onU) ok = (onL ? farU : farL) && (t_abs(Jy) >= thr);
                                    else ok = farL && farU;
                                    inert = (lambda >= (T)4 * nu) && __all_sync(gmask, !valid || ok);
                                }
                            }
                        }
                        if (inert) {
                            for (;;) {                         // replay LS:1112, 1125-1130 and the next pass's LS:979-983
                                ++fCalls;
                                lambda *= st.lambdaIncrease * mu; mu *= (T)2;
                                ++sPasses;
                                if (!(lambda <= st.maxLambda)) break;
                            }
                            status = mir_ls_furtherImprovement; finished = true;
                        } else if (needJacobian) {                                                           // LS:996-998
                            needJacobian = false;
                            if (age < maxAge) { ++age; jacMode = MUX_JAC_BROYDEN; ++sBroyden; }              // LS:999-1007
                            else {
                                age = 0; jacMode = MUX_JAC_FRESH; ++sFresh;                                  // LS:1010
                                if (FD) fCalls += N; else gCalls += 1;                                       // LS:1049 (counts tasks) / LS:1014
                                if constexpr (FD) {                                                          // LS:1026-1033, parameter gl
                                    fd_xm = t_max(x - st.jacobianEpsilon, lo);
                                    fd_xp = t_min(x + st.jacobianEpsilon, up);
                                    const T twh = fd_xp - fd_xm;
                                    fd_rt = (twh != (T)0) ? rcp_ni(twh) : (T)0;      // 0 marks "column = 0" (LS:1045-1047); 1 / twh is never 0
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (__all_sync(FULLW, retired)) break;

        // =============================================================== phase 1 (warp): Jacobian step of each slot that needs one, then J^T y, J^T J
        {
            const int flags = (active && !finished) ? jacMode : 0;
#pragma unroll 1
            for (int s = 0; s < S; ++s) {
                const int mode = __shfl_sync(FULLW, flags, s * G);
                if (mode == MUX_JAC_NONE) continue;
                T p[N];
#pragma unroll
                for (int j = 0; j < N; ++j) p[j] = __shfl_sync(FULLW, x, s * G + j);
                const unsigned long long sprob = __shfl_sync(FULLW, prob, s * G);
                T tk[R];
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const int row = lane + 32 * k;
                    tk[k] = gridPerProblem ? ((Model::kHasData && row < m) ? tptr[sprob * (unsigned long long)m + row] : (T)0) : tts[k];
                }
                if (mode == MUX_JAC_FRESH) {
                    if constexpr (FD) {                                                                      // LS:1018-1049
                        T yos[R];
#pragma unroll
                        for (int k = 0; k < R; ++k) yos[k] = slot_get(yo, s, k);
#pragma unroll 1
                        for (int j = 0; j < N; ++j) {
                            const T xp_ = __shfl_sync(FULLW, fd_xp, s * G + j), xm_ = __shfl_sync(FULLW, fd_xm, s * G + j);
                            const T rt_ = __shfl_sync(FULLW, fd_rt, s * G + j);
                            T col[R];
#pragma unroll
                            for (int k = 0; k < R; ++k) col[k] = (T)0;
                            if (rt_ != (T)0) {
                                T fp[R];
#pragma unroll
                                for (int k = 0; k < R; ++k) fp[k] = (T)0;
#pragma unroll 1
                                for (int sgn = 0; sgn < 2; ++sgn) {
                                    T pp[N], f[R];
#pragma unroll
                                    for (int i = 0; i < N; ++i) pp[i] = (i == j) ? (sgn ? xm_ : xp_) : p[i];
                                    eval_rows(pp, tk, yos, f);
#pragma unroll
                                    for (int k = 0; k < R; ++k) {
                                        if (sgn == 0) fp[k] = f[k];
                                        else col[k] = (fp[k] - f[k]) * rt_;                                  // LS:1040-1042
                                    }
                                }
                                if (lane == s * G) sEvals += 2;
                            }
#pragma unroll
                            for (int k = 0; k < R; ++k) sm.J[s][j][lane + 32 * k] = col[k];
                        }
                    } else {                                                                                 // LS:1011-1015
                        const typename Model::Pre pre = Model::prepare(p);
#pragma unroll
                        for (int k = 0; k < R; ++k) {
                            const int row = lane + 32 * k;
                            T Jr[N];
#pragma unroll
                            for (int i = 0; i < N; ++i) Jr[i] = (T)0;
                            if (row < m) Model::jacobian(pre, p, row, tk[k], Jr);
#pragma unroll
                            for (int i = 0; i < N; ++i) sm.J[s][i][row] = Jr[i];
                        }
                    }
                } else {                                                                                     // Broyden, LS:999-1007
                    T dxs[N];
#pragma unroll
                    for (int j = 0; j < N; ++j) dxs[j] = __shfl_sync(FULLW, dX, s * G + j);
                    const T negd = -rcp_ni(__shfl_sync(FULLW, deltaX_dot, s * G));                           // LS:1001
#pragma unroll
                    for (int k = 0; k < R; ++k) {
                        const int row = lane + 32 * k;
                        T Jr[N];
#pragma unroll
                        for (int i = 0; i < N; ++i) Jr[i] = sm.J[s][i][row];
                        T acc = (T)0;                                            // here y = f_new, mBuffer = f_old (after the swap, LS:1136)
#pragma unroll
                        for (int i = 0; i < N; ++i) acc = fma(Jr[i], dxs[i], acc);                           // gemv(1, J, deltaX, 1, mBuffer)
                        const T v = ((slot_get(mb, s, k) - slot_get(yv, s, k)) + acc) * negd;                // axpy(-1, y, mBuffer); scal(-d, mBuffer)
#pragma unroll
                        for (int i = 0; i < N; ++i) sm.J[s][i][row] = fma(v, dxs[i], Jr[i]);                 // ger(1, mBuffer, deltaX, J)
                    }
                }
                T ys[R];
#pragma unroll
                for (int k = 0; k < R; ++k) ys[k] = slot_get(yv, s, k);
                // J^T y (LS:1052) and J^T J (syrk, LS:1065): MUX_RED of the K values at a time -- my rows' partial sums, one
                // padded shared-memory row per value, lane v adds the 32 partials of value v in a fixed order
#pragma unroll
                for (int c0 = 0; c0 < K; c0 += MUX_RED) {
                    T part[MUX_RED];
#pragma unroll
                    for (int e = 0; e < MUX_RED; ++e) part[e] = (T)0;
#pragma unroll
                    for (int k = 0; k < R; ++k) {
                        const int row = lane + 32 * k;
                        T Jr[N];
#pragma unroll
                        for (int i = 0; i < N; ++i) Jr[i] = sm.J[s][i][row];
#pragma unroll
                        for (int e = 0; e < MUX_RED; ++e) {
                            const int v = c0 + e;
                            if (v < NP) { const int i = untri_row(v), j = v - i * (i + 1) / 2; part[e] = fma(Jr[i], Jr[j], part[e]); }
                            else if (v < K) part[e] = fma(Jr[v - NP], ys[k], part[e]);
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int e = 0; e < MUX_RED; ++e) if (c0 + e < K) sm.red[e][lane] = part[e];
                    __syncwarp();
                    if (lane < MUX_RED && c0 + lane < K) {
                        const T* rr = sm.red[lane];
                        T s0 = (T)0, s1 = (T)0, s2 = (T)0, s3 = (T)0;
#pragma unroll
                        for (int i = 0; i < 32; i += 4) { s0 += rr[i]; s1 += rr[i + 1]; s2 += rr[i + 2]; s3 += rr[i + 3]; }
                        const T tot = (s0 + s1) + (s2 + s3);
                        const int v = c0 + lane;
                        if (v < NP) { const int i = untri_row(v), j = v - i * (i + 1) / 2; sm.JJ[s][i * G + j] = tot; sm.JJ[s][j * G + i] = tot; }
                        else sm.Jy[s][v - NP] = tot;
                    }
                }
                __syncwarp();
            }
        }
        __syncwarp();

        // =============================================================== phase 2 (groups): g-test LS:1053-1062, lambda LS:1067-1072, BOXCQP LS:1074-1085, trial point LS:1087-1112
        if (active && !finished && !init) {
            if (jacMode != MUX_JAC_NONE) {
                Jy = valid ? sm.Jy[grp][gl] : (T)0;
                T gsel = gmax8(gmask, t_abs(Jy));                  // iamax picks the first max |.|: its magnitude is the max
                const T j0 = gshfl(gmask, Jy, 0);
                if (!(j0 == j0)) gsel = j0;                        // BLAS: a NaN wins iamax only as the first element
                if (!(gsel > st.gradTolerance)) {
                    if (age == 0) { status = mir_ls_gConverged; finished = true; }
                    else { age = maxAge; skipRest = true; }
                }
            }
            if (!finished && !skipRest) {
                T JJrow[G];
                T JJdiag = (T)0;
#pragma unroll
                for (int j = 0; j < G; ++j) { JJrow[j] = sm.JJ[grp][gl * G + j]; JJdiag = (j == gl) ? JJrow[j] : JJdiag; }
                if (!(lambda >= st.minLambda)) {                                                             // LS:1067-1072
                    const T dmax = gmax8(gmask, valid ? JJdiag : (T)0);   // diag[iamax]; the diagonal of J^T J is >= 0
                    lambda = (T)(0.001 * (double)dmax);
                    if (!(lambda >= st.minLambda)) lambda = (T)1;
                }
                QPCounters qc{0, 0};
                const int qps = boxqp_dist<T, N>(gmask, gl, gshift, st.qpSettings, JJrow, lambda, Jy, lo - x, up - x, dX, qc);   // LS:1074-1080
                sSolves += qc.solves; sQPIt += qc.iterations;
                const bool nan = __any_sync(gmask, valid && !(dX <= dX));                                    // LS:1087-1092
                if (qps != mir_qp_solved || nan) { status = mir_ls_numericError; finished = true; }          // LS:1080-1092
                else {
                    dX = valid ? add_rn(add_rn(dX, x), -x) : (T)0;                                           // LS:1096-1097
                    nd = gsum8(gmask, dX * dX);                                                              // LS:1099
                    if (!(sqrt_ni(nd) < st.maxStep)) { lambda *= st.lambdaIncrease * mu; mu *= (T)2; skipRest = true; }   // LS:1101-1106
                    else {
                        xt = valid ? t_max(t_min(add_rn(dX, x), up), lo) : (T)0;                             // LS:1108-1110
                        const bool same = __all_sync(gmask, !valid || ((xt == x) && (signbit(xt) == signbit(x))));
                        ++fCalls;                                                                            // LS:1112
                        if (same) trial = residual;           // f(xt) == y bit for bit: evaluation skipped, a rejection follows
                        else doEval = true;
                    }
                }
            }
        }
        __syncwarp();

        // =============================================================== phase 3 (warp): f at the trial point (LS:1113-1115) or at x (initial residual, LS:953-955)
        {
            const int flags = (active && !finished && doEval) ? (MUX_EVAL | (evalInit ? MUX_EVAL_INIT : 0)) : 0;
#pragma unroll 1
            for (int s = 0; s < S; ++s) {
                const int f = __shfl_sync(FULLW, flags, s * G);
                if (!(f & MUX_EVAL)) continue;
                const bool isInit = (f & MUX_EVAL_INIT) != 0;
                const T src = isInit ? x : xt;
                T p[N];
#pragma unroll
                for (int j = 0; j < N; ++j) p[j] = __shfl_sync(FULLW, src, s * G + j);
                const unsigned long long sprob = __shfl_sync(FULLW, prob, s * G);
                T tk[R], yos[R], out[R];
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const int row = lane + 32 * k;
                    tk[k] = gridPerProblem ? ((Model::kHasData && row < m) ? tptr[sprob * (unsigned long long)m + row] : (T)0) : tts[k];
                    if (isInit) {
                        const T v = (Model::kHasData && row < m) ? yptr[sprob * (unsigned long long)m + row] : (T)0;
                        slot_put(yo, s, k, v);
                        yos[k] = v;
                    } else yos[k] = slot_get(yo, s, k);
                }
                T part = eval_rows(p, tk, yos, out);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(FULLW, part, off);
#pragma unroll
                for (int k = 0; k < R; ++k) { if (isInit) slot_put(yv, s, k, out[k]); else slot_put(mb, s, k, out[k]); }
                if (grp == s) trial = part;
                if (lane == s * G) ++sEvals;
            }
        }
        __syncwarp();

        // =============================================================== phase 4 (groups): accept / reject, gain ratio, lambda, convergence, LS:1117-1175
        if (active && !finished) {
            if (init) {                                                                                      // LS:953-971
                init = false;
                residual = trial; fCalls = 1;
                fConverged = residual <= st.maxGoodResidual;
                needJacobian = true; age = maxAge;
            } else {
                if (!skipRest) {
                    if (!(trial <= Num<T>::inf())) { status = mir_ls_numericError; finished = true; }        // LS:1117-1122
                    else {
                        const T improvement = residual - trial;                                              // LS:1124
                        if (!(improvement > (T)0)) { lambda *= st.lambdaIncrease * mu; mu *= (T)2; }         // LS:1125-1130
                        else {
                            accepted = true;
                            needJacobian = true; mu = (T)1; ++iterations; ++sAccepted;                       // LS:1132-1139
                            x = xt;
                            residual = trial;
                            fConverged = residual <= st.maxGoodResidual;
                            deltaX_dot = nd;
                            T acc = (T)0;                                                                    // symv(Lower, 1, JJ, deltaX, 2, Jy), LS:1141
#pragma unroll
                            for (int j = 0; j < G; ++j) acc = fma(sm.JJ[grp][gl * G + j], gshfl(gmask, dX, j), acc);
                            Jy = acc + (T)2 * Jy;                      // (scratch from here, as in the reference)
                            const T pred = -gsum8(gmask, Jy * dX);                                           // LS:1142
                            if (!(pred > (T)0)) { status = mir_ls_furtherImprovement; finished = true; }     // LS:1144-1148
                            else {
                                const T rho = div_ni(pred, improvement);                                     // LS:1150
                                if (rho < st.minStepQuality) { lambda *= st.lambdaIncrease * mu; mu *= (T)2; }   // LS:1152-1156
                                else if (rho >= st.goodStepQuality) lambda = t_max(st.lambdaDecrease * lambda * mu, st.minLambda);   // LS:1158-1161
                                const T xmax = gmax8(gmask, valid ? t_abs(x) : (T)0);                        // LS:1164 (nrm2, scaled)
                                T xn = (T)0;
                                if (xmax > (T)0) {
                                    const T vx = valid ? x * rcp_ni(xmax) : (T)0;
                                    xn = xmax * sqrt_ni(gsum8(gmask, vx * vx));
                                }
                                const T sd = sqrt_ni(deltaX_dot);
                                if (!(sd > st.absTolerance && xn > sd * st.relTolerance)) {                  // LS:1164-1173
                                    if (age == 0) { status = mir_ls_xConverged; finished = true; }
                                    else age = maxAge;
                                }
                            }
                        }
                    }
                }
                if (!finished && !(iterations < st.maxIterations)) { status = mir_ls_maxIterations; finished = true; }   // LS:1175
            }
        }
        if (active && finished) {
            if (valid) static_cast<T*>(args.x)[prob * N + gl] = x;
            if (gl == 0) {
                Result ret;
                ret.status = status; ret.iterations = iterations; ret.fCalls = fCalls; ret.gCalls = gCalls;
                ret.residual = residual; ret.lambda = lambda;
                static_cast<Result*>(args.results)[prob] = ret;
            }
            active = false;
        }
        __syncwarp();
        // accepted steps: mBuffer becomes y (the reference swaps the slices, LS:1136) -- on every lane, they all hold rows
#pragma unroll
        for (int s = 0; s < S; ++s) {
            if (__shfl_sync(FULLW, accepted ? 1 : 0, s * G)) {
#pragma unroll
                for (int k = 0; k < R; ++k) { const T tmp = mb[s][k]; mb[s][k] = yv[s][k]; yv[s][k] = tmp; }
            }
        }
    }

    if (args.stats && gl == 0 && sProblems) {
        atomicAdd((unsigned long long*)&args.stats->problems, (unsigned long long)sProblems);
        atomicAdd((unsigned long long*)&args.stats->passes, (unsigned long long)sPasses);
        atomicAdd((unsigned long long*)&args.stats->accepted, (unsigned long long)sAccepted);
        atomicAdd((unsigned long long*)&args.stats->fresh_jacobians, (unsigned long long)sFresh);
        atomicAdd((unsigned long long*)&args.stats->broyden_updates, (unsigned long long)sBroyden);
        atomicAdd((unsigned long long*)&args.stats->model_evals, (unsigned long long)sEvals);
        atomicAdd((unsigned long long*)&args.stats->qp_solves, (unsigned long long)sSolves);
        atomicAdd((unsigned long long*)&args.stats->qp_iterations, (unsigned long long)sQPIt);
    }
}

}  // namespace mirb200
