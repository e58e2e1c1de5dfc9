// lm_tpp.cuh -- batched Levenberg-Marquardt, one THREAD per problem, for large batches of small problems
// (BASELINE configs[1], [2], [4]b).  Follows optimizeLeastSquaresImplGeneric!T, least_squares.d:877-1176.
//
// Why not a warp (or sub-warp group) per problem, as lm_small.cuh does: the n-sized part of a pass (BOXCQP /
// posvx / lambda control, ~55 % of the instructions at n = 4) is serial, so a group of L lanes executes it L
// times redundantly, and different problems in one warp diverge (measured on B200: 18 of 32 lanes active,
// FP64 pipe 20 % busy, 5.9 M fits/s).  With one thread per problem nothing is replicated and no shuffle or
// barrier is needed; the price is that the m-sized per-problem state no longer fits in registers.
//
// Data layout.  Per thread: x, bounds, trial point, step, J^T r, packed lower J^T J, lambda/mu/age and the
// counters live in registers.  The m-sized vectors -- observations, y (current residuals), mBuffer (trial /
// previous residuals) and the m x n Jacobian that the Broyden update needs in place (LS:1003-1006) -- live in a
// per-CTA slab in global memory laid out [element][thread], so the 32 problems of a warp read and write 32
// consecutive values: every access is a fully coalesced 256-byte (double) transaction served from L1/L2.
//
// Lock-step state machine.  All threads of a warp walk the same phase sequence once per LM pass
//     fetch -> guards -> Jacobian -> step (BOXCQP) -> trial evaluation -> accept/reject
// and a phase a problem does not need this pass is predicated off, so the warp reconverges after every phase
// instead of drifting apart; a thread that finishes its problem pulls the next index from an atomic counter
// at the next pass boundary (pass counts vary 10x across a batch).
//
// Equivalences (bit-exact w.r.t. this file's arithmetic; same as lm_small.cuh): J^T J rebuilt only when J
// changed; trial == x skips the model evaluation; the inert lambda-overflow tail is fast-forwarded.
#pragma once
#include "lm_small.cuh"

namespace mirb200 {

constexpr int TPP_THREADS = 128;
enum { JAC_NONE_ = 0, JAC_BROYDEN_ = 1, JAC_FRESH_ = 2 };

#ifndef MIRB200_TPP_MINBLOCKS
#define MIRB200_TPP_MINBLOCKS 1
#endif

// YOS: the observations of the thread's current problem live in shared memory ([row][thread], conflict-free)
// instead of the slab -- they are read by every model evaluation, the most frequent phase.
template <class Model, class T, bool FD, bool YOS>
__global__ void __launch_bounds__(TPP_THREADS, MIRB200_TPP_MINBLOCKS)
lm_tpp_kernel(const typename Num<T>::Settings st, const SmallBatchArgs args, T* __restrict__ slabBase)
{
    constexpr int N = Model::N;
    constexpr int NP = N * (N + 1) / 2;
    constexpr int NT = TPP_THREADS;
    using Result = typename Num<T>::Result;
    const int m = (int)args.m;
    const int tid = threadIdx.x;
    const bool gridPerProblem = (args.flags & MIR_MODEL_GRID_PER_PROBLEM) != 0;
    const bool tailShortcut = (args.flags & MIR_MODEL_NO_TAIL_SHORTCUT) == 0;
    const T* __restrict__ tptr = static_cast<const T*>(args.t);
    const T* __restrict__ yptr = static_cast<const T*>(args.y);

    // abscissa shared by every problem: one copy in shared memory, then (YOS) the observations
    extern __shared__ __align__(16) unsigned char tpp_smem[];
    T* const st_ = reinterpret_cast<T*>(tpp_smem);

    // slab of this CTA: [buf0 m][buf1 m][J m*N] (+ [yobs m] first when !YOS), each element strided by NT
    constexpr int SLAB_VECS = YOS ? 2 : 3;
    T* const slab = slabBase + (size_t)blockIdx.x * ((size_t)m * (N + SLAB_VECS)) * NT + tid;
    T* const pYO = YOS ? st_ + ((m + 1) & ~1) + tid : slab;
    T* const pB0 = slab + (size_t)(SLAB_VECS - 2) * m * NT;
    T* const pB1 = pB0 + (size_t)m * NT;
    T* const pJ = pB1 + (size_t)m * NT;
    auto YO = [&](int row) -> T& { return pYO[row * NT]; };
    auto JE = [&](int row, int i) -> T& { return pJ[(row * N + i) * NT]; };

    if (Model::kHasData && !gridPerProblem) {
        for (int row = tid; row < m; row += NT) st_[row] = tptr[row];
        __syncthreads();
    }

    unsigned long long sPasses = 0, sAccepted = 0, sFresh = 0, sBroyden = 0, sEvals = 0, sSolves = 0, sQPIt = 0, sProblems = 0;

    // per-problem state
    bool active = false, retired = false, init = false;
    unsigned long long prob = 0;
    const T* tp = st_;
    T x[N], lo[N], up[N], xt[N], dX[N], Jy[N], JJ[NP];
    T lambda = (T)0, mu = (T)1, residual = Num<T>::inf(), deltaX_dot = (T)0, nd = (T)0;
    unsigned age = 0, maxAge = 1, iterations = 0, fCalls = 0, gCalls = 0;
    int status = mir_ls_numericError, ysel = 0;
    bool needJacobian = false, fConverged = false;
#pragma unroll
    for (int i = 0; i < N; ++i) { x[i] = lo[i] = up[i] = xt[i] = dX[i] = Jy[i] = (T)0; }
#pragma unroll
    for (int i = 0; i < NP; ++i) JJ[i] = (T)0;

    for (;;) {
        // ------------------------------------------------------------------ fetch
        if (!active && !retired) {
            const unsigned int idx = atomicAdd(args.counter, 1u);
            if (idx >= args.batch) retired = true;
            else {
                prob = idx; ++sProblems;
                const T* xp = static_cast<const T*>(args.x) + prob * N;
                const T* lp = static_cast<const T*>(args.l) + prob * args.bound_stride;
                const T* upp = static_cast<const T*>(args.u) + prob * args.bound_stride;
#pragma unroll
                for (int i = 0; i < N; ++i) { x[i] = xp[i]; lo[i] = lp[i]; up[i] = upp[i]; }
                // validation, LS:930-943 (first failure wins)
                bool finite = true, inb = true;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    finite = finite && (-Num<T>::inf() < x[i] && x[i] < Num<T>::inf());
                    inb = inb && (lo[i] <= x[i]) && (x[i] <= up[i]);
                }
                int vs = 0;
                if (m == 0 || !finite) vs = mir_ls_badGuess;
                else if (!inb) vs = mir_ls_badBounds;
                else if (!((T)0 <= st.minStepQuality && st.minStepQuality < (T)1)) vs = mir_ls_badMinStepQuality;
                else if (!((T)0 <= st.goodStepQuality && st.goodStepQuality <= (T)1)) vs = mir_ls_badGoodStepQuality;
                else if (!(st.minStepQuality < st.goodStepQuality)) vs = mir_ls_badStepQuality;
                else if (!((T)1 <= st.lambdaIncrease && st.lambdaIncrease <= Num<T>::sqrt_max())) vs = mir_ls_badLambdaParams;
                else if (!(Num<T>::sqrt_min_normal() <= st.lambdaDecrease && st.lambdaDecrease <= (T)1)) vs = mir_ls_badLambdaParams;
                if (vs) {
                    Result ret;
                    ret.status = vs; ret.iterations = 0; ret.fCalls = 0; ret.gCalls = 0; ret.residual = Num<T>::inf(); ret.lambda = (T)0;
                    static_cast<Result*>(args.results)[prob] = ret;          // x is left untouched
                } else {
                    active = true; init = true;
                    tp = gridPerProblem ? tptr + prob * (unsigned long long)m : st_;
                    if (Model::kHasData) {
                        const T* yp = yptr + prob * (unsigned long long)m;
#pragma unroll 4
                        for (int row = 0; row < m; ++row) YO(row) = yp[row];
                    }
#pragma unroll
                    for (int i = 0; i < N; ++i) { xt[i] = x[i]; Jy[i] = (T)0; dX[i] = (T)0; }
                    maxAge = st.maxAge ? st.maxAge : (FD ? 2u * N : 3u);                             // LS:945
                    ysel = 0; iterations = 0; fCalls = 0; gCalls = 0; status = mir_ls_maxIterations;
                    residual = Num<T>::inf(); lambda = (T)0; mu = (T)1; deltaX_dot = (T)0;
                }
            }
        }
        if (__all_sync(0xffffffffu, retired)) break;

        // ------------------------------------------------------------------ guards of this pass, LS:974-995
        int jacMode = JAC_NONE_;
        bool doEval = false, skipRest = false, finished = false;
        if (active) {
            if (init) doEval = true;                       // initial residual, LS:953-956
            else {
                ++sPasses;
                if (fConverged) { status = mir_ls_fConverged; finished = true; }                         // LS:974-978
                else if (!(lambda <= st.maxLambda)) { status = mir_ls_furtherImprovement; finished = true; }   // LS:979-983
                else {
                    if (mu > (T)16 && age) { needJacobian = true; age = maxAge; mu = (T)1; }             // LS:984-989
                    bool nan = false;                                                                    // LS:990-995
#pragma unroll
                    for (int i = 0; i < N; ++i) nan = nan || !(x[i] <= x[i]);
                    if (nan) { status = mir_ls_numericError; finished = true; }
                    else if (!needJacobian && age == 0 && tailShortcut && tail_is_inert<T, N>(x, Jy, lambda)) {
                        for (;;) {                         // replay LS:1112, 1125-1130 and the next pass's LS:979-983
                            ++fCalls;
                            lambda *= st.lambdaIncrease * mu; mu *= (T)2;
                            ++sPasses;
                            if (!(lambda <= st.maxLambda)) break;
                        }
                        status = mir_ls_furtherImprovement; finished = true;
                    } else if (needJacobian) {                                                           // LS:996-998
                        needJacobian = false;
                        if (age < maxAge) { ++age; jacMode = JAC_BROYDEN_; ++sBroyden; }                 // LS:999-1007
                        else { age = 0; jacMode = JAC_FRESH_; ++sFresh; if (FD) fCalls += N; else gCalls += 1; }   // LS:1010-1015, 1049
                    }
                }
            }
        }
        const bool go = active && !finished;

        // ------------------------------------------------------------------ Jacobian phase + J^T y, J^T J
        if (go && jacMode != JAC_NONE_) {
            if (jacMode == JAC_FRESH_ && FD) {                                                           // LS:1018-1049
#pragma unroll 1
                for (int j = 0; j < N; ++j) {
                    T save = (T)0, lj = (T)0, uj = (T)0;
#pragma unroll
                    for (int i = 0; i < N; ++i) if (i == j) { save = x[i]; lj = lo[i]; uj = up[i]; }
                    const T xmh = t_max(save - st.jacobianEpsilon, lj);
                    const T xph = t_min(save + st.jacobianEpsilon, uj);
                    const T twh = xph - xmh;
                    if (twh != (T)0) {
                        T pp[N], pm[N];
#pragma unroll
                        for (int i = 0; i < N; ++i) { pp[i] = (i == j) ? xph : x[i]; pm[i] = (i == j) ? xmh : x[i]; }
                        const typename Model::Pre prep = Model::prepare(pp);
                        const typename Model::Pre prem = Model::prepare(pm);
                        const T rt = rcp_ni(twh);
                        int row = 0;
#pragma unroll 1
                        for (; row + 2 <= m; row += 2) {
                            const T t0 = Model::kHasData ? tp[row] : (T)0, y0 = Model::kHasData ? YO(row) : (T)0;
                            const T t1 = Model::kHasData ? tp[row + 1] : (T)0, y1 = Model::kHasData ? YO(row + 1) : (T)0;
                            const T fp0 = Model::residual(prep, pp, row, t0, y0), fm0 = Model::residual(prem, pm, row, t0, y0);
                            const T fp1 = Model::residual(prep, pp, row + 1, t1, y1), fm1 = Model::residual(prem, pm, row + 1, t1, y1);
                            JE(row, j) = (fp0 - fm0) * rt;                                               // LS:1040-1042
                            JE(row + 1, j) = (fp1 - fm1) * rt;
                        }
                        for (; row < m; ++row) {
                            const T tt = Model::kHasData ? tp[row] : (T)0, yo = Model::kHasData ? YO(row) : (T)0;
                            JE(row, j) = (Model::residual(prep, pp, row, tt, yo) - Model::residual(prem, pm, row, tt, yo)) * rt;
                        }
                        sEvals += 2;
                    } else {
#pragma unroll 1
                        for (int row = 0; row < m; ++row) JE(row, j) = (T)0;                             // LS:1045-1047
                    }
                }
            }
            T pJy[N], pJJ[NP];
#pragma unroll
            for (int i = 0; i < N; ++i) pJy[i] = (T)0;
#pragma unroll
            for (int i = 0; i < NP; ++i) pJJ[i] = (T)0;
            const typename Model::Pre pre = Model::prepare(x);
            const T negd = (jacMode == JAC_BROYDEN_) ? -rcp_ni(deltaX_dot) : (T)0;                       // LS:1001
            const T* const yv = ysel ? pB1 : pB0;            // y  = f at the current point
            const T* const fo = ysel ? pB0 : pB1;            // mBuffer = f at the previous point (Broyden)
            auto jrow = [&](int row, T yr, T fold, T tt, T (&Jr)[N]) {
                if (jacMode == JAC_FRESH_ && !FD) {                                                      // LS:1011-1015
                    Model::jacobian(pre, x, row, tt, Jr);
#pragma unroll
                    for (int i = 0; i < N; ++i) JE(row, i) = Jr[i];
                } else if (jacMode == JAC_BROYDEN_) {                                                    // LS:1003-1006
                    T acc = (T)0;
#pragma unroll
                    for (int i = 0; i < N; ++i) acc += Jr[i] * dX[i];
                    const T v = ((fold - yr) + acc) * negd;
#pragma unroll
                    for (int i = 0; i < N; ++i) { Jr[i] += v * dX[i]; JE(row, i) = Jr[i]; }
                }
            };
            auto accum = [&](T yr, const T (&Jr)[N]) {                                                   // LS:1052, 1065
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    pJy[i] += Jr[i] * yr;
#pragma unroll
                    for (int j = 0; j <= i; ++j) pJJ[tri(i, j)] += Jr[i] * Jr[j];
                }
            };
            const bool loadJ = !(jacMode == JAC_FRESH_ && !FD);
            const bool isBro = jacMode == JAC_BROYDEN_;
            int row = 0;
#pragma unroll 1
            for (; row + 2 <= m; row += 2) {
                T Ja[N], Jb[N];
                const T ya = yv[row * NT], yb = yv[(row + 1) * NT];
                const T fa = isBro ? fo[row * NT] : (T)0, fb = isBro ? fo[(row + 1) * NT] : (T)0;
                const T ta = (Model::kHasData && !loadJ) ? tp[row] : (T)0, tb = (Model::kHasData && !loadJ) ? tp[row + 1] : (T)0;
#pragma unroll
                for (int i = 0; i < N; ++i) { Ja[i] = loadJ ? JE(row, i) : (T)0; Jb[i] = loadJ ? JE(row + 1, i) : (T)0; }
                jrow(row, ya, fa, ta, Ja);
                jrow(row + 1, yb, fb, tb, Jb);
                accum(ya, Ja);
                accum(yb, Jb);
            }
            for (; row < m; ++row) {
                T Ja[N];
                const T ya = yv[row * NT];
                const T fa = isBro ? fo[row * NT] : (T)0;
                const T ta = (Model::kHasData && !loadJ) ? tp[row] : (T)0;
#pragma unroll
                for (int i = 0; i < N; ++i) Ja[i] = loadJ ? JE(row, i) : (T)0;
                jrow(row, ya, fa, ta, Ja);
                accum(ya, Ja);
            }
#pragma unroll
            for (int i = 0; i < N; ++i) Jy[i] = pJy[i];
#pragma unroll
            for (int i = 0; i < NP; ++i) JJ[i] = pJJ[i];

            T gsel = Jy[0]; T gbest = t_abs(Jy[0]);                                                      // LS:1053 (iamax: first max |.|)
#pragma unroll
            for (int i = 1; i < N; ++i) { const T v = t_abs(Jy[i]); if (v > gbest) { gbest = v; gsel = Jy[i]; } }
            if (!(t_abs(gsel) > st.gradTolerance)) {                                                     // LS:1053-1062
                if (age == 0) { status = mir_ls_gConverged; finished = true; }
                else { age = maxAge; skipRest = true; }
            }
        }

        // ------------------------------------------------------------------ step: lambda init, BOXCQP, trial point
        if (active && !finished && !init && !skipRest) {
            if (!(lambda >= st.minLambda)) {                                                             // LS:1067-1072
                T dmax = JJ[0];
#pragma unroll
                for (int i = 1; i < N; ++i) if (t_abs(JJ[tri(i, i)]) > t_abs(dmax)) dmax = JJ[tri(i, i)];
                lambda = (T)(0.001 * (double)dmax);
                if (!(lambda >= st.minLambda)) lambda = (T)1;
            }
            T qpl[N], qpu[N];                                                                            // LS:1074-1077
#pragma unroll
            for (int i = 0; i < N; ++i) { qpl[i] = lo[i] - x[i]; qpu[i] = up[i] - x[i]; }
            QPCounters qc{0, 0};
            const int qps = boxqp_small<T, N>(st.qpSettings, JJ, lambda, Jy, qpl, qpu, dX, qc);          // LS:1078-1080
            sSolves += qc.solves; sQPIt += qc.iterations;
            bool nan = false;                                                                            // LS:1087-1092
#pragma unroll
            for (int i = 0; i < N; ++i) nan = nan || !(dX[i] <= dX[i]);
            if (qps != mir_qp_solved || nan) { status = mir_ls_numericError; finished = true; }          // LS:1080-1092
            else {
                nd = (T)0;                                                                               // LS:1096-1099
#pragma unroll
                for (int i = 0; i < N; ++i) { dX[i] = add_rn(add_rn(dX[i], x[i]), -x[i]); nd += dX[i] * dX[i]; }
                if (!(sqrt_ni(nd) < st.maxStep)) { lambda *= st.lambdaIncrease * mu; mu *= (T)2; skipRest = true; }   // LS:1101-1106
                else {
                    bool same = true;                                                                    // LS:1108-1110
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        xt[i] = t_max(t_min(add_rn(dX[i], x[i]), up[i]), lo[i]);
                        same = same && (xt[i] == x[i]) && (signbit(xt[i]) == signbit(x[i]));
                    }
                    ++fCalls;                                                                            // LS:1112
                    doEval = !same;        // f(xt) == y bit for bit when xt == x: evaluation skipped, trial = residual
                }
            }
        }

        // ------------------------------------------------------------------ trial evaluation, LS:1113-1115 (and LS:953-955)
        T trial = residual;
        if (active && !finished && doEval) {
            const typename Model::Pre pre = Model::prepare(xt);
            T acc = (T)0;
            T* const out = ysel ? pB0 : pB1;                   // mBuffer = the buffer that is not y
            constexpr int U = 4;                               // independent rows in flight: their exp chains interleave
            int row = 0;
#pragma unroll 1
            for (; row + U <= m; row += U) {
                T tt[U], yo[U], r[U];
#pragma unroll
                for (int u = 0; u < U; ++u) { tt[u] = Model::kHasData ? tp[row + u] : (T)0; yo[u] = Model::kHasData ? YO(row + u) : (T)0; }
#pragma unroll
                for (int u = 0; u < U; ++u) r[u] = Model::residual(pre, xt, row + u, tt[u], yo[u]);
#pragma unroll
                for (int u = 0; u < U; ++u) { out[(row + u) * NT] = r[u]; acc += r[u] * r[u]; }
            }
            for (; row < m; ++row) {
                const T r = Model::residual(pre, xt, row, Model::kHasData ? tp[row] : (T)0, Model::kHasData ? YO(row) : (T)0);
                out[row * NT] = r;
                acc += r * r;
            }
            trial = acc;
            ++sEvals;
        }

        // ------------------------------------------------------------------ accept / reject, LS:1117-1175
        if (active && !finished) {
            const bool wasInit = init;
            if (init) {                                                                                  // LS:953-971
                init = false;
                residual = trial; ysel ^= 1; fCalls = 1;
                fConverged = residual <= st.maxGoodResidual;
                needJacobian = true; age = maxAge;
            } else if (!skipRest) {
                if (!(trial <= Num<T>::inf())) { status = mir_ls_numericError; finished = true; }        // LS:1117-1122
                else {
                    const T improvement = residual - trial;                                              // LS:1124
                    if (!(improvement > (T)0)) { lambda *= st.lambdaIncrease * mu; mu *= (T)2; }         // LS:1125-1130
                    else {
                        needJacobian = true; mu = (T)1; ++iterations; ++sAccepted;                       // LS:1132-1139
#pragma unroll
                        for (int i = 0; i < N; ++i) x[i] = xt[i];
                        ysel ^= 1;
                        residual = trial;
                        fConverged = residual <= st.maxGoodResidual;
                        deltaX_dot = nd;
                        T pred = (T)0;                                                                   // LS:1141-1142
#pragma unroll
                        for (int i = 0; i < N; ++i) {
                            T acc = (T)0;
#pragma unroll
                            for (int j = 0; j < N; ++j) acc += JJ[trisym(i, j)] * dX[j];
                            Jy[i] = acc + (T)2 * Jy[i];
                            pred += Jy[i] * dX[i];
                        }
                        pred = -pred;
                        if (!(pred > (T)0)) { status = mir_ls_furtherImprovement; finished = true; }     // LS:1144-1148
                        else {
                            const T rho = div_ni(pred, improvement);                                     // LS:1150
                            if (rho < st.minStepQuality) { lambda *= st.lambdaIncrease * mu; mu *= (T)2; }   // LS:1152-1156
                            else if (rho >= st.goodStepQuality) lambda = t_max(st.lambdaDecrease * lambda * mu, st.minLambda);   // LS:1158-1161
                            T xmax = (T)0;                                                               // LS:1164 (nrm2, scaled)
#pragma unroll
                            for (int i = 0; i < N; ++i) xmax = t_max(xmax, t_abs(x[i]));
                            T xn = (T)0;
                            if (xmax > (T)0) {
                                const T inv = rcp_ni(xmax);
                                T ss = (T)0;
#pragma unroll
                                for (int i = 0; i < N; ++i) { const T v = x[i] * inv; ss += v * v; }
                                xn = xmax * sqrt_ni(ss);
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
            // LS:1175 (a do-while: the first pass always runs)
            if (!wasInit && !finished && !(iterations < st.maxIterations)) { status = mir_ls_maxIterations; finished = true; }
        }

        if (active && finished) {
            T* xp = static_cast<T*>(args.x) + prob * N;
#pragma unroll
            for (int i = 0; i < N; ++i) xp[i] = x[i];
            Result ret;
            ret.status = status; ret.iterations = iterations; ret.fCalls = fCalls; ret.gCalls = gCalls;
            ret.residual = residual; ret.lambda = lambda;
            static_cast<Result*>(args.results)[prob] = ret;
            active = false;
        }
    }

    if (args.stats) {
        auto wsum = [](unsigned long long v) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            return v;
        };
        sProblems = wsum(sProblems); sPasses = wsum(sPasses); sAccepted = wsum(sAccepted); sFresh = wsum(sFresh);
        sBroyden = wsum(sBroyden); sEvals = wsum(sEvals); sSolves = wsum(sSolves); sQPIt = wsum(sQPIt);
        if ((tid & 31) == 0 && sProblems) {
            atomicAdd((unsigned long long*)&args.stats->problems, sProblems);
            atomicAdd((unsigned long long*)&args.stats->passes, sPasses);
            atomicAdd((unsigned long long*)&args.stats->accepted, sAccepted);
            atomicAdd((unsigned long long*)&args.stats->fresh_jacobians, sFresh);
            atomicAdd((unsigned long long*)&args.stats->broyden_updates, sBroyden);
            atomicAdd((unsigned long long*)&args.stats->model_evals, sEvals);
            atomicAdd((unsigned long long*)&args.stats->qp_solves, sSolves);
            atomicAdd((unsigned long long*)&args.stats->qp_iterations, sQPIt);
        }
    }
}

}  // namespace mirb200
