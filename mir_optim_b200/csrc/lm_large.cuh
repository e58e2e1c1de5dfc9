// lm_large.cuh -- device side of the single-large-problem Levenberg-Marquardt engine
// (m >> n, n <= 128; BASELINE configs[0] and configs[3]).  Follows
// optimizeLeastSquaresImplGeneric!T, least_squares.d:877-1176 (line cites inline).
//
// One LM pass (one trip round the do-while at LS:972-1175) is a fixed sequence of kernels that all
// read a device-resident control block (LargeCtl) and turn themselves into no-ops when the pass
// does not need them, so the host enqueues passes blindly -- there is no per-iteration host
// round trip; the host only polls `done` every few passes:
//
//   jac / fd-jac / broyden   (row-parallel; fresh J or LS:1001-1006 rank-1 update, fused J^T y)
//   syrk_dmma + reduce       (J^T J on the FP64 tensor pipe, syrk_dmma.cuh)
//   [all-reduce packed]      (row-sharded runs: n(n+1)/2 + n doubles over NCCL)
//   ctl_mid                  (1 CTA: g-test, lambda init, BOXCQP step, trial point)
//   eval                     (row-parallel residuals at the trial point + ||r||^2)
//   [all-reduce rr]          (1 double)
//   ctl_post                 (1 CTA: accept/reject, gain ratio, lambda update, convergence,
//                             loop guards of the NEXT pass LS:974-995 and its Jacobian mode)
//
// Every rank of a sharded run executes ctl_mid / ctl_post redundantly on bit-identical inputs
// (NCCL all-reduce hands every rank the same bits), so all ranks take the same branches.
//
// J is materialised (the Broyden update needs it in place, LS:1003-1006): row-major, row pitch
// ldj = n rounded up to even, rows padded with zeros to a multiple of SYRK_KT.  y and mBuffer are
// two m-vectors selected by ctl->ysel (the reference swaps the slices, LS:1136).
#pragma once
#include "boxqp_cta.cuh"
#include "models_large.cuh"
#include "syrk_dmma.cuh"

namespace mirb200 {

constexpr int LARGE_NMAX = 128;
constexpr int LARGE_NP_MAX = LARGE_NMAX * (LARGE_NMAX + 1) / 2;
constexpr int LARGE_TILE = 32;       // rows per Jacobian tile
constexpr int LARGE_CTL_THREADS = 256;      // >= 136 = lower-triangular 8 x 8 tiles of a 128 x 128 system (cta_ldl_factor_blocked)

enum { JAC_NONE = 0, JAC_BROYDEN = 1, JAC_FRESH = 2 };

template <class T> struct LargeCtl {
    typename Num<T>::Settings st;
    int n, ldj;
    unsigned maxAge;
    int hasG;                    // analytic Jacobian supplied (g != null)
    int tailShortcut;            // fast-forward the inert lambda-overflow tail (lm_small.cuh, tail_is_inert)
    // loop state, LS:959-971
    T lambda, mu, residual, deltaX_dot, nd;
    unsigned age, iterations, fCalls, gCalls;
    int status, needJacobian, fConverged;
    // per-pass flags
    int jacMode, doEval, skipRest, done, ysel, initPhase;
    unsigned int ticket[4];      // last-block tickets of the row-parallel kernels
    int chunkSeq, doneChunk;     // chunks of passes completed so far; index of the chunk in which `done` was raised (-1: still running)
    unsigned long long passes, accepted, fresh, broyden, evals, qpSolves, qpIters;
    T rr;                        // ||f(trial)||^2 (all-reduced in place)
    T x[LARGE_NMAX], xt[LARGE_NMAX], l[LARGE_NMAX], u[LARGE_NMAX], dX[LARGE_NMAX], Jy[LARGE_NMAX];
    T JJ[LARGE_NMAX * LARGE_NMAX];                   // PACKED lower triangle (tri(i,j)), undamped: the copy of `packed` taken right after
                                                     // its all-reduce (the host all-reduces `packed` on every pass, also when J did not change)
    T packed[LARGE_NP_MAX + LARGE_NMAX];             // [lower(J^T J) by rows | J^T y]  (all-reduced in place)
};

// ---------------------------------------------------------------------------------------------
// small deterministic CTA reductions
// ---------------------------------------------------------------------------------------------
template <class T, int NT> __device__ __forceinline__ T cta_sum_fixed(T v, T* red)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    T r = red[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) r += red[w];
    return r;
}

// The scalar part of the control block (everything before x) as 32-bit words: the one-CTA control kernels copy it into
// shared memory on entry, work on the copy and write it back on exit.
template <class T> __host__ __device__ constexpr int large_head_words() { return (int)(offsetof(LargeCtl<T>, x) / 4); }
template <class T> __device__ __forceinline__ LargeCtl<T>* large_head_load(const LargeCtl<T>* c, unsigned int* buf)
{
    static_assert(offsetof(LargeCtl<T>, x) % 4 == 0, "header size");
    const unsigned int* src = reinterpret_cast<const unsigned int*>(c);
    for (int i = threadIdx.x; i < large_head_words<T>(); i += blockDim.x) buf[i] = __ldcg(src + i);
    return reinterpret_cast<LargeCtl<T>*>(buf);       // ONLY the scalar fields may be touched through this pointer
}
template <class T> __device__ __forceinline__ void large_head_store(LargeCtl<T>* c, const unsigned int* buf)
{
    unsigned int* dst = reinterpret_cast<unsigned int*>(c);
    for (int i = threadIdx.x; i < large_head_words<T>(); i += blockDim.x) dst[i] = buf[i];
}

// LS:974-995 guards + choice of the Jacobian work for the pass that follows (LS:996-1015 bookkeeping).
// Runs in ONE thread.  c is the caller's SHARED-MEMORY copy of the control block's scalar header (large_head_load) and
// x, Jy, l, u its shared copies of the vectors; JJ is the global packed matrix (rare certificate branch only).  The serial
// walk over the global control block this replaces cost ~15 us per pass: read-modify-write sequences on one cache line,
// every read an L2 round trip -- the Amdahl term of row-sharded runs.
template <class T> __device__ void large_begin_pass(LargeCtl<T>* c, const T* x, const T* Jy, const T* l, const T* u, const T* JJ)
{
    // single thread
    c->jacMode = JAC_NONE; c->doEval = 0; c->skipRest = 0;
    if (c->done) return;
    ++c->passes;
    if (c->fConverged) { c->status = mir_ls_fConverged; c->done = 1; return; }                     // LS:974-978
    if (!(c->lambda <= c->st.maxLambda)) { c->status = mir_ls_furtherImprovement; c->done = 1; return; }   // LS:979-983
    if (c->mu > (T)16 && c->age) { c->needJacobian = 1; c->age = c->maxAge; c->mu = (T)1; }        // LS:984-989
    bool nan = false;                                                                              // LS:990-995
    for (int i = 0; i < c->n; ++i) nan = nan || !(x[i] <= x[i]);
    if (nan) { c->status = mir_ls_numericError; c->done = 1; return; }
    if (!c->needJacobian && c->age == 0 && c->tailShortcut) {
        // inert lambda-overflow tail (proof at tail_is_inert in lm_small.cuh): replay the scalar recurrence only
        T q2 = (T)0, xmin = Num<T>::inf();
        for (int i = 0; i < c->n; ++i) { q2 += Jy[i] * Jy[i]; xmin = t_min(xmin, t_abs(x[i])); }
        // + maxStep > 0, and BOXCQP must provably return `solved`: no x_i on a bound, or the on-bound certificate
        // (tail_bounds_certificate in lm_small.cuh, same conditions, restated here for run-time n)
        bool ok = c->st.maxStep > (T)0 && xmin > (T)0 && sqrt_ni(q2) < c->lambda * (xmin * (Num<T>::lapack_eps() * (T)0.125));
        bool any = false;
        for (int i = 0; ok && i < c->n; ++i) {
            ok = (l[i] <= x[i]) && (x[i] <= u[i]);
            any = any || (l[i] == x[i]) || (u[i] == x[i]);
        }
        if (ok && any) {
            const int n = c->n;
            T nu = (T)0, qinf = (T)0;
            for (int i = 0; i < n; ++i) {
                T row = (T)0;
                for (int j = 0; j < n; ++j) row += t_abs(JJ[trisym(i, j)]);
                nu = t_max(nu, row); qinf = t_max(qinf, t_abs(Jy[i]));
            }
            ok = c->lambda >= (T)4 * nu;
            const T thr = ((T)8 * nu) * (qinf / c->lambda), dmax = xmin * (Num<T>::lapack_eps() * (T)0.25);
            const T relTol = c->st.qpSettings.relTolerance, absTol = c->st.qpSettings.absTolerance;
            for (int i = 0; ok && i < n; ++i) {
                const T ql = l[i] - x[i], qu = u[i] - x[i];
                const bool onL = ql == (T)0, onU = qu == (T)0;
                const bool farL = (-ql - dmax) >= (T)2 * (relTol + absTol * t_abs(ql));
                const bool farU = (qu - dmax) >= (T)2 * (relTol + absTol * t_abs(qu));
                if (onL && onU) ok = false;
                else if (onL || onU) ok = (onL ? farU : farL) && (t_abs(Jy[i]) >= thr);
                else ok = farL && farU;
            }
        }
        if (ok) {
            for (;;) {
                ++c->fCalls;
                c->lambda *= c->st.lambdaIncrease * c->mu; c->mu *= (T)2;
                ++c->passes;
                if (!(c->lambda <= c->st.maxLambda)) break;
            }
            c->status = mir_ls_furtherImprovement; c->done = 1; return;
        }
    }
    if (c->needJacobian) {                                                                         // LS:996-998
        c->needJacobian = 0;
        if (c->age < c->maxAge) { ++c->age; c->jacMode = JAC_BROYDEN; ++c->broyden; }              // LS:999-1007
        else {
            c->age = 0; c->jacMode = JAC_FRESH; ++c->fresh;                                        // LS:1010
            if (c->hasG) c->gCalls += 1;                                                           // LS:1014
            else c->fCalls += (unsigned)c->n;                                                      // LS:1049
        }
    }
}

// Last node of every chunk of passes.  The host enqueues chunks blindly and polls one chunk behind; what it polls is
// doneChunk, the index of the chunk in which the solve finished, never a bare flag: a rank of a row-sharded run whose
// host happens to read the mailbox late (after the NEXT chunk's copy landed) must still take the decision that belongs
// to the chunk it synchronised on, or ranks would enqueue different numbers of chunks -- and of all-reduces.
template <class T> __global__ void large_chunk_end_kernel(LargeCtl<T>* c)
{
    if (c->done && c->doneChunk < 0) c->doneChunk = c->chunkSeq;
    ++c->chunkSeq;
}

// ---------------------------------------------------------------------------------------------
// residual evaluation: out = f(xt) for this rank's rows, rr = sum of squares  (LS:953-955, 1113-1115)
// ---------------------------------------------------------------------------------------------
template <class T> struct EvalArgs {
    LargeCtl<T>* ctl;
    const T* t; const T* yobs;
    T* buf0; T* buf1;           // y / mBuffer pair
    T* partRR;                  // gridDim.x partial sums
    long long rows;
};

template <class Model, class T>
__global__ void __launch_bounds__(256) large_eval_kernel(const EvalArgs<T> a)
{
    LargeCtl<T>* c = a.ctl;
    if (c->done || !c->doEval) return;
    __shared__ T sp[LARGE_NMAX], saux[LARGE_NMAX], red[8];
    __shared__ bool last;
    const int n = c->n, tid = threadIdx.x;
    for (int k = tid; k < n; k += 256) { const T v = c->xt[k]; sp[k] = v; saux[k] = Model::aux_of(k, n, v); }
    __syncthreads();
    T* out = c->ysel ? a.buf0 : a.buf1;        // the trial residuals go to mBuffer = the buffer that is not y
    ParamView<T> pv{sp, saux, -1, (T)0, (T)0};
    T acc = (T)0;
    for (long long row = (long long)blockIdx.x * 256 + tid; row < a.rows; row += (long long)gridDim.x * 256) {
        const T r = Model::residual(pv, n, a.t[row], a.yobs[row]);
        out[row] = r;
        acc += r * r;
    }
    const T s = cta_sum_fixed<T, 256>(acc, red);
    if (tid == 0) {
        a.partRR[blockIdx.x] = s;
        __threadfence();
        last = atomicAdd(&c->ticket[0], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        T v = (T)0;
        for (int i = tid; i < (int)gridDim.x; i += 256) v += a.partRR[i];
        const T tot = cta_sum_fixed<T, 256>(v, red);
        if (tid == 0) { c->rr = tot; c->ticket[0] = 0; ++c->evals; }
    }
}

// rr of a residual vector that is already in the mBuffer (host-callback mode)
template <class T>
__global__ void __launch_bounds__(256) large_rr_kernel(const EvalArgs<T> a)
{
    LargeCtl<T>* c = a.ctl;
    if (c->done || !c->doEval) return;
    __shared__ T red[8];
    __shared__ bool last;
    const int tid = threadIdx.x;
    const T* in = c->ysel ? a.buf0 : a.buf1;
    T acc = (T)0;
    for (long long row = (long long)blockIdx.x * 256 + tid; row < a.rows; row += (long long)gridDim.x * 256) { const T r = in[row]; acc += r * r; }
    const T s = cta_sum_fixed<T, 256>(acc, red);
    if (tid == 0) { a.partRR[blockIdx.x] = s; __threadfence(); last = atomicAdd(&c->ticket[0], 1u) == gridDim.x - 1; }
    __syncthreads();
    if (last) {
        __threadfence();
        T v = (T)0;
        for (int i = tid; i < (int)gridDim.x; i += 256) v += a.partRR[i];
        const T tot = cta_sum_fixed<T, 256>(v, red);
        if (tid == 0) { c->rr = tot; c->ticket[0] = 0; ++c->evals; }
    }
}

// ---------------------------------------------------------------------------------------------
// Jacobian kernels.  All of them also produce J^T y (LS:1052) for the rows they touch: per-CTA
// partial sums, added in CTA order by the last CTA to finish -> packed[np .. np+n).
// ---------------------------------------------------------------------------------------------
template <class T> struct JacArgs {
    LargeCtl<T>* ctl;
    const T* t; const T* yobs;
    const T* buf0; const T* buf1;
    T* J;
    T* partJy;                  // gridDim.x * LARGE_NMAX
    long long rows;
};

template <class T, int NT>
__device__ __forceinline__ void large_finish_jy(LargeCtl<T>* c, T* partJy, const T* myJy /* smem, ldj */, int ldj, unsigned* ticket)
{
    __shared__ bool lastJ;
    const int tid = threadIdx.x;
    for (int k = tid; k < ldj; k += NT) partJy[(size_t)blockIdx.x * LARGE_NMAX + k] = myJy[k];
    __threadfence();
    __syncthreads();
    if (tid == 0) lastJ = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (lastJ) {
        __threadfence();
        const int n = c->n, np = n * (n + 1) / 2;
        for (int k = tid; k < n; k += NT) {
            T s0 = (T)0, s1 = (T)0;
            int b = 0;
            for (; b + 2 <= (int)gridDim.x; b += 2) { s0 += partJy[(size_t)b * LARGE_NMAX + k]; s1 += partJy[(size_t)(b + 1) * LARGE_NMAX + k]; }
            if (b < (int)gridDim.x) s0 += partJy[(size_t)b * LARGE_NMAX + k];
            c->packed[np + k] = s0 + s1;
        }
        if (tid == 0) *ticket = 0;
    }
}

// fresh Jacobian: analytic (LS:1011-1015) or central differences (LS:1018-1049; FD = true)
template <class Model, class T, bool FD>
__global__ void __launch_bounds__(256) large_jac_kernel(const JacArgs<T> a)
{
    LargeCtl<T>* c = a.ctl;
    if (c->done || c->jacMode != JAC_FRESH) return;
    constexpr int NT = 256;
    __shared__ T sp[LARGE_NMAX], saux[LARGE_NMAX], sjy[LARGE_NMAX];
    __shared__ T sxm[FD ? LARGE_NMAX : 1], sxp[FD ? LARGE_NMAX : 1], srt[FD ? LARGE_NMAX : 1], sam[FD ? LARGE_NMAX : 1], sap[FD ? LARGE_NMAX : 1];
    __shared__ T st_[LARGE_TILE], sy_[LARGE_TILE], syo_[LARGE_TILE];
    extern __shared__ __align__(16) unsigned char jac_smem[];
    T* tile = reinterpret_cast<T*>(jac_smem);                  // LARGE_TILE x (ldj + 1)
    const int n = c->n, ldj = c->ldj, tp = ldj + 1, tid = threadIdx.x;
    const T* y = c->ysel ? a.buf1 : a.buf0;

    for (int k = tid; k < n; k += NT) {
        const T v = c->x[k]; sp[k] = v; saux[k] = Model::aux_of(k, n, v);
        if (FD) {                                                                                  // LS:1026-1033
            T xmh = v - c->st.jacobianEpsilon, xph = v + c->st.jacobianEpsilon;
            xmh = t_max(xmh, c->l[k]); xph = t_min(xph, c->u[k]);
            const T twh = xph - xmh;
            sxm[k] = xmh; sxp[k] = xph; srt[k] = (twh != (T)0) ? rcp_ni(twh) : (T)0;               // rt == 0 marks "column = 0" (LS:1045-1047)
            sam[k] = Model::aux_of(k, n, xmh); sap[k] = Model::aux_of(k, n, xph);
        }
    }
    for (int k = tid; k < LARGE_NMAX; k += NT) sjy[k] = (T)0;
    for (int e = tid; e < LARGE_TILE * tp; e += NT) tile[e] = (T)0;          // pad column (ldj > n) stays 0
    __syncthreads();

    const long long tiles = (a.rows + LARGE_TILE - 1) / LARGE_TILE;
    const long long per = (tiles + gridDim.x - 1) / gridDim.x;
    const long long t0 = (long long)blockIdx.x * per, t1 = (t0 + per < tiles) ? t0 + per : tiles;
    const int items = FD ? n : Model::jac_items(n);
    T myJy = (T)0;                                             // thread k < ldj owns column k

    for (long long tl = t0; tl < t1; ++tl) {
        const long long row0 = tl * LARGE_TILE;
        if (tid < LARGE_TILE) {
            const long long row = row0 + tid;
            const bool ok = row < a.rows;
            st_[tid] = ok ? a.t[row] : (T)0;
            sy_[tid] = ok ? y[row] : (T)0;
            if (FD) syo_[tid] = ok ? a.yobs[row] : (T)0;
        }
        __syncthreads();
        // lane = row of the tile, warp = item (stride NT / 32): the item -- and with it every branch and parameter load
        // of jac_item -- is uniform over a warp, the shared-memory stores of a warp walk one column of the tile (pitch
        // ldj + 1: conflict-free), and no division by the run-time item count is needed
        static_assert(LARGE_TILE == 32, "lane = tile row");
        for (int item = tid >> 5; item < items; item += NT / 32) {
            const int r = tid & 31;
            if (row0 + r >= a.rows) continue;
            if (!FD) {
                ParamView<T> pv{sp, saux, -1, (T)0, (T)0};
                Model::jac_item(pv, n, item, st_[r], tile + r * tp);
            } else {
                T v = (T)0;
                if (srt[item] != (T)0) {
                    ParamView<T> pp{sp, saux, item, sxp[item], sap[item]};
                    ParamView<T> pm{sp, saux, item, sxm[item], sam[item]};
                    const T fp = Model::residual(pp, n, st_[r], syo_[r]);
                    const T fm = Model::residual(pm, n, st_[r], syo_[r]);
                    v = (fp - fm) * srt[item];                                                     // LS:1040-1042
                }
                tile[r * tp + item] = v;
            }
        }
        __syncthreads();
        if (tid < ldj) {
            T acc = myJy;
#pragma unroll 8
            for (int r = 0; r < LARGE_TILE; ++r) acc += tile[r * tp + tid] * sy_[r];
            myJy = acc;
        }
        {   // write-out: the tile is one contiguous block of J; (r, k) advance incrementally instead of a division per element
            int r = tid / ldj, k = tid - r * ldj;
            const int dr = NT / ldj, dk = NT - dr * ldj;
            T* const Jt = a.J + (size_t)row0 * ldj;
            for (int e = tid; e < LARGE_TILE * ldj; e += NT) {
                if (row0 + r < a.rows) Jt[e] = tile[r * tp + k];
                r += dr; k += dk;
                if (k >= ldj) { k -= ldj; ++r; }
            }
        }
        __syncthreads();
    }
    if (tid < ldj) sjy[tid] = myJy;
    __syncthreads();
    large_finish_jy<T, NT>(c, a.partJy, sjy, ldj, &c->ticket[1]);
}

// Broyden rank-1 update (LS:1001-1006), one warp per row, fused with J^T y.  UPDATE = false only
// forms J^T y of a Jacobian that is already in memory (host-callback mode after the user's g).
//   mBuffer = f_old, y = f_new:   v = ((f_old - f_new) + J dX) * (-1/||dX||^2);   J += v dX^T
template <class T, bool UPDATE>
__global__ void __launch_bounds__(256) large_broyden_kernel(const JacArgs<T> a)
{
    LargeCtl<T>* c = a.ctl;
    if (c->done || c->jacMode != (UPDATE ? JAC_BROYDEN : JAC_FRESH)) return;
    constexpr int NT = 256, NW = NT / 32, Q = LARGE_NMAX / 32;
    __shared__ T sw[NW][LARGE_NMAX], sjy[LARGE_NMAX];
    const int ldj = c->ldj, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const T* y = c->ysel ? a.buf1 : a.buf0;
    const T* mb = c->ysel ? a.buf0 : a.buf1;
    T dx[Q], jy[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) { const int k = lane + 32 * q; dx[q] = (UPDATE && k < c->n) ? c->dX[k] : (T)0; jy[q] = (T)0; }
    const T negd = UPDATE ? -((T)1 / c->deltaX_dot) : (T)0;                                        // LS:1001

    const long long per = (a.rows + gridDim.x - 1) / gridDim.x;
    const long long r0 = (long long)blockIdx.x * per, r1 = (r0 + per < a.rows) ? r0 + per : a.rows;
    for (long long row = r0 + warp; row < r1; row += NW) {
        T* Jr = a.J + (size_t)row * ldj;
        T jv[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) { const int k = lane + 32 * q; jv[q] = (k < ldj) ? Jr[k] : (T)0; }
        const T yr = y[row];
        if (UPDATE) {
            T acc = (T)0;
#pragma unroll
            for (int q = 0; q < Q; ++q) acc += jv[q] * dx[q];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            const T v = ((mb[row] - yr) + acc) * negd;                                             // LS:1003-1005
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int k = lane + 32 * q;
                jv[q] += v * dx[q];                                                                // LS:1006
                if (k < ldj) Jr[k] = jv[q];
            }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) jy[q] += jv[q] * yr;
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) sw[warp][lane + 32 * q] = jy[q];
    __syncthreads();
    if (tid < LARGE_NMAX) {
        T s = (T)0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += sw[w][tid];
        sjy[tid] = s;
    }
    __syncthreads();
    large_finish_jy<T, NT>(c, a.partJy, sjy, ldj, &c->ticket[1]);
}

// ---------------------------------------------------------------------------------------------
// control kernels (one CTA)
// ---------------------------------------------------------------------------------------------
template <class T> __device__ __forceinline__ void large_reject(LargeCtl<T>* c)
{
    c->lambda *= c->st.lambdaIncrease * c->mu; c->mu *= (T)2;                                      // LS:1103-1104, 1127-1128, 1154-1155
}

// after the (all-reduced) Jacobian products: LS:1052-1110
template <class T>
__global__ void __launch_bounds__(LARGE_CTL_THREADS) large_ctl_mid_kernel(LargeCtl<T>* g)
{
    if (g->done) return;
    MIRB200_PHASE(0);
    constexpr int NT = LARGE_CTL_THREADS;
    extern __shared__ __align__(16) unsigned char ctl_smem[];
    __shared__ __align__(16) unsigned int s_head[large_head_words<T>() + 2];
    LargeCtl<T>* c = large_head_load(g, s_head);     // scalars: c (shared copy);  vectors / matrices: g (global)
    const int tid = threadIdx.x;
    __shared__ int s_flag;
    __syncthreads();
    const int n = c->n;
    CtaQPScratch<T> w;
    w.carve(ctl_smem, n);
    T* vec = reinterpret_cast<T*>(ctl_smem + ((CtaQPScratch<T>::bytes(n) + 15) & ~(size_t)15));
    T* sq = vec; T* sl = vec + n; T* su = vec + 2 * n; T* sx = vec + 3 * n;
    T* sA = vec + 4 * n;                             // packed lower J^T J (undamped), staged once: every P(i, j) below is a shared load
    const int np = n * (n + 1) / 2;
    const bool fresh = c->jacMode != JAC_NONE;

    {
        // packed J^T J -> shared memory (and, when the Jacobian changed, into the control block's own copy: the host
        // all-reduces `packed` in place on every pass).  Sixteen loads in flight per thread: one L2 round trip per batch.
        const T* src = fresh ? g->packed : g->JJ;
        for (int e0 = 0; e0 < np; e0 += NT * 16) {
            T v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) { const int e = e0 + k * NT + tid; v[k] = (e < np) ? __ldcg(src + e) : (T)0; }
#pragma unroll
            for (int k = 0; k < 16; ++k) { const int e = e0 + k * NT + tid; if (e < np) { sA[e] = v[k]; if (fresh) g->JJ[e] = v[k]; } }
        }
    }
    // the n-vectors of the step, one parallel round trip: q = J^T y, l - x, u - x  (LS:1074-1077)
    T xk = (T)0, lk = (T)0, uk = (T)0;
    for (int k = tid; k < n; k += NT) {
        const T q = fresh ? __ldcg(g->packed + np + k) : __ldcg(g->Jy + k);
        if (fresh) g->Jy[k] = q;
        xk = __ldcg(g->x + k); lk = __ldcg(g->l + k); uk = __ldcg(g->u + k);
        sq[k] = q; sl[k] = lk - xk; su[k] = uk - xk; sx[k] = (T)0;
    }
    __syncthreads();
    int exitFlag = 0;
    if (fresh) {
        if (tid < 32) {                                                                            // LS:1053-1062
            // iamax: first index of max |.|; NaN never wins unless it is element 0 (reference BLAS behaviour)
            T bv = (T)-1; int bi = 0;
            for (int k = tid; k < n; k += 32) { const T v = t_abs(sq[k]); if (v > bv) { bv = v; bi = k; } }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const T ov = __shfl_xor_sync(0xffffffffu, bv, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (tid == 0) {
                const T sel = !(sq[0] == sq[0]) ? sq[0] : sq[bi];
                int f = 0;
                if (!(t_abs(sel) > c->st.gradTolerance)) {
                    if (c->age == 0) { c->status = mir_ls_gConverged; c->done = 1; f = 1; }
                    else { c->age = c->maxAge; c->skipRest = 1; f = 1; }
                }
                s_flag = f;
            }
        }
        __syncthreads();
        exitFlag = s_flag;
    }
    MIRB200_PHASE(1);
    if (!exitFlag) {
        if (tid == 0) {
            if (!(c->lambda >= c->st.minLambda)) {                                                 // LS:1067-1072
                T dmax = sA[0];
                for (int i = 1; i < n; ++i) { const T d = sA[tri(i, i)]; if (t_abs(d) > t_abs(dmax)) dmax = d; }
                c->lambda = (T)(0.001 * (double)dmax);
                if (!(c->lambda >= c->st.minLambda)) c->lambda = (T)1;
            }
        }
        __syncthreads();
        const T lambda = c->lambda;
        const PackedLowerShift<T> P{sA, lambda};                                                  // LS:1078-1079: J^T J + lambda I (i >= j)
        unsigned iters = 0, solves = 0;
        MIRB200_PHASE(2);
        const int qs = cta_boxqp<T, NT, true>(c->st.qpSettings, n, P, sq, sl, su, sx, w, iters, solves);   // LS:1080
        __syncthreads();
        MIRB200_PHASE(3);
        bool nan = false;                                                                          // LS:1087-1092
        for (int k = tid; k < n; k += NT) nan = nan || !(sx[k] <= sx[k]);
        const bool anyNan = cta_any<NT>(nan);
        if (tid == 0) {
            c->qpSolves += solves; c->qpIters += iters;
            int f = 0;
            if (qs != mir_qp_solved || anyNan) { c->status = mir_ls_numericError; c->done = 1; f = 1; }   // LS:1080-1092
            s_flag = f;
        }
        __syncthreads();
        if (!s_flag) {
            bool differs = false;                                                                  // LS:1096-1097, 1108-1110
            T trial = (T)0;
            for (int k = tid; k < n; k += NT) {            // (n <= NT: xk, lk, uk are still this thread's element k)
                const T d = add_rn(add_rn(sx[k], xk), -xk);
                sx[k] = d; g->dX[k] = d;
                trial = t_max(t_min(add_rn(d, xk), uk), lk);
                differs = differs || !((trial == xk) && (signbit(trial) == signbit(xk)));
            }
            const bool same = !cta_any<NT>(differs);       // (barrier: sx is complete)
            if (tid == 0) {
                T nd = (T)0;                                                                       // LS:1099
#pragma unroll 8
                for (int k = 0; k < n; ++k) { const T d = sx[k]; nd += d * d; }
                c->nd = nd;
                int take = 0;
                if (!(sqrt_ni(nd) < c->st.maxStep)) { large_reject(c); c->skipRest = 1; }          // LS:1101-1106
                else {
                    take = 1;
                    ++c->fCalls;                                                                   // LS:1112
                    c->doEval = same ? 0 : 1;      // f(xt) == y bit for bit when xt == x: the evaluation is skipped, trial = residual
                    if (same) c->rr = c->residual;
                }
                s_flag = take;
            }
            __syncthreads();
            if (s_flag) for (int k = tid; k < n; k += NT) g->xt[k] = trial;
        }
    }
    __syncthreads();
    large_head_store(g, s_head);
    MIRB200_PHASE(4);
    MIRB200_PHASE_DUMP();
}

// after the (all-reduced) trial residual: LS:1117-1175, then the guards of the next pass
template <class T>
__global__ void __launch_bounds__(LARGE_CTL_THREADS) large_ctl_post_kernel(LargeCtl<T>* g)
{
    if (g->done) return;
    __shared__ T sxv[LARGE_NMAX], sdx[LARGE_NMAX], sjy[LARGE_NMAX], slo[LARGE_NMAX], sup[LARGE_NMAX];
    __shared__ __align__(16) unsigned int s_head[large_head_words<T>() + 2];
    __shared__ int s_go;
    LargeCtl<T>* c = large_head_load(g, s_head);     // scalars: c (shared copy);  vectors / matrices: g (global)
    const int tid = threadIdx.x;
    const int n = g->n;

    // one parallel round trip for every n-vector the serial parts below walk over
    T xtk = (T)0;
    if (tid < n) {
        sxv[tid] = __ldcg(g->x + tid); sdx[tid] = __ldcg(g->dX + tid); sjy[tid] = __ldcg(g->Jy + tid);
        slo[tid] = __ldcg(g->l + tid); sup[tid] = __ldcg(g->u + tid); xtk = __ldcg(g->xt + tid);
    }
    __syncthreads();
    const int initPhase = c->initPhase;
    __syncthreads();                                 // (every thread has read the flag before thread 0 clears it in the shared copy)
    if (initPhase) {                                                                               // LS:953-971
        if (tid == 0) {
            c->initPhase = 0;
            c->residual = c->rr; c->ysel ^= 1; c->fCalls = 1;
            c->fConverged = (c->residual <= c->st.maxGoodResidual) ? 1 : 0;
            c->needJacobian = 1; c->age = c->maxAge; c->lambda = (T)0; c->iterations = 0; c->mu = (T)1;
            c->status = mir_ls_maxIterations; c->deltaX_dot = (T)0;
            large_begin_pass(c, sxv, sjy, slo, sup, g->JJ);
        }
        __syncthreads();
        large_head_store(g, s_head);
        return;
    }

    if (tid == 0) {
        int go = 0;
        if (!c->skipRest) {
            const T trial = c->rr;
            if (!(trial <= Num<T>::inf())) { c->status = mir_ls_numericError; c->done = 1; }       // LS:1117-1122
            else {
                const T improvement = c->residual - trial;                                         // LS:1124
                if (!(improvement > (T)0)) large_reject(c);                                        // LS:1125-1130
                else go = 1;
            }
        }
        s_go = go;
    }
    __syncthreads();
    if (s_go) {
        // accepted: LS:1132-1139
        if (tid < n) { g->x[tid] = xtk; sxv[tid] = xtk; }
        // symv(Lower, 1, JJ, deltaX, 2, Jy) with the undamped JJ, then pred = -Jy . deltaX          LS:1141-1142
        // (row i of the packed matrix, j ascending; 32 loads in flight per thread)
        if (tid < n) {
            const int i = tid;
            T acc = (T)0;
            for (int j0 = 0; j0 < n; j0 += 32) {
                T v[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) { const int j = j0 + k; v[k] = (j < n) ? __ldcg(g->JJ + trisym(i, j)) : (T)0; }
#pragma unroll
                for (int k = 0; k < 32; ++k) { const int j = j0 + k; if (j < n) acc += v[k] * sdx[j]; }
            }
            const T v = acc + (T)2 * sjy[i];
            g->Jy[i] = v;
            sjy[i] = v;
        }
        __syncthreads();
        if (tid == 0) {
            const T improvement = c->residual - c->rr;
            c->needJacobian = 1; c->mu = (T)1; ++c->iterations; ++c->accepted; c->ysel ^= 1;
            c->residual = c->rr;
            c->fConverged = (c->residual <= c->st.maxGoodResidual) ? 1 : 0;
            c->deltaX_dot = c->nd;
            T pred = (T)0;
#pragma unroll 8
            for (int k = 0; k < n; ++k) pred += sjy[k] * sdx[k];
            pred = -pred;
            if (!(pred > (T)0)) { c->status = mir_ls_furtherImprovement; c->done = 1; }            // LS:1144-1148
            else {
                const T rho = div_ni(pred, improvement);                                           // LS:1150
                if (rho < c->st.minStepQuality) large_reject(c);                                   // LS:1152-1156
                else if (rho >= c->st.goodStepQuality) c->lambda = t_max(c->st.lambdaDecrease * c->lambda * c->mu, c->st.minLambda);   // LS:1158-1161
                // LS:1164: !(sqrt(dd) > absTol && nrm2(x) > sqrt(dd) * relTol)
                T xmax = (T)0;
#pragma unroll 8
                for (int k = 0; k < n; ++k) xmax = t_max(xmax, t_abs(sxv[k]));
                T xn = (T)0;
                if (xmax > (T)0) {
                    const T inv = rcp_ni(xmax);
                    T ss = (T)0;
#pragma unroll 8
                    for (int k = 0; k < n; ++k) { const T v = sxv[k] * inv; ss += v * v; }
                    xn = xmax * sqrt_ni(ss);
                }
                const T sd = sqrt_ni(c->deltaX_dot);
                if (!(sd > c->st.absTolerance && xn > sd * c->st.relTolerance)) {                  // LS:1164-1173
                    if (c->age == 0) { c->status = mir_ls_xConverged; c->done = 1; }
                    else c->age = c->maxAge;
                }
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (!*(volatile int*)&c->done) {             // (volatile: not hoisted above the branch into the other threads)
            if (!(c->iterations < c->st.maxIterations)) { c->status = mir_ls_maxIterations; c->done = 1; }   // LS:1175
            else large_begin_pass(c, sxv, sjy, slo, sup, g->JJ);
        }
    }
    __syncthreads();
    large_head_store(g, s_head);
}

}  // namespace mirb200
