/*
 * lm_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of mir-optim's Levenberg-Marquardt loop and BOXCQP solver, used only as
 * the parity checker (tests/, __graft_entry__.smoke()) and as the timed CPU baseline
 * (bench.py `cpu_baseline` / `--impl reference`).  Nothing under mir_optim_b200/ may
 * include, link or call this file.
 *
 * What it follows (reference tree /root/reference, LS = source/mir/optim/least_squares.d,
 * BQ = source/mir/optim/boxcqp.d):
 *   lm_impl<T>      <- optimizeLeastSquaresImplGeneric!T       LS:877-1176
 *   boxqp_impl<T>   <- solveBoxQP!T (full overload)            BQ:122-379
 *   apply_bounds    <- applyBounds                             BQ:404-410
 *   all_le          <- allLessOrEqual                          LS:1185-1192
 *   work/iwork len  <- mir_least_squares_(i)work_length        LS:642-656, BQ:36-50
 *   status strings  <- leastSquaresStatusString                LS:528-557
 *   settings .init  <- LeastSquaresSettings!T / BoxQPSettings  LS:85-123, BQ:56-71
 *
 * The reference is D and cannot be compiled in this image (no ldc2/dmd/gdc/dub), and its
 * arithmetic lives in un-vendored third-party packages: mir-blas (>=1.x, wrappers over CBLAS),
 * mir-lapack >=1.2.3 (`posvx`), mir-algorithm >=3.7.19 (`Summator!(T, Summation.kbn)`), see
 * dub.sdl:7-8; there is no lock file.  Those calls are made here against the REAL
 * BLAS/LAPACK: OpenBLAS 0.3.31.dev as bundled with scipy 1.18.1
 * (site-packages/scipy.libs/libscipy_openblas-*.so, symbols scipy_cblas_*, scipy_dposvx_,
 * scipy_sposvx_), loaded with dlopen by oracle_init().  KBN summation is restated from its
 * published definition (Neumaier 1974) below.
 *
 * Pinning: tests/test_oracle_reference_scenarios.py checks this file against every
 * assertion of the reference's own unit tests (LS:217-434 T1-T6, BQ:381-402) -- all double.
 * The reference has no float test and its float entry point passes m=2 (LS:629), so the
 * float instantiation is "parity unpinned": it is this restatement with the real m.
 *
 * Exported with the reference's extern(C) names and struct layouts (include/mir_optim_b200.h,
 * part 1), so the same ctypes binding drives the oracle and the CUDA library.
 */
#include "../include/mir_optim_b200.h"

#include <dlfcn.h>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <type_traits>

namespace {

// ---------------------------------------------------------------------------------------
// BLAS / LAPACK binding (dlopen of scipy's OpenBLAS; CBLAS enums are the standard values)
// ---------------------------------------------------------------------------------------
enum { RowMajor = 101, NoTrans = 111, Trans = 112, Upper = 121, Lower = 122 };

template <class T> struct Blas {
    T      (*dot )(int, const T*, int, const T*, int);
    T      (*nrm2)(int, const T*, int);
    void   (*axpy)(int, T, const T*, int, T*, int);
    void   (*scal)(int, T, T*, int);
    void   (*copy)(int, const T*, int, T*, int);
    void   (*swap)(int, T*, int, T*, int);
    size_t (*iamax)(int, const T*, int);
    void   (*gemv)(int, int, int, int, T, const T*, int, const T*, int, T, T*, int);
    void   (*ger )(int, int, int, T, const T*, int, const T*, int, T*, int);
    void   (*symv)(int, int, int, T, const T*, int, const T*, int, T, T*, int);
    void   (*syrk)(int, int, int, int, int, T, const T*, int, T, T*, int);
    // ?posvx_(fact, uplo, n, nrhs, a, lda, af, ldaf, equed, s, b, ldb, x, ldx, rcond, ferr, berr,
    //         work, iwork, info, [hidden fortran string lengths])
    void   (*posvx)(const char*, const char*, const int*, const int*, T*, const int*, T*, const int*,
                    char*, T*, T*, const int*, T*, const int*, T*, T*, T*, T*, int*, int*,
                    size_t, size_t, size_t);
    static Blas& get() { static Blas b{}; return b; }
};

void* g_lib = nullptr;
void (*g_set_threads)(int) = nullptr;

template <class F> bool load(F& fn, const char* name) {
    fn = reinterpret_cast<F>(dlsym(g_lib, name));
    if (!fn) std::fprintf(stderr, "lm_oracle: missing symbol %s\n", name);
    return fn != nullptr;
}

template <class T> bool load_blas(char p /* 'd' or 's' */) {
    auto& b = Blas<T>::get();
    char nm[64];
    bool ok = true;
    auto N = [&](const char* fmt) { std::snprintf(nm, sizeof nm, fmt, p); return nm; };
    ok &= load(b.dot,  N("scipy_cblas_%cdot"));
    ok &= load(b.nrm2, N("scipy_cblas_%cnrm2"));
    ok &= load(b.axpy, N("scipy_cblas_%caxpy"));
    ok &= load(b.scal, N("scipy_cblas_%cscal"));
    ok &= load(b.copy, N("scipy_cblas_%ccopy"));
    ok &= load(b.swap, N("scipy_cblas_%cswap"));
    ok &= load(b.iamax, N("scipy_cblas_i%camax"));
    ok &= load(b.gemv, N("scipy_cblas_%cgemv"));
    ok &= load(b.ger,  N("scipy_cblas_%cger"));
    ok &= load(b.symv, N("scipy_cblas_%csymv"));
    ok &= load(b.syrk, N("scipy_cblas_%csyrk"));
    ok &= load(b.posvx, N("scipy_%cposvx_"));
    return ok;
}

// ---------------------------------------------------------------------------------------
// Type plumbing: reference PODs per precision
// ---------------------------------------------------------------------------------------
template <class T> struct Types;
template <> struct Types<double> {
    using Settings = mir_least_squares_settings_d; using Result = mir_least_squares_result_d;
    using QPSettings = mir_box_qp_settings_d;
    using F = mir_ls_function_d; using G = mir_ls_jacobian_d;
};
template <> struct Types<float> {
    using Settings = mir_least_squares_settings_s; using Result = mir_least_squares_result_s;
    using QPSettings = mir_box_qp_settings_s;
    using F = mir_ls_function_s; using G = mir_ls_jacobian_s;
};

// allLessOrEqual, LS:1185-1192 (false as soon as a NaN is involved)
template <class T> bool all_le(const T* a, const T* b, size_t n) {
    for (size_t i = 0; i < n; ++i) if (!(a[i] <= b[i])) return false;
    return true;
}

// applyBounds, BQ:404-410: x = fmax(fmin(x, u), l)
template <class T> void apply_bounds(T* x, const T* l, const T* u, size_t n) {
    for (size_t i = 0; i < n; ++i) x[i] = std::fmax(std::fmin(x[i], u[i]), l[i]);
}

// mir.math.sum Summator!(T, Summation.kbn): Kahan-Babuska-Neumaier compensated sum.
template <class T> struct KBN {
    T s, c;
    explicit KBN(T v) : s(v), c(0) {}
    void put(T v) {
        volatile T t = s + v;               // volatile: forbid re-association
        if (std::fabs(s) >= std::fabs(v)) { volatile T d = s - t; c += d + v; }
        else                              { volatile T d = v - t; c += d + s; }
        s = t;
    }
    T sum() const { return s + c; }
};

// ---------------------------------------------------------------------------------------
// BOXCQP, BQ:122-379.  P is n x n row-major ("Canonical"), only the lower triangle is read.
// work: 2n^2 + 8n elements, iwork: n + ceil(n/4) ints (BQ:36-50).
// ---------------------------------------------------------------------------------------
template <class T>
int boxqp_impl(const typename Types<T>::QPSettings& settings, size_t n, T* P, size_t ldp,
               const T* q, const T* l, const T* u, T* x, bool unconstrainedSolution,
               T* work, mir_lapackint* iwork, bool restoreUpperP, unsigned* iterations_out)
{
    auto& B = Blas<T>::get();
    if (iterations_out) *iterations_out = 0;
    if (n == 0) return mir_qp_solved;                                         // BQ:162-163

    signed char* flags = reinterpret_cast<signed char*>(iwork + n);          // BQ:165, 222
    const int one = 1;

    if (!unconstrainedSolution) {                                             // BQ:168-214
        T* buffer = work;
        /* Pdiagonal */ buffer += n;
        T* scaling = buffer; buffer += n;
        T* b = buffer; buffer += n;
        T* lapackWorkSpace = buffer; buffer += 3 * n;
        T* F = buffer; buffer += n * n;
        T* A = buffer; buffer += n * n;
        for (size_t i = 0; i < n; ++i)                                        // BQ:187-188
            for (size_t k = 0; k <= i; ++k) A[k * n + i] = P[i * ldp + k];
        for (size_t i = 0; i < n; ++i) b[i] = -q[i];                          // BQ:191
        char equed = 'N'; T rcond, ferr, berr; int info = 0; int ni = (int)n;
        B.posvx("E", "L", &ni, &one, A, &ni, F, &ni, &equed, scaling, b, &ni, x, &ni,
                &rcond, &ferr, &berr, lapackWorkSpace, iwork, &info, 1, 1, 1);
        if (info != 0 && info != ni + 1) return mir_qp_numericError;          // BQ:212-213
    }

    {
        bool inside = true;                                                   // BQ:216-219
        for (size_t i = 0; i < n; ++i) if (!(l[i] <= x[i] && x[i] <= u[i])) { inside = false; break; }
        if (inside) return mir_qp_solved;
    }

    unsigned maxIterations = settings.maxIterations;                          // BQ:224-226
    if (!maxIterations) maxIterations = (unsigned)n * 10 + 100;

    T* la = work; work += n;                                                  // BQ:228-232
    T* mu = work; work += n;
    for (size_t i = 0; i < n; ++i) la[i] = mu[i] = 0;

    for (unsigned step = 0; step < maxIterations; ++step) {                   // BQ:234
        if (iterations_out) *iterations_out = step + 1;
        size_t s = 0;
        for (size_t i = 0; i < n; ++i) {                                      // BQ:239-263
            T xl = x[i] - l[i];
            T ux = u[i] - x[i];
            if (xl < 0 || (xl < settings.relTolerance + settings.absTolerance * std::fabs(l[i]) && la[i] >= 0)) {
                flags[i] = -1; x[i] = l[i]; mu[i] = 0;
            } else if (ux < 0 || (ux < settings.relTolerance + settings.absTolerance * std::fabs(u[i]) && mu[i] >= 0)) {
                flags[i] = +1; x[i] = u[i]; la[i] = 0;
            } else {
                flags[i] = 0; iwork[s++] = (mir_lapackint)i; mu[i] = 0; la[i] = 0;
            }
        }
        if (s == n) break;                                                    // BQ:265-266 -> maxIterations

        {
            mir_lapackint* SI = iwork;                                        // BQ:269-280
            T* buffer = work;
            T* scaling = buffer; buffer += s;
            T* sX = buffer; buffer += s;
            T* b = buffer; buffer += s;
            T* lapackWorkSpace = buffer; buffer += 3 * s;
            T* F = buffer; buffer += s * s;
            T* A = buffer; buffer += s * s;

            for (size_t ii = 0; ii < s; ++ii) {                               // BQ:282-305
                size_t i = (size_t)SI[ii];
                KBN<T> sum(q[i]);
                size_t jj = 0;
                for (size_t j = 0; j < i; ++j) {
                    T Pij = P[i * ldp + j];
                    if (flags[j]) sum.put(Pij * (flags[j] < 0 ? l : u)[j]);
                    else A[(jj++) * s + ii] = Pij;
                }
                for (size_t j = i; j < n; ++j) {
                    T Pji = P[j * ldp + i];
                    if (flags[j]) sum.put(Pji * (flags[j] < 0 ? l : u)[j]);
                    else A[ii * s + (jj++)] = Pji;
                }
                b[ii] = -sum.sum();
            }
            if (s) {                                                          // BQ:307-325
                char equed = 'N'; T rcond, ferr, berr; int info = 0; int si = (int)s;
                B.posvx("E", "L", &si, &one, A, &si, F, &si, &equed, scaling, b, &si, sX, &si,
                        &rcond, &ferr, &berr, lapackWorkSpace, SI, &info, 1, 1, 1);
                if (info != 0 && info != si + 1) return mir_qp_numericError;
            }
            size_t ii = 0;                                                    // BQ:327-329
            for (size_t i = 0; i < n; ++i) if (flags[i] == 0) x[i] = sX[ii++];
        }

        for (size_t i = 0; i < n; ++i) if (flags[i]) {                        // BQ:333-337
            T val = B.dot((int)i, P + i * ldp, 1, x, 1)
                  + B.dot((int)(n - i), P + i * ldp + i, (int)ldp, x + i, 1) + q[i];
            if (flags[i] < 0) la[i] = val; else mu[i] = -val;
        }

        bool again = false;                                                   // BQ:339-347
        for (size_t i = 0; i < n && !again; ++i) {
            if (flags[i] < 0)      { if (!(la[i] >= 0)) again = true; }
            else if (flags[i] > 0) { if (!(mu[i] >= 0)) again = true; }
            else                   { if (!(x[i] >= l[i] && x[i] <= u[i])) again = true; }
        }
        if (again) continue;

        apply_bounds(x, l, u, n);                                             // BQ:349
        if (restoreUpperP)                                                    // BQ:365-373
            for (size_t i = 0; i < n; ++i)
                for (size_t j = i + 1; j < n; ++j) P[i * ldp + j] = P[j * ldp + i];
        return mir_qp_solved;
    }
    return mir_qp_maxIterations;                                              // BQ:378
}

// ---------------------------------------------------------------------------------------
// LM, LS:877-1176
// ---------------------------------------------------------------------------------------
template <class T> struct Consts {
    static T sqrt_max()        { return std::sqrt(std::numeric_limits<T>::max()); }
    static T sqrt_min_normal() { return std::sqrt(std::numeric_limits<T>::min()); }
};

struct OracleCounters { unsigned long long passes, qp_solves_iters; };
thread_local OracleCounters g_counters;

template <class T>
typename Types<T>::Result lm_impl(const typename Types<T>::Settings& st, size_t m, size_t n_, T* x,
                                  const T* lower, const T* upper, T* work, mir_lapackint* iwork,
                                  void* fctx, typename Types<T>::F f, void* gctx, typename Types<T>::G g,
                                  void* tmctx, mir_ls_thread_manager tm)
{
    auto& B = Blas<T>::get();
    typename Types<T>::Result ret;
    ret.status = mir_ls_numericError; ret.iterations = 0; ret.fCalls = 0; ret.gCalls = 0;   // LS:131-142
    ret.residual = std::numeric_limits<T>::infinity(); ret.lambda = 0;

    const unsigned n = (unsigned)n_;                                                         // LS:911
    const int ni = (int)n, mi = (int)m;
    T* deltaX = work; work += n;                                                             // LS:913-926
    T* Jy = work; work += n;
    T* nBuffer = work; work += n;
    T* JJ = work; work += (size_t)n * n;
    T* J = work; work += m * n;
    T* y = work; work += m;
    T* mBuffer = work; work += m;
    T* qpl = work; work += n;
    T* qpu = work; work += n;
    T* qpwork = work;

    // LS:930-943
    bool finite = true;
    for (unsigned i = 0; i < n; ++i)
        if (!(-std::numeric_limits<T>::infinity() < x[i] && x[i] < std::numeric_limits<T>::infinity())) finite = false;
    if (m == 0 || n == 0 || !finite) { ret.status = mir_ls_badGuess; return ret; }
    if (!all_le(lower, (const T*)x, n) || !all_le((const T*)x, upper, n)) { ret.status = mir_ls_badBounds; return ret; }
    if (!(0 <= st.minStepQuality && st.minStepQuality < 1)) { ret.status = mir_ls_badMinStepQuality; return ret; }
    if (!(0 <= st.goodStepQuality && st.goodStepQuality <= 1)) { ret.status = mir_ls_badGoodStepQuality; return ret; }
    if (!(st.minStepQuality < st.goodStepQuality)) { ret.status = mir_ls_badStepQuality; return ret; }
    if (!(1 <= st.lambdaIncrease && st.lambdaIncrease <= Consts<T>::sqrt_max())) { ret.status = mir_ls_badLambdaParams; return ret; }
    if (!(Consts<T>::sqrt_min_normal() <= st.lambdaDecrease && st.lambdaDecrease <= 1)) { ret.status = mir_ls_badLambdaParams; return ret; }

    const unsigned maxAge = st.maxAge ? st.maxAge : (g ? 3u : 2u * n);                       // LS:945

    f(fctx, m, n, x, y);                                                                     // LS:953-956
    ++ret.fCalls;
    ret.residual = B.dot(mi, y, 1, y, 1);
    bool fConverged = ret.residual <= st.maxGoodResidual;

    bool needJacobian = true;                                                                // LS:959-971
    unsigned age = maxAge;
    ret.lambda = 0;
    ret.iterations = 0;
    T deltaX_dot = 0;
    T mu = 1;
    const T suspiciousMu = 16;
    ret.status = mir_ls_maxIterations;

    // FD-Jacobian task body, LS:1019-1048 (serial thread-manager semantics: totalThreads=1, threadId=0)
    struct FDState {
        const typename Types<T>::Settings* st; size_t m; unsigned n; T* x; const T* lower; const T* upper;
        T* JJ; T* J; T* mBuffer; mir_lapackint* iwork; void* fctx; typename Types<T>::F f;
    } fd{&st, m, n, x, lower, upper, JJ, J, mBuffer, iwork, fctx, f};
    auto fd_task = [](mir_ls_task task, unsigned totalThreads, unsigned threadId, unsigned j) {
        FDState& s = *static_cast<FDState*>(task.context);
        auto& B = Blas<T>::get();
        unsigned idx = totalThreads >= s.n ? j : threadId;                                   // LS:1022
        T* p = s.JJ + (size_t)idx * s.n;
        if (s.iwork[idx]++ == 0) B.copy((int)s.n, s.x, 1, p, 1);                             // LS:1024-1025
        T save = p[j];
        T xmh = save - s.st->jacobianEpsilon;
        T xph = save + s.st->jacobianEpsilon;
        xmh = std::fmax(xmh, s.lower[j]);
        xph = std::fmin(xph, s.upper[j]);
        T* Jj = s.J + j;                                                                     // column j, stride n
        T twh = xph - xmh;
        if (twh != 0) {                                                                      // LS:1033-1043
            p[j] = xph;
            s.f(s.fctx, s.m, s.n, p, s.mBuffer);
            B.copy((int)s.m, s.mBuffer, 1, Jj, (int)s.n);
            p[j] = xmh;
            s.f(s.fctx, s.m, s.n, p, s.mBuffer);
            p[j] = save;
            B.axpy((int)s.m, (T)-1, s.mBuffer, 1, Jj, (int)s.n);
            B.scal((int)s.m, 1 / twh, Jj, (int)s.n);
        } else {
            for (size_t i = 0; i < s.m; ++i) Jj[i * s.n] = 0;                                // LS:1046
        }
    };

    do {                                                                                     // LS:972
        ++g_counters.passes;
        if (fConverged) { ret.status = mir_ls_fConverged; break; }                           // LS:974-978
        if (!(ret.lambda <= st.maxLambda)) { ret.status = mir_ls_furtherImprovement; break; } // LS:979-983
        if (mu > suspiciousMu && age) { needJacobian = true; age = maxAge; mu = 1; }         // LS:984-989
        if (!all_le((const T*)x, (const T*)x, n)) { ret.status = mir_ls_numericError; break; } // LS:990-995
        if (needJacobian) {                                                                  // LS:996
            needJacobian = false;
            if (age < maxAge) {                                                              // LS:999-1007 Broyden
                age++;
                T d = 1 / deltaX_dot;
                B.axpy(mi, (T)-1, y, 1, mBuffer, 1);
                B.gemv(RowMajor, NoTrans, mi, ni, (T)1, J, ni, deltaX, 1, (T)1, mBuffer, 1);
                B.scal(mi, -d, mBuffer, 1);
                B.ger(RowMajor, mi, ni, (T)1, mBuffer, 1, deltaX, 1, J, ni);
            } else {
                age = 0;                                                                     // LS:1010
                if (g) { g(gctx, m, n, x, J); ret.gCalls += 1; }                             // LS:1011-1015
                else {
                    for (unsigned i = 0; i < n; ++i) iwork[i] = 0;                           // LS:1018
                    mir_ls_task task{&fd, nullptr};
                    if (tm) tm(tmctx, n, task, fd_task);
                    else for (unsigned j = 0; j < n; ++j) fd_task(task, 1, 0, j);            // LS:947-951
                    unsigned sum = 0;
                    for (unsigned i = 0; i < n; ++i) sum += (unsigned)iwork[i];
                    ret.fCalls += sum;                                                       // LS:1049
                }
            }
            B.gemv(RowMajor, Trans, mi, ni, (T)1, J, ni, y, 1, (T)0, Jy, 1);                 // LS:1052
            if (!(std::fabs(Jy[B.iamax(ni, Jy, 1)]) > st.gradTolerance)) {                   // LS:1053-1062
                if (age == 0) { ret.status = mir_ls_gConverged; break; }
                age = maxAge;
                continue;
            }
        }

        B.syrk(RowMajor, Lower, Trans, ni, mi, (T)1, J, ni, (T)0, JJ, ni);                   // LS:1065

        if (!(ret.lambda >= st.minLambda)) {                                                 // LS:1067-1072
            ret.lambda = (T)(0.001 * JJ[(size_t)B.iamax(ni, JJ, ni + 1) * (n + 1)]);
            if (!(ret.lambda >= st.minLambda)) ret.lambda = 1;
        }

        for (unsigned i = 0; i < n; ++i) qpl[i] = lower[i];                                  // LS:1074-1079
        B.axpy(ni, (T)-1, x, 1, qpl, 1);
        for (unsigned i = 0; i < n; ++i) qpu[i] = upper[i];
        B.axpy(ni, (T)-1, x, 1, qpu, 1);
        for (unsigned i = 0; i < n; ++i) nBuffer[i] = JJ[(size_t)i * (n + 1)];
        for (unsigned i = 0; i < n; ++i) JJ[(size_t)i * (n + 1)] += ret.lambda;
        unsigned qpit = 0;
        int qps = boxqp_impl<T>(st.qpSettings, n, JJ, n, Jy, qpl, qpu, deltaX, false, qpwork, iwork, false, &qpit);
        g_counters.qp_solves_iters += 1 + qpit;
        if (qps != mir_qp_solved) { ret.status = mir_ls_numericError; break; }               // LS:1080-1085
        if (!all_le((const T*)deltaX, (const T*)deltaX, n)) { ret.status = mir_ls_numericError; break; } // LS:1087-1092

        for (unsigned i = 0; i < n; ++i) JJ[(size_t)i * (n + 1)] = nBuffer[i];               // LS:1094

        B.axpy(ni, (T)1, x, 1, deltaX, 1);                                                   // LS:1096-1097
        B.axpy(ni, (T)-1, x, 1, deltaX, 1);

        T newDeltaX_dot = B.dot(ni, deltaX, 1, deltaX, 1);                                   // LS:1099

        if (!(std::sqrt(newDeltaX_dot) < st.maxStep)) {                                      // LS:1101-1106
            ret.lambda *= st.lambdaIncrease * mu; mu *= 2; continue;
        }

        for (unsigned i = 0; i < n; ++i) nBuffer[i] = deltaX[i];                             // LS:1108-1110
        B.axpy(ni, (T)1, x, 1, nBuffer, 1);
        apply_bounds(nBuffer, lower, upper, n);

        ++ret.fCalls;                                                                        // LS:1112-1115
        f(fctx, m, n, nBuffer, mBuffer);
        T trialResidual = B.dot(mi, mBuffer, 1, mBuffer, 1);

        if (!(trialResidual <= std::numeric_limits<T>::infinity())) { ret.status = mir_ls_numericError; break; } // LS:1117-1122

        T improvement = ret.residual - trialResidual;                                        // LS:1124-1130
        if (!(improvement > 0)) { ret.lambda *= st.lambdaIncrease * mu; mu *= 2; continue; }

        needJacobian = true;                                                                 // LS:1132-1139
        mu = 1;
        ret.iterations++;
        for (unsigned i = 0; i < n; ++i) x[i] = nBuffer[i];
        B.swap(mi, mBuffer, 1, y, 1);
        ret.residual = trialResidual;
        fConverged = ret.residual <= st.maxGoodResidual;
        deltaX_dot = newDeltaX_dot;

        B.symv(RowMajor, Lower, ni, (T)1, JJ, ni, deltaX, 1, (T)2, Jy, 1);                   // LS:1141-1142
        T predictedImprovement = -B.dot(ni, Jy, 1, deltaX, 1);

        if (!(predictedImprovement > 0)) { ret.status = mir_ls_furtherImprovement; break; }  // LS:1144-1148

        T rho = predictedImprovement / improvement;                                          // LS:1150

        if (rho < st.minStepQuality) { ret.lambda *= st.lambdaIncrease * mu; mu *= 2; }      // LS:1152-1156
        else if (rho >= st.goodStepQuality)                                                  // LS:1158-1161
            ret.lambda = std::fmax(st.lambdaDecrease * ret.lambda * mu, st.minLambda);

        if (!(std::sqrt(deltaX_dot) > st.absTolerance
              && B.nrm2(ni, x, 1) > std::sqrt(deltaX_dot) * st.relTolerance)) {              // LS:1164-1173
            if (age == 0) { ret.status = mir_ls_xConverged; break; }
            age = maxAge;
            continue;
        }
    } while (ret.iterations < st.maxIterations);                                             // LS:1175
    return ret;
}

template <class S, class T> void settings_init(S* s) {                                       // LS:85-123, BQ:56-71
    const T eps = std::numeric_limits<T>::epsilon();
    s->maxIterations = 1000;
    s->maxAge = 0;
    s->jacobianEpsilon = (T)std::ldexp(1.0, (1 - std::numeric_limits<T>::digits) / 2);       // integer division, LS:98
    s->absTolerance = eps;
    s->relTolerance = 0;
    s->gradTolerance = eps;
    s->maxGoodResidual = eps * eps;
    s->maxStep = std::sqrt(std::numeric_limits<T>::max()) / 16;
    s->maxLambda = std::numeric_limits<T>::max() / 16;
    s->minLambda = std::numeric_limits<T>::min() * 16;
    s->minStepQuality = (T)0.1;
    s->goodStepQuality = (T)0.5;
    s->lambdaIncrease = 2;
    s->lambdaDecrease = (T)(1 / (1.6180339887498948482045868343656381L * 2));
    s->qpSettings.relTolerance = eps * 16;
    s->qpSettings.absTolerance = eps * 16;
    s->qpSettings.maxIterations = 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// exported surface (reference names) + oracle-only helpers
// ---------------------------------------------------------------------------------------
extern "C" {

int oracle_init(const char* openblas_path) {
    if (g_lib) return 0;
    g_lib = dlopen(openblas_path, RTLD_NOW | RTLD_LOCAL);
    if (!g_lib) { std::fprintf(stderr, "lm_oracle: dlopen(%s): %s\n", openblas_path, dlerror()); return 1; }
    bool ok = load_blas<double>('d') & load_blas<float>('s');
    load(g_set_threads, "scipy_openblas_set_num_threads");
    if (g_set_threads) g_set_threads(1);
    return ok ? 0 : 2;
}

void oracle_set_blas_threads(int n) { if (g_set_threads) g_set_threads(n); }
void oracle_counters(unsigned long long* passes, unsigned long long* qp) {
    *passes = g_counters.passes; *qp = g_counters.qp_solves_iters;
}
void oracle_counters_reset(void) { g_counters.passes = 0; g_counters.qp_solves_iters = 0; }

size_t mir_box_qp_work_length(size_t n)  { return n * n * 2 + n * 8; }                       // BQ:36-42
size_t mir_box_qp_iwork_length(size_t n) {                                                   // BQ:47-50
    return n + (n / sizeof(mir_lapackint) + (n % sizeof(mir_lapackint) != 0));
}
size_t mir_least_squares_work_length(size_t m, size_t n) {                                   // LS:642-646
    return mir_box_qp_work_length(n) + n * 5 + n * n + n * m + m * 2;
}
size_t mir_least_squares_iwork_length(size_t m, size_t n) {                                  // LS:651-656
    (void)m; size_t a = mir_box_qp_iwork_length(n); return a > n ? a : n;
}

const char* mir_least_squares_status_string(int st) {                                        // LS:528-557
    switch (st) {
        case mir_ls_furtherImprovement: return "The algorithm cann't improve the solution";
        case mir_ls_maxIterations:      return "Maximum number of iterations reached";
        case mir_ls_xConverged:         return "X converged";
        case mir_ls_gConverged:         return "Jacobian converged";
        case mir_ls_fConverged:         return "Residual is small enough";
        case mir_ls_badBounds:          return "Initial guess must be within bounds.";
        case mir_ls_badGuess:           return "Initial guess must be an array of finite numbers.";
        case mir_ls_badMinStepQuality:  return "0 <= minStepQuality < 1 must hold.";
        case mir_ls_badGoodStepQuality: return "0 < goodStepQuality <= 1 must hold.";
        case mir_ls_badStepQuality:     return "minStepQuality < goodStepQuality must hold.";
        case mir_ls_badLambdaParams:    return "1 <= lambdaIncrease && lambdaIncrease <= T.max.sqrt and T.min_normal.sqrt <= lambdaDecrease && lambdaDecrease <= 1 must hold.";
        case mir_ls_numericError:       return "Numeric Error";
    }
    return nullptr;   // D's `final switch` has no default; out-of-range input is a caller bug
}

void mir_least_squares_init_d (mir_least_squares_settings_d* s) { settings_init<mir_least_squares_settings_d, double>(s); }
void mir_least_squares_init_s (mir_least_squares_settings_s* s) { settings_init<mir_least_squares_settings_s, float>(s); }
void mir_least_squares_reset_d(mir_least_squares_settings_d* s) { settings_init<mir_least_squares_settings_d, double>(s); }
void mir_least_squares_reset_s(mir_least_squares_settings_s* s) { settings_init<mir_least_squares_settings_s, float>(s); }

mir_least_squares_result_d mir_optimize_least_squares_d(
    const mir_least_squares_settings_d* settings, size_t m, size_t n, double* x, const double* l, const double* u,
    mir_slice_d work, mir_slice_i iwork, void* fContext, mir_ls_function_d f, void* gContext, mir_ls_jacobian_d g,
    void* tmContext, mir_ls_thread_manager tm)
{
    return lm_impl<double>(*settings, m, n, x, l, u, work.ptr, iwork.ptr, fContext, f, gContext, g, tmContext, tm);
}

/* NB: the reference's float instantiation runs with m = 2 (LS:629).  The oracle uses the real m. */
mir_least_squares_result_s mir_optimize_least_squares_s(
    const mir_least_squares_settings_s* settings, size_t m, size_t n, float* x, const float* l, const float* u,
    mir_slice_s work, mir_slice_i iwork, void* fContext, mir_ls_function_s f, void* gContext, mir_ls_jacobian_s g,
    void* tmContext, mir_ls_thread_manager tm)
{
    return lm_impl<float>(*settings, m, n, x, l, u, work.ptr, iwork.ptr, fContext, f, gContext, g, tmContext, tm);
}

/* solveBoxQP simple overload, BQ:85-102: allocates, unconstrainedSolution=false, restoreUpperP=true.
 * P is modified only in its upper triangle (mirrored from the lower), as in the reference. */
int oracle_solve_box_qp_d(const mir_box_qp_settings_d* settings, size_t n, double* P, const double* q,
                          const double* l, const double* u, double* x, unsigned* iterations)
{
    mir_box_qp_settings_d def{DBL_EPSILON * 16, DBL_EPSILON * 16, 0};
    double* work = (double*)std::malloc(sizeof(double) * (mir_box_qp_work_length(n) + 1));
    mir_lapackint* iwork = (mir_lapackint*)std::malloc(sizeof(mir_lapackint) * (mir_box_qp_iwork_length(n) + 1));
    int r = boxqp_impl<double>(settings ? *settings : def, n, P, n, q, l, u, x, false, work, iwork, true, iterations);
    std::free(work); std::free(iwork);
    return r;
}
int oracle_solve_box_qp_s(const mir_box_qp_settings_s* settings, size_t n, float* P, const float* q,
                          const float* l, const float* u, float* x, unsigned* iterations)
{
    mir_box_qp_settings_s def{FLT_EPSILON * 16, FLT_EPSILON * 16, 0};
    float* work = (float*)std::malloc(sizeof(float) * (mir_box_qp_work_length(n) + 1));
    mir_lapackint* iwork = (mir_lapackint*)std::malloc(sizeof(mir_lapackint) * (mir_box_qp_iwork_length(n) + 1));
    int r = boxqp_impl<float>(settings ? *settings : def, n, P, n, q, l, u, x, false, work, iwork, true, iterations);
    std::free(work); std::free(iwork);
    return r;
}

/* Direct access to LAPACK ?posvx('E','L') on a row-major symmetric matrix (lower triangle read),
 * for unit-testing the device restatement of posvx.  Returns info. */
int oracle_posvx_d(int n, const double* Arow, const double* b, double* x, char* equed_out) {
    auto& B = Blas<double>::get();
    double* A = (double*)std::malloc(sizeof(double) * (size_t)n * n * 2 + sizeof(double) * 8 * n + 64);
    double* F = A + (size_t)n * n; double* s = F + (size_t)n * n; double* bb = s + n; double* w = bb + n;
    int* iw = (int*)std::malloc(sizeof(int) * (n + 1));
    for (int i = 0; i < n; ++i) for (int k = 0; k <= i; ++k) A[(size_t)k * n + i] = Arow[(size_t)i * n + k];
    for (int i = 0; i < n; ++i) bb[i] = b[i];
    char equed = 'N'; double rcond, ferr, berr; int info = 0; const int one = 1;
    B.posvx("E", "L", &n, &one, A, &n, F, &n, &equed, s, bb, &n, x, &n, &rcond, &ferr, &berr, w, iw, &info, 1, 1, 1);
    if (equed_out) *equed_out = equed;
    std::free(A); std::free(iw);
    return info;
}
int oracle_posvx_s(int n, const float* Arow, const float* b, float* x, char* equed_out) {
    auto& B = Blas<float>::get();
    float* A = (float*)std::malloc(sizeof(float) * (size_t)n * n * 2 + sizeof(float) * 8 * n + 64);
    float* F = A + (size_t)n * n; float* s = F + (size_t)n * n; float* bb = s + n; float* w = bb + n;
    int* iw = (int*)std::malloc(sizeof(int) * (n + 1));
    for (int i = 0; i < n; ++i) for (int k = 0; k <= i; ++k) A[(size_t)k * n + i] = Arow[(size_t)i * n + k];
    for (int i = 0; i < n; ++i) bb[i] = b[i];
    char equed = 'N'; float rcond, ferr, berr; int info = 0; const int one = 1;
    B.posvx("E", "L", &n, &one, A, &n, F, &n, &equed, s, bb, &n, x, &n, &rcond, &ferr, &berr, w, iw, &info, 1, 1, 1);
    if (equed_out) *equed_out = equed;
    std::free(A); std::free(iw);
    return info;
}

}  // extern "C"
