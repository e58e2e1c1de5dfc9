/*
 * models_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Host residual / Jacobian callbacks (the reference's LeastSquaresFunctionBetterC /
 * LeastSquaresJacobianBetterC signatures, LS:78-80) for the built-in models named by
 * mir_model_id, plus an OpenMP driver that runs the oracle LM (lm_oracle.cpp ==
 * LS:877-1176) over a batch of independent problems.  The formulas are written here
 * independently of the CUDA functors (mir_optim_b200/csrc/models.cuh); the first five are the
 * reference's own unit-test problems (LS:217-434).  exp is oracle_math::exp_repro (repro_math.h)
 * and the file is compiled with -ffp-contract=off, so every operation rounds exactly like the
 * device functor's explicitly rounded operation sequence: CPU and GPU see identical residuals.
 */
#include "../include/mir_optim_b200.h"
#include "repro_math.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

/* context handed to the callbacks: one problem's data */
typedef struct oracle_model_ctx {
    int         model;
    const void* t;
    const void* y;
} oracle_model_ctx;

}  // extern "C"

namespace {

using oracle_math::exp_repro;

template <class T>
void model_f(void* vctx, size_t m, size_t n, const T* p, T* r)
{
    const oracle_model_ctx* c = static_cast<const oracle_model_ctx*>(vctx);
    const T* t = static_cast<const T*>(c->t);
    const T* y = static_cast<const T*>(c->y);
    switch (c->model) {
    case MIR_MODEL_LINEAR2:    r[0] = p[0]; r[1] = 2 - p[1]; break;                        // LS:230-234
    case MIR_MODEL_ROSENBROCK: r[0] = 10 * (p[1] - p[0] * p[0]); r[1] = 1 - p[0]; break;   // LS:261-265
    case MIR_MODEL_SQRTCIRCLE: r[0] = std::sqrt(1 - (p[0] * p[0] + p[1] * p[1])); break;   // LS:427-430
    case MIR_MODEL_EXPDECAY2:                                                               // LS:347, 360
        for (size_t i = 0; i < m; ++i) r[i] = p[0] * exp_repro(-t[i] * p[1]) - y[i];
        break;
    case MIR_MODEL_EXPTAU3:                                                                 // LS:378, 390
        for (size_t i = 0; i < m; ++i) r[i] = p[0] * exp_repro(-t[i] / p[1]) + p[2] - y[i];
        break;
    case MIR_MODEL_EXPDECAY3:
        for (size_t i = 0; i < m; ++i) r[i] = p[0] * exp_repro(-p[1] * t[i]) + p[2] - y[i];
        break;
    case MIR_MODEL_GAUSS4: {
        const T is = 1 / p[2];
        for (size_t i = 0; i < m; ++i) {
            T z = (t[i] - p[1]) * is;
            r[i] = p[0] * exp_repro((T)-0.5 * (z * z)) + p[3] - y[i];
        }
        break; }
    case MIR_MODEL_SUMEXP:
        for (size_t i = 0; i < m; ++i) {
            T acc = 0;
            for (size_t k = 0; k + 1 < n; k += 2) acc += p[k] * exp_repro(-p[k + 1] * t[i]);
            r[i] = acc - y[i];
        }
        break;
    case MIR_MODEL_GAUSSMIX:
        // (rows are independent: the OpenMP loop gives the same bits for any thread count; it only matters for the
        //  single large problem of configs[3], where the CPU baseline is meant to use all host cores)
#pragma omp parallel for schedule(static) if (m >= 65536)
        for (long long i = 0; i < (long long)m; ++i) {
            T acc = p[n - 2] + p[n - 1] * t[i];
            for (size_t k = 0; k + 3 <= n - 2; k += 3) {
                T z = (t[i] - p[k + 1]) * (1 / p[k + 2]);
                acc += p[k] * exp_repro((T)-0.5 * (z * z));
            }
            r[i] = acc - y[i];
        }
        break;
    default: for (size_t i = 0; i < m; ++i) r[i] = NAN;
    }
}

template <class T>
void model_g(void* vctx, size_t m, size_t n, const T* p, T* J)
{
    const oracle_model_ctx* c = static_cast<const oracle_model_ctx*>(vctx);
    const T* t = static_cast<const T*>(c->t);
    switch (c->model) {
    case MIR_MODEL_LINEAR2:    J[0] = 1; J[1] = 0; J[2] = 0; J[3] = -1; break;             // LS:235-241
    case MIR_MODEL_ROSENBROCK: J[0] = -20 * p[0]; J[1] = 10; J[2] = -1; J[3] = 0; break;   // LS:295-301
    case MIR_MODEL_SQRTCIRCLE: {
        T s = std::sqrt(1 - (p[0] * p[0] + p[1] * p[1]));
        J[0] = -p[0] / s; J[1] = -p[1] / s; break; }
    case MIR_MODEL_EXPDECAY2:
        for (size_t i = 0; i < m; ++i) {
            T e = exp_repro(-t[i] * p[1]);
            J[i * n + 0] = e; J[i * n + 1] = -(p[0] * t[i]) * e;
        }
        break;
    case MIR_MODEL_EXPTAU3:
        for (size_t i = 0; i < m; ++i) {
            T e = exp_repro(-t[i] / p[1]);
            J[i * n + 0] = e; J[i * n + 1] = ((p[0] * e) * t[i]) / (p[1] * p[1]); J[i * n + 2] = 1;
        }
        break;
    case MIR_MODEL_EXPDECAY3:
        for (size_t i = 0; i < m; ++i) {
            T e = exp_repro(-p[1] * t[i]);
            J[i * n + 0] = e; J[i * n + 1] = -(p[0] * t[i]) * e; J[i * n + 2] = 1;
        }
        break;
    case MIR_MODEL_GAUSS4: {
        const T is = 1 / p[2];
        for (size_t i = 0; i < m; ++i) {
            T z = (t[i] - p[1]) * is;
            T e = exp_repro((T)-0.5 * (z * z));
            T ae = p[0] * e;
            J[i * n + 0] = e;
            J[i * n + 1] = (ae * z) * is;
            J[i * n + 2] = (ae * (z * z)) * is;
            J[i * n + 3] = 1;
        }
        break; }
    case MIR_MODEL_SUMEXP:
        for (size_t i = 0; i < m; ++i)
            for (size_t k = 0; k + 1 < n; k += 2) {
                T e = exp_repro(-p[k + 1] * t[i]);
                J[i * n + k] = e; J[i * n + k + 1] = -(p[k] * t[i]) * e;
            }
        break;
    case MIR_MODEL_GAUSSMIX:
#pragma omp parallel for schedule(static) if (m >= 65536)
        for (long long i = 0; i < (long long)m; ++i) {
            for (size_t k = 0; k + 3 <= n - 2; k += 3) {
                T is = 1 / p[k + 2];
                T z = (t[i] - p[k + 1]) * is;
                T e = exp_repro((T)-0.5 * (z * z));
                T ae = p[k] * e;
                J[i * n + k] = e; J[i * n + k + 1] = (ae * z) * is; J[i * n + k + 2] = (ae * (z * z)) * is;
            }
            J[i * n + n - 2] = 1; J[i * n + n - 1] = t[i];
        }
        break;
    default: for (size_t i = 0; i < m * n; ++i) J[i] = NAN;
    }
}

template <class T> struct API;
template <> struct API<double> {
    using S = mir_least_squares_settings_d; using R = mir_least_squares_result_d; using Sl = mir_slice_d;
    static R run(const S* s, size_t m, size_t n, double* x, const double* l, const double* u, Sl w, mir_slice_i iw,
                 void* ctx, bool fd) {
        return mir_optimize_least_squares_d(s, m, n, x, l, u, w, iw, ctx, model_f<double>, ctx,
                                            fd ? nullptr : model_g<double>, nullptr, nullptr);
    }
};
template <> struct API<float> {
    using S = mir_least_squares_settings_s; using R = mir_least_squares_result_s; using Sl = mir_slice_s;
    static R run(const S* s, size_t m, size_t n, float* x, const float* l, const float* u, Sl w, mir_slice_i iw,
                 void* ctx, bool fd) {
        return mir_optimize_least_squares_s(s, m, n, x, l, u, w, iw, ctx, model_f<float>, ctx,
                                            fd ? nullptr : model_g<float>, nullptr, nullptr);
    }
};

template <class T>
int batched(const typename API<T>::S* settings, const mir_model_desc* model, size_t batch, size_t m, size_t n,
            T* x, const T* l, const T* u, size_t bound_stride, typename API<T>::R* results, int nthreads)
{
    const bool fd = (model->flags & MIR_MODEL_FD_JACOBIAN) != 0;
    const bool per = (model->flags & MIR_MODEL_GRID_PER_PROBLEM) != 0;
    const size_t wl = mir_least_squares_work_length(m, n), iwl = mir_least_squares_iwork_length(m, n);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    if ((size_t)nthreads > batch) nthreads = (int)(batch ? batch : 1);   // one problem: leave the threads to OpenBLAS and the row loops
#pragma omp parallel num_threads(nthreads)
    {
        std::vector<T> work(wl + 8);
        std::vector<mir_lapackint> iwork(iwl + 8);
#pragma omp for schedule(dynamic, 16)
        for (long long b = 0; b < (long long)batch; ++b) {
            oracle_model_ctx ctx;
            ctx.model = (int)model->model;
            ctx.t = model->t ? static_cast<const T*>(model->t) + (per ? (size_t)b * m : 0) : nullptr;
            ctx.y = model->y ? static_cast<const T*>(model->y) + (size_t)b * m : nullptr;
            typename API<T>::Sl w{wl, work.data()};
            mir_slice_i iw{iwl, iwork.data()};
            results[b] = API<T>::run(settings, m, n, x + (size_t)b * n, l + (size_t)b * bound_stride,
                                     u + (size_t)b * bound_stride, w, iw, &ctx, fd);
        }
    }
    return nthreads;
}

}  // namespace

extern "C" {

void oracle_model_f_d(void* c, size_t m, size_t n, const double* x, double* y) { model_f<double>(c, m, n, x, y); }
void oracle_model_g_d(void* c, size_t m, size_t n, const double* x, double* J) { model_g<double>(c, m, n, x, J); }
void oracle_model_f_s(void* c, size_t m, size_t n, const float* x, float* y)   { model_f<float>(c, m, n, x, y); }
void oracle_model_g_s(void* c, size_t m, size_t n, const float* x, float* J)   { model_g<float>(c, m, n, x, J); }

/* Run the oracle LM on every problem of a batch (host pointers).  Returns the thread count used. */
int oracle_batched_d(const mir_least_squares_settings_d* s, const mir_model_desc* model, size_t batch, size_t m,
                     size_t n, double* x, const double* l, const double* u, size_t bound_stride,
                     mir_least_squares_result_d* results, int nthreads)
{ return batched<double>(s, model, batch, m, n, x, l, u, bound_stride, results, nthreads); }

int oracle_batched_s(const mir_least_squares_settings_s* s, const mir_model_desc* model, size_t batch, size_t m,
                     size_t n, float* x, const float* l, const float* u, size_t bound_stride,
                     mir_least_squares_result_s* results, int nthreads)
{ return batched<float>(s, model, batch, m, n, x, l, u, bound_stride, results, nthreads); }

int oracle_solve_box_qp_d(const mir_box_qp_settings_d*, size_t, double*, const double*, const double*, const double*, double*, unsigned*);
int oracle_solve_box_qp_s(const mir_box_qp_settings_s*, size_t, float*, const float*, const float*, const float*, float*, unsigned*);

/* Batched BOXCQP (BQ:85-102 per problem).  P is copied per problem because the solver mirrors
 * the lower triangle into the upper one. */
int oracle_box_qp_batched_d(const mir_box_qp_settings_d* s, size_t batch, size_t n, const double* P, const double* q,
                            const double* l, const double* u, double* x, int32_t* status, uint32_t* iters, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        std::vector<double> Pc(n * n);
#pragma omp for schedule(dynamic, 16)
        for (long long b = 0; b < (long long)batch; ++b) {
            std::memcpy(Pc.data(), P + (size_t)b * n * n, sizeof(double) * n * n);
            unsigned it = 0;
            status[b] = oracle_solve_box_qp_d(s, n, Pc.data(), q + b * n, l + b * n, u + b * n, x + b * n, &it);
            if (iters) iters[b] = it;
        }
    }
    return nthreads;
}
int oracle_box_qp_batched_s(const mir_box_qp_settings_s* s, size_t batch, size_t n, const float* P, const float* q,
                            const float* l, const float* u, float* x, int32_t* status, uint32_t* iters, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        std::vector<float> Pc(n * n);
#pragma omp for schedule(dynamic, 16)
        for (long long b = 0; b < (long long)batch; ++b) {
            std::memcpy(Pc.data(), P + (size_t)b * n * n, sizeof(float) * n * n);
            unsigned it = 0;
            status[b] = oracle_solve_box_qp_s(s, n, Pc.data(), q + b * n, l + b * n, u + b * n, x + b * n, &it);
            if (iters) iters[b] = it;
        }
    }
    return nthreads;
}

}  // extern "C"
