// shim.cpp -- host-only part of the reference's extern(C) surface: workspace sizing, status
// strings, settings initialisers.  (least_squares.d:637-669, 761-792; boxcqp.d:31-51)
#include <cfloat>
#include <cmath>
#include <limits>

#include "../../include/mir_optim_b200.h"

// ABI guard: layouts of SURVEY appendix C (== the D structs at least_squares.d:85-143, boxcqp.d:56-71).
static_assert(sizeof(mir_least_squares_settings_d) == 128, "LeastSquaresSettings!double is 128 bytes");
static_assert(sizeof(mir_least_squares_settings_s) == 68, "LeastSquaresSettings!float is 68 bytes");
static_assert(sizeof(mir_least_squares_result_d) == 32, "LeastSquaresResult!double is 32 bytes");
static_assert(sizeof(mir_least_squares_result_s) == 24, "LeastSquaresResult!float is 24 bytes");
static_assert(sizeof(mir_box_qp_settings_d) == 24 && sizeof(mir_box_qp_settings_s) == 12, "BoxQPSettings");
static_assert(offsetof(mir_least_squares_settings_d, qpSettings) == 104, "qpSettings offset (double)");
static_assert(offsetof(mir_least_squares_settings_s, qpSettings) == 56, "qpSettings offset (float)");
static_assert(offsetof(mir_least_squares_result_d, residual) == 16 && offsetof(mir_least_squares_result_d, lambda) == 24, "Result!double");
static_assert(offsetof(mir_least_squares_result_s, residual) == 16 && offsetof(mir_least_squares_result_s, lambda) == 20, "Result!float");
static_assert(sizeof(mir_slice_d) == 16 && sizeof(mir_ls_task) == 16, "Slice / delegate are 16 bytes");

namespace {
template <class S, class T> void settings_init(S* s)
{
    const T eps = std::numeric_limits<T>::epsilon();
    s->maxIterations = 1000;
    s->maxAge = 0;
    s->jacobianEpsilon = (T)std::ldexp(1.0, (1 - std::numeric_limits<T>::digits) / 2);   // T(2)^^((1-mant_dig)/2)
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
    s->lambdaDecrease = (T)(1 / (1.6180339887498948482045868343656381L * 2));            // 1 / (GoldenRatio * 2)
    s->qpSettings.relTolerance = eps * 16;
    s->qpSettings.absTolerance = eps * 16;
    s->qpSettings.maxIterations = 0;
}
}  // namespace

extern "C" {

size_t mir_box_qp_work_length(size_t n) { return n * n * 2 + n * 8; }
size_t mir_box_qp_iwork_length(size_t n) { return n + (n / sizeof(mir_lapackint) + (n % sizeof(mir_lapackint) != 0)); }
size_t mir_least_squares_work_length(size_t m, size_t n) { return mir_box_qp_work_length(n) + n * 5 + n * n + n * m + m * 2; }
size_t mir_least_squares_iwork_length(size_t m, size_t n)
{
    (void)m;
    const size_t a = mir_box_qp_iwork_length(n);
    return a > n ? a : n;
}

const char* mir_least_squares_status_string(int st)
{
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
    return nullptr;
}

void mir_least_squares_init_d(mir_least_squares_settings_d* s)  { settings_init<mir_least_squares_settings_d, double>(s); }
void mir_least_squares_init_s(mir_least_squares_settings_s* s)  { settings_init<mir_least_squares_settings_s, float>(s); }
void mir_least_squares_reset_d(mir_least_squares_settings_d* s) { settings_init<mir_least_squares_settings_d, double>(s); }
void mir_least_squares_reset_s(mir_least_squares_settings_s* s) { settings_init<mir_least_squares_settings_s, float>(s); }

}  // extern "C"
