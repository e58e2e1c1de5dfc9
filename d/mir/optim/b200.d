/++
D (BetterC) bindings of the B200 engine's C ABI (include/mir_optim_b200.h).

How it drops in (see INTEGRATION.md): `libmir_optim_b200.so` exports, unmangled, every `extern(C)` symbol that
`mir.optim.least_squares` / `mir.optim.boxcqp` define (least_squares.d:637-799, boxcqp.d:31-51) with identical
signatures and POD layouts, so mir-optim's own D wrappers (`optimize`, `optimizeLeastSquares`, least_squares.d:165-215,
459-519) keep working when the library is linked instead of the D/LAPACK object code.  This module only adds the
declarations of the NEW entry points (batched, device resident, row sharded) over the reference's own types.

NOT compiled in this repository's CI: the build image has no D compiler.  Layouts are enforced from the C++ side
(static_asserts in mir_optim_b200/csrc/shim.cpp against SURVEY appendix C).
+/
module mir.optim.b200;

import mir.optim.least_squares : LeastSquaresSettings, LeastSquaresResult, LeastSquaresStatus;
import mir.optim.boxcqp : BoxQPSettings, BoxQPStatus;
import mir.ndslice.slice : Slice, Contiguous;

extern(C) @system nothrow @nogc:

/// mir_b200_error
enum MirB200Error : int { ok = 0, noDevice = 1, invalid = 2, unsupported = 3, cuda = 4, nccl = 5 }

/// Built-in device residual models ("device functors"); see mir_model_id in the header.
enum MirModel : uint
{
    linear2 = 0, rosenbrock = 1, expDecay2 = 2, expTau3 = 3, sqrtCircle = 4,
    expDecay3 = 5, gauss4 = 6, sumExp = 7, gaussMix = 8,
    spline = 9,          /// fitSpline's residual (fit_splie.d:58-80): aux = knots, param = lambda
    userBase = 0x1000,   /// ids returned by mir_b200_model_compile
}

enum uint mirModelFdJacobian = 1;      /// g == null semantics (least_squares.d:1016-1050)
enum uint mirModelGridPerProblem = 2;  /// abscissa has batch*m entries
enum uint mirModelNoTailShortcut = 4;  /// verification only
enum uint mirModelAuxPerProblem = 8;   /// aux has batch*n entries
enum uint mirModelWarmStart = 16;      /// results[b].lambda on entry is the initial damping (least_squares.d:141-142, 966)

/// mir_model_desc
struct MirModelDesc
{
    uint model;
    uint flags;
    const(void)* t;   /// abscissa T[m] or T[batch*m]
    const(void)* y;   /// observations T[batch*m]
    const(void)* aux; /// model constants T[n] or T[batch*n] (the knots of MirModel.spline); else null
    double param = 0; /// model constant (lambda of MirModel.spline)
}

/// mir_batch_stats: device-side work counters summed over a batch
struct MirBatchStats
{
    ulong problems, passes, accepted, freshJacobians, broydenUpdates, modelEvals, qpSolves, qpIterations;
}

const(char)* mir_b200_last_error();
ulong mir_b200_kernel_launches();
int mir_b200_device_count();
const(char)* mir_b200_version();
double mir_b200_measure_peak_tflops(int kind, int reps);

/// Sentinels for `mir_optimize_least_squares_{d,s}`: pass as f / g with fContext = MirModelDesc* (host arrays).
void mir_b200_device_model_d(scope void* context, size_t m, size_t n, const(double)* x, double* y) pure;
void mir_b200_device_model_jac_d(scope void* context, size_t m, size_t n, const(double)* x, double* J) pure; /// ditto
void mir_b200_device_model_s(scope void* context, size_t m, size_t n, const(float)* x, float* y) pure;      /// ditto
void mir_b200_device_model_jac_s(scope void* context, size_t m, size_t n, const(float)* x, float* J) pure;  /// ditto

/// `batch` independent problems, each with the semantics of one mir_optimize_least_squares_{d,s} call. Host pointers.
int mir_optimize_least_squares_batched_d(scope const LeastSquaresSettings!double* settings, scope const MirModelDesc* model,
    size_t batch, size_t m, size_t n, double* x, const(double)* l, const(double)* u, size_t boundStride,
    LeastSquaresResult!double* results, MirBatchStats* stats, int device);
/// ditto
int mir_optimize_least_squares_batched_s(scope const LeastSquaresSettings!float* settings, scope const MirModelDesc* model,
    size_t batch, size_t m, size_t n, float* x, const(float)* l, const(float)* u, size_t boundStride,
    LeastSquaresResult!float* results, MirBatchStats* stats, int device);
/// Device-pointer forms, asynchronous on `cudaStream`.
int mir_optimize_least_squares_batched_dev_d(scope const LeastSquaresSettings!double* settings, scope const MirModelDesc* model,
    size_t batch, size_t m, size_t n, double* x, const(double)* l, const(double)* u, size_t boundStride,
    LeastSquaresResult!double* results, MirBatchStats* stats, void* cudaStream);
/// ditto
int mir_optimize_least_squares_batched_dev_s(scope const LeastSquaresSettings!float* settings, scope const MirModelDesc* model,
    size_t batch, size_t m, size_t n, float* x, const(float)* l, const(float)* u, size_t boundStride,
    LeastSquaresResult!float* results, MirBatchStats* stats, void* cudaStream);

/// fitSpline (fit_splie.d:26-85) on the GPU: values = the fitted spline values at the knots `x`.
int mir_fit_spline_d(scope const LeastSquaresSettings!double* settings, size_t points, const(double)* pointsX, const(double)* pointsY,
    size_t n, const(double)* x, const(double)* l, const(double)* u, double lambda, double* values, LeastSquaresResult!double* result);
int mir_fit_spline_s(scope const LeastSquaresSettings!float* settings, size_t points, const(float)* pointsX, const(float)* pointsY,
    size_t n, const(float)* x, const(float)* l, const(float)* u, float lambda, float* values, LeastSquaresResult!float* result); /// ditto
int mir_fit_spline_batched_d(scope const LeastSquaresSettings!double* settings, size_t batch, size_t points, const(double)* pointsX,
    const(double)* pointsY, size_t n, const(double)* x, const(double)* l, const(double)* u, double lambda, uint flags,
    double* values, LeastSquaresResult!double* results, int device); /// ditto
int mir_fit_spline_batched_s(scope const LeastSquaresSettings!float* settings, size_t batch, size_t points, const(float)* pointsX,
    const(float)* pointsY, size_t n, const(float)* x, const(float)* l, const(float)* u, float lambda, uint flags,
    float* values, LeastSquaresResult!float* results, int device); /// ditto

/// Residual models written in CUDA C++ (`template <class REAL> struct UserModel {...}`), compiled at run time (NVRTC).
int mir_b200_model_compile(const(char)* source, uint* modelId);
int mir_b200_model_release(uint modelId); /// ditto

/// solveBoxQP (boxcqp.d:85-102) on the GPU; returns BoxQPStatus or -MirB200Error.
int mir_solve_box_qp_d(scope const BoxQPSettings!double* settings, size_t n, const(double)* P, const(double)* q,
    const(double)* l, const(double)* u, double* x);
int mir_solve_box_qp_s(scope const BoxQPSettings!float* settings, size_t n, const(float)* P, const(float)* q,
    const(float)* l, const(float)* u, float* x); /// ditto
int mir_solve_box_qp_batched_d(scope const BoxQPSettings!double* settings, size_t batch, size_t n, const(double)* P,
    const(double)* q, const(double)* l, const(double)* u, double* x, int* status, uint* qpIterations, int device); /// ditto
int mir_solve_box_qp_batched_s(scope const BoxQPSettings!float* settings, size_t batch, size_t n, const(float)* P,
    const(float)* q, const(float)* l, const(float)* u, float* x, int* status, uint* qpIterations, int device); /// ditto

/// One large problem, rows sharded over the ranks of an NCCL communicator (t / y of `model` are DEVICE pointers to this rank's rows).
int mir_optimize_least_squares_sharded_d(scope const LeastSquaresSettings!double* settings, scope const MirModelDesc* model,
    size_t mLocal, size_t n, double* x, const(double)* l, const(double)* u, void* ncclComm, void* cudaStream,
    LeastSquaresResult!double* result, MirBatchStats* stats);
int mir_b200_nccl_unique_id(void* id128);                                            /// NCCL bootstrap
int mir_b200_nccl_comm_init(void** comm, int nranks, const(void)* id128, int rank);  /// ditto
int mir_b200_nccl_comm_destroy(void* comm);                                          /// ditto

extern(D):

/++
Batched counterpart of `optimizeLeastSquares` for the built-in device models: `x` is `batch x n` (in/out), `y` is
`batch x m`, `t` has `m` entries shared by all problems, `l`/`u` have `n` entries shared by all problems.
Returns a MirB200Error; per-problem statuses are in `results` (same codes as LeastSquaresStatus).
+/
MirB200Error optimizeLeastSquaresBatched(T)(
    scope const ref LeastSquaresSettings!T settings, MirModel model, bool finiteDifferences,
    Slice!(T*, 2) x, Slice!(const(T)*) l, Slice!(const(T)*) u,
    Slice!(const(T)*) t, Slice!(const(T)*, 2) y,
    Slice!(LeastSquaresResult!T*) results, MirBatchStats* stats = null, int device = -1) @trusted nothrow @nogc
    if (is(T == double) || is(T == float))
{
    assert(x.length!0 == y.length!0 && x.length!0 == results.length && l.length == x.length!1 && u.length == x.length!1);
    auto desc = MirModelDesc(model, finiteDifferences ? mirModelFdJacobian : 0, t.ptr, y.ptr);
    static if (is(T == double))
        alias fn = mir_optimize_least_squares_batched_d;
    else
        alias fn = mir_optimize_least_squares_batched_s;
    return cast(MirB200Error) fn(&settings, &desc, x.length!0, y.length!1, x.length!1, x.ptr, l.ptr, u.ptr, 0, results.ptr, stats, device);
}

/// Batched counterpart of `solveBoxQP` (boxcqp.d:85-102): P is `batch x n x n`, q/l/u/x are `batch x n`.
MirB200Error solveBoxQPBatched(T)(
    scope const ref BoxQPSettings!T settings, Slice!(const(T)*, 3) P, Slice!(const(T)*, 2) q,
    Slice!(const(T)*, 2) l, Slice!(const(T)*, 2) u, Slice!(T*, 2) x, Slice!(int*) status, int device = -1) @trusted nothrow @nogc
    if (is(T == double) || is(T == float))
{
    assert(P.length!1 == P.length!2 && q.length!1 == P.length!1 && status.length == P.length!0);
    static if (is(T == double))
        alias fn = mir_solve_box_qp_batched_d;
    else
        alias fn = mir_solve_box_qp_batched_s;
    return cast(MirB200Error) fn(&settings, P.length!0, P.length!1, P.ptr, q.ptr, l.ptr, u.ptr, x.ptr, status.ptr, null, device);
}
