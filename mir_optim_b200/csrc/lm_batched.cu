// lm_batched.cu -- C ABI of the batched LM entry points (header part 2): argument checks,
// staging of host buffers, stream-ordered scratch, launch of the warp-per-problem kernel.
#include <cstdlib>
#include <cstring>
#include <string>

#include "lm_small_launch.cuh"
#include "runtime.cuh"

namespace mirb200 {

template <class T>
static int batched_dev(const typename Num<T>::Settings* settings, const mir_model_desc* model, size_t batch, size_t m, size_t n,
                       T* x, const T* l, const T* u, size_t bound_stride, typename Num<T>::Result* results,
                       mir_batch_stats* stats, cudaStream_t stream, const unsigned int* ready = nullptr)
{
    clear_error();
    if (!settings || !model || (batch && (!x || !l || !u || !results))) { set_error("mir_optim_b200: null argument"); return MIR_B200_EINVAL; }
    if (n == 0) { set_error("mir_optim_b200: n must be > 0 for the batched entry point"); return MIR_B200_EINVAL; }
    if (bound_stride != 0 && bound_stride < n) { set_error("mir_optim_b200: bound_stride must be 0 or >= n"); return MIR_B200_EINVAL; }
    if (batch == 0) return MIR_B200_OK;
    int rc = require_device(-1);
    if (rc) return rc;

    if (batch >= 0xffffffffull) { set_error("mir_optim_b200: at most 2^32 - 2 problems per launch"); return MIR_B200_EINVAL; }
    unsigned int* counter = nullptr;
    MIRB200_CUDA(cudaMallocAsync((void**)&counter, sizeof(unsigned int), stream));
    MIRB200_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));

    SmallBatchArgs a;
    a.t = model->t; a.y = model->y; a.x = x; a.l = l; a.u = u; a.results = results;
    a.counter = counter; a.ready = ready; a.stats = stats; a.batch = batch; a.m = (unsigned)m;
    a.bound_stride = (unsigned)bound_stride; a.flags = model->flags;
    rc = launch_small_model<T>(model->model, n, *settings, a, stream);
    cudaFreeAsync(counter, stream);
    return rc;
}

template <class T>
static int batched_host(const typename Num<T>::Settings* settings, const mir_model_desc* model, size_t batch, size_t m, size_t n,
                        T* x, const T* l, const T* u, size_t bound_stride, typename Num<T>::Result* results,
                        mir_batch_stats* stats, int device)
{
    using Result = typename Num<T>::Result;
    clear_error();
    if (!settings || !model || (batch && (!x || !l || !u || !results))) { set_error("mir_optim_b200: null argument"); return MIR_B200_EINVAL; }
    int rc = require_device(device);
    if (rc) return rc;
    if (batch == 0) return MIR_B200_OK;

    const bool per = (model->flags & MIR_MODEL_GRID_PER_PROBLEM) != 0;
    const size_t tBytes = model->t ? sizeof(T) * (per ? batch * m : m) : 0;
    const size_t yBytes = model->y ? sizeof(T) * batch * m : 0;
    const size_t xBytes = sizeof(T) * batch * n;
    const size_t bBytes = sizeof(T) * (bound_stride ? batch * bound_stride : n);
    const size_t rBytes = sizeof(Result) * batch;

    // Pipeline.  ONE kernel launch for the whole batch (chunked launches lose ~15 % to queue drain at this grid size),
    // enqueued BEFORE its inputs: the per-problem inputs follow on a copy stream in chunks of 65536 problems, each
    // followed by a 4-byte watermark copy; a thread that pulls problem i from the queue waits until the watermark has
    // passed i (wait_staged).  The copy engine runs ~3x ahead of the solve, so only the first chunk's transfer is exposed.
    // MIRB200_NO_STAGING=1: plain copy -> kernel -> copy on one stream (for tools that serialise kernels against copies)
    static const bool noStaging = [] { const char* e = std::getenv("MIRB200_NO_STAGING"); return e && *e == '1'; }();
    const size_t chunk = 65536;
    const size_t nchunks = (batch + chunk - 1) / chunk;
    cudaStream_t cs = nullptr, ps = nullptr;
    cudaEvent_t ev = nullptr;
    unsigned int* h_wm = nullptr;
    rc = MIR_B200_OK;
    auto CK = [&](cudaError_t err, const char* what) { if (rc == MIR_B200_OK) rc = check_cuda(err, what); };
    CK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking), "cudaStreamCreate");
    CK(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking), "cudaStreamCreate");
    CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate");
    h_wm = static_cast<unsigned int*>(pinned_scratch(sizeof(unsigned int) * (nchunks + 1)));
    if (!h_wm && rc == MIR_B200_OK) { set_error("mir_optim_b200: cannot allocate pinned host scratch"); rc = MIR_B200_ECUDA; }
    auto cleanup = [&]() {
        if (cs) cudaStreamDestroy(cs);
        if (ps) cudaStreamDestroy(ps);
        if (ev) cudaEventDestroy(ev);
    };
    if (rc != MIR_B200_OK) { cleanup(); return rc; }

    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t total = align(tBytes) + align(yBytes) + align(xBytes) + 2 * align(bBytes) + align(rBytes) + align(sizeof(mir_batch_stats)) + 256;
    char* base = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&base, total, cs);
    if (e != cudaSuccess) { cleanup(); return check_cuda(e, "cudaMallocAsync(batch buffers)"); }
    char* p = base;
    T* dt = (T*)p; p += align(tBytes);
    T* dy = (T*)p; p += align(yBytes);
    T* dx = (T*)p; p += align(xBytes);
    T* dl = (T*)p; p += align(bBytes);
    T* du = (T*)p; p += align(bBytes);
    Result* dr = (Result*)p; p += align(rBytes);
    mir_batch_stats* ds = (mir_batch_stats*)p; p += align(sizeof(mir_batch_stats));
    unsigned int* d_ready = (unsigned int*)p;

    // shared inputs, counters and the watermark (0 = nothing staged), then the kernel
    if (tBytes && !per) CK(cudaMemcpyAsync(dt, model->t, tBytes, cudaMemcpyHostToDevice, cs), "H2D t");
    if (!bound_stride) {
        CK(cudaMemcpyAsync(dl, l, bBytes, cudaMemcpyHostToDevice, cs), "H2D l");
        CK(cudaMemcpyAsync(du, u, bBytes, cudaMemcpyHostToDevice, cs), "H2D u");
    }
    if (stats) CK(cudaMemsetAsync(ds, 0, sizeof(mir_batch_stats), cs), "memset stats");
    CK(cudaMemsetAsync(d_ready, 0, sizeof(unsigned int), cs), "memset watermark");
    CK(cudaEventRecord(ev, cs), "cudaEventRecord");
    CK(cudaStreamWaitEvent(ps, ev, 0), "cudaStreamWaitEvent");
    bool launched = false;
    auto launch = [&](const unsigned int* ready) {
        mir_model_desc dm = *model;
        dm.t = tBytes ? dt : nullptr; dm.y = yBytes ? dy : nullptr;
        rc = batched_dev<T>(settings, &dm, batch, m, n, dx, dl, du, bound_stride, dr, stats ? ds : nullptr, cs, ready);
        launched = rc == MIR_B200_OK;
    };
    if (rc == MIR_B200_OK && !noStaging) launch(d_ready);
    // the per-problem inputs, chunk by chunk, behind the running kernel
    size_t c = 0;
    for (size_t lo = 0; lo < batch && rc == MIR_B200_OK; lo += chunk, ++c) {
        const size_t nb = (lo + chunk <= batch) ? chunk : batch - lo;
        if (tBytes && per) CK(cudaMemcpyAsync(dt + lo * m, static_cast<const T*>(model->t) + lo * m, sizeof(T) * nb * m, cudaMemcpyHostToDevice, ps), "H2D t");
        if (yBytes) CK(cudaMemcpyAsync(dy + lo * m, static_cast<const T*>(model->y) + lo * m, sizeof(T) * nb * m, cudaMemcpyHostToDevice, ps), "H2D y");
        CK(cudaMemcpyAsync(dx + lo * n, x + lo * n, sizeof(T) * nb * n, cudaMemcpyHostToDevice, ps), "H2D x");
        if (bound_stride) {
            CK(cudaMemcpyAsync(dl + lo * bound_stride, l + lo * bound_stride, sizeof(T) * nb * bound_stride, cudaMemcpyHostToDevice, ps), "H2D l");
            CK(cudaMemcpyAsync(du + lo * bound_stride, u + lo * bound_stride, sizeof(T) * nb * bound_stride, cudaMemcpyHostToDevice, ps), "H2D u");
        }
        h_wm[c] = (unsigned int)(lo + nb);
        CK(cudaMemcpyAsync(d_ready, &h_wm[c], sizeof(unsigned int), cudaMemcpyHostToDevice, ps), "H2D watermark");
    }
    if (noStaging && rc == MIR_B200_OK) {
        CK(cudaEventRecord(ev, ps), "cudaEventRecord");
        CK(cudaStreamWaitEvent(cs, ev, 0), "cudaStreamWaitEvent");
        if (rc == MIR_B200_OK) launch(nullptr);
    }
    if (launched && !noStaging && rc != MIR_B200_OK) {             // a copy failed behind a running kernel: release the waiters
        h_wm[nchunks] = 0xffffffffu;
        cudaMemcpyAsync(d_ready, &h_wm[nchunks], sizeof(unsigned int), cudaMemcpyHostToDevice, ps);
    }
    if (rc == MIR_B200_OK) {
        CK(cudaMemcpyAsync(x, dx, xBytes, cudaMemcpyDeviceToHost, cs), "D2H x");
        CK(cudaMemcpyAsync(results, dr, rBytes, cudaMemcpyDeviceToHost, cs), "D2H results");
        if (stats) CK(cudaMemcpyAsync(stats, ds, sizeof(mir_batch_stats), cudaMemcpyDeviceToHost, cs), "D2H stats");
    }
    cudaError_t se = cudaStreamSynchronize(ps);
    if (rc == MIR_B200_OK) rc = check_cuda(se, "batched LM input staging");
    cudaFreeAsync(base, cs);
    se = cudaStreamSynchronize(cs);
    if (rc == MIR_B200_OK) rc = check_cuda(se, "batched LM kernel");
    cleanup();
    return rc;
}

// used by the legacy single-problem entry (lm_large.cu) for shapes the register-resident kernel covers
template <class T>
int batched_host_entry(const typename Num<T>::Settings* settings, const mir_model_desc* model, size_t batch, size_t m, size_t n,
                       T* x, const T* l, const T* u, size_t bound_stride, typename Num<T>::Result* results,
                       mir_batch_stats* stats, int device)
{ return batched_host<T>(settings, model, batch, m, n, x, l, u, bound_stride, results, stats, device); }
template int batched_host_entry<double>(const mir_least_squares_settings_d*, const mir_model_desc*, size_t, size_t, size_t, double*, const double*,
                                        const double*, size_t, mir_least_squares_result_d*, mir_batch_stats*, int);
template int batched_host_entry<float>(const mir_least_squares_settings_s*, const mir_model_desc*, size_t, size_t, size_t, float*, const float*,
                                       const float*, size_t, mir_least_squares_result_s*, mir_batch_stats*, int);

}  // namespace mirb200

using namespace mirb200;

extern "C" {

int mir_optimize_least_squares_batched_d(const mir_least_squares_settings_d* s, const mir_model_desc* model, size_t batch, size_t m,
    size_t n, double* x, const double* l, const double* u, size_t bound_stride, mir_least_squares_result_d* results,
    mir_batch_stats* stats, int device)
{ return batched_host<double>(s, model, batch, m, n, x, l, u, bound_stride, results, stats, device); }

int mir_optimize_least_squares_batched_s(const mir_least_squares_settings_s* s, const mir_model_desc* model, size_t batch, size_t m,
    size_t n, float* x, const float* l, const float* u, size_t bound_stride, mir_least_squares_result_s* results,
    mir_batch_stats* stats, int device)
{ return batched_host<float>(s, model, batch, m, n, x, l, u, bound_stride, results, stats, device); }

int mir_optimize_least_squares_batched_dev_d(const mir_least_squares_settings_d* s, const mir_model_desc* model, size_t batch, size_t m,
    size_t n, double* x, const double* l, const double* u, size_t bound_stride, mir_least_squares_result_d* results,
    mir_batch_stats* stats, void* stream)
{ return batched_dev<double>(s, model, batch, m, n, x, l, u, bound_stride, results, stats, (cudaStream_t)stream); }

int mir_optimize_least_squares_batched_dev_s(const mir_least_squares_settings_s* s, const mir_model_desc* model, size_t batch, size_t m,
    size_t n, float* x, const float* l, const float* u, size_t bound_stride, mir_least_squares_result_s* results,
    mir_batch_stats* stats, void* stream)
{ return batched_dev<float>(s, model, batch, m, n, x, l, u, bound_stride, results, stats, (cudaStream_t)stream); }

}  // extern "C"
