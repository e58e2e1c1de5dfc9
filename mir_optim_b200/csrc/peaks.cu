// peaks.cu -- roofline denominators that MEASURED_PEAKS.json does not carry: FP64 FMA (DFMA),
// FP64 tensor (mma.sync m8n8k4 f64 -> DMMA.8x8x4 on sm_100a; tcgen05 has no f64 kind) and FP32
// FMA pipe throughput, measured on the device the caller is about to be judged on.
#include "runtime.cuh"

namespace mirb200 {

template <class T, int ILP>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters, T a, T b)
{
    T acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == (T)123456789) out[0] = s;     // never true; keeps the loop alive
}

template <int ILP>
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a, double b)
{
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    if (s == 123456789.0) out[0] = s;
}

}  // namespace mirb200

using namespace mirb200;

extern "C" {

// kind: 0 = FP64 FMA, 1 = FP64 tensor (DMMA m8n8k4), 2 = FP32 FMA.  Returns TFLOP/s (FMA = 2 flops),
// best of `reps` timed launches after one warm-up, or a negative mir_b200_error.
double mir_b200_measure_peak_tflops(int kind, int reps)
{
    clear_error();
    if (require_device(-1)) return -(double)MIR_B200_ENODEVICE;
    const int sms = sm_count();
    const int blocks = sms * 8, threads = 256, iters = 4096;
    constexpr int ILP = 8;
    double* out = nullptr;
    if (cudaMalloc(&out, 64) != cudaSuccess) return -(double)MIR_B200_ECUDA;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int r = 0; r < reps + 1; ++r) {
        cudaEventRecord(e0);
        if (kind == 0) fma_peak_kernel<double, ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        else if (kind == 1) dmma_peak_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        else fma_peak_kernel<float, ILP><<<blocks, threads>>>((float*)out, iters, 1.0000001f, 1e-9f);
        count_launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        double flops;
        if (kind == 1) flops = (double)blocks * (threads / 32) * iters * ILP * (2.0 * 8 * 8 * 4);
        else flops = (double)blocks * threads * iters * ILP * 2.0;
        if (r > 0 && ms > 0) { double tf = flops / (ms * 1e-3) / 1e12; if (tf > best) best = tf; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    if (check_cuda(cudaGetLastError(), "peak probe")) return -(double)MIR_B200_ECUDA;
    return best;
}

}  // extern "C"
