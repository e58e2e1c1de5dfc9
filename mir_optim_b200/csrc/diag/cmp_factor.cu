// compares the blocked factor / warp solve with the column loop, element by element (diagnostic harness)
#include "../boxqp_cta.cuh"
#include <cstdio>
#include <vector>
#include <random>
#include <cstring>
using namespace mirb200;
namespace mirb200 { void set_error(const char*) {} }
constexpr int NT = 256;
__global__ void __launch_bounds__(NT) k_cmp(int s, const double* A, const double* b, double* Fblk, double* Fcol, double* dblk, double* dcol, double* xblk, double* xcol)
{
    extern __shared__ __align__(16) unsigned char smem[];
    CtaQPScratch<double> w; w.carve(smem, s);
    const int tid = threadIdx.x, ldf = w.ldf;
    auto Ag = [&](int i, int k) -> double { return A[i * s + k]; };
    // blocked
    cta_ldl_factor_blocked<double, NT>(s, Ag, w.F, ldf, w.dinv, w.blk);
    for (int e = tid; e < s * s; e += NT) { const int i = e / s, k = e % s; if (k <= i) Fblk[e] = w.F[blk_row(i, ldf) + k]; }
    for (int i = tid; i < s; i += NT) { dblk[i] = w.dinv[i]; w.sx[i] = b[i]; }
    cta_ldl_solve_blocked<double, NT>(s, w.F, ldf, w.dinv, w.sx);
    for (int i = tid; i < s; i += NT) xblk[i] = w.sx[i];
    __syncthreads();
    // column loop
    for (int e = tid; e < s * s; e += NT) { const int i = e / s, k = e % s; if (k <= i) w.F[i * ldf + k] = Ag(i, k); }
    __syncthreads();
    constexpr int TK = 8, TI = NT / TK;
    const int tx = tid % TK, ty = tid / TK;
    for (int j = 0; j < s; ++j) {
        const double d = w.F[j * ldf + j];
        const double inv = 1.0 / d;
        for (int i = j + 1 + ty; i < s; i += TI) {
            const double li = w.F[i * ldf + j] * inv;
            for (int k = j + 1 + tx; k <= i; k += TK) w.F[i * ldf + k] = fnma(li, w.F[k * ldf + j], w.F[i * ldf + k]);
        }
        if (tid == 0) w.dinv[j] = inv;
        __syncthreads();
    }
    for (int e = tid; e < s * s; e += NT) { const int i = e / s, k = e % s; if (k <= i) Fcol[e] = w.F[i * ldf + k]; }
    for (int i = tid; i < s; i += NT) { dcol[i] = w.dinv[i]; w.sx[i] = b[i]; }
    cta_ldl_solve<double, NT>(s, w.F, ldf, w.dinv, w.sx);
    for (int i = tid; i < s; i += NT) xcol[i] = w.sx[i];
}
int main(int argc, char** argv)
{
    for (int n : {64, 100, 128}) {
        std::mt19937_64 rng(5); std::normal_distribution<double> N01;
        const int m = 4 * n;
        std::vector<double> J((size_t)m * n), A((size_t)n * n), b(n);
        for (auto& v : J) v = N01(rng);
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < m; ++k) s += J[(size_t)k * n + i] * J[(size_t)k * n + j]; A[i * n + j] = s; }
        for (auto& v : b) v = N01(rng);
        double *dA, *db, *out; cudaMalloc(&dA, 8 * n * n); cudaMalloc(&db, 8 * n); cudaMalloc(&out, 8 * (2 * n * n + 4 * n)); cudaMemset(out, 0, 8 * (2 * n * n + 4 * n));
        cudaMemcpy(dA, A.data(), 8 * n * n, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), 8 * n, cudaMemcpyHostToDevice);
        const size_t smem = CtaQPScratch<double>::bytes(n) + 16;
        cudaFuncSetAttribute(k_cmp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_cmp<<<1, NT, smem>>>(n, dA, db, out, out + n * n, out + 2 * n * n, out + 2 * n * n + n, out + 2 * n * n + 2 * n, out + 2 * n * n + 3 * n);
        std::vector<double> h(2 * n * n + 4 * n); cudaMemcpy(h.data(), out, 8 * h.size(), cudaMemcpyDeviceToHost);
        printf("n=%d: %s\n", n, cudaGetErrorString(cudaGetLastError()));
        int badF = 0, firstI = -1, firstK = -1;
        for (int i = 0; i < n; ++i) for (int k = 0; k <= i; ++k) if (memcmp(&h[i * n + k], &h[n * n + i * n + k], 8)) { if (!badF) { firstI = i; firstK = k; } ++badF; }
        int badD = 0, badX = 0, fx = -1;
        for (int i = 0; i < n; ++i) { if (memcmp(&h[2 * n * n + i], &h[2 * n * n + n + i], 8)) ++badD; if (memcmp(&h[2 * n * n + 2 * n + i], &h[2 * n * n + 3 * n + i], 8)) { if (fx < 0) fx = i; ++badX; } }
        printf("  factor entries differing: %d (first at %d,%d)  dinv differing: %d  solve entries differing: %d (first %d)\n", badF, firstI, firstK, badD, badX, fx);
        if (n == 100) { FILE* f = fopen("cmp_n100.bin", "wb"); fwrite(h.data(), 8, h.size(), f); fwrite(b.data(), 8, n, f); fclose(f); }
        cudaFree(dA); cudaFree(db); cudaFree(out);
    }
    return 0;
}
