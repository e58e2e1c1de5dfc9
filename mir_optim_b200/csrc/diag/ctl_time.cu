// times large_ctl_mid_kernel<double> alone on a synthetic n = 128 control block (diagnostic harness, not shipped)
#define MIRB200_PHASE_CLOCK 1
#include "../lm_large.cuh"
#include <cstdio>
#include <vector>
#include <random>
#include <cstring>
#include <cfloat>
#include <cmath>
using namespace mirb200;
namespace mirb200 { void set_error(const char*) {} }
int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 128;
    using Ctl = LargeCtl<double>;
    std::vector<char> hb(sizeof(Ctl)); Ctl* h = (Ctl*)hb.data(); memset(h, 0, sizeof(Ctl));
    std::mt19937_64 rng(7); std::normal_distribution<double> N01;
    const int m = 1024, np = n * (n + 1) / 2;
    std::vector<double> J((size_t)m * n);
    for (auto& v : J) v = N01(rng);
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) J[(size_t)i * n + j] *= (1.0 + 0.5 * (j % 7));
    for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) { double s = 0; for (int k = 0; k < m; ++k) s += J[(size_t)k * n + i] * J[(size_t)k * n + j]; h->packed[i * (i + 1) / 2 + j] = s; }
    for (int i = 0; i < n; ++i) { h->packed[np + i] = N01(rng) * 10; h->x[i] = 1.0 + 0.01 * i; h->xt[i] = h->x[i]; h->l[i] = -1e300; h->u[i] = 1e300; }
    h->n = n; h->ldj = n; h->maxAge = 3; h->hasG = 1; h->tailShortcut = 1;
    h->st.maxIterations = 1000; h->st.absTolerance = DBL_EPSILON; h->st.gradTolerance = DBL_EPSILON; h->st.maxGoodResidual = DBL_EPSILON * DBL_EPSILON;
    h->st.maxStep = sqrt(DBL_MAX) / 16; h->st.maxLambda = DBL_MAX / 16; h->st.minLambda = DBL_MIN * 16; h->st.minStepQuality = 0.1; h->st.goodStepQuality = 0.5;
    h->st.lambdaIncrease = 2; h->st.lambdaDecrease = 1 / (2 * 1.618033988749895); h->st.jacobianEpsilon = ldexp(1.0, -26);
    h->st.qpSettings.relTolerance = 0; h->st.qpSettings.absTolerance = DBL_EPSILON; h->st.qpSettings.maxIterations = 0;
    h->lambda = 1e-3 * h->packed[0]; h->mu = 1; h->residual = 1.0; h->jacMode = JAC_FRESH; h->status = mir_ls_maxIterations;
    Ctl* d; cudaMalloc(&d, sizeof(Ctl)); cudaMemcpy(d, h, sizeof(Ctl), cudaMemcpyHostToDevice);
    const size_t smem = ((CtaQPScratch<double>::bytes(n) + 15) & ~(size_t)15) + sizeof(double) * (4 * (size_t)n + (size_t)np);
    cudaFuncSetAttribute(large_ctl_mid_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        const int iters = rep == 2 ? 50 : 1;
        cudaEventRecord(e0);
        for (int i = 0; i < iters; ++i) large_ctl_mid_kernel<double><<<1, LARGE_CTL_THREADS, smem, 0>>>(d);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("n=%d rep %d: %.2f us per launch (%s)\n", n, rep, ms * 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
        if (rep == 1) {
            long long acc[64]; int cnt[64];
            cudaMemcpyFromSymbol(acc, g_phase_acc, sizeof(acc)); cudaMemcpyFromSymbol(cnt, g_phase_cnt, sizeof(cnt));
            long long tot = 0;
            for (int a = 0; a < 64; ++a) if (cnt[a]) { printf("  phase ending at %2d : %8lld cycles in %3d intervals (%.0f each)\n", a, acc[a], cnt[a], (double)acc[a] / cnt[a]); tot += acc[a]; }
            printf("  total %lld cycles\n", tot);
        }
    }
    cudaMemcpy(h, d, sizeof(Ctl), cudaMemcpyDeviceToHost);
    {
        // ctl_post: accepted and rejected passes (state restored before every launch)
        std::vector<char> sb(sizeof(Ctl)); memcpy(sb.data(), h, sizeof(Ctl));
        for (int mode = 0; mode < 2; ++mode) {
            float best = 1e9f; Ctl* hs = (Ctl*)sb.data();
            hs->rr = mode == 0 ? hs->residual * 0.9 : hs->residual * 1.1; hs->skipRest = 0; hs->initPhase = 0; hs->done = 0; hs->age = 0; hs->needJacobian = 0;
            for (int rep = 0; rep < 6; ++rep) {
                cudaMemcpy(d, hs, sizeof(Ctl), cudaMemcpyHostToDevice); cudaDeviceSynchronize();
                cudaEventRecord(e0); large_ctl_post_kernel<double><<<1, LARGE_CTL_THREADS, 0, 0>>>(d); cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            Ctl* ho = (Ctl*)malloc(sizeof(Ctl)); cudaMemcpy(ho, d, sizeof(Ctl), cudaMemcpyDeviceToHost);
            double cj = 0; for (int i = 0; i < n; ++i) cj += ho->Jy[i] * (i + 1);
            printf("ctl_post %s: %.2f us (launch included); lambda %.17g iterations %u jacMode %d Jy checksum %.17g\n", mode == 0 ? "accepted" : "rejected", best * 1e3, ho->lambda, ho->iterations, ho->jacMode, cj);
            free(ho);
        }
    }
    double cs = 0; for (int i = 0; i < n; ++i) cs += h->xt[i] * (i + 1);
    printf("status %d done %d qpSolves %llu doEval %d checksum %.17g dX0 %.17g\n", h->status, h->done, h->qpSolves, h->doEval, cs, h->dX[0]);
    return 0;
}
