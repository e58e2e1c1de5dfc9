"""CUDA sources of user-defined residual models used by the NVRTC tests (and by __graft_entry__ / INTEGRATION.md as examples)."""

# r_i = p0 exp(-p1 t_i) + p2 - y_i with the library's own explicitly rounded operations and reproducible exp: bit-identical
# to the built-in MIR_MODEL_EXPDECAY3 functor, so the two can be compared exactly
EXPDECAY3_REPRO = r'''
template <class REAL> struct UserModel {
    static constexpr bool kAnalytic = true;
    __device__ static REAL residual(const REAL* p, int n, int m, int row, REAL t, REAL y, const REAL* aux, REAL param) {
        using namespace mirb200;
        return sub_rn(add_rn(mul_rn(p[0], exp_repro_tp(mul_rn(-p[1], t))), p[2]), y);
    }
    __device__ static void jacobian(const REAL* p, int n, int m, int row, REAL t, REAL y, const REAL* aux, REAL param, REAL* J) {
        using namespace mirb200;
        const REAL e = exp_repro_tp(mul_rn(-p[1], t));
        J[0] = e; J[1] = mul_rn(-mul_rn(p[0], t), e); J[2] = (REAL)1;
    }
};
'''

# a model the library does not ship: logistic growth curve, r_i = K / (1 + exp(-r (t_i - t0))) - y_i, p = (K, r, t0);
# no Jacobian (finite differences), plain CUDA math
LOGISTIC = r'''
template <class REAL> struct UserModel {
    static constexpr bool kAnalytic = false;
    __device__ static REAL residual(const REAL* p, int n, int m, int row, REAL t, REAL y, const REAL* aux, REAL param) {
        return p[0] / ((REAL)1 + exp(-p[1] * (t - p[2]))) - y;
    }
};
'''

# uses aux (per-parameter weights) and param (ridge weight): rows 0..m-n-1 are a polynomial fit, the last n rows are
# sqrt(param) * aux[k] * p[k] (Tikhonov rows) -- exercises aux / param / row-dependent residuals
RIDGE_POLY = r'''
template <class REAL> struct UserModel {
    static constexpr bool kAnalytic = true;
    __device__ static REAL residual(const REAL* p, int n, int m, int row, REAL t, REAL y, const REAL* aux, REAL param) {
        if (row >= m - n) { const int k = row - (m - n); return sqrt(param) * aux[k] * p[k]; }
        REAL acc = p[n - 1];
        for (int k = n - 2; k >= 0; --k) acc = acc * t + p[k];
        return acc - y;
    }
    __device__ static void jacobian(const REAL* p, int n, int m, int row, REAL t, REAL y, const REAL* aux, REAL param, REAL* J) {
        if (row >= m - n) { for (int k = 0; k < n; ++k) J[k] = 0; const int k = row - (m - n); J[k] = sqrt(param) * aux[k]; return; }
        REAL pw = 1;
        for (int k = 0; k < n; ++k) { J[k] = pw; pw *= t; }
    }
};
'''
