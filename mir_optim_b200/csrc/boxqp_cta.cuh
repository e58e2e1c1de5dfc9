// boxqp_cta.cuh -- CTA-cooperative BOXCQP and ?posvx('E','L') restatement for runtime n (<= 128),
// matrices in shared memory.  Used by the batched BoxQP kernel (one CTA per QP, BASELINE
// configs[4]) and by the control kernel of the large single-problem LM path (n = 128).
//
//   cta_posvx  <- LAPACK dposvx/sposvx FACT='E', UPLO='L' as called at boxcqp.d:194-205, 310-321:
//                 ?poequ/?laqsy equilibration decision, factorisation, solve, ?porfs refinement
//                 (<= 5 sweeps driven by the componentwise backward error), un-scaling.
//                 The factorisation is the square-root-free LDL^T form of Cholesky (same pivots
//                 d_j > 0 test as ?potrf's breakdown test); OpenBLAS' blocked potrf cannot be
//                 matched bit for bit anyway, the parity bar is the stated tolerance.
//   cta_boxqp  <- solveBoxQP!T full overload, boxcqp.d:122-379 (unconstrainedSolution = false).
//                 Active-set sub-systems are compacted through the ascending free-index list,
//                 exactly like boxcqp.d:269-305.
#pragma once
#include "boxqp_small.cuh"   // KBN
#include "common.cuh"

namespace mirb200 {

// Phase clocks for the single-CTA control kernel (diagnostic builds only, -DMIRB200_PHASE_CLOCK): thread 0 books the
// cycles between named points in shared memory, the kernel dumps the totals at its end.  Compiled out otherwise.
#ifdef MIRB200_PHASE_CLOCK
__device__ long long g_phase_acc[64];
__device__ int g_phase_cnt[64];
struct PhaseClock { long long last, acc[64]; int cnt[64]; };
__device__ __forceinline__ PhaseClock& phase_clock() { __shared__ PhaseClock pc; return pc; }
__device__ __forceinline__ void phase_stamp(int tag)      // the interval that ends here is booked under `tag`; tag 0 resets
{
    if (threadIdx.x != 0) return;
    PhaseClock& pc = phase_clock();
    const long long now = clock64();
    if (tag == 0) { for (int i = 0; i < 64; ++i) { pc.acc[i] = 0; pc.cnt[i] = 0; } }
    else { pc.acc[tag] += now - pc.last; ++pc.cnt[tag]; }
    pc.last = clock64();
}
__device__ __forceinline__ void phase_dump()
{
    if (threadIdx.x != 0) return;
    PhaseClock& pc = phase_clock();
    for (int i = 0; i < 64; ++i) { g_phase_acc[i] = pc.acc[i]; g_phase_cnt[i] = pc.cnt[i]; }
}
#define MIRB200_PHASE(tag) phase_stamp(tag)
#define MIRB200_PHASE_DUMP() phase_dump()
#else
#define MIRB200_PHASE(tag) do { } while (0)
#define MIRB200_PHASE_DUMP() do { } while (0)
#endif

// Accessor for a packed lower-triangular symmetric matrix plus a diagonal shift (P = A + shift I, the LM step matrix).
// cta_posvx recognises it and walks the packed rows / columns directly in its refinement residual instead of computing
// tri(i, j) and the diagonal test per element.
template <class T> struct PackedLowerShift {
    const T* a; T shift;
    __device__ __forceinline__ T operator()(int i, int j) const { const T v = a[tri(i, j)]; return i == j ? v + shift : v; }
};
template <class A> struct IsPackedLower { static constexpr bool value = false; };
template <class T> struct IsPackedLower<PackedLowerShift<T>> { static constexpr bool value = true; };

template <class T, int NT> __device__ __forceinline__ T cta_reduce_max(T v, T* red)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = t_max(v, __shfl_xor_sync(0xffffffffu, v, off));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    T r = red[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) r = t_max(r, red[w]);
    return r;
}
template <class T, int NT> __device__ __forceinline__ T cta_reduce_min(T v, T* red) { return -cta_reduce_max<T, NT>(-v, red); }
template <int NT> __device__ __forceinline__ bool cta_any(bool p) { return __syncthreads_or(p ? 1 : 0) != 0; }

// Shared-memory scratch of one CTA-level QP solve.  nmax = capacity, ldf = pitch(nmax).
template <class T> struct CtaQPScratch {
    T* F;        // nmax * ldf   factor (lower, LDL^T: unscaled Schur columns)
    T* sc;       // equilibration scale
    T* dinv;     // 1 / d_j
    T* b;        // right-hand side (scaled)
    T* sx;       // sub-system solution
    T* r;        // refinement residual
    T* la;       // multipliers of lower bounds
    T* mu;       // multipliers of upper bounds
    T* red;      // 32 reduction slots
    T* blk;      // blocked factorisation staging (blk_elems)
    int* idx;    // free-index list (ascending)
    signed char* flag;   // -1 lower, 0 free, +1 upper (boxcqp.d:153-158)
    int ldf;
    // Row pitch of F: odd, so that a column walk (fixed column, consecutive rows) touches every bank once; the blocked
    // path adds one element of skew per 16 rows (blk_row) so that the tile stores, whose rows are 8 apart, do too.
    __host__ __device__ static int pitch(int nmax) { return (nmax + 8) | 1; }
    __host__ __device__ static size_t bytes(int nmax) {
        return sizeof(T) * ((size_t)nmax * pitch(nmax) + 8 + 7 * (size_t)nmax + 32 + blk_elems(nmax)) + sizeof(int) * nmax + ((nmax + 15) & ~15);
    }
    // blocked factorisation staging: panel 64 x (tile rows | 1), two diagonal tiles 8 x 9 + 8 reciprocal pivots each
    __host__ __device__ static size_t blk_elems(int nmax) { return (size_t)64 * ((((nmax + 7) >> 3)) | 1) + 2 * 80; }
    __device__ void carve(void* base, int nmax) {
        ldf = pitch(nmax);
        T* p = static_cast<T*>(base);
        F = p; p += (size_t)nmax * ldf + 8;
        sc = p; p += nmax; dinv = p; p += nmax; b = p; p += nmax; sx = p; p += nmax; r = p; p += nmax;
        la = p; p += nmax; mu = p; p += nmax; red = p; p += 32; blk = p; p += blk_elems(nmax);
        idx = reinterpret_cast<int*>(p);
        flag = reinterpret_cast<signed char*>(idx + nmax);
    }
};

// L D L^T solve with warp 0; v in shared memory, overwritten by the solution.  Other warps wait.
template <class T, int NT>
__device__ __forceinline__ void cta_ldl_solve(int s, const T* F, int ldf, const T* dinv, T* v)
{
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int j = 0; j < s; ++j) {                       // forward, unit lower L' = F[i][j] * dinv[j]
            __syncwarp();
            const T t = v[j] * dinv[j];
            for (int i = j + 1 + lane; i < s; i += 32) v[i] = fnma(F[i * ldf + j], t, v[i]);
        }
        __syncwarp();
        for (int i = lane; i < s; i += 32) v[i] *= dinv[i]; // D^-1
        for (int j = s - 1; j > 0; --j) {                   // backward with L'^T
            __syncwarp();
            const T xj = v[j];
            for (int i = lane; i < j; i += 32) v[i] = fnma(F[j * ldf + i] * dinv[i], xj, v[i]);
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// Blocked variants (8 x 8 tiles) for ONE CTA working alone, where latency is all that matters (the control kernel of
// the large-problem path, n = 128).  They perform, for every matrix / vector element, the SAME floating-point
// operations in the SAME order as the column-by-column loops above (an element (i,k) is updated with pivots p = 0, 1,
// ... ascending; the blocking only groups eight pivots between barriers), so the results are bit-identical.
//   factor: the trailing matrix lives in registers, one lower-triangular 8 x 8 tile per thread, loaded straight from
//           the matrix accessor; 2 CTA barriers per eight columns (the owner of the next diagonal tile factors it right
//           after its own trailing update).  The panel is exchanged through shared memory in a layout whose bank is the
//           tile row, the factor is stored with pitch / skew chosen so that no phase has bank conflicts.
//   solve:  one warp, the right-hand side in registers (lane owns four consecutive rows), one shuffle round per four
//           columns and no barrier.
// Phase clocks on B200 (n = 128, cycles per posvx): round-2 first version staging 20.7 k + factor 98 k + 42 k per solve
// (32 CTA barriers and a one-thread diagonal part per direction) -> see DESIGN section 4.2 for the current figures.
// ---------------------------------------------------------------------------------------------------------------
template <int NT> __host__ __device__ constexpr bool cta_blocked_ok(int s) { return ((s + 7) / 8) * ((s + 7) / 8 + 1) / 2 <= NT && s <= 128; }

__host__ __device__ __forceinline__ int blk_row(int i, int ldf) { return i * ldf + (i >> 4); }

// LDL^T (square-root free) of the matrix Aload(i, k), i >= k; factor (unscaled Schur columns) in F (blk_row layout) and
// the reciprocal pivots in dinv on exit.  Returns 0 or the 1-based index of the first pivot <= 0 (uniform over the CTA).
template <class T, int NT, class ALoad>
__device__ int cta_ldl_factor_blocked(int s, ALoad Aload, T* F, int ldf, T* dinv, T* blk)
{
    const int tid = threadIdx.x;
    const int nb = (s + 7) >> 3;
    const int PT = nb | 1;
    T* Pn = blk;                                   // panel: entry (8 t + r, p) of the current block column at Pn[(p * 8 + r) * PT + t]
    T* Dg2 = blk + (size_t)64 * PT;                // two diagonal tiles (alternating): [k * 9 + p], then the 8 reciprocal pivots at [72 + p]
    __shared__ int s_info;
    // tile -> thread: row-major over the lower-triangular tile grid.  With 16 tile rows there are 136 tiles: the eight
    // column-0 tiles of rows 8..15 (which only take part in the very first panel) go to the fifth warp, so that from
    // then on every scheduler partition of the SM runs exactly one warp of tile owners.
    int ti = nb, tj = 0;
    {
        int e = tid;
        if (nb == 16) {
            if (tid >= 136) e = -1;
            else if (tid >= 128) { e = -1; ti = 8 + (tid - 128); tj = 0; }
            else {
#pragma unroll
                for (int k = 0; k < 8; ++k) { const int x = (8 + k) * (9 + k) / 2; if (e >= x) ++e; }
            }
        }
        if (e >= 0) {
            ti = 0;
            while ((ti + 1) * (ti + 2) / 2 <= e) ++ti;
            tj = e - ti * (ti + 1) / 2;
        }
    }
    const bool owner = ti < nb;                    // this thread owns tile (ti, tj), ti >= tj
    T a[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int i = 8 * ti + r, k = 8 * tj + c;
            a[r][c] = (i == k) ? (T)1 : (T)0;                                                        // identity padding
            if (owner && i < s && k <= i) a[r][c] = Aload(i, k);
        }
    if (tid == 0) s_info = 0;
    __syncthreads();

    // the diagonal tile of block column jt: plain right-looking LDL^T in the registers of its owner
    auto factor_diag = [&](int jt) {
        T* Dg = Dg2 + (jt & 1) * 80;
        int info = 0;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const T d = a[p][p];
            if (info == 0 && 8 * jt + p < s && d <= (T)0) info = 8 * jt + p + 1;     // a NaN pivot passes, as in OpenBLAS' potrf
            const T inv = (T)1 / d;
            if (8 * jt + p < s) dinv[8 * jt + p] = inv;
            Dg[72 + p] = inv;
#pragma unroll
            for (int i = p + 1; i < 8; ++i) {
                const T li = a[i][p] * inv;
#pragma unroll
                for (int k = p + 1; k <= i; ++k) a[i][k] = fnma(li, a[k][p], a[i][k]);
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) {
                Dg[r * 9 + c] = a[r][c];
                if (8 * jt + r < s) F[blk_row(8 * jt + r, ldf) + 8 * jt + c] = a[r][c];
            }
        if (info) s_info = info;
    };
    if (owner && ti == 0 && tj == 0) factor_diag(0);

    // Tried and dropped (phase clocks, n = 128): this kernel runs once per launch on one SM, every instruction is first
    // fetched cold (ncu: stall_no_instruction 31 % of the non-barrier samples, 330 KB of SASS), but the smaller-code
    // variants were slower -- a rolled pivot loop in the trailing update and ?posvx out of line (one copy for the two
    // call sites of BOXCQP): factorisation +60 % (exposed shared-memory latency per pivot, the scratch descriptor behind
    // a stack reference); a single inline site for the diagonal tile plus a packed-row tile load: 244 B of spills.
    for (int jt = 0; jt < nb; ++jt) {
        const T* Dg = Dg2 + (jt & 1) * 80;
        __syncthreads();                             // diagonal tile jt is out (and every trailing update of step jt - 1 is done)
        MIRB200_PHASE(20);
        if (s_info) return s_info;
        // the panel below it: columns of the block, left to right
        if (owner && tj == jt && ti > jt) {
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const T inv = Dg[72 + p];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const T li = a[r][p] * inv;
#pragma unroll
                    for (int k = p + 1; k < 8; ++k) a[r][k] = fnma(li, Dg[k * 9 + p], a[r][k]);
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int fr = blk_row(8 * ti + r, ldf) + 8 * jt;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    Pn[(c * 8 + r) * PT + ti] = a[r][c];
                    if (8 * ti + r < s) F[fr + c] = a[r][c];
                }
            }
        }
        __syncthreads();
        MIRB200_PHASE(21);
        // the trailing tiles: eight rank-1 updates from the panel, operands in registers; the owner of the next
        // diagonal tile goes straight on to factor it
        if (owner && tj > jt) {
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const T inv = Dg[72 + p];
                T li[8], ck[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) { li[r] = Pn[(p * 8 + r) * PT + ti] * inv; ck[r] = Pn[(p * 8 + r) * PT + tj]; }
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) a[r][c] = fnma(li[r], ck[c], a[r][c]);
            }
            if (ti == jt + 1 && tj == jt + 1) factor_diag(jt + 1);
        }
        MIRB200_PHASE(22);
    }
    __syncthreads();
    return 0;
}

// L D L^T solve by ONE warp (call with the 32 threads of a warp; v in shared memory, overwritten by the solution).
// Same operations per element and the same order as cta_ldl_solve: forward t_j = v_j dinv_j, v_i -= F_ij t_j with j
// ascending; D^-1; backward v_i -= (F_ji dinv_i) x_j with j descending.  F in the blk_row layout.
template <class T>
__device__ __forceinline__ void warp_ldl_solve(int s, const T* F, int ldf, const T* dinv, T* v)
{
    // Lane L owns rows 4L .. 4L+3 (s <= 128).  Per block of four columns: the owner lane substitutes inside its 4 x 4
    // diagonal block in registers, four values go round by shuffle, the other lanes apply them to their rows -- one
    // shuffle round per four columns on the dependent chain; the factor entries of the next block are loaded a block ahead.
    const int lane = threadIdx.x & 31;
    const int nblk = (s + 3) >> 2;
    T vr[4], di[4];
    int fo[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = 4 * lane + r;
        const bool ok = i < s;
        vr[r] = ok ? v[i] : (T)0; di[r] = ok ? dinv[i] : (T)0; fo[r] = blk_row(ok ? i : 0, ldf);
    }
    MIRB200_PHASE(30);
    // The loop bodies are kept lean on purpose: one warp works alone here, so the solve is bound by instruction issue as
    // much as by the dependent chain (phase clocks: a version with per-element predicates and register copies between
    // the prefetch buffers took 3x the cycles of its arithmetic).  Loads are unpredicated where the value cannot reach a
    // stored result: rows >= s of a lane are never written back, a lane above / below the block ignores its own 4 x 4
    // scratch, and the owner only uses the strictly lower part of its block.
    {   // forward: t_j = v_j dinv_j, v_i -= F_ij t_j, j ascending
        auto loadF = [&](int jb, T (&f)[4][4]) {       // f[r][c] = F(4 lane + r, 4 jb + c)
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) f[r][c] = F[fo[r] + 4 * jb + c];
        };
        auto step = [&](int jb, const T (&f)[4][4]) {
            // the owner's 4 x 4 block (every lane runs it on its own registers, only lane jb's values are used)
            T w[4], t[4];
            w[0] = vr[0];
            t[0] = w[0] * di[0];
            w[1] = fnma(f[1][0], t[0], vr[1]);
            t[1] = w[1] * di[1];
            w[2] = fnma(f[2][1], t[1], fnma(f[2][0], t[0], vr[2]));
            t[2] = w[2] * di[2];
            w[3] = fnma(f[3][2], t[2], fnma(f[3][1], t[1], fnma(f[3][0], t[0], vr[3])));
            t[3] = w[3] * di[3];
            T tb[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) tb[c] = __shfl_sync(0xffffffffu, t[c], jb);
            if (lane == jb) {
#pragma unroll
                for (int r = 0; r < 4; ++r) vr[r] = w[r];
            } else if (lane > jb) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) vr[r] = fnma(f[r][c], tb[c], vr[r]);
            }
        };
        T fa[4][4], fb[4][4];
        loadF(0, fa);
        for (int jb = 0; jb < nblk; jb += 2) {         // two blocks per trip: the prefetch buffers swap roles, nothing is copied
            loadF(jb + 1 < nblk ? jb + 1 : jb, fb);
            step(jb, fa);
            if (jb + 1 >= nblk) break;
            loadF(jb + 2 < nblk ? jb + 2 : jb + 1, fa);
            step(jb + 1, fb);
        }
    }
    MIRB200_PHASE(32);
#pragma unroll
    for (int r = 0; r < 4; ++r) vr[r] = (4 * lane + r < s) ? vr[r] * di[r] : (T)0;        // D^-1 (rows >= s: back to an exact zero)
    {   // backward: v_i -= (F_ji dinv_i) x_j, j descending
        auto loadG = [&](int jb, T (&g)[4][4]) {       // g[c][r] = F(4 jb + c, 4 lane + r) dinv_i; columns >= s (last block) give 0
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int j = 4 * jb + c;
                const bool okc = j < s;
                const int fj = blk_row(okc ? j : 0, ldf) + 4 * lane;
#pragma unroll
                for (int r = 0; r < 4; ++r) g[c][r] = okc ? F[fj + r] * di[r] : (T)0;
            }
        };
        auto step = [&](int jb, const T (&g)[4][4]) {
            T x[4];
            x[3] = vr[3];
            x[2] = fnma(g[3][2], x[3], vr[2]);
            x[1] = fnma(g[2][1], x[2], fnma(g[3][1], x[3], vr[1]));
            x[0] = fnma(g[1][0], x[1], fnma(g[2][0], x[2], fnma(g[3][0], x[3], vr[0])));
            T xb[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) xb[c] = __shfl_sync(0xffffffffu, x[c], jb);
            if (lane == jb) {
#pragma unroll
                for (int r = 0; r < 4; ++r) vr[r] = x[r];
            } else if (lane < jb) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 3; c >= 0; --c) vr[r] = fnma(g[c][r], xb[c], vr[r]);
            }
        };
        T ga[4][4], gb[4][4];
        loadG(nblk - 1, ga);
        for (int jb = nblk - 1; jb >= 0; jb -= 2) {
            loadG(jb > 0 ? jb - 1 : 0, gb);
            step(jb, ga);
            if (jb == 0) break;
            loadG(jb > 1 ? jb - 2 : 0, ga);
            step(jb - 1, gb);
        }
    }
    MIRB200_PHASE(34);
#pragma unroll
    for (int r = 0; r < 4; ++r) { const int i = 4 * lane + r; if (i < s) v[i] = vr[r]; }
}

// CTA wrapper: warp 0 solves, the other warps wait.
template <class T, int NT>
__device__ __forceinline__ void cta_ldl_solve_blocked(int s, const T* F, int ldf, const T* dinv, T* v)
{
    __syncthreads();
    if (threadIdx.x < 32) warp_ldl_solve<T>(s, F, ldf, dinv, v);
    __syncthreads();
}

// A(a, c) for a >= c returns the (unscaled) entry of the s x s system.  b: rhs (overwritten by its
// scaled copy), x: solution.  Returns LAPACK info (0, or k>0 = breakdown at pivot k), uniform over the CTA.
// BLK: use the blocked register-tiled factorisation / solves.  They win when ONE CTA works alone and latency is all
// that matters (the LM control kernel, n = 128: 0.47 -> 0.22 ms); the batched BoxQP kernel, with two CTAs per SM
// competing for issue slots, is faster with the all-threads column loop (measured 1.35 M vs 1.17 M QP/s at n = 64).
template <class T, int NT, bool BLK, class AGet>
__device__ int cta_posvx(int s, AGet A, CtaQPScratch<T>& w, T* b, T* x, int* equed_out = nullptr)
{
    const int tid = threadIdx.x;
    T* F = w.F; const int ldf = w.ldf;

    // ?poequ / ?laqsy
    MIRB200_PHASE(10);
    T mn = Num<T>::inf(), mx = -Num<T>::inf();
    for (int a = tid; a < s; a += NT) { const T d = A(a, a); mn = t_min(mn, d); mx = t_max(mx, d); }
    const T smin = cta_reduce_min<T, NT>(mn, w.red);
    const T amax = cta_reduce_max<T, NT>(mx, w.red);
    bool equil = false;
    if (smin > (T)0) {
        const T scond = sqrt_ni(smin) / sqrt_ni(amax);
        equil = !(scond >= (T)0.1 && amax >= Num<T>::small_() && amax <= Num<T>::large_());
    }
    if (equed_out && tid == 0) *equed_out = equil ? 1 : 0;      // (unit tests only: the ?laqsy decision, LAPACK's EQUED)
    for (int a = tid; a < s; a += NT) {
        const T sa = equil ? (T)1 / sqrt_ni(A(a, a)) : (T)1;
        w.sc[a] = sa;
        b[a] = equil ? sa * b[a] : b[a];
    }
    __syncthreads();
    MIRB200_PHASE(11);
    const bool blocked = BLK && cta_blocked_ok<NT>(s);
    if (blocked) {
        // register-tiled factorisation straight from the accessor (nothing is staged), warp-level solves
        MIRB200_PHASE(12);
        auto As = [&](int a, int c) -> T { return equil ? (w.sc[c] * w.sc[a]) * A(a, c) : A(a, c); };
        const int info = cta_ldl_factor_blocked<T, NT>(s, As, F, ldf, w.dinv, w.blk);
        if (info) return info;
        MIRB200_PHASE(13);
        for (int a = tid; a < s; a += NT) x[a] = b[a];
        cta_ldl_solve_blocked<T, NT>(s, F, ldf, w.dinv, x);
        MIRB200_PHASE(14);
    } else {
        for (int e = tid; e < s * s; e += NT) {
            const int a = e / s, c = e - a * s;
            if (c <= a) F[a * ldf + c] = equil ? (w.sc[c] * w.sc[a]) * A(a, c) : A(a, c);
        }
        __syncthreads();
        // factorisation: right-looking, square-root free.  Column j keeps its unscaled Schur values.
        constexpr int TK = 8, TI = NT / TK;
        const int tx = tid % TK, ty = tid / TK;
        for (int j = 0; j < s; ++j) {
            const T d = F[j * ldf + j];
            if (d <= (T)0) return j + 1;     // uniform (every thread reads the same d).  A NaN pivot passes, as in OpenBLAS' potrf
            const T inv = (T)1 / d;
            for (int i = j + 1 + ty; i < s; i += TI) {
                const T li = F[i * ldf + j] * inv;
                for (int k = j + 1 + tx; k <= i; k += TK) F[i * ldf + k] = fnma(li, F[k * ldf + j], F[i * ldf + k]);
            }
            if (tid == 0) w.dinv[j] = inv;
            __syncthreads();
        }
        for (int a = tid; a < s; a += NT) x[a] = b[a];
        cta_ldl_solve<T, NT>(s, F, ldf, w.dinv, x);
    }

    // ?porfs
    const T eps = Num<T>::lapack_eps();
    const T safe1 = (T)(s + 1) * Num<T>::safmin();
    const T safe2 = safe1 / eps;
    T lstres = (T)3;
    for (int count = 1;; ++count) {
        T q = (T)0;
        for (int a = tid; a < s; a += NT) {
            T ra = b[a], wa = t_abs(b[a]);
            const T sa = w.sc[a];
            auto term = [&](T aij, int c) {                        // one entry of row a: r_a -= a_ac x_c, w_a += |a_ac| |x_c|
                if (equil) aij = (sa * w.sc[c]) * aij;
                const T xc = x[c];
                ra = fnma(aij, xc, ra);
                wa = fma(t_abs(aij), t_abs(xc), wa);
            };
            // row part of the symmetric matrix (c <= a), then the column part (c > a); c ascending throughout
            if constexpr (IsPackedLower<AGet>::value) {
                const T* rowp = A.a + tri(a, 0);
#pragma unroll 4
                for (int c = 0; c < a; ++c) term(rowp[c], c);
                term(rowp[a] + A.shift, a);
                int idx = tri(a + 1, a);
#pragma unroll 4
                for (int c = a + 1; c < s; ++c) { term(A.a[idx], c); idx += c + 1; }
            } else {
#pragma unroll 4
                for (int c = 0; c <= a; ++c) term(A(a, c), c);
#pragma unroll 4
                for (int c = a + 1; c < s; ++c) term(A(c, a), c);
            }
            w.r[a] = ra;
            q = t_max(q, (wa > safe2) ? t_abs(ra) / wa : (t_abs(ra) + safe1) / (wa + safe1));
        }
        const T berr = cta_reduce_max<T, NT>(q, w.red);
        MIRB200_PHASE(15);
        if (berr > eps && (T)2 * berr <= lstres && count <= 5) {
            if (blocked) cta_ldl_solve_blocked<T, NT>(s, F, ldf, w.dinv, w.r);
            else cta_ldl_solve<T, NT>(s, F, ldf, w.dinv, w.r);
            for (int a = tid; a < s; a += NT) x[a] += w.r[a];
            __syncthreads();
            MIRB200_PHASE(16);
            lstres = berr;
            continue;
        }
        break;
    }
    if (equil) for (int a = tid; a < s; a += NT) x[a] *= w.sc[a];
    __syncthreads();
    return 0;
}

// P(i, j) for i >= j: lower triangle of the QP matrix.  q, l, u, x: length n (shared or global).
// Returns mir_box_qp_status (uniform).  iterations: BOXCQP main-loop count, solves: posvx calls.
template <class T, int NT, bool BLK, class PGet>
__device__ int cta_boxqp(const typename Num<T>::QPSettings& st, int n, PGet P, const T* q, const T* l, const T* u, T* x,
                         CtaQPScratch<T>& w, unsigned& iterations, unsigned& solves)
{
    const int tid = threadIdx.x;
    iterations = 0; solves = 0;
    if (n == 0) return mir_qp_solved;                                          // boxcqp.d:162-163

    for (int i = tid; i < n; i += NT) w.b[i] = -q[i];                          // boxcqp.d:191
    __syncthreads();
    ++solves;
    if (cta_posvx<T, NT, BLK>(n, P, w, w.b, x) != 0) return mir_qp_numericError;    // boxcqp.d:194-213

    bool out = false;                                                          // boxcqp.d:216-219
    for (int i = tid; i < n; i += NT) out = out || !(l[i] <= x[i] && x[i] <= u[i]);
    if (!cta_any<NT>(out)) return mir_qp_solved;

    const unsigned maxIterations = st.maxIterations ? st.maxIterations : (unsigned)n * 10u + 100u;   // boxcqp.d:224-226
    for (int i = tid; i < n; i += NT) { w.la[i] = (T)0; w.mu[i] = (T)0; }
    __shared__ int s_free;

    for (unsigned step = 0; step < maxIterations; ++step) {                    // boxcqp.d:234
        ++iterations;
        __syncthreads();
        for (int i = tid; i < n; i += NT) {                                    // boxcqp.d:239-263
            const T xl = x[i] - l[i];
            const T ux = u[i] - x[i];
            if (xl < (T)0 || (xl < st.relTolerance + st.absTolerance * t_abs(l[i]) && w.la[i] >= (T)0)) {
                w.flag[i] = -1; x[i] = l[i]; w.mu[i] = (T)0;
            } else if (ux < (T)0 || (ux < st.relTolerance + st.absTolerance * t_abs(u[i]) && w.mu[i] >= (T)0)) {
                w.flag[i] = 1; x[i] = u[i]; w.la[i] = (T)0;
            } else {
                w.flag[i] = 0; w.mu[i] = (T)0; w.la[i] = (T)0;
            }
        }
        __syncthreads();
        if (tid == 0) {                                                        // ascending free list
            int s = 0;
            for (int i = 0; i < n; ++i) if (w.flag[i] == 0) w.idx[s++] = i;
            s_free = s;
        }
        __syncthreads();
        const int s = s_free;
        if (s == n) break;                                                     // boxcqp.d:265-266 -> maxIterations

        if (s > 0) {
            for (int a = tid; a < s; a += NT) {                                // boxcqp.d:282-305
                const int i = w.idx[a];
                KBN<T> sum(q[i]);
                for (int j = 0; j < n; ++j) {
                    const int f = w.flag[j];
                    if (f) sum.put(mul_rn((i >= j) ? P(i, j) : P(j, i), f < 0 ? l[j] : u[j]));
                }
                w.b[a] = -sum.sum();
            }
            __syncthreads();
            ++solves;
            const int* idx = w.idx;
            auto Asub = [&](int a, int c) -> T { return P(idx[a], idx[c]); };  // idx ascending: a >= c => idx[a] >= idx[c]
            if (cta_posvx<T, NT, BLK>(s, Asub, w, w.b, w.sx) != 0) return mir_qp_numericError;   // boxcqp.d:310-324
            for (int a = tid; a < s; a += NT) x[w.idx[a]] = w.sx[a];           // boxcqp.d:327-329
            __syncthreads();
        }

        for (int i = tid; i < n; i += NT) if (w.flag[i]) {                     // boxcqp.d:333-337
            T d1 = (T)0, d2 = (T)0;
            for (int j = 0; j < i; ++j) d1 += P(i, j) * x[j];
            for (int j = i; j < n; ++j) d2 += P(j, i) * x[j];
            const T val = d1 + d2 + q[i];
            if (w.flag[i] < 0) w.la[i] = val; else w.mu[i] = -val;
        }
        __syncthreads();
        bool again = false;                                                    // boxcqp.d:339-347
        for (int i = tid; i < n; i += NT) {
            const int f = w.flag[i];
            if (f < 0)      again = again || !(w.la[i] >= (T)0);
            else if (f > 0) again = again || !(w.mu[i] >= (T)0);
            else            again = again || !(x[i] >= l[i] && x[i] <= u[i]);
        }
        if (cta_any<NT>(again)) continue;

        for (int i = tid; i < n; i += NT) x[i] = t_max(t_min(x[i], u[i]), l[i]);   // applyBounds, boxcqp.d:349
        __syncthreads();
        return mir_qp_solved;
    }
    return mir_qp_maxIterations;                                               // boxcqp.d:378
}

}  // namespace mirb200
