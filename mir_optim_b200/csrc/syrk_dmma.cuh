// syrk_dmma.cuh -- J^T J (lower triangle) of a tall row-major J on the FP64 tensor pipe.
//
// Replaces `syrk(Uplo.Lower, 1, J.transposed, 0, JJ)` at least_squares.d:1065 for the single large
// problem (BASELINE configs[3]: m = 4M rows, n = 128 columns).
//
// sm_100a has no tcgen05 f64 kind; FP64 tensor work is warp-level `mma.sync.m8n8k4.f64`
// (SASS DMMA.8x8x4, measured 37.0 TFLOP/s on B200 by peaks.cu).  Accumulators therefore live in
// registers, not TMEM.  Layout of one persistent CTA (1 per SM):
//
//   * warp 8 = producer: for every tile of KT = 32 rows it arms the stage's "full" mbarrier with the
//     byte count and issues one TMA bulk copy (cp.async.bulk, SASS UBLKCP) per row into a 4-stage
//     shared-memory ring.  Rows are ldj*8 bytes in HBM and 1056 bytes apart in shared memory:
//     1056 = 32 (mod 128), so the 4 rows x 4 columns a half-warp touches in one fragment load fall
//     into 16 distinct 8-byte bank pairs -- fragment loads are conflict-free without a swizzle.
//   * warps 0-7 = consumers.  The 128 x 128 result is 16 x 16 blocks of 8 x 8; only the 136 blocks
//     of the lower triangle are computed.  Warp w owns block-rows w and 15-w (w+1 and 16-w blocks:
//     17 blocks per warp, perfectly balanced) = 34 accumulator registers.  For each k-step (4 rows)
//     a warp loads 16-w fragments (one LDS.64 each; the A fragment of block-row i is the B fragment
//     of block-column i, so nothing is loaded twice) and issues 17 DMMAs.
//   * split-K over CTAs: each CTA owns a contiguous range of row tiles and writes its partial
//     128 x 128 block image; syrk_reduce_kernel adds the partials in CTA order (deterministic,
//     bit-identical on every run and every rank layout with the same grid).
//
// Algorithmic flops m*n*(n+1); executed m*2*64*136 = 1.054x that (upper halves of the 16
// diagonal blocks).  J traffic: one read of J (8*ldj B/row), fully overlapped with the DMMAs.
#pragma once
#include "common.cuh"

namespace mirb200 {

constexpr int SYRK_KT = 32;                         // rows per pipeline stage
constexpr int SYRK_STAGES = 4;
constexpr int SYRK_NPAD = 128;                      // columns covered by the block grid
constexpr int SYRK_PITCH = SYRK_NPAD * 8 + 32;      // bytes between rows in shared memory
constexpr int SYRK_PITCH_D = SYRK_PITCH / 8;        // ... in doubles
constexpr int SYRK_CONSUMERS = 8;
constexpr int SYRK_THREADS = (SYRK_CONSUMERS + 1) * 32;
constexpr size_t SYRK_STAGE_BYTES = (size_t)SYRK_KT * SYRK_PITCH;
constexpr size_t SYRK_SMEM_BYTES = SYRK_STAGES * SYRK_STAGE_BYTES + 2 * SYRK_STAGES * sizeof(unsigned long long) + 128;

// Optional fused Broyden update (least_squares.d:999-1006 + the gemv at :1052): when enabled and *gate == 1, every tile
// is updated IN the shared-memory ring before the DMMAs read it --
//     v = ((f_old - f_new) + J_row . dX) * (-1 / |dX|^2),   J_row += v dX',   J'y += J_row * f_new
// -- written back to HBM from registers, and J'y leaves through per-CTA partials added in CTA order by the last CTA.
// The separate Broyden kernel streamed J twice (read + write, 8.2 GB, 1.40 ms HBM-bound at 4M x 128) before this kernel
// read it a third time; fused, the update hides under the tensor pipe.  Row arithmetic is the standalone kernel's,
// operation for operation (same lane layout and shuffle tree), so J comes out bit-identical.
struct SyrkBroyden {
    int           enabled;
    double*       J;            // writable alias of SyrkArgs::J
    const double* buf0;         // y / mBuffer pair, selected by *ysel as in lm_large.cuh
    const double* buf1;
    const int*    ysel;
    const double* dX;           // accepted step, n values
    const double* deltaX_dot;   // |dX|^2
    long long     rows;         // valid rows of this rank (<= tiles * KT)
    double*       partJy;       // gridDim.x * 128
    unsigned*     ticket;       // last-CTA ticket (zero on entry, reset on exit)
    double*       outJy;        // n values
    int           n;
};

struct SyrkArgs {
    const double* J;        // rowsPadded x ldj, row-major; rows >= rows are zero (the engine pads to a multiple of KT)
    long long     tiles;    // rowsPadded / KT
    int           ldj;      // even, <= 128
    int           nblk;     // ceil(n / 8)
    double*       partial;  // gridDim.x images of 128 x 128 doubles
    const int*    gate;     // device flag: 0 => nothing to do this pass (J unchanged); 1 = Broyden pass, 2 = fresh Jacobian
    const int*    done;     // device flag: solve finished
    SyrkBroyden   bro;      // bro.enabled == 0: plain J^T J whatever the gate says
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// Consumer warp W owns block-rows A = W and B = 15 - W of the lower triangle: W + 1 and 16 - W blocks = 17 accumulator
// pairs for every W, kept in one uniform array so that the per-tile loop (and the fused Broyden update in it) exists
// ONCE in the kernel and only the DMMA block of a tile is specialised per warp -- eight fully specialised consumer
// loops plus an inlined update overflowed the instruction cache (ncu: stall_no_instruction 8.4, tensor pipe 37 %).
template <int W>
__device__ __forceinline__ void syrk_mma_tile(double (&c)[17][2], const double* tile, int nblk)
{
    constexpr int RA = W, RB = 15 - W;          // RA < RB
    constexpr int NF = RB + 1;                  // fragments per k-step: block-columns 0..RB
#pragma unroll 2
    for (int ks = 0; ks < SYRK_KT / 4; ++ks) {
        const double* row = tile + ks * 4 * SYRK_PITCH_D;
        double f[NF];
        if (nblk == 16) {
#pragma unroll
            for (int j = 0; j < NF; ++j) f[j] = row[8 * j];
#pragma unroll
            for (int j = 0; j <= RA; ++j) dmma884(c[j], f[RA], f[j]);
#pragma unroll
            for (int j = 0; j <= RB; ++j) dmma884(c[RA + 1 + j], f[RB], f[j]);
        } else {
#pragma unroll
            for (int j = 0; j < NF; ++j) f[j] = (j < nblk) ? row[8 * j] : 0.0;
            if (RA < nblk) {
#pragma unroll
                for (int j = 0; j <= RA; ++j) dmma884(c[j], f[RA], f[j]);
            }
            if (RB < nblk) {
#pragma unroll
                for (int j = 0; j <= RB; ++j) if (j < nblk) dmma884(c[RA + 1 + j], f[RB], f[j]);
            }
        }
    }
}

// epilogue: C fragment of m8n8k4 = row lane/4, columns 2*(lane%4) + {0,1}
template <int W>
__device__ __forceinline__ void syrk_store_blocks(const double (&c)[17][2], double* out, int lane)
{
    constexpr int RA = W, RB = 15 - W;
    const int r = lane >> 2, col = (lane & 3) * 2;
#pragma unroll
    for (int j = 0; j <= RA; ++j)
        *reinterpret_cast<double2*>(out + (size_t)(8 * RA + r) * SYRK_NPAD + 8 * j + col) = make_double2(c[j][0], c[j][1]);
#pragma unroll
    for (int j = 0; j <= RB; ++j)
        *reinterpret_cast<double2*>(out + (size_t)(8 * RB + r) * SYRK_NPAD + 8 * j + col) = make_double2(c[RA + 1 + j][0], c[RA + 1 + j][1]);
}

__global__ void __launch_bounds__(SYRK_THREADS, 1) syrk_dmma_kernel(const SyrkArgs a)
{
    if (*a.done || *a.gate == 0) return;
    const bool fuse = a.bro.enabled && *a.gate == 1;
    __shared__ double s_jy[SYRK_CONSUMERS * SYRK_NPAD];
    __shared__ bool s_last;
    extern __shared__ __align__(128) unsigned char syrk_smem[];
    unsigned char* stages = syrk_smem;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(syrk_smem + SYRK_STAGES * SYRK_STAGE_BYTES);
    unsigned long long* empty = full + SYRK_STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // columns >= ldj are never written by the row copies: zero the ring once so padded blocks contribute 0
    for (size_t i = tid; i < SYRK_STAGES * SYRK_STAGE_BYTES / 16; i += SYRK_THREADS)
        reinterpret_cast<uint4*>(stages)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int s = 0; s < SYRK_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], SYRK_CONSUMERS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy zero fill before async-proxy writes
    __syncthreads();

    // contiguous range of row tiles for this CTA
    const long long per = (a.tiles + gridDim.x - 1) / gridDim.x;
    long long tile0 = (long long)blockIdx.x * per;
    long long tile1 = tile0 + per < a.tiles ? tile0 + per : a.tiles;
    if (tile0 > tile1) tile0 = tile1;

    if (warp == SYRK_CONSUMERS) {
        const unsigned rowBytes = (unsigned)a.ldj * 8u;
        for (long long it = tile0; it < tile1; ++it) {
            const long long k = it - tile0;
            const int s = (int)(k % SYRK_STAGES);
            if (k >= SYRK_STAGES) mbar_wait(&empty[s], (unsigned)(((k / SYRK_STAGES) - 1) & 1));
            if (lane == 0) mbar_expect_tx(&full[s], rowBytes * SYRK_KT);
            __syncwarp();
            const double* src = a.J + ((size_t)it * SYRK_KT + lane) * a.ldj;
            tma_bulk_g2s(stages + (size_t)s * SYRK_STAGE_BYTES + (size_t)lane * SYRK_PITCH, src, rowBytes, &full[s]);
        }
    } else {
        double c[17][2];
#pragma unroll
        for (int j = 0; j < 17; ++j) { c[j][0] = 0.0; c[j][1] = 0.0; }
        const int nblk = a.nblk;
        const int fragOff = (lane & 3) * SYRK_PITCH_D + (lane >> 2);

        // fused Broyden state: this warp updates rows 4*warp .. 4*warp+3 of every tile; lane owns columns lane + 32 q
        double dx[4] = {0.0, 0.0, 0.0, 0.0}, jy[4] = {0.0, 0.0, 0.0, 0.0};
        double negd = 0.0, pf = 0.0;
        const double* yv = nullptr; const double* mbv = nullptr;
        auto fetch_pf = [&](long long tileIdx) -> double {       // lanes 0-3: f_new of my 4 rows, lanes 4-7: f_old
            const long long row = tileIdx * SYRK_KT + 4 * warp + (lane & 3);
            if (lane < 8 && tileIdx < tile1 && row < a.bro.rows) return (lane < 4 ? yv : mbv)[row];
            return 0.0;
        };
        if (fuse) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { const int col = lane + 32 * q; dx[q] = col < a.bro.n ? a.bro.dX[col] : 0.0; }
            negd = -(1.0 / *a.bro.deltaX_dot);                                                   // LS:1001
            const int ys = *a.bro.ysel;
            yv = ys ? a.bro.buf1 : a.bro.buf0; mbv = ys ? a.bro.buf0 : a.bro.buf1;
            pf = fetch_pf(tile0);
        }

        for (long long it = tile0; it < tile1; ++it) {
            const long long k = it - tile0;
            const int s = (int)(k % SYRK_STAGES);
            double pfNext = 0.0;
            if (fuse) pfNext = fetch_pf(it + 1);                  // next tile's residuals: in flight during this tile
            mbar_wait(&full[s], (unsigned)((k / SYRK_STAGES) & 1));
            if (fuse) {
                double* trow = reinterpret_cast<double*>(stages + (size_t)s * SYRK_STAGE_BYTES) + (size_t)(4 * warp) * SYRK_PITCH_D;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const long long row = it * SYRK_KT + 4 * warp + r;
                    double jv[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) jv[q] = trow[r * SYRK_PITCH_D + lane + 32 * q];
                    double acc = 0.0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc += jv[q] * dx[q];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
                    const double yr = __shfl_sync(0xffffffffu, pf, r), mbr = __shfl_sync(0xffffffffu, pf, 4 + r);
                    const double v = (row < a.bro.rows) ? ((mbr - yr) + acc) * negd : 0.0;       // LS:1003-1005 (padding rows stay zero)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int col = lane + 32 * q;
                        jv[q] += v * dx[q];                                                      // LS:1006
                        trow[r * SYRK_PITCH_D + col] = jv[q];
                        if (col < a.ldj && row < a.bro.rows) a.bro.J[(size_t)row * a.ldj + col] = jv[q];
                        jy[q] += jv[q] * yr;                                                     // LS:1052
                    }
                }
                pf = pfNext;
                asm volatile("bar.sync 1, %0;" ::"n"(SYRK_CONSUMERS * 32) : "memory");          // all 32 rows updated before any fragment load
            }
            const double* tile = reinterpret_cast<const double*>(stages + (size_t)s * SYRK_STAGE_BYTES) + fragOff;
            switch (warp) {
                case 0: syrk_mma_tile<0>(c, tile, nblk); break;
                case 1: syrk_mma_tile<1>(c, tile, nblk); break;
                case 2: syrk_mma_tile<2>(c, tile, nblk); break;
                case 3: syrk_mma_tile<3>(c, tile, nblk); break;
                case 4: syrk_mma_tile<4>(c, tile, nblk); break;
                case 5: syrk_mma_tile<5>(c, tile, nblk); break;
                case 6: syrk_mma_tile<6>(c, tile, nblk); break;
                default: syrk_mma_tile<7>(c, tile, nblk); break;
            }
            if (fuse) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy tile writes before the next TMA fill
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }

        double* out = a.partial + (size_t)blockIdx.x * (SYRK_NPAD * SYRK_NPAD);
        switch (warp) {
            case 0: syrk_store_blocks<0>(c, out, lane); break;
            case 1: syrk_store_blocks<1>(c, out, lane); break;
            case 2: syrk_store_blocks<2>(c, out, lane); break;
            case 3: syrk_store_blocks<3>(c, out, lane); break;
            case 4: syrk_store_blocks<4>(c, out, lane); break;
            case 5: syrk_store_blocks<5>(c, out, lane); break;
            case 6: syrk_store_blocks<6>(c, out, lane); break;
            default: syrk_store_blocks<7>(c, out, lane); break;
        }
        if (fuse) {
#pragma unroll
            for (int q = 0; q < 4; ++q) s_jy[warp * SYRK_NPAD + lane + 32 * q] = jy[q];
        }
    }
    if (!fuse) return;

    // J'y: warps in fixed order -> this CTA's partial -> the last CTA adds the partials in CTA order
    __syncthreads();
    if (tid < SYRK_NPAD) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < SYRK_CONSUMERS; ++w) sum += s_jy[w * SYRK_NPAD + tid];
        a.bro.partJy[(size_t)blockIdx.x * SYRK_NPAD + tid] = sum;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(a.bro.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (int col = tid; col < a.bro.n; col += SYRK_THREADS) {
            double s0 = 0.0, s1 = 0.0;
            int b = 0;
            for (; b + 2 <= (int)gridDim.x; b += 2) { s0 += a.bro.partJy[(size_t)b * SYRK_NPAD + col]; s1 += a.bro.partJy[(size_t)(b + 1) * SYRK_NPAD + col]; }
            if (b < (int)gridDim.x) s0 += a.bro.partJy[(size_t)b * SYRK_NPAD + col];
            a.bro.outJy[col] = s0 + s1;
        }
        if (tid == 0) *a.bro.ticket = 0;
    }
}

// Second stage: packed[tri(i, j)] = sum over CTAs of partial[c][i][j], in CTA order.
__global__ void __launch_bounds__(256) syrk_reduce_kernel(const double* __restrict__ partial, int nparts, int n, double* __restrict__ packed,
                                                          const int* gate, const int* done)
{
    if (*done || *gate == 0) return;
    const int np = n * (n + 1) / 2;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < np; e += gridDim.x * blockDim.x) {
        // invert e = i (i + 1) / 2 + j
        int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
        while (i * (i + 1) / 2 > e) --i;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        const int j = e - i * (i + 1) / 2;
        const double* p = partial + (size_t)i * SYRK_NPAD + j;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int c = 0;
        for (; c + 16 <= nparts; c += 16) {            // sixteen loads in flight (one L2 round trip), added in the order of the loop below
            double v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = __ldcg(p + (size_t)(c + k) * (SYRK_NPAD * SYRK_NPAD));
#pragma unroll
            for (int k = 0; k < 16; k += 4) { s0 += v[k]; s1 += v[k + 1]; s2 += v[k + 2]; s3 += v[k + 3]; }
        }
        for (; c + 4 <= nparts; c += 4) {
            s0 += p[(size_t)(c + 0) * (SYRK_NPAD * SYRK_NPAD)];
            s1 += p[(size_t)(c + 1) * (SYRK_NPAD * SYRK_NPAD)];
            s2 += p[(size_t)(c + 2) * (SYRK_NPAD * SYRK_NPAD)];
            s3 += p[(size_t)(c + 3) * (SYRK_NPAD * SYRK_NPAD)];
        }
        for (; c < nparts; ++c) s0 += p[(size_t)c * (SYRK_NPAD * SYRK_NPAD)];
        packed[e] = (s0 + s1) + (s2 + s3);
    }
}

}  // namespace mirb200
