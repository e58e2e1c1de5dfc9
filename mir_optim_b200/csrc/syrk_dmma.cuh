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

struct SyrkArgs {
    const double* J;        // rowsPadded x ldj, row-major; rows >= rows are zero (the engine pads to a multiple of KT)
    long long     tiles;    // rowsPadded / KT
    int           ldj;      // even, <= 128
    int           nblk;     // ceil(n / 8)
    double*       partial;  // gridDim.x images of 128 x 128 doubles
    const int*    gate;     // device flag: 0 => nothing to do this pass (J unchanged)
    const int*    done;     // device flag: solve finished
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

// Consumer warp W: block-rows A = W and B = 15 - W of the lower triangle.
template <int W>
__device__ __forceinline__ void syrk_consume(const SyrkArgs& a, const unsigned char* stages, unsigned long long* full,
                                             unsigned long long* empty, long long tile0, long long tile1, int lane)
{
    constexpr int RA = W, RB = 15 - W;          // RA < RB
    constexpr int NF = RB + 1;                  // fragments per k-step: block-columns 0..RB
    double ca[RA + 1][2], cb[RB + 1][2];
#pragma unroll
    for (int j = 0; j <= RA; ++j) { ca[j][0] = 0.0; ca[j][1] = 0.0; }
#pragma unroll
    for (int j = 0; j <= RB; ++j) { cb[j][0] = 0.0; cb[j][1] = 0.0; }
    const int nblk = a.nblk;
    const int fragOff = (lane & 3) * SYRK_PITCH_D + (lane >> 2);

    for (long long it = tile0; it < tile1; ++it) {
        const long long k = it - tile0;
        const int s = (int)(k % SYRK_STAGES);
        mbar_wait(&full[s], (unsigned)((k / SYRK_STAGES) & 1));
        const double* tile = reinterpret_cast<const double*>(stages + (size_t)s * SYRK_STAGE_BYTES) + fragOff;
#pragma unroll 2
        for (int ks = 0; ks < SYRK_KT / 4; ++ks) {
            const double* row = tile + ks * 4 * SYRK_PITCH_D;
            double f[NF];
            if (nblk == 16) {
#pragma unroll
                for (int j = 0; j < NF; ++j) f[j] = row[8 * j];
#pragma unroll
                for (int j = 0; j <= RA; ++j) dmma884(ca[j], f[RA], f[j]);
#pragma unroll
                for (int j = 0; j <= RB; ++j) dmma884(cb[j], f[RB], f[j]);
            } else {
#pragma unroll
                for (int j = 0; j < NF; ++j) f[j] = (j < nblk) ? row[8 * j] : 0.0;
                if (RA < nblk) {
#pragma unroll
                    for (int j = 0; j <= RA; ++j) dmma884(ca[j], f[RA], f[j]);
                }
                if (RB < nblk) {
#pragma unroll
                    for (int j = 0; j <= RB; ++j) if (j < nblk) dmma884(cb[j], f[RB], f[j]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    // epilogue: C fragment of m8n8k4 = row lane/4, columns 2*(lane%4) + {0,1}
    double* out = a.partial + (size_t)blockIdx.x * (SYRK_NPAD * SYRK_NPAD);
    const int r = lane >> 2, c = (lane & 3) * 2;
#pragma unroll
    for (int j = 0; j <= RA; ++j)
        *reinterpret_cast<double2*>(out + (size_t)(8 * RA + r) * SYRK_NPAD + 8 * j + c) = make_double2(ca[j][0], ca[j][1]);
#pragma unroll
    for (int j = 0; j <= RB; ++j)
        *reinterpret_cast<double2*>(out + (size_t)(8 * RB + r) * SYRK_NPAD + 8 * j + c) = make_double2(cb[j][0], cb[j][1]);
}

__global__ void __launch_bounds__(SYRK_THREADS, 1) syrk_dmma_kernel(const SyrkArgs a)
{
    if (*a.done || *a.gate == 0) return;
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
        switch (warp) {
            case 0: syrk_consume<0>(a, stages, full, empty, tile0, tile1, lane); break;
            case 1: syrk_consume<1>(a, stages, full, empty, tile0, tile1, lane); break;
            case 2: syrk_consume<2>(a, stages, full, empty, tile0, tile1, lane); break;
            case 3: syrk_consume<3>(a, stages, full, empty, tile0, tile1, lane); break;
            case 4: syrk_consume<4>(a, stages, full, empty, tile0, tile1, lane); break;
            case 5: syrk_consume<5>(a, stages, full, empty, tile0, tile1, lane); break;
            case 6: syrk_consume<6>(a, stages, full, empty, tile0, tile1, lane); break;
            default: syrk_consume<7>(a, stages, full, empty, tile0, tile1, lane); break;
        }
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
