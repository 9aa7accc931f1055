// Tensor-core (tcgen05 / TMEM / TMA) version of the per-edge stage.
//
// CTA = 6 warps.  Warps 0-3 ("row threads", thread t <-> tile row t <-> TMEM lane t) own one edge each of a
// 128-edge tile: they produce the A operands (tcgen05.st into TMEM, 3xTF32 hi/lo planes), consume the
// accumulators (tcgen05.ld) and do all per-edge math thread-locally (a head = 16 consecutive columns of one
// row, so the attention reduction needs no shuffles).  Warp 4 lane 0 streams the pre-swizzled weight images
// through a 4-stage shared-memory ring with 1-D TMA bulk copies; warp 5 lane 0 issues the MMAs
// (D[128x128] += A[128xK] * W^T, three tf32 MMAs per K-step: hi*hi + lo*hi + hi*lo ~ fp32 accuracy).
// Per-target segmented sums go through a padded shared tile and are finished by the same 128 threads in
// thread-per-channel mapping.
//
// Weight images: for every 128-column GEMM chunk and every K-slab of 32: a 16 KB "hi" plane then a 16 KB "lo"
// plane, each already in the K-major SWIZZLE_128B shared-memory layout (see tc_common.cuh), so one
// contiguous 32 KB bulk copy fills a ring stage.  Built by ai2bmd_b200/weights.py::tc_image().
#pragma once
#include "k_edge.cuh"
#include "tc_common.cuh"

namespace vb {

constexpr int TC_TE = 128;           // edges per tile (= MMA M)
constexpr int TC_STAGES = 4;
constexpr int TC_MAXJOBS = 8;
constexpr int TC_THREADS = 192;
constexpr int TC_LT = D + LDS_PAD;   // padded row length of the staging tile (132 floats)

// TMEM column map (512 columns): A hi plane, A lo plane, two accumulators
constexpr uint32_t TC_COL_AHI = 0, TC_COL_ALO = 128, TC_COL_D0 = 256, TC_COL_D1 = 384;

struct TcJob {
    const float* img;   // weight image of this 128x128 chunk (4 slabs x 32 KB)
    int d_col;          // accumulator column base (TC_COL_D0 / TC_COL_D1)
    int accumulate;     // 0: overwrite the accumulator with the first MMA, 1: add to what is there
};

struct TcShared {
    alignas(1024) uint8_t ring[TC_STAGES][tc::STAGE_BYTES];
    alignas(16) float tile[TC_TE][TC_LT];
    EdgeMeta<TC_TE> meta;
    alignas(8) uint64_t b_full[TC_STAGES];
    uint64_t b_empty[TC_STAGES];
    uint64_t go[TC_MAXJOBS];
    uint64_t done[TC_MAXJOBS];
    uint32_t tmem_base;
};

// one-time CTA setup: barriers + TMEM; returns the TMEM base address
__device__ __forceinline__ uint32_t tc_setup(TcShared& sh, int njobs) {
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; s++) { tc::mbar_init(&sh.b_full[s], 1); tc::mbar_init(&sh.b_empty[s], 1); }
        for (int j = 0; j < njobs; j++) { tc::mbar_init(&sh.go[j], TC_TE); tc::mbar_init(&sh.done[j], 1); }
        tc::fence_barrier_init();
    }
    if (warp == 4) tc::tmem_alloc(&sh.tmem_base, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    return sh.tmem_base;
}
__device__ __forceinline__ void tc_teardown(uint32_t tmem_base) {
    tc::fence_before_sync();
    __syncthreads();
    if ((threadIdx.x >> 5) == 4) tc::tmem_dealloc(tmem_base, 512);
}

// weight producer: one thread; streams the slabs of `njobs` jobs for `ntiles` tiles through the ring
__device__ __forceinline__ void tc_producer(TcShared& sh, const TcJob* jobs, int njobs, int ntiles) {
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < ntiles; t++) {
        for (int j = 0; j < njobs; j++) {
            const char* src = reinterpret_cast<const char*>(jobs[j].img);
#pragma unroll 1
            for (int s = 0; s < D / tc::SLAB_K; s++) {
                tc::mbar_wait(&sh.b_empty[stage], phase ^ 1);
                tc::mbar_arrive_expect_tx(&sh.b_full[stage], tc::STAGE_BYTES);
                tc::tma_load_1d(sh.ring[stage], src + (size_t)s * tc::STAGE_BYTES, tc::STAGE_BYTES, &sh.b_full[stage]);
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    }
}

// MMA issuer: one thread
__device__ __forceinline__ void tc_mma_issuer(TcShared& sh, const TcJob* jobs, int njobs, int ntiles, uint32_t tmem_base) {
    constexpr uint32_t idesc = tc::idesc_tf32(128, 128);
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < ntiles; t++) {
        const uint32_t tpar = (uint32_t)(t & 1);
        for (int j = 0; j < njobs; j++) {
            tc::mbar_wait(&sh.go[j], tpar);
            tc::fence_after_sync();
            const uint32_t d_addr = tmem_base + (uint32_t)jobs[j].d_col;
            uint32_t acc = (uint32_t)jobs[j].accumulate;
#pragma unroll 1
            for (int s = 0; s < D / tc::SLAB_K; s++) {
                tc::mbar_wait(&sh.b_full[stage], phase);
                tc::fence_after_sync();
                const uint32_t bhi = tc::smem_u32(sh.ring[stage]);
                const uint32_t blo = bhi + tc::SLAB_BYTES;
#pragma unroll
                for (int kk = 0; kk < tc::SLAB_K / 8; kk++) {
                    const uint32_t a_off = (uint32_t)(s * tc::SLAB_K + kk * 8);
                    const uint64_t dhi = tc::smem_desc_sw128(bhi + kk * 32);
                    const uint64_t dlo = tc::smem_desc_sw128(blo + kk * 32);
                    tc::mma_tf32_ts(d_addr, tmem_base + TC_COL_ALO + a_off, dhi, idesc, acc);   // lo * hi
                    tc::mma_tf32_ts(d_addr, tmem_base + TC_COL_AHI + a_off, dlo, idesc, 1u);    // hi * lo
                    tc::mma_tf32_ts(d_addr, tmem_base + TC_COL_AHI + a_off, dhi, idesc, 1u);    // hi * hi
                    acc = 1u;
                }
                tc::mma_commit(&sh.b_empty[stage]);
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
            tc::mma_commit(&sh.done[j]);
        }
    }
}

// row thread: publish "the A operand / accumulator for job j is ready"
__device__ __forceinline__ void tc_signal_go(TcShared& sh, int j) {
    tc::wait_st();
    tc::fence_before_sync();
    tc::mbar_arrive(&sh.go[j]);
}
__device__ __forceinline__ void tc_wait_done(TcShared& sh, int j, uint32_t tpar) {
    tc::mbar_wait(&sh.done[j], tpar);
    tc::fence_after_sync();
}
// barrier among the 128 row threads only (named barrier 1)
__device__ __forceinline__ void rows_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Self-test: Dout[128][128] = A[128][128] * W^T with W given as a tc image (validates descriptors,
// swizzle, TMEM lane/column conventions, the ring and every barrier before the edge kernels use them).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ img,
                                                                    float* __restrict__ Dout, int reps) {
    extern __shared__ __align__(1024) uint8_t dyn_raw[];
    TcShared& sh = *reinterpret_cast<TcShared*>(dyn_raw);
    __shared__ TcJob jobs[1];
    if (threadIdx.x == 0) jobs[0] = TcJob{img, (int)TC_COL_D0, 0};
    const uint32_t tmem = tc_setup(sh, 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 4) {
        if (lane == 0) tc_producer(sh, jobs, 1, reps);
    } else if (warp == 5) {
        if (lane == 0) tc_mma_issuer(sh, jobs, 1, reps, tmem);
    } else {
        const int row = threadIdx.x;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int t = 0; t < reps; t++) {
            for (int c0 = 0; c0 < D; c0 += 16) {
                float v[16];
#pragma unroll
                for (int q = 0; q < 16; q += 4) {
                    const float4 x = ld4(A + (size_t)row * D + c0 + q);
                    v[q] = x.x; v[q + 1] = x.y; v[q + 2] = x.z; v[q + 3] = x.w;
                }
                tc::store_a16(tmem + lane_base + TC_COL_AHI, tmem + lane_base + TC_COL_ALO, c0, v);
            }
            tc_signal_go(sh, 0);
            tc_wait_done(sh, 0, (uint32_t)(t & 1));
            for (int c0 = 0; c0 < D; c0 += 16) {
                float v[16];
                tc::tmem_ld16(tmem + lane_base + TC_COL_D0 + c0, v);
#pragma unroll
                for (int q = 0; q < 16; q += 4) st4(Dout + (size_t)row * D + c0 + q, f4(v[q], v[q + 1], v[q + 2], v[q + 3]));
            }
            tc::fence_before_sync();
            rows_sync();          // every row thread finished reading D before the next repetition overwrites it
            tc::fence_after_sync();
        }
    }
    tc_teardown(tmem);
}

}  // namespace vb
