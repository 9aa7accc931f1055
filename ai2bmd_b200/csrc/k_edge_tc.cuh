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
constexpr int TC_MAXJOBS = 12;
constexpr int TC_THREADS = 192;
constexpr int TC_LT = D + LDS_PAD;   // padded row length of the staging tile (132 floats)
constexpr int TC_TILE_EXT = 1792;    // floats appended to the staging tile for the fused kernels' node stage (7 KB)

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
    float tile_ext[TC_TILE_EXT];        // fused kernels (k_fused.cuh): the node stage's shared rows start at `tile` and may run on into here
    EdgeMeta<TC_TE> meta;
    alignas(8) uint64_t b_full[TC_STAGES];
    uint64_t b_empty[TC_STAGES];
    uint64_t go[TC_MAXJOBS];
    uint64_t done[TC_MAXJOBS];
    uint64_t b_tile;                    // per-edge feature rows landed in `tile` (one arrival + byte count per compute warp)
    uint32_t tmem_base;
    alignas(16) float eacc[TC_TE][4];   // per-edge adjoint scalars of the current tile: dE/dC, dE/dd[3]
    float gattn[TC_TE][H];              // adjoint kernel: dE/da_h per edge
};

// the ring must start on a 1024 B boundary of the shared window (swizzle atom): align the dynamic block by hand
__device__ __forceinline__ TcShared* tc_shared_base(uint8_t* raw) {
    const uint32_t s = tc::smem_u32(raw);
    return reinterpret_cast<TcShared*>(raw + (((s + 1023u) & ~1023u) - s));
}
constexpr size_t TC_SMEM_BYTES = sizeof(TcShared) + 1024;

// one-time CTA setup: barriers + TMEM; returns the TMEM base address
__device__ __forceinline__ uint32_t tc_setup(TcShared& sh, int njobs, int go_count = TC_TE) {
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; s++) { tc::mbar_init(&sh.b_full[s], 1); tc::mbar_init(&sh.b_empty[s], 1); }
        for (int j = 0; j < njobs; j++) { tc::mbar_init(&sh.go[j], go_count); tc::mbar_init(&sh.done[j], 1); }
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
__device__ __forceinline__ void tc_mma_issuer(TcShared& sh, const TcJob* jobs, int njobs, int ntiles, uint32_t tmem_base,
                                              unsigned long long* tl = nullptr) {
    constexpr uint32_t idesc = tc::idesc_tf32(128, 128);
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < ntiles; t++) {
        const uint32_t tpar = (uint32_t)(t & 1);
        for (int j = 0; j < njobs; j++) {
            tc::mbar_wait(&sh.go[j], tpar);
            tc::fence_after_sync();
            if (tl != nullptr && t == 0) tl[32 + 2 * j] = (unsigned long long)clock64();
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
            if (tl != nullptr && t == 0) tl[33 + 2 * j] = (unsigned long long)clock64();
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
    TcShared& sh = *tc_shared_base(dyn_raw);
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


// =====================================================================================================
// Tensor-core edge stage, hybrid layout.
//   * 16 compute warps keep the coalesced "lane owns 4 channels" layout of k_edge.cuh for every global
//     gather / scatter / elementwise step (one 512 B request per node row per warp);
//   * the five 128x128x128 contractions per tile run on tcgen05: the A operand is moved from the padded
//     shared staging tile into TMEM (thread-per-row, 3xTF32 planes), the accumulator comes back the same way;
//   * warp 16 lane 0 = TMA weight producer, warp 17 lane 0 = MMA issuer (tc_producer / tc_mma_issuer);
//   * a tile holds ROWS = 32 / 64 / 96 / 128 edges (template parameter): MMA M stays 128, TMEM lanes >= ROWS are never
//     written or read, compute warp w owns rows [w*ROWS/16, (w+1)*ROWS/16) in the coalesced phases.
// =====================================================================================================
namespace vb {

constexpr int TC2_CWARPS = 16;                      // compute warps
constexpr int TC2_CTHREADS = TC2_CWARPS * 32;       // 512
constexpr int TC2_THREADS = TC2_CTHREADS + 64;      // + producer warp + MMA warp
constexpr int TC2_CBLK = D / (TC2_CWARPS / 4);         // columns each warp moves between the staging tile and TMEM
constexpr int TC2_NGRP = TC2_CTHREADS / D;             // channel groups in the per-target aggregation phases

struct EdgeTcArgs {
    int layer;
    ModelW mw;
    Workspace ws;
    TcJob jobs[TC_MAXJOBS];
    int njobs;
    int tile_rows;              // edges per tile, <= the kernel's ROWS
    unsigned long long* tl;     // optional timeline (SM clock stamps of CTA 0, first tile); nullptr = off
};
constexpr int TC_TL_SLOTS = 64;  // [0,32): compute thread 0 phase stamps ; [32,48): MMA issuer (go seen / MMAs issued per job)
#define TC_TL(k) do { if (a.tl != nullptr && blockIdx.x == 0 && it == 0 && threadIdx.x == 0) a.tl[k] = (unsigned long long)clock64(); } while (0)

__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, %0;" ::"n"(TC2_CTHREADS) : "memory"); }

__device__ __forceinline__ uint32_t tc2_setup(TcShared& sh, int njobs) {
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; s++) { tc::mbar_init(&sh.b_full[s], 1); tc::mbar_init(&sh.b_empty[s], 1); }
        for (int j = 0; j < njobs; j++) { tc::mbar_init(&sh.go[j], TC2_CTHREADS); tc::mbar_init(&sh.done[j], 1); }
        tc::mbar_init(&sh.b_tile, TC2_CWARPS);
        tc::fence_barrier_init();
    }
    if (warp == TC2_CWARPS) tc::tmem_alloc(&sh.tmem_base, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    return sh.tmem_base;
}
__device__ __forceinline__ void tc2_teardown(uint32_t tmem_base) {
    tc::fence_before_sync();
    __syncthreads();
    if ((threadIdx.x >> 5) == TC2_CWARPS) tc::tmem_dealloc(tmem_base, 512);
}

// staging tile (fp32, row-major, padded) -> A operand planes in TMEM.  Compute warp w serves TMEM lane quarter
// w&3 (rows 32*(w&3)..+31) and column half w>>2.
template <int ROWS>
__device__ __forceinline__ void tc2_tile_to_a(TcShared& sh, uint32_t tmem, int warp, int lane) {
    if ((warp & 3) * 32 >= ROWS) return;          // short tiles: TMEM lanes >= ROWS stay stale (their D rows are never read)
    const int row = (warp & 3) * 32 + lane, ch = (warp >> 2) * TC2_CBLK;
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll
    for (int c0 = 0; c0 < TC2_CBLK; c0 += 16) {
        float v[16];
#pragma unroll
        for (int q = 0; q < 16; q += 4) {
            const float4 x = ld4(&sh.tile[row][ch + c0 + q]);
            v[q] = x.x; v[q + 1] = x.y; v[q + 2] = x.z; v[q + 3] = x.w;
        }
        tc::store_a16(tl + TC_COL_AHI, tl + TC_COL_ALO, ch + c0, v);
    }
}
// accumulator (TMEM) -> staging tile
template <int ROWS>
__device__ __forceinline__ void tc2_d_to_tile(TcShared& sh, uint32_t tmem, uint32_t d_col, int warp, int lane) {
    if ((warp & 3) * 32 >= ROWS) return;
    const int row = (warp & 3) * 32 + lane, ch = (warp >> 2) * TC2_CBLK;
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16) + d_col;
    constexpr int NB16 = TC2_CBLK / 16;
    uint32_t r[NB16][16];
#pragma unroll
    for (int b = 0; b < NB16; b++) tc::tmem_ld16_nowait(tl + ch + b * 16, r[b]);
    tc::wait_ld();
#pragma unroll
    for (int b = 0; b < NB16; b++)
#pragma unroll
        for (int q = 0; q < 16; q += 4)
            st4(&sh.tile[row][ch + b * 16 + q], f4(__uint_as_float(r[b][q]), __uint_as_float(r[b][q + 1]),
                                                    __uint_as_float(r[b][q + 2]), __uint_as_float(r[b][q + 3])));
}
__device__ __forceinline__ void tc2_go(TcShared& sh, int j) {      // every compute thread
    tc::wait_st();
    tc::fence_before_sync();
    tc::mbar_arrive(&sh.go[j]);
}

// ---------------------------------------------------------------------------------------------
// forward (math and reference lines: see edge_fwd_kernel in k_edge.cuh)
// job order: dk -> D0, dv -> D1, [f -> D0], s1 -> D1, s2 -> D0
// ---------------------------------------------------------------------------------------------
template <int ROWS>
__global__ void __launch_bounds__(TC2_THREADS, 1) edge_fwd_tc_kernel(const __grid_constant__ EdgeTcArgs a) {
    pdl_entry();
    extern __shared__ __align__(1024) uint8_t dyn_raw[];
    TcShared& sh = *tc_shared_base(dyn_raw);
    const Workspace& ws = a.ws;
    const int l = a.layer;
    const LayerW& lw = a.mw.layer[l];
    const bool upd = (l < L - 1);
    const int J_DK = 0, J_DV = 1, J_F = 2, J_S1 = upd ? 3 : 2, J_S2 = upd ? 4 : 3;
    constexpr int RPW = ROWS / TC2_CWARPS;      // rows per compute warp in the coalesced phases
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, col = lane * 4;
    const int E = ws.rowptr[ws.N];
    const int trows = min(ROWS, max(16, a.tile_rows));          // edges per tile (<= ROWS, chosen on the host so the tiles fill whole waves)
    const int ntiles_total = (E + trows - 1) / trows;
    const int my_tiles = ((int)blockIdx.x < ntiles_total) ? (ntiles_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (a.tl != nullptr && blockIdx.x == 0 && threadIdx.x == 0) a.tl[0] = (unsigned long long)clock64();
    const uint32_t tmem = tc2_setup(sh, a.njobs);
    if (a.tl != nullptr && blockIdx.x == 0 && threadIdx.x == 0) a.tl[1] = (unsigned long long)clock64();

    if (warp == TC2_CWARPS) {
        if (lane == 0) tc_producer(sh, a.jobs, a.njobs, my_tiles);
    } else if (warp == TC2_CWARPS + 1) {
        if (lane == 0) tc_mma_issuer(sh, a.jobs, a.njobs, my_tiles, tmem, blockIdx.x == 0 ? a.tl : nullptr);
    } else {
        const float* __restrict__ Fin = ws.F[l];
        float* __restrict__ Fout = upd ? ws.F[l + 1] : nullptr;
        const float* __restrict__ QKV = ws.QKV[l];
        const float* __restrict__ VN = ws.VN[l];
        const float* __restrict__ TU = ws.TU[l];
        float* __restrict__ P1 = ws.P1[l];
        float* __restrict__ SP = ws.SP[l];
        float* __restrict__ ATT = ws.ATT[l];
        const int cch = threadIdx.x & (D - 1), grp = threadIdx.x >> 7;      // aggregation role: channel, target parity
        for (int it = 0; it < my_tiles; it++) {
            const uint32_t tpar = (uint32_t)(it & 1);
            const int e0 = ((int)blockIdx.x + it * (int)gridDim.x) * trows;
            const int nvalid = min(trows, E - e0);
            // rows are dealt to the compute warps in contiguous runs of rpw = ceil(nvalid / 16): a tile shorter than ROWS
            // keeps every warp busy (slot s of a warp is row warp * rpw + s, valid while s < rpw and the row exists)
            const int rpw = (nvalid + TC2_CWARPS - 1) / TC2_CWARPS, r0 = warp * rpw;
            // ---- per-edge feature rows -> staging tile by TMA bulk copies (one 512 B row each, padded rows in shared
            //      memory), completion on an mbarrier ----
            {
                constexpr int RW = ROWS / TC2_CWARPS;                 // rows a warp issues
                const int w0 = warp * RW, wn = max(0, min(RW, nvalid - w0));
                if (lane == 0) {
                    if (wn > 0) tc::mbar_arrive_expect_tx(&sh.b_tile, (uint32_t)wn * D * 4);
                    else tc::mbar_arrive(&sh.b_tile);
                }
                __syncwarp();
                if (lane < wn) tc::tma_load_1d(&sh.tile[w0 + lane][0], Fin + (size_t)(e0 + w0 + lane) * D, D * 4, &sh.b_tile);
            }
            load_edge_meta<TC_TE, TC2_CTHREADS>(sh.meta, ws, e0, nvalid);
            tc::mbar_wait(&sh.b_tile, tpar);
            csync();
            TC_TL(2);
            tc2_tile_to_a<ROWS>(sh, tmem, warp, lane);
            tc2_go(sh, J_DK);
            tc2_go(sh, J_DV);
            TC_TL(3);
            // ---- dk -> attention weights ----
            float Areg[RPW];
            tc::mbar_wait(&sh.done[J_DK], tpar);
            tc::fence_after_sync();
            TC_TL(4);
            csync();                                              // everyone finished reading f from the tile
            tc2_d_to_tile<ROWS>(sh, tmem, TC_COL_D0, warp, lane);
            tc::fence_before_sync();
            csync();
            {
                TC_TL(5);
                const float4 bb = ldg4(lw.b1 + col);
#pragma unroll
                for (int r = 0; r < RPW; r++) {
                    if (r >= rpw) break;
                    const int row = r0 + r;
                    const float4 qi = ldg4(QKV + (size_t)sh.meta.dst[row] * 3 * D + col);
                    const float4 kj = ldg4(QKV + (size_t)sh.meta.src[row] * 3 * D + D + col);
                    const float4 P = ld4(&sh.tile[row][col]) + bb;
                    const float av = quad_sum(hsum4(qi * kj * silu4(P)));
                    Areg[r] = silu_(av) * sh.meta.C[row];
                    if ((r < rpw && row < nvalid)) {
                        st4(P1 + (size_t)(e0 + row) * 3 * D + col, P);
                        if ((lane & 3) == 0) ATT[(size_t)(e0 + row) * H + (lane >> 2)] = av;
                    }
                }
            }
            TC_TL(6);
            if (upd) { tc::fence_before_sync(); tc::mbar_arrive(&sh.go[J_F]); }     // D0 is free
            // ---- dv -> message m (in place in the tile) ----
            tc::mbar_wait(&sh.done[J_DV], tpar);
            tc::fence_after_sync();
            TC_TL(7);
            csync();
            tc2_d_to_tile<ROWS>(sh, tmem, TC_COL_D1, warp, lane);
            csync();
            {
                TC_TL(8);
                const float4 bb = ldg4(lw.b1 + D + col);
#pragma unroll
                for (int r = 0; r < RPW; r++) {
                    if (r >= rpw) break;
                    const int row = r0 + r;
                    const float4 vj = ldg4(QKV + (size_t)sh.meta.src[row] * 3 * D + 2 * D + col);
                    const float4 P = ld4(&sh.tile[row][col]) + bb;
                    st4(&sh.tile[row][col], vj * silu4(P) * Areg[r]);
                    if ((r < rpw && row < nvalid)) st4(P1 + (size_t)(e0 + row) * 3 * D + D + col, P);
                }
            }
            csync();
            TC_TL(9);
            // ---- xa_i = sum_e m_e ----
            {
                const int i_first = sh.meta.dst[0], i_last = sh.meta.dst[nvalid - 1];
                for (int i = i_first + grp; i <= i_last; i += TC2_NGRP) {
                    const int q0 = ws.rowptr[i], q1 = ws.rowptr[i + 1];
                    const int lo = max(q0, e0) - e0, hi = min(q1, e0 + nvalid) - e0;
                    float xa = 0.f;
                    for (int r = lo; r < hi; r++) xa += sh.tile[r][cch];
                    if (q0 >= e0 && q1 <= e0 + nvalid) ws.XA[(size_t)i * D + cch] = xa;
                    else atomicAdd(ws.XA + (size_t)i * D + cch, xa);
                }
            }
            TC_TL(10);
            // ---- A = m, start s1 (-> D1) ----
            if (upd) { tc::mbar_wait(&sh.done[J_F], tpar); tc::fence_after_sync(); }   // A planes no longer read
            tc2_tile_to_a<ROWS>(sh, tmem, warp, lane);
            tc2_go(sh, J_S1);
            TC_TL(11);
            // ---- edge update from the f chunk (D0) ----
            if (upd) {
                csync();                                          // m tile fully consumed (xa + A copy)
                tc2_d_to_tile<ROWS>(sh, tmem, TC_COL_D0, warp, lane);
                tc::fence_before_sync();
                csync();
                const float4 bb = ldg4(lw.b1 + 2 * D + col);
#pragma unroll 1
                for (int rb = 0; rb < RPW; rb += 2) {       // gathers of 2 rows in flight before the first global store
                    if (rb >= rpw) break;
                    float4 tir[2][3], ujr[2][3], fin[2];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int row = r0 + rb + u;
                        const size_t i3 = (size_t)sh.meta.dst[row] * 3, j3 = (size_t)sh.meta.src[row] * 3;
                        fin[u] = (rb + u < rpw && row < nvalid) ? ldg4(Fin + (size_t)(e0 + row) * D + col) : f4s(0.f);
#pragma unroll
                        for (int s = 0; s < 3; s++) {
                            tir[u][s] = ldg4(TU + (i3 + s) * 2 * D + col);
                            ujr[u][s] = ldg4(TU + (j3 + s) * 2 * D + D + col);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int row = r0 + rb + u;
                        const float4 dd = sh.meta.d[row];
                        const float4 Pf = ld4(&sh.tile[row][col]) + bb;
                        const float4 fp = silu4(Pf);
                        const float4 a1 = tir[u][0] * dd.x + tir[u][1] * dd.y + tir[u][2] * dd.z;
                        const float4 a2 = ujr[u][0] * dd.x + ujr[u][1] * dd.y + ujr[u][2] * dd.z;
                        const float4 wdot = (tir[u][0] - a1 * dd.x) * (ujr[u][0] - a2 * dd.x) + (tir[u][1] - a1 * dd.y) * (ujr[u][1] - a2 * dd.y) +
                                            (tir[u][2] - a1 * dd.z) * (ujr[u][2] - a2 * dd.z);
                        if ((rb + u < rpw && row < nvalid)) {
                            st4(P1 + (size_t)(e0 + row) * 3 * D + 2 * D + col, Pf);
                            st4(Fout + (size_t)(e0 + row) * D + col, fin[u] + fp * wdot);
                        }
                    }
                }
            }
            TC_TL(12);
            tc::fence_before_sync();
            tc::mbar_arrive(&sh.go[J_S2]);                        // D0 is free (A = m already published by go[J_S1])
            // ---- s1 (D1): va_i += sum_e vn_j * s1 ----
            float bnd[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
            tc::mbar_wait(&sh.done[J_S1], tpar);
            tc::fence_after_sync();
            TC_TL(13);
            csync();
            tc2_d_to_tile<ROWS>(sh, tmem, TC_COL_D1, warp, lane);
            csync();
            {
                TC_TL(14);
                const float b = __ldg(lw.bs + cch);
                const int i_first = sh.meta.dst[0], i_last = sh.meta.dst[nvalid - 1];
                int nb = 0;
                for (int i = i_first + grp; i <= i_last; i += TC2_NGRP) {
                    const int q0 = ws.rowptr[i], q1 = ws.rowptr[i + 1];
                    const int lo = max(q0, e0) - e0, hi = min(q1, e0 + nvalid) - e0;
                    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
                    int r = lo;
                    for (; r + 4 <= hi; r += 4) {              // 12 independent gathers in flight
                        float g[4][3], s1[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const size_t j3 = (size_t)sh.meta.src[r + u] * 3;
                            g[u][0] = __ldg(VN + (j3 + 0) * D + cch); g[u][1] = __ldg(VN + (j3 + 1) * D + cch); g[u][2] = __ldg(VN + (j3 + 2) * D + cch);
                            const float sp = sh.tile[r + u][cch] + b;
                            SP[(size_t)(e0 + r + u) * 2 * D + cch] = sp;
                            s1[u] = silu_(sp);
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) { v0 += g[u][0] * s1[u]; v1 += g[u][1] * s1[u]; v2 += g[u][2] * s1[u]; }
                    }
                    for (; r < hi; r++) {
                        const size_t j3 = (size_t)sh.meta.src[r] * 3;
                        const float sp = sh.tile[r][cch] + b;
                        SP[(size_t)(e0 + r) * 2 * D + cch] = sp;
                        const float s1 = silu_(sp);
                        v0 += __ldg(VN + (j3 + 0) * D + cch) * s1;
                        v1 += __ldg(VN + (j3 + 1) * D + cch) * s1;
                        v2 += __ldg(VN + (j3 + 2) * D + cch) * s1;
                    }
                    if (q0 >= e0 && q1 <= e0 + nvalid) {
                        ws.VA[((size_t)i * 3 + 0) * D + cch] = v0;
                        ws.VA[((size_t)i * 3 + 1) * D + cch] = v1;
                        ws.VA[((size_t)i * 3 + 2) * D + cch] = v2;
                    } else if (nb < 2) {
                        bnd[nb][0] = v0; bnd[nb][1] = v1; bnd[nb][2] = v2;
                        nb++;
                    }
                }
            }
            TC_TL(15);
            if (threadIdx.x == 0 && it + 1 < my_tiles) {             // next tile's feature rows -> L2 (bulk prefetch), shortly before use
                const int en = ((int)blockIdx.x + (it + 1) * (int)gridDim.x) * trows;
                tc::tma_prefetch_l2(Fin + (size_t)en * D, (uint32_t)min(trows, E - en) * D * 4);
            }
            // ---- s2 (D0): va_i += sum_e s2 * d ----
            tc::mbar_wait(&sh.done[J_S2], tpar);
            tc::fence_after_sync();
            TC_TL(16);
            csync();
            tc2_d_to_tile<ROWS>(sh, tmem, TC_COL_D0, warp, lane);
            tc::fence_before_sync();
            csync();
            {
                TC_TL(17);
                const float b = __ldg(lw.bs + D + cch);
                const int i_first = sh.meta.dst[0], i_last = sh.meta.dst[nvalid - 1];
                int nb = 0;
                for (int i = i_first + grp; i <= i_last; i += TC2_NGRP) {
                    const int q0 = ws.rowptr[i], q1 = ws.rowptr[i + 1];
                    const int lo = max(q0, e0) - e0, hi = min(q1, e0 + nvalid) - e0;
                    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
                    for (int r = lo; r < hi; r++) {
                        const float4 de = sh.meta.d[r];
                        const float sp = sh.tile[r][cch] + b;
                        SP[(size_t)(e0 + r) * 2 * D + D + cch] = sp;
                        const float s2 = silu_(sp);
                        v0 += s2 * de.x; v1 += s2 * de.y; v2 += s2 * de.z;
                    }
                    if (q0 >= e0 && q1 <= e0 + nvalid) {
                        ws.VA[((size_t)i * 3 + 0) * D + cch] += v0;
                        ws.VA[((size_t)i * 3 + 1) * D + cch] += v1;
                        ws.VA[((size_t)i * 3 + 2) * D + cch] += v2;
                    } else if (nb < 2) {
                        atomicAdd(ws.VA + ((size_t)i * 3 + 0) * D + cch, bnd[nb][0] + v0);
                        atomicAdd(ws.VA + ((size_t)i * 3 + 1) * D + cch, bnd[nb][1] + v1);
                        atomicAdd(ws.VA + ((size_t)i * 3 + 2) * D + cch, bnd[nb][2] + v2);
                        nb++;
                    }
                }
            }
            TC_TL(18);
            csync();                                              // tile / meta free for the next tile
        }
    }
    tc2_teardown(tmem);
    if (a.tl != nullptr && blockIdx.x == 0 && threadIdx.x == 0) a.tl[31] = (unsigned long long)clock64();
}

}  // namespace vb

namespace vb {

// ---------------------------------------------------------------------------------------------
// adjoint on tensor cores (math and reference lines: see edge_bwd_kernel in k_edge.cuh).  Pre-activations come
// from the forward stage (P1, SP, ATT), so the tile runs only the two adjoint contractions:
// jobs (upd):  0 g3a -> D1   1 g3b -> D1(+)   2 g4dv -> D0   3 g4dk -> D0(+)   4 g4f -> D0(+)
// ---------------------------------------------------------------------------------------------
template <int ROWS>
__global__ void __launch_bounds__(TC2_THREADS, 1) edge_bwd_tc_kernel(const __grid_constant__ EdgeTcArgs a) {
    pdl_entry();
    extern __shared__ __align__(1024) uint8_t dyn_raw[];
    TcShared& sh = *tc_shared_base(dyn_raw);
    const Workspace& ws = a.ws;
    const int l = a.layer;
    const bool upd = (l < L - 1);
    const int J_G3A = 0, J_G3B = 1, J_G4DV = 2, J_G4DK = 3, J_G4F = 4;
    const int J_LAST = upd ? J_G4F : J_G4DK;
    constexpr int RPW = ROWS / TC2_CWARPS;
    constexpr int RB4 = (RPW % 4 == 0) ? 4 : 2;     // rows whose loads are issued together
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, col = lane * 4, hd = lane >> 2;
    const int E = ws.rowptr[ws.N];
    const int trows = min(ROWS, max(16, a.tile_rows));          // edges per tile (<= ROWS, chosen on the host so the tiles fill whole waves)
    const int ntiles_total = (E + trows - 1) / trows;
    const int my_tiles = ((int)blockIdx.x < ntiles_total) ? (ntiles_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (a.tl != nullptr && blockIdx.x == 0 && threadIdx.x == 0) a.tl[0] = (unsigned long long)clock64();
    const uint32_t tmem = tc2_setup(sh, a.njobs);
    if (a.tl != nullptr && blockIdx.x == 0 && threadIdx.x == 0) a.tl[1] = (unsigned long long)clock64();

    if (warp == TC2_CWARPS) {
        if (lane == 0) tc_producer(sh, a.jobs, a.njobs, my_tiles);
    } else if (warp == TC2_CWARPS + 1) {
        if (lane == 0) tc_mma_issuer(sh, a.jobs, a.njobs, my_tiles, tmem, blockIdx.x == 0 ? a.tl : nullptr);
    } else {
        const float* __restrict__ QKV = ws.QKV[l];
        const float* __restrict__ VN = ws.VN[l];
        const float* __restrict__ TU = ws.TU[l];
        const float* __restrict__ P1 = ws.P1[l];
        const float* __restrict__ SP = ws.SP[l];
        const float* __restrict__ ATT = ws.ATT[l];
        const int cch = threadIdx.x & (D - 1), grp = threadIdx.x >> 7;
        auto wait_done = [&](int j, uint32_t tpar) { tc::mbar_wait(&sh.done[j], tpar); tc::fence_after_sync(); };
        for (int it = 0; it < my_tiles; it++) {
            const uint32_t tpar = (uint32_t)(it & 1);
            const int e0 = ((int)blockIdx.x + it * (int)gridDim.x) * trows;
            const int nvalid = min(trows, E - e0);
            // rows are dealt to the compute warps in contiguous runs of rpw = ceil(nvalid / 16): a tile shorter than ROWS
            // keeps every warp busy (slot s of a warp is row warp * rpw + s, valid while s < rpw and the row exists)
            const int rpw = (nvalid + TC2_CWARPS - 1) / TC2_CWARPS, r0 = warp * rpw;
            load_edge_meta<TC_TE, TC2_CTHREADS>(sh.meta, ws, e0, nvalid);
            csync();
            TC_TL(2);
            // ---- s1 half: g_Spre[:, 0:128] -> tile -> A ; source-side g_vn ----
            // (loads of 4 rows are issued together: the atomics below are compiler barriers for load hoisting)
#pragma unroll 1
            for (int rb = 0; rb < RPW; rb += RB4) {
                if (rb >= rpw) break;
                float4 sp[RB4], gM[RB4][3], vn[RB4][3];
#pragma unroll
                for (int u = 0; u < RB4; u++) {
                    const int row = r0 + rb + u;
                    const size_t e = (size_t)(e0 + ((rb + u < rpw && row < nvalid) ? row : 0));
                    const size_t i3 = (size_t)sh.meta.dst[row] * 3, j3 = (size_t)sh.meta.src[row] * 3;
                    sp[u] = ldg4(SP + e * 2 * D + col);
#pragma unroll
                    for (int s = 0; s < 3; s++) { gM[u][s] = ldg4(ws.GVEC + (i3 + s) * D + col); vn[u][s] = ldg4(VN + (j3 + s) * D + col); }
                }
#pragma unroll
                for (int u = 0; u < RB4; u++) {
                    const int row = r0 + rb + u;
                    const bool ok = (rb + u < rpw && row < nvalid);
                    const size_t j3 = (size_t)sh.meta.src[row] * 3;
                    const float4 s1 = silu4(sp[u]);
                    const float4 gs1 = gM[u][0] * vn[u][0] + gM[u][1] * vn[u][1] + gM[u][2] * vn[u][2];
                    if (rb + u < rpw) st4(&sh.tile[row][col], ok ? gs1 * dsilu4(sp[u]) : f4s(0.f));   // (a slot past the run is the next warp's row)
                    if (ok) {
                        red4(ws.GVNMSG + (j3 + 0) * D + col, gM[u][0] * s1);
                        red4(ws.GVNMSG + (j3 + 1) * D + col, gM[u][1] * s1);
                        red4(ws.GVNMSG + (j3 + 2) * D + col, gM[u][2] * s1);
                    }
                }
            }
            TC_TL(3);
            csync();
            tc2_tile_to_a<ROWS>(sh, tmem, warp, lane);
            tc2_go(sh, J_G3A);
            TC_TL(4);
            csync();
            // ---- s2 half ----
#pragma unroll 4
            for (int r = 0; r < RPW; r++) {
                if (r >= rpw) break;
                const int row = r0 + r;
                const bool ok = (r < rpw && row < nvalid);
                const size_t e = (size_t)(e0 + (ok ? row : 0));
                const size_t i3 = (size_t)sh.meta.dst[row] * 3;
                const float4 dd = sh.meta.d[row];
                const float4 sp = ldg4(SP + e * 2 * D + D + col);
                const float4 s2 = silu4(sp);
                const float4 gM0 = ldg4(ws.GVEC + (i3 + 0) * D + col), gM1 = ldg4(ws.GVEC + (i3 + 1) * D + col),
                             gM2 = ldg4(ws.GVEC + (i3 + 2) * D + col);
                const float gx_ = warp_sum(hsum4(gM0 * s2)), gy_ = warp_sum(hsum4(gM1 * s2)), gz_ = warp_sum(hsum4(gM2 * s2));
                if (lane == 0) { sh.eacc[row][1] = gx_; sh.eacc[row][2] = gy_; sh.eacc[row][3] = gz_; }
                st4(&sh.tile[row][col], ok ? (gM0 * dd.x + gM1 * dd.y + gM2 * dd.z) * dsilu4(sp) : f4s(0.f));
            }
            TC_TL(5);
            csync();
            wait_done(J_G3A, tpar);
            TC_TL(6);
            tc2_tile_to_a<ROWS>(sh, tmem, warp, lane);
            tc2_go(sh, J_G3B);
            TC_TL(7);
            // ---- g_m = g_xa_i + g_Spre Ws ; adjoint of m = v_j dv A ----
            wait_done(J_G3B, tpar);
            TC_TL(8);
            csync();
            tc2_d_to_tile<ROWS>(sh, tmem, TC_COL_D1, warp, lane);
            tc::fence_before_sync();
            csync();
            TC_TL(9);
#pragma unroll 1
            for (int rb = 0; rb < RPW; rb += RB4) {
                if (rb >= rpw) break;
                float4 gxa[RB4], vjr[RB4], pdvr[RB4];
                float avr[RB4];
#pragma unroll
                for (int u = 0; u < RB4; u++) {
                    const int row = r0 + rb + u;
                    const size_t e = (size_t)(e0 + ((rb + u < rpw && row < nvalid) ? row : 0));
                    gxa[u] = load_gxa(ws, (size_t)sh.meta.dst[row], col);
                    vjr[u] = ldg4(QKV + (size_t)sh.meta.src[row] * 3 * D + 2 * D + col);
                    pdvr[u] = ldg4(P1 + e * 3 * D + D + col);
                    avr[u] = (rb + u < rpw && row < nvalid) ? __ldg(ATT + e * H + hd) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < RB4; u++) {
                    const int row = r0 + rb + u;
                    const bool ok = (rb + u < rpw && row < nvalid);
                    const size_t j = sh.meta.src[row];
                    const float Ce = sh.meta.C[row];
                    const float av = avr[u], sa = silu_(av), A = sa * Ce;
                    const bool mine = rb + u < rpw;                 // a slot past the run is the next warp's row: no shared accesses
                    const float4 gm = mine ? ld4(&sh.tile[row][col]) + gxa[u] : f4s(0.f);   // (its owner rewrites it in this phase)
                    const float4 dv = silu4(pdvr[u]);
                    if (mine) st4(&sh.tile[row][col], ok ? gm * vjr[u] * A * dsilu4(pdvr[u]) : f4s(0.f));      // g_Pdv
                    const float gA = quad_sum(hsum4(gm * vjr[u] * dv));
                    if (mine && (lane & 3) == 0) sh.gattn[row][hd] = gA * Ce * dsilu_(av);
                    const float gc = warp_sum((lane & 3) == 0 ? gA * sa : 0.f);
                    if (mine && lane == 0) sh.eacc[row][0] = gc;
                    if (ok) red4(ws.GQKV + j * 3 * D + 2 * D + col, gm * dv * A);
                }
            }
            TC_TL(10);
            csync();
            tc2_tile_to_a<ROWS>(sh, tmem, warp, lane);               // A = g_Pdv (A planes free: g3b done)
            tc2_go(sh, J_G4DV);
            TC_TL(11);
            csync();
            // ---- adjoint of a_h = sum q_i k_j dk : first g_Pdk (next A operand), then the g_q tile ----
#pragma unroll 1
            for (int rb = 0; rb < RPW; rb += RB4) {
                if (rb >= rpw) break;
                float4 pdkr[RB4], qir[RB4], kjr[RB4];
#pragma unroll
                for (int u = 0; u < RB4; u++) {
                    const int row = r0 + rb + u;
                    const size_t e = (size_t)(e0 + ((rb + u < rpw && row < nvalid) ? row : 0));
                    pdkr[u] = ldg4(P1 + e * 3 * D + col);
                    qir[u] = ldg4(QKV + (size_t)sh.meta.dst[row] * 3 * D + col);
                    kjr[u] = ldg4(QKV + (size_t)sh.meta.src[row] * 3 * D + D + col);
                }
#pragma unroll
                for (int u = 0; u < RB4; u++) {
                    const int row = r0 + rb + u;
                    const bool ok = (rb + u < rpw && row < nvalid);
                    const size_t j = sh.meta.src[row];
                    const float4 dk = silu4(pdkr[u]);
                    const float gav = sh.gattn[row][hd];
                    if (rb + u < rpw) st4(&sh.tile[row][col], ok ? qir[u] * kjr[u] * gav * dsilu4(pdkr[u]) : f4s(0.f));   // g_Pdk
                    if (ok) red4(ws.GQKV + j * 3 * D + D + col, qir[u] * dk * gav);
                }
            }
            TC_TL(12);
            csync();
            wait_done(J_G4DV, tpar);
            TC_TL(13);
            tc2_tile_to_a<ROWS>(sh, tmem, warp, lane);               // A = g_Pdk
            tc2_go(sh, J_G4DK);
            TC_TL(14);
            csync();
#pragma unroll 4
            for (int r = 0; r < RPW; r++) {
                if (r >= rpw) break;
                const int row = r0 + r;
                const size_t e = (size_t)(e0 + ((r < rpw && row < nvalid) ? row : 0));
                const float4 dk = silu4(ldg4(P1 + e * 3 * D + col));
                const float4 kj = ldg4(QKV + (size_t)sh.meta.src[row] * 3 * D + D + col);
                st4(&sh.tile[row][col], kj * dk * sh.gattn[row][hd]);                    // per-edge g_q contribution
            }
            csync();
            {
                const int i_first = sh.meta.dst[0], i_last = sh.meta.dst[nvalid - 1];
                for (int i = i_first + grp; i <= i_last; i += TC2_NGRP) {
                    const int q0 = ws.rowptr[i], q1 = ws.rowptr[i + 1];
                    const int lo = max(q0, e0) - e0, hi = min(q1, e0 + nvalid) - e0;
                    float gq = 0.f;
                    for (int r = lo; r < hi; r++) gq += sh.tile[r][cch];
                    if (q0 >= e0 && q1 <= e0 + nvalid) ws.GQKV[(size_t)i * 3 * D + cch] = gq;
                    else atomicAdd(ws.GQKV + (size_t)i * 3 * D + cch, gq);
                }
            }
            TC_TL(15);
            // ---- adjoint of the edge update: first g_Pf (A operand), then the g_wdot tile ----
            if (upd) {
                csync();
#pragma unroll 1
                for (int rb = 0; rb < RPW; rb += 2) {
                    if (rb >= rpw) break;
                    float4 gfr[2], pfr[2], tir[2][3], ujr[2][3];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int row = r0 + rb + u;
                        const bool ok = (rb + u < rpw && row < nvalid);
                        const size_t e = (size_t)(e0 + (ok ? row : 0));
                        const size_t i3 = (size_t)sh.meta.dst[row] * 3, j3 = (size_t)sh.meta.src[row] * 3;
                        gfr[u] = ok ? ld4(ws.GF + e * D + col) : f4s(0.f);
                        pfr[u] = ldg4(P1 + e * 3 * D + 2 * D + col);
#pragma unroll
                        for (int s = 0; s < 3; s++) {
                            tir[u][s] = ldg4(TU + (i3 + s) * 2 * D + col);
                            ujr[u][s] = ldg4(TU + (j3 + s) * 2 * D + D + col);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int row = r0 + rb + u;
                        const bool ok = (rb + u < rpw && row < nvalid);
                        const size_t e = (size_t)(e0 + (ok ? row : 0));
                        const size_t j3 = (size_t)sh.meta.src[row] * 3;
                        const float4 dd = sh.meta.d[row];
                        const float4 gfn = gfr[u], pf = pfr[u];
                        const float4 fp = silu4(pf);
                        const float dv3[3] = {dd.x, dd.y, dd.z};
                        const float4 a1 = tir[u][0] * dd.x + tir[u][1] * dd.y + tir[u][2] * dd.z;
                        const float4 a2 = ujr[u][0] * dd.x + ujr[u][1] * dd.y + ujr[u][2] * dd.z;
                        float4 w1[3], w2[3];
#pragma unroll
                        for (int s = 0; s < 3; s++) { w1[s] = tir[u][s] - a1 * dv3[s]; w2[s] = ujr[u][s] - a2 * dv3[s]; }
                        const float4 wdot = w1[0] * w2[0] + w1[1] * w2[1] + w1[2] * w2[2];
                        const float4 gwd = gfn * fp;
                        if (rb + u < rpw) st4(&sh.tile[row][col], gfn * wdot * dsilu4(pf));                    // g_Pf
                        const float4 c1 = gwd * (w2[0] * dd.x + w2[1] * dd.y + w2[2] * dd.z);
                        const float4 c2 = gwd * (w1[0] * dd.x + w1[1] * dd.y + w1[2] * dd.z);
                        float gdl[3];
                        float4 gu[3];
#pragma unroll
                        for (int s = 0; s < 3; s++) {
                            const float4 gw1 = gwd * w2[s], gw2 = gwd * w1[s];
                            gu[s] = gw2 - c2 * dv3[s];
                            gdl[s] = warp_sum(hsum4(tir[u][s] * c1 + a1 * gw1 + ujr[u][s] * c2 + a2 * gw2));
                        }
                        if (lane == 0 && rb + u < rpw) { sh.eacc[row][1] -= gdl[0]; sh.eacc[row][2] -= gdl[1]; sh.eacc[row][3] -= gdl[2]; }
                        if (ok) {
                            red4(ws.GTU + (j3 + 0) * 2 * D + D + col, gu[0]);
                            red4(ws.GTU + (j3 + 1) * 2 * D + D + col, gu[1]);
                            red4(ws.GTU + (j3 + 2) * 2 * D + D + col, gu[2]);
                        }
                        (void)e;
                    }
                }
                TC_TL(16);
                csync();
                wait_done(J_G4DK, tpar);
                TC_TL(17);
                tc2_tile_to_a<ROWS>(sh, tmem, warp, lane);           // A = g_Pf
                tc2_go(sh, J_G4F);
                TC_TL(18);
                csync();
#pragma unroll 4
                for (int r = 0; r < RPW; r++) {
                    if (r >= rpw) break;
                    const int row = r0 + r;
                    const bool ok = (r < rpw && row < nvalid);
                    const size_t e = (size_t)(e0 + (ok ? row : 0));
                    const float4 gfn = ok ? ld4(ws.GF + e * D + col) : f4s(0.f);
                    st4(&sh.tile[row][col], gfn * silu4(ldg4(P1 + e * 3 * D + 2 * D + col)));   // g_wdot
                }
                csync();
                {
                    const int i_first = sh.meta.dst[0], i_last = sh.meta.dst[nvalid - 1];
                    for (int i = i_first + grp; i <= i_last; i += TC2_NGRP) {
                        const int q0 = ws.rowptr[i], q1 = ws.rowptr[i + 1];
                        const int lo = max(q0, e0) - e0, hi = min(q1, e0 + nvalid) - e0;
                        float gt0 = 0.f, gt1 = 0.f, gt2 = 0.f;
                        auto term = [&](int r, float u0, float u1, float u2) {
                            const float4 dd = sh.meta.d[r];
                            const float gw = sh.tile[r][cch];
                            const float a2 = u0 * dd.x + u1 * dd.y + u2 * dd.z;
                            const float w20 = u0 - a2 * dd.x, w21 = u1 - a2 * dd.y, w22 = u2 - a2 * dd.z;
                            const float wd = w20 * dd.x + w21 * dd.y + w22 * dd.z;
                            gt0 += gw * (w20 - wd * dd.x);
                            gt1 += gw * (w21 - wd * dd.y);
                            gt2 += gw * (w22 - wd * dd.z);
                        };
                        int r = lo;
                        for (; r + 4 <= hi; r += 4) {              // 12 independent gathers in flight
                            float u[4][3];
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const size_t j3 = (size_t)sh.meta.src[r + q] * 3;
                                u[q][0] = __ldg(TU + (j3 + 0) * 2 * D + D + cch);
                                u[q][1] = __ldg(TU + (j3 + 1) * 2 * D + D + cch);
                                u[q][2] = __ldg(TU + (j3 + 2) * 2 * D + D + cch);
                            }
#pragma unroll
                            for (int q = 0; q < 4; q++) term(r + q, u[q][0], u[q][1], u[q][2]);
                        }
                        for (; r < hi; r++) {
                            const size_t j3 = (size_t)sh.meta.src[r] * 3;
                            term(r, __ldg(TU + (j3 + 0) * 2 * D + D + cch), __ldg(TU + (j3 + 1) * 2 * D + D + cch),
                                 __ldg(TU + (j3 + 2) * 2 * D + D + cch));
                        }
                        if (q0 >= e0 && q1 <= e0 + nvalid) {
                            ws.GTU[((size_t)i * 3 + 0) * 2 * D + cch] = gt0;
                            ws.GTU[((size_t)i * 3 + 1) * 2 * D + cch] = gt1;
                            ws.GTU[((size_t)i * 3 + 2) * 2 * D + cch] = gt2;
                        } else {
                            atomicAdd(ws.GTU + ((size_t)i * 3 + 0) * 2 * D + cch, gt0);
                            atomicAdd(ws.GTU + ((size_t)i * 3 + 1) * 2 * D + cch, gt1);
                            atomicAdd(ws.GTU + ((size_t)i * 3 + 2) * 2 * D + cch, gt2);
                        }
                    }
                }
            }
            TC_TL(19);
            // next tile's stored pre-activations -> L2 (bulk prefetch, UBLKPF).  Issued late in the tile: a whole tile ahead
            // the rows were evicted again before their use (ncu: DRAM reads 1.0 -> 1.7 GB per launch on the 512-fragment batch)
            if (threadIdx.x < 3 && it + 1 < my_tiles) {
                const int en = ((int)blockIdx.x + (it + 1) * (int)gridDim.x) * trows;
                const uint32_t nn = (uint32_t)min(trows, E - en);
                if (threadIdx.x == 0) tc::tma_prefetch_l2(SP + (size_t)en * 2 * D, nn * 2 * D * 4);
                else if (threadIdx.x == 1) tc::tma_prefetch_l2(P1 + (size_t)en * 3 * D, nn * 3 * D * 4);
                else tc::tma_prefetch_l2(ATT + (size_t)en * H, nn * H * 4);
            }
            // ---- g_f = g_f_next + [g_Pdk|g_Pdv|g_Pf] W1 ----
            wait_done(J_LAST, tpar);
            TC_TL(20);
            csync();
            tc2_d_to_tile<ROWS>(sh, tmem, TC_COL_D0, warp, lane);
            tc::fence_before_sync();
            csync();
            TC_TL(21);
#pragma unroll 4
            for (int r = 0; r < RPW; r++) {
                if (r >= rpw) break;
                const int row = r0 + r;
                if ((r < rpw && row < nvalid)) {
                    float* g = ws.GF + (size_t)(e0 + row) * D + col;
                    float4 v = ld4(&sh.tile[row][col]);
                    if (upd) v = v + ld4(g);
                    st4(g, v);
                }
            }
            if (threadIdx.x < nvalid) {
                float* ea = ws.eacc + (size_t)(e0 + threadIdx.x) * 4;
                st4(ea, ld4(ea) + ld4(&sh.eacc[threadIdx.x][0]));
            }
            TC_TL(22);
            csync();
        }
    }
    tc2_teardown(tmem);
    if (a.tl != nullptr && blockIdx.x == 0 && threadIdx.x == 0) a.tl[31] = (unsigned long long)clock64();
}

}  // namespace vb
