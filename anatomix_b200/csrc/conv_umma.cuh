// 3x3x3 "valid" convolution over a reflect-padded planar bf16 buffer as an
// implicit GEMM on the sm_100a tensor cores (tcgen05.mma, accumulators in TMEM).
//
//   M = 128 voxels  = one 8(x) x 16(y) patch of one z plane  (TMEM lane = voxel)
//   N = Cout (x3 when the three dz taps are folded into one instruction)
//   K = 16 input channels per instruction, 27 taps x Cin/16 chunks per tile
//
// Per CTA tile (8 x 16 x bz voxels) and per 16-channel chunk, ONE TMA box copy
// brings the (bz+2) x 18 x 10 halo brick of both 8-channel groups into shared
// memory.  Because the planar layout stores a voxel's 8 channels as 16
// contiguous bytes and x-neighbours contiguously, the brick is already the
// canonical K-major non-swizzled operand: the A descriptor of tap (dz,dy,dx) is
// just the brick base advanced by ((dz*18+dy)*10+dx)*16 bytes, with SBO = 160
// (next y row = next group of 8 voxels) and LBO = the distance between the two
// channel groups.  No im2col copy exists anywhere, in HBM or in shared memory.
//
// dz folding (Cout <= 80): out[z] = sum_dz in[z+dz] * W[dz].  Input plane j of
// the brick contributes to output planes j-2, j-1, j (tile-local) with weights
// dz = +1, 0, -1.  Output plane b owns TMEM columns [b*ncols, (b+1)*ncols), so
// ONE instruction with N = 3*ncols and B = [W(+1) | W(0) | W(-1)] stacked along N,
// targeted at column (j-2)*ncols, accumulates all three at once.  That cuts the
// shared-memory reads of A (the bottleneck for thin layers: a 4 KB A tile feeds
// only N columns) by 3x.  The epilogue warps re-initialise the accumulators with
// the per-channel shift (folded BatchNorm / conv bias) right after draining them,
// so every MMA accumulates, the overlap needs no first-touch case and the
// epilogue needs no bias add.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA
// issuer (warp-uniform loop, one elected lane issues), warps 2..9 = epilogue (TMEM
// lane quadrant = warp_id % 4, the two warps of a quadrant take alternate output
// planes).  Rings: A bricks (stage = chunk), B weight slabs (stage = chunk x
// dz-group), TMEM accumulator stages.  Persistent CTAs stride over tiles.
#pragma once
#include "umma_epilogue.cuh"

namespace anx {

#ifndef ANX_EPI_WARPS
#define ANX_EPI_WARPS 8
#endif
constexpr int EPI_WARPS = ANX_EPI_WARPS;     // generic conv kernel: multiple of 4 (one or more warps per TMEM lane quadrant)
constexpr int STEM_EPI_WARPS = 8;
constexpr int UMMA_THREADS = 64 + 32 * EPI_WARPS;
constexpr int MAX_A_STAGES = 4;
constexpr int MAX_B_STAGES = 8;

struct UmmaShared {          // lives behind the A / B rings in dynamic shared memory
    uint64_t full_a[MAX_A_STAGES], empty_a[MAX_A_STAGES];
    uint64_t full_b[MAX_B_STAGES], empty_b[MAX_B_STAGES];
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_slot;
    uint32_t pad[3];
    float shift[1024];       // per-channel accumulator seed (folded BN shift / bias), all splits
};
// EPI_STATS on thin layers (ConvGeom::stats_acc): per-warp instance-norm sums of the current sample, in an extra
// region right behind UmmaShared that only such launches allocate
constexpr uint32_t STATS_ACC_BYTES = EPI_WARPS * 2 * STATS_ACC_MAX_COLS * sizeof(double);

struct TileCoord { int n, z0, y0, x0, split; };
__device__ __forceinline__ TileCoord decode_tile(const ConvGeom &g, int tile) {
    TileCoord t;
    t.split = tile % g.n_splits;       // consecutive CTAs share one brick and take different channel splits
    tile /= g.n_splits;
    t.n = tile / g.tiles_per_sample;
    int r = tile - t.n * g.tiles_per_sample;
    int tz = r / (g.tiles_y * g.tiles_x);
    r -= tz * g.tiles_y * g.tiles_x;
    int ty = r / g.tiles_x;
    t.z0 = tz * g.bz;
    t.y0 = ty * TILE_Y;
    t.x0 = (r - ty * g.tiles_x) * TILE_X;
    return t;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t hi, uint32_t lo) {
    return ((uint64_t)hi << 32) | lo;
}

// SEEDED: accumulators start from a stored partial-sum tensor (ep.seed_src) instead of the channel
// shift -- the skip half of a decoder conv whose upsampled half was evaluated at low resolution.
template <int MODE>   // EpiMode: the epilogue variant compiled into this instantiation
__global__ void __launch_bounds__(UMMA_THREADS, 1)
conv3_umma_kernel(const __grid_constant__ CUtensorMap tmap_in, const ConvGeom g, const uint8_t *__restrict__ wpack,
                  const Epilogue ep) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *a_ring = smem;
    uint8_t *b_ring = a_ring + (size_t)g.a_stages * g.a_stage_bytes;
    UmmaShared *sh = reinterpret_cast<UmmaShared *>(b_ring + (size_t)g.b_stages * g.b_stage_bytes);

    constexpr bool SEEDED = MODE == EPI_SEEDED;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const int acc_cols = g.bz * g.ncols;   // TMEM columns per accumulator stage

    if (threadIdx.x == 0) {
        for (int i = 0; i < g.a_stages; ++i) { mbar_init(&sh->full_a[i], 1); mbar_init(&sh->empty_a[i], 1); }
        for (int i = 0; i < g.b_stages; ++i) { mbar_init(&sh->full_b[i], 1); mbar_init(&sh->empty_b[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&sh->tmem_full[i], 1); mbar_init(&sh->tmem_empty[i], 32 * EPI_WARPS); }
        fence_barrier_init();
        tma_prefetch_desc(&tmap_in);
    }
    if (warp == 1) {
        tmem_alloc_dyn(&sh->tmem_slot, (uint32_t)g.tmem_cols);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < g.ncols * g.n_splits; i += blockDim.x) sh->shift[i] = ep.bias[i];
    if constexpr (MODE == EPI_F32_HEAD)   // ncols <= 32 here: the head sits behind the channel shift
        for (int i = threadIdx.x; i < HEAD_FLOATS; i += blockDim.x) sh->shift[HEAD_SMEM_OFFSET + i] = ep.head[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sh->tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ producer
        if (lane == 0) {
            uint32_t ka = 0, kb = 0;
            if (g.b_static && blockIdx.x < g.total_tiles) {   // the weights of this layer: resident for the whole launch
                if (ANX_ABL(g, 8)) {
                    mbar_arrive(&sh->full_b[0]);
                } else {
                    mbar_arrive_expect_tx(&sh->full_b[0], g.b_stage_bytes);
                    for (uint32_t off = 0; off < g.b_stage_bytes; off += 32768u) {      // bulk copies of at most 32 KB
                        const uint32_t len = g.b_stage_bytes - off < 32768u ? g.b_stage_bytes - off : 32768u;
                        bulk_load_1d(b_ring + off, wpack + off, len, &sh->full_b[0]);
                    }
                }
            }
            for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
                const TileCoord t = decode_tile(g, tile);
                for (int c = 0; c < g.cin_chunks; ++c) {
                    const uint32_t sa = ka % g.a_stages;
                    mbar_wait(&sh->empty_a[sa], ((ka / g.a_stages) & 1) ^ 1, 1);
                    if (ANX_ABL(g, 4)) {
                        mbar_arrive(&sh->full_a[sa]);
                    } else {
                        mbar_arrive_expect_tx(&sh->full_a[sa], g.a_stage_bytes);
                        tma_load_4d(a_ring + (size_t)sa * g.a_stage_bytes, &tmap_in, &sh->full_a[sa], t.x0 * 8,
                                    t.y0, t.z0, t.n * g.in_groups_total + g.in_group_offset + 2 * c);
                    }
                    ++ka;
                    if (g.b_static) continue;
                    for (int grp = 0; grp < g.groups; ++grp) {
                        const uint32_t sb = kb % g.b_stages;
                        mbar_wait(&sh->empty_b[sb], ((kb / g.b_stages) & 1) ^ 1, 2);
                        if (ANX_ABL(g, 8)) {
                            mbar_arrive(&sh->full_b[sb]);
                        } else {
                            mbar_arrive_expect_tx(&sh->full_b[sb], g.b_stage_bytes);
                            bulk_load_1d(b_ring + (size_t)sb * g.b_stage_bytes,
                                         wpack + ((size_t)(t.split * g.cin_chunks + c) * g.groups + grp) * g.b_stage_bytes,
                                         g.b_stage_bytes,
                                         &sh->full_b[sb]);
                        }
                        ++kb;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------------------------------------- MMA issuer (warp-uniform)
        uint32_t ka = 0, kb = 0, it = 0;
        const uint32_t R = g.b_rows;
        const uint32_t a_hi = (ROW_BYTES >> 4) | (1u << 14);     // SBO = 160 B, descriptor version 1
        const uint32_t b_hi = (128u >> 4) | (1u << 14);          // SBO = 128 B
        const uint32_t a_lbo_bits = ((g.a_lbo >> 4) & 0x3FFF) << 16;
        const uint32_t b_lbo_bits = (R & 0x3FFF) << 16;          // LBO = 16 * R bytes
        const uint32_t b_tap = 2 * R;                            // 32 * R bytes per tap, in 16 B units
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++it) {
            const uint32_t s = it % g.acc_stages;
            mbar_wait_warp(&sh->tmem_empty[s], (it / g.acc_stages) & 1, 3);
            tc_fence_after();
            const uint32_t acc = tmem_u + s * acc_cols;
            for (int c = 0; c < g.cin_chunks; ++c) {
                const uint32_t sa = ka % g.a_stages;
                mbar_wait_warp(&sh->full_a[sa], (ka / g.a_stages) & 1, 4);
                const uint32_t a_lo = ((smem_u32(a_ring + (size_t)sa * g.a_stage_bytes) & 0x3FFFF) >> 4) | a_lbo_bits;
                for (int grp = 0; grp < g.groups; ++grp) {
                    const uint32_t sb = g.b_static ? 0u : kb % g.b_stages;
                    // a static slab completes phase 0 once and is never recycled: parity 0 stays satisfied
                    mbar_wait_warp(&sh->full_b[sb], g.b_static ? 0u : (kb / g.b_stages) & 1, 5);
                    tc_fence_after();
                    const uint32_t b_lo = ((smem_u32(b_ring + (size_t)sb * g.b_stage_bytes) & 0x3FFFF) >> 4) | b_lbo_bits;
                    if (ANX_ABL(g, 1)) {
                    } else if (g.fold) {
                        for (int j = 0; j < g.bz + 2; ++j) {
                            const int lo = j - 2 > 0 ? j - 2 : 0;
                            const int hi = j < g.bz - 1 ? j : g.bz - 1;
                            const uint32_t idesc = idesc_m128((uint32_t)(hi - lo + 1) * g.ncols, g.dt);
                            const uint32_t dcol = acc + lo * g.ncols;
                            const uint32_t aj = a_lo + j * (HALO_Y * HALO_X);
                            const uint32_t bj = b_lo + (uint32_t)(lo - (j - 2)) * g.ncols;   // first B row, 16 B each
                            if (ANX_ABL(g, 32)) {   // timing experiment: 128-byte aligned core matrices (SBO = 128 B)
                                const uint32_t a_hi_al = (128u >> 4) | (1u << 14);
                                const uint32_t aj_al = a_lo + j * 176;
#pragma unroll
                                for (int t = 0; t < 9; ++t)
                                    umma_bf16_warp(dcol, make_desc(a_hi_al, aj_al + t * 8), make_desc(b_hi, bj + t * b_tap), idesc);
                                continue;
                            }
                            if (ANX_ABL(g, 16)) {   // timing experiment: every tap reads the aligned brick origin
#pragma unroll
                                for (int t = 0; t < 9; ++t)
                                    umma_bf16_warp(dcol, make_desc(a_hi, a_lo), make_desc(b_hi, bj + t * b_tap), idesc);
                                continue;
                            }
#pragma unroll
                            for (int t = 0; t < 9; ++t)
                                umma_bf16_warp(dcol, make_desc(a_hi, aj + (t / 3) * HALO_X + (t % 3)),
                                               make_desc(b_hi, bj + t * b_tap), idesc);
                        }
                    } else {
                        const uint32_t idesc = idesc_m128(g.ncols, g.dt);
                        for (int b = 0; b < g.bz; ++b) {
                            const uint32_t dcol = acc + b * g.ncols;
                            const uint32_t aj = a_lo + (b + grp) * (HALO_Y * HALO_X);   // grp = kz = dz + 1
                            if (g.trim == 2) { // ... and the compact tiles of every tap resident in shared memory
                                const uint32_t b_base = (smem_u32(b_ring) & 0x3FFFF) >> 4;
#pragma unroll
                                for (int t = 0; t < 9; ++t) {
                                    const uint32_t lo = 16u * g.trim_lo[grp * 9 + t], nn = 16u * g.trim_n[grp * 9 + t];
                                    const uint32_t bt = b_base + g.trim_off[(c * 3 + grp) * 9 + t];
                                    umma_bf16_warp(dcol + lo, make_desc(a_hi, aj + (t / 3) * HALO_X + (t % 3)),
                                                   make_desc(b_hi, bt | ((nn & 0x3FFF) << 16)), idesc_m128(nn, g.dt));
                                }
                                continue;
                            }
                            if (g.trim) {      // structurally sparse B: only the columns this tap can reach
#pragma unroll
                                for (int t = 0; t < 9; ++t) {
                                    const uint32_t lo = 16u * g.trim_lo[grp * 9 + t], nn = 16u * g.trim_n[grp * 9 + t];
                                    umma_bf16_warp(dcol + lo, make_desc(a_hi, aj + (t / 3) * HALO_X + (t % 3)),
                                                   make_desc(b_hi, b_lo + t * b_tap + lo), idesc_m128(nn, g.dt));
                                }
                                continue;
                            }
#pragma unroll
                            for (int t = 0; t < 9; ++t)
                                umma_bf16_warp(dcol, make_desc(a_hi, aj + (t / 3) * HALO_X + (t % 3)),
                                               make_desc(b_hi, b_lo + t * b_tap), idesc);
                        }
                    }
                    if (!g.b_static) umma_commit_warp(&sh->empty_b[sb]);
                    ++kb;
                }
                umma_commit_warp(&sh->empty_a[sa]);
                ++ka;
            }
            umma_commit_warp(&sh->tmem_full[s]);
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue
        const int q = warp & 3;                    // TMEM lane quadrant this warp may touch
        const PlaneSplit half{(warp - 2) >> 2, EPI_WARPS / 4, MODE == EPI_POOL};   // this warp's run of output planes
        const int r = q * 32 + lane;               // accumulator row = voxel within the 8x16 patch
        const int ly = r >> 3, lx = r & 7;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const int chunks = g.ncols >> 4;
        auto epi_tile_of = [&](int tile) {
            const TileCoord t = decode_tile(g, tile);
            EpiTile et;
            et.n = t.n; et.z0 = t.z0; et.y = t.y0 + ly; et.x = t.x0 + lx;
            et.chan0 = t.split * g.ncols;
            et.in_xy = (et.y < g.H) && (et.x < g.W);
            et.store = !(ANX_ABL(g, 2));
            return et;
        };
        // seed every accumulator stage for the first tile that will use it
        for (int s = 0; s < g.acc_stages; ++s) {
            const int tile0 = blockIdx.x + s * (int)gridDim.x;
            if (SEEDED) {
                const bool valid = tile0 < g.total_tiles;
                const EpiTile e0 = epi_tile_of(valid ? tile0 : 0);
                for (int b = plane_lo(half, g.bz); b < plane_hi(half, g.bz); ++b) {
                    uint4 q0, q1;
                    float sv[16];
                    load_seed16(ep, e0, valid, b, g.D, q0, q1);
                    unpack_x8(q0, sv, ep.dt);
                    unpack_x8(q1, sv + 8, ep.dt);
                    tmem_st16(lane_base + s * acc_cols + b * g.ncols, sv);
                }
            } else {
                const int split = tile0 % g.n_splits;
                for (int b = plane_lo(half, g.bz); b < plane_hi(half, g.bz); ++b)
                    for (int cb = 0; cb < chunks; ++cb)
                        tmem_st16(lane_base + s * acc_cols + b * g.ncols + cb * 16, sh->shift + split * g.ncols + cb * 16);
            }
        }
        tmem_wait_st();
        tc_fence_before();
        for (int s = 0; s < g.acc_stages; ++s) mbar_arrive(&sh->tmem_empty[s]);

        // EPI_STATS on thin layers: sums are kept per warp in shared memory and flushed once per sample
        double *my_acc = nullptr;
        int acc_n = -1;
        if constexpr (MODE == EPI_STATS) {
            if (g.stats_acc) {
                my_acc = reinterpret_cast<double *>(sh + 1) + (warp - 2) * 2 * STATS_ACC_MAX_COLS;
                for (int i = lane; i < 2 * g.ncols; i += 32) my_acc[i] = 0.0;
                __syncwarp();
            }
        }
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++it) {
            const EpiTile et = epi_tile_of(tile);
            if constexpr (MODE == EPI_STATS) {
                if (my_acc && et.n != acc_n) {
                    if (acc_n >= 0) warp_stats_flush(my_acc, chunks, ep.stats, ep.stats_stride, acc_n);
                    acc_n = et.n;
                }
            }
            const uint32_t s = it % g.acc_stages;
            const int next = tile + g.acc_stages * (int)gridDim.x;   // the tile that reuses this accumulator stage
            const bool next_valid = SEEDED && next < g.total_tiles;
            const EpiTile en = SEEDED ? epi_tile_of(next_valid ? next : tile) : et;
            if (SEEDED) prefetch_seeds(ep, en, next_valid, half, g.bz, g.D);   // in flight during the wait
            mbar_wait(&sh->tmem_full[s], (it / g.acc_stages) & 1, 6);
            tc_fence_after();
            umma_epilogue_tile<MODE>(ep, et, lane_base + s * acc_cols, sh->shift + (next % g.n_splits) * g.ncols, half,
                                       g.bz, g.ncols, g.D, en, next_valid, my_acc);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&sh->tmem_empty[s]);
        }
        if constexpr (MODE == EPI_STATS) {
            if (my_acc && acc_n >= 0) warp_stats_flush(my_acc, chunks, ep.stats, ep.stats_stride, acc_n);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
}

}   // namespace anx
