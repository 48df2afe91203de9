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
// only N columns) by 3x.  Accumulators are zeroed by the epilogue warps after
// they drain them, so every MMA accumulates and the overlap needs no special
// first-touch case.
//
// Warp roles (192 threads): warp 0 = TMA producer (1 lane), warp 1 = TMEM
// allocator + MMA issuer (1 lane), warps 2..5 = epilogue (TMEM lane quadrant =
// warp_id % 4).  Rings: A bricks (stage = chunk), B weight slabs (stage = chunk
// x dz-group), TMEM accumulator stages.  Persistent CTAs stride over tiles.
#pragma once
#include "epilogue.cuh"
#include "ptx.cuh"

namespace anx {

constexpr int UMMA_THREADS = 192;
constexpr int MAX_A_STAGES = 4;
constexpr int MAX_B_STAGES = 8;

struct UmmaBarriers {
    uint64_t full_a[MAX_A_STAGES], empty_a[MAX_A_STAGES];
    uint64_t full_b[MAX_B_STAGES], empty_b[MAX_B_STAGES];
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_slot;
    uint32_t pad;
};

struct TileCoord { int n, z0, y0, x0; };
__device__ __forceinline__ TileCoord decode_tile(const ConvGeom &g, int tile) {
    TileCoord t;
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

__global__ void __launch_bounds__(UMMA_THREADS, 1)
conv3_umma_kernel(const __grid_constant__ CUtensorMap tmap_in, const ConvGeom g, const uint8_t *__restrict__ wpack,
                  const Epilogue ep) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *a_ring = smem;
    uint8_t *b_ring = a_ring + (size_t)g.a_stages * g.a_stage_bytes;
    UmmaBarriers *bars = reinterpret_cast<UmmaBarriers *>(b_ring + (size_t)g.b_stages * g.b_stage_bytes);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int acc_cols = g.bz * g.ncols;   // TMEM columns per accumulator stage

    if (threadIdx.x == 0) {
        for (int i = 0; i < g.a_stages; ++i) { mbar_init(&bars->full_a[i], 1); mbar_init(&bars->empty_a[i], 1); }
        for (int i = 0; i < g.b_stages; ++i) { mbar_init(&bars->full_b[i], 1); mbar_init(&bars->empty_b[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], 128); }
        fence_barrier_init();
        tma_prefetch_desc(&tmap_in);
    }
    if (warp == 1) {
        tmem_alloc_dyn(&bars->tmem_slot, (uint32_t)g.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ producer
        if (lane == 0) {
            uint32_t ka = 0, kb = 0;
            for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
                const TileCoord t = decode_tile(g, tile);
                for (int c = 0; c < g.cin_chunks; ++c) {
                    const uint32_t sa = ka % g.a_stages;
                    mbar_wait(&bars->empty_a[sa], ((ka / g.a_stages) & 1) ^ 1, 1);
                    if (g.ablate & 4) {
                        mbar_arrive(&bars->full_a[sa]);
                    } else {
                        mbar_arrive_expect_tx(&bars->full_a[sa], g.a_stage_bytes);
                        tma_load_4d(a_ring + (size_t)sa * g.a_stage_bytes, &tmap_in, &bars->full_a[sa], t.x0 * 8,
                                    t.y0, t.z0, t.n * g.in_groups_total + g.in_group_offset + 2 * c);
                    }
                    ++ka;
                    for (int grp = 0; grp < g.groups; ++grp) {
                        const uint32_t sb = kb % g.b_stages;
                        mbar_wait(&bars->empty_b[sb], ((kb / g.b_stages) & 1) ^ 1, 2);
                        if (g.ablate & 8) {
                            mbar_arrive(&bars->full_b[sb]);
                        } else {
                            mbar_arrive_expect_tx(&bars->full_b[sb], g.b_stage_bytes);
                            bulk_load_1d(b_ring + (size_t)sb * g.b_stage_bytes,
                                         wpack + (size_t)(c * g.groups + grp) * g.b_stage_bytes, g.b_stage_bytes,
                                         &bars->full_b[sb]);
                        }
                        ++kb;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------------------------------------------------- MMA issuer
        if (lane == 0) {
            uint32_t ka = 0, kb = 0, it = 0;
            const uint32_t R = g.b_rows;
            for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++it) {
                const uint32_t s = it % g.acc_stages;
                mbar_wait(&bars->tmem_empty[s], (it / g.acc_stages) & 1, 3);
                tc_fence_after();
                const uint32_t acc = tmem_base + s * acc_cols;
                for (int c = 0; c < g.cin_chunks; ++c) {
                    const uint32_t sa = ka % g.a_stages;
                    mbar_wait(&bars->full_a[sa], (ka / g.a_stages) & 1, 4);
                    const uint32_t a0 = smem_u32(a_ring + (size_t)sa * g.a_stage_bytes);
                    for (int grp = 0; grp < g.groups; ++grp) {
                        const uint32_t sb = kb % g.b_stages;
                        mbar_wait(&bars->full_b[sb], (kb / g.b_stages) & 1, 5);
                        tc_fence_after();
                        const uint32_t b0 = smem_u32(b_ring + (size_t)sb * g.b_stage_bytes);
                        if (g.ablate & 1) {
                        } else if (g.fold) {
                            for (int j = 0; j < g.bz + 2; ++j) {
                                const int lo = j - 2 > 0 ? j - 2 : 0;
                                const int hi = j < g.bz - 1 ? j : g.bz - 1;
                                const uint32_t n_mma = (uint32_t)(hi - lo + 1) * g.ncols;
                                const uint32_t row0 = (uint32_t)(lo - (j - 2)) * g.ncols;
                                const uint32_t idesc = idesc_bf16_m128(n_mma);
                                const uint32_t dcol = acc + lo * g.ncols;
#pragma unroll
                                for (int t = 0; t < 9; ++t) {
                                    const int dy = t / 3, dx = t - dy * 3;
                                    const uint64_t ad = smem_desc_kmajor_noswz(
                                        a0 + ((j * HALO_Y + dy) * HALO_X + dx) * 16, g.a_lbo, ROW_BYTES);
                                    const uint64_t bd =
                                        smem_desc_kmajor_noswz(b0 + t * 32 * R + row0 * 16, 16 * R, 128);
                                    umma_bf16(dcol, ad, bd, idesc, 1);
                                }
                            }
                        } else {
                            const uint32_t idesc = idesc_bf16_m128(g.ncols);
                            for (int b = 0; b < g.bz; ++b) {
                                const int j = b + grp;   // grp = kz = dz + 1
                                const uint32_t dcol = acc + b * g.ncols;
#pragma unroll
                                for (int t = 0; t < 9; ++t) {
                                    const int dy = t / 3, dx = t - dy * 3;
                                    const uint64_t ad = smem_desc_kmajor_noswz(
                                        a0 + ((j * HALO_Y + dy) * HALO_X + dx) * 16, g.a_lbo, ROW_BYTES);
                                    const uint64_t bd = smem_desc_kmajor_noswz(b0 + t * 32 * R, 16 * R, 128);
                                    umma_bf16(dcol, ad, bd, idesc, 1);
                                }
                            }
                        }
                        umma_commit(&bars->empty_b[sb]);
                        ++kb;
                    }
                    umma_commit(&bars->empty_a[sa]);
                    ++ka;
                }
                umma_commit(&bars->tmem_full[s]);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue
        const int q = warp & 3;                    // TMEM lane quadrant this warp may touch
        const int r = q * 32 + lane;               // accumulator row = voxel within the 8x16 patch
        const int ly = r >> 3, lx = r & 7;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int col = 0; col < g.acc_stages * acc_cols; col += 16) tmem_st16_zero(lane_base + col);
        tmem_wait_st();
        tc_fence_before();
        for (int s = 0; s < g.acc_stages; ++s) mbar_arrive(&bars->tmem_empty[s]);

        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++it) {
            const TileCoord t = decode_tile(g, tile);
            const uint32_t s = it % g.acc_stages;
            mbar_wait(&bars->tmem_full[s], (it / g.acc_stages) & 1, 6);
            tc_fence_after();
            const int y = t.y0 + ly, x = t.x0 + lx;
            const bool in_xy = (y < g.H) && (x < g.W);
            const uint32_t acc = lane_base + s * acc_cols;
            for (int b = 0; b < g.bz; ++b) {
                const int z = t.z0 + b;
                for (int cb = 0; cb < g.ncols / 16; ++cb) {
                    float v[16];
                    tmem_ld16(acc + b * g.ncols + cb * 16, v);
                    if (in_xy && z < g.D && !(g.ablate & 2)) epilogue_store16(ep, t.n, z, y, x, cb, v);
                }
            }
            for (int col = 0; col < acc_cols; col += 16) tmem_st16_zero(acc + col);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&bars->tmem_empty[s]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
}

}   // namespace anx
