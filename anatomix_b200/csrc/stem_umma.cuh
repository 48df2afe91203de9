// Stem convolution (Cin = 1..4, fp32 NCDHW input) on the tensor cores.
//
// The stem is 27*Cin MACs per output channel: far too thin in K for the generic
// kernel's 16-channel chunks, and 432 FFMA per voxel on the CUDA cores made it the
// most expensive launch of the 6M forward.  Here the 9 in-plane taps (dy,dx) of
// every input channel become the K dimension (9*Cin values, zero padded to a
// multiple of 16) and the three dz taps are folded into N exactly as in
// conv3_umma_kernel: ONE set of MMAs per INPUT plane accumulates into the three
// output planes it touches.
//
// fp32 accuracy on bf16 tensor cores: both operands are split hi + lo in bf16
// (x = x_hi + x_lo, w = w_hi + w_lo, 16 mantissa bits each) and three products are
// accumulated, x_hi*w_hi + x_lo*w_hi + x_hi*w_lo; the dropped x_lo*w_lo term is
// 2^-16 relative.  The A tiles live only in shared memory: 4 builder warps read
// the fp32 halo brick (TMA box, zero filled outside the volume; the reflect
// padding of network.py:310-318 is applied by re-indexing inside the brick) and
// write the canonical K-major non-swizzled core matrices.
//
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA
// issuer, warps 2..9 = epilogue (same code path as the generic conv: shift-seeded
// accumulators, reflect-shell stores, instance-norm statistics), warps 10..17 =
// A-tile builders: two groups of 128 threads (thread = voxel of the 8x16 patch) that
// take alternate input planes -- the build (shared-memory reads, hi/lo split, stores,
// proxy fence) is a latency chain per plane, and one group alone left the MMA issuer
// and the epilogue waiting (ncu: 23 % of all samples in the epilogue's accumulator wait).
#pragma once
#include "conv_umma.cuh"

namespace anx {

constexpr int STEM_BUILD_GROUPS = 2;      // builder groups of 128 threads (thread = voxel); group g builds planes ka = g (mod groups)
constexpr int STEM_THREADS = 64 + 32 * STEM_EPI_WARPS + 128 * STEM_BUILD_GROUPS;
constexpr int STEM_BRICK_X = 16;          // x0-4 .. x0+11: TMA wants the innermost start 16-byte aligned
constexpr int STEM_X_LEAD = 4;            // brick index of coordinate x0
constexpr int STEM_MAX_KQ = 3;            // K chunks of 16: 9*Cin <= 48  (Cin <= 4 uses 36)
constexpr int STEM_MAX_A_SLOTS = 8;
// A-tile ring (one slot = one input plane, hi + lo).  The builder -> MMA -> builder hand-over is a
// round trip of mbarrier latencies, so the ring depth bounds the planes in flight: 8 slots of 8 KB for
// the single-channel stem, 4 (of 16 / 24 KB) when the K chunks are larger.
__host__ __device__ constexpr int stem_a_slots(int kq) { return kq == 1 ? 8 : 4; }
// Halo bricks in flight (fp32 TMA boxes of 64-byte rows straight from the network input in HBM): the
// load of brick k + depth is issued when brick k is released, so depth - 1 tiles of build time must
// cover the box latency.
constexpr int STEM_MAX_BRICKS = 4;
__host__ __device__ constexpr int stem_bricks(int kq) { return kq == 1 ? 4 : 2; }

struct StemGeom {
    int N, D, H, W, cin;
    int tiles_x, tiles_y, tiles_z, tiles_per_sample, total_tiles;
    int bz, ncols, kq;                    // kq = ceil(9*cin / 16)
    int acc_stages, tmem_cols;
    int z_halo;                           // depth-slab mode: input has D+2 planes (see ANX_FLAG_DEPTH_HALO_INPUT)
    uint32_t brick_bytes;                 // cin * (bz+2) * 18 * 16 * 4
    uint32_t a_tile_bytes;                // kq * 2 * 128 * 16   (one of hi / lo)
    uint32_t b_bytes;                     // kq * 2 * (3*ncols) * 16 (one of hi / lo)
    uint32_t smem_bytes;
    int stats_acc;                        // EPI_STATS: per-warp sums in shared memory behind StemShared (STATS_ACC_BYTES more)
    int dbg_shift;                        // experiments only
    uint32_t ablate;                      // timing experiments (ANX_ABLATE): 1 no MMA, 2 no stores, 4 no A build, 8 no brick load
};

struct StemShared {
    uint64_t full_brick[STEM_MAX_BRICKS], empty_brick[STEM_MAX_BRICKS];
    uint64_t full_a[STEM_MAX_A_SLOTS], empty_a[STEM_MAX_A_SLOTS];
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_slot;
    uint32_t pad[3];
    float shift[64];
};

__device__ __forceinline__ int brick_reflect(int i, int lo, int hi) {   // lo/hi: first/last valid brick index
    return i < lo ? 2 * lo - i : (i > hi ? 2 * hi - i : i);
}


template <int KQ, int MODE>   // K chunks of 16 (1 for Cin = 1, 2 for Cin = 2..3, 3 for Cin = 4); EpiMode
__global__ void __launch_bounds__(STEM_THREADS, 1)
stem_umma_kernel(const __grid_constant__ CUtensorMap tmap_in, const StemGeom g, const uint8_t *__restrict__ wpack,
                 const Epilogue ep) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int STEM_A_SLOTS = stem_a_slots(KQ);
    constexpr uint32_t NB = stem_bricks(KQ);
    // [brick 0][brick 1][A slots: hi, lo][B hi][B lo][StemShared]
    const uint32_t brick_stride = (g.brick_bytes + 127) & ~127u;
    uint8_t *bricks = smem;
    uint8_t *a_ring = bricks + NB * brick_stride;
    uint8_t *b_hi = a_ring + (size_t)STEM_A_SLOTS * 2 * g.a_tile_bytes;
    uint8_t *b_lo = b_hi + g.b_bytes;
    StemShared *sh = reinterpret_cast<StemShared *>(b_lo + g.b_bytes);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int acc_cols = g.bz * g.ncols;
    const int planes = g.bz + 2;

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < NB; ++i) {
            mbar_init(&sh->full_brick[i], 1);
            mbar_init(&sh->empty_brick[i], 128 * STEM_BUILD_GROUPS);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sh->tmem_full[i], 1);
            mbar_init(&sh->tmem_empty[i], 32 * STEM_EPI_WARPS);
        }
        for (int i = 0; i < STEM_A_SLOTS; ++i) { mbar_init(&sh->full_a[i], 128); mbar_init(&sh->empty_a[i], 1); }
        fence_barrier_init();
        tma_prefetch_desc(&tmap_in);
    }
    if (warp == 1) {
        tmem_alloc_dyn(&sh->tmem_slot, (uint32_t)g.tmem_cols);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < g.ncols; i += blockDim.x) sh->shift[i] = ep.bias[i];
    for (uint32_t i = threadIdx.x; i < 2 * g.b_bytes / 16; i += blockDim.x)   // weights stay resident
        reinterpret_cast<uint4 *>(b_hi)[i] = reinterpret_cast<const uint4 *>(wpack)[i];
    fence_proxy_async();   // generic-proxy writes of B -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sh->tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ producer
        if (lane == 0) {
            uint32_t k = 0;
            for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++k) {
                const int n = tile / g.tiles_per_sample;
                int r = tile - n * g.tiles_per_sample;
                const int tz = r / (g.tiles_y * g.tiles_x);
                r -= tz * g.tiles_y * g.tiles_x;
                const int ty = r / g.tiles_x, tx = r - ty * g.tiles_x;
                const uint32_t s = k % NB;
                mbar_wait(&sh->empty_brick[s], ((k / NB) & 1) ^ 1, 11);
                if (ANX_ABL(g, 8)) { mbar_arrive(&sh->full_brick[s]); continue; }
                mbar_arrive_expect_tx(&sh->full_brick[s], g.brick_bytes);
                // box {16 x, 18 y, bz+2 z, cin}; planes of (n, c) are consecutive along dim 3
                tma_load_4d(bricks + s * brick_stride, &tmap_in, &sh->full_brick[s], tx * TILE_X - STEM_X_LEAD + g.dbg_shift,
                            ty * TILE_Y - 1 + g.dbg_shift, tz * g.bz - 1 + g.z_halo + g.dbg_shift, n * g.cin);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------------------------------------- MMA issuer (warp-uniform)
        uint32_t ka = 0, it = 0;
        const uint32_t R = 3 * g.ncols;                        // B rows: (dz=+1 | 0 | -1) x ncols
        const uint32_t a_hi_bits = (128u >> 4) | (1u << 14);   // SBO = 128 B (dense 8-row groups)
        const uint32_t a_lbo = ((128u * 16u) >> 4) << 16;      // next K half: 128 rows * 16 B
        const uint32_t b_lbo = (R & 0x3FFF) << 16;
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t bh = (smem_u32(b_hi) & 0x3FFFF) >> 4, bl = (smem_u32(b_lo) & 0x3FFFF) >> 4;
        for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++it) {
            const uint32_t s = it % g.acc_stages;
            mbar_wait_warp(&sh->tmem_empty[s], (it / g.acc_stages) & 1, 13);
            tc_fence_after();
            const uint32_t acc = tmem_u + s * acc_cols;
            for (int j = 0; j < planes; ++j, ++ka) {
                const uint32_t sa = ka % STEM_A_SLOTS;
                mbar_wait_warp(&sh->full_a[sa], (ka / STEM_A_SLOTS) & 1, 14);
                tc_fence_after();
                const int lo = j - 2 > 0 ? j - 2 : 0;
                const int hi = j < g.bz - 1 ? j : g.bz - 1;
                const uint32_t idesc = idesc_m128((uint32_t)(hi - lo + 1) * g.ncols, DT_BF16);
                const uint32_t dcol = acc + lo * g.ncols;
                const uint32_t row0 = (uint32_t)(lo - (j - 2)) * g.ncols;
                const uint32_t ah = (smem_u32(a_ring + (size_t)sa * 2 * g.a_tile_bytes) & 0x3FFFF) >> 4;
                const uint32_t al = ah + (g.a_tile_bytes >> 4);
#pragma unroll
                for (int kc = 0; kc < KQ; ++kc) {
                    if (ANX_ABL(g, 1)) break;
                    const uint32_t ao = kc * (2 * 128), bo = kc * (2 * R) + row0;   // 16 B units per K chunk of 16
                    umma_bf16_warp(dcol, make_desc(a_hi_bits, (ah + ao) | a_lbo), make_desc(a_hi_bits, (bh + bo) | b_lbo), idesc);
                    umma_bf16_warp(dcol, make_desc(a_hi_bits, (al + ao) | a_lbo), make_desc(a_hi_bits, (bh + bo) | b_lbo), idesc);
                    umma_bf16_warp(dcol, make_desc(a_hi_bits, (ah + ao) | a_lbo), make_desc(a_hi_bits, (bl + bo) | b_lbo), idesc);
                }
                if (ANX_ABL(g, 16)) {          // timing experiment (no MMAs in flight): plain arrive instead of tcgen05.commit
                    if (lane == 0) mbar_arrive(&sh->empty_a[sa]);
                    __syncwarp();
                } else {
                    umma_commit_warp(&sh->empty_a[sa]);
                }
            }
            umma_commit_warp(&sh->tmem_full[s]);
        }
        __syncwarp();
    } else if (warp >= 2 + STEM_EPI_WARPS) {
        // ------------------------------------------------------- A-tile builders
        const int bt = threadIdx.x - (64 + 32 * STEM_EPI_WARPS);
        const uint32_t group = (uint32_t)bt >> 7;            // alternate input planes between the builder groups
        const int r = bt & 127;                               // voxel of the 8 x 16 patch
        const int ly = r >> 3, lx = r & 7;
        uint32_t k = 0, ka = 0;
        for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++k) {
            const int n = tile / g.tiles_per_sample;
            int rr = tile - n * g.tiles_per_sample;
            const int tz = rr / (g.tiles_y * g.tiles_x);
            rr -= tz * g.tiles_y * g.tiles_x;
            const int ty = rr / g.tiles_x, tx = rr - ty * g.tiles_x;
            const int x0 = tx * TILE_X, y0 = ty * TILE_Y, z0 = tz * g.bz;
            // valid brick index range per axis (brick index b <-> coordinate origin - 1 + b)
            // x: brick index b <-> coordinate x0 - STEM_X_LEAD + b
            const int xlo = x0 == 0 ? STEM_X_LEAD : 0, xhi = min(STEM_BRICK_X - 1, g.W - 1 - x0 + STEM_X_LEAD);
            const int ylo = y0 == 0 ? 1 : 0, yhi = min(HALO_Y - 1, g.H - y0);
            const int zlo = (z0 == 0 && !g.z_halo) ? 1 : 0;
            const int zhi = g.z_halo ? planes - 1 : min(planes - 1, g.D - z0);
            int bx[3], by[3];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                bx[t] = brick_reflect(min(lx + t + STEM_X_LEAD - 1, xhi + 1), xlo, xhi);
                by[t] = brick_reflect(min(ly + t, yhi + 1), ylo, yhi);
            }
            const uint32_t s = k % NB;
            mbar_wait(&sh->full_brick[s], (k / NB) & 1, 15);
            const float *brick = reinterpret_cast<const float *>(bricks + s * brick_stride);
            for (int j = 0; j < planes; ++j, ++ka) {
                if (ka % STEM_BUILD_GROUPS != group) continue;
                const uint32_t sa = ka % STEM_A_SLOTS;
                if (ANX_ABL(g, 4)) {
                    mbar_wait(&sh->empty_a[sa], ((ka / STEM_A_SLOTS) & 1) ^ 1, 16);
                    mbar_arrive(&sh->full_a[sa]);
                    continue;
                }
                const int bzj = brick_reflect(min(j, zhi + 1), zlo, zhi);
                // K index = c*9 + (dy*3+dx); 16*KQ slots, unused ones are zero
                float v[16 * KQ];
#pragma unroll
                for (int i = 0; i < 16 * KQ; ++i) v[i] = 0.0f;
#pragma unroll
                for (int c = 0; c < (16 * KQ) / 9; ++c) {
                    if (c < g.cin) {
                        const float *pl = brick + ((size_t)(c * planes + bzj) * HALO_Y) * STEM_BRICK_X;
#pragma unroll
                        for (int t = 0; t < 9; ++t) v[c * 9 + t] = pl[by[t / 3] * STEM_BRICK_X + bx[t % 3]];
                    }
                }
                uint32_t hi[8 * KQ], lo[8 * KQ];
#pragma unroll
                for (int i = 0; i < 8 * KQ; ++i) {
                    if (2 * i < 9 * ((16 * KQ) / 9)) split_hi_lo(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
                    else { hi[i] = 0u; lo[i] = 0u; }
                }
                mbar_wait(&sh->empty_a[sa], ((ka / STEM_A_SLOTS) & 1) ^ 1, 16);
                uint8_t *ah = a_ring + (size_t)sa * 2 * g.a_tile_bytes, *al = ah + g.a_tile_bytes;
#pragma unroll
                for (int h8 = 0; h8 < 2 * KQ; ++h8) {
                    // canonical K-major: [K half-chunk][row][8 elements], 16 B per row
                    *reinterpret_cast<uint4 *>(ah + ((size_t)h8 * 128 + r) * 16) =
                        make_uint4(hi[4 * h8], hi[4 * h8 + 1], hi[4 * h8 + 2], hi[4 * h8 + 3]);
                    *reinterpret_cast<uint4 *>(al + ((size_t)h8 * 128 + r) * 16) =
                        make_uint4(lo[4 * h8], lo[4 * h8 + 1], lo[4 * h8 + 2], lo[4 * h8 + 3]);
                }
                fence_proxy_async();          // make the generic-proxy stores visible to the MMA
                mbar_arrive(&sh->full_a[sa]);
            }
            mbar_arrive(&sh->empty_brick[s]);
        }
    } else {
        // ------------------------------------------------------------ epilogue
        const int q = warp & 3;
        const PlaneSplit half{(warp - 2) >> 2, STEM_EPI_WARPS / 4, 0};
        const int r = q * 32 + lane;
        const int ly = r >> 3, lx = r & 7;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const int chunks = g.ncols >> 4;
        for (int s = 0; s < g.acc_stages; ++s)
            for (int b = plane_lo(half, g.bz); b < plane_hi(half, g.bz); ++b)
                for (int cb = 0; cb < chunks; ++cb) tmem_st16(lane_base + s * acc_cols + b * g.ncols + cb * 16, sh->shift + cb * 16);
        tmem_wait_st();
        tc_fence_before();
        for (int s = 0; s < g.acc_stages; ++s) mbar_arrive(&sh->tmem_empty[s]);

        double *my_acc = nullptr;
        int acc_n = -1;
        if constexpr (MODE == EPI_STATS) {
            if (g.stats_acc) {               // region behind StemShared, allocated by EPI_STATS launches only
                my_acc = reinterpret_cast<double *>(sh + 1) + (warp - 2) * 2 * STATS_ACC_MAX_COLS;
                for (int i = lane; i < 2 * g.ncols; i += 32) my_acc[i] = 0.0;
                __syncwarp();
            }
        }
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++it) {
            const int n = tile / g.tiles_per_sample;
            if constexpr (MODE == EPI_STATS) {
                if (my_acc && n != acc_n) {
                    if (acc_n >= 0) warp_stats_flush(my_acc, chunks, ep.stats, ep.stats_stride, acc_n);
                    acc_n = n;
                }
            }
            int rr = tile - n * g.tiles_per_sample;
            const int tz = rr / (g.tiles_y * g.tiles_x);
            rr -= tz * g.tiles_y * g.tiles_x;
            const int ty = rr / g.tiles_x, tx = rr - ty * g.tiles_x;
            const uint32_t s = it % g.acc_stages;
            EpiTile et;
            et.n = n; et.z0 = tz * g.bz; et.y = ty * TILE_Y + ly; et.x = tx * TILE_X + lx;
            et.chan0 = 0;
            et.in_xy = (et.y < g.H) && (et.x < g.W);
            et.store = !(ANX_ABL(g, 2));
            mbar_wait(&sh->tmem_full[s], (it / g.acc_stages) & 1, 17);
            tc_fence_after();
            umma_epilogue_tile<MODE>(ep, et, lane_base + s * acc_cols, sh->shift, half, g.bz, g.ncols, g.D, et, false, my_acc);
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
