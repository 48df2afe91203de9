// 16 -> 16 channel 3x3x3 convolution at full-row granularity: the thin layers of the network (reference
// network.py:334-369, 413-461 at level 0), whose generic tile (M = 8 x 16 voxels, N = 3 * 16 with dz folded) is
// bound by the tensor core's shared-memory operand fetch -- 4 KB of A feed only 48 columns.
//
// Here the 128 TMEM lanes are 128 consecutive x voxels of ONE row, and the accumulator columns enumerate
// output ROWS: column block (yo * 3 + (z mod 3)) * 16 of a strip of 4 output rows.  One input row (y', z'),
// shifted by dx, then contributes to up to 3 x 3 output rows that are CONTIGUOUS in TMEM, so a single MMA
// with N = 144 (B = the 3 x 3 (dy, dz) weight blocks stacked along N) replaces nine N = 16 products: dy AND dz
// are folded into N, only dx remains as separate instructions.  Shared-memory wavefronts per output row of
// 128 voxels drop from 9 * 43.5 to about 250.
//
// A CTA owns a block of 8 output rows (two strips of 4, alternating so that the epilogue of one overlaps the
// MMAs of the other) of one 128-wide x tile and streams ZS output planes through a ring of input planes:
//   producer warp : per input plane, 10 rows x 2 channel groups of 130 voxels -> 20 bulk copies (rows are
//                   contiguous in the padded planar buffer; no tensor map needed)
//   MMA warp      : per plane and strip 6 input rows x 3 dx = 18 MMAs (N = 48 / 96 / 144 at the strip edges),
//                   weights resident in shared memory as three images (rotation of the z slots by z mod 3)
//   8 epilogue warps: after plane p, output plane p - 2 of the strip is complete: tcgen05.ld, re-seed with
//                   the channel shift, activation, 16-bit pack, 512-byte contiguous row stores (+ shell mirrors),
//                   or fp32 NCDHW / fused-head stores for the last conv.
// Output planes outside the z segment ("phantoms": contributions of the first / last two input planes) are
// simply re-seeded, so every plane runs the same instruction sequence.
#pragma once
#include <cstddef>
#include "conv_umma.cuh"

namespace anx {

constexpr int ROWS_X = 128;                     // x voxels per tile = TMEM lanes
constexpr int ROWS_BY = 4;                      // output rows per strip
constexpr int ROWS_YB = 2 * ROWS_BY;            // output rows per CTA unit
constexpr int ROWS_IN_Y = ROWS_YB + 2;          // input rows per plane
constexpr int ROWS_ROW_BYTES = (ROWS_X + 2) * 16;               // one input row of one channel group
constexpr int ROWS_GROUP_BYTES = ROWS_IN_Y * ROWS_ROW_BYTES;    // LBO: distance between the two K halves
constexpr int ROWS_PLANE_BYTES = 2 * ROWS_GROUP_BYTES;          // one ring stage
constexpr int ROWS_STAGES = 4;
constexpr int ROWS_N = 9 * 16;                  // rows of one B tap matrix
constexpr int ROWS_B_TAP_BYTES = 2 * ROWS_N * 16;
constexpr int ROWS_B_IMAGE_BYTES = 3 * ROWS_B_TAP_BYTES;        // three dx taps
constexpr int ROWS_STRIP_COLS = ROWS_BY * 48;   // TMEM columns of one strip
constexpr int ROWS_THREADS = 64 + 32 * 8;
// Epilogue warps: (TMEM lane quadrant) x (row group of RPW output rows of each strip).  RPW = 2: eight warps, each
// draining a row pair (needed by the fused pooling, whose y pair is the warp's two rows); RPW = 1: sixteen warps with
// one row each -- the drain is a latency chain per warp (tcgen05.ld, re-seed, pack, stores, barrier round trip), so
// twice the warps hide twice the latency (ablation at batch 4: 66 us of a 121 us launch is that skeleton).
__host__ __device__ constexpr int rows_threads(int rpw, bool stem) { return 64 + 32 * (16 / rpw) + (stem ? 128 : 0); }
// Stem variant (C_in = 1, fp32 NCDHW input): four builder warps (thread = x voxel) write the A tiles, see below.
constexpr int ROWS_STEM_THREADS = ROWS_THREADS + 128;
constexpr int ROWS_STEM_TILE_BYTES = 2 * ROWS_X * 16;           // one input row: two K halves x 128 lanes x 16 B
constexpr int ROWS_STEM_B_IMAGE_BYTES = 2 * ROWS_N * 16;        // one z rotation: all three dx taps live in K
static_assert(ROWS_IN_Y * ROWS_STEM_TILE_BYTES <= ROWS_PLANE_BYTES, "a stem plane must fit a ring stage");
// fp32 input staging of the stem: per plane ROWS_IN_Y row segments [x0 - 4, x0 + 132) (16-byte aligned on both
// sides), bulk-copied by the producer warp several planes ahead of the builders.  The ring lives in the part of
// the weight area the stem's small B images leave free.
constexpr int ROWS_STEM_IN_ROW_BYTES = (ROWS_X + 8) * 4;
constexpr int ROWS_STEM_IN_PLANE_BYTES = ROWS_IN_Y * ROWS_STEM_IN_ROW_BYTES;
constexpr int ROWS_STEM_IN_STAGES = 5;
static_assert(3 * ROWS_STEM_B_IMAGE_BYTES + ROWS_STEM_IN_STAGES * ROWS_STEM_IN_PLANE_BYTES <= 3 * ROWS_B_IMAGE_BYTES,
              "stem input staging must fit behind the stem's B images");

struct RowsGeom {
    int N, D, H, W;
    int zs;                  // output planes per unit
    int tiles_x, tiles_y, tiles_z, units_per_sample, total_units;
    int dt;
    uint32_t smem_bytes;
    uint32_t ablate;
    int z_halo;              // stem variant: the input carries one extra plane at each end of D (depth-slab mode)
};

struct RowsShared {
    uint64_t full_a[ROWS_STAGES], empty_a[ROWS_STAGES];
    uint64_t full_b;
    uint64_t acc_ready[2][2], drained[2][2];   // [strip][row pair]: hand-over at half-strip granularity
    uint64_t full_in[ROWS_STEM_IN_STAGES], empty_in[ROWS_STEM_IN_STAGES];   // stem: fp32 input staging ring
    uint32_t tmem_slot;
    uint32_t pad[5];         // keeps `shift` 16-byte aligned (float4 reads of the head weights)
    float shift[16 + HEAD_FLOATS + 16];
};
static_assert(offsetof(RowsShared, shift) % 16 == 0, "shift must be 16-byte aligned");

// STEM = true: the network's first conv (C_in = 1, reference network.py:309-326) in the same row form.  K = 16 of
// ONE instruction holds the three dx taps of the hi / lo bf16 split of the fp32 input,
//      A[x][k] = { hi(x-1), hi(x), hi(x+1),  lo(x-1), lo(x), lo(x+1),  hi(x-1), hi(x), hi(x+1),  0 ... }
//      B[n][k] = { w_hi(dx = 0, 1, 2),       w_hi(0, 1, 2),            w_lo(0, 1, 2),            0 ... }
// (x_hi w_hi + x_lo w_hi + x_hi w_lo: fp32-level accuracy, the dropped x_lo w_lo term is 2^-16 relative), so one
// N = 144 MMA per input row replaces the three dx instructions of the 16-channel layers.  Four builder warps
// (thread = x voxel) read the fp32 rows straight from global memory (reflect padding by index arithmetic) and write
// the canonical K-major A tiles into the plane ring; MMA issue, TMEM layout and epilogue are the shared code.
template <int MODE, bool STEM = false, int RPW = 2>   // EPI_PADDED, EPI_POOL (max), EPI_SEEDED, EPI_F32 (also the fused gather), EPI_CL16 or EPI_F32_HEAD
__global__ void __launch_bounds__(rows_threads(RPW, STEM), 1)
conv3_rows_kernel(const ActView src, const RowsGeom g, const uint8_t *__restrict__ wrows, const Epilogue ep,
                  const float *__restrict__ stem_in) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *a_ring = smem;
    uint8_t *b_img = a_ring + (size_t)ROWS_STAGES * ROWS_PLANE_BYTES;
    RowsShared *sh = reinterpret_cast<RowsShared *>(b_img + 3 * ROWS_B_IMAGE_BYTES);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int planes = g.zs + 2;

    if (threadIdx.x == 0) {
        for (int i = 0; i < ROWS_STAGES; ++i) { mbar_init(&sh->full_a[i], STEM ? 128 : 1); mbar_init(&sh->empty_a[i], 1); }
        mbar_init(&sh->full_b, 1);
        for (int i = 0; i < ROWS_STEM_IN_STAGES; ++i) { mbar_init(&sh->full_in[i], 1); mbar_init(&sh->empty_in[i], 128); }
        for (int i = 0; i < 4; ++i) { mbar_init(&sh->acc_ready[i >> 1][i & 1], 1); mbar_init(&sh->drained[i >> 1][i & 1], 4 * (2 / RPW)); }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc_dyn(&sh->tmem_slot, 512u);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < 16; i += blockDim.x) sh->shift[i] = ep.bias[i];
    if constexpr (MODE == EPI_F32_HEAD)
        for (int i = threadIdx.x; i < HEAD_FLOATS; i += blockDim.x) sh->shift[HEAD_SMEM_OFFSET + i] = ep.head[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sh->tmem_slot;

    auto decode = [&](int unit, int &n, int &x0, int &y0, int &z0) {
        n = unit / g.units_per_sample;
        int r = unit - n * g.units_per_sample;
        const int tz = r / (g.tiles_y * g.tiles_x);
        r -= tz * g.tiles_y * g.tiles_x;
        const int ty = r / g.tiles_x;
        // the last x tile of a width that is not a multiple of 128 is pulled back to end at the border: it recomputes
        // (and re-stores, identically) the voxels it shares with its left neighbour -- no masked lanes anywhere
        x0 = (r - ty * g.tiles_x) * ROWS_X;
        if (x0 + ROWS_X > g.W) x0 = g.W - ROWS_X;
        y0 = ty * ROWS_YB;
        z0 = tz * g.zs;
    };

    constexpr uint32_t B_BYTES = STEM ? 3 * ROWS_STEM_B_IMAGE_BYTES : 3 * ROWS_B_IMAGE_BYTES;
    if (warp == 0) {
        // ------------------------------------------------------------ producer
        if (lane == 0 && blockIdx.x < g.total_units) {
            mbar_arrive_expect_tx(&sh->full_b, B_BYTES);
            bulk_load_1d(b_img, wrows, B_BYTES, &sh->full_b);
        }
        uint32_t ka = 0;
        if constexpr (STEM) {
            // fp32 rows of input plane z0 - 1 + p (reflect at the global faces, or the attached neighbour planes):
            // row i of the stage = input row reflect(y0 - 1 + i), floats [x0 - 4, x0 + 132) clipped to the volume
            uint8_t *stage0 = b_img + 3 * ROWS_STEM_B_IMAGE_BYTES;
            auto refl = [](int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); };
            for (int unit = blockIdx.x; unit < g.total_units; unit += gridDim.x) {
                int n, x0, y0, z0;
                decode(unit, n, x0, y0, z0);
                const int gx_lo = x0 >= 4 ? x0 - 4 : 0, gx_hi = x0 + ROWS_X + 4 <= g.W ? x0 + ROWS_X + 4 : g.W;
                const uint32_t bytes = (uint32_t)(gx_hi - gx_lo) * 4u, dst_off = (uint32_t)(gx_lo - (x0 - 4)) * 4u;
                for (int p = 0; p < planes; ++p, ++ka) {
                    const uint32_t k = ka % ROWS_STEM_IN_STAGES;
                    if (lane == 0) {
                        mbar_wait(&sh->empty_in[k], ((ka / ROWS_STEM_IN_STAGES) & 1) ^ 1, 27);
                        mbar_arrive_expect_tx(&sh->full_in[k], ROWS_IN_Y * bytes);
                    }
                    __syncwarp();
                    if (lane < ROWS_IN_Y) {
                        const int zi = g.z_halo ? z0 + p : refl(z0 - 1 + p, g.D);
                        const float *row = stem_in + (((size_t)n * (g.D + 2 * g.z_halo) + zi) * g.H + refl(y0 - 1 + lane, g.H)) * (size_t)g.W + gx_lo;
                        bulk_load_1d(stage0 + (size_t)k * ROWS_STEM_IN_PLANE_BYTES + (size_t)lane * ROWS_STEM_IN_ROW_BYTES + dst_off,
                                     row, bytes, &sh->full_in[k]);
                    }
                }
            }
        }
        if constexpr (!STEM)
        for (int unit = blockIdx.x; unit < g.total_units; unit += gridDim.x) {
            int n, x0, y0, z0;
            decode(unit, n, x0, y0, z0);
            for (int p = 0; p < planes; ++p, ++ka) {
                const uint32_t st = ka % ROWS_STAGES;
                if (lane == 0) {
                    mbar_wait(&sh->empty_a[st], ((ka / ROWS_STAGES) & 1) ^ 1, 21);
                    if (ANX_ABL(g, 4)) mbar_arrive(&sh->full_a[st]);
                    else mbar_arrive_expect_tx(&sh->full_a[st], ROWS_PLANE_BYTES);
                }
                __syncwarp();
                if (!(ANX_ABL(g, 4)) && lane < 2 * ROWS_IN_Y) {
                    const int grp = lane / ROWS_IN_Y, row = lane - grp * ROWS_IN_Y;
                    // padded coordinates: plane z0 + p, row y0 + row, 130 voxels from x0 (= interior x0 - 1)
                    bulk_load_1d(a_ring + (size_t)st * ROWS_PLANE_BYTES + (size_t)grp * ROWS_GROUP_BYTES +
                                     (size_t)row * ROWS_ROW_BYTES,
                                 src.at(n, grp, z0 + p, y0 + row, x0), ROWS_ROW_BYTES, &sh->full_a[st]);
                }
            }
        }
        __syncwarp();
    } else if (STEM && warp >= 2 + 16 / RPW) {
        // ------------------------------------------------- stem: A-tile builders
        const int bx = threadIdx.x - rows_threads(RPW, false);   // x voxel of the 128-wide tile
        const int Ww = g.W;
        const uint8_t *stage0 = b_img + 3 * ROWS_STEM_B_IMAGE_BYTES;
        uint32_t ka = 0;
        for (int unit = blockIdx.x; unit < g.total_units; unit += gridDim.x) {
            int n, x0, y0, z0;
            decode(unit, n, x0, y0, z0);
            const int x = x0 + bx;
            // staged floats start at x0 - 4: x sits at bx + 4; reflect padding at the volume's x faces
            const int im = x == 0 ? bx + 5 : bx + 3, ip = x == Ww - 1 ? bx + 3 : bx + 5;
            for (int p = 0; p < planes; ++p, ++ka) {
                const uint32_t st = ka % ROWS_STAGES;
                const uint32_t k = ka % ROWS_STEM_IN_STAGES;
                mbar_wait(&sh->full_in[k], (ka / ROWS_STEM_IN_STAGES) & 1, 28);
                const float *stg = reinterpret_cast<const float *>(stage0 + (size_t)k * ROWS_STEM_IN_PLANE_BYTES);
                float v[ROWS_IN_Y][3];
#pragma unroll
                for (int i = 0; i < ROWS_IN_Y; ++i) {
                    const float *row = stg + i * (ROWS_STEM_IN_ROW_BYTES / 4);
                    v[i][0] = row[im];
                    v[i][1] = row[bx + 4];
                    v[i][2] = row[ip];
                }
                mbar_arrive(&sh->empty_in[k]);                // values are in registers: the producer may refill the stage
                mbar_wait(&sh->empty_a[st], ((ka / ROWS_STAGES) & 1) ^ 1, 26);
                uint8_t *tile = a_ring + (size_t)st * ROWS_PLANE_BYTES + (size_t)bx * 16;
#pragma unroll
                for (int i = 0; i < ROWS_IN_Y; ++i) {
                    // bf16 hi / lo split of the three taps: hi = rn(x), lo = rn(x - hi)
                    uint32_t h01, l01, h2z, l2z;
                    split_hi_lo(v[i][0], v[i][1], h01, l01);
                    split_hi_lo(v[i][2], 0.0f, h2z, l2z);
                    const uint32_t h2 = h2z & 0xffffu, l0 = l01 & 0xffffu, l1 = l01 >> 16, l2 = l2z & 0xffffu;
                    // K = { hi0 hi1 | hi2 lo0 | lo1 lo2 | hi0 hi1 }  { hi2 0 | 0 0 | 0 0 | 0 0 }
                    *reinterpret_cast<uint4 *>(tile + (size_t)i * ROWS_STEM_TILE_BYTES) =
                        make_uint4(h01, h2 | (l0 << 16), l1 | (l2 << 16), h01);
                    *reinterpret_cast<uint4 *>(tile + (size_t)i * ROWS_STEM_TILE_BYTES + ROWS_X * 16) = make_uint4(h2, 0u, 0u, 0u);
                }
                fence_proxy_async();          // generic-proxy stores -> visible to the tensor core
                mbar_arrive(&sh->full_a[st]);
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------- MMA issuer (warp-uniform)
        uint32_t ka = 0;
        const uint32_t hi_bits = (128u >> 4) | (1u << 14);                   // SBO = 128 B for A and B
        const uint32_t a_lbo = ((uint32_t)((STEM ? ROWS_X * 16 : ROWS_GROUP_BYTES) >> 4) & 0x3FFF) << 16;
        const uint32_t b_lbo = ((uint32_t)ROWS_N & 0x3FFF) << 16;            // 16 * 144 bytes between the K halves
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t b0 = (smem_u32(b_img) & 0x3FFFF) >> 4;
        constexpr uint32_t A_ROW16 = (STEM ? ROWS_STEM_TILE_BYTES : ROWS_ROW_BYTES) >> 4;          // 16-byte units per input row
        constexpr uint32_t B_IMG16 = (STEM ? ROWS_STEM_B_IMAGE_BYTES : ROWS_B_IMAGE_BYTES) >> 4;
        const int mma_dt = STEM ? (int)DT_BF16 : g.dt;                       // the stem's operands are always bf16 hi / lo pairs
        mbar_wait_warp(&sh->full_b, 0, 22);
        for (int unit = blockIdx.x; unit < g.total_units; unit += gridDim.x) {
            for (int p = 0; p < planes; ++p, ++ka) {
                const uint32_t st = ka % ROWS_STAGES;
                mbar_wait_warp(&sh->full_a[st], (ka / ROWS_STAGES) & 1, 23);
                if constexpr (STEM) tc_fence_after();
                const uint32_t a0 = (smem_u32(a_ring + (size_t)st * ROWS_PLANE_BYTES) & 0x3FFFF) >> 4;
                const uint32_t bimg = b0 + (uint32_t)(p % 3) * B_IMG16;
                for (int s = 0; s < 2; ++s) {
                    // Hand-over with the epilogue per ROW PAIR of the strip: output rows {0, 1} are complete after
                    // input row 3 and are drained (and re-seeded) while input rows 4, 5 still accumulate into rows
                    // {2, 3}; the next plane's rows 0, 1 only need pair 0 drained.  Twice the pipeline depth of a
                    // whole-strip hand-over for the same TMEM.
#pragma unroll
                    for (int i = 0; i < ROWS_BY + 2; ++i) {
                        if (i == 0 || i == 2) {                     // input rows 0, 1 touch pair 0 only; row 2 is the first to touch pair 1
                            mbar_wait_warp(&sh->drained[s][i >> 1], ka & 1, 24);
                            tc_fence_after();
                        }
                        if (!(ANX_ABL(g, 1))) {
                            const int yo_min = i - 2 > 0 ? i - 2 : 0;
                            const int yo_max = i < ROWS_BY - 1 ? i : ROWS_BY - 1;
                            const uint32_t nrows = (uint32_t)(yo_max - yo_min + 1);
                            const uint32_t idesc = idesc_m128(nrows * 48u, mma_dt);
                            const uint32_t dcol = tmem_u + (uint32_t)s * ROWS_STRIP_COLS + (uint32_t)yo_min * 48u;
                            const uint32_t arow = a0 + (uint32_t)(s * ROWS_BY + i) * A_ROW16;
                            const uint32_t brow = bimg + (uint32_t)(2 - (i - yo_min)) * 48u;   // 16 B per B row
                            if constexpr (STEM) {
                                umma_bf16_warp(dcol, make_desc(hi_bits, arow | a_lbo), make_desc(hi_bits, brow | b_lbo), idesc);
                            } else {
#pragma unroll
                                for (int dx = 0; dx < 3; ++dx)
                                    umma_bf16_warp(dcol, make_desc(hi_bits, (arow + dx) | a_lbo),
                                                   make_desc(hi_bits, (brow + dx * (ROWS_B_TAP_BYTES >> 4)) | b_lbo), idesc);
                            }
                        }
                        if (i == 3) umma_commit_warp(&sh->acc_ready[s][0]);      // rows 0, 1 of the strip have all their taps
                    }
                    umma_commit_warp(&sh->acc_ready[s][1]);
                }
                umma_commit_warp(&sh->empty_a[st]);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue
        constexpr bool SEEDED = MODE == EPI_SEEDED;
        constexpr bool POOL = MODE == EPI_POOL;
        constexpr bool PADDED = MODE == EPI_PADDED || SEEDED || POOL;
        const int q = warp & 3;                    // TMEM lane quadrant
        static_assert(RPW == 2 || MODE != EPI_POOL, "the fused pooling needs a row pair per warp");
        const int h = (warp - 2) >> 2;             // row group: rows {RPW * h + r} of each strip
        const int pair = RPW == 2 ? h : h >> 1;    // the strip's row pair these rows belong to (hand-over barriers)
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const int lx = q * 32 + lane;
        const int Dd = g.D, Hh = g.H, Ww = g.W;
        const size_t vol = (size_t)Dd * Hh * Ww;
        // loop invariants of the padded stores: row / plane / group strides (uint4 units), shell semantics
        const size_t rowp = (size_t)ep.dst.pitch, plane = rowp * (Hh + 2), gstride = plane * (Dd + 2);
        const int rep = ep.dst.shell_rep, zopen = ep.dst.z_open;
        float sd[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) sd[i] = SEEDED ? 0.0f : sh->shift[i];
        // SEEDED: the accumulators of output plane o start from stored partial sums (the low-resolution half of
        // a decoder conv, see engine.cu "upconv") instead of the channel shift
        const size_t sgstride = SEEDED ? (size_t)ep.seed_src.pitch * (Hh + 2) * (Dd + 2) : 0;
        auto load_seed = [&](bool valid, int n_, int z_, int y_, int x_, uint4 &s0, uint4 &s1) {
            s0 = s1 = make_uint4(0u, 0u, 0u, 0u);
            if (SEEDED && valid) {
                const uint4 *ps = ep.seed_src.at(n_, 0, z_ + 1, y_ + 1, x_ + 1);
                s0 = __ldg(ps);
                s1 = __ldg(ps + sgstride);
            }
        };
        auto seed_store = [&](uint32_t col, const uint4 &s0, const uint4 &s1) {
            if constexpr (SEEDED) {
                float sv[16];
                unpack_x8(s0, sv, ep.dt);
                unpack_x8(s1, sv + 8, ep.dt);
                tmem_st16(col, sv);
            } else {
                tmem_st16(col, sd);
            }
        };
        // start: slot 0 holds the seeds of the first unit's plane 0 (slots 1 and 2 are re-seeded by the two
        // phantom planes before a real plane touches them); without seeds every column starts from the shift
        {
            int n, x0, y0, z0;
            decode(blockIdx.x < g.total_units ? blockIdx.x : 0, n, x0, y0, z0);
            for (int s = 0; s < 2; ++s)
                for (int r = 0; r < RPW; ++r) {
                    uint4 s0, s1;
                    load_seed(blockIdx.x < g.total_units, n, z0, y0 + s * ROWS_BY + RPW * h + r, x0 + lx, s0, s1);
                    for (int slot = 0; slot < 3; ++slot) {
                        const uint32_t col = lane_base + s * ROWS_STRIP_COLS + ((RPW * h + r) * 3 + slot) * 16;
                        if (slot == 0) seed_store(col, s0, s1);
                        else tmem_st16(col, sd);
                    }
                }
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&sh->drained[0][pair]); mbar_arrive(&sh->drained[1][pair]); }

        uint32_t held[2][8];                        // POOL: row-pair maxima of the even plane of a z pair, per strip
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int i = 0; i < 8; ++i) held[s][i] = 0u;

        // SEEDED: seeds the drain of (plane p, strip s) writes back -- those of the plane that uses the slot
        // next: o + 3 of the same unit, or plane 0 of the next unit after the last plane.  They are loaded one
        // drain ahead (`nsq`), so the L2 / HBM latency hides behind a drain's worth of work.
        auto seeds_for = [&](int p_, int s_, int n_, int x0_, int y0_, int z0_, bool nvalid, int nn_, int nx0_, int ny0_,
                             int nz0_, uint4 (&out)[RPW][2]) {
            const int o_ = p_ - 2;
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int yo = s_ * ROWS_BY + RPW * h + r;
                if (o_ + 3 < g.zs) load_seed(true, n_, z0_ + o_ + 3, y0_ + yo, x0_ + lx, out[r][0], out[r][1]);
                else if (o_ == g.zs - 1) load_seed(nvalid, nn_, nz0_, ny0_ + yo, nx0_ + lx, out[r][0], out[r][1]);
                else load_seed(false, 0, 0, 0, 0, out[r][0], out[r][1]);
            }
        };
        uint4 nsq[RPW][2];
        if constexpr (SEEDED) {
            int n, x0, y0, z0;
            decode(blockIdx.x < g.total_units ? blockIdx.x : 0, n, x0, y0, z0);
            seeds_for(0, 0, n, x0, y0, z0, false, 0, 0, 0, 0, nsq);     // p = 0 never needs the next unit
        }

        uint32_t ka = 0;
        for (int unit = blockIdx.x; unit < g.total_units; unit += gridDim.x) {
            int n, x0, y0, z0;
            decode(unit, n, x0, y0, z0);
            const int x = x0 + lx;
            const int mdx = mirror_delta(x, Ww, rep);              // this lane's x never changes within a unit
            // fp32 outputs: this lane's column of the unit's sample (one pointer per unit, rows add an offset)
            const size_t out_off = (size_t)(ep.sample_offset + n) * ep.out_nstride + (size_t)x;
            const int next_unit = unit + (int)gridDim.x;
            const bool nvalid = next_unit < g.total_units;
            int nn = 0, nx0 = 0, ny0 = 0, nz0 = 0;
            if (SEEDED && nvalid) decode(next_unit, nn, nx0, ny0, nz0);
            for (int p = 0; p < planes; ++p, ++ka) {
                const int o = p - 2;                               // output plane completed by input plane p
                const int slot = (o + 3) % 3;
                const int z = z0 + o;
                const int mdz = mirror_delta_z(z, Dd, rep, zopen);  // one plane per drain: uniform over the warp
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    // seeds of the plane that uses this slot next: o + 3 of this unit, or plane 0 of the next unit
                    uint4 sq[RPW][2];
                    if constexpr (SEEDED) {
#pragma unroll
                        for (int r = 0; r < RPW; ++r) { sq[r][0] = nsq[r][0]; sq[r][1] = nsq[r][1]; }
                        // the drain after this one: the other strip, the next plane, or the first drain of the next unit
                        if (s == 0) seeds_for(p, 1, n, x0, y0, z0, nvalid, nn, nx0, ny0, nz0, nsq);
                        else if (p + 1 < planes) seeds_for(p + 1, 0, n, x0, y0, z0, nvalid, nn, nx0, ny0, nz0, nsq);
                        else if (nvalid) seeds_for(0, 0, nn, nx0, ny0, nz0, false, 0, 0, 0, 0, nsq);
                    }
                    mbar_wait(&sh->acc_ready[s][pair], ka & 1, 25);
                    tc_fence_after();
                    uint32_t pm[8];                                // POOL: maximum over this warp's two rows
#pragma unroll
                    for (int r = 0; r < RPW; ++r) {
                        const int yo = RPW * h + r;
                        const uint32_t col = lane_base + s * ROWS_STRIP_COLS + (yo * 3 + slot) * 16;
                        __syncwarp();
                        if (o < 0) {                               // phantom plane below the segment: discard
                            seed_store(col, sq[r][0], sq[r][1]);
                            continue;
                        }
                        uint32_t rr[16];
                        tmem_ld16_nowait(col, rr);
                        tmem_wait_ld();
                        tmem_ld_ready16(rr);
                        seed_store(col, sq[r][0], sq[r][1]);
                        const int y = y0 + s * ROWS_BY + yo;
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = activate(__uint_as_float(rr[i]), ep.act, ep.slope);
                        if constexpr (PADDED) {
                            const uint4 q0 = pack_x8(v, ep.dt), q1 = pack_x8(v + 8, ep.dt);
                            if constexpr (POOL) {
                                const uint32_t pk[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
                                for (int i = 0; i < 8; ++i) pm[i] = r == 0 ? pk[i] : max16x2(pm[i], pk[i], ep.dt);
                            }
                            if (ANX_ABL(g, 2)) continue;
                            uint4 *pd = ep.dst.at(n, 0, z + 1, y + 1, x + 1);
                            *pd = q0;
                            pd[gstride] = q1;
                            const int mdy = mirror_delta(y, Hh, rep);
                            if (mdx | mdy | mdz) {
                                store_mirrors(pd, q0, mdz, mdy, mdx, rowp, plane);
                                store_mirrors(pd + gstride, q1, mdz, mdy, mdx, rowp, plane);
                            }
                        } else if constexpr (MODE == EPI_F32) {
                            if (ANX_ABL(g, 2)) continue;
                            // fp32 NCDHW: a warp writes one 128-byte run per channel; with a fused feature
                            // all-gather the same runs go to every rank's gather buffer over NVLink
                            const size_t off = out_off + ((size_t)z * Hh + y) * Ww;
                            if (ep.n_peers == 0) {
                                float *po = ep.out_f32 + off;
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    if (i < ep.cout) *po = v[i];
                                    po += vol;
                                }
                            } else {
                                for (int pr = 0; pr < ep.n_peers; ++pr) {
                                    float *po = ep.out_peers[pr] + off;
#pragma unroll
                                    for (int i = 0; i < 16; ++i) {
                                        if (i < ep.cout) *po = v[i];
                                        po += vol;
                                    }
                                }
                            }
                        } else if constexpr (MODE == EPI_CL16) {
                            if (ANX_ABL(g, 2)) continue;
                            // 16-bit channels-last [N, D, H, W, 16]: 32 contiguous bytes per lane, 1 KB per warp
                            const uint4 q0 = pack_x8(v, ep.dt), q1 = pack_x8(v + 8, ep.dt);
                            const size_t off = ((((size_t)(ep.sample_offset + n) * Dd + z) * Hh + y) * Ww + x) * 2;
                            const int targets = ep.n_peers > 0 ? ep.n_peers : 1;
                            for (int pr = 0; pr < targets; ++pr) {
                                uint4 *po = reinterpret_cast<uint4 *>(ep.n_peers > 0 ? ep.out_peers[pr] : ep.out_f32) + off;
                                po[0] = q0;
                                po[1] = q1;
                            }
                        } else {
                            if (ANX_ABL(g, 2)) continue;
                            const float *hb = sh->shift + HEAD_SMEM_OFFSET, *hw = hb + HEAD_MAX;
                            float *po = ep.out_f32 + out_off + ((size_t)z * Hh + y) * Ww;
#pragma unroll 2
                            for (int k = 0; k < ep.head_nc; ++k) {
                                float a = hb[k];
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float4 w4 = reinterpret_cast<const float4 *>(hw + k * 16)[i];
                                    a = fmaf(w4.x, v[4 * i], a);
                                    a = fmaf(w4.y, v[4 * i + 1], a);
                                    a = fmaf(w4.z, v[4 * i + 2], a);
                                    a = fmaf(w4.w, v[4 * i + 3], a);
                                }
                                po[(size_t)k * vol] = a;
                            }
                        }
                    }
                    if constexpr (POOL) {
                        // 2x2x2 max pooling (reference network.py:297,368) of the values just stored: y pair = this
                        // warp's two rows, z pair = planes (o, o + 1) via `held`, x pair = neighbouring lanes
                        if (o >= 0) {
                            if (!(o & 1)) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) held[s][i] = pm[i];
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    pm[i] = max16x2(pm[i], held[s][i], ep.dt);
                                    pm[i] = max16x2(pm[i], __shfl_xor_sync(0xffffffffu, pm[i], 1), ep.dt);
                                }
                                if (!(lane & 1) && !(ANX_ABL(g, 2)))
                                    store_padded_groups(ep.pool_dst, n, 0, 2, z >> 1, (y0 + s * ROWS_BY + 2 * h) >> 1, x >> 1,
                                                        make_uint4(pm[0], pm[1], pm[2], pm[3]),
                                                        make_uint4(pm[4], pm[5], pm[6], pm[7]));
                            }
                        }
                    }
                    if (p == planes - 1) {          // phantom planes above the segment: neutral again
#pragma unroll
                        for (int r = 0; r < RPW; ++r)
#pragma unroll
                            for (int e = 1; e <= 2; ++e) {
                                __syncwarp();
                                tmem_st16(lane_base + s * ROWS_STRIP_COLS + ((RPW * h + r) * 3 + (slot + e) % 3) * 16, sd);
                            }
                    }
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sh->drained[s][pair]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

}   // namespace anx
