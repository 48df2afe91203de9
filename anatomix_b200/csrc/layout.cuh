// Activation layout in HBM and the parameter blocks shared by the kernels.
//
// Every intermediate activation lives in a reflect-PADDED, channel-group-planar
// bf16 buffer:
//
//      [n][g][zp][yp][lead + xp][8]      zp in [0, D+2), yp in [0, H+2), xp in [0, W+2)
//
// with `pitch` voxels per row and `lead` unused voxels in front of the x shell.  The default is the dense
// form (lead 0, pitch W + 2).  With lead 1 (pitch W + 4) interior voxel x = 0 sits on a 32-byte sector
// boundary and an 8-voxel tile row stored by the conv epilogue is four FULL sectors instead of three full
// and two half ones; lead 7 aligns tile rows to 128-byte lines.  The epilogue's store pattern ALONE drains
// at 2.7 / 5.7 / 5.3 TB/s for lead 0 / 1 / 7 (tools/micro/store_bw.cu on B200), but inside the conv kernel
// the stores hide behind the MMA stream and the halo-brick loads grow from 5 to 6 sectors per row, so the
// whole forward measured 2080 / 2060 / 2070 volumes/s on one box: the variants stay selectable
// (ANX_X_LEAD) for kernels whose MMA phase gets short enough to expose the stores.
//
// g indexes groups of 8 channels (16 bytes per voxel per group).  The one-voxel
// shell holds the reflect padding of nn.Conv3d(padding_mode='reflect')
// (reference network.py:310-318): shell[0] = interior[1], shell[S+1] =
// interior[S-2], written by the PRODUCER of the tensor, so every conv is a plain
// "valid" 3x3x3 correlation on the padded buffer.  16-byte voxels with x
// contiguous are exactly the 8x16-byte "core matrix" rows a K-major,
// non-swizzled tcgen05 shared-memory descriptor wants, so a TMA box copy of a
// halo brick is directly a valid A operand for all 27 taps.
//
// A skip concat (network.py:545) is zero-copy: the encoder conv and the
// upsample write disjoint group ranges [0, Cskip/8) and [Cskip/8, ...) of one
// buffer.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

// Timing experiments (ablation bits in *Geom::ablate, ANX_* environment switches in engine.cu) exist only in
// builds made with -DANX_EXPERIMENTS (ANX_LIB_VARIANT builds, anatomix_b200/build.py): they make launches
// skip work and return WRONG results, so the product library compiles them out.
#ifdef ANX_EXPERIMENTS
#define ANX_ABL(g, bit) (((g).ablate & (bit)) != 0)
#else
#define ANX_ABL(g, bit) false
#endif

namespace anx {

// Row geometry of a padded planar buffer of interior width W for a given lead (host side picks the lead).
struct RowLayout { int lead, pitch; };
__host__ __device__ __forceinline__ RowLayout layout_of(int W, int lead) {
    if (lead <= 0) return RowLayout{0, W + 2};
    if (lead == 1) return RowLayout{1, (W + 4) & ~1};
    return RowLayout{7, (W + 16) & ~7};
}

struct ActView {            // one tensor inside a padded planar buffer
    __nv_bfloat16 *base;    // start of the whole buffer
    int groups_total;       // channel groups per sample in the buffer
    int group_offset;       // first group of this tensor
    int D, H, W;            // interior size
    int shell_rep;          // shell semantics written by the producer: 0 reflect (x[1] / x[S-2]), 1 replicate (x[0] / x[S-1])
    int z_open;             // depth-slab mode: bit 0 / bit 1 = the lower / upper z face borders a neighbouring slab.  That
                            // shell plane belongs to the NEIGHBOUR (its boundary plane is delivered into it, possibly
                            // before this slab's own producer has run), so the producer writes no mirror copy there.
    int lead, pitch;        // row geometry: voxels in front of the x shell, voxels per row (layout_of)
    __host__ __device__ __forceinline__ size_t voxel_index(int n, int g, int zp, int yp, int xp) const {
        return ((((size_t)n * groups_total + (group_offset + g)) * (D + 2) + zp) * (H + 2) + yp) * (size_t)pitch +
               lead + xp;
    }
    __host__ __device__ __forceinline__ uint4 *at(int n, int g, int zp, int yp, int xp) const {
        return reinterpret_cast<uint4 *>(base) + voxel_index(n, g, zp, yp, xp);
    }
};

enum OutMode : int {
    OUT_PADDED_BF16 = 0,    // next layer's padded planar buffer (+ reflect shell)
    OUT_NCDHW_F32 = 1       // network output, fp32 [N, C, D, H, W]
};

enum StoreType : int { DT_BF16 = 0, DT_FP16 = 1 };   // 16-bit storage type of activations / packed weights

struct Epilogue {
    int mode;
    ActView dst;            // OUT_PADDED_BF16 (name kept; holds bf16 or fp16 per `dt`)
    float *out_f32;         // OUT_NCDHW_F32
    int cout;               // real output channels of the whole conv
    const float *bias;      // [cout rounded up to 16] folded BN shift / conv bias (zeros if none)
    int act;                // 0 none, 1 relu, 2 leaky
    float slope;
    int dt;                 // StoreType of dst
    double *stats;          // instance-norm sums [N][cout_padded][2] (sum, sum of squares) or nullptr
    int stats_stride;       // cout rounded up to 16 (channels per sample in `stats`)
    // Fused feature all-gather (OUT_NCDHW_F32 only): when n_peers > 0 the output of local sample n is
    // stored into EVERY peer's gather buffer at sample index sample_offset + n (peer pointers are
    // NVLink-mapped device addresses; out_f32 is ignored).
    float *out_peers[8];
    int n_peers;
    int sample_offset;
    // OUT_NCDHW_F32 only: 1 = store the network output as 16-bit channels-last [N, D, H, W, cout] (type `dt`) instead
    // of fp32 NCDHW -- the compact payload of the feature all-gather and of the host download.
    int cl16;
    // fp32 NCDHW output: floats between consecutive samples (>= channels written * D*H*W).  Larger than that when the
    // output is a channel slice of a wider tensor (zero-copy concat with other features, anx_engine_forward_concat).
    size_t out_nstride;
    // Fused 2x2x2 pooling (tensor-core kernel only): besides the full-resolution store, the epilogue
    // reduces each 2x2x2 block (z pair in registers, y / x pairs by warp shuffles) and writes the
    // pooled tensor with its shell.  pool_kind: -1 off, 0 max, 1 mean.
    int pool_kind;
    ActView pool_dst;
    // Depth-to-space store (low-resolution half of a decoder conv, see engine.cu "upconv"): the conv's
    // columns are (parity, channel) with d2s_cout channels per parity p = a*4 + b*2 + c, and the 16
    // channels of a chunk go to voxel (2z+a, 2y+b, 2x+c) of `dst`, a HIGH-resolution partial-sum tensor.
    int d2s_cout;
    // Accumulator seeding from a stored tensor instead of the channel shift (the skip half of that conv):
    // the partial sums written by the depth-to-space launch.
    int seed_on;
    ActView seed_src;
    // Linear head fused behind the last conv (OUT_NCDHW_F32, cout <= 16): out[k] = head[k] + sum_c
    // head[HEAD_MAX + k*16 + c] * y[c] for k < head_nc, written as fp32 [N, head_nc, D, H, W].
    int head_nc;
    const float *head;      // device: [HEAD_MAX] bias then [HEAD_MAX][16] weights
};

constexpr int HEAD_MAX = 32;                            // most output channels a fused head may have
constexpr int HEAD_FLOATS = HEAD_MAX + HEAD_MAX * 16;   // bias + weights

// tile geometry of the tensor-core conv: 8 (x) x 16 (y) voxels per MMA (M = 128),
// `bz` output planes per CTA tile.
constexpr int TILE_X = 8;
constexpr int TILE_Y = 16;
constexpr int HALO_X = TILE_X + 2;      // 10
constexpr int HALO_Y = TILE_Y + 2;      // 18
constexpr int ROW_BYTES = HALO_X * 16;  // 160: one x-row of the halo brick per group

struct ConvGeom {
    int N, D, H, W;
    int tiles_x, tiles_y, tiles_z, tiles_per_sample, total_tiles;
    int bz;                 // output planes per tile
    int cin_chunks;         // Cin / 16
    int in_groups_total;    // groups per sample in the INPUT buffer (TMA dim 3 = n * this + g)
    int in_group_offset;
    int ncols;              // output channels handled per CTA tile, multiple of 16 (TMEM columns per plane)
    int n_splits;           // Cout > 256 is split over CTAs: tile -> (spatial tile, split); channel0 = split*ncols
    int dt;                 // StoreType of the input activations and packed weights
    int fold;               // 1: one MMA covers the three dz taps (N = 3*ncols)
    int groups;             // B stages per chunk: 1 when folded, 3 (one per dz) otherwise
    int acc_stages;         // TMEM accumulator double buffering
    int tmem_cols;          // power of two >= acc_stages * bz * ncols
    int a_stages, b_stages;
    uint32_t a_stage_bytes; // 2 * (bz+2) * HALO_Y * ROW_BYTES
    uint32_t a_lbo;         // (bz+2) * HALO_Y * ROW_BYTES : distance between the two 8-channel planes
    uint32_t b_stage_bytes; // 9 * 32 * R, R = rows per tap matrix
    uint32_t b_rows;        // R = fold ? 3*ncols : ncols
    uint32_t smem_bytes;
    int fuse_pool;          // this launch also writes the pooled tensor (needs bz % 4 == 0)
    int b_static;           // 1: one B slab serves every tile; loaded once per CTA, never recycled
    int alt;                // 1: the layer's second packing (64-column splits, unfolded) is in use for this small problem
    int stats_acc;          // EPI_STATS on a thin layer: per-warp sums in shared memory behind UmmaShared (STATS_ACC_BYTES more)
    uint32_t ablate;        // timing experiments only (ANX_ABLATE): 1 no MMA, 2 no stores, 4 no A load, 8 no B load,
                            // 16 every tap reads the brick origin, 32 128-byte aligned core matrices (results are wrong)
    // Column trimming (unfolded layers whose packed weights are structurally sparse: the low-resolution half of a
    // decoder conv, engine.cu "upconv"): tap (dz group, dy*3+dx) only touches output columns
    // [trim_lo, trim_lo + trim_n) -- the other columns of its B tile are zeros -- so its MMA is issued with N = trim_n
    // on that column range.  trim = 0: every tap covers all ncols columns.
    int trim;
    uint8_t trim_lo[27], trim_n[27];   // in units of 16 columns
    // trim == 2: the trimmed weights of the WHOLE layer are resident in shared memory (loaded once per CTA): tap
    // (chunk, dz group, dy*3+dx) is a compact K-major tile of trim_n rows at trim_off (16-byte units) from the image.
    uint16_t trim_off[4 * 27];
    uint32_t trim_bytes;
};

}   // namespace anx
