// CUDA-core kernels around the tensor-core conv: the Cin=1 stem conv (reads the
// fp32 NCDHW network input), 2x2x2 pooling, x2 upsampling into a concat buffer,
// and a slow direct conv that consumes the SAME packed weights as the tcgen05
// kernel (debug / cross-check only, selected with ANX_FLAG_FORCE_SIMT).
#pragma once
#include "epilogue.cuh"

namespace anx {

__device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// ------------------------------------------------------------------ stem conv
// in: fp32 [N, CIN, D, H, W]; w: fp32 [CIN][27][COUT] (BN folded).  K = 27*CIN is far
// too thin for a tensor-core tile, so this runs on the CUDA cores: a thread owns ZT
// consecutive z voxels at one (y, x) and all COUT channels (COUT*ZT accumulators);
// lanes run along x so every load and store is coalesced, and each loaded input
// value feeds up to 3 output planes.  FFMA-bound (432 FMA per voxel for 1->16).
template <int COUT, int ZT>
__global__ void __launch_bounds__(256, 2)
stem_conv_kernel(const float *__restrict__ in, const float *__restrict__ w, int cin, int N, int D, int H, int W,
                 int z_halo, Epilogue ep) {
    extern __shared__ float sw[];   // [cin][27][COUT]
    for (int i = threadIdx.x; i < cin * 27 * COUT; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int ztiles = (D + ZT - 1) / ZT;
    const int n = blockIdx.z / ztiles;
    const int z0 = (blockIdx.z - n * ztiles) * ZT;
    const bool valid = (x < W) && (y < H);       // no early exit: the statistics path shuffles across the warp
    // z_halo: the input carries one extra plane at each end of D (depth-slab mode: the
    // neighbour slab's boundary plane, or the caller's reflect copy at a global face)
    const size_t plane = (size_t)(D + 2 * z_halo) * H * W;
    float acc[ZT][COUT];
#pragma unroll
    for (int o = 0; o < ZT; ++o)
#pragma unroll
        for (int i = 0; i < COUT; ++i) acc[o][i] = 0.0f;
    int xx[3], yy[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        xx[k] = reflect_idx(valid ? x + k - 1 : 0, W);
        yy[k] = reflect_idx(valid ? y + k - 1 : 0, H);
    }
    for (int ci = 0; ci < cin; ++ci) {
        const float *src = in + ((size_t)n * cin + ci) * plane;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float *col = src + (size_t)yy[ky] * W + xx[kx];
                float a[ZT + 2];
#pragma unroll
                for (int pz = 0; pz < ZT + 2; ++pz) {
                    int zz = z0 + pz - 1;                          // beyond D only for masked outputs
                    zz = z_halo ? (zz < D ? zz + 1 : D + 1) : reflect_idx(zz < D ? zz : D, D);
                    a[pz] = __ldg(col + (size_t)zz * H * W);
                }
#pragma unroll
                for (int kz = 0; kz < 3; ++kz) {
                    const float *wt = sw + (ci * 27 + (kz * 3 + ky) * 3 + kx) * COUT;
#pragma unroll
                    for (int i = 0; i < COUT; ++i) {
                        const float wv = wt[i];
#pragma unroll
                        for (int o = 0; o < ZT; ++o) acc[o][i] = fmaf(a[o + kz], wv, acc[o][i]);
                    }
                }
            }
    }
    if (ep.stats) {   // instance norm: warp-reduced sums of the raw conv output (lanes = 32 x voxels of sample n)
#pragma unroll
        for (int cb = 0; cb < COUT / 16; ++cb) {
            float s16[16], q16[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { s16[i] = 0.0f; q16[i] = 0.0f; }
#pragma unroll
            for (int o = 0; o < ZT; ++o)
                if (valid && z0 + o < D) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float v = acc[o][cb * 16 + i];
                        s16[i] += v;
                        q16[i] = fmaf(v, v, q16[i]);
                    }
                }
            warp_stats_add(s16, q16, ep.stats + ((size_t)n * ep.stats_stride + cb * 16) * 2);
        }
    }
    Epilogue ep_store = ep;
    ep_store.stats = nullptr;
    if (!valid) return;
#pragma unroll
    for (int o = 0; o < ZT; ++o) {
        if (z0 + o >= D) break;
#pragma unroll
        for (int cb = 0; cb < COUT / 16; ++cb) epilogue_store16(ep_store, n, z0 + o, y, x, cb, acc[o] + cb * 16);
    }
}

// ------------------------------------------------------------- debug direct conv
// One thread per (voxel, block of 16 output channels).  Reads the packed B slabs
// [chunk][group][tap9][kchunk2][R][8] exactly as the tensor-core kernel does.
__global__ void __launch_bounds__(128)
conv3_simt_kernel(ActView src, ConvGeom g, const __nv_bfloat16 *__restrict__ wpack, Epilogue ep) {
    const size_t plane = (size_t)g.D * g.H * g.W;
    const int cblocks = g.ncols * g.n_splits / 16;   // all output channels, every split
    const size_t total = (size_t)g.N * plane * cblocks;
    const size_t slab = (size_t)9 * 2 * g.b_rows * 8;   // elements per (chunk, group)
    for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
        const int cb = (int)(id % cblocks);
        size_t v = id / cblocks;
        const int x = (int)(v % g.W); v /= g.W;
        const int y = (int)(v % g.H); v /= g.H;
        const int z = (int)(v % g.D);
        const int n = (int)(v / g.D);
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
        for (int c = 0; c < g.cin_chunks; ++c)
            for (int kz = 0; kz < 3; ++kz) {
                const int grp = g.fold ? 0 : kz;
                const int rowbase = g.fold ? (2 - kz) * g.ncols : 0;   // blocks ordered dz = +1, 0, -1
                const int split = cb * 16 / g.ncols, cl = cb * 16 - split * g.ncols;   // channel slice and offset in it
                const __nv_bfloat16 *wb = wpack + ((size_t)(split * g.cin_chunks + c) * g.groups + grp) * slab;
                for (int t = 0; t < 9; ++t) {
                    const int ky = t / 3, kx = t % 3;
                    for (int kc = 0; kc < 2; ++kc) {
                        float a[8];
                        unpack_x8(*src.at(n, 2 * c + kc, z + kz, y + ky, x + kx), a, g.dt);
                        const __nv_bfloat16 *wr = wb + ((size_t)(t * 2 + kc) * g.b_rows + rowbase + cl) * 8;
                        for (int i = 0; i < 16; ++i) {
                            float wv[8];
                            unpack_x8(*reinterpret_cast<const uint4 *>(wr + i * 8), wv, g.dt);
#pragma unroll
                            for (int e = 0; e < 8; ++e) acc[i] = fmaf(a[e], wv[e], acc[i]);
                        }
                    }
                }
            }
        epilogue_store16(ep, n, z, y, x, cb, acc);
    }
}

// ---------------------------------------------------------------------- pooling
// 2x2x2 stride-2 max / mean (reference network.py:297,368); one thread per output
// voxel and 8-channel group; writes the pooled tensor with its reflect shell.
__global__ void __launch_bounds__(256)
pool2_kernel(ActView src, ActView dst, int N, int groups, int kind, int dt) {
    const size_t total = (size_t)N * groups * dst.D * dst.H * dst.W;
    for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
        size_t v = id;
        const int x = (int)(v % dst.W); v /= dst.W;
        const int y = (int)(v % dst.H); v /= dst.H;
        const int z = (int)(v % dst.D); v /= dst.D;
        const int gidx = (int)(v % groups);
        const int n = (int)(v / groups);
        float m[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = kind == 0 ? -INFINITY : 0.0f;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    float f[8];
                    unpack_x8(*src.at(n, gidx, 2 * z + a + 1, 2 * y + b + 1, 2 * x + c + 1), f, dt);
#pragma unroll
                    for (int i = 0; i < 8; ++i) m[i] = kind == 0 ? fmaxf(m[i], f[i]) : m[i] + f[i];
                }
        if (kind != 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) m[i] *= 0.125f;
        }
        const uint4 q = pack_x8(m, dt);
        store_padded_groups(dst, n, gidx, 1, z, y, x, q, q);
    }
}

// Grid-mapped form: blockIdx = (output y, output z, n * groups + g), threads along the output x; storage type as a
// template parameter (no per-voxel index decomposition, one arm of the conversion).
template <int DT>
__global__ void __launch_bounds__(128)
pool2_grid_kernel(ActView src, ActView dst, int groups, int kind) {
    const int y = blockIdx.x, z = blockIdx.y;
    const int n = blockIdx.z / groups, gidx = blockIdx.z - n * groups;
    const size_t srow = (size_t)src.pitch, splane = srow * (src.H + 2);
    const size_t drow = (size_t)dst.pitch, dplane = drow * (dst.H + 2);
    const uint4 *s00 = src.at(n, gidx, 2 * z + 1, 2 * y + 1, 1);
    uint4 *prow = dst.at(n, gidx, z + 1, y + 1, 1);
    const int mdy = mirror_delta(y, dst.H, dst.shell_rep), mdz = mirror_delta_z(z, dst.D, dst.shell_rep, dst.z_open);
    for (int x = threadIdx.x; x < dst.W; x += blockDim.x) {
        float m[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = kind == 0 ? -INFINITY : 0.0f;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    float f[8];
                    unpack_x8(__ldg(s00 + (size_t)a * splane + (size_t)b * srow + 2 * x + c), f, DT);
#pragma unroll
                    for (int i = 0; i < 8; ++i) m[i] = kind == 0 ? fmaxf(m[i], f[i]) : m[i] + f[i];
                }
        if (kind != 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) m[i] *= 0.125f;
        }
        const uint4 q = pack_x8(m, DT);
        uint4 *pd = prow + x;
        *pd = q;
        const int mdx = mirror_delta(x, dst.W, dst.shell_rep);
        if (mdx | mdy | mdz) store_mirrors(pd, q, mdz, mdy, mdx, drow, dplane);
    }
}

// -------------------------------------------------------------------- upsampling
// x2 nearest / trilinear (align_corners=False) of `src` written into the group
// range of `dst` (the decoder half of a concat buffer), reference network.py:407,545.
__device__ __forceinline__ void tri_src(int o, int n, int &i0, int &i1, float &t) {
    float s = fmaxf((o + 0.5f) * 0.5f - 0.5f, 0.0f);
    i0 = (int)s;
    i1 = i0 + 1 < n ? i0 + 1 : n - 1;
    t = s - (float)i0;
}

// nearest only: one thread per (low z, low y, HIGH x); it reads its low-res voxel
// once and writes the 2x2 (z, y) copies, so every store is a coalesced 512-byte run.
__global__ void __launch_bounds__(256)
upsample2_nearest_kernel(ActView src, ActView dst, int N, int groups) {
    const size_t total = (size_t)N * groups * src.D * src.H * dst.W;
    for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
        size_t v = id;
        const int x = (int)(v % dst.W); v /= dst.W;
        const int yl = (int)(v % src.H); v /= src.H;
        const int zl = (int)(v % src.D); v /= src.D;
        const int gidx = (int)(v % groups);
        const int n = (int)(v / groups);
        const uint4 q = *src.at(n, gidx, zl + 1, yl + 1, (x >> 1) + 1);
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) store_padded_groups(dst, n, gidx, 1, 2 * zl + a, 2 * yl + b, x, q, q);
    }
}

// Nearest x2, grid-mapped: blockIdx = (low y, low z, n * groups + g), threads along the HIGH-resolution x.  Each thread
// reads its low-resolution voxel once and writes the 2 x 2 (z, y) copies: every warp store is a contiguous 512-byte
// run; strides and the y / z mirror offsets are per block, no per-voxel index decomposition.
__global__ void __launch_bounds__(128)
upsample2_nearest_grid_kernel(ActView src, ActView dst, int groups) {
    const int yl = blockIdx.x, zl = blockIdx.y;
    const int n = blockIdx.z / groups, gidx = blockIdx.z - n * groups;
    const uint4 *srow = src.at(n, gidx, zl + 1, yl + 1, 1);
    const size_t drow = (size_t)dst.pitch, dplane = drow * (dst.H + 2);
    uint4 *d00 = dst.at(n, gidx, 2 * zl + 1, 2 * yl + 1, 1);
    int mdy[2], mdz[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        mdy[a] = mirror_delta(2 * yl + a, dst.H, dst.shell_rep);
        mdz[a] = mirror_delta_z(2 * zl + a, dst.D, dst.shell_rep, dst.z_open);
    }
    for (int x = threadIdx.x; x < dst.W; x += blockDim.x) {
        const uint4 q = __ldg(srow + (x >> 1));
        const int mdx = mirror_delta(x, dst.W, dst.shell_rep);
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                uint4 *pd = d00 + (size_t)a * dplane + (size_t)b * drow + x;
                *pd = q;
                if (mdx | mdy[b] | mdz[a]) store_mirrors(pd, q, mdz[a], mdy[b], mdx, drow, dplane);
            }
    }
}

// z axis of a depth slab (one oversized volume split over several GPUs): at an interior slab face the
// interpolation reads the neighbour's boundary plane from the shell (index -1 / n) instead of clamping.
__device__ __forceinline__ void tri_src_slab(int o, int n, bool lo_open, bool hi_open, int &i0, int &i1, float &t) {
    float s = (o + 0.5f) * 0.5f - 0.5f;
    if (!lo_open) s = fmaxf(s, 0.0f);
    i0 = (int)floorf(s);
    t = s - (float)i0;
    i1 = i0 + 1;
    if (i1 > n - 1 && !hi_open) i1 = n - 1;
}

__global__ void __launch_bounds__(256)
upsample2_kernel(ActView src, ActView dst, int N, int groups, int kind, int dt, int z_lo_open, int z_hi_open) {
    const size_t total = (size_t)N * groups * dst.D * dst.H * dst.W;
    for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
        size_t v = id;
        const int x = (int)(v % dst.W); v /= dst.W;
        const int y = (int)(v % dst.H); v /= dst.H;
        const int z = (int)(v % dst.D); v /= dst.D;
        const int gidx = (int)(v % groups);
        const int n = (int)(v / groups);
        uint4 q;
        if (kind == 0) {
            q = *src.at(n, gidx, (z >> 1) + 1, (y >> 1) + 1, (x >> 1) + 1);
        } else {
            int z0, z1, y0, y1, x0, x1;
            float tz, ty, tx;
            tri_src_slab(z, src.D, z_lo_open != 0, z_hi_open != 0, z0, z1, tz);
            tri_src(y, src.H, y0, y1, ty);
            tri_src(x, src.W, x0, x1, tx);
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = 0.0f;
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const float wgt = (a ? tz : 1.0f - tz) * (b ? ty : 1.0f - ty) * (c ? tx : 1.0f - tx);
                        float f[8];
                        unpack_x8(*src.at(n, gidx, (a ? z1 : z0) + 1, (b ? y1 : y0) + 1, (c ? x1 : x0) + 1), f, dt);
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] = fmaf(wgt, f[i], o[i]);
                    }
            q = pack_x8(o, dt);
        }
        store_padded_groups(dst, n, gidx, 1, z, y, x, q, q);
    }
}

// Trilinear x2, one thread per LOW-resolution voxel and channel group: the 3 x 3 x 3 neighbourhood is read once
// (27 loads and conversions for 8 output voxels instead of 8 each) and interpolated separably -- along x while a
// plane's rows are read, then y, then z:  out[2i] = 0.25 in[i-1] + 0.75 in[i],  out[2i+1] = 0.75 in[i] + 0.25 in[i+1]
// with clamped neighbours (= align_corners=False with the source index clamped at 0, network.py:407), or the
// neighbour slab's plane from the shell at an open z face.  blockIdx = (low y, low z, n * groups + g), threads along
// low x: a warp writes 1 KB contiguous per output row.  (The generic upsample2_kernel spends most of its instructions
// on the 64-bit index decomposition of its grid-stride loop and both arms of the run-time bf16 / fp16 conversion:
// 1.14 ms for the 94M model's last upsample against 0.40 ms here.)
template <int DT>
__global__ void __launch_bounds__(64)
upsample2_tri_block_kernel(ActView src, ActView dst, int groups, int z_lo_open, int z_hi_open) {
    const int yl = blockIdx.x, zl = blockIdx.y;
    const int n = blockIdx.z / groups, gidx = blockIdx.z - n * groups;
    const size_t srow = (size_t)src.pitch, splane = srow * (src.H + 2);
    const size_t drow = (size_t)dst.pitch, dplane = drow * (dst.H + 2);
    // padded source coordinates of the three planes / rows (clamped, or the shell plane at an open z face)
    const int zm = zl > 0 ? zl : (z_lo_open ? 0 : 1), zp = zl + 1 < src.D ? zl + 2 : (z_hi_open ? src.D + 1 : src.D);
    const int zs[3] = {zm, zl + 1, zp};
    const int ys[3] = {yl > 0 ? yl : 1, yl + 1, yl + 1 < src.H ? yl + 2 : src.H};
    const uint4 *sbase = src.at(n, gidx, 0, 0, 0);
    uint4 *d00 = dst.at(n, gidx, 2 * zl + 1, 2 * yl + 1, 1);
    int mdy[2], mdz[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        mdy[a] = mirror_delta(2 * yl + a, dst.H, dst.shell_rep);
        mdz[a] = mirror_delta_z(2 * zl + a, dst.D, dst.shell_rep, dst.z_open);
    }
    for (int xl = threadIdx.x; xl < src.W; xl += blockDim.x) {
        const int xs[3] = {xl > 0 ? xl : 1, xl + 1, xl + 1 < src.W ? xl + 2 : src.W};
        float P[2][4][8];                           // xy-interpolated planes: [which][b * 2 + c][channel]
        auto plane_xy = [&](int k, float (&out)[4][8]) {
            float R[3][2][8];                       // per source row: the two x outputs
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const uint4 *row = sbase + (size_t)zs[k] * splane + (size_t)ys[j] * srow;
                float a[8], b[8], c[8];
                unpack_x8(__ldg(row + xs[0]), a, DT);
                unpack_x8(__ldg(row + xs[1]), b, DT);
                unpack_x8(__ldg(row + xs[2]), c, DT);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    R[j][0][i] = fmaf(0.75f, b[i], 0.25f * a[i]);
                    R[j][1][i] = fmaf(0.75f, b[i], 0.25f * c[i]);
                }
            }
#pragma unroll
            for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    out[0 * 2 + cc][i] = fmaf(0.75f, R[1][cc][i], 0.25f * R[0][cc][i]);
                    out[1 * 2 + cc][i] = fmaf(0.75f, R[1][cc][i], 0.25f * R[2][cc][i]);
                }
        };
        auto emit = [&](int a, const float (&near)[4][8], const float (&far)[4][8]) {   // 0.75 near + 0.25 far
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    float o[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = fmaf(0.75f, near[b * 2 + cc][i], 0.25f * far[b * 2 + cc][i]);
                    const uint4 q = pack_x8(o, DT);
                    const int x = 2 * xl + cc;
                    uint4 *pd = d00 + (size_t)a * dplane + (size_t)b * drow + x;
                    *pd = q;
                    const int mdx = mirror_delta(x, dst.W, dst.shell_rep);
                    if (mdx | mdy[b] | mdz[a]) store_mirrors(pd, q, mdz[a], mdy[b], mdx, drow, dplane);
                }
        };
        plane_xy(0, P[0]);
        plane_xy(1, P[1]);
        emit(0, P[1], P[0]);
        plane_xy(2, P[0]);
        emit(1, P[1], P[0]);
    }
}

// ------------------------------------------------------------------ tap export
// One stored tensor (padded planar 16-bit) -> fp32 NCDHW, for `forward(layers=[...])` feature taps
// (reference network.py:475-529).  One thread per (n, group, z, y, x), lanes along x: a warp reads 512
// contiguous bytes and writes eight 128-byte runs.
__global__ void __launch_bounds__(256)
export_ncdhw_kernel(ActView src, int N, int groups, int channels, float *__restrict__ out, int dt) {
    const size_t vol = (size_t)src.D * src.H * src.W;
    const size_t total = (size_t)N * groups * vol;
    for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
        size_t v = id;
        const int x = (int)(v % src.W); v /= src.W;
        const int y = (int)(v % src.H); v /= src.H;
        const int z = (int)(v % src.D); v /= src.D;
        const int gidx = (int)(v % groups);
        const int n = (int)(v / groups);
        float f[8];
        unpack_x8(*src.at(n, gidx, z + 1, y + 1, x + 1), f, dt);
        float *o = out + ((size_t)n * channels + gidx * 8) * vol + ((size_t)z * src.H + y) * src.W + x;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (gidx * 8 + i < channels) o[(size_t)i * vol] = f[i];
    }
}

// ------------------------------------------------------- scaled average pooling
// out = scale * avg_pool3d(in, k, stride=k) on fp32 [NC, D, H, W] (floor mode): the feature post-processing
// of the registration caller (reference run_convex_adam_with_network_feats.py:166-167, 198-205).  One
// thread per output voxel, lanes along x; streaming, HBM-bound (reads every input byte once).
__global__ void __launch_bounds__(256)
avgpool3d_scale_kernel(const float *__restrict__ in, float *__restrict__ out, size_t NC, int D, int H, int W, int k,
                       float scale) {
    const int Do = D / k, Ho = H / k, Wo = W / k;
    const size_t total = NC * Do * Ho * Wo;
    const float norm = scale / (float)(k * k * k);
    for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
        size_t v = id;
        const int x = (int)(v % Wo); v /= Wo;
        const int y = (int)(v % Ho); v /= Ho;
        const int z = (int)(v % Do);
        const size_t nc = v / Do;
        const float *src = in + ((nc * D + (size_t)z * k) * H + (size_t)y * k) * W + (size_t)x * k;
        float a = 0.0f;
        for (int dz = 0; dz < k; ++dz)
            for (int dy = 0; dy < k; ++dy) {
                const float *row = src + ((size_t)dz * H + dy) * W;
                for (int dx = 0; dx < k; ++dx) a += __ldg(row + dx);
            }
        out[id] = a * norm;
    }
}

// k = 2, W a multiple of 4, 16-byte aligned input: one thread per PAIR of output voxels, 128-bit loads
// (a warp reads four 512-byte runs and writes 256 contiguous bytes).
__global__ void __launch_bounds__(256)
avgpool3d_scale_k2_kernel(const float *__restrict__ in, float *__restrict__ out, size_t NC, int D, int H, int W,
                          float scale) {
    const int Do = D / 2, Ho = H / 2, Wq = W / 4;
    const size_t total = NC * Do * Ho * Wq;
    const float norm = scale * 0.125f;
    for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
        size_t v = id;
        const int xq = (int)(v % Wq); v /= Wq;
        const int y = (int)(v % Ho); v /= Ho;
        const int z = (int)(v % Do);
        const size_t nc = v / Do;
        const float *src = in + ((nc * D + (size_t)z * 2) * H + (size_t)y * 2) * W + (size_t)xq * 4;
        float a = 0.0f, b = 0.0f;
#pragma unroll
        for (int dz = 0; dz < 2; ++dz)
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const float4 p = __ldg(reinterpret_cast<const float4 *>(src + ((size_t)dz * H + dy) * W));
                a += p.x + p.y;
                b += p.z + p.w;
            }
        *reinterpret_cast<float2 *>(out + ((nc * Do + z) * Ho + y) * (size_t)(W / 2) + (size_t)xq * 2) =
            make_float2(a * norm, b * norm);
    }
}

// ------------------------------------------------- voxelwise channel normalisation
// Per voxel, across the C channels of fp32 [N, C, D, H, W] features (reference README.md:13,49: "voxelwise normalize the
// features across channels to have unit norm or zero mean with unit standard deviation" before registration /
// visualisation with the dev models):
//   mode 0: y = x / max(||x||_2, eps)                       (torch.nn.functional.normalize(x, dim=1))
//   mode 1: y = (x - mean_c) / (std_c + eps), unbiased std  ((x - x.mean(1, True)) / x.std(1, keepdim=True))
// One thread per voxel, lanes along x: every channel row is read and written in 128-byte runs; two passes over the
// channels (statistics, then scale) -- the second pass hits L1 / L2.  In place (out == in) is allowed.
__global__ void __launch_bounds__(256)
channel_normalize_kernel(const float *__restrict__ in, float *__restrict__ out, size_t n_samples, size_t vol, int C,
                         int mode, float eps) {
    const size_t total = n_samples * vol;
    for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
        const size_t n = id / vol, v = id - n * vol;
        const float *src = in + n * (size_t)C * vol + v;
        float *dst = out + n * (size_t)C * vol + v;
        float s = 0.0f, q = 0.0f;
        for (int c = 0; c < C; ++c) {
            const float x = src[(size_t)c * vol];
            s += x;
            q = fmaf(x, x, q);
        }
        float shift = 0.0f, scale;
        if (mode == 0) {
            scale = 1.0f / fmaxf(sqrtf(q), eps);
        } else {
            shift = s / (float)C;
            const float var = fmaxf(q - s * shift, 0.0f) / (float)(C > 1 ? C - 1 : 1);
            scale = 1.0f / (sqrtf(var) + eps);
        }
        for (int c = 0; c < C; ++c) dst[(size_t)c * vol] = (src[(size_t)c * vol] - shift) * scale;
    }
}

// ------------------------------------------------------- sliding-window blend
// One window of a sliding-window scan (MONAI-style inferer, reference convex_adam_utils.py:202-219):
//   out[c, z0+z, y0+y, x0+x] += pred[c, z, y, x] * weight[z, y, x]      norm[z0+z, y0+y, x0+x] += weight[z, y, x]
// One thread per window voxel, lanes along x, all channels in a loop (the weight is read once).  Windows of
// one scan overlap, so the caller launches them one after another on one stream.
__global__ void __launch_bounds__(256)
blend_window_kernel(const float *__restrict__ pred, const float *__restrict__ weight, float *__restrict__ out,
                    float *__restrict__ norm, int C, int d, int h, int w, int D, int H, int W, int z0, int y0, int x0) {
    const size_t wvol = (size_t)d * h * w, vol = (size_t)D * H * W;
    for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < wvol; id += (size_t)gridDim.x * blockDim.x) {
        size_t v = id;
        const int x = (int)(v % w); v /= w;
        const int y = (int)(v % h);
        const int z = (int)(v / h);
        const float g = __ldg(weight + id);
        const size_t o = ((size_t)(z0 + z) * H + (y0 + y)) * W + (x0 + x);
        norm[o] += g;
        for (int c = 0; c < C; ++c) out[(size_t)c * vol + o] = fmaf(__ldg(pred + (size_t)c * wvol + id), g, out[(size_t)c * vol + o]);
    }
}

// ---------------------------------------------------------- instance norm + act
// In place over one tensor inside a padded planar buffer, shell included (the shell
// holds mirror copies, so normalising it element-wise equals mirroring afterwards):
//   y = act((x - mean[n,c]) * rsqrt(var[n,c] + eps)),  biased variance
// (nn.InstanceNorm3d(affine=False), reference network.py:157-158).  mean / var come
// from the double-precision sums the producing conv accumulated from its fp32
// accumulators.  blockIdx.y = n * groups + g; blockIdx.x strides over the plane.
__global__ void __launch_bounds__(256)
inorm_act_kernel(ActView t, int groups, const double *__restrict__ stats, int stats_stride, double inv_count,
                 float eps, int act, float slope, int dt) {
    __shared__ float s_mean[8], s_rstd[8];
    const int n = blockIdx.y / groups, gidx = blockIdx.y - n * groups;
    if (threadIdx.x < 8) {
        const double *st = stats + ((size_t)n * stats_stride + gidx * 8 + threadIdx.x) * 2;
        const double m = st[0] * inv_count;
        double var = st[1] * inv_count - m * m;
        var = var > 0.0 ? var : 0.0;
        s_mean[threadIdx.x] = (float)m;
        s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    float mean[8], rstd[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { mean[i] = s_mean[i]; rstd[i] = s_rstd[i]; }
    const size_t count = (size_t)(t.D + 2) * (t.H + 2) * (t.W + 2);
    const unsigned wp = (unsigned)(t.W + 2), pitch = (unsigned)t.pitch;
    uint4 *base = t.at(n, gidx, 0, 0, 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const size_t row = i / wp;
        uint4 *cell = base + row * pitch + (i - row * wp);      // rows carry unused lead / tail voxels
        float v[8];
        unpack_x8(*cell, v, dt);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = activate((v[k] - mean[k]) * rstd[k], act, slope);
        *cell = pack_x8(v, dt);
    }
}

}   // namespace anx
