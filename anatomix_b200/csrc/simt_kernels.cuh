// CUDA-core kernels around the tensor-core conv: the Cin=1 stem conv (reads the
// fp32 NCDHW network input), 2x2x2 pooling, x2 upsampling into a concat buffer,
// and a slow direct conv that consumes the SAME packed weights as the tcgen05
// kernel (debug / cross-check only, selected with ANX_FLAG_FORCE_SIMT).
#pragma once
#include "epilogue.cuh"

namespace anx {

__device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// ------------------------------------------------------------------ stem conv
// in: fp32 [N, CIN, D, H, W]; w: fp32 [CIN][27][COUT] (BN folded); one thread
// per voxel computes all COUT channels.  Memory-bound (reads 4 B, writes
// 2*COUT B per voxel), so CUDA cores are the right tool (K = 27 is far too thin
// for a tensor-core tile).
template <int COUT>
__global__ void __launch_bounds__(256)
stem_conv_kernel(const float *__restrict__ in, const float *__restrict__ w, int cin, int N, int D, int H, int W,
                 Epilogue ep) {
    extern __shared__ float sw[];   // [cin][27][COUT]
    for (int i = threadIdx.x; i < cin * 27 * COUT; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const size_t plane = (size_t)D * H * W;
    const size_t total = (size_t)N * plane;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(v / plane);
        size_t r = v - (size_t)n * plane;
        const int z = (int)(r / ((size_t)H * W));
        r -= (size_t)z * H * W;
        const int y = (int)(r / W), x = (int)(r - (size_t)y * W);
        float acc[COUT];
#pragma unroll
        for (int i = 0; i < COUT; ++i) acc[i] = 0.0f;
        for (int ci = 0; ci < cin; ++ci) {
            const float *src = in + ((size_t)n * cin + ci) * plane;
#pragma unroll
            for (int kz = 0; kz < 3; ++kz) {
                const int zz = reflect_idx(z + kz - 1, D);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const int yy = reflect_idx(y + ky - 1, H);
                    const float *row = src + ((size_t)zz * H + yy) * W;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float a = __ldg(row + reflect_idx(x + kx - 1, W));
                        const float *wt = sw + (ci * 27 + (kz * 3 + ky) * 3 + kx) * COUT;
#pragma unroll
                        for (int i = 0; i < COUT; ++i) acc[i] = fmaf(a, wt[i], acc[i]);
                    }
                }
            }
        }
#pragma unroll
        for (int cb = 0; cb < COUT / 16; ++cb) epilogue_store16(ep, n, z, y, x, cb, acc + cb * 16);
    }
}

// ------------------------------------------------------------- debug direct conv
// One thread per (voxel, block of 16 output channels).  Reads the packed B slabs
// [chunk][group][tap9][kchunk2][R][8] exactly as the tensor-core kernel does.
__global__ void __launch_bounds__(128)
conv3_simt_kernel(ActView src, ConvGeom g, const __nv_bfloat16 *__restrict__ wpack, Epilogue ep) {
    const size_t plane = (size_t)g.D * g.H * g.W;
    const int cblocks = g.ncols / 16;
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
                const __nv_bfloat16 *wb = wpack + (size_t)(c * g.groups + grp) * slab;
                for (int t = 0; t < 9; ++t) {
                    const int ky = t / 3, kx = t % 3;
                    for (int kc = 0; kc < 2; ++kc) {
                        float a[8];
                        unpack_bf16x8(*src.at(n, 2 * c + kc, z + kz, y + ky, x + kx), a);
                        const __nv_bfloat16 *wr = wb + ((size_t)(t * 2 + kc) * g.b_rows + rowbase + cb * 16) * 8;
                        for (int i = 0; i < 16; ++i) {
                            float wv[8];
                            unpack_bf16x8(*reinterpret_cast<const uint4 *>(wr + i * 8), wv);
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
pool2_kernel(ActView src, ActView dst, int N, int groups, int kind) {
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
                    unpack_bf16x8(*src.at(n, gidx, 2 * z + a + 1, 2 * y + b + 1, 2 * x + c + 1), f);
#pragma unroll
                    for (int i = 0; i < 8; ++i) m[i] = kind == 0 ? fmaxf(m[i], f[i]) : m[i] + f[i];
                }
        if (kind != 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) m[i] *= 0.125f;
        }
        const uint4 q = pack_bf16x8(m);
        store_padded_groups(dst, n, gidx, 1, z, y, x, q, q);
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

__global__ void __launch_bounds__(256)
upsample2_kernel(ActView src, ActView dst, int N, int groups, int kind) {
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
            tri_src(z, src.D, z0, z1, tz);
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
                        unpack_bf16x8(*src.at(n, gidx, (a ? z1 : z0) + 1, (b ? y1 : y0) + 1, (c ? x1 : x0) + 1), f);
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] = fmaf(wgt, f[i], o[i]);
                    }
            q = pack_bf16x8(o);
        }
        store_padded_groups(dst, n, gidx, 1, z, y, x, q, q);
    }
}

}   // namespace anx
