// Shared epilogue: bias + activation + store of one voxel's 16 output channels,
// either into the next layer's reflect-padded planar bf16 buffer (writing the
// mirrored shell copies as well) or into the fp32 NCDHW network output.
#pragma once
#include <cuda_fp16.h>

#include "layout.cuh"

namespace anx {

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}
// hi/lo split of two fp32 values into packed bf16 pairs: hi = rn(x), lo = rn(x - hi)
__device__ __forceinline__ void split_hi_lo(float x0, float x1, uint32_t &hi, uint32_t &lo) {
    hi = pack_bf16x2(x0, x1);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    lo = pack_bf16x2(x0 - h0, x1 - h1);
}
// 8 floats -> 16 bytes of the storage type (round to nearest even)
__device__ __forceinline__ uint4 pack_x8(const float *v, int dt) {
    if (dt == DT_BF16)
        return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                          pack_bf16x2(v[6], v[7]));
    return make_uint4(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]), pack_f16x2(v[4], v[5]),
                      pack_f16x2(v[6], v[7]));
}
__device__ __forceinline__ void unpack_x8(const uint4 &q, float *v, int dt) {
    if (dt == DT_BF16) {
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&q);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __bfloat1622float2(h[i]);
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    } else {
        const __half2 *h = reinterpret_cast<const __half2 *>(&q);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __half22float2(h[i]);
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
}

// element-wise max of two packed 16-bit pairs of the storage type
__device__ __forceinline__ uint32_t max16x2(uint32_t a, uint32_t b, int dt) {
    if (dt == DT_BF16) {
        __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162 *>(&a), *reinterpret_cast<__nv_bfloat162 *>(&b));
        return *reinterpret_cast<uint32_t *>(&r);
    }
    __half2 r = __hmax2(*reinterpret_cast<__half2 *>(&a), *reinterpret_cast<__half2 *>(&b));
    return *reinterpret_cast<uint32_t *>(&r);
}

// Warp-level instance-norm statistics: every lane brings 16 per-channel partial sums
// `s` and 16 partial sums of squares `q` (its own voxels).  A butterfly
// reduce-scatter (31 shuffles for the 32 values) leaves lane l with the warp total of
// value l, which it adds to `stats16[(l & 15) * 2 + (l >> 4)]` in double precision.
// `acc` (optional): this warp's private 32-double row in shared memory for the chunk -- lane l adds its total there
// instead of issuing a global atomic; the kernel flushes the rows when the sample changes and at its end
// (STATS_ACC_MAX_COLS).  With one atomic per warp, chunk and TILE the thin 128^3 layers issued 4 M double atomics
// onto 256 addresses per launch.
constexpr int STATS_ACC_MAX_COLS = 64;      // per-CTA accumulation rows exist for layers of at most this many columns
__device__ __forceinline__ void warp_stats_add(const float *s, const float *q, double *stats16, double *acc = nullptr) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 16; ++i) { v[i] = s[i]; v[16 + i] = q[i]; }
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < m; ++i) {
            const float keep = up ? v[i + m] : v[i];
            const float send = up ? v[i] : v[i + m];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
    }
    if (acc) acc[lane] += (double)v[0];
    else atomicAdd(stats16 + (lane & 15) * 2 + (lane >> 4), (double)v[0]);
}
// Adds one warp's accumulated rows (`chunks` x 32 doubles, layout of warp_stats_add) for sample n to the global
// sums and clears them.
__device__ __forceinline__ void warp_stats_flush(double *acc, int chunks, double *stats, int stats_stride, int n) {
    const int lane = threadIdx.x & 31;
    for (int cb = 0; cb < chunks; ++cb) {
        const double v = acc[cb * 32 + lane];
        if (v != 0.0) atomicAdd(stats + ((size_t)n * stats_stride + cb * 16) * 2 + (lane & 15) * 2 + (lane >> 4), v);
        acc[cb * 32 + lane] = 0.0;
    }
}

// Padded-coordinate targets of interior coordinate v on an axis of size S:
// always v+1 (k = 0); additionally the shell cell 0 when v == 1 (k = 1) and S+1
// when v == S-2 (k = 2) (reflect: shell[0] = x[1], shell[S+1] = x[S-2]).  -1 = none.
// With replicate shells (rep = 1) the sources are x[0] and x[S-1] instead.
__device__ __forceinline__ int mirror_target(int v, int S, int k, int rep = 0) {
    if (k == 0) return v + 1;
    if (k == 1) return v == (rep ? 0 : 1) ? 0 : -1;
    return v == (rep ? S - 1 : S - 2) ? S + 1 : -1;
}

// Shell copies of one 16-byte voxel group whose interior store went to `p`:
// reflect padding puts x[1] into shell cell 0 (padded index v+1-2) and x[S-2] into
// shell cell S+1 (padded index v+1+2), so every mirror copy sits at -2 / +2 cells
// along each mirrored axis.  d* are those offsets in cells (0 = axis not mirrored);
// valid for S >= 4 (a voxel then mirrors to at most one side per axis).
__device__ __forceinline__ void store_mirrors(uint4 *p, const uint4 &q, int dz, int dy, int dx, size_t row,
                                              size_t plane) {
    if (dx) p[dx] = q;
    if (dy) {
        uint4 *py = p + (ptrdiff_t)dy * (ptrdiff_t)row;
        *py = q;
        if (dx) py[dx] = q;
    }
    if (dz) {
        uint4 *pz = p + (ptrdiff_t)dz * (ptrdiff_t)plane;
        *pz = q;
        if (dx) pz[dx] = q;
        if (dy) {
            uint4 *pzy = pz + (ptrdiff_t)dy * (ptrdiff_t)row;
            *pzy = q;
            if (dx) pzy[dx] = q;
        }
    }
}
__device__ __forceinline__ int mirror_delta(int v, int S, int rep = 0) {
    if (rep) return v == 0 ? -1 : (v == S - 1 ? 1 : 0);      // replicate: shell cell next to the border voxel
    return v == 1 ? -2 : (v == S - 2 ? 2 : 0);
}
// z axis: no mirror copy into a shell plane that a neighbouring depth slab owns (ActView::z_open)
__device__ __forceinline__ int mirror_delta_z(int v, int S, int rep, int open) {
    const int d = mirror_delta(v, S, rep);
    return (d < 0 && (open & 1)) || (d > 0 && (open & 2)) ? 0 : d;
}

// Stores `ngroups` (1 or 2) packed 8-channel groups of voxel (n,z,y,x) starting at
// group g0 into a padded planar buffer, including its reflect-shell copies.
__device__ __forceinline__ void store_padded_groups(const ActView &dst, int n, int g0, int ngroups, int z, int y,
                                                    int x, const uint4 &q0, const uint4 &q1) {
    const size_t row = (size_t)dst.pitch, plane = row * (dst.H + 2), gstride = plane * (dst.D + 2);
    uint4 *p = dst.at(n, g0, z + 1, y + 1, x + 1);
    if (dst.D >= 4 && dst.H >= 4 && dst.W >= 4) {
        const int dz = mirror_delta_z(z, dst.D, dst.shell_rep, dst.z_open), dy = mirror_delta(y, dst.H, dst.shell_rep),
                  dx = mirror_delta(x, dst.W, dst.shell_rep);
        *p = q0;
        if (ngroups > 1) p[gstride] = q1;
        if (dz | dy | dx) {
            store_mirrors(p, q0, dz, dy, dx, row, plane);
            if (ngroups > 1) store_mirrors(p + gstride, q1, dz, dy, dx, row, plane);
        }
        return;
    }
    // tiny tensors (a size-2 or size-3 axis mirrors one voxel to both sides): generic loops
    for (int a = 0; a < 3; ++a) {
        const int zt = mirror_target(z, dst.D, a, dst.shell_rep);
        if (zt < 0 || (a == 1 && (dst.z_open & 1)) || (a == 2 && (dst.z_open & 2))) continue;
        for (int b = 0; b < 3; ++b) {
            const int yt = mirror_target(y, dst.H, b, dst.shell_rep);
            if (yt < 0) continue;
            for (int c = 0; c < 3; ++c) {
                const int xt = mirror_target(x, dst.W, c, dst.shell_rep);
                if (xt < 0) continue;
                *dst.at(n, g0, zt, yt, xt) = q0;
                if (ngroups > 1) *dst.at(n, g0 + 1, zt, yt, xt) = q1;
            }
        }
    }
}

__device__ __forceinline__ float activate(float v, int act, float slope) {
    if (act == 1) return fmaxf(v, 0.0f);
    if (act == 2) return v > 0.0f ? v : v * slope;
    return v;
}

// v[16]: raw accumulators of channels [16*cb, 16*cb+16) of voxel (n,z,y,x).  Used by
// the CUDA-core kernels (stem, debug conv); statistics go out one atomic per value
// there, the tensor-core kernel has its own warp-reduced path.
__device__ __forceinline__ void epilogue_store16(const Epilogue &ep, int n, int z, int y, int x, int cb, float *v) {
    const int c0 = cb * 16;
    if (ep.stats) {
        double *st = ep.stats + ((size_t)n * ep.stats_stride + c0) * 2;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            atomicAdd(st + 2 * i, (double)v[i]);
            atomicAdd(st + 2 * i + 1, (double)v[i] * (double)v[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = activate(v[i] + __ldg(ep.bias + c0 + i), ep.act, ep.slope);
    if (ep.mode == OUT_PADDED_BF16) {
        const int ngroups = (ep.cout - c0) >= 16 ? 2 : ((ep.cout - c0 + 7) >> 3);
        if (ngroups <= 0) return;
        store_padded_groups(ep.dst, n, 2 * cb, ngroups, z, y, x, pack_x8(v, ep.dt), pack_x8(v + 8, ep.dt));
    } else {
        const size_t plane = (size_t)ep.dst.D * ep.dst.H * ep.dst.W;
        const int targets = ep.n_peers > 0 ? ep.n_peers : 1;
        for (int pr = 0; pr < targets; ++pr) {
            float *base = ep.n_peers > 0 ? ep.out_peers[pr] : ep.out_f32;
            const int ns = ep.n_peers > 0 ? ep.sample_offset + n : n;
            float *o = base + (size_t)ns * ep.out_nstride + (size_t)c0 * plane + ((size_t)z * ep.dst.H + y) * ep.dst.W + x;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (c0 + i < ep.cout) o[(size_t)i * plane] = v[i];
        }
    }
}

}   // namespace anx
