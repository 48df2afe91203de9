// Shared epilogue: bias + activation + store of one voxel's 16 output channels,
// either into the next layer's reflect-padded planar bf16 buffer (writing the
// mirrored shell copies as well) or into the fp32 NCDHW network output.
#pragma once
#include "layout.cuh"

namespace anx {

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ uint4 pack_bf16x8(const float *v) {
    return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                      pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void unpack_bf16x8(const uint4 &q, float *v) {
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}

// Padded-coordinate targets of interior coordinate v on an axis of size S:
// always v+1; additionally the shell cell 0 when v == 1 and S+1 when v == S-2
// (reflect: shell[0] = x[1], shell[S+1] = x[S-2]).
__device__ __forceinline__ int mirror_targets(int v, int S, int *t) {
    int n = 0;
    t[n++] = v + 1;
    if (v == 1) t[n++] = 0;
    if (v == S - 2) t[n++] = S + 1;
    return n;
}

// Stores `ngroups` (1 or 2) packed 8-channel groups of voxel (n,z,y,x) starting at
// group g0 into a padded planar buffer, including its reflect-shell copies.
__device__ __forceinline__ void store_padded_groups(const ActView &dst, int n, int g0, int ngroups, int z, int y,
                                                    int x, const uint4 &q0, const uint4 &q1) {
    int zt[3], yt[3], xt[3];
    const int nz = mirror_targets(z, dst.D, zt), ny = mirror_targets(y, dst.H, yt), nx = mirror_targets(x, dst.W, xt);
    if (nz + ny + nx == 3) {   // interior voxel: the common case
        uint4 *p = dst.at(n, g0, zt[0], yt[0], xt[0]);
        *p = q0;
        if (ngroups > 1) *dst.at(n, g0 + 1, zt[0], yt[0], xt[0]) = q1;
        return;
    }
    for (int a = 0; a < nz; ++a)
        for (int b = 0; b < ny; ++b)
            for (int c = 0; c < nx; ++c) {
                *dst.at(n, g0, zt[a], yt[b], xt[c]) = q0;
                if (ngroups > 1) *dst.at(n, g0 + 1, zt[a], yt[b], xt[c]) = q1;
            }
}

__device__ __forceinline__ float activate(float v, int act, float slope) {
    if (act == 1) return fmaxf(v, 0.0f);
    if (act == 2) return v > 0.0f ? v : v * slope;
    return v;
}

// v[16]: raw accumulators of channels [16*cb, 16*cb+16) of voxel (n,z,y,x).
__device__ __forceinline__ void epilogue_store16(const Epilogue &ep, int n, int z, int y, int x, int cb, float *v) {
    const int c0 = cb * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = activate(v[i] + __ldg(ep.bias + c0 + i), ep.act, ep.slope);
    if (ep.mode == OUT_PADDED_BF16) {
        const int ngroups = (ep.cout - c0) >= 16 ? 2 : ((ep.cout - c0 + 7) >> 3);
        if (ngroups <= 0) return;
        store_padded_groups(ep.dst, n, 2 * cb, ngroups, z, y, x, pack_bf16x8(v), pack_bf16x8(v + 8));
    } else {
        const size_t plane = (size_t)ep.dst.D * ep.dst.H * ep.dst.W;
        float *o = ep.out_f32 + ((size_t)n * ep.cout + c0) * plane + ((size_t)z * ep.dst.H + y) * ep.dst.W + x;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (c0 + i < ep.cout) o[(size_t)i * plane] = v[i];
    }
}

}   // namespace anx
