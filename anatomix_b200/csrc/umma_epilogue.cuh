// Epilogue of one CTA tile, shared by the generic conv kernel and the stem kernel:
// TMEM -> registers -> (instance-norm statistics) -> activation -> 16-bit pack ->
// stores into the next layer's reflect-padded planar buffer (shell copies included),
// or fp32 NCDHW / fused all-gather stores for the network output; and the re-seeding
// of the drained accumulator columns with the per-channel shift.
//
// One call handles the planes [plane_lo, plane_hi) of this warp's TMEM lane quadrant.
#pragma once
#include "epilogue.cuh"
#include "ptx.cuh"

namespace anx {

struct EpiTile {
    int n, z0, y, x;        // sample, first output plane, this thread's voxel row / column
    int chan0;              // first output channel of this CTA's channel split
    bool in_xy;             // voxel inside the volume in y and x
    bool store;             // false only in timing experiments
};

// Planes of a tile are split between the `parts` warps of a TMEM lane quadrant in contiguous runs:
// warp `part` owns [plane_lo, plane_hi).  `pairs`: runs of even length (the fused pooling reduces z pairs).
struct PlaneSplit { int part, parts, pairs; };
__device__ __forceinline__ int plane_run(int bz, const PlaneSplit &ps) {
    const int per = (bz + ps.parts - 1) / ps.parts;
    return ps.pairs ? (per + 1) & ~1 : per;
}
__device__ __forceinline__ int plane_lo(const PlaneSplit &ps, int bz) { return min(bz, ps.part * plane_run(bz, ps)); }
__device__ __forceinline__ int plane_hi(const PlaneSplit &ps, int bz) { return min(bz, plane_lo(ps, bz) + plane_run(bz, ps)); }

// Seeded accumulators (ncols = 16: two 8-channel groups per plane): the partial sums of the tile that
// will reuse an accumulator stage are pulled towards L1 while the warp waits for the MMAs, and loaded
// for real next to the tcgen05.ld of the plane they replace -- no registers are held across the wait.
__device__ __forceinline__ const uint4 *seed_ptr(const Epilogue &ep, const EpiTile &t, int b, int group) {
    return ep.seed_src.at(t.n, group, t.z0 + b + 1, t.y + 1, t.x + 1);
}
__device__ __forceinline__ bool seed_valid(const EpiTile &t, bool tile_valid, int b, int D) {
    return tile_valid && t.in_xy && (t.z0 + b) < D;
}
__device__ __forceinline__ void prefetch_seeds(const Epilogue &ep, const EpiTile &t, bool tile_valid, const PlaneSplit &half, int bz,
                                               int D) {
    for (int b = plane_lo(half, bz); b < plane_hi(half, bz); ++b)
        if (seed_valid(t, tile_valid, b, D)) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(seed_ptr(ep, t, b, 0)));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(seed_ptr(ep, t, b, 1)));
        }
}
__device__ __forceinline__ void load_seed16(const Epilogue &ep, const EpiTile &t, bool tile_valid, int b, int D,
                                            uint4 &q0, uint4 &q1) {
    q0 = q1 = make_uint4(0u, 0u, 0u, 0u);
    if (seed_valid(t, tile_valid, b, D)) {
        q0 = __ldg(seed_ptr(ep, t, b, 0));
        q1 = __ldg(seed_ptr(ep, t, b, 1));
    }
}

// Compile-time epilogue variants: each launch carries only the code of its own path.
enum EpiMode : int {
    EPI_PADDED = 0,      // 16-bit padded planar store (+ shell)
    EPI_POOL = 1,        // ... and the fused 2x2x2 pooled tensor
    EPI_D2S = 2,         // depth-to-space partial sums (low-resolution half of a decoder conv)
    EPI_STATS = 3,       // ... raw output + instance-norm sums
    EPI_F32 = 4,         // fp32 NCDHW network output
    EPI_F32_PEERS = 5,   // ... into every peer's gather buffer
    EPI_SEEDED = 6,      // padded store, accumulators re-seeded from stored partial sums
    EPI_F32_HEAD = 7,    // fp32 NCDHW output of a linear head applied to the conv's 16 channels
    EPI_CL16 = 8         // 16-bit channels-last [N, D, H, W, C] network output (into every peer's buffer when gathering)
};
constexpr int HEAD_SMEM_OFFSET = 32;   // floats behind the channel shift in shared memory where the head lives

// `next` / `next_valid`: the tile that will reuse this accumulator stage (seeded kernels only).
// `stats_acc`: EPI_STATS only -- this warp's shared-memory accumulation rows (ncols / 16 x 32 doubles), or nullptr to
// add to the global sums directly.
template <int MODE>
__device__ __forceinline__ void umma_epilogue_tile(const Epilogue &ep, const EpiTile &t, uint32_t acc,
                                                   const float *seed, const PlaneSplit &half, int bz, int ncols, int D,
                                                   const EpiTile &next, bool next_valid, double *stats_acc = nullptr) {
    const int Dd = ep.dst.D, Hh = ep.dst.H, Ww = ep.dst.W;
    const size_t plane = (size_t)(Hh + 2) * ep.dst.pitch;      // uint4 units
    const size_t gstride = (size_t)(Dd + 2) * plane;
    const size_t vol = (size_t)Dd * Hh * Ww;
    constexpr bool SEEDED = MODE == EPI_SEEDED;
    constexpr bool PADDED = MODE != EPI_F32 && MODE != EPI_F32_PEERS && MODE != EPI_F32_HEAD && MODE != EPI_CL16;
    const bool big = (Dd >= 4) & (Hh >= 4) & (Ww >= 4);         // else: generic mirror loops
    const int rep = ep.dst.shell_rep;
    const int mdx = mirror_delta(t.x, Ww, rep), mdy = mirror_delta(t.y, Hh, rep);
    const size_t rowp = (size_t)ep.dst.pitch;
    uint4 *pbase = nullptr;
    float *fbase = nullptr;
    if constexpr (PADDED)
        pbase = ep.dst.at(t.n, t.chan0 >> 3, t.z0 + 1, t.y + 1, t.x + 1);
    else if constexpr (MODE != EPI_CL16)
        fbase = ep.out_f32 + (size_t)t.n * ep.out_nstride + (size_t)t.chan0 * vol + ((size_t)t.z0 * Hh + t.y) * Ww + t.x;
    const int chunks = ncols >> 4;
    for (int cb = 0; cb < chunks; ++cb) {
        const int c0 = t.chan0 + cb * 16;
        float sd[16];   // this chunk's seed values (16-byte aligned in shared memory), reused for every plane
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 f = reinterpret_cast<const float4 *>(seed + cb * 16)[i];
            sd[4 * i] = f.x; sd[4 * i + 1] = f.y; sd[4 * i + 2] = f.z; sd[4 * i + 3] = f.w;
        }
        float s16[16], q16[16];
        if constexpr (MODE == EPI_STATS) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { s16[i] = 0.0f; q16[i] = 0.0f; }
        }
        const int b_end = plane_hi(half, bz);
        if constexpr (MODE == EPI_POOL) {
            // ---- fused 2x2x2 pooling: planes in pairs (b, b+1); requires an even, pair-aligned plane range
            const int ngroups = (ep.cout - c0) >= 16 ? 2 : ((ep.cout - c0 + 7) >> 3);
            for (int b = plane_lo(half, bz); b < b_end; b += 2) {
                uint32_t r0[16], r1[16];
                __syncwarp();
                tmem_ld16_nowait(acc + b * ncols + cb * 16, r0);
                tmem_ld16_nowait(acc + (b + 1) * ncols + cb * 16, r1);
                tmem_wait_ld();
                tmem_ld_ready16(r0);
                tmem_ld_ready16(r1);
                tmem_st16(acc + b * ncols + cb * 16, sd);
                tmem_st16(acc + (b + 1) * ncols + cb * 16, sd);
                const int z = t.z0 + b;
                const bool ok = t.in_xy && z < D && t.store;     // the pair is inside or outside together (even sizes)
                float v0[16], v1[16], m[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    v0[i] = activate(__uint_as_float(r0[i]), ep.act, ep.slope);
                    v1[i] = activate(__uint_as_float(r1[i]), ep.act, ep.slope);
                }
                const uint4 a0 = pack_x8(v0, ep.dt), a1 = pack_x8(v0 + 8, ep.dt);
                const uint4 c0q = pack_x8(v1, ep.dt), c1q = pack_x8(v1 + 8, ep.dt);
                if (ok && ngroups > 0) {
                    if (big) {
                        uint4 *p = pbase + (size_t)b * plane + (size_t)(2 * cb) * gstride;
                        p[0] = a0;
                        p[plane] = c0q;
                        if (ngroups > 1) { p[gstride] = a1; p[gstride + plane] = c1q; }
                        const int mz0 = mirror_delta_z(z, Dd, rep, ep.dst.z_open), mz1 = mirror_delta_z(z + 1, Dd, rep, ep.dst.z_open);
                        if (mdx | mdy | mz0) {
                            store_mirrors(p, a0, mz0, mdy, mdx, rowp, plane);
                            if (ngroups > 1) store_mirrors(p + gstride, a1, mz0, mdy, mdx, rowp, plane);
                        }
                        if (mdx | mdy | mz1) {
                            store_mirrors(p + plane, c0q, mz1, mdy, mdx, rowp, plane);
                            if (ngroups > 1) store_mirrors(p + gstride + plane, c1q, mz1, mdy, mdx, rowp, plane);
                        }
                    } else {
                        store_padded_groups(ep.dst, t.n, c0 >> 3, ngroups, z, t.y, t.x, a0, a1);
                        store_padded_groups(ep.dst, t.n, c0 >> 3, ngroups, z + 1, t.y, t.x, c0q, c1q);
                    }
                }
                if (ep.pool_kind == 0) {
                    // max pooling on the packed 16-bit pairs: 16 shuffles + 24 two-wide max, no unpack / repack
                    uint32_t pk[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    const uint32_t pc[8] = {c0q.x, c0q.y, c0q.z, c0q.w, c1q.x, c1q.y, c1q.z, c1q.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        pk[i] = max16x2(pk[i], pc[i], ep.dt);
                        pk[i] = max16x2(pk[i], __shfl_xor_sync(0xffffffffu, pk[i], 1), ep.dt);
                        pk[i] = max16x2(pk[i], __shfl_xor_sync(0xffffffffu, pk[i], 8), ep.dt);
                    }
                    if (ok && ngroups > 0 && !((t.x | t.y) & 1))
                        store_padded_groups(ep.pool_dst, t.n, c0 >> 3, ngroups, z >> 1, t.y >> 1, t.x >> 1,
                                            make_uint4(pk[0], pk[1], pk[2], pk[3]), make_uint4(pk[4], pk[5], pk[6], pk[7]));
                    continue;
                }
                // mean pooling of the STORED (rounded) values, as the reference pools the stored activation
                unpack_x8(a0, v0, ep.dt); unpack_x8(a1, v0 + 8, ep.dt);
                unpack_x8(c0q, v1, ep.dt); unpack_x8(c1q, v1 + 8, ep.dt);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float a = ep.pool_kind == 0 ? fmaxf(v0[i], v1[i]) : v0[i] + v1[i];
                    const float bx = __shfl_xor_sync(0xffffffffu, a, 1);       // x neighbour
                    a = ep.pool_kind == 0 ? fmaxf(a, bx) : a + bx;
                    const float by = __shfl_xor_sync(0xffffffffu, a, 8);       // y neighbour (rows are 8 lanes apart)
                    a = ep.pool_kind == 0 ? fmaxf(a, by) : a + by;
                    m[i] = ep.pool_kind == 0 ? a : a * 0.125f;
                }
                if (ok && ngroups > 0 && !((t.x | t.y) & 1))
                    store_padded_groups(ep.pool_dst, t.n, c0 >> 3, ngroups, z >> 1, t.y >> 1, t.x >> 1, pack_x8(m, ep.dt),
                                        pack_x8(m + 8, ep.dt));
            }
            continue;
        }
        uint4 nq0, nq1;   // seeds of the NEXT plane, loaded one plane ahead (L2 latency hides behind a plane of work)
        const uint4 *sbase = nullptr;          // seed tensor at (next tile, plane 0, this voxel), group 0
        size_t splane = 0, sgroup = 0;
        bool s_xy = false;
        auto seed_load = [&](int b, uint4 &q0, uint4 &q1) {
            q0 = q1 = make_uint4(0u, 0u, 0u, 0u);
            if (s_xy && (next.z0 + b) < D) {
                q0 = __ldg(sbase + (size_t)b * splane);
                q1 = __ldg(sbase + (size_t)b * splane + sgroup);
            }
        };
        if (SEEDED) {
            s_xy = next_valid && next.in_xy;
            splane = (size_t)(ep.seed_src.H + 2) * ep.seed_src.pitch;
            sgroup = splane * (ep.seed_src.D + 2);
            sbase = seed_ptr(ep, next, 0, 0);
            seed_load(plane_lo(half, bz), nq0, nq1);
        }
        for (int b = plane_lo(half, bz); b < b_end; ++b) {
            uint32_t r[1][16];
            uint4 sq0, sq1;
            __syncwarp();   // tcgen05.ld / st are warp-collective
            tmem_ld16_nowait(acc + b * ncols + cb * 16, r[0]);
            if (SEEDED) {
                sq0 = nq0; sq1 = nq1;
                if (b + 1 < b_end) seed_load(b + 1, nq0, nq1);
            }
            tmem_wait_ld();
            tmem_ld_ready16(r[0]);
            if (SEEDED) {   // re-seed with the stored partial sums of the tile that reuses this stage
                float sv[16];
                unpack_x8(sq0, sv, ep.dt);
                unpack_x8(sq1, sv + 8, ep.dt);
                tmem_st16(acc + b * ncols + cb * 16, sv);
            } else {
                tmem_st16(acc + b * ncols + cb * 16, sd);   // re-seed for a later tile
            }
            {
                const int z = t.z0 + b;
                const bool ok = t.in_xy && z < D && t.store;
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[0][i]);
                if constexpr (MODE == EPI_STATS) {
                    if (ok) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) { s16[i] += v[i]; q16[i] = fmaf(v[i], v[i], q16[i]); }
                    }
                }
                if (!ok) continue;
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = activate(v[i], ep.act, ep.slope);
                if constexpr (PADDED) {
                    const int ngroups = (ep.cout - c0) >= 16 ? 2 : ((ep.cout - c0 + 7) >> 3);
                    if (ngroups <= 0) continue;
                    const uint4 q0 = pack_x8(v, ep.dt), q1 = pack_x8(v + 8, ep.dt);
                    if constexpr (MODE == EPI_D2S) {
                        // depth-to-space: chunk = 16 channels of one parity; no shell (only the interior is read back)
                        const int par = c0 / ep.d2s_cout, co0 = c0 - par * ep.d2s_cout;
                        uint4 *p = ep.dst.at(t.n, co0 >> 3, 2 * z + ((par >> 2) & 1) + 1, 2 * t.y + ((par >> 1) & 1) + 1,
                                             2 * t.x + (par & 1) + 1);
                        *p = q0;
                        p[gstride] = q1;
                    } else if (big) {
                        uint4 *p = pbase + (size_t)b * plane + (size_t)(2 * cb) * gstride;
                        *p = q0;
                        if (ngroups > 1) p[gstride] = q1;
                        const int mdz = mirror_delta_z(z, Dd, rep, ep.dst.z_open);
                        if (mdx | mdy | mdz) {   // shell copies: a few predicated stores, no loops
                            store_mirrors(p, q0, mdz, mdy, mdx, rowp, plane);
                            if (ngroups > 1) store_mirrors(p + gstride, q1, mdz, mdy, mdx, rowp, plane);
                        }
                    } else {
                        store_padded_groups(ep.dst, t.n, c0 >> 3, ngroups, z, t.y, t.x, q0, q1);
                    }
                } else if constexpr (MODE == EPI_F32) {
                    float *o = fbase + (size_t)(cb * 16) * vol + (size_t)b * Hh * Ww;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < ep.cout) o[(size_t)i * vol] = v[i];
                } else if constexpr (MODE == EPI_CL16) {
                    // 16-bit channels-last: a voxel's channels are contiguous (32 bytes per 16-channel chunk)
                    const uint4 q0 = pack_x8(v, ep.dt), q1 = pack_x8(v + 8, ep.dt);
                    const size_t cg = (size_t)(ep.cout >> 3);       // 8-channel groups per voxel
                    const size_t off = ((((size_t)(ep.sample_offset + t.n) * Dd + z) * Hh + t.y) * Ww + t.x) * cg + (c0 >> 3);
                    const int targets = ep.n_peers > 0 ? ep.n_peers : 1;
                    for (int pr = 0; pr < targets; ++pr) {
                        uint4 *o = reinterpret_cast<uint4 *>(ep.n_peers > 0 ? ep.out_peers[pr] : ep.out_f32) + off;
                        o[0] = q0;
                        if (c0 + 8 < ep.cout) o[1] = q1;
                    }
                } else if constexpr (MODE == EPI_F32_HEAD) {
                    // 1x1x1 conv on the voxel's channel vector (registers) with the head in shared memory
                    const float *hb = seed + HEAD_SMEM_OFFSET, *hw = hb + HEAD_MAX;
                    float *o = ep.out_f32 + (size_t)t.n * ep.out_nstride + ((size_t)z * Hh + t.y) * Ww + t.x;
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
                        o[(size_t)k * vol] = a;
                    }
                } else {
                    // fused all-gather: the same values go to every rank's gather buffer over NVLink
                    const size_t off = (size_t)(ep.sample_offset + t.n) * ep.out_nstride + (size_t)c0 * vol +
                                       ((size_t)z * Hh + t.y) * Ww + t.x;
                    for (int pr = 0; pr < ep.n_peers; ++pr) {
                        float *o = ep.out_peers[pr] + off;
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (c0 + i < ep.cout) o[(size_t)i * vol] = v[i];
                    }
                }
            }
        }
        if constexpr (MODE == EPI_STATS) {   // whole warp converged: the loops above have warp-uniform trip counts
            __syncwarp();
            warp_stats_add(s16, q16, ep.stats + ((size_t)t.n * ep.stats_stride + c0) * 2,
                           stats_acc ? stats_acc + cb * 32 : nullptr);
        }
    }
}

}   // namespace anx
