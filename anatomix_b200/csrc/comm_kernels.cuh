// Kernels of the multi-GPU paths; peers are reached through NVLink-mapped device
// pointers (torch symmetric memory / CUDA IPC / cuMem handles -- the caller's business), never through a library
// collective:
//   * halo_exchange_kernel  -- depth-slab partition of one oversized volume: pushes the two boundary planes
//                              of a freshly produced tensor into the neighbours' shell planes, publishes a
//                              sequence number to them and waits for theirs.  ONE launch per exchange; the
//                              next conv of the same stream then finds both shell planes in place.
//   * widen_cl16_kernel     -- 16-bit channels-last features [N, D, H, W, C] (the compact payload of the feature
//                              all-gather / of the host download) -> fp32 NCDHW, the reference's output layout.
#pragma once
#include "epilogue.cuh"

namespace anx {

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct HaloArgs {
    const uint4 *src;        // this rank's tensor: voxel (n = 0, first group, padded z = 0, y = 0, row start)
    uint4 *lower, *upper;    // the SAME tensor inside the lower / upper neighbour's workspace (nullptr: global face)
    size_t plane;            // uint4 per padded plane: (H + 2) * pitch
    size_t group_stride;     // uint4 between consecutive channel groups: (D + 2) * plane
    size_t sample_stride;    // uint4 between samples: groups_total * group_stride
    int groups, N, D;        // groups of this tensor, samples, interior depth of the slab (equal on all ranks)
    uint32_t *my_flags;      // [0]: written by the lower neighbour, [1]: by the upper neighbour
    uint32_t *lower_flag;    // = lower neighbour's my_flags + 1   (this rank is ITS upper neighbour)
    uint32_t *upper_flag;    // = upper neighbour's my_flags + 0
    uint32_t seq;            // sequence number of this exchange (monotonic across forwards, same on all ranks)
    unsigned int *ticket;    // device counter, zero between launches
};

// Wrap-safe "flag has reached seq".  A watchdog turns a protocol bug (ranks out of step) into a trap.
__device__ __forceinline__ void wait_flag(const uint32_t *flag, uint32_t seq) {
    long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(flag) - seq) < 0) {
        if (clock64() - t0 > 8000000000LL) {   // ~4 s
            printf("anx: halo flag watchdog: waiting for %u, flag holds %u\n", seq, ld_acquire_sys(flag));
            __trap();
        }
    }
}

__global__ void __launch_bounds__(256)
halo_exchange_kernel(const HaloArgs a) {
    // my first interior plane (padded z = 1) -> lower neighbour's upper shell plane (padded z = D + 1);
    // my last interior plane (padded z = D)  -> upper neighbour's lower shell plane (padded z = 0)
    const size_t per_dir = (size_t)a.N * a.groups * a.plane;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int dir = 0; dir < 2; ++dir) {
        uint4 *dst = dir == 0 ? a.lower : a.upper;
        if (!dst) continue;
        const size_t zs = (size_t)(dir == 0 ? 1 : a.D) * a.plane, zd = (size_t)(dir == 0 ? a.D + 1 : 0) * a.plane;
        for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < per_dir; i0 += 4 * stride) {
            uint4 v[4];
            size_t off[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {        // four independent 16-byte loads in flight per thread
                const size_t i = i0 + k * stride;
                const size_t chunk = i / a.plane, within = i - chunk * a.plane;
                const size_t n = chunk / a.groups, g = chunk - n * a.groups;
                off[k] = n * a.sample_stride + g * a.group_stride + within;
                if (i < per_dir) v[k] = __ldg(a.src + off[k] + zs);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (i0 + k * stride < per_dir) dst[off[k] + zd] = v[k];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x != 0) return;
    if (atomicAdd(a.ticket, 1u) != gridDim.x - 1) return;
    // last CTA: every CTA's peer stores are fenced; publish, then wait for the neighbours' planes
    *a.ticket = 0;
    __threadfence_system();
    if (a.lower) st_release_sys(a.lower_flag, a.seq);
    if (a.upper) st_release_sys(a.upper_flag, a.seq);
    if (a.lower) wait_flag(a.my_flags + 0, a.seq);
    if (a.upper) wait_flag(a.my_flags + 1, a.seq);
}

// 16-bit channels-last [N, D, H, W, C] (C a multiple of 8) -> fp32 NCDHW.  One thread per voxel, lanes along x: a warp
// reads 32 * 2C contiguous bytes and writes C runs of 128 bytes.
__global__ void __launch_bounds__(256)
widen_cl16_kernel(const uint4 *__restrict__ src, float *__restrict__ dst, size_t n_samples, size_t vol, int C, int dt) {
    const size_t total = n_samples * vol;
    const int groups = C >> 3;
    for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
        const size_t n = id / vol, v = id - n * vol;
        float *o = dst + n * (size_t)C * vol + v;
        for (int g = 0; g < groups; ++g) {
            float f[8];
            unpack_x8(__ldg(src + id * groups + g), f, dt);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[(size_t)(g * 8 + i) * vol] = f[i];
        }
    }
}

}   // namespace anx
