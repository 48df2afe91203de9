// Store-pattern microbenchmark: how fast can the epilogue's store patterns drain to HBM on their own?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bw store_bw.cu && ./store_bw
// Patterns (all write 8 x 16ch x 128^3 values; 537 MB as 16-bit, 1074 MB as fp32):
//   0 contiguous uint4 per lane (ideal)                      1 padded planar, x origin at +1 voxel (today)
//   2 padded planar, rows 128-byte aligned (x origin +8)     3 fp32 NCDHW, 16 scalar stores per lane (today)
//   4 fp32 NCDHW, float4 per lane after a 4x4 lane transpose (values irrelevant here)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int N = 8, D = 128, H = 128, W = 128;

__global__ void k_contig(uint4 *dst, size_t n16) {
    const uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}

// one "plane tile" = 8(x) x 16(y) voxels, 128 threads of a 256-thread block handle one tile (2 tiles per block iter)
template <int XOFF, int PITCH_EXTRA>
__global__ void k_padded(uint4 *dst) {
    const int Wp = W + 2 + PITCH_EXTRA, Hp = H + 2, Dp = D + 2;
    const size_t plane = (size_t)Hp * Wp, gstride = plane * Dp;
    const int tiles_x = W / 8, tiles_y = H / 16;
    const size_t total = (size_t)N * D * tiles_y * tiles_x;   // plane tiles
    const int sub = threadIdx.x >> 7, r = threadIdx.x & 127, ly = r >> 3, lx = r & 7;
    const uint4 v = make_uint4(r, 1, 2, 3);
    for (size_t t = (size_t)blockIdx.x * 2 + sub; t < total; t += (size_t)gridDim.x * 2) {
        size_t q = t;
        const int tx = q % tiles_x; q /= tiles_x;
        const int ty = q % tiles_y; q /= tiles_y;
        const int z = q % D; const int n = q / D;
        uint4 *p = dst + ((size_t)n * 2) * gstride + (size_t)(z + 1) * plane + (size_t)(ty * 16 + ly + 1) * Wp + tx * 8 + lx + XOFF;
        p[0] = v;
        p[gstride] = v;
    }
}

template <int VEC>
__global__ void k_f32(float *dst) {
    const size_t vol = (size_t)D * H * W;
    const int tiles_x = W / 8, tiles_y = H / 16;
    const size_t total = (size_t)N * D * tiles_y * tiles_x;
    const int sub = threadIdx.x >> 7, r = threadIdx.x & 127, ly = r >> 3, lx = r & 7;
    for (size_t t = (size_t)blockIdx.x * 2 + sub; t < total; t += (size_t)gridDim.x * 2) {
        size_t q = t;
        const int tx = q % tiles_x; q /= tiles_x;
        const int ty = q % tiles_y; q /= tiles_y;
        const int z = q % D; const int n = q / D;
        float *o = dst + (size_t)n * 16 * vol + ((size_t)z * H + ty * 16 + ly) * W + tx * 8;
        if (VEC == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) o[(size_t)i * vol + lx] = (float)i;
        } else {
            // lane (ly, lx): x-half = lx >> 2, channel sub-index j = lx & 3; stores channels 4g + j for g = 0..3
#pragma unroll
            for (int g = 0; g < 4; ++g)
                *reinterpret_cast<float4 *>(o + (size_t)(4 * g + (lx & 3)) * vol + (lx >> 2) * 4) = make_float4(1, 2, 3, 4);
        }
    }
}

template <typename F>
float time_it(F f, int reps = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    const size_t bytes16 = (size_t)N * 2 * (D + 2) * (H + 2) * (W + 16) * 16;
    const size_t bytes32 = (size_t)N * 16 * D * H * W * 4;
    void *buf; cudaMalloc(&buf, bytes32 > bytes16 ? bytes32 : bytes16);
    const double mb16 = (double)N * 16 * D * H * W * 2 / 1e6, mb32 = (double)bytes32 / 1e6;
    for (int blocks : {148, 296, 592, 1184}) {
        float t0 = time_it([&] { k_contig<<<blocks, 256>>>((uint4 *)buf, (size_t)N * 2 * D * H * W); });
        float t1 = time_it([&] { k_padded<1, 0><<<blocks, 256>>>((uint4 *)buf); });
        float t2 = time_it([&] { k_padded<8, 14><<<blocks, 256>>>((uint4 *)buf); });
        float t5 = time_it([&] { k_padded<2, 2><<<blocks, 256>>>((uint4 *)buf); });
        float t6 = time_it([&] { k_padded<4, 6><<<blocks, 256>>>((uint4 *)buf); });
        float t3 = time_it([&] { k_f32<1><<<blocks, 256>>>((float *)buf); });
        float t4 = time_it([&] { k_f32<4><<<blocks, 256>>>((float *)buf); });
        printf("blocks %4d  contiguous %.0f GB/s | padded(+1) %.0f | padded aligned %.0f | padded(+2: sector aligned) %.0f | padded(+4: 64B) %.0f | f32 scalar %.0f | f32 float4 %.0f\n", blocks,
               mb16 / t0, mb16 / t1, mb16 / t2, mb16 / t5, mb16 / t6, mb32 / t3, mb32 / t4);
    }
    float tm = time_it([&] { cudaMemsetAsync(buf, 0, bytes32); });
    printf("cudaMemset %.0f GB/s   (%s)\n", mb32 / tm, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
