// anx_selftest: probes of the tcgen05 / TMA primitives the conv kernel relies on,
// each against host arithmetic.  Exists because the kernels are developed without
// a local GPU: when a parity test fails on the box, the report says whether the
// shared-memory descriptor layout, the TMEM column addressing or the TMA brick
// geometry is to blame.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/anatomix_b200.h"
#include "layout.cuh"
#include "ptx.cuh"

using namespace anx;

namespace {

inline uint16_t bf16_bits(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
inline float bf16_val(uint16_t b) {
    uint32_t u = (uint32_t)b << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

struct GemmProbe {
    uint32_t a_off, a_lbo, a_sbo;      // byte offsets inside the smem image
    uint32_t b_off, b_lbo, b_sbo;
    uint32_t n;                        // MMA N
    uint32_t d_col;                    // TMEM column offset of D
    uint32_t repeats;                  // issue the same MMA this many times (accumulation check)
    uint32_t image_bytes;
};

// One CTA, 128 threads.  Copies a prepared shared-memory image, zeroes 64 TMEM
// columns, issues `repeats` MMAs (M=128, N=n, K=16) and dumps columns [0, 64).
__global__ void __launch_bounds__(128) gemm_probe_kernel(const uint8_t *image, GemmProbe p, float *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (uint32_t i = threadIdx.x; i < p.image_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(smem)[i] = reinterpret_cast<const uint4 *>(image)[i];
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < 32) {
        tmem_alloc<64>(&slot);
        tmem_relinquish();
    }
    fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t lane_base = tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
    for (int c = 0; c < 64; c += 16) tmem_st16_zero(lane_base + c);
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t s0 = smem_u32(smem);
        const uint64_t ad = smem_desc_kmajor_noswz(s0 + p.a_off, p.a_lbo, p.a_sbo);
        const uint64_t bd = smem_desc_kmajor_noswz(s0 + p.b_off, p.b_lbo, p.b_sbo);
        for (uint32_t r = 0; r < p.repeats; ++r) umma_bf16(tmem + p.d_col, ad, bd, idesc_bf16_m128(p.n), 1);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0, 100);
    tc_fence_after();
    for (int c = 0; c < 64; c += 16) {
        float v[16];
        tmem_ld16(lane_base + c, v);
        for (int i = 0; i < 16; ++i) out[threadIdx.x * 64 + c + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 64);
}

// TMA probe: loads one halo brick with the conv kernel's box geometry and copies
// the shared-memory image back out.
__global__ void __launch_bounds__(128)
tma_probe_kernel(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int c2, int c3, uint32_t bytes,
                 uint8_t *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, bytes);
        tma_load_4d(smem, &tmap, &bar, c0, c1, c2, c3);
    }
    mbar_wait(&bar, 0, 101);
    for (uint32_t i = threadIdx.x; i < bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(out)[i] = reinterpret_cast<const uint4 *>(smem)[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Report {
    std::string text;
    bool ok = true;
    void line(const char *fmt, ...) {
        char tmp[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(tmp, sizeof tmp, fmt, ap);
        va_end(ap);
        text += tmp;
        text += "\n";
    }
};

#define ST_CUDA(call)                                                                    \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            rep.line("CUDA error %s at %s:%d", cudaGetErrorString(e__), __FILE__, __LINE__); \
            rep.ok = false;                                                              \
            return;                                                                      \
        }                                                                                \
    } while (0)

// A: 128 x 16, 8-row groups `a_sbo` apart, K halves `a_lbo` apart; B: rows x 16.
void run_gemm_probe(Report &rep, const char *name, uint32_t a_sbo, uint32_t a_lbo, uint32_t a_shift, uint32_t b_rows,
                    uint32_t b_row0, uint32_t n, uint32_t d_col, uint32_t repeats) {
    std::vector<float> A(128 * 16), B(b_rows * 16);
    uint32_t seed = 12345u + a_sbo * 7 + n;
    auto rnd = [&]() {
        seed = seed * 1664525u + 1013904223u;
        return ((seed >> 9) & 0xffff) / 32768.0f - 1.0f;
    };
    for (auto &v : A) v = bf16_val(bf16_bits(rnd()));
    for (auto &v : B) v = bf16_val(bf16_bits(rnd()));
    const uint32_t a_bytes = std::max(a_lbo + 16 * a_sbo + 256, 2 * a_lbo) + a_shift + 256;
    const uint32_t b_off = (a_bytes + 1023) / 1024 * 1024;
    const uint32_t b_lbo = 16 * b_rows, b_sbo = 128;
    const uint32_t total = (b_off + 2 * b_lbo + 1023) / 1024 * 1024;
    std::vector<uint8_t> img(total, 0);
    auto put = [&](uint32_t off, float v) {
        uint16_t b = bf16_bits(v);
        std::memcpy(&img[off], &b, 2);
    };
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 16; ++k)
            put(a_shift + (k / 8) * a_lbo + (r / 8) * a_sbo + (r % 8) * 16 + (k % 8) * 2, A[r * 16 + k]);
    for (uint32_t r = 0; r < b_rows; ++r)
        for (int k = 0; k < 16; ++k)
            put(b_off + (k / 8) * b_lbo + (r / 8) * b_sbo + (r % 8) * 16 + (k % 8) * 2, B[r * 16 + k]);
    uint8_t *d_img = nullptr;
    float *d_out = nullptr;
    ST_CUDA(cudaMalloc(&d_img, total));
    ST_CUDA(cudaMalloc(&d_out, 128 * 64 * sizeof(float)));
    ST_CUDA(cudaMemcpy(d_img, img.data(), total, cudaMemcpyHostToDevice));
    GemmProbe p{a_shift, a_lbo, a_sbo, b_off + b_row0 * 16, b_lbo, b_sbo, n, d_col, repeats, total};
    ST_CUDA(cudaFuncSetAttribute(gemm_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    gemm_probe_kernel<<<1, 128, total>>>(d_img, p, d_out);
    ST_CUDA(cudaGetLastError());
    ST_CUDA(cudaDeviceSynchronize());
    std::vector<float> out(128 * 64);
    ST_CUDA(cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(d_img);
    cudaFree(d_out);
    double worst = 0, worst_outside = 0;
    for (int r = 0; r < 128; ++r)
        for (uint32_t c = 0; c < 64; ++c) {
            double ref = 0;
            if (c >= d_col && c < d_col + n) {
                for (int k = 0; k < 16; ++k) ref += (double)A[r * 16 + k] * B[(b_row0 + c - d_col) * 16 + k];
                ref *= repeats;
                worst = std::max(worst, std::fabs(ref - out[r * 64 + c]));
            } else {
                worst_outside = std::max(worst_outside, (double)std::fabs(out[r * 64 + c]));
            }
        }
    const bool pass = worst < 1e-3 && worst_outside == 0.0;
    rep.line("[%s] gemm probe %-28s sbo=%u lbo=%u shift=%u N=%u dcol=%u rep=%u: max|err|=%.3g outside=%.3g",
             pass ? "ok" : "FAIL", name, a_sbo, a_lbo, a_shift, n, d_col, repeats, worst, worst_outside);
    if (!pass) {
        rep.ok = false;
        rep.line("   row0: got %.4f %.4f %.4f %.4f ...; row1: %.4f %.4f; row8: %.4f %.4f", out[d_col], out[d_col + 1],
                 out[d_col + 2], out[d_col + 3], out[64 + d_col], out[64 + d_col + 1], out[8 * 64 + d_col],
                 out[8 * 64 + d_col + 1]);
    }
}

void run_tma_probe(Report &rep) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
        rep.line("[FAIL] cuTensorMapEncodeTiled entry point missing");
        rep.ok = false;
        return;
    }
    EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(fn);
    // padded planar tensor: n*groups = 4 planes, D+2=6, H+2=20, W+2=18 voxels of 8 channels
    const int G = 4, Dp = 6, Hp = 20, Wp = 18, bz = 2;
    const size_t vox = (size_t)G * Dp * Hp * Wp;
    std::vector<uint16_t> host(vox * 8);
    for (size_t i = 0; i < host.size(); ++i) host[i] = (uint16_t)(i * 2654435761u >> 16);
    uint16_t *d_src = nullptr;
    uint8_t *d_out = nullptr;
    const uint32_t bytes = 2u * (bz + 2) * HALO_Y * ROW_BYTES;
    ST_CUDA(cudaMalloc(&d_src, host.size() * 2));
    ST_CUDA(cudaMalloc(&d_out, bytes));
    ST_CUDA(cudaMemcpy(d_src, host.data(), host.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap tmap;
    cuuint64_t dims[4] = {(cuuint64_t)Wp * 8, Hp, Dp, G};
    cuuint64_t strides[3] = {(cuuint64_t)Wp * 16, (cuuint64_t)Wp * Hp * 16, (cuuint64_t)Wp * Hp * Dp * 16};
    cuuint32_t box[4] = {HALO_X * 8, HALO_Y, bz + 2, 2};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d_src, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rep.line("[FAIL] cuTensorMapEncodeTiled returned %d", (int)r);
        rep.ok = false;
        return;
    }
    // brick at padded origin (x=8, y=16, z=2), groups 2..3: y rows 16..33 overhang Hp=20 -> zero filled
    const int x0 = 8, y0 = 16, z0 = 2, g0 = 2;
    ST_CUDA(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    tma_probe_kernel<<<1, 128, bytes>>>(tmap, x0 * 8, y0, z0, g0, bytes, d_out);
    ST_CUDA(cudaGetLastError());
    ST_CUDA(cudaDeviceSynchronize());
    std::vector<uint16_t> got(bytes / 2);
    ST_CUDA(cudaMemcpy(got.data(), d_out, bytes, cudaMemcpyDeviceToHost));
    cudaFree(d_src);
    cudaFree(d_out);
    size_t bad = 0, first_bad = 0;
    for (int g = 0; g < 2; ++g)
        for (int z = 0; z < bz + 2; ++z)
            for (int y = 0; y < HALO_Y; ++y)
                for (int x = 0; x < HALO_X; ++x)
                    for (int c = 0; c < 8; ++c) {
                        const size_t si = ((((size_t)g * (bz + 2) + z) * HALO_Y + y) * HALO_X + x) * 8 + c;
                        const int gz = z0 + z, gy = y0 + y, gx = x0 + x;
                        uint16_t want = 0;
                        if (gz < Dp && gy < Hp && gx < Wp)
                            want = host[((((size_t)(g0 + g) * Dp + gz) * Hp + gy) * Wp + gx) * 8 + c];
                        if (got[si] != want) {
                            if (!bad) first_bad = si;
                            ++bad;
                        }
                    }
    rep.line("[%s] tma brick probe: %zu mismatching elements of %u (first at %zu)", bad ? "FAIL" : "ok", bad,
             bytes / 2, first_bad);
    if (bad) rep.ok = false;
}

}   // namespace

extern "C" anx_status anx_selftest(int32_t device, char *report, size_t report_bytes) {
    Report rep;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return ANX_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    cudaSetDevice(device);
    rep.line("device %d: %s sm_%d%d, %d SMs, %zu KB smem/block", device, prop.name, prop.major, prop.minor,
             prop.multiProcessorCount, prop.sharedMemPerBlockOptin / 1024);
    if (prop.major != 10) {
        rep.line("[FAIL] not an sm_100 device");
        rep.ok = false;
    } else {
        // dense canonical layout, as a plain GEMM tile would use it
        run_gemm_probe(rep, "dense", 128, 2048, 0, 32, 0, 32, 0, 1);
        // the conv kernel's A geometry: y rows 160 B apart, channel planes far apart, tap shift of 16 B
        run_gemm_probe(rep, "halo-brick strides", 160, 28800, 0, 48, 0, 48, 0, 1);
        run_gemm_probe(rep, "halo-brick shifted +16B", 160, 28800, 16, 48, 0, 48, 0, 1);
        run_gemm_probe(rep, "halo-brick shifted +176B", 160, 28800, 176, 48, 0, 48, 0, 1);
        // B sub-range (rows 16..47) written at TMEM column 16, accumulated twice
        run_gemm_probe(rep, "B rows 16.., D col 16, x2", 160, 28800, 16, 48, 16, 32, 16, 2);
        run_gemm_probe(rep, "N=16 at D col 48", 160, 11520, 32, 48, 32, 16, 48, 3);
        run_tma_probe(rep);
    }
    if (report && report_bytes) {
        std::strncpy(report, rep.text.c_str(), report_bytes - 1);
        report[report_bytes - 1] = 0;
    }
    return rep.ok ? ANX_OK : ANX_ERR_CUDA;
}
