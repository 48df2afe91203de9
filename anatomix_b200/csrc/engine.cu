// Engine behind the C ABI in include/anatomix_b200.h: layer program of the
// anatomix U-Net (reference network.py:309-465), weight folding / packing,
// workspace planning, TMA descriptors and the launch sequence of one forward
// (reference network.py:530-548).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/anatomix_b200.h"
#include "comm_kernels.cuh"
#include "conv_rows.cuh"
#include "conv_umma.cuh"
#include "simt_kernels.cuh"
#include "stem_umma.cuh"

using namespace anx;

namespace {

// ------------------------------------------------------------------ utilities
inline uint16_t f32_to_bf16_rne(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
inline uint16_t f32_to_f16_rne(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (x > 0x7f800000u ? 0x200u : 0));   // inf / nan
    if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);                                      // overflow -> inf
    if (x < 0x33000001u) return (uint16_t)sign;                                                   // underflow -> 0
    if (x < 0x38800000u) {                                                                        // subnormal half
        const int shift = 126 - (int)(x >> 23);              // 14..24
        const uint32_t mant = (x & 0x7fffffu) | 0x800000u;
        uint32_t h = mant >> shift;
        const uint32_t rem = mant & ((1u << shift) - 1), halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (h & 1))) ++h;
        return (uint16_t)(sign | h);
    }
    uint32_t h = ((x - 0x38000000u) >> 13);
    const uint32_t rem = x & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;
    return (uint16_t)(sign | h);
}
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Environment switches for A/B timing experiments: read only by -DANX_EXPERIMENTS builds (see layout.cuh).
inline const char *exp_env(const char *name) {
#ifdef ANX_EXPERIMENTS
    return getenv(name);
#else
    (void)name;
    return nullptr;
#endif
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn load_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// ---------------------------------------------------------------- layer program
enum StepKind { STEP_STEM, STEP_CONV, STEP_POOL, STEP_UP, STEP_NORM };

struct ConvLayer {            // one nn.Conv3d of the Sequential, network order
    int module_index;         // position in the flat Sequential (state-dict key)
    int cin, cout, ncols;     // ncols = cout rounded up to 16
    int level;                // resolution level
    bool has_norm, has_act, is_stem, is_final;
    bool is_tap = false;      // un-folded clone of a conv that writes its PRE-norm output as fp32 NCDHW (feature taps)
    int src_buf, dst_buf, dst_group_offset;   // buffer ids (-1: network input / output)
    // device parameters
    bool ready = false;
    int fold = 0, groups = 3;
    int n_splits = 1, ncols_split = 0;        // Cout > 256: channel splits of 256 handled by different CTAs
    bool inorm = false;                       // followed by InstanceNorm: raw output + statistics
    int pool_dst_buf = -1;                    // >= 0: the following 2x2x2 pool may be fused into this conv's epilogue
    int d2s_cout = 0;                         // > 0: low-resolution half of a decoder conv (depth-to-space store)
    int seed_buf = -1;                        // >= 0: skip half of that conv, accumulators seeded from this buffer
    size_t stats_index = 0;                   // first double of this conv's [N][ncols][2] block, per sample-channel
    void *d_wpack = nullptr;  // 16-bit slabs (tensor-core convs) or fp32 [cin][27][cout] (CUDA-core stem)
    void *d_wstem = nullptr;  // stem on tensor cores: bf16 hi|lo images of B, [kq][half][3*ncols][8] each
    void *d_wrows = nullptr;  // 16 -> 16 row kernel: three B images [z rotation][dx][k half][144 rows][8] (conv_rows.cuh)
    int alt_splits = 0;       // > 0: a second packing in 64-column channel splits, unfolded (small problems, see make_geom)
    void *d_wpack_alt = nullptr;
    void *d_wtrim = nullptr;  // low-resolution decoder half: compact per-tap tiles of the non-zero column spans (resident B)
    size_t wtrim_bytes = 0;
    void *d_wstem_rows = nullptr;   // row-form stem (Cin = 1, 16 columns): three bf16 B images [z rotation][k half][144 rows][8]
    float *d_bias = nullptr;  // [ncols]
    size_t wpack_bytes = 0;
};

struct Buffer {               // padded planar activation buffer
    int level, groups;
    int shell_rep = 0;        // producer writes replicate instead of reflect copies into the shell
    int first = -1, last = -1;   // first step that writes it, last step that touches it (workspace liveness)
};

// One nn.Conv3d of the reference's Sequential as the binding sees it (anx_engine_set_conv ordinal).
// A decoder conv behind a nearest upsample may run as TWO launches (see build_program): conv_b >= 0.
struct Logical {
    int module_index, cin, cout;
    bool has_norm;
    int conv_a, conv_b;
};

struct Step {
    StepKind kind;
    int conv = -1;                    // STEP_STEM / STEP_CONV
    int src_buf = -1, dst_buf = -1;   // STEP_POOL / STEP_UP
    int groups = 0, dst_group_offset = 0;
    int module_index = -1;            // STEP_POOL / STEP_UP: position of the nn.MaxPool3d / nn.Upsample slot
    char name[32];
};

// A feature tap (reference network.py:475-529, `forward(layers=[...])`): the activation that exists after
// Sequential slot `module_index`, when the engine materialises it (post-activation conv outputs, pooled
// tensors, the [skip | upsampled] concat after an Upsample slot, the network output).
struct TapSite {
    int module_index;
    int buffer, group_offset, groups;   // buffer -1: the network output
    int channels, level;
    int last_step;                      // the tensor is complete once steps [0, last_step] have run
    int prenorm_conv = -1;              // >= 0: PRE-norm output of logical conv `prenorm_conv` (buffer = -2): not stored,
                                        // re-evaluated on request by an un-folded clone of that conv right after last_step
};

struct GatherArgs {           // where and how the final conv stores (one forward call)
    float *peers[8] = {nullptr};     // fused feature all-gather: every rank's gather buffer (n_peers = 0: `out` only)
    int n_peers = 0, sample_offset = 0;
    int payload = ANX_PAYLOAD_F32_NCDHW;
    size_t out_nstride = 0;          // fp32 output: floats between samples (0 = dense)
};

struct ShapePlan {            // everything that depends on (N, D, H, W, workspace)
    int N = 0, D = 0, H = 0, W = 0;
    void *workspace = nullptr;
    std::vector<size_t> buf_offset;
    size_t stats_offset = 0, stats_bytes = 0;   // instance-norm sums at the end of the workspace
    std::vector<size_t> conv_stats_offset;      // per conv, bytes from workspace start
    std::vector<CUtensorMap> tmaps;   // per conv (unused for the stem)
    std::vector<ConvGeom> geoms;      // per conv
};

}   // namespace

struct anx_engine {
    anx_unet_desc desc;
    int dt = DT_BF16;         // storage type of activations / packed weights
    int use_rows = 1;         // thin 16 -> 16 layers run on conv3_rows_kernel when the shape allows (ANX_ROWS=0: never)
    int rows_rpw1 = 0;        // row-kernel variants that run 16 one-row epilogue warps instead of 8 row-pair warps:
                              // bit 0 plain padded store, 1 stem, 2 seeded, 3 fp32 / 16-bit / head outputs
    int x_lead = 0;           // row layout of the padded planar buffers (layout.cuh layout_of; ANX_X_LEAD)
    int num_sms = 148;
    int max_smem = 0;
    std::vector<ConvLayer> convs;
    std::vector<Logical> logical;
    std::vector<Buffer> bufs;
    std::vector<Step> steps;
    std::vector<TapSite> taps;
    std::vector<ConvLayer> tap_convs;   // per logical conv: un-folded clone for pre-norm taps (ready once anx_engine_set_tap_conv ran)
    // optional linear head fused into the last conv's epilogue (anx_engine_set_head)
    // depth-slab mode (anx_engine_set_slab): which z faces of this slab have a neighbour, and the depth of the
    // whole volume (instance-norm statistics are per whole volume)
    int slab_lower = 0, slab_upper = 0, slab_depth_total = 0;
    int head_nc = 0;
    float *d_head = nullptr;          // [HEAD_MAX][16] weights then [HEAD_MAX] bias
    std::mutex mu;
    std::vector<std::shared_ptr<ShapePlan>> plans;
    // host-buffer pipeline (anx_engine_forward_host): copy streams and fork/join events
    std::mutex host_mu;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_in[8] = {nullptr}, ev_out[8] = {nullptr};
    // pipelined host-buffer forwards (anx_engine_forward_host_pipelined): last download that read a given dev_out
    // buffer (two buffer sets alternate), and whether `ev_join` marks downloads the caller has not waited for yet
    void *pipe_out[2] = {nullptr, nullptr};
    cudaEvent_t ev_pipe_done[2] = {nullptr, nullptr};
    int pipe_next = 0;
    bool pipe_pending = false;
    // copy-engine push of the feature all-gather (anx_push_to_peers): one stream per destination
    cudaStream_t push_stream[8] = {nullptr};
    cudaEvent_t ev_push_fork = nullptr, ev_push_done[8] = {nullptr};
    // depth-slab forward (anx_engine_forward_slab): exchange counter (same on every rank) and the ticket the
    // halo kernel's CTAs draw
    uint32_t slab_seq = 0;
    unsigned int *d_ticket = nullptr;
    mutable std::string last_error;

    anx_status fail(anx_status st, const char *fmt, ...) const {
        char tmp[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(tmp, sizeof tmp, fmt, ap);
        va_end(ap);
        last_error = tmp;
        return st;
    }
};

#define ANX_CUDA(e, call)                                                                            \
    do {                                                                                             \
        cudaError_t err__ = (call);                                                                  \
        if (err__ != cudaSuccess)                                                                    \
            return (e)->fail(ANX_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), \
                             __FILE__, __LINE__);                                                    \
    } while (0)

namespace {

int add_buffer(anx_engine *e, int level, int channels) {
    e->bufs.push_back(Buffer{level, channels / 8, 0});
    return (int)e->bufs.size() - 1;
}

void add_conv(anx_engine *e, int &module_index, int cin, int cout, int level, bool last, bool stem, int src, int dst,
              int dst_goff) {
    const anx_unet_desc &d = e->desc;
    ConvLayer c{};
    c.module_index = module_index;
    c.cin = cin;
    c.cout = cout;
    c.ncols = (cout + 15) / 16 * 16;
    c.level = level;
    c.has_norm = !last && d.norm_kind != ANX_NORM_NONE;
    c.has_act = !last && d.act_kind != ANX_ACT_NONE;
    c.is_stem = stem;
    c.is_final = last;
    c.src_buf = src;
    c.dst_buf = dst;
    c.dst_group_offset = dst_goff;
    c.inorm = c.has_norm && d.norm_kind == ANX_NORM_INSTANCE;
    c.n_splits = c.ncols > 256 ? c.ncols / 256 : 1;
    c.ncols_split = c.ncols / c.n_splits;
    c.alt_splits = (!stem && c.ncols >= 128 && c.ncols % 64 == 0) ? c.ncols / 64 : 0;
    module_index += 1 + (c.has_norm ? 1 : 0) + (c.has_act ? 1 : 0);
    Step s{};
    s.kind = stem ? STEP_STEM : STEP_CONV;
    s.conv = (int)e->convs.size();
    snprintf(s.name, sizeof s.name, "conv%d_%dto%d_L%d", c.module_index, cin, cout, level);
    e->logical.push_back(Logical{c.module_index, cin, cout, c.has_norm, (int)e->convs.size(), -1});
    e->convs.push_back(c);
    e->steps.push_back(s);
    if (c.inorm) {   // normalise + activate in place once the whole tensor's statistics exist
        Step nrm{};
        nrm.kind = STEP_NORM;
        nrm.conv = s.conv;
        nrm.dst_buf = dst;
        nrm.dst_group_offset = dst_goff;
        nrm.groups = cout / 8;
        snprintf(nrm.name, sizeof nrm.name, "inorm%d_L%d", c.module_index + 1, level);
        e->steps.push_back(nrm);
    }
}

// Walks the constructor logic of reference network.py:309-465 and lays out the
// buffers: every conv writes straight into the buffer its consumer reads; the
// second conv of encoder level i writes groups [0, w_i/8) of the level's concat
// buffer and the decoder's upsample fills the rest (zero-copy torch.cat,
// encoder channels first, network.py:545).
void build_program(anx_engine *e) {
    const anx_unet_desc &d = e->desc;
    const int nd = d.num_downs, g = d.ngf;
    int mi = 0;
    std::vector<int> cat(nd), width(nd + 1);
    for (int i = 0; i <= nd; ++i) width[i] = g << i;
    // (see below) decoder level 0 behind a nearest upsample never materialises the upsampled tensor
    const bool use_upconv = d.interp_kind == ANX_INTERP_NEAREST && d.norm_kind != ANX_NORM_INSTANCE &&
                            !(d.flags & (ANX_FLAG_FORCE_SIMT | ANX_FLAG_NO_UPCONV)) && width[0] == 16 &&
                            !exp_env("ANX_NO_UPCONV");   // the seeded skip conv handles one 16-channel chunk
    // concat buffers [skip | upsampled]; with the low-resolution decoder conv the level-0 one holds the skip only
    for (int i = 0; i < nd; ++i) cat[i] = add_buffer(e, i, (i == 0 && use_upconv) ? width[0] : 3 * width[i]);

    int cur = add_buffer(e, 0, g);
    add_conv(e, mi, d.input_nc, g, 0, false, true, -1, cur, 0);
    for (int i = 0; i < nd; ++i) {
        int t = add_buffer(e, i, width[i]);
        add_conv(e, mi, i == 0 ? g : width[i - 1], width[i], i, false, false, cur, t, 0);
        add_conv(e, mi, width[i], width[i], i, false, false, t, cat[i], 0);
        int p = add_buffer(e, i + 1, width[i]);
        if (!(d.flags & ANX_FLAG_FORCE_SIMT) && d.norm_kind != ANX_NORM_INSTANCE && !exp_env("ANX_NO_POOL_FUSION"))
            e->convs.back().pool_dst_buf = p;     // epilogue-fused when the tile shape allows (see make_geom)
        Step s{};
        s.kind = STEP_POOL;
        s.conv = (int)e->convs.size() - 1;
        s.src_buf = cat[i];
        s.dst_buf = p;
        s.groups = width[i] / 8;
        s.module_index = mi;
        snprintf(s.name, sizeof s.name, "pool%d_L%d", mi, i);
        e->steps.push_back(s);
        mi += 1;
        cur = p;
    }
    {
        int t = add_buffer(e, nd, width[nd]);
        add_conv(e, mi, width[nd - 1], width[nd], nd, false, false, cur, t, 0);
        int u = add_buffer(e, nd, width[nd]);
        add_conv(e, mi, width[nd], width[nd], nd, false, false, t, u, 0);
        cur = u;
    }
    // Decoder level 0 behind a NEAREST upsample: conv(cat(skip, up(L))) = conv_skip(skip) + conv_up(up(L)).
    // up(L) is piecewise constant, so conv_up(up(L)) at output voxel (2z+a, 2y+b, 2x+c) is a 2x2x2 conv of L
    // with parity-summed weights: all eight parities together are ONE 3x3x3 conv at LOW resolution with
    // 8*w output columns (zero weights where a parity does not see a tap) on a replicate-padded L (reflect
    // padding of the upsampled tensor equals replicate padding of L).  That launch runs with N = 8*w per A
    // tile instead of 3*w, never materialises the upsampled tensor, and stores its (shift-seeded) result
    // depth-to-space as 16-bit partial sums, which then seed the accumulators of the w -> w skip conv.
    std::vector<std::pair<int, int>> pairs;
    for (int l = nd - 1; l >= 0; --l) {
        if (use_upconv && l == 0) {
            e->bufs[cur].shell_rep = 1;
            mi += 1;                                     // the nn.Upsample slot
            const int w = width[l], cup = width[l + 1];
            const bool has_norm = d.norm_kind != ANX_NORM_NONE, has_act = d.act_kind != ANX_ACT_NONE;
            Logical lg{mi, 3 * w, w, has_norm, (int)e->convs.size(), (int)e->convs.size() + 1};
            ConvLayer a{};
            a.module_index = mi; a.cin = cup; a.cout = 8 * w; a.ncols = 8 * w; a.level = l + 1;
            a.has_norm = false; a.has_act = false; a.is_stem = false; a.is_final = false;
            a.src_buf = cur; a.dst_buf = -2; a.dst_group_offset = 0; a.d2s_cout = w;
            a.n_splits = a.ncols > 256 ? a.ncols / 256 : 1;
            a.ncols_split = a.ncols / a.n_splits;
            int v = add_buffer(e, l, w);
            ConvLayer b{};
            b.module_index = mi; b.cin = w; b.cout = w; b.ncols = (w + 15) / 16 * 16; b.level = l;
            b.has_norm = has_norm; b.has_act = has_act; b.is_stem = false; b.is_final = false;
            b.src_buf = cat[l]; b.dst_buf = v; b.dst_group_offset = 0; b.seed_buf = -2;
            b.n_splits = 1; b.ncols_split = b.ncols;
            Step sa{}, sb{};
            sa.kind = sb.kind = STEP_CONV;
            sa.conv = lg.conv_a; sb.conv = lg.conv_b;
            snprintf(sa.name, sizeof sa.name, "upconv%d_%dto8x%d_L%d", mi, cup, w, l + 1);
            snprintf(sb.name, sizeof sb.name, "conv%d_skip%dto%d_L%d", mi, w, w, l);
            pairs.push_back({lg.conv_a, lg.conv_b});
            e->logical.push_back(lg);
            e->convs.push_back(a);
            e->convs.push_back(b);
            e->steps.push_back(sa);
            e->steps.push_back(sb);
            mi += 1 + (has_norm ? 1 : 0) + (has_act ? 1 : 0);
            int x = add_buffer(e, l, w);
            add_conv(e, mi, w, w, l, false, false, v, x, 0);
            cur = x;
            continue;
        }
        Step s{};
        s.kind = STEP_UP;
        s.src_buf = cur;
        s.dst_buf = cat[l];
        s.groups = width[l + 1] / 8;
        s.dst_group_offset = width[l] / 8;
        s.module_index = mi;
        snprintf(s.name, sizeof s.name, "up%d_L%d", mi, l);
        e->steps.push_back(s);
        mi += 1;
        int v = add_buffer(e, l, width[l]);
        add_conv(e, mi, 3 * width[l], width[l], l, false, false, cat[l], v, 0);
        int x = add_buffer(e, l, width[l]);
        add_conv(e, mi, width[l], width[l], l, false, false, v, x, 0);
        cur = x;
    }
    add_conv(e, mi, g, d.output_nc, 0, true, false, cur, -1, 0);
    for (auto &pr : pairs) {   // partial-sum buffers go last so the other buffers keep their indices
        const int P = add_buffer(e, e->convs[pr.second].level, e->convs[pr.second].cout);
        e->convs[pr.first].dst_buf = P;
        e->convs[pr.second].seed_buf = P;
    }
}

// Which Sequential slots leave a tensor the engine actually stores (see TapSite).
void build_taps(anx_engine *e) {
    e->taps.clear();
    for (int si = 0; si < (int)e->steps.size(); ++si) {
        const Step &s = e->steps[si];
        if (s.kind == STEP_STEM || s.kind == STEP_CONV) {
            const ConvLayer &c = e->convs[s.conv];
            if (c.d2s_cout) continue;            // partial sums of a split decoder conv: not a network tensor
            int last = si;
            if (c.inorm) last = si + 1;          // the STEP_NORM that follows normalises + activates in place
            if (c.is_final) {
                e->taps.push_back(TapSite{c.module_index, -1, 0, (c.cout + 7) / 8, c.cout, 0, last});
                continue;
            }
            // The conv slot itself, when a norm follows: the reference taps the conv's output BEFORE the norm
            // (network.py:504-515; the pretraining defaults tap these slots).  The engine never stores that tensor
            // (BatchNorm is folded into the weights, InstanceNorm overwrites it in place): an un-folded clone of the
            // conv re-evaluates it on request.  Not for a conv that runs as two launches (see build_program).
            if (c.has_norm)
                for (int k = 0; k < (int)e->logical.size(); ++k)
                    if (e->logical[k].conv_a == s.conv && e->logical[k].conv_b < 0) {
                        TapSite t{c.module_index, -2, 0, c.cout / 8, c.cout, c.level, si};
                        t.prenorm_conv = k;
                        e->taps.push_back(t);
                    }
            // the stored tensor is the conv's output after norm and activation: the tap of the LAST slot of
            // the conv block (act if present, else norm, else the conv itself)
            const int idx = c.module_index + (c.has_norm ? 1 : 0) + (c.has_act ? 1 : 0);
            e->taps.push_back(TapSite{idx, c.dst_buf, c.dst_group_offset, c.cout / 8, c.cout, c.level, last});
            // The reference's activations are in-place modules (network.py:171-204): the tensor tapped at the
            // slot just before the activation is overwritten by it, so that tap holds the same values.
            if (c.has_act)
                e->taps.push_back(TapSite{idx - 1, c.dst_buf, c.dst_group_offset, c.cout / 8, c.cout, c.level, last});
        } else if (s.kind == STEP_POOL) {
            const Buffer &b = e->bufs[s.dst_buf];
            e->taps.push_back(TapSite{s.module_index, s.dst_buf, 0, s.groups, s.groups * 8, b.level, si});
        } else if (s.kind == STEP_UP) {
            // reference network.py:545: the tap after an Upsample slot sees cat(skip, upsampled)
            const Buffer &b = e->bufs[s.dst_buf];
            e->taps.push_back(TapSite{s.module_index, s.dst_buf, 0, b.groups, b.groups * 8, b.level, si});
        }
    }
}

// First writer / last user of every buffer over the launch sequence.
void build_liveness(anx_engine *e) {
    auto touch = [&](int buf, int step) {
        if (buf < 0) return;
        Buffer &b = e->bufs[buf];
        if (b.first < 0) b.first = step;
        b.last = std::max(b.last, step);
    };
    for (int si = 0; si < (int)e->steps.size(); ++si) {
        const Step &s = e->steps[si];
        if (s.kind == STEP_STEM || s.kind == STEP_CONV) {
            const ConvLayer &c = e->convs[s.conv];
            touch(c.src_buf, si);
            touch(c.seed_buf, si);
            touch(c.dst_buf, si);
            touch(c.pool_dst_buf, si);        // written here when the pooling is fused into this conv's epilogue
        } else {
            touch(s.src_buf, si);
            touch(s.dst_buf, si);
        }
    }
}

bool shape_ok(const anx_engine *e, int n, int d, int h, int w) {
    const int unit = 1 << e->desc.num_downs;
    if (n < 1 || d < 2 * unit || h < 2 * unit || w < 2 * unit) return false;
    return d % unit == 0 && h % unit == 0 && w % unit == 0;
}

size_t buffer_bytes(const anx_engine *e, const Buffer &b, int n, int d, int h, int w) {
    const size_t dp = (d >> b.level) + 2, hp = (h >> b.level) + 2, wp = layout_of(w >> b.level, e->x_lead).pitch;
    return align_up((size_t)n * b.groups * dp * hp * wp * 16, 256);
}

// Byte offset of every activation buffer inside the workspace; returns the bytes they span.  Buffers whose lifetimes
// (first writer .. last user) do not overlap share memory: placed largest first, each at the lowest offset where it
// collides with no already-placed buffer that is live at the same time.  Engines in depth-slab mode keep one
// region per buffer (a neighbour's halo plane may arrive while this rank is still several launches behind), as do
// engines created with ANX_FLAG_NO_WS_REUSE (debugging: every intermediate tensor survives the forward).
size_t plan_offsets(const anx_engine *e, int n, int d, int h, int w, std::vector<size_t> &off) {
    const size_t nb = e->bufs.size();
    off.assign(nb, 0);
    std::vector<size_t> bytes(nb);
    for (size_t i = 0; i < nb; ++i) bytes[i] = buffer_bytes(e, e->bufs[i], n, d, h, w);
    size_t total = 0;
    if (e->desc.flags & (ANX_FLAG_DEPTH_HALO_INPUT | ANX_FLAG_NO_WS_REUSE)) {
        for (size_t i = 0; i < nb; ++i) { off[i] = total; total += bytes[i]; }
        return total;
    }
    std::vector<size_t> order(nb);
    for (size_t i = 0; i < nb; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return bytes[a] > bytes[b]; });
    std::vector<size_t> placed;
    for (size_t k : order) {
        const Buffer &bk = e->bufs[k];
        size_t at = 0;
        bool moved = true;
        while (moved) {                                   // lowest offset free of live neighbours
            moved = false;
            for (size_t j : placed) {
                const Buffer &bj = e->bufs[j];
                const bool live_together = !(bk.last < bj.first || bj.last < bk.first);
                if (live_together && at < off[j] + bytes[j] && off[j] < at + bytes[k]) {
                    at = off[j] + bytes[j];
                    moved = true;
                }
            }
        }
        off[k] = at;
        placed.push_back(k);
        total = std::max(total, at + bytes[k]);
    }
    return total;
}

// Low-resolution half of a decoder conv: column (parity p = a*4 + b*2 + c, channel) of low-resolution tap offset o
// in {-1, 0, +1} along an axis is non-zero only for parity 0 (o = -1), both (o = 0) or parity 1 (o = +1) of that
// axis (see anx_engine_set_conv).  Span of 16-column blocks tap (kz, ky, kx) reaches; `blocks` = blocks per parity.
void trim_span(int kz, int ky, int kx, int blocks, int &lo, int &n) {
    auto span = [](int o, int &l, int &h) { l = o == 2 ? 1 : 0; h = o == 0 ? 0 : 1; };   // o = offset + 1
    int al, ah, bl, bh, cl, ch;
    span(kz, al, ah); span(ky, bl, bh); span(kx, cl, ch);
    const int pmin = al * 4 + bl * 2 + cl, pmax = ah * 4 + bh * 2 + ch;
    lo = pmin * blocks;
    n = (pmax - pmin + 1) * blocks;
}

// Output planes per unit of the row kernels: a unit streams zs + 2 input planes, and units are dealt round-robin to
// the persistent CTAs, so the launch takes ceil(units / SMs) * (zs + 2) plane times.  Longer z segments have less halo
// but fewer units: pick the segment length (a divisor of D among 16, 32, 64) that minimises that product.
constexpr int ROWS_RPW1_DEFAULT = 1 | 2 | 4;   // see anx_engine::rows_rpw1 (measured: padded -4 %, stem -14 %, seeded -15 %, fp32 +7 %)

// Widths the row kernels take: whole 128-voxel x tiles, or a last tile pulled back to the border when the voxels it
// recomputes are at most a fifth of the row (224 = 128 + 96: 14 % extra work beats the generic tile kernel; 160 does not).
bool rows_width_ok(int W) { return W >= ROWS_X && ((W + ROWS_X - 1) / ROWS_X) * ROWS_X * 5 <= W * 6; }

int rows_segment(const anx_engine *e, int N, int D, int H, int W) {
    int best = 16;
    size_t best_cost = ~(size_t)0;
    for (int zs : {16, 32, 64}) {
        if (D % zs) continue;
        const size_t units = (size_t)N * ((W + ROWS_X - 1) / ROWS_X) * (H / ROWS_YB) * (D / zs);
        const size_t cost = ((units + e->num_sms - 1) / e->num_sms) * (size_t)(zs + 2);
        if (cost < best_cost) { best_cost = cost; best = zs; }
    }
    return best;
}

// Tile / pipeline configuration of one tensor-core conv at one shape.
ConvGeom make_geom(const anx_engine *e, const ConvLayer &c, int N, int D, int H, int W, int in_groups_total) {
    ConvGeom g{};
    g.N = N; g.D = D; g.H = H; g.W = W;
    g.ncols = c.ncols_split;
    g.n_splits = c.n_splits;
    g.dt = e->dt;
    g.fold = c.fold;
    g.groups = c.groups;
    // Small problems on wide layers (the 16^3 / 8^3 levels, above all at the batch-2 call shape of the sliding-window
    // predictor): few tiles, each a long serial chain of N = 128 / 256 MMAs, leave most SMs idle.  When the regular
    // tiling fills less than half the SMs the launch uses the layer's second packing: unfolded 64-column channel
    // splits, one plane per tile -- (ncols / 64) x more tiles, each with cheaper MMAs (consecutive CTAs share one
    // activation brick in L2).  More MMA time per FLOP, so only then.
    bool alt = false;
    if (c.alt_splits > 0 && c.d_wpack_alt && !exp_env("ANX_NO_ALT_SPLITS")) {
        // tiles of the alternative tiling (one plane each): taken only while they still fit one wave -- measured at
        // batch 8 (256 tiles): conv34 0.060 -> 0.070 ms; at batch 2 (64 tiles): 0.059 -> 0.041 ms
        const size_t alt_tiles = (size_t)((W + TILE_X - 1) / TILE_X) * ((H + TILE_Y - 1) / TILE_Y) * D * N * c.alt_splits;
        alt = alt_tiles <= (size_t)e->num_sms;
    }
    if (alt) {
        g.ncols = 64;
        g.n_splits = c.alt_splits;
        g.fold = 0;
        g.groups = 3;
        g.alt = 1;
    }
    const bool fold = g.fold != 0;
    const int n_splits = g.n_splits;
    g.cin_chunks = c.cin / 16;
    g.in_groups_total = in_groups_total;
    g.in_group_offset = 0;
    // Single-slab layers (one 16-channel chunk, folded dz, no channel split: the 16 -> 16 convs) use the
    // same B image for every tile: it is loaded once per CTA and stays in shared memory.
    g.b_static = (fold && g.cin_chunks == 1 && n_splits == 1 && !exp_env("ANX_NO_BSTATIC")) ? 1 : 0;
    // output planes per tile: as many as TMEM double buffering allows, at most 8 -- or 16 for the thin
    // single-slab layers, whose MMA count per output plane is 9 * (bz + 2) / bz (all of TMEM, two A stages)
    int bz = std::max(1, std::min(8, 256 / g.ncols));
    if (g.b_static && g.ncols == 16 && D >= 16 && !exp_env("ANX_NO_BZ16")) bz = 16;
    g.acc_stages = 2;
    bz = std::min(bz, D);
    if (alt) bz = 1;
    // Unfolded (wide) layers at the deep levels: a tile is a long serial chain of MMAs (27 taps x Cin / 16 per
    // plane) and there are few tiles, so with two planes per tile most SMs idle while a few grind: one plane per
    // tile doubles the parallelism at no extra MMA work (unfolded tiles have no z halo cost in the MMAs).  Only
    // while the doubled tile count still fits one wave: every tile streams the layer's whole weight set from
    // L2, so past one wave the extra weight traffic costs more than the parallelism brings (measured at batch 8:
    // conv38 0.086 -> 0.110 ms with 256 tiles; at batch 2: 0.086 -> 0.060 ms with 64).
    if (!fold && bz > 1 && !exp_env("ANX_NO_BZ1")) {
        const size_t tiles = (size_t)((W + TILE_X - 1) / TILE_X) * ((H + TILE_Y - 1) / TILE_Y) * ((D + bz - 1) / bz) * N * n_splits;
        if (2 * tiles <= (size_t)e->num_sms) bz = 1;
    }
    g.bz = bz;
    g.fuse_pool = (c.pool_dst_buf >= 0 && bz % 4 == 0) ? 1 : 0;
    int cols = g.acc_stages * bz * g.ncols;
    int p2 = 32;
    while (p2 < cols) p2 *= 2;
    g.tmem_cols = p2;
    g.tiles_x = (W + TILE_X - 1) / TILE_X;
    g.tiles_y = (H + TILE_Y - 1) / TILE_Y;
    g.tiles_z = (D + bz - 1) / bz;
    g.tiles_per_sample = g.tiles_x * g.tiles_y * g.tiles_z;
    g.total_tiles = g.tiles_per_sample * N * g.n_splits;
    g.a_lbo = (uint32_t)(bz + 2) * HALO_Y * ROW_BYTES;
    g.a_stage_bytes = 2 * g.a_lbo;
    g.b_rows = fold ? 3 * g.ncols : g.ncols;
    g.b_stage_bytes = 9 * 32 * g.b_rows;
    // InstanceNorm statistics of thin layers are accumulated per CTA in shared memory (one flush per sample instead
    // of one double atomic per warp, chunk and tile)
    g.stats_acc = (c.inorm && g.ncols <= STATS_ACC_MAX_COLS && n_splits == 1) ? 1 : 0;
    const size_t shared_tail = sizeof(UmmaShared) + (g.stats_acc ? STATS_ACC_BYTES : 0);
    const size_t budget = (size_t)e->max_smem - shared_tail - 1024;
    g.a_stages = 3;
    g.b_stages = 2;
    const size_t b_reserve = (g.b_static ? 1u : 2u) * (size_t)g.b_stage_bytes;
    while (g.a_stages > 1 && (size_t)g.a_stages * g.a_stage_bytes + b_reserve > budget) --g.a_stages;
    size_t left = budget - (size_t)g.a_stages * g.a_stage_bytes;
    g.b_stages = (int)std::min<size_t>(MAX_B_STAGES, left / g.b_stage_bytes);
    g.b_stages = std::min(g.b_stages, std::max(2, 2 * g.groups));
    if (g.b_static) g.b_stages = std::min(g.b_stages, 1);
    if (const char *ab = exp_env("ANX_ABLATE")) g.ablate = (uint32_t)atoi(ab);
    g.smem_bytes = (uint32_t)((size_t)g.a_stages * g.a_stage_bytes + (size_t)g.b_stages * g.b_stage_bytes + shared_tail);
    // Low-resolution half of a decoder conv: each tap's MMA covers just the span of its parities (trim_span); when the
    // compact tiles of the whole layer fit next to the A ring they stay resident in shared memory, loaded once per CTA
    // instead of streaming 221 KB of (mostly zero) slabs from L2 for every tile.
    if (!alt && c.d2s_cout > 0 && !c.fold && c.n_splits == 1 && c.d2s_cout % 16 == 0 && !exp_env("ANX_NO_TRIM")) {
        g.trim = 1;
        const int blocks = c.d2s_cout / 16;        // 16-column blocks per parity
        uint32_t off16 = 0;
        for (int t = 0; t < 27; ++t) {
            int lo, n;
            trim_span(t / 9, (t % 9) / 3, t % 3, blocks, lo, n);
            g.trim_lo[t] = (uint8_t)lo;
            g.trim_n[t] = (uint8_t)n;
        }
        if (c.d_wtrim && g.cin_chunks <= 4 && !exp_env("ANX_NO_BRESIDENT")) {
            for (int ch = 0; ch < g.cin_chunks; ++ch)
                for (int t = 0; t < 27; ++t) {
                    g.trim_off[ch * 27 + t] = (uint16_t)off16;
                    off16 += 2u * 16u * g.trim_n[t];             // two K halves x rows, one 16-byte unit per row
                }
            const size_t need = (size_t)g.a_stages * g.a_stage_bytes + (size_t)off16 * 16 + sizeof(UmmaShared);
            if ((size_t)off16 * 16 == c.wtrim_bytes && need + 1024 <= (size_t)e->max_smem && off16 < 65536u) {
                g.trim = 2;
                g.b_static = 1;
                g.b_stages = 1;
                g.b_stage_bytes = (uint32_t)c.wtrim_bytes;
                g.trim_bytes = (uint32_t)c.wtrim_bytes;
                g.smem_bytes = (uint32_t)need;
            }
        }
    }
    return g;
}

ActView view_of(const anx_engine *e, const ShapePlan &p, int buf, int group_offset) {
    const Buffer &b = e->bufs[buf];
    ActView v;
    v.base = reinterpret_cast<__nv_bfloat16 *>(static_cast<char *>(p.workspace) + p.buf_offset[buf]);
    v.groups_total = b.groups;
    v.group_offset = group_offset;
    v.D = p.D >> b.level;
    v.H = p.H >> b.level;
    v.W = p.W >> b.level;
    v.shell_rep = b.shell_rep;
    v.z_open = (e->slab_lower ? 1 : 0) | (e->slab_upper ? 2 : 0);
    const RowLayout rl = layout_of(v.W, e->x_lead);
    v.lead = rl.lead;
    v.pitch = rl.pitch;
    return v;
}

anx_status get_plan(anx_engine *e, int N, int D, int H, int W, void *workspace, std::shared_ptr<ShapePlan> &out) {
    std::lock_guard<std::mutex> lock(e->mu);
    for (auto &p : e->plans)
        if (p->N == N && p->D == D && p->H == H && p->W == W && p->workspace == workspace) {
            out = p;
            return ANX_OK;
        }
    EncodeTiledFn encode = load_encode_tiled();
    if (!encode) return e->fail(ANX_ERR_NO_DEVICE, "cuTensorMapEncodeTiled entry point not available");
    auto p = std::make_shared<ShapePlan>();
    p->N = N; p->D = D; p->H = H; p->W = W;
    p->workspace = workspace;
    size_t off = plan_offsets(e, N, D, H, W, p->buf_offset);
    p->stats_offset = off;
    p->conv_stats_offset.assign(e->convs.size(), 0);
    for (size_t i = 0; i < e->convs.size(); ++i)
        if (e->convs[i].inorm) {
            p->conv_stats_offset[i] = off;
            off += align_up((size_t)N * e->convs[i].ncols * 2 * sizeof(double), 256);
        }
    p->stats_bytes = off - p->stats_offset;
    p->tmaps.resize(e->convs.size());
    p->geoms.resize(e->convs.size());
    for (size_t i = 0; i < e->convs.size(); ++i) {
        const ConvLayer &c = e->convs[i];
        if (c.is_stem) continue;
        const Buffer &sb = e->bufs[c.src_buf];
        const int d = D >> c.level, h = H >> c.level, w = W >> c.level;
        p->geoms[i] = make_geom(e, c, N, d, h, w, sb.groups);
        const ConvGeom &g = p->geoms[i];
        if (g.smem_bytes > (uint32_t)e->max_smem || g.a_stages < 1 || g.b_stages < 1)
            return e->fail(ANX_ERR_UNSUPPORTED, "conv %d (%d->%d) does not fit shared memory", c.module_index, c.cin,
                           c.cout);
        // 4-D view of the padded planar buffer: [n*groups][zp][yp][xp*8 + ch], rows of `pitch` voxels
        // with the x shell starting `lead` voxels into each row
        const RowLayout rl = layout_of(w, e->x_lead);
        const cuuint64_t pitch = (cuuint64_t)rl.pitch;
        cuuint64_t dims[4] = {(cuuint64_t)(w + 2) * 8, (cuuint64_t)(h + 2), (cuuint64_t)(d + 2),
                              (cuuint64_t)N * sb.groups};
        cuuint64_t strides[3] = {pitch * 16, pitch * (h + 2) * 16, pitch * (h + 2) * (d + 2) * 16};
        cuuint32_t box[4] = {HALO_X * 8, HALO_Y, (cuuint32_t)(g.bz + 2), 2};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        void *base = static_cast<char *>(workspace) + p->buf_offset[c.src_buf] + (size_t)rl.lead * 16;
        CUresult r = encode(&p->tmaps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
            return e->fail(ANX_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for conv %d", (int)r, c.module_index);
    }
    if (e->plans.size() >= 8) e->plans.erase(e->plans.begin());
    e->plans.push_back(p);
    out = p;
    return ANX_OK;
}

Epilogue make_epilogue(const anx_engine *e, const ShapePlan &p, const ConvLayer &c, float *out,
                       const GatherArgs *ga = nullptr, const ConvGeom *g = nullptr) {
    Epilogue ep{};
    ep.pool_kind = -1;
    ep.d2s_cout = c.d2s_cout;
    ep.seed_on = 0;
    if (c.seed_buf >= 0 && g) {
        ep.seed_on = 1;
        ep.seed_src = view_of(e, p, c.seed_buf, 0);
    }
    if (g && g->fuse_pool) {
        ep.pool_kind = e->desc.pool_kind;
        ep.pool_dst = view_of(e, p, c.pool_dst_buf, 0);
    }
    ep.head_nc = 0;
    ep.head = nullptr;
    if (c.is_final && !c.is_tap && e->head_nc > 0) {
        ep.head_nc = e->head_nc;
        ep.head = e->d_head;
    }
    if (ga && c.is_final) {
        for (int i = 0; i < ga->n_peers; ++i) ep.out_peers[i] = ga->peers[i];
        ep.n_peers = ga->n_peers;
        ep.sample_offset = ga->sample_offset;
        ep.cl16 = ga->payload == ANX_PAYLOAD_CL16 ? 1 : 0;
    }
    if (c.is_final) {
        const size_t chans = (e->head_nc > 0 && !c.is_tap) ? (size_t)e->head_nc : (size_t)c.cout;
        ep.out_nstride = (ga && ga->out_nstride) ? ga->out_nstride
                                                 : chans * (size_t)(p.D >> c.level) * (p.H >> c.level) * (p.W >> c.level);
    }
    ep.cout = c.cout;
    ep.bias = c.d_bias;
    // with InstanceNorm the conv stores its raw output; the activation runs after the normalisation
    ep.act = (c.has_act && !c.inorm) ? e->desc.act_kind : ANX_ACT_NONE;
    ep.slope = e->desc.act_slope;
    ep.dt = e->dt;
    ep.stats = nullptr;
    ep.stats_stride = c.ncols;
    if (c.inorm) {
        const size_t k = &c - e->convs.data();
        ep.stats = reinterpret_cast<double *>(static_cast<char *>(p.workspace) + p.conv_stats_offset[k]);
    }
    if (c.is_final) {
        ep.mode = OUT_NCDHW_F32;
        ep.out_f32 = out;
        ep.dst.base = nullptr;
        ep.dst.groups_total = 0;
        ep.dst.group_offset = 0;
        ep.dst.D = p.D >> c.level; ep.dst.H = p.H >> c.level; ep.dst.W = p.W >> c.level;   // level 0 except for tap clones
        ep.dst.lead = 0; ep.dst.pitch = p.W >> c.level;
    } else {
        ep.mode = OUT_PADDED_BF16;
        ep.dst = view_of(e, p, c.dst_buf, c.dst_group_offset);
        ep.out_f32 = nullptr;
    }
    return ep;
}

int grid_for(size_t work_items, int threads, int num_sms, int waves) {
    size_t blocks = (work_items + threads - 1) / threads;
    size_t cap = (size_t)num_sms * waves;
    return (int)std::max<size_t>(1, std::min(blocks, cap));
}

anx_status launch_step(anx_engine *e, const ShapePlan &p, const Step &s, const float *in, float *out,
                       cudaStream_t st, const GatherArgs *ga = nullptr, const ConvLayer *conv_override = nullptr,
                       const ConvGeom *geom_override = nullptr) {
    const bool force_simt = (e->desc.flags & ANX_FLAG_FORCE_SIMT) != 0;
    switch (s.kind) {
    case STEP_STEM: {
        const ConvLayer &c = conv_override ? *conv_override : e->convs[s.conv];
        Epilogue ep = make_epilogue(e, p, c, out);
        const int zh0 = (e->desc.flags & ANX_FLAG_DEPTH_HALO_INPUT) ? 1 : 0;
        if (!force_simt && e->use_rows && c.d_wstem_rows && !ep.stats && rows_width_ok(p.W) && p.H % ROWS_YB == 0 &&
            p.D % 16 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && !exp_env("ANX_NO_STEM_ROWS")) {
            // row-form stem: the 16 -> 16 row kernel's pipeline with builder warps in front of it
            RowsGeom rg{};
            rg.N = p.N; rg.D = p.D; rg.H = p.H; rg.W = p.W;
            rg.zs = rows_segment(e, p.N, p.D, p.H, p.W);
            rg.tiles_x = (p.W + ROWS_X - 1) / ROWS_X; rg.tiles_y = p.H / ROWS_YB; rg.tiles_z = p.D / rg.zs;
            rg.units_per_sample = rg.tiles_x * rg.tiles_y * rg.tiles_z;
            rg.total_units = rg.units_per_sample * p.N;
            rg.dt = e->dt;
            rg.z_halo = zh0;
            if (const char *ab = exp_env("ANX_ABLATE")) rg.ablate = (uint32_t)atoi(ab);
            rg.smem_bytes = (uint32_t)(ROWS_STAGES * ROWS_PLANE_BYTES + 3 * ROWS_B_IMAGE_BYTES + sizeof(RowsShared));
            const int grid = std::min(rg.total_units, e->num_sms);
            if (ep.mode == OUT_NCDHW_F32)      // pre-norm tap of the stem
                conv3_rows_kernel<EPI_F32, true><<<grid, ROWS_STEM_THREADS, rg.smem_bytes, st>>>(
                    ActView{}, rg, (const uint8_t *)c.d_wstem_rows, ep, in);
            else if (e->rows_rpw1 & 2)
                conv3_rows_kernel<EPI_PADDED, true, 1><<<grid, rows_threads(1, true), rg.smem_bytes, st>>>(
                    ActView{}, rg, (const uint8_t *)c.d_wstem_rows, ep, in);
            else
                conv3_rows_kernel<EPI_PADDED, true><<<grid, ROWS_STEM_THREADS, rg.smem_bytes, st>>>(
                    ActView{}, rg, (const uint8_t *)c.d_wstem_rows, ep, in);
            break;
        }
        if (!force_simt && c.d_wstem && p.W % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
            !exp_env("ANX_SIMT_STEM")) {
            // tensor-core stem: per-call tensor map over the caller's fp32 input
            EncodeTiledFn encode = load_encode_tiled();
            StemGeom g{};
            g.N = p.N; g.D = p.D; g.H = p.H; g.W = p.W; g.cin = c.cin;
            g.ncols = c.ncols;
            g.kq = (9 * c.cin + 15) / 16;
            g.bz = std::min(std::max(1, std::min(8, 256 / c.ncols)), p.D);
            if (c.ncols == 16 && c.cin == 1 && p.D >= 16 && !exp_env("ANX_NO_BZ16")) g.bz = 16;   // all of TMEM: 18 input planes per 16
            g.acc_stages = 2;
            int cols = 32;
            while (cols < g.acc_stages * g.bz * g.ncols) cols *= 2;
            g.tmem_cols = cols;
            g.z_halo = zh0;
            if (const char *ds = exp_env("ANX_STEM_SHIFT")) g.dbg_shift = atoi(ds);
            if (const char *ab = exp_env("ANX_ABLATE")) g.ablate = (uint32_t)atoi(ab);
            g.tiles_x = (p.W + TILE_X - 1) / TILE_X;
            g.tiles_y = (p.H + TILE_Y - 1) / TILE_Y;
            g.tiles_z = (p.D + g.bz - 1) / g.bz;
            g.tiles_per_sample = g.tiles_x * g.tiles_y * g.tiles_z;
            g.total_tiles = g.tiles_per_sample * p.N;
            g.brick_bytes = (uint32_t)(c.cin * (g.bz + 2) * HALO_Y * STEM_BRICK_X * 4);
            g.a_tile_bytes = (uint32_t)(g.kq * 2 * 128 * 16);
            g.b_bytes = (uint32_t)(g.kq * 2 * 3 * g.ncols * 16);
            g.stats_acc = (ep.stats && g.ncols <= STATS_ACC_MAX_COLS) ? 1 : 0;
            g.smem_bytes = stem_bricks(g.kq) * ((g.brick_bytes + 127) & ~127u) + stem_a_slots(g.kq) * 2 * g.a_tile_bytes + 2 * g.b_bytes +
                           (uint32_t)sizeof(StemShared) + (g.stats_acc ? STATS_ACC_BYTES : 0);
            if (g.smem_bytes > (uint32_t)e->max_smem)
                return e->fail(ANX_ERR_UNSUPPORTED, "stem tile does not fit shared memory (%u bytes)", g.smem_bytes);
            CUtensorMap tm;
            cuuint64_t dims[4] = {(cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)(p.D + 2 * zh0), (cuuint64_t)p.N * c.cin};
            cuuint64_t strides[3] = {(cuuint64_t)p.W * 4, (cuuint64_t)p.W * p.H * 4,
                                     (cuuint64_t)p.W * p.H * (p.D + 2 * zh0) * 4};
            cuuint32_t box[4] = {STEM_BRICK_X, HALO_Y, (cuuint32_t)(g.bz + 2), (cuuint32_t)c.cin};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(in), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return e->fail(ANX_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for the stem input", (int)r);
            const int grid = std::min(g.total_tiles, e->num_sms);
            const uint8_t *ws_ = (const uint8_t *)c.d_wstem;
#define ANX_STEM(KQ_, MODE_) stem_umma_kernel<KQ_, MODE_><<<grid, STEM_THREADS, g.smem_bytes, st>>>(tm, g, ws_, ep)
            if (ep.mode == OUT_NCDHW_F32) {   // pre-norm tap of the stem
                if (g.kq == 1) ANX_STEM(1, EPI_F32); else if (g.kq == 2) ANX_STEM(2, EPI_F32); else ANX_STEM(3, EPI_F32);
            } else if (ep.stats) {
                if (g.kq == 1) ANX_STEM(1, EPI_STATS); else if (g.kq == 2) ANX_STEM(2, EPI_STATS); else ANX_STEM(3, EPI_STATS);
            } else {
                if (g.kq == 1) ANX_STEM(1, EPI_PADDED); else if (g.kq == 2) ANX_STEM(2, EPI_PADDED); else ANX_STEM(3, EPI_PADDED);
            }
#undef ANX_STEM
            break;
        }
        const size_t sm = (size_t)c.cin * 27 * c.ncols * sizeof(float);
        const float *wp = (const float *)c.d_wpack;
        auto grid_of = [&](int zt) { return dim3((p.W + 31) / 32, (p.H + 7) / 8, p.N * ((p.D + zt - 1) / zt)); };
        const int zh = (e->desc.flags & ANX_FLAG_DEPTH_HALO_INPUT) ? 1 : 0;
        if (c.ncols == 16)
            stem_conv_kernel<16, 4><<<grid_of(4), 256, sm, st>>>(in, wp, c.cin, p.N, p.D, p.H, p.W, zh, ep);
        else if (c.ncols == 32)
            stem_conv_kernel<32, 2><<<grid_of(2), 256, sm, st>>>(in, wp, c.cin, p.N, p.D, p.H, p.W, zh, ep);
        else if (c.ncols == 48)
            stem_conv_kernel<48, 1><<<grid_of(1), 256, sm, st>>>(in, wp, c.cin, p.N, p.D, p.H, p.W, zh, ep);
        else
            stem_conv_kernel<64, 1><<<grid_of(1), 256, sm, st>>>(in, wp, c.cin, p.N, p.D, p.H, p.W, zh, ep);
        break;
    }
    case STEP_CONV: {
        const ConvLayer &c = conv_override ? *conv_override : e->convs[s.conv];
        const ConvGeom &g = geom_override ? *geom_override : p.geoms[s.conv];
        Epilogue ep = make_epilogue(e, p, c, out, ga, force_simt ? nullptr : &g);
        if (ep.head_nc > 0 && (force_simt || ep.n_peers > 0 || ep.cl16))
            return e->fail(ANX_ERR_UNSUPPORTED, "a fused output head needs the tensor-core path with the plain fp32 output");
        if (ep.cl16 && (force_simt || (c.cout & 7)))
            return e->fail(ANX_ERR_UNSUPPORTED, "the 16-bit channels-last output needs the tensor-core path and output_nc % 8 == 0");
        if (!force_simt && e->use_rows && c.d_wrows && !(g.fuse_pool && ep.pool_kind != 0) && !ep.stats && !ep.d2s_cout &&
            !((e->use_rows & 2) && (g.fuse_pool || ep.seed_on)) &&      // ANX_ROWS=3: plain / fp32 variants only
            !(ep.cl16 && c.cout != 16) && rows_width_ok(g.W) && g.H % ROWS_YB == 0 && g.D % 16 == 0) {
            RowsGeom rg{};
            rg.N = g.N; rg.D = g.D; rg.H = g.H; rg.W = g.W;
            rg.zs = rows_segment(e, g.N, g.D, g.H, g.W);
            rg.tiles_x = (g.W + ROWS_X - 1) / ROWS_X; rg.tiles_y = g.H / ROWS_YB; rg.tiles_z = g.D / rg.zs;
            rg.units_per_sample = rg.tiles_x * rg.tiles_y * rg.tiles_z;
            rg.total_units = rg.units_per_sample * g.N;
            rg.dt = e->dt;
            rg.ablate = g.ablate;
            rg.smem_bytes = (uint32_t)(ROWS_STAGES * ROWS_PLANE_BYTES + 3 * ROWS_B_IMAGE_BYTES + sizeof(RowsShared));
            const ActView src = view_of(e, p, c.src_buf, 0);
            const int grid = std::min(rg.total_units, e->num_sms);
            const uint8_t *wr = (const uint8_t *)c.d_wrows;
#define ANX_ROWS_LAUNCH(MODE_) conv3_rows_kernel<MODE_><<<grid, ROWS_THREADS, rg.smem_bytes, st>>>(src, rg, wr, ep, nullptr)
#define ANX_ROWS_LAUNCH1(MODE_) conv3_rows_kernel<MODE_, false, 1><<<grid, rows_threads(1, false), rg.smem_bytes, st>>>(src, rg, wr, ep, nullptr)
            if (ep.mode == OUT_NCDHW_F32) {
                const bool one = (e->rows_rpw1 & 8) != 0;
                if (ep.head_nc > 0) { if (one) ANX_ROWS_LAUNCH1(EPI_F32_HEAD); else ANX_ROWS_LAUNCH(EPI_F32_HEAD); }
                else if (ep.cl16) { if (one) ANX_ROWS_LAUNCH1(EPI_CL16); else ANX_ROWS_LAUNCH(EPI_CL16); }
                else { if (one) ANX_ROWS_LAUNCH1(EPI_F32); else ANX_ROWS_LAUNCH(EPI_F32); }   // also the fused gather (peer loop inside)
            }
            else if (ep.seed_on) { if (e->rows_rpw1 & 4) ANX_ROWS_LAUNCH1(EPI_SEEDED); else ANX_ROWS_LAUNCH(EPI_SEEDED); }
            else if (g.fuse_pool) ANX_ROWS_LAUNCH(EPI_POOL);
            else { if (e->rows_rpw1 & 1) ANX_ROWS_LAUNCH1(EPI_PADDED); else ANX_ROWS_LAUNCH(EPI_PADDED); }
#undef ANX_ROWS_LAUNCH
#undef ANX_ROWS_LAUNCH1
            break;
        }
        if (force_simt) {
            ActView src = view_of(e, p, c.src_buf, 0);
            const size_t items = (size_t)g.N * g.D * g.H * g.W * (g.ncols * g.n_splits / 16);
            conv3_simt_kernel<<<grid_for(items, 128, e->num_sms, 64), 128, 0, st>>>(
                src, g, (const __nv_bfloat16 *)(g.alt ? c.d_wpack_alt : c.d_wpack), ep);
        } else {
            const int grid = std::min(g.total_tiles, e->num_sms);
            const uint8_t *wp_ = (const uint8_t *)(g.trim == 2 ? c.d_wtrim : (g.alt ? c.d_wpack_alt : c.d_wpack));
#define ANX_CONV(MODE_) conv3_umma_kernel<MODE_><<<grid, UMMA_THREADS, g.smem_bytes, st>>>(p.tmaps[s.conv], g, wp_, ep)
            if (ep.mode == OUT_NCDHW_F32) {
                if (ep.cl16) ANX_CONV(EPI_CL16);
                else if (ep.n_peers > 0) ANX_CONV(EPI_F32_PEERS);
                else if (ep.head_nc > 0) ANX_CONV(EPI_F32_HEAD);
                else ANX_CONV(EPI_F32);
            }
            else if (ep.seed_on) ANX_CONV(EPI_SEEDED);
            else if (ep.stats) ANX_CONV(EPI_STATS);
            else if (ep.d2s_cout) ANX_CONV(EPI_D2S);
            else if (ep.pool_kind >= 0) ANX_CONV(EPI_POOL);
            else ANX_CONV(EPI_PADDED);
#undef ANX_CONV
        }
        break;
    }
    case STEP_POOL: {
        if (!force_simt && p.geoms[s.conv].fuse_pool) return ANX_OK;   // already written by the conv's epilogue
        ActView src = view_of(e, p, s.src_buf, 0), dst = view_of(e, p, s.dst_buf, 0);
        const size_t items = (size_t)p.N * s.groups * dst.D * dst.H * dst.W;
        if (dst.D >= 4 && dst.H >= 4 && dst.W >= 32 && dst.D <= 65535 && (size_t)p.N * s.groups <= 65535) {
            const dim3 grid(dst.H, dst.D, p.N * s.groups);
            const int threads = std::min(128, (dst.W + 31) / 32 * 32);
            if (e->dt == DT_BF16) pool2_grid_kernel<DT_BF16><<<grid, threads, 0, st>>>(src, dst, s.groups, e->desc.pool_kind);
            else pool2_grid_kernel<DT_FP16><<<grid, threads, 0, st>>>(src, dst, s.groups, e->desc.pool_kind);
            break;
        }
        pool2_kernel<<<grid_for(items, 256, e->num_sms, 32), 256, 0, st>>>(src, dst, p.N, s.groups,
                                                                           e->desc.pool_kind, e->dt);
        break;
    }
    case STEP_UP: {
        ActView src = view_of(e, p, s.src_buf, 0), dst = view_of(e, p, s.dst_buf, s.dst_group_offset);
        const size_t items = (size_t)p.N * s.groups * dst.D * dst.H * dst.W;
        if (e->desc.interp_kind == ANX_INTERP_NEAREST) {
            if (dst.D >= 4 && dst.H >= 4 && dst.W >= 32 && src.D <= 65535 && (size_t)p.N * s.groups <= 65535)
                upsample2_nearest_grid_kernel<<<dim3(src.H, src.D, p.N * s.groups), std::min(128, dst.W), 0, st>>>(src, dst, s.groups);
            else
                upsample2_nearest_kernel<<<grid_for(items / 4, 256, e->num_sms, 32), 256, 0, st>>>(src, dst, p.N, s.groups);
        }
        else if (dst.D >= 4 && dst.H >= 4 && dst.W >= 32 && dst.D <= 65535 && (size_t)p.N * s.groups <= 65535) {
            // one thread per low-resolution voxel (27 loads per 8 outputs, separable weights)
            const dim3 grid(src.H, src.D, p.N * s.groups);
            const int threads = std::min(64, (src.W + 31) / 32 * 32);
            if (e->dt == DT_BF16)
                upsample2_tri_block_kernel<DT_BF16><<<grid, threads, 0, st>>>(src, dst, s.groups, e->slab_lower, e->slab_upper);
            else
                upsample2_tri_block_kernel<DT_FP16><<<grid, threads, 0, st>>>(src, dst, s.groups, e->slab_lower, e->slab_upper);
        }
        else
            upsample2_kernel<<<grid_for(items, 256, e->num_sms, 32), 256, 0, st>>>(
                src, dst, p.N, s.groups, e->desc.interp_kind, e->dt, e->slab_lower, e->slab_upper);
        break;
    }
    case STEP_NORM: {
        const ConvLayer &c = e->convs[s.conv];
        ActView t = view_of(e, p, s.dst_buf, s.dst_group_offset);
        const size_t count = (size_t)(t.D + 2) * (t.H + 2) * (t.W + 2);
        // statistics are over the whole volume: in depth-slab mode the caller has all-reduced the sums
        const int depth = e->slab_depth_total > 0 ? e->slab_depth_total >> e->bufs[s.dst_buf].level : t.D;
        const double inv = 1.0 / ((double)depth * t.H * t.W);
        const unsigned gx = (unsigned)std::max<size_t>(1, std::min<size_t>(1024, (count + 2047) / 2048));
        const double *stats = reinterpret_cast<const double *>(static_cast<char *>(p.workspace) + p.conv_stats_offset[s.conv]);
        inorm_act_kernel<<<dim3(gx, (unsigned)(p.N * s.groups)), 256, 0, st>>>(
            t, s.groups, stats, c.ncols, inv, e->desc.norm_eps, c.has_act ? e->desc.act_kind : ANX_ACT_NONE,
            e->desc.act_slope, e->dt);
        break;
    }
    }
    ANX_CUDA(e, cudaGetLastError());
    return ANX_OK;
}

anx_status check_forward_args(anx_engine *e, const void *in, const void *out, int n, int d, int h, int w,
                              void *workspace, size_t ws_bytes) {
    if (!e) return ANX_ERR_BAD_ARG;
    if (!in || !out) return e->fail(ANX_ERR_BAD_ARG, "null input/output pointer");
    if (!shape_ok(e, n, d, h, w))
        return e->fail(ANX_ERR_BAD_SHAPE,
                       "input [%d,%d,%d,%d,%d]: each of D,H,W must be a multiple of %d and >= %d", n,
                       e->desc.input_nc, d, h, w, 1 << e->desc.num_downs, 2 << e->desc.num_downs);
    for (auto &c : e->convs)
        if (!c.ready) return e->fail(ANX_ERR_NOT_READY, "conv at module index %d has no parameters", c.module_index);
    const size_t need = anx_engine_workspace_bytes(e, n, d, h, w);
    if (!workspace || ws_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255))
        return e->fail(ANX_ERR_WORKSPACE, "workspace %p of %zu bytes; need %zu bytes, 256-byte aligned", workspace,
                       ws_bytes, need);
    return ANX_OK;
}

}   // namespace

// ========================================================================= C ABI
extern "C" {

int32_t anx_version(void) { return 100; }

const char *anx_status_string(anx_status s) {
    switch (s) {
    case ANX_OK: return "ok";
    case ANX_ERR_BAD_ARG: return "bad argument";
    case ANX_ERR_BAD_SHAPE: return "unsupported input shape";
    case ANX_ERR_UNSUPPORTED: return "configuration not supported by the engine";
    case ANX_ERR_WORKSPACE: return "workspace missing, misaligned or too small";
    case ANX_ERR_CUDA: return "CUDA error";
    case ANX_ERR_NOT_READY: return "parameters not set";
    case ANX_ERR_NO_DEVICE: return "no usable sm_100 device";
    }
    return "unknown status";
}

const char *anx_engine_last_error(const anx_engine *e) { return e ? e->last_error.c_str() : "null engine"; }

anx_status anx_engine_create(const anx_unet_desc *desc, anx_engine **out) {
    if (!desc || !out || desc->struct_size != sizeof(anx_unet_desc)) return ANX_ERR_BAD_ARG;
    *out = nullptr;
    if (desc->input_nc < 1 || desc->output_nc < 1 || desc->output_nc > 256 || desc->num_downs < 1 ||
        desc->num_downs > 7 || desc->ngf < 16 || desc->ngf % 16 != 0 || (desc->ngf << desc->num_downs) > 1024)
        return ANX_ERR_UNSUPPORTED;
    for (int i = 0; i <= desc->num_downs; ++i) {   // widths above 256 are split into 256-channel CTA slices
        const int wdt = desc->ngf << i;
        if (wdt > 256 && wdt % 256 != 0) return ANX_ERR_UNSUPPORTED;
    }
    if (desc->norm_kind < ANX_NORM_NONE || desc->norm_kind > ANX_NORM_INSTANCE) return ANX_ERR_BAD_ARG;
    if (desc->act_kind < ANX_ACT_NONE || desc->act_kind > ANX_ACT_LEAKY) return ANX_ERR_BAD_ARG;
    if (desc->pool_kind != ANX_POOL_MAX && desc->pool_kind != ANX_POOL_AVG) return ANX_ERR_BAD_ARG;
    if (desc->interp_kind != ANX_INTERP_NEAREST && desc->interp_kind != ANX_INTERP_TRILINEAR) return ANX_ERR_BAD_ARG;
    if (desc->ngf > 64 || desc->input_nc > 4) return ANX_ERR_UNSUPPORTED;
    if ((desc->flags & ANX_FLAG_STORE_FP16) && (desc->flags & ANX_FLAG_STORE_BF16)) return ANX_ERR_BAD_ARG;

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || desc->device < 0 || desc->device >= ndev) return ANX_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, desc->device) != cudaSuccess) return ANX_ERR_NO_DEVICE;
    if (prop.major != 10) return ANX_ERR_NO_DEVICE;   // tcgen05 kernels: sm_100a only, no fallback
    if (cudaSetDevice(desc->device) != cudaSuccess) return ANX_ERR_NO_DEVICE;

    anx_engine *e = new anx_engine();
    e->desc = *desc;
    // InstanceNorm bounds every stored activation, so fp16 (3 more mantissa bits than bf16) is safe and
    // is needed to stay within tolerance through 24 normalised layers; BatchNorm networks keep bf16 range.
    e->dt = desc->norm_kind == ANX_NORM_INSTANCE ? DT_FP16 : DT_BF16;
    if (desc->flags & ANX_FLAG_STORE_FP16) e->dt = DT_FP16;
    if (desc->flags & ANX_FLAG_STORE_BF16) e->dt = DT_BF16;
    e->num_sms = prop.multiProcessorCount;
    e->max_smem = (int)prop.sharedMemPerBlockOptin;
    if (const char *xl = exp_env("ANX_X_LEAD")) e->x_lead = atoi(xl);
    if (desc->flags & ANX_FLAG_NO_ROWS) e->use_rows = 0;
    e->rows_rpw1 = ROWS_RPW1_DEFAULT;
    if (const char *rp = exp_env("ANX_RPW1")) e->rows_rpw1 = atoi(rp);
    if (const char *rw = exp_env("ANX_ROWS")) e->use_rows = atoi(rw);
    build_program(e);
    build_taps(e);
    build_liveness(e);
    for (auto &c : e->convs) {
        c.fold = (!c.is_stem && c.n_splits == 1 && 3 * c.ncols <= 256 && !exp_env("ANX_NOFOLD")) ? 1 : 0;
        c.groups = c.fold ? 1 : 3;
    }
    cudaError_t err = cudaSuccess;
#define ANX_SMEM(K_) if (err == cudaSuccess) err = cudaFuncSetAttribute(K_, cudaFuncAttributeMaxDynamicSharedMemorySize, e->max_smem)
    ANX_SMEM(conv3_umma_kernel<EPI_PADDED>); ANX_SMEM(conv3_umma_kernel<EPI_POOL>); ANX_SMEM(conv3_umma_kernel<EPI_D2S>);
    ANX_SMEM(conv3_umma_kernel<EPI_STATS>); ANX_SMEM(conv3_umma_kernel<EPI_F32>); ANX_SMEM(conv3_umma_kernel<EPI_F32_PEERS>);
    ANX_SMEM(conv3_umma_kernel<EPI_SEEDED>); ANX_SMEM(conv3_umma_kernel<EPI_F32_HEAD>); ANX_SMEM(conv3_umma_kernel<EPI_CL16>);
    ANX_SMEM(conv3_rows_kernel<EPI_CL16>); ANX_SMEM((conv3_rows_kernel<EPI_PADDED, true>));
    ANX_SMEM(conv3_rows_kernel<EPI_PADDED>); ANX_SMEM(conv3_rows_kernel<EPI_F32>); ANX_SMEM(conv3_rows_kernel<EPI_F32_HEAD>);
    ANX_SMEM(conv3_rows_kernel<EPI_SEEDED>); ANX_SMEM(conv3_rows_kernel<EPI_POOL>);
    ANX_SMEM((stem_umma_kernel<1, EPI_PADDED>)); ANX_SMEM((stem_umma_kernel<2, EPI_PADDED>)); ANX_SMEM((stem_umma_kernel<3, EPI_PADDED>));
    ANX_SMEM((stem_umma_kernel<1, EPI_STATS>)); ANX_SMEM((stem_umma_kernel<2, EPI_STATS>)); ANX_SMEM((stem_umma_kernel<3, EPI_STATS>));
    ANX_SMEM((stem_umma_kernel<1, EPI_F32>)); ANX_SMEM((stem_umma_kernel<2, EPI_F32>)); ANX_SMEM((stem_umma_kernel<3, EPI_F32>));
    ANX_SMEM((conv3_rows_kernel<EPI_F32, true>));
    ANX_SMEM((conv3_rows_kernel<EPI_PADDED, false, 1>)); ANX_SMEM((conv3_rows_kernel<EPI_PADDED, true, 1>));
    ANX_SMEM((conv3_rows_kernel<EPI_SEEDED, false, 1>)); ANX_SMEM((conv3_rows_kernel<EPI_F32, false, 1>));
    ANX_SMEM((conv3_rows_kernel<EPI_CL16, false, 1>)); ANX_SMEM((conv3_rows_kernel<EPI_F32_HEAD, false, 1>));
#undef ANX_SMEM
    if (err == cudaSuccess)
        err = cudaFuncSetAttribute(stem_conv_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (err != cudaSuccess) {
        delete e;
        return ANX_ERR_CUDA;
    }
    *out = e;
    return ANX_OK;
}

void anx_engine_destroy(anx_engine *e) {
    if (!e) return;
    if (e->h2d_stream) cudaStreamDestroy(e->h2d_stream);
    if (e->d2h_stream) cudaStreamDestroy(e->d2h_stream);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    for (int i = 0; i < 2; ++i)
        if (e->ev_pipe_done[i]) cudaEventDestroy(e->ev_pipe_done[i]);
    for (int i = 0; i < 8; ++i) {
        if (e->ev_in[i]) cudaEventDestroy(e->ev_in[i]);
        if (e->ev_out[i]) cudaEventDestroy(e->ev_out[i]);
    }
    if (e->d_head) cudaFree(e->d_head);
    if (e->d_ticket) cudaFree(e->d_ticket);
    if (e->ev_push_fork) cudaEventDestroy(e->ev_push_fork);
    for (int i = 0; i < 8; ++i) {
        if (e->push_stream[i]) cudaStreamDestroy(e->push_stream[i]);
        if (e->ev_push_done[i]) cudaEventDestroy(e->ev_push_done[i]);
    }
    for (auto &c : e->tap_convs) {
        if (c.d_wpack) cudaFree(c.d_wpack);
        if (c.d_wstem) cudaFree(c.d_wstem);
        if (c.d_wstem_rows) cudaFree(c.d_wstem_rows);
        if (c.d_wtrim) cudaFree(c.d_wtrim);
        if (c.d_wpack_alt) cudaFree(c.d_wpack_alt);
        if (c.d_wrows) cudaFree(c.d_wrows);
        if (c.d_bias) cudaFree(c.d_bias);
    }
    for (auto &c : e->convs) {
        if (c.d_wpack) cudaFree(c.d_wpack);
        if (c.d_wstem) cudaFree(c.d_wstem);
        if (c.d_wstem_rows) cudaFree(c.d_wstem_rows);
        if (c.d_wtrim) cudaFree(c.d_wtrim);
        if (c.d_wpack_alt) cudaFree(c.d_wpack_alt);
        if (c.d_wrows) cudaFree(c.d_wrows);
        if (c.d_bias) cudaFree(c.d_bias);
    }
    delete e;
}

int32_t anx_engine_num_convs(const anx_engine *e) { return e ? (int32_t)e->logical.size() : -1; }

anx_status anx_engine_conv_info(const anx_engine *e, int32_t k, int32_t *module_index, int32_t *cin, int32_t *cout,
                                int32_t *has_norm) {
    if (!e || k < 0 || k >= (int)e->logical.size()) return ANX_ERR_BAD_ARG;
    const Logical &c = e->logical[k];
    if (module_index) *module_index = c.module_index;
    if (cin) *cin = c.cin;
    if (cout) *cout = c.cout;
    if (has_norm) *has_norm = c.has_norm ? 1 : 0;
    return ANX_OK;
}

// Packs and uploads one launch's parameters: `w` = [c.cout][c.cin][27] fp32 with every fold already
// applied, `shift` = [c.ncols] accumulator seeds.
static anx_status upload_conv(anx_engine *e, ConvLayer &c, const std::vector<float> &w, const std::vector<float> &shift) {
    if (c.d_wpack) { cudaFree(c.d_wpack); c.d_wpack = nullptr; }
    if (c.d_bias) { cudaFree(c.d_bias); c.d_bias = nullptr; }
    c.ready = false;
    if (c.is_stem) {
        // fp32 [cin][27][ncols] for the CUDA-core stem
        std::vector<float> pk((size_t)c.cin * 27 * c.ncols, 0.0f);
        for (int o = 0; o < c.cout; ++o)
            for (int i = 0; i < c.cin; ++i)
                for (int t = 0; t < 27; ++t) pk[((size_t)i * 27 + t) * c.ncols + o] = w[((size_t)o * c.cin + i) * 27 + t];
        c.wpack_bytes = pk.size() * sizeof(float);
        ANX_CUDA(e, cudaMalloc(&c.d_wpack, c.wpack_bytes));
        ANX_CUDA(e, cudaMemcpy(c.d_wpack, pk.data(), c.wpack_bytes, cudaMemcpyHostToDevice));
        if (9 * c.cin <= 16 * STEM_MAX_KQ && 3 * c.ncols <= 256) {
            // tensor-core stem: K index = ci*9 + (ky*3+kx), rows = (dz=+1 | 0 | -1) x ncols, w = hi + lo in bf16
            const int kq = (9 * c.cin + 15) / 16, R = 3 * c.ncols;
            const size_t half_img = (size_t)kq * 2 * R * 8;
            std::vector<uint16_t> img(2 * half_img, 0);
            for (int o = 0; o < c.cout; ++o)
                for (int i = 0; i < c.cin; ++i)
                    for (int kz = 0; kz < 3; ++kz)
                        for (int t = 0; t < 9; ++t) {
                            const float v = w[((size_t)o * c.cin + i) * 27 + kz * 9 + t];
                            const uint16_t hi = f32_to_bf16_rne(v);
                            uint32_t hb32 = (uint32_t)hi << 16;
                            float hf;
                            std::memcpy(&hf, &hb32, 4);
                            const uint16_t lo = f32_to_bf16_rne(v - hf);
                            const int k = i * 9 + t, row = (2 - kz) * c.ncols + o;
                            const size_t idx = ((size_t)((k / 16) * 2 + (k % 16) / 8) * R + row) * 8 + k % 8;
                            img[idx] = hi;
                            img[half_img + idx] = lo;
                        }
            if (c.d_wstem) { cudaFree(c.d_wstem); c.d_wstem = nullptr; }
            ANX_CUDA(e, cudaMalloc(&c.d_wstem, img.size() * 2));
            ANX_CUDA(e, cudaMemcpy(c.d_wstem, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
        }
        if (c.d_wstem_rows) { cudaFree(c.d_wstem_rows); c.d_wstem_rows = nullptr; }
        if (c.cin == 1 && c.ncols == 16) {
            // row-form stem (conv3_rows_kernel<.., true>): image r (= input plane mod 3), row (j*3 + s)*16 + co holds
            // the taps ky = 2 - j, kz = (r - s) mod 3 of output channel co; K = { w_hi(dx 0..2), w_hi(dx 0..2),
            // w_lo(dx 0..2), 0 x 7 } against A = { x_hi, x_lo, x_hi } of the three dx neighbours
            std::vector<uint16_t> img((size_t)3 * ROWS_STEM_B_IMAGE_BYTES / 2, 0);
            for (int r = 0; r < 3; ++r)
                for (int j = 0; j < 3; ++j)
                    for (int s = 0; s < 3; ++s) {
                        const int ky = 2 - j, kz = (r - s + 3) % 3;
                        for (int o = 0; o < c.cout; ++o)
                            for (int dx = 0; dx < 3; ++dx) {
                                const float v = w[(size_t)o * 27 + kz * 9 + ky * 3 + dx];
                                const uint16_t hi = f32_to_bf16_rne(v);
                                uint32_t hb32 = (uint32_t)hi << 16;
                                float hf;
                                std::memcpy(&hf, &hb32, 4);
                                const uint16_t lo = f32_to_bf16_rne(v - hf);
                                const size_t row = (size_t)(j * 3 + s) * 16 + o;
                                const uint16_t vals[3] = {hi, hi, lo};
                                for (int part = 0; part < 3; ++part) {
                                    const int k = part * 3 + dx;
                                    img[(size_t)r * (ROWS_STEM_B_IMAGE_BYTES / 2) + ((size_t)(k / 8) * ROWS_N + row) * 8 + k % 8] = vals[part];
                                }
                            }
                    }
            ANX_CUDA(e, cudaMalloc(&c.d_wstem_rows, img.size() * 2));
            ANX_CUDA(e, cudaMemcpy(c.d_wstem_rows, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
        }
    } else {
        // 16-bit slabs [split][chunk][group][tap(ky,kx)][kchunk][row][8]; folded rows = (dz=+1 | 0 | -1) x ncols
        const int W = c.ncols_split;
        const int R = c.fold ? 3 * W : W;
        const int chunks = c.cin / 16;
        const size_t slab = (size_t)9 * 2 * R * 8;
        std::vector<uint16_t> pk((size_t)c.n_splits * chunks * c.groups * slab, 0);
        for (int o = 0; o < c.cout; ++o)
            for (int i = 0; i < c.cin; ++i) {
                const int ch = i / 16, kc = (i % 16) / 8, el = i % 8;
                const int split = o / W, ol = o % W;
                for (int kz = 0; kz < 3; ++kz) {
                    const int grp = c.fold ? 0 : kz;
                    const int row = c.fold ? (2 - kz) * W + ol : ol;
                    for (int t = 0; t < 9; ++t) {
                        const float v = w[((size_t)o * c.cin + i) * 27 + kz * 9 + t];
                        pk[((size_t)(split * chunks + ch) * c.groups + grp) * slab + ((size_t)(t * 2 + kc) * R + row) * 8 + el] =
                            e->dt == DT_BF16 ? f32_to_bf16_rne(v) : f32_to_f16_rne(v);
                    }
                }
            }
        c.wpack_bytes = pk.size() * sizeof(uint16_t);
        ANX_CUDA(e, cudaMalloc(&c.d_wpack, c.wpack_bytes));
        ANX_CUDA(e, cudaMemcpy(c.d_wpack, pk.data(), c.wpack_bytes, cudaMemcpyHostToDevice));
        if (c.d_wpack_alt) { cudaFree(c.d_wpack_alt); c.d_wpack_alt = nullptr; }
        if (c.alt_splits > 0 && !c.d2s_cout) {
            // second packing for small problems (make_geom): 64-column splits, unfolded, one slab per dz
            const int Wa = 64, nsp = c.alt_splits;
            const size_t slab_a = (size_t)9 * 2 * Wa * 8;
            std::vector<uint16_t> pa((size_t)nsp * chunks * 3 * slab_a, 0);
            for (int o = 0; o < c.cout; ++o)
                for (int i = 0; i < c.cin; ++i) {
                    const int ch = i / 16, kc = (i % 16) / 8, el = i % 8, split = o / Wa, ol = o % Wa;
                    for (int kz = 0; kz < 3; ++kz)
                        for (int t = 0; t < 9; ++t) {
                            const float v = w[((size_t)o * c.cin + i) * 27 + kz * 9 + t];
                            pa[((size_t)(split * chunks + ch) * 3 + kz) * slab_a + ((size_t)(t * 2 + kc) * Wa + ol) * 8 + el] =
                                e->dt == DT_BF16 ? f32_to_bf16_rne(v) : f32_to_f16_rne(v);
                        }
                }
            ANX_CUDA(e, cudaMalloc(&c.d_wpack_alt, pa.size() * sizeof(uint16_t)));
            ANX_CUDA(e, cudaMemcpy(c.d_wpack_alt, pa.data(), pa.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        }
        if (c.d_wrows) { cudaFree(c.d_wrows); c.d_wrows = nullptr; }
        if (c.fold && c.cin == 16 && c.ncols == 16 && c.n_splits == 1) {
            // conv3_rows_kernel: image r (= input plane mod 3), tap dx, K half, row (j*3 + s)*16 + co holds
            // w[co][ci][kz][ky][dx] with ky = 2 - j (output row i - ky) and kz = (r - s) mod 3 (z slot s)
            std::vector<uint16_t> img((size_t)3 * ROWS_B_IMAGE_BYTES / 2, 0);
            for (int r = 0; r < 3; ++r)
                for (int dx = 0; dx < 3; ++dx)
                    for (int j = 0; j < 3; ++j)
                        for (int s = 0; s < 3; ++s) {
                            const int ky = 2 - j, kz = (r - s + 3) % 3;
                            for (int o = 0; o < c.cout; ++o)
                                for (int i = 0; i < 16; ++i) {
                                    const float v = w[((size_t)o * c.cin + i) * 27 + kz * 9 + ky * 3 + dx];
                                    const size_t row = (size_t)(j * 3 + s) * 16 + o;
                                    const size_t idx = (size_t)r * (ROWS_B_IMAGE_BYTES / 2) + (size_t)dx * (ROWS_B_TAP_BYTES / 2) +
                                                       ((size_t)(i / 8) * ROWS_N + row) * 8 + i % 8;
                                    img[idx] = e->dt == DT_BF16 ? f32_to_bf16_rne(v) : f32_to_f16_rne(v);
                                }
                        }
            ANX_CUDA(e, cudaMalloc(&c.d_wrows, img.size() * 2));
            ANX_CUDA(e, cudaMemcpy(c.d_wrows, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
        }
    }
    if (c.d_wtrim) { cudaFree(c.d_wtrim); c.d_wtrim = nullptr; c.wtrim_bytes = 0; }
    if (!c.is_stem && c.d2s_cout > 0 && !c.fold && c.n_splits == 1 && c.d2s_cout % 16 == 0) {
        // compact tiles [chunk][kz][tap(ky,kx)] -> [k half][rows of the tap's column span][8] (make_geom: trim == 2)
        const int chunks = c.cin / 16, blocks = c.d2s_cout / 16;
        std::vector<uint16_t> img;
        for (int ch = 0; ch < chunks; ++ch)
            for (int t = 0; t < 27; ++t) {
                int lo, n;
                trim_span(t / 9, (t % 9) / 3, t % 3, blocks, lo, n);
                const size_t at = img.size();
                img.resize(at + (size_t)2 * 16 * n * 8, 0);
                for (int kc = 0; kc < 2; ++kc)
                    for (int r = 0; r < 16 * n; ++r)
                        for (int el = 0; el < 8; ++el) {
                            const int o = 16 * lo + r, i = ch * 16 + kc * 8 + el;
                            const float v = o < c.cout ? w[((size_t)o * c.cin + i) * 27 + t] : 0.0f;
                            img[at + ((size_t)kc * 16 * n + r) * 8 + el] = e->dt == DT_BF16 ? f32_to_bf16_rne(v) : f32_to_f16_rne(v);
                        }
            }
        c.wtrim_bytes = img.size() * 2;
        ANX_CUDA(e, cudaMalloc(&c.d_wtrim, c.wtrim_bytes));
        ANX_CUDA(e, cudaMemcpy(c.d_wtrim, img.data(), c.wtrim_bytes, cudaMemcpyHostToDevice));
    }
    ANX_CUDA(e, cudaMalloc(&c.d_bias, c.ncols * sizeof(float)));
    ANX_CUDA(e, cudaMemcpy(c.d_bias, shift.data(), c.ncols * sizeof(float), cudaMemcpyHostToDevice));
    c.ready = true;
    return ANX_OK;
}

anx_status anx_engine_set_conv(anx_engine *e, int32_t k, const float *weight, const float *bias,
                               const float *bn_w, const float *bn_b, const float *bn_mean, const float *bn_var,
                               int32_t location) {
    if (!e) return ANX_ERR_BAD_ARG;
    if (k < 0 || k >= (int)e->logical.size() || !weight) return e->fail(ANX_ERR_BAD_ARG, "bad conv ordinal or null weight");
    const Logical &L = e->logical[k];
    ConvLayer &ca = e->convs[L.conv_a];
    const bool fold_bn = L.has_norm && e->desc.norm_kind == ANX_NORM_BATCH_EVAL;
    const bool inorm = L.has_norm && e->desc.norm_kind == ANX_NORM_INSTANCE;
    if (fold_bn && (!bn_w || !bn_b || !bn_mean || !bn_var))
        return e->fail(ANX_ERR_BAD_ARG, "conv %d is followed by BatchNorm: its four arrays are required", L.module_index);
    ANX_CUDA(e, cudaSetDevice(e->desc.device));
    const size_t nw = (size_t)L.cout * L.cin * 27;
    std::vector<float> hw(nw), hb, g1, g2, g3, g4;
    auto fetch = [&](const float *src, std::vector<float> &dst, size_t n) -> cudaError_t {
        dst.resize(n);
        if (location == ANX_LOC_DEVICE) return cudaMemcpy(dst.data(), src, n * sizeof(float), cudaMemcpyDeviceToHost);
        std::memcpy(dst.data(), src, n * sizeof(float));
        return cudaSuccess;
    };
    ANX_CUDA(e, fetch(weight, hw, nw));
    if (bias) ANX_CUDA(e, fetch(bias, hb, L.cout));
    const int ncols = (L.cout + 15) / 16 * 16;
    std::vector<float> scale(L.cout, 1.0f), shift(ncols, 0.0f);
    if (fold_bn) {
        ANX_CUDA(e, fetch(bn_w, g1, L.cout));
        ANX_CUDA(e, fetch(bn_b, g2, L.cout));
        ANX_CUDA(e, fetch(bn_mean, g3, L.cout));
        ANX_CUDA(e, fetch(bn_var, g4, L.cout));
        // eval BatchNorm folded into the conv: y = (conv + b - mean) * gamma / sqrt(var + eps) + beta
        for (int o = 0; o < L.cout; ++o) {
            scale[o] = g1[o] / std::sqrt(g4[o] + e->desc.norm_eps);
            shift[o] = g2[o] - g3[o] * scale[o] + (bias ? hb[o] * scale[o] : 0.0f);
        }
    } else if (bias && !inorm) {   // a bias in front of InstanceNorm is cancelled by the mean subtraction
        for (int o = 0; o < L.cout; ++o) shift[o] = hb[o];
    }
    for (int o = 0; o < L.cout; ++o)
        for (size_t j = 0; j < (size_t)L.cin * 27; ++j) hw[(size_t)o * L.cin * 27 + j] *= scale[o];

    if (L.conv_b < 0) return upload_conv(e, ca, hw, shift);

    // Decoder conv run as two launches (see build_program).  Input channels [0, w) are the skip tensor,
    // [w, cin) the upsampled one.
    ConvLayer &cb = e->convs[L.conv_b];
    const int w = L.cout, cup = L.cin - w;
    // low-resolution half: column (p*w + co), p = a*4 + b*2 + c; tap (kz,ky,kx) of the high-resolution kernel
    // lands on low-resolution offset floor((parity + k - 1) / 2) along each axis
    std::vector<float> wa((size_t)8 * w * cup * 27, 0.0f), sa((size_t)8 * w, 0.0f);
    auto lowtap = [](int par, int kk) { const int v = par + kk - 1; return (v < 0 ? -1 : v / 2) + 1; };
    for (int par = 0; par < 8; ++par) {
        const int pa = (par >> 2) & 1, pb = (par >> 1) & 1, pc = par & 1;
        for (int co = 0; co < w; ++co) {
            sa[(size_t)par * w + co] = shift[co];
            for (int ci = 0; ci < cup; ++ci)
                for (int kz = 0; kz < 3; ++kz)
                    for (int ky = 0; ky < 3; ++ky)
                        for (int kx = 0; kx < 3; ++kx) {
                            const int t = (lowtap(pa, kz) * 3 + lowtap(pb, ky)) * 3 + lowtap(pc, kx);
                            wa[(((size_t)par * w + co) * cup + ci) * 27 + t] +=
                                hw[((size_t)co * L.cin + w + ci) * 27 + (kz * 3 + ky) * 3 + kx];
                        }
        }
    }
    anx_status st = upload_conv(e, ca, wa, sa);
    if (st != ANX_OK) return st;
    // skip half: the first w input channels; its accumulators are seeded from the partial sums
    std::vector<float> wb((size_t)w * w * 27), sb(cb.ncols, 0.0f);
    for (int co = 0; co < w; ++co)
        for (int ci = 0; ci < w; ++ci)
            for (int t = 0; t < 27; ++t) wb[((size_t)co * w + ci) * 27 + t] = hw[((size_t)co * L.cin + ci) * 27 + t];
    return upload_conv(e, cb, wb, sb);
}

size_t anx_engine_workspace_bytes(const anx_engine *e, int32_t n, int32_t d, int32_t h, int32_t w) {
    if (!e || !shape_ok(e, n, d, h, w)) return 0;
    std::vector<size_t> offs;
    size_t total = plan_offsets(e, n, d, h, w, offs);
    for (auto &c : e->convs)
        if (c.inorm) total += align_up((size_t)n * c.ncols * 2 * sizeof(double), 256);
    return total;
}

anx_status anx_engine_buffer_info(const anx_engine *e, int32_t n, int32_t d, int32_t h, int32_t w, int32_t index,
                                  size_t *offset, size_t *bytes, int32_t *level, int32_t *groups) {
    if (!e || !shape_ok(e, n, d, h, w) || index < 0 || index >= (int)e->bufs.size()) return ANX_ERR_BAD_ARG;
    std::vector<size_t> offs;
    plan_offsets(e, n, d, h, w, offs);
    if (offset) *offset = offs[index];
    if (bytes) *bytes = buffer_bytes(e, e->bufs[index], n, d, h, w);
    if (level) *level = e->bufs[index].level;
    if (groups) *groups = e->bufs[index].groups;
    return ANX_OK;
}

anx_status anx_engine_row_layout(const anx_engine *e, int32_t w, int32_t *lead, int32_t *pitch) {
    if (!e || w < 1) return ANX_ERR_BAD_ARG;
    const RowLayout rl = layout_of(w, e->x_lead);
    if (lead) *lead = rl.lead;
    if (pitch) *pitch = rl.pitch;
    return ANX_OK;
}

int32_t anx_engine_num_buffers(const anx_engine *e) { return e ? (int32_t)e->bufs.size() : -1; }

int32_t anx_engine_launches_per_forward(const anx_engine *e, int32_t n, int32_t d, int32_t h, int32_t w) {
    if (!e || !shape_ok(e, n, d, h, w)) return -1;
    int32_t count = 0;
    for (auto &s : e->steps) {
        if (s.kind == STEP_POOL && !(e->desc.flags & ANX_FLAG_FORCE_SIMT)) {
            const ConvLayer &c = e->convs[s.conv];
            const int bz = std::min(std::max(1, std::min(8, 256 / c.ncols_split)), d >> c.level);
            if (c.pool_dst_buf >= 0 && bz % 4 == 0) continue;   // fused into the producing conv
        }
        ++count;
    }
    return count;
}

anx_status anx_engine_forward(anx_engine *e, const float *in, float *out, int32_t n, int32_t d, int32_t h, int32_t w,
                              void *workspace, size_t ws_bytes, void *stream) {
    anx_status st = check_forward_args(e, in, out, n, d, h, w, workspace, ws_bytes);
    if (st != ANX_OK) return st;
    ANX_CUDA(e, cudaSetDevice(e->desc.device));
    std::shared_ptr<ShapePlan> p;
    st = get_plan(e, n, d, h, w, workspace, p);
    if (st != ANX_OK) return st;
    if (p->stats_bytes)
        ANX_CUDA(e, cudaMemsetAsync(static_cast<char *>(workspace) + p->stats_offset, 0, p->stats_bytes,
                                    static_cast<cudaStream_t>(stream)));
    for (auto &s : e->steps) {
        st = launch_step(e, *p, s, in, out, static_cast<cudaStream_t>(stream));
        if (st != ANX_OK) return st;
    }
    return ANX_OK;
}

int32_t anx_engine_num_steps(const anx_engine *e) { return e ? (int32_t)e->steps.size() : -1; }

anx_status anx_engine_step_info(const anx_engine *e, int32_t step, int32_t *kind, int32_t *out_buffer,
                                int32_t *out_group_offset, int32_t *out_groups, char name[32]) {
    if (!e || step < 0 || step >= (int)e->steps.size()) return ANX_ERR_BAD_ARG;
    const Step &s = e->steps[step];
    int buf = s.dst_buf, goff = s.dst_group_offset, groups = s.groups;
    if (s.kind == STEP_STEM || s.kind == STEP_CONV) {
        const ConvLayer &c = e->convs[s.conv];
        buf = c.dst_buf;
        goff = c.dst_group_offset;
        groups = (c.cout + 7) / 8;
    }
    if (kind) *kind = (int32_t)s.kind;
    if (out_buffer) *out_buffer = buf;
    if (out_group_offset) *out_group_offset = goff;
    if (out_groups) *out_groups = groups;
    if (name) std::memcpy(name, s.name, 32);
    return ANX_OK;
}

anx_status anx_engine_run_steps(anx_engine *e, const float *in, float *out, int32_t n, int32_t d, int32_t h,
                                int32_t w, void *workspace, size_t ws_bytes, void *stream, int32_t first,
                                int32_t last) {
    anx_status st = check_forward_args(e, in, out, n, d, h, w, workspace, ws_bytes);
    if (st != ANX_OK) return st;
    if (first < 0 || last > (int)e->steps.size() || first > last) return e->fail(ANX_ERR_BAD_ARG, "bad step range");
    ANX_CUDA(e, cudaSetDevice(e->desc.device));
    std::shared_ptr<ShapePlan> p;
    st = get_plan(e, n, d, h, w, workspace, p);
    if (st != ANX_OK) return st;
    if (first == 0 && p->stats_bytes)
        ANX_CUDA(e, cudaMemsetAsync(static_cast<char *>(workspace) + p->stats_offset, 0, p->stats_bytes,
                                    static_cast<cudaStream_t>(stream)));
    for (int i = first; i < last; ++i) {
        st = launch_step(e, *p, e->steps[i], in, out, static_cast<cudaStream_t>(stream));
        if (st != ANX_OK) return st;
    }
    return ANX_OK;
}

anx_status anx_engine_forward_gather(anx_engine *e, const float *in, void *const *out_peers, int32_t world,
                                     int32_t rank, int32_t payload, int32_t n, int32_t d, int32_t h, int32_t w,
                                     void *workspace, size_t ws_bytes, void *stream) {
    if (!e) return ANX_ERR_BAD_ARG;
    if (!out_peers || world < 1 || world > 8 || rank < 0 || rank >= world)
        return e->fail(ANX_ERR_BAD_ARG, "bad peer list (world %d, rank %d; at most 8 peers)", world, rank);
    if (payload != ANX_PAYLOAD_F32_NCDHW && payload != ANX_PAYLOAD_CL16) return e->fail(ANX_ERR_BAD_ARG, "bad payload kind");
    for (int i = 0; i < world; ++i)
        if (!out_peers[i]) return e->fail(ANX_ERR_BAD_ARG, "null gather buffer for peer %d", i);
    anx_status st = check_forward_args(e, in, out_peers[rank], n, d, h, w, workspace, ws_bytes);
    if (st != ANX_OK) return st;
    ANX_CUDA(e, cudaSetDevice(e->desc.device));
    std::shared_ptr<ShapePlan> p;
    st = get_plan(e, n, d, h, w, workspace, p);
    if (st != ANX_OK) return st;
    GatherArgs ga;
    for (int i = 0; i < world; ++i) ga.peers[i] = static_cast<float *>(out_peers[i]);
    ga.n_peers = world;
    ga.sample_offset = rank * n;
    ga.payload = payload;
    if (p->stats_bytes)
        ANX_CUDA(e, cudaMemsetAsync(static_cast<char *>(workspace) + p->stats_offset, 0, p->stats_bytes,
                                    static_cast<cudaStream_t>(stream)));
    for (auto &s : e->steps) {
        st = launch_step(e, *p, s, in, static_cast<float *>(out_peers[rank]), static_cast<cudaStream_t>(stream), &ga);
        if (st != ANX_OK) return st;
    }
    return ANX_OK;
}

anx_status anx_engine_forward_allgather(anx_engine *e, const float *in, float *const *out_peers, int32_t world,
                                        int32_t rank, int32_t n, int32_t d, int32_t h, int32_t w, void *workspace,
                                        size_t ws_bytes, void *stream) {
    return anx_engine_forward_gather(e, in, reinterpret_cast<void *const *>(out_peers), world, rank,
                                     ANX_PAYLOAD_F32_NCDHW, n, d, h, w, workspace, ws_bytes, stream);
}

anx_status anx_engine_forward_cl16(anx_engine *e, const float *in, void *out_cl16, int32_t n, int32_t d, int32_t h,
                                   int32_t w, void *workspace, size_t ws_bytes, void *stream) {
    void *one[1] = {out_cl16};
    return anx_engine_forward_gather(e, in, one, 1, 0, ANX_PAYLOAD_CL16, n, d, h, w, workspace, ws_bytes, stream);
}

int32_t anx_engine_storage_type(const anx_engine *e) { return e ? e->dt : -1; }

anx_status anx_engine_forward_concat(anx_engine *e, const float *in, float *dst, int32_t dst_channels,
                                     int32_t channel_offset, int32_t n, int32_t d, int32_t h, int32_t w,
                                     void *workspace, size_t ws_bytes, void *stream) {
    if (!e) return ANX_ERR_BAD_ARG;
    const int32_t mine = anx_engine_out_channels(e);
    if (!dst || channel_offset < 0 || dst_channels < channel_offset + mine)
        return e->fail(ANX_ERR_BAD_ARG, "channels [%d, %d) do not fit a tensor of %d channels", channel_offset,
                       channel_offset + mine, dst_channels);
    const size_t vol = (size_t)d * h * w;
    float *out = dst + (size_t)channel_offset * vol;
    anx_status st = check_forward_args(e, in, out, n, d, h, w, workspace, ws_bytes);
    if (st != ANX_OK) return st;
    ANX_CUDA(e, cudaSetDevice(e->desc.device));
    std::shared_ptr<ShapePlan> p;
    st = get_plan(e, n, d, h, w, workspace, p);
    if (st != ANX_OK) return st;
    GatherArgs ga;
    ga.out_nstride = (size_t)dst_channels * vol;
    if (p->stats_bytes)
        ANX_CUDA(e, cudaMemsetAsync(static_cast<char *>(workspace) + p->stats_offset, 0, p->stats_bytes,
                                    static_cast<cudaStream_t>(stream)));
    for (auto &s : e->steps) {
        st = launch_step(e, *p, s, in, out, static_cast<cudaStream_t>(stream), &ga);
        if (st != ANX_OK) return st;
    }
    return ANX_OK;
}

anx_status anx_channel_normalize_f32(const float *in, float *out, int64_t n, int32_t channels, int32_t d, int32_t h,
                                     int32_t w, int32_t mode, float eps, void *stream) {
    if (!in || !out || n < 1 || channels < 1 || d < 1 || h < 1 || w < 1 || (mode != 0 && mode != 1)) return ANX_ERR_BAD_ARG;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess) return ANX_ERR_NO_DEVICE;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t vol = (size_t)d * h * w;
    channel_normalize_kernel<<<grid_for((size_t)n * vol, 256, sms, 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, out, (size_t)n, vol, channels, mode, eps);
    return cudaGetLastError() == cudaSuccess ? ANX_OK : ANX_ERR_CUDA;
}

anx_status anx_widen_cl16_f32(const void *src_cl16, float *dst_ncdhw, int64_t n, int32_t channels, int32_t d,
                              int32_t h, int32_t w, int32_t storage_type, void *stream) {
    if (!src_cl16 || !dst_ncdhw || n < 1 || channels < 8 || (channels & 7) || d < 1 || h < 1 || w < 1 ||
        (storage_type != DT_BF16 && storage_type != DT_FP16))
        return ANX_ERR_BAD_ARG;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess) return ANX_ERR_NO_DEVICE;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t vol = (size_t)d * h * w;
    widen_cl16_kernel<<<grid_for((size_t)n * vol, 256, sms, 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4 *>(src_cl16), dst_ncdhw, (size_t)n, vol, channels, storage_type);
    return cudaGetLastError() == cudaSuccess ? ANX_OK : ANX_ERR_CUDA;
}

anx_status anx_push_to_peers(anx_engine *e, const void *src, void *const *peer_dst, int32_t world, int32_t rank,
                             size_t bytes, void *stream) {
    if (!e) return ANX_ERR_BAD_ARG;
    if (!src || !peer_dst || world < 1 || world > 8 || rank < 0 || rank >= world)
        return e->fail(ANX_ERR_BAD_ARG, "bad peer list (world %d, rank %d; at most 8 peers)", world, rank);
    ANX_CUDA(e, cudaSetDevice(e->desc.device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!e->ev_push_fork) {
        ANX_CUDA(e, cudaEventCreateWithFlags(&e->ev_push_fork, cudaEventDisableTiming));
        for (int i = 0; i < 8; ++i) {
            ANX_CUDA(e, cudaStreamCreateWithFlags(&e->push_stream[i], cudaStreamNonBlocking));
            ANX_CUDA(e, cudaEventCreateWithFlags(&e->ev_push_done[i], cudaEventDisableTiming));
        }
    }
    // one copy per destination, each on its own stream so the copy engines drive all NVLink ports at
    // once; `stream` continues when the last copy has landed
    ANX_CUDA(e, cudaEventRecord(e->ev_push_fork, st));
    for (int k = 1; k < world; ++k) {
        const int r = (rank + k) % world;           // staggered start: rank i begins with peer i + 1
        if (!peer_dst[r]) return e->fail(ANX_ERR_BAD_ARG, "null destination for peer %d", r);
        ANX_CUDA(e, cudaStreamWaitEvent(e->push_stream[r], e->ev_push_fork, 0));
        ANX_CUDA(e, cudaMemcpyAsync(peer_dst[r], src, bytes, cudaMemcpyDeviceToDevice, e->push_stream[r]));
        ANX_CUDA(e, cudaEventRecord(e->ev_push_done[r], e->push_stream[r]));
        ANX_CUDA(e, cudaStreamWaitEvent(st, e->ev_push_done[r], 0));
    }
    return ANX_OK;
}

anx_status anx_engine_forward_slab(anx_engine *e, const float *in, float *out, int32_t n, int32_t d, int32_t h,
                                   int32_t w, void *workspace, size_t ws_bytes, const anx_slab_links *links,
                                   void *stream) {
    anx_status st = check_forward_args(e, in, out, n, d, h, w, workspace, ws_bytes);
    if (st != ANX_OK) return st;
    if (!links || links->struct_size != sizeof(anx_slab_links) || !links->flags)
        return e->fail(ANX_ERR_BAD_ARG, "bad slab links");
    if ((links->lower_workspace != nullptr) != (e->slab_lower != 0) || (links->upper_workspace != nullptr) != (e->slab_upper != 0) ||
        (links->lower_workspace && !links->lower_flags) || (links->upper_workspace && !links->upper_flags))
        return e->fail(ANX_ERR_BAD_ARG, "slab links do not match anx_engine_set_slab (lower %d, upper %d)", e->slab_lower,
                       e->slab_upper);
    for (auto &c : e->convs)
        if (c.inorm)
            return e->fail(ANX_ERR_UNSUPPORTED, "InstanceNorm networks need whole-volume statistics between launches: "
                                                "use anx_engine_run_steps / anx_engine_step_stats");
    ANX_CUDA(e, cudaSetDevice(e->desc.device));
    if (!e->d_ticket) {
        ANX_CUDA(e, cudaMalloc(&e->d_ticket, sizeof(unsigned int)));
        ANX_CUDA(e, cudaMemset(e->d_ticket, 0, sizeof(unsigned int)));
    }
    std::shared_ptr<ShapePlan> p;
    st = get_plan(e, n, d, h, w, workspace, p);
    if (st != ANX_OK) return st;
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    // One launch per exchange: push my two boundary planes of the tensor into the neighbours' shells,
    // publish the sequence number, wait for theirs.  groups = 0: handshake only (start of a forward: the
    // neighbours have finished the previous one, so their shells may be overwritten).
    auto exchange = [&](int buf, int goff, int groups) -> anx_status {
        HaloArgs a{};
        size_t off_bytes = 0;
        if (groups > 0) {
            const ActView v = view_of(e, *p, buf, goff);
            a.plane = (size_t)(v.H + 2) * v.pitch;
            a.group_stride = (size_t)(v.D + 2) * a.plane;
            a.sample_stride = (size_t)e->bufs[buf].groups * a.group_stride;
            a.D = v.D;
            off_bytes = p->buf_offset[buf] + (size_t)goff * a.group_stride * 16;
        }
        a.src = reinterpret_cast<const uint4 *>(static_cast<char *>(workspace) + off_bytes);
        a.lower = links->lower_workspace ? reinterpret_cast<uint4 *>(static_cast<char *>(links->lower_workspace) + off_bytes) : nullptr;
        a.upper = links->upper_workspace ? reinterpret_cast<uint4 *>(static_cast<char *>(links->upper_workspace) + off_bytes) : nullptr;
        a.groups = groups;
        a.N = n;
        a.my_flags = links->flags;
        a.lower_flag = links->lower_workspace ? links->lower_flags + 1 : nullptr;
        a.upper_flag = links->upper_workspace ? links->upper_flags + 0 : nullptr;
        a.seq = ++e->slab_seq;
        a.ticket = e->d_ticket;
        const size_t per_dir = (size_t)n * groups * a.plane;
        const int grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)e->num_sms, (per_dir + 1023) / 1024));
        halo_exchange_kernel<<<grid, 256, 0, cs>>>(a);
        ANX_CUDA(e, cudaGetLastError());
        return ANX_OK;
    };
    st = exchange(0, 0, 0);
    if (st != ANX_OK) return st;
    for (auto &s : e->steps) {
        st = launch_step(e, *p, s, in, out, cs);
        if (st != ANX_OK) return st;
        int buf = -1, goff = 0, groups = 0;
        if (s.kind == STEP_STEM || s.kind == STEP_CONV) {
            const ConvLayer &c = e->convs[s.conv];
            if (!c.is_final && !c.d2s_cout) { buf = c.dst_buf; goff = c.dst_group_offset; groups = c.cout / 8; }
        } else if (s.kind == STEP_POOL || s.kind == STEP_UP) {
            buf = s.dst_buf; goff = s.dst_group_offset; groups = s.groups;
        }
        if (buf >= 0 && (e->slab_lower || e->slab_upper)) {
            st = exchange(buf, goff, groups);
            if (st != ANX_OK) return st;
        }
    }
    return ANX_OK;
}

anx_status anx_engine_forward_host(anx_engine *e, const float *in_host, float *out_host, int32_t n, int32_t d,
                                   int32_t h, int32_t w, float *dev_in, float *dev_out, void *workspace,
                                   size_t ws_bytes, void *stream) {
    return anx_engine_forward_host_ex(e, in_host, out_host, ANX_PAYLOAD_F32_NCDHW, n, d, h, w, dev_in, dev_out,
                                      workspace, ws_bytes, stream);
}

static anx_status forward_host_impl(anx_engine *e, const float *in_host, void *out_host_v, int32_t payload, int32_t n,
                                    int32_t d, int32_t h, int32_t w, float *dev_in, void *dev_out_v, void *workspace,
                                    size_t ws_bytes, void *stream, bool pipelined);

anx_status anx_engine_forward_host_ex(anx_engine *e, const float *in_host, void *out_host_v, int32_t payload,
                                      int32_t n, int32_t d, int32_t h, int32_t w, float *dev_in, void *dev_out_v,
                                      void *workspace, size_t ws_bytes, void *stream) {
    return forward_host_impl(e, in_host, out_host_v, payload, n, d, h, w, dev_in, dev_out_v, workspace, ws_bytes, stream, false);
}

anx_status anx_engine_forward_host_pipelined(anx_engine *e, const float *in_host, void *out_host_v, int32_t payload,
                                             int32_t n, int32_t d, int32_t h, int32_t w, float *dev_in,
                                             void *dev_out_v, void *workspace, size_t ws_bytes, void *stream) {
    return forward_host_impl(e, in_host, out_host_v, payload, n, d, h, w, dev_in, dev_out_v, workspace, ws_bytes, stream, true);
}

anx_status anx_engine_host_wait(anx_engine *e, void *stream) {
    if (!e) return ANX_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(e->host_mu);
    if (e->pipe_pending) {
        ANX_CUDA(e, cudaSetDevice(e->desc.device));
        ANX_CUDA(e, cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), e->ev_join, 0));
        e->pipe_pending = false;
    }
    return ANX_OK;
}

static anx_status forward_host_impl(anx_engine *e, const float *in_host, void *out_host_v, int32_t payload, int32_t n,
                                    int32_t d, int32_t h, int32_t w, float *dev_in, void *dev_out_v, void *workspace,
                                    size_t ws_bytes, void *stream, bool pipelined) {
    if (!e) return ANX_ERR_BAD_ARG;
    if (payload != ANX_PAYLOAD_F32_NCDHW && payload != ANX_PAYLOAD_CL16) return e->fail(ANX_ERR_BAD_ARG, "bad payload kind");
    char *out_host = static_cast<char *>(out_host_v), *dev_out = static_cast<char *>(dev_out_v);
    if (!in_host || !out_host || !dev_in || !dev_out) return e->fail(ANX_ERR_BAD_ARG, "null buffer");
    if (!shape_ok(e, n, d, h, w)) return e->fail(ANX_ERR_BAD_SHAPE, "bad shape for the host-buffer forward");
    std::lock_guard<std::mutex> lock(e->host_mu);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ANX_CUDA(e, cudaSetDevice(e->desc.device));
    if (!e->h2d_stream) {
        ANX_CUDA(e, cudaStreamCreateWithFlags(&e->h2d_stream, cudaStreamNonBlocking));
        ANX_CUDA(e, cudaStreamCreateWithFlags(&e->d2h_stream, cudaStreamNonBlocking));
        ANX_CUDA(e, cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
        ANX_CUDA(e, cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
        for (int i = 0; i < 8; ++i) {
            ANX_CUDA(e, cudaEventCreateWithFlags(&e->ev_in[i], cudaEventDisableTiming));
            ANX_CUDA(e, cudaEventCreateWithFlags(&e->ev_out[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < 2; ++i) ANX_CUDA(e, cudaEventCreateWithFlags(&e->ev_pipe_done[i], cudaEventDisableTiming));
    }
    // A download issued by an earlier pipelined call may still be reading this dev_out buffer: the convs of this call
    // (which overwrite it) wait for that download.  Two buffer sets used alternately never wait on each other.
    int slot = -1;
    for (int i = 0; i < 2; ++i)
        if (e->pipe_out[i] == dev_out_v) slot = i;
    if (slot >= 0) ANX_CUDA(e, cudaStreamWaitEvent(st, e->ev_pipe_done[slot], 0));
    // Three-stage pipeline over chunks of the batch: the upload of chunk i+1 and the download of
    // chunk i-1 run on their own streams while chunk i computes on the caller's stream.  The
    // download (output_nc/input_nc times larger than the upload) is what bounds the call.
    const int chunks = std::min(n, 4);
    // bytes per sample of the output in the chosen payload
    const size_t in_vol = (size_t)d * h * w * e->desc.input_nc,
                 out_vol = (size_t)d * h * w * anx_engine_out_channels(e) * (payload == ANX_PAYLOAD_CL16 ? 2 : 4);
    ANX_CUDA(e, cudaEventRecord(e->ev_fork, st));
    ANX_CUDA(e, cudaStreamWaitEvent(e->h2d_stream, e->ev_fork, 0));
    ANX_CUDA(e, cudaStreamWaitEvent(e->d2h_stream, e->ev_fork, 0));
    int lo = 0;
    for (int i = 0; i < chunks; ++i) {
        const int cnt = n / chunks + (i < n % chunks ? 1 : 0);
        ANX_CUDA(e, cudaMemcpyAsync(dev_in + lo * in_vol, in_host + lo * in_vol, cnt * in_vol * sizeof(float),
                                    cudaMemcpyHostToDevice, e->h2d_stream));
        ANX_CUDA(e, cudaEventRecord(e->ev_in[i], e->h2d_stream));
        ANX_CUDA(e, cudaStreamWaitEvent(st, e->ev_in[i], 0));
        anx_status r = payload == ANX_PAYLOAD_CL16
            ? anx_engine_forward_cl16(e, dev_in + lo * in_vol, dev_out + lo * out_vol, cnt, d, h, w, workspace, ws_bytes, stream)
            : anx_engine_forward(e, dev_in + lo * in_vol, reinterpret_cast<float *>(dev_out + lo * out_vol), cnt, d, h, w,
                                 workspace, ws_bytes, stream);
        if (r != ANX_OK) return r;
        ANX_CUDA(e, cudaEventRecord(e->ev_out[i], st));
        ANX_CUDA(e, cudaStreamWaitEvent(e->d2h_stream, e->ev_out[i], 0));
        ANX_CUDA(e, cudaMemcpyAsync(out_host + lo * out_vol, dev_out + lo * out_vol, cnt * out_vol,
                                    cudaMemcpyDeviceToHost, e->d2h_stream));
        lo += cnt;
    }
    ANX_CUDA(e, cudaEventRecord(e->ev_join, e->d2h_stream));
    if (pipelined) {
        // the caller's stream does NOT wait for the download: the next call's upload and convs overlap it;
        // anx_engine_host_wait joins.  Remember which download last read this dev_out buffer.
        if (slot < 0) { slot = e->pipe_next; e->pipe_next ^= 1; e->pipe_out[slot] = dev_out_v; }
        ANX_CUDA(e, cudaEventRecord(e->ev_pipe_done[slot], e->d2h_stream));
        e->pipe_pending = true;
    } else {
        ANX_CUDA(e, cudaStreamWaitEvent(st, e->ev_join, 0));   // the caller's stream completes after the last download
        e->pipe_pending = false;
    }
    return ANX_OK;
}

anx_status anx_engine_set_slab(anx_engine *e, int32_t has_lower, int32_t has_upper, int32_t depth_total) {
    if (!e) return ANX_ERR_BAD_ARG;
    if (depth_total < 0 || (depth_total % (1 << e->desc.num_downs)) != 0)
        return e->fail(ANX_ERR_BAD_ARG, "total depth %d is not a multiple of %d", depth_total, 1 << e->desc.num_downs);
    e->slab_lower = has_lower ? 1 : 0;
    e->slab_upper = has_upper ? 1 : 0;
    e->slab_depth_total = depth_total;
    return ANX_OK;
}

anx_status anx_engine_step_stats(const anx_engine *e, int32_t step, int32_t n, int32_t d, int32_t h, int32_t w,
                                 size_t *offset, size_t *bytes) {
    if (!e || step < 0 || step >= (int)e->steps.size() || !shape_ok(e, n, d, h, w)) return ANX_ERR_BAD_ARG;
    const Step &s = e->steps[step];
    if (offset) *offset = 0;
    if (bytes) *bytes = 0;
    if ((s.kind != STEP_STEM && s.kind != STEP_CONV) || !e->convs[s.conv].inorm) return ANX_OK;
    std::vector<size_t> offs;
    size_t off = plan_offsets(e, n, d, h, w, offs);
    for (int i = 0; i < s.conv; ++i)
        if (e->convs[i].inorm) off += align_up((size_t)n * e->convs[i].ncols * 2 * sizeof(double), 256);
    if (offset) *offset = off;
    if (bytes) *bytes = (size_t)n * e->convs[s.conv].ncols * 2 * sizeof(double);
    return ANX_OK;
}

int32_t anx_engine_out_channels(const anx_engine *e) {
    if (!e) return -1;
    return e->head_nc > 0 ? e->head_nc : e->desc.output_nc;
}

anx_status anx_engine_set_head(anx_engine *e, int32_t head_nc, const float *weight, const float *bias,
                               int32_t location) {
    if (!e) return ANX_ERR_BAD_ARG;
    if (head_nc == 0) {            // back to the plain network output
        e->head_nc = 0;
        return ANX_OK;
    }
    if (head_nc < 0 || head_nc > HEAD_MAX || !weight) return e->fail(ANX_ERR_BAD_ARG, "head of %d channels (1..%d)", head_nc, HEAD_MAX);
    if (e->desc.output_nc > 16)
        return e->fail(ANX_ERR_UNSUPPORTED, "a fused head needs output_nc <= 16 (one accumulator chunk per voxel)");
    if (e->desc.flags & ANX_FLAG_FORCE_SIMT) return e->fail(ANX_ERR_UNSUPPORTED, "no fused head on the CUDA-core debug path");
    ANX_CUDA(e, cudaSetDevice(e->desc.device));
    const int cin = e->desc.output_nc;
    std::vector<float> hw((size_t)head_nc * cin), hb(head_nc, 0.0f), img(HEAD_FLOATS, 0.0f);
    const cudaMemcpyKind kind = location == ANX_LOC_DEVICE ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost;
    ANX_CUDA(e, cudaMemcpy(hw.data(), weight, hw.size() * sizeof(float), kind));
    if (bias) ANX_CUDA(e, cudaMemcpy(hb.data(), bias, hb.size() * sizeof(float), kind));
    for (int k = 0; k < head_nc; ++k) {
        img[k] = hb[k];
        for (int c = 0; c < cin; ++c) img[HEAD_MAX + k * 16 + c] = hw[(size_t)k * cin + c];
    }
    if (!e->d_head) ANX_CUDA(e, cudaMalloc(&e->d_head, HEAD_FLOATS * sizeof(float)));
    ANX_CUDA(e, cudaMemcpy(e->d_head, img.data(), HEAD_FLOATS * sizeof(float), cudaMemcpyHostToDevice));
    e->head_nc = head_nc;
    return ANX_OK;
}

int32_t anx_engine_num_taps(const anx_engine *e) { return e ? (int32_t)e->taps.size() : -1; }

anx_status anx_engine_tap_info(const anx_engine *e, int32_t k, int32_t *module_index, int32_t *channels,
                               int32_t *level, int32_t *last_step, int32_t *is_output) {
    if (!e || k < 0 || k >= (int)e->taps.size()) return ANX_ERR_BAD_ARG;
    const TapSite &t = e->taps[k];
    if (module_index) *module_index = t.module_index;
    if (channels) *channels = t.channels;
    if (level) *level = t.level;
    if (last_step) *last_step = t.last_step;
    if (is_output) *is_output = t.buffer == -1 ? 1 : 0;
    return ANX_OK;
}

anx_status anx_engine_export_tap(anx_engine *e, int32_t k, int32_t n, int32_t d, int32_t h, int32_t w,
                                 void *workspace, size_t ws_bytes, float *out, void *stream) {
    if (!e) return ANX_ERR_BAD_ARG;
    if (k < 0 || k >= (int)e->taps.size() || !out) return e->fail(ANX_ERR_BAD_ARG, "bad tap ordinal or null output");
    const TapSite &t = e->taps[k];
    if (t.buffer < 0) return e->fail(ANX_ERR_BAD_ARG, "tap %d is not a stored tensor (network output or pre-norm tap)", k);
    if (!shape_ok(e, n, d, h, w)) return e->fail(ANX_ERR_BAD_SHAPE, "bad shape for a tap export");
    const size_t need = anx_engine_workspace_bytes(e, n, d, h, w);
    if (!workspace || ws_bytes < need) return e->fail(ANX_ERR_WORKSPACE, "workspace of %zu bytes, need %zu", ws_bytes, need);
    ANX_CUDA(e, cudaSetDevice(e->desc.device));
    std::shared_ptr<ShapePlan> p;
    anx_status st = get_plan(e, n, d, h, w, workspace, p);
    if (st != ANX_OK) return st;
    ActView src = view_of(e, *p, t.buffer, t.group_offset);
    const size_t items = (size_t)n * t.groups * src.D * src.H * src.W;
    export_ncdhw_kernel<<<grid_for(items, 256, e->num_sms, 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, n, t.groups, t.channels, out, e->dt);
    ANX_CUDA(e, cudaGetLastError());
    return ANX_OK;
}

anx_status anx_engine_tap_kind(const anx_engine *e, int32_t k, int32_t *kind, int32_t *conv_ordinal) {
    if (!e || k < 0 || k >= (int)e->taps.size()) return ANX_ERR_BAD_ARG;
    const TapSite &t = e->taps[k];
    if (kind) *kind = t.prenorm_conv >= 0 ? ANX_TAP_PRENORM : (t.buffer < 0 ? ANX_TAP_OUTPUT : ANX_TAP_STORED);
    if (conv_ordinal) *conv_ordinal = t.prenorm_conv;
    return ANX_OK;
}

anx_status anx_engine_set_tap_conv(anx_engine *e, int32_t k, const float *weight, const float *bias, int32_t location) {
    if (!e) return ANX_ERR_BAD_ARG;
    if (k < 0 || k >= (int)e->logical.size() || !weight) return e->fail(ANX_ERR_BAD_ARG, "bad conv ordinal or null weight");
    const Logical &L = e->logical[k];
    if (L.conv_b >= 0 || !L.has_norm)
        return e->fail(ANX_ERR_UNSUPPORTED, "conv %d has no pre-norm tap on this engine", L.module_index);
    ANX_CUDA(e, cudaSetDevice(e->desc.device));
    if (e->tap_convs.size() != e->logical.size()) e->tap_convs.resize(e->logical.size());
    ConvLayer &t = e->tap_convs[k];
    const ConvLayer &src = e->convs[L.conv_a];
    // the clone owns its device buffers: keep them across the struct copy (upload_conv frees / replaces them)
    void *keep[7] = {t.d_wpack, t.d_wstem, t.d_wrows, t.d_wtrim, t.d_wstem_rows, t.d_bias, t.d_wpack_alt};
    t = src;
    t.d_wpack = keep[0]; t.d_wstem = keep[1]; t.d_wrows = keep[2]; t.d_wtrim = keep[3]; t.d_wstem_rows = keep[4];
    t.d_bias = static_cast<float *>(keep[5]); t.d_wpack_alt = keep[6];
    t.is_tap = true; t.is_final = true; t.has_norm = false; t.has_act = false; t.inorm = false;
    t.pool_dst_buf = -1; t.seed_buf = -1; t.d2s_cout = 0; t.dst_buf = -1; t.ready = false;
    const size_t nw = (size_t)L.cout * L.cin * 27;
    std::vector<float> hw(nw), shift(t.ncols, 0.0f);
    const cudaMemcpyKind kind = location == ANX_LOC_DEVICE ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost;
    ANX_CUDA(e, cudaMemcpy(hw.data(), weight, nw * sizeof(float), kind));
    if (bias) ANX_CUDA(e, cudaMemcpy(shift.data(), bias, L.cout * sizeof(float), kind));
    return upload_conv(e, t, hw, shift);      // no fold: the clone computes conv(x) + bias
}

anx_status anx_engine_export_prenorm_tap(anx_engine *e, int32_t tap, const float *in, int32_t n, int32_t d, int32_t h,
                                         int32_t w, void *workspace, size_t ws_bytes, float *out, void *stream) {
    if (!e) return ANX_ERR_BAD_ARG;
    if (tap < 0 || tap >= (int)e->taps.size() || e->taps[tap].prenorm_conv < 0 || !out || !in)
        return e->fail(ANX_ERR_BAD_ARG, "tap %d is not a pre-norm tap (or null pointers)", tap);
    const TapSite &t = e->taps[tap];
    if ((int)e->tap_convs.size() <= t.prenorm_conv || !e->tap_convs[t.prenorm_conv].ready)
        return e->fail(ANX_ERR_NOT_READY, "anx_engine_set_tap_conv has not been called for conv ordinal %d", t.prenorm_conv);
    if (!shape_ok(e, n, d, h, w)) return e->fail(ANX_ERR_BAD_SHAPE, "bad shape for a tap export");
    const size_t need = anx_engine_workspace_bytes(e, n, d, h, w);
    if (!workspace || ws_bytes < need) return e->fail(ANX_ERR_WORKSPACE, "workspace of %zu bytes, need %zu", ws_bytes, need);
    ANX_CUDA(e, cudaSetDevice(e->desc.device));
    std::shared_ptr<ShapePlan> p;
    anx_status st = get_plan(e, n, d, h, w, workspace, p);
    if (st != ANX_OK) return st;
    // Re-run the conv's launch (its input tensor is still live: this is called right after step `last_step`) with
    // the un-folded clone and the fp32 NCDHW epilogue; tiles / tensor map are those of the original launch.
    const Step &s = e->steps[t.last_step];
    ConvGeom g = p->geoms[s.conv];
    g.fuse_pool = 0;
    return launch_step(e, *p, s, in, out, static_cast<cudaStream_t>(stream), nullptr, &e->tap_convs[t.prenorm_conv], &g);
}

anx_status anx_avgpool3d_scale_f32(const float *in, float *out, int64_t nc, int32_t d, int32_t h, int32_t w,
                                   int32_t k, float scale, void *stream) {
    if (!in || !out || nc < 1 || k < 1 || d < k || h < k || w < k) return ANX_ERR_BAD_ARG;
    const size_t items = (size_t)nc * (d / k) * (h / k) * (w / k);
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess) return ANX_ERR_NO_DEVICE;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (k == 2 && w % 4 == 0 && !(reinterpret_cast<uintptr_t>(in) & 15) && !(reinterpret_cast<uintptr_t>(out) & 7))
        avgpool3d_scale_k2_kernel<<<grid_for(items / 2, 256, sms, 32), 256, 0, st>>>(in, out, (size_t)nc, d, h, w, scale);
    else
        avgpool3d_scale_kernel<<<grid_for(items, 256, sms, 32), 256, 0, st>>>(in, out, (size_t)nc, d, h, w, k, scale);
    return cudaGetLastError() == cudaSuccess ? ANX_OK : ANX_ERR_CUDA;
}

anx_status anx_blend_window_f32(const float *pred, const float *weight, float *out, float *norm, int32_t channels,
                                int32_t d, int32_t h, int32_t w, int32_t D, int32_t H, int32_t W, int32_t z0,
                                int32_t y0, int32_t x0, void *stream) {
    if (!pred || !weight || !out || !norm || channels < 1 || d < 1 || h < 1 || w < 1) return ANX_ERR_BAD_ARG;
    if (z0 < 0 || y0 < 0 || x0 < 0 || z0 + d > D || y0 + h > H || x0 + w > W) return ANX_ERR_BAD_SHAPE;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess) return ANX_ERR_NO_DEVICE;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    blend_window_kernel<<<grid_for((size_t)d * h * w, 256, sms, 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        pred, weight, out, norm, channels, d, h, w, D, H, W, z0, y0, x0);
    return cudaGetLastError() == cudaSuccess ? ANX_OK : ANX_ERR_CUDA;
}

anx_status anx_engine_profile(anx_engine *e, const float *in, float *out, int32_t n, int32_t d, int32_t h, int32_t w,
                              void *workspace, size_t ws_bytes, void *stream, float *ms_out, char (*names)[32],
                              int32_t capacity, int32_t *count) {
    anx_status st = check_forward_args(e, in, out, n, d, h, w, workspace, ws_bytes);
    if (st != ANX_OK) return st;
    if (!ms_out || !names || !count) return e->fail(ANX_ERR_BAD_ARG, "null profile outputs");
    ANX_CUDA(e, cudaSetDevice(e->desc.device));
    std::shared_ptr<ShapePlan> p;
    st = get_plan(e, n, d, h, w, workspace, p);
    if (st != ANX_OK) return st;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int k = (int)e->steps.size();
    std::vector<cudaEvent_t> ev(k + 1);
    for (auto &x : ev) ANX_CUDA(e, cudaEventCreate(&x));
    if (p->stats_bytes)
        ANX_CUDA(e, cudaMemsetAsync(static_cast<char *>(workspace) + p->stats_offset, 0, p->stats_bytes, s));
    ANX_CUDA(e, cudaEventRecord(ev[0], s));
    for (int i = 0; i < k; ++i) {
        st = launch_step(e, *p, e->steps[i], in, out, s);
        if (st != ANX_OK) return st;
        ANX_CUDA(e, cudaEventRecord(ev[i + 1], s));
    }
    ANX_CUDA(e, cudaStreamSynchronize(s));
    *count = std::min(k, capacity);
    for (int i = 0; i < *count; ++i) {
        ANX_CUDA(e, cudaEventElapsedTime(&ms_out[i], ev[i], ev[i + 1]));
        std::memcpy(names[i], e->steps[i].name, 32);
    }
    for (auto &x : ev) cudaEventDestroy(x);
    return ANX_OK;
}

}   // extern "C"
