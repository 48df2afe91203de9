/* anatomix_b200 -- C ABI of the B200 (sm_100a) U-Net forward engine.
 *
 * Drop-in boundary for ONE hot path of neel-dey/anatomix: the standard branch of
 * `Unet.forward` (reference anatomix/model/network.py:530-548) of the network
 * built by `Unet.__init__` (network.py:262-465).  The reference has no FFI of its
 * own (it is pure Python over torch.nn); the entry points below are what a
 * binding for this path needs: describe the network the constructor would
 * build, hand over each conv's parameters (the state-dict tensors
 * `model.<idx>.weight/.bias` and the following norm's
 * `weight/bias/running_mean/running_var`), then run forwards on caller-owned
 * device (or host) buffers.  Plain C types only; no torch types, no exceptions,
 * no allocation on the forward path, nothing aborts.
 *
 * Layout contract (same as the reference module):
 *   input   fp32  NCDHW  [N, input_nc, D, H, W]   contiguous
 *   output  fp32  NCDHW  [N, output_nc, D, H, W]  contiguous
 *   each of D, H, W a multiple of 2^num_downs and >= 2 * 2^num_downs
 *   (smaller / ragged shapes fail in the reference too: reflect padding of a
 *   size-1 bottleneck, or a pool/upsample size mismatch at the skip concat).
 */
#ifndef ANATOMIX_B200_H
#define ANATOMIX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct anx_engine anx_engine;

typedef enum anx_status {
    ANX_OK = 0,
    ANX_ERR_BAD_ARG = 1,       /* null pointer, bad ordinal, bad enum            */
    ANX_ERR_BAD_SHAPE = 2,     /* N/D/H/W not usable with this network depth     */
    ANX_ERR_UNSUPPORTED = 3,   /* configuration outside the engine's fast path   */
    ANX_ERR_WORKSPACE = 4,     /* workspace null / too small / misaligned        */
    ANX_ERR_CUDA = 5,          /* a CUDA call failed; see anx_engine_last_error  */
    ANX_ERR_NOT_READY = 6,     /* forward before every conv was set              */
    ANX_ERR_NO_DEVICE = 7      /* no sm_100 device / driver entry point missing  */
} anx_status;

enum { ANX_NORM_NONE = 0, ANX_NORM_BATCH_EVAL = 1, ANX_NORM_INSTANCE = 2 };
enum { ANX_ACT_NONE = 0, ANX_ACT_RELU = 1, ANX_ACT_LEAKY = 2 };
enum { ANX_POOL_MAX = 0, ANX_POOL_AVG = 1 };
enum { ANX_INTERP_NEAREST = 0, ANX_INTERP_TRILINEAR = 1 };
enum { ANX_LOC_HOST = 0, ANX_LOC_DEVICE = 1 };

/* Mirrors the constructor arguments of reference network.py:262-279 that the
 * engine supports (dimension=3, pad_type='reflect', doubleconv=True,
 * use_skip_connection=True, residual_connection=False, final_act='none'). */
typedef struct anx_unet_desc {
    uint32_t struct_size;   /* = sizeof(anx_unet_desc), for forward compatibility   */
    int32_t input_nc;       /* network.py:265                                       */
    int32_t output_nc;      /* network.py:266                                       */
    int32_t num_downs;      /* network.py:267                                       */
    int32_t ngf;            /* network.py:268; must be a multiple of 16             */
    int32_t norm_kind;      /* ANX_NORM_*: 'batch' in eval mode / 'instance' / none */
    float   norm_eps;       /* network.py:278                                       */
    int32_t act_kind;       /* ANX_ACT_*: 'relu' / 'lrelu' / 'none'                 */
    float   act_slope;      /* 0.3 for 'lrelu' (network.py:191)                     */
    int32_t pool_kind;      /* ANX_POOL_*    (network.py:297)                       */
    int32_t interp_kind;    /* ANX_INTERP_*  (network.py:407)                       */
    int32_t device;         /* CUDA device ordinal the engine lives on              */
    uint32_t flags;         /* ANX_FLAG_* below                                     */
} anx_unet_desc;

/* debugging / measurement switches */
#define ANX_FLAG_FORCE_SIMT 1u   /* run every conv on the CUDA-core debug kernel  */
/* 16-bit storage type of activations and packed weights (fp32 accumulate either way).
 * Default: bf16 for BatchNorm / no-norm networks (activations are unbounded), fp16 for
 * InstanceNorm networks (every stored tensor is bounded by the normalisation and the
 * extra mantissa bits are needed through 24 normalised layers). */
#define ANX_FLAG_STORE_FP16 2u
#define ANX_FLAG_STORE_BF16 4u
/* Depth-slab mode (one oversized volume split along D over several GPUs): the input
 * tensor is [N, C, D+2, H, W] with one extra plane at each end of D (the neighbouring
 * slab's boundary plane, or the caller's reflect copy at a global face). */
#define ANX_FLAG_DEPTH_HALO_INPUT 8u
/* Keep the decoder's level-0 conv as one launch over the materialised upsampled tensor instead of
 * evaluating its upsampled half at low resolution (tests use this to compare code paths). */
#define ANX_FLAG_NO_UPCONV 16u
/* Run the thin 16 -> 16 layers on the generic tile kernel instead of the row kernel (tests compare the two). */
#define ANX_FLAG_NO_ROWS 32u
/* Give every intermediate tensor its own region of the workspace (no liveness-based sharing), so that all of them
 * survive the forward: debugging / layer-by-layer comparisons.  Depth-slab engines never share regions. */
#define ANX_FLAG_NO_WS_REUSE 64u

/* Replaces: Unet.__init__ (network.py:262-465).  Builds the layer program and
 * device constants; no parameters yet. */
anx_status anx_engine_create(const anx_unet_desc *desc, anx_engine **out_engine);
void anx_engine_destroy(anx_engine *engine);

/* Number of convolutions in network order (20 for `anatomix`, 24 for
 * `anatomix-dev`) and the shape of each, so a binding can check its state dict. */
int32_t anx_engine_num_convs(const anx_engine *engine);
anx_status anx_engine_conv_info(const anx_engine *engine, int32_t ordinal,
                                int32_t *module_index, int32_t *cin, int32_t *cout,
                                int32_t *has_norm);

/* Replaces: load_state_dict for one conv block (load_from_hf.py:48).
 * `weight` is `model.<idx>.weight` fp32 [cout, cin, 3, 3, 3]; `bias` is
 * `model.<idx>.bias` or NULL (network.py:292); the four bn_* arrays are the
 * following BatchNorm3d's weight / bias / running_mean / running_var (all NULL
 * unless norm_kind == ANX_NORM_BATCH_EVAL and the conv has a norm).  The engine
 * folds, rounds and repacks into its own device buffers; the caller keeps
 * ownership of the inputs.  `location` says whether the pointers are host or
 * device memory.  Not thread-safe against concurrent forwards. */
anx_status anx_engine_set_conv(anx_engine *engine, int32_t ordinal,
                               const float *weight, const float *bias,
                               const float *bn_weight, const float *bn_bias,
                               const float *bn_running_mean, const float *bn_running_var,
                               int32_t location);

/* Scratch the forward needs for a given input shape (activations between
 * layers; caller-owned so torch's caching allocator can serve it).  0 on a bad
 * shape.  Must be 256-byte aligned.  Tensors whose lifetimes do not overlap share
 * memory, so after a forward only the tensors still live at its end are intact:
 * feature taps are exported right after the launch that completes them
 * (anx_engine_tap_info: last_step). */
size_t anx_engine_workspace_bytes(const anx_engine *engine, int32_t n, int32_t d,
                                  int32_t h, int32_t w);

/* Replaces: Unet.forward standard branch (network.py:530-548) with device
 * buffers.  Asynchronous on `stream` (a cudaStream_t); no allocation, no
 * synchronisation.  Distinct workspaces allow concurrent calls. */
anx_status anx_engine_forward(anx_engine *engine, const float *in_ncdhw, float *out_ncdhw,
                              int32_t n, int32_t d, int32_t h, int32_t w,
                              void *workspace, size_t workspace_bytes, void *stream);

/* Step-wise execution, for callers that must act between layers (the depth-slab
 * partition exchanges one halo plane with its neighbours after every producer).
 * A step is one kernel launch of the forward; `anx_engine_step_info` names the
 * activation buffer (index into anx_engine_buffer_info, -1 = network output) and the
 * 8-channel group range the step writes.  `anx_engine_run_steps` runs steps
 * [first, last) of the same program `anx_engine_forward` runs in one go. */
int32_t anx_engine_num_steps(const anx_engine *engine);
anx_status anx_engine_step_info(const anx_engine *engine, int32_t step, int32_t *kind,
                                int32_t *out_buffer, int32_t *out_group_offset,
                                int32_t *out_groups, char name[32]);
anx_status anx_engine_run_steps(anx_engine *engine, const float *in_ncdhw, float *out_ncdhw,
                                int32_t n, int32_t d, int32_t h, int32_t w,
                                void *workspace, size_t workspace_bytes, void *stream,
                                int32_t first_step, int32_t last_step);

/* Depth-slab mode for networks with InstanceNorm and / or trilinear upsampling (`anatomix-dev`):
 * `anx_engine_set_slab` tells the engine which z faces of its slab touch a neighbouring slab (the
 * trilinear upsample then reads the neighbour's boundary plane from the shell instead of clamping) and
 * the depth of the WHOLE volume (InstanceNorm statistics are per whole volume: mean / variance are taken
 * over depth_total x H x W).  Every conv that feeds an InstanceNorm accumulates its sums into a block of
 * doubles [n][cout rounded up to 16][sum, sum of squares] inside the workspace; `anx_engine_step_stats`
 * gives that block for a step (bytes = 0 for any other step), and the caller all-reduces it over the
 * slabs between that step and the normalisation step that follows.  depth_total = 0 switches the mode off.
 * anx_engine_set_slab changes engine state: call it between forwards, not during one. */
anx_status anx_engine_set_slab(anx_engine *engine, int32_t has_lower_neighbour,
                               int32_t has_upper_neighbour, int32_t depth_total);
anx_status anx_engine_step_stats(const anx_engine *engine, int32_t step, int32_t n, int32_t d,
                                 int32_t h, int32_t w, size_t *offset, size_t *bytes);

/* Batch-sharded forward fused with the feature all-gather: the last conv's epilogue
 * stores this rank's `n` output volumes straight into EVERY rank's gather buffer
 * (`out_peers[r]` = NVLink-mapped device pointer to rank r's fp32
 * [world*n, output_nc, D, H, W] buffer, e.g. from torch symmetric memory or CUDA IPC)
 * at sample offset rank*n, so the transfer overlaps the conv tile by tile and no
 * separate collective runs.  The caller synchronises the ranks afterwards (a
 * barrier) before anyone reads its buffer.  At most 8 peers. */
anx_status anx_engine_forward_allgather(anx_engine *engine, const float *in_ncdhw,
                                        float *const *out_peers, int32_t world, int32_t rank,
                                        int32_t n, int32_t d, int32_t h, int32_t w,
                                        void *workspace, size_t workspace_bytes, void *stream);

/* The same fused gather with a selectable payload.  ANX_PAYLOAD_F32_NCDHW: fp32 [world*n, C, D, H, W] buffers
 * (bit-identical to anx_engine_forward on every rank).  ANX_PAYLOAD_CL16: 16-bit channels-last
 * [world*n, D, H, W, C] buffers in the engine's storage type (anx_engine_storage_type) -- half the NVLink
 * bytes, 1 KB contiguous per warp store; the values are the fp32 results rounded once to 16 bits, and
 * anx_widen_cl16_f32 turns a gathered buffer into the reference's fp32 NCDHW layout locally.  With
 * world = 1, rank = 0 this is a plain forward that writes the chosen payload (anx_engine_forward_cl16). */
enum { ANX_PAYLOAD_F32_NCDHW = 0, ANX_PAYLOAD_CL16 = 1 };
anx_status anx_engine_forward_gather(anx_engine *engine, const float *in_ncdhw,
                                     void *const *out_peers, int32_t world, int32_t rank, int32_t payload,
                                     int32_t n, int32_t d, int32_t h, int32_t w,
                                     void *workspace, size_t workspace_bytes, void *stream);
anx_status anx_engine_forward_cl16(anx_engine *engine, const float *in_ncdhw, void *out_cl16,
                                   int32_t n, int32_t d, int32_t h, int32_t w,
                                   void *workspace, size_t workspace_bytes, void *stream);
/* 0 = bf16, 1 = fp16: the 16-bit type of stored activations and of the CL16 payload. */
int32_t anx_engine_storage_type(const anx_engine *engine);
/* 16-bit channels-last [n, D, H, W, channels] -> fp32 NCDHW [n, channels, D, H, W] on the current device. */
anx_status anx_widen_cl16_f32(const void *src_cl16, float *dst_ncdhw, int64_t n, int32_t channels,
                              int32_t d, int32_t h, int32_t w, int32_t storage_type, void *stream);

/* Feature all-gather by the copy engines: copies `bytes` from `src` (this rank's slice, already inside its
 * own gather buffer) to `peer_dst[r]` for every r != rank (NVLink-mapped addresses of this rank's slot in
 * rank r's gather buffer), one copy per destination on internal streams forked from / joined back into
 * `stream`.  No SM is used, so the pushes of step k overlap the convs of step k + 1 when `stream` is a side
 * stream.  The ranks synchronise afterwards (a barrier) before reading.  Replaces: the `ncclAllGather` of
 * BASELINE.json configs[3]. */
anx_status anx_push_to_peers(anx_engine *engine, const void *src, void *const *peer_dst,
                             int32_t world, int32_t rank, size_t bytes, void *stream);

/* Depth-slab forward of ONE oversized volume split along D over the GPUs of a box (BASELINE.json
 * configs[4]) with the halo exchange done by the engine: after every launch that produces an activation
 * tensor, one small kernel stores this slab's two boundary planes of that tensor straight into the
 * neighbours' shell planes (through `lower_workspace` / `upper_workspace`, the neighbours' workspaces
 * mapped into this process, e.g. torch symmetric memory; NULL at a global face), publishes a sequence
 * number in their flag words and waits for theirs -- no host code, library collective or staging copy
 * between launches.  Requirements: every rank calls this the same number of times with the same shape
 * (slabs of equal depth: the workspace layout must be identical on all ranks); `flags` points to two
 * zero-initialised uint32 in peer-visible memory ([0] written by the lower, [1] by the upper neighbour),
 * `lower_flags` / `upper_flags` are the neighbours' `flags` mapped here; anx_engine_set_slab has been
 * called with the matching faces; the engine was created with ANX_FLAG_DEPTH_HALO_INPUT (the input
 * carries the neighbours' boundary planes).  BatchNorm(eval) / no-norm networks only: InstanceNorm needs
 * the caller's all-reduce between launches (anx_engine_run_steps + anx_engine_step_stats). */
typedef struct anx_slab_links {
    uint32_t struct_size;                  /* = sizeof(anx_slab_links) */
    void *lower_workspace, *upper_workspace;
    uint32_t *flags, *lower_flags, *upper_flags;
} anx_slab_links;
anx_status anx_engine_forward_slab(anx_engine *engine, const float *in_ncdhw, float *out_ncdhw,
                                   int32_t n, int32_t d, int32_t h, int32_t w,
                                   void *workspace, size_t workspace_bytes,
                                   const anx_slab_links *links, void *stream);

/* Same call for HOST buffers (pinned memory recommended): copies the input to
 * `dev_in`, runs the forward, copies `dev_out` back, all queued on `stream`.
 * `dev_in` / `dev_out` are caller-owned device staging buffers of the input /
 * output size.  This is the end-to-end entry point bench.py times. */
anx_status anx_engine_forward_host(anx_engine *engine, const float *in_host, float *out_host,
                                   int32_t n, int32_t d, int32_t h, int32_t w,
                                   float *dev_in, float *dev_out,
                                   void *workspace, size_t workspace_bytes, void *stream);

/* Host-buffer forward with a selectable output payload: ANX_PAYLOAD_CL16 downloads the features as 16-bit
 * channels-last [N, D, H, W, C] (half the PCIe bytes of the fp32 drop-in output; `out_host` / `dev_out` are
 * then buffers of that size and type). */
anx_status anx_engine_forward_host_ex(anx_engine *engine, const float *in_host, void *out_host,
                                      int32_t payload, int32_t n, int32_t d, int32_t h, int32_t w,
                                      float *dev_in, void *dev_out,
                                      void *workspace, size_t workspace_bytes, void *stream);

/* Pipelined form for back-to-back calls: like anx_engine_forward_host_ex, but `stream` does NOT wait for the download,
 * so the upload and the convs of the next call overlap it and the PCIe link never idles.  Results (of every call so
 * far) are complete once `stream` has passed an anx_engine_host_wait.  Alternate two {dev_out, out_host} sets: a call
 * waits for the earlier download that still reads the dev_out buffer it is about to overwrite (with a single set
 * the calls simply serialise).  The host input buffer must stay untouched until its upload has run, as with any
 * asynchronous copy. */
anx_status anx_engine_forward_host_pipelined(anx_engine *engine, const float *in_host, void *out_host,
                                             int32_t payload, int32_t n, int32_t d, int32_t h, int32_t w,
                                             float *dev_in, void *dev_out,
                                             void *workspace, size_t workspace_bytes, void *stream);
anx_status anx_engine_host_wait(anx_engine *engine, void *stream);

/* ---- rows next to the hot path -------------------------------------------
 *
 * Linear head fused into the last conv's epilogue: out[k] = bias[k] + sum_c weight[k][c] * y[c]
 * for k < head_nc, i.e. a 1x1x1 Conv3d on the network output evaluated on the accumulators
 * before they leave the SM.  Replaces: `nn.Sequential(Unet, UnetOutBlock(3, C, n_classes + 1))(x)`
 * (reference anatomix/segmentation/segmentation_utils.py:114-115) and, with a diagonal weight,
 * `pred * downscale_feat_scalar` (anatomix/registration/run_convex_adam_with_network_feats.py:166-167).
 * `weight` is fp32 [head_nc][output_nc] (the Conv3d weight with its 1x1x1 tail dropped), `bias`
 * fp32 [head_nc] or NULL.  Afterwards every forward writes fp32 [N, head_nc, D, H, W]
 * (anx_engine_out_channels).  head_nc = 0 removes the head.  Needs output_nc <= 16, head_nc <= 32.
 * Like anx_engine_set_conv it changes engine state: not thread-safe against concurrent forwards. */
anx_status anx_engine_set_head(anx_engine *engine, int32_t head_nc, const float *weight,
                               const float *bias, int32_t location);
int32_t anx_engine_out_channels(const anx_engine *engine);

/* Feature taps -- `Unet.forward(input, layers=[...], encode_only)` (network.py:475-529).  The engine
 * can hand out the activation after Sequential slot `module_index` when it stores that tensor:
 * the last slot of every conv block (its post-norm, post-activation output), the pooling slots,
 * the Upsample slots (where the reference taps cat(skip, upsampled), network.py:545) and the last
 * conv (the network output itself).  Pre-norm conv outputs are not stored; they are served by the separate
 * entry points below (anx_engine_tap_kind / set_tap_conv / export_prenorm_tap).
 * `anx_engine_tap_info` lists the available sites; `last_step` is the launch after which the tensor
 * is complete (run [0, last_step] with anx_engine_run_steps for `encode_only`); `is_output` marks the
 * last conv, whose tensor is the forward's fp32 output buffer itself;
 * `anx_engine_export_tap` converts tap `k` to fp32 NCDHW [N, channels, D>>level, H>>level, W>>level]. */
int32_t anx_engine_num_taps(const anx_engine *engine);
anx_status anx_engine_tap_info(const anx_engine *engine, int32_t k, int32_t *module_index,
                               int32_t *channels, int32_t *level, int32_t *last_step,
                               int32_t *is_output);
anx_status anx_engine_export_tap(anx_engine *engine, int32_t k, int32_t n, int32_t d, int32_t h,
                                 int32_t w, void *workspace, size_t workspace_bytes,
                                 float *out_ncdhw, void *stream);

/* Zero-copy channel concat: the forward writes its output channels as channels [channel_offset, channel_offset + C)
 * of a wider fp32 [N, dst_channels, D, H, W] tensor whose other channels the caller fills.  Replaces:
 * `torch.concatenate([mind_fixed, pred_fixed], dim=1)` (MIND-SSC descriptors in front of the network features,
 * anatomix/registration/instance_optimization.py:16-119): the 134 MB-per-volume feature tensor is not copied. */
anx_status anx_engine_forward_concat(anx_engine *engine, const float *in_ncdhw, float *dst_ncdhw,
                                     int32_t dst_channels, int32_t channel_offset,
                                     int32_t n, int32_t d, int32_t h, int32_t w,
                                     void *workspace, size_t workspace_bytes, void *stream);

/* Voxelwise normalisation across the channels of a contiguous fp32 [n, channels, D, H, W] device tensor, in place
 * allowed: mode 0 = unit L2 norm (x / max(||x||, eps)), mode 1 = zero mean / unit (unbiased) standard deviation
 * ((x - mean) / (std + eps)).  Replaces: the per-voxel feature normalisation the reference prescribes for the dev
 * models before registration / visualisation (README.md:13,49). */
anx_status anx_channel_normalize_f32(const float *in, float *out, int64_t n, int32_t channels,
                                     int32_t d, int32_t h, int32_t w, int32_t mode, float eps, void *stream);

/* PRE-norm conv taps (network.py:504-515: `layers` holding the index of an nn.Conv3d that a norm follows; the
 * pretraining code taps these slots, pretrain_anatomix.py:383-387, supcl_model.py:275-280).  The engine never stores
 * that tensor (eval BatchNorm is folded into the packed weights, InstanceNorm normalises in place), so such a tap is
 * served by an un-folded clone of the conv: anx_engine_tap_kind says which sites are of this kind and which conv
 * ordinal they belong to; anx_engine_set_tap_conv hands over that conv's plain `weight` / `bias` once (and again
 * whenever they change); anx_engine_export_prenorm_tap re-runs the conv's launch on its still-live input tensor --
 * call it right after anx_engine_run_steps has executed step `last_step` of the site -- and writes
 * conv(x) + bias as fp32 NCDHW [N, channels, D>>level, H>>level, W>>level].  `in_ncdhw` is the forward's input
 * (read again only when the tapped conv is the stem).  Not offered for the decoder conv that runs as two launches
 * (use an engine created with ANX_FLAG_NO_UPCONV for a tap there). */
enum { ANX_TAP_STORED = 0, ANX_TAP_OUTPUT = 1, ANX_TAP_PRENORM = 2 };
anx_status anx_engine_tap_kind(const anx_engine *engine, int32_t k, int32_t *kind, int32_t *conv_ordinal);
anx_status anx_engine_set_tap_conv(anx_engine *engine, int32_t conv_ordinal, const float *weight,
                                   const float *bias, int32_t location);
anx_status anx_engine_export_prenorm_tap(anx_engine *engine, int32_t k, const float *in_ncdhw,
                                         int32_t n, int32_t d, int32_t h, int32_t w,
                                         void *workspace, size_t workspace_bytes,
                                         float *out_ncdhw, void *stream);

/* out = scale * avg_pool3d(in, k, stride=k) on a contiguous fp32 [nc, D, H, W] device tensor (floor
 * mode), on the current device.  Replaces: `pred * downscale_feat_scalar` followed by
 * `F.avg_pool3d(features, grid_sp, stride=grid_sp)`
 * (anatomix/registration/run_convex_adam_with_network_feats.py:166-167, 198-205). */
anx_status anx_avgpool3d_scale_f32(const float *in, float *out, int64_t nc, int32_t d, int32_t h,
                                   int32_t w, int32_t k, float scale, void *stream);

/* One window of a sliding-window scan: out[c, z0+z, y0+y, x0+x] += pred[c, z, y, x] * weight[z, y, x] and
 * norm[z0+z, ...] += weight[z, y, x], for fp32 device tensors pred [channels, d, h, w], weight [d, h, w],
 * out [channels, D, H, W], norm [D, H, W].  Replaces the accumulate step of MONAI's
 * `sliding_window_inference` as the reference calls it (anatomix/registration/convex_adam_utils.py:202-219,
 * anatomix/segmentation/train_segmentation.py:194-199).  Windows of one scan overlap: launch them one
 * after another on one stream. */
anx_status anx_blend_window_f32(const float *pred, const float *weight, float *out, float *norm,
                                int32_t channels, int32_t d, int32_t h, int32_t w, int32_t D, int32_t H,
                                int32_t W, int32_t z0, int32_t y0, int32_t x0, void *stream);

/* Number of kernel launches one forward of this shape issues (for bench.py's
 * `gpu_launches`); -1 on a bad shape. */
int32_t anx_engine_launches_per_forward(const anx_engine *engine, int32_t n, int32_t d,
                                        int32_t h, int32_t w);

/* Per-conv timing hook for profiling: runs the forward once with CUDA events
 * around every launch and writes `count` (<= capacity) milliseconds into
 * `ms_out` plus a short name into `names_out[i][32]`.  Synchronises. */
anx_status anx_engine_profile(anx_engine *engine, const float *in_ncdhw, float *out_ncdhw,
                              int32_t n, int32_t d, int32_t h, int32_t w,
                              void *workspace, size_t workspace_bytes, void *stream,
                              float *ms_out, char (*names_out)[32], int32_t capacity,
                              int32_t *count);

/* Introspection of the workspace layout (tests / debugging / the depth-slab halo exchange):
 * activation buffer `index` holds `groups` 8-channel groups of a reflect-padded planar
 * 16-bit tensor [n][g][D/2^level+2][H/2^level+2][pitch][8] at `offset` in the workspace;
 * anx_engine_row_layout gives, for an interior width w = W/2^level, the `pitch` (voxels per
 * row) and the `lead`: voxels [lead, lead + w + 2) of a row are the x shell + interior (the
 * lead keeps interior rows on 32-byte sector boundaries; lead / tail voxels are never read). */
int32_t anx_engine_num_buffers(const anx_engine *engine);
anx_status anx_engine_buffer_info(const anx_engine *engine, int32_t n, int32_t d, int32_t h,
                                  int32_t w, int32_t index, size_t *offset, size_t *bytes,
                                  int32_t *level, int32_t *groups);

anx_status anx_engine_row_layout(const anx_engine *engine, int32_t w, int32_t *lead, int32_t *pitch);

const char *anx_status_string(anx_status status);
/* Detail of the last failure on this engine (thread-unsafe convenience). */
const char *anx_engine_last_error(const anx_engine *engine);
int32_t anx_version(void);

/* Self-test of the tcgen05 / TMA primitives the conv kernel relies on (small
 * GEMM through hand-built shared-memory descriptors, a TMA brick load and one
 * conv tile against a CUDA-core reference).  Returns ANX_OK when every probe
 * matches; writes a human-readable report (NUL-terminated) into `report`. */
anx_status anx_selftest(int32_t device, char *report, size_t report_bytes);

#ifdef __cplusplus
}
#endif
#endif /* ANATOMIX_B200_H */
