/* pml.h -- C ABI of libpml.so, the B200 (sm_100a) photometric-loss library.
 *
 * Drop-in boundary for the hot path of MariBax/self-supervised-depth-estimation
 * (all file:line citations are into that repository):
 *
 *   Trainer.generate_images_pred      trainer.py:465-515  (trainer_fusion.py:421-472,
 *                                     trainer_fusion_v3.py:447-482, trainer_gru.py:864-908)
 *   Trainer.compute_reprojection_loss trainer.py:517-529
 *   Trainer.compute_losses            trainer.py:531-622  (trainer_fusion.py:488-579,
 *                                     trainer_fusion_v3.py:498-590, trainer_gru.py:926-1023)
 *   layers.disp_to_depth :16-25, BackprojectDepth :139-168, Project3D :171-193,
 *   get_smooth_loss :202-215, SSIM :218-248, transformation_from_parameters :28-103
 *
 * The reference has no FFI / plugin layer (it is pure Python calling ATen); the entry points
 * below are what a ctypes binding for that path binds (INTEGRATION.md shows the stub).  Rules:
 *   - every pointer is a DEVICE pointer to contiguous fp32 (uint8 for argmin) unless noted;
 *   - the caller (PyTorch's caching allocator) owns every buffer including the workspace: the
 *     library never allocates, never synchronises, keeps no global mutable state, and only
 *     enqueues work on the given stream (CUDA-graph capturable, re-entrant);
 *   - every function returns PML_OK or a negative pml_status; nothing throws or aborts;
 *   - there is no CPU fallback.
 */
#ifndef PML_H_
#define PML_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ABI history: 1 = fused loss + layer kernels + depth metrics; 2 = pml_pass.frame_weight / grad_frame_weight
 * (predictive mask), pml_upsample_*, pml_bce_ones_*, pml_pyramid_u8; 3 = PML_MAX_SOURCES 4 -> 8 (array sizes in
 * pml_problem) and pml_problem.loss_vector; 4 = pml_problem.loss_total / loss_total_div (the mean over scales,
 * trainer.py:621, written by the library), pml_problem.segments (per-timestep tensors of the sequence trainer consumed
 * in place), pml_problem.seed_device, pml_scale_grads takes the upstream of the total, pml_selection_masks, pml_disp_head_*, PML_FLAG_KERNEL_CTA.
 * The Python binding refuses a library of another version. */
#define PML_ABI_VERSION 4
#define PML_MAX_SOURCES 8 /* source frames per target, e.g. (-1, 1, "s") = 3; BASELINE config 5 sweeps 2/4/8 */
#define PML_MAX_PASSES 8  /* scales handled by one call */
#define PML_MAX_SEGMENTS 16 /* separately allocated chunks of the batch (len_sequence of trainer_gru.py, default 10) */

typedef struct CUstream_st* pml_stream_t; /* == cudaStream_t */

typedef enum pml_status {
    PML_OK = 0,
    PML_ERR_INVALID = -1,     /* bad size / null pointer / misaligned pointer */
    PML_ERR_UNSUPPORTED = -2, /* S > PML_MAX_SOURCES, non power-of-two scale ratio, ... */
    PML_ERR_WORKSPACE = -3,   /* workspace too small (see pml_workspace_bytes) */
    PML_ERR_CUDA = -4         /* a CUDA runtime call failed (launch error) */
} pml_status;

/* flags, mirroring the loss-affecting options (options.py:184-198) */
#define PML_FLAG_NO_SSIM 1u      /* --no_ssim             trainer.py:523 */
#define PML_FLAG_NO_AUTOMASK 2u  /* --disable_automasking trainer.py:556 */
#define PML_FLAG_AVG_REPROJ 4u   /* --avg_reprojection    trainer.py:565,585 */
#define PML_FLAG_KERNEL_CTA 256u /* testing: run the first-generation CTA-strip kernel (S <= 4, no frame weights)
                                    instead of the warp-strip sweep; same results, kept as an independent cross-check */

/* One scale ("pass") of the loss: trainer.py:469 / :538 loop body. */
typedef struct pml_pass {
    int32_t hd, wd;            /* resolution of disp and smooth_color; H/hd == W/wd == 2^k (or 1) */
    float smooth_weight;       /* disparity_smoothness / 2^scale (trainer.py:616) */
    int32_t reserved;
    const float* disp;         /* [B,1,hd,wd]  outputs[("disp", s)] */
    const float* smooth_color; /* [B,3,hd,wd]  inputs[("color", 0, s)] (trainer.py:547) */
    const float* noise;        /* [B,n_id,H,W] tie-break randn (trainer.py:594) or NULL => Philox(seed) */
    uint8_t* argmin;           /* [B,H,W] out: index returned by torch.min (trainer.py:604); nullable */
    float* depth;              /* [B,1,H,W] out: outputs[("depth",0,s)] (trainer.py:480); nullable */
    float* warped;             /* [S,B,3,H,W] out: outputs[("color",f,s)] (trainer.py:508); nullable */
    float* grad_disp;          /* [B,1,hd,wd] out: d loss_s / d disp_s WITHOUT the per-image mean term
                                  of the smoothness normalisation (that constant is grad_disp_const);
                                  forward_backward only */
    const float* frame_weight; /* [B,S,H,W] or NULL: --predictive_mask (trainer.py:571-579), the mask
                                  already resized to H x W; multiplies the reprojection loss of each
                                  source frame before the mean / min.  Requires PML_FLAG_NO_AUTOMASK
                                  (the reference's `elif`), set on every pass or on none */
    float* grad_frame_weight;  /* [B,S,H,W] out: d loss_s / d frame_weight (forward_backward only) */
} pml_pass;

/* Optional: the image batch arrives as n_seg separately allocated chunks of seg_size images each
 * (B == n_seg * seg_size).  trainer_gru.py keeps one tensor per time step, inputs[("color", f, s, i)],
 * inputs[("K", s, i)], and concatenates them on the fly for every scale and frame (trainer_gru.py:890-899,
 * 943-957); with this table the kernels read the per-timestep tensors in place: image b lives in chunk
 * b / seg_size at index b % seg_size.  Pointers a problem does not use may be NULL.  When `segments` is set,
 * pml_problem.target / sources / K / inv_K and pml_pass.smooth_color are ignored.  HOST struct of DEVICE pointers. */
typedef struct pml_segments {
    int32_t n_seg, seg_size;
    const float* target[PML_MAX_SEGMENTS];                        /* [seg_size,3,H,W] each */
    const float* sources[PML_MAX_SOURCES][PML_MAX_SEGMENTS];      /* [seg_size,3,H,W] */
    const float* K[PML_MAX_SEGMENTS];                             /* [seg_size,4,4] */
    const float* inv_K[PML_MAX_SEGMENTS];                         /* [seg_size,4,4] */
    const float* smooth_color[PML_MAX_PASSES][PML_MAX_SEGMENTS];  /* [seg_size,3,hd,wd] */
} pml_segments;

/* A group of passes that share images, intrinsics and poses (all scales when
 * v1_multiscale is off; one scale per group when it is on). */
typedef struct pml_problem {
    int32_t B, H, W; /* resolution of target / sources (source_scale, trainer.py:471-476) */
    int32_t S;       /* number of source frames (reference hard-codes [-1, 1], trainer.py:482) */
    int32_t n_pass;
    uint32_t flags;
    float min_depth, max_depth; /* disp_to_depth, layers.py:16-25 */
    float eps;                  /* Project3D eps, layers.py:174 */
    int32_t reserved;
    uint64_t seed;                          /* Philox seed used when pass.noise == NULL */
    const float* target;                    /* [B,3,H,W] inputs[("color",0,source_scale)] */
    const float* sources[PML_MAX_SOURCES];  /* [B,3,H,W] inputs[("color",f,source_scale)] */
    const float* K;                         /* [B,4,4] */
    const float* inv_K;                     /* [B,4,4] */
    const float* T[PML_MAX_SOURCES];        /* [B,4,4] cam_T_cam / stereo_T per source */
    pml_pass pass[PML_MAX_PASSES];
    float* losses;          /* [n_pass,4] out: loss_s, photometric mean, smoothness term, 0 */
    float* grad_T;          /* [n_pass,S,B,4,4] out: d loss_s / d T_f          (forward_backward) */
    float* grad_disp_const; /* [n_pass,B] out: per-image constant to add to grad_disp
                               (gradient through disp.mean(), trainer.py:612)  (forward_backward) */
    void* prof_start;       /* optional cudaEvent_t recorded right before the fused sweep kernel */
    void* prof_stop;        /* optional cudaEvent_t recorded right after it (bench.py roofline) */
    float* loss_vector;     /* [n_pass] out, nullable: loss_s again as a contiguous vector (what autograd returns) */
    float* loss_total;      /* [1] out, nullable: (loss_0 + loss_1 + ...) / loss_total_div in scale order =
                               losses["loss"] (trainer.py:618-621) when loss_total_div == num_scales */
    float loss_total_div;   /* divisor of loss_total; 0 is read as n_pass */
    int32_t reserved2;
    const pml_segments* segments; /* nullable, see pml_segments */
    const uint64_t* seed_device;  /* nullable DEVICE pointer: *seed_device is xor'ed into `seed` by the kernels, so a captured
                                     CUDA graph draws fresh tie-break noise on every replay (the caller advances the value) */
} pml_problem;

int pml_abi_version(void);
const char* pml_strerror(int status);

/* Bytes of device workspace the two calls below need for this problem (16-byte aligned). */
size_t pml_workspace_bytes(const pml_problem* p);

/* generate_images_pred + compute_losses, forward only: losses, argmin, optional depth/warped. */
int pml_loss_forward(const pml_problem* p, void* workspace, size_t workspace_bytes, pml_stream_t stream);

/* Same pass, additionally producing the gradients of every loss_s with respect to disp_s and
 * T_f in the same sweep (the adjoint is evaluated while the tiles are on chip; autograd later
 * only scales them by the incoming gradient, pml_scale_grads). */
int pml_loss_forward_backward(const pml_problem* p, void* workspace, size_t workspace_bytes, pml_stream_t stream);

/* Backward of the autograd node: grad_disp_s <- up[s] * (grad_disp_s + const[s,b]) in place and
 * grad_T_out[f,b] = sum_s up[s] * grad_T[s,f,b], with up[s] = upstream[s] + upstream_total[0] / total_div.
 * `upstream` (device [n_pass]) is the gradient arriving at loss_vector, `upstream_total` (device [1]) the one
 * arriving at loss_total; either may be NULL (read as zero), not both. */
int pml_scale_grads(int32_t n_pass, int32_t B, int32_t S, const int32_t* hd, const int32_t* wd,
                    float* const* grad_disp, const float* grad_disp_const, const float* grad_T,
                    const float* upstream, const float* upstream_total, float total_div,
                    float* grad_T_out, pml_stream_t stream);

/* outputs["identity_selection/{s}"] (trainer.py:606-608) for every scale in one launch:
 * out[s][i] = argmin[s][i] > n_id - 1 ? 1.f : 0.f over n_pix = B*H*W pixels per scale. */
int pml_selection_masks(int32_t n_pass, int64_t n_pix, const uint8_t* const* argmin, int32_t n_id,
                        float* const* out, pml_stream_t stream);

/* ---- layer-level drop-ins (layers.py signatures), each one kernel forward + one backward ---- */
/* disp_to_depth, layers.py:16-25 */
int pml_disp_to_depth_fwd(const float* disp, float* scaled, float* depth, int64_t n, float min_depth, float max_depth, pml_stream_t);
int pml_disp_to_depth_bwd(const float* disp, const float* g_scaled, const float* g_depth, float* g_disp, int64_t n, float min_depth, float max_depth, pml_stream_t);
/* BackprojectDepth.forward, layers.py:163-168: depth [B,1,H,W], inv_K [B,4,4] -> cam [B,4,H*W] */
int pml_backproject_fwd(const float* depth, const float* inv_K, float* cam, int32_t B, int32_t H, int32_t W, pml_stream_t);
int pml_backproject_bwd(const float* g_cam, const float* inv_K, float* g_depth, int32_t B, int32_t H, int32_t W, pml_stream_t);
/* Project3D.forward, layers.py:182-193: points [B,4,H*W], K,T [B,4,4] -> grid [B,H,W,2] */
int pml_project_fwd(const float* points, const float* K, const float* T, float* grid, int32_t B, int32_t H, int32_t W, float eps, pml_stream_t);
/* g_points [B,4,H*W]; g_T_partial [B,nblk,12] + pml_project_bwd reduces it into g_T [B,4,4] */
int pml_project_bwd(const float* points, const float* K, const float* T, const float* g_grid, float* g_points, float* g_T,
                    void* workspace, size_t workspace_bytes, int32_t B, int32_t H, int32_t W, float eps, pml_stream_t);
size_t pml_project_bwd_workspace_bytes(int32_t B, int32_t H, int32_t W);
/* SSIM.forward, layers.py:234-248: x,y [B,C,H,W] -> [B,C,H,W] */
int pml_ssim_fwd(const float* x, const float* y, float* out, int32_t planes, int32_t H, int32_t W, pml_stream_t);
int pml_ssim_bwd(const float* x, const float* y, const float* g_out, float* g_x, float* g_y, int32_t planes, int32_t H, int32_t W, pml_stream_t);
/* get_smooth_loss, layers.py:202-215: disp [B,1,H,W], img [B,C,H,W] -> scalar (out[0]) */
int pml_smooth_fwd(const float* disp, const float* img, float* out, void* workspace, size_t workspace_bytes, int32_t B, int32_t C, int32_t H, int32_t W, pml_stream_t);
int pml_smooth_bwd(const float* disp, const float* img, const float* g_out, float* g_disp, float* g_img, int32_t B, int32_t C, int32_t H, int32_t W, pml_stream_t);
size_t pml_smooth_workspace_bytes(int32_t B, int32_t H, int32_t W);
/* transformation_from_parameters, layers.py:28-103: axisangle,translation [B,3] -> T [B,4,4] */
int pml_pose_fwd(const float* axisangle, const float* translation, float* T, int32_t B, int32_t invert, pml_stream_t);
int pml_pose_bwd(const float* axisangle, const float* translation, const float* g_T, float* g_axisangle, float* g_translation, int32_t B, int32_t invert, pml_stream_t);
/* F.interpolate(x, [H, W], mode="bilinear", align_corners=False) (trainer.py:474-475 on the
 * disparities, :574-576 on the predictive mask): x [planes,h,w] -> out [planes,H,W]; bwd is the
 * transposed resize as a deterministic gather */
int pml_upsample_fwd(const float* x, float* out, int32_t planes, int32_t h, int32_t w, int32_t H, int32_t W, pml_stream_t);
int pml_upsample_bwd(const float* g_out, float* g_x, int32_t planes, int32_t h, int32_t w, int32_t H, int32_t W, pml_stream_t);
/* nn.BCELoss()(mask, ones) (trainer.py:582): out[0] = mean(-max(log(mask), -100)); bwd follows ATen's
 * binary_cross_entropy_backward, g * (x - 1) / max((1 - x) x, 1e-12) / n */
size_t pml_bce_workspace_bytes(void);
int pml_bce_ones_fwd(const float* mask, int64_t n, float* out, void* workspace, size_t workspace_bytes, pml_stream_t);
int pml_bce_ones_bwd(const float* mask, const float* g_out, float* g_mask, int64_t n, pml_stream_t);


/* ---- DepthDecoder disparity heads: networks/depth_decoder.py:46-47,62-66 over
 * layers.Conv3x3 (layers.py:121-136): disp = sigmoid(conv3x3(reflection_pad1(x), weight) + bias), x [B,C,h,w],
 * weight [1,C,3,3], bias [1] -> disp [B,1,h,w], one kernel.  Backward (one kernel + a fixed-order reduction):
 * from the saved output and g_disp to g_x [B,C,h,w] (nullable), g_weight [1,C,3,3], g_bias [1]. ---- */
int pml_disp_head_fwd(const float* x, const float* weight, const float* bias, float* disp,
                      int32_t B, int32_t C, int32_t h, int32_t w, pml_stream_t);
size_t pml_disp_head_bwd_workspace_bytes(int32_t B, int32_t C, int32_t h, int32_t w);
int pml_disp_head_bwd(const float* x, const float* weight, const float* disp, const float* g_disp, float* g_x,
                      float* g_weight, float* g_bias, void* workspace, size_t workspace_bytes,
                      int32_t B, int32_t C, int32_t h, int32_t w, pml_stream_t);

/* ---- input colour pyramid: datasets/mono_dataset.py:84-111 resizes every
 * frame on the CPU with PIL (scale i = Resize((H >> i, W >> i), Image.ANTIALIAS) of scale i-1), applies
 * ToTensor and the trainer uploads all scales as fp32 (trainer.py:233-237).  Here the caller uploads the
 * uint8 scale-0 frames [N,H,W,3] (PIL / numpy layout) and gets out[s] = fp32 [N,3,H>>s,W>>s] for
 * s = 0..n_scales-1, bit-exact with Pillow's 8-bit Lanczos resampling + torchvision's ToTensor.  `out` is
 * a HOST array of n_scales device pointers; H and W must be multiples of 2^(n_scales-1). ---- */
size_t pml_pyramid_workspace_bytes(int32_t N, int32_t H, int32_t W, int32_t n_scales);
int pml_pyramid_u8(const uint8_t* frames, int32_t N, int32_t H, int32_t W, int32_t n_scales, float* const* out,
                   void* workspace, size_t workspace_bytes, pml_stream_t);

/* ---- monitoring metrics: Trainer.compute_depth_losses (trainer.py:624-652) over
 * layers.compute_depth_errors (layers.py:251-269).  prepare resizes depth [B,1,H,W] to the ground
 * truth's [B,1,Hg,Wg] (bilinear, align_corners=False), clamps, applies mask = gt>0 & crop and writes
 * dense pred/gt arrays (+inf where masked) plus the valid count; median_ratio finds the two medians
 * (element (count-1)/2 of each sorted array == torch.median) and reduce, given their ratio,
 * returns abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3. ---- */
size_t pml_depth_metrics_workspace_bytes(void);
/* ratio[0] = median(gt) / median(pred) over the valid entries (trainer.py:645), torch.median's lower-middle element,
 * found by radix selection on the device (no sort).  The workspace is the one of pml_depth_metrics_workspace_bytes. */
int pml_depth_metrics_median_ratio(const float* pred, const float* gt, int64_t n, const int32_t* count, float* ratio,
                                   void* workspace, size_t workspace_bytes, pml_stream_t);
int pml_depth_metrics_prepare(const float* depth, const float* depth_gt, float* pred_out, float* gt_out, int32_t* count,
                              int32_t B, int32_t H, int32_t W, int32_t Hg, int32_t Wg, int32_t crop_y0, int32_t crop_y1,
                              int32_t crop_x0, int32_t crop_x1, float min_depth, float max_depth, pml_stream_t);
int pml_depth_metrics_reduce(const float* pred, const float* gt, int64_t n, const float* ratio, const int32_t* count,
                             float min_depth, float max_depth, float* out7, void* workspace, size_t workspace_bytes, pml_stream_t);

#ifdef __cplusplus
}
#endif
#endif /* PML_H_ */
