/*
 * libmonovifi_b200.so -- C ABI of the B200-native Mono-ViFI training inner loop (sm_100a).
 *
 * The reference (LiuJF1226/Mono-ViFI) is pure Python over torch; it has no FFI of its own.  The boundary
 * below is what its `layers.py` / `train.py` call sites bind through ctypes (see INTEGRATION.md and
 * mono_vifi_b200/_lib.py).  Each entry point names the reference code it replaces (paths relative to
 * the reference checkout).
 *
 * Conventions
 *   - every tensor pointer is a DEVICE pointer to a contiguous fp32 NCHW buffer owned by the caller
 *     (a torch allocation); the library allocates nothing persistent;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream); all work is
 *     enqueued on it, nothing synchronises;
 *   - return value: 0 on success, negative mvf_status on failure; never throws, never exits;
 *     mvf_last_error() returns a thread-local description of the last failure;
 *   - `*_host` variants take HOST pointers and run H2D -> kernel -> D2H on the given stream, then
 *     synchronise it (the end-to-end path a non-torch caller would use).
 */
#ifndef MONOVIFI_B200_H
#define MONOVIFI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    MVF_OK = 0,
    MVF_ERR_INVALID = -1, /* bad argument (null pointer, non-positive size, H or W < 3 ...) */
    MVF_ERR_CUDA = -2,    /* a CUDA runtime call failed; see mvf_last_error() */
    MVF_ERR_WORKSPACE = -3 /* workspace too small */
} mvf_status;

/* option flags: options.py:173-181 (no_ssim, avg_reprojection, disable_automasking) */
#define MVF_NO_SSIM 1
#define MVF_AVG_REPROJECTION 2
#define MVF_DISABLE_AUTOMASKING 4

typedef struct {
    int B, H, W;
    float min_disp;   /* (float)(1/max_depth)               layers.py:21 */
    float disp_range; /* (float)(1/min_depth - 1/max_depth)  layers.py:22-23 */
    float smooth_w;   /* opt.disparity_smoothness            train.py:1049 */
    int flags;
} mvf_f1_params;

int mvf_version(void);
const char* mvf_last_error(void);

/* ---- fused view synthesis + photometric loss (F1) ------------------------------------------------
 * Replaces, for one loss group (one target, two sources):
 *   Trainer.generate_images_pred x2      train.py:956-971   (disp_to_depth layers.py:16-25,
 *                                         BackprojectDepth layers.py:192-197, Project3D layers.py:211-222,
 *                                         F.grid_sample border/align_corners=True)
 *   Trainer.compute_reprojection_loss x4 train.py:973-985   (SSIM layers.py:277-290)
 *   Trainer.compute_losses_base          train.py:987-1051  (get_smooth_loss layers.py:231-242)
 *
 * workspace: mvf_f1_workspace_bytes(B) bytes of device memory, zeroed ONCE with mvf_workspace_init and
 * then reused by every call on the same stream (the kernels leave it zeroed).
 */
size_t mvf_f1_workspace_bytes(int B);
int mvf_workspace_init(void* workspace, size_t bytes, void* stream);

/* inputs : disp[B,1,H,W] tgt/src0/src1[B,3,H,W] inv_K[B,4,4] P0/P1[B,3,4] (= (K@T)[:, :3])
 *          noise[B,nid,H,W] or NULL (nid = 2; 1 with avg_reprojection; the tensor train.py:1023 draws)
 *          mask_rec[B,1,H,W] or NULL
 * outputs: loss[4] = {loss, photometric mean, smoothness (unweighted), 0}
 *          stats[B,4] (saved for backward), idx[B,H,W] uint8 argmin of `combined` (auto_mask = idx > 1)
 * debug  : x0y0 int32 [2][2][B,H,W], warp0/warp1 [B,3,H,W], to_optimise [B,H,W]; all NULL in production */
int mvf_f1_forward(const mvf_f1_params* p, const float* disp, const float* tgt, const float* src0,
                   const float* src1, const float* inv_K, const float* P0, const float* P1, const float* noise,
                   const float* mask_rec, float* loss, float* stats, uint8_t* idx, int32_t* x0y0, float* warp0,
                   float* warp1, float* to_optimise, void* workspace, size_t workspace_bytes, void* stream);

/* gradient of loss[0] w.r.t. disp and P0/P1 (the only inputs autograd needs: SURVEY.md 9.3).
 * gout: device scalar dL/dloss or NULL (= 1).  idx/stats: as written by mvf_f1_forward. */
int mvf_f1_backward(const mvf_f1_params* p, const float* disp, const float* tgt, const float* src0,
                    const float* src1, const float* inv_K, const float* P0, const float* P1,
                    const float* mask_rec, const uint8_t* idx, const float* stats, const float* gout,
                    float* g_disp, float* g_P0, float* g_P1, void* workspace, size_t workspace_bytes,
                    void* stream);

/* host-buffer variant of the forward (allocates device scratch per call, synchronises) */
int mvf_f1_forward_host(const mvf_f1_params* p, const float* disp, const float* tgt, const float* src0,
                        const float* src1, const float* inv_K, const float* P0, const float* P1,
                        const float* noise, const float* mask_rec, float* loss, uint8_t* idx);

/* ---- stand-alone ops behind the reference's layers.py API (autograd pairs) --------------------------------
 * scratch for the reductions is the same zeroed workspace as F1's (mvf_f1_workspace_bytes(B)). */

/* disp_to_depth, layers.py:16-25.  scaled_disp / depth may be NULL. */
int mvf_disp_to_depth_fwd(const float* disp, float* scaled_disp, float* depth, size_t n, float min_disp,
                          float disp_range, void* stream);
int mvf_disp_to_depth_bwd(const float* disp, const float* g_scaled_disp, const float* g_depth, float* g_disp, size_t n,
                          float min_disp, float disp_range, void* stream);
/* BackprojectDepth.forward, layers.py:192-197: depth[B,1,H,W], inv_K[B,4,4] -> cam_points[B,4,H*W] */
int mvf_backproject_fwd(const float* depth, const float* inv_K, float* cam_points, int B, int H, int W, void* stream);
int mvf_backproject_bwd(const float* g_cam_points, const float* inv_K, float* g_depth, int B, int H, int W, void* stream);
/* Project3D.forward, layers.py:214-222 with P = (K@T)[:, :3] ([B,3,4], layers.py:212 stays a torch matmul):
 * points[B,4,H*W] -> pix_coords[B,H,W,2] in [-1,1] */
int mvf_project_fwd(const float* points, const float* P, float* pix_coords, int B, int H, int W, float eps, void* stream);
int mvf_project_bwd(const float* points, const float* P, const float* g_pix_coords, float* g_points, float* g_P,
                    void* workspace, size_t workspace_bytes, int B, int H, int W, float eps, void* stream);
/* SSIM.forward, layers.py:277-290 on N = B*C planes.  bwd gives the gradient w.r.t. x (call it with x and y
 * swapped for y: the formula is symmetric); coef_scratch: 3*N*H*W floats. */
int mvf_ssim_fwd(const float* x, const float* y, float* out, int N, int H, int W, void* stream);
int mvf_ssim_bwd(const float* x, const float* y, const float* g_out, float* g_x, float* coef_scratch, int N, int H,
                 int W, void* stream);
/* get_smooth_loss, layers.py:231-242: disp[B,1,H,W], img[B,3,H,W] -> loss[1]; gradient w.r.t. disp */
int mvf_smooth_loss_fwd(const float* disp, const float* img, float* loss, void* workspace, size_t workspace_bytes,
                        int B, int H, int W, void* stream);
int mvf_smooth_loss_bwd(const float* disp, const float* img, const float* gout, float* g_disp, int B, int H, int W,
                        void* stream);
/* Trainer.compute_SI_log_depth_loss, train.py:924-941: pred/target[B,1,H,W] (HW = H*W), mask or NULL ->
 * loss[1], stats[B,2] (saved for bwd); bwd gives gradients w.r.t. pred and target (either may be NULL) */
int mvf_si_log_fwd(const float* pred, const float* target, const float* mask, float* loss, float* stats,
                   void* workspace, size_t workspace_bytes, int B, size_t HW, float beta, void* stream);
int mvf_si_log_bwd(const float* pred, const float* target, const float* mask, const float* stats, const float* gout,
                   float* g_pred, float* g_target, int B, size_t HW, float beta, void* stream);

/* ---- tensor-core convolutions (tcgen05 / TMEM / TMA implicit GEMM, TF32 inputs, fp32 accumulate) ----------------
 * Replace the cuDNN calls behind nn.Conv2d in the reference's networks (layers.py:131,146 Conv3x3 / Conv1x1;
 * networks/monodepth2.py:16-96, networks/posenet.py:10-137 via torchvision ResNet) on CHANNELS-LAST fp32 tensors
 * (NCHW-shaped torch tensors in torch.channels_last memory format: element (b,c,y,x) at b*sB + y*sH + x*sW + c).
 * The filter bank is re-packed once per weight update with mvf_conv2d_pack_filters (dgrad = 0: forward;
 * dgrad = 1: the flipped / transposed bank, with which mvf_conv2d_forward on grad_out and pad' = k-1-pad computes
 * the input gradient of a stride-1 convolution).  Strides are in elements; the channel stride is 1.
 * Constraints (mvf_conv2d_supported): stride 1 or 2, Cin % 4 == 0, strides % 4 == 0, pointers 16-byte aligned. */
typedef struct {
    int B, Cin, H, W;               /* input  [B,Cin,H,W]  */
    int Cout, KH, KW, pad, stride;  /* output [B,Cout,(H+2pad-KH)/stride+1,(W+2pad-KW)/stride+1] */
    long long x_stride[3];          /* batch, row (y), column (x) */
    long long y_stride[3];
    int stride_x;                   /* column stride when it differs from `stride` (rows); 0 = same.  Used by the
                                       row-packed 7x7 stem: the input is viewed as [B, 8 px * C, H+6, Wo] with an
                                       overlapping column stride of 2 px, so the column stride of the conv is 1 */
} mvf_conv2d_desc;
#define MVF_ACT_NONE 0
#define MVF_ACT_RELU 1
#define MVF_ACT_ELU 2
size_t mvf_conv2d_packed_filter_floats(int N, int K, int KH, int KW);
int mvf_conv2d_pack_filters(const float* w /*[Cout,Cin,KH,KW]*/, float* packed, int Cout, int Cin, int KH, int KW,
                            int dgrad, void* stream);
int mvf_conv2d_supported(const mvf_conv2d_desc* d); /* 1 / 0 (reason in mvf_last_error) */
int mvf_conv2d_forward(const mvf_conv2d_desc* d, const float* x, const float* w_packed, const float* bias /*or NULL*/,
                       float* y, int act, void* stream);

/* Data gradient of a STRIDE-2 convolution (the three down-sampling 3x3 convolutions and 1x1 shortcuts of every ResNet
 * encoder, networks/monodepth2.py:16-31 via torchvision's BasicBlock): d describes the forward problem, x_stride the
 * gradient being produced (grad_x [B,Cin,H,W]) and y_stride the incoming one (grad_y [B,Cout,Ho,Wo]); w_packed is the
 * dgrad = 1 bank.  The input pixels split into four parity classes, each a small stride-1 convolution over grad_y that
 * lands on a stride-2 lattice of grad_x; classes without taps (1x1: three of four) are not written -- pass a zeroed
 * grad_x when KH or KW is 1.  mvf_conv2d_dgrad_s2_plan writes the class / tap table the kernel walks (n_classes, then
 * per class py, px, ntaps and ntaps x (dy, dx, packed tap index)) and returns the number of ints; it runs on the host.
 * Status: written at the end of round 1 behind MVF_DGRAD_S2=1 in the Python binding; the plan is checked on the CPU
 * (tests/test_dgrad_s2_plan.py), the kernel has not been run on hardware yet. */
int mvf_conv2d_dgrad_s2_supported(const mvf_conv2d_desc* d);
int mvf_conv2d_dgrad_s2(const mvf_conv2d_desc* d, const float* grad_y, const float* w_packed, float* grad_x, void* stream);
int mvf_conv2d_dgrad_s2_plan(int KH, int KW, int pad, int* table, int capacity);

/* weight gradient: grad_w[Cout,Cin,KH,KW] (contiguous, the parameter's layout) = sum over output pixels of
 * grad_out (x) input patches; d->x_stride describes x, d->y_stride describes grad_out (both channels-last).
 * Split-K partial sums go to `workspace` (mvf_conv2d_wgrad_workspace_floats(d) floats, caller-owned scratch) and are
 * added in a fixed order: results are bitwise reproducible.  Constraints (mvf_conv2d_wgrad_supported): at most 9
 * taps, channel counts % 4 == 0, Cout <= 32 or Cout % 32 == 0, stride 1 or 2. */
int mvf_conv2d_wgrad_supported(const mvf_conv2d_desc* d);
size_t mvf_conv2d_wgrad_workspace_floats(const mvf_conv2d_desc* d);
int mvf_conv2d_wgrad(const mvf_conv2d_desc* d, const float* x, const float* grad_out, float* grad_w, float* workspace,
                     size_t workspace_floats, void* stream);

/* Packs every filter bank of a training step in ONE launch (instead of one mvf_conv2d_pack_filters launch per layer and direction:
 * 155 per step of the ResNet18 configuration).  table (device memory): n_entries rows of 8 int64 {w pointer, out pointer, Cout, Cin,
 * KH, KW, dgrad, first block}; bank e occupies blocks [first block_e, first block_e + ceil(packed floats_e / mvf_conv2d_pack_chunk())),
 * total_blocks = their sum.  Same layouts as mvf_conv2d_pack_filters. */
int mvf_conv2d_pack_chunk(void);
int mvf_conv2d_pack_filters_multi(const long long* table, int n_entries, long long total_blocks, void* stream);

/* Convolution + bias + nn.PReLU(Cout) in the epilogue: the `convrelu` block of the frozen VFI network (networks/IFRNet.py:121-125,
 * every encoder / decoder convolution of IFRNet.py:153-330).  Same descriptor / packed bank as mvf_conv2d_forward; slope[Cout]. */
int mvf_conv2d_forward_prelu(const mvf_conv2d_desc* d, const float* x, const float* w_packed, const float* bias, const float* slope,
                             float* y, void* stream);
/* nn.ConvTranspose2d(Cin_t, Cout_t, k, stride 2, padding p) forward (IFRNet.py:194: 4x4, stride 2, padding 1) = the data gradient of
 * the stride-2 convolution with the same weight tensor, plus bias: four output-parity classes, each a small stride-1 convolution,
 * one persistent tcgen05 launch (the kernel of mvf_conv2d_dgrad_s2).  d describes that convolution: d->Cout = Cin_t (channels of x),
 * d->Cin = Cout_t (channels of y), H / W = the OUTPUT size, x_stride = strides of y, y_stride = strides of x; w_packed =
 * mvf_conv2d_pack_filters(weight viewed as [Cin_t, Cout_t, k, k], dgrad = 1). */
int mvf_conv_transpose2d_s2_fwd(const mvf_conv2d_desc* d, const float* x, const float* w_packed, const float* bias, float* y, void* stream);

/* ---- SyncBatchNorm exchange over NVLink peer memory (csrc/peer.cu) ------------------------------------------------------------
 * In-place SUM over all ranks of vec[n] (float64, n <= 2056: the [2C + 1] statistics vector of mvf_bn_sync_*), as ONE single-CTA kernel
 * on `stream`: remote stores of this rank's vector into a slot of every rank's symmetric buffer, release / acquire flags carrying a
 * sequence number, fixed rank-order sum (bitwise identical on every rank).  Replaces the per-call NCCL all-reduce of
 * torch.nn.SyncBatchNorm (train.py:205-208).  peers_dev: device array of `world` pointers to the ranks' symmetric buffers
 * (mvf_peer_buffer_bytes() bytes each, zero-initialised, e.g. torch.distributed._symmetric_memory); channel < 8: exchanges issued from
 * different streams must use different channels; seq_local: this rank's 64-bit exchange counter of that channel (device memory,
 * starts at 0, advanced by the kernel -- so CUDA-graph replays keep counting).  Every rank must issue the same exchanges in the same
 * order per channel.  world <= 16. */
size_t mvf_peer_buffer_bytes(void);
int mvf_peer_allreduce_f64(double* vec, int n, void* const* peers_dev, int rank, int world, int channel, unsigned long long* seq_local,
                           void* stream);

/* ---- evaluation path (csrc/eval.cu; train.py:419-483 test_kitti, layers.py:293-311 compute_depth_errors) -----------------------
 * mvf_bn_eval_fwd: y = relu?( BatchNorm2d_eval(x) + identity ) with the running statistics, dense channels-last [P][C], C % 4 == 0.
 * mvf_depth_eval: one image -- disp[h,w] (the scaled disparity of disp_to_depth) is resized to the ground truth's [Hg,Wg]
 *   (bilinear, align_corners=False), inverted to depth, masked (eigen_crop = 1: 1e-3 < gt < 80 inside the Eigen crop, = 0: gt > 0;
 *   min_depth / max_depth are those two bounds), scaled by median(gt) / median(pred) (torch.median's lower median; stereo_scale > 0
 *   uses that fixed factor instead), clamped to [min_depth, max_depth]; metrics8 = {abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3,
 *   scale ratio}.  workspace: mvf_depth_eval_workspace_bytes(Hg, Wg) bytes, 16-byte aligned.  No host synchronisation. */
int mvf_bn_eval_fwd(const float* x, const float* identity, float* y, const float* gamma, const float* beta, const float* running_mean,
                    const float* running_var, long long P, int C, float eps, int relu, void* stream);
size_t mvf_depth_eval_workspace_bytes(int Hg, int Wg);
int mvf_depth_eval(const float* disp, int h, int w, const float* gt, int Hg, int Wg, float min_depth, float max_depth, int eigen_crop,
                   float stereo_scale, void* workspace, size_t workspace_bytes, float* metrics8, void* stream);

/* ---- GPU input pipeline (csrc/input.cu; datasets/mono_dataset.py:102-184, 206-238) ------------------------------------------------
 * From a batch of uint8 frames already resized to the network resolution, frames[B, F, H, W, 3] (HWC), produce what
 * MonoDataset.__getitem__ / preprocess hand to the trainer: color[f] = ToTensor(frame f) and color_aug[f] = ColorJitter(frame f),
 * each fp32 [B, 3, H, W], with the item's horizontal flip applied to both (mono_dataset.py:224-226) and ONE jitter parameter set per
 * item for all of its frames (mono_dataset.py:228-233).  prm_f[B,4] = brightness, contrast, saturation, hue factors;
 * prm_i[B,6] = the order torchvision drew (4 entries: 0 brightness, 1 contrast, 2 saturation, 3 hue), do_color_aug, do_flip.
 * color_dev / color_aug_dev: DEVICE arrays of F output pointers.  Arithmetic = torchvision's tensor kernels in fp32 (the reference's
 * PIL path rounds to 8 bits after every operation: agreement to ~1/255 per operation).  workspace:
 * mvf_input_pipeline_workspace_floats(B, F) floats.  A step's host-to-device copy shrinks from six fp32 tensors to the 8-bit frames. */
size_t mvf_input_pipeline_workspace_floats(int B, int F);
int mvf_input_pipeline(const unsigned char* frames, const float* prm_f, const int* prm_i, float* workspace, size_t workspace_floats,
                       float* const* color_dev, float* const* color_aug_dev, int B, int F, int H, int W, void* stream);

/* ---- disparity head (csrc/dispconv.cu): Conv3x3(C, 1) on the padded channels-last decoder feature (networks/monodepth2.py:76-77, 94;
 * LiteMono.py:468-469, 502).  xp: dense channels-last [B, H+2, W+2, C] (the reflection-padded input of mvf_upcat_pad_fwd), C % 4 == 0,
 * C <= 64; w: the module's [1, C, 3, 3] weight; y / grad_y: [B, H, W].  HBM-bound direct kernels instead of an N = 16 tensor-core
 * tile with 15 zero columns.  wgrad also returns the bias gradient (grad_b may be NULL); workspace:
 * mvf_dispconv_wgrad_workspace_floats(B*H*W, C) floats, partials added in a fixed order. */
int mvf_dispconv_fwd(const float* xp, const float* w, const float* bias, float* y, int B, int C, int H, int W, void* stream);
int mvf_dispconv_dgrad(const float* grad_y, const float* w, float* grad_xp, int B, int C, int H, int W, void* stream);
size_t mvf_dispconv_wgrad_workspace_floats(long long P, int C);
int mvf_dispconv_wgrad(const float* xp, const float* grad_y, float* grad_w, float* grad_b, float* workspace, size_t workspace_floats, int B,
                       int C, int H, int W, void* stream);

/* ---- fused nearest-upsample x2 + channel concat + ReflectionPad2d(1), channels-last ---------------------------------
 * y[B,Ca+Cs,H+2,W+2] = pad(cat(upsample ? up2(a[B,Ca,H/2,W/2]) : a[B,Ca,H,W], skip[B,Cs,H,W])): the data movement
 * the reference does with F.interpolate + torch.cat + nn.ReflectionPad2d(1) before each decoder convolution
 * (monodepth2.py:86-93, layers.py:126-139), in one pass; bwd is the exact adjoint.  Dense channels-last buffers
 * (element (b,c,y,x) at ((b*H + y)*W + x)*C + c), channel counts % 4 == 0, H, W >= 4.  skip may be NULL (Cs = 0). */
int mvf_upcat_pad_fwd(const float* a, const float* skip, float* y, int B, int Ca, int Cs, int H, int W, int upsample, void* stream);
int mvf_upcat_pad_bwd(const float* grad_y, float* grad_a, float* grad_skip, int B, int Ca, int Cs, int H, int W, int upsample,
                      void* stream);

/* MaxPool2d(kernel 3, stride 2, padding 1) of the ResNet stems (torchvision ResNet.maxpool, used by monodepth2.py:39 and
 * posenet.py:91) on dense channels-last tensors; idx[B,Ho,Wo,C] (uint8) holds the in-window position of each maximum and
 * feeds the gather-form (atomic-free, deterministic) backward.  C % 4 == 0; Ho = (H-1)/2+1. */
int mvf_maxpool3s2_fwd(const float* x, float* y, unsigned char* idx, int B, int C, int H, int W, void* stream);
int mvf_maxpool3s2_bwd(const float* grad_y, const unsigned char* idx, float* grad_x, int B, int C, int H, int W, void* stream);

/* ---- feature warp ("F2"), bilinear resize, PReLU tail, pose matrices (csrc/warp_cl.cu) -------------------------------------
 * layout: 0 = dense NCHW (any C), 1 = dense channels-last with C % 2 == 0 (element (b,c,y,x) at ((b*H + y)*W + x)*C + c;
 * float4 accesses when C % 4 == 0, float2 otherwise).
 *
 * mvf_flow_warp_*: IFRNet.warp (networks/IFRNet.py:7-15, used by IFRNet.forward :400-441 and FusionModule.warp_features,
 *   fusion_module.py:78-90): y[b,:,v,u] = bilinear sample of x[b] at (u + flow[b,0,v,u], v + flow[b,1,v,u]), border padding,
 *   align_corners=True; flow is NCHW [B,2,H,W] in pixels.  bwd (channels-last only) gives grad_x; it replaces ATen's float-atomic
 *   grid_sampler_2d_backward by a 64-bit fixed-point scatter (order-independent, bitwise deterministic); workspace:
 *   mvf_flow_warp_bwd_workspace_bytes(B, C, H, W) bytes of scratch (zeroed by the call).
 * mvf_resize_bilinear_*: F.interpolate(mode="bilinear", align_corners=...) (hrnet_encoder.py:275-280, IFRNet.py:118,383-423,
 *   fusion_module.py:68-99, layers.py:225-228 as used by LiteMono.py:495,502); scale_h / scale_w are torch's source-index scales
 *   (in/out, 1/scale_factor, or (in-1)/(out-1) when align_corners); channel c of an NCHW tensor is multiplied by mul_even / mul_odd
 *   by parity (the flow rescaling of fusion_module.py:86-87, IFRNet.py:417-421); bwd is the exact adjoint in gather form.
 * mvf_prelu_cl_fwd: y = PReLU_C(x + res) (res may be NULL), channels-last [P pixels][C]  (IFRNet.py:121-150).
 * mvf_pose_matrix_*: transformation_from_parameters (layers.py:28-103): axisangle[B,3], translation[B,3] -> M[B,4,4]. */
int mvf_flow_warp_fwd(const float* x, const float* flow, float* y, int B, int C, int H, int W, int layout, void* stream);
size_t mvf_flow_warp_bwd_workspace_bytes(int B, int C, int H, int W);
int mvf_flow_warp_bwd(const float* grad_y, const float* flow, float* grad_x, int B, int C, int H, int W, void* workspace,
                      size_t workspace_bytes, void* stream);
int mvf_resize_bilinear_fwd(const float* x, float* y, int B, int C, int Hin, int Win, int Hout, int Wout, float scale_h, float scale_w,
                            int align_corners, float mul_even, float mul_odd, int layout, void* stream);
int mvf_resize_bilinear_bwd(const float* grad_y, float* grad_x, int B, int C, int Hin, int Win, int Hout, int Wout, float scale_h,
                            float scale_w, int align_corners, float mul_even, float mul_odd, int layout, void* stream);
int mvf_prelu_cl_fwd(const float* x, const float* res, const float* slope, float* y, long long P, int C, void* stream);
int mvf_pose_matrix_fwd(const float* axisangle, const float* translation, float* M, int B, int invert, void* stream);
int mvf_pose_matrix_bwd(const float* axisangle, const float* translation, const float* grad_M, float* grad_axisangle,
                        float* grad_translation, int B, int invert, void* stream);

/* ---- Lite-Mono block kernels (csrc/litemono.cu); dense channels-last tensors [P pixels][C], C % 4 == 0 -------------------------
 * mvf_dwconv3x3_*: depth-wise dilated 3x3 convolution, stride 1, zero padding = dilation (CDilated, networks/LiteMono.py:140-155,
 *   called by DilatedConv.forward :179-201).  w_taps is the filter in tap-major order [9][C] (tap = kh*3 + kw); flip = 1 mirrors the
 *   taps (= the data gradient: call it on grad_y).  wgrad writes grad_w in the module's [C,1,3,3] order; workspace:
 *   mvf_dwconv3x3_wgrad_workspace_floats(P, C) floats; per-CTA partials added in a fixed order (reproducible).
 * mvf_gelu_*: exact (erf) GELU of nn.GELU (LiteMono.py:127,172,216), backward from the saved input; n % 4 == 0.
 * mvf_layernorm_cl_*: LayerNorm over C of channels-last tokens (LiteMono.py:93-121, data_format="channels_last"), eps inside the
 *   square root; fwd keeps mean / rstd [P] for the backward; bwd gives grad_x, grad_weight, grad_bias (fixed-order partials).  C <= 512. */
int mvf_dwconv3x3_fwd(const float* x, const float* w_taps, const float* bias, float* y, int B, int C, int H, int W, int dilation, int flip,
                      void* stream);
size_t mvf_dwconv3x3_wgrad_workspace_floats(long long P, int C);
int mvf_dwconv3x3_wgrad(const float* x, const float* grad_y, float* grad_w, float* workspace, size_t workspace_floats, int B, int C, int H,
                        int W, int dilation, void* stream);
int mvf_gelu_fwd(const float* x, float* y, long long n, void* stream);
int mvf_gelu_bwd(const float* x, const float* grad_y, float* grad_x, long long n, void* stream);
int mvf_layernorm_cl_fwd(const float* x, const float* weight, const float* bias, float* y, float* mean, float* rstd, long long P, int C,
                         float eps, void* stream);
size_t mvf_layernorm_bwd_workspace_floats(long long P, int C);
int mvf_layernorm_cl_bwd(const float* x, const float* grad_y, const float* weight, const float* mean, const float* rstd, float* grad_x,
                         float* grad_weight, float* grad_bias, float* workspace, size_t workspace_floats, long long P, int C, void* stream);

/* ---- fused training-mode BatchNorm2d (+ residual add) + ReLU on dense channels-last tensors [P pixels][C] -----------
 * The bn -> (+= identity) -> relu tail of torchvision's BasicBlock / Bottleneck (networks/monodepth2.py:16-31,
 * networks/posenet.py:10-52, hrnet_encoder.py:58-139).  fwd: batch statistics (biased variance for the normalisation),
 * running statistics updated as nn.BatchNorm2d does (momentum, unbiased variance, the int64 num_batches_tracked counter
 * incremented on the device; pass NULL to skip), save_mean /
 * save_invstd [C] kept for the backward.  bwd: grad_x, grad_gamma, grad_beta and (optional) grad_identity = masked grad_y.
 * identity / y may be NULL (no residual / relu = 0).  workspace: mvf_bn_workspace_floats(P, C) floats of scratch; the
 * per-CTA partial sums are added in a fixed order (bitwise reproducible).  C <= 1024; C % 4 == 0, or C % 4 == 2 with an even P (HRNet's
 * 18-channel branch): two pixels are then processed as one row of 2C channels and save_mean / save_invstd must hold 2C floats. */
size_t mvf_bn_workspace_floats(long long P, int C);
int mvf_bn_relu_fwd(const float* x, const float* identity, float* y, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, long long* num_batches_tracked, float* save_mean, float* save_invstd, float* workspace,
                    size_t workspace_floats, long long P, int C, float eps, float momentum, int relu, void* stream);
int mvf_bn_relu_bwd(const float* x, const float* grad_y, const float* y, const float* gamma, const float* save_mean,
                    const float* save_invstd, float* grad_x, float* grad_identity, float* grad_gamma, float* grad_beta,
                    float* workspace, size_t workspace_floats, long long P, int C, int relu, void* stream);

/* ---- SyncBatchNorm (train.py:205-208: nn.SyncBatchNorm.convert_sync_batchnorm under data parallelism) -----------------
 * The same kernels split around a cross-rank SUM of `sums`, [2C + 1] doubles {sum_0[C], sum_1[C], pixel count}:
 *   forward : mvf_bn_sync_stats_fwd (this rank's sum x, sum x^2, count) -> all-reduce(sum) -> mvf_bn_sync_apply_fwd
 *             (mean / biased variance over the GLOBAL count, running statistics with the global unbiased variance, apply);
 *   backward: mvf_bn_sync_stats_bwd (this rank's sum g, sum g*xhat; grad_beta / grad_gamma are these LOCAL sums, as in
 *             torch.nn.SyncBatchNorm -- they are averaged with every other parameter gradient) -> all-reduce(sum) ->
 *             mvf_bn_sync_apply_bwd (input gradient with the two means taken over the global count).
 * The exchange itself is the caller's (torch.distributed all_reduce on the float64 vector, NCCL over NVLink).
 * scratch of mvf_bn_sync_apply_bwd: 2C + 4 floats. */
int mvf_bn_sync_stats_fwd(const float* x, double* sums, float* workspace, size_t workspace_floats, long long P, int C, void* stream);
int mvf_bn_sync_apply_fwd(const float* x, const float* identity, float* y, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, long long* num_batches_tracked, float* save_mean, float* save_invstd,
                          const double* global_sums, long long P, int C, float eps, float momentum, int relu, void* stream);
int mvf_bn_sync_stats_bwd(const float* x, const float* grad_y, const float* y, const float* save_mean, const float* save_invstd,
                          double* sums, float* grad_gamma, float* grad_beta, float* workspace, size_t workspace_floats, long long P,
                          int C, int relu, void* stream);
int mvf_bn_sync_apply_bwd(const float* x, const float* grad_y, const float* y, const float* gamma, const float* save_mean,
                          const float* save_invstd, float* grad_x, float* grad_identity, const double* global_sums, float* scratch,
                          size_t scratch_floats, long long P, int C, int relu, void* stream);

/* Backward of the activation fused into a convolution's epilogue plus the bias gradient, one pass over dense
 * channels-last [P pixels][C]: grad_pre = grad_y * act'(y) (act 0 none, 1 relu, 2 elu(alpha=1), from the saved OUTPUT y)
 * and grad_bias[c] = sum_p grad_pre[p][c] (fixed-order, reproducible).  Replaces elu_backward / threshold_backward +
 * `grad.sum((0,2,3))` behind layers.py:68-117 (ConvBlock / Conv3x3 with bias).  grad_bias may be NULL (activation only);
 * grad_pre may be NULL when act == 0 (bias gradient only).  workspace: mvf_bn_workspace_floats(P, C). */
int mvf_act_bwd_bias(const float* grad_y, const float* y, float* grad_pre, float* grad_bias, float* workspace, size_t workspace_floats,
                     long long P, int C, int act, void* stream);

/* ---- fused torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW.step over ONE flat fp32 arena (train.py:661-666) --------
 * params / grads / exp_avg / exp_avg_sq: n floats each, 16-byte aligned.
 * state[4] (device): {step count, last gradient norm, learning rate, gradient scale}.  The step count is read and
 * incremented on the device and the learning rate / scale are read there, so the call can be recorded into a CUDA graph
 * and still follow a scheduler (train.py:239-242) -- the host writes state[2] before a replay.  The gradient scale
 * (1 / world size after a sum-all-reduce) is applied before the norm.  max_norm <= 0 disables clipping.
 * n_duplicated: the first n_duplicated floats (multiple of 4) belong to parameters the reference lists twice in its
 * optimizer (train.py:198-200, shared encoder): their norm counts twice, their gradient is clipped twice and they take two
 * AdamW updates per call (their own step counter at 2t - 1 and 2t), as torch.optim.AdamW + clip_grad_norm_ do to a
 * duplicated list entry.  skip_groups (may be NULL): one byte per 4-float group, non-zero = parameter without a gradient
 * this step, left untouched (torch skips `grad is None`).  workspace: mvf_adamw_workspace_bytes() bytes of scratch
 * (gradient-norm partials, added in a fixed order).  Same arithmetic as torch's (decoupled weight decay, bias-corrected
 * moments, eps outside the square root). */
size_t mvf_adamw_workspace_bytes(void);
int mvf_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, long long n_duplicated,
                   const unsigned char* skip_groups, float* state, void* workspace, size_t workspace_bytes, float beta1, float beta2,
                   float eps, float weight_decay, float max_norm, void* stream);

/* Gathers the per-parameter gradient tensors autograd produced into the flat arena mvf_adamw_step reads (and the one
 * data-parallel all-reduce runs on): grads[i] (device pointer, host array; NULL = no gradient this step, slice zeroed)
 * -> arena[offsets[i] .. offsets[i] + sizes[i]).  Offsets must be multiples of 4 floats.  One launch per 128 tensors;
 * replaces the ~200 `param.grad += g` launches of torch's AccumulateGrad + DDP's bucket copies (train.py:205-208). */
int mvf_gather_grads(float* arena, const void* const* grads, const long long* offsets, const long long* sizes, int n_tensors,
                     void* stream);

/* Id of the CUDA-graph capture `stream` is part of, 0 when it is not capturing.  The host side keys its per-step caches
 * (packed filter banks) on it so that a recorded graph never depends on buffers produced outside its own capture. */
unsigned long long mvf_stream_capture_id(void* stream);

/* device self-test: q_sequence[i] = the kernels' shared-reciprocal division of a[i] by b[i], q_ieee[i] = the
 * IEEE quotient (div.rn.f32); the two must be bit-identical for operands in the normal range. */
int mvf_selftest_division(const float* a, const float* b, float* q_sequence, float* q_ieee, size_t n, void* stream);

/* device self-test of the tcgen05 plumbing: D[128,N] = A . B^T on one CTA (TF32 in, fp32 out).  A is [160][K] (first 128 rows used;
 * a_mn_major = 0) or [K][128] (a_mn_major = 1, the layout NCHW activations have); B is [N][K].
 * K % 32 == 0, N % 16 == 0, 16 <= N <= 256. */
int mvf_selftest_umma(const float* A, const float* B, float* D, int N, int K, int a_mn_major, void* stream);
/* same, K-major A given as [160][K]: D = A[row_off : row_off+128] . B^T, i.e. an operand whose start address is not
 * aligned to the 1024-byte swizzle pattern (base_off_mode 1 sets the descriptor's base-offset field, 0 leaves it 0). */
int mvf_selftest_umma_rows(const float* A, const float* B, float* D, int N, int K, int row_off, int base_off_mode, void* stream);
/* development aid: when non-NULL, CTA (0,0) of mvf_conv2d_forward dumps its pipeline stage 0 there */
void mvf_conv2d_debug_buffer(float* p);

#ifdef __cplusplus
}
#endif
#endif
