/*
 * epn_b200.h -- C ABI of libepn_b200.so, the B200 (sm_100a) engine for the
 * SE(3) separable point-convolution hot path of nintendops/EPN_PointCloud.
 *
 * Drop-in boundary: every entry point replaces one pybind function of the
 * reference's `vgtk.cuda.*` extensions, or one PyTorch op chain of
 * `vgtk.so3conv` / `vgtk.spconv` (file:line cited per function; paths are
 * relative to the reference root, commit b625483).
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller; the library never
 *     allocates, frees or keeps device memory.  No call leaves per-call state
 *     behind: what a forward must tell its backward travels through the caller
 *     (`grouped` buffer + `grouped_layout` word).  Process-wide state is limited
 *     to (a) a thread-local error string, (b) a launch counter and the optional
 *     profiling record, (c) the CONFIGURATION KNOBS at the end of this header
 *     (epn_set_gemm_backend / epn_set_slab_bytes / epn_set_fused_inter /
 *     epn_set_fused_inter_bwd and their EPN_* environment defaults): atomics read once at the start of every call,
 *     shared by all threads and devices of the process.  Calls are re-entrant and
 *     thread-safe; changing a knob while another thread is inside a call is
 *     allowed and affects only calls that start afterwards;
 *   - the per-device kernel attributes (dynamic shared memory limits) are set
 *     lazily on every device the library is used on (one process may drive
 *     several GPUs, as the reference's nn.DataParallel does,
 *     vgtk/vgtk/app/trainer.py:153-160);
 *   - tensors are dense, row-major, in the reference's layouts:
 *       xyz [B,3,P]   feats [B,C,P,A] (A innermost)   idx [B,P,K] int32
 *       inter_w [B,P,A,KS,K]   grouped [B,C,KS,P,A]   W [C_out, C_in*KS]
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *     every launch goes to that stream; nothing synchronises;
 *   - return value 0 = success; >0 = cudaError_t of the failed launch;
 *     <0 = argument check failed (EPN_ERR_*); epn_last_error() describes it.
 *     No exception ever crosses the boundary.
 *   - fp32 only (the reference's double dispatch is not on the model path).
 */
#ifndef EPN_B200_H_
#define EPN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EPN_B200_VERSION 101 /* major*1000 + minor */

#define EPN_ERR_NULL (-1)     /* required pointer is NULL              */
#define EPN_ERR_SHAPE (-2)    /* non-positive or unsupported extent    */
#define EPN_ERR_WORKSPACE (-3) /* workspace too small / missing        */
#define EPN_ERR_ALIGN (-4)    /* pointer not aligned as required       */

int epn_version(void);
/* Thread-local, valid until the next failing call on the same thread. */
const char *epn_last_error(void);
/* 1 if the library holds a kernel image for the current device (sm_100). */
int epn_device_supported(void);
/* Number of kernels this library has launched in this process (all threads). */
unsigned long long epn_launch_count(void);
/* Optional per-kernel-class device timing (bench.py's roofline pass; off by default).
 * While enabled every launch site is bracketed by a cudaEvent pair on its stream.
 * epn_profile_read synchronises on the recorded events, sums their elapsed ms per class
 * into ms_per_class[0..n_class) / scopes_per_class[0..n_class) and clears the record.
 * Classes: 0 index ops, 1 inter grouping fwd, 2 inter grouping bwd (scatter),
 * 3 intra grouping, 4 channel GEMMs, 5 other. */
void epn_profile_enable(int on);
int epn_profile_read(double *ms_per_class, long long *scopes_per_class, int n_class);

/* ------------------------------------------------------------------ grouping
 * vgtk.cuda.grouping.ball_query  (vgtk/vgtk/cuda/grouping_cuda.cpp:71-86,
 * kernel grouping_cuda_kernel.cu:67-113).  new_xyz [b,3,m], xyz [b,3,n] ->
 * idx [b,m,nsample]: first `nsample` support indices (ascending) with
 * d2 < radius^2 (fp32, FMUL/FFMA/FFMA), cyclic repeat fill if fewer than
 * nsample-1 hits, a single trailing 0 if exactly nsample-1.  Bit-exact.
 * Every output slot is written (no pre-zeroing needed). */
int epn_ball_query_f32(const float *new_xyz, const float *xyz, int32_t *idx, int b, int n, int m,
                       float radius, int nsample, void *stream);

/* vgtk.cuda.grouping.furthest_point_sampling (grouping_cuda.cpp:160-174,
 * kernel grouping_cuda_kernel.cu:340-466).  xyz [b,3,n] -> idx [b,m], start at
 * index 0, points with |p|^2 <= 1e-3 never selected, the reference's
 * thread/tree tie-breaking reproduced.  Bit-exact.
 * temp: workspace of epn_fps_workspace_bytes(b,n) bytes (may be NULL when that
 * returns 0). */
size_t epn_fps_workspace_bytes(int b, int n);
int epn_fps_f32(const float *xyz, void *temp, int32_t *idx, int b, int n, int m, void *stream);

/* ----------------------------------------------------------------- gathering
 * vgtk.cuda.gathering.gather_points_forward / _backward
 * (gathering_cuda.cpp:29-65, kernels gathering_cuda_kernel.cu:42-98).
 * points [b,c,n], idx [b,m] -> out [b,c,m];  backward accumulates into
 * grad_points [b,c,n] which the CALLER must have zeroed. */
int epn_gather_fwd_f32(const float *points, const int32_t *idx, float *out, int b, int c, int n,
                       int m, void *stream);
int epn_gather_bwd_f32(const float *grad_out, const int32_t *idx, float *grad_points, int b, int c,
                       int n, int m, void *stream);

/* ------------------------------------------------------------ zpconv surface
 * vgtk.cuda.zpconv.* (zpconv_cuda.cpp:41-118, kernels zpconv_cuda_kernel.cu:32-195).
 * inter: nbr/w [b,np,na,ks,ann], feats [b,c,nq,na] -> out [b,c,ks,np,na]
 * intra: nbr [na_out,ann], w [na_out,ks,ann], feats [b,c,np,na_in] -> out [b,c,ks,np,na_out]
 * Forward outputs are fully written.  Backward outputs are accumulated with
 * fp32 atomics: the CALLER zeroes dfeats first. */
int epn_zp_inter_fwd_f32(const int32_t *nbr, const float *w, const float *feats, float *out, int b,
                         int c, int nq, int np, int na, int ks, int ann, void *stream);
int epn_zp_inter_bwd_f32(const int32_t *nbr, const float *w, const float *dout, float *dfeats, int b,
                         int c, int nq, int np, int na, int ks, int ann, void *stream);
int epn_zp_intra_fwd_f32(const int32_t *nbr, const float *w, const float *feats, float *out, int b,
                         int c, int np, int na_in, int na_out, int ks, int ann, void *stream);
int epn_zp_intra_bwd_f32(const int32_t *nbr, const float *w, const float *dout, float *dfeats, int b,
                         int c, int np, int na_in, int na_out, int ks, int ann, void *stream);

/* ------------------------------------------------- live-path grouping stages
 * inter_so3conv_grouping_anchor (vgtk/vgtk/so3conv/functional.py:180-218):
 *   inter_w[b,p,a,k,n] = relu(1 - |xyz[b,:,idx[b,p,n]] - centers[b,:,p] - anchors[a]@kernels[k]|^2 / sigma)
 * xyz [b,3,p_in], centers [b,3,p], idx [b,p,nn], anchors [na,3,3], kernels [ks,3]. */
int epn_inter_weights_f32(const float *xyz, const float *centers, const int32_t *idx,
                          const float *anchors, const float *kernels, float sigma, float *inter_w,
                          int b, int p_in, int p, int nn, int na, int ks, void *stream);

/* inter_zpconv_grouping_naive (vgtk/vgtk/spconv/functional.py:372-390):
 *   out[b,c,k,p,a] = sum_n feats[b,c,idx[b,p,n],a] * inter_w[b,p,a,k,n]
 * If inter_w == NULL the weights are recomputed on the fly from
 * (xyz, centers, anchors, kernels, sigma) and never touch HBM. */
int epn_inter_group_fwd_f32(const float *feats, const int32_t *idx, const float *inter_w,
                            const float *xyz, const float *centers, const float *anchors,
                            const float *kernels, float sigma, float *out, int b, int c, int p_in,
                            int p, int nn, int na, int ks, void *stream);
/* adjoint w.r.t. feats; dfeats [b,c,p_in,na] must be zeroed by the caller. */
int epn_inter_group_bwd_f32(const float *dout, const int32_t *idx, const float *inter_w,
                            const float *xyz, const float *centers, const float *anchors,
                            const float *kernels, float sigma, float *dfeats, int b, int c,
                            int p_in, int p, int nn, int na, int ks, void *stream);

/* intra_so3conv_grouping (vgtk/vgtk/so3conv/functional.py:221-268):
 *   out[b,c,k,p,a] = feats[b,c,p,intra_idx[a,k]];  intra_idx [na,kn] int32.
 * Backward: dfeats [b,c,p,na] fully written (no atomics, no pre-zeroing);
 * requires every column of intra_idx to be a permutation (true for the
 * icosahedral index) -- otherwise use epn_zp_intra_bwd_f32. */
int epn_intra_group_fwd_f32(const float *feats, const int32_t *intra_idx, float *out, int b, int c,
                            int p, int na, int kn, void *stream);
int epn_intra_group_bwd_f32(const float *dout, const int32_t *intra_idx, float *dfeats, int b, int c,
                            int p, int na, int kn, void *stream);

/* ------------------------------------------------------------- fused convs
 * InterSO3Conv.forward minus sampling/ball query
 * (vgtk/vgtk/so3conv/modules.py:157-174 = so3conv/functional.py:174-176 +
 * modules.py:48-55):
 *   out[b,o,p,a] = sum_{c,k} W[o,c*ks+k] * sum_n w(b,p,a,k,n) * feats[b,c,idx[b,p,n],a]
 * feats == NULL means feats == 1 with c_in == 1 (layer 0: occupancy features,
 * so3conv/functional.py:25-44).  inter_w is never materialised.
 * workspace: epn_inter_so3conv_workspace_bytes(...) bytes, 256-B aligned.
 *
 * grouped (optional, NULL = off): what autograd would keep of the grouped tensor for the weight gradient
 * (the reference keeps the whole [b,c,ks,p,na] fp32 tensor, so3conv/modules.py:48-55 under autograd).
 * A training forward may pass a buffer of exactly epn_*_grouped_bytes(...) bytes (256-B aligned); the
 * tensor-core operand tiles of the grouped tensor are then written there instead of into the workspace, and
 * a backward given the same buffer computes dW from them without re-running the spatial contraction.
 * epn_*_grouped_bytes returns 0 when the shape / backend cannot do that (then pass NULL).
 * grouped_layout: a forward given `grouped` stores one word describing how it laid the tiles out (order of the
 * K dimension, slab plan); the caller hands that word to the backward together with the buffer, so the pair
 * stays consistent even if the configuration knobs change between the two calls.  Ignored when grouped == NULL. */
size_t epn_inter_so3conv_workspace_bytes(int b, int c_in, int c_out, int p_in, int p, int nn, int na,
                                         int ks, int backward);
size_t epn_inter_so3conv_grouped_bytes(int b, int c_in, int p, int nn, int na, int ks);
int epn_inter_so3conv_fwd_f32(const float *feats, const float *xyz, const float *centers,
                              const int32_t *idx, const float *anchors, const float *kernels,
                              float sigma, const float *W, float *out, void *workspace,
                              size_t workspace_bytes, void *grouped, size_t grouped_bytes,
                              unsigned long long *grouped_layout, int b, int c_in, int c_out, int p_in, int p,
                              int nn, int na, int ks, void *stream);
/* dout [b,c_out,p,na] -> dfeats [b,c_in,p_in,na] (NULL to skip; else fully
 * written) and dW [c_out,c_in*ks] (NULL to skip; else fully written). */
int epn_inter_so3conv_bwd_f32(const float *dout, const float *feats, const float *xyz,
                              const float *centers, const int32_t *idx, const float *anchors,
                              const float *kernels, float sigma, const float *W, float *dfeats,
                              float *dW, void *workspace, size_t workspace_bytes, const void *grouped,
                              size_t grouped_bytes, unsigned long long grouped_layout, int b, int c_in,
                              int c_out, int p_in, int p, int nn, int na, int ks, void *stream);

/* IntraSO3Conv.forward (vgtk/vgtk/so3conv/modules.py:197-200):
 *   out[b,o,p,a] = sum_{c,k} W[o,c*kn+k] * feats[b,c,p,intra_idx[a,k]] */
size_t epn_intra_so3conv_workspace_bytes(int b, int c_in, int c_out, int p, int na, int kn,
                                         int backward);
size_t epn_intra_so3conv_grouped_bytes(int b, int c_in, int p, int na, int kn);
int epn_intra_so3conv_fwd_f32(const float *feats, const int32_t *intra_idx, const float *W, float *out,
                              void *workspace, size_t workspace_bytes, void *grouped, size_t grouped_bytes,
                              unsigned long long *grouped_layout, int b, int c_in, int c_out, int p, int na,
                              int kn, void *stream);
/* IntraSO3ConvBlock fed by an Inter/IntraSO3ConvBlock WITHOUT materialising the normalised activation
 * (SPConvNets/utils/base_so3conv.py:116-126 followed by :52-62): `x` [b,c_in,p,na] is the RAW output of the preceding
 * convolution, `stats` its (mean, rstd) per group from epn_norm_stats_f32 (or, for an evaluation-mode BatchNorm, the
 * running statistics), and feats = leaky_relu(norm(x) * gamma + beta, slope) is applied as the operand tiles are built.
 * Everything else as epn_intra_so3conv_fwd_f32; shapes outside the tile routes return EPN_ERR_SHAPE. */
int epn_intra_so3conv_fwd_norm_f32(const float *x, const float *stats, const float *gamma, const float *beta,
                                   int norm_mode, float slope, const int32_t *intra_idx, const float *W, float *out,
                                   void *workspace, size_t workspace_bytes, void *grouped, size_t grouped_bytes,
                                   unsigned long long *grouped_layout, int b, int c_in, int c_out, int p, int na,
                                   int kn, void *stream);
/* feats may be NULL when grouped is given (dW then needs nothing else). */
int epn_intra_so3conv_bwd_f32(const float *dout, const float *feats, const int32_t *intra_idx,
                              const float *W, float *dfeats, float *dW, void *workspace,
                              size_t workspace_bytes, const void *grouped, size_t grouped_bytes,
                              unsigned long long grouped_layout, int b, int c_in, int c_out, int p, int na,
                              int kn, void *stream);

/* BasicSO3Conv.forward (vgtk/vgtk/so3conv/modules.py:48-55) on an already
 * grouped tensor: x [b, ck, pa], W [co, ck] -> out [b, co, pa].
 * workspace: epn_basic_conv_workspace_bytes(...) bytes, 256-B aligned (operand tiles). */
size_t epn_basic_conv_workspace_bytes(int b, int ck, int co, int pa);
int epn_basic_conv_fwd_f32(const float *x, const float *W, float *out, void *workspace,
                           size_t workspace_bytes, int b, int ck, int co, int pa, void *stream);
/* dx [b,ck,pa] (NULL to skip), dW [co,ck] (NULL to skip; fully written). */
int epn_basic_conv_bwd_f32(const float *dout, const float *x, const float *W, float *dx, float *dW,
                           void *workspace, size_t workspace_bytes, int b, int ck, int co, int pa,
                           void *stream);

/* ------------------------------------------------ fused norm + activation (block wrappers)
 * y = leaky_relu(norm(x) * gamma + beta, slope) on x [b, c, n] (n = points*anchors):
 *   mode 0: InstanceNorm2d(affine=False) -> leaky_relu  (SPConvNets/utils/base_so3conv.py:43,55-57)
 *   mode 1: BatchNorm2d, batch statistics, affine -> leaky_relu  (base_so3conv.py:107,119-125,193,209)
 *   mode 2 (forward only): BatchNorm2d in evaluation mode -- stats [2*c] is an INPUT holding the per-channel
 *           (mean, 1/sqrt(running_var + eps)); one apply pass, no statistics kernels
 * gamma/beta [c] may be NULL (= 1 / 0).  stats [2*G] receives (mean, rstd) per group, G = b*c (mode 0)
 * or c (mode 1); the backward needs it, and the caller derives BatchNorm's running statistics from it.
 * residual [b, c, n] (NULL = none): the skip connection of SeparableSO3ConvBlock fused into the same pass,
 *   y = leaky_relu(...) + residual   (base_so3conv.py:209-211); its gradient is dy itself.
 * Backward writes dx fully and, for mode 1, dgamma/dbeta [c] (NULL to skip).
 * workspace: epn_norm_act_workspace_bytes(b, c) bytes. */
size_t epn_norm_act_workspace_bytes(int b, int c);
/* statistics only (mode 0 / 1; same workspace): for consumers that apply the normalisation while they load */
int epn_norm_stats_f32(const float *x, float *stats, void *workspace, size_t workspace_bytes, int b, int c, int n,
                       int mode, float eps, void *stream);
int epn_norm_act_fwd_f32(const float *x, const float *gamma, const float *beta, const float *residual, float *y,
                         float *stats, void *workspace, size_t workspace_bytes, int b, int c, int n, int mode,
                         float eps, float slope, void *stream);
/* Running-statistics bookkeeping of a training-mode BatchNorm2d (what nn.BatchNorm2d.forward does next to the
 * normalisation) from the (mean, rstd) [2*c] of epn_norm_act_fwd_f32 / epn_norm_stats_f32 (mode 1), as one launch:
 *   mean += bias (a per-channel constant that was left out of the kernel's input, NULL = none);
 *   var = (1 / rstd^2 - eps) * count / max(count - 1, 1);   num_batches_tracked (int64, device) += 1;
 *   running = running * (1 - momentum) + momentum * batch    (momentum < 0: cumulative average, factor 1 / num_batches_tracked). */
int epn_bn_track_f32(const float *stats, const float *bias, float *running_mean, float *running_var,
                     long long *num_batches_tracked, int c, long long count, float momentum, float eps, void *stream);
int epn_norm_act_bwd_f32(const float *dy, const float *x, const float *gamma, const float *beta,
                         const float *stats, float *dx, float *dgamma, float *dbeta, void *workspace,
                         size_t workspace_bytes, int b, int c, int n, int mode, float slope, void *stream);

/* ------------------------------------------------ configuration knobs (PROCESS-GLOBAL, see Conventions)
 * Channel-GEMM engine used by the three convs: 0 = tcgen05 tensor cores with bf16 hi/lo
 * operand splitting (default; fp32-faithful to ~1e-5 relative), 1 = fp32 SIMT GEMM (the
 * in-library cross-check used by the GPU tests; also selectable with EPN_GEMM=simt). */
void epn_set_gemm_backend(int simt);
/* Upper bound of the grouped tensor one slab of a conv call may hold (default 1 GiB, or the
 * EPN_SLAB_BYTES environment variable).  Changes the *_workspace_bytes() / *_grouped_bytes() results. */
void epn_set_slab_bytes(size_t bytes);
size_t epn_get_slab_bytes(void);
int epn_get_gemm_backend(void);
/* InterSO3Conv forward schedule on the tensor-core engine: 1 (default) = ONE fused kernel (gather + kernel weights +
 * spatial contraction feeding the channel GEMM from shared memory; replaces the op chain
 * vgtk/vgtk/so3conv/functional.py:118-218 -> spconv/functional.py:361-390 -> so3conv/modules.py:48-55);
 * 0 (or EPN_FUSED=0) = grouping kernel writing operand tiles followed by the GEMM kernel.  Same results (tests). */
void epn_set_fused_inter(int on);
int epn_get_fused_inter(void);
/* Data gradient of InterSO3Conv (60 anchors, 24 kernel points, C_out a multiple of 64 up to 256).  Levels
 * (EPN_FUSED_BWD=0/1/2): 1 = rows of <= 16 neighbour slots run ONE fused kernel (dG = dout . W^T accumulated in TMEM,
 * transposed spatial contraction + scatter straight from TMEM: the gradient of the grouped tensor never reaches HBM);
 * 2 (default) = rows of 17..32 slots too (two CTAs per point pair, 16 distinct neighbours each: +6 % on the rotation
 * network, neutral on the classification network); 0 = data-gradient GEMM into a slab followed by the scatter kernel.
 * Same results (tests). */
void epn_set_fused_inter_bwd(int on);
int epn_get_fused_inter_bwd(void);

/* ------------------------------------------------ operand format of the forward GEMMs (THREAD-LOCAL)
 * 0 (default) = bf16 hi/lo split operands: fp32's exponent range, ~2^-18 relative per operand.
 * 1           = fp16 hi/lo split operands for the forwards of the CALLING THREAD that keep no operand tiles
 *               (grouped == NULL: inference; BasicSO3Conv forwards always): 22 significand bits at the same speed,
 *               but fp16's RANGE -- activations must satisfy |x| * (neighbours per row) < 65504 (beyond that the
 *               output is NaN, never silently wrong), are resolved to 2^-25 absolute below 2^-3, and weights
 *               need 1e-6 < |W| < 64 to keep their precision (they are scaled by 2^10 internally).  Meant for
 *               callers that know their inputs are normalised activations: the block wrappers of this package set
 *               it for the convs that follow their own norm + activation under torch.no_grad().
 * The setting is per host thread (threads of nn.DataParallel do not disturb each other) and is read at the
 * start of each forward call. */
void epn_set_forward_operands(int fmt);
int epn_get_forward_operands(void);

#ifdef __cplusplus
}
#endif
#endif /* EPN_B200_H_ */
