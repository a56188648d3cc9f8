/*
 * mofa_b200.h — C ABI of the B200-native MoFaNeRF ray-marching engine.
 *
 * The reference (zhuhao-nju/mofanerf) has no FFI: its boundary for this path is the Python class
 * models/render_class.py:40 `myRenderer`.  Each entry point below names the reference function it
 * replaces (file:line relative to the reference tree).  Plain pointers and sizes only; every device
 * pointer is a CUDA device address on the context's device; every call is asynchronous on `stream`
 * (a cudaStream_t passed as void*), performs no hidden synchronisation and no allocation on the hot
 * path (the caller passes the workspace).  All entry points return 0 on success; on failure they
 * return non-zero and mofa_b200_last_error() describes the error (thread-local).
 */
#ifndef MOFA_B200_H
#define MOFA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOFA_B200_ABI_VERSION 2   /* 2: bwd_args.loss_scale_dev, generate_rays, packed-weight blob */

typedef struct mofa_b200_ctx mofa_b200_ctx;

/* flags for mofa_b200_render_args.flags */
#define MOFA_FLAG_LINDISP     1u /* render_rays(lindisp=True)          models/render_class.py:292-295 */
#define MOFA_FLAG_WHITE_BKGD  2u /* raw2outputs(white_bkgd=True)       models/render_class.py:479-480 */
#define MOFA_FLAG_GEMM_SIMT   8u /* verification only: run the dense layers on the plain SIMT CUDA GEMM
                                    instead of the tcgen05 kernel (same fp16 operands, fp32 accumulate) */

/* which network: index into the two MLPs created by tools/create_model_condition.py:23-35 */
#define MOFA_NET_COARSE 0
#define MOFA_NET_FINE   1

int mofa_b200_abi_version(void);
const char* mofa_b200_last_error(void);

/* Lifetime.  `device` is a CUDA ordinal.  Replaces nothing in the reference (it has no handle). */
int mofa_b200_create(mofa_b200_ctx** out, int device);
int mofa_b200_destroy(mofa_b200_ctx* ctx);

/*
 * Upload one network (models/model.py:80-137 `NeRF`) and repack it for the tensor-core kernels:
 * fp16 [out,in] K-major tiles per concat segment, fp32 biases and fp32 latent-column blocks.
 * `tensors` are device fp32 pointers, (weight, bias) pairs in this canonical order (== state_dict
 * order): xyzEncode.linears1.Linear0..3, linear_BiM_xyz.linears1.Linear0..4,
 * linear_BiM_xyz.linears2.Linear0..D-6, linear_uv_xyzBiM.linears1.Linear0..4,
 * linear_uv_xyzBiM.linears2.Linear0..D-6, linear_view_xyBMuv.0, alpha_linear.0, rgb_linear.
 * n_tensors must be 2*(4 + 2*(5 + D-5) + 3).  W must be a multiple of 256, D >= 6.
 */
int mofa_b200_load_weights(mofa_b200_ctx* ctx, int net, int W, int D, const float* const* tensors,
                           int n_tensors, void* stream);

/*
 * Packed-weight blob (the reference re-reads its .tar state_dicts through torch.load and
 * load_state_dict on every start, tools/create_model_condition.py:62-89).  export writes the engine's own layout of a
 * loaded network — fp16 K-major images per concat segment (+ the low images of the split-precision coarse net, + the
 * transposed copies the backward pass uses), fp32 biases, latent columns and heads — into HOST memory; import rebuilds
 * the network from such a blob without any fp32 source tensors (and without the PyTorch modules).  The blob is specific
 * to this library's layout version (checked through its magic) and independent of the device it was made on.
 */
size_t mofa_b200_packed_bytes(mofa_b200_ctx* ctx, int net);
int mofa_b200_export_packed(mofa_b200_ctx* ctx, int net, void* host_dst, size_t bytes, void* stream);
int mofa_b200_import_packed(mofa_b200_ctx* ctx, int net, const void* host_src, size_t bytes, void* stream);

/*
 * Per-call latent conditioning (models/render_class.py:74-85,104): shape code [50], modulated
 * expression code exp_scale*expCodes_Sigma[expType]+exp_bias [30], texture code [256]; device fp32.
 * Folds the latent columns of the five concat layers of each loaded net into per-layer bias vectors.
 */
int mofa_b200_set_latents(mofa_b200_ctx* ctx, const float* shape50, const float* exp_mod30,
                          const float* tex256, void* stream);

typedef struct mofa_b200_render_args {
  uint32_t struct_size;      /* sizeof(mofa_b200_render_args) */
  uint32_t flags;            /* MOFA_FLAG_* */
  const float* rays;         /* [n_rays, ray_stride] fp32: o(3) d(3) near far viewdir(3)  (render_class.py:176-179) */
  int64_t n_rays;
  int32_t ray_stride;        /* floats per ray row, >= 11 */
  int32_t n_samples;         /* N_samples   (coarse samples per ray, 2..256; the reference degenerates at 1) */
  int32_t n_importance;      /* N_importance (0 = single pass) */
  int32_t run_fine;          /* myRenderer.is_run_fineNet (render_class.py:52,321) */
  int32_t fine_net;          /* MOFA_NET_FINE, or MOFA_NET_COARSE when network_fine is None (render_class.py:332) */
  int32_t chunk_rays;        /* rays per internal pass (0 = engine default); results do not depend on it */
  float perturb;             /* > 0: stratified jitter (render_class.py:299-313) */
  float raw_noise_std;       /* > 0: sigma noise (render_class.py:462-468) */
  uint64_t seed;             /* Philox seed used when the explicit random inputs below are NULL */
  const float* t_rand;       /* optional [n_rays, n_samples] uniforms for the jitter */
  const float* u;            /* optional [n_rays, n_importance] uniforms for sample_pdf (perturb > 0) */
  const float* noise_c;      /* optional [n_rays, n_samples] additive sigma noise (already scaled) */
  const float* noise_f;      /* optional [n_rays, n_samples+n_importance] */
  /* outputs (device fp32); rgb/disp/acc are the final maps (fine pass if it ran). NULL = not wanted */
  float* rgb;                /* [n_rays,3] */
  float* disp;               /* [n_rays]   */
  float* acc;                /* [n_rays]   */
  float* rgb0;               /* [n_rays,3] coarse maps, written only when the fine pass runs */
  float* disp0;
  float* acc0;
  float* z_std;              /* [n_rays]   (render_class.py:345) */
  float* raw;                /* [n_rays, S_last, 4] retraw=True (render_class.py:339-340) */
  float* weights;            /* [n_rays, S_last]  compositing weights of the last pass (parity aid) */
  float* z_vals;             /* [n_rays, S_last]  sample depths of the last pass (parity aid) */
  void* workspace;           /* >= mofa_b200_workspace_bytes(...) bytes, 1024-byte aligned */
  size_t workspace_bytes;
} mofa_b200_render_args;

/* Workspace needed by mofa_b200_render_rays_fwd for these sizes (chunk_rays as it will be passed). */
size_t mofa_b200_workspace_bytes(mofa_b200_ctx* ctx, int64_t n_rays, int n_samples, int n_importance,
                                 int chunk_rays);

/*
 * The hot path: myRenderer.batchify_rays -> render_rays (models/render_class.py:111-123, 239-352):
 * stratified sampling, positional encoding, latent-conditioned coarse MLP, raw2outputs (:440-482),
 * sample_pdf (tools/run_nerf_helpers.py:203-247) + sort (:328), fine MLP, raw2outputs.
 */
int mofa_b200_render_rays_fwd(mofa_b200_ctx* ctx, const mofa_b200_render_args* args, void* stream);

/*
 * myRenderer.run_network (models/render_class.py:69-94) + NeRF.forward (models/model.py:121-137):
 * query `net` at n_pts explicit points.  pts [n_pts,3], viewdirs [n_pts,3] (already normalised and
 * expanded per point), raw_out [n_pts,4] = (r,g,b,sigma) pre-activation.  Latents as last set.
 */
size_t mofa_b200_query_workspace_bytes(mofa_b200_ctx* ctx, int64_t n_pts);
int mofa_b200_run_network(mofa_b200_ctx* ctx, int net, const float* pts, const float* viewdirs,
                          int64_t n_pts, float* raw_out, uint32_t flags, void* workspace,
                          size_t workspace_bytes, void* stream);

/* ---- fitting: forward that keeps activations + backward (run_fit.py:305-313) ----
 * Differentiable inputs: ray origins / directions / view directions and the three latent codes; network
 * weights are constants (run_fit.py optimises pose and codes only).  d(disp) is not propagated.
 * The whole ray batch is one pass (no chunking): the train workspace holds every layer's activation. */
size_t mofa_b200_train_workspace_bytes(mofa_b200_ctx* ctx, int64_t n_rays, int n_samples, int n_importance,
                                       int fine_net);
/* Same arguments and outputs as mofa_b200_render_rays_fwd; args->workspace must be a train workspace and must
 * be passed unchanged to mofa_b200_render_rays_bwd. */
int mofa_b200_render_rays_train_fwd(mofa_b200_ctx* ctx, const mofa_b200_render_args* args, void* stream);

typedef struct mofa_b200_bwd_args {
  uint32_t struct_size;      /* sizeof(mofa_b200_bwd_args) */
  uint32_t flags;            /* as in the forward call */
  const float* rays;         /* the forward call's rays */
  int64_t n_rays;
  int32_t ray_stride;
  int32_t n_samples;
  int32_t n_importance;
  int32_t run_fine;
  int32_t fine_net;
  int32_t reserved;
  const float* noise_c;      /* the forward call's explicit sigma noise (or NULL) */
  const float* noise_f;
  const float* d_rgb;        /* [n,3] dL/d rgb_map   (NULL = zero) */
  const float* d_acc;        /* [n]   dL/d acc_map */
  const float* d_rgb0;       /* [n,3] dL/d rgb0      (coarse maps; only when the fine pass ran) */
  const float* d_acc0;
  float loss_scale;          /* power of two applied to the fp16 inter-layer gradients, removed at the outputs */
  float reserved_f;
  float* d_rays;             /* [n,11] out: d/d o(3), d(3), 0, 0, viewdir(3) */
  float* d_shape;            /* [50]  out */
  float* d_expmod;           /* [30]  out: w.r.t. the modulated expression code passed to set_latents */
  float* d_tex;              /* [256] out */
  /* optional (training): fp32 gradient buffers for the network parameters, in the (weight, bias) order
   * of mofa_b200_load_weights and with the reference tensors' shapes; must be zero-initialised by the caller; gradients
   * are ACCUMULATED into them.  NULL = weights are constants (fitting). */
  float* const* d_params_coarse;
  float* const* d_params_fine;
  int32_t n_params_coarse;
  int32_t n_params_fine;
  void* workspace;
  size_t workspace_bytes;
  const float* loss_scale_dev; /* optional: DEVICE pointer to {scale, 1/scale}; when set it replaces loss_scale, so the caller
                                  can derive the scale from the upstream gradients on the device without a host sync */
} mofa_b200_bwd_args;

int mofa_b200_render_rays_bwd(mofa_b200_ctx* ctx, const mofa_b200_bwd_args* args, void* stream);

/*
 * get_rays (tools/run_nerf_helpers.py:153-168) + the ray-batch packing at the top of myRenderer.render /
 * render_fitting (models/render_class.py:158-179, 393-415), SURVEY.md §8 row f3: the rays of the row-major range
 * [first_ray, first_ray + n_rays) of an H x W pinhole image, generated on the device from 21 scalars — K (HOST, 3x3
 * row-major: fx = K[0], cx = K[2], fy = K[4], cy = K[5]), c2w (HOST, [3,4] row-major), near, far — so a frame needs no
 * host->device input and every rank of a ray-sharded render produces only its own range.  rays_out [n_rays, ray_stride]
 * (device): o(3) d(3) near far viewdir(3); ray_stride >= 11, and 12 gives 16-byte rows that every kernel reads with
 * 128-bit loads.  fp32 operation order as torch evaluates get_rays: bit-identical origins and directions.
 */
int mofa_b200_generate_rays(mofa_b200_ctx* ctx, int H, int W, const float* K9, const float* c2w12, float near_,
                            float far_, int64_t first_ray, int64_t n_rays, float* rays_out, int ray_stride,
                            void* stream);

/* ---- op-level entry points (each one is also a stage of render_rays_fwd) ---- */

/* Embedder.embed (models/model.py:15-63): x [n,3] fp32 -> out [n, 3+6*multires] fp32. */
int mofa_b200_embed(mofa_b200_ctx* ctx, const float* x, int64_t n, int multires, float* out, void* stream);

/* raw2outputs (models/render_class.py:440-482).  raw [n,S,4], z [n,S], rays_d [n,3] (d_stride floats
 * apart), noise [n,S] or NULL.  Outputs may be NULL. */
int mofa_b200_raw2outputs(mofa_b200_ctx* ctx, const float* raw, const float* z, const float* rays_d,
                          int d_stride, const float* noise, int64_t n, int S, int white_bkgd,
                          float* rgb, float* disp, float* acc, float* weights, float* depth, void* stream);

/* sample_pdf(z_mid, weights[1:-1], N_importance, det) + sort(cat(z, samples)) + std
 * (models/render_class.py:324-328,345; tools/run_nerf_helpers.py:203-247).
 * z [n,S], weights [n,S]; u [n,N_i] explicit uniforms or NULL (det linspace).
 * z_samples [n,N_i] (may be NULL), z_merged [n,S+N_i], z_std [n] (may be NULL). */
int mofa_b200_sample_pdf_merge(mofa_b200_ctx* ctx, const float* z, const float* weights, const float* u,
                               int64_t n, int S, int N_i, float* z_samples, float* z_merged,
                               float* z_std, void* stream);

/* Adjoint of raw2outputs w.r.t. raw and |rays_d| (one stage of mofa_b200_render_rays_bwd).  rays [n, stride>=6]
 * (direction at +3), d_rgb [n,3] / d_acc [n] upstream gradients (may be NULL), outputs d_raw [n,S,4] and d_rays
 * [n,11] (only columns 3..5 are written; the buffer is zeroed first). */
int mofa_b200_raw2outputs_bwd(mofa_b200_ctx* ctx, const float* raw, const float* z, const float* rays, int stride,
                              const float* noise, const float* d_rgb, const float* d_acc, int64_t n, int S,
                              int white_bkgd, float* d_raw, float* d_rays, void* stream);

/* The weight-gradient contraction as the engine runs it: C[Mp, ldc] (+)= scale * A^T · B with A [P, Mp], B [P, Kb] fp16
 * row-major and the reduction over the P rows (both operands MN-major for the tensor core); columns >= n_valid are not
 * written.  Mp % 128 == 0, Kb % 64 == 0, P % 64 == 0.  use_simt selects the verification kernel. */
int mofa_b200_wgrad(mofa_b200_ctx* ctx, const void* A, int Mp, const void* B, int Kb, int n_valid, int64_t P, float scale,
                    float* C, int ldc, int use_simt, void* stream);

/* One dense layer as the engine runs it: C[M,N] = act(A0[M,K0]·B0[N,K0]^T (+ A1[M,K1]·B1[N,K1]^T) + bias).
 * fp16 operands (device, row-major, K0/K1 multiples of 64, N multiple of 128, M multiple of 128),
 * fp32 bias (may be NULL), fp16 output.  use_simt: 0 = production kernel (CTA-pair tcgen05 when N % 256 == 0),
 * 1 = SIMT verification kernel, 2 = single-CTA tcgen05 kernel. */
int mofa_b200_dense(mofa_b200_ctx* ctx, const void* A0, const void* B0, int K0, const void* A1,
                    const void* B1, int K1, const float* bias, void* C, int64_t M, int N, int relu,
                    int use_simt, void* stream);

/* Number of kernels this library has launched on this context since creation (bench accounting). */
int64_t mofa_b200_launch_count(mofa_b200_ctx* ctx);

/* Measurement aid (bench.py roofline): while enabled, every tensor-core dense launch is bracketed by a
 * pair of CUDA events recorded on the launch stream.  profile_read synchronises the device and returns,
 * per network n in {0 coarse, 1 fine}: out6[3n] = summed launch duration (ms), out6[3n+1] = summed
 * ALGORITHMIC FLOPs of those launches (2 * rows * out_features * in_features of the reference nn.Linear,
 * latent columns included), out6[3n+2] = launch count; then clears the records. */
int mofa_b200_profile_enable(mofa_b200_ctx* ctx, int on);
int mofa_b200_profile_read(mofa_b200_ctx* ctx, double* out6);

#ifdef __cplusplus
}
#endif
#endif /* MOFA_B200_H */
