// Internal declarations shared by the engine's translation units (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mofa {

// ---- dense layer (tcgen05 kernel and SIMT verification kernel) ----------------------------------
struct DenseLaunch {
  CUtensorMap tmA[2];   // activations  [M, K_seg] fp16, box {64, 128}, SWIZZLE_128B
  CUtensorMap tmB[2];   // weights      [N, K_seg] fp16, box {64, BN},  SWIZZLE_128B
  CUtensorMap tmB2[2];  // weights      [N, K_seg] fp16, box {64, 128}: one CTA's half of a pair tile
  CUtensorMap tmC;      // output       [M, N]     fp16, box {64, 128}, SWIZZLE_128B
  // raw pointers of the same operands (SIMT verification path)
  const __half* A[2];
  const __half* B[2];
  __half* C;
  int lda[2];           // row pitch (elements) of A segments
  int ldc;              // row pitch of C
  int K[2];             // K of each segment (multiple of 64; K[1] == 0 when single segment)
  const float* bias;    // [N] fp32 or nullptr
  int64_t M;            // multiple of 128
  int N;                // multiple of BN
  int BN;               // 128 or 256
  int relu;
  // fused output head (tensor-core kernels only): partial dot products of the activated tile rows with
  // head_w [head_n, N] go to head_out[row * head_stride + head_slot0 + slot * head_n + q], slot = n_tile (x groups + group)
  const float* head_w;
  float* head_out;
  int head_n, head_stride, head_slot0;
  int store_c;          // 0: do not write C (only the head consumes this layer)
  // backward-pass epilogue options: C = mask ⊙ (acc + r1_row ⊗ r1_col)
  const __half* mask;
  const float* r1_row;
  const float* r1_col;
  int r1_stride;
  int64_t M_valid;      // rows that exist in mask / head_out (0 = M)
};

cudaError_t launch_dense_tc(const DenseLaunch& L, int num_sms, cudaStream_t stream);
cudaError_t launch_dense_simt(const DenseLaunch& L, cudaStream_t stream);
cudaError_t dense_tc_configure();   // one-time cudaFuncSetAttribute for the kernel instantiations
// CTA-pair (cta_group::2) kernel, N % 256 == 0: 256x256 tile per pair of SMs (dense_tc2.cu)
cudaError_t launch_dense_tc2(const DenseLaunch& L, int num_sms, cudaStream_t stream);
cudaError_t dense_tc2_configure();
// head partial slots the pair kernel writes per 256-column tile (1, or 2 with the two-group epilogue)
int dense_tc2_head_groups();

// ---- element-wise / per-ray kernels (sampling.cu) -----------------------------------------------
cudaError_t launch_zvals_coarse(const float* rays, int stride, int64_t n, int S, int lindisp, float perturb,
                                const float* t_rand, uint64_t seed, int64_t ray_offset, float* z,
                                cudaStream_t s);
// points from rays + z  ->  fp16 PE rows X0 [P,64] (63 features + 0) and view PE rows V [P,64] (27 + 0)
// X0lo (optional): fp16(value - float(X0)) — the low image for the split-precision coarse kernel; V may be nullptr
cudaError_t launch_encode_rays(const float* rays, int stride, const float* z, int64_t n, int S,
                               int multires, int multires_views, __half* X0, __half* V, cudaStream_t s,
                               __half* X0lo = nullptr);
// explicit points (run_network): pts [P,3], viewdirs [P,3]
cudaError_t launch_encode_points(const float* pts, const float* viewdirs, int64_t P, int multires,
                                 int multires_views, __half* X0, __half* V, cudaStream_t s, __half* X0lo = nullptr);
// rays [n, stride] for the row-major ray range [first, first + n) of an H x W pinhole image (K9, c2w12: HOST arrays)
cudaError_t launch_generate_rays(int H, int W, const float* K9, const float* c2w12, float nr, float fr, int64_t first,
                                 int64_t n, float* rays, int stride, cudaStream_t s);
cudaError_t launch_embed_f32(const float* x, int64_t n, int multires, float* out, cudaStream_t s);
// out[p*4 + off + j] = A[p,:]·Wh[j,:] + b[j]    (alpha_linear / rgb_linear)
cudaError_t launch_head(const __half* A, int K, const float* Wh, const float* b, int nout, float* raw,
                        int off, int64_t P, cudaStream_t s);
cudaError_t launch_composite(const float* raw, const float* z, const float* rays_d, int d_stride,
                             const float* noise, float noise_std, uint64_t seed, int64_t ray_offset,
                             int64_t n, int S, int white_bkgd, float* rgb, float* disp, float* acc,
                             float* weights, float* depth, cudaStream_t s);
cudaError_t launch_sample_pdf_merge(const float* z, const float* weights, const float* u, int det,
                                    uint64_t seed, int64_t ray_offset, int64_t n, int S, int Ni,
                                    float* z_samples, float* z_merged, float* z_std, cudaStream_t s);
// dst[n, kpad] (fp16) = src[n, c0 : c0+k] (fp32, row pitch ld), zero padded to kpad columns
cudaError_t launch_pack_weight(const float* src, int ld, int c0, int k, int kpad, int nrows, __half* dst,
                               cudaStream_t s);
// out[n] = b[n] + sum_j Wc[n*ld + c0 + j] * lat[j]
cudaError_t launch_fold_bias(const float* Wsrc, int ld, int c0, int nlat, const float* b, const float* lat,
                             int nrows, float* out, cudaStream_t s);
cudaError_t launch_copy_f32(const float* src, float* dst, int64_t n, cudaStream_t s);
// raw[p] = (b_rgb[0..2] + sum of rgb partial slots, b_alpha + sum of alpha partial slots)
cudaError_t launch_finalize_raw(const float* hp, int stride, int a_slot0, int a_tiles, int r_slot0, int r_tiles,
                                const float* b_alpha, const float* b_rgb, float* raw, int64_t P, cudaStream_t s);

// ---- fused coarse-net kernel (coarse_fused.cu) ---------------------------------------------------
struct FusedLayerDesc {
  const float* bias;   // folded bias [n_out]
  int n_out;           // 256, or 128 for the view layer
  int kb_prim;         // 64-wide K blocks of the primary A operand (previous layer's output / X0)
  int kb_sec;          // K blocks of the secondary A operand (0 if none)
  int sec_kind;        // 1 = parked skip tensor (scratch), 2 = view encoding V
  int map_prim;        // index into the weight tensor-map array for each segment
  int map_sec;
  int store;           // write the activation for the next layer
  int save;            // also park the activation in the scratch (needed again by a later skip layer)
  int head;            // 0 none, 1 alpha_linear, 2 rgb_linear (computed in the epilogue)
};
struct FusedLaunch {
  CUtensorMap tmX0, tmV, tmScratch;
  const CUtensorMap* wmaps;        // device array
  const FusedLayerDesc* layers;    // device array
  int n_layers;
  int64_t P_rows;
  const float *w_alpha, *b_alpha, *w_rgb, *b_rgb;
  float* raw;
};
cudaError_t launch_coarse_fused(const FusedLaunch& F, int num_sms, cudaStream_t stream);
cudaError_t coarse_fused_configure();

// ---- split-precision fused coarse-net kernel (coarse_split.cu) -----------------------------------
struct SplitLayerDesc {
  const float* bias;   // folded bias [n_out] fp32; nullptr for a virtual (park) layer
  int n_out;           // 256, or 128 for the view layer
  int kb;              // 64-wide K blocks of the resident activation this layer reads: 1 (point encoding) or 4
  int map_hi, map_lo;  // weight tensor maps (fp16 hi / lo images, box = n_out/2 rows x 64 columns)
  int wait_act;        // 1: first consumer of a freshly written activation (waits on the per-block barriers)
  int store;           // write ReLU(out) as fp16 hi + lo images in place (the next layer's A operand)
  int park;            // 1: virtual layer — raw fp32 accumulators go to the per-CTA scratch (skip partial product)
  int add_park;        // 1: add the parked partial product (the skip layer)
  int add_ray;         // 1: add the per-ray view vector (view layer)
  int head;            // 0 none, 1 alpha_linear, 2 rgb_linear (computed in the epilogue)
};
struct SplitLaunch {
  CUtensorMap tmX0hi, tmX0lo;      // [P_pad, 64] fp16 point encodings (hi, lo), box {64, 128}
  const CUtensorMap* wmaps;        // device array
  const SplitLayerDesc* layers;    // device array
  int n_layers;
  int64_t P_rows;
  const float *w_alpha, *b_alpha, *w_rgb, *b_rgb;
  const float* ray_vec;            // [groups, 128] fp32 (launch_view_vec)
  int rows_per_group;              // consecutive point rows that share a view direction (samples per ray)
  float* park;                     // coarse_split_park_bytes()
  float* raw;
};
cudaError_t launch_coarse_split(const SplitLaunch& S, int num_sms, cudaStream_t stream);
cudaError_t coarse_split_configure();
size_t coarse_split_park_bytes(int num_sms);
// out[g, c] = sum_k Wv[c, k] * PE_4(dirs[g])[k], Wv [n_out, 27] fp32
cudaError_t launch_view_vec(const float* dirs, int stride, int64_t n, const float* Wv, int n_out, float* out,
                            cudaStream_t s);
// dst = fp16(w - float(fp16(w))): the low image of the split-precision weights (same layout as launch_pack_weight)
// e4m3 image of a plain [nrows, k] layer with one scale per output row: dst8 = e4m3(w / s_n), s_n = max|w_n| / 448;
// colscale[n] = s_n / act_scale (what the accumulator of an fp8 x fp8 product is multiplied by)
cudaError_t launch_pack_weight_fp8(const float* src, int ld, int c0, int k, int nrows, float act_scale, uint8_t* dst8,
                                   float* colscale, cudaStream_t s);
cudaError_t launch_pack_weight_lo(const float* src, int ld, int c0, int k, int kpad, int nrows, __half* dst,
                                  cudaStream_t s);

// ---- fine-net chain kernel (fine_chain.cu): all dense layers of a W >= 512 net in one persistent launch -------------
struct ChainLayerDesc {
  int mapA0, mapA1, mapB0, mapB1, mapC;   // indices into ChainParams::maps
  int kb0, kb1;                           // 64-wide K blocks per segment
  int n_tiles;                            // N / 256
  int a0_global, a1_global;               // A rows: 1 = global point rows (encodings), 0 = slab-local activation buffers
  int relu, store_c;
  int head_n, head_slot0;                 // fused output head (0 = none)
  int N;
  const float* bias;
  const float* head_w;
  // opt-in FP8 variant (MOFA_B200_FP8): e4m3 operands for this layer (K blocks of 128 elements), e4m3 output for the next
  int fp8_in, fp8_out;
  const float* colscale;                  // [N] weight scale / activation scale, applied to the accumulator (fp8_in)
};
struct ChainParams {
  const CUtensorMap* maps;                // device array
  const ChainLayerDesc* layers;           // device array
  int n_layers;
  int nt, nt_last;                        // n-tiles of every layer but the last / of the last layer
  int tiles_per_mb;                       // sum of n_tiles over the layers
  int slab_mb;                            // m-blocks (256 rows) per slab
  int total_mb;                           // ceil(P_rows / 256)
  float* head_out;                        // [P_rows, head_stride] head partial slots
  int head_stride;
  int64_t P_rows;
  uint32_t* counters;                     // [n_layers, total_mb], zero before the launch
  int nodep;                              // MEASUREMENT ONLY (MOFA_B200_CHAIN_NODEP=1): skip the dependency waits (wrong results)
};
cudaError_t launch_fine_chain(const ChainParams& p, int num_sms, cudaStream_t stream);
cudaError_t fine_chain_configure();
constexpr int kChainSlabMb = 56;          // m-blocks per slab: 56 x 4 n-tiles = three rounds of the 74 CTA pairs per layer
                                          // (measured 37 / 56 / 74 / 111: 162.2 / 168.6 / 166.1 / 157.6 k rays/s on one box)

// ---- backward pass (backward.cu) ----------------------------------------------------------------
cudaError_t launch_composite_bwd(const float* raw, const float* z, const float* rays, int stride, const float* noise,
                                 const float* d_rgb, const float* d_acc, float gscale, int64_t n, int S,
                                 int white_bkgd, float* d_raw, float* d_rays, cudaStream_t s, const float* sc = nullptr);
// (every scalar factor below is multiplied by *sc when sc != nullptr: a device-resident loss scale / its inverse)
cudaError_t launch_view_head_bwd(const float* d_raw, const float* w_rgb, const __half* HV, int Nh, int64_t P,
                                 __half* dZ, cudaStream_t s);
cudaError_t launch_colsum(const __half* dZ, int N, int64_t P, float* out, cudaStream_t s);
cudaError_t launch_fold_bwd(const float* fold_w, int nlat, int N, const float* d_beff, float inv_scale, float* d_lat,
                            cudaStream_t s, const float* sc = nullptr);
cudaError_t launch_pe_bwd(const float* rays, int stride, const float* z, const __half* dX0, const __half* dV, int ld,
                          int64_t n, int S, float* d_rays, cudaStream_t s);
cudaError_t launch_pack_weight_t(const float* src, int ld, int c0, int K, int krows_pad, int N, __half* dst,
                                 cudaStream_t s);
cudaError_t launch_scale_f32(float* x, float a, int64_t n, cudaStream_t s, const float* sc = nullptr);
cudaError_t launch_axpy_f32(const float* x, float a, float* y, int n, cudaStream_t s, const float* sc = nullptr);
cudaError_t launch_outer_add(const float* u, const float* v, int rows, int cols, float a, float* G, int ld, int c0,
                             cudaStream_t s, const float* sc = nullptr);
cudaError_t launch_head_wgrad(const float* g, int q0, int nq, const __half* act, int N, int64_t P, float a, float* gW,
                              float* gb, cudaStream_t s, const float* sc = nullptr);

// ---- weight-gradient GEMM (dense_wgrad.cu): C[Mp, :] += scale * A^T B, reduction over the P rows of A [P,Mp], B [P,Np]
struct WgradLaunch {
  CUtensorMap tmA;   // [P, Mp] fp16, box {64 cols, 64 rows}, SWIZZLE_128B
  CUtensorMap tmB;   // [P, cols of B] fp16, same box
  float* C;          // fp32, already offset to the first output column
  int ldc;
  int n_valid;       // real columns of B
  float scale;
  const float* scale_dev;   // optional device-resident factor multiplied into `scale`
  int Mp;            // multiple of 128
  int Np;            // multiple of BN (columns beyond the tensor are zero-filled by TMA)
  int BN;            // 128 or 256
  int64_t P;         // multiple of 64
};
cudaError_t launch_wgrad_tc(const WgradLaunch& W, int num_sms, cudaStream_t stream);
cudaError_t wgrad_configure();
cudaError_t launch_wgrad_simt(const __half* A, int lda, const __half* B, int ldb, int64_t P, int Mp, int n_valid,
                              float scale, float* C, int ldc, cudaStream_t stream);

}  // namespace mofa
