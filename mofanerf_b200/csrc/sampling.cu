// Per-ray / per-point stages of render_rays that are not dense layers:
//   stratified depth sampling       models/render_class.py:291-313
//   point generation + sin/cos PE   models/render_class.py:315, models/model.py:15-63
//   sigma / rgb heads               models/model.py:130,134   (N = 1 and N = 3: SIMT dot products)
//   raw2outputs compositing         models/render_class.py:440-482  (warp prefix product)
//   sample_pdf + sort + std         tools/run_nerf_helpers.py:203-247, render_class.py:324-328,345
// plus the load-time weight repack and the per-call latent fold.
// All of it is fp32 arithmetic in the reference's operation order where that is cheap (explicit
// __fmul_rn/__fadd_rn where the compiler would otherwise contract to FMA), so that these stages agree
// with the oracle to fp32 rounding and the only reduced-precision step is the fp16 dense chain.
#include "engine.h"

namespace mofa {

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (used only when the caller passes no explicit random tensors).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float u01(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }  // [0,1)
__device__ __forceinline__ float rng_uniform(uint64_t seed, uint32_t stream, uint64_t ray, uint32_t idx) {
  uint4 c = make_uint4(static_cast<uint32_t>(ray), static_cast<uint32_t>(ray >> 32), idx, stream);
  uint2 k = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  return u01(philox4x32_10(c, k).x);
}
__device__ __forceinline__ float rng_normal(uint64_t seed, uint32_t stream, uint64_t ray, uint32_t idx) {
  uint4 c = make_uint4(static_cast<uint32_t>(ray), static_cast<uint32_t>(ray >> 32), idx, stream);
  uint2 k = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  uint4 r = philox4x32_10(c, k);
  const float a = fmaxf(u01(r.x), 5.9604645e-8f), b = u01(r.y);
  return sqrtf(-2.0f * logf(a)) * cospif(2.0f * b);
}

// torch.linspace(0, 1, steps)[i] in fp32: start + i*step below the midpoint, end - (steps-1-i)*step above.
__device__ __forceinline__ float linspace01(int i, int steps) {
  if (steps == 1) return 0.0f;
  const float step = __fdiv_rn(1.0f, static_cast<float>(steps - 1));
  // ATen evaluates the upper half as one fused multiply-subtract (single rounding); verified bit-exact
  // against torch.linspace on CPU for steps in {32, 40, 64, 128}.
  return (i < steps / 2) ? __fmul_rn(step, static_cast<float>(i))
                         : __fmaf_rn(-step, static_cast<float>(steps - 1 - i), 1.0f);
}

// ------------------------------------------------------------------------------------------------
// Ray rows: o(3) d(3) near far viewdir(3) [pad]   (models/render_class.py:176-179).  The reference's rows are 11
// floats = 44 bytes; rows padded to 12 floats (what mofa_b200_generate_rays and the renderer's pack_rays produce) are
// 16-byte aligned and are read as three 128-bit loads.
// ------------------------------------------------------------------------------------------------
struct RayRow {
  float o[3], d[3], nr, fr, v[3];
};
__device__ __forceinline__ bool rays_vec4(const float* rays, int stride) {
  return (stride & 3) == 0 && (reinterpret_cast<uintptr_t>(rays) & 15) == 0;
}
__device__ __forceinline__ RayRow load_ray(const float* __restrict__ rays, int stride, int64_t r) {
  RayRow q;
  const float* p = rays + r * stride;
  if (rays_vec4(rays, stride)) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    const float4 c = __ldg(reinterpret_cast<const float4*>(p) + 2);
    q.o[0] = a.x; q.o[1] = a.y; q.o[2] = a.z; q.d[0] = a.w; q.d[1] = b.x; q.d[2] = b.y;
    q.nr = b.z; q.fr = b.w; q.v[0] = c.x; q.v[1] = c.y; q.v[2] = c.z;
  } else {
    q.o[0] = p[0]; q.o[1] = p[1]; q.o[2] = p[2]; q.d[0] = p[3]; q.d[1] = p[4]; q.d[2] = p[5];
    q.nr = p[6]; q.fr = p[7]; q.v[0] = p[8]; q.v[1] = p[9]; q.v[2] = p[10];
  }
  return q;
}

// ------------------------------------------------------------------------------------------------
// Ray generation: get_rays (tools/run_nerf_helpers.py:153-168) + the packing of myRenderer.render
// (models/render_class.py:158-179) in one kernel — the image's rays come from 21 scalars (K, c2w, near, far) passed by
// value, so a frame needs no host->device input at all and a rank generates only its own ray range.
// Operation order as torch evaluates it in fp32: dirs = [(i - cx) / fx, -(j - cy) / fy, -1];
// rays_d[k] = (dirs0 * R[k][0] + dirs1 * R[k][1]) + dirs2 * R[k][2]; viewdir = rays_d / sqrt((d0^2 + d1^2) + d2^2).
// Ray index = row * W + col (row-major, the "bit-exact ray indices" ordering).
// ------------------------------------------------------------------------------------------------
struct RayGen {
  float cx, fx, cy, fy;
  float R[9];
  float o[3];
  float nr, fr;
  int W;
};

__global__ void generate_rays_kernel(const RayGen g, int64_t first, int64_t n, float* __restrict__ rays, int stride) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int64_t r = first + idx;
  const int row = static_cast<int>(r / g.W), col = static_cast<int>(r - static_cast<int64_t>(row) * g.W);
  const float x = __fdiv_rn(__fsub_rn(static_cast<float>(col), g.cx), g.fx);
  const float y = -__fdiv_rn(__fsub_rn(static_cast<float>(row), g.cy), g.fy);
  float d[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    // torch.sum starts from +0: (0 + a) turns a = -0 into +0, which matters for the sign of exact zeros
    d[k] = __fadd_rn(__fadd_rn(__fadd_rn(0.0f, __fmul_rn(x, g.R[3 * k])), __fmul_rn(y, g.R[3 * k + 1])), -g.R[3 * k + 2]);
  // torch.norm's reduction is acc + x * x compiled with contraction: two fused multiply-adds after the first square
  const float nrm = sqrtf(__fmaf_rn(d[2], d[2], __fmaf_rn(d[1], d[1], __fmul_rn(d[0], d[0]))));
  const float v0 = __fdiv_rn(d[0], nrm), v1 = __fdiv_rn(d[1], nrm), v2 = __fdiv_rn(d[2], nrm);
  float* p = rays + idx * stride;
  if (rays_vec4(rays, stride)) {
    reinterpret_cast<float4*>(p)[0] = make_float4(g.o[0], g.o[1], g.o[2], d[0]);
    reinterpret_cast<float4*>(p)[1] = make_float4(d[1], d[2], g.nr, g.fr);
    reinterpret_cast<float4*>(p)[2] = make_float4(v0, v1, v2, 0.0f);
  } else {
    p[0] = g.o[0]; p[1] = g.o[1]; p[2] = g.o[2]; p[3] = d[0]; p[4] = d[1]; p[5] = d[2];
    p[6] = g.nr; p[7] = g.fr; p[8] = v0; p[9] = v1; p[10] = v2;
  }
}

cudaError_t launch_generate_rays(int H, int W, const float* K9, const float* c2w12, float nr, float fr, int64_t first,
                                 int64_t n, float* rays, int stride, cudaStream_t s) {
  (void)H;
  if (n == 0) return cudaSuccess;
  RayGen g;
  g.fx = K9[0]; g.cx = K9[2]; g.fy = K9[4]; g.cy = K9[5];
  for (int k = 0; k < 3; ++k) {
    for (int j = 0; j < 3; ++j) g.R[3 * k + j] = c2w12[4 * k + j];
    g.o[k] = c2w12[4 * k + 3];
  }
  g.nr = nr; g.fr = fr; g.W = W;
  generate_rays_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(g, first, n, rays, stride);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// z_vals of the coarse pass
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float coarse_z(float nr, float fr, int i, int S, int lindisp) {
  const float t = linspace01(i, S);
  const float omt = __fsub_rn(1.0f, t);
  if (!lindisp) return __fadd_rn(__fmul_rn(nr, omt), __fmul_rn(fr, t));
  const float inv = __fadd_rn(__fmul_rn(__fdiv_rn(1.0f, nr), omt), __fmul_rn(__fdiv_rn(1.0f, fr), t));
  return __fdiv_rn(1.0f, inv);
}

__global__ void zvals_coarse_kernel(const float* __restrict__ rays, int stride, int64_t n, int S, int lindisp,
                                    float perturb, const float* __restrict__ t_rand, uint64_t seed,
                                    int64_t ray_offset, float* __restrict__ z) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n * S) return;
  const int64_t r = idx / S;
  const int i = static_cast<int>(idx - r * S);
  float nr, fr;
  if (rays_vec4(rays, stride)) {        // near / far sit in the second 16-byte word of a padded row
    const float4 b = __ldg(reinterpret_cast<const float4*>(rays + r * stride) + 1);
    nr = b.z; fr = b.w;
  } else {
    nr = rays[r * stride + 6]; fr = rays[r * stride + 7];
  }
  float zi = coarse_z(nr, fr, i, S, lindisp);
  if (perturb > 0.0f) {
    const float zl = (i > 0) ? coarse_z(nr, fr, i - 1, S, lindisp) : zi;
    const float zu = (i < S - 1) ? coarse_z(nr, fr, i + 1, S, lindisp) : zi;
    const float lower = (i > 0) ? __fmul_rn(0.5f, __fadd_rn(zi, zl)) : zi;
    const float upper = (i < S - 1) ? __fmul_rn(0.5f, __fadd_rn(zu, zi)) : zi;
    const float tr = t_rand ? t_rand[idx]
                            : rng_uniform(seed, 1u, static_cast<uint64_t>(ray_offset + r), static_cast<uint32_t>(i));
    zi = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), tr));
  }
  z[idx] = zi;
}

cudaError_t launch_zvals_coarse(const float* rays, int stride, int64_t n, int S, int lindisp, float perturb,
                                const float* t_rand, uint64_t seed, int64_t ray_offset, float* z,
                                cudaStream_t s) {
  const int64_t tot = n * S;
  if (tot == 0) return cudaSuccess;
  zvals_coarse_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, s>>>(rays, stride, n, S, lindisp,
                                                                              perturb, t_rand, seed,
                                                                              ray_offset, z);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// positional encoding
// ------------------------------------------------------------------------------------------------
// Writes one 64-half row: [x, y, z, sin(2^0 p), cos(2^0 p), ..., sin(2^(L-1) p), cos(2^(L-1) p), 0...]
// row_lo (optional): the fp16 remainder value - float(fp16(value)) of every feature (split-precision coarse kernel)
template <int L>
__device__ __forceinline__ void pe_row_f16(float x, float y, float z, __half* __restrict__ row,
                                           __half* __restrict__ row_lo = nullptr) {
  static_assert(3 + 6 * L <= 64, "PE row must fit one 64-wide K block");
  __align__(16) __half h[64];
  __align__(16) __half lo[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    h[i] = __float2half_rn(0.0f);
    lo[i] = __float2half_rn(0.0f);
  }
  auto put = [&](int i, float v) {
    h[i] = __float2half_rn(v);
    lo[i] = __float2half_rn(v - __half2float(h[i]));
  };
  put(0, x);
  put(1, y);
  put(2, z);
#pragma unroll
  for (int f = 0; f < L; ++f) {
    const float fr = static_cast<float>(1 << f);
    float s, c;
    sincosf(x * fr, &s, &c);
    put(3 + 6 * f + 0, s);
    put(3 + 6 * f + 3, c);
    sincosf(y * fr, &s, &c);
    put(3 + 6 * f + 1, s);
    put(3 + 6 * f + 4, c);
    sincosf(z * fr, &s, &c);
    put(3 + 6 * f + 2, s);
    put(3 + 6 * f + 5, c);
  }
  uint4* dst = reinterpret_cast<uint4*>(row);
  const uint4* src = reinterpret_cast<const uint4*>(h);
#pragma unroll
  for (int i = 0; i < 8; ++i) dst[i] = src[i];
  if (row_lo != nullptr) {
    uint4* dl = reinterpret_cast<uint4*>(row_lo);
    const uint4* sl = reinterpret_cast<const uint4*>(lo);
#pragma unroll
    for (int i = 0; i < 8; ++i) dl[i] = sl[i];
  }
}

template <int LX, int LV>
__global__ void encode_rays_kernel(const float* __restrict__ rays, int stride, const float* __restrict__ z,
                                   int64_t n, int S, __half* __restrict__ X0, __half* __restrict__ V,
                                   __half* __restrict__ X0lo) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n * S) return;
  const int64_t r = p / S;
  const RayRow ray = load_ray(rays, stride, r);
  const float zi = z[p];
  // pts = rays_o + rays_d * z   (render_class.py:315,329): separate multiply and add, as torch does
  const float px = __fadd_rn(ray.o[0], __fmul_rn(ray.d[0], zi));
  const float py = __fadd_rn(ray.o[1], __fmul_rn(ray.d[1], zi));
  const float pz = __fadd_rn(ray.o[2], __fmul_rn(ray.d[2], zi));
  pe_row_f16<LX>(px, py, pz, X0 + p * 64, X0lo ? X0lo + p * 64 : nullptr);
  if (V != nullptr) pe_row_f16<LV>(ray.v[0], ray.v[1], ray.v[2], V + p * 64);
}

template <int LX, int LV>
__global__ void encode_points_kernel(const float* __restrict__ pts, const float* __restrict__ vd, int64_t P,
                                     __half* __restrict__ X0, __half* __restrict__ V, __half* __restrict__ X0lo) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= P) return;
  pe_row_f16<LX>(pts[p * 3 + 0], pts[p * 3 + 1], pts[p * 3 + 2], X0 + p * 64, X0lo ? X0lo + p * 64 : nullptr);
  if (V != nullptr) pe_row_f16<LV>(vd[p * 3 + 0], vd[p * 3 + 1], vd[p * 3 + 2], V + p * 64);
}

cudaError_t launch_encode_rays(const float* rays, int stride, const float* z, int64_t n, int S, int multires,
                               int multires_views, __half* X0, __half* V, cudaStream_t s, __half* X0lo) {
  if (multires != 10 || multires_views != 4) return cudaErrorInvalidValue;
  const int64_t tot = n * S;
  if (tot == 0) return cudaSuccess;
  encode_rays_kernel<10, 4><<<static_cast<unsigned>((tot + 127) / 128), 128, 0, s>>>(rays, stride, z, n, S, X0, V, X0lo);
  return cudaGetLastError();
}

cudaError_t launch_encode_points(const float* pts, const float* viewdirs, int64_t P, int multires,
                                 int multires_views, __half* X0, __half* V, cudaStream_t s, __half* X0lo) {
  if (multires != 10 || multires_views != 4) return cudaErrorInvalidValue;
  if (P == 0) return cudaSuccess;
  encode_points_kernel<10, 4><<<static_cast<unsigned>((P + 127) / 128), 128, 0, s>>>(pts, viewdirs, P, X0, V, X0lo);
  return cudaGetLastError();
}

__global__ void embed_f32_kernel(const float* __restrict__ x, int64_t n, int L, float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int D = 3 + 6 * L;
  float* o = out + i * D;
  const float v[3] = {x[i * 3], x[i * 3 + 1], x[i * 3 + 2]};
  for (int c = 0; c < 3; ++c) o[c] = v[c];
  for (int f = 0; f < L; ++f) {
    const float fr = exp2f(static_cast<float>(f));
    for (int c = 0; c < 3; ++c) {
      float s, co;
      sincosf(v[c] * fr, &s, &co);
      o[3 + 6 * f + c] = s;
      o[3 + 6 * f + 3 + c] = co;
    }
  }
}

cudaError_t launch_embed_f32(const float* x, int64_t n, int multires, float* out, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  embed_f32_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, s>>>(x, n, multires, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// heads: raw[p, off + j] = A[p, :] · Wh[j, :] + b[j]       (one warp per point row)
// ------------------------------------------------------------------------------------------------
template <int NOUT>
__global__ void head_kernel(const __half* __restrict__ A, int K, const float* __restrict__ Wh,
                            const float* __restrict__ b, float* __restrict__ raw, int off, int64_t P) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= P) return;
  float acc[NOUT];
#pragma unroll
  for (int j = 0; j < NOUT; ++j) acc[j] = 0.0f;
  const uint4* arow = reinterpret_cast<const uint4*>(A + row * K);
  for (int c = lane; c < K / 8; c += 32) {
    const uint4 a = arow[c];
    const __half2* h2 = reinterpret_cast<const __half2*>(&a);
    float av[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(h2[e]);
      av[2 * e] = f.x;
      av[2 * e + 1] = f.y;
    }
#pragma unroll
    for (int j = 0; j < NOUT; ++j) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(Wh + static_cast<int64_t>(j) * K + c * 8));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(Wh + static_cast<int64_t>(j) * K + c * 8 + 4));
      acc[j] += av[0] * w0.x + av[1] * w0.y + av[2] * w0.z + av[3] * w0.w + av[4] * w1.x + av[5] * w1.y +
                av[6] * w1.z + av[7] * w1.w;
    }
  }
#pragma unroll
  for (int j = 0; j < NOUT; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < NOUT; ++j) raw[row * 4 + off + j] = acc[j] + b[j];
  }
}

cudaError_t launch_head(const __half* A, int K, const float* Wh, const float* b, int nout, float* raw, int off,
                        int64_t P, cudaStream_t s) {
  if (P == 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>((P + 7) / 8);
  if (nout == 1) head_kernel<1><<<grid, 256, 0, s>>>(A, K, Wh, b, raw, off, P);
  else if (nout == 3) head_kernel<3><<<grid, 256, 0, s>>>(A, K, Wh, b, raw, off, P);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// raw2outputs: one warp per ray, each lane owns a contiguous run of samples; the exclusive
// transmittance product is a local product + warp shuffle scan (front-to-back order preserved).
// ------------------------------------------------------------------------------------------------
constexpr int kMaxPer = 8;   // S <= 256

__global__ void composite_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                                 const float* __restrict__ rays_d, int d_stride,
                                 const float* __restrict__ noise, float noise_std, uint64_t seed,
                                 int64_t ray_offset, int64_t n, int S, int white_bkgd, float* __restrict__ rgb,
                                 float* __restrict__ disp, float* __restrict__ acc, float* __restrict__ weights,
                                 float* __restrict__ depth) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const int per = (S + 31) / 32;
  const float* d = rays_d + r * d_stride;
  // torch.norm(rays_d): sqrt(sum of squares)
  const float nd = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
  const float* zr = z + r * S;
  const float4* rr = reinterpret_cast<const float4*>(raw) + r * S;

  float alpha[kMaxPer], cr[kMaxPer], cg[kMaxPer], cb[kMaxPer], zz[kMaxPer];
  float local = 1.0f;
#pragma unroll
  for (int j = 0; j < kMaxPer; ++j) {
    const int i = lane * per + j;
    alpha[j] = 0.0f; cr[j] = cg[j] = cb[j] = 0.0f; zz[j] = 0.0f;
    if (j < per && i < S) {
      const float4 v = rr[i];
      const float zi = zr[i];
      float dist = (i < S - 1) ? __fsub_rn(zr[i + 1], zi) : 1e10f;            // :454-455
      dist = __fmul_rn(dist, nd);                                            // :458
      float sg = v.w;
      if (noise != nullptr) sg = __fadd_rn(sg, noise[r * S + i]);
      else if (noise_std > 0.0f)
        sg = __fadd_rn(sg, noise_std * rng_normal(seed, 2u, static_cast<uint64_t>(ray_offset + r), static_cast<uint32_t>(i)));
      sg = fmaxf(sg, 0.0f);
      alpha[j] = __fsub_rn(1.0f, expf(-__fmul_rn(sg, dist)));                // :453
      cr[j] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v.x)));                  // sigmoid :460
      cg[j] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v.y)));
      cb[j] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v.z)));
      zz[j] = zi;
      local = __fmul_rn(local, __fadd_rn(__fsub_rn(1.0f, alpha[j]), 1e-10f));   // :471
    }
  }
  // exclusive prefix product across lanes
  float incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl = __fmul_rn(incl, t);
  }
  float T = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) T = 1.0f;

  float sr = 0.f, sg_ = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxPer; ++j) {
    const int i = lane * per + j;
    if (j < per && i < S) {
      const float w = __fmul_rn(alpha[j], T);
      T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.0f, alpha[j]), 1e-10f));
      if (weights != nullptr) weights[r * S + i] = w;
      sr += w * cr[j];
      sg_ += w * cg[j];
      sb += w * cb[j];
      sd += w * zz[j];
      sa += w;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sr += __shfl_xor_sync(0xffffffffu, sr, o);
    sg_ += __shfl_xor_sync(0xffffffffu, sg_, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
    sd += __shfl_xor_sync(0xffffffffu, sd, o);
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
  }
  if (lane == 0) {
    if (white_bkgd) {                                                         // :479-480
      const float bg = __fsub_rn(1.0f, sa);
      sr += bg; sg_ += bg; sb += bg;
    }
    if (rgb != nullptr) { rgb[r * 3 + 0] = sr; rgb[r * 3 + 1] = sg_; rgb[r * 3 + 2] = sb; }
    if (acc != nullptr) acc[r] = sa;
    if (depth != nullptr) depth[r] = sd;
    if (disp != nullptr) {
      // 1 / max(1e-10, depth/acc): torch.max propagates NaN (0/0 when acc == 0)           :476
      const float q = __fdiv_rn(sd, sa);
      const float m = (q != q) ? q : fmaxf(1e-10f, q);
      disp[r] = __fdiv_rn(1.0f, m);
    }
  }
}

cudaError_t launch_composite(const float* raw, const float* z, const float* rays_d, int d_stride,
                             const float* noise, float noise_std, uint64_t seed, int64_t ray_offset, int64_t n,
                             int S, int white_bkgd, float* rgb, float* disp, float* acc, float* weights,
                             float* depth, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  if (S < 1 || S > 32 * kMaxPer) return cudaErrorInvalidValue;
  composite_kernel<<<static_cast<unsigned>((n + 3) / 4), 128, 0, s>>>(raw, z, rays_d, d_stride, noise, noise_std,
                                                                     seed, ray_offset, n, S, white_bkgd, rgb,
                                                                     disp, acc, weights, depth);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// sample_pdf + merge + std: one warp per ray, everything in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int kPdfMaxS = 256;      // coarse samples
constexpr int kPdfMaxTot = 512;    // coarse + importance samples (padded to a power of two)

__global__ void __launch_bounds__(128)
sample_pdf_merge_kernel(const float* __restrict__ z, const float* __restrict__ weights,
                        const float* __restrict__ u_in, int det, uint64_t seed, int64_t ray_offset, int64_t n,
                        int S, int Ni, float* __restrict__ z_samples, float* __restrict__ z_merged,
                        float* __restrict__ z_std) {
  __shared__ float s_cdf[4][kPdfMaxS];
  __shared__ float s_bins[4][kPdfMaxS];
  __shared__ float s_m[4][kPdfMaxTot];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 4 + wid;
  if (r >= n) return;
  float* cdf = s_cdf[wid];
  float* bins = s_bins[wid];
  float* m = s_m[wid];
  const float* zr = z + r * S;
  const float* wr = weights + r * S;
  const int nb = S - 1;   // bins (z_mid) == cdf entries
  const int nw = S - 2;   // weights[1:-1]

  for (int i = lane; i < nb; i += 32) bins[i] = __fmul_rn(0.5f, __fadd_rn(zr[i + 1], zr[i]));   // :324
  for (int i = lane; i < S; i += 32) m[i] = zr[i];
  // pdf / cdf: contiguous run per lane + warp scan
  const int per = (nw + 31) / 32;
  float loc[kMaxPer];
  float tot = 0.0f;
#pragma unroll
  for (int j = 0; j < kMaxPer; ++j) {
    const int i = lane * per + j;
    loc[j] = (j < per && i < nw) ? __fadd_rn(wr[i + 1], 1e-5f) : 0.0f;                           // :205
    tot += loc[j];
  }
  float sum = tot;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  // pdf into shared memory, then a sequential running sum by one lane: torch.cumsum (CPU) is a
  // sequential fp32 scan, and searchsorted decisions at u == cdf[k] depend on its exact rounding.
#pragma unroll
  for (int j = 0; j < kMaxPer; ++j) {
    const int i = lane * per + j;
    if (j < per && i < nw) cdf[i + 1] = __fdiv_rn(loc[j], sum);                                  // pdf
  }
  __syncwarp();
  if (lane == 0) {
    cdf[0] = 0.0f;
    float run = 0.0f;
    for (int i = 1; i <= nw; ++i) {
      run = __fadd_rn(run, cdf[i]);
      cdf[i] = run;
    }
  }
  __syncwarp();

  // inverse-CDF samples
  float ssum = 0.0f;
  for (int k = lane; k < Ni; k += 32) {
    float u;
    if (u_in != nullptr) u = u_in[r * Ni + k];
    else if (det) u = linspace01(k, Ni);
    else u = rng_uniform(seed, 3u, static_cast<uint64_t>(ray_offset + r), static_cast<uint32_t>(k));
    // searchsorted(cdf, u, right=True): number of entries <= u
    int lo = 0, hi = nb;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    const int below = max(0, lo - 1);
    const int above = min(nb - 1, lo);
    float denom = __fsub_rn(cdf[above], cdf[below]);
    if (denom < 1e-5f) denom = 1.0f;                                                             // :243
    const float t = __fdiv_rn(__fsub_rn(u, cdf[below]), denom);
    const float smp = __fadd_rn(bins[below], __fmul_rn(t, __fsub_rn(bins[above], bins[below])));
    m[S + k] = smp;
    ssum += smp;
    if (z_samples != nullptr) z_samples[r * Ni + k] = smp;
  }
  __syncwarp();
  if (z_std != nullptr) {                                                                        // :345
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
    const float mean = ssum / static_cast<float>(Ni);
    float var = 0.0f;
    for (int k = lane; k < Ni; k += 32) {
      const float dlt = m[S + k] - mean;
      var += dlt * dlt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    if (lane == 0) z_std[r] = sqrtf(var / static_cast<float>(Ni));
  }
  // sort(cat(z_vals, z_samples))  (:328): bitonic network on the padded array
  const int tot_n = S + Ni;
  int n2 = 1;
  while (n2 < tot_n) n2 <<= 1;
  for (int i = tot_n + lane; i < n2; i += 32) m[i] = __int_as_float(0x7f800000);
  __syncwarp();
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < n2; i += 32) {
        const int l = i ^ j;
        if (l > i) {
          const float a = m[i], b = m[l];
          const bool up = (i & k) == 0;
          if (up ? (a > b) : (a < b)) { m[i] = b; m[l] = a; }
        }
      }
      __syncwarp();
    }
  }
  for (int i = lane; i < tot_n; i += 32) z_merged[r * tot_n + i] = m[i];
}

cudaError_t launch_sample_pdf_merge(const float* z, const float* weights, const float* u, int det, uint64_t seed,
                                    int64_t ray_offset, int64_t n, int S, int Ni, float* z_samples,
                                    float* z_merged, float* z_std, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  if (S < 3 || S > kPdfMaxS || Ni < 1 || S + Ni > kPdfMaxTot) return cudaErrorInvalidValue;
  sample_pdf_merge_kernel<<<static_cast<unsigned>((n + 3) / 4), 128, 0, s>>>(z, weights, u, det, seed, ray_offset,
                                                                            n, S, Ni, z_samples, z_merged, z_std);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// load-time repack and per-call latent fold
// ------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ src, int ld, int c0, int k, int kpad, int nrows,
                                   __half* __restrict__ dst) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(nrows) * kpad) return;
  const int r = static_cast<int>(i / kpad), c = static_cast<int>(i % kpad);
  dst[i] = __float2half_rn(c < k ? src[static_cast<int64_t>(r) * ld + c0 + c] : 0.0f);
}

cudaError_t launch_pack_weight(const float* src, int ld, int c0, int k, int kpad, int nrows, __half* dst,
                               cudaStream_t s) {
  const int64_t tot = static_cast<int64_t>(nrows) * kpad;
  pack_weight_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, s>>>(src, ld, c0, k, kpad, nrows, dst);
  return cudaGetLastError();
}

// one warp per output row: row maximum -> scale -> e4m3 (round to nearest even, saturating)
__global__ void pack_weight_fp8_kernel(const float* __restrict__ src, int ld, int c0, int k, int nrows, float act_scale,
                                       uint8_t* __restrict__ dst8, float* __restrict__ colscale) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const float* w = src + static_cast<int64_t>(row) * ld + c0;
  float m = 0.f;
  for (int j = lane; j < k; j += 32) m = fmaxf(m, fabsf(w[j]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float sc = fmaxf(m, 1e-12f) / 448.0f;
  if (lane == 0) colscale[row] = sc / act_scale;
  for (int j = lane * 2; j < k; j += 64) {
    const float a = w[j] / sc, b = (j + 1 < k) ? w[j + 1] / sc : 0.f;
    uint16_t pk;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(pk) : "f"(b), "f"(a));     // first operand -> upper byte
    *reinterpret_cast<uint16_t*>(dst8 + static_cast<int64_t>(row) * k + j) = pk;
  }
}

cudaError_t launch_pack_weight_fp8(const float* src, int ld, int c0, int k, int nrows, float act_scale, uint8_t* dst8,
                                   float* colscale, cudaStream_t s) {
  if ((k & 1) != 0) return cudaErrorInvalidValue;
  pack_weight_fp8_kernel<<<(nrows + 7) / 8, 256, 0, s>>>(src, ld, c0, k, nrows, act_scale, dst8, colscale);
  return cudaGetLastError();
}

__global__ void pack_weight_lo_kernel(const float* __restrict__ src, int ld, int c0, int k, int kpad, int nrows,
                                      __half* __restrict__ dst) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(nrows) * kpad) return;
  const int r = static_cast<int>(i / kpad), c = static_cast<int>(i % kpad);
  const float w = c < k ? src[static_cast<int64_t>(r) * ld + c0 + c] : 0.0f;
  dst[i] = __float2half_rn(w - __half2float(__float2half_rn(w)));
}

cudaError_t launch_pack_weight_lo(const float* src, int ld, int c0, int k, int kpad, int nrows, __half* dst,
                                  cudaStream_t s) {
  const int64_t tot = static_cast<int64_t>(nrows) * kpad;
  pack_weight_lo_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, s>>>(src, ld, c0, k, kpad, nrows, dst);
  return cudaGetLastError();
}

__global__ void fold_bias_kernel(const float* __restrict__ Wsrc, int ld, int c0, int nlat,
                                 const float* __restrict__ b, const float* __restrict__ lat, int nrows,
                                 float* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  float acc = 0.0f;
  for (int j = lane; j < nlat; j += 32) acc += Wsrc[static_cast<int64_t>(row) * ld + c0 + j] * lat[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[row] = acc + b[row];
}

cudaError_t launch_fold_bias(const float* Wsrc, int ld, int c0, int nlat, const float* b, const float* lat,
                             int nrows, float* out, cudaStream_t s) {
  fold_bias_kernel<<<(nrows + 7) / 8, 256, 0, s>>>(Wsrc, ld, c0, nlat, b, lat, nrows, out);
  return cudaGetLastError();
}

__global__ void copy_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

cudaError_t launch_copy_f32(const float* src, float* dst, int64_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  copy_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(src, dst, n);
  return cudaGetLastError();
}

__global__ void finalize_raw_kernel(const float* __restrict__ hp, int stride, int a_slot0, int a_tiles, int r_slot0,
                                    int r_tiles, const float* __restrict__ b_alpha, const float* __restrict__ b_rgb,
                                    float* __restrict__ raw, int64_t P) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float* h = hp + p * stride;
  float a = 0.f, r = 0.f, g = 0.f, b = 0.f;
  for (int t = 0; t < a_tiles; ++t) a += h[a_slot0 + t];          // fixed order: deterministic
  for (int t = 0; t < r_tiles; ++t) {
    r += h[r_slot0 + 3 * t + 0];
    g += h[r_slot0 + 3 * t + 1];
    b += h[r_slot0 + 3 * t + 2];
  }
  reinterpret_cast<float4*>(raw)[p] = make_float4(r + b_rgb[0], g + b_rgb[1], b + b_rgb[2], a + b_alpha[0]);
}

// Same reduction with the slot layout known at compile time: the 128-byte slot row is read with 16-byte loads (only the
// quarters that hold live slots) instead of one scalar load per slot.
template <int STRIDE, int A0, int R0>
__global__ void finalize_raw_vec_kernel(const float* __restrict__ hp, int a_tiles, int r_tiles,
                                        const float* __restrict__ b_alpha, const float* __restrict__ b_rgb,
                                        float* __restrict__ raw, int64_t P) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float4* h4 = reinterpret_cast<const float4*>(hp + p * STRIDE);
  float row[STRIDE];
#pragma unroll
  for (int i = 0; i < STRIDE / 4; ++i) {
    const bool need = (4 * i < A0 + a_tiles && 4 * i + 4 > A0) || (4 * i < R0 + 3 * r_tiles && 4 * i + 4 > R0);
    const float4 v = need ? __ldg(h4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    row[4 * i] = v.x; row[4 * i + 1] = v.y; row[4 * i + 2] = v.z; row[4 * i + 3] = v.w;
  }
  float a = 0.f, r = 0.f, g = 0.f, b = 0.f;
#pragma unroll
  for (int t = 0; t < R0 - A0; ++t)
    if (t < a_tiles) a += row[A0 + t];                              // fixed order: deterministic
#pragma unroll
  for (int t = 0; t < (STRIDE - R0) / 3; ++t)
    if (t < r_tiles) {
      r += row[R0 + 3 * t + 0];
      g += row[R0 + 3 * t + 1];
      b += row[R0 + 3 * t + 2];
    }
  reinterpret_cast<float4*>(raw)[p] = make_float4(r + b_rgb[0], g + b_rgb[1], b + b_rgb[2], a + b_alpha[0]);
}

cudaError_t launch_finalize_raw(const float* hp, int stride, int a_slot0, int a_tiles, int r_slot0, int r_tiles,
                                const float* b_alpha, const float* b_rgb, float* raw, int64_t P, cudaStream_t s) {
  if (P == 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>((P + 255) / 256);
  const bool vec = stride == 32 && a_slot0 == 0 && r_slot0 == 8 && a_tiles <= 8 && r_tiles <= 8 &&
                   (reinterpret_cast<uintptr_t>(hp) & 15u) == 0;
  if (vec)
    finalize_raw_vec_kernel<32, 0, 8><<<grid, 256, 0, s>>>(hp, a_tiles, r_tiles, b_alpha, b_rgb, raw, P);
  else
    finalize_raw_kernel<<<grid, 256, 0, s>>>(hp, stride, a_slot0, a_tiles, r_slot0, r_tiles, b_alpha, b_rgb, raw, P);
  return cudaGetLastError();
}

}  // namespace mofa
