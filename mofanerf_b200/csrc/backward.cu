// Backward pass of the ray-marching path for fitting (run_fit.py:305-313):
// gradients of the rendered maps w.r.t. ray origins/directions/view directions and the shape / texture /
// modulated-expression codes.  Network weights are constants here (they are not optimised in run_fit.py).
//
//   composite_bwd      d(rgb_map, acc_map) -> d(raw) per sample, d|rays_d|        raw2outputs   render_class.py:440-482
//   view_head_bwd      d(raw.rgb) -> dZ of the view layer (rgb_linear^T, ReLU')                 model.py:133-134
//   dense backward     dZ_T = ReLU'(T) ⊙ (Σ_consumers dZ_c · W_c,seg (+ d_alpha ⊗ w_alpha))     dense_tc*.cu, BWD epilogue
//   colsum + fold_bwd  d(folded bias) -> d(latent code)                                         the per-call fold's adjoint
//   pe_bwd             d(X0), d(V) -> d(point), d(viewdir) -> per-ray d(o), d(d)                 model.py:15-63, :315
//
// Gradients are carried in fp16 between layers multiplied by a caller-chosen power-of-two loss scale
// (removed in fp32 at the outputs) so that small gradients stay above fp16's subnormal range.
#include "engine.h"

namespace mofa {

constexpr int kBwdMaxPer = 8;   // S <= 256

__global__ void composite_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                                     const float* __restrict__ rays, int stride, const float* __restrict__ noise,
                                     const float* __restrict__ d_rgb, const float* __restrict__ d_acc, float gscale_h,
                                     const float* __restrict__ sc,
                                     int64_t n, int S, int white_bkgd, float* __restrict__ d_raw,
                                     float* __restrict__ d_rays) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const int per = (S + 31) / 32;
  const float* d = rays + r * stride + 3;
  const float nd = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  const float* zr = z + r * S;
  const float4* rr = reinterpret_cast<const float4*>(raw) + r * S;
  const float gscale = gscale_h * (sc ? __ldg(sc) : 1.0f);     // host factor x optional device-resident factor
  const float gr = d_rgb ? d_rgb[r * 3 + 0] * gscale : 0.f;
  const float gg = d_rgb ? d_rgb[r * 3 + 1] * gscale : 0.f;
  const float gb = d_rgb ? d_rgb[r * 3 + 2] * gscale : 0.f;
  float ga = d_acc ? d_acc[r] * gscale : 0.f;
  if (white_bkgd) ga -= (gr + gg + gb);                     // rgb_map += 1 - acc_map   (:479-480)

  float alpha[kBwdMaxPer], tt[kBwdMaxPer], sg[kBwdMaxPer], dist[kBwdMaxPer], gw[kBwdMaxPer];
  float cr[kBwdMaxPer], cg[kBwdMaxPer], cb[kBwdMaxPer];
  bool act[kBwdMaxPer];
  float local = 1.0f;
#pragma unroll
  for (int j = 0; j < kBwdMaxPer; ++j) {
    const int i = lane * per + j;
    alpha[j] = 0.f; tt[j] = 1.f; sg[j] = 0.f; dist[j] = 0.f; gw[j] = 0.f; cr[j] = cg[j] = cb[j] = 0.f; act[j] = false;
    if (j < per && i < S) {
      const float4 v = rr[i];
      const float dz = (i < S - 1) ? (zr[i + 1] - zr[i]) : 1e10f;
      dist[j] = dz;
      float s0 = v.w + (noise ? noise[r * S + i] : 0.f);
      act[j] = s0 > 0.f;
      sg[j] = fmaxf(s0, 0.f);
      alpha[j] = 1.0f - expf(-sg[j] * dz * nd);
      tt[j] = (1.0f - alpha[j]) + 1e-10f;
      cr[j] = 1.0f / (1.0f + expf(-v.x));
      cg[j] = 1.0f / (1.0f + expf(-v.y));
      cb[j] = 1.0f / (1.0f + expf(-v.z));
      gw[j] = gr * cr[j] + gg * cg[j] + gb * cb[j] + ga;    // dL/dw_i
      local *= tt[j];
    }
  }
  // exclusive prefix product of tt across lanes -> T at the start of this lane's run
  float incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= t;
  }
  float T0 = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) T0 = 1.0f;
  // per-sample T and w; lane-local sum of gw*w, then exclusive suffix sum across lanes
  float Ti[kBwdMaxPer], wi[kBwdMaxPer];
  float T = T0, lsum = 0.f;
#pragma unroll
  for (int j = 0; j < kBwdMaxPer; ++j) {
    Ti[j] = T;
    wi[j] = alpha[j] * T;
    T *= tt[j];
    lsum += gw[j] * wi[j];
  }
  float suf = lsum;                                         // inclusive suffix over lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_down_sync(0xffffffffu, suf, o);
    if (lane + o < 32) suf += t;
  }
  float after = suf - lsum;                                 // sum over later lanes
  float dnorm = 0.f;
  // walk the lane's run back to front: S_i = sum_{j>i} gw_j w_j
#pragma unroll
  for (int j = kBwdMaxPer - 1; j >= 0; --j) {
    const int i = lane * per + j;
    if (j < per && i < S) {
      const float dalpha = gw[j] * Ti[j] - after / tt[j];
      after += gw[j] * wi[j];
      const float one_m_a = 1.0f - alpha[j];                // = exp(-sigma*delta)
      const float delta = dist[j] * nd;
      const float dsig = act[j] ? dalpha * one_m_a * delta : 0.f;
      dnorm += dalpha * one_m_a * sg[j] * dist[j];
      float4 o4;
      o4.x = wi[j] * gr * cr[j] * (1.f - cr[j]);
      o4.y = wi[j] * gg * cg[j] * (1.f - cg[j]);
      o4.z = wi[j] * gb * cb[j] * (1.f - cb[j]);
      o4.w = dsig;
      reinterpret_cast<float4*>(d_raw)[r * S + i] = o4;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dnorm += __shfl_xor_sync(0xffffffffu, dnorm, o);
  if (lane == 0 && d_rays != nullptr) {                     // d|d| -> d(rays_d): += (not scaled down yet)
    const float inv = nd > 0.f ? 1.0f / nd : 0.f;
    d_rays[r * 11 + 3] += dnorm * d[0] * inv;
    d_rays[r * 11 + 4] += dnorm * d[1] * inv;
    d_rays[r * 11 + 5] += dnorm * d[2] * inv;
  }
}

cudaError_t launch_composite_bwd(const float* raw, const float* z, const float* rays, int stride, const float* noise,
                                 const float* d_rgb, const float* d_acc, float gscale, int64_t n, int S,
                                 int white_bkgd, float* d_raw, float* d_rays, cudaStream_t s, const float* sc) {
  if (n == 0) return cudaSuccess;
  composite_bwd_kernel<<<static_cast<unsigned>((n + 3) / 4), 128, 0, s>>>(raw, z, rays, stride, noise, d_rgb, d_acc,
                                                                         gscale, sc, n, S, white_bkgd, d_raw, d_rays);
  return cudaGetLastError();
}

// dZ_view[p, c] = (HV[p, c] > 0) ? sum_q d_raw[p, q] * W_rgb[q, c] : 0         (8 columns per thread)
__global__ void view_head_bwd_kernel(const float* __restrict__ d_raw, const float* __restrict__ w_rgb,
                                     const __half* __restrict__ HV, int Nh, int64_t P, __half* __restrict__ dZ) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int cpr = Nh / 8;
  if (idx >= P * cpr) return;
  const int64_t p = idx / cpr;
  const int c0 = static_cast<int>(idx % cpr) * 8;
  const float4 g = reinterpret_cast<const float4*>(d_raw)[p];
  const uint4 hv = reinterpret_cast<const uint4*>(HV + p * Nh + c0)[0];
  const __half* hh = reinterpret_cast<const __half*>(&hv);
  __align__(16) __half out[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float v = g.x * w_rgb[c0 + e] + g.y * w_rgb[Nh + c0 + e] + g.z * w_rgb[2 * Nh + c0 + e];
    if (!(__half2float(hh[e]) > 0.f)) v = 0.f;
    out[e] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  }
  reinterpret_cast<uint4*>(dZ + p * Nh + c0)[0] = *reinterpret_cast<const uint4*>(out);
}

cudaError_t launch_view_head_bwd(const float* d_raw, const float* w_rgb, const __half* HV, int Nh, int64_t P,
                                 __half* dZ, cudaStream_t s) {
  const int64_t tot = P * (Nh / 8);
  if (tot == 0) return cudaSuccess;
  view_head_bwd_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, s>>>(d_raw, w_rgb, HV, Nh, P, dZ);
  return cudaGetLastError();
}

// out[n] += sum_p dZ[p, n]     grid (N/64, row blocks of 2048); block 256 = 8 column groups x 32 row phases
// 16-byte row segments: lane & 7 selects 8 of the block's 64 columns, the other thread bits one of 32 row phases — a warp
// load covers four whole 128-byte lines, four loads are in flight per thread.  (2-byte-per-thread loads with one row per
// iteration ran at 4 TB/s: 12 % of a training step went into these column sums.)
__device__ __forceinline__ void add8(float (&a)[8], const uint4 v) {
  const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
  const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
  const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&v.z));
  const float2 f3 = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
  a[0] += f0.x; a[1] += f0.y; a[2] += f1.x; a[3] += f1.y; a[4] += f2.x; a[5] += f2.y; a[6] += f3.x; a[7] += f3.y;
}

__global__ void __launch_bounds__(256) colsum_kernel(const __half* __restrict__ dZ, int N, int64_t P, float* __restrict__ out) {
  __shared__ float red[8][64];
  const int cg = threadIdx.x & 7, rp = threadIdx.x >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 2048;
  const int64_t r1 = (r0 + 2048 < P) ? r0 + 2048 : P;
  const __half* base = dZ + blockIdx.x * 64 + cg * 8;
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int64_t r = r0 + rp;
  for (; r + 96 < r1; r += 128) {
    const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(base + r * N));
    const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(base + (r + 32) * N));
    const uint4 v2 = __ldg(reinterpret_cast<const uint4*>(base + (r + 64) * N));
    const uint4 v3 = __ldg(reinterpret_cast<const uint4*>(base + (r + 96) * N));
    add8(a, v0); add8(a, v1); add8(a, v2); add8(a, v3);
  }
  for (; r < r1; r += 32) add8(a, __ldg(reinterpret_cast<const uint4*>(base + r * N)));
#pragma unroll
  for (int i = 0; i < 8; ++i) {       // the four row phases of a warp hold the same columns
    a[i] += __shfl_xor_sync(0xffffffffu, a[i], 8);
    a[i] += __shfl_xor_sync(0xffffffffu, a[i], 16);
  }
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = a[i];
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(out + blockIdx.x * 64 + threadIdx.x, t);
  }
}

cudaError_t launch_colsum(const __half* dZ, int N, int64_t P, float* out, cudaStream_t s) {
  if (P == 0) return cudaSuccess;
  dim3 grid(N / 64, static_cast<unsigned>((P + 2047) / 2048));
  colsum_kernel<<<grid, 256, 0, s>>>(dZ, N, P, out);
  return cudaGetLastError();
}

// d_lat[j] += inv_scale * sum_n fold_w[n, j] * d_beff[n]
// grid (nlat / 64, 8 slices of n); 256 threads = 64 latent columns x 4 row phases.  (One thread per latent column walking
// all N rows was 81 us per launch — 4 % of a fitting iteration for a 1 MB matrix-vector product.)
__global__ void fold_bwd_kernel(const float* __restrict__ fold_w, int nlat, int N, const float* __restrict__ d_beff,
                                float inv_scale_h, float* __restrict__ d_lat, const float* __restrict__ sc) {
  __shared__ float red[4][64];
  const int c = threadIdx.x & 63, ph = threadIdx.x >> 6;
  const int j = blockIdx.x * 64 + c;
  const int per = (N + gridDim.y - 1) / gridDim.y;
  const int n0 = blockIdx.y * per, n1 = min(N, n0 + per);
  float acc = 0.f;
  if (j < nlat)
    for (int n = n0 + ph; n < n1; n += 4) acc += fold_w[static_cast<size_t>(n) * nlat + j] * d_beff[n];
  red[ph][c] = acc;
  __syncthreads();
  if (ph == 0 && j < nlat) {
    const float inv_scale = inv_scale_h * (sc ? __ldg(sc) : 1.0f);
    atomicAdd(d_lat + j, (red[0][c] + red[1][c] + red[2][c] + red[3][c]) * inv_scale);
  }
}

cudaError_t launch_fold_bwd(const float* fold_w, int nlat, int N, const float* d_beff, float inv_scale, float* d_lat,
                            cudaStream_t s, const float* sc) {
  fold_bwd_kernel<<<dim3((nlat + 63) / 64, 8), 256, 0, s>>>(fold_w, nlat, N, d_beff, inv_scale, d_lat, sc);
  return cudaGetLastError();
}

// Adjoint of the positional encoding and of pts = o + d*z; one warp per ray.
// dX0 / dV rows are `ld` halves wide (columns 0..62 / 0..26 are real).
template <int LX, int LV>
__global__ void pe_bwd_kernel(const float* __restrict__ rays, int stride, const float* __restrict__ z,
                              const __half* __restrict__ dX0, const __half* __restrict__ dV, int ld, int64_t n, int S,
                              float* __restrict__ d_rays) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const float* ray = rays + r * stride;
  float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // d_o(3), d_d(3), d_v(3)
  for (int i = lane; i < S; i += 32) {
    const int64_t p = r * S + i;
    const float zi = z[p];
    const __half* gx = dX0 + p * ld;
    const __half* gv = dV + p * ld;
    float dpt[3], dvv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = __fadd_rn(ray[c], __fmul_rn(ray[3 + c], zi));
      float g = __half2float(gx[c]);
#pragma unroll
      for (int f = 0; f < LX; ++f) {
        const float fr = static_cast<float>(1 << f);
        float sn, cs;
        sincosf(x * fr, &sn, &cs);
        g += fr * (cs * __half2float(gx[3 + 6 * f + c]) - sn * __half2float(gx[3 + 6 * f + 3 + c]));
      }
      dpt[c] = g;
      const float v = ray[8 + c];
      float h = __half2float(gv[c]);
#pragma unroll
      for (int f = 0; f < LV; ++f) {
        const float fr = static_cast<float>(1 << f);
        float sn, cs;
        sincosf(v * fr, &sn, &cs);
        h += fr * (cs * __half2float(gv[3 + 6 * f + c]) - sn * __half2float(gv[3 + 6 * f + 3 + c]));
      }
      dvv[c] = h;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      acc[c] += dpt[c];
      acc[3 + c] += dpt[c] * zi;
      acc[6 + c] += dvv[c];
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  }
  if (lane == 0) {
    float* o = d_rays + r * 11;      // accumulates in loss-scaled units (both passes); unscaled once at the end
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[c] += acc[c];
      o[3 + c] += acc[3 + c];
      o[8 + c] += acc[6 + c];
    }
  }
}

cudaError_t launch_pe_bwd(const float* rays, int stride, const float* z, const __half* dX0, const __half* dV, int ld,
                          int64_t n, int S, float* d_rays, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  pe_bwd_kernel<10, 4><<<static_cast<unsigned>((n + 3) / 4), 128, 0, s>>>(rays, stride, z, dX0, dV, ld, n, S, d_rays);
  return cudaGetLastError();
}

// dst[k, n] (fp16, [krows_pad, N]) = src[n * ld + c0 + k] for k < K, zero rows above: transposed weight segment,
// the K-major "B" operand of the backward GEMMs.
__global__ void pack_weight_t_kernel(const float* __restrict__ src, int ld, int c0, int K, int krows_pad, int N,
                                     __half* __restrict__ dst) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(krows_pad) * N) return;
  const int k = static_cast<int>(i / N), n = static_cast<int>(i % N);
  dst[i] = __float2half_rn(k < K ? src[static_cast<int64_t>(n) * ld + c0 + k] : 0.0f);
}

cudaError_t launch_pack_weight_t(const float* src, int ld, int c0, int K, int krows_pad, int N, __half* dst,
                                 cudaStream_t s) {
  const int64_t tot = static_cast<int64_t>(krows_pad) * N;
  pack_weight_t_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, s>>>(src, ld, c0, K, krows_pad, N, dst);
  return cudaGetLastError();
}

// y[i] += a * x[i]
__global__ void axpy_f32_kernel(const float* __restrict__ x, float a, float* __restrict__ y, int n,
                                const float* __restrict__ sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += a * (sc ? __ldg(sc) : 1.0f) * x[i];
}
cudaError_t launch_axpy_f32(const float* x, float a, float* y, int n, cudaStream_t s, const float* sc) {
  if (n == 0) return cudaSuccess;
  axpy_f32_kernel<<<(n + 255) / 256, 256, 0, s>>>(x, a, y, n, sc);
  return cudaGetLastError();
}

// G[r * ld + c0 + j] += a * u[r] * v[j]      (latent columns of a folded layer: d bias (x) latent)
__global__ void outer_add_kernel(const float* __restrict__ u, const float* __restrict__ v, int rows, int cols, float a,
                                 float* __restrict__ G, int ld, int c0, const float* __restrict__ sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int r = i / cols, j = i % cols;
  G[static_cast<size_t>(r) * ld + c0 + j] += a * (sc ? __ldg(sc) : 1.0f) * u[r] * v[j];
}
cudaError_t launch_outer_add(const float* u, const float* v, int rows, int cols, float a, float* G, int ld, int c0,
                             cudaStream_t s, const float* sc) {
  const int tot = rows * cols;
  if (tot == 0) return cudaSuccess;
  outer_add_kernel<<<(tot + 255) / 256, 256, 0, s>>>(u, v, rows, cols, a, G, ld, c0, sc);
  return cudaGetLastError();
}

// Head weight gradients: gW[q, c] += a * sum_p g[p*4 + q0 + q] * act[p, c];  gb[q] += a * sum_p g[p*4 + q0 + q]
// grid (N/64, row blocks of 2048); 256 threads = 64 columns x 4 row phases
__global__ void __launch_bounds__(256) head_wgrad_kernel(const float* __restrict__ g, int q0, int nq, const __half* __restrict__ act, int N,
                                  int64_t P, float a_h, float* __restrict__ gW, float* __restrict__ gb,
                                  const float* __restrict__ sc) {
  // same thread layout as colsum_kernel (16-byte activation loads, 32 row phases, two rows in flight per thread); the
  // upstream gradient row g[r, 0..3] is one 16-byte broadcast load per row
  const float a = a_h * (sc ? __ldg(sc) : 1.0f);
  __shared__ float red[8][3][64];
  __shared__ float redb[8][3];
  const int cg = threadIdx.x & 7, rp = threadIdx.x >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 2048;
  const int64_t r1 = (r0 + 2048 < P) ? r0 + 2048 : P;
  const __half* base = act + blockIdx.x * 64 + cg * 8;
  float acc[3][8], accb[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int q = 0; q < 3; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[q][i] = 0.f;
  auto row = [&](const uint4 v, const float4 g4) {
    float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    add8(x, v);
    const float gq[3] = {q0 == 3 ? g4.w : g4.x, g4.y, g4.z};    // (q0, nq) is (3, 1) for alpha_linear, (0, 3) for rgb_linear
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      if (q >= nq) break;
      const float gv = gq[q];
      accb[q] += gv;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[q][i] += gv * x[i];
    }
  };
  int64_t r = r0 + rp;
  for (; r + 32 < r1; r += 64) {
    const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(base + r * N));
    const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(base + (r + 32) * N));
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + r * 4));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(g + (r + 32) * 4));
    row(v0, g0);
    row(v1, g1);
  }
  for (; r < r1; r += 32)
    row(__ldg(reinterpret_cast<const uint4*>(base + r * N)), __ldg(reinterpret_cast<const float4*>(g + r * 4)));
#pragma unroll
  for (int q = 0; q < 3; ++q) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[q][i] += __shfl_xor_sync(0xffffffffu, acc[q][i], 8);
      acc[q][i] += __shfl_xor_sync(0xffffffffu, acc[q][i], 16);
    }
    accb[q] += __shfl_xor_sync(0xffffffffu, accb[q], 8);     // every column group of a row phase saw the same g rows
    accb[q] += __shfl_xor_sync(0xffffffffu, accb[q], 16);
  }
  if (lane < 8) {
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int i = 0; i < 8; ++i) red[warp][q][lane * 8 + i] = acc[q][i];
  }
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < 3; ++q) redb[warp][q] = accb[q];
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    for (int q = 0; q < nq; ++q) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][q][threadIdx.x];
      atomicAdd(gW + static_cast<size_t>(q) * N + blockIdx.x * 64 + threadIdx.x, a * t);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < nq) {     // bias: rows of this block, counted once (first column block only)
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += redb[w][threadIdx.x];
    atomicAdd(gb + threadIdx.x, a * t);
  }
}
cudaError_t launch_head_wgrad(const float* g, int q0, int nq, const __half* act, int N, int64_t P, float a, float* gW,
                              float* gb, cudaStream_t s, const float* sc) {
  if (P == 0) return cudaSuccess;
  dim3 grid(N / 64, static_cast<unsigned>((P + 2047) / 2048));
  head_wgrad_kernel<<<grid, 256, 0, s>>>(g, q0, nq, act, N, P, a, gW, gb, sc);
  return cudaGetLastError();
}

__global__ void scale_f32_kernel(float* __restrict__ x, float a, int64_t n, const float* __restrict__ sc) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= a * (sc ? __ldg(sc) : 1.0f);
}

cudaError_t launch_scale_f32(float* x, float a, int64_t n, cudaStream_t s, const float* sc) {
  if (n == 0) return cudaSuccess;
  scale_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(x, a, n, sc);
  return cudaGetLastError();
}

}  // namespace mofa
