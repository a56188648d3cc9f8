// Weight-gradient GEMM (run_train.py:333-357):
//
//   dW[M', N'] += scale * sum_p  dZ[p, m'] * X[p, n']          dZ: [P, M'] fp16,  X: [P, N'] fp16,  dW fp32
//
// i.e. C = A^T · B with the REDUCTION over the row index of two row-major matrices.  Both operands are therefore
// "MN-major" for the tensor core (the M / N index is the contiguous one): TMA brings [64 p-rows x 64 columns] boxes
// (128-byte rows, 128B swizzle) and the UMMA shared-memory descriptors describe them as MN-major SWIZZLE_128B atoms
// (64 MN x 8 K per 1024-byte atom; LBO = distance between 64-wide MN blocks, SBO = distance between 8-row K groups;
// layout from CUTLASS cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>).  No transposed copies are made.
// The reduction dimension (P ~ 1e5) is split across CTAs; partial tiles are accumulated into the fp32 gradient with
// red.global.add.  Same warp roles as dense_tc.cu (TMA producer / MMA issuer / TMEM allocator / 4 epilogue warps).
#include "engine.h"
#include "ptx.cuh"

namespace mofa {

struct WgradParams {
  float* C;          // [M', ldc] fp32 (+ column offset applied by the host)
  int ldc;
  int n_valid;       // columns < n_valid are written (padded encodings have 64 stored, 63 / 27 real columns)
  float scale;
  const float* scale_dev;
  int m_tiles, n_tiles, splits;
  int kb_per_split;  // 64-row K blocks per split
  int kb_total;
};

template <int BN>
struct WgradSmem {
  static constexpr int STAGES = 4;
  static constexpr int A_BYTES = 2 * 8192;            // 128 MN x 64 K  = two [64 K x 64 MN] boxes
  static constexpr int B_BYTES = (BN / 64) * 8192;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OFF_BAR = STAGES * STAGE_BYTES;
  static constexpr int N_BARS = 2 * STAGES + 4;
  static constexpr int OFF_TPTR = OFF_BAR + N_BARS * 8;
  static constexpr int DYN_BYTES = OFF_TPTR + 16 + 1024;
};

// MN-major, 128B swizzle: start address, LBO (next 64-wide MN block), SBO (next 8-row K group), version 1, layout 2.
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16, fp16 operands, fp32 accumulate, A and B both MN-major (bits 15 and 16)
__host__ __device__ constexpr uint32_t umma_idesc_f16_f32_mn(int M, int N) {
  return (1u << 4) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

template <int BN>
__global__ void __launch_bounds__(256, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradParams p) {
  using L = WgradSmem<BN>;
  constexpr int STAGES = L::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = base + L::OFF_BAR;
  const uint32_t empty0 = full0 + 8 * STAGES;
  const uint32_t tfull0 = empty0 + 8 * STAGES;
  const uint32_t tempty0 = tfull0 + 16;
  const uint32_t tptr = base + L::OFF_TPTR;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + L::OFF_TPTR);

  const int tiles_mn = p.m_tiles * p.n_tiles;
  const int num_tiles = tiles_mn * p.splits;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int sp = t / tiles_mn, mn = t % tiles_mn;
      const int m0 = (mn / p.n_tiles) * 128, n0 = (mn % p.n_tiles) * BN;
      const int kb0 = sp * p.kb_per_split;
      const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty0 + 8 * stage, phase ^ 1u);
        if (elect_one()) {
          const uint32_t fb = full0 + 8 * stage;
          mbar_expect_tx(fb, L::STAGE_BYTES);
          const uint32_t sa = base + stage * L::STAGE_BYTES;
#pragma unroll
          for (int j = 0; j < 2; ++j) tma_load_2d(sa + j * 8192, &tmA, fb, m0 + j * 64, kb * 64);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sa + L::A_BYTES + j * 8192, &tmB, fb, n0 + j * 64, kb * 64);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_f16_f32_mn(128, BN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int sp = t / tiles_mn;
      const int kb0 = sp * p.kb_per_split;
      const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      const int as = it & 1;
      mbar_wait(tempty0 + 8 * as, ((it >> 1) & 1) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * 256;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        const uint32_t sa = base + stage * L::STAGE_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {     // K = 16 p-rows per MMA = two 8-row groups = +2048 B
            const uint64_t da = umma_desc_sw128_mnmajor(sa + k * 2048, 8192, 1024);
            const uint64_t db = umma_desc_sw128_mnmajor(sa + L::A_BYTES + k * 2048, 8192, 1024);
            umma_f16_ss(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty0 + 8 * stage);
          if (kb == kb1 - 1) umma_commit(tfull0 + 8 * as);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int mn = t % tiles_mn;
      const int m0 = (mn / p.n_tiles) * 128, n0 = (mn % p.n_tiles) * BN;
      const int as = it & 1;
      mbar_wait(tfull0 + 8 * as, (it >> 1) & 1);
      tc_fence_after();
      float* crow = p.C + static_cast<size_t>(m0 + row) * p.ldc;
      const float scale = p.scale * (p.scale_dev ? __ldg(p.scale_dev) : 1.0f);
      // 16-byte vector reductions (red.global.add.v4.f32: a quarter of the atomic instructions) wherever this row's
      // columns are 16-byte aligned — every plain W -> W layer (ldc = W) — scalar reductions otherwise (the layers whose
      // reference weight has latent columns in front: ldc = 50 + W, 256 + 2W, ...)
      const bool vec = ((reinterpret_cast<uintptr_t>(crow + n0) & 15) == 0) && ((p.ldc & 3) == 0);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + lane_base + as * 256 + c * 32, v);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if (vec && col0 + 32 <= p.n_valid) {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow + col0 + e),
                         "f"(__uint_as_float(v[e]) * scale), "f"(__uint_as_float(v[e + 1]) * scale),
                         "f"(__uint_as_float(v[e + 2]) * scale), "f"(__uint_as_float(v[e + 3]) * scale)
                         : "memory");
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int col = col0 + e;
            if (col < p.n_valid) atomicAdd(crow + col, __uint_as_float(v[e]) * scale);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * as);
    }
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

cudaError_t wgrad_configure() {
  cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       WgradSmem<256>::DYN_BYTES);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(wgrad_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              WgradSmem<128>::DYN_BYTES);
}

cudaError_t launch_wgrad_tc(const WgradLaunch& W, int num_sms, cudaStream_t stream) {
  WgradParams p;
  p.C = W.C;
  p.ldc = W.ldc;
  p.n_valid = W.n_valid;
  p.scale = W.scale;
  p.scale_dev = W.scale_dev;
  p.m_tiles = W.Mp / 128;
  p.n_tiles = W.Np / W.BN;
  p.kb_total = static_cast<int>(W.P / 64);
  const int tiles_mn = p.m_tiles * p.n_tiles;
  if (tiles_mn <= 0 || p.kb_total <= 0) return cudaSuccess;
  // Splits of the reduction: the persistent grid runs ceil(tiles / num_sms) rounds of tiles, each as long as one split,
  // so the split count is chosen to minimise  rounds x (K-blocks per split + epilogue)  — not just "enough tiles to fill
  // the machine": a 1024 x 1024 gradient has 32 output tiles, and 5 splits (160 tiles on 148 SMs) meant two rounds of 410
  // K-blocks where 9 splits give two rounds of 228.  At least 8 K-blocks (512 rows) per split; the epilogue (a 128 x BN
  // fp32 red.add per tile) is charged as 8 K-blocks.
  const int max_splits = (p.kb_total + 7) / 8;
  int splits = 1;
  long long best = -1;
  for (int s = 1; s <= max_splits && s <= num_sms; ++s) {
    const int per = (p.kb_total + s - 1) / s;
    const int eff = (p.kb_total + per - 1) / per;            // splits that actually get work
    const long long rounds = (static_cast<long long>(tiles_mn) * eff + num_sms - 1) / num_sms;
    const long long cost = rounds * (per + 8);
    if (best < 0 || cost < best) {
      best = cost;
      splits = s;
    }
  }
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  const int tiles = tiles_mn * p.splits;
  const int grid = tiles < num_sms ? tiles : num_sms;
  if (W.BN == 256)
    wgrad_tc_kernel<256><<<grid, 256, WgradSmem<256>::DYN_BYTES, stream>>>(W.tmA, W.tmB, p);
  else
    wgrad_tc_kernel<128><<<grid, 256, WgradSmem<128>::DYN_BYTES, stream>>>(W.tmA, W.tmB, p);
  return cudaGetLastError();
}

// Plain SIMT version of the same contraction (verification of the MN-major descriptors).
__global__ void wgrad_simt_kernel(const __half* __restrict__ A, int lda, const __half* __restrict__ B, int ldb, int64_t P,
                                  int Mp, int n_valid, float scale, float* __restrict__ C, int ldc) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= n_valid || m >= Mp) return;
  float acc = 0.f;
  for (int64_t p = 0; p < P; ++p) acc += __half2float(A[p * lda + m]) * __half2float(B[p * ldb + n]);
  atomicAdd(C + static_cast<size_t>(m) * ldc + n, acc * scale);
}

cudaError_t launch_wgrad_simt(const __half* A, int lda, const __half* B, int ldb, int64_t P, int Mp, int n_valid,
                              float scale, float* C, int ldc, cudaStream_t stream) {
  dim3 grid((n_valid + 127) / 128, Mp);
  wgrad_simt_kernel<<<grid, 128, 0, stream>>>(A, lda, B, ldb, P, Mp, n_valid, scale, C, ldc);
  return cudaGetLastError();
}

}  // namespace mofa
