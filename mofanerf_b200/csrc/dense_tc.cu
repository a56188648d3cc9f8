// Dense layer of the MoFaNeRF MLPs on the sm_100a tensor cores.
//
//   C[M,N] = act( A0[M,K0]·B0[N,K0]^T  (+ A1[M,K1]·B1[N,K1]^T)  + bias[N] )      fp16 in, fp32 accumulate, fp16 out
//
// replaces the reference's nn.Linear + ReLU (+ torch.cat of the skip input) calls in
// models/model.py:126-135, 226-230.  The skip / concat layers never materialise the concatenation:
// each concat segment is its own K range accumulated into the same TMEM tile, and the latent-code
// columns are folded into `bias` once per render call (see engine.cu).
//
// Structure (one persistent CTA per SM, 256 threads, warp-specialised):
//   warp 0 lane 0 : TMA producer  — cp.async.bulk.tensor 2-D loads of A (128x64) and B (BNx64) tiles,
//                   128-byte swizzle, 4-stage mbarrier ring
//   warp 1 lane 0 : MMA issuer    — tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16 x4 per stage,
//                   accumulators in TMEM (2 stages x BN columns), tcgen05.commit -> mbarriers
//   warp 2        : TMEM allocator
//   warps 4..7    : epilogue      — tcgen05.ld (32 lanes x 32 columns), +bias, ReLU, fp16 pack,
//                   swizzled st.shared, TMA store; overlaps the next tile's MMAs
#include "dense_epilogue.cuh"
#include "engine.h"
#include "ptx.cuh"

namespace mofa {

struct DenseParams {
  EpiParams epi;
  int m_tiles;
  int n_tiles;
  int kb0;   // 64-wide K blocks in segment 0
  int kb1;   // ... in segment 1 (0 if none)
};

template <int BN, int STAGES>
struct DenseSmem {
  static constexpr int A_BYTES = 128 * 64 * 2;
  static constexpr int B_BYTES = BN * 64 * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int C_BYTES = 128 * 64 * 2;
  static constexpr int OFF_C = STAGES * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_C + 2 * C_BYTES;
  static constexpr int N_BARS = 2 * STAGES + 4;
  static constexpr int OFF_TPTR = OFF_BAR + N_BARS * 8;
  static constexpr int TOTAL = OFF_TPTR + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;   // slack for manual 1024-byte alignment
};

template <int BN, int STAGES, bool BWD>
__global__ void __launch_bounds__(256, 1)
dense_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                const __grid_constant__ CUtensorMap tmC, const DenseParams p) {
  using L = DenseSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t full0 = base + L::OFF_BAR;
  const uint32_t empty0 = full0 + 8 * STAGES;
  const uint32_t tfull0 = empty0 + 8 * STAGES;
  const uint32_t tempty0 = tfull0 + 16;
  const uint32_t tptr = base + L::OFF_TPTR;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmB0);
    prefetch_tmap(&tmC);
    if (p.kb1 > 0) {
      prefetch_tmap(&tmA1);
      prefetch_tmap(&tmB1);
    }
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, 128);
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

  const int num_tiles = p.m_tiles * p.n_tiles;
  const int total_kb = p.kb0 + p.kb1;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp, one elected lane issues)
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m0 = (t / p.n_tiles) * 128;
      const int n0 = (t % p.n_tiles) * BN;
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(empty0 + 8 * stage, phase ^ 1u);
        const uint32_t fb = full0 + 8 * stage;
        const uint32_t sa = base + stage * L::STAGE_BYTES;
        const uint32_t sb = sa + L::A_BYTES;
        if (elect_one()) {
          mbar_expect_tx(fb, L::STAGE_BYTES);
          if (kb < p.kb0) {
            tma_load_2d(sa, &tmA0, fb, kb * 64, m0);
            tma_load_2d(sb, &tmB0, fb, kb * 64, n0);
          } else {
            const int k = (kb - p.kb0) * 64;
            tma_load_2d(sa, &tmA1, fb, k, m0);
            tma_load_2d(sb, &tmB1, fb, k, n0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
    constexpr uint32_t idesc = umma_idesc_f16_f32(128, BN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(tempty0 + 8 * as, aphase ^ 1u);     // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(full0 + 8 * stage, phase);        // TMA bytes have landed
        tc_fence_after();
        const uint32_t sa = base + stage * L::STAGE_BYTES;
        const uint64_t da = umma_desc_sw128_kmajor(sa);
        const uint64_t db = umma_desc_sw128_kmajor(sa + L::A_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {             // 4 x (K=16): +32 B along the swizzled row
            umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty0 + 8 * stage);          // frees the smem slot when these MMAs retire
          if (kb == total_kb - 1) umma_commit(tfull0 + 8 * as);   // accumulator ready for the epilogue
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 4;                        // == warp % 4: the TMEM lane quadrant this warp may read
    const int ep_tid = threadIdx.x - 128;
    const int row = ew * 32 + lane;                 // tile row == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    int it = 0;
    uint32_t cnt = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int m0 = (t / p.n_tiles) * 128;
      const int n0 = (t % p.n_tiles) * BN;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(tfull0 + 8 * as, aphase);
      tc_fence_after();
      const EpiGroup g{base + L::OFF_C, 0, BN / 64, 1, t % p.n_tiles, ep_tid};
      epilogue_tile<BN, BWD, 2>(p.epi, &tmC, tmem_base + lane_base + as * BN, g, cnt, m0, n0, row);
      tc_fence_before();
      mbar_arrive(tempty0 + 8 * as);                // 128 arrivals release the accumulator stage
    }
    if (ep_tid == 0) tma_store_wait_all<0>();
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

EpiParams make_epi(const DenseLaunch& L) {
  EpiParams e;
  e.bias = L.bias;
  e.head_w = L.head_w;
  e.head_out = L.head_out;
  e.relu = L.relu;
  e.store_c = L.store_c;
  e.head_n = L.head_w ? L.head_n : 0;
  e.head_stride = L.head_stride;
  e.head_slot0 = L.head_slot0;
  e.N = L.N;
  e.M = static_cast<int>(L.M_valid > 0 ? L.M_valid : L.M);
  e.mask = L.mask;
  e.r1_row = L.r1_row;
  e.r1_col = L.r1_col;
  e.r1_stride = L.r1_stride;
  return e;
}

static constexpr int kStages256 = 4;
static constexpr int kStages128 = 4;

cudaError_t dense_tc_configure() {
  cudaError_t e = cudaSuccess;
  auto set = [&](const void* fn, int bytes) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  };
  set(reinterpret_cast<const void*>(dense_tc_kernel<256, kStages256, false>), DenseSmem<256, kStages256>::DYN_BYTES);
  set(reinterpret_cast<const void*>(dense_tc_kernel<256, kStages256, true>), DenseSmem<256, kStages256>::DYN_BYTES);
  set(reinterpret_cast<const void*>(dense_tc_kernel<128, kStages128, false>), DenseSmem<128, kStages128>::DYN_BYTES);
  set(reinterpret_cast<const void*>(dense_tc_kernel<128, kStages128, true>), DenseSmem<128, kStages128>::DYN_BYTES);
  return e;
}

cudaError_t launch_dense_tc(const DenseLaunch& L, int num_sms, cudaStream_t stream) {
  DenseParams p;
  p.epi = make_epi(L);
  p.m_tiles = static_cast<int>(L.M / 128);
  p.n_tiles = L.N / L.BN;
  p.kb0 = L.K[0] / 64;
  p.kb1 = L.K[1] / 64;
  const long long tiles = static_cast<long long>(p.m_tiles) * p.n_tiles;
  if (tiles <= 0) return cudaSuccess;
  const int grid = static_cast<int>(tiles < num_sms ? tiles : num_sms);
  const int s1 = p.kb1 > 0 ? 1 : 0;
  const bool bwd = L.mask != nullptr || L.r1_row != nullptr;
#define MOFA_LAUNCH(BN_, ST_, BW_)                                                                        \
  dense_tc_kernel<BN_, ST_, BW_><<<grid, 256, DenseSmem<BN_, ST_>::DYN_BYTES, stream>>>(L.tmA[0], L.tmA[s1],  \
                                                                                        L.tmB[0], L.tmB[s1], L.tmC, p)
  if (L.BN == 256) {
    if (bwd) MOFA_LAUNCH(256, kStages256, true); else MOFA_LAUNCH(256, kStages256, false);
  } else {
    if (bwd) MOFA_LAUNCH(128, kStages128, true); else MOFA_LAUNCH(128, kStages128, false);
  }
#undef MOFA_LAUNCH
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Plain SIMT kernel with the same contract (fp16 operands, fp32 accumulate, fp16 output).
// Verification aid only (MOFA_FLAG_GEMM_SIMT): lets the tests separate "tensor-core kernel wrong"
// from "pipeline wrong".  64x64 tile, 16x16 threads, 4x4 outputs per thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dense_simt_kernel(const __half* __restrict__ A0, int lda0, const __half* __restrict__ B0, int K0,
                  const __half* __restrict__ A1, int lda1, const __half* __restrict__ B1, int K1,
                  const float* __restrict__ bias, __half* __restrict__ C, int ldc, int relu) {
  __shared__ float sA[16][64 + 1];
  __shared__ float sB[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * 64;
  const int n0 = blockIdx.y * 64;
  float acc[4][4] = {};
  for (int seg = 0; seg < 2; ++seg) {
    const __half* A = seg ? A1 : A0;
    const __half* B = seg ? B1 : B0;
    const int K = seg ? K1 : K0;
    const int lda = seg ? lda1 : lda0;
    for (int k0 = 0; k0 < K; k0 += 16) {
      for (int i = threadIdx.x; i < 64 * 16; i += 256) {
        const int r = i >> 4, c = i & 15;
        sA[c][r] = __half2float(A[(m0 + r) * lda + k0 + c]);
        sB[c][r] = __half2float(B[static_cast<int64_t>(n0 + r) * K + k0 + c]);
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a[i] = sA[k][ty * 4 + i];
          b[i] = sB[k][tx * 4 + i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      float x = acc[i][j] + (bias ? bias[n] : 0.0f);
      if (relu) x = fmaxf(x, 0.0f);
      x = fminf(fmaxf(x, -65504.0f), 65504.0f);
      C[(m0 + ty * 4 + i) * ldc + n] = __float2half_rn(x);
    }
}

cudaError_t launch_dense_simt(const DenseLaunch& L, cudaStream_t stream) {
  dim3 grid(static_cast<unsigned>(L.M / 64), L.N / 64);
  dense_simt_kernel<<<grid, 256, 0, stream>>>(L.A[0], L.lda[0], L.B[0], L.K[0], L.A[1], L.lda[1], L.B[1],
                                              L.K[1], L.bias, L.C, L.ldc, L.relu);
  return cudaGetLastError();
}

}  // namespace mofa
