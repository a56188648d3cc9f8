// Dense layer, CTA-pair version: tcgen05.mma.cta_group::2 on a 256 x 256 output tile per pair of SMs.
//
// Same contract as dense_tc.cu (C = act(sum_seg A_seg·B_seg^T + bias), fp16 in / fp32 accumulate / fp16 out)
// for N % 256 == 0.  Why pairs: with one CTA per 128x256 tile every SM streams the whole 256-row weight
// block from L2 for its own 128 rows (ncu on the 1-CTA kernel: 6.4 GB L2->SM per 0.55 TFLOP launch,
// tensor pipe 72-75 % active).  In cta_group::2 mode the two SMs of a TPC hold 128 A-rows each and one
// HALF of the B tile each; the tensor cores read both halves, so L2->SM bytes per FLOP drop by a third and
// a stage shrinks to 32 KB per SM (6 stages instead of 4 in the same shared memory).
//
// Per CTA (128 + 32*EW threads; EW = 4 or 8 epilogue warps), cluster of 2:
//   warp 0 lane 0 : TMA producer — own A half (128x64) + own B half (128x64) per stage, completion bytes
//                   routed to the LEADER CTA's full barrier (cp.async.bulk.tensor ... .cta_group::2)
//   warp 1 lane 0 : MMA issuer (leader CTA only) — UMMA 256x256x16, commits multicast to both CTAs
//   warp 2        : TMEM allocator (cta_group::2, both CTAs)
//   warps 4..     : epilogue of this CTA's 128 rows (TMEM -> +bias, ReLU -> fp16 -> swizzled smem -> TMA store);
//                   one lane per warp releases the accumulator stage on the leader's barrier (remote arrive).
//                   A warp reads the TMEM lane quadrant warp % 4; with EW = 8 warps 4..7 take columns 0..127 of the
//                   tile and warps 8..11 columns 128..255 (two independent groups: own staging buffer, named barrier
//                   and head slot), and the register file is re-split with setmaxnreg (40 for warps 0..3, 232 for the
//                   epilogue) — the epilogue of a 128x256 tile was busy ~90 % of a K=1024 tile's MMA time with 4 warps
//                   (ncu source page), so every layer with extra epilogue work (heads, K=64) was epilogue-bound.
#include <cstdlib>

#include "dense_epilogue.cuh"
#include "engine.h"
#include "pair.cuh"
#include "ptx.cuh"

namespace mofa {

struct Dense2Params {
  EpiParams epi;
  int m_tiles;   // 256-row pair tiles
  int n_tiles;   // 256-column tiles
  int kb0, kb1;
  int prefetch;   // L2-prefetch the next m-block's activation rows
};

template <int STAGES, int CBUFS>
struct Dense2Smem {
  static constexpr int A_BYTES = 128 * 64 * 2;
  static constexpr int B_BYTES = 128 * 64 * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;          // per CTA
  static constexpr int C_BYTES = 128 * 64 * 2;
  static constexpr int OFF_C = STAGES * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_C + CBUFS * C_BYTES;
  static constexpr int N_BARS = 2 * STAGES + 4;
  static constexpr int OFF_TPTR = OFF_BAR + N_BARS * 8;
  static constexpr int TOTAL = OFF_TPTR + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};

// EW: epilogue warps (4 or 8).  NBUF: staging buffers per epilogue group (EW/4 groups).
template <int STAGES, bool BWD, int EW, int NBUF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 32 * EW, 1)
dense_tc2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                 const __grid_constant__ CUtensorMap tmC, const Dense2Params p) {
  using L = Dense2Smem<STAGES, (EW / 4) * NBUF>;
  constexpr int BN = 256;
  constexpr int G = EW / 4;   // epilogue groups
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t full0 = base + L::OFF_BAR;            // used in the leader only (bytes of both CTAs)
  const uint32_t empty0 = full0 + 8 * STAGES;          // per CTA
  const uint32_t tfull0 = empty0 + 8 * STAGES;         // per CTA
  const uint32_t tempty0 = tfull0 + 16;                // used in the leader only (EW warp arrivals per CTA)
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
      mbar_init(tempty0 + 8 * i, 2 * EW);
    }
    fence_mbar_init();
  }
  __syncthreads();
  cluster_sync_all();                                  // barriers of both CTAs initialised before any remote use
  if (warp == 2) {
    tmem_alloc_cg2(tptr, 512);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + L::OFF_TPTR);

  const int num_tiles = p.m_tiles * p.n_tiles;
  const int total_kb = p.kb0 + p.kb1;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp < 4) {
  // warpgroup 0 gives registers back, the epilogue warpgroups take them: 128*40 + 256*232 = 64512 = 384*168
  if constexpr (EW == 8) setmaxnreg_dec_40();
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs; whole warp, one elected lane issues)
    int stage = 0;
    uint32_t phase = 0;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      const int m0 = (t / p.n_tiles) * 256 + static_cast<int>(rank) * 128;
      const int n0 = (t % p.n_tiles) * BN + static_cast<int>(rank) * 128;
      // Activation rows come from HBM on first touch.  The four n-tiles of an m-block run on four neighbouring pairs
      // at about the same time; the pair whose NEXT tile is the n = 0 tile of an m-block starts that m-block's rows
      // towards L2 one whole tile (~6 us) ahead, so that the 6-stage smem ring only has to cover L2 latency.
      {
        const int tn = t + num_pairs;
        if (p.prefetch && tn < num_tiles && (tn % p.n_tiles) == 0 && elect_one()) {
          const int pm0 = (tn / p.n_tiles) * 256 + static_cast<int>(rank) * 128;
          for (int kb = 0; kb < p.kb0; ++kb) tma_prefetch_l2_2d(&tmA0, kb * 64, pm0);
          for (int kb = 0; kb < p.kb1; ++kb) tma_prefetch_l2_2d(&tmA1, kb * 64, pm0);
        }
      }
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(empty0 + 8 * stage, phase ^ 1u);
        const uint32_t fb_local = full0 + 8 * stage;
        const uint32_t fb = mapa_u32(fb_local, 0);
        const uint32_t sa = base + stage * L::STAGE_BYTES;
        const uint32_t sb = sa + L::A_BYTES;
        if (elect_one()) {
          if (leader) mbar_expect_tx(fb_local, 2 * L::STAGE_BYTES);    // bytes of both CTAs land on this barrier
          if (kb < p.kb0) {
            tma_load_2d_cg2(sa, &tmA0, fb, kb * 64, m0);
            tma_load_2d_cg2(sb, &tmB0, fb, kb * 64, n0);
          } else {
            const int k = (kb - p.kb0) * 64;
            tma_load_2d_cg2(sa, &tmA1, fb, k, m0);
            tma_load_2d_cg2(sb, &tmB1, fb, k, n0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA; whole warp, one elected lane issues)
    constexpr uint32_t idesc = umma_idesc_f16_f32(256, BN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(tempty0 + 8 * as, aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        const uint32_t sa = base + stage * L::STAGE_BYTES;
        const uint64_t da = umma_desc_sw128_kmajor(sa);
        const uint64_t db = umma_desc_sw128_kmajor(sa + L::A_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16_ss_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_cg2_mc(empty0 + 8 * stage, 0x3);   // frees this stage in BOTH CTAs
          if (kb == total_kb - 1) umma_commit_cg2_mc(tfull0 + 8 * as, 0x3);   // accumulator halves ready in both CTAs
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  }
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
    if constexpr (EW == 8) setmaxnreg_inc_232();
    const int ew = warp & 3;                            // TMEM lane quadrant
    const int grp = (warp - 4) >> 2;                    // column group
    const int row = ew * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    EpiGroup g;
    g.cbuf0 = base + L::OFF_C + grp * NBUF * L::C_BYTES;
    g.cb0 = grp * (BN / 64 / G);
    g.cb1 = g.cb0 + BN / 64 / G;
    g.bar_id = 1 + grp;
    g.gtid = static_cast<int>(threadIdx.x) - 128 - grp * 128;
    int it = 0;
    uint32_t cnt = 0;
    for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
      const int m0 = (t / p.n_tiles) * 256 + static_cast<int>(rank) * 128;
      const int n0 = (t % p.n_tiles) * BN;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      g.slot = (t % p.n_tiles) * G + grp;
      mbar_wait(tfull0 + 8 * as, aphase);
      tc_fence_after();
      epilogue_tile<BN, BWD, NBUF>(p.epi, &tmC, tmem_base + lane_base + as * BN, g, cnt, m0, n0, row);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(tempty0 + 8 * as, 0));   // EW warps x 2 CTAs release the stage
    }
    if (g.gtid == 0) tma_store_wait_all<0>();
  }
  __syncthreads();
  cluster_sync_all();                                  // peer smem / TMEM stay valid until both CTAs are done
  if (warp == 2) tmem_dealloc_cg2(tmem_base, 512);
}

EpiParams make_epi(const DenseLaunch& L);   // dense_tc.cu

// Variants (MOFA_B200_PAIR_EPI): 0 = 4 epilogue warps, 6 stages, 2 staging buffers;
//                                1 = 8 epilogue warps (2 groups), 6 stages, 1 staging buffer per group  [default];
//                                2 = 8 epilogue warps (2 groups), 5 stages, 2 staging buffers per group.
// Same box, 800x800 FULL frame, fine-net TFLOP/s / rays/s: variant 0 1268 / 166.6 k, variant 1 1301 / 170.0 k,
// variant 2 1282 / 167.9 k (profiles/r01_pair_epilogue_variants.json): the sixth ring stage is worth more than a
// second staging buffer once two groups overlap each other's TMA-store wait.
static int pair_variant() {
  static const int v = [] {
    const char* e = getenv("MOFA_B200_PAIR_EPI");
    return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
  }();
  return v;
}

int dense_tc2_head_groups() { return pair_variant() == 0 ? 1 : 2; }

template <int STAGES, bool BWD, int EW, int NBUF>
static cudaError_t configure_one() {
  return cudaFuncSetAttribute(dense_tc2_kernel<STAGES, BWD, EW, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              Dense2Smem<STAGES, (EW / 4) * NBUF>::DYN_BYTES);
}

cudaError_t dense_tc2_configure() {
  cudaError_t e;
  if ((e = configure_one<6, false, 4, 2>()) != cudaSuccess) return e;
  if ((e = configure_one<6, true, 4, 2>()) != cudaSuccess) return e;
  if ((e = configure_one<6, false, 8, 1>()) != cudaSuccess) return e;
  if ((e = configure_one<6, true, 8, 1>()) != cudaSuccess) return e;
  if ((e = configure_one<5, false, 8, 2>()) != cudaSuccess) return e;
  return configure_one<5, true, 8, 2>();
}

template <int STAGES, int EW, int NBUF>
static void launch_variant(const DenseLaunch& L, const Dense2Params& p, int pairs, int s1, cudaStream_t stream) {
  constexpr int kSmem = Dense2Smem<STAGES, (EW / 4) * NBUF>::DYN_BYTES;
  constexpr int kThreads = 128 + 32 * EW;
  if (L.mask != nullptr || L.r1_row != nullptr)
    dense_tc2_kernel<STAGES, true, EW, NBUF><<<2 * pairs, kThreads, kSmem, stream>>>(L.tmA[0], L.tmA[s1], L.tmB2[0],
                                                                                  L.tmB2[s1], L.tmC, p);
  else
    dense_tc2_kernel<STAGES, false, EW, NBUF><<<2 * pairs, kThreads, kSmem, stream>>>(L.tmA[0], L.tmA[s1], L.tmB2[0],
                                                                                   L.tmB2[s1], L.tmC, p);
}

cudaError_t launch_dense_tc2(const DenseLaunch& L, int num_sms, cudaStream_t stream) {
  Dense2Params p;
  p.epi = make_epi(L);
  p.m_tiles = static_cast<int>((L.M + 255) / 256);
  p.n_tiles = L.N / 256;
  p.kb0 = L.K[0] / 64;
  p.kb1 = L.K[1] / 64;
  {
    // measured: -4 % on the fine-net layers (the extra L2 requests cost more than the latency they hide) => opt-in only
    static const int pf = [] { const char* v = getenv("MOFA_B200_L2_PREFETCH"); return (v && v[0] == '1') ? 1 : 0; }();
    p.prefetch = pf;
  }
  const long long tiles = static_cast<long long>(p.m_tiles) * p.n_tiles;
  if (tiles <= 0) return cudaSuccess;
  const int max_pairs = num_sms / 2;
  const int pairs = static_cast<int>(tiles < max_pairs ? tiles : max_pairs);
  const int s1 = p.kb1 > 0 ? 1 : 0;
  switch (pair_variant()) {
    case 1: launch_variant<6, 8, 1>(L, p, pairs, s1, stream); break;
    case 2: launch_variant<5, 8, 2>(L, p, pairs, s1, stream); break;
    default: launch_variant<6, 4, 2>(L, p, pairs, s1, stream); break;
  }
  return cudaGetLastError();
}

}  // namespace mofa
