// One persistent kernel for the whole coarse MLP (W = 256): 23 dense layers + both heads per 128-point tile,
// activations never leave the SM.
//
// Why: per-layer launches of the W=256 net move 1 KB/point/layer through HBM for 131 kFLOP/point/layer (HBM-bound, and
// every 128x256 tile re-streams the 128 KB weight matrix from L2): measured 345-415 TFLOP/s = 9 % of a frame for 3 % of
// its FLOPs.  Here a CTA owns 128 points (2 rays x 64 samples) for all layers:
//   * activations ping-pong between two 64 KB shared-memory buffers in the UMMA K-major / 128B-swizzle layout (the
//     epilogue writes exactly the layout the next layer's tcgen05.mma reads as its A operand);
//   * weights stream from L2 through a 3-stage TMA ring of 256x64 K-blocks (3.2 MB per tile, L2-resident);
//   * accumulators double-buffer in TMEM (2 x 256 columns), so layer l+1's MMAs on K-block j start as soon as layer l's
//     epilogue has produced column block j (per-column-block mbarriers) — MMA and epilogue overlap across layers;
//   * the skip inputs (xyz_code, sigmaCodes: needed again 6 layers later) are parked in a 64 KB-per-CTA global scratch
//     (L2-resident) by TMA store and brought back by TMA into the idle ping-pong buffer just before the skip layer;
//   * alpha_linear / rgb_linear are dot products in the epilogue; raw (r,g,b,sigma) is written directly.
// Replaces models/model.py:121-137 for the coarse net; layer wiring comes from the engine's program (LayerDesc table).
#include "engine.h"
#include "ptx.cuh"

namespace mofa {

constexpr int kFusedStages = 3;
constexpr int kActBytes = 128 * 256 * 2;          // one activation buffer: 4 K-blocks of [128 rows x 64 cols] fp16
constexpr int kKbBytes = 128 * 64 * 2;            // one K-block of A
constexpr int kRingBytes = 256 * 64 * 2;          // one K-block of B (N = 256)

struct FusedSmem {
  static constexpr int OFF_ACT = 0;                                   // 2 x 64 KB
  static constexpr int OFF_RING = 2 * kActBytes;                      // 3 x 32 KB
  static constexpr int OFF_BAR = OFF_RING + kFusedStages * kRingBytes;
  // barriers: full[3], empty[3], x0_full, sec_full, mma_done[2], act_ready[2][4], tile_done, saved_ready
  static constexpr int N_BARS = 2 * kFusedStages + 2 + 2 + 8 + 2;
  static constexpr int OFF_TPTR = OFF_BAR + N_BARS * 8;
  static constexpr int OFF_HEAD = OFF_TPTR + 16;               // [128][3] fp32: head partials of epilogue group 1
  static constexpr int TOTAL = OFF_HEAD + 128 * 3 * 4;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};

// The two epilogue groups finish column blocks {0,2} first and {1,3} second; layers with four primary K-blocks consume
// them in that order (the producer streams the weight K-blocks in the same order).
__device__ __forceinline__ int kb_order(int i, int kb_prim) { return (kb_prim == 4) ? (((i & 1) << 1) | (i >> 1)) : i; }

__global__ void __launch_bounds__(384, 1)
coarse_fused_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmV,
                    const __grid_constant__ CUtensorMap tmScratch, const CUtensorMap* __restrict__ wmaps,
                    const FusedLayerDesc* __restrict__ layers, int n_layers, int num_tiles, int64_t P_rows,
                    const float* __restrict__ w_alpha, const float* __restrict__ b_alpha,
                    const float* __restrict__ w_rgb, const float* __restrict__ b_rgb, float* __restrict__ raw) {
  using L = FusedSmem;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t act0 = base + L::OFF_ACT;
  const uint32_t ring0 = base + L::OFF_RING;
  const uint32_t full0 = base + L::OFF_BAR;
  const uint32_t empty0 = full0 + 8 * kFusedStages;
  const uint32_t x0_full = empty0 + 8 * kFusedStages;
  const uint32_t sec_full = x0_full + 8;
  const uint32_t mma_done0 = sec_full + 8;          // [2] by accumulator stage
  const uint32_t act_ready0 = mma_done0 + 16;       // [2][4]
  const uint32_t tile_done = act_ready0 + 64;
  const uint32_t saved_ready = tile_done + 8;
  const uint32_t tptr = base + L::OFF_TPTR;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX0);
    prefetch_tmap(&tmV);
    prefetch_tmap(&tmScratch);
    for (int i = 0; i < kFusedStages; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
    }
    mbar_init(x0_full, 1);
    mbar_init(sec_full, 1);
    mbar_init(mma_done0, 1);
    mbar_init(mma_done0 + 8, 1);
    for (int i = 0; i < 8; ++i) mbar_init(act_ready0 + 8 * i, 4);   // one arrival per epilogue warp
    mbar_init(tile_done, 8);
    mbar_init(saved_ready, 1);
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

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    int stage = 0;
    uint32_t phase = 0;
    uint32_t done_ph[2] = {0, 0};       // phases of mma_done[acc]
    uint32_t saved_ph = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m0 = t * 128;
      // X0 K-block of this tile -> ACT[0].kb0.  ACT[0] is free once the previous tile's last MMA has retired
      // (waited for at the end of the previous iteration).
      if (elect_one()) {
        mbar_expect_tx(x0_full, kKbBytes);
        tma_load_2d(act0, &tmX0, x0_full, 0, m0);
      }
      __syncwarp();
      for (int l = 0; l < n_layers; ++l) {
        const FusedLayerDesc d = layers[l];
        // secondary A operand of THIS layer goes to ACT[(l+1)&1], which MMA(l-1) read: wait for it to retire
        if (d.kb_sec > 0) {
          if (l > 0) {
            const int a = (l - 1) & 1;
            mbar_wait(mma_done0 + 8 * a, done_ph[a]);      // peek only: the phase counter advances below
          }
          if (d.sec_kind == 1) {                            // parked skip tensor: its TMA store must have completed
            mbar_wait(saved_ready, saved_ph);
            saved_ph ^= 1u;
          }
          const uint32_t dst = act0 + ((l + 1) & 1) * kActBytes;
          if (elect_one()) {
            mbar_expect_tx(sec_full, d.kb_sec * kKbBytes);
            for (int kb = 0; kb < d.kb_sec; ++kb) {
              if (d.sec_kind == 1) tma_load_2d(dst + kb * kKbBytes, &tmScratch, sec_full, kb * 64, blockIdx.x * 128);
              else tma_load_2d(dst + kb * kKbBytes, &tmV, sec_full, kb * 64, m0);
            }
          }
          __syncwarp();
        }
        if (l > 0) done_ph[(l - 1) & 1] ^= 1u;   // one completion of mma_done[(l-1)&1] per layer, waited for or not
        // weight K-blocks of this layer: primary segment then secondary
        const int nkb = d.kb_prim + d.kb_sec;
        const uint32_t bytes = d.n_out * 128;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty0 + 8 * stage, phase ^ 1u);
          if (elect_one()) {
            const uint32_t fb = full0 + 8 * stage;
            mbar_expect_tx(fb, bytes);
            if (kb < d.kb_prim) tma_load_2d(ring0 + stage * kRingBytes, wmaps + d.map_prim, fb, kb_order(kb, d.kb_prim) * 64, 0);
            else tma_load_2d(ring0 + stage * kRingBytes, wmaps + d.map_sec, fb, (kb - d.kb_prim) * 64, 0);
          }
          __syncwarp();
          if (++stage == kFusedStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      // the next tile's X0 overwrites ACT[0].kb0, read by the last layer's MMAs
      {
        const int a = (n_layers - 1) & 1;
        mbar_wait(mma_done0 + 8 * a, done_ph[a]);
        done_ph[a] ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    uint32_t x0_ph = 0, sec_ph = 0, tile_ph = 0;
    uint32_t ready_ph[2] = {0, 0};      // phase of act_ready[b][*] (all four flip together, once per use of buffer b)
    bool first_tile = true;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      if (!first_tile) {                // accumulator stage 0 is still being drained by the previous tile's last epilogue
        mbar_wait(tile_done, tile_ph);
        tile_ph ^= 1u;
      }
      first_tile = false;
      for (int l = 0; l < n_layers; ++l) {
        const FusedLayerDesc d = layers[l];
        const int cur = l & 1;
        const uint32_t a_prim = act0 + cur * kActBytes;
        const uint32_t a_sec = act0 + (cur ^ 1) * kActBytes;
        const uint32_t d_tmem = tmem_base + (l & 1) * 256;
        const uint32_t idesc = umma_idesc_f16_f32(128, d.n_out);
        const int nkb = d.kb_prim + d.kb_sec;
        for (int kb = 0; kb < nkb; ++kb) {
          // A operand ready?
          const int kbp = kb_order(kb, d.kb_prim);     // which K-block of the primary operand this iteration consumes
          if (kb < d.kb_prim) {
            if (l == 0) {
              if (kb == 0) {
                mbar_wait(x0_full, x0_ph);
                x0_ph ^= 1u;
              }
            } else {
              mbar_wait(act_ready0 + 8 * (cur * 4 + kbp), ready_ph[cur]);
            }
          } else if (kb == d.kb_prim) {
            mbar_wait(sec_full, sec_ph);
            sec_ph ^= 1u;
          }
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a_addr = (kb < d.kb_prim) ? a_prim + kbp * kKbBytes : a_sec + (kb - d.kb_prim) * kKbBytes;
          const uint64_t da = umma_desc_sw128_kmajor(a_addr);
          const uint64_t db = umma_desc_sw128_kmajor(ring0 + stage * kRingBytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(empty0 + 8 * stage);
            if (kb == nkb - 1) umma_commit(mma_done0 + 8 * (l & 1));
          }
          __syncwarp();
          if (++stage == kFusedStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (l > 0) ready_ph[cur] ^= 1u;     // buffer `cur` was produced once (by layer l-1) and is now consumed
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: 8 warps = 2 groups x 4 lane quadrants.
    // The W=256 layers are epilogue-bound (4 K-blocks of MMA per 128x256 outputs), so two warps per SM sub-partition
    // split each tile's column blocks: group 0 takes the first half, group 1 the second.
    const int grp = (warp - 4) >> 2;
    const int ew = warp & 3;                        // TMEM lane quadrant this warp may read
    const int ep_tid = threadIdx.x - 128;
    const int row = ew * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    float* head_sh = reinterpret_cast<float*>(base_ptr + L::OFF_HEAD);
    uint32_t done_ph[2] = {0, 0};
    bool pending_save = false;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int64_t grow = static_cast<int64_t>(t) * 128 + row;
      for (int l = 0; l < n_layers; ++l) {
        const FusedLayerDesc d = layers[l];
        const int acc = l & 1;
        const uint32_t nxt = act0 + ((l + 1) & 1) * kActBytes;
        if (pending_save) {               // the parked tensor's TMA store (issued one layer ago) must land before reuse
          if (ep_tid == 0) {
            tma_store_wait_all<0>();
            mbar_arrive(saved_ready);
          }
          pending_save = false;
        }
        mbar_wait(mma_done0 + 8 * acc, done_ph[acc]);
        done_ph[acc] ^= 1u;
        tc_fence_after();
        float hacc[3] = {0.f, 0.f, 0.f};
        const float* hw = d.head == 1 ? w_alpha : w_rgb;
        const int hn = d.head == 1 ? 1 : (d.head == 2 ? 3 : 0);
        const int ncb = d.n_out / 64, half_cb = ncb / 2;
#pragma unroll 1
        for (int cb = grp * half_cb; cb < (grp + 1) * half_cb; ++cb) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            // (no software pipelining of tcgen05.ld: the destination registers of an in-flight load must not be touched
            //  by the compiler, which cannot be guaranteed across this much unrolled code — tried, produced wrong data)
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + lane_base + acc * 256 + cb * 64 + h * 32, v);
            tmem_ld_wait();
            const int ncol = cb * 64 + h * 32;
            const float4* bias4 = reinterpret_cast<const float4*>(d.bias + ncol);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b0 = __ldg(bias4 + 2 * j), b1 = __ldg(bias4 + 2 * j + 1);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j * 8 + e]) + bb[e];
              if (hn > 0) {
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                  if (q >= hn) break;
                  const float4* w4 = reinterpret_cast<const float4*>(hw + q * d.n_out + ncol) + 2 * j;
                  const float4 w0 = __ldg(w4), w1 = __ldg(w4 + 1);
                  hacc[q] += fmaxf(f[0], 0.f) * w0.x + fmaxf(f[1], 0.f) * w0.y + fmaxf(f[2], 0.f) * w0.z +
                             fmaxf(f[3], 0.f) * w0.w + fmaxf(f[4], 0.f) * w1.x + fmaxf(f[5], 0.f) * w1.y +
                             fmaxf(f[6], 0.f) * w1.z + fmaxf(f[7], 0.f) * w1.w;
                }
              }
              if (d.store) {
                // ReLU + fp16 conversion in one instruction per pair (cvt.rn.relu.f16x2.f32), then a packed
                // min against the fp16 maximum so that an overflow saturates instead of becoming +inf
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(pk[e]) : "f"(f[2 * e + 1]), "f"(f[2 * e]));
                  asm("min.f16x2 %0, %0, %1;" : "+r"(pk[e]) : "r"(0x7bff7bffu));
                }
                const int chunk = h * 4 + j;
                const uint32_t addr = nxt + cb * kKbBytes + row * 128 + ((chunk ^ (row & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                             "r"(pk[3])
                             : "memory");
              }
            }
          }
          if (d.store) {                  // column block cb of the next layer's A operand is complete for this warp's rows
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(act_ready0 + 8 * (((l + 1) & 1) * 4 + cb));
          }
        }
        if (hn > 0) {                     // combine the two groups' partial dot products (rows are shared, columns split)
          if (grp == 1) {
#pragma unroll
            for (int q = 0; q < 3; ++q) head_sh[row * 3 + q] = hacc[q];
          }
          named_bar_sync(2, 256);
          if (grp == 0 && grow < P_rows) {
            if (d.head == 1) raw[grow * 4 + 3] = hacc[0] + head_sh[row * 3] + b_alpha[0];
            else {
              raw[grow * 4 + 0] = hacc[0] + head_sh[row * 3 + 0] + b_rgb[0];
              raw[grow * 4 + 1] = hacc[1] + head_sh[row * 3 + 1] + b_rgb[1];
              raw[grow * 4 + 2] = hacc[2] + head_sh[row * 3 + 2] + b_rgb[2];
            }
          }
          named_bar_sync(2, 256);         // head_sh may be rewritten by the next head layer
        }
        if (d.save) {                     // park this layer's output (all 128 rows x 256) for the skip layer
          named_bar_sync(1, 256);         // every epilogue warp has written + fenced its rows / columns
          if (ep_tid == 0) {
            for (int kb = 0; kb < 4; ++kb) tma_store_2d(&tmScratch, nxt + kb * kKbBytes, kb * 64, blockIdx.x * 128);
            tma_store_commit();
          }
          pending_save = true;
        }
        tc_fence_before();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(tile_done);
    }
    if (ep_tid == 0) tma_store_wait_all<0>();
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

cudaError_t coarse_fused_configure() {
  return cudaFuncSetAttribute(coarse_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedSmem::DYN_BYTES);
}

cudaError_t launch_coarse_fused(const FusedLaunch& F, int num_sms, cudaStream_t stream) {
  const int num_tiles = static_cast<int>((F.P_rows + 127) / 128);
  if (num_tiles <= 0) return cudaSuccess;
  const int grid = num_tiles < num_sms ? num_tiles : num_sms;
  coarse_fused_kernel<<<grid, 384, FusedSmem::DYN_BYTES, stream>>>(F.tmX0, F.tmV, F.tmScratch, F.wmaps, F.layers,
                                                                  F.n_layers, num_tiles, F.P_rows, F.w_alpha, F.b_alpha,
                                                                  F.w_rgb, F.b_rgb, F.raw);
  return cudaGetLastError();
}

}  // namespace mofa
