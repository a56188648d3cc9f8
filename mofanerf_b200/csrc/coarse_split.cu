// Split-precision fused coarse network (W = 256): one persistent CTA-pair kernel for the whole coarse MLP with
// fp32-class arithmetic on the tensor cores.
//
// Why: the coarse pass decides WHERE the fine pass samples (sample_pdf, tools/run_nerf_helpers.py:203-247).  With
// single fp16 operands its sigma carries ~1e-3 relative error, the resampled depths move by ~1e-3 and a random-init
// field with 2^9 positional frequencies answers with 2e-2 .. 5e-2 colour differences on single rays (round-1 parity
// table).  The coarse net is 3 % of a frame's FLOPs, so it can afford three tensor-core products per layer:
//     x = x_hi + x_lo,  w = w_hi + w_lo   (fp16 pairs, |lo| <= 2^-11 |hi|)
//     x·w  ~=  x_hi·w_hi + x_lo·w_hi + x_hi·w_lo        (fp32 accumulation in TMEM; the dropped term is 2^-22)
// which leaves the layer outputs within ~1e-6 relative of fp32 SGEMM (the reference's arithmetic, models/model.py:121-137).
//
// Structure (per CTA pair = 256 points, persistent over tiles, cluster of 2, 384 threads per CTA):
//   * activations live in shared memory for all 23 layers as TWO swizzled K-major images (hi, lo: 2 x 64 KB per CTA) and
//     are updated IN PLACE: layer l's epilogue may overwrite its own input because every MMA that read it has retired
//     (it waits for that layer's tcgen05.commit);
//   * weights stream from L2 through a ring of six 16 KB slots: per 64-wide K block the hi image then the lo image, each
//     CTA loading its half of the N rows (cta_group::2: the pair's tensor cores read both halves, so L2->SM bytes per
//     point are those of the single-precision kernel although the weights are twice as large);
//   * accumulators double-buffer in TMEM, so layer l+1's MMAs on K block j start when layer l's epilogue has produced
//     column block j (per-block mbarriers; the peer CTA's warps arrive remotely on the leader's barriers);
//   * no second A operand anywhere: the skip layers' partial product  W[:, x_in] · x_in  is computed while x_in is
//     still the resident activation (a "virtual" layer right after its producer), parked as fp32 in a per-CTA scratch
//     (L2-resident, coalesced) and added in the skip layer's epilogue; the view-direction columns of the view layer are
//     a per-ray fp32 vector (view_vec_kernel) added in that layer's epilogue — exact, and one K block less;
//   * alpha_linear / rgb_linear are fp32 dot products of the un-rounded activations in the epilogue.
// Layer wiring comes from the engine's program (SplitLayerDesc table built in engine.cu).
#include "engine.h"
#include "pair.cuh"
#include "ptx.cuh"

namespace mofa {

constexpr int kSplitSlots = 6;
constexpr int kSlotBytes = 128 * 64 * 2;          // one CTA's half of a [256 x 64] weight K-block
constexpr int kAKb = 128 * 64 * 2;                // one K-block of this CTA's activation rows

struct SplitSmem {
  static constexpr int OFF_HI = 0;                                     // 4 K-blocks, 64 KB
  static constexpr int OFF_LO = 4 * kAKb;                              // 64 KB
  static constexpr int OFF_RING = 8 * kAKb;                            // 6 x 16 KB
  static constexpr int OFF_BAR = OFF_RING + kSplitSlots * kSlotBytes;
  // barriers: full[6] (leader), empty[6], x0_full (leader), mma_done[2], act_ready[4] (leader)
  static constexpr int N_BARS = 2 * kSplitSlots + 1 + 2 + 4;
  static constexpr int OFF_TPTR = OFF_BAR + N_BARS * 8;
  static constexpr int OFF_HEAD = OFF_TPTR + 16;                       // [128][3] fp32: head partials of epilogue group 1
  static constexpr int TOTAL = OFF_HEAD + 128 * 3 * 4;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};
static_assert(SplitSmem::DYN_BYTES <= 232448, "split coarse kernel exceeds the 227 KB shared-memory limit");

// The epilogue finishes the column blocks in ascending order (both groups work on the same block), and a layer consumes
// its K blocks in that order.
__device__ __forceinline__ int split_kb_order(int i, int kb) { (void)kb; return i; }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
coarse_split_kernel(const __grid_constant__ CUtensorMap tmX0hi, const __grid_constant__ CUtensorMap tmX0lo,
                    const CUtensorMap* __restrict__ wmaps, const SplitLayerDesc* __restrict__ layers, int n_layers,
                    int num_tiles, int64_t P_rows, const float* __restrict__ w_alpha, const float* __restrict__ b_alpha,
                    const float* __restrict__ w_rgb, const float* __restrict__ b_rgb,
                    const float* __restrict__ ray_vec, int rows_per_group, float* __restrict__ park,
                    float* __restrict__ raw) {
  using L = SplitSmem;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  const uint32_t act_hi = base + L::OFF_HI;
  const uint32_t act_lo = base + L::OFF_LO;
  const uint32_t ring0 = base + L::OFF_RING;
  const uint32_t full0 = base + L::OFF_BAR;
  const uint32_t empty0 = full0 + 8 * kSplitSlots;
  const uint32_t x0_full = empty0 + 8 * kSplitSlots;
  const uint32_t mma_done0 = x0_full + 8;           // [2] by accumulator stage
  const uint32_t act_ready0 = mma_done0 + 16;       // [4] by column / K block
  const uint32_t tptr = base + L::OFF_TPTR;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX0hi);
    prefetch_tmap(&tmX0lo);
    for (int i = 0; i < kSplitSlots; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
    }
    mbar_init(x0_full, 1);
    mbar_init(mma_done0, 1);
    mbar_init(mma_done0 + 8, 1);
    for (int i = 0; i < 4; ++i) mbar_init(act_ready0 + 8 * i, 16);  // one arrival per epilogue warp, both CTAs
    fence_mbar_init();
  }
  __syncthreads();
  cluster_sync_all();                                // barriers of both CTAs exist before any remote use
  if (warp == 2) {
    tmem_alloc_cg2(tptr, 512);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + L::OFF_TPTR);

  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp < 4) {
    setmaxnreg_dec_40();
    if (warp == 0) {
      // ---------------------------------------------------------------- producer (both CTAs)
      int slot = 0;
      uint32_t phase = 0;
      uint32_t done_cnt[2] = {0, 0};
      uint32_t g = 0;                                 // running layer counter: accumulator stage = g & 1
      for (int t = pair; t < num_tiles; t += num_pairs) {
        const int m0 = t * 256 + static_cast<int>(rank) * 128;
        // K block 0 of the activation images <- point encoding (hi, lo).  The images are free: first tile, or the previous
        // tile's last MMA has retired (waited for at the end of the previous iteration).
        if (elect_one()) {
          if (leader) mbar_expect_tx(x0_full, 4 * kAKb);          // hi + lo of both CTAs
          const uint32_t xb = mapa_u32(x0_full, 0);
          tma_load_2d_cg2(act_hi, &tmX0hi, xb, 0, m0);
          tma_load_2d_cg2(act_lo, &tmX0lo, xb, 0, m0);
        }
        __syncwarp();
        for (int l = 0; l < n_layers; ++l) {
          const SplitLayerDesc d = layers[l];
          const int half_n = d.n_out >> 1;
          const uint32_t bytes = static_cast<uint32_t>(half_n) * 128u;     // this CTA's rows of one K block
          for (int i = 0; i < 2 * d.kb; ++i) {
            const int kb = split_kb_order(i >> 1, d.kb);
            mbar_wait(empty0 + 8 * slot, phase ^ 1u);
            const uint32_t fb_local = full0 + 8 * slot;
            const uint32_t fb = mapa_u32(fb_local, 0);
            if (elect_one()) {
              if (leader) mbar_expect_tx(fb_local, 2 * bytes);            // bytes of both CTAs land on this barrier
              tma_load_2d_cg2(ring0 + slot * kSlotBytes, wmaps + ((i & 1) ? d.map_lo : d.map_hi), fb, kb * 64,
                              static_cast<int>(rank) * half_n);
            }
            __syncwarp();
            if (++slot == kSplitSlots) {
              slot = 0;
              phase ^= 1u;
            }
          }
          ++done_cnt[g & 1];
          ++g;
        }
        // the next tile's encoding overwrites K block 0, read by this tile's MMAs: wait for the last layer to retire
        {
          const uint32_t a = (g - 1) & 1;
          mbar_wait(mma_done0 + 8 * a, (done_cnt[a] - 1) & 1u);
        }
      }
    } else if (warp == 1 && leader) {
      // ---------------------------------------------------------------- MMA issuer (leader CTA)
      int slot = 0;
      uint32_t phase = 0;
      uint32_t x0_ph = 0, ready_ph = 0;
      uint32_t g = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        for (int l = 0; l < n_layers; ++l, ++g) {
          const SplitLayerDesc d = layers[l];
          const uint32_t d_tmem = tmem_base + (g & 1) * 256;
          const uint32_t idesc = umma_idesc_f16_f32(256, d.n_out);
          const uint32_t done_bar = mma_done0 + 8 * (g & 1);
          for (int kbi = 0; kbi < d.kb; ++kbi) {
            const int kb = split_kb_order(kbi, d.kb);
            if (l == 0) {
              if (kbi == 0) {
                mbar_wait(x0_full, x0_ph);
                x0_ph ^= 1u;
              }
            } else if (d.wait_act) {
              mbar_wait(act_ready0 + 8 * kb, ready_ph);           // both CTAs' epilogues have written K block kb
            }
            const uint64_t a_hi = umma_desc_sw128_kmajor(act_hi + kb * kAKb);
            const uint64_t a_lo = umma_desc_sw128_kmajor(act_lo + kb * kAKb);
            // ---- hi weights: x_hi·w_hi + x_lo·w_hi
            mbar_wait(full0 + 8 * slot, phase);
            tc_fence_after();
            {
              const uint64_t b = umma_desc_sw128_kmajor(ring0 + slot * kSlotBytes);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss_cg2(d_tmem, a_hi + 2 * k, b + 2 * k, idesc, (kbi | k) != 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss_cg2(d_tmem, a_lo + 2 * k, b + 2 * k, idesc, 1u);
                umma_commit_cg2_mc(empty0 + 8 * slot, 0x3);
              }
              __syncwarp();
              if (++slot == kSplitSlots) {
                slot = 0;
                phase ^= 1u;
              }
            }
            // ---- lo weights: x_hi·w_lo
            mbar_wait(full0 + 8 * slot, phase);
            tc_fence_after();
            {
              const uint64_t b = umma_desc_sw128_kmajor(ring0 + slot * kSlotBytes);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss_cg2(d_tmem, a_hi + 2 * k, b + 2 * k, idesc, 1u);
                umma_commit_cg2_mc(empty0 + 8 * slot, 0x3);
                if (kbi == d.kb - 1) umma_commit_cg2_mc(done_bar, 0x3);   // accumulator halves ready in both CTAs
              }
              __syncwarp();
              if (++slot == kSplitSlots) {
                slot = 0;
                phase ^= 1u;
              }
            }
          }
          if (l > 0 && d.wait_act) ready_ph ^= 1u;     // one completion of the four block barriers per stored activation
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 8 warps = 2 column groups x 4 lane quadrants
    // (12 warps in 512-thread CTAs were measured too: 2.33 ms instead of 2.16 ms per 265 k points — the compiler is held to
    //  128 registers per thread by the launch bound and the third warp per scheduler does not make up for the spills)
    setmaxnreg_inc_232();
    const int grp = (warp - 4) >> 2;
    const int ew = warp & 3;                        // TMEM lane quadrant this warp may read
    const int row = ew * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    float* head_sh = reinterpret_cast<float*>(base_ptr + L::OFF_HEAD);
    float4* park4 = reinterpret_cast<float4*>(park) + static_cast<size_t>(blockIdx.x) * (128 * 64);
    uint32_t done_ph[2] = {0, 0};
    uint32_t g = 0;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      const int64_t grow = static_cast<int64_t>(t) * 256 + static_cast<int64_t>(rank) * 128 + row;
      const int64_t vrow = grow < P_rows ? grow : (P_rows - 1);
      const float4* rv4 = reinterpret_cast<const float4*>(ray_vec + (vrow / rows_per_group) * 128);
      for (int l = 0; l < n_layers; ++l, ++g) {
        const SplitLayerDesc d = layers[l];
        const uint32_t acc = g & 1;
        mbar_wait(mma_done0 + 8 * acc, done_ph[acc]);
        done_ph[acc] ^= 1u;
        tc_fence_after();
        float hacc[3] = {0.f, 0.f, 0.f};
        const float* hw = d.head == 1 ? w_alpha : w_rgb;
        const int hn = d.head == 1 ? 1 : (d.head == 2 ? 3 : 0);
        const int ncb = d.n_out >> 6;
        // This warp's 32-column chunks of the tile: half `grp` of every 64-column block, blocks in ascending order — both
        // groups work on block 0 first, so the next layer's K block 0 is complete after a quarter of the epilogue, block 1
        // after half, ... (with each group owning two whole blocks the first blocks were ready only at half time and the
        // MMAs of the next layer waited for them).  Chunk q covers columns [q * 64 + grp * 32, +32).  The
        // additive terms of chunk q + 1 (bias, parked skip partial or per-ray view vector) are fetched while chunk q is
        // being processed, and chunk q's own TMEM load is issued before anything else, so the only exposed latency per
        // chunk is the TMEM load itself (with every fetch after the TMEM wait the epilogue was 75 % stalled on them).
        const int nchunk = ncb;
        const int col0 = grp * 32;
        auto fetch_add = [&](int ncol, float4 (&ad)[8]) {
          if (d.park) return;
          const float4* bias4 = reinterpret_cast<const float4*>(d.bias + ncol);
#pragma unroll
          for (int i = 0; i < 8; ++i) ad[i] = __ldg(bias4 + i);
          if (d.add_park) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 p = park4[(ncol / 4 + i) * 128 + row];
              ad[i].x += p.x; ad[i].y += p.y; ad[i].z += p.z; ad[i].w += p.w;
            }
          } else if (d.add_ray) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 p = __ldg(rv4 + ncol / 4 + i);
              ad[i].x += p.x; ad[i].y += p.y; ad[i].z += p.z; ad[i].w += p.w;
            }
          }
        };
        float4 ad[8], adn[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ad[i] = adn[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        fetch_add(col0, ad);
        // one chunk; `cur` holds its additive terms, `nxt` receives the next chunk's (the caller alternates the two
        // arrays, so no register copies are needed between chunks)
        auto do_chunk = [&](const int q, float4 (&cur)[8], float4 (&nxt)[8]) {
          const int ncol = col0 + q * 64;
          const int cb = q, h = grp;
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + lane_base + acc * 256 + ncol, v);
          if (q + 1 < nchunk) fetch_add(ncol + 64, nxt);
          tmem_ld_wait();
          if (d.park) {            // virtual layer: raw accumulators to the scratch, nothing else
#pragma unroll
            for (int i = 0; i < 8; ++i)
              park4[(ncol / 4 + i) * 128 + row] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                              __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            return;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 a0 = cur[2 * j], a1 = cur[2 * j + 1];
            float f[8];
            f[0] = fmaxf(__uint_as_float(v[j * 8 + 0]) + a0.x, 0.f);
            f[1] = fmaxf(__uint_as_float(v[j * 8 + 1]) + a0.y, 0.f);
            f[2] = fmaxf(__uint_as_float(v[j * 8 + 2]) + a0.z, 0.f);
            f[3] = fmaxf(__uint_as_float(v[j * 8 + 3]) + a0.w, 0.f);
            f[4] = fmaxf(__uint_as_float(v[j * 8 + 4]) + a1.x, 0.f);
            f[5] = fmaxf(__uint_as_float(v[j * 8 + 5]) + a1.y, 0.f);
            f[6] = fmaxf(__uint_as_float(v[j * 8 + 6]) + a1.z, 0.f);
            f[7] = fmaxf(__uint_as_float(v[j * 8 + 7]) + a1.w, 0.f);
            if (hn > 0) {
#pragma unroll
              for (int qq = 0; qq < 3; ++qq) {
                if (qq >= hn) break;
                const float4* w4 = reinterpret_cast<const float4*>(hw + qq * d.n_out + ncol) + 2 * j;
                const float4 w0 = __ldg(w4), w1 = __ldg(w4 + 1);
                hacc[qq] += f[0] * w0.x + f[1] * w0.y + f[2] * w0.z + f[3] * w0.w + f[4] * w1.x + f[5] * w1.y +
                            f[6] * w1.z + f[7] * w1.w;
              }
            }
            if (d.store) {
              // hi = fp16(x) (saturated at the fp16 maximum), lo = fp16(x - hi): the pair carries ~22 mantissa bits
              uint32_t ph[4], pl[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(ph[e]) : "f"(f[2 * e + 1]), "f"(f[2 * e]));
                asm("min.f16x2 %0, %0, %1;" : "+r"(ph[e]) : "r"(0x7bff7bffu));
                const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&ph[e]));
                asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pl[e]) : "f"(f[2 * e + 1] - hf.y), "f"(f[2 * e] - hf.x));
                asm("min.f16x2 %0, %0, %1;" : "+r"(pl[e]) : "r"(0x7bff7bffu));
              }
              const int chunk = h * 4 + j;
              const uint32_t off = cb * kAKb + row * 128 + ((chunk ^ (row & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(act_hi + off), "r"(ph[0]), "r"(ph[1]),
                           "r"(ph[2]), "r"(ph[3])
                           : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(act_lo + off), "r"(pl[0]), "r"(pl[1]),
                           "r"(pl[2]), "r"(pl[3])
                           : "memory");
            }
          }
          if (d.store) {                  // this warp's half of K block cb of the next layer's A operand is complete
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            // The peer CTA arrives on the leader's barrier with the default (CTA-scope release) semantics, as the pair
            // kernel's accumulator hand-off does: the data are this thread's own st.shared, made visible to the async
            // proxy by the fence above, and the MMA that reads them is issued only after the leader has seen the arrive.
            // (The explicit .release.cluster form compiled to MEMBAR.ALL.GPU + CCTL.IVALL + ERRBAR: 20 % of the peer
            // epilogue's samples, and the L1 invalidation made every bias / park fetch that followed miss.)
            if (lane == 0) {
              if (leader) mbar_arrive(act_ready0 + 8 * cb);
              else mbar_arrive_cluster(mapa_u32(act_ready0 + 8 * cb, 0));
            }
          }
        };
#pragma unroll 1
        for (int q = 0; q < nchunk; q += 2) {
          do_chunk(q, ad, adn);
          do_chunk(q + 1, adn, ad);        // n_out is 128 or 256: always an even number of chunks
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
        tc_fence_before();
      }
    }
  }
  __syncthreads();
  cluster_sync_all();                     // peer shared memory / TMEM stay valid until both CTAs are done
  if (warp == 2) tmem_dealloc_cg2(tmem_base, 512);
}

cudaError_t coarse_split_configure() {
  return cudaFuncSetAttribute(coarse_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SplitSmem::DYN_BYTES);
}

size_t coarse_split_park_bytes(int num_sms) { return static_cast<size_t>(num_sms) * 128 * 256 * sizeof(float); }

cudaError_t launch_coarse_split(const SplitLaunch& S, int num_sms, cudaStream_t stream) {
  const int num_tiles = static_cast<int>((S.P_rows + 255) / 256);
  if (num_tiles <= 0) return cudaSuccess;
  const int max_pairs = num_sms / 2;
  const int pairs = num_tiles < max_pairs ? num_tiles : max_pairs;
  coarse_split_kernel<<<2 * pairs, 384, SplitSmem::DYN_BYTES, stream>>>(
      S.tmX0hi, S.tmX0lo, S.wmaps, S.layers, S.n_layers, num_tiles, S.P_rows, S.w_alpha, S.b_alpha, S.w_rgb, S.b_rgb,
      S.ray_vec, S.rows_per_group, S.park, S.raw);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Per-ray view vector: out[g, c] = sum_k Wv[c, k] * PE_4(viewdir_g)[k]   (the 27 view-direction columns of
// linear_view_xyBMuv, models/model.py:133 with input order cat[views, rgbCodes]; PE as models/model.py:24-45)
// ------------------------------------------------------------------------------------------------
__global__ void view_vec_kernel(const float* __restrict__ dirs, int stride, int64_t n, const float* __restrict__ Wv,
                                int n_out, float* __restrict__ out) {
  __shared__ float pe[8][28];
  const int64_t g0 = static_cast<int64_t>(blockIdx.x) * 8;
  if (threadIdx.x < 8 * 3) {
    const int r = threadIdx.x / 3, c = threadIdx.x % 3;
    const int64_t gi = g0 + r;
    if (gi < n) {
      const float x = dirs[gi * stride + c];
      pe[r][c] = x;
      for (int f = 0; f < 4; ++f) {
        float s, co;
        sincosf(x * static_cast<float>(1 << f), &s, &co);
        pe[r][3 + 6 * f + c] = s;
        pe[r][3 + 6 * f + 3 + c] = co;
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < n_out; c += blockDim.x) {
    float w[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) w[k] = Wv[c * 27 + k];
    for (int r = 0; r < 8; ++r) {
      if (g0 + r >= n) break;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 27; ++k) acc += w[k] * pe[r][k];
      out[(g0 + r) * n_out + c] = acc;
    }
  }
}

cudaError_t launch_view_vec(const float* dirs, int stride, int64_t n, const float* Wv, int n_out, float* out,
                            cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  view_vec_kernel<<<static_cast<unsigned>((n + 7) / 8), 128, 0, s>>>(dirs, stride, n, Wv, n_out, out);
  return cudaGetLastError();
}

}  // namespace mofa
