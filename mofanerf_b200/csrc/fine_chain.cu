// Fine network as ONE persistent kernel per pass: all dense layers of a W >= 512 net (27 nn.Linear of the W = 1024 fine
// net: models/model.py:121-137, 226-230) chained inside a single launch of 74 CTA pairs, activations L2-resident.
//
// Why: launched layer by layer over a 4144-ray chunk, every 1024 -> 1024 layer streams its 1.09 GB input from HBM and its
// 1.05 GB output back (ncu: 2.14 GB DRAM traffic per 1.11-TFLOP launch, 2.8 TB/s) — not time-limiting (DRAM 36 %), but the
// GPU sits at its 1 kW power cap and every HBM byte is paid for in SM clock (round-1 verdict: 0.80 pJ/FLOP against cuBLAS's
// 0.73).  The only traffic the algorithm needs is 365 KB of rays per chunk.
//
// How: the point rows are processed in SLABS of 56 m-blocks (14336 rows = 112 rays x 128 samples).  For one slab the
// three rotating activation buffers are 3 x 29 MB — together with the layer's weights they stay inside the 126 MB L2 —
// and every slab re-uses the SAME physical buffers, so activations are written and read back through L2 and only a
// small remainder reaches HBM (ncu, whole 25-layer pass over 4144 rays: 8.2 GB of DRAM traffic instead of 53 GB).
// Work item = (slab, layer, m-block, n-tile) = one 256 x 256 output tile of one layer; items are numbered slab-major,
// then layer, then m-block, then n-tile, and pair p takes items p, p + 74, p + 148, ... — a layer of a full slab is about
// three rounds of the 74 pairs.  A tile of layer l needs the four n-tiles of layer l-1 of ITS m-block only, which lie
// three rounds back in that order: by the time a pair's producer warp polls the m-block's completion counter it is
// almost always already there (skipping the waits altogether changes the frame time by ~1 %).  Counters are bumped by
// the epilogue (red.release.gpu) after the tile's TMA stores have completed (cp.async.bulk.wait_group 0) and a proxy
// fence; the consumer fences the async proxy again before its TMA loads.  Layer 0 of slab s waits for the last layer of
// slab s-1 at the same local m-block (the buffers it overwrites are free then).  Every dependency points to a lower item
// number and all 74 pairs are co-resident, so the pair holding the lowest unfinished item can always run: no deadlock.
// Slab size is a trade: fewer m-blocks = fewer rounds between a tile and its dependencies (stalls), more = activations
// that no longer fit L2 (HBM traffic, i.e. power and clock): 37 / 56 / 74 / 111 m-blocks measured 162.2 / 168.6 / 166.1 /
// 157.6 k rays/s on one box (one launch per layer: 157-165 k).
//
// The tile pipeline itself is the CTA-pair kernel of dense_tc2.cu (cta_group::2 UMMA 256x256x16, 6-stage TMA ring,
// double-buffered TMEM accumulators, 8 epilogue warps, fused alpha / rgb heads); what changes per tile are the tensor
// maps (device array), the K extents, the bias / head pointers and the row coordinates (slab-local for the activation
// buffers, global for the point encodings and the head partials).
#include "dense_epilogue.cuh"
#include "engine.h"
#include "pair.cuh"
#include "ptx.cuh"

namespace mofa {

constexpr int kChainStages = 6;

struct ChainSmem {
  static constexpr int A_BYTES = 128 * 64 * 2;
  static constexpr int B_BYTES = 128 * 64 * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;          // per CTA
  static constexpr int C_BYTES = 128 * 64 * 2;
  static constexpr int OFF_C = kChainStages * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_C + 2 * C_BYTES;            // one staging buffer per epilogue group
  static constexpr int N_BARS = 2 * kChainStages + 4;
  static constexpr int OFF_TPTR = OFF_BAR + N_BARS * 8;
  static constexpr int OFF_LAYERS = OFF_TPTR + 16;              // compact per-layer table (48 bytes x 32 layers)
  static constexpr int TOTAL = OFF_LAYERS + 48 * 32;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};
static_assert(ChainSmem::DYN_BYTES <= 232448, "chain kernel exceeds the 227 KB shared-memory limit");

// What the three roles need per tile, kept in shared memory: fetched from global memory per tile it cost every role an
// L2 round trip at each tile start (the SM's L1 is invalidated twice per tile, see chain_epilogue) — for the MMA issuer
// that is tensor-pipe idle time.
struct ChainLayerSm {
  const float* bias;
  const float* head_w;
  const float* colscale;
  uint16_t mapA0, mapA1, mapB0, mapB1, mapC, N;
  uint8_t kb0, kb1, a0_global, a1_global, relu, store_c, head_n, head_slot0;
  uint8_t fp8_in, fp8_out, pad[2];
};
static_assert(sizeof(ChainLayerSm) == 48, "ChainLayerSm layout");

struct ChainTile {
  int layer, mb_local, mb_global, n;
};

// item number -> (slab, layer, m-block, n-tile).  Every layer but the last has nt n-tiles, the last has nt_last
// (checked on the host), so the layer follows from one division.
__device__ __forceinline__ ChainTile chain_decode(const ChainParams& p, long long t) {
  const int full = p.slab_mb * p.tiles_per_mb;
  const int s = static_cast<int>(t / full);
  const int r = static_cast<int>(t - static_cast<long long>(s) * full);
  const int left = p.total_mb - s * p.slab_mb;
  const int mb_s = left < p.slab_mb ? left : p.slab_mb;
  const int q = r / mb_s;                                   // in units of "n-tiles of one m-block"
  int l = q / p.nt;
  if (l > p.n_layers - 1) l = p.n_layers - 1;
  const int ntl = (l == p.n_layers - 1) ? p.nt_last : p.nt;
  const int r2 = r - l * p.nt * mb_s;
  ChainTile c;
  c.layer = l;
  c.mb_local = r2 / ntl;
  c.n = r2 - c.mb_local * ntl;
  c.mb_global = s * p.slab_mb + c.mb_local;
  return c;
}

// Counter poll: a relaxed GPU-scope load (LDG.STRONG.GPU, served by L2).  ld.acquire.gpu would add a CCTL.IVALL — an
// invalidation of the SM's whole L1 — to EVERY poll; the data the counter guards are read by TMA (async proxy, straight
// from L2, never through L1), so the only ordering needed is "counter read before the TMA loads are issued": program
// order of the issuing thread plus the proxy fence below.
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic <-> async proxy ordering for GLOBAL memory only (FENCE.VIEW.ASYNC.G; the unqualified form also emits a
// GPU-scope MEMBAR)
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Bounded spin on an m-block's completion counter (a protocol bug must trap, not hang the box).
__device__ __forceinline__ void wait_counter(const uint32_t* ctr, uint32_t need) {
  uint32_t spins = 0;
  long long t0 = 0;
  while (ld_relaxed_gpu(ctr) < need) {
    __nanosleep(64);
    if ((++spins & 0x3FFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) __trap();
    }
  }
}


// Epilogue of one tile for one epilogue group (4 warps x 32 rows, 128 of the tile's 256 columns = four 32-column chunks).
// Same arithmetic and staging as dense_epilogue.cuh (bit-identical results) with one difference that matters here: the
// bias of chunk q + 1 is fetched while chunk q is processed and chunk q's own fetches are issued before its TMEM wait.
// In this kernel the SM's L1 is invalidated twice per tile (cp.async.bulk.wait_group before a tile is published), so
// bias loads issued where they are used — fine in the per-layer kernels, where they hit L1 — each exposed an L2 round
// trip: the epilogue became the critical path (tensor pipe 63 % active, ncu).
// F8IN: the accumulator is an e4m3 x e4m3 product and is multiplied by colscale[col] before the bias.  F8OUT: the
// activation is written as e4m3 (x kFp8ActScale, saturating) in 32-byte rows instead of fp16 in 64-byte rows.
constexpr float kFp8ActScale = 8.0f;
template <int HEAD, bool F8IN, bool F8OUT>
__device__ __forceinline__ void chain_epilogue(const EpiParams& p, const float* __restrict__ colscale, const void* tmC,
                                               uint32_t acc_addr, const EpiGroup& g, int m0, int n0, int row,
                                               long long hrow0) {
  float hacc[3] = {0.f, 0.f, 0.f};
  const int col_base = n0 + g.cb0 * 64;
  float4 bcur[8], bnext[8];
  {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col_base);
#pragma unroll
    for (int i = 0; i < 8; ++i) bcur[i] = ldg_f4_pinned(b4 + i);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int cb = g.cb0 + (q >> 1), h = q & 1;
    const int ncol = col_base + q * 32;
    (void)h;
    // staging: two 8 KB half-buffers per group (32 columns x 128 rows, 64-byte rows, SWIZZLE_64B), alternating per chunk,
    // so the TMA store of chunk q reads its buffer while chunk q + 1 is computed into the other one
    const uint32_t cbuf = g.cbuf0 + (q & 1) * (128 * 32 * 2);
    uint32_t v[32];
    tmem_ld_32x32b_x32(acc_addr + cb * 64 + h * 32, v);
    float4 cs[F8IN ? 8 : 1];
    if constexpr (F8IN) {
      const float4* c4 = reinterpret_cast<const float4*>(colscale + ncol);
#pragma unroll
      for (int i = 0; i < 8; ++i) cs[i] = ldg_f4_pinned(c4 + i);
    }
    constexpr int HB = (HEAD == 3) ? 1 : 4;           // 8-column blocks of head weights per fetch
    constexpr int NHW = HEAD > 0 ? HEAD * 2 * HB : 1;
    float4 hw[2][NHW];
    if constexpr (HEAD > 0) {
#pragma unroll
      for (int qq = 0; qq < HEAD; ++qq) {
        const float4* w4 = reinterpret_cast<const float4*>(p.head_w + static_cast<size_t>(qq) * p.N + ncol);
#pragma unroll
        for (int i = 0; i < 2 * HB; ++i) hw[0][qq * 2 * HB + i] = ldg_f4_pinned(w4 + i);
      }
    }
    if constexpr (HEAD == 0) {        // (the head layers keep their registers for the head weights: own chunk only)
      if (q + 1 < 4) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + ncol + 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) bnext[i] = ldg_f4_pinned(b4 + i);
      }
    } else if (q > 0) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + ncol);
#pragma unroll
      for (int i = 0; i < 8; ++i) bcur[i] = ldg_f4_pinned(b4 + i);
    }
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 b0 = bcur[2 * j], b1 = bcur[2 * j + 1];
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float x = __uint_as_float(v[j * 8 + e]);
        if constexpr (F8IN) {
          const float4 c0 = cs[2 * j], c1 = cs[2 * j + 1];
          const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
          x = x * cc[e] + bb[e];
        } else {
          x += bb[e];
        }
        if (p.relu) x = fmaxf(x, 0.0f);
        f[e] = fminf(fmaxf(x, -65504.0f), 65504.0f);
      }
      if constexpr (HEAD > 0 && HB == 1) {
        if (j + 1 < 4) {
#pragma unroll
          for (int qq = 0; qq < HEAD; ++qq) {
            const float4* w4 = reinterpret_cast<const float4*>(p.head_w + static_cast<size_t>(qq) * p.N + ncol) + 2 * (j + 1);
            hw[(j + 1) & 1][qq * 2] = ldg_f4_pinned(w4);
            hw[(j + 1) & 1][qq * 2 + 1] = ldg_f4_pinned(w4 + 1);
          }
        }
      }
      if constexpr (HEAD > 0) {
#pragma unroll
        for (int qq = 0; qq < HEAD; ++qq) {
          const float4 w0 = HB == 1 ? hw[j & 1][qq * 2] : hw[0][qq * 2 * HB + 2 * j];
          const float4 w1 = HB == 1 ? hw[j & 1][qq * 2 + 1] : hw[0][qq * 2 * HB + 2 * j + 1];
          hacc[qq] += f[0] * w0.x + f[1] * w0.y + f[2] * w0.z + f[3] * w0.w + f[4] * w1.x + f[5] * w1.y +
                      f[6] * w1.z + f[7] * w1.w;
        }
      }
      if constexpr (F8OUT) {
        if (p.store_c) {      // 8 columns -> 8 bytes of this row's 32-byte staging row (no swizzle)
          uint16_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(pk[e]) : "f"(f[2 * e + 1] * kFp8ActScale), "f"(f[2 * e] * kFp8ActScale));
          const uint32_t w0 = static_cast<uint32_t>(pk[0]) | (static_cast<uint32_t>(pk[1]) << 16);
          const uint32_t w1 = static_cast<uint32_t>(pk[2]) | (static_cast<uint32_t>(pk[3]) << 16);
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(cbuf + row * 32 + j * 8), "r"(w0), "r"(w1) : "memory");
        }
      } else if (p.store_c) {
        __half2 h0 = __floats2half2_rn(f[0], f[1]);
        __half2 h1 = __floats2half2_rn(f[2], f[3]);
        __half2 h2 = __floats2half2_rn(f[4], f[5]);
        __half2 h3 = __floats2half2_rn(f[6], f[7]);
        // 64-byte swizzle: 16-byte chunk index (2 bits) xor address bits [7,9) = (row >> 1) & 3
        const uint32_t addr = cbuf + row * 64 + ((j ^ ((row >> 1) & 3)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                     "r"(*reinterpret_cast<uint32_t*>(&h0)), "r"(*reinterpret_cast<uint32_t*>(&h1)),
                     "r"(*reinterpret_cast<uint32_t*>(&h2)), "r"(*reinterpret_cast<uint32_t*>(&h3))
                     : "memory");
      }
    }
    if (p.store_c) {
      fence_proxy_async_smem();
      // one barrier per chunk: before it, thread 0 makes sure the store of chunk q - 1 has finished READING its buffer —
      // the buffer chunk q + 1 will be written into — so that nobody has to wait for a store that was only just issued
      if (g.gtid == 0) tma_store_wait_read<0>();
      named_bar_sync(g.bar_id, 128);
      if (g.gtid == 0) {
        tma_store_2d(tmC, cbuf, ncol, m0);
        tma_store_commit();
      }
    }
    if constexpr (HEAD == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) bcur[i] = bnext[i];
    }
  }
  if constexpr (HEAD > 0) {
    if (hrow0 + row < p.M) {
      float* dst = p.head_out + static_cast<size_t>(hrow0 + row) * p.head_stride + p.head_slot0 + g.slot * HEAD;
#pragma unroll
      for (int qq = 0; qq < HEAD; ++qq) dst[qq] = hacc[qq];
    }
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
fine_chain_kernel(const ChainParams p) {
  using L = ChainSmem;
  constexpr int STAGES = kChainStages;
  constexpr int BN = 256;
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
  const uint32_t tempty0 = tfull0 + 16;                // used in the leader only (8 warp arrivals per CTA)
  const uint32_t tptr = base + L::OFF_TPTR;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, 16);
    }
    fence_mbar_init();
  }
  {
    ChainLayerSm* lt = reinterpret_cast<ChainLayerSm*>(base_ptr + L::OFF_LAYERS);
    for (int i = threadIdx.x; i < p.n_layers; i += blockDim.x) {
      const ChainLayerDesc d = p.layers[i];
      ChainLayerSm q;
      q.bias = d.bias; q.head_w = d.head_w;
      q.mapA0 = d.mapA0; q.mapA1 = d.mapA1; q.mapB0 = d.mapB0; q.mapB1 = d.mapB1; q.mapC = d.mapC; q.N = d.N;
      q.kb0 = d.kb0; q.kb1 = d.kb1; q.a0_global = d.a0_global; q.a1_global = d.a1_global; q.relu = d.relu;
      q.store_c = d.store_c; q.head_n = d.head_n; q.head_slot0 = d.head_slot0;
      q.colscale = d.colscale;
      q.fp8_in = d.fp8_in; q.fp8_out = d.fp8_out;
      q.pad[0] = q.pad[1] = 0;
      lt[i] = q;
    }
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tmem_alloc_cg2(tptr, 512);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + L::OFF_TPTR);

  ChainLayerSm* lsm = reinterpret_cast<ChainLayerSm*>(base_ptr + L::OFF_LAYERS);
  const long long num_items = static_cast<long long>(p.total_mb) * p.tiles_per_mb;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp < 4) {
    setmaxnreg_dec_40();
    if (warp == 0) {
      // ------------------------------------------------------------------ TMA producer (both CTAs)
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = pair; t < num_items; t += num_pairs) {
        const ChainTile c = chain_decode(p, t);
        const ChainLayerSm d = lsm[c.layer];
        // dependency: the previous layer's n-tiles of this m-block (or, for layer 0 of a later slab, the last layer of
        // the previous slab at the same local m-block: the activation buffers it still reads are overwritten from here on)
        if (lane == 0 && !p.nodep) {
          if (c.layer > 0) {
            wait_counter(p.counters + static_cast<size_t>(c.layer - 1) * p.total_mb + c.mb_global, 4u * p.nt);
          } else if (c.mb_global >= p.slab_mb) {
            wait_counter(p.counters + static_cast<size_t>(p.n_layers - 1) * p.total_mb + (c.mb_global - p.slab_mb),
                         4u * p.nt_last);
          }
        }
        __syncwarp();
        fence_proxy_async_global();   // the loads below are async-proxy reads of data other CTAs' TMA stores wrote
        const int ml = c.mb_local * 256 + static_cast<int>(rank) * 128;
        const int mg = c.mb_global * 256 + static_cast<int>(rank) * 128;
        const int m_a0 = d.a0_global ? mg : ml;
        const int m_a1 = d.a1_global ? mg : ml;
        const int n0 = c.n * BN + static_cast<int>(rank) * 128;
        const int total_kb = d.kb0 + d.kb1;
        const int kstep = d.fp8_in ? 128 : 64;       // elements per 128-byte K block (e4m3 / fp16)
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(empty0 + 8 * stage, phase ^ 1u);
          const uint32_t fb_local = full0 + 8 * stage;
          const uint32_t fb = mapa_u32(fb_local, 0);
          const uint32_t sa = base + stage * L::STAGE_BYTES;
          const uint32_t sb = sa + L::A_BYTES;
          if (elect_one()) {
            if (leader) mbar_expect_tx(fb_local, 2 * L::STAGE_BYTES);
            if (kb < d.kb0) {
              tma_load_2d_cg2(sa, p.maps + d.mapA0, fb, kb * kstep, m_a0);
              tma_load_2d_cg2(sb, p.maps + d.mapB0, fb, kb * kstep, n0);
            } else {
              const int k = (kb - d.kb0) * kstep;
              tma_load_2d_cg2(sa, p.maps + d.mapA1, fb, k, m_a1);
              tma_load_2d_cg2(sb, p.maps + d.mapB1, fb, k, n0);
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
      // ------------------------------------------------------------------ MMA issuer (leader CTA)
      constexpr uint32_t idesc = umma_idesc_f16_f32(256, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (long long t = pair; t < num_items; t += num_pairs, ++it) {
        const ChainTile c = chain_decode(p, t);
        const int total_kb = lsm[c.layer].kb0 + lsm[c.layer].kb1;
        const bool f8 = lsm[c.layer].fp8_in != 0;
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
            if (f8) {      // same descriptors and K stepping (32 bytes per instruction): only the operand kind differs
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f8_ss_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16_ss_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_cg2_mc(empty0 + 8 * stage, 0x3);
            if (kb == total_kb - 1) umma_commit_cg2_mc(tfull0 + 8 * as, 0x3);
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
    // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows; two column groups)
    setmaxnreg_inc_232();
    const int ew = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int row = ew * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    EpiGroup g;
    g.cbuf0 = base + L::OFF_C + grp * L::C_BYTES;
    g.cb0 = grp * 2;
    g.cb1 = g.cb0 + 2;
    g.bar_id = 1 + grp;
    g.gtid = static_cast<int>(threadIdx.x) - 128 - grp * 128;
    int it = 0;
    for (long long t = pair; t < num_items; t += num_pairs, ++it) {
      const ChainTile c = chain_decode(p, t);
      const ChainLayerSm d = lsm[c.layer];
      EpiParams e;
      e.bias = d.bias;
      e.head_w = d.head_w;
      e.head_out = p.head_out;
      e.relu = d.relu;
      e.store_c = d.store_c;
      e.head_n = d.head_n;
      e.head_stride = p.head_stride;
      e.head_slot0 = d.head_slot0;
      e.N = d.N;
      e.M = static_cast<int>(p.P_rows);
      e.mask = nullptr;
      e.r1_row = nullptr;
      e.r1_col = nullptr;
      e.r1_stride = 0;
      const int ml = c.mb_local * 256 + static_cast<int>(rank) * 128;
      const long long mg = static_cast<long long>(c.mb_global) * 256 + static_cast<long long>(rank) * 128;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      g.slot = c.n * 2 + grp;
      mbar_wait(tfull0 + 8 * as, aphase);
      tc_fence_after();
      {
        const uint32_t acc_addr = tmem_base + lane_base + as * BN;
        const void* mc = p.maps + d.mapC;
        const int nn = c.n * BN;
        if (d.head_n == 0) {
          if (!d.fp8_in && !d.fp8_out) chain_epilogue<0, false, false>(e, d.colscale, mc, acc_addr, g, ml, nn, row, mg);
          else if (d.fp8_in && d.fp8_out) chain_epilogue<0, true, true>(e, d.colscale, mc, acc_addr, g, ml, nn, row, mg);
          else if (d.fp8_in) chain_epilogue<0, true, false>(e, d.colscale, mc, acc_addr, g, ml, nn, row, mg);
          else chain_epilogue<0, false, true>(e, d.colscale, mc, acc_addr, g, ml, nn, row, mg);
        } else if (d.head_n == 1) {
          if (d.fp8_in) chain_epilogue<1, true, false>(e, d.colscale, mc, acc_addr, g, ml, nn, row, mg);
          else chain_epilogue<1, false, false>(e, d.colscale, mc, acc_addr, g, ml, nn, row, mg);
        } else {
          chain_epilogue<3, false, false>(e, d.colscale, mc, acc_addr, g, ml, nn, row, mg);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(tempty0 + 8 * as, 0));
      // publish: this group's part of the tile (128 rows x 128 columns) is in global memory.  Thread 0 of the group issued
      // the TMA stores: wait for their completion (writes performed, not just the staging buffer read), order them before
      // the counter update across proxies, then release.
      if (g.gtid == 0) {
        tma_store_wait_all<0>();
        fence_proxy_async_global();
        red_release_gpu_add(p.counters + static_cast<size_t>(c.layer) * p.total_mb + c.mb_global, 1u);
      }
    }
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_cg2(tmem_base, 512);
}

cudaError_t fine_chain_configure() {
  return cudaFuncSetAttribute(fine_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ChainSmem::DYN_BYTES);
}

cudaError_t launch_fine_chain(const ChainParams& p, int num_sms, cudaStream_t stream) {
  const long long items = static_cast<long long>(p.total_mb) * p.tiles_per_mb;
  if (items <= 0) return cudaSuccess;
  const int max_pairs = num_sms / 2;
  const int pairs = static_cast<int>(items < max_pairs ? items : max_pairs);
  fine_chain_kernel<<<2 * pairs, 384, ChainSmem::DYN_BYTES, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace mofa
