// Epilogue shared by the single-CTA and CTA-pair dense kernels: one 128-row x BN-column accumulator tile
//   TMEM --tcgen05.ld--> registers --(+bias, ReLU, clamp)--> fp16 --swizzled st.shared--> TMA store
// plus the optional fused output head (alpha_linear / rgb_linear, models/model.py:130,134): each thread owns one
// output row, so the head is a running dot product of the row's activated values with the head weights; the
// per-tile partial goes to its own slot (row, n_tile) of `head_out` with a plain store — deterministic, no
// atomics; finalize_raw_kernel adds the slots and the head bias.
#pragma once
#include "ptx.cuh"

namespace mofa {

struct EpiParams {
  const float* bias;      // [N] or nullptr
  const float* head_w;    // [head_n, N] fp32 or nullptr
  float* head_out;        // [M, head_stride] fp32 partial slots
  int relu;
  int store_c;            // 0: the layer's activations are not needed (view layer with fused rgb head)
  int head_n;             // 0, 1 or 3
  int head_stride;        // floats per row in head_out
  int head_slot0;         // first slot of this head in a row
  int N;                  // layer width (row pitch of head_w)
  int M;                  // valid rows of head_out
  // backward-pass options: out = mask ⊙ (acc + r1_row[row] * r1_col[col])
  const __half* mask;     // [M, N] forward activation of the tensor whose gradient this is (ReLU'), or nullptr
  const float* r1_row;    // per-row factor (element stride r1_stride) of a rank-1 term, or nullptr
  const float* r1_col;    // [N] per-column factor
  int r1_stride;
};

// Read-only 16-byte load that the compiler may not move across the other volatile asm statements of the epilogue
// (TMEM load / wait, staging stores): pins WHERE a prefetch is issued, which __ldg does not.
__device__ __forceinline__ float4 ldg_f4_pinned(const float4* ptr) {
  float4 r;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(ptr));
  return r;
}

__device__ __forceinline__ uint4 ldg_u4_pinned(const uint4* ptr) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(ptr));
  return r;
}

// One epilogue GROUP = 4 warps (one per TMEM lane quadrant) that own a range of 64-column blocks of the tile, their
// own staging buffer(s), their own named barrier and their own head slot.  The single-CTA kernel and the 4-warp pair
// kernel run one group over all BN columns; the 8-warp pair kernel runs two groups on one half of the columns each.
struct EpiGroup {
  uint32_t cbuf0;   // shared address of this group's staging buffer(s), 16 KB each
  int cb0, cb1;     // 64-column blocks [cb0, cb1) of the tile
  int bar_id;       // named barrier shared by the group's 128 threads
  int slot;         // head partial slot of this (n-tile, group)
  int gtid;         // thread index within the group, 0..127
};

// acc_addr: TMEM address of this warp's lane quadrant at the accumulator stage's first column.
// cnt: running staging-buffer counter (NBUF == 2).  NBUF: staging buffers the group alternates between.
// HEAD = number of fused head outputs (0, 1, 3), a compile-time copy of p.head_n: the head weights of a 32-column
// chunk are fetched while the chunk's TMEM load is in flight.  (Fetched where they are used — inside a branch on
// p.head_n — every 8-column group exposed a full L1/L2 latency: ncu source page, 30 % of the sigma layer's time.)
template <int BN, bool BWD, int HEAD, int NBUF>
// hrow0: row of head_out that corresponds to the tile's first row (the fine-net chain kernel stores C into slab-local
// buffers but writes head partials per global point row); the per-layer kernels pass m0.
__device__ __forceinline__ void epilogue_tile_impl(const EpiParams& p, const void* tmC, uint32_t acc_addr,
                                                   const EpiGroup& g, uint32_t& cnt, int m0, int n0, int row,
                                                   long long hrow0) {
  float hacc[3] = {0.f, 0.f, 0.f};
  const float r1 = (BWD && p.r1_row != nullptr && m0 + row < p.M) ? p.r1_row[static_cast<size_t>(m0 + row) * p.r1_stride] : 0.0f;
  // Backward: the ReLU mask of a 32-column chunk (this row's 64 bytes of the forward activation, which comes from HBM —
  // the tensor is far larger than L2) is fetched ONE CHUNK AHEAD, while the previous chunk is being processed.  Fetched
  // where it is used (inside the 8-column loop) every block exposed a full memory latency: the dX kernel ran 47 % longer
  // than the forward kernel of the same layer.
  const bool use_mask = BWD && p.mask != nullptr && m0 + row < p.M;
  const __half* mrow = use_mask ? p.mask + static_cast<size_t>(m0 + row) * p.N + n0 : nullptr;
  uint4 mk_cur[4], mk_nxt[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) mk_cur[j] = mk_nxt[j] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);   // fp16 ones: keep everything
  if (use_mask) {
#pragma unroll
    for (int j = 0; j < 4; ++j) mk_cur[j] = ldg_u4_pinned(reinterpret_cast<const uint4*>(mrow + g.cb0 * 64) + j);
  }
#pragma unroll 1
  for (int cb = g.cb0; cb < g.cb1; ++cb) {
    const uint32_t cbuf = g.cbuf0 + (NBUF == 2 ? (cnt & 1u) * (128 * 64 * 2) : 0u);
    if (p.store_c) {
      if (g.gtid == 0) tma_store_wait_read<NBUF - 1>();    // the store that last read this buffer is done
      named_bar_sync(g.bar_id, 128);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(acc_addr + cb * 64 + h * 32, v);
      const int ncol = n0 + cb * 64 + h * 32;
      if (use_mask) {       // next chunk of this row: the other half of the block, or the first half of the next block
        const int nxt = cb * 64 + h * 32 + 32;
        if (nxt < g.cb1 * 64) {
#pragma unroll
          for (int j = 0; j < 4; ++j) mk_nxt[j] = ldg_u4_pinned(reinterpret_cast<const uint4*>(mrow + nxt) + j);
        }
      }
      // head weights: HEAD == 1 fetches the whole 32-column chunk here (32 registers); HEAD == 3 would need 96, so it
      // fetches one 8-column block ahead of the block being reduced (two rotating sets of 24 registers)
      constexpr int HB = (HEAD == 3) ? 1 : 4;           // 8-column blocks per fetch
      constexpr int NHW = HEAD > 0 ? HEAD * 2 * HB : 1;
      float4 hw[2][NHW];
      if constexpr (HEAD > 0) {
#pragma unroll
        for (int q = 0; q < HEAD; ++q) {
          const float4* w4 = reinterpret_cast<const float4*>(p.head_w + static_cast<size_t>(q) * p.N + ncol);
#pragma unroll
          for (int i = 0; i < 2 * HB; ++i) hw[0][q * 2 * HB + i] = ldg_f4_pinned(w4 + i);
        }
      }
      tmem_ld_wait();
      const float4* bias4 = reinterpret_cast<const float4*>(p.bias + ncol);   // 128-byte aligned (ncol % 32 == 0)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f[8];
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
        if (p.bias != nullptr) {
          b0 = __ldg(bias4 + 2 * j);
          b1 = __ldg(bias4 + 2 * j + 1);
        }
        float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        if (BWD && p.r1_row != nullptr) {
          const float4* c4 = reinterpret_cast<const float4*>(p.r1_col + ncol) + 2 * j;
          const float4 c0 = __ldg(c4), c1 = __ldg(c4 + 1);
          bb[0] += r1 * c0.x; bb[1] += r1 * c0.y; bb[2] += r1 * c0.z; bb[3] += r1 * c0.w;
          bb[4] += r1 * c1.x; bb[5] += r1 * c1.y; bb[6] += r1 * c1.z; bb[7] += r1 * c1.w;
        }
        const uint4 mk = mk_cur[j];
        const __half* mh = reinterpret_cast<const __half*>(&mk);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float x = __uint_as_float(v[j * 8 + e]) + bb[e];
          if (p.relu) x = fmaxf(x, 0.0f);
          if (BWD && !(__half2float(mh[e]) > 0.0f)) x = 0.0f;
          f[e] = fminf(fmaxf(x, -65504.0f), 65504.0f);
        }
        if constexpr (HEAD > 0 && HB == 1) {
          if (j + 1 < 4) {
#pragma unroll
            for (int q = 0; q < HEAD; ++q) {
              const float4* w4 = reinterpret_cast<const float4*>(p.head_w + static_cast<size_t>(q) * p.N + ncol) + 2 * (j + 1);
              hw[(j + 1) & 1][q * 2] = ldg_f4_pinned(w4);
              hw[(j + 1) & 1][q * 2 + 1] = ldg_f4_pinned(w4 + 1);
            }
          }
        }
        if constexpr (HEAD > 0) {
#pragma unroll
          for (int q = 0; q < HEAD; ++q) {
            const float4 w0 = HB == 1 ? hw[j & 1][q * 2] : hw[0][q * 2 * HB + 2 * j];
            const float4 w1 = HB == 1 ? hw[j & 1][q * 2 + 1] : hw[0][q * 2 * HB + 2 * j + 1];
            hacc[q] += f[0] * w0.x + f[1] * w0.y + f[2] * w0.z + f[3] * w0.w + f[4] * w1.x + f[5] * w1.y +
                       f[6] * w1.z + f[7] * w1.w;
          }
        }
        if (p.store_c) {
          __half2 h0 = __floats2half2_rn(f[0], f[1]);
          __half2 h1 = __floats2half2_rn(f[2], f[3]);
          __half2 h2 = __floats2half2_rn(f[4], f[5]);
          __half2 h3 = __floats2half2_rn(f[6], f[7]);
          const int chunk = h * 4 + j;              // 16-byte chunk within the 128-byte row
          const uint32_t addr = cbuf + row * 128 + ((chunk ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                       "r"(*reinterpret_cast<uint32_t*>(&h0)), "r"(*reinterpret_cast<uint32_t*>(&h1)),
                       "r"(*reinterpret_cast<uint32_t*>(&h2)), "r"(*reinterpret_cast<uint32_t*>(&h3))
                       : "memory");
        }
      }
      if constexpr (BWD) {
#pragma unroll
        for (int j = 0; j < 4; ++j) mk_cur[j] = mk_nxt[j];
      }
    }
    if (p.store_c) {
      fence_proxy_async_smem();
      named_bar_sync(g.bar_id, 128);
      if (g.gtid == 0) {
        tma_store_2d(tmC, cbuf, n0 + cb * 64, m0);
        tma_store_commit();
      }
      ++cnt;
    }
  }
  if constexpr (HEAD > 0) {
    if (hrow0 + row < p.M) {
      float* dst = p.head_out + static_cast<size_t>(hrow0 + row) * p.head_stride + p.head_slot0 + g.slot * HEAD;
#pragma unroll
      for (int q = 0; q < HEAD; ++q) dst[q] = hacc[q];
    }
  }
}

template <int BN, bool BWD, int NBUF>
__device__ __forceinline__ void epilogue_tile(const EpiParams& p, const void* tmC, uint32_t acc_addr, const EpiGroup& g,
                                              uint32_t& cnt, int m0, int n0, int row, long long hrow0 = -1) {
  if (hrow0 < 0) hrow0 = m0;
  if constexpr (BWD) {
    epilogue_tile_impl<BN, BWD, 0, NBUF>(p, tmC, acc_addr, g, cnt, m0, n0, row, hrow0);
  } else {
    if (p.head_n == 0) epilogue_tile_impl<BN, BWD, 0, NBUF>(p, tmC, acc_addr, g, cnt, m0, n0, row, hrow0);
    else if (p.head_n == 1) epilogue_tile_impl<BN, BWD, 1, NBUF>(p, tmC, acc_addr, g, cnt, m0, n0, row, hrow0);
    else epilogue_tile_impl<BN, BWD, 3, NBUF>(p, tmC, acc_addr, g, cnt, m0, n0, row, hrow0);
  }
}

}  // namespace mofa
