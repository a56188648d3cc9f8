// C-ABI implementation (include/mofa_b200.h): context, weight repack, latent fold, the layer program
// of the two MoFaNeRF MLPs and the render_rays orchestration.
#include "../../include/mofa_b200.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "engine.h"

namespace {

thread_local std::string g_err;

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}

#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);     \
  } while (0)

constexpr int kNShape = 50, kNExp = 30, kNTex = 256, kMultires = 10, kMultiresViews = 4;
constexpr int kPeXyz = 3 + 6 * kMultires;        // 63
constexpr int kPeView = 3 + 6 * kMultiresViews;  // 27
constexpr int kMaxSms = 160;
constexpr int kHeadStride = 32;                  // fp32 partial-head slots per point: alpha at 0..7, rgb (x3) at 8..31
constexpr int kRgbSlot0 = 8;

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

enum Lat { LAT_NONE = 0, LAT_EXP = 1, LAT_SHAPE = 2, LAT_TEX = 3 };
enum Src { SRC_X0 = 0, SRC_V = 1, SRC_T0 = 2 };  // SRC_T0 + i = activation buffer i

struct Layer {
  int N = 0;
  int nseg = 0;
  int K[2] = {0, 0};             // padded K per segment
  __half* w[2] = {nullptr, nullptr};
  CUtensorMap tmB[2];            // box rows = BN (one CTA per tile)
  CUtensorMap tmB2[2];           // box rows = 128 (CTA-pair kernel: each CTA loads half of the 256 rows)
  float* bias_raw = nullptr;     // [N]
  float* bias_eff = nullptr;     // [N] (== bias_raw when nothing is folded)
  float* fold_w = nullptr;       // [N, fold_n] fp32 latent columns
  int fold_n = 0;
  int fold_lat = LAT_NONE;
  int BN = 256;
  int in_ref = 0;                // reference in_features (algorithmic FLOP accounting)
  // backward: transposed segment weights [rows_t, N] (rows_t = K padded to >= 128), the
  // K-major B operand of dX = dZ · W
  int seg_c0[2] = {0, 0};        // first column / real width of each activation segment in the reference weight
  int seg_kreal[2] = {0, 0};
  int fold_c0 = 0;               // first latent column
  int in_total = 0;              // row pitch of the reference weight
  __half* wt[2] = {nullptr, nullptr};
  int rows_t[2] = {0, 0};
  int BN_t[2] = {256, 256};
  CUtensorMap tmBt[2];           // box rows = BN_t
  CUtensorMap tmBt2[2];          // box rows = 128 (CTA-pair kernel)
  // split-precision coarse kernel (W == 256 nets only): low image fp16(w - fp16(w)) and half-N boxes for the CTA pair
  __half* wlo[2] = {nullptr, nullptr};
  CUtensorMap tmS_hi[2];         // box rows = N / 2
  CUtensorMap tmS_lo[2];
  // opt-in FP8 variant (MOFA_B200_FP8; plain W -> W layers of a wide net only): e4m3 image, per-row scale / activation scale
  uint8_t* w8 = nullptr;
  float* colscale8 = nullptr;
  CUtensorMap tmB8;              // box 128 rows x 128 bytes
};

struct Step {
  int kind;       // 0 dense, 1 alpha head, 2 rgb head
  int layer;      // index into Net::layers (dense)
  int in[2];      // Src ids
  int out;        // Src id (dense)
  int head;       // dense step whose output feeds a head: 1 alpha_linear, 2 rgb_linear (fused in the epilogue)
  int in_step[2]; // producer of each input: dense ordinal >= 0, -1 = X0 (point encoding), -2 = V (view encoding), -3 none
  int ord;        // ordinal among the dense steps
};

struct Net {
  bool loaded = false;
  int W = 0, D = 0;
  std::vector<Layer> layers;
  std::vector<Step> program;
  float *w_alpha = nullptr, *b_alpha = nullptr, *w_rgb = nullptr, *b_rgb = nullptr;
  // fused persistent kernel (W == 256 only): device-resident layer table and weight tensor maps
  mofa::FusedLayerDesc* fused_layers = nullptr;
  CUtensorMap* fused_wmaps = nullptr;
  int fused_n = 0;
  // split-precision fused kernel (W == 256 only): layer table, weight maps, fp32 view-direction columns [W/2, 27]
  mofa::SplitLayerDesc* split_layers = nullptr;
  CUtensorMap* split_wmaps = nullptr;
  int split_n = 0;
  float* w_view = nullptr;
  std::vector<void*> allocs;
  // the allocations that hold WEIGHT DATA (packed fp16 images, biases, latent columns, heads), in creation order —
  // what mofa_b200_export_packed writes and mofa_b200_import_packed restores; derived tables are rebuilt on import
  std::vector<std::pair<void*, size_t>> data_allocs;
};

}  // namespace

struct mofa_b200_ctx {
  int device = 0;
  int num_sms = 148;
  EncodeTiledFn encode = nullptr;
  Net nets[2];
  float* lat[4] = {nullptr, nullptr, nullptr, nullptr};  // device copies of the current latents
  bool latents_set = false;
  bool pair_kernel = true;       // cta_group::2 kernel for N % 256 == 0 (MOFA_B200_DENSE_1CTA=1 disables)
  bool fused_coarse = true;      // one persistent kernel for a W == 256 net (MOFA_B200_NO_FUSED_COARSE=1 disables)
  int fp8_layers = 0;            // MOFA_B200_FP8=1|all: every plain W -> W layer of a wide net runs with e4m3 operands;
                                 // MOFA_B200_FP8=layers=K: the first K of them.  NOT the default: fails the stated tolerance (profiles/r02_fp8_parity_study.json)
  bool importing = false;        // load_weights is being driven by mofa_b200_import_packed: build the structure, read no sources
  bool chain_fine = true;        // all dense layers of a W >= 512 net in one persistent launch, activations L2-resident
                                 // (fine_chain.cu); MOFA_B200_FINE_PER_LAYER=1 selects one launch per layer (round 1)
  // device tables of the chain kernel (tensor maps + layer descriptors), rebuilt when the buffers they point to change
  CUtensorMap* chain_maps = nullptr;
  mofa::ChainLayerDesc* chain_layers = nullptr;
  const void* chain_key[4] = {nullptr, nullptr, nullptr, nullptr};
  mofa::ChainParams chain_proto;
  bool split_coarse = true;      // ... in split precision (fp16 hi+lo, 3 products): MOFA_B200_COARSE_FP16=1 selects the
                                 // single-fp16 fused kernel of round 1 instead
  int64_t launches = 0;
  // profiling (bench): CUDA-event pairs around every tensor-core dense launch, on the launch stream
  bool profiling = false;
  std::vector<cudaEvent_t> ev;        // pool: 2 per record
  struct Rec { int net; double flops; };
  std::vector<Rec> recs;
};

namespace {

using namespace mofa;

// box_cols = 64 (128-byte rows, SWIZZLE_128B: every operand load and the per-layer kernels' stores) or 32 (64-byte rows,
// SWIZZLE_64B: the chain kernel's half-block stores)
int make_tmap_2d(mofa_b200_ctx* c, CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch,
                 uint32_t box_rows, uint32_t box_cols = 64) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch * sizeof(__half)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = c->encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu pitch=%llu box_rows=%u ptr=%p", (int)r,
                (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch, box_rows, ptr);
  return 0;
}

// 8-bit (e4m3) operand / output maps of the FP8 variant: bytes as elements.  box_cols = 128 (operands, SWIZZLE_128B) or 32
// (the chain kernel's half-block stores, no swizzle)
int make_tmap_u8(mofa_b200_ctx* c, CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch,
                 uint32_t box_rows, uint32_t box_cols) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = c->encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         box_cols == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (u8) failed (%d)", (int)r);
  return 0;
}

int dev_alloc(Net& n, void** p, size_t bytes) {
  CK(cudaMalloc(p, bytes));
  n.allocs.push_back(*p);
  return 0;
}
int dev_alloc_data(Net& n, void** p, size_t bytes) {
  if (dev_alloc(n, p, bytes)) return 1;
  n.data_allocs.emplace_back(*p, bytes);
  return 0;
}

void free_net(Net& n) {
  for (void* p : n.allocs) cudaFree(p);
  n = Net();
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// One linear layer of the reference -> engine layer.  Column layout of the reference weight
// [N, n_lat + K0 (+ K1)] : latent (or PE-extra) columns first except for xyzEncode.Linear0 where the
// modulated expression code comes *after* the 63 PE columns (render_class.py:83).
struct LayerSpec {
  int N;
  int in_total;       // reference in_features
  int nseg;
  int seg_c0[2];      // first column of each activation segment in the reference weight
  int seg_k[2];       // real width
  int seg_kpad[2];    // padded to 64
  int fold_c0, fold_n, fold_lat;
};

int build_layer(mofa_b200_ctx* c, Net& net, const LayerSpec& sp, const float* w, const float* b, cudaStream_t s) {
  const bool imp = c->importing;
  Layer L;
  L.N = sp.N;
  L.nseg = sp.nseg;
  L.in_ref = sp.in_total;
  L.in_total = sp.in_total;
  L.BN = (sp.N % 256 == 0) ? 256 : 128;
  if (sp.N % L.BN != 0) return fail("layer width %d is not a multiple of 128", sp.N);
  for (int i = 0; i < sp.nseg; ++i) {
    L.K[i] = sp.seg_kpad[i];
    L.seg_c0[i] = sp.seg_c0[i];
    L.seg_kreal[i] = sp.seg_k[i];
    if (dev_alloc_data(net, reinterpret_cast<void**>(&L.w[i]), sizeof(__half) * (size_t)sp.N * L.K[i])) return 1;
    if (!imp) {
      CK(launch_pack_weight(w, sp.in_total, sp.seg_c0[i], sp.seg_k[i], L.K[i], sp.N, L.w[i], s));
      c->launches++;
    }
    if (make_tmap_2d(c, &L.tmB[i], L.w[i], sp.N, L.K[i], L.K[i], L.BN)) return 1;
    if (make_tmap_2d(c, &L.tmB2[i], L.w[i], sp.N, L.K[i], L.K[i], 128)) return 1;
    if (net.W == 256) {
      if (dev_alloc_data(net, reinterpret_cast<void**>(&L.wlo[i]), sizeof(__half) * (size_t)sp.N * L.K[i])) return 1;
      if (!imp) {
        CK(launch_pack_weight_lo(w, sp.in_total, sp.seg_c0[i], sp.seg_k[i], L.K[i], sp.N, L.wlo[i], s));
        c->launches++;
      }
      if (make_tmap_2d(c, &L.tmS_hi[i], L.w[i], sp.N, L.K[i], L.K[i], sp.N / 2)) return 1;
      if (make_tmap_2d(c, &L.tmS_lo[i], L.wlo[i], sp.N, L.K[i], L.K[i], sp.N / 2)) return 1;
    }
    L.rows_t[i] = L.K[i] < 128 ? 128 : L.K[i];
    L.BN_t[i] = (L.rows_t[i] % 256 == 0) ? 256 : 128;
    if (dev_alloc_data(net, reinterpret_cast<void**>(&L.wt[i]), sizeof(__half) * (size_t)L.rows_t[i] * sp.N)) return 1;
    if (!imp) {
      CK(launch_pack_weight_t(w, sp.in_total, sp.seg_c0[i], sp.seg_k[i], L.rows_t[i], sp.N, L.wt[i], s));
      c->launches++;
    }
    if (make_tmap_2d(c, &L.tmBt[i], L.wt[i], L.rows_t[i], sp.N, sp.N, L.BN_t[i])) return 1;
    if (make_tmap_2d(c, &L.tmBt2[i], L.wt[i], L.rows_t[i], sp.N, sp.N, 128)) return 1;
  }
  if (c->fp8_layers > 0 && !imp && net.W >= 512 && sp.nseg == 1 && sp.fold_n == 0 && sp.seg_k[0] == net.W && sp.N == net.W) {
    // (not part of the packed blob: the FP8 variant is a measurement mode)
    if (dev_alloc(net, reinterpret_cast<void**>(&L.w8), (size_t)sp.N * sp.seg_k[0])) return 1;
    if (dev_alloc(net, reinterpret_cast<void**>(&L.colscale8), sizeof(float) * sp.N)) return 1;
    CK(launch_pack_weight_fp8(w, sp.in_total, sp.seg_c0[0], sp.seg_k[0], sp.N, 8.0f, L.w8, L.colscale8, s));
    c->launches++;
    if (make_tmap_u8(c, &L.tmB8, L.w8, sp.N, sp.seg_k[0], sp.seg_k[0], 128, 128)) return 1;
  }
  if (dev_alloc_data(net, reinterpret_cast<void**>(&L.bias_raw), sizeof(float) * sp.N)) return 1;
  if (!imp) CK(cudaMemcpyAsync(L.bias_raw, b, sizeof(float) * sp.N, cudaMemcpyDeviceToDevice, s));
  L.bias_eff = L.bias_raw;
  if (sp.fold_n > 0) {
    L.fold_n = sp.fold_n;
    L.fold_lat = sp.fold_lat;
    L.fold_c0 = sp.fold_c0;
    if (dev_alloc_data(net, reinterpret_cast<void**>(&L.fold_w), sizeof(float) * (size_t)sp.N * sp.fold_n)) return 1;
    if (!imp)
      CK(cudaMemcpy2DAsync(L.fold_w, sizeof(float) * sp.fold_n, w + sp.fold_c0, sizeof(float) * sp.in_total,
                           sizeof(float) * sp.fold_n, sp.N, cudaMemcpyDeviceToDevice, s));
    if (dev_alloc(net, reinterpret_cast<void**>(&L.bias_eff), sizeof(float) * sp.N)) return 1;
  }
  net.layers.push_back(L);
  return 0;
}

int pad64(int k) { return (k + 63) / 64 * 64; }

// Emits dense steps for a run of layers and keeps the 3-buffer allocation invariant:
// at most {pinned, cur} are live, so one activation buffer is always free.
struct ProgBuilder {
  Net& net;
  int cur = SRC_X0;
  int pinned = -1;
  int writer[SRC_T0 + 3] = {-1, -2, -3, -3, -3};   // last dense ordinal that wrote each buffer (X0 = -1, V = -2)
  int n_dense = 0;
  explicit ProgBuilder(Net& n) : net(n) {}
  int free_buf(int a, int b) const {
    for (int t = SRC_T0; t < SRC_T0 + 3; ++t)
      if (t != a && t != b && t != pinned && t != cur) return t;
    return -1;
  }
  void dense(int layer, int in0, int in1) {
    Step st;
    st.kind = 0;
    st.layer = layer;
    st.in[0] = in0;
    st.in[1] = in1;
    st.out = free_buf(in0, in1);
    st.head = 0;
    st.in_step[0] = writer[in0];
    st.in_step[1] = in1 >= 0 ? writer[in1] : -3;
    st.ord = n_dense++;
    writer[st.out] = st.ord;
    net.program.push_back(st);
    cur = st.out;
  }
};

int fold_net(mofa_b200_ctx* c, Net& net, cudaStream_t s) {
  for (Layer& L : net.layers) {
    if (L.fold_n == 0) continue;
    CK(launch_fold_bias(L.fold_w, L.fold_n, 0, L.fold_n, L.bias_raw, c->lat[L.fold_lat], L.N, L.bias_eff, s));
    c->launches++;
  }
  return 0;
}


// Layer table of the split-precision fused kernel (coarse_split.cu).  Every layer reads the activation resident in shared
// memory; a skip layer's second operand becomes a "virtual" layer right after the producer of that operand (raw fp32
// partial product parked in a scratch, added in the skip layer's epilogue), the view layer's view-direction segment
// becomes a per-ray vector.  Leaves split_n == 0 (the caller falls back to the single-fp16 fused kernel) when the
// program does not have the shape the kernel's hazard analysis assumes.
int build_split_table(mofa_b200_ctx* c, Net& net, cudaStream_t s) {
  std::vector<const Step*> dense;
  for (const Step& st : net.program)
    if (st.kind == 0) dense.push_back(&st);
  struct Pending { int producer, layer, seg; };
  std::vector<Pending> pend;
  for (const Step* st : dense) {
    const Layer& L = net.layers[st->layer];
    if (L.nseg == 2 && st->in_step[0] != -2 && st->in_step[1] != -2) {
      const int prim = st->in_step[0] == st->ord - 1 ? 0 : 1;
      if (st->in_step[prim] != st->ord - 1) return 0;
      pend.push_back({st->in_step[1 - prim], st->layer, 1 - prim});
    }
  }
  std::vector<SplitLayerDesc> sl;
  std::vector<CUtensorMap> maps;
  bool fresh = false, parked = false, ok = true;
  auto push_maps = [&](const Layer& L, int seg, SplitLayerDesc& d) {
    d.map_hi = static_cast<int>(maps.size());
    maps.push_back(L.tmS_hi[seg]);
    d.map_lo = static_cast<int>(maps.size());
    maps.push_back(L.tmS_lo[seg]);
    d.kb = L.K[seg] / 64;
  };
  for (const Step* st : dense) {
    const Layer& L = net.layers[st->layer];
    SplitLayerDesc d{};
    d.bias = L.bias_eff;
    d.n_out = L.N;
    d.head = st->head;
    d.store = st->head == 2 ? 0 : 1;
    int prim = 0;
    if (L.nseg == 2) {
      if (st->in_step[0] == -2 || st->in_step[1] == -2) {
        prim = st->in_step[0] == -2 ? 1 : 0;
        d.add_ray = 1;
        if (L.N != 128 || st->head != 2) ok = false;     // the per-ray vector has 128 columns: the view layer
      } else {
        prim = st->in_step[0] == st->ord - 1 ? 0 : 1;
        d.add_park = 1;
        if (!parked) ok = false;
        parked = false;
      }
    }
    push_maps(L, prim, d);
    if (sl.empty()) {
      if (st->in_step[prim] != -1 || d.kb != 1) ok = false;    // first layer reads the point encoding (one K block)
    } else {
      if (st->in_step[prim] != st->ord - 1 || d.kb != 4) ok = false;
    }
    if (d.n_out != 256 && !(d.n_out == 128 && !d.store)) ok = false;
    d.wait_act = fresh ? 1 : 0;
    fresh = d.store != 0;
    if (!d.store && st != dense.back()) ok = false;             // only the last layer may skip the store
    sl.push_back(d);
    for (const Pending& p : pend) {
      if (p.producer != st->ord) continue;
      const Layer& PL = net.layers[p.layer];
      SplitLayerDesc v{};
      v.n_out = PL.N;
      v.park = 1;
      push_maps(PL, p.seg, v);
      v.wait_act = fresh ? 1 : 0;
      if (!fresh || parked || v.kb != 4 || v.n_out != 256) ok = false;   // directly after a storing layer, one park at a time
      fresh = false;
      parked = true;
      sl.push_back(v);
    }
  }
  if (!ok || sl.empty() || net.w_view == nullptr) return 0;
  if (dev_alloc(net, reinterpret_cast<void**>(&net.split_layers), sizeof(SplitLayerDesc) * sl.size())) return 1;
  if (dev_alloc(net, reinterpret_cast<void**>(&net.split_wmaps), sizeof(CUtensorMap) * maps.size())) return 1;
  CK(cudaMemcpyAsync(net.split_layers, sl.data(), sizeof(SplitLayerDesc) * sl.size(), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(net.split_wmaps, maps.data(), sizeof(CUtensorMap) * maps.size(), cudaMemcpyHostToDevice, s));
  CK(cudaStreamSynchronize(s));   // the host vectors go out of scope
  net.split_n = static_cast<int>(sl.size());
  return 0;
}

struct Workspace {
  float *z_c, *w_c, *z_f, *raw, *hp;
  __half *X0, *V, *T[3];
  __half* fscratch;    // fused coarse kernel: [num_sms * 128, 256] parked skip tensors
  __half* X0lo;        // split-precision coarse kernel: low image of the point encoding [P_pad, 64]
  float* ray_vec;      // ... per-ray view vector [groups, 128]
  float* park;         // ... parked fp32 skip partial products, one [128 x 256] block per CTA
  uint32_t* chain_ctr; // fine-net chain kernel: per (layer, m-block) completion counters
  int64_t t_rows;      // rows of each activation buffer T[i] (== P_pad, or a few slabs when the chain kernel is in use)
  int64_t P_pad;
  size_t total;
};

// Where the view directions of the point rows come from: one direction per `rows_per_group` consecutive rows.
struct ViewSrc {
  const float* dirs;
  int stride;
  int rows_per_group;
};

constexpr int kChainMaxLayers = 32;     // the chain kernel keeps its layer table in shared memory (48 bytes per layer)

// m-blocks per slab of the fine-net chain kernel (default kChainSlabMb = 56: three rounds of the 74 pairs per layer,
// 3 x 29 MB of activations); MOFA_B200_CHAIN_SLAB overrides it for measurements
int chain_slab_mb() {
  static const int v = [] {
    const char* e = getenv("MOFA_B200_CHAIN_SLAB");
    const int n = e ? atoi(e) : 0;
    return (n >= 8 && n <= 1024) ? n : kChainSlabMb;
  }();
  return v;
}
constexpr int64_t kSmallRows = 65536;     // activation-buffer rows kept for the per-layer / SIMT paths when the chain kernel is on

// Rows of the three rotating activation buffers.  With the chain kernel the fine net only ever touches one slab of them
// (that is the point: they stay in L2), so a chunk may be arbitrarily long without growing them; the per-layer paths
// (SIMT verification, MOFA_B200_FINE_PER_LAYER / NO_FUSED_COARSE) need one row per point.
int64_t t_rows_for(const mofa_b200_ctx* c, int64_t P_pad) {
  if (c == nullptr) return P_pad;     // training workspaces: activations are kept per layer elsewhere
  bool chain = c->chain_fine && c->pair_kernel && c->fused_coarse;
  bool any = false;
  for (int i = 0; i < 2; ++i)
    if (c->nets[i].loaded && c->nets[i].W >= 512) {
      any = true;
      if (c->nets[i].W % 512 != 0 || c->nets[i].W > 1024) chain = false;
    }
  if (!chain || !any) return P_pad;
  const int64_t slab = static_cast<int64_t>(chain_slab_mb()) * 256;
  const int64_t small = P_pad < kSmallRows ? P_pad : kSmallRows;
  return small > slab ? small : slab;
}

Workspace carve(const mofa_b200_ctx* c, void* base, int64_t n_chunk, int S_c, int S_f, int Wmax) {
  Workspace w;
  const int S_max = S_f > S_c ? S_f : S_c;
  w.P_pad = (n_chunk * S_max + 127) / 128 * 128;
  w.t_rows = t_rows_for(c, w.P_pad);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  uint8_t* b = static_cast<uint8_t*>(base);
  w.z_c = reinterpret_cast<float*>(b + take(sizeof(float) * n_chunk * S_c));
  w.w_c = reinterpret_cast<float*>(b + take(sizeof(float) * n_chunk * S_c));
  w.z_f = reinterpret_cast<float*>(b + take(sizeof(float) * n_chunk * (S_f > 0 ? S_f : 1)));
  w.raw = reinterpret_cast<float*>(b + take(sizeof(float) * 4 * w.P_pad));
  w.hp = reinterpret_cast<float*>(b + take(sizeof(float) * kHeadStride * w.P_pad));
  w.X0 = reinterpret_cast<__half*>(b + take(sizeof(__half) * 64 * w.P_pad));
  w.V = reinterpret_cast<__half*>(b + take(sizeof(__half) * 64 * w.P_pad));
  for (int i = 0; i < 3; ++i) w.T[i] = reinterpret_cast<__half*>(b + take(sizeof(__half) * (size_t)Wmax * w.t_rows));
  w.fscratch = reinterpret_cast<__half*>(b + take(sizeof(__half) * 256 * 128 * kMaxSms));
  w.X0lo = reinterpret_cast<__half*>(b + take(sizeof(__half) * 64 * w.P_pad));
  w.ray_vec = reinterpret_cast<float*>(b + take(sizeof(float) * 128 * n_chunk));
  w.park = reinterpret_cast<float*>(b + take(coarse_split_park_bytes(kMaxSms)));
  w.chain_ctr = reinterpret_cast<uint32_t*>(b + take(sizeof(uint32_t) * kChainMaxLayers * (w.P_pad / 256 + 1)));
  w.total = off;
  return w;
}

int max_width(mofa_b200_ctx* c) {
  int w = 0;
  for (int i = 0; i < 2; ++i)
    if (c->nets[i].loaded && c->nets[i].W > w) w = c->nets[i].W;
  return w;
}

// Runs the MLP program of `net` over the first P_pad rows of the workspace buffers.

// All dense layers of a wide net (W = 512 or 1024) in ONE persistent launch with L2-resident activations (fine_chain.cu).
// The tensor-map / layer tables live in device memory and are rebuilt only when the buffers they describe change.
bool chain_applies(const mofa_b200_ctx* c, const Net& net) {
  return c->chain_fine && c->pair_kernel && net.W >= 512 && net.W % 512 == 0 && net.W <= 1024 &&
         static_cast<int>(net.layers.size()) <= kChainMaxLayers;
}

int run_chain(mofa_b200_ctx* c, Net& net, const Workspace& ws, int64_t P_rows, cudaStream_t s) {
  const int net_id = static_cast<int>(&net - c->nets);
  const int64_t slab_rows = static_cast<int64_t>(chain_slab_mb()) * 256;
  if (ws.t_rows < slab_rows) return fail("chain kernel: activation buffers have %lld rows, need %lld", (long long)ws.t_rows, (long long)slab_rows);
  const void* key[4] = {ws.T[0], ws.X0, ws.V, &net};
  if (c->chain_maps == nullptr) {
    CK(cudaMalloc(reinterpret_cast<void**>(&c->chain_maps), sizeof(CUtensorMap) * kChainMaxLayers * 5));
    CK(cudaMalloc(reinterpret_cast<void**>(&c->chain_layers), sizeof(ChainLayerDesc) * kChainMaxLayers));
  }
  if (memcmp(key, c->chain_key, sizeof(key)) != 0) {
    std::vector<CUtensorMap> maps;
    std::vector<ChainLayerDesc> descs;
    auto src_map = [&](int id, int K, int* is_global) -> int {
      CUtensorMap m;
      int rc;
      if (id == SRC_X0 || id == SRC_V) {
        *is_global = 1;
        rc = make_tmap_2d(c, &m, id == SRC_X0 ? ws.X0 : ws.V, (uint64_t)ws.P_pad, 64, 64, 128);
      } else {
        *is_global = 0;
        rc = make_tmap_2d(c, &m, ws.T[id - SRC_T0], (uint64_t)slab_rows, (uint64_t)K, (uint64_t)K, 128);
      }
      if (rc) return -1;
      maps.push_back(m);
      return static_cast<int>(maps.size()) - 1;
    };
    const int pair_groups = 2;
    int tiles_per_mb = 0, nt = 0, nt_last = 0;
    // FP8 variant: which dense steps take e4m3 operands (the first fp8_layers plain layers), and which therefore must
    // PRODUCE e4m3 (every consumer of a tensor has the same operand type by construction of the network)
    std::vector<int> f8_in, f8_out;
    {
      std::vector<const Step*> dn;
      for (const Step& st : net.program)
        if (st.kind == 0) dn.push_back(&st);
      f8_in.assign(dn.size(), 0);
      f8_out.assign(dn.size(), 0);
      int plain = 0;
      for (size_t i = 0; i < dn.size(); ++i)
        if (net.layers[dn[i]->layer].w8 != nullptr && plain++ < c->fp8_layers) f8_in[i] = 1;
      for (size_t i = 0; i < dn.size(); ++i) {
        int n8 = 0, n16 = 0;
        for (size_t j = 0; j < dn.size(); ++j)
          for (int k = 0; k < net.layers[dn[j]->layer].nseg; ++k)
            if (dn[j]->in_step[k] == dn[i]->ord) (f8_in[j] ? n8 : n16)++;
        if (n8 > 0 && n16 > 0) return fail("chain kernel (FP8): a tensor has consumers of both operand types");
        f8_out[i] = n8 > 0 ? 1 : 0;
      }
    }
    int dense_i = -1;
    for (const Step& st : net.program) {
      if (st.kind != 0) continue;
      ++dense_i;
      const Layer& L = net.layers[st.layer];
      ChainLayerDesc d;
      memset(&d, 0, sizeof(d));
      d.fp8_in = f8_in[dense_i];
      d.fp8_out = f8_out[dense_i];
      d.colscale = d.fp8_in ? L.colscale8 : nullptr;
      d.kb0 = L.K[0] / (d.fp8_in ? 128 : 64);
      d.kb1 = L.nseg > 1 ? L.K[1] / 64 : 0;
      d.n_tiles = L.N / 256;
      d.N = L.N;
      d.bias = L.bias_eff;
      d.relu = 1;
      d.store_c = 1;
      if (d.fp8_in) {        // e4m3 view of the same activation buffer: [slab_rows, K] bytes
        CUtensorMap m;
        if (make_tmap_u8(c, &m, ws.T[st.in[0] - SRC_T0], (uint64_t)slab_rows, (uint64_t)L.K[0], (uint64_t)L.K[0], 128, 128)) return 1;
        maps.push_back(m);
        d.mapA0 = static_cast<int>(maps.size()) - 1;
        d.a0_global = 0;
        maps.push_back(L.tmB8);
        d.mapB0 = static_cast<int>(maps.size()) - 1;
      } else {
        if ((d.mapA0 = src_map(st.in[0], L.K[0], &d.a0_global)) < 0) return 1;
        maps.push_back(L.tmB2[0]);
        d.mapB0 = static_cast<int>(maps.size()) - 1;
      }
      d.mapA1 = d.mapA0;
      d.mapB1 = d.mapB0;
      if (L.nseg > 1) {
        if ((d.mapA1 = src_map(st.in[1], L.K[1], &d.a1_global)) < 0) return 1;
        maps.push_back(L.tmB2[1]);
        d.mapB1 = static_cast<int>(maps.size()) - 1;
      }
      if (st.head == 1) {
        d.head_w = net.w_alpha; d.head_n = 1; d.head_slot0 = 0;
      } else if (st.head == 2) {
        d.head_w = net.w_rgb; d.head_n = 3; d.head_slot0 = kRgbSlot0;
        d.store_c = 0;                       // the view layer feeds rgb_linear only
      }
      {
        CUtensorMap m;
        // output map: 32-column boxes (the epilogue stages and stores half a 64-column block at a time, double-buffered)
        if (d.fp8_out) {
          if (make_tmap_u8(c, &m, ws.T[st.out - SRC_T0], (uint64_t)slab_rows, (uint64_t)L.N, (uint64_t)L.N, 128, 32)) return 1;
        } else if (make_tmap_2d(c, &m, ws.T[st.out - SRC_T0], (uint64_t)slab_rows, (uint64_t)L.N, (uint64_t)L.N, 128, 32)) return 1;
        maps.push_back(m);
        d.mapC = static_cast<int>(maps.size()) - 1;
      }
      tiles_per_mb += d.n_tiles;
      descs.push_back(d);
    }
    const int nl = static_cast<int>(descs.size());
    if (nl < 2 || nl > kChainMaxLayers || static_cast<int>(maps.size()) > kChainMaxLayers * 5) return fail("chain kernel: unsupported program");
    nt = descs[0].n_tiles;
    nt_last = descs[nl - 1].n_tiles;
    for (int i = 0; i < nl - 1; ++i)
      if (descs[i].n_tiles != nt) return fail("chain kernel: layers of different widths");
    if (nt * pair_groups > 8 || nt_last * pair_groups * 3 > kHeadStride - kRgbSlot0) return fail("chain kernel: head slots do not fit");
    CK(cudaMemcpyAsync(c->chain_maps, maps.data(), sizeof(CUtensorMap) * maps.size(), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->chain_layers, descs.data(), sizeof(ChainLayerDesc) * nl, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));            // the host vectors go out of scope; happens once per (workspace, net)
    memcpy(c->chain_key, key, sizeof(key));
    ChainParams& P = c->chain_proto;
    memset(&P, 0, sizeof(P));
    P.maps = c->chain_maps;
    P.layers = c->chain_layers;
    P.n_layers = nl;
    P.nt = nt;
    P.nt_last = nt_last;
    P.tiles_per_mb = tiles_per_mb;
    P.slab_mb = chain_slab_mb();
    P.head_stride = kHeadStride;
    {
      const char* v = getenv("MOFA_B200_CHAIN_NODEP");
      P.nodep = (v && v[0] == '1') ? 1 : 0;
    }
  }
  ChainParams P = c->chain_proto;
  P.total_mb = static_cast<int>((P_rows + 255) / 256);
  P.head_out = ws.hp;
  P.P_rows = P_rows;
  P.counters = ws.chain_ctr;
  CK(cudaMemsetAsync(ws.chain_ctr, 0, sizeof(uint32_t) * (size_t)P.n_layers * P.total_mb, s));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (c->profiling) {
    const size_t idx = c->recs.size() * 2;
    while (c->ev.size() < idx + 2) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      c->ev.push_back(e);
    }
    e0 = c->ev[idx];
    e1 = c->ev[idx + 1];
    CK(cudaEventRecord(e0, s));
  }
  CK(launch_fine_chain(P, c->num_sms, s));
  if (c->profiling) {
    double fl = 0.0;
    for (const Step& st : net.program)
      if (st.kind == 0) fl += 2.0 * (double)P_rows * (double)net.layers[st.layer].N * (double)net.layers[st.layer].in_ref;
    CK(cudaEventRecord(e1, s));
    c->recs.push_back({net_id, fl});
  }
  c->launches++;
  const int a_slots = (net.W / 256) * 2, r_slots = (net.W / 2 / 256) * 2;
  CK(launch_finalize_raw(ws.hp, kHeadStride, 0, a_slots, kRgbSlot0, r_slots, net.b_alpha, net.b_rgb, ws.raw, P_rows, s));
  c->launches++;
  return 0;
}

// True when `net` runs in the split-precision fused kernel for an inference pass (the caller must then have produced the
// low image of the point encoding, ws.X0lo).
bool use_split(const mofa_b200_ctx* c, const Net& net, uint32_t flags) {
  return !(flags & MOFA_FLAG_GEMM_SIMT) && c->fused_coarse && c->split_coarse && net.split_n > 0 && c->num_sms <= kMaxSms;
}

int run_program(mofa_b200_ctx* c, Net& net, const Workspace& ws, int64_t P_rows, uint32_t flags, cudaStream_t s,
                __half* const* act = nullptr, const ViewSrc* view = nullptr) {
  // act != nullptr: training mode — dense step k writes act[k] (kept for the backward pass) instead of the ping-pong buffers
  const int net_id = static_cast<int>(&net - c->nets);
  (void)net_id;
  const int64_t M = (P_rows + 127) / 128 * 128;
  auto src_ptr = [&](int id) -> __half* { return id == SRC_X0 ? ws.X0 : id == SRC_V ? ws.V : ws.T[id - SRC_T0]; };
  const bool tc = !(flags & MOFA_FLAG_GEMM_SIMT);
  if (!act && view != nullptr && use_split(c, net, flags)) {
    // whole network in one persistent CTA-pair kernel, fp16 hi+lo operands (fp32-class results)
    const int64_t groups = (P_rows + view->rows_per_group - 1) / view->rows_per_group;
    CK(launch_view_vec(view->dirs, view->stride, groups, net.w_view, net.W / 2, ws.ray_vec, s));
    c->launches++;
    SplitLaunch S;
    memset(&S, 0, sizeof(S));
    if (make_tmap_2d(c, &S.tmX0hi, ws.X0, (uint64_t)M, 64, 64, 128)) return 1;
    if (make_tmap_2d(c, &S.tmX0lo, ws.X0lo, (uint64_t)M, 64, 64, 128)) return 1;
    S.wmaps = net.split_wmaps;
    S.layers = net.split_layers;
    S.n_layers = net.split_n;
    S.P_rows = P_rows;
    S.w_alpha = net.w_alpha; S.b_alpha = net.b_alpha; S.w_rgb = net.w_rgb; S.b_rgb = net.b_rgb;
    S.ray_vec = ws.ray_vec;
    S.rows_per_group = view->rows_per_group;
    S.park = ws.park;
    S.raw = ws.raw;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->profiling) {
      const size_t idx = c->recs.size() * 2;
      while (c->ev.size() < idx + 2) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        c->ev.push_back(e);
      }
      e0 = c->ev[idx];
      e1 = c->ev[idx + 1];
      CK(cudaEventRecord(e0, s));
    }
    CK(launch_coarse_split(S, c->num_sms, s));
    if (c->profiling) {
      double fl = 0.0;
      for (const Step& st : net.program)
        if (st.kind == 0) fl += 2.0 * (double)P_rows * (double)net.layers[st.layer].N * (double)net.layers[st.layer].in_ref;
      CK(cudaEventRecord(e1, s));
      c->recs.push_back({net_id, fl});
    }
    c->launches++;
    return 0;
  }
  if (tc && !act && c->fused_coarse && net.fused_n > 0 && c->num_sms <= kMaxSms) {
    // whole network in one persistent kernel (activations stay in shared memory)
    FusedLaunch F;
    memset(&F, 0, sizeof(F));
    if (make_tmap_2d(c, &F.tmX0, ws.X0, (uint64_t)M, 64, 64, 128)) return 1;
    if (make_tmap_2d(c, &F.tmV, ws.V, (uint64_t)M, 64, 64, 128)) return 1;
    if (make_tmap_2d(c, &F.tmScratch, ws.fscratch, (uint64_t)128 * kMaxSms, 256, 256, 128)) return 1;
    F.wmaps = net.fused_wmaps;
    F.layers = net.fused_layers;
    F.n_layers = net.fused_n;
    F.P_rows = P_rows;
    F.w_alpha = net.w_alpha; F.b_alpha = net.b_alpha; F.w_rgb = net.w_rgb; F.b_rgb = net.b_rgb;
    F.raw = ws.raw;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->profiling) {
      const size_t idx = c->recs.size() * 2;
      while (c->ev.size() < idx + 2) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        c->ev.push_back(e);
      }
      e0 = c->ev[idx];
      e1 = c->ev[idx + 1];
      CK(cudaEventRecord(e0, s));
    }
    CK(launch_coarse_fused(F, c->num_sms, s));
    if (c->profiling) {
      double fl = 0.0;
      for (const Step& st : net.program)
        if (st.kind == 0) fl += 2.0 * (double)P_rows * (double)net.layers[st.layer].N * (double)net.layers[st.layer].in_ref;
      CK(cudaEventRecord(e1, s));
      c->recs.push_back({net_id, fl});
    }
    c->launches++;
    return 0;
  }
  if (tc && !act && chain_applies(c, net)) return run_chain(c, net, ws, P_rows, s);
  if (M > ws.t_rows && !act)
    return fail("the per-layer path needs %lld activation rows but the workspace keeps %lld (chain-kernel layout): render "
                "fewer rays per call or set chunk_rays", (long long)M, (long long)ws.t_rows);
  // head fusion needs every partial slot to fit: W/256 alpha tiles (<= 4), (W/2)/BN rgb tiles (<= 4)
  const int a_tiles = net.W / 256;
  const int r_bn = ((net.W / 2) % 256 == 0) ? 256 : 128;
  const int r_tiles = (net.W / 2) / r_bn;
  const bool fuse_heads = tc && a_tiles <= 4 && r_tiles <= 4;
  // a layer run by the pair kernel writes one partial slot per (n-tile, epilogue group)
  const int pair_groups = c->pair_kernel ? dense_tc2_head_groups() : 1;
  const int a_slots = a_tiles * ((net.W >= 512) ? pair_groups : 1);
  const int r_slots = r_tiles * ((r_bn == 256 && net.W / 2 >= 512) ? pair_groups : 1);
  for (const Step& st : net.program) {
    if (st.kind == 0) {
      const Layer& L = net.layers[st.layer];
      DenseLaunch d;
      memset(&d, 0, sizeof(d));
      for (int i = 0; i < L.nseg; ++i) {
        d.A[i] = (act && st.in_step[i] >= 0) ? act[st.in_step[i]] : src_ptr(st.in[i]);
        d.B[i] = L.w[i];
        d.K[i] = L.K[i];
        d.lda[i] = L.K[i];   // every source buffer is dense with pitch == its K
        d.tmB[i] = L.tmB[i];
        d.tmB2[i] = L.tmB2[i];
      }
      d.C = act ? act[st.ord] : src_ptr(st.out);
      d.ldc = L.N;
      d.bias = L.bias_eff;
      d.M = M;
      d.N = L.N;
      d.BN = L.BN;
      d.relu = 1;
      d.store_c = 1;
      if (fuse_heads && st.head == 1) {
        d.head_w = net.w_alpha; d.head_out = ws.hp; d.head_n = 1; d.head_stride = kHeadStride; d.head_slot0 = 0;
      } else if (fuse_heads && st.head == 2) {
        d.head_w = net.w_rgb; d.head_out = ws.hp; d.head_n = 3; d.head_stride = kHeadStride; d.head_slot0 = kRgbSlot0;
        d.store_c = act ? 1 : 0;   // the backward pass needs the view layer's activation (ReLU')
      }
      if (!tc) {
        CK(launch_dense_simt(d, s));
      } else {
        for (int i = 0; i < L.nseg; ++i)
          if (make_tmap_2d(c, &d.tmA[i], d.A[i], (uint64_t)M, (uint64_t)L.K[i], (uint64_t)L.K[i], 128)) return 1;
        if (make_tmap_2d(c, &d.tmC, d.C, (uint64_t)M, (uint64_t)L.N, (uint64_t)L.N, 128)) return 1;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (c->profiling) {
          const size_t idx = c->recs.size() * 2;
          while (c->ev.size() < idx + 2) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            c->ev.push_back(e);
          }
          e0 = c->ev[idx];
          e1 = c->ev[idx + 1];
          CK(cudaEventRecord(e0, s));
        }
        // CTA-pair kernel for the wide (fine-net) layers; the narrow coarse layers (N = 256: one n-tile, K = 256)
        // are epilogue-bound and measured faster on the single-CTA kernel
        if (c->pair_kernel && L.BN == 256 && L.N >= 512) CK(launch_dense_tc2(d, c->num_sms, s));
        else CK(launch_dense_tc(d, c->num_sms, s));
        if (c->profiling) {
          CK(cudaEventRecord(e1, s));
          c->recs.push_back({net_id, 2.0 * (double)P_rows * (double)L.N * (double)L.in_ref});
        }
      }
      c->launches++;
    } else if (fuse_heads) {
      continue;   // computed in the producing layer's epilogue
    } else if (st.kind == 1) {
      CK(launch_head(src_ptr(st.in[0]), net.W, net.w_alpha, net.b_alpha, 1, ws.raw, 3, P_rows, s));
      c->launches++;
    } else {
      CK(launch_head(src_ptr(st.in[0]), net.W / 2, net.w_rgb, net.b_rgb, 3, ws.raw, 0, P_rows, s));
      c->launches++;
    }
  }
  if (fuse_heads) {
    CK(launch_finalize_raw(ws.hp, kHeadStride, 0, a_slots, kRgbSlot0, r_slots, net.b_alpha, net.b_rgb, ws.raw, P_rows, s));
    c->launches++;
  }
  return 0;
}

}  // namespace

// =================================================================================================
extern "C" {

int mofa_b200_abi_version(void) { return MOFA_B200_ABI_VERSION; }
const char* mofa_b200_last_error(void) { return g_err.c_str(); }

int mofa_b200_create(mofa_b200_ctx** out, int device) {
  if (!out) return fail("mofa_b200_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail("mofa_b200_create: no CUDA device (%s); this engine has no CPU path", cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail("mofa_b200_create: device %d out of range (%d devices)", device, count);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail("mofa_b200_create: device %d is sm_%d%d; this library contains sm_100a code only", device,
                prop.major, prop.minor);
  mofa_b200_ctx* c = new mofa_b200_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    delete c;
    return fail("mofa_b200_create: cuTensorMapEncodeTiled not available from the driver");
  }
  c->encode = reinterpret_cast<EncodeTiledFn>(fn);
  e = mofa::dense_tc_configure();
  if (e == cudaSuccess) e = mofa::dense_tc2_configure();
  {
    const char* v = getenv("MOFA_B200_DENSE_1CTA");
    c->pair_kernel = !(v && v[0] == '1');
    v = getenv("MOFA_B200_NO_FUSED_COARSE");
    c->fused_coarse = !(v && v[0] == '1');
    v = getenv("MOFA_B200_FINE_PER_LAYER");
    c->chain_fine = !(v && v[0] == '1');
    if (e == cudaSuccess) e = mofa::fine_chain_configure();
    v = getenv("MOFA_B200_FP8");
    if (v && strncmp(v, "layers=", 7) == 0) c->fp8_layers = atoi(v + 7);         // the first K plain layers (parity sweep)
    else if (v && (strcmp(v, "1") == 0 || strcmp(v, "all") == 0)) c->fp8_layers = 1000;   // all of them
    v = getenv("MOFA_B200_COARSE_FP16");
    c->split_coarse = !(v && v[0] == '1');
    if (e == cudaSuccess) e = mofa::coarse_fused_configure();
    if (e == cudaSuccess) e = mofa::coarse_split_configure();
    if (e == cudaSuccess) e = mofa::wgrad_configure();
  }
  if (e != cudaSuccess) {
    delete c;
    return fail("mofa_b200_create: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
  }
  const int nlat[4] = {1, kNExp, kNShape, kNTex};
  for (int i = 1; i < 4; ++i) {
    e = cudaMalloc(reinterpret_cast<void**>(&c->lat[i]), sizeof(float) * nlat[i]);
    if (e != cudaSuccess) {
      delete c;
      return fail("mofa_b200_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    }
  }
  *out = c;
  return 0;
}

int mofa_b200_destroy(mofa_b200_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  for (int i = 0; i < 2; ++i) free_net(c->nets[i]);
  for (int i = 1; i < 4; ++i) cudaFree(c->lat[i]);
  for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
  cudaFree(c->chain_maps);
  cudaFree(c->chain_layers);
  delete c;
  return 0;
}

int64_t mofa_b200_launch_count(mofa_b200_ctx* c) { return c ? c->launches : 0; }

int mofa_b200_profile_enable(mofa_b200_ctx* c, int on) {
  if (!c) return fail("profile_enable: ctx is NULL");
  c->profiling = on != 0;
  c->recs.clear();
  return 0;
}

int mofa_b200_profile_read(mofa_b200_ctx* c, double* out6) {
  if (!c || !out6) return fail("profile_read: NULL argument");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  for (int i = 0; i < 6; ++i) out6[i] = 0.0;
  for (size_t i = 0; i < c->recs.size(); ++i) {
    float ms = 0.0f;
    CK(cudaEventElapsedTime(&ms, c->ev[2 * i], c->ev[2 * i + 1]));
    const int n = c->recs[i].net;
    out6[3 * n + 0] += ms;
    out6[3 * n + 1] += c->recs[i].flops;
    out6[3 * n + 2] += 1.0;
  }
  c->recs.clear();
  return 0;
}

int mofa_b200_load_weights(mofa_b200_ctx* c, int net_id, int W, int D, const float* const* t, int n_tensors,
                           void* stream) {
  if (!c) return fail("load_weights: ctx is NULL");
  if (net_id < 0 || net_id > 1) return fail("load_weights: net must be 0 (coarse) or 1 (fine)");
  if (W % 256 != 0 || W <= 0) return fail("load_weights: W=%d must be a positive multiple of 256", W);
  if (D < 6) return fail("load_weights: D=%d must be >= 6 (skipMLP(D, skip=4))", D);
  const int n2 = D - 5;  // layers in linears2 of each skipMLP: 1 + (D - 6)
  const int expect = 2 * (4 + 2 * (5 + n2) + 3);
  if (n_tensors != expect) return fail("load_weights: expected %d tensors for D=%d, got %d", expect, D, n_tensors);
  CK(cudaSetDevice(c->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Net& net = c->nets[net_id];
  const bool imp = c->importing;
  if (!imp && !t) return fail("load_weights: tensors is NULL");
  if (!imp && net.loaded && net.W == W && net.D == D) {
    // Same architecture (a training step changed the values): repack into the existing buffers — no allocation, tensor
    // maps and the fused-kernel tables stay valid.
    const int nd = static_cast<int>(net.layers.size());
    for (int li = 0; li < nd; ++li) {
      Layer& L = net.layers[li];
      const float* w = t[2 * li];
      const float* b = t[2 * li + 1];
      for (int i = 0; i < L.nseg; ++i) {
        CK(launch_pack_weight(w, L.in_total, L.seg_c0[i], L.seg_kreal[i], L.K[i], L.N, L.w[i], s));
        CK(launch_pack_weight_t(w, L.in_total, L.seg_c0[i], L.seg_kreal[i], L.rows_t[i], L.N, L.wt[i], s));
        c->launches += 2;
        if (L.wlo[i]) {
          CK(launch_pack_weight_lo(w, L.in_total, L.seg_c0[i], L.seg_kreal[i], L.K[i], L.N, L.wlo[i], s));
          c->launches++;
        }
      }
      if (L.w8) {
        CK(launch_pack_weight_fp8(w, L.in_total, L.seg_c0[0], L.seg_kreal[0], L.N, 8.0f, L.w8, L.colscale8, s));
        c->launches++;
      }
      CK(cudaMemcpyAsync(L.bias_raw, b, sizeof(float) * L.N, cudaMemcpyDeviceToDevice, s));
      if (L.fold_n > 0)
        CK(cudaMemcpy2DAsync(L.fold_w, sizeof(float) * L.fold_n, w + L.fold_c0, sizeof(float) * L.in_total,
                             sizeof(float) * L.fold_n, L.N, cudaMemcpyDeviceToDevice, s));
    }
    CK(cudaMemcpyAsync(net.w_alpha, t[2 * nd], sizeof(float) * W, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(net.b_alpha, t[2 * nd + 1], sizeof(float), cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(net.w_rgb, t[2 * nd + 2], sizeof(float) * 3 * (W / 2), cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(net.b_rgb, t[2 * nd + 3], sizeof(float) * 3, cudaMemcpyDeviceToDevice, s));
    if (net.w_view)   // view-direction columns of linear_view_xyBMuv (the last dense layer)
      CK(cudaMemcpy2DAsync(net.w_view, sizeof(float) * kPeView, t[2 * (nd - 1)], sizeof(float) * (kPeView + W),
                           sizeof(float) * kPeView, W / 2, cudaMemcpyDeviceToDevice, s));
    if (c->latents_set && fold_net(c, net, s)) return 1;
    return 0;
  }
  free_net(net);
  for (int i = 0; i < 4; ++i) c->chain_key[i] = nullptr;     // the chain kernel's tables point into the freed network
  net.W = W;
  net.D = D;
  int ti = 0;
  auto next = [&](const float*& w, const float*& b) {
    w = imp ? nullptr : t[ti];
    b = imp ? nullptr : t[ti + 1];
    ti += 2;
  };
  const float *w, *b;
  ProgBuilder pb(net);
  int li = 0;
  // xyzEncode: skipMLP(D=3, skip=None) -> 4 layers; Linear0 input = [PE(63), exp_mod(30)]
  for (int i = 0; i < 4; ++i) {
    next(w, b);
    LayerSpec sp{};
    sp.N = W;
    sp.nseg = 1;
    if (i == 0) {
      sp.in_total = kPeXyz + kNExp;
      sp.seg_c0[0] = 0; sp.seg_k[0] = kPeXyz; sp.seg_kpad[0] = pad64(kPeXyz);
      sp.fold_c0 = kPeXyz; sp.fold_n = kNExp; sp.fold_lat = LAT_EXP;
    } else {
      sp.in_total = W;
      sp.seg_c0[0] = 0; sp.seg_k[0] = W; sp.seg_kpad[0] = W;
    }
    if (build_layer(c, net, sp, w, b, s)) return 1;
    pb.dense(li++, pb.cur, -1);
  }
  // two skipMLP(D, skip=4): latent = shape (linear_BiM_xyz) then texture (linear_uv_xyzBiM)
  for (int blk = 0; blk < 2; ++blk) {
    const int nl = blk == 0 ? kNShape : kNTex;
    const int lat = blk == 0 ? LAT_SHAPE : LAT_TEX;
    const int x_in = pb.cur;     // xyz_code / sigmaCodes: must survive until linears2.Linear0
    pb.pinned = x_in;
    for (int i = 0; i < 5; ++i) {   // linears1.Linear0..4
      next(w, b);
      LayerSpec sp{};
      sp.N = W;
      sp.nseg = 1;
      if (i == 0) {
        sp.in_total = nl + W;
        sp.seg_c0[0] = nl; sp.seg_k[0] = W; sp.seg_kpad[0] = W;
        sp.fold_c0 = 0; sp.fold_n = nl; sp.fold_lat = lat;
      } else {
        sp.in_total = W;
        sp.seg_c0[0] = 0; sp.seg_k[0] = W; sp.seg_kpad[0] = W;
      }
      if (build_layer(c, net, sp, w, b, s)) return 1;
      pb.dense(li++, pb.cur, -1);
    }
    for (int i = 0; i < n2; ++i) {  // linears2.Linear0..: Linear0 input = cat[x_in(lat, x), h]
      next(w, b);
      LayerSpec sp{};
      sp.N = W;
      if (i == 0) {
        sp.nseg = 2;
        sp.in_total = nl + 2 * W;
        sp.seg_c0[0] = nl;     sp.seg_k[0] = W; sp.seg_kpad[0] = W;
        sp.seg_c0[1] = nl + W; sp.seg_k[1] = W; sp.seg_kpad[1] = W;
        sp.fold_c0 = 0; sp.fold_n = nl; sp.fold_lat = lat;
        if (build_layer(c, net, sp, w, b, s)) return 1;
        const int h = pb.cur;
        pb.dense(li++, x_in, h);
        pb.pinned = -1;
      } else {
        sp.nseg = 1;
        sp.in_total = W;
        sp.seg_c0[0] = 0; sp.seg_k[0] = W; sp.seg_kpad[0] = W;
        if (build_layer(c, net, sp, w, b, s)) return 1;
        pb.dense(li++, pb.cur, -1);
      }
    }
    if (blk == 0) {   // alpha = alpha_linear(sigmaCodes)   (model.py:130)
      net.program.back().head = 1;
      Step st{1, -1, {pb.cur, -1}, -1, 0, {-3, -3}, -1};
      net.program.push_back(st);
    }
  }
  // linear_view_xyBMuv: Linear(27 + W -> W/2) on cat[views, rgbCodes]
  {
    next(w, b);
    LayerSpec sp{};
    sp.N = W / 2;
    sp.nseg = 2;
    sp.in_total = kPeView + W;
    sp.seg_c0[0] = 0;       sp.seg_k[0] = kPeView; sp.seg_kpad[0] = pad64(kPeView);
    sp.seg_c0[1] = kPeView; sp.seg_k[1] = W;       sp.seg_kpad[1] = W;
    if (build_layer(c, net, sp, w, b, s)) return 1;
    pb.dense(li++, SRC_V, pb.cur);
    if (W == 256) {   // fp32 view-direction columns for the per-ray view vector of the split-precision kernel
      if (dev_alloc_data(net, reinterpret_cast<void**>(&net.w_view), sizeof(float) * kPeView * (W / 2))) return 1;
      if (!imp)
        CK(cudaMemcpy2DAsync(net.w_view, sizeof(float) * kPeView, w, sizeof(float) * (kPeView + W), sizeof(float) * kPeView,
                             W / 2, cudaMemcpyDeviceToDevice, s));
    }
  }
  // alpha_linear (W -> 1), rgb_linear (W/2 -> 3): fp32 copies for the SIMT heads
  next(w, b);
  if (dev_alloc_data(net, reinterpret_cast<void**>(&net.w_alpha), sizeof(float) * W)) return 1;
  if (dev_alloc_data(net, reinterpret_cast<void**>(&net.b_alpha), sizeof(float))) return 1;
  if (!imp) {
    CK(cudaMemcpyAsync(net.w_alpha, w, sizeof(float) * W, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(net.b_alpha, b, sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  next(w, b);
  if (dev_alloc_data(net, reinterpret_cast<void**>(&net.w_rgb), sizeof(float) * 3 * (W / 2))) return 1;
  if (dev_alloc_data(net, reinterpret_cast<void**>(&net.b_rgb), sizeof(float) * 3)) return 1;
  if (!imp) {
    CK(cudaMemcpyAsync(net.w_rgb, w, sizeof(float) * 3 * (W / 2), cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(net.b_rgb, b, sizeof(float) * 3, cudaMemcpyDeviceToDevice, s));
  }
  {
    for (auto it = net.program.rbegin(); it != net.program.rend(); ++it)
      if (it->kind == 0) {
        it->head = 2;   // the view layer feeds rgb_linear only
        break;
      }
    Step st{2, -1, {pb.cur, -1}, -1, 0, {-3, -3}, -1};
    net.program.push_back(st);
  }
  if (W == 256) {   // tables for the fused persistent kernel (coarse_fused.cu)
    std::vector<FusedLayerDesc> fl;
    std::vector<CUtensorMap> maps;
    int prev = -1;
    bool ok = true;
    for (const Step& st : net.program) {
      if (st.kind != 0) continue;
      const Layer& L = net.layers[st.layer];
      FusedLayerDesc d{};
      d.bias = L.bias_eff;
      d.n_out = L.N;
      d.store = st.head == 2 ? 0 : 1;
      d.head = st.head;
      int prim = -1;
      for (int i = 0; i < L.nseg; ++i)
        if (st.in_step[i] == prev || (prev == -1 && st.in_step[i] == -1)) prim = i;
      if (prim < 0) { ok = false; break; }
      d.kb_prim = L.K[prim] / 64;
      d.map_prim = static_cast<int>(maps.size());
      maps.push_back(L.tmB[prim]);
      if (L.nseg == 2) {
        const int sec = 1 - prim;
        d.kb_sec = L.K[sec] / 64;
        d.sec_kind = st.in_step[sec] == -2 ? 2 : 1;
        d.map_sec = static_cast<int>(maps.size());
        maps.push_back(L.tmB[sec]);
        if (d.sec_kind == 1) {   // mark the producer of the parked tensor
          if (st.in_step[sec] < 0 || st.in_step[sec] >= (int)fl.size()) { ok = false; break; }
          fl[st.in_step[sec]].save = 1;
        }
      }
      fl.push_back(d);
      prev = st.ord;
    }
    if (ok) {
      if (dev_alloc(net, reinterpret_cast<void**>(&net.fused_layers), sizeof(FusedLayerDesc) * fl.size())) return 1;
      if (dev_alloc(net, reinterpret_cast<void**>(&net.fused_wmaps), sizeof(CUtensorMap) * maps.size())) return 1;
      CK(cudaMemcpyAsync(net.fused_layers, fl.data(), sizeof(FusedLayerDesc) * fl.size(), cudaMemcpyHostToDevice, s));
      CK(cudaMemcpyAsync(net.fused_wmaps, maps.data(), sizeof(CUtensorMap) * maps.size(), cudaMemcpyHostToDevice, s));
      CK(cudaStreamSynchronize(s));   // the host vectors go out of scope
      net.fused_n = static_cast<int>(fl.size());
    }
  }
  if (W == 256 && build_split_table(c, net, s)) return 1;
  net.loaded = true;
  if (!imp && c->latents_set && fold_net(c, net, s)) return 1;
  return 0;
}

// ---- packed-weight blob: the engine's own layout of one network, for an on-disk cache ----------
namespace {
struct PackedHeader {
  char magic[8];          // "MOFAPK02"
  int32_t W, D, n_allocs, reserved;
  uint64_t total_bytes;   // header + table of sizes + data
};
const char kPackedMagic[8] = {'M', 'O', 'F', 'A', 'P', 'K', '0', '2'};
size_t packed_total(const Net& n) {
  size_t t = sizeof(PackedHeader) + sizeof(uint64_t) * n.data_allocs.size();
  for (const auto& a : n.data_allocs) t += (a.second + 15) / 16 * 16;
  return t;
}
}  // namespace

size_t mofa_b200_packed_bytes(mofa_b200_ctx* c, int net_id) {
  if (!c || net_id < 0 || net_id > 1 || !c->nets[net_id].loaded) return 0;
  return packed_total(c->nets[net_id]);
}

int mofa_b200_export_packed(mofa_b200_ctx* c, int net_id, void* host_dst, size_t bytes, void* stream) {
  if (!c || net_id < 0 || net_id > 1 || !c->nets[net_id].loaded) return fail("export_packed: net %d not loaded", net_id);
  const Net& n = c->nets[net_id];
  if (!host_dst || bytes < packed_total(n)) return fail("export_packed: buffer too small (%zu < %zu)", bytes, packed_total(n));
  CK(cudaSetDevice(c->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* p = static_cast<uint8_t*>(host_dst);
  PackedHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, kPackedMagic, 8);
  h.W = n.W; h.D = n.D; h.n_allocs = static_cast<int32_t>(n.data_allocs.size());
  h.total_bytes = packed_total(n);
  memcpy(p, &h, sizeof(h));
  p += sizeof(h);
  for (const auto& a : n.data_allocs) {
    const uint64_t sz = a.second;
    memcpy(p, &sz, sizeof(sz));
    p += sizeof(sz);
  }
  for (const auto& a : n.data_allocs) {
    CK(cudaMemcpyAsync(p, a.first, a.second, cudaMemcpyDeviceToHost, s));
    p += (a.second + 15) / 16 * 16;
  }
  CK(cudaStreamSynchronize(s));
  return 0;
}

int mofa_b200_import_packed(mofa_b200_ctx* c, int net_id, const void* host_src, size_t bytes, void* stream) {
  if (!c) return fail("import_packed: ctx is NULL");
  if (net_id < 0 || net_id > 1) return fail("import_packed: net must be 0 or 1");
  if (!host_src || bytes < sizeof(PackedHeader)) return fail("import_packed: blob too small");
  PackedHeader h;
  memcpy(&h, host_src, sizeof(h));
  if (memcmp(h.magic, kPackedMagic, 8) != 0) return fail("import_packed: not a mofa_b200 packed-weight blob (or an older layout)");
  if (h.total_bytes != bytes) return fail("import_packed: blob is %zu bytes, header says %llu", bytes, (unsigned long long)h.total_bytes);
  const int n2 = h.D - 5;
  const int expect = 2 * (4 + 2 * (5 + n2) + 3);
  CK(cudaSetDevice(c->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  free_net(c->nets[net_id]);
  c->importing = true;
  const int rc = mofa_b200_load_weights(c, net_id, h.W, h.D, nullptr, expect, stream);
  c->importing = false;
  if (rc) return rc;
  Net& n = c->nets[net_id];
  if (static_cast<int>(n.data_allocs.size()) != h.n_allocs || packed_total(n) != bytes) {
    free_net(n);
    return fail("import_packed: blob layout does not match this build (allocs %d vs %zu)", h.n_allocs, n.data_allocs.size());
  }
  const uint8_t* p = static_cast<const uint8_t*>(host_src) + sizeof(PackedHeader);
  for (const auto& a : n.data_allocs) {
    uint64_t sz;
    memcpy(&sz, p, sizeof(sz));
    p += sizeof(sz);
    if (sz != a.second) {
      free_net(n);
      return fail("import_packed: tensor size mismatch");
    }
  }
  for (const auto& a : n.data_allocs) {
    CK(cudaMemcpyAsync(a.first, p, a.second, cudaMemcpyHostToDevice, s));
    p += (a.second + 15) / 16 * 16;
  }
  CK(cudaStreamSynchronize(s));      // the caller's buffer may go away
  if (c->latents_set && fold_net(c, n, s)) return 1;
  for (int i = 0; i < 4; ++i) c->chain_key[i] = nullptr;     // device tables point into the freed network
  return 0;
}

int mofa_b200_set_latents(mofa_b200_ctx* c, const float* shape50, const float* exp_mod30, const float* tex256,
                          void* stream) {
  if (!c) return fail("set_latents: ctx is NULL");
  if (!shape50 || !exp_mod30 || !tex256) return fail("set_latents: NULL latent pointer");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaMemcpyAsync(c->lat[LAT_EXP], exp_mod30, sizeof(float) * kNExp, cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpyAsync(c->lat[LAT_SHAPE], shape50, sizeof(float) * kNShape, cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpyAsync(c->lat[LAT_TEX], tex256, sizeof(float) * kNTex, cudaMemcpyDeviceToDevice, s));
  c->latents_set = true;
  for (int i = 0; i < 2; ++i)
    if (c->nets[i].loaded && fold_net(c, c->nets[i], s)) return 1;
  return 0;
}

static int effective_chunk(int64_t n_rays, int chunk_rays, int S_last = 128) {
  // Default: at most 4144 rays per pass (with 64 + 128 samples: 37 slabs of the fine-net chain kernel = 2072 m-blocks,
  // 14 rounds of coarse pair tiles), and the passes of a call BALANCED: n rays are split into ceil(n / 4144) passes of
  // equal length, rounded up to whole slabs — a rank's 80 000 rays of an 8-way sharded 800x800 frame become 19 passes of
  // 4032 rays + one of 3392 instead of 19 x 4144 + a 1264-ray stub whose partial rounds and fixed per-pass work were the
  // 8-GPU tail of round 1.  An explicit chunk_rays is taken as is (results never depend on the chunking).
  if (chunk_rays > 0) return static_cast<int>(chunk_rays < n_rays ? chunk_rays : (n_rays > 0 ? n_rays : 1));
  const int64_t cap = 4144;
  if (n_rays <= cap) return static_cast<int>(n_rays > 0 ? n_rays : 1);
  const int64_t passes = (n_rays + cap - 1) / cap;
  int64_t per = (n_rays + passes - 1) / passes;
  const int64_t unit = std::max<int64_t>(1, static_cast<int64_t>(chain_slab_mb()) * 256 / (S_last > 0 ? S_last : 1));
  per = (per + unit - 1) / unit * unit;
  if (per > cap) per = cap;
  return static_cast<int>(per);
}

size_t mofa_b200_workspace_bytes(mofa_b200_ctx* c, int64_t n_rays, int n_samples, int n_importance, int chunk_rays) {
  if (!c) return 0;
  const int S_f = n_importance > 0 ? n_samples + n_importance : 0;
  const int ch = effective_chunk(n_rays, chunk_rays, S_f > 0 ? S_f : n_samples);
  int Wmax = max_width(c);
  if (Wmax == 0) Wmax = 1024;
  return carve(c, nullptr, ch, n_samples, S_f, Wmax).total + 1024;
}

int mofa_b200_render_rays_fwd(mofa_b200_ctx* c, const mofa_b200_render_args* a, void* stream) {
  if (!c || !a) return fail("render_rays_fwd: NULL argument");
  if (a->struct_size != sizeof(mofa_b200_render_args))
    return fail("render_rays_fwd: struct_size %u != %zu (ABI mismatch)", a->struct_size, sizeof(mofa_b200_render_args));
  if (a->n_rays < 0) return fail("render_rays_fwd: n_rays < 0");
  if (a->n_rays == 0) return 0;
  if (!a->rays || a->ray_stride < 11) return fail("render_rays_fwd: rays NULL or ray_stride < 11");
  const int S_c = a->n_samples, N_i = a->n_importance;
  const bool fine = (N_i > 0) && a->run_fine;
  const int S_f = fine ? S_c + N_i : 0;
  if (S_c < 2 || S_c > 256) return fail("render_rays_fwd: n_samples=%d out of range [2,256]", S_c);
  if (fine && (S_c < 3 || S_f > 256)) return fail("render_rays_fwd: n_samples+n_importance=%d out of range", S_f);
  Net& nc = c->nets[MOFA_NET_COARSE];
  if (!nc.loaded) return fail("render_rays_fwd: coarse network not loaded");
  if (a->fine_net < 0 || a->fine_net > 1) return fail("render_rays_fwd: fine_net must be 0 or 1");
  Net& nf = c->nets[a->fine_net];
  if (fine && !nf.loaded) return fail("render_rays_fwd: fine network (net %d) not loaded", a->fine_net);
  if (!c->latents_set) return fail("render_rays_fwd: mofa_b200_set_latents has not been called");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  int ch = effective_chunk(a->n_rays, a->chunk_rays, S_f > 0 ? S_f : S_c);
  if ((a->flags & MOFA_FLAG_GEMM_SIMT) && ch > 256) ch = 256;    // verification path: per-layer buffers, one row per point
  const int Wmax = max_width(c);
  if (!a->workspace) return fail("render_rays_fwd: workspace is NULL");
  uint8_t* wbase = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(a->workspace), 1024));
  const size_t slack = wbase - static_cast<uint8_t*>(a->workspace);
  Workspace ws = carve(c, wbase, ch, S_c, S_f, Wmax);
  if (ws.total + slack > a->workspace_bytes)
    return fail("render_rays_fwd: workspace too small (%zu < %zu)", a->workspace_bytes, ws.total + slack);
  const int lindisp = (a->flags & MOFA_FLAG_LINDISP) ? 1 : 0;
  const int white = (a->flags & MOFA_FLAG_WHITE_BKGD) ? 1 : 0;
  const int S_last = fine ? S_f : S_c;

  for (int64_t r0 = 0; r0 < a->n_rays; r0 += ch) {
    const int64_t n = (a->n_rays - r0 < ch) ? (a->n_rays - r0) : ch;
    const float* rays = a->rays + r0 * a->ray_stride;
    auto off = [&](float* p, int64_t per_ray) -> float* { return p ? p + r0 * per_ray : nullptr; };
    auto coff = [&](const float* p, int64_t per_ray) -> const float* { return p ? p + r0 * per_ray : nullptr; };

    // ---- coarse pass
    CK(launch_zvals_coarse(rays, a->ray_stride, n, S_c, lindisp, a->perturb, coff(a->t_rand, S_c), a->seed, r0,
                           ws.z_c, s));
    {
      const bool sp = use_split(c, nc, a->flags);
      CK(launch_encode_rays(rays, a->ray_stride, ws.z_c, n, S_c, kMultires, kMultiresViews, ws.X0, sp ? nullptr : ws.V, s,
                            sp ? ws.X0lo : nullptr));
    }
    c->launches += 2;
    const ViewSrc view_c{rays + 8, a->ray_stride, S_c};
    if (run_program(c, nc, ws, n * S_c, a->flags, s, nullptr, &view_c)) return 1;
    float* o_rgb = fine ? off(a->rgb0, 3) : off(a->rgb, 3);
    float* o_disp = fine ? off(a->disp0, 1) : off(a->disp, 1);
    float* o_acc = fine ? off(a->acc0, 1) : off(a->acc, 1);
    CK(launch_composite(ws.raw, ws.z_c, rays + 3, a->ray_stride, coff(a->noise_c, S_c), a->raw_noise_std,
                        a->seed, r0, n, S_c, white, o_rgb, o_disp, o_acc, ws.w_c, nullptr, s));
    c->launches++;
    const float* z_last = ws.z_c;
    if (fine) {
      // ---- hierarchical resampling + fine pass
      const int det = (a->perturb == 0.0f) ? 1 : 0;
      CK(launch_sample_pdf_merge(ws.z_c, ws.w_c, coff(a->u, N_i), det, a->seed, r0, n, S_c, N_i, nullptr, ws.z_f,
                                 off(a->z_std, 1), s));
      {
        const bool sp = use_split(c, nf, a->flags);
        CK(launch_encode_rays(rays, a->ray_stride, ws.z_f, n, S_f, kMultires, kMultiresViews, ws.X0, sp ? nullptr : ws.V, s,
                              sp ? ws.X0lo : nullptr));
      }
      c->launches += 2;
      const ViewSrc view_f{rays + 8, a->ray_stride, S_f};
      if (run_program(c, nf, ws, n * S_f, a->flags, s, nullptr, &view_f)) return 1;
      CK(launch_composite(ws.raw, ws.z_f, rays + 3, a->ray_stride, coff(a->noise_f, S_f), a->raw_noise_std,
                          a->seed + 0x9E3779B97F4A7C15ull, r0, n, S_f, white, off(a->rgb, 3), off(a->disp, 1),
                          off(a->acc, 1), a->weights ? off(a->weights, S_f) : nullptr, nullptr, s));
      c->launches++;
      z_last = ws.z_f;
    } else if (a->weights) {
      CK(launch_copy_f32(ws.w_c, off(a->weights, S_c), n * S_c, s));
      c->launches++;
    }
    if (a->raw) {
      CK(launch_copy_f32(ws.raw, off(a->raw, (int64_t)S_last * 4), n * S_last * 4, s));
      c->launches++;
    }
    if (a->z_vals) {
      CK(launch_copy_f32(z_last, off(a->z_vals, S_last), n * S_last, s));
      c->launches++;
    }
  }
  return 0;
}

size_t mofa_b200_query_workspace_bytes(mofa_b200_ctx* c, int64_t n_pts) {
  if (!c) return 0;
  int Wmax = max_width(c);
  if (Wmax == 0) Wmax = 1024;
  return carve(c, nullptr, n_pts, 1, 0, Wmax).total + 1024;
}

int mofa_b200_run_network(mofa_b200_ctx* c, int net_id, const float* pts, const float* viewdirs, int64_t n_pts,
                          float* raw_out, uint32_t flags, void* workspace, size_t workspace_bytes, void* stream) {
  if (!c) return fail("run_network: ctx is NULL");
  if (net_id < 0 || net_id > 1 || !c->nets[net_id].loaded) return fail("run_network: net %d not loaded", net_id);
  if (!c->latents_set) return fail("run_network: mofa_b200_set_latents has not been called");
  if (n_pts == 0) return 0;
  if (!pts || !viewdirs || !raw_out || !workspace) return fail("run_network: NULL pointer");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* wbase = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(workspace), 1024));
  const size_t slack = wbase - static_cast<uint8_t*>(workspace);
  Workspace ws = carve(c, wbase, n_pts, 1, 0, max_width(c));
  if (ws.total + slack > workspace_bytes)
    return fail("run_network: workspace too small (%zu < %zu)", workspace_bytes, ws.total + slack);
  {
    const bool sp = use_split(c, c->nets[net_id], flags);
    CK(launch_encode_points(pts, viewdirs, n_pts, kMultires, kMultiresViews, ws.X0, sp ? nullptr : ws.V, s,
                            sp ? ws.X0lo : nullptr));
  }
  c->launches++;
  const ViewSrc view{viewdirs, 3, 1};
  if (run_program(c, c->nets[net_id], ws, n_pts, flags, s, nullptr, &view)) return 1;
  CK(launch_copy_f32(ws.raw, raw_out, n_pts * 4, s));
  c->launches++;
  return 0;
}

int mofa_b200_generate_rays(mofa_b200_ctx* c, int H, int W, const float* K9, const float* c2w12, float near_, float far_,
                            int64_t first_ray, int64_t n_rays, float* rays_out, int ray_stride, void* stream) {
  if (!c) return fail("generate_rays: ctx is NULL");
  if (!K9 || !c2w12 || !rays_out) return fail("generate_rays: NULL pointer");
  if (H <= 0 || W <= 0 || ray_stride < 11) return fail("generate_rays: bad sizes (H=%d W=%d stride=%d)", H, W, ray_stride);
  if (first_ray < 0 || n_rays < 0 || first_ray + n_rays > static_cast<int64_t>(H) * W)
    return fail("generate_rays: ray range [%lld, %lld) outside the %dx%d image", (long long)first_ray,
                (long long)(first_ray + n_rays), H, W);
  if (K9[0] == 0.0f || K9[4] == 0.0f) return fail("generate_rays: zero focal length");
  CK(cudaSetDevice(c->device));
  CK(launch_generate_rays(H, W, K9, c2w12, near_, far_, first_ray, n_rays, rays_out, ray_stride,
                          static_cast<cudaStream_t>(stream)));
  c->launches++;
  return 0;
}

int mofa_b200_embed(mofa_b200_ctx* c, const float* x, int64_t n, int multires, float* out, void* stream) {
  if (!c) return fail("embed: ctx is NULL");
  if (multires < 0 || multires > 16) return fail("embed: multires out of range");
  CK(cudaSetDevice(c->device));
  CK(launch_embed_f32(x, n, multires, out, static_cast<cudaStream_t>(stream)));
  c->launches++;
  return 0;
}

int mofa_b200_raw2outputs(mofa_b200_ctx* c, const float* raw, const float* z, const float* rays_d, int d_stride,
                          const float* noise, int64_t n, int S, int white_bkgd, float* rgb, float* disp, float* acc,
                          float* weights, float* depth, void* stream) {
  if (!c) return fail("raw2outputs: ctx is NULL");
  if (S < 2 || S > 256) return fail("raw2outputs: S=%d out of range [2,256]", S);
  CK(cudaSetDevice(c->device));
  CK(launch_composite(raw, z, rays_d, d_stride, noise, 0.0f, 0, 0, n, S, white_bkgd, rgb, disp, acc, weights, depth,
                      static_cast<cudaStream_t>(stream)));
  c->launches++;
  return 0;
}

int mofa_b200_sample_pdf_merge(mofa_b200_ctx* c, const float* z, const float* weights, const float* u, int64_t n,
                               int S, int N_i, float* z_samples, float* z_merged, float* z_std, void* stream) {
  if (!c) return fail("sample_pdf_merge: ctx is NULL");
  if (S < 3 || S > 256 || N_i < 1 || S + N_i > 512) return fail("sample_pdf_merge: S=%d N_i=%d out of range", S, N_i);
  CK(cudaSetDevice(c->device));
  CK(launch_sample_pdf_merge(z, weights, u, 1, 0, 0, n, S, N_i, z_samples, z_merged, z_std,
                             static_cast<cudaStream_t>(stream)));
  c->launches++;
  return 0;
}

int mofa_b200_dense(mofa_b200_ctx* c, const void* A0, const void* B0, int K0, const void* A1, const void* B1, int K1,
                    const float* bias, void* C, int64_t M, int N, int relu, int use_simt, void* stream) {
  if (!c) return fail("dense: ctx is NULL");
  if (M % 128 != 0 || N % 128 != 0 || K0 % 64 != 0 || K1 % 64 != 0 || K0 <= 0)
    return fail("dense: shape constraint violated (M%%128, N%%128, K%%64)");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DenseLaunch d;
  memset(&d, 0, sizeof(d));
  d.A[0] = static_cast<const __half*>(A0);
  d.B[0] = static_cast<const __half*>(B0);
  d.A[1] = static_cast<const __half*>(A1);
  d.B[1] = static_cast<const __half*>(B1);
  d.K[0] = K0;
  d.K[1] = (A1 && B1) ? K1 : 0;
  d.lda[0] = K0;
  d.lda[1] = K1;
  d.C = static_cast<__half*>(C);
  d.ldc = N;
  d.bias = bias;
  d.M = M;
  d.N = N;
  d.BN = (N % 256 == 0) ? 256 : 128;
  d.relu = relu;
  d.store_c = 1;
  if (use_simt == 1) {
    CK(launch_dense_simt(d, s));
  } else {
    const int nseg = d.K[1] > 0 ? 2 : 1;
    for (int i = 0; i < nseg; ++i) {
      if (make_tmap_2d(c, &d.tmA[i], d.A[i], (uint64_t)M, (uint64_t)d.K[i], (uint64_t)d.K[i], 128)) return 1;
      if (make_tmap_2d(c, &d.tmB[i], d.B[i], (uint64_t)N, (uint64_t)d.K[i], (uint64_t)d.K[i], d.BN)) return 1;
      if (make_tmap_2d(c, &d.tmB2[i], d.B[i], (uint64_t)N, (uint64_t)d.K[i], (uint64_t)d.K[i], 128)) return 1;
    }
    if (make_tmap_2d(c, &d.tmC, d.C, (uint64_t)M, (uint64_t)N, (uint64_t)N, 128)) return 1;
    if (use_simt != 2 && c->pair_kernel && d.BN == 256) CK(launch_dense_tc2(d, c->num_sms, s));
    else CK(launch_dense_tc(d, c->num_sms, s));
  }
  c->launches++;
  return 0;
}

}  // extern "C"

// =================================================================================================
// fitting: training-mode forward + backward
// =================================================================================================
namespace {

struct PassBufs {
  float *z, *raw;                 // [n,S], [P,4]
  __half *X0, *V;                 // [P,64] encodings (kept for the weight gradients of the first / view layer)
  std::vector<__half*> act;       // per dense step: [P_pad, N]
};

struct TrainWS {
  Workspace fw;                   // forward scratch shared by both passes (X0, V, hp, z_c, w_c, z_f, raw)
  PassBufs pass[2];               // 0 = coarse pass, 1 = fine pass
  std::vector<__half*> dz;        // backward: per dense step of the widest/deepest net: [P_pad, Wmax]
  __half *dX0, *dV;               // [P_pad,128]
  float* d_raw;                   // [P_pad,4]
  float* d_beff;                  // [Wmax]
  size_t total;
};

int n_dense_steps(const Net& n) {
  int k = 0;
  for (const Step& s : n.program) k += s.kind == 0;
  return k;
}

TrainWS carve_train(void* base, int64_t n, int S_c, int S_f, const Net& nc, const Net* nf) {
  TrainWS t;
  const int Wmax = nf && nf->W > nc.W ? nf->W : nc.W;
  t.fw = carve(nullptr, base, n, S_c, S_f, Wmax);
  size_t off = align_up(t.fw.total, 1024);
  uint8_t* b = static_cast<uint8_t*>(base);
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return b + o;
  };
  const int64_t P_c = (n * S_c + 127) / 128 * 128;
  const int64_t P_f = (n * (S_f > 0 ? S_f : 1) + 127) / 128 * 128;
  for (int ps = 0; ps < 2; ++ps) {
    const Net* net = ps == 0 ? &nc : nf;
    if (!net) continue;
    const int64_t P = ps == 0 ? P_c : P_f;
    const int S = ps == 0 ? S_c : S_f;
    t.pass[ps].z = reinterpret_cast<float*>(take(sizeof(float) * n * S));
    t.pass[ps].raw = reinterpret_cast<float*>(take(sizeof(float) * 4 * P));
    t.pass[ps].X0 = reinterpret_cast<__half*>(take(sizeof(__half) * 64 * P));
    t.pass[ps].V = reinterpret_cast<__half*>(take(sizeof(__half) * 64 * P));
    for (const Step& st : net->program)
      if (st.kind == 0)
        t.pass[ps].act.push_back(reinterpret_cast<__half*>(take(sizeof(__half) * (size_t)P * net->layers[st.layer].N)));
  }
  const int64_t P_max = P_f > P_c && nf ? P_f : P_c;
  int nd = n_dense_steps(nc);
  if (nf && n_dense_steps(*nf) > nd) nd = n_dense_steps(*nf);
  for (int k = 0; k < nd; ++k) t.dz.push_back(reinterpret_cast<__half*>(take(sizeof(__half) * (size_t)P_max * Wmax)));
  t.dX0 = reinterpret_cast<__half*>(take(sizeof(__half) * (size_t)P_max * 128));
  t.dV = reinterpret_cast<__half*>(take(sizeof(__half) * (size_t)P_max * 128));
  t.d_raw = reinterpret_cast<float*>(take(sizeof(float) * 4 * P_max));
  t.d_beff = reinterpret_cast<float*>(take(sizeof(float) * Wmax));
  t.total = off;
  return t;
}

// Backward of one pass through `net`: d_raw (loss-scaled) -> dX0/dV and latent gradients.
// d_params (optional): fp32 gradient buffers in the canonical (weight, bias) order of mofa_b200_load_weights
int run_backward(mofa_b200_ctx* c, Net& net, const TrainWS& t, const PassBufs& pb, int64_t P_rows, float inv_scale,
                 float* const d_lat[4], float* const* d_params, cudaStream_t s, const float* inv_dev = nullptr) {
  const int64_t M = (P_rows + 127) / 128 * 128;
  std::vector<const Step*> dense;
  for (const Step& st : net.program)
    if (st.kind == 0) dense.push_back(&st);
  const int nd = static_cast<int>(dense.size());
  auto gemm = [&](int target /* dense ordinal, -1 X0, -2 V */, int width, __half* out, const __half* mask,
                  bool rank1) -> int {
    DenseLaunch d;
    memset(&d, 0, sizeof(d));
    int nseg = 0;
    for (int cidx = 0; cidx < nd; ++cidx) {
      const Step& cs = *dense[cidx];
      const Layer& CL = net.layers[cs.layer];
      for (int i = 0; i < CL.nseg; ++i) {
        if (cs.in_step[i] != target) continue;
        if (nseg == 2) return fail("backward: tensor with more than two consumers");
        d.A[nseg] = t.dz[cidx];
        d.B[nseg] = CL.wt[i];
        d.K[nseg] = CL.N;
        d.lda[nseg] = CL.N;
        d.tmB[nseg] = CL.tmBt[i];
        d.tmB2[nseg] = CL.tmBt2[i];
        if (CL.rows_t[i] != width) return fail("backward: width mismatch (%d vs %d)", CL.rows_t[i], width);
        if (make_tmap_2d(c, &d.tmA[nseg], d.A[nseg], (uint64_t)M, (uint64_t)CL.N, (uint64_t)CL.N, 128)) return 1;
        ++nseg;
      }
    }
    if (nseg == 0) return fail("backward: tensor %d has no consumer", target);
    d.C = out;
    d.ldc = width;
    d.M = M;
    d.M_valid = M;
    d.N = width;
    d.BN = (width % 256 == 0) ? 256 : 128;
    d.relu = 0;
    d.store_c = 1;
    d.mask = mask;
    if (rank1) {
      d.r1_row = t.d_raw + 3;
      d.r1_stride = 4;
      d.r1_col = net.w_alpha;
    }
    if (make_tmap_2d(c, &d.tmC, d.C, (uint64_t)M, (uint64_t)width, (uint64_t)width, 128)) return 1;
    if (c->pair_kernel && d.BN == 256 && width >= 512) CK(launch_dense_tc2(d, c->num_sms, s));
    else CK(launch_dense_tc(d, c->num_sms, s));
    c->launches++;
    return 0;
  };
  for (int k = nd - 1; k >= 0; --k) {
    const Step& st = *dense[k];
    const Layer& L = net.layers[st.layer];
    if (st.head == 2) {   // view layer: only rgb_linear consumes it
      CK(launch_view_head_bwd(t.d_raw, net.w_rgb, pb.act[k], L.N, P_rows, t.dz[k], s));
      c->launches++;
      if (M > P_rows)     // padding rows feed the dX GEMMs and the weight-gradient reduction: must be zero, not stale
        CK(cudaMemsetAsync(t.dz[k] + P_rows * L.N, 0, sizeof(__half) * (size_t)(M - P_rows) * L.N, s));
    } else {
      if (gemm(k, L.N, t.dz[k], pb.act[k], st.head == 1)) return 1;
    }
    if (L.fold_n > 0 || d_params) {   // d(bias_eff) = column sums of dZ (adjoint of the latent fold; bias gradient)
      CK(cudaMemsetAsync(t.d_beff, 0, sizeof(float) * L.N, s));
      CK(launch_colsum(t.dz[k], L.N, P_rows, t.d_beff, s));
      c->launches++;
    }
    if (L.fold_n > 0) {
      CK(launch_fold_bwd(L.fold_w, L.fold_n, L.N, t.d_beff, inv_scale, d_lat[L.fold_lat], s, inv_dev));
      c->launches++;
    }
    if (d_params) {   // weight gradients: dW_seg += dZ^T · X_seg, db += colsum(dZ), latent columns += db (x) latent
      float* gW = d_params[2 * st.layer];
      float* gb = d_params[2 * st.layer + 1];
      CK(launch_axpy_f32(t.d_beff, inv_scale, gb, L.N, s, inv_dev));
      if (L.fold_n > 0)
        CK(launch_outer_add(t.d_beff, c->lat[L.fold_lat], L.N, L.fold_n, inv_scale, gW, L.in_ref, L.fold_c0, s, inv_dev));
      for (int i = 0; i < L.nseg; ++i) {
        const __half* X = st.in_step[i] >= 0 ? pb.act[st.in_step[i]] : (st.in_step[i] == -1 ? pb.X0 : pb.V);
        WgradLaunch W;
        memset(&W, 0, sizeof(W));
        if (make_tmap_2d(c, &W.tmA, t.dz[k], (uint64_t)M, (uint64_t)L.N, (uint64_t)L.N, 64)) return 1;
        if (make_tmap_2d(c, &W.tmB, X, (uint64_t)M, (uint64_t)L.K[i], (uint64_t)L.K[i], 64)) return 1;
        W.C = gW + L.seg_c0[i];
        W.ldc = L.in_ref;
        W.n_valid = L.seg_kreal[i];
        W.scale = inv_scale;
        W.scale_dev = inv_dev;
        W.Mp = L.N;
        W.BN = (L.K[i] % 256 == 0) ? 256 : 128;
        W.Np = (L.K[i] + W.BN - 1) / W.BN * W.BN;
        W.P = M;
        CK(launch_wgrad_tc(W, c->num_sms, s));
        c->launches++;
      }
      c->launches += 2;
    }
  }
  if (d_params) {   // heads: alpha_linear (W -> 1) on sigmaCodes, rgb_linear (W/2 -> 3) on the view layer's activation
    const int n_dense = nd;
    for (int k = 0; k < nd; ++k) {
      const Step& st = *dense[k];
      if (st.head == 1)
        CK(launch_head_wgrad(t.d_raw, 3, 1, pb.act[k], net.W, P_rows, inv_scale, d_params[2 * (n_dense + 0)],
                             d_params[2 * (n_dense + 0) + 1], s, inv_dev));
      else if (st.head == 2)
        CK(launch_head_wgrad(t.d_raw, 0, 3, pb.act[k], net.W / 2, P_rows, inv_scale, d_params[2 * (n_dense + 1)],
                             d_params[2 * (n_dense + 1) + 1], s, inv_dev));
    }
    c->launches += 2;
  }
  if (gemm(-1, 128, t.dX0, nullptr, false)) return 1;
  if (gemm(-2, 128, t.dV, nullptr, false)) return 1;
  return 0;
}

}  // namespace

extern "C" {

size_t mofa_b200_train_workspace_bytes(mofa_b200_ctx* c, int64_t n_rays, int n_samples, int n_importance, int fine_net) {
  if (!c || !c->nets[0].loaded) return 0;
  const bool fine = n_importance > 0;
  if (fine && (fine_net < 0 || fine_net > 1 || !c->nets[fine_net].loaded)) return 0;
  return carve_train(nullptr, n_rays, n_samples, fine ? n_samples + n_importance : 0, c->nets[0],
                     fine ? &c->nets[fine_net] : nullptr).total + 1024;
}

int mofa_b200_render_rays_train_fwd(mofa_b200_ctx* c, const mofa_b200_render_args* a, void* stream) {
  if (!c || !a) return fail("train_fwd: NULL argument");
  if (a->struct_size != sizeof(mofa_b200_render_args)) return fail("train_fwd: struct_size mismatch");
  if (a->n_rays <= 0) return fail("train_fwd: n_rays must be positive");
  if (!a->rays || a->ray_stride < 11) return fail("train_fwd: rays NULL or ray_stride < 11");
  if (a->flags & MOFA_FLAG_GEMM_SIMT) return fail("train_fwd: the SIMT verification kernel has no training mode");
  const int S_c = a->n_samples, N_i = a->n_importance;
  const bool fine = (N_i > 0) && a->run_fine;
  const int S_f = fine ? S_c + N_i : 0;
  if (S_c < 2 || S_c > 256 || (fine && (S_c < 3 || S_f > 256))) return fail("train_fwd: sample counts out of range");
  Net& nc = c->nets[0];
  if (!nc.loaded) return fail("train_fwd: coarse network not loaded");
  if (fine && (a->fine_net < 0 || a->fine_net > 1 || !c->nets[a->fine_net].loaded))
    return fail("train_fwd: fine network not loaded");
  Net* nf = fine ? &c->nets[a->fine_net] : nullptr;
  if (nc.W > 1024 || (nf && nf->W > 1024))
    return fail("train_fwd: networks wider than 1024 are not supported in training mode (the unfused head path reads the "
                "ping-pong buffers, which the activation-keeping forward does not write)");
  if (!c->latents_set) return fail("train_fwd: set_latents has not been called");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!a->workspace) return fail("train_fwd: workspace is NULL");
  uint8_t* wbase = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(a->workspace), 1024));
  const size_t slack = wbase - static_cast<uint8_t*>(a->workspace);
  TrainWS t = carve_train(wbase, a->n_rays, S_c, S_f, nc, nf);
  if (t.total + slack > a->workspace_bytes)
    return fail("train_fwd: workspace too small (%zu < %zu)", a->workspace_bytes, t.total + slack);
  const int64_t n = a->n_rays;
  const int lindisp = (a->flags & MOFA_FLAG_LINDISP) ? 1 : 0;
  const int white = (a->flags & MOFA_FLAG_WHITE_BKGD) ? 1 : 0;
  Workspace ws = t.fw;
  // Rows [n*S, round_up(n*S, 128)) of the encodings are read by the dense layers and by the weight-gradient GEMM (which
  // reduces over all padded rows) but never written by encode_rays: zero them so every padding row of every activation
  // stays finite (the workspace is uninitialised allocator memory).
  for (int ps = 0; ps < (fine ? 2 : 1); ++ps) {
    const int64_t rows = n * (ps == 0 ? S_c : S_f);
    const int64_t pad = (rows + 127) / 128 * 128 - rows;
    if (pad > 0) {
      CK(cudaMemsetAsync(t.pass[ps].X0 + rows * 64, 0, sizeof(__half) * 64 * pad, s));
      CK(cudaMemsetAsync(t.pass[ps].V + rows * 64, 0, sizeof(__half) * 64 * pad, s));
    }
  }
  // ---- coarse pass (activations kept)
  ws.X0 = t.pass[0].X0;
  ws.V = t.pass[0].V;
  CK(launch_zvals_coarse(a->rays, a->ray_stride, n, S_c, lindisp, a->perturb, a->t_rand, a->seed, 0, t.pass[0].z, s));
  CK(launch_encode_rays(a->rays, a->ray_stride, t.pass[0].z, n, S_c, kMultires, kMultiresViews, ws.X0, ws.V, s));
  c->launches += 2;
  ws.raw = t.pass[0].raw;
  if (run_program(c, nc, ws, n * S_c, a->flags, s, t.pass[0].act.data())) return 1;
  CK(launch_composite(ws.raw, t.pass[0].z, a->rays + 3, a->ray_stride, a->noise_c, a->raw_noise_std, a->seed, 0, n,
                      S_c, white, fine ? a->rgb0 : a->rgb, fine ? a->disp0 : a->disp, fine ? a->acc0 : a->acc, ws.w_c,
                      nullptr, s));
  c->launches++;
  const float* z_last = t.pass[0].z;
  const float* raw_last = t.pass[0].raw;
  if (fine) {
    const int det = (a->perturb == 0.0f) ? 1 : 0;
    CK(launch_sample_pdf_merge(t.pass[0].z, ws.w_c, a->u, det, a->seed, 0, n, S_c, N_i, nullptr, t.pass[1].z, a->z_std, s));
    ws.X0 = t.pass[1].X0;
    ws.V = t.pass[1].V;
    CK(launch_encode_rays(a->rays, a->ray_stride, t.pass[1].z, n, S_f, kMultires, kMultiresViews, ws.X0, ws.V, s));
    c->launches += 2;
    ws.raw = t.pass[1].raw;
    if (run_program(c, *nf, ws, n * S_f, a->flags, s, t.pass[1].act.data())) return 1;
    CK(launch_composite(ws.raw, t.pass[1].z, a->rays + 3, a->ray_stride, a->noise_f, a->raw_noise_std,
                        a->seed + 0x9E3779B97F4A7C15ull, 0, n, S_f, white, a->rgb, a->disp, a->acc, a->weights, nullptr,
                        s));
    c->launches++;
    z_last = t.pass[1].z;
    raw_last = t.pass[1].raw;
  } else if (a->weights) {
    CK(launch_copy_f32(ws.w_c, a->weights, n * S_c, s));
  }
  const int S_last = fine ? S_f : S_c;
  if (a->raw) CK(launch_copy_f32(raw_last, a->raw, n * S_last * 4, s));
  if (a->z_vals) CK(launch_copy_f32(z_last, a->z_vals, n * S_last, s));
  return 0;
}

int mofa_b200_render_rays_bwd(mofa_b200_ctx* c, const mofa_b200_bwd_args* a, void* stream) {
  if (!c || !a) return fail("bwd: NULL argument");
  if (a->struct_size != sizeof(mofa_b200_bwd_args)) return fail("bwd: struct_size mismatch");
  if (a->n_rays <= 0 || !a->rays || !a->workspace) return fail("bwd: bad arguments");
  if (!a->d_rays || !a->d_shape || !a->d_expmod || !a->d_tex) return fail("bwd: output pointers must not be NULL");
  const int S_c = a->n_samples, N_i = a->n_importance;
  const bool fine = (N_i > 0) && a->run_fine;
  const int S_f = fine ? S_c + N_i : 0;
  Net& nc = c->nets[0];
  Net* nf = fine ? &c->nets[a->fine_net] : nullptr;
  if (!nc.loaded || (fine && !nf->loaded)) return fail("bwd: networks not loaded");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* wbase = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(a->workspace), 1024));
  const size_t slack = wbase - static_cast<uint8_t*>(a->workspace);
  TrainWS t = carve_train(wbase, a->n_rays, S_c, S_f, nc, nf);
  if (t.total + slack > a->workspace_bytes) return fail("bwd: workspace too small");
  const int64_t n = a->n_rays;
  const int white = (a->flags & MOFA_FLAG_WHITE_BKGD) ? 1 : 0;
  // loss scale: a host value, or (loss_scale_dev) two floats in device memory {scale, 1 / scale} computed by the caller
  // on the device — then nothing on this path ever waits for the GPU
  const float* sdev = a->loss_scale_dev;
  const float scale = sdev ? 1.0f : (a->loss_scale > 0.0f ? a->loss_scale : 1.0f);
  const float inv = 1.0f / scale;
  const float* idev = sdev ? sdev + 1 : nullptr;
  float* d_lat[4] = {nullptr, a->d_expmod, a->d_shape, a->d_tex};
  CK(cudaMemsetAsync(a->d_rays, 0, sizeof(float) * 11 * n, s));
  CK(cudaMemsetAsync(a->d_shape, 0, sizeof(float) * kNShape, s));
  CK(cudaMemsetAsync(a->d_expmod, 0, sizeof(float) * kNExp, s));
  CK(cudaMemsetAsync(a->d_tex, 0, sizeof(float) * kNTex, s));
  // pass 1 = fine (rgb_map / acc_map), pass 0 = coarse (rgb0 / acc0, or the final maps when no fine pass ran)
  for (int ps = fine ? 1 : 0; ps >= 0; --ps) {
    const float* g_rgb = (ps == 1 || !fine) ? a->d_rgb : a->d_rgb0;
    const float* g_acc = (ps == 1 || !fine) ? a->d_acc : a->d_acc0;
    if (!g_rgb && !g_acc) continue;
    Net& net = ps == 1 ? *nf : nc;
    const int S = ps == 1 ? S_f : S_c;
    const float* noise = ps == 1 ? a->noise_f : a->noise_c;
    CK(launch_composite_bwd(t.pass[ps].raw, t.pass[ps].z, a->rays, a->ray_stride, noise, g_rgb, g_acc, scale, n, S, white,
                            t.d_raw, a->d_rays, s, sdev));
    c->launches++;
    {
      const int64_t rows = n * S, pad = (rows + 127) / 128 * 128 - rows;
      if (pad > 0) CK(cudaMemsetAsync(t.d_raw + rows * 4, 0, sizeof(float) * 4 * pad, s));   // rank-1 sigma term of padding rows
    }
    float* const* dp = (&net == &nc) ? a->d_params_coarse : a->d_params_fine;
    if (dp && (&net == &nc ? a->n_params_coarse : a->n_params_fine) != 2 * (n_dense_steps(net) + 2))
      return fail("bwd: d_params has the wrong number of tensors");
    if (run_backward(c, net, t, t.pass[ps], n * S, inv, d_lat, dp, s, idev)) return 1;
    CK(launch_pe_bwd(a->rays, a->ray_stride, t.pass[ps].z, t.dX0, t.dV, 128, n, S, a->d_rays, s));
    c->launches++;
  }
  CK(launch_scale_f32(a->d_rays, inv, 11 * n, s, idev));
  c->launches++;
  return 0;
}

int mofa_b200_raw2outputs_bwd(mofa_b200_ctx* c, const float* raw, const float* z, const float* rays, int stride,
                              const float* noise, const float* d_rgb, const float* d_acc, int64_t n, int S,
                              int white_bkgd, float* d_raw, float* d_rays, void* stream) {
  if (!c) return fail("raw2outputs_bwd: ctx is NULL");
  if (S < 2 || S > 256 || stride < 6) return fail("raw2outputs_bwd: bad sizes");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaMemsetAsync(d_rays, 0, sizeof(float) * 11 * n, s));
  CK(launch_composite_bwd(raw, z, rays, stride, noise, d_rgb, d_acc, 1.0f, n, S, white_bkgd, d_raw, d_rays, s));
  c->launches++;
  return 0;
}

int mofa_b200_wgrad(mofa_b200_ctx* c, const void* A, int Mp, const void* B, int Kb, int n_valid, int64_t P, float scale,
                    float* C, int ldc, int use_simt, void* stream) {
  if (!c) return fail("wgrad: ctx is NULL");
  if (Mp % 128 != 0 || Kb % 64 != 0 || P % 64 != 0) return fail("wgrad: Mp%%128, Kb%%64, P%%64 required");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (use_simt) {
    CK(launch_wgrad_simt(static_cast<const __half*>(A), Mp, static_cast<const __half*>(B), Kb, P, Mp, n_valid, scale, C, ldc, s));
  } else {
    WgradLaunch W;
    memset(&W, 0, sizeof(W));
    if (make_tmap_2d(c, &W.tmA, A, (uint64_t)P, (uint64_t)Mp, (uint64_t)Mp, 64)) return 1;
    if (make_tmap_2d(c, &W.tmB, B, (uint64_t)P, (uint64_t)Kb, (uint64_t)Kb, 64)) return 1;
    W.C = C; W.ldc = ldc; W.n_valid = n_valid; W.scale = scale; W.scale_dev = nullptr; W.Mp = Mp;
    W.BN = (Kb % 256 == 0) ? 256 : 128;
    W.Np = (Kb + W.BN - 1) / W.BN * W.BN;
    W.P = P;
    CK(launch_wgrad_tc(W, c->num_sms, s));
  }
  c->launches++;
  return 0;
}

}  // extern "C"
