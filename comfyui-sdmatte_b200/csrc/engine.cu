// Engine: weight repack + execution plan for the SDMatte single-pass matte path.
//
// The op graph follows the reference exactly (citations into /root/reference):
//   SDMatte.forward                 src/modeling/SDMatte/meta_arch.py:127-261
//   CustomUNet.forward              src/utils/replace.py:379-549   (module tree built at :184-362)
//   diffusers internals             restated in SURVEY.md Appendix A (A.2 embeddings, A.3 UNet blocks, A.4 VAE)
// Data layout: every activation is NHWC fp16 in one caller-provided arena; tokens (B, L, C) alias NHWC.
// The CLIP text encoder of the reference is dead compute on this path (meta_arch.py:220-234 never reaches
// the UNet because use_encoder_hidden_states_list=[True]*3, replace.py:414-416) and is not executed.
#include "engine.h"

#include "common.cuh"
#include "kernels.h"
#include "umma_gemm.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace sdm {

// ================================================================================================
// weights
// ================================================================================================
struct HostT {
  int dtype = 0;
  std::vector<int64_t> shape;
  const void* data = nullptr;
  bool used = false;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

static inline float bf16_to_f(uint16_t v) {
  uint32_t u = (uint32_t)v << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

struct Weights {
  std::unordered_map<std::string, HostT> host;  // valid during engine_load only
  std::unordered_map<std::string, void*> dev;   // packed device buffers
  std::vector<void*> allocs;
  std::vector<std::string> missing;
  bool loading = false;
  bool loaded = false;
  size_t bytes = 0;
  int n_used = 0, n_unexpected = 0;

  const HostT* find(const std::string& name) {
    auto it = host.find(name);
    if (it != host.end()) { it->second.used = true; return &it->second; }
    // legacy VAE attention names (query/key/value/proj_attn)
    static const char* alias[4][2] = {{".to_q.", ".query."}, {".to_k.", ".key."}, {".to_v.", ".value."}, {".to_out.0.", ".proj_attn."}};
    for (auto& a : alias) {
      auto pos = name.find(a[0]);
      if (pos != std::string::npos) {
        std::string alt = name;
        alt.replace(pos, strlen(a[0]), a[1]);
        auto it2 = host.find(alt);
        if (it2 != host.end()) { it2->second.used = true; return &it2->second; }
      }
    }
    return nullptr;
  }
  // fetch as fp32 host vector; records a missing key (and returns zeros) if absent
  std::vector<float> fetch(const std::string& name, int64_t expect_numel) {
    std::vector<float> out((size_t)expect_numel, 0.f);
    const HostT* t = find(name);
    if (!t) { missing.push_back(name); return out; }
    if (t->numel() != expect_numel)
      throw Error{"shape mismatch for '" + name + "': got " + std::to_string(t->numel()) + " elements, expected " + std::to_string(expect_numel)};
    if (t->dtype == 0) memcpy(out.data(), t->data, (size_t)expect_numel * 4);
    else if (t->dtype == 1) { const __half* h = (const __half*)t->data; for (int64_t i = 0; i < expect_numel; ++i) out[i] = __half2float(h[i]); }
    else if (t->dtype == 2) { const uint16_t* h = (const uint16_t*)t->data; for (int64_t i = 0; i < expect_numel; ++i) out[i] = bf16_to_f(h[i]); }
    else throw Error{"unsupported dtype for '" + name + "'"};
    return out;
  }
  void* upload(const std::string& key, const void* src, size_t nbytes) {
    void* d = nullptr;
    SDM_CUDA_OK(cudaMalloc(&d, std::max<size_t>(nbytes, 16)));
    SDM_CUDA_OK(cudaMemcpy(d, src, nbytes, cudaMemcpyHostToDevice));
    dev[key] = d;
    allocs.push_back(d);
    bytes += nbytes;
    return d;
  }
  void* get(const std::string& key) {
    auto it = dev.find(key);
    return it == dev.end() ? nullptr : it->second;
  }
  void require_loading(const std::string& key) {
    if (!loading) throw Error{"weight '" + key + "' was not packed at load time (internal error)"};
  }

  // conv weight OIHW -> [O(+pad)][kh*kw][I(+pad)] fp16
  const __half* conv(const std::string& name, int O, int I, int k, int Ipad = 0, int Opad = 0) {
    const std::string key = "c:" + name;
    if (void* p = get(key)) return (const __half*)p;
    require_loading(key);
    const int Ip = Ipad ? Ipad : I, Op = Opad ? Opad : O;
    std::vector<float> w = fetch(name + ".weight", (int64_t)O * I * k * k);
    std::vector<__half> pk((size_t)Op * k * k * Ip, __float2half_rn(0.f));
    for (int o = 0; o < O; ++o)
      for (int i = 0; i < I; ++i)
        for (int t = 0; t < k * k; ++t) pk[((size_t)o * k * k + t) * Ip + i] = __float2half_rn(w[((size_t)o * I + i) * k * k + t]);
    return (const __half*)upload(key, pk.data(), pk.size() * 2);
  }
  // "nearest x2 upsample -> 3x3 conv" as four polyphase 2x2-tap convs over the LOW-resolution input (Upsample2D of diffusers:
  // F.interpolate(scale 2, nearest) then conv; reference call sites: the decoder / UNet up blocks built at
  // /root/reference/src/utils/replace.py via diffusers).  Output pixel (2y + py, 2x + px) sees, through the nearest upsampling, the
  // low-resolution rows {y - 1, y, y} (py = 0) or {y, y, y + 1} (py = 1) under kernel rows ky = 0, 1, 2 — the taps that land on the same
  // input pixel are summed: 4 products per input channel instead of 9, and the upsampled tensor is never materialised.  The sums are
  // formed in fp32 from the fp16-rounded taps and rounded to fp16 once (relative error <= 2^-11 per combined weight, random sign:
  // below the fp16 rounding of the activations they multiply).  Layout: [parity q = 2 py + px][O][tap t = 2 dy + dx][I].
  const __half* conv_poly(const std::string& name, int O, int I) {
    const std::string key = "p:" + name;
    if (void* p = get(key)) return (const __half*)p;
    require_loading(key);
    std::vector<float> w = fetch(name + ".weight", (int64_t)O * I * 9);
    std::vector<__half> pk((size_t)4 * O * 4 * I);
    // kernel rows (columns) feeding low-resolution offset d of parity p: p = 0: d = 0 <- {0}, d = 1 <- {1, 2}; p = 1: d = 0 <- {0, 1}, d = 1 <- {2}
    auto lo = [](int par, int d) { return par == 0 ? (d == 0 ? 0 : 1) : (d == 0 ? 0 : 2); };
    auto hi = [](int par, int d) { return par == 0 ? (d == 0 ? 0 : 2) : (d == 0 ? 1 : 2); };
    for (int q = 0; q < 4; ++q) {
      const int py = q >> 1, px = q & 1;
      for (int o = 0; o < O; ++o)
        for (int t = 0; t < 4; ++t) {
          const int dy = t >> 1, dx = t & 1;
          for (int i = 0; i < I; ++i) {
            float acc = 0.f;
            for (int ky = lo(py, dy); ky <= hi(py, dy); ++ky)
              for (int kx = lo(px, dx); kx <= hi(px, dx); ++kx)
                acc += __half2float(__float2half_rn(w[((size_t)o * I + i) * 9 + ky * 3 + kx]));
            pk[(((size_t)q * O + o) * 4 + t) * I + i] = __float2half_rn(acc);
          }
        }
    }
    return (const __half*)upload(key, pk.data(), pk.size() * 2);
  }
  // 3x3 conv with O <= 3 output channels as a 1x1 GEMM producing the 9*O per-tap partial products of every INPUT pixel
  // (rows tap*O + o, padded to 32 rows); alpha_col2im_kernel then sums the 9 shifted partials per output pixel
  const __half* conv_taprows(const std::string& name, int O, int I) {
    const std::string key = "t:" + name;
    if (void* p = get(key)) return (const __half*)p;
    require_loading(key);
    std::vector<float> w = fetch(name + ".weight", (int64_t)O * I * 9);
    std::vector<__half> pk((size_t)32 * I, __float2half_rn(0.f));
    for (int o = 0; o < O; ++o)
      for (int i = 0; i < I; ++i)
        for (int t = 0; t < 9; ++t) pk[(size_t)(t * O + o) * I + i] = __float2half_rn(w[((size_t)o * I + i) * 9 + t]);
    return (const __half*)upload(key, pk.data(), pk.size() * 2);
  }
  // 3x3 conv with I <= 4 input channels as a GEMM over im2col rows: [O][64], k = tap*4 + ci
  const __half* conv_im2col(const std::string& name, int O, int I) {
    const std::string key = "i:" + name;
    if (void* p = get(key)) return (const __half*)p;
    require_loading(key);
    std::vector<float> w = fetch(name + ".weight", (int64_t)O * I * 9);
    std::vector<__half> pk((size_t)O * 64, __float2half_rn(0.f));
    for (int o = 0; o < O; ++o)
      for (int i = 0; i < I; ++i)
        for (int t = 0; t < 9; ++t) pk[(size_t)o * 64 + t * 4 + i] = __float2half_rn(w[((size_t)o * I + i) * 9 + t]);
    return (const __half*)upload(key, pk.data(), pk.size() * 2);
  }
  const __half* linear(const std::string& name, int N, int K) {
    const std::string key = "l:" + name;
    if (void* p = get(key)) return (const __half*)p;
    require_loading(key);
    std::vector<float> w = fetch(name + ".weight", (int64_t)N * K);
    std::vector<__half> pk(w.size());
    for (size_t i = 0; i < w.size(); ++i) pk[i] = __float2half_rn(w[i]);
    return (const __half*)upload(key, pk.data(), pk.size() * 2);
  }
  const float* vec(const std::string& name, int n, int npad = 0) {
    const std::string key = "v:" + name;
    if (void* p = get(key)) return (const float*)p;
    require_loading(key);
    std::vector<float> v = fetch(name, n);
    if (npad > n) v.resize(npad, 0.f);
    return (const float*)upload(key, v.data(), v.size() * 4);
  }
  // GEGLU projection [8C][C]: rows re-ordered so that every 256-row tile holds 128 value rows followed by the
  // 128 matching gate rows (reference FeedForward/GEGLU: value = first half, gate = second half)
  const __half* geglu_w(const std::string& name, int C) {
    const std::string key = "g:" + name;
    if (void* p = get(key)) return (const __half*)p;
    require_loading(key);
    const int F = 4 * C;
    std::vector<float> w = fetch(name + ".weight", (int64_t)2 * F * C);
    std::vector<__half> pk(w.size());
    for (int t = 0; t < F / 128; ++t)
      for (int half = 0; half < 2; ++half)
        for (int r = 0; r < 128; ++r) {
          const size_t src = (size_t)(half * F + t * 128 + r) * C, dst = (size_t)(t * 256 + half * 128 + r) * C;
          for (int c = 0; c < C; ++c) pk[dst + c] = __float2half_rn(w[src + c]);
        }
    return (const __half*)upload(key, pk.data(), pk.size() * 2);
  }
  const float* geglu_b(const std::string& name, int C) {
    const std::string key = "gb:" + name;
    if (void* p = get(key)) return (const float*)p;
    require_loading(key);
    const int F = 4 * C;
    std::vector<float> b = fetch(name + ".bias", 2 * F), pk(2 * F);
    for (int t = 0; t < F / 128; ++t)
      for (int half = 0; half < 2; ++half)
        for (int r = 0; r < 128; ++r) pk[t * 256 + half * 128 + r] = b[half * F + t * 128 + r];
    return (const float*)upload(key, pk.data(), pk.size() * 4);
  }
  const float* raw_vec(const std::string& key, const std::vector<float>& v) {
    if (void* p = get(key)) return (const float*)p;
    require_loading(key);
    return (const float*)upload(key, v.data(), v.size() * 4);
  }
  void free_all() {
    for (void* p : allocs) cudaFree(p);
    allocs.clear();
    dev.clear();
    bytes = 0;
  }
};

// ================================================================================================
// plan-time arena: offsets inside the caller's workspace, first-fit free list, explicit free
// ================================================================================================
struct Arena {
  struct Blk { size_t off, size; };
  std::vector<Blk> free_list;
  size_t top = 0, peak = 0;
  static size_t align(size_t x) { return (x + 1023) & ~(size_t)1023; }
  size_t alloc(size_t bytes) {
    bytes = align(std::max<size_t>(bytes, 1024));
    size_t best = (size_t)-1;
    for (size_t i = 0; i < free_list.size(); ++i)
      if (free_list[i].size >= bytes && (best == (size_t)-1 || free_list[i].size < free_list[best].size)) best = i;
    if (best != (size_t)-1) {
      const size_t off = free_list[best].off;
      if (free_list[best].size == bytes) free_list.erase(free_list.begin() + best);
      else { free_list[best].off += bytes; free_list[best].size -= bytes; }
      return off;
    }
    // extend the top (absorb a trailing free block if it touches the top)
    for (size_t i = 0; i < free_list.size(); ++i)
      if (free_list[i].off + free_list[i].size == top) {
        const size_t off = free_list[i].off;
        top = off + bytes;
        free_list.erase(free_list.begin() + i);
        peak = std::max(peak, top);
        return off;
      }
    const size_t off = top;
    top += bytes;
    peak = std::max(peak, top);
    return off;
  }
  void release(size_t off, size_t bytes) {
    bytes = align(std::max<size_t>(bytes, 1024));
    free_list.push_back({off, bytes});
    std::sort(free_list.begin(), free_list.end(), [](const Blk& a, const Blk& b) { return a.off < b.off; });
    for (size_t i = 0; i + 1 < free_list.size();) {
      if (free_list[i].off + free_list[i].size == free_list[i + 1].off) {
        free_list[i].size += free_list[i + 1].size;
        free_list.erase(free_list.begin() + i + 1);
      } else ++i;
    }
  }
};

struct T {  // NHWC fp16 activation
  __half* p = nullptr;
  size_t off = 0, bytes = 0;
  int B = 0, H = 0, W = 0, C = 0;
  // GroupNorm partial statistics written by the producing conv's epilogue (optional)
  float* stats = nullptr;
  size_t stats_off = 0, stats_bytes = 0;
  int stat_slots = 0;
  long long ld() const { return C; }
  long long HW() const { return (long long)H * W; }
  bool valid() const { return bytes != 0; }
};

using Op = std::function<void(cudaStream_t)>;
struct OpRec {
  Op fn;
  std::string kind;
  double flops = 0;   // algorithmic tensor FLOPs (0 for HBM-bound ops)
  double bytes = 0;   // algorithmic HBM bytes (inputs read once + outputs written once)
};
struct ProfRec {
  std::string kind;
  float ms;
  double flops, bytes;
};

struct Plan {
  int B = 0, R = 0;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  std::vector<OpRec> ops;
  int n_launches = 0;
  double tensor_flops = 0;
  // fixed buffers
  float* d_image = nullptr;   // used by forward_host only
  float* d_trimap = nullptr;
  __half* d_alpha = nullptr;
  const float** image_slot = nullptr;
  std::map<std::string, T> taps;
  std::vector<std::string> tap_order;
  bool keep_taps = false;
  // CUDA graph of the op list (engine option "cuda_graph"): kernel arguments are baked at capture, so a graph is valid for
  // exactly one set of run-time pointers (`graph_slots`) and one set of per-sample flags
  cudaGraphExec_t graph_exec = nullptr;
  int eager_runs = 0;
  // per-sample conditioning (point / bbox / mask prompts): plan-owned device tables, coordinates buffer in the workspace
  int prompt_kind = -1, prompt_ncoords = 0;
  int4* d_row_map = nullptr;
  int* d_iota = nullptr;
  float* d_coords = nullptr;
  ~Plan() {
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    if (d_row_map) cudaFree(d_row_map);
    if (d_iota) cudaFree(d_iota);
  }
  // run-time argument slots (device pointers change per call without rebuilding the plan)
  struct Slots { const float* image; const float* trimap; __half* alpha; __half* premean; } slots{}, graph_slots{}, last_slots{};
  int* d_is_trans = nullptr;
};

struct Engine {
  int device = 0;
  int num_sms = 148;
  bool keep_taps = false;  // parity diagnostics: tapped block outputs are never recycled by the arena (bigger workspace)
  int prompt_kind = -1;    // -1: the node's trimap prompt (embeddings folded at load); 0 bbox-like (4 coords), 1 point coords (n3)
  int prompt_ncoords = 0;
  std::vector<std::pair<std::string, int>> temb_order;  // resnets in the order of the stacked time_emb_proj rows, with their C
  int temb_rows = 0;
  bool has_point_embedding = false;
  bool use_graph = true;   // replay the plan as one CUDA graph when the caller's pointers repeat (option "cuda_graph")
  cudaStream_t cap_stream = nullptr;  // capture happens here (the caller's stream may be the legacy default stream, which cannot capture)
  int32_t* flags_pinned = nullptr;    // staging for the per-sample is_transparent flags (uploaded only when they change)
  char* pinned = nullptr;             // page-locked staging of the node call's host tensors (grown on demand, engine_apply_host)
  size_t pinned_bytes = 0;
  int copy_threads = 8;
  double node_ms[4] = {0, 0, 0, 0};   // last node call: host staging + H2D enqueue, kernel enqueue, wait for the GPU, copy-out
  cudaEvent_t flags_event = nullptr;
  int graph_launches = 0, graph_captures = 0;
  Weights W;
  std::unique_ptr<Plan> plan;
  int last_launches = 0;
  double last_flops = 0;
  std::vector<ProfRec> prof;
};

// ================================================================================================
// graph builder
// ================================================================================================
struct Builder {
  Engine& E;
  Weights& W;
  Plan* plan;       // null in dry mode
  Arena arena;
  bool dry;
  int B, R, S;
  char* ws;
  int n_launches = 0;
  double flops = 0;
  int* d_is_trans = nullptr;
  // per-sample conditioning (E.prompt_kind >= 0): bias table of every resnet, [B][C] each, selected per sample by d_iota
  float* dyn_tables = nullptr;
  const int* d_iota = nullptr;
  std::map<std::string, size_t> dyn_off;  // resnet -> float offset of its table

  Builder(Engine& e, Plan* p, int B_, int R_, void* ws_) : E(e), W(e.W), plan(p), dry(p == nullptr), B(B_), R(R_), S(R_ / 8), ws((char*)ws_) {}

  T alloc(int b, int h, int w, int c) {
    T t;
    t.B = b; t.H = h; t.W = w; t.C = c;
    t.bytes = (size_t)b * h * w * c * 2;
    t.off = arena.alloc(t.bytes);
    t.p = dry ? nullptr : (__half*)(ws + t.off);
    return t;
  }
  void* alloc_raw(size_t bytes, size_t* off_out) {
    const size_t off = arena.alloc(bytes);
    *off_out = off;
    return dry ? nullptr : (void*)(ws + off);
  }
  // parity taps: the named block output stays readable after the forward (sdm_debug_tensor).  Without the engine's keep_taps
  // option only tensors that are never recycled anyway (unet_in, ctx, unet_out_scaled) are valid afterwards.
  std::vector<size_t> pinned;
  void tap(const std::string& name, const T& t) {
    if (E.keep_taps) pinned.push_back(t.off);
    if (plan && (E.keep_taps || name == "unet_in" || name == "ctx" || name == "unet_out_scaled")) {
      plan->taps[name] = t;
      plan->tap_order.push_back(name);
    }
  }
  void free(T& t) {
    const bool keep = t.valid() && std::find(pinned.begin(), pinned.end(), t.off) != pinned.end();
    if (t.valid() && !keep) arena.release(t.off, t.bytes);
    if (t.stats_bytes) arena.release(t.stats_off, t.stats_bytes);
    t.bytes = 0;
    t.stats_bytes = 0;
  }
  void push(Op op, int launches = 1, const std::string& kind = "misc", double fl = 0, double by = 0) {
    n_launches += launches;
    if (!dry) plan->ops.push_back(OpRec{std::move(op), kind, fl, by});
  }

  // ---------------------------------------------------------------- primitive emitters
  struct GemmOpt {
    const float* bias = nullptr;
    const int* bias_sel = nullptr;
    const T* res = nullptr;
    int mode = EPI_F16;
    int ups2 = 0;
    int stride = 1;
    int pad = PAD_SAME;
    float scale = 1.0f;
    float post_div = 1.0f;
    int n_store = 0;
    T* stats_for = nullptr;    // EPI_F16 (no ups2): attach GroupNorm partial statistics of the output to this tensor
    const float* gn_ab = nullptr;  // fused GroupNorm(+SiLU) of the input (the conv normalises its resident halo tile itself)
    int gn_silu = 0;
    const char* label = nullptr;
  };
  // generic tensor-core conv (ksize 1/3) over one or two channel-concatenated sources
  void conv_tc(const T& a, const T* a2, const __half* w, int N, int ksize, const T& out, const GemmOpt& o) {
    const int Hout = a.H / o.stride, Wout = a.W / o.stride;
    const double fl = 2.0 * a.B * Hout * Wout * (double)N * ksize * ksize * (a.C + (a2 ? a2->C : 0));
    flops += fl;
    float* stats_ptr = nullptr;
    if (o.stats_for && o.mode == EPI_F16 && !o.ups2) {
      T& st_t = *o.stats_for;
      st_t.stat_slots = conv_gemm_tiles_per_image(Hout, Wout);
      st_t.stats_bytes = (size_t)a.B * st_t.stat_slots * N * 2 * sizeof(float);
      stats_ptr = (float*)alloc_raw(st_t.stats_bytes, &st_t.stats_off);
      st_t.stats = dry ? (float*)1 : stats_ptr;  // non-null marker in dry mode
    }
    if (dry) { n_launches++; return; }
    ConvGemmDesc d;
    d.B = a.B; d.Hin = a.H; d.Win = a.W;
    d.nsrc = a2 ? 2 : 1;
    d.src[0] = {a.p, a.C, a.ld()};
    if (a2) d.src[1] = {a2->p, a2->C, a2->ld()};
    d.ksize = ksize; d.stride = o.stride; d.pad = o.pad;
    d.w = w; d.N = N;
    d.mode = o.mode; d.ups2 = o.ups2;
    d.out = out.p;
    if (o.mode == EPI_F16_T) {
      d.out_ld = out.W;                          // out is [B][N][Lld] stored as T{B, H=N, W=Lld, C=1}
      d.out_bstride = (long long)out.H * out.W;
    } else {
      d.out_ld = out.C;
      d.out_bstride = out.HW() * out.C;
    }
    d.bias = o.bias; d.bias_sel = o.bias_sel;
    if (o.res) { d.res = o.res->p; d.res_ld = o.res->C; d.res_bstride = o.res->HW() * o.res->C; }
    d.scale = o.scale;
    d.post_div = o.post_div;
    d.n_store = o.n_store;
    d.stats = stats_ptr;
    d.gn_ab = o.gn_ab; d.gn_silu = o.gn_silu;
    if (o.mode == EPI_ALPHA) { d.out_ld = 1; d.out_bstride = (long long)Hout * Wout; }
    auto l = conv_gemm_build(d, E.num_sms);
    const double by = 2.0 * a.B * ((double)a.HW() * (a.C + (a2 ? a2->C : 0)) + (double)Hout * Wout * N * (o.ups2 ? 4 : 1) * (o.mode == EPI_GEGLU ? 0.5 : 1.0) +
                                   (o.res ? (double)Hout * Wout * N : 0.0)) + 2.0 * N * ksize * ksize * (a.C + (a2 ? a2->C : 0));
    std::string kind = o.label ? o.label : (ksize == 3 ? (o.stride == 2 ? "conv3x3_s2" : "conv3x3") : (a.H == 1 ? "linear" : "conv1x1"));
    if (o.mode == EPI_GEGLU) kind = "linear_geglu";
    if (o.mode == EPI_F16_T) kind = "linear_vT";
    push([l](cudaStream_t st) { conv_gemm_run(*l, st); }, 1, "tc:" + kind, fl, by);
  }
  // token GEMM: x [B][L][K] -> out [B][L][N]
  void linear(const T& x, const __half* w, int N, const T& out, const GemmOpt& o) {
    T xv = x; xv.H = 1; xv.W = (int)x.HW();
    T ov = out;
    if (o.mode != EPI_F16_T) { ov.H = 1; ov.W = (int)out.HW(); }
    GemmOpt oo = o;
    T rv;
    if (o.res) { rv = *o.res; rv.H = 1; rv.W = (int)o.res->HW(); oo.res = &rv; }
    conv_tc(xv, nullptr, w, N, 1, ov, oo);
  }
  T groupnorm(const T& a, const T* a2, const std::string& name, float eps, int silu) {
    const int Ctot = a.C + (a2 ? a2->C : 0);
    T out = alloc(a.B, a.H, a.W, Ctot);
    const float* gamma = W.vec(name + ".weight", Ctot);
    const float* beta = W.vec(name + ".bias", Ctot);
    size_t soff;
    const size_t sbytes = groupnorm_scratch_floats(a.B, (int)a.HW(), Ctot) * 4;
    float* scratch = (float*)alloc_raw(sbytes, &soff);
    if (!dry) {
      GroupNormDesc d;
      d.B = a.B; d.HW = (int)a.HW(); d.nsrc = a2 ? 2 : 1;
      d.src[0] = a.p; d.C[0] = a.C; d.ld[0] = a.ld();
      if (a2) { d.src[1] = a2->p; d.C[1] = a2->C; d.ld[1] = a2->ld(); }
      d.gamma = gamma; d.beta = beta; d.eps = eps; d.silu = silu;
      d.out = out.p; d.scratch = scratch;
      if (a.stats && (!a2 || (a2->stats && a2->stat_slots == a.stat_slots))) {
        d.pre_partial[0] = a.stats;
        d.pre_partial[1] = a2 ? a2->stats : nullptr;
        d.pre_slots = a.stat_slots;
      }
      push([d](cudaStream_t st) { groupnorm_run(d, st); }, groupnorm_num_launches(d), "groupnorm", 0, 4.0 * a.B * (double)a.HW() * Ctot);
    } else n_launches += 3;
    arena.release(soff, sbytes);
    return out;
  }
  // GroupNorm whose apply pass is fused into the consuming 3x3 conv (conv_gemm_can_fuse_gn): statistics + finalize only.
  // The (scale, shift) table lives in the scratch block, which must stay allocated until the consuming conv has been emitted.
  struct GnFused { const float* ab = nullptr; size_t soff = 0, sbytes = 0; };
  GnFused groupnorm_stats_only(const T& a, const T* a2, const std::string& name, float eps) {
    const int Ctot = a.C + (a2 ? a2->C : 0);
    const float* gamma = W.vec(name + ".weight", Ctot);
    const float* beta = W.vec(name + ".bias", Ctot);
    GnFused g;
    g.sbytes = groupnorm_scratch_floats(a.B, (int)a.HW(), Ctot) * 4;
    float* scratch = (float*)alloc_raw(g.sbytes, &g.soff);
    if (!dry) {
      GroupNormDesc d;
      d.B = a.B; d.HW = (int)a.HW(); d.nsrc = a2 ? 2 : 1;
      d.src[0] = a.p; d.C[0] = a.C; d.ld[0] = a.ld();
      if (a2) { d.src[1] = a2->p; d.C[1] = a2->C; d.ld[1] = a2->ld(); }
      d.gamma = gamma; d.beta = beta; d.eps = eps; d.silu = 1;
      d.out = nullptr; d.scratch = scratch;
      if (a.stats && (!a2 || (a2->stats && a2->stat_slots == a.stat_slots))) {
        d.pre_partial[0] = a.stats;
        d.pre_partial[1] = a2 ? a2->stats : nullptr;
        d.pre_slots = a.stat_slots;
      }
      g.ab = groupnorm_ab(d);
      push([d](cudaStream_t st) { groupnorm_run(d, st); }, groupnorm_num_launches(d), "groupnorm_stats", 0, 0);
    } else n_launches += 2;
    return g;
  }
  T layernorm(const T& x, const std::string& name) {
    T out = alloc(x.B, x.H, x.W, x.C);
    const float* g = W.vec(name + ".weight", x.C);
    const float* b = W.vec(name + ".bias", x.C);
    if (!dry) {
      const __half* xp = x.p; __half* yp = out.p;
      const long long rows = (long long)x.B * x.HW();
      const int C = x.C;
      push([=](cudaStream_t st) { layernorm_run(xp, yp, g, b, rows, C, 1e-5f, st); }, 1, "layernorm", 0, 4.0 * rows * C);
    } else n_launches++;
    return out;
  }
  void direct(const DirectConvDesc& d) {
    if (!dry) push([d](cudaStream_t st) { direct_conv_run(d, st); }, 1, "direct_conv", 0, 2.0 * d.B * (double)d.H * d.W * (d.Cin + (d.cout_limit ? d.cout_limit : d.Cout)));
    else n_launches++;
  }
  void attention(const T& q, const T& k, const T& vt, int heads, const float* bias, long long bias_bs, const int* ntiles, const T& out) {
    const int Lq = (int)q.HW(), Lk = (int)k.HW();
    const double fl = 4.0 * q.B * heads * (double)Lq * Lk * 64;
    flops += fl;
    if (dry) { n_launches++; return; }
    AttnDesc d;
    d.B = q.B; d.heads = heads; d.Lq = Lq; d.Lk = Lk;
    d.q = q.p; d.ldq = q.C; d.k = k.p; d.ldk = k.C;
    d.vt = vt.p; d.ldvt = vt.W;
    d.bias = bias; d.bias_bstride = bias_bs; d.ntiles = ntiles;
    d.out = out.p; d.ldo = out.C; d.scale = 0.125f;
    auto l = attn_build(d);
    push([l](cudaStream_t st) { attn_run(*l, st); }, 1, bias ? "tc:attention_self" : "tc:attention_cross", fl, 2.0 * q.B * heads * 64.0 * (2.0 * Lq + 2.0 * Lk));
  }

  // ---------------------------------------------------------------- blocks
  // ResnetBlock2D.  x2 != null: input is cat([x, x2], channel).  temb_bias: folded
  // conv1.bias + time_emb_proj(silu(emb)) with two rows (is_transparent = 0 / 1), or null for the VAE.
  T resnet(const T& x, const T* x2, const std::string& p, int Cout, float eps, bool has_temb, int ups2) {
    const int Cin = x.C + (x2 ? x2->C : 0);
    // GroupNorm -> SiLU -> conv3x3: where the consuming conv can normalise its resident input tile (conv_gemm_can_fuse_gn, a
    // function of the per-sample geometry only) there is no apply pass and no normalised copy of the activation
    const bool fuse1 = conv_gemm_can_fuse_gn(3, 1, EPI_F16, 0, Cout, 0, x.H, x.W);
    const bool fuse2 = conv_gemm_can_fuse_gn(3, 1, EPI_F16, ups2, Cout, 1, x.H, x.W);
    T h = alloc(x.B, x.H, x.W, Cout);
    {
      GemmOpt o;
      if (has_temb && E.prompt_kind >= 0) { o.bias = dry ? nullptr : dyn_tables + dyn_off.at(p); o.bias_sel = d_iota; }
      else if (has_temb) { o.bias = W.raw_vec("temb:" + p, {}); o.bias_sel = d_is_trans; }
      else o.bias = W.vec(p + ".conv1.bias", Cout);
      o.stats_for = &h;  // norm2 statistics come out of this conv's epilogue
      if (fuse1) {
        GnFused g = groupnorm_stats_only(x, x2, p + ".norm1", eps);
        o.gn_ab = g.ab; o.gn_silu = 1;
        conv_tc(x, x2, W.conv(p + ".conv1", Cout, Cin, 3), Cout, 3, h, o);
        arena.release(g.soff, g.sbytes);
      } else {
        T n1 = groupnorm(x, x2, p + ".norm1", eps, 1);
        conv_tc(n1, nullptr, W.conv(p + ".conv1", Cout, Cin, 3), Cout, 3, h, o);
        free(n1);
      }
    }
    T n2;
    GnFused g2;
    if (fuse2) g2 = groupnorm_stats_only(h, nullptr, p + ".norm2", eps);
    else { n2 = groupnorm(h, nullptr, p + ".norm2", eps, 1); free(h); }
    T sc;
    const T* resid = &x;
    if (Cin != Cout) {
      sc = alloc(x.B, x.H, x.W, Cout);
      GemmOpt o;
      o.bias = W.vec(p + ".conv_shortcut.bias", Cout);
      conv_tc(x, x2, W.conv(p + ".conv_shortcut", Cout, Cin, 1), Cout, 1, sc, o);
      resid = &sc;
    } else {
      SDM_CHECK(x2 == nullptr, "concat input without shortcut conv");
    }
    T out = ups2 ? alloc(x.B, x.H * 2, x.W * 2, Cout) : alloc(x.B, x.H, x.W, Cout);
    {
      GemmOpt o;
      o.bias = W.vec(p + ".conv2.bias", Cout);
      o.res = resid;
      o.ups2 = ups2;
      o.stats_for = ups2 ? nullptr : &out;
      if (fuse2) { o.gn_ab = g2.ab; o.gn_silu = 1; }
      conv_tc(fuse2 ? h : n2, nullptr, W.conv(p + ".conv2", Cout, Cout, 3), Cout, 3, out, o);
    }
    if (fuse2) { arena.release(g2.soff, g2.sbytes); free(h); }
    else free(n2);
    free(sc);
    return out;
  }

  // Transformer2DModel with one BasicTransformerBlock; ctx = trimap tokens [B][Lctx][1024]
  // per-level key bias of attn1 (+ the compacted form, see key_compact_kernel; idx == null: compaction off)
  struct KeyBias { const float* bias = nullptr; long long bs = 0; const float* cbias = nullptr; const int* idx = nullptr; const int* ntiles = nullptr; };
  T transformer(const T& x, const std::string& p, int heads, const T& ctx, const KeyBias& kbias, int ups2) {
    const int C = x.C, L = (int)x.HW(), Bq = x.B;
    const std::string tb = p + ".transformer_blocks.0";
    T n = groupnorm(x, nullptr, p + ".norm", 1e-6f, 0);
    T h = alloc(Bq, x.H, x.W, C);
    { GemmOpt o; o.bias = W.vec(p + ".proj_in.bias", C); linear(n, W.linear(p + ".proj_in", C, C), C, h, o); }
    free(n);
    auto attn = [&](const std::string& ap, const T& qsrc, const T& kvsrc, int Ckv, const float* bias, long long bias_bs, const int* ntiles) {
      const int Lk = (int)kvsrc.HW();
      const int Lld = (Lk + 7) & ~7;
      T q = alloc(Bq, x.H, x.W, C);
      T k = alloc(Bq, kvsrc.H, kvsrc.W, C);
      T vt; vt.B = Bq; vt.H = C; vt.W = Lld; vt.C = 1; vt.bytes = (size_t)Bq * C * Lld * 2; vt.off = arena.alloc(vt.bytes);
      vt.p = dry ? nullptr : (__half*)(ws + vt.off);
      { GemmOpt o; linear(qsrc, W.linear(ap + ".to_q", C, C), C, q, o); }
      { GemmOpt o; linear(kvsrc, W.linear(ap + ".to_k", C, Ckv), C, k, o); }
      { GemmOpt o; o.mode = EPI_F16_T; linear(kvsrc, W.linear(ap + ".to_v", C, Ckv), C, vt, o); }
      T o_ = alloc(Bq, x.H, x.W, C);
      attention(q, k, vt, heads, bias, bias_bs, ntiles, o_);
      free(q); free(k); free(vt);
      { GemmOpt o; o.bias = W.vec(ap + ".to_out.0.bias", C); o.res = &h; linear(o_, W.linear(ap + ".to_out.0", C, C), C, h, o); }
      free(o_);
    };
    {  // attn1: self attention with the trimap key bias (replace.py:20-122)
      T ln = layernorm(h, tb + ".norm1");
      if (kbias.idx) {
        // keys/values only for the keys that can have a non-zero probability: gather their LayerNorm rows in front
        // (rows beyond 128*ntiles[b] of lnc, and so of K / V^T, are never read by the attention kernel)
        T lnc = alloc(Bq, x.H, x.W, C);
        if (!dry) {
          const __half* sp = ln.p; __half* dp = lnc.p; const int* ip = kbias.idx; const int* np = kbias.ntiles; const int bs = (int)kbias.bs;
          push([=](cudaStream_t st) { gather_rows_run(sp, dp, ip, np, Bq, L, C, bs, st); }, 1, "gather_keys", 0, 0);
        } else n_launches++;
        attn(tb + ".attn1", ln, lnc, C, kbias.cbias, kbias.bs, kbias.ntiles);
        free(lnc);
      } else {
        attn(tb + ".attn1", ln, ln, C, kbias.bias, kbias.bs, nullptr);
      }
      free(ln);
    }
    {  // attn2: cross attention to the trimap tokens, no mask (replace.py:96-98 beta=0)
      T ln = layernorm(h, tb + ".norm2");
      attn(tb + ".attn2", ln, ctx, 1024, nullptr, 0, nullptr);
      free(ln);
    }
    {  // feed-forward (GEGLU)
      T ln = layernorm(h, tb + ".norm3");
      T g = alloc(Bq, x.H, x.W, 4 * C);
      { GemmOpt o; o.mode = EPI_GEGLU; o.bias = W.geglu_b(tb + ".ff.net.0.proj", C); linear(ln, W.geglu_w(tb + ".ff.net.0.proj", C), 8 * C, g, o); }
      free(ln);
      { GemmOpt o; o.bias = W.vec(tb + ".ff.net.2.bias", C); o.res = &h; linear(g, W.linear(tb + ".ff.net.2", C, 4 * C), C, h, o); }
      free(g);
    }
    T out = ups2 ? alloc(Bq, x.H * 2, x.W * 2, C) : alloc(Bq, x.H, x.W, C);
    { GemmOpt o; o.bias = W.vec(p + ".proj_out.bias", C); o.res = &x; o.ups2 = ups2; o.stats_for = ups2 ? nullptr : &out;
      conv_tc(h, nullptr, W.linear(p + ".proj_out", C, C), C, 1, out, o); }
    free(h);
    (void)L;
    return out;
  }

  // VAE mid-block attention: single head, d = 512.  Unfused: QK^T -> fp32 scores -> softmax -> P V.
  T vae_attention(const T& x, const std::string& p) {
    const int C = 512, L = (int)x.HW(), Bv = x.B;
    T n = groupnorm(x, nullptr, p + ".group_norm", 1e-6f, 0);
    T q = alloc(Bv, x.H, x.W, C), k = alloc(Bv, x.H, x.W, C);
    T vt; vt.B = Bv; vt.H = C; vt.W = L; vt.C = 1; vt.bytes = (size_t)Bv * C * L * 2; vt.off = arena.alloc(vt.bytes);
    vt.p = dry ? nullptr : (__half*)(ws + vt.off);
    { GemmOpt o; o.bias = W.vec(p + ".to_q.bias", C); linear(n, W.linear(p + ".to_q", C, C), C, q, o); }
    { GemmOpt o; o.bias = W.vec(p + ".to_k.bias", C); linear(n, W.linear(p + ".to_k", C, C), C, k, o); }
    { GemmOpt o; o.bias = W.vec(p + ".to_v.bias", C); o.mode = EPI_F16_T; linear(n, W.linear(p + ".to_v", C, C), C, vt, o); }
    free(n);
    T att = alloc(Bv, x.H, x.W, C);
    const int chunk = std::max(1, std::min(Bv, (int)((6ull << 30) / ((size_t)L * L * 6))));
    size_t s_off, p_off;
    const size_t s_bytes = (size_t)chunk * L * L * 4, p_bytes = (size_t)chunk * L * L * 2;
    float* scores = (float*)alloc_raw(s_bytes, &s_off);
    __half* probs = (__half*)alloc_raw(p_bytes, &p_off);
    for (int b0 = 0; b0 < Bv; b0 += chunk) {
      const int nb = std::min(chunk, Bv - b0);
      flops += 4.0 * nb * (double)L * L * C;
      if (dry) { n_launches += 3; continue; }
      ConvGemmDesc d1;  // scores[b] = scale * Q[b] K[b]^T
      d1.B = nb; d1.Hin = 1; d1.Win = L; d1.nsrc = 1;
      d1.src[0] = {q.p + (size_t)b0 * L * C, C, C};
      d1.ksize = 1; d1.w = k.p + (size_t)b0 * L * C; d1.N = L; d1.w_bstride = (long long)L * C;
      d1.mode = EPI_F32; d1.out = scores; d1.out_ld = L; d1.out_bstride = (long long)L * L;
      d1.scale = 1.0f / sqrtf((float)C);
      auto l1 = conv_gemm_build(d1, E.num_sms);
      push([l1](cudaStream_t st) { conv_gemm_run(*l1, st); }, 1, "tc:vae_qk", 2.0 * nb * (double)L * L * C, nb * ((double)L * L * 4 + 4.0 * L * C));
      const long long rows = (long long)nb * L;
      push([=](cudaStream_t st) { softmax_rows_run(scores, probs, rows, L, st); }, 1, "softmax_rows", 0, 6.0 * rows * L);
      ConvGemmDesc d2;  // att[b] = P[b] V[b]   (B operand = V^T [512][L])
      d2.B = nb; d2.Hin = 1; d2.Win = L; d2.nsrc = 1;
      d2.src[0] = {probs, L, L};
      d2.ksize = 1; d2.w = vt.p + (size_t)b0 * C * L; d2.N = C; d2.w_bstride = (long long)C * L;
      d2.mode = EPI_F16; d2.out = att.p + (size_t)b0 * L * C; d2.out_ld = C; d2.out_bstride = (long long)L * C;
      auto l2 = conv_gemm_build(d2, E.num_sms);
      push([l2](cudaStream_t st) { conv_gemm_run(*l2, st); }, 1, "tc:vae_pv", 2.0 * nb * (double)L * L * C, nb * ((double)L * L * 2 + 4.0 * L * C));
    }
    arena.release(s_off, s_bytes);
    arena.release(p_off, p_bytes);
    free(q); free(k); free(vt);
    T out = alloc(Bv, x.H, x.W, C);
    { GemmOpt o; o.bias = W.vec(p + ".to_out.0.bias", C); o.res = &x; o.stats_for = &out; linear(att, W.linear(p + ".to_out.0", C, C), C, out, o); }
    free(att);
    return out;
  }

  T conv3_plain(const T& x, const std::string& name, int Cout, int stride, int pad) {
    T out = alloc(x.B, x.H / stride, x.W / stride, Cout);
    GemmOpt o;
    o.bias = W.vec(name + ".bias", Cout);
    o.stride = stride; o.pad = pad;
    o.stats_for = &out;
    conv_tc(x, nullptr, W.conv(name, Cout, x.C, 3), Cout, 3, out, o);
    return out;
  }

  // Upsample2D (nearest x2 + conv3x3) of a LOW-resolution x as four polyphase launches (Weights::conv_poly); the caller checked
  // conv_gemm_can_poly(Cout, x.H, x.W).  The output carries its GroupNorm partials (4 x the low-resolution slot count).
  T conv3_poly(const T& x, const std::string& name, int Cout) {
    T out = alloc(x.B, x.H * 2, x.W * 2, Cout);
    const __half* w = W.conv_poly(name, Cout, x.C);
    const float* bias = W.vec(name + ".bias", Cout);
    out.stat_slots = 4 * conv_gemm_tiles_per_image(x.H, x.W);
    out.stats_bytes = (size_t)x.B * out.stat_slots * Cout * 2 * sizeof(float);
    float* stats_ptr = (float*)alloc_raw(out.stats_bytes, &out.stats_off);
    out.stats = dry ? (float*)1 : stats_ptr;
    for (int q = 0; q < 4; ++q) {
      const double fl = 2.0 * x.B * (double)x.HW() * Cout * 4.0 * x.C;
      flops += fl;
      if (dry) { n_launches++; continue; }
      ConvGemmDesc d;
      d.B = x.B; d.Hin = x.H; d.Win = x.W;
      d.nsrc = 1;
      d.src[0] = {x.p, x.C, x.ld()};
      d.ksize = 3; d.stride = 1; d.pad = PAD_SAME;
      d.w = w + (size_t)q * Cout * 4 * x.C;
      d.N = Cout;
      d.mode = EPI_F16;
      d.out = out.p; d.out_ld = out.C; d.out_bstride = out.HW() * out.C;
      d.bias = bias;
      d.stats = stats_ptr;
      d.poly = q + 1;
      auto l = conv_gemm_build(d, E.num_sms);
      const double by = 2.0 * x.B * ((double)x.HW() * x.C + (double)x.HW() * Cout) + 2.0 * Cout * 4.0 * x.C;
      push([l](cudaStream_t st) { conv_gemm_run(*l, st); }, 1, "tc:conv3x3_poly", fl, by);
    }
    return out;
  }

  // ---------------------------------------------------------------- constant folding of the embeddings
  // emb = time_embedding(time_proj(trans)) + bbox_embedding(emb320([0,0,1,1]))   (replace.py:430-459, meta_arch.py:178-187)
  void fold_embeddings() {
    if (W.get("temb:unet.mid_block.resnets.0")) return;
    W.require_loading("temb");
    auto sinus = [](float t, std::vector<float>& out) {  // get_timestep_embedding(dim 320, flip_sin_to_cos, shift 0)
      const int half = 160;
      for (int i = 0; i < half; ++i) {
        const float f = expf(-logf(10000.0f) * (float)i / (float)half);
        out.push_back(cosf(t * f));
      }
      for (int i = 0; i < half; ++i) {
        const float f = expf(-logf(10000.0f) * (float)i / (float)half);
        out.push_back(sinf(t * f));
      }
    };
    auto matvec = [](const std::vector<float>& w, const std::vector<float>& b, const std::vector<float>& x, int N, int K) {
      std::vector<float> y(N);
      for (int n = 0; n < N; ++n) {
        double a = b[n];
        for (int k = 0; k < K; ++k) a += (double)w[(size_t)n * K + k] * x[k];
        y[n] = (float)a;
      }
      return y;
    };
    auto silu = [](std::vector<float> v) { for (auto& x : v) x = x / (1.0f + expf(-x)); return v; };
    auto te_w1 = W.fetch("unet.time_embedding.linear_1.weight", 1280 * 320), te_b1 = W.fetch("unet.time_embedding.linear_1.bias", 1280);
    auto te_w2 = W.fetch("unet.time_embedding.linear_2.weight", 1280 * 1280), te_b2 = W.fetch("unet.time_embedding.linear_2.bias", 1280);
    auto bb_w1 = W.fetch("unet.bbox_embedding.linear_1.weight", 1280 * 1280), bb_b1 = W.fetch("unet.bbox_embedding.linear_1.bias", 1280);
    auto bb_w2 = W.fetch("unet.bbox_embedding.linear_2.weight", 1280 * 1280), bb_b2 = W.fetch("unet.bbox_embedding.linear_2.bias", 1280);
    std::vector<float> coords;
    for (float c : {0.f, 0.f, 1.f, 1.f}) sinus(c, coords);  // (4, 320) flattened -> 1280
    auto aug = matvec(bb_w2, bb_b2, silu(matvec(bb_w1, bb_b1, coords, 1280, 1280)), 1280, 1280);
    std::vector<float> emb[2];
    for (int is_trans = 0; is_trans < 2; ++is_trans) {
      std::vector<float> tp;
      sinus((float)(1 - is_trans), tp);  // trans = 1 - is_trans (meta_arch.py:237-238)
      auto op = matvec(te_w2, te_b2, silu(matvec(te_w1, te_b1, tp, 1280, 320)), 1280, 1280);
      emb[is_trans].resize(1280);
      for (int i = 0; i < 1280; ++i) emb[is_trans][i] = op[i] + aug[i];
      emb[is_trans] = silu(emb[is_trans]);  // every resnet applies SiLU to emb before time_emb_proj
    }
    // n3: the un-folded chain for per-sample coordinates (cond_embed.cu): fp32 embedding MLPs, all time_emb_proj stacked in fp16
    W.upload("emb:te_w1", te_w1.data(), te_w1.size() * 4); W.upload("emb:te_b1", te_b1.data(), te_b1.size() * 4);
    W.upload("emb:te_w2", te_w2.data(), te_w2.size() * 4); W.upload("emb:te_b2", te_b2.data(), te_b2.size() * 4);
    W.upload("emb:bb_w1", bb_w1.data(), bb_w1.size() * 4); W.upload("emb:bb_b1", bb_b1.data(), bb_b1.size() * 4);
    W.upload("emb:bb_w2", bb_w2.data(), bb_w2.size() * 4); W.upload("emb:bb_b2", bb_b2.data(), bb_b2.size() * 4);
    E.has_point_embedding = W.find("unet.point_embedding.linear_1.weight") != nullptr;
    if (E.has_point_embedding) {  // optional: a checkpoint stripped of the unused prompt heads still loads
      auto w1 = W.fetch("unet.point_embedding.linear_1.weight", 1280 * 1680), b1 = W.fetch("unet.point_embedding.linear_1.bias", 1280);
      auto w2 = W.fetch("unet.point_embedding.linear_2.weight", 1280 * 1280), b2 = W.fetch("unet.point_embedding.linear_2.bias", 1280);
      W.upload("emb:pt_w1", w1.data(), w1.size() * 4); W.upload("emb:pt_b1", b1.data(), b1.size() * 4);
      W.upload("emb:pt_w2", w2.data(), w2.size() * 4); W.upload("emb:pt_b2", b2.data(), b2.size() * 4);
    }
    std::vector<__half> tp_w;
    std::vector<float> tp_b;
    E.temb_order.clear();
    auto fold = [&](const std::string& p, int C) {
      auto w = W.fetch(p + ".time_emb_proj.weight", (int64_t)C * 1280), b = W.fetch(p + ".time_emb_proj.bias", C);
      auto cb = W.fetch(p + ".conv1.bias", C);
      E.temb_order.push_back({p, C});
      for (auto v : w) tp_w.push_back(__float2half_rn(v));
      for (int i = 0; i < C; ++i) tp_b.push_back(cb[i] + b[i]);
      std::vector<float> out(2 * C);
      for (int v = 0; v < 2; ++v) {
        auto t = matvec(w, b, emb[v], C, 1280);
        for (int i = 0; i < C; ++i) out[v * C + i] = cb[i] + t[i];
      }
      W.upload("temb:" + p, out.data(), out.size() * 4);
    };
    const int ch[4] = {320, 640, 1280, 1280};
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 2; ++j) fold("unet.down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), ch[i]);
    for (int j = 0; j < 2; ++j) fold("unet.mid_block.resnets." + std::to_string(j), 1280);
    const int rch[4] = {1280, 1280, 640, 320};
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 3; ++j) fold("unet.up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), rch[i]);
    E.temb_rows = (int)tp_b.size();
    W.upload("emb:tp_w", tp_w.data(), tp_w.size() * 2);
    W.upload("emb:tp_b", tp_b.data(), tp_b.size() * 4);
  }

  // n3: head of the plan for per-sample conditioning — coordinates -> emb -> the [B][C] bias table of every resnet
  void emit_cond_embed() {
    SDM_CHECK(E.prompt_kind == 0 || E.prompt_kind == 1, "prompt kind");
    int npad = 4, dim = 320;
    if (E.prompt_kind == 0) SDM_CHECK(E.prompt_ncoords == 4, "bbox / mask / trimap prompts carry 4 coordinates per sample");
    else {
      SDM_CHECK(E.has_point_embedding, "the checkpoint has no unet.point_embedding.* weights");
      SDM_CHECK(E.prompt_ncoords >= 1 && E.prompt_ncoords < 1680, "number of point coordinates");
      npad = 0;
      for (int i = E.prompt_ncoords; i < 1680; ++i)  // meta_arch.py:153-160: first i >= N that divides 1680
        if (1680 % i == 0) { npad = i; break; }
      SDM_CHECK(npad > 0, "point coordinates: no divisor of 1680");
      dim = 1680 / npad;
    }
    const int Kc = npad * dim;
    size_t off;
    size_t cum = 0;
    for (auto& pr : E.temb_order) { dyn_off[pr.first] = cum; cum += (size_t)B * pr.second; }
    float* coords = (float*)alloc_raw((size_t)B * E.prompt_ncoords * 4, &off);
    float* xt = (float*)alloc_raw((size_t)B * 320 * 4, &off);
    float* xc = (float*)alloc_raw((size_t)B * Kc * 4, &off);
    float* ht = (float*)alloc_raw((size_t)B * 1280 * 4, &off);
    float* hc = (float*)alloc_raw((size_t)B * 1280 * 4, &off);
    float* semb = (float*)alloc_raw((size_t)B * 1280 * 4, &off);
    dyn_tables = (float*)alloc_raw(cum * 4, &off);
    if (dry) { n_launches += 5; d_iota = (const int*)1; return; }
    // plan-owned tables (not in the caller's workspace: they must survive whatever the caller does with that memory)
    std::vector<int4> map((size_t)E.temb_rows);
    {
      size_t r = 0;
      for (auto& pr : E.temb_order)
        for (int c = 0; c < pr.second; ++c) map[r++] = make_int4((int)dyn_off[pr.first], pr.second, c, 0);
    }
    std::vector<int> iota((size_t)B);
    for (int i = 0; i < B; ++i) iota[(size_t)i] = i;
    SDM_CUDA_OK(cudaMalloc((void**)&plan->d_row_map, map.size() * sizeof(int4)));
    SDM_CUDA_OK(cudaMemcpy(plan->d_row_map, map.data(), map.size() * sizeof(int4), cudaMemcpyHostToDevice));
    SDM_CUDA_OK(cudaMalloc((void**)&plan->d_iota, (size_t)B * 4));
    SDM_CUDA_OK(cudaMemcpy(plan->d_iota, iota.data(), (size_t)B * 4, cudaMemcpyHostToDevice));
    d_iota = plan->d_iota;
    plan->d_coords = coords;
    CondEmbedDesc d;
    d.B = B; d.is_trans = d_is_trans; d.coords = coords;
    d.ncoords = E.prompt_ncoords; d.npad = npad; d.dim = dim; d.Kc = Kc;
    d.te_w1 = (const float*)W.get("emb:te_w1"); d.te_b1 = (const float*)W.get("emb:te_b1");
    d.te_w2 = (const float*)W.get("emb:te_w2"); d.te_b2 = (const float*)W.get("emb:te_b2");
    const char* pre = E.prompt_kind == 1 ? "emb:pt_" : "emb:bb_";
    d.ce_w1 = (const float*)W.get(std::string(pre) + "w1"); d.ce_b1 = (const float*)W.get(std::string(pre) + "b1");
    d.ce_w2 = (const float*)W.get(std::string(pre) + "w2"); d.ce_b2 = (const float*)W.get(std::string(pre) + "b2");
    d.tp_w = (const __half*)W.get("emb:tp_w"); d.tp_b = (const float*)W.get("emb:tp_b");
    d.rows = E.temb_rows; d.row_map = plan->d_row_map;
    d.xt = xt; d.xc = xc; d.ht = ht; d.hc = hc; d.semb = semb; d.tables = dyn_tables;
    SDM_CHECK(d.te_w1 && d.ce_w1 && d.tp_w, "conditioning weights were not packed at load time");
    push([d](cudaStream_t st) { cond_embed_run(d, st); }, 5, "cond_embed", 0, 0);
  }

  // ---------------------------------------------------------------- the whole path
  void build() {
    SDM_CHECK(R % 64 == 0 && R >= 64 && R <= 1024, "inference size must be a multiple of 64 in [64, 1024]");
    SDM_CHECK(B >= 1, "batch");
    if (W.loading) fold_embeddings();
    const int B2 = 2 * B;
    size_t off;
    // run-time inputs
    d_is_trans = (int*)alloc_raw((size_t)B * 4, &off);
    if (plan) plan->d_is_trans = d_is_trans;
    Plan::Slots* slots = plan ? &plan->slots : nullptr;
    if (E.prompt_kind >= 0 && !W.loading) emit_cond_embed();

    // ---- a1: input preparation (sdmatte_nodes.py:343,351; meta_arch.py:141)
    T x0 = alloc(B2, R, R, 64);  // im2col rows of the VAE conv_in
    if (!dry) {
      __half* xp = x0.p; const int Bc = B, Rc = R;
      push([=](cudaStream_t st) { prep_inputs_run(slots->image, slots->trimap, xp, 64, Bc, Rc, st); }, 1, "prep_inputs", 0, (double)Bc * Rc * Rc * (16 + 256));
    } else n_launches++;
    // ---- a4/a6: additive key bias per level
    int lpad[4];
    float* kb[4];
    for (int k = 0; k < 4; ++k) {
      const int s = S >> k;
      lpad[k] = (s * s + 127) & ~127;
      kb[k] = (float*)alloc_raw((size_t)B * lpad[k] * 4, &off);
    }
    if (!dry) {
      const int Bc = B, Rc = R;
      float* k0 = kb[0]; float* k1 = kb[1]; float* k2 = kb[2]; float* k3 = kb[3];
      const int l0 = lpad[0], l1 = lpad[1], l2 = lpad[2], l3 = lpad[3];
      push([=](cudaStream_t st) { const int lp[4] = {l0, l1, l2, l3}; key_bias_run(slots->trimap, Bc, Rc, k0, k1, k2, k3, lp, st); }, 1, "key_bias");
    } else n_launches++;
    // compacted keys for attn1 (SDM_ATTN_COMPACT=0: stream all keys, the A/B baseline)
    const bool compact = [] { const char* e = getenv("SDM_ATTN_COMPACT"); return e ? atoi(e) != 0 : true; }();  // read per plan build
    KeyBias kbl[4];
    {
      float* cb[4]; int* ix[4]; int* nt[4];
      for (int k = 0; k < 4; ++k) {
        kbl[k].bias = kb[k]; kbl[k].bs = lpad[k];
        if (!compact) continue;
        cb[k] = (float*)alloc_raw((size_t)B * lpad[k] * 4, &off);
        ix[k] = (int*)alloc_raw((size_t)B * lpad[k] * 4, &off);
        nt[k] = (int*)alloc_raw((size_t)B * 4, &off);
        kbl[k].cbias = cb[k]; kbl[k].idx = ix[k]; kbl[k].ntiles = nt[k];
      }
      if (compact) {
        if (!dry) {
          const int Bc = B, Sc = S;
          struct Ptrs { const float* b[4]; float* c[4]; int* i[4]; int* n[4]; int lp[4]; } pp;
          for (int k = 0; k < 4; ++k) { pp.b[k] = kb[k]; pp.c[k] = cb[k]; pp.i[k] = ix[k]; pp.n[k] = nt[k]; pp.lp[k] = lpad[k]; }
          push([=](cudaStream_t st) { key_compact_run(pp.b, pp.c, pp.i, pp.n, pp.lp, Bc, Sc, st); }, 1, "key_bias");
        } else n_launches++;
      }
    }

    // ---- a2: VAE encoder over [rgb ; trimap x3] as one batch of 2B (meta_arch.py:139-145,209-212)
    T unet_in = alloc(B, S, S, 8);
    {
      const std::string e = "vae.encoder";
      T h = alloc(B2, R, R, 128);
      { GemmOpt o; o.bias = W.vec(e + ".conv_in.bias", 128); o.stats_for = &h; o.label = "conv_in_im2col";
        conv_tc(x0, nullptr, W.conv_im2col(e + ".conv_in", 128, 3), 128, 1, h, o); }
      free(x0);
      tap("enc.conv_in", h);
      const int ch[4] = {128, 256, 512, 512};
      for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 2; ++j) {
          T o = resnet(h, nullptr, e + ".down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), ch[i], 1e-6f, false, 0);
          free(h); h = o;
        }
        if (i < 3) {
          T o = conv3_plain(h, e + ".down_blocks." + std::to_string(i) + ".downsamplers.0.conv", ch[i], 2, PAD_VAE_DOWN);
          free(h); h = o;
        }
        tap("enc.down" + std::to_string(i), h);
      }
      { T o = resnet(h, nullptr, e + ".mid_block.resnets.0", 512, 1e-6f, false, 0); free(h); h = o; }
      { T o = vae_attention(h, e + ".mid_block.attentions.0"); free(h); h = o; }
      tap("enc.mid_attn", h);
      { T o = resnet(h, nullptr, e + ".mid_block.resnets.1", 512, 1e-6f, false, 0); free(h); h = o; }
      tap("enc.mid", h);
      T n = groupnorm(h, nullptr, e + ".conv_norm_out", 1e-6f, 1);
      free(h);
      T mom = alloc(B2, S, S, 8);
      { GemmOpt o; o.bias = W.vec(e + ".conv_out.bias", 8); o.label = "conv3x3_skinny";
        conv_tc(n, nullptr, W.conv(e + ".conv_out", 8, 512, 3), 8, 3, mom, o); }
      free(n);
      // quant_conv 1x1 (8->8), keep the mean half (channels 0..3), * scaling_factor; rgb -> ch 0..3, trimap -> ch 4..7
      for (int part = 0; part < 2; ++part) {
        DirectConvDesc d; d.B = B; d.H = S; d.W = S; d.Cin = 8; d.Cout = 8; d.ksize = 1;
        d.x = dry ? nullptr : mom.p + (size_t)part * B * S * S * 8; d.x_ld = 8;
        d.w = W.conv("vae.quant_conv", 8, 8, 1); d.bias = W.vec("vae.quant_conv.bias", 8);
        d.out = unet_in.p; d.out_ld = 8; d.out_coff = part * 4; d.out_scale = 0.18215f; d.cout_limit = 4;
        direct(d);
      }
      free(mom);
    }
    tap("unet_in", unet_in);

    // ---- a5: trimap tokens = aux_conv_in(trimap latent) (meta_arch.py:215-218, utils.py:33-41)
    T ctx = alloc(B, S, S, 1024);
    { DirectConvDesc d; d.B = B; d.H = S; d.W = S; d.Cin = 4; d.Cout = 1024; d.ksize = 3; d.x = dry ? nullptr : unet_in.p + 4; d.x_ld = 8;
      d.w = W.conv("unet.aux_conv_in", 1024, 4, 3); d.bias = W.vec("unet.aux_conv_in.bias", 1024); d.out = ctx.p; d.out_ld = 1024; direct(d); }
    tap("ctx", ctx);

    // ---- a7..a14: UNet (replace.py:462-544)
    T unet_out = alloc(B, S, S, 4);
    {
      const int ch[4] = {320, 640, 1280, 1280};
      const int heads[4] = {5, 10, 20, 20};
      std::vector<T> skips;
      T h = alloc(B, S, S, 320);
      { DirectConvDesc d; d.B = B; d.H = S; d.W = S; d.Cin = 8; d.Cout = 320; d.ksize = 3; d.x = unet_in.p; d.x_ld = 8;
        d.w = W.conv("unet.conv_in", 320, 8, 3); d.bias = W.vec("unet.conv_in.bias", 320); d.out = h.p; d.out_ld = 320; direct(d); }
      tap("unet.conv_in", h);
      skips.push_back(h);
      for (int i = 0; i < 4; ++i) {
        const std::string bp = "unet.down_blocks." + std::to_string(i);
        for (int j = 0; j < 2; ++j) {
          T o = resnet(h, nullptr, bp + ".resnets." + std::to_string(j), ch[i], 1e-5f, true, 0);
          if (i < 3) {
            T o2 = transformer(o, bp + ".attentions." + std::to_string(j), heads[i], ctx, kbl[i], 0);
            free(o); o = o2;
          }
          h = o;
          tap("unet.down" + std::to_string(i) + "." + std::to_string(j), h);
          skips.push_back(h);
        }
        if (i < 3) {
          h = conv3_plain(h, bp + ".downsamplers.0.conv", ch[i], 2, PAD_SAME);
          tap("unet.down" + std::to_string(i) + ".ds", h);
          skips.push_back(h);
        }
      }
      // mid (h is the last skip; it stays alive as a skip)
      {
        T o = resnet(h, nullptr, "unet.mid_block.resnets.0", 1280, 1e-5f, true, 0);
        T o2 = transformer(o, "unet.mid_block.attentions.0", 20, ctx, kbl[3], 0);
        free(o);
        T o3 = resnet(o2, nullptr, "unet.mid_block.resnets.1", 1280, 1e-5f, true, 0);
        free(o2);
        h = o3;
        tap("unet.mid", h);
      }
      const int rch[4] = {1280, 1280, 640, 320};
      const int rheads[4] = {20, 20, 10, 5};
      for (int i = 0; i < 4; ++i) {
        const std::string bp = "unet.up_blocks." + std::to_string(i);
        const int level = 3 - i;
        // Upsample2D = nearest x2 + conv3x3: as four polyphase convs over the low-resolution tensor where the geometry allows
        // (conv3_poly), else nearest x2 fused into the producer's store + a plain 3x3 conv over the upsampled tensor
        const bool poly = i < 3 && conv_gemm_can_poly(rch[i], h.H, h.W);
        for (int j = 0; j < 3; ++j) {
          T skip = skips.back();
          skips.pop_back();
          const bool last = (j == 2) && (i < 3);
          const bool has_attn = i > 0;
          T o = resnet(h, &skip, bp + ".resnets." + std::to_string(j), rch[i], 1e-5f, true, (last && !has_attn && !poly) ? 1 : 0);
          free(h); free(skip);
          if (has_attn) {
            T o2 = transformer(o, bp + ".attentions." + std::to_string(j), rheads[i], ctx, kbl[level], (last && !poly) ? 1 : 0);
            free(o); o = o2;
          }
          h = o;
          // j == 2, i < 3: already nearest-x2 upsampled, unless the polyphase form keeps it at low resolution (".lo")
          tap("unet.up" + std::to_string(i) + "." + std::to_string(j) + ((last && poly) ? ".lo" : ""), h);
        }
        if (i < 3) {
          if (W.loading) W.conv_poly(bp + ".upsamplers.0.conv", rch[i], h.C);  // both weight forms are packed: the plan's geometry picks one
          T o = poly ? conv3_poly(h, bp + ".upsamplers.0.conv", rch[i]) : conv3_plain(h, bp + ".upsamplers.0.conv", rch[i], 1, PAD_SAME);
          free(h); h = o;
          tap("unet.up" + std::to_string(i) + ".us", h);
        }
      }
      T n = groupnorm(h, nullptr, "unet.conv_norm_out", 1e-5f, 1);
      free(h);
      // conv_out (320->4), then label_latent / scaling_factor (meta_arch.py:254)
      { GemmOpt o; o.bias = W.vec("unet.conv_out.bias", 4, 8); o.post_div = 0.18215f; o.n_store = 4; o.label = "conv3x3_skinny";
        conv_tc(n, nullptr, W.conv("unet.conv_out", 4, 320, 3, 0, 8), 8, 3, unet_out, o); }
      free(n);
      SDM_CHECK(skips.empty(), "skip bookkeeping");
    }
    // ctx / unet_in / unet_out stay allocated: they are small and double as parity-test taps
    tap("unet_out_scaled", unet_out);

    // ---- a15: VAE decode (meta_arch.py:255-256)
    {
      const std::string dcd = "vae.decoder";
      T z = alloc(B, S, S, 4);
      { DirectConvDesc d; d.B = B; d.H = S; d.W = S; d.Cin = 4; d.Cout = 8; d.ksize = 1; d.x = unet_out.p; d.x_ld = 4;
        d.w = W.conv("vae.post_quant_conv", 4, 4, 1, 0, 8); d.bias = W.vec("vae.post_quant_conv.bias", 4, 8); d.out = z.p; d.out_ld = 4; d.cout_limit = 4; direct(d); }
      T h = alloc(B, S, S, 512);
      { DirectConvDesc d; d.B = B; d.H = S; d.W = S; d.Cin = 4; d.Cout = 512; d.ksize = 3; d.x = z.p; d.x_ld = 4;
        d.w = W.conv(dcd + ".conv_in", 512, 4, 3); d.bias = W.vec(dcd + ".conv_in.bias", 512); d.out = h.p; d.out_ld = 512; direct(d); }
      free(z);
      tap("dec.conv_in", h);
      { T o = resnet(h, nullptr, dcd + ".mid_block.resnets.0", 512, 1e-6f, false, 0); free(h); h = o; }
      { T o = vae_attention(h, dcd + ".mid_block.attentions.0"); free(h); h = o; }
      { T o = resnet(h, nullptr, dcd + ".mid_block.resnets.1", 512, 1e-6f, false, 0); free(h); h = o; }
      tap("dec.mid", h);
      const int ch[4] = {512, 512, 256, 128};
      for (int i = 0; i < 4; ++i) {
        const std::string bp = dcd + ".up_blocks." + std::to_string(i);
        const bool poly = i < 3 && conv_gemm_can_poly(ch[i], h.H, h.W);  // see the UNet up blocks
        for (int j = 0; j < 3; ++j) {
          T o = resnet(h, nullptr, bp + ".resnets." + std::to_string(j), ch[i], 1e-6f, false, (j == 2 && i < 3 && !poly) ? 1 : 0);
          free(h); h = o;
        }
        if (i < 3) {
          if (W.loading) W.conv_poly(bp + ".upsamplers.0.conv", ch[i], h.C);
          T o = poly ? conv3_poly(h, bp + ".upsamplers.0.conv", ch[i]) : conv3_plain(h, bp + ".upsamplers.0.conv", ch[i], 1, PAD_SAME);
          free(h); h = o;
        }
        tap("dec.up" + std::to_string(i), h);
      }
      T n = groupnorm(h, nullptr, dcd + ".conv_norm_out", 1e-6f, 1);
      free(h);
      // ---- a16: conv_out (128->3) + channel mean + clip + (x+1)/2 (meta_arch.py:258-260)
      // r1a-r1o: ONE tcgen05 conv with N = 16 and the alpha math in the epilogue — 3.8 ms at bs=8: the nine shifted TMA boxes
      // re-read the 2.1 GB input nine times through L2 -> shared memory for 0.15 TFLOP of math.  r1p: the 27 per-tap partial
      // products of every input pixel as ONE 1x1 GEMM (input read once, fp32 out), then a col2im sum of the nine shifted
      // partials + the alpha math (the EPI_ALPHA conv survives as a kernel-level test only).
      const float* cbias = W.vec(dcd + ".conv_out.bias", 3, 8);
      {
        const __half* wt = W.conv_taprows(dcd + ".conv_out", 3, 128);
        size_t yoff;
        const size_t ybytes = (size_t)B * R * R * 32 * sizeof(float);
        float* y = (float*)alloc_raw(ybytes, &yoff);
        { GemmOpt o; o.mode = EPI_F32; o.label = "alpha_head_taps";
          T yv; yv.B = B; yv.H = R; yv.W = R; yv.C = 32; yv.p = (__half*)y;
          conv_tc(n, nullptr, wt, 32, 1, yv, o); }
        if (!dry) {
          Plan::Slots* slots = &plan->slots;
          const int Bc = B, Rc = R;
          push([=](cudaStream_t st) { alpha_col2im_run(y, cbias, Bc, Rc, Rc, slots->alpha, slots->premean, st); }, 1, "alpha_col2im", 0,
               (double)B * R * R * (128.0 + 2.0));
        } else n_launches++;
        arena.release(yoff, ybytes);
      }
      free(n);
    }
  }
};

// ================================================================================================
// engine API
// ================================================================================================
Engine* engine_create(int device) {
  SDM_CUDA_OK(cudaSetDevice(device));
  Engine* e = new Engine();
  e->device = device;
  SDM_CUDA_OK(cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, device));
  int major = 0, minor = 0;
  SDM_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  SDM_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10) {
    delete e;
    throw Error{"sdmatte_b200 requires an sm_100a (B200) device; found sm_" + std::to_string(major) + std::to_string(minor) + " — there is no fallback path"};
  }
  return e;
}

void engine_destroy(Engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  e->plan.reset();
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  if (e->flags_event) cudaEventDestroy(e->flags_event);
  if (e->flags_pinned) cudaFreeHost(e->flags_pinned);
  if (e->pinned) cudaFreeHost(e->pinned);
  e->W.free_all();
  delete e;
}

void engine_load(Engine* e, const sdm_tensor_desc* tensors, int n) {
  SDM_CUDA_OK(cudaSetDevice(e->device));
  Weights& W = e->W;
  W.loaded = false;  // a failed (re)load must leave the engine "not loaded", never half-loaded
  W.free_all();
  W.host.clear();
  W.missing.clear();
  e->plan.reset();
  SDM_CHECK(n >= 0 && (n == 0 || tensors != nullptr), "sdm_load_weights: null descriptor array");
  for (int i = 0; i < n; ++i) {
    // the descriptors come across the C ABI: validate before touching shape[] (4 entries) or dereferencing anything
    SDM_CHECK(tensors[i].name != nullptr, "sdm_load_weights: descriptor without a name");
    if (tensors[i].ndim < 0 || tensors[i].ndim > 4)
      throw Error{std::string("sdm_load_weights: '") + tensors[i].name + "' has rank " + std::to_string(tensors[i].ndim) + " (0..4 supported)"};
    if (tensors[i].dtype < 0 || tensors[i].dtype > 2)
      throw Error{std::string("sdm_load_weights: '") + tensors[i].name + "' has an unsupported dtype code " + std::to_string(tensors[i].dtype)};
    SDM_CHECK(tensors[i].data != nullptr, "sdm_load_weights: descriptor without data");
    for (int d = 0; d < tensors[i].ndim; ++d) SDM_CHECK(tensors[i].shape[d] >= 0, "sdm_load_weights: negative extent");
    HostT t;
    t.dtype = tensors[i].dtype;
    for (int d = 0; d < tensors[i].ndim; ++d) t.shape.push_back(tensors[i].shape[d]);
    t.data = tensors[i].data;
    W.host[tensors[i].name] = t;
  }
  W.loading = true;
  try {
    Builder b(*e, nullptr, 1, 64, nullptr);
    b.build();
  } catch (...) {
    W.loading = false;
    W.host.clear();
    throw;
  }
  W.loading = false;
  W.n_used = 0;
  W.n_unexpected = 0;
  for (auto& kv : W.host) (kv.second.used ? W.n_used : W.n_unexpected)++;
  W.host.clear();
  if (!W.missing.empty()) {
    std::string msg = std::to_string(W.missing.size()) + " required checkpoint keys are missing, e.g.:";
    for (size_t i = 0; i < std::min<size_t>(W.missing.size(), 8); ++i) msg += " " + W.missing[i];
    W.free_all();
    throw Error{msg};
  }
  W.loaded = true;
}

void engine_load_report(Engine* e, int* n_used, int* n_unexpected) {
  if (n_used) *n_used = e->W.n_used;
  if (n_unexpected) *n_unexpected = e->W.n_unexpected;
}

size_t engine_workspace_bytes(Engine* e, int B, int R) {
  SDM_CHECK(e->W.loaded, "weights not loaded");
  Builder b(*e, nullptr, B, R, nullptr);
  b.build();
  return b.arena.peak + 4096;
}

size_t engine_workspace_bytes_prompt(Engine* e, int B, int R, int prompt_kind, int ncoords) {
  const int k0 = e->prompt_kind, n0 = e->prompt_ncoords;
  e->prompt_kind = prompt_kind; e->prompt_ncoords = prompt_kind >= 0 ? ncoords : 0;
  size_t r = 0;
  try { r = engine_workspace_bytes(e, B, R); } catch (...) { e->prompt_kind = k0; e->prompt_ncoords = n0; throw; }
  e->prompt_kind = k0; e->prompt_ncoords = n0;
  return r;
}

static Plan& get_plan(Engine* e, int B, int R, void* ws, size_t ws_bytes) {
  SDM_CHECK(e->W.loaded, "weights not loaded");
  SDM_CHECK((reinterpret_cast<uintptr_t>(ws) & 1023) == 0, "workspace must be 1024-byte aligned");
  if (e->plan && e->plan->B == B && e->plan->R == R && e->plan->ws == ws && e->plan->ws_bytes == ws_bytes && e->plan->keep_taps == e->keep_taps &&
      e->plan->prompt_kind == e->prompt_kind && e->plan->prompt_ncoords == e->prompt_ncoords) return *e->plan;
  const size_t need = engine_workspace_bytes(e, B, R);
  if (ws_bytes < need) throw Error{"workspace too small: need " + std::to_string(need) + " bytes, got " + std::to_string(ws_bytes)};
  auto plan = std::make_unique<Plan>();
  plan->B = B; plan->R = R; plan->ws = ws; plan->ws_bytes = ws_bytes; plan->keep_taps = e->keep_taps;
  plan->prompt_kind = e->prompt_kind; plan->prompt_ncoords = e->prompt_ncoords;
  Builder b(*e, plan.get(), B, R, ws);
  b.build();
  plan->n_launches = b.n_launches;
  plan->tensor_flops = b.flops;
  e->plan = std::move(plan);
  return *e->plan;
}

// per-sample flags (and prompt coordinates): staged through a small page-locked buffer and uploaded at the head of EVERY forward —
// the workspace is the caller's memory and may have been recycled since the last call, so nothing in it is assumed to persist
static void upload_flags(Engine* e, Plan& p, const int32_t* is_trans, int B, const float* coords, int ncoords, cudaStream_t st) {
  for (int i = 0; i < B; ++i) SDM_CHECK(is_trans[i] == 0 || is_trans[i] == 1, "is_trans must be 0/1");
  const size_t nflag = (size_t)B * 4, ncoor = coords ? (size_t)B * ncoords * 4 : 0;
  SDM_CHECK(nflag + ncoor <= 64 * 1024, "batch too large for the flag staging buffer");
  if (!e->flags_pinned) {
    SDM_CUDA_OK(cudaHostAlloc((void**)&e->flags_pinned, 64 * 1024, cudaHostAllocDefault));
    SDM_CUDA_OK(cudaEventCreateWithFlags(&e->flags_event, cudaEventDisableTiming));
  } else {
    SDM_CUDA_OK(cudaEventSynchronize(e->flags_event));  // the previous upload has left the staging buffer
  }
  memcpy(e->flags_pinned, is_trans, nflag);
  SDM_CUDA_OK(cudaMemcpyAsync(p.d_is_trans, e->flags_pinned, nflag, cudaMemcpyHostToDevice, st));
  if (coords) {
    SDM_CHECK(p.d_coords != nullptr, "plan without a coordinates buffer");
    memcpy((char*)e->flags_pinned + nflag, coords, ncoor);
    SDM_CUDA_OK(cudaMemcpyAsync(p.d_coords, (char*)e->flags_pinned + nflag, ncoor, cudaMemcpyHostToDevice, st));
  }
  SDM_CUDA_OK(cudaEventRecord(e->flags_event, st));
}

static void forward_impl(Engine* e, const float* image_dev, const float* trimap_dev, int B, int R, const int32_t* is_trans,
                         int prompt_kind, const float* coords_host, int ncoords,
                         void* alpha_dev, void* premean_dev, void* ws, size_t ws_bytes, cudaStream_t st);

void engine_forward(Engine* e, const float* image_dev, const float* trimap_dev, int B, int R, const int32_t* is_trans,
                    void* alpha_dev, void* premean_dev, void* ws, size_t ws_bytes, cudaStream_t st) {
  forward_impl(e, image_dev, trimap_dev, B, R, is_trans, -1, nullptr, 0, alpha_dev, premean_dev, ws, ws_bytes, st);
}
// n3: the auxiliary image is a mask / bbox mask (prompt_kind 0, 4 coordinates per sample) or a point mask (1, n coordinates)
void engine_forward_prompt(Engine* e, const float* image_dev, const float* aux_dev, int B, int R, const int32_t* is_trans, int prompt_kind,
                           const float* coords_host, int ncoords, void* alpha_dev, void* premean_dev, void* ws, size_t ws_bytes,
                           cudaStream_t st) {
  SDM_CHECK(prompt_kind == 0 || prompt_kind == 1, "prompt_kind: 0 = bbox / mask (4 coordinates), 1 = points");
  SDM_CHECK(coords_host != nullptr && ncoords >= 1, "coordinates");
  forward_impl(e, image_dev, aux_dev, B, R, is_trans, prompt_kind, coords_host, ncoords, alpha_dev, premean_dev, ws, ws_bytes, st);
}

static void forward_impl(Engine* e, const float* image_dev, const float* trimap_dev, int B, int R, const int32_t* is_trans,
                         int prompt_kind, const float* coords_host, int ncoords,
                         void* alpha_dev, void* premean_dev, void* ws, size_t ws_bytes, cudaStream_t st) {
  SDM_CUDA_OK(cudaSetDevice(e->device));
  e->prompt_kind = prompt_kind;
  e->prompt_ncoords = prompt_kind >= 0 ? ncoords : 0;
  Plan& p = get_plan(e, B, R, ws, ws_bytes);
  p.slots.image = image_dev;
  p.slots.trimap = trimap_dev;
  p.slots.alpha = (__half*)alpha_dev;
  p.slots.premean = (__half*)premean_dev;
  upload_flags(e, p, is_trans, B, coords_host, ncoords, st);
  e->last_launches = p.n_launches;
  e->last_flops = p.tensor_flops;
  auto same = [](const Plan::Slots& a, const Plan::Slots& b) { return memcmp(&a, &b, sizeof(Plan::Slots)) == 0; };
  if (e->use_graph && p.graph_exec && same(p.graph_slots, p.slots)) {
    SDM_CUDA_OK(cudaGraphLaunch(p.graph_exec, st));
    e->graph_launches++;
    return;
  }
  // capture when the same pointers come back (first run of a plan is always eager: it also sets the per-device kernel
  // attributes); callers that hand in fresh buffers every time simply stay on the eager path
  if (e->use_graph && p.eager_runs >= 1 && same(p.last_slots, p.slots)) {
    if (!e->cap_stream) SDM_CUDA_OK(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
    cudaGraph_t g = nullptr;
    SDM_CUDA_OK(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
    try {
      for (auto& op : p.ops) op.fn(e->cap_stream);
    } catch (...) {
      cudaStreamEndCapture(e->cap_stream, &g);
      if (g) cudaGraphDestroy(g);
      throw;
    }
    SDM_CUDA_OK(cudaStreamEndCapture(e->cap_stream, &g));
    if (p.graph_exec) { cudaGraphExecDestroy(p.graph_exec); p.graph_exec = nullptr; }
    const cudaError_t ie = cudaGraphInstantiate(&p.graph_exec, g, 0);
    cudaGraphDestroy(g);
    SDM_CUDA_OK(ie);
    p.graph_slots = p.slots;
    e->graph_captures++;
    SDM_CUDA_OK(cudaGraphLaunch(p.graph_exec, st));
    e->graph_launches++;
    return;
  }
  for (auto& op : p.ops) op.fn(st);
  p.eager_runs++;
  p.last_slots = p.slots;
}

void engine_forward_profiled(Engine* e, const float* image_dev, const float* trimap_dev, int B, int R, const int32_t* is_trans,
                             void* alpha_dev, void* ws, size_t ws_bytes, cudaStream_t st) {
  SDM_CUDA_OK(cudaSetDevice(e->device));
  Plan& p = get_plan(e, B, R, ws, ws_bytes);
  p.slots.image = image_dev;
  p.slots.trimap = trimap_dev;
  p.slots.alpha = (__half*)alpha_dev;
  p.slots.premean = nullptr;
  upload_flags(e, p, is_trans, B, nullptr, 0, st);
  std::vector<cudaEvent_t> ev(p.ops.size() + 1);
  for (auto& x : ev) SDM_CUDA_OK(cudaEventCreate(&x));
  SDM_CUDA_OK(cudaEventRecord(ev[0], st));
  for (size_t i = 0; i < p.ops.size(); ++i) {
    p.ops[i].fn(st);
    SDM_CUDA_OK(cudaEventRecord(ev[i + 1], st));
  }
  SDM_CUDA_OK(cudaStreamSynchronize(st));
  e->prof.clear();
  for (size_t i = 0; i < p.ops.size(); ++i) {
    float ms = 0;
    SDM_CUDA_OK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
    e->prof.push_back({p.ops[i].kind, ms, p.ops[i].flops, p.ops[i].bytes});
  }
  for (auto& x : ev) cudaEventDestroy(x);
}
int engine_profile_count(Engine* e) { return (int)e->prof.size(); }
void engine_profile_entry(Engine* e, int i, char* kind, int kind_len, float* ms, double* flops, double* bytes) {
  SDM_CHECK(i >= 0 && i < (int)e->prof.size(), "profile index");
  snprintf(kind, kind_len, "%s", e->prof[i].kind.c_str());
  *ms = e->prof[i].ms; *flops = e->prof[i].flops; *bytes = e->prof[i].bytes;
}

void engine_forward_host(Engine* e, const float* image_host, const float* trimap_host, int B, int R, const int32_t* is_trans,
                         void* alpha_host_f16, void* ws, size_t ws_bytes, cudaStream_t st) {
  SDM_CUDA_OK(cudaSetDevice(e->device));
  // staging buffers live at the tail of the workspace: image, trimap, alpha
  const size_t img_b = (size_t)B * R * R * 3 * 4, tri_b = (size_t)B * R * R * 4, al_b = (size_t)B * R * R * 2;
  const size_t stage = ((img_b + tri_b + al_b) + 4095) & ~(size_t)4095;
  SDM_CHECK(ws_bytes > stage, "workspace too small for host staging");
  const size_t main_bytes = (ws_bytes - stage) & ~(size_t)1023;
  char* tail = (char*)ws + main_bytes;
  float* d_img = (float*)tail;
  float* d_tri = (float*)(tail + img_b);
  __half* d_alpha = (__half*)(tail + img_b + tri_b);
  SDM_CUDA_OK(cudaMemcpyAsync(d_img, image_host, img_b, cudaMemcpyHostToDevice, st));
  SDM_CUDA_OK(cudaMemcpyAsync(d_tri, trimap_host, tri_b, cudaMemcpyHostToDevice, st));
  engine_forward(e, d_img, d_tri, B, R, is_trans, d_alpha, nullptr, ws, main_bytes, st);
  SDM_CUDA_OK(cudaMemcpyAsync(alpha_host_f16, d_alpha, al_b, cudaMemcpyDeviceToHost, st));
  SDM_CUDA_OK(cudaStreamSynchronize(st));
}

// ================================================================================================
// the node call: host tensors of ANY size in, host tensors out (replaces sdmatte_nodes.py:339-397 around the forward)
// ================================================================================================
// Layout of the caller's workspace: [ plan arena (engine_workspace_bytes) | image HxW | trimap HxW | image RxR | trimap RxR |
// alpha RxR | alpha HxW | matted HxW ] — fixed offsets for a given geometry, so the plan's CUDA graph sees the same pointers
// on every call.
struct NodeLayout {
  size_t main_bytes, img, tri, img_r, tri_r, alpha_r, out, matted, total;
  int mch;
};
static NodeLayout node_layout(Engine* e, int B, int H, int W, int R, int output_mode) {
  auto al = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
  NodeLayout L{};
  L.mch = output_mode == 1 ? 4 : (output_mode == 0 ? 0 : 3);
  const size_t hw = (size_t)B * H * W, rr = (size_t)B * R * R;
  const bool resize = !(H == R && W == R);
  L.main_bytes = al(engine_workspace_bytes(e, B, R));
  size_t o = L.main_bytes;
  L.img = o; o += al(hw * 12);
  L.tri = o; o += al(hw * 4);
  L.img_r = resize ? o : L.img; if (resize) o += al(rr * 12);
  L.tri_r = resize ? o : L.tri; if (resize) o += al(rr * 4);
  L.alpha_r = o; o += al(rr * 2);
  L.out = o; o += al(hw * 2);
  L.matted = o; o += al(hw * 4 * L.mch);
  L.total = o + 1024;
  return L;
}
size_t engine_node_workspace_bytes(Engine* e, int B, int H, int W, int R, int output_mode) {
  SDM_CHECK(B >= 1 && H >= 1 && W >= 1, "node geometry");
  SDM_CHECK(output_mode >= 0 && output_mode <= 3, "output_mode");
  return node_layout(e, B, H, W, R, output_mode).total;
}

// Host copy with non-temporal stores: the destination (page-locked staging, or the caller's result tensor) is not read again by
// this core, so bypassing the cache saves the read-for-ownership of every destination line (r2j: 134 MB of inputs staged at
// ~30 GB/s by 8 memcpy threads = 4.4 ms in front of a 266 ms forward; glibc only switches to streaming stores above a size
// threshold that the 4 MB chunks stay under).
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
__attribute__((target("avx2"))) static void copy_nt_avx2(char* dst, const char* src, size_t n) {
  size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
  if (head > n) head = n;
  memcpy(dst, src, head);
  dst += head; src += head; n -= head;
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
    const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
  }
  _mm_sfence();
  memcpy(dst + i, src + i, n - i);
}
static void copy_nt(char* dst, const char* src, size_t n) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2 && n >= 4096) copy_nt_avx2(dst, src, n);
  else memcpy(dst, src, n);
}
#else
static void copy_nt(char* dst, const char* src, size_t n) { memcpy(dst, src, n); }
#endif

// pageable -> pinned -> device: `nt` host threads copy disjoint chunks into the page-locked buffer and each enqueues the H2D of
// its chunk right behind it, so the staging memcpy (the slow half: host DRAM bandwidth) overlaps the DMA of earlier chunks
static void stage_h2d(Engine* e, char* pin, void* dst_dev, const void* src_host, size_t bytes, cudaStream_t st) {
  const size_t chunk = 4u << 20;
  const size_t nchunks = (bytes + chunk - 1) / chunk;
  const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)e->copy_threads, nchunks));
  std::vector<std::thread> th;
  std::vector<cudaError_t> err((size_t)nt, cudaSuccess);
  const int device = e->device;
  for (int t = 0; t < nt; ++t)
    th.emplace_back([=, &err] {
      cudaSetDevice(device);
      for (size_t c = (size_t)t; c < nchunks; c += (size_t)nt) {
        const size_t off = c * chunk, n = std::min(chunk, bytes - off);
        copy_nt(pin + off, (const char*)src_host + off, n);
        const cudaError_t r = cudaMemcpyAsync((char*)dst_dev + off, pin + off, n, cudaMemcpyHostToDevice, st);
        if (r != cudaSuccess) err[(size_t)t] = r;
      }
    });
  for (auto& x : th) x.join();
  for (auto r : err) SDM_CUDA_OK(r);
}
static void parallel_memcpy(Engine* e, void* dst, const void* src, size_t bytes) {
  const size_t chunk = 4u << 20;
  const size_t nchunks = (bytes + chunk - 1) / chunk;
  const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)e->copy_threads, nchunks));
  if (nt == 1) { copy_nt((char*)dst, (const char*)src, bytes); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t)
    th.emplace_back([=] {
      for (size_t c = (size_t)t; c < nchunks; c += (size_t)nt) {
        const size_t off = c * chunk, n = std::min(chunk, bytes - off);
        copy_nt((char*)dst + off, (const char*)src + off, n);
      }
    });
  for (auto& x : th) x.join();
}

void engine_apply_host(Engine* e, const float* image_host, const float* trimap_host, int B, int H, int W, int R, const int32_t* is_trans,
                       int mask_refine, double trimap_constraint, int output_mode, void* alpha_out_host_f16, float* matted_out_host,
                       void* ws, size_t ws_bytes, cudaStream_t st) {
  SDM_CUDA_OK(cudaSetDevice(e->device));
  SDM_CHECK(image_host && trimap_host && alpha_out_host_f16, "null host tensor");
  const NodeLayout L = node_layout(e, B, H, W, R, output_mode);
  SDM_CHECK(ws_bytes >= L.total, "workspace too small for the node call (sdm_node_workspace_bytes)");
  SDM_CHECK(L.mch == 0 || matted_out_host != nullptr, "matted_out needed for this output_mode");
  char* base = (char*)ws;
  const size_t hw = (size_t)B * H * W, rr = (size_t)B * R * R;
  const size_t in_b = hw * 16, out_b = hw * 2 + hw * 4 * L.mch;
  const size_t need = ((in_b + 4095) & ~(size_t)4095) + out_b;
  if (e->pinned_bytes < need) {
    if (e->pinned) { cudaFreeHost(e->pinned); e->pinned = nullptr; e->pinned_bytes = 0; }
    SDM_CUDA_OK(cudaHostAlloc((void**)&e->pinned, need, cudaHostAllocDefault));
    e->pinned_bytes = need;
  }
  char* pin_in = e->pinned;
  char* pin_out = e->pinned + ((in_b + 4095) & ~(size_t)4095);
  const auto t0 = std::chrono::steady_clock::now();
  // trimap first (the key-bias / compaction kernels need only it), then the image
  stage_h2d(e, pin_in, base + L.tri, trimap_host, hw * 4, st);
  stage_h2d(e, pin_in + hw * 4, base + L.img, image_host, hw * 12, st);
  const auto t1 = std::chrono::steady_clock::now();
  if (!(H == R && W == R))
    preprocess_run((const float*)(base + L.img), (const float*)(base + L.tri), B, H, W, R, (float*)(base + L.img_r), (float*)(base + L.tri_r), st);
  engine_forward(e, (const float*)(base + L.img_r), (const float*)(base + L.tri_r), B, R, is_trans, base + L.alpha_r, nullptr, ws, L.main_bytes, st);
  postprocess_run((const __half*)(base + L.alpha_r), B, R, H, W, (const float*)(base + L.img), (const float*)(base + L.tri), mask_refine,
                  trimap_constraint, output_mode, (__half*)(base + L.out), L.mch ? (float*)(base + L.matted) : nullptr, st);
  SDM_CUDA_OK(cudaMemcpyAsync(pin_out, base + L.out, hw * 2, cudaMemcpyDeviceToHost, st));
  if (L.mch) SDM_CUDA_OK(cudaMemcpyAsync(pin_out + hw * 2, base + L.matted, hw * 4 * L.mch, cudaMemcpyDeviceToHost, st));
  const auto t2 = std::chrono::steady_clock::now();
  SDM_CUDA_OK(cudaStreamSynchronize(st));
  const auto t3 = std::chrono::steady_clock::now();
  parallel_memcpy(e, alpha_out_host_f16, pin_out, hw * 2);
  if (L.mch) parallel_memcpy(e, matted_out_host, pin_out + hw * 2, hw * 4 * L.mch);
  const auto t4 = std::chrono::steady_clock::now();
  auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  e->node_ms[0] = ms(t0, t1); e->node_ms[1] = ms(t1, t2); e->node_ms[2] = ms(t2, t3); e->node_ms[3] = ms(t3, t4);
  (void)rr;
}

void engine_stats(Engine* e, int* n_launches, double* tensor_flops) {
  if (n_launches) *n_launches = e->last_launches;
  if (tensor_flops) *tensor_flops = e->last_flops;
}

void engine_node_timing(Engine* e, double* ms4) { for (int i = 0; i < 4; ++i) ms4[i] = e->node_ms[i]; }
void engine_graph_stats(Engine* e, int* captures, int* launches) {
  if (captures) *captures = e->graph_captures;
  if (launches) *launches = e->graph_launches;
}

void engine_debug_tensor(Engine* e, const char* name, void* dst_dev, size_t dst_bytes, int64_t* shape4, int* dtype) {
  SDM_CHECK(e->plan != nullptr, "no forward has run yet");
  auto it = e->plan->taps.find(name);
  if (it == e->plan->taps.end()) throw Error{std::string("unknown debug tensor '") + name + "'"};
  const T& t = it->second;
  const size_t bytes = (size_t)t.B * t.H * t.W * t.C * 2;
  if (shape4) { shape4[0] = t.B; shape4[1] = t.H; shape4[2] = t.W; shape4[3] = t.C; }
  if (dtype) *dtype = 1;
  if (!dst_dev) return;  // shape query
  SDM_CHECK(dst_bytes >= bytes, "debug tensor destination too small");
  SDM_CUDA_OK(cudaSetDevice(e->device));
  SDM_CUDA_OK(cudaMemcpy(dst_dev, (const char*)e->plan->ws + t.off, bytes, cudaMemcpyDeviceToDevice));
}

int engine_debug_tensor_count(Engine* e) { return e->plan ? (int)e->plan->tap_order.size() : 0; }
const char* engine_debug_tensor_name(Engine* e, int i) {
  SDM_CHECK(e->plan && i >= 0 && i < (int)e->plan->tap_order.size(), "tap index");
  return e->plan->tap_order[i].c_str();
}

void engine_set_option(Engine* e, const char* name, int value) {
  const std::string n = name ? name : "";
  if (n == "keep_taps") { e->keep_taps = value != 0; e->plan.reset(); return; }
  if (n == "cuda_graph") { e->use_graph = value != 0; e->plan.reset(); return; }
  if (n == "copy_threads") { e->copy_threads = std::max(1, std::min(64, value)); return; }
  throw Error{"unknown engine option '" + n + "'"};
}

}  // namespace sdm
