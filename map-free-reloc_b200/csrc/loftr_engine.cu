// LoFTR matcher engine: the whole coarse-to-fine network as a static launch list of hand-written
// sm_100a kernels (tcgen05 implicit-GEMM convs / linears + SIMT glue), replayed through a CUDA graph.
// Mirrors etc/feature_matching_baselines/LoFTR/src/loftr/loftr.py:29-75 of the reference
// (backbone -> pos-enc -> coarse transformer -> dual-softmax mutual-NN -> fine windows -> fine
// transformer -> expectation), for B image pairs per call, channel-last fp32 activations.
#include "loftr_engine.cuh"

#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "conv_gemm.cuh"
#include "encoder_tail.cuh"
#include "loftr_ops.cuh"

namespace mfr {

struct WeightSpec {
  std::string name;
  size_t numel;
  size_t offset;  // floats, 64-float aligned
};

struct DebugTensor {
  float* ptr;
  long long numel;
};

struct StepInfo {
  std::string name;
  int kind;             // 1 = tcgen05 conv/GEMM launch, 0 = SIMT kernel(s)
  double flops;         // algorithmic FLOPs (static part); for dyn steps: FLOPs per live row
  const int* dyn_rows;  // device row count (per image) or null
  int dyn_images;
  int launches = 1;     // kernels this step launches
};

// bf16 mode: the coarse correlation reads the bf16 shadow of the final tokens (kind::f16, half the operand bytes)
constexpr bool kCorrelationBf16 = true;
constexpr size_t kW16Capacity = static_cast<size_t>(16) * 1024 * 1024;   // bf16 weight copies (elements): ~11.5 M used

struct LoftrEngine {
  int B, H, W, NI;
  int H2, W2, H4, W4, H8, W8, L, Mcap;
  float thr, temperature;
  int border;
  int splits_kv = 96, splits_col = 32;   // splits_kv is set from the image count in build()
  int precision = 0;   // MFR_PRECISION_* (include/mfr_b200.h)
  bool fused_tail = getenv("MFR_FUSED_TAIL") == nullptr || atoi(getenv("MFR_FUSED_TAIL")) != 0;   // bf16 mode: encoder_tail.cu
  bool fused_fine_attn = getenv("MFR_FUSED_FINE_ATTN") != nullptr && atoi(getenv("MFR_FUSED_FINE_ATTN")) != 0;   // bf16 mode, fine level (opt-in)
  // bf16 mode: the q|k|v GEMMs store raw projections and the attention kernels apply elu + 1 to their mma fragments
  // The pose back end of the previous batch runs on its own stream under the first layers of this batch's matcher. Its blocks
  // do not fit next to the persistent GEMM CTAs, and a persistent kernel with a static tile assignment whose CTA has to wait
  // for an SM is late by that CTA's whole share: the first `reserve_layers` conv launches (and the stem) therefore leave
  // `reserve_sms` SMs free (MFR_RESERVE_SMS / MFR_RESERVE_LAYERS; 0 = use every SM).
  int reserve_sms = getenv("MFR_RESERVE_SMS") ? atoi(getenv("MFR_RESERVE_SMS")) : 0;
  int reserve_layers = getenv("MFR_RESERVE_LAYERS") ? atoi(getenv("MFR_RESERVE_LAYERS")) : 4;
  int conv_count = 0;
  bool elu_in_attn = getenv("MFR_ELU_IN_ATTN") == nullptr || atoi(getenv("MFR_ELU_IN_ATTN")) != 0;
  struct Bf16Conv { std::string name; int cout, taps, cin, cin_pad; size_t off; };
  std::vector<Bf16Conv> bf16_convs;
  unsigned short* wblob16 = nullptr;
  float* wblob3 = nullptr;     // precision 2: [hi | hi | lo] split of every GEMM weight (3 * Cin per row)
  float* x3 = nullptr;         // precision 2: scratch for the [hi | lo | hi] split of a GEMM's activations
  size_t x3_cap = 0;
  size_t w16_total = 0;
  bool w16_ready = false;

  std::vector<WeightSpec> specs;
  std::map<std::string, size_t> spec_index;
  float* wblob = nullptr;
  size_t wtotal = 0;
  bool weights_loaded = false;
  std::vector<char> weight_set;   // per weight tensor: uploaded at least once
  int last_src_h = -1, last_src_w = -1;

  std::vector<void*> allocs;
  std::map<std::string, DebugTensor> dbg;
  std::vector<std::function<int(cudaStream_t)>> steps;
  std::vector<StepInfo> infos;
  void push(const std::string& name, std::function<int(cudaStream_t)> f, int launches = 1) {
    steps.push_back(std::move(f));
    StepInfo inf{name, 0, 0.0, nullptr, 0};
    inf.launches = launches;
    infos.push_back(inf);
  }
  // buffers
  float* img = nullptr;
  float *a0, *a1, *a2, *b0, *b1, *b2, *c0, *c1, *c2, *fc, *p4a, *p4b, *p4c, *p2a, *p2b;
  float *tok, *qkv, *msg, *mrg, *h1, *h2, *kvp, *kv, *S;
  float *tokb = nullptr, *ftokb = nullptr;   // bf16 mode: bf16 shadows of tok / ftok (same shape; float-typed like the other bf16 buffers)
  size_t ftok_elems = 0;
  float* shadow(float* x) const {            // same element offset inside the shadow buffer
    if (x >= ftok && x < ftok + ftok_elems) return ftokb + (x - ftok) / 2;
    return tokb + (x - tok) / 2;
  }
  CoarseMatchBuffers cm;
  float *cin, *cproj, *win, *ftok, *fqkv, *fmsg, *fmrg, *fh1, *fh2, *mk1f, *expec;
  // graph
  cudaGraphExec_t graph_exec = nullptr;
  bool use_graph = true;
  int build_rc = MFR_OK;

  float* W_(const std::string& n) {
    auto it = spec_index.find(n);
    if (it == spec_index.end()) {
      fprintf(stderr, "[mfr_b200] unknown weight %s\n", n.c_str());
      build_rc = MFR_ERR_ARG;
      return wblob;
    }
    return wblob + specs[it->second].offset;
  }
  void add_spec(const std::string& n, size_t numel) {
    WeightSpec s{n, numel, wtotal};
    spec_index[n] = specs.size();
    specs.push_back(s);
    wtotal += (numel + 63) / 64 * 64;
  }
  template <typename T>
  T* alloc(size_t n, const char* name = nullptr) {
    void* p = nullptr;
    if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) {
      fprintf(stderr, "[mfr_b200] cudaMalloc of %zu bytes failed (%s)\n", n * sizeof(T), name ? name : "");
      build_rc = MFR_ERR_CUDA;
      return nullptr;
    }
    cudaMemset(p, 0, n * sizeof(T));
    allocs.push_back(p);
    if (name) dbg[name] = DebugTensor{reinterpret_cast<float*>(p), static_cast<long long>(n)};
    return reinterpret_cast<T*>(p);
  }

  // ---- step builders -----------------------------------------------------------------------
  void conv(const std::string& wname, const float* x, int N, int Hh, int Ww, int Cin, long long ldx,
            int Cout, int R, int stride, bool bias, const float* residual, int act, float* out,
            long long ldo, int act_cols = 0, const int* dyn_w = nullptr, const float* w_override = nullptr,
            const float* bias_override = nullptr, int in_bf16 = 0, int out_bf16 = 0, int w_rows_total = 0,
            int w_row_off = 0) {
    ConvGemmDesc d{};
    int cin_gemm = Cin;                       // K of the GEMM per tap (3 * Cin in the 3xTF32 mode)
    if (precision == 2 && !in_bf16) {
      // fp32-faithful mode: A = [hi | lo | hi] of the activations (split into the scratch x3 right before the GEMM),
      // B = [hi | hi | lo] of the weights (wblob3, built once per weight upload) -> one TF32 GEMM over 3 * Cin
      const long long rows_img = static_cast<long long>(Hh) * Ww;
      if (static_cast<size_t>(N) * rows_img * 3 * Cin > x3_cap) {
        fprintf(stderr, "[mfr_b200] split scratch too small for %s\n", wname.c_str());
        build_rc = MFR_ERR_ARG;
        return;
      }
      float* x3_ = x3; const float* xs = x; const int N_ = N, C_ = Cin; const long long ldx_ = ldx;
      push("split3", [=](cudaStream_t st) { return rows_split3(xs, ldx_, rows_img * ldx_, x3_, rows_img, N_, C_, 0, dyn_w, st); });
      size_t off = w16_total;
      bool found = false;
      for (const auto& e : bf16_convs)
        if (e.name == wname) { off = e.off; found = true; }
      if (!found) {
        const int rows_total = w_rows_total > 0 ? w_rows_total : Cout;
        bf16_convs.push_back(Bf16Conv{wname, rows_total, R * R, Cin, 3 * Cin, off});
        w16_total += (static_cast<size_t>(rows_total) * R * R * 3 * Cin + 63) / 64 * 64;
      }
      w_override = wblob3 + off + static_cast<size_t>(w_row_off) * R * R * 3 * Cin;
      x = x3; ldx = 3 * Cin; cin_gemm = 3 * Cin;
    }
    d.x = x; d.N = N; d.H = Hh; d.W = Ww; d.Cin = cin_gemm; d.ldx = ldx;
    d.in_bf16 = in_bf16; d.out_bf16 = out_bf16;
    if (in_bf16) {   // bf16 copy of the weights, Cin padded to a multiple of 8 (16-byte TMA strides)
      const int cin_pad = (Cin + 7) / 8 * 8;
      size_t off = w16_total;
      bool found = false;
      for (const auto& e : bf16_convs)
        if (e.name == wname) { off = e.off; found = true; }     // cross layers build the same weights twice
      if (!found) {
        // a GEMM may use a row window [w_row_off, w_row_off + Cout) of a larger weight (q / k|v parts of qkv)
        const int rows_total = w_rows_total > 0 ? w_rows_total : Cout;
        bf16_convs.push_back(Bf16Conv{wname, rows_total, R * R, Cin, cin_pad, off});
        w16_total += (static_cast<size_t>(rows_total) * R * R * cin_pad + 63) / 64 * 64;
      }
      d.ldw = cin_pad;
      w_override = reinterpret_cast<const float*>(wblob16 + off + static_cast<size_t>(w_row_off) * R * R * cin_pad);
    }
    d.w = w_override ? w_override : W_(wname + ".w");
    d.Cout = Cout; d.R = R; d.S = R; d.stride = stride; d.pad = R / 2;
    d.b_batched = 0; if (!in_bf16) d.ldw = 0; d.act_cols = act_cols; d.dyn_w = dyn_w;
    d.bias = bias ? (bias_override ? bias_override : W_(wname + ".b")) : nullptr;
    d.residual = residual; d.ldr = ldo;
    d.act = act; d.alpha = 1.f; d.out = out; d.ldo = ldo;
    ConvGemmParams p;
    int rc = conv_gemm_build(d, &p);
    if (rc != MFR_OK) { build_rc = rc; return; }
    if (reserve_sms > 0 && conv_count < reserve_layers) p.grid_limit = conv_gemm_num_sms() - reserve_sms;
    ++conv_count;
    steps.push_back([p](cudaStream_t st) { return conv_gemm_launch(p, st); });
    StepInfo inf;
    inf.name = wname.empty() ? "qkv" : wname;
    inf.kind = 1;
    inf.dyn_rows = dyn_w;
    inf.dyn_images = N;
    const double per_pixel = 2.0 * Cout * Cin * R * R;
    inf.flops = dyn_w ? per_pixel : per_pixel * N * p.Ho * p.Wo;
    infos.push_back(inf);
  }

  // bf16 copy of a GEMM weight [rows][cin] in wblob16 (converted from the fp32 blob by the next forward)
  const unsigned short* w16(const std::string& wname, int rows, int cin) {
    for (const auto& e : bf16_convs)
      if (e.name == wname) return wblob16 + e.off;
    const size_t off = w16_total;
    bf16_convs.push_back(Bf16Conv{wname, rows, 1, cin, cin, off});
    w16_total += (static_cast<size_t>(rows) * cin + 63) / 64 * 64;
    return wblob16 + off;
  }

  void specs_backbone() {
    add_spec("stem.w", 128 * 49);
    add_spec("stem.b", 128);
    auto cv = [&](const std::string& n, int cin, int cout, int r, bool bias) {
      add_spec(n + ".w", static_cast<size_t>(cout) * r * r * cin);
      if (bias) add_spec(n + ".b", cout);
    };
    int cin = 128;
    const int dims[3] = {128, 196, 256};
    for (int li = 0; li < 3; ++li) {
      const int c = dims[li];
      for (int bi = 0; bi < 2; ++bi) {
        const std::string p = "layer" + std::to_string(li + 1) + "." + std::to_string(bi);
        cv(p + ".conv1", bi == 0 ? cin : c, c, 3, true);
        cv(p + ".conv2", c, c, 3, true);
        if (bi == 0 && li > 0) cv(p + ".down", cin, c, 1, true);
      }
      cin = c;
    }
    cv("layer3_outconv", 256, 256, 1, false);
    cv("layer2_outconv", 196, 256, 1, false);
    cv("layer2_outconv2.0", 256, 256, 3, true);
    cv("layer2_outconv2.3", 256, 196, 3, false);
    cv("layer1_outconv", 128, 196, 1, false);
    cv("layer1_outconv2.0", 196, 196, 3, true);
    cv("layer1_outconv2.3", 196, 128, 3, false);
    add_spec("pos_enc", static_cast<size_t>(L) * 256);
  }
  void specs_transformer(const std::string& pre, int C, int layers) {
    for (int i = 0; i < layers; ++i) {
      const std::string p = pre + "." + std::to_string(i);
      add_spec(p + ".qkv.w", static_cast<size_t>(3) * C * C);
      add_spec(p + ".merge.w", static_cast<size_t>(C) * C);
      add_spec(p + ".mlp0.w", static_cast<size_t>(4) * C * C);
      add_spec(p + ".mlp2.w", static_cast<size_t>(2) * C * C);
      add_spec(p + ".norm1.g", C);
      add_spec(p + ".norm1.b", C);
      add_spec(p + ".norm2.g", C);
      add_spec(p + ".norm2.b", C);
    }
  }

  // One LoFTREncoderLayer (transformer.py:35-58) on `n` "images" of `rows` tokens each.
  //   x, src : token buffers with row stride 2C (cols [0,C) = features, [C,2C) = message scratch)
  //   q/k/v live in `qkv_x` / `qkv_s` (row stride 3C) at the rows of the x / src images.
  void encoder_layer(const std::string& p, int C, float* x, float* src, float* qkv_x, float* qkv_s,
                     float* msg_x, float* mrg_x, float* h1_x, float* h2_x, int n, int rows,
                     const int* dyn_rows, const int* dyn_items, bool windows) {
    float* wq = W_(p + ".qkv.w");
    const long long ld2 = 2 * C, ld3 = 3 * C;
    // bf16 mode: the big intermediates (q|k|v, attention message, MLP hidden) are stored in bf16 — these GEMMs and
    // the attention kernels are HBM-bound — while the residual stream, LayerNorm inputs and outputs stay fp32.
    // The buffers are float-typed: `tb` bf16 elements per float when converting element offsets.
    const int tbf = precision == 1 ? 1 : 0, tb = tbf ? 2 : 1;
    auto E = [tb](float* ptr, size_t elems) { return ptr + elems / tb; };
    // bf16 mode: the GEMMs that read the token stream (q|k|v, mlp0) take their A operand from its bf16 shadow
    // (written by the LayerNorms below): half the operand bytes and kind::f16 MMAs instead of kind::tf32.
    float* xb = tbf ? shadow(x) : x;
    float* srcb = tbf ? shadow(src) : src;
    const std::string qn = p + ".qkv";
    if (windows && tbf && fused_fine_attn) {
      // fine level, bf16 mode: projection + window attention in one kernel (q, k, v stay on the SM; loftr_ops.cu)
      const unsigned short* wq16 = w16(qn, 3 * C, C);
      const int items_cap = rows / 25;
      for (int side = 0; side < n; ++side) {
        const size_t off2 = static_cast<size_t>(side) * rows * ld2, off1 = static_cast<size_t>(side) * rows * C;
        const float *xq_ = E(xb, off2), *xs_ = E(srcb, off2);
        float* mp = E(msg_x, off1);
        push("fine_attn_fused", [=](cudaStream_t st) {
          return fine_attention_fused(xq_, ld2, xs_, ld2, wq16, mp, C, dyn_items, items_cap, 25, st);
        });
      }
    } else {
    const bool raw = tbf && elu_in_attn;            // feature map applied by the attention kernels (bf16 = 2 below)
    const int qk_act = raw ? ACT_NONE : ACT_ELU1, att_mode = raw ? 2 : tbf;
    if (x == src) {
      conv(qn, xb, n, 1, rows, C, ld2, 3 * C, 1, 1, false, nullptr, qk_act, qkv_x, ld3, 2 * C, dyn_rows, wq, nullptr, tbf, tbf,
           3 * C, 0);
    } else {
      conv(qn, xb, n, 1, rows, C, ld2, C, 1, 1, false, nullptr, qk_act, qkv_x, ld3, C, dyn_rows, wq, nullptr, tbf, tbf, 3 * C, 0);
      conv(qn, srcb, n, 1, rows, C, ld2, 2 * C, 1, 1, false, nullptr, qk_act, E(qkv_s, C), ld3, C, dyn_rows,
           wq + static_cast<size_t>(C) * C, nullptr, tbf, tbf, 3 * C, C);
    }
    if (!windows) {
      float* kvp_ = kvp; float* kv_ = kv;
      const int Lr = rows, sk = splits_kv * NI / n;   // same partial-buffer footprint for n = B or n = 2B images
      float *kp = E(qkv_s, C), *vp = E(qkv_s, 2 * C);
      push("attn_long", [=](cudaStream_t st) {
        return linear_attention_long(qkv_x, ld3, kp, ld3, vp, ld3, msg_x, C, kvp_, kv_, n, Lr, Lr, 8, 32, sk, st, att_mode);
      }, 3);
    } else {
      const int items_cap = rows / 25;
      for (int side = 0; side < n; ++side) {
        const size_t off3 = static_cast<size_t>(side) * rows * ld3;
        const size_t off1 = static_cast<size_t>(side) * rows * C;
        float *qp = E(qkv_x, off3), *kp = E(qkv_s, off3 + C), *vp = E(qkv_s, off3 + 2 * C), *mp = E(msg_x, off1);
        push("attn_windows", [=](cudaStream_t st) {
          return linear_attention_windows(qp, ld3, kp, ld3, vp, ld3, mp, C, dyn_items, items_cap, 25, st, att_mode);
        });
      }
    }
    }
    const float *g1 = W_(p + ".norm1.g"), *b1 = W_(p + ".norm1.b");
    const float *g2 = W_(p + ".norm2.g"), *b2 = W_(p + ".norm2.b");
    if (tbf && fused_tail) {
      // merge -> norm1 -> mlp([x | message]) -> norm2 -> residual as ONE kernel: the merge output, the norm1 output, the
      // 2C-wide hidden tensor and the mlp output never leave the SM (encoder_tail.cu)
      EncoderTailDesc d{};
      d.msg = msg_x; d.ld_msg = C; d.xb_in = xb; d.ld_xb = ld2;
      d.w_merge = w16(p + ".merge", C, C); d.w_mlp0 = w16(p + ".mlp0", 2 * C, 2 * C); d.w_mlp2 = w16(p + ".mlp2", C, 2 * C);
      d.g1 = g1; d.b1 = b1; d.g2 = g2; d.b2 = b2; d.x = x; d.ldx = ld2; d.C = C;
      if (dyn_rows) { d.n_img = n; d.rows_cap = rows; d.dyn_rows = dyn_rows; }
      else { d.n_img = 1; d.rows_cap = static_cast<long long>(n) * rows; d.dyn_rows = nullptr; }
      EncoderTailParams tp;
      int rc = encoder_tail_build(d, &tp);
      if (rc != MFR_OK) { build_rc = rc; return; }
      steps.push_back([tp](cudaStream_t st) { return encoder_tail_launch(tp, st); });
      StepInfo inf;
      inf.name = p + ".tail";
      inf.kind = 1;
      inf.dyn_rows = dyn_rows;
      inf.dyn_images = n;
      const double per_row = 14.0 * C * C;      // 2 * (C*C + 2C*2C + 2C*C)
      inf.flops = dyn_rows ? per_row : per_row * n * rows;
      infos.push_back(inf);
      return;
    }
    conv(p + ".merge", msg_x, n, 1, rows, C, C, C, 1, 1, false, nullptr, ACT_NONE, mrg_x, C, 0, dyn_rows, nullptr, nullptr,
         tbf, 0);
    const long long tot = static_cast<long long>(n) * rows;
    // norm1(message) is only read by mlp0: in bf16 mode it goes straight to the shadow's message half
    if (dyn_rows == nullptr) {
      push("layernorm", [=](cudaStream_t st) {
        if (tbf) return layernorm_rows(mrg_x, C, g1, b1, nullptr, 0, nullptr, 0, tot, C, nullptr, st, E(xb, C), ld2);
        return layernorm_rows(mrg_x, C, g1, b1, nullptr, 0, x + C, ld2, tot, C, nullptr, st);
      });
    } else {
      for (int side = 0; side < n; ++side) {
        const size_t o1 = static_cast<size_t>(side) * rows * C, o2 = static_cast<size_t>(side) * rows * ld2;
        push("layernorm", [=](cudaStream_t st) {
          if (tbf) return layernorm_rows(mrg_x + o1, C, g1, b1, nullptr, 0, nullptr, 0, rows, C, dyn_rows, st, E(xb, o2 + C), ld2);
          return layernorm_rows(mrg_x + o1, C, g1, b1, nullptr, 0, x + o2 + C, ld2, rows, C, dyn_rows, st);
        });
      }
    }
    conv(p + ".mlp0", xb, n, 1, rows, 2 * C, ld2, 2 * C, 1, 1, false, nullptr, ACT_RELU, h1_x, ld2, 0, dyn_rows, nullptr,
         nullptr, tbf, tbf);
    conv(p + ".mlp2", h1_x, n, 1, rows, 2 * C, ld2, C, 1, 1, false, nullptr, ACT_NONE, h2_x, C, 0, dyn_rows, nullptr,
         nullptr, tbf, 0);
    if (dyn_rows == nullptr) {
      push("layernorm", [=](cudaStream_t st) {
        return layernorm_rows(h2_x, C, g2, b2, x, ld2, x, ld2, tot, C, nullptr, st, tbf ? xb : nullptr, ld2);
      });
    } else {
      for (int side = 0; side < n; ++side) {
        const size_t o1 = static_cast<size_t>(side) * rows * C, o2 = static_cast<size_t>(side) * rows * ld2;
        push("layernorm", [=](cudaStream_t st) {
          return layernorm_rows(h2_x + o1, C, g2, b2, x + o2, ld2, x + o2, ld2, rows, C, dyn_rows, st,
                                tbf ? E(xb, o2) : nullptr, ld2);
        });
      }
    }
  }

  int build() {
    NI = 2 * B;
    H2 = H / 2; W2 = W / 2; H4 = H / 4; W4 = W / 4; H8 = H / 8; W8 = W / 8;
    L = H8 * W8;
    Mcap = B * L;
    specs_backbone();
    specs_transformer("coarse", 256, 8);
    add_spec("fine.down_proj.w", 128 * 256);
    add_spec("fine.down_proj.b", 128);
    add_spec("fine.merge_feat.w", 128 * 256);
    add_spec("fine.merge_feat.b", 128);
    specs_transformer("fine", 128, 2);
    wblob = alloc<float>(wtotal, "weights");

    const size_t px2 = static_cast<size_t>(NI) * H2 * W2, px4 = static_cast<size_t>(NI) * H4 * W4,
                 px8 = static_cast<size_t>(NI) * H8 * W8;
    img = alloc<float>(static_cast<size_t>(NI) * H * W, "img");
    const bool bf = precision == 1;
    const int c196 = bf ? 200 : 196;                  // pixel stride of the 196-channel maps (16-byte rows in bf16)
    const size_t esz = bf ? 2 : 4;
    auto balloc = [&](size_t elems, const char* name) {  // backbone activation buffer (fp32 or bf16 storage)
      return reinterpret_cast<float*>(alloc<unsigned char>(elems * esz, name));
    };
    // the bf16 weight blob must exist before the conv steps are built (its size is known only afterwards:
    // reserve the upper bound = all backbone conv weights)
    if (bf) wblob16 = alloc<unsigned short>(kW16Capacity, "weights_bf16");
    if (precision == 2) {
      wblob3 = alloc<float>(wtotal * 3 + 64 * specs.size(), "weights_split3");
      const size_t frows_ = static_cast<size_t>(2) * Mcap * 25;
      x3_cap = 3 * std::max(std::max(px2 * 196, px4 * 256), std::max(px8 * 512, frows_ * 256));
      x3 = alloc<float>(x3_cap, "split3_scratch");
    }
    a0 = balloc(px2 * 128, "a0"); a1 = balloc(px2 * 128, "a1"); a2 = balloc(px2 * 128, "a2");
    b0 = balloc(px4 * c196, "b0"); b1 = balloc(px4 * c196, "b1"); b2 = balloc(px4 * c196, "b2");
    c0 = balloc(px8 * 256, "c0"); c1 = balloc(px8 * 256, "c1"); c2 = balloc(px8 * 256, "c2");
    fc = alloc<float>(px8 * 256, "feat_c_backbone");
    p4a = balloc(px4 * 256, "p4a"); p4b = balloc(px4 * 256, "p4b"); p4c = balloc(px4 * c196, "p4c");
    p2a = balloc(px2 * c196, "p2a"); p2b = balloc(px2 * c196, "p2b");
    float* featf = alloc<float>(px2 * 128, "feat_f");
    tok = alloc<float>(px8 * 512, "tok");
    if (bf) tokb = reinterpret_cast<float*>(alloc<unsigned short>(px8 * 512, "tok_bf16"));
    qkv = alloc<float>(px8 * 768, "qkv"); msg = alloc<float>(px8 * 256, "msg"); mrg = alloc<float>(px8 * 256, "mrg");
    h1 = alloc<float>(px8 * 512, "h1"); h2 = alloc<float>(px8 * 256, "h2");
    splits_kv = std::max(1, std::min(L / 16, (2 * 148 + NI - 1) / NI));   // ~2 blocks per SM when all images run
    kvp = alloc<float>(static_cast<size_t>(NI) * 8 * splits_kv * 1056, "kvp");
    kv = alloc<float>(static_cast<size_t>(NI) * 8 * 1056, "kv");
    S = alloc<float>(static_cast<size_t>(B) * L * L, "sim");
    const size_t BL = static_cast<size_t>(B) * L;
    cm.row_max = alloc<float>(BL); cm.row_rsum = alloc<float>(BL);
    cm.col_max = alloc<float>(BL); cm.col_rsum = alloc<float>(BL);
    splits_col = coarse_match_splits(B, L);
    cm.col_part = alloc<float>(BL * splits_col * 2);
    cm.rconf_max = alloc<float>(BL); cm.rconf_arg = alloc<int>(BL); cm.rconf_ties = alloc<int>(BL);
    cm.cconf_max = alloc<float>(BL); cm.jsel = alloc<int>(BL);
    cm.m_total = alloc<int>(4); cm.m_offsets = alloc<int>(B + 1);
    cm.b_ids = alloc<int>(BL); cm.i_ids = alloc<int>(BL); cm.j_ids = alloc<int>(BL);
    cm.mconf = alloc<float>(BL, "mconf"); cm.mkpts0_c = alloc<float>(BL * 2, "mkpts0_c");
    cm.mkpts1_c = alloc<float>(BL * 2, "mkpts1_c");
    const size_t frows = static_cast<size_t>(2) * Mcap * 25;
    cin = alloc<float>(static_cast<size_t>(2) * Mcap * 256, "fine_cin");
    cproj = alloc<float>(static_cast<size_t>(2) * Mcap * 128, "fine_cproj");
    win = alloc<float>(frows * 256, "fine_win");
    ftok = alloc<float>(frows * 256, "fine_tok");
    ftok_elems = frows * 256;
    if (bf) ftokb = reinterpret_cast<float*>(alloc<unsigned short>(frows * 256, "fine_tok_bf16"));
    fqkv = alloc<float>(frows * 384, "fine_qkv"); fmsg = alloc<float>(frows * 128); fmrg = alloc<float>(frows * 128);
    fh1 = alloc<float>(frows * 256); fh2 = alloc<float>(frows * 128);
    mk1f = alloc<float>(BL * 2, "mkpts1_f"); expec = alloc<float>(BL * 3, "expec_f");
    if (build_rc != MFR_OK) return build_rc;

    // ---------------------------------------------------------------- backbone (resnet_fpn.py:100-119)
    const int B16 = bf ? 1 : 0;
    {
      float *img_ = img, *a0_ = a0; float* ws = W_("stem.w"); float* bs = W_("stem.b");
      const int NI_ = NI, H_ = H, W_w = W, ob = B16, exact = precision == 2 ? 1 : 0;
      const int stem_limit = reserve_sms > 0 ? conv_gemm_num_sms() - reserve_sms : 0;
      push("stem", [=](cudaStream_t st) { return stem_conv7x7(img_, ws, bs, a0_, NI_, H_, W_w, st, ob, exact, stem_limit); });
    }
    auto bconv = [&](const std::string& name, const float* x, int Hh, int Ww, int Cin, int ldx, int Cout, int R,
                     int stride, bool bias, const float* res, int act, float* out, int ldo, int out16) {
      conv(name, x, NI, Hh, Ww, Cin, ldx, Cout, R, stride, bias, res, act, out, ldo, 0, nullptr, nullptr, nullptr, B16,
           out16);
    };
    bconv("layer1.0.conv1", a0, H2, W2, 128, 128, 128, 3, 1, true, nullptr, ACT_RELU, a1, 128, B16);
    bconv("layer1.0.conv2", a1, H2, W2, 128, 128, 128, 3, 1, true, a0, ACT_RELU, a2, 128, B16);
    bconv("layer1.1.conv1", a2, H2, W2, 128, 128, 128, 3, 1, true, nullptr, ACT_RELU, a1, 128, B16);
    bconv("layer1.1.conv2", a1, H2, W2, 128, 128, 128, 3, 1, true, a2, ACT_RELU, a0, 128, B16);  // x1 = a0
    bconv("layer2.0.conv1", a0, H2, W2, 128, 128, 196, 3, 2, true, nullptr, ACT_RELU, b1, c196, B16);
    bconv("layer2.0.down", a0, H2, W2, 128, 128, 196, 1, 2, true, nullptr, ACT_NONE, b2, c196, B16);
    bconv("layer2.0.conv2", b1, H4, W4, 196, c196, 196, 3, 1, true, b2, ACT_RELU, b0, c196, B16);
    bconv("layer2.1.conv1", b0, H4, W4, 196, c196, 196, 3, 1, true, nullptr, ACT_RELU, b1, c196, B16);
    bconv("layer2.1.conv2", b1, H4, W4, 196, c196, 196, 3, 1, true, b0, ACT_RELU, b2, c196, B16);  // x2 = b2
    bconv("layer3.0.conv1", b2, H4, W4, 196, c196, 256, 3, 2, true, nullptr, ACT_RELU, c1, 256, B16);
    bconv("layer3.0.down", b2, H4, W4, 196, c196, 256, 1, 2, true, nullptr, ACT_NONE, c2, 256, B16);
    bconv("layer3.0.conv2", c1, H8, W8, 256, 256, 256, 3, 1, true, c2, ACT_RELU, c0, 256, B16);
    bconv("layer3.1.conv1", c0, H8, W8, 256, 256, 256, 3, 1, true, nullptr, ACT_RELU, c1, 256, B16);
    bconv("layer3.1.conv2", c1, H8, W8, 256, 256, 256, 3, 1, true, c0, ACT_RELU, c2, 256, B16);  // x3 = c2
    bconv("layer3_outconv", c2, H8, W8, 256, 256, 256, 1, 1, false, nullptr, ACT_NONE, fc, 256, 0);   // fp32 out
    bconv("layer2_outconv", b2, H4, W4, 196, c196, 256, 1, 1, false, nullptr, ACT_NONE, p4a, 256, B16);
    {
      float *s_ = fc, *d_ = p4a; const int n = NI, hs = H8, wsz = W8, ob = B16;
      push("upsample", [=](cudaStream_t st) { return upsample2x_add(s_, d_, n, hs, wsz, 256, 256, 256, 0, ob, st); });
    }
    bconv("layer2_outconv2.0", p4a, H4, W4, 256, 256, 256, 3, 1, true, nullptr, ACT_LEAKY, p4b, 256, B16);
    bconv("layer2_outconv2.3", p4b, H4, W4, 256, 256, 196, 3, 1, false, nullptr, ACT_NONE, p4c, c196, B16);
    bconv("layer1_outconv", a0, H2, W2, 128, 128, 196, 1, 1, false, nullptr, ACT_NONE, p2a, c196, B16);
    {
      float *s_ = p4c, *d_ = p2a; const int n = NI, hs = H4, wsz = W4, ld = c196, ob = B16;
      push("upsample", [=](cudaStream_t st) { return upsample2x_add(s_, d_, n, hs, wsz, 196, ld, ld, ob, ob, st); });
    }
    bconv("layer1_outconv2.0", p2a, H2, W2, 196, c196, 196, 3, 1, true, nullptr, ACT_LEAKY, p2b, c196, B16);
    bconv("layer1_outconv2.3", p2b, H2, W2, 196, c196, 128, 3, 1, false, nullptr, ACT_NONE, featf, 128, 0);   // fp32 out

    // ---------------------------------------------------------------- pos-enc + coarse transformer
    {
      float *f = fc, *t = tok; float* pe = W_("pos_enc"); const int n = NI, l = L;
      push("posenc", [=](cudaStream_t st) { return posenc_tokens(f, pe, t, 512, n, l, 256, st); });
      if (bf) {
        float* tb_ = tokb;
        push("to_bf16", [=](cudaStream_t st) { return rows_to_bf16(t, 512, tb_, 512, static_cast<long long>(n) * l, 256, nullptr, st); });
      }
    }
    const size_t half_tok = static_cast<size_t>(B) * L * 512, half_qkv = static_cast<size_t>(B) * L * 768;
    const size_t half_c = static_cast<size_t>(B) * L * 256, half_2c = static_cast<size_t>(B) * L * 512;
    for (int i = 0; i < 8; ++i) {
      const std::string p = "coarse." + std::to_string(i);
      if (i % 2 == 0) {  // self: both image sets at once
        encoder_layer(p, 256, tok, tok, qkv, qkv, msg, mrg, h1, h2, NI, L, nullptr, nullptr, false);
      } else {  // cross: feat0 <- (feat0, feat1), then feat1 <- (feat1, updated feat0)
        const int tb = precision == 1 ? 2 : 1;   // qkv / msg / h1 hold bf16 in that mode (see encoder_layer)
        encoder_layer(p, 256, tok, tok + half_tok, qkv, qkv + half_qkv / tb, msg, mrg, h1, h2, B, L, nullptr,
                      nullptr, false);
        encoder_layer(p, 256, tok + half_tok, tok, qkv + half_qkv / tb, qkv, msg + half_c / tb, mrg + half_c,
                      h1 + half_2c / tb, h2 + half_c, B, L, nullptr, nullptr, false);
      }
    }

    // ---------------------------------------------------------------- coarse matching
    {
      ConvGemmDesc d{};
      d.x = tok; d.N = B; d.H = 1; d.W = L; d.Cin = 256; d.ldx = 512;
      d.w = tok + half_tok; d.Cout = L; d.R = 1; d.S = 1; d.stride = 1; d.pad = 0;
      if (kCorrelationBf16 && precision == 1) { d.x = tokb; d.w = tokb + half_tok / 2; d.in_bf16 = 1; }
      if (precision == 2) {   // 3xTF32: X -> [hi | lo | hi], Y -> [hi | hi | lo], one GEMM over K = 768
        const size_t BLr = static_cast<size_t>(B) * L;
        float *xa = x3, *xb = x3 + BLr * 768; const float *t0 = tok, *t1 = tok + half_tok; const int B_ = B, L_ = L;
        push("split3", [=](cudaStream_t st) { return rows_split3(t0, 512, static_cast<long long>(L_) * 512, xa, L_, B_, 256, 0, nullptr, st); });
        push("split3", [=](cudaStream_t st) { return rows_split3(t1, 512, static_cast<long long>(L_) * 512, xb, L_, B_, 256, 1, nullptr, st); });
        d.x = xa; d.w = xb; d.Cin = 768; d.ldx = 768; d.ldw = 768;
      }
      d.b_batched = 1; if (precision != 2) d.ldw = 512; d.act = ACT_NONE; d.alpha = 1.f / (256.f * temperature);
      d.out = S; d.ldo = L; d.ldr = L;
      ConvGemmParams p;
      int rc = conv_gemm_build(d, &p);
      if (rc != MFR_OK) return rc;
      steps.push_back([p](cudaStream_t st) { return conv_gemm_launch(p, st); });
      infos.push_back(StepInfo{"correlation", 1, 2.0 * B * static_cast<double>(L) * L * 256, nullptr, 0});
      CoarseMatchBuffers cmb = cm; float* S_ = S;
      const int B_ = B, L_ = L, hc = H8, wc = W8, bd = border, sc = splits_col; const float th = thr;
      const float scale = static_cast<float>(H) / static_cast<float>(H8);
      push("coarse_match", [=](cudaStream_t st) {
        return coarse_match(S_, cmb, B_, L_, hc, wc, th, bd, scale, sc, st);
      }, 8);
    }

    // ---------------------------------------------------------------- fine level
    {
      CoarseMatchBuffers cmb = cm; float *tok_ = tok, *cin_ = cin, *cproj_ = cproj, *win_ = win;
      const int B_ = B, L_ = L, Mc = Mcap, Hf = H2, Wf = W2, wc = W8;
      push("fine_gather_coarse", [=](cudaStream_t st) {
        return fine_gather_coarse(tok_, 512, B_, L_, cmb.b_ids, cmb.i_ids, cmb.j_ids, cmb.m_total, Mc, cin_, st);
      });
      conv("fine.down_proj", cin, 2, 1, Mcap, 256, 256, 128, 1, 1, true, nullptr, ACT_NONE, cproj, 128, 0,
           cm.m_total);
      push("fine_gather_windows", [=](cudaStream_t st) {
        return fine_gather_windows(featf, B_, Hf, Wf, wc, cmb.b_ids, cmb.i_ids, cmb.j_ids, cmb.m_total, Mc,
                                   cproj_, win_, st);
      });
      conv("fine.merge_feat", win, 2, 1, Mcap * 25, 256, 256, 128, 1, 1, true, nullptr, ACT_NONE, ftok, 256, 0,
           cm.m_total + 2);
    }
    {
      const int rows = Mcap * 25;
      if (precision == 1) {   // bf16 shadow of the fine tokens (the encoder layers keep it current from here on)
        for (int side = 0; side < 2; ++side) {
          const float* src = ftok + static_cast<size_t>(side) * rows * 256;
          float* dst = ftokb + static_cast<size_t>(side) * rows * 256 / 2;
          const int* dr = cm.m_total + 2;
          push("to_bf16", [=](cudaStream_t st) { return rows_to_bf16(src, 256, dst, 256, rows, 128, dr, st); });
        }
      }
      const size_t s_tok = static_cast<size_t>(rows) * 256, s_qkv = static_cast<size_t>(rows) * 384,
                   s_c = static_cast<size_t>(rows) * 128, s_2c = static_cast<size_t>(rows) * 256;
      encoder_layer("fine.0", 128, ftok, ftok, fqkv, fqkv, fmsg, fmrg, fh1, fh2, 2, rows, cm.m_total + 2,
                    cm.m_total, true);
      const int tb = precision == 1 ? 2 : 1;   // qkv / msg / h1 hold bf16 in that mode (see encoder_layer)
      encoder_layer("fine.1", 128, ftok, ftok + s_tok, fqkv, fqkv + s_qkv / tb, fmsg, fmrg, fh1, fh2, 1, rows,
                    cm.m_total + 2, cm.m_total, true);
      encoder_layer("fine.1", 128, ftok + s_tok, ftok, fqkv + s_qkv / tb, fqkv, fmsg + s_c / tb, fmrg + s_c,
                    fh1 + s_2c / tb, fh2 + s_c, 1, rows, cm.m_total + 2, cm.m_total, true);
      CoarseMatchBuffers cmb = cm; float *f0 = ftok, *f1 = ftok + s_tok, *mk = mk1f, *ex = expec;
      const int Mc = Mcap; const float scale_f = static_cast<float>(H) / static_cast<float>(H2);
      push("fine_match", [=](cudaStream_t st) {
        return fine_match(f0, f1, 256, cmb.mkpts1_c, cmb.m_total, Mc, scale_f, mk, ex, st);
      });
    }
    if (precision == 2 && w16_total > wtotal * 3 + 64 * specs.size()) {
      fprintf(stderr, "[mfr_b200] split weight blob too small\n");
      return MFR_ERR_ARG;
    }
    if (bf && w16_total > kW16Capacity) {
      fprintf(stderr, "[mfr_b200] bf16 weight blob too small\n");
      return MFR_ERR_ARG;
    }
    return build_rc;
  }

  int run_steps(cudaStream_t st) {
    for (auto& s : steps) {
      int rc = s(st);
      if (rc != MFR_OK) return rc;
    }
    return MFR_OK;
  }

  // src images are [B][src_h][src_w] (src_h <= H, src_w <= W); the right/bottom remainder of the
  // engine's [H][W] frame stays zero: the reference's ConstantPad2d (matchers.py:41-46).
  int forward(const float* img0, const float* img1, int src_h, int src_w, cudaStream_t st) {
    if (!weights_loaded) {
      fprintf(stderr, "[mfr_b200] loftr forward before weights were loaded\n");
      return MFR_ERR_ARG;
    }
    if (src_h > H || src_w > W || src_h < 1 || src_w < 1) return MFR_ERR_ARG;
    if (precision == 1 && !w16_ready) {   // bf16 copies of the (folded) backbone conv weights
      for (const auto& bc : bf16_convs) {
        int rc = weights_to_bf16(W_(bc.name + ".w"), wblob16 + bc.off, static_cast<size_t>(bc.cout) * bc.taps, bc.cin,
                                 bc.cin_pad, st);
        if (rc != MFR_OK) return rc;
      }
      w16_ready = true;
    }
    if (precision == 2 && !w16_ready) {   // [hi | hi | lo] split of every GEMM weight
      for (const auto& bc : bf16_convs) {
        int rc = rows_split3(W_(bc.name + ".w"), bc.cin, 0, wblob3 + bc.off, static_cast<long long>(bc.cout) * bc.taps, 1, bc.cin,
                             1, nullptr, st);
        if (rc != MFR_OK) return rc;
      }
      w16_ready = true;
    }
    const size_t half = static_cast<size_t>(B) * H * W;
    if ((src_h != H || src_w != W) && (src_h != last_src_h || src_w != last_src_w) && last_src_h > 0)
      MFR_CUDA_CHECK(cudaMemsetAsync(img, 0, 2 * half * 4, st));   // a smaller source than last time: clear the old pixels of the pad
    last_src_h = src_h; last_src_w = src_w;
    if (src_h == H && src_w == W) {
      MFR_CUDA_CHECK(cudaMemcpyAsync(img, img0, half * 4, cudaMemcpyDeviceToDevice, st));
      MFR_CUDA_CHECK(cudaMemcpyAsync(img + half, img1, half * 4, cudaMemcpyDeviceToDevice, st));
    } else {
      for (int b = 0; b < B; ++b) {
        MFR_CUDA_CHECK(cudaMemcpy2DAsync(img + static_cast<size_t>(b) * H * W, static_cast<size_t>(W) * 4,
                                         img0 + static_cast<size_t>(b) * src_h * src_w, static_cast<size_t>(src_w) * 4,
                                         static_cast<size_t>(src_w) * 4, src_h, cudaMemcpyDeviceToDevice, st));
        MFR_CUDA_CHECK(cudaMemcpy2DAsync(img + half + static_cast<size_t>(b) * H * W, static_cast<size_t>(W) * 4,
                                         img1 + static_cast<size_t>(b) * src_h * src_w, static_cast<size_t>(src_w) * 4,
                                         static_cast<size_t>(src_w) * 4, src_h, cudaMemcpyDeviceToDevice, st));
      }
    }
    if (!use_graph) return run_steps(st);
    if (!graph_exec) {
      // warm (sets function attributes outside capture), then capture the launch list once
      int rc = run_steps(st);
      if (rc != MFR_OK) return rc;
      MFR_CUDA_CHECK(cudaStreamSynchronize(st));
      cudaGraph_t g;
      MFR_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      rc = run_steps(st);
      cudaError_t e = cudaStreamEndCapture(st, &g);
      if (rc != MFR_OK || e != cudaSuccess) {
        fprintf(stderr, "[mfr_b200] graph capture failed (%d, %s); running eagerly\n", rc, cudaGetErrorString(e));
        use_graph = false;
        cudaGetLastError();
        return run_steps(st);
      }
      MFR_CUDA_CHECK(cudaGraphInstantiate(&graph_exec, g, 0));
      cudaGraphDestroy(g);
      return MFR_OK;  // the warm-up pass above already produced this call's results
    }
    MFR_CUDA_CHECK(cudaGraphLaunch(graph_exec, st));
    return MFR_OK;
  }

  // eager run with a CUDA event pair around every step; ms[i] = device time of step i
  int profile(const float* img0, const float* img1, int src_h, int src_w, cudaStream_t st, int max_steps,
              float* ms, double* flops, int* kinds, char* names) {
    const bool g = use_graph;
    use_graph = false;
    int rc = forward(img0, img1, src_h, src_w, st);  // warm
    if (rc != MFR_OK) { use_graph = g; return -rc; }
    const int n = static_cast<int>(steps.size());
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& e : ev) cudaEventCreate(&e);
    cudaEventRecord(ev[0], st);
    for (int i = 0; i < n; ++i) {
      rc = steps[i](st);
      if (rc != MFR_OK) break;
      cudaEventRecord(ev[i + 1], st);
    }
    cudaStreamSynchronize(st);
    int m_total_h[4] = {0, 0, 0, 0};
    cudaMemcpy(m_total_h, cm.m_total, sizeof(m_total_h), cudaMemcpyDeviceToHost);
    for (int i = 0; i < n && i < max_steps && rc == MFR_OK; ++i) {
      cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
      double f = infos[i].flops;
      if (infos[i].dyn_rows) {
        const int idx = static_cast<int>(infos[i].dyn_rows - cm.m_total);
        f *= static_cast<double>(m_total_h[idx]) * infos[i].dyn_images;
      }
      flops[i] = f;
      kinds[i] = infos[i].kind;
      strncpy(names + 32 * i, infos[i].name.c_str(), 31);
      names[32 * i + 31] = 0;
    }
    for (auto& e : ev) cudaEventDestroy(e);
    use_graph = g;
    return rc == MFR_OK ? n : -rc;
  }

  ~LoftrEngine() {
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    for (void* p : allocs) cudaFree(p);
  }
};

// ---------------------------------------------------------------------------------------------
int loftr_create(int B, int H, int W, float thr, int border, float temperature, int precision, LoftrEngine** out) {
  if (B < 1 || H % 8 || W % 8 || H < 32 || W < 32 || precision < 0 || precision > 2) return MFR_ERR_ARG;
  auto* e = new LoftrEngine();
  e->B = B; e->H = H; e->W = W; e->thr = thr; e->border = border; e->temperature = temperature;
  e->precision = precision;
  int rc = e->build();
  if (rc != MFR_OK) { delete e; return rc; }
  *out = e;
  return MFR_OK;
}
void loftr_destroy(LoftrEngine* e) { delete e; }
int loftr_num_weights(const LoftrEngine* e) { return static_cast<int>(e->specs.size()); }
const char* loftr_weight_name(const LoftrEngine* e, int i) { return e->specs[i].name.c_str(); }
long long loftr_weight_numel(const LoftrEngine* e, int i) { return static_cast<long long>(e->specs[i].numel); }
int loftr_set_weight(LoftrEngine* e, int i, const float* host, long long numel, cudaStream_t st) {
  if (i < 0 || i >= static_cast<int>(e->specs.size()) || numel != static_cast<long long>(e->specs[i].numel))
    return MFR_ERR_ARG;
  MFR_CUDA_CHECK(cudaMemcpyAsync(e->wblob + e->specs[i].offset, host, numel * 4, cudaMemcpyHostToDevice, st));
  MFR_CUDA_CHECK(cudaStreamSynchronize(st));
  e->w16_ready = false;                       // derived copies (bf16 / split operands) are rebuilt by the next forward
  if (e->weight_set.size() != e->specs.size()) e->weight_set.assign(e->specs.size(), 0);
  e->weight_set[i] = 1;
  bool all = true;
  for (char c : e->weight_set) all = all && c;
  e->weights_loaded = all;                    // every tensor of the state dict has arrived
  return MFR_OK;
}
void loftr_mark_weights_loaded(LoftrEngine* e) { e->weights_loaded = true; }
int loftr_forward(LoftrEngine* e, const float* img0, const float* img1, int src_h, int src_w, cudaStream_t st) {
  return e->forward(img0, img1, src_h, src_w, st);
}
void loftr_set_graph(LoftrEngine* e, int on) { e->use_graph = on != 0; }
LoftrOutputs loftr_outputs(LoftrEngine* e) {
  LoftrOutputs o;
  o.m_total = e->cm.m_total; o.m_offsets = e->cm.m_offsets; o.b_ids = e->cm.b_ids;
  o.i_ids = e->cm.i_ids; o.j_ids = e->cm.j_ids; o.mconf = e->cm.mconf;
  o.mkpts0_f = e->cm.mkpts0_c; o.mkpts1_f = e->mk1f; o.mkpts1_c = e->cm.mkpts1_c; o.expec_f = e->expec;
  o.capacity = e->Mcap;
  return o;
}
int loftr_debug_tensor(LoftrEngine* e, const char* name, float** ptr, long long* numel) {
  auto it = e->dbg.find(name);
  if (it == e->dbg.end()) return MFR_ERR_ARG;
  *ptr = it->second.ptr;
  *numel = it->second.numel;
  return MFR_OK;
}
int loftr_num_steps(const LoftrEngine* e) { return static_cast<int>(e->steps.size()); }
int loftr_num_launches(const LoftrEngine* e) {
  int n = 0;
  for (const auto& i : e->infos) n += i.launches;
  return n;
}
int loftr_profile(LoftrEngine* e, const float* img0, const float* img1, int src_h, int src_w, cudaStream_t st,
                  int max_steps, float* ms, double* flops, int* kinds, char* names) {
  return e->profile(img0, img1, src_h, src_w, st, max_steps, ms, flops, kinds, names);
}

}  // namespace mfr
