// SuperPoint + SuperGlue matcher engine (B image pairs per call): static launch list of tcgen05
// conv/GEMM launches + SIMT kernels, replayed through a CUDA graph. Mirrors the reference's
// SuperGlue/models/matching.py:56-84 (SuperPoint x2 -> SuperGlue) with the settings hard-wired in
// etc/feature_matching_baselines/matchers.py:65-71 (nms 4, threshold 0.005, <= 1024 keypoints,
// 20 Sinkhorn iterations, match threshold 0.2). Images are stored pair-interleaved
// (image 2p = pair p's image0, 2p+1 = image1) so that the cross-attention partner is `img ^ 1`.
#include "spsg_engine.cuh"

#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "conv_gemm.cuh"
#include "loftr_ops.cuh"   // rows_split3 (3xTF32 operand split of the fp32-faithful mode)
#include "spsg_ops.cuh"

namespace mfr {

struct SpsgEngine {
  int B, H, W, NI, K;  // K = max keypoints per image
  float kp_thr, match_thr;
  int nms_r, border, sk_iters;
  struct Spec { std::string name; size_t numel, offset; };
  std::vector<Spec> specs;
  std::map<std::string, size_t> idx;
  float* wblob = nullptr;
  size_t wtotal = 0;
  bool loaded = false;
  float bin_score = 1.f;
  std::vector<void*> allocs;
  std::map<std::string, std::pair<float*, long long>> dbg;
  std::vector<std::function<int(cudaStream_t)>> steps;
  int launches = 0;
  size_t sg_begin = 0;   // first step of the SuperGlue part
  int rc = MFR_OK;
  cudaGraphExec_t graph = nullptr;
  bool use_graph = true;
  // buffers
  float* img;
  float *kpts, *kscores, *X, *D, *mk0, *mk1, *mscores0;
  int *counts, *matches0, *m_offsets;
  // precision 2 (fp32-faithful, MFR_PRECISION_FP32X3): every tensor-core GEMM runs as a 3xTF32 split product over 3x the
  // K extent; A operands are split [hi | lo | hi] into the scratch x3 right before the GEMM, weights [hi | hi | lo] into
  // wblob3 once per upload (see loftr_ops.cu: rows_split3)
  int precision = 0;
  float* x3 = nullptr; size_t x3_cap = 0;
  float* wblob3 = nullptr;
  bool w3_ready = false;
  struct W3 { std::string name; long long rows; int cin; };
  std::vector<W3> w3_list;

  void add(const std::string& n, size_t numel) {
    idx[n] = specs.size();
    specs.push_back({n, numel, wtotal});
    wtotal += (numel + 63) / 64 * 64;
  }
  float* Wp(const std::string& n) {
    auto it = idx.find(n);
    if (it == idx.end()) { fprintf(stderr, "[mfr_b200] spsg: unknown weight %s\n", n.c_str()); rc = MFR_ERR_ARG; return wblob; }
    return wblob + specs[it->second].offset;
  }
  template <typename T>
  T* alloc(size_t n, const char* name = nullptr) {
    void* p = nullptr;
    if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) { rc = MFR_ERR_CUDA; return nullptr; }
    cudaMemset(p, 0, n * sizeof(T));
    allocs.push_back(p);
    if (name) dbg[name] = {reinterpret_cast<float*>(p), static_cast<long long>(n)};
    return reinterpret_cast<T*>(p);
  }
  void push(std::function<int(cudaStream_t)> f, int n_launch = 1) { steps.push_back(std::move(f)); launches += n_launch; }

  void gemm(const std::string& w, const float* x, int N, int Hh, int Ww, int Cin, long long ldx, int Cout, int R,
            int act, const float* residual, float* out, long long ldo, long long ldr = 0) {
    ConvGemmDesc d{};
    d.x = x; d.N = N; d.H = Hh; d.W = Ww; d.Cin = Cin; d.ldx = ldx;
    d.w = Wp(w + ".w");
    if (precision == 2) {
      const long long rows_img = static_cast<long long>(Hh) * Ww;
      if (static_cast<size_t>(N) * rows_img * 3 * Cin > x3_cap || (Cin & 3)) { rc = MFR_ERR_ARG; return; }
      float* x3_ = x3; const int N_ = N, C_ = Cin;
      push([=](cudaStream_t st) { return rows_split3(x, ldx, rows_img * ldx, x3_, rows_img, N_, C_, 0, nullptr, st); });
      bool found = false;
      for (const auto& e : w3_list) found = found || e.name == w;
      if (!found) w3_list.push_back(W3{w, static_cast<long long>(Cout) * R * R, Cin});
      d.x = x3; d.Cin = 3 * Cin; d.ldx = 3 * Cin;
      d.w = wblob3 + 3 * specs[idx[w + ".w"]].offset;
    }
    d.Cout = Cout; d.R = R; d.S = R; d.stride = 1; d.pad = R / 2;
    d.bias = Wp(w + ".b"); d.residual = residual; d.ldr = ldr > 0 ? ldr : ldo; d.act = act; d.alpha = 1.f; d.out = out; d.ldo = ldo;
    ConvGemmParams p;
    int r = conv_gemm_build(d, &p);
    if (r != MFR_OK) { rc = r; return; }
    push([p](cudaStream_t st) { return conv_gemm_launch(p, st); });
  }

  int build() {
    NI = 2 * B;
    // ---- weight table
    add("conv1a.w", 64 * 9); add("conv1a.b", 64);
    auto cv = [&](const std::string& n, int ci, int co, int r) { add(n + ".w", static_cast<size_t>(co) * r * r * ci); add(n + ".b", co); };
    cv("conv1b", 64, 64, 3); cv("conv2a", 64, 64, 3); cv("conv2b", 64, 64, 3); cv("conv3a", 64, 128, 3);
    cv("conv3b", 128, 128, 3); cv("conv4a", 128, 128, 3); cv("conv4b", 128, 128, 3);
    cv("convPa", 128, 256, 3); cv("convPb", 256, 65, 1); cv("convDa", 128, 256, 3); cv("convDb", 256, 256, 1);
    cv("kenc.0", 4, 32, 1); cv("kenc.1", 32, 64, 1); cv("kenc.2", 64, 128, 1); cv("kenc.3", 128, 256, 1); cv("kenc.4", 256, 256, 1);
    for (int i = 0; i < 18; ++i) {
      const std::string p = "gnn." + std::to_string(i);
      cv(p + ".qkv", 256, 768, 1); cv(p + ".merge", 256, 256, 1); cv(p + ".mlp0", 512, 512, 1); cv(p + ".mlp3", 512, 256, 1);
    }
    cv("final_proj", 256, 256, 1);
    add("bin_score", 1);
    wblob = alloc<float>(wtotal, "weights");
    if (precision == 2) {
      wblob3 = alloc<float>(3 * wtotal, "weights_split3");
      const size_t act = static_cast<size_t>(2) * B * H * W * 64;                         // largest conv input (conv1b)
      const size_t att = static_cast<size_t>(2) * B * (4 * static_cast<size_t>(K) * K + 256 * static_cast<size_t>(K));   // P + V^T
      x3_cap = 3 * std::max(act, std::max(att, static_cast<size_t>(2) * B * K * 768));
      x3 = alloc<float>(x3_cap, "split3_scratch");
    }

    const int H2 = H / 2, W2 = W / 2, H4 = H2 / 2, W4 = W2 / 2, h8 = H4 / 2, w8 = W4 / 2;
    const int Hs = h8 * 8, Ws = w8 * 8;
    auto px = [&](int hh, int ww) { return static_cast<size_t>(NI) * hh * ww; };
    img = alloc<float>(px(H, W), "img");
    float* x1a = alloc<float>(px(H, W) * 64, "x1a"); float* x1b = alloc<float>(px(H, W) * 64);
    float* p1 = alloc<float>(px(H2, W2) * 64); float* x2a = alloc<float>(px(H2, W2) * 64); float* x2b = alloc<float>(px(H2, W2) * 64);
    float* p2 = alloc<float>(px(H4, W4) * 64); float* x3a = alloc<float>(px(H4, W4) * 128); float* x3b = alloc<float>(px(H4, W4) * 128);
    float* p3 = alloc<float>(px(h8, w8) * 128); float* x4a = alloc<float>(px(h8, w8) * 128); float* x4b = alloc<float>(px(h8, w8) * 128, "x4b");
    float* cPa = alloc<float>(px(h8, w8) * 256); float* logits = alloc<float>(px(h8, w8) * 65, "logits");
    float* scores = alloc<float>(px(Hs, Ws), "scores"); float* mask = alloc<float>(px(Hs, Ws), "nms_mask");
    float* t0 = alloc<float>(px(Hs, Ws)); float* t1 = alloc<float>(px(Hs, Ws)); float* t2 = alloc<float>(px(Hs, Ws)); float* t3 = alloc<float>(px(Hs, Ws));
    float* cDa = alloc<float>(px(h8, w8) * 256); float* dense = alloc<float>(px(h8, w8) * 256, "dense_desc");
    kpts = alloc<float>(static_cast<size_t>(NI) * K * 2, "keypoints"); kscores = alloc<float>(static_cast<size_t>(NI) * K, "kscores");
    counts = alloc<int>(NI);
    X = alloc<float>(static_cast<size_t>(NI) * K * 512, "tokens");
    D = alloc<float>(static_cast<size_t>(NI) * K * 256, "descriptors");
    const size_t rows = static_cast<size_t>(NI) * K;
    float* kin = alloc<float>(rows * 4); float* k1 = alloc<float>(rows * 32); float* k2 = alloc<float>(rows * 64);
    float* k3 = alloc<float>(rows * 128); float* k4 = alloc<float>(rows * 256);
    float* qkv = alloc<float>(rows * 768, "qkv"); float* msg = alloc<float>(rows * 256, "msg"); float* h1 = alloc<float>(rows * 512);
    float* Mb = alloc<float>(rows * 256, "mdesc");
    // tensor-core attention: scores / probabilities [NI][4 heads][K][K] and V transposed [NI][256][K]
    float* Sat = alloc<float>(static_cast<size_t>(NI) * 4 * K * K, "attn_scores");
    float* Vt = alloc<float>(static_cast<size_t>(NI) * 256 * K, "attn_vt");
    float* S = alloc<float>(static_cast<size_t>(B) * K * K, "sg_scores");
    float* u = alloc<float>(static_cast<size_t>(B) * (K + 1), "sk_u"); float* v = alloc<float>(static_cast<size_t>(B) * (K + 1), "sk_v");
    float* rmax = alloc<float>(static_cast<size_t>(B) * K); int* rarg = alloc<int>(static_cast<size_t>(B) * K); int* carg = alloc<int>(static_cast<size_t>(B) * K);
    matches0 = alloc<int>(static_cast<size_t>(B) * K); mscores0 = alloc<float>(static_cast<size_t>(B) * K, "mscores0");
    mk0 = alloc<float>(static_cast<size_t>(B) * K * 2); mk1 = alloc<float>(static_cast<size_t>(B) * K * 2);
    m_offsets = alloc<int>(B + 1);
    if (rc != MFR_OK) return rc;

    // ---- SuperPoint (superpoint.py:145-196)
    {
      float *w = Wp("conv1a.w"), *b = Wp("conv1a.b"), *im = img; const int n = NI, hh = H, ww = W;
      push([=](cudaStream_t st) { return sp_conv1a(im, w, b, x1a, n, hh, ww, st); });
    }
    auto pool = [&](float* in, float* out, int hh, int ww, int c) {
      const int n = NI;
      push([=](cudaStream_t st) { return maxpool2x2(in, out, n, hh, ww, c, st); });
    };
    gemm("conv1b", x1a, NI, H, W, 64, 64, 64, 3, ACT_RELU, nullptr, x1b, 64);
    pool(x1b, p1, H, W, 64);
    gemm("conv2a", p1, NI, H2, W2, 64, 64, 64, 3, ACT_RELU, nullptr, x2a, 64);
    gemm("conv2b", x2a, NI, H2, W2, 64, 64, 64, 3, ACT_RELU, nullptr, x2b, 64);
    pool(x2b, p2, H2, W2, 64);
    gemm("conv3a", p2, NI, H4, W4, 64, 64, 128, 3, ACT_RELU, nullptr, x3a, 128);
    gemm("conv3b", x3a, NI, H4, W4, 128, 128, 128, 3, ACT_RELU, nullptr, x3b, 128);
    pool(x3b, p3, H4, W4, 128);
    gemm("conv4a", p3, NI, h8, w8, 128, 128, 128, 3, ACT_RELU, nullptr, x4a, 128);
    gemm("conv4b", x4a, NI, h8, w8, 128, 128, 128, 3, ACT_RELU, nullptr, x4b, 128);
    gemm("convPa", x4b, NI, h8, w8, 128, 128, 256, 3, ACT_RELU, nullptr, cPa, 256);
    gemm("convPb", cPa, NI, h8, w8, 256, 256, 65, 1, ACT_NONE, nullptr, logits, 65);
    {
      const int n = NI, hh = h8, ww = w8, r = nms_r, bd = border, kk = K; const float th = kp_thr;
      float* kp = kpts; float* ks = kscores; int* cn = counts;
      push([=](cudaStream_t st) { return sp_score_map(logits, 65, scores, n, hh, ww, st); });
      push([=](cudaStream_t st) { return sp_nms(scores, mask, t0, t1, t2, t3, n, hh * 8, ww * 8, r, st); }, 15);
      push([=](cudaStream_t st) { return sp_select_keypoints(scores, mask, n, hh * 8, ww * 8, th, bd, kk, kp, ks, cn, st); });
    }
    gemm("convDa", x4b, NI, h8, w8, 128, 128, 256, 3, ACT_RELU, nullptr, cDa, 256);
    gemm("convDb", cDa, NI, h8, w8, 256, 256, 256, 1, ACT_NONE, nullptr, dense, 256);
    {
      const int n = NI, hh = h8, ww = w8, kk = K; float* kp = kpts; int* cn = counts; float* Dp = D;
      const size_t rws = px(h8, w8);
      push([=](cudaStream_t st) { return l2norm256(dense, rws, st); });
      push([=](cudaStream_t st) { return sp_sample_descriptors(dense, kp, cn, n, hh, ww, kk, Dp, 256, st); });
    }
    // ---- SuperGlue (superglue.py:228-283). Token buffer X: [NI][K][512] = [desc | message]
    sg_begin = steps.size();
    {
      float* kp = kpts; float* ks = kscores; const int tot = NI * K, hh = H, ww = W;
      push([=](cudaStream_t st) { return sg_kenc_input(kp, ks, kin, tot, hh, ww, st); });
    }
    gemm("kenc.0", kin, NI, 1, K, 4, 4, 32, 1, ACT_RELU, nullptr, k1, 32);
    gemm("kenc.1", k1, NI, 1, K, 32, 32, 64, 1, ACT_RELU, nullptr, k2, 64);
    gemm("kenc.2", k2, NI, 1, K, 64, 64, 128, 1, ACT_RELU, nullptr, k3, 128);
    gemm("kenc.3", k3, NI, 1, K, 128, 128, 256, 1, ACT_RELU, nullptr, k4, 256);
    gemm("kenc.4", k4, NI, 1, K, 256, 256, 256, 1, ACT_NONE, D, X, 512, 256);   // X[:, :256] = desc + kenc
    for (int i = 0; i < 18; ++i) {
      const std::string p = "gnn." + std::to_string(i);
      gemm(p + ".qkv", X, NI, 1, K, 256, 512, 768, 1, ACT_NONE, nullptr, qkv, 768);
      {
        // multi-head softmax attention (superglue.py:85-89) as two tcgen05 GEMMs with the heads as the H dimension:
        //   S[img][h][q][k] = Q_h K_h^T / 8   (B operand = keys of image img ^ cross, head h)
        //   P = softmax_k(S) in place, masked to the source image's keypoint count
        //   msg[img][q][64 h + d] = sum_k P[img][h][q][k] V[img ^ cross][k][64 h + d]   (B operand = V^T of that image / head)
        const int cross = i % 2, kk = K, n = NI; int* cn = counts;
        ConvGemmDesc d{};
        d.x = qkv; d.N = NI; d.H = 4; d.W = K; d.Cin = 64; d.ldx = 768; d.x_y_stride = 64; d.x_img_stride = static_cast<long long>(K) * 768;
        d.w = qkv + 256; d.Cout = K; d.R = 1; d.S = 1; d.stride = 1; d.pad = 0;
        d.b_batched = 2; d.ldw = 768; d.w_y_stride = 64; d.w_img_stride = static_cast<long long>(K) * 768; d.b_xor = cross;
        d.act = ACT_NONE; d.alpha = 0.125f; d.out = Sat; d.ldo = K; d.ldr = K;
        if (precision == 2) {   // q / k heads (64 channels each) -> [hi | lo | hi] / [hi | hi | lo] of 192, rows keep a stride of 768
          float *q3 = x3, *k3 = x3 + static_cast<size_t>(NI) * K * 768; const int rows = NI * K;
          if (rows > 65535) return MFR_ERR_ARG;
          push([=](cudaStream_t st) { return rows_split3(qkv, 64, 768, q3, 4, rows, 64, 0, nullptr, st); });
          push([=](cudaStream_t st) { return rows_split3(qkv + 256, 64, 768, k3, 4, rows, 64, 1, nullptr, st); });
          d.x = q3; d.Cin = 192; d.x_y_stride = 192;
          d.w = k3; d.w_y_stride = 192;
        }
        ConvGemmParams pq;
        int r = conv_gemm_build(d, &pq);
        if (r != MFR_OK) return r;
        push([pq](cudaStream_t st) { return conv_gemm_launch(pq, st); });
        push([=](cudaStream_t st) { return sg_softmax_rows(Sat, cn, cross, kk, n, st); });
        push([=](cudaStream_t st) { return sg_transpose_v(qkv, 768, Vt, kk, n, st); });
        ConvGemmDesc e{};
        e.x = Sat; e.N = NI; e.H = 4; e.W = K; e.Cin = K; e.ldx = K;
        e.w = Vt; e.Cout = 64; e.R = 1; e.S = 1; e.stride = 1; e.pad = 0;
        e.b_batched = 2; e.ldw = K; e.w_y_stride = 64LL * K; e.w_img_stride = 256LL * K; e.b_xor = cross;
        e.act = ACT_NONE; e.alpha = 1.f; e.out = msg; e.ldo = 256; e.ldr = 256; e.out_y_stride = 64;
        e.out_img_stride = static_cast<long long>(K) * 256;
        if (precision == 2) {   // P rows [K] -> [3K] (A pattern), V^T rows [K] -> [3K] (B pattern)
          float *p3 = x3, *v3 = x3 + static_cast<size_t>(NI) * 4 * K * 3 * K;
          const long long prow = static_cast<long long>(NI) * 4 * K, vrow = static_cast<long long>(NI) * 256;
          push([=](cudaStream_t st) { return rows_split3(Sat, kk, 0, p3, prow, 1, kk, 0, nullptr, st); });
          push([=](cudaStream_t st) { return rows_split3(Vt, kk, 0, v3, vrow, 1, kk, 1, nullptr, st); });
          e.x = p3; e.Cin = 3 * K; e.ldx = 3 * K;
          e.w = v3; e.ldw = 3 * K; e.w_y_stride = 64LL * 3 * K; e.w_img_stride = 256LL * 3 * K;
        }
        ConvGemmParams pv;
        r = conv_gemm_build(e, &pv);
        if (r != MFR_OK) return r;
        push([pv](cudaStream_t st) { return conv_gemm_launch(pv, st); });
      }
      gemm(p + ".merge", msg, NI, 1, K, 256, 256, 256, 1, ACT_NONE, nullptr, X + 256, 512);
      gemm(p + ".mlp0", X, NI, 1, K, 512, 512, 512, 1, ACT_RELU, nullptr, h1, 512);
      gemm(p + ".mlp3", h1, NI, 1, K, 512, 512, 256, 1, ACT_NONE, X, X, 512);   // x + delta (in place)
    }
    gemm("final_proj", X, NI, 1, K, 256, 512, 256, 1, ACT_NONE, nullptr, Mb, 256);
    {
      ConvGemmDesc d{};
      d.x = Mb; d.N = B; d.H = 1; d.W = K; d.Cin = 256; d.ldx = 256; d.x_img_stride = 2LL * K * 256;
      d.w = Mb + static_cast<size_t>(K) * 256; d.Cout = K; d.R = 1; d.S = 1; d.stride = 1; d.pad = 0;
      d.b_batched = 1; d.ldw = 256; d.w_img_stride = 2LL * K * 256; d.act = ACT_NONE; d.alpha = 1.f / 16.f;
      d.out = S; d.ldo = K; d.ldr = K;
      if (precision == 2) {   // matching descriptors of all images, once in the A pattern and once in the B pattern
        const long long rows = static_cast<long long>(NI) * K;
        float *ma = x3, *mb = x3 + static_cast<size_t>(rows) * 768;
        push([=](cudaStream_t st) { return rows_split3(Mb, 256, 0, ma, rows, 1, 256, 0, nullptr, st); });
        push([=](cudaStream_t st) { return rows_split3(Mb, 256, 0, mb, rows, 1, 256, 1, nullptr, st); });
        d.x = ma; d.Cin = 768; d.ldx = 768; d.x_img_stride = 2LL * K * 768;
        d.w = mb + static_cast<size_t>(K) * 768; d.ldw = 768; d.w_img_stride = 2LL * K * 768;
      }
      ConvGemmParams p;
      int r = conv_gemm_build(d, &p);
      if (r != MFR_OK) return r;
      push([p](cudaStream_t st) { return conv_gemm_launch(p, st); });
    }
    {
      const int kk = K, nb = B, iters = sk_iters; int* cn = counts; float* kp = kpts; const float th = match_thr;
      int* m0 = matches0; float* ms = mscores0; float *a0 = mk0, *a1 = mk1; int* mo = m_offsets;
      SpsgEngine* self = this;
      push([=](cudaStream_t st) { return sg_sinkhorn(S, kk, cn, self->bin_score, iters, kk, nb, u, v, st); }, 2 + 2 * sk_iters);
      push([=](cudaStream_t st) {
        return sg_extract_matches(S, kk, cn, u, v, th, kp, kk, rmax, rarg, carg, m0, ms, a0, a1, mo, nb, st);
      }, 3);
    }
    return rc;
  }

  int run(cudaStream_t st) {
    int r = prepare_weights(st);
    if (r != MFR_OK) return r;
    for (auto& s : steps) { r = s(st); if (r != MFR_OK) return r; }
    return MFR_OK;
  }
  int prepare_weights(cudaStream_t st) {
    if (precision != 2 || w3_ready) return MFR_OK;
    for (const auto& e : w3_list) {
      const size_t off = specs[idx[e.name + ".w"]].offset;
      int r = rows_split3(wblob + off, e.cin, 0, wblob3 + 3 * off, e.rows, 1, e.cin, 1, nullptr, st);
      if (r != MFR_OK) return r;
    }
    w3_ready = true;
    return MFR_OK;
  }

  // img0/img1: device [B][H][W]; copied into the pair-interleaved frame buffer
  int forward(const float* img0, const float* img1, cudaStream_t st) {
    if (!loaded) return MFR_ERR_ARG;
    const size_t hw = static_cast<size_t>(H) * W;
    MFR_CUDA_CHECK(cudaMemcpy2DAsync(img, 2 * hw * 4, img0, hw * 4, hw * 4, B, cudaMemcpyDeviceToDevice, st));
    MFR_CUDA_CHECK(cudaMemcpy2DAsync(img + hw, 2 * hw * 4, img1, hw * 4, hw * 4, B, cudaMemcpyDeviceToDevice, st));
    { int r0 = prepare_weights(st); if (r0 != MFR_OK) return r0; }   // outside any capture
    if (!use_graph) return run(st);
    if (!graph) {
      int r = run(st);
      if (r != MFR_OK) return r;
      MFR_CUDA_CHECK(cudaStreamSynchronize(st));
      cudaGraph_t g;
      MFR_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      r = run(st);
      cudaError_t e = cudaStreamEndCapture(st, &g);
      if (r != MFR_OK || e != cudaSuccess) {
        fprintf(stderr, "[mfr_b200] spsg graph capture failed (%d, %s); running eagerly\n", r, cudaGetErrorString(e));
        use_graph = false;
        cudaGetLastError();
        return run(st);
      }
      MFR_CUDA_CHECK(cudaGraphInstantiate(&graph, g, 0));
      cudaGraphDestroy(g);
      return MFR_OK;
    }
    MFR_CUDA_CHECK(cudaGraphLaunch(graph, st));
    return MFR_OK;
  }

  // SuperGlue only, on keypoints / scores / descriptors / counts uploaded by the caller (host arrays,
  // pair-interleaved image order) -- stage-boundary entry point for parity tests
  int forward_superglue(const int* h_counts, const float* h_kpts, const float* h_scores, const float* h_desc,
                        cudaStream_t st) {
    if (!loaded) return MFR_ERR_ARG;
    MFR_CUDA_CHECK(cudaMemcpyAsync(counts, h_counts, sizeof(int) * NI, cudaMemcpyHostToDevice, st));
    MFR_CUDA_CHECK(cudaMemcpyAsync(kpts, h_kpts, sizeof(float) * NI * K * 2, cudaMemcpyHostToDevice, st));
    MFR_CUDA_CHECK(cudaMemcpyAsync(kscores, h_scores, sizeof(float) * NI * K, cudaMemcpyHostToDevice, st));
    MFR_CUDA_CHECK(cudaMemcpyAsync(D, h_desc, sizeof(float) * NI * K * 256, cudaMemcpyHostToDevice, st));
    { int r0 = prepare_weights(st); if (r0 != MFR_OK) return r0; }
    for (size_t i = sg_begin; i < steps.size(); ++i) {
      int r = steps[i](st);
      if (r != MFR_OK) return r;
    }
    return MFR_OK;
  }

  ~SpsgEngine() {
    if (graph) cudaGraphExecDestroy(graph);
    for (void* p : allocs) cudaFree(p);
  }
};

int spsg_create(int B, int H, int W, int max_kp, int nms_radius, float kp_thr, int border, int sinkhorn_iters,
                float match_thr, int precision, SpsgEngine** out) {
  if (B < 1 || H < 64 || W < 64 || max_kp < 16 || max_kp > 2048 || (max_kp % 16)) return MFR_ERR_ARG;
  if (precision != 0 && precision != 2) return MFR_ERR_ARG;   // 0 = TF32 GEMM inputs, 2 = 3xTF32 split (fp32-faithful)
  auto* e = new SpsgEngine();
  e->precision = precision;
  e->B = B; e->H = H; e->W = W; e->K = max_kp; e->nms_r = nms_radius; e->kp_thr = kp_thr; e->border = border;
  e->sk_iters = sinkhorn_iters; e->match_thr = match_thr;
  int rc = e->build();
  if (rc != MFR_OK) { delete e; return rc; }
  *out = e;
  return MFR_OK;
}
void spsg_destroy(SpsgEngine* e) { delete e; }
int spsg_num_weights(const SpsgEngine* e) { return static_cast<int>(e->specs.size()); }
const char* spsg_weight_name(const SpsgEngine* e, int i) { return e->specs[i].name.c_str(); }
long long spsg_weight_numel(const SpsgEngine* e, int i) { return static_cast<long long>(e->specs[i].numel); }
int spsg_set_weight(SpsgEngine* e, int i, const float* host, long long numel, cudaStream_t st) {
  if (i < 0 || i >= static_cast<int>(e->specs.size()) || numel != static_cast<long long>(e->specs[i].numel)) return MFR_ERR_ARG;
  MFR_CUDA_CHECK(cudaMemcpyAsync(e->wblob + e->specs[i].offset, host, numel * 4, cudaMemcpyHostToDevice, st));
  MFR_CUDA_CHECK(cudaStreamSynchronize(st));
  e->w3_ready = false;
  if (e->specs[i].name == "bin_score") e->bin_score = host[0];
  if (i == static_cast<int>(e->specs.size()) - 1) e->loaded = true;
  return MFR_OK;
}
int spsg_forward(SpsgEngine* e, const float* img0, const float* img1, cudaStream_t st) { return e->forward(img0, img1, st); }
void spsg_set_graph(SpsgEngine* e, int on) { e->use_graph = on != 0; }
SpsgOutputs spsg_outputs(SpsgEngine* e) {
  SpsgOutputs o;
  o.counts = e->counts; o.keypoints = e->kpts; o.scores = e->kscores; o.descriptors = e->D; o.desc_ld = 256;
  o.matches0 = e->matches0; o.matching_scores0 = e->mscores0; o.m_offsets = e->m_offsets; o.mkpts0 = e->mk0;
  o.mkpts1 = e->mk1; o.max_keypoints = e->K;
  return o;
}
int spsg_debug_tensor(SpsgEngine* e, const char* name, float** ptr, long long* numel) {
  auto it = e->dbg.find(name);
  if (it == e->dbg.end()) return MFR_ERR_ARG;
  *ptr = it->second.first; *numel = it->second.second;
  return MFR_OK;
}
int spsg_num_launches(const SpsgEngine* e) { return e->launches; }
int spsg_forward_superglue(SpsgEngine* e, const int* counts, const float* kpts, const float* scores, const float* desc,
                           cudaStream_t st) {
  return e->forward_superglue(counts, kpts, scores, desc, st);
}

}  // namespace mfr
