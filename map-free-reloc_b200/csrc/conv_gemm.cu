// tcgen05 implicit-GEMM convolution kernel (sm_100a). See conv_gemm.cuh for the op definition.
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0      : TMA producer. Per k-block it loads a [TH x TW pixels] x [32 channels] activation
//                 box (rank-4 tensor map over NHWC; halo / padding comes back as zeros through the
//                 TMA out-of-bounds fill, stride-2 through the map's traversal stride) and a
//                 [BN filters] x [32 channels] weight box, both 128B-swizzled, into a 4-stage ring.
//   warp 1      : MMA issuer. One elected lane issues 4 x tcgen05.mma (128 x BN x 8, kind::tf32)
//                 per stage into one of two TMEM accumulators (2 x 256 columns), then
//                 tcgen05.commit frees the stage / publishes the accumulator.
//   warps 2..5  : epilogue. tcgen05.ld the accumulator (thread = output pixel, 32 channels per
//                 load), apply alpha / bias / residual / activation, write NHWC fp32.
// The epilogue of tile i overlaps the main loop of tile i+1 through the double-buffered TMEM.
#include "conv_gemm.cuh"

#include <algorithm>
#include <mutex>

namespace mfr {

namespace {

constexpr int kMaxStages = 8;
constexpr int kBlockK = 32;                     // fp32 elements per k-block = 128 B = swizzle span
constexpr int kUmmaK = 8;                       // tf32: 32 B of K per instruction
constexpr int kTileM = 128;
constexpr int kMaxBN = 256;
constexpr int kAStage = kTileM * kBlockK * 4;   // 16 KB
constexpr int kRingBytes = 192 * 1024;          // operand ring: stages = kRingBytes / (16 KB + BN * 128 B)
constexpr int kStgStride = 36;                  // floats per staged row (144 B: conflict-free 128-bit access)
constexpr int kStgBytes = 4 * 32 * kStgStride * 4;  // epilogue transpose buffers, one per epilogue warp
constexpr int kSmemBytes = kRingBytes + kStgBytes + 1024 /*align slack*/;
constexpr int kThreads = 192;
constexpr int kTmemCols = 512;

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define MFR_STAMP(i) do { if (p.dbg != nullptr && blockIdx.x == 0) p.dbg[i] = gtime(); } while (0)

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_LEAKY) return v > 0.f ? v : 0.01f * v;
  if (act == ACT_ELU1) return v > 0.f ? v + 1.f : __expf(v);
  return v;
}
template <int ACT>
__device__ __forceinline__ float act_fn(float v) {
  if constexpr (ACT == ACT_RELU) return fmaxf(v, 0.f);
  if constexpr (ACT == ACT_LEAKY) return v > 0.f ? v : 0.01f * v;
  if constexpr (ACT == ACT_ELU1) return v > 0.f ? v + 1.f : __expf(v);
  return v;
}

// Coalesced part of the epilogue for one 32-channel chunk: 8 x (LDS.128 -> alpha/bias/residual/act ->
// STG.128); lane (sub_r, sub_c) handles staged row k*4+sub_r, channels sub_c..sub_c+3.
__device__ __forceinline__ float4 bf16x4_to_float4(uint2 u) {
  float4 f;
  f.x = __uint_as_float(u.x << 16); f.y = __uint_as_float(u.x & 0xffff0000u);
  f.z = __uint_as_float(u.y << 16); f.w = __uint_as_float(u.y & 0xffff0000u);
  return f;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <int ACT, bool RES, bool OBF>
__device__ __forceinline__ void store_chunk(const float* __restrict__ stg_rd, const long long (&off_o)[8],
                                            const long long (&off_r)[8], uint32_t okmask, float alpha,
                                            float4 b4, const float* __restrict__ res_n,
                                            float* __restrict__ out_n) {
  float4 vv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) vv[k] = *reinterpret_cast<const float4*>(stg_rd + k * (4 * kStgStride));
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float4 v = vv[k];
    if (okmask & (1u << k)) {
      v.x = fmaf(v.x, alpha, b4.x); v.y = fmaf(v.y, alpha, b4.y);
      v.z = fmaf(v.z, alpha, b4.z); v.w = fmaf(v.w, alpha, b4.w);
      if constexpr (RES) {
        float4 r4;
        if constexpr (OBF)   // res_n / out_n point at bf16 storage: element offsets count 2-byte units
          r4 = bf16x4_to_float4(__ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(res_n) + off_r[k])));
        else
          r4 = __ldg(reinterpret_cast<const float4*>(res_n + off_r[k]));
        v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
      }
      v.x = act_fn<ACT>(v.x); v.y = act_fn<ACT>(v.y); v.z = act_fn<ACT>(v.z); v.w = act_fn<ACT>(v.w);
      if constexpr (OBF)
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(out_n) + off_o[k]) =
            make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
      else
        *reinterpret_cast<float4*>(out_n + off_o[k]) = v;
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_tf32_kernel(const __grid_constant__ ConvGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bars[2 * kMaxStages + 4];
  __shared__ uint32_t s_tmem_base;
  const int kStages = p.num_stages;
  const int kBStage = p.b_bytes;                 // BN rows x 128 B
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + kStages * kAStage;
  const uint32_t smem_stg = smem_base + kRingBytes;
  const uint32_t bar_base = smem_u32(s_bars);
  // barriers: full[kMaxStages], empty[kMaxStages], tmem_full[2], tmem_empty[2]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + 2 + a); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) MFR_STAMP(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&s_tmem_base), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  if (threadIdx.x == 0) MFR_STAMP(1);

  int w_lim = p.Wo, tiles_x = p.tiles_x;
  if (p.dyn_w != nullptr) {
    w_lim = min(p.Wo, *p.dyn_w);
    tiles_x = (w_lim + p.TW - 1) / p.TW;
  }
  const int m_tiles = p.N * p.tiles_y * tiles_x;
  const int total_tiles = m_tiles * p.n_tiles;
  const int num_kb = p.R * p.S * p.kb_per_tap;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile % m_tiles, nt = tile / m_tiles;
        const int tx = mt % tiles_x;
        const int ty = (mt / tiles_x) % p.tiles_y;
        const int img = mt / (tiles_x * p.tiles_y);
        const int x_in0 = tx * p.TW * p.stride - p.pad;
        const int y_in0 = ty * p.TH * p.stride - p.pad;
        const int n0 = nt * p.BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / p.kb_per_tap;
          const int cb = kb - tap * p.kb_per_tap;
          const int ky = tap / p.S, kx = tap - ky * p.S;
          mbar_wait(empty_bar(stage), phase ^ 1u, 1);
          mbar_arrive_expect_tx(full_bar(stage), static_cast<uint32_t>(p.a_bytes + p.b_bytes));
          tma_load_4d(smem_a + stage * kAStage, &p.tmA, full_bar(stage), cb * p.kblk, x_in0 + kx,
                      y_in0 + ky, img);
          tma_load_3d(smem_b + stage * kBStage, &p.tmB, full_bar(stage), cb * p.kblk,
                      p.b_batched ? img : tap, n0);
          if (kb == 0 && tile == static_cast<int>(blockIdx.x)) MFR_STAMP(2);
          if (kb == num_kb - 1 && tile == static_cast<int>(blockIdx.x)) MFR_STAMP(3);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(p.in_bf16 ? 1u : 2u, kTileM, static_cast<uint32_t>(p.BN));
      const bool bf16_in = p.in_bf16 != 0;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 2);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(acc * kMaxBN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, 3);
          tc_fence_after();
          if (kb == 0 && it == 0) MFR_STAMP(4);
          if (kb == num_kb - 1 && it == 0) MFR_STAMP(5);
          const uint32_t a_addr = smem_a + stage * kAStage;
          const uint32_t b_addr = smem_b + stage * kBStage;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t adesc = umma_desc_sw128_kmajor(a_addr + k * kUmmaK * 4);
            const uint64_t bdesc = umma_desc_sw128_kmajor(b_addr + k * kUmmaK * 4);
            if (bf16_in) tc_mma_bf16(tmem_acc, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            else tc_mma_tf32(tmem_acc, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(empty_bar(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(acc));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps)
    // TMEM -> registers (thread = output pixel) -> per-warp shared-memory transpose -> coalesced
    // global stores (8 lanes cover 128 contiguous bytes of one pixel, 4 pixels per instruction).
    const int q = warp & 3;  // TMEM lane quadrant this warp is allowed to read
    const int row = q * 32 + lane;
    const int ly = row / p.TW, lx = row - ly * p.TW;
    const bool vec_ok = ((p.Cout & 3) == 0) && ((p.ldo & 3) == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.out) & (p.out_bf16 ? 7 : 15)) == 0) &&
                        (p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) &&
                        (p.residual == nullptr ||
                         (((p.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & (p.out_bf16 ? 7 : 15)) == 0)));
    float* stg = reinterpret_cast<float*>(smem_raw + (smem_stg - smem_u32(smem_raw))) + q * (32 * kStgStride);
    const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int mt = tile % m_tiles, nt = tile / m_tiles;
      const int tx = mt % tiles_x;
      const int ty = (mt / tiles_x) % p.tiles_y;
      const int img = mt / (tiles_x * p.tiles_y);
      const int y = ty * p.TH + ly, x = tx * p.TW + lx;
      const bool row_ok = (row < p.TW * p.TH) && (y < p.Ho) && (x < w_lim);
      const long long pix = static_cast<long long>(y) * p.Wo + x;   // pixel inside the image
      const int n0 = nt * p.BN;
      // rows this lane stores in the coalesced phase: staged row k*4+sub_r of this warp's 32 pixels
      long long off_o[8], off_r[8];
      uint32_t okmask = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int rr = k * 4 + sub_r;
        const long long pr = __shfl_sync(0xffffffffu, pix, rr);
        okmask |= (__shfl_sync(0xffffffffu, row_ok ? 1 : 0, rr) != 0 ? 1u : 0u) << k;
        off_o[k] = img * p.out_img_stride + pr * p.ldo;
        off_r[k] = img * p.res_img_stride + pr * p.ldr;
      }

      mbar_wait(tfull_bar(acc), acc_phase, 4);
      tc_fence_after();
      if (it == 0 && threadIdx.x == 64) MFR_STAMP(6);
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * kMaxBN) +
                             (static_cast<uint32_t>(q * 32) << 16);
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t r[32];
        if (it == 0 && c0 == 0 && threadIdx.x == 64) MFR_STAMP(9);
        tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c0), r);
        tmem_ld_wait();
        if (it == 0 && c0 == 0 && threadIdx.x == 64) MFR_STAMP(10);
        if (vec_ok) {
          // stage: this thread's pixel, 32 consecutive channels
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stg + lane * kStgStride + j * 4) =
                make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          __syncwarp();
          if (it == 0 && c0 == 0 && threadIdx.x == 64) MFR_STAMP(11);
          const int n = n0 + c0 + sub_c;
          const uint32_t okm = n < p.Cout ? okmask : 0u;
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias != nullptr && okm) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
          const int act = n < p.act_cols ? p.act : ACT_NONE;   // n..n+3 never straddle act_cols (multiple of 4)
          const float* stg_rd = stg + sub_r * kStgStride + sub_c;
          // base pointers advanced by n elements of the output type
          float* out_n = p.out_bf16 ? reinterpret_cast<float*>(reinterpret_cast<uint16_t*>(p.out) + n) : p.out + n;
          const float* res_n = p.out_bf16 ? reinterpret_cast<const float*>(reinterpret_cast<const uint16_t*>(p.residual) + n)
                                          : p.residual + n;
#define MFR_STORE(ACTV, RESV)                                                                                   \
  do {                                                                                                          \
    if (p.out_bf16) store_chunk<ACTV, RESV, true>(stg_rd, off_o, off_r, okm, p.alpha, b4, res_n, out_n);        \
    else store_chunk<ACTV, RESV, false>(stg_rd, off_o, off_r, okm, p.alpha, b4, res_n, out_n);                  \
  } while (0)
          if (p.residual != nullptr) {
            switch (act) {
              case ACT_RELU: MFR_STORE(ACT_RELU, true); break;
              case ACT_LEAKY: MFR_STORE(ACT_LEAKY, true); break;
              case ACT_ELU1: MFR_STORE(ACT_ELU1, true); break;
              default: MFR_STORE(ACT_NONE, true); break;
            }
          } else {
            switch (act) {
              case ACT_RELU: MFR_STORE(ACT_RELU, false); break;
              case ACT_LEAKY: MFR_STORE(ACT_LEAKY, false); break;
              case ACT_ELU1: MFR_STORE(ACT_ELU1, false); break;
              default: MFR_STORE(ACT_NONE, false); break;
            }
          }
#undef MFR_STORE
          __syncwarp();
          if (it == 0 && c0 == 0 && threadIdx.x == 64) MFR_STAMP(12);
        } else if (row_ok) {
          float* orow = p.out + img * p.out_img_stride + pix * p.ldo;
          const float* rrow = p.residual ? p.residual + img * p.res_img_stride + pix * p.ldr : nullptr;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + c0 + j;
            if (n < p.Cout) {
              float v = __uint_as_float(r[j]) * p.alpha;
              if (p.bias) v += __ldg(p.bias + n);
              if (rrow) v += __ldg(rrow + n);
              orow[n] = n < p.act_cols ? apply_act(v, p.act) : v;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (it == 0 && threadIdx.x == 64) MFR_STAMP(7);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
  if (threadIdx.x == 0) MFR_STAMP(8);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(f);
    }
  });
  return fn;
}

// Pick the output tile (TW x TH <= 128 pixels) that wastes the fewest MMA rows.
void choose_tile(int Wo, int Ho, int stride, int* TW, int* TH) {
  double best = -1.0;
  int bw = 1, bh = 1;
  for (int tw = 1; tw <= 128; ++tw) {
    if (tw * stride > 256) break;
    int th = 128 / tw;
    if (th > Ho) th = Ho;
    if (th * stride > 256) th = 256 / stride;
    if (th < 1) continue;
    const long long tiles = static_cast<long long>((Wo + tw - 1) / tw) * ((Ho + th - 1) / th);
    const double eff = static_cast<double>(Wo) * Ho / (static_cast<double>(tiles) * 128.0);
    if (eff > best + 1e-9 || (eff > best - 1e-9 && tw > bw)) {
      best = std::max(best, eff);
      bw = tw;
      bh = th;
    }
  }
  *TW = bw;
  *TH = bh;
}

}  // namespace

int conv_gemm_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  }
  return sms;
}

int conv_gemm_build(const ConvGemmDesc& d, ConvGemmParams* p) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    fprintf(stderr, "[mfr_b200] cuTensorMapEncodeTiled unavailable (no CUDA driver?)\n");
    return MFR_ERR_CUDA;
  }
  const int es = d.in_bf16 ? 2 : 4;           // operand element size
  const int kblk = 128 / es;                  // K elements per 128-byte k-block
  const long long ldw_e = d.ldw > 0 ? d.ldw : d.Cin;
  if ((d.ldx * es) % 16 != 0 || (ldw_e * es) % 16 != 0 || (reinterpret_cast<uintptr_t>(d.x) & 15) ||
      (reinterpret_cast<uintptr_t>(d.w) & 15)) {
    fprintf(stderr, "[mfr_b200] conv_gemm: operand row strides must be multiples of 16 bytes, pointers 16B aligned\n");
    return MFR_ERR_ARG;
  }
  if (d.out_bf16 && ((d.Cout & 3) || (d.ldo & 3))) return MFR_ERR_ARG;   // bf16 output: vector path only
  if (d.stride != 1 && d.stride != 2) return MFR_ERR_ARG;
  memset(p, 0, sizeof(*p));
  p->N = d.N;
  p->Ho = (d.H + 2 * d.pad - d.R) / d.stride + 1;
  p->Wo = (d.W + 2 * d.pad - d.S) / d.stride + 1;
  p->Cout = d.Cout;
  p->Cin = d.Cin;
  p->R = d.R; p->S = d.S; p->stride = d.stride; p->pad = d.pad;
  choose_tile(p->Wo, p->Ho, d.stride, &p->TW, &p->TH);
  p->tiles_x = (p->Wo + p->TW - 1) / p->TW;
  p->tiles_y = (p->Ho + p->TH - 1) / p->TH;
  // N tile: whole Cout when it fits one UMMA (<= 256), else equal 16-aligned slices; problems with
  // fewer tiles than SMs are cut into narrower N tiles (down to 64) so that more SMs take part.
  const int c16 = (d.Cout + 15) / 16 * 16;
  const int m_tiles = p->N * p->tiles_x * p->tiles_y;
  int max_bn = kMaxBN;
  while (max_bn > 64 && static_cast<long long>(m_tiles) * ((c16 + max_bn - 1) / max_bn) < conv_gemm_num_sms() &&
         c16 > max_bn / 2)
    max_bn /= 2;
  p->n_tiles = (c16 + max_bn - 1) / max_bn;
  p->BN = ((c16 / 16 + p->n_tiles - 1) / p->n_tiles) * 16;
  p->kb_per_tap = (d.Cin + kblk - 1) / kblk;
  p->kblk = kblk;
  p->in_bf16 = d.in_bf16;
  p->out_bf16 = d.out_bf16;
  p->b_batched = d.b_batched;
  p->a_bytes = p->TW * p->TH * 128;
  p->b_bytes = p->BN * 128;
  p->num_stages = std::min(kMaxStages, kRingBytes / (kAStage + p->b_bytes));  // BN*128 B is 2 KB aligned
  p->act = d.act;
  p->act_cols = d.act_cols > 0 ? d.act_cols : d.Cout;
  p->dyn_w = d.dyn_w;
  if (d.dyn_w != nullptr && (d.H != 1 || d.R != 1 || d.S != 1 || d.stride != 1)) return MFR_ERR_ARG;
  p->alpha = d.alpha;
  p->bias = d.bias;
  p->residual = d.residual;
  p->ldr = d.ldr;
  p->out = d.out;
  p->ldo = d.ldo;
  p->out_img_stride = d.out_img_stride > 0 ? d.out_img_stride : static_cast<long long>(p->Ho) * p->Wo * d.ldo;
  p->res_img_stride = static_cast<long long>(p->Ho) * p->Wo * d.ldr;
  p->dbg = nullptr;

  {  // activations: {C, W, H, N}
    cuuint64_t dims[4] = {(cuuint64_t)d.Cin, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.N};
    cuuint64_t strides[3] = {(cuuint64_t)d.ldx * es, (cuuint64_t)d.ldx * es * d.W,
                             d.x_img_stride > 0 ? (cuuint64_t)d.x_img_stride * es : (cuuint64_t)d.ldx * es * d.W * d.H};
    cuuint32_t box[4] = {(cuuint32_t)kblk, (cuuint32_t)(p->TW * d.stride),
                         (cuuint32_t)(p->TH * d.stride), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)d.stride, (cuuint32_t)d.stride, 1};
    CUresult r = enc(&p->tmA, d.in_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                     const_cast<float*>(d.x), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      fprintf(stderr, "[mfr_b200] cuTensorMapEncodeTiled(A) failed: %d\n", (int)r);
      return MFR_ERR_CUDA;
    }
  }
  {  // weights: {Cin, Z, Cout}; Z = tap (shared weights) or image (batched B operand)
    const int taps = d.R * d.S;
    cuuint64_t dims[3];
    cuuint64_t strides[2];
    if (d.b_batched) {
      const long long ldw = ldw_e;
      dims[0] = d.Cin; dims[1] = d.N; dims[2] = d.Cout;
      strides[0] = d.w_img_stride > 0 ? (cuuint64_t)d.w_img_stride * es : (cuuint64_t)d.Cout * ldw * es;  // next image
      strides[1] = (cuuint64_t)ldw * es;           // next output row
    } else {
      dims[0] = d.Cin; dims[1] = taps; dims[2] = d.Cout;
      strides[0] = (cuuint64_t)ldw_e * es;
      strides[1] = (cuuint64_t)ldw_e * es * taps;
    }
    cuuint32_t box[3] = {(cuuint32_t)kblk, 1, (cuuint32_t)p->BN};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&p->tmB, d.in_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                     const_cast<float*>(d.w), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      fprintf(stderr, "[mfr_b200] cuTensorMapEncodeTiled(B) failed: %d\n", (int)r);
      return MFR_ERR_CUDA;
    }
  }
  return MFR_OK;
}

int conv_gemm_launch(const ConvGemmParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    MFR_CUDA_CHECK(cudaFuncSetAttribute(conv_gemm_tf32_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  const int total = p.N * p.tiles_y * p.tiles_x * p.n_tiles;
  if (total <= 0) return MFR_OK;
  const int grid = std::min(total, conv_gemm_num_sms());
  conv_gemm_tf32_kernel<<<grid, kThreads, kSmemBytes, stream>>>(p);
  MFR_CUDA_CHECK(cudaGetLastError());
  return MFR_OK;
}

// ------------------------------------------------------------------------------------------------
// launch-overhead probes (debug): same launch shape as the GEMM kernel, progressively more set-up
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(kThreads, 1) overhead_probe_kernel(int mode, int* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bars[2 * kMaxStages + 4];
  __shared__ uint32_t s_tmem_base;
  if (mode == 0) return;
  const int warp = threadIdx.x >> 5;
  if (mode >= 2 && threadIdx.x == 0) {
    for (int s = 0; s < 2 * kMaxStages + 4; ++s) mbar_init(smem_u32(s_bars) + 8u * s, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&s_tmem_base), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (mode >= 3 && threadIdx.x == 0 && sink != nullptr) smem_raw[0] = static_cast<uint8_t>(*sink);
  __syncthreads();
  if (warp == 1) tmem_dealloc(s_tmem_base, kTmemCols);
}
}  // namespace

float conv_gemm_probe_overhead(int mode, int smem_bytes, int grid, int iters, cudaStream_t st) {
  cudaFuncSetAttribute(overhead_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 5; ++i) overhead_probe_kernel<<<grid, kThreads, smem_bytes, st>>>(mode, nullptr);
  cudaEventRecord(e0, st);
  for (int i = 0; i < iters; ++i) overhead_probe_kernel<<<grid, kThreads, smem_bytes, st>>>(mode, nullptr);
  cudaEventRecord(e1, st);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return ms * 1000.f / iters;
}

float conv_gemm_time_prebuilt(const ConvGemmParams& p, int iters, cudaStream_t st) {
  {  // one instrumented launch: milestone stamps of block 0
    long long* d = nullptr;
    cudaMalloc(&d, 16 * sizeof(long long));
    cudaMemset(d, 0, 16 * sizeof(long long));
    ConvGemmParams q = p;
    q.dbg = d;
    conv_gemm_launch(q, st);
    conv_gemm_launch(q, st);
    cudaStreamSynchronize(st);
    long long h[16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    const char* names[13] = {"start", "setup done", "first TMA issued", "last TMA issued", "first stage landed",
                             "last stage landed", "accumulator ready", "epilogue done", "exit", "c0 begin",
                             "c0 tmem loaded", "c0 staged", "c0 stored"};
    fprintf(stderr, "[conv_gemm stamps, block 0, ns since start]");
    for (int i = 0; i < 13; ++i) fprintf(stderr, " %s=%lld", names[i], h[i] ? h[i] - h[0] : -1);
    fprintf(stderr, "\n");
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 5; ++i) conv_gemm_launch(p, st);
  cudaEventRecord(e0, st);
  for (int i = 0; i < iters; ++i) conv_gemm_launch(p, st);
  cudaEventRecord(e1, st);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return ms * 1000.f / iters;
}

}  // namespace mfr
