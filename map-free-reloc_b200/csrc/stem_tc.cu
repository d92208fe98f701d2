// The 7x7 stride-2 stem (1 -> 128 channels, folded BN, ReLU; resnet_fpn.py:52-62,102) on tcgen05, bf16 output.
//
// Same arithmetic as stem_mma_kernel (loftr_ops.cu): image pixels and weights are split into bf16 hi + lo parts and
// out = hi*hi + lo*hi + hi*lo accumulated in fp32 (error ~2^-16 relative), K = 8 ky x 8 kx per term (row ky = 7 and tap
// kx = 7 are zero). Here the three terms are three 64-wide k-blocks of ONE GEMM per 128-pixel tile:
//   A [128 px][192] = [hi | lo | hi] patches, built in shared memory by 128 builder threads (thread = output pixel) from a
//                     21 x 37 fp32 image patch, K-major rows of 128 bytes per k-block, 128B swizzle;
//   B [128 ch][192] = [hi | hi | lo] weights, built once per CTA;
//   D [128 px][128 ch] fp32 in tensor memory, 12 tcgen05.mma (M = N = 128, K = 16) per tile, double-buffered.
// Persistent CTAs, 13 warps: 0-7 builders (the next tile's patch is in flight as cp.async while this one is converted), 8 MMA
// issuer, 9-12 epilogue (bias + ReLU -> bf16 -> swizzled staging tile -> TMA store of 2 rows x 16 px x 64 channels). The mma.sync version needed 3 x 16 warp-level MMAs per 16 pixels and was
// bound by them (0.59 ms per batch); this one is bound by the 0.8 GB it writes.
#include "loftr_ops.cuh"

#include <algorithm>
#include <cstdio>
#include <mutex>

namespace mfr {

namespace {

constexpr int kTW = 16, kTH = 8;               // output pixels of a tile
constexpr int kPR = 2 * kTH + 5, kPC = 40;     // patch rows, padded patch row length (37 used)
constexpr int kKb = 16384;                     // one k-block of an operand: [128 rows][64 bf16], 128B-swizzled
constexpr int kBuilders = 256;                 // two threads per output pixel, four ky rows each
constexpr int kThreadsStem = kBuilders + 5 * 32;
constexpr int kSmemStem = 3 * kKb /*B*/ + 2 * 3 * kKb /*A*/ + 2 * kPR * kPC * 4 /*patches*/ + 4 * 4096 /*staging*/ + 1024;
constexpr uint64_t kDescHiStem = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
__device__ __forceinline__ uint64_t sdesc(uint32_t addr) { return kDescHiStem | ((addr >> 4) & 0x3fffu); }

__device__ __forceinline__ unsigned short f2bf_rn(float f) {
  unsigned u = __float_as_uint(f);
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<unsigned short>(u >> 16);
}
__device__ __forceinline__ float bf2f_(unsigned short b) { return __uint_as_float(static_cast<unsigned>(b) << 16); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

struct StemParams {
  const float* img;
  const float* w;      // [128][49]
  const float* bias;   // [128]
  int NI, H, W, Ho, Wo, tiles_x, tiles_y;
  CUtensorMap tmOut;   // bf16 {128 ch, Wo, Ho, NI}, box {64, 16, 2, 1}, 128B swizzle
};

__global__ void __launch_bounds__(kThreadsStem, 1) stem_tc_kernel(const __grid_constant__ StemParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bars[8];
  __shared__ uint32_t s_tmem_base;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));          // generic pointer to the aligned base
  const uint32_t sB = base, sA = base + 3 * kKb, sStg = base + 9 * kKb;      // operand and staging tiles are 1 KB aligned (swizzle)
  float* patch = reinterpret_cast<float*>(gen + 9 * kKb + 4 * 4096);
  const uint32_t bb = smem_u32(s_bars);
  auto a_full = [&](int b) { return bb + 8u * b; };
  auto a_empty = [&](int b) { return bb + 16u + 8u * b; };
  auto d_full = [&](int a) { return bb + 32u + 8u * a; };
  auto d_empty = [&](int a) { return bb + 48u + 8u * a; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
  if (t == 0) {
    tma_prefetch_desc(&p.tmOut);
    for (int b = 0; b < 2; ++b) {
      mbar_init(a_full(b), kBuilders / 32); mbar_init(a_empty(b), 1); mbar_init(d_full(b), 1); mbar_init(d_empty(b), 4);
    }
    fence_barrier_init();
  }
  if (warp == kBuilders / 32) {
    tmem_alloc(smem_u32(&s_tmem_base), 256);
    tmem_relinquish();
  }
  // B operand: [hi | hi | lo] of the weights; row n = output channel, 16-byte chunk ky = taps (ky, kx = 0..7), zero at ky = 7 / kx = 7
  for (int idx = t; idx < 128 * 8; idx += kThreadsStem) {
    const int n = idx >> 3, ky = idx & 7;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int kx = 2 * k2 + e;
        v[e] = (ky < 7 && kx < 7) ? p.w[n * 49 + ky * 7 + kx] : 0.f;
      }
      const unsigned short h0 = f2bf_rn(v[0]), h1 = f2bf_rn(v[1]);
      hi[k2] = static_cast<uint32_t>(h0) | (static_cast<uint32_t>(h1) << 16);
      lo[k2] = pack2(v[0] - bf2f_(h0), v[1] - bf2f_(h1));
    }
    const uint32_t off = static_cast<uint32_t>(n) * 128u + ((static_cast<uint32_t>(ky) ^ (n & 7)) << 4);
    *reinterpret_cast<uint4*>(gen + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(gen + kKb + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(gen + 2 * kKb + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  const int tiles_img = p.tiles_x * p.tiles_y, total = p.NI * tiles_img;

  if (warp < kBuilders / 32) {
    // -------------------------------------------------------------------------------------------- builders (two threads per pixel)
    const int pix = t & 127, half = t >> 7, py = pix >> 4, px = pix & 15;
    // image patch of a tile -> shared memory as 4-byte cp.async (zero fill outside the image)
    auto load_patch = [&](int tile, int buf) {
      const int img = tile / tiles_img, r_ = tile - img * tiles_img, ty = r_ / p.tiles_x, tx = r_ - ty * p.tiles_x;
      const int y_in0 = 2 * ty * kTH - 3, x_in0 = 2 * tx * kTW - 3;
      const float* im = p.img + static_cast<size_t>(img) * p.H * p.W;
      const uint32_t pb = smem_u32(patch + buf * (kPR * kPC));
      for (int idx = t; idx < kPR * kPC; idx += kBuilders) {
        const int r = idx / kPC, c = idx - r * kPC;
        const int iy = y_in0 + r, ix = x_in0 + c;
        const bool ok = c < 2 * kTW + 5 && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
        const float* src = ok ? im + static_cast<size_t>(iy) * p.W + ix : im;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(pb + 4u * idx), "l"(src), "r"(ok ? 4 : 0) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int it = 0;
    if (static_cast<int>(blockIdx.x) < total) load_patch(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");         // this tile's patch is complete; everyone is done reading the other one
      if (tile + static_cast<int>(gridDim.x) < total) load_patch(tile + gridDim.x, buf ^ 1);
      if (it >= 2) mbar_wait(a_empty(buf), ((it >> 1) - 1) & 1, 41);          // the MMAs of tile it - 2 have read this A buffer
      const float* pb = patch + buf * (kPR * kPC);
      uint8_t* ab = gen + 3 * kKb + buf * 3 * kKb + pix * 128;
      const uint32_t sw = static_cast<uint32_t>(pix & 7);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const int ky = 4 * half + k4;
        uint32_t hi[4] = {0u, 0u, 0u, 0u}, lo[4] = {0u, 0u, 0u, 0u};
        if (ky < 7) {
          const float* pr = pb + (2 * py + ky) * kPC + 2 * px;
          float v[8];
#pragma unroll
          for (int kx = 0; kx < 7; ++kx) v[kx] = pr[kx];
          v[7] = 0.f;
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) {
            const unsigned short h0 = f2bf_rn(v[2 * k2]), h1 = f2bf_rn(v[2 * k2 + 1]);
            hi[k2] = static_cast<uint32_t>(h0) | (static_cast<uint32_t>(h1) << 16);
            lo[k2] = pack2(v[2 * k2] - bf2f_(h0), v[2 * k2 + 1] - bf2f_(h1));
          }
        }
        const uint32_t co = (static_cast<uint32_t>(ky) ^ sw) << 4;
        *reinterpret_cast<uint4*>(ab + co) = make_uint4(hi[0], hi[1], hi[2], hi[3]);               // [hi | lo | hi]
        *reinterpret_cast<uint4*>(ab + kKb + co) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(ab + 2 * kKb + co) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(buf));
    }
  } else if (warp == kBuilders / 32) {
    // -------------------------------------------------------------------------------------------- MMA issuer
    const uint32_t idesc = umma_idesc(1u, 128u, 128u);
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(a_full(buf), (it >> 1) & 1, 42);
      if (it >= 2) mbar_wait(d_empty(buf), ((it >> 1) - 1) & 1, 43);          // the epilogue has drained this accumulator
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t d = tmem_base + static_cast<uint32_t>(buf * 128);
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {
          const uint64_t ad = sdesc(sA + (buf * 3 + kb) * kKb), bd = sdesc(sB + kb * kKb);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, (kb == 0 && k == 0) ? 0u : 1u);
        }
        tc_commit(a_empty(buf));
        tc_commit(d_full(buf));
      }
      __syncwarp();
    }
  } else {
    // -------------------------------------------------------------------------------------------- epilogue
    const int q = warp & 3;                                  // TMEM lane quadrant = pixels 32 q .. 32 q + 31 = tile rows 2 q, 2 q + 1
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t stg = sStg + static_cast<uint32_t>(q) * 4096u, row_addr = stg + static_cast<uint32_t>(lane) * 128u;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int img = tile / tiles_img, r_ = tile - img * tiles_img, ty = r_ / p.tiles_x, tx = r_ - ty * p.tiles_x;
      mbar_wait(d_full(buf), (it >> 1) & 1, 44);
      tc_fence_after();
      const uint32_t d = tmem_base + static_cast<uint32_t>(buf * 128) + lane_off;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t ra[2][32];
        tmem_ld_32x32b_x32(d + static_cast<uint32_t>(64 * g), ra[0]);
        tmem_ld_32x32b_x32(d + static_cast<uint32_t>(64 * g + 32), ra[1]);
        tmem_ld_wait();
        if (g == 1) {                                        // last read of this accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d_empty(buf));
        }
        uint32_t o[32];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + 64 * g + 32 * h) + j4);
            o[16 * h + 2 * j4] = pack2(fmaxf(__uint_as_float(ra[h][4 * j4]) + b4.x, 0.f), fmaxf(__uint_as_float(ra[h][4 * j4 + 1]) + b4.y, 0.f));
            o[16 * h + 2 * j4 + 1] = pack2(fmaxf(__uint_as_float(ra[h][4 * j4 + 2]) + b4.z, 0.f), fmaxf(__uint_as_float(ra[h][4 * j4 + 3]) + b4.w, 0.f));
          }
        if (elect_one_sync()) tma_store_wait_read<0>();      // the previous store has read the staging tile
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + ((static_cast<uint32_t>(j) ^ sw) << 4)), "r"(o[4 * j]),
                       "r"(o[4 * j + 1]), "r"(o[4 * j + 2]), "r"(o[4 * j + 3])
                       : "memory");
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one_sync()) {
          tma_store_4d(&p.tmOut, stg, 64 * g, tx * kTW, ty * kTH + 2 * q, img);
          tma_store_commit();
        }
      }
    }
    tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kBuilders / 32) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn stem_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

}  // namespace

int stem_conv7x7_tc(const float* img, const float* w, const float* bias, void* out_bf16, int NI, int H, int W, cudaStream_t st, int grid_limit) {
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  EncodeTiledFn enc = stem_encode_fn();
  if (!enc) return MFR_ERR_CUDA;
  if ((reinterpret_cast<uintptr_t>(out_bf16) & 15) || (reinterpret_cast<uintptr_t>(bias) & 15)) return MFR_ERR_ARG;
  StemParams p{};
  p.img = img; p.w = w; p.bias = bias; p.NI = NI; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo;
  p.tiles_x = (Wo + kTW - 1) / kTW; p.tiles_y = (Ho + kTH - 1) / kTH;
  cuuint64_t dims[4] = {128, static_cast<cuuint64_t>(Wo), static_cast<cuuint64_t>(Ho), static_cast<cuuint64_t>(NI)};
  cuuint64_t strides[3] = {256, static_cast<cuuint64_t>(Wo) * 256, static_cast<cuuint64_t>(Ho) * Wo * 256};
  cuuint32_t box[4] = {64, kTW, 2, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&p.tmOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out_bf16, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[mfr_b200] stem: cuTensorMapEncodeTiled failed: %d\n", static_cast<int>(r));
    return MFR_ERR_CUDA;
  }
  static bool attr_seen[64] = {};
  if (first_call_on_device(attr_seen))
    MFR_CUDA_CHECK(cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemStem));
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int total = NI * p.tiles_x * p.tiles_y;
  if (total <= 0) return MFR_OK;
  if (grid_limit > 0) sms = std::min(sms, grid_limit);
  stem_tc_kernel<<<std::min(total, sms), kThreadsStem, kSmemStem, st>>>(p);
  MFR_CUDA_CHECK(cudaGetLastError());
  return MFR_OK;
}

}  // namespace mfr
