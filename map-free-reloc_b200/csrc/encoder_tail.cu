// Fused tail of a LoFTREncoderLayer on the 5th-gen tensor cores (sm_100a, bf16 engine mode).
//
// Reference (LoFTR/src/loftr/loftr_module/transformer.py:52-58), per token row:
//     m  = LayerNorm1(message @ Wm^T)                       merge + norm1
//     h  = relu([x | m] @ W0^T)                             mlp.0 on the concatenation (2C -> 2C)
//     x += LayerNorm2(h @ W2^T)                             mlp.2 (2C -> C) + norm2 + residual
// Unfused this is 3 GEMM launches + 2 LayerNorm launches whose intermediates (merge output, norm1 output, the 2C-wide
// hidden tensor, the mlp.2 output) each make a round trip through HBM: ~7 KB of the ~14 KB a token row moves per layer.
// Rows are independent, so one persistent CTA per SM takes 128-row tiles through the whole chain with every
// intermediate kept on chip:
//
//   TMEM   D3  [128 x C]  fp32   accumulator of GEMM3 (mlp.2), columns [0, C)
//          scratch, 256 columns  D1 = accumulator of GEMM1 (merge); once LayerNorm1 has consumed it, the two D2 buffers
//                                [128 x 128] = double-buffered accumulator of GEMM2 (mlp.0), one 128-column chunk q of the hidden layer
//   SMEM   L   [128 x C]  bf16   LayerNorm1 output = k-blocks C/64 .. 2C/64-1 of GEMM2's A operand (K-major, 128B swizzle)
//          M   64 KB             two [128 x 128] bf16 buffers H1[q & 1]: relu(D2_q), the A operand of GEMM3
//          ring of 16 KB slots   every streamed tile [128 rows x 64 bf16]: message / x k-blocks of this row tile and 128
//                                filter rows of a weight k-block, by TMA from L2 (6-8 slots in flight: the per-tile traffic
//                                of ~1.1 MB needs depth, not width, to hide the L2 latency)
//
//   warp 0     TMA producer: the ring items in exactly the order the MMA warp consumes them
//   warp 1     MMA issuer  : GEMM1; then GEMM2 chunk q+1 is issued BEFORE GEMM3 chunk q, so the tensor pipe runs the
//                            next chunk while the epilogue warps turn chunk q into the bf16 A operand of GEMM3
//   warps 2-5  front epilogue, thread = row: LayerNorm1 (TMEM -> registers -> swizzled smem), relu + bf16 (TMEM -> smem)
//   warps 6-9  back epilogue, thread = row: LayerNorm2 + residual of the PREVIOUS tile (TMEM -> staging -> coalesced fp32
//              store of x and its bf16 shadow; the residual rows are prefetched a chunk ahead) while the front group and
//              the tensor pipe already work on the next tile
//
// The hidden layer is produced and consumed in 128-column chunks because GEMM3 accumulates over exactly those
// columns: D3 += relu(D2_q) @ W2[:, q]^T can start as soon as chunk q exists, and 2C-wide rows never exist anywhere.
#include "encoder_tail.cuh"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "conv_gemm.cuh"

namespace mfr {

namespace {

constexpr int kEpiWarps = 8;          // front group (warps 2-5) and back group (warps 6-9), one warp per TMEM lane quadrant each
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kRelBatch = 3;        // ring slots handed back per tcgen05.commit batch
constexpr int kKbBytes = 16384;     // one k-block of an A operand: [128 rows x 64 bf16], 128B-swizzled
constexpr int kTmemCols = 512;
constexpr uint64_t kDescHi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
__device__ __forceinline__ uint64_t desc_at(uint32_t addr) { return kDescHi | ((addr >> 4) & 0x3fffu); }

template <int C>
struct Cfg {
  static constexpr int KB = C / 64;            // k-blocks of a C-wide operand
  static constexpr int NQ = 2 * C / 128;       // 128-column chunks of the hidden layer
  static constexpr int L_BYTES = KB * kKbBytes;
  static constexpr int M_BYTES = 4 * kKbBytes; // message tile (KB k-blocks) / two H1 buffers of two k-blocks
  static constexpr int NH = C / 128;           // 128-column halves of a C-wide output (every MMA is N = 128)
  static constexpr int NS = (C == 256) ? 5 : 7;  // ring slots of 16 KB
  static constexpr int STG_BYTES = 4 * 4096;     // staging tiles of the LayerNorm2 warps (32 rows x 32 fp32 each)
  static constexpr int SMEM = L_BYTES + M_BYTES + NS * kKbBytes + STG_BYTES + 1024;
  // TMEM: D3 (mlp.2 accumulator) in columns [0, C); a 256-column scratch at [256, 512) holds D1 (merge accumulator, read by
  // LayerNorm1) and, once that is consumed, the two 128-column D2 buffers of the hidden-layer chunks.
  static constexpr int SCRATCH = 256;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <int C>
__global__ void __launch_bounds__(kThreads, 1) encoder_tail_kernel(const __grid_constant__ EncoderTailParams p) {
  using G = Cfg<C>;
  constexpr int KB = G::KB, NQ = G::NQ, NS = G::NS, NH = G::NH;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bars[2 * NS + 12];
  __shared__ uint32_t s_tmem_base;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sL = base, sM = base + G::L_BYTES, sRing = sM + G::M_BYTES, sStg = sRing + NS * kKbBytes;
  const uint32_t bb = smem_u32(s_bars);
  auto full = [&](int s) { return bb + 8u * s; };
  auto empty = [&](int s) { return bb + 8u * (NS + s); };
  const uint32_t m_full = bb + 8u * (2 * NS), d1_full = m_full + 8, ln1_ready = m_full + 16, d3_full = m_full + 24;
  auto d2_full = [&](int b) { return m_full + 32 + 8u * b; };
  auto h1_full = [&](int b) { return m_full + 48 + 8u * b; };
  auto h1_empty = [&](int b) { return m_full + 64 + 8u * b; };
  auto d13_empty = [&](int b) { return m_full + 80 + 8u * b; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int dbg_n = 0;
  auto stamp = [&](int role, int id) {           // debug timeline of block 0 (first 32 events per role)
    if (p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && dbg_n < 32) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg[role * 64 + 2 * dbg_n] = id;
      p.dbg[role * 64 + 2 * dbg_n + 1] = t;
      ++dbg_n;
    }
  };
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmMsg); tma_prefetch_desc(&p.tmX); tma_prefetch_desc(&p.tmWm); tma_prefetch_desc(&p.tmW0); tma_prefetch_desc(&p.tmW2);
    for (int s = 0; s < 2 * NS; ++s) mbar_init(bb + 8u * s, 1);
    mbar_init(m_full, 1); mbar_init(d1_full, 1); mbar_init(ln1_ready, 4); mbar_init(d3_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(d2_full(b), 1); mbar_init(h1_full(b), 4); mbar_init(h1_empty(b), 1); mbar_init(d13_empty(b), 4);
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

  long long live = p.rows_cap;
  if (p.dyn_rows != nullptr) live = min(live, static_cast<long long>(*p.dyn_rows));
  const int tpi = static_cast<int>((live + 127) / 128);        // tiles per row block
  const int total = p.n_img * tpi;
  auto tile_row0 = [&](int tile, int& r_in_img) -> long long {
    const int img = tile / tpi;
    r_in_img = (tile - img * tpi) * 128;
    return static_cast<long long>(img) * p.rows_cap + r_in_img;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA producer
    // Every ring item is one 16 KB tile [128 rows x 64 bf16]: an A k-block (message or x rows of this tile) or 128 filter
    // rows of a weight k-block. Items are produced in exactly the order the MMA warp consumes them.
    int s = 0;
    uint32_t ph = 0;
    auto put = [&](const CUtensorMap* tm, int c0, int c1) {
      mbar_wait(empty(s), ph ^ 1u, 12);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(full(s), kKbBytes);
        tma_load_2d(sRing + s * kKbBytes, tm, full(s), c0, c1);
      }
      if (++s == NS) { s = 0; ph ^= 1u; }
    };
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      int rin;
      const int row0 = static_cast<int>(tile_row0(tile, rin));
      stamp(2, tile);
      for (int kb = 0; kb < KB; ++kb) {                         // GEMM1: message k-block + merge weights
        put(&p.tmMsg, kb * 64, row0);
        for (int h = 0; h < NH; ++h) put(&p.tmWm, kb * 64, h * 128);
      }
      auto g2 = [&](int q) {                                    // GEMM2 chunk q: W0 rows [128 q, 128 q + 128), all 2C input channels
        for (int kb = 0; kb < 2 * KB; ++kb) {
          if (kb < KB) put(&p.tmX, kb * 64, row0);              // x half of the A operand (the other half is region L)
          put(&p.tmW0, kb * 64, q * 128);
        }
      };
      auto g3 = [&](int q) {                                    // GEMM3 chunk q: W2 input channels [128 q, 128 q + 128)
        for (int j = 0; j < 2; ++j)
          for (int h = 0; h < NH; ++h) put(&p.tmW2, (2 * q + j) * 64, h * 128);
      };
      g2(0);
      for (int q = 1; q < NQ; ++q) { g2(q); g3(q - 1); }
      g3(NQ - 1);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = umma_idesc(1u, 128u, 128u);          // every MMA is 128 x 128 x 16 (bf16)
    int s = 0;
    uint32_t ph = 0;
    int rel_from = 0, n_unrel = 0;                              // consumed ring slots not yet handed back
    // take the next ring item: wait for its bytes, return its shared-memory address
    auto take = [&]() -> uint32_t {
      mbar_wait(full(s), ph, 23);
      const uint32_t a = sRing + s * kKbBytes;
      if (++s == NS) { s = 0; ph ^= 1u; }
      ++n_unrel;
      return a;
    };
    // hand consumed slots back in batches: a tcgen05.commit costs the issuing thread a pipe drain however many barriers it
    // signals, so several slots (and, where one is due, an accumulator barrier) go out together
    auto flush = [&](uint32_t extra_bar) {
      if (elect_one_sync()) {
        int r = rel_from;
        for (int i = 0; i < n_unrel; ++i) {
          tc_commit(empty(r));
          if (++r == NS) r = 0;
        }
        if (extra_bar) tc_commit(extra_bar);
      }
      rel_from += n_unrel;
      if (rel_from >= NS) rel_from -= NS;
      n_unrel = 0;
    };
    auto mma4 = [&](uint32_t d, uint32_t a_addr, uint32_t b_addr, bool first) {
      if (elect_one_sync()) {
        const uint64_t ad = desc_at(a_addr), bd = desc_at(b_addr);
        for (int k = 0; k < 4; ++k) tc_mma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
      }
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const uint32_t d3 = tmem_base;                                       // mlp.2 accumulator
      const uint32_t d1 = tmem_base + G::SCRATCH, d2b = tmem_base + G::SCRATCH;   // merge accumulator, then the two D2 buffers
      stamp(0, 100 * it);
      // ---- GEMM1: D1 = message @ Wm^T  (the scratch is free: every D2 chunk of the previous tile was consumed before its last GEMM3)
      for (int kb = 0; kb < KB; ++kb) {
        const uint32_t a = take();
        tc_fence_after();
        for (int h = 0; h < NH; ++h) {
          const uint32_t b = take();
          tc_fence_after();
          mma4(d1 + 128u * h, a, b, kb == 0);
        }
        if (n_unrel >= kRelBatch) flush(0);
      }
      flush(d1_full);
      stamp(0, 100 * it + 1);
      mbar_wait(ln1_ready, it & 1, 26);          // LayerNorm1 has read D1 (the D2 buffers overwrite it) and written region L
      bool d3_ok = it == 0;
      auto g2 = [&](int q) {
        const uint32_t d2 = d2b + static_cast<uint32_t>((q & 1) * 128);
        for (int kb = 0; kb < 2 * KB; ++kb) {
          uint32_t a;
          if (kb < KB) {
            a = take();
          } else {
            a = sL + (kb - KB) * kKbBytes;
          }
          const uint32_t b = take();
          tc_fence_after();
          mma4(d2, a, b, kb == 0);
          if (n_unrel >= kRelBatch && kb != 2 * KB - 1) flush(0);
        }
        flush(d2_full(q & 1));
        stamp(0, 100 * it + 10 + q);
      };
      auto g3 = [&](int q) {
        const int u = it * (NQ / 2) + (q >> 1);                  // use index of H1 buffer q & 1
        mbar_wait(h1_full(q & 1), u & 1, 27);
        if (!d3_ok) { mbar_wait(d13_empty(0), (it - 1) & 1, 21); d3_ok = true; }   // LayerNorm2 of the previous tile has read D3
        tc_fence_after();
        for (int j = 0; j < 2; ++j)
          for (int h = 0; h < NH; ++h) {
            const uint32_t b = take();
            tc_fence_after();
            mma4(d3 + 128u * h, sM + (q & 1) * 2 * kKbBytes + j * kKbBytes, b, q == 0 && j == 0);
          }
        flush(q == NQ - 1 ? 0u : h1_empty(q & 1));
        if (q == NQ - 1 && elect_one_sync()) { tc_commit(h1_empty(q & 1)); tc_commit(d3_full); }
        stamp(0, 100 * it + 20 + q);
      };
      g2(0);
      for (int q = 1; q < NQ; ++q) { g2(q); g3(q - 1); }
      g3(NQ - 1);
    }
  } else {
    // ------------------------------------------------------------------------------------------ epilogue (thread = row)
    // Two groups of four warps (one warp per TMEM lane quadrant each) that work on DIFFERENT tiles at the same time:
    //   front (warps 2-5): LayerNorm1 of the merge output -> region L, then relu + bf16 of every hidden chunk -> H1 buffers
    //   back  (warps 6-9): LayerNorm2 of the mlp.2 output + residual -> fp32 x and its bf16 shadow in global memory
    // While the back group streams tile i to HBM (the long, bandwidth-bound phase), the front group and the tensor pipe are
    // already in tile i+1; the MMA issuer only waits for the back group before the first GEMM3 of tile i+1 overwrites D3.
    const int q4 = warp & 3;                                    // TMEM lane quadrant of this warp
    const bool back = warp >= 6;
    const int r = q4 * 32 + lane;                               // row of the tile
    const uint32_t lane_off = static_cast<uint32_t>(q4 * 32) << 16;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    const float inv_c = 1.f / static_cast<float>(C);
    // row statistics of a [128 x C] accumulator in one pass: sums of (v - v0), v0 = the row's first element
    auto row_stats = [&](uint32_t acc, float& mean, float& rstd) {
      float v0 = 0.f, s1 = 0.f, s2 = 0.f;
      for (int c0 = 0; c0 < C; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(acc + static_cast<uint32_t>(c0), v);
        tmem_ld_wait();
        if (c0 == 0) v0 = __uint_as_float(v[0]);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = __uint_as_float(v[j]) - v0;
          s1 += d;
          s2 = fmaf(d, d, s2);
        }
      }
      const float md = s1 * inv_c;
      mean = v0 + md;
      rstd = rsqrtf(fmaxf(fmaf(-md, md, s2 * inv_c), 0.f) + 1e-5f);
    };
    if (!back) {
      const uint32_t d1 = tmem_base + G::SCRATCH + lane_off, d2b = d1;
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        // ---- E1: LayerNorm1 of the merge output -> bf16 A operand (region L)
        mbar_wait(d1_full, it & 1, 31);
        tc_fence_after();
        if (warp == 2) stamp(1, 100 * it + 1);
        {
          float mean, rstd;
          row_stats(d1, mean, rstd);
          for (int c0 = 0; c0 < C; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(d1 + static_cast<uint32_t>(c0), v);
            tmem_ld_wait();
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {                         // the same 4 gammas / betas for every lane: broadcast 16-byte loads
              const float4 g = __ldg(reinterpret_cast<const float4*>(p.g1 + c0) + j);
              const float4 bt = __ldg(reinterpret_cast<const float4*>(p.b1 + c0) + j);
              o[2 * j] = pack_bf16x2(fmaf((__uint_as_float(v[4 * j]) - mean) * rstd, g.x, bt.x),
                                     fmaf((__uint_as_float(v[4 * j + 1]) - mean) * rstd, g.y, bt.y));
              o[2 * j + 1] = pack_bf16x2(fmaf((__uint_as_float(v[4 * j + 2]) - mean) * rstd, g.z, bt.z),
                                         fmaf((__uint_as_float(v[4 * j + 3]) - mean) * rstd, g.w, bt.w));
            }
            const uint32_t rowa = sL + (c0 >> 6) * kKbBytes + r * 128;
            const uint32_t cc0 = static_cast<uint32_t>((c0 & 63) >> 3);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + (((cc0 + i) ^ sw) << 4)), "r"(o[4 * i]),
                           "r"(o[4 * i + 1]), "r"(o[4 * i + 2]), "r"(o[4 * i + 3]) : "memory");
          }
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(ln1_ready);
          if (warp == 2) stamp(1, 100 * it + 2);
        }
        // ---- E2: relu + bf16 of hidden chunk q -> H1 buffer q & 1 (A operand of GEMM3)
        for (int q = 0; q < NQ; ++q) {
          const int b = q & 1;
          const int u = it * (NQ / 2) + (q >> 1);
          mbar_wait(d2_full(b), u & 1, 32);
          tc_fence_after();
          if (u >= 1) mbar_wait(h1_empty(b), (u - 1) & 1, 33);   // GEMM3 of the previous use of this buffer has read it
          if (warp == 2) stamp(1, 100 * it + 10 + q);
          const uint32_t acc = d2b + static_cast<uint32_t>(b * 128);
          for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(acc + static_cast<uint32_t>(c0), v);
            tmem_ld_wait();
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              o[j] = pack_bf16x2(fmaxf(__uint_as_float(v[2 * j]), 0.f), fmaxf(__uint_as_float(v[2 * j + 1]), 0.f));
            const uint32_t rowa = sM + b * 2 * kKbBytes + (c0 >> 6) * kKbBytes + r * 128;
            const uint32_t cc0 = static_cast<uint32_t>((c0 & 63) >> 3);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + (((cc0 + i) ^ sw) << 4)), "r"(o[4 * i]),
                           "r"(o[4 * i + 1]), "r"(o[4 * i + 2]), "r"(o[4 * i + 3]) : "memory");
          }
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(h1_full(b));
          if (warp == 2) stamp(1, 100 * it + 20 + q);
        }
      }
    } else {
      // ---- E3: x += LayerNorm2(mlp.2 output); fp32 x and its bf16 shadow, coalesced through this warp's private staging tile
      // (32 rows x 32 floats, 16-byte chunk j of row i at chunk position j ^ (i & 7): conflict-free both ways)
      const uint32_t d3 = tmem_base + lane_off;
      float* stg = reinterpret_cast<float*>(smem_raw + (sStg - smem_u32(smem_raw))) + q4 * 1024;
      const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        int rin;
        const long long row0 = tile_row0(tile, rin);
        // the residual rows of a chunk are requested one chunk ahead of their use (the first before the accumulator is ready):
        // 8 independent 16-byte loads in flight per lane instead of one DRAM round trip per store
        const long long grow0 = row0 + q4 * 32;                  // first row of this warp
        const bool ok = rin + r < live;
        uint32_t okm = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) okm |= (__shfl_sync(0xffffffffu, ok ? 1 : 0, k * 4 + sub_r) ? 1u : 0u) << k;
        float4 xn[8];
        auto load_res = [&](int c0) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            xn[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((okm >> k) & 1u) xn[k] = *reinterpret_cast<const float4*>(p.x + (grow0 + k * 4 + sub_r) * p.ldx + c0 + sub_c);
          }
        };
        load_res(0);
        mbar_wait(d3_full, it & 1, 34);
        tc_fence_after();
        if (warp == 6) stamp(3, 100 * it + 30);
        float mean, rstd;
        row_stats(d3, mean, rstd);
        float mk[8], rk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rr = k * 4 + sub_r;
          mk[k] = __shfl_sync(0xffffffffu, mean, rr);
          rk[k] = __shfl_sync(0xffffffffu, rstd, rr);
        }
        for (int c0 = 0; c0 < C; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(d3 + static_cast<uint32_t>(c0), v);
          tmem_ld_wait();
          if (c0 + 32 >= C) {                                    // last read of D3: the tensor pipe may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d13_empty(0));
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.g2 + c0 + sub_c));
          const float4 bt = __ldg(reinterpret_cast<const float4*>(p.b2 + c0 + sub_c));
          float4 xc[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) xc[k] = xn[k];
          if (c0 + 32 < C) load_res(c0 + 32);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (!((okm >> k) & 1u)) continue;
            const int rr = k * 4 + sub_r;
            const float4 a = *reinterpret_cast<const float4*>(stg + rr * 32 + (((sub_c >> 2) ^ (rr & 7)) << 2));
            float* xr = p.x + (grow0 + rr) * p.ldx + c0 + sub_c;
            const float4 xo = xc[k];
            float4 o;
            o.x = xo.x + fmaf((a.x - mk[k]) * rk[k], g.x, bt.x);
            o.y = xo.y + fmaf((a.y - mk[k]) * rk[k], g.y, bt.y);
            o.z = xo.z + fmaf((a.z - mk[k]) * rk[k], g.z, bt.z);
            o.w = xo.w + fmaf((a.w - mk[k]) * rk[k], g.w, bt.w);
            *reinterpret_cast<float4*>(xr) = o;
            *reinterpret_cast<uint2*>(p.xb + (grow0 + rr) * p.ld_xb + c0 + sub_c) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
          }
          __syncwarp();
        }
        if (warp == 6) stamp(3, 100 * it + 31);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
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

// bf16 [rows][cols] row-major with row stride ld (elements): box {64 columns, box_rows}, 128B swizzle, zero OOB fill
int map2d(CUtensorMap* tm, const void* ptr, long long rows, int cols, long long ld, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return MFR_ERR_CUDA;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld * 2) % 16) return MFR_ERR_ARG;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[mfr_b200] encoder_tail: cuTensorMapEncodeTiled failed: %d\n", static_cast<int>(r));
    return MFR_ERR_CUDA;
  }
  return MFR_OK;
}

}  // namespace

int encoder_tail_build(const EncoderTailDesc& d, EncoderTailParams* p) {
  if ((d.C != 128 && d.C != 256) || d.n_img < 1 || d.rows_cap < 1) return MFR_ERR_ARG;
  if ((d.ldx & 3) || (d.ld_xb & 3) || (reinterpret_cast<uintptr_t>(d.x) & 15) || (reinterpret_cast<uintptr_t>(d.xb_in) & 7)) return MFR_ERR_ARG;
  memset(p, 0, sizeof(*p));
  const long long rows = d.rows_cap * d.n_img;
  int rc;
  if ((rc = map2d(&p->tmMsg, d.msg, rows, d.C, d.ld_msg, 128)) != MFR_OK) return rc;
  if ((rc = map2d(&p->tmX, d.xb_in, rows, d.C, d.ld_xb, 128)) != MFR_OK) return rc;
  if ((rc = map2d(&p->tmWm, d.w_merge, d.C, d.C, d.C, 128)) != MFR_OK) return rc;
  if ((rc = map2d(&p->tmW0, d.w_mlp0, 2 * d.C, 2 * d.C, 2 * d.C, 128)) != MFR_OK) return rc;
  if ((rc = map2d(&p->tmW2, d.w_mlp2, d.C, 2 * d.C, 2 * d.C, 128)) != MFR_OK) return rc;
  p->g1 = d.g1; p->b1 = d.b1; p->g2 = d.g2; p->b2 = d.b2;
  p->x = d.x; p->xb = static_cast<unsigned short*>(const_cast<void*>(d.xb_in));
  p->ldx = d.ldx; p->ld_xb = d.ld_xb;
  p->C = d.C; p->n_img = d.n_img; p->rows_cap = d.rows_cap; p->dyn_rows = d.dyn_rows;
  p->dbg = nullptr;
  return MFR_OK;
}

int encoder_tail_launch(const EncoderTailParams& p, cudaStream_t st) {
  static bool attr_seen[64] = {};
  if (first_call_on_device(attr_seen)) {
    MFR_CUDA_CHECK(cudaFuncSetAttribute(encoder_tail_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<256>::SMEM));
    MFR_CUDA_CHECK(cudaFuncSetAttribute(encoder_tail_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM));
  }
  const long long tiles = static_cast<long long>(p.n_img) * ((p.rows_cap + 127) / 128);
  const int grid = static_cast<int>(std::min<long long>(tiles, conv_gemm_num_sms()));
  if (grid <= 0) return MFR_OK;
  if (p.C == 256) encoder_tail_kernel<256><<<grid, kThreads, Cfg<256>::SMEM, st>>>(p);
  else encoder_tail_kernel<128><<<grid, kThreads, Cfg<128>::SMEM, st>>>(p);
  MFR_CUDA_CHECK(cudaGetLastError());
  return MFR_OK;
}

}  // namespace mfr
