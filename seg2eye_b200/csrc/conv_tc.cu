// Tap-convolution as implicit GEMM on tcgen05 / TMEM, operands staged by TMA (sm_100a).
//
// Forward-like kernel (also used for the data gradient with transposed weights / negated taps):
//   D[128 pixels x BN couts] = sum over taps, 64-channel blocks of  A(tap) [128 x 64] * W(tap)^T [64 x BN]
//   A tile = one 4-D TMA box {64 ch, TW, TH, TB} of the NHWC input at (w0+dx, h0+dy): halo and padding are
//   TMA out-of-bounds zero fill, so no im2col is ever materialised.  Both operands K-major, SWIZZLE_128B.
//   Persistent CTAs; warp 0 = TMA producer, warp 1 = MMA issuer (one lane), warps 2-17 = epilogue
//   (tcgen05.ld -> scale/bias/act -> bf16 -> swizzled smem -> TMA store).  TMEM accumulator double buffered.
//
// Weight-gradient kernel:
//   dW[tap][128 couts x BN cins] += sum over pixel tiles  dY^T [128 x KP] * X(tap) [KP x BN]
//   Both operands are MN-major (channels contiguous, pixels = K), again straight out of 4-D TMA boxes.
//   Split-K over pixel tiles; fp32 accumulators are reduced into global memory with vector red.add.
#include <stdio.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // bf16 elements = 128 bytes = swizzle span
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int OUT_BUF_BYTES = BM * 128;
constexpr int NUM_THREADS = 192;      // weight-gradient kernel: TMA warp, MMA warp, 4 epilogue warps
constexpr int FWD_THREADS = 576;      // forward kernel: TMA warp, MMA warp, 16 epilogue warps
constexpr int FWD_EPI_THREADS = 512;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

// NHWC bf16 tensor [B][H][W][C] -> 4-D map {C, W, H, B}, box {64, bw, bh, bb}
int make_map_nhwc(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int bw, int bh, int bb) {
  EncodeTiledFn enc = get_encode();
  S2E_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  S2E_REQUIRE(((uintptr_t)ptr & 15) == 0 && (C % 8) == 0, "TMA needs 16B-aligned base and C %% 8 == 0 (C=%d)", C);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  S2E_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(nhwc B%d H%d W%d C%d box %d,%d,%d) failed: %d", B, H, W, C, bw,
              bh, bb, (int)r);
  return S2E_OK;
}
// packed weights [T][N][K] bf16 -> 3-D map {K, N, T}, box {64, bn, 1}
int make_map_w(CUtensorMap* m, const void* ptr, int T, int N, int K, int bn) {
  EncodeTiledFn enc = get_encode();
  S2E_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  S2E_REQUIRE(((uintptr_t)ptr & 15) == 0 && (K % 8) == 0, "TMA needs 16B-aligned base and K %% 8 == 0");
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, (cuuint64_t)T};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)N * K * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)bn, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  S2E_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights T%d N%d K%d) failed: %d", T, N, K, (int)r);
  return S2E_OK;
}

struct TapTable {
  int n;
  int dy[S2E_MAX_TAPS];
  int dx[S2E_MAX_TAPS];
};

// ================================================================================ forward-like kernel
struct FwdParams {
  int dbg;  // timing experiments only (debug key 3): bit0 skip TMA store
  int tiles_w, tiles_h, tiles_b, tiles_n, num_tiles, tiles_m;  // tiles_m = pixel tiles; num_tiles = ceil(tiles_m/MT)*tiles_n
  int TW, TH, TB;
  int mt;   // 128-pixel sub-tiles per work item (<= FwdCfg::MT; 1 on maps too small to fill the SMs with double tiles)
  int Cout, kc_per_tap, act;
  int B, Ho, Wo;
  const bf16* mask;   // fused ReLU backward: zero the output where mask <= 0
  const bf16* res;    // fused residual: output += res (same shape as y); never together with mask
  int bias_n;         // valid bias entries (<= Cout; the rest of a channel-padded output gets no bias)
  // fused SPADE+Style modulation (inference): the accumulator holds gamma | beta (sC channels each) and the kernel writes
  // act(0.5 * [(x*ka + kb) * (1 + gamma) + beta + x*kc + s1]) with sC channels instead of gamma | beta
  const bf16* sx;     // block input x, [B][Ho][Wo][sC] or, with sup != 0, its half-resolution source [B][Ho/2][Wo/2][sC]
  const float* spar;  // [B][4][sC]: ka = rstd, kb = -mean*rstd, kc = 1 + s0, s1
  int sC, sact, sup;
  float sscale;       // 0.5 for SPADE+Style (normalization.py:190), 1 for plain SPADE (par rows 2, 3 are zero then)
  int sgamma;         // training: gamma (sC channels, bf16) is written through tmG as well -- backward needs it
  uint8_t* smask;     // training: one bit per output element, set where the output is > 0 ([B*Ho*Wo][sC/8] bytes)
  uint32_t a_box_bytes;
  uint32_t halo_box_bytes;   // HALO kernels: bytes of one {64 ch, TW+2, TH+2, 1} box
  int halo_bo;               // HALO kernels: put (addr >> 7) & 7 into the descriptor's base-offset field
  const float* bias;
  const float* scale;
  TapTable taps;
};

// MT = number of 128-pixel sub-tiles a CTA processes per k-step against ONE weight tile.  The kernel is bound by the
// bytes it can keep in flight (TMA latency x shared-memory capacity), so narrow N tiles get MT = 2: a 256 x 128 tile
// moves the same bytes per FLOP as the 128 x 256 one (ncu: tensor pipe 42 % -> see profiles/).
template <int BN>
struct FwdCfg {
  static constexpr int MT = BN == 256 ? 1 : 2;
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = MT * A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = 4;
  static constexpr int TMEM_COLS = 2 * MT * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * OUT_BUF_BYTES + 1024 + 256 + BN * 4;
};

// HALO variant (3x3 / stride 1, tile = 8 wide x 16 high): per 64-channel block ONE input tile with a one-pixel halo
// ({64 ch, 10, 18, 1} = 23 KB) is staged and serves all nine taps -- the A operand of tap (dy, dx) is the window of that
// tile starting at row (dy+1)*10 + (dx+1), read through a UMMA descriptor whose 8-row group stride (SBO) is the halo
// pitch 10*128 B (tools/umma_shift_probe.cu: the tensor core swizzles on absolute shared-memory address bits, so a
// window may start at any 128-byte row).  Only the weight tiles still stream once per tap: shared-memory fill per FLOP
// drops 2.3x for N = 128 and 3.5x for N = 64 (the N <= 128 kernels were bound by exactly that).
template <int BN>
struct HaloCfg {
  static constexpr int MT = BN == 256 ? 1 : 2;
  static constexpr int TW = 8, TH = 16;
  static constexpr int HALO_ROWS = (TW + 2) * (TH + 2);
  static constexpr int HALO_BYTES = ((HALO_ROWS * 128 + 1023) / 1024) * 1024;
  static constexpr int A_SLOTS = 2;
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int B_STAGES = BN == 64 ? 8 : (BN == 128 ? 6 : 4);
  static constexpr int TMEM_COLS = 2 * MT * BN;
  static constexpr int SMEM_BYTES = A_SLOTS * MT * HALO_BYTES + B_STAGES * B_STAGE_BYTES + 2 * OUT_BUF_BYTES + 1024 + 256 + BN * 4;
};

// pixel sub-tile m -> tile origin; sub-tiles past the end land fully out of bounds (TMA zero-fills loads, clips stores)
__device__ __forceinline__ void subtile_origin(const FwdParams& p, int m, int& w0, int& h0, int& b0) {
  if (m >= p.tiles_m) {
    w0 = 0;
    h0 = 0;
    b0 = p.tiles_b * p.TB + p.TB;
    return;
  }
  const int w_idx = m % p.tiles_w;
  m /= p.tiles_w;
  const int h_idx = m % p.tiles_h;
  const int b_idx = m / p.tiles_h;
  w0 = w_idx * p.TW;
  h0 = h_idx * p.TH;
  b0 = b_idx * p.TB;
}

template <int ACT>
__device__ __forceinline__ float act_t(float v) {
  if (ACT == S2E_ACT_LRELU) return fmaxf(v, 0.2f * v);
  if (ACT == S2E_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

// Per-thread state of the forward kernel's epilogue warps, carried from tile to tile.
struct EpiState {
  int q, sub, row, et;
  bool store_thread;
  float scale;
  int acc, buf;
  uint32_t acc_phase;
  uint32_t tmem_base, out_u32, bias_u32;
  uint8_t* out_buf;
  float* s_bias;
  uint64_t* tmem_full;
  uint64_t* tmem_empty;
};

// One output tile (MT sub-tiles of 128 pixels x BN channels): TMEM -> registers -> scale / bias / activation (or the fused
// SPADE+Style modulation) -> bf16 -> swizzled staging buffer -> TMA store, 64 channels at a time.
// SIDE = per-pixel side input shaped like the output: 0 none, 1 ReLU-backward mask, 2 residual added in fp32.
template <int BN, int ACT, int MODE, int SIDE>
__device__ __forceinline__ void epilogue_tile(const FwdParams& p, EpiState& es, const int t, const CUtensorMap* tmY,
                                              const CUtensorMap* tmG) {
  using Cfg = FwdCfg<BN>;
  constexpr bool SPADE = MODE == 1 || MODE == 2, SPADE_TRAIN = MODE == 2;
  constexpr bool LOADS = SPADE || SIDE != 0;   // reads 16 bf16 per thread and chunk next to the accumulator
  const int q = es.q, sub = es.sub, row = es.row, et = es.et;
  const float scale = es.scale;
  const int n0 = (t % p.tiles_n) * BN;
  const int grp = t / p.tiles_n;
  if (et < BN) {
    float bv = (p.bias && n0 + et < p.bias_n) ? __ldg(p.bias + n0 + et) : 0.f;
    if (SPADE && et >= p.sC && et < 2 * p.sC) {   // fused SPADE: beta's bias absorbs the style offset s1 of this tile's sample
      int w0t, h0t, b0t;
      subtile_origin(p, grp * p.mt, w0t, h0t, b0t);
      bv += __ldg(p.spar + ((size_t)min(b0t, p.B - 1) * 4 + 3) * p.sC + (et - p.sC));
    }
    es.s_bias[et] = bv;
  }
  ptx::mbar_wait(&es.tmem_full[es.acc], es.acc_phase);
  ptx::tc_fence_after();
#pragma unroll 1
  for (int j = 0; j < p.mt; ++j) {
    int w0, h0, b0;
    subtile_origin(p, grp * p.mt + j, w0, h0, b0);
    const bf16* mrow = nullptr;    // this pixel's row of the side input (mask / residual / x of the fused SPADE)
    const float* par = nullptr;
    uint8_t* mask_row = nullptr;   // fused SPADE, training: this pixel's activation-mask bytes
    if (LOADS) {
      const int tw = row % p.TW, r2 = row / p.TW;
      const int th = r2 % p.TH, tb = r2 / p.TH;
      if (tb < p.TB && b0 + tb < p.B && h0 + th < p.Ho && w0 + tw < p.Wo) {
        const size_t pix = ((size_t)(b0 + tb) * p.Ho + h0 + th) * p.Wo + w0 + tw;
        if (SPADE_TRAIN && p.smask) mask_row = p.smask + pix * (size_t)(p.sC >> 3);
        if (SIDE == 1) {
          mrow = p.mask + pix * p.Cout;
        } else if (SIDE == 2) {
          mrow = p.res + pix * p.Cout;
        } else if (p.sup) {   // x lives at half resolution (nearest 2x up-sampling folded in)
          mrow = p.sx + ((((size_t)(b0 + tb) * (p.Ho >> 1) + ((h0 + th) >> 1)) * (p.Wo >> 1) + ((w0 + tw) >> 1)) * p.sC);
        } else {
          mrow = p.sx + pix * p.sC;
        }
      }
      if (SPADE) par = p.spar + (size_t)min(b0, p.B - 1) * 4 * p.sC;   // host guarantees TB == 1: one sample per tile
    }
    const int nch = SPADE ? (p.sC >> 6) : BN / 64;
#pragma unroll 1
    for (int ch = 0; ch < nch; ++ch) {
      const int nbase = n0 + ch * 64;
      if (!SPADE && nbase >= p.Cout) break;
      uint32_t r[16], rb[16];
      const uint32_t taddr =
          es.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(es.acc * Cfg::MT * BN + j * BN + ch * 64 + sub * 16);
      ptx::tmem_ld_32x16(taddr, r);
      if (SPADE) ptx::tmem_ld_32x16(taddr + (uint32_t)p.sC, rb);   // beta sits sC columns after gamma
      constexpr uint32_t dflt = SIDE == 1 ? 0x3f803f80u : 0u;   // pixel outside the map: mask keeps everything, residual / x are zero
      uint4 mk[2] = {make_uint4(dflt, dflt, dflt, dflt), make_uint4(dflt, dflt, dflt, dflt)};
      if (LOADS && mrow) {  // host guarantees Cout % 64 == 0 when a mask / residual is given
#pragma unroll
        for (int i = 0; i < 2; ++i) mk[i] = __ldg(reinterpret_cast<const uint4*>(mrow + nbase + sub * 16) + i);
      }
      // the TMA store that last read this staging buffer must have finished reading it
      if (es.store_thread) {
        if (SPADE_TRAIN) ptx::tma_store_wait_read<0>();   // this chunk fills BOTH staging buffers (output and gamma)
        else ptx::tma_store_wait_read<1>();
      }
      ptx::named_bar_sync(1, FWD_EPI_THREADS);   // also orders the s_bias writes of this tile before the reads below
      ptx::tmem_ld_wait();
      const uint32_t ob = es.out_u32 + (uint32_t)(es.buf * OUT_BUF_BYTES + row * 128);
      const uint32_t bsrc = es.bias_u32 + (uint32_t)((ch * 64 + sub * 16) * 4);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float4 b0v = ptx::lds128f(bsrc + 32 * i), b1v = ptx::lds128f(bsrc + 32 * i + 16);
        uint32_t pk[4];
        float v[8];
        v[0] = act_t<ACT>(fmaf(__uint_as_float(r[8 * i + 0]), scale, b0v.x));
        v[1] = act_t<ACT>(fmaf(__uint_as_float(r[8 * i + 1]), scale, b0v.y));
        v[2] = act_t<ACT>(fmaf(__uint_as_float(r[8 * i + 2]), scale, b0v.z));
        v[3] = act_t<ACT>(fmaf(__uint_as_float(r[8 * i + 3]), scale, b0v.w));
        v[4] = act_t<ACT>(fmaf(__uint_as_float(r[8 * i + 4]), scale, b1v.x));
        v[5] = act_t<ACT>(fmaf(__uint_as_float(r[8 * i + 5]), scale, b1v.y));
        v[6] = act_t<ACT>(fmaf(__uint_as_float(r[8 * i + 6]), scale, b1v.z));
        v[7] = act_t<ACT>(fmaf(__uint_as_float(r[8 * i + 7]), scale, b1v.w));
        if (SIDE == 2) {  // residual add in fp32, one rounding on the sum (bf16 pair: low half = even channel)
          const uint32_t rw[4] = {mk[i].x, mk[i].y, mk[i].z, mk[i].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[2 * e] += __uint_as_float(rw[e] << 16);
            v[2 * e + 1] += __uint_as_float(rw[e] & 0xffff0000u);
          }
        }
        if (SPADE) {  // v = gamma (+bias); form the SPADE+Style output from x, beta and the per-channel constants
          const int c0 = ch * 64 + sub * 16 + 8 * i;
          const uint32_t bb = es.bias_u32 + (uint32_t)((p.sC + c0) * 4);
          const float4 bb0 = ptx::lds128f(bb), bb1 = ptx::lds128f(bb + 16);
          const float betab[8] = {bb0.x, bb0.y, bb0.z, bb0.w, bb1.x, bb1.y, bb1.z, bb1.w};
          const float4* pa = reinterpret_cast<const float4*>(par + c0);
          const float4* pb = reinterpret_cast<const float4*>(par + p.sC + c0);
          const float4* pc = reinterpret_cast<const float4*>(par + 2 * p.sC + c0);
          const float4 a0 = __ldg(pa), a1 = __ldg(pa + 1), k0 = __ldg(pb), k1 = __ldg(pb + 1), c0v = __ldg(pc), c1v = __ldg(pc + 1);
          const float ka[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          const float kb[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
          const float kc[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
          const uint32_t xw[4] = {mk[i].x, mk[i].y, mk[i].z, mk[i].w};
          if (SPADE_TRAIN) {   // gamma itself goes to the other staging buffer (rounded to bf16 exactly like backward reads it)
            const uint32_t og = es.out_u32 + (uint32_t)((es.buf ^ 1) * OUT_BUF_BYTES + row * 128);
            ptx::sts128(og + (uint32_t)(((sub * 2 + i) ^ (row & 7)) << 4), pack2_bf16(v[0], v[1]), pack2_bf16(v[2], v[3]),
                        pack2_bf16(v[4], v[5]), pack2_bf16(v[6], v[7]));
          }
          uint32_t bits = 0;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float xv = __uint_as_float((e & 1) ? (xw[e >> 1] & 0xffff0000u) : (xw[e >> 1] << 16));
            const float beta = fmaf(__uint_as_float(rb[8 * i + e]), scale, betab[e]);   // includes the style offset s1
            float o = p.sscale * (fmaf(fmaf(xv, ka[e], kb[e]), 1.f + v[e], beta) + xv * kc[e]);
            if (p.sact == S2E_ACT_LRELU) o = fmaxf(o, 0.2f * o);
            bits |= (o > 0.f ? 1u : 0u) << e;
            v[e] = o;
          }
          if (SPADE_TRAIN && mask_row) mask_row[(ch * 64 + sub * 16 + 8 * i) >> 3] = (uint8_t)bits;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) pk[e] = pack2_bf16(v[2 * e], v[2 * e + 1]);
        if (SIDE == 1) {  // keep a value only where the bf16 mask element is > 0 (sign clear and magnitude non-zero)
          const uint32_t mw[4] = {mk[i].x, mk[i].y, mk[i].z, mk[i].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t lo = mw[e] & 0xffffu, hi = mw[e] >> 16;
            const uint32_t keep = (((lo & 0x8000u) == 0u && (lo & 0x7fffu) != 0u) ? 0x0000ffffu : 0u) |
                                  (((hi & 0x8000u) == 0u && (hi & 0x7fffu) != 0u) ? 0xffff0000u : 0u);
            pk[e] &= keep;
          }
        }
        ptx::sts128(ob + (uint32_t)(((sub * 2 + i) ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
      }
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(2, FWD_EPI_THREADS);
      if (es.store_thread && !(p.dbg & 1)) {
        ptx::tma_store_4d(tmY, es.out_buf + es.buf * OUT_BUF_BYTES, nbase, w0, h0, b0);
        if (SPADE_TRAIN) ptx::tma_store_4d(tmG, es.out_buf + (es.buf ^ 1) * OUT_BUF_BYTES, nbase, w0, h0, b0);
        ptx::tma_store_commit();
      }
      if (!SPADE_TRAIN) es.buf ^= 1;
    }
  }
  ptx::tc_fence_before();
  ptx::mbar_arrive(&es.tmem_empty[es.acc]);
  es.acc ^= 1;
  if (es.acc == 0) es.acc_phase ^= 1;
}

// Swapped-operand mode (MODE 3, Cout <= 128): the accumulator is D^T -- TMEM lane = output channel, column = pixel (two
// 128-pixel sub-tiles side by side).  A thread owns ONE channel (lane) and 32 pixels (columns) of a sub-tile and scatters
// bf16 values into the pixel-major staging buffers (one per 64-channel half, both filled in the same phase).
template <int ACT, int SIDE>
__device__ __forceinline__ void epilogue_tile_swapped(const FwdParams& p, EpiState& es, const int t, const CUtensorMap* tmY) {
  const int q = es.q, sub = es.sub;
  const int co = es.row;                       // q * 32 + lane: this thread's output channel within the 128-wide tile
  const int n0 = (t % p.tiles_n) * 128;
  const int grp = t / p.tiles_n;
  const int cout = n0 + co;
  const bool live = cout < p.Cout;
  const float bv = (p.bias && cout < p.bias_n) ? __ldg(p.bias + cout) : 0.f;
  const float scale = es.scale;
  ptx::mbar_wait(&es.tmem_full[es.acc], es.acc_phase);
  ptx::tc_fence_after();
  // staging address of (pixel row r, this channel): half (co >> 6), row r * 128, 16-byte chunk ((co & 63) >> 3) ^ (r & 7)
  const int c8 = (co & 63) >> 3;
  uint32_t sbase[8];
#pragma unroll
  for (int m = 0; m < 8; ++m)
    sbase[m] = es.out_u32 + (uint32_t)((co >> 6) * OUT_BUF_BYTES + sub * 32 * 128 + ((c8 ^ m) << 4) + (co & 7) * 2);
#pragma unroll 1
  for (int j = 0; j < p.mt; ++j) {
    int w0, h0, b0;
    subtile_origin(p, grp * p.mt + j, w0, h0, b0);
    uint32_t r[32];
    ptx::tmem_ld_32x32(es.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(es.acc * 256 + j * 128 + sub * 32), r);
    // Side input (mask / residual), pixel-major in memory: all 512 threads copy the tile's 128 x 128 values into the staging
    // buffers with 16-byte accesses (thread = 2 pixel rows x 2 channel halves x 8 channels, the layout the TMA store reads);
    // each channel thread then reads its 32 values from there and overwrites them with the result.  (Per-channel 2-byte global
    // loads, the direct way, cost 1.1 ns per pixel on B200.)
    uint4 sd[4];
    if (SIDE != 0) {
      const bf16* side = SIDE == 1 ? p.mask : p.res;
      constexpr uint32_t dflt = SIDE == 1 ? 0x3f803f80u : 0u;   // pixel outside the map: mask keeps everything, residual is zero
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int prow = (es.et >> 3) + 64 * (v & 1), half = v >> 1;
        const int tw = prow % p.TW, r2 = prow / p.TW;
        const int th = r2 % p.TH, tb = r2 / p.TH;
        sd[v] = make_uint4(dflt, dflt, dflt, dflt);
        if (tb < p.TB && b0 + tb < p.B && h0 + th < p.Ho && w0 + tw < p.Wo && n0 + half * 64 < p.Cout)
          sd[v] = __ldg(reinterpret_cast<const uint4*>(side + (((size_t)(b0 + tb) * p.Ho + h0 + th) * p.Wo + w0 + tw) * p.Cout + n0 +
                                                       half * 64 + (es.et & 7) * 8));
      }
    }
    if (es.store_thread) ptx::tma_store_wait_read<0>();   // both staging buffers are refilled below
    ptx::named_bar_sync(1, FWD_EPI_THREADS);
    if (SIDE != 0) {
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int prow = (es.et >> 3) + 64 * (v & 1), half = v >> 1;
        ptx::sts128(es.out_u32 + (uint32_t)(half * OUT_BUF_BYTES + prow * 128 + (((es.et & 7) ^ (prow & 7)) << 4)), sd[v].x, sd[v].y,
                    sd[v].z, sd[v].w);
      }
      ptx::named_bar_sync(3, FWD_EPI_THREADS);
    }
    ptx::tmem_ld_wait();
    if (live) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const uint32_t addr = sbase[i & 7] + (uint32_t)(i * 128);
        float v = act_t<ACT>(fmaf(__uint_as_float(r[i]), scale, bv));
        uint32_t sv = 0;
        if (SIDE != 0) sv = ptx::lds16(addr);
        if (SIDE == 2) v += __uint_as_float(sv << 16);   // residual add in fp32, one rounding on the sum
        uint32_t h = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
        if (SIDE == 1 && ((sv & 0x8000u) != 0u || (sv & 0x7fffu) == 0u)) h = 0u;   // keep only where the mask is > 0
        ptx::sts16(addr, h);
      }
    }
    ptx::fence_proxy_async_smem();
    ptx::named_bar_sync(2, FWD_EPI_THREADS);
    if (es.store_thread && !(p.dbg & 1)) {
      ptx::tma_store_4d(tmY, es.out_buf, n0, w0, h0, b0);
      if (n0 + 64 < p.Cout) ptx::tma_store_4d(tmY, es.out_buf + OUT_BUF_BYTES, n0 + 64, w0, h0, b0);
      ptx::tma_store_commit();
    }
  }
  ptx::tc_fence_before();
  ptx::mbar_arrive(&es.tmem_empty[es.acc]);
  es.acc ^= 1;
  if (es.acc == 0) es.acc_phase ^= 1;
}

// MODE 0: plain convolution epilogue; 3: the same with swapped MMA operands (weights = A, pixels = B; see below); 1: fused SPADE+Style modulation (inference); 2: the same, also writing gamma and the
// activation mask for backward.  A template parameter so that the ordinary instantiations do not carry the extra code.
template <int BN, int ACT, int MODE, bool HALO>
__global__ void __launch_bounds__(FWD_THREADS, 1)
tapconv_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmG, const FwdParams p) {
  using Cfg = FwdCfg<BN>;
  using HCfg = HaloCfg<BN>;
  constexpr bool SPADE = MODE == 1 || MODE == 2, SPADE_TRAIN = MODE == 2, SWAPPED = MODE == 3;
  static_assert(!SWAPPED || (BN == 128 && !HALO), "swapped-operand mode: N tile 128, no halo staging");
  // ring barriers: non-HALO: full/empty[STAGES] guard (A | B) stages.  HALO: full/empty[0..B_STAGES) guard the weight
  // ring, full/empty[B_STAGES .. B_STAGES + A_SLOTS) guard the halo tiles
  constexpr int NBARS = HALO ? HCfg::B_STAGES + HCfg::A_SLOTS : Cfg::STAGES;
  constexpr int PIPE_BYTES = HALO ? HCfg::A_SLOTS * HCfg::MT * HCfg::HALO_BYTES + HCfg::B_STAGES * HCfg::B_STAGE_BYTES
                                  : Cfg::STAGES * Cfg::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* out_buf = smem + PIPE_BYTES;
  uint64_t* bars = (uint64_t*)(out_buf + 2 * OUT_BUF_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + NBARS;
  uint64_t* tmem_full = bars + 2 * NBARS;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  float* s_bias = (float*)((uint8_t*)bars + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NBARS; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], FWD_EPI_THREADS);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmY);
    if (SPADE_TRAIN) ptx::prefetch_tmap(&tmG);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb = p.taps.n * p.kc_per_tap;

  if (warp == 0 && HALO) {
    // ------------------------------------------------------------------ TMA producer, one halo tile per 64-channel block
    if (lane == 0) {
      int bs = 0, as = 0;
      uint32_t bph = 0, aph = 0;
      uint8_t* sB = smem + HCfg::A_SLOTS * HCfg::MT * HCfg::HALO_BYTES;
      uint64_t* fullA = full + HCfg::B_STAGES;
      uint64_t* emptyA = empty + HCfg::B_STAGES;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int n0 = (t % p.tiles_n) * BN;
        const int grp = t / p.tiles_n;
        int w0[HCfg::MT], h0[HCfg::MT], b0[HCfg::MT];
#pragma unroll
        for (int j = 0; j < HCfg::MT; ++j) subtile_origin(p, grp * HCfg::MT + j, w0[j], h0[j], b0[j]);
        for (int kc = 0; kc < p.kc_per_tap; ++kc) {
          ptx::mbar_wait(&emptyA[as], aph ^ 1);
          ptx::mbar_expect_tx(&fullA[as], HCfg::MT * p.halo_box_bytes);
#pragma unroll
          for (int j = 0; j < HCfg::MT; ++j)
            ptx::tma_load_4d(smem + (as * HCfg::MT + j) * HCfg::HALO_BYTES, &tmA, &fullA[as], kc * BK, w0[j] - 1, h0[j] - 1, b0[j]);
          if (++as == HCfg::A_SLOTS) {
            as = 0;
            aph ^= 1;
          }
          for (int tap = 0; tap < p.taps.n; ++tap) {
            ptx::mbar_wait(&empty[bs], bph ^ 1);
            ptx::mbar_expect_tx(&full[bs], HCfg::B_STAGE_BYTES);
            ptx::tma_load_3d(sB + bs * HCfg::B_STAGE_BYTES, &tmB, &full[bs], kc * BK, n0, tap);
            if (++bs == HCfg::B_STAGES) {
              bs = 0;
              bph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1 && HALO) {
    // ------------------------------------------------------------------ MMA issuer, shifted windows of the halo tile
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN, 0, 0);
    constexpr uint32_t SBO_A = (HCfg::TW + 2) * 128;     // 8-row groups of the window are one halo row (10 pixels) apart
    int bs = 0, as = 0;
    uint32_t bph = 0, aph = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint8_t* sB = smem + HCfg::A_SLOTS * HCfg::MT * HCfg::HALO_BYTES;
    uint64_t* fullA = full + HCfg::B_STAGES;
    uint64_t* emptyA = empty + HCfg::B_STAGES;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * HCfg::MT * BN);
      for (int kc = 0; kc < p.kc_per_tap; ++kc) {
        ptx::mbar_wait(&fullA[as], aph);
        const uint32_t a_base = ptx::smem_u32(smem + as * HCfg::MT * HCfg::HALO_BYTES);
        for (int tap = 0; tap < p.taps.n; ++tap) {
          ptx::mbar_wait(&full[bs], bph);
          ptx::tc_fence_after();
          if (lane == 0) {
            const uint32_t b_addr = ptx::smem_u32(sB + bs * HCfg::B_STAGE_BYTES);
            const uint32_t win = (uint32_t)((p.taps.dy[tap] + 1) * (HCfg::TW + 2) + (p.taps.dx[tap] + 1)) * 128u;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t bd = ptx::umma_desc_sw128(b_addr + k * 32, 0, 1024);
#pragma unroll
              for (int j = 0; j < HCfg::MT; ++j) {
                const uint32_t a_addr = a_base + j * HCfg::HALO_BYTES + win + k * 32;
                uint64_t ad = ptx::umma_desc_sw128(a_addr, 0, SBO_A);
                if (p.halo_bo) ad |= (uint64_t)((a_addr >> 7) & 7u) << 49;
                ptx::umma_bf16(d_tmem + (uint32_t)(j * BN), ad, bd, idesc, (kc | tap | k) != 0 ? 1u : 0u);
              }
            }
            ptx::umma_commit(&empty[bs]);
            if (tap == p.taps.n - 1) ptx::umma_commit(&emptyA[as]);
            if (tap == p.taps.n - 1 && kc == p.kc_per_tap - 1) ptx::umma_commit(&tmem_full[acc]);
          }
          __syncwarp();
          if (++bs == HCfg::B_STAGES) {
            bs = 0;
            bph ^= 1;
          }
        }
        if (++as == HCfg::A_SLOTS) {
          as = 0;
          aph ^= 1;
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int n0 = (t % p.tiles_n) * BN;
        const int grp = t / p.tiles_n;
        int w0[Cfg::MT], h0[Cfg::MT], b0[Cfg::MT];
#pragma unroll
        for (int j = 0; j < Cfg::MT; ++j) subtile_origin(p, grp * p.mt + j, w0[j], h0[j], b0[j]);
        for (int tap = 0; tap < p.taps.n; ++tap) {
          const int dy = p.taps.dy[tap], dx = p.taps.dx[tap];
          for (int kc = 0; kc < p.kc_per_tap; ++kc) {
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            uint8_t* sb = sa + Cfg::MT * A_STAGE_BYTES;
            ptx::mbar_expect_tx(&full[stage], p.mt * p.a_box_bytes + Cfg::B_STAGE_BYTES);
#pragma unroll
            for (int j = 0; j < Cfg::MT; ++j)
              if (j < p.mt)
                ptx::tma_load_4d(sa + j * A_STAGE_BYTES, &tmA, &full[stage], kc * BK, w0[j] + dx, h0[j] + dy, b0[j]);
            ptx::tma_load_3d(sb, &tmB, &full[stage], kc * BK, n0, tap);
            if (++stage == Cfg::STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * Cfg::MT * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::MT * A_STAGE_BYTES;
          if (SWAPPED) {
            // weights are the A operand (M = 128 output channels; rows past Cout are TMA zero fill), the mt pixel sub-tiles
            // -- contiguous in the stage -- one B operand of N = mt * 128: one MMA per k-step.  In SS mode an MMA costs at
            // least the A read (~110 clk for 128 rows x 16), whatever its N; N = 64 / 128 issue could not hide that.
            const uint32_t idesc_s = ptx::umma_idesc_bf16(128, p.mt * 128, 0, 0);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              ptx::umma_bf16(d_tmem, ptx::umma_desc_sw128(b_addr + k * 32, 0, 1024), ptx::umma_desc_sw128(a_addr + k * 32, 0, 1024),
                             idesc_s, (kb | k) != 0 ? 1u : 0u);
          } else
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t bd = ptx::umma_desc_sw128(b_addr + k * 32, 0, 1024);
#pragma unroll
            for (int j = 0; j < Cfg::MT; ++j) {
              if (j >= p.mt) break;
              const uint64_t ad = ptx::umma_desc_sw128(a_addr + j * A_STAGE_BYTES + k * 32, 0, 1024);
              ptx::umma_bf16(d_tmem + (uint32_t)(j * BN), ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          ptx::umma_commit(&empty[stage]);
          if (kb == num_kb - 1) ptx::umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..17)
    // Four warps per TMEM lane quarter (hardware: warp_id % 4 selects the 32 lanes a warp may read); each takes one
    // 16-column slice of every 64-column chunk.  Sixteen warps, not eight: the epilogue is a chain of dependent
    // short-latency steps (tcgen05.ld -> convert -> st.shared -> barrier), so its throughput is set by how many warps
    // the schedulers can interleave (ncu: 41 % issue slots busy with two epilogue warps per scheduler).
    // Bias is staged in shared memory once per tile.  The per-pixel side input (ReLU-backward mask / residual) is a
    // compile-time mode of the tile body: layers with K = 64 (mlp_shared, 1x1 shortcuts) are paced by this epilogue's
    // instruction count, and the mask / residual arithmetic was two thirds of it.
    EpiState es;
    es.q = warp & 3;
    es.sub = (warp - 2) >> 2;          // 0..3: columns [sub*16, sub*16+16) of the chunk
    es.row = es.q * 32 + lane;
    es.et = threadIdx.x - 64;
    es.store_thread = (es.et == 0);
    es.scale = p.scale ? __ldg(p.scale) : 1.0f;
    es.acc = 0;
    es.acc_phase = 0;
    es.buf = 0;
    es.tmem_base = tmem_base;
    es.out_buf = out_buf;
    es.out_u32 = ptx::smem_u32(out_buf);
    es.bias_u32 = ptx::smem_u32(s_bias);
    es.s_bias = s_bias;
    es.tmem_full = tmem_full;
    es.tmem_empty = tmem_empty;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      if (SWAPPED) {
        if (p.mask) epilogue_tile_swapped<ACT, 1>(p, es, t, &tmY);
        else if (p.res) epilogue_tile_swapped<ACT, 2>(p, es, t, &tmY);
        else epilogue_tile_swapped<ACT, 0>(p, es, t, &tmY);
      } else if (SPADE) epilogue_tile<BN, ACT, MODE, 0>(p, es, t, &tmY, &tmG);
      else if (p.mask) epilogue_tile<BN, ACT, MODE, 1>(p, es, t, &tmY, &tmG);
      else if (p.res) epilogue_tile<BN, ACT, MODE, 2>(p, es, t, &tmY, &tmG);
      else epilogue_tile<BN, ACT, MODE, 0>(p, es, t, &tmY, &tmG);
    }
    if (es.store_thread) ptx::tma_store_wait<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

void choose_fwd_tile(int B, int H, int W, int* tw, int* th, int* tb) {
  long long best = -1;
  int bw = 1, bh = 1, bb = 1;
  for (int w = 1; w <= W && w <= BM; ++w) {
    for (int h = 1; h <= H && w * h <= BM; ++h) {
      int b = BM / (w * h);
      if (b > B) b = B;
      if (b < 1) b = 1;
      long long tiles = (long long)ceil_div(W, w) * ceil_div(H, h) * ceil_div(B, b);
      // prefer fewer tiles; tie -> wider rows (longer contiguous TMA runs)
      long long score = tiles * 1024 - w;
      if (best < 0 || score < best) {
        best = score;
        bw = w;
        bh = h;
        bb = b;
      }
    }
  }
  *tw = bw;
  *th = bh;
  *tb = bb;
}

// 3x3 / stride-1 tap set (possibly negated: data gradient) on a map large enough for 8 x 16 tiles
bool halo_eligible(const s2e_conv_t* d) {
  if (d->ntaps != 9 || d->Wo < 8 || d->Ho < 16 || d->Hi != d->Ho || d->Wi != d->Wo) return false;
  if (d->tile_w > 0 && !(d->tile_w == 8 && d->tile_h == 16 && d->tile_b == 1)) return false;
  int seen = 0;
  for (int i = 0; i < 9; ++i) {
    const int dy = d->tap_dy[i], dx = d->tap_dx[i];
    if (dy < -1 || dy > 1 || dx < -1 || dx > 1) return false;
    seen |= 1 << ((dy + 1) * 3 + dx + 1);
  }
  return seen == 0x1ff;
}

template <int BN, int ACT, int MODE, bool HALO>
int launch_fwd(const s2e_conv_t* d, const void* x, const void* wp, const float* bias, const float* scale, void* y,
               cudaStream_t stream) {
  using Cfg = FwdCfg<BN>;
  using HCfg = HaloCfg<BN>;
  constexpr int SMEM = HALO ? HCfg::SMEM_BYTES : Cfg::SMEM_BYTES;
  static_assert(SMEM <= 232448, "shared memory budget");
  int tw = d->tile_w, th = d->tile_h, tb = d->tile_b;
  if (HALO) {
    tw = HCfg::TW;
    th = HCfg::TH;
    tb = 1;
  } else if (tw <= 0 || th <= 0 || tb <= 0) {
    choose_fwd_tile(d->B, d->Ho, d->Wo, &tw, &th, &tb);
  }
  S2E_REQUIRE(tw * th * tb <= BM && tw <= 256 && th <= 256 && tb <= 256, "bad forward tile %dx%dx%d", tw, th, tb);
  CUtensorMap tmA, tmB, tmY, tmG;
  int rc;
  if (HALO) {
    if ((rc = make_map_nhwc(&tmA, x, d->B, d->Hi, d->Wi, d->Cin, tw + 2, th + 2, 1)) != S2E_OK) return rc;
  } else {
    if ((rc = make_map_nhwc(&tmA, x, d->B, d->Hi, d->Wi, d->Cin, tw, th, tb)) != S2E_OK) return rc;
  }
  if ((rc = make_map_w(&tmB, wp, d->ntaps, d->Cout, d->Cin, BN)) != S2E_OK) return rc;
  const bool spade = MODE == 1 || MODE == 2;
  S2E_REQUIRE(spade == (d->spade_x != nullptr) && (MODE == 2) == (spade && d->spade_gamma_out != nullptr), "tapconv_fwd: mode mismatch");
  if (spade) {
    S2E_REQUIRE(d->spade_par && (d->spade_C == 64 || d->spade_C == 128) && d->Cout == 2 * d->spade_C && BN == d->Cout,
                "tapconv_fwd: fused SPADE needs Cout = 2*C = the N tile, C in {64, 128} (C=%d Cout=%d)", d->spade_C, d->Cout);
    S2E_REQUIRE(tb == 1 && d->act == S2E_ACT_NONE && !d->relu_mask && !d->residual && (d->bias_n == 0 || d->bias_n == d->Cout),
                "tapconv_fwd: fused SPADE needs one sample per tile and a plain gamma|beta convolution");
    S2E_REQUIRE(!d->spade_up || (d->Ho % 2 == 0 && d->Wo % 2 == 0), "tapconv_fwd: fused SPADE with up-sampling needs even H, W");
  }
  if ((rc = make_map_nhwc(&tmY, y, d->B, d->Ho, d->Wo, spade ? d->spade_C : d->Cout, tw, th, tb)) != S2E_OK) return rc;
  tmG = tmY;
  if (spade && d->spade_gamma_out &&
      (rc = make_map_nhwc(&tmG, d->spade_gamma_out, d->B, d->Ho, d->Wo, d->spade_C, tw, th, tb)) != S2E_OK)
    return rc;
  FwdParams p;
  p.dbg = s2e_debug_get(3);
  p.tiles_w = ceil_div(d->Wo, tw);
  p.tiles_h = ceil_div(d->Ho, th);
  p.tiles_b = ceil_div(d->B, tb);
  p.tiles_n = ceil_div(d->Cout, BN);
  p.tiles_m = p.tiles_w * p.tiles_h * p.tiles_b;
  p.mt = Cfg::MT;
  if (!HALO && !spade && Cfg::MT == 2 && !(s2e_debug_get(6) & 4)) {
    // double tiles halve the number of work items: on small maps (the low-resolution gamma|beta data gradients: 120 or 30
    // single tiles) they leave most SMs idle.  A single tile costs ~0.6 of a double one (the weight tile is not shared).
    const int sms = s2e_num_sms();
    const long long w1 = (long long)ceil_div(p.tiles_m * p.tiles_n, sms) * 6, w2 = (long long)ceil_div(ceil_div(p.tiles_m, 2) * p.tiles_n, sms) * 10;
    if (w1 < w2) p.mt = 1;
  }
  p.num_tiles = ceil_div(p.tiles_m, p.mt) * p.tiles_n;
  p.TW = tw;
  p.TH = th;
  p.TB = tb;
  p.Cout = d->Cout;
  p.kc_per_tap = d->Cin / BK;
  p.act = d->act;
  p.B = d->B;
  p.Ho = d->Ho;
  p.Wo = d->Wo;
  p.mask = (const bf16*)d->relu_mask;
  p.res = (const bf16*)d->residual;
  p.bias_n = d->bias_n > 0 ? d->bias_n : d->Cout;
  p.sx = (const bf16*)d->spade_x;
  p.spar = d->spade_par;
  p.sC = d->spade_C;
  p.sact = d->spade_act;
  p.sup = d->spade_up;
  p.sscale = d->spade_plain ? 1.0f : 0.5f;
  p.sgamma = (spade && d->spade_gamma_out) ? 1 : 0;
  p.smask = spade ? (uint8_t*)d->spade_mask_out : nullptr;
  S2E_REQUIRE(!(p.mask && p.res), "tapconv_fwd: relu_mask and residual are mutually exclusive");
  S2E_REQUIRE(d->in_act == S2E_ACT_NONE && d->mask_slope == 0.f && !d->img_out, "tapconv_fwd: in_act / mask_slope / image head exist on the CUDA-core path only");
  S2E_REQUIRE(!(p.mask || p.res) || d->Cout % 64 == 0, "tapconv_fwd: relu_mask / residual need Cout %% 64 == 0 on the tcgen05 path");
  // fused SPADE: the MT sub-tiles of one work item must belong to one sample; a partial last row of tiles is fine (TMA clips)
  S2E_REQUIRE(!spade || (p.tiles_w * p.tiles_h) % p.mt == 0, "tapconv_fwd: fused SPADE: sub-tiles of one tile must share a sample");
  p.a_box_bytes = (uint32_t)(tw * th * tb * BK * 2);
  p.halo_box_bytes = (uint32_t)((tw + 2) * (th + 2) * BK * 2);
  p.halo_bo = s2e_debug_get(6) & 2 ? 1 : 0;
  p.bias = bias;
  p.scale = scale;
  p.taps.n = d->ntaps;
  for (int i = 0; i < d->ntaps; ++i) {
    p.taps.dy[i] = d->tap_dy[i];
    p.taps.dx[i] = d->tap_dx[i];
  }
  static int attr_dev_mask = 0;     // per device: one process may drive several GPUs
  int dev = 0;
  S2E_CHECK_CUDA(cudaGetDevice(&dev));
  if (!(attr_dev_mask & (1 << (dev & 31)))) {
    S2E_CHECK_CUDA(cudaFuncSetAttribute(tapconv_fwd_kernel<BN, ACT, MODE, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_dev_mask |= 1 << (dev & 31);
  }
  int grid = p.num_tiles < s2e_num_sms() ? p.num_tiles : s2e_num_sms();
  tapconv_fwd_kernel<BN, ACT, MODE, HALO><<<grid, FWD_THREADS, SMEM, stream>>>(tmA, tmB, tmY, tmG, p);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

// ================================================================================ weight-gradient kernel
struct WgParams {
  int swap;  // 0: D[Cout x Cin] = dY^T X ; 1: D[Cin x Cout] = X^T dY (chosen so that the N tile is as wide as possible)
  int Cout, Cin;
  int mt, nt, ksplit;            // tiles over Cout (128), Cin (BN), split-K
  int kt_w, kt_h, kt_b, kt_total;  // pixel tiles
  int KTW, KTH, KTB, KP;
  int swap_lbo_sbo;
  int merge_taps;   // multi-tap kernel, 64-channel taps on the N side: one MMA of N = ntap * 64 per k-step
  float* dwp;
  TapTable taps;
};

template <int BN>
struct WgCfg {
  static constexpr int A_BYTES = 2 * 64 * 128;          // 128 couts x 64 pixels max
  static constexpr int B_BYTES = (BN / 64) * 64 * 128;  // BN cins x 64 pixels max
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN == 256 ? 4 : 6;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tapconv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgParams p) {
  using Cfg = WgCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* acc_full = bars + 2 * Cfg::STAGES;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(acc_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmDY);
    ptx::prefetch_tmap(&tmX);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item decode: blockIdx.x = ((ks * mt + m) * nt + n) * ntaps + tap -- the taps of one pixel range are adjacent
  // CTAs, i.e. co-resident, so the dY / X tiles they share are fetched from HBM once and served from L2 to the others
  int wi = blockIdx.x;
  const int tap = wi % p.taps.n;
  wi /= p.taps.n;
  const int n_idx = wi % p.nt;
  wi /= p.nt;
  const int m_idx = wi % p.mt;
  const int ks = wi / p.mt;
  const int m0 = m_idx * 128, n0 = n_idx * BN;
  const int k_begin = (int)(((long long)p.kt_total * ks) / p.ksplit);
  const int k_end = (int)(((long long)p.kt_total * (ks + 1)) / p.ksplit);
  const int nk = k_end - k_begin;
  const uint32_t blk_bytes = (uint32_t)p.KP * 128u;  // one 64-channel block of KP pixels

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int dy = p.taps.dy[tap], dx = p.taps.dx[tap];
      for (int kt = k_begin; kt < k_end; ++kt) {
        int r = kt;
        const int w_idx = r % p.kt_w;
        r /= p.kt_w;
        const int h_idx = r % p.kt_h;
        const int b_idx = r / p.kt_h;
        const int w0 = w_idx * p.KTW, h0 = h_idx * p.KTH, b0 = b_idx * p.KTB;
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        ptx::mbar_expect_tx(&full[stage], blk_bytes * (2 + BN / 64));
        const CUtensorMap* mapA = p.swap ? &tmX : &tmDY;
        const CUtensorMap* mapB = p.swap ? &tmDY : &tmX;
        const int ax = p.swap ? dx : 0, ay = p.swap ? dy : 0, bx = p.swap ? 0 : dx, by = p.swap ? 0 : dy;
#pragma unroll
        for (int j = 0; j < 2; ++j) ptx::tma_load_4d(sa + j * blk_bytes, mapA, &full[stage], m0 + j * 64, w0 + ax, h0 + ay, b0);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j)
          ptx::tma_load_4d(sb + j * blk_bytes, mapB, &full[stage], n0 + j * 64, w0 + bx, h0 + by, b0);
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, BN, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    const int kmma = p.KP / 16;
    const uint32_t lbo = p.swap_lbo_sbo ? 1024u : blk_bytes;
    const uint32_t sbo = p.swap_lbo_sbo ? blk_bytes : 1024u;
    for (int i = 0; i < nk; ++i) {
      ptx::mbar_wait(&full[stage], phase);
      ptx::tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t b_addr = a_addr + Cfg::A_BYTES;
        for (int k = 0; k < kmma; ++k) {
          const uint64_t ad = ptx::umma_desc_sw128(a_addr + k * 2048, lbo, sbo);
          const uint64_t bd = ptx::umma_desc_sw128(b_addr + k * 2048, lbo, sbo);
          ptx::umma_bf16(tmem_base, ad, bd, idesc, (i | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty[stage]);
        if (i == nk - 1) ptx::umma_commit(acc_full);
      }
      __syncwarp();
      if (++stage == Cfg::STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (nk > 0) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int mrow = m0 + row;
    const int Mdim = p.swap ? p.Cin : p.Cout, Ndim = p.swap ? p.Cout : p.Cin;
    ptx::mbar_wait(acc_full, 0);
    ptx::tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= Ndim) break;  // warp-uniform
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      ptx::tmem_ld_wait();
      if (mrow < Mdim) {
        if (!p.swap) {
          float* dst_row = p.dwp + ((size_t)tap * p.Cout + mrow) * p.Cin + n0 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (n0 + c0 + j + 3 < Ndim) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_row + j), "f"(__uint_as_float(r[j])),
                           "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                           : "memory");
            } else {
              for (int e = 0; e < 4; ++e)
                if (n0 + c0 + j + e < Ndim) atomicAdd(dst_row + j + e, __uint_as_float(r[j + e]));
            }
          }
        } else {
          // accumulator row = input channel (contiguous across the warp's lanes -> coalesced reductions)
          float* dst = p.dwp + ((size_t)tap * p.Cout + n0 + c0) * p.Cin + mrow;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c0 + j < Ndim) atomicAdd(dst + (size_t)j * p.Cin, __uint_as_float(r[j]));
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

void choose_k_tile(int B, int H, int W, int* tw, int* th, int* tb) {
  long long best = -1;
  int bw = 16, bh = 1, bb = 1;
  for (int w = 1; w <= 64; ++w) {
    for (int h = 1; w * h <= 64; ++h) {
      for (int b = 1; w * h * b <= 64; ++b) {
        int prod = w * h * b;
        if (prod % 16) continue;
        if ((w > W && w > 16) || (h > H && h > 1 && prod > 16) || (b > B && b > 1)) continue;
        long long tiles = (long long)ceil_div(W, w) * ceil_div(H, h) * ceil_div(B, b);
        long long padded = tiles * prod;
        // Boxes whose rows are one or two pixels long (the zero-padding choice for the odd-sized discriminator maps, e.g.
        // {1, 2, 32} for 82 x 50) load far slower than boxes of >= 4 adjacent pixels with a few per cent of padding
        // (B200, B32 82x50 256->512 4x4: 0.95 ms with {1,2,32}, 0.51-0.53 ms with {2,2,16} / {4,2,8}; B32 161x97 256->128
        // 2x2: 0.334 / 0.178 / 0.131 ms with w = 1 / 2 / 4) -- profiles/r02e_ktile_probe.txt
        const int pen = w >= 4 ? 100 : (w == 3 ? 108 : (w == 2 ? 120 : 160));
        long long score = padded * pen * 41 + tiles * 8 - (w >= 8 ? 1 : 0);
        if (best < 0 || score < best) {
          best = score;
          bw = w;
          bh = h;
          bb = b;
        }
      }
    }
  }
  *tw = bw;
  *th = bh;
  *tb = bb;
}

template <int BN>
int launch_wgrad(const s2e_conv_t* d, const void* x, const void* dy, float* dwp, int swap, cudaStream_t stream) {
  using Cfg = WgCfg<BN>;
  int tw = d->ktile_w, th = d->ktile_h, tb = d->ktile_b;
  if (tw <= 0 || th <= 0 || tb <= 0) choose_k_tile(d->B, d->Ho, d->Wo, &tw, &th, &tb);
  const int KP = tw * th * tb;
  S2E_REQUIRE(KP % 16 == 0 && KP <= 64, "bad wgrad pixel tile %dx%dx%d", tw, th, tb);
  CUtensorMap tmDY, tmX;
  int rc;
  if ((rc = make_map_nhwc(&tmDY, dy, d->B, d->Ho, d->Wo, d->Cout, tw, th, tb)) != S2E_OK) return rc;
  if ((rc = make_map_nhwc(&tmX, x, d->B, d->Hi, d->Wi, d->Cin, tw, th, tb)) != S2E_OK) return rc;
  WgParams p;
  p.swap = swap;
  p.Cout = d->Cout;
  p.Cin = d->Cin;
  p.mt = ceil_div(swap ? d->Cin : d->Cout, 128);
  p.nt = ceil_div(swap ? d->Cout : d->Cin, BN);
  p.kt_w = ceil_div(d->Wo, tw);
  p.kt_h = ceil_div(d->Ho, th);
  p.kt_b = ceil_div(d->B, tb);
  p.kt_total = p.kt_w * p.kt_h * p.kt_b;
  p.KTW = tw;
  p.KTH = th;
  p.KTB = tb;
  p.KP = KP;
  p.swap_lbo_sbo = s2e_debug_get(0);
  p.merge_taps = 0;
  p.dwp = dwp;
  p.taps.n = d->ntaps;
  for (int i = 0; i < d->ntaps; ++i) {
    p.taps.dy[i] = d->tap_dy[i];
    p.taps.dx[i] = d->tap_dx[i];
  }
  // split-K so that the CTA count fills whole waves of SMs (a 297-CTA launch on 148 SMs runs 3 waves, the last with
  // one CTA): try 1..4 waves, keep >= 4 pixel tiles per CTA, take the best fill (ties -> fewer waves, fewer atomics)
  const int base = d->ntaps * p.mt * p.nt;
  const int sms = s2e_num_sms();
  int max_split = p.kt_total / 4;
  if (max_split < 1) max_split = 1;
  int ksplit = 1;
  double best_fill = -1.0;
  for (int w = 1; w <= 4; ++w) {
    int ks = (sms * w) / base;
    if (ks < 1) ks = 1;
    if (ks > max_split) ks = max_split;
    const int ctas = base * ks;
    const int waves = ceil_div(ctas, sms);
    const double fill = (double)ctas / ((double)waves * sms);
    if (fill > best_fill + 0.02) {
      best_fill = fill;
      ksplit = ks;
    }
  }
  p.ksplit = ksplit;
  static int attr_dev_mask = 0;     // per device: one process may drive several GPUs
  int dev = 0;
  S2E_CHECK_CUDA(cudaGetDevice(&dev));
  if (!(attr_dev_mask & (1 << (dev & 31)))) {
    S2E_CHECK_CUDA(cudaFuncSetAttribute(tapconv_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_dev_mask |= 1 << (dev & 31);
  }
  tapconv_wgrad_kernel<BN><<<base * ksplit, NUM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmDY, tmX, p);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

// ================================================================================ weight gradient, several taps per CTA
// One CTA accumulates TPC taps of the same (M, N) tile (the default for N sides < 256).
// The tap-independent operand (dY) is staged ONCE per pixel tile and multiplied against TPC shifted copies of X, each
// into its own TMEM accumulator, instead of TPC CTAs re-reading dY through L2.  Same operand layouts, descriptors and
// split-K reduction as tapconv_wgrad_kernel; TPC * BN <= 512 TMEM columns.
constexpr int WG_BLK = 64 * 128;   // one 64-channel block of up to 64 pixels

template <int BN, int TPC>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tapconv_wgrad_mt_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgParams p,
                        int stages, int ngroups) {
  constexpr int MAX_STAGES = 6;
  constexpr int TMEM_COLS = TPC * BN <= 32 ? 32 : (TPC * BN <= 64 ? 64 : (TPC * BN <= 128 ? 128 : (TPC * BN <= 256 ? 256 : 512)));
  static_assert(TPC * BN <= 512, "accumulators of all taps must fit TMEM");
  constexpr int A_BLOCKS = 2, B_BLOCKS = BN / 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // stage = [shared operand (dY)] [tap 0 operand (X)] ... [tap TPC-1 operand]; which of A / B is shared depends on swap
  const int shared_blocks = p.swap ? B_BLOCKS : A_BLOCKS;
  const int tap_blocks = p.swap ? A_BLOCKS : B_BLOCKS;
  const int stage_bytes = (shared_blocks + TPC * tap_blocks) * WG_BLK;
  uint64_t* bars = (uint64_t*)(smem + stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + MAX_STAGES;
  uint64_t* acc_full = bars + 2 * MAX_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(acc_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmDY);
    ptx::prefetch_tmap(&tmX);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item decode: blockIdx.x = ((ks * mt + m) * nt + n) * ngroups + tap group
  int wi = blockIdx.x;
  const int tg = wi % ngroups;
  wi /= ngroups;
  const int n_idx = wi % p.nt;
  wi /= p.nt;
  const int m_idx = wi % p.mt;
  const int ks = wi / p.mt;
  const int tap0 = tg * TPC;
  const int ntap = min(TPC, p.taps.n - tap0);
  const int m0 = m_idx * 128, n0 = n_idx * BN;
  const int k_begin = (int)(((long long)p.kt_total * ks) / p.ksplit);
  const int k_end = (int)(((long long)p.kt_total * (ks + 1)) / p.ksplit);
  const int nk = k_end - k_begin;
  const uint32_t blk_bytes = (uint32_t)p.KP * 128u;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const CUtensorMap* map_shared = &tmDY;
      const CUtensorMap* map_tap = &tmX;
      const int c_shared = p.swap ? n0 : m0;   // channel origin of the shared operand (dY: Cout axis)
      const int c_tap = p.swap ? m0 : n0;      // channel origin of the per-tap operand (X: Cin axis)
      for (int kt = k_begin; kt < k_end; ++kt) {
        int r = kt;
        const int w_idx = r % p.kt_w;
        r /= p.kt_w;
        const int h_idx = r % p.kt_h;
        const int b_idx = r / p.kt_h;
        const int w0 = w_idx * p.KTW, h0 = h_idx * p.KTH, b0 = b_idx * p.KTB;
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* ss = smem + stage * stage_bytes;
        ptx::mbar_expect_tx(&full[stage], blk_bytes * (uint32_t)(shared_blocks + ntap * tap_blocks));
        for (int j = 0; j < shared_blocks; ++j)
          ptx::tma_load_4d(ss + j * blk_bytes, map_shared, &full[stage], c_shared + j * 64, w0, h0, b0);
        for (int tt = 0; tt < ntap; ++tt) {
          uint8_t* st = ss + (shared_blocks + tt * tap_blocks) * WG_BLK;
          const int dy = p.taps.dy[tap0 + tt], dx = p.taps.dx[tap0 + tt];
          for (int j = 0; j < tap_blocks; ++j)
            ptx::tma_load_4d(st + j * blk_bytes, map_tap, &full[stage], c_tap + j * 64, w0 + dx, h0 + dy, b0);
        }
        if (++stage == stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, BN, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    const int kmma = p.KP / 16;
    const uint32_t lbo = p.swap_lbo_sbo ? 1024u : blk_bytes;
    const uint32_t sbo = p.swap_lbo_sbo ? blk_bytes : 1024u;
    for (int i = 0; i < nk; ++i) {
      ptx::mbar_wait(&full[stage], phase);
      ptx::tc_fence_after();
      if (lane == 0) {
        const uint32_t s_addr = ptx::smem_u32(smem + stage * stage_bytes);
        if (p.merge_taps) {
          // taps on the N side whose operand blocks sit WG_BLK apart ARE one MN-major operand of g * BN columns (64-column
          // atoms WG_BLK apart): one MMA of N <= 256 per k-step covers g = 256 / BN taps and reads the shared dY operand once.
          // In SS mode an MMA costs at least its A read (~110 clk for 128 rows x 16) whatever its N, so three N = 64 MMAs take
          // 330 clk where one N = 192 MMA takes 110.
          constexpr int G = 256 / BN;
          for (int tt = 0; tt < ntap; tt += G) {
            const int g = min(G, ntap - tt);
            const uint32_t idesc_n = ptx::umma_idesc_bf16(128, g * BN, 1, 1);
            const uint32_t t_addr = s_addr + (uint32_t)((shared_blocks + tt * tap_blocks) * WG_BLK);
            for (int k = 0; k < kmma; ++k) {
              const uint64_t ad = ptx::umma_desc_sw128(s_addr + k * 2048, lbo, sbo);
              const uint64_t bd = ptx::umma_desc_sw128(t_addr + k * 2048, (uint32_t)WG_BLK, sbo);
              ptx::umma_bf16(tmem_base + (uint32_t)(tt * BN), ad, bd, idesc_n, (i | k) != 0 ? 1u : 0u);
            }
          }
        } else
        for (int tt = 0; tt < ntap; ++tt) {
          const uint32_t t_addr = s_addr + (uint32_t)((shared_blocks + tt * tap_blocks) * WG_BLK);
          const uint32_t a_addr = p.swap ? t_addr : s_addr;   // A = M side: dY (shared) unless swapped
          const uint32_t b_addr = p.swap ? s_addr : t_addr;
          for (int k = 0; k < kmma; ++k) {
            const uint64_t ad = ptx::umma_desc_sw128(a_addr + k * 2048, lbo, sbo);
            const uint64_t bd = ptx::umma_desc_sw128(b_addr + k * 2048, lbo, sbo);
            ptx::umma_bf16(tmem_base + (uint32_t)(tt * BN), ad, bd, idesc, (i | k) != 0 ? 1u : 0u);
          }
        }
        ptx::umma_commit(&empty[stage]);
        if (i == nk - 1) ptx::umma_commit(acc_full);
      }
      __syncwarp();
      if (++stage == stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (nk > 0) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int mrow = m0 + row;
    const int Mdim = p.swap ? p.Cin : p.Cout, Ndim = p.swap ? p.Cout : p.Cin;
    ptx::mbar_wait(acc_full, 0);
    ptx::tc_fence_after();
#pragma unroll 1
    for (int tt = 0; tt < ntap; ++tt) {
      const int tap = tap0 + tt;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= Ndim) break;  // warp-uniform
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tt * BN + c0), r);
        ptx::tmem_ld_wait();
        if (mrow < Mdim) {
          if (!p.swap) {
            float* dst_row = p.dwp + ((size_t)tap * p.Cout + mrow) * p.Cin + n0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (n0 + c0 + j + 3 < Ndim) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_row + j), "f"(__uint_as_float(r[j])),
                             "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                             : "memory");
              } else {
                for (int e = 0; e < 4; ++e)
                  if (n0 + c0 + j + e < Ndim) atomicAdd(dst_row + j + e, __uint_as_float(r[j + e]));
              }
            }
          } else {
            float* dst = p.dwp + ((size_t)tap * p.Cout + n0 + c0) * p.Cin + mrow;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c0 + j < Ndim) atomicAdd(dst + (size_t)j * p.Cin, __uint_as_float(r[j]));
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, int TPC>
int launch_wgrad_mt(const s2e_conv_t* d, const void* x, const void* dy, float* dwp, int swap, cudaStream_t stream) {
  int tw = d->ktile_w, th = d->ktile_h, tb = d->ktile_b;
  if (tw <= 0 || th <= 0 || tb <= 0) choose_k_tile(d->B, d->Ho, d->Wo, &tw, &th, &tb);
  const int KP = tw * th * tb;
  S2E_REQUIRE(KP % 16 == 0 && KP <= 64, "bad wgrad pixel tile %dx%dx%d", tw, th, tb);
  CUtensorMap tmDY, tmX;
  int rc;
  if ((rc = make_map_nhwc(&tmDY, dy, d->B, d->Ho, d->Wo, d->Cout, tw, th, tb)) != S2E_OK) return rc;
  if ((rc = make_map_nhwc(&tmX, x, d->B, d->Hi, d->Wi, d->Cin, tw, th, tb)) != S2E_OK) return rc;
  WgParams p;
  p.swap = swap;
  p.Cout = d->Cout;
  p.Cin = d->Cin;
  p.mt = ceil_div(swap ? d->Cin : d->Cout, 128);
  p.nt = ceil_div(swap ? d->Cout : d->Cin, BN);
  p.kt_w = ceil_div(d->Wo, tw);
  p.kt_h = ceil_div(d->Ho, th);
  p.kt_b = ceil_div(d->B, tb);
  p.kt_total = p.kt_w * p.kt_h * p.kt_b;
  p.KTW = tw;
  p.KTH = th;
  p.KTB = tb;
  p.KP = KP;
  p.swap_lbo_sbo = s2e_debug_get(0);
  p.merge_taps = 0;
  p.dwp = dwp;
  p.taps.n = d->ntaps;
  for (int i = 0; i < d->ntaps; ++i) {
    p.taps.dy[i] = d->tap_dy[i];
    p.taps.dx[i] = d->tap_dx[i];
  }
  // (BN = 128: the two 64-channel blocks of a tap are blk_bytes apart, the taps 2 * WG_BLK: uniform only for 64-pixel k tiles)
  p.merge_taps = ((BN == 64 || KP * 128 == WG_BLK) && BN <= 128 && !swap && !p.swap_lbo_sbo && s2e_debug_get(5) != 2) ? 1 : 0;
  const int ngroups = ceil_div(d->ntaps, TPC);
  const int shared_blocks = swap ? BN / 64 : 2, tap_blocks = swap ? 2 : BN / 64;
  const int stage_bytes = (shared_blocks + TPC * tap_blocks) * WG_BLK;
  int stages = (232448 - 1024 - 256) / stage_bytes;
  if (stages > 6) stages = 6;
  S2E_REQUIRE(stages >= 2, "wgrad_mt: stage of %d bytes leaves no pipeline", stage_bytes);
  const int smem_bytes = stages * stage_bytes + 1024 + 256;
  const int base = ngroups * p.mt * p.nt;
  const int sms = s2e_num_sms();
  int max_split = p.kt_total / 4;
  if (max_split < 1) max_split = 1;
  int ksplit = 1;
  double best_fill = -1.0;
  for (int w = 1; w <= 4; ++w) {
    int ks = (sms * w) / base;
    if (ks < 1) ks = 1;
    if (ks > max_split) ks = max_split;
    const int ctas = base * ks;
    const int waves = ceil_div(ctas, sms);
    const double fill = (double)ctas / ((double)waves * sms);
    if (fill > best_fill + 0.02) {
      best_fill = fill;
      ksplit = ks;
    }
  }
  p.ksplit = ksplit;
  static int attr_dev_mask = 0;
  int dev = 0;
  S2E_CHECK_CUDA(cudaGetDevice(&dev));
  if (!(attr_dev_mask & (1 << (dev & 31)))) {
    S2E_CHECK_CUDA(cudaFuncSetAttribute(tapconv_wgrad_mt_kernel<BN, TPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_dev_mask |= 1 << (dev & 31);
  }
  tapconv_wgrad_mt_kernel<BN, TPC><<<base * ksplit, NUM_THREADS, smem_bytes, stream>>>(tmDY, tmX, p, stages, ngroups);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

}  // namespace

int s2e_tapconv_fwd_tc(const s2e_conv_t* d, const void* x, const void* wp, const float* bias, const float* scale,
                       void* y, cudaStream_t stream) {
  S2E_REQUIRE(d->Cin % 64 == 0 && d->Cout % 8 == 0, "tcgen05 tapconv needs Cin %% 64 == 0, Cout %% 8 == 0 (Cin=%d Cout=%d)",
              d->Cin, d->Cout);
  // debug key 6, bit 0: 3x3 / stride-1 layers with an N tile <= 128 take the halo-tile kernel (one staged input tile per
  // 64-channel block serves all nine taps); bit 1: descriptors carry an explicit base offset
  const bool halo = (s2e_debug_get(6) & 1) && halo_eligible(d);
#define S2E_FWD_DISPATCH(BN_, HALO_)                                                                          \
  switch (d->act) {                                                                                           \
    case S2E_ACT_LRELU: return launch_fwd<BN_, S2E_ACT_LRELU, 0, HALO_>(d, x, wp, bias, scale, y, stream);    \
    case S2E_ACT_RELU: return launch_fwd<BN_, S2E_ACT_RELU, 0, HALO_>(d, x, wp, bias, scale, y, stream);      \
    default: return launch_fwd<BN_, S2E_ACT_NONE, 0, HALO_>(d, x, wp, bias, scale, y, stream);                \
  }
  if (d->spade_x) {   // fused SPADE+Style epilogue: gamma | beta fill exactly one N tile
    S2E_REQUIRE(d->act == S2E_ACT_NONE && (d->Cout == 256 || d->Cout == 128), "tapconv_fwd: fused SPADE needs Cout in {128, 256}");
    if (d->spade_gamma_out) {
      if (d->Cout == 256) return launch_fwd<256, S2E_ACT_NONE, 2, false>(d, x, wp, bias, scale, y, stream);
      if (halo) return launch_fwd<128, S2E_ACT_NONE, 2, true>(d, x, wp, bias, scale, y, stream);
      return launch_fwd<128, S2E_ACT_NONE, 2, false>(d, x, wp, bias, scale, y, stream);
    }
    if (d->Cout == 256) return launch_fwd<256, S2E_ACT_NONE, 1, false>(d, x, wp, bias, scale, y, stream);
    if (halo) return launch_fwd<128, S2E_ACT_NONE, 1, true>(d, x, wp, bias, scale, y, stream);
    return launch_fwd<128, S2E_ACT_NONE, 1, false>(d, x, wp, bias, scale, y, stream);
  }
  if (d->Cout >= 256) { S2E_FWD_DISPATCH(256, false) }
  // Cout <= 128: swapped operands (MODE 3) unless debug key 6 bit 3 is set or the halo kernel is asked for.  With Cout = 64 only
  // half of the epilogue warps own live TMEM lanes, so short reductions (K < 512: 1x1 shortcuts, the 2x2 taps of stride-2
  // 64-channel layers), which are paced by the epilogue, and layers with a mask / residual input (B200: 64->64 + residual at
  // 640x384 x 16: 0.59 -> 0.74 ms swapped) stay on the N = 64 kernel.
  if (!halo && d->Cout >= 64 && !(s2e_debug_get(6) & 8) &&
      (d->Cout >= 128 || (d->ntaps * d->Cin >= 512 && !d->relu_mask && !d->residual))) {
    switch (d->act) {
      case S2E_ACT_LRELU: return launch_fwd<128, S2E_ACT_LRELU, 3, false>(d, x, wp, bias, scale, y, stream);
      case S2E_ACT_RELU: return launch_fwd<128, S2E_ACT_RELU, 3, false>(d, x, wp, bias, scale, y, stream);
      default: return launch_fwd<128, S2E_ACT_NONE, 3, false>(d, x, wp, bias, scale, y, stream);
    }
  }
  if (d->Cout >= 128) {
    if (halo) { S2E_FWD_DISPATCH(128, true) }
    S2E_FWD_DISPATCH(128, false)
  }
  if (halo) { S2E_FWD_DISPATCH(64, true) }
  S2E_FWD_DISPATCH(64, false)
#undef S2E_FWD_DISPATCH
}

int s2e_tapconv_wgrad_tc(const s2e_conv_t* d, const void* x, const void* dy, float* dwp, cudaStream_t stream) {
  S2E_REQUIRE(d->in_act == S2E_ACT_NONE, "tapconv_wgrad: in_act exists on the CUDA-core path only");
  S2E_REQUIRE(d->Cin % 8 == 0 && d->Cout % 8 == 0 && d->Cin >= 64 && d->Cout >= 64,
              "tcgen05 wgrad needs Cin,Cout %% 8 == 0 and >= 64 (Cin=%d Cout=%d)", d->Cin, d->Cout);
  // orientation: the N side of the accumulator should be the wide one (N = 256 halves the shared-memory traffic
  // per MMA), and a 64-channel side should not occupy the 128-row M side
  // (debug key 5 = 2: the previous rule, which also swapped Cout < 128 <= Cin so that the 64-channel side left the M side;
  //  with merged taps the unswapped form -- M = 64 live rows, N = 2 taps x 128 -- is the faster one)
  const int swap = (d->Cout > d->Cin) || (s2e_debug_get(5) == 2 && d->Cout < 128 && d->Cin >= 128);
  const int nside = swap ? d->Cout : d->Cin;
  // narrow layers (N side < 256): several taps per CTA against one staged dY tile instead of one CTA per tap re-reading dY
  // nine times through L2 (B200, round 2: 64->64 196 -> 243, 128->64 389 -> 482, 128->128 752 -> 937 TFLOP/s at 640x384 x 16).
  // Debug key 5 = 1 restores the one-tap-per-CTA kernel.
  if (s2e_debug_get(5) != 1 && d->ntaps >= 3 && nside < 256) {
    if (nside >= 128) return launch_wgrad_mt<128, 3>(d, x, dy, dwp, swap, stream);
    return launch_wgrad_mt<64, 3>(d, x, dy, dwp, swap, stream);
  }
  if (nside >= 256) return launch_wgrad<256>(d, x, dy, dwp, swap, stream);
  if (nside >= 128) return launch_wgrad<128>(d, x, dy, dwp, swap, stream);
  return launch_wgrad<64>(d, x, dy, dwp, swap, stream);
}
