// CUDA-core tap convolution (same contract as the tcgen05 kernels in conv_tc.cu).
// Used for the MMA-unfriendly layers (Cin in {1,4,5,20}, Cout = 1: G.fc, mlp_shared, conv_img, D model0/model4,
// E layer0 -- < 1.5 % of the step's FLOPs) and as the on-device cross-check of the tensor-core path.
#include "common.cuh"

namespace {

struct Geom {
  int B, Hi, Wi, Cin, Ho, Wo, Cout, ntaps, act;
  int dy[S2E_MAX_TAPS], dx[S2E_MAX_TAPS];
};

// 64 pixels x (16*TN) couts per block, 256 threads, each thread 4 pixels x TN couts.
template <int TN>
__global__ void __launch_bounds__(256) simt_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ wp,
                                                       const float* __restrict__ bias, const float* __restrict__ scale,
                                                       bf16* __restrict__ y, const Geom g, const bf16* __restrict__ mask,
                                                       float mask_slope) {
  constexpr int BNT = 16 * TN;
  __shared__ float As[16][64 + 1];
  __shared__ float Bs[16][BNT + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long P = (long long)g.B * g.Ho * g.Wo;
  const long long p0 = (long long)blockIdx.x * 64;
  const int n0 = blockIdx.y * BNT;

  // A-load assignment: element idx = tid + i*256 -> pixel (tid>>4) + 16 i, channel tid & 15
  int ab[4], ah[4], aw[4];
  bool av[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long p = p0 + (tid >> 4) + 16 * i;
    av[i] = p < P;
    long long pp = av[i] ? p : 0;
    aw[i] = (int)(pp % g.Wo);
    pp /= g.Wo;
    ah[i] = (int)(pp % g.Ho);
    ab[i] = (int)(pp / g.Ho);
  }
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int t = 0; t < g.ntaps; ++t) {
    const int dy = g.dy[t], dx = g.dx[t];
    for (int c0 = 0; c0 < g.Cin; c0 += 16) {
      const int c = c0 + (tid & 15);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int hi = ah[i] + dy, wi = aw[i] + dx;
        float v = 0.f;
        if (av[i] && c < g.Cin && hi >= 0 && hi < g.Hi && wi >= 0 && wi < g.Wi)
          v = __bfloat162float(x[(((long long)ab[i] * g.Hi + hi) * g.Wi + wi) * g.Cin + c]);
        As[tid & 15][(tid >> 4) + 16 * i] = v;
      }
      for (int idx = tid; idx < BNT * 16; idx += 256) {
        const int n = idx >> 4, cc = c0 + (idx & 15);
        float v = 0.f;
        if (n0 + n < g.Cout && cc < g.Cin) v = __bfloat162float(wp[((long long)t * g.Cout + n0 + n) * g.Cin + cc]);
        Bs[idx & 15][n] = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float a[4], b[TN];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  const float sc = scale ? __ldg(scale) : 1.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long p = p0 + ty + 16 * i;
    if (p >= P) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= g.Cout) continue;
      float v = act_apply(acc[i][j] * sc + (bias ? __ldg(bias + n) : 0.f), g.act);
      if (mask && !(__bfloat162float(mask[p * g.Cout + n]) > 0.f)) v *= mask_slope;
      y[p * g.Cout + n] = __float2bfloat16(v);
    }
  }
}

// dW[t][co][ci] += sum_p dy[p][co] * x[p+tap][ci]; 64 co x 64 ci per block, split over pixel ranges.
__global__ void __launch_bounds__(256) simt_wgrad_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                         float* __restrict__ dwp, const Geom g, int ksplit) {
  __shared__ float As[16][64 + 1];  // [pixel][co]
  __shared__ float Bs[16][64 + 1];  // [pixel][ci]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int mt = (g.Cout + 63) / 64, nt = (g.Cin + 63) / 64;
  int wi = blockIdx.x;
  const int ks = wi % ksplit;
  wi /= ksplit;
  const int n_idx = wi % nt;
  wi /= nt;
  const int m_idx = wi % mt;
  const int t = wi / mt;
  const int m0 = m_idx * 64, n0 = n_idx * 64;
  const long long P = (long long)g.B * g.Ho * g.Wo;
  const long long pb = P * ks / ksplit, pe = P * (ks + 1) / ksplit;
  const int tdy = g.dy[t], tdx = g.dx[t];
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long pc = pb; pc < pe; pc += 16) {
    // each thread loads 4 elements of each tile: pixel = tid >> 4, channels (tid & 15) + 16 i
    const long long p = pc + (tid >> 4);
    const bool pv = p < pe;
    long long pp = pv ? p : 0;
    const int wo = (int)(pp % g.Wo);
    pp /= g.Wo;
    const int ho = (int)(pp % g.Ho);
    const int b = (int)(pp / g.Ho);
    const int hi = ho + tdy, wi2 = wo + tdx;
    const bool xv = pv && hi >= 0 && hi < g.Hi && wi2 >= 0 && wi2 < g.Wi;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = (tid & 15) + 16 * i;
      float a = 0.f, bb = 0.f;
      if (pv && m0 + c < g.Cout) a = __bfloat162float(dy[p * g.Cout + m0 + c]);
      if (xv && n0 + c < g.Cin) bb = __bfloat162float(x[(((long long)b * g.Hi + hi) * g.Wi + wi2) * g.Cin + n0 + c]);
      As[tid >> 4][c] = a;
      Bs[tid >> 4][c] = bb;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = m0 + ty + 16 * i;
    if (co >= g.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = n0 + tx + 16 * j;
      if (ci >= g.Cin) continue;
      atomicAdd(dwp + ((long long)t * g.Cout + co) * g.Cin + ci, acc[i][j]);
    }
  }
}

Geom make_geom(const s2e_conv_t* d) {
  Geom g;
  g.B = d->B;
  g.Hi = d->Hi;
  g.Wi = d->Wi;
  g.Cin = d->Cin;
  g.Ho = d->Ho;
  g.Wo = d->Wo;
  g.Cout = d->Cout;
  g.ntaps = d->ntaps;
  g.act = d->act;
  for (int i = 0; i < d->ntaps; ++i) {
    g.dy[i] = d->tap_dy[i];
    g.dx[i] = d->tap_dx[i];
  }
  return g;
}

}  // namespace

int s2e_thin_fwd(const s2e_conv_t*, const void*, const void*, const float*, const float*, void*, cudaStream_t);
int s2e_thin_wgrad(const s2e_conv_t*, const void*, const void*, float*, cudaStream_t);

int s2e_tapconv_fwd_simt(const s2e_conv_t* d, const void* x, const void* wp, const float* bias, const float* scale,
                         void* y, cudaStream_t stream) {
  S2E_REQUIRE(!d->residual && !d->spade_x && (d->bias_n == 0 || d->bias_n == d->Cout),
              "tapconv_fwd: fused residual / SPADE epilogue / channel-padded bias exist on the tcgen05 path only");
  if (!s2e_debug_get(2)) {  // debug key 2 = keep thin layers on the generic kernel
    const int rc = s2e_thin_fwd(d, x, wp, bias, scale, y, stream);
    if (rc != 0) return rc < 0 ? rc : S2E_OK;
  }
  S2E_REQUIRE(d->in_act == S2E_ACT_NONE && !d->img_out, "tapconv_fwd: in_act / the tanh image head are implemented by the thin-layer kernels only");
  Geom g = make_geom(d);
  const long long P = (long long)d->B * d->Ho * d->Wo;
  if (P == 0) return S2E_OK;
  if (d->Cout > 16) {
    dim3 grid((unsigned)ceil_div_ll(P, 64), (unsigned)ceil_div(d->Cout, 64));
    simt_fwd_kernel<4><<<grid, 256, 0, stream>>>((const bf16*)x, (const bf16*)wp, bias, scale, (bf16*)y, g, (const bf16*)d->relu_mask, d->mask_slope);
  } else {
    dim3 grid((unsigned)ceil_div_ll(P, 64), (unsigned)ceil_div(d->Cout, 16));
    simt_fwd_kernel<1><<<grid, 256, 0, stream>>>((const bf16*)x, (const bf16*)wp, bias, scale, (bf16*)y, g, (const bf16*)d->relu_mask, d->mask_slope);
  }
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

int s2e_tapconv_wgrad_simt(const s2e_conv_t* d, const void* x, const void* dy, float* dwp, cudaStream_t stream) {
  if (!s2e_debug_get(2)) {
    const int rc = s2e_thin_wgrad(d, x, dy, dwp, stream);
    if (rc != 0) return rc < 0 ? rc : S2E_OK;
  }
  S2E_REQUIRE(d->in_act == S2E_ACT_NONE, "tapconv_wgrad: in_act is implemented by the thin-layer kernels only");
  Geom g = make_geom(d);
  const long long P = (long long)d->B * d->Ho * d->Wo;
  if (P == 0) return S2E_OK;
  const int base = d->ntaps * ceil_div(d->Cout, 64) * ceil_div(d->Cin, 64);
  int ksplit = ceil_div(4 * s2e_num_sms(), base);
  long long max_split = P / 256;
  if (max_split < 1) max_split = 1;
  if (ksplit > max_split) ksplit = (int)max_split;
  if (ksplit < 1) ksplit = 1;
  simt_wgrad_kernel<<<base * ksplit, 256, 0, stream>>>((const bf16*)x, (const bf16*)dy, dwp, g, ksplit);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}
