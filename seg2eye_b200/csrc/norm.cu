// SPADE+Style normalisation / modulation and InstanceNorm kernels (HBM-bound; NHWC bf16, fp32/fp64 statistics).
// One 16-byte vector (8 channels) per thread access, fully coalesced along the channel axis; per-channel
// partial sums live in registers, are combined across the block through shared memory and leave the block as one
// fp64 atomic per channel (so E[x^2]-E[x]^2 is formed without cancellation trouble).
#include "common.cuh"

namespace {

constexpr int NT = 256;

// ---------------------------------------------------------------- per-channel sum / sum of squares
// grid = (chunks, G) ; G = B when per_sample else 1 (then the chunk range spans all samples)
// body(p, c, acc) consumes one pixel; body4(p, stride, c, acc) consumes pixels p, p+stride, p+2 stride, p+3 stride with all
// of their loads issued before the first use (memory-level parallelism).
template <int NACC, typename F, typename F4>
__device__ __forceinline__ void block_channel_reduce(int C, long long pix_begin, long long pix_end, double* out /*[NACC][C]*/,
                                                     F&& body, F4&& body4) {
  extern __shared__ float red[];  // [NT][NACC*8] worst case handled by looping
  const int cg = C >> 3;          // channel groups of 8
  const int tid = threadIdx.x;
  // thread -> (channel group, pixel lane); when cg > NT loop over channel-group passes
  for (int cg0 = 0; cg0 < cg; cg0 += NT) {
    const int ncg = min(NT, cg - cg0);
    const int lanes = NT / ncg;
    const int my_cg = tid % ncg, my_lane = tid / ncg;
    float acc[NACC][8];
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[a][j] = 0.f;
    if (my_lane < lanes) {
      long long p = pix_begin + my_lane;
      for (; p + 3LL * lanes < pix_end; p += 4LL * lanes) body4(p, (long long)lanes, (cg0 + my_cg) * 8, acc);
      for (; p < pix_end; p += lanes) body(p, (cg0 + my_cg) * 8, acc);
    }
    // combine pixel lanes through smem: layout [lane][ncg][NACC*8]
    for (int a = 0; a < NACC; ++a) {
      __syncthreads();
      if (my_lane < lanes) {
#pragma unroll
        for (int j = 0; j < 8; ++j) red[(my_lane * ncg + my_cg) * 8 + j] = acc[a][j];
      }
      __syncthreads();
      for (int idx = tid; idx < ncg * 8; idx += NT) {
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += red[l * ncg * 8 + idx];
        atomicAdd(out + (size_t)a * C + cg0 * 8 + idx, (double)s);
      }
    }
  }
}

__global__ void __launch_bounds__(NT) stats_kernel(const bf16* __restrict__ x, int HW, int C, long long pix_total,
                                                   int per_sample, double* __restrict__ acc) {
  const int g = blockIdx.y;
  const long long span = per_sample ? HW : pix_total;
  const long long base = per_sample ? (long long)g * HW : 0;
  const long long chunk = (span + gridDim.x - 1) / gridDim.x;
  const long long b0 = base + (long long)blockIdx.x * chunk;
  const long long b1 = min(base + span, b0 + chunk);
  auto one = [&](long long p, int c, float(*a)[8]) {
    float f[8];
    unpack8(ld_stream8(x + p * C + c), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[0][j] += f[j];
      a[1][j] = fmaf(f[j], f[j], a[1][j]);
    }
  };
  auto four = [&](long long p, long long st, int c, float(*a)[8]) {
    bf16x8 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = ld_stream8(x + (p + i * st) * C + c);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[0][j] += f[j];
        a[1][j] = fmaf(f[j], f[j], a[1][j]);
      }
    }
  };
  block_channel_reduce<2>(C, b0, b1, acc + (size_t)g * 2 * C, one, four);
}

// in_scale (optional, one value per `group` consecutive statistic groups): statistics are those of x * in_scale without
// the product ever being formed: mean stays that of x, rstd' = s / sqrt(var * s^2 + eps), so (x - mean) * rstd' == norm(s x).
__global__ void finalize_kernel(const double* __restrict__ acc, int G, int C, double count, float eps, float* mean,
                                float* rstd, float* running_mean, float* running_var, float momentum,
                                long long* nbt, const float* __restrict__ in_scale, int group, double count_unbiased) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && nbt) *nbt += 1;
  if (i >= G * C) return;
  const int g = i / C, c = i % C;
  const double s = acc[(size_t)g * 2 * C + c], ss = acc[(size_t)g * 2 * C + C + c];
  const double m = s / count;
  double var = ss / count - m * m;
  if (var < 0) var = 0;
  mean[i] = (float)m;
  const double sc = in_scale ? (double)in_scale[g / group] : 1.0;
  rstd[i] = (float)(sc / sqrt(var * sc * sc + (double)eps));
  if (running_mean && g == 0) {
    const double cu = count_unbiased > 0 ? count_unbiased : count;   // element count BatchNorm sees (x4 when the statistics
    const double unb = cu > 1 ? var * cu / (cu - 1) : var;           // come from the source of a nearest-2x up-sampling)
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

// Pixel index of x for output pixel q of sample b.  up_w == 0: x has the output's resolution.  up_w = W of the output:
// x is the (H/2, W/2) tensor the block input was nearest-2x up-sampled from (generator.py:50,86-93); the up-sampled copy
// is never materialised, the kernels read the source through this index map.
__device__ __forceinline__ long long x_pixel(int b, long long q, int HW, int up_w) {
  if (!up_w) return (long long)b * HW + q;
  const unsigned qq = (unsigned)q, uw = (unsigned)up_w;   // q < HW < 2^31: 32-bit division
  const unsigned h = qq / uw, w = qq - h * uw;
  return (long long)b * (HW >> 2) + (long long)((h >> 1) * (uw >> 1) + (w >> 1));
}

// bit j = (o[j] > 0): the activation mask of 8 channels in one byte
__device__ __forceinline__ uint8_t sign_bits8(const float* o) {
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) m |= (o[j] > 0.f ? 1u : 0u) << j;
  return (uint8_t)m;
}

// constants of the fused gamma|beta-convolution epilogue (conv_tc.cu): par[b][0..3][c]
__global__ void spade_params_kernel(const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ style,
                                    int B, int C, int per_sample, float* __restrict__ par) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i % C;
  const int so = per_sample ? b * C : 0;
  const float rs = rstd[so + c];
  float* o = par + (size_t)b * 4 * C + c;
  o[0] = rs;
  o[C] = -mean[so + c] * rs;
  o[2 * C] = style ? 1.f + style[(size_t)b * 2 * C + c] : 0.f;   // style == NULL: plain SPADE, no style term
  o[3 * C] = style ? style[(size_t)b * 2 * C + C + c] : 0.f;
}

// ---------------------------------------------------------------- SPADE+Style forward (elementwise, 8 B/elem)
// grid = (pixel chunks, B).  A thread owns one 8-channel group: its per-channel constants (mean, rstd, style) are
// loaded once into registers, then it streams pixels: 3 x 16-byte loads + 1 x 16-byte store per pixel, two pixels in
// flight per iteration.
__global__ void __launch_bounds__(NT) spade_style_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ gb,
                                                             const float* __restrict__ style, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, int HW, int C, int per_sample,
                                                             int act, bf16* __restrict__ out, uint8_t* __restrict__ amask, int up_w) {
  // amask (optional): one bit per element, set where out > 0 -- all the backward pass needs of `out` (16x fewer bytes)
  const int b = blockIdx.y;
  const long long chunk = ((long long)HW + gridDim.x - 1) / gridDim.x;
  const long long q0 = (long long)blockIdx.x * chunk;
  const long long q1 = min((long long)HW, q0 + chunk);
  const int cg = C >> 3;
  const float* mu_b = mean + (per_sample ? b * C : 0);
  const float* rs_b = rstd + (per_sample ? b * C : 0);
  // style == NULL: plain SPADE (normalization.py:91-105 alone): out = act((x*rs - mu*rs)*(1+g) + beta), no style term, no 1/2
  const float* st_b = style ? style + (size_t)b * 2 * C : nullptr;
  const float os = style ? 0.5f : 1.0f;
  for (int cg0 = 0; cg0 < cg; cg0 += NT) {
    const int ncg = min(NT, cg - cg0);
    const int lanes = NT / ncg;
    const int my_cg = threadIdx.x % ncg, my_lane = threadIdx.x / ncg;
    if (my_lane >= lanes) continue;
    const int c = (cg0 + my_cg) * 8;
    float k_a[8], k_b[8], k_c[8];  // out = os*( (x*rs - mu*rs)*(1+g) + beta + x*(1+s0) + s1 )
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float rs = rs_b[c + j];
      k_a[j] = rs;
      k_b[j] = -mu_b[c + j] * rs;
      k_c[j] = st_b ? 1.f + st_b[c + j] : 0.f;
    }
    float s1v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1v[j] = st_b ? st_b[C + c + j] : 0.f;
    const long long base = (long long)b * HW;
    long long q = q0 + my_lane;
    for (; q + lanes < q1; q += 2 * lanes) {
      const long long pA = base + q, pB = base + q + lanes;
      const bf16x8 xa = ld_stream8(x + x_pixel(b, q, HW, up_w) * C + c), xb = ld_stream8(x + x_pixel(b, q + lanes, HW, up_w) * C + c);
      const bf16x8 ga = ld_stream8(gb + pA * 2 * C + c), gbb = ld_stream8(gb + pB * 2 * C + c);
      const bf16x8 ba = ld_stream8(gb + pA * 2 * C + C + c), bb = ld_stream8(gb + pB * 2 * C + C + c);
      float xf[8], gf[8], bf_[8], o[8];
      unpack8(xa, xf); unpack8(ga, gf); unpack8(ba, bf_);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        o[j] = act_apply(os * (fmaf(fmaf(xf[j], k_a[j], k_b[j]), 1.f + gf[j], bf_[j]) + fmaf(xf[j], k_c[j], s1v[j])), act);
      st_stream8(out + pA * C + c, pack8(o));
      if (amask) amask[pA * cg + cg0 + my_cg] = sign_bits8(o);
      unpack8(xb, xf); unpack8(gbb, gf); unpack8(bb, bf_);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        o[j] = act_apply(os * (fmaf(fmaf(xf[j], k_a[j], k_b[j]), 1.f + gf[j], bf_[j]) + fmaf(xf[j], k_c[j], s1v[j])), act);
      st_stream8(out + pB * C + c, pack8(o));
      if (amask) amask[pB * cg + cg0 + my_cg] = sign_bits8(o);
    }
    for (; q < q1; q += lanes) {
      const long long pA = base + q;
      float xf[8], gf[8], bf_[8], o[8];
      unpack8(ld_stream8(x + x_pixel(b, q, HW, up_w) * C + c), xf);
      unpack8(ld_stream8(gb + pA * 2 * C + c), gf);
      unpack8(ld_stream8(gb + pA * 2 * C + C + c), bf_);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        o[j] = act_apply(os * (fmaf(fmaf(xf[j], k_a[j], k_b[j]), 1.f + gf[j], bf_[j]) + fmaf(xf[j], k_c[j], s1v[j])), act);
      st_stream8(out + pA * C + c, pack8(o));
      if (amask) amask[pA * cg + cg0 + my_cg] = sign_bits8(o);
    }
  }
}

// ---------------------------------------------------------------- SPADE+Style backward
// pass 1: per (sample, channel) sums  S1 = sum dxh, S2 = sum dxh*xh, S3 = sum g*x, S4 = sum g, S5 = sum g*xh
// (S4 / S5 are also the per-channel sums of dbeta / dgamma, i.e. the bias gradients of the gamma|beta convolution)
__global__ void __launch_bounds__(NT) spade_style_bwd_reduce_kernel(const bf16* __restrict__ dout, const uint8_t* __restrict__ amask,
                                                                    const bf16* __restrict__ x, const bf16* __restrict__ gb,
                                                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                    int HW, int C, int per_sample, int act, double* __restrict__ racc,
                                                                    int up_w, int gstride, float os) {
  const int b = blockIdx.y;
  const long long chunk = ((long long)HW + gridDim.x - 1) / gridDim.x;
  const long long b0 = (long long)b * HW + (long long)blockIdx.x * chunk;
  const long long b1 = min((long long)(b + 1) * HW, b0 + chunk);
  const float* mu = mean + (per_sample ? b * C : 0);
  const float* rs = rstd + (per_sample ? b * C : 0);
  auto one = [&](long long p, int c, float(*a)[8]) {
    float df[8], xf[8], gf[8];
    unpack8(ld_stream8(dout + p * C + c), df);
    unpack8(ld_stream8(x + x_pixel(b, p - (long long)b * HW, HW, up_w) * C + c), xf);
    unpack8(ld_stream8(gb + p * gstride + c), gf);
    const uint32_t mbits = act != S2E_ACT_NONE ? (uint32_t)amask[p * (C >> 3) + (c >> 3)] : 0xffu;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float g = os * df[j];
      if (act == S2E_ACT_LRELU) g *= ((mbits >> j) & 1u) ? 1.f : 0.2f;
      if (act == S2E_ACT_RELU) g *= ((mbits >> j) & 1u) ? 1.f : 0.f;
      const float xh = (xf[j] - __ldg(mu + c + j)) * __ldg(rs + c + j);
      const float dxh = g * (1.f + gf[j]);
      a[0][j] += dxh;
      a[1][j] = fmaf(dxh, xh, a[1][j]);
      a[2][j] = fmaf(g, xf[j], a[2][j]);
      a[3][j] += g;
      a[4][j] = fmaf(g, xh, a[4][j]);
    }
  };
  auto four = [&](long long p, long long st, int c, float(*a)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) one(p + i * st, c, a);   // the compiler hoists the independent streaming loads
  };
  block_channel_reduce<5>(C, b0, b1, racc + (size_t)b * 5 * C, one, four);
}

// fold the per-sample sums: m1/m2 per stat group (float [G][2][C]), dstyle [B][2C] and, optionally, chsum [3][C] =
// per-channel sums over the whole batch of dgamma, dbeta and dx.  The last one needs no extra pass: the normalisation
// part of dx sums to zero over the pixels its statistics were taken from, so sum dx = sum_b (1 + s0[b]) * S4[b].
__global__ void spade_style_bwd_fold_kernel(const double* __restrict__ racc, int B, int C, int per_sample, double count,
                                            const float* __restrict__ style, float* __restrict__ m12,
                                            float* __restrict__ dstyle, float* __restrict__ chsum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i % C;
  const double* r = racc + (size_t)b * 5 * C;
  if (dstyle) {
    dstyle[(size_t)b * 2 * C + c] = (float)r[2 * C + c];
    dstyle[(size_t)b * 2 * C + C + c] = (float)r[3 * C + c];
  }
  if (per_sample) {
    m12[(size_t)b * 2 * C + c] = (float)(r[c] / count);
    m12[(size_t)b * 2 * C + C + c] = (float)(r[C + c] / count);
  }
  if (b == 0 && (!per_sample || chsum)) {
    double s1 = 0, s2 = 0, sg = 0, sb = 0, sx = 0;
    for (int bb = 0; bb < B; ++bb) {
      const double* q = racc + (size_t)bb * 5 * C;
      s1 += q[c];
      s2 += q[C + c];
      sb += q[3 * C + c];
      sg += q[4 * C + c];
      if (style) sx += (1.0 + (double)style[(size_t)bb * 2 * C + c]) * q[3 * C + c];
    }
    if (!per_sample) {
      m12[c] = (float)(s1 / count);
      m12[C + c] = (float)(s2 / count);
    }
    if (chsum) {
      chsum[c] = (float)sg;
      chsum[C + c] = (float)sb;
      chsum[2 * C + c] = (float)sx;
    }
  }
}

// pass 2: dx = rstd*(dxh - m1 - xh*m2) + g*(1+s0) ; dgamma = g*xh ; dbeta = g     (same thread mapping as the forward)
__global__ void __launch_bounds__(NT) spade_style_bwd_apply_kernel(const bf16* __restrict__ dout, const uint8_t* __restrict__ amask,
                                                                   const bf16* __restrict__ x, const bf16* __restrict__ gb,
                                                                   const float* __restrict__ style, const float* __restrict__ mean,
                                                                   const float* __restrict__ rstd, const float* __restrict__ m12,
                                                                   int HW, int C, int per_sample, int act,
                                                                   bf16* __restrict__ dx, int dx_acc, bf16* __restrict__ dgb, int up_w,
                                                                   int gstride) {
  const int b = blockIdx.y;
  const long long chunk = ((long long)HW + gridDim.x - 1) / gridDim.x;
  const long long q0 = (long long)blockIdx.x * chunk;
  const long long q1 = min((long long)HW, q0 + chunk);
  const int cg = C >> 3;
  const int so = per_sample ? b * C : 0;
  const float* m1p = m12 + (per_sample ? (size_t)b * 2 * C : 0);
  const float* s0p = style ? style + (size_t)b * 2 * C : nullptr;
  const float os = style ? 0.5f : 1.0f;
  for (int cg0 = 0; cg0 < cg; cg0 += NT) {
    const int ncg = min(NT, cg - cg0);
    const int lanes = NT / ncg;
    const int my_cg = threadIdx.x % ncg, my_lane = threadIdx.x / ncg;
    if (my_lane >= lanes) continue;
    const int c = (cg0 + my_cg) * 8;
    float rsd[8], mu[8], m1[8], m2[8], s0[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      rsd[j] = rstd[so + c + j];
      mu[j] = mean[so + c + j];
      m1[j] = m1p[c + j];
      m2[j] = m1p[C + c + j];
      s0[j] = s0p ? 1.f + s0p[c + j] : 0.f;
    }
    // two pixels per iteration: all six 16-byte loads are issued before the first use (memory-level parallelism)
    auto finish = [&](long long p, const bf16x8& vd, const bf16x8& vx, const bf16x8& vg, uint32_t mbits) {
      float df[8], xf[8], gf[8], odx[8], odg[8], odb[8], prev[8];
      if (dx_acc) unpack8(*reinterpret_cast<const bf16x8*>(dx + p * C + c), prev);
      unpack8(vd, df); unpack8(vx, xf); unpack8(vg, gf);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float g = os * df[j];
        if (act == S2E_ACT_LRELU) g *= ((mbits >> j) & 1u) ? 1.f : 0.2f;
        if (act == S2E_ACT_RELU) g *= ((mbits >> j) & 1u) ? 1.f : 0.f;
        const float xh = (xf[j] - mu[j]) * rsd[j];
        const float dxh = g * (1.f + gf[j]);
        float d = rsd[j] * (dxh - m1[j] - xh * m2[j]) + g * s0[j];
        if (dx_acc) d += prev[j];
        odx[j] = d;
        odg[j] = g * xh;
        odb[j] = g;
      }
      *reinterpret_cast<bf16x8*>(dx + p * C + c) = pack8(odx);
      st_stream8(dgb + p * 2 * C + c, pack8(odg));
      st_stream8(dgb + p * 2 * C + C + c, pack8(odb));
    };
    const long long base = (long long)b * HW;
    const int mcol = cg0 + my_cg;
    long long q = q0 + my_lane;
    for (; q + lanes < q1; q += 2 * lanes) {
      const long long pA = base + q, pB = base + q + lanes;
      const bf16x8 da = ld_stream8(dout + pA * C + c), db = ld_stream8(dout + pB * C + c);
      const bf16x8 xa = ld_stream8(x + x_pixel(b, q, HW, up_w) * C + c), xb = ld_stream8(x + x_pixel(b, q + lanes, HW, up_w) * C + c);
      const bf16x8 ga = ld_stream8(gb + pA * gstride + c), gb2 = ld_stream8(gb + pB * gstride + c);
      const uint32_t ma = act != S2E_ACT_NONE ? (uint32_t)amask[pA * cg + mcol] : 0xffu;
      const uint32_t mb = act != S2E_ACT_NONE ? (uint32_t)amask[pB * cg + mcol] : 0xffu;
      finish(pA, da, xa, ga, ma);
      finish(pB, db, xb, gb2, mb);
    }
    for (; q < q1; q += lanes) {
      const long long p = base + q;
      const uint32_t mbits = act != S2E_ACT_NONE ? (uint32_t)amask[p * cg + mcol] : 0xffu;
      finish(p, ld_stream8(dout + p * C + c), ld_stream8(x + x_pixel(b, q, HW, up_w) * C + c), ld_stream8(gb + p * gstride + c), mbits);
    }
  }
}

// ---------------------------------------------------------------- InstanceNorm (+act)
// same streaming structure as the SPADE kernels: grid (pixel chunks, B), per-channel constants in registers
__global__ void __launch_bounds__(NT) instnorm_apply_kernel(const bf16* __restrict__ x, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, int HW, int C, int act,
                                                            bf16* __restrict__ y) {
  const int b = blockIdx.y;
  const long long chunk = ((long long)HW + gridDim.x - 1) / gridDim.x;
  const long long q0 = (long long)blockIdx.x * chunk;
  const long long q1 = min((long long)HW, q0 + chunk);
  const int cg = C >> 3;
  for (int cg0 = 0; cg0 < cg; cg0 += NT) {
    const int ncg = min(NT, cg - cg0);
    const int lanes = NT / ncg;
    const int my_cg = threadIdx.x % ncg, my_lane = threadIdx.x / ncg;
    if (my_lane >= lanes) continue;
    const int c = (cg0 + my_cg) * 8;
    float ka[8], kb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ka[j] = rstd[(size_t)b * C + c + j];
      kb[j] = -mean[(size_t)b * C + c + j] * ka[j];
    }
    const long long base = (long long)b * HW;
    long long q = q0 + my_lane;
    for (; q + lanes < q1; q += 2 * lanes) {
      const bf16x8 va = ld_stream8(x + (base + q) * C + c), vb = ld_stream8(x + (base + q + lanes) * C + c);
      float f[8], o[8];
      unpack8(va, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = act_apply(fmaf(f[j], ka[j], kb[j]), act);
      st_stream8(y + (base + q) * C + c, pack8(o));
      unpack8(vb, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = act_apply(fmaf(f[j], ka[j], kb[j]), act);
      st_stream8(y + (base + q + lanes) * C + c, pack8(o));
    }
    for (; q < q1; q += lanes) {
      float f[8], o[8];
      unpack8(ld_stream8(x + (base + q) * C + c), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = act_apply(fmaf(f[j], ka[j], kb[j]), act);
      st_stream8(y + (base + q) * C + c, pack8(o));
    }
  }
}

// reduce: S1 = sum g, S2 = sum g*xh  (g = dy * act'(y))
__global__ void __launch_bounds__(NT) instnorm_bwd_reduce_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y,
                                                                 const bf16* __restrict__ x, const float* __restrict__ mean,
                                                                 const float* __restrict__ rstd, int HW, int C, int act,
                                                                 double* __restrict__ racc) {
  const int b = blockIdx.y;
  const long long chunk = ((long long)HW + gridDim.x - 1) / gridDim.x;
  const long long b0 = (long long)b * HW + (long long)blockIdx.x * chunk;
  const long long b1 = min((long long)(b + 1) * HW, b0 + chunk);
  auto one = [&](long long p, int c, float(*a)[8]) {
    float df[8], yf[8], xf[8];
    unpack8(ld_stream8(dy + p * C + c), df);
    unpack8(ld_stream8(x + p * C + c), xf);
    if (act != S2E_ACT_NONE) unpack8(ld_stream8(y + p * C + c), yf);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float g = df[j];
      if (act == S2E_ACT_LRELU) g *= (yf[j] > 0.f ? 1.f : 0.2f);
      const float xh = (xf[j] - __ldg(mean + b * C + c + j)) * __ldg(rstd + b * C + c + j);
      a[0][j] += g;
      a[1][j] = fmaf(g, xh, a[1][j]);
    }
  };
  auto four = [&](long long p, long long st, int c, float(*a)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) one(p + i * st, c, a);
  };
  block_channel_reduce<2>(C, b0, b1, racc + (size_t)b * 2 * C, one, four);
}

__global__ void __launch_bounds__(NT) instnorm_bwd_apply_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y,
                                                                const bf16* __restrict__ x, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, const double* __restrict__ racc,
                                                                int HW, int C, int act, bf16* __restrict__ dx) {
  const int b = blockIdx.y;
  const long long chunk = ((long long)HW + gridDim.x - 1) / gridDim.x;
  const long long q0 = (long long)blockIdx.x * chunk;
  const long long q1 = min((long long)HW, q0 + chunk);
  const int cg = C >> 3;
  const float inv = 1.f / (float)HW;
  for (int cg0 = 0; cg0 < cg; cg0 += NT) {
    const int ncg = min(NT, cg - cg0);
    const int lanes = NT / ncg;
    const int my_cg = threadIdx.x % ncg, my_lane = threadIdx.x / ncg;
    if (my_lane >= lanes) continue;
    const int c = (cg0 + my_cg) * 8;
    float rsd[8], mu[8], m1[8], m2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      rsd[j] = rstd[(size_t)b * C + c + j];
      mu[j] = mean[(size_t)b * C + c + j];
      m1[j] = (float)racc[(size_t)b * 2 * C + c + j] * inv;
      m2[j] = (float)racc[(size_t)b * 2 * C + C + c + j] * inv;
    }
    const long long base = (long long)b * HW;
    for (long long q = q0 + my_lane; q < q1; q += lanes) {
      const long long p = base + q;
      float df[8], yf[8], xf[8], o[8];
      const bf16x8 vd = ld_stream8(dy + p * C + c), vx = ld_stream8(x + p * C + c);
      if (act != S2E_ACT_NONE) unpack8(ld_stream8(y + p * C + c), yf);
      unpack8(vd, df);
      unpack8(vx, xf);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float g = df[j];
        if (act == S2E_ACT_LRELU) g *= (yf[j] > 0.f ? 1.f : 0.2f);
        const float xh = (xf[j] - mu[j]) * rsd[j];
        o[j] = rsd[j] * (g - m1[j] - xh * m2[j]);
      }
      st_stream8(dx + p * C + c, pack8(o));
    }
  }
}

// pixel chunks per sample for the streaming elementwise kernels: ~16 resident blocks per SM in total, and at least
// 8 pixels per pixel-lane so the per-thread constant setup is amortised
int ew_chunks(int HW, int B, int C) {
  const int ncg = (C >> 3) < NT ? (C >> 3) : NT;
  const int lanes = NT / ncg;
  long long want = ((long long)s2e_num_sms() * 16 + B - 1) / B;
  long long maxc = ((long long)HW + 8LL * lanes - 1) / (8LL * lanes);
  if (want > maxc) want = maxc;
  return (int)(want < 1 ? 1 : want);
}
int red_chunks(long long pixels, int G) {
  // aim for ~4 blocks per SM overall, at least 64 pixels per block
  long long want = ((long long)s2e_num_sms() * 4 + G - 1) / G;
  long long maxc = (pixels + 63) / 64;
  if (want > maxc) want = maxc;
  return (int)(want < 1 ? 1 : want);
}
size_t red_smem(int C) {
  int ncg = (C >> 3) < NT ? (C >> 3) : NT;
  int lanes = NT / ncg;
  return (size_t)lanes * ncg * 8 * sizeof(float);
}

}  // namespace

extern "C" {

int s2e_norm_stats(const void* x, int B, int HW, int C, int per_sample, double* acc, void* stream) {
  S2E_REQUIRE(C % 8 == 0 && C >= 8, "norm_stats needs C %% 8 == 0 (C=%d)", C);
  cudaStream_t st = (cudaStream_t)stream;
  const int G = per_sample ? B : 1;
  S2E_CHECK_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * G * 2 * C, st));
  const long long span = per_sample ? HW : (long long)B * HW;
  dim3 grid(red_chunks(span, G), G);
  stats_kernel<<<grid, NT, red_smem(C), st>>>((const bf16*)x, HW, C, (long long)B * HW, per_sample, acc);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

int s2e_norm_finalize(const double* acc, int G, int C, double count, double count_unbiased, float eps, float* mean,
                      float* rstd, float* running_mean, float* running_var, float momentum, int64_t* nbt, void* stream) {
  finalize_kernel<<<ceil_div(G * C, 256), 256, 0, (cudaStream_t)stream>>>(acc, G, C, count, eps, mean, rstd, running_mean,
                                                                           running_var, momentum, (long long*)nbt, nullptr, 1,
                                                                           count_unbiased);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

// c[b] = (eps / is_b) * sum_{n in sample b, c} S2[n][c] * rstd'[n][c]^2   (one block per sample)
__global__ void sn_corr_coef_kernel(const double* __restrict__ racc, const float* __restrict__ rstd, const float* __restrict__ inv_sigma,
                                    int C, int group, float eps, float* __restrict__ coef) {
  const int b = blockIdx.x;
  float acc = 0.f;
  for (int i = threadIdx.x; i < group * C; i += blockDim.x) {
    const int n = b * group + i / C, c = i % C;
    const float r = rstd[(size_t)n * C + c];
    acc += (float)racc[(size_t)n * 2 * C + C + c] * r * r;
  }
  __shared__ float sh[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) coef[b] = eps / inv_sigma[b] * v;
  }
}
// dW[co][k] = - sum_b coef[b] * U[b][co] * V[b][k]
__global__ void sn_corr_apply_kernel(const float* __restrict__ coef, const float* __restrict__ U, const float* __restrict__ V, int Bn,
                                     int Cout, int K, float* __restrict__ dw) {
  const long long n = (long long)Cout * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i / K), k = (int)(i % K);
    float acc = 0.f;
    for (int b = 0; b < Bn; ++b) acc = fmaf(coef[b] * U[(size_t)b * Cout + co], V[(size_t)b * K + k], acc);
    dw[i] = -acc;
  }
}

int s2e_spade_params(const float* mean, const float* rstd, const float* style, int B, int C, int per_sample, float* par,
                     void* stream) {
  if (B * C == 0) return S2E_OK;
  spade_params_kernel<<<ceil_div(B * C, 256), 256, 0, (cudaStream_t)stream>>>(mean, rstd, style, B, C, per_sample, par);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

int s2e_spade_style_fwd(const void* x, const void* gb, const float* style, const float* mean, const float* rstd, int B,
                        int HW, int C, int per_sample, int act, void* out, uint8_t* act_mask, int up_w, void* stream) {
  S2E_REQUIRE(C % 8 == 0, "spade_style_fwd needs C %% 8 == 0 (C=%d)", C);
  S2E_REQUIRE(up_w == 0 || (up_w % 2 == 0 && HW % up_w == 0 && (HW / up_w) % 2 == 0), "spade_style_fwd: bad up-sampled width %d", up_w);
  if ((long long)B * HW == 0) return S2E_OK;
  dim3 grid(ew_chunks(HW, B, C), B);
  spade_style_fwd_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)gb, style, mean, rstd, HW, C,
                                                                 per_sample, act, (bf16*)out, act_mask, up_w);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

int s2e_spade_style_bwd(const void* dout, const uint8_t* act_mask, const void* x, const void* gb, const float* style,
                        const float* mean, const float* rstd, int B, int HW, int C, int per_sample, int act, double* racc,
                        void* dx, int dx_accumulate, void* dgb, float* dstyle, float* chsum, int up_w, int gb_stride,
                        void* stream) {
  S2E_REQUIRE(C % 8 == 0, "spade_style_bwd needs C %% 8 == 0 (C=%d)", C);
  S2E_REQUIRE(up_w == 0 || (up_w % 2 == 0 && HW % up_w == 0 && (HW / up_w) % 2 == 0), "spade_style_bwd: bad up-sampled width %d", up_w);
  cudaStream_t st = (cudaStream_t)stream;
  // racc: double [B][5][C] followed by float m12 [B][2][C]
  S2E_REQUIRE(act == S2E_ACT_NONE || act_mask, "spade_style_bwd: the activation mask written by the forward pass is required");
  const int gstride = gb_stride > 0 ? gb_stride : 2 * C;
  S2E_REQUIRE(gstride >= C && gstride % 8 == 0, "spade_style_bwd: bad gamma stride %d", gstride);
  S2E_REQUIRE(style || !dstyle, "spade_style_bwd: plain SPADE (style == NULL) has no style gradient");
  S2E_CHECK_CUDA(cudaMemsetAsync(racc, 0, sizeof(double) * B * 5 * C, st));
  float* m12 = (float*)(racc + (size_t)B * 5 * C);
  dim3 grid(red_chunks(HW, B), B);
  spade_style_bwd_reduce_kernel<<<grid, NT, red_smem(C), st>>>((const bf16*)dout, act_mask, (const bf16*)x,
                                                                (const bf16*)gb, mean, rstd, HW, C, per_sample, act, racc, up_w, gstride,
                                                                style ? 0.5f : 1.0f);
  S2E_LAUNCH_CHECK();
  const double count = per_sample ? (double)HW : (double)B * HW;
  spade_style_bwd_fold_kernel<<<ceil_div(B * C, 256), 256, 0, st>>>(racc, B, C, per_sample, count, style, m12, dstyle, chsum);
  S2E_LAUNCH_CHECK();
  dim3 grid2(ew_chunks(HW, B, C), B);
  spade_style_bwd_apply_kernel<<<grid2, NT, 0, st>>>((const bf16*)dout, act_mask, (const bf16*)x, (const bf16*)gb, style,
                                                     mean, rstd, m12, HW, C, per_sample, act, (bf16*)dx, dx_accumulate,
                                                     (bf16*)dgb, up_w, gstride);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

int s2e_sn_in_correction(const double* racc, const float* rstd, const float* inv_sigma, int Bn, int group, int C, float eps,
                         const float* U, const float* V, int K, float* coef, float* dw, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  sn_corr_coef_kernel<<<Bn, 256, 0, st>>>(racc, rstd, inv_sigma, C, group, eps, coef);
  S2E_LAUNCH_CHECK();
  const long long n = (long long)C * K;
  long long g = (n + 255) / 256;
  if (g > (long long)s2e_num_sms() * 16) g = (long long)s2e_num_sms() * 16;
  sn_corr_apply_kernel<<<(unsigned)g, 256, 0, st>>>(coef, U, V, Bn, C, K, dw);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

int s2e_instnorm_fwd(const void* x, int B, int HW, int C, int act, float eps, const float* in_scale, int group, double* acc,
                     float* mean, float* rstd, void* y, void* stream) {
  int rc = s2e_norm_stats(x, B, HW, C, 1, acc, stream);
  if (rc) return rc;
  finalize_kernel<<<ceil_div(B * C, 256), 256, 0, (cudaStream_t)stream>>>(acc, B, C, (double)HW, eps, mean, rstd, nullptr, nullptr,
                                                                           0.f, nullptr, in_scale, group > 0 ? group : 1, 0.0);
  S2E_LAUNCH_CHECK();
  dim3 grid(ew_chunks(HW, B, C), B);
  instnorm_apply_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>((const bf16*)x, mean, rstd, HW, C, act, (bf16*)y);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

int s2e_instnorm_bwd(const void* dy, const void* y, const void* x, const float* mean, const float* rstd, int B, int HW,
                     int C, int act, double* racc, void* dx, void* stream) {
  S2E_REQUIRE(C % 8 == 0, "instnorm_bwd needs C %% 8 == 0 (C=%d)", C);
  cudaStream_t st = (cudaStream_t)stream;
  S2E_CHECK_CUDA(cudaMemsetAsync(racc, 0, sizeof(double) * B * 2 * C, st));
  dim3 grid(red_chunks(HW, B), B);
  instnorm_bwd_reduce_kernel<<<grid, NT, red_smem(C), st>>>((const bf16*)dy, (const bf16*)y, (const bf16*)x, mean, rstd, HW, C,
                                                            act, racc);
  S2E_LAUNCH_CHECK();
  dim3 grid2(ew_chunks(HW, B, C), B);
  instnorm_bwd_apply_kernel<<<grid2, NT, 0, st>>>((const bf16*)dy, (const bf16*)y, (const bf16*)x, mean, rstd, racc, HW, C, act,
                                                  (bf16*)dx);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

}  // extern "C"
