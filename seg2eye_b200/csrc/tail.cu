// Validation / inference tail and style-feature aggregation (sm_100a).
//
//  * s2e_to255_resize: data/postprocessor.py:57-114 -- the reference moves the generated batch to the host, resizes every
//    image with cv2.resize(INTER_LINEAR) in float64, maps [-1,1] -> [0,255] and truncates with .int().  Here: one pass on
//    the device in float64 with cv2's coordinate rule (half-pixel centres, fractional part computed with a fused
//    multiply-add like the AVX2 build of OpenCV the reference runs on, border rows/columns replicated), the int cast and,
//    when a 0..255 target is given, the per-image sum of squared differences for the OpenEDS score in the same pass.
//  * s2e_openeds_score: models/networks/loss.py:102-133 -- sqrt(sum (p - t)^2) / (H * W) per image; the sum is an exact
//    64-bit integer here (the reference adds the squares in fp32).
//  * s2e_aggregate_{fwd,bwd}: models/pix2pix_model.py:271-303 -- mean | max over the ns style images of a sample, for the
//    style codes (fp32) and the encoder feature maps (bf16 NHWC) of the optional style losses.
#include "common.cuh"

namespace {

constexpr int NT = 256;

__device__ __forceinline__ void cv_coord(int d, double scale, int src_len, int& s, double& f) {
  // fx = (dx + 0.5) * scale - 0.5 ; sx = floor(fx) ; fx -= sx      (resize.cpp, INTER_LINEAR coefficient table)
  const double c = __fma_rn((double)d + 0.5, scale, -0.5);
  const double fl = floor(c);
  s = (int)fl;
  f = __dsub_rn(c, fl);
  (void)src_len;
}

// one thread per destination pixel; sqsum[n] accumulates the exact integer sum of (out - target)^2
__global__ void to255_resize_kernel(const float* __restrict__ x, int N, int h, int w, int H, int W, double scale_y, double scale_x,
                                    int f32_path, const int* __restrict__ target, int* __restrict__ out,
                                    unsigned long long* __restrict__ sqsum) {
  const long long total = (long long)N * H * W;
  unsigned long long local = 0ull;
  int img = -1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int dx = (int)(i % W);
    const int dy = (int)((i / W) % H);
    const int n = (int)(i / ((long long)W * H));
    const float* src = x + (size_t)n * h * w;
    int v;
    if (f32_path) {
      // ImageProcessor.to_255imagebatch on the fp32 tensor itself (no resize): ((x + 1) * 255) / 2 in fp32, then .int()
      const float t = __fdiv_rn(__fmul_rn(__fadd_rn(src[(size_t)dy * w + dx], 1.0f), 255.0f), 2.0f);
      v = (int)t;
    } else {
      int sx, sy;
      double fx, fy;
      cv_coord(dx, scale_x, w, sx, fx);
      cv_coord(dy, scale_y, h, sy, fy);
      // horizontal rule: sx < 0 -> (0, fx = 0); sx >= w-1 -> (w-1, fx = 0)
      if (sx < 0) { sx = 0; fx = 0.0; }
      if (sx >= w - 1) { sx = w - 1; fx = 0.0; }
      const int sx1 = min(sx + 1, w - 1);
      // vertical rule: rows are clipped to [0, h-1], the weights are NOT changed
      const int r0 = min(max(sy, 0), h - 1), r1 = min(max(sy + 1, 0), h - 1);
      const double a0 = __dsub_rn(1.0, fx), a1 = fx, b0 = __dsub_rn(1.0, fy), b1 = fy;
      const double p00 = (double)src[(size_t)r0 * w + sx], p01 = (double)src[(size_t)r0 * w + sx1];
      const double p10 = (double)src[(size_t)r1 * w + sx], p11 = (double)src[(size_t)r1 * w + sx1];
      const bool single = sx + 1 >= w;   // dx >= xmax: the row pass copies S[sx]
      const double h0 = single ? p00 : __dadd_rn(__dmul_rn(p00, a0), __dmul_rn(p01, a1));
      const double h1 = single ? p10 : __dadd_rn(__dmul_rn(p10, a0), __dmul_rn(p11, a1));
      const double val = __dadd_rn(__dmul_rn(h0, b0), __dmul_rn(h1, b1));
      // unnormalize (postprocessor.py:62-64): add 1, mul 255, div 2 in float64, then .int() (truncation)
      const double t = __ddiv_rn(__dmul_rn(__dadd_rn(val, 1.0), 255.0), 2.0);
      v = (int)t;
    }
    out[i] = v;
    if (target) {
      if (n != img) {
        if (img >= 0 && local) atomicAdd(sqsum + img, local);
        img = n;
        local = 0ull;
      }
      const long long d = (long long)v - (long long)target[i];
      local += (unsigned long long)(d * d);
    }
  }
  if (target && img >= 0 && local) atomicAdd(sqsum + img, local);
}

__global__ void sqdiff_kernel(const int* __restrict__ p, const int* __restrict__ t, int N, long long HW,
                              unsigned long long* __restrict__ sqsum) {
  const int n = blockIdx.y;
  unsigned long long local = 0ull;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    const long long d = (long long)p[(size_t)n * HW + i] - (long long)t[(size_t)n * HW + i];
    local += (unsigned long long)(d * d);
  }
  // warp reduction, one atomic per warp
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(sqsum + n, local);
}

__global__ void score_finalize_kernel(const unsigned long long* __restrict__ sqsum, int N, float hw, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  // loss.py:106-110: torch.sqrt(sum.float()) / (h * w), in fp32 (explicit _rn intrinsics: the library is built with fast-math)
  if (n < N) out[n] = __fdiv_rn(__fsqrt_rn((float)sqsum[n]), hw);
}

template <typename T>
__device__ __forceinline__ float ldf(const T* p, size_t i);
template <>
__device__ __forceinline__ float ldf<float>(const float* p, size_t i) { return p[i]; }
template <>
__device__ __forceinline__ float ldf<bf16>(const bf16* p, size_t i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ void stf(float* p, size_t i, float v) { p[i] = v; }
__device__ __forceinline__ void stf(bf16* p, size_t i, float v) { p[i] = __float2bfloat16(v); }

// x [G][ns][n] -> out [G][n] fp32 (mean or max over ns); arg[G][n] = index of the (first) maximum
template <typename T>
__global__ void aggregate_fwd_kernel(const T* __restrict__ x, int ns, long long n, long long total, int mode, float* __restrict__ out,
                                     uint8_t* __restrict__ arg) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long g = i / n, e = i - g * n;
    const T* base = x + (size_t)g * ns * n + e;
    if (mode == 0) {
      float s = 0.f;
      for (int k = 0; k < ns; ++k) s += ldf<T>(base, (size_t)k * n);
      out[i] = s / (float)ns;
    } else {
      float best = ldf<T>(base, 0);
      int bi = 0;
      for (int k = 1; k < ns; ++k) {
        const float v = ldf<T>(base, (size_t)k * n);
        if (v > best) { best = v; bi = k; }
      }
      out[i] = best;
      if (arg) arg[i] = (uint8_t)bi;
    }
  }
}
template <typename T>
__global__ void aggregate_bwd_kernel(const float* __restrict__ dout, const uint8_t* __restrict__ arg, int ns, long long n, long long total,
                                     int mode, T* __restrict__ dx) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long g = i / n, e = i - g * n;
    T* base = dx + (size_t)g * ns * n + e;
    const float d = dout[i];
    for (int k = 0; k < ns; ++k) {
      const float v = mode == 0 ? d / (float)ns : (k == (int)arg[i] ? d : 0.f);
      stf(base, (size_t)k * n, v);
    }
  }
}

inline int grid_for(long long n) {
  long long g = (n + NT - 1) / NT;
  const long long cap = (long long)s2e_num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" {

int s2e_to255_resize(const float* x, int N, int h, int w, int H, int W, int f32_path, const int* target, int* out,
                     unsigned long long* sqsum, float* score, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  S2E_REQUIRE(N > 0 && h > 0 && w > 0 && H > 0 && W > 0, "to255_resize: bad shape");
  S2E_REQUIRE(!f32_path || (h == H && w == W), "to255_resize: the fp32 path does not resize");
  S2E_REQUIRE(!target || sqsum, "to255_resize: a target needs the sqsum scratch (N x u64)");
  // cv2: inv_scale = (double)dst / src; scale = 1. / inv_scale
  const double scale_x = 1.0 / ((double)W / (double)w), scale_y = 1.0 / ((double)H / (double)h);
  if (target) S2E_CHECK_CUDA(cudaMemsetAsync(sqsum, 0, sizeof(unsigned long long) * N, st));
  to255_resize_kernel<<<grid_for((long long)N * H * W), NT, 0, st>>>(x, N, h, w, H, W, scale_y, scale_x, f32_path, target, out, sqsum);
  S2E_LAUNCH_CHECK();
  if (target && score) {
    score_finalize_kernel<<<(N + 127) / 128, 128, 0, st>>>(sqsum, N, (float)H * (float)W, score);
    S2E_LAUNCH_CHECK();
  }
  return S2E_OK;
}

int s2e_openeds_score(const int* produced, const int* target, int N, int H, int W, unsigned long long* sqsum, float* score,
                      void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  S2E_REQUIRE(N > 0 && H > 0 && W > 0, "openeds_score: bad shape");
  S2E_CHECK_CUDA(cudaMemsetAsync(sqsum, 0, sizeof(unsigned long long) * N, st));
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)min((long long)64, (HW + NT - 1) / NT), N);
  sqdiff_kernel<<<grid, NT, 0, st>>>(produced, target, N, HW, sqsum);
  S2E_LAUNCH_CHECK();
  score_finalize_kernel<<<(N + 127) / 128, 128, 0, st>>>(sqsum, N, (float)H * (float)W, score);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

int s2e_aggregate_fwd(const void* x, int x_is_f32, int G, int ns, long long n, int mode, float* out, uint8_t* argmax, void* stream) {
  S2E_REQUIRE(G > 0 && ns > 0 && ns < 256 && n > 0 && (mode == 0 || mode == 1), "aggregate_fwd: bad arguments");
  const long long total = (long long)G * n;
  if (x_is_f32)
    aggregate_fwd_kernel<float><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>((const float*)x, ns, n, total, mode, out, argmax);
  else
    aggregate_fwd_kernel<bf16><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>((const bf16*)x, ns, n, total, mode, out, argmax);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

int s2e_aggregate_bwd(const float* dout, const uint8_t* argmax, int x_is_f32, int G, int ns, long long n, int mode, void* dx,
                      void* stream) {
  S2E_REQUIRE(G > 0 && ns > 0 && ns < 256 && n > 0 && (mode == 0 || (mode == 1 && argmax)), "aggregate_bwd: bad arguments");
  const long long total = (long long)G * n;
  if (x_is_f32)
    aggregate_bwd_kernel<float><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>(dout, argmax, ns, n, total, mode, (float*)dx);
  else
    aggregate_bwd_kernel<bf16><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>(dout, argmax, ns, n, total, mode, (bf16*)dx);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

}  // extern "C"
