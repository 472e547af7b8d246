// Device-side data layer: the per-sample preprocessing of data/openeds_dataset.py:82-119 with
// data/base_dataset.py:50-80 in 'fixed' mode, on raw uint8 OpenEDS frames already resident in HBM.
//   label : cv2.resize(mask, (w, h), INTER_NEAREST) -> optional horizontal flip -> int64 (N,1,h,w)
//   image : PIL Image.resize((w, h), BICUBIC) -- separable, 22-bit fixed-point taps, 8-bit intermediate -- -> flip ->
//           ToTensor (u8 / 255) -> Normalize(0.5, 0.5) -> fp32 (N,1,h,w)
// Integer results are bit-exact by construction (the tap tables are computed on the host exactly like Pillow's
// precompute_coeffs / normalize_coeffs_8bpc and passed in); the float results use explicitly rounded fp32 operations.
#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr int PIL_PRECISION_BITS = 32 - 8 - 2;

inline int grid_for(long long n) {
  long long g = (n + NT - 1) / NT;
  const long long cap = (long long)s2e_num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

__global__ void label_nearest_kernel(const uint8_t* __restrict__ mask, int N, int H0, int W0, int h, int w, double ify, double ifx,
                                     const uint8_t* __restrict__ flip, long long* __restrict__ out) {
  const long long total = (long long)N * h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w), y = (int)((i / w) % h), n = (int)(i / ((long long)w * h));
    const int xs = (flip && flip[n]) ? w - 1 - x : x;      // flip AFTER the resize: output x reads resized column w-1-x
    int sx = (int)floor((double)xs * ifx), sy = (int)floor((double)y * ify);
    sx = min(sx, W0 - 1);
    sy = min(sy, H0 - 1);
    out[i] = (long long)mask[((size_t)n * H0 + sy) * W0 + sx];
  }
}

// one separable pass of Pillow's 8-bit resampling along x (horizontal != 0) or y
__global__ void pil_resample_kernel(const uint8_t* __restrict__ in, int N, int Hin, int Win, int Hout, int Wout, int horizontal,
                                    const int* __restrict__ kk, const int* __restrict__ bounds, int ksize, uint8_t* __restrict__ out) {
  const long long total = (long long)N * Hout * Wout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wout), y = (int)((i / Wout) % Hout), n = (int)(i / ((long long)Wout * Hout));
    const int o = horizontal ? x : y;
    const int first = bounds[2 * o], cnt = bounds[2 * o + 1];
    const int* k = kk + (size_t)o * ksize;
    const uint8_t* src = in + (size_t)n * Hin * Win;
    int ss = 1 << (PIL_PRECISION_BITS - 1);
    if (horizontal) {
      const uint8_t* row = src + (size_t)y * Win + first;
      for (int t = 0; t < cnt; ++t) ss += (int)row[t] * k[t];
    } else {
      const uint8_t* col = src + (size_t)first * Win + x;
      for (int t = 0; t < cnt; ++t) ss += (int)col[(size_t)t * Win] * k[t];
    }
    const int v = ss >> PIL_PRECISION_BITS;     // arithmetic shift, then clip8
    out[i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
  }
}

__global__ void u8_normalize_kernel(const uint8_t* __restrict__ in, int N, int per_flag, int h, int w, const uint8_t* __restrict__ flip,
                                    float* __restrict__ out) {
  const long long total = (long long)N * h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int n = (int)(i / ((long long)w * h));
    const bool f = flip && flip[n / per_flag];
    const uint8_t v = in[f ? i - x + (w - 1 - x) : i];
    // ToTensor: uint8 -> float32, div(255); Normalize: sub(0.5).div(0.5)  (torchvision functional_tensor, fp32 throughout)
    out[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), 0.5f), 0.5f);
  }
}

__global__ void u8_flip_to_i32_kernel(const uint8_t* __restrict__ in, int N, int h, int w, const uint8_t* __restrict__ flip, int* __restrict__ out) {
  const long long total = (long long)N * h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int n = (int)(i / ((long long)w * h));
    out[i] = (int)in[(flip && flip[n]) ? i - x + (w - 1 - x) : i];
  }
}

}  // namespace

extern "C" {

int s2e_label_nearest_flip(const uint8_t* mask, int N, int H0, int W0, int h, int w, const uint8_t* flip, int64_t* out, void* stream) {
  S2E_REQUIRE(N > 0 && H0 > 0 && W0 > 0 && h > 0 && w > 0, "label_nearest_flip: bad shape");
  // cv2 resizeNN: ifx = 1 / (dsize.width / ssize.width) in double, sx = min(cvFloor(x * ifx), ssize.width - 1)
  const double ifx = 1.0 / ((double)w / (double)W0), ify = 1.0 / ((double)h / (double)H0);
  label_nearest_kernel<<<grid_for((long long)N * h * w), NT, 0, (cudaStream_t)stream>>>(mask, N, H0, W0, h, w, ify, ifx, flip, (long long*)out);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

int s2e_pil_resample_u8(const uint8_t* in, int N, int Hin, int Win, int out_size, int horizontal, const int* kk, const int* bounds,
                        int ksize, uint8_t* out, void* stream) {
  S2E_REQUIRE(N > 0 && Hin > 0 && Win > 0 && out_size > 0 && ksize > 0, "pil_resample_u8: bad shape");
  const int Hout = horizontal ? Hin : out_size, Wout = horizontal ? out_size : Win;
  pil_resample_kernel<<<grid_for((long long)N * Hout * Wout), NT, 0, (cudaStream_t)stream>>>(in, N, Hin, Win, Hout, Wout, horizontal, kk,
                                                                                          bounds, ksize, out);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

int s2e_u8_flip_normalize(const uint8_t* in, int N, int images_per_flag, int h, int w, const uint8_t* flip, float* out, void* stream) {
  S2E_REQUIRE(N > 0 && h > 0 && w > 0 && images_per_flag > 0, "u8_flip_normalize: bad shape");
  u8_normalize_kernel<<<grid_for((long long)N * h * w), NT, 0, (cudaStream_t)stream>>>(in, N, images_per_flag, h, w, flip, out);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

int s2e_u8_flip_to_i32(const uint8_t* in, int N, int h, int w, const uint8_t* flip, int* out, void* stream) {
  S2E_REQUIRE(N > 0 && h > 0 && w > 0, "u8_flip_to_i32: bad shape");
  u8_flip_to_i32_kernel<<<grid_for((long long)N * h * w), NT, 0, (cudaStream_t)stream>>>(in, N, h, w, flip, out);
  S2E_LAUNCH_CHECK();
  return S2E_OK;
}

}  // extern "C"
