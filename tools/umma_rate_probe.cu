// How long does one tcgen05.mma (cta_group::1, kind::f16, bf16 operands from shared memory, fp32 accumulator in TMEM) take as
// a function of its shape and of the operand layouts?  The forward / weight-gradient kernels of conv_tc.cu are scheduled
// around the answer:
//
//   * M = 128, N in {64, 128, 192, 256}: is the cost N / 2 cycles (the tensor-pipe floor) or is there a per-instruction floor
//     (the read of the 128 x 16 A operand) that makes narrow N expensive?
//   * K-major (forward kernel) versus MN-major (weight-gradient kernel) SWIZZLE_128B operands: same rate?
//
// Method: every SM runs one CTA whose shared memory holds an A region and a B region (contents irrelevant: zeros); one thread
// issues REPS x 4 MMAs (K = 16 each, the four k-slices of a 64-element swizzle row, exactly like the kernels' inner loop),
// commits to an mbarrier, waits, and reports clock64() / (REPS * 4).  No TMA, no epilogue: pure issue / operand-fetch rate.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/umma_rate_probe tools/umma_rate_probe.cu && /tmp/umma_rate_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../seg2eye_b200/csrc/common.cuh"

void s2e_set_error(const char* fmt, ...) { (void)fmt; }
int s2e_num_sms() { return 148; }
int s2e_debug_get(int) { return 0; }

namespace {

constexpr int REPS = 2000;

// layout 0: K-major (rows = M or N index, 128 B = 64 k per row; k-slice stride 32 B, 8-row group stride 1024 B)
// layout 1: MN-major (rows = k, 128 B = 64 M/N indices per row, 64-index atoms `blk` bytes apart; k-slice of 16 = 2048 B)
__global__ void __launch_bounds__(128, 1) rate_kernel(int M, int N, int a_mn, int b_mn, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 128 rows x 128 B (K-major) or 2 blocks x 64 k-rows x 128 B (MN-major): 16 KB
  uint8_t* sB = smem + 16384;         // up to 256 rows x 128 B = 32 KB
  uint64_t* bar = (uint64_t*)(smem + 16384 + 32768);
  uint32_t* tmem_slot = (uint32_t*)(bar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0u;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::umma_idesc_bf16(M, N, a_mn, b_mn);
    const uint32_t a0 = ptx::smem_u32(sA), b0 = ptx::smem_u32(sB);
    const long long t0 = clock64();
    for (int r = 0; r < REPS; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = a_mn ? ptx::umma_desc_sw128(a0 + k * 2048, 8192, 1024) : ptx::umma_desc_sw128(a0 + k * 32, 0, 1024);
        const uint64_t bd = b_mn ? ptx::umma_desc_sw128(b0 + k * 2048, 8192, 1024) : ptx::umma_desc_sw128(b0 + k * 32, 0, 1024);
        ptx::umma_bf16(tmem, ad, bd, idesc, 1u);
      }
    }
    ptx::umma_commit(bar);
    ptx::mbar_wait(bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 256);
}

}  // namespace

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d_out = nullptr;
  cudaMalloc(&d_out, sizeof(long long) * sms);
  const int smem_bytes = 16384 + 32768 + 1024 + 256;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  std::vector<long long> h(sms);
  printf("# tcgen05.mma cta_group::1 kind::f16, bf16 operands from shared memory, K = 16 per instruction, %d SMs busy\n", sms);
  printf("# %-4s %-4s %-9s %-9s %10s %12s %10s\n", "M", "N", "A layout", "B layout", "clk/MMA", "floor N*M/256", "MAC/clk");
  const int Ms[] = {128, 64};
  const int Ns[] = {64, 128, 192, 256};
  for (int mi = 0; mi < 2; ++mi)
    for (int a_mn = 0; a_mn < 2; ++a_mn)
      for (int b_mn = 0; b_mn < 2; ++b_mn)
        for (int ni = 0; ni < 4; ++ni) {
          const int M = Ms[mi], N = Ns[ni];
          if (M == 64 && (a_mn != b_mn)) continue;
          for (int rep = 0; rep < 2; ++rep) {   // first pass warms the clocks
            rate_kernel<<<sms, 128, smem_bytes>>>(M, N, a_mn, b_mn, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
              printf("M%d N%d a_mn%d b_mn%d: %s\n", M, N, a_mn, b_mn, cudaGetErrorString(e));
              return 1;
            }
          }
          cudaMemcpy(h.data(), d_out, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
          double sum = 0;
          for (int i = 0; i < sms; ++i) sum += (double)h[i];
          const double clk = sum / sms / (REPS * 4.0);
          printf("  %-4d %-4d %-9s %-9s %10.1f %12.1f %10.0f\n", M, N, a_mn ? "MN-major" : "K-major", b_mn ? "MN-major" : "K-major", clk,
                 (M < 128 ? 128 : M) * N / 256.0, (double)M * N * 16 / clk);
        }
  cudaFree(d_out);
  return 0;
}
