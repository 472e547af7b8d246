// Hardware question behind DESIGN.md section 8 item 2(b) (one halo tile serving all nine taps of a 3x3 convolution):
// can a K-major SWIZZLE_128B UMMA A-operand descriptor start at an ARBITRARY 128-byte row of a tile that TMA wrote with
// the same swizzle, and with a group stride (SBO) that is not 1024 bytes?  If yes, the shifted windows of a halo tile can
// be fed to tcgen05.mma directly and the forward / data-gradient kernels stop re-reading the input once per tap.
//
// Experiment: A_full = [ROWS x 64] bf16 with A_full[r][k] = (7 r + 3 k) mod 251, B = 64x64 identity, so
// D[m][n] = A_window[m][n] reveals which shared-memory row / 16-byte chunk the tensor core read for output row m.
// For every (row shift s in 0..9) x (group pitch in {8, 10, 16} rows) x (base_offset in {0, (addr>>7)&7}) the program
// prints whether D equals the expected window.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/umma_shift_probe tools/umma_shift_probe.cu && /tmp/umma_shift_probe
//
// Not part of the library; run it under gpurun before building the halo kernel.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../seg2eye_b200/csrc/common.cuh"

void s2e_set_error(const char* fmt, ...) { (void)fmt; }   // common.cuh declares these; the probe does not use them
int s2e_num_sms() { return 148; }
int s2e_debug_get(int) { return 0; }

namespace {

constexpr int ROWS = 512;   // rows of the A tile in shared memory (64 KB)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool make_map_2d(CUtensorMap* m, const void* ptr, int rows, int box_rows) {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) return false;
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  return ((EncodeTiledFn)fp)(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   ptx::smem_u32(dst)),
               "l"(m), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

// descriptor with an explicit base_offset field (bits [49, 52))
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t base_offset) {
  return ptx::umma_desc_sw128(addr, lbo, sbo) | ((uint64_t)(base_offset & 7) << 49);
}

// one CTA, 128 threads: TMA-load A (ROWS x 64) and B (64 x 64), one M128 x N64 x K64 MMA from the shifted window, D -> global
__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                       int shift_rows, int pitch_rows, int use_base_offset, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                        // ROWS * 128 bytes
  uint8_t* sB = smem + ROWS * 128;           // 64 * 128 bytes
  uint64_t* bar_load = (uint64_t*)(sB + 64 * 128);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = (uint32_t*)(bar_mma + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_load, 1);
    ptx::mbar_init(bar_mma, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, 64);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(bar_load, ROWS * 128 + 64 * 128);
    for (int r0 = 0; r0 < ROWS; r0 += 256) tma_load_2d(sA + r0 * 128, &tmA, bar_load, 0, r0);
    tma_load_2d(sB, &tmB, bar_load, 0, 0);
    ptx::mbar_wait(bar_load, 0);
    ptx::tc_fence_after();
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, 64, 0, 0);
    const uint32_t a0 = ptx::smem_u32(sA) + (uint32_t)shift_rows * 128u;
    const uint32_t sbo = (uint32_t)pitch_rows * 128u;                       // stride between 8-row groups
    const uint32_t bo = use_base_offset ? ((a0 >> 7) & 7u) : 0u;
    for (int k = 0; k < 4; ++k) {
      const uint64_t ad = desc_sw128(a0 + k * 32, 0, sbo, bo);
      const uint64_t bd = desc_sw128(ptx::smem_u32(sB) + k * 32, 0, 1024, 0);
      ptx::umma_bf16(tmem_base, ad, bd, idesc, k != 0 ? 1u : 0u);
    }
    ptx::umma_commit(bar_mma);
  }
  __syncthreads();
  ptx::mbar_wait(bar_mma, 0);
  ptx::tc_fence_after();
  uint32_t r[32];
  for (int c0 = 0; c0 < 64; c0 += 32) {
    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 64);
}

// Second question (weight-gradient kernel): both operands are MN-major there (pixels = K, one 128-byte row per pixel).
// Can the B-operand descriptor of an MN-major SWIZZLE_128B tile start at an arbitrary pixel row, so that ONE staged X tile
// with a halo serves the dx = -1 / 0 / +1 taps?  A = K-major selection matrix (A[m][k] = 1 iff k == m % 64), B = rows
// [shift, shift + 64) of the staged [ROWS x 64] tile read MN-major, so D[m][n] = X[shift + m % 64][n].
__global__ void __launch_bounds__(128, 1) probe_mn_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmSel,
                                                          int shift_rows, int use_base_offset, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sX = smem;                        // ROWS * 128 bytes
  uint8_t* sSel = smem + ROWS * 128;         // 128 * 128 bytes (M = 128 rows x K = 64)
  uint64_t* bar_load = (uint64_t*)(sSel + 128 * 128);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = (uint32_t*)(bar_mma + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_load, 1);
    ptx::mbar_init(bar_mma, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, 64);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(bar_load, ROWS * 128 + 128 * 128);
    for (int r0 = 0; r0 < ROWS; r0 += 256) tma_load_2d(sX + r0 * 128, &tmX, bar_load, 0, r0);
    tma_load_2d(sSel, &tmSel, bar_load, 0, 0);
    ptx::mbar_wait(bar_load, 0);
    ptx::tc_fence_after();
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, 64, 0, 1);   // A K-major, B MN-major
    const uint32_t b0 = ptx::smem_u32(sX) + (uint32_t)shift_rows * 128u;
    const uint32_t bo = use_base_offset ? ((b0 >> 7) & 7u) : 0u;
    for (int k = 0; k < 4; ++k) {
      const uint64_t ad = desc_sw128(ptx::smem_u32(sSel) + k * 32, 0, 1024, 0);
      const uint64_t bd = desc_sw128(b0 + k * 2048, 8192, 1024, bo);   // 16 pixel rows per K = 16 step
      ptx::umma_bf16(tmem_base, ad, bd, idesc, k != 0 ? 1u : 0u);
    }
    ptx::umma_commit(bar_mma);
  }
  __syncthreads();
  ptx::mbar_wait(bar_mma, 0);
  ptx::tc_fence_after();
  uint32_t r[32];
  for (int c0 = 0; c0 < 64; c0 += 32) {
    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 64);
}

}  // namespace

int main() {
  // A_full[r][k] = (7 r + 3 k) mod 251: exactly representable in bf16, and any wrong row or 16-byte chunk changes the result
  std::vector<__nv_bfloat16> hA(ROWS * 64), hB(64 * 64);
  for (int r = 0; r < ROWS; ++r)
    for (int k = 0; k < 64; ++k) hA[r * 64 + k] = __float2bfloat16((float)((r * 7 + k * 3) % 251));   // exact in bf16 (< 256)
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *dA, *dB;
  float* dOut;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dOut, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  if (!make_map_2d(&tmA, dA, ROWS, 256) || !make_map_2d(&tmB, dB, 64, 64)) {
    printf("tensor map creation failed\n");
    return 1;
  }
  const int smem_bytes = ROWS * 128 + 64 * 128 + 1024 + 256;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  std::vector<float> hOut(128 * 64);
  printf("pitch = rows between consecutive 8-row groups (8 = dense tile, 10 = halo tile of width 8+2, 16 = padded halo tile)\n");
  for (int pitch : {8, 10, 16}) {
    for (int bo = 0; bo < 2; ++bo) {
      printf("pitch %2d  base_offset %s :", pitch, bo ? "(addr>>7)&7" : "0          ");
      for (int shift = 0; shift < 10; ++shift) {
        cudaMemset(dOut, 0, 128 * 64 * 4);
        probe_kernel<<<1, 128, smem_bytes>>>(tmA, tmB, shift, pitch, bo, dOut);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf(" [shift %d: %s]", shift, cudaGetErrorString(e));
          return 2;
        }
        cudaMemcpy(hOut.data(), dOut, hOut.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m) {
          const int src = shift + (m / 8) * pitch + (m % 8);   // the row the window is supposed to read
          for (int n = 0; n < 64; ++n)
            if (hOut[m * 64 + n] != __bfloat162float(hA[src * 64 + n])) ++bad;
        }
        printf(" s%d:%s", shift, bad == 0 ? "OK" : "x ");
      }
      printf("\n");
    }
  }

  // ---- MN-major B operand (weight-gradient layout) read from a shifted pixel row
  {
    std::vector<__nv_bfloat16> hSel(128 * 64);
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < 64; ++k) hSel[m * 64 + k] = __float2bfloat16(k == m % 64 ? 1.f : 0.f);
    __nv_bfloat16* dSel;
    cudaMalloc(&dSel, hSel.size() * 2);
    cudaMemcpy(dSel, hSel.data(), hSel.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tmSel;
    if (!make_map_2d(&tmSel, dSel, 128, 128)) {
      printf("tensor map creation failed (sel)\n");
      return 1;
    }
    const int smem2 = ROWS * 128 + 128 * 128 + 1024 + 256;
    cudaFuncSetAttribute(probe_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
    for (int bo = 0; bo < 2; ++bo) {
      printf("MN-major B, base_offset %s :", bo ? "(addr>>7)&7" : "0          ");
      for (int shift = 0; shift < 12; ++shift) {
        cudaMemset(dOut, 0, 128 * 64 * 4);
        probe_mn_kernel<<<1, 128, smem2>>>(tmA, tmSel, shift, bo, dOut);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf(" [shift %d: %s]", shift, cudaGetErrorString(e));
          return 2;
        }
        cudaMemcpy(hOut.data(), dOut, hOut.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n)
            if (hOut[m * 64 + n] != __bfloat162float(hA[(shift + m % 64) * 64 + n])) ++bad;
        printf(" s%d:%s", shift, bad == 0 ? "OK" : "x ");
      }
      printf("\n");
    }
  }
  printf("OK for all shifts with one (pitch, base_offset) row => shifted windows of a halo tile can feed tcgen05.mma directly.\n");
  return 0;
}
