// Where do the rows of a cta_group::2, M = 128 accumulator live?  Two CTAs, each supplies 64 rows of A (value
// row + 1 + 100 * cta in every K position) and 8 of the 16 rows of B (value (n + 1) / 16), so that after one K = 16 MMA
// D[m][n] = (m_id)(n + 1) names its own row and column.  Every CTA dumps its 128 TMEM lanes x 16 columns.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I chess-alpha-zero_b200/csrc -o tools/ub/ub_pair_layout tools/ub/ub_pair_layout.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"
using namespace caz;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_kernel(float* out, int m_total, int n) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __nv_bfloat16* a = reinterpret_cast<__nv_bfloat16*>(smem);            // [rows per CTA][64]
  __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(smem + 32768);    // [n/2][64]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t rank = ptx::cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = m_total / 2;
  for (int i = threadIdx.x; i < rows * 64; i += 128) a[i] = __float2bfloat16(static_cast<float>(i / 64 + 1 + 100 * rank));
  for (int i = threadIdx.x; i < (n / 2) * 64; i += 128) b[i] = __float2bfloat16(static_cast<float>(rank * (n / 2) + i / 64 + 1) / 16.0f);
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::mbar_fence_init();
  }
  ptx::fence_proxy_async();
  if (warp == 2) {
    ptx::tmem_alloc_2cta(&tmem_slot, 32);
    ptx::tmem_relinquish_2cta();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  if (rank == 0 && warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::idesc_bf16_f32(m_total, n);
      ptx::mma_bf16_ss_2cta(tmem, desc_sw128(ptx::smem_u32(a), 1024), desc_sw128(ptx::smem_u32(b), 1024), idesc, 0u);
      ptx::mma_commit_2cta(&bar);
    }
    __syncwarp();
  }
  ptx::mbar_wait(&bar, 0, nullptr, 0);
  ptx::tcgen05_fence_after();
  uint32_t v[16];
  ptx::tmem_ld_32x16(tmem + (static_cast<uint32_t>(warp * 32) << 16), v);
  ptx::tmem_ld_wait();
  for (int j = 0; j < 16; ++j) out[(rank * 128 + warp * 32 + lane) * 16 + j] = __uint_as_float(v[j]);
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc_2cta(tmem, 32);
  }
}

int main() {
  float* d;
  cudaMalloc(&d, 2 * 128 * 16 * sizeof(float));
  cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  for (int m_total : {128, 256}) {
    cudaMemset(d, 0xff, 2 * 128 * 16 * sizeof(float));
    pair_kernel<<<2, 128, 80 * 1024>>>(d, m_total, 16);
    cudaError_t e = cudaDeviceSynchronize();
    printf("M=%d: %s\n", m_total, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> h(2 * 128 * 16);
    cudaMemcpy(h.data(), d, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
    for (int c = 0; c < 2; ++c) {
      printf(" CTA %d: lane -> (row id from column 0, column 1 / column 0)\n", c);
      for (int l = 0; l < 128; l += 1) {
        const float x0 = h[(c * 128 + l) * 16], x1 = h[(c * 128 + l) * 16 + 1], x15 = h[(c * 128 + l) * 16 + 15];
        if (l % 8 == 0) printf("  lane %3d:", l);
        printf(" %6.0f/%4.1f/%5.1f", x0, x0 != 0 ? x1 / x0 : 0.f, x0 != 0 ? x15 / x0 : 0.f);
        if (l % 8 == 7) printf("\n");
      }
    }
  }
  return 0;
}
