// Microbenchmark: what does one layer-to-layer activation exchange cost when the eight CTAs that own the output-channel
// slices of one board are a thread-block cluster and hand their 16-bit activations to each other through distributed
// shared memory (st.shared::cluster + remote mbarrier arrive), instead of global memory + progress flags + TMA?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ub/ub_dsmem tools/ub/ub_dsmem.cu && tools/ub/ub_dsmem
// Per iteration every CTA: 64 "row" threads store 4 x 16 B to each of the 8 CTAs' next tile buffer, the producing warps
// meet on a named barrier, one thread release-arrives on the 8 CTAs' mbarriers; one consumer thread waits for its own
// barrier's 8 arrivals (acquire.cluster) and issues fence.proxy.async, then tells the producers to go on.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CL = 8, TILE = 51200, ITERS = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_remote_v4(uint32_t a, uint4 v) {
  asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void arrive_remote(uint32_t a) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ bool try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(256) exchange_kernel(long long* cycles, unsigned* sums, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[2], go;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[b])), "r"(CL));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&go)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 2 * TILE / 16; i += 256) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  long long t0 = 0;
  if (tid == 0) t0 = clock64();
  if (tid >= 128) {   // producers: 4 warps, 64 of their threads hold a board row each
    const int p = tid - 128, row = p & 63;
    const bool has_row = p < 64 || mode == 1;   // mode 1: all 128 threads store (two boards)
    for (int it = 0; it < ITERS; ++it) {
      const int buf = (it + 1) & 1;
      if (it > 0) {   // wait until this CTA's consumer has seen the previous layer (stands for "MMAs done, accumulator read")
        while (true) {
          uint32_t ok;
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(smem_u32(&go)), "r"((it - 1) & 1) : "memory");
          if (ok) break;
        }
      }
      if (has_row) {
        const uint32_t base = smem_u32(smem) + buf * TILE + (rank >> 1) * 12800 + (11 + row + 2 * (row >> 3)) * 128 + (rank & 1) * 64;
        const uint4 v = make_uint4(it + 1, rank, row, 0x1234);
#pragma unroll
        for (int d = 0; d < CL; ++d) {
          const uint32_t r = mapa(base, d);
#pragma unroll
          for (int u = 0; u < 4; ++u) st_remote_v4(r + ((u ^ (row & 3)) << 4), v);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (p == 0) {
        asm volatile("fence.acq_rel.cluster;" ::: "memory");
        const uint32_t b = smem_u32(&full[buf]);
#pragma unroll
        for (int d = 0; d < CL; ++d) arrive_remote(mapa(b, d));
      }
    }
  } else if (tid == 0) {   // consumer
    unsigned acc = 0;
    for (int it = 0; it < ITERS; ++it) {
      const int buf = (it + 1) & 1;
      unsigned spins = 0;
      while (!try_wait_cluster(smem_u32(&full[buf]), (it >> 1) & 1) && ++spins < (1u << 24)) {}
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      acc += reinterpret_cast<volatile unsigned*>(smem + buf * TILE + 11 * 128)[0];   // the stored iteration number
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&go)) : "memory");
    }
    cycles[blockIdx.x] = clock64() - t0;
    sums[blockIdx.x] = acc;
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int main() {
  const int grid = 128, smem = 2 * TILE + 100 * 1024;   // one CTA per SM, like the forward kernel
  cudaFuncSetAttribute(exchange_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int max_clusters = 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
  at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaOccupancyMaxActiveClusters(&max_clusters, exchange_kernel, &cfg);
  printf("co-resident clusters of %d at %d B of shared memory: %d (needed: %d)\n", CL, smem, max_clusters, grid / CL);
  long long* cyc;
  unsigned* sums;
  cudaMallocManaged(&cyc, grid * sizeof(long long));
  cudaMallocManaged(&sums, grid * sizeof(unsigned));
  for (int mode = 0; mode < 2; ++mode)
    for (int rep = 0; rep < 3; ++rep) {
      exchange_kernel<<<grid, 256, smem>>>(cyc, sums, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      long long mx = 0, mn = 1ll << 60;
      bool ok = true;
      for (int i = 0; i < grid; ++i) {
        mx = cyc[i] > mx ? cyc[i] : mx;
        mn = cyc[i] < mn ? cyc[i] : mn;
        ok = ok && sums[i] == (unsigned)(ITERS * (ITERS + 1) / 2);
      }
      printf("mode %d (%d storing threads): %.0f..%.0f cycles per exchange, data %s\n", mode, mode ? 128 : 64, (double)mn / ITERS, (double)mx / ITERS, ok ? "ok" : "WRONG");
    }
  return 0;
}
