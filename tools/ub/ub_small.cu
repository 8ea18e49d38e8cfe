// Micro-benchmarks behind the small-batch kernel's design: what does a tcgen05.mma cost at small N,
// how fast is a DSMEM push inside a cluster, how long is a cluster barrier, how many clusters fit at once.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I chess-alpha-zero_b200/csrc -o tools/ub/ub_small tools/ub/ub_small.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace caz;

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);           \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// ---- A: cost of T back-to-back MMAs (M x N x 16), rotating over `nacc` accumulators, issued by `issuers` warps ----
struct MmaArgs { int m, n, nacc, issuers, total; long long* out; int halo; };  // halo: A read as shifted views of a 10-file halo tile

__global__ void __launch_bounds__(128, 1) mma_cost_kernel(MmaArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, a.issuers);
    ptx::mbar_fence_init();
    ptx::fence_proxy_async();
  }
  if (warp == 2) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp < a.issuers) {
    const uint32_t idesc = ptx::idesc_bf16_f32(a.m, a.n);
    // A: 32 KB region at 0 (128 rows x 64 ch x 4 chunks), B: region at 48 KB
    const uint64_t a0 = desc_sw128(ptx::smem_u32(smem), a.halo ? 1280 : 1024);
    const int row_pitch = a.m == 128 ? 20 : 10;
    const uint64_t b0 = desc_sw128(ptx::smem_u32(smem + 48 * 1024), 1024);
    const int per = a.total / a.issuers;
    const int bch = (48 * 1024) / (a.n * 128) > 0 ? (48 * 1024) / (a.n * 128) : 1;
    const int acc_per = a.nacc / a.issuers > 0 ? a.nacc / a.issuers : 1;
    __syncwarp();
    t0 = clock64();
    if (ptx::elect_one()) {
      const uint32_t tm = tmem + warp * 2 * a.n;          // two accumulators per issuer
      const uint64_t rp8 = static_cast<uint64_t>(row_pitch * 8);
      const uint64_t bstep = static_cast<uint64_t>((a.n * 128) >> 4);
#pragma unroll 1
      for (int it = 0; it < per / 36; ++it) {
        const uint64_t ac = a.halo ? a0 : a0 + static_cast<uint64_t>(((it & 1) * 16384) >> 4);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint64_t ash = a.halo ? (tap / 3) * rp8 + (tap % 3) * 8 : 0;
          const uint64_t bsh = (tap % 3) * bstep;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            ptx::mma_bf16_ss(tm + (kk & 1) * a.n, ac + ash + kk * 2, b0 + bsh + kk * 2, idesc, (it | tap | (kk >> 1)) ? 1u : 0u);
        }
      }
      ptx::mma_commit(&bar);
    }
    __syncwarp();
    t1 = clock64();
    ptx::mbar_wait(&bar, 0, nullptr, 0);
    t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) {
      a.out[warp * 2] = t1 - t0;
      a.out[warp * 2 + 1] = t2 - t0;
    }
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// ---- B: DSMEM push: every CTA of the cluster writes `bytes_per_peer` to each other CTA, then a cluster barrier ----
struct PushArgs { int bytes_per_peer; int mode; int iters; long long* out; };

__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void bulk_copy_cluster(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}

__global__ void __launch_bounds__(256, 1) push_kernel(PushArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // [0, 64K): source; [64K, 192K): receive slots, one per source rank
  uint8_t* src = smem;
  uint8_t* rcv = smem + 64 * 1024;
  __shared__ uint64_t bar;
  const uint32_t rank = ptx::cluster_ctarank(), csz = cluster_nctarank();
  for (int i = threadIdx.x; i < 16 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(src)[i] = i + rank;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::mbar_fence_init();
    ptx::fence_proxy_async();
  }
  __syncthreads();
  ptx::cluster_sync_all();
  const int t = threadIdx.x - 128;  // the 128 "epilogue" threads push
  long long t0 = clock64();
  for (int it = 0; it < a.iters; ++it) {
    if (a.mode == 0) {
      if (t >= 0) {
        // thread-per-16B, peers interleaved so that all links are busy at once
        const int vec_per_peer = a.bytes_per_peer / 16;
        for (int v = t; v < vec_per_peer; v += 128) {
          const uint4 val = *reinterpret_cast<const uint4*>(src + v * 16);
          for (uint32_t p = 1; p < csz; ++p) {
            const uint32_t peer = (rank + p) % csz;
            st_cluster_v4(ptx::mapa_u32(rcv + rank * a.bytes_per_peer + v * 16, peer), val);
          }
        }
      }
      ptx::cluster_sync_all();
    } else {
      // bulk copies: one elected thread, completion counted on the destination's mbarrier
      if (threadIdx.x == 0) {
        ptx::fence_proxy_async();
        ptx::mbar_expect_tx(&bar, (csz - 1) * a.bytes_per_peer);
        for (uint32_t p = 1; p < csz; ++p) {
          const uint32_t peer = (rank + p) % csz;
          bulk_copy_cluster(ptx::mapa_u32(rcv + rank * a.bytes_per_peer, peer), ptx::smem_u32(src), a.bytes_per_peer,
                            ptx::mapa_u32(&bar, peer));
        }
      }
      ptx::mbar_wait(&bar, it & 1, nullptr, 0);
      ptx::cluster_sync_all();   // receive slots may be overwritten by the next round
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 128 && blockIdx.x == 0) a.out[0] = (t1 - t0) / a.iters;
  // checksum so that nothing is optimised away
  if (threadIdx.x == 0 && blockIdx.x == 0) a.out[1] = reinterpret_cast<uint32_t*>(rcv)[((rank + 1) % csz) * a.bytes_per_peer / 4 + 1];
}

// ---- C: cluster barrier latency ----
__global__ void __launch_bounds__(256, 1) csync_kernel(int iters, long long* out) {
  ptx::cluster_sync_all();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) ptx::cluster_sync_all();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / iters;
}

template <typename K>
static void launch_cluster(K kernel, int grid, int block, int smem, int csz, void** params) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csz;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CK(cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(kernel), params));
}

int main() {
  long long* d_out;
  CK(cudaMalloc(&d_out, 64 * sizeof(long long)));
  long long h[64];

  // D: how many clusters are co-resident with ~200 KB of shared memory per CTA
  {
    const int smem = 200 * 1024;
    CK(cudaFuncSetAttribute(push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(push_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (int csz : {2, 4, 8, 16}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(csz * 32);
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = csz;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, reinterpret_cast<const void*>(push_kernel), &cfg);
      printf("max active clusters: cluster_size=%d smem=%d -> %d (%s)\n", csz, smem, n, cudaGetErrorString(e));
    }
  }

  // A
  CK(cudaFuncSetAttribute(mma_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  printf("\nMMA cost (cycles per MMA: issue-side, issue->complete), 144 MMAs K=16\n");
  for (int m : {64, 128})
    for (int n : {16, 32, 64})
      for (int issuers : {1, 2, 4})
        for (int halo : {0, 1}) {
          const int nacc = 2 * issuers;
          if (nacc * n > 512) continue;
          MmaArgs a{m, n, nacc, issuers, 144, d_out, halo};
          for (int rep = 0; rep < 2; ++rep) mma_cost_kernel<<<1, 128, 100 * 1024>>>(a);
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(h, d_out, 8 * sizeof(long long), cudaMemcpyDeviceToHost));
          for (int w = 1; w < issuers; ++w) if (h[w * 2 + 1] > h[1]) h[1] = h[w * 2 + 1];
          printf("M=%3d N=%3d issuers=%d nacc=%d halo=%d : issue %6.1f  complete %6.1f cyc/MMA (total %lld)\n", m, n, issuers, nacc, halo,
                 (double)h[0] / (144 / issuers), (double)h[1] / 144, h[1]);
        }
  // same with all SMs busy (148 CTAs) for one configuration: is there any chip-level effect
  {
    MmaArgs a{128, 16, 8, 2, 144, d_out, 1};
    for (int rep = 0; rep < 2; ++rep) mma_cost_kernel<<<148, 128, 100 * 1024>>>(a);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d_out, 4 * sizeof(long long), cudaMemcpyDeviceToHost));
    printf("148 CTAs: M=128 N=16 issuers=2 nacc=8 : complete %6.1f cyc/MMA\n", (double)h[1] / 144);
  }

  // B
  printf("\nDSMEM push + cluster barrier (cycles per round, every CTA pushes to all peers)\n");
  for (int csz : {8}) {
    for (int mode : {0})
      for (int bytes : {1024, 4096, 8192, 16384}) {
        if (bytes * csz > 128 * 1024) continue;
        PushArgs a{bytes, mode, 20, d_out};
        void* params[] = {&a};
        launch_cluster(push_kernel, csz * (128 / csz), 256, 200 * 1024, csz, params);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d_out, 2 * sizeof(long long), cudaMemcpyDeviceToHost));
        const double out_bytes = (double)bytes * (csz - 1);
        printf("cluster=%d mode=%s bytes/peer=%5d out=%6.0f B : %6lld cyc/round -> %5.1f B/cyc/CTA\n", csz,
               mode ? "bulk" : "st.v4", bytes, out_bytes, h[0], out_bytes / (double)h[0]);
      }
  }
  // C
  printf("\ncluster barrier (arrive.release + wait.acquire), 256 threads\n");
  for (int csz : {2, 4, 8}) {
    int iters = 100;
    void* params[] = {&iters, &d_out};
    launch_cluster(csync_kernel, csz * (128 / csz), 256, 0, csz, params);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d_out, sizeof(long long), cudaMemcpyDeviceToHost));
    printf("cluster=%d : %lld cyc/sync\n", csz, h[0]);
  }
  return 0;
}
