// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05) for 8x8 boards.
//
// Replaces, per layer, what Keras/TF ran for
//   Conv2D(k x k, padding="same", use_bias=False) + BatchNormalization (+ Add) + ReLU
// in ChessModel.build / _build_residual_block (reference src/chess_zero/agent/model_chess.py:65-69,96-111).
//
// GEMM view:  D[m, oc] = sum_{tap, ic} A_tap[m, ic] * W[oc, tap*Cin + ic]
//   m   = position*64 + rank*8 + file            (M = 64 * batch, 128 rows = 2 boards per tile)
//   A_tap[m, ic] = x[position, rank+dy-pad, file+dx-pad, ic], zero outside the board
//   activations are NHWC bf16, so A_tap for one (tap, 64-channel chunk) is a 4-D TMA box
//   {64 ch, 8 files, 8 ranks, 2 boards} fetched at signed coordinates (c0, dx-pad, dy-pad, n0): the TMA
//   unit's out-of-bounds zero fill IS the "same" padding, no halo is ever materialised.
//   The box lands as 128 rows x 128 bytes with the 128-byte swizzle = the canonical K-major UMMA layout.
//
// Warp roles (256 threads, persistent over M tiles):
//   warp 0   : TMA producer (one elected lane)      -> full[stage]
//   warp 1   : tcgen05.mma issuer (one elected lane) -> empty[stage] (tcgen05.commit), tmem_full[acc]
//   warp 2   : TMEM allocator / deallocator
//   warps 4-7: epilogue: tcgen05.ld accumulators, + folded-BN bias, + fp32 residual, ReLU,
//              fp32 store of the residual stream and/or bf16 store of the next layer's operand
// Two TMEM accumulator stages (2 x 256 fp32 columns) let the epilogue of tile i overlap the MMAs of i+1.
#pragma once
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace caz {
namespace tc {

constexpr int BLOCK_M = 128;  // rows per tile: two boards
constexpr int BLOCK_K = 64;   // bf16 elements per 128-byte swizzled row
constexpr int UMMA_K = 16;    // K per tcgen05.mma, 16-bit operands
constexpr int STAGES = 4;
constexpr int THREADS = 256;
constexpr int EPI_WARP0 = 4;
constexpr int ACC_STAGES = 2;
constexpr int ACC_COLS = 256;  // TMEM columns per accumulator stage (N <= 256)
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int B_STAGE_BYTES = 256 * BLOCK_K * 2;      // 32 KiB (N = 256 rows)
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

struct ConvArgs {
  int batch;        // positions
  int num_tiles;    // ceil(batch / 2)
  int taps_w;       // kernel width  (3 or 5)
  int taps;         // kernel height * width
  int pad;          // (k-1)/2
  int kchunks;      // Cin_padded / 64
  int n;            // Cout, multiple of 16, <= 256
  int relu;
  const float* bias;        // [n] folded BatchNorm shift
  const float* residual;    // nullptr or fp32 [batch*64][n]: the residual stream stays in fp32
  float* out_f32;           // nullptr or fp32 [batch*64][n] (may alias residual: same element, same thread)
  __nv_bfloat16* out_bf16;  // nullptr or bf16 [batch*64][n]: the next convolution's A operand
  int* err_flag;
};

enum : int { ERR_PRODUCER = 101, ERR_MMA_FULL = 102, ERR_MMA_ACC = 103, ERR_EPILOGUE = 104 };

__global__ void __launch_bounds__(THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_w,
               const ConvArgs args) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                            // [STAGES]
  uint64_t* empty = bars + STAGES;                  // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;          // [ACC_STAGES]
  uint64_t* tmem_empty = bars + 2 * STAGES + ACC_STAGES;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_k_blocks = args.taps * args.kchunks;
  const uint32_t b_bytes = static_cast<uint32_t>(args.n) * BLOCK_K * 2;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_act);
    ptx::prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    ptx::mbar_fence_init();
    ptx::fence_proxy_async();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_base_slot, ACC_STAGES * ACC_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) {
        const int n0 = tile * 2;
        for (int tap = 0; tap < args.taps; ++tap) {
          const int dy = tap / args.taps_w - args.pad;
          const int dx = tap % args.taps_w - args.pad;
          for (int kc = 0; kc < args.kchunks; ++kc) {
            ptx::mbar_wait(&empty[stage], phase ^ 1, args.err_flag, ERR_PRODUCER);
            ptx::mbar_expect_tx(&full[stage], A_STAGE_BYTES + b_bytes);
            ptx::tma_load_4d(smem_a + stage * A_STAGE_BYTES, &tmap_act, &full[stage], kc * BLOCK_K, dx, dy, n0);
            ptx::tma_load_2d(smem_b + stage * B_STAGE_BYTES, &tmap_w, &full[stage],
                             (tap * args.kchunks + kc) * BLOCK_K, 0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = ptx::idesc_bf16_f32(BLOCK_M, args.n);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1, args.err_flag, ERR_MMA_ACC);
        ptx::tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          ptx::mbar_wait(&full[stage], phase, args.err_flag, ERR_MMA_FULL);
          ptx::tcgen05_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem_a + stage * A_STAGE_BYTES);
          const uint32_t b_addr = ptx::smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adesc = ptx::smem_desc_sw128(a_addr + k * UMMA_K * 2);
            const uint64_t bdesc = ptx::smem_desc_sw128(b_addr + k * UMMA_K * 2);
            ptx::mma_bf16_ss(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::mma_commit(&empty[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    int acc = 0;
    uint32_t acc_phase = 0;
    const long total_rows = static_cast<long>(args.batch) * 64;
    for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tmem_full[acc], acc_phase, args.err_flag, ERR_EPILOGUE);
      ptx::tcgen05_fence_after();
      const long row = static_cast<long>(tile) * BLOCK_M + quad * 32 + lane;
      const bool live = row < total_rows;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * ACC_COLS;
      for (int c0 = 0; c0 < args.n; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(taddr + c0, v);
        ptx::tmem_ld_wait();
        if (live) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + __ldg(args.bias + c0 + j);
          if (args.residual) {
            const float4* rp = reinterpret_cast<const float4*>(args.residual + row * args.n + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 rv = rp[q];
              f[q * 4 + 0] += rv.x;
              f[q * 4 + 1] += rv.y;
              f[q * 4 + 2] += rv.z;
              f[q * 4 + 3] += rv.w;
            }
          }
          if (args.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
          }
          if (args.out_f32) {
            float4* op = reinterpret_cast<float4*>(args.out_f32 + row * args.n + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) op[q] = make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
          }
          if (args.out_bf16) {
            uint4* op = reinterpret_cast<uint4*>(args.out_bf16 + row * args.n + c0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 b2 = __floats2bfloat162_rn(f[q * 8 + e * 2], f[q * 8 + e * 2 + 1]);
                w[e] = *reinterpret_cast<const uint32_t*>(&b2);
              }
              op[q] = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
      }
      ptx::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, ACC_STAGES * ACC_COLS);
  }
}

}  // namespace tc
}  // namespace caz
