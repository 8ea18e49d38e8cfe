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
constexpr int EPI_WARP_BYTES = 4096;                  // per-warp 32x32 fp32 transpose buffer
constexpr int HEAD_MAX_F = 8;                          // policy + value 1x1 filters fused into the last epilogue
constexpr int EPI_BYTES = 4 * EPI_WARP_BYTES + 1024 + HEAD_MAX_F * 256 * 4;  // + bias (<= 256 floats) + head filters
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

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
  // Last layer of the tower only: the two heads' 1x1 convolutions + folded BN + ReLU + Flatten
  // (model_chess.py:77-81,86-90) are computed from the fp32 epilogue registers, so the final activation never
  // goes to HBM.  head_w: fp32 [pf+vf][n]; outputs in Keras channels-first Flatten order [b][c*64 + square].
  const float* head_w;      // nullptr = not fused
  const float* head_b;      // [pf+vf]
  float* featp;             // [batch][pf*64]
  float* featv;             // [batch][vf*64]
  int pf, vf;
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
  uint8_t* smem_epi = smem + STAGES * STAGE_BYTES;
  float* s_bias = reinterpret_cast<float*>(smem_epi + 4 * EPI_WARP_BYTES);
  float* s_head = s_bias + 256;  // [pf+vf][n]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + EPI_BYTES);
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
  if (threadIdx.x < args.n) s_bias[threadIdx.x] = args.bias[threadIdx.x];
  if (args.head_w)
    for (int i = threadIdx.x; i < (args.pf + args.vf) * args.n; i += THREADS) s_head[i] = args.head_w[i];
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  // The single-thread roles run as WARP-UNIFORM loops: all 32 lanes wait on the mbarriers and walk the loop
  // counters, and only the instruction issue sits under elect.sync.  (Issuing from inside an `if (lane == 0)` region
  // makes the compiler wrap every UTCHMMA / UTMALDG in an election loop with R2UR moves: ~700 issue cycles per K
  // block, more than the 512 cycles the four MMAs take.)
  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) {
      const int n0 = tile * 2;
      int kcol = 0;  // K coordinate of the weight block: (tap * kchunks + kc) * 64
      for (int dy = -args.pad; dy <= args.pad; ++dy)
        for (int dx = -args.pad; dx <= args.pad; ++dx)
          for (int kc = 0; kc < args.kchunks; ++kc, kcol += BLOCK_K) {
            ptx::mbar_wait(&empty[stage], phase ^ 1, args.err_flag, ERR_PRODUCER);
            if (ptx::elect_one()) {
              ptx::mbar_expect_tx(&full[stage], A_STAGE_BYTES + b_bytes);
              ptx::tma_load_4d(smem_a + stage * A_STAGE_BYTES, &tmap_act, &full[stage], kc * BLOCK_K, dx, dy, n0);
              ptx::tma_load_2d(smem_b + stage * B_STAGE_BYTES, &tmap_w, &full[stage], kcol, 0);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = ptx::idesc_bf16_f32(BLOCK_M, args.n);
    // descriptors differ only in the 14-bit start-address field: add (byte offset >> 4) to a template
    const uint64_t a_desc0 = ptx::smem_desc_sw128(ptx::smem_u32(smem_a));
    const uint64_t b_desc0 = ptx::smem_desc_sw128(ptx::smem_u32(smem_b));
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
        if (ptx::elect_one()) {
          const uint64_t adesc = a_desc0 + static_cast<uint64_t>((stage * A_STAGE_BYTES) >> 4);
          const uint64_t bdesc = b_desc0 + static_cast<uint64_t>((stage * B_STAGE_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            ptx::mma_bf16_ss(tmem_d, adesc + k * ((UMMA_K * 2) >> 4), bdesc + k * ((UMMA_K * 2) >> 4), idesc,
                             (kb | k) != 0 ? 1u : 0u);
          ptx::mma_commit(&empty[stage]);                             // smem slot reusable once these MMAs have read it
          if (kb == num_k_blocks - 1) ptx::mma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue =====================
    // TMEM hands thread t row t of the tile (tcgen05.ld 32x32b), but a row is 1 KiB (fp32) / 512 B (bf16) of
    // global memory, so row-per-thread global accesses touch 32 lines per instruction.  Every 32x32 block is
    // therefore transposed through a 4 KiB per-warp staging buffer (XOR-swizzled, conflict-free both ways) so
    // that each global load/store instruction covers whole 128-byte lines: 4 rows x 128 B (fp32) or
    // 8 rows x 64 B (bf16) per instruction.  The fp32 residual chunk is prefetched one chunk ahead.
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    float4* stage = reinterpret_cast<float4*>(smem_epi + quad * EPI_WARP_BYTES);
    uint4* stage_b = reinterpret_cast<uint4*>(stage);
    const int fr = lane >> 3, fq = lane & 7;   // fp32 transposed mapping: row i*4+fr, float4 column fq
    const int br = lane >> 2, bc = lane & 3;   // bf16 transposed mapping: row i*8+br, 16-byte column bc
    int acc = 0;
    uint32_t acc_phase = 0;
    const long total_rows = static_cast<long>(args.batch) * 64;
    for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) {
      const long row0 = static_cast<long>(tile) * BLOCK_M + quad * 32;
      const long left = total_rows - row0;
      const int rows_live = left >= 32 ? 32 : (left > 0 ? static_cast<int>(left) : 0);
      float4 res_pf[8];
      if (args.residual) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + fr;
          res_pf[i] = r < rows_live ? *reinterpret_cast<const float4*>(args.residual + (row0 + r) * args.n + fq * 4)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      ptx::mbar_wait(&tmem_full[acc], acc_phase, args.err_flag, ERR_EPILOGUE);
      ptx::tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * ACC_COLS;
      float hacc[HEAD_MAX_F];
#pragma unroll
      for (int hf = 0; hf < HEAD_MAX_F; ++hf) hacc[hf] = 0.0f;
      for (int c0 = 0; c0 < args.n; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(taddr + c0, v);
        if (args.residual) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + fr;
            stage[r * 8 + (fq ^ (r & 7))] = res_pf[i];
          }
          __syncwarp();
          if (c0 + 32 < args.n) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = i * 4 + fr;
              if (r < rows_live)
                res_pf[i] = *reinterpret_cast<const float4*>(args.residual + (row0 + r) * args.n + c0 + 32 + fq * 4);
            }
          }
        }
        ptx::tmem_ld_wait();
        if (c0 + 32 >= args.n) {
          // all accumulator columns of this tile are in registers: hand the TMEM stage back to the MMA warp
          ptx::tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
        }
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + s_bias[c0 + j];
        if (args.residual) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 rv = stage[lane * 8 + (q ^ (lane & 7))];
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
        if (args.head_w) {
          const int hf_total = args.pf + args.vf;
#pragma unroll
          for (int hf = 0; hf < HEAD_MAX_F; ++hf) {
            if (hf < hf_total) {
              const float4* wp = reinterpret_cast<const float4*>(s_head + hf * args.n + c0);  // warp-uniform: broadcast
              float a = hacc[hf];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 wv = wp[q];
                a = fmaf(f[q * 4 + 0], wv.x, a);
                a = fmaf(f[q * 4 + 1], wv.y, a);
                a = fmaf(f[q * 4 + 2], wv.z, a);
                a = fmaf(f[q * 4 + 3], wv.w, a);
              }
              hacc[hf] = a;
            }
          }
        }
        if (args.out_f32) {
          // own row only: no cross-lane hazard with the residual reads above
#pragma unroll
          for (int q = 0; q < 8; ++q)
            stage[lane * 8 + (q ^ (lane & 7))] = make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + fr;
            const float4 o = stage[r * 8 + (fq ^ (r & 7))];
            if (r < rows_live) *reinterpret_cast<float4*>(args.out_f32 + (row0 + r) * args.n + c0 + fq * 4) = o;
          }
          __syncwarp();
        }
        if (args.out_bf16) {
          if (args.residual && !args.out_f32) __syncwarp();  // residual rows fully read before reuse
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __nv_bfloat162 b2 = __floats2bfloat162_rn(f[c * 8 + e * 2], f[c * 8 + e * 2 + 1]);
              w[e] = *reinterpret_cast<const uint32_t*>(&b2);
            }
            stage_b[lane * 4 + (c ^ ((lane >> 1) & 3))] = make_uint4(w[0], w[1], w[2], w[3]);
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = i * 8 + br;
            const uint4 o = stage_b[r * 4 + (bc ^ ((r >> 1) & 3))];
            if (r < rows_live) *reinterpret_cast<uint4*>(args.out_bf16 + (row0 + r) * args.n + c0 + bc * 8) = o;
          }
          __syncwarp();
        }
      }
      if (args.head_w && lane < rows_live) {
        // this thread's row is one square of one board: 32 consecutive lanes = 32 consecutive squares
        const long row = row0 + lane;
        const long b = row >> 6;
        const int sq = static_cast<int>(row & 63);
#pragma unroll
        for (int hf = 0; hf < HEAD_MAX_F; ++hf) {
          if (hf < args.pf + args.vf) {
            const float y = fmaxf(hacc[hf] + args.head_b[hf], 0.0f);
            if (hf < args.pf) args.featp[b * (args.pf * 64) + hf * 64 + sq] = y;
            else args.featv[b * (args.vf * 64) + (hf - args.pf) * 64 + sq] = y;
          }
        }
      }
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
