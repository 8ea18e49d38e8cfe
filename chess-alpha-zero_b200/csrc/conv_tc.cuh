// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05) for 8x8 boards.
//
// Replaces, per layer, what Keras/TF ran for
//   Conv2D(k x k, padding="same", use_bias=False) + BatchNormalization (+ Add) + ReLU
// in ChessModel.build / _build_residual_block (reference src/chess_zero/agent/model_chess.py:65-69,96-111).
//
// GEMM view:  D[m, oc] = sum_{tap, ic} A_tap[m, ic] * W[oc, tap*Cin + ic],  one tile = 128 rows = two boards.
//
// What bounds this kernel is NOT the tensor pipe but the bytes each SM pulls from L2 per tile: a B200 SM sustains
// ~100-110 GB/s of TMA fill with all 148 SMs pulling (measured: V1 fetched 1.77 MB per tile -- nine shifted copies
// of the activations plus the layer's weights -- and took 17.8 us per tile against 9.6 us of MMA time).  So:
//   * activations: ONE zero-padded halo tile per 64-channel chunk, [10 ranks][2 boards][10 files][64 ch] = 25.6 KB,
//     fetched by a single 4-D TMA box at signed coordinates (c0, -pad, n0, -pad) -- the out-of-bounds zero fill
//     is the convolution's "same" padding -- and shared by all k*k taps: tap (dy, dx) is the same tile read through
//     a UMMA descriptor whose start address is shifted by (2*hw*dy + dx) rows of 128 bytes and whose 8-row-group
//     stride is hw rows.  The 128-byte swizzle is a function of the shared-memory address, so the shifted view is
//     consistent with what TMA wrote.  102 KB of activations per tile instead of 590 KB.
//   * the K loop runs chunk-major (kc outer, taps inner); a chunk tile lives for k*k K-blocks (~2.7 us), far longer
//     than a TMA round trip, so the activation ring is just two chunk tiles.
//   * weights: [oc][tap*Cin + ic] bf16, one 32 KB TMA box per (tap, chunk).  With the activations nearly free, what
//     bounds a tile is the weight ring's DEPTH: each K block needs 32 KB that take ~1.2 us to arrive under load, so
//     K blocks complete at (stages / latency); every byte of shared memory the halo trick freed goes to that ring.
// TMEM lane m of the accumulator is row (rank h = m>>4, board p = (m>>3)&1, file w = m&7).
//
// Warp roles (256 threads, persistent over tiles; single-thread roles are warp-uniform loops with only the
// instruction issue under elect.sync -- issuing from an `if (lane == 0)` region makes the compiler wrap every
// UTCHMMA / UTMALDG in an election loop):
//   warp 0   : weight (B) TMA producer
//   warp 3   : activation (A) halo-tile TMA producer
//   warp 1   : tcgen05.mma issuer; tcgen05.commit frees the weight stage / the chunk tile / hands over the accumulator
//   warp 2   : TMEM allocator / deallocator
//   warps 4-7: epilogue: tcgen05.ld accumulators, + folded-BN bias, + fp32 residual, ReLU, fp32 store of the
//              residual stream and/or bf16 store of the next layer's operand (or, on the last layer, the heads'
//              1x1 convolutions straight from the registers)
// Two TMEM accumulator stages (2 x 256 fp32 columns) let the epilogue of tile i overlap the MMAs of i+1.
#pragma once
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace caz {
namespace tc {

constexpr int BLOCK_M = 128;  // rows per tile: two boards
constexpr int BLOCK_K = 64;   // bf16 elements per 128-byte swizzled row
constexpr int UMMA_K = 16;    // K per tcgen05.mma, 16-bit operands
constexpr int B_STAGES = 4;
constexpr int THREADS = 256;
constexpr int EPI_WARP0 = 4;
constexpr int ACC_STAGES = 2;
constexpr int ACC_COLS = 256;                          // TMEM columns per accumulator stage (N <= 256)
constexpr int A_RING_BYTES = 2 * (10 * 2 * 10 * 128);  // 51 200: two 3x3 chunk tiles (or one 5x5 / 7x7 one)
constexpr int MAX_A_SLOTS = 8;   // 3x3 and larger chunk tiles: two fit the ring; the 16 KB tiles of the GEMM uses (ksize 1): three, or up to
                                 // eight in the weight-gradient GEMM, which re-balances the two rings (conv_tc2.cuh)
constexpr int B_STAGE_BYTES = 256 * BLOCK_K * 2;       // 32 KiB (N = 256 rows)
constexpr int EPI_WARP_BYTES = 4096;                   // per-warp 32x32 fp32 transpose buffer
constexpr int HEAD_MAX_F = 8;                          // policy + value 1x1 filters fused into the last epilogue
constexpr int EPI_BYTES = 4 * EPI_WARP_BYTES + 1024 + HEAD_MAX_F * 256 * 4;  // + bias (<= 256 floats) + head filters
constexpr int SMEM_BYTES = A_RING_BYTES + B_STAGES * B_STAGE_BYTES + EPI_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

struct ConvArgs {
  int batch;        // positions
  int num_tiles;    // ceil(batch / 2)
  int ksize;        // kernel height = width (3 or 5)
  int kchunks;      // Cin_padded / (row_bytes / 2)
  int row_bytes;    // bytes of one operand row in shared memory = swizzle span: 128 (64 channels per K block) or
                    // 64 (32 channels: the stem, whose 18 input planes would otherwise be padded to 64)
  int n;            // Cout, multiple of 32, <= 256
  int relu;
  const float* bias;        // [n] folded BatchNorm shift
  const float* residual;    // nullptr or the fp32 residual stream, epilogue-native layout (see xf_index)
  float* out_f32;           // nullptr or the same stream to write (may alias residual: same element, same thread)
  __nv_bfloat16* out_bf16;  // nullptr or bf16 [batch*64][n]: the next convolution's A operand
  // Last layer of the tower only: the two heads' 1x1 convolutions + folded BN + ReLU + Flatten
  // (model_chess.py:77-81,86-90) are computed from the fp32 epilogue registers, so the final activation never
  // goes to HBM.  head_w: fp32 [pf+vf][n]; outputs in Keras channels-first Flatten order [b][c*64 + square].
  const float* head_w;      // nullptr = not fused
  const float* head_b;      // [pf+vf]
  float* featp;             // [batch][pf*64]
  float* featv;             // [batch][vf*64]
  int pf, vf;
  int fp16;                 // operand formats: ptx::FMT_A_FP16 (activations) | ptx::FMT_B_FP16 (weights); 0 = both bf16
  // Plain-GEMM use of the same kernel (the heads' Dense layers: ksize 1, rows = positions in groups of 64):
  // the N dimension is walked in n_tiles tiles of `n` columns, the result goes to a ROW-MAJOR fp32 matrix.
  int n_tiles;              // 1 for convolutions
  int n_total;              // valid output columns (<= n_tiles * n)
  int m_rows;               // valid output rows
  float* out_rm;            // nullptr or fp32 [m_rows][n_total] row-major; bias is then [n_tiles * n]
  int reverse;              // walk the tiles from the last to the first (CTA-pair kernel; see conv_tc2.cuh)
  int l2_prefetch;          // pull the next tile's activations / residual into L2 one tile ahead
  int pdl;                  // launched with programmatic stream serialization (CTA-pair kernel, see conv_tc2.cuh)
  // Weight-gradient use of the CTA-pair kernel (training, caz_train.inc.cu): out[tap][split] = A[:, K range] . B[:, K range + shift(tap)]^T
  // with A = dz^T [n_out = 256 rows][K] and B = x^T [n_in rows][K], both over the zero-padded board index K = board * pitch^2 +
  // (rank + pad) * pitch + file + pad, so that a filter tap is a plain shift of B's K coordinate.  wg_splits > 0 switches it on:
  // work item t = (tap t / wg_splits, split t % wg_splits) covers K chunks [split * kchunks, (split + 1) * kchunks).
  int wg_splits, wg_k, wg_pitch, wg_dx;   // K shift of tap (dy, dx): (dy - pad) * wg_pitch + (dx - pad) * wg_dx
  // wg_cpc > 0: K runs over the 64 REAL cells only (wg_cpc chunks of 64 boards per cell): chunk j of the work list is chunk
  // (cell(j / wg_cpc) * wg_cpc + j % wg_cpc) of the padded index -- A (dz^T) is zero on the halo cells, so those chunks (36 of
  // 100 for a 3x3 kernel, 80 of 144 for 5x5) need neither loading nor multiplying
  int wg_cpc;
  int mixed_whole;          // CTA-pair kernel, n_tiles == 2: this many leading work items take both N tiles at once (conv_tc2.cuh)
  // Training (BatchNormalization batch statistics out of the convolution's epilogue, with out_rm): nullptr, or
  // [2][n_total][stats_slots] floats -- every epilogue warp leaves the sum and the sum of squares of its 32 rows of each of
  // its output columns in slot (CTA tile * 4 + warp); rows of boards beyond the batch are left out.  Each
  // (column, slot) is written exactly once per launch; bn_finish_stats_f32_kernel adds the slots in a fixed order.
  float* stats;
  int stats_slots;
  int all_lanes_fence;      // one-launch tower, A/B switch (CAZ_TOWER_ALL_LANES_FENCE): every lane fences before a readiness arrive
  int* err_flag;
};

enum : int { ERR_PRODUCER_B = 101, ERR_PRODUCER_A = 105, ERR_MMA_FULL = 102, ERR_MMA_ACC = 103, ERR_EPILOGUE = 104, ERR_MMA_A = 106 };

// K-major swizzled descriptor with an explicit 8-row-group stride (the halo view needs hw*row_bytes, weights
// 8*row_bytes).  row_bytes 128 -> SWIZZLE_128B (layout type 2), 64 -> SWIZZLE_64B (layout type 4).
__device__ __forceinline__ uint64_t desc_sw(uint32_t smem_addr, uint32_t sbo_bytes, int row_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(row_bytes == 128 ? 2 : 4) << 61;
  return d;
}

constexpr uint32_t kLocalCta = 0xffffffffu;

// The fp32 residual stream is private to these epilogues (written by one, read by the next, by the SAME thread of the
// SAME tile), so it is not stored row-major but in the order the epilogue touches it:
//   xf[tile][32-column chunk][float4 q of the chunk (8)][TMEM lane m (128)]   (float4 units)
// A warp's load/store of one float4 per thread is then 512 contiguous bytes -- no shared-memory transposition, which
// matters because the tensor cores' operand reads already use most of the shared-memory bandwidth.
__device__ __forceinline__ size_t xf_index(int tile, int chunks, int chunk, int q, int m) {
  return ((static_cast<size_t>(tile) * chunks + chunk) * 8 + q) * 128 + m;
}

// Epilogue of ONE tile by one of the four epilogue warps (shared by the 1-CTA and the 2-CTA kernel).
//   n0          first board of the tile (boards n0, n0+1; a board >= batch does not exist)
//   taddr       TMEM address of this warp's lane quadrant in the accumulator stage
//   full_bar    mbarrier the MMA commit arrives on when the accumulator is complete
//   empty_bar   mbarrier to arrive on once the accumulator has been read out; in CTA `release_rank` of the cluster,
//               or in this CTA when release_rank == kLocalCta
// Warp `quad` owns TMEM lanes 32*quad .. +31 = ranks 2*quad, 2*quad+1 of both boards: local row r is
// (rank 2*quad + (r>>4), board (r>>3)&1, file r&7), eight consecutive r = eight consecutive global rows.
__device__ __forceinline__ void epilogue_tile(const ConvArgs& args, uint8_t* smem_epi, const float* s_bias,
                                              const float* s_head, int quad, int lane, int n0, uint32_t taddr,
                                              uint64_t* full_bar, uint32_t full_phase, uint64_t* empty_bar,
                                              uint32_t release_rank, int nt = 0, int next_n0 = -1, int n0_store = -1,
                                              uint64_t* chunk_ready = nullptr) {
  // chunk_ready: nullptr, or mbarriers [n / 64] of THIS CTA on which the warp arrives as soon as it has stored and fenced
  // 64 more output channels (the one-launch tower starts the next layer on them while the rest is still being drained)
  // n0_store: first board of the tile in the ACTIVATION buffers when they are addressed by something else than the board
  // number (the one-launch tower keeps them in a compact per-CTA scratch, conv_tower.cuh); -1 = n0
  if (n0_store < 0) n0_store = n0;
  uint4* stage_b = reinterpret_cast<uint4*>(smem_epi + quad * EPI_WARP_BYTES);
  float4* stage_f = reinterpret_cast<float4*>(stage_b);
  const int fr = lane >> 3, fq = lane & 7;   // fp32 transposed mapping: row i*4+fr, float4 column fq
  const int br = lane >> 2, bc = lane & 3;   // 16-bit transposed mapping: row i*8+br, 16-byte column bc
  const int left = args.batch - n0;
  const int boards_live = left >= 2 ? 2 : (left > 0 ? left : 0);
  // global (NHWC) row of local row r, and whether its board exists
  auto grow = [&](int r) -> long {
    return (static_cast<long>(n0_store + ((r >> 3) & 1)) << 6) + ((2 * quad + (r >> 4)) << 3) + (r & 7);
  };
  auto live = [&](int r) -> bool { return ((r >> 3) & 1) < boards_live; };
  // (two N tiles per board group, nt = 0 / 1: this epilogue owns 32-column chunks [chunk0, chunk0 + n / 32) of the row's
  // n_total / 32; the residual stream and the 16-bit operand keep their full-width layouts)
  const int tile = n0_store >> 1, chunks = (args.n_tiles > 1 ? args.n_total : args.n) >> 5, chunk0 = nt * (args.n >> 5);
  const int m = quad * 32 + lane, opitch = args.n_tiles > 1 ? args.n_total : args.n;
  const float4* res4 = reinterpret_cast<const float4*>(args.residual);
  float4* out4 = reinterpret_cast<float4*>(args.out_f32);
  float4 res_pf[8];
  if (args.residual) {
#pragma unroll
    for (int q = 0; q < 8; ++q) res_pf[q] = res4[xf_index(tile, chunks, chunk0, q, m)];
    // The residual stream comes from HBM (it is larger than L2 at big batches) and is consumed 4 KiB per warp and chunk
    // with only one chunk of lookahead in registers: pull the NEXT tile's slice into L2 now, a whole tile ahead.
    // A warp's slice of one chunk is 8 x 512 contiguous bytes = 32 lines: one prefetch per lane and chunk.
    if (next_n0 >= 0 && args.l2_prefetch) {
      const int ntile = next_n0 >> 1;
      for (int ch = 0; ch < (args.n >> 5); ++ch)
        ptx::prefetch_l2(res4 + xf_index(ntile, chunks, chunk0 + ch, lane >> 2, quad * 32 + (lane & 3) * 8));
    }
  }
  ptx::mbar_wait(full_bar, full_phase, args.err_flag, ERR_EPILOGUE);
  ptx::tcgen05_fence_after();
  float hacc[HEAD_MAX_F];
#pragma unroll
  for (int hf = 0; hf < HEAD_MAX_F; ++hf) hacc[hf] = 0.0f;
  for (int c0 = 0; c0 < args.n; c0 += 32) {
    const int chunk = chunk0 + (c0 >> 5);
    uint32_t v[32];
    ptx::tmem_ld_32x32(taddr + c0, v);
    ptx::tmem_ld_wait();
    if (c0 + 32 >= args.n) {
      // all accumulator columns of this tile are in registers: hand the TMEM stage back to the MMA warp
      ptx::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (release_rank == kLocalCta) ptx::mbar_arrive(empty_bar);
        else ptx::mbar_arrive_cluster(empty_bar, release_rank);
      }
    }
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j)
      f[j] = __uint_as_float(v[j]) + (args.n_tiles > 1 ? __ldg(args.bias + nt * args.n + c0 + j) : s_bias[c0 + j]);
    if (args.residual) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        f[q * 4 + 0] += res_pf[q].x;
        f[q * 4 + 1] += res_pf[q].y;
        f[q * 4 + 2] += res_pf[q].z;
        f[q * 4 + 3] += res_pf[q].w;
      }
      if (c0 + 32 < args.n) {  // next chunk's residual, in flight during this chunk's math and stores
#pragma unroll
        for (int q = 0; q < 8; ++q) res_pf[q] = res4[xf_index(tile, chunks, chunk + 1, q, m)];
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
#pragma unroll
      for (int q = 0; q < 8; ++q)
        out4[xf_index(tile, chunks, chunk, q, m)] = make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
    }
    if (args.out_rm) {
      // row-major fp32 result (GEMM use): transpose the 32x32 block through the per-warp staging buffer so that each
      // store instruction covers 4 rows x 128 contiguous bytes
#pragma unroll
      for (int q = 0; q < 8; ++q)
        stage_f[lane * 8 + (q ^ (lane & 7))] = make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
      __syncwarp();
      if (args.stats && tile * 4 + quad < args.stats_slots) {   // column c0 + lane of this warp's 32 rows (a CTA tile past the batch has no slot)
        const float* sf = reinterpret_cast<const float*>(stage_f);
        float s = 0.0f, q2 = 0.0f;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          // (a board past the batch reads whatever an earlier, larger batch left in the operand buffer: not part of the sums)
          const float x = live(r) ? sf[(r * 8 + ((lane >> 2) ^ (r & 7))) * 4 + (lane & 3)] : 0.0f;
          s += x;
          q2 = fmaf(x, x, q2);
        }
        const size_t at = static_cast<size_t>(nt * args.n + c0 + lane) * args.stats_slots + tile * 4 + quad;
        args.stats[at] = s;
        args.stats[static_cast<size_t>(args.n_total) * args.stats_slots + at] = q2;
      }
      const int col = nt * args.n + c0 + fq * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + fr;
        const float4 o = stage_f[r * 8 + (fq ^ (r & 7))];
        const long row = grow(r);
        if (row < args.m_rows && col + 3 < args.n_total)
          *reinterpret_cast<float4*>(args.out_rm + row * args.n_total + col) = o;
        else if (row < args.m_rows) {
          const float ov[4] = {o.x, o.y, o.z, o.w};
          for (int j = 0; j < 4; ++j)
            if (col + j < args.n_total) args.out_rm[row * args.n_total + col + j] = ov[j];
        }
      }
      __syncwarp();
    }
    if (args.out_bf16) {
      // the 16-bit operand copy must be NHWC for the next layer's TMA: transpose the 32x32 block through a 2 KiB
      // XOR-swizzled per-warp staging buffer so that each store instruction covers 8 rows x 64 contiguous bytes
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) w[e] = ptx::pack_h16x2(f[c * 8 + e * 2], f[c * 8 + e * 2 + 1], (args.fp16 & ptx::FMT_A_FP16) != 0);
        stage_b[lane * 4 + (c ^ ((lane >> 1) & 3))] = make_uint4(w[0], w[1], w[2], w[3]);
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i * 8 + br;
        const uint4 o = stage_b[r * 4 + (bc ^ ((r >> 1) & 3))];
        if (live(r)) *reinterpret_cast<uint4*>(args.out_bf16 + grow(r) * opitch + nt * args.n + c0 + bc * 8) = o;
      }
      __syncwarp();
    }
    if (chunk_ready && ((c0 + 32) & 63) == 0) {
      // channels [c0 - 32, c0 + 32) of this warp's rows are stored: visible to the TMA reads of the same CTA (async proxy)
      // (one gpu-scope fence by the arriving lane, cumulative over the warp's stores through __syncwarp -- the TMA reads them
      // from L2; every lane fencing twice cost 0.85 us per fence point instead of 0.3, measured in conv_sb.cuh)
      if (args.all_lanes_fence) { __threadfence(); ptx::fence_proxy_async_all(); }
      __syncwarp();
      if (lane == 0) {
        if (!args.all_lanes_fence) ptx::fence_acq_rel_gpu();
        ptx::mbar_arrive(&chunk_ready[c0 >> 6]);
      }
    }
  }
  if (args.head_w && live(lane)) {
    // this thread's row is one square of one board
    const long b = n0 + ((lane >> 3) & 1);
    const int sq = ((2 * quad + (lane >> 4)) << 3) + (lane & 7);
#pragma unroll
    for (int hf = 0; hf < HEAD_MAX_F; ++hf) {
      if (hf < args.pf + args.vf) {
        const float y = fmaxf(hacc[hf] + args.head_b[hf], 0.0f);
        if (hf < args.pf) args.featp[b * (args.pf * 64) + hf * 64 + sq] = y;
        else args.featv[b * (args.vf * 64) + (hf - args.pf) * 64 + sq] = y;
      }
    }
  }
}

// tmap_halo: NHWC bf16 activations as (channel, file, position, rank) with box 64 x hw x 2 x hw (see caz_abi.cu)
__global__ void __launch_bounds__(THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_halo, const __grid_constant__ CUtensorMap tmap_w,
               const ConvArgs args) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + A_RING_BYTES;
  uint8_t* smem_epi = smem_b + B_STAGES * B_STAGE_BYTES;
  float* s_bias = reinterpret_cast<float*>(smem_epi + 4 * EPI_WARP_BYTES);
  float* s_head = s_bias + 256;  // [pf+vf][n]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + EPI_BYTES);
  uint64_t* a_full = bars;                          // [MAX_A_SLOTS]
  uint64_t* a_empty = a_full + MAX_A_SLOTS;         // [MAX_A_SLOTS]
  uint64_t* b_full = a_empty + MAX_A_SLOTS;         // [B_STAGES]
  uint64_t* b_empty = b_full + B_STAGES;            // [B_STAGES]
  uint64_t* tmem_full = b_empty + B_STAGES;         // [ACC_STAGES]
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;    // [ACC_STAGES]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int k = args.ksize, pad = (k - 1) / 2, hw = 8 + k - 1;
  const int rb = args.row_bytes, kblk = rb >> 1;        // operand row bytes, K elements per K block
  const int plane_bytes = hw * 2 * hw * rb;            // one chunk tile; a multiple of 1024
  const int a_slots = A_RING_BYTES / plane_bytes < MAX_A_SLOTS ? A_RING_BYTES / plane_bytes : MAX_A_SLOTS;
  const uint32_t b_bytes = static_cast<uint32_t>(args.n) * rb;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_halo);
    ptx::prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < a_slots; ++i) {
      ptx::mbar_init(&a_full[i], 1);
      ptx::mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < B_STAGES; ++i) {
      ptx::mbar_init(&b_full[i], 1);
      ptx::mbar_init(&b_empty[i], 1);
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

  if (warp == 0) {
    // ===================== weight producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x)
      for (int kc = 0; kc < args.kchunks; ++kc)
        for (int tap = 0; tap < k * k; ++tap) {
          ptx::mbar_wait(&b_empty[stage], phase ^ 1, args.err_flag, ERR_PRODUCER_B);
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(&b_full[stage], b_bytes);
            ptx::tma_load_2d(smem_b + stage * B_STAGE_BYTES, &tmap_w, &b_full[stage], (tap * args.kchunks + kc) * kblk, 0);
          }
          __syncwarp();
          if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
        }
  } else if (warp == 3) {
    // ===================== activation producer: one halo tile per (board pair, channel chunk) =====================
    int slot = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x)
      for (int kc = 0; kc < args.kchunks; ++kc) {
        ptx::mbar_wait(&a_empty[slot], phase ^ 1, args.err_flag, ERR_PRODUCER_A);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&a_full[slot], plane_bytes);
          ptx::tma_load_4d(smem_a + slot * plane_bytes, &tmap_halo, &a_full[slot], kc * kblk, -pad, tile * 2, -pad);
        }
        __syncwarp();
        if (++slot == a_slots) { slot = 0; phase ^= 1; }
      }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = ptx::idesc_h16_f32(BLOCK_M, args.n, args.fp16);
    // descriptors differ only in the 14-bit start-address field: add (byte offset >> 4) to a template
    const uint64_t a_desc0 = desc_sw(ptx::smem_u32(smem_a), hw * rb, rb);  // next 8-row group = next (rank, board)
    const uint64_t b_desc0 = desc_sw(ptx::smem_u32(smem_b), 8 * rb, rb);
    const int tap_rows16 = rb >> 4;   // one halo row in descriptor units (16 bytes)
    int stage = 0, slot = 0, acc = 0;
    uint32_t phase = 0, a_phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1, args.err_flag, ERR_MMA_ACC);
      ptx::tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
      for (int kc = 0; kc < args.kchunks; ++kc) {
        ptx::mbar_wait(&a_full[slot], a_phase, args.err_flag, ERR_MMA_A);
        const uint64_t a_plane = a_desc0 + static_cast<uint64_t>((slot * plane_bytes) >> 4);
        for (int dy = 0; dy < k; ++dy)
          for (int dx = 0; dx < k; ++dx) {
            ptx::mbar_wait(&b_full[stage], phase, args.err_flag, ERR_MMA_FULL);
            ptx::tcgen05_fence_after();
            if (ptx::elect_one()) {
              // tap (dy, dx) = the same halo tile, viewed (2*hw*dy + dx) rows further on
              const uint64_t adesc = a_plane + static_cast<uint64_t>((dy * 2 * hw + dx) * tap_rows16);
              const uint64_t bdesc = b_desc0 + static_cast<uint64_t>((stage * B_STAGE_BYTES) >> 4);
              const uint32_t first = (kc | dy | dx) == 0 ? 0u : 1u;
              ptx::mma_bf16_ss(tmem_d, adesc, bdesc, idesc, first);
              ptx::mma_bf16_ss(tmem_d, adesc + 2, bdesc + 2, idesc, 1u);
              if (rb == 128) {
                ptx::mma_bf16_ss(tmem_d, adesc + 4, bdesc + 4, idesc, 1u);
                ptx::mma_bf16_ss(tmem_d, adesc + 6, bdesc + 6, idesc, 1u);
              }
              ptx::mma_commit(&b_empty[stage]);  // weight stage reusable once these MMAs have read it
              if (dy == k - 1 && dx == k - 1) {
                ptx::mma_commit(&a_empty[slot]);                                  // chunk tile fully consumed
                if (kc == args.kchunks - 1) ptx::mma_commit(&tmem_full[acc]);     // accumulator complete -> epilogue
              }
            }
            __syncwarp();
            if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
          }
        if (++slot == a_slots) { slot = 0; a_phase ^= 1; }
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue =====================
    // TMEM hands thread t the tile row in lane t (tcgen05.ld 32x32b), but a row is 1 KiB (fp32) / 512 B (bf16) of
    // global memory, so row-per-thread global accesses touch 32 lines per instruction.  Every 32x32 block is
    // therefore transposed through a 4 KiB per-warp staging buffer (XOR-swizzled, conflict-free both ways) so
    // that each global load/store instruction covers whole 128-byte lines: 4 rows x 128 B (fp32) or
    // 8 rows x 64 B (bf16) per instruction.  The fp32 residual chunk is prefetched one chunk ahead.
    // Warp `quad` owns TMEM lanes 32*quad .. +31 = ranks 2*quad, 2*quad+1 of both boards: local row r is
    // (rank 2*quad + (r>>4), board (r>>3)&1, file r&7), eight consecutive r = eight consecutive global rows.
    const int quad = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) {
      const int n0 = tile * 2;
      const int next_tile = tile + static_cast<int>(gridDim.x);
      epilogue_tile(args, smem_epi, s_bias, s_head, quad, lane, n0,
                    tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * ACC_COLS, &tmem_full[acc], acc_phase,
                    &tmem_empty[acc], kLocalCta, 0, next_tile < args.num_tiles ? next_tile * 2 : -1);
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
