// The whole convolutional tower (stem + every residual block + the heads' 1x1 convolutions) in ONE persistent launch
// for batches beyond the small-batch kernel's reach.
//
// Replaces, for a batch, the chain  Conv2D+BN+ReLU, n x [Conv2D+BN+ReLU, Conv2D+BN+Add+ReLU], policy/value Conv2D(1x1)+BN+ReLU
// of ChessModel.build / _build_residual_block (reference src/chess_zero/agent/model_chess.py:65-72,77-81,86-90,96-111).
//
// Why one launch: a 3x3 convolution on 8x8 boards has no dependency ACROSS boards, and a CTA pair of the per-layer
// kernel (conv_tc2.cuh) already computes all 256 output channels of its four boards.  So a pair can run layer l+1 of its
// boards as soon as ITS OWN epilogue of layer l has been written -- no grid-wide step, no flags between CTAs.  Each CTA of
// the pair keeps two boards; everything it reads back (16-bit operand copies, fp32 residual stream) it wrote itself,
// tens of microseconds earlier: the working set of the whole grid is 148 x 2 groups x 256 KB = 76 MB and stays in the 126 MB L2
// together with the 17.4 MB of weights.  The per-layer chain streamed every layer's activations of the WHOLE batch through
// HBM instead (0.75 GB per conv2-type launch at 4096 boards, 2.8x the algorithmic bytes; tensor pipe 81 % busy against
// 92 % for conv1-type launches), and at a few hundred boards most of a launch was launch latency, prologue and tail.
//
// Work order of a pair: super-groups of EIGHT boards (group A = tiles 4s, 4s+1; group B = tiles 4s+2, 4s+3; rank r of the
// pair owns tile 2 * (2s + g) + r; a tile = two boards).  Layers are walked with the two groups interleaved -- (A,0) (B,0) (A,1)
// (B,1) ... -- so that while the epilogue warps drain layer l of group A, the tensor cores run layer l of group B, and
// the two TMEM accumulator stages map one to each group.  The only new synchronisation is CTA-local: the epilogue warps
// fence their global stores towards the async proxy and arrive on `act_ready[group][chunk]` after every 64 output channels;
// the activation producer waits for a chunk's barrier before it requests that 64-channel halo tile of the group's next
// layer with TMA -- so a lone group's next layer starts on the first channels while the rest is still being drained.
//
// Everything else (halo tile + shifted UMMA descriptors, cta_group::2 MMAs, weight halves per CTA, multicast commits,
// the fused epilogue) is conv_tc2.cuh / conv_tc.cuh.
#pragma once
#include <cuda_bf16.h>

#include "conv_tc.cuh"
#include "conv_tc2.cuh"
#include "ptx.cuh"

namespace caz {
namespace tw {

constexpr int THREADS = 256;
constexpr int B_STAGES = 8;
constexpr int B_STAGE_BYTES = 128 * 64 * 2;         // this CTA's 128 weight rows of one K block
constexpr int A_SLOT_BYTES = 10 * 2 * 10 * 128;     // one 3x3 chunk tile; the stem's tile (64-byte rows) must fit too
constexpr int A_SLOTS = 2;
constexpr int ACC_STAGES = 2;
constexpr int ACC_COLS = 256;
constexpr int EPI_BYTES = 4 * tc::EPI_WARP_BYTES + 2 * 1024 /* two bias slots */ + tc::HEAD_MAX_F * 256 * 4;
constexpr int SMEM_BYTES = A_SLOTS * A_SLOT_BYTES + B_STAGES * B_STAGE_BYTES + EPI_BYTES + 1024 /*alignment slack*/ + 512 /*barriers*/;

enum : int { ERRT_PRODUCER_B = 401, ERRT_PRODUCER_A = 402, ERRT_MMA_B = 403, ERRT_MMA_A = 404, ERRT_MMA_ACC = 405, ERRT_ACT_READY = 406 };

struct TowerArgs {
  int batch;             // boards
  int n_layers;          // 1 + 2 * residual blocks
  int stem_k;            // stem kernel size (odd, <= 7)
  int stem_row_bytes;    // bytes of one stem operand row (64: up to 32 input planes)
  int n;                 // filters (= output channels of every layer), multiple of 64, <= 256
  int fmt;               // operand formats (ptx::FMT_*)
  int pdl;               // launched with programmatic stream serialization: wait for the previous grid before the first read
  int all_lanes_fence;   // A/B switch (CAZ_TOWER_ALL_LANES_FENCE): the epilogue's readiness arrives behind two fences of every lane
  int interleave;        // groups per super-group: 2 (the groups alternate layer by layer: epilogue under the other group's
                         // MMAs) or 1 (few boards: every group gets its own CTA pair, the epilogue is exposed)
  const float* bias_all; // [n_layers][n] folded BatchNorm shifts
  float* xf;             // fp32 residual stream, epilogue-native layout (tc::xf_index)
  __nv_bfloat16* act0;   // 16-bit copy of the residual stream  = A operand of conv1      [boards][64][n]
  __nv_bfloat16* act1;   // 16-bit conv1 output                 = A operand of conv2
  const float* head_w;   // [pf + vf][n]
  const float* head_b;
  float* featp;
  float* featv;
  int pf, vf;
  int* err_flag;
  long long* trace;      // nullptr, or [n_layers][8] clock64 stamps of pair 0's leader CTA for its first super-group, group 0
                         // (caz_debug_tower_trace): 0 first 64 input channels ready, 1 tile requested, 2 first chunk landed, 3 MMAs issued,
                         // 4 accumulator complete, 5 epilogue done
};

// tmap_in: stem input halo view; tmap_x / tmap_y: halo views of act0 / act1; tmap_w_stem: [n][k*k*cin_pad], box (row/2) x n/2;
// tmap_w_tower: all tower layers as one matrix [(n_layers-1) * n][9 * n], box 64 x n/2
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
tower_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_x,
             const __grid_constant__ CUtensorMap tmap_y, const __grid_constant__ CUtensorMap tmap_w_stem,
             const __grid_constant__ CUtensorMap tmap_w_tower, const TowerArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + A_SLOTS * A_SLOT_BYTES;
  uint8_t* smem_epi = smem_b + B_STAGES * B_STAGE_BYTES;
  float* s_bias = reinterpret_cast<float*>(smem_epi + 4 * tc::EPI_WARP_BYTES);   // [2][256]
  float* s_head = s_bias + 512;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + EPI_BYTES);
  uint64_t* a_full = bars;                          // [A_SLOTS]      leader's copy live
  uint64_t* a_empty = a_full + A_SLOTS;             // [A_SLOTS]      per CTA, multicast arrive
  uint64_t* b_full = a_empty + A_SLOTS;             // [B_STAGES]     leader's copy live
  uint64_t* b_empty = b_full + B_STAGES;            // [B_STAGES]     per CTA, multicast arrive
  uint64_t* tmem_full = b_empty + B_STAGES;         // [ACC_STAGES]   per CTA, multicast arrive
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;    // [ACC_STAGES]   leader's copy live, 8 arrivals
  uint64_t* act_ready = tmem_empty + ACC_STAGES;    // [2][4]         per CTA, group and 64-channel chunk: that part of the group's layer output is in memory
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(act_ready + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int pair0 = blockIdx.x >> 1, n_pairs_grid = gridDim.x >> 1;
  const int num_tiles = (args.batch + 1) >> 1;          // two boards each
  const int num_groups = (num_tiles + 1) >> 1;          // four boards: one tile per CTA of the pair
  const int G = args.interleave;                        // groups per super-group
  // A lone group has nothing to hide its epilogue under: it signals every 64 output channels (a fence each: twice the
  // epilogue time, but the next layer starts ~7 000 cycles earlier).  Two interleaved groups signal once per tile: their
  // epilogue must stay shorter than the other group's MMA phase.
  const bool fine = G == 1;
  const int num_super = (num_groups + G - 1) / G;
  const int L = args.n_layers;
  const int n_half = args.n >> 1;

  // per-layer geometry: layer 0 is the stem (k x k, one chunk of stem_row_bytes / 2 channels), the others 3x3 over n channels
  auto lk = [&](int l) -> int { return l == 0 ? args.stem_k : 3; };
  auto lrb = [&](int l) -> int { return l == 0 ? args.stem_row_bytes : 128; };
  auto lchunks = [&](int l) -> int { return l == 0 ? 1 : args.n >> 6; };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_in);
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_y);
    ptx::prefetch_tmap(&tmap_w_stem);
    ptx::prefetch_tmap(&tmap_w_tower);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < A_SLOTS; ++i) {
      ptx::mbar_init(&a_full[i], 1);
      ptx::mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < B_STAGES; ++i) {
      ptx::mbar_init(&b_full[i], 1);
      ptx::mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 8);
    }
    for (int i = 0; i < 8; ++i) ptx::mbar_init(&act_ready[i], 4);   // one arrive per epilogue warp
    ptx::mbar_fence_init();
    ptx::fence_proxy_async();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2cta(tmem_base_slot, ACC_STAGES * ACC_COLS);
    ptx::tmem_relinquish_2cta();
  }
  for (int i = threadIdx.x; i < (args.pf + args.vf) * args.n; i += THREADS) s_head[i] = args.head_w[i];
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (args.pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp != 0) asm volatile("griddepcontrol.wait;" ::: "memory");   // weights are constants; everything else waits
  }

  // Intermediate activations (act0, act1, the fp32 stream) are SCRATCH: written and read back by the same CTA, dead after the
  // last layer.  They are addressed by (pair, group slot, rank) instead of the board number, so that every round of
  // super-groups rewrites the same <= 148 x 2 tiles (76 MB with the fp32 stream: L2-sized) instead of walking through
  // batch-sized buffers whose dead lines the L2 would have to write back to HBM.
  auto scratch_tile = [&](int g) -> int { return 2 * (G * pair0 + g) + static_cast<int>(rank); };
  // number of groups of super-group s that exist (1 or 2)
  auto groups_in = [&](int s) -> int { return G * s + G <= num_groups ? G : num_groups - G * s; };

  if (warp == 0) {
    // ===================== weight producer: this CTA's half of the output channels, every (group, layer) tile =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int s = pair0; s < num_super; s += n_pairs_grid) {
      const int ng = groups_in(s);
      for (int l = 0; l < L; ++l) {
        const int k = lk(l), kblk = lrb(l) >> 1, kch = lchunks(l);
        const uint32_t b_bytes = static_cast<uint32_t>(n_half) * lrb(l);
        const CUtensorMap* tm = l == 0 ? &tmap_w_stem : &tmap_w_tower;
        const int row0 = (l == 0 ? 0 : (l - 1) * args.n) + static_cast<int>(rank) * n_half;
        for (int g = 0; g < ng; ++g)
          for (int kc = 0; kc < kch; ++kc)
            for (int tap = 0; tap < k * k; ++tap) {
              ptx::mbar_wait(&b_empty[stage], phase ^ 1, args.err_flag, ERRT_PRODUCER_B);
              if (ptx::elect_one()) {
                if (leader) ptx::mbar_expect_tx(&b_full[stage], 2 * b_bytes);
                ptx::tma_load_2d_2cta(smem_b + stage * B_STAGE_BYTES, tm, ptx::mapa_u32(&b_full[stage], 0), (tap * kch + kc) * kblk, row0);
              }
              __syncwarp();
              if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
            }
      }
    }
  } else if (warp == 3) {
    // ===================== activation producer: this CTA's two boards of every (group, layer) tile =====================
    int slot = 0;
    uint32_t phase = 0;
    uint32_t ready_phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int s = pair0; s < num_super; s += n_pairs_grid) {
      const int ng = groups_in(s);
      for (int l = 0; l < L; ++l) {
        const int k = lk(l), pad = (k - 1) / 2, hw = 8 + k - 1, rb = lrb(l), kblk = rb >> 1, kch = lchunks(l);
        const int plane_bytes = hw * 2 * hw * rb;
        // stem reads the packed input planes; conv1 (odd l) reads act0, conv2 (even l >= 2) reads act1
        const CUtensorMap* tm = l == 0 ? &tmap_in : ((l & 1) ? &tmap_x : &tmap_y);
        for (int g = 0; g < ng; ++g) {
          // (activations of layers > 0 live in this CTA's scratch tile, see `scratch_tile`)
          const int tile = l == 0 ? 2 * (G * s + g) + static_cast<int>(rank) : scratch_tile(g);
          const bool tr = args.trace && blockIdx.x == 0 && s == pair0 && g == 0 && lane == 0;
          for (int kc = 0; kc < kch; ++kc) {
            if (l > 0 && (fine || kc == 0)) {
              // this CTA's epilogue of (group g, layer l - 1) has stored and fenced what we are about to read: the whole
              // tile, or (lone group) these 64 channels -- the layer starts while the rest of its input is still being drained
              const int r = g * 4 + (fine ? kc : 0);
              ptx::mbar_wait(&act_ready[r], ready_phase[r], args.err_flag, ERRT_ACT_READY);
              ready_phase[r] ^= 1;
              ptx::fence_proxy_async_all();
            }
            if (tr && kc == 0) args.trace[l * 8 + 0] = clock64();
            ptx::mbar_wait(&a_empty[slot], phase ^ 1, args.err_flag, ERRT_PRODUCER_A);
            if (ptx::elect_one()) {
              if (leader) ptx::mbar_expect_tx(&a_full[slot], 2 * plane_bytes);
              ptx::tma_load_4d_2cta(smem_a + slot * A_SLOT_BYTES, tm, ptx::mapa_u32(&a_full[slot], 0), kc * kblk, -pad, tile * 2, -pad);
            }
            __syncwarp();
            if (tr && kc == 0) args.trace[l * 8 + 1] = clock64();
            if (++slot == A_SLOTS) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only) =====================
    if (leader) {
      int stage = 0, slot = 0, acc = 0;
      uint32_t phase = 0, a_phase = 0, acc_phase = 0;
      for (int s = pair0; s < num_super; s += n_pairs_grid) {
        const int ng = groups_in(s);
        for (int l = 0; l < L; ++l) {
          const int k = lk(l), hw = 8 + k - 1, rb = lrb(l), kch = lchunks(l);
          const uint32_t idesc = ptx::idesc_h16_f32(256, args.n, args.fmt);
          const uint64_t a_desc0 = tc::desc_sw(ptx::smem_u32(smem_a), hw * rb, rb);
          const uint64_t b_desc0 = tc::desc_sw(ptx::smem_u32(smem_b), 8 * rb, rb);
          const int tap_rows16 = rb >> 4;
          for (int g = 0; g < ng; ++g) {
            ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1, args.err_flag, ERRT_MMA_ACC);
            ptx::tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
            const bool tr = args.trace && blockIdx.x == 0 && s == pair0 && g == 0 && lane == 0;
            for (int kc = 0; kc < kch; ++kc) {
              ptx::mbar_wait(&a_full[slot], a_phase, args.err_flag, ERRT_MMA_A);
              if (tr && kc == 0) args.trace[l * 8 + 2] = clock64();
              const uint64_t a_plane = a_desc0 + static_cast<uint64_t>((slot * A_SLOT_BYTES) >> 4);
              for (int dy = 0; dy < k; ++dy)
                for (int dx = 0; dx < k; ++dx) {
                  ptx::mbar_wait(&b_full[stage], phase, args.err_flag, ERRT_MMA_B);
                  ptx::tcgen05_fence_after();
                  if (ptx::elect_one()) {
                    const uint64_t adesc = a_plane + static_cast<uint64_t>((dy * 2 * hw + dx) * tap_rows16);
                    const uint64_t bdesc = b_desc0 + static_cast<uint64_t>((stage * B_STAGE_BYTES) >> 4);
                    const uint32_t first = (kc | dy | dx) == 0 ? 0u : 1u;
                    ptx::mma_bf16_ss_2cta(tmem_d, adesc, bdesc, idesc, first);
                    ptx::mma_bf16_ss_2cta(tmem_d, adesc + 2, bdesc + 2, idesc, 1u);
                    if (rb == 128) {
                      ptx::mma_bf16_ss_2cta(tmem_d, adesc + 4, bdesc + 4, idesc, 1u);
                      ptx::mma_bf16_ss_2cta(tmem_d, adesc + 6, bdesc + 6, idesc, 1u);
                    }
                    ptx::mma_commit_2cta(&b_empty[stage]);
                    if (dy == k - 1 && dx == k - 1) {
                      ptx::mma_commit_2cta(&a_empty[slot]);
                      if (kc == kch - 1) ptx::mma_commit_2cta(&tmem_full[acc]);
                    }
                  }
                  __syncwarp();
                  if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
                }
              if (++slot == A_SLOTS) { slot = 0; a_phase ^= 1; }
            }
            if (tr) args.trace[l * 8 + 3] = clock64();
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: this CTA's 128 rows of every (group, layer) tile =====================
    const int quad = warp & 3;
    const int et = threadIdx.x - 128;   // 0..127
    int acc = 0;
    uint32_t acc_phase = 0;
    tc::ConvArgs ca{};
    ca.batch = args.batch;
    ca.num_tiles = num_tiles;
    ca.n = args.n;
    ca.relu = 1;
    ca.head_b = args.head_b;
    ca.featp = args.featp;
    ca.featv = args.featv;
    ca.pf = args.pf;
    ca.vf = args.vf;
    ca.fp16 = args.fmt;
    ca.n_tiles = 1;
    ca.n_total = args.n;
    ca.m_rows = args.batch * 64;
    ca.out_rm = nullptr;
    ca.l2_prefetch = 0;
    ca.err_flag = args.err_flag;
    ca.all_lanes_fence = args.all_lanes_fence;
    for (int s = pair0; s < num_super; s += n_pairs_grid) {
      const int ng = groups_in(s);
      for (int l = 0; l < L; ++l) {
        const bool conv1 = (l & 1) != 0, last = l == L - 1;
        // stem: -> xf + act0;  conv1: -> act1;  conv2: + xf -> xf + act0;  last layer: heads' 1x1 convolutions, no stores
        ca.residual = (l > 0 && !conv1) ? args.xf : nullptr;
        ca.out_f32 = (conv1 || last) ? nullptr : args.xf;
        ca.out_bf16 = last ? nullptr : (conv1 ? args.act1 : args.act0);
        ca.head_w = last ? args.head_w : nullptr;
        for (int g = 0; g < ng; ++g) {
          const int tile = 2 * (G * s + g) + static_cast<int>(rank);
          float* sb = s_bias + acc * 256;
          // (the other epilogue warps may still read this slot's previous contents -- two tiles ago -- only if they were two
          // tiles behind, which the named barrier below rules out: every warp passes it once per tile)
          sb[et] = args.bias_all[l * args.n + (et < args.n ? et : 0)];
          if (et + 128 < args.n) sb[et + 128] = args.bias_all[l * args.n + et + 128];
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const bool tr = args.trace && blockIdx.x == 0 && s == pair0 && g == 0 && threadIdx.x == 128;
          if (tr) {   // when the accumulator is complete (the epilogue proper waits for the same barrier again: no cost)
            ptx::mbar_wait(&tmem_full[acc], acc_phase, args.err_flag, tc::ERR_EPILOGUE);
            args.trace[l * 8 + 4] = clock64();
          }
          tc::epilogue_tile(ca, smem_epi, sb, s_head, quad, lane, tile * 2,
                            tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * ACC_COLS, &tmem_full[acc], acc_phase,
                            &tmem_empty[acc], 0u /* the leader owns the accumulator-empty barrier */, 0, -1, scratch_tile(g) * 2,
                            (last || !fine) ? nullptr : &act_ready[g * 4]);
          if (tr) args.trace[l * 8 + 5] = clock64();
          if (!last && !fine) {
            // what this warp stored must be visible to the TMA reads of the group's next layer (same CTA, async proxy)
            if (args.all_lanes_fence) { __threadfence(); ptx::fence_proxy_async_all(); }
            __syncwarp();
            if (lane == 0) {
              if (!args.all_lanes_fence) ptx::fence_acq_rel_gpu();   // cumulative over the warp's stores (see conv_tc.cuh)
              ptx::mbar_arrive(&act_ready[g * 4]);
            }
          }
          if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc_2cta(tmem_base, ACC_STAGES * ACC_COLS);
  }
}

}  // namespace tw
}  // namespace caz
