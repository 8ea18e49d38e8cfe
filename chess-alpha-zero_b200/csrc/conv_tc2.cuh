// Two-SM variant of the implicit-GEMM convolution (conv_tc.cuh): the two CTAs of a cluster pair run ONE
// tcgen05.mma.cta_group::2 per K step, M = 256 = four boards, each CTA contributing its 128 rows.
//
// Why: with the halo tile the activations cost almost nothing, and a tile is bound by the weight stream -- 32 KB per K
// block that take ~1.2 us to arrive under load, so K blocks complete at (ring depth / latency).  In a CTA pair each
// CTA stages only HALF of every weight block (its 128 of the 256 output channels; the tensor cores exchange the halves
// over the pair's private path), so the same shared memory holds twice as many K blocks in flight, and each SM pulls
// half the weight bytes from L2.
//
// Pair protocol (rank 0 = leader):
//   * both CTAs load their own halo tiles and their own half of each weight block; every load counts its bytes on
//     the LEADER's mbarrier (cp.async.bulk.tensor ... .cta_group::2 with a shared::cluster barrier address)
//   * only the leader's warp 1 issues MMAs; tcgen05.commit.cta_group::2 ... multicast arrives on the stage-empty /
//     accumulator-full barriers of BOTH CTAs
//   * both CTAs' epilogue warps drain their own 128 accumulator rows and arrive on the leader's accumulator-empty
//     barrier (8 arrivals: 4 local, 4 remote)
//   * consecutive layers overlap through programmatic dependent launch (args.pdl, see below)
#pragma once
#include <cuda_bf16.h>

#include "conv_tc.cuh"
#include "ptx.cuh"

namespace caz {
namespace tc2 {

using tc::ConvArgs;

constexpr int THREADS = 256;
constexpr int BLOCK_K = 64;
constexpr int B_STAGES = 8;
constexpr int B_STAGE_BYTES = 128 * BLOCK_K * 2;   // 16 KiB: this CTA's 128 weight rows of one K block
constexpr int A_RING_BYTES = tc::A_RING_BYTES;
constexpr int MAX_A_SLOTS = tc::MAX_A_SLOTS;
constexpr int ACC_STAGES = 2;
constexpr int ACC_COLS = 256;
constexpr int SMEM_BYTES = A_RING_BYTES + B_STAGES * B_STAGE_BYTES + tc::EPI_BYTES + 1024 /*alignment slack*/ + 512 /*barriers*/;

enum : int { ERR2_PRODUCER_B = 301, ERR2_PRODUCER_A = 302, ERR2_MMA_B = 303, ERR2_MMA_A = 304, ERR2_MMA_ACC = 305 };

// tmap_w: weight rows [oc][K] with box 64 x (n/2); tmap_w_whole: the same matrix with box 64 x n, used by the full-width work
// items of a mixed launch (ConvArgs::mixed_whole; any valid map otherwise)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmap_halo, const __grid_constant__ CUtensorMap tmap_w,
                const __grid_constant__ CUtensorMap tmap_w_whole, const ConvArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_epi = smem + A_RING_BYTES + B_STAGES * B_STAGE_BYTES;
  float* s_bias = reinterpret_cast<float*>(smem_epi + 4 * tc::EPI_WARP_BYTES);
  float* s_head = s_bias + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + tc::EPI_BYTES);
  uint64_t* a_full = bars;                          // [MAX_A_SLOTS]   (the leader's copy is the live one)
  uint64_t* a_empty = a_full + MAX_A_SLOTS;         // [MAX_A_SLOTS]   per CTA, multicast arrive
  uint64_t* b_full = a_empty + MAX_A_SLOTS;         // [B_STAGES]      leader's copy live
  uint64_t* b_empty = b_full + B_STAGES;            // [B_STAGES]      per CTA, multicast arrive
  uint64_t* tmem_full = b_empty + B_STAGES;         // [ACC_STAGES]    per CTA, multicast arrive
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;    // [ACC_STAGES]    leader's copy live, 8 arrivals
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const bool wg = args.wg_splits > 0;               // weight-gradient GEMM (see ConvArgs): work items are (tap, K split)
  const int pair0 = blockIdx.x >> 1, n_pairs_grid = gridDim.x >> 1;
  const int num_pairs = (args.num_tiles + 1) >> 1;
  // Two N tiles per board group (n_tiles == 2, args.n = half the width) may be MIXED: the first mixed_whole work items take a
  // whole group (both N tiles in one 2 n wide tile -- one MMA sweep at full efficiency), the groups after them are split.  96
  // groups on 74 CTA pairs: 74 whole tiles, then 44 half tiles = one whole + one half tile time instead of three half tiles.
  const int mixed_whole = (!wg && args.n_tiles == 2) ? args.mixed_whole : 0;
  const int num_work = wg ? args.wg_k * args.wg_k * args.wg_splits   // (tap, K split), else (board-pair pair, N tile)
                          : (args.n_tiles == 2 ? mixed_whole + 2 * (num_pairs - mixed_whole) : num_pairs * args.n_tiles);
  auto item_whole = [&](int t) -> bool { return t < mixed_whole; };
  auto item_nt = [&](int t) -> int { return args.n_tiles == 2 ? (t < mixed_whole ? 0 : (t - mixed_whole) & 1) : t % args.n_tiles; };
  // K-coordinate shift of the B operand for the tap of work item t, and its first K chunk
  auto wg_shift = [&](int t) -> int { const int tap = t / args.wg_splits; return (tap / args.wg_k - (args.wg_k - 1) / 2) * args.wg_pitch + (tap % args.wg_k - (args.wg_k - 1) / 2) * args.wg_dx; };
  auto wg_kc0 = [&](int t) -> int { return (t % args.wg_splits) * args.kchunks; };
  // Consecutive layers walk the batch in opposite directions ("snake"): a layer starts with the tiles its predecessor
  // wrote LAST, which are still in the 126 MB L2 -- the activations of a big batch are larger than L2.
  auto pair_of = [&](int t) -> int {
    const int pr = args.n_tiles == 2 ? (t < mixed_whole ? t : mixed_whole + ((t - mixed_whole) >> 1)) : t / args.n_tiles;
    return args.reverse ? num_pairs - 1 - pr : pr;
  };
  const int k = args.ksize, pad = (k - 1) / 2, hw = 8 + k - 1;
  const int rb = args.row_bytes, kblk = rb >> 1;
  auto wg_col = [&](int t, int kc) -> int {   // K coordinate (elements) of chunk kc of work item t, see ConvArgs::wg_cpc
    int j = wg_kc0(t) + kc;
    if (args.wg_cpc > 0) {
      const int cc = j / args.wg_cpc, sub = j - cc * args.wg_cpc, pitch = args.wg_pitch / args.wg_dx, wpad = (args.wg_k - 1) / 2;
      j = (((cc >> 3) + wpad) * pitch + (cc & 7) + wpad) * args.wg_cpc + sub;
    }
    return j * kblk;
  };
  const int plane_bytes = hw * 2 * hw * rb;
  const int n_half = args.n >> 1;
  const uint32_t b_bytes = static_cast<uint32_t>(n_half) * rb;
  // Ring geometry.  Convolutions: 8 weight stages of 16 KB, and as many chunk tiles as fit 51 KB.  The weight-gradient GEMM
  // streams BOTH operands once (16 KB of A and n/2 rows of B per K chunk, four MMAs each): its weight stages shrink to what
  // a chunk needs, six of them, and the rest of the operand area becomes activation slots (five at 256 columns, eight at 32)
  // -- with three slots the GEMM waited for A most of the time (63 -> 46 us per 3x3 layer with three, measured).
  const int bstage_bytes = wg ? static_cast<int>((b_bytes + 1023u) & ~1023u) : B_STAGE_BYTES;
  const int n_bstages = wg ? 6 : B_STAGES;
  const int a_room = wg ? A_RING_BYTES + B_STAGES * B_STAGE_BYTES - n_bstages * bstage_bytes : A_RING_BYTES;
  const int a_slots = a_room / plane_bytes < MAX_A_SLOTS ? a_room / plane_bytes : MAX_A_SLOTS;
  uint8_t* smem_b = smem + (wg ? a_slots * plane_bytes : A_RING_BYTES);

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
      ptx::mbar_init(&tmem_empty[i], 8);  // four epilogue warps in each CTA of the pair
    }
    ptx::mbar_fence_init();
    ptx::fence_proxy_async();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2cta(tmem_base_slot, ACC_STAGES * ACC_COLS);
    ptx::tmem_relinquish_2cta();
  }
  if (threadIdx.x < args.n && args.n_tiles == 1) s_bias[threadIdx.x] = args.bias[threadIdx.x];
  if (args.head_w)
    for (int i = threadIdx.x; i < (args.pf + args.vf) * args.n; i += THREADS) s_head[i] = args.head_w[i];
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them remotely
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  // Programmatic dependent launch: a layer is one launch, and at a few hundred boards most of a launch is prologue, launch
  // latency and tail.  The NEXT layer's grid may therefore start as soon as every CTA of this one has got here: it sets up
  // its barriers and TMEM and streams its first weights (constants) on idle SMs, and only then waits for this grid to
  // complete (griddepcontrol.wait: completion and visibility of everything the previous grid wrote) before it touches an
  // activation -- every global read of activations and every global write sits behind that wait.
  if (args.pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp != 0) asm volatile("griddepcontrol.wait;" ::: "memory");
  }

  if (warp == 0) {
    // ===================== weight producer: this CTA's half of the output channels =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int t = pair0; t < num_work; t += n_pairs_grid) {
      const int nt = wg ? 0 : item_nt(t);
      const bool whole = !wg && item_whole(t);
      const int rows = whole ? args.n : n_half;   // this CTA's half of the item's output channels
      for (int kc = 0; kc < args.kchunks; ++kc)
        for (int tap = 0; tap < k * k; ++tap) {
          ptx::mbar_wait(&b_empty[stage], phase ^ 1, args.err_flag, ERR2_PRODUCER_B);
          if (ptx::elect_one()) {
            if (leader) ptx::mbar_expect_tx(&b_full[stage], 2u * static_cast<uint32_t>(rows) * rb);   // both CTAs' halves
            const int col = wg ? wg_col(t, kc) + wg_shift(t) : (tap * args.kchunks + kc) * kblk;
            ptx::tma_load_2d_2cta(smem_b + stage * bstage_bytes, whole ? &tmap_w_whole : &tmap_w, ptx::mapa_u32(&b_full[stage], 0),
                                  col, nt * args.n + static_cast<int>(rank) * rows);
          }
          __syncwarp();
          if (++stage == n_bstages) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp == 3) {
    // ===================== activation producer: this CTA's two boards =====================
    int slot = 0;
    uint32_t phase = 0;
    for (int t = pair0; t < num_work; t += n_pairs_grid) {
      const int tile = wg ? static_cast<int>(rank) : 2 * pair_of(t) + static_cast<int>(rank);
      const int kc_first = wg ? wg_kc0(t) : 0;
      // the activations of a big batch come from HBM and the ring holds only two chunk tiles: ask L2 for the NEXT
      // work item's tiles now, a whole tile ahead of need
      if (!wg && args.l2_prefetch && t + n_pairs_grid < num_work && ptx::elect_one()) {
        const int ntile = 2 * pair_of(t + n_pairs_grid) + static_cast<int>(rank);
        for (int kc = 0; kc < args.kchunks; ++kc) ptx::tma_prefetch_4d(&tmap_halo, kc * kblk, -pad, ntile * 2, -pad);
      }
      __syncwarp();
      for (int kc = 0; kc < args.kchunks; ++kc) {
        ptx::mbar_wait(&a_empty[slot], phase ^ 1, args.err_flag, ERR2_PRODUCER_A);
        if (ptx::elect_one()) {
          if (leader) ptx::mbar_expect_tx(&a_full[slot], 2 * plane_bytes);
          ptx::tma_load_4d_2cta(smem_a + slot * plane_bytes, &tmap_halo, ptx::mapa_u32(&a_full[slot], 0),
                                wg ? wg_col(t, kc) : (kc_first + kc) * kblk, -pad, tile * 2, -pad);
        }
        __syncwarp();
        if (++slot == a_slots) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only) =====================
    if (leader) {
      const uint32_t idesc_item = ptx::idesc_h16_f32(256, args.n, args.fp16);
      const uint32_t idesc_whole = mixed_whole > 0 ? ptx::idesc_h16_f32(256, 2 * args.n, args.fp16) : idesc_item;
      const uint64_t a_desc0 = tc::desc_sw(ptx::smem_u32(smem_a), hw * rb, rb);
      const uint64_t b_desc0 = tc::desc_sw(ptx::smem_u32(smem_b), 8 * rb, rb);
      const int tap_rows16 = rb >> 4;
      int stage = 0, slot = 0, acc = 0;
      uint32_t phase = 0, a_phase = 0, acc_phase = 0;
      for (int t = pair0; t < num_work; t += n_pairs_grid) {
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1, args.err_flag, ERR2_MMA_ACC);
        ptx::tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
        const uint32_t idesc = (!wg && item_whole(t)) ? idesc_whole : idesc_item;
        for (int kc = 0; kc < args.kchunks; ++kc) {
          ptx::mbar_wait(&a_full[slot], a_phase, args.err_flag, ERR2_MMA_A);
          const uint64_t a_plane = a_desc0 + static_cast<uint64_t>((slot * plane_bytes) >> 4);
          for (int dy = 0; dy < k; ++dy)
            for (int dx = 0; dx < k; ++dx) {
              ptx::mbar_wait(&b_full[stage], phase, args.err_flag, ERR2_MMA_B);
              ptx::tcgen05_fence_after();
              if (ptx::elect_one()) {
                const uint64_t adesc = a_plane + static_cast<uint64_t>((dy * 2 * hw + dx) * tap_rows16);
                const uint64_t bdesc = b_desc0 + static_cast<uint64_t>((stage * bstage_bytes) >> 4);
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
                  if (kc == args.kchunks - 1) ptx::mma_commit_2cta(&tmem_full[acc]);
                }
              }
              __syncwarp();
              if (++stage == n_bstages) { stage = 0; phase ^= 1; }
            }
          if (++slot == a_slots) { slot = 0; a_phase ^= 1; }
        }
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: this CTA's 128 rows =====================
    const int quad = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = pair0; t < num_work; t += n_pairs_grid) {
      if (wg) {
        // one [m_rows][n_total] partial product per (tap, K split), this CTA's 128 of the 256 rows
        ConvArgs wa = args;
        wa.out_rm = args.out_rm + static_cast<size_t>(t) * args.m_rows * args.n_total;
        tc::epilogue_tile(wa, smem_epi, s_bias, s_head, quad, lane, static_cast<int>(rank) * 2,
                          tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * ACC_COLS, &tmem_full[acc], acc_phase,
                          &tmem_empty[acc], 0u, 0, -1);
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      const int tile = 2 * pair_of(t) + static_cast<int>(rank);
      const int tn = t + n_pairs_grid;
      const int next_tile = tn < num_work ? 2 * pair_of(tn) + static_cast<int>(rank) : -1;
      ConvArgs ea = args;
      if (item_whole(t)) ea.n = 2 * args.n;   // columns [0, 2 n) of the n_total wide rows, N tile 0
      tc::epilogue_tile(ea, smem_epi, s_bias, s_head, quad, lane, tile * 2,
                        tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * ACC_COLS, &tmem_full[acc], acc_phase,
                        &tmem_empty[acc], 0u /* the leader owns the accumulator-empty barrier */, item_nt(t),
                        next_tile >= 0 && next_tile < args.num_tiles ? next_tile * 2 : -1);
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();  // nobody frees TMEM or exits while the peer may still signal it
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc_2cta(tmem_base, ACC_STAGES * ACC_COLS);
  }
}

}  // namespace tc2
}  // namespace caz
