// Small-batch path: the WHOLE forward pass (input conversion, stem, every residual conv, both heads) in ONE persistent,
// cooperatively launched kernel, for MCTS-sized batches (reference batches are <= 48 positions:
// search_threads x max_processes, configs/normal.py:30-31).
//
// At batch 16 a forward pass is ~17 GFLOP -- nothing for the tensor cores -- but every layer needs the 1.18 MB of its
// weights and has to wait for the previous layer: the pass is bound by latency chains and by operand bandwidth.  So:
//   * grid = (board tiles) x (output-channel slices): a tile's rows and N_s = 256/S output channels belong to the same
//     CTA(s) in EVERY layer.  Up to 36 boards a slice of a PAIR of boards is run by a CTA pair (cta_group::2, M = 128:
//     one board per CTA, each CTA staging half of the slice's weight rows, the leader issuing for both); above that one
//     CTA takes two boards (M = 128) and a 64-channel slice.  The fp32 residual stream of a thread's row and channels
//     stays in its registers; only 16-bit activations cross CTAs, through L2.
//   * a tcgen05.mma at these shapes costs its operand reads from shared memory, (M + N) x 32 B at 128 B/cycle
//     (tools/ub/ub_small.cu), IF the issuing thread keeps up: descriptors are hoisted, tap geometry is compile-time, two
//     warps issue alternate 64-channel chunks into their own accumulators (4 TMEM column groups, added by the epilogue).
//   * weights: the 16-bit weight set (17.4 MB) is L2-resident; each CTA streams its [N_s x K] slice through a
//     shared-memory ring with TMA, running AHEAD across layer boundaries.
//   * activations: ONE zero-padded halo tile per 64-channel chunk, [10 ranks][nb boards][10 files][64 ch] (TMA box
//     with out-of-bounds fill), serves all nine taps: tap (dy, dx) is the same tile read through a UMMA descriptor
//     whose start address is shifted by (10*nb*dy + dx) rows and whose 8-row-group stride is 10 rows.  The stem's tile
//     is built in place by the CTA itself from the fp32 planes / 68-byte records (no pack phase, no global round trip).
//   * layers rendezvous per board tile through progress flags (L2 atomics, one polling thread, one acquire fence);
//     every tile has a flag line on each die's L2 and polls the near one (see Args::flag_line).  Cooperative launch
//     guarantees co-residency; every spin is bounded and traps instead of hanging.
//   * heads per tile: head-conv partial sums from the last epilogue, Dense weights staged over the tower's tiles while
//     it runs, softmax / value statistics exchanged through the tile's flags; canonical summation order, so outputs
//     are bit-identical for every batch size.
// TMEM lane m of an M = 128 accumulator is row (rank h = m>>4, board p = (m>>3)&1, file w = m&7).
#pragma once
#include <cuda_bf16.h>

#include "kernels_simt.cuh"
#include "ptx.cuh"

namespace caz {
namespace sb {

constexpr int THREADS = 256;
constexpr int A_PLANES = 4;                          // up to 256 input channels
constexpr int A_PLANE_BYTES_3 = 10 * 2 * 10 * 128;   // 25 600: 3x3 halo tile of one 64-channel chunk
constexpr int A_BYTES = A_PLANES * A_PLANE_BYTES_3;  // 102 400 (a 5x5 / 7x7 single-plane stem tile fits inside)
constexpr int W_RING_BYTES = 108 * 1024;
constexpr int W_SLOT_MAX_BYTES = 36 * 1024;   // a whole 3x3 filter of 32 output channels
constexpr int MAX_W_SLOTS = 16;
constexpr int NACC = 4;   // accumulator groups in TMEM: one per K sub-step of a 64-channel chunk row
constexpr int BAR_BYTES = 1024;
constexpr int SMEM_BYTES = A_BYTES + W_RING_BYTES + BAR_BYTES + 1024 /*alignment*/;
constexpr int HEAD_SMEM_BYTES = A_BYTES + W_RING_BYTES;        // what the head phases may use for staged weights
constexpr int MAX_BATCH = 148;  // 74 two-board tiles x 2 slices of 128 channels = 148 CTAs (up to 72 boards: 36 CTA pairs x 2 slices x 2)
constexpr int MAX_TILES = 74;   // board tiles per launch (one flag line each)

struct Args {
  int batch;        // positions (<= MAX_BATCH)
  int n_layers;     // 1 + 2 * res_blocks
  int nb;           // boards per CTA tile: 2 (M = 128) or 1 (M = 64: half the A-operand bytes per MMA)
  int ns;           // output channels per slice (16, 32, 64 or 128)
  int pair;         // 1: the two CTAs of a cluster share a slice (cta_group::2, M = 128: one board each, half of the
                    //    slice's weights each); needs nb = 1 and ns >= 32
  int n_slices;     // filters / ns
  int filters;
  int stem_k, stem_kchunks;  // stem kernel size, input channel chunks (in_cpad / 64)
  int blk_k;                  // tower kernel size
  int in_planes, in_cpad;
  int fp16;                   // operand formats: ptx::FMT_A_FP16 (activations) | ptx::FMT_B_FP16 (weights); 0 = both bf16
  int dbg;                    // measurement switch (CAZ_SB_DBG), 0 in production: bits 8.. = the CTA that writes the trace
  int slot_bytes, n_slots;    // weight ring geometry
  const float* planes;        // fp32 [batch][in_planes][64] or nullptr
  const uint8_t* records;     // u8 [batch][68] or nullptr (exactly one of the two)
  const float* bias_all;      // [n_layers][filters]
  __nv_bfloat16* xb;          // bf16 copy of it         (A operand of conv1)
  __nv_bfloat16* y;           // bf16 conv1 output       (A operand of conv2)
  // heads (fp32)
  int pf, vf, value_fc, n_labels;
  const float* head_w;        // [pf+vf][filters]  1x1 head convolutions, BN folded
  const float* head_b;        // [pf+vf]
  float* hpart;               // [batch][16 channel groups][8 head filters][64 squares] partial head-conv sums
  const float* pfc_w;         // [pf*64][n_labels]
  const float* pfc_wb;        // the same weights in 128-label blocks, [units][pf*64][128] zero-padded: ONE bulk copy per unit
  const float* pfc_b;
  const float* vfc1_w;        // [vf*64][value_fc]
  const float* vfc1_wb;       // in 16-unit blocks, [groups][vf*64][16] zero-padded: one bulk copy per hidden group
  const float* vfc1_b;
  const float* vfc2_w;        // [value_fc]
  const float* vfc2_b;
  float* hstats;              // [batch][HSTAT_FLOATS]: per 128-label unit (max, sum of exp), per 16-unit group value partial
  const uint8_t* mask;        // nullptr or [batch][n_labels]
  float* policy;              // [batch][n_labels]
  float* value;               // [batch]
  // Progress flags: base + number of phases a CTA has completed; one 128-byte line per board tile (words 0-15: the
  // slices' flags, words 16-31: scratch for the start-up probe).  B200 is two dies, and an L2 atomic on a line homed in
  // the OTHER die's L2 takes twice as long to become visible to a gpu-scope poll (~720 vs ~340 cycles, measured); with
  // all flags on one die half of the tiles ran 6.2 us per layer instead of 5.4.  So every tile has two flag lines, one
  // from each latency class (caz_abi.cu classifies candidate lines at creation): a CTA publishes to both and polls the
  // one that answers faster from its own SM (probed once per launch, off the critical path).
  unsigned* flag_base;
  const unsigned char* sm_near;       // nullptr or [%smid] -> the copy that is near for that SM (0xFF: unknown), caz_abi.cu
  uint16_t flag_line[2][MAX_TILES];   // [copy][tile] -> line index (x 32 words)
  unsigned base;              // flag value that means "nothing of this launch done yet" (monotonic across launches)
  long long* trace;           // nullptr or [n_layers+1][8] clock64 stamps of CTA 0 (see caz_debug_sb_trace)
  int* err_flag;
};

enum : int { ERR_HEADW = 207, ERR_W = 201, ERR_A = 202, ERR_GRID = 203, ERR_MMA_A = 204, ERR_MMA_W = 205, ERR_EPI = 206 };

// Grid phases: 0 input pack, 1..L the L conv layers, L+1 head 1x1 convs, L+2 head dense layers.
// A CTA publishes flag = base + 1 + l after layer l and base + 1 + n_layers after its head statistics.
constexpr int FLAG_STEPS_EXTRA = 1;
// Heads: every quantity that is summed across CTAs is split into CANONICAL pieces that do not depend on how many
// slices the grid has -- 16-channel groups for the 1x1 head convolutions, 128-label units for the softmax statistics,
// 16-unit groups for the value head's hidden layer -- and the pieces are combined in a fixed order, so the outputs
// are bit-identical for every batch size / grid geometry.
constexpr int HEAD_FILTERS_MAX = 8;    // pf + vf
constexpr int LABEL_UNIT = 128;        // labels per softmax unit (4 warps)
constexpr int LABEL_UNITS_MAX = 16;    // => n_labels <= 2048
constexpr int HID_GROUP = 16;          // hidden units per value-head group
constexpr int HID_GROUPS_MAX = 16;     // => value_fc <= 256
constexpr int HSTAT_FLOATS = 2 * LABEL_UNITS_MAX + HID_GROUPS_MAX;   // per position
constexpr int LABEL_CHUNK = 256;       // labels per pass over the staged policy weights (one per thread)
constexpr int LCHUNKS_MAX = 4;         // passes per CTA: 1 024 labels, i.e. two slices per board at the least

// taps per weight-ring slot: the whole filter, one filter row, or one tap -- the largest that fits a slot
__host__ __device__ inline int taps_per_slot(int k, int ns) {
  if (k * k * ns * 128 <= W_SLOT_MAX_BYTES) return k * k;
  if (k * ns * 128 <= W_SLOT_MAX_BYTES) return k;
  return 1;
}

// ... for a layer in a ring whose slots were sized by ANOTHER layer (the stem's 5x5 filter rows do not fit where the tower's
// 3x3 rows do): as many single taps as the slot holds, instead of one tap in a mostly empty slot -- with four 24 KB slots
// holding 8 KB each, the 25 taps of the stem were a chain of 25 weight round trips (7.2 us at 48 boards, 2.6 us per tower layer)
__host__ __device__ inline int taps_in_slot(int k, int ns, int slot_bytes) {
  const int t = taps_per_slot(k, ns);
  if (t > 1) return t;
  const int fit = slot_bytes / (ns * 128);
  return fit < 1 ? 1 : (fit > k * k ? k * k : fit);
}

__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void sync_trap(const Args& args) {
  if (args.err_flag) atomicExch(args.err_flag, ERR_GRID);
  __threadfence_system();
  __trap();
}

// ONE thread: wait until every CTA of board tile `tile` has published `target`, then acquire.  (Tile-local on purpose:
// a tile's CTAs only ever read what CTAs of the same tile wrote.  Polling the flags of every tile from every CTA was
// measured to swamp the L2 slices that hold them.)
__device__ __forceinline__ void wait_tile(const Args& args, const unsigned* tile_flags, unsigned target) {
  uint32_t spins = 0;
  for (;;) {
    // 16 flag words per tile (n_slices <= 16 of them in use), fetched as four independent 16-byte loads
    uint4 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      v[q] = q * 4 < args.n_slices ? ptx::ld_relaxed_gpu_v4(tile_flags + q * 4) : make_uint4(target, target, target, target);
    int behind = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      behind |= static_cast<int>(v[q].x - target);
      if (q * 4 + 1 < args.n_slices) behind |= static_cast<int>(v[q].y - target);
      if (q * 4 + 2 < args.n_slices) behind |= static_cast<int>(v[q].z - target);
      if (q * 4 + 3 < args.n_slices) behind |= static_cast<int>(v[q].w - target);
    }
    if (behind >= 0) break;
    if (++spins > (1u << 24)) sync_trap(args);
  }
  ptx::fence_acq_rel_gpu();   // measured cost: 2.4 us per pass; the data is consumed through TMA, but the model wants it
}

__device__ __forceinline__ unsigned* tile_flag_line(const Args& args, int copy, int tile) {
  return args.flag_base + static_cast<size_t>(args.flag_line[copy][tile]) * 32;
}
// one thread: publish this CTA's progress in both copies (fire and forget)
__device__ __forceinline__ void publish(const Args& args, int tile, int slice, unsigned value) {
  ptx::publish_flag_gpu(tile_flag_line(args, 0, tile) + slice, value);
  ptx::publish_flag_gpu(tile_flag_line(args, 1, tile) + slice, value);
}

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   ptx::smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}

// Geometry of the head phases for a grid with n_slices output-channel slices per board tile.
struct HeadGeom {
  int units_total, upl;      // 128-label units in all / per slice
  int groups_total, gps;     // 16-unit hidden groups in all / per slice
  int hsp;                   // hidden units per slice = gps * 16 (row pitch of the staged value weights)
};
__host__ __device__ inline HeadGeom head_geom(int n_labels, int value_fc, int n_slices) {
  HeadGeom g;
  g.units_total = (n_labels + LABEL_UNIT - 1) / LABEL_UNIT;
  g.upl = (g.units_total + n_slices - 1) / n_slices;
  g.groups_total = (value_fc + HID_GROUP - 1) / HID_GROUP;
  g.gps = (g.groups_total + n_slices - 1) / n_slices;
  g.hsp = g.gps * HID_GROUP;
  return g;
}
// can the head phases run for this architecture with `n_slices` slices (staged weights must fit the tower's smem)
__host__ __device__ inline bool heads_fit(int pf, int vf, int n_labels, int value_fc, int n_slices) {
  if (pf + vf > HEAD_FILTERS_MAX || (n_labels & 3) || (value_fc & 3)) return false;
  const HeadGeom g = head_geom(n_labels, value_fc, n_slices);
  if (g.units_total > LABEL_UNITS_MAX || g.groups_total > HID_GROUPS_MAX) return false;
  if ((g.upl * LABEL_UNIT + LABEL_CHUNK - 1) / LABEL_CHUNK > LCHUNKS_MAX) return false;   // logits live in registers
  // (the value Dense columns are staged next to the policy chunk if they fit, else read from L2: value_staged())
  return pf * 64 * LABEL_CHUNK * 4 <= HEAD_SMEM_BYTES && (pf + vf) * 64 <= 512;
}

// staged value weights: [group][vf*64][16], groups 16 floats further apart than their size (bank spread)
__host__ __device__ inline int value_group_pitch(int vf) { return vf * 64 * HID_GROUP + 16; }
__host__ __device__ inline bool value_staged(int pf, int vf, int hsp) {
  return (pf * 64 * LABEL_CHUNK + hsp / HID_GROUP * value_group_pitch(vf)) * 4 <= HEAD_SMEM_BYTES;
}

// Stage head weights for this CTA: the policy Dense kernel of 128-label units [unit0, unit0 + n_units) (n_units <= 2)
// -> s_wp[unit][k][128]; with_value: hidden groups [grp0, grp0 + n_groups) of the value Dense kernel ->
// s_wv[group][k][16].  One bulk copy per unit / group from the blocked copies of the weights (a bulk copy per weight
// ROW, 384 of them, took the issuing warp 10 us: the copy engine accepts one every ~50 cycles).  One warp; completion
// on `bar` (one phase per call).
__device__ __forceinline__ void stage_head_weights(const Args& args, float* s_wp, float* s_wv, uint64_t* bar, int lane,
                                                   int unit0, int n_units, bool with_value, int grp0, int n_groups) {
  const int kp = args.pf * 64, kv = args.vf * 64;
  const uint32_t unit_bytes = static_cast<uint32_t>(kp) * LABEL_UNIT * 4u, grp_bytes = static_cast<uint32_t>(kv) * HID_GROUP * 4u;
  if (!with_value) n_groups = 0;
  if (lane == 0) ptx::mbar_expect_tx(bar, static_cast<uint32_t>(n_units) * unit_bytes + static_cast<uint32_t>(n_groups) * grp_bytes);
  __syncwarp();
  if (lane < n_units)
    bulk_g2s(s_wp + static_cast<size_t>(lane) * kp * LABEL_UNIT, args.pfc_wb + static_cast<size_t>(unit0 + lane) * kp * LABEL_UNIT,
             unit_bytes, bar);
  else if (lane >= 8 && lane - 8 < n_groups)
    bulk_g2s(s_wv + static_cast<size_t>(lane - 8) * value_group_pitch(args.vf),
             args.vfc1_wb + static_cast<size_t>(grp0 + lane - 8) * kv * HID_GROUP, grp_bytes, bar);
}

// The MMAs of `TPS` consecutive taps of one 64-channel chunk (one weight-ring slot), issued by ONE thread.
// A tcgen05.mma at these shapes occupies the tensor pipe for (M + N) * 32 B / 128 B per cycle of operand reads
// (M = 64: 25 cycles, M = 128: 41 cycles at N = 16; tools/ub/ub_small.cu) -- IF the issuing thread keeps up: every
// scalar instruction between two MMAs costs the pipe time.  So everything is hoisted: the caller passes descriptors
// that already point at this chunk and slot, the tap geometry is compile-time where the slot shape allows it, and
// what is left per MMA is two 64-bit adds.
// Two warps issue, taking alternate 64-channel chunks of a layer, so that one of them is feeding the pipe while the
// other waits for weights; each accumulates into its own two TMEM column groups (even / odd K sub-steps).  Which MMA
// lands in which accumulator, and in what order, depends only on (chunk, tap, sub-step) -- not on the slot shape or
// the grid geometry -- so results are bit-identical across batch sizes.
//   a_plane: A descriptor of the chunk's halo tile;  b_slot: B descriptor of the ring slot
//   (dy0, dx0): filter position of the slot's first tap; rp8 = halo rows between ranks * 8 (descriptor units of 16 B)
template <bool PAIR>
__device__ __forceinline__ void mma_h16(uint32_t tm, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if constexpr (PAIR) ptx::mma_bf16_ss_2cta(tm, adesc, bdesc, idesc, accum);
  else ptx::mma_bf16_ss(tm, adesc, bdesc, idesc, accum);
}
template <bool PAIR>
__device__ __forceinline__ void commit_to(uint64_t* bar) {
  if constexpr (PAIR) ptx::mma_commit_2cta(bar);   // arrives on the barrier at this offset in BOTH CTAs of the pair
  else ptx::mma_commit(bar);
}

template <int K, int TPS, bool PAIR>
__device__ __forceinline__ void issue_taps(uint64_t a_plane, uint64_t b_slot, int dy0, int dx0, uint64_t rp8,
                                           uint64_t tap_desc, uint32_t tm0, uint32_t tm1, uint32_t idesc, bool fresh) {
  const uint64_t a_row = a_plane + static_cast<uint64_t>(dy0) * rp8 + static_cast<uint64_t>(dx0 * 8);
#pragma unroll
  for (int t = 0; t < TPS; ++t) {
    // TPS == K*K: the whole filter (dy0 = dx0 = 0); TPS == K: one filter row (dx0 = 0); TPS == 1: one tap
    const uint64_t adesc = TPS == K * K ? a_row + static_cast<uint64_t>(t / K) * rp8 + static_cast<uint64_t>((t % K) * 8)
                                        : a_row + static_cast<uint64_t>(t * 8);
    const uint64_t bdesc = b_slot + static_cast<uint64_t>(t) * tap_desc;
    const uint32_t accum = (t > 0 || !fresh) ? 1u : 0u;   // this issuer's first tap of a layer initialises its accumulators
    mma_h16<PAIR>(tm0, adesc, bdesc, idesc, accum);
    mma_h16<PAIR>(tm1, adesc + 2, bdesc + 2, idesc, accum);
    mma_h16<PAIR>(tm0, adesc + 4, bdesc + 4, idesc, 1u);
    mma_h16<PAIR>(tm1, adesc + 6, bdesc + 6, idesc, 1u);
  }
}

// RES: 16-channel chunks of the slice a thread's epilogue carries (4: slices up to 64 channels per CTA; 8: the 128-channel
// slices of 73-148 boards -- a separate instantiation, so that the latency-critical small forms keep their register budget)
template <bool PAIR, int RES = 4>
__global__ void __launch_bounds__(THREADS, 1)
forward_sb_kernel(const __grid_constant__ CUtensorMap tm_xb,
                  const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_w_stem,
                  const __grid_constant__ CUtensorMap tm_w_tower, const Args args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_w = smem + A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_BYTES + W_RING_BYTES);
  uint64_t* a_full = bars;                       // [A_PLANES]
  uint64_t* w_full = bars + A_PLANES;            // [MAX_W_SLOTS]
  uint64_t* w_empty = w_full + MAX_W_SLOTS;      // [MAX_W_SLOTS]
  uint64_t* acc_full = w_empty + MAX_W_SLOTS;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  uint64_t* hw_full = acc_full + 2;              // [1] staged head weights
  uint64_t* tower_done = acc_full + 3;           // [1] the last layer's MMAs have completed (both issuers)
  uint64_t* stem_ready = acc_full + 4;           // [1] the stem tile(s) built in place are complete (both CTAs of a pair)
  __shared__ __align__(16) float s_hw[HEAD_FILTERS_MAX * 128];  // this slice's columns of the 1x1 head convolutions
  __shared__ __align__(16) float s_bias[2][128];                // this slice's folded BatchNorm shifts of the current / next layer
  __shared__ __align__(16) float s_feat[2 * 512];               // head-conv outputs of this tile's boards (Flatten order)
  __shared__ float s_red[64];
  __shared__ float s_vb[256], s_vw[256];                        // this slice's hidden units: bias, output weight
  __shared__ float s_stat[2 * HSTAT_FLOATS];                    // the tile's boards' head statistics (H3)
  __shared__ uint32_t s_rec[2 * 17];                            // this tile's board records (68 bytes each)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // which CTA takes which (tile, slice): consecutive block ids go to different GPCs (classic placement), so
  // tile = id / n_slices spreads a tile over the GPCs, tile = id % n_tiles (dbg bit 1) keeps strides together
  const unsigned trace_cta = static_cast<unsigned>(args.dbg >> 8);   // which CTA writes the per-layer trace (debug)
  // pair mode: cluster c = blockIdx.x / 2 is (board pair c / n_slices, slice c % n_slices); its CTA `rank` takes board
  // 2 * pair + rank and stages rows [rank * ns/2, +ns/2) of the slice's weights; rank 0 (the leader) issues the MMAs.
  const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int unit = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile = PAIR ? (unit / args.n_slices) * 2 + static_cast<int>(rank) : unit / args.n_slices;
  const int slice = unit % args.n_slices;
  const int n0 = tile * args.nb, col0 = slice * args.ns;
  const int ns_cta = PAIR ? args.ns / 2 : args.ns;   // weight rows this CTA stages = TMEM columns per accumulator
  __shared__ int s_near_copy;   // which of the tile's two flag lines this SM polls (set by warp 3 during layer 0)
  const int tap_bytes = ns_cta * 128;
  const uint32_t tmem_cols = static_cast<uint32_t>(NACC * ns_cta);  // 64 .. 512: powers of two
  const unsigned L = static_cast<unsigned>(args.n_layers);
  // Heads geometry.  The two CTAs of a pair hold the same slice, i.e. they would stage the same Dense weights for one
  // board each.  With few slices (37-72 boards: 2 slices = 1 024 labels and 128 hidden units per CTA, four rounds of
  // weight staging) they split the slice's labels and hidden units instead, and each does its share for BOTH boards:
  // half the staging, and the value Dense columns fit next to the policy chunk again.  The canonical pieces (128-label
  // units, 16-unit groups) and their summation order do not change, so neither do the results.
  const bool hshare = PAIR && args.n_slices <= 4;
  const int h_slices = hshare ? 2 * args.n_slices : args.n_slices;
  const int hslice = hshare ? 2 * slice + static_cast<int>(rank) : slice;
  const int hn0 = hshare ? (tile & ~1) : n0, hnb = hshare ? 2 : args.nb;   // the boards this CTA's heads serve
  const HeadGeom hg = head_geom(args.n_labels, args.value_fc, h_slices);
  const int lab0 = hslice * hg.upl * LABEL_UNIT;                        // this CTA's labels: [lab0, lab0 + lab_cnt)
  const int lab_cnt = max(0, min(hg.upl * LABEL_UNIT, args.n_labels - lab0));
  const int hid0 = hslice * hg.hsp;                                     // this CTA's hidden units: [hid0, hid0 + hid_cnt)
  const int hid_cnt = max(0, min(hg.hsp, args.value_fc - hid0));
  float* s_wp = reinterpret_cast<float*>(smem);                         // heads: [2 units][pf*64][128], over the tower's tiles
  float* s_wv = s_wp + args.pf * 64 * LABEL_CHUNK;                      //        [gps groups][vf*64][16] (+16 between groups)
  const int unit_first = hslice * hg.upl;                               // this CTA's 128-label units: [unit_first, +upl)
  auto units_in_chunk = [&](int c) {                                    // units of label chunk c that exist (0..2)
    return max(0, min(min(2, hg.upl - 2 * c), hg.units_total - (unit_first + 2 * c)));
  };
  long long* ktrace = (args.trace && blockIdx.x == trace_cta && threadIdx.x == 32) ? args.trace + args.n_layers * 8 : nullptr;
  if (args.trace && threadIdx.x == 0) {   // where does this CTA run
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    args.trace[1024 + (15 * 256 + blockIdx.x) * 2] = smid;
  }
  if (ktrace) {
    ktrace[0] = clock64();
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    ktrace[8] = gt;
  }

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_xb);
    ptx::prefetch_tmap(&tm_y);
    ptx::prefetch_tmap(&tm_w_stem);
    ptx::prefetch_tmap(&tm_w_tower);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < A_PLANES; ++i) ptx::mbar_init(&a_full[i], 1);
    for (int i = 0; i < args.n_slots; ++i) {
      ptx::mbar_init(&w_full[i], 1);
      ptx::mbar_init(&w_empty[i], 1);   // the tcgen05.commit of the issuer warp that consumed the slot
    }
    ptx::mbar_init(acc_full, 2);
    ptx::mbar_init(hw_full, 1);
    ptx::mbar_init(tower_done, 2);
    ptx::mbar_init(stem_ready, PAIR ? 2 : 1);
    ptx::mbar_fence_init();
    ptx::fence_proxy_async();
  }
  if (warp == 2) {
    if constexpr (PAIR) { ptx::tmem_alloc_2cta(tmem_slot, tmem_cols); ptx::tmem_relinquish_2cta(); }
    else { ptx::tmem_alloc(tmem_slot, tmem_cols); ptx::tmem_relinquish(); }
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  if constexpr (PAIR) ptx::cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them remotely
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ===================== phase 0: the stem's halo tile, built in place =====================
  // fp32 NCHW planes (what predict_on_batch receives) or 68-byte board records -> the zero-padded tile
  // [hw ranks][nb boards][hw files][64 channels] of 16-bit operands, written straight into shared memory in the layout a
  // TMA box with the 128-byte swizzle would have produced (16-byte chunk index XOR row & 7; the tile is 1024-byte
  // aligned).  Every CTA of a tile builds its own copy: 1 152 values per board, no global round trip, no rendezvous.
  // The weight producer does not take part: it starts streaming layer 0's weights immediately.
  if (warp != 0) {
    const int nthr = THREADS - 32, t = threadIdx.x - 32;
    const int hw = 8 + args.stem_k - 1, pad = (args.stem_k - 1) / 2;
    const int rows = hw * args.nb * hw;
    uint4* tile16 = reinterpret_cast<uint4*>(smem_a);
    const int per_board = args.in_planes * 64, total = args.nb * per_board;
    constexpr int PF = 8;   // values per thread requested at once (the loads used to be waited for one by one)
    for (int base = 0; base < total || base == 0; base += nthr * PF) {
      float v[PF];
      if (args.planes) {
        // fp32 NCHW planes: this tile's boards are contiguous, element i of the tile is planes[n0 * per_board + i]
#pragma unroll
        for (int j = 0; j < PF; ++j) {
          const int i = base + j * nthr + t;
          v[j] = (i < total && n0 + i / per_board < args.batch) ? args.planes[static_cast<size_t>(n0) * per_board + i] : 0.0f;
        }
      }
      if (base == 0) {
        for (int i = t; i < rows * 8; i += nthr) tile16[i] = make_uint4(0u, 0u, 0u, 0u);
        // board records are fetched once (17 words per board; they may live in the caller's pinned host memory) and
        // decoded from shared memory
        if (args.records && t < args.nb * 17) {
          const int p = t / 17, wd = t - p * 17;
          s_rec[p * 17 + wd] = n0 + p < args.batch ? reinterpret_cast<const uint32_t*>(args.records + static_cast<size_t>(n0 + p) * 68)[wd] : 0u;
        }
        asm volatile("bar.sync 2, 224;" ::: "memory");
      }
#pragma unroll
      for (int j = 0; j < PF; ++j) {
        const int i = base + j * nthr + t;
        if (i >= total) break;
        const int p = i / per_board, rem = i - p * per_board;
        const int c = rem >> 6, sq = rem & 63;   // consecutive threads -> consecutive squares of one plane (coalesced)
        if (n0 + p >= args.batch) continue;
        const float val = args.planes ? v[j] : record_plane_value(reinterpret_cast<const uint8_t*>(s_rec + p * 17), c, sq);
        const int row = ((sq >> 3) + pad) * args.nb * hw + p * hw + (sq & 7) + pad;
        const int off = row * 128 + (((c >> 3) ^ (row & 7)) << 4) + ((c & 7) << 1);
        *reinterpret_cast<uint16_t*>(smem_a + off) = ptx::pack_h16(val, (args.fp16 & ptx::FMT_A_FP16) != 0);
      }
    }
    ptx::fence_proxy_async();   // generic-proxy writes -> the tensor core's async-proxy reads
    asm volatile("bar.sync 2, 224;" ::: "memory");
    if (ktrace) ktrace[1] = clock64();
  }

  // Single-thread roles are warp-uniform loops with only the instruction issue under elect.sync (see conv_tc.cuh).
  if (warp == 0) {
    // ============ weight producer: streams this CTA's slice of every layer, ahead of the barriers ============
    int slot = 0;
    uint32_t phase = 0;
    for (int l = 0; l < args.n_layers; ++l) {
      const int k = l == 0 ? args.stem_k : args.blk_k;
      const int kch = l == 0 ? args.stem_kchunks : args.filters / 64;
      const int tps = taps_in_slot(k, ns_cta, args.slot_bytes);
      const CUtensorMap* tm = l == 0 ? &tm_w_stem : &tm_w_tower;
      const int row = (l == 0 ? 0 : (l - 1) * args.filters) + col0 + static_cast<int>(rank) * ns_cta;
      for (int kc = 0; kc < kch; ++kc)
        for (int tap0 = 0; tap0 < k * k; tap0 += tps) {
          ptx::mbar_wait(&w_empty[slot], phase ^ 1, args.err_flag, ERR_W);
          const int nt = min(tps, k * k - tap0);   // (a filter whose taps do not divide into slots: the last slot is short)
          if (ptx::elect_one()) {
            if constexpr (!PAIR) {
              ptx::mbar_expect_tx(&w_full[slot], nt * tap_bytes);
              for (int t = 0; t < nt; ++t)
                ptx::tma_load_2d(smem_w + slot * args.slot_bytes + t * tap_bytes, tm, &w_full[slot],
                                 ((tap0 + t) * kch + kc) * 64, row);
            } else {
              // both CTAs' halves of the slot count on the LEADER's barrier; each CTA waits for its own slot to be free
              // (the leader's tcgen05.commit arrives on w_empty in both CTAs)
              if (leader) ptx::mbar_expect_tx(&w_full[slot], 2 * nt * tap_bytes);
              const uint32_t bar0 = ptx::mapa_u32(&w_full[slot], 0);
              for (int t = 0; t < nt; ++t)
                ptx::tma_load_2d_2cta(smem_w + slot * args.slot_bytes + t * tap_bytes, tm, bar0,
                                      ((tap0 + t) * kch + kc) * 64, row);
            }
          }
          __syncwarp();
          if (++slot == args.n_slots) { slot = 0; phase ^= 1; }
        }
    }
    // The last layer's MMAs are done with the tiles and the ring: stage this CTA's head weights over them while the
    // epilogue warps are still busy (the Dense weights do not depend on the activations).
    ptx::mbar_wait(tower_done, 0, args.err_flag, ERR_HEADW);
    stage_head_weights(args, s_wp, s_wv, hw_full, lane, unit_first, units_in_chunk(0), value_staged(args.pf, args.vf, hg.hsp),
                       hslice * hg.gps, max(0, min(hg.gps, hg.groups_total - hslice * hg.gps)));
  } else if (warp == 3) {
    // ============ activation producer: one halo tile per 64-channel chunk, as soon as the CTAs that own those
    // channels of this board tile have published the previous layer (no grid-wide barrier: a tile's CTAs only ever
    // read what CTAs of the same tile wrote, and a chunk can start while other slices are still in their epilogue) ====
    const unsigned* tile_flags = tile_flag_line(args, 0, tile);   // lane 0 replaces it with the nearer copy during layer 0
    for (int l = 0; l < args.n_layers; ++l) {
      const int k = l == 0 ? args.stem_k : args.blk_k;
      const int kch = l == 0 ? args.stem_kchunks : args.filters / 64;
      const int hw = 8 + k - 1, pad = (k - 1) / 2;
      const int plane_bytes = hw * args.nb * hw * 128;
      const CUtensorMap* tm = (l & 1) ? &tm_xb : &tm_y;   // layer 0 builds its tile in place, see above
      const unsigned target = args.base + static_cast<unsigned>(l);   // layers 0..l-1
      if (l == 0) {
        // the stem's tile was built in shared memory above (this warp took part and has passed its barrier)
        if (lane == 0) {
          if (args.trace && blockIdx.x == trace_cta) args.trace[0] = clock64();
          if (leader) ptx::mbar_arrive(stem_ready);
          else ptx::mbar_arrive_cluster(stem_ready, 0);   // release.cluster: this CTA's tile is complete
          // while layer 0 computes: which of the tile's two flag lines answers faster from this SM?  Publish a value
          // to this slice's scratch word of each line and time until a poll sees it (two rounds, best of each).
          // which of the tile's two flag lines is the near one for this SM: from the table made at creation, else
          // (no table, unknown SM) measured here while layer 0 computes
          unsigned smid;
          asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
          const unsigned known = args.sm_near ? args.sm_near[smid] : 0xFFu;
          unsigned best[2] = {known == 1u ? 1u : 0u, known == 1u ? 0u : 1u};
          if (known > 1u) best[0] = best[1] = 0xffffffffu;
          for (unsigned round = 1; round <= 2 && known > 1u; ++round)
            for (int c = 0; c < 2; ++c) {
              unsigned* pw = tile_flag_line(args, c, tile) + 16 + slice;
              const unsigned val = args.base + round;
              const long long t0 = clock64();
              ptx::publish_flag_gpu(pw, val);
              unsigned spins = 0;
              while (ptx::ld_relaxed_gpu(pw) != val && ++spins < (1u << 16)) {}
              const unsigned d = static_cast<unsigned>(clock64() - t0);
              best[c] = d < best[c] ? d : best[c];
            }
          s_near_copy = best[1] < best[0] ? 1 : 0;
          tile_flags = tile_flag_line(args, s_near_copy, tile);
        }
        __syncwarp();
        continue;
      }
      if (lane == 0) {
        // (Requesting each plane as soon as the slices that produce ITS channels have published was tried: the planes
        // become ready in any order but are consumed in order, and every request needs its own acquire fence --
        // 5.9 us per layer instead of 4.85.)
        wait_tile(args, tile_flags, target);   // the tile's CTAs have published the previous layer
        ptx::fence_proxy_async_all();  // other CTAs' generic-proxy stores -> our async-proxy (TMA) reads
        if (args.trace && blockIdx.x == trace_cta) args.trace[l * 8 + 0] = clock64();
        for (int kc = 0; kc < kch; ++kc) {
          if constexpr (!PAIR) {
            ptx::mbar_expect_tx(&a_full[kc], plane_bytes);
            ptx::tma_load_4d(smem_a + kc * plane_bytes, tm, &a_full[kc], kc * 64, -pad, n0, -pad);
          } else {
            // each CTA of the pair fetches its own board's tile when ITS tile group is ready; the bytes of both count
            // on the leader's barrier
            if (leader) ptx::mbar_expect_tx(&a_full[kc], 2 * plane_bytes);
            ptx::tma_load_4d_2cta(smem_a + kc * plane_bytes, tm, ptx::mapa_u32(&a_full[kc], 0), kc * 64, -pad, n0, -pad);
          }
        }
      }
      __syncwarp();
      if (args.trace && blockIdx.x == trace_cta && leader) {   // when did the layer's last activation plane land (debug only)
        ptx::mbar_wait(&a_full[kch - 1], (l - 1) & 1, args.err_flag, ERR_A);   // plane's first TMA fill is layer 1's
        if (lane == 0) args.trace[l * 8 + 7] = clock64();
      }
    }
  } else if ((warp == 1 || warp == 2) && leader) {   // (pair mode: the leader issues for both CTAs)
    // ============ MMA issuers (two warps: a tcgen05.mma costs its issuing thread ~60 cycles whatever its shape,
    // so with N = 16..64 the issue rate, not the tensor pipe, bounds a layer) ============
    const int which = warp - 1;
    const uint32_t idesc = ptx::idesc_h16_f32(PAIR ? 128 : 64 * args.nb, args.ns, args.fp16);
    const uint64_t b_desc0 = desc_sw128(ptx::smem_u32(smem_w), 1024);
    const uint64_t tap_desc = static_cast<uint64_t>(tap_bytes >> 4);
    const uint64_t slot_desc = static_cast<uint64_t>(args.slot_bytes >> 4);
    const uint32_t tm0 = tmem_base + static_cast<uint32_t>(2 * which * ns_cta), tm1 = tm0 + static_cast<uint32_t>(ns_cta);
    int slot = 0;
    uint32_t phase = 0, a_phase = 0;
    uint64_t b_slot = b_desc0;
    for (int l = 0; l < args.n_layers; ++l) {
      const int k = l == 0 ? args.stem_k : args.blk_k;
      const int kch = l == 0 ? args.stem_kchunks : args.filters / 64;
      const int tps = taps_in_slot(k, ns_cta, args.slot_bytes);
      const int hw = 8 + k - 1;
      const uint64_t plane_desc = static_cast<uint64_t>((hw * args.nb * hw * 128) >> 4);
      const uint64_t rp8 = static_cast<uint64_t>(args.nb * hw * 8);  // halo rows between consecutive ranks, x 128 B / 16
      // next 8-row group = next (rank, board) = hw halo rows further on
      uint64_t a_plane = desc_sw128(ptx::smem_u32(smem_a), hw * 128);
      bool fresh = true;   // this issuer has not written its accumulators in this layer yet
      const int shape = (k == 3 && tps == 9) ? 0 : (k == 3 && tps == 3) ? 1 : (k == 5 && tps == 5) ? 2 : 3;
      for (int kc = 0; kc < kch; ++kc, a_plane += plane_desc) {
        // plane kc completes once per layer that uses it (the stem may use fewer planes than the tower)
        if (l == 0) {
          // the stem's tile was written by this CTA's (and, in pair mode, the peer's) threads, not by TMA
          ptx::mbar_wait_cluster(stem_ready, 0, args.err_flag, ERR_MMA_A);
        } else {
          ptx::mbar_wait(&a_full[kc], (a_phase >> kc) & 1u, args.err_flag, ERR_MMA_A);
          a_phase ^= 1u << kc;
        }
        if (args.trace && blockIdx.x == trace_cta && lane == 0 && kc == 0 && which == 0) args.trace[l * 8 + 1] = clock64();
        if (args.trace && blockIdx.x == trace_cta && lane == 0 && kc == kch - 1 && which == 0) args.trace[l * 8 + 2] = clock64();
        int dy0 = 0, dx0 = 0;
        for (int tap0 = 0; tap0 < k * k; tap0 += tps) {
          if ((kc & 1) == which) {   // issuer `which` takes the chunks with kc % 2 == which
            ptx::mbar_wait(&w_full[slot], phase, args.err_flag, ERR_MMA_W);
            ptx::tcgen05_fence_after();
            if (ptx::elect_one()) {
              if constexpr (!PAIR) {
                if (shape == 0) issue_taps<3, 9, false>(a_plane, b_slot, 0, 0, rp8, tap_desc, tm0, tm1, idesc, fresh);
                else if (shape == 1) issue_taps<3, 3, false>(a_plane, b_slot, dy0, 0, rp8, tap_desc, tm0, tm1, idesc, fresh);
                else if (shape == 2) issue_taps<5, 5, false>(a_plane, b_slot, dy0, 0, rp8, tap_desc, tm0, tm1, idesc, fresh);
                else
                  for (int t = 0, dy = dy0, dx = dx0; t < min(tps, k * k - tap0); ++t) {
                    issue_taps<1, 1, false>(a_plane, b_slot + static_cast<uint64_t>(t) * tap_desc, dy, dx, rp8, tap_desc, tm0, tm1,
                                            idesc, fresh && t == 0);
                    if (++dx == k) { dx = 0; ++dy; }
                  }
                ptx::mma_commit(&w_empty[slot]);
              } else {
                if (shape == 0) issue_taps<3, 9, true>(a_plane, b_slot, 0, 0, rp8, tap_desc, tm0, tm1, idesc, fresh);
                else if (shape == 1) issue_taps<3, 3, true>(a_plane, b_slot, dy0, 0, rp8, tap_desc, tm0, tm1, idesc, fresh);
                else if (shape == 2) issue_taps<5, 5, true>(a_plane, b_slot, dy0, 0, rp8, tap_desc, tm0, tm1, idesc, fresh);
                else
                  for (int t = 0, dy = dy0, dx = dx0; t < min(tps, k * k - tap0); ++t) {
                    issue_taps<1, 1, true>(a_plane, b_slot + static_cast<uint64_t>(t) * tap_desc, dy, dx, rp8, tap_desc, tm0, tm1,
                                           idesc, fresh && t == 0);
                    if (++dx == k) { dx = 0; ++dy; }
                  }
                ptx::mma_commit_2cta(&w_empty[slot]);   // frees the slot in both CTAs
              }
            }
            __syncwarp();
            fresh = false;
          }
          dx0 += tps;
          while (dx0 >= k) { dx0 -= k; ++dy0; }
          b_slot += slot_desc;
          if (++slot == args.n_slots) { slot = 0; phase ^= 1; b_slot = b_desc0; }
        }
      }
      if (args.trace && blockIdx.x == trace_cta && lane == 0 && which == 0) args.trace[l * 8 + 3] = clock64();
      if (ptx::elect_one()) {
        if constexpr (!PAIR) {
          ptx::mma_commit(acc_full);
          if (l == args.n_layers - 1) ptx::mma_commit(tower_done);
        } else {   // both CTAs' epilogues (and head-weight stagers) wait on their own copies
          ptx::mma_commit_2cta(acc_full);
          if (l == args.n_layers - 1) ptx::mma_commit_2cta(tower_done);
        }
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ============ epilogue: bias (+ fp32 residual) + ReLU, stores, grid barrier arrive ============
    // M = 128 (two boards): TMEM lane m = quad*32 + lane is row (rank m>>4, board (m>>3)&1, file m&7).
    // M = 64 (one board): the 64 rows sit in lanes 0-15 of each 32-lane quadrant, row i = quad*16 + lane.
    // Pair (cta_group::2, M = 128): this CTA's 64 rows sit in lanes 0-63 for the slice's first ns/2 columns and again in
    // lanes 64-127 for the other ns/2 (measured, tools/ub/ub_pair_layout.cu): row = (quad & 1) * 32 + lane, half = quad >> 1.
    const int quad = warp & 3;
    const int m = PAIR ? (quad & 1) * 32 + lane : (args.nb == 2 ? quad * 32 + lane : quad * 16 + (lane & 15));
    const int h = args.nb == 2 ? m >> 4 : m >> 3, p = args.nb == 2 ? (m >> 3) & 1 : 0, w = m & 7;
    const int pos = n0 + p;
    const bool live = pos < args.batch && (PAIR || args.nb == 2 || lane < 16);
    const int cb = PAIR ? (quad >> 1) * ns_cta : 0;   // this thread's first channel within the slice
    const size_t row = static_cast<size_t>(pos) * 64 + h * 8 + w;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int hf_all = args.pf + args.vf;
    for (int i = threadIdx.x - 128; i < HEAD_FILTERS_MAX * 128; i += 128) {
      const int f6 = i >> 7, j = i & 127;
      s_hw[i] = (f6 < hf_all && j < args.ns) ? args.head_w[f6 * args.filters + col0 + j] : 0.0f;
    }
    // The layer's shifts are staged one layer ahead, behind the epilogue's named barrier: fetched chunk by chunk right before
    // their use they cost an L2 round trip per 16-channel chunk (0.7 us per layer with four chunks, measured).
    if (threadIdx.x - 128 < args.ns) s_bias[0][threadIdx.x - 128] = args.bias_all[col0 + threadIdx.x - 128];
    asm volatile("bar.sync 1, 128;" ::: "memory");
    float res[RES][16];   // fp32 residual stream: this thread's row, this CTA's channels (up to 16 RES)
#pragma unroll
    for (int ci = 0; ci < RES; ++ci)
#pragma unroll
      for (int j = 0; j < 16; ++j) res[ci][j] = 0.0f;
    for (int l = 0; l < args.n_layers; ++l) {
      const bool last = l == args.n_layers - 1;   // its output only feeds the heads: no activation stores
      const bool conv1 = l > 0 && (l & 1);     // bf16 out to y only
      const bool has_res = l > 0 && !(l & 1);  // conv2: + residual, out to xf and xb
      const bool two_issuers = (l == 0 ? args.stem_kchunks : args.filters / 64) > 1;
      const float* bias = s_bias[l & 1] + cb;
      // next layer's shift of channel (thread - 128): requested now, in flight during the wait for the accumulators, parked in
      // shared memory right before the layer's barrier (requested there it put an L2 round trip in front of the release)
      const bool stage_bias = !last && threadIdx.x - 128 < args.ns;
      const float next_bias = stage_bias ? __ldg(args.bias_all + static_cast<size_t>(l + 1) * args.filters + col0 + threadIdx.x - 128) : 0.0f;
#pragma unroll
      for (int ci = 0; ci < RES; ++ci) {
        const int c0 = ci * 16;
        if (c0 >= ns_cta) break;
        // the shifts come from shared memory (staged a layer ahead); the fp32 residual stream of this row and these channels never leaves the
        // thread: it lives in registers from one block to the next (res)
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = bias[c0 + j];
        if (has_res) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] += res[ci][j];
        }
        if (c0 == 0) {
          ptx::mbar_wait(acc_full, l & 1, args.err_flag, ERR_EPI);
          ptx::tcgen05_fence_after();
          if (args.trace && blockIdx.x == trace_cta && threadIdx.x == 128) args.trace[l * 8 + 4] = clock64();
        }
        {
          // a single-chunk layer (the stem) is issued by issuer 0 alone: accumulators 2 and 3 were not written
          uint32_t v[NACC][16];
          ptx::tmem_ld_32x16(taddr + c0, v[0]);
          ptx::tmem_ld_32x16(taddr + ns_cta + c0, v[1]);
          if (two_issuers) {
            ptx::tmem_ld_32x16(taddr + 2 * ns_cta + c0, v[2]);
            ptx::tmem_ld_32x16(taddr + 3 * ns_cta + c0, v[3]);
          }
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float s01 = __uint_as_float(v[0][j]) + __uint_as_float(v[1][j]);
            const float s23 = two_issuers ? __uint_as_float(v[2][j]) + __uint_as_float(v[3][j]) : 0.0f;
            f[j] += s01 + s23;
          }
        }
        if (live && last) {
          // 1x1 head convolutions (model_chess.py:77,86): this thread's row, this 16-channel group -> one partial sum
          // per head filter; the groups are added up, in order, by the head phase below
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.0f);
          float* hp = args.hpart + ((static_cast<size_t>(pos) * 16 + ((col0 + cb + c0) >> 4)) * HEAD_FILTERS_MAX) * 64 + h * 8 + w;
#pragma unroll
          for (int f6 = 0; f6 < HEAD_FILTERS_MAX; ++f6) {
            if (f6 < hf_all) {
              const float4* wq = reinterpret_cast<const float4*>(s_hw + f6 * 128 + cb + c0);
              float acc = 0.0f;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 wv = wq[q];
                acc = fmaf(f[q * 4 + 0], wv.x, acc); acc = fmaf(f[q * 4 + 1], wv.y, acc);
                acc = fmaf(f[q * 4 + 2], wv.z, acc); acc = fmaf(f[q * 4 + 3], wv.w, acc);
              }
              hp[f6 * 64] = acc;
            }
          }
        } else if (live) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.0f);
          if (!conv1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) res[ci][j] = f[j];
          }
          __nv_bfloat16* orow = (conv1 ? args.y : args.xb) + row * args.filters + col0 + cb + c0;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              pk[e] = ptx::pack_h16x2(f[q * 8 + e * 2], f[q * 8 + e * 2 + 1], (args.fp16 & ptx::FMT_A_FP16) != 0);
            }
            *reinterpret_cast<uint4*>(orow + q * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      // Release of this layer's activations: the CTA barrier below orders every epilogue thread's stores before thread 128,
      // whose ONE gpu-scope fence is cumulative over them, then the flag.  (All 128 threads used to fence -- __threadfence
      // plus fence.proxy.async each: 0.85 us per layer where this costs 0.3, measured; the consumer's own fence.proxy.async
      // after its acquire orders its TMA reads.  CAZ_SB_DBG bit 2 brings the old sequence back for A/B runs.)
      if (args.dbg & 4) {
        __threadfence();
        ptx::fence_proxy_async_all();
      }
      ptx::tcgen05_fence_before();   // TMEM reads done before the next layer's MMAs overwrite the accumulators
      if (stage_bias) s_bias[(l + 1) & 1][threadIdx.x - 128] = next_bias;
      if (args.trace && blockIdx.x == trace_cta && threadIdx.x == 128) args.trace[l * 8 + 5] = clock64();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (args.trace && blockIdx.x == trace_cta && threadIdx.x == 128) args.trace[l * 8 + 6] = clock64();
      if (args.trace && threadIdx.x == 128 && l < 16) {
        long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        args.trace[1024 + (l * 256 + blockIdx.x) * 2] = gt;
      }
      if (threadIdx.x == 128) {
        if (!(args.dbg & 4)) ptx::fence_acq_rel_gpu();
        publish(args, tile, slice, args.base + 1u + static_cast<unsigned>(l));
      }
    }
  }
  if (args.trace && blockIdx.x == trace_cta && lane == 0) args.trace[200 + warp] = clock64();
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (args.trace && blockIdx.x == trace_cta && lane == 0) args.trace[216 + warp] = clock64();
  if constexpr (PAIR) ptx::cluster_sync_all();  // nobody frees TMEM while the peer may still signal or read it
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    if constexpr (PAIR) ptx::tmem_dealloc_2cta(tmem_base, tmem_cols);
    else ptx::tmem_dealloc(tmem_base, tmem_cols);
  }

  // ===================== heads, per board tile (no grid-wide step) =====================
  // H1: this tile's head-conv outputs from the 16 channel-group partials (+ folded BN shift, ReLU, Flatten order).
  // H2: this CTA's labels of the policy Dense layer and its hidden units of the value head, from staged weights;
  //     per-128-label softmax statistics and per-16-unit value partials are published for the tile.
  // H3: after the tile's second rendezvous every CTA normalises its own labels; slice 0 writes the value.
  // Measurement switch (CAZ_SB_DBG bit 3, tools/sb_trace.py): the heads run twice in one launch.  The second pass costs
  // what the first does, i.e. the heads are not bound by cold instruction or data caches.  (Its outputs are not valid:
  // the value phase reuses the policy weights' shared memory.)
#pragma unroll 1
  for (int hpass = 0; hpass < ((args.dbg & 8) ? 2 : 1); ++hpass) {
  long long* kt = hpass == 0 ? ktrace : (ktrace ? args.trace + 222 : nullptr);
  const int hf_total = args.pf + args.vf;
  const int kp = args.pf * 64, kv = args.vf * 64;
  if (kt) kt[2] = clock64();
  // requested now, used after H1: this thread's label biases and legal-move mask bytes, this slice's hidden-unit constants
  const int n_lchunks = (hg.upl * LABEL_UNIT + LABEL_CHUNK - 1) / LABEL_CHUNK;
  float pbias[LCHUNKS_MAX];
  bool pmask[LCHUNKS_MAX][2];
  const float vbias2 = args.vfc2_b[0];
#pragma unroll
  for (int c = 0; c < LCHUNKS_MAX; ++c) {
    const int n = lab0 + c * LABEL_CHUNK + threadIdx.x;
    const bool valid = c < n_lchunks && c * LABEL_CHUNK + threadIdx.x < lab_cnt;
    pbias[c] = valid ? args.pfc_b[n] : 0.0f;
#pragma unroll
    for (int bl = 0; bl < 2; ++bl)
      pmask[c][bl] = !(valid && args.mask && bl < hnb && hn0 + bl < args.batch) ||
                     args.mask[static_cast<size_t>(hn0 + bl) * args.n_labels + n] != 0;
  }
  // (kept in registers until H1's loads are on their way: a shared-memory store would wait for its load right here)
  const float vb_r = threadIdx.x < hid_cnt ? args.vfc1_b[hid0 + threadIdx.x] : 0.0f;
  const float vw_r = threadIdx.x < hid_cnt ? args.vfc2_w[hid0 + threadIdx.x] : 0.0f;
  const unsigned* tile_flags = tile_flag_line(args, s_near_copy, tile);   // written by warp 3 long before the __syncthreads above
  const unsigned* peer_flags = tile_flag_line(args, s_near_copy, tile ^ 1);   // (shared heads: the pair's other board)
  if (threadIdx.x == 0) {
    wait_tile(args, tile_flags, args.base + L);   // the tile's CTAs have stored their partials
    if (hshare) wait_tile(args, peer_flags, args.base + L);
  }
  __syncthreads();
  if (kt) kt[3] = clock64();
  {
    // one thread per (board, head filter, four squares): sixteen 16-byte loads, all requested before any is summed
    const int n_quads = hnb * hf_total * 16;   // <= 2 * 8 * 16 = THREADS
    const int t = threadIdx.x;
    const int q = t & 15, f6 = (t >> 4) % hf_total, bl = t / (16 * hf_total);
    const bool on = t < n_quads && hn0 + bl < args.batch;
    const float4* hp = reinterpret_cast<const float4*>(
        args.hpart + ((static_cast<size_t>(on ? hn0 + bl : 0) * 16) * HEAD_FILTERS_MAX + f6) * 64 + q * 4);
    float4 part[16];
#pragma unroll
    for (int g = 0; g < 16; ++g)
      part[g] = (on && g * 16 < args.filters) ? hp[static_cast<size_t>(g) * HEAD_FILTERS_MAX * 16] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float hb = on ? args.head_b[f6] : 0.0f;
    s_vb[threadIdx.x] = vb_r;
    s_vw[threadIdx.x] = vw_r;
    if (t < n_quads) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int g = 0; g < 16; ++g) { v.x += part[g].x; v.y += part[g].y; v.z += part[g].z; v.w += part[g].w; }
      v.x = on ? fmaxf(v.x + hb, 0.0f) : 0.0f; v.y = on ? fmaxf(v.y + hb, 0.0f) : 0.0f;
      v.z = on ? fmaxf(v.z + hb, 0.0f) : 0.0f; v.w = on ? fmaxf(v.w + hb, 0.0f) : 0.0f;
      // Keras channels-first Flatten: index c*64 + square (model_chess.py:80,89); policy filters first, then value
      *reinterpret_cast<float4*>(&s_feat[bl * 512 + (f6 < args.pf ? f6 * 64 : kp + (f6 - args.pf) * 64) + q * 4]) = v;
    }
  }
  __syncthreads();
  if (kt) kt[4] = clock64();

  // ---- H2 ----
  float lg[LCHUNKS_MAX][2];   // [label chunk][board of the tile]: this thread's logits (n_lchunks is checked on the host)
  const int unit_in_chunk = threadIdx.x >> 7;   // threads 0-127 / 128-255 = the chunk's first / second 128-label unit
#pragma unroll
  for (int c = 0; c < LCHUNKS_MAX; ++c) {
    if (c >= n_lchunks) break;
    if (c > 0) {
      __syncthreads();   // everyone is done with the previous chunk's weights
      if (warp == 0)
        stage_head_weights(args, s_wp, s_wv, hw_full, lane, unit_first + 2 * c, units_in_chunk(c), false, 0, 0);
    }
    ptx::mbar_wait(hw_full, c & 1, args.err_flag, ERR_HEADW);
    if (kt && c == 0) kt[5] = clock64();
    const int n = lab0 + c * LABEL_CHUNK + threadIdx.x;
    const bool valid = c * LABEL_CHUNK + threadIdx.x < lab_cnt;
    float a0 = 0.0f, a1 = 0.0f;
    if (valid) {
      const float bn = pbias[c];
      const float* wcol = s_wp + (threadIdx.x >> 7) * kp * LABEL_UNIT + (threadIdx.x & 127);
      // (the features are the same for every thread: one 16-byte broadcast load per four k)
      const float4* f4 = reinterpret_cast<const float4*>(s_feat);
#pragma unroll 4
      for (int k = 0; k < kp; k += 4) {
        const float4 fa = f4[k >> 2];
        a0 = fmaf(fa.x, wcol[k * LABEL_UNIT], a0);
        a0 = fmaf(fa.y, wcol[(k + 1) * LABEL_UNIT], a0);
        a0 = fmaf(fa.z, wcol[(k + 2) * LABEL_UNIT], a0);
        a0 = fmaf(fa.w, wcol[(k + 3) * LABEL_UNIT], a0);
      }
      if (hnb == 2) {
#pragma unroll 4
        for (int k = 0; k < kp; k += 4) {
          const float4 fa = f4[128 + (k >> 2)];
          a1 = fmaf(fa.x, wcol[k * LABEL_UNIT], a1);
          a1 = fmaf(fa.y, wcol[(k + 1) * LABEL_UNIT], a1);
          a1 = fmaf(fa.z, wcol[(k + 2) * LABEL_UNIT], a1);
          a1 = fmaf(fa.w, wcol[(k + 3) * LABEL_UNIT], a1);
        }
      }
      a0 += bn; a1 += bn;
    }
#pragma unroll
    for (int bl = 0; bl < 2; ++bl) {
      if (bl >= hnb) break;
      const int b = hn0 + bl;
      const bool on = valid && bl < hnb && b < args.batch && pmask[c][bl];
      const float x = on ? (bl ? a1 : a0) : -INFINITY;
      lg[c][bl] = x;
      // statistics of this 128-label unit: max, then the sum of exp(x - max), reduced in a fixed order
      float mx = x;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (lane == 0) s_red[warp] = mx;
      __syncthreads();
      mx = fmaxf(fmaxf(s_red[unit_in_chunk * 4], s_red[unit_in_chunk * 4 + 1]),
                 fmaxf(s_red[unit_in_chunk * 4 + 2], s_red[unit_in_chunk * 4 + 3]));
      float e = x == -INFINITY ? 0.0f : expf(x - mx);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
      __syncthreads();
      if (lane == 0) s_red[8 + warp] = e;
      __syncthreads();
      const int unit = hslice * hg.upl + c * (LABEL_CHUNK / LABEL_UNIT) + unit_in_chunk;
      if ((threadIdx.x & 127) == 0 && bl < hnb && b < args.batch && unit < LABEL_UNITS_MAX &&
          c * (LABEL_CHUNK / LABEL_UNIT) + unit_in_chunk < hg.upl) {
        const float se = (s_red[8 + unit_in_chunk * 4] + s_red[8 + unit_in_chunk * 4 + 1]) +
                         (s_red[8 + unit_in_chunk * 4 + 2] + s_red[8 + unit_in_chunk * 4 + 3]);
        float* st = args.hstats + static_cast<size_t>(b) * HSTAT_FLOATS + unit * 2;
        st[0] = mx;
        st[1] = se;
      }
    }
  }
  if (kt) kt[6] = clock64();
  // value head hidden layer (model_chess.py:91): unit u = hid0 + (t & (hsp-1))... every thread takes one (unit, k-eighth):
  // eight partial dot products per unit over k = kq, kq+8, ..., added in order
  {
    float* s_part = s_wp;   // the policy weights are consumed: [2 boards][8][hsp] partials
    __syncthreads();
    const int per = hg.hsp * 8;
    for (int i = threadIdx.x; i < per; i += THREADS) {
      const int u = i % hg.hsp, kq = i / hg.hsp;
      float a0 = 0.0f, a1 = 0.0f;
      const bool staged = value_staged(args.pf, args.vf, hg.hsp);
      if (u < hid_cnt)
        for (int k = kq; k < kv; k += 8) {
          const float wv = staged ? s_wv[(u >> 4) * value_group_pitch(args.vf) + k * HID_GROUP + (u & 15)]
                                  : __ldg(args.vfc1_w + static_cast<size_t>(k) * args.value_fc + hid0 + u);
          a0 = fmaf(s_feat[kp + k], wv, a0);
          if (hnb == 2) a1 = fmaf(s_feat[512 + kp + k], wv, a1);
        }
      s_part[i] = a0;
      s_part[per + i] = a1;
    }
    __syncthreads();
    // one thread per (board, hidden unit): hidden = relu(sum of the 8 partials + bias) times its output weight; the 16
    // products of a group are added by a fixed shuffle tree (whole warps take part: a group is half a warp)
    const int n_hid = hnb * hg.hsp;
    for (int i = threadIdx.x; i < ((n_hid + 31) & ~31); i += THREADS) {
      const int bl = i / hg.hsp, u = i - bl * hg.hsp, b = hn0 + bl;
      float pv = 0.0f;
      if (i < n_hid && u < hid_cnt) {
        float hsum = 0.0f;
#pragma unroll
        for (int kq = 0; kq < 8; ++kq) hsum += s_part[bl * per + kq * hg.hsp + u];
        pv = fmaxf(hsum + s_vb[u], 0.0f) * s_vw[u];
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) pv += __shfl_xor_sync(0xffffffffu, pv, o);
      const int grp = hslice * hg.gps + (u >> 4);
      if (i < n_hid && (u & 15) == 0 && b < args.batch && grp < hg.groups_total)
        args.hstats[static_cast<size_t>(b) * HSTAT_FLOATS + 2 * LABEL_UNITS_MAX + grp] = pv;
    }
  }
  if (args.dbg & 4) __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (!(args.dbg & 4)) ptx::fence_acq_rel_gpu();   // one cumulative fence behind the barrier (see the tower's release)
    publish(args, tile, slice, args.base + 1u + L);
    wait_tile(args, tile_flags, args.base + 1u + L);
    if (hshare) wait_tile(args, peer_flags, args.base + 1u + L);
  }
  __syncthreads();
  if (kt) kt[7] = clock64();

  // ---- H3: softmax over all of the board's labels (model_chess.py:83) from the units' statistics; tanh value ----
  // the tile's statistics come in with one L2 read per thread (ld.cg: other CTAs wrote them)
  if (threadIdx.x < 2 * HSTAT_FLOATS) {
    const int bl = threadIdx.x / HSTAT_FLOATS, j = threadIdx.x - bl * HSTAT_FLOATS;
    s_stat[threadIdx.x] = (bl < hnb && hn0 + bl < args.batch) ? ptx::ld_cg_f32(args.hstats + static_cast<size_t>(hn0 + bl) * HSTAT_FLOATS + j) : 0.0f;
  }
  __syncthreads();
  // one warp per board: the maximum over the units and the normaliser (one expf per lane, fixed shuffle trees)
  if (warp < hnb) {
    const float* st = s_stat + warp * HSTAT_FLOATS;
    const float mu = lane < hg.units_total ? st[2 * lane] : -INFINITY, su = lane < hg.units_total ? st[2 * lane + 1] : 0.0f;
    float mx = mu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float z = mu != -INFINITY ? su * expf(mu - mx) : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
    if (lane == 0) { s_red[2 * warp] = mx; s_red[2 * warp + 1] = z > 0.0f ? 1.0f / z : 0.0f; }
  }
  __syncthreads();
#pragma unroll
  for (int bl = 0; bl < 2; ++bl) {
    const int b = hn0 + bl;
    if (bl >= hnb || b >= args.batch) continue;
    const float* st = s_stat + bl * HSTAT_FLOATS;
    const float mx = s_red[2 * bl], inv = s_red[2 * bl + 1];
#pragma unroll
    for (int c = 0; c < LCHUNKS_MAX; ++c) {
      if (c >= n_lchunks) break;
      const int n = lab0 + c * LABEL_CHUNK + threadIdx.x;
      if (c * LABEL_CHUNK + threadIdx.x < lab_cnt) {
        const float x = lg[c][bl];
        args.policy[static_cast<size_t>(b) * args.n_labels + n] = x == -INFINITY ? 0.0f : expf(x - mx) * inv;
      }
    }
    if (hslice == 0 && threadIdx.x == 0) {
      float v = vbias2;
      for (int g = 0; g < hg.groups_total; ++g) v += st[2 * LABEL_UNITS_MAX + g];
      args.value[b] = tanhf(v);
    }
  }
  if (kt) kt[11] = clock64();
  }
  if (ktrace) {
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    ktrace[9] = gt;
    ktrace[10] = clock64();
  }
}

}  // namespace sb
}  // namespace caz
