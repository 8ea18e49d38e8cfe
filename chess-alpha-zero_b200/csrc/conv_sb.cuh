// Small-batch path: the WHOLE forward pass (input pack, stem, every residual conv, both heads) in ONE persistent,
// cooperatively launched kernel, for MCTS-sized batches (reference batches are <= 48 positions:
// search_threads x max_processes, configs/normal.py:30-31).
//
// At batch 16 a forward pass is ~17 GFLOP -- nothing for the tensor cores -- but every layer needs the 1.18 MB
// of its weights and has to wait for the previous layer: the pass is bound by latency and by L2->SM bandwidth.
// So:
//   * grid = (board pairs) x (output-channel slices): CTA (t, s) owns rows of boards 2t, 2t+1 and N_s = 256/S
//     output channels in EVERY layer (batch 16: 8 x 16 = 128 CTAs, N_s = 16).  Its residual slice and its bias
//     never leave the CTA; only the bf16 activations cross CTAs, through L2.
//   * weights: the bf16 weight set (17.4 MB) is L2-resident; each CTA streams its [N_s x K] slice through a
//     shared-memory ring with TMA, running AHEAD across layer boundaries (weights do not depend on the previous
//     layer), so after the first layer the weight fetch is off the critical path.  A ring slot holds several
//     filter taps, so the MMA warp waits once per slot and then issues its MMAs back to back.
//   * activations: ONE zero-padded halo tile per 64-channel chunk, [10 ranks][2 boards][10 files][64 ch] (TMA box
//     with out-of-bounds fill), serves all nine taps: tap (dy, dx) is the same tile read through a UMMA descriptor
//     whose start address is shifted by (20*dy + dx) rows and whose 8-row-group stride is 10 rows.  The 128-byte
//     swizzle is a function of the shared-memory address, so the shifted view stays consistent with what TMA
//     wrote.  Activation traffic per CTA and layer: 102 KB instead of 9 x 64 KB.
//   * consecutive tcgen05.mma into one accumulator serialise; with N = 16..64 an MMA is far shorter than that
//     latency, so the K loop rotates over NACC accumulators (TMEM column groups) that the epilogue adds up.
//   * phases are separated by a grid barrier (one release-add + acquire-spin on an L2 counter; cooperative launch
//     guarantees co-residency; the spin is bounded and traps instead of hanging).  (A cluster-per-board-pair
//     variant with DSMEM mbarriers was measured slower: 16-CTA clusters do not all fit at once and the 16 CTAs of
//     a cluster then pull the same activation tile through one GPC's port.)
//   * the input pack runs as phase 0 and the heads as three SIMT phases after the tower (1x1 convs per row;
//     policy logits / value hidden layer sliced over CTAs; softmax / tanh per position): no extra launches.
//     Those phases run once per launch, i.e. from a cold instruction cache: their loops are kept rolled.
// TMEM lane m of the accumulator is row (rank h = m>>4, board p = (m>>3)&1, file w = m&7).
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
constexpr int SCRATCH_FLOATS = (A_BYTES + W_RING_BYTES) / 4;  // what the head phases may use
constexpr int MAX_BATCH = 64;

struct Args {
  int batch;        // positions (<= MAX_BATCH)
  int n_layers;     // 1 + 2 * res_blocks
  int nb;           // boards per CTA tile: 2 (M = 128) or 1 (M = 64: half the A-operand bytes per MMA)
  int ns;           // output channels per CTA (16, 32 or 64)
  int n_slices;     // filters / ns
  int filters;
  int stem_k, stem_kchunks;  // stem kernel size, input channel chunks (in_cpad / 64)
  int blk_k;                  // tower kernel size
  int in_planes, in_cpad;
  int fp16;                   // 16-bit operands are fp16 instead of bf16
  int head_reps;              // 1; > 1 repeats the head phases (measurement aid)
  int slot_bytes, n_slots;    // weight ring geometry
  const float* planes;        // fp32 [batch][in_planes][64] or nullptr
  const uint8_t* records;     // u8 [batch][68] or nullptr (exactly one of the two)
  __nv_bfloat16* in_act;      // bf16 [batch][64][in_cpad]
  const float* bias_all;      // [n_layers][filters]
  float* xf;                  // fp32 residual stream   [batch*64][filters]
  __nv_bfloat16* xb;          // bf16 copy of it         (A operand of conv1)
  __nv_bfloat16* y;           // bf16 conv1 output       (A operand of conv2)
  // heads (fp32)
  int pf, vf, value_fc, n_labels;
  const float* head_w;        // [pf+vf][filters]
  const float* head_b;        // [pf+vf]
  float* featp;               // [batch][pf*64]
  float* featv;               // [batch][vf*64]
  const float* pfc_w;         // [pf*64][n_labels]
  const float* pfc_b;
  const float* vfc1_w;        // [vf*64][value_fc]
  const float* vfc1_b;
  const float* vfc2_w;        // [value_fc]
  const float* vfc2_b;
  float* vhid;                // [batch][value_fc]
  const uint8_t* mask;        // nullptr or [batch][n_labels]
  float* policy;              // [batch][n_labels]
  float* value;               // [batch]
  unsigned* flags;            // [gridDim.x] progress flag of every CTA: base + number of phases it has completed
  unsigned base;              // flag value that means "nothing of this launch done yet" (monotonic across launches)
  unsigned* counter;          // grid-wide barrier counter for the three head phases (one poller per CTA, one address)
  unsigned cbase;             // its value when this launch started
  long long* trace;           // nullptr or [n_layers+1][8] clock64 stamps of CTA 0 (see caz_debug_sb_trace)
  int* err_flag;
};

enum : int { ERR_W = 201, ERR_A = 202, ERR_GRID = 203, ERR_MMA_A = 204, ERR_MMA_W = 205, ERR_EPI = 206 };

// Grid phases: 0 input pack, 1..L the L conv layers, L+1 head 1x1 convs, L+2 head dense layers.
// A CTA that has packed its share of the input publishes flag = base + 1, after layer l flag = base + 2 + l.
constexpr int GRID_PHASES = 3;   // grid-wide arrivals per launch: tower done, head convs done, head dense layers done
constexpr int MAX_TILES = 64;   // flags: 16 words per board tile

// taps per weight-ring slot: the whole filter, one filter row, or one tap -- the largest that fits a slot
__host__ __device__ inline int taps_per_slot(int k, int ns) {
  if (k * k * ns * 128 <= W_SLOT_MAX_BYTES) return k * k;
  if (k * ns * 128 <= W_SLOT_MAX_BYTES) return k;
  return 1;
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

// Grid-wide barrier of the head phases (3 per launch): one release-add per CTA on one L2 counter, one polling thread
// per CTA.  (The tower layers do NOT use it: they synchronise per board tile through the progress flags.  Polling the
// flags of every tile from every CTA was measured to swamp the L2 slices that hold them.)
__device__ __forceinline__ void grid_wait(const Args& args, unsigned n) {
  const unsigned target = args.cbase + n * gridDim.x;
  uint32_t spins = 0;
  // relaxed polls (an acquire per poll is a fence per poll), one acquire fence once the count is there
  while (static_cast<int>(ptx::ld_relaxed_gpu(args.counter) - target) < 0)
    if (++spins > (1u << 26)) sync_trap(args);
  ptx::fence_acq_rel_gpu();
}

// all threads of the CTA: publish this CTA's stores, count it as arrived, wait for grid-wide arrival number n
__device__ __forceinline__ void grid_sync_all(const Args& args, unsigned n) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::red_relaxed_gpu_add(args.counter, 1u);
    grid_wait(args, n);
  }
  __syncthreads();
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
template <int K, int TPS>
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
    ptx::mma_bf16_ss(tm0, adesc, bdesc, idesc, accum);
    ptx::mma_bf16_ss(tm1, adesc + 2, bdesc + 2, idesc, accum);
    ptx::mma_bf16_ss(tm0, adesc + 4, bdesc + 4, idesc, 1u);
    ptx::mma_bf16_ss(tm1, adesc + 6, bdesc + 6, idesc, 1u);
  }
}

__global__ void __launch_bounds__(THREADS, 1)
forward_sb_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_xb,
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / args.n_slices, slice = blockIdx.x % args.n_slices;
  const int n0 = tile * args.nb, col0 = slice * args.ns;
  const int tap_bytes = args.ns * 128;
  const uint32_t tmem_cols = static_cast<uint32_t>(NACC * args.ns);  // 128, 256 or 512: powers of two
  const unsigned L = static_cast<unsigned>(args.n_layers);
  long long* ktrace = (args.trace && blockIdx.x == 0 && threadIdx.x == 32) ? args.trace + args.n_layers * 8 : nullptr;
  if (ktrace) ktrace[0] = clock64();

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_in);
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
    ptx::mbar_fence_init();
    ptx::fence_proxy_async();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ===================== phase 0: input pack =====================
  // fp32 NCHW planes (what predict_on_batch receives) or 68-byte board records -> bf16 NHWC, channels zero-padded.
  // The weight producer does not take part: it starts streaming layer 0's weights immediately.
  if (warp != 0) {
    const int per_pos = 64 * args.in_cpad;
    const int nthr = THREADS - 32, t = threadIdx.x - 32;
    // board b is packed by the n_slices CTAs of its own pair, CTA `slice` taking every n_slices-th element
    for (int b = n0; b < n0 + args.nb && b < args.batch; ++b) {
      __nv_bfloat16* dst = args.in_act + static_cast<size_t>(b) * per_pos;
#pragma unroll 1
      for (int i = slice * nthr + t; i < per_pos; i += nthr * args.n_slices) {
        const int sq = i / args.in_cpad, c = i - sq * args.in_cpad;
        float v = 0.0f;
        if (c < args.in_planes)
          v = args.planes ? args.planes[(static_cast<size_t>(b) * args.in_planes + c) * 64 + sq]
                          : record_plane_value(args.records + static_cast<size_t>(b) * 68, c, sq);
        reinterpret_cast<uint16_t*>(dst)[i] = ptx::pack_h16(v, args.fp16 != 0);
      }
    }
    __threadfence();
    ptx::fence_proxy_async_all();
    asm volatile("bar.sync 2, 224;" ::: "memory");
    if (threadIdx.x == 32) ptx::publish_flag_gpu(args.flags + tile * 16 + slice, args.base + 1u);
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
      const int tps = taps_per_slot(k, args.ns);
      const CUtensorMap* tm = l == 0 ? &tm_w_stem : &tm_w_tower;
      const int row = (l == 0 ? 0 : (l - 1) * args.filters) + col0;
      for (int kc = 0; kc < kch; ++kc)
        for (int tap0 = 0; tap0 < k * k; tap0 += tps) {
          ptx::mbar_wait(&w_empty[slot], phase ^ 1, args.err_flag, ERR_W);
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(&w_full[slot], tps * tap_bytes);
            for (int t = 0; t < tps; ++t)
              ptx::tma_load_2d(smem_w + slot * args.slot_bytes + t * tap_bytes, tm, &w_full[slot],
                               ((tap0 + t) * kch + kc) * 64, row);
          }
          __syncwarp();
          if (++slot == args.n_slots) { slot = 0; phase ^= 1; }
        }
    }
  } else if (warp == 3) {
    // ============ activation producer: one halo tile per 64-channel chunk, as soon as the CTAs that own those
    // channels of this board tile have published the previous layer (no grid-wide barrier: a tile's CTAs only ever
    // read what CTAs of the same tile wrote, and a chunk can start while other slices are still in their epilogue) ====
    const unsigned* tile_flags = args.flags + tile * 16;
    for (int l = 0; l < args.n_layers; ++l) {
      const int k = l == 0 ? args.stem_k : args.blk_k;
      const int kch = l == 0 ? args.stem_kchunks : args.filters / 64;
      const int hw = 8 + k - 1, pad = (k - 1) / 2;
      const int plane_bytes = hw * args.nb * hw * 128;
      const CUtensorMap* tm = l == 0 ? &tm_in : ((l & 1) ? &tm_xb : &tm_y);
      const unsigned target = args.base + 1u + static_cast<unsigned>(l);   // pack + layers 0..l-1
      if (lane == 0) {
        // one thread reads all of the tile's flags (independent loads: one L2 round trip per poll), then one acquire fence
        uint32_t spins = 0;
        for (;;) {
          // 16 flag words per tile (n_slices <= 16 of them in use), fetched as four independent 16-byte loads
          uint4 v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = q * 4 < args.n_slices ? ptx::ld_relaxed_gpu_v4(tile_flags + q * 4) : make_uint4(target, target, target, target);
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
        ptx::fence_acq_rel_gpu();
        ptx::fence_proxy_async_all();  // other CTAs' generic-proxy stores -> our async-proxy (TMA) reads
        if (args.trace && blockIdx.x == 0) args.trace[l * 8 + 0] = clock64();
        for (int kc = 0; kc < kch; ++kc) {
          ptx::mbar_expect_tx(&a_full[kc], plane_bytes);
          ptx::tma_load_4d(smem_a + kc * plane_bytes, tm, &a_full[kc], kc * 64, -pad, n0, -pad);
        }
      }
      __syncwarp();
      if (args.trace && blockIdx.x == 0) {   // when did the layer's last activation plane land (debug only)
        ptx::mbar_wait(&a_full[kch - 1], l == 0 ? 0u : (kch - 1 < args.stem_kchunks ? (l & 1) : ((l - 1) & 1)), args.err_flag, ERR_A);
        if (lane == 0) args.trace[l * 8 + 7] = clock64();
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ============ MMA issuers (two warps: a tcgen05.mma costs its issuing thread ~60 cycles whatever its shape,
    // so with N = 16..64 the issue rate, not the tensor pipe, bounds a layer) ============
    const int which = warp - 1;
    const uint32_t idesc = ptx::idesc_h16_f32(64 * args.nb, args.ns, args.fp16 != 0);
    const uint64_t b_desc0 = desc_sw128(ptx::smem_u32(smem_w), 1024);
    const uint64_t tap_desc = static_cast<uint64_t>(tap_bytes >> 4);
    const uint64_t slot_desc = static_cast<uint64_t>(args.slot_bytes >> 4);
    const uint32_t tm0 = tmem_base + static_cast<uint32_t>(2 * which * args.ns), tm1 = tm0 + static_cast<uint32_t>(args.ns);
    int slot = 0;
    uint32_t phase = 0, a_phase = 0;
    uint64_t b_slot = b_desc0;
    for (int l = 0; l < args.n_layers; ++l) {
      const int k = l == 0 ? args.stem_k : args.blk_k;
      const int kch = l == 0 ? args.stem_kchunks : args.filters / 64;
      const int tps = taps_per_slot(k, args.ns);
      const int hw = 8 + k - 1;
      const uint64_t plane_desc = static_cast<uint64_t>((hw * args.nb * hw * 128) >> 4);
      const uint64_t rp8 = static_cast<uint64_t>(args.nb * hw * 8);  // halo rows between consecutive ranks, x 128 B / 16
      // next 8-row group = next (rank, board) = hw halo rows further on
      uint64_t a_plane = desc_sw128(ptx::smem_u32(smem_a), hw * 128);
      bool fresh = true;   // this issuer has not written its accumulators in this layer yet
      const int shape = (k == 3 && tps == 9) ? 0 : (k == 3 && tps == 3) ? 1 : (k == 5 && tps == 5) ? 2 : 3;
      for (int kc = 0; kc < kch; ++kc, a_plane += plane_desc) {
        // plane kc completes once per layer that uses it (the stem may use fewer planes than the tower)
        ptx::mbar_wait(&a_full[kc], (a_phase >> kc) & 1u, args.err_flag, ERR_MMA_A);
        a_phase ^= 1u << kc;
        if (args.trace && blockIdx.x == 0 && lane == 0 && kc == 0 && which == 0) args.trace[l * 8 + 1] = clock64();
        if (args.trace && blockIdx.x == 0 && lane == 0 && kc == kch - 1 && which == 0) args.trace[l * 8 + 2] = clock64();
        int dy0 = 0, dx0 = 0;
        for (int tap0 = 0; tap0 < k * k; tap0 += tps) {
          if ((kc & 1) == which) {   // issuer `which` takes the chunks with kc % 2 == which
            ptx::mbar_wait(&w_full[slot], phase, args.err_flag, ERR_MMA_W);
            ptx::tcgen05_fence_after();
            if (ptx::elect_one()) {
              if (shape == 0) issue_taps<3, 9>(a_plane, b_slot, 0, 0, rp8, tap_desc, tm0, tm1, idesc, fresh);
              else if (shape == 1) issue_taps<3, 3>(a_plane, b_slot, dy0, 0, rp8, tap_desc, tm0, tm1, idesc, fresh);
              else if (shape == 2) issue_taps<5, 5>(a_plane, b_slot, dy0, 0, rp8, tap_desc, tm0, tm1, idesc, fresh);
              else
                for (int t = 0, dy = dy0, dx = dx0; t < tps; ++t) {
                  issue_taps<1, 1>(a_plane, b_slot + static_cast<uint64_t>(t) * tap_desc, dy, dx, rp8, tap_desc, tm0, tm1, idesc,
                                   fresh && t == 0);
                  if (++dx == k) { dx = 0; ++dy; }
                }
              ptx::mma_commit(&w_empty[slot]);
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
      if (args.trace && blockIdx.x == 0 && lane == 0 && which == 0) args.trace[l * 8 + 3] = clock64();
      if (ptx::elect_one()) ptx::mma_commit(acc_full);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ============ epilogue: bias (+ fp32 residual) + ReLU, stores, grid barrier arrive ============
    // M = 128 (two boards): TMEM lane m = quad*32 + lane is row (rank m>>4, board (m>>3)&1, file m&7).
    // M = 64 (one board): the 64 rows sit in lanes 0-15 of each 32-lane quadrant, row i = quad*16 + lane.
    const int quad = warp & 3;
    const int m = args.nb == 2 ? quad * 32 + lane : quad * 16 + (lane & 15);
    const int h = args.nb == 2 ? m >> 4 : m >> 3, p = args.nb == 2 ? (m >> 3) & 1 : 0, w = m & 7;
    const int pos = n0 + p;
    const bool live = pos < args.batch && (args.nb == 2 || lane < 16);
    const size_t row = static_cast<size_t>(pos) * 64 + h * 8 + w;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    for (int l = 0; l < args.n_layers; ++l) {
      const bool conv1 = l > 0 && (l & 1);     // bf16 out to y only
      const bool has_res = l > 0 && !(l & 1);  // conv2: + residual, out to xf and xb
      const bool two_issuers = (l == 0 ? args.stem_kchunks : args.filters / 64) > 1;
      const float* bias = args.bias_all + static_cast<size_t>(l) * args.filters + col0;
      for (int c0 = 0; c0 < args.ns; c0 += 16) {
        // bias and the residual slice (written by this very thread two layers ago) are fetched BEFORE the wait
        float f[16];
        float* xrow = args.xf + row * args.filters + col0 + c0;
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __ldg(bias + c0 + j);
        if (has_res && live) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 r = *reinterpret_cast<const float4*>(xrow + q * 4);
            f[q * 4 + 0] += r.x; f[q * 4 + 1] += r.y; f[q * 4 + 2] += r.z; f[q * 4 + 3] += r.w;
          }
        }
        if (c0 == 0) {
          ptx::mbar_wait(acc_full, l & 1, args.err_flag, ERR_EPI);
          ptx::tcgen05_fence_after();
          if (args.trace && blockIdx.x == 0 && threadIdx.x == 128) args.trace[l * 8 + 4] = clock64();
        }
        {
          // a single-chunk layer (the stem) is issued by issuer 0 alone: accumulators 2 and 3 were not written
          uint32_t v[NACC][16];
          ptx::tmem_ld_32x16(taddr + c0, v[0]);
          ptx::tmem_ld_32x16(taddr + args.ns + c0, v[1]);
          if (two_issuers) {
            ptx::tmem_ld_32x16(taddr + 2 * args.ns + c0, v[2]);
            ptx::tmem_ld_32x16(taddr + 3 * args.ns + c0, v[3]);
          }
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float s01 = __uint_as_float(v[0][j]) + __uint_as_float(v[1][j]);
            const float s23 = two_issuers ? __uint_as_float(v[2][j]) + __uint_as_float(v[3][j]) : 0.0f;
            f[j] += s01 + s23;
          }
        }
        if (live) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.0f);
          if (!conv1) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<float4*>(xrow + q * 4) = make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
          }
          __nv_bfloat16* orow = (conv1 ? args.y : args.xb) + row * args.filters + col0 + c0;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              pk[e] = ptx::pack_h16x2(f[q * 8 + e * 2], f[q * 8 + e * 2 + 1], args.fp16 != 0);
            }
            *reinterpret_cast<uint4*>(orow + q * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      __threadfence();               // this thread's stores visible at gpu scope ...
      ptx::fence_proxy_async_all();  // ... and ordered before later async-proxy reads
      ptx::tcgen05_fence_before();   // TMEM reads done before the next layer's MMAs overwrite the accumulators
      if (args.trace && blockIdx.x == 0 && threadIdx.x == 128) args.trace[l * 8 + 5] = clock64();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (args.trace && blockIdx.x == 0 && threadIdx.x == 128) args.trace[l * 8 + 6] = clock64();
      if (args.trace && threadIdx.x == 128 && l < 16) {
        long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        args.trace[1024 + (l * 256 + blockIdx.x) * 2] = gt;
      }
      if (threadIdx.x == 128) {
        ptx::publish_flag_gpu(args.flags + tile * 16 + slice, args.base + 2u + static_cast<unsigned>(l));
        if (l == args.n_layers - 1) ptx::red_relaxed_gpu_add(args.counter, 1u);   // the heads read every tile
      }
    }
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, tmem_cols);
  }

  // ===================== heads: three SIMT phases over the whole grid =====================
  // The tower's shared memory is free now: reuse it as fp32 scratch.
  float* scratch = reinterpret_cast<float*>(smem);
  const int hf_total = args.pf + args.vf;
  const int F = args.filters;
  for (int i = threadIdx.x; i < hf_total * F; i += THREADS) scratch[i] = args.head_w[i];
  if (ktrace) ktrace[2] = clock64();
  if (threadIdx.x == 0) grid_wait(args, 1);  // every CTA has stored the last layer's output
  __syncthreads();
  // ---- H1: 1x1 convs + folded BN + ReLU + Flatten (model_chess.py:77-81,86-90), one warp per board square ----
  for (int rep = 0; rep < args.head_reps; ++rep) {   // head_reps > 1 is a measurement aid (cold vs warm instruction cache)
  if (ktrace) ktrace[3] = clock64();
  for (int r = blockIdx.x * 8 + warp; r < args.batch * 64; r += gridDim.x * 8) {
    const int b = r >> 6, sq = r & 63;
    const float* x = args.xf + static_cast<size_t>(r) * F;
#pragma unroll 1
    for (int f = 0; f < hf_total; ++f) {
      float v = 0.0f;
#pragma unroll 1
      for (int c = lane; c < F; c += 32) v = fmaf(x[c], scratch[f * F + c], v);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) {
        v = fmaxf(v + args.head_b[f], 0.0f);
        if (f < args.pf) args.featp[static_cast<size_t>(b) * args.pf * 64 + f * 64 + sq] = v;
        else args.featv[static_cast<size_t>(b) * args.vf * 64 + (f - args.pf) * 64 + sq] = v;
      }
    }
  }
  }
  if (ktrace) ktrace[4] = clock64();
  grid_sync_all(args, 2);
  // ---- H2: policy logits (labels sliced over CTAs) and the value head's hidden layer (units sliced) ----
  for (int rep = 0; rep < args.head_reps; ++rep) {
    if (ktrace) ktrace[5] = clock64();
    __syncthreads();
    const int kp = args.pf * 64, kv = args.vf * 64;
    // labels / hidden units per CTA, rounded up to 4 so that every row slice is 16-byte aligned for cp.async
    const int ls = (((args.n_labels + gridDim.x - 1) / gridDim.x) + 3) & ~3;
    const int lab0 = blockIdx.x * ls;
    const int hs = (((args.value_fc + gridDim.x - 1) / gridDim.x) + 3) & ~3;
    const int h0 = blockIdx.x * hs;
    float* s_fp = scratch;                        // [batch][kp]
    float* s_fv = s_fp + args.batch * kp;         // [batch][kv]
    float* s_wv = s_fv + args.batch * kv;         // [kv][hs]  this CTA's columns of the value Dense kernel
    float* s_wp = s_wv + kv * hs;                 // [kp][ls]  this CTA's columns of the policy Dense kernel
    const bool staged = args.batch * (kp + kv) + kv * hs + kp * ls <= SCRATCH_FLOATS;   // else straight from L2
    const bool vec = (args.n_labels & 3) == 0 && (args.value_fc & 3) == 0;
    if (staged) {
      // asynchronous copies, all in flight at once: one L2 latency, no register staging
#pragma unroll 1
      for (int i = threadIdx.x; i < args.batch * kp; i += THREADS) ptx::cp_async_4(s_fp + i, args.featp + i);
#pragma unroll 1
      for (int i = threadIdx.x; i < args.batch * kv; i += THREADS) ptx::cp_async_4(s_fv + i, args.featv + i);
      if (vec) {
        const int lq = ls >> 2, hq = hs >> 2;
#pragma unroll 1
        for (int i = threadIdx.x; i < kp * lq; i += THREADS) {
          const int k = i / lq, n = lab0 + (i - k * lq) * 4;
          if (n < args.n_labels) ptx::cp_async_16(s_wp + k * ls + (n - lab0), args.pfc_w + static_cast<size_t>(k) * args.n_labels + n);
        }
#pragma unroll 1
        for (int i = threadIdx.x; i < kv * hq; i += THREADS) {
          const int k = i / hq, hh = h0 + (i - k * hq) * 4;
          if (hh < args.value_fc) ptx::cp_async_16(s_wv + k * hs + (hh - h0), args.vfc1_w + static_cast<size_t>(k) * args.value_fc + hh);
        }
      } else {
#pragma unroll 1
        for (int i = threadIdx.x; i < kp * ls; i += THREADS) {
          const int k = i / ls, n = lab0 + i % ls;
          if (n < args.n_labels) ptx::cp_async_4(s_wp + i, args.pfc_w + static_cast<size_t>(k) * args.n_labels + n);
        }
#pragma unroll 1
        for (int i = threadIdx.x; i < kv * hs; i += THREADS) {
          const int k = i / hs, hh = h0 + i % hs;
          if (hh < args.value_fc) ptx::cp_async_4(s_wv + i, args.vfc1_w + static_cast<size_t>(k) * args.value_fc + hh);
        }
      }
      if (ktrace) ktrace[8] = clock64();
      ptx::cp_async_wait_all();
      __syncthreads();
      if (ktrace) ktrace[9] = clock64();
    }
#pragma unroll 1
    for (int i = threadIdx.x; i < args.batch * ls; i += THREADS) {
      const int b = i / ls, j = i - b * ls, n = lab0 + j;   // consecutive threads -> consecutive labels
      if (n < args.n_labels) {
        float a0 = 0.f;
        if (staged) {
          const float* fp = s_fp + b * kp;
#pragma unroll 4
          for (int k = 0; k < kp; ++k) a0 = fmaf(fp[k], s_wp[k * ls + j], a0);
        } else {
          const float* fp = args.featp + static_cast<size_t>(b) * kp;
#pragma unroll 8
          for (int k = 0; k < kp; ++k) a0 = fmaf(fp[k], __ldg(args.pfc_w + static_cast<size_t>(k) * args.n_labels + n), a0);
        }
        args.policy[static_cast<size_t>(b) * args.n_labels + n] = a0 + args.pfc_b[n];
      }
    }
    if (ktrace) ktrace[10] = clock64();
#pragma unroll 1
    for (int i = threadIdx.x; i < args.batch * hs; i += THREADS) {
      const int b = i / hs, j = i - b * hs, hh = h0 + j;
      if (hh < args.value_fc) {
        float a0 = 0.f;
        if (staged) {
          const float* fv = s_fv + b * kv;
#pragma unroll 4
          for (int k = 0; k < kv; ++k) a0 = fmaf(fv[k], s_wv[k * hs + j], a0);
        } else {
          const float* fv = args.featv + static_cast<size_t>(b) * kv;
#pragma unroll 8
          for (int k = 0; k < kv; ++k) a0 = fmaf(fv[k], __ldg(args.vfc1_w + static_cast<size_t>(k) * args.value_fc + hh), a0);
        }
        args.vhid[static_cast<size_t>(b) * args.value_fc + hh] = fmaxf(a0 + args.vfc1_b[hh], 0.0f);
      }
    }
  }
  if (ktrace) ktrace[6] = clock64();
  grid_sync_all(args, 3);
  if (ktrace) ktrace[7] = clock64();
  // ---- H3: softmax (optionally legal-masked) and Dense(1, tanh), one CTA per position ----
  for (int b = blockIdx.x; b < args.batch; b += gridDim.x) {
    __shared__ float s_red[8];
    __shared__ float s_bc;
    float* prow = args.policy + static_cast<size_t>(b) * args.n_labels;
    const uint8_t* mrow = args.mask ? args.mask + static_cast<size_t>(b) * args.n_labels : nullptr;
    float x[PNJ];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < PNJ; ++j) {
      const int n = threadIdx.x + j * THREADS;
      const bool on = n < args.n_labels && (!mrow || mrow[n] != 0);
      x[j] = on ? prow[n] : -INFINITY;
      mx = fmaxf(mx, x[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = s_red[0];
      for (int wi = 1; wi < 8; ++wi) v = fmaxf(v, s_red[wi]);
      s_bc = v;
    }
    __syncthreads();
    mx = s_bc;
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < PNJ; ++j) {
      x[j] = x[j] == -INFINITY ? 0.0f : expf(x[j] - mx);
      sum += x[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = 0.0f;
      for (int wi = 0; wi < 8; ++wi) v += s_red[wi];
      s_bc = v;
    }
    __syncthreads();
    const float inv = s_bc > 0.0f ? 1.0f / s_bc : 0.0f;
#pragma unroll
    for (int j = 0; j < PNJ; ++j) {
      const int n = threadIdx.x + j * THREADS;
      if (n < args.n_labels) prow[n] = x[j] * inv;
    }
    // value: tanh(vhid . w2 + b2)
    float part = 0.0f;
    for (int hh = threadIdx.x; hh < args.value_fc; hh += THREADS)
      part = fmaf(args.vhid[static_cast<size_t>(b) * args.value_fc + hh], args.vfc2_w[hh], part);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = args.vfc2_b[0];
      for (int wi = 0; wi < 8; ++wi) v += s_red[wi];
      args.value[b] = tanhf(v);
    }
    __syncthreads();
  }
}

}  // namespace sb
}  // namespace caz
