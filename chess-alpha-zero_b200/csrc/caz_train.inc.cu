// Training step of the reference's `opt` worker on the GPU (included at the end of caz_abi.cu: one
// translation unit, so the tensor-core launch helpers above are in scope).
//
// Replaces  model.compile(Adam(), ['categorical_crossentropy', 'mean_squared_error'], loss_weights=[1.25, 1.0]) and one
// batch of model.fit(state_ary, [policy_ary, value_ary], batch_size=384)  (reference src/chess_zero/worker/optimize.py:79-103)
// on the graph of agent/model_chess.py:57-111: BatchNormalization in training mode, l2(1e-4) on every Conv2D and Dense
// kernel (configs/normal.py:46-68).  Parameters, Adam moments and gradients are fp32 vectors in the tensor order of
// caz_create (= Keras layer order); the gradient vector may live in caller-provided device memory so that a
// torch.distributed all-reduce (NCCL) can run on it between caz_train_forward_backward and caz_train_apply.
#include "train.cuh"

namespace {

struct TrainConv {
  int k = 0, cin = 0, cout = 0;
  size_t w = 0, gamma = 0, beta = 0, mmean = 0, mvar = 0;   // offsets into the flat parameter vector
  float *z = nullptr, *a = nullptr;                         // [rows][cout]: convolution output, activation after BN (+skip) + ReLU
  float *mean = nullptr, *invstd = nullptr;                 // batch statistics of this step
  // 16-bit builds: fp16 NHWC copy of the activation (A operand of the next layer's forward convolution) and K-major
  // 16-bit copies of the kernel for the forward (fp16) and the input-gradient (bf16, rotated + transposed) convolutions
  uint16_t *a16 = nullptr, *wf = nullptr, *wb = nullptr;
  // ... and the activation once more as the K-major bf16 operand [cout][cell][board] of the NEXT layer's weight-gradient GEMM
  // (written by bn_apply_tile_kernel in the forward pass; halo cells stay zero from the allocation)
  uint16_t* aT = nullptr;
  int cin_pad = 0;
  CUtensorMap tm_a16, tm_wf, tm_wb, tm_aT;
  CUtensorMap tm_wf_h, tm_wb_h;   // the same weight matrices in boxes of cout / 4 (cin / 4) rows: two N tiles per board group
};

}  // namespace

struct caz_trainer {
  caz_arch arch{};
  caz_train_config cfg{};
  int device = 0, precision = CAZ_PRECISION_FP32, max_batch = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  size_t n_params = 0;
  std::vector<size_t> tensor_off, tensor_len;   // per tensor of the caz_create order
  float *params = nullptr, *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr;
  bool grads_external = false;
  uint8_t* seg = nullptr;
  int64_t step = 0;
  std::vector<TrainConv> convs;     // stem, conv1 / conv2 of every block
  TrainConv pconv, vconv;           // heads' 1x1 convolutions (+ BN)
  size_t pfc_w = 0, pfc_b = 0, vfc1_w = 0, vfc1_b = 0, vfc2_w = 0, vfc2_b = 0;
  // batch-sized buffers
  float *x0 = nullptr, *planes = nullptr, *policy_t = nullptr, *value_t = nullptr;
  float *featp = nullptr, *logits = nullptr, *dlogits = nullptr, *probs = nullptr, *featv = nullptr, *hid = nullptr, *vpre = nullptr,
        *vout = nullptr, *dvpre = nullptr, *dhid = nullptr, *dfeatp = nullptr, *dfeatv = nullptr, *gemm_partial = nullptr;
  size_t gemm_partial_floats = 0;
  float *d_a = nullptr, *d_b = nullptr, *d_skip = nullptr, *dz = nullptr, *dzp = nullptr, *dzv = nullptr, *dap = nullptr, *dav = nullptr;
  float *w_flip = nullptr, *zero_bias = nullptr;
  double *partial = nullptr, *loss_acc = nullptr;
  float last_losses[4] = {0, 0, 0, 0};
  // the ~230 launches of a step are captured once per batch size and replayed as one CUDA graph (CAZ_TRAIN_NO_GRAPH=1: off)
  bool use_graph = true;
  std::vector<std::pair<int, cudaGraphExec_t>> graphs;
  // data-parallel overlap: ev_mid fires when the gradients of the upper half of the network (vector elements >= grad_split)
  // are complete, ev_end when all are -- the caller's all-reduce stream waits on them (caz_train_wait_gradients)
  cudaEvent_t ev_mid = nullptr, ev_end = nullptr;
  size_t grad_split = 0;
  bool capturing = false;
  bool profiling = false;
  float* conv_stats = nullptr;   // tensor-core build: BatchNormalization partial sums written by the convolution epilogues
  bool whole_tiles = false, no_mixed = false, wg_padded = false;   // CAZ_TRAIN_WHOLE_TILES / _NO_MIXED / _WG_PADDED (A/B switches, read at creation)
  bool mask_fp32 = false;   // CAZ_TRAIN_MASK_FP32: the backward passes read the ReLU mask from the fp32 activations
  bool no_tile = false;   // CAZ_TRAIN_NO_TILE: BatchNormalization passes and transposes as separate kernels (the fallback for odd shapes)
  std::vector<std::pair<const char*, cudaEvent_t>> prof;
  int pending_batch = 0;
  // 16-bit builds (tensor-core convolutions): forward fp16 x fp16, input and weight gradients bf16 x bf16, fp32 accumulation
  bool tc = false;
  int num_sms = 0, nb = 0;                 // nb: max_batch rounded up to 64 (board-minor extent of the weight-gradient operands)
  uint16_t* x0_16 = nullptr;               // [rows][32] fp16: the stem's input, planes zero-padded to one 64-byte swizzle row
  uint16_t* dz16 = nullptr;                // [rows][filters] bf16: A operand of the input-gradient convolution
  // [channels][padded cell][nb] bf16: K-major operands of the weight-gradient GEMM (3x3 layers; the stem's own pitch)
  __nv_bfloat16 *dzT = nullptr, *aT = nullptr, *dzT_stem = nullptr, *x0T = nullptr;
  float* wg_partial = nullptr;             // [tap][K split][n_out][n_in] partial products
  CUtensorMap tm_x0, tm_dz, tm_dzT, tm_dzT_stem, tm_aT, tm_x0T;
};

namespace {

int tfail(caz_trainer* t, int code, const std::string& msg) {
  if (t) t->err = msg;
  else {
    std::lock_guard<std::mutex> g(g_err_mutex);
    g_create_error = msg;
  }
  return code;
}

#define TR_TRY(t, expr)                                                                                      \
  do {                                                                                                       \
    cudaError_t e__ = (expr);                                                                                \
    if (e__ != cudaSuccess)                                                                                  \
      return tfail((t), CAZ_ERR_CUDA, fmt("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__)); \
  } while (0)

int tr_launch_ok(caz_trainer* t, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return tfail(t, CAZ_ERR_CUDA, fmt("launch of %s failed: %s", what, cudaGetErrorString(e)));
  if (t->profiling) {   // CAZ_TRAIN_PROFILE: an event behind every launch (group); caz_train_forward_backward_end prints the table
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) == cudaSuccess && cudaEventRecord(ev, t->stream) == cudaSuccess) t->prof.emplace_back(what, ev);
  }
  return CAZ_OK;
}

// CAZ_TRAIN_PROFILE=1: device time between consecutive launch groups of the last step, summed per kernel name, to stderr
void tr_profile_report(caz_trainer* t) {
  std::vector<std::pair<std::string, std::pair<double, int>>> rows;
  double total = 0.0;
  for (size_t i = 1; i < t->prof.size(); ++i) {
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, t->prof[i - 1].second, t->prof[i].second);
    total += ms;
    bool found = false;
    for (auto& r : rows)
      if (r.first == t->prof[i].first) { r.second.first += ms; r.second.second += 1; found = true; }
    if (!found) rows.push_back({t->prof[i].first, {ms, 1}});
  }
  std::sort(rows.begin(), rows.end(), [](const auto& a, const auto& b) { return a.second.first > b.second.first; });
  fprintf(stderr, "caz_train profile (one step, plain launches, %.1f us):\n", total * 1e3);
  for (auto& r : rows) fprintf(stderr, "  %9.1f us %4d x %7.1f us  %s\n", r.second.first * 1e3, r.second.second, r.second.first * 1e3 / r.second.second, r.first.c_str());
  for (auto& pr : t->prof) cudaEventDestroy(pr.second);
  t->prof.clear();
}

template <typename T>
int tr_alloc(caz_trainer* t, T** p, size_t count) {
  TR_TRY(t, cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
  TR_TRY(t, cudaMemsetAsync(*p, 0, count * sizeof(T), t->stream));
  return CAZ_OK;
}

constexpr int HEAD_WGRAD_SPLITS = 128;   // row slices of the heads' 1x1 weight gradients
unsigned blocks_for(size_t n, int per = 256) { return static_cast<unsigned>((n + per - 1) / per); }

// "same" convolution on fp32 NHWC rows, raw output (no bias, no activation): out = conv(in, w) (+ add)
int tr_conv_fp32(caz_trainer* t, const float* in, const float* w, const float* add, float* out, int batch, int k, int cin, int cout) {
  const size_t smem = static_cast<size_t>(8 + k - 1) * (8 + k - 1) * cin * sizeof(float);
  switch (k) {
    case 1: caz::conv_fp32_kernel<1><<<batch, 256, smem, t->stream>>>(in, w, t->zero_bias, add, out, cin, cin, cout, 0); break;
    case 3: caz::conv_fp32_kernel<3><<<batch, 256, smem, t->stream>>>(in, w, t->zero_bias, add, out, cin, cin, cout, 0); break;
    case 5: caz::conv_fp32_kernel<5><<<batch, 256, smem, t->stream>>>(in, w, t->zero_bias, add, out, cin, cin, cout, 0); break;
    case 7: caz::conv_fp32_kernel<7><<<batch, 256, smem, t->stream>>>(in, w, t->zero_bias, add, out, cin, cin, cout, 0); break;
    default: return tfail(t, CAZ_ERR_INVALID, "unsupported kernel size");
  }
  return tr_launch_ok(t, "conv_fp32_kernel");
}

// C[m][n] = A . B (+ bias) with arbitrary strides (see gemm_kernel); ksplit > 1 (no bias, no relu): K in ksplit slices whose
// partial products go through t->gemm_partial and are summed in a fixed order
int tr_gemm(caz_trainer* t, const float* a, long a_si, long a_sk, const float* b, long b_sk, long b_sj, const float* bias, float* c,
            int m, int n, int k, int relu = 0, int ksplit = 1) {
  const size_t count = static_cast<size_t>(m) * n;
  if (ksplit > 1 && count * ksplit > t->gemm_partial_floats) return tfail(t, CAZ_ERR_INVALID, "tr_gemm: split-K scratch too small");
  const dim3 grid((n + 31) / 32, (m + 31) / 32, ksplit);
  caz::train::gemm_kernel<<<grid, 256, 0, t->stream>>>(a, a_si, a_sk, b, b_sk, b_sj, ksplit > 1 ? nullptr : bias,
                                                       ksplit > 1 ? t->gemm_partial : c, m, n, k, 0, ksplit > 1 ? 0 : relu);
  int rc;
  if ((rc = tr_launch_ok(t, "gemm_kernel")) || ksplit <= 1) return rc;
  caz::train::sum_partials_kernel<<<blocks_for(count), 256, 0, t->stream>>>(t->gemm_partial, ksplit, count, c);
  return tr_launch_ok(t, "sum_partials_kernel");
}

// 16-bit builds, tower layers: the BatchNormalization passes run as tile kernels that also write the K-major operands of the
// weight-gradient GEMMs (train.cuh); needs 64-channel tiles and float2 slices of the parameter vectors
bool tr_tile_path(const caz_trainer* t, const TrainConv& L) {
  return t->tc && L.a16 && (L.cout & 63) == 0 && (L.gamma & 3) == 0 && (L.beta & 3) == 0 && !t->no_tile;
}

// batch statistics of z (and the moving averages), then a = relu(BN(z) (+ skip))
int tr_bn_forward(caz_trainer* t, TrainConv& L, const float* skip, int rows) {
  int rc;
  if (t->conv_stats && L.a16) {
    // tower layer of the tensor-core build: the convolution's epilogue has left the partial sums (ConvArgs::stats)
    caz::train::bn_finish_stats_f32_kernel<<<(L.cout + 7) / 8, 256, 0, t->stream>>>(t->conv_stats, ((rows / 64 + 1) / 2) * 4, rows, L.cout,
                                                                                      t->arch.bn_epsilon, L.mean, L.invstd, t->grads + L.mmean,
                                                                                      t->grads + L.mvar);
    if ((rc = tr_launch_ok(t, "bn_finish_stats_f32_kernel"))) return rc;
  } else {
    const dim3 grid((L.cout + caz::train::STAT_CH - 1) / caz::train::STAT_CH, caz::train::STAT_SPLITS);
    caz::train::channel_sums_kernel<<<grid, caz::train::STAT_THREADS, 0, t->stream>>>(0, L.z, nullptr, nullptr, nullptr, nullptr, nullptr, rows, L.cout, t->partial);
    if ((rc = tr_launch_ok(t, "channel_sums_kernel"))) return rc;
    caz::train::bn_finish_stats_kernel<<<(L.cout + 7) / 8, 256, 0, t->stream>>>(t->partial, rows, L.cout, t->arch.bn_epsilon, L.mean, L.invstd,
                                                                                  t->grads + L.mmean, t->grads + L.mvar);
    if ((rc = tr_launch_ok(t, "bn_finish_stats_kernel"))) return rc;
  }
  const size_t total = static_cast<size_t>(rows) * L.cout;
  const int vec4 = (L.cout & 3) == 0 && (L.gamma & 3) == 0 && (L.beta & 3) == 0;   // float4 slices of the parameter vectors
  if (tr_tile_path(t, L)) {
    const int pad = (t->arch.block_kernel - 1) / 2;
    const dim3 tg(t->nb / 64, L.cout / 64, 64);
    caz::train::bn_apply_tile_kernel<<<tg, 256, 0, t->stream>>>(L.z, skip, t->params + L.gamma, t->params + L.beta, L.mean, L.invstd, rows / 64, L.cout,
                                                               t->nb, pad, L.a, reinterpret_cast<__half*>(L.a16),
                                                               reinterpret_cast<__nv_bfloat16*>(L.aT));
    return tr_launch_ok(t, "bn_apply_tile_kernel");
  }
  caz::train::bn_apply_kernel<__half><<<blocks_for(vec4 ? total / 4 : total), 256, 0, t->stream>>>(L.z, skip, t->params + L.gamma, t->params + L.beta, L.mean,
                                                                                                   L.invstd, total, L.cout, L.a,
                                                                                                   reinterpret_cast<__half*>(L.a16), vec4);
  return tr_launch_ok(t, "bn_apply_kernel");
}

// da (gradient wrt the layer's activation) -> dz (wrt its convolution output), dgamma / dbeta; dskip = gradient passed to the Add's other input
int tr_bn_backward(caz_trainer* t, TrainConv& L, const float* da, float* dz, float* dskip, int dskip_accumulate, int rows,
                   const float* da2 = nullptr, uint16_t* dz16 = nullptr, __nv_bfloat16* dzT = nullptr) {
  const dim3 grid((L.cout + caz::train::STAT_CH - 1) / caz::train::STAT_CH, caz::train::STAT_SPLITS);
  const bool tile_path = dzT && !dskip_accumulate;   // (the caller checked tr_tile_path)
  caz::train::channel_sums_kernel<<<grid, caz::train::STAT_THREADS, 0, t->stream>>>(1, L.z, L.a, da, da2, L.mean, L.invstd, rows, L.cout, t->partial,
                                                                                    tile_path && !t->mask_fp32 ? reinterpret_cast<const __half*>(L.a16) : nullptr);
  int rc;
  if ((rc = tr_launch_ok(t, "channel_sums_kernel (backward)"))) return rc;
  caz::train::bn_param_grads_kernel<<<(L.cout + 7) / 8, 256, 0, t->stream>>>(t->partial, L.cout, t->grads + L.gamma, t->grads + L.beta);
  if ((rc = tr_launch_ok(t, "bn_param_grads_kernel"))) return rc;
  const size_t total = static_cast<size_t>(rows) * L.cout;
  const int vec4 = (L.cout & 3) == 0 && (L.gamma & 3) == 0 && (L.beta & 3) == 0;
  if (tile_path) {   // no fp32 dz, the transposed operand instead
    const dim3 tg(t->nb / 64, L.cout / 64, 64);
    caz::train::bn_backward_tile_kernel<<<tg, 256, 0, t->stream>>>(L.z, L.a, t->mask_fp32 ? nullptr : reinterpret_cast<const __half*>(L.a16), da, da2, t->params + L.gamma, L.mean, L.invstd, t->grads + L.beta,
                                                                  t->grads + L.gamma, rows / 64, L.cout, t->nb, (L.k - 1) / 2,
                                                                  reinterpret_cast<__nv_bfloat16*>(dz16), dzT, dskip);
    return tr_launch_ok(t, "bn_backward_tile_kernel");
  }
  caz::train::bn_backward_kernel<__nv_bfloat16><<<blocks_for(vec4 ? total / 4 : total), 256, 0, t->stream>>>(L.z, L.a, da, da2, t->params + L.gamma, L.mean, L.invstd,
                                                                                          t->grads + L.beta, t->grads + L.gamma, rows, L.cout, dz, reinterpret_cast<__nv_bfloat16*>(dz16), dskip,
                                                                                          dskip_accumulate, vec4);
  return tr_launch_ok(t, "bn_backward_kernel");
}

// dW of a k x k convolution (fp32 build)
int tr_wgrad_fp32(caz_trainer* t, const float* x, const float* dz, int batch, const TrainConv& L) {
  const int taps = L.k * L.k;
  int splits = 1;
  while (splits < 16 && taps * splits * ((L.cin + 31) / 32) * ((L.cout + 31) / 32) < 4 * 148 && splits * 2 <= batch) splits *= 2;
  float* dw = t->grads + L.w;
  if (splits > 1) TR_TRY(t, cudaMemsetAsync(dw, 0, static_cast<size_t>(taps) * L.cin * L.cout * sizeof(float), t->stream));
  const dim3 grid((L.cout + 31) / 32, (L.cin + 31) / 32, taps * splits);
  caz::train::conv_wgrad_fp32_kernel<<<grid, 256, 0, t->stream>>>(x, dz, batch, L.cin, L.cout, L.k, splits, dw);
  return tr_launch_ok(t, "conv_wgrad_fp32_kernel");
}

// din = conv(dz, rot180(w)^T) (+ add): the input gradient of a "same" convolution
int tr_dgrad_fp32(caz_trainer* t, const float* dz, const TrainConv& L, const float* add, float* din, int batch) {
  const size_t total = static_cast<size_t>(L.k) * L.k * L.cin * L.cout;
  caz::train::flip_weights_kernel<<<blocks_for(total), 256, 0, t->stream>>>(t->params + L.w, t->w_flip, L.k * L.k, L.cin, L.cout);
  int rc;
  if ((rc = tr_launch_ok(t, "flip_weights_kernel"))) return rc;
  return tr_conv_fp32(t, dz, t->w_flip, add, din, batch, L.k, L.cout, L.cin);
}

// ---- tensor-core convolutions of the 16-bit builds (conv_tc2_kernel, see conv_tc2.cuh) ------------------------------------
int tr_tmap(caz_trainer* t, CUtensorMap* m, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
            int row_bytes = 128) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return tfail(t, CAZ_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return tfail(t, CAZ_ERR_CUDA, fmt("cuTensorMapEncodeTiled failed with CUresult %d", (int)r));
  return CAZ_OK;
}

// halo view of 16-bit NHWC rows [boards][8][8][c] (see make_halo_tmap)
int tr_halo_tmap(caz_trainer* t, CUtensorMap* m, void* base, int c, int k, int boards, int row_bytes = 128) {
  const cuuint32_t hw = 8 + k - 1;
  cuuint64_t dims[4] = {(cuuint64_t)c, 8, (cuuint64_t)boards, 8};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)c * 128, (cuuint64_t)c * 16};
  cuuint32_t box[4] = {(cuuint32_t)row_bytes / 2, hw, 2, hw};
  return tr_tmap(t, m, base, 4, dims, strides, box, row_bytes);
}

// K-major matrix [rows][kdim] 16-bit, boxes of (row_bytes / 2) x box_rows
int tr_kmajor_tmap(caz_trainer* t, CUtensorMap* m, void* base, size_t kdim, int rows, int box_rows, int row_bytes = 128) {
  cuuint64_t dims[2] = {(cuuint64_t)kdim, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kdim * 2};
  cuuint32_t box[2] = {(cuuint32_t)row_bytes / 2, (cuuint32_t)box_rows};
  return tr_tmap(t, m, base, 2, dims, strides, box, row_bytes);
}

int tr_launch_tc(caz_trainer* t, const CUtensorMap& tm_a, const CUtensorMap& tm_b, caz::tc::ConvArgs& args, int work, const char* what,
                 const CUtensorMap* tm_b_whole = nullptr) {
  args.err_flag = nullptr;
  const int max_pairs = t->num_sms / 2;
  const int grid = 2 * (work < max_pairs ? work : max_pairs);
  caz::tc2::conv_tc2_kernel<<<grid, caz::tc2::THREADS, caz::tc2::SMEM_BYTES, t->stream>>>(tm_a, tm_b, tm_b_whole ? *tm_b_whole : tm_b, args);
  return tr_launch_ok(t, what);
}

// out[rows][n] = conv(in16, w16): "same" convolution as an implicit GEMM on the tensor cores, raw fp32 row-major output
int tr_conv_tc(caz_trainer* t, const CUtensorMap& tm_in, const CUtensorMap& tm_w, const CUtensorMap& tm_w_half, int ksize, int row_bytes,
               int kchunks, int n, int fmt, float* out, int batch, bool stats = false) {
  caz::tc::ConvArgs args{};
  args.batch = batch;
  args.num_tiles = (batch + 1) / 2;
  args.ksize = ksize;
  args.row_bytes = row_bytes;
  args.kchunks = kchunks;
  args.n = n;
  args.bias = t->zero_bias;
  args.fp16 = fmt;
  args.n_tiles = 1;
  args.n_total = n;
  args.m_rows = batch * 64;
  args.out_rm = out;
  if (stats) {   // per-channel sum and sum of squares of the output, per epilogue warp (ConvArgs::stats)
    args.stats = t->conv_stats;
    args.stats_slots = args.num_tiles * 4;
  }
  // Two N tiles of n / 2 channels per board group when that fills the last wave better: batch 384 = 96 groups on 74 CTA pairs
  // is two waves of whole tiles but 2.6 waves of half tiles (= 1.3 tile times)
  // ... or MIXED: whole tiles for as many full rounds as the groups fill, half tiles for the rest (ConvArgs::mixed_whole).
  // Measured tile times on 74 pairs: whole 12.5 us, half 8.65 us (an N = 128 MMA sweep is not half an N = 256 one).
  const int pairs = (args.num_tiles + 1) / 2, maxp = t->num_sms / 2;
  const double t_whole = 12.5, t_half = 8.65;
  const int full = pairs / maxp * maxp;   // groups that make full rounds of whole tiles
  const double c_whole = (pairs + maxp - 1) / maxp * t_whole, c_half = (2 * pairs + maxp - 1) / maxp * t_half,
               c_mixed = full / maxp * t_whole + (2 * (pairs - full) + maxp - 1) / maxp * t_half;
  if (n % 64 == 0 && !t->whole_tiles && (c_half < c_whole || c_mixed < c_whole)) {
    args.n_tiles = 2;
    args.n = n / 2;
    args.mixed_whole = (c_mixed < c_half && !t->no_mixed) ? full : 0;
    return tr_launch_tc(t, tm_in, tm_w_half, args, args.mixed_whole + 2 * (pairs - args.mixed_whole),
                        args.mixed_whole ? "conv_tc2_kernel (training, whole + half N tiles)" : "conv_tc2_kernel (training, two N tiles)", &tm_w);
  }
  return tr_launch_tc(t, tm_in, tm_w, args, pairs, "conv_tc2_kernel (training)");
}

// dW[tap][ic][oc] of a k x k convolution: x rows [batch * 64][cin] (fp32), dz rows [batch * 64][cout] (fp32) -> t->grads + L.w
// x / dz == nullptr: that operand has already been written in K-major form (tile kernels)
int tr_wgrad_tc(caz_trainer* t, const float* x, const float* dz, int batch, const TrainConv& L, const CUtensorMap& tm_dzT, const CUtensorMap& tm_xT) {
  const int pad = (L.k - 1) / 2, pitch = 8 + 2 * pad, cells = pitch * pitch, taps = L.k * L.k;
  const int n_in_pad = L.cin < 32 ? 32 : L.cin;
  // K-major bf16 operands over the zero-padded board index (transpose_pad_kernel); the halo cells were zeroed at allocation
  // and are never written (the stem's layout, pitch 8 + 2 pad, has its own buffers)
  {
    const dim3 g1(t->nb / 64, (L.cout + 63) / 64, 64), g2(t->nb / 64, (n_in_pad + 63) / 64, 64);
    const bool stem = &L == &t->convs[0];
    if (dz) caz::train::transpose_pad_kernel<<<g1, 256, 0, t->stream>>>(dz, batch, L.cout, L.cout, t->nb, pad, stem ? t->dzT_stem : t->dzT);
    if (x) caz::train::transpose_pad_kernel<<<g2, 256, 0, t->stream>>>(x, batch, L.cin, n_in_pad, t->nb, pad, stem ? t->x0T : t->aT);
    int rc;
    if ((dz || x) && (rc = tr_launch_ok(t, "transpose_pad_kernel"))) return rc;
  }
  // K over the real cells only (ConvArgs::wg_cpc): dz^T is zero on the halo cells
  const bool compact = !t->wg_padded;
  const int chunks = (compact ? 64 : cells) * t->nb / 64;
  int splits = (t->num_sms / 2) / taps;
  if (splits < 1) splits = 1;
  if (splits > 16) splits = 16;
  const int cps = (chunks + splits - 1) / splits;
  caz::tc::ConvArgs args{};
  args.batch = 4;            // M = 256 rows = four "boards" of 64 output channels (missing ones read as zeros)
  args.num_tiles = 2;
  args.ksize = 1;
  args.row_bytes = 128;
  args.kchunks = cps;
  args.n = n_in_pad;
  args.bias = t->zero_bias;
  args.fp16 = 0;             // bf16 x bf16
  args.n_tiles = 1;
  args.n_total = n_in_pad;
  args.m_rows = L.cout;
  args.out_rm = t->wg_partial;
  args.wg_splits = splits;
  args.wg_k = L.k;
  args.wg_pitch = pitch * t->nb;
  args.wg_dx = t->nb;
  args.wg_cpc = compact ? t->nb / 64 : 0;
  int rc;
  if ((rc = tr_launch_tc(t, tm_dzT, tm_xT, args, taps * splits, "conv_tc2_kernel (weight gradient)"))) return rc;
  const dim3 gr((L.cin + 31) / 32, (L.cout + 31) / 32, taps);
  caz::train::wgrad_reduce_kernel<<<gr, 256, 0, t->stream>>>(t->wg_partial, taps, splits, L.cout, L.cin, n_in_pad, t->grads + L.w);
  return tr_launch_ok(t, "wgrad_reduce_kernel");
}

// the 16-bit copies of the convolution kernels follow the fp32 master weights (after creation and after every update)
int tr_refresh_weights16(caz_trainer* t) {
  if (!t->tc) return CAZ_OK;
  if (t->convs.size() > caz::train::PACK_MAX_LAYERS) return tfail(t, CAZ_ERR_INVALID, "too many convolution layers for the weight packing table");
  caz::train::PackTable tab{};
  for (size_t l = 0; l < t->convs.size(); ++l) {
    TrainConv& L = t->convs[l];
    tab.layer[l] = caz::train::PackLayer{t->params + L.w, L.wf, l == 0 ? nullptr : L.wb, L.k * L.k, L.cin, L.cin_pad, L.cout};
  }
  const dim3 grid(64, static_cast<unsigned>(t->convs.size()));
  caz::train::pack_all_conv_weights_kernel<<<grid, 256, 0, t->stream>>>(tab, 1, 0);
  return tr_launch_ok(t, "pack_all_conv_weights_kernel");
}

int tr_forward_backward(caz_trainer* t, int batch) {
  const caz_arch& a = t->arch;
  const int rows = batch * 64, F = a.filters, pf = a.policy_filters, vf = a.value_filters, nl = a.n_labels, vfc = a.value_fc;
  const int kp = pf * 64, kv = vf * 64;
  int rc;
  float* P = t->params;
  float* G = t->grads;
  // ---- forward ----
  {
    const size_t smem = static_cast<size_t>(a.input_planes) * 65 * sizeof(float);
    caz::pack_planes_kernel<float><<<batch, 256, smem, t->stream>>>(t->planes, t->x0, batch, a.input_planes, a.input_planes);
    if ((rc = tr_launch_ok(t, "pack_planes_kernel"))) return rc;
  }
  const int n_convs = static_cast<int>(t->convs.size());
  if (t->tc) {
    const size_t smem = static_cast<size_t>(a.input_planes) * 65 * sizeof(float);
    caz::pack_planes_kernel<__half><<<batch, 256, smem, t->stream>>>(t->planes, reinterpret_cast<__half*>(t->x0_16), batch, a.input_planes, 32);
    if ((rc = tr_launch_ok(t, "pack_planes_kernel (fp16)"))) return rc;
  }
  for (int l = 0; l < n_convs; ++l) {
    TrainConv& L = t->convs[l];
    const float* in = l == 0 ? t->x0 : t->convs[l - 1].a;
    if (t->tc) {
      const int fmt = caz::ptx::FMT_A_FP16 | caz::ptx::FMT_B_FP16;
      if (l == 0) rc = tr_conv_tc(t, t->tm_x0, L.tm_wf, L.tm_wf_h, L.k, 64, 1, L.cout, fmt, L.z, batch, t->conv_stats != nullptr);
      else rc = tr_conv_tc(t, t->convs[l - 1].tm_a16, L.tm_wf, L.tm_wf_h, L.k, 128, L.cin / 64, L.cout, fmt, L.z, batch, t->conv_stats != nullptr);
      if (rc) return rc;
    } else if ((rc = tr_conv_fp32(t, in, P + L.w, nullptr, L.z, batch, L.k, L.cin, L.cout))) return rc;
    const bool conv2 = l > 0 && (l & 1) == 0;
    if ((rc = tr_bn_forward(t, L, conv2 ? t->convs[l - 2].a : nullptr, rows))) return rc;
  }
  const float* top = t->convs[n_convs - 1].a;
  // heads' 1x1 convolutions as GEMMs over the rows
  caz::train::head_conv_forward_kernel<<<blocks_for(static_cast<size_t>(rows) * 32), 256, 0, t->stream>>>(top, P + t->pconv.w, P + t->vconv.w, t->pconv.z,
                                                                                                      t->vconv.z, rows, F, pf, vf);
  if ((rc = tr_launch_ok(t, "head_conv_forward_kernel"))) return rc;
  if ((rc = tr_bn_forward(t, t->pconv, nullptr, rows))) return rc;
  if ((rc = tr_bn_forward(t, t->vconv, nullptr, rows))) return rc;
  caz::train::flatten_kernel<<<blocks_for(static_cast<size_t>(rows) * pf), 256, 0, t->stream>>>(t->pconv.a, t->featp, batch, pf, 0);
  caz::train::flatten_kernel<<<blocks_for(static_cast<size_t>(rows) * vf), 256, 0, t->stream>>>(t->vconv.a, t->featv, batch, vf, 0);
  if ((rc = tr_launch_ok(t, "flatten_kernel"))) return rc;
  if ((rc = tr_gemm(t, t->featp, kp, 1, P + t->pfc_w, nl, 1, P + t->pfc_b, t->logits, batch, nl, kp))) return rc;
  if ((rc = tr_gemm(t, t->featv, kv, 1, P + t->vfc1_w, vfc, 1, P + t->vfc1_b, t->hid, batch, vfc, kv, 1))) return rc;
  if ((rc = tr_gemm(t, t->hid, vfc, 1, P + t->vfc2_w, 1, 1, P + t->vfc2_b, t->vpre, batch, 1, vfc))) return rc;
  // ---- losses and their gradients ----
  TR_TRY(t, cudaMemsetAsync(t->loss_acc, 0, 4 * sizeof(double), t->stream));
  caz::train::loss_kernel<<<batch, 256, 0, t->stream>>>(t->logits, t->policy_t, t->vpre, t->value_t, nl, t->cfg.policy_loss_weight, t->cfg.value_loss_weight,
                                                        batch, t->dlogits, t->dvpre, t->probs, t->vout, t->loss_acc);
  if ((rc = tr_launch_ok(t, "loss_kernel"))) return rc;
  // ---- backward: heads ----
  // policy Dense: dW = featp^T . dlogits, db = column sums, dfeatp = dlogits . W^T
  if ((rc = tr_gemm(t, t->featp, 1, kp, t->dlogits, nl, 1, nullptr, G + t->pfc_w, kp, nl, batch))) return rc;
  caz::train::column_sum_kernel<<<(nl + 31) / 32, 256, 0, t->stream>>>(t->dlogits, batch, nl, G + t->pfc_b);
  if ((rc = tr_gemm(t, t->dlogits, nl, 1, P + t->pfc_w, 1, nl, nullptr, t->dfeatp, batch, kp, nl, 0, 8))) return rc;
  // value head: Dense(1, tanh) <- Dense(vfc, relu)
  if ((rc = tr_gemm(t, t->hid, 1, vfc, t->dvpre, 1, 1, nullptr, G + t->vfc2_w, vfc, 1, batch))) return rc;
  caz::train::column_sum_kernel<<<1, 256, 0, t->stream>>>(t->dvpre, batch, 1, G + t->vfc2_b);
  if ((rc = tr_gemm(t, t->dvpre, 1, 1, P + t->vfc2_w, 1, 1, nullptr, t->dhid, batch, vfc, 1))) return rc;
  caz::train::relu_backward_kernel<<<blocks_for(static_cast<size_t>(batch) * vfc), 256, 0, t->stream>>>(t->hid, t->dhid, static_cast<size_t>(batch) * vfc);
  if ((rc = tr_gemm(t, t->featv, 1, kv, t->dhid, vfc, 1, nullptr, G + t->vfc1_w, kv, vfc, batch))) return rc;
  caz::train::column_sum_kernel<<<(vfc + 31) / 32, 256, 0, t->stream>>>(t->dhid, batch, vfc, G + t->vfc1_b);
  if ((rc = tr_gemm(t, t->dhid, vfc, 1, P + t->vfc1_w, 1, vfc, nullptr, t->dfeatv, batch, kv, vfc))) return rc;
  if ((rc = tr_launch_ok(t, "head backward"))) return rc;
  // un-flatten, BN + ReLU backward of the head convolutions, their weight gradients and the gradient wrt the tower's output
  caz::train::flatten_kernel<<<blocks_for(static_cast<size_t>(rows) * pf), 256, 0, t->stream>>>(t->dap, t->dfeatp, batch, pf, 1);
  caz::train::flatten_kernel<<<blocks_for(static_cast<size_t>(rows) * vf), 256, 0, t->stream>>>(t->dav, t->dfeatv, batch, vf, 1);
  if ((rc = tr_bn_backward(t, t->pconv, t->dap, t->dzp, nullptr, 0, rows))) return rc;
  if ((rc = tr_bn_backward(t, t->vconv, t->dav, t->dzv, nullptr, 0, rows))) return rc;
  {
    const dim3 grid((F + 255) / 256, HEAD_WGRAD_SPLITS);
    caz::train::head_conv_backward_weight_kernel<<<grid, 256, 0, t->stream>>>(top, t->dzp, t->dzv, t->gemm_partial, rows, F, pf, vf);
    if ((rc = tr_launch_ok(t, "head_conv_backward_weight_kernel"))) return rc;
    caz::train::head_conv_weight_reduce_kernel<<<blocks_for(static_cast<size_t>(F) * (pf + vf)), 256, 0, t->stream>>>(
        t->gemm_partial, HEAD_WGRAD_SPLITS, F, pf, vf, G + t->pconv.w, G + t->vconv.w);
    if ((rc = tr_launch_ok(t, "head_conv_weight_reduce_kernel"))) return rc;
  }
  float* dcur = t->d_a;   // gradient wrt the activation of the layer being processed
  float* dalt = t->d_b;
  caz::train::head_conv_backward_data_kernel<<<blocks_for(static_cast<size_t>(rows) * F), 256, 0, t->stream>>>(t->dzp, t->dzv, P + t->pconv.w, P + t->vconv.w,
                                                                                                           dcur, rows, F, pf, vf);
  if ((rc = tr_launch_ok(t, "head_conv_backward_data_kernel"))) return rc;
  // ---- backward: tower ----
  for (int l = n_convs - 1; l >= 0; --l) {
    TrainConv& L = t->convs[l];
    const bool conv2 = l > 0 && (l & 1) == 0, conv1 = (l & 1) == 1;
    // conv2: the Add hands dy to the block's input as well (kept in d_skip until the gradient wrt that input is complete).
    // fp32 build: conv1's input-gradient convolution adds d_skip itself; 16-bit builds: the layer below reads it as da2.
    const float* da2 = (t->tc && (l & 1) == 0 && l < n_convs - 1) ? t->d_skip : nullptr;
    const bool tile = tr_tile_path(t, L);   // dz leaves the BatchNormalization pass as dz16 + dzT; the layer below left its aT
    const bool in_tile = l > 0 && tr_tile_path(t, t->convs[l - 1]) && t->convs[l - 1].aT;
    if ((rc = tr_bn_backward(t, L, dcur, t->dz, conv2 ? t->d_skip : nullptr, 0, rows, da2, t->tc ? t->dz16 : nullptr,
                             tile ? (l == 0 ? t->dzT_stem : t->dzT) : nullptr))) return rc;
    const float* in = l == 0 ? t->x0 : t->convs[l - 1].a;
    if (t->tc) rc = tr_wgrad_tc(t, in_tile ? nullptr : in, tile ? nullptr : t->dz, batch, L, l == 0 ? t->tm_dzT_stem : t->tm_dzT,
                                l == 0 ? t->tm_x0T : (in_tile ? t->convs[l - 1].tm_aT : t->tm_aT));
    else rc = tr_wgrad_fp32(t, in, t->dz, batch, L);
    if (rc) return rc;
    if (l == 0) break;
    if (t->tc) rc = tr_conv_tc(t, t->tm_dz, L.tm_wb, L.tm_wb_h, L.k, 128, L.cout / 64, L.cin, 0, dalt, batch);
    else rc = tr_dgrad_fp32(t, t->dz, L, conv1 ? t->d_skip : nullptr, dalt, batch);
    if (rc) return rc;
    std::swap(dcur, dalt);
    if (l == n_convs / 2 && t->ev_mid)   // everything from this layer upwards (and the heads) has its gradients
      TR_TRY(t, cudaEventRecordWithFlags(t->ev_mid, t->stream, t->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
  }
  if (n_convs < 3 && t->ev_mid)   // no upper half to speak of: "mid" = end
    TR_TRY(t, cudaEventRecordWithFlags(t->ev_mid, t->stream, t->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
  return CAZ_OK;
}

}  // namespace

extern "C" {

void caz_train_destroy(caz_trainer* t) {
  if (!t) return;
  cudaSetDevice(t->device);
  if (t->stream) cudaStreamSynchronize(t->stream);
  std::vector<void*> ptrs = {t->params, t->adam_m, t->adam_v, t->seg, t->x0, t->planes, t->policy_t, t->value_t, t->featp, t->logits, t->dlogits,
                             t->probs, t->featv, t->hid, t->vpre, t->vout, t->dvpre, t->dhid, t->dfeatp, t->dfeatv, t->d_a, t->d_b, t->d_skip, t->dz,
                             t->dzp, t->dzv, t->dap, t->dav, t->w_flip, t->zero_bias, t->partial, t->loss_acc, t->gemm_partial, t->conv_stats};
  if (!t->grads_external) ptrs.push_back(t->grads);
  for (void* p : {(void*)t->x0_16, (void*)t->dz16, (void*)t->dzT, (void*)t->aT, (void*)t->dzT_stem, (void*)t->x0T, (void*)t->wg_partial}) ptrs.push_back(p);
  auto free_conv = [&](TrainConv& L) {
    for (void* p : {(void*)L.z, (void*)L.a, (void*)L.mean, (void*)L.invstd, (void*)L.a16, (void*)L.wf, (void*)L.wb, (void*)L.aT}) ptrs.push_back(p);
  };
  for (auto& L : t->convs) free_conv(L);
  free_conv(t->pconv);
  free_conv(t->vconv);
  for (void* p : ptrs) if (p) cudaFree(p);
  for (auto& g : t->graphs) cudaGraphExecDestroy(g.second);
  if (t->ev_mid) cudaEventDestroy(t->ev_mid);
  if (t->ev_end) cudaEventDestroy(t->ev_end);
  if (t->stream) cudaStreamDestroy(t->stream);
  delete t;
}

const char* caz_train_last_error(const caz_trainer* t) {
  if (t) return t->err.c_str();
  return caz_last_error(nullptr);
}

int64_t caz_train_param_count(const caz_arch* arch) {
  if (!arch) return 0;
  const caz_arch& a = *arch;
  int64_t n = (int64_t)a.stem_kernel * a.stem_kernel * a.input_planes * a.filters + 4 * a.filters;
  n += (int64_t)2 * a.res_blocks * ((int64_t)a.block_kernel * a.block_kernel * a.filters * a.filters + 4 * a.filters);
  n += (int64_t)a.filters * a.policy_filters + 4 * a.policy_filters + (int64_t)a.policy_filters * 64 * a.n_labels + a.n_labels;
  n += (int64_t)a.filters * a.value_filters + 4 * a.value_filters + (int64_t)a.value_filters * 64 * a.value_fc + a.value_fc + a.value_fc + 1;
  return n;
}

int caz_train_create(const caz_arch* arch, const caz_tensor* tensors, int32_t n_tensors, int32_t device, int32_t precision,
                     int32_t max_batch, const caz_train_config* cfg, float* d_grads, caz_trainer** out) {
  if (!arch || !tensors || !out || !cfg) return tfail(nullptr, CAZ_ERR_INVALID, "null argument");
  *out = nullptr;
  const caz_arch& a = *arch;
  if (precision != CAZ_PRECISION_FP32 && precision != CAZ_PRECISION_BF16)
    return tfail(nullptr, CAZ_ERR_INVALID, "caz_train_create: precision must be CAZ_PRECISION_FP32 or CAZ_PRECISION_BF16");
  if (precision == CAZ_PRECISION_BF16 && (a.filters % 64 != 0 || a.filters > 256 || a.input_planes > 32 || a.block_kernel != 3 || a.stem_kernel > 5))
    return tfail(nullptr, CAZ_ERR_INVALID, "the tensor-core training build needs filters in {64, 128, 192, 256}, <= 32 input planes, 3x3 blocks and a stem of at most 5x5");
  if (max_batch < 1 || n_tensors != expected_tensor_count(a)) return tfail(nullptr, CAZ_ERR_INVALID, "caz_train_create: bad batch size or tensor count");
  if (a.policy_filters + a.value_filters > caz::HEAD_MAX_F || (a.stem_kernel & 1) == 0 || (a.block_kernel & 1) == 0 || a.stem_kernel > 7 || a.block_kernel > 7)
    return tfail(nullptr, CAZ_ERR_INVALID, "architecture outside the supported envelope");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
    return tfail(nullptr, CAZ_ERR_NO_DEVICE, fmt("no CUDA device %d; this library has no CPU path", device));
  caz_trainer* t = new caz_trainer();
  t->arch = a;
  t->cfg = *cfg;
  t->device = device;
  t->precision = precision;
  t->max_batch = max_batch;
  auto bail = [&](int rc) {
    tfail(nullptr, rc, t->err);
    caz_train_destroy(t);
    return rc;
  };
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking) != cudaSuccess) {
    t->err = "cudaSetDevice / cudaStreamCreate failed";
    return bail(CAZ_ERR_CUDA);
  }
  // ---- flat parameter layout in the tensor order of caz_create ----
  std::vector<uint8_t> seg;
  std::vector<float> host;
  int ti = 0;
  auto take = [&](int64_t expect, uint8_t kind) -> size_t {   // returns the offset, SIZE_MAX on a size mismatch
    if (tensors[ti].numel != expect) return SIZE_MAX;
    const size_t off = host.size();
    host.insert(host.end(), tensors[ti].data, tensors[ti].data + expect);
    seg.insert(seg.end(), static_cast<size_t>(expect), kind);
    t->tensor_off.push_back(off);
    t->tensor_len.push_back(static_cast<size_t>(expect));
    ++ti;
    return off;
  };
  bool ok = true;
  auto take_conv = [&](TrainConv& L, int k, int cin, int cout) {
    L.k = k; L.cin = cin; L.cout = cout;
    ok &= (L.w = take((int64_t)k * k * cin * cout, 2)) != SIZE_MAX;
    ok &= (L.gamma = take(cout, 1)) != SIZE_MAX;
    ok &= (L.beta = take(cout, 1)) != SIZE_MAX;
    ok &= (L.mmean = take(cout, 3)) != SIZE_MAX;
    ok &= (L.mvar = take(cout, 3)) != SIZE_MAX;
  };
  t->convs.resize(1 + 2 * a.res_blocks);
  take_conv(t->convs[0], a.stem_kernel, a.input_planes, a.filters);
  for (int i = 1; ok && i < (int)t->convs.size(); ++i) take_conv(t->convs[i], a.block_kernel, a.filters, a.filters);
  if (ok) {
    take_conv(t->pconv, 1, a.filters, a.policy_filters);
    ok &= (t->pfc_w = take((int64_t)a.policy_filters * 64 * a.n_labels, 2)) != SIZE_MAX;
    ok &= (t->pfc_b = take(a.n_labels, 1)) != SIZE_MAX;
  }
  if (ok) {
    take_conv(t->vconv, 1, a.filters, a.value_filters);
    ok &= (t->vfc1_w = take((int64_t)a.value_filters * 64 * a.value_fc, 2)) != SIZE_MAX;
    ok &= (t->vfc1_b = take(a.value_fc, 1)) != SIZE_MAX;
    ok &= (t->vfc2_w = take(a.value_fc, 2)) != SIZE_MAX;
    ok &= (t->vfc2_b = take(1, 1)) != SIZE_MAX;
  }
  if (!ok) { t->err = fmt("tensor %d has an unexpected size", ti); return bail(CAZ_ERR_INVALID); }
  t->n_params = host.size();
  int rc;
#define TC(expr) do { if ((rc = (expr))) return bail(rc); } while (0)
  TC(tr_alloc(t, &t->params, t->n_params));
  TC(tr_alloc(t, &t->adam_m, t->n_params));
  TC(tr_alloc(t, &t->adam_v, t->n_params));
  TC(tr_alloc(t, &t->seg, t->n_params));
  if (d_grads) { t->grads = d_grads; t->grads_external = true; }
  else TC(tr_alloc(t, &t->grads, t->n_params));
  if (cudaMemcpyAsync(t->params, host.data(), host.size() * 4, cudaMemcpyHostToDevice, t->stream) != cudaSuccess ||
      cudaMemcpyAsync(t->seg, seg.data(), seg.size(), cudaMemcpyHostToDevice, t->stream) != cudaSuccess ||
      cudaMemsetAsync(t->grads, 0, t->n_params * 4, t->stream) != cudaSuccess || cudaStreamSynchronize(t->stream) != cudaSuccess) {
    t->err = "uploading the parameters failed";
    return bail(CAZ_ERR_CUDA);
  }
  // ---- batch-sized buffers ----
  const size_t B = (size_t)max_batch, R = B * 64, F = (size_t)a.filters;
  auto alloc_conv = [&](TrainConv& L) -> int {
    int r;
    if ((r = tr_alloc(t, &L.z, R * L.cout))) return r;
    if ((r = tr_alloc(t, &L.a, R * L.cout))) return r;
    if ((r = tr_alloc(t, &L.mean, (size_t)L.cout))) return r;
    return tr_alloc(t, &L.invstd, (size_t)L.cout);
  };
  for (auto& L : t->convs) TC(alloc_conv(L));
  TC(alloc_conv(t->pconv));
  TC(alloc_conv(t->vconv));
  const size_t kp = (size_t)a.policy_filters * 64, kv = (size_t)a.value_filters * 64;
  TC(tr_alloc(t, &t->planes, B * a.input_planes * 64));
  TC(tr_alloc(t, &t->x0, R * a.input_planes));
  TC(tr_alloc(t, &t->policy_t, B * a.n_labels));
  TC(tr_alloc(t, &t->value_t, B));
  TC(tr_alloc(t, &t->featp, B * kp));
  TC(tr_alloc(t, &t->dfeatp, B * kp));
  t->gemm_partial_floats = std::max(8 * B * kp, (size_t)HEAD_WGRAD_SPLITS * F * (a.policy_filters + a.value_filters));
  TC(tr_alloc(t, &t->gemm_partial, t->gemm_partial_floats));
  TC(tr_alloc(t, &t->logits, B * a.n_labels));
  TC(tr_alloc(t, &t->dlogits, B * a.n_labels));
  TC(tr_alloc(t, &t->probs, B * a.n_labels));
  TC(tr_alloc(t, &t->featv, B * kv));
  TC(tr_alloc(t, &t->dfeatv, B * kv));
  TC(tr_alloc(t, &t->hid, B * a.value_fc));
  TC(tr_alloc(t, &t->dhid, B * a.value_fc));
  TC(tr_alloc(t, &t->vpre, B));
  TC(tr_alloc(t, &t->vout, B));
  TC(tr_alloc(t, &t->dvpre, B));
  TC(tr_alloc(t, &t->d_a, R * F));
  TC(tr_alloc(t, &t->d_b, R * F));
  TC(tr_alloc(t, &t->d_skip, R * F));
  TC(tr_alloc(t, &t->dz, R * F));
  TC(tr_alloc(t, &t->dzp, R * a.policy_filters));
  TC(tr_alloc(t, &t->dzv, R * a.value_filters));
  TC(tr_alloc(t, &t->dap, R * a.policy_filters));
  TC(tr_alloc(t, &t->dav, R * a.value_filters));
  const size_t wmax = std::max((size_t)a.block_kernel * a.block_kernel * F * F, (size_t)a.stem_kernel * a.stem_kernel * a.input_planes * F);
  TC(tr_alloc(t, &t->w_flip, wmax));
  TC(tr_alloc(t, &t->zero_bias, std::max(F, (size_t)a.input_planes) + 64));
  TC(tr_alloc(t, &t->partial, (size_t)2 * caz::train::STAT_SPLITS * std::max(F, (size_t)64)));
  TC(tr_alloc(t, &t->loss_acc, (size_t)8));   // [0] CE, [1] MSE of the step; [4] sum of squares of the l2-regularised parameters
#undef TC
  // the fp32 convolution keeps a zero-padded board of all input channels in shared memory
  const size_t s3 = (size_t)(8 + a.block_kernel - 1) * (8 + a.block_kernel - 1) * a.filters * 4;
  const size_t s1 = (size_t)(8 + a.stem_kernel - 1) * (8 + a.stem_kernel - 1) * std::max(a.input_planes, a.filters) * 4;
  const int need = (int)std::max(s3, s1);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || need > (int)prop.sharedMemPerBlockOptin) {
    t->err = "board tile does not fit shared memory";
    return bail(CAZ_ERR_INVALID);
  }
  cudaFuncSetAttribute(caz::conv_fp32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
  cudaFuncSetAttribute(caz::conv_fp32_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
  cudaFuncSetAttribute(caz::conv_fp32_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
  cudaFuncSetAttribute(caz::conv_fp32_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
  t->num_sms = prop.multiProcessorCount;
  if (getenv("CAZ_TRAIN_NO_GRAPH")) t->use_graph = false;
  if (getenv("CAZ_TRAIN_PROFILE")) { t->use_graph = false; t->profiling = true; }
  if (cudaEventCreateWithFlags(&t->ev_mid, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&t->ev_end, cudaEventDisableTiming) != cudaSuccess) {
    t->err = "cudaEventCreate failed";
    return bail(CAZ_ERR_CUDA);
  }
  t->grad_split = t->convs[t->convs.size() / 2].w;   // gradients of convs[n / 2 ...] and the heads: complete at ev_mid
  if (t->convs.size() < 3) t->grad_split = 0;
  if (precision == CAZ_PRECISION_BF16) {
    if (prop.major != 10) { t->err = "the tensor-core training build needs an sm_100 GPU"; return bail(CAZ_ERR_NO_DEVICE); }
    t->tc = true;
    t->nb = (max_batch + 63) / 64 * 64;
    const int boards = (max_batch + 3) & ~3;
    const size_t Rp = (size_t)boards * 64;
#define TC2(expr) do { if ((rc = (expr))) return bail(rc); } while (0)
    for (size_t l = 0; l < t->convs.size(); ++l) {
      TrainConv& L = t->convs[l];
      const int taps = L.k * L.k;
      L.cin_pad = l == 0 ? 32 : L.cin;
      TC2(tr_alloc(t, &L.a16, Rp * L.cout));
      TC2(tr_alloc(t, &L.wf, (size_t)L.cout * taps * L.cin_pad));
      TC2(tr_halo_tmap(t, &L.tm_a16, L.a16, L.cout, a.block_kernel, boards));
      TC2(tr_kmajor_tmap(t, &L.tm_wf, L.wf, (size_t)taps * L.cin_pad, L.cout, L.cout / 2, l == 0 ? 64 : 128));
      TC2(tr_kmajor_tmap(t, &L.tm_wf_h, L.wf, (size_t)taps * L.cin_pad, L.cout, L.cout / 4, l == 0 ? 64 : 128));
      if (l > 0) {
        TC2(tr_alloc(t, &L.wb, (size_t)L.cin * taps * L.cout));
        TC2(tr_kmajor_tmap(t, &L.tm_wb, L.wb, (size_t)taps * L.cout, L.cin, L.cin / 2));
        TC2(tr_kmajor_tmap(t, &L.tm_wb_h, L.wb, (size_t)taps * L.cout, L.cin, L.cin / 4));
      }
    }
    TC2(tr_alloc(t, &t->x0_16, Rp * 32));
    TC2(tr_halo_tmap(t, &t->tm_x0, t->x0_16, 32, a.stem_kernel, boards, 64));
    TC2(tr_alloc(t, &t->dz16, Rp * F));
    TC2(tr_halo_tmap(t, &t->tm_dz, t->dz16, (int)F, a.block_kernel, boards));
    const int pitch3 = 8 + a.block_kernel - 1, pitch_s = 8 + a.stem_kernel - 1;
    const size_t k3 = (size_t)pitch3 * pitch3 * t->nb, ks = (size_t)pitch_s * pitch_s * t->nb;
    TC2(tr_alloc(t, &t->dzT, F * k3));
    TC2(tr_alloc(t, &t->aT, F * k3));
    if (!getenv("CAZ_TRAIN_NO_CONV_STATS")) TC2(tr_alloc(t, &t->conv_stats, (size_t)2 * F * ((boards + 1) / 2) * 4));
    if (getenv("CAZ_TRAIN_NO_TILE")) t->no_tile = true;
    if (getenv("CAZ_TRAIN_WHOLE_TILES")) t->whole_tiles = true;
    if (getenv("CAZ_TRAIN_NO_MIXED")) t->no_mixed = true;
    if (getenv("CAZ_TRAIN_WG_PADDED")) t->wg_padded = true;
    if (getenv("CAZ_TRAIN_MASK_FP32")) t->mask_fp32 = true;
    for (size_t l = 0; l + 1 < t->convs.size() && !t->no_tile; ++l) {   // inputs of the tower's weight-gradient GEMMs
      TrainConv& L = t->convs[l];
      if ((L.cout & 63) != 0 || (size_t)L.cout != F) continue;
      TC2(tr_alloc(t, &L.aT, F * k3));
      TC2(tr_kmajor_tmap(t, &L.tm_aT, L.aT, k3, (int)F, (int)F / 2));
    }
    TC2(tr_alloc(t, &t->dzT_stem, F * ks));
    TC2(tr_alloc(t, &t->x0T, (size_t)32 * ks));
    TC2(tr_halo_tmap(t, &t->tm_dzT, t->dzT, (int)k3, 1, (int)F / 64));
    TC2(tr_halo_tmap(t, &t->tm_dzT_stem, t->dzT_stem, (int)ks, 1, (int)F / 64));
    TC2(tr_kmajor_tmap(t, &t->tm_aT, t->aT, k3, (int)F, (int)F / 2));
    TC2(tr_kmajor_tmap(t, &t->tm_x0T, t->x0T, ks, 32, 16));
    const size_t taps_max = (size_t)std::max(a.block_kernel * a.block_kernel, a.stem_kernel * a.stem_kernel);
    TC2(tr_alloc(t, &t->wg_partial, (taps_max * 16 + 16) * F * std::max(F, (size_t)32)));
#undef TC2
    if (cudaFuncSetAttribute(caz::tc2::conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, caz::tc2::SMEM_BYTES) != cudaSuccess) {
      t->err = "cudaFuncSetAttribute(conv_tc2_kernel) failed";
      return bail(CAZ_ERR_CUDA);
    }
    if ((rc = tr_refresh_weights16(t))) return bail(rc);
  }
  // the l2 term of the first step's loss; afterwards adam_kernel leaves the sum of squares of the weights it has just written
  caz::train::l2_sum_kernel<<<t->num_sms * 8, 256, 0, t->stream>>>(t->params, t->seg, t->n_params, t->loss_acc + 4);
  if (cudaStreamSynchronize(t->stream) != cudaSuccess) { t->err = "initialisation failed"; return bail(CAZ_ERR_CUDA); }
  *out = t;
  return CAZ_OK;
}

int caz_train_forward_backward_begin(caz_trainer* t, const float* planes, const float* policy_target, const float* value_target,
                                     int32_t batch) {
  if (!t) return CAZ_ERR_INVALID;
  if (!planes || !policy_target || !value_target || batch < 1 || batch > t->max_batch)
    return tfail(t, CAZ_ERR_INVALID, fmt("caz_train_forward_backward: bad argument (batch %d, max_batch %d)", batch, t->max_batch));
  TR_TRY(t, cudaSetDevice(t->device));
  const caz_arch& a = t->arch;
  TR_TRY(t, cudaMemcpyAsync(t->planes, planes, (size_t)batch * a.input_planes * 64 * 4, cudaMemcpyHostToDevice, t->stream));
  TR_TRY(t, cudaMemcpyAsync(t->policy_t, policy_target, (size_t)batch * a.n_labels * 4, cudaMemcpyHostToDevice, t->stream));
  TR_TRY(t, cudaMemcpyAsync(t->value_t, value_target, (size_t)batch * 4, cudaMemcpyHostToDevice, t->stream));
  int rc;
  if (t->profiling) {
    for (auto& pr : t->prof) cudaEventDestroy(pr.second);
    t->prof.clear();
    tr_launch_ok(t, "start");
  }
  cudaGraphExec_t exec = nullptr;
  for (auto& g : t->graphs)
    if (g.first == batch) exec = g.second;
  if (t->use_graph && !exec) {
    // first step at this batch size: record the launch sequence (pointers and sizes are fixed per batch size)
    cudaGraph_t graph = nullptr;
    TR_TRY(t, cudaStreamBeginCapture(t->stream, cudaStreamCaptureModeThreadLocal));
    t->capturing = true;
    rc = tr_forward_backward(t, batch);
    t->capturing = false;
    const cudaError_t ce = cudaStreamEndCapture(t->stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
      cudaGetLastError();
      exec = nullptr;
      t->use_graph = false;   // fall back to plain launches
    } else {
      t->graphs.emplace_back(batch, exec);
    }
    if (graph) cudaGraphDestroy(graph);
  }
  if (exec) TR_TRY(t, cudaGraphLaunch(exec, t->stream));
  else if ((rc = tr_forward_backward(t, batch))) return rc;
  TR_TRY(t, cudaEventRecord(t->ev_end, t->stream));
  t->pending_batch = batch;
  return CAZ_OK;
}

int caz_train_wait_gradients(caz_trainer* t, int32_t part, void* stream) {
  if (!t || part < 0 || part > 1) return CAZ_ERR_INVALID;
  TR_TRY(t, cudaSetDevice(t->device));
  TR_TRY(t, cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), part == 0 ? t->ev_mid : t->ev_end, 0));
  return CAZ_OK;
}

int64_t caz_train_gradient_split(const caz_trainer* t) { return t ? (int64_t)t->grad_split : 0; }

int caz_train_forward_backward_end(caz_trainer* t, float* losses) {
  if (!t) return CAZ_ERR_INVALID;
  if (t->pending_batch < 1) return tfail(t, CAZ_ERR_INVALID, "caz_train_forward_backward_end without a begin");
  const int batch = t->pending_batch;
  t->pending_batch = 0;
  TR_TRY(t, cudaSetDevice(t->device));
  double acc[5];
  TR_TRY(t, cudaMemcpyAsync(acc, t->loss_acc, sizeof acc, cudaMemcpyDeviceToHost, t->stream));
  TR_TRY(t, cudaStreamSynchronize(t->stream));
  if (t->profiling && t->step == 3) tr_profile_report(t);
  const float ce = (float)(acc[0] / batch), mse = (float)(acc[1] / batch), reg = (float)(t->cfg.l2 * acc[4]);
  t->last_losses[0] = t->cfg.policy_loss_weight * ce + t->cfg.value_loss_weight * mse + reg;
  t->last_losses[1] = ce;
  t->last_losses[2] = mse;
  t->last_losses[3] = reg;
  if (losses) memcpy(losses, t->last_losses, sizeof t->last_losses);
  return CAZ_OK;
}

int caz_train_forward_backward(caz_trainer* t, const float* planes, const float* policy_target, const float* value_target, int32_t batch,
                               float* losses) {
  int rc;
  if ((rc = caz_train_forward_backward_begin(t, planes, policy_target, value_target, batch))) return rc;
  return caz_train_forward_backward_end(t, losses);
}

int caz_train_apply(caz_trainer* t, float grad_scale) {
  if (!t) return CAZ_ERR_INVALID;
  TR_TRY(t, cudaSetDevice(t->device));
  ++t->step;
  const double b1 = t->cfg.beta_1, b2 = t->cfg.beta_2;
  const float lr_t = (float)(t->cfg.learning_rate * std::sqrt(1.0 - std::pow(b2, (double)t->step)) / (1.0 - std::pow(b1, (double)t->step)));
  TR_TRY(t, cudaMemsetAsync(t->loss_acc + 4, 0, sizeof(double), t->stream));
  caz::train::adam_kernel<<<t->num_sms * 8, 256, 0, t->stream>>>(t->params, t->grads, t->adam_m, t->adam_v, t->seg, t->n_params, lr_t, t->cfg.beta_1,
                                                              t->cfg.beta_2, t->cfg.epsilon, t->cfg.l2, grad_scale, t->cfg.bn_momentum,
                                                              t->loss_acc + 4);
  int rc;
  if ((rc = tr_launch_ok(t, "adam_kernel"))) return rc;
  if ((rc = tr_refresh_weights16(t))) return rc;
  TR_TRY(t, cudaStreamSynchronize(t->stream));
  return CAZ_OK;
}

int caz_train_read(caz_trainer* t, int32_t what, float* host, int64_t count) {
  if (!t || !host) return CAZ_ERR_INVALID;
  if (count != (int64_t)t->n_params) return tfail(t, CAZ_ERR_INVALID, fmt("caz_train_read: expected %lld values", (long long)t->n_params));
  const float* src = what == 0 ? t->params : (what == 1 ? t->grads : (what == 2 ? t->adam_m : t->adam_v));
  TR_TRY(t, cudaSetDevice(t->device));
  TR_TRY(t, cudaStreamSynchronize(t->stream));
  TR_TRY(t, cudaMemcpy(host, src, (size_t)count * 4, cudaMemcpyDeviceToHost));
  return CAZ_OK;
}

int caz_train_read_outputs(caz_trainer* t, int32_t batch, float* policy, float* value) {
  if (!t || batch < 1 || batch > t->max_batch) return CAZ_ERR_INVALID;
  TR_TRY(t, cudaSetDevice(t->device));
  TR_TRY(t, cudaStreamSynchronize(t->stream));
  if (policy) TR_TRY(t, cudaMemcpy(policy, t->probs, (size_t)batch * t->arch.n_labels * 4, cudaMemcpyDeviceToHost));
  if (value) TR_TRY(t, cudaMemcpy(value, t->vout, (size_t)batch * 4, cudaMemcpyDeviceToHost));
  return CAZ_OK;
}

float* caz_train_grad_device_ptr(caz_trainer* t) { return t ? t->grads : nullptr; }
int64_t caz_train_num_params(const caz_trainer* t) { return t ? (int64_t)t->n_params : 0; }
int64_t caz_train_steps(const caz_trainer* t) { return t ? t->step : 0; }

}  // extern "C"
