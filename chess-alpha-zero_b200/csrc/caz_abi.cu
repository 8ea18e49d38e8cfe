// C ABI (include/caz_b200.h) and host-side orchestration of the sm_100a evaluator.
// Host work done here: BatchNorm folding (fp32), weight re-layout for the kernels, TMA tensor-map
// construction, buffer ownership, the per-layer launch sequence of the forward pass.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/caz_b200.h"
#include "conv_sb.cuh"
#include "conv_tc.cuh"
#include "conv_tc2.cuh"
#include "conv_tower.cuh"
#include "kernels_simt.cuh"

namespace {

std::mutex g_err_mutex;
std::string g_create_error;

std::string fmt(const char* f, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof buf, f, ap);
  va_end(ap);
  return buf;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct ConvLayer {
  int k = 0, cin = 0, cin_pad = 0, cout = 0;
  void* w = nullptr;      // bf16 [cout][k*k*cin_pad] (tensor path) or fp32 [k*k][cin][cout] (fp32 path)
  float* bias = nullptr;  // [cout]
  CUtensorMap tmap_w;    // box 64 x cout       (one-CTA kernel)
  CUtensorMap tmap_w2;   // box 64 x cout / 2   (CTA-pair kernel: each CTA stages half of the output channels)
  CUtensorMap tmap_w2h;  // box 64 x cout / 4   (the same with two N tiles of cout / 2 channels)
};

}  // namespace

struct caz_handle {
  caz_arch arch{};
  int device = 0;
  int precision = 0;
  int max_batch = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  std::vector<ConvLayer> convs;  // stem, then conv1/conv2 of each block
  bool snake = true;             // odd layers walk the batch backwards (CAZ_SNAKE=0 turns it off)
  bool l2_prefetch = false;       // CAZ_L2_PREFETCH=0 turns the one-tile-ahead L2 prefetches off
  bool split_n = true;           // two 128-channel N tiles per board group when the grid would not fill the device; CAZ_NO_SPLIT_N=1
  bool pdl = true;               // programmatic dependent launch between the per-layer convolutions; CAZ_NO_PDL=1 turns it off
  bool use_2cta = true;          // per-layer convolutions as CTA pairs (cta_group::2); CAZ_CONV_2CTA=0 turns it off
  // every environment switch is read ONCE, in caz_create (no getenv on the call path)
  bool sb_no_pairs = false;      // CAZ_SB_NO_PAIRS
  bool no_direct_output = false; // CAZ_NO_DIRECT_OUTPUT
  int sb_dbg = 0;                // CAZ_SB_DBG
  bool tw_all_lanes_fence = false;   // CAZ_TOWER_ALL_LANES_FENCE
  bool host_timing = false;          // CAZ_HOST_TIMING
  double ht_sum[3] = {0, 0, 0};
  long ht_calls = 0;
  int dbg_epi = 0;               // CAZ_DBG_EPI (only with -DCAZ_TIMING_EXPERIMENTS)
  // weight slabs the layers point into: the tower's layers are contiguous so that one tensor map covers them
  void* w_stem = nullptr;
  void* sb_w_stem = nullptr;   // stem weights padded to 64 input channels (small-batch kernel)
  void* w_tower = nullptr;   // [2*res_blocks][filters][9*filters] (tensor-core build)
  float* bias_all = nullptr; // [1 + 2*res_blocks][filters]
  // whole tower in one persistent launch (conv_tower.cuh): every batch the small-batch kernel does not take
  bool tw_ok = false;        // architecture fits it
  bool tw_enabled = true;    // CAZ_NO_TOWER=1 keeps the per-layer launch chain (conv_tc2.cuh)
  int tw_mode = 0;           // caz_set_fused_tower / CAZ_TOWER_MODE: 0 = by batch size, 1 / 2 = forced groups per pair
  CUtensorMap tw_tm_w_tower; // all tower layers' weights as one matrix, box 64 x filters / 2
  // the tower's intermediate activations: ONE slab [fp32 stream | 16-bit x | 16-bit y] for <= 8 boards per CTA pair, behind an
  // L2 persistence window (CAZ_NO_L2_PERSIST=1: no window)
  void* tw_scratch = nullptr;
  size_t tw_scratch_bytes = 0;
  int tw_scratch_boards = 0;
  CUtensorMap tw_tm_x, tw_tm_y;
  long long* tw_trace = nullptr;   // [64][8] clock stamps (caz_debug_tower_trace)
  bool tw_trace_on = false;
  // small-batch persistent tower kernel (conv_sb.cuh)
  bool sb_ok = false;        // architecture fits it
  int sb_limit = -1;         // batches <= this use it; -1 = automatic
  CUtensorMap sb_tm_xb[2], sb_tm_y[2];  // halo views for 1 and 2 boards per CTA tile
  CUtensorMap sb_tm_w_stem[4], sb_tm_w_tower[4];      // weight maps for N_s = 16, 32, 64, 128
  int sb_force_nb = 0;
  unsigned char* sb_sm_near = nullptr;   // [smid] -> which flag-line class (copy) is the near one, 0xFF unknown
  unsigned* sb_counter = nullptr;   // progress flags of the small-batch kernel: 32 candidate grains of 4 KB
  uint16_t sb_flag_line[2][caz::sb::MAX_TILES];   // [copy][tile] -> candidate line of that tile's flags (copy 0 / 1: the two latency classes)
  bool sb_two_dies = false;
  long long* sb_trace = nullptr;   // [64][8] clock stamps of CTA 0 (debug)
  bool sb_trace_on = false;
  unsigned sb_epoch = 0;
  float* sb_hpart = nullptr;    // partial head-conv sums of the small-batch kernel
  float* sb_hstats = nullptr;   // its softmax / value statistics
  float *head_w = nullptr, *head_b = nullptr;
  float *pfc_wb = nullptr, *vfc1_wb = nullptr;   // blocked copies for the small-batch kernel's bulk staging (conv_sb.cuh)
  float *pfc_w = nullptr, *pfc_b = nullptr, *vfc1_w = nullptr, *vfc1_b = nullptr, *vfc2_w = nullptr,
        *vfc2_b = nullptr;
  void* in_act = nullptr;  // [cap][64][in_cpad]
  int in_cpad = 0;
  // [cap][64][filters] each.  fp32 build: three fp32 buffers in rotation.  Tensor-core build:
  // act[0] = bf16 copy of the residual stream (A operand of conv1), act[1] = bf16 conv1 output
  // (A operand of conv2), act[2] = the residual stream itself, kept in fp32.
  void* act[3] = {nullptr, nullptr, nullptr};
  CUtensorMap tmap_in, tmap_act[2];
  float *featp = nullptr, *featv = nullptr, *vhid = nullptr;
  // heads' Dense layers on the tensor cores (3-term 16-bit split, see split3_kernel)
  bool dense_tc = false;
  int mpad = 0;                       // rows of the split operands: max_batch rounded up to 128
  uint16_t *a3p = nullptr, *a3v = nullptr, *w3p = nullptr, *w3v = nullptr;
  float *bias3p = nullptr;            // policy bias padded to a multiple of 256
  int np_pad = 0;
  CUtensorMap tm_a3p, tm_a3v, tm_w3p, tm_w3v;
  // staging for the host-pointer entry points
  float *d_planes = nullptr, *d_policy = nullptr, *d_value = nullptr;
  uint8_t *d_mask = nullptr, *d_records = nullptr;
  // second staging slot + copy streams: caz_eval pipelines H2D / forward / D2H over chunks of a large batch
  float *d_planes2 = nullptr, *d_policy2 = nullptr, *d_value2 = nullptr;
  uint8_t *d_mask2 = nullptr, *d_records2 = nullptr;
  // caz_submit_records_priors: CSR legal-move labels in, one float64 prior per legal move out (two staging slots)
  int32_t* d_off[2] = {nullptr, nullptr};      // [cap + 1]
  int32_t* d_lab[2] = {nullptr, nullptr};      // [cap * CAZ_MAX_MOVES]
  double* d_pri[2] = {nullptr, nullptr};       // [cap * CAZ_MAX_MOVES]
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  cudaEvent_t ev_in[2] = {}, ev_fwd[2] = {}, ev_out[2] = {};
  // caz_submit / caz_collect: tickets count up; ticket t completes on ev_done[t % TICKET_RING], its staging slot is t & 1
  // (a third submission simply queues behind the first one's slot on the streams).  caz_collect may run on another
  // thread than caz_submit: it touches only its ring entry.
  static constexpr int TICKET_RING = 16;
  cudaEvent_t ev_done[TICKET_RING] = {};
  std::atomic<int> ticket_state[TICKET_RING];   // 0 free, 1 in flight
  int32_t next_ticket = 0;
  std::mutex err_mutex;          // last_error is written under it (collect and submit may fail concurrently)
  int* err_flag_host = nullptr;  // mapped pinned: readable even after a device trap
  int* err_flag_dev = nullptr;
  // PUCT scratch
  void* puct_buf = nullptr;
  size_t puct_cap = 0;
  int64_t launches = 0;
  int64_t weight_bytes = 0;
  int64_t inexact_weights = 0;   // bf16 build: bf16-valued weights the fp16 encoding could not hold exactly (see weight_bits)
  // profiling
  // one event set per forward pass while profiling; drained (without blocking the forward passes) on read
  struct EvSet { cudaEvent_t ev[CAZ_STAGES + 1]; int64_t launches[CAZ_STAGES]; };
  bool profiling = false;
  std::vector<EvSet> ev_pool;
  size_t ev_used = 0;
  cudaEvent_t* ev = nullptr;  // set being recorded by the current forward
  float stage_ms[CAZ_STAGES] = {};
  int64_t stage_launches[CAZ_STAGES] = {};
  int64_t* pending_launches = nullptr;
};

namespace {

int fail(caz_handle* h, int code, const std::string& msg) {
  if (h) {
    std::lock_guard<std::mutex> g(h->err_mutex);
    h->err = msg;
  } else {
    std::lock_guard<std::mutex> g(g_err_mutex);
    g_create_error = msg;
  }
  return code;
}

#define CU_TRY(h, expr)                                                                              \
  do {                                                                                               \
    cudaError_t e__ = (expr);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      int code__ = CAZ_ERR_CUDA;                                                                     \
      std::string m__ = fmt("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      if ((h) && (h)->err_flag_host && *(h)->err_flag_host) {                                        \
        code__ = CAZ_ERR_KERNEL;                                                                     \
        m__ += fmt(" [kernel watchdog code %d]", *(h)->err_flag_host);                               \
      }                                                                                              \
      return fail((h), code__, m__);                                                                 \
    }                                                                                                \
  } while (0)

int expected_tensor_count(const caz_arch& a) { return 5 + 10 * a.res_blocks + 7 + 9; }

bool tensor_path(const caz_handle* h) { return h->precision != CAZ_PRECISION_FP32; }
// 16-bit formats of the two tensor-core operands (include/caz_b200.h).  tcgen05.mma kind::f16 has separate A and B format
// fields, but a launch with A = fp16 and B = bf16 dies with "an illegal instruction was encountered" on sm_100a (measured
// in round 2): both operands must have the same format.  The bf16 build therefore hands its bf16-VALUED weights to the
// tensor core in the fp16 encoding -- every bf16 value with 2^-17 <= |w| < 65504 is exactly representable in fp16 (8
// significant bits, fp16 keeps 11 down to 2^-14 and 8 down to 2^-17) -- next to fp16 activation operands.
bool act_fp16(const caz_handle* h) { return h->precision == CAZ_PRECISION_FP16 || h->precision == CAZ_PRECISION_BF16; }
bool w_fp16(const caz_handle* h) { return act_fp16(h); }   // the encoding the tensor core reads
bool w_bf16_grid(const caz_handle* h) { return h->precision == CAZ_PRECISION_BF16 || h->precision == CAZ_PRECISION_BF16_PURE; }
int operand_fmt(const caz_handle* h) { return (act_fp16(h) ? caz::ptx::FMT_A_FP16 : 0) | (w_fp16(h) ? caz::ptx::FMT_B_FP16 : 0); }

// One folded convolution weight -> the 16-bit word the tensor core reads.  *inexact counts bf16 values that the fp16
// encoding cannot hold exactly (|w| < 2^-17 and not 0, or >= 65504; saturated / rounded to nearest then).
uint16_t weight_bits(const caz_handle* h, float v, int64_t* inexact) {
  uint16_t bits;
  if (w_bf16_grid(h)) {
    const __nv_bfloat16 bv = __float2bfloat16_rn(v);
    if (!w_fp16(h)) { memcpy(&bits, &bv, 2); return bits; }
    const float g = __bfloat162float(bv);
    const __half hv = __float2half_rn(fminf(fmaxf(g, -65504.0f), 65504.0f));
    if (__half2float(hv) != g && inexact) ++*inexact;
    memcpy(&bits, &hv, 2);
    return bits;
  }
  const __half hv = __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f));
  memcpy(&bits, &hv, 2);
  return bits;
}

// BatchNorm inference folded into the preceding bias-free convolution (Keras semantics,
// model_chess.py:65-69): y = conv(x)*inv + (beta - mean*inv), inv = gamma / sqrt(var + eps).
void bn_fold(const caz_tensor* bn, int c, float eps, std::vector<float>& inv, std::vector<float>& shift) {
  inv.resize(c);
  shift.resize(c);
  for (int i = 0; i < c; ++i) {
    const double s = static_cast<double>(bn[0].data[i]) / std::sqrt(static_cast<double>(bn[3].data[i]) + eps);
    inv[i] = static_cast<float>(s);
    shift[i] = static_cast<float>(static_cast<double>(bn[1].data[i]) - static_cast<double>(bn[2].data[i]) * s);
  }
}

int upload(caz_handle* h, void* dst, const void* src, size_t bytes) {
  CU_TRY(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  return CAZ_OK;
}

int make_tmap(caz_handle* h, CUtensorMap* m, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
              const cuuint32_t* box, int row_bytes = 128) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(h, CAZ_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, act_fp16(h) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, dims, strides, box, ones,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, CAZ_ERR_CUDA, fmt("cuTensorMapEncodeTiled failed with CUresult %d", (int)r));
  return CAZ_OK;
}

// Halo view of NHWC bf16 activations [cap][8][8][c]: dims (c, file, position, rank), box 64 x hw x nb x hw, fetched at
// (c0, -pad, n0, -pad): the out-of-bounds fill supplies the zero border on all four sides (hw = 8 + k - 1).
int make_halo_tmap(caz_handle* h, CUtensorMap* m, void* base, int c, int k, int nb, int cap = 0, int row_bytes = 128) {
  const cuuint32_t hw = 8 + k - 1;
  cuuint64_t dims[4] = {(cuuint64_t)c, 8, (cuuint64_t)(cap ? cap : h->max_batch), 8};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)c * 128, (cuuint64_t)c * 16};
  cuuint32_t box[4] = {(cuuint32_t)row_bytes / 2, hw, (cuuint32_t)nb, hw};
  return make_tmap(h, m, base, 4, dims, strides, box, row_bytes);
}

int setup_dense_tc(caz_handle* h, const float* pfc_w, const float* pfc_b, const float* vfc1_w);

// Installs (folds, re-lays-out, uploads) one full weight set. Buffers are allocated on first use.
int install_weights(caz_handle* h, const caz_tensor* t, int n_tensors) {
  const caz_arch& a = h->arch;
  if (n_tensors != expected_tensor_count(a))
    return fail(h, CAZ_ERR_INVALID, fmt("expected %d tensors for %d residual blocks, got %d",
                                        expected_tensor_count(a), a.res_blocks, n_tensors));
  const bool tcp = tensor_path(h);
  const int F = a.filters;
  int ti = 0;
  int64_t wbytes = 0;
  std::vector<float> inv, shift;
  const int n_convs = 1 + 2 * a.res_blocks;
  const bool first_install = h->convs.empty();
  h->inexact_weights = 0;
  if (first_install) {
    h->convs.resize(n_convs);
    const int scin_pad = h->in_cpad;
    const size_t stem_elems = (size_t)a.stem_kernel * a.stem_kernel * scin_pad * F;
    const size_t blk_elems = (size_t)a.block_kernel * a.block_kernel * F * F;
    const size_t esz = tcp ? 2 : 4;
    CU_TRY(h, cudaMalloc(&h->w_stem, stem_elems * esz));
    if (a.res_blocks > 0) CU_TRY(h, cudaMalloc(&h->w_tower, blk_elems * esz * 2 * a.res_blocks));
    CU_TRY(h, cudaMalloc(&h->bias_all, (size_t)n_convs * F * sizeof(float)));
    for (int li = 0; li < n_convs; ++li) {
      h->convs[li].w = li == 0 ? h->w_stem : static_cast<char*>(h->w_tower) + (size_t)(li - 1) * blk_elems * esz;
      h->convs[li].bias = h->bias_all + (size_t)li * F;
    }
  }
  for (int li = 0; li < n_convs; ++li) {
    ConvLayer& L = h->convs[li];
    const int k = li == 0 ? a.stem_kernel : a.block_kernel;
    const int cin = li == 0 ? a.input_planes : F;
    const int cin_pad = li == 0 ? h->in_cpad : (tcp ? ((cin + 63) / 64) * 64 : cin);
    const caz_tensor& kt = t[ti];
    if (kt.numel != (int64_t)k * k * cin * F)
      return fail(h, CAZ_ERR_INVALID, fmt("conv %d kernel has %lld elements, expected %lld", li,
                                          (long long)kt.numel, (long long)k * k * cin * F));
    for (int j = 1; j <= 4; ++j)
      if (t[ti + j].numel != F) return fail(h, CAZ_ERR_INVALID, fmt("conv %d BatchNorm tensor %d: expected %d values", li, j, F));
    bn_fold(t + ti + 1, F, a.bn_epsilon, inv, shift);
    ti += 5;
    const int taps = k * k;
    const size_t ktot = (size_t)taps * cin_pad;
    if (first_install) {
      L.k = k; L.cin = cin; L.cin_pad = cin_pad; L.cout = F;
      if (tcp) {
        cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)F};
        cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
        const int rbytes = li == 0 ? h->in_cpad * 2 : 128;   // operand row = swizzle span
        cuuint32_t box[2] = {(cuuint32_t)rbytes / 2, (cuuint32_t)F};
        int rc = make_tmap(h, &L.tmap_w, L.w, 2, dims, strides, box, rbytes);
        if (rc) return rc;
        cuuint32_t box2[2] = {(cuuint32_t)rbytes / 2, (cuuint32_t)F / 2};
        if ((rc = make_tmap(h, &L.tmap_w2, L.w, 2, dims, strides, box2, rbytes))) return rc;
        cuuint32_t box2h[2] = {(cuuint32_t)rbytes / 2, (cuuint32_t)(F >= 4 ? F / 4 : 1)};
        if ((rc = make_tmap(h, &L.tmap_w2h, L.w, 2, dims, strides, box2h, rbytes))) return rc;
      }
    }
    if (tcp) {
      // W[oc][tap*cin_pad + ic] bf16, K-major rows: the B operand of the implicit GEMM
      // 16-bit words: bf16, or fp16 in the FP16 build (both have a +0.0 bit pattern of 0)
      std::vector<uint16_t> wp(ktot * F, 0);
      for (int tap = 0; tap < taps; ++tap)
        for (int ic = 0; ic < cin; ++ic) {
          const float* src = kt.data + ((size_t)tap * cin + ic) * F;
          for (int oc = 0; oc < F; ++oc)
            wp[(size_t)oc * ktot + (size_t)tap * cin_pad + ic] = weight_bits(h, src[oc] * inv[oc], &h->inexact_weights);
        }
      int rc = upload(h, L.w, wp.data(), wp.size() * 2);
      if (rc) return rc;
      wbytes += (int64_t)taps * cin * F * 2;
      if (li == 0) {
        // the small-batch kernel keeps 128-byte operand rows everywhere: its own copy of the stem, padded to 64 channels
        const size_t k64 = (size_t)taps * 64;
        std::vector<uint16_t> w64(k64 * F, 0);
        for (int oc = 0; oc < F; ++oc)
          for (int tap = 0; tap < taps; ++tap)
            for (int ic = 0; ic < cin && ic < 64; ++ic) w64[(size_t)oc * k64 + (size_t)tap * 64 + ic] = wp[(size_t)oc * ktot + (size_t)tap * cin_pad + ic];
        if (!h->sb_w_stem) CU_TRY(h, cudaMalloc(&h->sb_w_stem, w64.size() * 2));
        if ((rc = upload(h, h->sb_w_stem, w64.data(), w64.size() * 2))) return rc;
      }
    } else {
      std::vector<float> wp((size_t)taps * cin * F);
      for (size_t i = 0; i < wp.size(); ++i) wp[i] = kt.data[i] * inv[i % F];
      int rc = upload(h, L.w, wp.data(), wp.size() * 4);
      if (rc) return rc;
      wbytes += (int64_t)wp.size() * 4;
    }
    int rc = upload(h, L.bias, shift.data(), F * sizeof(float));
    if (rc) return rc;
    wbytes += F * 4;
  }
  // ---- heads: 1x1 convs (policy filters first, then value filters), fp32 in both builds ----
  const int pf = a.policy_filters, vf = a.value_filters, ft = pf + vf;
  std::vector<float> hw((size_t)ft * F), hb(ft);
  const caz_tensor* pol = t + ti;       // conv, g, b, m, v, dense k, dense b
  const caz_tensor* val = t + ti + 7;   // conv, g, b, m, v, d1 k, d1 b, d2 k, d2 b
  if (pol[0].numel != (int64_t)F * pf || val[0].numel != (int64_t)F * vf)
    return fail(h, CAZ_ERR_INVALID, "head conv kernel size mismatch");
  for (int j = 1; j <= 4; ++j)
    if (pol[j].numel != pf || val[j].numel != vf) return fail(h, CAZ_ERR_INVALID, "head BatchNorm size mismatch");
  bn_fold(pol + 1, pf, a.bn_epsilon, inv, shift);
  for (int f = 0; f < pf; ++f) {
    for (int c = 0; c < F; ++c) hw[(size_t)f * F + c] = pol[0].data[(size_t)c * pf + f] * inv[f];
    hb[f] = shift[f];
  }
  bn_fold(val + 1, vf, a.bn_epsilon, inv, shift);
  for (int f = 0; f < vf; ++f) {
    for (int c = 0; c < F; ++c) hw[(size_t)(pf + f) * F + c] = val[0].data[(size_t)c * vf + f] * inv[f];
    hb[pf + f] = shift[f];
  }
  const int kp = pf * 64, kv = vf * 64;
  if (pol[5].numel != (int64_t)kp * a.n_labels || pol[6].numel != a.n_labels)
    return fail(h, CAZ_ERR_INVALID, "policy Dense size mismatch");
  if (val[5].numel != (int64_t)kv * a.value_fc || val[6].numel != a.value_fc || val[7].numel != a.value_fc ||
      val[8].numel != 1)
    return fail(h, CAZ_ERR_INVALID, "value Dense size mismatch");
  if (!h->head_w) {
    CU_TRY(h, cudaMalloc(&h->head_w, hw.size() * 4));
    CU_TRY(h, cudaMalloc(&h->head_b, hb.size() * 4));
    CU_TRY(h, cudaMalloc(&h->pfc_w, (size_t)pol[5].numel * 4));
    CU_TRY(h, cudaMalloc(&h->pfc_b, (size_t)pol[6].numel * 4));
    CU_TRY(h, cudaMalloc(&h->vfc1_w, (size_t)val[5].numel * 4));
    CU_TRY(h, cudaMalloc(&h->vfc1_b, (size_t)val[6].numel * 4));
    CU_TRY(h, cudaMalloc(&h->vfc2_w, (size_t)val[7].numel * 4));
    CU_TRY(h, cudaMalloc(&h->vfc2_b, 4));
  }
  int rc;
  {
    // blocked copies: policy Dense kernel as [128-label unit][k][128], value Dense kernel as [16-unit group][k][16], zero-padded
    const int units = (a.n_labels + caz::sb::LABEL_UNIT - 1) / caz::sb::LABEL_UNIT;
    const int groups = (a.value_fc + caz::sb::HID_GROUP - 1) / caz::sb::HID_GROUP;
    std::vector<float> pb((size_t)units * kp * caz::sb::LABEL_UNIT, 0.0f), vb((size_t)groups * kv * caz::sb::HID_GROUP, 0.0f);
    for (int k = 0; k < kp; ++k)
      for (int n = 0; n < a.n_labels; ++n)
        pb[((size_t)(n / caz::sb::LABEL_UNIT) * kp + k) * caz::sb::LABEL_UNIT + n % caz::sb::LABEL_UNIT] = pol[5].data[(size_t)k * a.n_labels + n];
    for (int k = 0; k < kv; ++k)
      for (int u = 0; u < a.value_fc; ++u)
        vb[((size_t)(u / caz::sb::HID_GROUP) * kv + k) * caz::sb::HID_GROUP + u % caz::sb::HID_GROUP] = val[5].data[(size_t)k * a.value_fc + u];
    if (!h->pfc_wb) {
      CU_TRY(h, cudaMalloc(&h->pfc_wb, pb.size() * 4));
      CU_TRY(h, cudaMalloc(&h->vfc1_wb, vb.size() * 4));
    }
    if ((rc = upload(h, h->pfc_wb, pb.data(), pb.size() * 4))) return rc;
    if ((rc = upload(h, h->vfc1_wb, vb.data(), vb.size() * 4))) return rc;
  }
  if ((rc = upload(h, h->head_w, hw.data(), hw.size() * 4))) return rc;
  if ((rc = upload(h, h->head_b, hb.data(), hb.size() * 4))) return rc;
  if ((rc = upload(h, h->pfc_w, pol[5].data, (size_t)pol[5].numel * 4))) return rc;
  if ((rc = upload(h, h->pfc_b, pol[6].data, (size_t)pol[6].numel * 4))) return rc;
  if ((rc = upload(h, h->vfc1_w, val[5].data, (size_t)val[5].numel * 4))) return rc;
  if ((rc = upload(h, h->vfc1_b, val[6].data, (size_t)val[6].numel * 4))) return rc;
  if ((rc = upload(h, h->vfc2_w, val[7].data, (size_t)val[7].numel * 4))) return rc;
  if ((rc = upload(h, h->vfc2_b, val[8].data, 4))) return rc;
  wbytes += (int64_t)(hw.size() + hb.size() + pol[5].numel + pol[6].numel + val[5].numel + val[6].numel + val[7].numel + 1) * 4;
  h->weight_bytes = wbytes;
  if (tcp && (rc = setup_dense_tc(h, pol[5].data, pol[6].data, val[5].data))) return rc;
  return CAZ_OK;
}

// Launch with (or without) the programmatic-stream-serialization attribute: see pdl_sync() in kernels_simt.cuh.
template <typename... KArgs, typename... Args>
cudaError_t launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int check_launch(caz_handle* h, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, CAZ_ERR_CUDA, fmt("launch of %s failed: %s", what, cudaGetErrorString(e)));
  h->launches++;
  return CAZ_OK;
}

// residual is fp32 in both builds.  fp32 build: `out` is the fp32 output.  Tensor-core build: `out` is the
// bf16 output (or null) and `out_f32` the fp32 residual-stream output (or null).
int launch_conv(caz_handle* h, int li, const void* in, const CUtensorMap* tmap_in, int in_stride, const float* residual,
                void* out, float* out_f32, int batch, bool fuse_heads = false) {
  const ConvLayer& L = h->convs[li];
  if (tensor_path(h)) {
    caz::tc::ConvArgs args{};
    args.batch = batch;
    args.num_tiles = (batch + 1) / 2;
    args.ksize = L.k;
    args.row_bytes = li == 0 ? h->in_cpad * 2 : 128;
    args.kchunks = L.cin_pad / (args.row_bytes / 2);
    args.n = L.cout;
    args.relu = 1;
    args.bias = L.bias;
    args.residual = residual;
    args.out_f32 = out_f32;
    args.out_bf16 = static_cast<__nv_bfloat16*>(out);
    args.head_w = fuse_heads ? h->head_w : nullptr;
    args.head_b = h->head_b;
    args.featp = h->featp;
    args.featv = h->featv;
    args.pf = h->arch.policy_filters;
    args.vf = h->arch.value_filters;
    args.fp16 = operand_fmt(h);
    args.n_tiles = 1;
    args.n_total = L.cout;
    args.m_rows = batch * 64;
    args.out_rm = nullptr;
    args.l2_prefetch = h->l2_prefetch ? 1 : 0;
    args.pdl = 0;
    args.reverse = (h->snake && (li & 1)) ? 1 : 0;
#ifdef CAZ_TIMING_EXPERIMENTS   // never in the shipped library: results are wrong with these (drops epilogue traffic)
    if (h->dbg_epi & 1) args.residual = nullptr;
    if (h->dbg_epi & 2) args.out_f32 = nullptr;
    if (h->dbg_epi & 4) args.out_bf16 = nullptr;
#endif
    args.err_flag = h->err_flag_dev;
    if (h->use_2cta && args.num_tiles >= 2 && L.cout % 64 == 0) {
      // CTA pairs: one cta_group::2 MMA covers two tiles (four boards)
      const int pairs = (args.num_tiles + 1) / 2;
      const int max_pairs = h->num_sms / 2;
      // Few boards (up to ~148): one M = 256 tile per CTA pair leaves half of the device idle, and a lone tile cannot
      // overlap its epilogue with anything.  Two N tiles of 128 channels per board group: twice the CTAs, half the tile.
      // (Not the last layer: the fused head convolutions want all 256 channels of a row in one thread.  Not for bigger
      // batches either: splitting to fill a partly empty last wave was measured at 600-1 200 boards and costs 10 % --
      // a half tile reads the same activations for half the work.)
      const bool split_n = h->split_n && !fuse_heads && L.cout == 256 && 2 * pairs <= max_pairs;
      if (split_n) {
        args.n_tiles = 2;
        args.n = L.cout / 2;
      }
      const int work = pairs * args.n_tiles;
      const int grid = 2 * (work < max_pairs ? work : max_pairs);
      // consecutive layers overlap (prologue and first weights of layer l+1 under the tail of layer l): see conv_tc2.cuh
      args.pdl = h->pdl ? 1 : 0;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(caz::tc2::THREADS);
      cfg.dynamicSmemBytes = caz::tc2::SMEM_BYTES;
      cfg.stream = h->stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at;
      cfg.numAttrs = h->pdl ? 1 : 0;
      cudaError_t e = cudaLaunchKernelEx(&cfg, caz::tc2::conv_tc2_kernel, *tmap_in, split_n ? L.tmap_w2h : L.tmap_w2, L.tmap_w2, args);
      if (e != cudaSuccess) return fail(h, CAZ_ERR_CUDA, fmt("launch of conv_tc2_kernel failed: %s", cudaGetErrorString(e)));
      return check_launch(h, "conv_tc2_kernel");
    }
    const int grid = args.num_tiles < h->num_sms ? args.num_tiles : h->num_sms;
    caz::tc::conv_tc_kernel<<<grid, caz::tc::THREADS, caz::tc::SMEM_BYTES, h->stream>>>(*tmap_in, L.tmap_w, args);
    return check_launch(h, "conv_tc_kernel");
  }
  const size_t smem = (size_t)(8 + L.k - 1) * (8 + L.k - 1) * L.cin * sizeof(float);
  const float* fin = static_cast<const float*>(in);
  const float* fres = residual;
  float* fout = static_cast<float*>(out);
  const float* fw = static_cast<const float*>(L.w);
  switch (L.k) {
    case 1: caz::conv_fp32_kernel<1><<<batch, 256, smem, h->stream>>>(fin, fw, L.bias, fres, fout, L.cin, in_stride, L.cout, 1); break;
    case 3: caz::conv_fp32_kernel<3><<<batch, 256, smem, h->stream>>>(fin, fw, L.bias, fres, fout, L.cin, in_stride, L.cout, 1); break;
    case 5: caz::conv_fp32_kernel<5><<<batch, 256, smem, h->stream>>>(fin, fw, L.bias, fres, fout, L.cin, in_stride, L.cout, 1); break;
    case 7: caz::conv_fp32_kernel<7><<<batch, 256, smem, h->stream>>>(fin, fw, L.bias, fres, fout, L.cin, in_stride, L.cout, 1); break;
    default: return fail(h, CAZ_ERR_INVALID, "unsupported kernel size");
  }
  return check_launch(h, "conv_fp32_kernel");
}

// ---- small-batch persistent tower (conv_sb.cuh) -------------------------------------------------------
constexpr int FLAG_CANDIDATES = 512;   // candidate 128-byte lines for the progress flags

// ONE thread on one SM: for every candidate line, how long until a flag published the way the kernel does it
// (L2 atomic) is visible to a gpu-scope poll from the same SM.  Lines homed in the other die's L2 take about twice as
// long (measured on B200: ~340 vs ~720 cycles), which splits the candidates into the two dies.
__global__ void flag_line_probe_kernel(unsigned* cand, unsigned* out_cycles) {
  for (int c = 0; c < FLAG_CANDIDATES; ++c) {
    unsigned* p = cand + static_cast<size_t>(c) * 32 + 31;
    unsigned best = 0xffffffffu;
    for (unsigned trial = 1; trial <= 4; ++trial) {
      const long long t0 = clock64();
      caz::ptx::publish_flag_gpu(p, trial);
      unsigned spins = 0;
      while (caz::ptx::ld_relaxed_gpu(p) != trial && ++spins < (1u << 20)) {}
      const unsigned d = static_cast<unsigned>(clock64() - t0);
      best = d < best ? d : best;
    }
    caz::ptx::publish_flag_gpu(p, 0u);
    out_cycles[c] = best;
  }
}

// Which of the two latency classes is the NEAR one for each SM?  One CTA per SM (the shared-memory request keeps them
// apart on an idle device) times publish -> visible on a line of each class, best of 32 rounds, and records the winner
// under its %smid.  The forward kernel used to decide this itself during layer 0 with two rounds per launch, and a CTA
// that guessed wrong slowed its whole board tile by 0.4 us per layer (5-7 us per pass, in ~half of the launches).
__global__ void sm_near_probe_kernel(unsigned* cand, const uint16_t* line0, const uint16_t* line1, int n_lines, unsigned char* near_of_sm,
                                     unsigned* cycles_of_sm, unsigned* turn) {
  if (threadIdx.x != 0) return;
  // one SM at a time (CTAs are dispatched in block order, so waiting for the lower block ids cannot deadlock; the wait is
  // bounded all the same): concurrent probes of the same lines made a few SMs look equidistant from both classes
  for (unsigned spins = 0; caz::ptx::ld_acquire_gpu(turn) != blockIdx.x && spins < (1u << 22); ++spins) __nanosleep(200);
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  unsigned best[2] = {0xffffffffu, 0xffffffffu};
  for (unsigned round = 1; round <= 32; ++round)
    for (int c = 0; c < 2; ++c) {
      unsigned* p = cand + static_cast<size_t>((c ? line1 : line0)[round % (n_lines < 8 ? n_lines : 8)]) * 32 + 30;
      const unsigned val = 0x5a000000u + round + (blockIdx.x << 8);
      const long long t0 = clock64();
      caz::ptx::publish_flag_gpu(p, val);
      unsigned spins = 0;
      while (caz::ptx::ld_relaxed_gpu(p) != val && ++spins < (1u << 20)) {}
      const unsigned d = static_cast<unsigned>(clock64() - t0);
      best[c] = d < best[c] ? d : best[c];
    }
  near_of_sm[smid] = best[1] < best[0] ? 1 : 0;
  cycles_of_sm[2 * smid] = best[0];
  cycles_of_sm[2 * smid + 1] = best[1];
  __threadfence();
  atomicAdd(turn, 1u);
}

// Lays out the two copies of the per-tile flag lines (conv_sb.cuh, Args::flag_line) on lines of the two latency
// classes.  Without a clear split (one die, or noise) both copies still get distinct lines: nothing depends on the
// split being right except speed.
int calibrate_flag_copies(caz_handle* h) {
  unsigned* d_cyc = nullptr;
  CU_TRY(h, cudaMalloc(&d_cyc, FLAG_CANDIDATES * sizeof(unsigned)));
  for (int rep = 0; rep < 2; ++rep) flag_line_probe_kernel<<<1, 1, 0, h->stream>>>(h->sb_counter, d_cyc);
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  std::vector<unsigned> cyc(FLAG_CANDIDATES);
  CU_TRY(h, cudaMemcpy(cyc.data(), d_cyc, cyc.size() * sizeof(unsigned), cudaMemcpyDeviceToHost));
  cudaFree(d_cyc);
  std::vector<unsigned> sorted(cyc);
  std::sort(sorted.begin(), sorted.end());
  unsigned gap = 0, split = 0;
  for (int i = FLAG_CANDIDATES / 8; i + 1 < FLAG_CANDIDATES - FLAG_CANDIDATES / 8; ++i)
    if (sorted[i + 1] - sorted[i] > gap) { gap = sorted[i + 1] - sorted[i]; split = (sorted[i + 1] + sorted[i]) / 2; }
  h->sb_two_dies = gap * 4 > sorted[FLAG_CANDIDATES / 2];   // a jump of more than a quarter of the median
  int n[2] = {0, 0};
  for (int c = 0; c < FLAG_CANDIDATES; ++c) {
    const int cls = h->sb_two_dies ? (cyc[c] > split ? 1 : 0) : (c & 1);
    if (n[cls] < caz::sb::MAX_TILES) h->sb_flag_line[cls][n[cls]++] = (uint16_t)c;
  }
  if (n[0] < caz::sb::MAX_TILES || n[1] < caz::sb::MAX_TILES) {   // lopsided: fall back to alternating lines
    h->sb_two_dies = false;
    for (int t = 0; t < caz::sb::MAX_TILES; ++t) { h->sb_flag_line[0][t] = (uint16_t)(2 * t); h->sb_flag_line[1][t] = (uint16_t)(2 * t + 1); }
  }
  if (getenv("CAZ_SB_VERBOSE"))
    fprintf(stderr, "caz: flag lines: two latency classes=%d (gap %u, split %u, min %u, median %u, max %u)\n", (int)h->sb_two_dies, gap,
            split, sorted.front(), sorted[FLAG_CANDIDATES / 2], sorted.back());
  // per-SM table of the nearer class (only with a clear split; otherwise the forward kernel probes for itself)
  if (h->sb_two_dies && !getenv("CAZ_SB_NO_SM_TABLE")) {
    const int n_sm_max = 1024;
    uint16_t* d_lines = nullptr;
    unsigned* d_c = nullptr;
    CU_TRY(h, cudaMalloc(&h->sb_sm_near, n_sm_max));
    CU_TRY(h, cudaMemset(h->sb_sm_near, 0xFF, n_sm_max));
    CU_TRY(h, cudaMalloc(&d_lines, 2 * caz::sb::MAX_TILES * sizeof(uint16_t)));
    CU_TRY(h, cudaMalloc(&d_c, (2 * n_sm_max + 1) * sizeof(unsigned)));
    CU_TRY(h, cudaMemset(d_c, 0, (2 * n_sm_max + 1) * sizeof(unsigned)));
    CU_TRY(h, cudaMemcpy(d_lines, h->sb_flag_line, 2 * caz::sb::MAX_TILES * sizeof(uint16_t), cudaMemcpyHostToDevice));
    const int probe_smem = 160 * 1024;   // one CTA per SM
    CU_TRY(h, cudaFuncSetAttribute(sm_near_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, probe_smem));
    sm_near_probe_kernel<<<h->num_sms, 32, probe_smem, h->stream>>>(h->sb_counter, d_lines, d_lines + caz::sb::MAX_TILES, caz::sb::MAX_TILES,
                                                                    h->sb_sm_near, d_c, d_c + 2 * n_sm_max);
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    if (getenv("CAZ_SB_VERBOSE")) {
      std::vector<unsigned char> nr(n_sm_max);
      std::vector<unsigned> cy(2 * n_sm_max);
      CU_TRY(h, cudaMemcpy(nr.data(), h->sb_sm_near, n_sm_max, cudaMemcpyDeviceToHost));
      CU_TRY(h, cudaMemcpy(cy.data(), d_c, cy.size() * sizeof(unsigned), cudaMemcpyDeviceToHost));
      int cnt[3] = {0, 0, 0};
      unsigned min_margin = 0xffffffffu;
      for (int i = 0; i < n_sm_max; ++i) {
        cnt[nr[i] <= 1 ? nr[i] : 2]++;
        if (nr[i] <= 1) { const unsigned m = cy[2 * i] > cy[2 * i + 1] ? cy[2 * i] - cy[2 * i + 1] : cy[2 * i + 1] - cy[2 * i]; if (m < min_margin) min_margin = m; }
      }
      fprintf(stderr, "caz: per-SM near class: %d SMs -> class 0, %d -> class 1, smallest margin %u cycles\n", cnt[0], cnt[1], min_margin);
      if (getenv("CAZ_SB_VERBOSE")[0] == '2')
        for (int i = 0; i < h->num_sms + 8; ++i) fprintf(stderr, "  sm %d: %u / %u -> %d\n", i, cy[2 * i], cy[2 * i + 1], (int)nr[i]);
    }
    cudaFree(d_lines);
    cudaFree(d_c);
  }
  return CAZ_OK;
}

int setup_small_batch(caz_handle* h) {
  const caz_arch& a = h->arch;
  h->sb_ok = false;
  if (!tensor_path(h) || a.block_kernel != 3 || a.filters != 256 && a.filters != 128 && a.filters != 64) return CAZ_OK;
  if (a.input_planes > 64 || (8 + a.stem_kernel - 1) * 2 * (8 + a.stem_kernel - 1) * 128 > caz::sb::A_BYTES) return CAZ_OK;
  const int sb_cap = caz::sb::MAX_BATCH + 2;
  int rc;
  for (int nb = 1; nb <= 2; ++nb) {
    if ((rc = make_halo_tmap(h, &h->sb_tm_xb[nb - 1], h->act[0], a.filters, a.block_kernel, nb))) return rc;
    if ((rc = make_halo_tmap(h, &h->sb_tm_y[nb - 1], h->act[1], a.filters, a.block_kernel, nb))) return rc;
  }
  if (const char* e = getenv("CAZ_SB_BOARDS")) h->sb_force_nb = atoi(e);
  const int F = a.filters;
  for (int i = 0; i < 4; ++i) {
    const cuuint32_t ns = 16u << i;
    if (ns > (cuuint32_t)F) { h->sb_tm_w_stem[i] = h->sb_tm_w_stem[0]; h->sb_tm_w_tower[i] = h->sb_tm_w_tower[0]; continue; }
    const cuuint64_t kst = (cuuint64_t)a.stem_kernel * a.stem_kernel * 64;
    cuuint64_t dims_s[2] = {kst, (cuuint64_t)F};
    cuuint64_t str_s[1] = {kst * 2};
    cuuint32_t box[2] = {64, ns};
    if ((rc = make_tmap(h, &h->sb_tm_w_stem[i], h->sb_w_stem, 2, dims_s, str_s, box))) return rc;
    if (a.res_blocks > 0) {
      const cuuint64_t kt = (cuuint64_t)9 * F;
      cuuint64_t dims_t[2] = {kt, (cuuint64_t)F * 2 * a.res_blocks};
      cuuint64_t str_t[1] = {kt * 2};
      if ((rc = make_tmap(h, &h->sb_tm_w_tower[i], h->w_tower, 2, dims_t, str_t, box))) return rc;
    } else {
      h->sb_tm_w_tower[i] = h->sb_tm_w_stem[i];
    }
  }
  CU_TRY(h, cudaMalloc(&h->sb_hpart, (size_t)sb_cap * 16 * caz::sb::HEAD_FILTERS_MAX * 64 * sizeof(float)));
  CU_TRY(h, cudaMalloc(&h->sb_hstats, (size_t)sb_cap * caz::sb::HSTAT_FLOATS * sizeof(float)));
  CU_TRY(h, cudaMalloc(&h->sb_counter, FLAG_CANDIDATES * 128));
  CU_TRY(h, cudaMemset(h->sb_counter, 0, FLAG_CANDIDATES * 128));
  if ((rc = calibrate_flag_copies(h))) return rc;
  CU_TRY(h, cudaMalloc(&h->sb_trace, (1024 + 16 * 256 * 2) * sizeof(long long)));
  CU_TRY(h, cudaMemset(h->sb_trace, 0, (1024 + 16 * 256 * 2) * sizeof(long long)));
  CU_TRY(h, cudaFuncSetAttribute(caz::sb::forward_sb_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, caz::sb::SMEM_BYTES));
  CU_TRY(h, cudaFuncSetAttribute(caz::sb::forward_sb_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, caz::sb::SMEM_BYTES));
  CU_TRY(h, cudaFuncSetAttribute(caz::sb::forward_sb_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, caz::sb::SMEM_BYTES));
  h->sb_ok = true;
  return CAZ_OK;
}

// Grid geometry for a batch: boards per CTA tile (1 -> M = 64 MMAs, 2 -> M = 128) and output-channel slices, such that
// every CTA is co-resident.
bool sb_geometry(const caz_handle* h, int batch, int* nb_out, int* slices_out, int* pair_out = nullptr) {
  if (!h->sb_ok) return false;
  if (pair_out) *pair_out = 0;
  // Up to 72 boards: CTA pairs (cta_group::2, one board per CTA; a 32-channel slice per pair up to 18 boards, 64 channels up
  // to 36, 128 above).  Per MMA a CTA reads its 64 activation rows plus HALF of the slice's weight rows and stages half of the
  // slice's weights: 484 KB of shared-memory traffic per layer instead of 630 KB at 16 boards, 630 KB
  // instead of 969 KB at 32; and all 128 epilogue threads have rows (an M = 64 single-CTA accumulator leaves half idle),
  // which is why the pair form also wins below 9 boards (measured: 111 us vs 115 us).
  if (batch >= 1 && !h->sb_force_nb && !h->sb_no_pairs) {
    const caz_arch& a = h->arch;
    for (int ns = 32; ns <= 128; ns *= 2) {   // 128-channel slices: 37-72 boards
      const int s = a.filters / ns;
      if (a.filters % ns != 0 || s < 1 || s > 16 || ((batch + 1) / 2) * s * 2 > h->num_sms) continue;
      // (with up to 4 slices the two CTAs of a pair split their slice's labels and hidden units: conv_sb.cuh, hshare)
      if (!caz::sb::heads_fit(a.policy_filters, a.value_filters, a.n_labels, a.value_fc, s <= 4 ? 2 * s : s)) continue;
      *nb_out = 1;
      *slices_out = s;
      if (pair_out) *pair_out = 1;
      return true;
    }
  }
  for (int pass = 0; pass < 2; ++pass)
    for (int nb = 1; nb <= 2; ++nb) {
      if (h->sb_force_nb && nb != h->sb_force_nb) continue;
      const int tiles = (batch + nb - 1) / nb;
      for (int s = 16; s >= 2; s >>= 1) {   // (two 128-channel slices of two-board tiles: 73-148 boards)
        const int ns = h->arch.filters / s;
        if (ns < 16 || ns > 128 || h->arch.filters % s != 0 || tiles * s > h->num_sms) continue;
        if (!caz::sb::heads_fit(h->arch.policy_filters, h->arch.value_filters, h->arch.n_labels, h->arch.value_fc, s)) continue;
        // first pass: one board per tile (M = 64: the tensor pipe is bound by operand reads from shared memory,
        // (M + N) * 32 B per MMA, tools/ub/ub_small.cu) as long as that leaves slices of at most 32 channels
        if (pass == 0 && nb == 1 && s < 8) break;
        *nb_out = nb;
        *slices_out = s;
        return true;
      }
    }
  return false;
}

// Batches up to here take the persistent small-batch kernel unless told otherwise: 37 two-board tiles x 4 slices of 64 channels
// = 148 CTAs.  (caz_set_small_batch_limit(h, 148) extends it to two 128-channel slices per tile: measured 281 us at 128 boards,
// no better than the one-launch tower -- 15.3 us per layer, of which 5 us epilogue over 128 channels per thread -- so not automatic.)
constexpr int SB_AUTO_LIMIT = 74;
bool use_small_batch(const caz_handle* h, int batch) {
  const int limit = h->sb_limit < 0 ? SB_AUTO_LIMIT : h->sb_limit;
  int nb, sl;
  return batch <= limit && batch <= caz::sb::MAX_BATCH && sb_geometry(h, batch, &nb, &sl);
}

// The whole forward pass of a small batch: one cooperative launch (conv_sb.cuh).
int launch_forward_small(caz_handle* h, const float* d_planes, const uint8_t* d_records, int batch, float* d_policy,
                         float* d_value, const uint8_t* d_mask) {
  const caz_arch& a = h->arch;
  int nb = 2, slices = 4, pair = 0;
  sb_geometry(h, batch, &nb, &slices, &pair);
  const int ns = a.filters / slices;
  const int ns_cta = pair ? ns / 2 : ns;   // weight rows one CTA stages
  const int wi = ns_cta == 16 ? 0 : (ns_cta == 32 ? 1 : (ns_cta == 64 ? 2 : 3));
  caz::sb::Args args;
  args.batch = batch;
  args.nb = nb;
  args.pair = pair;
  args.n_layers = 1 + 2 * a.res_blocks;
  args.ns = ns;
  args.n_slices = slices;
  args.filters = a.filters;
  args.stem_k = a.stem_kernel;
  args.stem_kchunks = 1;
  args.blk_k = a.block_kernel;
  args.in_planes = a.input_planes;
  args.in_cpad = 64;
  args.fp16 = operand_fmt(h);
  args.dbg = h->sb_dbg;
  const int sb_stem = caz::sb::taps_per_slot(a.stem_kernel, ns_cta) * ns_cta * 128;
  const int sb_blk = a.res_blocks > 0 ? caz::sb::taps_per_slot(a.block_kernel, ns_cta) * ns_cta * 128 : 0;
  args.slot_bytes = sb_stem > sb_blk ? sb_stem : sb_blk;
  args.n_slots = caz::sb::W_RING_BYTES / args.slot_bytes;
  if (args.n_slots > caz::sb::MAX_W_SLOTS) args.n_slots = caz::sb::MAX_W_SLOTS;
  args.planes = d_planes;
  args.records = d_records;
  args.bias_all = h->bias_all;
  args.xb = static_cast<__nv_bfloat16*>(h->act[0]);
  args.y = static_cast<__nv_bfloat16*>(h->act[1]);
  args.pf = a.policy_filters;
  args.vf = a.value_filters;
  args.value_fc = a.value_fc;
  args.n_labels = a.n_labels;
  args.head_w = h->head_w;
  args.head_b = h->head_b;
  args.hpart = h->sb_hpart;
  args.hstats = h->sb_hstats;
  args.pfc_w = h->pfc_w;
  args.pfc_wb = h->pfc_wb;
  args.vfc1_wb = h->vfc1_wb;
  args.pfc_b = h->pfc_b;
  args.vfc1_w = h->vfc1_w;
  args.vfc1_b = h->vfc1_b;
  args.vfc2_w = h->vfc2_w;
  args.vfc2_b = h->vfc2_b;
  args.mask = d_mask;
  args.policy = d_policy;
  args.value = d_value;
  args.flag_base = h->sb_counter;
  args.sm_near = h->sb_sm_near;
  for (int c = 0; c < 2; ++c)
    for (int t = 0; t < caz::sb::MAX_TILES; ++t) args.flag_line[c][t] = h->sb_flag_line[c][t];

  if (h->sb_epoch > 0x40000000u) {   // long before the flag epochs could wrap past a stale tile's value: start over
    CU_TRY(h, cudaMemsetAsync(h->sb_counter, 0, FLAG_CANDIDATES * 128, h->stream));
    h->sb_epoch = 0;
  }
  args.base = h->sb_epoch;
  args.trace = (h->sb_trace_on && args.n_layers <= 64) ? h->sb_trace : nullptr;
  args.err_flag = h->err_flag_dev;
  const int grid = pair ? ((batch + 1) / 2) * slices * 2 : ((batch + nb - 1) / nb) * slices;
  h->sb_epoch += (unsigned)(args.n_layers + caz::sb::FLAG_STEPS_EXTRA);
  void* params[] = {&h->sb_tm_xb[nb - 1], &h->sb_tm_y[nb - 1], &h->sb_tm_w_stem[wi], &h->sb_tm_w_tower[wi], &args};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(caz::sb::THREADS);
  cfg.dynamicSmemBytes = caz::sb::SMEM_BYTES;
  cfg.stream = h->stream;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeCooperative;   // every CTA co-resident: the kernel's rendezvous spin on each other
  attrs[0].val.cooperative = 1;
  attrs[1].id = cudaLaunchAttributeClusterDimension;
  attrs[1].val.clusterDim.x = pair ? 2 : 1;
  attrs[1].val.clusterDim.y = 1;
  attrs[1].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pair ? 2 : 1;
  // Nsight Compute cannot launch a kernel that is cooperative AND clustered (LaunchFailed before the first pass, grid
  // reported as 0x0x0).  Under its injection the pair form is launched with the cluster attribute only: the profiler
  // serialises kernels, so the grid (<= one CTA per SM) is co-resident anyway.
  static const bool under_ncu = getenv("NV_NSIGHT_INJECTION_TRANSPORT_TYPE") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR");
  if (pair && under_ncu) {
    attrs[0] = attrs[1];
    cfg.numAttrs = 1;
  }
  const void* kernel = pair ? reinterpret_cast<const void*>(caz::sb::forward_sb_kernel<true>)
                            : (ns_cta > 64 ? reinterpret_cast<const void*>(caz::sb::forward_sb_kernel<false, 8>)
                                           : reinterpret_cast<const void*>(caz::sb::forward_sb_kernel<false>));
  cudaError_t e = cudaLaunchKernelExC(&cfg, kernel, params);
  if (e != cudaSuccess) return fail(h, CAZ_ERR_CUDA, fmt("cooperative launch of forward_sb_kernel failed: %s", cudaGetErrorString(e)));
  h->launches++;
  return CAZ_OK;
}

// ---- whole tower in one launch (conv_tower.cuh) ---------------------------------------------------------
int setup_tower(caz_handle* h) {
  const caz_arch& a = h->arch;
  h->tw_ok = false;
  if (!tensor_path(h) || !h->use_2cta || a.block_kernel != 3 || a.filters % 64 != 0 || a.filters > 256) return CAZ_OK;
  const int hw = 8 + a.stem_kernel - 1;
  if (hw * 2 * hw * h->in_cpad * 2 > caz::tw::A_SLOT_BYTES) return CAZ_OK;   // the stem's halo tile must fit a ring slot
  if (a.res_blocks > 0) {
    const cuuint64_t kt = (cuuint64_t)9 * a.filters;
    cuuint64_t dims[2] = {kt, (cuuint64_t)a.filters * 2 * a.res_blocks};
    cuuint64_t str[1] = {kt * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)a.filters / 2};
    int rc;
    if ((rc = make_tmap(h, &h->tw_tm_w_tower, h->w_tower, 2, dims, str, box))) return rc;
  } else {
    h->tw_tm_w_tower = h->convs[0].tmap_w2;
  }
  CU_TRY(h, cudaFuncSetAttribute(caz::tw::tower_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, caz::tw::SMEM_BYTES));
  if (!h->tw_scratch) {
    // Scratch of the one-launch tower: every CTA pair rewrites the same <= 8 boards (two groups) layer after layer.  Left to
    // itself the L2 writes those dirty lines back to HBM almost as fast as they are produced (3.1 GB per launch at 4 096
    // boards, although every line is rewritten two layers later); marked PERSISTING they stay: 0.8 GB, and the power the HBM
    // traffic cost goes to the SM clock (+3.8 % positions/s, measured back to back on one box).
    const int boards = std::min((h->max_batch + 3) & ~3, 8 * (h->num_sms / 2));
    const size_t esz = 2, f32 = (size_t)boards * 64 * a.filters * 4, h16 = (size_t)boards * 64 * a.filters * esz;
    h->tw_scratch_boards = boards;
    h->tw_scratch_bytes = f32 + 2 * h16;
    CU_TRY(h, cudaMalloc(&h->tw_scratch, h->tw_scratch_bytes));
    CU_TRY(h, cudaMemset(h->tw_scratch, 0, h->tw_scratch_bytes));
    char* base = static_cast<char*>(h->tw_scratch);
    int rc;
    if ((rc = make_halo_tmap(h, &h->tw_tm_x, base + f32, a.filters, a.block_kernel, 2, boards))) return rc;
    if ((rc = make_halo_tmap(h, &h->tw_tm_y, base + f32 + h16, a.filters, a.block_kernel, 2, boards))) return rc;
    if (!getenv("CAZ_NO_L2_PERSIST")) {
      cudaDeviceProp prop;
      CU_TRY(h, cudaGetDeviceProperties(&prop, h->device));
      size_t cur = 0;
      cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize);
      const size_t lim = std::min(std::max(cur, h->tw_scratch_bytes), (size_t)prop.persistingL2CacheMaxSize);
      cudaStreamAttrValue av = {};
      av.accessPolicyWindow.base_ptr = h->tw_scratch;
      av.accessPolicyWindow.num_bytes = std::min(h->tw_scratch_bytes, (size_t)prop.accessPolicyMaxWindowSize);
      av.accessPolicyWindow.hitRatio = lim >= h->tw_scratch_bytes ? 1.0f : (float)((double)lim / (double)h->tw_scratch_bytes);
      av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      // (a device without the feature, or a refused limit, only costs the speed-up)
      if (lim > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, lim) == cudaSuccess)
        cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av);
      cudaGetLastError();
    }
  }
  if (!h->tw_trace) {
    CU_TRY(h, cudaMalloc(&h->tw_trace, 64 * 8 * sizeof(long long)));
    CU_TRY(h, cudaMemset(h->tw_trace, 0, 64 * 8 * sizeof(long long)));
  }
  h->tw_ok = true;
  return CAZ_OK;
}

bool use_tower(const caz_handle* h, int batch) { return h->tw_ok && h->tw_enabled && batch >= 1; }

int launch_tower(caz_handle* h, int batch) {
  const caz_arch& a = h->arch;
  caz::tw::TowerArgs args{};
  args.batch = batch;
  args.n_layers = 1 + 2 * a.res_blocks;
  args.stem_k = a.stem_kernel;
  args.stem_row_bytes = h->in_cpad * 2;
  args.n = a.filters;
  args.fmt = operand_fmt(h);
  args.pdl = h->pdl ? 1 : 0;   // chained to the pack kernel before it and the head kernels after it
  args.bias_all = h->bias_all;
  {
    char* base = static_cast<char*>(h->tw_scratch);
    const size_t f32 = (size_t)h->tw_scratch_boards * 64 * a.filters * 4, h16 = (size_t)h->tw_scratch_boards * 64 * a.filters * 2;
    args.xf = reinterpret_cast<float*>(base);
    args.act0 = reinterpret_cast<__nv_bfloat16*>(base + f32);
    args.act1 = reinterpret_cast<__nv_bfloat16*>(base + f32 + h16);
  }
  args.head_w = h->head_w;
  args.head_b = h->head_b;
  args.featp = h->featp;
  args.featv = h->featv;
  args.pf = a.policy_filters;
  args.vf = a.value_filters;
  args.err_flag = h->err_flag_dev;
  args.trace = (h->tw_trace_on && args.n_layers <= 64) ? h->tw_trace : nullptr;
  const int groups = ((batch + 1) / 2 + 1) / 2, max_pairs = h->num_sms / 2;
  // Enough boards for every pair to hold two groups: interleave them (the epilogue of one runs under the other's MMAs).
  // Fewer: one group per pair on as many pairs as there are groups -- an exposed epilogue costs less than an idle SM.
  // (Cutting every layer of a lone group into two N tiles, first half drained and consumed while the second is computed,
  // was built and measured in round 2: bit-identical, and not a microsecond faster -- 283.6 vs 286.1 us at 100 boards.)
  args.interleave = h->tw_mode ? h->tw_mode : (groups > max_pairs ? 2 : 1);
  args.all_lanes_fence = h->tw_all_lanes_fence ? 1 : 0;
  const int supers = (groups + args.interleave - 1) / args.interleave;
  const int pairs = supers < max_pairs ? supers : max_pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(caz::tw::THREADS);
  cfg.dynamicSmemBytes = caz::tw::SMEM_BYTES;
  cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = h->pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, caz::tw::tower_kernel, h->tmap_in, h->tw_tm_x, h->tw_tm_y, h->convs[0].tmap_w2,
                                     h->tw_tm_w_tower, args);
  if (e != cudaSuccess) return fail(h, CAZ_ERR_CUDA, fmt("launch of tower_kernel failed: %s", cudaGetErrorString(e)));
  return check_launch(h, "tower_kernel");
}

// Picks the event set for this forward pass (no synchronisation).
int prof_begin(caz_handle* h) {
  h->ev = nullptr;
  if (!h->profiling) return CAZ_OK;
  if (h->ev_used == h->ev_pool.size()) {
    caz_handle::EvSet set{};
    for (auto& e : set.ev) CU_TRY(h, cudaEventCreate(&e));
    h->ev_pool.push_back(set);
  }
  h->ev = h->ev_pool[h->ev_used].ev;
  h->pending_launches = h->ev_pool[h->ev_used].launches;
  h->ev_used++;
  return CAZ_OK;
}

// Waits for the recorded forward passes and folds their stage times into the accumulators.
int prof_drain(caz_handle* h) {
  if (h->ev_used == 0) return CAZ_OK;
  CU_TRY(h, cudaEventSynchronize(h->ev_pool[h->ev_used - 1].ev[CAZ_STAGES]));
  for (size_t i = 0; i < h->ev_used; ++i)
    for (int s = 0; s < CAZ_STAGES; ++s) {
      float ms = 0;
      CU_TRY(h, cudaEventElapsedTime(&ms, h->ev_pool[i].ev[s], h->ev_pool[i].ev[s + 1]));
      h->stage_ms[s] += ms;
      h->stage_launches[s] += h->ev_pool[i].launches[s];
    }
  h->ev_used = 0;
  return CAZ_OK;
}

// ---- heads' Dense layers as tensor-core GEMMs (conv_tc2_kernel in GEMM mode) -----------------------------
uint16_t round16(float v, bool f16) {
  uint16_t bits;
  if (f16) { const __half hv = __float2half_rn(v); memcpy(&bits, &hv, 2); }
  else { const __nv_bfloat16 bv = __float2bfloat16_rn(v); memcpy(&bits, &bv, 2); }
  return bits;
}
float widen16(uint16_t bits, bool f16) {
  if (f16) { __half hv; memcpy(&hv, &bits, 2); return __half2float(hv); }
  __nv_bfloat16 bv; memcpy(&bv, &bits, 2); return __bfloat162float(bv);
}

// W [k][n] fp32 (Keras Dense kernel) -> [npad][3k] 16-bit rows = [wh | wl | wh]
int upload_w3(caz_handle* h, uint16_t* dst, const float* w, int k, int n, int npad) {
  const bool f16 = w_fp16(h);
  std::vector<uint16_t> w3((size_t)npad * 3 * k, 0);
  for (int col = 0; col < n; ++col)
    for (int kk = 0; kk < k; ++kk) {
      const float v = w[(size_t)kk * n + col];
      const uint16_t hi = round16(v, f16), lo = round16(v - widen16(hi, f16), f16);
      uint16_t* row = w3.data() + (size_t)col * 3 * k;
      row[kk] = hi;
      row[k + kk] = lo;
      row[2 * k + kk] = hi;
    }
  return upload(h, dst, w3.data(), w3.size() * 2);
}

int setup_dense_tc(caz_handle* h, const float* pfc_w, const float* pfc_b, const float* vfc1_w) {
  const caz_arch& a = h->arch;
  const int kp = a.policy_filters * 64, kv = a.value_filters * 64;
  h->dense_tc = tensor_path(h) && a.value_fc % 64 == 0 && a.value_fc <= 256 && h->use_2cta;
  if (!h->dense_tc) return CAZ_OK;
  int rc;
  const bool first = h->a3p == nullptr;
  if (first) {
    h->mpad = ((h->max_batch + 127) / 128) * 128;
    h->np_pad = ((a.n_labels + 255) / 256) * 256;
    CU_TRY(h, cudaMalloc(&h->a3p, (size_t)h->mpad * 3 * kp * 2));
    CU_TRY(h, cudaMalloc(&h->a3v, (size_t)h->mpad * 3 * kv * 2));
    CU_TRY(h, cudaMalloc(&h->w3p, (size_t)h->np_pad * 3 * kp * 2));
    CU_TRY(h, cudaMalloc(&h->w3v, (size_t)a.value_fc * 3 * kv * 2));
    CU_TRY(h, cudaMalloc(&h->bias3p, (size_t)h->np_pad * 4));
    // A operands: rows in groups of 64 play the role of boards; ksize 1 -> the "halo" tile is the plain 128 x 64 tile
    if ((rc = make_halo_tmap(h, &h->tm_a3p, h->a3p, 3 * kp, 1, 2, h->mpad / 64))) return rc;
    if ((rc = make_halo_tmap(h, &h->tm_a3v, h->a3v, 3 * kv, 1, 2, h->mpad / 64))) return rc;
    cuuint64_t dp[2] = {(cuuint64_t)3 * kp, (cuuint64_t)h->np_pad}, sp[1] = {(cuuint64_t)3 * kp * 2};
    cuuint32_t bp[2] = {64, 128};
    if ((rc = make_tmap(h, &h->tm_w3p, h->w3p, 2, dp, sp, bp))) return rc;
    cuuint64_t dv[2] = {(cuuint64_t)3 * kv, (cuuint64_t)a.value_fc}, sv[1] = {(cuuint64_t)3 * kv * 2};
    cuuint32_t bv[2] = {64, (cuuint32_t)a.value_fc / 2};
    if ((rc = make_tmap(h, &h->tm_w3v, h->w3v, 2, dv, sv, bv))) return rc;
  }
  if ((rc = upload_w3(h, h->w3p, pfc_w, kp, a.n_labels, h->np_pad))) return rc;
  if ((rc = upload_w3(h, h->w3v, vfc1_w, kv, a.value_fc, a.value_fc))) return rc;
  std::vector<float> b3(h->np_pad, 0.0f);
  memcpy(b3.data(), pfc_b, (size_t)a.n_labels * 4);
  return upload(h, h->bias3p, b3.data(), b3.size() * 4);
}

// out[m][n_total] = act(in[m][k] . W + bias) through the CTA-pair tensor-core kernel
int launch_dense_tc(caz_handle* h, const float* in, uint16_t* a3, const CUtensorMap* tm_a, const CUtensorMap* tm_w, int k,
                    int n_tile, int n_tiles, int n_total, const float* bias, int relu, float* out, int m) {
  int rc;
  const size_t total = (size_t)(((m + 127) / 128) * 128) * k;
  launch_chained(caz::split3_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, h->stream, h->pdl, in, a3, m,
                 ((m + 127) / 128) * 128, k, act_fp16(h) ? 1 : 0);
  if ((rc = check_launch(h, "split3_kernel"))) return rc;
  caz::tc::ConvArgs args{};
  args.batch = (m + 63) / 64;          // "boards" = groups of 64 rows
  args.num_tiles = (args.batch + 1) / 2;
  args.ksize = 1;
  args.row_bytes = 128;
  args.kchunks = 3 * k / 64;
  args.n = n_tile;
  args.relu = relu;
  args.bias = bias;
  args.fp16 = operand_fmt(h);
  args.n_tiles = n_tiles;
  args.n_total = n_total;
  args.m_rows = m;
  args.out_rm = out;
  args.err_flag = h->err_flag_dev;
  args.pdl = h->pdl ? 1 : 0;
  const int work = ((args.num_tiles + 1) / 2) * n_tiles, max_pairs = h->num_sms / 2;
  const int grid = 2 * (work < max_pairs ? work : max_pairs);
  launch_chained(caz::tc2::conv_tc2_kernel, dim3(grid), dim3(caz::tc2::THREADS), caz::tc2::SMEM_BYTES, h->stream, h->pdl, *tm_a, *tm_w, *tm_w, args);
  return check_launch(h, "conv_tc2_kernel (dense)");
}

// Forward pass over device-resident input. Exactly one of d_planes / d_records is non-null.
int forward(caz_handle* h, const float* d_planes, const uint8_t* d_records, int batch, float* d_policy, float* d_value,
            const uint8_t* d_mask) {
  const caz_arch& a = h->arch;
  const bool tcp = tensor_path(h);
  int rc;
  if ((rc = prof_begin(h))) return rc;
  const bool prof = h->ev != nullptr;
  int64_t l0 = h->launches;
  if (prof) CU_TRY(h, cudaEventRecord(h->ev[0], h->stream));
  if (tcp && use_small_batch(h, batch)) {
    // one persistent kernel does pack + stem + tower + heads; its time is booked under the tower stage
    if (prof) { CU_TRY(h, cudaEventRecord(h->ev[1], h->stream)); CU_TRY(h, cudaEventRecord(h->ev[2], h->stream)); h->pending_launches[0] = h->pending_launches[1] = 0; }
    if ((rc = launch_forward_small(h, d_planes, d_records, batch, d_policy, d_value, d_mask))) return rc;
    if (prof) { CU_TRY(h, cudaEventRecord(h->ev[3], h->stream)); CU_TRY(h, cudaEventRecord(h->ev[4], h->stream)); h->pending_launches[2] = 1; h->pending_launches[3] = 0; }
    return CAZ_OK;
  }
  // ---- stage 0: input pack (NCHW fp32 planes -> NHWC activations, channels zero-padded) ----
  if (d_planes) {
    const size_t smem = (size_t)a.input_planes * 65 * sizeof(float);
    if (tcp && act_fp16(h)) caz::pack_planes_kernel<__half><<<batch, 256, smem, h->stream>>>(d_planes, static_cast<__half*>(h->in_act), batch, a.input_planes, h->in_cpad);
    else if (tcp) caz::pack_planes_kernel<__nv_bfloat16><<<batch, 256, smem, h->stream>>>(d_planes, static_cast<__nv_bfloat16*>(h->in_act), batch, a.input_planes, h->in_cpad);
    else caz::pack_planes_kernel<float><<<batch, 256, smem, h->stream>>>(d_planes, static_cast<float*>(h->in_act), batch, a.input_planes, h->in_cpad);
    if ((rc = check_launch(h, "pack_planes_kernel"))) return rc;
  } else {
    const size_t total = (size_t)batch * 64 * h->in_cpad;
    const int blocks = (int)((total + 255) / 256);
    if (tcp && act_fp16(h)) caz::encode_pack_kernel<__half><<<blocks, 256, 0, h->stream>>>(d_records, static_cast<__half*>(h->in_act), batch, h->in_cpad);
    else if (tcp) caz::encode_pack_kernel<__nv_bfloat16><<<blocks, 256, 0, h->stream>>>(d_records, static_cast<__nv_bfloat16*>(h->in_act), batch, h->in_cpad);
    else caz::encode_pack_kernel<float><<<blocks, 256, 0, h->stream>>>(d_records, static_cast<float*>(h->in_act), batch, h->in_cpad);
    if ((rc = check_launch(h, "encode_pack_kernel"))) return rc;
  }
  if (prof) { CU_TRY(h, cudaEventRecord(h->ev[1], h->stream)); h->pending_launches[0] = h->launches - l0; l0 = h->launches; }
  float* xf = static_cast<float*>(h->act[2]);  // tensor-core build: fp32 residual stream
  int cur_fp32 = 0;
  if (tcp && use_tower(h, batch)) {
    // stem + tower + the heads' 1x1 convolutions: one persistent launch, booked under the tower stage
    if (prof) { CU_TRY(h, cudaEventRecord(h->ev[2], h->stream)); h->pending_launches[1] = 0; }
    if ((rc = launch_tower(h, batch))) return rc;
    if (prof) { CU_TRY(h, cudaEventRecord(h->ev[3], h->stream)); h->pending_launches[2] = h->launches - l0; l0 = h->launches; }
  } else {
  // ---- stage 1: stem ----
  // tensor-core build: the LAST convolution's epilogue also computes the heads' 1x1 convs (no activation store)
  const bool stem_is_last = tcp && a.res_blocks == 0;
  if ((rc = launch_conv(h, 0, h->in_act, &h->tmap_in, h->in_cpad, nullptr, stem_is_last ? nullptr : h->act[0],
                        (tcp && !stem_is_last) ? xf : nullptr, batch, stem_is_last))) return rc;
  if (prof) { CU_TRY(h, cudaEventRecord(h->ev[2], h->stream)); h->pending_launches[1] = h->launches - l0; l0 = h->launches; }
  // ---- stage 2: residual tower ----
  int cur = 0;
  for (int blk = 0; blk < a.res_blocks; ++blk) {
    if (tcp) {
      // x (bf16 copy) -> conv1 -> y (bf16);  y -> conv2 (+ fp32 x) -> x in place (fp32) and its bf16 copy
      if ((rc = launch_conv(h, 1 + 2 * blk, h->act[0], &h->tmap_act[0], a.filters, nullptr, h->act[1], nullptr, batch))) return rc;
      const bool last = blk == a.res_blocks - 1;
      if ((rc = launch_conv(h, 2 + 2 * blk, h->act[1], &h->tmap_act[1], a.filters, xf, last ? nullptr : h->act[0],
                            last ? nullptr : xf, batch, last))) return rc;
      continue;
    }
    const int t1 = (cur + 1) % 3, t2 = (cur + 2) % 3;
    if ((rc = launch_conv(h, 1 + 2 * blk, h->act[cur], nullptr, a.filters, nullptr, h->act[t1], nullptr, batch))) return rc;
    if ((rc = launch_conv(h, 2 + 2 * blk, h->act[t1], nullptr, a.filters, static_cast<const float*>(h->act[cur]), h->act[t2], nullptr, batch))) return rc;
    cur = t2;
    cur_fp32 = cur;
  }
  if (prof) { CU_TRY(h, cudaEventRecord(h->ev[3], h->stream)); h->pending_launches[2] = h->launches - l0; l0 = h->launches; }
  }
  // ---- stage 3: heads ----
  const int pf = a.policy_filters, vf = a.value_filters;
  const size_t hsmem = (size_t)(pf + vf) * a.filters * sizeof(float);
  if (!tcp) {
    caz::head_conv_kernel<float><<<batch, 256, hsmem, h->stream>>>(static_cast<const float*>(h->act[cur_fp32]), h->head_w, h->head_b, h->featp, h->featv, batch, a.filters, pf, vf);
    if ((rc = check_launch(h, "head_conv_kernel"))) return rc;
  }
  {
    if (h->dense_tc) {
      if ((rc = launch_dense_tc(h, h->featp, h->a3p, &h->tm_a3p, &h->tm_w3p, pf * 64, 256, h->np_pad / 256, a.n_labels,
                                h->bias3p, 0, d_policy, batch))) return rc;
      if ((rc = launch_dense_tc(h, h->featv, h->a3v, &h->tm_a3v, &h->tm_w3v, vf * 64, a.value_fc, 1, a.value_fc,
                                h->vfc1_b, 1, h->vhid, batch))) return rc;
    } else {
      const dim3 gp((batch + caz::PG_TM - 1) / caz::PG_TM, (a.n_labels + caz::PG_TN - 1) / caz::PG_TN);
      caz::dense_tiled_kernel<<<gp, 256, 0, h->stream>>>(h->featp, h->pfc_w, h->pfc_b, d_policy, batch, pf * 64, a.n_labels, 0);
      if ((rc = check_launch(h, "dense_tiled_kernel (policy)"))) return rc;
      const dim3 gv((batch + caz::PG_TM - 1) / caz::PG_TM, (a.value_fc + caz::PG_TN - 1) / caz::PG_TN);
      caz::dense_tiled_kernel<<<gv, 256, 0, h->stream>>>(h->featv, h->vfc1_w, h->vfc1_b, h->vhid, batch, vf * 64, a.value_fc, 1);
      if ((rc = check_launch(h, "dense_tiled_kernel (value)"))) return rc;
    }
    launch_chained(caz::policy_softmax_kernel, dim3(batch), dim3(256), 0, h->stream, h->pdl && h->dense_tc, d_policy, d_mask, a.n_labels,
                   h->vhid, h->vfc2_w, h->vfc2_b, d_value, a.value_fc);
    if ((rc = check_launch(h, "policy_softmax_kernel"))) return rc;
  }
  if (prof) { CU_TRY(h, cudaEventRecord(h->ev[4], h->stream)); h->pending_launches[3] = h->launches - l0; }
  return CAZ_OK;
}

int ensure_puct(caz_handle* h, size_t bytes) {
  if (bytes <= h->puct_cap) return CAZ_OK;
  if (h->puct_buf) CU_TRY(h, cudaFree(h->puct_buf));
  h->puct_buf = nullptr;
  h->puct_cap = 0;
  const size_t cap = bytes + bytes / 2 + 4096;
  CU_TRY(h, cudaMalloc(&h->puct_buf, cap));
  h->puct_cap = cap;
  return CAZ_OK;
}

size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

int set_device(caz_handle* h) {
  CU_TRY(h, cudaSetDevice(h->device));
  return CAZ_OK;
}

}  // namespace

extern "C" {

int32_t caz_abi_version(void) { return CAZ_ABI_VERSION; }

const char* caz_last_error(const caz_handle* h) {
  if (h) return h->err.c_str();
  std::lock_guard<std::mutex> g(g_err_mutex);
  static thread_local std::string copy;
  copy = g_create_error;
  return copy.c_str();
}

void caz_destroy(caz_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->host_timing && h->ht_calls > 0)
    fprintf(stderr, "caz host timing over %ld synchronous calls: before the launch %.2f us, launch %.2f us, copies back + synchronize %.2f us\n",
            h->ht_calls, h->ht_sum[0] / h->ht_calls * 1e-3, h->ht_sum[1] / h->ht_calls * 1e-3, h->ht_sum[2] / h->ht_calls * 1e-3);
  if (h->w_stem) cudaFree(h->w_stem);
  if (h->sb_w_stem) cudaFree(h->sb_w_stem);
  if (h->w_tower) cudaFree(h->w_tower);
  if (h->bias_all) cudaFree(h->bias_all);
  if (h->sb_counter) cudaFree(h->sb_counter);
  if (h->sb_hpart) cudaFree(h->sb_hpart);
  if (h->sb_hstats) cudaFree(h->sb_hstats);
  if (h->sb_trace) cudaFree(h->sb_trace);
  if (h->tw_trace) cudaFree(h->tw_trace);
  if (h->tw_scratch) cudaFree(h->tw_scratch);
  if (h->sb_sm_near) cudaFree(h->sb_sm_near);
  void* ptrs[] = {h->pfc_wb, h->vfc1_wb, h->head_w, h->head_b, h->pfc_w, h->pfc_b, h->vfc1_w, h->vfc1_b, h->vfc2_w, h->vfc2_b, h->in_act,
                  h->act[0], h->act[1], h->act[2], h->featp, h->featv, h->vhid, h->a3p, h->a3v, h->w3p, h->w3v, h->bias3p, h->d_planes, h->d_policy, h->d_value,
                  h->d_mask, h->d_records, h->puct_buf, h->d_planes2, h->d_policy2, h->d_value2, h->d_mask2,
                  h->d_records2};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (int i = 0; i < 2; ++i) {
    if (h->d_off[i]) cudaFree(h->d_off[i]);
    if (h->d_lab[i]) cudaFree(h->d_lab[i]);
    if (h->d_pri[i]) cudaFree(h->d_pri[i]);
  }
  if (h->err_flag_host) cudaFreeHost(h->err_flag_host);
  for (auto& e : h->ev_done) if (e) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
    if (h->ev_fwd[i]) cudaEventDestroy(h->ev_fwd[i]);
    if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
  }
  if (h->copy_in) cudaStreamDestroy(h->copy_in);
  if (h->copy_out) cudaStreamDestroy(h->copy_out);
  for (auto& set : h->ev_pool) for (auto& e : set.ev) if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int caz_create(const caz_arch* arch, const caz_tensor* tensors, int32_t n_tensors, int32_t device, int32_t precision,
               int32_t max_batch, caz_handle** out) {
  if (!arch || !tensors || !out) return fail(nullptr, CAZ_ERR_INVALID, "null argument");
  *out = nullptr;
  const caz_arch& a = *arch;
  if (precision != CAZ_PRECISION_BF16 && precision != CAZ_PRECISION_FP32 && precision != CAZ_PRECISION_FP16 &&
      precision != CAZ_PRECISION_BF16_PURE)
    return fail(nullptr, CAZ_ERR_INVALID, "precision must be CAZ_PRECISION_BF16, CAZ_PRECISION_BF16_PURE, CAZ_PRECISION_FP16 or CAZ_PRECISION_FP32");
  if (max_batch < 1) return fail(nullptr, CAZ_ERR_INVALID, "max_batch must be >= 1");
  if (a.input_planes < 1 || a.input_planes > 64 || a.res_blocks < 0 || a.filters < 1 || a.policy_filters < 1 ||
      a.value_filters < 1 || a.policy_filters + a.value_filters > caz::HEAD_MAX_F || a.value_fc < 1 ||
      a.n_labels < 1 || a.n_labels > caz::PNJ * 256 || (a.stem_kernel & 1) == 0 || (a.block_kernel & 1) == 0 ||
      a.stem_kernel > 7 || a.block_kernel > 7)
    return fail(nullptr, CAZ_ERR_INVALID, "architecture outside the supported envelope");
  if (precision != CAZ_PRECISION_FP32 && (a.filters % 64 != 0 || a.filters > 256))
    return fail(nullptr, CAZ_ERR_INVALID,
                "the tensor-core build needs filters to be a multiple of 64 and <= 256; use CAZ_PRECISION_FP32");
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count <= 0 || device < 0 || device >= count)
    return fail(nullptr, CAZ_ERR_NO_DEVICE, fmt("no CUDA device %d (%s); this library has no CPU path", device,
                                                ce == cudaSuccess ? "count is 0 or index out of range" : cudaGetErrorString(ce)));
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
    return fail(nullptr, CAZ_ERR_NO_DEVICE, fmt("device %d is not an sm_100 GPU (compute %d.%d)", device, prop.major, prop.minor));
  caz_handle* h = new caz_handle();
  h->arch = a;
  h->device = device;
  h->precision = precision;
  h->max_batch = max_batch;
  h->num_sms = prop.multiProcessorCount;
  auto bail = [&](int rc) {
    fail(nullptr, rc, h->err);
    caz_destroy(h);
    return rc;
  };
#define CREATE_TRY(expr)                                                                                          \
  do {                                                                                                            \
    cudaError_t e__ = (expr);                                                                                     \
    if (e__ != cudaSuccess) {                                                                                     \
      h->err = fmt("%s failed: %s", #expr, cudaGetErrorString(e__));                                              \
      return bail(CAZ_ERR_CUDA);                                                                                  \
    }                                                                                                             \
  } while (0)
  CREATE_TRY(cudaSetDevice(device));
  CREATE_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CREATE_TRY(cudaHostAlloc(&h->err_flag_host, sizeof(int), cudaHostAllocMapped));
  *h->err_flag_host = 0;
  CREATE_TRY(cudaHostGetDevicePointer(&h->err_flag_dev, h->err_flag_host, 0));
  const bool tcp = precision != CAZ_PRECISION_FP32;
  const size_t esz = tcp ? 2 : 4;
  const size_t cap = (size_t)max_batch;
  // 16-bit builds: the stem's input rows are one swizzle span: 32 channels (64-byte rows) when the planes fit, else 64
  h->in_cpad = tcp ? (a.input_planes <= 32 ? 32 : 64) : a.input_planes;
  CREATE_TRY(cudaMalloc(&h->in_act, cap * 64 * h->in_cpad * esz));
  for (int i = 0; i < 3; ++i) {
    // a CTA pair covers four boards (two tiles): the fp32 stream of a pair's second tile is stored even when its boards do
    // not exist, so the buffers hold whole groups of four
    const size_t bytes = ((cap + 3) & ~size_t(3)) * 64 * a.filters * (i == 2 ? 4 : esz);
    CREATE_TRY(cudaMalloc(&h->act[i], bytes));
    // rows of boards past the batch are read (never used) by tiles that hold two boards: keep them finite
    CREATE_TRY(cudaMemset(h->act[i], 0, bytes));
  }
  CREATE_TRY(cudaMalloc(&h->featp, cap * a.policy_filters * 64 * 4));
  CREATE_TRY(cudaMalloc(&h->featv, cap * a.value_filters * 64 * 4));
  CREATE_TRY(cudaMalloc(&h->vhid, cap * a.value_fc * 4));
  CREATE_TRY(cudaMalloc(&h->d_planes, cap * a.input_planes * 64 * 4));
  CREATE_TRY(cudaMalloc(&h->d_policy, cap * a.n_labels * 4));
  CREATE_TRY(cudaMalloc(&h->d_value, cap * 4));
  CREATE_TRY(cudaMalloc(&h->d_mask, cap * a.n_labels));
  CREATE_TRY(cudaMalloc(&h->d_records, cap * CAZ_RECORD_BYTES));
  CREATE_TRY(cudaMalloc(&h->d_planes2, cap * a.input_planes * 64 * 4));
  CREATE_TRY(cudaMalloc(&h->d_policy2, cap * a.n_labels * 4));
  CREATE_TRY(cudaMalloc(&h->d_value2, cap * 4));
  CREATE_TRY(cudaMalloc(&h->d_mask2, cap * a.n_labels));
  CREATE_TRY(cudaMalloc(&h->d_records2, cap * CAZ_RECORD_BYTES));
  CREATE_TRY(cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking));
  CREATE_TRY(cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    CREATE_TRY(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&h->ev_fwd[i], cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < caz_handle::TICKET_RING; ++i) {
    h->ticket_state[i].store(0);
    CREATE_TRY(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
  }
  int rc;
  if (tcp) {
    if ((rc = make_halo_tmap(h, &h->tmap_in, h->in_act, h->in_cpad, a.stem_kernel, 2, 0, h->in_cpad * 2))) return bail(rc);
    for (int i = 0; i < 2; ++i)
      if ((rc = make_halo_tmap(h, &h->tmap_act[i], h->act[i], a.filters, a.block_kernel, 2))) return bail(rc);
    CREATE_TRY(cudaFuncSetAttribute(caz::tc::conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, caz::tc::SMEM_BYTES));
    CREATE_TRY(cudaFuncSetAttribute(caz::tc2::conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, caz::tc2::SMEM_BYTES));
    if (const char* e = getenv("CAZ_CONV_2CTA")) h->use_2cta = atoi(e) != 0;
    if (getenv("CAZ_NO_PDL")) h->pdl = false;
    if (getenv("CAZ_NO_TOWER")) h->tw_enabled = false;
    if (const char* e = getenv("CAZ_TOWER_MODE")) h->tw_mode = atoi(e) >= 1 && atoi(e) <= 2 ? atoi(e) : 0;
    if (getenv("CAZ_NO_SPLIT_N")) h->split_n = false;
    if (const char* e = getenv("CAZ_L2_PREFETCH")) h->l2_prefetch = atoi(e) != 0;
    if (const char* e = getenv("CAZ_SNAKE")) h->snake = atoi(e) != 0;
    h->sb_no_pairs = getenv("CAZ_SB_NO_PAIRS") != nullptr;
    h->no_direct_output = getenv("CAZ_NO_DIRECT_OUTPUT") != nullptr;
    if (const char* e = getenv("CAZ_SB_DBG")) h->sb_dbg = atoi(e);
    if (getenv("CAZ_TOWER_ALL_LANES_FENCE")) h->tw_all_lanes_fence = true;
    if (getenv("CAZ_HOST_TIMING")) h->host_timing = true;
    if (const char* e = getenv("CAZ_DBG_EPI")) h->dbg_epi = atoi(e);
  } else {
    const size_t s3 = (size_t)(8 + a.block_kernel - 1) * (8 + a.block_kernel - 1) * a.filters * 4;
    const size_t s1 = (size_t)(8 + a.stem_kernel - 1) * (8 + a.stem_kernel - 1) * a.input_planes * 4;
    const size_t need = s3 > s1 ? s3 : s1;
    if (need > (size_t)prop.sharedMemPerBlockOptin) { h->err = "fp32 build: board tile does not fit shared memory"; return bail(CAZ_ERR_INVALID); }
    CREATE_TRY(cudaFuncSetAttribute(caz::conv_fp32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    CREATE_TRY(cudaFuncSetAttribute(caz::conv_fp32_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    CREATE_TRY(cudaFuncSetAttribute(caz::conv_fp32_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    CREATE_TRY(cudaFuncSetAttribute(caz::conv_fp32_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
  }
#undef CREATE_TRY
  if ((rc = install_weights(h, tensors, n_tensors))) return bail(rc);
  if ((rc = setup_tower(h))) return bail(rc);
  if ((rc = setup_small_batch(h))) return bail(rc);
  *out = h;
  return CAZ_OK;
}

int caz_reload(caz_handle* h, const caz_tensor* tensors, int32_t n_tensors) {
  if (!h || !tensors) return fail(h, CAZ_ERR_INVALID, "null argument");
  int rc;
  if ((rc = set_device(h))) return rc;
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  return install_weights(h, tensors, n_tensors);
}

int caz_sync(caz_handle* h) {
  if (!h) return CAZ_ERR_INVALID;
  int rc;
  if ((rc = set_device(h))) return rc;
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  return CAZ_OK;
}

int caz_eval_device(caz_handle* h, const float* d_planes, int32_t batch, float* d_policy, float* d_value,
                    const uint8_t* d_legal_mask, uint32_t flags) {
  if (!h) return CAZ_ERR_INVALID;
  if (batch == 0) return CAZ_OK;
  if (!d_planes || !d_policy || !d_value || batch < 0 || batch > h->max_batch)
    return fail(h, CAZ_ERR_INVALID, fmt("caz_eval_device: bad argument (batch %d, max_batch %d)", batch, h->max_batch));
  if ((flags & CAZ_EVAL_MASKED_SOFTMAX) && !d_legal_mask) return fail(h, CAZ_ERR_INVALID, "masked softmax needs a legal mask");
  int rc;
  if ((rc = set_device(h))) return rc;
  return forward(h, d_planes, nullptr, batch, d_policy, d_value, (flags & CAZ_EVAL_MASKED_SOFTMAX) ? d_legal_mask : nullptr);
}

// One chunk (<= max_batch positions) through the three-stream pipeline: H2D on copy_in, forward on the compute
// stream, D2H on copy_out, chained by events; `slot` picks one of the two staging buffer sets.  Nothing here blocks
// the host.  Waiting on an event that was never recorded is a no-op, so first uses need no special case.
struct PriorsRequest {   // caz_submit_records_priors: gather instead of the full-policy read-back
  const int32_t* offsets;  // host [nb + 1]
  const int32_t* labels;   // host [offsets[nb]]
  double* priors;          // host [offsets[nb]] out
};

static int enqueue_pipelined(caz_handle* h, int slot, const float* planes, const uint8_t* records, int32_t nb,
                             float* policy, float* value, const uint8_t* legal_mask, bool masked,
                             const PriorsRequest* pr = nullptr) {
  const caz_arch& a = h->arch;
  const size_t pin = (size_t)a.input_planes * 64;
  float* dpl = slot ? h->d_planes2 : h->d_planes;
  float* dpo = slot ? h->d_policy2 : h->d_policy;
  float* dva = slot ? h->d_value2 : h->d_value;
  uint8_t* dma = slot ? h->d_mask2 : h->d_mask;
  uint8_t* dre = slot ? h->d_records2 : h->d_records;
  int rc;
  // this slot's inputs are free once the last forward pass that read them has finished
  CU_TRY(h, cudaStreamWaitEvent(h->copy_in, h->ev_fwd[slot], 0));
  if (planes) CU_TRY(h, cudaMemcpyAsync(dpl, planes, (size_t)nb * pin * 4, cudaMemcpyHostToDevice, h->copy_in));
  else CU_TRY(h, cudaMemcpyAsync(dre, records, (size_t)nb * CAZ_RECORD_BYTES, cudaMemcpyHostToDevice, h->copy_in));
  if (masked) CU_TRY(h, cudaMemcpyAsync(dma, legal_mask, (size_t)nb * a.n_labels, cudaMemcpyHostToDevice, h->copy_in));
  const int32_t n_edges = pr ? pr->offsets[nb] : 0;
  if (pr) {
    if (!h->d_off[slot]) {   // first use: the staging for CSR labels and priors (11 MB per slot at 4096 boards)
      const size_t cap = (size_t)h->max_batch;
      CU_TRY(h, cudaMalloc(&h->d_off[slot], (cap + 1) * sizeof(int32_t)));
      CU_TRY(h, cudaMalloc(&h->d_lab[slot], cap * CAZ_MAX_MOVES * sizeof(int32_t)));
      CU_TRY(h, cudaMalloc(&h->d_pri[slot], cap * CAZ_MAX_MOVES * sizeof(double)));
    }
    CU_TRY(h, cudaMemcpyAsync(h->d_off[slot], pr->offsets, (size_t)(nb + 1) * 4, cudaMemcpyHostToDevice, h->copy_in));
    if (n_edges > 0) CU_TRY(h, cudaMemcpyAsync(h->d_lab[slot], pr->labels, (size_t)n_edges * 4, cudaMemcpyHostToDevice, h->copy_in));
  }
  CU_TRY(h, cudaEventRecord(h->ev_in[slot], h->copy_in));
  CU_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_in[slot], 0));
  // ... and its outputs once the last D2H that drained them has finished
  CU_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_out[slot], 0));
  if ((rc = forward(h, planes ? dpl : nullptr, planes ? nullptr : dre, nb, dpo, dva, masked ? dma : nullptr))) return rc;
  if (pr && n_edges > 0) {
    // the prior push of select_action_q_and_u (player_chess.py:267-275) on the policy rows that never leave the device
    const int wpb = 8;
    caz::push_priors_kernel<<<(nb + wpb - 1) / wpb, wpb * 32, 0, h->stream>>>(dpo, a.n_labels, h->d_off[slot], nb, h->d_lab[slot], h->d_pri[slot]);
    if ((rc = check_launch(h, "push_priors_kernel"))) return rc;
  }
  CU_TRY(h, cudaEventRecord(h->ev_fwd[slot], h->stream));
  CU_TRY(h, cudaStreamWaitEvent(h->copy_out, h->ev_fwd[slot], 0));
  if (pr) {
    if (n_edges > 0) CU_TRY(h, cudaMemcpyAsync(pr->priors, h->d_pri[slot], (size_t)n_edges * 8, cudaMemcpyDeviceToHost, h->copy_out));
  } else {
    CU_TRY(h, cudaMemcpyAsync(policy, dpo, (size_t)nb * a.n_labels * 4, cudaMemcpyDeviceToHost, h->copy_out));
  }
  CU_TRY(h, cudaMemcpyAsync(value, dva, (size_t)nb * 4, cudaMemcpyDeviceToHost, h->copy_out));
  CU_TRY(h, cudaEventRecord(h->ev_out[slot], h->copy_out));
  return CAZ_OK;
}

// Device-side alias of a caller buffer if it is pinned, mapped host memory (what caz_host_alloc returns), else nullptr.
// Asked on every call (~1 us): an address can change hands between a pinned and a pageable allocation.
static void* mapped_alias(const void* host) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, host) == cudaSuccess && at.type == cudaMemoryTypeHost) return at.devicePointer;
  cudaGetLastError();
  return nullptr;
}

static int eval_host(caz_handle* h, const float* planes, const uint8_t* records, int32_t batch, float* policy,
                     float* value, const uint8_t* legal_mask, uint32_t flags) {
  if (!h) return CAZ_ERR_INVALID;
  if (batch == 0) return CAZ_OK;
  if ((!planes && !records) || !policy || !value || batch < 0) return fail(h, CAZ_ERR_INVALID, "caz_eval: bad argument");
  const bool masked = (flags & CAZ_EVAL_MASKED_SOFTMAX) != 0;
  if (masked && !legal_mask) return fail(h, CAZ_ERR_INVALID, "masked softmax needs a legal mask");
  int rc;
  if ((rc = set_device(h))) return rc;
  const caz_arch& a = h->arch;
  const size_t pin = (size_t)a.input_planes * 64;
  if (batch <= h->max_batch) {
    // CAZ_HOST_TIMING=1: where a synchronous call's host time goes (averages printed by caz_destroy)
    const bool ht = h->host_timing;
    auto now = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e9 + ts.tv_nsec; };
    const double t_in = ht ? now() : 0.0;
    // One pass on one stream, no cross-stream events: the latency path.  (Cutting a batch into pipelined chunks was
    // measured: the smaller launches lose what the copy overlap gains; overlap ACROSS calls is caz_submit/caz_collect.)
    CU_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_out[0], 0));   // a still-draining caz_submit on slot 0
    // (same for board records in pinned + mapped memory: 68 bytes per board, read by the kernel itself)
    const bool sb_direct = tensor_path(h) && !h->no_direct_output && use_small_batch(h, batch);
    const uint8_t* drec = h->d_records;
    if (planes) CU_TRY(h, cudaMemcpyAsync(h->d_planes, planes, (size_t)batch * pin * 4, cudaMemcpyHostToDevice, h->stream));
    else if (sb_direct && (reinterpret_cast<uintptr_t>(records) & 3) == 0 && mapped_alias(records)) drec = static_cast<const uint8_t*>(mapped_alias(records));
    else CU_TRY(h, cudaMemcpyAsync(h->d_records, records, (size_t)batch * CAZ_RECORD_BYTES, cudaMemcpyHostToDevice, h->stream));
    if (masked) CU_TRY(h, cudaMemcpyAsync(h->d_mask, legal_mask, (size_t)batch * a.n_labels, cudaMemcpyHostToDevice, h->stream));
    // MCTS-sized batches whose outputs are pinned + mapped: the persistent kernel stores policy and value straight into
    // the caller's memory (posted PCIe writes under the kernel's tail) instead of two device-to-host copies after it.
    float* dpol = h->d_policy;
    float* dval = h->d_value;
    bool direct = false;
    if (sb_direct) {
      void* ap = mapped_alias(policy);
      void* av = mapped_alias(value);
      if (ap && av) { dpol = static_cast<float*>(ap); dval = static_cast<float*>(av); direct = true; }
    }
    const double t_pre = ht ? now() : 0.0;
    if ((rc = forward(h, planes ? h->d_planes : nullptr, planes ? nullptr : drec, batch, dpol, dval, masked ? h->d_mask : nullptr))) return rc;
    const double t_launched = ht ? now() : 0.0;
    if (!direct) {
      CU_TRY(h, cudaMemcpyAsync(policy, h->d_policy, (size_t)batch * a.n_labels * 4, cudaMemcpyDeviceToHost, h->stream));
      CU_TRY(h, cudaMemcpyAsync(value, h->d_value, (size_t)batch * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    if (ht) {
      const double t_out = now();
      h->ht_sum[0] += t_pre - t_in; h->ht_sum[1] += t_launched - t_pre; h->ht_sum[2] += t_out - t_launched; ++h->ht_calls;
    }
    return CAZ_OK;
  }
  // larger than the device buffers: max_batch-sized chunks, alternating staging slots, copies under compute
  int c = 0;
  for (int32_t off = 0; off < batch; ++c) {
    const int32_t nb = batch - off < h->max_batch ? batch - off : h->max_batch;
    if ((rc = enqueue_pipelined(h, c & 1, planes ? planes + (size_t)off * pin : nullptr,
                                records ? records + (size_t)off * CAZ_RECORD_BYTES : nullptr, nb,
                                policy + (size_t)off * a.n_labels, value + off,
                                masked ? legal_mask + (size_t)off * a.n_labels : nullptr, masked))) return rc;
    off += nb;
  }
  CU_TRY(h, cudaStreamSynchronize(h->copy_out));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  return CAZ_OK;
}

// Common tail of the two submit entry points: claim the next ticket's ring entry, enqueue, record its completion event.
static int submit_common(caz_handle* h, const float* planes, const uint8_t* records, int32_t batch, float* policy, float* value,
                         const uint8_t* legal_mask, bool masked, int32_t* ticket, const PriorsRequest* pr = nullptr) {
  const int32_t t = h->next_ticket;
  const int entry = t % caz_handle::TICKET_RING;
  if (h->ticket_state[entry].load(std::memory_order_acquire) != 0)
    return fail(h, CAZ_ERR_INVALID, fmt("caz_submit: %d submissions are outstanding; caz_collect one first", caz_handle::TICKET_RING));
  int rc;
  if ((rc = enqueue_pipelined(h, t & 1, planes, records, batch, policy, value, legal_mask, masked, pr))) return rc;
  CU_TRY(h, cudaEventRecord(h->ev_done[entry], h->copy_out));
  h->ticket_state[entry].store(1, std::memory_order_release);
  h->next_ticket = t + 1 == (1 << 30) ? 0 : t + 1;   // wraps at a multiple of the ring size
  *ticket = t;
  return CAZ_OK;
}

int caz_submit(caz_handle* h, const float* planes, int32_t batch, float* policy, float* value, const uint8_t* legal_mask,
               uint32_t flags, int32_t* ticket) {
  if (!h) return CAZ_ERR_INVALID;
  if (!planes || !policy || !value || !ticket || batch < 1 || batch > h->max_batch)
    return fail(h, CAZ_ERR_INVALID, fmt("caz_submit: bad argument (batch %d, max_batch %d)", batch, h->max_batch));
  const bool masked = (flags & CAZ_EVAL_MASKED_SOFTMAX) != 0;
  if (masked && !legal_mask) return fail(h, CAZ_ERR_INVALID, "masked softmax needs a legal mask");
  int rc;
  if ((rc = set_device(h))) return rc;
  return submit_common(h, planes, nullptr, batch, policy, value, legal_mask, masked, ticket);
}

int caz_submit_records(caz_handle* h, const uint8_t* records, int32_t batch, float* policy, float* value,
                       const uint8_t* legal_mask, uint32_t flags, int32_t* ticket) {
  if (!h) return CAZ_ERR_INVALID;
  if (!records || !policy || !value || !ticket || batch < 1 || batch > h->max_batch)
    return fail(h, CAZ_ERR_INVALID, fmt("caz_submit_records: bad argument (batch %d, max_batch %d)", batch, h->max_batch));
  const bool masked = (flags & CAZ_EVAL_MASKED_SOFTMAX) != 0;
  if (masked && !legal_mask) return fail(h, CAZ_ERR_INVALID, "masked softmax needs a legal mask");
  if (h->arch.input_planes != CAZ_PLANES) return fail(h, CAZ_ERR_INVALID, "board records need an 18-plane network");
  int rc;
  if ((rc = set_device(h))) return rc;
  return submit_common(h, nullptr, records, batch, policy, value, legal_mask, masked, ticket);
}

int caz_submit_records_priors(caz_handle* h, const uint8_t* records, int32_t batch, const int32_t* offsets, const int32_t* labels,
                              double* priors, float* value, int32_t* ticket) {
  if (!h) return CAZ_ERR_INVALID;
  if (!records || !offsets || !labels || !priors || !value || !ticket || batch < 1 || batch > h->max_batch)
    return fail(h, CAZ_ERR_INVALID, fmt("caz_submit_records_priors: bad argument (batch %d, max_batch %d)", batch, h->max_batch));
  if (h->arch.input_planes != CAZ_PLANES) return fail(h, CAZ_ERR_INVALID, "board records need an 18-plane network");
  if (offsets[0] != 0 || offsets[batch] < 0 || (int64_t)offsets[batch] > (int64_t)h->max_batch * CAZ_MAX_MOVES)
    return fail(h, CAZ_ERR_INVALID, "caz_submit_records_priors: offsets must start at 0 and hold at most max_batch * CAZ_MAX_MOVES edges");
  int rc;
  if ((rc = set_device(h))) return rc;
  const PriorsRequest pr{offsets, labels, priors};
  return submit_common(h, nullptr, records, batch, nullptr, value, nullptr, false, ticket, &pr);
}

int caz_collect(caz_handle* h, int32_t ticket) {
  if (!h) return CAZ_ERR_INVALID;
  if (ticket < 0) return fail(h, CAZ_ERR_INVALID, "caz_collect: bad ticket");
  const int entry = ticket % caz_handle::TICKET_RING;
  if (h->ticket_state[entry].load(std::memory_order_acquire) != 1)
    return fail(h, CAZ_ERR_INVALID, fmt("caz_collect: ticket %d is not outstanding (never issued, or already collected)", ticket));
  int rc;
  if ((rc = set_device(h))) return rc;
  CU_TRY(h, cudaEventSynchronize(h->ev_done[entry]));
  h->ticket_state[entry].store(0, std::memory_order_release);
  return CAZ_OK;
}

int caz_eval(caz_handle* h, const float* planes, int32_t batch, float* policy, float* value, const uint8_t* legal_mask,
             uint32_t flags) {
  return eval_host(h, planes, nullptr, batch, policy, value, legal_mask, flags);
}

int caz_eval_records(caz_handle* h, const uint8_t* records, int32_t batch, float* policy, float* value,
                     const uint8_t* legal_mask, uint32_t flags) {
  return eval_host(h, nullptr, records, batch, policy, value, legal_mask, flags);
}

int caz_encode(caz_handle* h, const uint8_t* records, int32_t batch, float* planes) {
  if (!h) return CAZ_ERR_INVALID;
  if (batch == 0) return CAZ_OK;
  if (!records || !planes || batch < 0) return fail(h, CAZ_ERR_INVALID, "caz_encode: bad argument");
  if (h->arch.input_planes != CAZ_PLANES) return fail(h, CAZ_ERR_INVALID, "caz_encode needs an 18-plane network");
  int rc;
  if ((rc = set_device(h))) return rc;
  for (int32_t off = 0; off < batch; off += h->max_batch) {
    const int32_t nb = batch - off < h->max_batch ? batch - off : h->max_batch;
    CU_TRY(h, cudaMemcpyAsync(h->d_records, records + (size_t)off * CAZ_RECORD_BYTES, (size_t)nb * CAZ_RECORD_BYTES, cudaMemcpyHostToDevice, h->stream));
    const size_t total = (size_t)nb * CAZ_PLANE_FLOATS;
    caz::encode_records_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(h->d_records, h->d_planes, nb);
    if ((rc = check_launch(h, "encode_records_kernel"))) return rc;
    CU_TRY(h, cudaMemcpyAsync(planes + (size_t)off * CAZ_PLANE_FLOATS, h->d_planes, total * 4, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
  }
  return CAZ_OK;
}

int caz_push_priors(caz_handle* h, const float* policy, int32_t n_labels, const int32_t* offsets, int32_t n_nodes,
                    const int32_t* label_index, double* priors) {
  if (!h) return CAZ_ERR_INVALID;
  if (n_nodes == 0) return CAZ_OK;
  if (!policy || !offsets || !label_index || !priors || n_nodes < 0 || n_labels < 1)
    return fail(h, CAZ_ERR_INVALID, "caz_push_priors: bad argument");
  const int32_t n_edges = offsets[n_nodes];
  if (offsets[0] != 0 || n_edges < 0) return fail(h, CAZ_ERR_INVALID, "caz_push_priors: offsets must start at 0");
  if (n_edges == 0) return CAZ_OK;
  int rc;
  if ((rc = set_device(h))) return rc;
  const size_t b_pol = align256((size_t)n_nodes * n_labels * 4), b_off = align256((size_t)(n_nodes + 1) * 4),
               b_idx = align256((size_t)n_edges * 4), b_out = align256((size_t)n_edges * 8);
  if ((rc = ensure_puct(h, b_pol + b_off + b_idx + b_out))) return rc;
  char* base = static_cast<char*>(h->puct_buf);
  float* d_pol = reinterpret_cast<float*>(base);
  int* d_off = reinterpret_cast<int*>(base + b_pol);
  int* d_idx = reinterpret_cast<int*>(base + b_pol + b_off);
  double* d_out = reinterpret_cast<double*>(base + b_pol + b_off + b_idx);
  CU_TRY(h, cudaMemcpyAsync(d_pol, policy, (size_t)n_nodes * n_labels * 4, cudaMemcpyHostToDevice, h->stream));
  CU_TRY(h, cudaMemcpyAsync(d_off, offsets, (size_t)(n_nodes + 1) * 4, cudaMemcpyHostToDevice, h->stream));
  CU_TRY(h, cudaMemcpyAsync(d_idx, label_index, (size_t)n_edges * 4, cudaMemcpyHostToDevice, h->stream));
  const int wpb = 8;
  caz::push_priors_kernel<<<(n_nodes + wpb - 1) / wpb, wpb * 32, 0, h->stream>>>(d_pol, n_labels, d_off, n_nodes, d_idx, d_out);
  if ((rc = check_launch(h, "push_priors_kernel"))) return rc;
  CU_TRY(h, cudaMemcpyAsync(priors, d_out, (size_t)n_edges * 8, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  return CAZ_OK;
}

int caz_puct_select(caz_handle* h, const int32_t* offsets, int32_t n_nodes, const int32_t* n, const double* q,
                    const double* p, const int32_t* sum_n, const double* noise, const uint8_t* is_root, double c_puct,
                    double noise_eps, int32_t* best_edge) {
  if (!h) return CAZ_ERR_INVALID;
  if (n_nodes == 0) return CAZ_OK;
  if (!offsets || !n || !q || !p || !sum_n || !best_edge || n_nodes < 0) return fail(h, CAZ_ERR_INVALID, "caz_puct_select: bad argument");
  const int32_t n_edges = offsets[n_nodes];
  if (offsets[0] != 0 || n_edges < 0) return fail(h, CAZ_ERR_INVALID, "caz_puct_select: offsets must start at 0");
  int rc;
  if ((rc = set_device(h))) return rc;
  const bool with_noise = noise && is_root;
  const size_t ne = (size_t)(n_edges > 0 ? n_edges : 1);
  const size_t b_off = align256((size_t)(n_nodes + 1) * 4), b_n = align256(ne * 4), b_d = align256(ne * 8),
               b_sn = align256((size_t)n_nodes * 4), b_root = align256((size_t)n_nodes), b_best = align256((size_t)n_nodes * 4);
  if ((rc = ensure_puct(h, b_off + b_n + 3 * b_d + b_sn + b_root + b_best))) return rc;
  char* c = static_cast<char*>(h->puct_buf);
  int* d_off = reinterpret_cast<int*>(c); c += b_off;
  int* d_n = reinterpret_cast<int*>(c); c += b_n;
  double* d_q = reinterpret_cast<double*>(c); c += b_d;
  double* d_p = reinterpret_cast<double*>(c); c += b_d;
  double* d_noise = reinterpret_cast<double*>(c); c += b_d;
  int* d_sn = reinterpret_cast<int*>(c); c += b_sn;
  uint8_t* d_root = reinterpret_cast<uint8_t*>(c); c += b_root;
  int* d_best = reinterpret_cast<int*>(c);
  CU_TRY(h, cudaMemcpyAsync(d_off, offsets, (size_t)(n_nodes + 1) * 4, cudaMemcpyHostToDevice, h->stream));
  if (n_edges > 0) {
    CU_TRY(h, cudaMemcpyAsync(d_n, n, (size_t)n_edges * 4, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemcpyAsync(d_q, q, (size_t)n_edges * 8, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemcpyAsync(d_p, p, (size_t)n_edges * 8, cudaMemcpyHostToDevice, h->stream));
    if (with_noise) CU_TRY(h, cudaMemcpyAsync(d_noise, noise, (size_t)n_edges * 8, cudaMemcpyHostToDevice, h->stream));
  }
  CU_TRY(h, cudaMemcpyAsync(d_sn, sum_n, (size_t)n_nodes * 4, cudaMemcpyHostToDevice, h->stream));
  if (with_noise) CU_TRY(h, cudaMemcpyAsync(d_root, is_root, (size_t)n_nodes, cudaMemcpyHostToDevice, h->stream));
  const int wpb = 8;
  caz::puct_select_kernel<<<(n_nodes + wpb - 1) / wpb, wpb * 32, 0, h->stream>>>(
      d_off, n_nodes, d_n, d_q, d_p, d_sn, with_noise ? d_noise : nullptr, with_noise ? d_root : nullptr, c_puct, noise_eps, d_best);
  if ((rc = check_launch(h, "puct_select_kernel"))) return rc;
  CU_TRY(h, cudaMemcpyAsync(best_edge, d_best, (size_t)n_nodes * 4, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  return CAZ_OK;
}

int caz_puct_select_device(caz_handle* h, const int32_t* d_offsets, int32_t n_nodes, const int32_t* d_n, const double* d_q,
                            const double* d_p, const int32_t* d_sum_n, const double* d_noise, const uint8_t* d_is_root, double c_puct,
                            double noise_eps, int32_t* d_best_edge) {
  if (!h) return CAZ_ERR_INVALID;
  if (n_nodes == 0) return CAZ_OK;
  if (!d_offsets || !d_n || !d_q || !d_p || !d_sum_n || !d_best_edge || n_nodes < 0)
    return fail(h, CAZ_ERR_INVALID, "caz_puct_select_device: bad argument");
  int rc;
  if ((rc = set_device(h))) return rc;
  const bool with_noise = d_noise && d_is_root;
  const int wpb = 8;
  caz::puct_select_kernel<<<(n_nodes + wpb - 1) / wpb, wpb * 32, 0, h->stream>>>(d_offsets, n_nodes, d_n, d_q, d_p, d_sum_n, with_noise ? d_noise : nullptr,
                                                                               with_noise ? d_is_root : nullptr, c_puct, noise_eps, d_best_edge);
  return check_launch(h, "puct_select_kernel");
}

int caz_debug_conv_layer(caz_handle* h, int32_t layer, const float* in, const float* residual, int32_t batch,
                         float* out) {
  if (!h) return CAZ_ERR_INVALID;
  if (!in || !out || batch < 1 || batch > h->max_batch || layer < 0 || layer >= (int)h->convs.size())
    return fail(h, CAZ_ERR_INVALID, "caz_debug_conv_layer: bad argument");
  int rc;
  if ((rc = set_device(h))) return rc;
  const ConvLayer& L = h->convs[layer];
  const bool tcp = tensor_path(h);
  const size_t rows = (size_t)batch * 64;
  const int F = h->arch.filters;
  float *d_in = nullptr, *d_res = nullptr, *d_out = nullptr;
  CU_TRY(h, cudaMalloc(&d_in, rows * L.cin * 4));
  CU_TRY(h, cudaMalloc(&d_out, rows * F * 4));
  CU_TRY(h, cudaMemcpyAsync(d_in, in, rows * L.cin * 4, cudaMemcpyHostToDevice, h->stream));
  if (residual) {
    CU_TRY(h, cudaMalloc(&d_res, rows * F * 4));
    CU_TRY(h, cudaMemcpyAsync(d_res, residual, rows * F * 4, cudaMemcpyHostToDevice, h->stream));
  }
  // stage into the handle's own activation buffers so the real tensor maps and kernels are exercised:
  // input -> in_act (stem) or act[1]; residual -> act[2] (fp32 in both builds); output -> act[0] (+ act[2])
  void* a_in = layer == 0 ? h->in_act : h->act[1];
  const CUtensorMap* tm = layer == 0 ? &h->tmap_in : &h->tmap_act[1];
  const int cpad = layer == 0 ? h->in_cpad : F;
  const unsigned gi = (unsigned)((rows * cpad + 255) / 256), go = (unsigned)((rows * F + 255) / 256);
  if (tcp && act_fp16(h)) caz::cast_pad_kernel<__half><<<gi, 256, 0, h->stream>>>(d_in, (__half*)a_in, rows, L.cin, cpad);
  else if (tcp) caz::cast_pad_kernel<__nv_bfloat16><<<gi, 256, 0, h->stream>>>(d_in, (__nv_bfloat16*)a_in, rows, L.cin, cpad);
  else caz::cast_pad_kernel<float><<<gi, 256, 0, h->stream>>>(d_in, (float*)a_in, rows, L.cin, cpad);
  if ((rc = check_launch(h, "cast_pad_kernel"))) return rc;
  float* xf = static_cast<float*>(h->act[2]);
  if (residual && tcp) {
    const size_t total4 = ((rows + 127) / 128) * (F / 32) * 8 * 128;
    caz::xf_to_native_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, h->stream>>>(d_res, reinterpret_cast<float4*>(xf), rows, F);
    if ((rc = check_launch(h, "xf_to_native_kernel"))) return rc;
  } else if (residual) {
    CU_TRY(h, cudaMemcpyAsync(xf, d_res, rows * F * 4, cudaMemcpyDeviceToDevice, h->stream));
  }
  if ((rc = launch_conv(h, layer, a_in, tm, cpad, residual ? xf : nullptr, h->act[0], (tcp && (residual || layer == 0)) ? xf : nullptr, batch))) return rc;
  if (tcp && act_fp16(h)) caz::widen_kernel<__half><<<go, 256, 0, h->stream>>>((const __half*)h->act[0], d_out, rows * F);
  else if (tcp) caz::widen_kernel<__nv_bfloat16><<<go, 256, 0, h->stream>>>((const __nv_bfloat16*)h->act[0], d_out, rows * F);
  else caz::widen_kernel<float><<<go, 256, 0, h->stream>>>((const float*)h->act[0], d_out, rows * F);
  if ((rc = check_launch(h, "widen_kernel"))) return rc;
  CU_TRY(h, cudaMemcpyAsync(out, d_out, rows * F * 4, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_in); cudaFree(d_out); if (d_res) cudaFree(d_res);
  return CAZ_OK;
}

int caz_debug_sb_trace(caz_handle* h, int32_t enable, int64_t* stamps, int32_t n) {
  if (!h || !h->sb_ok) return fail(h, CAZ_ERR_INVALID, "small-batch kernel not available");
  int rc;
  if ((rc = set_device(h))) return rc;
  h->sb_trace_on = enable != 0;
  if (stamps && n > 0) {
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    CU_TRY(h, cudaMemcpy(stamps, h->sb_trace, (size_t)(n > 1024 + 16 * 256 * 2 ? 1024 + 16 * 256 * 2 : n) * sizeof(long long), cudaMemcpyDeviceToHost));
  }
  return CAZ_OK;
}

int caz_debug_tower_trace(caz_handle* h, int32_t enable, int64_t* stamps, int32_t n) {
  if (!h || !h->tw_ok) return fail(h, CAZ_ERR_INVALID, "one-launch tower kernel not available");
  int rc;
  if ((rc = set_device(h))) return rc;
  h->tw_trace_on = enable != 0;
  if (stamps && n > 0) {
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    CU_TRY(h, cudaMemcpy(stamps, h->tw_trace, (size_t)(n > 512 ? 512 : n) * sizeof(long long), cudaMemcpyDeviceToHost));
  }
  return CAZ_OK;
}

int caz_set_small_batch_limit(caz_handle* h, int32_t limit) {
  if (!h) return CAZ_ERR_INVALID;
  h->sb_limit = limit;
  return CAZ_OK;
}

int caz_set_fused_tower(caz_handle* h, int32_t mode) {
  if (!h) return CAZ_ERR_INVALID;
  if (mode < -1 || mode > 2) return fail(h, CAZ_ERR_INVALID, "caz_set_fused_tower: mode must be -1, 0, 1 or 2");
  h->tw_enabled = mode != 0;
  h->tw_mode = mode > 0 ? mode : 0;
  return CAZ_OK;
}

int32_t caz_max_batch(const caz_handle* h) { return h ? h->max_batch : 0; }
int32_t caz_precision(const caz_handle* h) { return h ? h->precision : -1; }
void* caz_stream(const caz_handle* h) { return h ? static_cast<void*>(h->stream) : nullptr; }
int64_t caz_kernel_launches(const caz_handle* h) { return h ? h->launches : 0; }
int64_t caz_weight_bytes(const caz_handle* h) { return h ? h->weight_bytes : 0; }
int64_t caz_inexact_weights(const caz_handle* h) { return h ? h->inexact_weights : 0; }

int caz_profile(caz_handle* h, int32_t enable) {
  if (!h) return CAZ_ERR_INVALID;
  h->profiling = enable != 0;
  return CAZ_OK;
}

int caz_profile_read(caz_handle* h, float* ms_per_stage, int64_t* launches_per_stage, int32_t reset) {
  if (!h) return CAZ_ERR_INVALID;
  int rc;
  if ((rc = set_device(h))) return rc;
  if ((rc = prof_drain(h))) return rc;
  for (int s = 0; s < CAZ_STAGES; ++s) {
    if (ms_per_stage) ms_per_stage[s] = h->stage_ms[s];
    if (launches_per_stage) launches_per_stage[s] = h->stage_launches[s];
    if (reset) { h->stage_ms[s] = 0; h->stage_launches[s] = 0; }
  }
  return CAZ_OK;
}

void* caz_host_alloc(int64_t bytes) {
  void* p = nullptr;
  if (bytes <= 0 || cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void caz_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"

#include "caz_train.inc.cu"
