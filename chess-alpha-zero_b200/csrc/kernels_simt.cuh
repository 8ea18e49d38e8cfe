// CUDA-core kernels of the evaluator: input packing / encoding, the fp32 parity build's convolution,
// the heads (1x1 convs, tiled Dense, softmax + value output, operand split for the tensor-core Dense) and PUCT.
// Reference semantics cited per kernel; paths relative to /root/reference/src/chess_zero.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace caz {

// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while
// its predecessor in the stream is still running; this lets ITS successor do the same and then waits until the
// predecessor has completed and its writes are visible -- before the first global access of the caller.  Without the
// launch attribute both instructions do nothing.
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory");
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f)); }
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ------------------------------------------------------------------------------------------------
// planes f32 [B][Cin][64] (NCHW, what predict_on_batch receives, agent/api_chess.py:66)
//   -> activations T [B][64][Cpad] (NHWC, channels zero-padded to Cpad)
// One block per position; coalesced reads, smem transpose, coalesced writes.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void pack_planes_kernel(const float* __restrict__ planes, T* __restrict__ act, int batch, int cin,
                                   int cpad) {
  extern __shared__ float s_pl[];  // [cin][65]
  const int b = blockIdx.x;
  if (b >= batch) return;
  const float* src = planes + static_cast<size_t>(b) * cin * 64;
  for (int i = threadIdx.x; i < cin * 64; i += blockDim.x) s_pl[(i >> 6) * 65 + (i & 63)] = src[i];
  __syncthreads();
  T* dst = act + static_cast<size_t>(b) * 64 * cpad;
  if (sizeof(T) == 2 && (cpad & 7) == 0) {
    // 16-bit activations: one 16-byte store = 8 channels of one square
    const int chunks = cpad >> 3;
    for (int i = threadIdx.x; i < 64 * chunks; i += blockDim.x) {
      const int sq = i / chunks, c0 = (i - sq * chunks) * 8;
      T v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = from_f32<T>(c0 + j < cin ? s_pl[(c0 + j) * 65 + sq] : 0.0f);
      *reinterpret_cast<uint4*>(dst + static_cast<size_t>(sq) * cpad + c0) = *reinterpret_cast<const uint4*>(v);
    }
    return;
  }
  for (int i = threadIdx.x; i < 64 * cpad; i += blockDim.x) {
    const int sq = i / cpad, c = i - sq * cpad;
    dst[i] = from_f32<T>(c < cin ? s_pl[c * 65 + sq] : 0.0f);
  }
}

// rows x cin fp32 -> rows x cpad T (zero padded), and back (test hook caz_debug_conv_layer)
template <typename T>
__global__ void cast_pad_kernel(const float* __restrict__ in, T* __restrict__ out, size_t rows, int cin, int cpad) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * cpad) return;
  const size_t r = i / cpad;
  const int c = static_cast<int>(i - r * cpad);
  out[i] = from_f32<T>(c < cin ? in[r * cin + c] : 0.0f);
}
// fp32 [m][k] -> 16-bit [mpad][3k] = [hi | hi | lo] with hi = round16(x), lo = round16(x - hi): the A operand of the
// "3-term split" tensor-core GEMM  x.w ~= hi.wh + hi.wl + lo.wh  (error ~2^-17 relative with bf16, fp32 accumulate)
__global__ void split3_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, int m, int mpad, int k, int fp16) {
  pdl_sync();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(mpad) * k) return;
  const int row = static_cast<int>(i / k), col = static_cast<int>(i - static_cast<size_t>(row) * k);
  float x = row < m ? in[i] : 0.0f;
  uint16_t hi, lo;
  if (fp16) {
    const __half h = __float2half_rn(fminf(fmaxf(x, -65504.0f), 65504.0f));
    const __half l = __float2half_rn(x - __half2float(h));
    hi = *reinterpret_cast<const uint16_t*>(&h);
    lo = *reinterpret_cast<const uint16_t*>(&l);
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    hi = *reinterpret_cast<const uint16_t*>(&h);
    lo = *reinterpret_cast<const uint16_t*>(&l);
  }
  uint16_t* o = out + static_cast<size_t>(row) * 3 * k;
  o[col] = hi;
  o[k + col] = hi;
  o[2 * k + col] = lo;
}

// row-major fp32 [rows][n] -> the tensor-core epilogues' native residual layout (conv_tc.cuh: xf_index):
// [tile][n/32 chunks][8 float4][128 lanes], lane m of a tile = (rank m>>4, board (m>>3)&1, file m&7)
__global__ void xf_to_native_kernel(const float* __restrict__ in, float4* __restrict__ out, size_t rows, int n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // one float4 of the output
  const int chunks = n >> 5;
  const size_t total = ((rows + 127) / 128) * chunks * 8 * 128;
  if (i >= total) return;
  const int m = static_cast<int>(i & 127);
  const int q = static_cast<int>((i >> 7) & 7);
  const size_t t = i >> 10;
  const int chunk = static_cast<int>(t % chunks);
  const size_t tile = t / chunks;
  const size_t row = (tile * 2 + ((m >> 3) & 1)) * 64 + (m >> 4) * 8 + (m & 7);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < rows) v = *reinterpret_cast<const float4*>(in + row * n + chunk * 32 + q * 4);
  out[i] = v;
}

template <typename T>
__global__ void widen_kernel(const T* __restrict__ in, float* __restrict__ out, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = to_f32<T>(in[i]);
}

// ------------------------------------------------------------------------------------------------
// Board records (68 bytes, include/caz_b200.h) -> canonical (18,8,8) float32 planes.
// Reproduces env/chess_env.py:231-348 bit for bit: black to move mirrors ranks and swaps piece colour
// (maybe_flip_fen :251-265), castling planes follow the case-swapped field (:279-283), plane 16 is the raw
// halfmove clock (:276-277), plane 17 is the en-passant square UN-mirrored (:271-274).
// One thread per output float; every thread of a block reads the same 68-byte record (L1 broadcast).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float record_plane_value(const uint8_t* __restrict__ rec, int plane, int sq) {
  const uint8_t flags = rec[64];
  const bool black = flags & 1;
  if (plane < 12) {
    const int r = sq >> 3, f = sq & 7;
    const int code = rec[black ? ((7 - r) << 3 | f) : sq];  // mirrored source square for black
    if (code == 0) return 0.0f;
    int p = code - 1;
    if (black) p = p >= 6 ? p - 6 : p + 6;  // swap case
    return p == plane ? 1.0f : 0.0f;
  }
  if (plane < 16) {
    int bit = plane - 12;          // K Q k q as seen by the side to move
    if (black) bit ^= 2;           // swapcase exchanges KQ <-> kq
    return (flags >> (1 + bit)) & 1 ? 1.0f : 0.0f;
  }
  if (plane == 16) return static_cast<float>(static_cast<int>(rec[66]) | (static_cast<int>(rec[67]) << 8));
  return rec[65] == sq ? 1.0f : 0.0f;  // 255 never matches
}

__global__ void encode_records_kernel(const uint8_t* __restrict__ records, float* __restrict__ planes, int batch) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(batch) * 1152;
  if (i >= total) return;
  const int b = static_cast<int>(i / 1152);
  const int rem = static_cast<int>(i - static_cast<size_t>(b) * 1152);
  planes[i] = record_plane_value(records + static_cast<size_t>(b) * 68, rem >> 6, rem & 63);
}

// records -> NHWC activations directly (no fp32 planes round trip through HBM)
template <typename T>
__global__ void encode_pack_kernel(const uint8_t* __restrict__ records, T* __restrict__ act, int batch, int cpad) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(batch) * 64 * cpad;
  if (i >= total) return;
  const int c = static_cast<int>(i % cpad);
  const size_t t = i / cpad;
  const int sq = static_cast<int>(t & 63);
  const size_t b = t >> 6;
  act[i] = from_f32<T>(c < 18 ? record_plane_value(records + b * 68, c, sq) : 0.0f);
}

// ------------------------------------------------------------------------------------------------
// fp32 parity build: direct "same" convolution on NHWC fp32, BN folded, optional residual + ReLU
// (agent/model_chess.py:65-69,96-111).  One block per position; the zero-padded input board lives in
// shared memory; each thread owns one output channel and all 64 squares in registers.
// w layout: [tap][ic][oc] (= Keras HWIO flattened, scaled by the folded BN factor).
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256)
conv_fp32_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                 const float* __restrict__ residual, float* __restrict__ out, int cin, int cin_stride, int cout,
                 int relu) {
  constexpr int P = (K - 1) / 2;
  constexpr int HW = 8 + K - 1;
  extern __shared__ float s_in[];  // [HW*HW][cin]
  const int b = blockIdx.x;
  const float* src = in + static_cast<size_t>(b) * 64 * cin_stride;
  for (int i = threadIdx.x; i < HW * HW * cin; i += blockDim.x) {
    const int cell = i / cin, c = i - cell * cin;
    const int r = cell / HW - P, f = cell % HW - P;
    s_in[i] = (r >= 0 && r < 8 && f >= 0 && f < 8) ? src[(r * 8 + f) * cin_stride + c] : 0.0f;
  }
  __syncthreads();
  for (int oc = threadIdx.x; oc < cout; oc += blockDim.x) {
    float acc[64];
#pragma unroll
    for (int s = 0; s < 64; ++s) acc[s] = 0.0f;
    for (int tap = 0; tap < K * K; ++tap) {
      const int dy = tap / K, dx = tap % K;
      const float* wp = w + static_cast<size_t>(tap) * cin * cout + oc;
      const float* sp = s_in + (dy * HW + dx) * cin;
      for (int ic = 0; ic < cin; ++ic) {
        const float wv = __ldg(wp + static_cast<size_t>(ic) * cout);
#pragma unroll
        for (int s = 0; s < 64; ++s) acc[s] = fmaf(wv, sp[((s >> 3) * HW + (s & 7)) * cin + ic], acc[s]);
      }
    }
    const float bv = bias[oc];
#pragma unroll
    for (int s = 0; s < 64; ++s) {
      const size_t o = (static_cast<size_t>(b) * 64 + s) * cout + oc;
      float v = acc[s] + bv;
      if (residual) v += residual[o];
      if (relu) v = fmaxf(v, 0.0f);
      out[o] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Head 1x1 convolutions + folded BN + ReLU + Flatten (agent/model_chess.py:77-81,86-90).
// in: T [B][64][C]; w: [F][C] fp32 (policy filters first, then value filters), bias [F].
// Output in Keras channels_first Flatten order: featp[b][c*64 + s], featv[b][c*64 + s].
// One block (8 warps) per position; a warp takes 8 squares, lanes split the channels.
// ------------------------------------------------------------------------------------------------
constexpr int HEAD_MAX_F = 8;

template <typename T>
__global__ void __launch_bounds__(256)
head_conv_kernel(const T* __restrict__ act, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ featp, float* __restrict__ featv, int batch, int c, int pf, int vf) {
  extern __shared__ float s_w[];  // [F][c]
  const int f_total = pf + vf;
  for (int i = threadIdx.x; i < f_total * c; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int sq = warp; sq < 64; sq += 8) {
    const T* row = act + (static_cast<size_t>(b) * 64 + sq) * c;
    float acc[HEAD_MAX_F];
#pragma unroll
    for (int f = 0; f < HEAD_MAX_F; ++f) acc[f] = 0.0f;
    for (int ch = lane; ch < c; ch += 32) {
      const float x = to_f32<T>(row[ch]);
#pragma unroll
      for (int f = 0; f < HEAD_MAX_F; ++f)
        if (f < f_total) acc[f] = fmaf(x, s_w[f * c + ch], acc[f]);
    }
#pragma unroll
    for (int f = 0; f < HEAD_MAX_F; ++f) {
      if (f < f_total) {
        float v = acc[f];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
          v = fmaxf(v + bias[f], 0.0f);
          if (f < pf) featp[static_cast<size_t>(b) * pf * 64 + f * 64 + sq] = v;
          else featv[static_cast<size_t>(b) * vf * 64 + (f - pf) * 64 + sq] = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Policy head: Dense(n_labels) + softmax (agent/model_chess.py:82-83), optionally restricted to a legal
// mask (how player_chess.py:267-275 consumes the row).  fp32 throughout, two kernels:
//   policy_logits_kernel : logits = featp . wk + bk as a shared-memory tiled GEMM, 64 positions x 128 labels
//                          per block (the 1 MB weight matrix is read once per 64 positions, not once per 8)
//   policy_softmax_kernel: in-place row softmax, one block per position, the row lives in registers
// wk: [kin][n_labels] (Keras Dense kernel layout).
// ------------------------------------------------------------------------------------------------
constexpr int PG_TM = 128;   // rows (positions) per block
constexpr int PG_TN = 128;   // columns (labels / hidden units) per block
constexpr int PG_TK = 16;    // K slab
constexpr int PNJ = 8;       // labels per thread in the softmax kernel (8 * 256 >= 1968)

// C[M][N] = act(A[M][K] . W[K][N] + bias[N]), fp32 on the CUDA cores.  128 x 128 block tile, 8 x 8 register tile per
// thread, K slabs of 16 double-buffered through registers (the next slab's global loads are in flight while the
// current one is multiplied).  Used for the policy logits (N = 1968, K = 128) and the value head's hidden layer
// (N = 256, K = 256, ReLU): agent/model_chess.py:82-83,91.
__global__ void __launch_bounds__(256, 2)
dense_tiled_kernel(const float* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ c, int m, int kdim, int n, int relu) {
  __shared__ __align__(16) float sA[PG_TK][PG_TM + 4];  // [k][row]
  __shared__ __align__(16) float sW[PG_TK][PG_TN];      // [k][col]
  const int m0 = blockIdx.x * PG_TM, n0 = blockIdx.y * PG_TN;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;  // rows ty*8..+7, cols tx*4..+3 and 64+tx*4..+3
  // global -> register staging: A slab is 128 rows x 16 k = 2048 floats (8 per thread), W slab 16 k x 128 cols (8 per thread)
  const int a_row = threadIdx.x >> 1, a_k0 = (threadIdx.x & 1) * 8;   // 8 consecutive k of one row
  const int w_k = threadIdx.x >> 4, w_c0 = (threadIdx.x & 15) * 8;    // 8 consecutive cols of one k
  float ra[8], rw[8];
  auto load_slab = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kk = k0 + a_k0 + j;
      ra[j] = (m0 + a_row < m && kk < kdim) ? a[static_cast<size_t>(m0 + a_row) * kdim + kk] : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = n0 + w_c0 + j;
      rw[j] = (k0 + w_k < kdim && col < n) ? __ldg(w + static_cast<size_t>(k0 + w_k) * n + col) : 0.0f;
    }
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
  load_slab(0);
  for (int k0 = 0; k0 < kdim; k0 += PG_TK) {
    __syncthreads();  // previous slab fully consumed
#pragma unroll
    for (int j = 0; j < 8; ++j) sA[a_k0 + j][a_row] = ra[j];
#pragma unroll
    for (int j = 0; j < 8; ++j) sW[w_k][w_c0 + j] = rw[j];
    __syncthreads();
    if (k0 + PG_TK < kdim) load_slab(k0 + PG_TK);
#pragma unroll
    for (int k = 0; k < PG_TK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sA[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sA[k][ty * 8 + 4]);
      const float4 w0 = *reinterpret_cast<const float4*>(&sW[k][tx * 4]);
      const float4 w1 = *reinterpret_cast<const float4*>(&sW[k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + ty * 8 + i;
    if (row >= m) continue;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int col = n0 + half * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = col + j < n ? acc[i][half * 4 + j] + bias[col + j] : 0.0f;
        if (relu) v[j] = fmaxf(v[j], 0.0f);
      }
      float* dst = c + static_cast<size_t>(row) * n + col;
      if (col + 3 < n && (n & 3) == 0) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      else
        for (int j = 0; j < 4; ++j)
          if (col + j < n) dst[j] = v[j];
    }
  }
}

// In-place row softmax of the policy logits (optionally restricted to a legal mask), one block per position, the row
// in registers; the same block finishes the value head: value = tanh(vhid . w2 + b2) (agent/model_chess.py:92).
__global__ void __launch_bounds__(256)
policy_softmax_kernel(float* __restrict__ policy, const uint8_t* __restrict__ mask, int n_labels,
                      const float* __restrict__ vhid, const float* __restrict__ w2, const float* __restrict__ b2,
                      float* __restrict__ value, int hid) {
  pdl_sync();
  __shared__ float s_red[8];
  __shared__ float s_bc;
  float* row = policy + static_cast<size_t>(blockIdx.x) * n_labels;
  const uint8_t* mrow = mask ? mask + static_cast<size_t>(blockIdx.x) * n_labels : nullptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float x[PNJ];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < PNJ; ++j) {
    const int n = threadIdx.x + j * 256;
    const bool on = n < n_labels && (!mrow || mrow[n] != 0);
    x[j] = on ? row[n] : -INFINITY;
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
  __syncthreads();  // s_red reuse
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
    const int n = threadIdx.x + j * 256;
    if (n < n_labels) row[n] = x[j] * inv;
  }
  if (value) {
    float part = 0.0f;
    for (int hh = threadIdx.x; hh < hid; hh += 256) part = fmaf(vhid[static_cast<size_t>(blockIdx.x) * hid + hh], w2[hh], part);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = b2[0];
      for (int wi = 0; wi < 8; ++wi) v += s_red[wi];
      value[blockIdx.x] = tanhf(v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// PUCT, one warp per tree node (agent/player_chess.py:251-299).  IEEE binary64, one rounding per
// operation in the reference's evaluation order (Python floats): explicit _rn intrinsics so that the
// compiler cannot contract a*b+c into an FMA.
// ------------------------------------------------------------------------------------------------
// prior push (:267-275): tot = 1e-8 + sum (left to right) of P[idx(a)]; p_a = P[idx(a)] / tot
__global__ void push_priors_kernel(const float* __restrict__ policy, int n_labels, const int* __restrict__ offsets,
                                   int n_nodes, const int* __restrict__ label_index, double* __restrict__ priors) {
  const int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (node >= n_nodes) return;
  const int beg = offsets[node], end = offsets[node + 1];
  const float* row = policy + static_cast<size_t>(node) * n_labels;
  // the sum must be sequential to match the reference bit for bit; every lane computes it
  // redundantly from L1-resident values (<= 218 edges) instead of shuffling a serial chain.
  // (a move without a label -- index < 0 -- counts as probability 0, as in the host-side search)
  double tot = 1e-8;
  for (int e = beg; e < end; ++e) {
    const int idx = label_index[e];
    tot = __dadd_rn(tot, idx >= 0 ? static_cast<double>(row[idx]) : 0.0);
  }
  for (int e = beg + lane; e < end; e += 32) {
    const int idx = label_index[e];
    priors[e] = __ddiv_rn(idx >= 0 ? static_cast<double>(row[idx]) : 0.0, tot);
  }
}

// select (:277-299): b = q + c_puct * p_ * sqrt(sum_n + 1) / (1 + n); first strict maximum from -999
__global__ void puct_select_kernel(const int* __restrict__ offsets, int n_nodes, const int* __restrict__ n,
                                   const double* __restrict__ q, const double* __restrict__ p,
                                   const int* __restrict__ sum_n, const double* __restrict__ noise,
                                   const uint8_t* __restrict__ is_root, double c_puct, double eps,
                                   int* __restrict__ best_edge) {
  const int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (node >= n_nodes) return;
  const int beg = offsets[node], end = offsets[node + 1];
  const double xx = __dsqrt_rn(static_cast<double>(sum_n[node] + 1));
  const bool root = noise != nullptr && is_root != nullptr && is_root[node] != 0;
  const double one_minus_eps = __dsub_rn(1.0, eps);
  double best_s = -999.0;
  int best_i = 0x7fffffff;
  for (int e = beg + lane; e < end; e += 32) {
    double p_ = p[e];
    if (root) p_ = __dadd_rn(__dmul_rn(one_minus_eps, p_), __dmul_rn(eps, noise[e]));
    const double u = __ddiv_rn(__dmul_rn(__dmul_rn(c_puct, p_), xx), static_cast<double>(1 + n[e]));
    const double s = __dadd_rn(q[e], u);
    if (s > best_s) {  // ascending e within a lane: strict '>' keeps the first maximum
      best_s = s;
      best_i = e - beg;
    }
  }
  // warp argmax, ties broken towards the smaller edge index (= the reference's first maximum)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double os = __shfl_xor_sync(0xffffffffu, best_s, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (os > best_s || (os == best_s && oi < best_i)) {
      best_s = os;
      best_i = oi;
    }
  }
  if (lane == 0) best_edge[node] = best_i == 0x7fffffff ? -1 : best_i;
}

}  // namespace caz
