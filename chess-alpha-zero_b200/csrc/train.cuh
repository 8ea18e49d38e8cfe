// CUDA-core kernels of the training step: what Keras runs for
//   model.compile(Adam(), ['categorical_crossentropy', 'mean_squared_error'], loss_weights=[1.25, 1.0]); model.fit(...)
// (reference src/chess_zero/worker/optimize.py:79-103) on the graph of agent/model_chess.py:57-111 -- BatchNormalization in
// TRAINING mode (batch statistics, moving averages with momentum 0.99), l2(1e-4) on every Conv2D / Dense kernel.
//
// Everything here is fp32 on row-major NHWC tensors [rows = batch * 64][channels]; the convolutions themselves are
// conv_fp32_kernel (fp32 build) or the tcgen05 implicit GEMM (16-bit builds), called from caz_train.inc.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels_simt.cuh"

namespace caz {
namespace train {

constexpr int STAT_SPLITS = 64;    // row splits of the per-channel reductions (partials are summed in a fixed order)
constexpr int STAT_THREADS = 512;
constexpr int STAT_CH = 64;        // channels per block: 16 lanes x float4
constexpr int STAT_ROW_LANES = STAT_THREADS / 16;

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- per-channel sums over the rows: partial[split][c] = sum over the split's rows of f(row, c) -------------------------
// mode 0: (sum z, sum z^2)                                -> BatchNormalization batch statistics
// mode 1: dy = da * (a > 0);  (sum dy, sum dy * xhat)     -> BatchNormalization backward, xhat = (z - mean) * invstd
// out: partial[2][STAT_SPLITS][c] (double); grid ((c + 63) / 64, STAT_SPLITS), STAT_THREADS threads
__global__ void __launch_bounds__(STAT_THREADS)
channel_sums_kernel(int mode, const float* __restrict__ z, const float* __restrict__ a, const float* __restrict__ da,
                    const float* __restrict__ da2, const float* __restrict__ mean, const float* __restrict__ invstd, int rows, int c,
                    double* __restrict__ partial, const __half* __restrict__ a16 = nullptr) {
  // a16 (tensor-core build, c % 4 == 0): the ReLU mask is read from the fp16 copy of the activation (half the bytes of `a`)
  // a block covers 64 channels (16 lanes x float4) x 32 row lanes; c < 4 or not a multiple of 4: scalar path below
  const int split = blockIdx.y;
  const int per = (rows + STAT_SPLITS - 1) / STAT_SPLITS;
  const int r0 = split * per, r1 = min(rows, r0 + per);
  const int sub = threadIdx.x >> 4, lane = threadIdx.x & 15;
  __shared__ double sh[2][STAT_ROW_LANES][STAT_CH];
  double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
  if ((c & 3) == 0) {
    const int ch = blockIdx.x * STAT_CH + lane * 4;
    if (ch < c) {
      float4 mu = make_float4(0, 0, 0, 0), is = mu;
      if (mode) { mu = *reinterpret_cast<const float4*>(mean + ch); is = *reinterpret_cast<const float4*>(invstd + ch); }
#pragma unroll 4
      for (int r = r0 + sub; r < r1; r += STAT_ROW_LANES) {
        const size_t i = static_cast<size_t>(r) * c + ch;
        const float4 zv = *reinterpret_cast<const float4*>(z + i);
        if (mode == 0) {
          s0[0] += zv.x; s0[1] += zv.y; s0[2] += zv.z; s0[3] += zv.w;
          s1[0] += static_cast<double>(zv.x) * zv.x; s1[1] += static_cast<double>(zv.y) * zv.y;
          s1[2] += static_cast<double>(zv.z) * zv.z; s1[3] += static_cast<double>(zv.w) * zv.w;
        } else {
          float4 av;
          if (a16) {
            const uint2 raw = *reinterpret_cast<const uint2*>(a16 + i);
            const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
            av = make_float4(lo.x, lo.y, hi.x, hi.y);
          } else {
            av = *reinterpret_cast<const float4*>(a + i);
          }
          float4 dv = *reinterpret_cast<const float4*>(da + i);
          if (da2) {   // the gradient that arrived over the skip connection
            const float4 d2 = *reinterpret_cast<const float4*>(da2 + i);
            dv.x += d2.x; dv.y += d2.y; dv.z += d2.z; dv.w += d2.w;
          }
          const float dy[4] = {av.x > 0.0f ? dv.x : 0.0f, av.y > 0.0f ? dv.y : 0.0f, av.z > 0.0f ? dv.z : 0.0f, av.w > 0.0f ? dv.w : 0.0f};
          const float xh[4] = {(zv.x - mu.x) * is.x, (zv.y - mu.y) * is.y, (zv.z - mu.z) * is.z, (zv.w - mu.w) * is.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) { s0[q] += dy[q]; s1[q] += static_cast<double>(dy[q]) * xh[q]; }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) { sh[0][sub][lane * 4 + q] = s0[q]; sh[1][sub][lane * 4 + q] = s1[q]; }
  } else {
    // scalar path (the heads' 1x1 convolutions have 1, 2 or 4 channels... and anything not a multiple of 4): channel lane + 16 q
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ch = blockIdx.x * STAT_CH + lane + 16 * q;
      if (ch < c) {
        const float mu = mode ? mean[ch] : 0.0f, is = mode ? invstd[ch] : 0.0f;
        for (int r = r0 + sub; r < r1; r += STAT_ROW_LANES) {
          const size_t i = static_cast<size_t>(r) * c + ch;
          if (mode == 0) {
            const float v = z[i];
            s0[q] += v;
            s1[q] += static_cast<double>(v) * v;
          } else {
            const float dy = a[i] > 0.0f ? da[i] + (da2 ? da2[i] : 0.0f) : 0.0f;
            s0[q] += dy;
            s1[q] += static_cast<double>(dy) * ((z[i] - mu) * is);
          }
        }
      }
      sh[0][sub][lane + 16 * q] = s0[q];
      sh[1][sub][lane + 16 * q] = s1[q];
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * STAT_CH) {
    const int which = threadIdx.x / STAT_CH, cl = threadIdx.x % STAT_CH, cc = blockIdx.x * STAT_CH + cl;
    if (cc < c) {
      double t = 0.0;
      for (int k = 0; k < STAT_ROW_LANES; ++k) t += sh[which][k][cl];
      partial[(which * STAT_SPLITS + split) * c + cc] = t;
    }
  }
}

// sum over the STAT_SPLITS partials of channel `ch` by one warp (lane-strided, then a shuffle tree): s0 / s1 valid in every lane
__device__ __forceinline__ void warp_sum_partials(const double* __restrict__ partial, int c, int ch, int lane, double& s0, double& s1) {
  s0 = 0.0;
  s1 = 0.0;
  for (int k = lane; k < STAT_SPLITS; k += 32) { s0 += partial[(0 * STAT_SPLITS + k) * c + ch]; s1 += partial[(1 * STAT_SPLITS + k) * c + ch]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
}

// Batch statistics from the partial sums (Keras BatchNormalization in training mode: biased variance, epsilon inside the
// square root).  The batch mean / variance are also left in the GRADIENT vector's slots of the moving statistics: the
// moving-average update  moving = moving * momentum + batch * (1 - momentum)  (keras 2.0.8 normalization.py:
// K.moving_average_update with the biased batch variance) happens in adam_kernel, after the data-parallel all-reduce has
// averaged the ranks' batch statistics, so that every rank keeps identical moving statistics.
__global__ void bn_finish_stats_kernel(const double* __restrict__ partial, int rows, int c, float eps,
                                       float* __restrict__ mean, float* __restrict__ invstd,
                                       float* __restrict__ batch_mean_out, float* __restrict__ batch_var_out) {
  const int ch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;   // one warp per channel
  if (ch >= c) return;
  double s0, s1;
  warp_sum_partials(partial, c, ch, lane, s0, s1);
  if (lane != 0) return;
  const double mu = s0 / rows;
  double var = s1 / rows - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[ch] = static_cast<float>(mu);
  invstd[ch] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  batch_mean_out[ch] = static_cast<float>(mu);
  batch_var_out[ch] = static_cast<float>(var);
}

// The same from the partial sums the tensor-core convolution's epilogue left (ConvArgs::stats: [2][c][slots] floats, one slot
// per epilogue warp and CTA tile): one warp per channel, double accumulation, fixed order.
__global__ void bn_finish_stats_f32_kernel(const float* __restrict__ partial, int slots, int rows, int c, float eps,
                                           float* __restrict__ mean, float* __restrict__ invstd,
                                           float* __restrict__ batch_mean_out, float* __restrict__ batch_var_out) {
  const int ch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (ch >= c) return;
  const float* p0 = partial + static_cast<size_t>(ch) * slots;
  const float* p1 = partial + (static_cast<size_t>(c) + ch) * slots;
  double s0 = 0.0, s1 = 0.0;
  for (int k = lane; k < slots; k += 32) { s0 += p0[k]; s1 += p1[k]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
  if (lane != 0) return;
  const double mu = s0 / rows;
  double var = s1 / rows - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[ch] = static_cast<float>(mu);
  invstd[ch] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  batch_mean_out[ch] = static_cast<float>(mu);
  batch_var_out[ch] = static_cast<float>(var);
}

// a = relu(gamma * (z - mean) * invstd + beta (+ skip));  optional 16-bit NHWC copy for the tensor-core convolutions
template <typename H>
__global__ void bn_apply_kernel(const float* __restrict__ z, const float* __restrict__ skip, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ mean,
                                const float* __restrict__ invstd, size_t total, int c, float* __restrict__ a,
                                H* __restrict__ a16, int vec4) {
  if (vec4) {   // four channels per thread (the host checked c % 4 == 0 and the 16-byte alignment of the parameter slices)
    const size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    if (i >= total) return;
    const int ch = static_cast<int>(i % c);
    const float4 zv = *reinterpret_cast<const float4*>(z + i), g = *reinterpret_cast<const float4*>(gamma + ch),
                 b = *reinterpret_cast<const float4*>(beta + ch), mu = *reinterpret_cast<const float4*>(mean + ch),
                 is = *reinterpret_cast<const float4*>(invstd + ch);
    float4 v = make_float4(g.x * ((zv.x - mu.x) * is.x) + b.x, g.y * ((zv.y - mu.y) * is.y) + b.y, g.z * ((zv.z - mu.z) * is.z) + b.z,
                           g.w * ((zv.w - mu.w) * is.w) + b.w);
    if (skip) {
      const float4 sk = *reinterpret_cast<const float4*>(skip + i);
      v.x += sk.x; v.y += sk.y; v.z += sk.z; v.w += sk.w;
    }
    v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f);
    *reinterpret_cast<float4*>(a + i) = v;
    if (a16) {
      H h[4] = {from_f32<H>(v.x), from_f32<H>(v.y), from_f32<H>(v.z), from_f32<H>(v.w)};
      *reinterpret_cast<uint2*>(a16 + i) = *reinterpret_cast<const uint2*>(h);
    }
    return;
  }
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ch = static_cast<int>(i % c);
  float v = gamma[ch] * ((z[i] - mean[ch]) * invstd[ch]) + beta[ch];
  if (skip) v += skip[i];
  v = fmaxf(v, 0.0f);
  a[i] = v;
  if (a16) a16[i] = from_f32<H>(v);
}

// BatchNormalization backward (training mode) through the ReLU in front of it:
//   dy = da * (a > 0);  dz = gamma * invstd * (dy - sum_dy / M - xhat * sum_dyx / M);  dgamma = sum_dyx;  dbeta = sum_dy
// dskip (conv2 of a residual block): the Add passes dy on to the block's input.  `dskip_accumulate`: add instead of store.
template <typename H>
__global__ void bn_backward_kernel(const float* __restrict__ z, const float* __restrict__ a, const float* __restrict__ da,
                                   const float* da2, const float* __restrict__ gamma, const float* __restrict__ mean,
                                   const float* __restrict__ invstd, const float* __restrict__ sum_dy,
                                   const float* __restrict__ sum_dyx, int rows, int c,
                                   float* __restrict__ dz, H* __restrict__ dz16, float* dskip, int dskip_accumulate, int vec4) {
  // sum_dy / sum_dyx: the channel sums, = dbeta / dgamma as bn_param_grads_kernel left them in the gradient vector
  // (da2 and dskip may be the same buffer: each thread reads its element of da2 before it writes that element of dskip)
  const size_t total = static_cast<size_t>(rows) * c;
  if (vec4) {   // four channels per thread
    const size_t i4 = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    if (i4 >= total) return;
    const int ch4 = static_cast<int>(i4 % c);
    const float4 zv = *reinterpret_cast<const float4*>(z + i4), av = *reinterpret_cast<const float4*>(a + i4);
    float4 dv = *reinterpret_cast<const float4*>(da + i4);
    if (da2) {
      const float4 d2 = *reinterpret_cast<const float4*>(da2 + i4);
      dv.x += d2.x; dv.y += d2.y; dv.z += d2.z; dv.w += d2.w;
    }
    const float4 g = *reinterpret_cast<const float4*>(gamma + ch4), mu = *reinterpret_cast<const float4*>(mean + ch4),
                 is = *reinterpret_cast<const float4*>(invstd + ch4), s0 = *reinterpret_cast<const float4*>(sum_dy + ch4),
                 s1 = *reinterpret_cast<const float4*>(sum_dyx + ch4);
    const float4 dy = make_float4(av.x > 0.0f ? dv.x : 0.0f, av.y > 0.0f ? dv.y : 0.0f, av.z > 0.0f ? dv.z : 0.0f, av.w > 0.0f ? dv.w : 0.0f);
    const float4 o = make_float4(g.x * is.x * (dy.x - s0.x / rows - (zv.x - mu.x) * is.x * (s1.x / rows)),
                                 g.y * is.y * (dy.y - s0.y / rows - (zv.y - mu.y) * is.y * (s1.y / rows)),
                                 g.z * is.z * (dy.z - s0.z / rows - (zv.z - mu.z) * is.z * (s1.z / rows)),
                                 g.w * is.w * (dy.w - s0.w / rows - (zv.w - mu.w) * is.w * (s1.w / rows)));
    *reinterpret_cast<float4*>(dz + i4) = o;
    if (dz16) {
      H h[4] = {from_f32<H>(o.x), from_f32<H>(o.y), from_f32<H>(o.z), from_f32<H>(o.w)};
      *reinterpret_cast<uint2*>(dz16 + i4) = *reinterpret_cast<const uint2*>(h);
    }
    if (dskip) {
      float4 sk = dy;
      if (dskip_accumulate) { const float4 old = *reinterpret_cast<const float4*>(dskip + i4); sk.x += old.x; sk.y += old.y; sk.z += old.z; sk.w += old.w; }
      *reinterpret_cast<float4*>(dskip + i4) = sk;
    }
    return;
  }
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ch = static_cast<int>(i % c);
  const float s0 = sum_dy[ch], s1 = sum_dyx[ch];
  const float dy = a[i] > 0.0f ? da[i] + (da2 ? da2[i] : 0.0f) : 0.0f;
  const float xhat = (z[i] - mean[ch]) * invstd[ch];
  const float v = gamma[ch] * invstd[ch] * (dy - s0 / rows - xhat * (s1 / rows));
  dz[i] = v;
  if (dz16) dz16[i] = from_f32<H>(v);
  if (dskip) dskip[i] = dskip_accumulate ? dskip[i] + dy : dy;
}

// ---- tensor-core build: the two BatchNormalization passes over a TILE of 64 boards x 64 channels of one square ----------
// Same arithmetic as bn_apply_kernel / bn_backward_kernel (bit-identical results when the ReLU mask is read from the fp32
// activation, CAZ_TRAIN_MASK_FP32; by default it comes from the fp16 copy -- half the bytes -- which differs where
// 0 < a < 2^-25: about one element per few layers, far fewer than the forward pass's 16-bit operands flip), but a block owns 64 boards of one square, so that besides the
// row-major outputs it can write, through shared memory, the K-major bf16 operand of the weight-gradient GEMM
//   outT[ch][cell][board],  cell = (rank + pad) * pitch + file + pad over the zero-padded board (see transpose_pad_kernel)
// in 128-byte runs -- the separate transposes (a read of the fp32 tensor and a launch each) are gone, and the backward
// pass no longer writes the fp32 dz at all.  c % 64 == 0; grid (nb / 64, c / 64, 64 squares), 256 threads; boards >= batch
// are written as zeros to outT.
__device__ __forceinline__ void tile_store_transposed(const float (*tile)[65], __nv_bfloat16* __restrict__ outT, int c0, int b0, int nb,
                                                      int pad, int sq, int tx, int ty) {
  const int pitch = 8 + 2 * pad, cells = pitch * pitch, cell = ((sq >> 3) + pad) * pitch + (sq & 7) + pad;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int chl = ty + r * 8, b = b0 + 2 * tx;
    if (b < nb)
      *reinterpret_cast<__nv_bfloat162*>(outT + (static_cast<size_t>(c0 + chl) * cells + cell) * nb + b) =
          __floats2bfloat162_rn(tile[2 * tx][chl], tile[2 * tx + 1][chl]);
  }
}

__global__ void __launch_bounds__(256)
bn_apply_tile_kernel(const float* __restrict__ z, const float* __restrict__ skip, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ invstd, int batch, int c,
                     int nb, int pad, float* __restrict__ a, __half* __restrict__ a16, __nv_bfloat16* __restrict__ aT) {
  __shared__ float tile[64][65];
  const int b0 = blockIdx.x * 64, c0 = blockIdx.y * 64, sq = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, ch = c0 + 2 * tx;
  const float2 g = *reinterpret_cast<const float2*>(gamma + ch), be = *reinterpret_cast<const float2*>(beta + ch),
               mu = *reinterpret_cast<const float2*>(mean + ch), is = *reinterpret_cast<const float2*>(invstd + ch);
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int bl = ty + r * 8, b = b0 + bl;
    float2 v = make_float2(0.0f, 0.0f);
    if (b < batch) {
      const size_t i = (static_cast<size_t>(b) * 64 + sq) * c + ch;
      const float2 zv = *reinterpret_cast<const float2*>(z + i);
      v = make_float2(g.x * ((zv.x - mu.x) * is.x) + be.x, g.y * ((zv.y - mu.y) * is.y) + be.y);
      if (skip) {
        const float2 sk = *reinterpret_cast<const float2*>(skip + i);
        v.x += sk.x; v.y += sk.y;
      }
      v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f);
      *reinterpret_cast<float2*>(a + i) = v;
      *reinterpret_cast<__half2*>(a16 + i) = __floats2half2_rn(v.x, v.y);
    }
    tile[bl][2 * tx] = v.x;
    tile[bl][2 * tx + 1] = v.y;
  }
  if (!aT) return;
  __syncthreads();
  tile_store_transposed(tile, aT, c0, b0, nb, pad, sq, tx, ty);
}

__global__ void __launch_bounds__(256)
bn_backward_tile_kernel(const float* __restrict__ z, const float* __restrict__ a, const __half* __restrict__ a16, const float* __restrict__ da, const float* da2,
                        const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ invstd,
                        const float* __restrict__ sum_dy, const float* __restrict__ sum_dyx, int batch, int c, int nb, int pad,
                        __nv_bfloat16* __restrict__ dz16, __nv_bfloat16* __restrict__ dzT, float* dskip) {
  // (da2 and dskip may be the same buffer: each thread reads its elements of da2 before it writes them to dskip)
  __shared__ float tile[64][65];
  const int b0 = blockIdx.x * 64, c0 = blockIdx.y * 64, sq = blockIdx.z, rows = batch * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, ch = c0 + 2 * tx;
  const float2 g = *reinterpret_cast<const float2*>(gamma + ch), mu = *reinterpret_cast<const float2*>(mean + ch),
               is = *reinterpret_cast<const float2*>(invstd + ch), s0 = *reinterpret_cast<const float2*>(sum_dy + ch),
               s1 = *reinterpret_cast<const float2*>(sum_dyx + ch);
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int bl = ty + r * 8, b = b0 + bl;
    float2 o = make_float2(0.0f, 0.0f);
    if (b < batch) {
      const size_t i = (static_cast<size_t>(b) * 64 + sq) * c + ch;
      const float2 zv = *reinterpret_cast<const float2*>(z + i),
                   av = a16 ? __half22float2(*reinterpret_cast<const __half2*>(a16 + i)) : *reinterpret_cast<const float2*>(a + i);
      float2 dv = *reinterpret_cast<const float2*>(da + i);
      if (da2) {
        const float2 d2 = *reinterpret_cast<const float2*>(da2 + i);
        dv.x += d2.x; dv.y += d2.y;
      }
      const float2 dy = make_float2(av.x > 0.0f ? dv.x : 0.0f, av.y > 0.0f ? dv.y : 0.0f);
      o = make_float2(g.x * is.x * (dy.x - s0.x / rows - (zv.x - mu.x) * is.x * (s1.x / rows)),
                      g.y * is.y * (dy.y - s0.y / rows - (zv.y - mu.y) * is.y * (s1.y / rows)));
      *reinterpret_cast<__nv_bfloat162*>(dz16 + i) = __floats2bfloat162_rn(o.x, o.y);
      if (dskip) *reinterpret_cast<float2*>(dskip + i) = dy;
    }
    tile[bl][2 * tx] = o.x;
    tile[bl][2 * tx + 1] = o.y;
  }
  __syncthreads();
  tile_store_transposed(tile, dzT, c0, b0, nb, pad, sq, tx, ty);
}

// dgamma / dbeta from the partial sums of channel_sums_kernel mode 1
__global__ void bn_param_grads_kernel(const double* __restrict__ partial, int c, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta) {
  const int ch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;   // one warp per channel
  if (ch >= c) return;
  double s0, s1;
  warp_sum_partials(partial, c, ch, lane, s0, s1);
  if (lane != 0) return;
  dbeta[ch] = static_cast<float>(s0);
  dgamma[ch] = static_cast<float>(s1);
}

// ---- a small general GEMM on the CUDA cores (heads' 1x1 convolutions and Dense layers, forward and backward) ----------
//   C[i][j] (+)= sum_k A(i, k) * B(k, j) (+ bias[j]),  A(i, k) = a[i * a_si + k * a_sk],  B(k, j) = b[k * b_sk + j * b_sj]
// 32 x 32 output tile per block, K in slabs of 32; gridDim.z splits K: slice z writes its partial sums to c + z * m * n and
// sum_partials_kernel adds the slices in a fixed order (no atomicAdd on the data-gradient path: run-to-run rounding noise there
// is amplified by every 16-bit rounding of the tensor-core build's backward pass).
__global__ void __launch_bounds__(256)
gemm_kernel(const float* __restrict__ a, long a_si, long a_sk, const float* __restrict__ b, long b_sk, long b_sj,
            const float* __restrict__ bias, float* __restrict__ c, int m, int n, int k, int accumulate, int relu) {
  __shared__ float sa[32][33], sb[32][33];   // sa[i][k], sb[k][j]
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // ty 0..7: rows ty, ty+8, ty+16, ty+24
  const int kz = gridDim.z, kper = ((k + kz - 1) / kz + 31) / 32 * 32;
  const int k0 = blockIdx.z * kper, k1 = min(k, k0 + kper);
  // the lanes of a warp run along whichever index of the operand is contiguous in memory (A may be a transposed view, B too)
  const bool a_lanes_k = a_sk == 1 || a_si != 1, b_lanes_j = b_sj == 1 || b_sk != 1;
  float ra[4], rb[4];
  auto fetch = [&](int kk) {   // this thread's elements of the slab starting at kk, into registers
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int q = ty + r * 8;
      const int ia = i0 + (a_lanes_k ? q : tx), ka = kk + (a_lanes_k ? tx : q);
      ra[r] = (ia < m && ka < k1) ? a[ia * a_si + ka * a_sk] : 0.0f;
      const int kb = kk + (b_lanes_j ? q : tx), jb = j0 + (b_lanes_j ? tx : q);
      rb[r] = (kb < k1 && jb < n) ? b[kb * b_sk + jb * b_sj] : 0.0f;
    }
  };
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  if (k0 < k1) fetch(k0);
  for (int kk = k0; kk < k1; kk += 32) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int q = ty + r * 8;
      if (a_lanes_k) sa[q][tx] = ra[r]; else sa[tx][q] = ra[r];
      if (b_lanes_j) sb[q][tx] = rb[r]; else sb[tx][q] = rb[r];
    }
    __syncthreads();
    if (kk + 32 < k1) fetch(kk + 32);   // the next slab's loads fly under this slab's arithmetic
#pragma unroll 8
    for (int q = 0; q < 32; ++q) {
      const float bv = sb[q][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fmaf(sa[ty + r * 8][q], bv, acc[r]);
    }
    __syncthreads();
  }
  const int j = j0 + tx;
  if (j >= n) return;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty + r * 8;
    if (i >= m) continue;
    float v = acc[r];
    float* dst = c + static_cast<size_t>(i) * n + j;
    if (kz > 1) { dst[static_cast<size_t>(blockIdx.z) * m * n] = v; continue; }
    if (bias) v += bias[j];
    if (accumulate) v += *dst;
    if (relu) v = fmaxf(v, 0.0f);
    *dst = v;
  }
}

__global__ void sum_partials_kernel(const float* __restrict__ partial, int splits, size_t count, float* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = partial[i];
  for (int k = 1; k < splits; ++k) s += partial[k * count + i];
  out[i] = s;
}

// ---- the heads' 1x1 convolutions (model_chess.py:77-78,86-87) over all rows, both heads in one pass over the tower's output ----
constexpr int HEADS_MAX_F = 8;   // policy + value filters

// zp[row][pf], zv[row][vf] = top[row][:] . wp[:, :], wv[:, :]   (w: Keras (1, 1, C, f) = [c][f]); one warp per row
__global__ void __launch_bounds__(256)
head_conv_forward_kernel(const float* __restrict__ top, const float* __restrict__ wp, const float* __restrict__ wv, float* __restrict__ zp,
                         float* __restrict__ zv, int rows, int c, int pf, int vf) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  float acc[HEADS_MAX_F];
#pragma unroll
  for (int f = 0; f < HEADS_MAX_F; ++f) acc[f] = 0.0f;
  for (int ch = lane; ch < c; ch += 32) {
    const float x = top[static_cast<size_t>(row) * c + ch];
#pragma unroll
    for (int f = 0; f < HEADS_MAX_F; ++f) {
      if (f < pf) acc[f] = fmaf(x, __ldg(wp + ch * pf + f), acc[f]);
      else if (f < pf + vf) acc[f] = fmaf(x, __ldg(wv + ch * vf + f - pf), acc[f]);
    }
  }
#pragma unroll
  for (int f = 0; f < HEADS_MAX_F; ++f) {
    if (f >= pf + vf) break;
    float v = acc[f];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) {
      if (f < pf) zp[static_cast<size_t>(row) * pf + f] = v;
      else zv[static_cast<size_t>(row) * vf + f - pf] = v;
    }
  }
}

// dtop[row][ch] = sum_f dzp[row][f] wp[ch][f] + sum_f dzv[row][f] wv[ch][f]: the gradient wrt the tower's output, written once
__global__ void head_conv_backward_data_kernel(const float* __restrict__ dzp, const float* __restrict__ dzv, const float* __restrict__ wp,
                                               const float* __restrict__ wv, float* __restrict__ dtop, int rows, int c, int pf, int vf) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(rows) * c) return;
  const int ch = static_cast<int>(i % c);
  const size_t row = i / c;
  float v = 0.0f;
  for (int f = 0; f < pf; ++f) v = fmaf(dzp[row * pf + f], __ldg(wp + ch * pf + f), v);
  for (int f = 0; f < vf; ++f) v = fmaf(dzv[row * vf + f], __ldg(wv + ch * vf + f), v);
  dtop[i] = v;
}

// partial[split][ch][f] = sum over this block's rows of top[row][ch] dz[row][f] (f: policy filters, then value filters);
// grid (c / 256, row splits); head_conv_weight_reduce_kernel adds the splits in a fixed order
__global__ void __launch_bounds__(256)
head_conv_backward_weight_kernel(const float* __restrict__ top, const float* __restrict__ dzp, const float* __restrict__ dzv,
                                 float* __restrict__ partial, int rows, int c, int pf, int vf) {
  const int ch = blockIdx.x * 256 + threadIdx.x;
  const int per = (rows + gridDim.y - 1) / gridDim.y, r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  __shared__ float sd[64][HEADS_MAX_F];
  float acc[HEADS_MAX_F];
#pragma unroll
  for (int f = 0; f < HEADS_MAX_F; ++f) acc[f] = 0.0f;
  const int ft = pf + vf;
  for (int rb = r0; rb < r1; rb += 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * ft; i += 256) {
      const int r = rb + i / ft, f = i % ft;
      sd[i / ft][f] = r < r1 ? (f < pf ? dzp[static_cast<size_t>(r) * pf + f] : dzv[static_cast<size_t>(r) * vf + f - pf]) : 0.0f;
    }
    __syncthreads();
    if (ch < c) {
      const int n = min(64, r1 - rb);
      for (int r = 0; r < n; ++r) {
        const float x = top[static_cast<size_t>(rb + r) * c + ch];
#pragma unroll
        for (int f = 0; f < HEADS_MAX_F; ++f)
          if (f < ft) acc[f] = fmaf(x, sd[r][f], acc[f]);
      }
    }
  }
  if (ch >= c) return;
  float* dst = partial + (static_cast<size_t>(blockIdx.y) * c + ch) * ft;
#pragma unroll
  for (int f = 0; f < HEADS_MAX_F; ++f)
    if (f < ft) dst[f] = acc[f];
}

__global__ void head_conv_weight_reduce_kernel(const float* __restrict__ partial, int splits, int c, int pf, int vf, float* __restrict__ dwp,
                                               float* __restrict__ dwv) {
  const int ft = pf + vf, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c * ft) return;
  float s = 0.0f;
  for (int k = 0; k < splits; ++k) s += partial[static_cast<size_t>(k) * c * ft + i];
  const int ch = i / ft, f = i - ch * ft;
  if (f < pf) dwp[ch * pf + f] = s;
  else dwv[ch * vf + f - pf] = s;
}

// column sums of a row-major matrix (Dense bias gradients): out[j] = sum_i x[i][j]
// block = 32 columns x 8 row lanes (256 threads); the row lanes' sums are added in a fixed order
__global__ void __launch_bounds__(256) column_sum_kernel(const float* __restrict__ x, int m, int n, float* __restrict__ out) {
  __shared__ double sh[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, j = blockIdx.x * 32 + tx;
  double s = 0.0;
  if (j < n)
    for (int i = ty; i < m; i += 8) s += x[static_cast<size_t>(i) * n + j];
  sh[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && j < n) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += sh[k][tx];
    out[j] = static_cast<float>(t);
  }
}

// Flatten of a channels_first tensor (model_chess.py:81,90): head activations [b*64 + s][f] <-> features [b][f*64 + s]
__global__ void flatten_kernel(float* __restrict__ rows_f, float* __restrict__ feat, int batch, int f, int to_rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * 64 * f) return;
  const int b = i / (64 * f), rem = i - b * 64 * f, ch = rem / 64, s = rem - ch * 64;
  const size_t r = (static_cast<size_t>(b) * 64 + s) * f + ch, q = static_cast<size_t>(b) * f * 64 + ch * 64 + s;
  if (to_rows) rows_f[r] = feat[q];
  else feat[q] = rows_f[r];
}

// ---- losses (optimize.py:100-103; keras losses on the TensorFlow backend) -------------------------------------------
// policy: softmax over the logits; categorical_crossentropy = -sum_j t_j log(clip(p_j / sum p, 1e-7, 1 - 1e-7));
//         d loss / d logits = p * sum_j t_j - t   (exact wherever the clip is inactive), scaled by weight / batch
// value : v = tanh(pre);  mean_squared_error = (v - t)^2;  d / d pre = 2 (v - t) (1 - v^2), scaled by weight / batch
// One block per position.  loss_acc[0] += CE, loss_acc[1] += MSE (sums over the batch, double).
__global__ void __launch_bounds__(256)
loss_kernel(const float* __restrict__ logits, const float* __restrict__ policy_t, const float* __restrict__ vpre,
            const float* __restrict__ value_t, int n_labels, float w_policy, float w_value, int batch,
            float* __restrict__ dlogits, float* __restrict__ dvpre, float* __restrict__ probs, float* __restrict__ vout,
            double* __restrict__ loss_acc) {
  __shared__ float s_red[8];
  __shared__ float s_bc;
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* row = logits + static_cast<size_t>(b) * n_labels;
  const float* trow = policy_t + static_cast<size_t>(b) * n_labels;
  auto block_reduce = [&](float v, bool is_max) -> float {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float u = __shfl_xor_sync(0xffffffffu, v, o);
      v = is_max ? fmaxf(v, u) : v + u;
    }
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float r = s_red[0];
      for (int k = 1; k < 8; ++k) r = is_max ? fmaxf(r, s_red[k]) : r + s_red[k];
      s_bc = r;
    }
    __syncthreads();
    return s_bc;
  };
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < n_labels; j += 256) mx = fmaxf(mx, row[j]);
  mx = block_reduce(mx, true);
  float se = 0.0f, st = 0.0f;
  for (int j = threadIdx.x; j < n_labels; j += 256) { se += expf(row[j] - mx); st += trow[j]; }
  se = block_reduce(se, false);
  st = block_reduce(st, false);
  float ce = 0.0f;
  for (int j = threadIdx.x; j < n_labels; j += 256) {
    const float p = expf(row[j] - mx) / se, t = trow[j];
    if (probs) probs[static_cast<size_t>(b) * n_labels + j] = p;
    dlogits[static_cast<size_t>(b) * n_labels + j] = (p * st - t) * (w_policy / batch);
    if (t != 0.0f) ce -= t * logf(fminf(fmaxf(p, 1e-7f), 1.0f - 1e-7f));
  }
  ce = block_reduce(ce, false);
  if (threadIdx.x == 0) {
    const float v = tanhf(vpre[b]), t = value_t[b];
    if (vout) vout[b] = v;
    dvpre[b] = 2.0f * (v - t) * (1.0f - v * v) * (w_value / batch);
    atomicAdd(&loss_acc[0], static_cast<double>(ce));
    atomicAdd(&loss_acc[1], static_cast<double>((v - t) * (v - t)));
  }
}

// dh *= (h > 0)   (ReLU backward of the value head's hidden layer)
__global__ void relu_backward_kernel(const float* __restrict__ h, float* __restrict__ dh, size_t total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < total && h[i] <= 0.0f) dh[i] = 0.0f;
}

// ---- weight gradient of a k x k "same" convolution on the CUDA cores (fp32 build) ---------------------------------------
//   dw[tap][ic][oc] = sum over boards b and squares s of  x[b][s + tap][ic] * dz[b][s][oc]      (zero outside the board)
// grid (oc tiles of 32, ic tiles of 32, taps * splits); a block walks its share of the boards, 64 squares at a time.
// dw must be zero on entry when splits > 1 (atomicAdd).
__global__ void __launch_bounds__(256)
conv_wgrad_fp32_kernel(const float* __restrict__ x, const float* __restrict__ dz, int batch, int cin, int cout, int ksize,
                       int splits, float* __restrict__ dw) {
  __shared__ float sx[64][33], sd[64][33];
  const int oc0 = blockIdx.x * 32, ic0 = blockIdx.y * 32;
  const int tap = blockIdx.z / splits, split = blockIdx.z - tap * splits;
  const int pad = (ksize - 1) / 2, dy = tap / ksize - pad, dx = tap % ksize - pad;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // thread: oc = oc0 + tx, ic = ic0 + ty + 8 r
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  for (int b = split; b < batch; b += splits) {
    // 64 squares x 32 channels of the shifted input and of dz
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int s = ty + q * 8;
      const int r = (s >> 3) + dy, f = (s & 7) + dx;
      const bool in = r >= 0 && r < 8 && f >= 0 && f < 8 && ic0 + tx < cin;
      sx[s][tx] = in ? x[(static_cast<size_t>(b) * 64 + r * 8 + f) * cin + ic0 + tx] : 0.0f;
      sd[s][tx] = oc0 + tx < cout ? dz[(static_cast<size_t>(b) * 64 + s) * cout + oc0 + tx] : 0.0f;
    }
    __syncthreads();
#pragma unroll 8
    for (int s = 0; s < 64; ++s) {
      const float dv = sd[s][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fmaf(sx[s][ty + r * 8], dv, acc[r]);
    }
    __syncthreads();
  }
  if (oc0 + tx >= cout) return;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int ic = ic0 + ty + r * 8;
    if (ic >= cin) continue;
    float* dst = dw + (static_cast<size_t>(tap) * cin + ic) * cout + oc0 + tx;
    if (splits > 1) atomicAdd(dst, acc[r]);
    else *dst = acc[r];
  }
}

// w'[tap'][oc][ic] = w[tap][ic][oc] with tap' the 180-degree rotation of tap: the weights of the INPUT-gradient convolution
// (dx = conv(dz, w')), in the same [tap][in][out] layout conv_fp32_kernel reads
__global__ void flip_weights_kernel(const float* __restrict__ w, float* __restrict__ wt, int taps, int cin, int cout) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(taps) * cin * cout;
  if (i >= total) return;
  const int oc = static_cast<int>(i % cout);
  const int ic = static_cast<int>((i / cout) % cin);
  const int tap = static_cast<int>(i / (static_cast<size_t>(cout) * cin));
  wt[(static_cast<size_t>(taps - 1 - tap) * cout + oc) * cin + ic] = w[i];
}

// ---- operand preparation for the tensor-core convolutions of the 16-bit training builds ------------------------------
// fp32 convolution kernel [tap][ic][oc] (Keras HWIO) ->
//   wf [oc][tap * cin_pad + ic]               K-major rows for the FORWARD implicit GEMM (16-bit, format f_fp16)
//   wb [ic][tap' * cout + oc], tap' = 180-degree rotation of tap:  K-major rows for the INPUT-GRADIENT convolution
//                                                                  dx = conv(dz, rot180(w)^T) (16-bit, format b_fp16)
__global__ void pack_conv_weights_kernel(const float* __restrict__ w, int taps, int cin, int cin_pad, int cout, uint16_t* __restrict__ wf,
                                         int f_fp16, uint16_t* __restrict__ wb, int b_fp16) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(taps) * cin * cout;
  if (i >= total) return;
  const int oc = static_cast<int>(i % cout);
  const int ic = static_cast<int>((i / cout) % cin);
  const int tap = static_cast<int>(i / (static_cast<size_t>(cout) * cin));
  const float v = w[i];
  auto bits = [](float x, int fp16) -> uint16_t {
    if (fp16) { const __half h = __float2half_rn(fminf(fmaxf(x, -65504.0f), 65504.0f)); return *reinterpret_cast<const uint16_t*>(&h); }
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    return *reinterpret_cast<const uint16_t*>(&h);
  };
  wf[static_cast<size_t>(oc) * taps * cin_pad + static_cast<size_t>(tap) * cin_pad + ic] = bits(v, f_fp16);
  if (wb) wb[static_cast<size_t>(ic) * taps * cout + static_cast<size_t>(taps - 1 - tap) * cout + oc] = bits(v, b_fp16);
}

// the same for every convolution of the network in ONE launch: blockIdx.y = layer
struct PackLayer { const float* w; uint16_t* wf; uint16_t* wb; int taps, cin, cin_pad, cout; };
constexpr int PACK_MAX_LAYERS = 48;
struct PackTable { PackLayer layer[PACK_MAX_LAYERS]; };
// a block walks (tap, 32 ic x 32 oc) tiles: reads and the wb stores run along oc, the wf stores along ic (through shared memory)
__global__ void __launch_bounds__(256) pack_all_conv_weights_kernel(const PackTable tab, int f_fp16, int b_fp16) {
  const PackLayer& L = tab.layer[blockIdx.y];
  __shared__ float tile[32][33];
  auto bits = [](float x, int fp16) -> uint16_t {
    if (fp16) { const __half h = __float2half_rn(fminf(fmaxf(x, -65504.0f), 65504.0f)); return *reinterpret_cast<const uint16_t*>(&h); }
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    return *reinterpret_cast<const uint16_t*>(&h);
  };
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ict = (L.cin + 31) / 32, oct = (L.cout + 31) / 32, tiles = L.taps * ict * oct;
  for (int tl = blockIdx.x; tl < tiles; tl += gridDim.x) {
    const int tap = tl / (ict * oct), ic0 = (tl / oct) % ict * 32, oc0 = tl % oct * 32;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ic = ic0 + ty + r * 8, oc = oc0 + tx;
      float v = 0.0f;
      if (ic < L.cin && oc < L.cout) {
        v = L.w[(static_cast<size_t>(tap) * L.cin + ic) * L.cout + oc];
        if (L.wb) L.wb[static_cast<size_t>(ic) * L.taps * L.cout + static_cast<size_t>(L.taps - 1 - tap) * L.cout + oc] = bits(v, b_fp16);
      }
      tile[ty + r * 8][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int oc = oc0 + ty + r * 8, ic = ic0 + tx;
      if (ic < L.cin && oc < L.cout)
        L.wf[static_cast<size_t>(oc) * L.taps * L.cin_pad + static_cast<size_t>(tap) * L.cin_pad + ic] = bits(tile[tx][ty + r * 8], f_fp16);
    }
  }
}

// rows [board * 64 + square][c] fp32  ->  K-major operand of the weight-gradient GEMM, bf16:
//   out[ch][cell][board],  cell = (rank + pad) * pitch + file + pad over the zero-padded board (pitch = 8 + 2 pad),
//   board < nb (nb = boards rounded up to 64; boards >= batch and channels >= c are written as zeros, halo cells stay zero
//   from the allocation).  With this order a filter tap is a shift of the K index by a multiple of nb elements: 16-byte
//   aligned, as the TMA box origin of the shifted operand has to be.
// grid (nb / 64, (c_pad + 63) / 64, 64 squares), 256 threads: a block turns 64 boards x 64 channels of one square
// (reads: 64 channels = 256 B per board; stores: 64 boards = 128 B per channel, two boards per thread)
__global__ void __launch_bounds__(256)
transpose_pad_kernel(const float* __restrict__ x, int batch, int c, int c_pad, int nb, int pad, __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[64][65];
  const int b0 = blockIdx.x * 64, c0 = blockIdx.y * 64, sq = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int pitch = 8 + 2 * pad, cells = pitch * pitch, cell = ((sq >> 3) + pad) * pitch + (sq & 7) + pad;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int b = b0 + ty + r * 8;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int ch = c0 + tx + q * 32;
      tile[ty + r * 8][tx + q * 32] = (b < batch && ch < c) ? x[(static_cast<size_t>(b) * 64 + sq) * c + ch] : 0.0f;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int chl = ty + r * 8, ch = c0 + chl, b = b0 + 2 * tx;   // nb is a multiple of 64: b + 1 < nb whenever b < nb
    if (ch < c_pad && b < nb)
      *reinterpret_cast<__nv_bfloat162*>(out + (static_cast<size_t>(ch) * cells + cell) * nb + b) =
          __floats2bfloat162_rn(tile[2 * tx][chl], tile[2 * tx + 1][chl]);
  }
}

// partial [tap][split][n_out][n_in_pad] (fp32, one tensor-core GEMM tile each) -> dw [tap][ic][oc] (Keras layout), summed over the splits
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int taps, int splits, int n_out, int n_in, int n_in_pad, float* __restrict__ dw) {
  __shared__ float tile[32][33];
  const int ic0 = blockIdx.x * 32, oc0 = blockIdx.y * 32, tap = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int oc = oc0 + ty + r * 8, ic = ic0 + tx;
    float s = 0.0f;
    if (oc < n_out && ic < n_in)
      for (int k = 0; k < splits; ++k) s += partial[((static_cast<size_t>(tap) * splits + k) * n_out + oc) * n_in_pad + ic];
    tile[ty + r * 8][tx] = s;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int ic = ic0 + ty + r * 8, oc = oc0 + tx;
    if (ic < n_in && oc < n_out) dw[(static_cast<size_t>(tap) * n_in + ic) * n_out + oc] = tile[tx][ty + r * 8];
  }
}

// ---- Adam (keras 2.0.8 optimizers.Adam: lr 1e-3, beta_1 0.9, beta_2 0.999, epsilon 1e-8, decay 0) ----------------------
//   g += 2 * l2 * p on regularised kernels (d/dp of l2 * sum p^2);  lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= lr_t * m / (sqrt(v) + eps)
// seg: per parameter 0 = frozen, 1 = trainable, 2 = trainable + l2, 3 = BatchNormalization moving statistic (its "gradient"
// slot holds this step's batch statistic, summed over the ranks: moving = moving * momentum + mean batch statistic * (1 - momentum))
// l2_out (zeroed by the caller): sum of squares of the UPDATED l2-regularised parameters, the regularisation part of the next
// step's loss (saves a pass over the parameters per step)
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            const uint8_t* __restrict__ seg, size_t n, float lr_t, float b1, float b2, float eps, float l2,
            float grad_scale, float bn_momentum, double* __restrict__ l2_out) {
  double sq = 0.0;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint8_t sg = seg[i];
    if (sg == 0) continue;
    if (sg == 3) {
      p[i] = p[i] * bn_momentum + g[i] * grad_scale * (1.0f - bn_momentum);
      continue;
    }
    float gi = g[i] * grad_scale;
    const float pi = p[i];
    if (sg == 2) gi += 2.0f * l2 * pi;
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float pn = pi - lr_t * mi / (sqrtf(vi) + eps);
    p[i] = pn;
    if (sg == 2) sq += static_cast<double>(pn) * pn;
  }
  __shared__ double sh[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0 && l2_out) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += sh[k];
    atomicAdd(l2_out, t);
  }
}

// sum of squares of the l2-regularised parameters (reported as the regularisation part of the loss)
__global__ void l2_sum_kernel(const float* __restrict__ p, const uint8_t* __restrict__ seg, size_t n, double* __restrict__ out) {
  double s = 0.0;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    if (seg[i] == 2) s += static_cast<double>(p[i]) * p[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

}  // namespace train
}  // namespace caz
