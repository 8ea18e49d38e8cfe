/*
 * caz_b200.h -- C ABI of the B200 (sm_100a) policy/value evaluator for chess-alpha-zero.
 *
 * This is the drop-in boundary for ONE path of the reference: the batched network
 * evaluation that ChessModelAPI._predict_batch_worker performs with
 *     policy_ary, value_ary = self.agent_model.model.predict_on_batch(data)
 * (reference src/chess_zero/agent/api_chess.py:66-67), plus the two small kernels
 * either side of it that north_star names: the input-plane encoder
 * (src/chess_zero/env/chess_env.py:231-348) and PUCT select
 * (src/chess_zero/agent/player_chess.py:251-299).
 *
 * Plain C types only: no torch, no Python objects, no exceptions across the boundary.
 * Every entry point returns an int status (0 = CAZ_OK); text for the last failure is
 * available from caz_last_error().  Buffers are caller-owned; one in-flight call per handle
 * (the reference has exactly one prediction_worker thread per model, api_chess.py:37-39);
 * several handles per process are fine (evaluate.py runs two models).
 *
 * There is no CPU fallback anywhere behind this header.
 */
#ifndef CAZ_B200_H
#define CAZ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAZ_ABI_VERSION 1

/* status codes */
#define CAZ_OK 0
#define CAZ_ERR_INVALID 1     /* bad argument / unsupported architecture        */
#define CAZ_ERR_CUDA 2        /* CUDA runtime or driver error (text in last_error) */
#define CAZ_ERR_NO_DEVICE 3   /* no sm_100 device visible                       */
#define CAZ_ERR_KERNEL 4      /* a kernel watchdog fired (pipeline never completed) */

/* numeric builds (BASELINE.json north_star: "fp32-accumulate build" and "bf16 build") */
#define CAZ_PRECISION_BF16 0  /* THE bf16 build: every convolution weight is a bf16 VALUE (BN-folded in fp32, rounded to
                                 bf16); fp32 TMEM accumulators, fp32 residual stream, fp32 heads.  The 16-bit operand COPIES
                                 of the activations are fp16, saturated to +-65504: activation rounding was two thirds of
                                 the value head's error variance with bf16 operands (profiles/r2_precision_ablation.md:
                                 max|dv| 0.8-0.95e-2 instead of 1.6-2.4e-2).  tcgen05.mma kind::f16 refuses A = fp16 with
                                 B = bf16 (illegal instruction on sm_100a, measured), so the bf16-valued weights are handed
                                 over in the fp16 ENCODING, which holds every bf16 value with 2^-17 <= |w| < 65504 exactly;
                                 caz_inexact_weights() counts the ones outside (rounded to nearest / saturated). */
#define CAZ_PRECISION_FP32 1  /* fp32 operands and accumulation on the CUDA cores (1e-4 parity build)      */
#define CAZ_PRECISION_FP16 2  /* the tensor-core build with fp16 weights AND activations: same kernels and throughput,
                                 8x finer operand rounding; weights below 6e-5 lose bits (fp16 subnormals)         */
#define CAZ_PRECISION_BF16_PURE 3 /* bf16 weights and bf16 activation operands (what round 1 called BF16): kept for
                                 comparison; misses north_star's 2e-2 on the value of the reference-init net in about
                                 half of the summation orders                                                      */

/* eval flags */
#define CAZ_EVAL_DEFAULT 0
#define CAZ_EVAL_MASKED_SOFTMAX 1 /* legal_mask given: softmax over mask!=0 entries only, 0 elsewhere */

#define CAZ_PLANES 18
#define CAZ_SQUARES 64
#define CAZ_PLANE_FLOATS (CAZ_PLANES * CAZ_SQUARES) /* 1152 floats = one (18,8,8) observation */
#define CAZ_RECORD_BYTES 68                         /* compact board record, see caz_encode   */

typedef struct caz_handle caz_handle;

/*
 * Network architecture, as read from the Keras Model.get_config() JSON that
 * ChessModel.load feeds to Model.from_config (model_chess.py:144-145); graph
 * shape per ChessModel.build/_build_residual_block (model_chess.py:57-111).
 */
typedef struct caz_arch {
  int32_t input_planes;   /* 18  (InputLayer batch_input_shape[1])              */
  int32_t stem_kernel;    /* 5   (cnn_first_filter_size; 3 in the shipped JSON) */
  int32_t filters;        /* 256 (cnn_filter_num)                               */
  int32_t res_blocks;     /* 7   (res_layer_num)                                */
  int32_t block_kernel;   /* 3   (cnn_filter_size)                              */
  int32_t policy_filters; /* 2                                                  */
  int32_t value_filters;  /* 4   (1 in the shipped JSON)                        */
  int32_t value_fc;       /* 256 (value_fc_size)                                */
  int32_t n_labels;       /* 1968 (Config.n_labels, config.py:143-146)          */
  float bn_epsilon;       /* 1e-3 (BatchNormalization epsilon in the JSON)      */
} caz_arch;

/*
 * One fp32 tensor in Keras layout, exactly as Keras stores it in the weight file:
 * Conv2D kernel (kh, kw, Cin, Cout); Dense kernel (in, out) and bias (out);
 * BatchNormalization gamma, beta, moving_mean, moving_variance (C each).
 */
typedef struct caz_tensor {
  const float* data;
  int64_t numel;
} caz_tensor;

/*
 * Tensor order expected by caz_create / caz_reload (graph order of model_chess.py:62-92):
 *   stem:      kernel, gamma, beta, mean, var
 *   block i:   conv1 kernel, bn1 gamma, beta, mean, var, conv2 kernel, bn2 gamma, beta, mean, var
 *   policy:    conv kernel, gamma, beta, mean, var, dense kernel, dense bias
 *   value:     conv kernel, gamma, beta, mean, var, dense1 kernel, dense1 bias, dense2 kernel, dense2 bias
 * => 5 + 10*res_blocks + 7 + 9 tensors.  BatchNorm is folded into the convolutions here,
 * in fp32, on the host side of this library (w' = w*gamma*rsqrt(var+eps), b' = beta - mean*that).
 */

/* Replaces ChessModel.build/.load (model_chess.py:57-94,121-153): materialise the network on `device`.
 * max_batch bounds the positions one device pass holds (larger eval calls are chunked). */
int caz_create(const caz_arch* arch, const caz_tensor* tensors, int32_t n_tensors, int32_t device,
               int32_t precision, int32_t max_batch, caz_handle** out);

/* Replaces the weight swap of reload_best_model_weight_if_changed (lib/model_helper.py:26-41):
 * same architecture, new tensors; pipes/handles stay open. */
int caz_reload(caz_handle* h, const caz_tensor* tensors, int32_t n_tensors);

void caz_destroy(caz_handle* h);

/* Replaces model.predict_on_batch(data) (api_chess.py:67).
 *   planes : host float32 [batch][18][8][8], C-contiguous (api_chess.py:66)
 *   policy : host float32 [batch][n_labels]   (softmax; rows sum to 1)
 *   value  : host float32 [batch]             (tanh)
 *   legal_mask : NULL, or host uint8 [batch][n_labels] used with CAZ_EVAL_MASKED_SOFTMAX
 * Host<->device copies and the stream synchronisation are inside the call. batch == 0 is a no-op. */
int caz_eval(caz_handle* h, const float* planes, int32_t batch, float* policy, float* value,
             const uint8_t* legal_mask, uint32_t flags);

/* Same computation with every buffer already resident on the handle's device; enqueues on the
 * handle's stream and returns without synchronising (caz_sync waits). batch <= max_batch. */
int caz_eval_device(caz_handle* h, const float* d_planes, int32_t batch, float* d_policy, float* d_value,
                    const uint8_t* d_legal_mask, uint32_t flags);

/* Same, but from compact board records (see caz_encode): the encoder kernel writes the input planes
 * on the device, so only 68 bytes per position cross PCIe. records: host uint8 [batch][68]. */
int caz_eval_records(caz_handle* h, const uint8_t* records, int32_t batch, float* policy, float* value,
                     const uint8_t* legal_mask, uint32_t flags);

/* Pipelined form of caz_eval for callers that keep several batches in flight (batch <= max_batch): caz_submit enqueues
 * H2D, forward pass and D2H on three streams and returns at once with a ticket; caz_collect(ticket) blocks until
 * that batch's outputs are in the caller's host buffers.  Batch i+1's H2D and batch i-1's D2H then run under batch
 * i's forward pass.  Tickets count up from 0 (they wrap at 2^30); each has its own completion event, so collecting a
 * ticket never waits for a later submission.  Up to 16 tickets may be outstanding (a 17th caz_submit fails with
 * CAZ_ERR_INVALID until one is collected); with more than two in flight the later ones queue on the streams behind
 * the two staging slots.  A ticket is collected exactly once; collecting an unknown ticket is CAZ_ERR_INVALID.
 * Threads: the submit calls follow the one-call-per-handle rule; caz_collect touches only its ticket and may run on
 * another thread concurrently with a submit.  Host buffers must stay valid (and should be pinned, caz_host_alloc)
 * until collected. */
int caz_submit(caz_handle* h, const float* planes, int32_t batch, float* policy, float* value,
               const uint8_t* legal_mask, uint32_t flags, int32_t* ticket);
/* caz_submit from board records (68 bytes per position, see caz_eval_records): what a batched search driver with two or
 * more groups of games calls, so that one group's policy read-back runs under another group's forward pass. */
int caz_submit_records(caz_handle* h, const uint8_t* records, int32_t batch, float* policy, float* value,
                       const uint8_t* legal_mask, uint32_t flags, int32_t* ticket);
/* caz_submit_records for a search driver that knows the legal moves of its leaves: instead of the whole policy row
 * (7 872 bytes per position) the device pushes the priors of select_action_q_and_u (player_chess.py:267-275) itself and
 * returns one float64 per legal move, priors[offsets[i] + e] = P_i[labels[offsets[i] + e]] / (1e-8 + sum over the
 * position's labels, left to right) -- bit-identical to caz_push_priors on the full row (labels < 0 count as 0).
 *   offsets: host int32 [batch + 1] (CSR, offsets[0] = 0), labels: host int32 [offsets[batch]],
 *   priors: host float64 [offsets[batch]] out, value: host float32 [batch] out; collected with caz_collect. */
#define CAZ_MAX_MOVES 256 /* bound on the legal moves of one position (218 is the known maximum) */
int caz_submit_records_priors(caz_handle* h, const uint8_t* records, int32_t batch, const int32_t* offsets,
                              const int32_t* labels, double* priors, float* value, int32_t* ticket);
int caz_collect(caz_handle* h, int32_t ticket);

int caz_sync(caz_handle* h);

/* Replaces canon_input_planes (env/chess_env.py:231-348) for a batch of positions.
 * record layout (68 bytes): [0..63] square codes in FEN text order (rank 8 first, file a..h) of the
 * position as written: 0 empty, 1..12 = index into "KQRBNPkqrbnp" + 1; [64] bit0 black to move,
 * bits1..4 castling K,Q,k,q as written; [65] en-passant square (8-rank)*8+file or 255; [66..67]
 * halfmove clock, little endian.  planes: host float32 [batch][18][8][8], bit-exact with the reference. */
int caz_encode(caz_handle* h, const uint8_t* records, int32_t batch, float* planes);

/* Replaces the prior push of select_action_q_and_u (player_chess.py:267-275) for many nodes at once.
 *   policy      : host float32 [n_nodes][n_labels]  (one network policy row per node)
 *   offsets     : host int32 [n_nodes+1], CSR offsets into the edge arrays
 *   label_index : host int32 [n_edges], move label of each legal edge (self.move_lookup[mov])
 *   priors      : host float64 [n_edges] out: P[idx] / (1e-8 + sum of the node's P[idx], left to right) */
int caz_push_priors(caz_handle* h, const float* policy, int32_t n_labels, const int32_t* offsets,
                    int32_t n_nodes, const int32_t* label_index, double* priors);

/* Replaces the scoring loop of select_action_q_and_u (player_chess.py:277-299), one warp per node.
 *   n, q, p   : host int32 / float64 / float64 [n_edges] (ActionStats.n, .q, .p; player_chess.py:33-55)
 *   sum_n     : host int32 [n_nodes] (VisitStats.sum_n)
 *   noise     : NULL or host float64 [n_edges], Dirichlet draw for nodes with is_root[i] != 0
 *   is_root   : NULL or host uint8 [n_nodes]
 *   best_edge : host int32 [n_nodes] out: index WITHIN the node of the first maximum of
 *               q + c_puct*p*sqrt(sum_n+1)/(1+n)  (strict '>' from -999, so -1 for an empty node) */
int caz_puct_select(caz_handle* h, const int32_t* offsets, int32_t n_nodes, const int32_t* n, const double* q,
                    const double* p, const int32_t* sum_n, const double* noise, const uint8_t* is_root,
                    double c_puct, double noise_eps, int32_t* best_edge);
/* The same on DEVICE-resident node arrays (a driver that keeps its trees on the GPU): enqueues the kernel on the handle's
 * stream and returns; no copies, no synchronisation (caz_sync or a stream event before reading d_best_edge). */
int caz_puct_select_device(caz_handle* h, const int32_t* d_offsets, int32_t n_nodes, const int32_t* d_n, const double* d_q,
                            const double* d_p, const int32_t* d_sum_n, const double* d_noise, const uint8_t* d_is_root,
                            double c_puct, double noise_eps, int32_t* d_best_edge);

/* --- introspection / measurement ------------------------------------------------------------ */
const char* caz_last_error(const caz_handle* h); /* h == NULL: last error of a failed caz_create */
int32_t caz_abi_version(void);
int32_t caz_max_batch(const caz_handle* h);
int32_t caz_precision(const caz_handle* h);
void* caz_stream(const caz_handle* h);            /* cudaStream_t the handle launches on */
int64_t caz_kernel_launches(const caz_handle* h); /* kernels this handle has launched so far */
int64_t caz_weight_bytes(const caz_handle* h);    /* folded device weight set, bytes */
int64_t caz_inexact_weights(const caz_handle* h); /* bf16 build: folded bf16 weights not exactly representable in the fp16
                                                     encoding the tensor core reads (0 for every network in the tests) */

/* Batches of at most `limit` positions run the persistent small-batch tower kernel (one cooperative launch for
 * stem + all residual convs, weights streamed from L2); larger ones run one implicit-GEMM launch per layer.
 * limit < 0 restores the automatic choice (74 positions), 0 disables the small-batch kernel; the kernel itself goes up
 * to 148 positions (limits above that are clamped).  Results are identical either way up to fp32 summation order inside
 * the MMA. */
int caz_set_small_batch_limit(caz_handle* h, int32_t limit);

/* Batches beyond the small-batch kernel run stem + tower + head convolutions as ONE persistent launch in which every
 * CTA pair carries its own boards through all layers.  mode -1 (default): the form is chosen by batch size; forced
 * forms: 1 = one four-board group per pair, 2 = two groups interleaved layer by layer.  mode 0 selects the older chain
 * of one implicit-GEMM launch per layer.  Results are bit-identical in every mode. */
int caz_set_fused_tower(caz_handle* h, int32_t mode);
/* Test hook: per-layer clock64 stamps of the one-launch tower kernel (pair 0's leader CTA, its first super-group, group
 * 0; stamps: host int64 [n <= 512], 8 per layer: activations ready, tile requested, first chunk landed, MMAs issued,
 * accumulator complete, epilogue done, unused, unused). */
int caz_debug_tower_trace(caz_handle* h, int32_t enable, int64_t* stamps, int32_t n);

/* Test hook: switch on per-layer clock64 stamps of CTA 0 of the small-batch kernel and/or read them back
 * (stamps: host int64 [n], 8 per layer: barrier passed, first A tile landed, last A tile landed, MMAs issued,
 * accumulator ready, stores fenced, epilogue warps joined, unused). */
int caz_debug_sb_trace(caz_handle* h, int32_t enable, int64_t* stamps, int32_t n);

/* Per-stage device timing of the forward pass (CUDA events on the handle's stream).
 * stage ids: 0 input pack, 1 stem conv, 2 residual tower convs, 3 heads. After caz_profile(h,1),
 * every eval accumulates; caz_profile_read returns accumulated milliseconds and launch counts. */
#define CAZ_STAGES 4
int caz_profile(caz_handle* h, int32_t enable);
int caz_profile_read(caz_handle* h, float* ms_per_stage, int64_t* launches_per_stage, int32_t reset);

/* Test hook: run ONE convolution layer of the installed network (0 = stem, 1.. = conv1/conv2 of each
 * residual block, in order) on caller-supplied NHWC activations, through the build's real kernel.
 *   in       : host float32 [batch][64][cin]   (cin = input_planes for layer 0, else filters)
 *   residual : NULL or host float32 [batch][64][filters]
 *   out      : host float32 [batch][64][filters] = relu(conv(in) * bn_scale + bn_shift (+ residual))
 * In the bf16 build inputs are rounded to bf16 on the way in and the bf16 result is widened on the way out. */
int caz_debug_conv_layer(caz_handle* h, int32_t layer, const float* in, const float* residual, int32_t batch,
                         float* out);

/* Pinned host memory for callers that want true async DMA (numpy arrays can wrap it). */
void* caz_host_alloc(int64_t bytes);
void caz_host_free(void* p);

/* ---------------------------------------------------------------------------------------------------------------
 * Training step: what the reference's `opt` worker runs per batch,
 *     model.compile(Adam(), ['categorical_crossentropy', 'mean_squared_error'], loss_weights=[1.25, 1.0])
 *     model.fit(state_ary, [policy_ary, value_ary], batch_size=384, ...)        (worker/optimize.py:79-103)
 * on the graph of agent/model_chess.py:57-111: BatchNormalization in training mode (batch statistics, moving averages),
 * l2(1e-4) on every Conv2D / Dense kernel.  A trainer is separate from an evaluator handle; its weights go back
 * through caz_train_read(what = 0) in the tensor order of caz_create (-> ChessModel.save, caz_reload).
 * Multi-GPU: one trainer per GPU on its shard of the batch; the flat gradient vector (caz_train_grad_device_ptr, or
 * caller-provided device memory) is summed over ranks between caz_train_forward_backward and caz_train_apply -- the
 * Python side does that with torch.distributed.all_reduce (NCCL) -- and caz_train_apply takes 1 / world as grad_scale.
 * --------------------------------------------------------------------------------------------------------------- */
typedef struct caz_trainer caz_trainer;

typedef struct caz_train_config {
  float learning_rate;      /* 1e-3  keras.optimizers.Adam defaults (keras 2.0.8)        */
  float beta_1;             /* 0.9                                                        */
  float beta_2;             /* 0.999                                                      */
  float epsilon;            /* 1e-8                                                       */
  float l2;                 /* 1e-4  ModelConfig.l2_reg (configs/normal.py:64)            */
  float policy_loss_weight; /* 1.25  TrainerConfig.loss_weights (configs/normal.py:57)    */
  float value_loss_weight;  /* 1.0                                                        */
  float bn_momentum;        /* 0.99  BatchNormalization default                           */
} caz_train_config;

/* tensors: the initial weights, same order and layout as caz_create.  precision: CAZ_PRECISION_FP32 (all arithmetic fp32
 * on the CUDA cores: the parity build) or CAZ_PRECISION_BF16 (mixed precision: fp32 master weights, moments, BatchNormalization
 * and heads; the stem's and the tower's convolutions on the tensor cores with fp32 accumulation -- forward fp16 x fp16,
 * input gradients and weight gradients bf16 x bf16).  d_grads: NULL, or device memory for caz_train_param_count(arch) floats that receives the gradients. */
int caz_train_create(const caz_arch* arch, const caz_tensor* tensors, int32_t n_tensors, int32_t device, int32_t precision,
                     int32_t max_batch, const caz_train_config* cfg, float* d_grads, caz_trainer** out);
void caz_train_destroy(caz_trainer* t);
const char* caz_train_last_error(const caz_trainer* t);
int64_t caz_train_param_count(const caz_arch* arch); /* floats in the flat parameter / gradient vector (moving statistics included) */
int64_t caz_train_num_params(const caz_trainer* t);
int64_t caz_train_steps(const caz_trainer* t);       /* caz_train_apply calls so far (Adam's t) */
/* Forward pass in training mode, losses, backward pass: fills the gradient vector (sum over this batch of the loss
 * gradients, already divided by `batch`; the l2 term is added in caz_train_apply; the slots of the BatchNormalization moving
 * statistics receive this batch's mean / variance, which caz_train_apply folds into the moving averages -- after an
 * all-reduce they are the mean over the ranks, so every rank keeps identical statistics).  planes: host float32 [batch][18][8][8]; policy_target: host float32 [batch][n_labels]; value_target: host
 * float32 [batch].  losses (may be NULL): host float32 [4] = total, mean categorical cross-entropy, mean squared error,
 * l2 term. */
int caz_train_forward_backward(caz_trainer* t, const float* planes, const float* policy_target, const float* value_target,
                               int32_t batch, float* losses);
/* The same in two halves, for overlapping the data-parallel all-reduce with the backward pass: _begin enqueues everything
 * on the trainer's stream and returns; caz_train_wait_gradients(t, part, stream) makes `stream` (a cudaStream_t of the
 * caller, e.g. the one the NCCL all-reduce is issued from) wait until part 0 = the gradient elements
 * [caz_train_gradient_split(t), n) -- the upper half of the tower and the heads, which the backward pass finishes first --
 * or part 1 = all of them are complete; _end waits for the stream and returns the losses. */
int caz_train_forward_backward_begin(caz_trainer* t, const float* planes, const float* policy_target, const float* value_target,
                                     int32_t batch);
int caz_train_wait_gradients(caz_trainer* t, int32_t part, void* stream);
int64_t caz_train_gradient_split(const caz_trainer* t);
int caz_train_forward_backward_end(caz_trainer* t, float* losses);
/* One Adam update from grad_scale * gradient vector (+ 2 * l2 * w on kernels). */
int caz_train_apply(caz_trainer* t, float grad_scale);
/* what: 0 parameters, 1 gradients, 2 / 3 Adam first / second moments; host float32 [caz_train_num_params] */
int caz_train_read(caz_trainer* t, int32_t what, float* host, int64_t count);
/* softmax policy and tanh value of the last forward pass (training-mode BatchNormalization): host [batch][n_labels], [batch] */
int caz_train_read_outputs(caz_trainer* t, int32_t batch, float* policy, float* value);
float* caz_train_grad_device_ptr(caz_trainer* t);

#ifdef __cplusplus
}
#endif
#endif /* CAZ_B200_H */
