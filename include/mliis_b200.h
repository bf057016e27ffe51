/*
 * mliis_b200.h - C ABI of the B200-native inner-loop adaptation engine for ml4ai/mliis.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own:
 * its "operator API" is the model-object contract consumed by the meta-learner, i.e.
 *   sess.run(model.minimize_op, {input_ph, label_ph[, lr_ph, drop_rate_ph]})
 *       (meta_learners/supervised_reptile/supervised_reptile/reptile.py:115-121, :269-279, :639-643)
 *   sess.run(model.predictions, {input_ph, is_training_ph: False})          (reptile.py:503-506, :520)
 *   VariableState.export_variables() / import_variables()                     (meta_learners/variables.py:70-80)
 *   numpy list arithmetic of the meta-update                                  (meta_learners/variables.py:9-45)
 * Each entry point below names the reference interface it replaces.
 *
 * Conventions
 *   - plain C, no C++ types or exceptions cross the boundary;
 *   - every function returns 0 on success, a negative mliis_status otherwise;
 *     mliis_last_error() returns a thread-local message for the last failure;
 *   - every pointer marked "dev" is a DEVICE pointer into caller-owned memory (the Python host
 *     owns all buffers through torch); the library never frees caller memory;
 *   - "stream" is a cudaStream_t passed as void*; NULL means the legacy default stream;
 *   - tensors are fp32 NHWC exactly as the reference feeds them (images in [0,255],
 *     labels [B,H,W,2] with channel 1 = class of interest);
 *   - one ctx per (process, GPU); a ctx is not thread-safe; distinct ctxs are independent;
 *   - there is NO CPU fallback: on a device that is not sm_100 every compute call fails with
 *     MLIIS_ERR_DEVICE.
 */
#ifndef MLIIS_B200_H_
#define MLIIS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mliis_ctx mliis_ctx;

enum mliis_status {
  MLIIS_OK = 0,
  MLIIS_ERR_ARG = -1,      /* bad argument (ValueError in the reference)               */
  MLIIS_ERR_CUDA = -2,     /* a CUDA runtime call or kernel launch failed               */
  MLIIS_ERR_DEVICE = -3,   /* not an sm_100 device / no device: no fallback exists      */
  MLIIS_ERR_STATE = -4     /* call sequence error (e.g. backward without forward)       */
};

enum mliis_optimizer {     /* meta_learners/args.py:151-154 */
  MLIIS_OPT_ADAM = 0,      /* tf.train.AdamOptimizer(beta1=0) */
  MLIIS_OPT_SGD = 1        /* tf.train.GradientDescentOptimizer (--sgd) */
};

enum mliis_loss_flags {    /* models/efficientlab.py:294-313 */
  MLIIS_LOSS_DICE = 1,     /* loss = CE - ln(2 IoU/(IoU+1))  (--loss_name bce_dice) */
  MLIIS_LOSS_L2 = 2        /* + 0.0005 * sum l2_loss(non-BN vars)  (--l2)           */
};

enum mliis_gemm_mode {     /* numeric mode of the dense contractions */
  MLIIS_GEMM_FP32 = 0,     /* fp32 FFMA kernels (exact-order reference mode)                */
  MLIIS_GEMM_TF32 = 1,     /* tcgen05: single-pass TF32 on the decoder convs, 3xTF32 on the backbone (opt-in) */
  MLIIS_GEMM_TF32X3 = 2    /* tcgen05 3xTF32 split everywhere (fp32-class accuracy; the parity-clean default) */
};

typedef struct mliis_config {
  int32_t image_size;          /* --image_size (square); must be a multiple of 32          */
  int32_t max_batch;           /* largest batch any call will use (inner batch / query set) */
  int32_t n_slots;             /* independent task slots (each has its own state/workspace) */
  int32_t optimizer;           /* mliis_optimizer                                           */
  int32_t loss_flags;          /* mliis_loss_flags                                          */
  int32_t gemm_mode;           /* mliis_gemm_mode                                           */
  float   label_smoothing;     /* --label_smoothing                                         */
  float   final_dropout_rate;  /* --final_layer_dropout_rate (0 = layer absent)             */
  int32_t rsd[4];              /* --rsd reduction indices, 0-terminated (canonical {2,4})   */
  int32_t n_classes;           /* 0/1: the binary few-shot head (2 channels, dense one-hot labels).  > 1: the
                                  joint-training head of joint_train.py:307 - n_classes + 1 channels (channel 0 =
                                  background, --seperate_background_channel), multi-class soft IoU
                                  (binary_iou_loss=False, efficientlab.py:369-382), SPARSE labels: dev_labels is a
                                  mask pool [n,H,W] (> 0.5 = foreground) + mliis_set_class_ids             */
} mliis_config;

/* One trainable variable of the reference graph (tf.trainable_variables() order). */
typedef struct mliis_param_info {
  const char* name;            /* expected TF variable name                                */
  int64_t     offset;          /* float offset into the engine's flat parameter buffer     */
  int64_t     size;            /* number of floats                                         */
  int32_t     ndim;
  int32_t     shape[4];        /* TF layout: HWIO kernels, [k,k,C,1] depthwise, [C] vectors */
  int32_t     l2;              /* 1 if included in regularizers.l2_term                    */
} mliis_param_info;

/* One BatchNorm layer: moving_mean at bn_state[offset..], moving_variance at bn_state[n_bn+offset..]. */
typedef struct mliis_bn_info {
  const char* scope;
  int32_t     channels;
  int32_t     offset;
  int32_t     fused;           /* 1 = decoder FusedBatchNorm (Bessel-corrected EMA variance) */
} mliis_bn_info;

const char* mliis_last_error(void);
const char* mliis_version(void);

/* ---- context ------------------------------------------------------------------------------
 * Replaces: EfficientLab.__init__/build_model graph construction (models/efficientlab.py:23-119)
 * and tf.Session() (run_metasegnet.py:109).  device < 0 builds a table-only ctx (no CUDA call is
 * made; used to query the variable tables on a machine without a GPU). */
int mliis_ctx_create(const mliis_config* cfg, int device, mliis_ctx** out);
int mliis_ctx_destroy(mliis_ctx* ctx);

int64_t mliis_num_params(const mliis_ctx* ctx);        /* P = 2 071 714 for the canonical config */
int32_t mliis_num_param_tensors(const mliis_ctx* ctx); /* 169 */
int32_t mliis_num_bn_layers(const mliis_ctx* ctx);     /* 39  */
int32_t mliis_num_bn_channels(const mliis_ctx* ctx);   /* 8752 */
int32_t mliis_num_dc_blocks(const mliis_ctx* ctx);     /* 6 blocks with drop-connect */
int mliis_param_table(const mliis_ctx* ctx, mliis_param_info* out, int32_t capacity);
int mliis_bn_table(const mliis_ctx* ctx, mliis_bn_info* out, int32_t capacity);
/* bytes of device workspace ONE slot needs (activations, gradients, scratch). */
int64_t mliis_workspace_bytes(const mliis_ctx* ctx);
/* floats in one slot's state vector: [theta | moving_mean n_bn | moving_variance n_bn | adam_v | beta1_power,
 * beta2_power, 2 pad], where theta / adam_v are mliis_theta_floats() long: the P parameters laid out with the
 * L2-regularised tensors first and every tensor 16-byte aligned (offsets: mliis_param_table). */
int64_t mliis_state_floats(const mliis_ctx* ctx);
int64_t mliis_theta_floats(const mliis_ctx* ctx);

/* Attach caller-owned device memory to a slot.  state: mliis_state_floats() floats;
 * workspace: mliis_workspace_bytes() bytes, 256-byte aligned. */
int mliis_slot_bind(mliis_ctx* ctx, int32_t slot, float* dev_state, void* dev_workspace);

/* ---- state save / restore -----------------------------------------------------------------
 * Replaces VariableState.export_variables/import_variables round trips through host numpy
 * (meta_learners/variables.py:70-80; reptile.py:258, :293, :102, :123).  Device-to-device. */
int mliis_state_copy(mliis_ctx* ctx, float* dev_dst, const float* dev_src, int32_t what, void* stream);
enum { MLIIS_STATE_TRAINABLES = 1, MLIIS_STATE_BN = 2, MLIIS_STATE_OPT = 4, MLIIS_STATE_ALL = 7 };

/* ---- one inner step: sess.run(minimize_op, feed_dict) ---------------------------------------
 * Replaces reptile.py:115-121 / :269-279 / :639-643 -> models/efficientlab.py:315-317
 * (forward, loss, backward, BN moving-average UPDATE_OPS, optimizer apply).
 *   dev_images  [n_pool,H,W,3] fp32 in [0,255]; dev_labels [n_pool,H,W,2] fp32
 *   dev_index   [B] int32: rows of the pool forming this mini-batch (NULL = 0..B-1)
 *   dev_dc_mask [n_dc_blocks*B] fp32 {0,1} drop-connect binary_tensor (NULL = all keep)
 *   dev_drop_mask [B*h*w*D] fp32 {0,1} final-layer dropout keep mask (NULL: rate must be 0,
 *                 or a counter-based device RNG keyed by `seed` is used when rate > 0)
 *   pre_decay_rate: reptile.py:112-113 pre_step_op (var *= rate); 1 = none
 *   dev_loss_out [1] fp32 (may be NULL). */
typedef struct mliis_step_args {
  const float*   dev_images;
  const float*   dev_labels;
  const int32_t* dev_index;
  int32_t        batch;
  float          lr;
  float          pre_decay_rate;
  const float*   dev_dc_mask;
  const float*   dev_drop_mask;
  uint64_t       seed;
  float*         dev_loss_out;
  const uint64_t* dev_seed;        /* optional device scalar added to `seed` when the dropout mask is drawn: lets a
                                      captured CUDA graph draw fresh masks on every replay (NULL = seed only) */
} mliis_step_args;
/* After mliis_kernel_group(n, stride) (below) the step runs for the n slots slot .. slot+n-1 in lockstep - one launch per
 * kernel; every dev_* pointer of args is the first slot's, slot k uses pointer + k*stride (task-batched meta-training). */
int mliis_train_step(mliis_ctx* ctx, int32_t slot, const mliis_step_args* args, void* stream);

/* The pieces of a step, exposed for parity tests (same buffers as mliis_train_step). */
int mliis_forward(mliis_ctx* ctx, int32_t slot, const float* dev_images, const int32_t* dev_index,
                  int32_t batch, int32_t training, const float* dev_dc_mask, const float* dev_drop_mask,
                  uint64_t seed, float* dev_logits_out /* [B,H,W,2] or NULL */, void* stream);
int mliis_loss_backward(mliis_ctx* ctx, int32_t slot, const float* dev_labels, const int32_t* dev_index,
                        int32_t batch, float* dev_grads_out /* [P] or NULL */, float* dev_loss_out, void* stream);
int mliis_optimizer_step(mliis_ctx* ctx, int32_t slot, float lr, float pre_decay_rate, void* stream);
/* Data-parallel training (joint_train.py): overwrite the slot's flat gradient buffer, e.g. with the
 * NCCL all-reduced mean of every rank's mliis_loss_backward(dev_grads_out), before mliis_optimizer_step. */
int mliis_set_grads(mliis_ctx* ctx, int32_t slot, const float* dev_grads /* [mliis_theta_floats] */, void* stream);

/* ---- multi-class head (n_classes > 1; joint_train.py:295-343, efficientlab.py:294-327, :369-396) --------------
 * Labels are sparse: one foreground class id per pool example (1..n_classes; 0 is the background channel) and a
 * binary mask.  The reference's dense [H,W,1001] one-hot label / logit / probability tensors are never built.
 * mliis_forward(dev_logits_out) returns the LOW-resolution head output [B, H/4, W/4, n_classes+1] in this mode. */
int mliis_set_class_ids(mliis_ctx* ctx, int32_t slot, const int32_t* dev_class_ids /* [n_pool], stays referenced */);
/* compute_iou_metric / iou_callback (joint_train.py:248-269): eval-mode forward; class_map = argmax channel where its
 * probability > 0.5 else -1 (the one-hot content of float(p > 0.5)); inter/union = integer counts over all channels */
int mliis_predict_classes(mliis_ctx* ctx, int32_t slot, const float* dev_images, const float* dev_masks,
                          const int32_t* dev_index, int32_t batch, int32_t* dev_class_map_out /* [B,H,W] or NULL */,
                          uint32_t* dev_inter_out, uint32_t* dev_union_out, void* stream);

/* ---- predictions + IoU: sess.run(predictions, {input_ph, is_training_ph: False}) -------------
 * Replaces reptile.py:482-524 (_test_predictions), models/efficientlab.py:291-292 (threshold) and
 * the integer part of Gecko._iou (reptile.py:526-549).  BN uses moving statistics.
 *   dev_pred_out  [B,H,W,2] fp32 {0,1} (the reference's `predictions` tensor) or NULL
 *   dev_inter_out / dev_union_out [B] uint32 counts on channel 1 (need dev_labels) or NULL. */
int mliis_predict(mliis_ctx* ctx, int32_t slot, const float* dev_images, const float* dev_labels,
                  const int32_t* dev_index, int32_t batch, float* dev_pred_out, float* dev_logits_out,
                  uint32_t* dev_inter_out, uint32_t* dev_union_out, void* stream);

/* ---- whole task: Gecko._evaluate (reptile.py:235-294) on device ------------------------------
 * state <- init_state (full state incl. optimizer slots, like _full_state export/import), then
 * n_steps inner steps with mini-batches dev_batch_index[n_steps*batch] drawn from the task pool,
 * then transductive prediction + IoU counts on the query rows dev_query_index[n_query]. */
typedef struct mliis_task_args {
  const float*   dev_init_state;   /* mliis_state_floats() floats, not modified              */
  const float*   dev_images;       /* task pool [n_pool,H,W,3]                               */
  const float*   dev_labels;       /* task pool [n_pool,H,W,2]                               */
  const int32_t* dev_batch_index;  /* [n_steps*batch]                                        */
  const float*   dev_lr;           /* [n_steps] per-step learning rate (lr_scheduler)        */
  int32_t        n_steps;
  int32_t        batch;
  const int32_t* dev_query_index;  /* [n_query]                                              */
  int32_t        n_query;
  const float*   dev_dc_mask;      /* [n_steps*n_dc_blocks*batch] or NULL                    */
  uint64_t       seed;
  float          pre_decay_rate;
  uint32_t*      dev_inter_out;    /* [n_query]                                              */
  uint32_t*      dev_union_out;    /* [n_query]                                              */
  float*         dev_loss_out;     /* [n_steps] or NULL                                      */
  const uint64_t* dev_seed;        /* optional device scalar: step t draws its final-layer dropout mask from
                                      *dev_seed + seed + t, read at RUN time (graph replays see the staged value) */
  /* Task-batched execution: n_group consecutive slots (slot, slot+1, ...) adapt their
   * tasks in LOCKSTEP - every kernel of the step is launched once with the task slot as a grid dimension, so the
   * 14x14 / 28x28 layers of n_group tasks fill the SMs together.  0 or 1 = a single slot.  All per-slot memory must
   * share one layout at a uniform stride: the state and workspace buffers bound to slot k (mliis_slot_bind) and
   * every dev_* pointer of this struct except dev_init_state are the FIRST slot's; slot k uses
   * pointer + k * group_stride_bytes (a multiple of 256).  Results equal n_group single-slot calls to fp32 rounding
   * (the reduction partials are sized by the launch: deterministic for a given n_group); with MLIIS_GROUP_CANONICAL=1 in
   * the environment the single-slot partition is kept and the results are bit-identical. */
  int32_t        n_group;
  int64_t        group_stride_bytes;
} mliis_task_args;
int mliis_adapt_eval_task(mliis_ctx* ctx, int32_t slot, const mliis_task_args* args, void* stream);

/* The same task as ONE CUDA graph: capture once per slot with pointers that stay valid (per-slot staging
 * buffers for the pool, index lists, learning rates and outputs), then replay per task.  ~2300 kernel
 * launches collapse into one graph launch; slots replay concurrently on their own streams.  `stream` must
 * be a non-default stream.  With final-layer dropout pass mliis_task_args.dev_seed (a staged device scalar): the
 * host `seed` is a captured constant, the device scalar is read by every replay. */
int mliis_task_graph_capture(mliis_ctx* ctx, int32_t slot, const mliis_task_args* args, void* stream);
int mliis_task_graph_launch(mliis_ctx* ctx, int32_t slot, void* stream);

/* Kernels launched by this library in this process (graph replays count their captured kernels). */
uint64_t mliis_launch_count(void);

/* CRC-32C (Castagnoli) of a HOST buffer, continuing from `crc` (0 to start): the checksum of the TF V2 checkpoint
 * bundle entries and of the TFRecord framing (mliis_b200/checkpoint.py, tfrecord.py; the reference gets it from
 * TensorFlow's C++ runtime through tf.train.Saver / tf.data, run_metasegnet.py:125-133, data/input_fn.py:68-110).
 * Host code only: works without a GPU. */
uint32_t mliis_crc32c(const void* data, uint64_t n_bytes, uint32_t crc);

/* ---- meta-update (meta_learners/variables.py:9-45; reptile.py:122-125, :644-647) --------------
 * delta_sum += (theta_a - theta_b)             [Reptile: a = adapted, b = old; FOMAML: a = theta_T, b = theta_{T-1}]
 * theta     += scale * delta_sum               [scale = meta_step_size / meta_batch_size]
 * All buffers are flat [P] fp32 in the engine's parameter order. */
int mliis_delta_accumulate(mliis_ctx* ctx, float* dev_delta_sum, const float* dev_theta_a,
                           const float* dev_theta_b, int32_t first /* 1: overwrite */, void* stream);
int mliis_meta_apply(mliis_ctx* ctx, float* dev_theta, const float* dev_delta_sum, float scale, void* stream);

/* ---- the ONE exchange step of a meta-update when tasks are sharded over GPUs ----------------
 * Replaces the host-side np.mean over per-task variable lists (meta_learners/variables.py:16-23 called at
 * reptile.py:124, :646) across processes.  The exchanged buffer is
 *   [ sum of this rank's task deltas (mliis_theta_floats) | sum of its slots' BN moving statistics (2*n_bn) |
 *     number of contributing slots | 3 pad ]                                  = mliis_meta_buffer_floats() floats.
 * mliis_meta_reduce builds it from n_rows per-slot delta sums (rows row_stride floats apart) and the BN statistics of
 * slots first_slot.. ; mliis_allreduce_delta is ONE ncclAllReduce(sum) over NVLink (a no-op without a communicator);
 * mliis_meta_finish applies theta += scale * delta_sum and writes the contributor-averaged BN statistics into n_slots
 * slots.  Adam slots stay rank-local (documented semantics of SURVEY 8e; exact for --sgd).
 * The communicator lives in the ctx: rank 0 calls mliis_comm_unique_id, the host broadcasts the 128 bytes over its own
 * control plane (torch.distributed, MPI, a file), every rank calls mliis_comm_init.  NCCL is bound with dlopen. */
int64_t mliis_meta_buffer_floats(const mliis_ctx* ctx);
int mliis_comm_unique_id(uint8_t* out_id_128_bytes);
int mliis_comm_init(mliis_ctx* ctx, const uint8_t* id_128_bytes, int32_t rank, int32_t world);
int mliis_comm_destroy(mliis_ctx* ctx);
int mliis_allreduce_delta(mliis_ctx* ctx, float* dev_buf, int64_t count, void* stream);
int mliis_meta_reduce(mliis_ctx* ctx, float* dev_buf, const float* dev_delta_rows, int64_t row_stride_floats,
                      int32_t first_slot, int32_t n_rows, void* stream);
int mliis_meta_finish(mliis_ctx* ctx, float* dev_theta, const float* dev_buf, float scale, int32_t first_slot,
                      int32_t n_slots, void* stream);

/* ---- per-kernel entry points (unit tests / micro-benchmarks) --------------------------------- */
int mliis_dwconv_fwd(const float* dev_x, const float* dev_w, float* dev_y, int32_t B, int32_t H, int32_t W,
                     int32_t C, int32_t k, int32_t stride, const float* dev_bn_a, const float* dev_bn_b,
                     void* stream);
int mliis_gemm_nn(const float* dev_a, const float* dev_w, float* dev_c, int32_t M, int32_t K, int32_t N,
                  int32_t mode, void* stream);
int mliis_conv3x3_fwd(const float* dev_x, const float* dev_w, const float* dev_bias, float* dev_y, int32_t B,
                      int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t dilation, int32_t mode,
                      void* stream);
/* The tcgen05 / TMA / TMEM contraction alone (no allocation, no synchronisation): re-lay-out the HWIO kernel
 * once with mliis_tc_prep_weights (dev_wt: taps*Cin*Cout floats, twice that for MLIIS_GEMM_TF32X3), then run
 * y[B,H,W,Cout] = conv(x[B,H,W,Cin]) + bias with taps = 9 (3x3, SAME, dilation) or taps = 1 (1x1 / GEMM). */
int mliis_tc_prep_weights(const float* dev_w, float* dev_wt, int32_t taps, int32_t Cin, int32_t Cout, int32_t dgrad,
                          int32_t mode, void* stream);
int mliis_tc_conv(const float* dev_x, const float* dev_wt, const float* dev_bias, float* dev_y, int32_t B, int32_t H,
                  int32_t W, int32_t Cin, int32_t Cout, int32_t taps, int32_t dilation, int32_t mode, void* stream);
/* MBConv project conv with its fused prologue (efficientnet_model.py:225-232, :266, :271-273):
 * y[B*HW, Cout] = (swish(bn_a[c] * x[m, c] + bn_b[c]) * gate[image(m), c]) * W   (gate may be null; dev_wt from
 * mliis_tc_prep_weights with taps = 1).  The activated / gated tensor is never written to memory. */
int mliis_tc_project_conv(const float* dev_x, const float* dev_wt, const float* dev_bn_a, const float* dev_bn_b,
                          const float* dev_gate /* [B, Cin] or null */, float* dev_y, int32_t B, int32_t HW, int32_t Cin,
                          int32_t Cout, int32_t mode, void* stream);
/* Weight gradient on the tensor cores: dev_dw[taps*Cin, Cout] = sum over pixels of a[pixel+tap, :]^T g[pixel, :]
 * (a: [B,H,W,Cin], g: [B,H,W,Cout]; taps = 9 for a 3x3 SAME conv with `dilation`, 1 for a 1x1 conv). */
int mliis_tc_wgrad(const float* dev_a, const float* dev_g, float* dev_dw, int32_t B, int32_t H, int32_t W, int32_t Cin,
                   int32_t Cout, int32_t taps, int32_t dilation, int32_t mode, void* stream);
int mliis_bilinear_fwd(const float* dev_x, float* dev_y, int32_t B, int32_t Hin, int32_t Win, int32_t Hout,
                       int32_t Wout, int32_t C, void* stream);
int mliis_adam_step(float* dev_theta, float* dev_v, const float* dev_grad, int64_t n, int64_t n_l2, float lr,
                    float beta2_power, float l2_coef, void* stream);

/* Task-batched per-kernel calls: after mliis_kernel_group(n, stride) every per-kernel entry point called by this thread
 * launches ONCE for n slot copies laid out `stride` bytes apart (all pointer arguments are slot 0's; n = 1 resets). */
int mliis_kernel_group(int32_t n_group, int64_t group_stride_bytes);
/* floats of scratch that is enough for any of the entry points below on a [B,H,W,C] tensor */
int64_t mliis_kernel_scratch_floats(int32_t B, int32_t H, int32_t W, int32_t C);
/* DepthwiseConv2dNative backward (efficientnet_model.py:190-196): dx and dw of y = dw(swish(bn_a*x+bn_b)) given dy */
int mliis_dwconv_bwd(const float* dev_x, const float* dev_bn_a, const float* dev_bn_b, const float* dev_w,
                     const float* dev_dy, float* dev_dx, float* dev_dw, float* dev_scratch, int32_t B, int32_t H, int32_t W,
                     int32_t C, int32_t k, int32_t stride, void* stream);
/* train-mode BatchNorm bookkeeping (models/efficientnet/utils.py:111-134; efficientlab.py:190): statistics of x [M,C]
 * -> dev_stats = [mean | rstd | a = gamma*rstd | b = beta - mean*a], EMA of the moving statistics (fused=1: Bessel) */
int mliis_bn_stats_fwd(const float* dev_x, const float* dev_gamma, const float* dev_beta, float* dev_moving_mean,
                       float* dev_moving_var, float* dev_stats, float* dev_scratch, int32_t M, int32_t C, int32_t fused,
                       void* stream);
/* backward of swish(BN(x)) at the MBConv expand / stem sites: dgamma, dbeta, dx */
int mliis_bn_swish_bwd(const float* dev_x, const float* dev_g, float* dev_dx, const float* dev_stats,
                       const float* dev_gamma, float* dev_dgamma, float* dev_dbeta, float* dev_scratch, int32_t M,
                       int32_t C, void* stream);
/* squeeze-excite forward (efficientnet_model.py:238-251): gate [B,C] from the pre-BN depthwise output x [B,HW,C] */
int mliis_se_fwd(const float* dev_x, const float* dev_bn_a, const float* dev_bn_b, const float* dev_w1,
                 const float* dev_b1, const float* dev_w2, const float* dev_b2, float* dev_pool, float* dev_hidpre,
                 float* dev_gate, float* dev_scratch, int32_t B, int32_t HW, int32_t C, int32_t Cr, void* stream);
/* fused binary-head loss (efficientlab.py:294-327, :385-396): upsample, softmax-CE - ln(dice), IoU sums, d loss / d logits */
int mliis_softmax_ce_iou(const float* dev_z_lo, const float* dev_labels, float* dev_p1, float* dev_dz_hi,
                         float* dev_scratch, float* dev_loss_out, int32_t B, int32_t h, int32_t w, int32_t H, int32_t W,
                         int32_t dice, float label_smoothing, void* stream);
/* conv2d_2 of a residual skip decoder module with the image-pooling branch folded into a per-image, per-border-class
 * bias (efficientlab.py:192-197, :220-224): operand over the first Cin of the Cs = Cin + Cp input channels */
int mliis_tc_prep_weights_sub(const float* dev_w, float* dev_wt, int32_t taps, int32_t Cin, int32_t Cs, int32_t Cout,
                              int32_t dgrad, int32_t mode, void* stream);
int mliis_rsd_conv2_fwd(const float* dev_x, int32_t ldx, const float* dev_pooled, const float* dev_w_hwio,
                        const float* dev_wt, const float* dev_bias, float* dev_bias9 /* [B,9,Cout] scratch */,
                        float* dev_y, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cp, int32_t Cout, int32_t mode,
                        void* stream);

/* Measured tensor-pipe peak of the mode the dense kernels run in (tcgen05.mma kind::tf32, cta_group::1, M=128 N=256 K=8,
 * operands resident in shared memory, no loads, every SM busy): the roofline denominator beside the bf16 cuBLAS one. */
int mliis_tc_peak_tf32(int32_t iters, double* tflops_out, void* stream);

/* Issue-rate microbenchmark of the MMA patterns the convolutions use: SM clocks per k-step (K = 8 of TF32) on one SM's
 * tensor pipe with every SM busy.  pattern 0: one M=128 x N=n MMA; 1: the 3xTF32 pair (N=2n with A_hi, N=n with A_lo);
 * 2: pattern 1 alternating between two accumulators; 3: three N=n MMAs.  shift_rows: A start shifted by 128-byte rows. */
int mliis_tc_mma_rate(int32_t iters, int32_t n, int32_t pattern, int32_t shift_rows, double* clk_per_kstep_out,
                      void* stream);

/* ---- debugging: copy a named activation / gradient buffer of a slot (tests only) -------------- */
int mliis_debug_buffer(mliis_ctx* ctx, int32_t slot, const char* name, const float** dev_ptr,
                       int64_t* rows_per_image, int32_t* channels, int32_t* ld);

#ifdef __cplusplus
}
#endif
#endif /* MLIIS_B200_H_ */
