/* msf_b200.h — C ABI of the B200-native HybridFusion hot path.
 *
 * The reference (Rutgers-ECE-MML4SS/multimodal-sensor-fusion-with-attention-RajeevAtla)
 * is pure Python/PyTorch and has no FFI of its own; this header is the boundary
 * a maintainer binds (ctypes stub in INTEGRATION.md) to move the arithmetic of
 *
 *   src/fusion.py:331-479     HybridFusion.forward / compute_adaptive_weights
 *   src/attention.py:68-146   CrossModalAttention.forward (q_len = k_len = 1 inside HybridFusion)
 *   src/train.py:185-186,310  CrossEntropyLoss(label_smoothing)
 *   src/train.py:378-382,416-430  AdamW step + global-norm clip
 *   src/eval.py:89-90         softmax -> max -> (confidence, prediction)
 *   src/uncertainty.py:84-171,218-241  ECE / MCE / reliability-diagram binning
 *
 * onto sm_100a kernels.  Conventions:
 *   - plain C, POD structs, raw DEVICE pointers, sizes, a cudaStream_t passed as void*;
 *   - every entry point returns 0 on success or a negative MSF_E_* code and never
 *     throws; msf_last_error() returns a thread-local message for the last failure;
 *   - the library BORROWS all buffers for the duration of the call (PyTorch owns
 *     them); it allocates nothing persistent except small per-device caches
 *     (TMA descriptors) that it owns;
 *   - all work is enqueued on the given stream; no entry point synchronises.
 *
 * Parameter ("master") arena layout, fp32, in the reference's registration
 * order (src/fusion.py:291-328) so that it equals
 * torch.nn.utils.parameters_to_vector(model.parameters()):
 *
 *   for m in modalities:            projections.m.0.weight (H x D_m), .bias (H)
 *   for q in modalities, k != q:    attention_modules.q_to_k.{query,key,value,out}_proj
 *                                   .weight (H x H), .bias (H)        [4 x (H*H + H)]
 *   for m in modalities:            gating_layers.m.weight (1 x H), .bias (1)
 *   classifier.0.weight (H x H), .bias (H); classifier.3.weight (C x H), .bias (C)
 *
 * A pair module deleted from the ModuleDict (src/fusion.py:388-389) keeps its
 * slot; clear its bit in msf_fusion_shape.pair_present.
 */
#ifndef MSF_B200_H_
#define MSF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSF_ABI_VERSION 4
#define MSF_MAX_MODALITIES 8

enum {
  MSF_OK = 0,
  MSF_E_INVALID = -1,      /* bad argument / unsupported shape */
  MSF_E_CUDA = -2,         /* a CUDA runtime/driver call failed */
  MSF_E_WORKSPACE = -3,    /* workspace too small */
  MSF_E_UNSUPPORTED = -4   /* shape not eligible for the requested precision */
};

/* Arithmetic the path computes in. */
enum {
  MSF_PREC_F32 = 0,   /* fp32 FFMA kernels: parity mode, max-abs <= 1e-5 vs the fp32 oracle */
  MSF_PREC_BF16 = 1   /* bf16 operands, fp32 accumulate on tcgen05/TMEM: <= 1e-2 */
};

typedef struct msf_fusion_shape {
  int32_t num_modalities;               /* M, 1..MSF_MAX_MODALITIES            */
  int32_t hidden;                       /* H  (fusion.py hidden_dim)           */
  int32_t num_heads;                    /* heads; H % heads == 0               */
  int32_t num_classes;                  /* C                                   */
  int32_t in_dims[MSF_MAX_MODALITIES];  /* D_m                                 */
  uint64_t pair_present;                /* bit (q*M + k): attention_modules[q_to_k] exists */
} msf_fusion_shape;

/* One forward or backward invocation over a batch of windows. */
typedef struct msf_fusion_call {
  int32_t batch;        /* B windows                                          */
  int32_t precision;    /* MSF_PREC_*                                         */
  int32_t training;     /* nonzero: dropout active with probability dropout_p */
  float dropout_p;
  uint64_t seed;        /* Philox4x32-10 key; identical (seed, offset) in     */
  uint64_t offset;      /* forward and backward regenerate identical masks    */
  const uint64_t* rng_state;    /* optional DEVICE {seed, offset[, step]}: overrides seed/offset so a
                                   captured CUDA graph draws fresh masks on every replay */
  const float* params;          /* master arena, fp32                          */
  const void* params_bf16;      /* compute arena made by msf_fusion_pack_bf16 (BF16 precision) */
  const float* x[MSF_MAX_MODALITIES];  /* (B, D_m) row-major encoder outputs: fp32, or bf16 when x_bf16 != 0 */
  const float* mask;            /* (B, M) fp32 availability, or NULL = all ones */
  void* workspace;              /* msf_fusion_workspace_bytes(); forward fills it, backward reads it */
  size_t workspace_bytes;
  /* forward outputs */
  float* logits;                /* (B, C)                                      */
  float* fusion_weights;        /* (B, M) or NULL                              */
  float* attn_gates;            /* (M*(M-1), B, heads): attention_maps values, q-major pair order, or NULL */
  /* backward inputs / outputs */
  const float* grad_logits;     /* (B, C)                                      */
  float* grad_params;           /* master layout, fully overwritten (dead q/k slots = 0) */
  float* grad_x[MSF_MAX_MODALITIES];   /* (B, D_m) or NULL                     */
  double* grad_sq;              /* optional, msf_fusion_train_pass (fused BF16 path) only: receives the sum of
                                   squares of grad_params as written (weight matrices: summed by the weight-gradient
                                   GEMM's epilogue; bias / gating slots: by a small kernel beside it); see
                                   MSF_OPT_NORM_GIVEN */
  /* Optional per-modality LayerNorm between the encoders and the fusion model (src/train.py:170-171,267-268:
   * nn.LayerNorm(D_m) on every encoder output): x[m] are then the RAW encoder outputs and the projection kernel
   * normalises each row in its input phase, y = (x - mean) * rstd * ln_weight + ln_bias (biased variance), before
   * mask and dropout.  Both NULL for a modality = no LayerNorm there.  Only where
   * msf_fusion_layer_norm_fused() says 1 (otherwise call msf_layer_norm_forward first and leave these NULL).
   * grad_x[m] is then the gradient with respect to the NORMALISED rows: msf_layer_norm_backward turns it into the
   * gradient with respect to x[m] and the gradients of ln_weight / ln_bias. */
  const float* ln_weight[MSF_MAX_MODALITIES];
  const float* ln_bias[MSF_MAX_MODALITIES];
  float ln_eps;
  /* nonzero: x[m] point to bf16 rows (B, D_m), 2 bytes per element.  The tensor-core path rounds the inputs to bf16
   * anyway (xt = bf16(dropout(x * mask))), so features that cross PCIe as bf16 halve the host->device bytes of a
   * step; only where the fused projection kernel runs (BF16 precision, in_dims % 64 == 0, <= 256), else
   * MSF_E_UNSUPPORTED.  grad_x stays fp32. */
  int32_t x_bf16;
} msf_fusion_call;

/* ---- per-modality LayerNorm (src/train.py:170-171,267-268) ------------------ */
/* 1 if msf_fusion_forward / _train_pass / _infer_pass apply msf_fusion_call.ln_* inside the projection kernel for
 * this shape and precision (tensor-core path, in_dims % 64 == 0), else 0. */
int msf_fusion_layer_norm_fused(const msf_fusion_shape* shape, int32_t precision);
/* y[rows, dim] = (x - mean) * rstd * gamma + beta per row (gamma / beta may be NULL), fp32, dim <= 512. */
int msf_layer_norm_forward(const float* x, const float* gamma, const float* beta, float* y, int64_t rows, int32_t dim,
                           float eps, void* stream);
/* dx (may be NULL), dgamma, dbeta (may be NULL; overwritten) from dy and the forward input x. */
int msf_layer_norm_backward(const float* x, const float* gamma, const float* dy, float* dx, float* dgamma, float* dbeta,
                            int64_t rows, int32_t dim, float eps, void* stream);

/* ---- library ------------------------------------------------------------ */
/* Concurrency: every call enqueues on the caller's stream and returns.  The optimizer kernels
 * (msf_fusion_optimizer_step*, msf_dp*_optimizer_step*) and msf_cross_entropy keep their grid-wide ticket / barrier words
 * in process-global device variables: launches of these entry points must not run CONCURRENTLY on one device (two
 * streams, two engines in one process) — serialise them on one stream or with events.  One process per GPU, as
 * the drop-in and the engine use the library, satisfies this by construction. */
int msf_abi_version(void);
/* sizeof(msf_fusion_shape), sizeof(msf_fusion_call) as compiled, for binding self-checks */
int msf_struct_sizes(int32_t* shape_bytes, int32_t* call_bytes);
const char* msf_last_error(void);
/* kernels launched by this library so far in this process (host-side counter) */
uint64_t msf_launch_count(void);
/* Per-launch timing of the tensor-core GEMM launches with CUDA events on the launching stream.
 * enable(1) clears earlier records and starts recording, enable(0) stops; enable(R > 1) additionally makes the
 * chained pair-GEMM launches issue their (idempotent) kernel R times back to back inside one event bracket, so
 * the average includes the prologue overlap of programmatic dependent launches and excludes the event records.
 * report() synchronises the device and writes one line per launch label:
 * "label\tlaunches\ttotal_ms\ttotal_flops\n".
 * Must not be enabled while the stream is being captured into a CUDA graph. */
int msf_prof_enable(int32_t on);
int msf_prof_report(char* buf, size_t cap);
/* n asynchronous copies (cudaMemcpyDefault: host-pinned or device pointers) on one stream in one call:
 * the per-step host->device staging of a batch (M feature tensors, mask, labels) without per-tensor
 * framework dispatch on the host. */
int msf_memcpy_batch(void* const* dst, const void* const* src, const size_t* bytes, int32_t n, void* stream);
/* sm count / compute capability of the current device; fails unless cc >= 10.0 */
int msf_device_check(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ---- HybridFusion (src/fusion.py:248-479) -------------------------------- */
int msf_fusion_param_count(const msf_fusion_shape* shape, int64_t* count);
/* element offset of a tensor in the master arena.
 * kind: 0 proj.weight 1 proj.bias (idx = m); 2..9 pair {q,k,v,o}x{weight,bias}
 * (idx = q*M + k); 10 gate.weight 11 gate.bias (idx = m); 12..15 classifier
 * {0.weight, 0.bias, 3.weight, 3.bias}. */
int msf_fusion_param_offset(const msf_fusion_shape* shape, int32_t kind, int32_t idx, int64_t* offset);
int msf_fusion_workspace_bytes(const msf_fusion_shape* shape, int32_t batch, int32_t precision,
                               size_t* bytes);
/* bytes of the bf16 compute arena and the conversion master -> compute layout */
int msf_fusion_bf16_arena_bytes(const msf_fusion_shape* shape, size_t* bytes);
int msf_fusion_pack_bf16(const msf_fusion_shape* shape, const float* params, void* params_bf16,
                         void* stream);
int msf_fusion_forward(const msf_fusion_shape* shape, const msf_fusion_call* call, void* stream);
int msf_fusion_backward(const msf_fusion_shape* shape, const msf_fusion_call* call, void* stream);
/* One training pass of src/train.py:302-324 around the fusion model: forward, CrossEntropyLoss(label_smoothing)
 * (mean over `batch`; d logits scaled by grad_scale, normally 1/(batch*world)) and backward, as one enqueue.
 * call->logits receives the logits, call->grad_params the gradients, loss_out[0] the mean loss (may be NULL),
 * row_loss (batch) is scratch.  With MSF_PREC_BF16 and num_classes <= 32, hidden <= 256 everything between
 * the aggregated modality tokens and their gradients (gating softmax, classifier, loss, their backward) is ONE
 * kernel; otherwise this is msf_fusion_forward + msf_cross_entropy + msf_fusion_backward and
 * grad_logits_scratch (batch x C) is required.
 * flags: MSF_TRAIN_DEAD_SLOTS_ZERO = the caller guarantees that the dead query/key slots of grad_params
 * already hold zeros (a persistent gradient arena that only this entry point writes): they are then not
 * rewritten, and no pass over the whole 13 MB arena is needed. */
#define MSF_TRAIN_DEAD_SLOTS_ZERO 1
int msf_fusion_train_pass(const msf_fusion_shape* shape, const msf_fusion_call* call, const int64_t* labels,
                          float smoothing, float grad_scale, float* row_loss, float* loss_out,
                          float* grad_logits_scratch, int32_t flags, void* stream);
/* 1 if msf_fusion_train_pass runs this shape / precision through the fused path (one head kernel, call->grad_sq
 * honoured), 0 if it composes forward + cross-entropy + backward (call->grad_sq must then be NULL). */
int msf_fusion_train_pass_is_fused(const msf_fusion_shape* shape, int32_t precision);
/* Inference pass of src/eval.py:84-90: logits, then softmax -> max -> (confidence, prediction), with the
 * softmax fused into the classifier epilogue on the tensor-core path.
 * present_hint: 0, or the set of modalities (bit m) that call->mask marks present in EVERY row while all others
 * are absent in EVERY row — the missing-modality subset sweep of src/eval.py:342-404.  The result is the same as
 * without the hint; the projections of absent modalities, the pair GEMMs with an absent key (attention gate 0:
 * exactly the out_proj bias remains) and the rows of absent queries are not computed (tensor-core path). */
int msf_fusion_infer_pass(const msf_fusion_shape* shape, const msf_fusion_call* call, float* conf, int64_t* pred,
                          uint32_t present_hint, void* stream);
/* The subset sweep of src/eval.py:342-404 with every attention module's value_proj -> out_proj pair FOLDED into one
 * matrix: in inference (no attention dropout) the gate of a present key is 1 for every head, so
 * out_proj(value_proj(P_k)) = P_k (Wo Wv)^T + (Wo bv + bo).  Per present query one tcgen05 GEMM whose K-segments are
 * the present keys (accumulated in TMEM) with the mean epilogue, then the fused head kernel: half the pair FLOPs
 * of msf_fusion_infer_pass, same result within the BF16 tolerance.  Tensor-core path, shapes of the fused kernels.
 *   wov_bf16  [M*(M-1)][H][H] bf16, pair order of attn_gates: Wo_qk * Wv_qk, row-major [out][in]
 *   bias_sum  [M][H] fp32 for THIS subset: row q = sum_{k present, k != q} (Wo_qk bv_qk + bo_qk) + sum_{k absent} bo_qk
 *   present   bit m = modality m is present in every row (call->mask must be that uniform mask)
 *   flags     bit 0: the projections of the present modalities are still in the workspace from an earlier call on
 *             the same features (the sweep projects once); bit 1: compute the projections only */
#define MSF_FOLD_REUSE_PROJECTIONS 1
#define MSF_FOLD_PROJECTIONS_ONLY 2
int msf_fusion_infer_folded(const msf_fusion_shape* shape, const msf_fusion_call* call, const void* wov_bf16,
                            const float* bias_sum, uint32_t present, int32_t flags, float* conf, int64_t* pred,
                            void* stream);
/* Debugging aid: clock64 stamps (SM cycles) of the phases of CTA 0 in the last fused head-kernel launch
 * (P0 start/end, then acquire/finish of E1..E4, P5 start/end).  Synchronises the device. */
int msf_debug_head_stamps(int64_t* out16);
/* Same for the last chained pair-GEMM launch: wait-time accounting of CTA 0 (chain3_gemm.cu; compiled into the timeline build only,
 * MSF_E_UNSUPPORTED otherwise; MSF_CHAIN=v2: chain2_gemm.cu). */
int msf_debug_chain_stamps(int64_t* out32);   /* 32 entries (chain2_gemm.cu fills the first 16) */
/* Phase stamps of CTA 0 of the last input-projection launch (proj_gemm.cu). */
int msf_debug_proj_stamps(int64_t* out16);
/* HybridFusion.compute_adaptive_weights (src/fusion.py:429-479) stand-alone:
 * agg (M, B, H), gate_w (M, H), gate_b (M), mask (B, M) -> weights (B, M). */
int msf_adaptive_weights(const float* agg, const float* gate_w, const float* gate_b, const float* mask,
                         int64_t batch, int32_t num_modalities, int32_t hidden, float* weights,
                         void* stream);
/* Dropout multipliers (0 or 1/(1-p)) exactly as the kernels draw them, for
 * injecting into the oracle.  site: 0 input (sub = m, rows x D_m), 1 projection
 * (sub = m, rows x H), 2 attention weights (sub = q*M + k, rows x heads),
 * 3 classifier hidden (sub = 0, rows x H). */
int msf_dropout_mask(uint64_t seed, uint64_t offset, int32_t site, int32_t sub, int64_t rows,
                     int64_t cols, float p, float* out, void* stream);

/* gather/scatter between per-tensor storage and a flat arena.  `table` is a
 * DEVICE array of n_tensors {int64 ptr, int64 arena_offset, int64 numel}. */
int msf_arena_gather(const int64_t* table, int32_t n_tensors, int64_t total, float* arena, void* stream);
int msf_arena_scatter(const int64_t* table, int32_t n_tensors, int64_t total, const float* arena, void* stream);

/* ---- loss / confidence (src/train.py:185-186,310; src/eval.py:89-90) ------ */
/* loss_out[0] = mean CE with label smoothing; grad_logits = d loss / d logits * grad_scale.
 * row_loss: (B) scratch. labels int64. */
int msf_cross_entropy(const float* logits, const int64_t* labels, int64_t batch, int32_t classes,
                      float smoothing, float grad_scale, float* row_loss, float* loss_out,
                      float* grad_logits, void* stream);
/* Rows handed to msf_cross_entropy whose label was outside [0, classes) since the last reset (synchronises the
 * device).  Such rows never index the logits: they are treated as having no target class.  The fused head kernel
 * (msf_fusion_train_pass) compares class indices instead of indexing, so it cannot read out of bounds either;
 * FusionEngine.load_batch validates host labels before they are copied. */
int msf_bad_label_count(int64_t* count, int32_t reset);
int msf_softmax_conf_pred(const float* logits, int64_t batch, int32_t classes, float* conf,
                          int64_t* pred, void* stream);

/* ---- ECE / reliability binning (src/uncertainty.py:84-171,218-241) -------- */
/* Single pass. edges: HOST array of num_bins+1 ascending doubles (fp32
 * torch.linspace edges promoted, or np.linspace f64 edges). Bin i holds
 * edges[i] <= c < edges[i+1]; the last bin also takes c == edges[num_bins].
 * NaN / out-of-range confidences fall in no bin.  Accumulates INTO
 * count/correct (int64[num_bins]) and conf_sum_q32 (uint64[num_bins], sum of
 * round(c * 2^32)) — zero them first; integer adds make the result
 * order-independent, hence deterministic and shard-mergeable. */
int msf_ece_bin(const float* conf, const int64_t* pred, const int64_t* label, int64_t n,
                const double* edges, int32_t num_bins, int64_t* count, int64_t* correct,
                uint64_t* conf_sum_q32, void* stream);
/* Evaluation epilogue in one pass over the logits (src/eval.py:39-130, src/uncertainty.py:495-553): per window
 * softmax -> (confidence, first arg-max) and NLL = lse - logit[label]; ACCUMULATED (all integer, order-independent,
 * shards merge with one integer all-reduce) into
 *   confusion[classes * classes]  counts of (label, prediction): accuracy and macro-F1 follow from it
 *   scalars[3]                    {windows seen, sum of NLL in Q24 fixed point, windows whose label is outside [0, C)}
 *   bins[3 * num_bins]            ECE / reliability bins as msf_ece_bin: count, correct, confidence sum (Q32)
 * conf / pred (batch each) may be NULL.  Windows with an out-of-range label only count in scalars[0] and scalars[2]. */
int msf_eval_accumulate(const float* logits, const int64_t* labels, int64_t batch, int32_t classes,
                        const double* edges, int32_t num_bins, float* conf, int64_t* pred, int64_t* confusion,
                        int64_t* scalars, int64_t* bins, void* stream);

/* ---- optimizer (src/train.py:378-382,416-430) ------------------------------ */
/* sq_norm[0] += sum(g^2) (double); zero it first. */
int msf_grad_sq_norm(const float* grad, int64_t n, double* sq_norm, void* stream);
/* AdamW over a flat arena.  clip: if max_norm > 0 the gradient is scaled by
 * min(1, max_norm / (sqrt(sq_norm[0]) * grad_scale + 1e-6)) after grad_scale
 * (1/world_size under data parallelism).  step is 1-based. */
int msf_adamw_step(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                   int64_t step, float lr, float beta1, float beta2, float eps, float weight_decay,
                   float grad_scale, float max_norm, const double* sq_norm, void* stream);
/* Same, with the 1-based step read from device memory (train_state[2]) so the
 * update can live inside a replayed CUDA graph. */
int msf_adamw_step_dev(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                       const uint64_t* train_state, float lr, float beta1, float beta2, float eps,
                       float weight_decay, float grad_scale, float max_norm, const double* sq_norm,
                       void* stream);
/* Clip + AdamW over a HybridFusion master arena, aware of its layout: the gradient square-norm is
 * computed here (sq_norm is scratch, zeroed by the call) and the dead query/key projection slots —
 * whose gradient and Adam moments are identically zero — only receive the decoupled weight decay,
 * which is exactly what AdamW does to them (src/train.py:378-382).  step is read from train_state[2]. */
int msf_fusion_optimizer_step(const msf_fusion_shape* shape, float* params, const float* grad, float* exp_avg,
                              float* exp_avg_sq, const uint64_t* train_state, float lr, float beta1, float beta2,
                              float eps, float weight_decay, float grad_scale, float max_norm, double* sq_norm,
                              void* stream);
/* The same step fused with what follows it in a bf16 training loop: every live parameter is updated, its
 * bf16 copy (and transposed copy) in the compute arena `params_bf16` rewritten from registers, and — if
 * advance_state & 1 — train_state {seed, offset, step} moved on by the last CTA (offset += 1, step += 1),
 * replacing msf_fusion_optimizer_step + msf_fusion_pack_bf16 + msf_train_state_advance.
 * advance_state & MSF_OPT_NORM_GIVEN: sq_norm points at TWO doubles and sq_norm[1] already holds the sum of
 * squares of `grad` (msf_fusion_train_pass with call->grad_sq = sq_norm + 1 earlier on the same stream), so the
 * pass over the gradient arena and its grid-wide barrier disappear.  sq_norm[0] receives the value used. */
#define MSF_OPT_NORM_GIVEN 2
int msf_fusion_optimizer_step_packed(const msf_fusion_shape* shape, float* params, const float* grad,
                                     float* exp_avg, float* exp_avg_sq, uint64_t* train_state, float lr,
                                     float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                                     float max_norm, double* sq_norm, void* params_bf16, int32_t advance_state,
                                     void* stream);
/* Data-parallel variant: gradient reduction over NVLink peer memory fused with clip + AdamW (one
 * process per GPU).  grads[r] / stages[r] / reds[r] / sigs[r] are the peer-mapped (symmetric) gradient
 * arena, staging arena, reduced-gradient arena (all msf_fusion_param_count floats) and 64-word uint64
 * signal block (zero-initialised) of rank r, in rank order, as seen from THIS process.  Only stages,
 * reds and sigs of the peers are written; nothing is read across NVLink.  Every rank must call it once per step with the same
 * arguments; the call replaces all-reduce + msf_fusion_optimizer_step.  Each rank's gradient must be
 * pre-scaled so that the SUM over ranks is the wanted gradient (engine: 1/(B*world)). */
#define MSF_DP_MAX_RANKS 8
typedef struct msf_dp_comm {
  int32_t rank, world;
  const float* grads[MSF_DP_MAX_RANKS];
  float* stages[MSF_DP_MAX_RANKS];
  float* reds[MSF_DP_MAX_RANKS];
  uint64_t* sigs[MSF_DP_MAX_RANKS];
} msf_dp_comm;
int msf_dp_optimizer_step(const msf_fusion_shape* shape, const msf_dp_comm* comm, float* params, float* exp_avg,
                          float* exp_avg_sq, const uint64_t* train_state, float lr, float beta1, float beta2,
                          float eps, float weight_decay, float grad_scale, float max_norm, void* stream);
/* The same exchange with the update fused like msf_fusion_optimizer_step_packed: clip + AdamW from the
 * reduced arena, bf16 re-pack of `params_bf16` and (advance_state != 0) the train-state advance in one launch. */
int msf_dp_optimizer_step_packed(const msf_fusion_shape* shape, const msf_dp_comm* comm, float* params,
                                 float* exp_avg, float* exp_avg_sq, uint64_t* train_state, float lr, float beta1,
                                 float beta2, float eps, float weight_decay, float grad_scale, float max_norm,
                                 void* params_bf16, int32_t advance_state, void* stream);
/* Sharded ("owner computes") data-parallel step for the tensor-core path: the master arena is cut into optimizer
 * units (32 x 32 tiles of the weight matrices, 1024-element pieces of the vectors); live unit u belongs to rank
 * u % world.  Two kernels over NVLink peer memory:
 *   (1) every rank pushes its gradient values of the units it does not own into the owner's staging arena
 *       (stages[owner] + rank * stride + arena offset, stride = total rounded up to a multiple of 4 floats); after a cross-GPU flag barrier each owner sums its units over
 *       the ranks in fixed rank order, IN PLACE in `grad`, and publishes their square norm;
 *   (2) after the norm barrier each owner applies clip + AdamW to its units (master weights and Adam moments of a
 *       unit live on its owner only), pushes the bf16 copy and transposed copy of every updated weight tile into
 *       EVERY rank's compute arena and the updated vector slots (biases, gating layers: the forward kernels read
 *       them from the fp32 master) into every rank's master arena; dead query / key slots are decayed by every
 *       rank; the kernel ends when all ranks' pushes have landed.
 * The redundant full-arena AdamW of msf_dp_optimizer_step_packed becomes 1/world of it, and the reduced gradients
 * never travel: (world-1)/world x 4 bytes per live parameter cross NVLink twice per step, as before.
 * All pointer tables are peer-mapped (symmetric memory), indexed by rank; params == comm->params[comm->rank].
 * msf_dpz_owner_map fills owner[total] on the HOST: rank owning each master element's weights and moments,
 * -1 for replicated elements, -2 - rank for vector slots (moments on the owner, value valid everywhere). */
typedef struct msf_dpz_comm {
  int32_t rank, world;
  float* stages[MSF_DP_MAX_RANKS];       /* world x ((total + 3) / 4 * 4) floats each */
  void* arenas_bf16[MSF_DP_MAX_RANKS];   /* compute arenas (msf_fusion_pack_bf16 layout) */
  float* params[MSF_DP_MAX_RANKS];       /* master arenas */
  uint64_t* sigs[MSF_DP_MAX_RANKS];      /* 64 x uint64 signal blocks, zero-initialised */
  /* Optional NVLink multicast (NVLS) addresses, all three or none: the address that stands for the same offset in
   * EVERY rank's gradient arena (`grad` of each rank's call must be its own window of it), compute arena and master
   * arena.  With them the owner of a unit reads its gradient already summed over the ranks by the switch
   * (multimem.ld_reduce: nothing is pushed into `stages`, which may then be NULL) and delivers every updated
   * weight to all ranks with one multimem.st instead of one store per rank. */
  const float* mc_grad;
  void* mc_arena_bf16;
  float* mc_params;
} msf_dpz_comm;
int msf_dpz_optimizer_step_packed(const msf_fusion_shape* shape, const msf_dpz_comm* comm, float* params, float* grad,
                                  float* exp_avg, float* exp_avg_sq, uint64_t* train_state, float lr, float beta1,
                                  float beta2, float eps, float weight_decay, float grad_scale, float max_norm,
                                  int32_t advance_state, void* stream);
int msf_dpz_owner_map(const msf_fusion_shape* shape, int32_t world, int8_t* owner_host);
/* Gradient accumulation over micro-batches (config/base.yaml:75 accumulate_grad_batches, src/train.py:519-521):
 * acc[i] = (first ? 0 : acc[i]) + grad[i] over n floats; if both are given, *loss_acc = (first ? 0 : *loss_acc) +
 * loss_scale * *loss (the mean micro-batch loss).  Each micro-batch's pass scales its gradient by
 * 1 / (batch * micro_batches) (grad_scale of msf_fusion_train_pass), so the sum is the mean over the whole step. */
int msf_grad_accumulate(float* acc, const float* grad, int64_t n, float* loss_acc, const float* loss, float loss_scale,
                        int32_t first, void* stream);
/* train_state = DEVICE {seed, offset, step[, lr bits]}: offset += 1, step += 1.
 * Every optimizer entry point that takes a train_state accepts lr < 0: the learning rate is then the fp32 whose bits
 * are the low word of train_state[3] (a fourth uint64 the caller owns and updates between launches), so a captured
 * CUDA graph follows a learning-rate schedule (src/train.py:395-404) without being re-captured. */
int msf_train_state_advance(uint64_t* train_state, void* stream);

/* ---- dense layer (nn.Linear) on the fp32 FFMA path ------------------------- */
/* y[rows,out] = x[rows,in] . w[out,in]^T + b  (b may be NULL); relu != 0 applies max(.,0).
 * Used by the stand-alone CrossModalAttention / encoder projection modules
 * (src/attention.py:104-106,140; src/encoders.py:165). */
int msf_linear_forward(const float* x, const float* w, const float* b, float* y, int64_t rows,
                       int32_t in_dim, int32_t out_dim, int32_t relu, void* stream);
/* dx = dy . w (NULL to skip); dw = dy^T . x; db = colsum(dy) (NULL to skip).
 * If relu != 0, dy is first masked by y > 0 into dy_scratch (rows x out). */
int msf_linear_backward(const float* x, const float* w, const float* y, const float* dy,
                        float* dy_scratch, float* dx, float* dw, float* db, int64_t rows,
                        int32_t in_dim, int32_t out_dim, int32_t relu, void* stream);

/* ---- generic attention core (src/attention.py:108-139) --------------------- */
/* q (B,Lq,H), k/v (B,Lk,H) already projected; mask (B,Lk) or NULL (0 = masked key).
 * weights (B,heads,Lq,Lk) receives the post-dropout attention weights the
 * reference returns; out (B,Lq,H) the merged-head context before out_proj. */
int msf_attention_core_forward(const float* q, const float* k, const float* v, const float* mask,
                               int64_t batch, int32_t q_len, int32_t k_len, int32_t hidden, int32_t heads,
                               float dropout_p, int32_t training, uint64_t seed, uint64_t offset,
                               float* weights, float* out, void* stream);
/* scratch: (B,heads,Lq,Lk).  dk/dv are zeroed then accumulated. */
int msf_attention_core_backward(const float* q, const float* k, const float* v, const float* mask,
                                int64_t batch, int32_t q_len, int32_t k_len, int32_t hidden, int32_t heads,
                                float dropout_p, int32_t training, uint64_t seed, uint64_t offset,
                                const float* weights, const float* grad_out, float* scratch, float* dq,
                                float* dk, float* dv, void* stream);

/* ---- LSTM recurrence of SequenceEncoder (src/encoders.py:54-65,135-166) ------ */
/* nn.LSTM(F, H, num_layers=1, batch_first=True) forward over T steps with zero initial state, bf16 operands /
 * fp32 accumulate and state, for up to MSF_LSTM_MAX_SEQS encoders of the same batch, length and hidden size.
 * hidden <= 256: ONE persistent launch over all time steps (lstm_seq.cu: a cluster of hidden/64 CTAs per share of the
 * windows, its W_hh / W_ih columns resident in shared memory, h exchanged through the L2 with one cluster barrier
 * per step); otherwise one grouped tensor-core launch per time step.  Operand layouts (prepared once per model / batch by the caller):
 *   x_bf16  [T][B][64]        bf16, time-major, the F <= 64 features zero-padded to 64 columns
 *   w_hh    [H/64][4H][64]    bf16: row 4u+g = gate g (i,f,g,o) of hidden unit u, columns split into 64-wide k-blocks
 *   w_ih    [4H][64]          bf16: same row order, F columns zero-padded to 64
 *   bias    [4H]              fp32: bias_ih + bias_hh in the same row order
 *   h_a,h_b [H/64][B][64]     bf16 scratch; h_a holds h_0 (zeros);  cell B*H fp32: zeros on entry (c_0), scratch
 *                             afterwards (c_T in an implementation-defined order)
 *   h_out   [B][H]            fp32: h_T (what SequenceEncoder feeds to its projection)
 *   lengths [B] int32 or NULL  valid steps per window (1..T): the state of a window stops after its last valid step,
 *                             as with pack_padded_sequence (src/encoders.py:140-152); h_out is then the state at that
 *                             step.  Persistent kernel only (hidden <= 256), MSF_E_UNSUPPORTED otherwise. */
#define MSF_LSTM_MAX_SEQS 4
typedef struct msf_lstm_seq {
  const void* x_bf16;
  const void* w_hh;
  const void* w_ih;
  const float* bias;
  void* h_a;
  void* h_b;
  float* cell;
  float* h_out;
  const int32_t* lengths;
  /* stacked layers / training mode (all NULL for single-layer inference) */
  void* h_all;           /* [T+1][B][H] bf16: hidden state BEFORE step t at [t]; [0] zeros on entry (h_a, h_b unused):
                            every hidden state is kept — the next layer's input, the backward pass's operand */
  const void* z_in;      /* layers above the first: [T][B][4H] bf16, columns 4u+g: in_t W_ih^T of every step, computed
                            beforehand by one GEMM (msf_gemm_bf16); replaces x_bf16 / w_ih; may be `gates` itself */
  /* training mode: set by the caller of msf_lstm_forward when msf_lstm_backward follows (cell unused) */
  void* gates;           /* [T][B][4H] bf16, columns 4u+g: gate activations after the forward pass, gradients of the
                            gate pre-activations after the backward pass (in place) */
  float* c_all;          /* [T][B*H] fp32: cell state after every step (implementation-defined order inside a step) */
  /* msf_lstm_backward only */
  const void* w_hh_t;    /* [H][4H] bf16: recurrent weights transposed, w_hh_t[n][4u+g] = weight_hh[g*H+u][n] */
  const float* d_h_out;  /* [B][H] fp32: gradient of the loss with respect to h_out (NULL for a layer below the top) */
  const void* d_h_all;   /* [T][B][H] bf16 or NULL: gradient with respect to EVERY step's hidden state, from the layer
                            above (d a of that layer times its W_ih, through the inter-layer dropout mask) */
  float* dc;             /* [B*H] fp32 scratch, ZERO on entry */
  float* partial;        /* fp32 scratch, msf_lstm_backward_scratch_bytes() */
  float* d_w_ih;         /* [4H][F] fp32, nn.LSTM row order (gate-major): gradient of weight_ih_l0 */
  float* d_w_hh;         /* [4H][H] fp32: gradient of weight_hh_l0 */
  float* d_bias;         /* [4H]    fp32: gradient of bias_ih_l0 (= that of bias_hh_l0) */
  int32_t features;      /* F = input_dim of this layer: 1..64 for the first layer, H for the layers above */
  int32_t in_cols;       /* row length of x_bf16 in msf_lstm_backward: 0 for the first layer (64 padded columns); H for a layer above,
                            whose x_bf16 is the [T][B][H] input the forward GEMM read (no ones column: the bias gradient
                            is a column sum of d a) */
  int32_t cell_type;     /* 0 = LSTM; 1 = GRU: the four columns of a unit are (r, z, n_x, n_h) — w_ih row 4u+3 and w_hh row
                            4u+2 are zero, bias = (b_ir + b_hr, b_iz + b_hz, b_in, b_hn) — and `cell` holds h in fp32 (zeros
                            on entry, also in training mode; c_all is unused).  Training mode keeps (r, z, n, a_nh + b_hn)
                            in `gates`; msf_lstm_backward returns the gradients in the same four-gate row order
                            (d_w_ih rows of n_h and d_w_hh rows of n_x are meaningless; d_bias = (d b_r, d b_z, d b_in, d b_hn)) */
} msf_lstm_seq;
int msf_lstm_forward(const msf_lstm_seq* seqs, int32_t n, int64_t batch, int32_t steps, int32_t hidden, void* stream);
/* x (B, T, F) fp32 windows (batch_first, contiguous) -> x_bf16 [T][B][64] as msf_lstm_forward reads it: time-major, bf16,
 * F <= 64 features zero-padded to 64 columns; ones_column != 0 (F <= 63): 1.0 in column F (bias gradient of
 * msf_lstm_backward).  out_bf16 must be 32-byte aligned. */
int msf_lstm_pack_input(const float* x, int64_t batch, int32_t steps, int32_t features, int32_t ones_column,
                        void* out_bf16, void* stream);
/* sizeof(msf_lstm_seq) as compiled, for binding self-checks */
int msf_lstm_seq_bytes(void);

/* Backward pass of the recurrence (training mode of SequenceEncoder, src/encoders.py:135-166 under autograd): given
 * d_h_out and what msf_lstm_forward kept (h_all, gates, c_all), ONE persistent launch walks the steps backwards
 * (lstm_bwd.cu: d h_{t-1} = d a_t W_hh on tcgen05 with this CTA's rows of W_hh^T resident in shared memory, the cell
 * backward in the epilogue, d a_t written over the gate activations), then the weight gradients
 * d W_hh = sum_t d a_t^T h_{t-1}, d W_ih = sum_t d a_t^T x_t are tensor-core GEMMs contracting over (t, window), split
 * over chunks of steps, and one reduction kernel sums the chunks into nn.LSTM's layout.  The bias gradient is the
 * column `features` of d W_ih: the caller stores 1.0 in column `features` of x_bf16 (features <= 63) whose weight
 * column is zero (features == 64, and the layers above the first: a fixed-order column sum of d a instead).  No gradient with respect to x (the encoders' inputs are data).  hidden <= 256. */
int msf_lstm_backward_scratch_bytes(int64_t batch, int32_t steps, int32_t hidden, size_t* bytes);
int msf_lstm_backward(const msf_lstm_seq* seqs, int32_t n, int64_t batch, int32_t steps, int32_t hidden, void* stream);
/* Inter-layer dropout of a stacked nn.LSTM in training mode: out = in * m over rows x cols bf16 (cols % 8 == 0), m the
 * library's Philox multipliers (0 or 1/(1-p)) of site 4, sub = layer (msf_dropout_mask(seed, offset, 4, layer, ...)
 * returns the same ones); in == out is allowed.  Forward: the input of layer `layer`; backward: its gradient. */
int msf_lstm_dropout(const void* in_bf16, void* out_bf16, int64_t rows, int32_t cols, float p, uint64_t seed,
                     uint64_t offset, int32_t layer, void* stream);

/* fp32 parity path of the same recurrence (lstm_f32.cu: FFMA GEMMs + pointwise cell kernels, one launch pair per time
 * step), ONE LAYER per call, forward and backward incl. the gradient with respect to the input: what SequenceEncoder
 * runs by default (precision "fp32", max-abs <= 1e-5 against the reference's nn.LSTM).  Time-major contiguous layouts:
 *   x [T][B][in_dim];  h_seq, c_seq [T+1][B][H]: slice t = state BEFORE step t, slice 0 ZERO on entry;  gates [T][B][4H],
 *   nn.LSTM's gate-major columns (i | f | g | o): activations after forward, d pre-activations after backward (in place)
 *   w_ih [4H][in_dim], w_hh [4H][H], b_ih / b_hh [4H] or NULL: nn.LSTM's own parameter tensors
 *   lengths [B] int32 or NULL: a window's state stands still from step lengths[b] on (pack_padded_sequence)
 *   forward scratch: B*4H floats;  backward scratch: 3*B*H floats
 *   d_h_last [B][H] or NULL: gradient of the state after a window's last valid step;  d_h_seq [T][B][H] or NULL: gradient
 *   of every step's hidden state (from the layer above);  d_x [T][B][in_dim] or NULL;  d_bias [4H] or NULL (= d b_ih = d b_hh) */
int msf_lstm_f32_forward(const float* x, int32_t in_dim, const float* w_ih, const float* w_hh, const float* b_ih,
                         const float* b_hh, const int32_t* lengths, int64_t batch, int32_t steps, int32_t hidden,
                         float* h_seq, float* c_seq, float* gates, float* scratch, void* stream);
int msf_lstm_f32_backward(const float* x, int32_t in_dim, const float* w_ih, const float* w_hh, const int32_t* lengths,
                          int64_t batch, int32_t steps, int32_t hidden, const float* h_seq, const float* c_seq,
                          float* gates, const float* d_h_last, const float* d_h_seq, float* scratch, float* d_x,
                          float* d_w_ih, float* d_w_hh, float* d_bias, void* stream);

/* The GRU cell of SequenceEncoder (src/encoders.py:66-72: nn.GRU, gate order r | z | n) on the same fp32 kernels, one
 * layer per call, same conventions as msf_lstm_f32_*: gates is [T][B][4H] ((r | z | n | h-share of the candidate gate)
 * after the forward pass; the gradient of the input's share of the pre-activations in its first 3H columns after the
 * backward pass), dzh [T][B][3H] receives the gradient of the recurrent share.  w_ih [3H][in_dim], w_hh [3H][H], biases
 * [3H] or NULL.  forward scratch: B*3H floats; backward scratch: 2*B*H floats. */
int msf_gru_f32_forward(const float* x, int32_t in_dim, const float* w_ih, const float* w_hh, const float* b_ih,
                        const float* b_hh, const int32_t* lengths, int64_t batch, int32_t steps, int32_t hidden,
                        float* h_seq, float* gates, float* scratch, void* stream);
int msf_gru_f32_backward(const float* x, int32_t in_dim, const float* w_ih, const float* w_hh, const int32_t* lengths,
                         int64_t batch, int32_t steps, int32_t hidden, const float* h_seq, float* gates, float* dzh,
                         const float* d_h_last, const float* d_h_seq, float* scratch, float* d_x, float* d_w_ih,
                         float* d_w_hh, float* d_b_ih, float* d_b_hh, void* stream);

/* ---- BatchNorm1d -> ReLU -> Dropout behind a Linear layer (src/encoders.py:339-397, BatchNorm at :374-375) ----- */
/* out = dropout(relu((y - mean) * invstd * gamma + beta)) over y (rows x cols, row-major fp32: the Linear output).
 * training != 0: batch statistics (biased variance; fp64 column sums), running_mean / running_var moved on like
 * nn.BatchNorm1d (momentum, unbiased variance; either may be NULL), Philox dropout keyed by `seed`;
 * training == 0: the running statistics, no dropout.  gamma / beta may be NULL (affine=False), relu == 0 skips the
 * ReLU.  save_mean / save_invstd (cols floats each) receive what the backward pass needs; scratch = 2 * cols
 * doubles, ZERO on entry, zero again on return (reusable across calls on one stream).  rows == 1 in training mode
 * is refused like PyTorch does. */
int msf_bn_act_forward(const float* y, float* out, int64_t rows, int32_t cols, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, float momentum, float eps, int32_t training,
                       int32_t relu, float dropout_p, uint64_t seed, float* save_mean, float* save_invstd,
                       double* scratch, void* stream);
/* Gradients of the above: dy (rows x cols) with respect to the Linear output, dgamma / dbeta (cols, may be NULL).
 * `out` is the forward result: the ReLU / dropout gate is recovered from out != 0. */
int msf_bn_act_backward(const float* dout, const float* y, const float* out, int64_t rows, int32_t cols,
                        const float* gamma, const float* save_mean, const float* save_invstd, int32_t training,
                        int32_t relu, float dropout_p, float* dy, float* dgamma, float* dbeta, double* scratch,
                        void* stream);

/* ---- temporal pooling of FrameEncoder (src/encoders.py:258-336) -------------------------------- */
/* x (B, T, H) fp32 frame features, mask (B, T) fp32 or NULL (0 = frame absent).  mode 0: attention pooling — scores
 * s_t = x_t . w + bias (w: H floats, bias: 1 float or NULL), masked softmax over the frames (a clip without a valid
 * frame pools to zeros, like the reference's nan_to_num), pooled = sum_t p_t x_t; mode 1: (masked) average; mode 2:
 * (masked) maximum.  pooled (B, H).  weights (B, T) receives p_t (modes 0 / 1), argmax (B, H) int32 the frame of the
 * maximum (mode 2, -1 for a clip without a valid frame): the tape of the backward pass.  One block per clip, the clip
 * is read once from DRAM.  T <= 8192. */
int msf_frame_pool_forward(const float* x, const float* mask, const float* w, const float* bias, int64_t batch,
                           int32_t frames, int32_t hidden, int32_t mode, float* pooled, float* weights, int32_t* argmax,
                           void* stream);
/* dx (B, T, H); mode 0 also dw (H) and db (1, may be NULL); scratch: (B * H + B) floats (mode 0). */
int msf_frame_pool_backward(const float* x, const float* w, const float* weights, const int32_t* argmax,
                            const float* d_pooled, int64_t batch, int32_t frames, int32_t hidden, int32_t mode, float* dx,
                            float* dw, float* db, float* scratch, void* stream);

/* ---- tensor-core GEMM building block (encoder/attention projections) ------ */
/* bf16 operands, fp32 accumulate in TMEM via tcgen05.mma, operands staged by TMA
 * into 128B-swizzled shared memory; D is fp32 or bf16 row-major (ldd elements).
 *   mn_major == 0:  D[m,n] = sum_k A[m,k] * B[n,k]   A (m x k, lda), B (n x k, ldb): nn.Linear forward / dgrad
 *   mn_major != 0:  D[m,n] = sum_k A[k,m] * B[k,n]   A (k x m, lda), B (k x n, ldb): weight gradient dY^T . X
 * then D = D + bias[n] (bias fp32 or NULL), then max(D, 0) if relu != 0.
 * m, n, k need not be tile multiples (TMA zero-fills out-of-range boxes); operand
 * base pointers must be 16-byte aligned and lda, ldb multiples of 8 elements. */
int msf_gemm_bf16(const void* a_bf16, const void* b_bf16, void* d, int32_t d_is_bf16, int64_t m, int64_t n,
                  int64_t k, int64_t lda, int64_t ldb, int64_t ldd, int32_t mn_major, const float* bias,
                  int32_t relu, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSF_B200_H_ */
