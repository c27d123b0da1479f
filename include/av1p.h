/* av1p.h - C ABI of the B200-native AV1 partition-prediction cascade (libav1p.so).
 *
 * The reference (chiarorosa/cnn-av1-research) is pure Python: its "operator interface" for this
 * path is a set of Python callables.  Each entry point below names the reference callable it
 * replaces (paths relative to the reference repository root); the Python package
 * cnn_av1_research_b200 re-exposes them under the reference's own names and signatures.
 *
 * Conventions: plain pointers and sizes only; every *_dev pointer is a CUDA device pointer owned by
 * the caller; `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on
 * that stream and no function synchronises unless stated; functions return 0 on success or a
 * negative AV1P_E* code, with a human-readable message available from av1p_last_error() (thread
 * local).  No C++ exception crosses this boundary.  A handle is immutable after creation and may
 * be shared by threads that use distinct streams and distinct workspaces.
 */
#ifndef AV1P_H_
#define AV1P_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AV1P_OK 0
#define AV1P_EINVAL (-1)   /* bad argument / malformed weight blob */
#define AV1P_ECUDA (-2)    /* CUDA runtime or driver error */
#define AV1P_ENOMEM (-3)   /* workspace too small */
#define AV1P_ENODEV (-4)   /* no sm_100 device */

typedef struct av1p_model av1p_model;       /* one packed stage network resident in HBM */
typedef struct av1p_stage av1p_stage;       /* a model bound to a workspace (ready-to-launch plan) */
typedef struct av1p_cascade av1p_cascade;   /* Stage1 -> Stage2 -> Stage3-RECT / Stage3-AB */
typedef struct av1p_flat_cascade av1p_flat_cascade;   /* Stage1 -> 7-way Stage2FlatModel (008b) */

/* Where the 16x16 luma blocks come from. */
typedef struct av1p_input {
  int32_t kind;             /* 0: planar YUV 4:2:0 10-bit LE frames in HBM; 1: float32 blocks [n][256] */
  int32_t width, height;    /* kind 0: luma geometry in samples */
  int32_t pitch;            /* kind 0: luma row pitch in samples (>= width) */
  int32_t n_frames;         /* kind 0 */
  int64_t frame_stride;     /* kind 0: samples between consecutive frames (Y+U+V = W*H*3/2 when packed) */
  const uint16_t* frames_dev;
  const float* images_dev;  /* kind 1: what HierarchicalPipelineV6.predict(images) receives */
} av1p_input;

const char* av1p_last_error(void);
int av1p_version(void);
/* Value written by a kernel watchdog (pipeline barrier that never completed), 0 if none. */
int av1p_debug_watchdog(void);
/* Per-device runtime switches (the reference has no counterpart: scheduling knobs of this build).  Every device
 * ordinal has its own context, initialised on first use while that device is current (cudaSetDevice / torch.cuda.device).
 *   "grid_sms": SMs a persistent kernel's grid may occupy (even, 2 .. SM count; 0 restores the SM count);
 *   "fc_pair" : 1 = FC layers on CTA pairs (default), 0 = single-CTA kernel;
 *   "stem_tma": 1 = frame input goes through the TMA-staged stem when a tensor map can describe the frames (default),
 *               0 = per-thread gather stem (read at every launch);
 *   "pdl"     : 1 = programmatic dependent launch between the kernels of a stage (default), 0 = plain stream order
 *               (read at every launch);
 *   "speculate": 1 = av1p_cascade_predict runs calls of up to 4,096 blocks speculatively - every stage on every block,
 *               side by side on four streams, results compacted by the routing lists; bit-identical outputs (default),
 *               0 = always the routed order (read at every call);
 *   "cr_resid_epi": layer1 residual convolutions add the identity branch 1 = in the epilogue, in place in the TMA-store
 *               staging tiles (default), 2 = in the epilogue from per-thread global loads, 0 = on the tensor core (read
 *               when a stage / cascade is created).
 * av1p_get_option also answers "sms"; it returns -1 for an unknown name. */
int av1p_set_option(const char* name, int32_t value);
int av1p_get_option(const char* name);

/* ---- model: replaces nn.Module construction + load_state_dict + .to(device).eval()
 *      (pesquisa_v6/scripts/008_run_pipeline_eval_v6.py:219-250, models.py:206-251,
 *       scripts/006_train_stage3_ab_fgvc.py:246-297).  `blob` is produced by
 *      cnn_av1_research_b200.packer from a reference state_dict (BN folded, convolutions unrolled to
 *      block-Toeplitz fp16 tiles).  The library copies it to the device. */
int av1p_model_create(const void* blob, size_t bytes, av1p_model** out);
void av1p_model_destroy(av1p_model* m);
int av1p_model_num_outputs(const av1p_model* m);

/* ---- single stage forward: replaces StageXModel.forward / FGVCModel.forward
 *      (models.py:213-251, 006:277-297) fused with block extraction + normalisation
 *      (pesquisa_v5/005_rearrange_video_YUV_420_10bit_LOSSLESS.py:353-457, data_hub.py:70-77). */
size_t av1p_stage_workspace_bytes(const av1p_model* m, int32_t capacity_rows);
int av1p_stage_create(const av1p_model* m, int32_t capacity_rows, void* workspace_dev, size_t workspace_bytes,
                      av1p_stage** out);
void av1p_stage_destroy(av1p_stage* s);
/* Row r of the output is the network applied to block id idx_dev[r] (or r when idx_dev is NULL).
 * n is the host-side upper bound on rows (sizes the grids); if n_dev is non-NULL the kernels read
 * the actual row count from it on the device, so no host synchronisation is needed between stages.
 * logits_dev: float32 [n][num_outputs]. */
int av1p_stage_forward(av1p_stage* s, const av1p_input* in, const int32_t* idx_dev, const int32_t* n_dev, int32_t n,
                       float* logits_dev, void* stream);
/* FGVCModel.forward(x, return_features=True) (scripts/006_train_stage3_ab_fgvc.py:277-296): when set, forwards of an
 * FGVC stage also write the L2-normalised features, fp32 [rows][512], into the caller's buffer; NULL switches it off. */
int av1p_stage_set_features_out(av1p_stage* s, float* features_dev);

/* ---- full cascade: replaces HierarchicalPipelineV6.__init__/predict (008:41-127).
 *      models[] = {stage1, stage2, stage3_rect, stage3_ab}.  Labels use predict()'s label space:
 *      0 NONE, 1 SPLIT, 2 HORZ, 3 VERT, 4 HORZ_A, 5 HORZ_B, 6 VERT_A, 7 VERT_B. */
size_t av1p_cascade_workspace_bytes(const av1p_model* const models[4], int32_t capacity_blocks);
int av1p_cascade_create(const av1p_model* const models[4], int32_t capacity_blocks, void* workspace_dev,
                        size_t workspace_bytes, av1p_cascade** out);
void av1p_cascade_destroy(av1p_cascade* c);
/* Either label pointer may be NULL.  n_blocks <= capacity_blocks. */
int av1p_cascade_predict(av1p_cascade* c, const av1p_input* in, int32_t n_blocks, float stage1_threshold,
                         uint8_t* labels_u8_dev, int64_t* labels_i64_dev, void* stream);
/* Introspection for the parity tests (device pointers into the workspace, valid after predict):
 * which: 0 stage-1 logits [n][1], 1 stage-2 logits [n2][3], 2 RECT logits [nR][2], 3 AB logits [nA][4],
 *        4 idx2 [n2], 5 idxR [nR], 6 idxA [nA], 7 counts int32[4] = {n2, unused, nR, nA},
 *        8 input-range flag int32[1]: frame input (kind 0) is 10-bit content (005:198-204 warns above 1023); the fused
 *          extraction keeps a sample as one fp16 integer, exact up to 2048, and sets this word to 1 when it meets a larger
 *          one (12-bit / corrupt data), where the reference would divide the exact value.  The caller zeroes and reads it
 *          in stream order; such content belongs on the float-block entry (kind 1), which has no limit. */
const void* av1p_cascade_buffer(const av1p_cascade* c, int32_t which);
/* Number of kernels one predict() call enqueues (for launch accounting). */
int av1p_cascade_launches_per_predict(const av1p_cascade* c);
int av1p_stage_launches_per_forward(const av1p_stage* s);

/* ---- extraction: replaces extract_blocks_with_validation (005:353-457) and
 *      BlockRecord.to_torch (data_hub.py:70-77).  y_dev is one luma plane [height][pitch] uint16;
 *      out is [ceil(H/b)*ceil(W/b)][b][b] in row-major grid order, zero padded bottom/right.
 *      block in {8,16,32,64} (005:32). */
int av1p_extract_u16(const uint16_t* y_dev, int32_t width, int32_t height, int32_t pitch, int32_t block,
                     uint16_t* out_dev, void* stream);
int av1p_extract_norm_u16(const uint16_t* y_dev, int32_t width, int32_t height, int32_t pitch, int32_t block,
                          float* out_dev, void* stream);
/* Same tiling for n_frames luma planes `frame_stride_elems` samples apart (a planar YUV 4:2:0 sequence resident in HBM,
 * 005:166-172 `frame_offset = n * frame_size`) in ONE launch; frame f's blocks follow frame f-1's in `out_dev`. */
int av1p_extract_frames_u16(const uint16_t* frames_dev, int32_t n_frames, int64_t frame_stride_elems, int32_t width, int32_t height,
                            int32_t pitch_elems, int32_t block_size, uint16_t* out_dev, void* stream);
int av1p_extract_frames_norm_u16(const uint16_t* frames_dev, int32_t n_frames, int64_t frame_stride_elems, int32_t width,
                                 int32_t height, int32_t pitch_elems, int32_t block_size, float* out_dev, void* stream);

/* ---- routing operators (008:77-125), exposed on their own so they can be checked bit-exactly on
 *      reference logits.  scratch_dev: av1p_route_scratch_bytes() bytes, zero-initialised once.
 *      idx outputs are ascending block ids; counts are written to device memory:
 *      count_dev is int32[2] ({n_routed, unused}), counts2_dev is int32[2] ({n_rect, n_ab}). */
size_t av1p_route_scratch_bytes(void);
int av1p_route_stage1(const float* logits_dev, const int32_t* n_dev, int32_t n, float threshold, int32_t* idx_dev,
                      int32_t* count_dev, uint8_t* labels_u8_dev, int64_t* labels_i64_dev, void* scratch_dev,
                      void* stream);
int av1p_route_stage2(const float* logits3_dev, const int32_t* idx_in_dev, const int32_t* n_dev, int32_t n,
                      int32_t* idx_rect_dev, int32_t* idx_ab_dev, int32_t* counts2_dev, uint8_t* labels_u8_dev,
                      int64_t* labels_i64_dev, void* scratch_dev, void* stream);
int av1p_finalize_labels(const float* logits_dev, int32_t num_classes, int32_t label_base, const int32_t* idx_dev,
                         const int32_t* n_dev, int32_t n, uint8_t* labels_u8_dev, int64_t* labels_i64_dev,
                         void* stream);

/* Same scatter with a plain argmax over the raw logits (first max wins), num_classes <= 8: the flatten cascade's
 * `stage2_logits.argmax(dim=1)` (scripts/008b_run_pipeline_flatten_eval.py:213). */
int av1p_finalize_labels_argmax(const float* logits_dev, int32_t num_classes, int32_t label_base, const int32_t* idx_dev,
                                const int32_t* n_dev, int32_t n, uint8_t* labels_u8_dev, int64_t* labels_i64_dev,
                                void* stream);

/* ---- flatten cascade: replaces run_pipeline_inference's per-batch body
 *      (pesquisa_v6/scripts/008b_run_pipeline_flatten_eval.py:177-229): Stage1Model -> threshold -> compaction ->
 *      Stage2FlatModel (backbone + Linear/BN1d/ReLU/Linear head, 008b:110-127) -> label = 1 + argmax (008b:163-175),
 *      label 0 for blocks below the threshold.  models[] = {stage1, flat7}. */
size_t av1p_flat_cascade_workspace_bytes(const av1p_model* const models[2], int32_t capacity_blocks);
int av1p_flat_cascade_create(const av1p_model* const models[2], int32_t capacity_blocks, void* workspace_dev,
                             size_t workspace_bytes, av1p_flat_cascade** out);
void av1p_flat_cascade_destroy(av1p_flat_cascade* c);
int av1p_flat_cascade_predict(av1p_flat_cascade* c, const av1p_input* in, int32_t n_blocks, float stage1_threshold,
                              uint8_t* labels_u8_dev, int64_t* labels_i64_dev, void* stream);
/* which: 0 stage-1 logits [n][1], 1 flat logits [n2][7], 2 idx2 [n2], 3 counts int32[2] = {n2, unused}. */
const void* av1p_flat_cascade_buffer(const av1p_flat_cascade* c, int32_t which);

/* ---- stage-1 threshold sweep: replaces the accumulation inside evaluate_with_threshold
 *      (pesquisa_v6/scripts/007_optimize_thresholds.py:24-71) for up to 32 thresholds in one pass over the
 *      stage-1 logits: prob = sigmoid(logit) (fp32), pred = (double)prob >= threshold (the reference compares a
 *      float32 array with np.float64 thresholds); counts_dev[t][4] = {tn, fp, fn, tp} against labels_dev (uint8
 *      0/1).  probs_dev (float32 [n]) is optional.  thresholds_host is read at call time. */
int av1p_threshold_sweep(const float* logits_dev, const uint8_t* labels_dev, int32_t n, const double* thresholds_host,
                         int32_t n_thresholds, float* probs_dev, uint64_t* counts_dev, void* stream);

/* ---- ensemble voting over the logits of several Stage-3-AB models: replaces the voting inside ABEnsemble.predict /
 *      predict_with_uncertainty and WeightedEnsemble.predict (pesquisa_v6/v6_pipeline/ensemble.py:29-116, 165-183).
 *      logits_dev: float32 [n_models][n][k] (n_models <= 8, k <= 8).  mode 0 = hard (majority, ties to the smallest class
 *      id as torch.unique + argmax give; conf = majority share), 1 = soft (mean of the per-model softmax), 2 = weighted
 *      soft (weights_dev [n_models], already normalised).  pred_dev int64 [n]; conf_dev float32 [n]; the remaining outputs
 *      are optional (NULL): mean_probs / std_probs [n][k] (unbiased std over the models), agreement [n], all_probs
 *      [n_models][n][k]. */
int av1p_ensemble_vote(const float* logits_dev, int32_t n_models, int32_t n, int32_t k, int32_t mode, const float* weights_dev,
                       int64_t* pred_dev, float* conf_dev, float* mean_probs_dev, float* std_probs_dev, float* agreement_dev,
                       float* all_probs_dev, void* stream);

/* ---- measurement support (bench.py): bracket every kernel launch of the calling thread with CUDA
 *      events on its stream.  av1p_profile_end synchronises on those events and returns the summed
 *      device time and launch count per kernel class: 0 stem, 1 tcgen05 FC, 2 SAM gate, 3 FGVC tail,
 *      4 routing, 5 label finalize, 6 squeeze-excite, 7 resident-weight layer1 conv (arrays of 8). */
int av1p_profile_begin(void);
int av1p_profile_end(float* ms_by_class, int32_t* launches_by_class);
/* Same, per launch in issue order: ms[i] / cls[i] for the first `cap` launches, *n_out = launches recorded. */
int av1p_profile_end_launches(float* ms, int32_t* cls, int32_t cap, int32_t* n_out);

/* ---- host -> device staging of the luma planes of planar 4:2:0 frames (pinned host memory
 *      recommended): copies n_frames * width*height samples, skipping chroma. */
int av1p_upload_luma(const uint16_t* frames_host, int32_t n_frames, int32_t width, int32_t height,
                     int64_t frame_stride, uint16_t* luma_dev, void* stream);

/* ---- kernel-level test hook: one block-sparse FC layer  out = epi(act . W^T)  on the tcgen05 path.
 *      a_dev[i]: fp16 [rows][a_cols[i]] activation sources (unused entries NULL);
 *      w: fp16 [n_w_chunks*block_n][64]; kb_begin[n_tiles+1]; schedule entry e multiplies the 64-wide
 *      K block (kb_src[e] & 0x3fff) of source (kb_src[e] >> 14) with weight chunk kb_w[e].
 *      See csrc/fc_tcgen05.cuh for the epilogue codes.
 *      Every activation matrix (a_dev, aux_dev, out_dev, their lo planes) is in the library's TILED layout:
 *      rows padded to a multiple of 128, columns (a_cols / aux_ld / out_ld) a multiple of 64, stored as
 *      [rows/128][cols/64][128][64] - element (r, c) at ((r/128 * cols/64 + c/64) * 128 + r%128) * 64 + c%64. */
typedef struct av1p_fc_desc {
  const void* a_dev[4]; int32_t a_cols[4];
  int32_t rows;
  const int32_t* n_dev;
  const void* w_dev; int32_t n_w_chunks;
  int32_t n_kb_total, n_tiles, block_n, epi;
  const int32_t* kb_begin;
  const uint16_t* kb_src;
  const uint16_t* kb_w;
  const float* bias_dev;
  const float* row_scale_dev;
  float acc_scale;
  int32_t pair_mode;   /* 1: entries are (x_hi,w_hi),(x_lo,w_lo) pairs of one K block -> hi*hi + hi*lo + lo*hi */
  const void* aux_dev; const void* aux_lo_dev; int32_t aux_ld;
  void* out_dev; void* out_lo_dev; int32_t out_ld;
  const float* tail_w_dev; const float* tail_b_dev; float* logits_dev; int32_t tail_n;
} av1p_fc_desc;
int av1p_fc_forward(const av1p_fc_desc* d, void* stream);

/* ---- kernel-level test hook: one layer1 convolution (3x3, stride 1, 64 -> 64 channels on the 4x4 map,
 *      torchvision BasicBlock conv3x3 as used by models.py:110) on the resident-weight tcgen05 path
 *      (csrc/conv_res_tcgen05.cuh).  x: fp16 [rows][1024] ([position][channel] per block), x_lo its low
 *      plane (split != 0); w: fp16 [planes][ky][kx = 2,1,0][64 co][64 ci]; bias float32[1024];
 *      epi 0 linear, 1 relu, 2 relu(acc + bias + aux).  x, aux and out use the tiled layout described above
 *      (rows padded to 128, 16 column blocks). */
typedef struct av1p_conv_res_desc {
  const void* x_dev; const void* x_lo_dev;
  int32_t rows;
  const int32_t* n_dev;
  const void* w_dev;
  int32_t split, epi;
  const float* bias_dev;
  float acc_scale;
  const void* aux_dev; const void* aux_lo_dev;
  void* out_dev; void* out_lo_dev;
} av1p_conv_res_desc;
int av1p_conv_res_forward(const av1p_conv_res_desc* d, void* stream);

/* ---- Stage-1 training step (BASELINE configs[4]): the two non-convolution pieces as single launches.
 *      av1p_focal_loss_binary replaces FocalLoss.forward's binary branch + mean (pesquisa_v6/v6_pipeline/losses.py:29-38,
 *      48-49) AND its autograd backward: *loss_dev = mean_i a_t (1 - pt)^gamma bce_i, dlogits_dev[i] = d loss / d logit_i
 *      (dlogits_dev may be NULL).  logits float32 [n], targets int64 [n] in {0, 1}.
 *      av1p_adamw_flat replaces optimizer.step() of torch.optim.AdamW(lr, weight_decay)
 *      (pesquisa_v6/scripts/003_train_stage1_improved.py:73, 250-254; betas / eps are torch's defaults there) for ALL
 *      parameters at once: param / grad / exp_avg / exp_avg_sq are flat float32 arrays of n elements (same address modulo 16),
 *      grad is multiplied by grad_scale first (1 / world size after a summing all-reduce); *step_dev is the device-side
 *      step counter: incremented by this call when advance_step != 0 (the first segment of a step; parameters that got
 *      no gradient are skipped like torch does, so a step may update several ranges), then used for the bias corrections -
 *      a captured CUDA graph of the step therefore replays correctly. */
int av1p_focal_loss_binary(const float* logits_dev, const int64_t* targets_dev, int32_t n, float alpha, float gamma,
                           float* loss_dev, float* dlogits_dev, void* stream);
int av1p_adamw_flat(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t n, double lr,
                    double beta1, double beta2, double eps, double weight_decay, double grad_scale, int32_t* step_dev, int32_t advance_step,
                    void* stream);

/* ---- data-parallel training step: gradient reduce-scatter + AdamW + parameter all-gather in ONE kernel over NVLink peer
 *      memory (replaces `dist.all_reduce(grads)` + `optimizer.step()` of a replicated data-parallel step,
 *      pesquisa_v6/scripts/003_train_stage1_improved.py:71-73 under DDP).  grad_ptrs / param_ptrs / flag_ptrs: HOST arrays
 *      of `world` device pointers, entry r = rank r's flat gradient buffer / flat parameter buffer (n float32 each, 16-byte
 *      aligned) / flag words (av1p_dp_flag_words() int32, zero-initialised once), each mapped into this process (CUDA IPC)
 *      with peer access enabled.  Rank r owns elements [r * shard, min(n, (r + 1) * shard)) (shard a multiple of 4) and
 *      keeps only that range's moments (exp_avg_shard_dev / exp_avg_sq_shard_dev, `shard` floats each).  [skip_lo, skip_hi):
 *      flat range that received no gradient on any rank (left untouched, like torch skips grad-less parameters; empty if
 *      skip_lo == skip_hi).  epoch: 1, 2, 3, ... identical on every rank for the same step.  *err_dev becomes non-zero if a
 *      peer did not arrive within ~10 s of SM clocks (the kernel then leaves; results are invalid).  All ranks must call this
 *      for every step, each on its own device / stream; *step_dev is advanced like av1p_adamw_flat does. */
int av1p_dp_adamw_fused(const float* const* grad_ptrs, float* const* param_ptrs, int32_t* const* flag_ptrs, int32_t rank,
                        int32_t world, int64_t n, int64_t shard, float* exp_avg_shard_dev, float* exp_avg_sq_shard_dev, double lr,
                        double beta1, double beta2, double eps, double weight_decay, int64_t skip_lo, int64_t skip_hi,
                        int32_t* step_dev, int32_t epoch, int32_t* err_dev, void* stream);
int av1p_dp_flag_words(void);
/* Enable peer access from the CURRENT device to `peer_device` (idempotent): required before av1p_dp_adamw_fused
 * dereferences that device's buffers. */
int av1p_enable_peer_access(int32_t peer_device);
/* CUDA IPC plumbing for those buffers.  av1p_ipc_export: 64-byte handle of the device allocation that holds dev_ptr + the
 * pointer's byte offset inside it (send both to the peer processes).  av1p_ipc_import: maps a peer's allocation into this
 * process for the CURRENT device (peer access is enabled as needed) and returns its base address; a handle may be imported
 * once per process.  av1p_ipc_close unmaps it.  The exporting process must keep the allocation alive. */
int av1p_ipc_export(const void* dev_ptr, uint8_t handle_out[64], int64_t* offset_out);
int av1p_ipc_import(const uint8_t handle[64], void** base_out);
int av1p_ipc_close(void* base);

#ifdef __cplusplus
}
#endif
#endif /* AV1P_H_ */
