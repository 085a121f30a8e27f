/*
 * refinedet_b200.h — C ABI of librefinedet_b200.so
 *
 * B200 (sm_100a) implementation of the RefineDet detection post-processing and
 * anchor-matching hot path.  This library replaces the reference's native
 * extension (utils/nms/gpu_nms.hpp:1-2, utils/nms/nms_kernel.cu, utils/build.py)
 * and backs the Python drop-ins for layers/box_utils.py,
 * layers/functions/detection_refinedet.py and
 * layers/modules/refinedet_multibox_loss.py (paths relative to the reference
 * checkout).
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends
 *     in `_host`; all float tensors are contiguous fp32.
 *   - `stream` is a `cudaStream_t` passed as `void*`; calls are asynchronous on
 *     it.  The library never allocates device memory, never synchronises and
 *     never changes the current device (the one exception is `rd_nms_host`,
 *     which mirrors the reference's synchronous host-pointer ABI).
 *   - every function returns 0 on success, a positive `cudaError_t` value on a
 *     CUDA failure, or a negative `RD_ERR_*` code on an argument error.
 *     `rd_error_string` turns either into text.
 *   - workspaces are caller-owned and reusable; `rd_*_workspace_bytes` gives the
 *     size.  A detect workspace should be passed through `rd_detect_workspace_reset` once
 *     before first use; every call leaves it ready for the next.
 */
#ifndef REFINEDET_B200_H_
#define REFINEDET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RD_ABI_VERSION 1

/* only the C ABI below is exported from the shared library */
#if defined(__GNUC__)
#define RD_API __attribute__((visibility("default")))
#else
#define RD_API
#endif

/* argument errors (negative so they never collide with cudaError_t) */
#define RD_ERR_BAD_ARG      (-1)   /* null pointer / non-positive size */
#define RD_ERR_ALIGNMENT    (-2)   /* a float4-accessed tensor is not 16-byte aligned */
#define RD_ERR_UNSUPPORTED  (-3)   /* size beyond what the kernels support (see RD_MAX_*) */
#define RD_ERR_WORKSPACE    (-4)   /* workspace too small */

/* largest number of boxes one NMS problem may keep in flight (top_k, or n for the
 * standalone calls after top_k truncation) */
#define RD_MAX_NMS_BOXES 4096
/* largest number of ground-truth boxes per image in rd_refine_match */
#define RD_MAX_GT 1024

/* NMS flavour flags */
#define RD_NMS_NORMALISED   0  /* layers/box_utils.py:241-285: area=(x2-x1)*(y2-y1),
                                  IoU = inter/((area_j-inter)+area_i), keep IoU <= thr */
#define RD_NMS_PIXEL_PLUS1  1  /* utils/nms/py_cpu_nms.py:18-36 == nms_kernel.cu:24-32:
                                  +1 widths, IoU = inter/(S_i+S_j-inter), keep IoU <= thr */
#define RD_NMS_SUPPRESS_EQ  2  /* OR-able: suppress on IoU >= thr (utils/nms/cpu_nms.pyx:65) */
/* OR-able into the nms_flags of the fused detect stage only: arm_conf / odm_conf hold LOGITS and the
 * softmax of models/refinedet.py:143-147 (over the last dimension, max-subtracted, fp32) is folded
 * into the stage — the separate read+write pass over odm_conf disappears.  Scores agree with
 * torch.softmax to fp32 rounding (the sum is reduced in a different order). */
#define RD_INPUT_LOGITS     4
/* TEST-ONLY, OR-able into the nms_flags of the fused detect stage: force one instantiation of the per-class
 * kernel (nms_small_kernel) regardless of the grid-size heuristic, so that the parity tests can hold every
 * instantiation to the oracle at small sizes too.  0 = choose automatically (the only value product code uses). */
#define RD_DEBUG_INSTANCE_SHIFT 8
#define RD_DEBUG_INSTANCE_MASK  (3 << RD_DEBUG_INSTANCE_SHIFT)
#define RD_DEBUG_INSTANCE_256   (1 << RD_DEBUG_INSTANCE_SHIFT)   /* <= 256 candidates, 128 threads  */
#define RD_DEBUG_INSTANCE_1024  (2 << RD_DEBUG_INSTANCE_SHIFT)   /* <= 1024 candidates, 256 threads */
/* output row layout of the fused detect stage */
#define RD_ROW_BOX_SCORE    0  /* x1,y1,x2,y2,score  (eval_refinedet_coco.py:226) */
#define RD_ROW_SCORE_BOX    1  /* score,x1,y1,x2,y2  (detection_refinedet.py:106-108) */

RD_API int         rd_abi_version(void);
RD_API const char* rd_error_string(int code);
/* number of kernels this library has launched in this process (bench accounting) */
RD_API unsigned long long rd_launch_count(void);

/* ---- layers/box_utils.py element-wise functions --------------------------- */
/* point_form (box_utils.py:5-14): [n,4] (cx,cy,w,h) -> (x1,y1,x2,y2) */
RD_API int rd_point_form(const float* boxes, float* out, int n, void* stream);
/* center_size (box_utils.py:17-26): [n,4] (x1,y1,x2,y2) -> (cx,cy,w,h) */
RD_API int rd_center_size(const float* boxes, float* out, int n, void* stream);
/* decode (box_utils.py:187-205): loc[n,4], priors[n,4] (cx,cy,w,h) -> [n,4] point form */
RD_API int rd_decode(const float* loc, const float* priors, float v0, float v1,
              float* out, int n, void* stream);
/* encode (box_utils.py:162-183): matched[n,4] point form, priors[n,4] -> [n,4] */
RD_API int rd_encode(const float* matched, const float* priors, float v0, float v1,
              float* out, int n, void* stream);
/* intersect / jaccard (box_utils.py:29-68): box_a[A,4] x box_b[Bn,4] -> out[A,Bn] */
RD_API int rd_intersect(const float* box_a, const float* box_b, float* out, int A, int Bn, void* stream);
RD_API int rd_jaccard(const float* box_a, const float* box_b, float* out, int A, int Bn, void* stream);

/* ---- Detect_RefineDet.forward (detection_refinedet.py:27-65) --------------- */
/* ARM filter + two-stage decode, dense outputs.  `odm_conf` is modified IN PLACE
 * exactly like the reference (:40-42): rows with arm_conf[...,1] <= objectness_thre
 * become all-zero.  boxes_out[B,P,4] point form, unclipped; scores_out[B,P,C]. */
RD_API int rd_detect_forward(const float* arm_loc, const float* arm_conf, const float* odm_loc,
                      float* odm_conf, const float* priors, int B, int P, int C,
                      float objectness_thre, float v0, float v1,
                      float* boxes_out, float* scores_out, void* stream);

/* The in-place half of the above on its own (detection_refinedet.py:79-81, the first step of
 * forward_python_nms): rows r of odm_conf[rows,C] with arm_conf[r,1] <= objectness_thre become all-zero. */
RD_API int rd_arm_zero_rows(const float* arm_conf, float* odm_conf, long long rows, int C,
                     float objectness_thre, void* stream);

/* SURVEY §8b name of the same entry point ("rd_decode_filter": ARM-objectness filter + two-stage decode) */
RD_API int rd_decode_filter(const float* arm_loc, const float* arm_conf, const float* odm_loc,
                      float* odm_conf, const float* priors, int B, int P, int C,
                      float objectness_thre, float v0, float v1,
                      float* boxes_out, float* scores_out, void* stream);

/* ---- stand-alone threshold + top-k select ---------------------------------------- */
/* Per (image b, class c >= first_class): the anchors with scores[b,p,c] > conf_thresh
 * (eval_refinedet_coco.py:214, detection_refinedet.py:98), the top_k highest of them in
 * score-descending order (eval :222 `argsort()[::-1][:top_k]`; box_utils.py:242-244), lower anchor
 * index first on equal scores.  scores[B,P,C] (the a3 `scores` output or any [B,P,C] tensor);
 *   idx_out   [B,C,top_k] int32   anchor index per rank; only the first count_out[b,c] entries are written
 *   score_out [B,C,top_k] float   their scores (may be NULL)
 *   count_out [B,C]       int32   min(#candidates, top_k); 0 for classes < first_class
 * No workspace.  min(top_k, P) <= 4 * RD_MAX_NMS_BOXES.  The fused stage does not call this (it selects
 * inside its per-class CTAs); it is the §8b `rd_select_topk` for callers that want the lists. */
RD_API int rd_select_topk(const float* scores, int B, int P, int C, float conf_thresh, int top_k,
                    int first_class, int* idx_out, float* score_out, int* count_out, void* stream);

/* ---- fused detect stage ---------------------------------------------------- */
/* (Detect_RefineDet.forward + the per-class loop of eval_refinedet_coco.py:205-232,
 *  or Detect_RefineDet.forward_python_nms, detection_refinedet.py:67-113, depending
 *  on the flags.)  Inputs are read once and are NOT modified. */
RD_API size_t rd_detect_workspace_bytes(int B, int P, int C);
RD_API int    rd_detect_workspace_reset(void* workspace, size_t workspace_bytes, void* stream);
/* per (image b, class c>=1):
 *   candidates = anchors with arm_conf[b,p,1] > objectness_thre and
 *                odm_conf[b,p,c] > conf_thresh; the top_k highest scores are kept
 *                (ties: lower anchor index first), boxes are the two-stage decode
 *                multiplied by img_scale[b] (may be NULL = no scaling), greedy NMS in
 *                score order with `nms_flags`, at most `max_out` rows are emitted.
 * outputs:
 *   out_counts [B,C]            int32   (class 0 is always 0)
 *   out_dets   [B,C,max_out,5]  float   rows per `row_layout`; only the first
 *                                       out_counts[b,c] rows of a slot are written
 *   out_anchor [B,C,max_out]    int32   anchor index of each row (may be NULL)      */
RD_API int rd_detect_fused(const float* arm_loc, const float* arm_conf, const float* odm_loc,
                    const float* odm_conf, const float* priors, int B, int P, int C,
                    float objectness_thre, float conf_thresh, float nms_thresh,
                    int top_k, int max_out, const float* img_scale, int nms_flags,
                    int row_layout, float v0, float v1,
                    void* workspace, size_t workspace_bytes,
                    int* out_counts, float* out_dets, int* out_anchor, void* stream);

/* SURVEY §8b names of rd_detect_fused / rd_detect_workspace_bytes (same arguments, same behaviour) */
RD_API int rd_detect(const float* arm_loc, const float* arm_conf, const float* odm_loc,
                    const float* odm_conf, const float* priors, int B, int P, int C,
                    float objectness_thre, float conf_thresh, float nms_thresh,
                    int top_k, int max_out, const float* img_scale, int nms_flags,
                    int row_layout, float v0, float v1,
                    void* workspace, size_t workspace_bytes,
                    int* out_counts, float* out_dets, int* out_anchor, void* stream);
RD_API size_t rd_workspace_bytes(int B, int P, int C);

/* Diagnostics twin of rd_detect_fused: records CUDA events between the stage's kernels on
 * `stream`, WAITS for the stage, and writes the device time in ms of
 * {collect_kernel, graph_kernel, nms_small_kernel + nms_large_kernel, 0} to
 * stage_ms_host[4] (host pointer).  The events serialise graph_kernel and nms_small_kernel, which overlap
 * in rd_detect_fused. */
RD_API int rd_detect_fused_timed(const float* arm_loc, const float* arm_conf, const float* odm_loc,
                    const float* odm_conf, const float* priors, int B, int P, int C,
                    float objectness_thre, float conf_thresh, float nms_thresh,
                    int top_k, int max_out, const float* img_scale, int nms_flags,
                    int row_layout, float v0, float v1,
                    void* workspace, size_t workspace_bytes,
                    int* out_counts, float* out_dets, int* out_anchor, void* stream,
                    float* stage_ms_host);

/* Plans: the launch chain of rd_detect_fused for FIXED buffers, captured once into a CUDA graph
 * (programmatic-launch edges included) and replayed with one driver call.  A serving loop whose
 * model writes its head outputs into static buffers replays the plan every batch; several plans
 * over different workspaces / output buffers replayed on different streams keep more than one
 * batch in flight (bench.py's pipelined mode).  Replaces nothing in the reference — its eval loop
 * (eval_refinedet_coco.py:205-232) issues ~160 host-synchronous calls per image.
 *   rd_detect_plan_create : same arguments as rd_detect_fused (no stream); the pointers are baked
 *                           into the plan and must stay valid while it lives.  Creates and destroys
 *                           a private capture stream; allocates host/driver objects only.
 *   rd_detect_plan_launch : asynchronous replay on `stream`.
 *   rd_detect_plan_destroy: releases the graph.                                              */
typedef struct rd_detect_plan rd_detect_plan;
RD_API int rd_detect_plan_create(const float* arm_loc, const float* arm_conf, const float* odm_loc,
                    const float* odm_conf, const float* priors, int B, int P, int C,
                    float objectness_thre, float conf_thresh, float nms_thresh,
                    int top_k, int max_out, const float* img_scale, int nms_flags,
                    int row_layout, float v0, float v1,
                    void* workspace, size_t workspace_bytes,
                    int* out_counts, float* out_dets, int* out_anchor, rd_detect_plan** plan_out);
/* Generic capture: everything the caller enqueues on `stream` (a non-default stream) between begin and end --
 * rd_detect_fused followed by the exchange of its result, for example -- becomes one plan.  Thread-local capture. */
RD_API int rd_plan_capture_begin(void* stream);
RD_API int rd_plan_capture_end(void* stream, rd_detect_plan** plan_out);
RD_API int rd_detect_plan_launch(rd_detect_plan* plan, void* stream);
RD_API int rd_detect_plan_destroy(rd_detect_plan* plan);

/* compact [B,C,max_out,5] slots into packed rows (score-descending inside a class,
 * classes ascending, images ascending).  out_offsets[B*C+1] = exclusive prefix sum
 * of counts; packed[total,5]; packed_capacity = rows available in `packed`. */
RD_API int rd_pack_detections(const int* counts, const float* dets, int B, int C, int max_out,
                       int* out_offsets, float* packed, int packed_capacity, void* stream);

/* Result wire format of the reference's eval (data/sarship_coco.py:293-336, built there by a Python double loop):
 * COCO result records straight from the device-side detections, in the reference's order -- classes ascending
 * (class 0 and classes with class_to_cat[c] < 0 skipped; class_to_cat may be NULL), images ascending inside a
 * class, rows score-descending -- with bbox = [x, y, x2 - x + 1, y2 - y + 1] in float64 (:296-303).
 *   dets: the slot layout [B,C,max_out,5] (max_out > 0), or packed rows [total,5] with src_offsets[B*C+1] = their
 *         image-major exclusive prefix (max_out = 0: the output of rd_pack_detections or of the multi-GPU gather)
 *   out_ids  [capacity,2] int32 (image index b, class c);  out_vals [capacity,5] float64 (x, y, w, h, score);
 *   out_total[1] = number of records (may exceed capacity: then only the first `capacity` were written).       */
RD_API int rd_coco_records(const int* counts, const float* dets, int B, int C, int max_out,
                    const int* src_offsets, const int* class_to_cat, int* out_ids, double* out_vals,
                    int capacity, int* out_total, void* stream);

/* Multi-GPU exchange of the compact detections (the reference has no counterpart: its
 * nn.DataParallel gathers the dense head outputs on GPU 0, train_refinedet.py:138-139).  Packing fused
 * with the gather: this rank's counts and packed rows are stored straight into its slot of EVERY peer's
 * exchange buffer (peer-mapped device memory: P2P stores over NVLink / NVSwitch), no staging copy, no
 * collective call; the caller places a barrier between the ranks afterwards.
 *   slot = [64 x i32 header: rows stored, B, C, rows total | counts B*C x i32, padded to 256 B |
 *           rows capacity_rows x 5 x f32], rd_exchange_slot_bytes(B, C, capacity_rows) bytes, 256-B aligned.
 *   peer_slots_host: HOST array of `world` device pointers, entry w = base of this rank's slot in the
 *   buffer of rank w (entry `rank` is the local one).  scratch_offsets[B*C+1] is local scratch.  slot_B >= B
 *   is the batch the slots were sized for (ranks may hold one image less than the largest shard). */
RD_API size_t rd_exchange_slot_bytes(int B, int C, int capacity_rows);
RD_API int rd_pack_scatter(const int* counts, const float* dets, int B, int C, int max_out,
                           int* scratch_offsets, void* const* peer_slots_host, int world, int rank,
                           int slot_B, int capacity_rows, void* stream);
/* Same, with the two knobs of the copy phase.  The rows are first packed into this rank's OWN slot (local HBM), then
 * the slot's used prefix is streamed to the peers as 16-byte words:
 *   multicast_slot != NULL: the address of this rank's slot in the MULTICAST mapping of the symmetric buffer
 *                           (NVSwitch replicates every store to all ranks: `multimem.st`, the rows leave the GPU once);
 *   multicast_slot == NULL: unicast P2P stores, one group of CTAs per peer.
 *   copy_ctas: CTAs per destination (<= 0 = default).  scratch_offsets may be NULL. */
RD_API int rd_pack_scatter_ex(const int* counts, const float* dets, int B, int C, int max_out,
                           int* scratch_offsets, void* const* peer_slots_host, int world, int rank,
                           int slot_B, int capacity_rows, void* multicast_slot, int copy_ctas, void* stream);

/* The exchange with its synchronisation folded in: two launches and ONE cross-GPU rendezvous per round, no
 * barrier kernels, replayable from a CUDA graph (the round counter lives on the device).
 *   Every rank owns one symmetric buffer  [control block, rd_exchange_ctrl_bytes() | world slots, parity 0 |
 *   world slots, parity 1], zero before the first round; peer_bases_host[w] = base of rank w's buffer as mapped
 *   into THIS process (entry `rank` = the local one); multicast_base = base of the multicast mapping, or NULL.
 *   Round e writes this rank's slot of parity e & 1 on every rank (layout of a slot: rd_pack_scatter), publishes e in
 *   flags[rank] of every rank's control block and returns -- on the stream -- once every rank's flag has arrived
 *   here: the rows of all ranks of round e are then readable in the local buffer (header word 4 of a slot = e).
 *   Consumers read on `stream` (or after it) BEFORE the next round is enqueued; the second parity is what makes a
 *   separate "buffer free" rendezvous unnecessary.  timeout_ms (<= 0: 10 s): a wait that long sets word `error`
 *   of the control block (offset 264) instead of hanging the GPU.  All ranks must call in lockstep. */
RD_API size_t rd_exchange_ctrl_bytes(void);
RD_API int rd_exchange_round(const int* counts, const float* dets, int B, int C, int max_out,
                      void* const* peer_bases_host, void* multicast_base, int world, int rank, int slot_B,
                      int capacity_rows, int copy_ctas, int timeout_ms, void* stream);

/* ---- stand-alone NMS ------------------------------------------------------- */
RD_API size_t rd_nms_workspace_bytes(int n);
/* box_utils.nms (box_utils.py:222-286) / utils.nms_wrapper.nms on device tensors:
 * boxes[n,4], scores[n]; keeps the top_k highest scores, greedy NMS, writes the
 * ORIGINAL indices of the kept boxes (score-descending) to keep_out[n] (int64,
 * entries past the count are left untouched) and the count to count_out[1] (int32). */
RD_API int rd_nms(const float* boxes, const float* scores, int n, float thresh, int top_k,
           int nms_flags, void* workspace, size_t workspace_bytes,
           long long* keep_out, int* count_out, void* stream);
/* Drop-in for `_nms` (utils/nms/gpu_nms.hpp:1-2, nms_kernel.cu:91-144): HOST pointers,
 * boxes_host[boxes_num, boxes_dim>=5] rows x1,y1,x2,y2,score ALREADY sorted by score
 * descending (gpu_nms.pyx:26-29); pixel +1 IoU, suppress IoU > thresh; synchronous.  The reference
 * cudaMalloc'd and freed its scratch on every call; this keeps one grow-only scratch per device. */
RD_API int rd_nms_host(int* keep_out_host, int* num_out_host, const float* boxes_host,
                int boxes_num, int boxes_dim, float nms_overlap_thresh, int device_id);
/* same with explicit flavour flags (RD_NMS_PIXEL_PLUS1 | RD_NMS_SUPPRESS_EQ reproduces the
 * Cython cpu_nms the reference dispatches to under force_cpu, utils/nms_wrapper.py:28-30) */
RD_API int rd_nms_host_ex(int* keep_out_host, int* num_out_host, const float* boxes_host,
                   int boxes_num, int boxes_dim, float nms_overlap_thresh, int device_id,
                   int nms_flags);

/* ---- training side: refine_match + hard-negative mining -------------------- */
/* refine_match / match over a batch (box_utils.py:70-160, called per image at
 * refinedet_multibox_loss.py:75-86).
 *   truths [B,Gmax,4] point form, labels [B,Gmax] float, gt_count [B] int32 (<= Gmax)
 *   priors [P,4] (cx,cy,w,h);  arm_loc [B,P,4] or NULL (ARM criterion / SSD match)
 *   label_mode: 0 = conf = (int64)labels[g]          (ODM, 1-based labels)
 *               1 = conf = labels[g] >= 0 ? 1 : 0    (ARM 2-class, :78-79)
 *               2 = conf = (int64)labels[g] + 1      (SSD match(), box_utils.py:107)
 *   outputs loc_t [B,P,4] float, conf_t [B,P] int64, best_truth_idx [B,P] int32 and
 *   best_truth_overlap [B,P] float (after the forced matches of :146-150; both are
 *   REQUIRED — they double as the scratch between the two kernels).
 * An image with gt_count == 0 yields conf_t = 0 and loc_t = 0 (the reference raises). */
RD_API size_t rd_match_workspace_bytes(int B, int Gmax);
RD_API int rd_refine_match(const float* truths, const float* labels, const int* gt_count,
                    const float* priors, const float* arm_loc, int B, int P, int Gmax,
                    float threshold, float v0, float v1, int label_mode,
                    void* workspace, size_t workspace_bytes,
                    float* loc_t, long long* conf_t, int* best_truth_idx,
                    float* best_truth_overlap, void* stream);

/* target ingestion (data/__init__.py:9-27 detection_collate + the per-image slicing of
 * refinedet_multibox_loss.py:76-77): flat[total,5] = the step's ragged target tensors concatenated
 * (x1,y1,x2,y2,label), offsets[B+1] int32 (exclusive prefix sum of the per-image counts; device memory, or pinned
 * host memory read over PCIe -- 33 ints, no copy call) ->
 * truths[B,Gmax,4], labels[B,Gmax] (zero padded), gt_count[B]. */
RD_API int rd_pad_targets(const float* flat, const int* offsets, int B, int Gmax,
                   float* truths, float* labels, int* gt_count, void* stream);

/* hard-negative mining (refinedet_multibox_loss.py:117-123):
 *   loss_c [B,P] float (values at positives are ignored = treated as 0, :117),
 *   pos [B,P] uint8;  neg_out [B,P] uint8 = 1 for the num_neg = min(ratio*num_pos, P-1)
 *   largest entries of each row (ties: lower index first); num_pos_out [B] int32. */
RD_API int rd_hnm_select(const float* loss_c, const unsigned char* pos, int B, int P,
                  int negpos_ratio, unsigned char* neg_out, int* num_pos_out, void* stream);

/* ---- loss tail of RefineDetMultiBoxLoss (refinedet_multibox_loss.py:96-139) ---------- */
/* Per row r of conf[rows,C] (rows = B*P, 2 <= C <= 128):
 *   lse_out[r] = log(sum_c exp(conf[r,c]))         (log_sum_exp, box_utils.py:208-216; max-subtracted)
 *   ce_out[r]  = lse_out[r] - conf[r, conf_t[r]]   (the mining loss of :114 = the cross-entropy term of :130);
 *                NaN when conf_t[r] is outside [0, C) (the reference's gather raises there)
 *   pos_out[r] = conf_t[r] > 0, and — when arm_conf[rows,2] (LOGITS) is given — not
 *                softmax(arm_conf[r])[1] <= theta  (:96-101)                                   */
RD_API int rd_conf_loss(const float* conf, const long long* conf_t, const float* arm_conf, float theta,
                 long long rows, int C, float* ce_out, float* lse_out, unsigned char* pos_out,
                 void* stream);
/* loss_l = sum_{pos} SmoothL1(loc - loc_t) / N (:105-110), loss_c = sum_{pos|neg} ce / N (:126-130),
 * N = sum(num_pos) (:134); all three are written as device scalars; N < 1 gives zeros (:135-136).
 * fp64 accumulation in a fixed order (deterministic). */
RD_API size_t rd_multibox_loss_workspace_bytes(int B);
RD_API int rd_multibox_loss_reduce(const float* loc, const float* loc_t, const float* ce,
                 const unsigned char* pos, const unsigned char* neg, const int* num_pos, int B, int P,
                 void* workspace, size_t workspace_bytes, float* loss_l, float* loss_c, float* n_out,
                 void* stream);
/* The whole criterion forward (refinedet_multibox_loss.py:62-138) as ONE call: refine_match / match over the
 * padded batch, the per-anchor confidence loss with the ARM-theta gate, hard-negative mining, both reductions
 * and the division by N -- six kernels chained with programmatic dependent launch, no host involvement.
 *   inputs : truths / labels / gt_count / priors / arm_loc / threshold / label_mode as rd_refine_match,
 *            loc_data [B,P,4], conf_data [B,P,C] (the branch's predictions), arm_conf_gate [B,P,2] LOGITS or NULL
 *            (ODM criterion: :96-101), theta, negpos_ratio
 *   outputs: loc_t [B,P,4], conf_t [B,P] int64, ce / lse [B,P] float, pos / neg [B,P] uint8, num_pos [B] int32,
 *            losses [3] float = { loss_l, loss_c, N }  (zeros when N < 1).  Everything the backward needs.   */
RD_API size_t rd_multibox_criterion_workspace_bytes(int B, int P, int Gmax);
RD_API int rd_multibox_criterion(const float* truths, const float* labels, const int* gt_count, const float* priors,
                 const float* arm_loc, const float* loc_data, const float* conf_data, const float* arm_conf_gate,
                 int B, int P, int C, int Gmax, float threshold, float v0, float v1, int label_mode, float theta,
                 int negpos_ratio, void* workspace, size_t workspace_bytes, float* loc_t, long long* conf_t,
                 float* ce, float* lse, unsigned char* pos, unsigned char* neg, int* num_pos, float* losses,
                 void* stream);
/* backward of the two losses (what autograd derives from :105-138): with g_l = *grad_loss_l,
 * g_c = *grad_loss_c (device scalars, either may be NULL = 0) and N = *n_dev,
 *   grad_conf[r,:] = (softmax(conf[r,:]) - onehot(conf_t[r])) * g_c / N   on pos|neg rows, 0 elsewhere
 *   grad_loc[r,:]  = clamp(loc[r,:] - loc_t[r,:], -1, 1) * g_l / N        on pos rows, 0 elsewhere
 * Either output may be NULL.  Every element of a given output is written. */
RD_API int rd_multibox_loss_backward(const float* loc, const float* loc_t, const float* conf,
                 const long long* conf_t, const float* lse, const unsigned char* pos,
                 const unsigned char* neg, const float* grad_loss_l, const float* grad_loss_c,
                 const float* n_dev, long long rows, int C, float* grad_loc, float* grad_conf,
                 void* stream);

/* ---- the ARM and the ODM criterion of a training step as ONE call (train_refinedet.py:252-253 calls
 * refinedet_multibox_loss.py:50-139 twice on the same predictions and targets) ------------------------------------
 * A criterion STATE is one device allocation of rd_criterion_state_bytes(B, P, Gmax) bytes (256-byte aligned) that
 * holds everything a forward leaves for its backward; rd_criterion_state_layout writes the byte offsets of its
 * regions: { loc_t f32[B,P,4], conf_t i64[B,P], ce f32[B,P], lse f32[B,P], pos u8[B,P], neg u8[B,P],
 * num_pos i32[B], losses f32[3] = { loss_l, loss_c, N }, workspace } into offsets[9].
 *
 * rd_multibox_criterion_pair: the ARM criterion (use_ARM=False: match against the priors, labels by arm_label_mode,
 * predictions arm_loc / arm_conf[B,P,2]) on `stream`, the ODM criterion (use_ARM=True: refine_match against
 * decode(arm_loc), predictions odm_loc / odm_conf[B,P,C], positives gated by softmax(arm_conf)[1] <= theta) on
 * `side_stream` -- forked from and joined back into `stream` with events, so the two chains of six kernels run
 * CONCURRENTLY and the caller sees one stream-ordered call.  side_stream == NULL or == stream: one after the other.
 * rd_multibox_loss_backward_pair: both backward kernels the same way.  g_* are device scalars (NULL = 0), grad_*
 * outputs may be NULL (that gradient is not needed). */
RD_API size_t rd_criterion_state_bytes(int B, int P, int Gmax);
RD_API int rd_criterion_state_layout(int B, int P, int Gmax, size_t* offsets /* [9] */);
RD_API int rd_multibox_criterion_pair(const float* truths, const float* labels, const int* gt_count,
                 const float* priors, const float* arm_loc, const float* arm_conf, const float* odm_loc,
                 const float* odm_conf, int B, int P, int C, int Gmax, float arm_threshold, float odm_threshold,
                 float v0, float v1, int arm_label_mode, float theta, int arm_negpos_ratio, int odm_negpos_ratio,
                 void* arm_state, void* odm_state, size_t state_bytes, void* stream, void* side_stream);
RD_API int rd_multibox_loss_backward_pair(const float* arm_loc, const float* arm_conf, const float* odm_loc,
                 const float* odm_conf, const void* arm_state, const void* odm_state, int B, int P, int C, int Gmax,
                 const float* g_arm_l, const float* g_arm_c, const float* g_odm_l, const float* g_odm_c,
                 float* grad_arm_loc, float* grad_arm_conf, float* grad_odm_loc, float* grad_odm_conf,
                 void* stream, void* side_stream);

#ifdef __cplusplus
}
#endif
#endif  /* REFINEDET_B200_H_ */
