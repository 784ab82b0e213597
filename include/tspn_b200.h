/*
 * tspn_b200.h — C ABI of the B200-native TSPN tracklet-pair stage.
 *
 * One shared library (csrc/libtspn_b200.so, sm_100a only) behind plain pointers and sizes:
 * no torch types, no C++ types, no exceptions across the boundary.  The host side that
 * mirrors the reference's Python interface (package tspn_b200) binds these with ctypes
 * and passes tensor.data_ptr(); INTEGRATION.md shows the stub a maintainer of the
 * reference would add.  Each entry point names the reference interface it replaces
 * (paths relative to the reference repository root).
 *
 * Conventions
 *   - every `d_*` pointer is DEVICE memory owned by the caller; the library never
 *     allocates persistent memory and never frees caller memory; scratch space is a
 *     caller-provided workspace whose size comes from the matching *_workspace_bytes();
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden
 *     synchronisation, no host reads of device data;
 *   - return value: TSPN_OK (0) or a negative TSPN_E* code; tspn_last_error() returns
 *     the thread-local message of the last failure;
 *   - a device that is not compute capability 10.x makes every compute entry return
 *     TSPN_EARCH: there is no fallback path by design;
 *   - base pointers must be 16-byte aligned (TMA, 128-bit accesses) -> TSPN_EALIGN.
 *
 * Batch layout (several videos per launch)
 *   A batch of V videos is described by the *video table*: (table_rows + 1) rows of
 *   TSPN_VT_COLS int64, built on the host by tspn_build_video_table() from the per-video
 *   tracklet and frame counts, then copied to the device by the caller.  Rows [0, V) are the
 *   videos, rows [V, table_rows) are empty videos (N = 0), and row `table_rows` is the SENTINEL:
 *   its offset columns hold the batch's totals (tracklets, pairs, geo floats, work items, boxes,
 *   scores).  Every entry point takes `num_videos` = table_rows and reads the true totals from the
 *   sentinel on the device, so the scalar totals a caller passes are only UPPER BOUNDS (they size
 *   the grids): a launch sequence recorded once for a capacity (a CUDA graph) serves every batch
 *   that fits it - ragged batches need no re-capture (ABI 5).  With N = tracklets, T = frames,
 *   P = N(N-1) ordered pairs of a video:
 *     boxes   float [sum N*Tb][4]   (x1,y1,x2,y2) inclusive pixels, row of tracklet n at
 *                                   box_off + n*Tb, Tb = T rounded up to 8 (pad = 0)
 *     span    int32 [sum N][2]      [pstart, pend)   (lib/modeling/trajectory.py:21-22)
 *     cls     float [sum N][C]      track_cls_logits (lib/dataset/vrdataset.py:61-83)
 *     motion  float [sum N][4000]   4 BoW blocks of 1000 (lib/dataset/vrdataset.py:227-236)
 *     geo     float [sum P*8*Tp]    per video [P][8][Tp], Tp = T rounded up to 4 (pad = 0)
 *     pair row p of (s,o): s*(N-1) + o - [o>s]  (order of lib/modeling/predict.py:133-140)
 */
#ifndef TSPN_B200_H
#define TSPN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSPN_ABI_VERSION 6

/* error codes */
#define TSPN_OK 0
#define TSPN_EBADARG (-1)
#define TSPN_ESHAPE (-2)
#define TSPN_EALIGN (-3)
#define TSPN_ECUDA (-4)
#define TSPN_EARCH (-5)

/* video table columns */
#define TSPN_VT_COLS 12
#define TSPN_VT_N 0         /* tracklets */
#define TSPN_VT_T 1         /* frames */
#define TSPN_VT_TP 2        /* geo row length: T rounded up to 4 */
#define TSPN_VT_TB 3        /* box row length: T rounded up to 8 */
#define TSPN_VT_TRK_OFF 4   /* first tracklet row (span / cls / motion) */
#define TSPN_VT_PAIR_OFF 5  /* first pair row */
#define TSPN_VT_GEO_OFF 6   /* first geo float */
#define TSPN_VT_ITEM_OFF 7  /* first work item of the pair-geometry kernel */
#define TSPN_VT_BOX_OFF 8   /* first box (in boxes, i.e. units of 16 bytes) */
#define TSPN_VT_SCORE_OFF 9 /* first relationness score (sum of N*N) */

/* totals[] written by tspn_build_video_table */
#define TSPN_TOT_COLS 10
#define TSPN_TOT_TRACKLETS 0
#define TSPN_TOT_PAIRS 1
#define TSPN_TOT_GEO_FLOATS 2
#define TSPN_TOT_ITEMS 3
#define TSPN_TOT_BOXES 4
#define TSPN_TOT_SCORES 5
#define TSPN_TOT_MAX_N 6
#define TSPN_TOT_MAX_T 7
#define TSPN_TOT_GEO_CHUNK 8 /* frames per work item of the pair-geometry kernel: tspn_geo_chunk(max T) */
#define TSPN_TOT_MAX_CHUNKS 9 /* max over videos of ceil(T / geo_chunk): chunk slots per pair in the workspace */

/* Work items of the pair-geometry kernel: (video, subject, group of TSPN_GEO_OBJ_GROUP other
 * tracklets, chunk of tspn_geo_chunk(max T of the batch) frames), numbered chunk-fastest; the table's
 * TSPN_VT_ITEM_OFF column and totals[TSPN_TOT_ITEMS] count them. */
#ifndef TSPN_GEO_OBJ_GROUP
#define TSPN_GEO_OBJ_GROUP 32
#endif

/* geometry channels of geo[P][8][Tp] ([SPEC] s2) */
#define TSPN_GEO_CHANNELS 8
#define TSPN_MOTION_DIM 4000
#define TSPN_MOTION_BLOCK 1000
#define TSPN_REL_BINS 500
#define TSPN_REL_DIM 3000

/* flags */
#define TSPN_VIOU_FULL 0       /* volumes over each full span: evaluation/common.py:65-106 (V2);
                                  equals trajectory.py:127-141 (V1) when all spans are equal */
#define TSPN_VIOU_CLIPPED 1    /* volumes over the overlap only: association.py:35-48 (V3) */
#define TSPN_GEO_DENSE_CTAS 2   /* tspn_pair_geo_viou: 1024 threads per SM with a 2-stage TMA ring and 64 registers
                                  instead of the default ~512 threads per SM, 3 stages, 103 registers (same
                                  results bit for bit; measured slower, kept for A/B - DESIGN.md section 4.1) */
/* tspn_pair_geo_viou runs three kernels: PRE (per-tracklet volumes), MAIN (the pair kernel), POST (per-pair
 * vIoU / tIoU).  With none of these bits set all three run; a caller that wants to time the pair kernel alone,
 * or to overlap the phases, issues them as separate calls (same arguments).  Every (pair, chunk) has exactly one
 * writer - the CTA of that work item stores the chunk's fixed-point sums into the pair's chunk slot of the
 * workspace ([pair][max_chunks][3] uint64) and POST adds a pair's slots in ascending chunk order - so nothing is
 * zeroed and there are no global atomics: PRE may run on another stream concurrently with MAIN; POST needs both.
 * MAIN writes d_overlap; POST writes d_viou / d_tiou. */
#define TSPN_GEO_PHASE_PRE 8
#define TSPN_GEO_PHASE_MAIN 16
#define TSPN_GEO_PHASE_POST 32
/* MAIN as persistent CTAs (one per SM slot for the whole launch) that pull work items from a queue in the
 * workspace, instead of one CTA per work item: same results bit for bit.  Kernels issued concurrently on
 * other streams can then only co-reside with the pair kernel's CTAs - they can never take over an SM between
 * two of its CTAs and lock the next one out. */
#define TSPN_GEO_PERSISTENT 128
/* With TSPN_GEO_PERSISTENT: leave that many SM slots free ((n) << TSPN_GEO_RESERVE_SHIFT, n <= 255).  Kernels of
 * concurrent streams run several times slower beside a pair CTA than on an SM of their own; a handful of free SMs
 * lets a side branch finish under the pair kernel at a small cost to it (bench workload, 148 SMs: 8 free SMs =
 * +1 % pair-kernel time, -1.7 % step time). */
#define TSPN_GEO_RESERVE_SHIFT 16
#define TSPN_TOPK_KEEP_DIAGONAL 0    /* reference behaviour, ppn.py:84-85 (quirk Q1) */
#define TSPN_TOPK_EXCLUDE_DIAGONAL 1 /* survivors are real pairs (sparsify mode) */
#define TSPN_PREC_FP32_EXACT 0 /* CUDA cores, fixed k-ascending fma order: bit-reproducible */
#define TSPN_PREC_TENSOR 1     /* tcgen05 tensor cores: bf16 (or tf32 on fp32 storage) operands,
                                  fp32 accumulation in TMEM */

/* ---- library ------------------------------------------------------------------------ */
int tspn_version(void);
/* copies the calling thread's last error message; returns its length */
int tspn_last_error(char* buf, int len);
/* TSPN_OK when the current CUDA device is compute capability 10.x, else TSPN_EARCH */
int tspn_check_device(void);

/* ---- host helper: batch layout ------------------------------------------------------- */
/* table_host: [table_rows + 1][TSPN_VT_COLS] int64 (host), totals: [TSPN_TOT_COLS] int64 (host).
 * table_rows >= num_videos (0 = num_videos): the table is padded with empty videos up to table_rows, then
 * the sentinel row.  geo_chunk: 0 = tspn_geo_chunk(max T of the batch), else 512 / 1024 / 2048 (a capacity
 * bucket fixes the chunk its launches were recorded with).  Pure host arithmetic on sizes; no device access. */
int tspn_build_video_table(int num_videos, const int32_t* n_tracklets, const int32_t* n_frames,
                           int table_rows, int geo_chunk, int64_t* table_host, int64_t* totals);
/* 512, 1024 or 2048: the smallest chunk that covers max_t (2048 beyond); one CTA of chunk/4 threads
 * writes whole geometry rows where it can - HBM absorbs few wide store streams best */
int tspn_geo_chunk(int64_t max_t);

/* ---- a1: pair enumeration -------------------------------------------------------------
 * Replaces the h5 `pairs` table (lib/dataset/vrdataset.py:208; order of predict.py:133-140).
 * d_pairs: int64 [sum P][2] = (s, o) local tracklet indices. */
int tspn_enumerate_pairs(const int64_t* d_table, int num_videos, int64_t total_pairs,
                         int64_t* d_pairs, void* stream);

/* ---- a3/a4/a6/a7 + [SPEC] s2/s3: all-pairs per-frame geometry and trajectory vIoU ----------
 * Replaces cubic_iou/_intersect/_union (lib/modeling/trajectory.py:85-141) applied to all
 * ordered pairs of each video, viou (lib/evaluation/common.py:65-106) and _traj_iou
 * (lib/modeling/association.py:35-48, flag TSPN_VIOU_CLIPPED); adds the per-frame channels.
 * d_geo may be NULL (reductions only).  d_workspace: tspn_pair_geo_workspace_bytes() bytes
 * (per-tracklet volumes + per-(pair, chunk) fixed-point volume sums + the persistent kernel's work-item
 * queue), 16-byte aligned; max_chunks = totals[TSPN_TOT_MAX_CHUNKS] (an upper bound, like the other totals).  Outputs by phase (flags below): MAIN writes d_geo and d_overlap, POST writes d_viou and
 * d_tiou from the sums MAIN left in the workspace and the volumes of PRE. */
int64_t tspn_pair_geo_workspace_bytes(int64_t total_tracklets, int64_t total_pairs, int max_chunks);
int tspn_pair_geo_viou(const int64_t* d_table, int num_videos, int64_t total_items, int geo_chunk, int max_chunks,
                       int64_t total_tracklets, int64_t total_pairs, int64_t total_boxes,
                       const float* d_boxes, const int32_t* d_span,
                       float* d_geo, float* d_viou, float* d_tiou, int32_t* d_overlap,
                       int flags, void* d_workspace, void* stream);

/* ---- opt-in WINDOWED geometry layout --------------------------------------------------------------------
 * Every channel of a pair's dense row is zero outside the pair's temporal overlap window [a, b) and channel 7 is the
 * window's indicator, so the dense [P][8][Tp] tensor is mostly zeros (60 % on the bench workload).  The windowed
 * layout stores, per pair, only the frames [a & ~3, (b + 3) & ~3) (length Lw, a multiple of 4) of channels 0..6:
 * [7][Lw] floats at d_geo + d_geo_off[p] - bit-identical to the dense rows on those frames; d_overlap gives [a, b).
 * tspn_geo_window_offsets fills d_geo_off (int64 [total_pairs], floats, exclusive prefix sum of 7 * Lw in pair
 * order) and *d_total (the floats the batch needs; at most 7/8 of totals[TSPN_TOT_GEO_FLOATS]) from the tracklet
 * spans - once per batch upload, not per step.  tspn_pair_geo_viou_windowed = tspn_pair_geo_viou writing that layout:
 * same phases, same flags (TSPN_GEO_DENSE_CTAS is rejected), same vIoU / tIoU / overlap bit for bit; its MAIN phase is
 * a kernel of its own (csrc/geo_windowed.cu: a warp per pair walking only the window; TSPN_GEO_RESERVE_SHIFT counts
 * CTA slots of two per SM).  The dense layout stays the parity layout; the kernels that read stored rows
 * (tspn_assemble_features, tspn_assemble_relative, tspn_span_head, tspn_span_proposals) read the dense one. */
int tspn_geo_window_offsets(const int64_t* d_table, int num_videos, int64_t total_pairs, const int32_t* d_span,
                            int64_t* d_geo_off, int64_t* d_total, void* stream);
int tspn_pair_geo_viou_windowed(const int64_t* d_table, int num_videos, int64_t total_items, int geo_chunk,
                                int max_chunks, int64_t total_tracklets, int64_t total_pairs, int64_t total_boxes,
                                const float* d_boxes, const int32_t* d_span, float* d_geo, const int64_t* d_geo_off,
                                float* d_viou, float* d_tiou, int32_t* d_overlap, int flags, void* d_workspace,
                                void* stream);

/* ---- a4 matrix form: cubic_iou(bboxes1, bboxes2) -----------------------------------------
 * lib/modeling/trajectory.py:127-141.  b1 [n1][t][4], b2 [n2][t][4] -> out [n1][n2]. */
int tspn_cubic_iou(const float* d_b1, int n1, const float* d_b2, int n2, int t,
                   float* d_out, void* stream);

/* ---- a6 batched: viou over an explicit list of trajectory pairs ---------------------------
 * lib/evaluation/common.py:65-106 for every (a[i], b[i]); trajectories live in a pool:
 * trajectory j has its boxes at d_pool[d_traj_off[j] ...] for frames [span[j][0], span[j][1]). */
int tspn_viou_pairs(const float* d_pool, const int64_t* d_traj_off, const int32_t* d_traj_span,
                    const int32_t* d_a, const int32_t* d_b, int64_t n_pairs, int flags,
                    float* d_out, void* stream);

/* ---- N3: the same for evaluation (lib/evaluation/visual_relation_detection.py:8-36 calls viou for
 * every (prediction, ground truth) pair of equal triplet): per-trajectory volumes are summed once
 * (d_workspace: tspn_viou_pairs_workspace_bytes(n_traj) bytes), each pair then reads only its overlap
 * window; sums and the ratio stay in fp64, so for integer boxes d_out[i] is bit-identical to python's
 * float(v_overlap) / (v1 + v2 - v_overlap) and the host-side threshold decisions cannot flip.
 * TSPN_VIOU_CLIPPED gives association.py:35-48 (volumes over the overlap only) in fp64.
 * d_traj_len (int32 [n_traj], may be NULL = duration length): number of boxes of each trajectory.  The
 * reference sums a trajectory's volume over its whole box list (common.py:100-105), and association
 * emits relations whose lists are longer than their duration (shared trajectories extended by other
 * relations' merges); the overlap window still comes from the durations, so len >= duration is required. */
int64_t tspn_viou_pairs_workspace_bytes(int64_t n_traj);
int tspn_viou_pairs_f64(const float* d_pool, const int64_t* d_traj_off, const int32_t* d_traj_span,
                        const int32_t* d_traj_len, int64_t n_traj, const int32_t* d_a, const int32_t* d_b,
                        int64_t n_pairs, int flags, double* d_out, void* d_workspace, void* stream);

/* ---- a2: feature rows -------------------------------------------------------------------
 * L1-normalise each 1000-wide BoW block (lib/dataset/vrdataset.py:227-236 with
 * lib/utils/miscellaneous.py:32-35). */
int tspn_normalize_motion(const float* d_motion, int64_t n_tracklets, float* d_out, void* stream);
/* Compact host->device transport (batch.py packs these when the values allow it, losslessly): motion
 * histograms as u8 counts [n][4000] (normalised exactly like tspn_normalize_motion: the integer sum is
 * exact), boxes as u16 pixel coordinates [n_boxes][4] expanded to the fp32 layout the kernels read. */
int tspn_normalize_motion_u8(const uint8_t* d_motion, int64_t n_tracklets, float* d_out, void* stream);
int tspn_unpack_boxes_u16(const uint16_t* d_src, int64_t n_boxes, float* d_dst, void* stream);
/* The same with the boxes SPAN-PACKED: a tracklet only has boxes on [pstart, pend) (lib/modeling/trajectory.py:21-22;
 * nothing on the path reads a box outside its tracklet's span), so only those travel - tracklet n's at
 * d_packed[d_packed_off[n] ...] (int64 [total_tracklets], units of boxes, in tracklet order) - and the expansion
 * writes them at frames pstart .. of the dense row and zeros everywhere else.  total_tracklets is an upper bound
 * (the sentinel row carries the true count).  Per tracklet two encodings, chosen by the packer: RAW - one 8-byte slot
 * (4 x u16) per frame - or, with TSPN_PACKED_DELTA set in d_packed_off[n], DELTA - slot 0 = the first frame's 4 x u16,
 * then 4 x i8 per further frame (two frames per slot): each coordinate's difference to the previous frame, for
 * tracklets whose boxes never move by more than [-128, 127] pixels per frame (half the bytes).  The expansion of a
 * delta tracklet is an integer prefix sum over its frames: the boxes are the packer's integers exactly. */
#define TSPN_PACKED_DELTA (1ll << 62)
int tspn_unpack_boxes_spans(const int64_t* d_table, int num_videos, int64_t total_tracklets, const int32_t* d_span,
                            const int64_t* d_packed_off, const uint16_t* d_packed, float* d_dst, void* stream);
/* HOST side of that transport (plain CPU code, no device access; serving.host_batches_for packs pinned batches
 * with it).  tspn_host_pack_boxes_spans: one video's boxes [n][t][4] fp32
 * (host) -> slots from slot0 on in dst (the arena's u16 box field, dst_slots slots of 8 bytes in all), tracklet after
 * tracklet, frames [pstart, pend) only, raw or - allow_delta and every frame-to-frame difference in [-128, 127] - delta
 * coded; writes box_off[n] (host, with TSPN_PACKED_DELTA) and *slots_used.  A coordinate inside a span that is not an
 * integer in [0, 65535], a span outside [0, t] or more slots than the arena holds -> TSPN_ESHAPE. */
int tspn_host_pack_boxes_spans(const float* boxes, int n_tracklets, int n_frames, const int32_t* span, int allow_delta,
                               uint16_t* dst, int64_t dst_slots, int64_t slot0, int64_t* box_off, int64_t* slots_used);
/* Build rows [2C | 8000 motion | 3000 relative] (lib/dataset/vrdataset.py:219-243); the last
 * 3000 columns are the adaptive-average-pooled geometry ([SPEC] s4).  d_rows: global pair rows
 * to build (int64, NULL = all total_pairs rows in order).  Output row stride ld_feat floats
 * (>= 2C+11000, multiple of 4).  d_feat_bf16 (optional, may be NULL): same rows in bf16 with
 * stride ld_bf16 (multiple of 8) for the tensor-core predicate head.  max_frames =
 * totals[TSPN_TOT_MAX_T] (0 = unknown; validated, otherwise unused since ABI 4: the pooled channels are
 * read straight from global memory). */
int tspn_assemble_features(const int64_t* d_table, int num_videos, int64_t total_pairs, int max_frames,
                           const float* d_cls, int n_classes, const float* d_motion_norm,
                           const float* d_geo, const int32_t* d_overlap,
                           const int64_t* d_rows, int64_t n_rows,
                           float* d_feat, int64_t ld_feat,
                           void* d_feat_bf16, int64_t ld_bf16, void* stream);

/* ---- a8/a9: relationness scoring and top-K -------------------------------------------------
 * PPNHead.forward (lib/modeling/relpn/ppn.py:92-112): scores [sum N*N] (per video N x N,
 * diagonal included).  weights: sub_emb.0 [H][C],[H]; sub_emb.2 [C][H],[C]; obj_emb likewise
 * (state_dict order).  d_workspace: tspn_relationness_workspace_bytes(). */
int64_t tspn_relationness_workspace_bytes(int64_t total_tracklets, int n_classes, int hidden);
/* precision: TSPN_PREC_FP32_EXACT = CUDA cores, fixed k-ascending fma order (scores bit-identical to oracle/exact,
 * hence a bit-exact top-K); TSPN_PREC_TENSOR = the three contractions (two MLP layers, S O^T) as tcgen05.mma with
 * fp32 accumulators in TMEM, operands tf32 on the fp32 storage (scores within 1e-2 of the float64 definition; bf16
 * operands measure 1.4e-2 on the test weights - see csrc/relationness_tc.cu).  max_tracklets =
 * totals[TSPN_TOT_MAX_N]; tspn_relationness_tc_supported: N <= 256, C <= 128, H <= 128, H % 8 == 0. */
int tspn_relationness_tc_supported(int max_tracklets, int n_classes, int hidden);
int tspn_relationness(const int64_t* d_table, int num_videos, int64_t total_tracklets, int max_tracklets,
                      const float* d_cls, int n_classes, int hidden,
                      const float* d_sub_w0, const float* d_sub_b0,
                      const float* d_sub_w2, const float* d_sub_b2,
                      const float* d_obj_w0, const float* d_obj_b0,
                      const float* d_obj_w2, const float* d_obj_b2,
                      float* d_scores, int precision, void* d_workspace, void* stream);
/* PPN._forward_test (lib/modeling/relpn/ppn.py:79-90): per video the first min(K, N*N)
 * flat indices s*N+o in descending score order, ties to the lower index ([SPEC] s6).
 * d_topk_idx int64 [V][K] (-1 beyond K_eff), d_topk_score float [V][K],
 * d_topk_row int64 [V][K] global pair row of (s,o) or -1 for s==o / padding (may be NULL). */
int tspn_topk_pairs(const int64_t* d_table, int num_videos, const float* d_scores, int k,
                    int flags, int64_t* d_topk_idx, float* d_topk_score, int64_t* d_topk_row,
                    void* stream);

/* a8 + a9 in two launches instead of three: the embedding kernel, then one CTA per video that computes the
 * video's N x N scores from the embeddings in shared memory, writes them and selects the top K from the keys it
 * has just produced.  Same outputs, bit for bit, as tspn_relationness followed by tspn_topk_pairs (k == 0: scores
 * are NOT produced - use tspn_relationness).  TSPN_PREC_FP32_EXACT: supported when max_tracklets^2 <= 8192
 * (max_tracklets = totals[TSPN_TOT_MAX_N]); otherwise use the two entry points above.  TSPN_PREC_TENSOR: see
 * tspn_relationness_tc_supported (videos of more than 128 tracklets take a third launch for the top-K). */
int tspn_relationness_topk_supported(int max_tracklets, int n_classes);
int tspn_relationness_topk(const int64_t* d_table, int num_videos, int64_t total_tracklets, int max_tracklets,
                           const float* d_cls, int n_classes, int hidden,
                           const float* d_sub_w0, const float* d_sub_b0,
                           const float* d_sub_w2, const float* d_sub_b2,
                           const float* d_obj_w0, const float* d_obj_b0,
                           const float* d_obj_w2, const float* d_obj_b2,
                           float* d_scores, int k, int flags, int precision, int64_t* d_topk_idx,
                           float* d_topk_score, int64_t* d_topk_row, void* d_workspace, void* stream);

/* ---- a14: predicate classifier -----------------------------------------------------------
 * RelationPredictor.forward (lib/modeling/model.py:76-88): y = sigmoid(x W^T + b).
 * x [m][ld_x] fp32 (or bf16 when x_is_bf16), W [r][f] fp32, y [m][r].
 * TSPN_PREC_TENSOR needs the packed weights of tspn_pack_predicate_weights. */
int64_t tspn_predicate_packed_bytes(int n_predicates, int feature_dim);
int tspn_pack_predicate_weights(const float* d_w, int n_predicates, int feature_dim,
                                void* d_packed, void* stream);
int64_t tspn_predicate_workspace_bytes(int64_t m, int feature_dim, int n_predicates, int precision);
int tspn_predicate_head(const void* d_x, int x_is_bf16, int64_t ld_x, int64_t m, int feature_dim,
                        const float* d_w, const void* d_w_packed, const float* d_bias,
                        int n_predicates, float* d_y, int precision, void* d_workspace,
                        void* stream);

/* ---- a14 decomposed (TSPN_PREC_TENSOR, features built on the GPU) --------------------------------------
 * The classifier is linear in the feature row [cls_s | cls_o | motion_s | motion_o | relative], so
 *   x W^T = A_s[subject] + A_o[object] + rel(pair) W_rel^T
 * with per-TRACKLET terms A_s = [cls | motion_norm] W_s^T, A_o = [cls | motion_norm] W_o^T (W_s / W_o: the
 * columns of rel_predictor.weight that multiply the subject / object blocks).  The [rows, F] feature matrix
 * then never has to exist: 2 x 3000 bytes per scored pair instead of 2 x (2C + 11000).
 *   tspn_tracklet_rows      [n][ld] bf16 rows [cls | L1-normalised motion | 0] (motion fp32 or u8 counts)
 *   tspn_predicate_head_affine   y = act(x Wp^T + bias + row_bias[row]) on tcgen05; Wp from
 *                           tspn_pack_predicate_weights; bias / row_bias may be NULL; flags TSPN_AFFINE_RAW
 *                           skips the sigmoid (used for the tracklet terms: Wp = W_s, then Wp = W_o)
 *   tspn_assemble_relative  per scored row: the pooled 3000-wide relative block in bf16 ([SPEC] s4) and the
 *                           bias row A_s[s] + A_o[o] gathered from d_terms_subject / d_terms_object [n_tracklets][R]
 *                           (padding rows: zeros).  Workspace of the head: tspn_predicate_workspace_bytes. */
#define TSPN_AFFINE_RAW 1
/* The call runs on a side stream underneath a kernel that keeps every SM occupied (the pair-geometry kernel:
 * one CTA with 134 KB of shared memory and 53 K registers per SM): a 3-stage TMA ring (<= 80 KB) and at most 4
 * K-splits, so that its CTAs co-reside with that kernel's instead of each waiting for - and then holding - a
 * whole SM.  Same results as without the flag up to the split-K summation order. */
#define TSPN_AFFINE_BACKGROUND 2
int tspn_tracklet_rows(const float* d_cls, int n_classes, const void* d_motion, int motion_is_u8,
                       int64_t n_tracklets, void* d_out_bf16, int64_t ld, void* stream);
int tspn_predicate_head_affine(const void* d_x, int x_is_bf16, int64_t ld_x, int64_t m, int feature_dim,
                               const void* d_w_packed, const float* d_bias, const float* d_row_bias,
                               int64_t ld_row_bias, int n_outputs, float* d_y, int flags, void* d_workspace,
                               void* stream);
int tspn_assemble_relative(const int64_t* d_table, int num_videos, int64_t total_pairs, int max_frames,
                           const float* d_geo, const int32_t* d_overlap, const int64_t* d_rows, int64_t n_rows,
                           void* d_rel_bf16, int64_t ld_rel, const float* d_terms_subject,
                           const float* d_terms_object, int n_outputs, float* d_row_bias, void* stream);

/* ---- a11/a12 + [SPEC] s5: temporal-span head ---------------------------------------------
 * DPNHead.forward (lib/modeling/relpn/dpn.py:55-73): Conv1d(k3,p1) -> ReLU -> Conv1d(k1).
 * x rows are gathered: pair i reads row d_rows[i] - row_base of x (NULL = identity, negative =
 * padding -> zeros), each row [cin][ld_t], rows row_stride floats apart; out [k][a2][t].
 * conv_w [cin][cin][3], pred_w [a2][cin]. */
int64_t tspn_span_head_workspace_bytes(int64_t k, int cin, int t, int a2, int precision);
int tspn_span_head(const float* d_x, const int64_t* d_rows, int64_t row_base, int64_t row_stride,
                   int64_t ld_t, int64_t k, int cin, int t,
                   const float* d_conv_w, const float* d_conv_b,
                   const float* d_pred_w, const float* d_pred_b, int a2,
                   float* d_out, int precision, void* d_workspace, void* stream);
/* anchors of anchor_generator.py:48-104 + decode: reg [k][2a][t] -> spans int32 [k][n_loc*a][2] */
int tspn_span_num_locations(int t, float stride);
int tspn_span_decode(const float* d_reg, int64_t k, int n_anchors, int t,
                     const float* d_sizes, float stride, int32_t* d_spans, void* stream);
/* tspn_span_head (fp32 exact order) + tspn_span_decode fused: the head is evaluated only at the
 * n_loc anchor columns floor(l*stride) the decode reads, the [k][2a][t] regressions never reach
 * memory.  Same arguments as the two calls it replaces (dpn.py:55-73 + anchor_generator.py:48-104),
 * bit-identical spans int32 [k][n_loc*a][2]. */
int tspn_span_proposals(const float* d_x, const int64_t* d_rows, int64_t row_base, int64_t row_stride,
                        int64_t ld_t, int64_t k, int cin, int t,
                        const float* d_conv_w, const float* d_conv_b,
                        const float* d_pred_w, const float* d_pred_b, int n_anchors,
                        const float* d_sizes, float stride, int32_t* d_spans, void* stream);

/* ---- a12 + [SPEC] s8: span suppression and top-n (RelNMS) -----------------------------------------
 * What RelNMS (lib/modeling/relpn/rel_nms.py:6-15: nms_threshold 0.5, top_k_proposals =
 * RELPN.DPN.NUM_DURATION_PROPOSALS = 64, lib/config/defaults.py:62) is meant to do; its forward is a stub in the
 * reference, so the rule is [SPEC] (oracle/heads.py:select_spans), integers only, bit-exact:
 *   candidate i of a pair = decoded span [s_i, e_i) in the decode's order (location major, anchor minor);
 *   rank key q_i = floor(2^15 * |span_i ^ W| / |span_i v W|), W = the pair's temporal overlap window; ties to the
 *   lower i; greedy: keep the best live candidate, drop every live candidate whose temporal IoU with it exceeds
 *   nms_threshold (inter * 1024 > round(1024 * thr) * union), until n_keep are kept or none is left.
 * d_cand int32 [n_rows][ld_cand]: row r holds its candidates as (start, end) pairs.  Two ways to describe the rows:
 *   - d_table != NULL: row r scores global pair row d_rows[r] (NULL = r; negative = padding -> count 0); it has
 *     locations(T_video) * n_anchors candidates and, unless d_windows is given, its window comes from the tracklet
 *     spans d_span (so the call does not depend on the all-pairs kernel); max_frames = totals[TSPN_TOT_MAX_T];
 *   - d_table == NULL: every row has n_cand candidates and the window d_windows[r] (int32 [n_rows][2]).
 * d_out: int16 [n_rows][n_keep][2] with TSPN_SPANS_I16 (frames < 65536 always required), else int32; kept spans
 * in keep order, zero padded; d_counts int32 [n_rows] (may be NULL). */
#define TSPN_SPANS_I16 1
int tspn_span_select(const int64_t* d_table, int num_videos, int max_frames, const int32_t* d_span,
                     const int64_t* d_rows, int64_t n_rows, const int32_t* d_windows,
                     const int32_t* d_cand, int64_t ld_cand, int n_cand, int n_anchors, float stride, int n_keep,
                     float nms_threshold, int flags, void* d_out, int32_t* d_counts, void* stream);

/* ---- surviving pairs: relative block + bias row + span proposals, recomputed from the boxes ------------
 * For every row the top-K kept (d_rows: global pair rows, [V][rows_per_video], -1 = padding) this produces
 * exactly what tspn_assemble_relative and tspn_span_proposals produce from the stored geometry rows of
 * tspn_pair_geo_viou - bit for bit (same per-frame device code, same summation orders) - but from the boxes:
 * it does not depend on the all-pairs kernel and is shaped to co-reside with it (128 threads, <= 96
 * registers, 16 KB of shared memory per 512 frames of max_frames + 8 KB), so the heads of the K survivors run on a
 * side stream underneath the HBM-bound all-pairs kernel instead of re-reading 64 KB per row behind it.
 * d_spans (may be NULL: no span head) int32 [n_rows][ld_spans]: row r holds [locations(T_v) * A][2] frame
 * bounds of its video, zero-filled up to ld_spans >= locations(max_frames) * 2A.  Span weights as for
 * tspn_span_proposals with Cin = 8.  tspn_survivor_rows_supported: n_anchors == 4 and the tile of
 * max_frames frames fits shared memory; otherwise use the two stored-row entry points. */
/* d_terms_subject / d_terms_object / d_row_bias may all be NULL: the bias rows then come from
 * tspn_gather_pair_terms (d_row_bias [n_rows][n_outputs] = A_s[subject] + A_o[object], zeros for padding rows;
 * d_rows NULL = all pair rows in order), so that the recomputation need not wait for the per-tracklet terms. */
int tspn_gather_pair_terms(const int64_t* d_table, int num_videos, const int64_t* d_rows, int64_t n_rows,
                           const float* d_terms_subject, const float* d_terms_object, int n_outputs,
                           float* d_row_bias, void* stream);
int tspn_survivor_rows_supported(int max_frames, int n_anchors);
int tspn_survivor_rows(const int64_t* d_table, int num_videos, int max_frames, const float* d_boxes,
                       const int32_t* d_span, const int64_t* d_rows, int64_t n_rows, int64_t rows_per_video,
                       void* d_rel_bf16, int64_t ld_rel, const float* d_terms_subject,
                       const float* d_terms_object, int n_outputs, float* d_row_bias,
                       const float* d_conv_w, const float* d_conv_b, const float* d_pred_w,
                       const float* d_pred_b, int n_anchors, const float* d_sizes, float stride,
                       int32_t* d_spans, int64_t ld_spans, void* stream);

/* ---- N1: predict.py:66-117 post-processing ------------------------------------------------
 * per video: top `topk_per_pair` predicates per scored row (predict.py:70-73), then top
 * `topk_per_video` overall (predict.py:76-81), both descending with ties to the lower index;
 * one 32-byte record per kept triplet: {score f32, s_cls, pred, o_cls, s_tid, o_tid, start,
 * end : i32} (start/end = the pair's temporal overlap window, read from d_overlap [P][2] - or, when that
 * is NULL, derived from the tracklet spans d_span, so that the records need not wait for the pair kernel).
 * d_logits [n_rows][r]; row i scores global pair row d_rows[i] (NULL = identity, -1 = padding);
 * video v owns rows [d_row_video_off[v], d_row_video_off[v+1]) (NULL = its pair rows).
 * flags: TSPN_POST_MIRROR_Q4 reproduces predict.py:89's wrong-row object class.
 * d_records int32 [V][topk_per_video][8] (unused slots: score 0, ids -1), d_counts int32 [V]. */
#define TSPN_POST_MIRROR_Q4 1
int64_t tspn_postprocess_workspace_bytes(int64_t n_rows, int topk_per_pair);
int tspn_postprocess(const int64_t* d_table, int num_videos, const float* d_logits,
                     const int64_t* d_rows, const int64_t* d_row_video_off, int64_t n_rows,
                     int n_predicates, const float* d_cls, int n_classes, const int32_t* d_overlap,
                     const int32_t* d_span, int topk_per_pair, int topk_per_video, int flags,
                     int32_t* d_records, int32_t* d_counts, void* d_workspace, void* stream);

/* ---- the exchange step over peer memory (multi-GPU serving) ---------------------------------------
 * The path's one exchange - every rank's per-video triplet records to every rank, what lib/utils/comm.py:48-88
 * (pickle + two all-gathers) would do - as plain stores into peer-mapped gather buffers over NVLink instead of a
 * collective launch per step.  Buffers (symmetric across ranks, allocated and exchanged by the caller, e.g.
 * torch.distributed._symmetric_memory): gather buffer int32 [ring][world][n_int32] and flags int32 [2][ring][world]
 * (zeroed) on every rank; d_peer_bufs / d_peer_flags: device arrays [world] of the ranks' base pointers (own rank
 * included); d_local_flags = this rank's flags; d_state: int32 [2 + ring], zeroed {error word, reserved, one CTA
 * counter per slot}.  `step` = 1, 2, 3, ... is the caller's count of exchanges (the same sequence on every rank; an
 * explicit argument because consecutive steps launched on different streams may execute in either order); step s uses
 * slot s % ring, and ring must exceed the steps in flight between a scatter and its release.
 *   tspn_records_scatter  producer, on the compute stream behind tspn_postprocess: waits for the slot's credits,
 *                         stores d_records into slot[rank] of every rank, publishes written[slot][rank] = s
 *   tspn_records_wait     consumer, before reading its gather buffer's slot: waits for written[slot][p] >= s, all p
 *   tspn_records_release  consumer, after the slot was read: credit[slot][rank] = s on every rank
 * Waits are bounded spins (max_spin polls of ~0.2 us): on expiry d_state[0] becomes non-zero (1 = credits, 2 =
 * records) and the kernel returns - check it on the host; the device is never left spinning. */
int tspn_records_scatter(const int32_t* d_records, int64_t n_int32, const uint64_t* d_peer_bufs,
                         const uint64_t* d_peer_flags, const int32_t* d_local_flags, int world, int rank, int ring,
                         int step, int32_t* d_state, int max_spin, void* stream);
int tspn_records_wait(const int32_t* d_local_flags, int world, int ring, int step, int32_t* d_state, int max_spin,
                      void* stream);
int tspn_records_release(const uint64_t* d_peer_flags, int world, int rank, int ring, int step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TSPN_B200_H */
