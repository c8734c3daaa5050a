/*
 * frb200.h — C ABI of libfrb200.so, the B200 (sm_100a) identification-stage kernels.
 *
 * The reference (sin0235/FaceRecognition) is pure Python and has no FFI of its own
 * (SURVEY.md §8b): its seam is a handful of Python methods.  Each entry point below
 * names the reference call site whose arithmetic it replaces; the Python mirror of
 * those methods (facerecognition_b200/) binds these symbols with ctypes
 * (INTEGRATION.md shows the stub).  Paths are relative to the reference root.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no torch / C++ types.
 *  - Every function returns an frb_status (0 = ok, <0 = error); nothing throws.
 *    frb_last_error() returns a thread-local description of the last failure.
 *  - "device" pointers are CUDA device pointers owned by the caller; `stream` is a
 *    cudaStream_t passed as void* (NULL = legacy default stream).  All launches are
 *    asynchronous on that stream; no function synchronises unless it says so.
 *  - Workspaces are caller-owned device scratch; query the size first.
 *  - Row indices are int64 and are offset by `idx_base`, so a rank that owns rows
 *    [base, base+N) of a sharded gallery reports GLOBAL ids.
 *  - Ties are always resolved to the LOWEST row index (the reference's stable sort /
 *    OpenCV's strict '<' scan keep the first row).
 *  - There is no CPU fallback: without a CUDA device of compute capability 10.x the
 *    calls fail with FRB_ERR_CUDA / FRB_ERR_UNSUPPORTED.
 */
#ifndef FRB200_H
#define FRB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum frb_status {
    FRB_OK = 0,
    FRB_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, k out of range ...) */
    FRB_ERR_UNSUPPORTED = -2, /* valid request outside what the kernels implement            */
    FRB_ERR_CUDA = -3,        /* CUDA runtime / driver error (see frb_last_error)             */
    FRB_ERR_WORKSPACE = -4    /* workspace missing or too small                               */
} frb_status;

typedef enum frb_dtype { FRB_F32 = 0, FRB_BF16 = 1, FRB_F16 = 2 } frb_dtype;

/* How a query row x is scaled before the dot product. */
typedef enum frb_qnorm {
    FRB_QNORM_NONE = 0,   /* use as given                                                       */
    FRB_QNORM_CLAMP = 1,  /* x / max(||x||, 1e-12)  — torch F.normalize, inference/extract_embeddings.py:381,434 */
    FRB_QNORM_EPS = 2     /* x / (||x|| + 1e-8)     — inference/recognition_engine.py:302, web_app.py:540       */
} frb_qnorm;

/* How a (query, gallery-row) score is formed from the dot product. */
typedef enum frb_score {
    FRB_SCORE_IP = 0,         /* inner product of the (optionally normalised) query with the stored row:
                                 faiss.IndexFlatIP.search, inference/recognition_engine.py:304;
                                 np.dot(E, P.T), notebooks/evaluate_arcface_kaggle.ipynb:618             */
    FRB_SCORE_REF_COSINE = 1  /* cosine_similarity(), inference/recognition_engine.py:41-63: 0 if either
                                 norm is 0; raw dot if both norms are within 1e-3 of 1; else dot/(na*nb).
                                 Needs q_norms and g_norms.                                            */
} frb_score;

#define FRB_MAX_K 64

/* ---- library -------------------------------------------------------------------------- */
int frb_version(void);
const char *frb_last_error(void);
/* sm_count / cc_major / cc_minor of the current CUDA device. */
int frb_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ---- measurement --------------------------------------------------------------------- */
/* Kernel ids for the profiling counters below. */
typedef enum frb_kernel {
    FRB_K_COSINE_TC = 0,   /* cosine_tc_kernel   (bf16 tcgen05 similarity + top-k) */
    FRB_K_COSINE_SIMT = 1, /* cosine_simt_kernel (fp32 FFMA similarity + top-k)    */
    FRB_K_LBP_HIST = 2,    /* lbp_hist_kernel                                       */
    FRB_K_CHISQ = 3,       /* chisq_kernel                                          */
    FRB_K_BGR2GRAY = 4,    /* bgr2gray_kernel                                       */
    FRB_K_COSINE_GEMV = 5, /* cosine_gemv_kernel (1..4 queries, HBM-bound row streaming) */
    FRB_K_RESIZE = 6,      /* resize_linear_kernel (cv2.resize INTER_LINEAR, optionally fused with BGR2GRAY) */
    FRB_K_CHISQ_FILTER = 7, /* chisq_filter_kernel (fp16 tcgen05 candidate filter in front of the chi-square scan) */
    FRB_K_COUNT = 8
} frb_kernel;

/* When enabled, every launch of the kernels listed above is bracketed by a CUDA event pair on the
 * launching stream (what bench.py's roofline line is computed from).  Off by default. */
int frb_profile_enable(int on);
/* Waits for the recorded launches of `kernel` to finish, returns their summed device time and count,
 * and clears the record. */
int frb_profile_read(int kernel, float *total_ms, int *launches);

/* ---- cosine path (K1) ------------------------------------------------------------------ */

/* ||x_r||_2 per row, fp32.  Replaces the two np.linalg.norm calls of cosine_similarity
 * (inference/recognition_engine.py:52-53), hoisted out of the per-pair loop. */
int frb_row_norms_f32(const float *x_dev, int64_t rows, int dim, float *out_norms_dev, void *stream);

/* Row-wise L2 normalisation with the reference's two conventions, writing fp32 or bf16.
 * (out_dtype FRB_F32, FRB_BF16 or FRB_F16).  mode = FRB_QNORM_CLAMP | FRB_QNORM_EPS (NONE = plain cast).  Replaces
 * extract_embeddings.py:622-623 (index rows), web_app.py:549 (db rows), :540 (query). */
int frb_normalize_rows(const float *x_dev, int64_t rows, int dim, int mode, void *out_dev, int out_dtype,
                       void *stream);

/* Scratch bytes for frb_cosine_topk at this problem size. */
size_t frb_cosine_topk_workspace_bytes(int64_t n_query, int64_t n_gallery, int dim, int gallery_dtype, int k);

/* Fused similarity + top-k: for each of n_query fp32 queries [n_query, dim] find the k best
 * rows of gallery [n_gallery, dim] (fp32 or bf16, row-major).  The n_query x n_gallery score
 * matrix never reaches HBM.
 *   gallery_dtype FRB_F32 : exact fp32 FFMA kernel (<=1e-5 of the reference's numpy scores).
 *   gallery_dtype FRB_BF16: tcgen05 tensor-core kernel (queries normalised in fp32, rounded to
 *                           bf16 in the prologue; fp32 accumulate in TMEM; <=1e-3).
 *   gallery_dtype FRB_F16 : the same kernel on fp16 operands (11 significant bits: a unit-norm row pair's inner product
 *                           moves by <= 1.1e-3 worst case, ~1e-5 typically) — the first pass of the exact-fp32 path,
 *                           see frb_cosine_rescore_topk.
 *   q_norms_dev / g_norms_dev : fp32 row norms, required for FRB_SCORE_REF_COSINE, else NULL.
 *   out_scores_dev [n_query, k] fp32 descending; out_idx_dev [n_query, k] int64 (idx_base + row);
 *   slots beyond n_gallery hold (-inf, -1) like faiss.
 * Replaces: RecognitionEngine.recognize_with_db loop+sort (inference/recognition_engine.py:277-289),
 * recognize_with_faiss search (:304), web_app.py:545-554, the notebooks' np.dot+argsort. */
int frb_cosine_topk(const float *queries_dev, int64_t n_query, const void *gallery_dev, int gallery_dtype,
                    int64_t n_gallery, int dim, const float *q_norms_dev, const float *g_norms_dev,
                    int score_mode, int qnorm_mode, int k, int64_t idx_base, float *out_scores_dev,
                    int64_t *out_idx_dev, void *workspace_dev, size_t workspace_bytes, void *stream);

/* The bf16 path of frb_cosine_topk for queries that are ALREADY L2-normalised and rounded to bf16 ([n_query, dim],
 * e.g. by frb_normalize_rows(..., FRB_BF16)): the tensor-core kernel reads them in place and no prologue runs.  This is
 * what a sharded search calls after all-gathering the ranks' normalised query slices (half the NVLink bytes of the fp32
 * batch, and no rank normalises another rank's queries).  Same results as frb_cosine_topk with FRB_QNORM_CLAMP/EPS on
 * the fp32 rows those bf16 rows came from.  Workspace: frb_cosine_topk_workspace_bytes(..., FRB_BF16, k). */
int frb_cosine_topk_bf16q(const void *queries_bf16_dev, int64_t n_query, const void *gallery_bf16_dev, int64_t n_gallery,
                          int dim, int k, int64_t idx_base, float *out_scores_dev, int64_t *out_idx_dev,
                          void *workspace_dev, size_t workspace_bytes, void *stream);

/* Exact fp32 top-k from a tensor-core first pass.  cand_idx / cand_approx [n_query, kp] are the (local row, score) lists
 * frb_cosine_topk returned for a unit-norm fp16 (or bf16) copy of the gallery with kp > k.  Each listed row is re-scored
 * in fp32 under score_mode (queries as that rule takes them: normalised rows for FRB_SCORE_IP, raw rows + norms for
 * FRB_SCORE_REF_COSINE), the best k are written (ties -> lowest row, ids + idx_base), and the query FAILS if its list
 * cannot be PROVEN complete.  Proof: a row outside the list has a first-pass score <= a_min (the list's smallest), hence
 * an exact score <= x + eps_rel * |x| + 1e-6 with x = a_min + eps_abs; that must be below the k-th exact score.
 *   eps_abs >= the worst |true cosine - first-pass score|: 1.1e-3 for fp16 unit vectors (2 * 2^-11 by Cauchy-Schwarz,
 *              subnormal tails and the fp32 accumulation included), 8.0e-3 for bf16;
 *   eps_rel >= the worst relative gap between the reference's score and the true cosine: 2.001e-3 for
 *              cosine_similarity()'s raw-dot branch (both norms within 1e-3 of 1), 0 for FRB_SCORE_IP on unit rows.
 * *fail_count_dev is incremented per failing query and, if fail_flags_dev (int32 [n_query], may be NULL) is given, the
 * query's flag is set to 1 (0 otherwise): the caller reruns frb_cosine_topk on the fp32 gallery for exactly those
 * queries.  Same result as recognize_with_db / np.dot + argsort (inference/recognition_engine.py:277-289,
 * notebooks/evaluate_arcface_kaggle.ipynb:618,713), reached through tcgen05. */
int frb_cosine_rescore_topk(const float *queries_dev, int64_t n_query, const float *gallery_f32_dev, int64_t n_gallery,
                            int dim, const float *q_norms_dev, const float *g_norms_dev, int score_mode,
                            const int64_t *cand_idx_dev, const float *cand_approx_dev, int kp, int k, float eps_abs,
                            float eps_rel, int64_t idx_base, float *out_scores_dev, int64_t *out_idx_dev,
                            int *fail_count_dev, int *fail_flags_dev, void *stream);

/* Gallery builder (K4): out[g] = mean(emb[order[offsets[g] .. offsets[g+1])]) / (||mean|| + 1e-8), float32 sums in
 * the given order then a division by the count (numpy's mean(axis=0)); groups with no rows come out all-zero.
 * emb f32 [M, dim]; order i64 [M] = sample indices grouped by identity; offsets i64 [n_groups + 1].
 * out_f32 [n_groups, dim] and/or out_bf16 [n_groups, dim] (either may be NULL; the bf16 copy is the K1 gallery).
 * Replaces: compute_prototypes (inference/extract_embeddings.py:573-584), the per-identity mean + renorm of
 * extract_embedding_for_folder (:758-760) / build_db (:808-831) and RecognitionEngine.add_to_db
 * (inference/recognition_engine.py:413-419). */
int frb_group_mean_renorm(const float *emb_dev, const int64_t *order_dev, const int64_t *offsets_dev, int64_t n_groups,
                          int dim, float *out_f32_dev, void *out_bf16_dev, void *stream);

/* Merge R candidate lists per query (e.g. one per GPU after the all-gather, or one per
 * gallery chunk): cand_scores [R, n_query, k], cand_idx [R, n_query, k] -> best k.
 * largest != 0: descending score (cosine); largest == 0: ascending distance (LBPH).
 * Entries with idx < 0 are padding.  Ties -> lowest idx, so the result does not depend on R. */
int frb_topk_merge(const float *cand_scores_dev, const int64_t *cand_idx_dev, int n_lists, int64_t n_query,
                   int k, int largest, float *out_scores_dev, int64_t *out_idx_dev, void *stream);

/* Same merge over lists that are not back to back: list l's scores start at cand_scores + l * score_list_stride
 * and its ids at cand_idx + l * idx_list_stride (strides in ELEMENTS, each >= n_query * k).  This is the layout
 * ONE all-gather of a packed per-rank record {ids i64 [n_query, k] | scores f32 [n_query, k]} leaves behind
 * (facerecognition_b200/sharded.py), so the cross-GPU exchange is a single collective. */
int frb_topk_merge_strided(const float *cand_scores_dev, const int64_t *cand_idx_dev, int64_t score_list_stride,
                           int64_t idx_list_stride, int n_lists, int64_t n_query, int k, int largest,
                           float *out_scores_dev, int64_t *out_idx_dev, void *stream);

/* BGR -> gray front end of the LBPH path: bgr u8 [n_pixels, 3] (interleaved, as cv2 frames) -> gray u8 [n_pixels],
 * OpenCV's 8-bit fixed point (3735 B + 19235 G + 9798 R + 2^14) >> 15, bit-exact with cv2.cvtColor(COLOR_BGR2GRAY)
 * of OpenCV 4.13 over all 2^24 colours.  Replaces the cv2.cvtColor calls that feed LBPH
 * (models/lbphmodel/train_lbph_script.py:72, web_app.py:475,486). */
int frb_bgr2gray_u8(const uint8_t *bgr_dev, int64_t n_pixels, uint8_t *out_gray_dev, void *stream);

/* Resize front end of the LBPH path: src u8 [count, src_rows, src_cols, channels] (channels 1 or 3, interleaved) ->
 * dst u8 [count, dst_rows, dst_cols, channels], bit-exact with cv2.resize(img, (dst_cols, dst_rows)) at its default
 * INTER_LINEAR (OpenCV 4.x 11-bit fixed point, including its switch to the 2x2 INTER_AREA average when both axes
 * halve exactly).  to_gray = 1 (channels must be 3) fuses cv2.cvtColor(..., COLOR_BGR2GRAY) behind it and writes
 * dst u8 [count, dst_rows, dst_cols]: the resized colour image never reaches memory.  Replaces
 * `cv2.resize(image, target_size)` + `cv2.cvtColor` in _preprocess_image_for_lbph
 * (models/lbphmodel/train_lbph_script.py:67-72) and web_app.py:472-475,484-486; destination sides <= 4096. */
int frb_resize_linear_u8(const uint8_t *src_dev, int64_t count, int src_rows, int src_cols, int channels,
                         uint8_t *dst_dev, int dst_rows, int dst_cols, int to_gray, void *stream);

/* ---- cross-GPU exchange fused with the merge, over NVLink peer memory ------------------------------------- */
/* One context per rank (process).  frb_exchange_create cudaMallocs this rank's buffer — `world` record slots of
 * max_query x max_k candidates, double-buffered by step parity, plus one flag per (rank, 128-query CTA) — and
 * returns its CUDA IPC handle (FRB_IPC_HANDLE_BYTES bytes).  The host all-gathers the handles (any transport) and
 * passes all `world` of them, in rank order, to frb_exchange_open, which maps the peers' buffers.
 * frb_exchange_topk_merge is then ONE kernel per step: each CTA stores its queries' local candidates into every
 * rank's buffer (peer stores), release-stores the step's epoch into the peers' flags, acquire-waits for the same
 * CTA of every rank, and merges (ties -> lowest global id).  All ranks must call it in lockstep.  The epoch is kept in
 * device memory and advanced by the kernel itself, so the step (local search + exchange) can be captured once into a
 * CUDA graph and replayed.  A rank that does not arrive within ~10 s is counted (frb_exchange_status) and the kernel
 * returns with that step's output undefined; frb_exchange_reset, called by every rank between two host barriers,
 * restarts the epochs — e.g. after one rank raised an error and skipped a step.
 * Replaces the all-gather + frb_topk_merge_strided pair of facerecognition_b200/sharded.py; no reference
 * counterpart (the reference is single-device). */
#define FRB_EXCHANGE_MAX_WORLD 8
#define FRB_IPC_HANDLE_BYTES 64
typedef struct frb_exchange frb_exchange;
int frb_exchange_create(int world, int rank, int64_t max_query, int max_k, frb_exchange **out,
                        unsigned char *ipc_handle_out);
int frb_exchange_open(frb_exchange *ex, const unsigned char *ipc_handles /* world x FRB_IPC_HANDLE_BYTES */);
int frb_exchange_destroy(frb_exchange *ex);
int frb_exchange_topk_merge(frb_exchange *ex, const float *local_scores_dev, const int64_t *local_idx_dev,
                            int64_t n_query, int k, int largest, float *out_scores_dev, int64_t *out_idx_dev,
                            void *stream);
/* timeouts: kernels of this rank that gave up waiting for a peer since creation / the last reset; epoch: steps
 * finished.  Synchronises with the device. */
int frb_exchange_status(frb_exchange *ex, int *timeouts, unsigned *epoch);
/* Clears this rank's flags, epoch and timeout count (device-synchronising).  Collective by convention: every rank calls
 * it after a barrier that guarantees no exchange kernel is in flight anywhere, and barriers again before the next step. */
int frb_exchange_reset(frb_exchange *ex);
/* Test hook: `world` contexts created in ONE process on one device act as the ranks of a single launch
 * (local_* and out_* are [world, n_query, k]); exercises the store / flag / wait / merge protocol without
 * a second GPU. */
int frb_exchange_emulate(frb_exchange *const *ranks, int world, const float *local_scores_dev,
                         const int64_t *local_idx_dev, int64_t n_query, int k, int largest, float *out_scores_dev,
                         int64_t *out_idx_dev, void *stream);

/* ---- LBPH path (K2, K3) ---------------------------------------------------------------- */

/* LBP codes, OpenCV elbp_ semantics (radius 1, 8 neighbours, float32 bilinear diagonals,
 * (t > c) || |t - c| < FLT_EPSILON): images u8 [count, rows, cols] -> codes u8
 * [count, rows-2, cols-2].  Exposed for bit-exact parity checks of the code stage. */
int frb_lbp_codes_u8(const uint8_t *images_dev, int64_t count, int rows, int cols, int radius, int neighbors,
                     uint8_t *out_codes_dev, void *stream);

/* LBP codes + spatial histogram in one pass (codes never reach HBM): images u8
 * [count, rows, cols] -> integer cell histograms u16 [count, grid_x*grid_y*256], cell
 * (i, j) at row i*grid_x + j, cell size floor((cols-2)/grid_x) x floor((rows-2)/grid_y).
 * *out_cell_px (host, may be NULL) receives the pixel count per cell; OpenCV's float view
 * is (float)count * (float)(1.0 / cell_px).
 * Replaces the per-image body of cv2.face LBPH train()/predict() (models/lbphmodel/train_lbph.py:35,
 * models/lbphmodel/inference_lbph.py:5): elbp + spatial_histogram. */
int frb_lbp_hist_u8(const uint8_t *images_dev, int64_t count, int rows, int cols, int radius, int neighbors,
                    int grid_x, int grid_y, uint16_t *out_hist_dev, int *out_cell_px, void *stream);

/* The same pass writing u8 counts [count, grid_x*grid_y*256] — the gallery form the chi-square kernels stream
 * (frb_chisq_topk_g8, frb_chisq_top1_filtered_g8): valid when a cell has <= 255 pixels (100x100 and 112x112 faces under
 * the 8x8 grid: 144 / 169), else FRB_ERR_UNSUPPORTED.  Half the bytes written per trained face, and no conversion pass
 * between LBPH train()/update() and the gallery. */
int frb_lbp_hist_u8_counts8(const uint8_t *images_dev, int64_t count, int rows, int cols, int radius, int neighbors,
                            int grid_x, int grid_y, uint8_t *out_hist_dev, int *out_cell_px, void *stream);

/* u16 counts -> u8 counts, n elements (every count must be <= 255): adopting a u16 histogram matrix as a u8 gallery. */
int frb_counts_u16_to_u8(const uint16_t *src_dev, int64_t n, uint8_t *dst_dev, void *stream);

/* idx[i] = idx[i] >= 0 ? table[idx[i]] : idx[i], in place, n entries: local gallery rows -> the caller's global row ids
 * (LBPH galleries that mix image sizes keep one histogram group per cell size; facerecognition_b200/lbph.py). */
int frb_index_remap(int64_t *idx_dev, int64_t n, const int64_t *table_dev, int64_t table_len, void *stream);

size_t frb_chisq_topk_workspace_bytes(int64_t n_query, int64_t n_gallery, int hist_len, int k);

/* Chi-square (HISTCMP_CHISQR_ALT) nearest neighbours: for each query histogram find the k
 * gallery rows with the smallest d = 2 * sum_j (h_j - q_j)^2 / (h_j + q_j), h = count/cell_px.
 *   q_hist_dev  u16 [n_query, hist_len], counts with q_cell_px pixels per cell
 *   gallery_dev u16 [n_gallery, hist_len], counts with g_cell_px pixels per cell
 *   out_dist_dev fp32 [n_query, k] ascending; out_idx_dev int64 [n_query, k] (idx_base + row),
 *   (+inf, -1) beyond n_gallery.  First row wins ties (StandardCollector's strict '<').
 * Replaces the compareHist scan of cv2.face LBPH predict() (web_app.py:587,
 * models/lbphmodel/evaluate_lbph.py:32, models/lbphmodel/threshold_lbph.py:48). */
int frb_chisq_topk(const uint16_t *q_hist_dev, int64_t n_query, int q_cell_px, const uint16_t *gallery_dev,
                   int64_t n_gallery, int hist_len, int g_cell_px, int k, int64_t idx_base,
                   float *out_dist_dev, int64_t *out_idx_dev, void *workspace_dev, size_t workspace_bytes,
                   void *stream);

/* All distances: out_dist_dev fp32 [n_query, n_gallery] (parity checks, threshold sweeps). */
int frb_chisq_dist(const uint16_t *q_hist_dev, int64_t n_query, int q_cell_px, const uint16_t *gallery_dev,
                   int64_t n_gallery, int hist_len, int g_cell_px, float *out_dist_dev, void *stream);

/* The same two scans over a gallery stored as u8 counts [n_gallery, hist_len] — valid when g_cell_px <= 255 (a cell of
 * a 100x100 or 112x112 face under the 8x8 grid has 144 / 169 pixels, so every count fits a byte) and hist_len % 16 == 0.
 * Same arithmetic on the same integers (queries stay u16, as frb_lbp_hist_u8 writes them).  A 1M-histogram gallery is
 * 16.4 GB instead of 32.8 GB (OpenCV keeps 65.5 GB of float32 for it), and this is the form the tensor-core filter below
 * streams.  It does NOT halve the time of a single predict(): the u16 scan runs at the HBM roof (100k rows in 0.51 ms,
 * 6.4 TB/s) and the u8 scan, with half the bytes, hits the FP32/MUFU limit at about the same time (0.50 ms, 3.25 TB/s). */
int frb_chisq_topk_g8(const uint16_t *q_hist_dev, int64_t n_query, int q_cell_px, const uint8_t *gallery_dev,
                      int64_t n_gallery, int hist_len, int g_cell_px, int k, int64_t idx_base,
                      float *out_dist_dev, int64_t *out_idx_dev, void *workspace_dev, size_t workspace_bytes,
                      void *stream);
int frb_chisq_dist_g8(const uint16_t *q_hist_dev, int64_t n_query, int q_cell_px, const uint8_t *gallery_dev,
                      int64_t n_gallery, int hist_len, int g_cell_px, float *out_dist_dev, void *stream);

/* ---- batched LBPH predict through the tensor cores (candidate filter + exact re-score) --------------------------- */
/* The N predicts of the reference's evaluation loops (models/lbphmodel/evaluate_lbph.py:31-33, threshold_lbph.py:47-50,
 * web_app.py:587 per frame) as ONE call that returns, for every query, exactly what frb_chisq_topk_g8(k = 1) returns —
 * same fp32 distance bits, same row, first row wins ties — while almost every (query, row) pair is decided by an fp16
 * GEMM on tcgen05 instead of the 65 k-flop exact formula:
 *   sum_j (g_j - q_j)^2 / (g_j + q_j) = sum g + sum q - 4 sum_j f(g_j, q_j),  f(a, b) = ab / (a + b);
 *   f ~ <u(a), v(b)> with rank-8 fp16 feature tables (frb_chisq_filter_tables), so sum_j f is an inner product of
 *   length 8 * hist_len; the table error bounds |approx - exact| per query for EVERY row, rows outside the resulting
 *   window around the best approximate score cannot win, and the survivors are re-scored with the exact kernel's own
 *   arithmetic.  A query whose survivor list overflows is answered by the plain exact scan inside the same call.
 * Needs equal cell sizes on both sides (cell_px <= 255, u8 gallery), hist_len % 16 == 0, hist_len <= 16384, and every
 * count (gallery and query) <= cell_px — all a cell of cell_px pixels can hold; the feature tables and their error bound
 * cover exactly 0..cell_px, larger counts give undefined candidates (not checked per call).  Rows need not sum to the
 * same total (LBPH rows always do): the kernel ranks by sum_j f - (row total) / 4.
 *   stats_dev  int32 [4] or NULL, incremented: [0] queries answered by the exact fallback, [1] survivors re-scored,
 *              [2] raw candidates appended by the filter kernel, [3] re-scored rows whose filter score missed the exact
 *              one by more than the bound (audit of every survivor; such a query is re-answered by the exact scan).
 *   approx_scores_dev  fp32 [n_query, n_gallery] or NULL: every approximate sum_j f (tests / calibration only).
 *   workspace_dev  >= frb_chisq_filter_workspace_bytes(...) bytes, 256-byte aligned; histograms 16-byte aligned.
 * The first call for a (device, cell_px) builds and uploads the tables with blocking copies; later calls are
 * stream-ordered and capturable. */
size_t frb_chisq_filter_workspace_bytes(int64_t n_query, int64_t n_gallery, int hist_len);
int frb_chisq_top1_filtered_g8(const uint16_t *q_hist_dev, int64_t n_query, const uint8_t *gallery_dev, int64_t n_gallery,
                               int hist_len, int cell_px, int64_t idx_base, float *out_dist_dev, int64_t *out_idx_dev,
                               int *stats_dev, float *approx_scores_dev, void *workspace_dev, size_t workspace_bytes,
                               void *stream);
/* The filter's tables for one cell size (host pointers): u_f16 / v_f16 [256, 8] fp16 bit patterns (gallery / query side,
 * rows above cell_px zero), emax [256] = max_a |<u(a), v(b)> - f(a, b)| and absmax [256] = max_a sum_m |u_m(a) v_m(b)|
 * per query count b.  Exposed so that tests can check the bound against the float64 oracle (oracle/chisq_filter.py). */
int frb_chisq_filter_tables(int cell_px, uint16_t *u_f16, uint16_t *v_f16, float *emax, float *absmax);

#ifdef __cplusplus
}
#endif
#endif /* FRB200_H */
