// cosine_api.cu — frb_cosine_topk: argument checks, workspace carving, kernel selection by shape and gallery dtype.
//   few queries (<= 2 bf16, <= 16 fp32), dim a multiple of 256 <= 512 -> cosine_gemv.cu (HBM-bound row streaming)
//   FRB_F32  gallery -> cosine_simt.cu (exact fp32 FFMA kernel)
//   FRB_BF16 gallery -> cosine_tc.cu   (tcgen05 / TMEM / TMA kernel)
// One kernel per (shape class, dtype); nothing falls back to anything else.
#include "frb_common.cuh"

namespace frb {
int64_t simt_chunks(int64_t n_query, int64_t n_gallery, int64_t *tiles_per_chunk);
int launch_cosine_simt(const float *queries, int64_t nq, const void *gallery, int gallery_dtype, int64_t ng, int dim,
                       const float *q_norms, const float *g_norms, int score_mode, int k, int64_t idx_base,
                       float *cand_scores, int64_t *cand_idx, int64_t tiles_per_chunk, int64_t chunks, cudaStream_t st);

size_t cosine_tc_workspace_bytes(int64_t n_query, int64_t n_gallery, int dim, int k);
int launch_cosine_tc(const float *queries, const void *queries_bf16, int64_t nq, const void *gallery_bf16, int64_t ng, int dim,
                     int qnorm_mode, int k, int64_t idx_base, float *out_scores, int64_t *out_idx, void *ws, size_t ws_bytes,
                     cudaStream_t st, int op_dtype);

bool gemv_applicable(int64_t n_query, int dim, int gallery_dtype);
int64_t gemv_ctas(int64_t n_gallery, int64_t *rows_per_cta);
int launch_cosine_gemv(const void *queries, int64_t nq, const void *gallery, int gallery_dtype, int64_t ng, int dim, const float *q_norms,
                       const float *g_norms, int score_mode, int k, int64_t idx_base, float *cs, int64_t *ci, int *cnt, int64_t rpc,
                       int64_t ctas, cudaStream_t st);

struct GemvPlan {
    int64_t rows_per_cta, ctas;
    size_t q_bytes, cnt_bytes, idx_bytes, score_bytes;
};

static GemvPlan gemv_plan(int64_t nq, int64_t ng, int dim, int k)
{
    GemvPlan p;
    p.ctas = gemv_ctas(ng > 0 ? ng : 1, &p.rows_per_cta);
    const size_t n = (size_t)p.ctas * (size_t)nq * (size_t)k;
    p.q_bytes = align_up((size_t)nq * dim * sizeof(float), 256);  // fp32 or bf16 copy of the (normalised) queries
    p.cnt_bytes = align_up((size_t)nq * sizeof(int), 256);
    p.idx_bytes = align_up(n * sizeof(int64_t), 256);
    p.score_bytes = align_up(n * sizeof(float), 256);
    return p;
}

struct SimtPlan {
    int64_t tiles_per_chunk, chunks;
    size_t qn_bytes, idx_bytes, score_bytes;
};

static SimtPlan simt_plan(int64_t nq, int64_t ng, int dim, int k, bool normalise)
{
    SimtPlan p;
    p.chunks = simt_chunks(nq, ng > 0 ? ng : 1, &p.tiles_per_chunk);
    size_t n = (size_t)p.chunks * (size_t)nq * (size_t)k;
    p.qn_bytes = normalise ? align_up((size_t)nq * dim * sizeof(float), 256) : 0;
    p.idx_bytes = align_up(n * sizeof(int64_t), 256);
    p.score_bytes = align_up(n * sizeof(float), 256);
    return p;
}
}  // namespace frb

using namespace frb;

extern "C" {

size_t frb_cosine_topk_workspace_bytes(int64_t n_query, int64_t n_gallery, int dim, int gallery_dtype, int k)
{
    if (n_query <= 0 || dim <= 0 || k <= 0) return 0;
    if (gemv_applicable(n_query, dim, gallery_dtype)) {
        GemvPlan p = gemv_plan(n_query, n_gallery, dim, k);
        return p.q_bytes + p.cnt_bytes + p.idx_bytes + p.score_bytes;
    }
    if (gallery_dtype == FRB_BF16 || gallery_dtype == FRB_F16) return cosine_tc_workspace_bytes(n_query, n_gallery, dim, k);
    SimtPlan p = simt_plan(n_query, n_gallery, dim, k, true);
    return p.qn_bytes + p.idx_bytes + p.score_bytes;
}

int frb_cosine_topk(const float *queries, int64_t n_query, const void *gallery, int gallery_dtype, int64_t n_gallery,
                    int dim, const float *q_norms, const float *g_norms, int score_mode, int qnorm_mode, int k,
                    int64_t idx_base, float *out_scores, int64_t *out_idx, void *workspace, size_t workspace_bytes,
                    void *stream)
{
    FRB_CHECK_ARG(n_query >= 0 && n_gallery >= 0, "frb_cosine_topk: n_query=%lld n_gallery=%lld", (long long)n_query,
                  (long long)n_gallery);
    FRB_CHECK_ARG(dim > 0 && (dim % 8) == 0, "frb_cosine_topk: dim=%d must be a positive multiple of 8", dim);
    FRB_CHECK_ARG(k >= 1 && k <= FRB_MAX_K, "frb_cosine_topk: k=%d (1..%d)", k, FRB_MAX_K);
    FRB_CHECK_ARG(gallery_dtype == FRB_F32 || gallery_dtype == FRB_BF16 || gallery_dtype == FRB_F16, "frb_cosine_topk: gallery_dtype=%d",
                  gallery_dtype);
    FRB_CHECK_ARG(score_mode == FRB_SCORE_IP || score_mode == FRB_SCORE_REF_COSINE, "frb_cosine_topk: score_mode=%d",
                  score_mode);
    FRB_CHECK_ARG(qnorm_mode >= FRB_QNORM_NONE && qnorm_mode <= FRB_QNORM_EPS, "frb_cosine_topk: qnorm_mode=%d", qnorm_mode);
    if (n_query == 0) return FRB_OK;
    FRB_CHECK_ARG(queries && out_scores && out_idx, "frb_cosine_topk: null pointer");
    FRB_CHECK_ARG(n_gallery == 0 || gallery, "frb_cosine_topk: null gallery");
    FRB_CHECK_ARG(((uintptr_t)queries & 15) == 0 && ((uintptr_t)gallery & 15) == 0,
                  "frb_cosine_topk: queries and gallery must be 16-byte aligned");
    if (score_mode == FRB_SCORE_REF_COSINE) {
        FRB_CHECK_ARG(q_norms && (g_norms || n_gallery == 0), "frb_cosine_topk: FRB_SCORE_REF_COSINE needs q_norms and g_norms");
        FRB_CHECK_ARG(qnorm_mode == FRB_QNORM_NONE, "frb_cosine_topk: FRB_SCORE_REF_COSINE takes raw queries (qnorm NONE)");
    }
    size_t need = frb_cosine_topk_workspace_bytes(n_query, n_gallery, dim, gallery_dtype, k);
    if (!workspace || workspace_bytes < need) {
        set_error("frb_cosine_topk: workspace %zu B < %zu B", workspace_bytes, need);
        return FRB_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (gallery_dtype != FRB_F32 && score_mode != FRB_SCORE_IP) {
        set_error("frb_cosine_topk: bf16 / fp16 galleries hold pre-normalised rows; only FRB_SCORE_IP is implemented");
        return FRB_ERR_UNSUPPORTED;
    }

    if (gemv_applicable(n_query, dim, gallery_dtype)) {
        GemvPlan p = gemv_plan(n_query, n_gallery, dim, k);
        char *ws = (char *)workspace;
        int *cnt = (int *)(ws + p.q_bytes);
        int64_t *ci = (int64_t *)(ws + p.q_bytes + p.cnt_bytes);
        float *cs = (float *)(ws + p.q_bytes + p.cnt_bytes + p.idx_bytes);
        const void *q_use = queries;
        if (gallery_dtype == FRB_BF16) {
            // the same rounding the tcgen05 path applies: normalise in fp32, round to bf16
            int rc = normalize_rows_impl(queries, n_query, dim, qnorm_mode, ws, FRB_BF16, nullptr, nullptr, st);
            if (rc != FRB_OK) return rc;
            q_use = ws;
        } else if (qnorm_mode != FRB_QNORM_NONE) {
            int rc = frb_normalize_rows(queries, n_query, dim, qnorm_mode, ws, FRB_F32, stream);
            if (rc != FRB_OK) return rc;
            q_use = ws;
        }
        int rc = launch_cosine_gemv(q_use, n_query, gallery, gallery_dtype, n_gallery, dim, q_norms, g_norms, score_mode, k, idx_base, cs, ci,
                                    cnt, p.rows_per_cta, p.ctas, st);
        if (rc != FRB_OK) return rc;
        return topk_merge_compact(cs, ci, cnt, p.ctas * k, n_query, k, /*largest=*/1, out_scores, out_idx, st);
    }

    if (gallery_dtype == FRB_BF16 || gallery_dtype == FRB_F16) {
        return launch_cosine_tc(queries, nullptr, n_query, gallery, n_gallery, dim, qnorm_mode, k, idx_base, out_scores, out_idx,
                                workspace, workspace_bytes, st, gallery_dtype);
    }

    SimtPlan p = simt_plan(n_query, n_gallery, dim, k, true);
    char *ws = (char *)workspace;
    float *qn = (float *)ws;
    int64_t *ci = (int64_t *)(ws + p.qn_bytes);
    float *cs = (float *)(ws + p.qn_bytes + p.idx_bytes);
    const float *q_use = queries;
    if (qnorm_mode != FRB_QNORM_NONE) {
        // element-wise x / denom exactly as the reference does before its dot products
        int rc = frb_normalize_rows(queries, n_query, dim, qnorm_mode, qn, FRB_F32, stream);
        if (rc != FRB_OK) return rc;
        q_use = qn;
    }
    int rc = launch_cosine_simt(q_use, n_query, gallery, FRB_F32, n_gallery, dim, q_norms, g_norms, score_mode, k, idx_base,
                                cs, ci, p.tiles_per_chunk, p.chunks, st);
    if (rc != FRB_OK) return rc;
    return frb_topk_merge(cs, ci, (int)p.chunks, n_query, k, /*largest=*/1, out_scores, out_idx, stream);
}

int frb_cosine_topk_bf16q(const void *queries_bf16, int64_t n_query, const void *gallery_bf16, int64_t n_gallery, int dim, int k,
                          int64_t idx_base, float *out_scores, int64_t *out_idx, void *workspace, size_t workspace_bytes, void *stream)
{
    const char *fn = "frb_cosine_topk_bf16q";
    FRB_CHECK_ARG(n_query >= 0 && n_gallery >= 0, "%s: n_query=%lld n_gallery=%lld", fn, (long long)n_query, (long long)n_gallery);
    FRB_CHECK_ARG(dim > 0 && (dim % 8) == 0, "%s: dim=%d must be a positive multiple of 8", fn, dim);
    FRB_CHECK_ARG(k >= 1 && k <= FRB_MAX_K, "%s: k=%d (1..%d)", fn, k, FRB_MAX_K);
    if (n_query == 0) return FRB_OK;
    FRB_CHECK_ARG(queries_bf16 && out_scores && out_idx, "%s: null pointer", fn);
    FRB_CHECK_ARG(n_gallery == 0 || gallery_bf16, "%s: null gallery", fn);
    FRB_CHECK_ARG(((uintptr_t)queries_bf16 & 15) == 0 && ((uintptr_t)gallery_bf16 & 15) == 0, "%s: queries and gallery must be 16-byte aligned", fn);
    const size_t need = frb_cosine_topk_workspace_bytes(n_query, n_gallery, dim, FRB_BF16, k);
    if (!workspace || workspace_bytes < need) {
        set_error("%s: workspace %zu B < %zu B", fn, workspace_bytes, need);
        return FRB_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (gemv_applicable(n_query, dim, FRB_BF16)) {
        GemvPlan p = gemv_plan(n_query, n_gallery, dim, k);
        char *ws = (char *)workspace;
        int *cnt = (int *)(ws + p.q_bytes);
        int64_t *ci = (int64_t *)(ws + p.q_bytes + p.cnt_bytes);
        float *cs = (float *)(ws + p.q_bytes + p.cnt_bytes + p.idx_bytes);
        int rc = launch_cosine_gemv(queries_bf16, n_query, gallery_bf16, FRB_BF16, n_gallery, dim, nullptr, nullptr, FRB_SCORE_IP, k, idx_base,
                                    cs, ci, cnt, p.rows_per_cta, p.ctas, st);
        if (rc != FRB_OK) return rc;
        return topk_merge_compact(cs, ci, cnt, p.ctas * k, n_query, k, /*largest=*/1, out_scores, out_idx, st);
    }
    return launch_cosine_tc(nullptr, queries_bf16, n_query, gallery_bf16, n_gallery, dim, FRB_QNORM_NONE, k, idx_base, out_scores, out_idx,
                            workspace, workspace_bytes, st, FRB_BF16);
}

}  // extern "C"
