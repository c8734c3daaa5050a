// cosine_refine.cu — exact fp32 answers at tensor-core speed: re-score a short candidate list.
//
// For fp32 galleries the 1e-5 bar rules out 16-bit operands, but not a 16-bit FIRST PASS: the tcgen05 kernel ranks the
// gallery on a unit-norm fp16 (or bf16) copy and returns kp > k candidates per query; this kernel then computes the
// EXACT fp32 score of those kp rows under the reference's rule (cosine_similarity()'s three branches,
// inference/recognition_engine.py:52-63, or an inner product), keeps the best k (ties -> lowest row) and PROVES
// completeness: every row outside the list has a first-pass score <= a_min (the list's smallest), hence a true
// cosine <= x = a_min + eps_abs (fp16 keeps 11 significant bits: rounding two unit vectors moves their inner product
// by at most 2 * 2^-11 by Cauchy-Schwarz, + subnormal tails + fp32 accumulation = 1.1e-3; bf16: 8.0e-3), hence a
// reference score <= x + eps_rel |x| + 1e-6 (cosine_similarity()'s raw-dot branch returns cos * |q| |g| with both norms
// within 1e-3 of 1: eps_rel = 2.001e-3).  If that is below the k-th exact score no outside row can enter or tie;
// otherwise the query is counted in *fail_count (and flagged) and the caller reruns the exact kernel for it.  One CTA
// per query: warps take candidates round-robin, lanes split the dot product.
#include "frb_common.cuh"

namespace frb {

constexpr int kRfThreads = 128;

__device__ __forceinline__ float rf_ref_cosine(float dot, float na, float nb)
{
    if (na == 0.f || nb == 0.f) return 0.f;
    if (fabsf(na - 1.0f) < 1e-3f && fabsf(nb - 1.0f) < 1e-3f) return dot;
    return __fdiv_rn(dot, __fmul_rn(na, nb));
}

__global__ void __launch_bounds__(kRfThreads)
cosine_rescore_kernel(const float *__restrict__ queries, const float *__restrict__ gallery, int dim, const float *__restrict__ q_norms,
                      const float *__restrict__ g_norms, int score_mode, const int64_t *__restrict__ cand_idx,
                      const float *__restrict__ cand_approx, int kp, int k, float eps, float eps_rel, int64_t n_gallery,
                      int64_t idx_base, float *__restrict__ out_scores, int64_t *__restrict__ out_idx, int *__restrict__ fail_count,
                      int *__restrict__ fail_flags)
{
    __shared__ float s_exact[FRB_MAX_K];
    const int64_t q = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *qrow = queries + q * dim;
    const float qn = score_mode == FRB_SCORE_REF_COSINE ? q_norms[q] : 0.f;
    for (int j = warp; j < kp; j += kRfThreads / 32) {
        const int64_t row = cand_idx[q * kp + j];
        float acc = 0.f;
        if (row >= 0) {
            const float *g = gallery + row * dim;
            for (int d = lane * 4; d < dim; d += 128) {
                const float4 a = __ldg(reinterpret_cast<const float4 *>(qrow + d));
                const float4 b = __ldg(reinterpret_cast<const float4 *>(g + d));
                acc = fmaf(a.x, b.x, acc);
                acc = fmaf(a.y, b.y, acc);
                acc = fmaf(a.z, b.z, acc);
                acc = fmaf(a.w, b.w, acc);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            if (row >= 0 && score_mode == FRB_SCORE_REF_COSINE) acc = rf_ref_cosine(acc, qn, __ldg(g_norms + row));
            s_exact[j] = acc;
        }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    float s[FRB_MAX_K];
    int64_t id[FRB_MAX_K];
    list_init<true>(s, id, k);
    float a_min = INFINITY;
    int64_t valid = 0;
    for (int j = 0; j < kp; j++) {
        const int64_t row = cand_idx[q * kp + j];
        if (row < 0) continue;
        valid++;
        a_min = fminf(a_min, cand_approx[q * kp + j]);
        const float v = s_exact[j];
        if (!better<true>(v, row, s[k - 1], id[k - 1])) continue;
        int p = k - 1;
        while (p > 0 && better<true>(v, row, s[p - 1], id[p - 1])) {
            s[p] = s[p - 1];
            id[p] = id[p - 1];
            --p;
        }
        s[p] = v;
        id[p] = row;
    }
    for (int j = 0; j < k; j++) {
        out_scores[q * k + j] = s[j];
        out_idx[q * k + j] = id[j] >= 0 ? id[j] + idx_base : -1;
    }
    // complete when the list holds the whole gallery, or no outside row can reach the k-th exact score
    const bool whole = valid >= n_gallery;
    const float kth = id[k - 1] >= 0 ? s[k - 1] : -INFINITY;
    const float x = a_min + eps;
    const bool fail = !whole && !(x + eps_rel * fabsf(x) + 1e-6f < kth);
    if (fail) atomicAdd(fail_count, 1);
    if (fail_flags) fail_flags[q] = fail ? 1 : 0;
}

}  // namespace frb

using namespace frb;

extern "C" int frb_cosine_rescore_topk(const float *queries, int64_t n_query, const float *gallery, int64_t n_gallery, int dim,
                                       const float *q_norms, const float *g_norms, int score_mode, const int64_t *cand_idx,
                                       const float *cand_approx, int kp, int k, float eps, float eps_rel, int64_t idx_base,
                                       float *out_scores, int64_t *out_idx, int *fail_count, int *fail_flags, void *stream)
{
    FRB_CHECK_ARG(n_query >= 0 && n_gallery >= 0 && dim > 0 && dim % 4 == 0, "frb_cosine_rescore_topk: n_query=%lld n_gallery=%lld dim=%d",
                  (long long)n_query, (long long)n_gallery, dim);
    FRB_CHECK_ARG(k >= 1 && kp >= k && kp <= FRB_MAX_K, "frb_cosine_rescore_topk: k=%d kp=%d (k <= kp <= %d)", k, kp, FRB_MAX_K);
    FRB_CHECK_ARG(score_mode == FRB_SCORE_IP || score_mode == FRB_SCORE_REF_COSINE, "frb_cosine_rescore_topk: score_mode=%d", score_mode);
    FRB_CHECK_ARG(eps >= 0.f && eps_rel >= 0.f, "frb_cosine_rescore_topk: eps_abs=%g eps_rel=%g", (double)eps, (double)eps_rel);
    if (n_query == 0) return FRB_OK;
    FRB_CHECK_ARG(queries && cand_idx && cand_approx && out_scores && out_idx && fail_count && (gallery || n_gallery == 0),
                  "frb_cosine_rescore_topk: null pointer");
    FRB_CHECK_ARG(score_mode != FRB_SCORE_REF_COSINE || (q_norms && (g_norms || n_gallery == 0)),
                  "frb_cosine_rescore_topk: FRB_SCORE_REF_COSINE needs q_norms and g_norms");
    FRB_CHECK_ARG(n_query <= 2147483647LL, "frb_cosine_rescore_topk: n_query too large");
    cosine_rescore_kernel<<<(unsigned)n_query, kRfThreads, 0, (cudaStream_t)stream>>>(queries, gallery, dim, q_norms, g_norms, score_mode,
                                                                                     cand_idx, cand_approx, kp, k, eps, eps_rel, n_gallery,
                                                                                     idx_base, out_scores, out_idx, fail_count, fail_flags);
    FRB_LAUNCH_OK("cosine_rescore_kernel");
    return FRB_OK;
}
