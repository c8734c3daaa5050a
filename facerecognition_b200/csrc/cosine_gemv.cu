// cosine_gemv.cu — K1 for 1..4 queries: the single-image path of RecognitionEngine.recognize()
// (reference inference/recognition_engine.py:267-289, 291-326; web_app.py:545-554), where the work is one pass
// over the gallery and the bound is HBM, not the tensor pipe (intensity ~Q flop/B, ridge ~220).
//
// The tiled kernels waste a 64- or 128-query tile on one query; this one keeps the queries in registers and
// streams gallery rows: a warp takes one row at a time, lane l owns the 128-bit groups {l + 32 j} of the row
// (512 contiguous bytes per load instruction), accumulates Q partial dot products in fp32, and a 5-step xor
// (segmented) butterfly leaves query q's sum in lane group q, whose first lane keeps the running best-k list; rows arrive in
// increasing order and admission is strict, so equal scores keep the lowest row.  Warp lists are merged per
// CTA in shared memory, CTA lists by the warp-per-query compact merge (topk_merge_compact).
//
// Operand rounding matches the batched kernels, so an answer does not depend on how many queries were sent
// together: fp32 galleries use fp32 queries (after the reference's element-wise normalisation, if asked);
// bf16 galleries use the bf16-rounded normalised queries the tcgen05 path feeds its MMAs, fp32 accumulation.
#include "frb_common.cuh"

namespace frb {

constexpr int kGvThreads = 256, kGvWarps = kGvThreads / 32;
constexpr int kGvMaxQ = 4;
constexpr int kGvMaxPerLane = 16;  // dim <= 512
constexpr int kGvUnroll = 4;       // rows in flight per warp (fp32: 8 KB); bf16 rows are half as long and take 8

__device__ __forceinline__ float gv_ref_cosine(float dot, float na, float nb)  // cosine_similarity(), recognition_engine.py:52-63
{
    if (na == 0.f || nb == 0.f) return 0.f;
    if (fabsf(na - 1.0f) < 1e-3f && fabsf(nb - 1.0f) < 1e-3f) return dot;
    return __fdiv_rn(dot, __fmul_rn(na, nb));
}

__device__ __forceinline__ uint4 gv_ld(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// values of one 128-bit group as fp32: 4 floats (fp32 gallery) or 8 bf16 (bf16 gallery)
template <typename GT>
struct GvGroup;
template <>
struct GvGroup<float> {
    static constexpr int kVals = 4;
    static __device__ __forceinline__ void unpack(const uint4 &w, float *v)
    {
        v[0] = __uint_as_float(w.x); v[1] = __uint_as_float(w.y); v[2] = __uint_as_float(w.z); v[3] = __uint_as_float(w.w);
    }
};
template <>
struct GvGroup<__nv_bfloat16> {
    static constexpr int kVals = 8;
    static __device__ __forceinline__ void unpack(const uint4 &w, float *v)
    {
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            v[2 * e] = __uint_as_float(ww[e] << 16);
            v[2 * e + 1] = __uint_as_float(ww[e] & 0xFFFF0000u);
        }
    }
};

// queries: fp32 [NQ, dim] (fp32 gallery) or bf16 [NQ, dim] (bf16 gallery), already normalised as requested
template <typename GT, int NQ>
__global__ void __launch_bounds__(kGvThreads)
cosine_gemv_kernel(const GT *__restrict__ queries, const GT *__restrict__ gallery, int64_t n_gallery, int dim,
                   const float *__restrict__ q_norms, const float *__restrict__ g_norms, int score_mode, int64_t rows_per_cta,
                   int k, int64_t idx_base, float *__restrict__ cand_scores, int64_t *__restrict__ cand_idx, int *__restrict__ cand_cnt)
{
    constexpr int V = GvGroup<GT>::kVals;
    constexpr int G = kGvMaxPerLane / V;  // 128-bit groups per lane per row at dim = 512
    constexpr int U = kGvUnroll * V / 4;  // rows in flight per warp: the same 8 KB for either dtype
    extern __shared__ __align__(16) unsigned char gv_smem[];  // [warps][NQ][k] scores, then ids

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int groups = dim / (32 * V);  // <= G
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
    int64_t r_end = r_begin + rows_per_cta;
    if (r_end > n_gallery) r_end = n_gallery;

    float qv[NQ][G][V];
#pragma unroll
    for (int q = 0; q < NQ; q++)
#pragma unroll
        for (int j = 0; j < G; j++) {
            uint4 w = make_uint4(0, 0, 0, 0);
            if (j < groups) w = __ldg(reinterpret_cast<const uint4 *>(queries + (int64_t)q * dim) + lane + 32 * j);
            GvGroup<GT>::unpack(w, qv[q][j]);
        }
    // after the reduction query q's total sits in lane group q: lanes 0 (1 query), 0/16 (2), 0/8/16/24 (3-4) keep the lists
    constexpr int kOwnerStride = NQ == 1 ? 32 : (NQ == 2 ? 16 : 8);
    const int my_q = lane / kOwnerStride;
    const int owner = my_q < NQ ? my_q * kOwnerStride : -1;   // == lane for list-keeping lanes
    float my_qn = 0.f;
    if (score_mode == FRB_SCORE_REF_COSINE && lane == owner) my_qn = q_norms[my_q];

    float best_s[FRB_MAX_K];
    int64_t best_i[FRB_MAX_K];
    float kth = -INFINITY;
    if (lane == owner) list_init<true>(best_s, best_i, k);

    const uint4 *gal4 = reinterpret_cast<const uint4 *>(gallery);
    const int64_t vec_per_row = dim / V;
    for (int64_t r0 = r_begin + warp * U; r0 < r_end; r0 += kGvWarps * U) {
        uint4 w[U][G];
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int j = 0; j < G; j++) {
                w[u][j] = make_uint4(0, 0, 0, 0);
                if (j < groups && r0 + u < r_end) w[u][j] = gv_ld(gal4 + (r0 + u) * vec_per_row + lane + 32 * j);
            }
#pragma unroll
        for (int u = 0; u < U; u++) {
            float acc[NQ];
#pragma unroll
            for (int q = 0; q < NQ; q++) acc[q] = 0.f;
#pragma unroll
            for (int j = 0; j < G; j++) {
                float g[V];
                GvGroup<GT>::unpack(w[u][j], g);
#pragma unroll
                for (int e = 0; e < V; e++)
#pragma unroll
                    for (int q = 0; q < NQ; q++) acc[q] = fmaf(g[e], qv[q][j][e], acc[q]);
            }
            // reduce across the warp; with several queries a segmented butterfly (halves swap values instead of
            // every lane summing everything) leaves query q's total in lanes [owner_stride * q, owner_stride * (q + 1))
            float mine;
            if (NQ == 1) {
                mine = acc[0];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
            } else if (NQ == 2) {
                const bool hi = lane & 16;
                mine = (hi ? acc[1] : acc[0]) + __shfl_xor_sync(0xffffffffu, hi ? acc[0] : acc[1], 16);
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
            } else {
                const float a3 = NQ > 3 ? acc[NQ > 3 ? 3 : 0] : 0.f;
                const bool hi = lane & 16;
                float k0 = hi ? acc[2] : acc[0], k1 = hi ? a3 : acc[1];
                k0 += __shfl_xor_sync(0xffffffffu, hi ? acc[0] : acc[2], 16);
                k1 += __shfl_xor_sync(0xffffffffu, hi ? acc[1] : a3, 16);
                const bool hi2 = lane & 8;
                mine = (hi2 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, hi2 ? k0 : k1, 8);
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
            }
            const int64_t row = r0 + u;
            if (row < r_end && lane == owner) {
                if (score_mode == FRB_SCORE_REF_COSINE) mine = gv_ref_cosine(mine, my_qn, __ldg(g_norms + row));
                if (mine > kth) kth = list_insert_stream<true>(best_s, best_i, k, mine, idx_base + row);
            }
        }
    }

    // warp lists -> shared -> one list per (CTA, query)
    float *sm_s = reinterpret_cast<float *>(gv_smem);
    int64_t *sm_i = reinterpret_cast<int64_t *>(gv_smem + (((size_t)kGvWarps * NQ * k * sizeof(float) + 15) & ~(size_t)15));
    if (lane == owner)
        for (int j = 0; j < k; j++) {
            sm_s[(warp * NQ + my_q) * k + j] = best_s[j];
            sm_i[(warp * NQ + my_q) * k + j] = best_i[j];
        }
    __syncthreads();
    if (tid < NQ) {
        float s[FRB_MAX_K];
        int64_t id[FRB_MAX_K];
        list_init<true>(s, id, k);
        for (int w2 = 0; w2 < kGvWarps; w2++)
            for (int j = 0; j < k; j++) {
                const float v = sm_s[(w2 * NQ + tid) * k + j];
                const int64_t idx = sm_i[(w2 * NQ + tid) * k + j];
                if (idx < 0 || !better<true>(v, idx, s[k - 1], id[k - 1])) continue;
                int p = k - 1;
                while (p > 0 && better<true>(v, idx, s[p - 1], id[p - 1])) {
                    s[p] = s[p - 1];
                    id[p] = id[p - 1];
                    --p;
                }
                s[p] = v;
                id[p] = idx;
            }
        // compact layout of topk_merge_compact: query q owns cand[q * cap ...], cap = CTAs * k, this CTA's slice inside it
        const int64_t cap = (int64_t)gridDim.x * k;
        const int64_t o = (int64_t)tid * cap + (int64_t)blockIdx.x * k;
        for (int j = 0; j < k; j++) {
            cand_scores[o + j] = s[j];
            cand_idx[o + j] = id[j];
        }
        if (blockIdx.x == 0) cand_cnt[tid] = (int)cap;
    }
}

// Measured on B200 against a 1M x 512 gallery (profiles/run_latency.py): bf16, 1 query 0.28 ms here vs 0.33 ms on the
// tcgen05 path (which wins from 3 queries up); fp32, 1 / 4 queries 0.44 / 0.56 ms here vs 3.9 ms for 8 queries on the
// FFMA-tiled path, so fp32 batches up to 16 queries run here as passes of 4.
bool gemv_applicable(int64_t n_query, int dim, int gallery_dtype)
{
    if (n_query < 1 || dim % 256 != 0 || dim > 32 * kGvMaxPerLane || gallery_dtype == FRB_F16) return false;
    return n_query <= (gallery_dtype == FRB_BF16 ? 2 : 16);
}

int64_t gemv_ctas(int64_t n_gallery, int64_t *rows_per_cta)
{
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    const int64_t min_rows = kGvWarps * kGvUnroll * 2;  // one round of every warp (bf16 takes 8 rows per round)
    int64_t want = (int64_t)sms * 4;
    int64_t max_ctas = (n_gallery + min_rows - 1) / min_rows;
    if (max_ctas < 1) max_ctas = 1;
    if (want > max_ctas) want = max_ctas;
    int64_t rpc = (n_gallery + want - 1) / want;
    rpc = (rpc + min_rows - 1) / min_rows * min_rows;
    if (rpc < min_rows) rpc = min_rows;
    *rows_per_cta = rpc;
    int64_t ctas = (n_gallery + rpc - 1) / rpc;
    return ctas < 1 ? 1 : ctas;
}

template <typename GT>
static int launch_gemv_t(const GT *queries, int64_t nq, const GT *gallery, int64_t ng, int dim, const float *q_norms, const float *g_norms,
                         int score_mode, int k, int64_t idx_base, float *cs, int64_t *ci, int *cnt, int64_t rpc, int64_t ctas, cudaStream_t st)
{
    const int64_t cap = ctas * k;  // compact candidate slots per query
    for (int64_t q0 = 0; q0 < nq; q0 += kGvMaxQ) {   // passes of up to 4 queries, each streaming the gallery once
        const int n = (int)(nq - q0 < kGvMaxQ ? nq - q0 : kGvMaxQ);
        const size_t smem = align_up((size_t)kGvWarps * n * k * sizeof(float), 16) + (size_t)kGvWarps * n * k * sizeof(int64_t);
        const GT *qp = queries + q0 * dim;
        const float *qn = q_norms ? q_norms + q0 : nullptr;
        float *csp = cs + q0 * cap;
        int64_t *cip = ci + q0 * cap;
        int *cntp = cnt + q0;
        ProfileScope prof(FRB_K_COSINE_GEMV, st);
#define FRB_GV(NQ)                                                                                                                  \
    cosine_gemv_kernel<GT, NQ><<<(unsigned)ctas, kGvThreads, smem, st>>>(qp, gallery, ng, dim, qn, g_norms, score_mode, rpc, k, idx_base, \
                                                                         csp, cip, cntp)
        switch (n) {
            case 1: FRB_GV(1); break;
            case 2: FRB_GV(2); break;
            case 3: FRB_GV(3); break;
            default: FRB_GV(4); break;
        }
#undef FRB_GV
        FRB_LAUNCH_OK("cosine_gemv_kernel");
    }
    return FRB_OK;
}

int launch_cosine_gemv(const void *queries, int64_t nq, const void *gallery, int gallery_dtype, int64_t ng, int dim, const float *q_norms,
                       const float *g_norms, int score_mode, int k, int64_t idx_base, float *cs, int64_t *ci, int *cnt, int64_t rpc,
                       int64_t ctas, cudaStream_t st)
{
    if (gallery_dtype == FRB_F32)
        return launch_gemv_t<float>((const float *)queries, nq, (const float *)gallery, ng, dim, q_norms, g_norms, score_mode, k, idx_base,
                                    cs, ci, cnt, rpc, ctas, st);
    return launch_gemv_t<__nv_bfloat16>((const __nv_bfloat16 *)queries, nq, (const __nv_bfloat16 *)gallery, ng, dim, q_norms, g_norms,
                                        score_mode, k, idx_base, cs, ci, cnt, rpc, ctas, st);
}

}  // namespace frb
