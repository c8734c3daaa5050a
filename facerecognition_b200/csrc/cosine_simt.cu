// cosine_simt.cu — K1 (fp32): exact-fp32 similarity + fused top-k on the FFMA pipe.
//
// Replaces RecognitionEngine.recognize_with_db's Python loop + sort (reference
// inference/recognition_engine.py:277-289), the FaceNet matcher loop (web_app.py:545-554) and the
// notebooks' np.dot + argsort (notebooks/evaluate_arcface_kaggle.ipynb:618,713) for fp32 galleries,
// where the 1e-5 tolerance rules out tensor-core input rounding (tcgen05 has no fp32-input MMA).
//
// One CTA = (64-query tile, range of 128-row gallery tiles).  Classic register-tiled SGEMM
// (256 threads, 4x8 accumulators each, BK = 16) whose 64x128 score tile goes to shared memory only;
// thread q (< 64) then scans its query's 128 scores, applies the reference's score rule
// (frb_score) and updates a running best-k list that persists across the CTA's gallery tiles.
// Per-CTA lists go to the workspace and are merged by topk_merge_kernel.
#include "frb_common.cuh"

namespace frb {

constexpr int kBQ = 64, kBN = 128, kBK = 16, kSimtThreads = 256;

template <typename T>
__device__ __forceinline__ float4 load4(const T *p, bool ok);

template <>
__device__ __forceinline__ float4 load4<float>(const float *p, bool ok)
{
    return ok ? __ldg(reinterpret_cast<const float4 *>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16 *p, bool ok)
{
    if (!ok) return make_float4(0.f, 0.f, 0.f, 0.f);
    uint2 w = __ldg(reinterpret_cast<const uint2 *>(p));
    return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xFFFF0000u), __uint_as_float(w.y << 16),
                       __uint_as_float(w.y & 0xFFFF0000u));
}

// cosine_similarity(), inference/recognition_engine.py:52-63
__device__ __forceinline__ float ref_cosine(float dot, float na, float nb)
{
    if (na == 0.f || nb == 0.f) return 0.f;
    if (fabsf(na - 1.0f) < 1e-3f && fabsf(nb - 1.0f) < 1e-3f) return dot;
    return __fdiv_rn(dot, __fmul_rn(na, nb));
}

template <typename GT>
__global__ void __launch_bounds__(kSimtThreads)
cosine_simt_kernel(const float *__restrict__ queries, int64_t n_query, const GT *__restrict__ gallery, int64_t n_gallery,
                   int dim, const float *__restrict__ q_norms, const float *__restrict__ g_norms, int score_mode,
                   int64_t tiles_per_chunk, int k, int64_t idx_base, float *__restrict__ cand_scores,
                   int64_t *__restrict__ cand_idx)
{
    __shared__ __align__(16) float As[kBK][kBQ + 4];
    __shared__ __align__(16) float Bs[kBK][kBN + 4];
    __shared__ float St[kBQ][kBN + 1];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads: tx -> 8 gallery rows, ty -> 4 queries
    const int64_t q0 = (int64_t)blockIdx.x * kBQ;
    const int64_t chunk = blockIdx.y;
    const int64_t n_tiles = (n_gallery + kBN - 1) / kBN;
    const int64_t tile_begin = chunk * tiles_per_chunk;
    int64_t tile_end = tile_begin + tiles_per_chunk;
    if (tile_end > n_tiles) tile_end = n_tiles;

    // loader mapping: 4 threads per row cover 16 consecutive k (one float4 each)
    const int lr = tid >> 2, lk = (tid & 3) * 4;

    float best_s[FRB_MAX_K];
    int64_t best_i[FRB_MAX_K];
    float kth = -INFINITY;
    float my_qn = 0.f;
    if (tid < kBQ) {
        list_init<true>(best_s, best_i, k);
        if (score_mode == FRB_SCORE_REF_COSINE && q0 + tid < n_query) my_qn = q_norms[q0 + tid];
    }

    for (int64_t tile = tile_begin; tile < tile_end; tile++) {
        const int64_t n0 = tile * kBN;
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

        for (int k0 = 0; k0 < dim; k0 += kBK) {
            // queries: 64 rows x 16 k -> As[k][row];  gallery: 128 rows x 16 k -> Bs[k][row]
            {
                const bool ok = (q0 + lr < n_query) && (k0 + lk < dim);
                float4 v = load4<float>(queries + (q0 + lr) * dim + k0 + lk, ok);
                As[lk + 0][lr] = v.x; As[lk + 1][lr] = v.y; As[lk + 2][lr] = v.z; As[lk + 3][lr] = v.w;
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int r = lr + h * 64;
                const bool ok = (n0 + r < n_gallery) && (k0 + lk < dim);
                float4 v = load4<GT>(gallery + (n0 + r) * dim + k0 + lk, ok);
                Bs[lk + 0][r] = v.x; Bs[lk + 1][r] = v.y; Bs[lk + 2][r] = v.z; Bs[lk + 3][r] = v.w;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < kBK; kk++) {
                const float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
                const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 8]);
                const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 8 + 4]);
                const float av[4] = {a.x, a.y, a.z, a.w};
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) St[ty * 4 + i][tx * 8 + j] = acc[i][j];
        __syncthreads();
        if (tid < kBQ && q0 + tid < n_query) {
            const int lim = (int)((n_gallery - n0) < kBN ? (n_gallery - n0) : kBN);
            for (int j = 0; j < lim; j++) {
                float s = St[tid][j];
                if (score_mode == FRB_SCORE_REF_COSINE) s = ref_cosine(s, my_qn, __ldg(g_norms + n0 + j));
                if (s > kth) kth = list_insert_stream<true>(best_s, best_i, k, s, idx_base + n0 + j);
            }
        }
        // St is rewritten only after the next tile's k-loop barriers
    }
    if (tid < kBQ && q0 + tid < n_query) {
        const int64_t o = (chunk * n_query + q0 + tid) * k;
        for (int j = 0; j < k; j++) {
            cand_scores[o + j] = best_s[j];
            cand_idx[o + j] = best_i[j];
        }
    }
}

int64_t simt_chunks(int64_t n_query, int64_t n_gallery, int64_t *tiles_per_chunk)
{
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    const int64_t q_tiles = (n_query + kBQ - 1) / kBQ;
    int64_t n_tiles = (n_gallery + kBN - 1) / kBN;
    if (n_tiles < 1) n_tiles = 1;
    // ~4 CTAs per SM in total: small galleries (configs[1]: 79 tiles x 4 query tiles) then run one tile per CTA, 2-3 CTAs per SM
    int64_t want = ((int64_t)sms * 4 + q_tiles - 1) / (q_tiles > 0 ? q_tiles : 1);
    if (want < 1) want = 1;
    if (want > n_tiles) want = n_tiles;
    if (want > 65535) want = 65535;
    int64_t tpc = (n_tiles + want - 1) / want;
    *tiles_per_chunk = tpc;
    return (n_tiles + tpc - 1) / tpc;
}

int launch_cosine_simt(const float *queries, int64_t nq, const void *gallery, int gallery_dtype, int64_t ng, int dim,
                       const float *q_norms, const float *g_norms, int score_mode, int k, int64_t idx_base,
                       float *cand_scores, int64_t *cand_idx, int64_t tiles_per_chunk, int64_t chunks, cudaStream_t st)
{
    dim3 grid((unsigned)((nq + kBQ - 1) / kBQ), (unsigned)chunks);
    ProfileScope prof(FRB_K_COSINE_SIMT, st);
    if (gallery_dtype == FRB_F32)
        cosine_simt_kernel<float><<<grid, kSimtThreads, 0, st>>>(queries, nq, (const float *)gallery, ng, dim, q_norms, g_norms,
                                                                 score_mode, tiles_per_chunk, k, idx_base, cand_scores, cand_idx);
    else
        cosine_simt_kernel<__nv_bfloat16><<<grid, kSimtThreads, 0, st>>>(queries, nq, (const __nv_bfloat16 *)gallery, ng, dim,
                                                                         q_norms, g_norms, score_mode, tiles_per_chunk, k,
                                                                         idx_base, cand_scores, cand_idx);
    FRB_LAUNCH_OK("cosine_simt_kernel");
    return FRB_OK;
}

}  // namespace frb
