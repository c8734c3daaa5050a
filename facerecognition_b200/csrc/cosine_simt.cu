// cosine_simt.cu — K1 (fp32): exact-fp32 similarity + fused top-k on the FFMA pipe.
//
// Replaces RecognitionEngine.recognize_with_db's Python loop + sort (reference
// inference/recognition_engine.py:277-289), the FaceNet matcher loop (web_app.py:545-554) and the
// notebooks' np.dot + argsort (notebooks/evaluate_arcface_kaggle.ipynb:618,713) for fp32 galleries,
// where the 1e-5 tolerance rules out tensor-core input rounding (tcgen05 has no fp32-input MMA).
//
// One CTA = (64-query tile, range of 128-row gallery tiles).  Register-tiled SGEMM: 256 threads as 16 x 16, each
// thread TM x 8 accumulators (TM = 4; a 128-query TM = 8 tile measured no faster), packed FFMA2 with a scalar
// broadcast operand, BK = 16, shared tiles double-buffered behind a two-slab-deep register prefetch (one barrier
// per k-step; a whole slab of FMAs hides the global latency).  The score tile goes to shared memory only; four threads
// per query apply the reference's score rule (frb_score) in place and pre-filter the tile by its maximum, and the
// query's owner thread walks the 128 scores only when that maximum beats the list: a running best-k
// list that persists across the CTA's gallery tiles (four threads per query pre-filter each tile by its maximum).
// Per-CTA lists go to the workspace and are merged by topk_merge_kernel.
#include "frb_common.cuh"

namespace frb {

constexpr int kBN = 128, kBK = 16, kSimtThreads = 256;

template <typename T>
__device__ __forceinline__ float4 load4(const T *p, bool ok);

// The loads are volatile asm on purpose: the optimiser otherwise sinks the next slab's prefetch below the FMA block
// (seen in SASS: LDG right before the STS that consumes it), which serialises global latency with the math.
template <>
__device__ __forceinline__ float4 load4<float>(const float *p, bool ok)
{
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16 *p, bool ok)
{
    uint2 w = make_uint2(0u, 0u);
    if (ok) asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(w.x), "=r"(w.y) : "l"(p));
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

// shared memory (dynamic): As[2][kBK][BQ + 4], Bs[2][kBK][kBN + 4], St[BQ][kBN + 4]
template <int TM>
struct SimtSmem {
    static constexpr int BQ = 16 * TM;
    static constexpr int kAs = kBK * (BQ + 4), kBs = kBK * (kBN + 4), kSt = BQ * (kBN + 4);
    static constexpr size_t kBytes = (size_t)(2 * kAs + 2 * kBs + kSt) * sizeof(float);
};

template <typename GT, int TM>
__global__ void __launch_bounds__(kSimtThreads, TM == 8 ? 1 : 2)
cosine_simt_kernel(const float *__restrict__ queries, int64_t n_query, const GT *__restrict__ gallery, int64_t n_gallery,
                   int dim, const float *__restrict__ q_norms, const float *__restrict__ g_norms, int score_mode,
                   int64_t tiles_per_chunk, int k, int64_t idx_base, float *__restrict__ cand_scores,
                   int64_t *__restrict__ cand_idx)
{
    constexpr int BQ = 16 * TM;
    constexpr int A_LD = BQ + 4, B_LD = kBN + 4, S_LD = kBN + 4;  // S_LD = 4 mod 32: the 4-threads-per-query scan is conflict-free
    constexpr int A_PER_THREAD = BQ * kBK / 4 / kSimtThreads;  // float4 loads per thread per k-slab: 1 (TM=4) or 2
    extern __shared__ __align__(16) float simt_smem[];
    float *As = simt_smem;                                  // [2][kBK][A_LD]
    float *Bs = As + 2 * SimtSmem<TM>::kAs;                 // [2][kBK][B_LD]
    float *St = Bs + 2 * SimtSmem<TM>::kBs;                 // [BQ][S_LD]

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads: tx -> gallery rows {tx*4.., 64 + tx*4..}, ty -> queries
    const int64_t q0 = (int64_t)blockIdx.x * BQ;
    const int64_t chunk = blockIdx.y;
    const int64_t n_tiles = (n_gallery + kBN - 1) / kBN;
    const int64_t tile_begin = chunk * tiles_per_chunk;
    int64_t tile_end = tile_begin + tiles_per_chunk;
    if (tile_end > n_tiles) tile_end = n_tiles;

    // loader mapping: 4 threads per row cover 16 consecutive k (one float4 each); 64 rows per pass
    const int lr = tid >> 2, lk = (tid & 3) * 4;

    // score scan: 4 threads per query (sq = query, seg = which interleaved quarter of the tile's 128 scores);
    // thread seg == 0 owns the query's running best-k list
    static_assert(BQ * 4 == kSimtThreads, "the scan maps 4 threads to each query of the tile");
    const int sq = tid >> 2, seg = tid & 3;
    const bool q_live = q0 + sq < n_query;
    float best_s[FRB_MAX_K];
    int64_t best_i[FRB_MAX_K];
    float kth = -INFINITY;
    float my_qn = 0.f;
    if (seg == 0) list_init<true>(best_s, best_i, k);
    if (score_mode == FRB_SCORE_REF_COSINE && q_live) my_qn = q_norms[q0 + sq];

    for (int64_t tile = tile_begin; tile < tile_end; tile++) {
        const int64_t n0 = tile * kBN;
        float2 acc[TM][4];  // acc[i][j].{x,y} = scores of query i against rows 2j, 2j+1 of this thread's 8
#pragma unroll
        for (int i = 0; i < TM; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = make_float2(0.f, 0.f);

        float4 pa[A_PER_THREAD], pb[2];
        auto fetch = [&](int k0) {
#pragma unroll
            for (int h = 0; h < A_PER_THREAD; h++) {
                const int r = lr + h * 64;
                pa[h] = load4<float>(queries + (q0 + r) * dim + k0 + lk, (q0 + r < n_query) && (k0 + lk < dim));
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int r = lr + h * 64;
                pb[h] = load4<GT>(gallery + (n0 + r) * dim + k0 + lk, (n0 + r < n_gallery) && (k0 + lk < dim));
            }
        };
        auto stash = [&](int buf) {
            float *a = As + buf * SimtSmem<TM>::kAs, *b = Bs + buf * SimtSmem<TM>::kBs;
#pragma unroll
            for (int h = 0; h < A_PER_THREAD; h++) {
                const int r = lr + h * 64;
                a[(lk + 0) * A_LD + r] = pa[h].x; a[(lk + 1) * A_LD + r] = pa[h].y;
                a[(lk + 2) * A_LD + r] = pa[h].z; a[(lk + 3) * A_LD + r] = pa[h].w;
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int r = lr + h * 64;
                b[(lk + 0) * B_LD + r] = pb[h].x; b[(lk + 1) * B_LD + r] = pb[h].y;
                b[(lk + 2) * B_LD + r] = pb[h].z; b[(lk + 3) * B_LD + r] = pb[h].w;
            }
        };

        // Software pipeline, two slabs deep: slab s is fetched into registers at the END of iteration s - 2 and
        // stashed into shared memory at the end of iteration s - 1, so a whole iteration of FMAs hides the global
        // latency wherever ptxas places the loads (it sinks a same-iteration prefetch down to its stores).
        fetch(0);
        stash(0);
        if (kBK < dim) fetch(kBK);
        __syncthreads();
        int buf = 0;
        for (int k0 = 0; k0 < dim; k0 += kBK, buf ^= 1) {
            const float *a = As + buf * SimtSmem<TM>::kAs, *b = Bs + buf * SimtSmem<TM>::kBs;
#pragma unroll
            for (int kk = 0; kk < kBK; kk++) {
                float av[TM], bv[8];
#pragma unroll
                for (int h = 0; h < TM / 4; h++) {
                    const float4 v = *reinterpret_cast<const float4 *>(&a[kk * A_LD + h * 64 + ty * 4]);
                    av[4 * h] = v.x; av[4 * h + 1] = v.y; av[4 * h + 2] = v.z; av[4 * h + 3] = v.w;
                }
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const float4 v = *reinterpret_cast<const float4 *>(&b[kk * B_LD + h * 64 + tx * 4]);
                    bv[4 * h] = v.x; bv[4 * h + 1] = v.y; bv[4 * h + 2] = v.z; bv[4 * h + 3] = v.w;
                }
                // packed FFMA2 (scalar-broadcast form): one instruction = two IEEE fmas, half the issue slots
#pragma unroll
                for (int i = 0; i < TM; i++) {
                    const float2 a2 = make_float2(av[i], av[i]);
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = __ffma2_rn(a2, make_float2(bv[2 * j], bv[2 * j + 1]), acc[i][j]);
                }
            }
            if (k0 + kBK < dim) stash(buf ^ 1);             // slab k0 + BK: fetched one iteration ago; that buffer was last read one barrier ago
            if (k0 + 2 * kBK < dim) fetch(k0 + 2 * kBK);    // slab k0 + 2 BK: in flight during the whole next iteration
            __syncthreads();
        }
        // scores -> St (query-major); thread (ty, tx) holds queries {h*64 + ty*4 + i} x rows {h*64 + tx*4 + j}
#pragma unroll
        for (int i = 0; i < TM; i++)
#pragma unroll
            for (int j = 0; j < 8; j++)
                St[((i >> 2) * 64 + ty * 4 + (i & 3)) * S_LD + (j >> 2) * 64 + tx * 4 + (j & 3)] = (j & 1) ? acc[i][j >> 1].y : acc[i][j >> 1].x;
        __syncthreads();
        {
            // pass 1, all threads: apply the score rule in place and take the maximum of my 32 scores; the 4 threads
            // of a query combine with two shuffles.  Only when that maximum beats the query's admission threshold
            // (rare once the list has warmed up) does the owner walk the 128 scores in row order.
            const int lim = (int)((n_gallery - n0) < kBN ? (n_gallery - n0) : kBN);
            float m = -INFINITY;
            if (q_live) {
#pragma unroll 8
                for (int jj = 0; jj < kBN / 4; jj++) {
                    const int j = jj * 4 + seg;
                    if (j < lim) {
                        float sc = St[sq * S_LD + j];
                        if (score_mode == FRB_SCORE_REF_COSINE) {
                            sc = ref_cosine(sc, my_qn, __ldg(g_norms + n0 + j));
                            St[sq * S_LD + j] = sc;
                        }
                        m = fmaxf(m, sc);
                    }
                }
            }
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            __syncwarp();  // the in-place rewrites of the other three threads are visible to the owner
            if (seg == 0 && q_live && m > kth) {
                for (int j = 0; j < lim; j++) {
                    const float sc = St[sq * S_LD + j];
                    if (sc > kth) kth = list_insert_stream<true>(best_s, best_i, k, sc, idx_base + n0 + j);
                }
            }
        }
        __syncthreads();  // St and the tile buffers are rewritten by the next tile
    }
    if (seg == 0 && q_live) {
        const int64_t o = (chunk * n_query + q0 + sq) * k;
        for (int j = 0; j < k; j++) {
            cand_scores[o + j] = best_s[j];
            cand_idx[o + j] = best_i[j];
        }
    }
}

// query-tile height: 64 (TM = 4, two CTAs per SM, 33 TFLOP/s) beat 128 (TM = 8: 31 TFLOP/s with one CTA per SM, 26 with spills at two)
static int simt_tile_queries(int64_t n_query) { (void)n_query; return 64; }  // measured: the 128-query tile (TM = 8, one CTA per SM) is no faster

int64_t simt_chunks(int64_t n_query, int64_t n_gallery, int64_t *tiles_per_chunk)
{
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    const int bq = simt_tile_queries(n_query);
    const int64_t q_tiles = (n_query + bq - 1) / bq;
    int64_t n_tiles = (n_gallery + kBN - 1) / kBN;
    if (n_tiles < 1) n_tiles = 1;
    // ~4 CTAs per SM in total (two are resident): small galleries (configs[1]: 79 tiles x 4 query tiles) run one
    // tile per CTA
    int64_t want = ((int64_t)sms * 4 + q_tiles - 1) / (q_tiles > 0 ? q_tiles : 1);
    if (want < 1) want = 1;
    if (want > n_tiles) want = n_tiles;
    if (want > 65535) want = 65535;
    int64_t tpc = (n_tiles + want - 1) / want;
    *tiles_per_chunk = tpc;
    return (n_tiles + tpc - 1) / tpc;
}

template <typename GT, int TM>
static int launch_simt_t(const float *queries, int64_t nq, const GT *gallery, int64_t ng, int dim, const float *q_norms,
                         const float *g_norms, int score_mode, int k, int64_t idx_base, float *cand_scores, int64_t *cand_idx,
                         int64_t tiles_per_chunk, int64_t chunks, cudaStream_t st)
{
    constexpr int BQ = 16 * TM;
    const size_t smem = SimtSmem<TM>::kBytes;
    static thread_local int attr_dev = -1;
    int dev = 0;
    FRB_CUDA_OK(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        FRB_CUDA_OK(cudaFuncSetAttribute(cosine_simt_kernel<GT, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_dev = dev;
    }
    dim3 grid((unsigned)((nq + BQ - 1) / BQ), (unsigned)chunks);
    ProfileScope prof(FRB_K_COSINE_SIMT, st);
    cosine_simt_kernel<GT, TM><<<grid, kSimtThreads, smem, st>>>(queries, nq, gallery, ng, dim, q_norms, g_norms, score_mode,
                                                                tiles_per_chunk, k, idx_base, cand_scores, cand_idx);
    FRB_LAUNCH_OK("cosine_simt_kernel");
    return FRB_OK;
}

int launch_cosine_simt(const float *queries, int64_t nq, const void *gallery, int gallery_dtype, int64_t ng, int dim,
                       const float *q_norms, const float *g_norms, int score_mode, int k, int64_t idx_base,
                       float *cand_scores, int64_t *cand_idx, int64_t tiles_per_chunk, int64_t chunks, cudaStream_t st)
{
#define FRB_SIMT(GT, TM)                                                                                                     \
    launch_simt_t<GT, TM>(queries, nq, (const GT *)gallery, ng, dim, q_norms, g_norms, score_mode, k, idx_base, cand_scores,  \
                          cand_idx, tiles_per_chunk, chunks, st)
    if (gallery_dtype == FRB_F32) return FRB_SIMT(float, 4);
    return FRB_SIMT(__nv_bfloat16, 4);
#undef FRB_SIMT
}

}  // namespace frb
