// chisq_filter.cu — tensor-core candidate filter in front of K3 (chi-square nearest neighbour over LBPH histograms).
//
// Replaces, for batches of queries, the compareHist scan inside cv2.face LBPH predict() (reference call sites
// web_app.py:587, models/lbphmodel/evaluate_lbph.py:31-33, threshold_lbph.py:47-50 — N predicts in a Python loop):
// the exact scan (chisq.cu) costs ~65 k fp32 operations per (query, gallery row) pair and is FP32-pipe bound as soon
// as a gallery chunk is shared by several queries.  Here almost every pair is decided by a GEMM instead.
//
// Algebra.  With integer counts g, q of equal cell size,  sum_j (g_j - q_j)^2 / (g_j + q_j) = sum g + sum q - 4 S(g, q),
// S = sum_j f(g_j, q_j),  f(a, b) = a b / (a + b), so the nearest row is the row with the LARGEST  S - (sum g) / 4
// (for LBPH histograms sum g is the same constant for every row; the kernel does not rely on that: the generator
// threads see every count of their row anyway and hand the row totals to the epilogue).  f on [0, cell_px]^2 is a symmetric table whose weighted
// eigen-decomposition truncated to rank 8 gives per-count feature vectors u(a), v(b) in fp16 with
// f(a, b) = <u(a), v(b)> + E(a, b), |E| known exactly (host, float64).  S~(g, q) = sum_j <u(g_j), v(q_j)> is an inner
// product of length 8 * hist_len: a GEMM with K = 131072 for the 8x8x256 histogram.
//
// Filter rule (exact result).  |S~ - S| <= e(q) = sum_j max_a |E(a, q_j)| (+ the accumulation allowance) for
// EVERY gallery row, so a row can only be the exact scan's answer if S~ >= max_rows S~ - w(q), w = 2 e (+ the exact
// scan's own rounding).  Those survivors (a handful per query) are re-scored by the exact kernel's own arithmetic
// (chisq_kernel<GATHER>), so distances and ties are bit for bit those of the full scan.  A query whose survivor list
// overflows is flagged and answered by the plain exact scan in the same call (counted in `stats`).
//
// Kernel (chisq_filter_kernel).  D[128 queries x 256 rows] += A[128 x 16] * B[256 x 16]^T, tcgen05.mma kind::f16 with
// fp16 inputs and fp32 accumulators in TMEM.  A = query features, precomputed once per call (chisq_feat_kernel) and
// streamed by TMA (SWIZZLE_128B, K-major).  B = gallery features, GENERATED in shared memory: eight generator warps hold
// one gallery row each per thread, read 16 u8 counts at a time straight from HBM/L2 (16 KiB per row, the gallery is
// never expanded in memory) and write u(count) — one 16-byte table look-up per bin — into the 128-byte-swizzled
// K-major layout the UMMA descriptor expects.  A generated B stage is used by TWO query tiles (two 256-column
// accumulators = all 512 TMEM columns), which halves the generation cost per flop.  One unit = (pair of query tiles,
// 256-row gallery tile) over the whole K; the epilogue (one thread per query, TMEM lane == query) keeps the running
// per-query maximum in global memory and appends the rows within the window to the query's candidate list.
#include "tc_ptx.cuh"

#include <cuda_fp16.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <vector>

namespace frb {

// chisq.cu
int chisq_gather_g8(const uint16_t *qh, int64_t nq, const uint8_t *gal, int64_t ng, int L, int cell_px, int64_t idx_base,
                    const int *q_flag, const int *row_list, const int *row_cnt, int64_t cap, float *cd, int64_t *ci, int *cnt_out,
                    cudaStream_t st);
int chisq_flagged_topk_g8(const uint16_t *qh, int64_t nq, const uint8_t *gal, int64_t ng, int L, int cell_px, int k,
                          int64_t idx_base, const int *q_flag, int chunks, int64_t cap, float *cd, int64_t *ci, int *cnt_out,
                          cudaStream_t st);

constexpr int kCfRank = 8;              // features per bin: one 16-byte fp16x8 vector
constexpr int kCfBlockM = 128;          // queries per tile (TMEM lanes)
constexpr int kCfBlockN = 256;          // gallery rows per tile (TMEM columns per accumulator)
constexpr int kCfBlockK = 64;           // fp16 per 128-byte swizzle row = 8 bins x 8 features
constexpr int kCfQTiles = 2;            // query tiles that share one generated gallery stage
constexpr int kCfAStages = 3, kCfBStages = 3;
constexpr int kCfThreads = 512;
constexpr int kCfGenWarp0 = 8;          // warps 8..15 generate the gallery operand
constexpr uint32_t kCfATileBytes = kCfBlockM * kCfBlockK * 2;      // 16 KB
constexpr uint32_t kCfAStageBytes = kCfQTiles * kCfATileBytes;     // 32 KB
constexpr uint32_t kCfBStageBytes = kCfBlockN * kCfBlockK * 2;     // 32 KB
constexpr int kCfTableRows = 256;
// The generators' table look-ups hit arbitrary rows, so lanes of a quarter-warp would collide on banks (ncu, first
// version: 89.6 M bank conflicts, 8.6 instead of 4 wavefronts per 128-bit look-up).  The shared-memory copy of the
// table is therefore REPLICATED: row a holds COP copies of u(a) side by side and lane l reads copy l % COP, so the
// eight lanes of a quarter-warp always touch eight different bank groups whatever their rows are.  8 copies fit when
// cell_px + 1 <= 176 rows (22 KB; 100x100 and 112x112 faces: 145 / 170 rows); larger cells get 4 copies of 256 rows.
constexpr int kCfRows8 = 176;
constexpr int64_t kCfMaxQueriesPerPass = 1024;
constexpr int kCfFallbackChunks = 64;

// kind::f16 instruction descriptor: D = f32 (1 << 4), A = B = f16 (format 0), K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kCfIdesc = (1u << 4) | ((uint32_t)(kCfBlockN >> 3) << 17) | ((uint32_t)(kCfBlockM >> 4) << 24);

struct CfBarriers {
    uint64_t a_full[kCfAStages], a_empty[kCfAStages];
    uint64_t b_full[kCfBStages], b_empty[kCfBStages];
    uint64_t tmem_full, tmem_empty;
    uint32_t tmem_base;
};

struct CfParams {
    int64_t n_query, n_gallery;
    int hist_len, n_kblocks;            // n_kblocks = hist_len / 8
    int64_t n_qtiles, n_qpairs, n_gtiles;
    const uint8_t *gallery;             // [n_gallery, hist_len] u8 counts
    const uint4 *u_table;               // [256] gallery-side features, 8 x fp16 per count
    int table_rows;                     // rows of the shared-memory copy (counts are clamped to table_rows - 1)
    const float *window;                // [n_query] w(q) in S units
    float *best;                        // [n_query] running max of the score S~ - (row total) / 4, -inf on entry
    int *cnt;                           // [n_query] candidates appended (may exceed cap)
    int cap;
    int *cand_row;                      // [n_query, cap]
    float *cand_s;                      // [n_query, cap]
    float *all_scores;                  // debug: [n_query, n_gallery] S~ of every pair (the raw accumulators), or null
};

// explicit shared-state-space accesses for the generators' hot loop (the compiler emitted generic LD.E / ST.E for the
// pointer arithmetic on the dynamic shared-memory base: same bytes, but a longer path than LDS / STS)
__device__ __forceinline__ uint4 cf_lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void cf_sts128(uint32_t addr, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ float cf_atomic_max(float *addr, float v)   // returns the value before the update
{
    if (v >= 0.f) return __int_as_float(atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v)));
    return __uint_as_float(atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v)));
}

template <int COP>
__global__ void __launch_bounds__(kCfThreads, 1)
chisq_filter_kernel(const __grid_constant__ CUtensorMap tmap_f, const CfParams p)
{
    extern __shared__ unsigned char cf_smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(cf_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *smem_a = smem;
    unsigned char *smem_b = smem + (size_t)kCfAStages * kCfAStageBytes;
    uint4 *tbl = reinterpret_cast<uint4 *>(smem_b + (size_t)kCfBStages * kCfBStageBytes);
    float *rowsum = reinterpret_cast<float *>(tbl + p.table_rows * COP);   // [2][256]: (sum of the row's counts) / 4, by unit parity
    CfBarriers *bars = reinterpret_cast<CfBarriers *>(rowsum + 2 * kCfBlockN);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_units = p.n_qpairs * p.n_gtiles;

    for (int i = threadIdx.x; i < p.table_rows * COP; i += kCfThreads) tbl[i] = p.u_table[i / COP];
    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_f) : "memory");
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kCfAStages; s++) {
            mbar_init(&bars->a_full[s], 1);
            mbar_init(&bars->a_empty[s], 1);
        }
        for (int s = 0; s < kCfBStages; s++) {
            mbar_init(&bars->b_full[s], kCfThreads / 32 - kCfGenWarp0);   // one arrive per generator warp
            mbar_init(&bars->b_empty[s], 1);
        }
        mbar_init(&bars->tmem_full, 1);
        mbar_init(&bars->tmem_empty, 4);                                  // one arrive per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer: query features =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int64_t pair = u / p.n_gtiles;
                const int64_t left = p.n_qtiles - pair * kCfQTiles;
                const int nqt = left < kCfQTiles ? (int)left : kCfQTiles;
                for (int kb = 0; kb < p.n_kblocks; kb++) {
                    mbar_wait(&bars->a_empty[stage], phase ^ 1);
                    mbar_expect_tx(&bars->a_full[stage], (uint32_t)nqt * kCfATileBytes);
                    for (int t = 0; t < nqt; t++)
                        tma_load_2d(smem_a + (size_t)stage * kCfAStageBytes + (size_t)t * kCfATileBytes, &tmap_f, &bars->a_full[stage],
                                    kb * kCfBlockK, (int)((pair * kCfQTiles + t) * kCfBlockM));
                    if (++stage == kCfAStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0, acc_phase = 0;
            for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int64_t pair = u / p.n_gtiles;
                const int64_t left = p.n_qtiles - pair * kCfQTiles;
                const int nqt = left < kCfQTiles ? (int)left : kCfQTiles;
                mbar_wait(&bars->tmem_empty, acc_phase ^ 1);   // the epilogue has drained the previous unit
                tcgen05_fence_after();
                for (int kb = 0; kb < p.n_kblocks; kb++) {
                    mbar_wait(&bars->a_full[sa], pa);
                    mbar_wait(&bars->b_full[sb], pb);
                    tcgen05_fence_after();
                    const uint64_t db = make_sw128_desc(smem_u32(smem_b + (size_t)sb * kCfBStageBytes));
                    for (int t = 0; t < nqt; t++) {
                        const uint64_t da = make_sw128_desc(smem_u32(smem_a + (size_t)sa * kCfAStageBytes + (size_t)t * kCfATileBytes));
                        const uint32_t tmem_d = tmem_base + (uint32_t)t * kCfBlockN;
#pragma unroll
                        for (int k4 = 0; k4 < kCfBlockK / 16; k4++)
                            umma_bf16(tmem_d, da + (uint64_t)(k4 * 2), db + (uint64_t)(k4 * 2), kCfIdesc, (kb | k4) != 0 ? 1u : 0u);
                    }
                    tcgen05_commit(&bars->a_empty[sa]);
                    tcgen05_commit(&bars->b_empty[sb]);
                    if (++sa == kCfAStages) { sa = 0; pa ^= 1; }
                    if (++sb == kCfBStages) { sb = 0; pb ^= 1; }
                }
                tcgen05_commit(&bars->tmem_full);
                acc_phase ^= 1;
            }
        }
    } else if (warp >= kCfGenWarp0) {
        // ===================== generators: u8 counts -> fp16 features, swizzled K-major B stage =====================
        const int r = (warp - kCfGenWarp0) * 32 + lane;          // gallery row inside the tile
        const uint32_t xr = (uint32_t)(r & 7) << 4;              // 128-byte swizzle: 16-byte chunk c of row r sits at c ^ (r & 7)
        const int n_iter = p.n_kblocks >> 1;                     // 16 bins (one 128-bit load) = two k-blocks per iteration
        const uint32_t my_tbl = smem_u32(tbl + (lane & (COP - 1)));   // this lane's copy of every table row
        const uint32_t last_row = (uint32_t)p.table_rows - 1;
        const uint32_t b_base = smem_u32(smem_b) + (uint32_t)r * 128u;
        int sb = 0;
        uint32_t pb = 0, upar = 0;
        for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x, upar ^= 1) {
            const int64_t gt = u % p.n_gtiles;
            uint32_t rsum = 0;
            const int64_t row = gt * kCfBlockN + r;
            const bool valid = row < p.n_gallery;
            const uint4 *src = reinterpret_cast<const uint4 *>(p.gallery + (valid ? row : 0) * (int64_t)p.hist_len);
            // counts are fetched two iterations (four k-blocks) ahead of their use
            uint4 nxt = valid ? __ldg(src) : make_uint4(0, 0, 0, 0);
            uint4 nxt2 = (valid && n_iter > 1) ? __ldg(src + 1) : make_uint4(0, 0, 0, 0);
            for (int it = 0; it < n_iter; it++) {
                const uint4 cur = nxt;
                nxt = nxt2;
                if (valid && it + 2 < n_iter) nxt2 = __ldg(src + it + 2);
                const uint32_t wds[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
                for (int wq = 0; wq < 4; wq++) rsum = __dp4a(wds[wq], 0x01010101u, rsum);
                // the row total is complete with the last counts; it is published before this iteration's b_full arrivals,
                // which the MMA thread's final commit (tmem_full) orders before the epilogue's reads
                if (it == n_iter - 1) rowsum[upar * kCfBlockN + r] = 0.25f * (float)rsum;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    mbar_wait(&bars->b_empty[sb], pb ^ 1);
                    const uint32_t dst = b_base + (uint32_t)sb * kCfBStageBytes;
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const uint32_t cnt = min((wds[h * 2 + (c >> 2)] >> (8 * (c & 3))) & 0xFFu, last_row);
                        cf_sts128(dst + (((uint32_t)c << 4) ^ xr), cf_lds128(my_tbl + cnt * (COP * 16u)));
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the UMMA reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->b_full[sb]);
                    if (++sb == kCfBStages) { sb = 0; pb ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: running maximum + candidates within the window =====================
        const int ew = warp & 3;
        uint32_t acc_phase = 0, upar = 0;
        for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x, upar ^= 1) {
            const float *rs = rowsum + upar * kCfBlockN;
            const int64_t pair = u / p.n_gtiles, gt = u % p.n_gtiles;
            const int64_t left = p.n_qtiles - pair * kCfQTiles;
            const int nqt = left < kCfQTiles ? (int)left : kCfQTiles;
            const int64_t n0 = gt * kCfBlockN;
            const int valid = (int)((p.n_gallery - n0) < kCfBlockN ? (p.n_gallery - n0) : kCfBlockN);
            mbar_wait(&bars->tmem_full, acc_phase);
            acc_phase ^= 1;
            tcgen05_fence_after();
            for (int t = 0; t < nqt; t++) {
                const int64_t q = (pair * kCfQTiles + t) * kCfBlockM + ew * 32 + lane;
                const bool live = q < p.n_query;
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)t * kCfBlockN;
                float m = -INFINITY;
#pragma unroll 1
                for (int c0 = 0; c0 < kCfBlockN; c0 += 32) {
                    float v[32];
                    tmem_ld_32x32(taddr + (uint32_t)c0, v);
#pragma unroll
                    for (int j = 0; j < 32; j++) m = (c0 + j < valid) ? fmaxf(m, v[j] - rs[c0 + j]) : m;
                    if (p.all_scores && live) {
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (c0 + j < valid) p.all_scores[q * p.n_gallery + n0 + c0 + j] = v[j];
                    }
                }
                float thr = INFINITY;
                if (live) {
                    const float before = cf_atomic_max(p.best + q, m);
                    thr = fmaxf(before, m) - p.window[q];
                }
#pragma unroll 1
                for (int c0 = 0; c0 < kCfBlockN; c0 += 32) {
                    float v[32];
                    tmem_ld_32x32(taddr + (uint32_t)c0, v);
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const float sc = v[j] - rs[c0 + j];
                        if (c0 + j < valid && sc >= thr) {
                            const int pos = atomicAdd(p.cnt + q, 1);
                            if (pos < p.cap) {
                                p.cand_row[q * p.cap + pos] = (int)(n0 + c0 + j);
                                p.cand_s[q * p.cap + pos] = sc;
                            }
                        }
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty);
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- query features + per-query window ------------------------------------------------------------------------------
// One CTA per query: feat[q][bin][0..7] = v(count), window[q] = 2 (e_tab + e_acc) + e_scan, and the per-query filter
// state is reset (best = -inf, cnt = 0, flag = 0).
__global__ void __launch_bounds__(256)
chisq_feat_kernel(const uint16_t *__restrict__ qhist, int hist_len, int cell_px, int *__restrict__ stats, const uint4 *__restrict__ v_table,
                  const float *__restrict__ emax, const float *__restrict__ absmax, float acc_rel, uint4 *__restrict__ feat,
                  float *__restrict__ window, float *__restrict__ qtot, float *__restrict__ best, int *__restrict__ cnt,
                  int *__restrict__ flag)
{
    __shared__ float s_red[3][8];
    const int64_t q = blockIdx.x;
    const uint16_t *h = qhist + q * hist_len;
    uint4 *out = feat + q * hist_len;
    float e = 0.f, a = 0.f, tot = 0.f;
    int invalid = 0;
    for (int j = threadIdx.x; j < hist_len; j += 256) {
        int c = h[j];
        // a count above cell_px cannot occur in a histogram of cells with cell_px pixels; the tables and their bound do
        // not cover it, so such a query is answered by the exact scan (flag below)
        invalid |= c > cell_px;
        c = c < kCfTableRows ? c : kCfTableRows - 1;
        out[j] = __ldg(v_table + c);
        e += emax[c];
        a += absmax[c];
        tot += (float)c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        e += __shfl_xor_sync(0xffffffffu, e, o);
        a += __shfl_xor_sync(0xffffffffu, a, o);
        tot += __shfl_xor_sync(0xffffffffu, tot, o);
    }
    if ((threadIdx.x & 31) == 0) {
        s_red[0][threadIdx.x >> 5] = e;
        s_red[1][threadIdx.x >> 5] = a;
        s_red[2][threadIdx.x >> 5] = tot;
    }
    invalid = __syncthreads_or(invalid);
    if (threadIdx.x == 0) {
        float es = 0.f, as = 0.f, ts = 0.f;
        for (int w = 0; w < 8; w++) { es += s_red[0][w]; as += s_red[1][w]; ts += s_red[2][w]; }
        // e_tab: table error (exact tables, summed in fp32: + 1e-4 relative); e_acc: accumulation allowance of the fp32
        // tensor-core sum; last term: the exact scan's own rounding (<= 1e-5 relative on a distance <= 2 * total counts)
        window[q] = 2.0f * (es * 1.0001f + acc_rel * as) + 1e-5f * ts;
        qtot[q] = ts;                                   // exact: an integer below 2^24
        best[q] = -INFINITY;
        cnt[q] = 0;
        flag[q] = invalid ? 1 : 0;
        if (invalid && stats) atomicAdd(stats + 0, 1);
    }
}

// ---- survivors: raw candidates within the FINAL window -> compact row lists ------------------------------------------
// One warp per query.  n_raw > cap: the list is incomplete -> flag the query for the exact scan.
__global__ void __launch_bounds__(256)
chisq_survivor_kernel(int64_t n_query, int cap, const float *__restrict__ best, const float *__restrict__ window,
                      const int *__restrict__ raw_cnt, const int *__restrict__ raw_row, const float *__restrict__ raw_s,
                      int *__restrict__ list_row, float *__restrict__ list_s, int *__restrict__ list_cnt, int *__restrict__ flag,
                      int *__restrict__ stats)
{
    const int64_t q = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= n_query) return;
    if (flag[q]) {                       // already flagged (invalid query histogram): the exact scan answers it
        if (lane == 0) list_cnt[q] = 0;
        return;
    }
    const int n_raw = raw_cnt[q];
    if (n_raw > cap) {
        if (lane == 0) {
            flag[q] = 1;
            list_cnt[q] = 0;
            if (stats) {
                atomicAdd(stats + 0, 1);
                atomicAdd(stats + 2, cap);
            }
        }
        return;
    }
    const float thr = best[q] - window[q];
    int kept = 0;
    for (int i0 = 0; i0 < n_raw; i0 += 32) {
        const int i = i0 + lane;
        const float sc = i < n_raw ? raw_s[q * cap + i] : 0.f;
        const bool keep = i < n_raw && sc >= thr;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int64_t o = q * cap + kept + __popc(m & ((1u << lane) - 1u));
            list_row[o] = raw_row[q * cap + i];
            list_s[o] = sc;
        }
        kept += __popc(m);
    }
    if (lane == 0) {
        list_cnt[q] = kept;
        if (stats) {
            atomicAdd(stats + 1, kept);
            atomicAdd(stats + 2, n_raw);
        }
    }
}

// ---- audit: every re-scored row has both its exact distance and its filter score ----------------------------------------
// score = S~ - (sum g)/4 and d = (2 / cell_px) (sum g + sum q - 4 S), so the exact score is (sum q - d cell_px / 2) / 4.
// |score - exact score| must be within e(q) = window / 2 — the bound the survivor rule rests on.  A violation (which the
// table bound excludes and the accumulation allowance is sized never to see) flags the query for the exact scan and is
// counted in stats[3]; the call then still returns the exact answer for it.
__global__ void __launch_bounds__(256)
chisq_audit_kernel(int64_t n_query, int cap, float half_cell_px, const int *__restrict__ list_cnt, const float *__restrict__ list_s,
                   const float *__restrict__ cand_dist, const float *__restrict__ qtot, const float *__restrict__ window,
                   int *__restrict__ flag, int *__restrict__ stats)
{
    const int64_t q = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= n_query || flag[q]) return;
    const int n = list_cnt[q];
    const float e = 0.5f * window[q], tq = qtot[q];
    int bad = 0;
    for (int i = lane; i < n; i += 32) {
        const float exact = 0.25f * (tq - cand_dist[q * cap + i] * half_cell_px);
        bad += fabsf(list_s[q * cap + i] - exact) > e ? 1 : 0;
    }
    bad = __reduce_add_sync(0xffffffffu, bad);
    if (lane == 0 && bad) {
        flag[q] = 1;
        if (stats) {
            atomicAdd(stats + 0, 1);
            atomicAdd(stats + 3, bad);
        }
    }
}

// ---- host: feature tables -----------------------------------------------------------------------------------------
struct CfTables {
    uint4 *d_u = nullptr, *d_v = nullptr;     // [256] fp16 x 8 per count (gallery side u, query side v)
    float *d_emax = nullptr, *d_absmax = nullptr;   // [256] per query count b: max_a |E(a, b)|, max_a sum_m |u_m(a) v_m(b)|
};

// cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (row-major, destroyed); V's columns are the eigenvectors
static void jacobi_eigen(std::vector<double> &A, int n, std::vector<double> &V)
{
    V.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) V[(size_t)i * n + i] = 1.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) (i == j ? diag : off) += A[(size_t)i * n + j] * A[(size_t)i * n + j];
        if (off <= 1e-30 * diag) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = A[(size_t)p * n + q];
                if (fabs(apq) < 1e-300) continue;
                const double theta = (A[(size_t)q * n + q] - A[(size_t)p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) {
                    const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
                    A[(size_t)k * n + p] = c * akp - s * akq;
                    A[(size_t)k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
                    A[(size_t)p * n + k] = c * apk - s * aqk;
                    A[(size_t)q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
                    V[(size_t)k * n + p] = c * vkp - s * vkq;
                    V[(size_t)k * n + q] = s * vkp + c * vkq;
                }
            }
    }
}

static uint16_t f2h_bits(double x)
{
    const __half h = __float2half_rn((float)x);
    return *reinterpret_cast<const uint16_t *>(&h);
}
static double h2d(uint16_t b)
{
    __half h;
    *reinterpret_cast<uint16_t *>(&h) = b;
    return (double)__half2float(h);
}

// u, v [256][8] fp16 bits; emax, absmax [256].  Weighted rank-8 eigen-truncation of F = ab/(a+b) on [0, cell_px]^2:
// factor W F W with w(a) = (1 + a)^-1.5 (small counts carry almost all bins of an LBP histogram, so their entries are
// fitted tightest), then unscale.  Row 0 is exactly zero on both sides (f(0, .) = f(., 0) = 0), so empty bins and
// zero-filled padding contribute exactly 0.
void chisq_filter_build_tables(int cell_px, uint16_t *u, uint16_t *v, float *emax, float *absmax)
{
    const int n = cell_px + 1;
    std::vector<double> F((size_t)n * n), A((size_t)n * n), V, w(n);
    for (int a = 0; a < n; a++) w[a] = pow(1.0 + a, -1.5);
    for (int a = 0; a < n; a++)
        for (int b = 0; b < n; b++) {
            F[(size_t)a * n + b] = (a + b) > 0 ? (double)a * b / (double)(a + b) : 0.0;
            A[(size_t)a * n + b] = w[a] * F[(size_t)a * n + b] * w[b];
        }
    jacobi_eigen(A, n, V);
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    for (int i = 0; i < n; i++)   // selection sort by |eigenvalue|, descending (n <= 256)
        for (int j = i + 1; j < n; j++)
            if (fabs(A[(size_t)order[j] * n + order[j]]) > fabs(A[(size_t)order[i] * n + order[i]])) std::swap(order[i], order[j]);
    for (int i = 0; i < kCfTableRows * kCfRank; i++) u[i] = v[i] = 0;
    for (int m = 0; m < kCfRank && m < n; m++) {
        const int e = order[m];
        const double lam = A[(size_t)e * n + e], sq = sqrt(fabs(lam)), sg = lam < 0 ? -1.0 : 1.0;
        for (int a = 1; a < n; a++) {
            const double x = sq * V[(size_t)a * n + e] / w[a];
            u[a * kCfRank + m] = f2h_bits(x);
            v[a * kCfRank + m] = f2h_bits(sg * x);
        }
    }
    for (int b = 0; b < kCfTableRows; b++) emax[b] = absmax[b] = 0.f;
    for (int b = 0; b < n; b++) {
        double em = 0.0, am = 0.0;
        for (int a = 0; a < n; a++) {
            double dot = 0.0, ab = 0.0;
            for (int m = 0; m < kCfRank; m++) {
                const double t = h2d(u[a * kCfRank + m]) * h2d(v[b * kCfRank + m]);
                dot += t;
                ab += fabs(t);
            }
            em = fmax(em, fabs(dot - F[(size_t)a * n + b]));
            am = fmax(am, ab);
        }
        emax[b] = nextafterf((float)em, INFINITY);
        absmax[b] = nextafterf((float)am, INFINITY);
    }
}

static int get_tables(int cell_px, CfTables *out)
{
    static std::mutex mu;
    static std::map<std::pair<int, int>, CfTables> cache;
    int dev = 0;
    FRB_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find({dev, cell_px});
    if (it != cache.end()) {
        *out = it->second;
        return FRB_OK;
    }
    std::vector<uint16_t> u(kCfTableRows * kCfRank), v(kCfTableRows * kCfRank);
    std::vector<float> emax(kCfTableRows), absmax(kCfTableRows);
    chisq_filter_build_tables(cell_px, u.data(), v.data(), emax.data(), absmax.data());
    CfTables t;
    // first use per (device, cell size): plain allocations + blocking copies (not capturable into a CUDA graph)
    FRB_CUDA_OK(cudaMalloc(&t.d_u, kCfTableRows * sizeof(uint4)));
    FRB_CUDA_OK(cudaMalloc(&t.d_v, kCfTableRows * sizeof(uint4)));
    FRB_CUDA_OK(cudaMalloc(&t.d_emax, kCfTableRows * sizeof(float)));
    FRB_CUDA_OK(cudaMalloc(&t.d_absmax, kCfTableRows * sizeof(float)));
    FRB_CUDA_OK(cudaMemcpy(t.d_u, u.data(), kCfTableRows * sizeof(uint4), cudaMemcpyHostToDevice));
    FRB_CUDA_OK(cudaMemcpy(t.d_v, v.data(), kCfTableRows * sizeof(uint4), cudaMemcpyHostToDevice));
    FRB_CUDA_OK(cudaMemcpy(t.d_emax, emax.data(), kCfTableRows * sizeof(float), cudaMemcpyHostToDevice));
    FRB_CUDA_OK(cudaMemcpy(t.d_absmax, absmax.data(), kCfTableRows * sizeof(float), cudaMemcpyHostToDevice));
    cache[{dev, cell_px}] = t;
    *out = t;
    return FRB_OK;
}

// ---- workspace plan -------------------------------------------------------------------------------------------------
struct CfPlan {
    int64_t pass_q;        // queries per pass
    int cap;
    size_t feat, window, qtot, best, cnt, flag, list_cnt, raw_row, raw_s, list_row, list_s, cand_idx, total;
};

static CfPlan cf_plan(int64_t nq, int64_t ng, int L)
{
    CfPlan pl;
    pl.pass_q = nq < kCfMaxQueriesPerPass ? nq : kCfMaxQueriesPerPass;
    if (pl.pass_q < 1) pl.pass_q = 1;
    pl.cap = ng > 400000 ? 8192 : 4096;
    const size_t qpad = (size_t)((pl.pass_q + kCfBlockM - 1) / kCfBlockM * kCfBlockM);
    const size_t n = (size_t)pl.pass_q * pl.cap;
    size_t o = 0;
    auto take = [&o](size_t bytes) { const size_t at = o; o += align_up(bytes, 1024); return at; };
    pl.feat = take(qpad * (size_t)L * kCfRank * 2);
    pl.window = take(qpad * 4);
    pl.qtot = take(qpad * 4);
    pl.best = take(qpad * 4);
    pl.cnt = take(qpad * 4);
    pl.flag = take(qpad * 4);
    pl.list_cnt = take(qpad * 4);
    pl.raw_row = take(n * 4);
    pl.raw_s = take(n * 4);       // reused as the exact distances of the survivors (cand_dist)
    pl.list_row = take(n * 4);
    pl.list_s = take(n * 4);
    pl.cand_idx = take(n * 8);
    pl.total = o;
    return pl;
}

static float cf_acc_rel()
{
    // Allowance for the fp32 accumulation inside the tensor core, relative to sum_j max_a sum_m |u_m v_m|
    // (measured error of the accumulated sum vs float64 on the same fp16 features, times a margin).
    // The environment override exists for the tests (a huge value forces the overflow -> exact-scan fallback).
    const char *e = getenv("FRB_CHISQ_FILTER_ACC_REL");
    return e ? (float)atof(e) : 1e-4f;
}

}  // namespace frb

using namespace frb;

extern "C" {

int frb_chisq_filter_tables(int cell_px, uint16_t *u_f16, uint16_t *v_f16, float *emax, float *absmax)
{
    FRB_CHECK_ARG(cell_px >= 1 && cell_px <= 255, "frb_chisq_filter_tables: cell_px=%d (1..255)", cell_px);
    FRB_CHECK_ARG(u_f16 && v_f16 && emax && absmax, "frb_chisq_filter_tables: null pointer");
    chisq_filter_build_tables(cell_px, u_f16, v_f16, emax, absmax);
    return FRB_OK;
}

size_t frb_chisq_filter_workspace_bytes(int64_t n_query, int64_t n_gallery, int hist_len)
{
    if (n_query <= 0 || hist_len <= 0) return 0;
    return cf_plan(n_query, n_gallery, hist_len).total;
}

int frb_chisq_top1_filtered_g8(const uint16_t *q_hist, int64_t n_query, const uint8_t *gallery, int64_t n_gallery, int hist_len,
                               int cell_px, int64_t idx_base, float *out_dist, int64_t *out_idx, int *stats,
                               float *approx_scores, void *workspace, size_t workspace_bytes, void *stream)
{
    const char *fn = "frb_chisq_top1_filtered_g8";
    FRB_CHECK_ARG(n_query >= 0 && n_gallery >= 0, "%s: n_query=%lld n_gallery=%lld", fn, (long long)n_query, (long long)n_gallery);
    FRB_CHECK_ARG(hist_len > 0 && hist_len % 16 == 0 && hist_len <= 16384, "%s: hist_len=%d must be a multiple of 16, <= 16384", fn,
                  hist_len);
    FRB_CHECK_ARG(cell_px >= 1 && cell_px <= 255, "%s: cell_px=%d (1..255: u8 counts)", fn, cell_px);
    FRB_CHECK_ARG(n_gallery <= 2147483647LL, "%s: n_gallery too large for one call", fn);
    if (n_query == 0) return FRB_OK;
    FRB_CHECK_ARG(q_hist && out_dist && out_idx, "%s: null pointer", fn);
    FRB_CHECK_ARG(n_gallery == 0 || gallery, "%s: null gallery", fn);
    FRB_CHECK_ARG(((uintptr_t)q_hist & 15) == 0 && ((uintptr_t)gallery & 15) == 0, "%s: histograms must be 16-byte aligned", fn);
    int dev = 0, major = 0;
    FRB_CUDA_OK(cudaGetDevice(&dev));
    FRB_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
        set_error("%s needs an sm_100-class GPU (tcgen05); device is sm_%d", fn, major * 10);
        return FRB_ERR_UNSUPPORTED;
    }
    const CfPlan pl = cf_plan(n_query, n_gallery, hist_len);
    if (!workspace || workspace_bytes < pl.total) {
        set_error("%s: workspace %zu B < %zu B", fn, workspace_bytes, pl.total);
        return FRB_ERR_WORKSPACE;
    }
    FRB_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "%s: workspace must be 256-byte aligned", fn);  // what cudaMalloc and torch's allocator give
    cudaStream_t st = (cudaStream_t)stream;
    CfTables tb;
    int rc = get_tables(cell_px, &tb);
    if (rc != FRB_OK) return rc;
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
        return FRB_ERR_CUDA;
    }
    char *w = (char *)workspace;
    uint4 *feat = (uint4 *)(w + pl.feat);
    float *window = (float *)(w + pl.window), *best = (float *)(w + pl.best), *qtot = (float *)(w + pl.qtot);
    int *cnt = (int *)(w + pl.cnt), *flag = (int *)(w + pl.flag), *list_cnt = (int *)(w + pl.list_cnt);
    int *raw_row = (int *)(w + pl.raw_row), *list_row = (int *)(w + pl.list_row);
    float *raw_s = (float *)(w + pl.raw_s), *list_s = (float *)(w + pl.list_s);
    int64_t *cand_idx = (int64_t *)(w + pl.cand_idx);

    const bool cop8 = cell_px + 1 <= kCfRows8;
    const int table_rows = cop8 ? kCfRows8 : kCfTableRows, copies = cop8 ? 8 : 4;
    const size_t smem = 1024 + (size_t)kCfAStages * kCfAStageBytes + (size_t)kCfBStages * kCfBStageBytes +
                        (size_t)table_rows * copies * sizeof(uint4) + 2 * kCfBlockN * sizeof(float) + sizeof(CfBarriers);
    {
        static thread_local int attr_dev[2] = {-1, -1};
        if (attr_dev[cop8] != dev) {
            if (cop8)
                FRB_CUDA_OK(cudaFuncSetAttribute(chisq_filter_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            else
                FRB_CUDA_OK(cudaFuncSetAttribute(chisq_filter_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_dev[cop8] = dev;
        }
    }
    const int sms = sm_count() > 0 ? sm_count() : 148;
    for (int64_t q0 = 0; q0 < n_query; q0 += pl.pass_q) {
        const int64_t nq = (n_query - q0) < pl.pass_q ? (n_query - q0) : pl.pass_q;
        const uint16_t *qh = q_hist + q0 * hist_len;
        chisq_feat_kernel<<<(unsigned)nq, 256, 0, st>>>(qh, hist_len, cell_px, stats, tb.d_v, tb.d_emax, tb.d_absmax, cf_acc_rel(), feat, window, qtot,
                                                       best, cnt, flag);
        FRB_LAUNCH_OK("chisq_feat_kernel");
        if (n_gallery > 0) {
            CUtensorMap tf;
            const cuuint64_t kdim = (cuuint64_t)hist_len * kCfRank;
            cuuint64_t dims[2] = {kdim, (cuuint64_t)nq};
            cuuint64_t strides[1] = {kdim * 2};
            cuuint32_t box[2] = {(cuuint32_t)kCfBlockK, (cuuint32_t)kCfBlockM};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(&tf, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, feat, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                set_error("%s: cuTensorMapEncodeTiled failed with CUresult %d", fn, (int)r);
                return FRB_ERR_CUDA;
            }
            CfParams p;
            p.n_query = nq;
            p.n_gallery = n_gallery;
            p.hist_len = hist_len;
            p.n_kblocks = hist_len / 8;
            p.n_qtiles = (nq + kCfBlockM - 1) / kCfBlockM;
            p.n_qpairs = (p.n_qtiles + kCfQTiles - 1) / kCfQTiles;
            p.n_gtiles = (n_gallery + kCfBlockN - 1) / kCfBlockN;
            p.gallery = gallery;
            p.u_table = tb.d_u;
            p.table_rows = table_rows;
            p.window = window;
            p.best = best;
            p.cnt = cnt;
            p.cap = pl.cap;
            p.cand_row = raw_row;
            p.cand_s = raw_s;
            p.all_scores = approx_scores ? approx_scores + q0 * n_gallery : nullptr;
            const int64_t n_units = p.n_qpairs * p.n_gtiles;
            const int grid = (int)(n_units < sms ? n_units : sms);
            {
                ProfileScope prof(FRB_K_CHISQ_FILTER, st);
                if (cop8)
                    chisq_filter_kernel<8><<<grid, kCfThreads, smem, st>>>(tf, p);
                else
                    chisq_filter_kernel<4><<<grid, kCfThreads, smem, st>>>(tf, p);
                FRB_LAUNCH_OK("chisq_filter_kernel");
            }
        }
        chisq_survivor_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, st>>>(nq, pl.cap, best, window, cnt, raw_row, raw_s, list_row, list_s,
                                                                        list_cnt, flag, stats);
        FRB_LAUNCH_OK("chisq_survivor_kernel");
        // exact distances of the survivors (cnt[q] = survivors) ...
        rc = chisq_gather_g8(qh, nq, gallery, n_gallery, hist_len, cell_px, idx_base, flag, list_row, list_cnt, pl.cap, raw_s, cand_idx,
                             cnt, st);
        if (rc != FRB_OK) return rc;
        chisq_audit_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, st>>>(nq, pl.cap, 0.5f * (float)cell_px, list_cnt, list_s, raw_s, qtot,
                                                                     window, flag, stats);
        FRB_LAUNCH_OK("chisq_audit_kernel");
        // ... and the plain exact scan for the flagged queries (CTAs of the others exit at once)
        rc = chisq_flagged_topk_g8(qh, nq, gallery, n_gallery, hist_len, cell_px, 1, idx_base, flag, kCfFallbackChunks, pl.cap, raw_s,
                                   cand_idx, cnt, st);
        if (rc != FRB_OK) return rc;
        rc = topk_merge_compact(raw_s, cand_idx, cnt, pl.cap, nq, 1, /*largest=*/0, out_dist + q0, out_idx + q0, st);
        if (rc != FRB_OK) return rc;
    }
    return FRB_OK;
}

}  // extern "C"
