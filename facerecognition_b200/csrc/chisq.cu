// chisq.cu — K3: chi-square (HISTCMP_CHISQR_ALT) nearest-neighbour scan over u16 gallery histograms.
//
// Replaces the compareHist loop inside cv2.face LBPH predict() (reference call sites web_app.py:587,
// models/lbphmodel/inference_lbph.py:5, evaluate_lbph.py:32, threshold_lbph.py:48):
//     d_i = 2 * sum_j (h_ij - q_j)^2 / (h_ij + q_j),  h = count / cell_px,  argmin_i with first-wins ties.
//
// Arithmetic.  With integer counts g (gallery, n_g pixels per cell) and c (query, n_q pixels per cell),
//     d = (2 / n_g) * sum_j (g_j - q~_j)^2 / (g_j + q~_j),   q~ = c * n_g / n_q,
// so the kernel streams raw u16 counts (32 KiB per gallery row instead of OpenCV's 64 KiB of float32).
// Bins with g = q~ = 0 must contribute exactly 0 (OpenCV skips |h+q| <= DBL_EPSILON): q~ is clamped to
// 2^-70, so there d^2 = 2^-140 flushes to zero (mul.ftz) and 0 * rcp(2^-70) = 0.  Identical histograms
// therefore give exactly 0.0, as in OpenCV.  Per bin: PRMT, FADD (u16 -> f32 via the 2^23 trick),
// FADD d, FADD s, FMUL.FTZ, MUFU.RCP, FFMA.
//
// Layout.  One CTA = (query, chunk of gallery rows).  512 threads; thread t keeps the query bins
// {(j*512 + t)*8 .. +8} in registers and reads the same bins of 4 gallery rows at a time with 128-bit
// coalesced loads (a warp covers 512 contiguous bytes of a row per load).  Row partials are reduced
// with a segmented warp butterfly (6 shuffles for 4 rows) and one shared-memory exchange per 32 rows;
// warp 0 keeps the running best-k list.  The query index varies fastest across the grid so CTAs that
// share a gallery chunk run together and re-read it from L2.
#include "frb_common.cuh"

namespace frb {

constexpr int kChiThreads = 512;
constexpr int kChiWarps = kChiThreads / 32;
constexpr int kChiRowsPerGroup = 4;
constexpr int kChiBatch = 32;  // rows per shared-memory exchange

__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ float u16lo_to_f32(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)) - 8388608.0f; }
__device__ __forceinline__ float u16hi_to_f32(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)) - 8388608.0f; }

__device__ __forceinline__ float chi_term(float g, float q)
{
    float d = g - q;
    float s = g + q;
    float dd, r;
    asm("mul.ftz.f32 %0, %1, %1;" : "=f"(dd) : "f"(d));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
    return dd * r;
}

__device__ __forceinline__ float chi_word(uint32_t w, float q0, float q1, float acc)
{
    acc += chi_term(u16lo_to_f32(w), q0);
    acc += chi_term(u16hi_to_f32(w), q1);
    return acc;
}

// CHUNKS: 128-bit loads per thread per row (hist_len <= CHUNKS * 4096).  WRITE_ALL: emit every distance.
template <int CHUNKS, bool WRITE_ALL>
__global__ void __launch_bounds__(kChiThreads, 1)
chisq_kernel(const uint16_t *__restrict__ qhist, int64_t n_query, float q_scale, const uint16_t *__restrict__ gallery,
             int64_t n_gallery, int hist_len, float out_scale, int64_t rows_per_chunk, int k, int64_t idx_base,
             float *__restrict__ cand_dist, int64_t *__restrict__ cand_idx, float *__restrict__ all_dist)
{
    __shared__ float s_part[2][kChiBatch][kChiWarps + 1];
    __shared__ float s_best[FRB_MAX_K];
    __shared__ int64_t s_bidx[FRB_MAX_K];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q = blockIdx.x;
    const int64_t chunk = blockIdx.y;
    const int64_t row_begin = chunk * rows_per_chunk;
    int64_t row_end = row_begin + rows_per_chunk;
    if (row_end > n_gallery) row_end = n_gallery;
    const int vec_per_row = hist_len >> 3;  // uint4 per row

    // query bins -> registers (scaled to the gallery's cell size, clamped away from 0)
    float qv[CHUNKS][8];
    bool live[CHUNKS];
    const float tiny = __uint_as_float(0x1C800000u);  // 2^-70
#pragma unroll
    for (int j = 0; j < CHUNKS; j++) {
        const int v = j * kChiThreads + tid;
        live[j] = v < vec_per_row;
        uint4 w = make_uint4(0, 0, 0, 0);
        if (live[j]) w = __ldg(reinterpret_cast<const uint4 *>(qhist + q * hist_len) + v);
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            qv[j][2 * e] = fmaxf(u16lo_to_f32(ww[e]) * q_scale, tiny);
            qv[j][2 * e + 1] = fmaxf(u16hi_to_f32(ww[e]) * q_scale, tiny);
        }
    }

    float kth = INFINITY;
    if (!WRITE_ALL && warp == 0) {
        if (lane == 0) list_init<false>(s_best, s_bidx, k);
        __syncwarp();
    }

    const uint4 *gal4 = reinterpret_cast<const uint4 *>(gallery);
    int buf = 0;
    for (int64_t base = row_begin; base < row_end; base += kChiBatch, buf ^= 1) {
#pragma unroll 1
        for (int grp = 0; grp < kChiBatch / kChiRowsPerGroup; grp++) {
            const int64_t r0 = base + grp * kChiRowsPerGroup;
            float p[kChiRowsPerGroup];
            uint4 w[kChiRowsPerGroup][CHUNKS];
#pragma unroll
            for (int r = 0; r < kChiRowsPerGroup; r++) {
                const bool rv = (r0 + r) < row_end;
#pragma unroll
                for (int j = 0; j < CHUNKS; j++) {
                    w[r][j] = make_uint4(0, 0, 0, 0);
                    if (rv && live[j]) w[r][j] = ld_stream_u4(gal4 + (r0 + r) * vec_per_row + j * kChiThreads + tid);
                }
            }
#pragma unroll
            for (int r = 0; r < kChiRowsPerGroup; r++) {
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < CHUNKS; j++) {
                    acc = chi_word(w[r][j].x, qv[j][0], qv[j][1], acc);
                    acc = chi_word(w[r][j].y, qv[j][2], qv[j][3], acc);
                    acc = chi_word(w[r][j].z, qv[j][4], qv[j][5], acc);
                    acc = chi_word(w[r][j].w, qv[j][6], qv[j][7], acc);
                }
                p[r] = acc;
            }
            // segmented butterfly: 4 row partials x 32 lanes -> lanes 0/8/16/24 hold rows 0/1/2/3
            const bool hi = lane & 16;
            float k0 = hi ? p[2] : p[0], k1 = hi ? p[3] : p[1];
            float s0 = hi ? p[0] : p[2], s1 = hi ? p[1] : p[3];
            k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
            k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            const bool hi2 = lane & 8;
            float kk = hi2 ? k1 : k0, ss = hi2 ? k0 : k1;
            kk += __shfl_xor_sync(0xffffffffu, ss, 8);
            kk += __shfl_xor_sync(0xffffffffu, kk, 4);
            kk += __shfl_xor_sync(0xffffffffu, kk, 2);
            kk += __shfl_xor_sync(0xffffffffu, kk, 1);
            if ((lane & 7) == 0) s_part[buf][grp * kChiRowsPerGroup + (lane >> 3)][warp] = kk;
        }
        __syncthreads();
        if (warp == 0) {
            const int64_t row = base + lane;
            const bool valid = row < row_end;
            float d = 0.f;
#pragma unroll
            for (int w2 = 0; w2 < kChiWarps; w2++) d += s_part[buf][lane][w2];
            d *= out_scale;
            if (WRITE_ALL) {
                if (valid) all_dist[q * n_gallery + row] = d;
            } else {
                unsigned m = __ballot_sync(0xffffffffu, valid && d < kth);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float v = __shfl_sync(0xffffffffu, d, src);
                    if (v < kth) {
                        float nk = 0.f;
                        if (lane == 0) nk = list_insert_stream<false>(s_best, s_bidx, k, v, idx_base + base + src);
                        kth = __shfl_sync(0xffffffffu, nk, 0);
                    }
                }
            }
        }
        // no second barrier: s_part is double-buffered and warp 0 reaches the next barrier only after reading
    }
    if (!WRITE_ALL && warp == 0) {
        __syncwarp();
        const int64_t o = (chunk * n_query + q) * k;
        for (int j = lane; j < k; j += 32) {
            cand_dist[o + j] = s_best[j];
            cand_idx[o + j] = s_bidx[j];
        }
    }
}

static int64_t chi_chunks(int64_t n_query, int64_t n_gallery, int64_t *rows_per_chunk)
{
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    // aim for ~4 CTAs per SM in total, each at least one 32-row batch, chunk rows a multiple of 32
    int64_t want = ((int64_t)sms * 4 + n_query - 1) / (n_query > 0 ? n_query : 1);
    if (want < 1) want = 1;
    int64_t max_chunks = (n_gallery + kChiBatch - 1) / kChiBatch;
    if (max_chunks < 1) max_chunks = 1;
    if (want > max_chunks) want = max_chunks;
    if (want > 65535) want = 65535;
    int64_t rpc = (n_gallery + want - 1) / want;
    rpc = (rpc + kChiBatch - 1) / kChiBatch * kChiBatch;
    if (rpc < kChiBatch) rpc = kChiBatch;
    *rows_per_chunk = rpc;
    int64_t chunks = (n_gallery + rpc - 1) / rpc;
    return chunks < 1 ? 1 : chunks;
}

template <bool WRITE_ALL>
static int launch_chisq(const uint16_t *qh, int64_t nq, float q_scale, const uint16_t *gal, int64_t ng, int L,
                        float out_scale, int64_t rpc, int64_t chunks, int k, int64_t idx_base, float *cd, int64_t *ci,
                        float *all, cudaStream_t st)
{
    dim3 grid((unsigned)nq, (unsigned)chunks);
    const int chunks_per_thread = (L / 8 + kChiThreads - 1) / kChiThreads;
    ProfileScope prof(FRB_K_CHISQ, st);
    switch (chunks_per_thread) {
        case 1:
            chisq_kernel<1, WRITE_ALL><<<grid, kChiThreads, 0, st>>>(qh, nq, q_scale, gal, ng, L, out_scale, rpc, k, idx_base, cd, ci, all);
            break;
        case 2:
            chisq_kernel<2, WRITE_ALL><<<grid, kChiThreads, 0, st>>>(qh, nq, q_scale, gal, ng, L, out_scale, rpc, k, idx_base, cd, ci, all);
            break;
        case 3:
        case 4:
            chisq_kernel<4, WRITE_ALL><<<grid, kChiThreads, 0, st>>>(qh, nq, q_scale, gal, ng, L, out_scale, rpc, k, idx_base, cd, ci, all);
            break;
        default:
            set_error("chi-square: hist_len=%d exceeds the 16384 bins the kernel keeps in registers", L);
            return FRB_ERR_UNSUPPORTED;
    }
    FRB_LAUNCH_OK("chisq_kernel");
    return FRB_OK;
}

static int check_chisq_args(const char *fn, int64_t nq, int qpx, int64_t ng, int L, int gpx)
{
    FRB_CHECK_ARG(nq >= 0 && ng >= 0, "%s: n_query=%lld n_gallery=%lld", fn, (long long)nq, (long long)ng);
    FRB_CHECK_ARG(L > 0 && (L % 8) == 0, "%s: hist_len=%d must be a positive multiple of 8", fn, L);
    FRB_CHECK_ARG(qpx > 0 && gpx > 0, "%s: cell_px must be positive (q=%d, g=%d)", fn, qpx, gpx);
    FRB_CHECK_ARG(nq <= 2147483647LL, "%s: n_query too large", fn);
    return FRB_OK;
}

}  // namespace frb

using namespace frb;

extern "C" {

size_t frb_chisq_topk_workspace_bytes(int64_t n_query, int64_t n_gallery, int hist_len, int k)
{
    (void)hist_len;
    if (n_query <= 0 || k <= 0) return 0;
    int64_t rpc;
    int64_t chunks = chi_chunks(n_query, n_gallery > 0 ? n_gallery : 1, &rpc);
    size_t n = (size_t)chunks * (size_t)n_query * (size_t)k;
    return align_up(n * sizeof(int64_t), 256) + align_up(n * sizeof(float), 256);
}

int frb_chisq_topk(const uint16_t *q_hist, int64_t n_query, int q_cell_px, const uint16_t *gallery, int64_t n_gallery,
                   int hist_len, int g_cell_px, int k, int64_t idx_base, float *out_dist, int64_t *out_idx,
                   void *workspace, size_t workspace_bytes, void *stream)
{
    int rc = check_chisq_args("frb_chisq_topk", n_query, q_cell_px, n_gallery, hist_len, g_cell_px);
    if (rc != FRB_OK) return rc;
    FRB_CHECK_ARG(k >= 1 && k <= FRB_MAX_K, "frb_chisq_topk: k=%d (1..%d)", k, FRB_MAX_K);
    if (n_query == 0) return FRB_OK;
    FRB_CHECK_ARG(q_hist && out_dist && out_idx, "frb_chisq_topk: null pointer");
    FRB_CHECK_ARG(n_gallery == 0 || gallery, "frb_chisq_topk: null gallery");
    FRB_CHECK_ARG(((uintptr_t)q_hist & 15) == 0 && ((uintptr_t)gallery & 15) == 0,
                  "frb_chisq_topk: histograms must be 16-byte aligned");
    size_t need = frb_chisq_topk_workspace_bytes(n_query, n_gallery, hist_len, k);
    if (!workspace || workspace_bytes < need) {
        set_error("frb_chisq_topk: workspace %zu B < %zu B", workspace_bytes, need);
        return FRB_ERR_WORKSPACE;
    }
    int64_t rpc;
    int64_t chunks = chi_chunks(n_query, n_gallery > 0 ? n_gallery : 1, &rpc);
    size_t n = (size_t)chunks * (size_t)n_query * (size_t)k;
    int64_t *ci = (int64_t *)workspace;
    float *cd = (float *)((char *)workspace + align_up(n * sizeof(int64_t), 256));
    const float q_scale = (float)g_cell_px / (float)q_cell_px;
    const float out_scale = 2.0f / (float)g_cell_px;
    rc = launch_chisq<false>(q_hist, n_query, q_scale, gallery, n_gallery, hist_len, out_scale, rpc, chunks, k, idx_base,
                             cd, ci, nullptr, (cudaStream_t)stream);
    if (rc != FRB_OK) return rc;
    return frb_topk_merge(cd, ci, (int)chunks, n_query, k, /*largest=*/0, out_dist, out_idx, stream);
}

int frb_chisq_dist(const uint16_t *q_hist, int64_t n_query, int q_cell_px, const uint16_t *gallery, int64_t n_gallery,
                   int hist_len, int g_cell_px, float *out_dist, void *stream)
{
    int rc = check_chisq_args("frb_chisq_dist", n_query, q_cell_px, n_gallery, hist_len, g_cell_px);
    if (rc != FRB_OK) return rc;
    if (n_query == 0 || n_gallery == 0) return FRB_OK;
    FRB_CHECK_ARG(q_hist && gallery && out_dist, "frb_chisq_dist: null pointer");
    FRB_CHECK_ARG(((uintptr_t)q_hist & 15) == 0 && ((uintptr_t)gallery & 15) == 0,
                  "frb_chisq_dist: histograms must be 16-byte aligned");
    int64_t rpc;
    int64_t chunks = chi_chunks(n_query, n_gallery, &rpc);
    const float q_scale = (float)g_cell_px / (float)q_cell_px;
    const float out_scale = 2.0f / (float)g_cell_px;
    return launch_chisq<true>(q_hist, n_query, q_scale, gallery, n_gallery, hist_len, out_scale, rpc, chunks, 1, 0, nullptr,
                              nullptr, out_dist, (cudaStream_t)stream);
}

}  // extern "C"
