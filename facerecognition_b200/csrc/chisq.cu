// chisq.cu — K3: chi-square (HISTCMP_CHISQR_ALT) nearest-neighbour scan over u16 gallery histograms.
//
// Replaces the compareHist loop inside cv2.face LBPH predict() (reference call sites web_app.py:587,
// models/lbphmodel/inference_lbph.py:5, evaluate_lbph.py:32, threshold_lbph.py:48):
//     d_i = 2 * sum_j (h_ij - q_j)^2 / (h_ij + q_j),  h = count / cell_px,  argmin_i with first-wins ties.
//
// Arithmetic.  With integer counts g (gallery, n_g pixels per cell) and c (query, n_q pixels per cell),
//     d = (2 / n_g) * sum_j (g_j - q~_j)^2 / (g_j + q~_j),   q~ = c * n_g / n_q,
// so the kernel streams raw u16 counts (32 KiB per gallery row instead of OpenCV's 64 KiB of float32).
// Bins with g = q~ = 0 must contribute exactly 0 (OpenCV skips |h+q| <= DBL_EPSILON): the difference is formed
// from the unclamped query (exactly 0 there) and only the sum is clamped away from 0.  Identical histograms
// therefore give exactly 0.0, as in OpenCV.
//
// Layout.  One CTA = (query, chunk of gallery rows), 16 warps.  Whole gallery rows (contiguous, 32 KiB at 16384
// bins) stream into a ring of shared-memory slots through 1-D TMA bulk copies (mbarrier full/empty per slot);
// the warps take turns issuing them, four rows ahead of the row being consumed.  Thread t keeps the query bins {(j*512 + t)*8 .. +8} in
// registers and reads the same bins of a row with 128-bit shared loads (a warp covers 512 contiguous bytes:
// conflict-free).  Row partials are reduced with a segmented warp butterfly (6 shuffles for 4 rows) and one
// shared-memory exchange per 32 rows; warp 0 keeps the running best-k list.  The query index varies fastest
// across the grid so CTAs that share a gallery chunk run together and re-read it from L2.
//
// Per-bin arithmetic (the kernel is FP32/MUFU-bound once a chunk is shared by many queries): all float ops are
// packed f32x2, and ONE reciprocal serves two bins:  a^2/s + c^2/t = (a^2 t + c^2 s) * rcp(s t).
// Empty-empty bins use s = 2^-40 (so s t >= 2^-80 stays normal) and an exact d = 0, hence contribute exactly 0.
// When both histograms have the same cell size (INTQ, the usual case) counts are integers and
//     d = (2^23 + g) - (2^23 + c)   and   s = d + max(2c, 2^-40)
// are exact without a separate u16 -> f32 conversion.
#include "frb_common.cuh"

namespace frb {

constexpr int kChiThreads = 512;
constexpr int kChiWarps = kChiThreads / 32;
constexpr int kChiBlock = kChiThreads;
constexpr int kChiRowsPerGroup = 4;
constexpr int kChiBatch = 24;  // rows per shared-memory exchange (two 12-row blocks)
constexpr int kChiSlots = 6;   // shared-memory ring of whole rows
constexpr int kChiMinRows = 8; // fewest gallery rows worth a CTA (it first loads the 32 KB query into registers)

__device__ __forceinline__ uint32_t chi_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void chi_mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "CHI_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra CHI_DONE;\n\t"
        "bra CHI_WAIT;\n\t"
        "CHI_DONE:\n\t"
        "}" ::"r"(chi_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void chi_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(chi_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void chi_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(chi_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(chi_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(chi_smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ float2 chi_f2(float x) { return make_float2(x, x); }
// the two u16 counts of a word as 2^23 + count (exact): the count sits in the mantissa of 2^23
__device__ __forceinline__ float2 chi_magic(uint32_t w)
{
    return make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)), __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)));
}
__device__ __forceinline__ float chi_rcp(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// the four u8 counts of a word as 2^23 + count: bins (0, 1) and (2, 3)
__device__ __forceinline__ void chi_magic8(uint32_t w, float2 &ga, float2 &gb)
{
    ga = make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540)), __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7541)));
    gb = make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7542)), __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7543)));
}

// Four bins with gallery counts ga (bins 0,1) and gb (bins 2,3) given as 2^23 + count.  INTQ: qa/qb = -(2^23 + c),
// ta/tb = max(2c, 2^-40); otherwise qa/qb = -c~ and ta/tb = max(c~, 2^-40) with c~ the query count rescaled to the
// gallery's cell size.
template <bool INTQ>
__device__ __forceinline__ float2 chi_quad_core(float2 ga, float2 gb, float2 qa, float2 ta, float2 qb, float2 tb, float2 acc)
{
    float2 da, db, sa, sb;
    if (INTQ) {
        da = __fadd2_rn(ga, qa);
        db = __fadd2_rn(gb, qb);
        sa = __fadd2_rn(da, ta);
        sb = __fadd2_rn(db, tb);
    } else {
        ga = __fadd2_rn(ga, chi_f2(-8388608.0f));
        gb = __fadd2_rn(gb, chi_f2(-8388608.0f));
        da = __fadd2_rn(ga, qa);
        db = __fadd2_rn(gb, qb);
        sa = __fadd2_rn(ga, ta);
        sb = __fadd2_rn(gb, tb);
    }
    float2 u = __fmul2_rn(__fmul2_rn(da, da), sb);
    u = __ffma2_rn(__fmul2_rn(db, db), sa, u);
    const float2 prod = __fmul2_rn(sa, sb);
    const float2 r = make_float2(chi_rcp(prod.x), chi_rcp(prod.y));
    return __ffma2_rn(u, r, acc);
}
// u16 gallery: words wa (bins 0,1) and wb (bins 2,3)
template <bool INTQ>
__device__ __forceinline__ float2 chi_quad(uint32_t wa, uint32_t wb, float2 qa, float2 ta, float2 qb, float2 tb, float2 acc)
{
    return chi_quad_core<INTQ>(chi_magic(wa), chi_magic(wb), qa, ta, qb, tb, acc);
}
// u8 gallery: one word holds the four bins
template <bool INTQ>
__device__ __forceinline__ float2 chi_quad8(uint32_t w, float2 qa, float2 ta, float2 qb, float2 tb, float2 acc)
{
    float2 ga, gb;
    chi_magic8(w, ga, gb);
    return chi_quad_core<INTQ>(ga, gb, qa, ta, qb, tb, acc);
}

// Optional per-query selection (all-null = the plain scan); see chisq_kernel.
struct ChiSelect {
    const int *q_flag;     // [n_query] or null
    const int *row_list;   // GATHER: [n_query, cap] gallery rows to re-score
    const int *row_cnt;    // GATHER: [n_query] rows listed
    int64_t cap;           // candidate slots per query in cand_dist / cand_idx
    int *cnt_out;          // [n_query] candidates written per query (read by topk_merge_compact)
};

// ---- twelve rows = two trips round the 6-slot ring, so slots and mbarrier parities are compile-time constants -----
// Row i of the block (i = 0..11) lives in slot i % 6 during ring phase (i / 6) & 1.  While row i is consumed, warp i
// requests row i + 4 (slot (i + 4) % 6) once every warp has released that slot's previous tenant (row i - 2).
// GUARD: the block may run past the chunk's last row (tail block only).  FULL: every thread owns CHUNKS full groups.
// G8: the gallery holds u8 counts (16 bins per 128-bit group, E = 8 bin pairs) instead of u16 (8 bins, E = 4).
// GATHER: the rows to scan are rows[0 .. n_rows) of the whole gallery (gal_chunk = its base) instead of a contiguous chunk.
template <int CHUNKS, bool INTQ, bool FULL, bool GUARD, bool G8, bool GATHER>
__device__ __forceinline__ void chi_block12(const unsigned char *ring, uint64_t *s_full, uint64_t *s_empty, uint32_t row_bytes,
                                            const unsigned char *__restrict__ gal_chunk, const int *__restrict__ rows, int64_t it0,
                                            int64_t n_rows,
                                            const float2 (&qd)[CHUNKS][G8 ? 8 : 4], const float2 (&qs)[CHUNKS][G8 ? 8 : 4],
                                            const bool (&live)[CHUNKS], float (*part)[kChiWarps + 1], int part_row0)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int g = 0; g < 3; g++) {
        float p[kChiRowsPerGroup];
#pragma unroll
        for (int r = 0; r < kChiRowsPerGroup; r++) {
            const int i = g * 4 + r;
            p[r] = 0.f;
            if (GUARD && it0 + i >= n_rows) continue;
            if (tid == i * 32) {
                const int j = i + 4;                       // row to request, relative to the block
                if (it0 + j < n_rows) {   // also in a full block: rows 12..15 belong to the NEXT block, which may not exist
                    const int ns = j % 6;
                    const uint32_t par = (j < 6 || j >= 12) ? 1u : 0u;   // parity of the phase that released the slot
                    if (j >= 6 || it0 > 0) chi_mbar_wait(&s_empty[ns], par);
                    const int64_t src_row = GATHER ? (int64_t)rows[it0 + j] : it0 + j;
                    chi_bulk_load(const_cast<unsigned char *>(ring) + (size_t)ns * row_bytes, gal_chunk + src_row * (int64_t)row_bytes,
                                  row_bytes, &s_full[ns]);
                }
            }
            __syncwarp();
            const int slot = i % 6;
            chi_mbar_wait(&s_full[slot], (uint32_t)((i / 6) & 1));
            const uint4 *row4 = reinterpret_cast<const uint4 *>(ring + (size_t)slot * row_bytes);
            uint4 w[CHUNKS];
#pragma unroll
            for (int jj = 0; jj < CHUNKS; jj++) {
                if (FULL || live[jj])
                    w[jj] = row4[jj * kChiThreads + tid];
                else
                    w[jj] = make_uint4(0, 0, 0, 0);
            }
            __syncwarp();
            if (lane == 0) chi_mbar_arrive(&s_empty[slot]);  // this warp holds its part of the row in registers
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int jj = 0; jj < CHUNKS; jj++) {
                if (G8) {
                    acc = chi_quad8<INTQ>(w[jj].x, qd[jj][0], qs[jj][0], qd[jj][1], qs[jj][1], acc);
                    acc = chi_quad8<INTQ>(w[jj].y, qd[jj][2], qs[jj][2], qd[jj][3], qs[jj][3], acc);
                    acc = chi_quad8<INTQ>(w[jj].z, qd[jj][G8 ? 4 : 0], qs[jj][G8 ? 4 : 0], qd[jj][G8 ? 5 : 1], qs[jj][G8 ? 5 : 1], acc);
                    acc = chi_quad8<INTQ>(w[jj].w, qd[jj][G8 ? 6 : 2], qs[jj][G8 ? 6 : 2], qd[jj][G8 ? 7 : 3], qs[jj][G8 ? 7 : 3], acc);
                } else {
                    acc = chi_quad<INTQ>(w[jj].x, w[jj].y, qd[jj][0], qs[jj][0], qd[jj][1], qs[jj][1], acc);
                    acc = chi_quad<INTQ>(w[jj].z, w[jj].w, qd[jj][2], qs[jj][2], qd[jj][3], qs[jj][3], acc);
                }
            }
            p[r] = acc.x + acc.y;
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
        if ((lane & 7) == 0) part[part_row0 + g * kChiRowsPerGroup + (lane >> 3)][warp] = kk;
    }
}

// CHUNKS: 128-bit groups per thread per row (hist_len <= CHUNKS * 4096 for a u16 gallery, CHUNKS * 8192 for u8).
// WRITE_ALL: emit every distance.  G8: gallery of u8 counts (cell_px <= 255): half the bytes per row, same arithmetic.
// GATHER (behind the tensor-core candidate filter, chisq_filter.cu): query q scans only the rows listed in
// sel.row_list[q * cap ..] and writes one exact distance per listed row to the compact candidate buffer
// cand_dist / cand_idx [q * cap + i]; the arithmetic and the summation order are those of the full scan, so a listed
// row gets bit for bit the distance the full scan gives it.  sel.q_flag: GATHER skips flagged queries; the plain
// top-k scan, when given q_flag, runs ONLY the flagged ones (the filter's exact fallback) and writes chunk lists to
// cand_dist / cand_idx [q * cap + chunk * k ..].
template <int CHUNKS, bool WRITE_ALL, bool INTQ, bool FULL, bool G8, bool GATHER>
__global__ void __launch_bounds__(kChiBlock, 1)
chisq_kernel(const uint16_t *__restrict__ qhist, int64_t n_query, float q_scale, const void *__restrict__ gallery_v,
             int64_t n_gallery, int hist_len, float out_scale, int64_t rows_per_chunk, int k, int64_t idx_base,
             float *__restrict__ cand_dist, int64_t *__restrict__ cand_idx, float *__restrict__ all_dist, const ChiSelect sel)
{
    extern __shared__ __align__(128) unsigned char chi_smem[];
    __shared__ float s_part[2][kChiBatch][kChiWarps + 1];
    __shared__ float s_best[FRB_MAX_K];
    __shared__ int64_t s_bidx[FRB_MAX_K];
    __shared__ __align__(8) uint64_t s_full[kChiSlots], s_empty[kChiSlots];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q = blockIdx.x;
    const int64_t chunk = blockIdx.y;
    if (sel.q_flag) {   // uniform over the CTA
        const bool flagged = sel.q_flag[q] != 0;
        if (GATHER ? flagged : !flagged) return;
    }
    const int64_t row_begin = GATHER ? 0 : chunk * rows_per_chunk;
    int64_t row_end = row_begin + rows_per_chunk;
    if (row_end > n_gallery) row_end = n_gallery;
    const int *rows = nullptr;
    if (GATHER) {
        int64_t n = sel.row_cnt[q];
        row_end = n < sel.cap ? n : sel.cap;
        rows = sel.row_list + q * sel.cap;
    }
    const int64_t n_rows = row_end - row_begin;
    constexpr int E = G8 ? 8 : 4;                          // bin pairs per 128-bit gallery group
    const int vec_per_row = hist_len >> (G8 ? 4 : 3);      // uint4 per gallery row
    const uint32_t row_bytes = (uint32_t)hist_len * (G8 ? 1u : 2u);
    const unsigned char *gal_chunk = reinterpret_cast<const unsigned char *>(gallery_v) + row_begin * (int64_t)row_bytes;

    if (tid == 0) {
        for (int i = 0; i < kChiSlots; i++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(chi_smem_u32(&s_full[i])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(chi_smem_u32(&s_empty[i])), "n"(kChiWarps));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // the first four rows are requested up front; after that warp i of a 12-row block requests row i + 4
    if (tid == 0) {
        for (int64_t i = 0; i < 4 && i < n_rows; i++)
            chi_bulk_load(chi_smem + (size_t)i * row_bytes, gal_chunk + (GATHER ? (int64_t)rows[i] : i) * (int64_t)row_bytes, row_bytes,
                          &s_full[i]);
    }

    // ---- query bins -> registers ------------------------------------------------------------------------------
    // (the query is always u16, as K2 writes it; a u8 gallery group of 16 bins takes two 128-bit query loads)
    float2 qd[CHUNKS][E], qs[CHUNKS][E];
    bool live[CHUNKS];
    const float tiny = __uint_as_float(0x2B800000u);  // 2^-40
#pragma unroll
    for (int j = 0; j < CHUNKS; j++) {
        const int v = j * kChiThreads + tid;
        live[j] = FULL || v < vec_per_row;
        uint32_t ww[E];
#pragma unroll
        for (int h = 0; h < E / 4; h++) {
            uint4 w = make_uint4(0, 0, 0, 0);
            if (live[j]) w = __ldg(reinterpret_cast<const uint4 *>(qhist + q * hist_len) + v * (E / 4) + h);
            ww[4 * h] = w.x;
            ww[4 * h + 1] = w.y;
            ww[4 * h + 2] = w.z;
            ww[4 * h + 3] = w.w;
        }
#pragma unroll
        for (int e = 0; e < E; e++) {
            const float c0 = (float)(ww[e] & 0xFFFFu), c1 = (float)(ww[e] >> 16);
            if (INTQ) {
                qd[j][e] = make_float2(-(c0 + 8388608.0f), -(c1 + 8388608.0f));
                qs[j][e] = make_float2(fmaxf(2.0f * c0, tiny), fmaxf(2.0f * c1, tiny));
            } else {
                qd[j][e] = make_float2(-(c0 * q_scale), -(c1 * q_scale));
                qs[j][e] = make_float2(fmaxf(c0 * q_scale, tiny), fmaxf(c1 * q_scale, tiny));
            }
        }
    }

    float kth = INFINITY;
    if (!WRITE_ALL && warp == 0) {
        if (lane == 0) list_init<false>(s_best, s_bidx, k);
        __syncwarp();
    }

    int buf = 0;
    for (int64_t it0 = 0; it0 < n_rows; it0 += kChiBatch, buf ^= 1) {
        // a batch = two 12-row blocks (24 rows, one per lane of warp 0 in the final step)
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            const int64_t b0 = it0 + h * 12;
            if (b0 + 12 <= n_rows)
                chi_block12<CHUNKS, INTQ, FULL, false, G8, GATHER>(chi_smem, s_full, s_empty, row_bytes, gal_chunk, rows, b0, n_rows, qd, qs,
                                                                   live, s_part[buf], h * 12);
            else if (b0 < n_rows)
                chi_block12<CHUNKS, INTQ, FULL, true, G8, GATHER>(chi_smem, s_full, s_empty, row_bytes, gal_chunk, rows, b0, n_rows, qd, qs,
                                                                  live, s_part[buf], h * 12);
        }
        __syncthreads();
        if (warp == 0) {
            const int64_t row = row_begin + it0 + lane;
            const bool valid = lane < kChiBatch && row < row_end;
            float d = 0.f;
#pragma unroll
            for (int w2 = 0; w2 < kChiWarps; w2++) d += s_part[buf][lane < kChiBatch ? lane : 0][w2];
            d *= out_scale;
            if (WRITE_ALL) {
                if (valid) all_dist[q * n_gallery + row] = d;
            } else if (GATHER) {
                if (valid) {
                    cand_dist[q * sel.cap + row] = d;                       // row = position in the query's list
                    cand_idx[q * sel.cap + row] = idx_base + rows[row];
                }
            } else {
                unsigned m = __ballot_sync(0xffffffffu, valid && d < kth);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float v = __shfl_sync(0xffffffffu, d, src);
                    if (v < kth) {
                        float nk = 0.f;
                        if (lane == 0) nk = list_insert_stream<false>(s_best, s_bidx, k, v, idx_base + row_begin + it0 + src);
                        kth = __shfl_sync(0xffffffffu, nk, 0);
                    }
                }
            }
        }
        // no second barrier: s_part is double-buffered and warp 0 reaches the next barrier only after reading
    }
    if (GATHER) {
        if (tid == 0) sel.cnt_out[q] = (int)n_rows;
    } else if (!WRITE_ALL && warp == 0) {
        __syncwarp();
        const int64_t o = sel.q_flag ? q * sel.cap + chunk * k : (chunk * n_query + q) * k;
        for (int j = lane; j < k; j += 32) {
            cand_dist[o + j] = s_best[j];
            cand_idx[o + j] = s_bidx[j];
        }
        if (sel.q_flag && lane == 0 && chunk == 0) sel.cnt_out[q] = (int)gridDim.y * k;
    }
}

static int64_t chi_chunks(int64_t n_query, int64_t n_gallery, int64_t *rows_per_chunk)
{
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    // ~4 CTAs per SM in total (a whole number of waves when there are few queries), each at least kChiMinRows rows:
    // a single predict() against a 1000-row gallery ran 42 CTAs of 24 rows in 30 us; 125 CTAs of 8 rows fill the GPU
    int64_t want = ((int64_t)sms * 4 + n_query - 1) / (n_query > 0 ? n_query : 1);
    if (want < 1) want = 1;
    int64_t max_chunks = (n_gallery + kChiMinRows - 1) / kChiMinRows;
    if (max_chunks < 1) max_chunks = 1;
    if (want > max_chunks) want = max_chunks;
    if (want > 65535) want = 65535;
    int64_t rpc = (n_gallery + want - 1) / want;
    if (rpc < kChiMinRows) rpc = kChiMinRows;
    *rows_per_chunk = rpc;
    int64_t chunks = (n_gallery + rpc - 1) / rpc;
    return chunks < 1 ? 1 : chunks;
}

template <int CHUNKS, bool WRITE_ALL, bool INTQ, bool FULL, bool G8, bool GATHER = false>
static int launch_chisq_t(dim3 grid, size_t smem, const uint16_t *qh, int64_t nq, float q_scale, const void *gal, int64_t ng,
                          int L, float out_scale, int64_t rpc, int k, int64_t idx_base, float *cd, int64_t *ci, float *all,
                          cudaStream_t st, const ChiSelect sel = ChiSelect{nullptr, nullptr, nullptr, 0, nullptr})
{
    static thread_local int attr_dev = -1;
    static thread_local size_t attr_smem = 0;
    int dev = 0;
    FRB_CUDA_OK(cudaGetDevice(&dev));
    if (attr_dev != dev || attr_smem < smem) {
        FRB_CUDA_OK(cudaFuncSetAttribute(chisq_kernel<CHUNKS, WRITE_ALL, INTQ, FULL, G8, GATHER>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_dev = dev;
        attr_smem = smem;
    }
    ProfileScope prof(FRB_K_CHISQ, st);
    chisq_kernel<CHUNKS, WRITE_ALL, INTQ, FULL, G8, GATHER><<<grid, kChiBlock, smem, st>>>(qh, nq, q_scale, gal, ng, L, out_scale, rpc,
                                                                                          k, idx_base, cd, ci, all, sel);
    FRB_LAUNCH_OK("chisq_kernel");
    return FRB_OK;
}

template <int CHUNKS, bool WRITE_ALL, bool G8>
static int launch_chisq_c(bool intq, bool full, dim3 grid, size_t smem, const uint16_t *qh, int64_t nq, float q_scale,
                          const void *gal, int64_t ng, int L, float out_scale, int64_t rpc, int k, int64_t idx_base, float *cd,
                          int64_t *ci, float *all, cudaStream_t st)
{
#define FRB_CHI_ARGS grid, smem, qh, nq, q_scale, gal, ng, L, out_scale, rpc, k, idx_base, cd, ci, all, st
    if (intq)
        return full ? launch_chisq_t<CHUNKS, WRITE_ALL, true, true, G8>(FRB_CHI_ARGS)
                    : launch_chisq_t<CHUNKS, WRITE_ALL, true, false, G8>(FRB_CHI_ARGS);
    return full ? launch_chisq_t<CHUNKS, WRITE_ALL, false, true, G8>(FRB_CHI_ARGS)
                : launch_chisq_t<CHUNKS, WRITE_ALL, false, false, G8>(FRB_CHI_ARGS);
#undef FRB_CHI_ARGS
}

// gallery_bytes: 2 (u16 counts) or 1 (u8 counts, cell_px <= 255)
template <bool WRITE_ALL, bool G8>
static int launch_chisq_g(const uint16_t *qh, int64_t nq, int q_cell_px, const void *gal, int64_t ng, int L, int g_cell_px,
                          int64_t rpc, int64_t chunks, int k, int64_t idx_base, float *cd, int64_t *ci, float *all,
                          cudaStream_t st)
{
    dim3 grid((unsigned)nq, (unsigned)chunks);
    const int bins_per_group = G8 ? 16 : 8;                 // bins in one 128-bit gallery load
    const int groups = L / bins_per_group;
    const int chunks_per_thread = (groups + kChiThreads - 1) / kChiThreads;
    const float q_scale = (float)g_cell_px / (float)q_cell_px;
    const float out_scale = 2.0f / (float)g_cell_px;
    const bool intq = q_cell_px == g_cell_px;  // integer counts on both sides: exact differences without a conversion
    const size_t smem = (size_t)L * (G8 ? 1 : 2) * kChiSlots;  // ring of whole rows
    int c = chunks_per_thread == 3 ? 4 : chunks_per_thread;
    const bool full = groups == c * kChiThreads;
#define FRB_CHI_CALL(C) \
    launch_chisq_c<C, WRITE_ALL, G8>(intq, full, grid, smem, qh, nq, q_scale, gal, ng, L, out_scale, rpc, k, idx_base, cd, ci, all, st)
    if (c == 1) return FRB_CHI_CALL(1);
    if (c == 2) return FRB_CHI_CALL(2);
    if constexpr (!G8) {                                    // 4 x 16 bins per thread would not fit the register file
        if (c == 4) return FRB_CHI_CALL(4);
    }
#undef FRB_CHI_CALL
    set_error("chi-square: hist_len=%d exceeds the 16384 bins the kernel keeps in registers", L);
    return FRB_ERR_UNSUPPORTED;
}

template <bool WRITE_ALL>
static int launch_chisq(const uint16_t *qh, int64_t nq, int q_cell_px, const void *gal, int gallery_bytes, int64_t ng, int L,
                        int g_cell_px, int64_t rpc, int64_t chunks, int k, int64_t idx_base, float *cd, int64_t *ci, float *all,
                        cudaStream_t st)
{
    if (gallery_bytes == 1)
        return launch_chisq_g<WRITE_ALL, true>(qh, nq, q_cell_px, gal, ng, L, g_cell_px, rpc, chunks, k, idx_base, cd, ci, all, st);
    return launch_chisq_g<WRITE_ALL, false>(qh, nq, q_cell_px, gal, ng, L, g_cell_px, rpc, chunks, k, idx_base, cd, ci, all, st);
}

// ---- entry points for the candidate filter (chisq_filter.cu): u8 gallery, equal cell sizes --------------------------
// exact distances of the listed rows -> cand_dist / cand_idx [q * cap + i], cnt_out[q] = rows listed; flagged queries skipped
int chisq_gather_g8(const uint16_t *qh, int64_t nq, const uint8_t *gal, int64_t ng, int L, int cell_px, int64_t idx_base,
                    const int *q_flag, const int *row_list, const int *row_cnt, int64_t cap, float *cd, int64_t *ci, int *cnt_out,
                    cudaStream_t st)
{
    const int groups = L / 16;
    const int c = (groups + kChiThreads - 1) / kChiThreads;
    const bool full = groups == c * kChiThreads;
    const size_t smem = (size_t)L * kChiSlots;
    const dim3 grid((unsigned)nq, 1);
    const ChiSelect sel{q_flag, row_list, row_cnt, cap, cnt_out};
    const float out_scale = 2.0f / (float)cell_px;
#define FRB_CHI_GATHER(C, F) \
    launch_chisq_t<C, false, true, F, true, true>(grid, smem, qh, nq, 1.0f, gal, ng, L, out_scale, ng, 1, idx_base, cd, ci, nullptr, st, sel)
    if (c == 1) return full ? FRB_CHI_GATHER(1, true) : FRB_CHI_GATHER(1, false);
    if (c == 2) return full ? FRB_CHI_GATHER(2, true) : FRB_CHI_GATHER(2, false);
#undef FRB_CHI_GATHER
    set_error("chi-square gather: hist_len=%d exceeds the 16384 bins the kernel keeps in registers", L);
    return FRB_ERR_UNSUPPORTED;
}

// the plain exact top-k scan for the flagged queries only, `chunks` chunk lists per query -> cand [q * cap + chunk * k ..]
int chisq_flagged_topk_g8(const uint16_t *qh, int64_t nq, const uint8_t *gal, int64_t ng, int L, int cell_px, int k,
                          int64_t idx_base, const int *q_flag, int chunks, int64_t cap, float *cd, int64_t *ci, int *cnt_out,
                          cudaStream_t st)
{
    const int groups = L / 16;
    const int c = (groups + kChiThreads - 1) / kChiThreads;
    const bool full = groups == c * kChiThreads;
    const size_t smem = (size_t)L * kChiSlots;
    int64_t rpc = (ng + chunks - 1) / chunks;
    if (rpc < 1) rpc = 1;
    const dim3 grid((unsigned)nq, (unsigned)chunks);
    const ChiSelect sel{q_flag, nullptr, nullptr, cap, cnt_out};
    const float out_scale = 2.0f / (float)cell_px;
#define FRB_CHI_FLAGGED(C, F) \
    launch_chisq_t<C, false, true, F, true, false>(grid, smem, qh, nq, 1.0f, gal, ng, L, out_scale, rpc, k, idx_base, cd, ci, nullptr, st, sel)
    if (c == 1) return full ? FRB_CHI_FLAGGED(1, true) : FRB_CHI_FLAGGED(1, false);
    if (c == 2) return full ? FRB_CHI_FLAGGED(2, true) : FRB_CHI_FLAGGED(2, false);
#undef FRB_CHI_FLAGGED
    set_error("chi-square: hist_len=%d exceeds the 16384 bins the kernel keeps in registers", L);
    return FRB_ERR_UNSUPPORTED;
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

static int chisq_topk_impl(const char *fn, const uint16_t *q_hist, int64_t n_query, int q_cell_px, const void *gallery,
                           int gallery_bytes, int64_t n_gallery, int hist_len, int g_cell_px, int k, int64_t idx_base,
                           float *out_dist, int64_t *out_idx, void *workspace, size_t workspace_bytes, void *stream)
{
    int rc = check_chisq_args(fn, n_query, q_cell_px, n_gallery, hist_len, g_cell_px);
    if (rc != FRB_OK) return rc;
    FRB_CHECK_ARG(k >= 1 && k <= FRB_MAX_K, "%s: k=%d (1..%d)", fn, k, FRB_MAX_K);
    FRB_CHECK_ARG(gallery_bytes == 2 || (hist_len % 16 == 0 && g_cell_px <= 255),
                  "%s: a u8 gallery needs hist_len %% 16 == 0 and cell_px <= 255 (hist_len=%d, cell_px=%d)", fn, hist_len, g_cell_px);
    if (n_query == 0) return FRB_OK;
    FRB_CHECK_ARG(q_hist && out_dist && out_idx, "%s: null pointer", fn);
    FRB_CHECK_ARG(n_gallery == 0 || gallery, "%s: null gallery", fn);
    FRB_CHECK_ARG(((uintptr_t)q_hist & 15) == 0 && ((uintptr_t)gallery & 15) == 0, "%s: histograms must be 16-byte aligned", fn);
    size_t need = frb_chisq_topk_workspace_bytes(n_query, n_gallery, hist_len, k);
    if (!workspace || workspace_bytes < need) {
        set_error("%s: workspace %zu B < %zu B", fn, workspace_bytes, need);
        return FRB_ERR_WORKSPACE;
    }
    int64_t rpc;
    int64_t chunks = chi_chunks(n_query, n_gallery > 0 ? n_gallery : 1, &rpc);
    size_t n = (size_t)chunks * (size_t)n_query * (size_t)k;
    int64_t *ci = (int64_t *)workspace;
    float *cd = (float *)((char *)workspace + align_up(n * sizeof(int64_t), 256));
    rc = launch_chisq<false>(q_hist, n_query, q_cell_px, gallery, gallery_bytes, n_gallery, hist_len, g_cell_px, rpc, chunks, k,
                             idx_base, cd, ci, nullptr, (cudaStream_t)stream);
    if (rc != FRB_OK) return rc;
    return frb_topk_merge(cd, ci, (int)chunks, n_query, k, /*largest=*/0, out_dist, out_idx, stream);
}

static int chisq_dist_impl(const char *fn, const uint16_t *q_hist, int64_t n_query, int q_cell_px, const void *gallery,
                           int gallery_bytes, int64_t n_gallery, int hist_len, int g_cell_px, float *out_dist, void *stream)
{
    int rc = check_chisq_args(fn, n_query, q_cell_px, n_gallery, hist_len, g_cell_px);
    if (rc != FRB_OK) return rc;
    FRB_CHECK_ARG(gallery_bytes == 2 || (hist_len % 16 == 0 && g_cell_px <= 255),
                  "%s: a u8 gallery needs hist_len %% 16 == 0 and cell_px <= 255 (hist_len=%d, cell_px=%d)", fn, hist_len, g_cell_px);
    if (n_query == 0 || n_gallery == 0) return FRB_OK;
    FRB_CHECK_ARG(q_hist && gallery && out_dist, "%s: null pointer", fn);
    FRB_CHECK_ARG(((uintptr_t)q_hist & 15) == 0 && ((uintptr_t)gallery & 15) == 0, "%s: histograms must be 16-byte aligned", fn);
    int64_t rpc;
    int64_t chunks = chi_chunks(n_query, n_gallery, &rpc);
    return launch_chisq<true>(q_hist, n_query, q_cell_px, gallery, gallery_bytes, n_gallery, hist_len, g_cell_px, rpc, chunks, 1, 0,
                              nullptr, nullptr, out_dist, (cudaStream_t)stream);
}

int frb_chisq_topk(const uint16_t *q_hist, int64_t n_query, int q_cell_px, const uint16_t *gallery, int64_t n_gallery,
                   int hist_len, int g_cell_px, int k, int64_t idx_base, float *out_dist, int64_t *out_idx,
                   void *workspace, size_t workspace_bytes, void *stream)
{
    return chisq_topk_impl("frb_chisq_topk", q_hist, n_query, q_cell_px, gallery, 2, n_gallery, hist_len, g_cell_px, k, idx_base,
                           out_dist, out_idx, workspace, workspace_bytes, stream);
}

int frb_chisq_topk_g8(const uint16_t *q_hist, int64_t n_query, int q_cell_px, const uint8_t *gallery, int64_t n_gallery,
                      int hist_len, int g_cell_px, int k, int64_t idx_base, float *out_dist, int64_t *out_idx,
                      void *workspace, size_t workspace_bytes, void *stream)
{
    return chisq_topk_impl("frb_chisq_topk_g8", q_hist, n_query, q_cell_px, gallery, 1, n_gallery, hist_len, g_cell_px, k,
                           idx_base, out_dist, out_idx, workspace, workspace_bytes, stream);
}

int frb_chisq_dist(const uint16_t *q_hist, int64_t n_query, int q_cell_px, const uint16_t *gallery, int64_t n_gallery,
                   int hist_len, int g_cell_px, float *out_dist, void *stream)
{
    return chisq_dist_impl("frb_chisq_dist", q_hist, n_query, q_cell_px, gallery, 2, n_gallery, hist_len, g_cell_px, out_dist,
                           stream);
}

int frb_chisq_dist_g8(const uint16_t *q_hist, int64_t n_query, int q_cell_px, const uint8_t *gallery, int64_t n_gallery,
                      int hist_len, int g_cell_px, float *out_dist, void *stream)
{
    return chisq_dist_impl("frb_chisq_dist_g8", q_hist, n_query, q_cell_px, gallery, 1, n_gallery, hist_len, g_cell_px, out_dist,
                           stream);
}

}  // extern "C"
