// core.cu — library plumbing, row norms / normalisation, candidate-list merge.
#include "frb_common.cuh"

#include <mutex>
#include <utility>
#include <vector>

namespace frb {

// ---- launch profiling (frb_profile_enable / frb_profile_read) ------------------------------------
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof[FRB_K_COUNT];

ProfileScope::ProfileScope(int kernel_id, cudaStream_t st) : kernel(kernel_id), stream(st), start(nullptr), stop(nullptr), on(false)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_on) return;
    if (cudaEventCreate(&start) != cudaSuccess || cudaEventCreate(&stop) != cudaSuccess) return;
    on = true;
    cudaEventRecord(start, stream);
}

ProfileScope::~ProfileScope()
{
    if (!on) return;
    cudaEventRecord(stop, stream);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof[kernel].push_back({start, stop});
}

static thread_local char g_err[512] = "";

char *last_error_buf() { return g_err; }

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count()
{
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev != cached_dev) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
        cached = p.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}

// ---------------------------------------------------------------------------------------------
// One warp per row: sum of squares with 128-bit loads + shuffle reduction.
// sqrt of the fp32 sum of squares, accumulated pairwise in fp32 (numpy's nrm2 is also fp32).
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float row_sumsq(const float *__restrict__ row, int dim, int lane)
{
    float acc = 0.f;
    if ((dim & 3) == 0 && ((uintptr_t)row & 15) == 0) {
        const float4 *r4 = reinterpret_cast<const float4 *>(row);
        for (int i = lane; i < dim / 4; i += 32) {
            float4 v = __ldg(r4 + i);
            acc = fmaf(v.x, v.x, acc);
            acc = fmaf(v.y, v.y, acc);
            acc = fmaf(v.z, v.z, acc);
            acc = fmaf(v.w, v.w, acc);
        }
    } else {
        for (int i = lane; i < dim; i += 32) {
            float v = __ldg(row + i);
            acc = fmaf(v, v, acc);
        }
    }
    return warp_sum(acc);
}

__global__ void __launch_bounds__(256) row_norms_kernel(const float *__restrict__ x, int64_t rows, int dim,
                                                        float *__restrict__ out)
{
    int lane = threadIdx.x & 31;
    int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp; r < rows; r += nwarps) {
        float ss = row_sumsq(x + r * dim, dim, lane);
        if (lane == 0) out[r] = sqrtf(ss);
    }
}

template <typename OutT>
__device__ __forceinline__ OutT cast_out(float v);
template <>
__device__ __forceinline__ float cast_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 cast_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half cast_out<__half>(float v) { return __float2half_rn(v); }

template <typename OutT>
__global__ void __launch_bounds__(256) normalize_rows_kernel(const float *__restrict__ x, int64_t rows, int dim,
                                                             int mode, OutT *__restrict__ out, float *__restrict__ neg_inf_fill,
                                                             int *__restrict__ zero_fill)
{
    int lane = threadIdx.x & 31;
    int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp; r < rows; r += nwarps) {
        const float *row = x + r * dim;
        if (neg_inf_fill && lane == 0) neg_inf_fill[r] = -INFINITY;  // per-row scratch resets riding on this launch
        if (zero_fill && lane == 0) zero_fill[r] = 0;
        float denom = 1.f;
        if (mode != FRB_QNORM_NONE) {
            float n = sqrtf(row_sumsq(row, dim, lane));
            denom = (mode == FRB_QNORM_CLAMP) ? fmaxf(n, 1e-12f) : (n + 1e-8f);
        }
        OutT *o = out + r * dim;
        for (int i = lane; i < dim; i += 32) {
            float v = __ldg(row + i);
            o[i] = cast_out<OutT>(mode == FRB_QNORM_NONE ? v : __fdiv_rn(v, denom));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// BGR -> gray, OpenCV's 8-bit fixed point (cv2.cvtColor COLOR_BGR2GRAY, 4.x): (3735 B + 19235 G + 9798 R + 2^14) >> 15.
// HBM-bound: 3 bytes in, 1 byte out per pixel.  A thread converts 16 pixels: three 128-bit loads, one 128-bit store.
__device__ __forceinline__ uint32_t gray_of(uint32_t b, uint32_t g, uint32_t r)
{
    return (b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15;
}

__global__ void __launch_bounds__(256) bgr2gray_kernel(const uint8_t *__restrict__ bgr, int64_t n_px, uint8_t *__restrict__ out, int vec_ok)
{
    const int64_t n_vec = vec_ok ? n_px / 16 : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec; v += stride) {
        const uint4 *src = reinterpret_cast<const uint4 *>(bgr) + v * 3;
        const uint4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
        const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {  // 4 pixels = 12 bytes = 3 words
            const uint32_t w0 = w[3 * q], w1 = w[3 * q + 1], w2 = w[3 * q + 2];
            const uint32_t g0 = gray_of(w0 & 0xFF, (w0 >> 8) & 0xFF, (w0 >> 16) & 0xFF);
            const uint32_t g1 = gray_of(w0 >> 24, w1 & 0xFF, (w1 >> 8) & 0xFF);
            const uint32_t g2 = gray_of((w1 >> 16) & 0xFF, w1 >> 24, w2 & 0xFF);
            const uint32_t g3 = gray_of((w2 >> 8) & 0xFF, (w2 >> 16) & 0xFF, w2 >> 24);
            o[q] = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
        }
        reinterpret_cast<uint4 *>(out)[v] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    for (int64_t i = n_vec * 16 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += stride)
        out[i] = (uint8_t)gray_of(__ldg(bgr + 3 * i), __ldg(bgr + 3 * i + 1), __ldg(bgr + 3 * i + 2));
}

// ---------------------------------------------------------------------------------------------
// Gallery builder: out[g] = mean(rows of group g) / (||mean|| + 1e-8).  One CTA per group; thread t owns dims
// t, t + 128, ...; rows are added in the order given (ascending sample index), in float32, then divided by the
// count, as numpy's mean(axis=0) does; sum of squares by a block reduction.  Empty groups stay all-zero.
constexpr int kGroupThreads = 128;
constexpr int kGroupMaxPerThread = 8;  // dim <= 1024

__global__ void __launch_bounds__(kGroupThreads) group_mean_renorm_kernel(const float *__restrict__ emb, const int64_t *__restrict__ order,
                                                                          const int64_t *__restrict__ offsets, int dim,
                                                                          float *__restrict__ out_f32, __nv_bfloat16 *__restrict__ out_bf16)
{
    __shared__ float s_red[kGroupThreads / 32];
    const int64_t g = blockIdx.x;
    const int64_t lo = offsets[g], hi = offsets[g + 1];
    const int tid = threadIdx.x;
    float acc[kGroupMaxPerThread];
#pragma unroll
    for (int j = 0; j < kGroupMaxPerThread; j++) acc[j] = 0.f;
    for (int64_t i = lo; i < hi; i++) {
        const float *row = emb + order[i] * dim;
#pragma unroll
        for (int j = 0; j < kGroupMaxPerThread; j++) {
            const int d = tid + j * kGroupThreads;
            if (d < dim) acc[j] = __fadd_rn(acc[j], __ldg(row + d));
        }
    }
    const float n = (float)(hi - lo);
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < kGroupMaxPerThread; j++) {
        if (hi > lo) acc[j] = __fdiv_rn(acc[j], n);
        ss = fmaf(acc[j], acc[j], ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kGroupThreads / 32; w++) tot += s_red[w];
    const float denom = sqrtf(tot) + 1e-8f;
#pragma unroll
    for (int j = 0; j < kGroupMaxPerThread; j++) {
        const int d = tid + j * kGroupThreads;
        if (d < dim) {
            const float v = hi > lo ? __fdiv_rn(acc[j], denom) : 0.f;
            if (out_f32) out_f32[g * dim + d] = v;
            if (out_bf16) out_bf16[g * dim + d] = __float2bfloat16_rn(v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// cv2.resize(src, (dst_cols, dst_rows)) with the default INTER_LINEAR on 8-bit images, bit-exact with OpenCV 4.x's
// fixed-point path (imgproc/src/resize.cpp: resizeGeneric_Invoker / HResizeLinear / VResizeLinear<uchar,...>):
//   coordinate  f = (float)((d + 0.5) * scale - 0.5), scale = 1 / (dst / src) in double; s = floor(f); f -= s
//   weights     short(cvRound((1 - f) * 2048)), short(cvRound(f * 2048))   (round half to even)
//   columns     s < 0 -> (s, f) = (0, 0);  s >= src - 1 -> (src - 1, 0)   (weights clamped)
//   rows        weights kept, ROW INDICES clamped to [0, src - 1]          (so a border row is blended with itself)
//   horizontal  r = p[s] * a0 + p[s + 1] * a1                               (int32, 11 fractional bits)
//   vertical    ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
// and, when both axes shrink by exactly 2, OpenCV switches INTER_LINEAR to the INTER_AREA fast path:
// (p00 + p01 + p10 + p11 + 2) >> 2.  Pinned against the real cv2.resize (tests/test_oracle_lbph.py, oracle/resize.py).
// GRAY additionally applies cvtColor(BGR2GRAY) to the resized pixel, so the resized colour image never reaches HBM.
struct ResizeTap {
    int ofs;         // first source row / column (not clamped for rows)
    short w0, w1;    // 11-bit fixed-point weights
};

__device__ __forceinline__ ResizeTap resize_tap(int d, double scale, int src, bool clamp_weights)
{
    // every step individually rounded, as the host code of OpenCV does it (no fused multiply-add)
    float f = __double2float_rn(__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5));
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (clamp_weights) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= src - 1) { s = src - 1; f = 0.f; }
    }
    ResizeTap t;
    t.ofs = s;
    t.w0 = (short)__float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    t.w1 = (short)__float2int_rn(__fmul_rn(f, 2048.f));
    return t;
}

// Work unit = (image, block of `rows_per_unit` destination rows); a CTA walks the unit's pixels 256 at a time, so
// consecutive lanes write consecutive bytes and read neighbouring source pixels (L1 serves the overlap of the taps).
// The tables hold everything that depends on one coordinate only (clamped byte offsets of both taps, weights), the
// loop body is loads + the fixed-point blend; two pixels per thread per trip keep 24 loads in flight.
// No clamp to [0, 255] is needed: weights are >= 0 and each pair sums to <= 2049, so the result is <= 255.
// Byte loads on purpose: fetching each row's two taps as three aligned words + funnel shifts was measured slower.
struct ResizeColTap {
    int x0, x1;      // byte offsets of the two taps within a source row
    short w0, w1;
    int pad;
};
struct ResizeRowTap {
    int y0, y1;      // byte offsets of the two source rows within the image (row index clamped)
    short w0, w1;
    int pad;
};

template <int C, bool GRAY>
__device__ __forceinline__ void resize_pixel(const uint8_t *__restrict__ img, const ResizeColTap ax, const ResizeRowTap ay,
                                             uint8_t *__restrict__ out)
{
    const uint8_t *p00 = img + ay.y0 + ax.x0, *p01 = img + ay.y0 + ax.x1, *p10 = img + ay.y1 + ax.x0, *p11 = img + ay.y1 + ax.x1;
    uint32_t v[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int h0 = (int)__ldg(p00 + c) * ax.w0 + (int)__ldg(p01 + c) * ax.w1;
        const int h1 = (int)__ldg(p10 + c) * ax.w0 + (int)__ldg(p11 + c) * ax.w1;
        v[c] = (uint32_t)((((ay.w0 * (h0 >> 4)) >> 16) + ((ay.w1 * (h1 >> 4)) >> 16) + 2) >> 2);
    }
    if (GRAY) {
        out[0] = (uint8_t)gray_of(v[0], v[C > 1 ? 1 : 0], v[C > 2 ? 2 : 0]);
    } else {
#pragma unroll
        for (int c = 0; c < C; c++) out[c] = (uint8_t)v[c];
    }
}

template <int C, bool GRAY>
__global__ void __launch_bounds__(256) resize_linear_kernel(const uint8_t *__restrict__ src, int64_t count, int src_rows, int src_cols,
                                                            uint8_t *__restrict__ dst, int dst_rows, int dst_cols, double scale_x,
                                                            double scale_y, int area2, int rows_per_unit)
{
    extern __shared__ __align__(16) unsigned char resize_smem[];
    ResizeColTap *tx = reinterpret_cast<ResizeColTap *>(resize_smem);
    ResizeRowTap *ty = reinterpret_cast<ResizeRowTap *>(tx + dst_cols);
    const int row_bytes = src_cols * C;
    if (!area2) {
        for (int d = threadIdx.x; d < dst_cols; d += blockDim.x) {
            const ResizeTap t = resize_tap(d, scale_x, src_cols, true);
            ResizeColTap o;
            o.x0 = t.ofs * C;
            o.x1 = min(t.ofs + 1, src_cols - 1) * C;
            o.w0 = t.w0;
            o.w1 = t.w1;
            o.pad = 0;
            tx[d] = o;
        }
        for (int d = threadIdx.x; d < dst_rows; d += blockDim.x) {
            const ResizeTap t = resize_tap(d, scale_y, src_rows, false);
            ResizeRowTap o;
            o.y0 = min(max(t.ofs, 0), src_rows - 1) * row_bytes;
            o.y1 = min(max(t.ofs + 1, 0), src_rows - 1) * row_bytes;
            o.w0 = t.w0;
            o.w1 = t.w1;
            o.pad = 0;
            ty[d] = o;
        }
    }
    __syncthreads();
    constexpr int OC = GRAY ? 1 : C;
    const int unit_blocks = (dst_rows + rows_per_unit - 1) / rows_per_unit;
    const int64_t units = count * unit_blocks;
    const int64_t src_img = (int64_t)src_rows * row_bytes, dst_img = (int64_t)dst_rows * dst_cols * OC;
    const int nthr = (int)blockDim.x;
    const int step_y = nthr / dst_cols, step_x = nthr - step_y * dst_cols;
    for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
        const int64_t b = u / unit_blocks;
        const int dy0 = (int)(u - b * unit_blocks) * rows_per_unit;
        const int n = (min(dst_rows, dy0 + rows_per_unit) - dy0) * dst_cols;
        const uint8_t *img = src + b * src_img;
        uint8_t *out = dst + b * dst_img + (int64_t)dy0 * dst_cols * OC;
        int dy = dy0 + (int)threadIdx.x / dst_cols, dx = (int)threadIdx.x % dst_cols;
        if (area2) {
            for (int r = threadIdx.x; r < n; r += nthr) {
                const uint8_t *p0 = img + (int64_t)(2 * dy) * row_bytes + 2 * dx * C, *p1 = p0 + row_bytes;
                uint32_t v[C];
#pragma unroll
                for (int c = 0; c < C; c++) v[c] = (__ldg(p0 + c) + __ldg(p0 + C + c) + __ldg(p1 + c) + __ldg(p1 + C + c) + 2u) >> 2;
                if (GRAY) {
                    out[r] = (uint8_t)gray_of(v[0], v[C > 1 ? 1 : 0], v[C > 2 ? 2 : 0]);
                } else {
#pragma unroll
                    for (int c = 0; c < C; c++) out[r * OC + c] = (uint8_t)v[c];
                }
                dx += step_x;
                dy += step_y;
                if (dx >= dst_cols) { dx -= dst_cols; dy++; }
            }
            continue;
        }
        int r = threadIdx.x;
        constexpr int P = 4;                          // pixels per trip (r, r + nthr, ...): 48 byte loads in flight (2: 0.458 ms, 4: 0.418, 8: 0.480)
        for (; r + (P - 1) * nthr < n; r += P * nthr) {
            int px[P], py[P];
#pragma unroll
            for (int u = 0; u < P; u++) {
                px[u] = dx;
                py[u] = dy;
                dx += step_x;
                dy += step_y;
                if (dx >= dst_cols) { dx -= dst_cols; dy++; }
            }
#pragma unroll
            for (int u = 0; u < P; u++) resize_pixel<C, GRAY>(img, tx[px[u]], ty[py[u]], out + (int64_t)(r + u * nthr) * OC);
        }
        for (; r < n; r += nthr) {
            resize_pixel<C, GRAY>(img, tx[dx], ty[dy], out + (int64_t)r * OC);
            dx += step_x;
            dy += step_y;
            if (dx >= dst_cols) { dx -= dst_cols; dy++; }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// One thread per query merges n_lists sorted candidate lists of length k (k <= FRB_MAX_K).
template <bool LARGEST>
__global__ void __launch_bounds__(128) topk_merge_kernel(const float *__restrict__ cs, const int64_t *__restrict__ ci,
                                                         int64_t s_stride, int64_t i_stride, int n_lists, int64_t n_query,
                                                         int k, float *__restrict__ os, int64_t *__restrict__ oi)
{
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_query) return;
    float s[FRB_MAX_K];
    int64_t id[FRB_MAX_K];
    list_init<LARGEST>(s, id, k);
    for (int l = 0; l < n_lists; l++) {
        const float *ls = cs + (int64_t)l * s_stride + q * k;     // strides in elements between consecutive lists
        const int64_t *li = ci + (int64_t)l * i_stride + q * k;
        for (int j = 0; j < k; j++) {
            float v = ls[j];
            int64_t idx = li[j];
            if (idx < 0) continue;
            if (!better<LARGEST>(v, idx, s[k - 1], id[k - 1])) continue;
            int p = k - 1;
            while (p > 0 && better<LARGEST>(v, idx, s[p - 1], id[p - 1])) {
                s[p] = s[p - 1];
                id[p] = id[p - 1];
                --p;
            }
            s[p] = v;
            id[p] = idx;
        }
    }
    for (int j = 0; j < k; j++) {
        os[q * k + j] = s[j];
        oi[q * k + j] = id[j];
    }
}

int normalize_rows_impl(const float *x, int64_t rows, int dim, int mode, void *out, int out_dtype, float *neg_inf_fill,
                        int *zero_fill, cudaStream_t st)
{
    FRB_CHECK_ARG(rows >= 0 && dim > 0, "frb_normalize_rows: rows=%lld dim=%d", (long long)rows, dim);
    FRB_CHECK_ARG(mode >= FRB_QNORM_NONE && mode <= FRB_QNORM_EPS, "frb_normalize_rows: mode=%d", mode);
    FRB_CHECK_ARG(out_dtype == FRB_F32 || out_dtype == FRB_BF16 || out_dtype == FRB_F16, "frb_normalize_rows: out_dtype=%d", out_dtype);
    if (rows == 0) return FRB_OK;
    FRB_CHECK_ARG(x && out, "frb_normalize_rows: null pointer");
    int64_t blocks = (rows + 7) / 8;
    int grid = (int)(blocks < (int64_t)sm_count() * 8 ? blocks : (int64_t)sm_count() * 8);
    if (grid < 1) grid = 1;
    if (out_dtype == FRB_F32)
        normalize_rows_kernel<float><<<grid, 256, 0, st>>>(x, rows, dim, mode, (float *)out, neg_inf_fill, zero_fill);
    else if (out_dtype == FRB_F16)
        normalize_rows_kernel<__half><<<grid, 256, 0, st>>>(x, rows, dim, mode, (__half *)out, neg_inf_fill, zero_fill);
    else
        normalize_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x, rows, dim, mode, (__nv_bfloat16 *)out, neg_inf_fill, zero_fill);
    FRB_LAUNCH_OK("normalize_rows_kernel");
    return FRB_OK;
}

// Compact candidate merge: query q owns cnt[q] unordered (score, row) candidates at cand[q * cap ...].
// Lanes take candidates round-robin into private sorted lists (loads issued four at a time so their latencies
// overlap), then k rounds of a warp-wide arg-best over the lane heads.  Total order (key, then lowest row) =>
// order-independent result.  WQ = 1: one warp per query, 8 queries per CTA (large batches).  WQ = 8: one CTA per
// query, each warp reduces an eighth of the candidates into shared memory and warp 0 merges the eight lists --
// for a handful of queries with thousands of candidates each (the row-streaming kernel leaves k per CTA: 2895 for
// one query, which took one warp 119 us of dependent L2 round trips).
template <bool LARGEST>
__device__ __forceinline__ void merge_insert(float *s, int64_t *id, int k, float v, int64_t idx)
{
    if (idx < 0 || !better<LARGEST>(v, idx, s[k - 1], id[k - 1])) return;
    int p = k - 1;
    while (p > 0 && better<LARGEST>(v, idx, s[p - 1], id[p - 1])) {
        s[p] = s[p - 1];
        id[p] = id[p - 1];
        --p;
    }
    s[p] = v;
    id[p] = idx;
}

// candidate i of a query: compact layout (contiguous) or n_lists sorted lists of k, `stride` elements apart
struct CompactCands {
    const float *s;
    const int64_t *x;
    __device__ __forceinline__ void get(int64_t i, float &v, int64_t &idx) const { v = s[i]; idx = x[i]; }
};
struct ListCands {
    const float *s;
    const int64_t *x;
    int64_t s_stride, i_stride;
    int k;
    __device__ __forceinline__ void get(int64_t i, float &v, int64_t &idx) const
    {
        const int64_t l = i / k, j = i - l * k;
        v = s[l * s_stride + j];
        idx = x[l * i_stride + j];
    }
};

template <bool LARGEST, class Cands>
__device__ __forceinline__ void merge_scan(float *s, int64_t *id, int k, const Cands &c, int64_t first, int64_t n, int64_t step)
{
    int64_t i = first;
    for (; i + 3 * step < n; i += 4 * step) {
        float v[4];
        int64_t x[4];
#pragma unroll
        for (int u = 0; u < 4; u++) c.get(i + u * step, v[u], x[u]);
#pragma unroll
        for (int u = 0; u < 4; u++) merge_insert<LARGEST>(s, id, k, v[u], x[u]);
    }
    for (; i < n; i += step) {
        float v;
        int64_t x;
        c.get(i, v, x);
        merge_insert<LARGEST>(s, id, k, v, x);
    }
}

// k rounds of warp arg-best over the lanes' sorted lists; lane 0 stores round r through (os, oi)
template <bool LARGEST>
__device__ __forceinline__ void merge_extract(const float *s, const int64_t *id, int kl, int k, int lane, float *os, int64_t *oi)
{
    int head = 0;
    for (int r = 0; r < k; r++) {
        float v = head < kl ? s[head] : worst_value<LARGEST>();
        int64_t idx = head < kl ? id[head] : -1;
        const int64_t mine = idx;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int64_t oidx = __shfl_xor_sync(0xffffffffu, idx, o);
            if (better<LARGEST>(ov, oidx, v, idx)) { v = ov; idx = oidx; }
        }
        if (idx >= 0 && mine == idx) head++;
        if (lane == 0) {
            os[r] = idx >= 0 ? v : worst_value<LARGEST>();
            oi[r] = idx;
        }
    }
}

// shared by the two layouts: this warp's (WQ = 1) or this CTA's (WQ = 8) query q with n candidates
template <bool LARGEST, int WQ, class Cands>
__device__ __forceinline__ void merge_query(const Cands &c, int64_t n, int k, float *os, int64_t *oi)
{
    __shared__ float sh_s[WQ > 1 ? WQ * FRB_MAX_K : 1];
    __shared__ int64_t sh_i[WQ > 1 ? WQ * FRB_MAX_K : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float s[FRB_MAX_K];
    int64_t id[FRB_MAX_K];
    // a lane sees at most ceil(n / stride) candidates: its private list need not be longer (k = 64 with ~26 candidates
    // per lane spent most of its time shifting empty slots in local memory)
    const int stride = 32 * WQ;
    int kl = (int)((n + stride - 1) / stride);
    kl = kl < 1 ? 1 : (kl > k ? k : kl);
    list_init<LARGEST>(s, id, kl);
    if (WQ == 1) {
        merge_scan<LARGEST>(s, id, kl, c, lane, n, 32);
        merge_extract<LARGEST>(s, id, kl, k, lane, os, oi);
    } else {
        merge_scan<LARGEST>(s, id, kl, c, threadIdx.x, n, stride);
        merge_extract<LARGEST>(s, id, kl, k, lane, sh_s + warp * k, sh_i + warp * k);
        __syncthreads();
        if (warp == 0) {
            int k2 = (WQ * k + 31) / 32;
            k2 = k2 > k ? k : k2;
            list_init<LARGEST>(s, id, k2);
            const CompactCands mine = {sh_s, sh_i};
            merge_scan<LARGEST>(s, id, k2, mine, lane, (int64_t)WQ * k, 32);
            merge_extract<LARGEST>(s, id, k2, k, lane, os, oi);
        }
    }
}

template <bool LARGEST, int WQ>
__global__ void __launch_bounds__(256) topk_merge_compact_kernel(const float *__restrict__ cs, const int64_t *__restrict__ ci,
                                                                 const int *__restrict__ cnt, int64_t cap, int64_t n_query,
                                                                 int k, float *__restrict__ os, int64_t *__restrict__ oi)
{
    const int64_t q = WQ > 1 ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= n_query) return;   // WQ > 1: uniform over the CTA
    int64_t n = cnt[q];
    if (n > cap) n = cap;
    const CompactCands c = {cs + q * cap, ci + q * cap};
    merge_query<LARGEST, WQ>(c, n, k, os + q * k, oi + q * k);
}

// the strided-list layout ([n_lists] x [n_query, k], lists s_stride / i_stride elements apart) with many lists:
// same warp / CTA per query scheme instead of topk_merge_kernel's one thread per query
template <bool LARGEST, int WQ>
__global__ void __launch_bounds__(256) topk_merge_lists_kernel(const float *__restrict__ cs, const int64_t *__restrict__ ci,
                                                               int64_t s_stride, int64_t i_stride, int n_lists, int64_t n_query,
                                                               int k, float *__restrict__ os, int64_t *__restrict__ oi)
{
    const int64_t q = WQ > 1 ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= n_query) return;
    const ListCands c = {cs + q * k, ci + q * k, s_stride, i_stride, k};
    merge_query<LARGEST, WQ>(c, (int64_t)n_lists * k, k, os + q * k, oi + q * k);
}

// u16 counts -> u8 counts (the caller guarantees every count <= 255), 8 per thread
__global__ void __launch_bounds__(256) counts_narrow_kernel(const uint16_t *__restrict__ src, int64_t n, uint8_t *__restrict__ dst)
{
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 8;
    if (i >= n) return;
    if (i + 8 <= n && ((reinterpret_cast<uintptr_t>(src + i) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst + i) & 7) == 0)) {
        const uint4 v = *reinterpret_cast<const uint4 *>(src + i);
        *reinterpret_cast<uint2 *>(dst + i) = make_uint2(__byte_perm(v.x, v.y, 0x6420), __byte_perm(v.z, v.w, 0x6420));
    } else {
        for (int64_t j = i; j < n && j < i + 8; j++) dst[j] = (uint8_t)src[j];
    }
}

__global__ void __launch_bounds__(256) index_remap_kernel(int64_t *__restrict__ idx, int64_t n, const int64_t *__restrict__ table,
                                                          int64_t table_len)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int64_t v = idx[i];
    if (v >= 0 && v < table_len) idx[i] = table[v];
}

int topk_merge_compact(const float *cs, const int64_t *ci, const int *cnt, int64_t cap, int64_t n_query, int k, int largest,
                       float *os, int64_t *oi, cudaStream_t st)
{
    if (n_query == 0) return FRB_OK;
    // a CTA per query while that still leaves SMs idle and the lists are long enough to split eight ways
    if (n_query <= sm_count() && cap >= 512) {
        if (largest)
            topk_merge_compact_kernel<true, 8><<<(int)n_query, 256, 0, st>>>(cs, ci, cnt, cap, n_query, k, os, oi);
        else
            topk_merge_compact_kernel<false, 8><<<(int)n_query, 256, 0, st>>>(cs, ci, cnt, cap, n_query, k, os, oi);
    } else {
        const int grid = (int)((n_query + 7) / 8);
        if (largest)
            topk_merge_compact_kernel<true, 1><<<grid, 256, 0, st>>>(cs, ci, cnt, cap, n_query, k, os, oi);
        else
            topk_merge_compact_kernel<false, 1><<<grid, 256, 0, st>>>(cs, ci, cnt, cap, n_query, k, os, oi);
    }
    FRB_LAUNCH_OK("topk_merge_compact_kernel");
    return FRB_OK;
}

}  // namespace frb

using namespace frb;

extern "C" {

int frb_version(void) { return 100; }

int frb_profile_enable(int on)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
    return FRB_OK;
}

int frb_profile_read(int kernel, float *total_ms, int *launches)
{
    FRB_CHECK_ARG(kernel >= 0 && kernel < FRB_K_COUNT, "frb_profile_read: kernel=%d", kernel);
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        ev.swap(g_prof[kernel]);
    }
    float total = 0.f;
    for (auto &e : ev) {
        float ms = 0.f;
        FRB_CUDA_OK(cudaEventSynchronize(e.second));
        FRB_CUDA_OK(cudaEventElapsedTime(&ms, e.first, e.second));
        total += ms;
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    if (total_ms) *total_ms = total;
    if (launches) *launches = (int)ev.size();
    return FRB_OK;
}

const char *frb_last_error(void) { return last_error_buf(); }

int frb_device_info(int *sm, int *major, int *minor)
{
    int dev = 0;
    FRB_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    FRB_CUDA_OK(cudaGetDeviceProperties(&p, dev));
    if (sm) *sm = p.multiProcessorCount;
    if (major) *major = p.major;
    if (minor) *minor = p.minor;
    return FRB_OK;
}

int frb_row_norms_f32(const float *x, int64_t rows, int dim, float *out, void *stream)
{
    FRB_CHECK_ARG(rows >= 0 && dim > 0, "frb_row_norms_f32: rows=%lld dim=%d", (long long)rows, dim);
    if (rows == 0) return FRB_OK;
    FRB_CHECK_ARG(x && out, "frb_row_norms_f32: null pointer");
    int64_t blocks = (rows + 7) / 8;
    int grid = (int)(blocks < (int64_t)sm_count() * 8 ? blocks : (int64_t)sm_count() * 8);
    if (grid < 1) grid = 1;
    row_norms_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, rows, dim, out);
    FRB_LAUNCH_OK("row_norms_kernel");
    return FRB_OK;
}

int frb_normalize_rows(const float *x, int64_t rows, int dim, int mode, void *out, int out_dtype, void *stream)
{
    return normalize_rows_impl(x, rows, dim, mode, out, out_dtype, nullptr, nullptr, (cudaStream_t)stream);
}

static int topk_merge_launch(const char *fn, const float *cs, const int64_t *ci, int64_t s_stride, int64_t i_stride, int n_lists,
                             int64_t n_query, int k, int largest, float *os, int64_t *oi, void *stream)
{
    FRB_CHECK_ARG(n_lists >= 1 && n_query >= 0 && k >= 1 && k <= FRB_MAX_K, "%s: n_lists=%d n_query=%lld k=%d", fn, n_lists,
                  (long long)n_query, k);
    if (n_query == 0) return FRB_OK;
    FRB_CHECK_ARG(cs && ci && os && oi, "%s: null pointer", fn);
    FRB_CHECK_ARG(s_stride >= n_query * k && i_stride >= n_query * k, "%s: list strides %lld / %lld < n_query * k", fn,
                  (long long)s_stride, (long long)i_stride);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t cands = (int64_t)n_lists * k;
    if (cands >= 512 && n_query <= sm_count()) {
        // few queries, long candidate sets: a CTA per query
        if (largest)
            topk_merge_lists_kernel<true, 8><<<(int)n_query, 256, 0, st>>>(cs, ci, s_stride, i_stride, n_lists, n_query, k, os, oi);
        else
            topk_merge_lists_kernel<false, 8><<<(int)n_query, 256, 0, st>>>(cs, ci, s_stride, i_stride, n_lists, n_query, k, os, oi);
    } else if (cands >= 32 && n_query <= (int64_t)sm_count() * 64) {
        // a warp per query while one thread per query would leave most of the GPU idle walking its lists
        const int grid = (int)((n_query + 7) / 8);
        if (largest)
            topk_merge_lists_kernel<true, 1><<<grid, 256, 0, st>>>(cs, ci, s_stride, i_stride, n_lists, n_query, k, os, oi);
        else
            topk_merge_lists_kernel<false, 1><<<grid, 256, 0, st>>>(cs, ci, s_stride, i_stride, n_lists, n_query, k, os, oi);
    } else {
        const int grid = (int)((n_query + 127) / 128);
        if (largest)
            topk_merge_kernel<true><<<grid, 128, 0, st>>>(cs, ci, s_stride, i_stride, n_lists, n_query, k, os, oi);
        else
            topk_merge_kernel<false><<<grid, 128, 0, st>>>(cs, ci, s_stride, i_stride, n_lists, n_query, k, os, oi);
    }
    FRB_LAUNCH_OK("topk_merge_kernel");
    return FRB_OK;
}

int frb_counts_u16_to_u8(const uint16_t *src, int64_t n, uint8_t *dst, void *stream)
{
    FRB_CHECK_ARG(n >= 0, "frb_counts_u16_to_u8: n=%lld", (long long)n);
    if (n == 0) return FRB_OK;
    FRB_CHECK_ARG(src && dst, "frb_counts_u16_to_u8: null pointer");
    const int64_t threads = (n + 7) / 8;
    counts_narrow_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, n, dst);
    FRB_LAUNCH_OK("counts_narrow_kernel");
    return FRB_OK;
}

int frb_index_remap(int64_t *idx, int64_t n, const int64_t *table, int64_t table_len, void *stream)
{
    FRB_CHECK_ARG(n >= 0 && table_len >= 0, "frb_index_remap: n=%lld table_len=%lld", (long long)n, (long long)table_len);
    if (n == 0) return FRB_OK;
    FRB_CHECK_ARG(idx && (table || table_len == 0), "frb_index_remap: null pointer");
    index_remap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(idx, n, table, table_len);
    FRB_LAUNCH_OK("index_remap_kernel");
    return FRB_OK;
}

int frb_bgr2gray_u8(const uint8_t *bgr, int64_t n_pixels, uint8_t *out, void *stream)
{
    FRB_CHECK_ARG(n_pixels >= 0, "frb_bgr2gray_u8: n_pixels=%lld", (long long)n_pixels);
    if (n_pixels == 0) return FRB_OK;
    FRB_CHECK_ARG(bgr && out, "frb_bgr2gray_u8: null pointer");
    const int vec_ok = (((uintptr_t)bgr | (uintptr_t)out) & 15) == 0;
    int64_t blocks = (n_pixels / 16 + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    ProfileScope prof(FRB_K_BGR2GRAY, (cudaStream_t)stream);
    bgr2gray_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(bgr, n_pixels, out, vec_ok);
    FRB_LAUNCH_OK("bgr2gray_kernel");
    return FRB_OK;
}

int frb_resize_linear_u8(const uint8_t *src, int64_t count, int src_rows, int src_cols, int channels, uint8_t *dst,
                         int dst_rows, int dst_cols, int to_gray, void *stream)
{
    FRB_CHECK_ARG(count >= 0 && src_rows >= 1 && src_cols >= 1 && dst_rows >= 1 && dst_cols >= 1,
                  "frb_resize_linear_u8: count=%lld src=%dx%d dst=%dx%d", (long long)count, src_rows, src_cols, dst_rows, dst_cols);
    FRB_CHECK_ARG(channels == 1 || channels == 3, "frb_resize_linear_u8: channels=%d (1 or 3)", channels);
    FRB_CHECK_ARG(!to_gray || channels == 3, "frb_resize_linear_u8: to_gray needs 3 (BGR) channels");
    if (dst_rows > 4096 || dst_cols > 4096) {   // the two tap tables (16 B per destination row / column) live in shared memory
        set_error("frb_resize_linear_u8: destination %dx%d larger than 4096 per side", dst_rows, dst_cols);
        return FRB_ERR_UNSUPPORTED;
    }
    if (count == 0) return FRB_OK;
    FRB_CHECK_ARG(src && dst, "frb_resize_linear_u8: null pointer");
    // cv::resize: inv_scale = dsize / ssize; hal::resize: scale = 1. / inv_scale (both double)
    const double scale_x = 1.0 / ((double)dst_cols / (double)src_cols), scale_y = 1.0 / ((double)dst_rows / (double)src_rows);
    const int area2 = src_cols == 2 * dst_cols && src_rows == 2 * dst_rows;
    FRB_CHECK_ARG((int64_t)src_rows * src_cols * channels <= INT32_MAX, "frb_resize_linear_u8: source image of %dx%dx%d bytes",
                  src_rows, src_cols, channels);
    // a unit = one image x a block of destination rows: >= ~1024 pixels each, and ~16 units per SM when the batch allows
    const int sms = sm_count();
    int64_t rpu = (count * dst_rows + (int64_t)sms * 16 - 1) / ((int64_t)sms * 16);
    const int64_t min_rows = (1024 + dst_cols - 1) / dst_cols;
    if (rpu < min_rows) rpu = min_rows;
    if (rpu > dst_rows) rpu = dst_rows;
    const int64_t units = count * ((dst_rows + rpu - 1) / rpu);
    const size_t smem = (size_t)dst_cols * sizeof(ResizeColTap) + (size_t)dst_rows * sizeof(ResizeRowTap);
    cudaStream_t st = (cudaStream_t)stream;
    typedef void (*Kernel)(const uint8_t *, int64_t, int, int, uint8_t *, int, int, double, double, int, int);
    const Kernel kernel = channels == 1 ? resize_linear_kernel<1, false>
                                        : (to_gray ? resize_linear_kernel<3, true> : resize_linear_kernel<3, false>);
    if (smem > 48 * 1024) FRB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // exactly one resident wave: CTAs stride over the units, and a partial second wave would run at a fraction of
    // the occupancy while the first one has already finished
    int per_sm = 0;
    FRB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t blocks = units;
    if (blocks > (int64_t)sms * per_sm) blocks = (int64_t)sms * per_sm;
    ProfileScope prof(FRB_K_RESIZE, st);
    kernel<<<(unsigned)blocks, 256, smem, st>>>(src, count, src_rows, src_cols, dst, dst_rows, dst_cols, scale_x, scale_y, area2, (int)rpu);
    FRB_LAUNCH_OK("resize_linear_kernel");
    return FRB_OK;
}

int frb_group_mean_renorm(const float *emb, const int64_t *order, const int64_t *offsets, int64_t n_groups, int dim,
                          float *out_f32, void *out_bf16, void *stream)
{
    FRB_CHECK_ARG(n_groups >= 0 && dim > 0 && dim <= kGroupThreads * kGroupMaxPerThread, "frb_group_mean_renorm: n_groups=%lld dim=%d (dim <= %d)",
                  (long long)n_groups, dim, kGroupThreads * kGroupMaxPerThread);
    if (n_groups == 0) return FRB_OK;
    FRB_CHECK_ARG(emb && order && offsets && (out_f32 || out_bf16), "frb_group_mean_renorm: null pointer");
    FRB_CHECK_ARG(n_groups <= 2147483647LL, "frb_group_mean_renorm: too many groups");
    group_mean_renorm_kernel<<<(unsigned)n_groups, kGroupThreads, 0, (cudaStream_t)stream>>>(emb, order, offsets, dim, out_f32,
                                                                                            (__nv_bfloat16 *)out_bf16);
    FRB_LAUNCH_OK("group_mean_renorm_kernel");
    return FRB_OK;
}

int frb_topk_merge(const float *cs, const int64_t *ci, int n_lists, int64_t n_query, int k, int largest,
                   float *os, int64_t *oi, void *stream)
{
    return topk_merge_launch("frb_topk_merge", cs, ci, n_query * k, n_query * k, n_lists, n_query, k, largest, os, oi, stream);
}

int frb_topk_merge_strided(const float *cs, const int64_t *ci, int64_t score_list_stride, int64_t idx_list_stride, int n_lists,
                           int64_t n_query, int k, int largest, float *os, int64_t *oi, void *stream)
{
    return topk_merge_launch("frb_topk_merge_strided", cs, ci, score_list_stride, idx_list_stride, n_lists, n_query, k, largest,
                             os, oi, stream);
}

}  // extern "C"
