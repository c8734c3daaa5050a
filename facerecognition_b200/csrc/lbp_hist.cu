// lbp_hist.cu — K2: LBP codes (OpenCV elbp_, radius 1 / 8 neighbours) + grid histograms.
//
// Replaces the per-image body of cv2.face.LBPHFaceRecognizer train()/predict()
// (reference call sites models/lbphmodel/train_lbph.py:35, models/lbphmodel/inference_lbph.py:5):
// opencv_contrib lbph_faces.cpp elbp_<uchar> + spatial_histogram.
//
// Exactness.  elbp_ samples 8 points on the unit circle with float32 bilinear weights
//     A = 0x3e5413cd (0.20710678)  B = 0x3effffff (0.49999997)  C = 0x3dafb0ce (0.08578645)
// and sets bit n when (t > c) || (|t - c| < FLT_EPSILON), t = ((w1*p1 + w2*p2) + w3*p3) + w4*p4 with
// every product and sum rounded to float32 (no FMA).  Two identities, both checked exhaustively in
// tests/test_oracle_lbph.py, let the kernel do this with one compare per bit:
//   * axis bits (n = 0,2,4,6): the weights degenerate to (1, ~6e-17, 0, 0), so bit = neighbour >= centre;
//   * diagonal bits: for an integer centre c in 0..255, (t > c) || |t - c| < 2^-23  <=>  t >= thr(c) with
//     thr(1) = 1 - 2^-24 and thr(c) = c otherwise (float spacing at c >= 2 is >= 2^-23).
// The blend itself keeps OpenCV's left-to-right order with every product and sum rounded separately: scalar
// __fmul_rn / __fadd_rn in the codes-only kernel; in the histogram kernels packed __fadd2_rn sums and products
// formed as fma(a, b, +0) — ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (seen in SASS, also with
// -fmad=false, which this file is compiled with anyway), which would round once where OpenCV rounds twice.
//
// Kernels.  lbp_codes_kernel: codes only (parity checks).  lbp_hist_pipe_kernel: the fast path for images the
// TMA can fetch — double-buffered images and counters, coding warps + one writer warp, mbarrier hand-offs, no
// block-wide barrier.  lbp_hist_kernel: any other shape — same arithmetic, plain staging, two barriers per
// image.  See the comment blocks above each for the work decomposition.
#include "frb_common.cuh"

namespace frb {

__device__ __forceinline__ float u8_to_f32(unsigned v)
{
    // exact for 0..255: place the byte in the mantissa of 2^23 and subtract 2^23 (LOP3 + FADD, no I2F)
    return __uint_as_float(0x4B000000u | v) - 8388608.0f;
}

// 3x3 window, row-major a b c / d e f / g h i (e = centre); returns the 8-bit LBP code.
__device__ __forceinline__ unsigned lbp_code_r1p8(float a, float b, float c, float d, float e, float f, float g,
                                                  float h, float i)
{
    const float A = __uint_as_float(0x3e5413cdu);
    const float B = __uint_as_float(0x3effffffu);
    const float C = __uint_as_float(0x3dafb0ceu);
    const float thr = (e == 1.0f) ? __uint_as_float(0x3f7fffffu) : e;  // 1 - 2^-24 when centre == 1
    const float Ab = __fmul_rn(A, b), Ad = __fmul_rn(A, d), Af = __fmul_rn(A, f), Ah = __fmul_rn(A, h);
    const float Ce = __fmul_rn(C, e);
    // n=1 (NE): A*b + B*c + C*e + A*f        n=3 (NW): B*a + A*b + A*d + C*e
    // n=5 (SW): A*d + C*e + B*g + A*h        n=7 (SE): C*e + A*f + A*h + B*i
    const float t1 = __fadd_rn(__fadd_rn(__fadd_rn(Ab, __fmul_rn(B, c)), Ce), Af);
    const float t3 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(B, a), Ab), Ad), Ce);
    const float t5 = __fadd_rn(__fadd_rn(__fadd_rn(Ad, Ce), __fmul_rn(B, g)), Ah);
    const float t7 = __fadd_rn(__fadd_rn(__fadd_rn(Ce, Af), Ah), __fmul_rn(B, i));
    unsigned code = (f >= e) ? 1u : 0u;
    code |= (t1 >= thr) ? 2u : 0u;
    code |= (b >= e) ? 4u : 0u;
    code |= (t3 >= thr) ? 8u : 0u;
    code |= (d >= e) ? 16u : 0u;
    code |= (t5 >= thr) ? 32u : 0u;
    code |= (h >= e) ? 64u : 0u;
    code |= (t7 >= thr) ? 128u : 0u;
    return code;
}

// ---- codes only (parity checks of the code stage) ----------------------------------------------
__global__ void __launch_bounds__(256) lbp_codes_kernel(const uint8_t *__restrict__ img, int64_t count, int rows,
                                                        int cols, uint8_t *__restrict__ out)
{
    const int orows = rows - 2, ocols = cols - 2;
    const int64_t per = (int64_t)orows * ocols;
    const int64_t total = count * per;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t b = t / per;
        int r = (int)(t - b * per);
        int y = r / ocols, x = r - y * ocols;
        const uint8_t *p = img + b * rows * cols + (int64_t)y * cols + x;
        float w[9];
#pragma unroll
        for (int dy = 0; dy < 3; dy++)
#pragma unroll
            for (int dx = 0; dx < 3; dx++) w[dy * 3 + dx] = u8_to_f32(__ldg(p + dy * cols + dx));
        out[t] = (uint8_t)lbp_code_r1p8(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], w[8]);
    }
}

// ---- codes + grid histograms ----------------------------------------------------------------------
// The kernel is issue-bound (the exact float32 blend is ~17 flops per code, plus 8 compares), so the design
// goal is the fewest issue slots per code:
//   * every float op is a packed f32x2 instruction (FMUL2 / FADD2, sm_100a): a thread walks the SAME pair of
//     code columns down TWO cell rows ("bands") half a grid apart, lane .x = upper band, lane .y = lower
//     band, so both lanes always execute the same instruction stream and no operand ever needs re-pairing;
//   * a source row is loaded once (two 16-bit shared loads per lane), converted once, and the products it
//     contributes to its neighbours' blends are formed once (A*v, B*v; C*v for the centres); three row slots
//     rotate through an unrolled-by-3 loop so nothing is copied;
//   * compares are subtractions whose SIGN is the (inverted) bit: axis bits  nb - c, diagonal bits
//     t - thr with thr = c - 2^-24 rounded to float32 (== c for c >= 2, 1 - 2^-24 for c = 1, < 0 for c = 0:
//     exactly the thr(c) of the header comment); a funnel shift per bit collects the signs;
//   * a u32 counter word holds bin `code` of BOTH lanes' cells (low half = upper band, high half = lower band), so
//     lane .x always adds 1 and lane .y always adds 0x10000 at word `code` of the cell pair: no value select, no
//     predicate (a column past the grid goes to a scratch pair).  Pairs are 260 words apart, so the popular codes
//     (0, 255, ...) of neighbouring cells fall into banks 4 apart; the write-out splits the halves and stores
//     128 bits per cell.
constexpr unsigned kLbpPairVec = 65;  // uint4 groups per cell pair: 256 counters + 16 bytes of bank skew
constexpr int kLbpMaxThreads = 224;  // 7 warps: 8x8 grid on 112x112 = 208 work items

__device__ __forceinline__ uint32_t lbp_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lbp_mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LBP_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra LBP_DONE;\n\t"
        "bra LBP_WAIT;\n\t"
        "LBP_DONE:\n\t"
        "}" ::"r"(lbp_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// one elected thread: arm the barrier with the byte count and start a 1-D bulk (TMA) copy global -> shared
__device__ __forceinline__ void lbp_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lbp_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(lbp_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(lbp_smem_u32(bar))
                 : "memory");
}

struct LbpRow2 {
    float2 v[4];  // raw pixels of source columns x0..x0+3 (lane .x / .y = the two bands)
    float2 b[4];  // B * v
    float2 a[2];  // A * v for the two centre columns x0+1, x0+2
};

__device__ __forceinline__ float2 f2(float x) { return make_float2(x, x); }

// Rounded product that ptxas cannot contract into a following add.  ptxas 12.9 fuses mul.rn.f32x2 + add.rn.f32x2
// into FFMA2 despite the explicit .rn, also with -fmad=false (seen in SASS); that would round once where OpenCV
// rounds twice.  fma(a, b, +0) rounds a*b exactly once, equals the product for every non-negative product
// (all of ours; it differs from mul only in turning -0 into +0, which is why no optimiser may rewrite it as a
// mul), and an FMA result cannot be fused a second time.
__device__ __forceinline__ float2 mul2_exact(float2 a, float2 b) { return __ffma2_rn(a, b, f2(0.0f)); }

template <bool ALIGNED16>
__device__ __forceinline__ void lbp_load_row2(LbpRow2 &r, const uint8_t *px, const uint8_t *py)
{
    const float2 A = f2(__uint_as_float(0x3e5413cdu)), B = f2(__uint_as_float(0x3effffffu));
    unsigned mx[4], my[4];  // 0x4B0000pp: the byte in the mantissa of 2^23
    if (ALIGNED16) {
        const unsigned lx = *reinterpret_cast<const uint16_t *>(px), hx = *reinterpret_cast<const uint16_t *>(px + 2);
        const unsigned ly = *reinterpret_cast<const uint16_t *>(py), hy = *reinterpret_cast<const uint16_t *>(py + 2);
        mx[0] = __byte_perm(lx, 0x4B000000u, 0x7540); mx[1] = __byte_perm(lx, 0x4B000000u, 0x7541);
        mx[2] = __byte_perm(hx, 0x4B000000u, 0x7540); mx[3] = __byte_perm(hx, 0x4B000000u, 0x7541);
        my[0] = __byte_perm(ly, 0x4B000000u, 0x7540); my[1] = __byte_perm(ly, 0x4B000000u, 0x7541);
        my[2] = __byte_perm(hy, 0x4B000000u, 0x7540); my[3] = __byte_perm(hy, 0x4B000000u, 0x7541);
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            mx[i] = 0x4B000000u | px[i];
            my[i] = 0x4B000000u | py[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        r.v[i] = __fadd2_rn(make_float2(__uint_as_float(mx[i]), __uint_as_float(my[i])), f2(-8388608.0f));
        r.b[i] = mul2_exact(B, r.v[i]);
    }
    r.a[0] = mul2_exact(A, r.v[1]);
    r.a[1] = mul2_exact(A, r.v[2]);
}

// Adds the codes of the two centre columns (both lanes) of `mid` to their cell pairs.  Same blends, same order
// as lbp_code_r1p8.  pair_last[q]: shared-space byte address of bin 255 of column q's cell pair.
struct LbpCells {
    unsigned pair_last[2];
};

__device__ __forceinline__ unsigned sign_in(unsigned acc, float r)
{
    return __funnelshift_l(__float_as_uint(r), acc, 1);  // (acc << 1) | sign(r)
}

// not_code < 256 (exactly eight sign bits were shifted into a zero), so bin = 255 - not_code and the counter
// address is one multiply-add off the pair's LAST word.  volatile keeps the reduction ordered against the block
// barriers; no memory clobber, so the next row's shared loads may be scheduled above it.
__device__ __forceinline__ void lbp_hist_add(unsigned pair_last, unsigned not_code, unsigned one)
{
    const unsigned addr = pair_last - (not_code << 2);
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(one));
}

__device__ __forceinline__ void lbp_emit_row2(const LbpRow2 &top, const LbpRow2 &mid, const LbpRow2 &bot, const LbpCells &cells)
{
    const float2 A = f2(__uint_as_float(0x3e5413cdu)), C = f2(__uint_as_float(0x3dafb0ceu));
    const float2 a0 = mul2_exact(A, mid.v[0]), a3 = mul2_exact(A, mid.v[3]);
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const float2 e = mid.v[q + 1], m1 = f2(-1.0f);
        // -(thr(centre)) = 2^-24 - e; the products by -1 below are exact, so these FMAs round once like the FADDs they replace
        const float2 nthr = __ffma2_rn(e, m1, f2(5.9604644775390625e-08f));
        const float2 Ce = mul2_exact(C, mid.v[q + 1]);
        const float2 aL = q == 0 ? a0 : mid.a[0];          // A * mid[q]
        const float2 aR = q == 0 ? mid.a[1] : a3;          // A * mid[q + 2]
        // n=1 (NE): A*N + B*NE + C*e + A*E        n=3 (NW): B*NW + A*N + A*W + C*e
        // n=5 (SW): A*W + C*e + B*SW + A*S        n=7 (SE): C*e + A*E + A*S + B*SE
        const float2 t1 = __fadd2_rn(__fadd2_rn(__fadd2_rn(top.a[q], top.b[q + 2]), Ce), aR);
        const float2 t3 = __fadd2_rn(__fadd2_rn(__fadd2_rn(top.b[q], top.a[q]), aL), Ce);
        const float2 t5 = __fadd2_rn(__fadd2_rn(__fadd2_rn(aL, Ce), bot.b[q]), bot.a[q]);
        const float2 t7 = __fadd2_rn(__fadd2_rn(__fadd2_rn(Ce, aR), bot.a[q]), bot.b[q + 2]);
        const float2 r7 = __fadd2_rn(t7, nthr), r6 = __ffma2_rn(e, m1, bot.v[q + 1]);
        const float2 r5 = __fadd2_rn(t5, nthr), r4 = __ffma2_rn(e, m1, mid.v[q]);
        const float2 r3 = __fadd2_rn(t3, nthr), r2 = __ffma2_rn(e, m1, top.v[q + 1]);
        const float2 r1 = __fadd2_rn(t1, nthr), r0 = __ffma2_rn(e, m1, mid.v[q + 2]);
        unsigned ax = 0, ay = 0;  // sign bits = NOT code
        ax = sign_in(ax, r7.x); ay = sign_in(ay, r7.y);
        ax = sign_in(ax, r6.x); ay = sign_in(ay, r6.y);
        ax = sign_in(ax, r5.x); ay = sign_in(ay, r5.y);
        ax = sign_in(ax, r4.x); ay = sign_in(ay, r4.y);
        ax = sign_in(ax, r3.x); ay = sign_in(ay, r3.y);
        ax = sign_in(ax, r2.x); ay = sign_in(ay, r2.y);
        ax = sign_in(ax, r1.x); ay = sign_in(ay, r1.y);
        ax = sign_in(ax, r0.x); ay = sign_in(ay, r0.y);
        lbp_hist_add(cells.pair_last[q], ax, 1u);
        lbp_hist_add(cells.pair_last[q], ay, 0x10000u);
    }
}

// Write-out of one cell pair's counters (u32 words: low half = upper cell's bin, high half = lower cell's bin) and their
// reset for the next image.  OUT8 = false: [cell][bin] u16, 8 bins per 128-bit store; OUT8 = true: u8 counts (valid when
// a cell has <= 255 pixels: the gallery form the chi-square kernels stream), 16 bins per 128-bit store.
__device__ __forceinline__ uint32_t lbp_pack8(uint4 u, unsigned sel)
{
    return __byte_perm(__byte_perm(u.x, u.y, sel), __byte_perm(u.z, u.w, sel), 0x5410);
}
template <bool OUT8>
__device__ __forceinline__ void lbp_write_pair(uint4 *hist_pair, void *out_image, unsigned pair, unsigned j, bool has_lower,
                                               unsigned lower_cell)
{
    if (OUT8) {
        // same 8 bins per step as the u16 form (two 128-bit counter loads per lane, 32 bytes apart: conflict-free; a first
        // version read 16 bins = 64 bytes per lane and paid 2-way bank conflicts: ncu 183.8 M vs 120.9 M, 1.228 vs 1.176 ms)
        uint4 *grp = hist_pair + 2 * j;
        const uint4 u = grp[0], v = grp[1];
        grp[0] = make_uint4(0, 0, 0, 0);
        grp[1] = make_uint4(0, 0, 0, 0);
        uint2 *dst = reinterpret_cast<uint2 *>(out_image);   // [cell][bin] u8: 32 groups of 8 bytes per cell
        dst[pair * 32 + j] = make_uint2(lbp_pack8(u, 0x0040), lbp_pack8(v, 0x0040));
        if (has_lower) dst[lower_cell * 32 + j] = make_uint2(lbp_pack8(u, 0x0062), lbp_pack8(v, 0x0062));
    } else {
        uint4 *grp = hist_pair + 2 * j;                      // 8 bins
        const uint4 u = grp[0], v = grp[1];
        grp[0] = make_uint4(0, 0, 0, 0);
        grp[1] = make_uint4(0, 0, 0, 0);
        uint4 *dst = reinterpret_cast<uint4 *>(out_image);   // [cell][bin] u16: 32 groups per cell
        dst[pair * 32 + j] = make_uint4(__byte_perm(u.x, u.y, 0x5410), __byte_perm(u.z, u.w, 0x5410),
                                        __byte_perm(v.x, v.y, 0x5410), __byte_perm(v.z, v.w, 0x5410));
        if (has_lower)
            dst[lower_cell * 32 + j] = make_uint4(__byte_perm(u.x, u.y, 0x7632), __byte_perm(u.z, u.w, 0x7632),
                                                  __byte_perm(v.x, v.y, 0x7632), __byte_perm(v.z, v.w, 0x7632));
    }
}

template <bool ALIGNED16, bool OUT8>
__global__ void __launch_bounds__(kLbpMaxThreads, 3) lbp_hist_kernel(const uint8_t *__restrict__ img, int64_t count, int rows,
                                                                     int cols, int grid_x, int grid_y, int img_smem_bytes,
                                                                     void *__restrict__ out_v)
{
    unsigned char *out = reinterpret_cast<unsigned char *>(out_v);
    // shared: [2 mbarriers][image buffer 0][image buffer 1][counters]
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem);
    uint8_t *s_img0 = smem + 16;
    // [half * grid_x cell pairs + 1 scratch pair][256 bins] u32 = (upper band count) | (lower band count) << 16
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem + 16 + 2 * img_smem_bytes);

    const unsigned tid = threadIdx.x, nthreads = blockDim.x;
    const unsigned ocols = cols - 2, orows = rows - 2;
    const unsigned gx = grid_x, gy = grid_y;
    const unsigned cw = ocols / gx, ch = orows / gy;
    const unsigned used_cols = cw * gx;
    const unsigned strips = (used_cols + 1) >> 1;
    const unsigned half = (gy + 1) >> 1;  // lane .y works `half` bands below lane .x
    const unsigned items = strips * half;
    const unsigned pairs = half * gx;
    const unsigned hist_vec = (pairs + 1) * kLbpPairVec;  // uint4 groups, scratch pair included
    const unsigned img_bytes = rows * cols;
    // images that start 16-byte aligned are staged by the TMA (1-D bulk copy), one image ahead of the compute
    const bool bulk_ok = ((img_bytes & 15) == 0) && ((reinterpret_cast<uintptr_t>(img) & 15) == 0);
    const unsigned hist_addr = lbp_smem_u32(s_hist);

    for (unsigned w = tid; w < hist_vec; w += nthreads) reinterpret_cast<uint4 *>(s_hist)[w] = make_uint4(0, 0, 0, 0);
    // a buffer is read 2 bytes past the last pixel by the right-most strip: keep that defined
    if (tid < 32) s_img0[(tid >> 4) * img_smem_bytes + img_smem_bytes - 16 + (tid & 15)] = 0;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lbp_smem_u32(&s_bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lbp_smem_u32(&s_bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (bulk_ok && tid == 0 && (int64_t)blockIdx.x < count)
        lbp_bulk_load(s_img0, img + (int64_t)blockIdx.x * img_bytes, img_bytes, &s_bar[0]);

    unsigned n = 0;
    for (int64_t b = blockIdx.x; b < count; b += gridDim.x, n++) {
        const unsigned buf = n & 1;
        const uint8_t *s_img = s_img0 + buf * img_smem_bytes;
        if (bulk_ok) {
            // the other buffer was last read while image n-1 was coded, two block barriers ago
            if (tid == 0 && b + gridDim.x < count)
                lbp_bulk_load(s_img0 + (buf ^ 1) * img_smem_bytes, img + (b + gridDim.x) * img_bytes, img_bytes, &s_bar[buf ^ 1]);
            lbp_mbar_wait(&s_bar[buf], (n >> 1) & 1);
        } else {
            const uint8_t *src = img + b * img_bytes;
            uint8_t *dst = s_img0 + buf * img_smem_bytes;
            for (unsigned w = tid; w < img_bytes; w += nthreads) dst[w] = __ldg(src + w);
            __syncthreads();
        }

        for (unsigned it = tid; it < items; it += nthreads) {
            const unsigned band = it / strips, x0 = (it - band * strips) * 2;
            const bool two = x0 + 1 < used_cols;
            const bool lane_y = band + half < gy;  // false only for the last band of an odd grid: its high halves are never read
            const unsigned cx0 = x0 / cw, cx1 = (x0 + 1) / cw;
            LbpCells cells;
            cells.pair_last[0] = hist_addr + (band * gx + cx0) * (kLbpPairVec * 16) + 1020;
            cells.pair_last[1] = hist_addr + (two ? band * gx + cx1 : pairs) * (kLbpPairVec * 16) + 1020;
            const uint8_t *p = s_img + band * ch * cols + x0;
            const unsigned dy = lane_y ? half * ch * cols : 0;  // an absent lower band re-reads the upper one
            LbpRow2 r[3];
            lbp_load_row2<ALIGNED16>(r[0], p, p + dy);
            lbp_load_row2<ALIGNED16>(r[1], p + cols, p + cols + dy);
            p += 2 * cols;
            for (unsigned y = 0; y < ch; y += 3) {
#pragma unroll
                for (int ph = 0; ph < 3; ph++) {
                    if (y + ph < ch) {
                        lbp_load_row2<ALIGNED16>(r[(ph + 2) % 3], p, p + dy);
                        p += cols;
                        lbp_emit_row2(r[ph % 3], r[(ph + 1) % 3], r[(ph + 2) % 3], cells);
                    }
                }
            }
        }
        __syncthreads();

        // write-out: 8 bins of a cell pair per step = two 16-byte groups -> low halves to the upper cell, high halves
        // to the lower cell ([cell][bin] u16 order, 128-bit stores); the same thread clears them for the next image
        constexpr unsigned kSteps = 32;                     // write-out steps per cell (8 bins each)
        unsigned char *dst = out + b * (int64_t)gx * gy * 256 * (OUT8 ? 1 : 2);
        for (unsigned w = tid; w < pairs * kSteps; w += nthreads) {
            const unsigned pair = w / kSteps, j = w % kSteps;
            lbp_write_pair<OUT8>(reinterpret_cast<uint4 *>(s_hist) + pair * kLbpPairVec, dst, pair, j, pair + half * gx < gx * gy,
                                 pair + half * gx);
        }
        __syncthreads();  // counters are clear before the next image's atomics
    }
}

// ---- pipelined variant (TMA-staged images): no block-wide barrier in steady state ---------------------------------
// Warps 0..W-2 code images, the last warp is the writer.  Counters and image buffers are both double-buffered:
//   full[b]  (TMA tx)        image n has landed in image buffer b = n & 1
//   done[b]  (compute warps) every compute warp has finished coding image n into counter buffer b
//   clean[b] (writer)        counter buffer b has been written out and cleared
// A compute warp that finishes image n early goes straight on to image n + 1 (other buffers); it only waits for
// things that happened a whole image ago (clean[b] of image n - 1's predecessor), so warp skew is absorbed instead of
// being paid at a barrier.  The writer requests image n + 2 as soon as done[b] frees image buffer b.
constexpr int kLbpPipeThreads = 256;

// the writer warp waits for most of an image's coding time: back off so its polling does not take issue slots
__device__ __forceinline__ void lbp_mbar_wait_relaxed(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}"
            : "=r"(done)
            : "r"(lbp_smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(256);
    }
}

__device__ __forceinline__ void lbp_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(lbp_smem_u32(bar)) : "memory");
}

template <bool ALIGNED16, bool OUT8>
__global__ void __launch_bounds__(kLbpPipeThreads, 2) lbp_hist_pipe_kernel(const uint8_t *__restrict__ img, int64_t count, int rows,
                                                                          int cols, int grid_x, int grid_y, int img_smem_bytes,
                                                                          int hist_smem_bytes, void *__restrict__ out_v)
{
    unsigned char *out = reinterpret_cast<unsigned char *>(out_v);
    // shared: [full[2], done[2], clean[2] mbarriers][image buffers 0, 1][counter buffers 0, 1]
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t *s_full = reinterpret_cast<uint64_t *>(smem), *s_done = s_full + 2, *s_clean = s_full + 4;
    uint8_t *s_img0 = smem + 64;
    unsigned char *s_hist0 = smem + 64 + 2 * img_smem_bytes;

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned n_cwarps = (blockDim.x >> 5) - 1, n_cthreads = n_cwarps * 32;
    const unsigned ocols = cols - 2, orows = rows - 2;
    const unsigned gx = grid_x, gy = grid_y;
    const unsigned cw = ocols / gx, ch = orows / gy;
    const unsigned used_cols = cw * gx;
    const unsigned strips = (used_cols + 1) >> 1;
    const unsigned half = (gy + 1) >> 1;
    const unsigned items = strips * half;
    const unsigned pairs = half * gx;
    const unsigned img_bytes = rows * cols;

    for (unsigned w = tid; w < 2u * hist_smem_bytes / 16; w += blockDim.x) reinterpret_cast<uint4 *>(s_hist0)[w] = make_uint4(0, 0, 0, 0);
    if (tid < 32) s_img0[(tid >> 4) * img_smem_bytes + img_smem_bytes - 16 + (tid & 15)] = 0;
    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lbp_smem_u32(&s_full[i])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lbp_smem_u32(&s_done[i])), "r"(n_cwarps));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lbp_smem_u32(&s_clean[i])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == n_cwarps) {
        // ---- writer / producer warp ---------------------------------------------------------------------------
        if (lane == 0) {
            for (int i = 0; i < 2; i++) {
                const int64_t b = (int64_t)blockIdx.x + (int64_t)i * gridDim.x;
                if (b < count) lbp_bulk_load(s_img0 + i * img_smem_bytes, img + b * img_bytes, img_bytes, &s_full[i]);
            }
        }
        unsigned n = 0;
        for (int64_t b = blockIdx.x; b < count; b += gridDim.x, n++) {
            const unsigned buf = n & 1, par = (n >> 1) & 1;
            lbp_mbar_wait_relaxed(&s_done[buf], par);  // image n is coded: its image buffer is free, its counters are final
            const int64_t nxt = b + 2 * (int64_t)gridDim.x;
            if (lane == 0 && nxt < count) lbp_bulk_load(s_img0 + buf * img_smem_bytes, img + nxt * img_bytes, img_bytes, &s_full[buf]);
            uint4 *hist = reinterpret_cast<uint4 *>(s_hist0 + (size_t)buf * hist_smem_bytes);
            constexpr unsigned kSteps = 32;
            unsigned char *dst = out + b * (int64_t)gx * gy * 256 * (OUT8 ? 1 : 2);
            for (unsigned w = lane; w < pairs * kSteps; w += 32) {
                const unsigned pair = w / kSteps, j = w % kSteps;
                lbp_write_pair<OUT8>(hist + pair * kLbpPairVec, dst, pair, j, pair + half * gx < gx * gy, pair + half * gx);
            }
            __syncwarp();
            if (lane == 0) lbp_mbar_arrive(&s_clean[buf]);
        }
        return;
    }

    // ---- compute warps ------------------------------------------------------------------------------------------
    unsigned n = 0;
    for (int64_t b = blockIdx.x; b < count; b += gridDim.x, n++) {
        const unsigned buf = n & 1, par = (n >> 1) & 1;
        const uint8_t *s_img = s_img0 + buf * img_smem_bytes;
        const unsigned hist_addr = lbp_smem_u32(s_hist0 + (size_t)buf * hist_smem_bytes);
        lbp_mbar_wait(&s_full[buf], par);
        if (n >= 2) lbp_mbar_wait(&s_clean[buf], par ^ 1);  // image n - 2 (previous tenant of these counters) is written out

        for (unsigned it = tid; it < items; it += n_cthreads) {
            const unsigned band = it / strips, x0 = (it - band * strips) * 2;
            const bool two = x0 + 1 < used_cols;
            const bool lane_y = band + half < gy;
            const unsigned cx0 = x0 / cw, cx1 = (x0 + 1) / cw;
            LbpCells cells;
            cells.pair_last[0] = hist_addr + (band * gx + cx0) * (kLbpPairVec * 16) + 1020;
            cells.pair_last[1] = hist_addr + (two ? band * gx + cx1 : pairs) * (kLbpPairVec * 16) + 1020;
            const uint8_t *p = s_img + band * ch * cols + x0;
            const unsigned dy = lane_y ? half * ch * cols : 0;
            LbpRow2 r[3];
            lbp_load_row2<ALIGNED16>(r[0], p, p + dy);
            lbp_load_row2<ALIGNED16>(r[1], p + cols, p + cols + dy);
            p += 2 * cols;
            for (unsigned y = 0; y < ch; y += 3) {
#pragma unroll
                for (int ph = 0; ph < 3; ph++) {
                    if (y + ph < ch) {
                        lbp_load_row2<ALIGNED16>(r[(ph + 2) % 3], p, p + dy);
                        p += cols;
                        lbp_emit_row2(r[ph % 3], r[(ph + 1) % 3], r[(ph + 2) % 3], cells);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) lbp_mbar_arrive(&s_done[buf]);
    }
}

}  // namespace frb

using namespace frb;

static int check_lbp_args(const char *fn, int64_t count, int rows, int cols, int radius, int neighbors)
{
    FRB_CHECK_ARG(count >= 0, "%s: count=%lld", fn, (long long)count);
    if (radius < 1 || radius > 32 || neighbors < 1 || neighbors > 8) {
        set_error("%s: radius 1..32 and neighbors 1..8 are implemented (got %d, %d); more than 8 neighbours would need "
                  "histograms longer than the 16384 bins the chi-square kernels keep in registers", fn, radius, neighbors);
        return FRB_ERR_UNSUPPORTED;
    }
    FRB_CHECK_ARG(rows >= 2 * radius + 1 && cols >= 2 * radius + 1, "%s: image %dx%d is smaller than the %dx%d LBP window", fn, rows,
                  cols, 2 * radius + 1, 2 * radius + 1);
    return FRB_OK;
}

// ---- any (radius, neighbors): the general elbp_ of lbph_faces.cpp ----------------------------------------------------
// The reference exposes both as options (models/lbphmodel/train_lbph_script.py:353-363, configs/lbph_config.yaml); only
// its defaults (1, 8) are on the hot path and have the tuned kernels above.  Everything else takes these plain kernels:
// same arithmetic as OpenCV — sample point n at (x, y) = (r cos(2 pi n / P), -r sin(2 pi n / P)) in float, bilinear
// weights w1..w4 in float, t = w1 a + w2 b + w3 c + w4 d with every product and sum rounded to float32 left to right,
// bit n = (t > centre) || |t - centre| < FLT_EPSILON.
struct LbpTaps {
    int fx[8], fy[8], cx[8], cy[8];
    float w1[8], w2[8], w3[8], w4[8];
};

static LbpTaps make_taps(int radius, int neighbors)
{
    LbpTaps t;
    for (int n = 0; n < 8; n++) {
        t.fx[n] = t.fy[n] = t.cx[n] = t.cy[n] = 0;
        t.w1[n] = t.w2[n] = t.w3[n] = t.w4[n] = 0.f;
    }
    for (int n = 0; n < neighbors; n++) {
        const float x = (float)(radius * cos(2.0 * 3.1415926535897932384626433832795 * n / (float)neighbors));
        const float y = (float)(-radius * sin(2.0 * 3.1415926535897932384626433832795 * n / (float)neighbors));
        const int fx = (int)floor(x), fy = (int)floor(y), cx = (int)ceil(x), cy = (int)ceil(y);
        volatile float ty = y - fy, tx = x - fx;             // volatile: every step rounded to float32, as OpenCV's build does
        volatile float omx = 1 - tx, omy = 1 - ty;
        t.fx[n] = fx; t.fy[n] = fy; t.cx[n] = cx; t.cy[n] = cy;
        volatile float a = omx * omy, b = tx * omy, c = omx * ty, d = tx * ty;
        t.w1[n] = a; t.w2[n] = b; t.w3[n] = c; t.w4[n] = d;
    }
    return t;
}

__device__ __forceinline__ unsigned lbp_code_generic(const uint8_t *__restrict__ img, int cols, int y, int x, int neighbors,
                                                     const LbpTaps &tp)
{
    const float c = (float)img[y * cols + x];
    unsigned code = 0;
    for (int n = 0; n < neighbors; n++) {
        const float a = (float)img[(y + tp.fy[n]) * cols + (x + tp.fx[n])], b = (float)img[(y + tp.fy[n]) * cols + (x + tp.cx[n])];
        const float cc = (float)img[(y + tp.cy[n]) * cols + (x + tp.fx[n])], d = (float)img[(y + tp.cy[n]) * cols + (x + tp.cx[n])];
        float t = __fmul_rn(tp.w1[n], a);
        t = __fadd_rn(t, __fmul_rn(tp.w2[n], b));
        t = __fadd_rn(t, __fmul_rn(tp.w3[n], cc));
        t = __fadd_rn(t, __fmul_rn(tp.w4[n], d));
        const bool bit = (t > c) || (fabsf(__fsub_rn(t, c)) < 1.1920928955078125e-07f);
        code |= bit ? (1u << n) : 0u;
    }
    return code;
}

__global__ void __launch_bounds__(256) lbp_codes_generic_kernel(const uint8_t *__restrict__ img, int64_t count, int rows, int cols,
                                                                int radius, int neighbors, const __grid_constant__ LbpTaps tp,
                                                                uint8_t *__restrict__ out)
{
    const int orows = rows - 2 * radius, ocols = cols - 2 * radius;
    const int64_t per = (int64_t)orows * ocols, total = count * per;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = t / per;
        const int r = (int)(t - b * per), y = r / ocols, x = r - y * ocols;
        out[t] = (uint8_t)lbp_code_generic(img + b * rows * cols, cols, y + radius, x + radius, neighbors, tp);
    }
}

// one CTA per image at a time; u32 counters [cells][2^P] in shared memory; OUT8 as above
template <bool OUT8>
__global__ void __launch_bounds__(256) lbp_hist_generic_kernel(const uint8_t *__restrict__ img, int64_t count, int rows, int cols,
                                                               int radius, int neighbors, int grid_x, int grid_y,
                                                               const __grid_constant__ LbpTaps tp, void *__restrict__ out_v)
{
    extern __shared__ unsigned lbp_gen_hist[];
    const int orows = rows - 2 * radius, ocols = cols - 2 * radius;
    const int cw = ocols / grid_x, ch = orows / grid_y, bins = 1 << neighbors, n_ctr = grid_x * grid_y * bins;
    const int used = cw * grid_x * ch * grid_y;
    for (int64_t b = blockIdx.x; b < count; b += gridDim.x) {
        for (int i = threadIdx.x; i < n_ctr; i += blockDim.x) lbp_gen_hist[i] = 0;
        __syncthreads();
        const uint8_t *im = img + b * rows * cols;
        for (int i = threadIdx.x; i < used; i += blockDim.x) {
            const int y = i / (cw * grid_x), x = i - y * (cw * grid_x);
            const unsigned code = lbp_code_generic(im, cols, y + radius, x + radius, neighbors, tp);
            atomicAdd(&lbp_gen_hist[((y / ch) * grid_x + x / cw) * bins + code], 1u);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n_ctr; i += blockDim.x) {
            if (OUT8)
                reinterpret_cast<uint8_t *>(out_v)[b * n_ctr + i] = (uint8_t)lbp_gen_hist[i];
            else
                reinterpret_cast<uint16_t *>(out_v)[b * n_ctr + i] = (uint16_t)lbp_gen_hist[i];
        }
        __syncthreads();
    }
}

// OUT8: u8 counts (cells of <= 255 pixels) instead of u16
template <bool OUT8>
static int lbp_hist_impl(const uint8_t *images, int64_t count, int rows, int cols, int radius, int neighbors, int grid_x,
                         int grid_y, void *out_hist, int *out_cell_px, void *stream)
{
    int rc = check_lbp_args("frb_lbp_hist_u8", count, rows, cols, radius, neighbors);
    if (rc != FRB_OK) return rc;
    FRB_CHECK_ARG(grid_x >= 1 && grid_y >= 1, "frb_lbp_hist_u8: grid %dx%d", grid_x, grid_y);
    const int cw = (cols - 2 * radius) / grid_x, ch = (rows - 2 * radius) / grid_y;
    if (out_cell_px) *out_cell_px = cw * ch;
    if (cw * ch > (OUT8 ? 255 : 65535)) {
        set_error("frb_lbp_hist_u8: %d pixels per cell overflow the %s counters", cw * ch, OUT8 ? "u8" : "u16");
        return FRB_ERR_UNSUPPORTED;
    }
    if (count == 0) return FRB_OK;
    FRB_CHECK_ARG(images && out_hist, "frb_lbp_hist_u8: null pointer");
    if (radius != 1 || neighbors != 8) {
        const size_t ctr_bytes = (size_t)grid_x * grid_y * ((size_t)1 << neighbors) * sizeof(unsigned);
        if (ctr_bytes > 200 * 1024) {
            set_error("frb_lbp_hist_u8: grid %dx%d with %d neighbours needs %zu B of counters (> 200 KB)", grid_x, grid_y, neighbors, ctr_bytes);
            return FRB_ERR_UNSUPPORTED;
        }
        FRB_CUDA_OK(cudaFuncSetAttribute(lbp_hist_generic_kernel<OUT8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctr_bytes));
        const LbpTaps tp = make_taps(radius, neighbors);
        const int64_t cap = (int64_t)sm_count() * 2;
        const int grid = (int)(count < cap ? count : cap);
        ProfileScope prof(FRB_K_LBP_HIST, (cudaStream_t)stream);
        lbp_hist_generic_kernel<OUT8><<<grid, 256, ctr_bytes, (cudaStream_t)stream>>>(images, count, rows, cols, radius, neighbors, grid_x,
                                                                                        grid_y, tp, out_hist);
        FRB_LAUNCH_OK("lbp_hist_generic_kernel");
        return FRB_OK;
    }
    // +16: the right-most column pair reads up to 2 bytes past the image (zeroed, never used in a code it emits)
    const int img_smem = (int)align_up((size_t)rows * cols, 16) + 16;
    // counters: one u32 per (cell pair, bin), pairs = ceil(grid_y / 2) * grid_x, plus one scratch pair
    const size_t hist_bytes = ((size_t)grid_x * ((grid_y + 1) / 2) + 1) * (kLbpPairVec * 16);
    const bool aligned16 = (cols % 2) == 0;
    // one thread per (band pair, column pair) when that fits a CTA
    const int items = ((cw * grid_x + 1) / 2) * ((grid_y + 1) / 2);
    // images the TMA can fetch (16-byte aligned, a multiple of 16 bytes) take the pipelined kernel: double-buffered
    // counters, a writer warp, no block-wide barrier
    const bool bulk_ok = (((size_t)rows * cols) & 15) == 0 && ((uintptr_t)images & 15) == 0;
    const size_t pipe_smem = 64 + 2 * (size_t)img_smem + 2 * hist_bytes;
    if (bulk_ok && pipe_smem <= 113 * 1024) {
        int threads = (items + 31) / 32 * 32 + 32;
        if (threads > kLbpPipeThreads) threads = kLbpPipeThreads;
        int per_sm = 1;
        if (aligned16) {
            FRB_CUDA_OK(cudaFuncSetAttribute(lbp_hist_pipe_kernel<true, OUT8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipe_smem));
            FRB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lbp_hist_pipe_kernel<true, OUT8>, threads, pipe_smem));
        } else {
            FRB_CUDA_OK(cudaFuncSetAttribute(lbp_hist_pipe_kernel<false, OUT8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipe_smem));
            FRB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lbp_hist_pipe_kernel<false, OUT8>, threads, pipe_smem));
        }
        if (per_sm < 1) per_sm = 1;
        const int64_t cap = (int64_t)sm_count() * per_sm;
        const int grid = (int)(count < cap ? count : cap);
        {
            ProfileScope prof(FRB_K_LBP_HIST, (cudaStream_t)stream);
            if (aligned16)
                lbp_hist_pipe_kernel<true, OUT8><<<grid, threads, pipe_smem, (cudaStream_t)stream>>>(images, count, rows, cols, grid_x, grid_y, img_smem, (int)hist_bytes, out_hist);
            else
                lbp_hist_pipe_kernel<false, OUT8><<<grid, threads, pipe_smem, (cudaStream_t)stream>>>(images, count, rows, cols, grid_x, grid_y, img_smem, (int)hist_bytes, out_hist);
        }
        FRB_LAUNCH_OK("lbp_hist_pipe_kernel");
        return FRB_OK;
    }
    const size_t smem = 16 + 2 * (size_t)img_smem + hist_bytes;
    if (smem > 227 * 1024) {
        set_error("frb_lbp_hist_u8: image %dx%d with grid %dx%d needs %zu B of shared memory (> 227 KB)", rows, cols,
                  grid_x, grid_y, smem);
        return FRB_ERR_UNSUPPORTED;
    }
    if (aligned16)
        FRB_CUDA_OK(cudaFuncSetAttribute(lbp_hist_kernel<true, OUT8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
        FRB_CUDA_OK(cudaFuncSetAttribute(lbp_hist_kernel<false, OUT8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int threads = (items + 31) / 32 * 32;
    if (threads < 64) threads = 64;
    if (threads > kLbpMaxThreads) threads = kLbpMaxThreads;
    int per_sm = 1;  // persistent grid = exactly the CTAs that can be resident
    if (aligned16)
        FRB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lbp_hist_kernel<true, OUT8>, threads, smem));
    else
        FRB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lbp_hist_kernel<false, OUT8>, threads, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t cap = (int64_t)sm_count() * per_sm;
    int grid = (int)(count < cap ? count : cap);
    {
        ProfileScope prof(FRB_K_LBP_HIST, (cudaStream_t)stream);
        if (aligned16)
            lbp_hist_kernel<true, OUT8><<<grid, threads, smem, (cudaStream_t)stream>>>(images, count, rows, cols, grid_x, grid_y, img_smem, out_hist);
        else
            lbp_hist_kernel<false, OUT8><<<grid, threads, smem, (cudaStream_t)stream>>>(images, count, rows, cols, grid_x, grid_y, img_smem, out_hist);
    }
    FRB_LAUNCH_OK("lbp_hist_kernel");
    return FRB_OK;
}


extern "C" {

int frb_lbp_codes_u8(const uint8_t *images, int64_t count, int rows, int cols, int radius, int neighbors,
                     uint8_t *out_codes, void *stream)
{
    int rc = check_lbp_args("frb_lbp_codes_u8", count, rows, cols, radius, neighbors);
    if (rc != FRB_OK) return rc;
    if (count == 0) return FRB_OK;
    FRB_CHECK_ARG(images && out_codes, "frb_lbp_codes_u8: null pointer");
    int64_t total = count * (int64_t)(rows - 2 * radius) * (cols - 2 * radius);
    if (total == 0) return FRB_OK;
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 16;
    int grid = (int)(blocks < cap ? blocks : cap);
    if (radius != 1 || neighbors != 8) {
        const LbpTaps tp = make_taps(radius, neighbors);
        lbp_codes_generic_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(images, count, rows, cols, radius, neighbors, tp, out_codes);
        FRB_LAUNCH_OK("lbp_codes_generic_kernel");
        return FRB_OK;
    }
    lbp_codes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(images, count, rows, cols, out_codes);
    FRB_LAUNCH_OK("lbp_codes_kernel");
    return FRB_OK;
}

int frb_lbp_hist_u8(const uint8_t *images, int64_t count, int rows, int cols, int radius, int neighbors, int grid_x,
                    int grid_y, uint16_t *out_hist, int *out_cell_px, void *stream)
{
    return lbp_hist_impl<false>(images, count, rows, cols, radius, neighbors, grid_x, grid_y, out_hist, out_cell_px, stream);
}

int frb_lbp_hist_u8_counts8(const uint8_t *images, int64_t count, int rows, int cols, int radius, int neighbors, int grid_x,
                            int grid_y, uint8_t *out_hist, int *out_cell_px, void *stream)
{
    return lbp_hist_impl<true>(images, count, rows, cols, radius, neighbors, grid_x, grid_y, out_hist, out_cell_px, stream);
}

}  // extern "C"
