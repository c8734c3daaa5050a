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
// The blend itself uses __fmul_rn / __fadd_rn in OpenCV's left-to-right order so ptxas cannot contract it.
//
// Layout.  One CTA per image at a time (persistent, grid-stride).  The image is staged in shared
// memory; warp w owns cell-row ("band") w: lanes walk down image columns with a rolling 3x3 window
// (3 new pixels per code), and add into the band's cell histograms with shared-memory atomics.  A band's
// histograms are touched by its warp only (warp-private), so the only block barriers are around staging
// and write-out.  Counters are u16 pairs packed in u32 words in the final [cell][bin] order, so the
// write-out is a straight 128-bit copy of 32 KB per image.
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
constexpr int kLbpThreads = 256;

__global__ void __launch_bounds__(kLbpThreads) lbp_hist_kernel(const uint8_t *__restrict__ img, int64_t count, int rows,
                                                               int cols, int grid_x, int grid_y, int img_smem_bytes,
                                                               uint16_t *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    uint8_t *s_img = smem;
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem + img_smem_bytes);  // [grid_y*grid_x][128] u16 pairs

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kLbpThreads >> 5;
    const int ocols = cols - 2, orows = rows - 2;
    const int cw = ocols / grid_x, ch = orows / grid_y;
    const int used_cols = cw * grid_x;
    const int hist_words = grid_x * grid_y * 128;
    const int img_bytes = rows * cols;
    const bool vec_ok = ((img_bytes & 15) == 0) && ((reinterpret_cast<uintptr_t>(img) & 15) == 0);

    for (int w = tid; w < hist_words / 4; w += kLbpThreads) reinterpret_cast<uint4 *>(s_hist)[w] = make_uint4(0, 0, 0, 0);

    for (int64_t b = blockIdx.x; b < count; b += gridDim.x) {
        // stage the image (128-bit when every image starts 16-byte aligned)
        const uint8_t *src = img + b * img_bytes;
        if (vec_ok) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
            for (int w = tid; w < img_bytes / 16; w += kLbpThreads) reinterpret_cast<uint4 *>(s_img)[w] = __ldg(s4 + w);
        } else {
            for (int w = tid; w < img_bytes; w += kLbpThreads) s_img[w] = __ldg(src + w);
        }
        __syncthreads();

        for (int band = warp; band < grid_y; band += nwarps) {
            const int y0 = band * ch;  // first code row of the band == first source row of its window
            for (int x = lane; x < used_cols; x += 32) {
                uint32_t *cell = s_hist + (band * grid_x + x / cw) * 128;
                const uint8_t *p = s_img + y0 * cols + x;
                float a = u8_to_f32(p[0]), bb = u8_to_f32(p[1]), c = u8_to_f32(p[2]);
                p += cols;
                float d = u8_to_f32(p[0]), e = u8_to_f32(p[1]), f = u8_to_f32(p[2]);
                for (int y = 0; y < ch; y++) {
                    p += cols;
                    float g = u8_to_f32(p[0]), h = u8_to_f32(p[1]), i = u8_to_f32(p[2]);
                    unsigned code = lbp_code_r1p8(a, bb, c, d, e, f, g, h, i);
                    atomicAdd(cell + (code >> 1), 1u + (code & 1u) * 0xFFFFu);
                    a = d; bb = e; c = f;
                    d = g; e = h; f = i;
                }
            }
        }
        __syncthreads();

        // write-out (already in [cell][bin] u16 order) and clear for the next image
        uint4 *dst = reinterpret_cast<uint4 *>(out + b * (int64_t)hist_words * 2);
        for (int w = tid; w < hist_words / 4; w += kLbpThreads) {
            dst[w] = reinterpret_cast<uint4 *>(s_hist)[w];
            reinterpret_cast<uint4 *>(s_hist)[w] = make_uint4(0, 0, 0, 0);
        }
        // the barrier after the next staging pass orders these stores before the next atomics
    }
}

}  // namespace frb

using namespace frb;

static int check_lbp_args(const char *fn, int64_t count, int rows, int cols, int radius, int neighbors)
{
    FRB_CHECK_ARG(count >= 0, "%s: count=%lld", fn, (long long)count);
    if (radius != 1 || neighbors != 8) {
        set_error("%s: only radius=1, neighbors=8 is implemented (got %d, %d)", fn, radius, neighbors);
        return FRB_ERR_UNSUPPORTED;
    }
    FRB_CHECK_ARG(rows >= 3 && cols >= 3, "%s: image %dx%d is smaller than the 3x3 LBP window", fn, rows, cols);
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
    int64_t total = count * (int64_t)(rows - 2) * (cols - 2);
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 16;
    int grid = (int)(blocks < cap ? blocks : cap);
    lbp_codes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(images, count, rows, cols, out_codes);
    FRB_LAUNCH_OK("lbp_codes_kernel");
    return FRB_OK;
}

int frb_lbp_hist_u8(const uint8_t *images, int64_t count, int rows, int cols, int radius, int neighbors, int grid_x,
                    int grid_y, uint16_t *out_hist, int *out_cell_px, void *stream)
{
    int rc = check_lbp_args("frb_lbp_hist_u8", count, rows, cols, radius, neighbors);
    if (rc != FRB_OK) return rc;
    FRB_CHECK_ARG(grid_x >= 1 && grid_y >= 1, "frb_lbp_hist_u8: grid %dx%d", grid_x, grid_y);
    const int cw = (cols - 2) / grid_x, ch = (rows - 2) / grid_y;
    if (out_cell_px) *out_cell_px = cw * ch;
    if (cw * ch > 65535) {
        set_error("frb_lbp_hist_u8: %d pixels per cell overflow the u16 counters", cw * ch);
        return FRB_ERR_UNSUPPORTED;
    }
    if (count == 0) return FRB_OK;
    FRB_CHECK_ARG(images && out_hist, "frb_lbp_hist_u8: null pointer");
    const int img_smem = (int)align_up((size_t)rows * cols, 16);
    const size_t smem = (size_t)img_smem + (size_t)grid_x * grid_y * 512;
    if (smem > 227 * 1024) {
        set_error("frb_lbp_hist_u8: image %dx%d with grid %dx%d needs %zu B of shared memory (> 227 KB)", rows, cols,
                  grid_x, grid_y, smem);
        return FRB_ERR_UNSUPPORTED;
    }
    FRB_CUDA_OK(cudaFuncSetAttribute(lbp_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int64_t cap = (int64_t)sm_count() * per_sm;
    int grid = (int)(count < cap ? count : cap);
    {
        ProfileScope prof(FRB_K_LBP_HIST, (cudaStream_t)stream);
        lbp_hist_kernel<<<grid, kLbpThreads, smem, (cudaStream_t)stream>>>(images, count, rows, cols, grid_x, grid_y, img_smem,
                                                                           out_hist);
    }
    FRB_LAUNCH_OK("lbp_hist_kernel");
    return FRB_OK;
}

}  // extern "C"
