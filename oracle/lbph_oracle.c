/*
 * oracle/lbph_oracle.c — CPU restatement of the LBPH identification path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (facerecognition_b200/)
 * may import, link or execute this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, and only as the
 * checker or as the reported CPU baseline.
 *
 * What it restates.  The reference (sin0235/FaceRecognition) delegates all LBPH
 * arithmetic to OpenCV-contrib's cv2.face.LBPHFaceRecognizer:
 *   models/lbphmodel/train_lbph.py:24-35   LBPHFaceRecognizer_create(1,8,8,8).train(faces, labels)
 *   models/lbphmodel/inference_lbph.py:5   model.predict(face) -> (label, confidence)
 *   web_app.py:587                         label, distance = model.predict(img)
 * opencv-contrib is NOT vendored in /root/reference and NOT installed in this
 * image (requirements.txt:13 "opencv-contrib-python>=4.8.0", unpinned; the only
 * version seen in the notebooks is 4.12.0.88).  The functions below restate the
 * published algorithm of opencv_contrib/modules/face/src/lbph_faces.cpp (4.x):
 * elbp_<uchar>, histc_/spatial_histogram, LBPH::train, LBPH::predict with
 * StandardCollector (strict '<' => first index wins ties), and of
 * cv::compareHist(..., HISTCMP_CHISQR_ALT) from modules/imgproc/src/histogram.cpp.
 *
 * PARITY UNPINNED for the LBP-code / histogram stage: there is no cv2.face here
 * to run and the reference's own tests hold no golden vectors for it
 * (models/lbphmodel/test_lbph_logic.py asserts ranges on unseeded data only).
 * The chi-square stage IS pinned: tests/test_oracle_lbph.py checks
 * frb_oracle_chisq_alt against the real cv2.compareHist of the installed
 * OpenCV core build.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: the float32 bilinear blend
 * must not be contracted into FMAs — x86 OpenCV wheels build the face module for
 * an SSE3 baseline, i.e. separate multiply and add, each rounded to float32).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.1415926535897932384626433832795
#endif

/* Interpolation constants of elbp_ for one sample point n (lbph_faces.cpp, elbp_):
 *   x = (float)( radius*cos(2*pi*n/(float)neighbors)),  y = (float)(-radius*sin(...))
 *   fx=floor(x) fy=floor(y) cx=ceil(x) cy=ceil(y); tx=x-fx ty=y-fy (float)
 *   w1=(1-tx)(1-ty) w2=tx(1-ty) w3=(1-tx)ty w4=tx*ty            (float)          */
typedef struct {
    int fx, fy, cx, cy;
    float w1, w2, w3, w4;
} frb_lbp_tap;

void frb_oracle_lbp_taps(int radius, int neighbors, frb_lbp_tap *taps)
{
    for (int n = 0; n < neighbors; n++) {
        float x = (float)(radius * cos(2.0 * M_PI * n / (float)neighbors));
        float y = (float)(-radius * sin(2.0 * M_PI * n / (float)neighbors));
        int fx = (int)floor(x), fy = (int)floor(y);
        int cx = (int)ceil(x), cy = (int)ceil(y);
        float ty = y - fy;
        float tx = x - fx;
        taps[n].fx = fx; taps[n].fy = fy; taps[n].cx = cx; taps[n].cy = cy;
        taps[n].w1 = (1 - tx) * (1 - ty);
        taps[n].w2 = tx * (1 - ty);
        taps[n].w3 = (1 - tx) * ty;
        taps[n].w4 = tx * ty;
    }
}

/* elbp_<uchar>: src rows x cols u8 (row stride = cols) -> dst int32
 * (rows-2r) x (cols-2r).  Bit n is set when the float32 bilinear sample
 * t = w1*a + w2*b + w3*c + w4*d (left to right, no FMA) satisfies
 * (t > centre) || (|t - centre| < FLT_EPSILON). */
void frb_oracle_elbp(const uint8_t *src, int rows, int cols, int radius, int neighbors, int32_t *dst)
{
    int orows = rows - 2 * radius, ocols = cols - 2 * radius;
    if (orows <= 0 || ocols <= 0) return;
    memset(dst, 0, sizeof(int32_t) * (size_t)orows * (size_t)ocols);
    frb_lbp_tap *taps = (frb_lbp_tap *)malloc(sizeof(frb_lbp_tap) * (size_t)neighbors);
    frb_oracle_lbp_taps(radius, neighbors, taps);
    for (int n = 0; n < neighbors; n++) {
        const frb_lbp_tap tp = taps[n];
        for (int i = radius; i < rows - radius; i++) {
            for (int j = radius; j < cols - radius; j++) {
                float t = tp.w1 * src[(i + tp.fy) * cols + (j + tp.fx)] + tp.w2 * src[(i + tp.fy) * cols + (j + tp.cx)] +
                          tp.w3 * src[(i + tp.cy) * cols + (j + tp.fx)] + tp.w4 * src[(i + tp.cy) * cols + (j + tp.cx)];
                float c = (float)src[i * cols + j];
                int bit = (t > c) || (fabsf(t - c) < FLT_EPSILON);
                dst[(i - radius) * ocols + (j - radius)] += bit << n;
            }
        }
    }
    free(taps);
}

/* spatial_histogram with integer counts: cell width = ocols/grid_x, height =
 * orows/grid_y (integer division, right/bottom remainder ignored), row index
 * i*grid_x + j, 2^neighbors bins.  Returns cell pixel count (width*height).
 * The u16 counts are the bit-exact contract; OpenCV's float view is
 * (float)count * (float)(1.0/cell_px)  (histc_ "result /= total"). */
int frb_oracle_spatial_hist_u16(const int32_t *codes, int orows, int ocols, int grid_x, int grid_y,
                                int num_patterns, uint16_t *hist)
{
    int width = ocols / grid_x, height = orows / grid_y;
    memset(hist, 0, sizeof(uint16_t) * (size_t)grid_x * (size_t)grid_y * (size_t)num_patterns);
    for (int i = 0; i < grid_y; i++)
        for (int j = 0; j < grid_x; j++) {
            uint16_t *h = hist + (size_t)(i * grid_x + j) * (size_t)num_patterns;
            for (int y = i * height; y < (i + 1) * height; y++)
                for (int x = j * width; x < (j + 1) * width; x++) {
                    int32_t v = codes[y * ocols + x];
                    if (v >= 0 && v < num_patterns) h[v]++;
                }
        }
    return width * height;
}

/* u8 image -> u16 grid histogram in one call.  Returns cell_px, or <0 on error. */
int frb_oracle_lbp_hist(const uint8_t *img, int rows, int cols, int radius, int neighbors, int grid_x,
                        int grid_y, uint16_t *hist)
{
    int orows = rows - 2 * radius, ocols = cols - 2 * radius;
    if (orows <= 0 || ocols <= 0 || neighbors < 1 || neighbors > 15) return -1;
    int32_t *codes = (int32_t *)malloc(sizeof(int32_t) * (size_t)orows * (size_t)ocols);
    if (!codes) return -2;
    frb_oracle_elbp(img, rows, cols, radius, neighbors, codes);
    int px = frb_oracle_spatial_hist_u16(codes, orows, ocols, grid_x, grid_y, 1 << neighbors, hist);
    free(codes);
    return px;
}

void frb_oracle_lbp_hist_batch(const uint8_t *imgs, int count, int rows, int cols, int radius,
                               int neighbors, int grid_x, int grid_y, uint16_t *hists)
{
    size_t L = (size_t)grid_x * (size_t)grid_y * ((size_t)1 << neighbors);
    for (int b = 0; b < count; b++)
        frb_oracle_lbp_hist(imgs + (size_t)b * rows * cols, rows, cols, radius, neighbors, grid_x, grid_y,
                            hists + (size_t)b * L);
}

/* OpenCV's normalised float32 histogram from integer counts. */
void frb_oracle_hist_to_f32(const uint16_t *hist, size_t len, int cell_px, float *out)
{
    float scale = (float)(1.0 / (double)cell_px);
    for (size_t k = 0; k < len; k++) out[k] = (float)hist[k] * scale;
}

/* cv::compareHist(h1, h2, HISTCMP_CHISQR_ALT) on float32 histograms:
 *   a = (double)h1 - (double)h2, b = (double)h1 + (double)h2,
 *   result += a*a/b when |b| > DBL_EPSILON;  result *= 2. */
double frb_oracle_chisq_alt(const float *h1, const float *h2, size_t len)
{
    double result = 0.0;
    for (size_t j = 0; j < len; j++) {
        double a = (double)h1[j] - (double)h2[j];
        double b = (double)h1[j] + (double)h2[j];
        if (fabs(b) > DBL_EPSILON) result += a * a / b;
    }
    return 2.0 * result;
}

/* All N distances of one query against a float32 gallery (N x len). */
void frb_oracle_chisq_scan(const float *gallery, size_t n, size_t len, const float *query, double *dist)
{
    for (size_t i = 0; i < n; i++) dist[i] = frb_oracle_chisq_alt(gallery + i * len, query, len);
}

/* LBPH::predict with StandardCollector: only dist < threshold is considered,
 * strict '<' against the running minimum => first index wins ties; starts from
 * (label=-1, dist=DBL_MAX).  Returns index of the winner or -1. */
long frb_oracle_predict(const float *gallery, const int32_t *labels, size_t n, size_t len, const float *query,
                        double threshold, int32_t *out_label, double *out_dist)
{
    double best = DBL_MAX;
    long best_i = -1;
    int32_t best_label = -1;
    for (size_t i = 0; i < n; i++) {
        double d = frb_oracle_chisq_alt(gallery + i * len, query, len);
        if (d < threshold && d < best) {
            best = d; best_i = (long)i; best_label = labels[i];
        }
    }
    *out_label = best_label;
    *out_dist = best;
    return best_i;
}

/* Same scan straight from u16 counts (builds the float32 view on the fly, the
 * way a CPU implementation that kept counts would); used as the CPU baseline
 * for the match stage and to cross-check the float path. */
void frb_oracle_chisq_scan_u16(const uint16_t *gallery, size_t n, size_t len, int gallery_cell_px,
                               const uint16_t *query, int query_cell_px, double *dist)
{
    float sg = (float)(1.0 / (double)gallery_cell_px), sq = (float)(1.0 / (double)query_cell_px);
    for (size_t i = 0; i < n; i++) {
        const uint16_t *g = gallery + i * len;
        double result = 0.0;
        for (size_t j = 0; j < len; j++) {
            float hf = (float)g[j] * sg, qf = (float)query[j] * sq;
            double a = (double)hf - (double)qf, b = (double)hf + (double)qf;
            if (fabs(b) > DBL_EPSILON) result += a * a / b;
        }
        dist[i] = 2.0 * result;
    }
}
