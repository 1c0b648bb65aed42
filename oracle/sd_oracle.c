/*
 * sd_oracle.c -- CPU restatement of the superviseddescent / RCR hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see sd_oracle.h).  Plain C99, libm + OpenMP.
 * Compile with -ffp-contract=off: the reference is built for baseline x86-64
 * (no FMA), so every float expression below is an IEEE mul followed by an IEEE
 * add, and the float/double mix of the original expressions is kept on purpose
 * (it decides the orientation arg-max on exact ties).
 *
 * Citations are file:line under /root/reference.
 */
#include "sd_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.141592653589793 /* hog.h:30 (VL_PI) */

/* ------------------------------------------------------------------------- */
/* HOG                                                                       */
/* ------------------------------------------------------------------------- */

int orc_hog_dimension(int variant, int K)
{
    /* hog.c:212-223: UoCTTI = 3K+4, Dalal-Triggs = 4K */
    return variant == 1 ? 3 * K + 4 : 4 * K;
}

/* hog.h:52-58 (vl_floor_f): floor for a float, returned as an integer */
static long orc_floor_f(float x)
{
    long xi = (long)x;
    if (x >= 0 || (float)xi == x) return xi;
    return xi - 1;
}

/* Orientation arg-max of one pixel.  hog.c:631-672.
 * Returns the directed bin in [0, 2K) and writes the gradient modulus. */
static int orc_pixel_bin(const float* px, int width, const float* ox, const float* oy, int K,
                         float* modulus)
{
    float gx = 0.f, gy = 0.f, g2 = 0.f;
    {
        /* single channel: the "channel with the largest gradient" loop of
         * hog.c:634-644 degenerates to "take it if its squared modulus is > 0" */
        float gx_ = px[1] - px[-1];
        float gy_ = px[width] - px[-width];
        float g2_ = gx_ * gx_ + gy_ * gy_;
        if (g2_ > g2) { gx = gx_; gy = gy_; g2 = g2_; }
    }
    float g = sqrtf(g2);                               /* hog.c:645 */
    double den = ((double)g > 1e-10) ? (double)g : 1e-10; /* VL_MAX(grad, 1e-10) is a double, :646 */
    gx = (float)((double)gx / den);                    /* :646 */
    gy = (float)((double)gy / den);                    /* :647 */

    float best = 0.f;
    int best_bin = -1;
    for (int k = 0; k < K; ++k) {                      /* :656-672 */
        float s = gx * ox[k] + gy * oy[k];
        int b = k;
        if (s < 0) { s = -s; b += K; }
        if (s > best) { best = s; best_bin = b; }      /* strict >, ascending k */
    }
    *modulus = g;
    return best_bin;
}

static void orc_orientation_tables(int K, float* ox, float* oy)
{
    for (int k = 0; k < K; ++k) {                      /* hog.c:195-204, transposed = false */
        double angle = k * ORC_PI / K;
        ox[k] = (float)cos(angle);
        oy[k] = (float)sin(angle);
    }
}

void orc_hog_orientation_bins(const float* image, int width, int height, int K, int32_t* bins)
{
    float ox[64], oy[64];
    orc_orientation_tables(K, ox, oy);
    for (int i = 0; i < width * height; ++i) bins[i] = -1;
    for (int y = 1; y < height - 1; ++y)
        for (int x = 1; x < width - 1; ++x) {
            float g;
            bins[y * width + x] = orc_pixel_bin(image + y * width + x, width, ox, oy, K, &g);
        }
}

void orc_hog_core(const float* image, int width, int height, int cs, int K, int variant, float* out)
{
    const int cw = (width + cs / 2) / cs;              /* hog.c:542-543 */
    const int ch = (height + cs / 2) / cs;
    const int cells = cw * ch;
    float ox[64], oy[64];
    orc_orientation_tables(K, ox, oy);

    float* hist = (float*)calloc((size_t)cells * 2 * K, sizeof(float)); /* hog.c:569 */
    float* energy = (float*)calloc((size_t)cells, sizeof(float));        /* hog.c:570 */

    /* ---- vl_hog_put_image, hog.c:616-727 ---- */
    for (int y = 1; y < height - 1; ++y) {
        for (int x = 1; x < width - 1; ++x) {
            float g;
            int bin = orc_pixel_bin(image + y * width + x, width, ox, oy, K, &g);
            if (bin < 0) continue;                     /* :694 (no orientation won: zero gradient) */

            /* :697-709 ; orientationWeights[0] == 1 (hard assignment, :680) */
            float hx = (float)((x + 0.5) / (double)cs - 0.5);
            float hy = (float)((y + 0.5) / (double)cs - 0.5);
            long bx = orc_floor_f(hx);
            long by = orc_floor_f(hy);
            float wx2 = hx - (float)bx;
            float wy2 = hy - (float)by;
            float wx1 = (float)(1.0 - (double)wx2);
            float wy1 = (float)(1.0 - (double)wy2);
            wx1 *= 1.0f; wx2 *= 1.0f; wy1 *= 1.0f; wy2 *= 1.0f;

            float* plane = hist + (size_t)bin * cells;
            if (bx >= 0 && by >= 0)                  plane[bx + by * cw] += g * wx1 * wy1;           /* :713-715 */
            if (bx < cw - 1 && by >= 0)              plane[bx + 1 + by * cw] += g * wx2 * wy1;       /* :716-718 */
            if (bx < cw - 1 && by < ch - 1)          plane[bx + 1 + (by + 1) * cw] += g * wx2 * wy2; /* :719-721 */
            if (bx >= 0 && by < ch - 1)              plane[bx + (by + 1) * cw] += g * wx1 * wy2;     /* :722-724 */
        }
    }

    /* ---- vl_hog_extract, hog.c:875-890: undirected cell energy ---- */
    for (int k = 0; k < K; ++k)
        for (int c = 0; c < cells; ++c) {
            float h = hist[(size_t)k * cells + c] + hist[(size_t)(k + K) * cells + c];
            energy[c] += h * h;
        }

    /* ---- block normalisation, hog.c:924-1061 ---- */
    for (int y = 0; y < ch; ++y) {
        for (int x = 0; x < cw; ++x) {
            int xm = x - 1 > 0 ? x - 1 : 0, xp = x + 1 < cw - 1 ? x + 1 : cw - 1;   /* :930-933 */
            int ym = y - 1 > 0 ? y - 1 : 0, yp = y + 1 < ch - 1 ? y + 1 : ch - 1;
            double n1 = energy[xm + ym * cw], n2 = energy[x + ym * cw], n3 = energy[xp + ym * cw];
            double n4 = energy[xm + y * cw],  n5 = energy[x + y * cw],  n6 = energy[xp + y * cw];
            double n7 = energy[xm + yp * cw], n8 = energy[x + yp * cw], n9 = energy[xp + yp * cw];
            /* :978-981 (non-transposed branch) */
            double f1 = 1.0 / sqrt(n1 + n2 + n4 + n5 + 1e-4);
            double f2 = 1.0 / sqrt(n2 + n3 + n5 + n6 + 1e-4);
            double f3 = 1.0 / sqrt(n4 + n5 + n7 + n8 + 1e-4);
            double f4 = 1.0 / sqrt(n5 + n6 + n8 + n9 + 1e-4);
            double t1 = 0, t2 = 0, t3 = 0, t4 = 0;
            float* o = out + x + cw * y;
            const float* hc = hist + x + cw * y;
            for (int k = 0; k < K; ++k) {              /* :985-1044 */
                double ha = hc[(size_t)cells * k];
                double hb = hc[(size_t)cells * (k + K)];
                double ha1 = f1 * ha, ha2 = f2 * ha, ha3 = f3 * ha, ha4 = f4 * ha;
                double hb1 = f1 * hb, hb2 = f2 * hb, hb3 = f3 * hb, hb4 = f4 * hb;
                double hc1 = ha1 + hb1, hc2 = ha2 + hb2, hc3 = ha3 + hb3, hc4 = ha4 + hb4;
#define ORC_CLAMP(v) ((0.2 < (v)) ? 0.2 : (v))        /* VL_MIN(0.2, v) */
                ha1 = ORC_CLAMP(ha1); ha2 = ORC_CLAMP(ha2); ha3 = ORC_CLAMP(ha3); ha4 = ORC_CLAMP(ha4);
                hb1 = ORC_CLAMP(hb1); hb2 = ORC_CLAMP(hb2); hb3 = ORC_CLAMP(hb3); hb4 = ORC_CLAMP(hb4);
                hc1 = ORC_CLAMP(hc1); hc2 = ORC_CLAMP(hc2); hc3 = ORC_CLAMP(hc3); hc4 = ORC_CLAMP(hc4);
#undef ORC_CLAMP
                t1 += hc1; t2 += hc2; t3 += hc3; t4 += hc4;
                if (variant == 1) {                    /* UoCTTI, :1026-1033 */
                    o[(size_t)cells * k]           = (float)(0.5 * (ha1 + ha2 + ha3 + ha4));
                    o[(size_t)cells * (k + K)]     = (float)(0.5 * (hb1 + hb2 + hb3 + hb4));
                    o[(size_t)cells * (k + 2 * K)] = (float)(0.5 * (hc1 + hc2 + hc3 + hc4));
                } else {                               /* Dalal-Triggs, :1035-1040 */
                    o[(size_t)cells * k]           = (float)hc1;
                    o[(size_t)cells * (k + K)]     = (float)hc2;
                    o[(size_t)cells * (k + 2 * K)] = (float)hc3;
                    o[(size_t)cells * (k + 3 * K)] = (float)hc4;
                }
            }
            if (variant == 1) {                        /* texture dims, :1046-1053 */
                const float c = 1.0f / sqrtf(18.0f);
                o[(size_t)cells * (3 * K + 0)] = (float)(c * t1);
                o[(size_t)cells * (3 * K + 1)] = (float)(c * t2);
                o[(size_t)cells * (3 * K + 2)] = (float)(c * t3);
                o[(size_t)cells * (3 * K + 3)] = (float)(c * t4);
            }
        }
    }
    free(hist);
    free(energy);
}

/* ------------------------------------------------------------------------- */
/* OpenCV arithmetic (un-vendored dependency; pinned against cv2 4.13)       */
/* ------------------------------------------------------------------------- */

/* cvRound(float): round-half-to-even (SSE cvtss2si under the default MXCSR). */
int orc_cv_round(float v) { return (int)lrintf(v); }

static short orc_sat_short_round(float v)
{
    long r = lrintf(v);
    if (r > 32767) r = 32767;
    if (r < -32768) r = -32768;
    return (short)r;
}

static int orc_clip(int x, int a, int b) { return x >= a ? (x < b ? x : b - 1) : a; }

/* cv::resize(src, dst, dsize) with the default INTER_LINEAR on CV_8UC1
 * (adaptive_vlhog.hpp:154-155).  Fixed-point bilinear: 11-bit coefficients,
 * horizontal pass into int32, vertical pass ((b*(t>>4))>>16), +2 >>2. */
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                          uint8_t* dst, int dw, int dh, int dstride)
{
    const double inv_sx = (double)dw / sw, inv_sy = (double)dh / sh;
    const double scale_x = 1. / inv_sx, scale_y = 1. / inv_sy;
    int* xofs = (int*)malloc(sizeof(int) * dw);
    short* xa = (short*)malloc(sizeof(short) * 2 * dw);
    int* row0 = (int*)malloc(sizeof(int) * dw);
    int* row1 = (int*)malloc(sizeof(int) * dw);

    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        xa[2 * dx] = orc_sat_short_round((1.f - fx) * 2048.f);
        xa[2 * dx + 1] = orc_sat_short_round(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        short b0 = orc_sat_short_round((1.f - fy) * 2048.f);
        short b1 = orc_sat_short_round(fy * 2048.f);
        const uint8_t* s0 = src + (size_t)orc_clip(sy, 0, sh) * sstride;
        const uint8_t* s1 = src + (size_t)orc_clip(sy + 1, 0, sh) * sstride;
        for (int dx = 0; dx < dw; ++dx) {
            int sx = xofs[dx];
            int sx1 = sx + 1 < sw ? sx + 1 : sx;   /* tap with zero weight when clamped */
            row0[dx] = s0[sx] * xa[2 * dx] + s0[sx1] * xa[2 * dx + 1];
            row1[dx] = s1[sx] * xa[2 * dx] + s1[sx1] * xa[2 * dx + 1];
        }
        uint8_t* d = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; ++dx)
            d[dx] = (uint8_t)((((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2);
    }
    free(xofs); free(xa); free(row0); free(row1);
}

/* cv::cvtColor(BGR2GRAY) on 8U, cv2 >= 3 constants (15-bit), adaptive_vlhog.hpp:116 */
void orc_bgr2gray_u8(const uint8_t* bgr, int w, int h, int sstride, uint8_t* gray, int dstride)
{
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = bgr + (size_t)y * sstride;
        uint8_t* d = gray + (size_t)y * dstride;
        for (int x = 0; x < w; ++x)
            d[x] = (uint8_t)((3735 * s[3 * x] + 19235 * s[3 * x + 1] + 9798 * s[3 * x + 2] + (1 << 14)) >> 15);
    }
}

/* ------------------------------------------------------------------------- */
/* RCR glue                                                                  */
/* ------------------------------------------------------------------------- */

/* helpers.hpp:136-160: eye centre = mean of the named landmarks (float Vec2f
 * arithmetic), IED = cv::norm(right, left, NORM_L2): float difference, double
 * accumulation of squares, double sqrt. */
double orc_get_ied(const float* row, int L, const int32_t* ridx, int nr, const int32_t* lidx, int nl)
{
    float rx = 0.f, ry = 0.f, lx = 0.f, ly = 0.f;
    for (int i = 0; i < nr; ++i) { rx += row[ridx[i]]; ry += row[ridx[i] + L]; }
    { float inv = 1.f / (float)nr; rx = rx * inv; ry = ry * inv; }   /* cv::Vec /= float multiplies by 1.f/alpha */
    for (int i = 0; i < nl; ++i) { lx += row[lidx[i]]; ly += row[lidx[i] + L]; }
    { float inv = 1.f / (float)nl; lx = lx * inv; ly = ly * inv; }
    double dx = (double)(rx - lx), dy = (double)(ry - ly);
    return sqrt(dx * dx + dy * dy);
}

/* adaptive_vlhog.hpp:123: int half = std::round(float rel * double ied / 2) */
int orc_patch_half(float rel, double ied) { return (int)round((double)rel * ied / 2); }

/* adaptive_vlhog.hpp:135-151: [c-half, c+half)^2, zero outside the image
 * (copyMakeBorder BORDER_CONSTANT 0 then ROI == zero-padded sampling). */
void orc_crop_patch_u8(const uint8_t* image, int w, int h, int stride, int cx, int cy, int half,
                       uint8_t* patch)
{
    const int P = 2 * half;
    for (int py = 0; py < P; ++py) {
        int iy = cy - half + py;
        for (int px = 0; px < P; ++px) {
            int ix = cx - half + px;
            patch[py * P + px] = (ix >= 0 && ix < w && iy >= 0 && iy < h) ? image[(size_t)iy * stride + ix] : 0;
        }
    }
}

int orc_feature_length(int L, const orc_hog_param* p)
{
    return L * p->num_cells * p->num_cells * orc_hog_dimension(p->variant, p->num_bins) + 1;
}

void orc_patch_geometry(const float* params, int L, const orc_hog_param* p, const int32_t* ridx, int nr,
                        const int32_t* lidx, int nl, int32_t* cx, int32_t* cy, int32_t* half)
{
    double ied = orc_get_ied(params, L, ridx, nr, lidx, nl);
    int hf = orc_patch_half(p->relative_patch_size, ied);
    for (int i = 0; i < L; ++i) {
        cx[i] = orc_cv_round(params[i]);          /* adaptive_vlhog.hpp:132 */
        cy[i] = orc_cv_round(params[i + L]);      /* :133 */
        half[i] = hf;
    }
}

/* HogTransform::operator(), adaptive_vlhog.hpp:109-185 */
int orc_hog_transform(const uint8_t* image, int w, int h, int stride, const float* params, int L,
                      const orc_hog_param* p, const int32_t* ridx, int nr, const int32_t* lidx, int nl,
                      orc_hog_core_fn hog_core, float* out_row)
{
    if (!hog_core) hog_core = orc_hog_core;
    const int nc = p->num_cells, fs = p->num_cells * p->cell_size;   /* :154 */
    const int dd = orc_hog_dimension(p->variant, p->num_bins);
    const int per_lm = nc * nc * dd;
    double ied = orc_get_ied(params, L, ridx, nr, lidx, nl);
    int half = orc_patch_half(p->relative_patch_size, ied);           /* :123 */
    if (half <= 0) return 1;   /* cv::resize would throw on an empty ROI */
    const int P = 2 * half;

    uint8_t* patch = (uint8_t*)malloc((size_t)P * P);
    uint8_t* small = (uint8_t*)malloc((size_t)fs * fs);
    float* fimg = (float*)malloc(sizeof(float) * fs * fs);
    float* planar = (float*)malloc(sizeof(float) * per_lm);

    for (int i = 0; i < L; ++i) {
        int cx = orc_cv_round(params[i]);
        int cy = orc_cv_round(params[i + L]);
        orc_crop_patch_u8(image, w, h, stride, cx, cy, half, patch);
        orc_resize_linear_u8(patch, P, P, P, small, fs, fs, fs);       /* :155 */
        for (int k = 0; k < fs * fs; ++k) fimg[k] = (float)small[k];   /* :157 */
        hog_core(fimg, fs, fs, p->cell_size, p->num_bins, p->variant, planar);  /* :158-165 */
        /* :166-176: per dimension j, transpose the (hh x ww) plane then flatten:
         * out[j*nc*nc + cx*nc + cy] = planar[j*nc*nc + cy*nc + cx] */
        float* o = out_row + (size_t)i * per_lm;
        for (int j = 0; j < dd; ++j)
            for (int yy = 0; yy < nc; ++yy)
                for (int xx = 0; xx < nc; ++xx)
                    o[j * nc * nc + xx * nc + yy] = planar[j * nc * nc + yy * nc + xx];
    }
    out_row[(size_t)L * per_lm] = 1.0f;                                 /* bias, :182-183 */
    free(patch); free(small); free(fimg); free(planar);
    return 0;
}

/* examples/landmark_detection.cpp:195-261 -- the non-adaptive HogTransform of the hello-world example:
 * patch half-size = num_cells * (cell_size / 2) (:213), zero padding outside the frame (:222-234), NO resize, vl_hog on
 * the patch itself (:238-246), per-dimension transpose (:247-256), NO bias column.  The HOG grid is whatever vl_hog
 * derives from the patch: (P + cell_size/2) / cell_size cells per side (hog.c:548-549). out_row: L * hw*hw*dd floats;
 * *out_len (may be NULL) receives that length. */
int orc_hog_transform_fixed(const uint8_t* image, int w, int h, int stride, const float* params, int L,
                            const orc_hog_param* p, orc_hog_core_fn hog_core, float* out_row, int* out_len)
{
    if (!hog_core) hog_core = orc_hog_core;
    const int half = p->num_cells * (p->cell_size / 2);
    const int P = 2 * half;
    if (P <= 3) return 1;
    const int hw = (P + p->cell_size / 2) / p->cell_size;
    const int dd = orc_hog_dimension(p->variant, p->num_bins);
    const int per_lm = hw * hw * dd;
    if (out_len) *out_len = L * per_lm;
    if (!out_row) return 0;
    uint8_t* patch = (uint8_t*)malloc((size_t)P * P);
    float* fimg = (float*)malloc(sizeof(float) * P * P);
    float* planar = (float*)malloc(sizeof(float) * per_lm);
    for (int i = 0; i < L; ++i) {
        int cx = orc_cv_round(params[i]);
        int cy = orc_cv_round(params[i + L]);
        orc_crop_patch_u8(image, w, h, stride, cx, cy, half, patch);
        for (int k = 0; k < P * P; ++k) fimg[k] = (float)patch[k];
        hog_core(fimg, P, P, p->cell_size, p->num_bins, p->variant, planar);
        float* o = out_row + (size_t)i * per_lm;
        for (int j = 0; j < dd; ++j)
            for (int yy = 0; yy < hw; ++yy)
                for (int xx = 0; xx < hw; ++xx)
                    o[j * hw * hw + xx * hw + yy] = planar[j * hw * hw + yy * hw + xx];
    }
    free(patch); free(fimg); free(planar);
    return 0;
}

int orc_hog_transform_batch(const uint8_t* images, int count, int w, int h, int stride,
                            const float* params, int L, const orc_hog_param* p,
                            const int32_t* ridx, int nr, const int32_t* lidx, int nl,
                            orc_hog_core_fn hog_core, int threads, float* out, int out_ld)
{
    int rc = 0;
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 4)
    for (int i = 0; i < count; ++i) {
        int r = orc_hog_transform(images + (size_t)i * h * stride, w, h, stride, params + (size_t)i * 2 * L,
                                  L, p, ridx, nr, lidx, nl, hog_core, out + (size_t)i * out_ld);
        if (r) {
#pragma omp atomic write
            rc = r;
        }
    }
    return rc;
}

/* apps/rcr/rcr-train.cpp:130-146: perturb a face box by a translation (fractions of its size) and a scaling about
 * its centre.  All arithmetic in float; cv::Rect(int) truncates the float expressions toward zero. */
void orc_perturb_box(const int32_t box[4], float tx, float ty, float scaling, int32_t out[4])
{
    const float tx_pixel = tx * (float)box[2];
    const float ty_pixel = ty * (float)box[3];
    const float pw = (float)box[2] * scaling;
    const float ph = (float)box[3] * scaling;
    out[0] = (int32_t)((float)box[0] + ((float)box[2] - pw) / 2.0f + tx_pixel);
    out[1] = (int32_t)((float)box[1] + ((float)box[3] - ph) / 2.0f + ty_pixel);
    out[2] = (int32_t)pw;
    out[3] = (int32_t)ph;
}

/* apps/rcr/rcr-train.cpp:149-212: per-landmark L2 error of every row, normalised by the inter-eye distance of the
 * PREDICTION.  cv::norm(Vec2f, Vec2f): float differences, squares summed in double, sqrt in double, stored as float
 * (:169); .mul(1.0f / ied): the double quotient becomes a float factor, the product is a float multiply
 * (OpenCV arithm on CV_32F with a scalar operand -- "parity unpinned": OpenCV C++ cannot run here). */
void orc_normalised_landmark_errors(const float* pred, const float* gt, int N, int L, const int32_t* ridx, int nr,
                                    const int32_t* lidx, int nl, float* out)
{
    for (int r = 0; r < N; ++r) {
        const float* p = pred + (size_t)r * 2 * L;
        const float* g = gt + (size_t)r * 2 * L;
        const double ied = orc_get_ied(p, L, ridx, nr, lidx, nl);
        const float f = (float)((double)1.0f / ied);
        for (int i = 0; i < L; ++i) {
            const float dx = p[i] - g[i], dy = p[i + L] - g[i + L];
            const float n = (float)sqrt((double)dx * (double)dx + (double)dy * (double)dy);
            out[(size_t)r * L + i] = n * f;
        }
    }
}

/* model.hpp:64-76 */
void orc_align_mean(const float* mean, int L, int bx, int by, int bw, int bh,
                    float sx, float sy, float tx, float ty, float* out)
{
    /* OpenCV folds the whole MatExpr ((m*s + 0.5f + t) * w + x) into ONE
     * convertTo(alpha, beta): alpha = s*w, beta = (0.5 + t)*w + x (doubles), and
     * the 32f->32f scaled conversion computes m*(float)alpha + (float)beta in float.
     * (Un-fused mul+add: baseline-SSE2 behaviour; an AVX2 OpenCV build may fuse it.
     * This step is "parity unpinned" -- OpenCV C++ cannot be executed here.) */
    const float ax = (float)((double)sx * (double)bw);
    const float bxf = (float)(((double)0.5f + (double)tx) * (double)bw + (double)bx);
    const float ay = (float)((double)sy * (double)bh);
    const float byf = (float)(((double)0.5f + (double)ty) * (double)bh + (double)by);
    for (int i = 0; i < L; ++i) {
        out[i] = mean[i] * ax + bxf;
        out[i + L] = mean[i + L] * ay + byf;
    }
}

/* model.hpp:94-98 (ones/ied) and superviseddescent.hpp:213,301,338 (1/norm) */
void orc_ied_normaliser(double ied, float* norm, float* inv_norm)
{
    float n = (float)(1.0 / ied);   /* Mat::ones / double -> ones * (1/ied), rounded to float */
    *norm = n;
    *inv_norm = 1.0f / n;           /* 1 / Mat: per-element float reciprocal */
}

/* ------------------------------------------------------------------------- */
/* Regressor                                                                 */
/* ------------------------------------------------------------------------- */

void orc_gram(const float* A, int N, int D, int precision, float* AtA)
{
    /* regressors.hpp:208 / verbose_solver.hpp:67: AtA = A^T * A */
    if (precision == 0) {
        memset(AtA, 0, sizeof(float) * (size_t)D * D);
        /* rank-1 updates row by row: float accumulation, ascending sample order */
        for (int n = 0; n < N; ++n) {
            const float* a = A + (size_t)n * D;
#pragma omp parallel for schedule(static) if (D > 256)
            for (int i = 0; i < D; ++i) {
                float ai = a[i];
                float* g = AtA + (size_t)i * D;
                for (int j = 0; j < D; ++j) g[j] += ai * a[j];
            }
        }
    } else {
        double* G = (double*)calloc((size_t)D * D, sizeof(double));
        for (int n = 0; n < N; ++n) {
            const float* a = A + (size_t)n * D;
#pragma omp parallel for schedule(static) if (D > 256)
            for (int i = 0; i < D; ++i) {
                double ai = a[i];
                double* g = G + (size_t)i * D;
                for (int j = 0; j < D; ++j) g[j] += ai * (double)a[j];
            }
        }
        for (size_t k = 0; k < (size_t)D * D; ++k) AtA[k] = (float)G[k];
        free(G);
    }
}

float orc_regulariser_lambda(const orc_regulariser* r, const float* AtA, int D, int n_train)
{
    float lambda = r->lambda;
    if (r->type == 1) {
        /* regressors.hpp:133-136: lambda * (float)cv::norm(AtA) / (float)N ;
         * cv::norm(CV_32F, NORM_L2) accumulates squares in double */
        double s = 0;
        for (size_t k = 0; k < (size_t)D * D; ++k) s += (double)AtA[k] * (double)AtA[k];
        lambda = lambda * (float)sqrt(s) / (float)n_train;
    }
    return lambda;
}

/* Right-looking LU with partial (row) pivoting, then forward/back substitution.
 * Restates Eigen::PartialPivLU + solve as used at regressors.hpp:224-225. */
#define ORC_LU_IMPL(T, NAME)                                                                     \
    static int NAME(T* G, T* R, int D, int M)                                                    \
    {                                                                                            \
        int singular = 0;                                                                        \
        for (int k = 0; k < D; ++k) {                                                            \
            int piv = k;                                                                         \
            T best = (T)fabs((double)G[(size_t)k * D + k]);                                      \
            for (int i = k + 1; i < D; ++i) {                                                    \
                T v = (T)fabs((double)G[(size_t)i * D + k]);                                     \
                if (v > best) { best = v; piv = i; }                                             \
            }                                                                                    \
            if (best == 0) { singular = 1; continue; }                                           \
            if (piv != k) {                                                                      \
                for (int j = 0; j < D; ++j) { T t = G[(size_t)k * D + j]; G[(size_t)k * D + j] = G[(size_t)piv * D + j]; G[(size_t)piv * D + j] = t; } \
                for (int j = 0; j < M; ++j) { T t = R[(size_t)k * M + j]; R[(size_t)k * M + j] = R[(size_t)piv * M + j]; R[(size_t)piv * M + j] = t; } \
            }                                                                                    \
            T pivv = G[(size_t)k * D + k];                                                       \
            _Pragma("omp parallel for schedule(static) if (D - k > 512)")                        \
            for (int i = k + 1; i < D; ++i) {                                                    \
                T l = G[(size_t)i * D + k] / pivv;                                               \
                G[(size_t)i * D + k] = l;                                                        \
                if (l != 0) {                                                                    \
                    T* gi = G + (size_t)i * D; const T* gk = G + (size_t)k * D;                  \
                    for (int j = k + 1; j < D; ++j) gi[j] -= l * gk[j];                          \
                    T* ri = R + (size_t)i * M; const T* rk = R + (size_t)k * M;                  \
                    for (int j = 0; j < M; ++j) ri[j] -= l * rk[j];                              \
                }                                                                                \
            }                                                                                    \
        }                                                                                        \
        for (int i = D - 1; i >= 0; --i) {                                                       \
            T* ri = R + (size_t)i * M;                                                           \
            for (int k = i + 1; k < D; ++k) {                                                    \
                T u = G[(size_t)i * D + k]; const T* rk = R + (size_t)k * M;                     \
                for (int j = 0; j < M; ++j) ri[j] -= u * rk[j];                                  \
            }                                                                                    \
            T d = G[(size_t)i * D + i];                                                          \
            for (int j = 0; j < M; ++j) ri[j] /= d;                                              \
        }                                                                                        \
        return singular;                                                                         \
    }
ORC_LU_IMPL(float, orc_lu_solve_f)
ORC_LU_IMPL(double, orc_lu_solve_d)

int orc_solve(const float* A, const float* B, int N, int D, int M, const orc_regulariser* r,
              int precision, float* X, float* lambda_out)
{
    float* AtA = (float*)malloc(sizeof(float) * (size_t)D * D);
    orc_gram(A, N, D, precision, AtA);
    float lambda = orc_regulariser_lambda(r, AtA, D, N);       /* regressors.hpp:212 (data.rows) */
    if (lambda_out) *lambda_out = lambda;
    int rc;
    if (precision == 0) {
        float* R = (float*)calloc((size_t)D * M, sizeof(float));
        for (int n = 0; n < N; ++n)                              /* A^T * labels, :225 */
            for (int i = 0; i < D; ++i) {
                float ai = A[(size_t)n * D + i];
                for (int j = 0; j < M; ++j) R[(size_t)i * M + j] += ai * B[(size_t)n * M + j];
            }
        for (int i = 0; i < D; ++i)                              /* :215-221 */
            AtA[(size_t)i * D + i] += (i == D - 1 && !r->regularise_last_row) ? 0.0f : lambda;
        rc = orc_lu_solve_f(AtA, R, D, M);
        memcpy(X, R, sizeof(float) * (size_t)D * M);
        free(R);
    } else {
        double* G = (double*)malloc(sizeof(double) * (size_t)D * D);
        double* R = (double*)calloc((size_t)D * M, sizeof(double));
        /* recompute the Gram in double (AtA above was rounded to float for the lambda rule) */
        memset(G, 0, sizeof(double) * (size_t)D * D);
        for (int n = 0; n < N; ++n) {
            const float* a = A + (size_t)n * D;
            for (int i = 0; i < D; ++i) {
                double ai = a[i];
                for (int j = 0; j < D; ++j) G[(size_t)i * D + j] += ai * (double)a[j];
                for (int j = 0; j < M; ++j) R[(size_t)i * M + j] += ai * (double)B[(size_t)n * M + j];
            }
        }
        for (int i = 0; i < D; ++i)
            G[(size_t)i * D + i] += (i == D - 1 && !r->regularise_last_row) ? 0.0 : (double)lambda;
        rc = orc_lu_solve_d(G, R, D, M);
        for (size_t k = 0; k < (size_t)D * M; ++k) X[k] = (float)R[k];
        free(G); free(R);
    }
    free(AtA);
    return rc;
}

/* regressors.hpp:377-381: values * x (cv::gemm on CV_32F accumulates in double) */
void orc_predict(const float* values, int N, int D, const float* X, int M, float* out)
{
#pragma omp parallel for schedule(static) if (N > 64)
    for (int n = 0; n < N; ++n) {
        double acc[512];
        for (int j0 = 0; j0 < M; j0 += 512) {
            int mj = M - j0 < 512 ? M - j0 : 512;
            for (int j = 0; j < mj; ++j) acc[j] = 0;
            for (int i = 0; i < D; ++i) {
                double v = values[(size_t)n * D + i];
                const float* xr = X + (size_t)i * M + j0;
                for (int j = 0; j < mj; ++j) acc[j] += v * (double)xr[j];
            }
            for (int j = 0; j < mj; ++j) out[(size_t)n * M + j0 + j] = (float)acc[j];
        }
    }
}

/* regressors.hpp:361-369: ||pred - labels||_2 / ||labels||_2 */
double orc_test_residual(const float* data, const float* labels, int N, int D, const float* X, int M)
{
    float* pred = (float*)malloc(sizeof(float) * (size_t)N * M);
    orc_predict(data, N, D, X, M, pred);
    double num = 0, den = 0;
    for (size_t k = 0; k < (size_t)N * M; ++k) {
        double d = (double)(pred[k] - labels[k]);   /* float difference, double square (cv::norm) */
        num += d * d;
        den += (double)labels[k] * (double)labels[k];
    }
    free(pred);
    return sqrt(num) / sqrt(den);
}

/* ------------------------------------------------------------------------- */
/* Cascade                                                                   */
/* ------------------------------------------------------------------------- */

static void orc_norm_factors(const orc_normalisation* nm, const float* x_row, int P, float* n, float* inv_n)
{
    if (!nm || nm->kind == 0) { *n = 1.0f; *inv_n = 1.0f; return; }   /* superviseddescent.hpp:60-74 */
    double ied = orc_get_ied(x_row, P / 2, nm->right_idx, nm->n_right, nm->left_idx, nm->n_left);
    orc_ied_normaliser(ied, n, inv_n);
}

/* x_next = x - (A X) * (1/norm(x)), superviseddescent.hpp:209-215 / :296-301 / :337-339 */
static void orc_apply_level(const float* A, const float* X, const float* cur, int N, int D, int P,
                            const orc_normalisation* nm, float* next)
{
    float* upd = (float*)malloc(sizeof(float) * (size_t)N * P);
    orc_predict(A, N, D, X, P, upd);
    for (int i = 0; i < N; ++i) {
        float n, inv_n;
        orc_norm_factors(nm, cur + (size_t)i * P, P, &n, &inv_n);
        for (int j = 0; j < P; ++j)
            next[(size_t)i * P + j] = cur[(size_t)i * P + j] - upd[(size_t)i * P + j] * inv_n;
    }
    free(upd);
}

static void orc_project_all(const float* cur, const float* templates, int N, int P, int D, int level,
                            orc_projection_fn h, void* user, float* A)
{
    /* superviseddescent.hpp:173-197: h per sample (thread pool in the reference) */
    for (int i = 0; i < N; ++i) h(cur + (size_t)i * P, P, level, i, A + (size_t)i * D, user);
    if (templates)
        for (size_t k = 0; k < (size_t)N * D; ++k) A[k] = A[k] - templates[k];
}

int orc_cascade_train(const float* x_gt, const float* x0, const float* templates, int N, int P,
                      int num_levels, const int* feat_dims, const orc_regulariser* regs,
                      const orc_normalisation* nm, orc_projection_fn h, void* user, int precision,
                      float** weights, float* x_final, orc_epoch_cb cb, void* cb_user)
{
    float* cur = (float*)malloc(sizeof(float) * (size_t)N * P);
    float* next = (float*)malloc(sizeof(float) * (size_t)N * P);
    float* b = (float*)malloc(sizeof(float) * (size_t)N * P);
    memcpy(cur, x0, sizeof(float) * (size_t)N * P);
    int rc = 0;
    for (int level = 0; level < num_levels; ++level) {
        int D = feat_dims[level];
        float* A = (float*)malloc(sizeof(float) * (size_t)N * D);
        orc_project_all(cur, templates, N, P, D, level, h, user, A);
        for (int i = 0; i < N; ++i) {                          /* :199-205 */
            float n, inv_n;
            orc_norm_factors(nm, cur + (size_t)i * P, P, &n, &inv_n);
            for (int j = 0; j < P; ++j)
                b[(size_t)i * P + j] = (cur[(size_t)i * P + j] - x_gt[(size_t)i * P + j]) * n;
        }
        rc |= orc_solve(A, b, N, D, P, &regs[level], precision, weights[level], NULL);   /* :207 */
        orc_apply_level(A, weights[level], cur, N, D, P, nm, next);                      /* :209-215 */
        float* t = cur; cur = next; next = t;
        if (cb) cb(cur, N, P, level, cb_user);                 /* :217 */
        free(A);
    }
    if (x_final) memcpy(x_final, cur, sizeof(float) * (size_t)N * P);
    free(cur); free(next); free(b);
    return rc;
}

int orc_cascade_apply(const float* x0, const float* templates, int N, int P, int num_levels,
                      const int* feat_dims, float* const* weights, const orc_normalisation* nm,
                      orc_projection_fn h, void* user, float* x_final, orc_epoch_cb cb, void* cb_user)
{
    float* cur = (float*)malloc(sizeof(float) * (size_t)N * P);
    float* next = (float*)malloc(sizeof(float) * (size_t)N * P);
    memcpy(cur, x0, sizeof(float) * (size_t)N * P);
    for (int level = 0; level < num_levels; ++level) {
        int D = feat_dims[level];
        float* A = (float*)malloc(sizeof(float) * (size_t)N * D);
        orc_project_all(cur, templates, N, P, D, level, h, user, A);
        orc_apply_level(A, weights[level], cur, N, D, P, nm, next);
        float* t = cur; cur = next; next = t;
        if (cb) cb(cur, N, P, level, cb_user);                 /* :303 */
        free(A);
    }
    memcpy(x_final, cur, sizeof(float) * (size_t)N * P);
    free(cur); free(next);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Model file: cereal::BinaryOutputArchive, little-endian, no headers.       */
/*   detection_model::serialize            model.hpp:178-182                 */
/*   SupervisedDescentOptimiser::serialize superviseddescent.hpp:356-360     */
/*   LinearRegressor::serialize            regressors.hpp:395-399            */
/*   Regulariser::serialize                regressors.hpp:164-168            */
/*   cv::Mat save/load                     utils/mat_cerealisation.hpp:42-99 */
/*   InterEyeDistanceNormalisation         model.hpp:111-115                 */
/*   HoGParam::serialize                   adaptive_vlhog.hpp:55-59          */
/* ------------------------------------------------------------------------- */

typedef struct { FILE* f; int ok; } orc_rd;
static void rd_bytes(orc_rd* r, void* dst, size_t n) { if (r->ok && fread(dst, 1, n, r->f) != n) r->ok = 0; }
static int32_t rd_i32(orc_rd* r) { int32_t v = 0; rd_bytes(r, &v, 4); return v; }
static uint64_t rd_u64(orc_rd* r) { uint64_t v = 0; rd_bytes(r, &v, 8); return v; }
static uint8_t rd_u8(orc_rd* r) { uint8_t v = 0; rd_bytes(r, &v, 1); return v; }
static float rd_f32(orc_rd* r) { float v = 0; rd_bytes(r, &v, 4); return v; }
static char** rd_strvec(orc_rd* r, int32_t* count)
{
    uint64_t n = rd_u64(r);
    if (!r->ok || n > (1u << 20)) { r->ok = 0; *count = 0; return NULL; }
    char** v = (char**)calloc(n ? n : 1, sizeof(char*));
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t len = rd_u64(r);
        if (!r->ok || len > (1u << 20)) { r->ok = 0; len = 0; }
        v[i] = (char*)calloc(len + 1, 1);
        rd_bytes(r, v[i], len);
    }
    *count = (int32_t)n;
    return v;
}
static void free_strvec(char** v, int n) { if (!v) return; for (int i = 0; i < n; ++i) free(v[i]); free(v); }
static float* rd_mat(orc_rd* r, int32_t* rows, int32_t* cols)
{
    *rows = rd_i32(r); *cols = rd_i32(r);
    int32_t type = rd_i32(r);
    uint8_t cont = rd_u8(r);
    (void)cont;   /* row-by-row or in one block: identical bytes for a packed float matrix */
    if (!r->ok || type != 5 /* CV_32FC1 */ || *rows < 0 || *cols < 0) { r->ok = 0; return NULL; }
    size_t n = (size_t)(*rows) * (size_t)(*cols);
    float* d = (float*)malloc(sizeof(float) * (n ? n : 1));
    rd_bytes(r, d, n * 4);
    return d;
}

static int find_id(char** ids, int n, const char* s)
{
    for (int i = 0; i < n; ++i) if (strcmp(ids[i], s) == 0) return i;
    return -1;
}

orc_model* orc_model_load(const char* path, char* err, int errlen)
{
#define FAIL(msg) do { if (err) snprintf(err, errlen, "%s", msg); if (m) orc_model_free(m); if (r.f) fclose(r.f); return NULL; } while (0)
    orc_model* m = NULL;
    orc_rd r; r.f = fopen(path, "rb"); r.ok = 1;
    if (!r.f) FAIL("The given model file could not be opened");   /* model.hpp:199 */
    m = (orc_model*)calloc(1, sizeof(orc_model));
    uint64_t nreg = rd_u64(&r);
    if (!r.ok || nreg == 0 || nreg > 64) FAIL("bad regressor count");
    m->num_levels = (int32_t)nreg;
    m->rows = (int32_t*)calloc(nreg, 4); m->cols = (int32_t*)calloc(nreg, 4);
    m->weights = (float**)calloc(nreg, sizeof(float*));
    m->regularisers = (orc_regulariser*)calloc(nreg, sizeof(orc_regulariser));
    for (uint64_t i = 0; i < nreg; ++i) {
        m->weights[i] = rd_mat(&r, &m->rows[i], &m->cols[i]);
        m->regularisers[i].type = rd_i32(&r);
        m->regularisers[i].lambda = rd_f32(&r);
        m->regularisers[i].regularise_last_row = rd_u8(&r);
        if (!r.ok) FAIL("truncated regressor");
    }
    /* normaliser: three string vectors (its own copies) */
    int32_t n_lm2 = 0, n_r2 = 0, n_l2 = 0;
    char** lm2 = rd_strvec(&r, &n_lm2);
    char** r2 = rd_strvec(&r, &n_r2);
    char** l2 = rd_strvec(&r, &n_l2);
    int32_t mr = 0, mc = 0;
    m->mean = rd_mat(&r, &mr, &mc);
    int32_t n_ids = 0;
    m->landmark_ids = rd_strvec(&r, &n_ids);
    m->num_landmarks = n_ids;
    uint64_t nhog = rd_u64(&r);
    if (!r.ok || nhog != nreg) { free_strvec(lm2, n_lm2); free_strvec(r2, n_r2); free_strvec(l2, n_l2); FAIL("hog_params count != regressor count"); }
    m->hog_params = (orc_hog_param*)calloc(nhog, sizeof(orc_hog_param));
    for (uint64_t i = 0; i < nhog; ++i) {
        m->hog_params[i].variant = rd_i32(&r);
        m->hog_params[i].num_cells = rd_i32(&r);
        m->hog_params[i].cell_size = rd_i32(&r);
        m->hog_params[i].num_bins = rd_i32(&r);
        m->hog_params[i].relative_patch_size = rd_f32(&r);
    }
    m->right_ids = rd_strvec(&r, &m->n_right);
    m->left_ids = rd_strvec(&r, &m->n_left);
    int trailing = (r.ok && fgetc(r.f) != EOF);
    int consistent = r.ok && !trailing && mr == 1 && mc == 2 * n_ids && n_lm2 == n_ids && n_r2 == m->n_right && n_l2 == m->n_left;
    for (int i = 0; consistent && i < n_ids; ++i) consistent = strcmp(lm2[i], m->landmark_ids[i]) == 0;
    for (int i = 0; consistent && i < m->n_right; ++i) consistent = strcmp(r2[i], m->right_ids[i]) == 0;
    for (int i = 0; consistent && i < m->n_left; ++i) consistent = strcmp(l2[i], m->left_ids[i]) == 0;
    free_strvec(lm2, n_lm2); free_strvec(r2, n_r2); free_strvec(l2, n_l2);
    if (!consistent) FAIL("model file inconsistent or has trailing bytes");
    m->right_idx = (int32_t*)calloc(m->n_right ? m->n_right : 1, 4);
    m->left_idx = (int32_t*)calloc(m->n_left ? m->n_left : 1, 4);
    for (int i = 0; i < m->n_right; ++i) {
        m->right_idx[i] = find_id(m->landmark_ids, n_ids, m->right_ids[i]);
        if (m->right_idx[i] < 0) FAIL("one of given rightEyeIdentifiers ids not present in lms");   /* helpers.hpp:144 */
    }
    for (int i = 0; i < m->n_left; ++i) {
        m->left_idx[i] = find_id(m->landmark_ids, n_ids, m->left_ids[i]);
        if (m->left_idx[i] < 0) FAIL("one of given leftEyeIdentifiers ids not present in lms");     /* helpers.hpp:153 */
    }
    fclose(r.f);
    return m;
#undef FAIL
}

static void wr_strvec(FILE* f, char** v, int n)
{
    uint64_t c = (uint64_t)n; fwrite(&c, 8, 1, f);
    for (int i = 0; i < n; ++i) { uint64_t len = strlen(v[i]); fwrite(&len, 8, 1, f); fwrite(v[i], 1, len, f); }
}
static void wr_mat(FILE* f, const float* d, int32_t rows, int32_t cols)
{
    int32_t type = 5; uint8_t cont = 1;
    fwrite(&rows, 4, 1, f); fwrite(&cols, 4, 1, f); fwrite(&type, 4, 1, f); fwrite(&cont, 1, 1, f);
    fwrite(d, 4, (size_t)rows * cols, f);
}

int orc_model_save(const orc_model* m, const char* path)
{
    FILE* f = fopen(path, "wb");
    if (!f) return 1;
    uint64_t n = (uint64_t)m->num_levels; fwrite(&n, 8, 1, f);
    for (int i = 0; i < m->num_levels; ++i) {
        wr_mat(f, m->weights[i], m->rows[i], m->cols[i]);
        fwrite(&m->regularisers[i].type, 4, 1, f);
        fwrite(&m->regularisers[i].lambda, 4, 1, f);
        uint8_t b = (uint8_t)(m->regularisers[i].regularise_last_row != 0); fwrite(&b, 1, 1, f);
    }
    wr_strvec(f, m->landmark_ids, m->num_landmarks);
    wr_strvec(f, m->right_ids, m->n_right);
    wr_strvec(f, m->left_ids, m->n_left);
    wr_mat(f, m->mean, 1, 2 * m->num_landmarks);
    wr_strvec(f, m->landmark_ids, m->num_landmarks);
    fwrite(&n, 8, 1, f);
    for (int i = 0; i < m->num_levels; ++i) {
        fwrite(&m->hog_params[i].variant, 4, 1, f);
        fwrite(&m->hog_params[i].num_cells, 4, 1, f);
        fwrite(&m->hog_params[i].cell_size, 4, 1, f);
        fwrite(&m->hog_params[i].num_bins, 4, 1, f);
        fwrite(&m->hog_params[i].relative_patch_size, 4, 1, f);
    }
    wr_strvec(f, m->right_ids, m->n_right);
    wr_strvec(f, m->left_ids, m->n_left);
    fclose(f);
    return 0;
}

void orc_model_free(orc_model* m)
{
    if (!m) return;
    if (m->weights) for (int i = 0; i < m->num_levels; ++i) free(m->weights[i]);
    free(m->weights); free(m->rows); free(m->cols); free(m->regularisers); free(m->mean);
    free_strvec(m->landmark_ids, m->num_landmarks);
    free(m->hog_params);
    free_strvec(m->right_ids, m->n_right); free_strvec(m->left_ids, m->n_left);
    free(m->right_idx); free(m->left_idx);
    free(m);
}

/* predict(), superviseddescent.hpp:323-344 with HogTransform as projection */
int orc_detect_init(const orc_model* m, const uint8_t* image, int w, int h, int stride,
                    const float* init, orc_hog_core_fn hog_core, float* landmarks)
{
    const int L = m->num_landmarks, P = 2 * L;
    float cur[1024], upd[1024];
    if (P > 1024) return 2;
    memcpy(cur, init, sizeof(float) * P);
    for (int level = 0; level < m->num_levels; ++level) {
        int D = m->rows[level];
        if (D != orc_feature_length(L, &m->hog_params[level]) || m->cols[level] != P) return 3;
        float* feat = (float*)malloc(sizeof(float) * D);
        int rc = orc_hog_transform(image, w, h, stride, cur, L, &m->hog_params[level], m->right_idx, m->n_right,
                                   m->left_idx, m->n_left, hog_core, feat);
        if (rc) { free(feat); return rc; }
        orc_predict(feat, 1, D, m->weights[level], P, upd);           /* :336 */
        double ied = orc_get_ied(cur, L, m->right_idx, m->n_right, m->left_idx, m->n_left);
        float n, inv_n;
        orc_ied_normaliser(ied, &n, &inv_n);
        for (int j = 0; j < P; ++j) cur[j] = cur[j] - upd[j] * inv_n;  /* :338-339 */
        free(feat);
    }
    memcpy(landmarks, cur, sizeof(float) * P);
    return 0;
}

int orc_detect(const orc_model* m, const uint8_t* image, int w, int h, int stride,
               int bx, int by, int bw, int bh, orc_hog_core_fn hog_core, float* landmarks)
{
    float init[1024];
    if (2 * m->num_landmarks > 1024) return 2;
    orc_align_mean(m->mean, m->num_landmarks, bx, by, bw, bh, 1.0f, 1.0f, 0.0f, 0.0f, init);  /* model.hpp:135 */
    return orc_detect_init(m, image, w, h, stride, init, hog_core, landmarks);
}

int orc_detect_batch(const orc_model* m, const uint8_t* images, int count, int w, int h, int stride,
                     const int32_t* boxes, orc_hog_core_fn hog_core, int threads, float* landmarks)
{
    int rc = 0;
    const int P = 2 * m->num_landmarks;
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (int i = 0; i < count; ++i) {
        int r = orc_detect(m, images + (size_t)i * h * stride, w, h, stride, boxes[4 * i], boxes[4 * i + 1],
                           boxes[4 * i + 2], boxes[4 * i + 3], hog_core, landmarks + (size_t)i * P);
        if (r) {
#pragma omp atomic write
            rc = r;
        }
    }
    return rc;
}
