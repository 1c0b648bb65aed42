/*
 * sd_oracle.h -- CPU restatement of the superviseddescent / RCR hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by
 * or executed from the product (superviseddescent_b200/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use it, and only as the checker / reported baseline.
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose arithmetic it restates.  Pinning status (see DESIGN.md "Oracle"):
 *   - HOG core            : pinned bit-exactly against the reference's own
 *                           hog.c compiled verbatim (oracle/_ref).
 *   - 8-bit resize        : pinned bit-exactly against cv2 4.13 (goldens in
 *                           tests/golden/, generator tests/golden/gen_golden.py).
 *   - regressor / cascade : pinned against the literals of the reference's
 *                           gtest suite (tests/test_LinearRegressor*.cpp,
 *                           tests/test_SupervisedDescentOptimiser.cpp).
 *   - model file / detect : pinned by parsing the shipped .bin byte-for-byte
 *                           and by the landmark error on the 5 annotated
 *                           example images.
 */
#ifndef SD_ORACLE_H
#define SD_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* rcr::HoGParam, include/rcr/adaptive_vlhog.hpp:41-60 (same field order as the
 * cereal archive: variant, num_cells, cell_size, num_bins, relative_patch_size) */
typedef struct {
    int32_t variant;    /* 0 = DalalTriggs, 1 = UoCTTI (hog.h:70) */
    int32_t num_cells;
    int32_t cell_size;
    int32_t num_bins;
    float relative_patch_size;
} orc_hog_param;

/* superviseddescent::Regulariser, regressors.hpp:87-169 */
typedef struct {
    int32_t type;       /* 0 = Manual, 1 = MatrixNorm */
    float lambda;       /* lambda or the factor for MatrixNorm */
    int32_t regularise_last_row;
} orc_regulariser;

/* pluggable HOG core: (float image w*h, cell, K, variant) -> planar features */
typedef void (*orc_hog_core_fn)(const float* image, int width, int height, int cell_size,
                                int num_orientations, int variant, float* out);

/* ---- HOG (include/rcr/hog.c) ------------------------------------------------ */
int orc_hog_dimension(int variant, int num_orientations);             /* hog.c:212-223 */
void orc_hog_core(const float* image, int width, int height, int cell_size,
                  int num_orientations, int variant, float* out);       /* hog.c:174-204,595-728,857-1062 */
/* per-pixel orientation arg-max only (integer result), for bit-exact parity checks */
void orc_hog_orientation_bins(const float* image, int width, int height, int num_orientations,
                              int32_t* bins /* w*h, -1 on the border */);

/* ---- OpenCV arithmetic on the path (un-vendored dependency, pinned vs cv2 4.13) */
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                          uint8_t* dst, int dw, int dh, int dstride);  /* cv::resize, adaptive_vlhog.hpp:155 */
int orc_cv_round(float v);                                              /* cvRound, adaptive_vlhog.hpp:132-133 */
void orc_bgr2gray_u8(const uint8_t* bgr, int w, int h, int sstride, uint8_t* gray, int dstride); /* :116 */

/* ---- RCR glue --------------------------------------------------------------- */
double orc_get_ied(const float* row, int num_landmarks, const int32_t* right_idx, int n_right,
                   const int32_t* left_idx, int n_left);                /* helpers.hpp:136-160 */
int orc_patch_half(float relative_patch_size, double ied);              /* adaptive_vlhog.hpp:123 */
void orc_crop_patch_u8(const uint8_t* image, int w, int h, int stride, int cx, int cy, int half,
                       uint8_t* patch /* (2*half)^2 */);               /* adaptive_vlhog.hpp:135-151 */
int orc_feature_length(int num_landmarks, const orc_hog_param* p);      /* D = L*nc*nc*dd + 1 */
/* HogTransform::operator(), adaptive_vlhog.hpp:109-185.  image is 8UC1. hog_core==NULL -> orc_hog_core */
int orc_hog_transform(const uint8_t* image, int w, int h, int stride, const float* params,
                      int num_landmarks, const orc_hog_param* p, const int32_t* right_idx, int n_right,
                      const int32_t* left_idx, int n_left, orc_hog_core_fn hog_core, float* out_row);
/* apps/rcr/rcr-train.cpp:130-146 and :149-212 (training front-end helpers) */
void orc_perturb_box(const int32_t box[4], float tx, float ty, float scaling, int32_t out[4]);
void orc_normalised_landmark_errors(const float* pred, const float* gt, int N, int L, const int32_t* ridx, int nr,
                                    const int32_t* lidx, int nl, float* out);
/* debug/parity taps: geometry (cx, cy, half) per landmark and the resized u8 patch of one landmark */
void orc_patch_geometry(const float* params, int num_landmarks, const orc_hog_param* p,
                        const int32_t* right_idx, int n_right, const int32_t* left_idx, int n_left,
                        int32_t* cx, int32_t* cy, int32_t* half);
void orc_align_mean(const float* mean, int num_landmarks, int box_x, int box_y, int box_w, int box_h,
                    float sx, float sy, float tx, float ty, float* out);  /* model.hpp:64-76 */
void orc_ied_normaliser(double ied, float* norm /*1/IED as float*/, float* inv_norm /*1/(1/IED)*/); /* model.hpp:94-98, superviseddescent.hpp:213 */

/* ---- regressor (regressors.hpp) --------------------------------------------- */
/* lambda actually applied, regressors.hpp:126-148 */
float orc_regulariser_lambda(const orc_regulariser* r, const float* AtA, int D, int num_training_elements);
/* PartialPivLUSolver::solve, regressors.hpp:199-234.  A: N x D, B: N x M row-major; X: D x M.
 * precision 0 = float32 throughout (as Eigen does), 1 = float64 accumulation (truth for error budgets).
 * Returns 0, or 1 if a zero pivot was met (the reference would silently return inf/nan). */
int orc_solve(const float* A, const float* B, int N, int D, int M, const orc_regulariser* r,
              int precision, float* X, float* lambda_out);
void orc_gram(const float* A, int N, int D, int precision, float* AtA);
void orc_predict(const float* values, int N, int D, const float* X, int M, float* out); /* :377-381 */
double orc_test_residual(const float* data, const float* labels, int N, int D, const float* X, int M); /* :361-369 */

/* ---- cascade (superviseddescent.hpp) ---------------------------------------- */
/* projection callback h(x_row, level, sample_idx) -> feature row of length D(level) */
typedef void (*orc_projection_fn)(const float* x_row, int P, int level, int sample_idx,
                                  float* out, void* user);
/* normalisation: 0 = NoNormalisation (superviseddescent.hpp:60-74), 1 = InterEyeDistance (model.hpp:84-116) */
typedef struct {
    int32_t kind;
    const int32_t* right_idx; int32_t n_right;
    const int32_t* left_idx; int32_t n_left;
} orc_normalisation;
typedef void (*orc_epoch_cb)(const float* current_x, int N, int P, int level, void* user);

/* train(), superviseddescent.hpp:165-219.  x_gt, x0: N x P.  templates: NULL or N x D.
 * feat_dims[level] gives D per level.  weights[level] receives D x P (caller-allocated). */
int orc_cascade_train(const float* x_gt, const float* x0, const float* templates, int N, int P,
                      int num_levels, const int* feat_dims, const orc_regulariser* regs,
                      const orc_normalisation* norm, orc_projection_fn h, void* user,
                      int precision, float** weights, float* x_final, orc_epoch_cb cb, void* cb_user);
/* test()/predict(), superviseddescent.hpp:262-344 */
int orc_cascade_apply(const float* x0, const float* templates, int N, int P, int num_levels,
                      const int* feat_dims, float* const* weights, const orc_normalisation* norm,
                      orc_projection_fn h, void* user, float* x_final, orc_epoch_cb cb, void* cb_user);

/* ---- model file (cereal binary; model.hpp:178-219 and friends) --------------- */
typedef struct {
    int32_t num_levels;
    int32_t num_landmarks;
    int32_t* rows; int32_t* cols;          /* per level */
    float** weights;                        /* per level rows*cols */
    orc_regulariser* regularisers;          /* per level */
    float* mean;                            /* 2L */
    char** landmark_ids;                    /* L strings */
    orc_hog_param* hog_params;              /* per level */
    int32_t n_right, n_left;
    int32_t* right_idx; int32_t* left_idx;  /* row indices of the eye landmarks */
    char** right_ids; char** left_ids;
    /* normaliser's own copies (model.hpp:111-115) are validated to equal the model's */
} orc_model;

orc_model* orc_model_load(const char* path, char* err, int errlen);
int orc_model_save(const orc_model* m, const char* path);
void orc_model_free(orc_model* m);
/* detection_model::detect(image, facebox), model.hpp:132-144.  8UC1 image. */
int orc_detect(const orc_model* m, const uint8_t* image, int w, int h, int stride,
               int box_x, int box_y, int box_w, int box_h, orc_hog_core_fn hog_core, float* landmarks);
/* detection_model::detect(image, initialisation), model.hpp:147-157 */
int orc_detect_init(const orc_model* m, const uint8_t* image, int w, int h, int stride,
                    const float* init, orc_hog_core_fn hog_core, float* landmarks);
/* batch of independent detects over images of identical size, `threads` OpenMP threads (1 = the
 * reference-faithful sequential predict; >1 = one face per thread) */
int orc_detect_batch(const orc_model* m, const uint8_t* images, int count, int w, int h, int stride,
                     const int32_t* boxes /* count x 4 */, orc_hog_core_fn hog_core, int threads,
                     float* landmarks /* count x 2L */);
/* batched HogTransform over many samples (one image per sample), `threads` OpenMP threads */
/* examples/landmark_detection.cpp:195-261: fixed patch (half = num_cells * (cell_size / 2)), no resize, no bias */
int orc_hog_transform_fixed(const uint8_t* image, int w, int h, int stride, const float* params, int L,
                            const orc_hog_param* p, orc_hog_core_fn hog_core, float* out_row, int* out_len);
int orc_hog_transform_batch(const uint8_t* images, int count, int w, int h, int stride,
                            const float* params, int num_landmarks, const orc_hog_param* p,
                            const int32_t* right_idx, int n_right, const int32_t* left_idx, int n_left,
                            orc_hog_core_fn hog_core, int threads, float* out, int out_ld);

#ifdef __cplusplus
}
#endif
#endif /* SD_ORACLE_H */
