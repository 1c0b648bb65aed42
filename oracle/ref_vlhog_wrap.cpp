// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Thin C-ABI wrapper around the reference's OWN HOG implementation
// (/root/reference/include/rcr/hog.h + hog.c, compiled where they lie, see
// oracle/Makefile).  No reference source is copied into this repository: this
// file only #includes the header from the read-only reference tree and exposes
// one entry point that runs vl_hog_new / vl_hog_put_image / vl_hog_extract /
// vl_hog_delete exactly as rcr::HogTransform does
// (reference: include/rcr/adaptive_vlhog.hpp:158-165).
//
// The built artefact goes to oracle/_ref/libref_vlhog.so (git-ignored, but it
// travels to the GPU box with the gpurun snapshot).
// hog.h declares its prototypes outside, and #includes hog.c inside, an extern "C" block;
// g++ needs both in the same linkage, so the whole header is wrapped.
extern "C" {
#include "hog.h"  // -I/root/reference/include/rcr
}

extern "C" {

// Runs the reference HOG on a float image (values 0..255), one channel.
// out must hold dims[0]*dims[1]*dims[2] floats; layout is VLFeat's planar
// [dimension][hogHeight][hogWidth] (x fastest).  Returns 0 on success.
int ref_vl_hog(int variant, int num_orientations, const float* image, int width, int height,
               int cell_size, float* out, int* dims)
{
    VlHog* hog = vl_hog_new(variant == 0 ? VlHogVariantDalalTriggs : VlHogVariantUoctti,
                            (vl_size)num_orientations, VL_FALSE);
    if (!hog) return 1;
    vl_hog_put_image(hog, image, (vl_size)width, (vl_size)height, 1, (vl_size)cell_size);
    const int ww = (int)vl_hog_get_width(hog);
    const int hh = (int)vl_hog_get_height(hog);
    const int dd = (int)vl_hog_get_dimension(hog);
    if (dims) { dims[0] = ww; dims[1] = hh; dims[2] = dd; }
    if (out) vl_hog_extract(hog, out);
    vl_hog_delete(hog);
    return 0;
}

// Same signature as the oracle's pluggable "hog core" callback
// (oracle/sd_oracle.h: orc_hog_core_fn) so that the CPU baseline can run the
// reference's hog.c inside the restated HogTransform glue.
void ref_hog_core(const float* image, int width, int height, int cell_size, int num_orientations,
                  int variant, float* out)
{
    ref_vl_hog(variant, num_orientations, image, width, height, cell_size, out, 0);
}

}  // extern "C"
