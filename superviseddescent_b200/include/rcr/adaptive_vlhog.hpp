// B200 drop-in for include/rcr/adaptive_vlhog.hpp: HoGParam (:41-60) and the projection functor
// HogTransform (:70-195).  The functor keeps the reference's constructor and call signature
//     cv::Mat operator()(cv::Mat parameters, size_t regressorLevel, int trainingIndex = 0)
// (one sample, used by predict(), superviseddescent.hpp:332) and adds project_device(), which the
// optimiser uses to extract the features of ALL samples of a level with one kernel launch.  The images are
// uploaded to HBM once, on first use; crop / resize / HOG run in sd_hog_batch (sm_100a).
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "rcr/helpers.hpp"
#include "sd_b200/device.hpp"

typedef enum { VlHogVariantDalalTriggs = 0, VlHogVariantUoctti = 1 } VlHogVariant;   // hog.h:70-72

namespace rcr {

struct HoGParam {
    VlHogVariant vlhog_variant;
    int num_cells;
    int cell_size;
    int num_bins;
    float relative_patch_size;
    sd_hog_param c() const { sd_hog_param p; p.variant = vlhog_variant; p.num_cells = num_cells; p.cell_size = cell_size; p.num_bins = num_bins; p.relative_patch_size = relative_patch_size; return p; }
};

class HogTransform {
public:
    // Do not call with `images` that are temporaries (the reference holds a const&, adaptive_vlhog.hpp:188).
    HogTransform(const std::vector<cv::Mat>& images, std::vector<HoGParam> hog_params, std::vector<std::string> modelLandmarksList,
                 std::vector<std::string> rightEyeIdentifiers, std::vector<std::string> leftEyeIdentifiers)
        : images(images), hog_params(hog_params), modelLandmarksList(modelLandmarksList), rightEyeIdentifiers(rightEyeIdentifiers),
          leftEyeIdentifiers(leftEyeIdentifiers), dev(std::make_shared<DeviceImages>()) {}

    int feature_length(size_t level) const
    {
        const sd_hog_param p = hog_params[level].c();
        return sd_hog_feature_length(static_cast<int>(modelLandmarksList.size()), &p);
    }

    // Features of ONE sample (adaptive_vlhog.hpp:109-185)
    cv::Mat operator()(cv::Mat parameters, size_t regressorLevel, int trainingIndex = 0)
    {
        sd_ctx* ctx = sd_b200::context();
        ensure_uploaded();
        const int D = feature_length(regressorLevel);
        sd_b200::DeviceBuffer dx, dA(static_cast<size_t>(D) * sizeof(float)), didx(sizeof(int32_t));
        sd_b200::upload(parameters, dx, parameters.cols);
        const int32_t idx = trainingIndex;
        sd_b200::check(ctx, sd_memcpy_h2d(ctx, didx.as<int32_t>(), &idx, sizeof(idx)), "HogTransform");
        launch(dx.as<float>(), parameters.cols, 1, regressorLevel, dA.as<float>(), D, didx.as<int32_t>());
        return sd_b200::download(dA.as<float>(), 1, D, D);
    }

    // Features of samples 0..n-1 (sample i reads image i), straight into a device matrix with row stride ld.
    void project_device(const float* d_x, int64_t ldx, int n, size_t level, float* d_A, int64_t ld)
    {
        ensure_uploaded();
        launch(d_x, ldx, n, level, d_A, ld, nullptr);
    }

    sd_normalisation eyes() const
    {
        sd_normalisation nrm{};
        nrm.kind = 1;
        const auto r = eye_indices(modelLandmarksList, rightEyeIdentifiers, "right");
        const auto l = eye_indices(modelLandmarksList, leftEyeIdentifiers, "left");
        if (r.empty() || l.empty() || r.size() > 4 || l.size() > 4) throw std::runtime_error("HogTransform: 1..4 eye identifiers per eye are supported");
        nrm.n_right = static_cast<int>(r.size());
        nrm.n_left = static_cast<int>(l.size());
        for (size_t i = 0; i < r.size(); ++i) nrm.right_idx[i] = r[i];
        for (size_t i = 0; i < l.size(); ++i) nrm.left_idx[i] = l[i];
        return nrm;
    }

private:
    struct DeviceImages {
        sd_b200::DeviceBuffer buf, frames;
        sd_image_batch batch{};
        bool ready = false;
    };

    void ensure_uploaded()
    {
        if (dev->ready) return;
        if (images.empty()) throw std::runtime_error("HogTransform: no images");
        sd_ctx* ctx = sd_b200::context();
        // frames may differ in size (the reference's std::vector<cv::Mat>): packed back to back with 16-byte aligned rows and one
        // sd_frame descriptor each; equally sized frames take the plain strided layout (and the TMA route of the kernel)
        bool same = true;
        for (const cv::Mat& im : images) same = same && im.cols == images[0].cols && im.rows == images[0].rows;
        std::vector<sd_frame> frames(images.size());
        size_t total = 0;
        for (size_t i = 0; i < images.size(); ++i) {
            const int w = images[i].cols, h = images[i].rows;
            const int stride = same ? w : (w + 15) / 16 * 16;
            frames[i].width = w; frames[i].height = h; frames[i].row_stride = stride; frames[i].reserved = 0;
            frames[i].offset = static_cast<int64_t>(total);
            total += static_cast<size_t>(stride) * h;
        }
        dev->buf.allocate(total);
        sd_b200::check(ctx, sd_memset(ctx, dev->buf.as<unsigned char>(), 0, total), "HogTransform upload");
        sd_b200::DeviceBuffer bgr;                 // staging for one colour frame
        for (size_t i = 0; i < images.size(); ++i) {
            const cv::Mat& im = images[i];
            const int w = im.cols, h = im.rows;
            unsigned char* d_frame = dev->buf.as<unsigned char>() + frames[i].offset;
            if (im.channels() == 3) {
                // cv::cvtColor(BGR2GRAY), adaptive_vlhog.hpp:115-117: on the device, once per frame (sd_bgr2gray)
                const size_t frame = static_cast<size_t>(w) * h;
                bgr.allocate(3 * frame);
                sd_b200::check(ctx, sd_memcpy2d_h2d(ctx, bgr.as<unsigned char>(), 3 * static_cast<size_t>(w), im.ptr<unsigned char>(0), im.step(), 3 * static_cast<size_t>(w), h), "HogTransform upload");
                sd_b200::check(ctx, sd_bgr2gray(ctx, bgr.as<unsigned char>(), w, h, 3 * static_cast<int64_t>(w), 3 * static_cast<int64_t>(frame), 1,
                                                d_frame, frames[i].row_stride, static_cast<int64_t>(frames[i].row_stride) * h), "sd_bgr2gray");
            } else {
                sd_b200::check(ctx, sd_memcpy2d_h2d(ctx, d_frame, frames[i].row_stride, im.ptr<unsigned char>(0), im.step(), w, h), "HogTransform upload");
            }
        }
        dev->batch = sd_image_batch{};
        dev->batch.d_data = dev->buf.as<unsigned char>();
        dev->batch.count = static_cast<int32_t>(images.size());
        if (same) {
            dev->batch.width = images[0].cols; dev->batch.height = images[0].rows; dev->batch.row_stride = images[0].cols;
            dev->batch.image_stride = static_cast<int64_t>(images[0].cols) * images[0].rows;
        } else {
            dev->frames.allocate(frames.size() * sizeof(sd_frame));
            sd_b200::check(ctx, sd_memcpy_h2d(ctx, dev->frames.as<sd_frame>(), frames.data(), frames.size() * sizeof(sd_frame)), "HogTransform upload");
            dev->batch.d_frames = dev->frames.as<sd_frame>();
        }
        sd_b200::check(ctx, sd_sync(ctx), "HogTransform upload");
        dev->ready = true;
    }

    void launch(const float* d_x, int64_t ldx, int n, size_t level, float* d_A, int64_t ld, const int32_t* d_index)
    {
        sd_ctx* ctx = sd_b200::context();
        const sd_normalisation nrm = eyes();
        const sd_hog_param p = hog_params[level].c();
        sd_b200::check(ctx, sd_hog_batch(ctx, &dev->batch, d_index, d_x, ldx, n, static_cast<int>(modelLandmarksList.size()), &nrm, &p, d_A, ld), "sd_hog_batch");
    }

    const std::vector<cv::Mat>& images;
    std::vector<HoGParam> hog_params;
    std::vector<std::string> modelLandmarksList;
    std::vector<std::string> rightEyeIdentifiers;
    std::vector<std::string> leftEyeIdentifiers;
    std::shared_ptr<DeviceImages> dev;   // shared between the copies the optimiser makes of this functor
};

}  // namespace rcr
