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
        sd_b200::DeviceBuffer buf;
        sd_image_batch batch{};
        bool ready = false;
    };

    void ensure_uploaded()
    {
        if (dev->ready) return;
        if (images.empty()) throw std::runtime_error("HogTransform: no images");
        const int w = images[0].cols, h = images[0].rows;
        const size_t frame = static_cast<size_t>(w) * h;
        dev->buf.allocate(frame * images.size());
        sd_ctx* ctx = sd_b200::context();
        sd_b200::DeviceBuffer bgr;                 // staging for one colour frame
        for (size_t i = 0; i < images.size(); ++i) {
            const cv::Mat& im = images[i];
            if (im.cols != w || im.rows != h) throw std::runtime_error("HogTransform: the batched device path needs equally sized images");
            unsigned char* d_frame = dev->buf.as<unsigned char>() + i * frame;
            if (im.channels() == 3) {
                // cv::cvtColor(BGR2GRAY), adaptive_vlhog.hpp:115-117: on the device, once per frame (sd_bgr2gray)
                bgr.allocate(3 * frame);
                for (int y = 0; y < h; ++y)
                    sd_b200::check(ctx, sd_memcpy_h2d(ctx, bgr.as<unsigned char>() + static_cast<size_t>(y) * 3 * w, im.ptr<unsigned char>(y), 3 * static_cast<size_t>(w)), "HogTransform upload");
                sd_b200::check(ctx, sd_bgr2gray(ctx, bgr.as<unsigned char>(), w, h, 3 * static_cast<int64_t>(w), 3 * static_cast<int64_t>(frame), 1,
                                                d_frame, w, static_cast<int64_t>(frame)), "sd_bgr2gray");
            } else {
                for (int y = 0; y < h; ++y)
                    sd_b200::check(ctx, sd_memcpy_h2d(ctx, d_frame + static_cast<size_t>(y) * w, im.ptr<unsigned char>(y), w), "HogTransform upload");
            }
        }
        sd_b200::check(ctx, sd_sync(ctx), "HogTransform upload");
        dev->batch.d_data = dev->buf.as<unsigned char>();
        dev->batch.width = w; dev->batch.height = h; dev->batch.row_stride = w;
        dev->batch.image_stride = static_cast<int64_t>(frame);
        dev->batch.count = static_cast<int32_t>(images.size());
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
