// Glue between the C++14 header shells and the C ABI (include/sd_b200.h): one lazily created context
// per host thread, RAII device buffers, and translation of status codes into the exception types the
// reference throws (std::runtime_error, SURVEY.md 8b "Errors").
#pragma once

#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>

#include "sd_b200.h"
#include "sd_b200/mat.hpp"

namespace sd_b200 {

inline void check(sd_ctx* ctx, int rc, const char* what)
{
    if (rc == SD_OK) return;
    std::string msg = ctx ? sd_last_error(ctx) : "no CUDA device / context (the B200 engine has no CPU fallback)";
    throw std::runtime_error(std::string(what) + ": " + msg);
}

// One context per host thread (re-entrant on distinct contexts, SURVEY 8b "Threading").
// The device ordinal comes from SD_B200_DEVICE (default 0).
inline sd_ctx* context()
{
    struct Holder {
        sd_ctx* ctx = nullptr;
        Holder()
        {
            const char* env = std::getenv("SD_B200_DEVICE");
            const int dev = env ? std::atoi(env) : 0;
            const int rc = sd_ctx_create(dev, nullptr, &ctx);
            if (rc != SD_OK) throw std::runtime_error("sd_ctx_create failed: no usable CUDA device (the B200 engine has no CPU fallback)");
        }
        ~Holder() { sd_ctx_destroy(ctx); }
    };
    static thread_local Holder holder;
    return holder.ctx;
}

class DeviceBuffer {
public:
    DeviceBuffer() = default;
    explicit DeviceBuffer(size_t bytes) { allocate(bytes); }
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    DeviceBuffer(DeviceBuffer&& o) noexcept : ptr_(o.ptr_), bytes_(o.bytes_) { o.ptr_ = nullptr; o.bytes_ = 0; }
    DeviceBuffer& operator=(DeviceBuffer&& o) noexcept { std::swap(ptr_, o.ptr_); std::swap(bytes_, o.bytes_); return *this; }
    ~DeviceBuffer() { if (ptr_) sd_free(context(), ptr_); }
    void allocate(size_t bytes)
    {
        if (bytes <= bytes_) return;
        if (ptr_) { sd_free(context(), ptr_); ptr_ = nullptr; }
        check(context(), sd_malloc(context(), bytes, &ptr_), "sd_malloc");
        bytes_ = bytes;
    }
    template <class T> T* as() const { return static_cast<T*>(ptr_); }
    size_t bytes() const { return bytes_; }

private:
    void* ptr_ = nullptr;
    size_t bytes_ = 0;
};

// packed float32 cv::Mat -> device (row stride ld floats, ld >= cols)
inline void upload(const cv::Mat& m, DeviceBuffer& dst, int64_t ld)
{
    dst.allocate(static_cast<size_t>(m.rows) * ld * sizeof(float));
    sd_ctx* ctx = context();
    if (m.isContinuous() && ld == m.cols) {
        check(ctx, sd_memcpy_h2d(ctx, dst.as<float>(), m.ptr<float>(0), static_cast<size_t>(m.rows) * m.cols * sizeof(float)), "upload");
    } else {   // strided on either side: one 2-D copy
        check(ctx, sd_memcpy2d_h2d(ctx, dst.as<float>(), static_cast<size_t>(ld) * sizeof(float), m.ptr<float>(0), m.step(),
                                   static_cast<size_t>(m.cols) * sizeof(float), static_cast<size_t>(m.rows)), "upload");
    }
    check(ctx, sd_sync(ctx), "upload");   // the host Mat may go away after this call
}

inline cv::Mat download(const float* d, int rows, int cols, int64_t ld)
{
    cv::Mat m(rows, cols, CV_32FC1);
    sd_ctx* ctx = context();
    if (ld == cols) {
        check(ctx, sd_memcpy_d2h(ctx, m.ptr<float>(0), d, static_cast<size_t>(rows) * cols * sizeof(float)), "download");
    } else {
        check(ctx, sd_memcpy2d_d2h(ctx, m.ptr<float>(0), m.step(), d, static_cast<size_t>(ld) * sizeof(float),
                                   static_cast<size_t>(cols) * sizeof(float), static_cast<size_t>(rows)), "download");
    }
    check(ctx, sd_sync(ctx), "download");
    return m;
}

}  // namespace sd_b200
