// B200 drop-in for the reference's include/superviseddescent/superviseddescent.hpp.
//
// SupervisedDescentOptimiser<RegressorType, NormalisationStrategy> keeps train / test / predict with the
// reference's signatures (superviseddescent.hpp:85-361) and its callback types (:52-54).  Two routes:
//
//   device route   RegressorType is this package's LinearRegressor<>, the projection is a device
//                  projection (rcr::HogTransform) and the normalisation maps to sd_normalisation:
//                  features, targets, Gram, solve and update all stay in HBM (sd_hog_batch ->
//                  sd_cascade_targets -> sd_gram -> sd_solve_gram -> sd_cascade_update).
//   functor route  any other projection functor h(row, level, idx) -> Mat | float is USER host code; it is
//                  evaluated on a pool of host threads exactly as the reference does (:173-189) and the
//                  stacked feature matrix goes through RegressorType::learn / predict (which are GPU calls
//                  for LinearRegressor<>).  That is the reference's API for user functors, not a fallback.
#pragma once

#include <functional>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

#include "superviseddescent/regressors.hpp"

namespace superviseddescent {

inline void no_eval(const cv::Mat& /*current_predictions*/) {}   // superviseddescent.hpp:52-54

class NoNormalisation {   // superviseddescent.hpp:60-74
public:
    inline cv::Mat operator()(cv::Mat params) { return cv::Mat::ones(1, params.cols, params.type()); }
    sd_normalisation c_normalisation() const { sd_normalisation n{}; n.kind = 0; return n; }
};

namespace detail {

template <class...> struct voider { using type = void; };
template <class... T> using void_t = typename voider<T...>::type;

// projection that can fill a device matrix for all rows at once (rcr::HogTransform)
template <class P, class = void> struct is_device_projection : std::false_type {};
template <class P> struct is_device_projection<P, void_t<decltype(std::declval<P&>().project_device(static_cast<const float*>(nullptr), int64_t(0), 0, size_t(0), static_cast<float*>(nullptr), int64_t(0)))>> : std::true_type {};

template <class N, class = void> struct has_c_normalisation : std::false_type {};
template <class N> struct has_c_normalisation<N, void_t<decltype(std::declval<const N&>().c_normalisation())>> : std::true_type {};

template <class R, class = void> struct is_device_regressor : std::false_type {};
template <class R> struct is_device_regressor<R, void_t<decltype(std::declval<R&>().device_x()), decltype(std::declval<R&>().get_regulariser())>> : std::true_type {};

inline cv::Mat as_row(float v) { cv::Mat m(1, 1, CV_32FC1); m.at<float>(0, 0) = v; return m; }
inline cv::Mat as_row(double v) { return as_row(static_cast<float>(v)); }
inline cv::Mat as_row(const cv::Mat& m) { return m; }

// h(current_x.row(i), level, i) for every row, on hardware_concurrency() host threads (superviseddescent.hpp:173-189)
template <class ProjectionFunction>
cv::Mat project_on_host(const cv::Mat& current_x, size_t level, ProjectionFunction projection)
{
    const int n = current_x.rows;
    std::vector<cv::Mat> rows(n);
    unsigned threads = std::thread::hardware_concurrency();
    if (threads == 0) threads = 4;
    if (threads > static_cast<unsigned>(n)) threads = n > 0 ? n : 1;
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; ++t) {
        pool.emplace_back([&, t, projection]() mutable {      // each worker owns a copy of the functor
            for (int i = t; i < n; i += threads) rows[i] = as_row(projection(current_x.row(i), level, i)).clone();
        });
    }
    for (auto& th : pool) th.join();
    cv::Mat features;
    for (int i = 0; i < n; ++i) features.push_back(rows[i]);
    return features;
}

}  // namespace detail

template <class RegressorType, class NormalisationStrategy = NoNormalisation>
class SupervisedDescentOptimiser {
public:
    SupervisedDescentOptimiser() = default;
    SupervisedDescentOptimiser(std::vector<RegressorType> regressors, NormalisationStrategy normalisation = NormalisationStrategy())
        : regressors(std::move(regressors)), normalisation_strategy(std::move(normalisation)) {}

    template <class ProjectionFunction>
    void train(cv::Mat parameters, cv::Mat initialisations, cv::Mat templates, ProjectionFunction projection)
    {
        want_callback = false;   // no host copy of current_x per level when nobody listens ("Metrics")
        train(parameters, initialisations, templates, projection, no_eval);
        want_callback = true;
    }

    // Multi-GPU (one process per GPU, each passing ITS shard of the rows): the optional communicator of the C ABI.  Per level
    // the local [AtA | Atb] is exchanged and solved by sd_learn_dist; every rank ends with the same regressors.  route: 0 =
    // all-reduce + every rank solves, 1 = reduce to the panel owners + distributed Cholesky, 2 = all-reduce + CG shared by the
    // ranks.  The callback receives the rows of ALL ranks (superviseddescent.hpp:217).
    template <class ProjectionFunction, class OnTrainingEpochCallback>
    void train(cv::Mat parameters, cv::Mat initialisations, cv::Mat templates, ProjectionFunction projection,
               OnTrainingEpochCallback on_training_epoch_callback, sd_comm* communicator, int route = 0)
    {
        static_assert(detail::is_device_projection<ProjectionFunction>::value && detail::has_c_normalisation<NormalisationStrategy>::value &&
                          detail::is_device_regressor<RegressorType>::value,
                      "multi-GPU training needs the device route (HogTransform projection, device regressors)");
        comm = communicator;
        comm_route = route;
        train_impl(parameters, initialisations, templates, projection, on_training_epoch_callback, std::true_type());
        comm = nullptr;
    }

    // superviseddescent.hpp:165-219
    template <class ProjectionFunction, class OnTrainingEpochCallback>
    void train(cv::Mat parameters, cv::Mat initialisations, cv::Mat templates, ProjectionFunction projection,
               OnTrainingEpochCallback on_training_epoch_callback)
    {
        train_impl(parameters, initialisations, templates, projection, on_training_epoch_callback,
                   std::integral_constant<bool, detail::is_device_projection<ProjectionFunction>::value &&
                                                    detail::has_c_normalisation<NormalisationStrategy>::value &&
                                                    detail::is_device_regressor<RegressorType>::value>());
    }

    template <class ProjectionFunction>
    cv::Mat test(cv::Mat initialisations, cv::Mat templates, ProjectionFunction projection)
    {
        want_callback = false;
        cv::Mat out = test(initialisations, templates, projection, no_eval);
        want_callback = true;
        return out;
    }

    // superviseddescent.hpp:262-306
    template <class ProjectionFunction, class OnRegressorIterationCallback>
    cv::Mat test(cv::Mat initialisations, cv::Mat templates, ProjectionFunction projection,
                 OnRegressorIterationCallback on_regressor_iteration_callback)
    {
        return test_impl(initialisations, templates, projection, on_regressor_iteration_callback,
                         std::integral_constant<bool, detail::is_device_projection<ProjectionFunction>::value &&
                                                          detail::has_c_normalisation<NormalisationStrategy>::value &&
                                                          detail::is_device_regressor<RegressorType>::value>());
    }

    // superviseddescent.hpp:323-344 (single row or batch; same arithmetic as test without a callback)
    template <class ProjectionFunction>
    cv::Mat predict(cv::Mat initialisations, cv::Mat templates, ProjectionFunction projection)
    {
        return test(initialisations, templates, projection);
    }

    std::vector<RegressorType>& get_regressors() { return regressors; }
    NormalisationStrategy& get_normalisation() { return normalisation_strategy; }

private:
    std::vector<RegressorType> regressors;
    NormalisationStrategy normalisation_strategy;
    bool want_callback = true;

    // ------------------------------------------------------------------ functor route (host projection)
    template <class P, class CB>
    void train_impl(cv::Mat parameters, cv::Mat initialisations, cv::Mat templates, P projection, CB cb, std::false_type)
    {
        using cv::Mat;
        Mat current_x = initialisations;
        for (size_t level = 0; level < regressors.size(); ++level) {
            Mat features = detail::project_on_host(current_x, level, projection);
            Mat observed = templates.empty() ? features : Mat(features - templates);              // :191-197
            Mat b(current_x.rows, current_x.cols, CV_32FC1);                                     // :199-205
            for (int i = 0; i < current_x.rows; ++i) {
                Mat n = normalisation_strategy(current_x.row(i));
                for (int j = 0; j < current_x.cols; ++j)
                    b.at<float>(i, j) = (current_x.at<float>(i, j) - parameters.at<float>(i, j)) * n.at<float>(0, j);
            }
            regressors[level].learn(observed, b);                                                // :207
            current_x = apply_level_host(level, observed, current_x);                            // :209-215
            cb(current_x);                                                                       // :217
        }
    }

    template <class P, class CB>
    cv::Mat test_impl(cv::Mat initialisations, cv::Mat templates, P projection, CB cb, std::false_type)
    {
        using cv::Mat;
        Mat current_x = initialisations;
        for (size_t level = 0; level < regressors.size(); ++level) {
            Mat features = detail::project_on_host(current_x, level, projection);
            Mat observed = templates.empty() ? features : Mat(features - templates);
            current_x = apply_level_host(level, observed, current_x);
            cb(current_x);                                                                       // :303
        }
        return current_x;
    }

    cv::Mat apply_level_host(size_t level, const cv::Mat& observed, const cv::Mat& current_x)
    {
        cv::Mat update = regressors[level].predict(observed);      // one batched GPU GEMM instead of N GEMVs
        cv::Mat x_k(current_x.rows, current_x.cols, CV_32FC1);
        for (int i = 0; i < current_x.rows; ++i) {
            cv::Mat n = normalisation_strategy(current_x.row(i));
            for (int j = 0; j < current_x.cols; ++j)
                x_k.at<float>(i, j) = current_x.at<float>(i, j) - update.at<float>(i, j) * (1.0f / n.at<float>(0, j));
        }
        return x_k;
    }

    sd_comm* comm = nullptr;     // set for the duration of a multi-GPU train()
    int comm_route = 0;

    // ------------------------------------------------------------------ device route
    template <class P, class CB>
    void train_impl(cv::Mat parameters, cv::Mat initialisations, cv::Mat templates, P projection, CB cb, std::true_type)
    {
        sd_ctx* ctx = sd_b200::context();
        const int n = initialisations.rows, Pd = initialisations.cols;
        sd_b200::DeviceBuffer d_gt, d_cur, d_next(static_cast<size_t>(n) * Pd * sizeof(float)), d_tmpl;
        sd_b200::upload(parameters, d_gt, Pd);
        sd_b200::upload(initialisations, d_cur, Pd);
        const sd_normalisation norm = normalisation_strategy.c_normalisation();
        sd_b200::DeviceBuffer A, G, X, Xc, mu;
        int64_t n_global = n;
        const int nranks = comm ? sd_comm_size(comm) : 1;
        if (nranks > 1) sd_b200::check(ctx, sd_comm_sum_int64(ctx, comm, &n_global), "sd_comm_sum_int64");
        for (size_t level = 0; level < regressors.size(); ++level) {
            const int D = projection.feature_length(level);
            const int64_t ld = (static_cast<int64_t>(D) + Pd + 3) / 4 * 4;
            A.allocate(static_cast<size_t>(n) * ld * sizeof(float));
            projection.project_device(d_cur.as<float>(), Pd, n, level, A.as<float>(), ld);                    // 1) :173-189
            if (!templates.empty()) {
                sd_b200::upload(templates, d_tmpl, templates.cols);
                sd_b200::check(ctx, sd_subtract_templates(ctx, A.as<float>(), ld, d_tmpl.as<float>(), templates.cols, n, D), "sd_subtract_templates");
            }
            float* B = A.as<float>() + D;                                                                    // 2) :199-205
            sd_b200::check(ctx, sd_cascade_targets(ctx, d_cur.as<float>(), d_gt.as<float>(), n, Pd, &norm, B, ld), "sd_cascade_targets");
            X.allocate(static_cast<size_t>(D) * Pd * sizeof(float));                                          // 3) :207
            const sd_regulariser reg = regressors[level].get_regulariser().c();
            // learn on centred rows (sd_centre_features: column means over all ranks, subtracted in place; no-op for D <= 256);
            // X is the model, Xc the weights that go with the centred buffer
            Xc.allocate(static_cast<size_t>(D) * Pd * sizeof(float));
            mu.allocate(static_cast<size_t>(D) * sizeof(float));
            sd_comm* c = nranks > 1 ? comm : nullptr;
            sd_b200::check(ctx, sd_centre_features(ctx, c, A.as<float>(), ld, n, D, static_cast<int>(n_global), &reg, mu.as<float>()), "sd_centre_features");
            sd_b200::check(ctx, sd_learn_centred(ctx, c, A.as<float>(), ld, B, ld, n, D, Pd, &reg, static_cast<int>(n_global), nranks > 1 ? comm_route : 0,
                                                 mu.as<float>(), X.as<float>(), Xc.as<float>(), nullptr), "sd_learn_centred");
            regressors[level].set_x(sd_b200::download(X.as<float>(), D, Pd, Pd));
            regressors[level].report_solver();
            sd_b200::check(ctx, sd_cascade_update(ctx, A.as<float>(), ld, n, D, Xc.as<float>(), Pd, d_cur.as<float>(), &norm, d_next.as<float>()), "sd_cascade_update");   // 4) :209-215
            std::swap(d_cur, d_next);
            if (want_callback) {                                                                             // 5) :217
                if (nranks > 1) {
                    // every rank contributes its rows; shards are padded to the largest one for the gather
                    int64_t most = n;
                    std::vector<int64_t> counts(nranks, 0);
                    for (int r = 0; r < nranks; ++r) {
                        int64_t v = (r == sd_comm_rank(comm)) ? n : 0;
                        sd_b200::check(ctx, sd_comm_sum_int64(ctx, comm, &v), "sd_comm_sum_int64");
                        counts[r] = v;
                        most = v > most ? v : most;
                    }
                    const size_t row_bytes = static_cast<size_t>(Pd) * sizeof(float);
                    sd_b200::DeviceBuffer send(static_cast<size_t>(most) * row_bytes), recv(static_cast<size_t>(most) * row_bytes * nranks);
                    sd_b200::check(ctx, sd_memset(ctx, send.as<float>(), 0, static_cast<size_t>(most) * row_bytes), "gather");
                    sd_b200::check(ctx, sd_memcpy2d_d2d(ctx, send.as<float>(), row_bytes, d_cur.as<float>(), row_bytes, row_bytes, n), "gather");
                    sd_b200::check(ctx, sd_comm_allgather(ctx, comm, send.as<float>(), static_cast<size_t>(most) * row_bytes, recv.as<float>()), "sd_comm_allgather");
                    cv::Mat all;
                    for (int r = 0; r < nranks; ++r)
                        if (counts[r] > 0) all.push_back(sd_b200::download(recv.as<float>() + static_cast<size_t>(r) * most * Pd, static_cast<int>(counts[r]), Pd, Pd));
                    cb(all);
                } else {
                    cb(sd_b200::download(d_cur.as<float>(), n, Pd, Pd));
                }
            }
        }
    }

    template <class P, class CB>
    cv::Mat test_impl(cv::Mat initialisations, cv::Mat templates, P projection, CB cb, std::true_type)
    {
        sd_ctx* ctx = sd_b200::context();
        const int n = initialisations.rows, Pd = initialisations.cols;
        sd_b200::DeviceBuffer d_cur, d_next(static_cast<size_t>(n) * Pd * sizeof(float)), d_tmpl, A;
        sd_b200::upload(initialisations, d_cur, Pd);
        const sd_normalisation norm = normalisation_strategy.c_normalisation();
        for (size_t level = 0; level < regressors.size(); ++level) {
            const int D = projection.feature_length(level);
            const int64_t ld = (static_cast<int64_t>(D) + 3) / 4 * 4;
            A.allocate(static_cast<size_t>(n) * ld * sizeof(float));
            projection.project_device(d_cur.as<float>(), Pd, n, level, A.as<float>(), ld);
            if (!templates.empty()) {
                sd_b200::upload(templates, d_tmpl, templates.cols);
                sd_b200::check(ctx, sd_subtract_templates(ctx, A.as<float>(), ld, d_tmpl.as<float>(), templates.cols, n, D), "sd_subtract_templates");
            }
            sd_b200::check(ctx, sd_cascade_update(ctx, A.as<float>(), ld, n, D, regressors[level].device_x(), Pd, d_cur.as<float>(), &norm, d_next.as<float>()), "sd_cascade_update");
            std::swap(d_cur, d_next);
            if (want_callback) cb(sd_b200::download(d_cur.as<float>(), n, Pd, Pd));                          // :303
        }
        return sd_b200::download(d_cur.as<float>(), n, Pd, Pd);
    }
};

}  // namespace superviseddescent
