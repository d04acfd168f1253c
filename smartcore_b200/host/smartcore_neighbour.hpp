// smartcore_neighbour.hpp -- host-side mirror of the reference's brute-force neighbour search for row vectors under
// the Euclidian metric, with the search itself on the GPU:
//
//   smartcore::algorithm::neighbour::linear_search::LinearKNNSearch   src/algorithm/neighbour/linear_search.rs:35-84
//   with D = smartcore::metrics::distance::euclidian::Euclidian       src/metrics/distance/euclidian.rs:51-76
//
// `new` uploads the data once (the reference moves the Vec into the struct); `find(from, k)` and
// `find_radius(from, radius)` are the reference calls, `find_batch` the batched form of `find` (what KNNClassifier::predict does row by row, knn_classifier.rs).  Results are
// (index, distance) pairs ascending by (distance, index); the reference returns the same pairs in the internal order
// of its HeapSelection (see include/smartcore_kmeans_cuda.h, sckm_knn, for the tie rule).
#pragma once
#include <utility>
#include <vector>

#include "smartcore_kmeans.hpp"

namespace smartcore { namespace algorithm { namespace neighbour { namespace linear_search {

template <typename TX> class LinearKNNSearch {
public:
    LinearKNNSearch() = default;
    LinearKNNSearch(const LinearKNNSearch&) = delete;
    LinearKNNSearch& operator=(const LinearKNNSearch&) = delete;
    LinearKNNSearch(LinearKNNSearch&& o) noexcept { *this = std::move(o); }
    LinearKNNSearch& operator=(LinearKNNSearch&& o) noexcept {
        if (this != &o) { release(); ds_ = o.ds_; n_ = o.n_; d_ = o.d_; o.ds_ = nullptr; }
        return *this;
    }
    ~LinearKNNSearch() { release(); }

    // LinearKNNSearch::new(data, Distances::euclidian()) -- data: n rows of d values
    static error::Result<LinearKNNSearch> new_(const linalg::basic::matrix::DenseMatrix<TX>& data) {
        using R = error::Result<LinearKNNSearch>;
        auto dev = cluster::kmeans::Device::get();
        if (dev.is_err()) return R::Err(dev.unwrap_err());
        cluster::kmeans::Packed<TX> p; cluster::kmeans::pack(data, p);
        LinearKNNSearch s;
        if (sckm_dataset_upload(dev.unwrap(), p.ptr, data.nrows, data.ncols, p.dtype, p.column_major, 0, data.nrows, &s.ds_) != SCKM_OK)
            return R::Err(error::Failed::input(sckm_last_error(dev.unwrap())));
        s.n_ = data.nrows; s.d_ = data.ncols;
        return R::Ok(std::move(s));
    }

    // find(from, k) (linear_search.rs:52-84): Failed(FindFailed, "k should be >= 1 and <= length(data)") as there
    error::Result<std::vector<std::pair<size_t, double>>> find(const std::vector<TX>& from, size_t k) const {
        auto r = find_batch(from, 1, k);
        if (r.is_err()) return error::Result<std::vector<std::pair<size_t, double>>>::Err(r.unwrap_err());
        return error::Result<std::vector<std::pair<size_t, double>>>::Ok(std::move(r.unwrap()[0]));
    }

    // the same for nq query rows stored row-major in `queries`
    error::Result<std::vector<std::vector<std::pair<size_t, double>>>> find_batch(const std::vector<TX>& queries, size_t nq,
                                                                                 size_t k) const {
        using R = error::Result<std::vector<std::vector<std::pair<size_t, double>>>>;
        if (k < 1 || k > n_) return R::Err(error::Failed{error::FailedError::FindFailed, "k should be >= 1 and <= length(data)"});
        if (queries.size() != nq * d_) return R::Err(error::Failed{error::FailedError::FindFailed, "query length differs from the data"});
        auto dev = cluster::kmeans::Device::get();
        if (dev.is_err()) return R::Err(dev.unwrap_err());
        std::vector<int64_t> idx(nq * k); std::vector<double> dist(nq * k);
        int rc;
        if constexpr (std::is_same<TX, float>::value || std::is_same<TX, double>::value) {
            rc = sckm_knn(ds_, queries.data(), nq, k, idx.data(), dist.data());
        } else {                                                   // other Number types were widened to f64 on upload
            std::vector<double> w(queries.begin(), queries.end());
            rc = sckm_knn(ds_, w.data(), nq, k, idx.data(), dist.data());
        }
        if (rc != SCKM_OK) return R::Err(error::Failed{error::FailedError::FindFailed, sckm_last_error(dev.unwrap())});
        std::vector<std::vector<std::pair<size_t, double>>> out(nq);
        for (size_t q = 0; q < nq; q++)
            for (size_t j = 0; j < k; j++)
                if (idx[q * k + j] >= 0) out[q].emplace_back((size_t)idx[q * k + j], dist[q * k + j]);   // NaN: fewer tuples
        return R::Ok(std::move(out));
    }

    // find_radius(from, radius) (linear_search.rs:89-110): rows with distance <= radius, ascending row order
    error::Result<std::vector<std::pair<size_t, double>>> find_radius(const std::vector<TX>& from, double radius) const {
        using R = error::Result<std::vector<std::pair<size_t, double>>>;
        if (!(radius > 0.0)) return R::Err(error::Failed{error::FailedError::FindFailed, "radius should be > 0"});
        if (from.size() != d_) return R::Err(error::Failed{error::FailedError::FindFailed, "query length differs from the data"});
        auto dev = cluster::kmeans::Device::get();
        if (dev.is_err()) return R::Err(dev.unwrap_err());
        std::vector<double> w;
        const void* q = from.data();
        if constexpr (!(std::is_same<TX, float>::value || std::is_same<TX, double>::value)) { w.assign(from.begin(), from.end()); q = w.data(); }
        int64_t count = 0, offset = 0;
        if (sckm_radius_count(ds_, q, 1, radius, &count) != SCKM_OK)
            return R::Err(error::Failed{error::FailedError::FindFailed, sckm_last_error(dev.unwrap())});
        std::vector<int64_t> idx((size_t)count); std::vector<double> dist((size_t)count);
        if (sckm_radius_fill(ds_, q, 1, radius, &offset, (uint64_t)count, idx.data(), dist.data()) != SCKM_OK)
            return R::Err(error::Failed{error::FailedError::FindFailed, sckm_last_error(dev.unwrap())});
        std::vector<std::pair<size_t, double>> out;
        out.reserve((size_t)count);
        for (int64_t i = 0; i < count; i++) out.emplace_back((size_t)idx[(size_t)i], dist[(size_t)i]);
        return R::Ok(std::move(out));
    }

    size_t len() const { return n_; }

private:
    void release() { if (ds_) sckm_dataset_destroy(ds_); ds_ = nullptr; }
    sckm_dataset* ds_ = nullptr;
    size_t n_ = 0, d_ = 0;
};

}}}}  // namespace smartcore::algorithm::neighbour::linear_search
