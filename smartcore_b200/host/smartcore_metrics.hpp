// smartcore_metrics.hpp -- host-side mirror of the reference's cluster-quality scores, with the O(n) counting
// done on the device:
//
//   smartcore::metrics::cluster_helpers::contingency_matrix   src/metrics/cluster_helpers.rs:7-25
//   smartcore::metrics::cluster_helpers::entropy              src/metrics/cluster_helpers.rs:27-48
//   smartcore::metrics::cluster_helpers::mutual_info_score    src/metrics/cluster_helpers.rs:50-104
//   smartcore::metrics::cluster_hcv::HCVScore                 src/metrics/cluster_hcv.rs:12-55
//
// The reference maps labels to dense indices with `unique_with_indices` (arrays.rs:233-247: sort + dedup, then a
// linear search per element, O(n*u)) and counts pairs on the host.  Here the dense mapping is one ordered-map pass
// and the pair counting is sckm_contingency_host on the GPU; the scores are then functions of the small table.
// The reference's entropy sums over a HashMap (iteration order unspecified); this mirror sums in ascending label
// order, which is one of the orders the reference may produce.
#pragma once
#include <cmath>
#include <cstdint>
#include <map>
#include <optional>
#include <string>
#include <vector>

#include "smartcore_kmeans.hpp"

namespace smartcore { namespace metrics {

namespace cluster_helpers {

// dense index of every element in the sorted, de-duplicated label set (unique_with_indices, arrays.rs:233-247)
template <typename T>
inline std::pair<std::vector<T>, std::vector<uint32_t>> unique_with_indices(const std::vector<T>& v) {
    std::map<T, uint32_t> pos;
    for (const T& e : v) pos.emplace(e, 0);
    std::vector<T> unique;
    unique.reserve(pos.size());
    uint32_t next = 0;
    for (auto& kv : pos) { kv.second = next++; unique.push_back(kv.first); }
    std::vector<uint32_t> idx(v.size());
    for (size_t i = 0; i < v.size(); i++) idx[i] = pos[v[i]];
    return {std::move(unique), std::move(idx)};
}

// contingency_matrix (cluster_helpers.rs:7-25): rows = classes of labels_true, columns = clusters of labels_pred
template <typename T>
inline error::Result<std::vector<std::vector<size_t>>> contingency_matrix(const std::vector<T>& labels_true,
                                                                          const std::vector<T>& labels_pred) {
    using R = error::Result<std::vector<std::vector<size_t>>>;
    if (labels_true.size() != labels_pred.size()) return R::Err(error::Failed::input("label vectors differ in length"));
    auto a = unique_with_indices(labels_true);
    auto b = unique_with_indices(labels_pred);
    const size_t na = a.first.size(), nb = b.first.size();
    std::vector<std::vector<size_t>> out(na, std::vector<size_t>(nb, 0));
    if (na == 0 || nb == 0) return R::Ok(std::move(out));
    auto dev = cluster::kmeans::Device::get();
    if (dev.is_err()) return R::Err(dev.unwrap_err());
    std::vector<int64_t> flat(na * nb);
    if (sckm_contingency_host(dev.unwrap(), a.second.data(), b.second.data(), labels_true.size(), na, nb, flat.data()) != SCKM_OK)
        return R::Err(error::Failed::input(sckm_last_error(dev.unwrap())));
    for (size_t r = 0; r < na; r++) for (size_t c = 0; c < nb; c++) out[r][c] = (size_t)flat[r * nb + c];
    return R::Ok(std::move(out));
}

// entropy of a histogram (cluster_helpers.rs:27-48 after the bin counting)
inline std::optional<double> entropy_of_counts(const std::vector<size_t>& counts) {
    double entropy = 0.0;
    long long sum = 0;
    for (size_t c : counts) sum += (long long)c;
    for (size_t c : counts)
        if (c > 0) {
            const double pi = (double)c;
            entropy -= (pi / (double)sum) * (std::log(pi) - std::log((double)sum));
        }
    return entropy;
}

template <typename T> inline std::optional<double> entropy(const std::vector<T>& data) {
    std::map<long long, size_t> bins;
    for (const T& e : data) bins[(long long)e]++;
    std::vector<size_t> counts;
    for (auto& kv : bins) counts.push_back(kv.second);
    return entropy_of_counts(counts);
}

// mutual_info_score (cluster_helpers.rs:50-104), same operation order
inline double mutual_info_score(const std::vector<std::vector<size_t>>& contingency) {
    if (contingency.empty() || contingency[0].empty()) return 0.0;
    const size_t nr = contingency.size(), nc = contingency[0].size();
    size_t contingency_sum = 0;
    std::vector<size_t> pi(nr, 0), pj(nc, 0), nzx, nzy, nz_val;
    for (size_t r = 0; r < nr; r++)
        for (size_t c = 0; c < nc; c++) {
            contingency_sum += contingency[r][c];
            pi[r] += contingency[r][c];
            pj[c] += contingency[r][c];
            if (contingency[r][c] > 0) { nzx.push_back(r); nzy.push_back(c); nz_val.push_back(contingency[r][c]); }
        }
    const double csum = (double)contingency_sum, csum_ln = std::log(csum);
    size_t pi_sum = 0, pj_sum = 0;
    for (size_t v : pi) pi_sum += v;
    for (size_t v : pj) pj_sum += v;
    const double pi_sum_l = std::log((double)pi_sum), pj_sum_l = std::log((double)pj_sum);
    double result = 0.0;
    for (size_t i = 0; i < nz_val.size(); i++) {
        const double log_nm = std::log((double)nz_val[i]);
        const double nm = (double)nz_val[i] / csum;
        const double log_outer = -std::log((double)(pi[nzx[i]] * pj[nzy[i]])) + pi_sum_l + pj_sum_l;
        result += (nm * (log_nm - csum_ln)) + nm * log_outer;
    }
    return result > 0.0 ? result : 0.0;
}

}  // namespace cluster_helpers

namespace cluster_hcv {

// HCVScore (cluster_hcv.rs:12-55)
template <typename T> class HCVScore {
public:
    std::optional<double> homogeneity() const { return homogeneity_; }
    std::optional<double> completeness() const { return completeness_; }
    std::optional<double> v_measure() const { return v_measure_; }

    // from a contingency table (rows = true classes, columns = predicted clusters): entropies are those of its margins
    void compute_from_table(const std::vector<std::vector<size_t>>& contingency) {
        std::vector<size_t> rows(contingency.size(), 0), cols(contingency.empty() ? 0 : contingency[0].size(), 0);
        for (size_t r = 0; r < contingency.size(); r++)
            for (size_t c = 0; c < contingency[r].size(); c++) { rows[r] += contingency[r][c]; cols[c] += contingency[r][c]; }
        const auto entropy_c = cluster_helpers::entropy_of_counts(rows);
        const auto entropy_k = cluster_helpers::entropy_of_counts(cols);
        const double mi = cluster_helpers::mutual_info_score(contingency);
        // `entropy_c.map(|e| mi / e).unwrap_or(0)` (cluster_hcv.rs:42-43): entropy is always Some, so 0/0 = NaN
        // survives exactly as in the reference
        const double h = entropy_c ? mi / *entropy_c : 0.0;
        const double c = entropy_k ? mi / *entropy_k : 0.0;
        const double v = (h + c == 0.0) ? 0.0 : 2.0 * h * c / (1.0 * h + c);
        homogeneity_ = h; completeness_ = c; v_measure_ = v;
    }

    // compute (cluster_hcv.rs:36-55)
    error::Result<bool> compute(const std::vector<T>& y_true, const std::vector<T>& y_pred) {
        auto t = cluster_helpers::contingency_matrix(y_true, y_pred);
        if (t.is_err()) return error::Result<bool>::Err(t.unwrap_err());
        compute_from_table(t.unwrap());
        return error::Result<bool>::Ok(true);
    }

private:
    std::optional<double> homogeneity_, completeness_, v_measure_;
};

}  // namespace cluster_hcv
}}  // namespace smartcore::metrics
