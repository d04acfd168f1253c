// smartcore_host_shim.cpp -- flat C entry points over the C++ host mirror (smartcore_kmeans.hpp) so
// that Python (tests/, bench.py, smartcore_b200/cluster.py) can drive KMeans::fit / predict exactly as
// a Rust caller would.  No arithmetic here: validation + RNG draws live in the header, the compute in
// libsmartcore_kmeans_cuda.so.
#include "smartcore_kmeans.hpp"
#include "smartcore_metrics.hpp"
#include "smartcore_neighbour.hpp"
#include <cstring>

using namespace smartcore;
using smartcore::cluster::kmeans::KMeansParameters;
using smartcore::cluster::kmeans::KMeansSearchParameters;
using smartcore::linalg::basic::matrix::DenseMatrix;

namespace {
enum { SCH_F32 = 0, SCH_F64 = 1, SCH_I32 = 2, SCH_I64 = 3 };

struct Model {  // type-erased KMeans<TX, usize>
    int dtype;
    cluster::kmeans::KMeans<float, size_t> f32;
    cluster::kmeans::KMeans<double, size_t> f64;
    cluster::kmeans::KMeans<int32_t, size_t> i32;
    cluster::kmeans::KMeans<int64_t, size_t> i64;
};

void set_err(char* buf, size_t len, const std::string& s) {
    if (!buf || !len) return;
    size_t m = s.size() < len - 1 ? s.size() : len - 1;
    memcpy(buf, s.data(), m); buf[m] = 0;
}

template <typename T> DenseMatrix<T> wrap(const void* values, size_t nrows, size_t ncols, int column_major) {
    const T* v = (const T*)values;
    return DenseMatrix<T>::new_(nrows, ncols, std::vector<T>(v, v + nrows * ncols), column_major != 0).unwrap();
}

template <typename T, typename M>
int do_fit(M& slot, const void* values, size_t nrows, size_t ncols, int column_major, const KMeansParameters& p,
           char* err, size_t errlen) {
    auto x = wrap<T>(values, nrows, ncols, column_major);
    auto r = cluster::kmeans::KMeans<T, size_t>::fit(x, p);
    if (r.is_err()) { set_err(err, errlen, r.unwrap_err().to_string()); return 1; }
    slot = std::move(r.unwrap());
    return 0;
}

template <typename T, typename M>
int do_predict(const M& m, const void* values, size_t nrows, size_t ncols, int column_major, int64_t* out,
               char* err, size_t errlen) {
    auto x = wrap<T>(values, nrows, ncols, column_major);
    auto r = m.predict(x);
    if (r.is_err()) { set_err(err, errlen, r.unwrap_err().to_string()); return 1; }
    for (size_t i = 0; i < nrows; i++) out[i] = (int64_t)r.unwrap()[i];
    return 0;
}
}  // namespace

extern "C" {

// KMeans::fit(&DenseMatrix::new(nrows, ncols, values, column_major), KMeansParameters{k, max_iter, seed})
int sch_kmeans_fit(int dtype, const void* values, size_t nrows, size_t ncols, int column_major, size_t k,
                   size_t max_iter, int has_seed, uint64_t seed, void** model_out, char* err, size_t errlen) {
    KMeansParameters p; p.k = k; p.max_iter = max_iter;
    if (has_seed) p.seed = seed;
    Model* m = new Model(); m->dtype = dtype;
    int rc = 2;
    switch (dtype) {
        case SCH_F32: rc = do_fit<float>(m->f32, values, nrows, ncols, column_major, p, err, errlen); break;
        case SCH_F64: rc = do_fit<double>(m->f64, values, nrows, ncols, column_major, p, err, errlen); break;
        case SCH_I32: rc = do_fit<int32_t>(m->i32, values, nrows, ncols, column_major, p, err, errlen); break;
        case SCH_I64: rc = do_fit<int64_t>(m->i64, values, nrows, ncols, column_major, p, err, errlen); break;
        default: set_err(err, errlen, "unsupported dtype");
    }
    if (rc) { delete m; return rc; }
    *model_out = m;
    return 0;
}

int sch_kmeans_predict(void* model, const void* values, size_t nrows, size_t ncols, int column_major, int64_t* out,
                       char* err, size_t errlen) {
    Model* m = (Model*)model;
    switch (m->dtype) {
        case SCH_F32: return do_predict<float>(m->f32, values, nrows, ncols, column_major, out, err, errlen);
        case SCH_F64: return do_predict<double>(m->f64, values, nrows, ncols, column_major, out, err, errlen);
        case SCH_I32: return do_predict<int32_t>(m->i32, values, nrows, ncols, column_major, out, err, errlen);
        case SCH_I64: return do_predict<int64_t>(m->i64, values, nrows, ncols, column_major, out, err, errlen);
    }
    return 2;
}

#define WITH_MODEL(m, expr)                                   \
    switch (((Model*)m)->dtype) {                             \
        case SCH_F32: { auto& M = ((Model*)m)->f32; expr; } break; \
        case SCH_F64: { auto& M = ((Model*)m)->f64; expr; } break; \
        case SCH_I32: { auto& M = ((Model*)m)->i32; expr; } break; \
        default:      { auto& M = ((Model*)m)->i64; expr; } break; \
    }

// field access (the reference's fields are private; in-crate tests read kmeans._y -- kmeans.rs:503)
void sch_kmeans_dims(void* model, size_t* k, size_t* n, size_t* d) {
    WITH_MODEL(model, { *k = M.k; *n = M._y.size(); *d = M.centroids.empty() ? 0 : M.centroids[0].size(); })
}
void sch_kmeans_get(void* model, int64_t* y, int64_t* size, double* centroids, double* distortion, int64_t* iters) {
    WITH_MODEL(model, {
        if (y) for (size_t i = 0; i < M._y.size(); i++) y[i] = (int64_t)M._y[i];
        if (size) for (size_t i = 0; i < M.size.size(); i++) size[i] = (int64_t)M.size[i];
        if (centroids) { size_t o = 0; for (auto& row : M.centroids) for (double v : row) centroids[o++] = v; }
        if (distortion) *distortion = M._distortion;
        if (iters) *iters = M._iterations;
    })
}
void sch_kmeans_free(void* model) { delete (Model*)model; }

// serde images of the model (format 0 = serde_json, 1 = bincode).  Returns the image size; copies it when it fits.
size_t sch_kmeans_serialize(void* model, int format, char* buf, size_t cap) {
    std::string img;
    WITH_MODEL(model, { img = format == 0 ? M.to_json() : M.to_bincode(); })
    if (buf && img.size() <= cap) memcpy(buf, img.data(), img.size());
    return img.size();
}
int sch_kmeans_deserialize(int dtype, int format, const char* bytes, size_t len, void** model_out, char* err, size_t errlen) {
    Model* m = new Model(); m->dtype = dtype;
    const std::string img(bytes, len);
    int rc = 0;
    WITH_MODEL(m, {
        auto r = format == 0 ? std::decay_t<decltype(M)>::from_json(img) : std::decay_t<decltype(M)>::from_bincode(img);
        if (r.is_err()) { set_err(err, errlen, r.unwrap_err().to_string()); rc = 1; }
        else M = std::move(r.unwrap());
    })
    if (rc) { delete m; return rc; }
    *model_out = m;
    return 0;
}
// PartialEq (kmeans.rs:85-107)
int sch_kmeans_eq(void* a, void* b) {
    if (((Model*)a)->dtype != ((Model*)b)->dtype) return 0;
    int eq = 0;
    switch (((Model*)a)->dtype) {
        case SCH_F32: eq = ((Model*)a)->f32 == ((Model*)b)->f32; break;
        case SCH_F64: eq = ((Model*)a)->f64 == ((Model*)b)->f64; break;
        case SCH_I32: eq = ((Model*)a)->i32 == ((Model*)b)->i32; break;
        default: eq = ((Model*)a)->i64 == ((Model*)b)->i64; break;
    }
    return eq;
}

// KMeansSearchParameters{k, max_iter, seed}.into_iter() flattened; returns the number of items written
size_t sch_search_parameters(const size_t* k, size_t nk, const size_t* max_iter, size_t nm, const uint64_t* seed,
                             const int* has_seed, size_t ns, size_t* out_k, size_t* out_max_iter, uint64_t* out_seed,
                             int* out_has_seed, size_t cap) {
    KMeansSearchParameters sp;
    sp.k.assign(k, k + nk); sp.max_iter.assign(max_iter, max_iter + nm); sp.seed.clear();
    for (size_t i = 0; i < ns; i++) sp.seed.push_back(has_seed[i] ? std::optional<uint64_t>(seed[i]) : std::nullopt);
    auto it = sp.into_iter();
    size_t n = 0;
    while (auto p = it.next()) {
        if (n >= cap) break;
        out_k[n] = p->k; out_max_iter[n] = p->max_iter; out_has_seed[n] = p->seed.has_value(); out_seed[n] = p->seed.value_or(0);
        n++;
    }
    return n;
}

// which smartcore build the host mirror follows: 0 = default features (SmallRng), 1 = `std_rand` (StdRng = ChaCha12)
void sch_set_std_rand(int on) { rand_custom::set_std_rand(on != 0); }
int sch_get_std_rand(void) { return rand_custom::std_rand_switch() ? 1 : 0; }
// raw generator output (KATs): `count` next_u64 values after seed_from_u64(seed) of the selected RngImpl
void sch_rng_next_u64(uint64_t seed, int std_rand, size_t count, uint64_t* out) {
    auto r = rand_custom::RngImpl::seed_from_u64(seed, false, std_rand != 0);
    for (size_t i = 0; i < count; i++) out[i] = r.next_u64();
}
void sch_chacha_block(const uint32_t* key8, uint64_t counter, int rounds, uint32_t* out16) {
    uint32_t key[8], out[16];
    for (int i = 0; i < 8; i++) key[i] = key8[i];
    rand_custom::RngImpl::chacha_block(key, counter, rounds, out);
    for (int i = 0; i < 16; i++) out16[i] = out[i];
}

// the host RNG draw sequence of kmeans_plus_plus for (seed, n, k): first index + k-1 uniforms
void sch_kmeanspp_draws(int has_seed, uint64_t seed, uint64_t n, size_t k, uint64_t* first, double* uniforms) {
    auto rng = rand_custom::get_rng_impl(has_seed ? std::optional<uint64_t>(seed) : std::nullopt);
    *first = rng.gen_range(n);
    for (size_t j = 0; j + 1 < k; j++) uniforms[j] = rng.gen_f64();
}

// DenseMatrix::get over both layouts (matrix.rs:367-381), f64 only; for layout tests
double sch_dense_get_f64(const double* values, size_t nrows, size_t ncols, int column_major, size_t r, size_t c) {
    DenseMatrix<double> m; m.nrows = nrows; m.ncols = ncols; m.column_major = column_major != 0;
    return column_major ? values[c * nrows + r] : values[c + ncols * r];
}

// ---- metrics::cluster_helpers / cluster_hcv over i64 labels ------------------------------------------------
// contingency_matrix(labels_true, labels_pred): writes the table row-major into out (capacity cap cells) and its
// shape into nr / nc; returns 0, 1 (error, message in err) or 3 (cap too small; shape still reported)
int sch_contingency_matrix(const int64_t* y_true, const int64_t* y_pred, size_t n, int64_t* out, size_t cap, size_t* nr,
                           size_t* nc, char* err, size_t errlen) {
    std::vector<int64_t> a(y_true, y_true + n), b(y_pred, y_pred + n);
    auto r = metrics::cluster_helpers::contingency_matrix(a, b);
    if (r.is_err()) { set_err(err, errlen, r.unwrap_err().to_string()); return 1; }
    auto& t = r.unwrap();
    *nr = t.size(); *nc = t.empty() ? 0 : t[0].size();
    if (*nr * *nc > cap) return 3;
    for (size_t i = 0; i < *nr; i++) for (size_t j = 0; j < *nc; j++) out[i * *nc + j] = (int64_t)t[i][j];
    return 0;
}
double sch_entropy(const int64_t* y, size_t n) {
    return *metrics::cluster_helpers::entropy(std::vector<int64_t>(y, y + n));
}
double sch_mutual_info_score(const int64_t* table, size_t nr, size_t nc) {
    std::vector<std::vector<size_t>> t(nr, std::vector<size_t>(nc));
    for (size_t i = 0; i < nr; i++) for (size_t j = 0; j < nc; j++) t[i][j] = (size_t)table[i * nc + j];
    return metrics::cluster_helpers::mutual_info_score(t);
}
// HCVScore::compute: out3 = homogeneity, completeness, v_measure
int sch_hcv_score(const int64_t* y_true, const int64_t* y_pred, size_t n, double* out3, char* err, size_t errlen) {
    metrics::cluster_hcv::HCVScore<int64_t> s;
    auto r = s.compute(std::vector<int64_t>(y_true, y_true + n), std::vector<int64_t>(y_pred, y_pred + n));
    if (r.is_err()) { set_err(err, errlen, r.unwrap_err().to_string()); return 1; }
    out3[0] = *s.homogeneity(); out3[1] = *s.completeness(); out3[2] = *s.v_measure();
    return 0;
}
// the same scores from a table that was counted elsewhere (sckm_contingency on resident labels)
void sch_hcv_from_table(const int64_t* table, size_t nr, size_t nc, double* out3) {
    std::vector<std::vector<size_t>> t(nr, std::vector<size_t>(nc));
    for (size_t i = 0; i < nr; i++) for (size_t j = 0; j < nc; j++) t[i][j] = (size_t)table[i * nc + j];
    metrics::cluster_hcv::HCVScore<int64_t> s;
    s.compute_from_table(t);
    out3[0] = *s.homogeneity(); out3[1] = *s.completeness(); out3[2] = *s.v_measure();
}

// ---- algorithm::neighbour::linear_search::LinearKNNSearch<f32|f64> with the Euclidian metric -----------------------
struct KnnHandle {
    int dtype;
    algorithm::neighbour::linear_search::LinearKNNSearch<float> f32;
    algorithm::neighbour::linear_search::LinearKNNSearch<double> f64;
};
int sch_knn_new(int dtype, const void* values, size_t nrows, size_t ncols, int column_major, void** out, char* err, size_t errlen) {
    KnnHandle* h = new KnnHandle(); h->dtype = dtype;
    int rc = 0;
    if (dtype == SCH_F32) {
        auto r = algorithm::neighbour::linear_search::LinearKNNSearch<float>::new_(wrap<float>(values, nrows, ncols, column_major));
        if (r.is_err()) { set_err(err, errlen, r.unwrap_err().to_string()); rc = 1; } else h->f32 = std::move(r.unwrap());
    } else if (dtype == SCH_F64) {
        auto r = algorithm::neighbour::linear_search::LinearKNNSearch<double>::new_(wrap<double>(values, nrows, ncols, column_major));
        if (r.is_err()) { set_err(err, errlen, r.unwrap_err().to_string()); rc = 1; } else h->f64 = std::move(r.unwrap());
    } else { set_err(err, errlen, "unsupported dtype"); rc = 2; }
    if (rc) { delete h; return rc; }
    *out = h;
    return 0;
}
// find for nq query rows; counts[q] = tuples returned for query q (k, or fewer when distances are NaN)
int sch_knn_find(void* handle, const void* queries, size_t nq, size_t d, size_t k, int64_t* idx_out, double* dist_out,
                 int64_t* counts, char* err, size_t errlen) {
    KnnHandle* h = (KnnHandle*)handle;
    auto emit = [&](auto& r) -> int {
        if (r.is_err()) { set_err(err, errlen, r.unwrap_err().to_string()); return 1; }
        auto& v = r.unwrap();
        for (size_t q = 0; q < nq; q++) {
            counts[q] = (int64_t)v[q].size();
            for (size_t j = 0; j < v[q].size(); j++) { idx_out[q * k + j] = (int64_t)v[q][j].first; dist_out[q * k + j] = v[q][j].second; }
        }
        return 0;
    };
    if (h->dtype == SCH_F32) {
        const float* q = (const float*)queries;
        auto r = h->f32.find_batch(std::vector<float>(q, q + nq * d), nq, k);
        return emit(r);
    }
    const double* q = (const double*)queries;
    auto r = h->f64.find_batch(std::vector<double>(q, q + nq * d), nq, k);
    return emit(r);
}
// find_radius for one query row: returns the number of neighbours (or -1 with the message in err); copies up to cap
long long sch_knn_find_radius(void* handle, const void* query, size_t d, double radius, int64_t* idx_out, double* dist_out,
                              size_t cap, char* err, size_t errlen) {
    KnnHandle* h = (KnnHandle*)handle;
    auto emit = [&](auto& r) -> long long {
        if (r.is_err()) { set_err(err, errlen, r.unwrap_err().to_string()); return -1; }
        auto& v = r.unwrap();
        for (size_t j = 0; j < v.size() && j < cap; j++) { idx_out[j] = (int64_t)v[j].first; dist_out[j] = v[j].second; }
        return (long long)v.size();
    };
    if (h->dtype == SCH_F32) {
        const float* q = (const float*)query;
        auto r = h->f32.find_radius(std::vector<float>(q, q + d), radius);
        return emit(r);
    }
    const double* q = (const double*)query;
    auto r = h->f64.find_radius(std::vector<double>(q, q + d), radius);
    return emit(r);
}
void sch_knn_free(void* handle) { delete (KnnHandle*)handle; }

}  // extern "C"
