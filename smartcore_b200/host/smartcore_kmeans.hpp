// smartcore_kmeans.hpp -- host-side mirror of smartcore's k-means API over the CUDA C ABI.
//
// The reference host language is Rust; no Rust toolchain exists in this image, so the host side
// above the C ABI is written in C++ with the reference's names, argument meaning and error
// behaviour (the Rust overlay that a maintainer would add is in rust/ and INTEGRATION.md):
//
//   smartcore::error::Failed                    src/error/mod.rs:11-37, Display :124-128
//   smartcore::linalg::basic::matrix::DenseMatrix   src/linalg/basic/matrix.rs:27-32,187-237,367-381
//   smartcore::rand_custom::get_rng_impl        src/rand_custom.rs:8-33 (default features -> SmallRng)
//   smartcore::cluster::kmeans::KMeansParameters / KMeansSearchParameters   src/cluster/kmeans.rs:109-232
//   smartcore::cluster::kmeans::KMeans::{fit,predict}                       src/cluster/kmeans.rs:254-352
//
// What stays on the host, exactly as in the reference: parameter validation and its messages
// (kmeans.rs:257-269), and the seeded RNG draw sequence of kmeans_plus_plus (kmeans.rs:355,359,385):
// one gen_range(0..n) then k-1 gen::<f64>() -- none of which depend on the data, so they are drawn
// up front and handed to the device path.  Everything else is device work behind
// include/smartcore_kmeans_cuda.h.  There is no CPU fallback.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstdlib>
#include <optional>
#include <random>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/smartcore_kmeans_cuda.h"
#include "smartcore_persist.hpp"

namespace smartcore {

// ---------------------------------------------------------------------------------------------
namespace error {
enum class FailedError { FitFailed = 1, PredictFailed, TransformFailed, FindFailed, DecompositionFailed,
                         SolutionFailed, ParametersError, InvalidStateError };
struct Failed {
    FailedError err;
    std::string msg;
    static Failed fit(const std::string& m) { return {FailedError::FitFailed, m}; }
    static Failed predict(const std::string& m) { return {FailedError::PredictFailed, m}; }
    static Failed input(const std::string& m) { return {FailedError::ParametersError, m}; }
    // Display: "{kind}: {msg}" (src/error/mod.rs:109-128)
    std::string to_string() const {
        const char* kind = "";
        switch (err) {
            case FailedError::FitFailed: kind = "Fit failed"; break;
            case FailedError::PredictFailed: kind = "Predict failed"; break;
            case FailedError::TransformFailed: kind = "Transform failed"; break;
            case FailedError::FindFailed: kind = "Find failed"; break;
            case FailedError::DecompositionFailed: kind = "Decomposition failed"; break;
            case FailedError::SolutionFailed: kind = "Can't find solution"; break;
            case FailedError::ParametersError: kind = "Error in input, check parameters"; break;
            case FailedError::InvalidStateError: kind = "Invalid state, this should never happen"; break;
        }
        return std::string(kind) + ": " + msg;
    }
};
// minimal Result<T, Failed>
template <typename T> struct Result {
    std::optional<T> value;
    std::optional<Failed> error;
    static Result Ok(T v) { Result r; r.value = std::move(v); return r; }
    static Result Err(Failed f) { Result r; r.error = std::move(f); return r; }
    bool is_ok() const { return value.has_value(); }
    bool is_err() const { return !is_ok(); }
    T& unwrap() { return *value; }
    const Failed& unwrap_err() const { return *error; }
};
}  // namespace error

// ---------------------------------------------------------------------------------------------
namespace linalg { namespace basic { namespace matrix {
// DenseMatrix: contiguous values, column-major by default (from_2d_array), row-major via new_(.., false)
template <typename T> class DenseMatrix {
public:
    size_t ncols = 0, nrows = 0;
    std::vector<T> values;
    bool column_major = true;

    static error::Result<DenseMatrix> new_(size_t nrows, size_t ncols, std::vector<T> values, bool column_major) {
        if (nrows * ncols != values.size())
            return error::Result<DenseMatrix>::Err(error::Failed::input(
                "The specified shape: (cols: " + std::to_string(ncols) + ", rows: " + std::to_string(nrows) +
                ") does not align with data len: " + std::to_string(values.size())));
        DenseMatrix m; m.ncols = ncols; m.nrows = nrows; m.values = std::move(values); m.column_major = column_major;
        return error::Result<DenseMatrix>::Ok(std::move(m));
    }
    static error::Result<DenseMatrix> from_2d_vec(const std::vector<std::vector<T>>& rows) {
        if (rows.empty() || rows[0].empty())
            return error::Result<DenseMatrix>::Err(error::Failed::input("The 2d vec provided is empty; cannot instantiate the matrix"));
        const size_t nrows = rows.size(), ncols = rows[0].size();
        std::vector<T> v; v.reserve(nrows * ncols);
        for (size_t c = 0; c < ncols; c++) for (size_t r = 0; r < nrows; r++) v.push_back(rows[r][c]);
        return new_(nrows, ncols, std::move(v), true);
    }
    std::pair<size_t, size_t> shape() const { return {nrows, ncols}; }
    const T& get(size_t row, size_t col) const {  // matrix.rs:367-381
        return column_major ? values[col * nrows + row] : values[col + ncols * row];
    }
};
}}}  // namespace linalg::basic::matrix

// ---------------------------------------------------------------------------------------------
namespace rand_custom {
// smartcore's RngImpl (src/rand_custom.rs:1-4) is a compile-time choice:
//   default features      -> rand 0.8.5 SmallRng = xoshiro256++ (64-bit targets)
//   feature `std_rand`    -> rand 0.8.5 StdRng   = rand_chacha 0.3 ChaCha12Rng   (forced by `datasets`, Cargo.toml:39-40)
// Both are mirrored here behind a process-wide switch (set_std_rand), so either smartcore build can be followed.
// seed_from_u64 is rand_core 0.6's default for both (PCG32 fill of the 32-byte seed; SmallRng does not forward
// xoshiro's SplitMix override in 0.8.x -- `splitmix = true` follows rand >= 0.9 instead).  gen::<f64>() and
// gen_range(0..n) are generator-independent (Standard: 53 high bits; UniformInt<usize>::sample_single on next_u64).
// (Published algorithms restated; rand is not part of the reference tree -- SURVEY App. B.  Pins: xoshiro256++ upstream
// vector, ChaCha TC1 vectors for 8/12/20 rounds, PCG32 XSH-RR demo vector -- checked by the CPU test suite under tests/.)
inline bool& std_rand_switch() { static bool on = false; return on; }
inline void set_std_rand(bool on) { std_rand_switch() = on; }

inline void pcg32_fill(uint64_t state, uint32_t (&w)[8]) {       // rand_core 0.6 SeedableRng::seed_from_u64
    const uint64_t MUL = 6364136223846793005ULL, INC = 11634580027462260723ULL;
    for (int i = 0; i < 8; i++) {
        state = state * MUL + INC;
        const uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27), rot = (uint32_t)(state >> 59);
        w[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
    }
}

class RngImpl {
public:
    // SmallRng (default) or StdRng (std_rand switch) seeded the way `RngImpl::seed_from_u64` does it
    static RngImpl seed_from_u64(uint64_t state, bool splitmix = false) { return seed_from_u64(state, splitmix, std_rand_switch()); }
    static RngImpl seed_from_u64(uint64_t state, bool splitmix, bool std_rand) {
        RngImpl r;
        r.chacha = std_rand;
        uint32_t w[8];
        if (std_rand) {                                   // StdRng: the PCG32 words are the ChaCha key
            pcg32_fill(state, w);
            for (int i = 0; i < 8; i++) r.key[i] = w[i];
            r.counter = 0; r.index = 16;
            return r;
        }
        if (!splitmix) {
            pcg32_fill(state, w);
            bool all_zero = true;
            for (int i = 0; i < 8; i++) all_zero = all_zero && w[i] == 0;
            if (!all_zero) { for (int i = 0; i < 4; i++) r.s[i] = ((uint64_t)w[2 * i + 1] << 32) | w[2 * i]; return r; }
            state = 0;
        }
        for (int i = 0; i < 4; i++) {
            state += 0x9e3779b97f4a7c15ULL;
            uint64_t z = state;
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
            r.s[i] = z ^ (z >> 31);
        }
        return r;
    }
    uint64_t next_u64() {
        if (chacha) {                                     // BlockRng: 16 words per block, two consecutive words, low first
            if (index >= 16) { block(); index = 0; }
            const uint64_t v = (uint64_t)buf[index] | ((uint64_t)buf[index + 1] << 32);
            index += 2;
            return v;
        }
        auto rotl = [](uint64_t x, int k) { return (x << k) | (x >> (64 - k)); };
        uint64_t result = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return result;
    }
    double gen_f64() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }  // Standard: 53 bits
    uint64_t gen_range(uint64_t high) {  // UniformInt<usize>::sample_single(0, high)
        if (high == 0) return next_u64();
        const uint64_t zone = (high << __builtin_clzll(high)) - 1;
        for (;;) {
            unsigned __int128 m = (unsigned __int128)next_u64() * high;
            if ((uint64_t)m <= zone) return (uint64_t)(m >> 64);
        }
    }
    // ChaCha block function, `rounds` rounds, 64-bit block counter in words 12-13, stream id 0 (rand_chacha 0.3 layout)
    static void chacha_block(const uint32_t (&key)[8], uint64_t counter, int rounds, uint32_t (&out)[16]) {
        const uint32_t st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3],
                                 key[4], key[5], key[6], key[7], (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
        uint32_t x[16];
        for (int i = 0; i < 16; i++) x[i] = st[i];
        auto rotl = [](uint32_t v, int k) { return (v << k) | (v >> (32 - k)); };
        auto qr = [&](int a, int b, int c, int d) {
            x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
            x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
        };
        for (int r = 0; r < rounds; r += 2) {
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
        }
        for (int i = 0; i < 16; i++) out[i] = x[i] + st[i];
    }
    uint64_t s[4] = {0, 0, 0, 0};
    bool chacha = false;
    uint32_t key[8] = {0, 0, 0, 0, 0, 0, 0, 0}, buf[16] = {0};
    uint64_t counter = 0;
    int index = 16;
private:
    void block() { chacha_block(key, counter++, 12, buf); }
};
// get_rng_impl (rand_custom.rs:8-33).  None: seed 0 in the default build (:24-28); in a std_rand build the reference
// seeds from thread_rng (:13-16, unreproducible by design) -- mirrored with the system entropy source.
inline RngImpl get_rng_impl(std::optional<uint64_t> seed) {
    if (seed) return RngImpl::seed_from_u64(*seed);
    if (!std_rand_switch()) return RngImpl::seed_from_u64(0);
    std::random_device rd;
    return RngImpl::seed_from_u64(((uint64_t)rd() << 32) | rd());
}
}  // namespace rand_custom

// ---------------------------------------------------------------------------------------------
namespace cluster { namespace kmeans {

struct KMeansParameters {
    size_t k = 2;           // Default (kmeans.rs:138-146)
    size_t max_iter = 100;
    std::optional<uint64_t> seed;
    KMeansParameters with_k(size_t v) const { KMeansParameters p = *this; p.k = v; return p; }
    KMeansParameters with_max_iter(size_t v) const { KMeansParameters p = *this; p.max_iter = v; return p; }
};

// grid iterator: k fastest, then max_iter, then seed (kmeans.rs:148-232)
struct KMeansSearchParameters {
    std::vector<size_t> k{2};
    std::vector<size_t> max_iter{100};
    std::vector<std::optional<uint64_t>> seed{std::nullopt};
    struct Iterator {
        const KMeansSearchParameters* sp; size_t ik = 0, im = 0, is = 0;
        std::optional<KMeansParameters> next() {
            if (ik == sp->k.size() && im == sp->max_iter.size() && is == sp->seed.size()) return std::nullopt;
            KMeansParameters p; p.k = sp->k[ik]; p.max_iter = sp->max_iter[im]; p.seed = sp->seed[is];
            if (ik + 1 < sp->k.size()) ik++;
            else if (im + 1 < sp->max_iter.size()) { ik = 0; im++; }
            else if (is + 1 < sp->seed.size()) { ik = 0; im = 0; is++; }
            else { ik++; im++; is++; }
            return p;
        }
    };
    Iterator into_iter() const { return Iterator{this}; }
};

// Process-wide device context used by KMeans: ONE context over every visible GPU (sckm_ctx_create_multi), so that a
// plain `KMeans::fit(&x, params)` shards its rows over the whole box with no change to the caller (kmeans.rs:254).
// SMARTCORE_CUDA_DEVICES="0,2,3" restricts / orders the devices ("0": single GPU).
class Device {
public:
    static error::Result<sckm_ctx*> get() {
        static sckm_ctx* ctx = nullptr;
        if (!ctx) {
            std::vector<int> ids;
            if (const char* e = std::getenv("SMARTCORE_CUDA_DEVICES")) {
                int v = 0; bool have = false;
                for (const char* p = e;; p++) {
                    if (*p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); have = true; }
                    else { if (have) ids.push_back(v); v = 0; have = false; if (!*p) break; }
                }
            }
            int rc = sckm_ctx_create_multi((int)ids.size(), ids.empty() ? nullptr : ids.data(), &ctx);
            if (rc != SCKM_OK) return error::Result<sckm_ctx*>::Err(error::Failed::fit(std::string("CUDA backend unavailable: ") + sckm_last_error(nullptr)));
        }
        return error::Result<sckm_ctx*>::Ok(ctx);
    }
};

template <typename TX> struct Packed {  // X as the C ABI wants it: f32/f64 pass through, other TX widen to f64
    const void* ptr; int dtype; int column_major; std::vector<double> widened;
};
template <typename TX> inline void pack(const linalg::basic::matrix::DenseMatrix<TX>& x, Packed<TX>& p) {
    p.column_major = x.column_major ? 1 : 0;
    if constexpr (std::is_same<TX, float>::value) { p.ptr = x.values.data(); p.dtype = SCKM_F32; }
    else if constexpr (std::is_same<TX, double>::value) { p.ptr = x.values.data(); p.dtype = SCKM_F64; }
    else { p.widened.assign(x.values.begin(), x.values.end()); p.ptr = p.widened.data(); p.dtype = SCKM_F64; }  // Number::to_f64
}

template <typename TX, typename TY> class KMeans {
public:
    using X = linalg::basic::matrix::DenseMatrix<TX>;
    size_t k = 0;
    std::vector<size_t> _y;
    std::vector<size_t> size;
    double _distortion = 0.0;
    std::vector<std::vector<double>> centroids;
    int64_t _iterations = 0;  // extra: number of clustering steps executed (not in the reference struct)

    // KMeans::fit (kmeans.rs:254-323)
    static error::Result<KMeans> fit(const X& data, const KMeansParameters& parameters) {
        using R = error::Result<KMeans>;
        if (parameters.k < 2)
            return R::Err(error::Failed::fit("invalid number of clusters: " + std::to_string(parameters.k)));
        if (parameters.max_iter == 0)
            return R::Err(error::Failed::fit("invalid maximum number of iterations: " + std::to_string(parameters.max_iter)));
        const size_t n = data.nrows, d = data.ncols;
        // kmeans_plus_plus's RNG draws, in the reference order (kmeans.rs:355,359,385)
        auto rng = rand_custom::get_rng_impl(parameters.seed);
        const uint64_t first = rng.gen_range((uint64_t)n);
        std::vector<double> uniforms(parameters.k - 1);
        for (auto& u : uniforms) u = rng.gen_f64();

        auto dev = Device::get();
        if (dev.is_err()) return R::Err(dev.unwrap_err());
        sckm_ctx* ctx = dev.unwrap();
        Packed<TX> p; pack(data, p);
        KMeans m; m.k = parameters.k;
        std::vector<uint64_t> y(n); std::vector<int64_t> sz(parameters.k); std::vector<double> c(parameters.k * d);
        int rc = sckm_kmeans_fit(ctx, p.ptr, n, d, p.dtype, p.column_major, parameters.k, parameters.max_iter, first,
                                 uniforms.data(), y.data(), 8, sz.data(), c.data(), &m._distortion, &m._iterations);
        if (rc != SCKM_OK) return R::Err(error::Failed::fit(sckm_last_error(ctx)));
        m._y.assign(y.begin(), y.end());
        m.size.assign(sz.begin(), sz.end());
        m.centroids.resize(parameters.k);
        for (size_t i = 0; i < parameters.k; i++) m.centroids[i].assign(c.begin() + i * d, c.begin() + (i + 1) * d);
        return R::Ok(std::move(m));
    }

    // KMeans::predict (kmeans.rs:327-352); labels converted with TY::from_usize
    error::Result<std::vector<TY>> predict(const X& x) const {
        using R = error::Result<std::vector<TY>>;
        const size_t n = x.nrows, d = x.ncols;
        if (centroids.size() != k) return R::Err(error::Failed::predict("model holds " + std::to_string(centroids.size()) + " centroids, k = " + std::to_string(k)));
        for (const auto& row : centroids)
            if (row.size() != d)
                return R::Err(error::Failed::predict("Input vector sizes are different."));  // reference panics (euclidian.rs:52-54)
        auto dev = Device::get();
        if (dev.is_err()) return R::Err(error::Failed::predict(dev.unwrap_err().msg));
        sckm_ctx* ctx = dev.unwrap();
        Packed<TX> p; pack(x, p);
        std::vector<double> c(k * d);
        for (size_t i = 0; i < k; i++) for (size_t j = 0; j < d; j++) c[i * d + j] = centroids[i][j];
        std::vector<uint32_t> lab(n);
        int rc = sckm_predict(ctx, p.ptr, n, d, p.dtype, p.column_major, c.data(), k, lab.data(), 4);
        if (rc != SCKM_OK) return R::Err(error::Failed::predict(sckm_last_error(ctx)));
        std::vector<TY> out(n);
        for (size_t i = 0; i < n; i++) out[i] = (TY)lab[i];
        return R::Ok(std::move(out));
    }

    // serde (kmeans.rs:70-83): serde_json and bincode images interchangeable with the reference's derive
    persist::KMeansFields fields() const {
        persist::KMeansFields f;
        f.k = k; f.y.assign(_y.begin(), _y.end()); f.size.assign(size.begin(), size.end());
        f.distortion = _distortion; f.centroids = centroids;
        return f;
    }
    static KMeans from_fields(const persist::KMeansFields& f) {
        KMeans m;
        m.k = (size_t)f.k; m._y.assign(f.y.begin(), f.y.end()); m.size.assign(f.size.begin(), f.size.end());
        m._distortion = f.distortion; m.centroids = f.centroids;
        return m;
    }
    std::string to_json() const { return persist::to_json(fields()); }
    std::string to_bincode() const { return persist::to_bincode(fields()); }
    static error::Result<KMeans> from_json(const std::string& text) {
        persist::KMeansFields f; std::string e;
        if (!persist::from_json(text, f, e)) return error::Result<KMeans>::Err(error::Failed::input(e));
        return error::Result<KMeans>::Ok(from_fields(f));
    }
    static error::Result<KMeans> from_bincode(const std::string& bytes) {
        persist::KMeansFields f; std::string e;
        if (!persist::from_bincode(bytes, f, e)) return error::Result<KMeans>::Err(error::Failed::input(e));
        return error::Result<KMeans>::Ok(from_fields(f));
    }

    // PartialEq (kmeans.rs:85-107)
    bool operator==(const KMeans& o) const {
        if (k != o.k || size != o.size || centroids.size() != o.centroids.size()) return false;
        for (size_t i = 0; i < centroids.size(); i++) {
            if (centroids[i].size() != o.centroids[i].size()) return false;
            for (size_t j = 0; j < centroids[i].size(); j++) {
                double diff = centroids[i][j] - o.centroids[i][j];
                if ((diff < 0 ? -diff : diff) > 2.220446049250313e-16) return false;
            }
        }
        return true;
    }
};

}}  // namespace cluster::kmeans
}  // namespace smartcore
