// smartcore_persist.hpp -- the on-disk forms of a fitted KMeans, kept compatible with what the reference's
// `#[derive(Serialize, Deserialize)]` produces (src/cluster/kmeans.rs:70-83), so
// that a model fitted here can be loaded by the reference and vice versa:
//
//   struct KMeans { k: usize, _y: Vec<usize>, size: Vec<usize>, _distortion: f64, centroids: Vec<Vec<f64>>,
//                   _phantom_tx/_phantom_ty/_phantom_x/_phantom_y: PhantomData }
//
//   * serde_json (the reference's own round-trip test, kmeans.rs:538-544): one object, fields in declaration order,
//     PhantomData as `null`, non-finite f64 as `null` (serde_json's rule); the reader accepts any field order and
//     ignores fields it does not know.
//   * bincode 1.3 default options (dev-dependency, Cargo.toml:53): little-endian, usize as u64, Vec = u64 length +
//     elements, f64 = 8 raw bytes, PhantomData = nothing.
//
// Host-only code: no device work, usable without a GPU.
#pragma once
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace smartcore { namespace persist {

struct KMeansFields {
    uint64_t k = 0;
    std::vector<uint64_t> y;
    std::vector<uint64_t> size;
    double distortion = 0.0;
    std::vector<std::vector<double>> centroids;
};

// ---- serde_json ------------------------------------------------------------------------------------------------
inline void json_f64(std::string& out, double v) {
    if (!std::isfinite(v)) { out += "null"; return; }
    char buf[40];
    auto r = std::to_chars(buf, buf + sizeof buf, v);            // shortest form that round-trips
    std::string s(buf, r.ptr);
    if (s.find_first_of(".eEn") == std::string::npos) s += ".0";   // serde_json always writes floats as floats
    out += s;
}
inline void json_u64s(std::string& out, const std::vector<uint64_t>& v) {
    out += '[';
    for (size_t i = 0; i < v.size(); i++) { if (i) out += ','; out += std::to_string(v[i]); }
    out += ']';
}
inline std::string to_json(const KMeansFields& m) {
    std::string o;
    o.reserve(64 + m.y.size() * 3 + m.centroids.size() * 32);
    o += "{\"k\":" + std::to_string(m.k) + ",\"_y\":";
    json_u64s(o, m.y);
    o += ",\"size\":";
    json_u64s(o, m.size);
    o += ",\"_distortion\":";
    json_f64(o, m.distortion);
    o += ",\"centroids\":[";
    for (size_t i = 0; i < m.centroids.size(); i++) {
        if (i) o += ',';
        o += '[';
        for (size_t j = 0; j < m.centroids[i].size(); j++) { if (j) o += ','; json_f64(o, m.centroids[i][j]); }
        o += ']';
    }
    o += "],\"_phantom_tx\":null,\"_phantom_ty\":null,\"_phantom_x\":null,\"_phantom_y\":null}";
    return o;
}

// minimal JSON reader for that schema
struct JsonReader {
    const char* p; const char* end; std::string err;
    explicit JsonReader(const std::string& s) : p(s.data()), end(s.data() + s.size()) {}
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
    bool fail(const std::string& m) { if (err.empty()) err = m; return false; }
    bool lit(char c) { ws(); if (p < end && *p == c) { p++; return true; } return fail(std::string("expected '") + c + "'"); }
    bool peek(char c) { ws(); return p < end && *p == c; }
    bool null_() { ws(); if (end - p >= 4 && !memcmp(p, "null", 4)) { p += 4; return true; } return false; }
    bool str(std::string& out) {
        if (!lit('"')) return false;
        out.clear();
        while (p < end && *p != '"') { if (*p == '\\' && p + 1 < end) p++; out += *p++; }
        return lit('"');
    }
    bool u64(uint64_t& v) {
        ws();
        auto r = std::from_chars(p, end, v);
        if (r.ec != std::errc()) return fail("expected an unsigned integer");
        p = r.ptr; return true;
    }
    bool f64(double& v) {
        if (null_()) { v = std::numeric_limits<double>::quiet_NaN(); return true; }
        ws();
        auto r = std::from_chars(p, end, v);
        if (r.ec != std::errc()) return fail("expected a number");
        p = r.ptr; return true;
    }
    template <typename F> bool array(F&& item) {
        if (!lit('[')) return false;
        if (peek(']')) { p++; return true; }
        for (;;) {
            if (!item()) return false;
            ws();
            if (p < end && *p == ',') { p++; continue; }
            return lit(']');
        }
    }
    bool skip() {                                                   // any JSON value
        ws();
        if (p >= end) return fail("unexpected end");
        if (*p == '"') { std::string s; return str(s); }
        if (*p == '[') return array([&] { return skip(); });
        if (*p == '{') {
            p++;
            if (peek('}')) { p++; return true; }
            for (;;) {
                std::string key;
                if (!str(key) || !lit(':') || !skip()) return false;
                ws();
                if (p < end && *p == ',') { p++; continue; }
                return lit('}');
            }
        }
        while (p < end && *p != ',' && *p != ']' && *p != '}' && *p != ' ' && *p != '\n') p++;
        return true;
    }
};

// Shape invariants every image written by the reference's derive(Serialize) satisfies (kmeans.rs:73-83): `size` and
// `centroids` have k entries, the centroid rows share one length, every label is a cluster index.  The Rust reference
// panics safely when a hand-edited image breaks them (index out of bounds); here the loaders refuse such an image,
// because the accessors and predict() size their buffers from k and centroids[0].
inline bool validate(const KMeansFields& m, std::string& err) {
    if (m.centroids.size() != m.k) { err = "field `centroids` has " + std::to_string(m.centroids.size()) + " rows, k = " + std::to_string(m.k); return false; }
    if (m.size.size() != m.k) { err = "field `size` has " + std::to_string(m.size.size()) + " entries, k = " + std::to_string(m.k); return false; }
    for (const auto& row : m.centroids)
        if (row.size() != m.centroids[0].size()) { err = "field `centroids` is ragged"; return false; }
    for (uint64_t v : m.y)
        if (v >= m.k) { err = "field `_y` holds label " + std::to_string(v) + " >= k = " + std::to_string(m.k); return false; }
    return true;
}

inline bool from_json(const std::string& text, KMeansFields& m, std::string& err) {
    JsonReader r(text);
    bool have_k = false, have_c = false;
    auto body = [&]() -> bool {
        if (!r.lit('{')) return false;
        if (r.peek('}')) { r.p++; return true; }
        for (;;) {
            std::string key;
            if (!r.str(key) || !r.lit(':')) return false;
            bool ok;
            if (key == "k") { ok = r.u64(m.k); have_k = true; }
            else if (key == "_y") { m.y.clear(); ok = r.array([&] { uint64_t v; if (!r.u64(v)) return false; m.y.push_back(v); return true; }); }
            else if (key == "size") { m.size.clear(); ok = r.array([&] { uint64_t v; if (!r.u64(v)) return false; m.size.push_back(v); return true; }); }
            else if (key == "_distortion") ok = r.f64(m.distortion);
            else if (key == "centroids") {
                m.centroids.clear(); have_c = true;
                ok = r.array([&] {
                    m.centroids.emplace_back();
                    return r.array([&] { double v; if (!r.f64(v)) return false; m.centroids.back().push_back(v); return true; });
                });
            } else ok = r.skip();
            if (!ok) return false;
            r.ws();
            if (r.p < r.end && *r.p == ',') { r.p++; continue; }
            return r.lit('}');
        }
    };
    if (!body()) { err = "invalid KMeans JSON: " + r.err; return false; }
    if (!have_k || !have_c) { err = "invalid KMeans JSON: missing field `k` or `centroids`"; return false; }
    std::string why;
    if (!validate(m, why)) { err = "invalid KMeans JSON: " + why; return false; }
    return true;
}

// ---- bincode 1.3, default options ----------------------------------------------------------------------------------
inline void put_u64(std::string& o, uint64_t v) { for (int i = 0; i < 8; i++) o += (char)((v >> (8 * i)) & 0xff); }
inline void put_f64(std::string& o, double v) { uint64_t b; memcpy(&b, &v, 8); put_u64(o, b); }
inline std::string to_bincode(const KMeansFields& m) {
    std::string o;
    o.reserve(40 + 8 * (m.y.size() + m.size.size()));
    put_u64(o, m.k);
    put_u64(o, m.y.size()); for (uint64_t v : m.y) put_u64(o, v);
    put_u64(o, m.size.size()); for (uint64_t v : m.size) put_u64(o, v);
    put_f64(o, m.distortion);
    put_u64(o, m.centroids.size());
    for (auto& row : m.centroids) { put_u64(o, row.size()); for (double v : row) put_f64(o, v); }
    return o;                                                      // PhantomData fields occupy no bytes
}
inline bool from_bincode(const std::string& bytes, KMeansFields& m, std::string& err) {
    size_t pos = 0;
    auto get = [&](uint64_t& v) -> bool {
        if (bytes.size() - pos < 8) return false;
        v = 0;
        for (int i = 0; i < 8; i++) v |= (uint64_t)(unsigned char)bytes[pos + i] << (8 * i);
        pos += 8; return true;
    };
    auto getf = [&](double& v) -> bool { uint64_t b; if (!get(b)) return false; memcpy(&v, &b, 8); return true; };
    auto vec = [&](std::vector<uint64_t>& out) -> bool {
        uint64_t len;
        if (!get(len) || len > (bytes.size() - pos) / 8) return false;
        out.resize(len);
        for (auto& v : out) if (!get(v)) return false;
        return true;
    };
    uint64_t rows = 0;
    bool ok = get(m.k) && vec(m.y) && vec(m.size) && getf(m.distortion) && get(rows) && rows <= (bytes.size() - pos) / 8;
    if (ok) {
        m.centroids.assign(rows, {});
        for (auto& row : m.centroids) {
            uint64_t len;
            if (!get(len) || len > (bytes.size() - pos) / 8) { ok = false; break; }
            row.resize(len);
            for (auto& v : row) if (!getf(v)) { ok = false; break; }
            if (!ok) break;
        }
    }
    if (!ok || pos != bytes.size()) { err = "invalid KMeans bincode image (truncated or trailing bytes)"; return false; }
    std::string why;
    if (!validate(m, why)) { err = "invalid KMeans bincode image: " + why; return false; }
    return true;
}

}}  // namespace smartcore::persist
