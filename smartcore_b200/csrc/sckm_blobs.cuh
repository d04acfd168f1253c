// sckm_blobs.cuh -- counter-based synthetic Gaussian-blob generator (host + device twins).
//
// Follows the recipe of smartcore's make_blobs (src/dataset/generator.rs:10-48):
// `n_centers` true centres with every coordinate ~ U[-10, 10); point i belongs to centre
// i % n_centers; coordinate = centre + N(0, 1).  The reference draws from thread_rng()
// (unreproducible), so the stream itself is ours: Philox-4x32-10 keyed by the data seed with
// counter (row, col, stream).  The normal deviate is Irwin-Hall-12 over exact 21-bit dyadic
// uniforms (integer sum, one exact conversion), so host and device produce identical bits and
// any row can be regenerated anywhere (shard independent).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SCKM_HD __host__ __device__ __forceinline__
#else
#define SCKM_HD inline
#endif

namespace sckm {

struct Philox4 { uint32_t v[4]; };

SCKM_HD void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32); lo = (uint32_t)p;
}

SCKM_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(M0, c0, hi0, lo0);
        mulhilo32(M1, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3; return o;
}

// sum of six 21-bit uniforms out of one Philox block
SCKM_HD uint32_t six21(const Philox4& p) {
    uint64_t a = ((uint64_t)p.v[1] << 32) | p.v[0], b = ((uint64_t)p.v[3] << 32) | p.v[2];
    const uint64_t M = 0x1FFFFFull;
    return (uint32_t)((a & M) + ((a >> 21) & M) + ((a >> 42) & M) + (b & M) + ((b >> 21) & M) + ((b >> 42) & M));
}

SCKM_HD double blob_center(uint64_t seed, uint64_t center, uint64_t col) {
    Philox4 p = philox4x32_10((uint32_t)center, (uint32_t)(center >> 32), (uint32_t)col, 2u,
                              (uint32_t)seed, (uint32_t)(seed >> 32));
    double u = (double)(p.v[0] >> 8) * (1.0 / 16777216.0);  // [0,1), 24 bits
    return -10.0 + 20.0 * u;
}

// value of X[row][col] in f64 (cast to f32 by the caller when the dataset is f32)
SCKM_HD double blob_value(uint64_t seed, uint64_t n_centers, uint64_t row, uint64_t col) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    Philox4 p0 = philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)col, 0u, k0, k1);
    Philox4 p1 = philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)col, 1u, k0, k1);
    uint32_t s = six21(p0) + six21(p1);                 // < 12 * 2^21
    double z = ((double)s - 12582912.0) * (1.0 / 2097152.0);  // (s - 6*2^21) / 2^21: mean 0, var 1
    return blob_center(seed, row % n_centers, col) + z;
}

}  // namespace sckm
