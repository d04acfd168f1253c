"""TEST / BASELINE INFRASTRUCTURE ONLY -- numpy restatement of the synthetic Gaussian-blob generator
(smartcore_b200/csrc/sckm_blobs.cuh: Philox-4x32-10 counters + Irwin-Hall-12 over 21-bit uniforms, recipe of
smartcore's make_blobs, /root/reference/src/dataset/generator.rs:10-48).

It exists so that `bench.py --impl reference` and the cpu_baseline leg can build their input WITHOUT loading the
product library, and it doubles as an independent check of the generator (tests/test_oracle.py compares it with
sckm_blobs_fill_host bit for bit).  Nothing under smartcore_b200/ imports it.
"""
import numpy as np

_M0, _M1, _W0, _W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
_LO = np.uint64(0xFFFFFFFF)


def _philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over equally shaped uint64 arrays holding 32-bit words."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) for v in (c0, c1, c2, c3))
    k0, k1 = int(k0), int(k1)
    for _ in range(10):
        p0 = _M0 * c0                      # 32 x 32 -> 64 bit products fit uint64
        p1 = _M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & _LO, p1 >> np.uint64(32), p1 & _LO
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _six21(v0, v1, v2, v3):
    a = (v1 << np.uint64(32)) | v0
    b = (v3 << np.uint64(32)) | v2
    m = np.uint64(0x1FFFFF)
    s = (a & m) + ((a >> np.uint64(21)) & m) + ((a >> np.uint64(42)) & m) + (b & m) + ((b >> np.uint64(21)) & m) + ((b >> np.uint64(42)) & m)
    return s


def blobs(row0, nrows, d, n_centers, seed, dtype=np.float64, chunk=1 << 16):
    """Rows [row0, row0 + nrows) of the n x d blob matrix with data seed `seed`; row i belongs to centre i % n_centers."""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    cols = np.arange(d, dtype=np.uint64)[None, :]
    cen = np.arange(n_centers, dtype=np.uint64)[:, None] + np.zeros((1, d), dtype=np.uint64)
    v0, _, _, _ = _philox4x32_10(cen & _LO, cen >> np.uint64(32), cols + np.zeros_like(cen), np.full_like(cen, 2), k0, k1)
    centers = -10.0 + 20.0 * ((v0 >> np.uint64(8)).astype(np.float64) * (1.0 / 16777216.0))
    out = np.empty((nrows, d), dtype=dtype)
    for lo in range(0, nrows, chunk):
        hi = min(nrows, lo + chunk)
        rows = (np.arange(row0 + lo, row0 + hi, dtype=np.uint64))[:, None] + np.zeros((1, d), dtype=np.uint64)
        cc = cols + np.zeros_like(rows)
        a = _philox4x32_10(rows & _LO, rows >> np.uint64(32), cc, np.zeros_like(rows), k0, k1)
        b = _philox4x32_10(rows & _LO, rows >> np.uint64(32), cc, np.ones_like(rows), k0, k1)
        s = (_six21(*a) + _six21(*b)).astype(np.float64)
        z = (s - 12582912.0) * (1.0 / 2097152.0)
        out[lo:hi] = (centers[(rows[:, 0] % np.uint64(n_centers)).astype(np.int64)] + z).astype(dtype)
    return out
