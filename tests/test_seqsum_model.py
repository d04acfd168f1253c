"""The reference's kmeans++ pick is a sequential f64 prefix scan (kmeans.rs:385-398).  (1) How often does a BLOCKED prefix
(what the CUDA path computes today) pick a different row?  (2) The block-parallel evaluation of the SEQUENTIAL rounding
prototyped in oracle/seqsum_np.py reproduces numpy's sequential cumsum bit for bit."""
import numpy as np
import pytest

from oracle import seqsum_np


@pytest.mark.parametrize("kind", ["gamma", "wide", "with_zeros", "tiny", "ties"])
def test_binade_block_sums_reproduce_the_sequential_prefix(kind):
    rng = np.random.default_rng(5)
    n = 200_000
    if kind == "gamma":
        d = rng.gamma(8.0, 4.0, size=n)
    elif kind == "wide":
        d = rng.gamma(2.0, 1.0, size=n) * 10.0 ** rng.uniform(-12, 6, size=n)
    elif kind == "with_zeros":
        d = rng.gamma(8.0, 4.0, size=n)
        d[rng.random(n) < 0.3] = 0.0                      # rows that ARE seeds have D^2 = 0
        d[:5000] = 0.0
    elif kind == "tiny":
        d = rng.gamma(8.0, 4.0, size=n) * 1e-300
    else:
        d = np.ldexp(rng.integers(1, 2 ** 20, size=n).astype(np.float64), -10)   # short mantissas: exact ties do occur
    want = np.cumsum(d)                                   # numpy's cumsum is the sequential loop
    got, fallbacks = seqsum_np.block_prefixes(d, 1024)
    nb = (n + 1023) // 1024
    ends = np.minimum(np.arange(1, nb + 1) * 1024, n) - 1
    assert np.array_equal(got[1:], want[ends]), "block-end prefixes differ from the sequential sum"
    if kind in ("gamma", "with_zeros"):
        assert fallbacks <= 40, "%d of %d blocks walked sequentially" % (fallbacks, nb)


def test_blocked_prefix_pick_differs_rarely_but_measurably():
    """Documents the size of the effect quoted in DESIGN.md: mean |blocked prefix - sequential prefix| in units of the mean
    element, i.e. the chance per draw that the two scans stop at different rows."""
    rng = np.random.default_rng(0)
    n = 2_000_000
    d = rng.gamma(8.0, 4.0, size=n)
    seq = np.cumsum(d)
    nb = (n + 1023) // 1024
    pad = np.zeros(nb * 1024)
    pad[:n] = d
    within = np.cumsum(pad.reshape(nb, 1024), axis=1)
    bpre = np.concatenate([[0.0], np.cumsum(within[:, -1])[:-1]])
    blk = (bpre[:, None] + within).reshape(-1)[:n]
    p = float(np.abs(blk - seq).mean() / d.mean())
    assert 1e-10 < p < 1e-6, p                              # ~2e-8 at 2M rows; grows like n^2: ~5e-7 at 1e7, ~1e-5 at 5e7
