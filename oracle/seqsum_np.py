"""TEST INFRASTRUCTURE ONLY -- prototype of a PARALLEL evaluation of the reference's sequential D^2 prefix.

kmeans++ in the reference picks the next seed with a strictly sequential f64 scan (`sum += d[i]` ... `cost += d[index];
if cost >= cutoff`, /root/reference/src/cluster/kmeans.rs:385-398).  The CUDA path sums D^2 in 1024-row blocks, which is
the same real number but not the same rounding: a numpy model (tests/test_seqsum_model.py) puts the chance of picking a
different row at ~1e-6 per draw for n = 1e7 and growing like n^2.  This file shows how to get the SEQUENTIAL rounding
from block-parallel work, for the kernel that should replace the blocked sums (known gaps):

  while the running sum s stays inside one binade [2^e, 2^(e+1)), its ulp U = 2^(e-52) is fixed and s is a multiple of
  U, so fl(s + d) = s + U * c(d) with c(d) = floor(d/U) + [frac(d/U) > 1/2] -- an INTEGER that does not depend on s
  (an exact tie, frac == 1/2, rounds to even and does depend on s: such blocks are walked sequentially).  A block's
  contribution is therefore U * sum(c(d_i)): an exact, order-free integer sum, valid if the binade assumed for the block
  (known beforehand from the approximate blocked prefix) is the binade of s at the block's start AND end.

Blocks that straddle a binade boundary, contain a tie, or start from s = 0 fall back to the sequential walk: a handful
per pass.  `block_prefixes` returns the sequential prefix at every block start, bit-identical to numpy's cumsum (which is
sequential), and how many blocks needed the fallback."""
import numpy as np


def _binade(s):
    """e with 2^e <= s < 2^(e+1) for a positive normal double"""
    return int(np.frexp(s)[1]) - 1


def block_prefixes(d, block=1024):
    d = np.asarray(d, dtype=np.float64)
    n = len(d)
    nb = (n + block - 1) // block
    # what the device has before the scan: blocked sums and their prefix (approximate, relative error ~1e-13)
    pad = np.zeros(nb * block)
    pad[:n] = d
    bsum = pad.reshape(nb, block).sum(axis=1)
    approx_start = np.concatenate([[0.0], np.cumsum(bsum)[:-1]])
    out = np.empty(nb + 1)
    s = 0.0
    fallbacks = 0
    for b in range(nb):
        out[b] = s
        blk = pad[b * block:(b + 1) * block]
        ok = False
        a = approx_start[b]
        if s > 0.0 and a > 0.0 and np.isfinite(a):
            e = _binade(a)                                   # binade GUESSED from the approximate prefix (parallel side)
            if e > -1000:
                U = np.ldexp(1.0, e - 52)
                x = blk / U                                   # exact: a power-of-two scaling
                q = np.floor(x)
                r = x - q                                     # exact fraction
                if not np.any(r == 0.5) and np.all(x < 2.0 ** 53):
                    Q = int(q.astype(np.int64).sum() + np.count_nonzero(r > 0.5))
                    # the scan side checks the guess: s must lie in that binade at the start and at the end of the block
                    if _binade(s) == e:
                        t = s + Q * U                         # exact: both are multiples of U below 2^(e+1) when the test passes
                        if Q < 2 ** 53 and t < np.ldexp(1.0, e + 1):
                            s = t
                            ok = True
        if not ok:
            fallbacks += 1
            for v in blk:
                s = s + v
    out[nb] = s
    return out, fallbacks
