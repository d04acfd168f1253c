"""CPU model of the ranking arithmetic of the 3xFP16 tcgen05 kernel (smartcore_b200/csrc/sckm_tc5h.cu), in numpy.

The kernel only RANKS with reduced precision; every row whose best / second gap is within H_TIE_REL * (||x||^2 + max||c||^2)
is re-decided exactly.  That contract holds if the error of a ranked score stays well inside the margin.  This file restates
the kernel's operand preparation -- power-of-two scaling of every row (from its own largest element, floored at 2^-20 of the
centroid scale) and of the centroids (from max||c||^2), the FP16 hi / lo split, the three products Xh.Ch + Xh.Cl + Xl.Ch, the
BF16 three-piece norm term -- and measures the error of the resulting score against the exact x.c - ||c||^2/2 on data of very
different magnitudes.  FP32 accumulation is modelled by rounding the running sum to f32 after every term (the tensor core keeps
at least that much).  No GPU needed: it pins the NUMBERS the margin was chosen from."""
import numpy as np
import pytest

H_TIE_REL = 2e-5          # sckm_tc5h.cu
H_XMAX_EXP = 13
H_ROW_FLOOR = 20


def centroid_exp(cmax):
    """h_centroid_exp: exponent m with 2^m >= sqrt(cmax), clamped to [-100, 100]; 0 when there is nothing to scale by."""
    if not (cmax > 0.0) or not np.isfinite(cmax):
        return 0
    ec = int(np.floor(np.log2(cmax)))            # cmax in [2^ec, 2^(ec+1))
    return int(np.clip((ec + 2) >> 1, -100, 100))


def bf16_round(v):
    """round-to-nearest-even to bfloat16, returned as float32"""
    u = np.asarray(v, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def ranked_scores(x, cent):
    """Scores of every (row, centroid) pair the way the kernel's tensor cores produce them, taken back to real units."""
    x = np.asarray(x, dtype=np.float32)
    cent = np.asarray(cent, dtype=np.float64)
    cnorm = (cent ** 2).sum(1)
    m_c = centroid_exp(float(cnorm.max()))
    s_c = H_XMAX_EXP - m_c
    # centroids: f32 copy, scaled, split
    cs = (cent.astype(np.float32) * np.float32(2.0 ** s_c)).astype(np.float32)
    ch = cs.astype(np.float16)
    cl = (cs - ch.astype(np.float32)).astype(np.float16)
    # rows: per-row scale from the row's own largest element
    mx = np.abs(x).max(1)
    with np.errstate(divide="ignore"):
        e_raw = np.where(mx > 0, np.floor(np.log2(np.maximum(mx, np.finfo(np.float32).tiny))), -127).astype(np.int64)
    s_row = H_XMAX_EXP - np.maximum(e_raw, m_c - H_ROW_FLOOR)
    assert np.all(s_row <= 126), "this model does not cover the forced-exact rows"
    xs = (x * np.exp2(s_row.astype(np.float64))[:, None].astype(np.float32)).astype(np.float32)
    xh = xs.astype(np.float16)
    xl = (xs - xh.astype(np.float32)).astype(np.float16)
    assert np.all(np.isfinite(xh.astype(np.float32))) and np.all(np.isfinite(ch.astype(np.float32)))
    # norm term: f32 value, three BF16 pieces, times the row's power of two (exact)
    v = (-0.5 * cnorm * 2.0 ** s_c).astype(np.float32)
    h1 = bf16_round(v); h2 = bf16_round(v - h1); h3 = bf16_round(v - h1 - h2)
    p = np.exp2(s_row.astype(np.float64)).astype(np.float32)
    # FP32 accumulation, one rounding per term: products of FP16 values are exact in f32
    acc = np.zeros((x.shape[0], cent.shape[0]), dtype=np.float32)
    A = [xh.astype(np.float32), xh.astype(np.float32), xl.astype(np.float32)]
    B = [ch.astype(np.float32), cl.astype(np.float32), ch.astype(np.float32)]
    for a, b in zip(A, B):
        for j in range(x.shape[1]):
            acc = (acc + a[:, j:j + 1] * b[None, :, j].reshape(1, -1)).astype(np.float32)
    for h in (h1, h2, h3):
        acc = (acc + p[:, None] * h[None, :]).astype(np.float32)
    rs = np.exp2(-(s_row + s_c).astype(np.float64))
    return acc.astype(np.float64) * rs[:, None]


def exact_scores(x, cent):
    x = np.asarray(x, dtype=np.float32).astype(np.float64)
    cent = np.asarray(cent, dtype=np.float64)
    return x @ cent.T - 0.5 * (cent ** 2).sum(1)[None, :]


CASES = ["blobs", "row_scales", "tiny_rows", "huge_scale", "tiny_scale", "offset", "one_big_feature", "sparse_rows"]


@pytest.mark.parametrize("case", CASES)
def test_ranked_score_error_is_far_inside_the_tie_margin(case):
    rng = np.random.default_rng(7)
    n, d, k = 600, 32, 160
    centers = rng.uniform(-10, 10, size=(k, d))
    x = centers[rng.integers(0, k, size=n)] + rng.normal(size=(n, d))
    if case == "row_scales":
        x *= 10.0 ** rng.uniform(-3, 3, size=(n, 1))
    elif case == "tiny_rows":
        x[::2] *= 1e-12
    elif case == "huge_scale":
        x *= 1e15
    elif case == "tiny_scale":
        x *= 1e-18
    elif case == "offset":
        x += 300.0
    elif case == "one_big_feature":
        x[:, 3] *= 1e4
    elif case == "sparse_rows":
        x[rng.random(size=x.shape) < 0.8] = 0.0
    x = x.astype(np.float32)
    cent = x[rng.choice(n, k, replace=False)].astype(np.float64) * (1 + 1e-3)
    got, want = ranked_scores(x, cent), exact_scores(x, cent)
    xn = (x.astype(np.float64) ** 2).sum(1)
    cmax = (cent ** 2).sum(1).max()
    rel = np.abs(got - want) / (xn + cmax)[:, None]
    # a decision compares two scores and the tie test works on 2 * (best - second): four score errors must fit the margin
    assert rel.max() <= H_TIE_REL / 8, "score error %.3g of (||x||^2 + max||c||^2): margin %.3g" % (rel.max(), H_TIE_REL)


def test_scaled_operands_stay_inside_fp16():
    """Row scaling keeps every scaled row below 2^14 and the centroids below 2^13 for centroid magnitudes between 2^-100 and
    2^100 (beyond that the exponent is clamped, scaled values may leave FP16's range, the scores turn non-finite and the
    tie test hands the row to the exact re-decision: slow, still correct -- the GPU tests run data at 1e30)."""
    rng = np.random.default_rng(3)
    for scale in (1e-25, 1e-6, 1.0, 1e9, 1e25):
        x = (rng.normal(size=(200, 32)) * scale).astype(np.float32)
        cent = x[:40].astype(np.float64)
        cnorm = (cent ** 2).sum(1)
        m_c = centroid_exp(float(cnorm.max()))
        assert np.abs(cent).max() <= 2.0 ** m_c * (1 + 1e-12)
        assert np.abs(cent.astype(np.float32) * np.float32(2.0 ** (H_XMAX_EXP - m_c))).max() <= 2.0 ** H_XMAX_EXP
        mx = np.abs(x).max(1)
        e_raw = np.floor(np.log2(mx)).astype(np.int64)
        s_row = H_XMAX_EXP - np.maximum(e_raw, m_c - H_ROW_FLOOR)
        assert np.all(np.abs(x) * np.exp2(s_row.astype(np.float64))[:, None] < 2.0 ** (H_XMAX_EXP + 1))
