"""CPU: host-side mirror of the reference API (validation, RNG draws, grid iterator, layouts), the
C-ABI library's exported symbols, and loud failure without a GPU.  No device compute here."""
import ctypes
import os
import re

import numpy as np
import pytest

import smartcore_b200 as sc
from smartcore_b200 import cabi, cluster, dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "smartcore_kmeans_cuda.h")).read()
    declared = sorted(set(re.findall(r"\b(sckm_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found"
    lib = ctypes.CDLL(cabi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libsmartcore_kmeans_cuda.so does not export %s" % name
    assert sorted(cabi.SYMBOLS) == declared
    assert lib.sckm_abi_version() == int(re.search(r"#define SCKM_ABI_VERSION (\d+)", header).group(1)) == 3


def test_no_cpu_fallback_when_device_missing():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sc.SckmError) as e:
        sc.Context(0)
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(sc.Failed) as e2:
        sc.KMeans.fit(sc.DenseMatrix.from_2d_array([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]]), sc.KMeansParameters())
    assert str(e2.value).startswith("Fit failed: CUDA backend unavailable")


def test_rust_ffi_declares_every_header_symbol():
    """rust/src/gpu/ffi.rs (uncompiled here: no rustc) must declare exactly the functions include/*.h exports, and the
    wrapper must not need `'static` / TypeId on `Number` (it is not `'static`: src/numbers/basenum.rs:8-23)."""
    header = open(os.path.join(ROOT, "include", "smartcore_kmeans_cuda.h")).read()
    declared = set(re.findall(r"\b(sckm_[a-z0-9_]+)\s*\(", header))
    ffi = open(os.path.join(ROOT, "rust", "src", "gpu", "ffi.rs")).read()
    bound = set(re.findall(r"pub fn (sckm_[a-z0-9_]+)\s*\(", ffi))
    assert declared == bound == set(cabi.SYMBOLS), (declared ^ bound, declared ^ set(cabi.SYMBOLS))
    assert "SCKM_ABI_VERSION: c_int = %d" % cabi.lib.sckm_abi_version() in ffi
    mod = open(os.path.join(ROOT, "rust", "src", "gpu", "mod.rs")).read()
    code = "\n".join(l for l in mod.splitlines() if not l.lstrip().startswith("//"))      # comments may mention them
    assert "TypeId" not in code and "'static" not in code
    assert os.path.exists(os.path.join(ROOT, "rust", "build.rs"))


def test_product_package_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "smartcore_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower(), "%s mentions the oracle" % f


def test_invalid_k(kat):  # kmeans.rs:426-443 (i32 matrix, validation before any device work)
    x = sc.DenseMatrix.from_2d_array([[1, 2, 3], [4, 5, 6]], dtype=np.int32)
    with pytest.raises(sc.Failed):
        sc.KMeans.fit(x, sc.KMeansParameters.default().with_k(0))
    with pytest.raises(sc.Failed) as e:
        sc.KMeans.fit(x, sc.KMeansParameters.default().with_k(1))
    assert str(e.value) == kat["invalid_k_message"]
    with pytest.raises(sc.Failed) as e:
        sc.KMeans.fit(x, sc.KMeansParameters.default().with_max_iter(0))
    assert str(e.value) == "Fit failed: invalid maximum number of iterations: 0"


def test_search_parameters():  # kmeans.rs:446-466
    it = iter(sc.KMeansSearchParameters(k=[2, 4], max_iter=[10, 100]))
    got = [(p.k, p.max_iter) for p in it]
    assert got == [(2, 10), (4, 10), (2, 100), (4, 100)]
    got = [(p.k, p.max_iter, p.seed) for p in sc.KMeansSearchParameters(k=[2, 3], max_iter=[5], seed=[None, 7])]
    assert got == [(2, 5, None), (3, 5, None), (2, 5, 7), (3, 5, 7)]
    d = sc.KMeansParameters.default()
    assert (d.k, d.max_iter, d.seed) == (2, 100, None)


def test_host_rng_draws_match_oracle_rng(O):
    for seed, n, k in [(None, 150, 3), (42, 150, 3), (7, 10**7, 256), (2**63 + 5, 12345, 17)]:
        first, u = cluster.kmeanspp_draws(seed, n, k)
        r = O.Rng(0 if seed is None else seed)
        assert first == r.gen_range(n)
        assert u.tolist() == [r.gen_f64() for _ in range(k - 1)]


def test_dense_matrix_layouts():
    a = np.arange(12, dtype=np.float64).reshape(4, 3)
    cm = sc.DenseMatrix.from_2d_array(a)          # column-major (matrix.rs:230-236)
    rm = sc.DenseMatrix.new(4, 3, a.reshape(-1), False)
    assert cm.column_major and not rm.column_major
    assert cm.values.tolist() == a.T.reshape(-1).tolist()
    for r in range(4):
        for c in range(3):
            assert cm.get((r, c)) == a[r, c] == rm.get((r, c))
    with pytest.raises(sc.Failed):
        sc.DenseMatrix.new(4, 3, np.zeros(11), True)


def test_blob_generator_host_twin():
    a = cabi.blobs_host(0, 64, 8, 4, 20260101)
    b = cabi.blobs_host(16, 8, 8, 4, 20260101)
    assert np.array_equal(a[16:24], b)            # any row regenerates independently of the shard
    assert not np.array_equal(a, cabi.blobs_host(0, 64, 8, 4, 20260102))
    f = cabi.blobs_host(0, 64, 8, 4, 20260101, dtype=np.float32)
    assert np.array_equal(f, a.astype(np.float32))
    big = cabi.blobs_host(0, 4000, 4, 4, 1)
    for c in range(4):                             # centre + N(0,1); centres in [-10, 10)
        pts = big[c::4]
        assert np.all(np.abs(pts.mean(0)) < 10.2) and np.all(np.abs(pts.std(0) - 1.0) < 0.1)


def test_shard_range_partitions_rows():
    for n in (0, 1, 150, 1024, 1025, 10**6 + 7, 10**7):
        for world in (1, 2, 3, 4, 8):
            spans = [dist.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            for lo, hi in spans[:-1]:
                assert (hi - lo) % dist.SHARD_ALIGN == 0 or hi == n


# ---- persistence: the serde images of the model (kmeans.rs:70-83; round-trip test kmeans.rs:538-544) --------------
SERDE_JSON_GOLDEN = ('{"k":2,"_y":[0,1,1],"size":[1,2],"_distortion":0.5,"centroids":[[1.0,2.5],[3.0,-4.25]],'
                     '"_phantom_tx":null,"_phantom_ty":null,"_phantom_x":null,"_phantom_y":null}')


def test_model_serde_json_layout_and_round_trip():
    """What serde_json writes for the derive on KMeans: declaration order, floats as floats, PhantomData as null."""
    m = sc.KMeans.from_json(SERDE_JSON_GOLDEN)
    assert m.k == 2 and m._y.tolist() == [0, 1, 1] and m.size.tolist() == [1, 2] and m._distortion == 0.5
    assert m.centroids.tolist() == [[1.0, 2.5], [3.0, -4.25]]
    assert m.to_json() == SERDE_JSON_GOLDEN                      # byte-identical image
    assert sc.KMeans.from_json(m.to_json()) == m                # PartialEq (kmeans.rs:85-107)
    # field order is free on input, unknown fields are skipped, shortest round-trip floats
    shuffled = ('{"centroids":[[0.1,1e-7],[1e300,2]],"extra":{"a":[1,2,{"b":"x"}]},"size":[5,6],"k":2,"_y":[],'
                '"_distortion":1.7976931348623157e308}')
    m2 = sc.KMeans.from_json(shuffled)
    assert m2.centroids.tolist() == [[0.1, 1e-7], [1e300, 2.0]] and m2._distortion == 1.7976931348623157e308
    assert sc.KMeans.from_json(m2.to_json()) == m2 and m2 != m
    import json
    assert json.loads(m2.to_json())["centroids"] == [[0.1, 1e-7], [1e300, 2.0]]


def test_model_serde_nan_centroids_and_errors():
    # an empty initial cluster leaves 0/0 centroids (kmeans.rs:284-290); serde_json writes non-finite f64 as null
    m = sc.KMeans.from_json('{"k":2,"_y":[0],"size":[1,0],"_distortion":0.0,"centroids":[[1.5],[null]]}')
    assert np.isnan(m.centroids[1][0]) and '[null]' in m.to_json()
    for bad in ('{"k":2', '{"k":"two","centroids":[]}', '{"_y":[1]}', '[1,2]'):
        with pytest.raises(sc.Failed):
            sc.KMeans.from_json(bad)


def test_model_images_with_inconsistent_shapes_are_refused():
    """A hand-edited image whose k / size / centroids / labels disagree must not reach the accessors (they size their
    buffers from k and centroids[0]); the Rust reference panics safely on such a model, here the loaders refuse it."""
    import struct
    for bad in ('{"k":3,"_y":[0],"size":[1,0],"_distortion":0.0,"centroids":[[1.5],[2.5]]}',        # k != rows
                '{"k":2,"_y":[0],"size":[1,0,7,7,7],"_distortion":0.0,"centroids":[[1.5],[2.5]]}',   # size too long
                '{"k":2,"_y":[0],"size":[1],"_distortion":0.0,"centroids":[[1.5],[2.5]]}',           # size too short
                '{"k":2,"_y":[0],"size":[1,0],"_distortion":0.0,"centroids":[[1.5],[2.5,3.5]]}',     # ragged rows
                '{"k":2,"_y":[0,2],"size":[1,1],"_distortion":0.0,"centroids":[[1.5],[2.5]]}'):      # label >= k
        with pytest.raises(sc.Failed):
            sc.KMeans.from_json(bad)
    ragged = (struct.pack("<Q", 2) + struct.pack("<Q", 0) + struct.pack("<Q2Q", 2, 1, 2) + struct.pack("<d", 0.5)
              + struct.pack("<Q", 2) + struct.pack("<Q2d", 2, 1.0, 2.5) + struct.pack("<Q1d", 1, 3.0))
    short = (struct.pack("<Q", 4) + struct.pack("<Q", 0) + struct.pack("<Q2Q", 2, 1, 2) + struct.pack("<d", 0.5)
             + struct.pack("<Q", 2) + struct.pack("<Q2d", 2, 1.0, 2.5) + struct.pack("<Q2d", 2, 3.0, 4.0))
    for bad in (ragged, short):
        with pytest.raises(sc.Failed):
            sc.KMeans.from_bincode(bad)


def test_model_bincode_layout():
    """bincode 1.3 default options: LE u64 for usize and lengths, raw f64, nothing for PhantomData."""
    import struct
    m = sc.KMeans.from_json(SERDE_JSON_GOLDEN, dtype=np.float32)
    want = (struct.pack("<Q", 2) + struct.pack("<Q3Q", 3, 0, 1, 1) + struct.pack("<Q2Q", 2, 1, 2) + struct.pack("<d", 0.5)
            + struct.pack("<Q", 2) + struct.pack("<Q2d", 2, 1.0, 2.5) + struct.pack("<Q2d", 2, 3.0, -4.25))
    assert m.to_bincode() == want
    back = sc.KMeans.from_bincode(want, dtype=np.float32)
    assert back == m and back._y.tolist() == [0, 1, 1]
    for bad in (want[:-1], want + b"\\0", struct.pack("<Q", 2) + struct.pack("<Q", 1 << 60)):
        with pytest.raises(sc.Failed):
            sc.KMeans.from_bincode(bad)


def test_every_environment_switch_is_documented():
    """INTEGRATION.md lists the SCKM_* environment switches the library reads (tests and experiments only): a switch added
    to the sources without a line there fails here."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    found = set()
    for path in glob.glob(os.path.join(root, "smartcore_b200", "csrc", "*.cu*")) + glob.glob(os.path.join(root, "smartcore_b200", "host", "*")) \
            + glob.glob(os.path.join(root, "smartcore_b200", "*.py")):
        text = open(path, errors="ignore").read()
        found.update(re.findall(r'getenv\("(SCKM_[A-Z0-9_]+)"\)', text))
        found.update(re.findall(r'environ[^\n]*?"(SCKM_[A-Z0-9_]+)"', text))
    assert found, "no switches found: the scan is broken"
    missing = sorted(v for v in found if v not in doc)
    assert not missing, "undocumented environment switches: %s" % missing
