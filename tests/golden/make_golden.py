"""Regenerates tests/golden/*.  Run in the build container (needs /root/reference for iris.xy):

    python tests/golden/make_golden.py

* iris_f32.bin      -- the 150x4 f32 matrix + 150 f32 targets decoded from the reference's embedded
                       dataset src/dataset/iris.xy (format: src/dataset/mod.rs:86-118: u64 LE num_features,
                       u64 LE num_samples, n*d f32 row-major, n f32 targets).  Data fixture, not source.
* kat.json          -- known answers typed from the reference's own tests (bbd_tree.rs:324-364,
                       euclidian.rs:84-91, kmeans.rs:426-443) and the upstream xoshiro256++ vector.
* oracle_fits.json  -- outputs of the CPU oracle (oracle/kmeans_oracle.cpp) for fixed inputs.  The
                       reference cannot be run here (no rustc), so these are REGRESSION fixtures for the
                       oracle + CUDA path, not reference outputs; each records the seeding candidate.
"""
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle_py as O  # noqa: E402

IRIS20 = [[5.1, 3.5, 1.4, 0.2], [4.9, 3.0, 1.4, 0.2], [4.7, 3.2, 1.3, 0.2], [4.6, 3.1, 1.5, 0.2], [5.0, 3.6, 1.4, 0.2],
          [5.4, 3.9, 1.7, 0.4], [4.6, 3.4, 1.4, 0.3], [5.0, 3.4, 1.5, 0.2], [4.4, 2.9, 1.4, 0.2], [4.9, 3.1, 1.5, 0.1],
          [7.0, 3.2, 4.7, 1.4], [6.4, 3.2, 4.5, 1.5], [6.9, 3.1, 4.9, 1.5], [5.5, 2.3, 4.0, 1.3], [6.5, 2.8, 4.6, 1.5],
          [5.7, 2.8, 4.5, 1.3], [6.3, 3.3, 4.7, 1.6], [4.9, 2.4, 3.3, 1.0], [6.6, 2.9, 4.6, 1.3], [5.2, 2.7, 3.9, 1.4]]


def main():
    src = "/root/reference/src/dataset/iris.xy"
    b = open(src, "rb").read()
    nf, ns = struct.unpack("<QQ", b[:16])
    assert (nf, ns) == (4, 150) and len(b) == 16 + ns * nf * 4 + ns * 4
    open(os.path.join(HERE, "iris_f32.bin"), "wb").write(b[16:])

    kat = {
        "iris20": IRIS20,
        "bbdtree_iris": {  # bbd_tree.rs:349-363
            "centroids": [[4.86, 3.22, 1.61, 0.29], [6.23, 2.92, 4.48, 1.42]],
            "cost": 10.68, "cost_tol": 1e-2, "sums_0_0": 48.6, "sums_1_3": 13.8, "sums_tol": 1e-2,
            "membership_17": 1,
        },
        "squared_distance": {"a": [1, 2, 3], "b": [4, 5, 6], "l2": 5.19615242, "tol": 1e-8},  # euclidian.rs:84-91
        "invalid_k_message": "Fit failed: invalid number of clusters: 1",  # kmeans.rs:435-442
        "xoshiro256pp_state_1234": [41943041, 58720359, 3588806011781223, 3591011842654386, 9228616714210784205,
                                    9973669472204895162, 14011001112246962877, 12406186145184390807,
                                    15849039046786891736, 10450023813501588000],
        "splitmix_seed0_state": [0xe220a8397b1dcdaf, 0x6e789e6aa1b965f4, 0x06c45d188009454f, 0xf88bb8a8724c81ec],
        "splitmix_seed0_out": [5987356902031041503, 7051070477665621255, 6633766593972829180, 211316841551650330],
    }
    json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"), indent=1)

    iris = np.frombuffer(b[16:16 + ns * nf * 4], dtype="<f4").reshape(ns, nf)
    fits = {}
    for name, x, k, seed in [("iris_f64_k3_seed42", iris.astype(np.float64), 3, 42),
                             ("iris_f32_k3_seed42", iris.copy(), 3, 42),
                             ("iris20_f64_k2_seedNone", np.array(IRIS20), 2, 0)]:
        for mode in (O.SEED_MODE_PCG, O.SEED_MODE_SPLITMIX):
            r = O.fit(x, k, 100, seed, mode, use_tree=True)
            fits["%s_mode%d" % (name, mode)] = dict(
                seeding="pcg32_fill(rand_core default)" if mode == 0 else "splitmix64", k=k, seed=seed,
                y=r.y.tolist(), size=r.size.tolist(), centroids=r.centroids.tolist(), distortion=r.distortion,
                iters=r.iters, seed_idx=r.seed_idx.tolist())
    json.dump(fits, open(os.path.join(HERE, "oracle_fits.json"), "w"), indent=1)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
