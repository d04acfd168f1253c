#!/usr/bin/env python
"""bench.py -- Lloyd point-iterations/sec on BASELINE.json's headline config, plus every other config of
BASELINE.json as extra keys of the same JSON line.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle)

A "step" is one Lloyd iteration (assignment + per-cluster sums/counts + all-reduce + centroid update + stop-rule
state on the device) over all rows.  Headline workload at every N: config C3 of BASELINE.json, synthetic Gaussian
blobs 10M x 64, k = 256, f64, PER GPU (weak scaling: rank r holds rows [r*10M, (r+1)*10M) of an N*10M-row matrix; the
only data-path collective is one NCCL all-reduce of [k*d sums | k counts | inertia] per step).

  value    : n_global * K / T, T = CUDA-event time of the K timed steps on the library's stream (max over ranks),
             X resident in HBM.  Inputs (5.12 GB/GPU) are far larger than L2, so no flush is needed.
  e2e      : the same metric through the reference-facing call with HOST buffers: ONE call of sckm_kmeans_fit (N = 1)
             or sckm_kmeans_fit_shard (one per rank, N > 1) on PAGEABLE host memory -- what the Rust shim of
             KMeans::fit hands over (a Vec<T>): upload + kmeans++ + initial means + Lloyd loop (reference stop rule,
             max_iter = K) + download of labels / centroids, all inside the timed region.
  roofline : the assignment kernel (dominant), 2*k*d flops per point against the FP64 peak measured by the library's
             own DFMA/DMMA micro-kernels on this GPU (MEASURED_PEAKS.json has no FP64 figure).
  cpu_baseline : the CPU oracle (a port of smartcore's BBD-tree path; rustc is not available) timed on this host,
             single-threaded like the reference, on a row sub-sample with the same k and d.
  configs  : C2 (1M x 16 k=8 f64, one GPU), C4 (100M x 128 k=1024 f64 over 8 GPUs: the 12.5M-row shard per GPU at every
             N, so N = 8 IS config C4 and N = 1 is its paired one-GPU rate), C5 (50M x 32 k=4096 f32 split over the N
             GPUs, kmeans++ on the GPU) -- each with ms_per_step, value, its per-shape roofline and clocks.
  strong   : C3 as a fixed 10M-row problem split over the N GPUs (strong scaling), device-timed steps and, on rank 0,
             the single-process drop-in call (sckm_ctx_create_multi over the N devices + sckm_kmeans_fit from one
             pageable host buffer) end to end.
  parity   : checked in THIS run, at this N: sizes sum to n_global, centroids bit-identical across ranks, labels of
             sampled rows against a float64 numpy brute force (north-star tolerance), and at N > 1 a 1M-row fit
             compared with the same fit on one rank.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_PER_GPU, D, K_CLUSTERS, DATA_SEED, KMEANS_SEED = 10_000_000, 64, 256, 20260101, 42
CPU_SAMPLE_ROWS = 100_000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=N_PER_GPU, help="rows per GPU (default: config C3)")
    # (--dim/--clusters rather than --d/--k: torchrun's own parser treats a bare --d as an abbreviation of its options)
    ap.add_argument("--dim", dest="d", type=int, default=D, help="features (default: config C3; other shapes: tuning only)")
    ap.add_argument("--clusters", dest="k", type=int, default=K_CLUSTERS, help="clusters (default: config C3)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="element type of X (default: config C3 = f64)")
    ap.add_argument("--init", default="kmeanspp", choices=["kmeanspp", "rows"],
                    help="tuning only: 'rows' seeds the timed steps from k evenly spaced rows instead of running kmeans++")
    ap.add_argument("--assign", type=int, default=0, help="tuning only: force an assignment kernel (SCKM_ASSIGN_*)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C2 / C4 / C5 / strong-scaling / parity sections")
    ap.add_argument("--only", default="", help="tuning only: comma list out of c2,c4,c5,strong,parity to run besides the headline")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed regions (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:  # noqa: BLE001
                pass
            self.proc = None

    def window(self, t0, t1):
        """Summarise the samples taken in [t0, t1]; if the region was shorter than the sampling period, the samples
        nearest to it."""
        rows = [r[1:] for r in self.rows if t0 <= r[0] <= t1 + 0.15]
        if not rows and self.rows:
            mid = 0.5 * (t0 + t1)
            rows = [r[1:] for r in sorted(self.rows, key=lambda r: abs(r[0] - mid))[:3]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic():
    """dram__bytes_read + dram__bytes_write of one assignment launch at the default C3 shape, from the newest committed
    `ncu --set full` capture of that kernel under profiles/ (a constant of the build, not measured in this run)."""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for name in ("ncu_r2_assign_dmma_summary.csv", "ncu_r1_assign_dmma_final_summary.csv"):
        path = os.path.join(ROOT, "profiles", name)
        total, seen = 0.0, 0
        try:
            for line in open(path):
                parts = line.strip().split(",")
                if len(parts) == 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and parts[1] in scale:
                    total += float(parts[2]) * scale[parts[1]]
                    seen += 1
        except OSError:
            continue
        if seen == 2:
            return total, "profiles/" + name
    return None, None


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (BBD-tree filter) on this host; needs oracle/ only, never the product library
# ------------------------------------------------------------------------------------------------------------------
def cpu_oracle_run(steps, warmup, rows):
    from oracle import oracle_py as O
    from oracle import blobs_np
    x = blobs_np.blobs(0, rows, D, K_CLUSTERS, DATA_SEED)
    t0 = time.perf_counter()
    tree = O.BBDTree(x)
    t_tree = time.perf_counter() - t0
    t0 = time.perf_counter()
    y, idx, _ = O.kmeanspp(x, K_CLUSTERS, seed=KMEANS_SEED)
    t_kpp = time.perf_counter() - t0
    cent = np.zeros((K_CLUSTERS, D))
    for c in range(K_CLUSTERS):
        cent[c] = x[y == c].mean(0) if np.any(y == c) else x[idx[c]]
    for _ in range(warmup):
        dist, sums, counts, _ = tree.clustering(cent)
        nz = counts > 0
        cent[nz] = sums[nz] / counts[nz, None]
    t0 = time.perf_counter()
    for _ in range(steps):
        dist, sums, counts, _ = tree.clustering(cent)
        nz = counts > 0
        cent[nz] = sums[nz] / counts[nz, None]
    t_steps = time.perf_counter() - t0
    return dict(rows=rows, t_tree=t_tree, t_kpp=t_kpp, t_steps=t_steps, steps=steps,
                value=rows * steps / t_steps, e2e=rows * steps / (t_tree + t_kpp + t_steps))


def reference_arm(args, world, rank):
    if rank != 0:
        return
    r = cpu_oracle_run(args.steps, args.warmup, CPU_SAMPLE_ROWS)
    line = {
        "impl": "reference", "metric": "lloyd_point_iters_per_sec", "value": r["value"], "unit": "point-iters/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["t_steps"] / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3 blobs 10M x 64 k=256 f64 per GPU (BASELINE.json configs[2])", "k": K_CLUSTERS, "d": D,
                   "sample_rows": r["rows"]},
        "cpu_baseline": {"value": r["value"], "unit": "point-iters/s", "cores": 1, "kind": "port",
                         "sample": "%d-row sub-sample of the C3 blobs, same k=%d d=%d; BBD-tree path of smartcore restated in "
                                   "C++ (oracle/, -O3), single thread like the reference; host has %d cores; tree build %.2fs, "
                                   "kmeans++ %.2fs" % (r["rows"], K_CLUSTERS, D, os.cpu_count(), r["t_tree"], r["t_kpp"])},
        "e2e": {"value": r["e2e"], "unit": "point-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def shape_roofline(n_local, k, d, dtype, t_assign, peaks, mp):
    """Per-shape roofline of the assignment launch: intensity = 2kd / (d*s) flop per byte against
    a machine balance of ~6 flop/B."""
    esize = 4 if dtype == "f32" else 8
    flops = 2.0 * n_local * k * d
    hbm_bytes = n_local * (d * esize + 4)
    achieved_tf = flops / t_assign / 1e12
    hbm_peak = mp.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json" if mp else "fallback (B200_PROFILING.md)"
    hbm_view = {"achieved_gbs": hbm_bytes / t_assign / 1e9, "peak_gbs": hbm_peak, "peak_source": hbm_src,
                "copy_gbs_measured_now": peaks["hbm_copy_gbs"]}
    intensity = 2.0 * k / esize
    if intensity < 6.0:
        roof = {"bound": "hbm", "achieved": hbm_view["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                "frac": hbm_view["achieved_gbs"] / hbm_peak,
                "peak_source": hbm_src + " (copy bandwidth; measured now: %.0f GB/s)" % peaks["hbm_copy_gbs"]}
    elif dtype == "f32":
        bf16 = mp.get("bf16_tflops_sustained", 1366.4)
        tf32x3 = bf16 / 2.0 / 3.0
        if d <= 32:
            # K2h (csrc/sckm_tc5h.cu): three FP16 products per real product (FP16 runs at the bf16 rate); the rank-one
            # norm term adds a seventh MMA to every six (not counted: it is overhead, not algorithmic work)
            f16x3 = bf16 / 3.0
            roof = {"bound": "tensor", "achieved": achieved_tf, "peak": f16x3, "unit": "TFLOP/s", "frac": achieved_tf / f16x3,
                    "peak_source": "3xFP16 roof = sustained dense bf16 (= f16 rate) of MEASURED_PEAKS.json / 3 (MMAs per product); "
                                   "the kernel is bound by its per-score epilogue and the power cap, not by the tensor pipe",
                    "frac_of_3xtf32_roof": achieved_tf / tf32x3,
                    "note_3xtf32": "rounds 1-2a ranked with TF32 operands (half the MMA rate): their fractions were quoted "
                                   "against bf16 / 2 / 3 = %.1f TFLOP/s; frac_of_3xtf32_roof keeps that scale for comparison" % tf32x3}
        else:
            roof = {"bound": "tensor", "achieved": achieved_tf, "peak": tf32x3, "unit": "TFLOP/s", "frac": achieved_tf / tf32x3,
                    "peak_source": "3xTF32 roof = sustained dense bf16 of MEASURED_PEAKS.json / 2 (TF32) / 3 (MMAs per product)"}
    else:
        fp64_peak = max(peaks["fp64_dfma_tflops"], peaks["fp64_dmma_tflops"])
        roof = {"bound": "tensor", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved_tf / fp64_peak if fp64_peak else None,
                "peak_source": "FP64 peak measured now by the library's DFMA/DMMA micro-kernels (dfma %.1f, dmma %.1f "
                               "TFLOP/s); MEASURED_PEAKS.json carries no FP64 figure" % (peaks["fp64_dfma_tflops"], peaks["fp64_dmma_tflops"])}
    roof.update({"kernel_ms": 1e3 * t_assign, "flop_per_byte": intensity, "hbm": hbm_view,
                 "algorithmic_bytes_per_launch": hbm_bytes, "algorithmic_flops_per_launch": flops})
    return roof


class Bench:
    def __init__(self, args):
        import torch
        import smartcore_b200 as sc
        from smartcore_b200 import cluster, dist as scd
        self.args, self.torch, self.sc, self.cluster, self.scd = args, torch, sc, cluster, scd
        self.world, self.rank, self.local = (int(os.environ.get(v, d)) for v, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
        torch.cuda.set_device(self.local)
        self.distributed = self.world > 1
        if self.distributed:
            import torch.distributed as tdist
            self.tdist = tdist
            tdist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.cpu_group = tdist.new_group(backend="gloo")   # a barrier that parks a rank on the CPU, its GPU left alone
        # N ranks share this host's cores: give each rank's pinned staging ring its share of them (the library alone sees
        # one process and would start 8 memcpy threads per rank)
        os.environ.setdefault("SCKM_INGEST_THREADS", str(max(2, min(8, (os.cpu_count() or 16) // max(self.world, 1)))))
        self.sampler = ClockSampler(self.local)
        self.sampler.start()                 # nvidia-smi takes seconds to start: begin now, select the timed windows later
        self.ctx = sc.Context(self.local)
        if self.distributed:
            scd.join_comm(self.ctx)
        self.peaks = self.ctx.device_peaks()
        try:
            self.mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            self.mp = {}

    def barrier(self):
        if self.distributed:
            self.tdist.barrier()
        self.torch.cuda.synchronize()

    def cpu_barrier(self):
        if self.distributed:
            self.tdist.barrier(group=self.cpu_group)

    def maxr(self, v):
        return self.scd.max_over_ranks(v) if self.distributed else v

    # -- one configuration: kmeans++ + means, W warm-up steps, K timed steps, device-timed ---------------------------------
    def run_steps(self, n_local, n_global, row0, d, k, dtype, steps, warmup, init="kmeanspp", keep=False):
        ctx = self.ctx
        np_dtype = np.float32 if dtype == "f32" else np.float64
        ds = ctx.generate_blobs(n_local, d, k, DATA_SEED, dtype=np_dtype, row_offset=row0, n_global=n_global)
        first, uniforms = self.cluster.kmeanspp_draws(KMEANS_SEED, n_global, k)
        self.barrier()
        t0 = time.perf_counter()
        if init == "rows":
            step = max(1, n_local // k)
            cent0 = np.vstack([ds.download_rows(i * step, 1) for i in range(k)]).astype(np.float64)
        else:
            ds.kmeanspp(k, first, uniforms)
            cent0, _ = ds.init_centroids(k)
        self.barrier()
        t_init = self.maxr(time.perf_counter() - t0)
        if warmup:
            ds.lloyd_iterate(cent0, warmup)
        self.barrier()
        l0 = ctx.launch_count()
        w0 = time.perf_counter()
        # the K timed steps, timed as a whole on the library's stream (two events, none between the steps: how a fit runs)
        out = ds.lloyd_iterate(cent0, steps, want_inertia=True, per_step_events=False)
        self.barrier()
        w1 = time.perf_counter()
        launches = ctx.launch_count() - l0
        t_dev = self.maxr(float(out["ms"].sum()) * 1e-3)
        # the same steps once more with events around every assignment launch: the dominant kernel's time for the roofline
        probe = ds.lloyd_iterate(cent0, steps)
        t_assign = self.maxr(float(probe["assign_ms"].mean()) * 1e-3)
        res = {"n_local": n_local, "n_global": n_global, "d": d, "k": k, "dtype": dtype, "steps": steps, "warmup": warmup,
               "t_dev": t_dev, "t_assign": t_assign, "wall": self.maxr(w1 - w0), "t_init": t_init, "launches": int(launches),
               "w0": w0, "w1": w1, "out": out, "cent0": cent0, "first": first, "uniforms": uniforms,
               "inertia_non_increasing": bool(np.all(np.diff(out["inertia"]) <= 1e-9 * np.abs(out["inertia"][:-1])))}
        if keep:
            res["ds"] = ds
        else:
            ds.close()
        return res

    def config_entry(self, r, what, scaling, extra=None):
        roof = shape_roofline(r["n_local"], r["k"], r["d"], r["dtype"], r["t_assign"], self.peaks, self.mp)
        e = {"workload": what, "n_gpus": self.world, "scaling": scaling, "n_per_gpu": r["n_local"], "n_global": r["n_global"],
             "d": r["d"], "k": r["k"], "dtype": r["dtype"], "steps": r["steps"], "warmup": r["warmup"],
             "ms_per_step": 1e3 * r["t_dev"] / r["steps"], "value": r["n_global"] * r["steps"] / r["t_dev"], "unit": "point-iters/s",
             "roofline": roof, "kmeanspp_init_s": r["t_init"], "gpu_launches": r["launches"],
             "inertia_non_increasing": r["inertia_non_increasing"], "clocks": self.sampler.window(r["w0"], r["w1"])}
        if self.world > 1:
            e["allreduce"] = self.ctx.allreduce_path()      # 'peer': summed over the ranks inside the finalize kernel; 'nccl'
        if extra:
            e.update(extra)
        return e

    # -- parity, checked in this run ---------------------------------------------------------------------------------------
    def parity(self, r):
        """r: the headline result with its dataset kept.  One more step from the final centroids, then: sizes sum to
        n_global, centroids identical on every rank, labels of sampled rows against a float64 numpy brute force."""
        torch, ds, k, d = self.torch, r["ds"], r["k"], r["d"]
        cent = r["out"]["centroids"]
        inertia, sums, counts = ds.lloyd_step(cent)
        p = {"n_gpus": self.world, "sizes_sum_equals_n_global": bool(int(counts.sum()) == r["n_global"])}
        if self.distributed:
            mine = torch.from_numpy(cent.copy()).cuda()
            allc = [torch.empty_like(mine) for _ in range(self.world)]
            self.tdist.all_gather(allc, mine)
            p["centroids_bit_identical_across_ranks"] = bool(all(torch.equal(allc[0].view(torch.int64), c.view(torch.int64)) for c in allc))
        # sampled rows of this rank's shard, regenerated on the host from the counter-based generator
        from smartcore_b200 import cabi
        rng = np.random.default_rng(7 + self.rank)
        rows = np.sort(rng.choice(r["n_local"], 2048, replace=False))
        row0 = self.rank * r["n_local"]
        np_dtype = np.float32 if r["dtype"] == "f32" else np.float64
        xs = np.vstack([cabi.blobs_host(int(row0 + i), 1, d, k, DATA_SEED, dtype=np_dtype) for i in rows]).astype(np.float64)
        got = ds.labels(width=4)[rows].astype(np.int64)
        dist = ((xs[:, None, :] - cent[None, :, :]) ** 2).sum(-1)                      # direct form, float64
        order = np.argsort(dist, axis=1, kind="stable")
        best, second = dist[np.arange(len(rows)), order[:, 0]], dist[np.arange(len(rows)), order[:, 1]]
        gap = (second - best) / np.maximum(best, 1e-300)
        tol = 1e-12 if r["dtype"] == "f64" else 1e-5
        bad = np.nonzero((got != order[:, 0]) & (gap >= tol))[0]
        nbad = int(len(bad))
        if self.distributed:
            t = torch.tensor([nbad], dtype=torch.int64, device="cuda"); self.tdist.all_reduce(t); nbad = int(t.item())
        p["sampled_rows"] = {"rows_per_rank": int(len(rows)), "label_mismatches_beyond_tolerance": nbad, "tolerance_rel_gap": tol,
                             "checker": "float64 numpy brute force on rows regenerated from the counter-based generator"}
        ds.close()
        # N > 1: a 1M-row fit over all ranks against the same fit on ONE rank
        if self.distributed:
            n, kk, dd = 1_000_000, 256, 64
            lo, hi = self.scd.shard_range(n, self.world, self.rank)
            first, u = self.cluster.kmeanspp_draws(KMEANS_SEED, n, kk)
            dsm = self.ctx.generate_blobs(hi - lo, dd, kk, DATA_SEED + 1, row_offset=lo, n_global=n)
            dsm.kmeanspp(kk, first, u)
            c0, _ = dsm.init_centroids(kk)
            fm = dsm.lloyd_fit(c0, 25)
            lab = torch.zeros(n, dtype=torch.int64, device="cuda")
            lab[lo:hi] = torch.from_numpy(dsm.labels(width=4).astype(np.int64)).cuda()
            self.tdist.all_reduce(lab)
            dsm.close()
            if self.rank == 0:
                solo = self.sc.Context(self.local)
                ds1 = solo.generate_blobs(n, dd, kk, DATA_SEED + 1)
                ds1.kmeanspp(kk, first, u)
                c1, _ = ds1.init_centroids(kk)
                f1 = ds1.lloyd_fit(c1, 25)
                l1 = ds1.labels(width=4).astype(np.int64)
                ds1.close(); solo.close()
                rel = float(np.max(np.abs(fm["centroids"] - f1["centroids"]) / np.maximum(np.abs(f1["centroids"]), 1e-300)))
                p["fit_1M_vs_one_rank"] = {
                    "shape": "1M x 64 k=256 f64, max_iter 25", "iters": [int(fm["iters"]), int(f1["iters"])],
                    "iters_equal": bool(fm["iters"] == f1["iters"]), "sizes_equal": bool(np.array_equal(fm["size"], f1["size"])),
                    "labels_equal": bool(np.array_equal(lab.cpu().numpy(), l1)), "max_rel_diff_centroids": rel,
                    "distortion_rel_diff": float(abs(fm["distortion"] - f1["distortion"]) / f1["distortion"]),
                    "within_north_star_tolerance": bool(rel <= 1e-9 and fm["iters"] == f1["iters"])}
        p["ok"] = bool(p["sizes_sum_equals_n_global"] and p.get("centroids_bit_identical_across_ranks", True)
                       and p["sampled_rows"]["label_mismatches_beyond_tolerance"] == 0
                       and p.get("fit_1M_vs_one_rank", {}).get("within_north_star_tolerance", True))
        return p

    # -- e2e through the one reference-facing call, pageable host memory ------------------------------------------------------
    def e2e(self, n_local, n_global, row0, d, k, dtype, max_iter):
        ctx = self.ctx
        np_dtype = np.float32 if dtype == "f32" else np.float64
        ds = ctx.generate_blobs(n_local, d, k, DATA_SEED, dtype=np_dtype, row_offset=row0, n_global=n_global)
        hx = np.empty((n_local, d), dtype=np_dtype)                       # pageable, like a Vec<T>
        chunk = 1 << 20
        for r in range(0, n_local, chunk):
            m = min(chunk, n_local - r)
            hx[r:r + m] = ds.download_rows(r, m)
        ds.close()
        first, uniforms = self.cluster.kmeanspp_draws(KMEANS_SEED, n_global, k)
        # a long-lived process has fitted before: its pinned staging lanes exist (they are pinned on first use, ~0.1 s
        # once per context) and the device pool holds the blocks; one small fit outside the timed region puts this
        # process in that state
        warm = min(n_local, 2_500_000)
        if self.distributed:
            ctx.kmeans_fit_shard(hx[:warm], row0, n_global, k, 2, first, uniforms)
        else:
            ctx.kmeans_fit(hx[:warm], k, 2, first % warm, uniforms)
        self.barrier()
        e0 = time.perf_counter()
        if self.distributed:
            fit = ctx.kmeans_fit_shard(hx, row0, n_global, k, max_iter, first, uniforms)
        else:
            fit = ctx.kmeans_fit(hx, k, max_iter, first, uniforms)
        self.barrier()
        t_e2e = self.maxr(time.perf_counter() - e0)
        ph = ctx.last_fit_times()
        iters = max(int(fit["iters"]), 1)
        # the same call when the caller's buffer happens to be pinned (direct DMA, no staging pass over host DRAM)
        pinned = None
        try:
            hp = self.torch.empty((n_local, d), dtype=self.torch.float32 if dtype == "f32" else self.torch.float64, pin_memory=True)
            hpn = hp.numpy(); hpn[:] = hx
            self.barrier()
            p0 = time.perf_counter()
            fit_p = ctx.kmeans_fit_shard(hpn, row0, n_global, k, max_iter, first, uniforms) if self.distributed else ctx.kmeans_fit(hpn, k, max_iter, first, uniforms)
            self.barrier()
            t_p = self.maxr(time.perf_counter() - p0)
            php = ctx.last_fit_times()
            pinned = {"value": n_global * max(int(fit_p["iters"]), 1) / t_p, "total_s": t_p, "upload_s": self.maxr(php["upload_s"]),
                      "same_result": bool(np.array_equal(fit_p["centroids"], fit["centroids"]) and fit_p["iters"] == fit["iters"])}
            del hp, hpn
        except Exception as exc:  # noqa: BLE001
            pinned = {"error": str(exc)[:200]}
        return {"value": n_global * iters / t_e2e, "unit": "point-iters/s",
                "h2d_bytes_per_step": int(hx.nbytes // iters),
                "d2h_bytes_per_step": int((fit["labels"].nbytes + fit["centroids"].nbytes + fit["size"].nbytes) // iters),
                "detail": {"what": "ONE call of %s on PAGEABLE host memory per rank (after one small warm-up fit that pins the staging lanes): upload + kmeans++ + initial means + Lloyd "
                                   "loop (device-side stop rule, max_iter = steps) + labels (usize) / centroids download; bytes "
                                   "are per fit divided by the iterations executed" % ("sckm_kmeans_fit_shard" if self.distributed else "sckm_kmeans_fit"),
                           "iters": iters, "total_s": t_e2e, "upload_s": self.maxr(ph["upload_s"]), "kmeanspp_init_s": self.maxr(ph["kmeanspp_init_s"]),
                           "lloyd_s": self.maxr(ph["lloyd_s"]), "download_s": self.maxr(ph["download_s"]), "distortion": fit["distortion"],
                           "upload_gbs_per_gpu": hx.nbytes / 1e9 / max(self.maxr(ph["upload_s"]), 1e-9),
                           "host_cores": os.cpu_count(), "staging_threads_per_rank": int(os.environ["SCKM_INGEST_THREADS"]),
                           "same_call_from_pinned_memory": pinned}}

    # -- strong scaling of C3 + the single-process drop-in call over the N devices -----------------------------------------------
    def strong(self, headline):
        a = self.args
        n = N_PER_GPU
        if self.world == 1:
            e = {"workload": "C3 blobs 10M x 64 k=256 f64, 10M rows GLOBAL split over the N GPUs (strong scaling)", "n_gpus": 1,
                 "n_global": n, "ms_per_step": headline["ms_per_step"], "value": headline["value"], "unit": "point-iters/s",
                 "note": "N = 1: identical to the headline run"}
            return e
        lo, hi = self.scd.shard_range(n, self.world, self.rank)
        r = self.run_steps(hi - lo, n, lo, D, K_CLUSTERS, "f64", a.steps, a.warmup)
        e = self.config_entry(r, "C3 blobs 10M x 64 k=256 f64, 10M rows GLOBAL split over the N GPUs (strong scaling)", "strong")
        # the drop-in form: ONE process, ONE context over the N devices, one pageable host buffer -> sckm_kmeans_fit.
        # Rank 0 drives all N devices; the other ranks wait on a CPU (gloo) barrier with their datasets closed -- a NCCL
        # barrier would spin a kernel on their GPUs and time-slice against rank 0's work there.
        self.barrier()
        if self.rank == 0:
            try:
                mc = self.sc.Context(devices=list(range(self.world)))
                gen = self.sc.Context(self.local)
                dsg = gen.generate_blobs(n, D, K_CLUSTERS, DATA_SEED)
                hx = np.empty((n, D))
                for q in range(0, n, 1 << 20):
                    m = min(1 << 20, n - q)
                    hx[q:q + m] = dsg.download_rows(q, m)
                dsg.close(); gen.close()
                first, u = self.cluster.kmeanspp_draws(KMEANS_SEED, n, K_CLUSTERS)
                mc.kmeans_fit(hx, K_CLUSTERS, 2, first, u)      # a first fit pins every device's staging lanes and warms NCCL
                t0 = time.perf_counter()
                fit = mc.kmeans_fit(hx, K_CLUSTERS, a.steps, first, u)
                t = time.perf_counter() - t0
                ph = mc.last_fit_times()
                it = max(int(fit["iters"]), 1)
                e["single_process_e2e"] = {
                    "what": "sckm_ctx_create_multi over the N devices, ONE sckm_kmeans_fit call from one pageable host buffer "
                            "(what KMeans::fit gets through the Rust shim), rank 0 only; measured INSIDE this torchrun job, whose other "
                            "ranks stay resident on the devices (bench/multi_probe.py is the same call in a process of its own)",
                    "value": n * it / t, "unit": "point-iters/s", "iters": it, "total_s": t, "devices": ph["devices"],
                    "upload_s": ph["upload_s"], "kmeanspp_init_s": ph["kmeanspp_init_s"], "lloyd_s": ph["lloyd_s"], "download_s": ph["download_s"],
                    "h2d_bytes_per_step": int(hx.nbytes // it), "d2h_bytes_per_step": int((fit["labels"].nbytes + fit["centroids"].nbytes) // it),
                    "sizes_sum_equals_n": bool(int(fit["size"].sum()) == n)}
                mc.close()
            except Exception as exc:  # noqa: BLE001
                e["single_process_e2e"] = {"error": str(exc)[:300]}
        self.cpu_barrier()
        return e

    def close(self):
        self.sampler.stop()
        self.ctx.close()
        if self.distributed:
            self.tdist.destroy_process_group()


def main():
    args = parse()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, world, rank)

    # NCCL prints its version banner on stdout at any debug level >= VERSION; rank 0 must print ONE JSON line, so send
    # NCCL's own log to a file instead (override with NCCL_DEBUG_FILE)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/sckm_nccl_%h_%p.log")
    # ... and whatever else a library writes to fd 1 (NCCL 2.28 still prints its banner there from some code paths) goes to
    # stderr for the whole run; stdout comes back for the one JSON line at the end
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    B = Bench(args)
    n_local, k, d = args.rows, args.k, args.d
    n_global, row0 = n_local * world, rank * n_local
    default_shape = (n_local, k, d, args.dtype) == (N_PER_GPU, K_CLUSTERS, D, "f64")
    if args.assign:
        B.ctx.set_assign_kernel(args.assign)
    only = set(s for s in args.only.split(",") if s)
    want = (lambda name: default_shape and not args.no_configs and (not only or name in only))

    # ---- headline: C3, weak ----
    head = B.run_steps(n_local, n_global, row0, d, k, args.dtype, args.steps, args.warmup, init=args.init, keep=want("parity"))
    head_path = B.ctx.allreduce_path()
    value = n_global * args.steps / head["t_dev"]
    roof = shape_roofline(n_local, k, d, args.dtype, head["t_assign"], B.peaks, B.mp)
    traffic, traffic_src = ncu_traffic() if default_shape else (None, None)
    roof.update({"traffic": traffic,
                 "traffic_note": "DRAM read+write bytes of one assignment launch: a constant of the build taken from the committed ncu "
                                 "--set full capture %s, NOT measured in this run; algorithmic bytes per launch = n*(d*s+4) = %.3e"
                                 % (traffic_src, roof["algorithmic_bytes_per_launch"]),
                 "kernel": "assignment kernel (dominant), CUDA events on the library stream, mean of %d launches" % args.steps})
    clocks = B.sampler.window(head["w0"], head["w1"])
    parity = B.parity(head) if want("parity") else None

    # ---- e2e ----
    e2e = None if args.no_e2e else B.e2e(n_local, n_global, row0, d, k, args.dtype, args.steps)

    # ---- the other configs of BASELINE.json ----
    configs, strong = {}, None
    sub_steps = max(3, min(args.steps, 10))
    if want("c2"):
        if world == 1:
            r = B.run_steps(1_000_000, 1_000_000, 0, 16, 8, "f64", max(args.steps, 20), max(args.warmup, 3))
            configs["C2"] = B.config_entry(r, "C2 blobs 1M x 16 k=8 f64 on 1 B200 (BASELINE.json configs[1]); X = 128 MB, just above the "
                                              "126 MB L2: no flush between steps", "none (1 GPU)")
        else:
            configs["C2"] = {"workload": "C2 blobs 1M x 16 k=8 f64", "skipped": "defined on 1 GPU; see the N = 1 line"}
    if want("c4"):
        r = B.run_steps(12_500_000, 12_500_000 * world, rank * 12_500_000, 128, 1024, "f64", sub_steps, args.warmup)
        configs["C4"] = B.config_entry(r, "C4 blobs 100M x 128 k=1024 f64 over 8 GPUs (BASELINE.json configs[3]): its 12.5M-row shard per GPU at "
                                          "every N -- N = 8 is config C4 itself, N = 1 its paired one-GPU rate", "weak")
    if want("c5"):
        lo, hi = B.scd.shard_range(50_000_000, world, rank)
        r = B.run_steps(hi - lo, 50_000_000, lo, 32, 4096, "f32", sub_steps, args.warmup)
        configs["C5"] = B.config_entry(r, "C5 blobs 50M x 32 k=4096 f32, kmeans++ on the GPU (BASELINE.json configs[4]): 50M rows GLOBAL split over "
                                          "the N GPUs", "strong")
    if want("strong"):
        strong = B.strong({"ms_per_step": 1e3 * head["t_dev"] / args.steps, "value": value})

    cpu = None
    if rank == 0 and not args.no_cpu:
        r = cpu_oracle_run(5, 1, CPU_SAMPLE_ROWS)
        cpu = {"value": r["value"], "unit": "point-iters/s", "cores": 1, "kind": "port",
               "sample": "%d-row sub-sample, same k and d, 5 timed BBD-tree clustering steps (oracle/ C++ port of "
                         "smartcore's path, -O3; single thread like the reference; host has %d cores); tree build %.2fs"
                         % (r["rows"], os.cpu_count(), r["t_tree"])}

    if rank == 0:
        line = {
            "metric": "lloyd_point_iters_per_sec", "value": value, "unit": "point-iters/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * head["t_dev"] / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "C3 blobs 10M x 64 k=256 f64 per GPU (BASELINE.json configs[2])" if default_shape
                                   else "tuning shape %d x %d k=%d %s per GPU" % (n_local, d, k, args.dtype), "n_per_gpu": n_local,
                       "n_global": n_global, "d": d, "k": k, "l2": "inputs (5.12 GB/GPU) larger than L2; no flush",
                       "parallelism": ("one GPU" if world == 1 else "rows sharded x%d; per step ONE sum of k*d+k+1 f64 over the ranks, %s"
                                       % (world, "inside the finalize kernel over NVLink peer memory (csrc/sckm_peer.cu)"
                                          if head_path == "peer" else "through ncclAllReduce")),
                       "kmeanspp_init_s": head["t_init"], "wall_s_timed_region": head["wall"],
                       "stop_rule": "evaluated on the device every step (finalize kernel); the timed steps include it"},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": head["launches"], "clocks": clocks,
            "configs": configs or None, "strong": {"C3": strong} if strong else None, "parity": parity,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    B.close()


if __name__ == "__main__":
    main()
