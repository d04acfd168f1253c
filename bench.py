#!/usr/bin/env python
"""bench.py -- Lloyd point-iterations/sec on BASELINE.json's headline config.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle)

A "step" is one Lloyd iteration (assignment + per-cluster sums/counts + all-reduce + centroid update +
inertia) over all rows.  Workload at every N: config C3 of BASELINE.json, synthetic Gaussian blobs
10M x 64, k = 256, f64, PER GPU (weak scaling: rank r holds rows [r*10M, (r+1)*10M) of an N*10M-row
matrix; the only data-path collective is one NCCL all-reduce of [k*d sums | k counts | inertia] per step).

  value : n_global * K / T, T = CUDA-event time of the K timed steps on the library's stream (max over
          ranks), X resident in HBM.  Inputs (5.12 GB/GPU) are far larger than L2, so no flush is needed.
  e2e   : the same metric through the reference-facing call sequence of KMeans::fit with HOST buffers:
          upload of X from pinned host memory + kmeans++ + initial means + Lloyd loop (reference stop
          rule, max_iter = K) + download of labels/centroids, all inside the timed region.
  roofline : the assignment kernel (dominant), 2*k*d flops per point against the FP64 peak measured by
          the library's own DFMA/DMMA micro-kernels on this GPU (MEASURED_PEAKS.json has no FP64 figure).
  cpu_baseline : the CPU oracle (a port of smartcore's BBD-tree path; rustc is not available) timed on
          this host, single-threaded like the reference, on a row sub-sample with the same k and d.
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
    ap.add_argument("--alt", action="store_true",
                    help="also time the experimental tcgen05-ranked path on f64 data (3xTF32 ranking of an f32 shadow, exact f64 result)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
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

    def mark(self):
        return time.perf_counter()

    def stop(self, t0=None, t1=None):
        """Summarise the samples taken in [t0, t1] (the timed region); if the region was shorter than the sampling
        period, fall back to the samples nearest to it."""
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:  # noqa: BLE001
                pass
        rows = [r[1:] for r in self.rows if t0 is None or (t0 <= r[0] <= t1 + 0.15)]
        if not rows and self.rows:
            mid = 0.5 * ((t0 or 0) + (t1 or 0))
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


def ncu_traffic_bytes():
    """dram__bytes_read + dram__bytes_write of one assignment launch at the default C3 shape, from the committed
    `ncu --set full` capture (profiles/ncu_r1_assign_dmma_final_summary.csv); None when the file is absent."""
    path = os.path.join(ROOT, "profiles", "ncu_r1_assign_dmma_final_summary.csv")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    total, seen = 0.0, 0
    try:
        for line in open(path):
            parts = line.strip().split(",")
            if len(parts) == 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and parts[1] in scale:
                total += float(parts[2]) * scale[parts[1]]
                seen += 1
    except OSError:
        return None
    return total if seen == 2 else None


def cpu_oracle_run(steps, warmup, rows):
    """The reference's algorithm (BBD-tree filter) on this host: single thread, row sub-sample."""
    from oracle import oracle_py as O
    from smartcore_b200 import cabi
    x = cabi.blobs_host(0, rows, D, K_CLUSTERS, DATA_SEED)
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
    cores = 1
    line = {
        "impl": "reference", "metric": "lloyd_point_iters_per_sec", "value": r["value"], "unit": "point-iters/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["t_steps"] / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3 blobs 10M x 64 k=256 f64 per GPU (BASELINE.json configs[2])", "k": K_CLUSTERS, "d": D,
                   "sample_rows": r["rows"]},
        "cpu_baseline": {"value": r["value"], "unit": "point-iters/s", "cores": cores, "kind": "port",
                         "sample": "%d-row sub-sample of the C3 blobs, same k=%d d=%d; BBD-tree path of smartcore restated in "
                                   "C++ (oracle/), single thread like the reference; host has %d cores; tree build %.2fs, "
                                   "kmeans++ %.2fs" % (r["rows"], K_CLUSTERS, D, os.cpu_count(), r["t_tree"], r["t_kpp"])},
        "e2e": {"value": r["e2e"], "unit": "point-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    world, rank, local_rank = (int(os.environ.get(v, d)) for v, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    if args.impl == "reference":
        return reference_arm(args, world, rank)

    # NCCL prints its version banner on stdout at any debug level >= VERSION; rank 0 must print ONE JSON line, so send
    # NCCL's own log to a file instead (override with NCCL_DEBUG_FILE)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/sckm_nccl_%h_%p.log")
    import torch
    import smartcore_b200 as sc
    from smartcore_b200 import cluster, dist as scd

    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        import torch.distributed as tdist
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if distributed:
            tdist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()                      # nvidia-smi takes seconds to start: begin now, select the timed window later
    ctx = sc.Context(local_rank)
    if distributed:
        scd.join_comm(ctx)
    n_local, k, d = args.rows, args.k, args.d
    n_global = n_local * world
    row0 = rank * n_local

    peaks = ctx.device_peaks()
    np_dtype = np.float32 if args.dtype == "f32" else np.float64
    esize = 4 if args.dtype == "f32" else 8
    if args.assign:
        ctx.set_assign_kernel(args.assign)
    ds = ctx.generate_blobs(n_local, d, k, DATA_SEED, dtype=np_dtype, row_offset=row0, n_global=n_global)
    first, uniforms = cluster.kmeanspp_draws(KMEANS_SEED, n_global, k)
    t0 = time.perf_counter()
    if args.init == "rows":
        step = max(1, n_local // k)
        cent0 = np.vstack([ds.download_rows(i * step, 1) for i in range(k)]).astype(np.float64)
    else:
        ds.kmeanspp(k, first, uniforms)
        cent0, _ = ds.init_centroids(k)
    t_init = time.perf_counter() - t0

    if args.warmup:
        ds.lloyd_iterate(cent0, args.warmup)
    barrier()
    l0 = ctx.launch_count()
    w0 = time.perf_counter()
    out = ds.lloyd_iterate(cent0, args.steps)
    barrier()
    w1 = time.perf_counter()
    wall = w1 - w0
    launches = ctx.launch_count() - l0
    t_dev = float(out["ms"].sum()) * 1e-3
    t_assign = float(out["assign_ms"].mean()) * 1e-3
    # ---- optional second figure: same steps with the tcgen05 kernel ranking an f32 shadow of X in 3xTF32 while every
    # decision, distance and sum stays exact f64 (opt-in kernel, SCKM_ASSIGN_TC5); reported beside the FP64 headline ----
    alt = None
    if args.alt and args.dtype == "f64" and not args.assign and d <= 64 and d % 4 == 0 and k >= 16:
        from smartcore_b200 import cabi as _cabi
        ctx.set_assign_kernel(_cabi.ASSIGN_TC5)
        ds.lloyd_iterate(cent0, max(args.warmup, 1))
        barrier()
        out2 = ds.lloyd_iterate(cent0, args.steps)
        barrier()
        ctx.set_assign_kernel(_cabi.ASSIGN_AUTO)
        t2 = float(out2["ms"].sum()) * 1e-3
        if distributed:
            t2 = scd.max_over_ranks(t2)
        denom = np.maximum(np.abs(out["centroids"]), 1e-300)
        alt = {"what": "same K steps with the tcgen05 kernel: 3xTF32 ranking of an f32 shadow of X on the tensor cores, "
                       "near-ties re-decided exactly, distances and sums in f64 (SCKM_ASSIGN_TC5, opt-in for f64 data)",
               "ms_per_step": 1e3 * t2 / args.steps, "value": n_global * args.steps / t2, "unit": "point-iters/s",
               "max_rel_diff_final_centroids_vs_fp64_path": float(np.max(np.abs(out2["centroids"] - out["centroids"]) / denom)),
               "sizes_equal": bool(np.array_equal(out2["size"], out["size"]))}
    if distributed:
        t_dev = scd.max_over_ranks(t_dev)
        t_assign = scd.max_over_ranks(t_assign)
        wall = scd.max_over_ranks(wall)
    value = n_global * args.steps / t_dev

    # ---- e2e: KMeans::fit call sequence from pinned host buffers (per rank: its shard) ----
    e2e = None
    if not args.no_e2e:
        host = torch.empty((n_local, d), dtype=torch.float32 if args.dtype == "f32" else torch.float64, pin_memory=True)
        hx = host.numpy()
        chunk = 1 << 20
        for r in range(0, n_local, chunk):
            m = min(chunk, n_local - r)
            hx[r:r + m] = ds.download_rows(r, m)
        ds.close()
        barrier()
        e0 = time.perf_counter()
        ds2 = ctx.upload(hx, column_major=False, row_offset=row0, n_global=n_global)
        t_up = time.perf_counter() - e0
        ds2.kmeanspp(k, first, uniforms)
        c0, _ = ds2.init_centroids(k)
        t_seed = time.perf_counter() - e0 - t_up
        fit = ds2.lloyd_fit(c0, args.steps)
        t_lloyd = time.perf_counter() - e0 - t_up - t_seed
        labels = ds2.labels(width=8)
        barrier()
        t_e2e = time.perf_counter() - e0
        t_down = t_e2e - t_lloyd - t_up - t_seed
        if distributed:
            t_e2e = scd.max_over_ranks(t_e2e)
        e2e = {"value": n_global * fit["iters"] / t_e2e, "unit": "point-iters/s",
               "h2d_bytes_per_step": int(hx.nbytes // max(fit["iters"], 1)),
               "d2h_bytes_per_step": int((labels.nbytes + fit["centroids"].nbytes) // max(fit["iters"], 1)),
               "detail": {"what": "upload(pinned host) + kmeans++ + init means + Lloyd loop (stop rule, max_iter=steps) + "
                                  "labels/centroids download; bytes are per fit divided by iterations executed",
                          "iters": int(fit["iters"]), "total_s": t_e2e, "upload_s": t_up, "kmeanspp_init_s": t_seed, "lloyd_s": t_lloyd, "download_s": t_down,
                          "distortion": fit["distortion"]}}
        ds2.close()
    else:
        ds.close()

    clocks = sampler.stop(w0, w1)
    cpu = None
    if rank == 0 and not args.no_cpu:
        r = cpu_oracle_run(5, 1, CPU_SAMPLE_ROWS)
        cpu = {"value": r["value"], "unit": "point-iters/s", "cores": 1, "kind": "port",
               "sample": "%d-row sub-sample, same k and d, 5 timed BBD-tree clustering steps (oracle/ C++ port of "
                         "smartcore's path; single thread like the reference; host has %d cores); tree build %.2fs"
                         % (r["rows"], os.cpu_count(), r["t_tree"])}

    if rank == 0:
        flops_per_launch = 2.0 * n_local * k * d
        fp64_peak = max(peaks["fp64_dfma_tflops"], peaks["fp64_dmma_tflops"])
        achieved = flops_per_launch / t_assign / 1e12
        hbm_bytes = n_local * (d * esize + 4)
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            mp = {}
        # per-shape roofline: intensity = 2kd / (d*s) flop per byte against a balance of ~6
        hbm_peak = mp.get("hbm_gbs", 6650.0)
        hbm_view = {"achieved_gbs": hbm_bytes / t_assign / 1e9, "peak_gbs": hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if mp else "fallback (B200_PROFILING.md)",
                    "copy_gbs_measured_now": peaks["hbm_copy_gbs"]}
        intensity = 2.0 * k / esize
        if intensity < 6.0:
            roof = {"bound": "hbm", "achieved": hbm_view["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": hbm_view["achieved_gbs"] / hbm_peak,
                    "peak_source": hbm_view["peak_source"] + " (copy bandwidth; measured now: %.0f GB/s)" % peaks["hbm_copy_gbs"]}
        elif args.dtype == "f32":
            # tcgen05 kind::tf32 runs at half the bf16 rate and the 3xTF32 split issues three MMAs per product
            tf32x3 = mp.get("bf16_tflops_sustained", 1366.4) / 2.0 / 3.0
            roof = {"bound": "tensor", "achieved": achieved, "peak": tf32x3, "unit": "TFLOP/s", "frac": achieved / tf32x3,
                    "peak_source": "3xTF32 roof = sustained dense bf16 of MEASURED_PEAKS.json / 2 (TF32) / 3 (MMAs per product)"}
        else:
            roof = {"bound": "tensor", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": achieved / fp64_peak if fp64_peak else None,
                    "peak_source": "FP64 peak measured now by the library's DFMA/DMMA micro-kernels (dfma %.1f, dmma %.1f "
                                   "TFLOP/s); MEASURED_PEAKS.json carries no FP64 figure" % (peaks["fp64_dfma_tflops"], peaks["fp64_dmma_tflops"])}
        roof.update({
            "traffic": ncu_traffic_bytes() if (n_local, k, d, args.dtype) == (N_PER_GPU, K_CLUSTERS, D, "f64") else None,
            "traffic_note": "DRAM read+write bytes of one assignment launch from the committed ncu --set full capture "
                            "(profiles/ncu_r1_assign_dmma_final_summary.csv); algorithmic bytes per launch = n*(d*s+4) = %.3e" % hbm_bytes,
            "kernel": "assignment kernel (dominant), CUDA events on the library stream, mean of %d launches" % args.steps,
            "kernel_ms": 1e3 * t_assign, "flop_per_byte": intensity, "hbm": hbm_view})
        line = {
            "metric": "lloyd_point_iters_per_sec", "value": value, "unit": "point-iters/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "C3 blobs 10M x 64 k=256 f64 per GPU (BASELINE.json configs[2])"
                                   if (n_local, k, d, args.dtype) == (N_PER_GPU, K_CLUSTERS, D, "f64")
                                   else "tuning shape %d x %d k=%d %s per GPU" % (n_local, d, k, args.dtype), "n_per_gpu": n_local,
                       "n_global": n_global, "d": d, "k": k, "l2": "inputs (5.12 GB/GPU) larger than L2; no flush",
                       "parallelism": "rows sharded x%d, one NCCL all-reduce of k*d+k+1 f64 per step" % world,
                       "kmeanspp_init_s": t_init, "wall_s_timed_region": wall},
            "roofline": roof,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "tf32_ranked": alt,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if distributed:
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
