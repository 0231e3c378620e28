"""bench.py -- EM throughput of the flat J=800 fit on bun000 (BASELINE.json configs[1]) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU reference arm (the oracle's C port, all host threads)

A "step" is ONE fit: 10 EM iterations (the reference's solve(..., 10, ...), gmm_kernels.cu:588) of the
full-covariance J=800 mixture over the 40 256-point cloud (per rank: weak scaling, every rank owns a
same-size shard and the O(J) sufficient statistics are all-reduced every iteration).
value  = points x EM-iterations processed by all ranks per second, cloud resident in HBM.
e2e    = the same through the public API with HOST buffers (H2D of the cloud + init, D2H of the model each step).
Prints one JSON line on rank 0.
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
PKG = os.path.join(ROOT, "gpu-accelerated-point-cloud-registration-using-hierarchical-gmm_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

J = 800
EM_ITERS = 10
SIGMA0_SQ = 1e-4
METRIC = "flat GMM EM throughput, J=800 full-cov on bun000 (EM iters/s x N pts)"
UNIT = "Mpoint-iters/s"


def load_cloud():
    path = os.path.join(ROOT, "tests", "golden", "bun000_xyz.npy")
    if os.path.exists(path):
        return np.load(path).astype(np.float32), "bun000.ply vertices (40256 pts, committed fixture)"
    from oracle import synth
    return synth.bunny_like(40256, seed=0), "synthetic bunny-like surface (40256 pts)"


def init_model(X, seed=1):
    rng = np.random.default_rng(seed)
    mu0 = X[rng.choice(len(X), J, replace=False)].astype(np.float32)
    cov0 = np.tile(np.eye(3, dtype=np.float32) * np.float32(SIGMA0_SQ), (J, 1, 1))
    w0 = np.full(J, 1.0 / J, np.float32)
    return mu0, cov0, w0


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args):
    """CPU arm: the oracle's plain-C/OpenMP port of the same fit on all host threads (the reference has no CPU
    implementation of its full-covariance fitter; its CUDA binary is timed separately)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle
    X, src = load_cloud()
    mu0, _, _ = init_model(X)
    # use whichever of {all logical CPUs, half of them (one per physical core)} runs the fit faster
    best = None
    for nt in sorted({os.cpu_count() or 1, max(1, (os.cpu_count() or 2) // 2)}):
        c_oracle.set_threads(nt)
        c_oracle.flat_fit(X, mu0, 1, SIGMA0_SQ)
        t0 = time.perf_counter()
        c_oracle.flat_fit(X, mu0, 2, SIGMA0_SQ)
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, nt)
    c_oracle.set_threads(best[1])
    cores = best[1]
    for _ in range(max(args.warmup, 1)):
        c_oracle.flat_fit(X, mu0, 1, SIGMA0_SQ)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_oracle.flat_fit(X, mu0, EM_ITERS, SIGMA0_SQ)
    dt = time.perf_counter() - t0
    val = len(X) * EM_ITERS * args.steps / dt / 1e6
    sample = "full workload: %d steps x %d EM iterations x %d pts x J=%d" % (args.steps, EM_ITERS, len(X), J)
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": src,
           "config": {"workload": "configs[1]: flat GMM J=800 full-cov on bun000, 10 EM iterations per step", "points": int(len(X)),
                      "components": J, "em_iters_per_step": EM_ITERS},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "em_iters_per_sec": EM_ITERS * args.steps / dt}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own")
    ap.add_argument("--extras", type=int, default=1, help="also time the tree / registration workloads (rank 0, N=1, untimed region)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import hgmm_b200
    from hgmm_b200 import dist as hdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    eng = hgmm_b200.Engine(local, stream=stream.cuda_stream)
    if world > 1:
        hdist.attach_communicator(eng)

    X, src = load_cloud()
    mu0, cov0, w0 = init_model(X)      # the replicated model starts identical on every rank
    if world > 1:       # weak scaling: a same-size shard per rank (the cloud jittered by a rank-seeded 10 um)
        X = (X + np.random.default_rng(100 + rank).normal(0, 1e-5, X.shape)).astype(np.float32)
    n = len(X)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")      # 256 MiB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------ device-resident throughput (value)
    eng.set_points(torch.from_numpy(X).cuda())
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # warm-up: at least W fits, and keep the GPU busy until nvidia-smi has produced samples under load (the timed
    # region itself is only a few tens of milliseconds long)
    t_w = time.perf_counter()
    nw = 0
    while nw < args.warmup or (rank == 0 and world == 1 and len(sampler.rows) < 3 and time.perf_counter() - t_w < 3.0):
        eng.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=EM_ITERS, want_outputs=False)
        nw += 1
    if world > 1:       # same warm-up length on every rank
        for _ in range(200):
            eng.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=EM_ITERS, want_outputs=False)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = eng.launch_count
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(k & 255)                   # L2 flush between steps, outside the timed events
            ev[k][0].record(stream)
        eng.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=EM_ITERS, want_outputs=False)
        ev[k][1].record(stream)
    barrier()
    launches = eng.launch_count - launches0
    t_ms = sum(a.elapsed_time(b) for a, b in ev)
    tt = torch.tensor([t_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms = float(tt.item())
    total_pts = n * world
    value = total_pts * EM_ITERS * args.steps / (t_ms * 1e-3) / 1e6

    # ------------------------------------------------------------ end to end through the public API, host buffers
    Xpin = torch.from_numpy(X).pin_memory()
    for _ in range(2):
        eng.set_points(Xpin)
        eng.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=EM_ITERS)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        eng.set_points(Xpin)                                                     # H2D of the step's cloud
        res = eng.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=EM_ITERS)    # H2D init, D2H fitted model + log-lik
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = total_pts * EM_ITERS * args.steps / float(te.item()) / 1e6
    h2d = n * 12 + J * (3 + 9 + 1) * 4
    d2h = J * (3 + 9 + 1) * 4 + EM_ITERS * 8 + 8 * 4

    # ------------------------------------------------------------ per-kernel roofline (separate, profiled pass)
    eng.set_points(torch.from_numpy(X).cuda())
    eng.set_profiling(True)
    k_ms, k_cnt = 0.0, 0
    for _ in range(5):
        with torch.cuda.stream(stream):
            flush.fill_(1)
        eng.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=EM_ITERS, want_outputs=False)
        tm = eng.last_timing_ms()
        k_ms += tm[1]
        k_cnt += int(tm[2])
    eng.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    k_avg_s = k_ms / max(k_cnt, 1) * 1e-3
    bytes_alg = 12.0 * n + 104.0 * J                 # 12 B/point + 104 B/component per sweep
    flops_alg = 52.0 * n * J                         # 52 flop per (point, component) pair
    hbm_peak, hbm_src = measured_peaks()
    peak_imm, peak_reg, peak_packed = eng.measure_fp32_peak()
    fp32_peak = peak_packed
    roof = {"bound": "hbm", "kernel": "em_flat7_kernel", "achieved": bytes_alg / k_avg_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": bytes_alg / k_avg_s / 1e9 / hbm_peak, "traffic": None, "peak_source": hbm_src,
            "avg_launch_us": k_avg_s * 1e6,
            "note": "J=800 makes this sweep FP32-issue bound (52*J/12 = 3467 flop/B >> the ~10 flop/B ridge); see roofline_fp32"}
    roof32 = {"bound": "fp32", "kernel": "em_flat7_kernel", "achieved": flops_alg / k_avg_s / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
              "frac": flops_alg / k_avg_s / 1e12 / fp32_peak if fp32_peak > 0 else None,
              "peak_source": "measured live: packed FFMA2 register loop (hgmm_measure_fp32_peak); scalar 3-register FFMA measures "
                             "%.1f, immediate-operand FFMA %.1f TFLOP/s on the same device" % (peak_reg, peak_imm)}
    prof_json = os.path.join(ROOT, "profiles", "em_flat_traffic.json")
    if os.path.exists(prof_json):
        try:
            roof["traffic"] = json.load(open(prof_json)).get("dram_bytes_per_launch")
        except Exception:
            pass

    # ------------------------------------------------------------ CPU baseline (rank 0, N=1) + other workloads
    cpu = None
    extras = {}
    if rank == 0 and world == 1:
        try:
            from oracle import c_oracle
            best = None
            for nt in sorted({os.cpu_count() or 1, max(1, (os.cpu_count() or 2) // 2)}):
                c_oracle.set_threads(nt)
                c_oracle.flat_fit(X, mu0, 1, SIGMA0_SQ)
                t0 = time.perf_counter()
                c_oracle.flat_fit(X, mu0, 2, SIGMA0_SQ)
                dtt = time.perf_counter() - t0
                if best is None or dtt < best[0]:
                    best = (dtt, nt)
            cores = best[1]
            c_oracle.set_threads(cores)
            t0 = time.perf_counter()
            reps = 0
            while time.perf_counter() - t0 < 10.0 and reps < 8:
                c_oracle.flat_fit(X, mu0, EM_ITERS, SIGMA0_SQ)
                reps += 1
            dt = time.perf_counter() - t0
            cpu = {"value": n * EM_ITERS * reps / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d full fits (10 EM iterations, %d pts, J=%d) of oracle/c/em_oracle.c, fp64, OpenMP" % (reps, n, J)}
        except Exception as e:      # the baseline is reported, never required for the GPU number
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
        if args.extras:
            try:
                extras = other_workloads(eng, stream)
            except Exception as e:
                extras = {"error": repr(e)}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": src,
               "config": {"workload": "configs[1]: flat GMM J=800 full-cov on bun000, 10 EM iterations per step",
                          "points_per_gpu": n, "components": J, "em_iters_per_step": EM_ITERS, "init": "seeded points, Sigma0=1e-4*I",
                          "l2": "256 MiB buffer written between steps (the 10 sweeps inside a step re-read the 483 kB cloud as the algorithm does)",
                          "parallelism": ("dp%d: points sharded; per EM iteration the J*10 fp64 moments are exchanged " % world) +
                                         ("inside the M-step kernel by NVLink stores into peer memory (no NCCL call)" if eng.p2p_enabled
                                          else "by one fp64 ncclAllReduce" if world > 1 else "(single rank: no exchange)")},
               "em_iters_per_sec": EM_ITERS * args.steps / (t_ms * 1e-3),
               "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": float(te.item()) / args.steps * 1e3, "timing": "wall clock, barrier+synchronize on both sides"},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "roofline_fp32": roof32, "cpu_baseline": cpu,
               "other_workloads": extras}
        print(json.dumps(out))
    if world > 1:
        eng.comm_destroy()
        dist.destroy_process_group()


def other_workloads(eng, stream):
    """configs[2] and configs[3] timed once each on the device (CUDA events via the library), for context."""
    import torch
    from oracle import synth
    from hgmm_b200 import hgmm as H
    out = {}
    # configs[2]: HGMM depth 4 on a 100k-point synthetic LiDAR sweep
    P = synth.lidar_sweep(100000, seed=2024)
    init = P[H.reference_init_indices(4)]
    eng.set_points(torch.from_numpy(P).cuda())
    for mode in ("estep", "level"):
        eng.fit_tree(init, 4, ls=20.0, ld=1e-4, sig2=4.0, ll_mode=mode, want_current=False, want_outputs=False)
        r = eng.fit_tree(init, 4, ls=20.0, ld=1e-4, sig2=4.0, ll_mode=mode, want_current=False, want_outputs=False)
        ms = float(eng.last_timing_ms()[0])
        its = int(r["iters"].sum())
        out["tree_L4_100k_lidar_ll_%s" % mode] = {"build_ms": ms, "em_iterations": its, "iters_per_level": r["iters"].tolist(),
                                                  "em_iters_per_sec": its / (ms * 1e-3)}
    # configs[3]: registration bun000 -> bun045 against a depth-3 tree (fit + <=20 iterations)
    b0 = os.path.join(ROOT, "tests", "golden", "bun000_xyz.npy")
    b45 = os.path.join(ROOT, "tests", "golden", "bun045_xyz.npy")
    if os.path.exists(b0) and os.path.exists(b45):
        S, T = np.load(b0), np.load(b45)
        init = S[H.reference_init_indices(3)]
        eng.set_points(torch.from_numpy(S).cuda())
        eng.reg_set_target(torch.from_numpy(T).cuda())
        for _ in range(2):
            t0 = time.perf_counter()
            eng.fit_tree(init, 3, ls=20.0, ld=1e-4, sig2=4e-4, ll_mode="estep", want_current=False, want_outputs=False)
            fit_ms = float(eng.last_timing_ms()[0])
            rot, t, q, it, _ = eng.register_tree(solver="twist_lstsq", maxiter=20, tol=1e-4)
            reg_ms = float(eng.last_timing_ms()[0])
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
        out["registration_bun000_bun045_L3"] = {"tree_fit_ms": fit_ms, "register_ms": reg_ms, "iterations": it,
                                                "fps_device": 1e3 / (fit_ms + reg_ms), "fps_wall": 1.0 / wall}
    return out


if __name__ == "__main__":
    main()
