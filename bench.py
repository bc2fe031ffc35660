#!/usr/bin/env python
"""bench.py — headline benchmark of the b200zk proving backend.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload msm|ntt] [--log-n L]
  python bench.py --impl reference ...        # CPU arm: the oracle's restatement of halo2's
                                              # rayon CPU path on this box's host cores

One "step" = one pass of the hot path over one batch of synthetic input:
  msm : one best_multiexp over 2^L uniform scalars and 2^L SRS points ([s^i]G, device-generated)
  ntt : one coeff_to_extended-sized best_fft over 2^L uniform Fr elements
Prints ONE JSON line (see the task contract): value = device-timed whole-job throughput with
inputs resident in HBM; e2e = the same through the host-buffer C-ABI call (pinned host scalars
-> H2D -> kernels -> D2H result) ; roofline for the dominant kernel; cpu_baseline on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# measured on this pool's B200 with tools/imad_bench.cu (profiles/r01_imad_microbench.jsonl):
# IMAD.WIDE.U32 issues at 32 lanes/clk/SM -> 9.19e12 wide multiply-adds per second
IMAD_WIDE_PEAK = 9.19e12
MUL32_PER_FIELD_MUL = 136           # 8x8 product + 8x8 reduction + 8 (m = t0 * inv)


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def random_scalars(n, seed, out=None):
    """n Montgomery-form Fr elements, uniform below r's top limb (so every value is < r)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64) if out is None else out
    if out is not None:
        out[:] = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] %= np.uint64(0x30644e72e131a029)
    return a


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local, dist


def max_over_ranks(dist, local, value):
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(dist, be):
    be.sync()
    if dist is not None:
        import torch
        dist.barrier()
        torch.cuda.synchronize()


# --------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    from __graft_entry__ import load_package
    zk = load_package()
    rank, world, local, dist = dist_setup(args.gpus)
    be = zk.Backend(local)
    peaks, peak_src = measured_peaks()
    L = args.log_n
    n = 1 << L
    line = {"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "data": "synthetic", "impl": "b200zk"}

    if args.workload == "msm":
        s = random_scalars(1, 4242)[0]
        params = zk.ParamsKZG.setup(be, L, s)                      # SRS resident in HBM, generated on device
        h_scalars = be.pinned_empty((n, 4))
        random_scalars(n, 100 + rank, out=h_scalars)
        d_scalars = be.to_device(h_scalars)
        launches0 = be.launch_count()

        def step_dev():
            return params.commit_dev(d_scalars, n, lagrange=False)

        def step_e2e():
            return params.commit(h_scalars)                        # H2D n*32 B + kernels + D2H 96 B

        unit, metric = "Mpts/s", "msm_mpts_per_s"
        units_per_step = n / 1e6
        h2d, d2h = n * 32, 96
        dtype = "u32x8 Montgomery (bn256 Fq/Fr, IMAD pipe)"
        workload = f"bn256 G1 MSM 2^{L} points (best_multiexp drop-in), uniform scalars, bases [s^i]G resident in HBM"
    else:
        dom = zk.EvaluationDomain(be, 2, L)
        h_a = be.pinned_empty((n, 4))
        random_scalars(n, 200 + rank, out=h_a)
        d_a = be.to_device(h_a)
        omega = dom.omega
        launches0 = be.launch_count()

        def step_dev():
            be.best_fft_dev(d_a, omega, L)

        def step_e2e():
            lib = zk.lib()
            be._check(lib.b200zk_fft(be._ctx, h_a.ctypes.data_as(__import__("ctypes").c_void_p),
                                     omega.ctypes.data_as(__import__("ctypes").c_void_p), L))

        unit, metric = "GB/s", "ntt_gb_per_s"
        units_per_step = 64.0 * n / 1e9                             # compulsory bytes: read + write once
        h2d, d2h = n * 32, n * 32
        dtype = "u32x8 Montgomery (bn256 Fr, IMAD pipe)"
        workload = f"bn256 Fr NTT 2^{L} (best_fft drop-in), uniform input"

    # ---- device-timed region: inputs resident in HBM --------------------------------
    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier(dist, be)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = be.launch_count()
    be.event_record(0)
    for _ in range(args.steps):
        step_dev()
    be.event_record(1)
    be.sync()
    ms_local = be.event_elapsed_ms(0, 1)
    launches = be.launch_count() - l0
    barrier(dist, be)
    ms = max_over_ranks(dist, local, ms_local)
    # ---- end-to-end region: host buffers through the C ABI ---------------------------
    for _ in range(2):
        step_e2e()
    barrier(dist, be)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    be.sync()
    e2e_ms = max_over_ranks(dist, local, (time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop() if rank == 0 else None

    ms_per_step = ms / args.steps
    value = units_per_step * world / (ms_per_step / 1e3)
    e2e_value = units_per_step * world / (e2e_ms / args.steps / 1e3)
    line.update({"metric": metric, "unit": unit, "value": value, "ms_per_step": ms_per_step, "dtype": dtype,
                 "config": {"workload": workload, "log_n": L, "l2": "inputs larger than L2" if n * 32 > 126e6 else "inputs fit L2; timed back to back"},
                 "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                         "ms_per_step": e2e_ms / args.steps},
                 "gpu_launches": int(launches), "clocks": clocks})

    if args.workload == "msm":
        # work model of SURVEY §8(d): c = 16, W = 16: (N*16*11 + 2*65536*16*16) * 136 mul32
        work = (n * 16 * 11 + 2 * 65536 * 16 * 16) * MUL32_PER_FIELD_MUL
        ach = work / (ms_per_step / 1e3)
        line["roofline"] = {"bound": "imad", "achieved": ach / 1e12, "peak": IMAD_WIDE_PEAK / 1e12, "unit": "T IMAD.WIDE/s",
                            "frac": ach / IMAD_WIDE_PEAK, "traffic": None,
                            "note": "peak = measured IMAD.WIDE.U32 issue rate (tools/imad_bench.cu); whole-MSM time, accumulate kernel dominates"}
    else:
        ach = 64.0 * n / (ms_per_step / 1e3) / 1e9
        line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                            "traffic": None, "peak_source": peak_src,
                            "note": "254-bit butterflies are IMAD-bound on B200 (see DESIGN.md): 64N bytes vs ~15N field muls"}

    if rank == 0:
        line["cpu_baseline"] = cpu_baseline(args, zk, be, params if args.workload == "msm" else None)
        print(json.dumps(line), flush=True)
    barrier(dist, be)
    be.close()
    if dist is not None:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- CPU legs
def cpu_baseline(args, zk, be, params):
    """The oracle's restatement of halo2's CPU algorithm, on a bounded sample, all host cores."""
    from oracle import binding as orc
    cores = orc.get_threads()
    if args.workload == "msm":
        Ls = min(args.log_n, args.cpu_log_n)
        ns = 1 << Ls
        g, _ = params.read(lagrange=False)
        bases = np.ascontiguousarray(g[:ns])
        sc = random_scalars(ns, 7)
        t0 = time.perf_counter()
        orc.best_multiexp(sc, bases)
        dt = time.perf_counter() - t0
        return {"value": ns / 1e6 / dt, "unit": "Mpts/s", "cores": cores, "kind": "port",
                "sample": f"best_multiexp restatement on 2^{Ls} of the same points, {dt:.2f} s"}
    Ls = min(args.log_n, args.cpu_log_n)
    ns = 1 << Ls
    a = random_scalars(ns, 8)
    from oracle import pyref
    w = orc.ints_to_mont([pyref.omega_for_k(Ls)])[0]
    t0 = time.perf_counter()
    orc.best_fft(a, w, Ls)
    dt = time.perf_counter() - t0
    return {"value": 64.0 * ns / 1e9 / dt, "unit": "GB/s", "cores": cores, "kind": "port",
            "sample": f"best_fft restatement on 2^{Ls} elements, {dt:.2f} s"}


def run_reference(args):
    """--impl reference: halo2's CPU algorithm (oracle restatement; the Rust crate cannot be built
    here) on this box's host cores, same metric/config, each step a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as orc
    from oracle import pyref
    orc.build()
    cores = orc.get_threads()
    Ls = min(args.log_n, args.cpu_log_n)
    ns = 1 << Ls
    if args.workload == "msm":
        s = orc.random_fr(1, 4242)[0]
        bases, _ = orc.params_setup(Ls, s, with_lagrange=False)
        sc = random_scalars(ns, 7)
        fn = lambda: orc.best_multiexp(sc, bases)
        units, unit, metric = ns / 1e6, "Mpts/s", "msm_mpts_per_s"
        workload = f"bn256 G1 MSM 2^{args.log_n} points (best_multiexp drop-in), uniform scalars, bases [s^i]G resident in HBM"
    else:
        a = random_scalars(ns, 8)
        w = orc.ints_to_mont([pyref.omega_for_k(Ls)])[0]
        fn = lambda: orc.best_fft(a, w, Ls)
        units, unit, metric = 64.0 * ns / 1e9, "GB/s", "ntt_gb_per_s"
        workload = f"bn256 Fr NTT 2^{args.log_n} (best_fft drop-in), uniform input"
    for _ in range(min(args.warmup, 1)):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = (time.perf_counter() - t0) / args.steps
    v = units / dt
    print(json.dumps({"impl": "reference", "metric": metric, "unit": unit, "value": v, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "u64x4 Montgomery (CPU)", "data": "synthetic",
                      "config": {"workload": workload, "log_n": args.log_n},
                      "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port",
                                       "sample": f"2^{Ls} units per step of the 2^{args.log_n} workload"},
                      "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200zk", choices=["b200zk", "reference"])
    ap.add_argument("--workload", default="msm", choices=["msm", "ntt"])
    ap.add_argument("--log-n", type=int, default=None)
    ap.add_argument("--cpu-log-n", type=int, default=None, help="size of the bounded CPU sample")
    args = ap.parse_args()
    if args.log_n is None:
        args.log_n = 24 if args.workload == "msm" else 24
    if args.cpu_log_n is None:
        args.cpu_log_n = 20 if args.workload == "msm" else 22
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
