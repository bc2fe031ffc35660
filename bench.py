#!/usr/bin/env python
"""bench.py — headline benchmark of the b200zk proving backend.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload prove|msm|ntt] [--k K] [--log-n L]
  python bench.py --impl reference ...     # CPU arm: the oracle's restatement of halo2's CPU
                                           # prover / best_multiexp / best_fft on this box's cores

Workloads (one "step" = one pass of the hot path over one batch of synthetic input):
  prove (default) : create_proof for the reference's Merkle Sum Tree circuit (chips.py: Poseidon
                    width 5, LtChip, u8 lookups; inclusion path of a 2^16-leaf tree, synthetic
                    values) padded to k = 20 — BASELINE.json's headline "create_proof ms (MST k=20)".
                    --circuit mst_dense is the same shape with every row in use (dense witness).
                    SRS and proving key resident in HBM.
  msm             : one best_multiexp over 2^L uniform scalars and 2^L SRS points
  ntt             : one best_fft over 2^L uniform Fr elements
Prints ONE JSON line: value = device-timed with inputs resident in HBM; e2e = the same call through
the host-buffer C-ABI entry point (pinned host witness / scalars -> H2D -> kernels -> D2H proof);
roofline for the dominant kernel; cpu_baseline on rank 0.  With N > 1 every rank proves its own
instance (weak scaling, no data-path collective); time is the max over ranks.
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# measured on this pool's B200 with tools/imad_bench.cu (profiles/r01_imad_microbench.jsonl):
# IMAD.WIDE.U32 issues at 32 lanes/clk/SM -> 9.19e12 wide multiply-adds per second
IMAD_WIDE_PEAK = 9.19e12
MUL32_PER_FIELD_MUL = 136           # 8x8 product + 8x8 reduction + 8 (m = t0 * inv)


def msm_work_mul32(n):
    """SURVEY.md §8(d) accounting (c = 16, W = 16): (N*16*11 + 2*65536*16*16) * 136 mul32."""
    return (n * 16 * 11 + 2 * 65536 * 16 * 16) * MUL32_PER_FIELD_MUL


def ntt_digits(log_n):
    """Pass structure the library plans for a transform of 2^log_n (csrc/ntt_plan.hpp)."""
    if 10 <= log_n <= 21:
        P = (log_n + 7) // 8
    else:
        P = min(3, max(1, (log_n + 9) // 10))
    return [log_n // P + (1 if p < log_n % P else 0) for p in range(P)]


def ntt_passes(log_n):
    return len(ntt_digits(log_n))


def ntt_work_mul32(log_n):
    """Field multiplications one transform needs as executed: butterflies whose twiddle is not 1
    (stage with half-size 2^lh: a fraction 1 - 2^-lh of N/2) plus one inter-pass twiddle per element
    and pass boundary; 132 IMAD.WIDE-equivalents each (128 wide + 8 low IMAD at twice the rate)."""
    n = 1 << log_n
    d = ntt_digits(log_n)
    muls = sum(n / 2 * (1 - 2.0 ** -lh) for m in d for lh in range(m)) + n * (len(d) - 1)
    return muls * 132


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
        sm, mx, pw, reasons = [], [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "power_w_max": max(pw), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def dist_setup():
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


def timed(be, dist, local, steps, fn):
    """CUDA events on the library's stream around `steps` calls, max over ranks -> total ms."""
    barrier(dist, be)
    be.event_record(0)
    for _ in range(steps):
        fn()
    be.event_record(1)
    be.sync()
    ms = be.event_elapsed_ms(0, 1)
    barrier(dist, be)
    return max_over_ranks(dist, local, ms)


CIRCUITS = {"mst": ("merkle_sum_tree_job", "the reference's MerkleSumTreeCircuit (Poseidon width 5, LtChip, 16-level path = tree of 2^16 leaves) padded to k={k}: 20 advice, 15 fixed, 8 u8 lookups, 16 permutation columns, degree 6; synthetic leaf / sibling values"),
            "v3": ("merkle_v3_job", "the reference's MerkleTreeV3Circuit (Poseidon width 3, 13-level path) at k={k}: 7 advice, 10 fixed, 10 permutation columns, degree 6"),
            "mst_dense": ("mst_shaped", "MST-shaped synthetic circuit k={k} with every row in use (dense witness, worst case for the commits): 20 advice, 8 u8 lookups, 16 permutation columns, degree 6"),
            "v3_dense": ("v3_shaped", "Merkle-v3-shaped synthetic circuit k={k} with every row in use: 9 advice, no lookups, 12 permutation columns, degree 6")}


def build_job(zk, circuit, k, seed):
    """Dense synthetic circuits come from circuits_synth, the reference's real circuits from chips."""
    mod = importlib.import_module(zk.__name__ + (".circuits_synth" if circuit.endswith("_dense") else ".chips"))
    return getattr(mod, CIRCUITS[circuit][0])(k, seed=seed)
WORKLOAD_TEXT = {
    "prove": "create_proof (KZG/SHPLONK/Blake2b), {circuit}; SRS + pk resident in HBM",
    "msm": "bn256 G1 MSM 2^{L} points (best_multiexp drop-in), uniform scalars, bases [s^i]G resident in HBM",
    "ntt": "bn256 Fr NTT 2^{L} (best_fft drop-in), uniform input",
}


# --------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    from __graft_entry__ import load_package
    zk = load_package()
    rank, world, local, dist = dist_setup()
    be = zk.Backend(local)
    peaks, peak_src = measured_peaks()
    line = {"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "scaling": "weak", "vs_baseline": None,
            "data": "synthetic", "impl": "b200zk"}
    extra = {}
    if args.workload == "prove":
        k = args.k
        n = 1 << k
        synth = importlib.import_module(zk.__name__ + ".circuits_synth")
        job = build_job(zk, args.circuit, k, 1 + rank)
        params = zk.ParamsKZG.setup(be, k, random_scalars(1, 4242)[0])
        pk = zk.ProvingKey(params, job.cs, k, job.fixed, job.map_col, job.map_row)
        A = job.cs.num_advice
        h_adv = be.pinned_empty((A * n, 4))
        for c in range(A):
            h_adv[c * n:(c + 1) * n] = job.advice[c]
        h_cols = [h_adv[c * n:(c + 1) * n] for c in range(A)]
        h_wide = be.pinned_empty((pk.rng_draws, 8))
        h_wide[:] = np.random.Generator(np.random.PCG64(99 + rank)).integers(0, 1 << 64, size=(pk.rng_draws, 8), dtype=np.uint64)
        d_adv, d_wide = be.to_device(h_adv), be.to_device(h_wide)
        lut = synth.mont_from_ints(job.instances[0] + [job.transcript_repr])
        inst, tr_repr = [lut[:-1]], lut[-1]

        def step_dev():
            return pk.create_proof_dev(d_adv, inst, d_wide, tr_repr)

        def step_e2e():
            return pk.create_proof(h_cols, inst, h_wide, tr_repr)

        unit, metric, hib = "ms", "create_proof_ms", False
        h2d, d2h = int(h_adv.nbytes + h_wide.nbytes), pk.proof_size
        dtype = "u32x8 Montgomery (bn256 Fr/Fq, IMAD pipe)"
        workload = WORKLOAD_TEXT["prove"].format(circuit=CIRCUITS[args.circuit][1].format(k=k))
    elif args.workload == "msm":
        # N > 1: ONE MSM of 2^L points sharded by point range (strong scaling): every rank holds
        # 2^L / N bases and scalars, partial sums are all-gathered (96 B per rank) and added.
        L = args.log_n
        shard_log = L - int(np.log2(world))
        n = 1 << shard_log
        params = zk.ParamsKZG.setup(be, shard_log, random_scalars(1, 4242 + rank)[0])
        h_scalars = be.pinned_empty((n, 4))
        random_scalars(n, 100 + rank, out=h_scalars)
        d_scalars = be.to_device(h_scalars)
        sharded = importlib.import_module(zk.__name__ + ".sharded")
        device = None
        if dist is not None:
            import torch
            device = torch.device("cuda", local)
        sc = sharded.ShardedCommit(sharded.GpuCommitEngine(zk, params), device)

        def step_dev():
            return sc.commit(d_scalars)

        def step_e2e():
            return sc.commit(h_scalars)

        unit, metric, hib = "Mpts/s", "msm_mpts_per_s", True
        units_per_step = n / 1e6                                  # per rank; value multiplies by world
        if world > 1:
            line["scaling"] = "strong"
        h2d, d2h = n * 32, 96
        dtype = "u32x8 Montgomery (bn256 Fq/Fr, IMAD pipe)"
        workload = WORKLOAD_TEXT["msm"].format(L=L)
    else:
        L = args.log_n
        n = 1 << L
        dom = zk.EvaluationDomain(be, 2, L)
        h_a = be.pinned_empty((n, 4))
        random_scalars(n, 200 + rank, out=h_a)
        d_a = be.to_device(h_a)
        omega = dom.omega

        def step_dev():
            be.best_fft_dev(d_a, omega, L)

        def step_e2e():
            be._check(zk.lib().b200zk_fft(be._ctx, h_a.ctypes.data_as(ctypes.c_void_p), omega.ctypes.data_as(ctypes.c_void_p), L))

        if world > 1:
            # ONE transform of 2^L elements sharded four-step (strong scaling): column blocks ->
            # column step -> NCCL all-to-all over NVLink -> row step (halo2-experiments_b200/sharded.py)
            import torch
            sharded = importlib.import_module(zk.__name__ + ".sharded")
            dev = torch.device("cuda", local)
            log_r = min(10, L // 2)
            # default: the exchange fused into the column-step kernel (NVLink peer stores, CUDA IPC);
            # B200ZK_NTT_NCCL=1 selects the NCCL all-to-all path
            fused = os.environ.get("B200ZK_NTT_NCCL", "0") != "1"
            fs = (sharded.FourStepNTTFused if fused else sharded.FourStepNTTDevice)(zk, be, L, log_r, rank, world, dev)
            extra["exchange"] = "fused into the column-step kernel (NVLink peer stores)" if fused else "NCCL all_to_all_single"
            blk_host = random_scalars((1 << L) // world, 300 + rank).reshape(1 << log_r, -1, 4)
            block = torch.from_numpy(blk_host.view(np.int64)).to(dev)
            omega_c = zk.EvaluationDomain(be, 2, L - log_r).omega
            pinned = torch.from_numpy(blk_host.view(np.int64)).pin_memory()
            pinned_np = pinned.numpy()

            def step_dev():
                fs.forward(block, omega, omega_c)

            def step_e2e():
                block.copy_(pinned, non_blocking=False)
                rows = fs.forward(block, omega, omega_c)
                if fused:
                    be._check(zk.lib().b200zk_download(be._ctx, pinned_np.ctypes.data_as(ctypes.c_void_p), rows.ptr, ctypes.c_size_t(pinned_np.nbytes)))
                else:
                    pinned.view(rows.shape).copy_(rows)

            n = (1 << L) // world                                   # per-rank elements; value multiplies by world
            line["scaling"] = "strong"

        unit, metric, hib = "GB/s", "ntt_gb_per_s", True
        units_per_step = 64.0 * n / 1e9
        h2d, d2h = n * 32, n * 32
        dtype = "u32x8 Montgomery (bn256 Fr, IMAD pipe)"
        workload = WORKLOAD_TEXT["ntt"].format(L=L)

    # ---- device-timed region: inputs resident in HBM --------------------------------
    for _ in range(max(args.warmup, 3)):
        step_dev()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = be.launch_count()
    ms = timed(be, dist, local, args.steps, step_dev)
    launches = (be.launch_count() - l0) // args.steps
    if args.workload == "prove":
        extra["phase_ms"] = pk.last_phase_ms()
        extra["phase_note"] = "device time per phase; the advice-coset NTTs run on a side stream concurrently with the lookup / permutation phases, so the phases overlap and sum to more than the step"
    # ---- end-to-end region: host buffers through the C ABI ---------------------------
    for _ in range(2):
        step_e2e()
    barrier(dist, be)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    be.sync()
    e2e_ms = max_over_ranks(dist, local, (time.perf_counter() - t0) * 1e3) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    if args.profile_step:                       # one extra step inside cudaProfilerStart/Stop for ncu
        be.profiler_range(True)
        step_dev()
        be.profiler_range(False)

    ms_per_step = ms / args.steps
    if args.workload == "prove":
        value, e2e_value = ms_per_step, e2e_ms
        extra["proofs_per_s_all_gpus"] = world / (ms_per_step / 1e3)
    else:
        value = units_per_step * world / (ms_per_step / 1e3)
        e2e_value = units_per_step * world / (e2e_ms / 1e3)
    line.update({"metric": metric, "unit": unit, "value": value, "ms_per_step": ms_per_step, "higher_is_better": hib, "dtype": dtype,
                 "config": {"workload": workload, "l2": "working set larger than L2 (126 MB); steps timed back to back"},
                 "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms},
                 "gpu_launches": int(launches), "clocks": clocks})

    # ---- roofline of the dominant kernel ------------------------------------------------
    imad_note = ("no HBM/tensor bound applies: 254-bit Montgomery arithmetic is IMAD-pipe bound; peak = measured IMAD.WIDE.U32 issue "
                 "rate on this B200 (tools/imad_bench.cu, profiles/r01_imad_microbench.jsonl)")
    if args.workload == "prove":
        # dominant kernel = the NTT pass kernel (per-coset size-n transforms of coeff_to_extended and
        # the iNTTs): time one size-n transform on device data with events
        omega = synth.mont_from_ints([pow(pow(7, (synth.R_MOD - 1) >> 28, synth.R_MOD), 1 << (28 - k), synth.R_MOD)])[0]
        d_vec = be.to_device(random_scalars(n, 5))
        for _ in range(3):
            be.best_fft_dev(d_vec, omega, k)
        reps = 20
        ms_ntt = timed(be, dist, local, reps, lambda: be.best_fft_dev(d_vec, omega, k)) / reps
        ph = extra["phase_ms"]
        ach = ntt_work_mul32(k) / (ms_ntt / 1e3)
        hbm = 64.0 * n / (ms_ntt / 1e3) / 1e9
        line["roofline"] = {"bound": "imad", "kernel": f"ntt pass kernel, one 2^{k} transform = {ntt_passes(k)} launches", "achieved": ach / 1e12,
                            "peak": IMAD_WIDE_PEAK / 1e12, "unit": "T IMAD.WIDE.U32/s", "frac": ach / IMAD_WIDE_PEAK,
                            # dram__bytes_read + write per pass launch of a 2^20 transform, ncu --set full
                            # (profiles/r01_ncu_ntt_warp_summary.txt): the 32 MB transform and its 32 MB twiddle
                            # table stay in L2, so traffic is below the 67 MB algorithmic bytes of a pass
                            "traffic": ({"dram_bytes_per_launch": [67.7e6, 34.1e6, 35.0e6], "algorithmic_bytes_per_launch": 64 * n,
                                         "source": "profiles/r01_ncu_ntt_warp_summary.txt"} if k == 20 else None),
                            "ms_per_launch_group": ms_ntt, "share_of_step": ph["ntt"] / max(sum(ph.values()), 1e-9),
                            "hbm_view": {"achieved": hbm, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm / peaks["hbm_gbs"],
                                         "algorithmic_bytes": 64 * n, "peak_source": peak_src},
                            "note": imad_note + "; work = non-trivial butterfly and inter-pass twiddle multiplications x 132 (DESIGN.md §4)"}
        # the launch shape the proof mostly runs (5/6 of its NTT work): the q quotient cosets of one column as ONE
        # batched launch per pass — here q independent 2^k transforms through b200zk_fft_rows_dev
        q = pk.degree - 1
        d_rows = be.to_device(random_scalars(q * n, 7))
        rows_call = lambda: be._check(zk.lib().b200zk_fft_rows_dev(be._ctx, d_rows.ptr, ctypes.c_uint32(q), zk._p(zk._fr(omega, 1)), ctypes.c_uint32(k)))
        for _ in range(3):
            rows_call()
        ms_rows = timed(be, dist, local, reps, rows_call) / reps
        line["roofline"]["batched"] = {"kernel": f"same kernel, {q} transforms of 2^{k} per launch (coeff_to_extended's quotient cosets)",
                                       "ms_per_launch_group": ms_rows, "achieved": q * ntt_work_mul32(k) / (ms_rows / 1e3) / 1e12,
                                       "frac": q * ntt_work_mul32(k) / (ms_rows / 1e3) / IMAD_WIDE_PEAK}
        d_rows.free()
        d_dense = be.to_device(random_scalars(n, 6))
        for _ in range(3):
            params.commit_dev(d_dense, n, lagrange=False)
        ms_msm = timed(be, dist, local, 5, lambda: params.commit_dev(d_dense, n, lagrange=False)) / 5
        line["roofline_msm"] = {"bound": "imad", "kernel": f"dense commit 2^{k} (msm_accumulate_task dominates)",
                                "achieved": msm_work_mul32(n) / (ms_msm / 1e3) / 1e12, "peak": IMAD_WIDE_PEAK / 1e12, "unit": "T IMAD.WIDE.U32/s",
                                "frac": msm_work_mul32(n) / (ms_msm / 1e3) / IMAD_WIDE_PEAK, "ms_per_launch_group": ms_msm,
                                "share_of_step": ph["msm"] / max(sum(ph.values()), 1e-9), "note": "work per SURVEY.md 8(d)"}
    elif args.workload == "msm":
        ach = msm_work_mul32(n) / (ms_per_step / 1e3)
        line["roofline"] = {"bound": "imad", "kernel": f"MSM 2^{int(np.log2(n))} (msm_accumulate dominates)", "achieved": ach / 1e12,
                            "peak": IMAD_WIDE_PEAK / 1e12, "unit": "T IMAD.WIDE.U32/s", "frac": ach / IMAD_WIDE_PEAK, "traffic": None,
                            "ms_per_launch_group": ms_per_step, "share_of_step": 1.0, "note": imad_note + "; work per SURVEY.md 8(d)"}
    else:
        ach = 64.0 * n / (ms_per_step / 1e3) / 1e9
        line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                            "traffic": None, "peak_source": peak_src,
                            "note": "254-bit butterflies are IMAD-bound on B200 (see DESIGN.md): 64N bytes vs ~15N field muls"}
    line.update(extra)
    if rank == 0:
        line["cpu_baseline"] = cpu_baseline(args)
        emit(line)
    barrier(dist, be)
    be.close()
    if dist is not None:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- CPU legs
def cpu_sample(args):
    """(callable, units, description, scale) for the oracle's restatement on a bounded sample."""
    from __graft_entry__ import load_package
    from oracle import binding as orc
    from oracle import pyref
    orc.build()
    if args.workload == "prove":
        from oracle import prover as OP
        zk = load_package()
        ks = min(args.k, args.cpu_k)
        job = build_job(zk, args.circuit, ks, 1)
        g, gl = orc.params_setup(ks, orc.random_fr(1, 4242)[0])
        pk = OP.keygen_pk(job.cs, ks, job.fixed, job.map_col, job.map_row)
        wide = np.random.Generator(np.random.PCG64(99)).integers(0, 1 << 64, size=(OP.rng_draws_needed(job.cs, ks), 8), dtype=np.uint64)
        fn = lambda: OP.create_proof(g, gl, pk, job.advice, job.instances, wide, job.transcript_repr)
        # scale from the sample to the full size: a complete oracle run of this circuit at k = 20 took
        # 147.7 s on the 16 host cores of the GPU box against 3.5 s at k = 14 (profiles/
        # r01_parity_mst_k20_gpu_vs_oracle.json), i.e. x42 rather than the x64 of the row count
        scale = float(1 << (args.k - ks))
        if args.circuit == "mst" and args.k == 20 and ks == 14:
            scale = 42.2
        if args.circuit == "mst_dense" and args.k == 20 and ks == 14:
            scale = 47.0                                         # 112.8 s at k = 20 (r01_parity_mst_dense_k20_gpu_vs_oracle.json) / 2.4 s
        return fn, None, f"oracle create_proof (restatement of halo2 v2023_02_02 CPU prover) on the same circuit at k={ks}", scale
    Ls = min(args.log_n, args.cpu_log_n)
    ns = 1 << Ls
    if args.workload == "msm":
        bases, _ = orc.params_setup(Ls, orc.random_fr(1, 4242)[0], with_lagrange=False)
        sc = random_scalars(ns, 7)
        return (lambda: orc.best_multiexp(sc, bases)), ns / 1e6, f"best_multiexp restatement on 2^{Ls} points", 1.0
    a = random_scalars(ns, 8)
    w = orc.ints_to_mont([pyref.omega_for_k(Ls)])[0]
    return (lambda: orc.best_fft(a, w, Ls)), 64.0 * ns / 1e9, f"best_fft restatement on 2^{Ls} elements", 1.0


def cpu_baseline(args):
    from oracle import binding as orc
    fn, units, desc, scale = cpu_sample(args)
    t0 = time.perf_counter()
    fn()
    dt = time.perf_counter() - t0
    cores = orc.get_threads()
    if args.workload == "prove":
        return {"value": dt * 1e3 * scale, "unit": "ms", "cores": cores, "kind": "port", "measured_ms": dt * 1e3,
                "sample": f"{desc}: {dt:.1f} s measured, x{scale:g} to k={args.k} (factor measured once with a complete k=20 oracle run: 147.7 s on 16 cores, profiles/r01_parity_mst_k20_gpu_vs_oracle.json; x2^(k-ks) for other sizes)"}
    return {"value": units / dt, "unit": "Mpts/s" if args.workload == "msm" else "GB/s", "cores": cores, "kind": "port",
            "sample": f"{desc}, {dt:.2f} s"}


def run_reference(args):
    """--impl reference: halo2's CPU algorithm (oracle restatement; the Rust crate cannot be built
    here) on this box's host cores, same metric/config, each step a bounded sample."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import binding as orc
    fn, units, desc, scale = cpu_sample(args)
    cores = orc.get_threads()
    for _ in range(min(args.warmup, 1)):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = (time.perf_counter() - t0) / args.steps
    if args.workload == "prove":
        v, unit, metric, hib = dt * 1e3 * scale, "ms", "create_proof_ms", False
        workload = WORKLOAD_TEXT["prove"].format(circuit=CIRCUITS[args.circuit][1].format(k=args.k))
        sample = f"{desc}: {dt:.1f} s per step measured, x{scale:g} to k={args.k} (factor from a complete k=20 oracle run, profiles/r01_parity_mst_k20_gpu_vs_oracle.json)"
    else:
        v = units / dt
        unit, metric, hib = ("Mpts/s", "msm_mpts_per_s", True) if args.workload == "msm" else ("GB/s", "ntt_gb_per_s", True)
        workload = WORKLOAD_TEXT[args.workload].format(L=args.log_n)
        sample = f"{desc} per step"
    emit({"impl": "reference", "metric": metric, "unit": unit, "value": v, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": hib,
                      "scaling": "weak", "vs_baseline": None, "dtype": "u64x4 Montgomery (CPU)", "data": "synthetic",
                      "config": {"workload": workload},
                      "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


_JSON_OUT = None


def emit(line):
    """The one JSON line of the contract, on the process's real stdout."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries exactly one JSON line: libraries that write to fd 1 (NCCL prints its version banner there
    # under torchrun) are sent to stderr, the line itself goes to a duplicate of the original descriptor
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200zk", choices=["b200zk", "reference"])
    ap.add_argument("--workload", default="prove", choices=["prove", "msm", "ntt"])
    ap.add_argument("--k", type=int, default=20, help="prove: circuit size")
    ap.add_argument("--circuit", default="mst", choices=sorted(CIRCUITS), help="prove: circuit shape")
    ap.add_argument("--cpu-k", type=int, default=14, help="prove: size of the bounded CPU sample")
    ap.add_argument("--log-n", type=int, default=24, help="msm / ntt size")
    ap.add_argument("--cpu-log-n", type=int, default=None, help="msm / ntt: size of the bounded CPU sample")
    ap.add_argument("--profile-step", action="store_true", help="bracket one extra step with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.cpu_log_n is None:
        args.cpu_log_n = 20 if args.workload == "msm" else 22
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
