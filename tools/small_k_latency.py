#!/usr/bin/env python
"""Small-circuit latency of create_proof (BASELINE config 2: Merkle v3 k=14; the reference's own test size, MST k=9):
ms per proof (device-resident inputs), launches per proof and the host timeline of one proof."""
import importlib, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
from bench import random_scalars, build_job
zk = load_package()
be = zk.Backend(0)
synth = importlib.import_module(zk.__name__ + ".circuits_synth")
for circuit, k in (("v3", 14), ("mst", 9), ("mst", 14)):
    chips = importlib.import_module(zk.__name__ + ".chips")
    job = chips.merkle_sum_tree_job(k, levels=5, seed=1) if (circuit, k) == ("mst", 9) else build_job(zk, circuit, k, 1)
    n = 1 << k
    params = zk.ParamsKZG.setup(be, k, random_scalars(1, 4242)[0])
    pk = zk.ProvingKey(params, job.cs, k, job.fixed, job.map_col, job.map_row)
    adv = np.concatenate([np.ascontiguousarray(a).reshape(-1, 4) for a in job.advice])
    d_adv = be.to_device(adv)
    wide = np.random.Generator(np.random.PCG64(99)).integers(0, 1 << 64, size=(pk.rng_draws, 8), dtype=np.uint64)
    d_wide = be.to_device(wide)
    lut = synth.mont_from_ints(job.instances[0] + [job.transcript_repr])
    inst, tr = [lut[:-1]], lut[-1]
    for _ in range(5):
        pk.create_proof_dev(d_adv, inst, d_wide, tr)
    reps = 20
    l0 = be.launch_count()
    be.sync(); be.event_record(0)
    for _ in range(reps):
        pk.create_proof_dev(d_adv, inst, d_wide, tr)
    be.event_record(1); be.sync()
    ms = be.event_elapsed_ms(0, 1) / reps
    if "--profile" in sys.argv and (circuit, k) == ("v3", 14):      # one proof inside cudaProfilerStart/Stop (ncu --profile-from-start off)
        be.profiler_range(True); pk.create_proof_dev(d_adv, inst, d_wide, tr); be.profiler_range(False)
    print(json.dumps({"circuit": circuit, "k": k, "ms": round(ms, 3), "launches": (be.launch_count() - l0) // reps,
                      "timeline_ms": [(a, round(b, 2)) for a, b in pk.last_trace()], "phase_ms": {a: round(b, 2) for a, b in pk.last_phase_ms().items()}}), flush=True)
    d_adv.free(); d_wide.free(); pk.close(); params.close()
