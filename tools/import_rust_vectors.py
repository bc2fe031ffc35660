#!/usr/bin/env python
"""Converts a directory written by rust-shim/src/dump.rs (stock halo2's prove job + proof) into the .npz layout the
golden tests read:  python tools/import_rust_vectors.py <dump_dir> <name>  ->  tests/golden/rust_<name>.npz"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(dump_dir):
    m = json.load(open(os.path.join(dump_dir, "manifest.json")))
    if m.get("format") != "b200zk-rust-vectors-1":
        raise ValueError("not a b200zk rust vector dump")
    rd = lambda name, dt: np.fromfile(os.path.join(dump_dir, name), dtype=dt)
    k, n = m["k"], 1 << m["k"]
    out = {"k": np.uint32(k), "blob": rd("cs_blob.u32", "<u4"),
           "fixed": rd("fixed.fr", "<u8").reshape(m["num_fixed"], n, 4), "advice": rd("advice.fr", "<u8").reshape(m["num_advice"], n, 4),
           "map_col": rd("map_col.u32", "<u4").reshape(m["num_permutation"], n), "map_row": rd("map_row.u32", "<u4").reshape(m["num_permutation"], n),
           "instances": rd("instances.fr", "<u8").reshape(-1, 4), "instance_lens": np.array(m["instance_lens"], dtype=np.uint32),
           "rng_wide": rd("rng_wide.bin", "<u8").reshape(-1, 8), "srs_secret_wide": rd("srs_secret_wide.bin", "<u8").reshape(1, 8),
           "transcript_repr": rd("transcript_repr.fr", "<u8").reshape(4),
           "fixed_commitments": rd("fixed_commitments.g1", "<u8").reshape(-1, 8), "sigma_commitments": rd("sigma_commitments.g1", "<u8").reshape(-1, 8),
           "g2": rd("g2.g2", "<u8").reshape(16), "s_g2": rd("s_g2.g2", "<u8").reshape(16), "proof": rd("proof.bin", np.uint8)}
    if m.get("has_params_bytes"):
        out["params_bytes"] = rd("params.bin", np.uint8)
    return out


def main():
    dump_dir, name = sys.argv[1], sys.argv[2]
    out = load(dump_dir)
    path = os.path.join(ROOT, "tests", "golden", f"rust_{name}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
