"""GPU: ONE create_proof sharded over G ranks (BASELINE config 5) produces exactly the
bytes of the single-GPU proof — and therefore of the oracle, which tests/test_gpu_prover.py pins the
single-GPU prover to.  The ranks are the threads of a b200zk_group; on the single-GPU test tier they
all share device 0 (same SPMD code path, same exchanges, peer copies degenerate to device-to-device
copies), on a multi-GPU box B200ZK_TEST_DEVICES=0,1,... spreads them over real devices."""
import importlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _synth(zk):
    return importlib.import_module(zk.__name__ + ".circuits_synth")


def _devices(world):
    env = os.environ.get("B200ZK_TEST_DEVICES")
    devs = [int(x) for x in env.split(",")] if env else [0]
    return [devs[r % len(devs)] for r in range(world)]


def _single(zk, backend, orc, job, s):
    params = zk.ParamsKZG.setup(backend, job.k, s)
    pk = zk.ProvingKey(params, job.cs, job.k, job.fixed, job.map_col, job.map_row)
    wide = orc.XorShiftWide().draw(pk.rng_draws)
    from oracle import prover as OP
    inst = [orc.ints_to_mont([v % OP.R for v in c]) if len(c) else np.zeros((0, 4), dtype=np.uint64) for c in job.instances]
    tr_repr = orc.ints_to_mont([job.transcript_repr])[0]
    proof = pk.create_proof(job.advice, inst, wide, tr_repr)
    pk.close(); params.close()
    return proof, wide, inst, tr_repr


def _sharded(zk, job, s, world, wide, inst, tr_repr, dev_inputs=False):
    group = zk.Group(_devices(world))
    try:
        params = [zk.ParamsKZG.setup(b, job.k, s) for b in group.backends]
        pks = [zk.ProvingKey(p, job.cs, job.k, job.fixed, job.map_col, job.map_row) for p in params]
        if dev_inputs:
            adv = np.concatenate([np.ascontiguousarray(a).reshape(-1, 4) for a in job.advice])
            d_adv = [b.to_device(adv) for b in group.backends]
            d_wide = [b.to_device(wide) for b in group.backends]
            proof = group.create_proof_dev(pks, d_adv, inst, d_wide, tr_repr)
            for d in d_adv + d_wide:
                d.free()
        else:
            proof = group.create_proof(pks, job.advice, inst, wide, tr_repr)
        for pk in pks:
            pk.close()
        for p in params:
            p.close()
        return proof
    finally:
        group.close()


@pytest.mark.parametrize("name,k,worlds", [("small", 6, (2, 3)), ("mst_shaped", 9, (2, 4, 8)), ("v3_shaped", 8, (2, 5)),
                                           ("generic_shapes", 8, (3, 8))])
def test_sharded_proof_equals_single_gpu(zk, backend, orc, name, k, worlds):
    job = getattr(_synth(zk), name)(k)
    s = orc.random_fr(1, 4321)[0]
    want, wide, inst, tr_repr = _single(zk, backend, orc, job, s)
    for world in worlds:
        got = _sharded(zk, job, s, world, wide, inst, tr_repr)
        assert got == want, f"world {world}: sharded proof differs from the single-GPU proof"
    assert _sharded(zk, job, s, worlds[0], wide, inst, tr_repr, dev_inputs=True) == want


def test_sharded_real_merkle_sum_tree_k11(zk, backend, orc):
    """The reference's MerkleSumTreeCircuit (8 lookups, 4 permutation sets, 5 quotient cosets, lookup terms on 4
    of them) at k = 11 over 2, 4 and 8 ranks: byte-identical to the single-GPU proof, which
    test_gpu_prover.py::test_real_merkle_sum_tree_k11 pins to the oracle."""
    chips = importlib.import_module(zk.__name__ + ".chips")
    job = chips.merkle_sum_tree_job(11, levels=9, seed=5)
    s = orc.random_fr(1, 4321)[0]
    want, wide, inst, tr_repr = _single(zk, backend, orc, job, s)
    for world in (2, 4, 8):
        assert _sharded(zk, job, s, world, wide, inst, tr_repr) == want, f"world {world}"


def test_sharded_lookup_failure_is_collective(zk, orc):
    """Error::ConstraintSystemFailure detected by the rank that owns the lookup reaches every rank (no hang)."""
    synth = _synth(zk)
    job = synth.small(6)
    byte_col = 5 + 3 + 2
    bad = np.array(job.advice[byte_col])
    bad[3] = orc.ints_to_mont([100000])[0]
    job.advice[byte_col] = bad
    s = orc.random_fr(1, 4321)[0]
    group = zk.Group(_devices(3))
    try:
        params = [zk.ParamsKZG.setup(b, job.k, s) for b in group.backends]
        pks = [zk.ProvingKey(p, job.cs, job.k, job.fixed, job.map_col, job.map_row) for p in params]
        wide = orc.XorShiftWide().draw(pks[0].rng_draws)
        from oracle import prover as OP
        inst = [orc.ints_to_mont([v % OP.R for v in c]) for c in job.instances]
        with pytest.raises(zk.B200zkError, match="-5"):
            group.create_proof(pks, job.advice, inst, wide, orc.ints_to_mont([job.transcript_repr])[0])
    finally:
        group.close()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_group_best_multiexp_and_best_fft(zk, backend, orc, world):
    """b200zk_group_msm / b200zk_group_fft: one best_multiexp by point range and one best_fft four-step (exchange fused into
    the column-step kernel) over the ranks of a group, host buffers in and out, against the oracle."""
    from oracle import pyref
    n = (1 << 14) + 37
    g, _ = orc.params_setup(15, orc.random_fr(1, 1234)[0], with_lagrange=False)
    coeffs = orc.random_fr(n, 77)
    want = orc.g1_batch_normalize(orc.best_multiexp(coeffs, g[:n]))[0]
    group = zk.Group(_devices(world))
    try:
        assert np.array_equal(group.best_multiexp(coeffs, g[:n])[:8], want)
        for k in (16, 17):
            a = orc.random_fr(1 << k, 90 + k)
            w = orc.ints_to_mont([pyref.omega_for_k(k)])[0]
            assert np.array_equal(group.best_fft(a, w, k), orc.best_fft(a, w, k)), f"k={k}"
        a = orc.random_fr(1 << 10, 5)                                      # below the four-step threshold: rank 0 alone
        w = orc.ints_to_mont([pyref.omega_for_k(10)])[0]
        assert np.array_equal(group.best_fft(a, w, 10), orc.best_fft(a, w, 10))
    finally:
        group.close()
