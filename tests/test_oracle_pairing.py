"""CPU: the oracle's bn256 pairing (test infrastructure) and the pairing form of the verifier —
the reference's `verify_proof(..).is_ok()` (/root/reference/src/circuits/utils.rs:52-63) without
knowledge of the SRS secret."""
import importlib

import pytest

from oracle import pairing as PR
from oracle import prover as OP
from oracle import pyref as P


def test_pairing_properties():
    assert PR.g2_on_curve(PR.G2_GEN)
    assert PR.g2_mul(PR.G2_GEN, PR.R - 1) == PR.g2_neg(PR.G2_GEN)            # r * G2 = identity
    x = [3, 5, 7, 0, 0, 0, 1, 0, 0, 0, 0, 2]
    assert PR.f12_mul(x, PR.f12_inv(x)) == PR.F12_ONE
    e1 = PR.pairing(PR.G2_GEN, P.G1_GEN)
    assert e1 != PR.F12_ONE and PR.f12_pow(e1, PR.R) == PR.F12_ONE           # non-degenerate, in mu_r
    a, b = 123456789, 987654321
    assert PR.pairing(PR.g2_mul(PR.G2_GEN, b), P.g1_mul(P.G1_GEN, a)) == PR.f12_pow(e1, a * b % PR.R)
    assert PR.pairing_product_is_one([(P.g1_mul(P.G1_GEN, a), PR.g2_mul(PR.G2_GEN, b)),
                                      (P.g1_mul(P.G1_GEN, (-a * b) % PR.R), PR.G2_GEN)])
    assert not PR.pairing_product_is_one([(P.g1_mul(P.G1_GEN, a), PR.g2_mul(PR.G2_GEN, b)),
                                          (P.g1_mul(P.G1_GEN, (-a * b + 1) % PR.R), PR.G2_GEN)])


def test_full_prover_verifies_by_pairing(zk, orc):
    """test_full_prover end to end: setup, keygen, create_proof (oracle), verify with e(., [s]_2)."""
    fe = importlib.import_module(zk.__name__ + ".frontend")
    chips = importlib.import_module(zk.__name__ + ".chips")
    leaf, elements, indices = (10, 100), [(1, 10), (5, 50), (6, 60), (9, 90), (9, 90)], [0] * 5
    root = chips.compute_merkle_sum_root(leaf, elements, indices)
    circuit = chips.MerkleSumTreeCircuit(leaf[0], leaf[1], [e[0] for e in elements], [e[1] for e in elements], indices, 500)
    job = fe.synthesize_job(circuit, 9, [[leaf[0], leaf[1], root[0], 500]])
    s = orc.random_fr(1, 2024)[0]
    g, gl = orc.params_setup(job.k, s)
    s_g2 = PR.g2_mul(PR.G2_GEN, orc.mont_to_ints(s)[0])                       # ParamsKZG.s_g2
    pk = OP.keygen_pk(job.cs, job.k, job.fixed, job.map_col, job.map_row)
    wide = orc.XorShiftWide().draw(OP.rng_draws_needed(job.cs, job.k))
    proof, _ = OP.create_proof(g, gl, pk, job.advice, job.instances, wide, job.transcript_repr)
    assert OP.verify_full(None, g, pk, job.instances, proof, job.transcript_repr, s_g2=s_g2)
    bad = bytearray(proof)
    bad[-40] ^= 1                                                               # inside the last commitment / evaluation block
    try:
        ok = OP.verify_full(None, g, pk, job.instances, bytes(bad), job.transcript_repr, s_g2=s_g2)
    except (AssertionError, ValueError):
        ok = False
    assert not ok
    assert not OP.verify_full(None, g, pk, [[leaf[0], leaf[1], root[0], 499]], proof, job.transcript_repr, s_g2=s_g2)
