import sys, ctypes, importlib, numpy as np
sys.path.insert(0, '/root/repo')
from __graft_entry__ import load_package
zk = load_package()
from oracle import binding as orc, prover as OP
synth = importlib.import_module(zk.__name__ + ".circuits_synth")
job = synth.small(5)
be = zk.Backend(0)
s = orc.random_fr(1, 4321)[0]
params = zk.ParamsKZG.setup(be, job.k, s)
g, gl = params.read()
pk = zk.ProvingKey(params, job.cs, job.k, job.fixed, job.map_col, job.map_row)
wide = orc.XorShiftWide().draw(pk.rng_draws)
inst = [orc.ints_to_mont(c) for c in job.instances]
got = pk.create_proof(job.advice, inst, wide, orc.ints_to_mont([job.transcript_repr])[0])
def dbg(name, count):
    out = np.zeros((count, 4), dtype=np.uint64); c = ctypes.c_size_t()
    be._check(zk.lib().b200zk_pk_debug_buffer(pk._h, name.encode(), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(count), ctypes.byref(c)))
    return out
n = 32
lk = dbg("lookups", 7 * 2 * n).reshape(2, 7, n, 4)
pk_cpu = OP.keygen_pk(job.cs, job.k, job.fixed, job.map_col, job.map_row)
# re-run oracle with instrumentation: monkeypatch commit to capture lookups
want, trace = OP.create_proof(g, gl, pk_cpu, job.advice, job.instances, wide, job.transcript_repr)
import oracle.prover as op
# recompute the oracle lookup data for lookup 0
theta, beta, gamma = trace["theta"], trace["beta"], trace["gamma"]
names = ["cin", "ctab", "pin", "ptab", "pin_poly", "ptab_poly", "z_poly"]
for i, nm in enumerate(names):
    print(nm, orc.mont_to_ints(lk[0, i][:3]))
dom = pk_cpu.dom
# oracle-side expected z_poly for lookup 0 from GPU's own cin/ctab/pin/ptab
den = op.vmul(op.vadds(lk[0,2], beta), op.vadds(lk[0,3], gamma))
den = orc.batch_invert(den)
prod = op.vmul(op.vmul(den, op.vadds(lk[0,0], beta)), op.vadds(lk[0,1], gamma))
pi = orc.mont_to_ints(prod)
zi = [1]
for row in range(1, n - 5): zi.append(zi[-1] * pi[row-1] % op.R)
print("z[usable] == 1 ?", zi[26] == 1)
zg = orc.mont_to_ints(dom.coeff_to_extended(lk[0,6])[::8]) if False else None
# convert GPU z_poly back to lagrange via NTT
zl = orc.best_fft(lk[0,6], dom.omega, 5)
zl = orc.mont_to_ints(zl)
print("gpu z lagrange first 5:", zl[:5])
print("expected          :", zi[:5])
print("match rows:", [a == b for a, b in zip(zl[:27], zi)])

oz = orc.mont_to_ints(trace["lookup_z"][0])
print("tail gpu   :", [hex(v)[:12] for v in zl[27:]])
print("tail oracle:", [hex(v)[:12] for v in oz[27:]])
print("rows equal :", [a == b for a, b in zip(zl, oz)])
rnd = orc.mont_to_ints(dbg("rnd", pk.rng_draws))
allr = orc.mont_to_ints(orc.from_u512(wide))
print("rnd equal:", rnd == allr, len(rnd))
for v in zl[27:]:
    print("gpu tail idx in stream:", allr.index(v) if v in allr else None)
for v in oz[27:]:
    print("oracle tail idx in stream:", allr.index(v) if v in allr else None)
