#!/usr/bin/env python
"""cfg 4: Lanczos replacement of linalg/exact_diagonalization.py on the 3x3 lattice (18 qubits), timed next to the
reference's route (sector matrix + scipy.sparse.linalg.eigsh, oracle/ed.py) on the host.
   python tests/perf_lanczos.py [--json out.json]"""
import json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "quantum-simulation-of-fermi-hubbard-model_b200")]
import numpy as np
from fhsim.backend import Context, DeviceTable, lanczos, lanczos_sector
from fhsim.symbolic import fermi_hubbard
from fhsim.tables import PauliTable
from oracle import ed, pauli

nx, ny, u, n = 3, 3, 6.0, 18
ctx = Context(0)
tab = DeviceTable(ctx, PauliTable.from_operator(fermi_hubbard(nx, ny, 1.0, u), n))
out = {"lattice": "3x3", "n_qubits": n, "U": u}

def timed(label, **kw):
    """Sector requests run on sector-compressed vectors (fh_lanczos_sector): the algorithmic bytes per iteration are
    96 B x the SECTOR dimension (Hv 32 + recurrence/dots 64 bytes per amplitude); full-space requests keep
    96 * 2^n."""
    sector = kw.get("n_up", -1) >= 0
    if sector:
        args = dict(k=kw["k"], tol=kw["tol"], max_iter=kw["max_iter"], seed=kw["seed"])
        lanczos_sector(tab, kw["n_up"], kw["n_dn"], **args)                # warm-up (allocations)
        ctx.sync()
        t0 = time.perf_counter()
        evals, _, _, info = lanczos_sector(tab, kw["n_up"], kw["n_dn"], **args)
        ctx.sync()
        dt = time.perf_counter() - t0
        iters, dim, loop = info["matvecs"], info["sector_dim"], info["loop_seconds"]
        out[label] = {"seconds": dt, "iterations": info["iterations"], "matvecs": iters, "evals": [float(e) for e in evals],
                      "sector_dim": dim, "host_syncs": info["host_syncs"], "loop_seconds": loop,
                      "matvecs_per_s": iters / dt, "matvecs_per_s_in_loop": iters / loop,
                      "effective_GBps": 96.0 * dim * iters / loop / 1e9}
    else:
        lanczos(tab, **kw)
        ctx.sync()
        t0 = time.perf_counter()
        evals, vecs, iters = lanczos(tab, **kw)
        ctx.sync()
        dt = time.perf_counter() - t0
        for v in vecs:
            v.close()
        out[label] = {"seconds": dt, "iterations": iters, "evals": [float(e) for e in evals],
                      "matvecs_per_s": iters / dt, "effective_GBps": 96.0 * (1 << n) * iters / dt / 1e9}
    print(label, out[label])

timed("gpu_sector_5_4_k1", k=1, n_up=5, n_dn=4, tol=1e-12, max_iter=2000, seed=7)
timed("gpu_sector_5_4_k4_degenerate", k=4, n_up=5, n_dn=4, tol=1e-11, max_iter=4000, seed=7)
timed("gpu_full_space_k1", k=1, tol=1e-11, max_iter=4000, seed=7)

o_h = pauli.compress(pauli.jw_table(pauli.hubbard_fermion_terms(nx, ny, 1.0, u), n))
t0 = time.perf_counter()
vals, _, idx = ed.ground_state(o_h, n, 9, 5, 4, k=10)
out["cpu_reference_route_sector_5_4_k10"] = {"seconds": time.perf_counter() - t0, "evals": [float(v) for v in vals[:5]],
                                              "what": "sector matrix build + scipy eigsh(k=10, which='SA') as "
                                                      "linalg/exact_diagonalization.py:181-229 (oracle/ed.py)",
                                              "sector_dim": int(len(idx))}
print("cpu", out["cpu_reference_route_sector_5_4_k10"])
assert abs(out["gpu_sector_5_4_k1"]["evals"][0] - vals[0]) < 1e-8
assert max(abs(a - vals[0]) for a in out["gpu_sector_5_4_k4_degenerate"]["evals"]) < 1e-7
if len(sys.argv) > 2 and sys.argv[1] == "--json":
    json.dump(out, open(sys.argv[2], "w"), indent=1)
