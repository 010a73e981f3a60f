#!/usr/bin/env python
"""Time K3 (whole pool) for one lattice and several tile sizes.   python tools/run_k3.py 3x3 0 10 11 12"""
import os, sys, statistics
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "quantum-simulation-of-fermi-hubbard-model_b200")]
import numpy as np
from fhsim.backend import Context, DevicePool, State
from fhsim.circuit import Circuit
from fhsim.symbolic import jordan_wigner
from fhsim.tables import GeneratorPlan
from operators.pool import hubbard_interaction_pool_simplified

lat = sys.argv[1]
nx, ny = map(int, lat.split("x"))
n = 2 * nx * ny
cap = 400 if n > 20 else 10 ** 6
ctx = Context(0)
plans = [GeneratorPlan(jordan_wigner(g), n) for g in hubbard_interaction_pool_simplified(nx, ny)[:cap]]
psi, lam = State(ctx, n), State(ctx, n)
c = Circuit(n, 0)
for q in range(n):
    c.ry(0.3 + 0.1 * q, q)
prog = c.compile(ctx)
psi.set_basis(0); prog.run(psi, [])
lam.set_basis(5); prog.run(lam, [])
ref = None
for tb in sys.argv[2:] or ["0", "11"]:
    os.environ["FHSIM_POOL_TILE_BITS"] = tb
    dpool = DevicePool(ctx, plans, n)
    g = dpool.gradients(psi, lam)
    if ref is None:
        ref = g
    ts = []
    for _ in range(5):
        ctx.timer_start()
        for _r in range(4):
            dpool.enqueue(psi, lam)
        ts.append(ctx.timer_stop() / 4)
    t = min(ts) * 1e-3
    print(f"{lat} n={n} pool={len(plans)} tile_bits={tb}: {t*1e6:9.1f} us  {4.0*(1<<n)*len(plans)/t/1e9:9.1f} GB/s effective  "
          f"max|dg vs first|={np.abs(g-ref).max():.2e}")
    dpool.close()
