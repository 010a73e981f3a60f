#!/usr/bin/env python
"""Run K2 (out = H psi, <psi|H|psi>) a few times on one lattice -- a target for ncu / quick timing.
   python tools/run_k2.py 3x4 [reps]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "quantum-simulation-of-fermi-hubbard-model_b200")]
import numpy as np
from fhsim.backend import Context, DeviceTable, State
from fhsim.circuit import Circuit
from fhsim.symbolic import fermi_hubbard
from fhsim.tables import PauliTable

lat = sys.argv[1] if len(sys.argv) > 1 else "3x4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
nx, ny = map(int, lat.split("x"))
n = 2 * nx * ny
ctx = Context(0)
tab = DeviceTable(ctx, PauliTable.from_operator(fermi_hubbard(nx, ny, 1.0, 4.0), n))
psi, out = State(ctx, n), State(ctx, n)
c = Circuit(n, 0)
for q in range(n):
    c.ry(0.3 + 0.1 * q, q)
prog = c.compile(ctx)
psi.set_basis(0)
prog.run(psi, [])
tab.apply(psi, out)
ts = []
for _ in range(reps):
    ctx.timer_start()
    e = tab.apply(psi, out)
    ts.append(ctx.timer_stop())
print(f"{lat} n={n} K2 min {min(ts)*1e3:.1f} us  ({32.0*(1<<n)/min(ts)/1e6:.1f} GB/s effective)  E={e.real:.9f} info={tab.info()}")
