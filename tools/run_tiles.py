#!/usr/bin/env python
"""Time the fused tile launches of the separable W network (+ a Coulomb diagonal layer) for one lattice.
   python tools/run_tiles.py 3x4 [tile_bits ...]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "quantum-simulation-of-fermi-hubbard-model_b200")]
import numpy as np
from fhsim.backend import Context, State
from fhsim.circuit import Circuit
from fhsim.symbolic import fermi_hubbard, jordan_wigner
from fhsim.tables import GeneratorPlan
from operators.tools import get_interacting_term

lat = sys.argv[1] if len(sys.argv) > 1 else "3x4"
tbs = [tuple(int(x) for x in v.split(":")) for v in sys.argv[2:]] or [(None,)]
nx, ny = map(int, lat.split("x"))
n = 2 * nx * ny
ctx = Context(0)
st = State(ctx, n)
for spec in tbs:
    tb = spec[0]
    lb = spec[1] if len(spec) > 1 else None
    c = Circuit(n, 0)
    c.basis_change_separable(nx, ny)
    prog = c.compile(ctx, tile_bits=tb, low_bits=lb)
    st.set_basis(3)
    prog.run(st, [])
    reps = 20 if n <= 24 else 3
    prog.time_items(st, 0, prog.n_items, False, 2)
    per = [1e3 * prog.time_items(st, i, 1, False, reps) for i in range(prog.n_items)]
    tot = 1e3 * prog.time_items(st, 0, prog.n_items, False, reps)
    ideal = 32.0 * (1 << n) / 6552.6e9 * 1e6
    print(f"{lat} n={n} tile_bits={prog.tile_bits} low_bits={prog.low_bits} W: {prog.n_items} launches, total {tot:.1f} us, per launch "
          f"{[round(p, 1) for p in per]} us; one streaming pass = {ideal:.1f} us; norm2={st.norm2():.12f}")
    prog.close()
    # Coulomb layer as a single diagonal op (k_diag) and fused in a tile together with ry's
    c = Circuit(n, 0)
    plan = GeneratorPlan(jordan_wigner(get_interacting_term(fermi_hubbard(nx, ny, 1.0, 4.0))), n)
    c.generator(plan, angle=0.37)
    prog = c.compile(ctx, fuse=False)
    prog.run(st, [])
    t = 1e3 * prog.time_items(st, 0, prog.n_items, False, reps)
    print(f"    coulomb diag layer ({len(plan.pieces[0].z)} terms): {t:.1f} us = {100 * ideal / t:.1f} % of roofline")
    prog.close()
