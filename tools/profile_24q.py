#!/usr/bin/env python
"""Launch every kernel class twice on the 3x4 lattice (24 qubits, 256 MiB state > L2): the target of
`ncu --set full` for the HBM-regime dram-traffic numbers.   python tools/profile_24q.py [lattice]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "quantum-simulation-of-fermi-hubbard-model_b200")]
import numpy as np
from fhsim.backend import Context, DevicePool, DeviceTable, State
from fhsim.circuit import Circuit
from fhsim.symbolic import fermi_hubbard, jordan_wigner
from fhsim.tables import GeneratorPlan, PauliTable
from operators.pool import hubbard_interaction_pool_simplified
from operators.tools import get_interacting_term

lat = sys.argv[1] if len(sys.argv) > 1 else "3x4"
nx, ny = map(int, lat.split("x"))
n = 2 * nx * ny
ctx = Context(0)
h = fermi_hubbard(nx, ny, 1.0, 4.0)
plans = [GeneratorPlan(jordan_wigner(g), n) for g in hubbard_interaction_pool_simplified(nx, ny)[:64]]
coulomb = GeneratorPlan(jordan_wigner(get_interacting_term(h)), n)
psi, lam = State(ctx, n), State(ctx, n)
c = Circuit(n, 0)
for q in range(n):
    c.ry(0.3 + 0.1 * q, q)
prog = c.compile(ctx)
psi.set_basis(0); prog.run(psi, []); prog.close()
lam.set_basis(3)

def once(build, fuse=False, reps=2):
    cc = Circuit(n, 0)
    build(cc)
    p = cc.compile(ctx, fuse=fuse)
    for _ in range(reps):
        p.run(psi, [])
    p.close()

x = (1 << (n - 1)) | (1 << (n // 2)) | 0b110
z = (1 << (n - 2)) | 0b011
once(lambda cc: cc.pauli_rotation(x, z, 0.5, angle=0.7))              # k_pair: dense Pauli rotation
once(lambda cc: cc.single_excitation(0.4, n // 2, n // 2 + 1))        # k_pair: Givens
once(lambda cc: cc.generator(plans[len(plans) // 2], angle=0.3))      # k_pair: fermionic double excitation
once(lambda cc: cc.generator(coulomb, angle=0.2))                     # k_diag_build + k_diag_tab
once(lambda cc: cc.basis_change_separable(nx, ny), fuse=True, reps=1) # k_tile
tab = DeviceTable(ctx, PauliTable.from_operator(h, n))
for _ in range(2):
    tab.apply(psi, lam)                                               # k_table_pass x 2 (K2 in tile passes)
ns = nx * ny
n_up, n_dn = (ns + 1) // 2, ns - (ns + 1) // 2
os.environ["FHSIM_K2_GATHER"] = "1"
tab.apply(psi, lam)                                                   # k_apply_table4 (gather kernel)
del os.environ["FHSIM_K2_GATHER"]
for _ in range(2):
    tab.apply_sector(psi, lam, n_up, n_dn)                            # k_sector_compress1 + k_sector_happly + k_sector_scatter1
for tb in ("0", "12"):
    os.environ["FHSIM_POOL_TILE_BITS"] = tb
    pool = DevicePool(ctx, plans, n)
    for _ in range(2):
        pool.gradients(psi, lam)                                      # k_pool32 / k_pool_tile
    if tb == "0":
        for _ in range(2):
            pool.gradients_sector(psi, lam, n_up, n_dn)               # k_sector_compress2 + k_sector_pool
    pool.close()
print("done", lat, n)
