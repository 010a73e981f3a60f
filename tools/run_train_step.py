#!/usr/bin/env python
"""Time one optimiser-step evaluation (energy + d<H>/dtheta by the adjoint sweep + <Sz>, <S^2>) of the cfg3 ansatz.
   python tools/run_train_step.py            (FHSIM_UNFUSED_ADJOINT=1 for the one-launch-per-op sweep)"""
import os, sys, statistics
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "quantum-simulation-of-fermi-hubbard-model_b200")]
import numpy as np
import bench
from fhsim.backend import Context, DeviceTable
from fhsim.symbolic import jordan_wigner
from fhsim.tables import PauliTable
from models.common import get_spin_operators

ctx = Context(0)
wl = bench.build_gpu_workload(ctx)
prog, dtab, basis, thetas = wl["prog"], wl["dtab"], wl["basis"], wl["thetas"]
n = bench.N_QUBITS
s2 = DeviceTable(ctx, PauliTable.from_operator(jordan_wigner(get_spin_operators(9, 'S^2')), n))
sz = DeviceTable(ctx, PauliTable.from_operator(jordan_wigner(get_spin_operators(9, 'Sz')), n))
for _ in range(5):
    res = prog.evaluate(basis, thetas, [dtab, sz, s2], grads=True)
ms = []
for _ in range(200):
    res = prog.evaluate(basis, thetas, [dtab, sz, s2], grads=True)
    ms.append(prog.last_stats()[0])
print("train step: %.1f us device (median), %d launches; E=%.12f Sz=%.6f S2=%.6f |grad|=%.12f g0=%.12f" % (
    1e3 * statistics.median(ms), prog.last_stats()[1], res["expvals"][0], res["expvals"][1], res["expvals"][2],
    np.linalg.norm(res["grads"]), res["grads"][0]))
np.save(os.path.join(R, "gpurun_out", "train_grads_%s.npy" % ("unfused" if os.environ.get("FHSIM_UNFUSED_ADJOINT") else "fused")), res["grads"])
