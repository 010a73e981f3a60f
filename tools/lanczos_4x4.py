#!/usr/bin/env python
"""4x4 Hubbard (32 qubits, U = 4, half filling): ground-state energy by sector-compressed Lanczos on ONE B200.

The reference's route (models/adapt_vqe.py:221-247 -> linalg/exact_diagonalization.py:34-51: 2^32 x 2^32 sparse matrix,
slice to the (8, 8) sector, ARPACK) is out of reach at this size; here the Krylov vectors live on the sector
(165 636 900 amplitudes = 2.65 GB each) and H is applied matrix-free from the packed Pauli table.

    python tools/lanczos_4x4.py [--json out.json] [--tol 1e-9]
Known answer (literature, exact diagonalisation of the 4x4 periodic cluster, U/t = 4, half filling): E0 = -13.6219 t."""
import json
import os
import sys
import time

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "quantum-simulation-of-fermi-hubbard-model_b200")]
from fhsim.symbolic import fermi_hubbard  # noqa: E402
from linalg.exact_diagonalization import get_sparse_operator, jw_get_ground_state_compressed  # noqa: E402

tol = float(sys.argv[sys.argv.index("--tol") + 1]) if "--tol" in sys.argv else 1e-9
t0 = time.perf_counter()
op = get_sparse_operator(fermi_hubbard(4, 4, 1.0, 4.0), 32)
e0, _, info = jw_get_ground_state_compressed(op, 16, 8, 8, tol=tol, max_iter=600, want_vector=False)
wall = time.perf_counter() - t0
out = {"lattice": "4x4", "n_qubits": 32, "U": 4.0, "sector": [8, 8], "E0": e0, "E0_literature": -13.6219,
       "abs_diff_to_literature": abs(e0 + 13.6219), "tol": tol, "wall_seconds": wall, **info}
out["matvecs_per_s"] = info["matvecs"] / info["loop_seconds"]
out["bytes_per_vector"] = 16 * info["sector_dim"]
print(json.dumps(out))
if "--json" in sys.argv:
    json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
assert abs(e0 + 13.6219) < 5e-4, e0
