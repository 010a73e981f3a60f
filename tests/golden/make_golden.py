#!/usr/bin/env python
"""Generate tests/golden/*.json by EXECUTING THE REFERENCE'S OWN PYTHON (from /root/reference).

The reference's third-party dependencies (openfermion, pennylane, qiskit, matplotlib) are not
installable in this container, so the reference modules are imported with

  * ``openfermion``  -> a shim module re-exporting this repo's ``fhsim.symbolic`` names
                       (FermionOperator, normal_ordered, hermitian_conjugated, up_index, ...)
  * ``pennylane``, ``qiskit``, ``matplotlib`` -> empty stub modules (nothing on the pinned paths calls them)

and then the *unmodified* reference functions are called:

  operators/pool.py:220-255     hubbard_interaction_pool_simplified        (pool order = operator ids)
  operators/fourier.py:13-37    fourier_transform_matrix
  operators/tools.py:3-23       get_quadratic_term / get_interacting_term
  linalg/exact_diagonalization.py:11-24   jw_number_spin_indices
  models/utils.py:304-333       get_hva_commuting_hopping_terms            (HVA layer colouring)
  models/adapt_vqe.py:104-122   get_non_interacting_ground_state_index     (source-extracted: module import needs torch+pennylane)

What this pins: every ordering / indexing decision that lives in the reference's own files.
What it does NOT pin: OpenFermion / PennyLane arithmetic itself (restated; see oracle/ headers).

This script only runs where /root/reference exists (the build container); the JSON it writes is
committed and is what the tests read.

    python tests/golden/make_golden.py
"""
import importlib
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, "quantum-simulation-of-fermi-hubbard-model_b200")
REF = os.environ.get("FH_REFERENCE", "/root/reference")


def _install_shims():
    sys.path.insert(0, PKG)
    import fhsim.symbolic as sym
    shim = types.ModuleType("openfermion")
    for name in sym.__all__:
        setattr(shim, name, getattr(sym, name))
    shim.get_sparse_operator = None            # imported by linalg/exact_diagonalization.py:2-6, unused on the pinned path
    sys.modules["openfermion"] = shim
    for name in ("pennylane", "qiskit", "qiskit.quantum_info", "qiskit.circuit", "matplotlib", "matplotlib.pyplot"):
        m = types.ModuleType(name)
        sys.modules[name] = m
    sys.modules["qiskit"].QuantumCircuit = object
    sys.modules["qiskit.quantum_info"].SparsePauliOp = object
    sys.modules["qiskit.circuit"].Parameter = object
    # our own drop-in packages are called operators/ linalg/ models/ too: make sure the REFERENCE ones win
    sys.path.remove(PKG)
    sys.path.insert(0, REF)
    sys.path.append(PKG)                       # fhsim itself stays importable
    for name in ("operators", "linalg", "models"):
        sys.modules.pop(name, None)
    return sym


def _fermion_terms(op):
    """FermionOperator -> [[ [[index, dagger], ...], re, im ], ...] in .terms order."""
    out = []
    for term, c in op.terms.items():
        c = complex(c)
        out.append([[list(map(int, f)) for f in term], c.real, c.imag])
    return out


def _qubit_terms(op):
    out = []
    for term, c in op.terms.items():
        c = complex(c)
        out.append([[[int(q), p] for q, p in term], c.real, c.imag])
    return out


def main():
    sym = _install_shims()
    ref_pool = importlib.import_module("operators.pool")
    ref_fourier = importlib.import_module("operators.fourier")
    ref_tools = importlib.import_module("operators.tools")
    ref_ed = importlib.import_module("linalg.exact_diagonalization")
    ref_utils = importlib.import_module("models.utils")
    assert ref_pool.__file__.startswith(REF), ref_pool.__file__

    # get_non_interacting_ground_state_index: exec its source lines only (models/adapt_vqe.py imports pennylane/torch at module level)
    import numpy as np
    src = open(os.path.join(REF, "models", "adapt_vqe.py")).read().splitlines()
    start = next(i for i, l in enumerate(src) if l.startswith("def get_non_interacting_ground_state_index"))
    end = next(i for i in range(start + 1, len(src)) if src[i] and not src[i].startswith((" ", "\t", "#")))
    ns = {"np": np, "FermionOperator": sym.FermionOperator}
    exec("\n".join(src[start:end]), ns)
    get_index = ns["get_non_interacting_ground_state_index"]

    golden = {"_generator": "tests/golden/make_golden.py", "_reference": "chuntse0514/Quantum-Simulation-of-Fermi-Hubbard-model",
              "lattices": {}}
    for (nx, ny, n_el, t, u) in [(2, 2, 4, 1.0, 4.0), (2, 3, 6, 1.0, 4.0), (3, 3, 9, 1.0, 6.0)]:
        n = 2 * nx * ny
        pool = ref_pool.hubbard_interaction_pool_simplified(nx, ny)
        ft = ref_fourier.fourier_transform_matrix(nx, ny)
        ham = sym.fermi_hubbard(nx, ny, t, u)
        quad = ref_tools.get_quadratic_term(ham)
        inter = ref_tools.get_interacting_term(ham)
        ft_quad = ref_fourier.fourier_transform(quad, nx, ny)
        n_up = (n_el + 1) // 2
        n_dn = n_el - n_up
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):          # the reference prints the orbital energies
            occ_up, occ_dn = get_index(ft_quad, n, n_up, n_dn)
        entry = {
            "n_qubits": n,
            "pool": [_fermion_terms(op) for op in pool],
            "ft_matrix_re": np.real(ft).tolist(), "ft_matrix_im": np.imag(ft).tolist(),
            "quadratic_terms": _fermion_terms(quad), "interacting_terms": _fermion_terms(inter),
            "ft_quadratic_terms": _fermion_terms(ft_quad),
            "sector_indices": [int(v) for v in ref_ed.jw_number_spin_indices(n_el, n_up, n_dn, n)],
            "sector": [n_el, n_up, n_dn],
        }
        entry["occupied_up"] = [int(v) for v in occ_up]
        entry["occupied_dn"] = [int(v) for v in occ_dn]
        hset, vset = ref_utils.get_hva_commuting_hopping_terms(nx, ny, True)
        entry["hva_horizontal"] = [_fermion_terms(g) for g in hset]
        entry["hva_vertical"] = [_fermion_terms(g) for g in vset]
        golden["lattices"][f"{nx}x{ny}"] = entry
        print(f"{nx}x{ny}: pool {len(pool)}, sector {len(entry['sector_indices'])}, occ {occ_up} {occ_dn}, "
              f"hva sets {len(hset)}h/{len(vset)}v")
    with open(os.path.join(HERE, "reference_host_tables.json"), "w") as f:
        json.dump(golden, f, separators=(",", ":"))
    print("wrote", os.path.join(HERE, "reference_host_tables.json"))

    # pools no driver uses (operators/pool.py:48-131, 257-340): signatures kept, outputs pinned all the same
    unused = {"_generator": "tests/golden/make_golden.py",
              "spin_complemented_pool": {}, "hubbard_interation_pool_modified": {}}
    for (n_el, n_orb, gen) in [(4, 4, True), (4, 4, False), (2, 3, True)]:
        ops = ref_pool.spin_complemented_pool(n_el, n_orb, gen)
        unused["spin_complemented_pool"][f"{n_el},{n_orb},{int(gen)}"] = [_fermion_terms(op) for op in ops]
    for (nx, ny) in [(2, 2), (2, 3)]:
        ch = ref_pool.hubbard_interation_pool_modified(nx, ny)
        unused["hubbard_interation_pool_modified"][f"{nx}x{ny}"] = {k: _fermion_terms(v) for k, v in ch.items()}
    with open(os.path.join(HERE, "reference_unused_pools.json"), "w") as f:
        json.dump(unused, f, separators=(",", ":"))
    print("wrote", os.path.join(HERE, "reference_unused_pools.json"))


if __name__ == "__main__":
    main()
