"""oracle/cpu_closed_form.c (the C + OpenMP closed form bench.py times as "cpu_closed_form") against the numpy oracle and
the test-side interpreter of the op semantics: pair ops, diagonal ops, H|psi>, pool gradients, one whole screening."""
import numpy as np
import pytest

import emulate
from fhsim.circuit import Circuit, Marker
from fhsim.symbolic import fermi_hubbard, givens_decomposition_square, jordan_wigner
from fhsim.tables import GeneratorPlan, PauliTable
from operators.fourier import fourier_transform_matrix
from operators.pool import hubbard_interaction_pool_simplified
from oracle import cpu_closed_form as cf, pauli, statevector as sv


def _rand(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return v / np.linalg.norm(v)


def test_ops_and_table_vs_interpreter():
    n = 9
    rng = np.random.default_rng(3)
    c = Circuit(n, 2)
    for _ in range(25):
        k = int(rng.integers(5))
        a, b = (int(v) for v in rng.choice(n, size=2, replace=False))
        if k == 0:
            c.ry(0.0, a, param=int(rng.integers(2)))
        elif k == 1:
            c.fermionic_single_excitation(float(rng.uniform(-2, 2)), a, b)
        elif k == 2:
            c.rz(float(rng.uniform(-2, 2)), a)
        elif k == 3:
            c.pauli_rotation(int(rng.integers(1, 1 << n)), int(rng.integers(1 << n)), 0.5, param=int(rng.integers(2)))
        else:
            c.cnot(a, b)
    th = rng.uniform(-1, 1, 2)
    psi0 = _rand(n, 1)
    want = emulate.run_circuit(c, psi0.copy(), th)
    for threads in (1, 4):
        cf.lib().cf_set_threads(threads)
        got = cf.run_ops(psi0.copy(), c.ops, th, n)
        assert np.abs(got - want).max() < 1e-13
        back = cf.run_ops(got.copy(), c.ops, th, n, dagger=True)
        assert np.abs(back - psi0).max() < 1e-13
    h = PauliTable.from_operator(fermi_hubbard(3, 1, 1.0, 4.0), 6)
    v = _rand(6, 2)
    assert np.abs(cf.apply_table(v, h, 6) - sv.apply_table(v, h.as_dict(), 6)).max() < 1e-13
    xs, zs = rng.integers(0, 1 << 7, 12), rng.integers(0, 1 << 7, 12)
    t = PauliTable(7, xs, zs, rng.normal(size=12) + 1j * rng.normal(size=12))
    w = _rand(7, 4)
    assert np.abs(cf.apply_table(w, t, 7) - sv.apply_table(w, t.as_dict(), 7)).max() < 1e-13


@pytest.mark.parametrize("nx,ny,u,up,dn", [(2, 2, 4.0, 2, 2), (2, 3, 4.0, 3, 3)])
def test_whole_screening_vs_numpy_oracle(nx, ny, u, up, dn):
    n = 2 * nx * ny
    h_tab = PauliTable.from_operator(fermi_hubbard(nx, ny, 1.0, u), n)
    plans = [GeneratorPlan(jordan_wigner(g), n) for g in hubbard_interaction_pool_simplified(nx, ny)]
    dec, diag = givens_decomposition_square(fourier_transform_matrix(nx, ny))
    o_h = pauli.compress(pauli.jw_table(pauli.hubbard_fermion_terms(nx, ny, 1.0, u), n))
    o_pool = [pauli.jw_table(op, n) for op in pauli.pool_fermion_terms(nx, ny)]
    o_up, o_dn, _ = pauli.k_space_occupation(nx, ny, 1.0, up, dn)
    occ = o_up + o_dn
    picks = [1, 4, 7, 11]
    th = np.random.default_rng(8).uniform(-0.4, 0.4, len(picks))
    circ = Circuit(n, len(picks))
    for j, k in enumerate(picks):
        circ.generator(plans[k], param=j)
    circ.marker("ansatz_end")
    circ.basis_change(diag, list(reversed(dec)))
    cut = next(i for i, op in enumerate(circ.ops) if isinstance(op, Marker))
    basis = sum(1 << (n - 1 - q) for q in occ)
    e, g = cf.screening(n, basis, circ.ops[:cut], circ.ops[cut + 1:], th, h_tab, plans, threads=4)
    psi = sv.adapt_state(n, occ, [o_pool[k] for k in picks], th)
    g_or, e_or, _ = sv.pool_gradients(psi, o_h, o_pool, diag, dec, n)
    assert abs(e - e_or) < 1e-12
    assert np.abs(g - g_or).max() < 1e-12
