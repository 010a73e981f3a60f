"""Pin the oracle to analytic / literature known answers.

The reference ships no tests or golden vectors, so these are the anchors that stand in for them.
"""
import numpy as np
import pytest

from oracle import ed, pauli, statevector as sv

LATTICES = {
    (2, 2): dict(u=4.0, up=2, dn=2, e_hf=0.0, e0=-2.1027484835, pool=24, nz=8, g=2.0, terms=29),
    (2, 3): dict(u=4.0, up=3, dn=3, e_hf=-2.0, e0=-3.7898230717, pool=90, nz=17, g=4.0 / 3, terms=55),
    (2, 4): dict(u=2.0, up=4, dn=4, e_hf=-8.0, e0=-8.4783032969, pool=224, nz=40, g=0.5, terms=73),
}


def build(nx, ny, u):
    n = 2 * nx * ny
    h = pauli.compress(pauli.jw_table(pauli.hubbard_fermion_terms(nx, ny, 1.0, u), n))
    pool = [pauli.jw_table(op, n) for op in pauli.pool_fermion_terms(nx, ny)]
    layers, diag = pauli.givens_network(pauli.ft_matrix(nx, ny))
    return n, h, pool, layers, diag


def test_dimer_analytic():
    n = 4
    h = pauli.compress(pauli.jw_table(pauli.hubbard_fermion_terms(2, 1, 1.0, 4.0), n))
    vals, _, _ = ed.ground_state(h, n, 2, 1, 1)
    assert abs(vals[0] - (4 - np.sqrt(16 + 16)) / 2) < 1e-10


@pytest.mark.parametrize("lat", list(LATTICES))
def test_known_answers(lat):
    nx, ny = lat
    ref = LATTICES[lat]
    n, h, pool, layers, diag = build(nx, ny, ref["u"])
    assert len(h) == ref["terms"]
    assert len(pool) == ref["pool"]
    for g in pool:                      # 8 strings, one shared x-mask, coefficients +-1/8
        assert len(g) == 8 and len({k[0] for k in g}) == 1
        assert all(abs(abs(c) - 0.125) < 1e-15 and abs(complex(c).imag) < 1e-15 for c in g.values())
    up, dn, _ = pauli.k_space_occupation(nx, ny, 1.0, ref["up"], ref["dn"])
    psi = sv.basis_state(n, up + dn)
    grads, e_hf, _ = sv.pool_gradients(psi, h, pool, diag, layers, n)
    assert abs(e_hf - ref["e_hf"]) < 1e-12
    nonzero = np.abs(grads) > 1e-9
    assert nonzero.sum() == ref["nz"]
    assert np.allclose(np.abs(grads[nonzero]), ref["g"], atol=1e-12)
    vals, _, _ = ed.ground_state(h, n, ref["up"] + ref["dn"], ref["up"], ref["dn"])
    assert abs(vals[0] - ref["e0"]) < 2e-10


def test_w_contract_one_hot():
    """W|1_p> = e^{i gamma} sum_q Q[p,q] |1_q>  (SURVEY A.2), complex 3x1 case."""
    nx, ny = 3, 1
    n = 6
    q = pauli.ft_matrix(nx, ny)
    layers, diag = pauli.givens_network(q)
    vac = sv.basis_change(sv.basis_state(n, []), diag, layers, n)
    gamma = vac[0]
    for p in range(n):
        out = sv.basis_change(sv.basis_state(n, [p]), diag, layers, n) / gamma
        expect = np.zeros(1 << n, complex)
        for r in range(n):
            expect[1 << (n - 1 - r)] = q[p, r]
        assert np.abs(out - expect).max() < 1e-12


def test_3x1_hf_energy():
    n = 6
    h = pauli.compress(pauli.jw_table(pauli.hubbard_fermion_terms(3, 1, 1.0, 4.0), n))
    layers, diag = pauli.givens_network(pauli.ft_matrix(3, 1))
    up, dn, _ = pauli.k_space_occupation(3, 1, 1.0, 2, 1)
    phi = sv.basis_change(sv.basis_state(n, up + dn), diag, layers, n)
    assert abs(sv.expval(phi, h, n).real - (-1.0 / 3)) < 1e-12
