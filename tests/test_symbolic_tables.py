"""Host compile parity: the product's symbolic route (tuple-keyed operators) and the oracle's
packed-integer route must give bit-identical Pauli tables, pools and Givens networks."""
import numpy as np
import pytest

from fhsim.symbolic import (FermionOperator, QubitOperator, fermi_hubbard, get_interaction_operator,
                            givens_decomposition_square, jordan_wigner, normal_ordered)
from fhsim.tables import GeneratorPlan, PauliTable
from operators.fourier import fourier_transform, fourier_transform_matrix
from operators.pool import hubbard_interaction_pool_simplified
from operators.tools import get_interacting_term, get_quadratic_term
from oracle import pauli

SIZES = {(2, 2): (29, 13, 9, 24), (2, 3): (55, 19, 19, 90), (2, 4): (73, 25, 25, 224), (3, 3): (100, 28, 37, 324),
         (4, 4): (177, 49, 65, 1920)}


@pytest.mark.parametrize("lat", list(SIZES))
def test_hamiltonian_table_bit_exact(lat):
    nx, ny = lat
    n = 2 * nx * ny
    terms, diag, groups, _ = SIZES[lat]
    tab = PauliTable.from_operator(fermi_hubbard(nx, ny, 1.0, 4.0), n)
    ox, oz, ok, oc = pauli.table_arrays(pauli.compress(pauli.jw_table(pauli.hubbard_fermion_terms(nx, ny, 1.0, 4.0), n)))
    assert len(tab) == terms and tab.n_groups == groups
    assert int((tab.x == 0).sum()) == diag
    assert np.array_equal(tab.x, ox) and np.array_equal(tab.z, oz) and np.array_equal(tab.k, ok)
    assert np.array_equal(tab.coeff, oc)           # bit-exact coefficients, same order


@pytest.mark.parametrize("lat", [(2, 2), (2, 3), (3, 3), (4, 4)])
def test_pool_tables_bit_exact(lat):
    nx, ny = lat
    n = 2 * nx * ny
    pool = hubbard_interaction_pool_simplified(nx, ny)
    opool = pauli.pool_fermion_terms(nx, ny)
    assert len(pool) == len(opool) == SIZES[lat][3]
    step = 1 if len(pool) < 400 else 37
    for k in range(0, len(pool), step):
        assert list(pool[k].terms.items()) == [(t, c) for t, c in opool[k]]
        tab = PauliTable.from_operator(jordan_wigner(pool[k]), n, compress=False)
        ox, oz, ok, oc = pauli.table_arrays(pauli.jw_table(opool[k], n))
        assert np.array_equal(tab.x, ox) and np.array_equal(tab.z, oz) and np.array_equal(tab.coeff, oc)
        assert len(tab) == 8 and len(set(tab.x.tolist())) == 1
        plan = GeneratorPlan(jordan_wigner(pool[k]), n)
        assert plan.exact and len(plan.pieces) == 1
        p = plan.pieces[0]
        assert abs(abs(p.b) - 1.0) < 1e-15 and bin(p.fixmask).count("1") == 4 and p.fixmask == p.x


@pytest.mark.parametrize("lat", [(2, 2), (2, 3), (3, 3), (3, 1)])
def test_givens_network_matches(lat):
    nx, ny = lat
    q = fourier_transform_matrix(nx, ny)
    assert np.array_equal(q, pauli.ft_matrix(nx, ny))
    dec, diag = givens_decomposition_square(q)
    odec, odiag = pauli.givens_network(q)
    assert [tuple(l) for l in dec] == [tuple(l) for l in odec]
    assert np.array_equal(diag, odiag)
    assert np.allclose(np.abs(diag), 1.0, atol=1e-12)


def test_givens_sizes_3x3():
    dec, _ = givens_decomposition_square(fourier_transform_matrix(3, 3))
    rot = [op for layer in dec for op in layer]
    assert len(dec) == 31 and len(rot) == 144
    assert sum(1 for op in rot if abs(abs(op[2]) - np.pi / 2) < 1e-12) == 72


def test_interaction_operator_route_same_operator():
    h = fermi_hubbard(2, 2, 1.0, 4.0)
    a = jordan_wigner(h)
    b = jordan_wigner(get_interaction_operator(h))
    a.compress()
    b.compress()
    assert a == b and len(a.terms) == len(b.terms) == 29


def test_symbolic_algebra_basics():
    assert normal_ordered(FermionOperator('1 1^')) == FermionOperator(()) - FermionOperator('1^ 1')
    assert normal_ordered(FermionOperator('2 2')) == FermionOperator()
    x, y, z = QubitOperator('X0'), QubitOperator('Y0'), QubitOperator('Z0')
    assert x * y == 1j * z and y * z == 1j * x and z * x == 1j * y
    assert FermionOperator('3^ 2', 2.0).many_body_order() == 2
    op = FermionOperator('0^ 1') + FermionOperator('1^ 0')
    assert len(jordan_wigner(op).terms) == 2            # (XX + YY)/2
    with pytest.raises(ValueError):
        FermionOperator('0^ x')


def test_k_space_occupation_matches_oracle():
    for (nx, ny, up, dn) in [(2, 2, 2, 2), (2, 3, 3, 3), (2, 4, 4, 4), (3, 3, 5, 4)]:
        h = fermi_hubbard(nx, ny, 1.0, 4.0)
        kq = fourier_transform(get_quadratic_term(h), nx, ny)
        n = 2 * nx * ny
        e_up = {q: 0 for q in range(0, n, 2)}
        e_dn = {q: 0 for q in range(1, n, 2)}
        for term, c in kq.terms.items():
            assert term[0][0] == term[1][0]            # diagonal in k space
            (e_up if term[0][0] % 2 == 0 else e_dn)[term[0][0]] = complex(c).real
        got_up = sorted(e_up, key=e_up.get)[:up]
        got_dn = sorted(e_dn, key=e_dn.get)[:dn]
        oup, odn, _ = pauli.k_space_occupation(nx, ny, 1.0, up, dn)
        assert got_up == oup and got_dn == odn
        assert len(get_interacting_term(h).terms) == nx * ny


def test_packed_iqcc_dressing_equals_symbolic():
    """PauliTable.dressed == the reference's symbolic update H + sin(t)(-i/2)[H,P] + (1/2)(1-cos t)(PHP - H)
    (models/iqcc_hubbard.py:184-189), chained over several generators, on the 2x3 Hubbard Hamiltonian."""
    import numpy as np
    from fhsim.symbolic import QubitOperator, fermi_hubbard, jordan_wigner
    from fhsim.tables import PauliTable, pack_term
    n = 12
    h = jordan_wigner(fermi_hubbard(2, 3, 1.0, 4.0))
    tab = PauliTable.from_operator(h, n)
    gens = [("Y0 X1 X4 X5", 0.37), ("Y2 X8", -0.81), ("Y3 X5 X6 X10", 1.3), ("Y0 X1 X4 X5", -0.2)]
    for string, tau in gens:
        P = QubitOperator(string)
        first = np.sin(tau) * (-1j / 2) * (h * P - P * h)
        second = 1 / 2 * (1 - np.cos(tau)) * (P * h * P - h)
        h = h + first + second
        (term, _), = P.terms.items()
        xp, zp = pack_term(term, n)
        tab = tab.dressed(xp, zp, tau)
    want = {}
    for term, c in h.terms.items():
        if abs(c) > 1e-12:
            want[pack_term(term, n)] = complex(c)
    got = tab.as_dict()
    assert set(got) == set(want)
    assert max(abs(got[k] - want[k]) for k in want) < 1e-13
    assert max(abs(complex(c).imag) for c in got.values()) < 1e-13          # stays Hermitian with real coefficients
    back = PauliTable.from_operator(tab.to_operator(), n, compress=False).as_dict()
    assert back == got
