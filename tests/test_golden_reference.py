"""Host tables of the product AND of the oracle against fixtures produced by executing the reference's
own Python (tests/golden/make_golden.py; fixtures: tests/golden/reference_host_tables.json).

Bit-exact requirements (BASELINE north_star: "the ADAPT operator-selection sequence and the Pauli tables
must match bit-exactly"): pool order and content, k-space occupation, sector index order, HVA layer
colouring, split of H into quadratic / quartic parts.
"""
import json
import os

import numpy as np
import pytest

from fhsim.symbolic import FermionOperator, fermi_hubbard, jordan_wigner
from oracle import pauli

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_host_tables.json")) as _f:
    GOLD = json.load(_f)["lattices"]

LATTICES = [(2, 2, 4.0), (2, 3, 4.0), (3, 3, 6.0)]


def _terms(op):
    return [[[list(map(int, f)) for f in term], complex(c).real, complex(c).imag] for term, c in op.terms.items()]


def _same_terms(got, want, tol=0.0):
    assert len(got) == len(want)
    for (tg, rg, ig), (tw, rw, iw) in zip(got, want):
        assert tg == tw
        assert abs(rg - rw) <= tol and abs(ig - iw) <= tol


@pytest.mark.parametrize("nx,ny,u", LATTICES)
def test_pool_order_and_content_bit_exact(nx, ny, u):
    from operators.pool import hubbard_interaction_pool_simplified
    gold = GOLD[f"{nx}x{ny}"]["pool"]
    mine = hubbard_interaction_pool_simplified(nx, ny)
    assert len(mine) == len(gold)
    for op, want in zip(mine, gold):
        _same_terms(_terms(op), want)
    # the oracle's integer restatement (different route: no symbolic normal ordering)
    ora = pauli.pool_fermion_terms(nx, ny)
    assert len(ora) == len(gold)
    for op, want in zip(ora, gold):
        got = [[[list(f) for f in term], complex(c).real, complex(c).imag] for term, c in op]
        _same_terms(got, want)


@pytest.mark.parametrize("nx,ny,u", LATTICES)
def test_pool_pauli_tables_product_equals_oracle(nx, ny, u):
    """JW tables of every pool operator: product (fhsim.symbolic + tables.pack_term) == oracle (integer ladder algebra)."""
    from fhsim.tables import PauliTable
    from operators.pool import hubbard_interaction_pool_simplified
    n = 2 * nx * ny
    for op, terms in zip(hubbard_interaction_pool_simplified(nx, ny), pauli.pool_fermion_terms(nx, ny)):
        tab = PauliTable.from_operator(jordan_wigner(op), n, compress=False).as_dict()
        want = pauli.jw_table(terms, n)
        assert set(tab) == set(want)
        for k in tab:
            assert tab[k] == want[k]          # coefficients are +-1/8 exactly


@pytest.mark.parametrize("nx,ny,u", LATTICES)
def test_fourier_matrix_and_split(nx, ny, u):
    from operators.fourier import fourier_transform, fourier_transform_matrix
    from operators.tools import get_interacting_term, get_quadratic_term
    g = GOLD[f"{nx}x{ny}"]
    ft = fourier_transform_matrix(nx, ny)
    want = np.array(g["ft_matrix_re"]) + 1j * np.array(g["ft_matrix_im"])
    assert np.array_equal(ft, want)
    assert np.abs(pauli.ft_matrix(nx, ny) - want).max() < 1e-15
    ham = fermi_hubbard(nx, ny, 1.0, u)
    _same_terms(_terms(get_quadratic_term(ham)), g["quadratic_terms"])
    _same_terms(_terms(get_interacting_term(ham)), g["interacting_terms"])
    _same_terms(_terms(fourier_transform(get_quadratic_term(ham), nx, ny)), g["ft_quadratic_terms"], tol=1e-15)


@pytest.mark.parametrize("nx,ny,u", LATTICES)
def test_k_space_occupation(nx, ny, u):
    from models.common import get_non_interacting_ground_state_index
    from operators.fourier import fourier_transform
    from operators.tools import get_quadratic_term
    g = GOLD[f"{nx}x{ny}"]
    n_el, n_up, n_dn = g["sector"]
    n = g["n_qubits"]
    quad = fourier_transform(get_quadratic_term(fermi_hubbard(nx, ny, 1.0, u)), nx, ny)
    up, dn = get_non_interacting_ground_state_index(quad, n, n_up, n_dn, verbose=False)
    assert list(up) == g["occupied_up"] and list(dn) == g["occupied_dn"]
    oup, odn, _ = pauli.k_space_occupation(nx, ny, 1.0, n_up, n_dn)
    assert oup == g["occupied_up"] and odn == g["occupied_dn"]


@pytest.mark.parametrize("nx,ny,u", LATTICES)
def test_sector_indices(nx, ny, u):
    from linalg.exact_diagonalization import jw_number_spin_indices
    from oracle import ed
    g = GOLD[f"{nx}x{ny}"]
    n_el, n_up, n_dn = g["sector"]
    assert [int(v) for v in jw_number_spin_indices(n_el, n_up, n_dn, g["n_qubits"])] == g["sector_indices"]
    assert [int(v) for v in ed.jw_number_spin_indices(n_el, n_up, n_dn, g["n_qubits"])] == g["sector_indices"]


@pytest.mark.parametrize("nx,ny,u", LATTICES)
def test_hva_layers(nx, ny, u):
    from models.utils import get_hva_commuting_hopping_terms
    g = GOLD[f"{nx}x{ny}"]
    hset, vset = get_hva_commuting_hopping_terms(nx, ny, True)
    assert len(hset) == len(g["hva_horizontal"]) and len(vset) == len(g["hva_vertical"])
    for op, want in zip(hset, g["hva_horizontal"]):
        _same_terms(_terms(op), want)
    for op, want in zip(vset, g["hva_vertical"]):
        _same_terms(_terms(op), want)
    assert all(isinstance(op, FermionOperator) for op in hset + vset)


# pools no driver uses (reference operators/pool.py:48-131, 257-340): signatures kept, outputs pinned by the reference's
# own loops all the same (tests/golden/reference_unused_pools.json)
with open(os.path.join(HERE, "golden", "reference_unused_pools.json")) as _f:
    UNUSED = json.load(_f)


@pytest.mark.parametrize("key", sorted(UNUSED["spin_complemented_pool"]))
def test_spin_complemented_pool_matches_reference_including_its_stale_variable(key):
    from operators.pool import spin_complemented_pool
    n_el, n_orb, gen = (int(v) for v in key.split(","))
    mine = spin_complemented_pool(n_el, n_orb, bool(gen))
    gold = UNUSED["spin_complemented_pool"][key]
    assert len(mine) == len(gold)
    for op, want in zip(mine, gold):
        _same_terms(_terms(op), want)


@pytest.mark.parametrize("key", sorted(UNUSED["hubbard_interation_pool_modified"]))
def test_hubbard_interation_pool_modified_matches_reference(key):
    from operators.pool import hubbard_interation_pool_modified
    nx, ny = (int(v) for v in key.split("x"))
    mine = hubbard_interation_pool_modified(nx, ny)
    gold = UNUSED["hubbard_interation_pool_modified"][key]
    assert list(mine) == list(gold)
    for name in gold:
        _same_terms(_terms(mine[name]), gold[name], tol=1e-12)
