"""Host-side lowering of circuits to pair/diag ops and the tile scheduler, validated on the CPU
against the oracle through a test-only numpy interpreter of the op semantics."""
import numpy as np
import pytest

from emulate import run_circuit, run_items
from fhsim.circuit import Circuit, schedule, Marker
from fhsim.symbolic import FermionOperator, QubitOperator, fermi_hubbard, givens_decomposition_square, jordan_wigner
from fhsim.tables import GeneratorPlan
from operators.fourier import fourier_transform_matrix
from operators.pool import hubbard_interaction_pool_simplified
from operators.tools import get_interacting_term
from oracle import pauli, statevector as sv


def rand_state(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return v / np.linalg.norm(v)


@pytest.mark.parametrize("lat", [(2, 2), (2, 3)])
def test_pool_generator_rotation_equals_trotter_product(lat):
    nx, ny = lat
    n = 2 * nx * ny
    pool = hubbard_interaction_pool_simplified(nx, ny)
    opool = [pauli.jw_table(op, n) for op in pauli.pool_fermion_terms(nx, ny)]
    psi = rand_state(n, 3)
    for k in range(0, len(pool), 5):
        c = Circuit(n, 1)
        c.generator(GeneratorPlan(jordan_wigner(pool[k]), n), param=0)
        got = run_circuit(c, psi, [0.37])
        want = sv.trotterize(psi, 0.37, opool[k], n)
        assert np.abs(got - want).max() < 1e-13


def test_single_qubit_and_cnot_gates():
    n = 5
    psi = rand_state(n, 1)
    c = Circuit(n, 3)
    c.rx(0.3, 1); c.ry(-0.7, 2); c.rz(1.1, 4); c.cnot(0, 3); c.pauli_x(2); c.cnot(4, 1)
    c.rx(0, 0, param=0); c.ry(0, 3, param=1); c.rz(0, 2, param=2)
    th = [0.21, -0.4, 0.9]
    want = sv.rx(psi, 0.3, 1, n); want = sv.ry(want, -0.7, 2, n); want = sv.rz(want, 1.1, 4, n)
    want = sv.cnot(want, 0, 3, n); want = sv.pauli_x(want, 2, n); want = sv.cnot(want, 4, 1, n)
    want = sv.rx(want, th[0], 0, n); want = sv.ry(want, th[1], 3, n); want = sv.rz(want, th[2], 2, n)
    assert np.abs(run_circuit(c, psi, th) - want).max() < 1e-13


def test_single_excitation_both_wire_orders():
    n = 4
    psi = rand_state(n, 2)
    for (i, j) in [(1, 2), (2, 1), (0, 3)]:
        c = Circuit(n)
        c.single_excitation(0.83, i, j)
        assert np.abs(run_circuit(c, psi) - sv.single_excitation(psi, 0.83, i, j, n)).max() < 1e-13


@pytest.mark.parametrize("lat", [(2, 2), (3, 1), (2, 3)])
def test_basis_change_matches_oracle(lat):
    nx, ny = lat
    n = 2 * nx * ny
    dec, diag = givens_decomposition_square(fourier_transform_matrix(nx, ny))
    c = Circuit(n)
    c.basis_change(diag, list(reversed(dec)))
    psi = rand_state(n, 5)
    assert np.abs(run_circuit(c, psi) - sv.basis_change(psi, diag, dec, n)).max() < 1e-12


def test_hva_generators_and_pauli_rotation():
    from models.utils import get_hva_commuting_hopping_terms
    nx, ny = 2, 3
    n = 12
    h_sets, v_sets = get_hva_commuting_hopping_terms(nx, ny, True)
    coulomb = jordan_wigner(get_interacting_term(fermi_hubbard(nx, ny, 1.0, 4.0)))
    psi = rand_state(n, 7)
    for gen in [jordan_wigner(g) for g in h_sets + v_sets] + [coulomb]:
        plan = GeneratorPlan(gen, n)
        assert plan.exact
        c = Circuit(n, 1)
        c.generator(plan, param=0)
        table = {}
        from fhsim.tables import pack_term
        for term, coef in gen.terms.items():
            table[pack_term(term, n)] = coef
        want = sv.trotterize(psi, -0.43, table, n)
        assert np.abs(run_circuit(c, psi, [-0.43]) - want).max() < 1e-12
    # iQCC style single string Y X X
    q = QubitOperator('Y1 X4 X9')
    (x, z), = [pack_term(t, n) for t in q.terms]
    c = Circuit(n, 1)
    c.pauli_rotation(x, z, 0.5, param=0)
    assert np.abs(run_circuit(c, psi, [0.77]) - sv.pauli_rotation(psi, 0.77, x, z, n)).max() < 1e-13


def test_non_commuting_generator_falls_back_to_literal_product():
    n = 3
    gen = QubitOperator('X0 Z1', 0.3) + QubitOperator('Z0', -0.2) + QubitOperator('Y2', 0.7)
    plan = GeneratorPlan(gen, n)
    assert not plan.exact and len(plan.pieces) == 3
    c = Circuit(n, 1)
    c.generator(plan, param=0)
    from fhsim.tables import pack_term
    table = {pack_term(t, n): cf for t, cf in gen.terms.items()}
    psi = rand_state(n, 9)
    assert np.abs(run_circuit(c, psi, [0.9]) - sv.trotterize(psi, 0.9, table, n)).max() < 1e-13


@pytest.mark.parametrize("tile_bits,low_bits", [(6, 1), (8, 2), (12, 1)])
def test_scheduler_preserves_the_unitary(tile_bits, low_bits):
    nx, ny = 2, 3
    n = 12
    pool = hubbard_interaction_pool_simplified(nx, ny)
    dec, diag = givens_decomposition_square(fourier_transform_matrix(nx, ny))
    rng = np.random.default_rng(11)
    picks = rng.choice(len(pool), size=14, replace=False)
    c = Circuit(n, len(picks))
    for p, k in enumerate(picks):
        c.generator(GeneratorPlan(jordan_wigner(pool[k]), n), param=p)
    c.basis_change(diag, list(reversed(dec)))
    th = rng.uniform(-0.5, 0.5, len(picks))
    psi = rand_state(n, 13)
    want = run_circuit(c, psi, th)
    ops = [o for o in c.ops if not isinstance(o, Marker)]
    import fhsim.circuit as fc
    old = fc._LAUNCH_BYTES
    fc._LAUNCH_BYTES = 1e12              # force fusion so the reordering logic is exercised
    try:
        items = schedule(ops, n, tile_bits, low_bits)
    finally:
        fc._LAUNCH_BYTES = old
    assert sum(1 if it[0] == "op" else len(it[2]) for it in items) == len(ops)
    assert any(it[0] == "tile" for it in items)
    assert np.abs(run_items(items, psi, th, n) - want).max() < 1e-12


def test_noncommuting_bitsets_equal_the_pairwise_string_rule():
    """The scheduler's vectorised commutation table answers exactly what the pairwise string-level rule answers,
    including ops without strings (commute with everything) and circuits with no strings at all."""
    from fhsim.circuit import DiagOpSpec, PairOpSpec, _noncommuting_bitsets, _ops_commute, absorb_phases
    nx, ny, n = 2, 3, 12
    pool = hubbard_interaction_pool_simplified(nx, ny)
    rng = np.random.default_rng(5)
    c = Circuit(n, 20)
    for p, k in enumerate(rng.choice(len(pool), size=20, replace=False)):
        c.generator(GeneratorPlan(jordan_wigner(pool[k]), n), param=p)
    for w in range(n):
        c.ry(0.3, w)
        c.rz(0.2, w)
    for w in range(n - 1):
        c.cnot(w, w + 1)
    c.generator(GeneratorPlan(jordan_wigner(get_interacting_term(fermi_hubbard(nx, ny, 1.0, 4.0))), n), angle=0.3)
    c.basis_change_separable(nx, ny)
    ops = absorb_phases(c.ops)
    ops.insert(7, PairOpSpec(1, 1, 0, 0))                 # an op that declares no strings
    ops.append(DiagOpSpec([3], [0.1]))
    sets = _noncommuting_bitsets(ops)
    assert len(sets) == len(ops)
    for i, a in enumerate(ops):
        for j, b in enumerate(ops):
            assert bool(sets[i] >> j & 1) == (not _ops_commute(a, b)), (i, j)
    assert _noncommuting_bitsets([PairOpSpec(1, 1, 0, 0), DiagOpSpec([1], [0.5])]) == [0, 0]
    assert _noncommuting_bitsets([]) == []


@pytest.mark.parametrize("lat", [(2, 2), (3, 1), (2, 3), (3, 2), (3, 3)])
def test_separable_basis_change_equals_reference_network(lat):
    """W compiled from the tensor structure of the FT matrix (36 fermionic Givens at 3x3) is the
    reference's 144-rotation network up to the global phase both vacuum phases predict."""
    nx, ny = lat
    n = 2 * nx * ny
    dec, diag = givens_decomposition_square(fourier_transform_matrix(nx, ny))
    ref = Circuit(n)
    ref.basis_change(diag, list(reversed(dec)))
    fast = Circuit(n)
    ph_fast = fast.basis_change_separable(nx, ny)
    ph_ref = Circuit.basis_change_vacuum_phase(diag, dec)
    assert len(fast.ops) < len(ref.ops)
    psi = rand_state(n, 17)
    want = sv.basis_change(psi, diag, dec, n)
    got = run_circuit(fast, psi) * np.exp(-1j * (ph_fast - ph_ref))
    assert np.abs(got - want).max() < 1e-12


@pytest.mark.parametrize("lat", [(2, 2), (2, 3), (3, 3)])
def test_phase_absorption_is_exact(lat):
    from fhsim.circuit import absorb_phases, DiagOpSpec
    nx, ny = lat
    n = 2 * nx * ny
    c = Circuit(n, 2)
    c.rz(0.37, 1)
    c.basis_change_separable(nx, ny)
    c.ry(0.0, 2, param=0)                       # cannot absorb (pattern does not pin x): forces a flush
    c.rz(-0.61, 2)
    pool = hubbard_interaction_pool_simplified(nx, ny)
    c.generator(GeneratorPlan(jordan_wigner(pool[3]), n), param=1)
    psi = rand_state(n, 23)
    th = [0.4, -0.3]
    want = run_circuit(c, psi, th)
    ops = [o for o in c.ops if not isinstance(o, Marker)]
    fused = absorb_phases(ops)
    assert sum(isinstance(o, DiagOpSpec) for o in fused) < sum(isinstance(o, DiagOpSpec) for o in ops)
    d = Circuit(n, 2)
    d.ops = fused
    assert np.abs(run_circuit(d, psi, th) - want).max() < 1e-13
    # and scheduling the absorbed ops still preserves the unitary
    items = schedule(fused, n, min(8, n), 1)
    assert np.abs(run_items(items, psi, th, n) - want).max() < 1e-12
