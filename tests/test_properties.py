"""Property tests (hypothesis) of the oracle and the host compile step -- SURVEY 4, item 3:
norm preservation, U(theta) U(-theta) = 1, Hermiticity <phi|H psi> = conj <psi|H phi>, commutation of the 8 strings of
every pool operator, gradient = central finite difference, scheduler / lowering invariants on random circuits."""
import numpy as np
from hypothesis import example, given, settings, strategies as st

from emulate import apply_op, run_circuit, run_items
from fhsim.circuit import Circuit, absorb_phases, schedule
from fhsim.sharded import QubitLayout, dagger_ops, lower_diag, lower_pair, plan_circuit, swap_steps
from fhsim.tables import PauliTable, strings_commute
from oracle import pauli, statevector as sv

N = 6
masks = st.integers(min_value=0, max_value=(1 << N) - 1)
angles = st.floats(min_value=-3.0, max_value=3.0, allow_nan=False)


def _state(seed, n=N):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return v / np.linalg.norm(v)


@settings(max_examples=60, deadline=None)
@given(x=masks.filter(lambda v: v != 0), z=masks, theta=angles, seed=st.integers(0, 10 ** 6))
def test_pauli_rotation_is_unitary_and_invertible(x, z, theta, seed):
    psi = _state(seed)
    out = sv.pauli_rotation(psi, theta, x, z, N)
    assert abs(np.linalg.norm(out) - 1.0) < 1e-12
    back = sv.pauli_rotation(out, -theta, x, z, N)
    assert np.abs(back - psi).max() < 1e-12
    # device-op form of the same rotation (host compile) == oracle closed form
    c = Circuit(N, 0)
    c.pauli_rotation(x, z, 0.5, angle=theta)
    assert np.abs(run_circuit(c, psi) - out).max() < 1e-12


@settings(max_examples=40, deadline=None)
@given(terms=st.lists(st.tuples(masks, masks, st.floats(-2, 2, allow_nan=False)), min_size=1, max_size=8),
       s1=st.integers(0, 10 ** 6), s2=st.integers(0, 10 ** 6))
def test_real_coefficient_tables_are_hermitian(terms, s1, s2):
    table = {}
    for x, z, c in terms:
        table[(x, z)] = table.get((x, z), 0.0) + c
    psi, phi = _state(s1), _state(s2)
    a = np.vdot(phi, sv.apply_table(psi, table, N))
    b = np.vdot(psi, sv.apply_table(phi, table, N))
    assert abs(a - np.conj(b)) < 1e-11


def test_pool_strings_commute_and_share_one_x_mask():
    for nx, ny in ((2, 2), (2, 3), (3, 3)):
        n = 2 * nx * ny
        for op in pauli.pool_fermion_terms(nx, ny):
            tab = pauli.jw_table(op, n)
            keys = [k for k in tab if k != (0, 0)]
            assert len(keys) == 8 and len({k[0] for k in keys}) == 1
            assert all(abs(abs(tab[k]) - 0.125) < 1e-15 and abs(complex(tab[k]).imag) < 1e-15 for k in keys)
            assert all(strings_commute(a, b) for a in keys for b in keys)


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10 ** 6), k=st.integers(0, 23), theta=st.floats(-1.0, 1.0, allow_nan=False))
def test_pool_gradient_equals_central_difference(seed, k, theta):
    """d/de <psi| e^{+i e G} H e^{-i e G} |psi> at e = 0 equals 2 Im <H psi | G psi> (the screening formula)."""
    n = 8
    h = pauli.compress(pauli.jw_table(pauli.hubbard_fermion_terms(2, 2, 1.0, 4.0), n))
    pool = [pauli.jw_table(op, n) for op in pauli.pool_fermion_terms(2, 2)]
    psi = _state(seed, n)
    g = pool[k]
    lam = sv.apply_table(psi, h, n)
    want = 2.0 * np.vdot(lam, sv.apply_table(psi, g, n)).imag
    hstep = 1e-5
    ep = sv.expval(sv.trotterize(psi, hstep, g, n), h, n).real
    em = sv.expval(sv.trotterize(psi, -hstep, g, n), h, n).real
    assert abs((ep - em) / (2 * hstep) - want) < 5e-9


def _random_circuit(rng, n, n_ops, max_weight=None):
    c = Circuit(n, 3)
    for _ in range(n_ops):
        kind = int(rng.integers(0, 6))
        if kind == 0:
            c.rz(float(rng.uniform(-2, 2)), int(rng.integers(n)))
        elif kind == 1:
            c.ry(0.0, int(rng.integers(n)), param=int(rng.integers(3)))
        elif kind == 2:
            a, b = rng.choice(n, size=2, replace=False)
            c.cnot(int(a), int(b))
        elif kind == 3:
            a, b = rng.choice(n, size=2, replace=False)
            c.fermionic_single_excitation(float(rng.uniform(-2, 2)), int(a), int(b))
        elif kind == 4:
            x = int(rng.integers(1, 1 << n))
            if max_weight is not None:          # a sharded slab must hold every x bit of an op plus g spare qubits
                bits = [b for b in range(n) if x >> b & 1][:max_weight]
                x = sum(1 << b for b in bits)
            c.pauli_rotation(x, int(rng.integers(0, 1 << n)), 0.5, param=int(rng.integers(3)))
        else:
            c.rx(float(rng.uniform(-2, 2)), int(rng.integers(n)))
    return c


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10 ** 6), tile_bits=st.integers(3, 6), low_bits=st.integers(0, 2))
def test_scheduler_and_phase_absorption_preserve_random_circuits(seed, tile_bits, low_bits):
    rng = np.random.default_rng(seed)
    n = 6
    c = _random_circuit(rng, n, 14)
    thetas = rng.uniform(-1, 1, 3)
    psi = _state(seed + 1, n)
    want = run_circuit(c, psi, thetas)
    ops = absorb_phases(c.ops)
    items = schedule(ops, n, tile_bits, min(low_bits, tile_bits))
    got = run_items(items, psi, thetas, n)
    assert np.abs(got - want).max() < 1e-11
    assert abs(np.linalg.norm(got) - 1.0) < 1e-11


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10 ** 6), g=st.integers(1, 2))
def test_sharded_plan_equals_global_circuit(seed, g):
    """Random circuit, rank-by-rank lowering along the planner's swap schedule == the circuit on the full vector;
    then the planner's inverse brings the state back."""
    rng = np.random.default_rng(seed)
    n = 7
    nl = n - g
    c = _random_circuit(rng, n, 12, max_weight=3)
    thetas = rng.uniform(-1, 1, 3)
    psi = _state(seed + 7, n)
    want = run_circuit(c, psi, thetas)

    def run_plan(ops, vec_logical, layout):
        idx = np.arange(1 << n, dtype=np.uint64)

        def phys_index(lay):
            p = np.zeros(1 << n, dtype=np.uint64)
            for b in range(n):
                p |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(lay.perm[b])
            return p

        steps, final = plan_circuit(ops, layout)
        full = np.zeros(1 << n, complex)
        full[phys_index(layout)] = vec_logical
        for step in steps:
            if step[0] == "swap":
                logical = full[phys_index(step[2])]
                full = np.zeros(1 << n, complex)
                full[phys_index(step[3])] = logical           # a relayout moves amplitudes, nothing else
                continue
            _, seg, lay = step
            for rank in range(1 << g):
                slab = full[rank << nl:(rank + 1) << nl].copy()
                for op in seg:
                    lowered = [lower_diag(op, lay, rank)] if hasattr(op, "coef") else lower_pair(op, lay, rank)
                    for lo in lowered:
                        slab = apply_op(slab, lo, thetas, nl)
                full[rank << nl:(rank + 1) << nl] = slab
        return full[phys_index(final)], final

    got, final = run_plan(c.ops, psi, QubitLayout(n, g))
    assert np.abs(got - want).max() < 1e-11
    back, _ = run_plan(dagger_ops(c.ops), got, final)
    assert np.abs(back - psi).max() < 1e-11


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 10 ** 6))
def test_swap_steps_is_a_permutation_with_the_requested_rank_bits(seed):
    rng = np.random.default_rng(seed)
    n, g = 9, int(rng.integers(1, 4))
    lay = QubitLayout(n, g, list(rng.permutation(n)))
    local = [b for b in range(n) if lay.perm[b] < n - g]
    new_globals = [int(v) for v in rng.choice(local, size=g, replace=False)]
    pairs, new = swap_steps(lay, new_globals)
    assert sorted(new.perm) == list(range(n))
    assert sorted(new.global_logical_bits()) == sorted(new_globals)
    assert len(pairs) <= g and all(0 <= a < n - g and 0 <= b < n - g for a, b in pairs)


@settings(max_examples=30, deadline=None)
@given(seed=st.integers(0, 10 ** 6), tau=angles)
@example(seed=267, tau=0.0)        # draws the string (x=18, z=0) twice: duplicates must be merged, not dropped
def test_dressing_preserves_the_spectrum(seed, tau):
    """exp(i tau P/2) H exp(-i tau P/2) is a similarity transform: <psi'|H'|psi'> with psi' = exp(i tau P/2) psi
    equals <psi|H|psi>."""
    rng = np.random.default_rng(seed)
    n = 5
    xs = rng.integers(0, 1 << n, size=6)
    zs = rng.integers(0, 1 << n, size=6)
    # Hermitian table: real coefficient times a Hermitian string (i^k convention makes every (x, z) Hermitian)
    table = PauliTable(n, xs, zs, rng.normal(size=6))
    xp, zp = int(rng.integers(1, 1 << n)), int(rng.integers(0, 1 << n))
    dressed = table.dressed(xp, zp, tau)
    psi = _state(seed + 3, n)
    e0 = sv.expval(psi, table.as_dict(), n).real
    psi_rot = sv.pauli_rotation(psi, -tau, xp, zp, n)       # exp(+i tau P / 2) psi
    e1 = sv.expval(psi_rot, dressed.as_dict(), n).real
    assert abs(e0 - e1) < 1e-11


def test_pauli_table_merges_duplicate_strings():
    """One entry per string: duplicates are summed in first-seen order, so as_dict / dressed / upload agree."""
    t = PauliTable(5, [18, 3, 18, 7], [0, 1, 0, 2], [0.5, 1.0, 0.25, -2.0])
    assert [int(v) for v in t.x] == [18, 3, 7] and [int(v) for v in t.z] == [0, 1, 2]
    assert np.allclose(t.coeff, [0.75, 1.0, -2.0])
    assert t.as_dict() == {(18, 0): 0.75, (3, 1): 1.0, (7, 2): -2.0}
    same = t.dressed(1, 0, 0.0)
    assert same.as_dict() == t.as_dict()
