"""Parity of the CUDA path (through the C-ABI) against the oracle on the same seeded inputs.

Tolerances: amplitudes 1e-12 abs, energies 1e-10 abs, gradients 1e-9 abs (BASELINE.json north_star).
"""
import numpy as np
import pytest

from fhsim.backend import Context, DevicePool, DeviceTable, State, lanczos, lanczos_sector, sector_indices
from fhsim.circuit import Circuit, DiagOpSpec
from fhsim.symbolic import QubitOperator, fermi_hubbard, givens_decomposition_square, jordan_wigner
from fhsim.tables import GeneratorPlan, PauliTable, pack_term
from operators.fourier import fourier_transform_matrix
from operators.pool import hubbard_interaction_pool_simplified
from oracle import ed, pauli, statevector as sv

pytestmark = pytest.mark.gpu

AMP_TOL, E_TOL, G_TOL = 1e-12, 1e-10, 1e-9


@pytest.fixture(scope="module")
def ctx():
    return Context(0)


def rand_state(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return v / np.linalg.norm(v)


def lattice(nx, ny, u):
    n = 2 * nx * ny
    h_op = fermi_hubbard(nx, ny, 1.0, u)
    h_tab = PauliTable.from_operator(h_op, n)
    pool_ops = [jordan_wigner(g) for g in hubbard_interaction_pool_simplified(nx, ny)]
    dec, diag = givens_decomposition_square(fourier_transform_matrix(nx, ny))
    o_h = pauli.compress(pauli.jw_table(pauli.hubbard_fermion_terms(nx, ny, 1.0, u), n))
    o_pool = [pauli.jw_table(op, n) for op in pauli.pool_fermion_terms(nx, ny)]
    return n, h_tab, pool_ops, dec, diag, o_h, o_pool


@pytest.mark.parametrize("n", [3, 8, 12, 18])
def test_pauli_rotation_batch_vs_oracle(ctx, n):
    rng = np.random.default_rng(n)
    psi = rand_state(n, 100 + n)
    xs = rng.integers(0, 1 << n, size=12, dtype=np.uint64)
    zs = rng.integers(0, 1 << n, size=12, dtype=np.uint64)
    xs[3] = 0                      # a diagonal string
    xs[5], zs[5] = 0, 0            # identity: skipped
    xs[7] = 1                      # lowest-bit flip
    xs[8] = 1 << (n - 1)           # highest-bit flip
    th = rng.uniform(-3, 3, size=12)
    st = State.from_numpy(ctx, psi)
    st.apply_pauli_rotations(xs, zs, 0.5 * th)
    want = psi
    for x, z, t in zip(xs, zs, th):
        if x == 0 and z == 0:
            continue
        want = sv.pauli_rotation(want, t, int(x), int(z), n)
    assert np.abs(st.numpy() - want).max() < AMP_TOL
    assert abs(st.norm2() - 1.0) < 1e-12


def test_rotation_inverse_is_identity(ctx):
    n = 14
    psi = rand_state(n, 5)
    st = State.from_numpy(ctx, psi)
    x, z = np.array([0b10110000110011], np.uint64), np.array([0b00110100010001], np.uint64)
    st.apply_pauli_rotations(x, z, [0.731])
    st.apply_pauli_rotations(x, z, [-0.731])
    assert np.abs(st.numpy() - psi).max() < 1e-14


@pytest.mark.parametrize("lat,u", [((2, 2), 4.0), ((2, 3), 4.0), ((3, 3), 6.0)])
def test_apply_table_and_expval(ctx, lat, u):
    n, h_tab, _, _, _, o_h, _ = lattice(*lat, u)
    psi = rand_state(n, 21)
    dtab = DeviceTable(ctx, h_tab)
    info = dtab.info()
    assert info["n_terms"] == len(o_h)
    st, out = State.from_numpy(ctx, psi), State(ctx, n)
    e = dtab.apply(st, out)
    want = sv.apply_table(psi, o_h, n)
    assert np.abs(out.numpy() - want).max() < 1e-12
    assert abs(e - np.vdot(psi, want)) < E_TOL
    assert abs(dtab.expval(st) - np.vdot(psi, want).real) < E_TOL
    # Hermiticity <phi|H psi> = conj <psi|H phi>
    phi = rand_state(n, 22)
    sp, hp = State.from_numpy(ctx, phi), State(ctx, n)
    dtab.apply(sp, hp)
    assert abs(sp.inner(out) - np.conj(st.inner(hp))) < 1e-11


def test_complex_coefficient_table(ctx):
    n = 6
    op = QubitOperator('X0 Y2', 0.3 + 0.2j) + QubitOperator('Z1 Z5', -0.7) + QubitOperator('Y3', 1.1j) + QubitOperator((), 0.25)
    tab = PauliTable.from_operator(op, n)
    psi = rand_state(n, 8)
    st, out = State.from_numpy(ctx, psi), State(ctx, n)
    e = DeviceTable(ctx, tab).apply(st, out)
    want = sv.apply_table(psi, tab.as_dict(), n)
    assert np.abs(out.numpy() - want).max() < 1e-13
    assert abs(e - np.vdot(psi, want)) < 1e-12


@pytest.mark.parametrize("n,seed", [(5, 1), (9, 2), (12, 3), (14, 4)])
def test_random_tables_all_group_shapes(ctx, n, seed):
    """K2 on random Hermitian-free (complex coefficient) tables: x-masks of weight 0..n (both the tabulated
    <=4-bit groups and the per-term fallback), several zeta classes per group, repeated x-masks."""
    rng = np.random.default_rng(seed)
    full = (1 << n) - 1
    xs = [0, 0, 0]
    for w in (1, 2, 2, 3, 4, 4, 5, min(7, n)):
        bits = rng.choice(n, size=w, replace=False)
        xs.append(int(sum(1 << int(b) for b in bits)))
    table = {}
    for x in xs:
        for _ in range(int(rng.integers(1, 6))):
            z = int(rng.integers(0, full + 1))
            if x == 0 and z == 0:
                continue
            table[(x, z)] = complex(rng.normal(), rng.normal())
    table[(0, 0)] = 0.37
    keys = list(table)
    tab = PauliTable(n, [k[0] for k in keys], [k[1] for k in keys], [table[k] for k in keys])
    psi = rand_state(n, 100 + seed)
    st, out = State.from_numpy(ctx, psi), State(ctx, n)
    e = DeviceTable(ctx, tab).apply(st, out)
    want = sv.apply_table(psi, table, n)
    assert np.abs(out.numpy() - want).max() < 1e-12
    assert abs(e - np.vdot(psi, want)) < 1e-11
    # real coefficients only -> the REAL kernel instantiation; only strings with an even number of Y's
    # have a real weight table, so keep those
    rtable = {k: v.real for k, v in table.items() if bin(k[0] & k[1]).count("1") % 2 == 0}
    rkeys = list(rtable)
    rtab = PauliTable(n, [k[0] for k in rkeys], [k[1] for k in rkeys], [rtable[k] for k in rkeys])
    e2 = DeviceTable(ctx, rtab).apply(st, out)
    want2 = sv.apply_table(psi, rtable, n)
    assert np.abs(out.numpy() - want2).max() < 1e-12
    assert abs(e2 - np.vdot(psi, want2)) < 1e-11


def test_diag_factor_tables_20_qubits(ctx):
    """n >= 20 and > 4 terms: the standalone diagonal kernel goes through three factor tables (index bits 0-11, 12-23,
    24-) plus per-amplitude evaluation of the terms that straddle two chunks."""
    from fhsim.circuit import DiagOpSpec
    n = 20
    rng = np.random.default_rng(17)
    zs = [1 << 3, (1 << 11) | (1 << 10), (1 << 12) | (1 << 11), (1 << 19) | 1, (1 << 13) | (1 << 15), 0b111 << 10,
          (1 << 19) | (1 << 18), 1 << 12, (1 << 5) | (1 << 17) | (1 << 2)]
    angles = rng.uniform(-1.0, 1.0, len(zs))
    psi = rand_state(n, 23)
    st = State.from_numpy(ctx, psi)
    st.apply_diag(zs, angles)
    idx = np.arange(1 << n, dtype=np.uint64)
    tot = np.zeros(1 << n)
    for z, a in zip(zs, angles):
        tot += a * (1 - 2 * (np.bitwise_count(idx & np.uint64(z)) & 1).astype(np.int64))
    assert np.abs(st.numpy() - psi * np.exp(-1j * tot)).max() < AMP_TOL
    # through a compiled program, forward then inverse
    c = Circuit(n, 1)
    c.ops.append(DiagOpSpec(zs, list(angles), 0, [(0, z) for z in zs]))
    prog = c.compile(ctx, fuse=False)
    st.upload(psi)
    prog.run(st, [0.7])
    assert np.abs(st.numpy() - psi * np.exp(-0.7j * tot)).max() < AMP_TOL
    prog.run(st, [0.7], dagger=True)
    assert np.abs(st.numpy() - psi).max() < AMP_TOL


def test_apply_table_22_qubits_eight_outputs_per_thread(ctx):
    """n >= 22 switches K2 to 8 outputs per thread (index bits 8..10): 1x11 Hubbard chain + a complex random table
    whose x / z masks hit bits 8, 9, 10."""
    n = 22
    o_h = pauli.compress(pauli.jw_table(pauli.hubbard_fermion_terms(11, 1, 1.0, 4.0), n))
    rng = np.random.default_rng(9)
    psi = rand_state(n, 31)
    st, out = State.from_numpy(ctx, psi), State(ctx, n)
    keys = list(o_h)
    tab = PauliTable(n, [k[0] for k in keys], [k[1] for k in keys], [o_h[k] for k in keys])
    e = DeviceTable(ctx, tab).apply(st, out)
    want = sv.apply_table(psi, o_h, n)
    assert np.abs(out.numpy() - want).max() < 1e-12
    assert abs(e - np.vdot(psi, want)) < E_TOL
    table = {}
    for x in (0, 1 << 8, (1 << 9) | (1 << 3), (1 << 10) | (1 << 8) | (1 << 20), (1 << 9) | (1 << 10) | 1 | (1 << 15),
              (1 << 21) | (1 << 2), 0b11111 << 7):
        for _ in range(3):
            z = int(rng.integers(0, 1 << n)) | (int(rng.integers(0, 8)) << 8)
            if x or z:
                table[(x, z)] = complex(rng.normal(), rng.normal())
    keys = list(table)
    tab = PauliTable(n, [k[0] for k in keys], [k[1] for k in keys], [table[k] for k in keys])
    e = DeviceTable(ctx, tab).apply(st, out)
    want = sv.apply_table(psi, table, n)
    assert np.abs(out.numpy() - want).max() < 1e-12
    assert abs(e - np.vdot(psi, want)) < 1e-10


def test_apply_table_accumulate_splits(ctx):
    """out = H1 psi, then out += H2 psi equals (H1+H2) psi: the sum-of-partial-tables mode used by the sharded path."""
    from fhsim import _cabi
    n, h_tab, _, _, _, o_h, _ = lattice(2, 3, 4.0)
    keys = list(o_h)
    half = len(keys) // 2
    parts = [dict((k, o_h[k]) for k in keys[:half]), dict((k, o_h[k]) for k in keys[half:])]
    psi = rand_state(n, 77)
    st, out = State.from_numpy(ctx, psi), State(ctx, n)
    tabs = [DeviceTable(ctx, PauliTable(n, [k[0] for k in p], [k[1] for k in p], [p[k] for k in p])) for p in parts]
    e1 = tabs[0].apply(st, out)
    re, im = _cabi.C.c_double(), _cabi.C.c_double()
    _cabi.check(_cabi.lib().fh_apply_table_accumulate(tabs[1]._h, st._h, out._h, _cabi.C.byref(re), _cabi.C.byref(im)))
    want = sv.apply_table(psi, o_h, n)
    assert np.abs(out.numpy() - want).max() < 1e-12
    assert abs(e1 + complex(re.value, im.value) - np.vdot(psi, want)) < E_TOL


def test_spin_squared_table(ctx):
    """S^2 (442 terms at 3x3 in the reference; 187 here at 2x3): 4-bit x-mask groups, <S^2> of a sector state."""
    from models.common import get_spin_operators
    n = 12
    s2 = jordan_wigner(get_spin_operators(6, 'S^2'))
    tab = PauliTable.from_operator(s2, n)
    psi = rand_state(n, 5)
    st, out = State.from_numpy(ctx, psi), State(ctx, n)
    e = DeviceTable(ctx, tab).apply(st, out)
    want = sv.apply_table(psi, tab.as_dict(), n)
    assert np.abs(out.numpy() - want).max() < 1e-11
    assert abs(e - np.vdot(psi, want)) < E_TOL


@pytest.mark.parametrize("lat,u,up,dn", [((2, 2), 4.0, 2, 2), ((2, 3), 4.0, 3, 3)])
def test_pool_gradients_vs_oracle(ctx, lat, u, up, dn):
    n, h_tab, pool_ops, dec, diag, o_h, o_pool = lattice(*lat, u)
    rng = np.random.default_rng(1234)
    occ_up, occ_dn, _ = pauli.k_space_occupation(*lat, 1.0, up, dn)
    picks = list(range(1, len(o_pool), max(1, len(o_pool) // 6)))[:6]
    th = rng.uniform(-0.3, 0.3, len(picks))
    psi = sv.adapt_state(n, occ_up + occ_dn, [o_pool[k] for k in picks], th)
    g_want, e_want, lam = sv.pool_gradients(psi, o_h, o_pool, diag, dec, n)
    plans = [GeneratorPlan(g, n) for g in pool_ops]
    dpool = DevicePool(ctx, plans, n)
    got = dpool.gradients(State.from_numpy(ctx, psi), State.from_numpy(ctx, lam))
    assert np.abs(got - g_want).max() < G_TOL
    part = dpool.gradients(State.from_numpy(ctx, psi), State.from_numpy(ctx, lam), first=5, count=7)
    assert np.array_equal(part, got[5:12])          # pool sharding gives bit-identical slices


@pytest.mark.parametrize("fuse,separable", [(False, False), (True, False), (True, True)])
@pytest.mark.parametrize("lat,u,up,dn", [((2, 2), 4.0, 2, 2), ((2, 3), 4.0, 3, 3)])
def test_program_evaluate_energy_grads_pool(ctx, lat, u, up, dn, fuse, separable):
    n, h_tab, pool_ops, dec, diag, o_h, o_pool = lattice(*lat, u)
    rng = np.random.default_rng(99)
    occ_up, occ_dn, _ = pauli.k_space_occupation(*lat, 1.0, up, dn)
    picks = list(rng.choice(len(pool_ops), size=7, replace=False))
    th = rng.uniform(-0.4, 0.4, len(picks))
    plans = [GeneratorPlan(g, n) for g in pool_ops]
    circ = Circuit(n, len(picks))
    for p, k in enumerate(picks):
        circ.generator(plans[k], param=p)
    circ.marker("ansatz_end")
    phase = 0.0
    if separable:
        phase = circ.basis_change_separable(*lat) - Circuit.basis_change_vacuum_phase(diag, dec)
    else:
        circ.basis_change(diag, list(reversed(dec)))
    prog = circ.compile(ctx, fuse=fuse)
    dtab = DeviceTable(ctx, h_tab)
    dpool = DevicePool(ctx, plans, n)
    basis = sum(1 << (n - 1 - q) for q in occ_up + occ_dn)
    out_state = State(ctx, n)
    res = prog.evaluate(basis, th, [dtab], grads=True, pool=dpool, pool_pos=prog.markers["ansatz_end"],
                        state_out=out_state)
    gens = [o_pool[k] for k in picks]
    e_want, g_want = sv.adjoint_gradient(n, occ_up + occ_dn, gens, th, o_h, diag, dec)
    psi_k = sv.adapt_state(n, occ_up + occ_dn, gens, th)
    pg_want, _, _ = sv.pool_gradients(psi_k, o_h, o_pool, diag, dec, n)
    assert abs(res["expvals"][0] - e_want) < E_TOL
    assert np.abs(res["grads"] - g_want).max() < G_TOL
    assert np.abs(res["pool"] - pg_want).max() < G_TOL
    phi = sv.basis_change(psi_k, diag, dec, n)
    assert np.abs(out_state.numpy() * np.exp(-1j * phase) - phi).max() < 1e-12
    # replay of the captured graph with new parameters
    th2 = th + 0.05
    res2 = prog.evaluate(basis, th2, [dtab], grads=True, pool=dpool, pool_pos=prog.markers["ansatz_end"],
                         state_out=out_state)
    e2, g2 = sv.adjoint_gradient(n, occ_up + occ_dn, gens, th2, o_h, diag, dec)
    assert abs(res2["expvals"][0] - e2) < E_TOL and np.abs(res2["grads"] - g2).max() < G_TOL


@pytest.mark.parametrize("tile_bits", [3, 4, 7, 11, 12, 13])
def test_tile_runs_random_gate_circuits(ctx, tile_bits):
    """k_tile on random gate sequences (1-qubit rotations, CNOT, Givens, fermionic Givens, RZ, Pauli rotations with
    Z tails, X) clustered on wire triples: every tile size, forward and dagger, against the numpy interpreter of
    the op semantics and against the unfused k_pair / k_diag path."""
    import emulate
    n, n_params = 14, 6
    rng = np.random.default_rng(1000 + tile_bits)
    th = rng.uniform(-1.0, 1.0, n_params)
    circ = Circuit(n, n_params)
    for _ in range(14):
        w = [int(v) for v in rng.choice(n, size=3, replace=False)]
        for _ in range(int(rng.integers(1, 6))):
            kind = int(rng.integers(0, 8))
            a, b, c = (w[int(i)] for i in rng.permutation(3))
            ang = float(rng.uniform(-2.0, 2.0))
            if kind == 0:
                circ.ry(0.0, a, param=int(rng.integers(n_params)))
            elif kind == 1:
                circ.rx(ang, a)
            elif kind == 2:
                circ.cnot(a, b)
            elif kind == 3:
                circ.single_excitation(ang, a, b)
            elif kind == 4:
                circ.fermionic_single_excitation(ang, a, b)
            elif kind == 5:
                circ.rz(ang, a)
            elif kind == 6:
                # Pauli string with X/Y letters on the triple and a Z tail elsewhere
                x = circ._bit(a) | circ._bit(b) | (circ._bit(c) if rng.integers(2) else 0)
                z = int(rng.integers(1 << n))
                circ.pauli_rotation(x, z, 0.5, param=int(rng.integers(n_params)))
            else:
                circ.pauli_x(a)
    psi0 = rand_state(n, 77 + tile_bits)
    want = emulate.run_circuit(circ, psi0.copy(), th)
    fused = circ.compile(ctx, tile_bits=tile_bits, low_bits=0)
    assert fused.n_tiles > 0
    st = State.from_numpy(ctx, psi0)
    fused.run(st, th)
    got = st.numpy()
    assert np.abs(got - want).max() < AMP_TOL
    plain = circ.compile(ctx, fuse=False)
    st2 = State.from_numpy(ctx, psi0)
    plain.run(st2, th)
    assert np.abs(st2.numpy() - got).max() < AMP_TOL
    fused.run(st, th, dagger=True)
    assert np.abs(st.numpy() - psi0).max() < AMP_TOL


def test_adjoint_gradient_vs_finite_difference_22_qubits(ctx):
    """The n >= 20/22 branches inside fh_program_evaluate (K2 with 8 outputs per thread + diagonal factor tables,
    standalone diagonal kernel through phase tables, fused adjoint tiles on 12-bit tiles) against central differences
    of the same program's energy, plus norm and energy consistency with the immediate-mode K2."""
    from operators.tools import get_interacting_term
    nx, ny, u, n = 11, 1, 4.0, 22
    h_op = fermi_hubbard(nx, ny, 1.0, u)
    h_tab = PauliTable.from_operator(h_op, n)
    coulomb = GeneratorPlan(jordan_wigner(get_interacting_term(h_op)), n)
    circ = Circuit(n, 4)
    for q in range(0, n, 3):
        circ.ry(0.2 + 0.05 * q, q)
    circ.generator(coulomb, param=0)
    for a, b, p in ((0, 2, 1), (7, 9, 2), (12, 20, 1), (1, 3, 3), (15, 17, 2)):
        circ.fermionic_single_excitation(0.3, a, b)
        circ.pauli_rotation((1 << (n - 1 - a)) | (1 << (n - 1 - b)), 0, 0.5, param=p)
    circ.generator(coulomb, param=3)
    prog = circ.compile(ctx)
    dtab = DeviceTable(ctx, h_tab)
    thetas = np.array([0.31, -0.22, 0.57, 0.13])
    basis = sum(1 << (n - 1 - q) for q in (0, 1, 4, 5, 8, 9, 12, 13, 16, 17, 20))
    out = State(ctx, n)
    res = prog.evaluate(basis, thetas, [dtab], grads=True, state_out=out)
    assert abs(out.norm2() - 1.0) < 1e-12
    assert abs(dtab.expval(out) - res["expvals"][0]) < 1e-10
    hstep = 1e-5
    for j in range(4):
        tp, tm = thetas.copy(), thetas.copy()
        tp[j] += hstep
        tm[j] -= hstep
        fd = (prog.evaluate(basis, tp, [dtab])["expvals"][0] - prog.evaluate(basis, tm, [dtab])["expvals"][0]) / (2 * hstep)
        assert abs(res["grads"][j] - fd) < 2e-8, (j, res["grads"][j], fd)
    unfused = circ.compile(ctx, fuse=False)
    res2 = unfused.evaluate(basis, thetas, [dtab], grads=True)
    assert abs(res2["expvals"][0] - res["expvals"][0]) < 1e-10
    assert np.abs(res2["grads"] - res["grads"]).max() < 1e-9


def test_known_answers_3x3_first_screening(ctx):
    """E_HF = -5/3 and exactly 52 operators at |g| = 4/3 through the CUDA path."""
    n, h_tab, pool_ops, dec, diag, _, _ = lattice(3, 3, 6.0)
    plans = [GeneratorPlan(g, n) for g in pool_ops]
    circ = Circuit(n, 0)
    circ.marker("ansatz_end")
    circ.basis_change_separable(3, 3)
    prog = circ.compile(ctx)
    occ = [0, 2, 4, 6, 12, 1, 3, 5, 7]
    basis = sum(1 << (n - 1 - q) for q in occ)
    res = prog.evaluate(basis, [], [DeviceTable(ctx, h_tab)], pool=DevicePool(ctx, plans, n), pool_pos=0)
    assert abs(res["expvals"][0] + 5.0 / 3.0) < E_TOL
    g = np.abs(res["pool"])
    assert (g > 1e-9).sum() == 52 and np.allclose(g[g > 1e-9], 4.0 / 3.0, atol=G_TOL)


def test_evaluate_graph_equals_item_by_item_run_18_qubits(ctx):
    """fh_program_evaluate launches its tile kernels with programmatic dependent launch (the next kernel's prologue
    overlaps the previous kernel's tail, griddepcontrol.wait before the first state access); fh_program_run launches
    the same items one by one without it.  Same circuit, same parameters: the final states must agree to rounding
    (a missing dependency would show up as O(1) differences), and replays of the graph must be bit-identical."""
    n, h_tab, pool_ops, _, _, _, _ = lattice(3, 3, 6.0)
    plans = [GeneratorPlan(g, n) for g in pool_ops]
    rng = np.random.default_rng(2718)
    picks = [int(k) for k in rng.choice(len(plans), size=40, replace=False)]
    th = rng.uniform(-0.3, 0.3, len(picks))
    circ = Circuit(n, len(picks))
    for p, k in enumerate(picks):
        circ.generator(plans[k], param=p)
    circ.basis_change_separable(3, 3)
    prog = circ.compile(ctx)
    assert prog.n_tiles >= 8
    occ = [0, 2, 4, 6, 12, 1, 3, 5, 7]
    basis = sum(1 << (n - 1 - q) for q in occ)
    dtab = DeviceTable(ctx, h_tab)
    out = State(ctx, n)
    res = prog.evaluate(basis, th, [dtab], state_out=out)
    via_graph = out.numpy().copy()
    st = State(ctx, n)
    st.set_basis(basis)
    prog.run(st, th)
    via_run = st.numpy()
    assert abs(np.vdot(via_run, via_run).real - 1.0) < 1e-12
    assert np.abs(via_graph - via_run).max() < 1e-14
    for _ in range(5):
        again = prog.evaluate(basis, th, [dtab], state_out=out)
        assert again["expvals"][0] == res["expvals"][0]
        assert np.array_equal(out.numpy(), via_graph)


def test_lanczos_vs_sector_ed(ctx):
    for (nx, ny, u, up, dn) in [(2, 2, 4.0, 2, 2), (2, 3, 4.0, 3, 3)]:
        n, h_tab, _, _, _, o_h, _ = lattice(nx, ny, u)
        want, _, _ = ed.ground_state(o_h, n, up + dn, up, dn, k=3)
        dtab = DeviceTable(ctx, h_tab)
        vals, vecs, iters = lanczos(dtab, k=3, n_up=up, n_dn=dn, tol=1e-11)
        assert np.abs(np.sort(vals) - want).max() < 1e-9
        for v, e in zip(vecs, vals):
            hv = State(ctx, n)
            assert abs(dtab.apply(v, hv).real - e) < 1e-9
            assert abs(v.norm2() - 1) < 1e-10


def test_lanczos_3x3_degenerate_ground_level(ctx):
    n, h_tab, _, _, _, _, _ = lattice(3, 3, 6.0)
    vals, vecs, iters = lanczos(DeviceTable(ctx, h_tab), k=5, n_up=5, n_dn=4, tol=1e-10, max_iter=600)
    vals = np.sort(vals)
    assert np.abs(vals[:4] + 5.5623088363).max() < 2e-9
    assert abs(vals[4] + 4.7084214853) < 2e-9


def test_error_paths(ctx):
    with pytest.raises(ValueError):
        State(ctx, 0)
    st = State(ctx, 4)
    with pytest.raises(ValueError):
        st.set_basis(16)
    with pytest.raises(ValueError):
        st.apply_pair(0, 1, 0, 0, [0, 0, 1, 0, 1, 0, 0, 0])
    tab = PauliTable(3, [1], [0], [1.0])
    with pytest.raises(ValueError):
        DeviceTable(ctx, tab).apply(st)              # qubit-count mismatch


def _random_clustered_circuit(n, n_params, seed, groups=14):
    rng = np.random.default_rng(seed)
    circ = Circuit(n, n_params)
    for _ in range(groups):
        w = [int(v) for v in rng.choice(n, size=3, replace=False)]
        for _ in range(int(rng.integers(1, 6))):
            kind = int(rng.integers(0, 8))
            a, b, c = (w[int(i)] for i in rng.permutation(3))
            ang = float(rng.uniform(-2.0, 2.0))
            if kind == 0:
                circ.ry(0.0, a, param=int(rng.integers(n_params)))
            elif kind == 1:
                circ.rx(ang, a)
            elif kind == 2:
                circ.cnot(a, b)
            elif kind == 3:
                circ.single_excitation(ang, a, b)
            elif kind == 4:
                circ.fermionic_single_excitation(ang, a, b)
            elif kind == 5:
                circ.rz(ang, a)
            elif kind == 6:
                x = circ._bit(a) | circ._bit(b) | (circ._bit(c) if rng.integers(2) else 0)
                z = int(rng.integers(1 << n))
                circ.pauli_rotation(x, z, 0.5, param=int(rng.integers(n_params)))
            else:
                circ.pauli_x(a)
    return circ


@pytest.mark.parametrize("n,tile_bits,low_bits", [(14, 8, 1), (14, 10, 3), (16, 11, 2), (18, 11, 1), (18, 12, 4),
                                                   (20, 13, 3), (22, 11, 3), (22, 12, 1)])
def test_tile_tma_path_equals_register_staged_path_and_interpreter(ctx, n, tile_bits, low_bits, monkeypatch):
    """k_tile_tma (TMA tile gather / scatter, hardware 128-byte swizzle, one or two tile buffers per CTA) against the
    register-staged k_tile (FHSIM_TILE_LDG=1) on the same compiled program, forward and dagger, and against the numpy
    interpreter of the op semantics where that finishes in seconds.  n = 22 runs persistent CTAs (2 048 / 1 024 tiles)."""
    import emulate
    n_params = 6
    th = np.random.default_rng(5 + n).uniform(-1.0, 1.0, n_params)
    circ = _random_clustered_circuit(n, n_params, 4000 + 31 * n + tile_bits, groups=18)
    psi0 = rand_state(n, 300 + n)
    prog = circ.compile(ctx, tile_bits=tile_bits, low_bits=low_bits)
    assert prog.n_tiles > 0
    monkeypatch.delenv("FHSIM_TILE_LDG", raising=False)
    st = State.from_numpy(ctx, psi0)
    prog.run(st, th)
    got = st.numpy()
    monkeypatch.setenv("FHSIM_TILE_LDG", "1")
    st2 = State.from_numpy(ctx, psi0)
    prog.run(st2, th)
    ref = st2.numpy()
    monkeypatch.delenv("FHSIM_TILE_LDG", raising=False)
    assert np.abs(got - ref).max() < 1e-13
    assert abs(np.vdot(got, got).real - 1.0) < 1e-12
    if n <= 18:
        want = emulate.run_circuit(circ, psi0.copy(), th)
        assert np.abs(got - want).max() < AMP_TOL
    prog.run(st, th, dagger=True)
    assert np.abs(st.numpy() - psi0).max() < AMP_TOL


def test_tile_tma_coulomb_layers_with_straddling_terms(ctx):
    """Diagonal ops inside TMA tiles: ZZ terms whose two bits sit in different halves of the tile (evaluated from the
    tile-local z-mask) and terms with bits outside the tile, interleaved with hopping rotations -- vs the interpreter."""
    import emulate
    n = 16
    rng = np.random.default_rng(99)
    circ = Circuit(n, 3)
    for rep in range(3):
        zs = [int((1 << int(a)) | (1 << int(b))) for a, b in (rng.choice(n, size=2, replace=False) for _ in range(20))]
        zs += [1 << int(q) for q in rng.choice(n, size=6, replace=False)]
        circ.ops.append(DiagOpSpec(zs, [float(v) for v in rng.uniform(-1, 1, len(zs))], rep, [(0, z) for z in zs]))
        for _ in range(4):
            a, b = (int(v) for v in rng.choice(n, size=2, replace=False))
            circ.fermionic_single_excitation(float(rng.uniform(-2, 2)), a, b)
    th = rng.uniform(-1, 1, 3)
    psi0 = rand_state(n, 5)
    want = emulate.run_circuit(circ, psi0.copy(), th)
    for tb, lb in ((10, 3), (12, 1), (9, 4)):
        prog = circ.compile(ctx, tile_bits=tb, low_bits=lb)
        st = State.from_numpy(ctx, psi0)
        prog.run(st, th)
        assert np.abs(st.numpy() - want).max() < AMP_TOL


@pytest.mark.parametrize("n,tile_bits", [(14, 9), (18, 11), (20, 12)])
def test_tile_chain_equals_one_launch_per_run(ctx, n, tile_bits, monkeypatch):
    """k_tile_chain (all tile runs of a range in one cooperative launch, run-boundary counters instead of kernel
    boundaries) vs one k_tile_tma launch per run (FHSIM_NO_CHAIN=1) vs the interpreter, forward and dagger."""
    import emulate
    n_params = 6
    th = np.random.default_rng(50 + n).uniform(-1.0, 1.0, n_params)
    circ = _random_clustered_circuit(n, n_params, 9000 + n, groups=24)
    psi0 = rand_state(n, 900 + n)
    prog = circ.compile(ctx, tile_bits=tile_bits, low_bits=3)
    assert prog.n_tiles >= 3
    monkeypatch.setenv("FHSIM_CHAIN", "1")                 # the chain is opt-in (not faster at 18 qubits)
    st = State.from_numpy(ctx, psi0)
    prog.run(st, th)
    got = st.numpy()
    monkeypatch.delenv("FHSIM_CHAIN", raising=False)
    st2 = State.from_numpy(ctx, psi0)
    prog.run(st2, th)
    ref = st2.numpy()
    monkeypatch.setenv("FHSIM_CHAIN", "1")
    assert np.array_equal(got, ref)                      # same kernels' arithmetic, only the launch structure differs
    if n <= 18:
        assert np.abs(got - emulate.run_circuit(circ, psi0.copy(), th)).max() < AMP_TOL
    prog.run(st, th, dagger=True)
    assert np.abs(st.numpy() - psi0).max() < AMP_TOL
    for _ in range(3):                                   # counters are reset per launch: repeated runs stay exact
        prog.run(st, th)
        prog.run(st, th, dagger=True)
    assert np.abs(st.numpy() - psi0).max() < 1e-11


def test_evaluate_with_chains_basis_synthesis_and_checkpoint_store(ctx, monkeypatch):
    """fh_program_evaluate at 3x3: the captured graph with chains (first run synthesises |HF>, the run before the pool
    position stores psi to the checkpoint as well) vs a graph captured with FHSIM_NO_CHAIN=1 (set-basis kernel, one
    launch per run, device-to-device checkpoint copy): energies, gradients, pool gradients, overlaps and state_out."""
    n, h_tab, pool_ops, dec, diag, o_h, o_pool = lattice(3, 3, 6.0)
    plans = [GeneratorPlan(g, n) for g in pool_ops]
    picks = [3, 17, 40, 77, 120, 200, 250, 300]
    th = np.random.default_rng(4).uniform(-0.3, 0.3, len(picks))
    occ = [0, 2, 4, 6, 12, 1, 3, 5, 7]
    basis = sum(1 << (n - 1 - q) for q in occ)
    dtab = DeviceTable(ctx, h_tab)
    dpool = DevicePool(ctx, plans, n)
    target = State.from_numpy(ctx, rand_state(n, 11))

    def build():
        circ = Circuit(n, len(picks))
        for j, k in enumerate(picks):
            circ.generator(plans[k], param=j)
        circ.marker("ansatz_end")
        circ.basis_change_separable(3, 3)
        return circ.compile(ctx)

    results = []
    for no_chain in (False, True):
        if no_chain:
            monkeypatch.delenv("FHSIM_CHAIN", raising=False)
        else:
            monkeypatch.setenv("FHSIM_CHAIN", "1")
        prog = build()
        out = State(ctx, n)
        r1 = prog.evaluate(basis, th, [dtab], grads=True, pool=dpool, pool_pos=prog.markers["ansatz_end"], targets=[target],
                           state_out=out)
        r2 = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=prog.markers["ansatz_end"])       # screening-only graph
        r3 = prog.evaluate(basis, th, [dtab])                                                        # energy-only graph
        r1b = prog.evaluate(basis, th, [dtab], grads=True, pool=dpool, pool_pos=prog.markers["ansatz_end"], targets=[target],
                            state_out=out)                                                           # re-captured, replayed
        assert np.array_equal(r1["pool"], r1b["pool"]) and np.array_equal(r1["grads"], r1b["grads"])
        results.append((r1, r2, r3, out.numpy(), prog.last_stats()[1]))
    monkeypatch.delenv("FHSIM_CHAIN", raising=False)
    (a1, a2, a3, sa, la), (b1, b2, b3, sb, lb) = results
    assert la < lb                                        # fewer launches with chains
    assert np.array_equal(sa, sb)
    for a, b in ((a1, b1), (a2, b2), (a3, b3)):
        assert np.array_equal(a["expvals"], b["expvals"])
        if a["pool"] is not None:
            assert np.array_equal(a["pool"], b["pool"])
        if a["grads"] is not None:
            assert np.array_equal(a["grads"], b["grads"])
    assert np.array_equal(a1["overlaps"], b1["overlaps"])
    # and against the oracle
    psi = sv.adapt_state(n, occ, [o_pool[k] for k in picks], th)
    g_or, e_or, _ = sv.pool_gradients(psi, o_h, o_pool, diag, dec, n)
    assert abs(a1["expvals"][0] - e_or) < E_TOL and np.abs(a1["pool"] - g_or).max() < G_TOL


@pytest.mark.parametrize("lat,u,up,dn,k", [((2, 2), 4.0, 2, 2, 3), ((2, 3), 4.0, 3, 3, 2), ((3, 3), 6.0, 5, 4, 4), ((2, 4), 2.0, 4, 4, 1)])
def test_sector_compressed_lanczos_vs_sector_ed(ctx, lat, u, up, dn, k):
    """fh_lanczos_sector (compressed vectors, device-resident iteration) vs scipy on the sector matrix: energies 1e-9,
    eigenvectors span the oracle's eigenspaces, compressed and scattered vectors agree, nothing leaves the sector."""
    n, h_tab, _, _, _, o_h, _ = lattice(*lat, u)
    tab = DeviceTable(ctx, h_tab)
    evals, vecs, comp, info = lanczos_sector(tab, up, dn, k=k, tol=1e-11, max_iter=3000, seed=7, want_vectors=True,
                                              want_compressed=True)
    want, wvecs, idx = ed.ground_state(o_h, n, up + dn, up, dn, k=max(k + 2, 6))
    assert np.abs(evals - want[:k]).max() < 1e-9
    assert info["sector_dim"] == len(idx) and info["matvecs"] >= info["iterations"] - 8 * k
    sidx = sector_indices(n, up, dn)
    assert sorted(int(v) for v in sidx) == sorted(int(v) for v in idx)
    level = [w for w, e in zip(wvecs, want) if abs(e - want[0]) < 1e-7]      # the oracle's (possibly degenerate) ground level
    # ARPACK's vectors of a degenerate level are not orthonormal (Gram error ~4e-3 on 3x3; the reference Gram-Schmidts them,
    # exact_diagonalization.py:210-222): orthonormalise before projecting
    level = list(np.linalg.qr(np.array(level).T)[0].T)
    for e in range(k):
        full = vecs[e].numpy()
        assert abs(np.linalg.norm(full) - 1.0) < 1e-10
        outside = np.ones(1 << n, bool)
        outside[idx] = False
        assert np.abs(full[outside]).max() == 0.0
        assert np.abs(full[sidx.astype(np.int64)] - comp[e]).max() == 0.0
        resid = sv.apply_table(full, o_h, n) - evals[e] * full
        assert np.linalg.norm(resid) < 1e-7
        if abs(evals[e] - want[0]) < 1e-7:
            proj = sum(abs(np.vdot(w, full)) ** 2 for w in level)
            assert abs(proj - 1.0) < 1e-7
    gram = np.array([[np.vdot(a.numpy(), b.numpy()) for b in vecs] for a in vecs])
    assert np.abs(gram - np.eye(k)).max() < 1e-9
    # fh_lanczos takes the compressed route by itself for sector requests and gives the same energies
    ev2, v2, it2 = lanczos(tab, k=k, n_up=up, n_dn=dn, tol=1e-11, max_iter=3000, seed=7)
    assert np.abs(ev2 - evals).max() < 1e-12


def test_sector_lanczos_rejects_tables_that_leave_the_sector(ctx):
    n = 8
    op = QubitOperator("X0 X1", 0.5) + QubitOperator("Z0 Z3", 0.25)         # X0 X1 creates / destroys an up-down pair
    tab = DeviceTable(ctx, PauliTable.from_operator(op, n))
    with pytest.raises(ValueError, match="conserve"):
        lanczos_sector(tab, 2, 2)


def test_device_dressing_is_bit_identical_to_the_host_restatement(ctx):
    """fh_ptable_dress vs PauliTable.dressed (reference iqcc_hubbard.py:184-189 on packed tables): same strings in the same
    order, coefficients bit for bit, over a chain of entanglers that grows the table from 100 to thousands of terms, plus
    random complex tables (duplicates of generated strings, cancellations, tau = 0)."""
    from fhsim.backend import DevicePauliTable
    n = 18
    host = PauliTable.from_operator(fermi_hubbard(3, 3, 1.0, 6.0), n)
    dev = DevicePauliTable(ctx, host)
    rng = np.random.default_rng(12)
    for step in range(7):
        xp = int(rng.integers(1, 1 << n))
        zp = int(rng.integers(0, 1 << n))
        tau = float(rng.uniform(-2, 2)) if step != 3 else 0.0
        host = host.dressed(xp, zp, tau)
        dev.dress(xp, zp, tau)
        got = dev.to_host()
        assert len(got) == len(host) == len(dev)
        assert np.array_equal(got.x, host.x) and np.array_equal(got.z, host.z)
        assert np.array_equal(got.coeff.real, host.coeff.real) and np.array_equal(got.coeff.imag, host.coeff.imag)
    assert len(host) > 1000
    dev.close()
    # random complex tables on few qubits: generated strings hit existing ones often
    for seed in range(6):
        r = np.random.default_rng(100 + seed)
        m = 7
        t = PauliTable(m, r.integers(0, 1 << m, 60), r.integers(0, 1 << m, 60), r.normal(size=60) + 1j * r.normal(size=60))
        d = DevicePauliTable(ctx, t)
        for _ in range(4):
            xp, zp, tau = int(r.integers(1, 1 << m)), int(r.integers(0, 1 << m)), float(r.uniform(-3, 3))
            t = t.dressed(xp, zp, tau)
            d.dress(xp, zp, tau)
        g = d.to_host()
        assert np.array_equal(g.x, t.x) and np.array_equal(g.z, t.z) and np.array_equal(g.coeff, t.coeff)
        d.close()


def test_register_fused_runs_equal_the_op_by_op_tile_kernel(ctx, monkeypatch):
    """FHSIM_RUNS=1: consecutive Givens-like ops inside three tile-local bits are applied to 8-amplitude groups held in
    registers (one shared-memory round trip per run).  The separable W of 3x3 (twelve 3-site Fourier transforms) and random
    Givens circuits, forward and dagger, against the op-by-op records of the same kernels."""
    n = 18
    th = []
    monkeypatch.delenv("FHSIM_RUNS", raising=False)
    circ = Circuit(n, 0)
    circ.basis_change_separable(3, 3)
    rng = np.random.default_rng(3)
    for _ in range(30):
        a = int(rng.integers(n - 2))
        circ.fermionic_single_excitation(float(rng.uniform(-2, 2)), a, a + int(rng.integers(1, 3)))
    psi0 = rand_state(n, 41)
    plain = circ.compile(ctx, tile_bits=11, low_bits=3)
    st = State.from_numpy(ctx, psi0)
    plain.run(st, th)
    ref = st.numpy()
    monkeypatch.setenv("FHSIM_RUNS", "1")
    fused = circ.compile(ctx, tile_bits=11, low_bits=3)        # the runs are detected when the program is finalised
    st2 = State.from_numpy(ctx, psi0)
    fused.run(st2, th)
    assert np.abs(st2.numpy() - ref).max() < 1e-13
    fused.run(st2, th, dagger=True)
    assert np.abs(st2.numpy() - psi0).max() < AMP_TOL


def test_apply_table_window_traversal_order_is_a_pure_reordering(ctx, monkeypatch):
    """K2 above L2 size walks the index space in windows (top bits fast, Gray-coded slow bits, kernels.cu): forced on at 22
    qubits it must give the same H|psi> bit for bit (every output is computed by the same thread arithmetic) and the same
    energy to rounding (the per-CTA partial sums collect different blocks)."""
    n, h_tab, _, _, _, o_h, _ = lattice(3, 3, 6.0)
    n = 22
    h22 = PauliTable.from_operator(fermi_hubbard(1, 11, 1.0, 4.0), n)
    psi = rand_state(n, 77)
    tab = DeviceTable(ctx, h22)
    st, out0, out1 = State.from_numpy(ctx, psi), State(ctx, n), State(ctx, n)
    monkeypatch.setenv("FHSIM_K2_SLOW_BITS", "0")
    e0 = tab.apply(st, out0)
    monkeypatch.setenv("FHSIM_K2_SLOW_BITS", "4")
    e1 = tab.apply(st, out1)
    monkeypatch.delenv("FHSIM_K2_SLOW_BITS", raising=False)
    assert np.array_equal(out0.numpy(), out1.numpy())
    assert abs(e0 - e1) < 1e-11
    want = sv.apply_table(psi, h22.as_dict(), n)
    assert np.abs(out1.numpy() - want).max() < AMP_TOL


def _coverable_random_table(n, rng, n_sets, complex_coeffs):
    """A random table whose x-masks lie inside `n_sets` sets of 12 index bits (so K2 takes the tile passes), with several
    z-masks per x-mask (multi-class groups), z bits inside and outside the sets, and diagonal terms of every chunk shape."""
    table = {}
    sets = {1: [sorted(int(b) for b in rng.choice(n, size=12, replace=False))],
            2: [list(range(1, n, 2)), list(range(0, n, 2))],                    # the two spin species of a Hubbard model
            3: [list(range(0, 7)), list(range(7, 14)), list(range(14, 21))]}[n_sets]
    for bits in sets:
        for _ in range(7):
            k = int(rng.integers(1, 5))
            x = sum(1 << b for b in rng.choice(bits, size=k, replace=False))
            for _ in range(int(rng.integers(1, 4))):
                z = int(rng.integers(0, 1 << n))
                table[(int(x), z)] = complex(rng.normal(), rng.normal() if complex_coeffs else 0.0)
    for z in (0, 1 << 3, (1 << 11) | (1 << 12), (1 << 23) | (1 << 24) if n > 24 else (1 << 5) | (1 << 20), (1 << 2) | (1 << 13) | (1 << (n - 1))):
        table[(0, int(z))] = complex(rng.normal(), 0.0)
    if not complex_coeffs:
        # a real-coefficient table is "all real" for the kernel only if every i^k is real: keep strings with an even number of Y
        table = {(x, z): v for (x, z), v in table.items() if bin(x & z).count("1") % 2 == 0}
    return table


@pytest.mark.parametrize("n_sets,complex_coeffs,seed", [(1, False, 1), (2, True, 2), (3, True, 3), (2, False, 4)])
def test_apply_table_tile_passes_vs_oracle_and_gather_kernel(ctx, n_sets, complex_coeffs, seed, monkeypatch):
    """K2 in shared-memory tile passes (n >= 22, csrc/table_tile.cu): out = H psi, <psi|H|psi> alone, and out += H psi against
    the oracle and against the gather kernel on the same table."""
    from fhsim import _cabi
    monkeypatch.setenv("FHSIM_K2_TILE", "2")            # general tables too (the default takes only uniform real passes)
    n = 22
    rng = np.random.default_rng(seed)
    table = _coverable_random_table(n, rng, n_sets, complex_coeffs)
    keys = list(table)
    dt = DeviceTable(ctx, PauliTable(n, [k[0] for k in keys], [k[1] for k in keys], [table[k] for k in keys]))
    assert 1 <= dt.tile_passes() <= 3
    psi = rand_state(n, 100 + seed)
    st, out = State.from_numpy(ctx, psi), State(ctx, n)
    want = sv.apply_table(psi, table, n)
    e = dt.apply(st, out)
    got = out.numpy()
    assert np.abs(got - want).max() < 1e-12
    assert abs(e - np.vdot(psi, want)) < E_TOL
    assert abs(dt.apply(st) - np.vdot(psi, want)) < E_TOL                   # expectation only
    re, im = _cabi.C.c_double(), _cabi.C.c_double()
    _cabi.check(_cabi.lib().fh_apply_table_accumulate(dt._h, st._h, out._h, _cabi.C.byref(re), _cabi.C.byref(im)))
    assert np.abs(out.numpy() - 2 * want).max() < 1e-12                      # out += H psi
    assert abs(complex(re.value, im.value) - np.vdot(psi, want)) < E_TOL
    monkeypatch.setenv("FHSIM_K2_GATHER", "1")
    e_g = dt.apply(st, out)
    assert np.abs(out.numpy() - got).max() < 1e-12 and abs(e_g - e) < 1e-11
    assert dt.apply(st, out) == e_g                                           # bit-reproducible
    monkeypatch.delenv("FHSIM_K2_GATHER")
    assert dt.apply(st, out) == e and np.array_equal(out.numpy(), got)


def test_apply_table_tile_passes_hubbard_3x4(ctx):
    """The 3x4 Hubbard Hamiltonian (24 qubits) is two passes (up-orbital bits, down-orbital bits); properties that do not need
    the oracle at this size: Hermiticity <a|H b> = conj(<b|H a>), energy = <psi|out>, HF energy of a basis state."""
    n = 24
    dt = DeviceTable(ctx, PauliTable.from_operator(fermi_hubbard(3, 4, 1.0, 4.0), n))
    assert dt.tile_passes() == 2
    a, b = rand_state(n, 1), rand_state(n, 2)
    sa, sb, out = State.from_numpy(ctx, a), State.from_numpy(ctx, b), State(ctx, n)
    ea = dt.apply(sa, out)
    ha = out.numpy()
    assert abs(ea - np.vdot(a, ha)) < 1e-10 and abs(ea.imag) < 1e-12
    dt.apply(sb, out)
    hb = out.numpy()
    assert abs(np.vdot(a, hb) - np.conj(np.vdot(b, ha))) < 1e-10
    basis = np.zeros(1 << n, complex)
    idx = sum(1 << (n - 1 - q) for q in (0, 1, 4, 7, 10, 13))      # sites 0 (doubly occupied), 2 up, 3 dn, 5 up, 6 dn
    basis[idx] = 1.0
    assert abs(dt.apply(State.from_numpy(ctx, basis)) - 4.0) < 1e-12     # U n_up n_dn on one doubly occupied site
