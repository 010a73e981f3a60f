"""Sector-resident evaluation (csrc/sector_eval.cu: the whole screening on the (N_up, N_dn)-compressed state inside one
thread-block cluster) against the oracle and against the full-space path of the same program, through the C-ABI.
Reference semantics: ADAPT.circuit + select_operator, models/adapt_vqe.py:297-361."""
import numpy as np
import pytest

from fhsim.backend import Context, DevicePool, DeviceTable
from fhsim.circuit import Circuit, DiagOpSpec
from fhsim.symbolic import fermi_hubbard, givens_decomposition_square, jordan_wigner
from fhsim.tables import GeneratorPlan, PauliTable
from operators.fourier import fourier_transform_matrix
from operators.pool import hubbard_interaction_pool_simplified
from oracle import pauli, statevector as sv

pytestmark = pytest.mark.gpu

E_TOL, G_TOL = 1e-10, 1e-9


@pytest.fixture(scope="module")
def ctx():
    return Context(0)


def lattice(nx, ny, u):
    n = 2 * nx * ny
    h_tab = PauliTable.from_operator(fermi_hubbard(nx, ny, 1.0, u), n)
    pool_ops = [jordan_wigner(g) for g in hubbard_interaction_pool_simplified(nx, ny)]
    dec, diag = givens_decomposition_square(fourier_transform_matrix(nx, ny))
    o_h = pauli.compress(pauli.jw_table(pauli.hubbard_fermion_terms(nx, ny, 1.0, u), n))
    o_pool = [pauli.jw_table(op, n) for op in pauli.pool_fermion_terms(nx, ny)]
    return n, h_tab, pool_ops, dec, diag, o_h, o_pool


@pytest.mark.parametrize("lat,u,up,dn", [((2, 2), 4.0, 2, 2), ((2, 3), 4.0, 3, 3), ((2, 3), 4.0, 4, 2), ((3, 3), 6.0, 5, 4)])
def test_sector_screening_vs_oracle(ctx, lat, u, up, dn, monkeypatch):
    """Energy and all pool gradients of an ADAPT ansatz state: oracle (closed form, float64) vs the cluster kernel.  W is the
    separable network (same unitary as the reference's, tests/test_circuit_host.py): the reference's own Givens network
    rotates between the up and down orbital of a site, so its intermediate states leave the (N_up, N_dn) sector (next test)."""
    monkeypatch.setenv("FHSIM_SECTOR", "1")
    n, h_tab, pool_ops, dec, diag, o_h, o_pool = lattice(*lat, u)
    rng = np.random.default_rng(7)
    occ_up, occ_dn, _ = pauli.k_space_occupation(*lat, 1.0, up, dn)
    picks = list(rng.choice(len(pool_ops), size=6, replace=False))
    th = rng.uniform(-0.4, 0.4, len(picks))
    plans = [GeneratorPlan(g, n) for g in pool_ops]
    circ = Circuit(n, len(picks))
    for p, k in enumerate(picks):
        circ.generator(plans[k], param=p)
    circ.marker("ansatz_end")
    circ.basis_change_separable(*lat)
    prog = circ.compile(ctx)
    dtab = DeviceTable(ctx, h_tab)
    dpool = DevicePool(ctx, plans, n)
    basis = sum(1 << (n - 1 - q) for q in occ_up + occ_dn)
    m = prog.markers["ansatz_end"]
    res = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=m)
    info = prog.sector_info()
    assert info["active"] and info["cluster"] >= 2
    assert prog.last_stats()[1] == 2                  # one cluster kernel + one pool kernel
    gens = [o_pool[k] for k in picks]
    psi_k = sv.adapt_state(n, occ_up + occ_dn, gens, th)
    pg_want, e_want, _ = sv.pool_gradients(psi_k, o_h, o_pool, diag, dec, n)
    assert abs(res["expvals"][0] - e_want) < E_TOL
    assert np.abs(res["pool"] - pg_want).max() < G_TOL
    # energy only, pool at the very start / very end, a sub-range of the pool; graph replay with new parameters
    e_only = prog.evaluate(basis, th, [dtab])
    assert prog.sector_info()["active"] and abs(e_only["expvals"][0] - e_want) < E_TOL
    monkeypatch.setenv("FHSIM_NO_SECTOR", "1")
    full0 = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=0)
    full_end = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=prog.n_items)
    assert not prog.sector_info()["active"]
    monkeypatch.delenv("FHSIM_NO_SECTOR")
    sec0 = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=0)
    sec_end = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=prog.n_items)
    assert prog.sector_info()["active"]
    assert np.abs(sec0["pool"] - full0["pool"]).max() < 1e-12 and np.abs(sec_end["pool"] - full_end["pool"]).max() < 1e-12
    lo, cnt = 3, max(1, dpool.n_out // 3)
    sub = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=m, pool_range=(lo, cnt))
    assert np.abs(sub["pool"] - pg_want[lo:lo + cnt]).max() < G_TOL
    th2 = th + 0.05
    res2 = prog.evaluate(basis, th2, [dtab], pool=dpool, pool_pos=m)
    pg2, e2, _ = sv.pool_gradients(sv.adapt_state(n, occ_up + occ_dn, gens, th2), o_h, o_pool, diag, dec, n)
    assert abs(res2["expvals"][0] - e2) < E_TOL and np.abs(res2["pool"] - pg2).max() < G_TOL
    assert np.array_equal(prog.evaluate(basis, th2, [dtab], pool=dpool, pool_pos=m)["pool"], res2["pool"])   # bit-reproducible


@pytest.mark.parametrize("seed,up,dn", [(1, 3, 2), (2, 2, 4), (3, 1, 1), (4, 5, 3)])
def test_sector_random_number_conserving_circuits(ctx, seed, up, dn, monkeypatch):
    """Random circuits of same-spin Givens rotations (with and without a Jordan-Wigner string), pool generators and diagonal
    layers with Z, same-species ZZ and up-down ZZ terms, the pool evaluated in the middle: cluster kernel vs full space."""
    nx, ny = 2, 3
    n, h_tab, pool_ops, dec, diag, o_h, o_pool = lattice(nx, ny, 3.0)
    rng = np.random.default_rng(seed)
    plans = [GeneratorPlan(g, n) for g in pool_ops]
    n_par = 10
    circ = Circuit(n, n_par)

    def random_ops(count):
        for _ in range(count):
            kind = int(rng.integers(0, 4))
            par = int(rng.integers(0, n_par))
            if kind == 0:
                s = int(rng.integers(0, 2))
                a, b = (int(v) for v in rng.choice(n // 2, size=2, replace=False))
                circ.fermionic_single_excitation(float(rng.uniform(-2, 2)), 2 * a + s, 2 * b + s)
            elif kind == 1:
                s = int(rng.integers(0, 2))
                a, b = (int(v) for v in rng.choice(n // 2, size=2, replace=False))
                circ.single_excitation(float(rng.uniform(-2, 2)), 2 * a + s, 2 * b + s)
            elif kind == 2:
                circ.generator(plans[int(rng.integers(0, len(plans)))], param=par)
            else:
                zs = [int((1 << int(a)) | (1 << int(b))) for a, b in (rng.choice(n, size=2, replace=False) for _ in range(8))]
                zs += [1 << int(q) for q in rng.choice(n, size=4, replace=False)]
                circ.ops.append(DiagOpSpec(zs, [float(v) for v in rng.uniform(-1, 1, len(zs))], par, [(0, z) for z in zs]))

    random_ops(14)
    circ.marker("mid")
    random_ops(14)
    prog = circ.compile(ctx)
    dtab = DeviceTable(ctx, h_tab)
    dpool = DevicePool(ctx, plans, n)
    occ = [2 * int(s) for s in rng.choice(n // 2, size=up, replace=False)] + \
          [2 * int(s) + 1 for s in rng.choice(n // 2, size=dn, replace=False)]
    basis = sum(1 << (n - 1 - q) for q in occ)
    th = rng.uniform(-1, 1, n_par)
    m = prog.markers["mid"]
    monkeypatch.setenv("FHSIM_NO_SECTOR", "1")
    full = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=m)
    assert not prog.sector_info()["active"]
    monkeypatch.delenv("FHSIM_NO_SECTOR")
    sec = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=m)          # 12 qubits: the default takes the sector path
    info = prog.sector_info()
    assert info["active"], info
    assert abs(sec["expvals"][0] - full["expvals"][0]) < 1e-12
    assert np.abs(sec["pool"] - full["pool"]).max() < 1e-12
    assert np.abs(full["pool"]).max() > 1e-3


def test_sector_path_is_not_taken_when_it_does_not_apply(ctx, monkeypatch):
    """Circuits that leave the sector (RX), gradient / state / overlap requests and FHSIM_NO_SECTOR use the full-space path;
    without FHSIM_SECTOR=1 so does a sector above the size limit (3x3: 15 876 amplitudes)."""
    n, h_tab, pool_ops, dec, diag, o_h, o_pool = lattice(2, 2, 4.0)
    plans = [GeneratorPlan(g, n) for g in pool_ops]
    dtab = DeviceTable(ctx, h_tab)
    basis = sum(1 << (n - 1 - q) for q in (0, 2, 1, 3))
    c1 = Circuit(n, 1)
    c1.generator(plans[0], param=0)
    c1.rx(0.3, 1)
    p1 = c1.compile(ctx)
    p1.evaluate(basis, [0.2], [dtab])
    assert not p1.sector_info()["active"]
    c2 = Circuit(n, 1)
    c2.generator(plans[0], param=0)
    p2 = c2.compile(ctx)
    a = p2.evaluate(basis, [0.2], [dtab])
    assert p2.sector_info()["active"]
    b = p2.evaluate(basis, [0.2], [dtab], grads=True)
    assert not p2.sector_info()["active"] and abs(a["expvals"][0] - b["expvals"][0]) < 1e-13
    monkeypatch.setenv("FHSIM_NO_SECTOR", "1")
    p2.evaluate(basis, [0.2], [dtab])
    assert not p2.sector_info()["active"]
    monkeypatch.delenv("FHSIM_NO_SECTOR")
    # the reference's W network (adjacent-wire Givens rotations: up <-> down of one site) conserves only N_up + N_dn
    c4 = Circuit(n, 1)
    c4.generator(plans[0], param=0)
    c4.basis_change(diag, list(reversed(dec)))
    p4 = c4.compile(ctx)
    e4 = p4.evaluate(basis, [0.2], [dtab])["expvals"][0]
    assert not p4.sector_info()["active"]
    c5 = Circuit(n, 1)
    c5.generator(plans[0], param=0)
    c5.basis_change_separable(2, 2)
    p5 = c5.compile(ctx)
    e5 = p5.evaluate(basis, [0.2], [dtab])["expvals"][0]
    assert p5.sector_info()["active"] and abs(e4 - e5) < 1e-12
    n3, h3, pool3, *_ = lattice(3, 3, 6.0)
    c3 = Circuit(n3, 1)
    c3.generator(GeneratorPlan(pool3[5], n3), param=0)
    p3 = c3.compile(ctx)
    occ_up, occ_dn, _ = pauli.k_space_occupation(3, 3, 1.0, 5, 4)
    b3 = sum(1 << (n3 - 1 - q) for q in occ_up + occ_dn)
    d3 = DeviceTable(ctx, h3)
    e_full = p3.evaluate(b3, [0.1], [d3])["expvals"][0]
    assert not p3.sector_info()["active"]
    monkeypatch.setenv("FHSIM_SECTOR", "1")
    e_sec = p3.evaluate(b3, [0.1], [d3])["expvals"][0]
    assert p3.sector_info()["active"] and p3.sector_info()["dim"] == 15876
    assert abs(e_sec - e_full) < 1e-12


@pytest.mark.parametrize("lat,u,up,dn", [((2, 3), 4.0, 3, 3), ((3, 3), 6.0, 5, 4)])
def test_pool_screening_on_sector_compressed_copies(ctx, lat, u, up, dn, monkeypatch):
    """Full-space circuit kernels + K3 on sector-compressed psi_s / lambda_s (the default at 3x3) against K3 in the full
    space, for a screening call, a training call that also screens (grads=True) and a sub-range of the pool."""
    monkeypatch.setenv("FHSIM_NO_SECTOR", "1")           # keep the cluster path out of the way at 2x3
    n, h_tab, pool_ops, dec, diag, o_h, o_pool = lattice(*lat, u)
    rng = np.random.default_rng(11)
    occ_up, occ_dn, _ = pauli.k_space_occupation(*lat, 1.0, up, dn)
    picks = list(rng.choice(len(pool_ops), size=5, replace=False))
    th = rng.uniform(-0.4, 0.4, len(picks))
    plans = [GeneratorPlan(g, n) for g in pool_ops]
    circ = Circuit(n, len(picks))
    for p, k in enumerate(picks):
        circ.generator(plans[k], param=p)
    circ.marker("ansatz_end")
    circ.basis_change_separable(*lat)
    prog = circ.compile(ctx)
    dtab = DeviceTable(ctx, h_tab)
    dpool = DevicePool(ctx, plans, n)
    basis = sum(1 << (n - 1 - q) for q in occ_up + occ_dn)
    m = prog.markers["ansatz_end"]
    a = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=m)
    info = prog.sector_info()
    assert info["pool_in_sector"] and not info["active"]
    ag = prog.evaluate(basis, th, [dtab], grads=True, pool=dpool, pool_pos=m)
    assert prog.sector_info()["pool_in_sector"]
    asub = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=m, pool_range=(2, 7))
    monkeypatch.setenv("FHSIM_NO_SECTOR_POOL", "1")
    b = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=m)
    assert not prog.sector_info()["pool_in_sector"]
    bg = prog.evaluate(basis, th, [dtab], grads=True, pool=dpool, pool_pos=m)
    assert np.abs(a["pool"] - b["pool"]).max() < 1e-12 and np.abs(b["pool"]).max() > 1e-3
    assert np.abs(ag["pool"] - bg["pool"]).max() < 1e-12 and np.abs(ag["grads"] - bg["grads"]).max() < 1e-13
    assert np.abs(asub["pool"] - b["pool"][2:9]).max() < 1e-12
    pg_want, e_want, _ = sv.pool_gradients(sv.adapt_state(n, occ_up + occ_dn, [o_pool[k] for k in picks], th), o_h, o_pool, diag, dec, n)
    assert np.abs(a["pool"] - pg_want).max() < G_TOL and abs(a["expvals"][0] - e_want) < E_TOL
    # the reference W network leaves the sector op by op: K3 stays in the full space
    c2 = Circuit(n, len(picks))
    for p, k in enumerate(picks):
        c2.generator(plans[k], param=p)
    c2.marker("ansatz_end")
    c2.basis_change(diag, list(reversed(dec)))
    monkeypatch.delenv("FHSIM_NO_SECTOR_POOL")
    p2 = c2.compile(ctx)
    r2 = p2.evaluate(basis, th, [dtab], pool=dpool, pool_pos=p2.markers["ansatz_end"])
    assert not p2.sector_info()["pool_in_sector"]
    assert np.abs(r2["pool"] - pg_want).max() < G_TOL


@pytest.mark.parametrize("seed", [1, 2])
def test_sector_screening_with_permuted_species_bits(ctx, seed):
    """fh_pool_gradients_sector_masks: the form a slab of a sharded state needs.  The index bits of a 2x3 ADAPT state are
    permuted at random (so up / down orbitals are no longer the odd / even bits), the pool is lowered to that layout, and
    the sector-compressed screening with the permuted species masks must equal the full-space kernel and the oracle."""
    from fhsim.backend import State
    from fhsim.sharded import QubitLayout, local_sector, lower_pool_entries, pool_entries_of
    lat, u, up, dn = (2, 3), 4.0, 3, 3
    n, h_tab, pool_ops, dec, diag, o_h, o_pool = lattice(*lat, u)
    rng = np.random.default_rng(seed)
    occ_up, occ_dn, _ = pauli.k_space_occupation(*lat, 1.0, up, dn)
    picks = list(rng.choice(len(pool_ops), size=5, replace=False))
    th = rng.uniform(-0.4, 0.4, len(picks))
    psi = sv.adapt_state(n, occ_up + occ_dn, [o_pool[k] for k in picks], th)
    g_want, _, lam = sv.pool_gradients(psi, o_h, o_pool, diag, dec, n)
    layout = QubitLayout(n, 0, [int(v) for v in rng.permutation(n)])
    idx = np.arange(1 << n)
    pidx = np.zeros_like(idx)
    for b in range(n):
        pidx |= ((idx >> b) & 1) << layout.perm[b]
    psi_p, lam_p = np.zeros_like(psi), np.zeros_like(lam)
    psi_p[pidx], lam_p[pidx] = psi, lam
    plans = [GeneratorPlan(g, n) for g in pool_ops]
    entries = pool_entries_of(plans)
    local, done = lower_pool_entries(entries, layout, 0, list(range(len(entries))))
    assert len(done) == len(entries)
    dpool = DevicePool.from_entries(ctx, n, local, len(plans))
    sp, sl = State.from_numpy(ctx, psi_p), State.from_numpy(ctx, lam_p)
    up_mask, dn_mask, nu, nd = local_sector(layout, 0, up, dn)
    assert (nu, nd) == (up, dn) and up_mask | dn_mask == (1 << n) - 1
    g_full = dpool.gradients(sp, sl)
    g_sec = dpool.gradients_sector(sp, sl, nu, nd, up_mask=up_mask, dn_mask=dn_mask)
    assert np.abs(g_full - g_want).max() < G_TOL
    assert np.abs(g_sec - g_full).max() < 1e-12
    with pytest.raises(ValueError):
        dpool.gradients_sector(sp, sl, nu, nd, up_mask=up_mask, dn_mask=dn_mask | 1 | up_mask)      # overlapping masks


@pytest.mark.parametrize("lat,u,up,dn", [((2, 3), 4.0, 3, 3), ((2, 3), 4.0, 2, 4), ((3, 3), 6.0, 5, 4)])
def test_apply_table_on_the_sector_compressed_state(ctx, lat, u, up, dn):
    """fh_apply_table_sector: compress, gather over the term groups in compact coordinates, scatter back -- vs the oracle."""
    from fhsim.backend import State
    n, h_tab, pool_ops, dec, diag, o_h, o_pool = lattice(*lat, u)
    rng = np.random.default_rng(3)
    occ_up, occ_dn, _ = pauli.k_space_occupation(*lat, 1.0, up, dn)
    picks = list(rng.choice(len(pool_ops), size=6, replace=False))
    psi = sv.adapt_state(n, occ_up + occ_dn, [o_pool[k] for k in picks], rng.uniform(-0.5, 0.5, len(picks)))
    phi = sv.basis_change(psi, diag, dec, n)                 # a dense vector of the sector (W conserves both numbers)
    want = sv.apply_table(phi, o_h, n)
    dt = DeviceTable(ctx, h_tab)
    st, out = State.from_numpy(ctx, phi), State.from_numpy(ctx, np.ones(1 << n, complex))      # out must be overwritten
    e = dt.apply_sector(st, out, up, dn)
    assert np.abs(out.numpy() - want).max() < 1e-12
    assert abs(e - np.vdot(phi, want)) < E_TOL
    assert abs(dt.apply_sector(st, None, up, dn) - e) < 1e-13
    assert abs(dt.apply(st) - e) < 1e-11                      # the full-space kernel agrees
    xx = DeviceTable(ctx, PauliTable(n, [0b11], [0], [1.0]))  # X X alone moves an electron pair in: not number conserving
    with pytest.raises(ValueError):
        xx.apply_sector(st, out, up, dn)


def test_evaluate_takes_k2_and_k3_in_the_sector_at_20_qubits(ctx, monkeypatch):
    """2x5 (20 qubits): from here on fh_program_evaluate applies H on the compressed state even when lambda is needed
    (memset + scatter), and screens in the sector; both against the full-space kernels of the same program."""
    lat, u, up, dn = (2, 5), 4.0, 5, 5
    n = 20
    h_tab = PauliTable.from_operator(fermi_hubbard(*lat, 1.0, u), n)
    pool_ops = [jordan_wigner(g) for g in hubbard_interaction_pool_simplified(*lat)]
    rng = np.random.default_rng(5)
    occ_up, occ_dn, _ = pauli.k_space_occupation(*lat, 1.0, up, dn)
    picks = list(rng.choice(len(pool_ops), size=4, replace=False))
    th = rng.uniform(-0.4, 0.4, len(picks))
    plans = [GeneratorPlan(g, n) for g in pool_ops]
    circ = Circuit(n, len(picks))
    for p, k in enumerate(picks):
        circ.generator(plans[k], param=p)
    circ.marker("ansatz_end")
    circ.basis_change_separable(*lat)
    prog = circ.compile(ctx)
    dtab = DeviceTable(ctx, h_tab)
    dpool = DevicePool(ctx, plans[:150], n)
    basis = sum(1 << (n - 1 - q) for q in occ_up + occ_dn)
    m = prog.markers["ansatz_end"]
    a = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=m)
    info = prog.sector_info()
    assert info["pool_in_sector"] and info["k2_in_sector"] and not info["active"]
    ag = prog.evaluate(basis, th, [dtab], grads=True)
    assert prog.sector_info()["k2_in_sector"]
    monkeypatch.setenv("FHSIM_NO_SECTOR_POOL", "1")
    b = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=m)
    bg = prog.evaluate(basis, th, [dtab], grads=True)
    assert not prog.sector_info()["k2_in_sector"]
    assert abs(a["expvals"][0] - b["expvals"][0]) < 1e-11 and np.abs(a["pool"] - b["pool"]).max() < 1e-11
    assert np.abs(b["pool"]).max() > 1e-3
    assert abs(ag["expvals"][0] - bg["expvals"][0]) < 1e-11 and np.abs(ag["grads"] - bg["grads"]).max() < 1e-11


@pytest.mark.parametrize("lat,u,up,dn", [((2, 2), 4.0, 2, 2), ((2, 3), 4.0, 3, 3), ((2, 3), 4.0, 4, 1), ((3, 3), 6.0, 5, 4)])
def test_dense_tail_screening_vs_oracle_and_full_space(ctx, lat, u, up, dn, monkeypatch):
    """W as two dense sector blocks (W = S (U_up x U_dn) S), then H, W^dagger and K3 on compressed vectors: against the
    oracle and against the op-by-op full-space kernels of the same program; energy-only calls; graph replay."""
    monkeypatch.setenv("FHSIM_NO_SECTOR", "1")           # not the cluster kernel
    n, h_tab, pool_ops, dec, diag, o_h, o_pool = lattice(*lat, u)
    rng = np.random.default_rng(21)
    occ_up, occ_dn, _ = pauli.k_space_occupation(*lat, 1.0, up, dn)
    picks = list(rng.choice(len(pool_ops), size=5, replace=False))
    th = rng.uniform(-0.4, 0.4, len(picks))
    plans = [GeneratorPlan(g, n) for g in pool_ops]
    circ = Circuit(n, len(picks))
    for p, k in enumerate(picks):
        circ.generator(plans[k], param=p)
    circ.marker("ansatz_end")
    circ.basis_change_separable(*lat)
    prog = circ.compile(ctx)
    dtab = DeviceTable(ctx, h_tab)
    dpool = DevicePool(ctx, plans, n)
    basis = sum(1 << (n - 1 - q) for q in occ_up + occ_dn)
    m = prog.markers["ansatz_end"]
    a = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=m)
    assert prog.sector_info()["dense_tail"]
    e_only = prog.evaluate(basis, th, [dtab])
    assert prog.sector_info()["dense_tail"]
    sub = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=m, pool_range=(1, 5))
    # opt-in: the ansatz in the cluster kernel as well -- the screening never touches a full-space state (7 launches)
    monkeypatch.setenv("FHSIM_SECTOR_PREFIX", "1")
    ap = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=m)
    info_p = prog.sector_info()
    assert info_p["cluster_prefix"] and info_p["dense_tail"] and prog.last_stats()[1] <= 7
    assert abs(ap["expvals"][0] - a["expvals"][0]) < 1e-12 and np.abs(ap["pool"] - a["pool"]).max() < 1e-12
    monkeypatch.delenv("FHSIM_SECTOR_PREFIX")
    monkeypatch.setenv("FHSIM_NO_SECTOR_DENSE", "1")
    b = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=m)
    assert not prog.sector_info()["dense_tail"]
    monkeypatch.delenv("FHSIM_NO_SECTOR_DENSE")
    gens = [o_pool[k] for k in picks]
    pg_want, e_want, _ = sv.pool_gradients(sv.adapt_state(n, occ_up + occ_dn, gens, th), o_h, o_pool, diag, dec, n)
    assert abs(a["expvals"][0] - e_want) < E_TOL and np.abs(a["pool"] - pg_want).max() < G_TOL
    assert abs(a["expvals"][0] - b["expvals"][0]) < 1e-12 and np.abs(a["pool"] - b["pool"]).max() < 1e-12
    assert abs(e_only["expvals"][0] - e_want) < E_TOL
    assert np.abs(sub["pool"] - pg_want[1:6]).max() < G_TOL
    th2 = th - 0.07
    a2 = prog.evaluate(basis, th2, [dtab], pool=dpool, pool_pos=m)
    pg2, e2, _ = sv.pool_gradients(sv.adapt_state(n, occ_up + occ_dn, gens, th2), o_h, o_pool, diag, dec, n)
    assert abs(a2["expvals"][0] - e2) < E_TOL and np.abs(a2["pool"] - pg2).max() < G_TOL
    # first-epoch screening (no ansatz at all) and a pool position that is not the start of the fixed tail
    c0 = Circuit(n, 0)
    c0.marker("ansatz_end")
    c0.basis_change_separable(*lat)
    p0 = c0.compile(ctx)
    g0 = p0.evaluate(basis, [], [dtab], pool=dpool, pool_pos=0)
    assert p0.sector_info()["dense_tail"]
    pg0, e0, _ = sv.pool_gradients(sv.basis_state(n, occ_up + occ_dn), o_h, o_pool, diag, dec, n)
    assert abs(g0["expvals"][0] - e0) < E_TOL and np.abs(g0["pool"] - pg0).max() < G_TOL
    mid = prog.evaluate(basis, th, [dtab], pool=dpool, pool_pos=max(m - 1, 0))
    if m > 0:
        assert not prog.sector_info()["dense_tail"]          # parametrised ops after the pool position: op-by-op path
