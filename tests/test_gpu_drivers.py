"""Drop-in drivers (models/*, linalg/*) on the CUDA backend vs the oracle.

Energies 1e-10, gradients 1e-9 (abs, float64 values before the float32 cast of ``.grad``), operator
selection sequences identical.
"""
import os

import numpy as np
import pytest
import torch

from oracle import ed, pauli, statevector as sv

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _workdir(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)          # drivers write ./results and ./images


def oracle_lattice(nx, ny, u):
    n = 2 * nx * ny
    h = pauli.compress(pauli.jw_table(pauli.hubbard_fermion_terms(nx, ny, 1.0, u), n))
    pool = [pauli.jw_table(op, n) for op in pauli.pool_fermion_terms(nx, ny)]
    layers, diag = pauli.givens_network(pauli.ft_matrix(nx, ny))
    return n, h, pool, layers, diag


def select_like_reference(grads32, ratio, threshold1):
    g = np.abs(grads32)
    ng = int(np.sum((g >= g.max() * ratio) * (g >= threshold1)))
    return np.argsort(g)[::-1][:ng].tolist()


def test_adapt_2x2_first_epochs_match_oracle():
    from models.adapt_vqe import ADAPT
    nx, ny, u = 2, 2, 4.0
    n, h, pool, layers, diag = oracle_lattice(nx, ny, u)
    vqe = ADAPT(n_epoch=2, threshold1=1e-2, threshold2=5e-2, x_dimension=nx, y_dimension=ny, n_electrons=4,
                n_spin_up=2, n_spin_down=2, tunneling=1, coulomb=u, verbose=False)
    assert abs(vqe.ground_state_energy + 2.1027484835) < 1e-9
    occ = vqe.spin_up_indices + vqe.spin_down_indices
    oup, odn, _ = pauli.k_space_occupation(nx, ny, 1.0, 2, 2)
    assert occ == oup + odn
    # epoch-1 screening: E_HF = 0, eight operators at |g| = 2, same sequence as the oracle
    ops, gates, max_grads = vqe.select_operator()
    psi0 = sv.basis_state(n, occ)
    g0, e0, _ = sv.pool_gradients(psi0, h, pool, diag, layers, n)
    want = select_like_reference(g0.astype(np.float32), 0.1, 1e-2)
    assert vqe.last_selected_indices == want and len(want) == 8
    assert np.allclose(max_grads, 2.0, atol=1e-6)
    vqe.run()
    assert len(vqe.results['epoch loss']) == 2
    assert abs(vqe.results['iteration loss'][0] - 0.0) < 1e-10           # first loss is E_HF
    assert np.all(np.diff(vqe.results['epoch loss']) < 1e-9)
    # energy, Sz, S^2, fidelity and gradient at the final parameters vs the oracle
    sel = [vqe.fermionOperatorPool.index(op) for op in vqe.results['selected operators']]
    th = vqe.params['t'].detach().to(torch.float64).numpy()
    vqe.params['t'].grad = None                                           # run() leaves its last gradient behind
    loss, sz, s2 = vqe.circuit(mode='train')
    loss.backward()
    e_or, g_or = sv.adjoint_gradient(n, occ, [pool[k] for k in sel], th, h, diag, layers)
    assert abs(loss.item() - e_or) < 1e-10
    assert np.abs(vqe.params['t'].grad.numpy() - g_or.astype(np.float32)).max() < 1e-6
    assert abs(sz.item()) < 1e-10
    state = vqe.circuit(mode='state').numpy()
    phi = sv.basis_change(sv.adapt_state(n, occ, [pool[k] for k in sel], th), diag, layers, n)
    assert np.abs(state - phi).max() < 1e-11                              # same global phase as the reference network
    fid = abs(np.vdot(vqe.ground_state_wf, phi)) ** 2
    assert abs(vqe._fidelity() - fid) < 1e-10
    # next screening at these parameters: identical operator sequence
    vqe.select_operator()
    g1, _, _ = sv.pool_gradients(sv.adapt_state(n, occ, [pool[k] for k in sel], th), h, pool, diag, layers, n)
    assert vqe.last_selected_indices == select_like_reference(g1.astype(np.float32), 0.1, 1e-2)
    # checkpoint round trip
    vqe.save_model()
    again = ADAPT(n_epoch=2, threshold1=1e-2, threshold2=5e-2, x_dimension=nx, y_dimension=ny, n_electrons=4,
                  n_spin_up=2, n_spin_down=2, tunneling=1, coulomb=u, verbose=False, load_model=True)
    assert torch.equal(again.params['t'], vqe.params['t']) and len(again.selected_gates) == len(vqe.selected_gates)
    assert abs(again.circuit(mode='train')[0].item() - e_or) < 1e-10
    # portable twin: remove the pickles, resume from <model>.npz
    os.remove(vqe.model_filepath)
    os.remove(vqe.result_filepath)
    twin = ADAPT(n_epoch=2, threshold1=1e-2, threshold2=5e-2, x_dimension=nx, y_dimension=ny, n_electrons=4,
                 n_spin_up=2, n_spin_down=2, tunneling=1, coulomb=u, verbose=False, load_model=True)
    assert torch.equal(twin.params['t'], vqe.params['t'])
    assert [twin._pool_index_of(g) for g in twin.selected_gates] == sel
    assert twin.results['epoch loss'] == vqe.results['epoch loss']
    assert twin.results['selected operators'] == vqe.results['selected operators']
    assert abs(twin.circuit(mode='train')[0].item() - e_or) < 1e-10


def test_adapt_2x3_selection_sequence():
    from models.adapt_vqe import ADAPT
    nx, ny, u = 2, 3, 4.0
    n, h, pool, layers, diag = oracle_lattice(nx, ny, u)
    vqe = ADAPT(n_epoch=1, threshold1=1e-2, threshold2=1e-1, x_dimension=nx, y_dimension=ny, n_electrons=6,
                n_spin_up=3, n_spin_down=3, tunneling=1, coulomb=u, verbose=False)
    occ = vqe.spin_up_indices + vqe.spin_down_indices
    vqe.run()
    assert abs(vqe.results['iteration loss'][0] + 2.0) < 1e-10           # E_HF = -2
    assert vqe.results['n_params'] == [17]
    sel = [vqe.fermionOperatorPool.index(op) for op in vqe.results['selected operators']]
    g0, _, _ = sv.pool_gradients(sv.basis_state(n, occ), h, pool, diag, layers, n)
    assert sel == select_like_reference(g0.astype(np.float32), 0.1, 1e-2)
    th = vqe.params['t'].detach().to(torch.float64).numpy()
    vqe.select_operator()
    g1, _, _ = sv.pool_gradients(sv.adapt_state(n, occ, [pool[k] for k in sel], th), h, pool, diag, layers, n)
    assert vqe.last_selected_indices == select_like_reference(g1.astype(np.float32), 0.1, 1e-2)
    assert abs(vqe.ground_state_energy + 3.7898230717) < 1e-9


def hva_oracle_energy(model, thetas_u, thetas_h, thetas_v, n, h, diag, layers):
    from fhsim.tables import pack_term

    def table(op):
        return {pack_term(t, n): c for t, c in op.terms.items()}
    psi = sv.basis_state(n, model.spin_up_indices + model.spin_down_indices)
    psi = sv.basis_change(psi, diag, layers, n)
    gens = model.hvaGenerators
    for rep in range(model.reps):
        psi = sv.trotterize(psi, thetas_u[rep], table(gens['coulomb']), n)
        for i in range(model.Nv):
            psi = sv.trotterize(psi, thetas_v[rep * model.Nv + i], table(gens['vertical'][i]), n)
        for i in range(model.Nh):
            psi = sv.trotterize(psi, thetas_h[rep * model.Nh + i], table(gens['horizontal'][i]), n)
    psi = sv.trotterize(psi, thetas_u[model.reps], table(gens['coulomb']), n)
    return sv.expval(psi, h, n).real, psi


def test_hva_2x3_energy_and_gradients():
    from models.hva import HVA
    nx, ny, u = 2, 3, 4.0
    n, h, pool, layers, diag = oracle_lattice(nx, ny, u)
    vqe = HVA(n_epoch=3, reps=2, lr=1e-2, threshold=1e-2, x_dimension=nx, y_dimension=ny, n_electrons=6, n_spin_up=3,
              n_spin_down=3, tunneling=1, coulomb=u, verbose=False)
    assert (vqe.Nh, vqe.Nv) == (1, 3)
    rng = np.random.default_rng(20260)
    with torch.no_grad():
        for k in ('theta_U', 'theta_h', 'theta_v'):
            vqe.params[k].copy_(torch.from_numpy(rng.uniform(-0.3, 0.3, vqe.params[k].numel()).astype(np.float32)))
    tu, thh, tv = (vqe.params[k].detach().to(torch.float64).numpy() for k in ('theta_U', 'theta_h', 'theta_v'))
    loss, sz, s2 = vqe.circuit(vqe.params['theta_U'], vqe.params['theta_h'], vqe.params['theta_v'], mode='train')
    loss.backward()
    e_or, psi_or = hva_oracle_energy(vqe, tu, thh, tv, n, h, diag, layers)
    assert abs(loss.item() - e_or) < 1e-10
    assert abs(vqe.fidelity_from_overlaps(vqe._last_overlaps) - abs(np.vdot(vqe.ground_state_wf, psi_or)) ** 2) < 1e-10
    state = vqe.circuit(vqe.params['theta_U'], vqe.params['theta_h'], vqe.params['theta_v'], mode='state').numpy()
    assert np.abs(state - psi_or).max() < 1e-11
    hstep = 1e-5
    for name, vec in (('theta_U', tu), ('theta_h', thh), ('theta_v', tv)):
        for j in range(len(vec)):
            plus, minus = vec.copy(), vec.copy()
            plus[j] += hstep
            minus[j] -= hstep
            args_p = [plus if name == k else v for k, v in (('theta_U', tu), ('theta_h', thh), ('theta_v', tv))]
            args_m = [minus if name == k else v for k, v in (('theta_U', tu), ('theta_h', thh), ('theta_v', tv))]
            fd = (hva_oracle_energy(vqe, *args_p, n, h, diag, layers)[0]
                  - hva_oracle_energy(vqe, *args_m, n, h, diag, layers)[0]) / (2 * hstep)
            assert abs(vqe.params[name].grad[j].item() - fd) < 2e-6        # .grad is float32
    vqe.run()
    assert len(vqe.results['loss']) == 3 and vqe.results['loss'][-1] < vqe.results['loss'][0] + 1e-9
    assert os.path.exists(vqe.model_filepath)


def test_iqcc_2x2_screening_and_dressing():
    from fhsim.backend import DeviceTable, default_context, lanczos
    from fhsim.symbolic import fermi_hubbard
    from fhsim.tables import PauliTable, pack_term
    from models.iqcc_hubbard import IQCC
    vqe = IQCC(fermi_hubbard(2, 2, 1.0, 4.0), n_epoch=1, lr=5e-2, threshold=5e-2, verbose=False)
    n = vqe.n_qubits
    # oracle: QMF state |11110000>, gradient of exp(-i tau P/2) at tau = 0 is Im <H psi| P psi>
    h = {pack_term(t, n): c for t, c in vqe.currentHamiltonian.terms.items()}
    psi = sv.basis_state(n, [])
    for q in range(n):
        psi = sv.ry(psi, float(vqe.params['theta'][q]), q, n)
    lam = sv.apply_table(psi, h, n)
    gates, names, grads = vqe.select_operator()
    from fhsim.symbolic import QubitOperator
    want = []
    for flip in vqe.partition_hamiltonian().keys():
        if not flip:
            continue
        (x, z), = [pack_term(t, n) for t in QubitOperator(' '.join(('Y' if k == 0 else 'X') + str(q) for k, q in enumerate(flip))).terms]
        want.append(abs(np.vdot(lam, sv.apply_pauli(psi, x, z, n)).imag))
    want = np.array(want, dtype=np.float32)
    assert len(want) == 8                                            # 9 x-mask groups minus the diagonal one
    ng = int(np.sum(want > want.max() * 0.1)) if want.max() * 0.1 > 5e-2 else int(np.sum(want > 5e-2))
    assert len(grads) == ng and np.allclose(np.sort(grads), np.sort(want)[-ng:], atol=1e-6)
    e_before = vqe.ground_state_energy
    vqe.run()
    assert len(vqe.loss_history['epoch']) == 1
    # dressing is a unitary transformation: the spectrum of the dressed Hamiltonian is unchanged
    tab = DeviceTable(default_context(), PauliTable.from_operator(vqe.currentHamiltonian, n))
    vals, _, _ = lanczos(tab, k=1, tol=1e-10)
    assert abs(vals[0] - e_before) < 1e-8
    dense = np.array([sv.apply_table(np.eye(1 << n, dtype=complex)[:, k], h, n) for k in range(1 << n)]).T
    assert abs(e_before - np.linalg.eigvalsh(dense)[0]) < 1e-9        # full-space ground state of the original H
    assert vqe.loss_history['epoch'][0] <= vqe.loss_history['iteration'][0] + 1e-9


def test_vqe_hea_energy_and_gradient():
    from fhsim.symbolic import fermi_hubbard
    from fhsim.tables import pack_term
    from models.vqe_hea import VQE
    vqe = VQE(fermi_hubbard(2, 2, 1.0, 4.0), n_epoch=2, reps=2, lr=1e-1, threshold=1e-3, seed=3, verbose=False)
    n = vqe.n_qubits
    h = {pack_term(t, n): c for t, c in vqe.qmlHamiltonian.operator.terms.items()}

    def energy(p):
        psi = sv.basis_state(n, [])
        for rep in range(vqe.reps):
            for q in range(n):
                psi = sv.rx(psi, p[rep, q, 0], q, n)
                psi = sv.ry(psi, p[rep, q, 1], q, n)
                psi = sv.rz(psi, p[rep, q, 2], q, n)
            for q in range(n):
                psi = sv.cnot(psi, q, (q + 1) % n, n)
        for q in range(n):
            psi = sv.rx(psi, p[vqe.reps - 1, q, 0], q, n)
            psi = sv.ry(psi, p[vqe.reps - 1, q, 1], q, n)
            psi = sv.rz(psi, p[vqe.reps - 1, q, 2], q, n)
        return sv.expval(psi, h, n).real
    p0 = vqe.params[0].detach().to(torch.float64).numpy()
    loss = vqe.circuit()
    loss.backward()
    assert abs(loss.item() - energy(p0)) < 1e-10
    g = vqe.params[0].grad.numpy()
    assert np.all(g[vqe.reps] == 0)                                   # row `reps` is never read (reference quirk)
    for (rep, q, k) in [(0, 0, 0), (1, 3, 1), (0, 5, 2), (1, 7, 0)]:
        plus, minus = p0.copy(), p0.copy()
        plus[rep, q, k] += 1e-5
        minus[rep, q, k] -= 1e-5
        assert abs(g[rep, q, k] - (energy(plus) - energy(minus)) / 2e-5) < 2e-6
    vqe.run()
    assert len(vqe.loss_history) == 2


def test_linalg_dropin_vs_scipy_sector_ed():
    from fhsim.symbolic import fermi_hubbard
    from linalg.exact_diagonalization import (get_sparse_operator, jw_get_ground_state, jw_get_ground_state_for_3x3,
                                              jw_number_spin_indices)
    n, h, _, _, _ = oracle_lattice(2, 3, 4.0)
    assert jw_number_spin_indices(6, 3, 3, n) == sorted(ed.jw_number_spin_indices(6, 3, 3, n))
    with pytest.raises(ValueError):
        jw_number_spin_indices(5, 3, 3, n)
    op = get_sparse_operator(fermi_hubbard(2, 3, 1.0, 4.0))
    e0, wf = jw_get_ground_state(op, 6, 3, 3)
    want, vecs, idx = ed.ground_state(h, n, 6, 3, 3, k=1)
    assert abs(e0 - want[0]) < 1e-9
    assert abs(abs(np.vdot(vecs[0], wf)) - 1.0) < 1e-8
    outside = np.ones(1 << n, bool)
    outside[idx] = False
    assert np.abs(wf[outside]).max() == 0.0                           # the state never leaves the sector
    e0b, basis = jw_get_ground_state_for_3x3(op, 6, 3, 3)
    assert abs(e0b - want[0]) < 1e-9 and len(basis) == 4
    gram = np.array([[np.vdot(a, b) for b in basis] for a in basis])
    assert np.abs(gram - np.eye(4)).max() < 1e-8


# ---------------------------------------------------------------------------------------------
# BASELINE configs 3 and 4 through the drop-in drivers (18 qubits)
# ---------------------------------------------------------------------------------------------
def test_adapt_3x3_first_epoch_matches_oracle():
    """models/adapt_vqe_for_3x3.ADAPT (reference adapt_vqe_for_3x3.py): 4-fold degenerate ED level, first screening
    (52 operators at |g| = 4/3) in the oracle's order, E_HF = -5/3, projected fidelity, and one
    energy + gradient evaluation with the 52 new parameters."""
    from models.adapt_vqe_for_3x3 import ADAPT
    nx, ny, u = 3, 3, 6.0
    n, h, pool, layers, diag = oracle_lattice(nx, ny, u)
    vqe = ADAPT(n_epoch=1, threshold1=1e-2, threshold2=1e-2, x_dimension=nx, y_dimension=ny, n_electrons=9,
                n_spin_up=5, n_spin_down=4, tunneling=1, coulomb=u, verbose=False)
    assert abs(vqe.ground_state_energy + 5.5623088363) < 1e-8
    wfs = np.array(vqe.ground_state_wfs)
    assert wfs.shape == (4, 1 << n)
    assert np.abs(wfs.conj() @ wfs.T - np.eye(4)).max() < 1e-9              # orthonormal degenerate subspace
    occ = vqe.spin_up_indices + vqe.spin_down_indices
    assert occ == [0, 2, 4, 6, 12, 1, 3, 5, 7]
    ops, gates, max_grads = vqe.select_operator()
    psi0 = sv.basis_state(n, occ)
    g0, e0, _ = sv.pool_gradients(psi0, h, pool, diag, layers, n)
    assert abs(e0 + 5.0 / 3.0) < 1e-12
    want = select_like_reference(g0.astype(np.float32), 0.1, 1e-2)
    assert len(want) == 52 and vqe.last_selected_indices == want
    assert np.allclose(max_grads, 4.0 / 3.0, atol=1e-6)
    # append them as run() does (reference adapt_vqe.py:388-389) and evaluate once at small random parameters
    vqe.selected_gates += gates
    rng = np.random.default_rng(1234)
    th = rng.uniform(-0.1, 0.1, len(gates)).astype(np.float32)
    vqe.params['t'] = torch.from_numpy(th)
    loss, sz, s2 = vqe.circuit(mode='train')
    loss.backward()
    th64 = th.astype(np.float64)
    e_or, g_or = sv.adjoint_gradient(n, occ, [pool[k] for k in want], th64, h, diag, layers)
    assert abs(loss.item() - e_or) < 1e-10
    assert np.abs(vqe.params['t'].grad.numpy() - g_or.astype(np.float32)).max() < 1e-6
    assert abs(sz.item() - 0.5) < 1e-10
    phi = sv.basis_change(sv.adapt_state(n, occ, [pool[k] for k in want], th64), diag, layers, n)
    fid = vqe.calculate_fidelity(vqe.ground_state_wfs, phi)
    assert abs(vqe._fidelity() - fid) < 1e-9 and 0.0 <= fid <= 1.0 + 1e-12
    state = vqe.circuit(mode='state').numpy()
    assert np.abs(state - phi).max() < 1e-11
    # second screening at these parameters: same sequence as the oracle
    vqe.select_operator()
    g1, _, _ = sv.pool_gradients(sv.adapt_state(n, occ, [pool[k] for k in want], th64), h, pool, diag, layers, n)
    assert vqe.last_selected_indices == select_like_reference(g1.astype(np.float32), 0.1, 1e-2)


def test_hva_3x3_energy_state_and_projected_fidelity():
    from models.hva_for_3x3 import HVA
    nx, ny, u = 3, 3, 6.0
    n, h, pool, layers, diag = oracle_lattice(nx, ny, u)
    vqe = HVA(n_epoch=1, reps=1, lr=1e-2, threshold=1e-2, x_dimension=nx, y_dimension=ny, n_electrons=9, n_spin_up=5,
              n_spin_down=4, tunneling=1, coulomb=u, verbose=False)
    assert (vqe.Nh, vqe.Nv) == (3, 3)                       # odd periodic length: even / odd / wrap-around bonds
    rng = np.random.default_rng(20260)
    with torch.no_grad():
        for k in ('theta_U', 'theta_h', 'theta_v'):
            vqe.params[k].copy_(torch.from_numpy(rng.uniform(-0.3, 0.3, vqe.params[k].numel()).astype(np.float32)))
    tu, thh, tv = (vqe.params[k].detach().to(torch.float64).numpy() for k in ('theta_U', 'theta_h', 'theta_v'))
    loss, sz, s2 = vqe.circuit(vqe.params['theta_U'], vqe.params['theta_h'], vqe.params['theta_v'], mode='train')
    loss.backward()
    e_or, psi_or = hva_oracle_energy(vqe, tu, thh, tv, n, h, diag, layers)
    assert abs(loss.item() - e_or) < 1e-10
    assert abs(sz.item() - 0.5) < 1e-10
    state = vqe.circuit(vqe.params['theta_U'], vqe.params['theta_h'], vqe.params['theta_v'], mode='state').numpy()
    assert np.abs(state - psi_or).max() < 1e-11
    fid = vqe.calculate_fidelity(vqe.ground_state_wfs, psi_or)
    assert abs(vqe.fidelity_from_overlaps(vqe._last_overlaps) - fid) < 1e-9
    # one gradient component by central differences of the oracle
    hstep = 1e-5
    plus, minus = tu.copy(), tu.copy()
    plus[0] += hstep
    minus[0] -= hstep
    fd = (hva_oracle_energy(vqe, plus, thh, tv, n, h, diag, layers)[0]
          - hva_oracle_energy(vqe, minus, thh, tv, n, h, diag, layers)[0]) / (2 * hstep)
    assert abs(vqe.params['theta_U'].grad[0].item() - fd) < 2e-6


def test_iqcc_3x3_partition_and_full_space_lanczos():
    """cfg 4: iQCC on 3x3 -- 36 non-zero x-mask groups of the Hubbard Hamiltonian (reference iqcc_hubbard.py:82-101),
    screening gradients of the YX..X generators at the QMF start state vs the oracle, and the Lanczos replacement
    of openfermion.get_ground_state over the FULL 2^18 space (iqcc_hubbard.py:57) vs sector ED minima."""
    from fhsim.symbolic import fermi_hubbard
    from models.iqcc_hubbard import IQCC
    nx, ny, u = 3, 3, 6.0
    n = 18
    model = IQCC(fermi_hubbard(nx, ny, 1.0, u), n_epoch=1, lr=1e-2, threshold=1e-2, verbose=False)
    groups = model.partition_hamiltonian()
    assert len([flip for flip in groups if len(flip)]) == 36
    # the global ground level of 3x3, U = 6 without a chemical potential lies at 6 electrons (degenerate in Sz:
    # (3,3), (4,2), (5,1) all give -9.7352272458, oracle/ed.py over all 55 (N_up >= N_dn) sectors)
    o_h = pauli.compress(pauli.jw_table(pauli.hubbard_fermion_terms(nx, ny, 1.0, u), n))
    best = ed.ground_state(o_h, n, 6, 3, 3, k=1)[0][0]
    assert abs(best + 9.735227245787) < 1e-9
    assert abs(model.ground_state_energy - best) < 1e-8
    wf = np.asarray(model.ground_state_wf)
    assert abs(np.linalg.norm(wf) - 1.0) < 1e-9
    assert abs(np.vdot(wf, sv.apply_table(wf, o_h, n)).real - best) < 1e-7      # it is an eigenvector of that level
