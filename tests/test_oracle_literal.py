"""The literal CPU path (oracle/literal.py: gate-by-gate PauliStringRotation recipe of reference
models/utils.py:58-83, per-term expval, append-every-pool-operator + loss.backward() of adapt_vqe.py:297-310)
against the closed form (oracle/statevector.py).  oracle/literal.py is what bench.py times as "the reference CPU
path", so it is itself checked here: same energies, same screening gradients, same ansatz gradients."""
import numpy as np
import pytest
import torch

from oracle import literal, pauli, statevector as sv


def _lattice(nx, ny, u):
    n = 2 * nx * ny
    h = pauli.compress(pauli.jw_table(pauli.hubbard_fermion_terms(nx, ny, 1.0, u), n))
    pool = [pauli.jw_table(op, n) for op in pauli.pool_fermion_terms(nx, ny)]
    layers, diag = pauli.givens_network(pauli.ft_matrix(nx, ny))
    return n, h, pool, layers, diag


def test_literal_pauli_rotation_equals_closed_form():
    """utils.py:58-83 recipe == cos(theta/2) - i sin(theta/2) P on random strings (weights 1..6)."""
    n = 6
    rng = np.random.default_rng(11)
    for _ in range(20):
        x, z = int(rng.integers(0, 1 << n)), int(rng.integers(0, 1 << n))
        if x == 0 and z == 0:
            continue
        theta = float(rng.uniform(-3, 3))
        v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
        v /= np.linalg.norm(v)
        (letters, wires, _), = literal.strings_of({(x, z): 1.0}, n)
        sim = literal.LiteralSimulator(n)
        sim.state = torch.from_numpy(v.reshape([2] * n).copy())
        sim.pauli_string_rotation(theta, letters, wires)
        assert np.abs(sim.state.reshape(-1).numpy() - sv.pauli_rotation(v, theta, x, z, n)).max() < 1e-14


def test_literal_screening_equals_closed_form_2x2_full_pool():
    """cfg 1: all 24 operators, at the HF state and at a 3-operator ansatz state, float64 parameters, 1e-12."""
    n, h, pool, layers, diag = _lattice(2, 2, 4.0)
    up, dn, _ = pauli.k_space_occupation(2, 2, 1.0, 2, 2)
    occ = up + dn
    h_terms = [(x, z, c) for (x, z), c in h.items()]
    strings = [literal.strings_of(p, n) for p in pool]
    for sel, thetas in (([], []), ([1, 2, 4], [0.3, -0.2, 0.11])):
        psi = sv.adapt_state(n, occ, [pool[k] for k in sel], thetas)
        g_closed, e_closed, _ = sv.pool_gradients(psi, h, pool, diag, layers, n)
        g_lit, passes = literal.screen_by_backprop(n, occ, [strings[k] for k in sel], thetas, strings, diag, layers,
                                                   h_terms, dtype=torch.float64)
        assert passes > 3000                                   # 3 264 gate passes for the pool part
        assert np.abs(g_lit - g_closed).max() < 1e-12
        loss, _ = literal.adapt_eval_circuit(n, occ, [strings[k] for k in sel], torch.tensor(thetas, dtype=torch.float64),
                                             [], torch.zeros(0, dtype=torch.float64), diag, layers, h_terms)
        assert abs(float(loss) - e_closed) < 1e-12
    # the reference's own precision: float32 parameters and gradients; first epoch = eight |g| = 2, sixteen 0
    g32, _ = literal.screen_by_backprop(n, occ, [], [], strings, diag, layers, h_terms)
    assert g32.dtype == np.float32
    assert sorted(np.round(np.abs(g32), 5).tolist()) == [0.0] * 16 + [2.0] * 8
    # chunked evaluation (what bench.py times at 3x3) is the same arithmetic
    g_chunk, _ = literal.screen_by_backprop(n, occ, [], [], strings, diag, layers, h_terms, chunk=5, dtype=torch.float64)
    g_full, _ = literal.screen_by_backprop(n, occ, [], [], strings, diag, layers, h_terms, dtype=torch.float64)
    assert np.abs(g_chunk - g_full).max() < 1e-13


def test_literal_ansatz_gradient_equals_adjoint_sweep_2x2():
    """loss.backward() through the literal train-mode circuit == oracle adjoint sweep (adapt_vqe.py:415-418)."""
    n, h, pool, layers, diag = _lattice(2, 2, 4.0)
    up, dn, _ = pauli.k_space_occupation(2, 2, 1.0, 2, 2)
    occ = up + dn
    sel, thetas = [1, 2, 4, 5, 9], np.array([0.3, -0.2, 0.11, 0.05, -0.4])
    strings = [literal.strings_of(pool[k], n) for k in sel]
    t = torch.tensor(thetas, dtype=torch.float64, requires_grad=True)
    loss, _ = literal.adapt_eval_circuit(n, occ, strings, t, [], torch.zeros(0, dtype=torch.float64), diag, layers,
                                         [(x, z, c) for (x, z), c in h.items()])
    loss.backward()
    e, g = sv.adjoint_gradient(n, occ, [pool[k] for k in sel], thetas, h, diag, layers)
    assert abs(float(loss) - e) < 1e-12
    assert np.abs(t.grad.numpy() - g).max() < 1e-12


def test_literal_screening_equals_closed_form_3x3_sample():
    """cfg 3 (the timed workload): 8 pool operators appended to a 6-operator ansatz state at 18 qubits."""
    nx, ny, u = 3, 3, 6.0
    n, h, pool, layers, diag = _lattice(nx, ny, u)
    occ = [0, 2, 4, 6, 12, 1, 3, 5, 7]
    rng = np.random.default_rng(1234)
    g0, e0, _ = sv.pool_gradients(sv.basis_state(n, occ), h, pool, diag, layers, n)
    assert abs(e0 + 5.0 / 3.0) < 1e-12
    picks = [k for k in range(len(pool)) if abs(g0[k]) > 1e-9]
    assert len(picks) == 52
    sel = picks[:6]
    thetas = rng.uniform(-0.1, 0.1, len(sel))
    sample = picks[6:10] + [0, 1, 2, 3]                      # four live operators and four arbitrary ones
    psi = sv.adapt_state(n, occ, [pool[k] for k in sel], thetas)
    g_closed, _, _ = sv.pool_gradients(psi, h, [pool[k] for k in sample], diag, layers, n)
    g_lit, _ = literal.screen_by_backprop(n, occ, [literal.strings_of(pool[k], n) for k in sel], thetas,
                                          [literal.strings_of(pool[k], n) for k in sample], diag, layers,
                                          [(x, z, c) for (x, z), c in h.items()], dtype=torch.float64)
    assert np.abs(g_closed).max() > 0.1
    assert np.abs(g_lit - g_closed).max() < 1e-12
