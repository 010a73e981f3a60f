"""Sharded-state path (fhsim/sharded.py) on CPU: planner/lowering unit tests and world_size 2 / 4 gloo runs
with the numpy slab engine (tests/emulate_sharded.py) against the unsharded oracle.

The -m gpu twin (tests/test_gpu_sharded.py) runs the same flow with CudaEngine over NCCL.
"""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

from fhsim.circuit import Circuit, PairOpSpec  # noqa: E402
from fhsim.sharded import (QubitLayout, ShardedSimulator, choose_globals_cover, dagger_ops, lower_pair,  # noqa: E402
                           plan_circuit, renormalise, swap_steps)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def build_case(nx, ny, u, n_ops, seed, separable):
    from fhsim.symbolic import fermi_hubbard, givens_decomposition_square, jordan_wigner
    from fhsim.tables import GeneratorPlan, PauliTable
    from operators.fourier import fourier_transform_matrix
    from operators.pool import hubbard_interaction_pool_simplified
    from oracle import pauli
    n = 2 * nx * ny
    h_tab = PauliTable.from_operator(fermi_hubbard(nx, ny, 1.0, u), n)
    plans = [GeneratorPlan(jordan_wigner(g), n) for g in hubbard_interaction_pool_simplified(nx, ny)]
    dec, diag = givens_decomposition_square(fourier_transform_matrix(nx, ny))
    rng = np.random.default_rng(seed)
    picks = [int(v) for v in rng.choice(len(plans), size=n_ops, replace=False)]
    thetas = rng.uniform(-0.4, 0.4, n_ops)
    ans = Circuit(n, n_ops)
    for j, k in enumerate(picks):
        ans.generator(plans[k], param=j)
    w = Circuit(n, n_ops)
    if separable:
        w.basis_change_separable(nx, ny)
    else:
        w.basis_change(diag, list(reversed(dec)))
    n_up = (nx * ny + 1) // 2
    up, dn, _ = pauli.k_space_occupation(nx, ny, 1.0, n_up, nx * ny - n_up)
    occ = up + dn
    basis = sum(1 << (n - 1 - q) for q in occ)
    o_h = pauli.compress(pauli.jw_table(pauli.hubbard_fermion_terms(nx, ny, 1.0, u), n))
    o_pool = [pauli.jw_table(op, n) for op in pauli.pool_fermion_terms(nx, ny)]
    return dict(n=n, h_tab=h_tab, plans=plans, dec=dec, diag=diag, picks=picks, thetas=thetas, ans=ans, w=w, occ=occ,
                basis=basis, o_h=o_h, o_pool=o_pool)


def run_case(engine_factory, case):
    """Sharded evaluation vs the unsharded oracle; returns the simulator (for counters)."""
    from oracle import statevector as sv
    c = case
    n = c["n"]
    psi_k = sv.adapt_state(n, c["occ"], [c["o_pool"][k] for k in c["picks"]], c["thetas"])
    want_g, want_e, want_lam = sv.pool_gradients(psi_k, c["o_h"], c["o_pool"], c["diag"], c["dec"], n)
    engine = engine_factory(n)
    sim = ShardedSimulator(engine, n)
    # state after the ansatz, gathered back into logical order
    st = sim.new_state()
    sim.set_basis(st, c["basis"])
    sim.apply_ops(st, c["ans"].ops, c["thetas"], len(c["thetas"]))
    got = sim.gather(st)
    assert np.abs(got - psi_k).max() < 1e-12
    # U^dagger U = 1 through the planner's own inverse
    sim.apply_ops(st, dagger_ops(c["ans"].ops), c["thetas"], len(c["thetas"]))
    ref = np.zeros(1 << n, complex)
    ref[c["basis"]] = 1.0
    assert np.abs(sim.gather(st) - ref).max() < 1e-12
    st.close()
    energy, grads = sim.adapt_screening(c["basis"], c["ans"].ops, c["w"].ops, c["h_tab"], c["plans"], c["thetas"],
                                        len(c["thetas"]))
    assert abs(energy.real - want_e) < 1e-10 and abs(energy.imag) < 1e-10
    assert np.abs(grads - want_g).max() < 1e-9
    e_only, none = sim.adapt_screening(c["basis"], c["ans"].ops, c["w"].ops, c["h_tab"], c["plans"], c["thetas"],
                                       len(c["thetas"]), want_gradients=False)
    assert none is None and abs(e_only.real - want_e) < 1e-10
    return sim


def _worker(rank, world, port, spec):
    import torch.distributed as dist
    sys.path[:0] = [HERE, os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "quantum-simulation-of-fermi-hubbard-model_b200")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from emulate_sharded import NumpyEngine
        g = world.bit_length() - 1
        case = build_case(*spec)
        sim = run_case(lambda n: NumpyEngine(n - g, dist), case)
        assert sim.swap_count > 0, "the test case never exercised a global<->local swap"
        assert sim.engine.calls["all_to_all"] == sim.swap_count
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,spec", [
    (2, (2, 2, 4.0, 5, 11, False)),
    (2, (2, 3, 4.0, 6, 12, True)),
    (4, (2, 3, 4.0, 4, 13, False)),
])
def test_sharded_matches_oracle_gloo(world, spec):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), spec), nprocs=world, join=True)


def test_single_rank_degenerate_case():
    from emulate_sharded import NumpyEngine
    case = build_case(2, 2, 4.0, 4, 3, True)
    sim = run_case(lambda n: NumpyEngine(n, None), case)
    assert sim.swap_count == 0


def test_swap_steps_and_layout_bookkeeping():
    lay = QubitLayout(10, 2)
    assert lay.global_logical_bits() == [8, 9]
    pairs, new = swap_steps(lay, [0, 7])
    # logical 7 already sits at the top of the slab (phys 7), logical 0 has to be brought to phys 6
    assert pairs == [(0, 6)]
    assert sorted(new.global_logical_bits()) == [0, 7]
    assert sorted(new.perm) == list(range(10))
    assert new.is_local(1 << 8) and new.is_local(1 << 9) and not new.is_local(1 << 0)
    # the all-to-all exchanges phys nl-g+k with phys nl+k
    assert new.perm[8] in (6, 7) and new.perm[9] in (6, 7)


def test_renormalise_covers_every_pair_once():
    rng = np.random.default_rng(0)
    n = 7
    for _ in range(200):
        x = int(rng.integers(1, 1 << n))
        xbits = [b for b in range(n) if x >> b & 1]
        # pattern pins a random non-empty subset of the x bits (plus some other bits), as device ops do
        sub = [b for b in xbits if rng.random() < 0.5] or [xbits[int(rng.integers(len(xbits)))]]
        fm = sum(1 << b for b in sub) | (int(rng.integers(0, 1 << n)) & ~x)
        fv = int(rng.integers(0, 1 << n)) & fm
        want = {(i, i ^ x) for i in range(1 << n) if (i & fm) == fv}
        got = set()
        for fm2, fv2, swapped in renormalise(x, fm, fv, 0):
            top = 1 << (x.bit_length() - 1)
            assert fm2 & top and not fv2 & top
            for i in range(1 << n):
                if (i & fm2) == fv2:
                    pair = (i ^ x, i) if swapped else (i, i ^ x)
                    assert pair not in got
                    got.add(pair)
        assert got == want


def test_lowered_ops_reproduce_the_global_op():
    """Every pair op lowered rank by rank (all layouts reachable by one swap) equals the op on the full vector."""
    from emulate import apply_op
    rng = np.random.default_rng(5)
    n, g = 8, 2
    nl = n - g
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    lay = QubitLayout(n, g)
    _, lay = swap_steps(lay, [1, 4])
    for trial in range(40):
        local_bits = [b for b in range(n) if lay.perm[b] < nl]
        xb = rng.choice(local_bits, size=int(rng.integers(1, 4)), replace=False)
        x = int(sum(1 << int(b) for b in xb))
        top = 1 << (x.bit_length() - 1)
        fm = top | (int(rng.integers(0, 1 << n)) & ~top)
        fv = int(rng.integers(0, 1 << n)) & fm & ~top
        ze = int(rng.integers(0, 1 << n))
        kind = int(rng.integers(0, 2))
        m = tuple(rng.normal(size=8))
        op = PairOpSpec(x, fm, fv, ze, kind=kind, param=0 if kind else -1, scale=0.7, bhat=np.exp(1j * rng.normal()),
                        matrix=m, strings=[(x, ze)])
        want = apply_op(psi, op, [0.3], n)
        # physical view of psi
        idx = np.arange(1 << n, dtype=np.uint64)
        phys = np.zeros(1 << n, dtype=np.uint64)
        for b in range(n):
            phys |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(lay.perm[b])
        full_phys = np.zeros(1 << n, complex)
        full_phys[phys] = psi
        out_phys = np.zeros(1 << n, complex)
        for rank in range(1 << g):
            slab = full_phys[rank << nl:(rank + 1) << nl].copy()
            for lo in lower_pair(op, lay, rank):
                slab = apply_op(slab, lo, [0.3], nl)
            out_phys[rank << nl:(rank + 1) << nl] = slab
        assert np.abs(out_phys[phys] - want).max() < 1e-12, trial


def test_planner_keeps_every_op_local_and_cover_choice_progresses():
    case = build_case(2, 3, 4.0, 8, 21, False)
    n, g = case["n"], 2
    steps, final = plan_circuit(case["ans"].ops + case["w"].ops, QubitLayout(n, g))
    n_swaps = sum(1 for s in steps if s[0] == "swap")
    assert n_swaps >= 1
    for s in steps:
        if s[0] == "ops":
            for op in s[1]:
                if isinstance(op, PairOpSpec):
                    assert s[2].is_local(op.x)
    lay = QubitLayout(n, g)
    masks = [int(x) for x in case["h_tab"].x if int(x)]
    remaining = [m for m in masks if not lay.is_local(m)]
    assert remaining
    _, lay2 = swap_steps(lay, choose_globals_cover(remaining, lay))
    assert all(lay2.is_local(m) for m in remaining)


def _pool_worker(rank, world, port, n_out):
    import torch.distributed as dist
    sys.path[:0] = [HERE, os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "quantum-simulation-of-fermi-hubbard-model_b200")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fhsim.parallel import pool_ranges, screen_pool_sharded
        truth = np.sin(np.arange(n_out) * 0.37)

        class FakePool:
            pass

        class FakeProgram:                       # stands in for DeviceProgram.evaluate (needs a GPU)
            def evaluate(self, basis, thetas, tables, pool=None, pool_pos=0, pool_range=None):
                first, count = pool_range
                return {"expvals": np.array([1.5]), "pool": truth[first:first + count].copy()}

        pool = FakePool()
        pool.n_out = n_out
        res = screen_pool_sharded(FakeProgram(), 0, [], [], pool, 0, dist)
        assert np.array_equal(res["pool"], truth)
        assert sum(c for _, c in pool_ranges(n_out, world)) == n_out
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_out", [(2, 324), (3, 24), (2, 1)])
def test_pool_sharded_gather_gloo(world, n_out):
    import torch.multiprocessing as mp
    mp.spawn(_pool_worker, args=(world, _free_port(), n_out), nprocs=world, join=True)


def test_pool_ranges_match_survey_split():
    from fhsim.parallel import pool_ranges
    assert [c for _, c in pool_ranges(324, 8)] == [41, 41, 41, 41, 40, 40, 40, 40]
    assert pool_ranges(5, 8)[5:] == [(5, 0), (5, 0), (5, 0)]
