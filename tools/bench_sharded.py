#!/usr/bin/env python
"""Sharded-state ADAPT screening over 1/2/4/8 B200 (BASELINE config 5 and its smaller siblings).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_sharded.py --lattice 4x4 --u 4 --n-ops 16 --steps 2 --json gpurun_out/shard_4x4.json

Workload: Lx x Ly Hubbard, t=1, U; k-space HF basis state evolved by the first ``n_ops``
non-zero-gradient pool operators (found by a first screening at the HF state), theta_j = 0.05 (-1)^j; then
phi = W psi, E = <phi|H|phi>, lambda = W^dagger H phi and the full pool-gradient scan -- on a state sharded by
its top log2(N) index bits, with global<->local qubit swaps (local bit permutation + NCCL all-to-all).

Parity inside the run (no CPU oracle can hold 2^32 amplitudes):
  * HF screening: E = sum eps_k + U N_up N_dn / N exactly, |g_k| in {0, 2U/N}           (analytic)
  * with --check-single (state fits one GPU): energy and every gradient against the unsharded single-GPU
    fh_program_evaluate on rank 0.
Times are CUDA-event / synchronised wall times, max over ranks.
"""
import argparse
import json
import os
import sys
import time

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "quantum-simulation-of-fermi-hubbard-model_b200")]

import numpy as np  # noqa: E402


def eps_k(nx, ny, t=1.0):
    def f(k, length):
        if length == 1:
            return 0.0
        c = np.cos(2 * np.pi * k / length)
        return c if length == 2 else 2 * c
    return [round(-t * (f(s % nx, nx) + f(s // nx, ny)), 12) for s in range(nx * ny)]


def run_sharded(lattice="4x4", u=4.0, n_ops=16, steps=2, dist=None, local_rank=0, check_single=False, profile=True,
                peak_gbs=None):
    """The sharded-state screening benchmark on an initialised process group (or a single GPU when ``dist`` is None).
    Every rank calls it; rank 0 gets the result dict, the others None."""
    import torch
    d = dist
    rank = d.get_rank() if d is not None else 0
    world = d.get_world_size() if d is not None else 1
    torch.cuda.set_device(local_rank)

    from fhsim.circuit import Circuit
    from fhsim.sharded import CudaEngine, ShardedSimulator
    from fhsim.symbolic import fermi_hubbard, jordan_wigner
    from fhsim.tables import GeneratorPlan, PauliTable
    from operators.pool import hubbard_interaction_pool_simplified

    nx, ny = map(int, lattice.split("x"))
    ns, n = nx * ny, 2 * nx * ny
    g = world.bit_length() - 1
    t0 = time.time()
    h_tab = PauliTable.from_operator(fermi_hubbard(nx, ny, 1.0, u), n)
    plans = [GeneratorPlan(jordan_wigner(op), n) for op in hubbard_interaction_pool_simplified(nx, ny)]
    n_up = (ns + 1) // 2
    n_dn = ns - n_up
    eps = eps_k(nx, ny)
    order = sorted(range(ns), key=lambda s: eps[s])                  # stable: ties by ascending orbital index
    occ = [2 * s for s in order[:n_up]] + [2 * s + 1 for s in order[:n_dn]]
    basis = sum(1 << (n - 1 - q) for q in occ)
    e_hf = sum(eps[s] for s in order[:n_up]) + sum(eps[s] for s in order[:n_dn]) + u * n_up * n_dn / ns
    w = Circuit(n, 0)
    w.basis_change_separable(nx, ny)
    host_s = time.time() - t0

    engine = CudaEngine(n - g, local_rank, d)
    sim = ShardedSimulator(engine, n)

    def sync_max(x):
        torch.cuda.synchronize()
        if d is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        d.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    # ---- first screening at the HF state: analytic known answers + operator picks -------------------------
    empty = Circuit(n, 0)
    torch.cuda.synchronize()
    t0 = time.time()
    e0, g0 = sim.adapt_screening(basis, empty.ops, w.ops, h_tab, plans)
    first_s = sync_max(time.time() - t0)
    gmax = 2.0 * u / ns
    spectrum_ok = bool(np.all((np.abs(g0) < 1e-9) | (np.abs(np.abs(g0) - gmax) < 1e-9)))
    picks = [k for k in range(len(plans)) if abs(g0[k]) > 1e-9][:n_ops]
    thetas = np.array([0.05 * (-1) ** j for j in range(len(picks))])
    ans = Circuit(n, len(picks))
    for j, k in enumerate(picks):
        ans.generator(plans[k], param=j)

    # ---- timed evaluations: exchanges are enqueued without host synchronisation -------------------------------
    res = []
    for step in range(steps + 1):                          # first one is the warm-up (program compilation)
        engine.a2a_ms = 0.0
        engine.time_exchanges = False
        s0, p0 = sim.swap_count, dict(sim.pass_count)
        if d is not None:
            d.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        e1, _ = sim.adapt_screening(basis, ans.ops, w.ops, h_tab, plans, thetas, len(picks), want_gradients=False)
        t_energy = sync_max(time.time() - t0)
        t0 = time.time()
        e2, g2 = sim.adapt_screening(basis, ans.ops, w.ops, h_tab, plans, thetas, len(picks))
        t_full = sync_max(time.time() - t0)
        res.append(dict(energy_s=t_energy, screening_s=t_full, swaps=sim.swap_count - s0,
                        table_passes=sim.pass_count["table"] - p0["table"], pool_passes=sim.pass_count["pool"] - p0["pool"]))
    timed = res[1:]
    best = min(timed, key=lambda r: r["screening_s"])
    phases = None
    if profile:
        # one more evaluation with a synchronisation after every phase and after every exchange: where the time goes
        sim.profiling = True
        engine.time_exchanges = True
        engine.a2a_ms = 0.0
        sim.adapt_screening(basis, ans.ops, w.ops, h_tab, plans, thetas, len(picks))
        phases = {k: round(sync_max(v), 5) for k, v in sim.profile.items()}
        phases["a2a_ms_inside"] = round(sync_max(engine.a2a_ms), 3)
        sim.profiling = False

    check = None
    if check_single and rank == 0 and n <= 28:
        from fhsim.backend import DevicePool, DeviceTable
        ctx = engine.ctx
        c1 = Circuit(n, len(picks))
        for j, k in enumerate(picks):
            c1.generator(plans[k], param=j)
        c1.marker("ansatz_end")
        c1.basis_change_separable(nx, ny)
        prog = c1.compile(ctx)
        one = prog.evaluate(basis, thetas, [DeviceTable(ctx, h_tab)], pool=DevicePool(ctx, plans, n),
                            pool_pos=prog.markers["ansatz_end"])
        check = {"max_abs_gradient_diff": float(np.abs(one["pool"] - g2).max()),
                 "energy_diff": float(abs(one["expvals"][0] - e2.real))}
        prog.close()

    out = None
    if rank == 0:
        n_pool = len(plans)
        alg_bytes = 4.0 * (1 << n) * n_pool
        out = {
            "lattice": lattice, "n_qubits": n, "n_gpus": world, "slab_GiB": 16.0 * (1 << (n - g)) / 2 ** 30,
            "pool": n_pool, "ansatz_ops": len(picks), "host_compile_s": round(host_s, 2),
            "communicator": "fh_comm (NCCL inside libfhsim, pipelined swap exchange)" if engine._comm is not None
                            else ("torch.distributed all_to_all_single" if world > 1 else "none"),
            "hf_screening": {"E": e0.real, "E_analytic": e_hf, "abs_err": abs(e0.real - e_hf),
                             "nonzero": int(np.sum(np.abs(g0) > 1e-9)), "gmax": float(np.abs(g0).max()),
                             "gmax_analytic": gmax, "spectrum_in_{0,2U/N}": spectrum_ok, "seconds_cold": first_s},
            "energy": e2.real, "energy_imag": e2.imag, "energy_consistency": abs(e1.real - e2.real),
            "grad_abs_max": float(np.abs(g2).max()), "grad_l2": float(np.linalg.norm(g2)),
            "timed_steps": timed, "best": best, "phase_seconds": phases,
            "h_evals_per_s": 1.0 / best["energy_s"],
            "gradients_per_s": n_pool / best["screening_s"],
            "screening_effective_GBps_all_gpus": alg_bytes / best["screening_s"] / 1e9,
            "check_vs_single_gpu": check,
            "pool_scan": ("every slab screened on its sector-compressed copy (fh_pool_gradients_sector_masks)"
                          if getattr(sim, "sector_pool_used", False) else "full local space (k_pool32 / k_pool_tile)"),
            "grad_l2_reference_full_space_scan": 9.543770434819237 if (lattice == "4x4" and len(picks) == 16) else None,
        }
        if peak_gbs and phases and phases.get("pool_scan"):
            # 4 * 2^n B per gradient is the full-space figure: with the sector scan this is an effective rate (can exceed 1)
            out["pool_scan_hbm_frac_per_gpu"] = alg_bytes / world / phases["pool_scan"] / 1e9 / peak_gbs
    engine.close()                       # slabs, cached tables / pools / programs, communicator
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lattice", default="4x4")
    ap.add_argument("--u", type=float, default=4.0)
    ap.add_argument("--n-ops", type=int, default=16)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--check-single", action="store_true")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    d = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        d = dist
    out = run_sharded(args.lattice, args.u, args.n_ops, args.steps, d, local_rank, args.check_single)
    if out is not None:
        print(json.dumps(out))
        if args.json:
            os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
            with open(args.json, "w") as f:
                json.dump(out, f, indent=1)
    if d is not None:
        d.barrier()
        d.destroy_process_group()


if __name__ == "__main__":
    main()
