#!/usr/bin/env python
"""bench.py -- ADAPT pool-gradients/s and <H> evals/s at 3x3 (18 qubits) on B200.

One "step" = one full ADAPT screening evaluation of BASELINE.json config 3
(3x3 Hubbard, U=6, (5 up, 4 down); k-space HF state evolved by the 52 first-epoch pool
operators, theta ~ U(-0.1, 0.1), default_rng(1234)):
    psi_k = U_52..U_1|HF>,  phi = W psi_k,  E = <phi|H|phi>,  lambda = W^dagger H phi,
    g_k = 2 Im <lambda|G_k|psi_k> for all 324 pool operators
i.e. what reference ADAPT.select_operator (models/adapt_vqe.py:297-323) computes, through ONE
C-ABI call (fh_program_evaluate, replayed as a CUDA graph).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 (under torchrun): every rank runs the same workload on its own GPU (independent screenings,
no data-path collective) -> "scaling": "weak"; time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "quantum-simulation-of-fermi-hubbard-model_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

NX, NY, U, N_UP, N_DN = 3, 3, 6.0, 5, 4
N_QUBITS = 2 * NX * NY
WORKLOAD = "cfg3: ADAPT screening, 3x3 Hubbard U=6 (5up,4dn), 18 qubits, 52-operator ansatz, 324-operator pool"
METRIC = "ADAPT pool-gradients/s at 3x3 (18q)"
L2_FLUSH_BYTES = 256 << 20
# dram__bytes_read.sum + dram__bytes_write.sum of one k_pool launch of this workload (ncu --set full,
# profiles/r01_s2_ncu_full_18q.md rows "k_pool"): psi and lambda (8 MiB) are read from HBM once, everything else hits L2
K3_DRAM_BYTES_PER_LAUNCH = 8431360
# the same for one k_tile_tma launch of this workload (profiles/r02_final_ncu_full_18q.md, dram_rd 4.245 MB + dram_wr ~0: ncu
# replays every kernel with a cold cache; inside the step the 4 MiB state comes from L2)
K_TILE_DRAM_BYTES_PER_LAUNCH = 4245000


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, device):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(device), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons, busy = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                clk, cmax, util = float(parts[0]), float(parts[1]), float(parts[6])
            except ValueError:
                continue
            sm.append(clk)
            mx.append(cmax)
            if util > 0:
                busy.append(clk)
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.tmp.name)
        if sm:
            out["sm_mhz"] = statistics.median(busy or sm)
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------
def build_tables():
    from fhsim.symbolic import fermi_hubbard, givens_decomposition_square, jordan_wigner
    from fhsim.tables import GeneratorPlan, PauliTable
    from operators.fourier import fourier_transform_matrix
    from operators.pool import hubbard_interaction_pool_simplified
    n = N_QUBITS
    h_tab = PauliTable.from_operator(fermi_hubbard(NX, NY, 1.0, U), n)
    plans = [GeneratorPlan(jordan_wigner(g), n) for g in hubbard_interaction_pool_simplified(NX, NY)]
    dec, diag = givens_decomposition_square(fourier_transform_matrix(NX, NY))
    occ = [0, 2, 4, 6, 12, 1, 3, 5, 7]        # stable sort of eps_k
    basis = sum(1 << (n - 1 - q) for q in occ)
    return h_tab, plans, dec, diag, basis


def build_gpu_workload(ctx, picks=None):
    from fhsim.backend import DevicePool, DeviceTable
    from fhsim.circuit import Circuit
    n = N_QUBITS
    h_tab, plans, dec, diag, basis = build_tables()
    dtab = DeviceTable(ctx, h_tab)
    dpool = DevicePool(ctx, plans, n)
    # first-epoch screening at the HF state picks the 52 operators with |g| = 2U/N
    if picks is None:
        c0 = Circuit(n, 0)
        c0.marker("ansatz_end")
        c0.basis_change(diag, list(reversed(dec)))
        p0 = c0.compile(ctx)
        g0 = p0.evaluate(basis, [], [dtab], pool=dpool, pool_pos=0)["pool"]
        picks = [k for k in range(len(plans)) if abs(g0[k]) > 1e-9]
        p0.close()
    assert len(picks) == 52, len(picks)
    thetas = np.random.default_rng(1234).uniform(-0.1, 0.1, len(picks))
    circ = Circuit(n, len(picks))
    for j, k in enumerate(picks):
        circ.generator(plans[k], param=j)
    circ.marker("ansatz_end")
    circ.basis_change_separable(NX, NY)      # same unitary as the reference W network (tests/test_circuit_host.py)
    prog = circ.compile(ctx)
    return dict(prog=prog, dtab=dtab, dpool=dpool, basis=basis, thetas=thetas, picks=picks, circ=circ,
                plans=plans, dec=dec, diag=diag, h_tab=h_tab)


def build_lattice_workload(ctx, nx, ny, u, n_ops):
    """The screening workload on an arbitrary lattice (as tools/bench_sharded.py builds it): k-space HF basis state, the
    first ``n_ops`` operators of the HF screening with theta_j = 0.05 (-1)^j, separable W, full pool."""
    from fhsim.backend import DevicePool, DeviceTable
    from fhsim.circuit import Circuit
    from fhsim.symbolic import fermi_hubbard, jordan_wigner
    from fhsim.tables import GeneratorPlan, PauliTable
    from operators.pool import hubbard_interaction_pool_simplified
    ns, n = nx * ny, 2 * nx * ny

    def f(k, length):
        if length == 1:
            return 0.0
        c = np.cos(2 * np.pi * k / length)
        return c if length == 2 else 2 * c
    eps = [round(-(f(s % nx, nx) + f(s // nx, ny)), 12) for s in range(ns)]
    n_up = (ns + 1) // 2
    n_dn = ns - n_up
    order = sorted(range(ns), key=lambda s: eps[s])
    occ = [2 * s for s in order[:n_up]] + [2 * s + 1 for s in order[:n_dn]]
    basis = sum(1 << (n - 1 - q) for q in occ)
    dtab = DeviceTable(ctx, PauliTable.from_operator(fermi_hubbard(nx, ny, 1.0, u), n))
    plans = [GeneratorPlan(jordan_wigner(g), n) for g in hubbard_interaction_pool_simplified(nx, ny)]
    dpool = DevicePool(ctx, plans, n)
    c0 = Circuit(n, 0)
    c0.marker("ansatz_end")
    c0.basis_change_separable(nx, ny)
    p0 = c0.compile(ctx)
    g0 = p0.evaluate(basis, [], [dtab], pool=dpool, pool_pos=0)["pool"]
    p0.close()
    picks = [k for k in range(len(plans)) if abs(g0[k]) > 1e-9][:n_ops]
    thetas = np.array([0.05 * (-1) ** j for j in range(len(picks))])
    circ = Circuit(n, len(picks))
    for j, k in enumerate(picks):
        circ.generator(plans[k], param=j)
    circ.marker("ansatz_end")
    circ.basis_change_separable(nx, ny)
    return dict(prog=circ.compile(ctx), dtab=dtab, dpool=dpool, basis=basis, thetas=thetas, picks=picks, plans=plans)


def cpu_reference_sample(n_ops, picks=None, thetas=None):
    """Reference CPU path (oracle/literal.py: gate-by-gate torch + autograd) on a bounded sample:
    screening of the first ``n_ops`` pool operators appended to the ansatz state.  The ansatz prefix
    is prepared untimed by the closed-form oracle (in the reference it is amortised over all 324
    operators); W, <H> and the backward pass are inside the timed region."""
    import torch
    from oracle import literal, pauli, statevector as sv
    n = N_QUBITS
    torch.set_num_threads(os.cpu_count() or 1)
    h = pauli.compress(pauli.jw_table(pauli.hubbard_fermion_terms(NX, NY, 1.0, U), n))
    opool = [pauli.jw_table(op, n) for op in pauli.pool_fermion_terms(NX, NY)]
    layers, diag = pauli.givens_network(pauli.ft_matrix(NX, NY))
    occ = [0, 2, 4, 6, 12, 1, 3, 5, 7]
    if picks is None:
        psi = sv.basis_state(n, occ)
        g0, _, _ = sv.pool_gradients(psi, h, opool, diag, layers, n)
        picks = [k for k in range(len(opool)) if abs(g0[k]) > 1e-9]
        thetas = np.random.default_rng(1234).uniform(-0.1, 0.1, len(picks))
    psi_k = sv.adapt_state(n, occ, [opool[k] for k in picks], thetas)
    h_terms = [(x, z, c) for (x, z), c in h.items()]
    sub = [literal.strings_of(p, n) for p in opool[:n_ops]]

    def one_step():
        t0 = time.perf_counter()
        e = torch.zeros(len(sub), dtype=torch.float32, requires_grad=True)
        sim = literal.LiteralSimulator(n)
        sim.state = torch.from_numpy(psi_k.reshape([2] * n).copy())
        for i, strings in enumerate(sub):
            sim.trotterize(e[i], strings)
        sim.basis_change(diag, layers)
        loss = sim.expval(h_terms)
        loss.backward()
        return time.perf_counter() - t0, e.grad.numpy().copy(), sim.gate_passes

    return one_step, torch.get_num_threads()


def reference_chunk_ops():
    """Pool operators per timed chunk of the reference-faithful CPU path: 32 ("3x3 screening is run in
    pool chunks of 32 ops"; ~19 GB of autograd-saved statevectors), fewer only if the host cannot hold that."""
    if os.environ.get("FH_BENCH_REF_OPS"):               # contract tests on small hosts
        return max(1, int(os.environ["FH_BENCH_REF_OPS"]))
    try:
        import psutil
        avail = psutil.virtual_memory().available / 2 ** 30
    except Exception:
        avail = 64.0
    return 32 if avail >= 40 else (16 if avail >= 24 else 8)


def run_cpu_closed_form(args):
    """'CPU-closed-form' of BASELINE.md 3: the SAME algorithm the GPU path runs (g_k = 2 Im <lambda|G_k|psi>,
    lambda = W^dagger H W psi) on the host cores -- oracle/cpu_closed_form.c (plain C + OpenMP) executing the same
    host-compiled op list as the CUDA program -- for the whole 324-operator screening step, on 1 thread and on all cores.
    It separates the algorithmic gain (closed form vs append-and-backprop) from the kernel/hardware gain.  No torch, no
    CUDA in this process."""
    from fhsim.circuit import Circuit, Marker
    from oracle import cpu_closed_form as cf
    n = N_QUBITS
    h_tab, plans, dec, diag, basis = build_tables()
    cores = os.cpu_count() or 1
    # first-epoch picks at the HF state, exactly as the GPU arm selects them
    c0 = Circuit(n, 0)
    c0.marker("ansatz_end")
    c0.basis_change_separable(NX, NY)
    _, g0 = cf.screening(n, basis, [], c0.ops[1:], [], h_tab, plans, threads=cores)
    picks = [k for k in range(len(plans)) if abs(g0[k]) > 1e-9]
    assert len(picks) == 52, len(picks)
    thetas = np.random.default_rng(1234).uniform(-0.1, 0.1, len(picks))
    circ = Circuit(n, len(picks))
    for j, k in enumerate(picks):
        circ.generator(plans[k], param=j)
    circ.marker("ansatz_end")
    circ.basis_change_separable(NX, NY)
    cut = next(i for i, op in enumerate(circ.ops) if isinstance(op, Marker))
    ansatz, w_ops = circ.ops[:cut], circ.ops[cut + 1:]
    out = {"cores": cores}
    for label, threads in (("1thread", 1), ("allcores", cores)):
        cf.screening(n, basis, ansatz, w_ops, thetas, h_tab, plans, threads)            # warm-up
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 3.0 and reps < 200):
            energy, grads = cf.screening(n, basis, ansatz, w_ops, thetas, h_tab, plans, threads)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        out[label] = {"value": len(plans) / dt, "unit": "gradients/s", "ms_per_screening": 1e3 * dt, "threads": threads,
                      "reps": reps}
    out["energy"] = float(energy)
    out["checksum"] = float(np.abs(grads).sum())
    out["sample"] = ("whole 324-operator screening of the bench workload (52 ansatz rotations, W, H, W^dagger, 324 pair-form "
                     "reductions) by oracle/cpu_closed_form.c: gcc -O3 -fopenmp, complex128, same op list as the CUDA program")
    print(json.dumps({"impl": "cpu-closed-form", **out}))


# ---------------------------------------------------------------------------------------------
def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    n_ops = reference_chunk_ops()
    one_step, cores = cpu_reference_sample(n_ops)
    budget_s = 150.0
    times = []
    t_first, _, passes = one_step()                       # warm-up / calibration (untimed)
    steps = max(1, min(args.steps, int(budget_s / max(t_first, 1e-3))))
    warm = 1
    for _ in range(steps):
        dt, _, _ = one_step()
        times.append(dt)
    total = sum(times)
    value = n_ops * len(times) / total
    sample = (f"one chunk of {n_ops} of the 324 pool operators appended to the 52-operator ansatz state and back-propagated "
              f"(reference algorithm, adapt_vqe.py:297-310, in pool chunks as BASELINE.md 3 specifies); W, the 100-term <H> and "
              f"their backward are inside the timed region and amortised over the {n_ops} operators of the chunk; the ansatz "
              f"prefix is prepared untimed (the reference would re-run it per chunk); {passes} gate passes per step; "
              f"steps bounded to ~{int(budget_s)} s of CPU work")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "gradients/s", "n_gpus": args.gpus,
        "steps": len(times), "steps_requested": args.steps, "warmup": warm, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": "gradients/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "gradients/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_gpu_arm(args, rank, world, local_rank):
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from fhsim.backend import Context, State
    ctx = Context(local_rank if world > 1 else 0)
    wl = build_gpu_workload(ctx)
    prog, dtab, dpool, basis, thetas = wl["prog"], wl["dtab"], wl["dpool"], wl["basis"], wl["thetas"]
    marker = prog.markers["ansatz_end"]
    n_pool = dpool.n_out

    def step():
        return prog.evaluate(basis, thetas, [dtab], pool=dpool, pool_pos=marker)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank if world > 1 else 0) if rank == 0 else None

    # ---- value: device-resident, CUDA events around the graph, L2 flushed before every step ----
    for _ in range(max(args.warmup, 3)):
        ctx.flush_l2(L2_FLUSH_BYTES)
        step()
    barrier()
    dev_ms = []
    launches = 0
    for _ in range(args.steps):
        ctx.flush_l2(L2_FLUSH_BYTES)
        ctx.sync()
        res = step()
        ms, launches = prog.last_stats()
        dev_ms.append(ms)
    barrier()
    total_ms = sum(dev_ms)
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    # ---- e2e: wall clock around the public call (host theta in, host gradients out) ----
    wall = []
    for _ in range(args.steps):
        ctx.flush_l2(L2_FLUSH_BYTES)
        ctx.sync()
        t0 = time.perf_counter()
        res = step()
        wall.append(time.perf_counter() - t0)
    e2e_total = sum(wall)
    if dist is not None:
        t = torch.tensor([e2e_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_total = float(t.item())

    # ---- the two paths that really shard across GPUs; every rank takes part ----
    # (1) Pool-sharded screening of ONE state.  At 18 qubits it is withdrawn: K3 is ~17 % of a 0.2 ms step, so splitting
    #     the pool cannot pay for a collective (round 1 measured 0.46-0.60 ms vs 0.23 ms on one GPU) -- 18-qubit
    #     screenings are replicas only.  It is measured where K3 dominates the step: the 3x4 lattice (24 qubits, 792
    #     operators, state 256 MiB), one screening on one GPU vs the same screening with the pool split over the ranks.
    # (2) cfg 5: the 4x4 lattice (32 qubits, 64 GiB state) sharded by its top index bits, global<->local qubit swaps
    #     through fh_comm (NCCL all-to-all pipelined against the local bit permutation).
    pool_sharded = None
    sharded_4x4 = None
    if dist is not None:
        from fhsim.parallel import screen_pool_sharded
        dev = torch.device("cuda", local_rank)
        try:
            wl24 = build_lattice_workload(ctx, 3, 4, 4.0, 16)
            p24, m24 = wl24["prog"], wl24["prog"].markers["ansatz_end"]
            one = p24.evaluate(wl24["basis"], wl24["thetas"], [wl24["dtab"]], pool=wl24["dpool"], pool_pos=m24)
            full = screen_pool_sharded(p24, wl24["basis"], wl24["thetas"], [wl24["dtab"]], wl24["dpool"], m24, dist, dev)
            parity = float(np.abs(full["pool"] - one["pool"]).max())
            times = {}
            for label in ("single_gpu", "pool_sharded"):
                ts = []
                for _ in range(5):
                    barrier()
                    t0 = time.perf_counter()
                    if label == "single_gpu":
                        p24.evaluate(wl24["basis"], wl24["thetas"], [wl24["dtab"]], pool=wl24["dpool"], pool_pos=m24)
                    else:
                        screen_pool_sharded(p24, wl24["basis"], wl24["thetas"], [wl24["dtab"]], wl24["dpool"], m24, dist, dev)
                    ctx.sync()
                    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ts.append(float(t.item()))
                times[label] = 1e3 * statistics.median(ts[1:])
            n24 = wl24["dpool"].n_out
            pool_sharded = {"workload": "3x4 Hubbard, 24 qubits, 16-operator ansatz, 792-operator pool, one screening",
                            "single_gpu_ms": times["single_gpu"], "pool_sharded_ms": times["pool_sharded"],
                            "speedup": times["single_gpu"] / times["pool_sharded"], "ops_per_rank": -(-n24 // world),
                            "max_abs_diff_vs_single_gpu": parity,
                            "collective": "all_gather of <= %d doubles per rank (NCCL); psi / lambda recomputed per rank" % -(-n24 // world),
                            "at_18_qubits": "withdrawn: replicas only (K3 is ~10 % of the step; see DESIGN 6)",
                            "note": "K3 runs on sector-compressed vectors inside fh_program_evaluate (400 gradients at 24 qubits: 0.21 ms "
                                    "instead of 4.09 ms), so it no longer dominates this screening and splitting the pool does not pay; "
                                    "kept as a measured negative result"}
            for o in (wl24["prog"], wl24["dpool"], wl24["dtab"]):
                o.close()
        except Exception as exc:
            print(f"pool-sharded measurement failed: {exc!r}", file=sys.stderr)
        if not args.no_sharded:
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                from bench_sharded import run_sharded
                sharded_4x4 = run_sharded("4x4", 4.0, 16, 1, dist, local_rank, False, True, measured_peak_gbs()[0])
            except Exception as exc:
                print(f"sharded 4x4 measurement failed: {exc!r}", file=sys.stderr)
            torch.cuda.set_device(local_rank)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- <H> evals/s (energy only: forward + K2 expectation) ----
    for _ in range(3):
        prog.evaluate(basis, thetas, [dtab])
    h_ms = []
    for _ in range(min(args.steps, 500)):
        ctx.flush_l2(L2_FLUSH_BYTES)
        ctx.sync()
        prog.evaluate(basis, thetas, [dtab])
        h_ms.append(prog.last_stats()[0])
    h_launches = prog.last_stats()[1]
    step()   # restore the screening graph

    # ---- one optimiser step of the same circuit: E + the 52 ansatz gradients by the adjoint sweep (information) ----
    train = None
    try:
        for _ in range(3):
            prog.evaluate(basis, thetas, [dtab], grads=True)
        t_ms = []
        for _ in range(min(args.steps, 200)):
            ctx.flush_l2(L2_FLUSH_BYTES)
            ctx.sync()
            prog.evaluate(basis, thetas, [dtab], grads=True)
            t_ms.append(prog.last_stats()[0])
        train = {"ms": statistics.median(t_ms), "launches": prog.last_stats()[1],
                 "what": "fh_program_evaluate(grads=True): energy + 52 ansatz gradients (adjoint sweep), device time"}
    except Exception as exc:
        print(f"train-step measurement failed: {exc!r}", file=sys.stderr)
    step()   # restore the screening graph

    # ---- dominant kernel (K3 pool screening) timed alone for the roofline ----
    n = N_QUBITS
    psi_k, lam = State(ctx, n), State(ctx, n)
    psi_k.set_basis(basis)
    prog.run(psi_k, thetas, 0, marker)
    lam.copy_from(psi_k)
    prog.run(lam, thetas, marker, prog.n_items - marker)
    phi = State(ctx, n)
    phi.copy_from(lam)
    dtab.apply(phi, lam)
    prog.run(lam, thetas, marker, prog.n_items - marker, dagger=True)
    g_check = dpool.gradients(psi_k, lam)
    assert np.abs(g_check - res["pool"]).max() < 1e-9
    k3_ms = []
    reps = 10
    for _ in range(10):
        ctx.flush_l2(L2_FLUSH_BYTES)
        # in the real step psi and lambda were produced just before the screening kernel, i.e. they are
        # L2-resident there too: touch them after the flush, then time `reps` back-to-back launches
        psi_k.norm2(); lam.norm2()
        ctx.timer_start()
        for _r in range(reps):
            dpool.enqueue(psi_k, lam)
        k3_ms.append(ctx.timer_stop() / reps)
    k3 = statistics.median(k3_ms) * 1e-3
    # K3 as the step runs it when the evaluation conserves (N_up, N_dn): on sector-compressed copies of psi_s / lambda_s
    step_info = prog.sector_info()
    k3_sector = None
    if step_info["pool_in_sector"] or step_info["active"]:
        g_sec = dpool.gradients_sector(psi_k, lam, N_UP, N_DN)
        assert np.abs(g_sec - res["pool"]).max() < 1e-9
        ks = []
        for _ in range(10):
            ctx.flush_l2(L2_FLUSH_BYTES)
            psi_k.norm2(); lam.norm2()
            ctx.timer_start()
            for _r in range(reps):
                dpool.gradients_sector(psi_k, lam, N_UP, N_DN, enqueue_only=True)
            ks.append(ctx.timer_stop() / reps)
        k3_sector = statistics.median(ks) * 1e-3
    peak, peak_src = measured_peak_gbs()
    alg_bytes = 4.0 * (1 << n) * n_pool              # 4*2^n B per gradient
    achieved = alg_bytes / k3 / 1e9

    # ---- the kernel that dominates the step by time: k_tile (15 of the 19 launches, ~73 % of the step in the ncu
    # launch list profiles/r01_s8_launches.csv).  Timed alone: the 13 forward runs on one state and the 2 W-dagger runs
    # of the lambda branch, launched back to back (warm L2, as inside the step), CUDA events on the library's stream.
    tile = None
    try:
        if prog.n_tiles == prog.n_items:
            if step_info.get("dense_tail"):
                # the step launches only the tile runs before the fixed tail (the ansatz); W / W^dagger run as dense sector blocks
                t_fwd = statistics.median(prog.time_items(phi, 0, marker, False, 20) for _ in range(5))
                n_tile_launches = marker
                tile_s = t_fwd * 1e-3
            else:
                n_dag = prog.n_items - marker
                t_fwd = statistics.median(prog.time_items(phi, 0, prog.n_items, False, 20) for _ in range(5))
                t_dag = statistics.median(prog.time_items(lam, marker, n_dag, True, 20) for _ in range(5))
                n_tile_launches = prog.n_items + n_dag
                tile_s = (t_fwd + t_dag) * 1e-3
            tile_alg = 32.0 * (1 << n) * n_tile_launches          # one read + one write of the state per launch
            tile = {"bound": "hbm", "kernel": "k_tile_tma (fused runs of rotations on TMA-staged shared-memory tiles)",
                    "achieved": tile_alg / tile_s / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": tile_alg / tile_s / 1e9 / peak, "traffic": K_TILE_DRAM_BYTES_PER_LAUNCH,
                    "peak_source": peak_src, "kernel_ms": 1e3 * tile_s / n_tile_launches,
                    "launches_per_step": n_tile_launches, "share_of_step": 1e3 * tile_s / (total_ms / args.steps),
                    "note": "18-qubit state (4 MiB) is L2-resident and the launches are latency-bound (128 CTAs, one "
                            "dependent chain per fused op, 6.3 us launch-to-launch floor): effective GB/s vs HBM peak; "
                            "algorithmic bytes = 32*2^n per launch; traffic = ncu dram bytes of one launch under cold-cache "
                            "replay; the HBM-bound figure for this kernel is hbm_regime.tile_W; the rest of the step (W, H, "
                            "W^dagger, K3 on sector-compressed vectors) is roofline_k3 / profiles/r02_final_ncu_full_18q.md"}
    except Exception as exc:                                       # never lose the bench line over a side measurement
        print(f"k_tile roofline measurement failed: {exc}", file=sys.stderr)
        tile = None

    # ---- the same kernels where the state no longer fits L2 (3x4 lattice, 24 qubits, 256 MiB): true HBM rooflines ----
    hbm = None
    if world == 1 and not args.no_hbm_regime:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from sweep_roofline import measure_lattice
        row = measure_lattice(ctx, "3x4", peak, pool_cap=400, verbose=False)
        hbm = {"workload": "3x4 Hubbard, 24 qubits, state 256 MiB (> L2), back-to-back launches, CUDA events",
               "unit": "GB/s", "peak": peak}
        for key, label in (("screening", "k_pool: 400 gradients, 4*2^n B each"),
                           ("pair_dense", "k_pair: Pauli-string rotation, 32*2^n B"),
                           ("givens", "k_pair: Givens rotation, 16*2^n B"),
                           ("pair_fermi4", "k_pair: fermionic double excitation, 4*2^n B"),
                           ("diag_coulomb", "k_diag: Coulomb layer, 32*2^n B"),
                           ("tile_W", "k_tile: one fused launch of W, 32*2^n B"),
                           ("h_apply", "k_table_pass x2 (K2 in shared-memory tile passes): H psi + <H>, 32*2^n B"),
                           ("h_apply_sector", "k_sector_compress1 + k_sector_happly + memset + k_sector_scatter1: the same H psi + <H> on "
                                              "the sector-compressed state (what fh_program_evaluate runs for number-conserving "
                                              "evaluations); effective rate on 32*2^n B"),
                           ("screening_sector", "k_sector_compress2 + k_sector_pool: the same 400 gradients on sector-compressed "
                                                "psi / lambda (853 776 amplitudes, L2-resident); effective rate on 4*2^n B each")):
            if key not in row:
                continue
            hbm[key] = {"what": label, "us": round(row[key]["us"], 2), "achieved": round(row[key]["GBps"], 1),
                        "frac": round(row[key]["frac"], 4)}

    # ---- K4: Lanczos on sector-compressed vectors (cfg 4's ED reference, 3x3 (5 up, 4 down), dim 15 876) ----
    k4 = None
    try:
        from fhsim.backend import lanczos_sector
        lanczos_sector(dtab, N_UP, N_DN, k=1, tol=1e-12, max_iter=2000, seed=7)             # warm-up (allocations)
        ev, _, _, info = lanczos_sector(dtab, N_UP, N_DN, k=1, tol=1e-12, max_iter=2000, seed=7)
        alg = 96.0 * info["sector_dim"] * info["matvecs"]           # 96 B per amplitude and iteration
        k4 = {"bound": "hbm", "kernel": "k_sector_matvec + CGS2 re-orthogonalisation (fh_lanczos_sector)",
              "achieved": alg / info["loop_seconds"] / 1e9, "peak": peak, "unit": "GB/s",
              "frac": alg / info["loop_seconds"] / 1e9 / peak, "traffic": None, "peak_source": peak_src,
              "matvecs_per_s": info["matvecs"] / info["loop_seconds"], "matvecs": info["matvecs"],
              "host_syncs": info["host_syncs"], "sector_dim": info["sector_dim"], "E0": float(ev[0]),
              "note": "3x3 sector vectors are 254 KB (L1/L2-resident): the iteration is launch-bound, the fraction of the HBM "
                      "roofline is nominal; the HBM-regime run is the 4x4 lattice (165 636 900 amplitudes, 2.65 GB per vector), "
                      "tools/lanczos_4x4.py -> profiles/r02_lanczos_4x4.json"}
    except Exception as exc:
        print(f"K4 measurement failed: {exc!r}", file=sys.stderr)

    clocks = sampler.stop() if sampler else {}

    # ---- CPU baseline (bounded sample of the same workload) ----
    cpu = None
    cpu_closed = None
    if not args.no_cpu_baseline and world == 1:          # rank 0 at N=1 only
        n_ops = reference_chunk_ops()
        warm_step, cores = cpu_reference_sample(2, wl["picks"], thetas)
        warm_step()                                       # thread pool / allocator warm-up on a 2-operator chunk
        one_step, cores = cpu_reference_sample(n_ops, wl["picks"], thetas)
        dt, grads_cpu, passes = one_step()
        cpu_val = n_ops / dt
        # float32 parameters and gradients, as the reference has them: agreement is bounded by float32 rounding here; the
        # 1e-12 agreement of this path with the closed form in float64 is tests/test_oracle_literal.py
        err = float(np.abs(np.abs(grads_cpu) - np.abs(res["pool"][:n_ops]).astype(np.float32)).max())
        cpu = {"value": cpu_val, "unit": "gradients/s", "cores": cores, "kind": "port",
               "sample": (f"one chunk of {n_ops} of the 324 pool operators by the reference's append-and-backprop algorithm "
                          f"(oracle/literal.py, torch CPU, {passes} gate passes, {dt:.1f} s; W + <H> + backward amortised over "
                          f"the chunk, ansatz prefix untimed); max |g_cpu - g_gpu| on the chunk = {err:.2e} (float32 gradients)")}
        try:                                              # fair algorithm-for-algorithm CPU number, in a fresh process
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "cpu-closed-form"], capture_output=True,
                                 text=True, timeout=600)
            cpu_closed = json.loads(out.stdout.strip().splitlines()[-1])
            cpu_closed.pop("impl", None)
            ref_sum = float(np.abs(res["pool"]).sum())
            cpu_closed["abs_sum_error_vs_gpu"] = abs(cpu_closed.pop("checksum") - ref_sum)
            cpu_closed["energy_error_vs_gpu"] = abs(cpu_closed.pop("energy") - float(res["expvals"][0]))
        except Exception as exc:
            print(f"closed-form CPU leg failed: {exc}", file=sys.stderr)

    roofline_k3_full = {"bound": "hbm", "kernel": "k_pool32 + k_pool_finalize (K3 in the full space; not in the step when K3 runs in the sector)",
                        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": K3_DRAM_BYTES_PER_LAUNCH, "peak_source": peak_src, "kernel_ms": k3 * 1e3}
    if k3_sector is not None:
        k3, achieved = k3_sector, alg_bytes / k3_sector / 1e9
    roofline_k3 = {"bound": "hbm", "kernel": ("k_sector_compress2 + k_sector_pool (K3 on sector-compressed psi_s / lambda_s)"
                                              if k3_sector is not None else "k_pool (K3 pool screening) + k_pool_finalize"),
                   "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                   "traffic": (None if k3_sector is not None else K3_DRAM_BYTES_PER_LAUNCH), "peak_source": peak_src, "kernel_ms": k3 * 1e3,
                   "share_of_step": k3 * 1e3 / (total_ms / args.steps),
                   "note": "the kernels that produce the metric's unit (324 gradients); algorithmic bytes = 4*2^n per gradient "
                           "x 324 (SURVEY 8(d), full-space figure), so this is an EFFECTIVE rate vs the HBM peak: in the sector "
                           "the screening touches 324 x 1 225 pairs of the 15 876-amplitude compressed vectors (L2-resident); "
                           "the same screening in the full space is roofline_k3_full_space"}
    value = world * n_pool * args.steps / (total_ms * 1e-3)
    e2e_value = world * n_pool * args.steps / e2e_total
    h2d, d2h = prog.payload_bytes()            # pinned op-payload arena in, scalars + pool gradients out (per step)
    line = {
        "metric": METRIC, "value": value, "unit": "gradients/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "complex128", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "flushed (256 MiB write) before every timed step",
                   "parallelism": f"replicas x{world}" if world > 1 else "single GPU",
                   "launch_items": prog.n_items, "tile_kernels": prog.n_tiles,
                   "path": ("sector-resident cluster kernel" if step_info["active"] else
                            "full-space tile kernels for the ansatz, then W (two dense sector blocks), H, W^dagger and K3 on "
                            "sector-compressed vectors" if step_info.get("dense_tail") else
                            "full-space tile kernels + K2, K3 on sector-compressed psi_s / lambda_s" if step_info["pool_in_sector"]
                            else "full-space tile kernels + K2 + K3")},
        "e2e": {"value": e2e_value, "unit": "gradients/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_total / args.steps},
        "gpu_launches": launches * args.steps,
        "launches_per_step": launches,
        "h_evals_per_s": world * len(h_ms) / (sum(h_ms) * 1e-3),
        "h_eval_ms": statistics.median(h_ms), "h_eval_launches": h_launches,
        "train_step": train,
        "roofline": tile if tile is not None else roofline_k3,
        "roofline_k3": roofline_k3,
        "roofline_k3_full_space": roofline_k3_full,
        "roofline_k4": k4,
        "hbm_regime": hbm,
        "cpu_baseline": cpu,
        "cpu_closed_form": cpu_closed,
        "pool_sharded": pool_sharded,
        "sharded_4x4": sharded_4x4,
        "clocks": clocks,
        "energy": float(res["expvals"][0]),
    }
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="fhsim", choices=["fhsim", "reference", "cpu-closed-form"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hbm-regime", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="skip the sharded 4x4 (32-qubit) block of multi-GPU runs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.impl == "cpu-closed-form":
        if rank == 0:
            run_cpu_closed_form(args)
        return
    if world == 1 and args.gpus > 1:
        print(json.dumps({"error": f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks"}))
        sys.exit(2)
    run_gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
