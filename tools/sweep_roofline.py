#!/usr/bin/env python
"""Roofline sweep of the four kernels over qubit count (SURVEY 8(d) scaling sweep).

For each synthetic Hubbard lattice (2x5 = 20 q, 3x4 = 24 q, 2x6 = 24 q, 2x7 = 28 q, 3x5 = 30 q ...) time,
with CUDA events on the launching stream and back-to-back repetitions:

  pair_fermi4   one fermionic double-excitation rotation (pool generator)    4 * 2^n  B algorithmic
  pair_dense    one Pauli-string rotation exp(-i theta P/2)                   32 * 2^n B
  givens        one adjacent Givens (SingleExcitation)                        16 * 2^n B
  diag          one diagonal (Coulomb-layer) phase op                         32 * 2^n B
  tile_W        the separable basis change W, per fused tile launch           32 * 2^n B per launch
  h_apply       out = H psi fused with <psi|H|psi>  (K2)                       32 * 2^n B
  screening     the whole pool, per gradient (K3)                             4 * 2^n  B per gradient

and print achieved GB/s and the fraction of the measured HBM copy peak (MEASURED_PEAKS.json).
States above ~126 MB (n >= 23) do not fit in L2, so those rows are true HBM numbers.

  python tools/sweep_roofline.py [--lattices 2x5,3x4,2x7] [--json out.json]
"""
import argparse
import json
import os
import sys
import time

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "quantum-simulation-of-fermi-hubbard-model_b200")]

import numpy as np  # noqa: E402

from fhsim.backend import Context, DevicePool, DeviceTable, State  # noqa: E402
from fhsim.circuit import Circuit  # noqa: E402
from fhsim.symbolic import fermi_hubbard, jordan_wigner  # noqa: E402
from fhsim.tables import GeneratorPlan, PauliTable  # noqa: E402
from operators.pool import hubbard_interaction_pool_simplified  # noqa: E402
from operators.tools import get_interacting_term  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(R, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def time_single(ctx, n, build, reps, fuse=False):
    c = Circuit(n, 0)
    build(c)
    prog = c.compile(ctx, fuse=fuse)
    st = State(ctx, n)
    st.set_basis(1)
    prog.run(st, [])
    prog.time_items(st, 0, prog.n_items, False, 2)
    ms = prog.time_items(st, 0, prog.n_items, False, reps)
    items = prog.n_items
    prog.close()
    st.close()
    return ms * 1e-3, items


def measure_lattice(ctx, lat, pk, pool_cap=400, verbose=True):
    """Time every kernel class on one lattice; returns the result row (see module docstring for the byte counts)."""
    nx, ny = (int(v) for v in lat.split("x"))
    n = 2 * nx * ny
    dim = 1 << n
    t0 = time.time()
    h_tab = PauliTable.from_operator(fermi_hubbard(nx, ny, 1.0, 4.0), n)
    pool_ops = hubbard_interaction_pool_simplified(nx, ny)
    plans = [GeneratorPlan(jordan_wigner(g), n) for g in pool_ops[:pool_cap]]
    coulomb = GeneratorPlan(jordan_wigner(get_interacting_term(fermi_hubbard(nx, ny, 1.0, 4.0))), n)
    reps = 20 if n <= 24 else (8 if n <= 28 else 3)
    res = {"lattice": lat, "n": n, "state_MiB": dim * 16 / 2 ** 20, "pool": len(pool_ops), "h_terms": len(h_tab),
           "h_groups": h_tab.n_groups, "host_compile_s": round(time.time() - t0, 2)}

    def rec(name, seconds, alg_bytes, extra=None):
        gbs = alg_bytes / seconds / 1e9
        res[name] = {"us": seconds * 1e6, "GBps": gbs, "frac": gbs / pk}
        if extra:
            res[name].update(extra)

    mid = plans[len(plans) // 2]
    s, _ = time_single(ctx, n, lambda c: c.generator(mid, angle=0.3), reps)
    rec("pair_fermi4", s, 4.0 * dim)
    x = (1 << (n - 1)) | (1 << (n // 2)) | 0b110
    z = (1 << (n - 2)) | 0b011
    s, _ = time_single(ctx, n, lambda c: c.pauli_rotation(x, z, 0.5, angle=0.7), reps)
    rec("pair_dense", s, 32.0 * dim)
    s, _ = time_single(ctx, n, lambda c: c.single_excitation(0.4, n // 2, n // 2 + 1), reps)
    rec("givens", s, 16.0 * dim)
    s, _ = time_single(ctx, n, lambda c: c.generator(coulomb, angle=0.2), reps)
    rec("diag_coulomb", s, 32.0 * dim, {"terms": sum(len(p.z) for p in coulomb.pieces)})
    s, items = time_single(ctx, n, lambda c: c.basis_change_separable(nx, ny), reps, fuse=True)
    rec("tile_W", s / items, 32.0 * dim, {"launches": items, "total_us": s * 1e6})

    # K2 and K3 on a generic (dense) state
    psi, lam = State(ctx, n), State(ctx, n)
    c = Circuit(n, 0)
    for q in range(n):
        c.ry(0.3 + 0.1 * q, q)
    prog = c.compile(ctx, fuse=True)
    psi.set_basis(0)
    prog.run(psi, [])
    prog.close()
    dtab = DeviceTable(ctx, h_tab)
    dtab.apply(psi, lam)
    ts = []
    for _ in range(max(3, reps // 2)):
        ctx.timer_start()
        dtab.apply(psi, lam)
        ts.append(ctx.timer_stop())
    rec("h_apply", min(ts) * 1e-3, 32.0 * dim)
    dpool = DevicePool(ctx, plans, n)
    dpool.gradients(psi, lam)
    ts = []
    for _ in range(3):
        ctx.timer_start()
        dpool.enqueue(psi, lam)
        ts.append(ctx.timer_stop())
    rec("screening", min(ts) * 1e-3, 4.0 * dim * len(plans), {"ops": len(plans), "per_gradient_us": min(ts) * 1e3 / len(plans)})
    # K2 on the sector-compressed copy (compress + gather on the compressed vector + memset / scatter of H psi)
    try:
        ns_ = nx * ny
        nu_, nd_ = (ns_ + 1) // 2, ns_ - (ns_ + 1) // 2
        dtab.apply_sector(psi, lam, nu_, nd_, enqueue_only=True)
        ts = []
        for _ in range(max(3, reps // 2)):
            ctx.timer_start()
            dtab.apply_sector(psi, lam, nu_, nd_, enqueue_only=True)
            ts.append(ctx.timer_stop())
        rec("h_apply_sector", min(ts) * 1e-3, 32.0 * dim)
    except Exception as exc:
        res["h_apply_sector"] = {"us": float("nan"), "GBps": float("nan"), "frac": float("nan"), "error": repr(exc)}
    # K3 on sector-compressed copies (what fh_program_evaluate does for number-conserving evaluations): same pool, the
    # half-filled sector; 4 * 2^n B per gradient is the FULL-SPACE algorithmic figure, so this row is an effective rate
    try:
        ns = nx * ny
        n_up = (ns + 1) // 2
        n_dn = ns - n_up
        dpool.gradients_sector(psi, lam, n_up, n_dn, enqueue_only=True)
        ts = []
        for _ in range(3):
            ctx.timer_start()
            dpool.gradients_sector(psi, lam, n_up, n_dn, enqueue_only=True)
            ts.append(ctx.timer_stop())
        from math import comb
        rec("screening_sector", min(ts) * 1e-3, 4.0 * dim * len(plans),
            {"ops": len(plans), "per_gradient_us": min(ts) * 1e3 / len(plans), "sector_dim": comb(ns, n_up) * comb(ns, n_dn)})
    except Exception as exc:                    # sector too large for the 12-bit list fields (more than 4096 patterns per spin)
        res["screening_sector"] = {"us": float("nan"), "GBps": float("nan"), "frac": float("nan"), "error": repr(exc)}
    for o in (dpool, dtab, psi, lam):
        o.close()
    if verbose:
        print(f"--- {lat}  n={n}  state={res['state_MiB']:.0f} MiB  pool={res['pool']}  H terms/groups={res['h_terms']}/{res['h_groups']}")
        for k in ("pair_fermi4", "pair_dense", "givens", "diag_coulomb", "tile_W", "h_apply", "h_apply_sector", "screening", "screening_sector"):
            v = res[k]
            print(f"    {k:13s} {v['us']:12.1f} us  {v['GBps']:9.1f} GB/s  {100 * v['frac']:6.1f} % of {pk:.0f}")
        sys.stdout.flush()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lattices", default="3x3,2x5,2x6,3x4,2x7")
    ap.add_argument("--json", default=None)
    ap.add_argument("--pool-cap", type=int, default=400, help="screen at most this many pool operators")
    args = ap.parse_args()
    ctx = Context(0)
    pk = peak()
    rows = [measure_lattice(ctx, lat, pk, args.pool_cap) for lat in args.lattices.split(",")]
    if args.json:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        json.dump({"peak_GBps": pk, "rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
