"""Sector-resident evaluation (csrc/sector_eval.cu) against the full-space path on the same programs: energies, pool
gradients, device times.  Run on a GPU box:  python tools/check_sector.py [steps]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "quantum-simulation-of-fermi-hubbard-model_b200"))
import bench  # noqa: E402


def both(prog, basis, thetas, tabs, **kw):
    os.environ.pop("FHSIM_NO_SECTOR", None)
    a = prog.evaluate(basis, thetas, tabs, **kw)
    info = prog.sector_info()
    ms_a = prog.last_stats()
    os.environ["FHSIM_NO_SECTOR"] = "1"
    b = prog.evaluate(basis, thetas, tabs, **kw)
    ms_b = prog.last_stats()
    info_b = prog.sector_info()
    os.environ.pop("FHSIM_NO_SECTOR", None)
    assert not info_b["active"]
    return a, b, info, ms_a, ms_b


def main():
    os.environ["FHSIM_SECTOR"] = "1"      # every size (the default limits the path to small sectors)
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    from fhsim.backend import Context
    ctx = Context(0)
    out = {}
    for (nx, ny, nops) in ((2, 2, 3), (2, 3, 8), (3, 3, 20)):
        wl = bench.build_lattice_workload(ctx, nx, ny, 4.0, nops)
        prog, m = wl["prog"], wl["prog"].markers["ansatz_end"]
        a, b, info, ms_a, ms_b = both(prog, wl["basis"], wl["thetas"], [wl["dtab"]], pool=wl["dpool"], pool_pos=m)
        de = abs(a["expvals"][0] - b["expvals"][0])
        dg = float(np.abs(a["pool"] - b["pool"]).max())
        print(f"{nx}x{ny}: sector {info}  dE={de:.2e} dpool={dg:.2e} |g|max={np.abs(b['pool']).max():.4f} "
              f"ms sector {ms_a[0]:.4f} ({ms_a[1]} launches) full {ms_b[0]:.4f} ({ms_b[1]} launches)", flush=True)
        a2, b2, info2, _, _ = both(prog, wl["basis"], wl["thetas"], [wl["dtab"]])
        print(f"   energy only: active={info2['active']} dE={abs(a2['expvals'][0] - b2['expvals'][0]):.2e}", flush=True)
        # pool at position 0 and a sub-range
        a3, b3, info3, _, _ = both(prog, wl["basis"], wl["thetas"], [wl["dtab"]], pool=wl["dpool"], pool_pos=0,
                                   pool_range=(1, max(1, wl["dpool"].n_out // 2)))
        print(f"   pool_pos=0, sub-range: active={info3['active']} dpool={float(np.abs(a3['pool'] - b3['pool']).max()):.2e}", flush=True)
        out[f"{nx}x{ny}"] = dict(info=info, dE=de, dpool=dg)
        for o in (prog, wl["dpool"], wl["dtab"]):
            o.close()

    wl = bench.build_gpu_workload(ctx)
    prog, m = wl["prog"], wl["prog"].markers["ansatz_end"]
    a, b, info, ms_a, ms_b = both(prog, wl["basis"], wl["thetas"], [wl["dtab"]], pool=wl["dpool"], pool_pos=m)
    de = abs(a["expvals"][0] - b["expvals"][0])
    dg = float(np.abs(a["pool"] - b["pool"]).max())
    print(f"bench cfg3: sector {info} dE={de:.2e} dpool={dg:.2e} E={a['expvals'][0]:.12f}", flush=True)
    res = {}
    for label in ("sector", "full"):
        if label == "full":
            os.environ["FHSIM_NO_SECTOR"] = "1"
        else:
            os.environ.pop("FHSIM_NO_SECTOR", None)
        for _ in range(5):
            prog.evaluate(wl["basis"], wl["thetas"], [wl["dtab"]], pool=wl["dpool"], pool_pos=m)
        dev, wall = [], []
        for _ in range(steps):
            ctx.flush_l2(bench.L2_FLUSH_BYTES)
            ctx.sync()
            t0 = time.perf_counter()
            prog.evaluate(wl["basis"], wl["thetas"], [wl["dtab"]], pool=wl["dpool"], pool_pos=m)
            wall.append(time.perf_counter() - t0)
            dev.append(prog.last_stats()[0])
        res[label] = dict(dev_ms=float(np.mean(dev)), dev_ms_min=float(np.min(dev)), e2e_ms=float(1e3 * np.mean(wall)),
                          launches=prog.last_stats()[1])
        print(label, res[label], flush=True)
    os.environ.pop("FHSIM_NO_SECTOR", None)
    out["bench"] = dict(info=info, dE=de, dpool=dg, timing=res)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
