#!/usr/bin/env python
"""In-kernel phase timeline of the fused tile kernel (k_tile) for every launch item of the bench circuit.

Builds a SEPARATE library with -DFH_TILE_TIMELINE (csrc/build_timeline/libfhsim_timeline.so; the shipped
fhsim/lib/libfhsim.so has the instrumentation compiled out), loads it in place of the shipped one and prints, per launch item, the
clock64() deltas thread 0 of CTA 0 stamped at the phase boundaries of tile_run (cycles at the SM clock):

    python tools/probe_timeline.py            # needs a GPU; the build step also works without one

Round-1 result at 18 qubits (T = 11, 512 threads, 1.965 GHz): prologue 2 100-2 500, tile load 2 900, ~700 per fermionic
double excitation, 900-1 000 per Givens of W, store 1 300 -- see DESIGN.md 7.
"""
import ctypes as C
import os
import subprocess
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(R, "quantum-simulation-of-fermi-hubbard-model_b200")
CSRC = os.path.join(PKG, "csrc")
VARIANT = os.environ.get("FH_OPLOOP_VARIANT", "0")         # see tile_tma.cu: strips the op loop piece by piece
EXTRA = os.environ.get("FH_PROBE_DEFS", "").split()      # extra -D flags, e.g. FH_PROBE_DEFS="-DPAIR_UNROLL=2"
TAG = "".join(c if c.isalnum() else "_" for c in "".join(EXTRA))
OUT = os.path.join(CSRC, "build_timeline", f"libfhsim_timeline_v{VARIANT}{TAG}.so")


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    srcs = ["kernels.cu", "tile_tma.cu", "api.cu", "program.cu", "lanczos.cu", "comm.cu", "dress.cu"]
    cmd = ["nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
           "-DFH_TILE_TIMELINE", f"-DFH_OPLOOP_VARIANT={VARIANT}", *EXTRA, "-shared", "-cudart", "static", "-o", OUT] + [os.path.join(CSRC, f) for f in srcs]
    subprocess.check_call(cmd)


def main():
    if not os.path.exists(OUT) or "--rebuild" in sys.argv:
        build()
    if "--build-only" in sys.argv:
        return
    sys.path[:0] = [R, PKG]
    from fhsim import _cabi
    _cabi.LIB_PATH = OUT                      # load the instrumented build instead of the shipped library
    import bench
    from fhsim.backend import Context, State
    ctx = Context(0)
    import json
    picks = json.load(open(os.path.join(R, "tools", "probes", "picks_3x3.json"))) if VARIANT != "0" else None
    wl = bench.build_gpu_workload(ctx, picks)      # stripped variants cannot screen: operator picks from the oracle
    prog, n = wl["prog"], bench.N_QUBITS
    st = State(ctx, n)
    st.set_basis(wl["basis"])
    lib = _cabi.lib()
    buf = (C.c_longlong * 64)()
    for item in range(prog.n_items):
        for _ in range(3):
            prog.run(st, wl["thetas"], item, 1)
        ctx.sync()
        tma = not os.environ.get("FHSIM_TILE_LDG")
        fn = lib.fh_debug_tile_tma_timeline if tma else lib.fh_debug_tile_timeline
        if fn(buf) != 0:
            raise RuntimeError("timeline read-back failed")
        t = list(buf)
        nops = sum(1 for k in range(4, 44) if t[2] <= t[k] <= t[3])
        marks = [t[4 + k] for k in range(nops)] + [t[3]]
        per_op = [marks[k + 1] - marks[k] for k in range(nops)]
        tail = f"store issue {t[62] - t[3]:6d}  store done {t[63] - t[62]:6d}" if tma else f"store {t[63] - t[3]:6d}"
        print(f"item {item:2d}: prologue {t[1] - t[0]:6d}  load {t[2] - t[1]:6d}  ops({nops}) {t[3] - t[2]:6d} {per_op}  "
              f"{tail}  total {t[63] - t[0]:6d} cycles   [{1e3 * prog.time_items(st, item, 1, False, 20):.2f} us/launch back to back]",
              flush=True)


if __name__ == "__main__":
    main()
