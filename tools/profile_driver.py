#!/usr/bin/env python
"""cProfile of the drop-in ADAPT drivers' optimiser loops (host overhead around fh_program_evaluate).
   python tools/profile_driver.py [2x2|3x3]"""
import contextlib, cProfile, io, os, pstats, sys, tempfile, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "quantum-simulation-of-fermi-hubbard-model_b200")]
which = sys.argv[1] if len(sys.argv) > 1 else "2x2"
os.chdir(tempfile.mkdtemp(prefix="fhsim_prof_"))
import torch
_w = torch.nn.Parameter(torch.zeros(2))          # pay torch's lazy optimizer imports (~2.5 s) outside the profile
_o = torch.optim.Adam([_w], lr=1e-3)
_w.sum().backward()
_o.step()
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    if which == "2x2":
        from models.adapt_vqe import ADAPT
        vqe = ADAPT(n_epoch=6, threshold1=1e-2, threshold2=1e-2, x_dimension=2, y_dimension=2, n_electrons=4,
                    n_spin_up=2, n_spin_down=2, tunneling=1, coulomb=4, verbose=False)
    else:
        from models.adapt_vqe_for_3x3 import ADAPT
        vqe = ADAPT(n_epoch=2, threshold1=1e-2, threshold2=5e-2, x_dimension=3, y_dimension=3, n_electrons=9,
                    n_spin_up=5, n_spin_down=4, tunneling=1, coulomb=6, verbose=False)
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    vqe.run()
    pr.disable()
    dt = time.perf_counter() - t0
its = len(vqe.results["iteration loss"])
print(f"{which}: {its} optimiser iterations in {dt:.3f} s = {1e3 * dt / its:.3f} ms per iteration (under cProfile)")
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(40)
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30)
print(s.getvalue()[:20000])
