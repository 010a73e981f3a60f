import sys; import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0]=[R, os.path.join(R,'quantum-simulation-of-fermi-hubbard-model_b200')]
import numpy as np, bench
from fhsim.backend import Context, State
ctx=Context(0)
wl=bench.build_gpu_workload(ctx)
prog=wl['prog']; n=18
st=State(ctx,n); st.set_basis(wl['basis'])
prog.run(st, wl['thetas'])
print('items', prog.n_items, 'marker', prog.markers)
for i in range(prog.n_items):
    print(i, 'fwd %.2f us' % (1e3*prog.time_items(st, i, 1, False, 50)), 'dag %.2f us' % (1e3*prog.time_items(st, i, 1, True, 50)))
print('all fwd %.2f us' % (1e3*prog.time_items(st, 0, prog.n_items, False, 20)))
res=prog.evaluate(wl['basis'], wl['thetas'], [wl['dtab']], pool=wl['dpool'], pool_pos=prog.markers['ansatz_end'])
print('eval ms', prog.last_stats())

# ---- synthetic slope test: one tile with K identical sub-ops ---------------------------------
import fhsim.circuit as fc
from fhsim.circuit import Circuit
fc._LAUNCH_BYTES = 1e12
plans = wl['plans']
for kind in ('fermi4', 'givens', 'rz', 'ry'):
    row = []
    for K in (1, 8, 32, 64):
        c = Circuit(n, 0)
        for k in range(K):
            if kind == 'fermi4':
                c.generator(plans[wl['picks'][0]], angle=0.01 * (k + 1))
            elif kind == 'givens':
                c.fermionic_single_excitation(0.1 * (k + 1), 3, 9)
            elif kind == 'rz':
                c.rz(0.1, 5); c.ry(0.1, 5)      # the ry stops the phases from being merged away
            else:
                c.ry(0.1 * (k + 1), 5)
        pr = c.compile(ctx)
        s2 = State(ctx, n); s2.set_basis(5)
        pr.run(s2, [])
        row.append((K, pr.n_items, round(1e3 * pr.time_items(s2, 0, pr.n_items, False, 50), 2)))
        pr.close()
    print(kind, row)
