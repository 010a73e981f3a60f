"""Host-side logic behind the sector-compressed paths (csrc/sector_eval.cu, fhsim/sharded.py): which operators and circuits
conserve (N_up, N_dn), the sector a slab of a sharded state holds, and the identity the dense tail relies on,
W = S (U_up x U_dn) S with S the interleaved -> blocked Jordan-Wigner sign.  CPU only (numpy interpreter of the circuits)."""
import itertools
from math import comb

import numpy as np
import pytest

import emulate
from fhsim.circuit import Circuit, DiagOpSpec, Marker
from fhsim.sharded import QubitLayout, conserves_species, local_sector
from fhsim.symbolic import fermi_hubbard, jordan_wigner
from fhsim.tables import GeneratorPlan, PauliTable
from operators.pool import hubbard_interaction_pool_simplified


def test_which_tables_conserve_both_particle_numbers():
    for lat in ((2, 2), (2, 3), (3, 3)):
        n = 2 * lat[0] * lat[1]
        assert PauliTable.from_operator(fermi_hubbard(*lat, 1.0, 4.0), n).conserves_species()
    assert not PauliTable(4, [0b0011], [0], [1.0]).conserves_species()                               # X X alone
    assert PauliTable(4, [0b0101, 0b0101], [0, 0b0101], [0.5, 0.5]).conserves_species()              # hop inside one species
    assert not PauliTable(4, [0b0011, 0b0011], [0, 0b0011], [0.5, 0.5]).conserves_species()          # hop between the species
    assert PauliTable(4, [0, 0], [0b0011, 0b1000], [1.0, -0.5]).conserves_species()                  # diagonal
    assert not PauliTable(6, [0b111110], [0], [1.0]).conserves_species()                             # five X: not analysed


def test_which_circuits_conserve_both_particle_numbers():
    from fhsim.symbolic import givens_decomposition_square
    from operators.fourier import fourier_transform_matrix
    nx, ny, n = 2, 3, 12
    plans = [GeneratorPlan(jordan_wigner(g), n) for g in hubbard_interaction_pool_simplified(nx, ny)]
    c = Circuit(n, 2)
    c.generator(plans[3], param=0)
    c.generator(plans[40], param=1)
    c.basis_change_separable(nx, ny)
    ops = [o for o in c.ops if not isinstance(o, Marker)]
    assert conserves_species(ops, n)
    r = Circuit(n, 0)
    dec, diag = givens_decomposition_square(fourier_transform_matrix(nx, ny))
    r.basis_change(diag, list(reversed(dec)))          # the reference network rotates between the up and down orbital of a site
    assert not conserves_species([o for o in r.ops if not isinstance(o, Marker)], n)
    x = Circuit(n, 0)
    x.rx(0.3, 1)
    assert not conserves_species(x.ops, n)


@pytest.mark.parametrize("n,g,n_up,n_dn,seed", [(8, 1, 2, 2, 0), (12, 2, 3, 3, 1), (12, 3, 4, 2, 2), (16, 3, 4, 4, 3)])
def test_the_slabs_of_a_sharded_state_partition_the_sector(n, g, n_up, n_dn, seed):
    """Over all ranks, the local sectors (species masks of the local bits + electrons left after the rank bits) tile the global
    (N_up, N_dn) sector exactly: dimensions add up and every sector index lands in exactly one slab's local sector."""
    rng = np.random.default_rng(seed)
    layout = QubitLayout(n, g, [int(v) for v in rng.permutation(n)])
    nl = n - g
    up = sum(1 << b for b in range(n) if (n - 1 - b) % 2 == 0)
    total = 0
    for rank in range(1 << g):
        loc = local_sector(layout, rank, n_up, n_dn)
        if loc is None:
            continue
        um, dm, lu, ld = loc
        assert um & dm == 0 and um | dm == (1 << nl) - 1
        total += comb(bin(um).count("1"), lu) * comb(bin(dm).count("1"), ld)
    assert total == comb(n // 2, n_up) * comb(n // 2, n_dn)
    for _ in range(50):                                   # a random sector index: its slab reports the matching local occupation
        ups = rng.choice([b for b in range(n) if up >> b & 1], size=n_up, replace=False)
        dns = rng.choice([b for b in range(n) if not up >> b & 1], size=n_dn, replace=False)
        idx = sum(1 << int(b) for b in itertools.chain(ups, dns))
        p = layout.phys(idx)
        rank, local = p >> nl, p & ((1 << nl) - 1)
        um, dm, lu, ld = local_sector(layout, rank, n_up, n_dn)
        assert bin(local & um).count("1") == lu and bin(local & dm).count("1") == ld


@pytest.mark.parametrize("lat,n_up,n_dn", [((2, 2), 2, 2), ((2, 3), 3, 3), ((2, 3), 4, 1)])
def test_basis_change_factorises_into_two_dense_sector_blocks(lat, n_up, n_dn):
    """W restricted to the (N_up, N_dn) sector equals S o (U_up (S o Psi) U_dn^T) with U_up / U_dn the action of the up-only /
    down-only ops on one species alone and S(u, d) = (-1)^#{(i, j): up orbital i and down orbital j occupied, j <= i}: what
    fh_sector_dense_prepare builds and k_sector_gemm applies (csrc/sector_eval.cu)."""
    nx, ny = lat
    n, half = 2 * nx * ny, nx * ny
    c = Circuit(n, 0)
    c.basis_change_separable(nx, ny)
    ops = [o for o in c.ops if not isinstance(o, Marker)]
    upmask = sum(1 << b for b in range(n) if b % 2 == 1)
    dnmask = sum(1 << b for b in range(n) if b % 2 == 0)

    def species(o):
        bits = 0
        for z in (o.z if isinstance(o, DiagOpSpec) else [o.x]):
            bits |= z
        return "U" if bits & upmask and not bits & dnmask else ("D" if bits & dnmask and not bits & upmask else "M")
    assert set(species(o) for o in ops) <= {"U", "D"}

    def deposit(cfg, mask):
        out, k = 0, 0
        for b in range(n):
            if mask >> b & 1:
                out |= ((cfg >> k) & 1) << b
                k += 1
        return out
    ucfgs = [u for u in range(1 << half) if bin(u).count("1") == n_up]
    dcfgs = [d for d in range(1 << half) if bin(d).count("1") == n_dn]

    def block(kind, cfgs, mask):
        sub = Circuit(n, 0)
        sub.ops = [o for o in ops if species(o) == kind]
        mat = np.zeros((len(cfgs), len(cfgs)), complex)
        for a, cfg in enumerate(cfgs):
            v = np.zeros(1 << n, complex)
            v[deposit(cfg, mask)] = 1.0
            w = emulate.run_circuit(sub, v, [])
            mat[:, a] = [w[deposit(c2, mask)] for c2 in cfgs]
        return mat
    u_up, u_dn = block("U", ucfgs, upmask), block("D", dcfgs, dnmask)
    sign = np.ones((len(ucfgs), len(dcfgs)))
    for a, u in enumerate(ucfgs):
        for b, d in enumerate(dcfgs):
            cnt = sum(bin(d & ((2 << i) - 1)).count("1") for i in range(half) if u >> i & 1)
            sign[a, b] = -1.0 if cnt & 1 else 1.0
    rng = np.random.default_rng(0)
    psi_mat = rng.normal(size=sign.shape) + 1j * rng.normal(size=sign.shape)
    psi = np.zeros(1 << n, complex)
    for a, u in enumerate(ucfgs):
        for b, d in enumerate(dcfgs):
            psi[deposit(u, upmask) | deposit(d, dnmask)] = psi_mat[a, b]
    full = Circuit(n, 0)
    full.ops = list(ops)
    out = emulate.run_circuit(full, psi.copy(), [])
    got = np.array([[out[deposit(u, upmask) | deposit(d, dnmask)] for d in dcfgs] for u in ucfgs])
    pred = sign * (u_up @ (sign * psi_mat) @ u_dn.T)
    assert np.abs(pred - got).max() < 1e-12
    assert abs(np.linalg.norm(out) - np.linalg.norm(psi)) < 1e-12          # and W keeps the state inside the sector
