"""Matrix-level pin of the host tables, independent of BOTH in-house symbolic layers.

tests/golden/reference_host_tables.json comes from the reference's own loops executed on the product's mini-OpenFermion
(fhsim.symbolic), and oracle/pauli.py is a second in-house restatement: two restatements agreeing does not pin OpenFermion
semantics.  Here every operator is rebuilt as an explicit sparse matrix from Jordan-Wigner ladder operators written as
Kronecker products (a_j = Z^{(x)j} (x) [[0,1],[0,0]] (x) 1..., wire 0 = most significant), with

* H   = -t sum_<ij>,s (a+_is a_js + h.c.) + U sum_i n_iup n_idn           (openfermion.fermi_hubbard, periodic, the
        2-site-dimension bond de-duplication of SURVEY A.1 written out as an explicit bond list),
* G_k = i (a+_i1 a+_i2 a_i3 a_i4 - a+_i3 a+_i4 a_i1 a_i2)                (reference operators/pool.py:233-253: same loop
        order, first-seen-of-+-op kept, de-duplicated on the MATRIX),

and compared with the dense form of the packed (x, z, coeff) tables expanded by Kronecker products of 2x2 Pauli
matrices.  Nothing from fhsim.symbolic / oracle.pauli takes part in building the expected matrices.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from fhsim.symbolic import fermi_hubbard, jordan_wigner
from fhsim.tables import PauliTable
from operators.pool import hubbard_interaction_pool_simplified
from oracle import pauli

I2 = sp.identity(2, format="csr", dtype=complex)
Z2 = sp.csr_matrix(np.diag([1.0, -1.0]).astype(complex))
X2 = sp.csr_matrix(np.array([[0, 1], [1, 0]], dtype=complex))
Y2 = sp.csr_matrix(np.array([[0, -1j], [1j, 0]], dtype=complex))
LOWER = sp.csr_matrix(np.array([[0, 1], [0, 0]], dtype=complex))      # |0><1|: annihilates an occupied orbital


def _kron(mats):
    out = mats[0]
    for m in mats[1:]:
        out = sp.kron(out, m, format="csr")
    return out


def ladder(n):
    a = [_kron([Z2] * j + [LOWER] + [I2] * (n - 1 - j)) for j in range(n)]
    return a, [m.conj().T.tocsr() for m in a]


def table_matrix(xs, zs, cs, n):
    """sum_t c_t P_t with P_t = kron over wires of {1, X, Y, Z}; wire q <-> bit n-1-q of the masks."""
    total = sp.csr_matrix((1 << n, 1 << n), dtype=complex)
    for x, z, c in zip(xs, zs, cs):
        mats = []
        for q in range(n):
            b = 1 << (n - 1 - q)
            mats.append(Y2 if (x & b and z & b) else X2 if x & b else Z2 if z & b else I2)
        total = total + complex(c) * _kron(mats)
    return total


def bonds(nx, ny):
    """Nearest-neighbour bonds of the periodic nx x ny lattice, each once (a length-2 dimension has ONE bond per pair)."""
    out = set()
    for y in range(ny):
        for x in range(nx):
            s = x + y * nx
            if nx > 1:
                out.add(tuple(sorted((s, (x + 1) % nx + y * nx))))
            if ny > 1:
                out.add(tuple(sorted((s, x + ((y + 1) % ny) * nx))))
    return sorted(b for b in out if b[0] != b[1])


def hubbard_matrix(nx, ny, t, u):
    n = 2 * nx * ny
    a, ad = ladder(n)
    h = sp.csr_matrix((1 << n, 1 << n), dtype=complex)
    for (i, j) in bonds(nx, ny):
        for s in (0, 1):
            p, q = 2 * i + s, 2 * j + s
            h = h - t * (ad[p] @ a[q] + ad[q] @ a[p])
    for i in range(nx * ny):
        h = h + u * (ad[2 * i] @ a[2 * i] @ ad[2 * i + 1] @ a[2 * i + 1])
    return h


def pool_matrices(nx, ny):
    n = 2 * nx * ny
    a, ad = ladder(n)
    n_sites = nx * ny
    rng = np.random.default_rng(5)
    probe = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    kept, prints = [], []
    for spin in (0, 1):
        for k1 in range(n_sites):
            for k2 in range(n_sites):
                for q in range(1, n_sites):
                    kx1, ky1, kx2, ky2, qx, qy = k1 % nx, k1 // nx, k2 % nx, k2 // nx, q % nx, q // nx
                    i1 = 2 * ((kx1 + qx) % nx + ((ky1 + qy) % ny) * nx) + spin
                    i2 = 2 * ((kx2 - qx) % nx + ((ky2 - qy) % ny) * nx) + (spin ^ 1)
                    i3 = 2 * (kx2 + ky2 * nx) + (spin ^ 1)
                    i4 = 2 * (kx1 + ky1 * nx) + spin
                    g = 1j * (ad[i1] @ ad[i2] @ a[i3] @ a[i4]) - 1j * (ad[i3] @ ad[i4] @ a[i1] @ a[i2])
                    fp = g @ probe                                           # fingerprint: G is fixed by its action
                    if any(np.abs(fp - f).max() < 1e-12 or np.abs(fp + f).max() < 1e-12 for f in prints):
                        continue
                    kept.append(g.tocsr())
                    prints.append(fp)
    return kept


def _maxabs(m):
    m = sp.csr_matrix(m)
    return 0.0 if m.nnz == 0 else float(np.abs(m.data).max())


@pytest.mark.parametrize("nx,ny,u", [(2, 2, 4.0), (2, 3, 4.0)])
def test_hamiltonian_tables_equal_the_ladder_operator_matrix(nx, ny, u):
    n = 2 * nx * ny
    want = hubbard_matrix(nx, ny, 1.0, u)
    assert _maxabs(want - want.conj().T) < 1e-14
    tab = PauliTable.from_operator(fermi_hubbard(nx, ny, 1.0, u), n)
    got = table_matrix([int(v) for v in tab.x], [int(v) for v in tab.z], tab.coeff, n)
    assert _maxabs(got - want) < 1e-12
    o_h = pauli.compress(pauli.jw_table(pauli.hubbard_fermion_terms(nx, ny, 1.0, u), n))
    got_o = table_matrix([k[0] for k in o_h], [k[1] for k in o_h], list(o_h.values()), n)
    assert _maxabs(got_o - want) < 1e-12


def test_2x2_ground_energy_of_the_ladder_operator_matrix():
    """the textbook 4-site-ring value, in the (2 up, 2 down) sector, straight from the Kronecker-product matrix."""
    n = 8
    h = hubbard_matrix(2, 2, 1.0, 4.0).toarray()
    idx = [i for i in range(1 << n)
           if bin(i & 0b10101010).count("1") == 2 and bin(i & 0b01010101).count("1") == 2]
    vals = np.linalg.eigvalsh(h[np.ix_(idx, idx)])
    assert abs(vals[0] + 2.1027484835) < 1e-9


@pytest.mark.parametrize("nx,ny,size", [(2, 2, 24), (2, 3, 90)])
def test_pool_tables_equal_the_ladder_operator_matrices(nx, ny, size):
    n = 2 * nx * ny
    want = pool_matrices(nx, ny)
    assert len(want) == size
    product = [PauliTable.from_operator(jordan_wigner(g), n, compress=False)
               for g in hubbard_interaction_pool_simplified(nx, ny)]
    oracle = [pauli.jw_table(op, n) for op in pauli.pool_fermion_terms(nx, ny)]
    assert len(product) == size and len(oracle) == size
    for k in range(size):
        assert _maxabs(want[k] - want[k].conj().T) < 1e-14                   # Hermitian generator
        got = table_matrix([int(v) for v in product[k].x], [int(v) for v in product[k].z], product[k].coeff, n)
        assert _maxabs(got - want[k]) < 1e-12, f"product pool operator {k}"
        keys = list(oracle[k])
        got_o = table_matrix([x for x, _ in keys], [z for _, z in keys], [oracle[k][kk] for kk in keys], n)
        assert _maxabs(got_o - want[k]) < 1e-12, f"oracle pool operator {k}"
