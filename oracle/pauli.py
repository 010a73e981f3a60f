"""ORACLE (test infrastructure only -- never imported by the product path).

Packed-mask restatement of the host compile step of the reference:
Jordan-Wigner (OpenFermion ``jordan_wigner``, called at reference ``models/adapt_vqe.py:143,166``),
``fermi_hubbard`` (``adapt_vqe.py:159``), the operator pool
(``operators/pool.py:220-255``) and the Givens network of ``fourier_transform_matrix``
(``operators/fourier.py:13-37`` + OpenFermion ``givens_decomposition_square``, ``adapt_vqe.py:187``).

OpenFermion / PennyLane are third-party dependencies that are NOT vendored in /root/reference and
have no pinned version (no requirements/lock file); their published algorithms are restated here
on integers instead of on symbolic tuples, i.e. by a different route than the product's
``fhsim.symbolic``.  PARITY UNPINNED by the reference (it ships no tests / golden vectors): the
oracle is pinned instead to analytic known answers (tests/test_oracle_known_answers.py).

Conventions: wire/orbital q <-> bit (n-1-q) of the flat index (wire 0 = MSB, PennyLane order,
reference ``linalg/exact_diagonalization.py:23``).  A Pauli string is (x, z) meaning
i^k X^x Z^z with k = popcount(x&z) mod 4 (each Y = iXZ) and
(P psi)[i] = i^k (-1)^popcount((i^x)&z) psi[i^x].
"""
from __future__ import annotations

import numpy as np

TOL = 1e-8


def bit(q: int, n: int) -> int:
    return 1 << (n - 1 - q)


def popcount(v: int) -> int:
    return bin(v).count("1")


# ---------------------------------------------------------------------------
# products of phase-tracked strings:  (x, z, k) == i^k X^x Z^z
# ---------------------------------------------------------------------------
def _mul(a, b):
    x1, z1, k1 = a
    x2, z2, k2 = b
    return (x1 ^ x2, z1 ^ z2, (k1 + k2 + 2 * popcount(z1 & x2)) & 3)


def _ladder(j: int, action: int, n: int):
    """a†_j = Z_{<j}(X_j - iY_j)/2, a_j = Z_{<j}(X_j + iY_j)/2 -> [(string, coeff)], X part first."""
    zs = 0
    for q in range(j):
        zs |= bit(q, n)
    bj = bit(j, n)
    xs = (bj, zs, 0)
    ys = (bj, zs | bj, 1)          # Y = i X Z
    return [(xs, 0.5), (ys, -0.5j if action else 0.5j)]


_I_POW = (1, 1j, -1, -1j)


def _canon(string, coeff):
    """(x,z,k),c -> key (x,z) and coefficient w.r.t. the Hermitian string i^popcount(x&z) X^x Z^z."""
    x, z, k = string
    extra = (k - popcount(x & z)) & 3
    return (x, z), coeff * _I_POW[extra]


def jw_table(fermion_terms, n: int):
    """[(term, coeff)] with term = ((index, 1|0), ...) -> insertion-ordered {(x, z): coeff}.

    Follows OpenFermion's ladder-by-ladder expansion: per fermion term start from the
    identity, multiply by (X-part + Y-part) left to right (left-term-major product order),
    then merge into the running sum, deleting entries that fall below 1e-8.
    """
    total = {}
    for term, coeff in fermion_terms:
        current = {(0, 0): ((0, 0, 0), coeff)}            # key -> (string, coeff)
        for index, action in term:
            lad = _ladder(index, action, n)
            nxt = {}
            for key, (s, c) in current.items():
                for ls, lc in lad:
                    prod = _mul(s, ls)
                    k2, c2 = _canon(prod, c * lc)
                    if k2 in nxt:
                        nxt[k2] = ((k2[0], k2[1], popcount(k2[0] & k2[1]) & 3), nxt[k2][1] + c2)
                    else:
                        nxt[k2] = ((k2[0], k2[1], popcount(k2[0] & k2[1]) & 3), c2)
            current = nxt
        for key, (s, c) in current.items():
            v = total.get(key, 0.0) + c
            total[key] = v
            if abs(v) < TOL:
                del total[key]
    return total


def compress(table: dict):
    out = {}
    for key, c in table.items():
        c = complex(c)
        if abs(c.imag) <= TOL:
            c = complex(c.real, 0.0)
        elif abs(c.real) <= TOL:
            c = complex(0.0, c.imag)
        if abs(c) > TOL:
            out[key] = c
    return out


def table_arrays(table: dict):
    """-> x[u64], z[u64], k[u8], coeff[c128] in dict order."""
    keys = list(table.keys())
    x = np.array([k[0] for k in keys], dtype=np.uint64)
    z = np.array([k[1] for k in keys], dtype=np.uint64)
    kk = np.array([popcount(k[0] & k[1]) & 3 for k in keys], dtype=np.uint8)
    c = np.array([complex(table[k]) for k in keys], dtype=np.complex128)
    return x, z, kk, c


# ---------------------------------------------------------------------------
# models
# ---------------------------------------------------------------------------
def hubbard_fermion_terms(nx, ny, t, u, periodic=True):
    """Spinful Hubbard model, site s = x + y*nx, orbital 2s+spin (OpenFermion ``fermi_hubbard``):
    right and bottom neighbour of each site; a periodic dimension of length 2 counts its bond once."""
    ns = nx * ny
    terms = []

    def hop(i, j):
        terms.append((((i, 1), (j, 0)), -t))
        terms.append((((j, 1), (i, 0)), -t))

    for s in range(ns):
        right = None
        if nx > 1:
            if (s + 1) % nx == 0:
                right = s + 1 - nx if periodic else None
            else:
                right = s + 1
        bottom = None
        if ny > 1:
            if s + nx + 1 > ns:
                bottom = s + nx - ns if periodic else None
            else:
                bottom = s + nx
        if nx == 2 and periodic and s % 2 == 1:
            right = None
        if ny == 2 and periodic and s >= nx:
            bottom = None
        if right is not None:
            hop(2 * s, 2 * right)
            hop(2 * s + 1, 2 * right + 1)
        if bottom is not None:
            hop(2 * s, 2 * bottom)
            hop(2 * s + 1, 2 * bottom + 1)
        terms.append((((2 * s, 1), (2 * s, 0), (2 * s + 1, 1), (2 * s + 1, 0)), u))
    # merge duplicates in first-seen order (FermionOperator +=)
    merged = {}
    for term, c in terms:
        merged[term] = merged.get(term, 0.0) + c
    return [(term, c) for term, c in merged.items() if abs(c) >= TOL]


def pool_fermion_terms(nx, ny):
    """reference operators/pool.py:220-255.  Each element: [(term, coeff), (term, coeff)] normal ordered.

    All four indices are distinct (q != 0), so normal ordering is a pure sign:
    a†_a a†_b a_c a_d -> (-1)^([a<b]+[c<d]) a†_max a†_min a_max' a_min'.
    """
    ns = nx * ny

    def t2i(ix, iy, spin):
        return 2 * (ix + iy * nx) + spin

    def ordered(a, b, c, d, coeff):
        sign = (-1) ** (int(a < b) + int(c < d))
        return ((max(a, b), 1), (min(a, b), 1), (max(c, d), 0), (min(c, d), 0)), sign * coeff

    pool, seen = [], set()
    for spin in (0, 1):
        for k1 in range(ns):
            for k2 in range(ns):
                for q in range(1, ns):
                    kx1, ky1 = k1 % nx, k1 // nx
                    kx2, ky2 = k2 % nx, k2 // nx
                    qx, qy = q % nx, q // nx
                    i1 = t2i((kx1 + qx) % nx, (ky1 + qy) % ny, spin)
                    i2 = t2i((kx2 - qx) % nx, (ky2 - qy) % ny, spin ^ 1)
                    i3 = t2i(kx2, ky2, spin ^ 1)
                    i4 = t2i(kx1, ky1, spin)
                    assert len({i1, i2, i3, i4}) == 4
                    ta, ca = ordered(i1, i2, i3, i4, 1j)
                    tb, cb = ordered(i3, i4, i1, i2, -1j)
                    # normal_ordered() accumulates term by term: order (ta, tb)
                    op = ((ta, ca), (tb, cb))
                    key_p = frozenset(op)
                    key_m = frozenset(((ta, -ca), (tb, -cb)))
                    if key_p in seen or key_m in seen:
                        continue
                    seen.add(key_p)
                    pool.append(list(op))
    return pool


def ft_matrix(nx, ny):
    """reference operators/fourier.py:13-37."""
    ns = nx * ny
    m = np.zeros((2 * ns, 2 * ns), dtype=complex)
    for row in range(2 * ns):
        for col in range(2 * ns):
            if row % 2 != col % 2:
                continue
            rx, ry = (row // 2) % nx, (row // 2) // nx
            cx, cy = (col // 2) % nx, (col // 2) // nx
            m[row, col] = np.exp(-2j * np.pi * cx * rx / nx) * np.exp(-2j * np.pi * cy * ry / ny)
    return m / np.sqrt(ns)


def givens_network(q):
    """OpenFermion ``givens_decomposition_square`` restated: -> (layers, diagonal)."""
    m = np.array(q, dtype=complex)
    n = m.shape[0]
    layers = []
    for k in range(2 * (n - 1) - 1):
        if k < n - 1:
            r0, c0 = 0, n - 1 - k
        else:
            r0, c0 = k - (n - 2), k - (n - 3)
        cols = list(range(c0, n, 2))
        ops = []
        for i, j in zip(range(r0, r0 + len(cols)), cols):
            r = np.conj(m[i, j])
            if abs(r) > TOL:
                l = np.conj(m[i, j - 1])
                if abs(l) < TOL:
                    c, s, ph = 1.0, 0.0, 1.0
                elif abs(r) < TOL:
                    c, s, ph = 0.0, 1.0, 1.0
                else:
                    h = np.sqrt(abs(l) ** 2 + abs(r) ** 2)
                    c, s = abs(r) / h, abs(l) / h
                    ph = (l / abs(l)) * np.conj(r / abs(r))
                if abs(np.imag(l)) < TOL and abs(np.imag(r)) < TOL:
                    g = np.array([[s, ph * c], [-ph * c, s]], dtype=complex)
                else:
                    g = np.array([[s, ph * c], [c, -ph * s]], dtype=complex)
                ops.append((j - 1, j, float(np.arcsin(np.real(g[1, 0]))), float(np.angle(g[1, 1]))))
                a, b = m[:, j - 1].copy(), m[:, j].copy()
                m[:, j - 1] = g[0, 0] * a + np.conj(g[0, 1]) * b
                m[:, j] = g[1, 0] * a + np.conj(g[1, 1]) * b
        if ops:
            layers.append(tuple(ops))
    return layers, np.diagonal(m).copy()


def k_space_occupation(nx, ny, t, n_up, n_dn):
    """Occupied k-orbitals: stable sort of eps_k = -t * sum_d f_d(k_d), f = cos for a length-2
    (de-duplicated) dimension else 2cos (reference adapt_vqe.py:104-122 on the symbolic FT of the
    quadratic term).  Ties broken by ascending orbital index."""
    ns = nx * ny

    def f(k, length):
        if length == 1:
            return 0.0
        c = np.cos(2 * np.pi * k / length)
        return c if length == 2 else 2 * c

    eps = [round(-t * (f(s % nx, nx) + f(s // nx, ny)), 6) for s in range(ns)]
    order = sorted(range(ns), key=lambda s: eps[s])
    up = [2 * s for s in order[:n_up]]
    dn = [2 * s + 1 for s in order[:n_dn]]
    return up, dn, eps
