"""Jordan-Wigner transform, the Fermi-Hubbard lattice model and the Givens network.

Restates the OpenFermion routines the reference calls by name (third-party, not
vendored, unpinned): ``fermi_hubbard`` (reference ``models/adapt_vqe.py:159``),
``jordan_wigner`` (``adapt_vqe.py:143,166``; ``iqcc_hubbard.py:43``),
``get_interaction_operator`` (``iqcc_hubbard.py:41``) and
``givens_decomposition_square`` (``adapt_vqe.py:187``).  Term insertion order follows
the published algorithms because the reference iterates ``.terms`` in dict order
(``adapt_vqe.py:91``, ``iqcc_hubbard.py:86``).
"""
from __future__ import annotations

import itertools

import numpy as np

from .ops import (EQ_TOLERANCE, FermionOperator, QubitOperator, count_qubits,
                  down_index, normal_ordered, number_operator, up_index)


# ---------------------------------------------------------------------------
# Jordan-Wigner
# ---------------------------------------------------------------------------
def _jw_ladder(index: int, action: int) -> QubitOperator:
    """a†_j = Z_0..Z_{j-1} (X_j - iY_j)/2 ;  a_j = Z_0..Z_{j-1} (X_j + iY_j)/2."""
    z_string = tuple((q, "Z") for q in range(index))
    op = QubitOperator._from_terms({z_string + ((index, "X"),): 0.5})
    op.terms[z_string + ((index, "Y"),)] = -0.5j if action else 0.5j
    return op


def jordan_wigner(operator):
    """FermionOperator | InteractionOperator -> QubitOperator (qubit j <-> mode j)."""
    if isinstance(operator, QubitOperator):
        return operator
    if isinstance(operator, InteractionOperator):
        return _jordan_wigner_interaction_op(operator)
    if not isinstance(operator, FermionOperator):
        raise TypeError("jordan_wigner expects a FermionOperator or InteractionOperator")
    ladder_cache = {}
    transformed = QubitOperator()
    for term, coeff in operator.terms.items():
        piece = QubitOperator((), coeff)
        for factor in term:
            lad = ladder_cache.get(factor)
            if lad is None:
                lad = ladder_cache[factor] = _jw_ladder(*factor)
            piece *= lad
        transformed += piece
    return transformed


# ---------------------------------------------------------------------------
# InteractionOperator (constant + one-body + two-body tensors)
# ---------------------------------------------------------------------------
class InteractionOperator:
    """``constant + sum h[p,q] a†_p a_q + sum g[p,q,r,s] a†_p a†_q a_r a_s``."""

    def __init__(self, constant, one_body_tensor, two_body_tensor):
        self.constant = constant
        self.one_body_tensor = np.asarray(one_body_tensor)
        self.two_body_tensor = np.asarray(two_body_tensor)
        self.n_qubits = self.one_body_tensor.shape[0]

    def __getitem__(self, key):
        if len(key) == 2:
            (p, _), (q, _) = key
            return self.one_body_tensor[p, q]
        (p, _), (q, _), (r, _), (s, _) = key
        return self.two_body_tensor[p, q, r, s]


def get_interaction_operator(fermion_operator: FermionOperator, n_qubits=None) -> InteractionOperator:
    """Normal-order, then scatter the <=2-body, particle-conserving terms into tensors."""
    if not isinstance(fermion_operator, FermionOperator):
        raise TypeError("Input operator must be a FermionOperator.")
    n = count_qubits(fermion_operator) if n_qubits is None else n_qubits
    if n < count_qubits(fermion_operator):
        raise ValueError("Invalid number of qubits specified.")
    ordered = normal_ordered(fermion_operator)
    constant = 0.0
    one_body = np.zeros((n, n), complex)
    two_body = np.zeros((n, n, n, n), complex)
    for term, coeff in ordered.terms.items():
        if len(term) == 0:
            constant = coeff
        elif len(term) == 2 and [a for _, a in term] == [1, 0]:
            p, q = (i for i, _ in term)
            one_body[p, q] = coeff
        elif len(term) == 4 and [a for _, a in term] == [1, 1, 0, 0]:
            p, q, r, s = (i for i, _ in term)
            two_body[p, q, r, s] = coeff
        else:
            raise ValueError("FermionOperator does not map to InteractionOperator "
                             "(not particle-conserving or more than two-body).")
    if not np.any(np.iscomplex(one_body)) and not np.any(np.iscomplex(two_body)):
        one_body, two_body = one_body.real.copy(), two_body.real.copy()
        constant = constant.real if isinstance(constant, complex) else constant
    return InteractionOperator(constant, one_body, two_body)


def _jw_one_body(p, q, coefficient) -> QubitOperator:
    op = QubitOperator()
    coefficient = complex(coefficient)
    if p != q:
        if p > q:
            p, q = q, p
            coefficient = coefficient.conjugate()
        parity = tuple((z, "Z") for z in range(p + 1, q))
        for c, (a, b) in ((coefficient.real, "XX"), (coefficient.real, "YY"),
                          (coefficient.imag, "XY"), (-coefficient.imag, "YX")):
            op += QubitOperator(((p, a),) + parity + ((q, b),), 0.5 * c)
    else:
        c = coefficient.real if coefficient.imag == 0 else coefficient
        op += QubitOperator((), 0.5 * c)
        op += QubitOperator(((p, "Z"),), -0.5 * c)
    return op


def _jw_two_body_density(p, q, coefficient) -> QubitOperator:
    """a†_p a†_q a_p a_q (p != q) = -n_p n_q."""
    c = complex(coefficient)
    c = c.real if c.imag == 0 else c
    coeff = 0.25 * c
    op = QubitOperator()
    op -= QubitOperator((), coeff)
    op += QubitOperator(((p, "Z"),), coeff)
    op += QubitOperator(((q, "Z"),), coeff)
    op -= QubitOperator(((min(p, q), "Z"), (max(p, q), "Z")), coeff)
    return op


def _jordan_wigner_interaction_op(iop: InteractionOperator) -> QubitOperator:
    """Structured JW of an InteractionOperator: constant, diagonal one-body, then per
    pair (p<q) the hopping and the density-density part, then every remaining two-body
    term through the ladder-operator transform."""
    n = iop.n_qubits
    h, g = iop.one_body_tensor, iop.two_body_tensor
    op = QubitOperator((), iop.constant)
    for p in range(n):
        op += _jw_one_body(p, p, h[p, p])
    for p, q in itertools.combinations(range(n), 2):
        op += _jw_one_body(p, q, 0.5 * (h[p, q] + np.conj(h[q, p])))
        coeff = g[p, q, p, q] - g[p, q, q, p] - g[q, p, p, q] + g[q, p, q, p]
        op += _jw_two_body_density(p, q, coeff)
    rest = FermionOperator()
    for p, q, r, s in zip(*np.nonzero(g)):
        if len({p, q, r, s}) == 2 and {p, q} == {r, s}:
            continue
        rest += FermionOperator(((int(p), 1), (int(q), 1), (int(r), 0), (int(s), 0)),
                                complex(g[p, q, r, s]))
    if rest.terms:
        op += jordan_wigner(rest)
    return op


# ---------------------------------------------------------------------------
# Fermi-Hubbard model on a rectangular lattice
# ---------------------------------------------------------------------------
def _right_neighbor(site, nx, ny, periodic):
    if nx == 1:
        return None
    if (site + 1) % nx == 0:
        return site + 1 - nx if periodic else None
    return site + 1


def _bottom_neighbor(site, nx, ny, periodic):
    if ny == 1:
        return None
    if site + nx + 1 > nx * ny:
        return site + nx - nx * ny if periodic else None
    return site + nx


def _hopping(i, j, coeff) -> FermionOperator:
    op = FermionOperator(((i, 1), (j, 0)), coeff)
    op += FermionOperator(((j, 1), (i, 0)), np.conj(coeff) if isinstance(coeff, complex) else coeff)
    return op


def _coulomb(n_modes, i, j, coeff, particle_hole_symmetry) -> FermionOperator:
    op = FermionOperator(((i, 1), (i, 0), (j, 1), (j, 0)), coeff)
    if particle_hole_symmetry:
        op -= number_operator(n_modes, i, 0.5 * coeff)
        op -= number_operator(n_modes, j, 0.5 * coeff)
        op += FermionOperator((), 0.25 * coeff)
    return op


def fermi_hubbard(x_dimension, y_dimension, tunneling, coulomb, chemical_potential=0.0,
                  magnetic_field=0.0, periodic=True, spinless=False,
                  particle_hole_symmetry=False) -> FermionOperator:
    """H = -t sum_<ij>,s (a†_is a_js + h.c.) + U sum_i n_i,up n_i,dn  (- mu, - h terms).

    Site s = x + y*Nx, spin orbital 2s (up) / 2s+1 (down).  Bonds are the right and
    bottom neighbour of every site; a periodic dimension of length 2 contributes each
    bond once.
    """
    nx, ny = x_dimension, y_dimension
    n_sites = nx * ny
    model = FermionOperator()
    if spinless:
        for site in range(n_sites):
            right = _right_neighbor(site, nx, ny, periodic)
            bottom = _bottom_neighbor(site, nx, ny, periodic)
            if nx == 2 and periodic and site % 2 == 1:
                right = None
            if ny == 2 and periodic and site >= nx:
                bottom = None
            if right is not None:
                model += _hopping(site, right, -tunneling)
                model += _coulomb(n_sites, site, right, coulomb, particle_hole_symmetry)
            if bottom is not None:
                model += _hopping(site, bottom, -tunneling)
                model += _coulomb(n_sites, site, bottom, coulomb, particle_hole_symmetry)
            model += number_operator(n_sites, site, -chemical_potential)
        return model

    n_modes = 2 * n_sites
    for site in range(n_sites):
        right = _right_neighbor(site, nx, ny, periodic)
        bottom = _bottom_neighbor(site, nx, ny, periodic)
        if nx == 2 and periodic and site % 2 == 1:
            right = None
        if ny == 2 and periodic and site >= nx:
            bottom = None
        if right is not None:
            model += _hopping(up_index(site), up_index(right), -tunneling)
            model += _hopping(down_index(site), down_index(right), -tunneling)
        if bottom is not None:
            model += _hopping(up_index(site), up_index(bottom), -tunneling)
            model += _hopping(down_index(site), down_index(bottom), -tunneling)
        model += _coulomb(n_modes, up_index(site), down_index(site), coulomb, particle_hole_symmetry)
        model += number_operator(n_modes, up_index(site), -chemical_potential - magnetic_field)
        model += number_operator(n_modes, down_index(site), -chemical_potential + magnetic_field)
    return model


# ---------------------------------------------------------------------------
# Givens decomposition of a square unitary (Gaussian basis change network)
# ---------------------------------------------------------------------------
def _givens_right(a, b):
    """2x2 unitary G whose column action zeroes ``b``: see givens_decomposition_square."""
    if abs(a) < EQ_TOLERANCE:
        c, s, phase = 1.0, 0.0, 1.0
    elif abs(b) < EQ_TOLERANCE:
        c, s, phase = 0.0, 1.0, 1.0
    else:
        hyp = np.sqrt(abs(a) ** 2 + abs(b) ** 2)
        c, s = abs(b) / hyp, abs(a) / hyp
        phase = (a / abs(a)) * np.conj(b / abs(b))
        if np.isreal(phase):
            phase = np.real(phase)
    if abs(np.imag(a)) < EQ_TOLERANCE and abs(np.imag(b)) < EQ_TOLERANCE:
        return np.array([[s, phase * c], [-phase * c, s]])
    return np.array([[s, phase * c], [c, -phase * s]])


def givens_decomposition_square(unitary_matrix, always_insert=False):
    """Decompose ``Q`` into layers of adjacent-column Givens rotations and a diagonal.

    Returns ``(decomposition, diagonal)`` where ``decomposition`` is a list of tuples of
    ``(j-1, j, theta, phi)`` that can run in parallel.  Elements are zeroed along
    anti-diagonals starting from the top-right corner.  The circuit realising it is:
    ``RZ(angle(diagonal[q]))`` on every wire, then the layers reversed, each entry as
    ``SingleExcitation(2 theta, [i, j])`` followed by ``RZ(phi, j)``
    (reference ``models/adapt_vqe.py:344-354``).
    """
    m = np.array(unitary_matrix, dtype=complex)
    n = m.shape[0]
    decomposition = []
    for k in range(2 * (n - 1) - 1):
        if k < n - 1:
            start_row, start_col = 0, n - 1 - k
        else:
            start_row, start_col = k - (n - 2), k - (n - 3)
        cols = range(start_col, n, 2)
        rows = range(start_row, start_row + len(cols))
        layer = []
        for i, j in zip(rows, cols):
            right = np.conj(m[i, j])
            if always_insert or abs(right) > EQ_TOLERANCE:
                left = np.conj(m[i, j - 1])
                g = _givens_right(left, right)
                theta = float(np.arcsin(np.real(g[1, 0])))
                phi = float(np.angle(g[1, 1]))
                layer.append((j - 1, j, theta, phi))
                col_a = m[:, j - 1].copy()
                col_b = m[:, j].copy()
                m[:, j - 1] = g[0, 0] * col_a + np.conj(g[0, 1]) * col_b
                m[:, j] = g[1, 0] * col_a + np.conj(g[1, 1]) * col_b
        if layer:
            decomposition.append(tuple(layer))
    return decomposition, m.diagonal().copy()
