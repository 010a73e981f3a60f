"""Host compile step: symbolic operators -> packed Pauli tables and pair/diagonal op plans.

"The Jordan-Wigner Hamiltonian and operator pool are compiled once on the host into packed
Pauli-term tables (x-mask, z-mask, phase, coefficient)".  This replaces
``QubitOperator_to_qmlHamiltonian`` (reference ``models/utils.py:30-56``) for observables and the
per-string gate expansion of ``Trotterize_generator`` / ``PauliStringRotation``
(``models/adapt_vqe.py:87-98``, ``models/utils.py:58-83``) for generators.

Bit convention: qubit/wire q <-> bit (n-1-q).  String (x, z) = i^k X^x Z^z, k = popcount(x&z).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .symbolic import FermionOperator, QubitOperator, jordan_wigner

_I_POW = (1, 1j, -1, -1j)
_TOL = 1e-12


def popcount(v: int) -> int:
    return bin(v).count("1")


def qubit_bit(q: int, n: int) -> int:
    return 1 << (n - 1 - q)


def pack_term(term, n: int):
    """((q, 'X'|'Y'|'Z'), ...) -> (x, z)."""
    x = z = 0
    for q, p in term:
        if q >= n:
            raise ValueError(f"qubit {q} outside a {n}-qubit register")
        b = 1 << (n - 1 - q)
        if p == "X":
            x |= b
        elif p == "Z":
            z |= b
        else:
            x |= b
            z |= b
    return x, z


class PauliTable:
    """Packed observable: arrays x, z (uint64), k (uint8 phase exponent), coeff (complex128)
    in ``QubitOperator.terms`` order (after ``compress()``, as the reference does)."""

    def __init__(self, n_qubits, x, z, coeff):
        self.n_qubits = int(n_qubits)
        x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1)
        z = np.ascontiguousarray(z, dtype=np.uint64).reshape(-1)
        coeff = np.ascontiguousarray(coeff, dtype=np.complex128).reshape(-1)
        if not (len(x) == len(z) == len(coeff)):
            raise ValueError("PauliTable: x, z and coeff must have the same length")
        # canonical form: one entry per string.  Duplicate (x, z) are merged (coefficients added, first-seen order),
        # exactly what accumulating into a QubitOperator's ``terms`` dict does, so every method (as_dict, dressed,
        # the device upload) sees the same operator.
        if len(x) > 1:
            keys = np.stack([x, z], axis=1)
            uniq, first, inv = np.unique(keys, axis=0, return_index=True, return_inverse=True)
            if len(uniq) != len(x):
                summed = np.zeros(len(uniq), dtype=np.complex128)
                np.add.at(summed, inv.reshape(-1), coeff)
                order = np.argsort(first, kind="stable")
                x, z, coeff = uniq[order, 0].copy(), uniq[order, 1].copy(), summed[order]
        self.x, self.z, self.coeff = x, z, coeff
        self.k = (np.bitwise_count(self.x & self.z) & 3).astype(np.uint8)

    @classmethod
    def from_operator(cls, op, n_qubits, compress=True):
        if isinstance(op, FermionOperator):
            op = jordan_wigner(op)
        if not isinstance(op, QubitOperator):
            raise TypeError("PauliTable needs a QubitOperator or FermionOperator")
        if compress:
            op = op.copy()
            op.compress()
        xs, zs, cs = [], [], []
        for term, c in op.terms.items():
            x, z = pack_term(term, n_qubits)
            xs.append(x)
            zs.append(z)
            cs.append(complex(c))
        return cls(n_qubits, xs, zs, cs)

    def __len__(self):
        return len(self.x)

    @property
    def n_groups(self):
        return len(set(int(v) for v in self.x))

    def as_dict(self):
        return {(int(a), int(b)): complex(c) for a, b, c in zip(self.x, self.z, self.coeff)}

    def conserves_species(self, tol: float = 0.0) -> bool:
        """Does the operator commute with N_up and N_dn (even wires = up orbitals)?  Per x-mask group: the weight of the
        partner with x-bit pattern ``pat`` is sum_m d_m (-1)^popcount(z_m & pat-bits) class by class (a class = the z bits
        outside x); every pattern with a non-zero weight must move as many electrons into as out of each species.
        Groups with more than four x bits are not analysed (-> False).  Same rule as the device planner (csrc/sector_eval.cu),
        which uses tol = 0 (exact cancellation, true for every table the drivers build)."""
        n = self.n_qubits
        up = sum(1 << b for b in range(n) if (n - 1 - b) % 2 == 0)
        i_pow = (1, 1j, -1, -1j)
        groups = {}
        for x, z, c, k in zip(self.x, self.z, self.coeff, self.k):
            groups.setdefault(int(x), []).append((int(z), complex(c) * i_pow[int(k)]))
        for x, terms in groups.items():
            if x == 0:
                continue
            pos = [b for b in range(n) if x >> b & 1]
            if len(pos) > 4:
                return False
            classes = {}
            for z, d in terms:
                classes.setdefault(z & ~x, []).append((z, d))
            for pat in range(1 << len(pos)):
                dep = sum(1 << pos[b] for b in range(len(pos)) if pat >> b & 1)
                live = any(abs(sum(d * (-1) ** bin(dep & z).count("1") for z, d in cl)) > tol for cl in classes.values())
                if not live:
                    continue
                # the partner j carries `pat` on the x bits, the output i = j ^ x the complement
                dup = sum((1 - 2 * (pat >> b & 1)) for b in range(len(pos)) if up >> pos[b] & 1)
                ddn = sum((1 - 2 * (pat >> b & 1)) for b in range(len(pos)) if not up >> pos[b] & 1)
                if dup != 0 or ddn != 0:
                    return False
        return True

    # -- iQCC dressing on packed masks (reference models/iqcc_hubbard.py:184-189) -----------------------------
    def dressed(self, xp: int, zp: int, tau: float, tol: float = 1e-12) -> "PauliTable":
        """exp(i tau P / 2) H exp(-i tau P / 2) for the Pauli string P = (xp, zp), i.e. the reference's
        ``H + sin(tau)(-i/2)[H, P] + (1/2)(1 - cos tau)(P H P - H)``, on the packed table.

        A term that commutes with P is unchanged; a term c_t P_t that anticommutes becomes
        ``cos(tau) c_t P_t - i sin(tau) c_t P_t P``, and P_t P = i^(k_t + k_P - k_3) (-1)^popcount(z_t & x_P) P_3
        with P_3 = (x_t ^ x_P, z_t ^ z_P).  Duplicate strings are merged and |c| < tol dropped.  Vectorised: the
        term table can grow to 10^5..10^6 entries."""
        xp64, zp64 = np.uint64(xp), np.uint64(zp)
        x, z, c = self.x, self.z, self.coeff
        par = (np.bitwise_count(x & zp64) + np.bitwise_count(z & xp64)) & 1      # 1: anticommutes with P
        anti = par.astype(bool)
        cs, sn = np.cos(tau), np.sin(tau)
        keep_c = np.where(anti, cs * c, c)
        xa, za, ca = x[anti], z[anti], c[anti]
        x3, z3 = xa ^ xp64, za ^ zp64
        k1 = np.bitwise_count(xa & za).astype(np.int64)
        k2 = int(bin(xp & zp).count("1"))
        k3 = np.bitwise_count(x3 & z3).astype(np.int64)
        phase = np.array(_I_POW)[(k1 + k2 - k3) & 3] * (1 - 2 * (np.bitwise_count(za & xp64) & 1).astype(np.int64))
        new_c = -1j * sn * ca * phase
        ax = np.concatenate([x, x3])
        az = np.concatenate([z, z3])
        ac = np.concatenate([keep_c, new_c])
        # merge equal strings, keeping first-seen order (as the reference's dict-based += does)
        keys = np.stack([ax, az], axis=1)
        uniq, first, inv = np.unique(keys, axis=0, return_index=True, return_inverse=True)
        summed = np.zeros(len(uniq), dtype=np.complex128)
        np.add.at(summed, inv.reshape(-1), ac)
        order = np.argsort(first, kind="stable")
        live = np.abs(summed[order]) > tol
        sel = order[live]
        return PauliTable(self.n_qubits, uniq[sel, 0], uniq[sel, 1], summed[sel])

    def to_operator(self):
        """QubitOperator with this table's terms (table order)."""
        op = QubitOperator()
        n = self.n_qubits
        for x, z, c in zip(self.x, self.z, self.coeff):
            x, z = int(x), int(z)
            term = []
            for q in range(n):
                b = 1 << (n - 1 - q)
                if x & b:
                    term.append((q, "Y" if z & b else "X"))
                elif z & b:
                    term.append((q, "Z"))
            op.terms[tuple(term)] = complex(c)
        return op


# ---------------------------------------------------------------------------------------------
# generator plans
# ---------------------------------------------------------------------------------------------
@dataclass
class PairPiece:
    """On every index i with (i & fixmask) == fixval:  (G psi)[i] = s B psi[i^x], (G psi)[i^x] = s conj(B) psi[i],
    s = (-1)^popcount(i & zeta).  fixmask contains the top bit of x and fixval has it clear."""
    x: int
    fixmask: int
    fixval: int
    zeta: int
    b: complex
    strings: list = field(default_factory=list)     # [(x, z)] Pauli strings it stands for (commutation checks)


@dataclass
class DiagPiece:
    """(G psi)[i] = sum_m coef[m] (-1)^popcount(i & z[m]) psi[i]."""
    z: list
    coef: list
    strings: list = field(default_factory=list)


def strings_commute(a, b) -> bool:
    (x1, z1), (x2, z2) = a, b
    return (popcount(x1 & z2) + popcount(x2 & z1)) % 2 == 0


def all_commute(strings) -> bool:
    for i in range(len(strings)):
        for j in range(i + 1, len(strings)):
            if not strings_commute(strings[i], strings[j]):
                return False
    return True


def _single_string_piece(x, z, c):
    k = popcount(x & z)
    d = c * _I_POW[k & 3]
    b = d * (-1) ** k               # w(i) = B (-1)^popcount(i&z)
    top = 1 << (x.bit_length() - 1)
    return PairPiece(x, top, 0, z, b, [(x, z)])


def _analyse_group(x, terms, max_patterns=8):
    """terms: [(z, real coeff)] sharing x != 0 -> list of PairPiece, or None when the group is not
    a union of simple 2x2 blocks (caller falls back to one piece per string)."""
    if len(terms) == 1:
        return [_single_string_piece(x, terms[0][0], terms[0][1])]
    xbits = [b for b in range(x.bit_length()) if x >> b & 1]
    if len(xbits) > 10:
        return None
    top = 1 << xbits[-1]
    classes = {}
    for z, c in terms:
        classes.setdefault(z & ~x, []).append((z & x, c * _I_POW[popcount(x & z) & 3]))
    pieces = []
    strings = [(x, z) for z, _ in terms]
    for pat in range(1 << (len(xbits) - 1)):          # top bit of the pattern stays 0
        a = 0
        for t, b in enumerate(xbits[:-1]):
            if pat >> t & 1:
                a |= 1 << b
        partner_bits = a ^ x
        active = []
        for zeta, members in classes.items():
            val = sum(d * (-1) ** popcount(partner_bits & zin) for zin, d in members)
            if abs(val) > _TOL:
                active.append((zeta, val))
        if not active:
            continue
        if len(active) > 1:
            return None
        pieces.append(PairPiece(x, x, a, active[0][0], complex(active[0][1]), strings))
        if len(pieces) > max_patterns:
            return None
    assert all(p.fixmask & top for p in pieces)
    return pieces


class GeneratorPlan:
    """exp(-i theta G) for G = sum_m Re(c_m) P_m as a list of pair / diagonal pieces.

    ``exact`` is True when every string of G commutes with every other: then the product of the
    pieces (any order) equals the reference's literal Trotter product exactly.  Otherwise the plan
    holds one piece per string in ``generator.terms`` order, which *is* the reference's product.
    """

    def __init__(self, generator, n_qubits):
        if isinstance(generator, FermionOperator):
            generator = jordan_wigner(generator)
        self.n_qubits = n_qubits
        strings = []
        for term, c in generator.terms.items():
            if not term:
                continue                                  # identity: global phase (reference adapt_vqe.py:92-93)
            x, z = pack_term(term, n_qubits)
            strings.append((x, z, complex(c).real))
        self.strings = strings
        self.exact = all_commute([(x, z) for x, z, _ in strings])
        self.pieces = []
        if self.exact:
            groups = {}
            for x, z, c in strings:
                groups.setdefault(x, []).append((z, c))
            for x, terms in groups.items():
                if x == 0:
                    self.pieces.append(DiagPiece([z for z, _ in terms], [c for _, c in terms], [(0, z) for z, _ in terms]))
                    continue
                got = _analyse_group(x, terms)
                if got is None:
                    got = [_single_string_piece(x, z, c) for z, c in terms]
                self.pieces.extend(got)
        else:
            for x, z, c in strings:
                if x == 0:
                    self.pieces.append(DiagPiece([z], [c], [(0, z)]))
                else:
                    self.pieces.append(_single_string_piece(x, z, c))
        self.pieces = [p for p in self.pieces if isinstance(p, DiagPiece) or abs(p.b) > _TOL]
