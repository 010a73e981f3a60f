"""Symbolic operator algebra used by the host-side compile step.

The reference takes and returns OpenFermion ``FermionOperator`` / ``QubitOperator``
objects everywhere in ``operators/`` and in every driver ``__init__``
(reference ``models/adapt_vqe.py:7-17``, ``operators/pool.py:3-9``).  OpenFermion
is a third-party dependency that is absent from ``/root/reference`` (unpinned,
no lock file), so this module restates the subset of its published behaviour the
hot path's callers rely on:

* ``.terms`` is an insertion-ordered dict ``{term_tuple: coefficient}``;
* ``+=`` merges coefficients and deletes entries that become smaller than 1e-8;
* ``*`` distributes left-term-major and keeps exact-zero products until ``compress``;
* ``FermionOperator`` keys are ``((index, 1|0), ...)`` in the written order,
  ``QubitOperator`` keys are ``((qubit, 'X'|'Y'|'Z'), ...)`` sorted by qubit, with
  same-qubit Paulis multiplied out.

Only host logic lives here -- no numerics of the statevector path.
"""
from __future__ import annotations

import numbers
import re

EQ_TOLERANCE = 1e-8

_COEFF_TYPES = (int, float, complex, numbers.Number)


def _is_small(value) -> bool:
    return abs(value) < EQ_TOLERANCE


class SymbolicOperator:
    """Shared machinery of :class:`FermionOperator` and :class:`QubitOperator`."""

    __slots__ = ("terms",)
    __hash__ = None

    # -- subclass hooks -----------------------------------------------------
    @staticmethod
    def _parse_string_factor(token: str):
        raise NotImplementedError

    @staticmethod
    def _check_factor(factor):
        raise NotImplementedError

    def _simplify(self, term, coefficient=1.0):
        return coefficient, tuple(term)

    @staticmethod
    def _format_factor(factor) -> str:
        raise NotImplementedError

    # -- pickling -----------------------------------------------------------
    # OpenFermion operators pickle as cls.__new__(cls) + the instance dict {'terms': {...}}; this class has __slots__,
    # whose default state is (None, {'terms': ...}).  Accept both so checkpoints written by the reference load
    # (fhsim/checkpoint.py maps the class paths).
    def __getstate__(self):
        return {"terms": self.terms}

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2:
            merged = {}
            for part in state:
                if part:
                    merged.update(part)
            state = merged
        self.terms = dict(state["terms"])

    # -- construction -------------------------------------------------------
    def __init__(self, term=None, coefficient=1.0):
        if not isinstance(coefficient, _COEFF_TYPES):
            raise ValueError("Coefficient must be a numeric type.")
        self.terms = {}
        if term is None:
            return
        if isinstance(term, str):
            parsed = self._parse_string(term)
        elif isinstance(term, (tuple, list)):
            parsed = self._parse_sequence(term)
        else:
            raise ValueError(f"term specified incorrectly: {term!r}")
        coefficient, parsed = self._simplify(parsed, coefficient)
        self.terms[parsed] = coefficient

    @classmethod
    def _parse_sequence(cls, term):
        if not term:
            return ()
        # a single factor written as (index, action) is promoted to a 1-term
        if not isinstance(term[0], (tuple, list)):
            term = (tuple(term),)
        out = []
        for factor in term:
            if len(factor) != 2:
                raise ValueError(f"Invalid factor {factor!r}")
            index, action = factor
            index = int(index)
            factor = (index, action if isinstance(action, str) else int(action))
            cls._check_factor(factor)
            out.append(factor)
        return tuple(out)

    @classmethod
    def _parse_string(cls, text: str):
        tokens = text.split()
        return tuple(cls._parse_string_factor(tok) for tok in tokens)

    @classmethod
    def zero(cls):
        return cls()

    @classmethod
    def identity(cls):
        return cls(())

    @classmethod
    def _from_terms(cls, terms: dict):
        new = cls()
        new.terms = terms
        return new

    def copy(self):
        return self._from_terms(dict(self.terms))

    __copy__ = copy

    def __deepcopy__(self, memo):
        return self.copy()

    # pickle support with __slots__
    def __getstate__(self):
        return {"terms": self.terms}

    def __setstate__(self, state):
        self.terms = state["terms"]

    # -- arithmetic ---------------------------------------------------------
    def __iadd__(self, addend):
        if isinstance(addend, type(self)):
            terms = self.terms
            for term, coeff in addend.terms.items():
                value = terms.get(term, 0.0) + coeff
                terms[term] = value
                if _is_small(value):
                    del terms[term]
            return self
        if isinstance(addend, _COEFF_TYPES):
            return self.__iadd__(type(self)((), addend))
        return NotImplemented

    def __add__(self, addend):
        result = self.copy()
        r = result.__iadd__(addend)
        return r

    def __radd__(self, addend):
        return self + addend

    def __isub__(self, subtrahend):
        if isinstance(subtrahend, type(self)):
            terms = self.terms
            for term, coeff in subtrahend.terms.items():
                value = terms.get(term, 0.0) - coeff
                terms[term] = value
                if _is_small(value):
                    del terms[term]
            return self
        if isinstance(subtrahend, _COEFF_TYPES):
            return self.__isub__(type(self)((), subtrahend))
        return NotImplemented

    def __sub__(self, subtrahend):
        result = self.copy()
        return result.__isub__(subtrahend)

    def __rsub__(self, minuend):
        return -1 * self + minuend

    def __neg__(self):
        return -1 * self

    def __imul__(self, multiplier):
        if isinstance(multiplier, _COEFF_TYPES):
            for term in self.terms:
                self.terms[term] *= multiplier
            return self
        if isinstance(multiplier, type(self)):
            result = {}
            simplify = self._simplify
            for left, lc in self.terms.items():
                for right, rc in multiplier.terms.items():
                    coeff, term = simplify(left + right, lc * rc)
                    if term in result:
                        result[term] += coeff
                    else:
                        result[term] = coeff
            self.terms = result
            return self
        return NotImplemented

    def __mul__(self, multiplier):
        if isinstance(multiplier, _COEFF_TYPES) or isinstance(multiplier, type(self)):
            product = self.copy()
            return product.__imul__(multiplier)
        return NotImplemented

    def __rmul__(self, multiplier):
        if isinstance(multiplier, _COEFF_TYPES):
            return self * multiplier
        return NotImplemented

    def __truediv__(self, divisor):
        if not isinstance(divisor, _COEFF_TYPES):
            raise TypeError("Cannot divide operator by non-scalar type.")
        return self * (1.0 / divisor)

    def __itruediv__(self, divisor):
        if not isinstance(divisor, _COEFF_TYPES):
            raise TypeError("Cannot divide operator by non-scalar type.")
        return self.__imul__(1.0 / divisor)

    def __pow__(self, exponent):
        if not isinstance(exponent, int) or exponent < 0:
            raise ValueError("exponent must be a non-negative int")
        result = type(self)(())
        for _ in range(exponent):
            result *= self
        return result

    # -- comparison ---------------------------------------------------------
    def isclose(self, other, tol=EQ_TOLERANCE):
        if not isinstance(other, type(self)):
            return NotImplemented
        mine, theirs = self.terms, other.terms
        for term, a in mine.items():
            if term in theirs:
                b = theirs[term]
                if abs(a - b) > tol * max(1.0, abs(a), abs(b)):
                    return False
            elif abs(a) > tol:
                return False
        for term, b in theirs.items():
            if term not in mine and abs(b) > tol:
                return False
        return True

    def __eq__(self, other):
        if not isinstance(other, type(self)):
            return NotImplemented
        return self.isclose(other)

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    # -- utilities ----------------------------------------------------------
    def compress(self, abs_tol=EQ_TOLERANCE):
        """Drop |c| < abs_tol, demote complex->real (and real->imag) parts below it."""
        new_terms = {}
        for term, coeff in self.terms.items():
            if isinstance(coeff, complex):
                if abs(coeff.imag) <= abs_tol:
                    coeff = coeff.real
                elif abs(coeff.real) <= abs_tol:
                    coeff = 1j * coeff.imag
            if abs(coeff) > abs_tol:
                new_terms[term] = coeff
        self.terms = new_terms

    def get_operators(self):
        for term, coeff in self.terms.items():
            yield type(self)._from_terms({term: coeff})

    def many_body_order(self):
        if not self.terms:
            return 0
        return max(len(term) for term, c in self.terms.items() if abs(c) > EQ_TOLERANCE)

    def induced_norm(self, order=1):
        return sum(abs(c) ** order for c in self.terms.values()) ** (1.0 / order)

    def __len__(self):
        return len(self.terms)

    def __iter__(self):
        return self.get_operators()

    def __str__(self):
        if not self.terms:
            return "0"
        pieces = []
        for term, coeff in sorted(self.terms.items(), key=lambda kv: _sort_key(kv[0])):
            body = " ".join(self._format_factor(f) for f in term)
            pieces.append(f"{coeff} [{body}]")
        return " +\n".join(pieces)

    def __repr__(self):
        return str(self)


def _sort_key(term):
    return tuple((i, str(a)) for i, a in term)


# ---------------------------------------------------------------------------
class FermionOperator(SymbolicOperator):
    """Sum of products of ladder operators; factor ``(p, 1)`` = a†_p, ``(p, 0)`` = a_p."""

    __slots__ = ()
    _token = re.compile(r"^(\d+)(\^?)$")

    @staticmethod
    def _parse_string_factor(token):
        m = FermionOperator._token.match(token)
        if not m:
            raise ValueError(f"Invalid ladder operator '{token}'")
        return (int(m.group(1)), 1 if m.group(2) else 0)

    @staticmethod
    def _check_factor(factor):
        index, action = factor
        if index < 0 or action not in (0, 1):
            raise ValueError(f"Invalid ladder operator {factor!r}")

    @staticmethod
    def _format_factor(factor):
        return f"{factor[0]}^" if factor[1] else f"{factor[0]}"

    def is_normal_ordered(self):
        for term in self.terms:
            for a, b in zip(term[:-1], term[1:]):
                if a[1] < b[1] or (a[1] == b[1] and a[0] <= b[0]):
                    return False
        return True


# ---------------------------------------------------------------------------
# Pauli products on one qubit: (left, right) -> (phase, result)
_PAULI_PRODUCT = {
    ("X", "X"): (1.0, "I"), ("Y", "Y"): (1.0, "I"), ("Z", "Z"): (1.0, "I"),
    ("X", "Y"): (1j, "Z"), ("Y", "X"): (-1j, "Z"),
    ("Y", "Z"): (1j, "X"), ("Z", "Y"): (-1j, "X"),
    ("Z", "X"): (1j, "Y"), ("X", "Z"): (-1j, "Y"),
    ("I", "X"): (1.0, "X"), ("X", "I"): (1.0, "X"),
    ("I", "Y"): (1.0, "Y"), ("Y", "I"): (1.0, "Y"),
    ("I", "Z"): (1.0, "Z"), ("Z", "I"): (1.0, "Z"),
    ("I", "I"): (1.0, "I"),
}


class QubitOperator(SymbolicOperator):
    """Sum of Pauli strings; factor ``(q, 'X'|'Y'|'Z')``; factors on distinct qubits commute."""

    __slots__ = ()
    _token = re.compile(r"^([XYZ])(\d+)$")

    @staticmethod
    def _parse_string_factor(token):
        m = QubitOperator._token.match(token)
        if not m:
            raise ValueError(f"Invalid Pauli factor '{token}'")
        return (int(m.group(2)), m.group(1))

    @staticmethod
    def _check_factor(factor):
        index, action = factor
        if index < 0 or action not in ("X", "Y", "Z"):
            raise ValueError(f"Invalid Pauli factor {factor!r}")

    @staticmethod
    def _format_factor(factor):
        return f"{factor[1]}{factor[0]}"

    def _simplify(self, term, coefficient=1.0):
        if not term:
            return coefficient, ()
        term = sorted(term, key=lambda f: f[0])   # stable: same-qubit order preserved
        out = []
        left_index, left_action = term[0]
        for right_index, right_action in term[1:]:
            if left_index == right_index:
                phase, left_action = _PAULI_PRODUCT[left_action, right_action]
                coefficient = coefficient * phase
            else:
                if left_action != "I":
                    out.append((left_index, left_action))
                left_index, left_action = right_index, right_action
        if left_action != "I":
            out.append((left_index, left_action))
        return coefficient, tuple(out)


# ---------------------------------------------------------------------------
def hermitian_conjugated(operator):
    """Reverse ladder order and flip daggers (Fermion) / conjugate coefficients (both)."""
    if isinstance(operator, FermionOperator):
        out = {}
        for term, coeff in operator.terms.items():
            conj_term = tuple((index, 1 - action) for index, action in reversed(term))
            c = coeff.conjugate() if isinstance(coeff, complex) else coeff
            out[conj_term] = c
        return FermionOperator._from_terms(out)
    if isinstance(operator, QubitOperator):
        out = {}
        for term, coeff in operator.terms.items():
            out[term] = coeff.conjugate() if isinstance(coeff, complex) else coeff
        return QubitOperator._from_terms(out)
    raise TypeError("hermitian_conjugated expects a FermionOperator or QubitOperator")


def count_qubits(operator) -> int:
    """1 + highest mode / qubit index appearing in the operator."""
    if hasattr(operator, "n_qubits") and not isinstance(operator, SymbolicOperator):
        return int(operator.n_qubits)
    highest = -1
    for term in operator.terms:
        for index, _ in term:
            if index > highest:
                highest = index
    return highest + 1


def up_index(site: int) -> int:
    return 2 * site


def down_index(site: int) -> int:
    return 2 * site + 1


def number_operator(n_modes, mode=None, coefficient=1.0):
    if mode is None:
        op = FermionOperator()
        for m in range(n_modes):
            op += number_operator(n_modes, m, coefficient)
        return op
    return FermionOperator(((mode, 1), (mode, 0)), coefficient)


def normal_ordered(operator: FermionOperator) -> FermionOperator:
    """Creation operators to the left, descending indices within each block."""
    ordered = FermionOperator()
    for term, coeff in operator.terms.items():
        ordered += _normal_ordered_term(term, coeff)
    return ordered


def _normal_ordered_term(term, coefficient) -> FermionOperator:
    term = list(term)
    ordered = FermionOperator()
    for i in range(1, len(term)):
        for j in range(i, 0, -1):
            right = term[j]
            left = term[j - 1]
            if right[1] and not left[1]:
                # a_p a†_q -> -a†_q a_p (+ delta_pq)
                term[j - 1], term[j] = right, left
                coefficient = -coefficient
                if right[0] == left[0]:
                    contracted = term[: j - 1] + term[j + 1:]
                    ordered += _normal_ordered_term(tuple(contracted), -coefficient)
            elif right[1] == left[1]:
                if right[0] == left[0]:
                    return ordered            # a_p a_p = 0
                if right[0] > left[0]:
                    term[j - 1], term[j] = right, left
                    coefficient = -coefficient
    ordered += FermionOperator(tuple(term), coefficient)
    return ordered
