"""Backend boundary helpers (reference ``models/utils.py``).

* ``QubitOperator_to_qmlHamiltonian`` (reference :30-56): symbolic operator -> packed observable.
* ``PauliStringRotation`` (reference :58-83): exp(-i theta P / 2).  The reference spells it as
  RY/RX basis changes + CNOT ladder + RZ; here it is one in-place pair-rotation op (same unitary).
* ``compile_hva_hopping_indices`` / ``get_hva_commuting_hopping_terms`` (reference :145-333): bond
  colouring of the lattice into sets of mutually commuting hopping terms.

The qiskit helpers of the reference file (:11-28, :85-143) have no caller and are not provided.
"""
from __future__ import annotations

from fhsim.recording import Param, active_circuit
from fhsim.symbolic import FermionOperator, QubitOperator, count_qubits, jordan_wigner
from fhsim.tables import PauliTable, pack_term


class Observable:
    """What ``qml.Hamiltonian(coeffs, obs)`` is to the reference: coefficients + Pauli words, plus
    the packed table the device consumes."""

    def __init__(self, op: QubitOperator, n_qubits=None):
        self.operator = op
        self.coeffs = list(op.terms.values())
        self.ops = list(op.terms.keys())
        self._tables = {}
        self.n_qubits = n_qubits

    def table(self, n_qubits=None) -> PauliTable:
        n = n_qubits or self.n_qubits or max(count_qubits(self.operator), 1)
        if n not in self._tables:
            self._tables[n] = PauliTable.from_operator(self.operator, n, compress=False)
        return self._tables[n]

    def __len__(self):
        return len(self.coeffs)


def QubitOperator_to_qmlHamiltonian(op, mapper=jordan_wigner):
    if isinstance(op, FermionOperator):
        op = mapper(op)
    op.compress()
    return Observable(op)


def PauliStringRotation(theta, pauliString):
    """exp(-i theta/2 P) for ``pauliString = (letters, wires)`` appended to the circuit being recorded."""
    letters, wires = pauliString
    circuit = active_circuit()
    x, z = pack_term(tuple(zip(wires, letters)), circuit.n)
    if isinstance(theta, Param):
        circuit.pauli_rotation(x, z, 0.5 * theta.mult, param=theta.index)
    else:
        circuit.pauli_rotation(x, z, 0.5, angle=float(theta))


# ---------------------------------------------------------------------------------------------
# HVA bond colouring
# ---------------------------------------------------------------------------------------------
def _bond_sets(length, periodic):
    """Sets of nearest-neighbour coordinate pairs along one dimension; bonds inside a set share no site.

    length 2 -> 1 set; odd periodic length > 2 -> 3 sets (even bonds, odd bonds, wrap-around bond);
    otherwise 2 sets (the wrap-around bond of an even periodic ring joins the odd set).
    """
    if length < 2:
        return []
    if length == 2:
        return [[(0, 1)]]
    even = [(a, a + 1) for a in range(0, length - 1, 2)]
    odd = [(a, a + 1) for a in range(1, length - 1, 2)]
    if periodic and length % 2 == 1:
        return [even, odd, [(0, length - 1)]]
    if periodic:
        return [even, odd + [(0, length - 1)]]
    return [even, odd]


def compile_hva_hopping_indices(x_dimension, y_dimension, periodic):
    def index(x, y, spin):
        return 2 * (x + y * x_dimension) + spin

    horizontal = []
    for bonds in _bond_sets(x_dimension, periodic):
        horizontal.append([(index(a, y, s), index(b, y, s))
                           for y in range(y_dimension) for (a, b) in bonds for s in (0, 1)])
    vertical = []
    for bonds in _bond_sets(y_dimension, periodic):
        vertical.append([(index(x, a, s), index(x, b, s))
                         for x in range(x_dimension) for (a, b) in bonds for s in (0, 1)])
    return horizontal, vertical


def _hopping_generator(pairs):
    generator = FermionOperator()
    for i, j in pairs:
        generator += FermionOperator(f'{i}^ {j}')
        generator += FermionOperator(f'{j}^ {i}')
    return generator


def get_hva_commuting_hopping_terms(x_dimesion, y_dimension, periodic):
    horizontal, vertical = compile_hva_hopping_indices(x_dimesion, y_dimension, periodic)
    return [_hopping_generator(s) for s in horizontal], [_hopping_generator(s) for s in vertical]
