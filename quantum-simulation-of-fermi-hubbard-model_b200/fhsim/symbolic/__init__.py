"""OpenFermion-compatible names for the host-side symbolic layer (see ops.py, transforms.py)."""
from .ops import (EQ_TOLERANCE, FermionOperator, QubitOperator, SymbolicOperator,
                  count_qubits, down_index, hermitian_conjugated, normal_ordered,
                  number_operator, up_index)
from .transforms import (InteractionOperator, fermi_hubbard, get_interaction_operator,
                         givens_decomposition_square, jordan_wigner)

__all__ = [
    "EQ_TOLERANCE", "FermionOperator", "QubitOperator", "SymbolicOperator", "InteractionOperator",
    "count_qubits", "up_index", "down_index", "hermitian_conjugated", "normal_ordered",
    "number_operator", "fermi_hubbard", "get_interaction_operator",
    "givens_decomposition_square", "jordan_wigner",
]
