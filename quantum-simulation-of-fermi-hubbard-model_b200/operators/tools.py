"""Split a fermionic Hamiltonian by many-body order (reference ``operators/tools.py:3-23``)."""
from fhsim.symbolic import FermionOperator


def _select_terms(operator: FermionOperator, keep) -> FermionOperator:
    picked = FermionOperator()
    for single in operator.get_operators():
        if keep(single.many_body_order()):
            picked += single
    picked.compress()
    return picked


def get_quadratic_term(operator: FermionOperator) -> FermionOperator:
    """Terms with exactly two ladder operators (hopping / on-site energies)."""
    return _select_terms(operator, lambda order: order == 2)


def get_interacting_term(operator: FermionOperator) -> FermionOperator:
    """Terms with more than two ladder operators (the Hubbard U part)."""
    return _select_terms(operator, lambda order: order > 2)
