"""Sector-restricted exact diagonalisation (reference ``linalg/exact_diagonalization.py``).

The reference builds OpenFermion's 2^n x 2^n sparse matrix, slices the (N_up, N_dn) block and calls
ARPACK (``scipy.sparse.linalg.eigsh(which='SA')``, reference :34-51, :181-229).  Here the
"sparse operator" is a :class:`SparseOperator` handle around the packed Pauli table and the
eigenproblem is solved matrix-free by Lanczos on the fused H|psi> CUDA kernel (``fh_lanczos``):
start vector supported on the sector, full re-orthogonalisation, deflation for degenerate levels.
Eigenvectors are defined up to a phase / a rotation inside degenerate levels (ARPACK's own start
vector is random too) -- compare energies and subspace projectors, not vectors.
"""
from __future__ import annotations

import itertools

import numpy as np

from fhsim.backend import DeviceTable, default_context, lanczos, lanczos_sector, sector_indices
from fhsim.symbolic import FermionOperator, QubitOperator, count_qubits, jordan_wigner
from fhsim.tables import PauliTable


class SparseOperator:
    """Stand-in for ``openfermion.get_sparse_operator(op)``: a 2^n x 2^n Hermitian operator kept in
    packed Pauli form (never materialised)."""

    def __init__(self, operator, n_qubits=None):
        if isinstance(operator, FermionOperator):
            operator = jordan_wigner(operator)
        elif hasattr(operator, "one_body_tensor"):
            operator = jordan_wigner(operator)
        if not isinstance(operator, QubitOperator):
            raise TypeError("SparseOperator needs a FermionOperator, InteractionOperator or QubitOperator")
        self.operator = operator
        self.n_qubits = int(n_qubits) if n_qubits is not None else count_qubits(operator)
        self.shape = (1 << self.n_qubits, 1 << self.n_qubits)
        self._device_table = None

    def table(self) -> PauliTable:
        return PauliTable.from_operator(self.operator, self.n_qubits)

    def device_table(self, ctx=None) -> DeviceTable:
        if self._device_table is None:
            self._device_table = DeviceTable(ctx or default_context(), self.table())
        return self._device_table


def get_sparse_operator(operator, n_qubits=None) -> SparseOperator:
    return SparseOperator(operator, n_qubits)


def jw_number_spin_indices(n_electrons, spin_up, spin_down, n_qubits):
    """Flat indices (wire 0 = MSB) of the basis states with the given spin populations, ascending."""
    if spin_up + spin_down != n_electrons:
        raise ValueError('spin up plus spin down must equal to n_electrons!')
    up_orbitals = range(0, n_qubits, 2)
    down_orbitals = range(1, n_qubits, 2)
    indices = []
    for ups in itertools.combinations(up_orbitals, spin_up):
        up_part = sum(1 << (n_qubits - q - 1) for q in ups)
        for downs in itertools.combinations(down_orbitals, spin_down):
            indices.append(up_part + sum(1 << (n_qubits - q - 1) for q in downs))
    return sorted(indices)


def jw_number_spin_restrict_operator(operator, n_electrons, spin_up, spin_down, n_qubits=None):
    """Slice a (scipy / numpy) 2^n x 2^n matrix to the sector block (host utility kept for API parity)."""
    if n_qubits is None:
        n_qubits = int(np.log2(operator.shape[0]))
    idx = jw_number_spin_indices(n_electrons, spin_up, spin_down, n_qubits)
    return operator[np.ix_(idx, idx)]


def _as_handle(sparse_operator) -> SparseOperator:
    if isinstance(sparse_operator, SparseOperator):
        return sparse_operator
    if isinstance(sparse_operator, (FermionOperator, QubitOperator)) or hasattr(sparse_operator, "one_body_tensor"):
        return SparseOperator(sparse_operator)
    raise TypeError("pass get_sparse_operator(op) (or the symbolic operator itself); dense/scipy matrices "
                    "are not accepted by the GPU eigensolver")


def _lowest(sparse_operator, particle_number, spin_up, spin_down, k, tol, seed):
    handle = _as_handle(sparse_operator)
    if spin_up + spin_down != particle_number:
        raise ValueError('spin up plus spin down must equal to n_electrons!')
    evals, vecs, _ = lanczos(handle.device_table(), k=k, n_up=spin_up, n_dn=spin_down, tol=tol,
                             max_iter=2000, seed=seed)
    order = np.argsort(evals)
    return evals[order], [vecs[i].numpy() for i in order]


def jw_get_ground_state(sparse_operator, particle_number, spin_up, spin_down, tol=1e-11, seed=7):
    """-> (E0, ground state as a length-2^n complex vector)."""
    handle = _as_handle(sparse_operator)
    if handle.n_qubits > 30:
        raise MemoryError(f"a {handle.n_qubits}-qubit state vector does not fit one device / the host as a dense array; "
                          "use jw_get_ground_state_compressed (sector-compressed eigenvector)")
    evals, vecs = _lowest(handle, particle_number, spin_up, spin_down, 1, tol, seed)
    return float(evals[0]), vecs[0]


class SectorVector:
    """Eigenvector stored on its (N_up, N_dn) sector only: ``amplitudes[r]`` belongs to the flat index ``indices()[r]``
    (wire 0 = MSB).  4x4: 165 636 900 amplitudes (2.65 GB) instead of 2^32 (64 GiB)."""

    def __init__(self, n_qubits, spin_up, spin_down, amplitudes):
        self.n_qubits, self.spin_up, self.spin_down = n_qubits, spin_up, spin_down
        self.amplitudes = amplitudes

    def indices(self):
        return sector_indices(self.n_qubits, self.spin_up, self.spin_down)

    def to_dense(self):
        if self.n_qubits > 30:
            raise MemoryError("dense form of a >30-qubit state requested")
        out = np.zeros(1 << self.n_qubits, dtype=np.complex128)
        out[self.indices().astype(np.int64)] = self.amplitudes
        return out


def jw_get_ground_state_compressed(sparse_operator, particle_number, spin_up, spin_down, tol=1e-10, seed=7, max_iter=2000,
                                   want_vector=True):
    """-> (E0, SectorVector, info): the ground state of the sector by Lanczos on sector-compressed vectors
    (``fh_lanczos_sector``).  The route for the 4x4 lattice (32 qubits) of reference adapt_vqe.py:221-247."""
    handle = _as_handle(sparse_operator)
    if spin_up + spin_down != particle_number:
        raise ValueError('spin up plus spin down must equal to n_electrons!')
    evals, _, comp, info = lanczos_sector(handle.device_table(), spin_up, spin_down, k=1, tol=tol, max_iter=max_iter, seed=seed,
                                          want_compressed=want_vector)
    vec = SectorVector(handle.n_qubits, spin_up, spin_down, comp[0]) if want_vector else None
    return float(evals[0]), vec, info


def jw_get_ground_state_for_3x3(sparse_operator, particle_number, spin_up, spin_down, tol=1e-11, seed=7):
    """Lowest 10 levels printed, the 4 lowest states returned orthonormalised (4-fold ground level of 3x3)."""
    handle = _as_handle(sparse_operator)
    dim = len(jw_number_spin_indices(particle_number, spin_up, spin_down, handle.n_qubits))
    k = min(10, dim)
    evals, vecs = _lowest(handle, particle_number, spin_up, spin_down, k, tol, seed)
    for value in evals:
        print(value)
    basis = []
    for v in vecs[:4]:
        w = v.copy()
        for u in basis:
            w = w - (u.conj() @ v) / (u.conj() @ u) * u
        basis.append(w / np.linalg.norm(w))
    return float(evals[0]), basis


def get_ground_state(sparse_operator, tol=1e-11, seed=7):
    """``openfermion.get_ground_state``: lowest eigenpair over the full Fock space (reference iqcc_hubbard.py:57)."""
    handle = _as_handle(sparse_operator)
    evals, vecs, _ = lanczos(handle.device_table(), k=1, n_up=-1, n_dn=-1, tol=tol, max_iter=2000, seed=seed)
    return float(evals[0]), vecs[0].numpy()
