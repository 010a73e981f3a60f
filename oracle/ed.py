"""ORACLE (test infrastructure only).

Sector-restricted exact diagonalisation as the reference does it
(``linalg/exact_diagonalization.py:11-51, 181-229``): indices of the (N_up, N_dn) sector with
wire 0 = MSB, restriction of the sparse operator, ``scipy.sparse.linalg.eigsh(which='SA')``.
The 2^n x 2^n matrix of OpenFermion's ``get_sparse_operator`` is replaced by a direct build of the
sector block from the packed Pauli table (same matrix elements).
"""
import itertools

import numpy as np
import scipy.sparse
import scipy.sparse.linalg


def jw_number_spin_indices(n_electrons, spin_up, spin_down, n_qubits):
    if spin_up + spin_down != n_electrons:
        raise ValueError('spin up plus spin down must equal to n_electrons!')
    picked = []
    for occ in itertools.combinations(range(n_qubits), n_electrons):
        if sum(1 for q in occ if q % 2 == 0) == spin_up:
            picked.append(occ)
    return [sum(2 ** (n_qubits - q - 1) for q in occ) for occ in reversed(picked)]


def sector_matrix(table, indices, n):
    idx = np.array(indices, dtype=np.uint64)
    order = np.argsort(idx)
    sorted_idx = idx[order]
    rows, cols, data = [], [], []
    for (x, z), c in table.items():
        j = idx ^ np.uint64(x)                       # column index: <i|P|j> with j = i^x
        pos = np.searchsorted(sorted_idx, j)
        pos = np.clip(pos, 0, len(idx) - 1)
        ok = sorted_idx[pos] == j
        k = bin(int(x) & int(z)).count("1") & 3
        sign = 1 - 2 * (np.bitwise_count(j & np.uint64(z)) & 1).astype(np.int64)
        val = c * np.array([1, 1j, -1, -1j])[k] * sign
        rows.append(np.nonzero(ok)[0])
        cols.append(order[pos[ok]])
        data.append(val[ok])
    rows, cols, data = map(np.concatenate, (rows, cols, data))
    m = scipy.sparse.csr_matrix((data, (rows, cols)), shape=(len(idx), len(idx)))
    m.sum_duplicates()
    return m


def ground_state(table, n, n_electrons, spin_up, spin_down, k=1, dense_below=3000):
    """-> (eigenvalues ascending (k of them), eigenvectors expanded to 2^n, sector indices)."""
    indices = jw_number_spin_indices(n_electrons, spin_up, spin_down, n)
    m = sector_matrix(table, indices, n)
    dim = m.shape[0]
    if dim <= dense_below:
        vals, vecs = np.linalg.eigh(m.toarray())
        vals, vecs = vals[:k], vecs[:, :k]
    else:
        vals, vecs = scipy.sparse.linalg.eigsh(m, k=k, which='SA')
        o = np.argsort(vals)
        vals, vecs = vals[o], vecs[:, o]
    full = np.zeros((k, 1 << n), dtype=complex)
    full[:, indices] = vecs.T
    return vals, full, indices
