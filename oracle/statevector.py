"""ORACLE (test infrastructure only -- never imported by the product path).

numpy complex128 restatement of the statevector arithmetic of the reference's hot path in
*closed form* (packed masks):

* ``pauli_rotation``            exp(-i theta/2 P)  == the gate recipe of reference ``models/utils.py:58-83``
* ``trotterize``                reference ``models/adapt_vqe.py:87-98`` (product over ``generator.terms``)
* ``apply_table`` / ``expval``  ``qml.expval(qml.Hamiltonian)``  (``adapt_vqe.py:357,361``)
* gates ``rx ry rz cnot pauli_x single_excitation``: PennyLane matrix conventions
* ``basis_change``              the W network of ``adapt_vqe.py:344-354``
* ``pool_gradients``            d<H>/d e_k at e = 0 of the eval-mode circuit (``adapt_vqe.py:297-310``)
                                 = 2 Im <lambda| G_k |psi>, lambda = W† H W psi
* ``adjoint_gradient``          d<H>/d theta_j of the train-mode circuit (``adapt_vqe.py:415-418``)

PennyLane is a third-party dependency absent from /root/reference (unpinned); PARITY UNPINNED by
the reference -- pinned to analytic known answers in tests/.  The literal gate-by-gate form with
autograd lives in oracle/literal.py and is cross-checked against this file.
"""
from __future__ import annotations

import numpy as np

_I_POW = np.array([1, 1j, -1, -1j])


def _indices(n):
    return np.arange(1 << n, dtype=np.uint64)


def _parity(v):
    return (np.bitwise_count(v) & 1).astype(np.int64)


def basis_state(n, occupied_wires):
    psi = np.zeros(1 << n, dtype=np.complex128)
    idx = 0
    for q in occupied_wires:
        idx |= 1 << (n - 1 - q)
    psi[idx] = 1.0
    return psi


def apply_pauli(psi, x, z, n):
    """(P psi)[i] = i^k (-1)^popcount((i^x)&z) psi[i^x], k = popcount(x&z)."""
    i = _indices(n)
    j = i ^ np.uint64(x)
    k = bin(int(x) & int(z)).count("1") & 3
    sign = 1 - 2 * _parity(j & np.uint64(z))
    return _I_POW[k] * sign * psi[j]


def pauli_rotation(psi, theta, x, z, n):
    """exp(-i theta/2 P) psi."""
    if x == 0 and z == 0:
        return psi * np.exp(-0.5j * theta)
    return np.cos(theta / 2) * psi - 1j * np.sin(theta / 2) * apply_pauli(psi, x, z, n)


def trotterize(psi, theta, table, n):
    """prod_m exp(-i theta Re(c_m) P_m) in table (dict) order; identity term skipped."""
    for (x, z), c in table.items():
        if x == 0 and z == 0:
            continue
        psi = pauli_rotation(psi, 2.0 * theta * complex(c).real, x, z, n)
    return psi


def apply_table(psi, table, n):
    out = np.zeros_like(psi)
    for (x, z), c in table.items():
        out += c * apply_pauli(psi, x, z, n)
    return out


def expval(psi, table, n):
    return np.vdot(psi, apply_table(psi, table, n))


# --- elementary gates (PennyLane conventions), wire w <-> bit n-1-w -----------------------
def _apply_1q(psi, mat, wire, n):
    b = n - 1 - wire
    t = psi.reshape(1 << wire, 2, 1 << b)
    return np.einsum("ab,xby->xay", mat, t).reshape(-1)


def rx(psi, phi, wire, n):
    c, s = np.cos(phi / 2), np.sin(phi / 2)
    return _apply_1q(psi, np.array([[c, -1j * s], [-1j * s, c]]), wire, n)


def ry(psi, phi, wire, n):
    c, s = np.cos(phi / 2), np.sin(phi / 2)
    return _apply_1q(psi, np.array([[c, -s], [s, c]], dtype=complex), wire, n)


def rz(psi, phi, wire, n):
    return _apply_1q(psi, np.diag([np.exp(-0.5j * phi), np.exp(0.5j * phi)]), wire, n)


def pauli_x(psi, wire, n):
    return _apply_1q(psi, np.array([[0, 1], [1, 0]], dtype=complex), wire, n)


def cnot(psi, control, target, n):
    i = _indices(n)
    cb, tb = np.uint64(1 << (n - 1 - control)), np.uint64(1 << (n - 1 - target))
    src = np.where((i & cb) != 0, i ^ tb, i)
    return psi[src]


def single_excitation(psi, phi, wire_i, wire_j, n):
    """|01> -> c|01> + s|10>, |10> -> -s|01> + c|10>  in the basis |q_i q_j>."""
    c, s = np.cos(phi / 2), np.sin(phi / 2)
    i = _indices(n)
    bi, bj = np.uint64(1 << (n - 1 - wire_i)), np.uint64(1 << (n - 1 - wire_j))
    qi, qj = (i & bi) != 0, (i & bj) != 0
    partner = psi[i ^ bi ^ bj]
    out = psi.copy()
    m01 = (~qi) & qj            # |01>: out = c*a01 - s*a10
    m10 = qi & (~qj)            # |10>: out = s*a01 + c*a10
    out[m01] = c * psi[m01] - s * partner[m01]
    out[m10] = s * partner[m10] + c * psi[m10]
    return out


def basis_change(psi, diagonal, decomposition, n, inverse=False):
    """W: RZ(angle(diagonal[q])) on every wire, then the layers reversed, each (i, j, theta, phi) as
    SingleExcitation(2 theta, [i, j]) then RZ(phi, j)   (reference adapt_vqe.py:344-354)."""
    gates = [("rz", q, float(np.angle(diagonal[q]))) for q in range(len(diagonal))]
    for layer in reversed(list(decomposition)):
        for (i, j, theta, phi) in layer:
            gates.append(("se", i, j, 2 * theta))
            gates.append(("rz", j, phi))
    if inverse:
        gates = [(g[0], *g[1:-1], -g[-1]) for g in reversed(gates)]
    for g in gates:
        if g[0] == "rz":
            psi = rz(psi, g[2], g[1], n)
        else:
            psi = single_excitation(psi, g[3], g[1], g[2], n)
    return psi


# --- ADAPT circuit pieces ------------------------------------------------------------------
def adapt_state(n, occupied, generators, thetas):
    """k-space HF state, then the selected generators (train/state mode before W)."""
    psi = basis_state(n, occupied)
    for table, theta in zip(generators, thetas):
        psi = trotterize(psi, theta, table, n)
    return psi


def pool_gradients(psi_k, h_table, pool_tables, diagonal, decomposition, n):
    """g_k = 2 Im <lambda| G_k |psi>, lambda = W† H W psi;  also returns E = <psi|lambda>."""
    phi = basis_change(psi_k, diagonal, decomposition, n)
    lam = basis_change(apply_table(phi, h_table, n), diagonal, decomposition, n, inverse=True)
    energy = np.vdot(psi_k, lam).real
    grads = np.array([2.0 * np.vdot(lam, apply_table(psi_k, g, n)).imag for g in pool_tables])
    return grads, energy, lam


def adjoint_gradient(n, occupied, generators, thetas, h_table, diagonal, decomposition):
    """E and dE/dtheta_j for psi = W prod_j exp(-i theta_j G_j)|occ> by the adjoint sweep,
    using the literal Trotter product for each U_j."""
    psi = adapt_state(n, occupied, generators, thetas)
    phi = basis_change(psi, diagonal, decomposition, n)
    lam = basis_change(apply_table(phi, h_table, n), diagonal, decomposition, n, inverse=True)
    energy = np.vdot(psi, lam).real
    grads = np.zeros(len(generators))
    for j in range(len(generators) - 1, -1, -1):
        g = {k: v for k, v in generators[j].items() if k != (0, 0)}
        grads[j] = 2.0 * np.vdot(lam, apply_table(psi, g, n)).imag
        inv = dict(reversed(list(generators[j].items())))
        psi = trotterize(psi, -thetas[j], inv, n)
        lam = trotterize(lam, -thetas[j], inv, n)
    return energy, grads


def energy_of(n, occupied, generators, thetas, h_table, diagonal, decomposition):
    psi = adapt_state(n, occupied, generators, thetas)
    phi = basis_change(psi, diagonal, decomposition, n)
    return expval(phi, h_table, n).real


def trotter_circuit_gradient(psi0, layers, thetas, h_table, n):
    """E = <psi|H|psi> and dE/dtheta_p for psi = prod_l Trotterize(theta_{p_l}, G_l) psi0, layers = [(p_l, table_l)]
    applied first to last; several layers may share one parameter (HVA: reference hva.py:283-298 applies one theta per
    layer of bonds, one theta_U per Coulomb layer).  Exact derivative of the LITERAL Trotter product
    prod_m exp(-i theta c_m P_m) (adapt_vqe.py:87-98), string by string -- no commutation assumption:
        d/dtheta exp(-i theta c P) = -i c P exp(-i theta c P)   =>   dE/dtheta += 2 c Im <lambda_m| P_m |psi_m>
    with psi_m, lambda_m the state and H|psi_final> pulled back to just after string m."""
    thetas = np.asarray(thetas, dtype=np.float64)
    psi = psi0.copy()
    for p, table in layers:
        psi = trotterize(psi, thetas[p], table, n)
    lam = apply_table(psi, h_table, n)
    energy = np.vdot(psi, lam).real
    grads = np.zeros(len(thetas))
    for p, table in reversed(layers):
        for (x, z), c in reversed(list(table.items())):
            if x == 0 and z == 0:
                continue
            cr = complex(c).real
            grads[p] += 2.0 * cr * np.vdot(lam, apply_pauli(psi, x, z, n)).imag
            psi = pauli_rotation(psi, -2.0 * thetas[p] * cr, x, z, n)
            lam = pauli_rotation(lam, -2.0 * thetas[p] * cr, x, z, n)
    return energy, grads
