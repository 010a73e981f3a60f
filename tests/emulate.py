"""Test-only numpy interpreter of the device op semantics (pair / diag ops), used to validate the
host compile step (tables.py, circuit.py, the tile scheduler) on a CPU-only box."""
import numpy as np

from fhsim.circuit import DiagOpSpec, Marker, PairOpSpec


def _par(v):
    return (np.bitwise_count(v) & 1).astype(np.int64)


def pair_matrix(op, thetas):
    if op.kind == 0:
        m = op.matrix
        return [complex(m[2 * i], m[2 * i + 1]) for i in range(4)]
    a = op.scale * (thetas[op.param] if op.param >= 0 else 1.0)
    c, s = np.cos(a), np.sin(a)
    b = complex(op.bhat)
    return [c, -1j * s * b, -1j * s * np.conj(b), c]


def apply_op(psi, op, thetas, n):
    idx = np.arange(1 << n, dtype=np.uint64)
    if isinstance(op, DiagOpSpec):
        tot = np.zeros(1 << n)
        for z, c in zip(op.z, op.coef):
            a = c * (thetas[op.param] if op.param >= 0 else 1.0)
            tot += a * (1 - 2 * _par(idx & np.uint64(z)))
        return psi * np.exp(-1j * tot)
    m00, m01, m10, m11 = pair_matrix(op, thetas)
    sel = (idx & np.uint64(op.fixmask)) == np.uint64(op.fixval)
    i = idx[sel]
    j = i ^ np.uint64(op.x)
    s = 1 - 2 * _par(i & np.uint64(op.zeta))
    out = psi.copy()
    out[i] = m00 * psi[i] + s * m01 * psi[j]
    out[j] = s * m10 * psi[i] + m11 * psi[j]
    return out


def run_circuit(circuit, psi, thetas=()):
    for op in circuit.ops:
        if isinstance(op, Marker):
            continue
        psi = apply_op(psi, op, thetas, circuit.n)
    return psi


def run_items(items, psi, thetas, n):
    for item in items:
        ops = [item[1]] if item[0] == "op" else item[2]
        if item[0] == "tile":
            bits = 0
            for b in item[1]:
                bits |= 1 << b
            for o in ops:
                assert o.tile_bits & ~bits == 0, "pair op x-mask escapes its tile"
        for o in ops:
            psi = apply_op(psi, o, thetas, n)
    return psi
