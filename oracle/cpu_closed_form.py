"""ORACLE / CPU BASELINE (test infrastructure only): ctypes binding of oracle/cpu_closed_form.c and a runner that executes
a host-compiled op list (the same ``fhsim.circuit.Circuit.ops`` the CUDA path is compiled from) on the host cores.

Only tests/, __graft_entry__ and bench.py's CPU-baseline leg may import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_closed_form.c")
LIB = os.path.join(HERE, "_build", "libcpu_closed_form.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        u64p, f64p, vp = C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.c_void_p
        L.cf_set_threads.argtypes = [C.c_int]
        L.cf_max_threads.restype = C.c_int
        L.cf_pair.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, f64p]
        L.cf_diag.argtypes = [vp, C.c_int, C.c_int, u64p, f64p]
        L.cf_table.argtypes = [vp, vp, C.c_int, C.c_int, u64p, u64p, f64p, f64p]
        L.cf_inner_re.argtypes = [vp, vp, C.c_int]
        L.cf_inner_re.restype = C.c_double
        L.cf_pool.argtypes = [vp, vp, C.c_int, C.c_int, u64p, u64p, u64p, u64p, f64p, f64p, f64p]
        _lib = L
    return _lib


def _u64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint64))


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _matrix(op, thetas, dagger):
    if op.kind == 0:
        m = [complex(op.matrix[2 * i], op.matrix[2 * i + 1]) for i in range(4)]
    else:
        a = op.scale * (thetas[op.param] if op.param >= 0 else 1.0)
        c, s, b = np.cos(a), np.sin(a), complex(op.bhat)
        m = [c, -1j * s * b, -1j * s * np.conj(b), c]
    if dagger:
        m = [np.conj(m[0]), np.conj(m[2]), np.conj(m[1]), np.conj(m[3])]
    return np.array([v for c in m for v in (complex(c).real, complex(c).imag)], dtype=np.float64)


def run_ops(psi, ops, thetas, n, dagger=False):
    """Apply ``fhsim.circuit`` op specs (pair / diag) to the numpy state in place; markers are skipped."""
    L = lib()
    seq = [op for op in ops if hasattr(op, "x") or hasattr(op, "z")]
    for op in (reversed(seq) if dagger else seq):
        if hasattr(op, "x"):
            m, mp = _f64(_matrix(op, thetas, dagger))
            L.cf_pair(psi.ctypes.data, n, int(op.x), int(op.fixmask), int(op.fixval), int(op.zeta), mp)
        else:
            scale = (thetas[op.param] if op.param >= 0 else 1.0) * (-1.0 if dagger else 1.0)
            z, zp = _u64(op.z)
            a, ap = _f64(np.asarray(op.coef, dtype=np.float64) * scale)
            L.cf_diag(psi.ctypes.data, n, len(z), zp, ap)
    return psi


def apply_table(psi, table, n):
    """table: fhsim.tables.PauliTable (x, z, coeff arrays)."""
    out = np.empty_like(psi)
    x, xp = _u64(table.x)
    z, zp = _u64(table.z)
    cr, crp = _f64(np.real(table.coeff))
    ci, cip = _f64(np.imag(table.coeff))
    lib().cf_table(psi.ctypes.data, out.ctypes.data, n, len(x), xp, zp, crp, cip)
    return out


def pool_gradients(psi, lam, plans, n):
    """plans: fhsim.tables.GeneratorPlan list with one pair piece each (the drivers' pool)."""
    pieces = [p.pieces for p in plans]
    assert all(len(ps) == 1 for ps in pieces), "closed-form CPU pool expects one pair piece per operator"
    ps = [p[0] for p in pieces]
    x, xp = _u64([q.x for q in ps])
    fm, fmp = _u64([q.fixmask for q in ps])
    fv, fvp = _u64([q.fixval for q in ps])
    ze, zep = _u64([q.zeta for q in ps])
    br, brp = _f64([complex(q.b).real for q in ps])
    bi, bip = _f64([complex(q.b).imag for q in ps])
    out = np.zeros(len(ps))
    lib().cf_pool(psi.ctypes.data, lam.ctypes.data, n, len(ps), xp, fmp, fvp, zep, brp, bip,
                  out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def screening(n, basis_index, ansatz_ops, w_ops, thetas, h_table, plans, threads):
    """One ADAPT screening step on the CPU: psi = ansatz|basis>, phi = W psi, lambda = W^dagger H phi, g_k; returns
    (energy, gradients)."""
    L = lib()
    L.cf_set_threads(int(threads))
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[basis_index] = 1.0
    run_ops(psi, ansatz_ops, thetas, n)
    phi = psi.copy()
    run_ops(phi, w_ops, thetas, n)
    lam = apply_table(phi, h_table, n)
    energy = L.cf_inner_re(phi.ctypes.data, lam.ctypes.data, n)
    run_ops(lam, w_ops, thetas, n, dagger=True)
    return energy, pool_gradients(psi, lam, plans, n)
