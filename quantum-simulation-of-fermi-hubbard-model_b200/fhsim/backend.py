"""Thin object wrappers over the C-ABI handles (context, state, observable table, pool).

PyTorch / numpy arrays are only the host-side buffer types here; all statevector arithmetic
happens inside ``libfhsim.so``.
"""
from __future__ import annotations

import numpy as np

from . import _cabi
from .tables import DiagPiece, GeneratorPlan, PauliTable

C = _cabi.C


class Context:
    """One CUDA device + stream (replaces ``qml.device(...)``, reference adapt_vqe.py:299-304)."""

    def __init__(self, device: int = 0, stream=None):
        self._h = _cabi._vp()
        _cabi.check(_cabi.lib().fh_ctx_create(int(device), stream, C.byref(self._h)))
        self.device = int(device)

    def sync(self):
        _cabi.check(_cabi.lib().fh_ctx_sync(self._h))

    def info(self):
        sm, free, total = C.c_int(), C.c_size_t(), C.c_size_t()
        _cabi.check(_cabi.lib().fh_ctx_info(self._h, C.byref(sm), C.byref(free), C.byref(total)))
        return {"sm_count": sm.value, "free_bytes": free.value, "total_bytes": total.value}

    def flush_l2(self, nbytes=256 << 20):
        _cabi.check(_cabi.lib().fh_ctx_flush_l2(self._h, int(nbytes)))

    def timer_start(self):
        _cabi.check(_cabi.lib().fh_ctx_timer_start(self._h))

    def timer_stop(self) -> float:
        """Elapsed device time in milliseconds since timer_start (CUDA events on this context's stream)."""
        ms = C.c_double()
        _cabi.check(_cabi.lib().fh_ctx_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def close(self):
        if self._h:
            _cabi.lib().fh_ctx_destroy(self._h)
            self._h = _cabi._vp()

    def __del__(self):
        try:
            if _cabi.alive():
                self.close()
        except Exception:
            pass


_DEFAULT_CTX = {}


def default_context(device: int = 0) -> Context:
    if device not in _DEFAULT_CTX:
        _DEFAULT_CTX[device] = Context(device)
    return _DEFAULT_CTX[device]


class State:
    """complex128[2^n] on the device."""

    def __init__(self, ctx: Context, n_qubits: int):
        self.ctx, self.n = ctx, int(n_qubits)
        self._h = _cabi._vp()
        _cabi.check(_cabi.lib().fh_state_create(ctx._h, self.n, C.byref(self._h)))

    @classmethod
    def from_numpy(cls, ctx, vec):
        vec = np.ascontiguousarray(vec, dtype=np.complex128)
        n = int(np.log2(vec.size))
        if 1 << n != vec.size:
            raise ValueError("state length must be a power of two")
        st = cls(ctx, n)
        st.upload(vec)
        return st

    def upload(self, vec):
        vec = np.ascontiguousarray(vec, dtype=np.complex128)
        if vec.size != 1 << self.n:
            raise ValueError(f"expected {1 << self.n} amplitudes, got {vec.size}")
        _cabi.check(_cabi.lib().fh_state_from_host(self._h, vec.view(np.float64).ctypes.data_as(_cabi._f64p)))

    def numpy(self):
        out = np.empty(1 << self.n, dtype=np.complex128)
        _cabi.check(_cabi.lib().fh_state_to_host(self._h, out.view(np.float64).ctypes.data_as(_cabi._f64p)))
        return out

    def set_basis(self, index: int):
        _cabi.check(_cabi.lib().fh_state_set_basis(self._h, int(index)))

    def copy_from(self, other: "State"):
        _cabi.check(_cabi.lib().fh_state_copy(self._h, other._h))

    def inner(self, other: "State") -> complex:
        """<self|other>."""
        re, im = C.c_double(), C.c_double()
        _cabi.check(_cabi.lib().fh_state_inner(self._h, other._h, C.byref(re), C.byref(im)))
        return complex(re.value, im.value)

    def norm2(self) -> float:
        out = C.c_double()
        _cabi.check(_cabi.lib().fh_state_norm2(self._h, C.byref(out)))
        return out.value

    # immediate-mode ops ------------------------------------------------------------------------
    def apply_pair(self, x, fixmask, fixval, zeta, matrix):
        ma, mp = _cabi.f64_array(matrix)
        _cabi.check(_cabi.lib().fh_apply_pair(self._h, int(x), int(fixmask), int(fixval), int(zeta), mp))

    def apply_diag(self, z, angles):
        za, zp = _cabi.u64_array(z)
        aa, ap = _cabi.f64_array(angles)
        _cabi.check(_cabi.lib().fh_apply_diag(self._h, len(za), zp, ap))

    def apply_pauli_rotations(self, x, z, half_angles):
        """prod_m exp(-i half_angles[m] P_m), in order (literal Trotterize_generator)."""
        xa, xp = _cabi.u64_array(x)
        za, zp = _cabi.u64_array(z)
        aa, ap = _cabi.f64_array(half_angles)
        _cabi.check(_cabi.lib().fh_apply_pauli_rot_batch(self._h, len(xa), xp, zp, ap))

    def close(self):
        if self._h:
            _cabi.lib().fh_state_destroy(self._h)
            self._h = _cabi._vp()

    def __del__(self):
        try:
            if _cabi.alive():
                self.close()
        except Exception:
            pass


class DeviceTable:
    """Observable uploaded to the device (K2)."""

    def __init__(self, ctx: Context, table: PauliTable):
        self.ctx, self.table, self.n = ctx, table, table.n_qubits
        self._h = _cabi._vp()
        xa, xp = _cabi.u64_array(table.x)
        za, zp = _cabi.u64_array(table.z)
        ra, rp = _cabi.f64_array(table.coeff.real)
        ia, ip = _cabi.f64_array(table.coeff.imag)
        _cabi.check(_cabi.lib().fh_table_upload(ctx._h, self.n, len(table), xp, zp, rp, ip, C.byref(self._h)))

    def info(self):
        nt, ng = C.c_int(), C.c_int()
        _cabi.check(_cabi.lib().fh_table_info(self._h, C.byref(nt), C.byref(ng)))
        return {"n_terms": nt.value, "n_groups": ng.value}

    def tile_passes(self) -> int:
        """Shared-memory tile passes K2 uses for this table (0: the gather kernel); csrc/table_tile.cu."""
        k = C.c_int()
        _cabi.check(_cabi.lib().fh_table_tile_passes(self._h, C.byref(k)))
        return k.value

    def apply(self, state: State, out: State | None = None) -> complex:
        """out <- H state (optional); returns <state|H|state>."""
        re, im = C.c_double(), C.c_double()
        _cabi.check(_cabi.lib().fh_apply_table(self._h, state._h, out._h if out is not None else None,
                                                C.byref(re), C.byref(im)))
        return complex(re.value, im.value)

    def expval(self, state: State) -> float:
        return self.apply(state).real

    def apply_sector(self, state: State, out: State | None, n_up: int, n_dn: int, enqueue_only=False):
        """K2 on the sector-compressed copy of a state confined to the (n_up, n_dn) sector (csrc/sector_eval.cu)."""
        re, im = C.c_double(), C.c_double()
        _cabi.check(_cabi.lib().fh_apply_table_sector(self._h, state._h, out._h if out is not None else None, int(n_up), int(n_dn),
                                                      None if enqueue_only else C.byref(re), None if enqueue_only else C.byref(im)))
        return None if enqueue_only else complex(re.value, im.value)

    def close(self):
        if self._h:
            _cabi.lib().fh_table_free(self._h)
            self._h = _cabi._vp()

    def __del__(self):
        try:
            if _cabi.alive():
                self.close()
        except Exception:
            pass


class DevicePool:
    """Screening pool (K3): one output per generator, entries = its x-mask pair pieces."""

    def __init__(self, ctx: Context, plans: list[GeneratorPlan], n_qubits: int):
        self.ctx, self.n, self.n_out = ctx, int(n_qubits), len(plans)
        x, fm, fv, ze, br, bi, out = [], [], [], [], [], [], []
        for k, plan in enumerate(plans):
            if not plan.exact:
                raise NotImplementedError("pool generators must consist of mutually commuting strings")
            for piece in plan.pieces:
                if isinstance(piece, DiagPiece):
                    raise NotImplementedError("diagonal pool generators are not supported by the screening kernel")
                x.append(piece.x); fm.append(piece.fixmask); fv.append(piece.fixval); ze.append(piece.zeta)
                br.append(piece.b.real); bi.append(piece.b.imag); out.append(k)
        self._upload(x, fm, fv, ze, br, bi, out)

    @classmethod
    def from_entries(cls, ctx: Context, n_qubits: int, entries, n_out: int):
        """entries: [(x, fixmask, fixval, zeta, b, out)] sorted by ``out`` (outputs without entries give 0)."""
        self = cls.__new__(cls)
        self.ctx, self.n, self.n_out = ctx, int(n_qubits), int(n_out)
        cols = list(zip(*entries)) if entries else [[] for _ in range(6)]
        self._upload(list(cols[0]), list(cols[1]), list(cols[2]), list(cols[3]), [complex(b).real for b in cols[4]],
                     [complex(b).imag for b in cols[4]], list(cols[5]))
        return self

    def _upload(self, x, fm, fv, ze, br, bi, out):
        ctx = self.ctx
        self.n_entries = len(x)
        self._h = _cabi._vp()
        xa, xp = _cabi.u64_array(x)
        fa, fp = _cabi.u64_array(fm)
        va, vp = _cabi.u64_array(fv)
        za, zp = _cabi.u64_array(ze)
        ra, rp = _cabi.f64_array(br)
        ia, ip = _cabi.f64_array(bi)
        oa, op = _cabi.i32_array(out)
        _cabi.check(_cabi.lib().fh_pool_upload(ctx._h, self.n, self.n_entries, xp, fp, vp, zp, rp, ip, op, self.n_out,
                                               C.byref(self._h)))

    def enqueue(self, psi: State, lam: State, first=0, count=None):
        """Launch the screening kernels only (no copy-back, no sync); for timing the kernel alone."""
        if count is None:
            count = self.n_out - first
        _cabi.check(_cabi.lib().fh_pool_gradients(self._h, psi._h, lam._h, int(first), int(count), None))

    def gradients_sector(self, psi: State, lam: State, n_up: int, n_dn: int, first=0, count=None, enqueue_only=False,
                         up_mask=None, dn_mask=None):
        """K3 on sector-compressed copies of psi / lambda (states confined to the (n_up, n_dn) sector); csrc/sector_eval.cu.
        up_mask / dn_mask: index bits of the up / down orbitals when they are not the standard odd / even bits (a slab of
        a sharded state)."""
        if count is None:
            count = self.n_out - first
        out = np.zeros(max(count, 1))
        optr = None if enqueue_only else out.ctypes.data_as(_cabi._f64p)
        if up_mask is None:
            _cabi.check(_cabi.lib().fh_pool_gradients_sector(self._h, psi._h, lam._h, int(n_up), int(n_dn), int(first), int(count),
                                                             optr))
        else:
            _cabi.check(_cabi.lib().fh_pool_gradients_sector_masks(self._h, psi._h, lam._h, int(up_mask), int(dn_mask), int(n_up),
                                                                   int(n_dn), int(first), int(count), optr))
        return None if enqueue_only else out[:count]

    def gradients(self, psi: State, lam: State, first=0, count=None) -> np.ndarray:
        if count is None:
            count = self.n_out - first
        out = np.zeros(max(count, 1))
        _cabi.check(_cabi.lib().fh_pool_gradients(self._h, psi._h, lam._h, int(first), int(count),
                                                  out.ctypes.data_as(_cabi._f64p)))
        return out[:count]

    def close(self):
        if self._h:
            _cabi.lib().fh_pool_free(self._h)
            self._h = _cabi._vp()

    def __del__(self):
        try:
            if _cabi.alive():
                self.close()
        except Exception:
            pass


def lanczos(table: DeviceTable, k=1, n_up=-1, n_dn=-1, tol=1e-10, max_iter=1000, seed=7, want_vectors=True):
    """k lowest eigenpairs of the observable in the (n_up, n_dn) sector (-1, -1: full space)."""
    ctx = table.ctx
    evals = np.zeros(k)
    vecs = [State(ctx, table.n) for _ in range(k)] if want_vectors else []
    arr = (C.c_void_p * max(k, 1))(*[v._h for v in vecs]) if want_vectors else None
    iters = C.c_int()
    _cabi.check(_cabi.lib().fh_lanczos(table._h, int(n_up), int(n_dn), int(k), float(tol), int(max_iter), int(seed),
                                       evals.ctypes.data_as(_cabi._f64p), arr, C.byref(iters)))
    return evals, vecs, iters.value


def lanczos_sector(table: DeviceTable, n_up, n_dn, k=1, tol=1e-10, max_iter=1000, seed=7, want_vectors=False,
                   want_compressed=False):
    """k lowest eigenpairs in the (n_up, n_dn) sector on SECTOR-COMPRESSED vectors (``fh_lanczos_sector``): the route
    for lattices whose full state does not fit one GPU (4x4: 2.65 GB per vector instead of 64 GiB).
    Returns (evals, full-space States or [], compressed eigenvectors [k, dim_sector] or None, info dict)."""
    from math import comb
    ctx = table.ctx
    half = table.n // 2
    dim = comb(half, int(n_up)) * comb(half, int(n_dn))
    evals = np.zeros(k)
    vecs = [State(ctx, table.n) for _ in range(k)] if want_vectors else []
    arr = (C.c_void_p * max(k, 1))(*[v._h for v in vecs]) if want_vectors else None
    comp = np.zeros((k, dim), dtype=np.complex128) if want_compressed else None
    iters = C.c_int()
    stats = np.zeros(4)
    _cabi.check(_cabi.lib().fh_lanczos_sector(
        table._h, int(n_up), int(n_dn), int(k), float(tol), int(max_iter), int(seed), evals.ctypes.data_as(_cabi._f64p), arr,
        comp.ctypes.data_as(_cabi._f64p) if want_compressed else None, C.byref(iters), stats.ctypes.data_as(_cabi._f64p)))
    info = {"sector_dim": int(stats[0]), "loop_seconds": float(stats[1]), "matvecs": int(stats[2]), "host_syncs": int(stats[3]),
            "iterations": iters.value}
    return evals, vecs, comp, info


def sector_indices(n, n_up, n_dn):
    """Full 2^n index of every compressed amplitude in rank order (rank = rank_up * D_dn + rank_dn; up = even wires)."""
    from itertools import combinations
    half = n // 2

    def patterns(count, shift):
        out = []
        for pat in range(1 << half):
            if bin(pat).count("1") == count:
                out.append(sum(1 << (2 * b + shift) for b in range(half) if pat >> b & 1))
        return np.array(out, dtype=np.uint64)
    up, dn = patterns(n_up, 1), patterns(n_dn, 0)
    return (up[:, None] | dn[None, :]).reshape(-1)


class DevicePauliTable:
    """Packed (x, z, coeff) term table resident on the device, for the iQCC dressing update
    (``fh_ptable_dress``: reference models/iqcc_hubbard.py:184-189 without symbolic operator products)."""

    def __init__(self, ctx: Context, table: PauliTable):
        self.ctx, self.n = ctx, table.n_qubits
        self._h = _cabi._vp()
        xa, xp = _cabi.u64_array(table.x)
        za, zp = _cabi.u64_array(table.z)
        ra, rp = _cabi.f64_array(np.real(table.coeff))
        ia, ip = _cabi.f64_array(np.imag(table.coeff))
        _cabi.check(_cabi.lib().fh_ptable_upload(ctx._h, self.n, len(xa), xp, zp, rp, ip, C.byref(self._h)))

    def __len__(self):
        n = C.c_int()
        _cabi.check(_cabi.lib().fh_ptable_size(self._h, C.byref(n)))
        return n.value

    def dress(self, xp: int, zp: int, tau: float, tol: float = 1e-12):
        """In place: exp(i tau P/2) H exp(-i tau P/2) for the Pauli string P = (xp, zp)."""
        _cabi.check(_cabi.lib().fh_ptable_dress(self._h, int(xp), int(zp), float(tau), float(tol)))
        return self

    def to_host(self) -> PauliTable:
        n = len(self)
        x, z = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        cr, ci = np.zeros(n), np.zeros(n)
        if n:
            _cabi.check(_cabi.lib().fh_ptable_download(self._h, x.ctypes.data_as(_cabi._u64p), z.ctypes.data_as(_cabi._u64p),
                                                       cr.ctypes.data_as(_cabi._f64p), ci.ctypes.data_as(_cabi._f64p)))
        return PauliTable(self.n, x, z, cr + 1j * ci)

    def close(self):
        if self._h:
            _cabi.lib().fh_ptable_free(self._h)
            self._h = _cabi._vp()

    def __del__(self):
        try:
            if _cabi.alive():
                self.close()
        except Exception:
            pass
