"""Circuit description on the host and its compilation into a device program.

A :class:`Circuit` is the ordered op list the reference builds by calling ``qml.*`` inside
``circuit()`` (``models/adapt_vqe.py:325-361``, ``hva.py:273-303``, ``iqcc_hubbard.py:59-80``,
``vqe_hea.py:43-57``), but already lowered to the two device op kinds:

* pair op  -- a 2x2 block on index pairs (i, i^x) selected by a bit pattern, with a parity sign;
* diag op  -- a phase exp(-i sum_m a_m (-1)^popcount(i & z_m)).

``compile`` schedules runs of ops into shared-memory tiles (one global read+write per run) and hands
the result to ``libfhsim`` through the C-ABI.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np

from . import _cabi
from .tables import DiagPiece, GeneratorPlan, PairPiece, popcount, qubit_bit, strings_commute

_X = (0.0, 0.0, 1.0, 0.0, 1.0, 0.0, 0.0, 0.0)      # [[0,1],[1,0]] as re/im pairs


@dataclass
class PairOpSpec:
    x: int
    fixmask: int
    fixval: int
    zeta: int
    kind: int = 0                 # 0 fixed matrix, 1 rotation exp(-i scale*theta Ghat)
    param: int = -1
    scale: float = 0.0
    bhat: complex = 1.0
    matrix: tuple = _X
    strings: list = field(default_factory=list)

    @property
    def tile_bits(self):
        return self.x


@dataclass
class DiagOpSpec:
    z: list
    coef: list
    param: int = -1
    strings: list = field(default_factory=list)

    @property
    def tile_bits(self):
        return 0


@dataclass
class Marker:
    name: str


def _ops_commute(a, b) -> bool:
    for s in a.strings:
        for t in b.strings:
            if not strings_commute(s, t):
                return False
    return True


def _normalise_pair(x, fixmask, fixval, zeta, m):
    """Make the pattern side the one whose top x bit is 0 (swap roles if needed)."""
    top = 1 << (x.bit_length() - 1)
    fixmask |= top
    if fixval & top:
        s = -1.0 if popcount(x & zeta) & 1 else 1.0
        fixval = (fixval ^ x) & fixmask
        m00, m01, m10, m11 = [complex(m[2 * i], m[2 * i + 1]) for i in range(4)]
        m00, m01, m10, m11 = m11, s * m10, s * m01, m00
        m = (m00.real, m00.imag, m01.real, m01.imag, m10.real, m10.imag, m11.real, m11.imag)
    return x, fixmask, fixval, zeta, tuple(m)


class Circuit:
    """Ordered list of device-level ops over ``n_qubits`` with ``n_params`` real parameters."""

    def __init__(self, n_qubits: int, n_params: int = 0):
        self.n = int(n_qubits)
        self.n_params = int(n_params)
        self.ops = []

    # -- elementary gates (PennyLane conventions, wire 0 = MSB) ---------------------------------
    def _bit(self, wire):
        if not 0 <= wire < self.n:
            raise ValueError(f"wire {wire} outside a {self.n}-qubit register")
        return qubit_bit(wire, self.n)

    def marker(self, name):
        self.ops.append(Marker(name))

    def pauli_x(self, wire):
        b = self._bit(wire)
        self.ops.append(PairOpSpec(b, b, 0, 0, matrix=_X, strings=[(b, 0)]))

    def cnot(self, control, target):
        c, t = self._bit(control), self._bit(target)
        self.ops.append(PairOpSpec(t, t | c, c, 0, matrix=_X, strings=[(t, 0), (0, c)]))

    def rz(self, angle, wire, param=-1, coef=0.5):
        """exp(-i angle Z/2); with ``param >= 0`` the angle is theta[param] (coef = 1/2)."""
        b = self._bit(wire)
        if param >= 0:
            self.ops.append(DiagOpSpec([b], [coef], param, [(0, b)]))
        else:
            self.ops.append(DiagOpSpec([b], [0.5 * float(angle)], -1, [(0, b)]))

    def _rot_1q(self, wire, letter, angle, param):
        b = self._bit(wire)
        z = b if letter == "Y" else 0
        bhat = -1j if letter == "Y" else 1.0
        if param >= 0:
            self.ops.append(PairOpSpec(b, b, 0, 0, kind=1, param=param, scale=0.5, bhat=bhat, strings=[(b, z)]))
        else:
            c, s = math.cos(0.5 * angle), math.sin(0.5 * angle)
            m01, m10 = -1j * s * bhat, -1j * s * np.conj(bhat)
            m = (c, 0.0, m01.real, m01.imag, m10.real, m10.imag, c, 0.0)
            self.ops.append(PairOpSpec(b, b, 0, 0, matrix=m, strings=[(b, z)]))

    def rx(self, angle, wire, param=-1):
        self._rot_1q(wire, "X", angle, param)

    def ry(self, angle, wire, param=-1):
        self._rot_1q(wire, "Y", angle, param)

    def single_excitation(self, phi, wire_i, wire_j):
        """qml.SingleExcitation(phi, [i, j]): |01> -> c|01> + s|10>, |10> -> -s|01> + c|10>."""
        bi, bj = self._bit(wire_i), self._bit(wire_j)
        c, s = math.cos(0.5 * phi), math.sin(0.5 * phi)
        x = bi | bj
        # pattern side |q_i q_j> = |01> : q_j set, q_i clear
        m = (c, 0.0, -s, 0.0, s, 0.0, c, 0.0)
        x, fixmask, fixval, zeta, m = _normalise_pair(x, x, bj, 0, m)
        self.ops.append(PairOpSpec(x, fixmask, fixval, zeta, matrix=m, strings=[(x, bi), (x, bj)]))

    def fermionic_single_excitation(self, phi, wire_i, wire_j):
        """exp((phi/2)(a†_i a_j - a†_j a_i)): the Givens rotation between two *modes* under Jordan-Wigner.
        Same 2x2 block as ``single_excitation`` on |q_i q_j> = |01>, |10>, times the parity of the
        occupied modes strictly between i and j (for adjacent modes the two gates coincide)."""
        bi, bj = self._bit(wire_i), self._bit(wire_j)
        c, s = math.cos(0.5 * phi), math.sin(0.5 * phi)
        x = bi | bj
        hi, lo = max(bi, bj), min(bi, bj)
        between = (hi - 1) & ~((lo << 1) - 1)
        m = (c, 0.0, -s, 0.0, s, 0.0, c, 0.0)
        x, fixmask, fixval, zeta, m = _normalise_pair(x, x, bj, between, m)
        self.ops.append(PairOpSpec(x, fixmask, fixval, zeta, matrix=m,
                                   strings=[(x, between | bi), (x, between | bj)]))

    def gaussian(self, unitary, wires):
        """Fermionic Gaussian unitary W with W a†_p W† = sum_q U[p,q] a†_q on the listed modes: the
        Givens network of ``givens_decomposition_square(U)`` with fermionic rotations, i.e. the recipe
        of reference adapt_vqe.py:344-354 for an arbitrary (not necessarily adjacent) set of modes.
        Returns the phase the network puts on the vacuum."""
        from .symbolic import givens_decomposition_square
        decomposition, diagonal = givens_decomposition_square(np.asarray(unitary, dtype=complex))
        total = 0.0
        phases = []
        for i, w in enumerate(wires):
            a = float(np.angle(diagonal[i]))
            if a != 0.0:
                phases.append((self._bit(w), 0.5 * a))
                total += a
        if phases:
            zs = [b for b, _ in phases]
            self.ops.append(DiagOpSpec(zs, [a for _, a in phases], -1, [(0, b) for b in zs]))
        for layer in reversed(decomposition):
            phases = []
            for (i, j, theta, phi) in layer:
                self.fermionic_single_excitation(2.0 * theta, wires[i], wires[j])
                if phi != 0.0:
                    phases.append((self._bit(wires[j]), 0.5 * float(phi)))
                    total += float(phi)
            if phases:
                zs = [b for b, _ in phases]
                self.ops.append(DiagOpSpec(zs, [a for _, a in phases], -1, [(0, b) for b in zs]))
        return -0.5 * total            # vacuum picks up exp(-i/2 sum of RZ angles)

    def basis_change_separable(self, x_dimension, y_dimension):
        """Same unitary as ``basis_change`` for ``fourier_transform_matrix`` (up to a global phase), compiled
        from its tensor structure: the matrix is F_x (x) F_y per spin, so W factorises into commuting row
        and column transforms -- N(Nx+Ny-2)... Givens rotations instead of n(n-1)/2 (3x3: 36 vs 144).
        Returns the vacuum phase of this network (see ``basis_change_vacuum_phase``)."""
        Nx, Ny = x_dimension, y_dimension

        def dft(L):
            k = np.arange(L)
            return np.exp(-1j * 2 * np.pi * np.outer(k, k) / L) / np.sqrt(L)

        phase = 0.0
        for spin in (0, 1):
            if Nx > 1:
                for y in range(Ny):
                    phase += self.gaussian(dft(Nx), [2 * (x + y * Nx) + spin for x in range(Nx)])
            if Ny > 1:
                for x in range(Nx):
                    phase += self.gaussian(dft(Ny), [2 * (x + y * Nx) + spin for y in range(Ny)])
        return phase

    @staticmethod
    def basis_change_vacuum_phase(diagonal, circuit_description):
        """Phase (radians) the reference W network (``basis_change``) puts on the vacuum."""
        total = sum(float(np.angle(d)) for d in diagonal)
        for layer in circuit_description:
            for (_, _, _, phi) in layer:
                total += float(phi)
        return -0.5 * total

    def pauli_rotation(self, x, z, coef, angle=0.0, param=-1):
        """exp(-i a coef P) with a = theta[param] (or the fixed ``angle``) for one string (x, z)."""
        if x == 0:
            if z == 0:
                return
            if param >= 0:
                self.ops.append(DiagOpSpec([z], [coef], param, [(0, z)]))
            else:
                self.ops.append(DiagOpSpec([z], [coef * angle], -1, [(0, z)]))
            return
        k = popcount(x & z)
        b = (1, 1j, -1, -1j)[k & 3] * (-1) ** k
        top = 1 << (x.bit_length() - 1)
        sgn = 1.0 if coef >= 0 else -1.0
        if param >= 0:
            self.ops.append(PairOpSpec(x, top, 0, z, kind=1, param=param, scale=abs(coef), bhat=b * sgn, strings=[(x, z)]))
        else:
            a = coef * angle
            c, s = math.cos(a), math.sin(a)
            m01, m10 = -1j * s * b, -1j * s * np.conj(b)
            m = (c, 0.0, m01.real, m01.imag, m10.real, m10.imag, c, 0.0)
            self.ops.append(PairOpSpec(x, top, 0, z, matrix=m, strings=[(x, z)]))

    # -- generators -------------------------------------------------------------------------
    def generator(self, plan: GeneratorPlan, param=-1, angle=0.0):
        """exp(-i theta G): theta = theta[param] or the fixed ``angle`` (Trotterize_generator)."""
        for piece in plan.pieces:
            if isinstance(piece, DiagPiece):
                if param >= 0:
                    self.ops.append(DiagOpSpec(list(piece.z), list(piece.coef), param, piece.strings))
                else:
                    self.ops.append(DiagOpSpec(list(piece.z), [c * angle for c in piece.coef], -1, piece.strings))
                continue
            r = abs(piece.b)
            bhat = piece.b / r
            if param >= 0:
                self.ops.append(PairOpSpec(piece.x, piece.fixmask, piece.fixval, piece.zeta, kind=1, param=param,
                                           scale=r, bhat=bhat, strings=piece.strings))
            else:
                a = r * angle
                c, s = math.cos(a), math.sin(a)
                m01, m10 = -1j * s * bhat, -1j * s * np.conj(bhat)
                m = (c, 0.0, m01.real, m01.imag, m10.real, m10.imag, c, 0.0)
                self.ops.append(PairOpSpec(piece.x, piece.fixmask, piece.fixval, piece.zeta, matrix=m,
                                           strings=piece.strings))

    def basis_change(self, diagonal, circuit_description):
        """The W network of reference adapt_vqe.py:344-354: RZ(angle(diagonal[q])) on every wire, then for
        each layer (already reversed by the caller) SingleExcitation(2 theta, [i, j]) and RZ(phi, j)."""
        first = [(self._bit(q), 0.5 * float(np.angle(diagonal[q]))) for q in range(len(diagonal))
                 if float(np.angle(diagonal[q])) != 0.0]
        if first:
            zs = [b for b, _ in first]
            self.ops.append(DiagOpSpec(zs, [a for _, a in first], -1, [(0, b) for b in zs]))
        for layer in circuit_description:
            phases = []
            for op in layer:
                i, j, theta, phi = op
                self.single_excitation(2.0 * theta, i, j)
                if phi != 0.0:
                    phases.append((self._bit(j), 0.5 * float(phi)))
            if phases:
                # the RZs of one layer sit on disjoint wires and commute with the layer's other rotations:
                # one diagonal op per layer
                zs = [b for b, _ in phases]
                self.ops.append(DiagOpSpec(zs, [a for _, a in phases], -1, [(0, b) for b in zs]))

    # -- compilation ------------------------------------------------------------------------
    def compile(self, ctx, fuse=True, tile_bits=None, low_bits=None, absorb=True):
        return DeviceProgram(ctx, self, fuse=fuse, tile_bits=tile_bits, low_bits=low_bits, absorb=absorb)


# ---------------------------------------------------------------------------------------------
# tile scheduling
# ---------------------------------------------------------------------------------------------
MAX_TILE_OPS = 96             # FH_TILE_MAX_SUB
MAX_TILE_TERMS = 256          # FH_TILE_MAX_TERMS
MAX_TILE_PATTERN_BITS = 4     # pattern bits a pair op may have to be fused into a tile
_LAUNCH_BYTES = 12e6          # bytes of traffic one kernel launch is "worth" (launch latency x bandwidth)


def _op_bytes(op, n):
    if isinstance(op, DiagOpSpec):
        return 32.0 * 2 ** n
    return 64.0 * 2 ** (n - popcount(op.fixmask))


def absorb_phases(ops):
    """Push fixed single-qubit Z phases (RZ-type diagonal terms) through the pair ops that follow them.

    With D = exp(-i sum_w a_w Z_w) pending after the ops seen so far, a pair op G whose pattern fixes
    every bit of its x-mask obeys G D = D (D^dagger G D), and D^dagger G D is the same op with its
    off-diagonal entries multiplied by a unit phase.  All absorbed phases end up in ONE diagonal op at
    the end of the segment instead of one diagonal pass per layer.  Exact (no approximation); ops that
    cannot absorb (pattern does not pin the x bits) first flush the pending phase.
    """
    import cmath
    pending = {}                                     # bit mask (single bit) -> angle a_w
    out = []

    def flush():
        if pending:
            zs = sorted(pending)
            out.append(DiagOpSpec(zs, [pending[b] for b in zs], -1, [(0, b) for b in zs]))
            pending.clear()

    for op in ops:
        if isinstance(op, DiagOpSpec):
            if op.param < 0 and all(popcount(z) == 1 for z in op.z):
                for z, a in zip(op.z, op.coef):
                    pending[z] = pending.get(z, 0.0) + a
            else:
                out.append(op)                       # diagonal ops commute with the pending phase
            continue
        touched = [b for b in pending if b & op.x]
        if not touched:
            out.append(op)
            continue
        if op.fixmask & op.x != op.x or popcount(op.x) > 6:
            flush()
            out.append(op)
            continue
        # rho = d(j)/d(i) for the pattern side i:  exp(+2i sum_{w in x} a_w (-1)^{i_w})
        expo = sum(pending[b] * (-1.0 if op.fixval & b else 1.0) for b in touched)
        rho = cmath.exp(2j * expo)
        # conjugation by Z phases on the x bits turns each Pauli string (x, z) of the op into a mix of
        # (x, z ^ s), s subset of x: keep all of them so the scheduler's commutation test stays valid
        xbits = [1 << b for b in range(op.x.bit_length()) if op.x >> b & 1]
        subsets = [0]
        for b in xbits:
            subsets += [s | b for s in subsets]
        strings = sorted({(sx, sz ^ s) for (sx, sz) in op.strings for s in subsets})
        new = PairOpSpec(op.x, op.fixmask, op.fixval, op.zeta, op.kind, op.param, op.scale, op.bhat, op.matrix,
                         strings)
        if op.kind == 1:
            new.bhat = complex(op.bhat) * rho
        else:
            m = [complex(op.matrix[2 * i], op.matrix[2 * i + 1]) for i in range(4)]
            m[1] *= rho
            m[2] *= rho.conjugate()
            new.matrix = tuple(v for c in m for v in (c.real, c.imag))
        out.append(new)
    flush()
    return out


def _noncommuting_bitsets(ops):
    """For every op the set (Python int, bit j) of ops j it does NOT commute with, by the string-level rule of
    ``_ops_commute``: some string of one anticommutes with some string of the other.  One vectorised sweep per op
    over all strings of the circuit instead of a Python double loop per op pair (the scheduler asks O(ops x tiles)
    such questions; 3x3 with 204 generators: 0.14 s -> ~0.02 s per compile)."""
    m = len(ops)
    starts, xs, zs = [], [], []
    for op in ops:
        starts.append(len(xs))
        for (x, z) in op.strings:
            xs.append(x)
            zs.append(z)
    if not xs:
        return [0] * m
    X, Z = np.array(xs, dtype=np.uint64), np.array(zs, dtype=np.uint64)
    starts_a = np.array(starts, dtype=np.intp)
    counts = np.diff(np.append(starts_a, len(xs)))
    nonempty = counts > 0
    # reduceat needs strictly valid segment starts: run it over the ops that own strings only
    seg = starts_a[nonempty]
    out = []
    for i, op in enumerate(ops):
        if not op.strings:
            out.append(0)
            continue
        anti = np.zeros(len(xs), dtype=bool)
        for (x, z) in op.strings:
            anti |= ((np.bitwise_count(X & np.uint64(z)) + np.bitwise_count(Z & np.uint64(x))) & 1).astype(bool)
        per_op = np.zeros(m, dtype=bool)
        per_op[nonempty] = np.logical_or.reduceat(anti, seg)
        out.append(int.from_bytes(np.packbits(per_op, bitorder="little").tobytes(), "little"))
    return out


def schedule(ops, n, tile_bits, low_bits, lookahead=512):
    """Greedy in-order packing of commuting-compatible ops into tiles.

    Returns a list of launch items: ('tile', [bit positions], [ops]) or ('op', op).  An op may jump
    ahead of skipped ops only if it commutes with every one of them (string-level check).
    """
    ops = list(ops)
    noncommuting = _noncommuting_bitsets(ops)
    items = []
    remaining = list(range(len(ops)))
    base_bits = (1 << low_bits) - 1
    while remaining:
        bits = base_bits
        chosen, skipped = [], []
        skipped_set = 0
        scanned = n_terms = 0
        for i in remaining:
            op = ops[i]
            scanned += 1
            if noncommuting[i] & skipped_set:
                skipped.append(i)
                skipped_set |= 1 << i
            else:
                need = bits | op.tile_bits
                nt = len(op.z) if isinstance(op, DiagOpSpec) else 0
                fusable = isinstance(op, DiagOpSpec) or popcount(op.fixmask) <= MAX_TILE_PATTERN_BITS
                if (fusable and popcount(need) <= tile_bits and len(chosen) < MAX_TILE_OPS
                        and n_terms + nt <= MAX_TILE_TERMS):
                    bits = need
                    n_terms += nt
                    chosen.append(i)
                else:
                    skipped.append(i)
                    skipped_set |= 1 << i
            if len(skipped) >= lookahead:
                break
        rest = skipped + remaining[scanned:]
        if not chosen:                                   # cannot happen (first op always fits) but stay safe
            items.append(("op", ops[remaining[0]]))
            remaining = remaining[1:]
            continue
        chosen_ops = [ops[i] for i in chosen]
        alone = sum(_op_bytes(o, n) + _LAUNCH_BYTES for o in chosen_ops)
        fused = 32.0 * 2 ** n + _LAUNCH_BYTES
        if len(chosen) == 1 or fused >= alone:
            # keep program order: only the leading run of chosen ops is emitted unfused
            items.extend(("op", o) for o in chosen_ops)
        else:
            # pad the tile with the lowest free bits so global accesses stay wide
            b = 0
            while popcount(bits) < min(tile_bits, n):
                if not bits >> b & 1:
                    bits |= 1 << b
                b += 1
            items.append(("tile", [p for p in range(n) if bits >> p & 1], chosen_ops))
        remaining = rest
    return items


class DeviceProgram:
    """A finalized ``fh_program`` plus the bookkeeping to call ``fh_program_evaluate``."""

    def __init__(self, ctx, circuit: Circuit, fuse=True, tile_bits=None, low_bits=None, absorb=True):
        self.absorb = absorb
        self.ctx = ctx
        self.n = circuit.n
        self.n_params = circuit.n_params
        self.markers = {}
        self._h = _cabi._vp()
        L = _cabi.lib()
        _cabi.check(L.fh_program_create(ctx._h, self.n, self.n_params, _cabi.C.byref(self._h)))
        n = self.n
        if tile_bits is None:
            tile_bits = max(1, min(12, n - 7)) if n > 8 else n
            if os.environ.get("FHSIM_TILE_BITS"):            # tuning knob (tools/, profiles/)
                tile_bits = int(os.environ["FHSIM_TILE_BITS"])
        tile_bits = min(tile_bits, n, 13)
        if low_bits is None:
            low_bits = 1 if n <= 20 else 2
        low_bits = min(low_bits, tile_bits)
        self.tile_bits, self.low_bits = tile_bits, low_bits
        self.n_items = 0
        self.n_tiles = 0
        segment = []
        for op in circuit.ops + [Marker("__end__")]:
            if isinstance(op, Marker):
                self._emit_segment(segment, fuse)
                segment = []
                self.markers[op.name] = self.n_items
            else:
                segment.append(op)
        _cabi.check(L.fh_program_finalize(self._h))

    def _emit_segment(self, ops, fuse):
        if not ops:
            return
        L = _cabi.lib()
        if self.absorb:
            ops = absorb_phases(ops)
        if fuse:
            items = schedule(ops, self.n, self.tile_bits, self.low_bits)
        else:
            items = [("op", o) for o in ops]
        for item in items:
            if item[0] == "op":
                self._add_op(item[1])
            else:
                _, bits, chosen = item
                arr, ptr = _cabi.i32_array(bits)
                _cabi.check(L.fh_program_begin_tile(self._h, len(bits), ptr))
                for o in chosen:
                    self._add_op(o)
                _cabi.check(L.fh_program_end_tile(self._h))
                self.n_tiles += 1
            self.n_items += 1

    def _add_op(self, op):
        L = _cabi.lib()
        if isinstance(op, DiagOpSpec):
            za, zp = _cabi.u64_array(op.z)
            ca, cp = _cabi.f64_array(op.coef)
            _cabi.check(L.fh_program_add_diag(self._h, len(op.z), zp, cp, op.param))
        else:
            ma, mp = _cabi.f64_array(op.matrix)
            b = complex(op.bhat)
            _cabi.check(L.fh_program_add_pair(self._h, op.x, op.fixmask, op.fixval, op.zeta, op.kind, op.param,
                                              float(op.scale), b.real, b.imag, mp))

    def close(self):
        if self._h:
            _cabi.lib().fh_program_destroy(self._h)
            self._h = _cabi._vp()

    def __del__(self):
        try:
            if _cabi.alive():
                self.close()
        except Exception:
            pass

    def last_stats(self):
        """(device milliseconds, kernel launches) of the most recent ``evaluate`` call."""
        ms, nl = _cabi.C.c_double(), _cabi.C.c_int()
        _cabi.check(_cabi.lib().fh_program_last_stats(self._h, _cabi.C.byref(ms), _cabi.C.byref(nl)))
        return ms.value, nl.value

    def sector_info(self):
        """Path of the most recent ``evaluate``: dict(active, pool_in_sector, cluster, dim, ops, transposes, remote_ops).
        ``active``: the call ran on the sector-compressed state resident in one thread-block cluster (csrc/sector_eval.cu);
        ``pool_in_sector`` / ``k2_in_sector``: full-space circuit kernels, but K3 screened the pool / K2 applied the first
        observable on sector-compressed copies of the state (K2 with lambda output only from 20 qubits on); ``dense_tail``:
        the trailing fixed single-species network (W), H, W^dagger and K3 all ran on compressed vectors (two dense blocks per W)."""
        C = _cabi.C
        act, cl, nops, ntr, nrem = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        dim = C.c_uint64()
        _cabi.check(_cabi.lib().fh_program_sector_info(self._h, C.byref(act), C.byref(cl), C.byref(dim), C.byref(nops),
                                                       C.byref(ntr), C.byref(nrem)))
        return dict(active=act.value == 1, pool_in_sector=bool(act.value & 2) and act.value != 1,
                    k2_in_sector=bool(act.value & 4) and act.value != 1,
                    dense_tail=bool(act.value & 8) and act.value != 1,
                    cluster_prefix=bool(act.value & 16) and act.value != 1, cluster=cl.value, dim=int(dim.value), ops=nops.value, transposes=ntr.value,
                    remote_ops=nrem.value)

    def payload_bytes(self):
        """(host->device, device->host) bytes of one ``evaluate`` call."""
        a, b = _cabi.C.c_size_t(), _cabi.C.c_size_t()
        _cabi.check(_cabi.lib().fh_program_payload_bytes(self._h, _cabi.C.byref(a), _cabi.C.byref(b)))
        return a.value, b.value

    def time_items(self, state, first=0, count=None, dagger=False, reps=20):
        """Average device milliseconds of items [first, first+count) launched back to back (measurement)."""
        if count is None:
            count = self.n_items - first
        ms = _cabi.C.c_double()
        _cabi.check(_cabi.lib().fh_program_time_items(self._h, state._h, first, count, int(dagger), reps,
                                                      _cabi.C.byref(ms)))
        return ms.value

    def run(self, state, thetas=(), first=0, count=None, dagger=False):
        """Apply launch items [first, first+count) to ``state`` in place."""
        if count is None:
            count = self.n_items - first
        ta, tp = _cabi.f64_array(thetas)
        _cabi.check(_cabi.lib().fh_program_run(self._h, state._h, tp, len(ta), first, count, int(dagger)))

    def evaluate(self, basis_index, thetas, tables, grads=False, pool=None, pool_pos=0, pool_range=None,
                 targets=(), state_out=None):
        """One fused evaluation; returns dict(expvals, grads, pool, overlaps)."""
        C = _cabi.C
        nt = len(tables)
        nv = len(targets)
        if pool is not None:
            first, count = pool_range if pool_range is not None else (0, pool.n_out)
        else:
            first = count = 0
        # argument / result buffers and their ctypes pointers are built once per call shape (numpy's .ctypes and the ctypes
        # array constructors cost microseconds each, which is visible next to a 0.14 ms evaluation)
        shape = (nt, nv, bool(grads), pool is not None, count)
        buf = self.__dict__.get("_eval_buf")
        if buf is None or buf[0] != shape:
            th = np.zeros(max(self.n_params, 1))
            ex = np.zeros(nt)
            g_ = np.zeros(max(self.n_params, 1))
            ov_ = np.zeros(2 * max(nv, 1))
            po = np.zeros(max(count, 1))
            buf = (shape, th, th.ctypes.data_as(_cabi._f64p), ex, ex.ctypes.data_as(_cabi._f64p), g_, g_.ctypes.data_as(_cabi._f64p),
                   ov_, ov_.ctypes.data_as(_cabi._f64p), po, po.ctypes.data_as(_cabi._f64p))
            self._eval_buf = buf
        _, th, th_p, ex, ex_p, g_, g_p, ov_, ov_p, po, po_p = buf
        n_th = len(thetas)
        if n_th != self.n_params:
            raise ValueError(f"program expects {self.n_params} parameters, got {n_th}")
        if n_th:
            th[:n_th] = thetas
        # handle arrays: rebuilt only when the handles change (the objects are kept alive by the key's references)
        hk = self.__dict__.get("_eval_handles")
        if hk is None or len(hk[0]) != nt or len(hk[1]) != nv or any(a is not b for a, b in zip(hk[0], tables)) or \
                any(a is not b for a, b in zip(hk[1], targets)) or any(t._h.value != v for t, v in zip(tables, hk[4])):
            hk = (list(tables), list(targets), (C.c_void_p * nt)(*[t._h for t in tables]),
                  (C.c_void_p * max(nv, 1))(*[t._h for t in targets]), [t._h.value for t in tables])
            self._eval_handles = hk
        tab_arr, tgt_arr = hk[2], hk[3]
        _cabi.check(_cabi.lib().fh_program_evaluate(
            self._h, int(basis_index), th_p, n_th, nt, tab_arr, ex_p,
            g_p if grads else None,
            pool._h if pool is not None else None, int(pool_pos), int(first), int(count),
            po_p if pool is not None else None,
            nv, tgt_arr, ov_p,
            state_out._h if state_out is not None else None))
        expvals = ex.copy()
        g = g_.copy() if grads else None
        pout = po.copy() if pool is not None else None
        ov = ov_.copy()
        return {
            "expvals": expvals,
            "grads": g[:self.n_params] if grads else None,
            "pool": pout[:count] if pool is not None else None,
            "overlaps": (ov[0:2 * nv:2] + 1j * ov[1:2 * nv:2]) if nv else np.zeros(0, complex),
        }
