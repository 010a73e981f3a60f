"""Statevector sharded over the ranks of one NVLink domain by its top index bits.

BASELINE config 5 (4x4 Hubbard, 32 qubits, complex128 = 64 GiB) does not fit one GPU with the
work buffers the evaluation needs, so the state is cut into G = 2^g slabs: rank r owns the
amplitudes whose top g *physical* index bits equal r.  The reference has no counterpart (it is
single-process, single-device: ``models/adapt_vqe.py:156``; it only switches PennyLane device at
``n_qubits >= 20``, ``adapt_vqe.py:299-304``).

Design
  * A :class:`QubitLayout` maps logical index bits (wire w <-> logical bit n-1-w) to physical
    bits.  The top g physical bits are the rank; the rest index the local slab.
  * Every op whose x-mask is local runs on the slab with the ordinary single-GPU kernels
    (``lower_pair`` / ``lower_diag`` fold the rank's global bits into patterns and signs at compile
    time: a pattern bit that lives in the rank either disables the op on that rank or is dropped; a
    Z/parity bit that lives in the rank flips a sign).
  * An op whose x-mask touches a global bit is made local by a **global<->local qubit swap**:
    a local bit-permutation pass (``fh_state_swap_bits``) that brings g chosen local qubits to the
    top of the slab, followed by ONE all-to-all of the 2^g top-local chunks
    (``torch.distributed.all_to_all_single``: NCCL over NVLink/NVSwitch on the GPU box, gloo in the
    CPU tests).  Which qubits become global is decided by a Belady rule over the upcoming ops.
  * H|phi> and the pool-gradient scan are sums over x-mask groups, so they are evaluated as a few
    *passes*: each pass handles the groups that are local in the current layout, then the states
    involved are re-laid out together so that the remaining groups become local.
  * Scalars (<H>, overlaps, the whole gradient vector) are all-reduced once.

The planner and the lowering are backend independent: ``CudaEngine`` (below) runs on libfhsim
through the C-ABI; ``tests/emulate_sharded.py`` is a numpy stand-in used by the world_size-2 gloo
tests on CPU.  There is no CPU fallback in the product: ``CudaEngine`` needs the extension and a GPU.
"""
from __future__ import annotations

import os

import numpy as np

from .circuit import DiagOpSpec, Marker, PairOpSpec
from .tables import DiagPiece, PauliTable, popcount


# ---------------------------------------------------------------------------------------------
# layout
# ---------------------------------------------------------------------------------------------
class QubitLayout:
    """perm[b] = physical bit position of logical index bit b; physical bits >= n_local are the rank."""

    def __init__(self, n_qubits: int, g: int, perm=None):
        self.n, self.g, self.nl = int(n_qubits), int(g), int(n_qubits) - int(g)
        self.perm = list(range(self.n)) if perm is None else list(perm)

    def copy(self):
        return QubitLayout(self.n, self.g, self.perm)

    def phys(self, mask: int) -> int:
        out, b = 0, 0
        while mask:
            if mask & 1:
                out |= 1 << self.perm[b]
            mask >>= 1
            b += 1
        return out

    def logical_at(self, phys_bit: int) -> int:
        return self.perm.index(phys_bit)

    def global_logical_bits(self):
        return [b for b in range(self.n) if self.perm[b] >= self.nl]

    def is_local(self, mask: int) -> bool:
        return self.phys(mask) >> self.nl == 0

    def __eq__(self, other):
        return isinstance(other, QubitLayout) and self.perm == other.perm and self.g == other.g


def swap_steps(layout: QubitLayout, new_globals):
    """Steps that make the logical bits ``new_globals`` (currently local) the rank bits.

    Returns (pairs, new_layout): ``pairs`` = disjoint (a, b) local physical positions to exchange before the
    all-to-all, which then exchanges physical bit nl-g+k with bit nl+k for k < g."""
    g, nl = layout.g, layout.nl
    new_globals = list(new_globals)
    assert len(new_globals) == g and len(set(new_globals)) == g
    assert all(layout.perm[b] < nl for b in new_globals), "new global qubits must currently be local"
    lay = layout.copy()
    top = list(range(nl - g, nl))
    want = set(new_globals)
    stay = [p for p in top if lay.logical_at(p) in want]
    free_top = [p for p in top if p not in stay]
    movers = [b for b in new_globals if lay.perm[b] not in top]
    pairs = []
    for b, p in zip(movers, free_top):
        q = lay.perm[b]
        other = lay.logical_at(p)
        pairs.append((q, p))
        lay.perm[b], lay.perm[other] = p, q
    for k in range(g):
        lo, hi = lay.logical_at(nl - g + k), lay.logical_at(nl + k)
        lay.perm[lo], lay.perm[hi] = nl + k, nl - g + k
    return pairs, lay


# ---------------------------------------------------------------------------------------------
# lowering of ops / tables / pool entries to one rank's slab
# ---------------------------------------------------------------------------------------------
def _parity(v: int) -> int:
    return popcount(v) & 1


def renormalise(x, fm, fv, ze):
    """Device pair ops enumerate index pairs from the side whose TOP x bit is 0 and need that bit in the pattern.
    A qubit permutation can move another x bit to the top, so re-express ``(x, fm, fv, ze)`` as a list of
    ``(fm', fv', swapped)``: ``swapped`` means the pattern side is now the former partner side, i.e. the 2x2 block
    must be transposed (m00<->m11, m01<->m10) with the off-diagonals times (-1)^popcount(x & ze)."""
    top = 1 << (x.bit_length() - 1)
    if fm & top:
        if fv & top:
            return [(fm, (fv ^ x) & fm, True)]
        return [(fm, fv, False)]
    # top bit unconstrained: split on it.  Pairs whose pattern side has top = 1 are enumerated from their partner.
    return [(fm | top, fv, False), (fm | top, ((fv | top) ^ x) & (fm | top), True)]


def _swap_roles(kind, bhat, matrix, x, ze):
    s = -1.0 if _parity(x & ze) else 1.0
    if kind == 1:
        return complex(bhat).conjugate() * s, matrix
    m = matrix
    return bhat, (m[6], m[7], s * m[4], s * m[5], s * m[2], s * m[3], m[0], m[1])


def lower_pair(op: PairOpSpec, layout: QubitLayout, rank: int):
    """Local version(s) of a pair op on ``rank``: a list (empty when the op does not act on this rank's slab)."""
    nl = layout.nl
    lm = (1 << nl) - 1
    x, fm, fv, ze = layout.phys(op.x), layout.phys(op.fixmask), layout.phys(op.fixval), layout.phys(op.zeta)
    if x >> nl:
        raise ValueError("pair op is not local in this layout")
    if (rank & (fm >> nl)) != (fv >> nl):
        return []
    sgn = -1.0 if _parity(rank & (ze >> nl)) else 1.0
    strings = [(layout.phys(sx), layout.phys(sz) & lm) for sx, sz in op.strings]
    bhat, matrix = complex(op.bhat), tuple(op.matrix)
    if sgn < 0:
        if op.kind == 1:
            bhat = -bhat
        else:
            m = list(matrix)
            m[2], m[3], m[4], m[5] = -m[2], -m[3], -m[4], -m[5]
            matrix = tuple(m)
    out = []
    for fm2, fv2, swapped in renormalise(x, fm & lm, fv & lm, ze & lm):
        b2, m2 = _swap_roles(op.kind, bhat, matrix, x, ze & lm) if swapped else (bhat, matrix)
        out.append(PairOpSpec(x, fm2, fv2, ze & lm, op.kind, op.param, op.scale, b2, m2, strings))
    return out


def lower_diag(op: DiagOpSpec, layout: QubitLayout, rank: int):
    nl = layout.nl
    lm = (1 << nl) - 1
    zs, cs = [], []
    for z, c in zip(op.z, op.coef):
        pz = layout.phys(z)
        zs.append(pz & lm)
        cs.append(-c if _parity(rank & (pz >> nl)) else c)
    return DiagOpSpec(zs, cs, op.param, [(0, z) for z in zs])


def lower_table(table: PauliTable, layout: QubitLayout, rank: int, todo=None):
    """Terms of ``table`` whose x-mask is local in ``layout`` -> (local PauliTable or None, indices handled).
    ``todo``: iterable of term indices still to be applied (default: all)."""
    nl = layout.nl
    lm = (1 << nl) - 1
    xs, zs, cs, done = [], [], [], []
    for t in (range(len(table)) if todo is None else todo):
        x, z = layout.phys(int(table.x[t])), layout.phys(int(table.z[t]))
        if x >> nl:
            continue
        done.append(t)
        sgn = -1.0 if _parity(rank & (z >> nl)) else 1.0
        xs.append(x)
        zs.append(z & lm)
        cs.append(sgn * complex(table.coeff[t]))
    if not done:
        return None, done
    return PauliTable(nl, xs, zs, cs), done


def lower_pool_entries(entries, layout: QubitLayout, rank: int, todo):
    """entries: list of (x, fixmask, fixval, zeta, b, out).  Returns (local entry arrays, indices handled);
    handled entries whose pattern excludes this rank contribute nothing and are dropped from the arrays."""
    nl = layout.nl
    lm = (1 << nl) - 1
    local, done = [], []
    for e in todo:
        x, fm, fv, ze, b, out = entries[e]
        px = layout.phys(x)
        if px >> nl:
            continue
        done.append(e)
        pfm, pfv, pze = layout.phys(fm), layout.phys(fv), layout.phys(ze)
        if (rank & (pfm >> nl)) != (pfv >> nl):
            continue
        sgn = -1.0 if _parity(rank & (pze >> nl)) else 1.0
        for fm2, fv2, swapped in renormalise(px, pfm & lm, pfv & lm, pze & lm):
            b2 = sgn * complex(b)
            if swapped:
                b2 = b2.conjugate() * (-1.0 if _parity(px & pze & lm) else 1.0)
            local.append((px, fm2, fv2, pze & lm, b2, out))
    local.sort(key=lambda t: t[5])
    return local, done


def _is_up_bit(logical_bit: int, n: int) -> bool:
    """wire = n - 1 - bit; even wires are the up orbitals (linalg/exact_diagonalization.py:34-51)."""
    return (n - 1 - logical_bit) % 2 == 0


def conserves_species(ops, n: int) -> bool:
    """True when every pair op pins all its x bits and moves as many electrons in as out of each spin species, i.e. the
    circuit maps every (N_up, N_dn) sector to itself op by op (diagonal ops always do)."""
    up = sum(1 << b for b in range(n) if _is_up_bit(b, n))
    for op in ops:
        if isinstance(op, DiagOpSpec):
            continue
        x, fm, fv = int(op.x), int(op.fixmask), int(op.fixval)
        if x & ~fm:
            return False
        for mask in (up, ~up & ((1 << n) - 1)):
            if 2 * bin(fv & x & mask).count("1") != bin(x & mask).count("1"):
                return False
    return True


def local_sector(layout: QubitLayout, rank: int, n_up: int, n_dn: int):
    """(up_mask, dn_mask, n_up_local, n_dn_local) of the slab of ``rank``: which local physical bits are up / down orbitals and
    how many electrons of each species they hold once the occupations of the rank bits are taken out; None when this slab
    holds no amplitude of the sector."""
    n, nl = layout.n, layout.nl
    up_mask = dn_mask = 0
    for b in range(n):
        p = layout.perm[b]
        isup = _is_up_bit(b, n)
        if p < nl:
            if isup:
                up_mask |= 1 << p
            else:
                dn_mask |= 1 << p
        elif (rank >> (p - nl)) & 1:
            if isup:
                n_up -= 1
            else:
                n_dn -= 1
    if n_up < 0 or n_dn < 0 or n_up > bin(up_mask).count("1") or n_dn > bin(dn_mask).count("1"):
        return None
    return up_mask, dn_mask, n_up, n_dn


def pool_entries_of(plans):
    entries = []
    for k, plan in enumerate(plans):
        if not plan.exact:
            raise NotImplementedError("pool generators must consist of mutually commuting strings")
        for piece in plan.pieces:
            if isinstance(piece, DiagPiece):
                raise NotImplementedError("diagonal pool generators are not supported by the screening kernel")
            entries.append((piece.x, piece.fixmask, piece.fixval, piece.zeta, complex(piece.b), k))
    return entries


def dagger_ops(ops):
    """Inverse circuit: reversed order, each op inverted."""
    out = []
    for op in reversed([o for o in ops if not isinstance(o, Marker)]):
        if isinstance(op, DiagOpSpec):
            out.append(DiagOpSpec(list(op.z), [-c for c in op.coef], op.param, op.strings))
        elif op.kind == 1:
            out.append(PairOpSpec(op.x, op.fixmask, op.fixval, op.zeta, 1, op.param, -op.scale, op.bhat, op.matrix,
                                  op.strings))
        else:
            m = op.matrix
            md = (m[0], -m[1], m[4], -m[5], m[2], -m[3], m[6], -m[7])
            out.append(PairOpSpec(op.x, op.fixmask, op.fixval, op.zeta, 0, -1, 0.0, op.bhat, md, op.strings))
    return out


# ---------------------------------------------------------------------------------------------
# planner
# ---------------------------------------------------------------------------------------------
def choose_globals_belady(ops, start, layout: QubitLayout):
    """g logical bits to make global before ops[start]: those whose next use as an x-bit is farthest away."""
    n, g = layout.n, layout.g
    next_use = [None] * n
    for k in range(start, len(ops)):
        op = ops[k]
        if isinstance(op, PairOpSpec):
            x, b = op.x, 0
            while x:
                if x & 1 and next_use[b] is None:
                    next_use[b] = k
                x >>= 1
                b += 1
        if all(v is not None for v in next_use):
            break
    inf = len(ops) + 1
    cand = [b for b in range(n) if layout.perm[b] < layout.nl]       # currently local
    # farthest next use first; among equals prefer qubits already sitting at the top of the slab (no local pass)
    cand.sort(key=lambda b: (-(next_use[b] if next_use[b] is not None else inf), -layout.perm[b]))
    chosen = cand[:g]
    if start < len(ops) and any(next_use[b] == start for b in chosen) and layout.nl < layout.n:
        raise ValueError("op needs more local qubits than the slab has")
    return chosen


def best_initial_layout(ops, n_qubits: int, g: int) -> QubitLayout:
    """Layout for a state that starts as a basis state (so its layout is free): the rank qubits are the g qubits
    whose first use as an x-bit lies farthest in the future, the rest keep their relative order."""
    ops = [o for o in ops if not isinstance(o, Marker)]
    probe = QubitLayout(n_qubits, 0)
    probe.g = g
    chosen = sorted(choose_globals_belady(ops, 0, probe)) if g else []
    rest = [b for b in range(n_qubits) if b not in chosen]
    perm = [0] * n_qubits
    for p, b in enumerate(rest + chosen):
        perm[b] = p
    return QubitLayout(n_qubits, g, perm)


def plan_circuit(ops, layout: QubitLayout):
    """Split ``ops`` (PairOpSpec / DiagOpSpec, logical masks) into local segments separated by qubit swaps.
    Returns (steps, final_layout); steps: ("ops", [ops], layout) | ("swap", pairs, layout_before, layout_after)."""
    ops = [o for o in ops if not isinstance(o, Marker)]
    lay = layout.copy()
    steps, seg = [], []
    for k, op in enumerate(ops):
        if isinstance(op, PairOpSpec) and not lay.is_local(op.x):
            if seg:
                steps.append(("ops", seg, lay.copy()))
                seg = []
            new_globals = choose_globals_belady(ops, k, lay)
            pairs, new_lay = swap_steps(lay, new_globals)
            steps.append(("swap", pairs, lay.copy(), new_lay.copy()))
            lay = new_lay
            assert lay.is_local(op.x)
        seg.append(op)
    if seg:
        steps.append(("ops", seg, lay.copy()))
    return steps, lay


def choose_globals_cover(masks, layout: QubitLayout):
    """Next global set for a multi-pass sum over x-mask groups: the g currently-local logical bits that occur
    in the fewest remaining masks (ties: already at the top of the slab first)."""
    n, g = layout.n, layout.g
    count = [0] * n
    for m in masks:
        b = 0
        while m:
            if m & 1:
                count[b] += 1
            m >>= 1
            b += 1
    cand = [b for b in range(n) if layout.perm[b] < layout.nl]
    cand.sort(key=lambda b: (count[b], -layout.perm[b]))
    return cand[:g]


# ---------------------------------------------------------------------------------------------
# sharded state + simulator (backend independent)
# ---------------------------------------------------------------------------------------------
class ShardedState:
    def __init__(self, engine, layout: QubitLayout):
        self.engine = engine
        self.layout = layout.copy()
        self.h = engine.new_state()

    def close(self):
        if self.h is not None:
            self.engine.free_state(self.h)
            self.h = None


class ShardedSimulator:
    """The hot path on a state sharded over ``engine.world`` = 2^g ranks."""

    def __init__(self, engine, n_qubits: int):
        self.engine = engine
        self.n = int(n_qubits)
        world = engine.world
        g = world.bit_length() - 1
        if 1 << g != world:
            raise ValueError("world size must be a power of two")
        self.g = g
        self.nl = self.n - g
        if engine.n_local != self.nl:
            raise ValueError("engine slab size does not match n_qubits - log2(world)")
        if self.nl < 2 * g:
            raise ValueError("slab too small for a global<->local swap")
        self.profiling = False         # True: adapt_screening syncs after every phase and fills self.profile (seconds)
        self.profile = {}
        self._keepalive = {}           # op lists whose ids key the engine's compiled-program cache
        self._dagger_cache = {}
        self._conserve_cache = {}
        self.sector_pool_used = False
        self.swap_count = 0            # all-to-alls issued (per state)
        self.pass_count = {"table": 0, "pool": 0}

    # -- state management ------------------------------------------------------------------------
    def new_state(self, layout=None) -> ShardedState:
        return ShardedState(self.engine, layout if layout is not None else QubitLayout(self.n, self.g))

    def set_basis(self, st: ShardedState, basis_index: int):
        p = st.layout.phys(int(basis_index))
        owner, local = p >> self.nl, p & ((1 << self.nl) - 1)
        self.engine.set_basis(st.h, local if owner == self.engine.rank else None)

    def copy(self, dst: ShardedState, src: ShardedState):
        self.engine.copy(dst.h, src.h)
        dst.layout = src.layout.copy()

    def _swap_exchange(self, h, pairs):
        """Local bit transpositions ``pairs`` then the all-to-all of the top-local chunks.  Engines that own a
        communicator (CudaEngine with fh_comm) run both as ONE pipelined exchange: the permutation of chunk k+1
        overlaps chunk k on the wire."""
        fused = getattr(self.engine, "swap_exchange", None)
        if fused is not None:
            fused(h, pairs)
            return
        if pairs:
            self.engine.swap_bits(h, pairs)
        self.engine.all_to_all(h)

    def relayout(self, states, new_globals):
        """Make ``new_globals`` the rank bits of every state in ``states`` (all must share one layout)."""
        lay = states[0].layout
        assert all(s.layout == lay for s in states)
        pairs, new_lay = swap_steps(lay, new_globals)
        for s in states:
            self._swap_exchange(s.h, pairs)
            s.layout = new_lay.copy()
            self.swap_count += 1

    def to_layout(self, st: ShardedState, target: QubitLayout):
        """Bring ``st`` to exactly ``target``: at most two all-to-alls plus local bit-permutation passes."""
        if st.layout == target:
            return
        nl, g = self.nl, self.g
        tg = [b for b in range(self.n) if target.perm[b] >= nl]
        if any(st.layout.perm[b] != target.perm[b] for b in tg):
            if any(st.layout.perm[b] >= nl for b in tg):
                # some target rank qubits are rank qubits already (wrong slot or wrong company): park a disjoint set
                spare = [b for b in range(self.n) if st.layout.perm[b] < nl and b not in tg][:g]
                self.relayout([st], spare)
            # line the target rank qubits up at the top of the slab in rank-bit order, then exchange
            self._local_permute(st, {b: target.perm[b] - g for b in tg})
            self.engine.all_to_all(st.h)
            lay = st.layout
            for k in range(g):
                lo, hi = lay.logical_at(nl - g + k), lay.logical_at(nl + k)
                lay.perm[lo], lay.perm[hi] = nl + k, nl - g + k
            self.swap_count += 1
        self._local_permute(st, {b: target.perm[b] for b in range(self.n) if target.perm[b] < nl})
        assert st.layout == target

    def _local_permute(self, st, want):
        """Move logical bits to the given local physical positions by successive disjoint transposition batches."""
        lay = st.layout
        while True:
            pairs, used = [], set()
            for b, p in want.items():
                q = lay.perm[b]
                if q == p or q in used or p in used:
                    continue
                pairs.append((q, p))
                used.update((q, p))
                if len(pairs) == 8:
                    break
            if not pairs:
                break
            self.engine.swap_bits(st.h, pairs)
            for q, p in pairs:
                a, c = lay.logical_at(q), lay.logical_at(p)
                lay.perm[a], lay.perm[c] = p, q

    # -- circuits --------------------------------------------------------------------------------
    def apply_ops(self, st: ShardedState, ops, thetas=(), n_params=0):
        """Apply PairOpSpec / DiagOpSpec ``ops`` (logical masks, circuit order) to ``st`` in place."""
        steps, _ = plan_circuit(ops, st.layout)
        rank = self.engine.rank
        for step in steps:
            if step[0] == "swap":
                _, pairs, _, new_lay = step
                self._swap_exchange(st.h, pairs)
                st.layout = new_lay.copy()
                self.swap_count += 1
                continue
            _, seg, lay = step
            local = []
            for op in seg:
                if isinstance(op, DiagOpSpec):
                    local.append(lower_diag(op, lay, rank))
                else:
                    local.extend(lower_pair(op, lay, rank))
            # every rank runs its (possibly shorter) local segment; no collective inside
            if local:
                key = (tuple(id(o) for o in seg), tuple(lay.perm), int(n_params))
                self._keepalive[key] = seg
                self.engine.run_ops(st.h, local, thetas, n_params, key=key)

    # -- K2 --------------------------------------------------------------------------------------
    def apply_table(self, table: PauliTable, phi: ShardedState, out: ShardedState | None = None) -> complex:
        """<phi|H|phi> (all-reduced) and, if ``out`` is given, out = H phi.  phi and out may come back in a
        different layout (shared by both)."""
        rank = self.engine.rank
        todo = list(range(len(table)))
        if out is not None:
            out.layout = phi.layout.copy()
        total = 0j
        first = True
        for _ in range(4 * self.n):
            local_tab, done = lower_table(table, phi.layout, rank, todo)
            if done:
                self.pass_count["table"] += 1
                total += self.engine.apply_table(local_tab, phi.h, out.h if out is not None else None,
                                                 accumulate=not first)
                first = False
                done_set = set(done)
                todo = [t for t in todo if t not in done_set]
            if not todo:
                break
            new_globals = choose_globals_cover([int(table.x[t]) for t in todo], phi.layout)
            together = [phi] + ([out] if out is not None and not first else [])
            self.relayout(together, new_globals)
            if out is not None and first:
                out.layout = phi.layout.copy()
        else:
            raise RuntimeError("sharded apply_table cannot make progress (x-mask wider than the slab?)")
        if out is not None and first:
            self.engine.set_basis(out.h, None)              # H has no terms: out = 0
        red = self.engine.all_reduce(np.array([total.real, total.imag]))
        return complex(red[0], red[1])

    # -- K3 --------------------------------------------------------------------------------------
    def pool_gradients(self, plans, psi: ShardedState, lam: ShardedState, sector=None) -> np.ndarray:
        """g_k = 2 Im <lam|G_k|psi> for every generator plan; psi and lam must share a layout (they are re-laid
        out together between passes).  sector = (N_up, N_dn): psi and lam are confined to that sector, so every slab is
        screened on its sector-compressed copy (``fh_pool_gradients_sector_masks``) instead of in its full 2^n_local space."""
        if not (psi.layout == lam.layout):
            self.to_layout(lam, psi.layout)
        entries = pool_entries_of(plans)
        n_out = len(plans)
        rank = self.engine.rank
        todo = list(range(len(entries)))
        acc = np.zeros(n_out)
        for _ in range(4 * self.n):
            local, done = lower_pool_entries(entries, psi.layout, rank, todo)
            if done:
                self.pass_count["pool"] += 1
                loc = local_sector(psi.layout, rank, *sector) if sector is not None else False
                if loc is None:
                    pass                                   # this slab holds no amplitude of the sector
                elif loc:
                    acc += self.engine.pool_partial(local, psi.h, lam.h, n_out, sector=loc)
                else:
                    acc += self.engine.pool_partial(local, psi.h, lam.h, n_out)
                done_set = set(done)
                todo = [e for e in todo if e not in done_set]
            if not todo:
                break
            new_globals = choose_globals_cover([entries[e][0] for e in todo], psi.layout)
            self.relayout([psi, lam], new_globals)
        else:
            raise RuntimeError("sharded pool scan cannot make progress")
        return self.engine.all_reduce(acc)

    # -- scalars ---------------------------------------------------------------------------------
    def inner(self, a: ShardedState, b: ShardedState) -> complex:
        if not (a.layout == b.layout):
            self.to_layout(b, a.layout)
        v = self.engine.inner(a.h, b.h)
        red = self.engine.all_reduce(np.array([v.real, v.imag]))
        return complex(red[0], red[1])

    def gather(self, st: ShardedState) -> np.ndarray:
        """Full 2^n vector in LOGICAL order on every rank (tests / small n only)."""
        slabs = self.engine.all_gather_slab(st.h)          # [world, 2^nl]
        full_phys = np.concatenate(slabs)
        n = self.n
        idx = np.arange(1 << n, dtype=np.uint64)
        phys = np.zeros(1 << n, dtype=np.uint64)
        for b in range(n):
            phys |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(st.layout.perm[b])
        return full_phys[phys]

    # -- the cfg-5 evaluation ---------------------------------------------------------------------
    def adapt_screening(self, basis_index, ansatz_ops, basis_change_ops, h_table: PauliTable, plans, thetas=(),
                        n_params=0, want_gradients=True, sector_pool=True):
        """psi_k = ansatz|basis>, phi = W psi_k, E = <phi|H|phi>, lam = W^dagger H phi, g = pool gradients
        (reference ADAPT.select_operator, models/adapt_vqe.py:297-323, on a sharded state)."""
        import time
        prof = self.profile = {}

        def lap(name, t0):
            if self.profiling:
                self.engine.sync()
                prof[name] = prof.get(name, 0.0) + time.perf_counter() - t0
            return time.perf_counter()

        t = time.perf_counter()
        psi = self.new_state(best_initial_layout(list(ansatz_ops) + list(basis_change_ops), self.n, self.g))
        self.set_basis(psi, basis_index)
        self.apply_ops(psi, ansatz_ops, thetas, n_params)
        t = lap("ansatz", t)
        phi = self.new_state()
        self.copy(phi, psi)
        self.apply_ops(phi, basis_change_ops, thetas, n_params)
        t = lap("W", t)
        lam = self.new_state(phi.layout) if want_gradients else None
        energy = self.apply_table(h_table, phi, lam)
        t = lap("H_apply", t)
        grads = None
        if want_gradients:
            phi.close()
            wd = self._dagger_cache.get(id(basis_change_ops))
            if wd is None:
                wd = self._dagger_cache[id(basis_change_ops)] = (dagger_ops(basis_change_ops), basis_change_ops)
            self.apply_ops(lam, wd[0], thetas, n_params)
            t = lap("W_dagger", t)
            self.to_layout(lam, psi.layout)
            t = lap("align_layouts", t)
            # every op conserves N_up and N_dn (and the Hubbard Hamiltonian does): psi_s and lambda_s live in the sector of
            # the basis state, so the pool is screened on sector-compressed slabs
            sector = None
            if sector_pool and os.environ.get("FHSIM_NO_SECTOR_POOL") is None:
                key = (id(ansatz_ops), id(basis_change_ops))
                ok = self._conserve_cache.get(key)
                if ok is None:
                    ok = self._conserve_cache[key] = (conserves_species(list(ansatz_ops) + list(basis_change_ops), self.n)
                                                      and h_table.conserves_species())
                if ok:
                    up = sum(1 << b for b in range(self.n) if _is_up_bit(b, self.n))
                    sector = (bin(int(basis_index) & up).count("1"), bin(int(basis_index) & ~up).count("1"))
            self.sector_pool_used = sector is not None
            grads = self.pool_gradients(plans, psi, lam, sector)
            t = lap("pool_scan", t)
            lam.close()
        else:
            phi.close()
        psi.close()
        return energy, grads


# ---------------------------------------------------------------------------------------------
# CUDA engine: libfhsim through the C-ABI, torch tensors as slab memory, torch.distributed for the collectives
# ---------------------------------------------------------------------------------------------
def pairwise_all_to_all(dist, chunks_in):
    """all-to-all of equal CPU chunks by ordered pairwise send/recv (gloo has no alltoall)."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    outs = [torch.empty_like(c) for c in chunks_in]
    for peer in range(world):
        if peer == rank:
            outs[peer].copy_(chunks_in[peer])
        elif peer > rank:
            dist.send(chunks_in[peer].contiguous(), peer)
            dist.recv(outs[peer], peer)
        else:
            dist.recv(outs[peer], peer)
            dist.send(chunks_in[peer].contiguous(), peer)
    return outs


class CudaEngine:
    """One rank's slab engine.  ``dist`` is an initialised torch.distributed (NCCL) or None for world = 1."""

    def __init__(self, n_local: int, device: int = 0, dist=None):
        import torch
        from . import _cabi
        from .backend import Context
        self.torch, self._cabi = torch, _cabi
        self.dist = dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.n_local = int(n_local)
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        # everything (our kernels, torch copies and NCCL) is ordered on ONE explicit stream.  torch's default
        # current stream is the legacy NULL stream, which fh_ctx_create would replace by a private one.
        self.stream = torch.cuda.Stream(self.device)
        torch.cuda.set_stream(self.stream)
        self.ctx = Context(device, stream=_cabi.C.c_void_p(self.stream.cuda_stream))
        self._spare = None
        self._programs = {}             # compiled local segments, keyed by the planner
        self._tables = {}               # uploaded per-layout observable tables
        self._pools = {}                # uploaded per-layout screening pools
        self.a2a_ms = 0.0
        self.time_exchanges = True      # False: exchanges are enqueued without a host synchronisation (a2a_ms not updated)
        self._host_staged = dist is not None and dist.get_backend() == "gloo"
        # communicator behind the C-ABI (fh_comm_*: NCCL resolved inside libfhsim, exchange pipelined against the local
        # bit permutation on a side stream).  torch.distributed only distributes the 128-byte id.  FHSIM_COMM=torch
        # keeps the round-1 path (all_to_all_single).
        self._comm = None
        self._spare2 = None
        import os
        if dist is not None and self.world > 1 and not self._host_staged and os.environ.get("FHSIM_COMM", "fhsim") != "torch":
            C = _cabi.C
            buf = (C.c_ubyte * 128)()
            if self.rank == 0:
                _cabi.check(_cabi.lib().fh_comm_unique_id(buf))
            t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=self.device)
            dist.broadcast(t, 0)
            idb = (C.c_ubyte * 128)(*t.cpu().tolist())
            h = _cabi._vp()
            _cabi.check(_cabi.lib().fh_comm_init(self.ctx._h, idb, self.rank, self.world, C.byref(h)))
            self._comm = h

    # slabs are torch tensors wrapped (borrowed) as fh_state handles
    class _Slab:
        __slots__ = ("t", "st")

    def _alloc(self):
        from .backend import State
        C = self._cabi.C
        s = CudaEngine._Slab()
        s.t = self.torch.empty(1 << self.n_local, dtype=self.torch.complex128, device=self.device)
        st = State.__new__(State)
        st.ctx, st.n = self.ctx, self.n_local
        st._h = self._cabi._vp()
        self._cabi.check(self._cabi.lib().fh_state_wrap(self.ctx._h, self.n_local, C.c_void_p(s.t.data_ptr()),
                                                         C.byref(st._h)))
        s.st = st
        return s

    def new_state(self):
        return [self._alloc()]          # one-element list so buffers can be exchanged in place

    def free_state(self, h):
        h[0].st.close()
        h[0] = None

    def _spare_slab(self):
        if self._spare is None:
            self._spare = self._alloc()
        return self._spare

    def set_basis(self, h, local_index):
        if local_index is None:
            h[0].t.zero_()
        else:
            h[0].st.set_basis(int(local_index))

    def copy(self, dst, src):
        dst[0].st.copy_from(src[0].st)

    def run_ops(self, h, ops, thetas, n_params, key=None):
        from .circuit import Circuit
        prog = self._programs.get(key) if key is not None else None
        if prog is None:
            circ = Circuit(self.n_local, n_params)
            circ.ops = list(ops)
            prog = circ.compile(self.ctx)
            if key is not None:
                self._programs[key] = prog
        try:
            prog.run(h[0].st, thetas)
        finally:
            if key is None:
                prog.close()

    def swap_bits(self, h, pairs):
        spare = self._spare_slab()
        aa, ap = self._cabi.i32_array([p[0] for p in pairs])
        ba, bp = self._cabi.i32_array([p[1] for p in pairs])
        self._cabi.check(self._cabi.lib().fh_state_swap_bits(spare.st._h, h[0].st._h, len(pairs), ap, bp))
        self._spare, h[0] = h[0], spare

    def swap_exchange(self, h, pairs):
        """Fused local permutation + all-to-all (see ShardedSimulator._swap_exchange)."""
        if self._comm is None:
            if pairs:
                self.swap_bits(h, pairs)
            self.all_to_all(h)
            return
        L, C = self._cabi.lib(), self._cabi.C
        spare = self._spare_slab()
        aa, ap = self._cabi.i32_array([p[0] for p in pairs])
        ba, bp = self._cabi.i32_array([p[1] for p in pairs])
        third = None
        if pairs:
            if self._spare2 is None:
                free, _ = self.torch.cuda.mem_get_info(self.device)
                if free > 2.0 * (16 << self.n_local):           # a third slab fits comfortably: pipelined exchange
                    self._spare2 = self._alloc()
            third = self._spare2
        if pairs and third is None:
            # not enough memory for a third slab (32-GiB slabs at N = 2): permute first, then exchange back into h
            self._cabi.check(L.fh_state_swap_bits(spare.st._h, h[0].st._h, len(pairs), ap, bp))
            self._cabi.check(L.fh_comm_swap_exchange(self._comm, spare.st._h, None, h[0].st._h, 0, None, None))
        elif pairs:
            self._cabi.check(L.fh_comm_swap_exchange(self._comm, h[0].st._h, spare.st._h, third.st._h, len(pairs), ap, bp))
            self._spare2, h[0] = h[0], third
        else:
            self._cabi.check(L.fh_comm_swap_exchange(self._comm, h[0].st._h, None, spare.st._h, 0, None, None))
            self._spare, h[0] = h[0], spare
        if self.time_exchanges:
            ms = C.c_double()
            self._cabi.check(L.fh_comm_last_exchange_ms(self._comm, C.byref(ms)))
            self.a2a_ms += ms.value

    def all_to_all(self, h):
        if self.world == 1:
            return
        if self._comm is not None:
            self.swap_exchange(h, [])
            return
        torch = self.torch
        spare = self._spare_slab()
        src = torch.view_as_real(h[0].t).reshape(-1)
        dst = torch.view_as_real(spare.t).reshape(-1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if self._host_staged:
            # ranks sharing one GPU (tests): NCCL refuses duplicate devices, so exchange through the host with gloo
            hs = src.cpu()
            dst.copy_(torch.cat(pairwise_all_to_all(self.dist, list(hs.chunk(self.world)))))
        else:
            self.dist.all_to_all_single(dst, src)
        e1.record()
        e1.synchronize()
        self.a2a_ms += e0.elapsed_time(e1)
        self._spare, h[0] = h[0], spare

    def apply_table(self, local_tab, h_in, h_out, accumulate):
        from .backend import DeviceTable
        C = self._cabi.C
        # per-layout tables recur every evaluation: keep the uploaded ones (small) instead of rebuilding them
        key = (local_tab.x.tobytes(), local_tab.z.tobytes(), local_tab.coeff.tobytes())
        tab = self._tables.get(key)
        if tab is None:
            if len(self._tables) >= 16:
                self._tables.pop(next(iter(self._tables))).close()
            tab = self._tables[key] = DeviceTable(self.ctx, local_tab)
        re, im = C.c_double(), C.c_double()
        fn = self._cabi.lib().fh_apply_table_accumulate if (accumulate and h_out is not None) \
            else self._cabi.lib().fh_apply_table
        self._cabi.check(fn(tab._h, h_in[0].st._h, h_out[0].st._h if h_out is not None else None,
                            C.byref(re), C.byref(im)))
        return complex(re.value, im.value)

    def pool_partial(self, local_entries, h_psi, h_lam, n_out, sector=None):
        from .backend import DevicePool
        if not local_entries:
            return np.zeros(n_out)
        # the x-mask cover of a per-layout pool costs host time (greedy over ~10^3 entries): cache the uploaded pools
        key = (n_out, tuple(local_entries))
        pool = self._pools.get(key)
        if pool is None:
            if len(self._pools) >= 8:
                self._pools.pop(next(iter(self._pools))).close()
            pool = self._pools[key] = DevicePool.from_entries(self.ctx, self.n_local, local_entries, n_out)
        if sector is not None and max(bin(sector[0]).count("1"), bin(sector[1]).count("1")) <= 16:
            up_mask, dn_mask, n_up, n_dn = sector
            return pool.gradients_sector(h_psi[0].st, h_lam[0].st, n_up, n_dn, up_mask=up_mask, dn_mask=dn_mask).copy()
        return pool.gradients(h_psi[0].st, h_lam[0].st).copy()

    def inner(self, ha, hb):
        return ha[0].st.inner(hb[0].st)

    def close(self):
        """Release cached device objects, the spare slabs and the communicator."""
        for cache in (self._programs, self._tables, self._pools):
            for obj in list(cache.values()):
                try:
                    obj.close()
                except Exception:
                    pass
            cache.clear()
        for name in ("_spare", "_spare2"):
            slab = getattr(self, name)
            if slab is not None:
                slab.st.close()
                setattr(self, name, None)
        if self._comm is not None:
            self.sync()
            self._cabi.lib().fh_comm_destroy(self._comm)
            self._comm = None

    def sync(self):
        self.stream.synchronize()

    def all_reduce(self, arr):
        if self.world == 1:
            return np.asarray(arr, dtype=np.float64)
        if self._comm is not None and np.size(arr) <= 4096:
            v = np.ascontiguousarray(arr, dtype=np.float64).copy()
            self._cabi.check(self._cabi.lib().fh_comm_all_reduce_sum(self._comm, v.ctypes.data_as(self._cabi._f64p), int(v.size)))
            return v
        t = self.torch.as_tensor(np.asarray(arr, dtype=np.float64), device="cpu" if self._host_staged else self.device)
        self.dist.all_reduce(t)
        return t.cpu().numpy()

    def all_gather_slab(self, h):
        if self.world == 1:
            return [h[0].t.cpu().numpy()]
        mine = h[0].t.cpu() if self._host_staged else h[0].t
        outs = [self.torch.empty_like(mine) for _ in range(self.world)]
        flat = [self.torch.view_as_real(o).reshape(-1) for o in outs]
        self.dist.all_gather(flat, self.torch.view_as_real(mine).reshape(-1))
        return [o.cpu().numpy() for o in outs]
