"""Operator pools (reference ``operators/pool.py``).

Only ``hubbard_interaction_pool_simplified`` (reference :220-255) is used by any driver
(``models/adapt_vqe.py:142``); it is the pool the screening kernel is compiled from.  The
other generators of the reference file are kept as symbolic helpers with the same
signatures.
"""
from functools import reduce

from fhsim.symbolic import FermionOperator, hermitian_conjugated, normal_ordered


def _index_maps(Nx):
    def to_index(ix, iy, spin):
        return 2 * (ix + iy * Nx) + spin

    def to_xy(site):
        return site % Nx, site // Nx

    return to_index, to_xy


def _canonical_key(operator, sign=1):
    """Hashable identity of an operator up to the 1e-8 tolerance ``==`` uses."""
    return frozenset((term, complex(round((sign * c).real, 7), round((sign * c).imag, 7)))
                     for term, c in operator.terms.items() if abs(c) > 1e-8)


def _momentum_exchange(i1, i2, i3, i4, hermitian=False):
    if hermitian:
        return FermionOperator(f'{i1}^ {i2}^ {i3} {i4}') + FermionOperator(f'{i3}^ {i4}^ {i1} {i2}')
    return FermionOperator(f'{i1}^ {i2}^ {i3} {i4}', 1j) - FermionOperator(f'{i3}^ {i4}^ {i1} {i2}', 1j)


def hubbard_interaction_pool_simplified(Nx, Ny):
    """i(c†_{k1+q,s} c†_{k2-q,-s} c_{k2,-s} c_{k1,s} - h.c.), q != 0, normal ordered, unique up to sign.

    Loop order spin -> k1 -> k2 -> q and "first seen wins" are the reference's; the O(pool^2)
    ``in`` scan is replaced by a hash of the canonical (tolerance-rounded) term set.
    """
    to_index, to_xy = _index_maps(Nx)
    n_sites = Nx * Ny
    pool, seen = [], set()
    for spin in (0, 1):
        for k1 in range(n_sites):
            kx1, ky1 = to_xy(k1)
            for k2 in range(n_sites):
                kx2, ky2 = to_xy(k2)
                for q in range(1, n_sites):
                    qx, qy = to_xy(q)
                    i1 = to_index((kx1 + qx) % Nx, (ky1 + qy) % Ny, spin)
                    i2 = to_index((kx2 - qx) % Nx, (ky2 - qy) % Ny, spin ^ 1)
                    i3 = to_index(kx2, ky2, spin ^ 1)
                    i4 = to_index(kx1, ky1, spin)
                    operator = normal_ordered(_momentum_exchange(i1, i2, i3, i4))
                    plus, minus = _canonical_key(operator), _canonical_key(operator, -1)
                    if plus in seen or minus in seen:
                        continue
                    seen.add(plus)
                    pool.append(operator)
    return pool


def hubbard_interaction_pool(Nx, Ny, hermitian=False):
    """ZS / ZS2 / BCS momentum channels (reference :133-218); no driver calls this."""
    to_index, to_xy = _index_maps(Nx)
    n_sites = Nx * Ny
    channels = {'ZS channel': [], 'ZS2 channel': [], 'BCS channel': []}
    seen_zs = set()
    for spin in (0, 1):
        other = spin ^ 1
        for k1 in range(n_sites):
            kx1, ky1 = to_xy(k1)
            for k2 in range(n_sites):
                kx2, ky2 = to_xy(k2)
                for q in range(n_sites):
                    qx, qy = to_xy(q)
                    plus = ((kx1 + qx) % Nx, (ky1 + qy) % Ny)
                    minus = ((kx2 - qx) % Nx, (ky2 - qy) % Ny)
                    zs = _momentum_exchange(to_index(*plus, spin), to_index(*minus, other),
                                            to_index(kx2, ky2, other), to_index(kx1, ky1, spin), hermitian)
                    if hermitian:
                        channels['ZS channel'].append(zs)
                    else:
                        zs = normal_ordered(zs)
                        a, b = _canonical_key(zs), _canonical_key(zs, -1)
                        if a not in seen_zs and b not in seen_zs:
                            seen_zs.add(a)
                            channels['ZS channel'].append(zs)
                    channels['ZS2 channel'].append(_momentum_exchange(
                        to_index(*plus, spin), to_index(*minus, other),
                        to_index(kx2, ky2, spin), to_index(kx1, ky1, other), hermitian))
                    channels['BCS channel'].append(_momentum_exchange(
                        to_index(kx1, ky1, spin), to_index((-kx1 + qx) % Nx, (-ky1 + qy) % Ny, other),
                        to_index((-kx2 + qx) % Nx, (-ky2 + qy) % Ny, other), to_index(kx2, ky2, spin), hermitian))
    return channels


def excitations(n_electrons, n_orbitals, delta_sz=0, generalized=True):
    """Spin-orbital index lists of single / double excitations with a given delta Sz (reference :15-46)."""
    n_so = 2 * n_orbitals
    sz = [0.5 if i % 2 == 0 else -0.5 for i in range(n_so)]
    end = n_so if generalized else n_electrons
    singles = [[q, p] for q in range(end)
               for p in range(q + 1 if generalized else n_so, n_so) if sz[p] - sz[q] == delta_sz]
    doubles = []
    for s in range(end - 1):
        for r in range(s + 1, end):
            for q in range(r + 1 if generalized else n_electrons, n_so - 1):
                for p in range(q + 1, n_so):
                    if sz[p] + sz[q] - sz[r] - sz[s] == delta_sz:
                        doubles.append([s, r, q, p])
    return singles, doubles


def general_operator_pool(Nx, Ny):
    """i(a†_k1 a_k2 - h.c.) and i(a†_k1 a†_k2 a_k3 a_k4 - h.c.) over all spin orbitals, unique
    (reference :342-363; the reference keeps its chained ``k1 != k2 != k3 != k4`` test, which
    only compares neighbours, and forgets to return the list -- we return it)."""
    n_so = 2 * Nx * Ny
    pool, seen = [], set()

    def push(op):
        if not op.terms:
            key = frozenset()
        else:
            key = _canonical_key(op)
        if key not in seen:
            seen.add(key)
            pool.append(op)

    for k1 in range(n_so):
        for k2 in range(n_so):
            if k1 != k2:
                push(normal_ordered(FermionOperator(f'{k1}^ {k2}', 1j) - FermionOperator(f'{k2}^ {k1}', 1j)))
            for k3 in range(n_so):
                for k4 in range(n_so):
                    if k1 != k2 != k3 != k4:
                        push(normal_ordered(FermionOperator(f'{k1}^ {k2}^ {k3} {k4}', 1j)
                                            - FermionOperator(f'{k3}^ {k4}^ {k1} {k2}', 1j)))
    return pool


def spin_complemented_pool(n_electrons, n_orbitals, generalized=True, faithful=True):
    """Spin-complemented singles A_pq = tau_{p up,q up} + tau_{p dn,q dn} and doubles A1 (same-spin), A2 (opposite-spin)
    (reference :48-131); no driver calls it.

    ``faithful=True`` reproduces the reference's output exactly, including its stale-variable slip: inside the doubles
    loop the reference re-assigns ``s_up/s_down`` where ``p_up/p_down`` was meant (reference :111-113), so every double
    excitation is built with the ``p`` left over from the singles loop (the last orbital).  ``faithful=False`` uses the
    loop's own ``p``."""
    n_occ = n_electrons // 2
    end = n_orbitals if generalized else n_occ
    pool = []
    p_left_over = None
    for q in range(end):
        for p in range(q + 1 if generalized else n_occ, n_orbitals):
            p_left_over = p
            up = FermionOperator(f'{2 * p}^ {2 * q}') - FermionOperator(f'{2 * q}^ {2 * p}')
            down = FermionOperator(f'{2 * p + 1}^ {2 * q + 1}') - FermionOperator(f'{2 * q + 1}^ {2 * p + 1}')
            op = normal_ordered(up + down)
            if op.many_body_order() > 0:
                pool.append(op)
    for s_ in range(end):
        for r in range(s_, end):
            for q in range(r + 1 if generalized else n_occ, n_orbitals):
                for p in range(q, n_orbitals):
                    pp = p_left_over if faithful else p
                    if pp is None:
                        raise NameError("p_up is not defined (the reference fails the same way when no single excitation exists)")
                    pu, pd, qu, qd, ru, rd, su, sd = 2 * pp, 2 * pp + 1, 2 * q, 2 * q + 1, 2 * r, 2 * r + 1, 2 * s_, 2 * s_ + 1
                    same = FermionOperator(f'{pu}^ {qu}^ {ru} {su}')
                    same += FermionOperator(f'{pd}^ {qd}^ {rd} {sd}')
                    same -= hermitian_conjugated(same)
                    same = normal_ordered(same)
                    mixed = FermionOperator(f'{pu}^ {qd}^ {ru} {sd}')
                    mixed += FermionOperator(f'{pd}^ {qu}^ {rd} {su}')
                    mixed -= hermitian_conjugated(mixed)
                    mixed = normal_ordered(mixed)
                    if same.many_body_order() > 0:
                        pool.append(same)
                    if mixed.many_body_order() > 0:
                        pool.append(mixed)
    return pool


def hubbard_interation_pool_modified(Nx, Ny):
    """Five momentum channels (ZS, ZS2, W, BCS, BCS2) restricted to nearest-neighbour transfers q, each returned as ONE
    summed FermionOperator (reference :257-340, name spelled as there); no driver calls it.  Q_N = (Nx//2, Ny//2)."""
    to_index, to_xy = _index_maps(Nx)
    names = ('ZS channel', 'ZS2 channel', 'W channel', 'BCS channel', 'BCS2 channel')
    members = {name: [] for name in names}
    seen = {name: set() for name in names}
    hx, hy = Nx // 2, Ny // 2

    def push(name, i1, i2, i3, i4):
        op = normal_ordered(FermionOperator(f'{i1}^ {i2}^ {i3} {i4}'))
        key = _canonical_key(op) if op.terms else frozenset()
        if key not in seen[name]:
            seen[name].add(key)
            members[name].append(op)

    for spin in (0, 1):
        o = spin ^ 1
        for k1 in range(Nx * Ny):
            kx1, ky1 = to_xy(k1)
            for k2 in range(Nx * Ny):
                kx2, ky2 = to_xy(k2)
                for qx, qy in ((1, 0), (0, 1), (-1, 0), (0, -1)):
                    plus = to_index((kx1 + qx) % Nx, (ky1 + qy) % Ny, spin)
                    minus = to_index((kx2 - qx) % Nx, (ky2 - qy) % Ny, o)
                    push('ZS channel', plus, minus, to_index(kx2, ky2, o), to_index(kx1, ky1, spin))
                    push('ZS2 channel', plus, minus, to_index(kx1, ky1, o), to_index(kx2, ky2, spin))
                    push('W channel', to_index(kx1, ky1, spin), to_index(kx2, ky2, o),
                         to_index((kx2 + hx + qx) % Nx, (ky2 + hy + qy) % Ny, o),
                         to_index((kx1 - hx - qx) % Nx, (ky1 - hy - qy) % Ny, spin))
                    push('BCS channel', to_index(kx1, ky1, spin), to_index((-kx1 + qx) % Nx, (-ky1 + qy) % Ny, o),
                         to_index((-kx2 + qx) % Nx, (-ky2 + qy) % Ny, o), to_index(kx2, ky2, spin))
                    push('BCS2 channel', to_index(kx1, ky1, spin),
                         to_index((-kx1 + hx + qx) % Nx, (-ky1 + hy + qy) % Ny, o),
                         to_index((-kx2 + hx + qx) % Nx, (-ky2 + hy + qy) % Ny, o), to_index(kx2, ky2, spin))
    return {name: reduce(lambda a, b: a + b, members[name]) for name in names}
