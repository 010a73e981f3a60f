"""Test-only numpy slab engine with the interface of ``fhsim.sharded.CudaEngine``.

Used by the world_size>1 gloo tests on CPU to exercise the sharded planner / lowering / collectives
(`fhsim/sharded.py`) without a GPU.  It enforces the same argument rules as the device entry points
(``fh_program_add_pair`` / ``fh_pool_upload``) so that an op the C-ABI would reject fails here too.
"""
import numpy as np

from emulate import apply_op
from fhsim.circuit import DiagOpSpec


def _par(v):
    return (np.bitwise_count(v) & 1).astype(np.int64)


def _check_pair_rules(x, fixmask, fixval, n):
    full = (1 << n) - 1
    assert x and (x | fixmask | fixval) & ~full == 0, "mask outside the slab"
    top = 1 << (x.bit_length() - 1)
    assert fixmask & top and not fixval & top, "pattern must pin the top x bit to 0 (device rule)"
    assert fixval & ~fixmask == 0
    assert bin(fixmask).count("1") <= 16


class NumpyEngine:
    def __init__(self, n_local, dist=None):
        self.n_local = int(n_local)
        self.dist = dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.calls = {"run_ops": 0, "swap_bits": 0, "all_to_all": 0, "apply_table": 0, "pool_partial": 0}

    def new_state(self):
        return [np.zeros(1 << self.n_local, dtype=np.complex128)]

    def free_state(self, h):
        h[0] = None

    def set_basis(self, h, local_index):
        h[0][:] = 0
        if local_index is not None:
            h[0][int(local_index)] = 1.0

    def copy(self, dst, src):
        dst[0] = src[0].copy()

    def run_ops(self, h, ops, thetas, n_params, key=None):
        self.calls["run_ops"] += 1
        psi = h[0]
        for op in ops:
            if isinstance(op, DiagOpSpec):
                assert all(0 <= z < (1 << self.n_local) for z in op.z)
            else:
                _check_pair_rules(op.x, op.fixmask, op.fixval, self.n_local)
            psi = apply_op(psi, op, thetas, self.n_local)
        h[0] = psi

    def swap_bits(self, h, pairs):
        self.calls["swap_bits"] += 1
        seen = set()
        for a, b in pairs:
            assert a != b and a not in seen and b not in seen and 0 <= a < self.n_local and 0 <= b < self.n_local
            seen.update((a, b))
        assert len(pairs) <= 8
        idx = np.arange(1 << self.n_local, dtype=np.uint64)
        j = idx.copy()
        for a, b in pairs:
            t = ((j >> np.uint64(a)) ^ (j >> np.uint64(b))) & np.uint64(1)
            j ^= (t << np.uint64(a)) | (t << np.uint64(b))
        h[0] = h[0][j]

    def all_to_all(self, h):
        self.calls["all_to_all"] += 1
        if self.world == 1:
            return
        import torch
        src = torch.from_numpy(np.ascontiguousarray(h[0]).view(np.float64).copy())
        from fhsim.sharded import pairwise_all_to_all
        chunks_out = pairwise_all_to_all(self.dist, list(src.chunk(self.world)))
        h[0] = torch.cat(chunks_out).numpy().view(np.complex128).copy()

    def apply_table(self, table, h_in, h_out, accumulate):
        self.calls["apply_table"] += 1
        n = self.n_local
        psi = h_in[0]
        idx = np.arange(1 << n, dtype=np.uint64)
        acc = np.zeros_like(psi)
        ipow = (1, 1j, -1, -1j)
        for x, z, c in zip(table.x, table.z, table.coeff):
            x, z = int(x), int(z)
            assert (x | z) >> n == 0
            k = bin(x & z).count("1") & 3
            j = idx ^ np.uint64(x)
            acc += c * ipow[k] * (1 - 2 * _par(j & np.uint64(z))) * psi[j]
        if h_out is not None:
            h_out[0] = (h_out[0] + acc) if accumulate else acc
        return complex(np.vdot(psi, acc))

    def pool_partial(self, entries, h_psi, h_lam, n_out, sector=None):
        self.calls["pool_partial"] += 1
        n = self.n_local
        psi, lam = h_psi[0], h_lam[0]
        idx = np.arange(1 << n, dtype=np.uint64)
        out = np.zeros(n_out)
        prev = -1
        for x, fm, fv, ze, b, o in entries:
            _check_pair_rules(x, fm, fv, n)
            assert o >= prev, "entries must be sorted by output"
            prev = o
            i = idx[(idx & np.uint64(fm)) == np.uint64(fv)]
            j = i ^ np.uint64(x)
            s = 1 - 2 * _par(i & np.uint64(ze))
            val = np.sum(np.conj(lam[i]) * s * b * psi[j] + np.conj(lam[j]) * s * np.conj(b) * psi[i])
            out[o] += 2.0 * val.imag
        return out

    def sync(self):
        pass

    def inner(self, ha, hb):
        return complex(np.vdot(ha[0], hb[0]))

    def all_reduce(self, arr):
        arr = np.asarray(arr, dtype=np.float64)
        if self.world == 1:
            return arr
        import torch
        t = torch.from_numpy(arr.copy())
        self.dist.all_reduce(t)
        return t.numpy()

    def all_gather_slab(self, h):
        if self.world == 1:
            return [h[0]]
        import torch
        mine = torch.from_numpy(np.ascontiguousarray(h[0]).view(np.float64).copy())
        outs = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(outs, mine)
        return [o.numpy().view(np.complex128) for o in outs]
