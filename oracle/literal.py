"""ORACLE / CPU BASELINE (test infrastructure only) -- the reference's literal CPU path.

What the reference executes through PennyLane ``default.qubit.torch`` + torch autograd
(third-party, not vendored, unpinned; behaviour restated from SURVEY A.3):

* state: complex128 torch tensor of shape [2]*n, wire 0 = first axis;
* every gate of ``PauliStringRotation`` (reference ``models/utils.py:58-83``) applied one at a time:
  RY/RX by tensordot, CNOT by slice-and-stack, RZ as a diagonal;
* ``Trotterize_generator`` (``models/adapt_vqe.py:87-98``): angle 2*theta*Re(c) per string, dict order;
* ``expval(Hamiltonian)`` term by term by index gather (``adapt_vqe.py:357,361``);
* pool screening = append every pool operator with e = 0 and ``loss.backward()``
  (``adapt_vqe.py:297-310, 336-341``).

This is the form timed as "the reference CPU path" (bench.py --impl reference / cpu_baseline).
"""
from __future__ import annotations

import math

import numpy as np
import torch

CD = torch.complex128


def _mat(rows):
    return torch.stack([torch.stack(r) for r in rows])


def _c(v):
    return v.to(CD) if isinstance(v, torch.Tensor) else torch.tensor(v, dtype=CD)


class LiteralSimulator:
    def __init__(self, n):
        self.n = n
        self.state = torch.zeros([2] * n, dtype=CD)
        self.state[(0,) * n] = 1.0
        self.gate_passes = 0

    # --- elementary gates ---------------------------------------------------------------
    def _apply_1q(self, mat, wire):
        s = torch.tensordot(mat, self.state, dims=([1], [wire]))
        self.state = torch.movedim(s, 0, wire)
        self.gate_passes += 1

    def pauli_x(self, wire):
        self.state = torch.flip(self.state, dims=[wire])
        self.gate_passes += 1

    def rx(self, phi, wire):
        phi = torch.as_tensor(phi, dtype=torch.float64)
        c, s = _c(torch.cos(phi / 2)), _c(torch.sin(phi / 2))
        self._apply_1q(_mat([[c, -1j * s], [-1j * s, c]]), wire)

    def ry(self, phi, wire):
        phi = torch.as_tensor(phi, dtype=torch.float64)
        c, s = _c(torch.cos(phi / 2)), _c(torch.sin(phi / 2))
        self._apply_1q(_mat([[c, -s], [s, c]]), wire)

    def rz(self, phi, wire):
        phi = torch.as_tensor(phi, dtype=torch.float64)
        ph = torch.exp(-0.5j * _c(phi))
        diag = torch.stack([ph, torch.conj(ph)])
        shape = [1] * self.n
        shape[wire] = 2
        self.state = self.state * diag.reshape(shape)
        self.gate_passes += 1

    def cnot(self, control, target):
        s0 = self.state.select(control, 0)
        s1 = self.state.select(control, 1)
        t = target if target < control else target - 1
        self.state = torch.stack([s0, torch.flip(s1, dims=[t])], dim=control)
        self.gate_passes += 1

    def single_excitation(self, phi, wire_i, wire_j):
        phi = torch.as_tensor(phi, dtype=torch.float64)
        c, s = _c(torch.cos(phi / 2)), _c(torch.sin(phi / 2))
        o, z = _c(1.0), _c(0.0)
        m = _mat([[o, z, z, z], [z, c, -s, z], [z, s, c, z], [z, z, z, o]]).reshape(2, 2, 2, 2)
        s_ = torch.tensordot(m, self.state, dims=([2, 3], [wire_i, wire_j]))
        self.state = torch.movedim(s_, [0, 1], [wire_i, wire_j])
        self.gate_passes += 1

    # --- reference composites -----------------------------------------------------------
    def pauli_string_rotation(self, theta, letters, wires):
        """reference models/utils.py:58-83."""
        for p, q in zip(letters, wires):
            if p == 'X':
                self.ry(-math.pi / 2, q)
            elif p == 'Y':
                self.rx(math.pi / 2, q)
        for a, b in zip(wires[:-1], wires[1:]):
            self.cnot(a, b)
        self.rz(theta, wires[-1])
        for a, b in zip(reversed(wires[:-1]), reversed(wires[1:])):
            self.cnot(a, b)
        for p, q in zip(letters, wires):
            if p == 'X':
                self.ry(math.pi / 2, q)
            elif p == 'Y':
                self.rx(-math.pi / 2, q)

    def trotterize(self, theta, strings):
        """strings: [(letters, wires, coeff)] in generator.terms order (identity skipped)."""
        for letters, wires, coeff in strings:
            if not wires:
                continue
            self.pauli_string_rotation(2 * theta * coeff.real, letters, wires)

    def basis_change(self, diagonal, decomposition):
        for q in range(self.n):
            self.rz(float(np.angle(diagonal[q])), q)
        for layer in reversed(list(decomposition)):
            for (i, j, theta, phi) in layer:
                self.single_excitation(2 * theta, i, j)
                self.rz(phi, j)

    def expval(self, terms):
        """terms: [(x_mask, z_mask, coeff)] packed with wire 0 = MSB; per-term gather."""
        flat = self.state.reshape(-1)
        n = self.n
        idx = torch.arange(1 << n, dtype=torch.int64)
        total = torch.zeros((), dtype=torch.float64)
        for x, z, coeff in terms:
            col = idx ^ int(x)
            k = bin(int(x) & int(z)).count("1") & 3
            par = col & int(z)                  # parity of the set bits: XOR-fold (vectorised, 6 passes)
            for sh in (32, 16, 8, 4, 2, 1):
                par = par ^ (par >> sh)
            par = par & 1
            data = (1 - 2 * par).to(CD) * complex([1, 1j, -1, -1j][k])
            val = torch.sum(torch.conj(flat) * data * flat[col])
            total = total + (complex(coeff) * val).real
        return total


def strings_of(table, n):
    """{(x,z): c} -> [(letters, wires, c)] with wires ascending (QubitOperator key order)."""
    out = []
    for (x, z), c in table.items():
        letters, wires = [], []
        for q in range(n):
            b = 1 << (n - 1 - q)
            if x & b and z & b:
                letters.append('Y'); wires.append(q)
            elif x & b:
                letters.append('X'); wires.append(q)
            elif z & b:
                letters.append('Z'); wires.append(q)
        out.append((letters, wires, complex(c)))
    return out


def adapt_eval_circuit(n, occupied, selected, t_params, pool, e_params, diagonal, decomposition, h_terms):
    """reference ADAPT.circuit(mode='eval') (models/adapt_vqe.py:325-361) -> <H> with autograd graph."""
    sim = LiteralSimulator(n)
    for q in occupied:
        sim.pauli_x(q)
    for i, strings in enumerate(selected):
        sim.trotterize(t_params[i], strings)
    for i, strings in enumerate(pool):
        sim.trotterize(e_params[i], strings)
    sim.basis_change(diagonal, decomposition)
    return sim.expval(h_terms), sim


def screen_by_backprop(n, occupied, selected, thetas, pool, diagonal, decomposition, h_terms,
                       chunk=None, dtype=torch.float32):
    """d<H>/d e_k at e=0 for every pool element by append-and-backprop, float32 parameters and result as
    the reference has them (adapt_vqe.py:306-310; ``dtype=torch.float64`` keeps the parameters in double so the
    result can be compared with the closed form at 1e-12).  ``chunk`` evaluates the pool in slices of that
    many operators (identical arithmetic per operator, bounded autograd memory)."""
    np_dtype = np.float32 if dtype == torch.float32 else np.float64
    t = torch.tensor(np.asarray(thetas, dtype=np_dtype), dtype=dtype)
    grads = np.zeros(len(pool), dtype=np_dtype)
    passes = 0
    step = len(pool) if not chunk else chunk
    for lo in range(0, len(pool), step):
        sub = pool[lo:lo + step]
        e = torch.zeros(len(sub), dtype=dtype, requires_grad=True)
        loss, sim = adapt_eval_circuit(n, occupied, selected, t, sub, e, diagonal, decomposition, h_terms)
        loss.backward()
        grads[lo:lo + step] = e.grad.numpy()
        passes += sim.gate_passes
    return grads, passes
