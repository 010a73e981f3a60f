"""Hardware-efficient-ansatz VQE (drop-in for reference ``models/vqe_hea.py``).

The reference takes an OpenFermion ``MolecularData`` (pyscf integrals, out of scope here); this driver
accepts that duck type (``n_qubits``, ``n_electrons``, ``n_orbitals``, ``get_molecular_hamiltonian()``,
``fci_energy``) *or* directly a FermionOperator / QubitOperator Hamiltonian such as the Hubbard model.
Circuit (reference :43-57): ``reps`` x [RX, RY, RZ on every qubit; CNOT ring q -> q+1 mod n], then a final
rotation layer that re-uses parameter row ``reps-1`` (row ``reps`` is allocated but never read -- kept).
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from fhsim.backend import DeviceTable, default_context
from fhsim.circuit import Circuit
from fhsim.symbolic import FermionOperator, QubitOperator, count_qubits, jordan_wigner

from .common import evaluate_with_grad
from .utils import QubitOperator_to_qmlHamiltonian


class VQE:
    def __init__(self, molecule, n_epoch: int, reps: int, lr: float, threshold: float, seed=None, verbose=True):
        self.molecule = molecule
        self.n_epoch, self.reps, self.lr, self.threshold = n_epoch, reps, lr, threshold
        self.verbose = verbose
        if isinstance(molecule, (FermionOperator, QubitOperator)):
            self.fermionHamiltonian = molecule
            self.n_qubits = count_qubits(molecule)
            self.n_electrons = self.n_qubits // 2
            self.n_orbitals = self.n_qubits // 2
            qubit_h = jordan_wigner(molecule)
            self.reference_energy = None
        else:
            self.n_qubits = molecule.n_qubits
            self.n_electrons = molecule.n_electrons
            self.n_orbitals = molecule.n_orbitals
            self.fermionHamiltonian = molecule.get_molecular_hamiltonian()
            qubit_h = jordan_wigner(self.fermionHamiltonian)
            self.reference_energy = getattr(molecule, 'fci_energy', None)
        self.qmlHamiltonian = QubitOperator_to_qmlHamiltonian(qubit_h)
        self.device = 'cpu'
        generator = torch.Generator().manual_seed(seed) if seed is not None else None
        init = (2 * torch.rand((reps + 1, self.n_qubits, 3), generator=generator) - 1) * np.pi
        self.params = nn.ParameterList([nn.Parameter(init, requires_grad=True)]).to(self.device)
        self.loss_history = []
        self._ctx = default_context()
        self._table = DeviceTable(self._ctx, self.qmlHamiltonian.table(self.n_qubits))
        self._program = None

    def _index(self, rep, q, k):
        return (rep * self.n_qubits + q) * 3 + k

    def build_circuit(self) -> Circuit:
        n = self.n_qubits
        circuit = Circuit(n, (self.reps + 1) * n * 3)
        for rep in range(self.reps):
            for q in range(n):
                circuit.rx(0.0, q, param=self._index(rep, q, 0))
                circuit.ry(0.0, q, param=self._index(rep, q, 1))
                circuit.rz(0.0, q, param=self._index(rep, q, 2))
            for q in range(n):
                circuit.cnot(q, (q + 1) % n)
        for q in range(n):
            circuit.rx(0.0, q, param=self._index(self.reps - 1, q, 0))
            circuit.ry(0.0, q, param=self._index(self.reps - 1, q, 1))
            circuit.rz(0.0, q, param=self._index(self.reps - 1, q, 2))
        return circuit

    def circuit(self):
        if self._program is None:
            self._program = self.build_circuit().compile(self._ctx)
        prog, table = self._program, self._table

        def evaluator(thetas):
            res = prog.evaluate(0, thetas, [table], grads=True)
            return res['expvals'][:1], res['grads']
        return evaluate_with_grad(evaluator, [self.params[0]])[0]

    def run(self):
        opt = optim.Adam(self.params, lr=self.lr)
        start_time = time.time()
        for i_epoch in range(self.n_epoch):
            opt.zero_grad()
            loss = self.circuit()
            loss.backward()
            opt.step()
            self.loss_history.append(loss.item())
            if self.verbose and (i_epoch + 1) % 5 == 0:
                print(f'epoch: {i_epoch+1}, total energy: {loss.item()}')
            grad_norm = torch.linalg.vector_norm(self.params[0].grad)
            if grad_norm < self.threshold:
                if self.verbose:
                    print(f'gradient norm is less than threshold {self.threshold}, break the loop!')
                break
        if self.verbose:
            print(f'total evaluation time: {time.time()-start_time}s')


if __name__ == '__main__':
    from fhsim.symbolic import fermi_hubbard
    vqe = VQE(fermi_hubbard(2, 2, 1.0, 4.0), n_epoch=100, reps=5, lr=1e-1, threshold=0.002, seed=0)
    vqe.run()
