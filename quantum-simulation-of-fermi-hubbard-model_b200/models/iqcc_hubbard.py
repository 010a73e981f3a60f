"""iQCC on a fermionic Hamiltonian (drop-in for reference ``models/iqcc_hubbard.py``).

Per epoch: partition the *current* qubit Hamiltonian by x-mask (``partition_hamiltonian``, reference
:82-101), screen one ``Y X...X`` candidate per non-zero x-mask (reference :103-143; here the closed form
g = 2 Im<lambda|P/2|psi> on the QMF state, one kernel for all candidates), optimise (theta, phi, tau) with
Adam, then dress the Hamiltonian symbolically (reference :184-189) and re-upload its packed table.
"""
from __future__ import annotations

from functools import partial

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from fhsim.backend import DevicePool, DeviceTable, default_context
from fhsim.circuit import Circuit
from fhsim.recording import Param, recording
from fhsim.symbolic import (FermionOperator, InteractionOperator, QubitOperator, count_qubits,
                            get_interaction_operator, jordan_wigner)
from fhsim.tables import GeneratorPlan
from linalg.exact_diagonalization import get_ground_state, get_sparse_operator

from .common import evaluate_with_grad, try_pyplot
from .utils import PauliStringRotation, QubitOperator_to_qmlHamiltonian


class IQCC:
    def __init__(self, fermion_hamiltonian, n_epoch: int, lr: float, threshold: float, verbose=True):
        self.n_epoch, self.lr, self.threshold = n_epoch, lr, threshold
        self.Ng = 0
        self.ratio = 0.1
        self.verbose = verbose
        self.n_qubits = count_qubits(fermion_hamiltonian)
        self.n_electrons = self.n_qubits // 2
        self.device = torch.device('cpu')
        if isinstance(fermion_hamiltonian, FermionOperator):
            fermion_hamiltonian = get_interaction_operator(fermion_hamiltonian)
        self.fermionHamiltonian = fermion_hamiltonian
        self.currentHamiltonian = jordan_wigner(self.fermionHamiltonian)
        self.qmlHamiltonian = QubitOperator_to_qmlHamiltonian(self.currentHamiltonian)
        self.params = nn.ParameterDict({
            'theta': nn.Parameter(torch.Tensor([np.pi] * self.n_electrons + [0] * (self.n_qubits - self.n_electrons)),
                                  requires_grad=True),
            'phi': nn.Parameter(torch.zeros(self.n_qubits), requires_grad=True),
            'tau': nn.Parameter(torch.zeros(self.Ng), requires_grad=True),
        }).to(self.device)
        self.loss_history = {'iteration': [], 'epoch': []}
        self.selectedGates = []
        self._ctx = default_context()
        self._table = None
        self._table_source = None
        self._program = None
        self._program_key = None
        self.ground_state_energy, self.ground_state_wf = get_ground_state(
            get_sparse_operator(fermion_hamiltonian, self.n_qubits))

    # -- backend plumbing ------------------------------------------------------------------------
    def _device_table(self):
        if self._table_source is not self.qmlHamiltonian:
            if self._table is not None:
                self._table.close()
            self._table = DeviceTable(self._ctx, self.qmlHamiltonian.table(self.n_qubits))
            self._table_source = self.qmlHamiltonian
        return self._table

    def _build(self, gates):
        """RY(theta_i) RZ(phi_i) on every qubit (QMF state), then the entanglers with parameters tau."""
        n = self.n_qubits
        circuit = Circuit(n, 2 * n + len(gates))
        for i in range(n):
            circuit.ry(0.0, i, param=i)
            circuit.rz(0.0, i, param=n + i)
        circuit.marker('qmf_end')
        with recording(circuit):
            for i, gate in enumerate(gates):
                gate(Param(2 * n + i))
        return circuit

    def get_circuit(self, tau=None, appendGates=None):
        """<H> of the current Hamiltonian.  Without ``appendGates``: the trainable circuit
        (theta, phi, tau all receive gradients).  With ``appendGates``: the screening circuit, QMF
        angles detached, gradient w.r.t. ``tau`` only (reference :59-80)."""
        table = self._device_table()
        n = self.n_qubits
        if not appendGates:
            key = tuple(id(g) for g in self.selectedGates)
            if self._program is None or key != self._program_key:
                if self._program is not None:
                    self._program.close()
                self._program = self._build(self.selectedGates).compile(self._ctx)
                self._program_key = key
            prog = self._program

            def evaluator(thetas):
                res = prog.evaluate(0, thetas, [table], grads=True)
                return res['expvals'][:1], res['grads']
            return evaluate_with_grad(evaluator, [self.params['theta'], self.params['phi'], self.params['tau']])[0]

        # screening: every appended gate is exp(-i tau P/2) at tau = 0  ->  g = 2 Im <lambda| P/2 |psi>
        prog = self._build([]).compile(self._ctx)
        plans = [GeneratorPlan(QubitOperator(tuple(zip(g.keywords['pauliString'][1], g.keywords['pauliString'][0])), 0.5),
                               n) for g in appendGates]
        pool = DevicePool(self._ctx, plans, n)
        qmf = torch.cat([self.params['theta'].detach().to(torch.float64), self.params['phi'].detach().to(torch.float64)])

        def evaluator(tau_values):
            res = prog.evaluate(0, qmf.numpy(), [table], pool=pool, pool_pos=prog.markers['qmf_end'])
            return res['expvals'][:1], res['pool']
        loss = evaluate_with_grad(evaluator, [tau])[0]
        self._screen_objects = (prog, pool)
        return loss

    def partition_hamiltonian(self):
        """Group the current Hamiltonian's strings by the set of qubits they flip (X or Y)."""
        partitioned = {}
        for pauliString, coeff in self.currentHamiltonian.terms.items():
            flip = tuple(index for index, pauli in pauliString if pauli in ('X', 'Y'))
            piece = coeff * QubitOperator(pauliString)
            if flip in partitioned:
                partitioned[flip] += piece
            else:
                partitioned[flip] = piece
        return partitioned

    def select_operator(self):
        DIS_gates, DIS_strings = [], []
        for flip in self.partition_hamiltonian().keys():
            if len(flip) == 0:
                continue
            Pk = ('Y' + 'X' * (len(flip) - 1), flip)
            DIS_gates.append(partial(PauliStringRotation, pauliString=Pk))
            DIS_strings.append(' '.join(p + str(i) for p, i in zip(*Pk)))
        tau = nn.Parameter(torch.zeros(len(DIS_gates)), requires_grad=True)
        loss = self.get_circuit(tau, DIS_gates)
        loss.backward()
        grads = np.abs(tau.grad.detach().cpu().numpy())
        self.last_screening = (list(DIS_strings), grads.copy())      # every candidate with its |gradient| (float32)
        prog, pool = self._screen_objects
        prog.close()
        pool.close()
        max_grad = np.max(grads) if len(grads) else 0.0
        if max_grad * self.ratio > self.threshold:
            self.Ng = int(np.sum(grads > max_grad * self.ratio))
        else:
            self.Ng = int(np.sum(grads > self.threshold))
        maxIndicies = np.argsort(grads)[::-1][:self.Ng]
        return ([DIS_gates[i] for i in maxIndicies], [DIS_strings[i] for i in maxIndicies], grads[maxIndicies])

    def dress_hamiltonian(self, operators, taus):
        """H <- H + sin(tau)(-i/2)[H, P] + (1/2)(1 - cos tau)(P H P - H), last entangler first (reference :184-189).

        Evaluated on the packed (x-mask, z-mask, coefficient) table ON THE DEVICE (``fh_ptable_dress``: anticommuting
        terms pick up cos/sin copies, the one possible duplicate per string is hash-merged; bit-identical to the host
        restatement ``PauliTable.dressed``), not by symbolic operator products: the term table grows geometrically with
        the number of entanglers.  ``currentHamiltonian`` stays a QubitOperator for the callers
        that iterate its terms (``partition_hamiltonian``)."""
        from fhsim.backend import DevicePauliTable
        from fhsim.tables import PauliTable, pack_term
        host = PauliTable.from_operator(self.currentHamiltonian, self.n_qubits, compress=False)
        dev = DevicePauliTable(self._ctx, host)              # every entangler is applied on the device: ONE download
        try:
            for P_k, tau_k in zip(operators[::-1], taus[::-1]):
                (term, coeff), = P_k.terms.items()
                if abs(coeff - 1.0) > 1e-12:
                    raise ValueError('entanglers must be bare Pauli strings')
                xp, zp = pack_term(term, self.n_qubits)
                # float32 parameter, promoted exactly: keeps the algebra in double precision
                dev.dress(xp, zp, float(tau_k))
            table = dev.to_host()
        finally:
            dev.close()
        self.currentHamiltonian = table.to_operator()
        self.qmlHamiltonian = QubitOperator_to_qmlHamiltonian(self.currentHamiltonian)

    def run(self):
        plt = try_pyplot()
        for i_epoch in range(self.n_epoch):
            maxGates, maxOperators, maxGrads = self.select_operator()
            if self.verbose:
                print(f'=== Found operators: {maxOperators} \n with gradients: {maxGrads} ====')
            if len(maxGrads) == 0:
                break
            self.selectedGates = maxGates
            self.params['tau'] = torch.zeros(self.Ng).to(self.device)
            opt = optim.Adam(self.params.values(), lr=self.lr)
            while True:
                opt.zero_grad()
                loss = self.get_circuit()
                loss.backward()
                opt.step()
                self.loss_history['iteration'].append(loss.item())
                grad_vec = torch.cat((self.params['theta'].grad, self.params['phi'].grad, self.params['tau'].grad))
                grad_norm = torch.linalg.vector_norm(grad_vec)
                if grad_norm < self.threshold:
                    break
                if self.verbose:
                    print(loss.item(), grad_norm.item())
            self.loss_history['epoch'].append(loss.item())
            self.selectedGates = []
            self.dress_hamiltonian([QubitOperator(op) for op in maxOperators],
                                   self.params['tau'].detach().cpu().numpy())
            if self.verbose:
                print(f'epoch: {i_epoch+1}, total energy: {loss.item()}')
        if plt:
            self._plot(plt)

    def _plot(self, plt):
        import os
        fig = plt.figure(figsize=(12, 6))
        for k, key in enumerate(('iteration', 'epoch')):
            ax = fig.add_subplot(1, 2, k + 1)
            ys = self.loss_history[key]
            xs = np.arange(len(ys)) + 1
            ax.plot(xs, ys, marker='x', color='r', label='iqcc')
            ax.plot(xs, np.full(len(ys), self.ground_state_energy), linestyle='-', color='g', label='ED')
            ax.set_xlabel(key + 's'); ax.set_ylabel('energy'); ax.legend(); ax.grid()
        os.makedirs('./images', exist_ok=True)
        fig.savefig('./images/`iqcc-hubbard-2x2.png')


if __name__ == '__main__':
    from fhsim.symbolic import fermi_hubbard
    hamiltonian = fermi_hubbard(x_dimension=2, y_dimension=2, tunneling=1, coulomb=4, periodic=True, spinless=False)
    vqe = IQCC(fermion_hamiltonian=hamiltonian, n_epoch=100, lr=1e-2, threshold=5e-3)
    vqe.run()
