"""Hamiltonian-variational ansatz on the fhsim backend (drop-in for reference ``models/hva.py``).

Circuit (reference :273-303): X prep -> W (the ansatz acts in real space) -> ``reps`` x [Coulomb layer,
vertical hopping sets, horizontal hopping sets] -> final Coulomb layer, every layer one generator with a
shared angle.  One C-ABI call per optimiser step returns <H>, <Sz>, <S^2>, the fidelity overlap and the
``5*reps+1`` parameter gradients (adjoint sweep; the reference back-propagates through the QNode).
"""
from __future__ import annotations

import os
import pickle

from fhsim import checkpoint

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from fhsim.symbolic import jordan_wigner
from linalg.exact_diagonalization import jw_get_ground_state
from operators.tools import get_interacting_term, get_quadratic_term

from .common import (Circuit, HubbardProblem, Param, State, Trotterize_generator, ensure_parent, evaluate_with_grad,
                     get_non_interacting_ground_state_index, get_particle_number_operator, get_spin_operators,
                     get_total_spin, recording, try_pyplot)
from .utils import QubitOperator_to_qmlHamiltonian, get_hva_commuting_hopping_terms

__all__ = ['HVA', 'Trotterize_generator', 'get_particle_number_operator', 'get_total_spin', 'get_spin_operators',
           'get_non_interacting_ground_state_index']


class HVA(HubbardProblem):
    ground_state_solver = staticmethod(jw_get_ground_state)

    def __init__(self, n_epoch: int, reps: int, lr: float, threshold: float, x_dimension: int, y_dimension: int,
                 n_electrons: int, n_spin_up: int, n_spin_down: int, tunneling: float, coulomb: float, periodic=True,
                 spinless=False, particle_hole_symmetry=False, load_model=False, verbose=True):
        self.n_epoch, self.reps, self.lr, self.threshold = n_epoch, reps, lr, threshold
        self.verbose = verbose
        self.setup_lattice(x_dimension, y_dimension, n_electrons, n_spin_up, n_spin_down, tunneling, coulomb,
                           periodic, spinless, particle_hole_symmetry, verbose=verbose)
        self.fermionOperators = {
            'hopping': get_quadratic_term(self.fermionHamiltonian),
            'coulomb': get_interacting_term(self.fermionHamiltonian),
            'particle number': get_particle_number_operator(x_dimension, y_dimension, spinless),
            'spin up': get_total_spin(x_dimension, y_dimension, 'spin-up'),
            'spin down': get_total_spin(x_dimension, y_dimension, 'spin-down'),
            'Sx': get_spin_operators(self.n_sites, spin_type='Sx'),
            'Sy': get_spin_operators(self.n_sites, spin_type='Sy'),
            'Sz': get_spin_operators(self.n_sites, spin_type='Sz'),
            'S^2': get_spin_operators(self.n_sites, spin_type='S^2'),
        }
        _h, _v = get_hva_commuting_hopping_terms(x_dimension, y_dimension, periodic)
        self.Nh, self.Nv = len(_h), len(_v)
        self.hvaGenerators = {
            'horizontal': [jordan_wigner(g) for g in _h],
            'vertical': [jordan_wigner(g) for g in _v],
            'coulomb': jordan_wigner(self.fermionOperators['coulomb']),
        }
        self.qmlOperators = {k: QubitOperator_to_qmlHamiltonian(self.fermionOperators[k])
                             for k in ('spin up', 'spin down', 'Sx', 'Sy', 'Sz', 'S^2')}
        tag = (f'{x_dimension}x{y_dimension} (t={tunneling}, U={coulomb}, n_electrons={n_electrons}, '
               f'up={n_spin_up}, down={n_spin_down}, reps={reps})')
        self.img_filepath = f'./images/HVA-{tag}.png'
        self.wf_filepath = (f'./results/ground_state_results/Hubbard-{x_dimension}x{y_dimension} '
                            f'(t={tunneling}, U={coulomb}, n_electrons={n_electrons}).pkl')
        self.result_filepath = f'./results/vqe_results/HVA-{tag}.pkl'
        self.model_filepath = f'./results/saved_model/HVA-{tag}.pkl'
        self._set_ground_state(*self.get_ground_state())
        self._program = None
        if load_model:
            self.load_model()
        else:
            self.params = nn.ParameterDict({
                'theta_U': nn.Parameter(torch.zeros(reps + 1), requires_grad=True),
                'theta_v': nn.Parameter(torch.zeros(reps * self.Nv), requires_grad=True),
                'theta_h': nn.Parameter(torch.zeros(reps * self.Nh), requires_grad=True),
            }).to(self.device)
            self.results = {'loss': [], 'Sz': [], 'S^2': [], 'fidelity': []}

    def _set_ground_state(self, energy, wf):
        self.ground_state_energy, self.ground_state_wf = energy, wf
        self._targets = self.upload_targets([wf])

    def fidelity_from_overlaps(self, overlaps):
        return float(np.abs(overlaps[0]) ** 2)

    def get_ground_state(self):
        return self.load_or_compute_ground_state(type(self).ground_state_solver)

    def save_model(self):
        ensure_parent(self.model_filepath)
        ensure_parent(self.result_filepath)
        with open(self.model_filepath, 'wb') as file:
            pickle.dump({'params': self.params}, file)
        with open(self.result_filepath, 'wb') as file:
            pickle.dump(self.results, file)

    def load_model(self):
        for path in (self.model_filepath, self.result_filepath):
            if not os.path.exists(path):
                raise ValueError('Please check if the file ' + path + 'exists!')
        with open(self.model_filepath, 'rb') as file:
            self.params = checkpoint.load(file)['params'].to(self.device)
        with open(self.result_filepath, 'rb') as file:
            self.results = checkpoint.load(file)

    # flat parameter layout handed to the backend: [theta_U | theta_h | theta_v]
    def build_circuit(self) -> Circuit:
        nU, nH = self.reps + 1, self.reps * self.Nh
        circuit = Circuit(self.n_qubits, nU + nH + self.reps * self.Nv)
        self._phase = self.append_basis_change(circuit)
        with recording(circuit):
            for rep in range(self.reps):
                Trotterize_generator(Param(rep), self.hvaGenerators['coulomb'])
                for i in range(self.Nv):
                    Trotterize_generator(Param(nU + nH + rep * self.Nv + i), self.hvaGenerators['vertical'][i])
                for i in range(self.Nh):
                    Trotterize_generator(Param(nU + rep * self.Nh + i), self.hvaGenerators['horizontal'][i])
            Trotterize_generator(Param(self.reps), self.hvaGenerators['coulomb'])
        return circuit

    def circuit(self, theta_U, theta_h, theta_v, mode='train'):
        if self._program is None:
            self._program = self.build_circuit().compile(self._ctx)
        prog, basis = self._program, self.basis_index()
        if mode == 'state':
            thetas = torch.cat([p.detach().reshape(-1).to(torch.float64).cpu() for p in (theta_U, theta_h, theta_v)]).numpy()
            out = State(self._ctx, self.n_qubits)
            prog.evaluate(basis, thetas, [self.device_table('H', self.qmlHamiltonian)], state_out=out)
            vec = out.numpy() * np.exp(-1j * self._phase)
            out.close()
            return torch.from_numpy(vec)
        tabs = [self.device_table('H', self.qmlHamiltonian), self.device_table('Sz', self.qmlOperators['Sz']),
                self.device_table('S^2', self.qmlOperators['S^2'])]

        def evaluator(thetas):
            res = prog.evaluate(basis, thetas, tabs, grads=True, targets=self._targets)
            self._last_overlaps = res['overlaps']
            return res['expvals'], res['grads']
        return evaluate_with_grad(evaluator, [theta_U, theta_h, theta_v])

    def run(self):
        plt = try_pyplot()
        fig = plt.figure(figsize=(12, 6)) if plt else None
        opt = optim.Adam(params=self.params.values(), lr=self.lr)
        i_epoch = len(self.results['loss'])
        while i_epoch < self.n_epoch:
            opt.zero_grad()
            loss, Sz, S_square = self.circuit(self.params['theta_U'], self.params['theta_h'], self.params['theta_v'],
                                              mode='train')
            fidelity = self.fidelity_from_overlaps(self._last_overlaps)
            loss.backward()
            opt.step()
            self.results['loss'].append(loss.item())
            self.results['Sz'].append(Sz.item())
            self.results['S^2'].append(S_square.item())
            self.results['fidelity'].append(fidelity)
            grad_vector = torch.cat((self.params['theta_U'].grad, self.params['theta_h'].grad, self.params['theta_v'].grad))
            grad_norm = torch.linalg.vector_norm(grad_vector).item()
            if self.verbose:
                print(f"iter: {len(self.results['loss'])} | loss: {loss.item(): 6f} | norm: {grad_norm: 6f} | "
                      f"fidelity: {fidelity: 6f} | Sz: {Sz.item(): 6f} | S^2: {S_square.item(): 6f}")
            if plt and (i_epoch + 1) % 10 == 0:
                self._plot(plt, fig)
            if (i_epoch + 1) % 10 == 0:
                self.save_model()
            i_epoch += 1
        self.save_model()

    def _plot(self, plt, fig):
        fig.clf()
        xs = np.arange(len(self.results['loss'])) + 1
        ax1 = fig.add_subplot(1, 2, 1)
        ax1.plot(xs, self.results['loss'], marker='X', color='r', label='HVA')
        ax1.plot(xs, np.full(len(xs), self.ground_state_energy), ls='-', color='g', label='ED')
        ax1.set_xlabel('epochs'); ax1.set_ylabel('energy'); ax1.legend(); ax1.grid()
        ax2 = fig.add_subplot(1, 2, 2)
        ax2.plot(xs, self.results['fidelity'], marker='X', ls=':', color='coral')
        ax2.set_xlabel('epochs'); ax2.set_ylabel('fidelity'); ax2.grid()
        ensure_parent(self.img_filepath)
        fig.savefig(self.img_filepath)


if __name__ == '__main__':
    vqe = HVA(n_epoch=1000, reps=10, lr=1e-2, threshold=1e-2, x_dimension=2, y_dimension=2, n_electrons=4,
              n_spin_up=2, n_spin_down=2, tunneling=1, coulomb=6, periodic=True, spinless=False,
              particle_hole_symmetry=False, load_model=False)
    vqe.run()
