"""ADAPT-VQE for the Fermi-Hubbard model on the fhsim backend (drop-in for reference ``models/adapt_vqe.py``).

Same class name, constructor arguments, methods (``get_ground_state``, ``get_ground_state_properties``,
``save_model``, ``load_model``, ``select_operator``, ``circuit``, ``run``), parameter names
(``params['e']``, ``params['t']``) and ``results`` keys as the reference.  What changed is underneath:

* pool screening (reference :297-323) is the closed form g_k = 2 Im <lambda|G_k|psi>, lambda = W† H W psi,
  evaluated for the whole pool by one CUDA kernel instead of appending 324 operators and back-propagating;
* the optimiser step (reference :402-419) is ONE C-ABI call returning <H>, <Sz>, <S^2>, the fidelity
  overlaps and d<H>/dt (adjoint sweep) -- the reference runs the circuit twice and back-propagates;
* exact diagonalisation (reference :221-247) is Lanczos on the GPU.
"""
from __future__ import annotations

import os
import pickle

from fhsim import checkpoint
import time
from functools import partial

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from linalg.exact_diagonalization import jw_get_ground_state
from operators.pool import hubbard_interaction_pool_simplified

from .common import (Circuit, DevicePool, HubbardProblem, Param, State, Trotterize_generator, ensure_parent,
                     evaluate_with_grad, generator_plan, get_non_interacting_ground_state_index,
                     get_particle_number_operator, get_spin_operators, get_total_spin, print_list, recording,
                     try_pyplot)
from .utils import PauliStringRotation, QubitOperator_to_qmlHamiltonian
from fhsim.symbolic import jordan_wigner

__all__ = ['ADAPT', 'Trotterize_generator', 'print_list', 'get_particle_number_operator', 'get_total_spin',
           'get_spin_operators', 'get_non_interacting_ground_state_index', 'PauliStringRotation']


class ADAPT(HubbardProblem):
    ground_state_solver = staticmethod(jw_get_ground_state)
    file_tag = 'ADAPT'

    def __init__(self, n_epoch: int, threshold1: float, threshold2: float, x_dimension: int, y_dimension: int,
                 n_electrons: int, n_spin_up: int, n_spin_down: int, tunneling: float, coulomb: float,
                 periodic=True, spinless=False, particle_hole_symmetry=False, load_model=False, verbose=True,
                 tie_break='numpy'):
        self.fermionOperatorPool = hubbard_interaction_pool_simplified(x_dimension, y_dimension)
        self.qubitOperatorPool = [jordan_wigner(g) for g in self.fermionOperatorPool]
        self.gateOperatorPool = [partial(Trotterize_generator, generator=g) for g in self.qubitOperatorPool]
        self.n_epoch, self.threshold1, self.threshold2 = n_epoch, threshold1, threshold2
        self.ratio = 0.1
        self.verbose = verbose
        self.tie_break = tie_break              # 'numpy': literal np.argsort(...)[::-1]; 'stable': (-|g|, index)
        self.setup_lattice(x_dimension, y_dimension, n_electrons, n_spin_up, n_spin_down, tunneling, coulomb,
                           periodic, spinless, particle_hole_symmetry, verbose=verbose)
        self.fermionOperators = {
            'spin up': get_total_spin(self.n_sites, spin_type='spin-up'),
            'spin down': get_total_spin(self.n_sites, spin_type='spin-down'),
            'Sx': get_spin_operators(self.n_sites, spin_type='Sx'),
            'Sy': get_spin_operators(self.n_sites, spin_type='Sy'),
            'Sz': get_spin_operators(self.n_sites, spin_type='Sz'),
            'S^2': get_spin_operators(self.n_sites, spin_type='S^2'),
        }
        self.qmlOperators = {k: QubitOperator_to_qmlHamiltonian(v) for k, v in self.fermionOperators.items()}
        self._k_interacting_term = None
        tag = (f'{x_dimension}x{y_dimension} (t={tunneling}, U={coulomb}, n_electrons={n_electrons}, '
               f'up={n_spin_up}, down={n_spin_down})')
        self.img_filepath = f'./images/{self.file_tag}-{tag}.png'
        self.wf_filepath = (f'./results/ground_state_results/Hubbard-{x_dimension}x{y_dimension} '
                            f'(t={tunneling}, U={coulomb}, n_electrons={n_electrons}).pkl')
        self.result_filepath = f'./results/vqe_results/{self.file_tag}-{tag}.pkl'
        self.model_filepath = f'./results/saved_model/{self.file_tag}-{tag}.pkl'
        self._set_ground_state(*self.get_ground_state())

        self._pool = DevicePool(self._ctx, [generator_plan(g, self.n_qubits) for g in self.qubitOperatorPool],
                                self.n_qubits)
        self._program = None
        self._program_key = None
        if load_model:
            self.load_model()
        else:
            self.params = nn.ParameterDict({
                'e': nn.Parameter(torch.zeros(len(self.gateOperatorPool)), requires_grad=True),
                't': nn.Parameter(torch.Tensor([]), requires_grad=True),
            }).to(self.device)
            self.selected_gates = []
            self.results = {'epoch loss': [], 'iteration loss': [], 'Sz': [], 'S^2': [], 'fidelity': [],
                            'n_params': [], 'selected operators': []}

    # the symbolic FT of the quartic term is expensive and never used by the reference: computed on demand
    @property
    def k_interacting_term(self):
        if self._k_interacting_term is None:
            from operators.fourier import fourier_transform
            self._k_interacting_term = fourier_transform(self.interacting_term, self.x_dimension, self.y_dimension)
        return self._k_interacting_term

    # -- ground state ---------------------------------------------------------------------------
    def _set_ground_state(self, energy, wf):
        self.ground_state_energy, self.ground_state_wf = energy, wf
        self._targets = self.upload_targets([wf])

    def get_ground_state(self):
        return self.load_or_compute_ground_state(type(self).ground_state_solver)

    def get_ground_state_properties(self):
        print('ground state energy: ', self.ground_state_energy)
        print('particle number: ', self.n_electrons)
        print('')

    def fidelity_from_overlaps(self, overlaps):
        return float(np.abs(overlaps[0]) ** 2)

    # -- checkpoints ----------------------------------------------------------------------------
    def _pool_index_of(self, gate):
        """Index of a selected gate in the pool (identity first, then by generator: unpickled gates are copies)."""
        for i, g in enumerate(self.gateOperatorPool):
            if g is gate:
                return i
        gen = gate.keywords['generator'] if hasattr(gate, 'keywords') else None
        for i, g in enumerate(self.gateOperatorPool):
            if gen is not None and g.keywords['generator'] == gen:
                return i
        raise ValueError('selected gate is not in the operator pool')

    def save_model(self):
        """Reference layout (pickle of the ParameterDict + gate closures, adapt_vqe.py:269-280) plus a portable twin
        ``<model>.npz`` (pool indices of the selected operators, parameters, results as JSON): the pickles embed
        torch / operator classes and do not travel between installations."""
        import json
        ensure_parent(self.model_filepath)
        ensure_parent(self.result_filepath)
        with open(self.model_filepath, 'wb') as file:
            pickle.dump({'params': self.params, 'circuit': self.selected_gates}, file)
        with open(self.result_filepath, 'wb') as file:
            pickle.dump(self.results, file)
        indices = [self._pool_index_of(g) for g in self.selected_gates]
        portable = {k: v for k, v in self.results.items() if k != 'selected operators'}
        np.savez(self.model_filepath + '.npz', selected_indices=np.asarray(indices, dtype=np.int64),
                 t=self.params['t'].detach().cpu().numpy(), results_json=np.asarray(json.dumps(portable)))

    def load_model(self):
        """Pickles if both exist (reference behaviour, adapt_vqe.py:282-295), else the portable ``.npz`` twin."""
        import json
        if os.path.exists(self.model_filepath) and os.path.exists(self.result_filepath):
            with open(self.model_filepath, 'rb') as file:
                state_dict = checkpoint.load(file)
            self.params = state_dict['params'].to(self.device)
            self.selected_gates = state_dict['circuit']
            with open(self.result_filepath, 'rb') as file:
                self.results = checkpoint.load(file)
            return
        twin = self.model_filepath + '.npz'
        if not os.path.exists(twin):
            raise ValueError('Please check if the file ' + self.model_filepath + 'exists!')
        data = np.load(twin)
        indices = [int(i) for i in data['selected_indices']]
        self.selected_gates = [self.gateOperatorPool[i] for i in indices]
        self.params = nn.ParameterDict({
            'e': nn.Parameter(torch.zeros(len(self.gateOperatorPool)), requires_grad=True),
            't': nn.Parameter(torch.from_numpy(np.asarray(data['t'], dtype=np.float32)), requires_grad=True),
        }).to(self.device)
        self.results = json.loads(str(data['results_json']))
        self.results['selected operators'] = [self.fermionOperatorPool[i] for i in indices]

    # -- circuit --------------------------------------------------------------------------------
    def build_circuit(self) -> Circuit:
        """Gate list of reference ``circuit()``: X prep is the initial basis state, then the selected
        generators (parameters t), then W."""
        circuit = Circuit(self.n_qubits, len(self.selected_gates))
        with recording(circuit):
            for i, gate in enumerate(self.selected_gates):
                gate(Param(i))
        circuit.marker('ansatz_end')
        self._phase = self.append_basis_change(circuit)
        return circuit

    def _compiled(self):
        key = tuple(id(g.keywords['generator']) if hasattr(g, 'keywords') else id(g) for g in self.selected_gates)
        if self._program is None or key != self._program_key:
            if self._program is not None:
                self._program.close()
            self._program = self.build_circuit().compile(self._ctx)
            self._program_key = key
        return self._program

    def _observables(self, mode):
        tabs = [self.device_table('H', self.qmlHamiltonian)]
        if mode == 'train':
            tabs += [self.device_table('Sz', self.qmlOperators['Sz']), self.device_table('S^2', self.qmlOperators['S^2'])]
        return tabs

    def circuit(self, mode='train'):
        """Execute the circuit.  'train' -> (<H>, <Sz>, <S^2>) with d<H>/dt attached; 'eval' -> <H> whose
        backward fills params['e'].grad with the pool gradients; 'state' -> the complex128 state vector."""
        prog = self._compiled()
        basis = self.basis_index()
        t = self.params['t']
        if mode == 'state':
            out = State(self._ctx, self.n_qubits)
            thetas = t.detach().to(torch.float64).cpu().numpy()
            prog.evaluate(basis, thetas, self._observables('eval'), state_out=out)
            vec = out.numpy() * np.exp(-1j * self._phase)
            out.close()
            return torch.from_numpy(vec)
        if mode == 'train':
            def evaluator(thetas):
                res = prog.evaluate(basis, thetas, self._observables('train'), grads=True, targets=self._targets)
                self._last_overlaps = res['overlaps']
                return res['expvals'], res['grads']
            return evaluate_with_grad(evaluator, [t])
        if mode == 'eval':
            e = self.params['e']
            nt = t.numel()

            def evaluator(thetas):
                res = prog.evaluate(basis, thetas[:nt], self._observables('eval'), pool=self._pool,
                                    pool_pos=prog.markers['ansatz_end'])
                grads = np.concatenate([np.zeros(nt), res['pool']])
                return res['expvals'][:1], grads
            return evaluate_with_grad(evaluator, [t, e])[0]
        raise ValueError(f'unknown mode {mode!r}')

    # -- operator selection (reference :297-323) ---------------------------------------------------
    def select_operator(self):
        loss = self.circuit(mode='eval')
        loss.backward()
        grads = self.params['e'].grad.cpu().numpy()
        grads = np.abs(grads)
        self.params['e'].grad.zero_()
        if self.params['t'].grad is not None:
            self.params['t'].grad.zero_()
        max_grad = np.max(grads)
        self.Ng = int(np.sum((grads >= max_grad * self.ratio) * (grads >= self.threshold1)))
        if self.tie_break == 'stable':
            order = np.argsort(-grads, kind='stable')
        else:
            order = np.argsort(grads)[::-1]
        selected_indices = order[:self.Ng].tolist()
        self.last_selected_indices = selected_indices
        selected_operator = [self.fermionOperatorPool[i] for i in selected_indices]
        selected_gates = [self.gateOperatorPool[i] for i in selected_indices]
        max_grads = [grads[i] for i in selected_indices]
        return selected_operator, selected_gates, max_grads

    def _fidelity(self):
        return self.fidelity_from_overlaps(self._last_overlaps)

    # -- main loop (reference :363-467) ----------------------------------------------------------
    def run(self):
        self.get_ground_state_properties()
        start_time = time.time()
        plt = try_pyplot()
        fig = plt.figure(figsize=(12, 6)) if plt else None
        i_epoch = len(self.results['epoch loss'])
        while i_epoch < self.n_epoch:
            selected_operators, selected_gates, max_grads = self.select_operator()
            if len(max_grads) == 0:
                print('\nconvergence criterion has satisfied, break the loop!')
                break
            self.selected_gates += selected_gates
            self.params['t'] = torch.cat((self.params['t'].detach(), torch.zeros(self.Ng))).to(self.device)
            self.results['selected operators'] += selected_operators
            self.results['n_params'].append(len(self.results['selected operators']))
            lr = torch.linalg.vector_norm(torch.Tensor(max_grads)).item() / np.sqrt(self.Ng) * 0.05
            opt = optim.Adam(params=self.params.values(), lr=lr)
            if self.verbose:
                print('learning rate = ', lr)
                print('Find operators')
                print_list(selected_operators)
                print('with max gradients')
                print(max_grads)
                print('')
            while True:
                opt.zero_grad()
                loss, Sz, S_square = self.circuit(mode='train')
                fidelity = self._fidelity()
                loss.backward()
                opt.step()
                self.results['iteration loss'].append(loss.item())
                self.results['Sz'].append(Sz.item())
                self.results['S^2'].append(S_square.item())
                self.results['fidelity'].append(fidelity)
                grad_norm = torch.linalg.vector_norm(self.params['t'].grad).item()
                if self.verbose:
                    print(f"iter: {len(self.results['iteration loss'])} | loss: {loss.item(): 6f} | norm: {grad_norm: 6f} "
                          f"| fidelity: {fidelity: 6f} | Sz: {Sz.item(): 6f} | S^2: {S_square.item(): 6f}")
                if grad_norm < self.threshold2:
                    break
            self.results['epoch loss'].append(self.results['iteration loss'][-1])
            i_epoch += 1
            if self.verbose:
                print('')
            self.save_model()
            if plt:
                self._plot(plt, fig)
        print('total run time: ', time.time() - start_time)

    def _plot(self, plt, fig):
        fig.clf()
        for k, (key, label) in enumerate((('iteration loss', 'iteration'), ('epoch loss', 'epoch'))):
            ax = fig.add_subplot(1, 2, k + 1)
            ys = self.results[key]
            xs = np.arange(len(ys)) + 1
            ax.plot(xs, ys, marker='X', ls='--', label='ADAPT')
            ax.plot(xs, np.full(len(ys), self.ground_state_energy), label='ED')
            ax.set_xlabel(label)
            ax.set_ylabel('energy')
            ax.legend()
            ax.grid()
        ensure_parent(self.img_filepath)
        fig.savefig(self.img_filepath)


if __name__ == '__main__':
    vqe = ADAPT(n_epoch=100, threshold1=1e-2, threshold2=1e-2, x_dimension=2, y_dimension=4, n_electrons=8,
                n_spin_up=4, n_spin_down=4, tunneling=1, coulomb=2, load_model=False)
    vqe.run()
