"""Shared machinery of the drop-in drivers (ADAPT / HVA / IQCC / VQE).

Holds what the reference copy-pastes into every driver file (spin operators, ``Trotterize_generator``,
k-space occupation -- reference ``models/adapt_vqe.py:33-122``, ``models/hva.py:32-115``) plus the
bridge between torch parameters and ``libfhsim``: one C-ABI call per evaluation
(``fh_program_evaluate``) wrapped in a ``torch.autograd.Function`` so ``loss.backward()`` still fills
``.grad`` of the (float32) ``nn.Parameter``s exactly as the reference's QNode does.
"""
from __future__ import annotations

import os
import pickle

from fhsim import checkpoint

import numpy as np
import torch

from fhsim.backend import DevicePool, DeviceTable, State, default_context
from fhsim.circuit import Circuit
from fhsim.recording import Param, active_circuit, recording
from fhsim.symbolic import (FermionOperator, QubitOperator, down_index, fermi_hubbard, givens_decomposition_square,
                            jordan_wigner, number_operator, up_index)
from fhsim.tables import GeneratorPlan
from linalg.exact_diagonalization import get_sparse_operator
from operators.fourier import fourier_transform, fourier_transform_matrix
from operators.tools import get_interacting_term, get_quadratic_term

from .utils import Observable, QubitOperator_to_qmlHamiltonian


# ---------------------------------------------------------------------------------------------
# symbolic helpers
# ---------------------------------------------------------------------------------------------
def get_particle_number_operator(x_dimension, y_dimension, spinless=False):
    n_sites = x_dimension * y_dimension
    total = FermionOperator()
    for site in range(n_sites):
        if spinless:
            total += number_operator(n_sites, site, 1)
        else:
            total += number_operator(2 * n_sites, up_index(site), 1)
            total += number_operator(2 * n_sites, down_index(site), 1)
    return total


def get_total_spin(n_sites, spin_type, *legacy):
    """Total number operator of one spin species.  ``hva.py`` calls this as
    ``get_total_spin(x_dimension, y_dimension, 'spin-up')``; both call shapes are accepted."""
    if legacy:
        n_sites, spin_type = n_sites * spin_type, legacy[0]
    if spin_type not in ('spin-up', 'spin-down'):
        raise ValueError('spin_type must be either spin-up or spin-down')
    index = up_index if spin_type == 'spin-up' else down_index
    total = FermionOperator()
    for site in range(n_sites):
        total += number_operator(2 * n_sites, index(site), 1)
    return total


def get_spin_operators(n_sites, spin_type):
    Sx, Sy, Sz = FermionOperator(), FermionOperator(), FermionOperator()
    for site in range(n_sites):
        u, d = up_index(site), down_index(site)
        Sx += FermionOperator(f'{u}^ {d}', 0.5) + FermionOperator(f'{d}^ {u}', 0.5)
        Sy += FermionOperator(f'{u}^ {d}', -0.5j) - FermionOperator(f'{d}^ {u}', -0.5j)
        Sz += FermionOperator(f'{u}^ {u}', 0.5) - FermionOperator(f'{d}^ {d}', 0.5)
    if spin_type == 'S^2':
        return Sx * Sx + Sy * Sy + Sz * Sz
    return {'Sx': Sx, 'Sy': Sy, 'Sz': Sz}.get(spin_type)


_PLAN_CACHE = {}


def generator_plan(generator, n_qubits) -> GeneratorPlan:
    """Pair/diagonal op plan of a generator, cached on the operator object."""
    key = (id(generator), n_qubits)
    hit = _PLAN_CACHE.get(key)
    if hit is None or hit[0] is not generator:
        hit = (generator, GeneratorPlan(generator, n_qubits))
        _PLAN_CACHE[key] = hit
    return hit[1]


def Trotterize_generator(theta, generator):
    """prod_m exp(-i theta Re(c_m) P_m) over ``generator.terms`` (reference adapt_vqe.py:87-98).

    Appends to the circuit being recorded.  When all strings commute (true for every generator the
    drivers build) the product is emitted as fused pair/diagonal ops -- the same unitary; otherwise one
    rotation per string in dict order, which is the reference's literal product.
    """
    circuit = active_circuit()
    plan = generator_plan(generator, circuit.n)
    if isinstance(theta, Param):
        if theta.mult != 1.0:
            raise ValueError("Trotterize_generator expects an unscaled parameter")
        circuit.generator(plan, param=theta.index)
    else:
        circuit.generator(plan, angle=float(theta))


def print_list(op_list):
    for op in op_list:
        print(str(op).replace('\n', ' '))


def get_non_interacting_ground_state_index(quadratic_hamiltonian, n_qubits, n_spin_up, n_spin_down, verbose=True):
    """Occupied k-orbitals: stable sort of the diagonal k-space energies (reference adapt_vqe.py:104-122)."""
    up = {x: 0 for x in range(0, n_qubits, 2)}
    down = {x: 0 for x in range(1, n_qubits, 2)}
    for term, coeff in quadratic_hamiltonian.terms.items():
        index = term[0][0]
        (up if index % 2 == 0 else down)[index] = coeff
    key_up = {k: complex(v).real for k, v in up.items()}
    key_down = {k: complex(v).real for k, v in down.items()}
    up_idx = sorted(key_up, key=key_up.get)[:n_spin_up]
    down_idx = sorted(key_down, key=key_down.get)[:n_spin_down]
    if verbose:
        print('spin up orbital energies:', up)
        print('spin down orbital energies: ', down)
    return up_idx, down_idx


# ---------------------------------------------------------------------------------------------
# torch <-> libfhsim bridge
# ---------------------------------------------------------------------------------------------
class _Evaluation(torch.autograd.Function):
    """values = evaluator(theta) with d values[0] / d theta supplied by the backend's adjoint sweep."""

    @staticmethod
    def forward(ctx, evaluator, *params):
        flat = [p.detach().reshape(-1).to(device='cpu', dtype=torch.float64) for p in params]
        thetas = torch.cat(flat).numpy() if flat else np.zeros(0)
        values, grads = evaluator(thetas)
        ctx.param_meta = [(p.shape, p.dtype, p.device, p.numel()) for p in params]
        ctx.grads = grads
        outs = tuple(torch.tensor(float(v), dtype=torch.float64) for v in values)
        if len(outs) > 1:
            ctx.mark_non_differentiable(*outs[1:])
        return outs

    @staticmethod
    def backward(ctx, *grad_outputs):
        g0 = grad_outputs[0]
        result = [None]
        offset = 0
        for shape, dtype, device, numel in ctx.param_meta:
            if ctx.grads is None:
                result.append(None)
            else:
                g = torch.from_numpy(np.ascontiguousarray(ctx.grads[offset:offset + numel])).reshape(shape)
                result.append((g * g0.to(torch.float64).cpu()).to(device=device, dtype=dtype))
            offset += numel
        return tuple(result)


def evaluate_with_grad(evaluator, params):
    """-> tuple of 0-d float64 tensors; the first carries a grad_fn w.r.t. ``params``."""
    return _Evaluation.apply(evaluator, *params)


class DeviceObservable:
    """Observable + its uploaded table, re-uploaded when the operator object changes."""

    def __init__(self, ctx, n_qubits):
        self.ctx, self.n = ctx, n_qubits
        self._source = None
        self._table = None

    def get(self, observable: Observable) -> DeviceTable:
        if self._source is not observable:
            if self._table is not None:
                self._table.close()
            self._table = DeviceTable(self.ctx, observable.table(self.n))
            self._source = observable
        return self._table


# ---------------------------------------------------------------------------------------------
# Hubbard lattice set-up shared by ADAPT and HVA
# ---------------------------------------------------------------------------------------------
class HubbardProblem:
    """Everything the reference's ADAPT.__init__ / HVA.__init__ derive from the lattice arguments."""

    def setup_lattice(self, x_dimension, y_dimension, n_electrons, n_spin_up, n_spin_down, tunneling, coulomb,
                      periodic, spinless, particle_hole_symmetry, verbose=True):
        self.x_dimension, self.y_dimension = x_dimension, y_dimension
        self.n_sites = x_dimension * y_dimension
        self.n_qubits = 2 * self.n_sites
        self.n_electrons, self.n_spin_up, self.n_spin_down = n_electrons, n_spin_up, n_spin_down
        self.device = torch.device('cpu')      # parameters stay on the host; the state lives in libfhsim
        self.fermionHamiltonian = fermi_hubbard(x_dimension, y_dimension, tunneling, coulomb, periodic=periodic,
                                                spinless=spinless, particle_hole_symmetry=particle_hole_symmetry)
        self.qubitHamiltonian = jordan_wigner(self.fermionHamiltonian)
        self.quadratic_term = get_quadratic_term(self.fermionHamiltonian)
        self.interacting_term = get_interacting_term(self.fermionHamiltonian)
        self.qmlHamiltonian = QubitOperator_to_qmlHamiltonian(self.fermionHamiltonian)
        self.FT_transformation_matrix = fourier_transform_matrix(x_dimension, y_dimension)
        self.decomposition, self.diagonal = givens_decomposition_square(self.FT_transformation_matrix)
        self.circuit_description = list(reversed(self.decomposition))
        self.k_quadratic_term = fourier_transform(self.quadratic_term, x_dimension, y_dimension)
        self.spin_up_indices, self.spin_down_indices = get_non_interacting_ground_state_index(
            self.k_quadratic_term, self.n_qubits, n_spin_up, n_spin_down, verbose=verbose)
        if verbose:
            print('spin up indices: ', self.spin_up_indices, '  ', 'spin down indices: ', self.spin_down_indices, '\n')
        # W is compiled from the tensor structure of the FT matrix (same unitary as the reference network)
        self.separable_basis_change = True
        self._ctx = default_context()
        self._tables = {}

    def basis_index(self):
        n = self.n_qubits
        return sum(1 << (n - 1 - q) for q in self.spin_up_indices + self.spin_down_indices)

    def device_table(self, name, observable):
        slot = self._tables.get(name)
        if slot is None:
            slot = self._tables[name] = DeviceObservable(self._ctx, self.n_qubits)
        return slot.get(observable)

    def append_basis_change(self, circuit: Circuit):
        """W: k-space -> real space (reference adapt_vqe.py:343-354); returns the global phase by which the
        compiled network differs from the reference's gate-by-gate network."""
        if self.separable_basis_change:
            fast = circuit.basis_change_separable(self.x_dimension, self.y_dimension)
            return fast - Circuit.basis_change_vacuum_phase(self.diagonal, self.decomposition)
        circuit.basis_change(self.diagonal, self.circuit_description)
        return 0.0

    # -- exact diagonalisation cache (reference adapt_vqe.py:221-247) --------------------------------
    def load_or_compute_ground_state(self, solver):
        path = self.wf_filepath
        if os.path.exists(path):
            with open(path, 'rb') as file:
                cached = checkpoint.load(file)
            return cached['energy'], cached['wave function']
        energy, wf = solver(sparse_operator=get_sparse_operator(self.fermionHamiltonian, self.n_qubits),
                            particle_number=self.n_electrons, spin_up=self.n_spin_up, spin_down=self.n_spin_down)
        os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
        with open(path, 'wb') as file:
            pickle.dump({'energy': energy, 'wave function': wf}, file)
        return energy, wf

    def upload_targets(self, vectors):
        return [State.from_numpy(self._ctx, v) for v in vectors]


def ensure_parent(path):
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)


def try_pyplot():
    try:
        import matplotlib
        matplotlib.use('Agg')
        import matplotlib.pyplot as plt
        return plt
    except Exception:
        return None


__all__ = [
    'QubitOperator', 'FermionOperator', 'Param', 'recording', 'Circuit', 'DevicePool', 'DeviceTable', 'State',
    'GeneratorPlan', 'Trotterize_generator', 'print_list', 'get_particle_number_operator', 'get_total_spin',
    'get_spin_operators', 'get_non_interacting_ground_state_index', 'evaluate_with_grad', 'HubbardProblem',
    'generator_plan', 'ensure_parent', 'try_pyplot', 'QubitOperator_to_qmlHamiltonian',
]
