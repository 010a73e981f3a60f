"""ADAPT-VQE for lattices with a degenerate ground level (drop-in for reference
``models/adapt_vqe_for_3x3.py``: 3x3, 9 electrons, 4-fold degenerate ground space).

Differs from :mod:`models.adapt_vqe` exactly where the reference's twin differs: the exact
diagonalisation keeps the four lowest states (``jw_get_ground_state_for_3x3``), the fidelity is the
weight of the state inside that subspace (``calculate_fidelity``, reference :361-368), and the
shipped configuration resumes from a checkpoint.
"""
from __future__ import annotations

import numpy as np

from linalg.exact_diagonalization import jw_get_ground_state_for_3x3

from .adapt_vqe import (ADAPT as _ADAPT, Trotterize_generator, get_non_interacting_ground_state_index,  # noqa: F401
                        get_particle_number_operator, get_spin_operators, get_total_spin, print_list)


class ADAPT(_ADAPT):
    ground_state_solver = staticmethod(jw_get_ground_state_for_3x3)

    def _set_ground_state(self, energy, wfs):
        self.ground_state_energy, self.ground_state_wfs = energy, wfs
        self._targets = self.upload_targets(list(wfs))

    def calculate_fidelity(self, ground_state_wfs, state):
        """|<state| P state / ||P state||>|^2 with P the projector on span(ground_state_wfs)."""
        projected = np.zeros_like(state)
        for wf in ground_state_wfs:
            projected += (wf.conj() @ state) * wf
        projected = projected / np.linalg.norm(projected, ord=2)
        return np.abs(state.conj() @ projected) ** 2

    def fidelity_from_overlaps(self, overlaps):
        # the target states are orthonormal, so the projected fidelity is the weight in the subspace;
        # the overlaps <gs_k|psi> come out of the same device evaluation as the energy
        return float(np.sum(np.abs(overlaps) ** 2))


if __name__ == '__main__':
    vqe = ADAPT(n_epoch=100, threshold1=1e-2, threshold2=1e-2, x_dimension=3, y_dimension=3, n_electrons=9,
                n_spin_up=5, n_spin_down=4, tunneling=1, coulomb=6, load_model=True)
    vqe.run()
