"""HVA for lattices with a degenerate ground level (drop-in for reference ``models/hva_for_3x3.py``)."""
from __future__ import annotations

import numpy as np

from linalg.exact_diagonalization import jw_get_ground_state_for_3x3

from .hva import HVA as _HVA


class HVA(_HVA):
    ground_state_solver = staticmethod(jw_get_ground_state_for_3x3)

    def _set_ground_state(self, energy, wfs):
        self.ground_state_energy, self.ground_state_wfs = energy, wfs
        self._targets = self.upload_targets(list(wfs))

    def calculate_fidelity(self, ground_state_wfs, state):
        projected = np.zeros_like(state)
        for wf in ground_state_wfs:
            projected += (wf.conj() @ state) * wf
        projected = projected / np.linalg.norm(projected, ord=2)
        return np.abs(state.conj() @ projected) ** 2

    def fidelity_from_overlaps(self, overlaps):
        return float(np.sum(np.abs(overlaps) ** 2))


if __name__ == '__main__':
    vqe = HVA(n_epoch=800, reps=10, lr=1e-2, threshold=1e-2, x_dimension=3, y_dimension=3, n_electrons=9,
              n_spin_up=5, n_spin_down=4, tunneling=1, coulomb=6, periodic=True, spinless=False,
              particle_hole_symmetry=False, load_model=False)
    vqe.run()
