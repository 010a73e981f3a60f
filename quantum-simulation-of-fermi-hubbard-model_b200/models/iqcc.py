"""iQCC for molecules (reference ``models/iqcc.py``).

The reference builds its Hamiltonian from an OpenFermion ``MolecularData`` (pyscf integrals).  That
front-end is out of scope (pyscf is not available and is unrelated to the Hubbard hot path); the
algorithm itself is identical to :class:`models.iqcc_hubbard.IQCC`, which this class re-uses for any
object offering ``get_molecular_hamiltonian()``.
"""
from .iqcc_hubbard import IQCC as _IQCC


class IQCC(_IQCC):
    def __init__(self, molecule, n_epoch: int, lr: float, threshold: float, **kw):
        hamiltonian = molecule.get_molecular_hamiltonian() if hasattr(molecule, 'get_molecular_hamiltonian') else molecule
        super().__init__(hamiltonian, n_epoch, lr, threshold, **kw)
        self.molecule = molecule
