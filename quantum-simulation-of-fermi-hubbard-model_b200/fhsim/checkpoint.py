"""Loading checkpoints / results / ED caches written by the REFERENCE code base.

The reference pickles (``models/adapt_vqe.py:269-280``, ``hva.py``, ED cache ``adapt_vqe.py:221-247``)

* ``{'params': nn.ParameterDict, 'circuit': [functools.partial(Trotterize_generator, generator=QubitOperator)]}``
* the results dict (``'selected operators'`` is a list of ``FermionOperator``)
* ``{'energy': float, 'wave function': ndarray}``

so the byte stream names classes by their import path in the *reference's* environment:
``openfermion.ops.operators.qubit_operator.QubitOperator`` (and ``fermion_operator.FermionOperator``; older OpenFermion
releases: ``openfermion.ops._qubit_operator`` ...), and the gate closure as ``models.adapt_vqe.Trotterize_generator`` --
or ``__main__.Trotterize_generator`` when the driver was started as a script, which is how every shipped ``__main__``
block runs it.  OpenFermion and PennyLane are not part of this build, so a plain ``pickle.load`` raises
``ModuleNotFoundError``.  :class:`ReferenceUnpickler` resolves those names onto this package's own classes
(``fhsim.symbolic`` operators, ``models.common.Trotterize_generator``); everything else (torch, numpy, builtins) is
resolved normally.  OpenFermion operators pickle as ``cls.__new__`` + ``{'terms': {...}}``, which the symbolic classes
accept through ``__setstate__``.
"""
from __future__ import annotations

import importlib
import io
import pickle

_SYMBOLIC = ("FermionOperator", "QubitOperator", "SymbolicOperator")
# functions the reference defines at module level in its drivers and pickles by reference inside functools.partial
_DRIVER_FUNCTIONS = ("Trotterize_generator", "PauliStringRotation")
_DRIVER_MODULES = ("__main__", "models.adapt_vqe", "models.adapt_vqe_for_3x3", "models.hva", "models.hva_for_3x3",
                   "models.iqcc_hubbard", "models.iqcc", "models.vqe_hea", "models.utils")


class ReferenceUnpickler(pickle.Unpickler):
    """``pickle.Unpickler`` whose ``find_class`` maps the reference environment's class paths onto this build."""

    def find_class(self, module, name):
        root = module.split(".")[0]
        if root == "openfermion":
            if name in _SYMBOLIC:
                import fhsim.symbolic as sym
                return getattr(sym, name)
            raise pickle.UnpicklingError(f"reference pickle needs openfermion.{name}, which this build does not provide")
        if root == "pennylane":
            raise pickle.UnpicklingError(f"reference pickle embeds the PennyLane object {module}.{name}; only parameters, "
                                         "gate closures, operators and result lists are supported")
        if name in _DRIVER_FUNCTIONS and module in _DRIVER_MODULES:
            target = "models.utils" if name == "PauliStringRotation" else "models.common"
            return getattr(importlib.import_module(target), name)
        return super().find_class(module, name)


def load(file):
    """``pickle.load`` for files written by the reference or by this build."""
    return ReferenceUnpickler(file).load()


def loads(data: bytes):
    return ReferenceUnpickler(io.BytesIO(data)).load()


def load_path(path):
    with open(path, "rb") as f:
        return load(f)
