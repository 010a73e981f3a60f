"""Checkpoints written by the reference (class paths of its environment) load through fhsim.checkpoint
(reference models/adapt_vqe.py:269-295; the shipped 3x3 configuration is load_model=True, adapt_vqe_for_3x3.py:482)."""
import pickle
from functools import partial

import numpy as np
import pytest
import torch
import torch.nn as nn

from fhsim import checkpoint
from fhsim.symbolic import FermionOperator, QubitOperator, jordan_wigner
from models.common import Trotterize_generator
from operators.pool import hubbard_interaction_pool_simplified
from refpickle import dumps_like_reference


def test_reference_pathed_checkpoint_loads_onto_this_builds_classes():
    pool = hubbard_interaction_pool_simplified(2, 2)
    qpool = [jordan_wigner(g) for g in pool]
    params = nn.ParameterDict({'e': nn.Parameter(torch.zeros(len(pool))),
                               't': nn.Parameter(torch.tensor([0.25, -0.5, 0.125]))})
    model = {'params': params, 'circuit': [partial(Trotterize_generator, generator=qpool[k]) for k in (1, 2, 4)]}
    results = {'epoch loss': [-1.0, -1.5], 'iteration loss': [0.0, -0.7, -1.5], 'Sz': [0.0], 'S^2': [0.0],
               'fidelity': [0.5], 'n_params': [3], 'selected operators': [pool[k] for k in (1, 2, 4)]}
    ed_cache = {'energy': -2.1027484835, 'wave function': np.arange(8, dtype=complex)}
    blobs = [dumps_like_reference(o) for o in (model, results, ed_cache)]
    assert b"openfermion.ops.operators.qubit_operator" in blobs[0] and b"__main__" in blobs[0]
    assert b"openfermion.ops.operators.fermion_operator" in blobs[1]
    with pytest.raises((ModuleNotFoundError, AttributeError)):
        pickle.loads(blobs[0])                                    # the plain loader cannot resolve the reference's paths
    m = checkpoint.loads(blobs[0])
    assert torch.equal(m['params']['t'], params['t']) and isinstance(m['params'], nn.ParameterDict)
    assert [type(g.keywords['generator']) for g in m['circuit']] == [QubitOperator] * 3
    assert all(g.func is Trotterize_generator for g in m['circuit'])
    assert [g.keywords['generator'] == qpool[k] for g, k in zip(m['circuit'], (1, 2, 4))] == [True] * 3
    assert [list(g.keywords['generator'].terms.items()) for g in m['circuit']] == [list(qpool[k].terms.items()) for k in (1, 2, 4)]
    r = checkpoint.loads(blobs[1])
    assert all(isinstance(op, FermionOperator) for op in r['selected operators'])
    assert r['selected operators'] == [pool[k] for k in (1, 2, 4)] and r['epoch loss'] == [-1.0, -1.5]
    c = checkpoint.loads(blobs[2])
    assert c['energy'] == -2.1027484835 and np.array_equal(c['wave function'], np.arange(8, dtype=complex))
    # operators of this build still round-trip through plain pickle (slots state) and through the shim
    again = pickle.loads(pickle.dumps(qpool[3]))
    assert again == qpool[3] and list(again.terms.items()) == list(qpool[3].terms.items())
    assert checkpoint.loads(pickle.dumps(pool[5])) == pool[5]


def test_unsupported_reference_objects_fail_loudly():
    import types, sys
    mod = types.ModuleType("pennylane")
    cls = type("QNode", (), {"__module__": "pennylane"})
    mod.QNode = cls
    sys.modules["pennylane"] = mod
    try:
        blob = pickle.dumps(cls())
    finally:
        del sys.modules["pennylane"]
    with pytest.raises(pickle.UnpicklingError):
        checkpoint.loads(blob)
