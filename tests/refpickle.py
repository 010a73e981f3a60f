"""Test helper: write pickles the way the REFERENCE environment would -- class paths of OpenFermion
(``openfermion.ops.operators.{qubit,fermion}_operator``), plain instance dicts ``{'terms': ...}``, and the gate closure
``Trotterize_generator`` pickled by reference from ``__main__`` (the reference drivers run as scripts)."""
import functools
import pickle
import sys
import types


def _fake_openfermion():
    mods = {}
    for path in ("openfermion", "openfermion.ops", "openfermion.ops.operators", "openfermion.ops.operators.symbolic_operator",
                 "openfermion.ops.operators.qubit_operator", "openfermion.ops.operators.fermion_operator"):
        mods[path] = types.ModuleType(path)

    class SymbolicOperator:                                    # no __slots__: state is the instance dict, as upstream
        def __init__(self, terms):
            self.terms = dict(terms)
    SymbolicOperator.__module__ = "openfermion.ops.operators.symbolic_operator"
    SymbolicOperator.__qualname__ = "SymbolicOperator"
    qo = type("QubitOperator", (SymbolicOperator,), {"__module__": "openfermion.ops.operators.qubit_operator"})
    fo = type("FermionOperator", (SymbolicOperator,), {"__module__": "openfermion.ops.operators.fermion_operator"})
    mods["openfermion.ops.operators.symbolic_operator"].SymbolicOperator = SymbolicOperator
    mods["openfermion.ops.operators.qubit_operator"].QubitOperator = qo
    mods["openfermion.ops.operators.fermion_operator"].FermionOperator = fo
    return mods, qo, fo


def dumps_like_reference(obj):
    """Pickle ``obj`` after replacing every fhsim.symbolic operator by an OpenFermion-pathed stand-in and every
    ``partial(Trotterize_generator, generator=...)`` by one whose function lives in ``__main__``.  The stand-in modules are
    removed again before returning, so loading needs fhsim.checkpoint."""
    from fhsim.symbolic import FermionOperator, QubitOperator
    mods, fake_q, fake_f = _fake_openfermion()

    def Trotterize_generator(theta, generator):                # body irrelevant: pickled by reference
        raise RuntimeError("stand-in")
    Trotterize_generator.__module__ = "__main__"
    Trotterize_generator.__qualname__ = "Trotterize_generator"

    def convert(o):
        if isinstance(o, QubitOperator):
            return fake_q(o.terms)
        if isinstance(o, FermionOperator):
            return fake_f(o.terms)
        if isinstance(o, functools.partial) and getattr(o.func, "__name__", "") == "Trotterize_generator":
            return functools.partial(Trotterize_generator, **{k: convert(v) for k, v in o.keywords.items()})
        if isinstance(o, dict):
            return {k: convert(v) for k, v in o.items()}
        if isinstance(o, list):
            return [convert(v) for v in o]
        if isinstance(o, tuple):
            return tuple(convert(v) for v in o)
        return o

    main = sys.modules["__main__"]
    had = getattr(main, "Trotterize_generator", None)
    saved = {k: sys.modules.get(k) for k in mods}
    try:
        sys.modules.update(mods)
        main.Trotterize_generator = Trotterize_generator
        return pickle.dumps(convert(obj))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        if had is None:
            delattr(main, "Trotterize_generator")
        else:
            main.Trotterize_generator = had
