"""Gate recording context.

The reference's ``circuit()`` methods call gate constructors for their queuing side effect
(PennyLane tapes).  Here the same call sites append device-level ops to the :class:`Circuit`
that is currently being recorded; ``theta`` arguments are :class:`Param` handles (parameter index +
multiplier) or plain floats for fixed angles.
"""
from __future__ import annotations

from contextlib import contextmanager
from dataclasses import dataclass

_ACTIVE = []


@dataclass(frozen=True)
class Param:
    """Symbolic angle ``mult * theta[index]``."""
    index: int
    mult: float = 1.0

    def __mul__(self, other):
        return Param(self.index, self.mult * float(other))

    __rmul__ = __mul__

    def __neg__(self):
        return Param(self.index, -self.mult)


@contextmanager
def recording(circuit):
    _ACTIVE.append(circuit)
    try:
        yield circuit
    finally:
        _ACTIVE.pop()


def active_circuit():
    if not _ACTIVE:
        raise RuntimeError("gate called outside a recording(circuit) context")
    return _ACTIVE[-1]
