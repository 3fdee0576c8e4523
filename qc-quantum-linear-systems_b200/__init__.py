"""B200-native statevector engine behind the ``QuantumLinearSolver`` surface of
SURFQuantum/qc-quantum-linear-systems (qiskit path only).

Import as ``qls_b200`` (alias of this directory).  Host code is Python; the gate stream runs in
hand-written sm_100a CUDA kernels reached through a C ABI (``include/qls_abi.h``,
``csrc/libqls_b200.so``).  There is no CPU fallback: constructing a ``Statevector`` or calling a
solver without the CUDA library and a GPU raises.
"""

__version__ = "0.1.0"

from .circuit import (  # noqa: F401
    ControlledGate,
    Gate,
    Parameter,
    ParameterVector,
    QuantumCircuit,
    QuantumRegister,
    StatePreparation,
    UCRYGate,
    flatten,
)
from .library import QFT, ExactReciprocal, NumPyMatrix, PhaseEstimation, RealAmplitudes  # noqa: F401
from .hhl import HHL, LinearSolverResult  # noqa: F401


def __getattr__(name):
    # heavy / GPU-facing members are imported lazily so that host-only users (circuit building,
    # lowering, planning) never touch torch or the CUDA library
    import importlib

    lazy = {
        "Statevector": ".statevector",
        "B200Estimator": ".estimator",
        "QuantumLinearSolver": ".quantum_linear_solver",
        "VQLS": ".vqls",
        "lower_circuit": ".lower",
    }
    if name in lazy:
        return getattr(importlib.import_module(lazy[name], __name__), name)
    raise AttributeError(name)
