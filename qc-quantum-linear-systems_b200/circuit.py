"""Circuit IR of the B200 statevector engine.

A deliberately small stand-in for the part of ``qiskit.QuantumCircuit`` that the reference's
qiskit path touches (``hhl_qiskit_implementation.py:35-77``, ``vqls_qiskit_implementation.py:38-89``
in the reference tree): registers, gate append helpers, ``compose`` / ``inverse`` / ``to_gate`` /
``control`` / ``power`` / ``decompose`` / ``depth`` / ``width`` and parameter binding.

Conventions are qiskit's:

* amplitude index bit ``q`` is qubit ``q`` (little endian, qubit 0 = LSB);
* a gate applied to ``qargs = [q0, q1, ...]`` uses ``q0`` as the LSB of its ``2**k`` matrix index;
* controlled gates take ``qargs = [controls..., targets...]``;
* every circuit carries a ``global_phase`` that the simulation multiplies in.

Nothing here touches the GPU; the lowering pass (``lower.py``) turns the flattened leaf stream into
fused blocks for the CUDA kernels and never mutates the circuit it is given.
"""
from __future__ import annotations

import cmath
import math
from dataclasses import dataclass
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

# --------------------------------------------------------------------------------------
# Parameters (just enough for RealAmplitudes-style ansaetze: a*theta + b)
# --------------------------------------------------------------------------------------


class ParameterExpression:
    """Affine expression ``coeff * parameter + offset`` of one free parameter."""

    __slots__ = ("parameter", "coeff", "offset")

    def __init__(self, parameter: "Parameter", coeff: float = 1.0, offset: float = 0.0):
        self.parameter = parameter
        self.coeff = float(coeff)
        self.offset = float(offset)

    def bind(self, values: Dict["Parameter", float]) -> Union[float, "ParameterExpression"]:
        if self.parameter in values:
            return self.coeff * float(values[self.parameter]) + self.offset
        return self

    def __neg__(self) -> "ParameterExpression":
        return ParameterExpression(self.parameter, -self.coeff, -self.offset)

    def __mul__(self, other: float) -> "ParameterExpression":
        return ParameterExpression(self.parameter, self.coeff * other, self.offset * other)

    __rmul__ = __mul__

    def __add__(self, other: float) -> "ParameterExpression":
        return ParameterExpression(self.parameter, self.coeff, self.offset + other)

    __radd__ = __add__

    def __sub__(self, other: float) -> "ParameterExpression":
        return self + (-other)

    def __repr__(self) -> str:
        return f"{self.coeff}*{self.parameter.name}+{self.offset}"


class Parameter(ParameterExpression):
    """A named free parameter (identity-hashed, like qiskit's)."""

    __slots__ = ("name",)

    def __init__(self, name: str):
        self.name = name
        super().__init__(self, 1.0, 0.0)

    def __hash__(self) -> int:
        return id(self)

    def __eq__(self, other: object) -> bool:
        return self is other

    def __repr__(self) -> str:
        return f"Parameter({self.name})"


class ParameterVector:
    def __init__(self, name: str, length: int):
        self.name = name
        self.params = [Parameter(f"{name}[{i}]") for i in range(length)]

    def __getitem__(self, i: int) -> Parameter:
        return self.params[i]

    def __len__(self) -> int:
        return len(self.params)

    def __iter__(self):
        return iter(self.params)


def _is_free(p: Any) -> bool:
    return isinstance(p, ParameterExpression)


# --------------------------------------------------------------------------------------
# Registers
# --------------------------------------------------------------------------------------


@dataclass(frozen=True)
class QuantumRegister:
    size: int
    name: str = "q"


# --------------------------------------------------------------------------------------
# Gates
# --------------------------------------------------------------------------------------

_SQ2 = 1.0 / math.sqrt(2.0)


def _u_matrix(theta: float, phi: float, lam: float) -> np.ndarray:
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array(
        [[c, -cmath.exp(1j * lam) * s], [cmath.exp(1j * phi) * s, cmath.exp(1j * (phi + lam)) * c]],
        dtype=complex,
    )


def _std_matrix(name: str, params: Sequence[float]) -> Optional[np.ndarray]:
    """Matrices of the parameter-free / numeric standard gates."""
    if name == "id":
        return np.eye(2, dtype=complex)
    if name == "h":
        return np.array([[_SQ2, _SQ2], [_SQ2, -_SQ2]], dtype=complex)
    if name == "x":
        return np.array([[0, 1], [1, 0]], dtype=complex)
    if name == "y":
        return np.array([[0, -1j], [1j, 0]], dtype=complex)
    if name == "z":
        return np.diag([1, -1]).astype(complex)
    if name == "s":
        return np.diag([1, 1j]).astype(complex)
    if name == "sdg":
        return np.diag([1, -1j]).astype(complex)
    if name == "t":
        return np.diag([1, cmath.exp(1j * math.pi / 4)]).astype(complex)
    if name == "tdg":
        return np.diag([1, cmath.exp(-1j * math.pi / 4)]).astype(complex)
    if name == "sx":
        return 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]], dtype=complex)
    if name == "p":
        return np.diag([1, cmath.exp(1j * params[0])]).astype(complex)
    if name == "rx":
        c, s = math.cos(params[0] / 2), math.sin(params[0] / 2)
        return np.array([[c, -1j * s], [-1j * s, c]], dtype=complex)
    if name == "ry":
        c, s = math.cos(params[0] / 2), math.sin(params[0] / 2)
        return np.array([[c, -s], [s, c]], dtype=complex)
    if name == "rz":
        return np.diag([cmath.exp(-0.5j * params[0]), cmath.exp(0.5j * params[0])]).astype(complex)
    if name == "u":
        return _u_matrix(*params)
    if name == "swap":
        return np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=complex)
    return None


_DIAGONAL_NAMES = {"id", "z", "s", "sdg", "t", "tdg", "p", "rz", "diagonal"}
_INVERSE_NAME = {"s": "sdg", "sdg": "s", "t": "tdg", "tdg": "t"}
_SELF_INVERSE = {"id", "h", "x", "y", "z", "swap"}


class Gate:
    """A unitary instruction on ``num_qubits`` qubits.

    Exactly one of the following describes its action:

    * a matrix (``to_matrix()`` returns it) -> a *leaf* for the simulator, like a qiskit gate with
      ``to_matrix``/``__array__``;
    * a ``definition`` circuit -> the simulator multiplies in ``definition.global_phase`` and
      recurses, like ``Statevector._evolve_instruction`` does for gates without a matrix.
    """

    def __init__(
        self,
        name: str,
        num_qubits: int,
        params: Sequence[Any] = (),
        matrix: Optional[np.ndarray] = None,
        definition: Optional["QuantumCircuit"] = None,
        label: Optional[str] = None,
    ):
        self.name = name
        self.num_qubits = int(num_qubits)
        self.params = list(params)
        self._matrix = matrix
        self.definition = definition
        self.label = label

    # -- description -----------------------------------------------------------------
    @property
    def is_parameterized(self) -> bool:
        if any(_is_free(p) for p in self.params):
            return True
        return self.definition is not None and self._matrix is None and self.definition.num_parameters > 0

    def to_matrix(self) -> Optional[np.ndarray]:
        if self._matrix is not None:
            return self._matrix
        if self.definition is None and not any(_is_free(p) for p in self.params):
            m = _std_matrix(self.name, self.params)
            if m is not None:
                self._matrix = m
            return m
        return None

    @property
    def is_diagonal(self) -> bool:
        return self.name in _DIAGONAL_NAMES

    # -- algebra ---------------------------------------------------------------------
    def bind(self, values: Dict[Parameter, float]) -> "Gate":
        if not self.is_parameterized:
            return self
        if self.definition is not None and self._matrix is None and not self.params:
            return Gate(self.name, self.num_qubits, [], None, self.definition.assign_parameters(values), self.label)
        new_params = [p.bind(values) if _is_free(p) else p for p in self.params]
        return Gate(self.name, self.num_qubits, new_params, None, None, self.label)

    def inverse(self) -> "Gate":
        if self.definition is not None and self._matrix is None:
            return Gate(self.name + "_dg", self.num_qubits, [], None, self.definition.inverse(), self.label)
        if self.name in _SELF_INVERSE:
            return self
        if self.name in _INVERSE_NAME:
            return Gate(_INVERSE_NAME[self.name], 1)
        if self.name in ("p", "rx", "ry", "rz"):
            return Gate(self.name, 1, [-self.params[0]])
        if self.name == "u":
            th, ph, lam = self.params
            return Gate("u", 1, [-th, -lam, -ph])
        m = self.to_matrix()
        if m is None:
            raise ValueError(f"cannot invert gate {self.name!r} with unbound parameters")
        name = self.name[:-3] if self.name.endswith("_dg") else self.name + "_dg"
        if self.name in ("unitary", "diagonal"):
            name = self.name
        return Gate(name, self.num_qubits, [], m.conj().T.copy(), None, self.label)

    def control(self, num_ctrl_qubits: int = 1, ctrl_state: Optional[int] = None) -> "ControlledGate":
        return ControlledGate(self, num_ctrl_qubits, ctrl_state)

    def __repr__(self) -> str:
        return f"Gate({self.name}, {self.num_qubits}q, params={self.params})"


class ControlledGate(Gate):
    """``base_gate`` controlled on ``num_ctrl_qubits`` qubits (``qargs = controls + targets``).

    Kept symbolic (base + controls) rather than as a ``2**(c+k)`` matrix so that the engine only
    touches the ``2**(n-c)`` amplitudes with the controls satisfied.  ``to_matrix`` still produces
    the full little-endian matrix for the oracle and for small-gate fusion.
    """

    def __init__(self, base_gate: Gate, num_ctrl_qubits: int = 1, ctrl_state: Optional[int] = None):
        if isinstance(base_gate, ControlledGate):
            # flatten nested control: new controls go first (qiskit order)
            inner = base_gate
            ctrl_state_inner = inner.ctrl_state
            new_state = (2**num_ctrl_qubits - 1) if ctrl_state is None else int(ctrl_state)
            ctrl_state = new_state | (ctrl_state_inner << num_ctrl_qubits)
            num_ctrl_qubits = num_ctrl_qubits + inner.num_ctrl_qubits
            base_gate = inner.base_gate
        self.base_gate = base_gate
        self.num_ctrl_qubits = int(num_ctrl_qubits)
        self.ctrl_state = (2**num_ctrl_qubits - 1) if ctrl_state is None else int(ctrl_state)
        super().__init__(
            "c" * num_ctrl_qubits + base_gate.name if num_ctrl_qubits <= 2 else f"mc{base_gate.name}",
            base_gate.num_qubits + num_ctrl_qubits,
            base_gate.params,
            None,
            None,
            base_gate.label,
        )

    @property
    def is_parameterized(self) -> bool:
        return self.base_gate.is_parameterized

    @property
    def is_diagonal(self) -> bool:
        return self.base_gate.is_diagonal

    def to_matrix(self) -> Optional[np.ndarray]:
        if self._matrix is None:
            base = self.base_gate.to_matrix()
            if base is None or self.num_qubits > 12:
                return None
            c = self.num_ctrl_qubits
            dim = 2**self.num_qubits
            full = np.eye(dim, dtype=complex)
            # controls are the low bits of the matrix index (qargs = controls + targets)
            sel = [i for i in range(dim) if (i & (2**c - 1)) == self.ctrl_state]
            full[np.ix_(sel, sel)] = base
            self._matrix = full
        return self._matrix

    def bind(self, values: Dict[Parameter, float]) -> "Gate":
        if not self.is_parameterized:
            return self
        return ControlledGate(self.base_gate.bind(values), self.num_ctrl_qubits, self.ctrl_state)

    def inverse(self) -> "Gate":
        return ControlledGate(self.base_gate.inverse(), self.num_ctrl_qubits, self.ctrl_state)


class UCRYGate(Gate):
    """Uniformly controlled RY: ``qargs = [target, control_0, control_1, ...]`` and the target gets
    ``RY(angles[v])`` with ``v = sum_k control_k * 2**k`` (qiskit ``UCRYGate`` semantics,
    SURVEY.md Appendix A)."""

    def __init__(self, angles: Sequence[float]):
        angles = np.asarray(angles, dtype=float)
        nc = int(round(math.log2(len(angles))))
        if 2**nc != len(angles):
            raise ValueError("UCRY needs 2**k angles")
        super().__init__("ucry", nc + 1, [angles])
        self.angles = angles

    def to_matrix(self) -> Optional[np.ndarray]:
        if self._matrix is None:
            if self.num_qubits > 12:
                return None
            dim = 2**self.num_qubits
            m = np.zeros((dim, dim), dtype=complex)
            for v, th in enumerate(self.angles):
                c, s = math.cos(th / 2), math.sin(th / 2)
                i0, i1 = 2 * v, 2 * v + 1  # target is bit 0 of the gate index
                m[i0, i0], m[i0, i1], m[i1, i0], m[i1, i1] = c, -s, s, c
            self._matrix = m
        return self._matrix

    def inverse(self) -> "Gate":
        return UCRYGate(-self.angles)


class DiagonalGate(Gate):
    """Diagonal unitary given by its ``2**k`` diagonal entries (qiskit ``DiagonalGate``)."""

    def __init__(self, diag: Sequence[complex]):
        d = np.asarray(diag, dtype=complex).reshape(-1)
        nq = int(round(math.log2(len(d))))
        if 2**nq != len(d):
            raise ValueError("diagonal needs 2**k entries")
        super().__init__("diagonal", nq, [d])
        self.diag = d

    def to_matrix(self) -> Optional[np.ndarray]:
        if self._matrix is None:
            if self.num_qubits > 12:
                return None
            self._matrix = np.diag(self.diag)
        return self._matrix

    def inverse(self) -> "Gate":
        return DiagonalGate(self.diag.conj())


class StatePreparation(Gate):
    """Isometry ``|0...0> -> sum_i amps[i] |i>`` (qiskit ``isometry`` / ``prepare_state``).

    As a matrix it is *a* unitary whose first column is ``amps`` (Householder completion); the
    engine short-cuts it to a direct initialisation when it is the first gate on fresh qubits
    (SURVEY.md section 8a row a7)."""

    def __init__(self, amplitudes: Sequence[complex]):
        amps = np.asarray(amplitudes, dtype=complex).reshape(-1)
        nq = int(round(math.log2(len(amps))))
        if 2**nq != len(amps):
            raise ValueError("state preparation needs 2**k amplitudes")
        nrm = np.linalg.norm(amps)
        if abs(nrm - 1.0) > 1e-8:
            raise ValueError("state preparation needs a normalised vector")
        super().__init__("state_preparation", nq, [amps])
        self.amplitudes = amps

    def to_matrix(self) -> Optional[np.ndarray]:
        if self._matrix is None:
            if self.num_qubits > 12:
                return None
            a = self.amplitudes
            dim = len(a)
            e0 = np.zeros(dim, dtype=complex)
            e0[0] = 1.0
            # unitary Householder-type reflection mapping e0 -> a
            phase = a[0] / abs(a[0]) if abs(a[0]) > 1e-300 else 1.0
            w = e0 - a / phase
            nw = np.linalg.norm(w)
            if nw < 1e-15:
                m = np.eye(dim, dtype=complex) * phase
            else:
                w = w / nw
                m = phase * (np.eye(dim, dtype=complex) - 2.0 * np.outer(w, w.conj()))
            self._matrix = m
        return self._matrix


@dataclass
class Instruction:
    operation: Gate
    qubits: Tuple[int, ...]


# --------------------------------------------------------------------------------------
# Circuit
# --------------------------------------------------------------------------------------


class QuantumCircuit:
    """Ordered list of gate applications on ``num_qubits`` qubits plus a global phase."""

    def __init__(self, *regs: Union[int, QuantumRegister], name: str = "circuit", global_phase: float = 0.0):
        self.name = name
        self.global_phase = float(global_phase)
        self.qregs: List[QuantumRegister] = []
        self.data: List[Instruction] = []
        self.num_clbits = 0
        for r in regs:
            if isinstance(r, QuantumRegister):
                self.qregs.append(r)
            else:
                self.qregs.append(QuantumRegister(int(r), "q"))
        self.num_qubits = sum(r.size for r in self.qregs)

    # -- registers --------------------------------------------------------------------
    def register_qubits(self, index: int) -> List[int]:
        off = sum(r.size for r in self.qregs[:index])
        return list(range(off, off + self.qregs[index].size))

    @property
    def qubits(self) -> List[int]:
        return list(range(self.num_qubits))

    # -- building ---------------------------------------------------------------------
    def append(self, gate: Union[Gate, "QuantumCircuit"], qargs: Sequence[int]) -> "QuantumCircuit":
        if isinstance(gate, QuantumCircuit):
            gate = gate.to_gate()
        qargs = tuple(int(q) for q in qargs)
        if len(qargs) != gate.num_qubits:
            raise ValueError(f"gate {gate.name} needs {gate.num_qubits} qubits, got {len(qargs)}")
        if len(set(qargs)) != len(qargs):
            raise ValueError("duplicate qubit arguments")
        if any(q < 0 or q >= self.num_qubits for q in qargs):
            raise ValueError("qubit index out of range")
        self.data.append(Instruction(gate, qargs))
        return self

    def _g(self, name: str, qubits: Sequence[int], params: Sequence[Any] = ()) -> "QuantumCircuit":
        return self.append(Gate(name, len(qubits), params), qubits)

    def id(self, q): return self._g("id", [q])
    def h(self, q): return self._g("h", [q])
    def x(self, q): return self._g("x", [q])
    def y(self, q): return self._g("y", [q])
    def z(self, q): return self._g("z", [q])
    def s(self, q): return self._g("s", [q])
    def sdg(self, q): return self._g("sdg", [q])
    def t(self, q): return self._g("t", [q])
    def tdg(self, q): return self._g("tdg", [q])
    def sx(self, q): return self._g("sx", [q])
    def p(self, lam, q): return self._g("p", [q], [lam])
    def rx(self, th, q): return self._g("rx", [q], [th])
    def ry(self, th, q): return self._g("ry", [q], [th])
    def rz(self, th, q): return self._g("rz", [q], [th])
    def u(self, th, ph, lam, q): return self._g("u", [q], [th, ph, lam])
    def swap(self, a, b): return self._g("swap", [a, b])
    def cx(self, c, t): return self.append(Gate("x", 1).control(1), [c, t])
    def cy(self, c, t): return self.append(Gate("y", 1).control(1), [c, t])
    def cz(self, c, t): return self.append(Gate("z", 1).control(1), [c, t])
    def ch(self, c, t): return self.append(Gate("h", 1).control(1), [c, t])
    def cp(self, lam, c, t): return self.append(Gate("p", 1, [lam]).control(1), [c, t])
    def crx(self, th, c, t): return self.append(Gate("rx", 1, [th]).control(1), [c, t])
    def cry(self, th, c, t): return self.append(Gate("ry", 1, [th]).control(1), [c, t])
    def crz(self, th, c, t): return self.append(Gate("rz", 1, [th]).control(1), [c, t])
    def ccx(self, c1, c2, t): return self.append(Gate("x", 1).control(2), [c1, c2, t])
    def mcx(self, controls, t): return self.append(Gate("x", 1).control(len(controls)), list(controls) + [t])

    def unitary(self, matrix: np.ndarray, qubits: Sequence[int], label: Optional[str] = None) -> "QuantumCircuit":
        m = np.asarray(matrix, dtype=complex)
        qubits = [qubits] if isinstance(qubits, (int, np.integer)) else list(qubits)
        if m.shape != (2 ** len(qubits), 2 ** len(qubits)):
            raise ValueError("unitary shape does not match qubit count")
        if m.shape[0] <= 64 and not np.allclose(m.conj().T @ m, np.eye(m.shape[0]), atol=1e-8):
            raise ValueError("matrix is not unitary")
        return self.append(Gate("unitary", len(qubits), [], m, None, label), qubits)

    def diagonal(self, diag: Sequence[complex], qubits: Sequence[int]) -> "QuantumCircuit":
        return self.append(DiagonalGate(diag), qubits)

    def ucry(self, angles: Sequence[float], controls: Sequence[int], target: int) -> "QuantumCircuit":
        return self.append(UCRYGate(angles), [target] + list(controls))

    def prepare_state(self, amplitudes: Sequence[complex], qubits: Optional[Sequence[int]] = None) -> "QuantumCircuit":
        g = StatePreparation(amplitudes)
        return self.append(g, list(range(g.num_qubits)) if qubits is None else qubits)

    def barrier(self, *_: Any) -> "QuantumCircuit":
        return self

    # -- composition ------------------------------------------------------------------
    def copy(self) -> "QuantumCircuit":
        qc = QuantumCircuit(*self.qregs, name=self.name, global_phase=self.global_phase)
        qc.data = list(self.data)
        return qc

    def compose(self, other: "QuantumCircuit", qubits: Optional[Sequence[int]] = None, inplace: bool = False) -> "QuantumCircuit":
        dest = self if inplace else self.copy()
        qmap = list(range(other.num_qubits)) if qubits is None else [int(q) for q in qubits]
        if len(qmap) != other.num_qubits:
            raise ValueError("compose: qubit map has the wrong length")
        dest.global_phase += other.global_phase
        for ins in other.data:
            dest.append(ins.operation, [qmap[q] for q in ins.qubits])
        return dest

    def inverse(self) -> "QuantumCircuit":
        qc = QuantumCircuit(*self.qregs, name=self.name + "_dg", global_phase=-self.global_phase)
        for ins in reversed(self.data):
            qc.data.append(Instruction(ins.operation.inverse(), ins.qubits))
        return qc

    def reverse_bits(self) -> "QuantumCircuit":
        n = self.num_qubits
        qc = QuantumCircuit(*self.qregs, name=self.name, global_phase=self.global_phase)
        for ins in self.data:
            qc.data.append(Instruction(ins.operation, tuple(n - 1 - q for q in ins.qubits)))
        return qc

    def to_gate(self, label: Optional[str] = None) -> Gate:
        return Gate(self.name, self.num_qubits, [], None, self, label)

    to_instruction = to_gate

    def repeat(self, reps: int) -> "QuantumCircuit":
        qc = QuantumCircuit(*self.qregs, name=f"{self.name}**{reps}")
        g = self.to_gate()
        for _ in range(int(reps)):
            qc.append(g, range(self.num_qubits))
        return qc

    def power(self, power: int) -> "QuantumCircuit":
        if int(power) != power or power < 0:
            raise ValueError("only non-negative integer powers are supported")
        return self.repeat(int(power))

    def control(self, num_ctrl_qubits: int = 1, ctrl_state: Optional[int] = None) -> Gate:
        """Controlled version of the whole circuit: every leaf is controlled and the circuit's
        global phase becomes a phase gate on the controls."""
        c = int(num_ctrl_qubits)
        qc = QuantumCircuit(c + self.num_qubits, name=f"c_{self.name}")
        ctrls = list(range(c))
        if abs(self.global_phase) > 0:
            if c == 1 and (ctrl_state is None or ctrl_state == 1):
                qc.p(self.global_phase, 0)
            else:
                qc.append(Gate("p", 1, [self.global_phase]).control(c - 1, None if ctrl_state is None else ctrl_state >> 1), ctrls)
        for ins in self.data:
            op = ins.operation
            if op.definition is not None and op.to_matrix() is None:
                sub = op.definition.control(c, ctrl_state)
            else:
                sub = op.control(c, ctrl_state)
            qc.append(sub, ctrls + [c + q for q in ins.qubits])
        return qc.to_gate()

    # -- parameters -------------------------------------------------------------------
    @property
    def parameters(self) -> List[Parameter]:
        seen: Dict[int, Parameter] = {}

        def walk(c: "QuantumCircuit") -> None:
            for ins in c.data:
                op = ins.operation
                base = op.base_gate if isinstance(op, ControlledGate) else op
                for p in base.params:
                    if _is_free(p):
                        seen.setdefault(id(p.parameter), p.parameter)
                if base.definition is not None and base._matrix is None:
                    walk(base.definition)

        walk(self)
        return sorted(seen.values(), key=lambda p: _param_sort_key(p.name))

    @property
    def num_parameters(self) -> int:
        return len(self.parameters)

    def assign_parameters(self, values: Union[Dict[Parameter, float], Sequence[float]]) -> "QuantumCircuit":
        if not isinstance(values, dict):
            plist = self.parameters
            vals = list(np.asarray(values, dtype=float).reshape(-1))
            if len(vals) != len(plist):
                raise ValueError(f"expected {len(plist)} parameter values, got {len(vals)}")
            values = dict(zip(plist, vals))
        qc = QuantumCircuit(*self.qregs, name=self.name, global_phase=self.global_phase)
        for ins in self.data:
            qc.data.append(Instruction(ins.operation.bind(values), ins.qubits))
        return qc

    bind_parameters = assign_parameters

    # -- inspection -------------------------------------------------------------------
    def width(self) -> int:
        return self.num_qubits + self.num_clbits

    def size(self) -> int:
        return len(self.data)

    def depth(self) -> int:
        level = [0] * self.num_qubits
        for ins in self.data:
            d = 1 + max(level[q] for q in ins.qubits)
            for q in ins.qubits:
                level[q] = d
        return max(level) if level else 0

    def count_ops(self) -> Dict[str, int]:
        out: Dict[str, int] = {}
        for ins in self.data:
            out[ins.operation.name] = out.get(ins.operation.name, 0) + 1
        return out

    def decompose(self, reps: int = 1) -> "QuantumCircuit":
        """Expand every gate that has a ``definition`` (one level per repetition).

        Standard gates with matrices are this IR's basis and are left alone, so after enough
        repetitions the circuit is the leaf stream the simulator applies."""
        cur = self
        for _ in range(int(reps)):
            nxt = QuantumCircuit(*cur.qregs, name=cur.name, global_phase=cur.global_phase)
            changed = False
            for ins in cur.data:
                op = ins.operation
                if op.definition is not None and op._matrix is None:
                    changed = True
                    nxt.global_phase += op.definition.global_phase
                    for sub in op.definition.data:
                        nxt.data.append(Instruction(sub.operation, tuple(ins.qubits[q] for q in sub.qubits)))
                else:
                    nxt.data.append(ins)
            cur = nxt
            if not changed:
                break
        return cur

    def draw(self, *_: Any, **__: Any) -> str:
        lines = [f"{self.name}: {self.num_qubits} qubits, global_phase={self.global_phase}"]
        for ins in self.data:
            lines.append(f"  {ins.operation.name}{list(ins.qubits)}")
        return "\n".join(lines)

    def __str__(self) -> str:
        return self.draw()

    def __len__(self) -> int:
        return len(self.data)


def _param_sort_key(name: str) -> Tuple[str, int]:
    if name.endswith("]") and "[" in name:
        base, idx = name[:-1].split("[", 1)
        try:
            return (base, int(idx))
        except ValueError:
            pass
    return (name, -1)


# --------------------------------------------------------------------------------------
# Flattening: the leaf stream Statevector._evolve_instruction would apply
# --------------------------------------------------------------------------------------


@dataclass
class Leaf:
    """One leaf gate of the flattened stream.

    ``kind``: ``"dense"`` (matrix on ``targets``), ``"diag"`` (``payload`` = diagonal over
    ``targets``), ``"ucry"`` (``targets[0]`` = rotated qubit, ``targets[1:]`` = selector qubits,
    ``payload`` = angles), ``"init"`` (``payload`` = amplitudes on ``targets``; only legal while
    those qubits are still |0>).  ``controls``/``ctrl_state`` restrict all kinds."""

    kind: str
    targets: Tuple[int, ...]
    payload: np.ndarray
    controls: Tuple[int, ...] = ()
    ctrl_state: int = 0
    name: str = ""


def _leaf_of(op: Gate, qubits: Tuple[int, ...]) -> Optional[Leaf]:
    controls: Tuple[int, ...] = ()
    ctrl_state = 0
    base = op
    if isinstance(op, ControlledGate):
        c = op.num_ctrl_qubits
        controls, qubits = qubits[:c], qubits[c:]
        ctrl_state = op.ctrl_state
        base = op.base_gate
    if isinstance(base, UCRYGate):
        return Leaf("ucry", qubits, base.angles, controls, ctrl_state, "ucry")
    if isinstance(base, StatePreparation):
        if controls:
            m = base.to_matrix()
            return None if m is None else Leaf("dense", qubits, m, controls, ctrl_state, base.name)
        return Leaf("init", qubits, base.amplitudes, (), 0, base.name)
    if isinstance(base, DiagonalGate):
        return Leaf("diag", qubits, base.diag, controls, ctrl_state, "diagonal")
    m = base.to_matrix()
    if m is None:
        return None
    if base.is_diagonal:
        return Leaf("diag", qubits, np.diag(m).copy(), controls, ctrl_state, base.name)
    return Leaf("dense", qubits, m, controls, ctrl_state, base.name)


def flatten(circuit: QuantumCircuit, parametric: bool = False) -> Tuple[List[Leaf], float]:
    """Leaf stream + accumulated global phase, walking definitions exactly as
    ``Statevector._evolve_instruction`` does: a gate with a matrix is
    applied as is, otherwise ``definition.global_phase`` is multiplied in and the definition is
    walked with remapped qubit arguments.

    With ``parametric=True`` unbound ``ry/rx/rz/p`` gates become ``"param"`` leaves whose payload
    is ``(gate name, parameter index, coeff, offset)`` -- the index is the parameter's position in
    ``circuit.parameters`` (the order ``Estimator.run`` binds ``parameter_values`` in)."""
    leaves: List[Leaf] = []
    phase = [float(circuit.global_phase)]
    pindex = {id(p): i for i, p in enumerate(circuit.parameters)} if parametric else {}

    def walk(c: QuantumCircuit, qmap: Sequence[int]) -> None:
        for ins in c.data:
            op = ins.operation
            qs = tuple(qmap[q] for q in ins.qubits)
            base = op.base_gate if isinstance(op, ControlledGate) else op
            if any(_is_free(p) for p in base.params):
                if not parametric:
                    raise ValueError(f"circuit has unbound parameters (gate {op.name!r})")
                if base.name not in ("ry", "rx", "rz", "p"):
                    raise NotImplementedError(f"parameterised {base.name!r} gates are not supported")
                expr = base.params[0]
                nc = op.num_ctrl_qubits if isinstance(op, ControlledGate) else 0
                leaves.append(Leaf("param", qs[nc:], (base.name, pindex[id(expr.parameter)], expr.coeff, expr.offset),
                                   qs[:nc], op.ctrl_state if nc else 0, base.name))
                continue
            leaf = _leaf_of(op, qs)
            if leaf is not None:
                leaves.append(leaf)
                continue
            if op.definition is None:
                raise ValueError(f"gate {op.name!r} has neither a matrix nor a definition")
            phase[0] += op.definition.global_phase
            walk(op.definition, qs)

    walk(circuit, list(range(circuit.num_qubits)))
    return leaves, phase[0]
