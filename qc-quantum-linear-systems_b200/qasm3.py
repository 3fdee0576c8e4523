"""OpenQASM 3 export / import of this IR (the wire format on the output side of the path:
``circuit_to_qasm3`` in the reference's ``utils.py:10-14``, files ``hhl_qiskit_circuit.qasm3`` /
``vqls_qiskit_circuit.qasm3``).

Dialect: header ``OPENQASM 3;`` + ``include "stdgates.inc";``, one ``qubit[n] <reg>;`` per
register, standard-library gate calls, ``ctrl(n) @`` / ``negctrl(n) @`` modifiers, hierarchical
``gate`` definitions for composite gates.  Gates that only exist as numeric tables (``unitary``,
``diagonal``, ``ucry``, ``state_preparation``) are declared by ``pragma qls`` lines carrying the
numbers (hex floats -> lossless), so ``loads(dumps(c))`` reproduces the circuit exactly and
``dumps(loads(text)) == text``.

The reference's text comes from ``qiskit.qasm3.dumps`` after ``decompose(reps=10)``, whose
synthesis passes (isometry, multiplexer and unitary decompositions) are not restated here; when
qiskit is importable the circuit should be exported by qiskit itself.
"""
from __future__ import annotations

import re
from typing import Dict, List, Tuple

import numpy as np

from .circuit import ControlledGate, DiagonalGate, Gate, QuantumCircuit, QuantumRegister, StatePreparation, UCRYGate

_STD = {"id", "h", "x", "y", "z", "s", "sdg", "t", "tdg", "sx", "p", "rx", "ry", "rz", "swap"}


def _hexc(values: np.ndarray) -> str:
    v = np.asarray(values, dtype=complex).reshape(-1)
    return " ".join(f"{float(z.real).hex()},{float(z.imag).hex()}" for z in v)


def _unhexc(text: str) -> np.ndarray:
    out = []
    for tok in text.split():
        re_, im_ = tok.split(",")
        out.append(complex(float.fromhex(re_), float.fromhex(im_)))
    return np.array(out, dtype=complex)


def _fmt(x: float) -> str:
    return repr(float(x))


class _Dumper:
    def __init__(self, circuit: QuantumCircuit):
        self.circuit = circuit
        self.pragmas: List[str] = []
        self.gate_defs: List[str] = []
        self.tables: Dict[int, str] = {}
        self.defs: Dict[int, str] = {}
        self.names: Dict[str, int] = {}

    def _unique(self, base: str) -> str:
        base = re.sub(r"[^A-Za-z0-9_]", "_", base) or "g"
        if base[0].isdigit():
            base = "g_" + base
        k = self.names.get(base, 0)
        self.names[base] = k + 1
        return f"{base}_{k}"

    def _table_gate(self, g: Gate) -> str:
        if id(g) in self.tables:
            return self.tables[id(g)]
        if isinstance(g, UCRYGate):
            nm = self._unique("ucry")
            self.pragmas.append(f"pragma qls ucry {nm} {g.num_qubits} " + " ".join(float(a).hex() for a in g.angles))
        elif isinstance(g, StatePreparation):
            nm = self._unique("state_preparation")
            self.pragmas.append(f"pragma qls state_preparation {nm} {g.num_qubits} " + _hexc(g.amplitudes))
        elif isinstance(g, DiagonalGate):
            nm = self._unique("diagonal")
            self.pragmas.append(f"pragma qls diagonal {nm} {g.num_qubits} " + _hexc(g.diag))
        else:
            nm = self._unique("unitary")
            self.pragmas.append(f"pragma qls unitary {nm} {g.num_qubits} " + _hexc(g.to_matrix()))
        self.tables[id(g)] = nm
        return nm

    def _def_gate(self, g: Gate) -> str:
        if id(g.definition) in self.defs:
            return self.defs[id(g.definition)]
        nm = self._unique(g.name)
        self.defs[id(g.definition)] = nm
        d = g.definition
        args = ", ".join(f"_q{i}" for i in range(d.num_qubits))
        body = [self._call(ins.operation, [f"_q{q}" for q in ins.qubits]) for ins in d.data]
        phase = f"  gphase({_fmt(d.global_phase)});\n" if d.global_phase else ""
        self.gate_defs.append(f"gate {nm} {args} {{\n{phase}" + "".join(f"  {b}\n" for b in body) + "}")
        return nm

    def _call(self, g: Gate, qs: List[str]) -> str:
        prefix = ""
        if isinstance(g, ControlledGate):
            c = g.num_ctrl_qubits
            if g.ctrl_state == 2**c - 1:
                prefix = "ctrl @ " if c == 1 else f"ctrl({c}) @ "
            else:  # mixed control state: one modifier per control, innermost last
                prefix = "".join(("ctrl @ " if (g.ctrl_state >> j) & 1 else "negctrl @ ") for j in range(c))
            g = g.base_gate
        if g.definition is not None and g._matrix is None:
            name, params = self._def_gate(g), ""
        elif g.name in _STD:
            name = g.name
            params = "(" + ", ".join(_fmt(p) for p in g.params) + ")" if g.params else ""
        elif g.name == "u":
            name, params = "U", "(" + ", ".join(_fmt(p) for p in g.params) + ")"
        else:
            name, params = self._table_gate(g), ""
        return f"{prefix}{name}{params} {', '.join(qs)};"

    def dumps(self) -> str:
        c = self.circuit
        regs, seen = [], {}
        for r in c.qregs:
            k = seen.get(r.name, 0)
            seen[r.name] = k + 1
            regs.append(r.name if k == 0 else f"{r.name}_{k}")
        qnames: List[str] = []
        for rn, r in zip(regs, c.qregs):
            qnames += [f"{rn}[{i}]" for i in range(r.size)]
        body = [self._call(ins.operation, [qnames[q] for q in ins.qubits]) for ins in c.data]
        out = ["OPENQASM 3;", 'include "stdgates.inc";']
        out += self.pragmas
        out += self.gate_defs
        out += [f"qubit[{r.size}] {rn};" for rn, r in zip(regs, c.qregs)]
        if c.global_phase:
            out.append(f"gphase({_fmt(c.global_phase)});")
        out += body
        return "\n".join(out) + "\n"


def dumps(circuit: QuantumCircuit) -> str:
    return _Dumper(circuit).dumps()


def dump(circuit: QuantumCircuit, stream) -> None:
    stream.write(dumps(circuit))


# --------------------------------------------------------------------------------------
# Import (the dialect written by dumps)
# --------------------------------------------------------------------------------------

_CALL = re.compile(r"^((?:(?:neg)?ctrl(?:\(\d+\))? @ )*)([A-Za-z_][A-Za-z0-9_]*)(?:\(([^)]*)\))? (.*);$")


def loads(text: str) -> QuantumCircuit:
    lines = [ln.rstrip() for ln in text.splitlines() if ln.strip()]
    if not lines or not lines[0].startswith("OPENQASM 3"):
        raise ValueError("not an OpenQASM 3 program")
    tables: Dict[str, Gate] = {}
    gate_defs: Dict[str, Tuple[List[str], List[str], float, str]] = {}
    regs: List[Tuple[str, int]] = []
    body: List[str] = []
    gphase = 0.0
    i = 1
    while i < len(lines):
        ln = lines[i]
        if ln.startswith("include"):
            pass
        elif ln.startswith("pragma qls "):
            _, _, kind, name, nq, rest = ln.split(" ", 5)
            nq = int(nq)
            if kind == "ucry":
                tables[name] = UCRYGate([float.fromhex(t) for t in rest.split()])
            elif kind == "state_preparation":
                tables[name] = StatePreparation(_unhexc(rest))
            elif kind == "diagonal":
                tables[name] = DiagonalGate(_unhexc(rest))
            elif kind == "unitary":
                tables[name] = Gate("unitary", nq, [], _unhexc(rest).reshape(2**nq, 2**nq))
            else:
                raise ValueError(f"unknown pragma {kind}")
        elif ln.startswith("gate "):
            head = ln[5:ln.index("{")].strip()
            name, args = head.split(" ", 1)
            argl = [a.strip() for a in args.split(",")]
            inner: List[str] = []
            ph = 0.0
            i += 1
            while lines[i].strip() != "}":
                st = lines[i].strip()
                if st.startswith("gphase("):
                    ph = float(st[7:st.index(")")])
                else:
                    inner.append(st)
                i += 1
            gate_defs[name] = (argl, inner, ph, name)
        elif ln.startswith("qubit["):
            m = re.match(r"qubit\[(\d+)\] ([A-Za-z_][A-Za-z0-9_]*);", ln)
            regs.append((m.group(2), int(m.group(1))))
        elif ln.startswith("gphase("):
            gphase = float(ln[7:ln.index(")")])
        else:
            body.append(ln)
        i += 1

    built: Dict[str, Gate] = {}

    def gate_of(name: str, params: List[float]) -> Gate:
        if name in tables:
            return tables[name]
        if name in gate_defs:
            if name not in built:
                argl, inner, ph, _ = gate_defs[name]
                sub = QuantumCircuit(len(argl), name=re.sub(r"_\d+$", "", name), global_phase=ph)
                amap = {a: j for j, a in enumerate(argl)}
                for st in inner:
                    apply(sub, st, amap)
                built[name] = sub.to_gate()
            return built[name]
        if name == "U":
            return Gate("u", 1, params)
        if name in _STD:
            return Gate(name, 2 if name == "swap" else 1, params)
        raise ValueError(f"unknown gate {name}")

    def apply(qc: QuantumCircuit, st: str, qmap: Dict[str, int]) -> None:
        m = _CALL.match(st)
        if not m:
            raise ValueError(f"cannot parse statement: {st}")
        mods, name, ptxt, qtxt = m.groups()
        params = [float(p) for p in ptxt.split(",")] if ptxt else []
        g = gate_of(name, params)
        bits: List[int] = []
        for mod in re.findall(r"(neg)?ctrl(?:\((\d+)\))? @ ", mods):
            cnt = int(mod[1]) if mod[1] else 1
            bits += [0 if mod[0] else 1] * cnt
        if bits:
            g = g.control(len(bits), sum(b << j for j, b in enumerate(bits)))
        qc.append(g, [qmap[q.strip()] for q in qtxt.split(",")])

    qc = QuantumCircuit(*[QuantumRegister(sz, nm) for nm, sz in regs], name="circuit", global_phase=gphase)
    qmap: Dict[str, int] = {}
    off = 0
    for nm, sz in regs:
        for j in range(sz):
            qmap[f"{nm}[{j}]"] = off + j
        off += sz
    for st in body:
        apply(qc, st, qmap)
    return qc
