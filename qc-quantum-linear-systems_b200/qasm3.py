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


_RAW_MIN = 256  # tables from this many entries on are written as one hexadecimal run of their bytes


def _hexc(values: np.ndarray) -> str:
    """Lossless text form of a complex table: ``re,im`` pairs of C99 hexadecimal floats for small tables (readable), one
    run ``h:<hex of the little-endian complex128 bytes>`` for large ones -- the 2^10 x 2^10 unitaries of a collapsed
    ``exp(iAt)`` power are 16 MiB each and a per-element format costs ~1.4 s per matrix (34 s per exported HHL circuit at
    N = 2^10; 1 s with the byte form)."""
    v = np.ascontiguousarray(np.asarray(values, dtype="<c16").reshape(-1))
    if v.size >= _RAW_MIN:
        return "h:" + v.tobytes().hex()
    return " ".join(f"{float(z.real).hex()},{float(z.imag).hex()}" for z in v)


def _unhexc(text: str) -> np.ndarray:
    text = text.strip()
    if text.startswith("h:"):
        return np.frombuffer(bytes.fromhex(text[2:]), dtype="<c16").astype(complex)
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


# --------------------------------------------------------------------------------------
# Import of the general "stdgates.inc" dialect (what qiskit 0.46's ``qasm3.dumps`` writes: the files
# ``hhl_qiskit_circuit.qasm3`` / ``vqls_qiskit_circuit.qasm3`` / ``{name}_{method}.qasm`` of the reference,
# utils.py:10-14, quantum_linear_solver.py:111-116), so that a circuit exported by the reference can be
# simulated here without qiskit.  Supported: ``qubit`` declarations (registers and single qubits),
# ``let`` aliases (index, slice, ``++``), every gate of stdgates.inc plus the built-in ``U`` and
# ``gphase``, gate definitions with angle parameters, ``ctrl`` / ``negctrl`` / ``inv`` / ``pow(int)``
# modifiers, register broadcasting, ``barrier`` (ignored), constant expressions with ``pi``.
# Classical control flow, ``measure`` and ``reset`` are rejected (the reference simulates with
# ``Statevector``, which rejects them as well).
# --------------------------------------------------------------------------------------

import ast as _ast
import math as _math

_STD1 = {"id": "id", "x": "x", "y": "y", "z": "z", "h": "h", "s": "s", "sdg": "sdg", "t": "t", "tdg": "tdg", "sx": "sx",
         "rx": "rx", "ry": "ry", "rz": "rz", "p": "p", "phase": "p", "u1": "p"}
_CTRL1 = {"cx": "x", "CX": "x", "cy": "y", "cz": "z", "ch": "h", "cp": "p", "cphase": "p", "cu1": "p", "crx": "rx", "cry": "ry",
          "crz": "rz"}
_FUNCS = {"sin": _math.sin, "cos": _math.cos, "tan": _math.tan, "exp": _math.exp, "ln": _math.log, "sqrt": _math.sqrt,
          "arcsin": _math.asin, "arccos": _math.acos, "arctan": _math.atan}
_CONSTS = {"pi": _math.pi, "tau": 2 * _math.pi, "euler": _math.e}


def _eval_expr(text: str, env: Dict[str, float]) -> float:
    """Constant angle expression of OpenQASM 3 (``pi/4``, ``-3*pi/8``, ``2**-3``, ``sin(x)``...)."""
    src = text.strip().replace("π", "pi").replace("τ", "tau").replace("ℇ", "euler")
    try:
        tree = _ast.parse(src, mode="eval")
    except SyntaxError as exc:
        raise ValueError(f"cannot parse expression {text!r}") from exc

    def ev(node) -> float:
        if isinstance(node, _ast.Expression):
            return ev(node.body)
        if isinstance(node, _ast.Constant) and isinstance(node.value, (int, float)):
            return float(node.value)
        if isinstance(node, _ast.Name):
            if node.id in env:
                return float(env[node.id])
            if node.id in _CONSTS:
                return _CONSTS[node.id]
            raise ValueError(f"unknown identifier {node.id!r} in expression {text!r}")
        if isinstance(node, _ast.UnaryOp) and isinstance(node.op, (_ast.USub, _ast.UAdd)):
            v = ev(node.operand)
            return -v if isinstance(node.op, _ast.USub) else v
        if isinstance(node, _ast.BinOp):
            a, b = ev(node.left), ev(node.right)
            if isinstance(node.op, _ast.Add):
                return a + b
            if isinstance(node.op, _ast.Sub):
                return a - b
            if isinstance(node.op, _ast.Mult):
                return a * b
            if isinstance(node.op, _ast.Div):
                return a / b
            if isinstance(node.op, _ast.Pow):
                return a**b
        if isinstance(node, _ast.Call) and isinstance(node.func, _ast.Name) and node.func.id in _FUNCS and len(node.args) == 1:
            return _FUNCS[node.func.id](ev(node.args[0]))
        raise ValueError(f"unsupported expression {text!r}")

    return ev(tree)


def _split_top(text: str, sep: str) -> List[str]:
    """Split at ``sep`` outside parentheses / brackets."""
    out, depth, cur = [], 0, []
    i = 0
    while i < len(text):
        ch = text[i]
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if depth == 0 and text.startswith(sep, i):
            out.append("".join(cur))
            cur = []
            i += len(sep)
            continue
        cur.append(ch)
        i += 1
    out.append("".join(cur))
    return [t.strip() for t in out if t.strip()]


def _statements(text: str) -> List[str]:
    """Top-level statements; a ``gate`` definition (with its braces) is one statement."""
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    out, depth, cur = [], 0, []
    for ch in text:
        if ch == "{":
            depth += 1
        if ch == "}":
            depth -= 1
            if depth == 0:
                cur.append(ch)
                out.append("".join(cur).strip())
                cur = []
                continue
        if ch == ";" and depth == 0:
            out.append("".join(cur).strip())
            cur = []
            continue
        cur.append(ch)
    if "".join(cur).strip():
        raise ValueError("unterminated statement at the end of the program")
    return [" ".join(s.split()) for s in out if s.strip()]


_HEAD = re.compile(r"^([A-Za-z_][A-Za-z0-9_]*)\s*(?:\((.*)\))?\s*(.*)$", re.S)
_MOD = re.compile(r"^(negctrl|ctrl|inv|pow)\s*(?:\(([^@]*)\))?\s*@\s*(.*)$", re.S)


def _mat_pow(gate: Gate, k: int) -> Gate:
    m = gate.to_matrix()
    if m is None:
        sub = QuantumCircuit(gate.num_qubits, name=gate.name)
        sub.append(gate, list(range(gate.num_qubits)))
        return (sub.repeat(k) if k >= 0 else sub.inverse().repeat(-k)).to_gate()
    return Gate("unitary", gate.num_qubits, [], np.linalg.matrix_power(m, k))


def loads_stdgates(text: str) -> QuantumCircuit:
    """Parse an OpenQASM 3 program written against ``stdgates.inc`` (see the comment above)."""
    sts = _statements(text)
    if not sts or not re.match(r"^OPENQASM 3(\.\d+)?$", sts[0]):
        raise ValueError("not an OpenQASM 3 program")
    regs: List[Tuple[str, int]] = []          # declaration order = qubit order
    operands: Dict[str, List[int]] = {}        # register / alias name -> global qubit indexes
    gate_defs: Dict[str, Tuple[List[str], List[str], List[str]]] = {}
    calls: List[str] = []
    phase0 = 0.0
    n_qubits = 0
    for st in sts[1:]:
        if re.match(r"^(include|barrier|bit|creg|defcalgrammar)\b", st):
            continue
        m = re.match(r"^(?:qubit|qreg)\s*(?:\[(\d+)\])?\s*([A-Za-z_$][A-Za-z0-9_]*)\s*(?:\[(\d+)\])?$", st)
        if m:
            size = int(m.group(1) or m.group(3) or 1)
            name = m.group(2)
            regs.append((name, size))
            operands[name] = list(range(n_qubits, n_qubits + size))
            n_qubits += size
            continue
        m = re.match(r"^let\s+([A-Za-z_][A-Za-z0-9_]*)\s*=\s*(.*)$", st)
        if m:
            operands[m.group(1)] = [q for part in _split_top(m.group(2), "++") for q in _operand(part, operands)]
            continue
        if st.startswith("gate "):
            head, body = st[5:st.index("{")].strip(), st[st.index("{") + 1:st.rindex("}")]
            hm = _HEAD.match(head)
            name, ptxt, qtxt = hm.group(1), hm.group(2), hm.group(3)
            gate_defs[name] = ([p.strip() for p in ptxt.split(",")] if ptxt and ptxt.strip() else [],
                               [q.strip() for q in qtxt.split(",")], [b.strip() for b in body.split(";") if b.strip()])
            continue
        if re.match(r"^(measure|reset|if|for|while|def|cal|extern|input|output|const|int|uint|float|angle|bool|duration|stretch|box|delay)\b", st) \
                or re.search(r"=\s*measure\b", st):
            raise ValueError(f"statement outside the unitary subset: {st!r}")
        calls.append(st)

    built: Dict[Tuple[str, Tuple[float, ...]], Gate] = {}

    def base_gate(name: str, params: List[float]) -> Gate:
        if name in gate_defs:  # a definition in the file shadows stdgates.inc, as in OpenQASM
            key = (name, tuple(params))
            if key not in built:
                pnames, qnames, body = gate_defs[name]
                if len(pnames) != len(params):
                    raise ValueError(f"gate {name} takes {len(pnames)} parameters")
                sub = QuantumCircuit(len(qnames), name=name)
                env = dict(zip(pnames, params))
                local = {q: [j] for j, q in enumerate(qnames)}
                for b in body:
                    if b.startswith("barrier"):
                        continue
                    if b.startswith("gphase"):
                        sub.global_phase += _eval_expr(b[b.index("(") + 1:b.rindex(")")], env)
                    else:
                        apply(sub, b, local, env)
                built[key] = sub.to_gate()
            return built[key]
        if name in _STD1:
            return Gate(_STD1[name], 1, params)
        if name in ("U", "u3", "u"):
            return Gate("u", 1, params)
        if name == "u2":
            return Gate("u", 1, [_math.pi / 2, params[0], params[1]])
        if name in _CTRL1:
            return Gate(_CTRL1[name], 1, params).control(1)
        if name == "swap":
            return Gate("swap", 2)
        if name == "ccx":
            return Gate("x", 1).control(2)
        if name == "cswap":
            return Gate("swap", 2).control(1)
        if name in ("cu", "cu3"):
            th, ph, lam = params[:3]
            gam = params[3] if name == "cu" else 0.0
            m = Gate("u", 1, [th, ph, lam]).to_matrix() * np.exp(1j * gam)
            return Gate("unitary", 1, [], m).control(1)
        raise ValueError(f"unknown gate {name!r}")

    def apply(qc: QuantumCircuit, st: str, ops: Dict[str, List[int]], env: Dict[str, float]) -> None:
        mods: List[Tuple[str, int]] = []
        rest = st
        while True:
            mm = _MOD.match(rest)
            if not mm:
                break
            kind, arg, rest = mm.group(1), mm.group(2), mm.group(3)
            val = 1
            if arg is not None and arg.strip():
                v = _eval_expr(arg, env)
                if v != int(v):
                    raise ValueError(f"non-integer modifier argument in {st!r}")
                val = int(v)
            mods.append((kind, val))
        hm = _HEAD.match(rest)
        if not hm:
            raise ValueError(f"cannot parse statement: {st!r}")
        name, ptxt, qtxt = hm.group(1), hm.group(2), hm.group(3)
        params = [_eval_expr(p, env) for p in _split_top(ptxt, ",")] if ptxt and ptxt.strip() else []
        if name == "gphase":
            if mods:
                raise ValueError("modified gphase is not supported")
            qc.global_phase += params[0]
            return
        g = base_gate(name, params)
        n_ctrl_total = 0
        for kind, val in reversed(mods):  # the modifier next to the gate applies first
            if kind == "inv":
                g = g.inverse()
            elif kind == "pow":
                g = _mat_pow(g, val)
            else:
                # new controls go in FRONT of the operand list; the control state is little-endian over them
                state = (2**val - 1) if kind == "ctrl" else 0
                if isinstance(g, ControlledGate):
                    c0 = g.num_ctrl_qubits
                    g = g.base_gate.control(val + c0, state | (g.ctrl_state << val))
                else:
                    g = g.control(val, state)
                n_ctrl_total += val
        groups = [_operand(q, ops) for q in _split_top(qtxt, ",")]
        if len(groups) != g.num_qubits:
            raise ValueError(f"{name} acts on {g.num_qubits} qubits, got {len(groups)} operands in {st!r}")
        width = max(len(gr) for gr in groups)
        if any(len(gr) not in (1, width) for gr in groups):
            raise ValueError(f"cannot broadcast operands of different sizes in {st!r}")
        for j in range(width):
            qc.append(g, [gr[j] if len(gr) > 1 else gr[0] for gr in groups])

    qc = QuantumCircuit(*[QuantumRegister(sz, nm) for nm, sz in regs], name="circuit", global_phase=phase0)
    for st in calls:
        apply(qc, st, operands, {})
    return qc


def _operand(text: str, operands: Dict[str, List[int]]) -> List[int]:
    """``name`` | ``name[i]`` | ``name[a:b]`` (inclusive) | ``name[a:s:b]`` | ``{a, b, ...}`` -> qubit indexes."""
    text = text.strip()
    if text.startswith("{") and text.endswith("}"):
        return [q for part in _split_top(text[1:-1], ",") for q in _operand(part, operands)]
    m = re.match(r"^([A-Za-z_$][A-Za-z0-9_]*)\s*(?:\[(.*)\])?$", text)
    if not m or m.group(1) not in operands:
        raise ValueError(f"unknown qubit operand {text!r}")
    base = operands[m.group(1)]
    if m.group(2) is None:
        return list(base)
    idx = [t.strip() for t in m.group(2).split(":")]
    n = len(base)

    def at(t: str, default: int) -> int:
        if t == "":
            return default
        v = int(_eval_expr(t, {}))
        return v + n if v < 0 else v

    if len(idx) == 1:
        return [base[at(idx[0], 0)]]
    if len(idx) == 2:
        a, b, s = at(idx[0], 0), at(idx[1], n - 1), 1
    else:
        a, s, b = at(idx[0], 0), int(_eval_expr(idx[1], {})), at(idx[2], n - 1)
    return [base[i] for i in range(a, b + (1 if s > 0 else -1), s)]


def load(path: str) -> QuantumCircuit:
    """Read a ``.qasm3`` / ``.qasm`` file: this package's own lossless dialect or the stdgates dialect."""
    with open(path) as fh:
        text = fh.read()
    if "pragma qls " in text:
        return loads(text)
    return loads_stdgates(text)
