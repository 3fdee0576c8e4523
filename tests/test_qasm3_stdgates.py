"""Reader for the stdgates.inc dialect of OpenQASM 3 (what qiskit 0.46's ``qasm3.dumps`` writes, i.e. the files the
reference leaves behind: utils.py:10-14).  The parsed circuit must act like the same circuit built through the API."""
import math

import numpy as np
import pytest

from oracle import numpy_sv
from qls_b200 import qasm3
from qls_b200.circuit import Gate, QuantumCircuit, QuantumRegister

from _util import circuit_to_stream

TEXT = """OPENQASM 3.0;
include "stdgates.inc";
// a definition with an angle parameter, as the exporter writes for gates outside stdgates.inc
gate rzx(_gate_p_0) _gate_q_0, _gate_q_1 {
  h _gate_q_1;
  cx _gate_q_0, _gate_q_1;
  rz(_gate_p_0) _gate_q_1;
  cx _gate_q_0, _gate_q_1;
  h _gate_q_1;
}
gate iqft_dg _gate_q_0, _gate_q_1, _gate_q_2 {
  gphase(pi/16);
  h _gate_q_2;
  cp(-pi/2) _gate_q_1, _gate_q_2;
  h _gate_q_1;
  cp(-pi/4) _gate_q_0, _gate_q_2;
  cp(-pi/2) _gate_q_0, _gate_q_1;
  h _gate_q_0;
}
bit[2] c;
qubit[1] q_b;
qubit[3] q_l;
qubit[1] q_f;
U(pi/2, 0.25, -pi) q_b[0];
h q_l;  /* broadcast over the register */
cp(3*pi/8) q_l[0], q_b[0];
rzx(pi/4) q_l[1], q_b[0];
rzx(-0.5) q_l[2], q_b[0];
iqft_dg q_l[0], q_l[1], q_l[2];
barrier q_b[0], q_l[0];
ctrl @ ry(2*arcsin(0.25)) q_l[2], q_f[0];
negctrl @ inv @ s q_l[0], q_f[0];
pow(3) @ t q_f[0];
ccx q_l[0], q_l[1], q_f[0];
cswap q_l[0], q_l[1], q_l[2];
cu(0.1, 0.2, 0.3, 0.4) q_l[0], q_b[0];
ctrl @ negctrl @ x q_l[0], q_l[1], q_b[0];
u2(0.3, -0.7) q_f[0];
let top = q_l[1:2];
sx top;
cx q_l[0:1], top;
gphase(pi/8);
"""


def _built_by_hand() -> QuantumCircuit:
    qb, ql, qf = QuantumRegister(1, "q_b"), QuantumRegister(3, "q_l"), QuantumRegister(1, "q_f")
    qc = QuantumCircuit(qb, ql, qf, global_phase=math.pi / 8)
    b, l0, l1, l2, f = 0, 1, 2, 3, 4

    def rzx(th, a, t):
        qc.h(t), qc.cx(a, t), qc.rz(th, t), qc.cx(a, t), qc.h(t)

    qc.u(math.pi / 2, 0.25, -math.pi, b)
    for q in (l0, l1, l2):
        qc.h(q)
    qc.cp(3 * math.pi / 8, l0, b)
    rzx(math.pi / 4, l1, b)
    rzx(-0.5, l2, b)
    qc.global_phase += math.pi / 16
    qc.h(l2), qc.cp(-math.pi / 2, l1, l2), qc.h(l1), qc.cp(-math.pi / 4, l0, l2), qc.cp(-math.pi / 2, l0, l1), qc.h(l0)
    qc.cry(2 * math.asin(0.25), l2, f)
    qc.append(Gate("sdg", 1).control(1, 0), [l0, f])
    qc.p(3 * math.pi / 4, f)
    qc.ccx(l0, l1, f)
    qc.append(Gate("swap", 2).control(1), [l0, l1, l2])
    m = Gate("u", 1, [0.1, 0.2, 0.3]).to_matrix() * np.exp(0.4j)
    qc.append(Gate("unitary", 1, [], m).control(1), [l0, b])
    qc.append(Gate("x", 1).control(2, 0b01), [l0, l1, b])  # control 0 (q_l[0]) on 1, control 1 (q_l[1]) on 0
    qc.u(math.pi / 2, 0.3, -0.7, f)
    qc.sx(l1), qc.sx(l2)
    qc.cx(l0, l1), qc.cx(l1, l2)
    return qc


def _state(qc, seed=3):
    """Evolve a random start state (so that every control value is exercised), not just |0...0>."""
    n = qc.num_qubits
    rng = np.random.default_rng(seed)
    prep = QuantumCircuit(n)
    amps = rng.standard_normal(2**n) + 1j * rng.standard_normal(2**n)
    prep.prepare_state(amps / np.linalg.norm(amps), list(range(n)))
    full = prep.compose(qc, list(range(n)))
    full.global_phase = prep.global_phase + qc.global_phase
    return numpy_sv.run(n, circuit_to_stream(full))


def test_stdgates_dialect_matches_hand_built_circuit():
    parsed = qasm3.loads_stdgates(TEXT)
    ref = _built_by_hand()
    assert [(r.name, r.size) for r in parsed.qregs] == [("q_b", 1), ("q_l", 3), ("q_f", 1)]
    assert np.max(np.abs(_state(parsed) - _state(ref))) < 1e-14


def test_stdgates_reader_takes_our_own_export_of_standard_gates(tmp_path):
    qc = QuantumCircuit(3)
    qc.h(0), qc.cx(0, 1), qc.rz(0.3, 2), qc.ccx(0, 1, 2), qc.cp(-1.25, 2, 0), qc.swap(1, 2), qc.u(0.1, 0.2, 0.3, 1)
    path = tmp_path / "c.qasm3"
    path.write_text(qasm3.dumps(qc))
    back = qasm3.load(str(path))  # no pragma lines -> general reader
    assert np.max(np.abs(_state(back) - _state(qc))) < 1e-14


@pytest.mark.parametrize("bad", [
    "OPENQASM 3;\nqubit[1] q;\nbit[1] c;\nc[0] = measure q[0];\n",
    "OPENQASM 3;\nqubit[1] q;\nreset q[0];\n",
    "OPENQASM 3;\nqubit[2] q;\nfoo q[0];\n",
    "OPENQASM 3;\nqubit[2] q;\nrz(theta) q[0];\n",
    "OPENQASM 3;\nqubit[2] q;\nqubit[3] r;\ncx q, r;\n",
    "OPENQASM 2.0;\nqreg q[1];\n",
])
def test_stdgates_reader_rejects(bad):
    with pytest.raises(ValueError):
        qasm3.loads_stdgates(bad)
