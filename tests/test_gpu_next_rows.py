"""GPU tests for the SURVEY.md section 8(f) "next" rows: the QASM3 files the reference writes, read back and
simulated on the device (row f1; reference ``utils.py:10-14``, ``hhl_qiskit_implementation.py:64-71``), and the
comparison harness / CLI driven end to end through the façade on the device (row f4; reference
``compare_classiq_qiskit.py:30-116``, ``execute_framework.py``)."""
import csv

import numpy as np
import pytest

from oracle import numpy_sv

from _util import circuit_to_stream

pytestmark = pytest.mark.gpu

TOL64 = 1e-12


def test_stdgates_file_on_the_device(cuda_device, tmp_path):
    """A stdgates.inc-dialect file (what qiskit 0.46's exporter leaves on disk) -> ``qasm3.load`` -> ``Statevector`` on
    the B200 against the oracle run on the same parsed circuit; also at a size where the tile-pass kernels (not the
    small-state path) do the work."""
    from qls_b200 import qasm3
    from qls_b200.statevector import Statevector
    from test_qasm3_stdgates import TEXT, _built_by_hand

    path = tmp_path / "hhl_circuit.qasm3"
    path.write_text(TEXT, encoding="utf-8")
    qc = qasm3.load(str(path))
    got = Statevector(qc).data
    want = numpy_sv.run(qc.num_qubits, circuit_to_stream(_built_by_hand()))
    assert np.max(np.abs(got - want)) <= TOL64

    # a wider register: the same definitions applied across 18 qubits
    n = 18
    lines = ["OPENQASM 3.0;", 'include "stdgates.inc";',
             "gate blk(_gate_p_0) _gate_q_0, _gate_q_1, _gate_q_2 {", "  h _gate_q_0;", "  cp(_gate_p_0) _gate_q_0, _gate_q_1;",
             "  ccx _gate_q_0, _gate_q_1, _gate_q_2;", "  ry(_gate_p_0/2) _gate_q_2;", "}", f"qubit[{n}] q;", "h q;"]
    rng = np.random.default_rng(5)
    for _ in range(40):
        a, b, c = (int(x) for x in rng.choice(n, 3, replace=False))
        lines.append(f"blk({float(rng.uniform(-3, 3))!r}) q[{a}], q[{b}], q[{c}];")
        lines.append(f"ctrl @ rz({float(rng.uniform(-3, 3))!r}) q[{c}], q[{a}];")
    path2 = tmp_path / "wide.qasm3"
    path2.write_text("\n".join(lines) + "\n", encoding="utf-8")
    qc2 = qasm3.load(str(path2))
    got2 = Statevector(qc2).data
    want2 = numpy_sv.run(n, circuit_to_stream(qc2))
    assert np.max(np.abs(got2 - want2)) <= TOL64
    assert abs(np.vdot(got2, got2).real - 1.0) <= 1e-12


def test_own_dialect_round_trip_on_the_device(cuda_device, tmp_path):
    """The file ``solve_hhl_qiskit`` writes (own lossless dialect) read back and simulated: same state as the circuit
    that was exported."""
    from qls_b200 import qasm3
    from qls_b200.hhl import HHL
    from qls_b200.statevector import Statevector
    from qls_b200.toymodels import ClassiqDemoExample

    model = ClassiqDemoExample()
    qc = HHL().construct_circuit(model.matrix_a, model.vector_b)
    path = tmp_path / "c.qasm3"
    with open(path, "w", encoding="utf-8") as f:
        qasm3.dump(qc, f)
    back = qasm3.load(str(path))
    a, b = Statevector(qc).data, Statevector(back).data
    assert np.max(np.abs(a - b)) <= 1e-14


def test_harness_and_cli_on_the_device(cuda_device, tmp_path, monkeypatch):
    """``compare_qls`` over two toy models and the CLI, no stand-ins: the solves run on the B200, the CSV carries the
    reference's columns, and the relative distance to the classical solution is what the reference's test accepts."""
    from qls_b200 import compare, execute_framework, toymodels

    monkeypatch.chdir(tmp_path)
    models = [toymodels.Qiskit4QubitExample(), toymodels.ClassiqDemoExample()]
    path = compare.compare_qls(models, filebasename="cmp")
    rows = list(csv.reader(open(path, encoding="utf-8")))
    assert rows[0] == compare.CSV_HEADERS and len(rows) == 3
    assert rows[1][0] == "Qiskit4QubitExample" and rows[1][1] == "hhl_qiskit"
    assert float(rows[1][6]) < 1e-9          # 2x2 KAT: exact
    assert float(rows[2][6]) < 20.0          # reference tests/test_hhl_implementations.py: < 20 % for the 4x4 demo

    np.savetxt(tmp_path / "a.csv", np.array([[1, -1 / 3], [-1 / 3, 1]]), delimiter=",")
    np.savetxt(tmp_path / "b.csv", np.array([1.0, 0.0]), delimiter=",")
    qls = execute_framework.main(["-i", "hhl_qiskit", "-m", "a.csv", "-v", "b.csv", "-y"])
    assert np.allclose(qls.solution, [1.125, 0.375], atol=1e-10)
    assert qls.circuit_width == 5 and qls.circuit_depth > 0 and qls.qasm_circuit.startswith("OPENQASM 3")
