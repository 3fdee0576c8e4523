"""generators, scaling toy models, CSV harness and CLI of the reference
(``utils.py:122-262``, ``toymodels.py:188-330``, ``compare_classiq_qiskit.py:30-116``,
``execute_framework.py``) restated on this engine.  Host logic only: the solver is stubbed by the
oracle-backed classical solution where a GPU would be needed."""
import csv
import os

import numpy as np
import pytest

from qls_b200 import compare, execute_framework, toymodels, utils


def test_generate_random_vector_levels():
    np.random.seed(3)
    for level in (0.0, 0.3, 1.0):
        v = utils.generate_random_vector(16, level)
        assert v.shape == (16, 1) and np.isclose(np.linalg.norm(v), 1.0)
    assert np.all(utils.generate_random_vector(8, 1.0) >= 0)  # all-uniform part is non-negative
    with pytest.raises(ValueError):
        utils.generate_random_vector(4, 1.5)


def test_entropy_of_uniform_vector():
    assert np.isclose(utils.vector_uniformity_entropy(np.ones(8)), 3.0, atol=1e-6)
    assert utils.vector_uniformity_entropy(np.array([1.0, 0, 0, 0])) < 1e-6


def test_s_sparse_matrix_rows():
    np.random.seed(5)
    m = utils.generate_s_sparse_matrix(8, 3)
    assert m.shape == (8, 8) and np.all((m != 0).sum(axis=1) == 3)
    for bad in ((8, 0), (0, 1), (4, 5)):
        with pytest.raises(ValueError):
            utils.generate_s_sparse_matrix(*bad)
    assert utils.is_matrix_well_conditioned(np.eye(3)) and not utils.is_matrix_well_conditioned(np.diag([1.0, 1e-3]))


def test_random_and_scaling_models():
    np.random.seed(11)
    m = toymodels.RandomNQubitProblem(2)
    assert m.matrix_a.shape == (4, 4) and np.allclose(m.matrix_a, m.matrix_a.T)
    r = m.matrix_a @ m.classical_solution  # ToyModel.normalize_model scales A, x and b by 1/|b|: A x is parallel to b
    assert np.allclose(r / np.linalg.norm(r), m.vector_b / np.linalg.norm(m.vector_b))
    s = toymodels.ScalingTestModel(matrix_size=4, matrix_s=2, matrix_well_conditioned=True, vector_uniformity=1.0)
    assert s.matrix_a.shape == (8, 8) and np.allclose(s.matrix_a, s.matrix_a.conj().T)  # Hermitian embedding
    assert s.name.startswith("TestNxN_n=4_s=2_c=") and s.classical_solution.shape == (4,)
    assert np.linalg.cond(s.matrix_a) <= 40


def test_integro_differential_matrix_blocks():
    a = np.array([[0.0, 1.0], [2.0, 3.0]])
    m = toymodels.integro_differential_a_matrix(a, 3)
    assert m.shape == (6, 6)
    assert np.allclose(m[:2, :2], np.eye(2)) and np.allclose(m[2:4, 2:4], np.eye(2))
    assert np.allclose(m[2:4, :2], -np.eye(2) - a / 3) and np.allclose(m[4:, 2:4], -np.eye(2) - a / 3)
    assert np.count_nonzero(m[:2, 2:]) == 0 and np.count_nonzero(m[4:, :2]) == 0


def test_csv_schema_and_cli(tmp_path, monkeypatch):
    """The harness writes the reference's columns; the CLI parses the reference's flags."""

    def fake_solve(self, matrix_a, vector_b, method, file_basename, **kwargs):
        self.solution = np.linalg.solve(matrix_a, vector_b)
        self.qasm_circuit, self.circuit_depth, self.circuit_width, self.run_time = "OPENQASM 3;", 7, 5, 0.25
        self.name, self.method = file_basename, method
        return self.solution

    monkeypatch.setattr(compare.QuantumLinearSolver, "solve", fake_solve)
    monkeypatch.chdir(tmp_path)
    path = compare.compare_qls([toymodels.Qiskit4QubitExample(), toymodels.ClassiqDemoExample()], filebasename="cmp")
    rows = list(csv.reader(open(path, encoding="utf-8")))
    assert rows[0] == compare.CSV_HEADERS and len(rows) == 3
    assert rows[1][0] == "Qiskit4QubitExample" and rows[1][1] == "hhl_qiskit" and float(rows[1][6]) < 1e-9

    np.savetxt(tmp_path / "a.csv", np.array([[1, -1 / 3], [-1 / 3, 1]]), delimiter=",")
    np.savetxt(tmp_path / "b.csv", np.array([1.0, 0.0]), delimiter=",")
    (tmp_path / "args.yaml").write_text("show_circuit: false\n", encoding="utf-8")
    args = execute_framework.parse_arguments(["-i", "hhl_qiskit", "-m", "a.csv", "-v", "b.csv", "-iargs", "args.yaml", "-y"])
    assert args.matrix_a.shape == (2, 2) and args.implementation_args == {"show_circuit": False}
    with pytest.raises(ValueError):
        execute_framework.parse_arguments(["-i", "nope"])
    monkeypatch.setattr(execute_framework.QuantumLinearSolver, "solve", fake_solve)
    qls = execute_framework.main(["-i", "hhl_qiskit", "-m", "a.csv", "-v", "b.csv", "-y"])
    assert np.allclose(qls.solution, [1.125, 0.375])
