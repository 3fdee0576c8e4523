"""Regenerate the golden statevectors under tests/golden/.

No part of the reference's qiskit stack can be imported in the build container (qiskit,
linear_solvers, vqls_prototype are absent and not installable), so these fixtures come from the
closed-form HHL state (``oracle/hhl_analytic.py``) -- for the 2x2 case that is exactly the vector
of SURVEY.md Appendix B, which the reference's own test pins (x = [1.125, 0.375], width 5) --
cross-checked against the gate-level NumPy oracle before writing.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import hhl_analytic, numpy_sv  # noqa: E402
from qls_b200.hhl import HHL  # noqa: E402
from qls_b200.toymodels import ClassiqDemoExample, Qiskit4QubitExample, VolterraProblem  # noqa: E402

from _util import circuit_to_stream, random_circuit  # noqa: E402


def main():
    for name, model in (("2x2", Qiskit4QubitExample()), ("classiq_demo", ClassiqDemoExample()),
                        ("volterra2", VolterraProblem(2))):
        ref = hhl_analytic.hhl_state(model.matrix_a, model.vector_b)
        qc = HHL(power_mode="eig").construct_circuit(model.matrix_a, model.vector_b)
        gate = numpy_sv.run(qc.num_qubits, circuit_to_stream(qc))
        assert np.max(np.abs(ref - gate)) < 1e-12, name
        np.save(os.path.join(HERE, f"hhl_{name}_state.npy"), ref)
    rng = np.random.default_rng(424242)
    qc = random_circuit(10, 80, rng, max_k=4)
    np.save(os.path.join(HERE, "random10_state.npy"), numpy_sv.run(10, circuit_to_stream(qc)))


if __name__ == "__main__":
    main()
