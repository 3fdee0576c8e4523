"""Parity of the CUDA path (through the C ABI) against the oracle.  Run on the B200 box:
``python -m pytest tests -m gpu``.

Tolerances (BASELINE.json north_star): per-amplitude |delta| <= 1e-12 in fp64 (1e-5 in fp32) and
state fidelity >= 1 - 1e-10."""
import math
import os

import numpy as np
import pytest

from oracle import hhl_analytic, numpy_sv

from _util import circuit_to_stream, random_circuit, random_unitary

pytestmark = pytest.mark.gpu

TOL64, TOL32, FID = 1e-12, 1e-5, 1 - 1e-10
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _fidelity(a, b):
    return abs(np.vdot(a, b)) ** 2 / (np.vdot(a, a).real * np.vdot(b, b).real)


def _gpu_state(qc, precision="fp64", **opt):
    from qls_b200.lower import LoweringOptions
    from qls_b200.statevector import Statevector

    options = LoweringOptions(dtype=precision, **opt) if opt else None
    return Statevector(qc, precision=precision, options=options)


def test_library_loads_and_reports_b200(cuda_device):
    import ctypes as C

    from qls_b200 import _abi

    lib = _abi.load()
    sm, smem, tot, fr = C.c_int(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    _abi.check(lib.qls_device_info(0, C.byref(sm), C.byref(smem), C.byref(tot), C.byref(fr)))
    assert sm.value > 0 and smem.value >= 64 * 1024


@pytest.mark.parametrize("seed", range(10))
def test_random_circuits_fp64(cuda_device, seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(3, 15))
    qc = random_circuit(n, 60, rng, max_k=5)
    want = numpy_sv.run(n, circuit_to_stream(qc))
    got = _gpu_state(qc).data
    assert np.max(np.abs(got - want)) <= TOL64
    assert _fidelity(got, want) >= FID


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("tile,high,fuse", [(5, 2, 1), (8, 3, 3), (12, 5, 4), (12, 5, 5), (10, 6, 2)])
def test_random_circuits_tile_shapes(cuda_device, seed, tile, high, fuse):
    rng = np.random.default_rng(2000 + seed)
    n = int(rng.integers(6, 15))
    qc = random_circuit(n, 50, rng, max_k=min(5, high + 1))
    want = numpy_sv.run(n, circuit_to_stream(qc))
    got = _gpu_state(qc, tile_bits=tile, max_high=high, max_fuse_qubits=fuse).data
    assert np.max(np.abs(got - want)) <= TOL64


@pytest.mark.parametrize("tile,high,fuse", [(6, 3, 2), (13, 5, 5), (13, 6, 3), (12, 5, 4)])
def test_random_circuits_tile_shapes_fp32(cuda_device, tile, high, fuse):
    rng = np.random.default_rng(2500 + tile + high)
    n = 15
    qc = random_circuit(n, 50, rng, max_k=min(5, high))
    want = numpy_sv.run(n, circuit_to_stream(qc))
    got = _gpu_state(qc, precision="fp32", tile_bits=tile, max_high=high, max_fuse_qubits=fuse).data
    assert np.max(np.abs(got - want)) <= TOL32


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_k5_dense_blocks_in_large_tiles(cuda_device, precision):
    """k = 5 dense blocks use two threads per amplitude group; cover tiles with more groups than
    one sweep of the CTA handles (13-bit tiles) and controls inside / outside the tile."""
    from qls_b200.circuit import Gate, QuantumCircuit

    rng = np.random.default_rng(31)
    n = 16
    qc = QuantumCircuit(n)
    for q in range(n):
        qc.h(q)
    for _ in range(6):
        qs = [int(q) for q in rng.choice(n, 6, replace=False)]
        qc.append(Gate("unitary", 5, [], random_unitary(5, rng)).control(1, int(rng.integers(2))), qs)
        qc.unitary(random_unitary(5, rng), [int(q) for q in rng.choice(n, 5, replace=False)])
    want = numpy_sv.run(n, circuit_to_stream(qc))
    for tile in (12, 13):
        got = _gpu_state(qc, precision=precision, tile_bits=tile, max_high=5).data
        assert np.max(np.abs(got - want)) <= (TOL64 if precision == "fp64" else TOL32)


@pytest.mark.parametrize("seed", range(4))
def test_random_circuits_fp32(cuda_device, seed):
    rng = np.random.default_rng(3000 + seed)
    n = int(rng.integers(3, 15))
    qc = random_circuit(n, 40, rng, max_k=4)
    want = numpy_sv.run(n, circuit_to_stream(qc))
    got = _gpu_state(qc, precision="fp32").data
    assert np.max(np.abs(got - want)) <= TOL32
    assert _fidelity(got, want) >= 1 - 1e-5


def test_random10_golden(cuda_device):
    from qls_b200.statevector import Statevector

    qc = random_circuit(10, 80, np.random.default_rng(424242), max_k=4)
    want = np.load(os.path.join(GOLDEN, "random10_state.npy"))
    assert np.max(np.abs(Statevector(qc).data - want)) <= TOL64


def test_fused_equals_unfused(cuda_device):
    from qls_b200.statevector import Statevector

    rng = np.random.default_rng(5)
    qc = random_circuit(13, 120, rng, max_k=3)
    a = Statevector(qc, fused=True).data
    b = Statevector(qc, fused=False).data
    assert np.max(np.abs(a - b)) <= 1e-13


def test_kat_hhl_2x2_through_the_facade(cuda_device, tmp_path, monkeypatch):
    """Reference tests/test_hhl_implementations.py:23-29: width 5, x = [1.125, 0.375]."""
    from qls_b200.quantum_linear_solver import QuantumLinearSolver
    from qls_b200.toymodels import Qiskit4QubitExample

    monkeypatch.chdir(tmp_path)
    model = Qiskit4QubitExample()
    qls = QuantumLinearSolver()
    x = qls.solve(model.matrix_a, model.vector_b, method="hhl_qiskit", file_basename="kat")
    assert qls.circuit_width == 5
    assert np.allclose(model.classical_solution, x)
    assert np.max(np.abs(x - np.array([1.125, 0.375]))) < 1e-12
    assert os.path.exists("hhl_qiskit_circuit.qasm3")  # reference side effect, utils.py:12-13
    assert qls.qasm_circuit.startswith("OPENQASM 3")
    assert qls.run_time > 0 and qls.circuit_depth > 0


def test_kat_state_matches_appendix_b(cuda_device):
    from qls_b200.hhl import HHL
    from qls_b200.statevector import Statevector
    from qls_b200.toymodels import Qiskit4QubitExample

    model = Qiskit4QubitExample()
    qc = HHL(power_mode="repeat").construct_circuit(model.matrix_a, model.vector_b)
    got = Statevector(qc).data
    want = np.load(os.path.join(GOLDEN, "hhl_2x2_state.npy"))
    assert np.max(np.abs(got - want)) <= TOL64


@pytest.mark.parametrize("name", ["classiq_demo", "volterra2"])
def test_hhl_toy_models_vs_golden(cuda_device, name):
    from qls_b200.hhl import HHL
    from qls_b200.statevector import Statevector
    from qls_b200.toymodels import ClassiqDemoExample, VolterraProblem

    model = ClassiqDemoExample() if name == "classiq_demo" else VolterraProblem(2)
    hhl = HHL()
    res = hhl.solve(model.matrix_a, model.vector_b)
    got = Statevector(res.state).data
    want = np.load(os.path.join(GOLDEN, f"hhl_{name}_state.npy"))
    assert np.max(np.abs(got - want)) <= TOL64
    assert _fidelity(got, want) >= FID
    n = model.matrix_a.shape[0]
    assert math.isclose(res.euclidean_norm, numpy_sv.hhl_norm(want, n, hhl.scaling), rel_tol=1e-12)


@pytest.mark.parametrize("in_place", [False, True])
@pytest.mark.parametrize("nb,n,ctrl", [(6, 22, (21,)), (7, 20, (9, 15)), (10, 22, ())])
def test_ctrl_gemm_direct(cuda_device, nb, n, ctrl, in_place):
    """qls_ctrl_gemm through the C ABI on a random state: out of place through the bounded scratch (nb = 6, n = 22: 2^15
    selected rows = two chunks of QLS_GEMM_CHUNK_ROWS) and in place (scratch = NULL: thread-block clusters of 1, 2 and 8
    CTAs), against rows @ U^T in NumPy."""
    import torch

    from qls_b200 import _abi
    from qls_b200.engine import DeviceState

    rng = np.random.default_rng(nb + n)
    psi = rng.standard_normal(2**n) + 1j * rng.standard_normal(2**n)
    psi /= np.linalg.norm(psi)
    u = random_unitary(nb, rng)
    mask = sum(1 << q for q in ctrl)
    st = DeviceState(n, "fp64")
    st.write(0, psi)
    u_dev = torch.from_numpy(np.ascontiguousarray(u)).cuda()
    rows_sel = 2 ** (n - nb - len(ctrl))
    scratch = None if in_place else torch.empty(min(rows_sel, _abi.QLS_GEMM_CHUNK_ROWS) * 2**nb, dtype=torch.complex128, device="cuda")
    _abi.check(st.lib.qls_ctrl_gemm(st.ptr, None if in_place else scratch.data_ptr(), n, st.abi_dtype, nb, mask, mask, 0,
                                    u_dev.data_ptr(), int(torch.cuda.current_stream().cuda_stream)))
    got = st.read()
    want = psi.reshape(-1, 2**nb).copy()
    rows = np.arange(2 ** (n - nb), dtype=np.int64) << nb
    sel = (rows & mask) == mask
    want[sel] = want[sel] @ u.T
    assert np.max(np.abs(got - want.reshape(-1))) <= 1e-14


@pytest.mark.parametrize("in_place", [False, True])
@pytest.mark.parametrize("nb", [6, 10])
def test_hhl_gemm_block_path(cuda_device, nb, in_place, monkeypatch):
    """Dense controlled-e^{iAt} blocks on the data register go through qls_ctrl_gemm.  nb = 10 is BASELINE.json
    configs[1] at its real size (tridiagonal N = 2**10, reference formula nl = 12, n = 23: 24 complex 1024x1024
    products on 2**22 rows each) against the closed-form state, at the north-star tolerance."""
    from qls_b200.hhl import HHL
    from qls_b200.lower import lower_circuit
    from qls_b200.statevector import Statevector
    from qls_b200.toymodels import TridiagonalToeplitz

    monkeypatch.setenv("QLS_GEMM_INPLACE", "1" if in_place else "0")  # both forms of qls_ctrl_gemm (csrc/ctrl_gemm.cu)
    model = TridiagonalToeplitz(nb)
    hhl = HHL(power_mode="eig")
    qc = hhl.construct_circuit(model.matrix_a, model.vector_b)
    if nb == 10:
        assert (qc.num_qubits, hhl.nl) == (23, 12)
    assert sum(s.kind == "gemm" for s in lower_circuit(qc).steps) == 2 * hhl.nl
    sv = Statevector(qc)
    got = sv.data
    want = hhl_analytic.hhl_state(model.matrix_a, model.vector_b)
    assert np.max(np.abs(got - want)) <= TOL64
    assert _fidelity(got, want) >= FID
    half = 2 ** (qc.num_qubits - 1)
    assert abs(sv.norm_sq_range(half, 2**nb) - float(np.sum(np.abs(want[half:half + 2**nb]) ** 2))) <= 1e-13


def test_reductions(cuda_device):
    from qls_b200.statevector import Statevector

    rng = np.random.default_rng(77)
    n = 15
    psi = rng.standard_normal(2**n) + 1j * rng.standard_normal(2**n)
    psi /= np.linalg.norm(psi)
    phi = rng.standard_normal(2**n) + 1j * rng.standard_normal(2**n)
    phi /= np.linalg.norm(phi)
    sv, sw = Statevector(psi), Statevector(phi)
    assert np.array_equal(sv.data, psi)  # write/read round trip is bit exact
    assert abs(sv.norm_sq_range(0, 2**n) - 1.0) < 1e-13
    assert abs(sv.norm_sq_range(2 ** (n - 1), 64) - np.sum(np.abs(psi[2 ** (n - 1):2 ** (n - 1) + 64]) ** 2)) < 1e-15
    assert sv.norm_sq_range(5, 0) == 0.0
    for q in (0, 7, n - 1):
        assert abs(sv.expectation_z(q) - numpy_sv.expectation_z(psi, q)) < 1e-13
    assert abs(sv.expectation_value("I" * (n - 1) + "Z") - numpy_sv.expectation_z(psi, 0)) < 1e-13
    mask, val = (1 << (n - 1)) | 0b110, (1 << (n - 1)) | 0b010
    assert abs(sv.probability(mask, val) - numpy_sv.prob_mask(psi, mask, val)) < 1e-13
    assert abs(sv.inner(sw) - np.vdot(phi, psi)) < 1e-13


def test_hhl_shaped_vs_oracle(cuda_device):
    from qls_b200.statevector import Statevector
    from qls_b200.synthetic import hhl_shaped_circuit

    n = 17
    qc = hhl_shaped_circuit(n)
    want = numpy_sv.run(n, circuit_to_stream(qc))
    got = Statevector(qc).data
    assert np.max(np.abs(got - want)) <= TOL64
    assert _fidelity(got, want) >= FID


@pytest.mark.parametrize("precision,tol", [("fp64", 1e-11), ("fp32", 2e-4)])
def test_large_state_properties(cuda_device, precision, tol):
    """Size-independent checks at a size the oracle cannot reach: unit norm, P(flag) consistency,
    and U^dagger U = identity (evolve, then evolve the inverse circuit, compare with the input)."""
    from qls_b200.engine import DeviceProgram, DeviceState
    from qls_b200.lower import LoweringOptions, lower_circuit
    from qls_b200.synthetic import hhl_shaped_circuit

    n = 27
    qc = hhl_shaped_circuit(n)
    opts = LoweringOptions(dtype=precision)
    fwd = DeviceProgram(lower_circuit(qc, opts))
    # inverse of everything after the state preparation (whose inverse is not an initialisation)
    from qls_b200.circuit import Instruction, QuantumCircuit
    inv_circ = QuantumCircuit(*qc.qregs, name="inv")
    for ins in reversed(qc.data[1:]):
        inv_circ.data.append(Instruction(ins.operation.inverse(), ins.qubits))
    bwd = DeviceProgram(lower_circuit(inv_circ, LoweringOptions(dtype=precision)))
    st = DeviceState(n, precision)
    fwd.run(st)
    assert abs(st.norm_sq_range(0, 2**n) - 1.0) < tol
    p1 = st.prob_mask(1 << (n - 1), 1 << (n - 1))
    p0 = st.prob_mask(1 << (n - 1), 0)
    assert abs(p0 + p1 - 1.0) < tol and 0.0 < p1 < 1.0
    assert abs(st.expectation_z(n - 1) - (p0 - p1)) < tol
    bwd.run(st, initialise=False)
    nb = qc.qregs[0].size
    rng = np.random.default_rng(99)
    b = rng.standard_normal(2**nb)
    b /= np.linalg.norm(b)
    back = st.read(0, 2**nb)
    assert np.max(np.abs(back - b)) < tol * 10
    assert abs(st.norm_sq_range(0, 2**nb) - 1.0) < tol * 10


# ---- VQLS path: batched estimator kernel ---------------------------------------------------------


@pytest.mark.parametrize("nq,precision", [(1, "fp64"), (2, "fp64"), (3, "fp64"), (3, "fp32")])
def test_estimator_matches_oracle_on_hadamard_tests(cuda_device, nq, precision):
    """Estimator.run values (reference call site vqls_qiskit_implementation.py:55-62) vs bind +
    oracle + <Z_0>, for every Hadamard-test circuit VQLS builds, 8 parameter vectors per launch."""
    from qls_b200.estimator import B200Estimator
    from qls_b200.library import RealAmplitudes
    from qls_b200.vqls import VQLS

    rng = np.random.default_rng(40 + nq)
    dim = 2**nq
    a = rng.standard_normal((dim, dim))
    a = a + a.T
    b = rng.standard_normal(dim)
    ansatz = RealAmplitudes(nq, reps=3)
    vq = VQLS(None, ansatz, None)
    norm_t, over_t = vq.construct_circuit(a, b)
    est = B200Estimator(precision=precision)
    thetas = rng.uniform(-np.pi, np.pi, (8, ansatz.num_parameters))
    tol = 1e-12 if precision == "fp64" else 1e-5
    for test in norm_t + over_t:
        for circ in test.circuits:
            got = est.run([circ] * 8, [("Z", 0)] * 8, list(thetas)).result().values
            for k in range(8):
                psi = numpy_sv.run(circ.num_qubits, circuit_to_stream(circ.assign_parameters(thetas[k])))
                assert abs(got[k] - numpy_sv.expectation_z(psi, 0)) <= tol


def test_estimator_wide_operator_matvec(cuda_device):
    """config 3 shape: ansatz on 9 data qubits, controlled dense 9-qubit unitary, 10 qubits total."""
    from qls_b200.circuit import QuantumCircuit
    from qls_b200.estimator import B200Estimator
    from qls_b200.library import RealAmplitudes
    from qls_b200.vqls import HadamardTest

    rng = np.random.default_rng(7)
    nq = 9
    ansatz = RealAmplitudes(nq, reps=3)
    op = QuantumCircuit(nq)
    op.unitary(random_unitary(nq, rng), list(range(nq)))
    test = HadamardTest([op], apply_initial_state=ansatz)
    thetas = rng.uniform(-np.pi, np.pi, (5, ansatz.num_parameters))
    est = B200Estimator()
    for circ in test.circuits:
        got = est.run([circ] * 5, ["I" * nq + "Z"] * 5, list(thetas)).result().values
        for k in range(5):
            psi = numpy_sv.run(nq + 1, circuit_to_stream(circ.assign_parameters(thetas[k])))
            assert abs(got[k] - numpy_sv.expectation_z(psi, 0)) <= 1e-12


def test_estimator_full_batch_of_4096(cuda_device):
    """BASELINE.json configs[2] at its real size: ONE launch of 4096 Hadamard-test circuits at 10 qubits (ansatz on 9
    data qubits, controlled dense 9-qubit unitary), real and imaginary part; 32 circuits sampled across the batch are
    checked against bind + oracle + <Z_0>."""
    from qls_b200.circuit import QuantumCircuit
    from qls_b200.estimator import B200Estimator
    from qls_b200.library import RealAmplitudes
    from qls_b200.vqls import HadamardTest

    rng = np.random.default_rng(7)
    nq, k = 9, 4096
    ansatz = RealAmplitudes(nq, reps=3)
    op = QuantumCircuit(nq)
    op.unitary(random_unitary(nq, rng), list(range(nq)))
    test = HadamardTest([op], apply_initial_state=ansatz)
    thetas = np.random.default_rng(4321).uniform(-np.pi, np.pi, (k, ansatz.num_parameters))
    est = B200Estimator()
    picks = sorted(set(int(x) for x in np.random.default_rng(1).integers(0, k, 30)) | {0, k - 1})
    for circ in test.circuits:
        got = est.run([circ] * k, ["I" * nq + "Z"] * k, list(thetas)).result().values
        assert got.shape == (k,) and np.all(np.abs(got) <= 1 + 1e-12)
        for j in picks:
            psi = numpy_sv.run(nq + 1, circuit_to_stream(circ.assign_parameters(thetas[j])))
            assert abs(got[j] - numpy_sv.expectation_z(psi, 0)) <= 1e-12


def test_vqls_2x2_through_the_facade(cuda_device, tmp_path, monkeypatch):
    """Reference tests/test_vqls_implementations.py:26-37: width 1, atol 1e-3."""
    from qls_b200.library import RealAmplitudes
    from qls_b200.quantum_linear_solver import QuantumLinearSolver
    from qls_b200.toymodels import Qiskit4QubitExample

    monkeypatch.chdir(tmp_path)
    model = Qiskit4QubitExample()
    ansatz = RealAmplitudes(num_qubits=1, entanglement="full", reps=3, insert_barriers=False)
    qls = QuantumLinearSolver()
    x = qls.solve(model.matrix_a, model.vector_b, method="vqls_qiskit", file_basename="kat", ansatz=ansatz, seed=3)
    assert qls.circuit_width == 1
    assert np.allclose(model.classical_solution, x, atol=1e-3)
    assert os.path.exists("vqls_qiskit_circuit.qasm3")


def test_vqls_4x4_default_ansatz(cuda_device, tmp_path, monkeypatch):
    from qls_b200.implementations.vqls_qiskit_implementation import solve_vqls_qiskit
    from qls_b200.toymodels import ClassiqDemoExample

    monkeypatch.chdir(tmp_path)
    model = ClassiqDemoExample()
    x, qasm, depth, width, run_time = solve_vqls_qiskit(model.matrix_a, model.vector_b, optimizer_max_iter=400, seed=1)
    assert width == 2 and depth > 0 and qasm.startswith("OPENQASM 3")
    rel = np.linalg.norm(model.classical_solution - x) / np.linalg.norm(model.classical_solution)
    assert rel < 0.2  # the reference's own acceptance gate (plotting.py:136-143)


# ---- register-stage kernel (full-size tiles) --------------------------------------------------------


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_register_stage_kernel_random(cuda_device, seed, precision):
    from qls_b200 import _abi
    from qls_b200.lower import LoweringOptions, lower_circuit
    from qls_b200.statevector import Statevector

    rng = np.random.default_rng(7000 + seed)
    n = int(rng.integers(13, 18))
    qc = random_circuit(n, 80, rng, max_k=2)
    opts = LoweringOptions(dtype=precision, max_high=int(rng.integers(2, 6)))
    assert any(p.rs_stages > 0 for p in lower_circuit(qc, opts).pass_info)
    want = numpy_sv.run(n, circuit_to_stream(qc))
    tol = TOL64 if precision == "fp64" else TOL32
    got = Statevector(qc, precision=precision, options=opts).data
    assert np.max(np.abs(got - want)) <= tol
    lib = _abi.load()
    lib.qls_set_sweep_only(1)
    try:
        sweep = Statevector(qc, precision=precision, options=LoweringOptions(dtype=precision, max_high=opts.max_high)).data
    finally:
        lib.qls_set_sweep_only(0)
    assert np.max(np.abs(sweep - want)) <= tol
    assert np.max(np.abs(sweep - got)) <= (1e-13 if precision == "fp64" else 1e-5)


def test_register_stage_kernel_hhl_shaped_many_tiles(cuda_device):
    """n = 22: 1024 tiles per pass, persistent CTAs loop over several tiles each."""
    from qls_b200.statevector import Statevector
    from qls_b200.synthetic import hhl_shaped_circuit

    n = 22
    qc = hhl_shaped_circuit(n)
    want = numpy_sv.run(n, circuit_to_stream(qc))
    got = Statevector(qc).data
    assert np.max(np.abs(got - want)) <= TOL64
    assert _fidelity(got, want) >= FID


def _controlled_chain_circuit(n, n_ctl, rng, always_controlled):
    """2-qubit unitaries and diagonals on the twelve low qubits (all five non-contiguous tile slots
    are used up by qubits 7..11), each controlled by one of ``n_ctl`` high qubits: the controls are
    OUTSIDE the 12-bit tile, i.e. they select the step-list variant in csrc/rs_pass.cu.  More than 6
    distinct ones force the lowering to split the pass; without unconditional ops a tile whose
    controls are all off is idle (neither read nor written)."""
    from qls_b200.circuit import Gate, QuantumCircuit

    qc = QuantumCircuit(n, name="ctl_chain")
    low = 12
    for q in range(low if always_controlled else 0, n):
        qc.h(q)
    for j in range(3 * n_ctl):
        c = n - 1 - (j % n_ctl)
        a = 7 + j % 5  # every one of 7..11 is touched
        b = int(rng.integers(0, 7))
        qc.append(Gate("unitary", 2, [], random_unitary(2, rng)).control(1, int(rng.integers(2))), [c, a, b])
        if j % 3 == 0:
            d = Gate("diagonal", 2, [], np.diag(np.exp(1j * rng.uniform(-3, 3, 4)))).control(1)
            qc.append(d, [c, a, b])
        if not always_controlled and j % 5 == 4:
            qc.unitary(random_unitary(2, rng), [a, b])
    return qc


@pytest.mark.parametrize("n_ctl,always", [(3, True), (5, False), (8, True), (8, False)])
def test_register_stage_variants_external_controls(cuda_device, n_ctl, always):
    """Step-list variants: idle tiles (every op controlled), > 6 control qubits (pass split)."""
    from qls_b200.lower import LoweringOptions, lower_circuit
    from qls_b200.statevector import Statevector

    rng = np.random.default_rng(9100 + n_ctl + always)
    n = 21
    qc = _controlled_chain_circuit(n, n_ctl, rng, always)
    low = lower_circuit(qc, LoweringOptions(dtype="fp64"))
    assert any(p.rs_stages > 0 for p in low.pass_info)
    if always and n_ctl == 8:
        assert any(p.touched_fraction < 1.0 for p in low.pass_info)  # idle tiles exist
    if n_ctl > 6:
        assert low.n_passes >= 5  # split by the cap on control qubits outside the tile
    want = numpy_sv.run(n, circuit_to_stream(qc))
    got = Statevector(qc).data
    assert np.max(np.abs(got - want)) <= TOL64
    assert _fidelity(got, want) >= FID


def test_register_stage_small_tiles(cuda_device):
    """2^11-amplitude tiles (128 threads, four CTAs per SM): the other fp64 instantiation of the kernel."""
    from qls_b200.lower import LoweringOptions
    from qls_b200.statevector import Statevector
    from qls_b200.synthetic import hhl_shaped_circuit

    n = 19
    qc = hhl_shaped_circuit(n)
    want = numpy_sv.run(n, circuit_to_stream(qc))
    got = Statevector(qc, options=LoweringOptions(dtype="fp64", tile_bits=11)).data
    assert np.max(np.abs(got - want)) <= TOL64


@pytest.mark.parametrize("seed,n", [(0, 14), (1, 17), (2, 20), (3, 20)])
def test_zero_qubit_tracking_on_gpu(cuda_device, seed, n):
    """Statevector(circuit) starts from |0...0>: passes skip the tiles in which a qubit that has not been
    acted on yet is 1, and ops controlled by such qubits are resolved at lowering time (dead ops)."""
    from qls_b200.lower import LoweringOptions, lower_circuit
    from qls_b200.statevector import Statevector

    from test_lowering import _wake_up_circuit

    qc = _wake_up_circuit(n, np.random.default_rng(8800 + seed))
    assert any(p.fixed_bits > 0 for p in lower_circuit(qc, LoweringOptions(dtype="fp64", zero_start=True)).pass_info)
    want = numpy_sv.run(n, circuit_to_stream(qc))
    got = Statevector(qc).data
    assert np.max(np.abs(got - want)) <= TOL64
    got32 = Statevector(qc, precision="fp32").data
    assert np.max(np.abs(got32 - want)) <= TOL32


# ---- edge cases ------------------------------------------------------------------------------------


def test_edge_case_circuits_on_gpu(cuda_device):
    from qls_b200.statevector import Statevector

    from test_lowering import _edge_circuits

    for name, qc in _edge_circuits().items():
        want = numpy_sv.run(qc.num_qubits, circuit_to_stream(qc))
        for fused in (True, False):
            got = Statevector(qc, fused=fused).data
            assert np.max(np.abs(got - want)) <= 1e-14, (name, fused)
        got32 = Statevector(qc, precision="fp32").data
        assert np.max(np.abs(got32 - want)) <= TOL32, name


def test_unbound_parameters_are_rejected(cuda_device):
    from qls_b200.library import RealAmplitudes
    from qls_b200.statevector import Statevector

    with pytest.raises(ValueError):
        Statevector(RealAmplitudes(2, reps=1))


def test_library_error_codes_surface_as_exceptions(cuda_device):
    import ctypes as C

    from qls_b200 import _abi

    lib = _abi.load()
    with pytest.raises(_abi.QlsError):
        _abi.check(lib.qls_sv_init_basis(None, 4, 0, 0, None))
    out = (C.c_double * 2)()
    with pytest.raises(_abi.QlsError):
        _abi.check(lib.qls_reduce(None, 4, 0, 0, 0, 0, 0, None, None, out, None))
