"""Specialised (NVRTC-compiled) pass kernels on the GPU: the producer-warp / bulk-copy / two-group kernel of
csrc/rs2_kernel.cuh around the generated per-tile bodies.  Parity against the oracle at sizes it reaches, and
against the interpreter kernels (a different code path through the same lowered program) beyond that.

Tolerances (BASELINE.json north_star): |delta| <= 1e-12 per amplitude in fp64, fidelity >= 1 - 1e-10."""
import numpy as np
import pytest

from oracle import numpy_sv

from _util import circuit_to_stream, random_circuit, random_unitary

pytestmark = pytest.mark.gpu

TOL64, FID = 1e-12, 1 - 1e-10


def _run(qc, psi0=None, jit_on=True, index_offset=0, n_local=None, zero_start=None):
    from qls_b200 import _abi
    from qls_b200.engine import DeviceProgram, DeviceState
    from qls_b200.lower import LoweringOptions, lower_circuit

    n = qc.num_qubits
    opts = LoweringOptions(n_local=n_local or 0, zero_start=(psi0 is None) if zero_start is None else zero_start)
    low = lower_circuit(qc, opts)
    prog = DeviceProgram(low)
    st = DeviceState(n_local or n, "fp64", index_offset=index_offset)
    _abi.check(_abi.load().qls_set_use_jit(1 if jit_on else 0))
    try:
        if psi0 is None:
            prog.run(st)
        else:
            st.write(0, psi0)
            prog.run(st, initialise=False)
        out = st.read()
    finally:
        _abi.check(_abi.load().qls_set_use_jit(1))
    return out, prog


def _fidelity(a, b):
    return abs(np.vdot(a, b)) ** 2 / (np.vdot(a, a).real * np.vdot(b, b).real)


@pytest.mark.parametrize("n", [20, 21])
def test_hhl_shaped_matches_oracle(cuda_device, n):
    from qls_b200.synthetic import hhl_shaped_circuit

    qc = hhl_shaped_circuit(n)
    want = numpy_sv.run(n, circuit_to_stream(qc))
    got, prog = _run(qc)
    assert prog.jit_passes > 0
    assert np.max(np.abs(got - want)) <= TOL64
    assert _fidelity(got, want) >= FID


@pytest.mark.parametrize("seed", range(3))
def test_random_circuits_match_oracle(cuda_device, seed):
    rng = np.random.default_rng(5000 + seed)
    n = 20
    qc = random_circuit(n, 70, rng, max_k=2)
    want = numpy_sv.run(n, circuit_to_stream(qc))
    got, prog = _run(qc)
    assert prog.jit_passes > 0
    assert np.max(np.abs(got - want)) <= TOL64
    assert _fidelity(got, want) >= FID


def test_idle_tiles_and_lazy_stages(cuda_device):
    """Every block is controlled from outside the tile: idle tiles are neither loaded nor stored."""
    from qls_b200.circuit import Gate, QuantumCircuit

    n = 20
    rng = np.random.default_rng(31)
    qc = QuantumCircuit(n)
    for ctl, tg in ((18, (0, 5)), (19, (3, 9)), (18, (1, 2)), (19, (10, 11)), (17, (4, 6)), (19, (0, 7)), (18, (8, 9))):
        qc.append(Gate("unitary", 2, [], random_unitary(2, rng)).control(1), [ctl] + list(tg))
    qc.cp(0.7, 19, 3)
    psi0 = rng.standard_normal(2**n) + 1j * rng.standard_normal(2**n)
    psi0 /= np.linalg.norm(psi0)
    want = numpy_sv.run(n, circuit_to_stream(qc), psi0=psi0.copy())
    got, prog = _run(qc, psi0=psi0)
    assert prog.jit_passes > 0
    assert np.max(np.abs(got - want)) <= TOL64


def test_sharded_index_offset(cuda_device):
    """Both shards of a 21-qubit state, one after the other on one GPU: rank bits reach controls and table
    indices through index_offset."""
    from qls_b200.circuit import Gate, QuantumCircuit

    n, n_local = 21, 20
    rng = np.random.default_rng(77)
    qc = QuantumCircuit(n)
    qc.h(0)
    qc.cp(0.37, n - 1, 2)
    qc.append(Gate("unitary", 2, [], random_unitary(2, rng)).control(1), [n - 1, 5, 1])
    qc.h(5)
    qc.cp(-1.1, 5, n - 1)
    for q in range(6, 18):
        qc.h(q)
        qc.cp(0.3 * q, n - 1, q)
    qc.append(Gate("unitary", 2, [], random_unitary(2, rng)).control(1, 0), [n - 1, 12, 3])
    psi0 = rng.standard_normal(2**n) + 1j * rng.standard_normal(2**n)
    psi0 /= np.linalg.norm(psi0)
    want = numpy_sv.run(n, circuit_to_stream(qc), psi0=psi0.copy())
    for rank in range(2):
        sl = slice(rank << n_local, (rank + 1) << n_local)
        got, prog = _run(qc, psi0=psi0[sl].copy(), index_offset=rank << n_local, n_local=n_local)
        assert prog.jit_passes > 0
        assert np.max(np.abs(got - want[sl])) <= TOL64


def test_large_state_matches_interpreter_kernels(cuda_device):
    """n = 27 (2 GiB): the specialised kernels against the interpreter kernels on the same lowered program --
    fidelity, max |delta| over sampled slices, P(flag = 1) and the solution-slice norm."""
    from qls_b200 import _abi
    from qls_b200.engine import DeviceProgram, DeviceState
    from qls_b200.lower import LoweringOptions, lower_circuit
    from qls_b200.synthetic import hhl_shaped_circuit

    n = 27
    qc = hhl_shaped_circuit(n)
    prog = DeviceProgram(lower_circuit(qc, LoweringOptions()))
    assert prog.jit_passes == prog.lowered.n_passes
    a, b = DeviceState(n), DeviceState(n)
    lib = _abi.load()
    prog.run(a)
    _abi.check(lib.qls_set_use_jit(0))
    try:
        prog.run(b)
    finally:
        _abi.check(lib.qls_set_use_jit(1))
    assert abs(a.norm_sq_range(0, 2**n) - 1.0) < 1e-11
    ov = a.inner(b)
    assert abs(abs(ov) ** 2 - 1.0) < 1e-10
    half = 2 ** (n - 1)
    assert abs(a.prob_mask(half, half) - b.prob_mask(half, half)) < 1e-12
    nb = qc.qregs[0].size
    assert abs(a.norm_sq_range(half, 2**nb) - b.norm_sq_range(half, 2**nb)) < 1e-13
    rng = np.random.default_rng(5)
    for off in [0, half] + [int(x) for x in rng.integers(0, 2**n - 2**16, 6)]:
        assert np.max(np.abs(a.read(off, 2**16) - b.read(off, 2**16))) <= TOL64


@pytest.mark.parametrize("jit_on", [True, False])
def test_multiplexer_over_seventeen_selectors(cuda_device, jit_on):
    """More than 16 selector qubits (ExactReciprocal of a sharded 34-36 qubit solve): register-stage form on both the
    specialised and the interpreter kernels; expected state by direct formula (tests/test_jit_codegen.py has the CPU
    version of this test)."""
    from qls_b200.circuit import QuantumCircuit, UCRYGate

    rng = np.random.default_rng(3)
    n, ns = 20, 17
    ang = rng.uniform(-3, 3, 2**ns)
    front, back = QuantumCircuit(n), QuantumCircuit(n)
    for q in range(n):
        front.ry(float(rng.uniform(-2, 2)), q)
    front.cx(2, 17), front.cp(0.3, 16, 4)
    back.h(3), back.cx(3, 17), back.rz(0.7, 9), back.cx(17, 19)
    qc = QuantumCircuit(n)
    qc.compose(front, list(range(n)), inplace=True)
    qc.append(UCRYGate(ang), [17] + list(range(ns)))
    qc.compose(back, list(range(n)), inplace=True)
    psi = numpy_sv.run(n, circuit_to_stream(front)).reshape(4, 2, 2**ns)  # [qubits 18-19, target, selectors]
    c, s = np.cos(ang / 2), np.sin(ang / 2)
    psi = np.stack([c * psi[:, 0] - s * psi[:, 1], s * psi[:, 0] + c * psi[:, 1]], axis=1).reshape(-1)
    want = numpy_sv.run(n, circuit_to_stream(back), psi0=psi)
    got, prog = _run(qc, jit_on=jit_on)
    assert all(p.rs_stages > 0 for p in prog.lowered.pass_info)
    assert np.max(np.abs(got - want)) <= TOL64
