"""Specialised pass units (qls_b200/jit.py) on the CPU: the source generated for NVRTC is compiled for the
host (tests/_jit_emu.py) and the evolved state is compared with the oracle.  Covers the generated index
arithmetic -- stage layouts, landing / exchange images, table indices, controls inside and outside the
tile, idle tiles, zero-qubit tracking, sharded index offsets; the producer/consumer kernel around it is a
GPU test (tests/test_gpu_jit.py)."""
import re

import numpy as np
import pytest

from oracle import numpy_sv
from qls_b200 import jit
from qls_b200.lower import LoweringOptions, lower_circuit
from qls_b200.synthetic import hhl_shaped_circuit

from _emulator import run_program
from _util import circuit_to_stream, random_circuit


def _c_eval(expr: str, **env) -> int:
    return int(eval(re.sub(r"(0x[0-9a-f]+|\d+)(ull|u)\b", r"\1", expr), {}, env))


GEOMETRIES = [(4, (16, 17, 18, 19, 20, 21, 22, 23)), (4, (5, 8, 9, 10, 12, 13, 14, 15)), (4, (7, 8, 10, 11, 13, 14, 15, 31)),
              (5, (6, 8, 10, 11, 12, 13, 14)), (6, (7, 8, 10, 11, 12, 13)), (7, (9, 10, 13, 14, 15)), (8, (12, 13, 14, 15)),
              (9, (10, 13, 15)), (12, ())]


@pytest.mark.parametrize("mode", ["tma", "bulk"])
@pytest.mark.parametrize("L,high", GEOMETRIES)
def test_layout_expressions_match_their_definitions(L, high, mode):
    lay = jit.Layout(L, high, mode)
    rng = np.random.default_rng(L)
    sw = lay.swz_expr("x")
    helpers = {"qls_swz128": lambda u: u ^ ((u >> 3) & 7),
               "qls_pad_land": lambda t: t + (t >> L) + (t >> (L + 3)) + (t >> (L + 6)) if L < jit.T else t}
    ld = lay.land_expr_from("x", [(p, p) for p in range(jit.T)])
    for i in rng.integers(0, 2**jit.T, 200):
        i = int(i)
        assert _c_eval(sw, x=i) == lay.swz(i)
        assert _c_eval(ld, x=i, **helpers) == lay.land(i)
    # both images are injective on the tile and fit the buffer
    assert len({lay.swz(i) for i in range(2**jit.T)}) == 2**jit.T
    assert len({lay.land(i) for i in range(2**jit.T)}) == 2**jit.T
    assert max(lay.land(i) for i in range(2**jit.T)) < lay.land_units()
    if lay.mode == "tma":
        # the landing order is: low run, tensor-map dimensions (ascending), split groups -- and one copy per split value
        plan = lay.plan
        assert sorted(plan.perm) == list(range(jit.T)) and plan.perm[:plan.low_bits] == list(range(plan.low_bits))
        assert 1 <= plan.n_ops <= jit.MAX_TMA_OPS and len(plan.dims) <= 3
        assert sum(b for _, _, b in plan.dims + plan.split) + plan.low_bits == jit.T


@pytest.mark.parametrize("mode", ["tma", "bulk"])
@pytest.mark.parametrize("L,high", GEOMETRIES)
def test_quarter_warps_hit_eight_bank_groups(L, high, mode):
    """A 16-byte access is served per quarter warp; with the chosen thread-bit order the eight lanes fall
    into eight different 16-byte bank groups in the exchange image (always) and in the landing image (almost
    always: it needs free tile positions of three different landing classes)."""
    lay = jit.Layout(L, high, mode)
    rng = np.random.default_rng(10 + L)
    landing_ok = 0
    for _ in range(50):
        regq = sorted(int(q) for q in rng.choice(jit.T, jit.KR, replace=False))
        order = jit.choose_thread_order(regq, lay)
        assert sorted(order + regq) == list(range(jit.T))
        lanes = [sum(((t >> j) & 1) << order[j] for j in range(3)) for t in range(8)]
        if {lay.ex_cls(p) for p in order} == {0, 1, 2}:  # (register qubits may use up a whole class)
            assert len({lay.swz(i) & 7 for i in lanes}) == 8
        landing_ok += len({lay.land(i) & 7 for i in lanes}) == 8
    if lay.mode == "tma" or L < jit.T:  # (an unpadded single run spreads over positions 0..2 only)
        assert landing_ok >= 25


@pytest.mark.parametrize("seed", range(4))
def test_random_circuits_through_generated_code(seed, monkeypatch):
    monkeypatch.setenv("QLS_JIT_LANDING", "tma" if seed % 2 == 0 else "bulk")
    rng = np.random.default_rng(900 + seed)
    n = 13 + seed % 2
    qc = random_circuit(n, 45, rng, max_k=2)
    want = numpy_sv.run(n, circuit_to_stream(qc))
    low = lower_circuit(qc, LoweringOptions())
    assert any(jit.extract_pass(low, i) is not None for i in range(low.n_passes))
    got = run_program(low, form="jit")
    assert np.max(np.abs(got - want)) < 1e-12


def test_hhl_shaped_controls_outside_the_tile():
    """QPE blocks controlled by clock qubits outside the tile (uniform predicates, lazily entered stages),
    the multiplexed RY and zero-qubit tracking (pass-level fixed bits)."""
    n = 19
    qc = hhl_shaped_circuit(n)
    want = numpy_sv.run(n, circuit_to_stream(qc))
    low = lower_circuit(qc, LoweringOptions())
    assert any(p.fixed_bits > 0 for p in low.pass_info)
    srcs = [jit.generate_source(jit.extract_pass(low, i)) for i in range(low.n_passes) if jit.extract_pass(low, i)]
    assert any("const bool f1" in s for s in srcs), "no pass with controls outside the tile in this program"
    got = run_program(low, form="jit")
    assert np.max(np.abs(got - want)) < 1e-12
    assert abs(np.vdot(got, want)) ** 2 > 1 - 1e-10


def test_idle_tiles_are_passed_through():
    """Every block of the pass is controlled from outside the tile: tiles on which nothing fires are
    neither loaded nor stored (qls_tile_idle), the others enter only the stages that fire."""
    from qls_b200.circuit import Gate, QuantumCircuit

    from _util import random_unitary

    n = 15
    rng = np.random.default_rng(31)
    qc = QuantumCircuit(n)
    for ctl, tg in ((13, (0, 5)), (14, (3, 9)), (13, (1, 2)), (14, (10, 11)), (13, (4, 6)), (14, (0, 7)), (13, (8, 9))):
        qc.append(Gate("unitary", 2, [], random_unitary(2, rng)).control(1), [ctl] + list(tg))
    qc.cp(0.7, 14, 3)
    psi0 = rng.standard_normal(2**n) + 1j * rng.standard_normal(2**n)
    psi0 /= np.linalg.norm(psi0)
    want = numpy_sv.run(n, circuit_to_stream(qc), psi0=psi0.copy())
    low = lower_circuit(qc, LoweringOptions())
    srcs = [jit.generate_source(jit.extract_pass(low, i)) for i in range(low.n_passes) if jit.extract_pass(low, i)]
    assert any("return !(" in s for s in srcs), "no pass with idle tiles in this program"
    got = run_program(low, psi0=psi0.copy(), form="jit")
    assert np.max(np.abs(got - want)) < 1e-12


def test_sharded_index_offset():
    """Rank bits above n_local reach controls and table indices through the tile's global base index."""
    n, n_local = 15, 14
    rng = np.random.default_rng(77)
    # only diagonal / control use of the global qubit keeps the program local
    from qls_b200.circuit import QuantumCircuit

    from qls_b200.circuit import Gate

    from _util import random_unitary

    qc2 = QuantumCircuit(n)
    qc2.h(0)
    qc2.cp(0.37, n - 1, 2)
    qc2.append(Gate("unitary", 2, [], random_unitary(2, rng)).control(1), [n - 1, 5, 1])
    qc2.h(5)
    qc2.cp(-1.1, 5, n - 1)
    for q in range(6, 12):
        qc2.h(q)
        qc2.cp(0.3 * q, n - 1, q)
    qc2.append(Gate("unitary", 2, [], random_unitary(2, rng)).control(1, 0), [n - 1, 12, 3])
    psi0 = rng.standard_normal(2**n) + 1j * rng.standard_normal(2**n)
    psi0 /= np.linalg.norm(psi0)
    want = numpy_sv.run(n, circuit_to_stream(qc2), psi0=psi0.copy())
    low = lower_circuit(qc2, LoweringOptions(n_local=n_local))
    for rank in range(2):
        sl = slice(rank << n_local, (rank + 1) << n_local)
        a = run_program(low, psi0=psi0[sl].copy(), index_offset=rank << n_local, form="rs")
        b = run_program(low, psi0=psi0[sl].copy(), index_offset=rank << n_local, form="jit")
        assert np.max(np.abs(a - b)) < 1e-13
        assert np.max(np.abs(b - want[sl])) < 1e-12


def test_program_form_with_shared_gate_shapes(monkeypatch):
    """QLS_JIT_SHAPES=1: the 4x4 gates of a pass become steps of a tile program whose switch cases are the distinct
    gate shapes (an experiment kept for heavy passes; off by default)."""
    monkeypatch.setenv("QLS_JIT_SHAPES", "1")
    monkeypatch.setenv("QLS_JIT_SHARE_MIN", "0")
    n = 17
    qc = hhl_shaped_circuit(n)
    want = numpy_sv.run(n, circuit_to_stream(qc))
    low = lower_circuit(qc, LoweringOptions())
    srcs = [jit.generate_source(jit.extract_pass(low, i)) for i in range(low.n_passes) if jit.extract_pass(low, i)]
    assert any("qls_prog[" in s for s in srcs)
    got = run_program(low, form="jit")
    assert np.max(np.abs(got - want)) < 1e-12


def test_multiplexer_over_seventeen_selectors():
    """A multiplexed RY over MORE than 16 selector qubits (the ExactReciprocal of a sharded 34-36 qubit solve has 17-18
    clock qubits) keeps a register-stage form: selectors inside the tile index the low table bits.  Expected state by
    direct formula (the oracle's dense UCRY matrix would be 2^18 x 2^18)."""
    from qls_b200.circuit import QuantumCircuit, UCRYGate

    rng = np.random.default_rng(3)
    n, ns = 18, 17
    ang = rng.uniform(-3, 3, 2**ns)
    front, back = QuantumCircuit(n), QuantumCircuit(n)
    for q in range(n):
        front.ry(float(rng.uniform(-2, 2)), q)
    front.cx(2, 17), front.cp(0.3, 16, 4)
    back.h(3), back.cx(3, 17), back.rz(0.7, 9)
    qc = QuantumCircuit(n)
    qc.compose(front, list(range(n)), inplace=True)
    qc.append(UCRYGate(ang), [17] + list(range(ns)))  # qargs = [target, selector_0, ...]
    qc.compose(back, list(range(n)), inplace=True)

    psi = numpy_sv.run(n, circuit_to_stream(front)).reshape(2, 2**ns)  # [target, selectors]
    c, s = np.cos(ang / 2), np.sin(ang / 2)
    psi = np.stack([c * psi[0] - s * psi[1], s * psi[0] + c * psi[1]]).reshape(-1)
    want = numpy_sv.run(n, circuit_to_stream(back), psi0=psi)

    low = lower_circuit(qc, LoweringOptions(dtype="fp64"))
    assert all(p.rs_stages > 0 for p in low.pass_info)
    for form in ("rs", "jit"):
        got = run_program(low, form=form)
        assert np.max(np.abs(got - want)) <= 1e-12, form
