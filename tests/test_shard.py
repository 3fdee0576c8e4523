"""Sharded execution on the CPU: the planner (layouts, remaps) and the runner's exchange protocol
are exercised with a NumPy stand-in for the local engine (``tests/_emulator.py`` interprets the
lowered local programs), first in one process with a fake communicator, then in two real processes
over ``torch.distributed`` gloo."""
import os
import sys

import numpy as np
import pytest

from oracle import numpy_sv
from qls_b200.lower import LoweringOptions
from qls_b200.shard import ShardedRunner, plan_sharded
from qls_b200.synthetic import hhl_shaped_circuit

from _emulator import run_pass, run_pass_rs
from _util import circuit_to_stream, random_circuit

OPTS = dict(tile_bits=5, max_high=2)


class NumpyShardEngine:
    """CPU stand-in for CudaShardEngine: a complex128 NumPy shard + the emulator."""

    def __init__(self, plan, rank):
        import torch

        self.torch = torch
        self.plan, self.rank = plan, rank
        self.n_local = plan.n_local
        self.state = np.zeros(2**plan.n_local, dtype=np.complex128)
        self.half = 2 ** (plan.n_local - 1)
        self.index_offset = rank << plan.n_local

    def run_local(self, lowered, initialise):
        steps = list(lowered.steps)
        if steps and steps[0].kind == "init":
            s = steps.pop(0)
            if initialise:
                self.state[:] = 0
                f = (1.0 if self.rank == 0 else 0.0) if s.rank_factors is None else s.rank_factors[self.rank]
                self.state[: len(s.amps)] = np.asarray(s.amps) * f
        elif initialise:
            self.state[:] = 0
            if self.rank == 0:
                self.state[0] = 1
        pool = lowered.pool.tobytes()
        for s in steps:
            assert s.kind == "passes"
            for pi in range(s.pass_begin, s.pass_end):
                run_pass(self.state, lowered.passes[pi], lowered.ops, pool, np.complex128, self.n_local, self.index_offset)

    def _block_index(self, lposs, value, begin, count):
        """Indices of the sub-block whose local bits lposs[m] (ascending) equal bit m of value."""
        i = np.arange(begin, begin + count, dtype=np.int64)
        for m, p in enumerate(lposs):
            assert m == 0 or p > lposs[m - 1]
            i = ((i >> p) << (p + 1)) | (i & ((1 << p) - 1))
        for m, p in enumerate(lposs):
            i |= ((value >> m) & 1) << p
        return i

    def pack(self, lposs, value, begin, count, slot=0):
        return self.torch.from_numpy(self.state[self._block_index(lposs, value, begin, count)].copy())

    def recv_buffer(self, count, slot=0):
        return self.torch.empty(count, dtype=self.torch.complex128)

    def unpack(self, buf, lposs, value, begin, count):
        self.state[self._block_index(lposs, value, begin, count)] = buf.numpy()


def _logical_state(plan, shards):
    """Undo the final layout: amplitude of logical index x sits at physical index P(x)."""
    n = plan.num_qubits
    phys_state = np.concatenate(shards)
    x = np.arange(2**n, dtype=np.int64)
    p = np.zeros_like(x)
    for q in range(n):
        p |= ((x >> q) & 1) << plan.final_layout[q]
    return phys_state[p]


class FakeDist:
    """Single-process stand-in for the torch.distributed calls the runner makes: ranks are run
    in lock step by the test, so a send just parks the tensor until the partner's recv."""

    class P2POp:
        def __init__(self, op, tensor, peer):
            self.op, self.tensor, self.peer = op, tensor, peer

    isend, irecv = "isend", "irecv"

    def __init__(self):
        self.mailbox = {}
        self.rank = 0

    def batch_isend_irecv(self, ops):
        class _Req:
            def wait(self_inner):
                return None

        for o in ops:
            if o.op == "isend":
                self.mailbox[(self.rank, o.peer)] = o.tensor.clone()
            else:
                self.pending_recv = (o.peer, o.tensor)
        return [_Req() for _ in ops]


def _run_fake(plan, world, chunk):
    """Lock-step execution of all ranks in one process."""
    engines = [NumpyShardEngine(plan, r) for r in range(world)]
    first = True
    for seg in plan.segments:
        if seg.kind == "local":
            for e in engines:
                e.run_local(seg.lowered, initialise=first)
            first = False
            continue
        # the k-qubit exchange of ShardedRunner.swap_many, all ranks in lock step
        pairs = sorted(((gpos - plan.n_local, lpos) for gpos, lpos in seg.pairs), key=lambda t: t[1])
        gbits, lposs = [t[0] for t in pairs], [t[1] for t in pairs]
        k = len(pairs)
        block = 2 ** (plan.n_local - k)
        for d in range(1, 2**k):
            def partner_of(r):
                p = r
                for j in range(k):
                    if (d >> j) & 1:
                        p ^= 1 << gbits[j]
                return p

            def value_of(p):
                return sum(((p >> gbits[j]) & 1) << j for j in range(k))

            for begin in range(0, block, chunk):
                count = min(chunk, block - begin)
                outs = {r: e.pack(lposs, value_of(partner_of(r)), begin, count) for r, e in enumerate(engines)}
                for r, e in enumerate(engines):
                    e.unpack(outs[partner_of(r)], lposs, value_of(partner_of(r)), begin, count)
    return engines


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("seed", range(3))
def test_plan_random_circuits_fake_comm(world, seed):
    rng = np.random.default_rng(50 + seed)
    n = 10
    qc = random_circuit(n, 60, rng, max_k=2)
    want = numpy_sv.run(n, circuit_to_stream(qc))
    plan = plan_sharded(qc, world, LoweringOptions(**OPTS), protect_low=2)
    engines = _run_fake(plan, world, chunk=37)
    got = _logical_state(plan, [e.state for e in engines])
    assert np.max(np.abs(got - want)) < 1e-12
    assert plan.n_swaps > 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_plan_hhl_shaped_fake_comm(world):
    n = 12
    qc = hhl_shaped_circuit(n)
    want = numpy_sv.run(n, circuit_to_stream(qc))
    plan = plan_sharded(qc, world, LoweringOptions(**OPTS), protect_low=2)
    engines = _run_fake(plan, world, chunk=1 << 20)
    got = _logical_state(plan, [e.state for e in engines])
    assert np.max(np.abs(got - want)) < 1e-12
    # shard invariance: same amplitudes as the unsharded program, up to rounding of identical arithmetic
    assert plan.initial_layout[: qc.qregs[0].size] == list(range(qc.qregs[0].size))  # data register stays low
    # controls / diagonals on global qubits must not have triggered remaps on their own
    assert plan.n_swaps <= 4 * int(np.log2(world)) + 2


@pytest.mark.parametrize("world", [2, 4, 8])
def test_first_gates_on_global_qubits_fold_into_the_state_preparation(world):
    """H on a clock qubit that starts on a global position is folded into the state preparation (every rank
    writes b / sqrt(2^g)): the flag stays local, and the result is the same as without the fold."""
    n = 13
    qc = hhl_shaped_circuit(n)
    want = numpy_sv.run(n, circuit_to_stream(qc))
    got = {}
    for fold in (True, False):
        plan = plan_sharded(qc, world, LoweringOptions(**OPTS), protect_low=2, fold_first_gates=fold)
        rf = plan.segments[0].lowered.steps[0].rank_factors
        assert (rf is not None) == fold
        if fold:
            g = int(np.log2(world))
            assert np.allclose(rf, [2 ** (-g / 2)] * world)  # Hadamards: every rank gets the same factor
            assert plan.initial_layout[n - 1] < plan.n_local  # the flag qubit is local
        engines = _run_fake(plan, world, chunk=1 << 20)
        got[fold] = _logical_state(plan, [e.state for e in engines])
        assert np.max(np.abs(got[fold] - want)) < 1e-12
    assert np.max(np.abs(got[True] - got[False])) < 1e-13


def _gloo_worker(rank, world, port, n, seed, out_dir):
    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        qc = hhl_shaped_circuit(n) if seed < 0 else random_circuit(n, 50, np.random.default_rng(seed), max_k=2)
        plan = plan_sharded(qc, world, LoweringOptions(**OPTS), protect_low=2)
        eng = NumpyShardEngine(plan, rank)
        runner = ShardedRunner(plan, eng, rank, world, chunk_amps=100, dist=dist)
        runner.run()
        np.save(os.path.join(out_dir, f"shard{rank}.npy"), eng.state)
        if rank == 0:
            np.save(os.path.join(out_dir, "layout.npy"), np.array(plan.final_layout))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("seed,world", [(-1, 2), (11, 2), (-1, 4), (12, 4)])
def test_multi_process_gloo_exchange(tmp_path, seed, world):
    """Real torch.distributed (gloo) runs of ShardedRunner: world 2 = single-qubit exchanges, world 4
    = two-qubit all-to-all exchanges (three pairwise rounds per exchange)."""
    import torch.multiprocessing as mp

    n = 11
    port = 29500 + (os.getpid() + (seed & 0xFF)) % 2000
    mp.spawn(_gloo_worker, args=(world, port, n, seed, str(tmp_path)), nprocs=world, join=True)
    qc = hhl_shaped_circuit(n) if seed < 0 else random_circuit(n, 50, np.random.default_rng(seed), max_k=2)
    want = numpy_sv.run(n, circuit_to_stream(qc))
    plan = plan_sharded(qc, world, LoweringOptions(**OPTS), protect_low=2)
    shards = [np.load(tmp_path / f"shard{r}.npy") for r in range(world)]
    assert list(np.load(tmp_path / "layout.npy")) == plan.final_layout
    got = _logical_state(plan, shards)
    assert np.max(np.abs(got - want)) < 1e-12


@pytest.mark.parametrize("world", [2, 4])
def test_gemm_register_is_pinned(world):
    """A GEMM block always executes on positions 0..nb-1 (csrc/ctrl_gemm.cu): the planner must never evict one of its
    target qubits, also when the register is wider than ``protect_low`` (nb = 9 > 8).  Without the pin, data qubit 8 is
    evicted during the inverse QFT and half of the GEMM blocks would act on the wrong qubits (the lowering now refuses
    such a block instead of emitting it)."""
    from qls_b200.hhl import HHL
    from qls_b200.toymodels import TridiagonalToeplitz

    model = TridiagonalToeplitz(9)
    hhl = HHL(power_mode="eig")
    qc = hhl.construct_circuit(model.matrix_a, model.vector_b)
    plan = plan_sharded(qc, world, LoweringOptions(dtype="fp64"))
    gemms = [s for seg in plan.segments if seg.kind == "local" for s in seg.lowered.steps if s.kind == "gemm"]
    assert len(gemms) == 2 * hhl.nl and all(s.nb == 9 for s in gemms)
    assert all(plan.final_layout[q] == q for q in range(9))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_deferred_final_gates_fold_into_the_slice_reduction(world):
    """Solver path: the final Hadamard of a clock qubit that is on a global position at the end is not applied; the
    clock = 0 slice is the rank-weighted sum of the local slices (``solution_slice_plan``).  Same slice as the full state of
    the oracle, fewer exchanged qubits than the full plan."""
    from qls_b200.shard import solution_slice_plan

    n = 13
    qc = hhl_shaped_circuit(n)
    nb = qc.qregs[0].size
    want = numpy_sv.run(n, circuit_to_stream(qc))
    full = plan_sharded(qc, world, LoweringOptions(**OPTS), protect_low=2)
    plan = plan_sharded(qc, world, LoweringOptions(**OPTS), protect_low=2, defer_final_1q=True)
    assert not full.deferred
    if world > 2:
        assert plan.deferred and plan.n_swaps < full.n_swaps
    assert all(plan.final_layout[q] >= plan.n_local and nb <= q < n - 1 for q in plan.deferred)  # clock qubits, global at the end
    engines = _run_fake(plan, world, chunk=1 << 20)
    x = np.zeros(2**nb, dtype=complex)
    p_flag = 0.0
    for r, e in enumerate(engines):
        sel, local, w = solution_slice_plan(plan, nb, r)
        x[sel] += w * e.state[local]
        gidx = np.arange(2**plan.n_local, dtype=np.int64) | (r << plan.n_local)
        p_flag += float(np.sum(np.abs(e.state[((gidx >> plan.final_layout[n - 1]) & 1) == 1]) ** 2))
    half = 2 ** (n - 1)
    assert np.max(np.abs(x - want[half:half + 2**nb])) < 1e-12
    assert abs(p_flag - float(np.sum(np.abs(want[half:]) ** 2))) < 1e-12  # P(flag = 1) does not see a unitary on the clock


def test_window_chunk_geometry_matches_the_gather_indices():
    """The pitched-copy description of a chunk (copy-engine exchange of consecutive local positions) addresses exactly the
    amplitudes the gather / scatter kernels' index deposit does."""
    from qls_b200.shard import window_chunk_geometry

    rng = np.random.default_rng(1)
    n_local, esz = 14, 16
    plan = plan_sharded(hhl_shaped_circuit(n_local + 1), 2, LoweringOptions(**OPTS), protect_low=2)
    eng = NumpyShardEngine(plan, 0)
    for _ in range(200):
        k = int(rng.integers(1, 4))
        p = int(rng.integers(0, n_local - k + 1))
        value = int(rng.integers(0, 2**k))
        block = 2 ** (n_local - k)
        count = 2 ** int(rng.integers(0, n_local - k + 1))
        begin = int(rng.integers(0, block // count)) * count
        off, pitch, width, height = window_chunk_geometry(p, k, value, begin, count, esz)
        got = np.concatenate([np.arange(width // esz) + (off + r * pitch) // esz for r in range(height)])
        want = eng._block_index(list(range(p, p + k)), value, begin, count)
        assert np.array_equal(got, want), (p, k, value, begin, count)


@pytest.mark.parametrize("world,n,chunk_log2", [(2, 34, 27), (4, 35, 27), (8, 36, 27), (8, 29, 20), (2, 23, 15)])
def test_overlap_plan_invariants(world, n, chunk_log2):
    """Host-side dry run of ShardedShapedSolver._plan_overlap (stand-ins for the device objects): the split positions are
    free in every overlapped pass and not exchanged, their element-index bits are at least log2(chunk), and the passes taken
    before / after neighbouring exchanges do not overlap inside one segment."""
    from qls_b200 import shard

    class Prog:
        def __init__(self, low):
            self.lowered, self.jit_attached = low, set(range(low.n_passes))

        def free_positions(self, i):
            p = self.lowered.passes[i]
            return [p.fseg[j].dst + w for j in range(p.n_fseg) for w in range(p.fseg[j].width)]

    class Slot:
        def numel(self):
            return 1 << chunk_log2

    class Runner:
        chunk = 1 << chunk_log2

        @staticmethod
        def ce_capable(lposs):
            ls = sorted(lposs)
            return all(b - a == 1 for a, b in zip(ls, ls[1:]))

    class Engine:
        pass

    plan = plan_sharded(hhl_shaped_circuit(n), world, LoweringOptions(dtype="fp64"), defer_final_1q=True)
    solver = shard.ShardedShapedSolver.__new__(shard.ShardedShapedSolver)
    eng = Engine()
    eng.peers, eng._recv = {1: 1}, [Slot()]
    eng.programs = {id(sg.lowered): Prog(sg.lowered) for sg in plan.segments if sg.kind == "local"}
    solver.plan, solver.engine, solver.runner = plan, eng, Runner()
    ov = solver._plan_overlap()
    assert ov, "the HHL-shaped plans have exchanges of consecutive positions with free positions beside them"
    taken = {}
    for j, (hs, ebits, pa, pb) in ov.items():
        seg = plan.segments[j]
        lposs = sorted(l for _, l in seg.pairs)
        assert 1 <= len(hs) <= 2 and len(set(hs)) == len(hs) and not set(hs) & set(lposs)
        for h, e in zip(hs, ebits):
            assert e == h - sum(1 for l in lposs if l < h) and (1 << e) >= (1 << chunk_log2)
        pa_prog, pb_prog = eng.programs[id(plan.segments[j - 1].lowered)], eng.programs[id(plan.segments[j + 1].lowered)]
        assert all(set(hs) <= set(pa_prog.free_positions(i)) for i in pa)
        assert all(set(hs) <= set(pb_prog.free_positions(i)) for i in pb)
        n_a = plan.segments[j - 1].lowered.n_passes
        assert pa == list(range(n_a - len(pa), n_a)) and pb == list(range(len(pb)))  # a suffix / a prefix
        for i in pa:
            assert (j - 1, i) not in taken
            taken[(j - 1, i)] = j
        for i in pb:
            assert (j + 1, i) not in taken
            taken[(j + 1, i)] = j
