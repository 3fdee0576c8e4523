"""Sharded statevectors: states larger than one GPU's HBM (BASELINE.json configs[4]: 34/35/36
qubits fp64 on 2/4/8 B200).

Partition: rank = the top ``g = log2(world)`` *physical* index bits; every GPU holds a contiguous
shard of ``2**(n-g)`` amplitudes.  A *layout* maps logical qubits to physical positions and changes
over the circuit:

* blocks that touch (act non-diagonally on) local positions only run as ordinary local programs;
  controls, diagonal tables and RY selectors on global positions need no communication -- the
  kernels read the rank bits from ``index_offset`` (SURVEY.md section 8e "no-comm cases");
* when a block touches a global position the planner inserts a *remap*: global position ``G`` is
  swapped with a local position ``l`` (pairwise exchange of half a shard with rank ``r ^ bit``),
  choosing to bring in the qubits touched soonest and to evict the local qubits whose next touch
  is farthest away (Belady), preferring high local positions so the pack/unpack kernels stream.

Execution: each remap is chunked (``chunk_amps``) through two small scratch buffers so that a
128 GiB shard is exchanged in place: pack chunk -> ``isend``/``irecv`` with the partner (NCCL over
NVLink; gloo in the CPU tests) -> unpack into the slots the chunk vacated.

The planner is host-only and backend-agnostic; the local engine is pluggable so that the same
runner drives CUDA shards (``CudaShardEngine``) and the CPU stand-in used by the gloo tests.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._abi import QlsError
from .circuit import QuantumCircuit, flatten
from .lower import Block, LoweredProgram, LoweringOptions, fuse, lower_blocks, normalise

def _check(rc: int) -> None:
    from . import _abi

    _abi.check(rc)


# --------------------------------------------------------------------------------------
# planning
# --------------------------------------------------------------------------------------


@dataclass
class Segment:
    kind: str  # "local" | "swap"
    lowered: Optional[LoweredProgram] = None
    pairs: List[Tuple[int, int]] = field(default_factory=list)  # swap: (global physical position, local position)


@dataclass
class ShardedPlan:
    num_qubits: int
    n_local: int
    n_global: int
    segments: List[Segment]
    initial_layout: List[int]  # physical position of logical qubit q
    final_layout: List[int]
    amp_updates: int  # whole-state sum over fused blocks of 2^(n-c)
    n_source_gates: int
    # logical qubit -> 2x2 matrix of its LAST gate, not applied to the state (``defer_final_1q``): the qubit is on a global
    # position at the end of the plan and a consumer that only projects it (the solver path) folds the gate into its reduction
    deferred: Dict[int, np.ndarray] = field(default_factory=dict)

    @property
    def n_swaps(self) -> int:
        return sum(len(s.pairs) for s in self.segments if s.kind == "swap")


def _relabel(b: Block, phys: Sequence[int]) -> Block:
    """The same block on physical positions."""
    nb = Block(b.kind, tuple(phys[q] for q in b.targets), {phys[q]: v for q, v in b.controls.items()}, b.matrix,
               [(tuple(phys[q] for q in fq), ft) for fq, ft in b.factors], tuple(phys[q] for q in b.selectors),
               b.angles, b.amps, b.param, b.n_source)
    if b.kind == "diag":
        # diag targets are kept sorted by the lowering
        order = np.argsort(nb.targets)
        nb.targets = tuple(nb.targets[i] for i in order)
    return nb


def plan_sharded(circuit: QuantumCircuit, world_size: int, opts: Optional[LoweringOptions] = None,
                 protect_low: int = 8, fold_first_gates: bool = True, defer_final_1q: bool = False,
                 contiguous_evict: Any = "auto") -> ShardedPlan:
    """``contiguous_evict``: "auto" plans twice -- evicting the farthest-touched qubits wherever they sit, and evicting them
    from consecutive positions even if one of them is touched a little sooner ("force") -- and keeps the second plan if it
    does not exchange more amplitudes or run more passes (its exchanges are pitched copies: no gather / scatter kernels)."""
    if contiguous_evict == "auto":
        a = _plan_sharded(circuit, world_size, opts, protect_low, fold_first_gates, defer_final_1q, True)
        b = _plan_sharded(circuit, world_size, opts, protect_low, fold_first_gates, defer_final_1q, "force")

        def cost(pl: ShardedPlan):
            return (sum(1.0 - 0.5 ** len(sg.pairs) for sg in pl.segments if sg.kind == "swap"),
                    sum(sg.lowered.n_passes for sg in pl.segments if sg.kind == "local"))

        def n_contig(pl: ShardedPlan) -> int:
            return sum(1 for sg in pl.segments if sg.kind == "swap" and _is_window(sg.pairs))

        return b if cost(b) <= cost(a) and n_contig(b) > n_contig(a) else a
    return _plan_sharded(circuit, world_size, opts, protect_low, fold_first_gates, defer_final_1q, contiguous_evict)


def _is_window(pairs: Sequence[Tuple[int, int]]) -> bool:
    """The local positions of an exchange are consecutive."""
    ls = sorted(l for _, l in pairs)
    return all(b - a == 1 for a, b in zip(ls, ls[1:]))


def _plan_sharded(circuit: QuantumCircuit, world_size: int, opts: Optional[LoweringOptions], protect_low: int,
                  fold_first_gates: bool, defer_final_1q: bool, contiguous_evict: Any) -> ShardedPlan:
    """``defer_final_1q``: an uncontrolled one-qubit gate that is the LAST block referring to its qubit at all is not applied
    if the qubit sits on a global position when the gate is reached; it is returned in ``ShardedPlan.deferred`` instead.
    For a consumer that only needs the state projected on one value of that qubit (the HHL solver path reads the
    clock = 0 slice; the final gates of inverse QPE are the Hadamards of the clock qubits) the gate becomes a weighted sum
    over ranks, <0|G = G[0][0]<0| + G[0][1]<1| -- an all-reduce of the 2^nb-amplitude slice instead of an exchange of up to
    7/8 of every shard.  ``Statevector``-style consumers keep the full plan (default)."""
    g = int(round(math.log2(world_size)))
    if 2**g != world_size:
        raise ValueError("world size must be a power of two")
    n = circuit.num_qubits
    n_local = n - g
    base = opts or LoweringOptions()
    if n_local < 2:
        raise ValueError("too few local qubits")
    leaves, phase = flatten(circuit)
    blocks = fuse(normalise(leaves), base.max_fuse_qubits, base.max_diag_qubits)
    if phase:
        ph = np.array([np.exp(1j * phase)])
        if blocks and blocks[0].kind == "init":
            b0 = blocks[0]
            blocks[0] = Block("init", b0.targets, amps=b0.amps * ph[0])
        else:
            blocks.insert(0, Block("diag", (), factors=[((), ph)], n_source=0))
    # candidates for deferral: uncontrolled 1-qubit gates after which nothing refers to their qubit any more
    last_ref = [-1] * n
    for i, b in enumerate(blocks):
        for q in b.support:
            last_ref[q] = i
    defer_ids = {id(b) for i, b in enumerate(blocks)
                 if defer_final_1q and b.kind == "dense" and len(b.targets) == 1 and not b.controls and last_ref[b.targets[0]] == i}
    # next-touch table (Belady): for every block index, the next index at which each logical qubit is touched
    # (a deferrable gate does not count: it never needs its qubit to be local)
    inf = len(blocks) + 1
    touch_at: List[List[int]] = [[] for _ in range(n)]
    for i, b in enumerate(blocks):
        if id(b) in defer_ids:
            continue
        for q in b.touched:
            touch_at[q].append(i)

    def next_touch(q: int, i: int) -> int:
        for t in touch_at[q]:
            if t >= i:
                return t
        return inf

    phys = list(range(n))  # logical -> physical
    # initial layout.  An initial state preparation keeps its register on the low local positions.  A qubit whose
    # FIRST use is an uncontrolled 1-qubit gate G can start on a global position for free: G|0> = G00|0> + G10|1> is
    # folded into the state preparation (rank bit b scales the prepared amplitudes by G[b][0]; no communication), which
    # keeps qubits that stay |0> for long (the HHL flag) LOCAL, where every rank profits from skipping their |1> half.
    # Remaining global positions: the qubits touched last.
    pinned = set(blocks[0].targets) if blocks and blocks[0].kind == "init" else set()
    # a GEMM block (dense controlled e^{iAt} on the data register) always executes on positions 0..nb-1
    # (csrc/ctrl_gemm.cu): its target qubits never leave their low positions, whatever ``protect_low`` says
    gemm_pinned: set = set()
    for b in blocks:
        if b.kind == "gemm":
            gemm_pinned |= set(b.targets)
    pinned |= gemm_pinned
    foldable: Dict[int, int] = {}
    if blocks and blocks[0].kind == "init" and fold_first_gates:
        seen: set = set()
        for i, b in enumerate(blocks[1:], start=1):
            if b.kind == "dense" and len(b.targets) == 1 and not b.controls and b.targets[0] not in seen \
                    and b.targets[0] not in pinned:
                foldable[b.targets[0]] = i
            seen |= set(b.support)
    global_logical = sorted(sorted(foldable, reverse=True)[:g])
    if len(global_logical) < g:
        order = sorted((q for q in range(n) if q not in pinned and q not in global_logical), key=lambda q: (-next_touch(q, 0), -q))
        global_logical = sorted(global_logical + order[: g - len(global_logical)])
    folded = {q: foldable[q] for q in global_logical if q in foldable}
    rest = [q for q in range(n) if q not in global_logical]
    for pos, q in enumerate(rest):
        phys[q] = pos
    for j, q in enumerate(global_logical):
        phys[q] = n_local + j
    rank_factors: Optional[List[complex]] = None
    if folded:
        rank_factors = []
        for r in range(world_size):
            f = 1.0 + 0.0j
            for q, i in folded.items():
                f *= complex(blocks[i].matrix[(r >> (phys[q] - n_local)) & 1, 0])
            rank_factors.append(f)
        drop = set(folded.values())
        blocks = [b for i, b in enumerate(blocks) if i not in drop]
        touch_at = [[] for _ in range(n)]
        for i, b in enumerate(blocks):
            if id(b) in defer_ids:
                continue
            for q in b.touched:
                touch_at[q].append(i)
    initial = list(phys)
    deferred: Dict[int, np.ndarray] = {}

    segments: List[Segment] = []
    cur: List[Block] = []
    amp_updates = 0

    def close_local() -> None:
        nonlocal cur
        if cur:
            awake = tuple(initial[q] for q in folded) if not segments else ()
            o = LoweringOptions(**{**base.__dict__, "n_local": n_local, "awake_after_init": awake,
                                   "zero_start": False})  # a leading state preparation implies the tracking; never after a remap
            low = lower_blocks(n, cur, o)
            if not segments and rank_factors is not None and low.steps and low.steps[0].kind == "init":
                low.steps[0].rank_factors = rank_factors
            segments.append(Segment("local", lowered=low))
            cur = []

    for i, b in enumerate(blocks):
        if id(b) in defer_ids and phys[b.targets[0]] >= n_local:
            deferred[b.targets[0]] = np.asarray(b.matrix, dtype=complex).reshape(2, 2)
            continue
        if b.kind != "init":
            amp_updates += 2 ** (n - len(b.controls))
        need_global = [q for q in b.touched if phys[q] >= n_local]
        if need_global:
            close_local()
            where = {p: q for q, p in enumerate(phys)}
            # bring in: every global qubit, soonest touch first (at least the ones this block needs)
            incoming = sorted((q for q in range(n) if phys[q] >= n_local), key=lambda q: next_touch(q, i))
            incoming = [q for q in incoming if q in need_global or next_touch(q, i) < inf]
            # evict: local qubits with the farthest next touch, high positions first, never the low run
            cands = [q for q in range(n) if protect_low <= phys[q] < n_local and q not in b.touched and q not in gemm_pinned]
            cands.sort(key=lambda q: (-next_touch(q, i), -phys[q]))
            chosen: List[Tuple[int, int]] = []  # (incoming qubit, evicted qubit)
            for qin in incoming:
                if len(chosen) == len(cands):
                    break
                qout = cands[len(chosen)]
                if qin not in need_global and next_touch(qout, i) <= next_touch(qin, i):
                    break  # not worth it
                chosen.append((qin, qout))
            if contiguous_evict and len(chosen) > 1:
                # Same number of qubits, but evicted from CONSECUTIVE local positions [p, p + k): a sub-block of the
                # exchange is then rows of 2^p amplitudes at one pitch -- one pitched copy per chunk for the copy engines,
                # no gather / scatter kernels.  Window = the one whose soonest-touched member is touched latest (never
                # sooner than the soonest-touched qubit of the greedy choice, or the greedy choice stands).
                k = len(chosen)
                ok = {phys[q]: q for q in cands}
                floor = min(next_touch(qo, i) for _, qo in chosen)
                best = None
                for p0 in range(protect_low, n_local - k + 1):
                    if all(p0 + j in ok for j in range(k)):
                        key = (min(next_touch(ok[p0 + j], i) for j in range(k)), p0)
                        if best is None or key > best[0]:
                            best = (key, p0)
                if best is not None and (best[0][0] >= floor or contiguous_evict == "force"):
                    window = [ok[best[1] + j] for j in range(k)]
                    chosen = [(qin, qout) for (qin, _), qout in zip(chosen, window)]
            pairs = []
            for qin, qout in chosen:
                pairs.append((phys[qin], phys[qout]))
                phys[qin], phys[qout] = phys[qout], phys[qin]
            if any(phys[q] >= n_local for q in b.touched):
                raise ValueError("cannot make the block local (too few evictable qubits)")
            segments.append(Segment("swap", pairs=pairs))
            del where
        cur.append(_relabel(b, phys))
    close_local()
    return ShardedPlan(n, n_local, g, segments, initial, list(phys), amp_updates, len(leaves), deferred)


def window_chunk_geometry(p: int, k: int, value: int, begin: int, count: int, esz: int) -> Tuple[int, int, int, int]:
    """(byte offset in the shard, pitch, width, height -- all in bytes / rows) of elements ``[begin, begin + count)`` of the
    sub-block whose local bits ``[p, p + k)`` equal ``value``: element ``e`` of the sub-block sits at amplitude index
    ``(e >> p) << (p + k) | value << p | e & (2^p - 1)``, i.e. rows of ``2^p`` amplitudes at a pitch of ``2^(p+k)``.
    ``begin`` and ``count`` are powers of two times each other (chunks of a power-of-two block)."""
    row = 1 << p
    if count <= row:  # inside one row
        off = (((begin >> p) << (p + k)) | (value << p) | (begin & (row - 1))) * esz
        return off, count * esz, count * esz, 1
    off = (((begin >> p) << (p + k)) | (value << p)) * esz
    return off, (row << k) * esz, row * esz, count >> p


def solution_slice_plan(plan: ShardedPlan, nb: int, rank: int) -> Tuple[np.ndarray, np.ndarray, complex]:
    """Which entries of the solution slice (last qubit = 1, qubits nb..n-2 = 0, qubits 0..nb-1 = i) ``rank`` contributes to,
    their addresses in its shard, and the rank's weight (product of G[0][own bit] over the deferred gates G)."""
    fl, nl, n = plan.final_layout, plan.n_local, plan.num_qubits
    i = np.arange(2**nb, dtype=np.int64)
    idx = np.zeros(2**nb, dtype=np.int64)
    for q in range(nb):
        idx |= ((i >> q) & 1) << fl[q]
    idx |= 1 << fl[n - 1]
    weight = 1.0 + 0.0j
    dmask = 0
    for q, m in plan.deferred.items():
        g = fl[q] - nl
        dmask |= 1 << g
        weight *= complex(m[0, (rank >> g) & 1])
    owner = idx >> nl
    sel = np.nonzero((owner & ~dmask) == (rank & ~dmask))[0]
    return sel, idx[sel] & (2**nl - 1), weight


# --------------------------------------------------------------------------------------
# execution
# --------------------------------------------------------------------------------------


class ShardedRunner:
    """Drives one rank's shard through a plan.  ``engine`` supplies the local work:

    ``run_local(lowered, initialise)``; ``pack(local_positions, value_bits, begin, count) -> tensor``
    (a view of engine-owned scratch: the sub-block whose local bits ``local_positions[m]`` equal bit
    ``m`` of ``value_bits``); ``unpack(tensor, local_positions, value_bits, begin, count)``;
    ``recv_buffer(count)``."""

    def __init__(self, plan: ShardedPlan, engine: Any, rank: int, world: int, chunk_amps: int = 1 << 26, dist: Any = None):
        self.plan, self.engine, self.rank, self.world = plan, engine, rank, world
        self.chunk = int(chunk_amps)
        if dist is None:
            import torch.distributed as dist  # noqa: PLC0415
        self.dist = dist
        self.bytes_sent = 0

    def run(self) -> None:
        first = True
        for seg in self.plan.segments:
            if seg.kind == "local":
                self.engine.run_local(seg.lowered, initialise=first)
                first = False
            else:
                if first:
                    raise ValueError("a plan cannot start with a remap")
                self.swap_many([(gpos - self.plan.n_local, lpos) for gpos, lpos in seg.pairs])

    def swap_many(self, pairs: Sequence[Tuple[int, int]]) -> None:
        """Exchange the global positions ``n_local + gbit`` with the local positions ``lpos`` of all
        ``k`` pairs at once: the amplitude at (global bits r, local bits v) moves to (global bits v,
        local bits r).  Every rank keeps the sub-block whose local bits equal its own global bits and
        swaps the sub-block with local bits ``p`` with the rank whose global bits are ``p``, in place:
        2^k - 1 pairwise exchanges of 2^(n_local-k) amplitudes = (1 - 2^-k) of a shard on the wire,
        instead of k/2 with one half-shard swap per qubit.  The rounds use the same XOR offset on all
        ranks, so they form perfect matchings (no ordering deadlock).

        Software pipeline over chunks with two scratch slots: while chunk ``i`` is on the wire,
        chunk ``i+1`` is packed and chunk ``i-1`` is unpacked (NCCL runs on its own stream)."""
        dist = self.dist
        order = sorted(range(len(pairs)), key=lambda j: pairs[j][1])  # local positions ascending
        gbits = [pairs[j][0] for j in order]
        lposs = [pairs[j][1] for j in order]
        k = len(pairs)
        block = 2 ** (self.plan.n_local - k)
        if getattr(self.engine, "peers", None) and os.environ.get("QLS_REMAP_PUSH", "1") != "0":
            if self.ce_capable(lposs):
                torch = self.engine.torch
                main = torch.cuda.current_stream()
                for ev in self.exchange_ce(pairs, [main.record_event()]):
                    main.wait_event(ev)
            else:
                self._swap_many_push(gbits, lposs, k, block)
            return
        if getattr(self.engine, "peers", None):
            # peer-memory path: one swap kernel per round and GPU on disjoint halves of the element range, the
            # partner's state mapped through CUDA IPC; a device-side barrier (tiny all-reduce) before and after
            self.engine.device_barrier(dist)
            my_value = sum(((self.rank >> gbits[j]) & 1) << j for j in range(k))
            for d in range(1, 2**k):
                partner = self.rank
                for j in range(k):
                    if (d >> j) & 1:
                        partner ^= 1 << gbits[j]
                value = sum(((partner >> gbits[j]) & 1) << j for j in range(k))
                half = block // 2
                begin, count = (0, half) if self.rank < partner else (half, block - half)
                self.engine.swap_peer(partner, lposs, value, my_value, begin, count)
                self.bytes_sent += block * self.engine.state.tensor.element_size()
            self.engine.device_barrier(dist)
            return
        # one flat job list over (round, chunk): the pack / wire / unpack pipeline is not drained between rounds
        jobs = []
        for d in range(1, 2**k):
            partner = self.rank
            for j in range(k):
                if (d >> j) & 1:
                    partner ^= 1 << gbits[j]
            value = sum(((partner >> gbits[j]) & 1) << j for j in range(k))  # partner's global bits, as local-bit values
            for b in range(0, block, self.chunk):
                jobs.append((partner, value, b, min(self.chunk, block - b)))

        def start(i: int):
            partner, value, begin, count = jobs[i]
            out = self.engine.pack(lposs, value, begin, count, slot=i & 1)
            inc = self.engine.recv_buffer(count, slot=i & 1)
            ops = [dist.P2POp(dist.isend, out, partner), dist.P2POp(dist.irecv, inc, partner)]
            if self.rank > partner:
                ops.reverse()
            self.bytes_sent += out.numel() * out.element_size()
            return dist.batch_isend_irecv(ops), inc

        inflight = start(0)
        for i, (partner, value, begin, count) in enumerate(jobs):
            nxt = start(i + 1) if i + 1 < len(jobs) else None
            reqs, inc = inflight
            for req in reqs:
                req.wait()
            self.engine.unpack(inc, lposs, value, begin, count)
            inflight = nxt


    def ce_capable(self, lposs: Sequence[int]) -> bool:
        """Exchanges whose local positions are consecutive can be done by the copy engines alone (pitched copies)."""
        ls = sorted(lposs)
        return (os.environ.get("QLS_REMAP_CE", "1") != "0" and hasattr(self.engine, "copy_rows")
                and all(b - a == 1 for a, b in zip(ls, ls[1:])))

    def exchange_ce(self, pairs: Sequence[Tuple[int, int]], ready: Sequence[Any], part_bits: Sequence[int] = ()) -> List[Any]:
        """Copy-engine exchange of ``k`` CONSECUTIVE local positions ``[p, p + k)``: the sub-block with local bits ``v`` is rows
        of ``2^p`` amplitudes at a pitch of ``2^(p+k)``, so a chunk of it is ONE pitched copy.  Per (round, chunk) job:
        my sub-block chunk -> the partner's receive slot (peer copy, stream c1) -> barrier -> receive slot -> the rows the
        chunk vacated (local pitched copy, stream c2).  No kernel touches the amplitudes, so local passes can run beside it.

        ``ready``: one event per part; the jobs of part ``i`` start after ``ready[i]``.  ``part_bits`` (one or two bits of the
        sub-block's element index, each >= log2(chunk)) cut the exchange into 2 or 4 parts: part ``i`` holds the elements whose
        bit ``part_bits[j]`` equals bit ``j`` of ``i`` -- the caller runs passes on the parts of the shard that are not on the
        wire.  Returns one event per part: the part has been scattered.  All ranks enqueue the same job order; the main
        stream is not touched."""
        torch, dist, eng = self.engine.torch, self.dist, self.engine
        order = sorted(range(len(pairs)), key=lambda j: pairs[j][1])
        gbits = [pairs[j][0] for j in order]
        lposs = [pairs[j][1] for j in order]
        k, p = len(pairs), lposs[0]
        block = 2 ** (self.plan.n_local - k)
        esz = eng.state.tensor.element_size()
        chunk = min(self.chunk, eng._recv[0].numel(), block)
        n_parts = 1 << len(part_bits)
        assert len(ready) == n_parts and all((1 << pb) >= chunk for pb in part_bits)

        def part_of(b: int) -> int:
            return sum(((b >> pb) & 1) << j for j, pb in enumerate(part_bits))
        jobs = []
        for part in range(n_parts):
            for d in range(1, 2**k):
                partner = self.rank
                for j in range(k):
                    if (d >> j) & 1:
                        partner ^= 1 << gbits[j]
                value = sum(((partner >> gbits[j]) & 1) << j for j in range(k))
                for b in range(0, block, chunk):
                    if part_of(b) == part:
                        jobs.append((part, partner, value, b, min(chunk, block - b)))

        def geometry(value: int, begin: int, count: int) -> Tuple[int, int, int, int]:
            return window_chunk_geometry(p, k, value, begin, count, esz)

        c1, c2 = eng.comm_stream, eng.scatter_stream
        pair_sync = os.environ.get("QLS_REMAP_PAIR_SYNC", "1") != "0" and self.world > 2
        base = eng.state.ptr
        scattered: List[Any] = []
        done: List[Any] = [None] * n_parts
        started = [False] * n_parts
        for i, (part, partner, value, begin, count) in enumerate(jobs):
            slot = i & 1
            off, pitch, width, height = geometry(value, begin, count)
            with torch.cuda.stream(c1):
                if not started[part]:
                    c1.wait_event(ready[part])  # the passes before the exchange have produced this part
                    started[part] = True
                eng.copy_rows(eng.peer_stage[partner][slot], width, base + off, pitch, width, height)
                if scattered:
                    c1.wait_event(scattered[-1])  # my scatter of job i - 1 precedes sync i (the receive slot of job i + 1 is that one)
                # sync i tells me that job i has arrived and the rank I send job i + 1 to that its slot is free: with the same
                # partner for both a pairwise hand-shake will do (a barrier over all ranks makes every job wait for the
                # slowest pair: 511 GB/s at 8 GPUs against 667 at 2); the last job of a round needs everybody
                if pair_sync and i + 1 < len(jobs) and jobs[i + 1][1] == partner:
                    eng.pair_barrier(dist, partner)
                else:
                    eng.device_barrier(dist)
                arrived = c1.record_event()
            with torch.cuda.stream(c2):
                c2.wait_event(arrived)
                eng.copy_rows(base + off, pitch, eng._recv[slot].data_ptr(), width, width, height)
                scattered.append(c2.record_event())
            done[part] = scattered[-1]  # c2 is in order: the last scatter of a part implies the earlier ones
            self.bytes_sent += count * esz
        with torch.cuda.stream(c1):  # nobody starts the next exchange's copies before every scatter of this one is done
            for ev in scattered[-2:]:
                c1.wait_event(ev)
            eng.device_barrier(dist)
            closed = c1.record_event()
        return [ev if ev is not None else closed for ev in done[:-1]] + [closed]

    def _swap_many_push(self, gbits: Sequence[int], lposs: Sequence[int], k: int, block: int) -> None:
        """Push-only exchange over peer memory.  Job (round, chunk): every rank gathers the chunk of the sub-block that
        belongs to its partner straight into the partner's receive slot -- posted remote WRITES only, no load ever crosses
        NVLink (the in-place swap kernel pays a remote read round trip per element: 474-590 GB/s per direction against
        the 770 GB/s of a peer copy) --, a device-side barrier, then every rank scatters what it received into the places
        the chunk vacated (local traffic).  Two receive slots and two streams: the push of job i + 1 runs on the
        communication stream while job i is scattered on the main stream.  Safety of slot reuse: the push of job i + 2
        follows barrier i + 1, which every rank enters only after its scatter of job i."""
        torch = self.engine.torch
        dist, eng = self.dist, self.engine
        jobs = []
        for d in range(1, 2**k):
            partner = self.rank
            for j in range(k):
                if (d >> j) & 1:
                    partner ^= 1 << gbits[j]
            value = sum(((partner >> gbits[j]) & 1) << j for j in range(k))  # partner's global bits, as local-bit values
            chunk = min(self.chunk, eng._recv[0].numel())
            for b in range(0, block, chunk):
                jobs.append((partner, value, b, min(chunk, block - b)))
        main, comm = torch.cuda.current_stream(), eng.comm_stream
        comm.wait_stream(main)  # the local passes before the exchange have produced what is pushed
        done: List[Any] = []
        if os.environ.get("QLS_REMAP_PUSH", "1") == "sm":
            # remote stores issued by the gather kernel itself (measured slower than the copy engines: 531 vs 586 GB/s at 2 GPUs)
            for i, (partner, value, begin, count) in enumerate(jobs):
                slot = i & 1
                with torch.cuda.stream(comm):
                    eng.push(partner, lposs, value, begin, count, slot)
                    if done:
                        comm.wait_event(done[-1])  # my scatter of job i - 1 precedes barrier i
                    eng.device_barrier(dist)
                    arrived = comm.record_event()
                main.wait_event(arrived)
                eng.unpack(eng._recv[slot][:count], lposs, value, begin, count)
                done.append(main.record_event())
                self.bytes_sent += count * eng.state.tensor.element_size()
        else:
            # gather into my send slot (main stream) -> copy engines to the partner's receive slot + barrier (communication
            # stream) -> scatter (main stream); job i + 1 is gathered while job i is on the wire
            copied: List[Any] = []
            eng.pack(lposs, jobs[0][1], jobs[0][2], jobs[0][3], slot=0)
            packed = main.record_event()
            for i, (partner, value, begin, count) in enumerate(jobs):
                slot = i & 1
                with torch.cuda.stream(comm):
                    comm.wait_event(packed)
                    eng.send_slot(partner, count, slot)
                    copied.append(comm.record_event())
                    if done:
                        comm.wait_event(done[-1])  # my scatter of job i - 1 precedes barrier i
                    eng.device_barrier(dist)
                    arrived = comm.record_event()
                if i + 1 < len(jobs):
                    if i >= 1:
                        main.wait_event(copied[i - 1])  # the send slot of job i + 1 was last read by the copy of job i - 1
                    eng.pack(lposs, jobs[i + 1][1], jobs[i + 1][2], jobs[i + 1][3], slot=(i + 1) & 1)
                    packed = main.record_event()
                main.wait_event(arrived)
                eng.unpack(eng._recv[slot][:count], lposs, value, begin, count)
                done.append(main.record_event())
                self.bytes_sent += count * eng.state.tensor.element_size()
        with torch.cuda.stream(comm):  # nobody starts the next exchange's pushes before every scatter of this one is done
            for ev in done[-2:]:
                comm.wait_event(ev)
            eng.device_barrier(dist)
            closed = comm.record_event()
        main.wait_event(closed)


class CudaShardEngine:
    """One CUDA shard: DeviceState + per-segment DevicePrograms + two scratch chunks."""

    def __init__(self, plan: ShardedPlan, rank: int, precision: str = "fp64", device: Optional[int] = None,
                 chunk_amps: int = 1 << 26):
        import torch

        from . import _abi
        from .engine import DeviceProgram, DeviceState

        self._abi = _abi
        self.torch = torch
        self.plan = plan
        self.rank = rank
        self.state = DeviceState(plan.n_local, precision, device, index_offset=rank << plan.n_local)
        self.half = 2 ** (plan.n_local - 1)
        self.programs: Dict[int, DeviceProgram] = {}
        for seg in plan.segments:
            if seg.kind == "local":
                self.programs[id(seg.lowered)] = DeviceProgram(seg.lowered, self.state.device)
        chunk = min(int(chunk_amps), self.half)
        with torch.cuda.device(self.state.device):
            self._send = [torch.empty(chunk, dtype=self.state.tensor.dtype, device=self.state.tensor.device) for _ in range(2)]
            self._recv = [torch.empty(chunk, dtype=self.state.tensor.dtype, device=self.state.tensor.device) for _ in range(2)]

    def _stream(self) -> int:
        return int(self.torch.cuda.current_stream().cuda_stream)

    def _ipc_export(self, t) -> Tuple[bytes, int]:
        """(64-byte CUDA IPC handle of the allocation ``t`` lives in, byte offset of ``t`` inside it)."""
        _dev, handle, _size, storage_off, *_ = t.untyped_storage()._share_cuda_()  # cudaIpcGetMemHandle of the allocation
        offset = int(storage_off) + t.storage_offset() * t.element_size()
        handle = bytes(handle)
        # torch prefixes the 64-byte cudaIpcMemHandle_t with [version byte,] allocation kind ('c' = plain cudaMalloc)
        if len(handle) in (65, 66):
            if handle[len(handle) - 65:len(handle) - 64] != b"c":
                raise QlsError("buffer is not a plain cudaMalloc allocation (expandable segments?): no IPC handle")
            handle = handle[-64:]
        if len(handle) != 64:
            raise QlsError(f"unexpected CUDA IPC handle size {len(handle)}")
        return handle, offset

    def map_peers(self, dist: Any) -> None:
        """Map every other rank's state buffer and its two receive slots into this process (CUDA IPC handles exchanged
        through torch.distributed; peer access over NVLink is enabled when a handle is opened)."""
        import ctypes as C

        t = self.state.tensor
        world = dist.get_world_size()
        gathered: List[Any] = [None] * world
        dist.all_gather_object(gathered, [self._ipc_export(x) for x in (t, self._recv[0], self._recv[1])])
        self._ipc_bases = []
        opened: Dict[Tuple[int, bytes], int] = {}  # an allocation can be opened once per process

        def open_(r: int, h: bytes, off: int) -> int:
            if (r, h) not in opened:
                base = C.c_void_p()
                buf = C.create_string_buffer(h, 64)
                self._abi.check(self.state.lib.qls_ipc_open(C.cast(buf, C.c_void_p), t.device.index, C.byref(base)))
                self._ipc_bases.append(base)
                opened[(r, h)] = base.value
            return opened[(r, h)] + off  # device pointer into rank r's memory, valid in kernels of this device

        peers: Dict[int, int] = {}
        stage: Dict[int, Tuple[int, int]] = {}
        for r, entries in enumerate(gathered):
            if r != self.rank:
                peers[r] = open_(r, *entries[0])
                stage[r] = (open_(r, *entries[1]), open_(r, *entries[2]))
        self.peers, self.peer_stage = peers, stage
        with self.torch.cuda.device(self.state.device):
            self._tiny = self.torch.zeros(1, dtype=self.torch.float32, device=t.device)
            self.comm_stream = self.torch.cuda.Stream(device=t.device)
            self.scatter_stream = self.torch.cuda.Stream(device=t.device)

    def copy_rows(self, dst: int, dst_pitch: int, src: int, src_pitch: int, width: int, height: int) -> None:
        """Pitched copy between device addresses (own or peer-mapped) by the copy engines, on the current stream."""
        st = self.state
        with self.torch.cuda.device(st.device):
            self._abi.check(st.lib.qls_copy2d_async(dst, dst_pitch, src, src_pitch, width, height, self._stream()))

    def send_slot(self, partner: int, count: int, slot: int) -> None:
        """Copy ``count`` amplitudes of my send slot into receive slot ``slot`` of ``partner`` (copy engines; ordered after the
        work of the current stream, which waits for the copy).  QLS_REMAP_COPY_WAYS > 1 splits the copy over that many
        streams so that several copy engines work on it."""
        st, torch = self.state, self.torch
        esz = self._send[slot].element_size()
        ways = max(1, int(os.environ.get("QLS_REMAP_COPY_WAYS", "1")))
        with torch.cuda.device(st.device):
            if ways == 1 or count < (1 << 20):
                self._abi.check(st.lib.qls_copy_async(self.peer_stage[partner][slot], self._send[slot].data_ptr(), count * esz,
                                                      self._stream()))
                return
            if len(getattr(self, "_copy_streams", [])) < ways:
                self._copy_streams = [torch.cuda.Stream(device=st.tensor.device) for _ in range(ways)]
            cur = torch.cuda.current_stream()
            ready = cur.record_event()
            part = (count + ways - 1) // ways
            for w in range(ways):
                lo, hi = w * part, min(count, (w + 1) * part)
                if lo >= hi:
                    break
                sw = self._copy_streams[w]
                sw.wait_event(ready)
                self._abi.check(st.lib.qls_copy_async(self.peer_stage[partner][slot] + lo * esz, self._send[slot].data_ptr() + lo * esz,
                                                      (hi - lo) * esz, int(sw.cuda_stream)))
                cur.wait_event(sw.record_event())

    def push(self, partner: int, lposs: Sequence[int], value: int, begin: int, count: int, slot: int) -> None:
        """Gather ``count`` amplitudes of my sub-block ``value`` straight into receive slot ``slot`` of ``partner`` (remote
        writes over NVLink, local reads; runs on the current stream)."""
        import ctypes as C

        st = self.state
        arr = (C.c_int * len(lposs))(*lposs)
        with self.torch.cuda.device(st.device):
            self._abi.check(st.lib.qls_pack_bits(st.ptr, self.peer_stage[partner][slot], st.n_local, st.abi_dtype, arr, len(lposs),
                                                 value, begin, count, self._stream()))

    def device_barrier(self, dist: Any) -> None:
        dist.all_reduce(self._tiny)  # stream ordered on every rank: nobody passes before everybody's earlier kernels are done

    def pair_barrier(self, dist: Any, partner: int) -> None:
        """Stream-ordered hand-shake with ONE rank (4-byte send + receive): neither side passes before the other's earlier work
        on its stream is done."""
        if not hasattr(self, "_tiny_rx"):
            self._tiny_rx = self.torch.zeros_like(self._tiny)
        ops = [dist.P2POp(dist.isend, self._tiny, partner), dist.P2POp(dist.irecv, self._tiny_rx, partner)]
        if self.rank > partner:
            ops.reverse()
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def swap_peer(self, partner: int, lposs: Sequence[int], my_value: int, peer_value: int, begin: int, count: int) -> None:
        import ctypes as C

        st = self.state
        arr = (C.c_int * len(lposs))(*lposs)
        with self.torch.cuda.device(st.device):
            self._abi.check(st.lib.qls_swap_peer(st.ptr, self.peers[partner], st.n_local, st.abi_dtype, arr, len(lposs),
                                                 my_value, peer_value, begin, count, self._stream()))

    def run_local(self, lowered: LoweredProgram, initialise: bool) -> None:
        self.programs[id(lowered)].run(self.state, initialise=initialise)

    def pack(self, lposs: Sequence[int], value: int, begin: int, count: int, slot: int = 0):
        import ctypes as C

        st = self.state
        arr = (C.c_int * len(lposs))(*lposs)
        with self.torch.cuda.device(st.device):
            self._abi.check(st.lib.qls_pack_bits(st.ptr, self._send[slot].data_ptr(), st.n_local, st.abi_dtype, arr, len(lposs),
                                                 value, begin, count, self._stream()))
        return self._send[slot][:count]

    def recv_buffer(self, count: int, slot: int = 0):
        return self._recv[slot][:count]

    def unpack(self, buf, lposs: Sequence[int], value: int, begin: int, count: int) -> None:
        import ctypes as C

        st = self.state
        arr = (C.c_int * len(lposs))(*lposs)
        with self.torch.cuda.device(st.device):
            self._abi.check(st.lib.qls_unpack_bits(st.ptr, buf.data_ptr(), st.n_local, st.abi_dtype, arr, len(lposs), value, begin,
                                                   count, self._stream()))


class ShardedShapedSolver:
    """Multi-GPU counterpart of ``ShapedSolver`` (one process per GPU, ``torch.distributed`` is
    initialised by the caller): ``solve(b_host)`` / ``run_device(b_dev)`` + ``reduce()``."""

    def __init__(self, circuit: QuantumCircuit, nb: int, precision: str = "fp64", device: Optional[int] = None,
                 options: Optional[LoweringOptions] = None, chunk_amps: int = 1 << 27):
        import torch.distributed as dist

        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        opts = options or LoweringOptions(dtype=precision)
        opts.dtype = precision
        self.n, self.nb = circuit.num_qubits, int(nb)
        # the solver path only reads the clock = 0 slice: the final Hadamards of clock qubits that are on global positions at
        # the end become rank weights of the slice's all-reduce instead of one more exchange (QLS_SHARD_DEFER=0: full plan)
        self.plan = plan_sharded(circuit, self.world, opts, defer_final_1q=os.environ.get("QLS_SHARD_DEFER", "1") != "0")
        chunk_amps = 1 << int(os.environ.get("QLS_REMAP_CHUNK_LOG2", int(chunk_amps).bit_length() - 1))
        self.engine = CudaShardEngine(self.plan, self.rank, precision, device, chunk_amps)
        self.runner = ShardedRunner(self.plan, self.engine, self.rank, self.world, chunk_amps, dist)
        self.transport_fallback: Optional[str] = None  # why the peer-memory remaps are not in use (None: they are, or were not asked for)
        if os.environ.get("QLS_REMAP_P2P", "1") == "1":
            # remaps through peer memory (CUDA IPC + NVLink loads/stores: 590 GB/s per direction measured) instead of
            # pack + NCCL send/recv + unpack (446 GB/s); all ranks must agree, so a failure anywhere falls back everywhere
            why = ""
            try:
                self.engine.map_peers(dist)
            except Exception as exc:  # noqa: BLE001 -- no IPC / no peer access on this box
                self.engine.peers = None
                why = f"rank {self.rank}: {type(exc).__name__}: {exc}"
            reasons: List[Any] = [None] * self.world
            dist.all_gather_object(reasons, why)
            if any(reasons):
                # never silent: the fall-back costs bandwidth (pack + NCCL send/recv + unpack, 446 vs 590 GB/s)
                self.engine.peers = None
                self.transport_fallback = "; ".join(r for r in reasons if r)
                if self.rank == 0:
                    import sys

                    sys.stderr.write(f"qls_b200.shard: peer-memory remaps unavailable, using the NCCL path ({self.transport_fallback})\n")
        self.state = self.engine.state
        first = self.plan.segments[0].lowered
        if not (first.steps and first.steps[0].kind == "init"):
            raise ValueError("circuit must start with a state preparation on the data register")
        self._init_amps = first.steps[0].amps
        self.lowered = first  # for accounting helpers (amp_bytes, n_local)
        self.program = self.engine.programs[id(first)]
        self.program.launches_per_run = sum(self.engine.programs[id(s.lowered)].launches_per_run
                                            for s in self.plan.segments if s.kind == "local")

    def amp_updates_global(self) -> int:
        return self.plan.amp_updates

    def total_pass_bytes_per_gpu(self) -> int:
        return sum(s.lowered.total_pass_bytes() for s in self.plan.segments if s.kind == "local")

    def _prepare(self, b_device=None, b_host=None) -> None:
        self.state.init_basis(None)
        rf = self.plan.segments[0].lowered.steps[0].rank_factors
        factor = (1.0 if self.rank == 0 else 0.0) if rf is None else rf[self.rank]
        if factor != 0:
            if b_device is not None:
                self.state.tensor[: 2**self.nb].copy_(b_device)
                if factor != 1.0:
                    self.state.tensor[: 2**self.nb].mul_(factor)
            else:
                amps = np.asarray(self._init_amps if b_host is None else b_host)
                self.state.write(0, amps if factor == 1.0 else amps * factor)

    # -- exchanges overlapped with the local passes on both sides of them ----------------------------------------------
    def _plan_overlap(self) -> Dict[int, Tuple[List[int], List[int], List[int], List[int]]]:
        """For every exchange that the copy engines can do alone: one or two local positions ``h`` that neither the exchange nor
        the passes next to it touch, the trailing passes of the segment before (``pa``) and the leading passes of the segment
        after (``pb``) that leave them alone.  Those passes then run on one part of the shard at a time (``h = 0`` / ``h = 1``
        halves, or the four quarters of two positions), and each part is exchanged while others are being computed on:

            pa on part 0 | pa on part 1 | ...   ||   exchange part 0 | exchange part 1 | ...   ||   pb on part 0 | pb on part 1 ...
                         '------ overlaps ------'                    '------------------ overlaps -----------------'

        Returns {segment index of the exchange: (positions, their bits in the sub-block's element index, pa, pb)}."""
        segs = self.plan.segments
        nl = self.plan.n_local
        out: Dict[int, Tuple[List[int], List[int], List[int], List[int]]] = {}
        if os.environ.get("QLS_SHARD_OVERLAP", "1") == "0" or not getattr(self.engine, "peers", None) \
                or os.environ.get("QLS_REMAP_PUSH", "1") == "0":
            return out
        max_pos = 2 if os.environ.get("QLS_SHARD_OVERLAP", "1") != "2way" else 1
        chunk = min(self.runner.chunk, self.engine._recv[0].numel())
        used_prefix: Dict[int, int] = {}  # local segment index -> passes taken as "pb" of the exchange before it
        for j, seg in enumerate(segs):
            if seg.kind != "swap" or j == 0 or j + 1 >= len(segs):
                continue
            lposs = sorted(l for _, l in seg.pairs)
            if not self.runner.ce_capable(lposs):
                continue
            la, lb = segs[j - 1].lowered, segs[j + 1].lowered
            pa_prog, pb_prog = self.engine.programs[id(la)], self.engine.programs[id(lb)]
            # trailing passes of A / leading passes of B (only if the segment ends / starts with tile passes)
            a_rng = range(la.steps[-1].pass_begin, la.steps[-1].pass_end) if la.steps and la.steps[-1].kind == "passes" else range(0)
            b_rng = range(lb.steps[0].pass_begin, lb.steps[0].pass_end) if lb.steps and lb.steps[0].kind == "passes" else range(0)
            a_lo = max(a_rng.start, used_prefix.get(j - 1, 0)) if len(a_rng) else 0
            free_a = {pi: set(pa_prog.free_positions(pi)) for pi in a_rng if pi in pa_prog.jit_attached}
            free_b = {pi: set(pb_prog.free_positions(pi)) for pi in b_rng if pi in pb_prog.jit_attached}
            cands = []
            for h in range(nl):
                ebit = h - sum(1 for l in lposs if l < h)
                if h not in lposs and chunk <= (1 << ebit) < (1 << (nl - len(lposs))):
                    cands.append((h, ebit))

            def runs(hs):
                pa: List[int] = []
                for pi in reversed(a_rng):
                    if pi < a_lo or pi not in free_a or not all(h in free_a[pi] for h in hs):
                        break
                    pa.insert(0, pi)
                pb: List[int] = []
                for pi in b_rng:
                    if pi not in free_b or not all(h in free_b[pi] for h in hs):
                        break
                    pb.append(pi)
                return pa, pb, sum(la.pass_bytes(pi) for pi in pa) + sum(lb.pass_bytes(pi) for pi in pb)

            best = None
            for h, ebit in cands:
                pa, pb, score = runs([h])
                if score > 0 and (best is None or score > best[0]):
                    best = (score, [h], [ebit], pa, pb)
            if best is not None and max_pos > 1:
                # a second position: quarters instead of halves (the exposed head and tail of the pipeline halve) if that
                # keeps at least 3/4 of the overlapped passes
                best2 = None
                for i1, (h1, e1) in enumerate(cands):
                    for h2, e2 in cands[i1 + 1:]:
                        pa, pb, score = runs([h1, h2])
                        if score > 0 and (best2 is None or score > best2[0]):
                            best2 = (score, [h1, h2], [e1, e2], pa, pb)
                if best2 is not None and 4 * best2[0] >= 3 * best[0]:
                    best = best2
            if best is not None:
                out[j] = best[1:]
                used_prefix[j + 1] = (best[4][-1] + 1) if best[4] else 0
        return out

    def _run_segments(self) -> None:
        segs = self.plan.segments
        if getattr(self, "_overlap", None) is None:
            self._overlap = self._plan_overlap()
        ov = self._overlap
        torch = self.engine.torch
        lib = self.state.lib
        for j, seg in enumerate(segs):
            if seg.kind == "local":
                prog = self.engine.programs[id(seg.lowered)]
                first = (ov[j - 1][3][-1] + 1) if (j - 1) in ov and ov[j - 1][3] else 0          # run as "pb" of the exchange before
                last = ov[j + 1][2][0] if (j + 1) in ov and ov[j + 1][2] else None               # run as "pa" of the exchange after
                prog.run(self.state, initialise=False, first_pass=first, last_pass=last)         # state prepared by _prepare
                continue
            pairs = [(gpos - self.plan.n_local, lpos) for gpos, lpos in seg.pairs]
            if j not in ov:
                self.runner.swap_many(pairs)
                continue
            hs, ebits, pa, pb = ov[j]
            prog_a, prog_b = self.engine.programs[id(segs[j - 1].lowered)], self.engine.programs[id(segs[j + 1].lowered)]
            main = torch.cuda.current_stream()
            parts = [[(h, (part >> i) & 1) for i, h in enumerate(hs)] for part in range(1 << len(hs))]  # part i: bit j of i = value of hs[j]
            # the exchange's hand-shake kernels (NCCL, one or two CTAs) need SMs beside the persistent pass kernels
            _check(lib.qls_set_pass_grid_reserve(int(os.environ.get("QLS_SHARD_OVERLAP_SMS", "2"))))
            try:
                ready = []
                for sel in parts:
                    for pi in pa:
                        prog_a.run_pass_part(self.state, pi, sel)
                    ready.append(main.record_event())
                done = self.runner.exchange_ce(pairs, ready, part_bits=ebits)
                for sel, ev in zip(parts, done):
                    main.wait_event(ev)
                    for pi in pb:
                        prog_b.run_pass_part(self.state, pi, sel)
                main.wait_event(done[-1])
            finally:
                _check(lib.qls_set_pass_grid_reserve(0))

    def run_device(self, b_device=None) -> None:
        self._prepare(b_device=b_device)
        self._run_segments()

    def _masks(self) -> Tuple[int, int, int, int]:
        fl = self.plan.final_layout
        flag = 1 << fl[self.n - 1]
        nondata = 0
        for q in range(self.nb, self.n):
            nondata |= 1 << fl[q]
        return flag, flag, nondata, flag

    def _slice_plan(self):
        """(indices of the solution slice this rank contributes to, their local addresses, this rank's weight).

        The slice is flag = 1, clock = 0, data = i.  A rank contributes the entries whose owner agrees with it on every
        global position EXCEPT those of deferred qubits; there its own bit b selects the weight G[0][b] of the deferred
        gate G (<0|G|psi> = G[0][0] psi_0 + G[0][1] psi_1)."""
        if getattr(self, "_slice_cache", None) is None:
            import torch

            sel, local, weight = solution_slice_plan(self.plan, self.nb, self.rank)
            dev = self.state.tensor.device
            self._slice_cache = (torch.from_numpy(sel).to(dev), torch.from_numpy(local).to(dev), weight)
        return self._slice_cache

    def _slice(self):
        """The solution slice (device tensor, complete on every rank after the all-reduce)."""
        import torch

        sel, local, weight = self._slice_plan()
        x = torch.zeros(2**self.nb, dtype=self.state.tensor.dtype, device=self.state.tensor.device)
        if sel.numel():
            x[sel] = self.state.tensor[local] * weight if weight != 1.0 else self.state.tensor[local]
        self.dist.all_reduce(torch.view_as_real(x))
        return x

    def reduce(self) -> Tuple[float, float]:
        import torch

        m1, v1, m2, v2 = self._masks()
        p_flag = self.state.prob_mask(m1, v1)
        if self.plan.deferred:
            # |x|^2 of the state AFTER the deferred gates = norm of the rank-weighted slice
            x = self._slice()
            t = torch.tensor([p_flag], dtype=torch.float64, device=self.state.tensor.device)
            self.dist.all_reduce(t)
            return float(t.item()), float((x.real.double() ** 2 + x.imag.double() ** 2).sum().item())
        norm_sq = self.state.prob_mask(m2, v2)
        t = torch.tensor([p_flag, norm_sq], dtype=torch.float64, device=self.state.tensor.device)
        self.dist.all_reduce(t)
        out = t.tolist()
        return float(out[0]), float(out[1])

    def solve(self, b_host: np.ndarray):
        self._prepare(b_host=np.asarray(b_host).reshape(-1))
        self._run_segments()
        p_flag, norm_sq = self.reduce()
        x = self._slice()
        return x.cpu().numpy(), p_flag, norm_sq

    def stats(self) -> Dict[str, int]:
        local = [s.lowered for s in self.plan.segments if s.kind == "local"]
        return {
            "fused_passes": sum(l.n_passes for l in local),
            "fused_blocks": sum(sum(len(p.op_controls) for p in l.pass_info) for l in local),
            "source_gates": self.plan.n_source_gates,
            "pass_bytes_per_gpu": sum(l.total_pass_bytes() for l in local),
            "remaps": self.plan.n_swaps,
            "remap_exchanges": sum(1 for s in self.plan.segments if s.kind == "swap"),
            # a k-qubit exchange puts (1 - 2^-k) of the shard on the wire
            "remap_bytes_sent_per_gpu": sum(int((1 - 0.5 ** len(s.pairs)) * 2 * self.engine.half) * self.lowered.amp_bytes()
                                            for s in self.plan.segments if s.kind == "swap"),
        }

    def time_passes(self, repeats: int = 1):
        """Tile-pass kernel time per solve on this rank (CUDA events), launches, algorithmic bytes.
        Also records ``self.remap_ms``: event time of the remaps (pack + send/recv + unpack)."""
        import torch

        tot, launches, nbytes, remap = 0.0, 0, 0, 0.0
        for _ in range(repeats):
            self._prepare()
            launches, nbytes, ms = 0, 0, 0.0
            for seg in self.plan.segments:
                if seg.kind == "local":
                    m, l, b = self.engine.programs[id(seg.lowered)].run_passes_timed(self.state)
                    ms, launches, nbytes = ms + m, launches + l, nbytes + b
                else:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    self.runner.swap_many([(gpos - self.plan.n_local, lpos) for gpos, lpos in seg.pairs])
                    e1.record()
                    torch.cuda.synchronize()
                    remap += e0.elapsed_time(e1)
            tot += ms
        self.remap_ms = remap / repeats
        return tot / repeats, launches, nbytes
