"""Circuit lowering: leaf gate stream -> fused blocks -> tile passes -> binary program.

Replaces the per-gate loop of ``Statevector._evolve_instruction`` (reference call site
``hhl_qiskit_implementation.py:49``) by:

1. **normalise**  every leaf becomes a dense block (k <= 5 targets + symbolic controls), a
   diagonal block (phase table over its qubits), a multiplexed-RY block (``UCRYGate`` and its
   controlled forms -- the eigenvalue inversion), an initialisation, or a GEMM block (dense
   ``e^{iAt}`` power on the low data register);
2. **fuse**       diagonal runs are merged (the ``cp`` ladders of the QFT inside QPE), RY
   multiplexers on the same target add their angle tables (``ExactReciprocal`` = one block), and
   neighbouring small dense gates are multiplied into k <= ``max_fuse_qubits`` unitaries;
3. **group**      consecutive blocks whose non-low target qubits fit the tile share one *pass*:
   one HBM read + write of the state, all blocks applied in shared memory;
4. **emit**       ``QlsPass`` / ``QlsOp`` records + constant pool (``include/qls_abi.h``).

The input circuit is never mutated (QASM / depth / width come from the un-fused circuit).
"""
from __future__ import annotations

import os

import ctypes as C
import dataclasses
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _abi
from .circuit import Leaf, QuantumCircuit, flatten

# --------------------------------------------------------------------------------------
# Blocks
# --------------------------------------------------------------------------------------


@dataclass
class Block:
    kind: str  # "dense" | "diag" | "muxry" | "init" | "gemm"
    targets: Tuple[int, ...] = ()  # dense: matrix-bit order; muxry: (target,); diag: table qubits
    controls: Dict[int, int] = field(default_factory=dict)
    matrix: Optional[np.ndarray] = None  # dense / gemm
    factors: List[Tuple[Tuple[int, ...], np.ndarray]] = field(default_factory=list)  # diag
    selectors: Tuple[int, ...] = ()  # muxry: table-bit order
    angles: Optional[np.ndarray] = None  # muxry
    amps: Optional[np.ndarray] = None  # init
    param: Optional[Tuple[str, int, float, float]] = None  # param: (gate, parameter index, coeff, offset)
    n_source: int = 1  # leaf gates folded into this block

    @property
    def touched(self) -> frozenset:
        """Qubits acted on non-diagonally."""
        if self.kind in ("dense", "gemm", "init", "param"):
            return frozenset(self.targets)
        if self.kind == "muxry":
            return frozenset(self.targets[:1])
        return frozenset()

    @property
    def support(self) -> frozenset:
        s = set(self.targets) | set(self.controls) | set(self.selectors)
        return frozenset(s)


def _commutes(a: Block, b: Block) -> bool:
    """Sufficient (not necessary) condition for two blocks to commute: neither acts
    non-diagonally on a qubit the other one uses at all.  (Controls, selectors and diagonal
    tables are all diagonal in the computational basis.)"""
    if a.touched & b.support:
        return False
    if b.touched & a.support:
        return False
    return True


def _is_diagonal(m: np.ndarray) -> bool:
    return np.count_nonzero(m - np.diag(np.diagonal(m))) == 0


def _fold_controls_diag(qubits: Tuple[int, ...], table: np.ndarray, controls: Dict[int, int]):
    """A controlled diagonal is a diagonal over controls + targets."""
    if not controls:
        return qubits, table
    cq = tuple(controls)
    nq = len(qubits)
    full = np.ones(2 ** (nq + len(cq)), dtype=complex)
    idx = np.arange(full.size)
    ok = np.ones(full.size, dtype=bool)
    for j, q in enumerate(cq):
        ok &= ((idx >> (nq + j)) & 1) == controls[q]
    full[ok] = table[idx[ok] & (2**nq - 1)]
    return qubits + cq, full


def normalise(leaves: Sequence[Leaf], max_dense: int = 5, batch: bool = False) -> List[Block]:
    out: List[Block] = []
    for lf in leaves:
        ctrl = {q: (lf.ctrl_state >> j) & 1 for j, q in enumerate(lf.controls)}
        if lf.kind == "param":
            out.append(Block("param", tuple(lf.targets), ctrl, param=tuple(lf.payload)))
        elif lf.kind == "init":
            out.append(Block("init", tuple(lf.targets), amps=np.asarray(lf.payload, dtype=complex)))
        elif lf.kind == "diag":
            tab = np.asarray(lf.payload, dtype=complex)
            if ctrl and len(lf.targets) >= 3:
                # a controlled diagonal *sub-circuit* (e.g. the phase run inside a controlled
                # e^{iAt} block): keep the controls symbolic so that a pass whose blocks all share
                # them can skip the other half of the tiles
                out.append(Block("diag", tuple(lf.targets), dict(ctrl), factors=[(tuple(lf.targets), tab)]))
            else:
                qs, tab = _fold_controls_diag(tuple(lf.targets), tab, ctrl)
                out.append(Block("diag", qs, factors=[(qs, tab)]))
        elif lf.kind == "ucry":
            tgt, sel = lf.targets[0], tuple(lf.targets[1:])
            ang = np.asarray(lf.payload, dtype=float)
            if ctrl:  # fold controls into the selector: angle 0 where the control is not satisfied
                cq = tuple(ctrl)
                ns = len(sel)
                full = np.zeros(2 ** (ns + len(cq)))
                idx = np.arange(full.size)
                ok = np.ones(full.size, dtype=bool)
                for j, q in enumerate(cq):
                    ok &= ((idx >> (ns + j)) & 1) == ctrl[q]
                full[ok] = ang[idx[ok] & (2**ns - 1)]
                sel, ang = sel + cq, full
            out.append(Block("muxry", (tgt,), selectors=sel, angles=ang))
        elif lf.kind == "dense":
            m = np.asarray(lf.payload, dtype=complex)
            k = len(lf.targets)
            if _is_diagonal(m) and k + len(ctrl) <= 10:
                qs, tab = _fold_controls_diag(tuple(lf.targets), np.diagonal(m).copy(), ctrl)
                out.append(Block("diag", qs, factors=[(qs, tab)]))
            elif k <= max_dense:
                out.append(Block("dense", tuple(lf.targets), ctrl, m))
            else:
                lo = min(lf.targets)
                want = list(range(lo, lo + k))
                if sorted(lf.targets) != want or (lo != 0 and not batch):
                    raise NotImplementedError(
                        f"dense {k}-qubit block must act on contiguous qubits"
                        f"{'' if batch else ' 0..' + str(k - 1)} (GEMM block); got {lf.targets}"
                    )
                if list(lf.targets) != want:
                    m = _permute_matrix(m, [lf.targets.index(q) for q in want])
                out.append(Block("gemm", tuple(want), ctrl, m))
        else:
            raise ValueError(lf.kind)
    return out


def _permute_matrix(m: np.ndarray, new_from_old: Sequence[int]) -> np.ndarray:
    """Matrix of the same gate with matrix bit ``j`` now being old matrix bit ``new_from_old[j]``."""
    k = len(new_from_old)
    idx = np.arange(2**k)
    old = np.zeros_like(idx)
    for j, o in enumerate(new_from_old):
        old |= ((idx >> j) & 1) << o
    return m[np.ix_(old, old)]


def _expand_dense(b: Block, qubits: Tuple[int, ...]) -> np.ndarray:
    """Full matrix of a dense block (controls folded in) on the ordered qubit tuple ``qubits``."""
    k = len(qubits)
    pos = {q: j for j, q in enumerate(qubits)}
    dim = 2**k
    full = np.zeros((dim, dim), dtype=complex)
    tpos = [pos[q] for q in b.targets]
    tmask = sum(1 << p for p in tpos)
    idx = np.arange(dim)
    ok = np.ones(dim, dtype=bool)
    for q, v in b.controls.items():
        ok &= ((idx >> pos[q]) & 1) == v
    sub = np.zeros(dim, dtype=np.int64)
    for j, p in enumerate(tpos):
        sub |= ((idx >> p) & 1) << j
    rest = idx & ~tmask
    kt = len(tpos)
    for col in range(dim):
        if not ok[col]:
            full[col, col] = 1.0
            continue
        for r in range(2**kt):
            row = rest[col]
            for j, p in enumerate(tpos):
                row |= ((r >> j) & 1) << p
            full[row, col] = b.matrix[r, sub[col]]
    return full


def _fold_diag_block(b: Block) -> Block:
    """Diagonal block with its symbolic controls folded into the tables."""
    if not b.controls:
        return b
    factors = []
    for fq, ft in b.factors:
        q2, t2 = _fold_controls_diag(tuple(fq), ft, b.controls)
        factors.append((q2, t2))
    qs = tuple(sorted(set(b.targets) | set(b.controls)))
    return Block("diag", qs, factors=factors, n_source=b.n_source)


def fuse(blocks: Sequence[Block], max_fuse_qubits: int = 3, max_diag_qubits: int = 10,
         max_mux_selectors: int = 20) -> List[Block]:
    """Greedy peephole fusion.  For every incoming block the *earliest legal position* ``pos`` is
    found by scanning back over blocks it commutes with; it may then be merged into any
    compatible block at or after ``pos`` (diagonals, RY multiplexers) or into the block that stops
    it (small dense gates), or it is inserted at ``pos`` (1-qubit gates -- this is what lets the
    QFT's later phase ladders merge into earlier tables) or appended."""
    out: List[Block] = []
    for b in blocks:
        if b.kind in ("init", "gemm", "param"):
            out.append(b)
            continue
        pos = 0
        for i in range(len(out) - 1, -1, -1):
            prev = out[i]
            if prev.kind in ("init", "gemm") or not _commutes(prev, b):
                if not (prev.kind == "muxry" and b.kind == "muxry" and prev.targets == b.targets):
                    pos = i + 1
                    break
        blocker = out[pos - 1] if pos > 0 else None

        if b.kind == "diag":
            done = False
            for i in range(pos, len(out)):
                prev = out[i]
                if prev.kind != "diag":
                    continue
                if prev.controls == b.controls:
                    union = set(prev.targets) | set(b.targets)
                    if len(union) + len(prev.controls) <= max_diag_qubits:
                        prev.targets = tuple(sorted(union))
                        prev.factors = prev.factors + b.factors
                        prev.n_source += b.n_source
                        done = True
                        break
                else:
                    fp, fb = _fold_diag_block(prev), _fold_diag_block(b)
                    union = set(fp.targets) | set(fb.targets)
                    if len(union) <= max_diag_qubits:
                        out[i] = Block("diag", tuple(sorted(union)), factors=fp.factors + fb.factors,
                                       n_source=prev.n_source + b.n_source)
                        done = True
                        break
            if not done:
                out.append(Block("diag", tuple(sorted(b.targets)), dict(b.controls), factors=list(b.factors),
                                 n_source=b.n_source))
            continue

        if b.kind == "muxry":
            done = False
            for i in range(pos, len(out)):
                prev = out[i]
                if prev.kind == "muxry" and prev.targets == b.targets:
                    union = list(prev.selectors) + [q for q in b.selectors if q not in prev.selectors]
                    if len(union) <= max_mux_selectors:
                        prev.angles = _mux_expand(prev.angles, prev.selectors, union) + _mux_expand(b.angles, b.selectors, union)
                        prev.selectors = tuple(union)
                        prev.n_source += b.n_source
                        done = True
                        break
            if not done:
                out.append(Block("muxry", b.targets, selectors=b.selectors, angles=np.array(b.angles), n_source=b.n_source))
            continue

        # dense
        cand = blocker
        if cand is not None and cand.kind == "dense":
            union = tuple(sorted(cand.support | b.support))
            small = len(b.support) <= 2 and len(cand.support) <= max_fuse_qubits
            tunion = tuple(sorted(set(cand.targets) | set(b.targets)))
            if cand.controls == b.controls and len(tunion) <= max(max_fuse_qubits, len(cand.targets), len(b.targets)):
                # same controls: multiply the target matrices, keep the controls symbolic
                ma = _expand_dense(Block("dense", cand.targets, {}, cand.matrix), tunion)
                mb = _expand_dense(Block("dense", b.targets, {}, b.matrix), tunion)
                cand.targets, cand.matrix = tunion, mb @ ma
                cand.n_source += b.n_source
                continue
            if small and len(union) <= max_fuse_qubits:
                ma = _expand_dense(cand, union)
                mb = _expand_dense(b, union)
                cand.targets, cand.controls, cand.matrix = union, {}, mb @ ma
                cand.n_source += b.n_source
                continue
        nb_ = Block("dense", b.targets, dict(b.controls), b.matrix, n_source=b.n_source)
        if len(b.targets) == 1 and not b.controls and pos < len(out) and all(o.kind == "diag" for o in out[pos:]):
            out.insert(pos, nb_)  # hop over trailing diagonal tables it commutes with
        else:
            out.append(nb_)
    # a fused dense block may have become diagonal or the identity
    final: List[Block] = []
    for b in out:
        if b.kind == "dense" and not b.controls and _is_diagonal(b.matrix):
            d = np.diagonal(b.matrix).copy()
            if np.allclose(d, 1.0, atol=0, rtol=0):
                continue
            final.append(Block("diag", b.targets, factors=[(b.targets, d)], n_source=b.n_source))
        else:
            final.append(b)
    return final


def tensor_merge(blocks: Sequence[Block], max_fuse_qubits: int = 3) -> List[Block]:
    """Tensor-product fusion inside one pass: an unconditional 1-2 qubit gate is multiplied into
    an earlier unconditional small gate it can reach by commutation -- one shared-memory sweep
    instead of two for the same traffic.  FP64 cost model: a complex k=2 or a real k=3 block is
    still balanced against the sweep's shared-memory time, so those are the caps."""
    out: List[Block] = []
    for b in blocks:
        if b.kind == "dense" and not b.controls and len(b.targets) <= 2:
            pos = 0
            for i in range(len(out) - 1, -1, -1):
                if not _commutes(out[i], b):
                    pos = i + 1
                    break
            b_real = not np.any(b.matrix.imag)
            merged = False
            for i in range(pos, len(out)):
                prev = out[i]
                if prev.kind != "dense" or prev.controls:
                    continue
                tunion = tuple(sorted(set(prev.targets) | set(b.targets)))
                both_real = b_real and not np.any(prev.matrix.imag)
                if len(tunion) <= min(max_fuse_qubits, 3 if both_real else 2):
                    ma = _expand_dense(prev, tunion)
                    mb = _expand_dense(b, tunion)
                    out[i] = Block("dense", tunion, {}, mb @ ma, n_source=prev.n_source + b.n_source)
                    merged = True
                    break
            if merged:
                continue
        out.append(b)
    return out


def split_table_controls(qubits: Sequence[int], table: np.ndarray):
    """A table qubit whose 0-half (or 1-half) is identically 1 is really a control: e.g. every
    controlled-phase ladder of the QFT is the identity on the control = 0 half.  Returns the reduced
    (qubits, table) and ``{qubit: value}`` -- half the lookups and multiplies per extracted qubit."""
    qubits = list(qubits)
    tab = np.asarray(table)
    controls: Dict[int, int] = {}
    changed = True
    while changed and qubits:
        changed = False
        for j in range(len(qubits)):
            view = tab.reshape(-1, 2, 2**j)  # [high bits][bit j][low bits]
            zero_half, one_half = view[:, 0, :], view[:, 1, :]
            if np.all(zero_half == 1):
                controls[qubits[j]] = 1
                tab = one_half.reshape(-1)
            elif np.all(one_half == 1):
                controls[qubits[j]] = 0
                tab = zero_half.reshape(-1)
            else:
                continue
            del qubits[j]
            changed = True
            break
    return qubits, tab, controls


def _mux_expand(angles: np.ndarray, sel: Sequence[int], union: Sequence[int]) -> np.ndarray:
    """Angle table over ``union`` (bit j = union[j]) of a table over ``sel``."""
    idx = np.arange(2 ** len(union))
    sub = np.zeros_like(idx)
    upos = {q: j for j, q in enumerate(union)}
    for j, q in enumerate(sel):
        sub |= ((idx >> upos[q]) & 1) << j
    return np.asarray(angles, dtype=float)[sub]


# --------------------------------------------------------------------------------------
# Pass grouping + emission
# --------------------------------------------------------------------------------------


@dataclass
class LoweringOptions:
    dtype: str = "fp64"  # "fp64" (complex128) | "fp32" (complex64)
    tile_bits: int = 0  # 0 -> 12 (fp64) / 13 (fp32)
    max_high: int = 8  # tile qubits above the contiguous low run (runs of >= 2^4 amplitudes = 256 B; 5 -> 2 KB runs: fewer ops per
    # pass and a higher HBM fraction per pass, but 8 more passes and 12 % more time per 33-qubit solve: profiles/README.md)
    max_fuse_qubits: int = 2  # 2 keeps every block expressible as register-stage micro-ops
    max_diag_qubits: int = 10
    n_local: int = 0  # 0 -> all qubits local (single GPU)
    batch: bool = False  # one whole-state pass for qls_batch_expect_z (parameterised, n <= 12/13)
    register_stages: bool = True  # also emit the register-stage form of passes that can use it
    zero_start: bool = False  # the program runs on |0...0> (Statevector(circuit)); a leading state preparation implies it
    awake_after_init: Tuple[int, ...] = ()  # positions that are NOT |0> after the state preparation (folded first gates)

    def resolved_tile_bits(self) -> int:
        if self.tile_bits:
            return self.tile_bits
        if self.dtype == "fp64":
            return int(os.environ.get("QLS_TILE_BITS", "0")) or 12
        return 13


@dataclass
class Step:
    kind: str  # "passes" | "init" | "gemm"
    pass_begin: int = 0
    pass_end: int = 0
    qubits: Tuple[int, ...] = ()
    amps: Optional[np.ndarray] = None
    nb: int = 0
    ctrl_mask: int = 0
    ctrl_val: int = 0
    matrix: Optional[np.ndarray] = None
    # init on a sharded state: rank r writes amps * rank_factors[r] (None: rank 0 only).  Set by the shard planner when
    # the first gate on a global qubit that is still |0> is folded into the state preparation.
    rank_factors: Optional[List[complex]] = None


@dataclass
class PassInfo:
    """Host-side description of one emitted pass (accounting for the roofline)."""
    tile_bits: int
    low_bits: int
    high: Tuple[int, ...]
    fixed_bits: int
    n_ops: int
    n_source: int
    op_controls: List[int]
    heavy: bool
    rs_stages: int = 0  # > 0: the pass also has a register-stage form with this many stages
    touched_fraction: float = 1.0  # tiles on which at least one op fires (the register-stage kernel skips the rest)


@dataclass
class LoweredProgram:
    num_qubits: int
    n_local: int
    dtype: str
    steps: List[Step]
    passes: "C.Array"
    ops: "C.Array"
    rs_ops: "C.Array"
    n_rs_ops: int
    pool: np.ndarray
    pass_info: List[PassInfo]
    n_source_gates: int
    global_phase: float
    zero_start: bool = False  # lowered for |0...0> input: passes skip tiles that a still-|0> qubit makes zero

    @property
    def n_passes(self) -> int:
        return len(self.pass_info)

    def amp_bytes(self) -> int:
        return 16 if self.dtype == "fp64" else 8

    def pass_bytes(self, i: int, n_local: Optional[int] = None) -> int:
        """Algorithmic HBM bytes of pass ``i``: one read + one write of every touched amplitude
        (``2*S*2**(n-c)``)."""
        n = self.n_local if n_local is None else n_local
        pi = self.pass_info[i]
        return int(2 * self.amp_bytes() * 2 ** (n - pi.fixed_bits) * pi.touched_fraction)

    def total_pass_bytes(self) -> int:
        return sum(self.pass_bytes(i) for i in range(self.n_passes))

    def fp64_fma_count(self, n_local: Optional[int] = None) -> float:
        """Real fused multiply-adds of the register-stage form (complex 4x4: 16 per amplitude, real
        4x4: 8, complex 2x2: 8, real 2x2: 4, table lookups: 4), halved per control.  Passes without a
        register-stage form are not counted."""
        return sum(self.pass_fma_count(i, n_local) for i in range(self.n_passes))

    def pass_fma_count(self, i: int, n_local: Optional[int] = None) -> float:
        """Real fused multiply-adds of pass ``i`` (register-stage form; see ``fp64_fma_count``)."""
        n = self.n_local if n_local is None else n_local
        per = {_abi.QLS_RS_G1: (8, 4), _abi.QLS_RS_G2: (16, 8), _abi.QLS_RS_DIAG: (4, 4), _abi.QLS_RS_MUXRY: (4, 4)}
        tot = 0.0
        p = self.passes[i]
        for o in range(p.rs_begin, p.rs_begin + p.rs_count):
            op = self.rs_ops[o]
            if op.kind not in per:
                continue
            c = per[op.kind][1 if (op.flags & _abi.QLS_OP_FLAG_REAL) else 0]
            # controls outside the tile and the pass-level fixed bits (shared controls, qubits still |0>) restrict the tiles
            nc = bin(op.xc_mask | p.fix_mask).count("1") + bin(op.ctl_tmask).count("1") + bin(op.ctl_rmask).count("1")
            tot += c * 0.5**nc * 2.0**n
        return tot

    def amp_updates(self) -> int:
        """Sum over fused blocks of ``2**(n-c)``."""
        tot = 0
        for p in self.pass_info:
            for c in p.op_controls:
                tot += 2 ** (self.n_local - c)
        for s in self.steps:
            if s.kind == "gemm":
                tot += 2 ** (self.n_local - bin(s.ctrl_mask & (2**self.n_local - 1)).count("1"))
        return tot


def _segments(pairs: List[Tuple[int, int]]) -> List[Tuple[int, int, int]]:
    """(src, dst) single-bit moves -> merged (src, width, dst) fields."""
    segs: List[List[int]] = []
    for src, dst in pairs:
        if segs and segs[-1][0] + segs[-1][1] == src and segs[-1][2] + segs[-1][1] == dst:
            segs[-1][1] += 1
        else:
            segs.append([src, 1, dst])
    return [tuple(s) for s in segs]


class _RsUnsupported(Exception):
    pass


class _Emitter:
    def __init__(self, n: int, opts: LoweringOptions):
        self.n = n
        self.opts = opts
        self.n_local = opts.n_local or n
        self.T = self.n_local if opts.batch else min(opts.resolved_tile_bits(), self.n_local)
        self.Hmax = 0 if opts.batch else min(int(os.environ.get("QLS_MAX_HIGH", opts.max_high)), self.T)
        self.Lmin = self.T - self.Hmax
        if opts.dtype == "fp32" and self.Lmin < 1:
            self.Lmin = min(1, self.T)
            self.Hmax = self.T - self.Lmin
        self.cdtype = np.complex128 if opts.dtype == "fp64" else np.complex64
        self.passes: List[_abi.QlsPass] = []
        self.ops: List[_abi.QlsOp] = []
        self.rs_ops: List[_abi.QlsRsOp] = []
        self.rs_T = opts.resolved_tile_bits()  # the register-stage kernel exists for 11 / 12 (complex128) and 13 (complex64)
        # 2^KR amplitudes per thread, 2^(T - KR) threads per tile (csrc/rs_pass.cu): complex64 5, complex128 4
        # (complex128 with 5 = 128 threads x 32 amplitudes halves the stage count but leaves 8 warps per SM:
        # measured slower, profiles/README.md)
        self.rs_KR = 5 if opts.dtype == "fp32" else 4
        self.pool = bytearray()
        self.pass_info: List[PassInfo] = []
        self.steps: List[Step] = []
        self.cur: List[Block] = []
        self.cur_high: set = set()
        # qubits known to be |0> (never acted on non-diagonally since the state was prepared); None: unknown
        self.zero: Optional[set] = None
        self.bzero: Optional[set] = None  # batch mode: qubits still |0> (see add)

    # -- pool -----------------------------------------------------------------------------
    def _pool_add(self, arr: np.ndarray) -> int:
        while len(self.pool) % 16:
            self.pool.append(0)
        off = len(self.pool)
        self.pool += np.ascontiguousarray(arr, dtype=self.cdtype).tobytes()
        return off

    # -- grouping -------------------------------------------------------------------------
    def _fit(self, needed: set, l_floor: int):
        """Largest contiguous low run L >= l_floor such that the touched qubits >= L fit the
        remaining T - L tile bits; returns (L, high qubits) or None."""
        n_loc = self.n_local
        for L in range(self.T, l_floor - 1, -1):
            high = {q for q in needed if q >= L}
            if len(high) <= self.T - L and self.T - L <= _abi.QLS_MAX_HIGH:
                q = L
                while len(high) < self.T - L:  # pad with the next unused qubits above the low run
                    if q >= n_loc:
                        break
                    if q not in high:
                        high.add(q)
                    q += 1
                if len(high) == self.T - L:
                    return L, high
        return None

    def add(self, b: Block) -> None:
        if self.opts.batch:
            if b.kind == "init":
                raise NotImplementedError("state preparation inside a batched circuit")
            # Batched circuits start from |0...0>.  A qubit nothing has acted on non-diagonally yet is still |0>: every
            # amplitude with that bit set is zero, so a block that does not use the qubit may skip them -- expressed as a
            # control on value 0 (exact; the Hadamard-test ancilla stays |0> through the whole ansatz: half the work)
            if self.bzero is None:
                self.bzero = set(range(self.n))
            if b.kind in ("dense", "param", "gemm"):
                extra = {q: 0 for q in self.bzero if q not in b.support}
                if any(v == 1 and q in self.bzero for q, v in b.controls.items()):
                    return  # controlled on a qubit that is still |0>: the identity
                if extra:
                    b = dataclasses.replace(b, controls={**b.controls, **extra})
            self.bzero -= set(b.touched)
            self.cur.append(b)
            self.cur_high |= set(b.touched)
            return
        if self.zero is not None and b.kind != "init" and b.controls:
            # a control on a qubit that is still |0>: value 1 never fires (the block is the identity),
            # value 0 always holds (drop the control)
            if any(v == 1 and q in self.zero for q, v in b.controls.items()):
                return
            if any(q in self.zero for q in b.controls):
                b = dataclasses.replace(b, controls={q: v for q, v in b.controls.items() if q not in self.zero})
        if b.kind in ("init", "gemm"):
            self.flush()
            if b.kind == "init":
                self.steps.append(Step("init", qubits=b.targets, amps=b.amps))
                self.zero = set(range(self.n)) - set(b.targets) - set(self.opts.awake_after_init)
            else:
                if tuple(b.targets) != tuple(range(len(b.targets))):
                    raise ValueError(f"a GEMM block acts on positions 0..nb-1 (csrc/ctrl_gemm.cu), got targets {tuple(b.targets)}")
                if self.zero is not None:
                    self.zero -= set(b.targets)
                mask = sum(1 << q for q in b.controls)
                val = sum(v << q for q, v in b.controls.items())
                self.steps.append(Step("gemm", nb=len(b.targets), ctrl_mask=mask, ctrl_val=val, matrix=b.matrix))
            return
        if any(q >= self.n_local for q in b.touched):
            raise ValueError(f"block acts on non-local qubit(s) {sorted(b.touched)}; remap first (shard planner)")
        need = set(b.touched)
        if self.cur and self._fit(self.cur_high | need, self.Lmin) is None:
            self.flush()
        floor = 1 if self.opts.dtype == "fp32" else 0
        if self._fit(self.cur_high | need, min(floor, self.T)) is None:
            raise ValueError(f"block touching {sorted(need)} does not fit a tile of {self.T} bits")
        if self.zero is not None:
            self.zero -= need  # eagerly (later blocks of the same pass may be controlled by these qubits), after the flush above
        self.cur.append(b)
        self.cur_high |= need

    MAX_EXT_CONTROLS = 6  # csrc/rs_pass.cu keeps one step list per assignment of the controls outside the tile

    def _ctl_prefix(self, blocks: Sequence[Block]) -> int:
        """Largest k such that blocks[:k] have at most MAX_EXT_CONTROLS distinct control qubits outside
        the qubits they touch (controls shared by all of them become pass-level fixed bits: free)."""
        touched: set = set()
        ctl: Dict[int, set] = {}
        common: Optional[Dict[int, int]] = None
        for k, b in enumerate(blocks):
            touched |= set(b.touched)
            bc = dict(b.controls) if b.kind in ("dense", "diag") else {}
            for q, v in bc.items():
                ctl.setdefault(q, set()).add(v)
            common = bc if common is None else {q: v for q, v in common.items() if bc.get(q) == v}
            ext = [q for q in ctl if q not in touched and q not in common]
            if len(ext) > self.MAX_EXT_CONTROLS:
                return max(k, 1)
        return len(blocks)

    def flush(self) -> None:
        if not self.cur:
            return
        if not self.opts.batch and self.opts.register_stages:
            k = self._ctl_prefix(self.cur)
            if k < len(self.cur):
                rest = self.cur[k:]
                self.cur = self.cur[:k]
                self.cur_high = set().union(*[set(b.touched) for b in self.cur])
                self._flush_one()
                self.cur = rest
                self.cur_high = set().union(*[set(b.touched) for b in self.cur])
                self.flush()
                return
        self._flush_one()

    def _flush_one(self) -> None:
        blocks, self.cur = tensor_merge(self.cur, self.opts.max_fuse_qubits), []
        needed = set(self.cur_high)
        self.cur_high = set()
        T = self.T
        floor = 1 if self.opts.dtype == "fp32" else 0
        fit = self._fit(needed, self.Lmin) or self._fit(needed, min(floor, T))
        L, high = fit
        high_sorted = tuple(sorted(high))
        L = T - len(high_sorted)
        local_pos = {q: q for q in range(L)}
        for j, q in enumerate(high_sorted):
            local_pos[q] = L + j

        # pass-level controls: external control bits shared by every op (tile skipping)
        common: Optional[Dict[int, int]] = None
        for b in blocks:
            ext = ({q: v for q, v in b.controls.items() if q not in local_pos and q < self.n_local}
                   if b.kind in ("dense", "diag") else {})
            common = ext if common is None else {q: v for q, v in common.items() if ext.get(q) == v}
        common = common or {}
        # qubits that are still |0> and that no block of the pass acts on: amplitudes with such a bit set are
        # zero before and after the pass, so the pass only enumerates the tiles where they are 0 (exact; the
        # bytes of the skipped tiles are not counted either)
        zero_fix: List[int] = []
        if self.zero is not None and not self.opts.batch:
            touched_pass = set().union(*[set(b.touched) for b in blocks]) if blocks else set()
            zero_fix = sorted(q for q in self.zero if q < self.n_local and q not in local_pos and q not in touched_pass
                              and q not in common)
            while zero_fix:
                free_try = [q for q in range(self.n_local) if q not in local_pos and q not in common and q not in zero_fix]
                if len(_segments([(j, q) for j, q in enumerate(free_try)])) <= _abi.QLS_MAX_FSEG:
                    break
                zero_fix.pop(0)  # too many index segments: give up the lowest one
            for q in zero_fix:
                common[q] = 0
        fix_mask = sum(1 << q for q in common)
        fix_val = sum(v << q for q, v in common.items())

        op_begin = len(self.ops)
        op_controls: List[int] = []
        heavy = False
        n_source = 0
        # Diagonal tables ride on a neighbouring unconditional dense block (applied to its inputs /
        # outputs in registers) or are chained onto one another: one tile sweep per host.
        pending: List[_abi.QlsOp] = []  # tables waiting for a host
        host_room = 0  # attachments the last emitted host can still take (0: no open host)

        def flush_pending() -> None:
            nonlocal host_room
            while pending:
                head = pending.pop(0)
                self.ops.append(head)
                take = pending[:_abi.QLS_MAX_ATTACH]
                del pending[:_abi.QLS_MAX_ATTACH]
                for t in take:
                    t.flags |= _abi.QLS_OP_FLAG_ATTACH_POST
                    self.ops.append(t)
            host_room = 0

        for b in blocks:
            n_source += b.n_source
            op_controls.append(len(b.controls))
            if b.kind == "diag":
                bb = b if all(q in common for q in b.controls) else _fold_diag_block(b)
                for op, _ in self._emit_block(bb, local_pos, T):
                    if host_room > 0:
                        op.flags |= _abi.QLS_OP_FLAG_ATTACH_POST
                        self.ops.append(op)
                        host_room -= 1
                    else:
                        pending.append(op)
                continue
            unconditional = b.kind == "dense" and all(q in common for q in b.controls)
            if unconditional:
                keep = pending[-_abi.QLS_MAX_ATTACH:]
                del pending[-_abi.QLS_MAX_ATTACH:]
                flush_pending()
                (op, _), = self._emit_block(b, local_pos, T)
                self.ops.append(op)
                for t in keep:
                    t.flags |= _abi.QLS_OP_FLAG_ATTACH_PRE
                    self.ops.append(t)
                host_room = _abi.QLS_MAX_ATTACH - len(keep)
            else:
                flush_pending()
                for op, _ in self._emit_block(b, local_pos, T):
                    self.ops.append(op)
            heavy |= b.kind == "dense" and len(b.targets) >= 4
        flush_pending()

        rs_begin = len(self.rs_ops)
        rs_stages = 0
        rs_ok = self.rs_T in ((11, 12) if self.opts.dtype == "fp64" else (13,)) and not (self.rs_T == 11 and self.rs_KR != 4)
        if self.opts.register_stages and not self.opts.batch and T == self.rs_T and rs_ok:
            rs_stages = self._emit_register_stages(blocks, local_pos, common)
        p = _abi.QlsPass()
        p.rs_begin, p.rs_count = rs_begin, len(self.rs_ops) - rs_begin
        p.tile_bits, p.low_bits, p.n_high = T, L, len(high_sorted)
        for j, q in enumerate(high_sorted):
            p.high[j] = q
        p.op_begin, p.op_count = op_begin, len(self.ops) - op_begin
        p.fix_mask, p.fix_val = fix_mask, fix_val
        free = [q for q in range(self.n_local) if q not in local_pos and not (fix_mask >> q) & 1]
        segs = _segments([(j, q) for j, q in enumerate(free)])
        if len(segs) > _abi.QLS_MAX_FSEG:
            raise ValueError("tile counter deposit needs too many segments")
        p.n_fseg = len(segs)
        for j, (src, w, dst) in enumerate(segs):
            p.fseg[j].src, p.fseg[j].width, p.fseg[j].dst = src, w, dst
        idx = len(self.passes)
        self.passes.append(p)
        touched = 1.0
        if rs_stages:
            # controls outside the tile: a tile whose control qubits switch every op off is skipped by the kernel
            recs = [r for r in self.rs_ops[rs_begin:] if r.kind != _abi.QLS_RS_STAGE]
            cmask = 0
            for r in recs:
                cmask |= r.xc_mask & ~fix_mask
            cb = [q for q in range(64) if (cmask >> q) & 1]
            if recs and all(r.xc_mask & ~fix_mask for r in recs) and len(cb) <= self.MAX_EXT_CONTROLS:
                live = 0
                for v in range(2 ** len(cb)):
                    a = sum(((v >> j) & 1) << q for j, q in enumerate(cb))
                    live += any((a & r.xc_mask & ~fix_mask) == (r.xc_val & ~fix_mask) for r in recs)
                touched = live / 2 ** len(cb)
        self.pass_info.append(PassInfo(T, L, high_sorted, len(common), p.op_count, n_source, op_controls, heavy, rs_stages, touched))
        if self.steps and self.steps[-1].kind == "passes" and self.steps[-1].pass_end == idx:
            self.steps[-1].pass_end = idx + 1
        else:
            self.steps.append(Step("passes", pass_begin=idx, pass_end=idx + 1))

    # -- register-stage form --------------------------------------------------------------
    def _swz(self, i: int) -> int:
        sh = 0 if self.opts.dtype == "fp64" else 1
        u = i >> sh
        return i ^ ((((u >> 3) ^ (u >> 6) ^ (u >> 9)) & 7) << sh)

    def _emit_register_stages(self, blocks: Sequence[Block], local_pos: Dict[int, int], common: Dict[int, int]) -> int:
        """Emit the pass as stages of register-resident micro-ops (csrc/rs_pass.cu).  Returns the
        number of stages, or 0 (and emits nothing) if some block cannot be expressed."""
        T, KR = self.rs_T, self.rs_KR
        for b in blocks:
            if b.kind == "dense" and len(b.targets) > 2:
                return 0
            if b.kind not in ("dense", "diag", "muxry"):
                return 0
        # greedy stage partition: a block joins the stage if its targets still fit the register qubits
        stages: List[Tuple[set, List[Block]]] = []
        cur_r: set = set()
        cur_b: List[Block] = []
        for b in blocks:
            need = {local_pos[q] for q in b.touched}
            if len(cur_r | need) > KR:
                stages.append((cur_r, cur_b))
                cur_r, cur_b = set(), []
            cur_r |= need
            cur_b.append(b)
        stages.append((cur_r, cur_b))
        if len(stages) > 24:
            return 0
        mark = len(self.rs_ops)
        n_records = 0
        try:
            for rset, sblocks in stages:
                regq = set(rset)
                pos = T - 1
                while len(regq) < KR:  # pad with the highest unused tile positions (lanes keep the low bits)
                    if pos not in regq:
                        regq.add(pos)
                    pos -= 1
                regq_sorted = sorted(regq)
                rho = {p_: j for j, p_ in enumerate(regq_sorted)}
                tpos = [p_ for p_ in range(T) if p_ not in regq]
                # global <-> register access without shared-memory staging: the lanes are the five lowest
                # non-register tile bits; every 32-byte sector (64 bytes with bit 1) of a warp access is fully
                # used as long as the lowest tile bits are lane bits, whatever the higher register qubits are
                low_lane_bits = 2 if self.opts.dtype == "fp64" else 3
                direct = min(regq_sorted) >= low_lane_bits
                if not direct and self.opts.dtype == "fp64":
                    tpos = self._conflict_free_lanes(tpos)
                tbit = {p_: j for j, p_ in enumerate(tpos)}
                st = _abi.QlsRsOp()
                st.kind = _abi.QLS_RS_STAGE
                for j, p_ in enumerate(regq_sorted[:4]):
                    st.r[j] = p_
                if KR > 4:
                    st.pad0 = regq_sorted[4]
                    st.flags |= 0x20  # five register qubits
                if direct:
                    st.flags |= 0x10  # direct global <-> register access is coalesced (lanes = tile bits 0..4)
                segs = _segments([(j, p_) for j, p_ in enumerate(tpos)])
                if len(segs) > _abi.QLS_MAX_SEG:
                    raise _RsUnsupported("thread deposit needs too many segments")
                st.n_tseg = len(segs)
                for j, (src, w, dst) in enumerate(segs):
                    st.tseg[j].src, st.tseg[j].width, st.tseg[j].dst = src, w, dst
                for r in range(2**KR):
                    off = sum(((r >> j) & 1) << regq_sorted[j] for j in range(KR))
                    st.rt[r] = self._swz(off)
                self.rs_ops.append(st)
                for b in sblocks:
                    for op in self._emit_rs_block(b, local_pos, rho, tbit, common):
                        self.rs_ops.append(op)
                n_records = len(self.rs_ops) - mark
                if n_records > 128:
                    raise _RsUnsupported("too many micro-ops")
        except _RsUnsupported:
            del self.rs_ops[mark:]
            return 0
        return len(stages)

    @staticmethod
    def _conflict_free_lanes(tpos: List[int]) -> List[int]:
        """Order of the thread bits of a stage that is exchanged through shared memory.  A 128-bit
        access is served per quarter warp (lanes 0..7 = thread bits 0..2) and the swizzle of
        csrc/rs_pass.cu puts 16-byte unit u into bank group (u ^ u>>3 ^ u>>6 ^ u>>9) & 7: the eight
        lanes hit eight different groups iff the tile positions of thread bits 0..2 are pairwise
        different mod 3.  Keep the ascending order when that already holds (or cannot be had)."""
        if len({p_ % 3 for p_ in tpos[:3]}) == 3:
            return tpos
        pick: List[int] = []
        for res in range(3):
            cand = [p_ for p_ in tpos if p_ % 3 == res]
            if not cand:
                return tpos
            pick.append(cand[0])
        pick.sort()
        out = pick + [p_ for p_ in tpos if p_ not in pick]
        if len(_segments([(j, p_) for j, p_ in enumerate(out)])) > _abi.QLS_MAX_SEG:
            return tpos
        return out

    def _rs_controls(self, op: _abi.QlsRsOp, controls: Dict[int, int], local_pos, rho, tbit) -> None:
        for q, v in controls.items():
            lp = local_pos.get(q)
            if lp is None:
                op.xc_mask |= 1 << q
                op.xc_val |= v << q
            elif lp in rho:
                op.ctl_rmask |= 1 << rho[lp]
                op.ctl_rval |= v << rho[lp]
            else:
                op.ctl_tmask |= 1 << tbit[lp]
                op.ctl_tval |= v << tbit[lp]

    def _rs_table_fields(self, op: _abi.QlsRsOp, qubits: Sequence[int], local_pos, rho, tbit) -> bool:
        """Index fields of a table whose bit j is ``qubits[j]``: thread bits -> tseg, register
        qubits -> rt[], everything else -> xseg on the global base index."""
        if len(qubits) > 30:
            return False
        tpairs, xpairs, rbits = [], [], []
        for j, q in enumerate(qubits):
            lp = local_pos.get(q)
            if lp is None:
                xpairs.append((q, j))
            elif lp in rho:
                rbits.append((rho[lp], j))
            else:
                tpairs.append((tbit[lp], j))
        tpairs.sort()
        ts, xs = _segments(tpairs), _segments(xpairs)
        if len(ts) > _abi.QLS_MAX_SEG or len(xs) > _abi.QLS_MAX_SEG:
            return False
        if any(j >= 16 for _, j in rbits):  # rt[] holds 16-bit words: register qubits must index the low table bits
            return False
        op.n_tseg, op.n_xseg = len(ts), len(xs)
        for j, (src, w, dst) in enumerate(ts):
            op.tseg[j].src, op.tseg[j].width, op.tseg[j].dst = src, w, dst
        for j, (src, w, dst) in enumerate(xs):
            op.xseg[j].src, op.xseg[j].width, op.xseg[j].dst = src, w, dst
        for r in range(2**self.rs_KR):
            op.rt[r] = sum(((r >> rb) & 1) << j for rb, j in rbits)
        return True

    def _emit_rs_block(self, b: Block, local_pos, rho, tbit, common):
        if b.kind == "dense":
            op = _abi.QlsRsOp()
            regs = [rho[local_pos[q]] for q in b.targets]
            m = b.matrix
            if len(regs) == 1:
                op.kind = _abi.QLS_RS_G1
                op.r[0] = regs[0]
                if not np.any(m.imag):
                    op.flags |= _abi.QLS_OP_FLAG_REAL
            else:
                op.kind = _abi.QLS_RS_G2
                if regs[0] > regs[1]:
                    m = _permute_matrix(m, [1, 0])
                    regs = [regs[1], regs[0]]
                op.r[0], op.r[1] = regs
                if not np.any(m.imag):
                    op.flags |= _abi.QLS_OP_FLAG_REAL
            op.pool_off = self._pool_add(m)
            self._rs_controls(op, b.controls, local_pos, rho, tbit)
            yield op
        elif b.kind == "muxry":
            op = _abi.QlsRsOp()
            op.kind = _abi.QLS_RS_MUXRY
            op.r[0] = rho[local_pos[b.targets[0]]]
            # selectors inside the tile index the low table bits, the ones outside the high bits: a table over more than 16
            # selectors (17-18 clock qubits on a sharded 34-36 qubit state) keeps its register-qubit fields within rt[]'s
            # 16-bit words -- without this the whole pass fell back to the shared-memory sweep kernel
            sel = sorted(b.selectors, key=lambda q: (q not in local_pos, q))
            ang = _mux_expand(b.angles, b.selectors, sel) if list(sel) != list(b.selectors) else b.angles
            if not self._rs_table_fields(op, sel, local_pos, rho, tbit):
                raise _RsUnsupported("multiplexer table fields")
            op.pool_off = self._pool_add(np.cos(ang / 2) + 1j * np.sin(ang / 2))
            yield op
        elif b.kind == "diag":
            bb = b if all(q in common for q in b.controls) else _fold_diag_block(b)
            cur: List[Tuple[Tuple[int, ...], np.ndarray]] = []
            groups: List[List[Tuple[Tuple[int, ...], np.ndarray]]] = []
            for f in bb.factors:
                trial = cur + [f]
                qs = sorted({q for fq, _ in trial for q in fq})
                if len(qs) <= self.opts.max_diag_qubits and self._rs_table_fields(_abi.QlsRsOp(), qs, local_pos, rho, tbit):
                    cur = trial
                else:
                    if not cur:
                        raise _RsUnsupported("diagonal table fields")
                    groups.append(cur)
                    cur = [f]
            if cur:
                groups.append(cur)
            for g in groups:
                qs = sorted({q for fq, _ in g for q in fq})
                tab = np.ones(2 ** len(qs), dtype=complex)
                pos = {q: j for j, q in enumerate(qs)}
                idx = np.arange(tab.size)
                for fq, ft in g:
                    sub = np.zeros_like(idx)
                    for j, q in enumerate(fq):
                        sub |= ((idx >> pos[q]) & 1) << j
                    tab *= ft[sub]
                op = _abi.QlsRsOp()
                op.kind = _abi.QLS_RS_DIAG
                qs2, tab2, ctl = split_table_controls(qs, tab)
                self._rs_table_fields(op, qs2, local_pos, rho, tbit)
                self._rs_controls(op, ctl, local_pos, rho, tbit)
                op.pool_off = self._pool_add(tab2)
                yield op
        else:
            raise _RsUnsupported(b.kind)

    # -- op emission ----------------------------------------------------------------------
    def _controls(self, op: _abi.QlsOp, controls: Dict[int, int], local_pos: Dict[int, int]) -> None:
        for q, v in controls.items():
            if q in local_pos:
                op.lc_mask |= 1 << local_pos[q]
                op.lc_val |= v << local_pos[q]
            else:
                op.xc_mask |= 1 << q
                op.xc_val |= v << q

    def _table_fields(self, op: _abi.QlsOp, qubits: Sequence[int], local_pos: Dict[int, int]) -> bool:
        """Index fields for a table whose bit j is ``qubits[j]`` (ascending).  False if too many."""
        lpairs = [(local_pos[q], j) for j, q in enumerate(qubits) if q in local_pos]
        xpairs = [(q, j) for j, q in enumerate(qubits) if q not in local_pos]
        lpairs.sort()
        ls, xs = _segments(lpairs), _segments(xpairs)
        if len(ls) > _abi.QLS_MAX_SEG or len(xs) > _abi.QLS_MAX_SEG:
            return False
        op.n_lseg, op.n_xseg = len(ls), len(xs)
        for j, (src, w, dst) in enumerate(ls):
            op.lseg[j].src, op.lseg[j].width, op.lseg[j].dst = src, w, dst
        for j, (src, w, dst) in enumerate(xs):
            op.xseg[j].src, op.xseg[j].width, op.xseg[j].dst = src, w, dst
        return True

    def _emit_block(self, b: Block, local_pos: Dict[int, int], T: int):
        if b.kind == "dense":
            op = _abi.QlsOp()
            op.kind, op.k = _abi.QLS_OP_DENSE, len(b.targets)
            lp = [local_pos[q] for q in b.targets]
            order = sorted(range(len(lp)), key=lambda j: lp[j])
            m = b.matrix if order == list(range(len(lp))) else _permute_matrix(b.matrix, order)
            for j, o in enumerate(order):
                op.tq[j] = lp[o]
            if not np.any(m.imag):
                op.flags |= _abi.QLS_OP_FLAG_REAL
            op.pool_off = self._pool_add(m)
            self._controls(op, b.controls, local_pos)
            yield op, len(b.controls)
        elif b.kind == "param":
            name, pidx, coeff, offset = b.param
            op = _abi.QlsOp()
            op.kind, op.k = _abi.QLS_OP_DENSE, 1
            op.flags |= _abi.QLS_OP_FLAG_PARAM
            op.pad0 = _abi.PARAM_KINDS[name]
            op.tq[0] = local_pos[b.targets[0]]
            op.param_index, op.param_coeff, op.param_offset = int(pidx), float(coeff), float(offset)
            self._controls(op, b.controls, local_pos)
            yield op, len(b.controls)
        elif b.kind == "gemm":
            op = _abi.QlsOp()
            op.kind, op.k = _abi.QLS_OP_MATVEC, len(b.targets)
            op.tq[0] = local_pos[b.targets[0]]
            op.pool_off = self._pool_add(b.matrix)
            self._controls(op, b.controls, local_pos)
            yield op, len(b.controls)
        elif b.kind == "muxry":
            op = _abi.QlsOp()
            op.kind, op.k = _abi.QLS_OP_MUXRY, 1
            op.tq[0] = local_pos[b.targets[0]]
            sel = sorted(b.selectors)
            ang = _mux_expand(b.angles, b.selectors, sel) if list(sel) != list(b.selectors) else b.angles
            if not self._table_fields(op, sel, local_pos):
                raise NotImplementedError("multiplexed RY with too many scattered selector qubits")
            tab = np.cos(ang / 2) + 1j * np.sin(ang / 2)  # (cos, sin) pairs
            op.pool_off = self._pool_add(tab)
            yield op, 0
        elif b.kind == "diag":
            # greedily regroup the factors so that every emitted table fits the field limit
            groups: List[List[Tuple[Tuple[int, ...], np.ndarray]]] = []
            cur: List[Tuple[Tuple[int, ...], np.ndarray]] = []
            for f in b.factors:
                trial = cur + [f]
                qs = sorted({q for fq, _ in trial for q in fq})
                probe = _abi.QlsOp()
                if len(qs) <= self.opts.max_diag_qubits and self._table_fields(probe, qs, local_pos):
                    cur = trial
                else:
                    if not cur:
                        raise NotImplementedError(f"diagonal gate on too many scattered qubits: {f[0]}")
                    groups.append(cur)
                    cur = [f]
            if cur:
                groups.append(cur)
            for g in groups:
                qs = sorted({q for fq, _ in g for q in fq})
                tab = np.ones(2 ** len(qs), dtype=complex)
                pos = {q: j for j, q in enumerate(qs)}
                idx = np.arange(tab.size)
                for fq, ft in g:
                    sub = np.zeros_like(idx)
                    for j, q in enumerate(fq):
                        sub |= ((idx >> pos[q]) & 1) << j
                    tab *= ft[sub]
                op = _abi.QlsOp()
                op.kind = _abi.QLS_OP_DIAG
                self._table_fields(op, qs, local_pos)
                op.pool_off = self._pool_add(tab)
                yield op, 0
        else:
            raise ValueError(b.kind)


def lower_blocks(num_qubits: int, blocks: Sequence[Block], opts: LoweringOptions, global_phase: float = 0.0,
                 n_source_gates: int = 0) -> LoweredProgram:
    em = _Emitter(num_qubits, opts)
    blocks = list(blocks)
    if global_phase:
        ph = np.array([np.exp(1j * global_phase)])
        if blocks and blocks[0].kind == "init":
            b0 = blocks[0]
            blocks[0] = Block("init", b0.targets, amps=b0.amps * ph[0])
        else:
            blocks.insert(0, Block("diag", (), factors=[((), ph)], n_source=0))
    if opts.zero_start and not opts.batch:
        em.zero = set(range(num_qubits))
    for b in blocks:
        em.add(b)
    em.flush()
    passes = (_abi.QlsPass * max(1, len(em.passes)))(*em.passes)
    ops = (_abi.QlsOp * max(1, len(em.ops)))(*em.ops)
    rs_ops = (_abi.QlsRsOp * max(1, len(em.rs_ops)))(*em.rs_ops)
    pool = np.frombuffer(bytes(em.pool) if em.pool else b"\0" * 16, dtype=np.uint8).copy()
    return LoweredProgram(num_qubits, em.n_local, opts.dtype, em.steps, passes, ops, rs_ops, len(em.rs_ops), pool,
                          em.pass_info, n_source_gates, global_phase, bool(opts.zero_start and not opts.batch))


def lower_leaves(num_qubits: int, leaves: Sequence[Leaf], global_phase: float = 0.0,
                 opts: Optional[LoweringOptions] = None, fused: bool = True) -> LoweredProgram:
    opts = opts or LoweringOptions()
    blocks = normalise(leaves, batch=opts.batch)
    if fused:
        blocks = fuse(blocks, opts.max_fuse_qubits, opts.max_diag_qubits)
    return lower_blocks(num_qubits, blocks, opts, global_phase, n_source_gates=len(leaves))


def lower_circuit(circuit: QuantumCircuit, opts: Optional[LoweringOptions] = None, fused: bool = True) -> LoweredProgram:
    leaves, phase = flatten(circuit, parametric=bool(opts and opts.batch))
    return lower_leaves(circuit.num_qubits, leaves, phase, opts, fused)
