"""Per-pass specialised CUDA source for register-stage passes, compiled at program-upload time.

The interpreter kernel of round 1 (``csrc/rs_pass.cu``) spends most of its issue slots decoding step
records.  Here every pass of a lowered program becomes its own translation unit: stage layouts,
register indices, table-index bit fields, control masks and the gate matrices are compile-time
constants of straight-line code, compiled with NVRTC for sm_100a through the C ABI
(``qls_jit_compile``) and attached to the program (``qls_program_set_pass_kernel``).  The kernel
around the generated per-tile body lives in ``csrc/rs2_runtime.cuh`` / ``csrc/rs2_kernel.cuh``
(producer warp with bulk asynchronous copies, two consumer groups, three tile buffers).

The generator consumes the SAME ``QlsRsOp`` records the interpreter kernel runs (validated against
the oracle by ``tests/_emulator.py``); it only re-derives a thread-bit order of its own per stage.
Compiled cubins are cached on disk (``jit_cache/``, keyed by a hash of the source), and NVRTC needs
no GPU, so the cache can be filled in the build container (``tools/prewarm_jit.py``).

Reference call site served: ``Statevector(hhl_circuit).data``
(``quantum_linear_systems/implementations/hhl_qiskit_implementation.py:49``).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
CACHE_DIR = os.environ.get("QLS_JIT_CACHE", os.path.join(HERE, "jit_cache"))
KERNEL_NAME = "qls_rs2_pass"
T = 12  # tile bits of the specialised kernels (complex128)
KR = 4  # register qubits per thread
THREADS = 2 * 256 + 128  # two consumer groups + the producer warpgroup
GEN_VERSION = 24  # bump when the generated code or the runtime headers change meaning


# --------------------------------------------------------------------------------------
# IR: a pass in terms of TILE POSITIONS (independent of the interpreter's thread-bit order)
# --------------------------------------------------------------------------------------


@dataclass
class JOp:
    kind: str  # "g1" | "g2" | "diag" | "mux"
    xc_mask: int = 0
    xc_val: int = 0
    tile_ctl: Dict[int, int] = field(default_factory=dict)  # tile position -> required value
    targets: Tuple[int, ...] = ()  # tile positions (g1 / g2: matrix bit j <-> targets[j]; mux: the RY target)
    real: bool = False
    matrix: Optional[np.ndarray] = None
    pool_off: int = 0
    idx_x: List[Tuple[int, int]] = field(default_factory=list)  # (global index bit, table bit)
    idx_t: List[Tuple[int, int]] = field(default_factory=list)  # (tile position, table bit)


@dataclass
class JStage:
    regq: Tuple[int, ...]
    ops: List[JOp]


@dataclass
class JPass:
    L: int
    high: Tuple[int, ...]
    fix_val: int
    fseg: List[Tuple[int, int, int]]
    stages: List[JStage]


def extract_pass(lowered, index: int) -> Optional[JPass]:
    """IR of pass ``index`` or None if it has no (supported) register-stage form."""
    p = lowered.passes[index]
    if lowered.dtype != "fp64" or p.rs_count == 0 or p.tile_bits != T:
        return None
    recs = [lowered.rs_ops[o] for o in range(p.rs_begin, p.rs_begin + p.rs_count)]
    if recs[0].kind != _abi.QLS_RS_STAGE or (recs[0].flags & 0x20):
        return None
    pool = lowered.pool
    fix_mask, fix_val = int(p.fix_mask), int(p.fix_val)
    stages: List[JStage] = []
    regq: Tuple[int, ...] = ()
    tpos_old: Dict[int, int] = {}
    for op in recs:
        if op.kind == _abi.QLS_RS_STAGE:
            if op.flags & 0x20:
                return None
            regq = tuple(int(op.r[j]) for j in range(KR))
            tpos_old = {}
            for s in range(op.n_tseg):
                sg = op.tseg[s]
                for w in range(sg.width):
                    tpos_old[sg.src + w] = sg.dst + w
            if sorted(list(regq) + list(tpos_old.values())) != list(range(T)):
                return None
            stages.append(JStage(regq, []))
            continue
        xm, xv = int(op.xc_mask), int(op.xc_val)
        # controls on pass-level fixed bits hold for every tile of the pass (or never: the op is dead)
        if (xv & fix_mask & xm) != (fix_val & xm & fix_mask):
            continue
        xm &= ~fix_mask
        xv &= ~fix_mask
        ctl: Dict[int, int] = {}
        for j in range(32):
            if (op.ctl_tmask >> j) & 1:
                ctl[tpos_old[j]] = (op.ctl_tval >> j) & 1
            if (op.ctl_rmask >> j) & 1:
                ctl[regq[j]] = (op.ctl_rval >> j) & 1
        real = bool(op.flags & _abi.QLS_OP_FLAG_REAL)
        if op.kind in (_abi.QLS_RS_G1, _abi.QLS_RS_G2):
            k = 1 if op.kind == _abi.QLS_RS_G1 else 2
            m = np.frombuffer(pool, dtype=np.complex128, count=4**k, offset=int(op.pool_off)).reshape(2**k, 2**k).copy()
            stages[-1].ops.append(JOp("g1" if k == 1 else "g2", xm, xv, ctl, tuple(regq[op.r[j]] for j in range(k)), real, m))
            continue
        jop = JOp("diag" if op.kind == _abi.QLS_RS_DIAG else "mux", xm, xv, ctl, pool_off=int(op.pool_off))
        if op.kind == _abi.QLS_RS_MUXRY:
            jop.targets = (regq[op.r[0]],)
        for s in range(op.n_xseg):
            sg = op.xseg[s]
            jop.idx_x += [(sg.src + w, sg.dst + w) for w in range(sg.width)]
        for s in range(op.n_tseg):
            sg = op.tseg[s]
            jop.idx_t += [(tpos_old[sg.src + w], sg.dst + w) for w in range(sg.width)]
        for j in range(KR):
            m_ = int(op.rt[1 << j])
            for b in range(16):
                if (m_ >> b) & 1:
                    jop.idx_t.append((regq[j], b))
        stages[-1].ops.append(jop)
    stages = [s for s in stages if s.ops]
    if not stages:
        return None
    fseg = [(p.fseg[j].src, p.fseg[j].width, p.fseg[j].dst) for j in range(p.n_fseg)]
    return JPass(int(p.low_bits), tuple(int(p.high[j]) for j in range(p.n_high)), fix_val, fseg, stages)


# --------------------------------------------------------------------------------------
# layouts: the two shared-memory images of a tile
# --------------------------------------------------------------------------------------
#
# EXCHANGE image (stage transitions; written and read by the group's threads): unit i lives at
# i ^ (xor over tile positions p >= 3 of bit_p << class(p)) -- an XOR swizzle over all 3-bit fields.
#
# LANDING image (what the producer warp's copies write and read):
#   "tma"   one tensor-map copy (cp.async.bulk.tensor, SWIZZLE_128B) per tile: a bit permutation of the
#           tile-local index (low run, the tensor-map dimensions, then the split bits) followed by the
#           hardware swizzle  u ^ ((u >> 3) & 7);
#   "bulk"  one cp.async.bulk per contiguous run, run r padded by r + r/8 + r/64 units (~58 cycles per
#           copy on B200: only competitive for runs of 2 KB and more).
# A quarter warp (thread bits 0..2) is one 128-byte wavefront of 16-byte accesses; the thread-bit order of
# every stage is chosen so that its eight lanes fall into eight different 16-byte bank groups in both images.

MAX_TMA_OPS = 8


@dataclass
class TmaPlan:
    low_bits: int                       # tile positions [0, low_bits) are dims 0/1 of the tensor map (8 amplitudes x rows)
    dims: List[Tuple[int, int, int]]    # (tile position, global index bit, bits) of dims 2.., ascending
    split: List[Tuple[int, int, int]]   # groups enumerated by separate copies (top of the landing order)
    perm: List[int]                     # tile position -> landing position

    @property
    def n_ops(self) -> int:
        return 1 << sum(b for _, _, b in self.split)


def tma_plan(L: int, high: Sequence[int]) -> Optional[TmaPlan]:
    low = min(L, 11)  # rows of 128 bytes: box dimension 1 holds at most 256 rows
    groups: List[Tuple[int, int, int]] = []
    if L > low:
        groups.append((low, low, L - low))
    j = 0
    while j < len(high):
        k = j
        while k + 1 < len(high) and high[k + 1] == high[k] + 1 and k + 1 - j < 8:
            k += 1
        groups.append((L + j, high[j], k - j + 1))
        j = k + 1
    by_size = sorted(groups, key=lambda g: (-g[2], g[0]))
    dims = sorted(by_size[:3])
    split = sorted(by_size[3:])
    plan_perm = list(range(T))
    pos = low
    for tp, _, bits in dims + split:
        for w in range(bits):
            plan_perm[tp + w] = pos
            pos += 1
    plan = TmaPlan(low, dims, split, plan_perm)
    if plan.n_ops > MAX_TMA_OPS or any(g + b > 36 for _, g, b in dims):
        return None
    return plan


class Layout:
    def __init__(self, L: int, high: Sequence[int], mode: Optional[str] = None):
        self.L = L
        mode = mode or os.environ.get("QLS_JIT_LANDING", "tma")
        self.plan = tma_plan(L, high) if mode == "tma" else None
        self.mode = "tma" if self.plan is not None else "bulk"

    # -- exchange image ----------------------------------------------------------------------
    def ex_cls(self, p: int) -> int:
        if p < 3:
            return p
        if self.mode == "tma":
            return self.plan.perm[p] % 3
        return p % 3 if p < self.L else (p - self.L) % 3

    def swz(self, i: int) -> int:
        x = 0
        for p in range(3, T):
            if (i >> p) & 1:
                x ^= 1 << self.ex_cls(p)
        return i ^ x

    def swz_expr(self, var: str) -> str:
        fields: List[Tuple[int, int, int]] = []  # (start, width, class of start)
        p = 3
        while p < T:
            c = self.ex_cls(p)
            w = 1
            while p + w < T and c + w < 3 and self.ex_cls(p + w) == c + w:
                w += 1
            fields.append((p, w, c))
            p += w
        terms = []
        for start, w, c in fields:
            t = f"(({var} >> {start}) & {(1 << w) - 1}u)"
            if c:
                t = f"({t} << {c})"
            terms.append(t)
        return f"({var} ^ " + " ^ ".join(terms) + ")"

    # -- landing image -----------------------------------------------------------------------
    def land(self, i: int) -> int:
        if self.mode == "tma":
            lam = sum(((i >> p) & 1) << self.plan.perm[p] for p in range(T))
            return lam ^ ((lam >> 3) & 7)
        if self.L >= T:
            return i
        r = i >> self.L
        return i + r + (r >> 3) + (r >> 6)

    def land_units(self) -> int:
        """Units (16 bytes) of a tile buffer."""
        if self.mode == "tma" or self.L >= T:
            return 1 << T
        r = (1 << (T - self.L)) - 1
        return (1 << T) + r + (r >> 3) + (r >> 6) + 8

    def land_expr_from(self, var: str, pairs: Sequence[Tuple[int, int]]) -> str:
        """Landing unit of the tile-local index whose bit ``dst`` is bit ``src`` of ``var`` (for every pair)."""
        if self.mode == "tma":
            return "qls_swz128(" + deposit_expr(var, [(src, self.plan.perm[dst]) for src, dst in pairs]) + ")"
        return "qls_pad_land(" + deposit_expr(var, pairs) + ")"

    @property
    def additive(self) -> bool:
        """bulk: land(a | b) = land(a) + land(b) for disjoint a, b; tma: land(a | b) = land(a) ^ land(b)."""
        return self.mode == "bulk"


def choose_thread_order(regq: Sequence[int], lay: Layout) -> List[int]:
    """Tile position of every thread bit: the triple for thread bits 0..2 is the one whose eight lanes hit
    eight different bank groups in the exchange image and, if possible, in the landing image as well."""
    from itertools import combinations

    free = [p for p in range(T) if p not in regq]
    best, best_score = None, -1
    for tri in combinations(free, 3):
        lanes = [sum(((t >> j) & 1) << tri[j] for j in range(3)) for t in range(8)]
        ex = len({lay.swz(i) & 7 for i in lanes})
        ld = len({lay.land(i) & 7 for i in lanes})
        score = ex * 16 + ld
        if score > best_score:
            best, best_score = tri, score
    pick = list(best)
    return pick + [p for p in free if p not in pick]


# --------------------------------------------------------------------------------------
# expression helpers
# --------------------------------------------------------------------------------------


def _segments(pairs: Sequence[Tuple[int, int]]) -> List[Tuple[int, int, int]]:
    segs: List[List[int]] = []
    for src, dst in sorted(pairs):
        if segs and segs[-1][0] + segs[-1][1] == src and segs[-1][2] + segs[-1][1] == dst:
            segs[-1][1] += 1
        else:
            segs.append([src, 1, dst])
    return [tuple(s) for s in segs]


def deposit_expr(var: str, pairs: Sequence[Tuple[int, int]], wide: bool = False) -> str:
    """C expression moving bit ``src`` of ``var`` to bit ``dst`` for every pair."""
    if not pairs:
        return "0ull" if wide else "0u"
    sfx = "ull" if wide else "u"
    terms = []
    for src, w, dst in _segments(pairs):
        mask = ((1 << w) - 1) << src
        t = f"({var} & {mask:#x}{sfx})"
        if dst > src:
            t = f"({t} << {dst - src})"
        elif dst < src:
            t = f"({t} >> {src - dst})"
        terms.append(t)
    return "(" + " | ".join(terms) + ")"


def _hexf(x: float) -> str:
    return float(x).hex()


# --------------------------------------------------------------------------------------
# code generation
# --------------------------------------------------------------------------------------


class _TooBig(Exception):
    pass


class _Gen:
    def __init__(self, jp: JPass, mode: Optional[str] = None):
        self.jp = jp
        self.L = jp.L
        self.cm: List[complex] = []  # gate-matrix constants
        self.lines: List[str] = []
        self.preds: Dict[Tuple[int, int], str] = {}  # external-control predicate -> variable
        self.lay = Layout(jp.L, jp.high, mode)
        # QLS_JIT_SHAPES=1 (default 0): the 4x4 gates become steps of a small tile program whose switch cases are the
        # distinct gate SHAPES of the pass, so that their code exists once.  Motivation: the straight-line form of a heavy
        # pass is ~80 KB of code, more than twice the instruction cache of an SM, streamed by both consumer groups at
        # different places (measured: GPC-level instruction cache at 78-89 % of its request rate, FP64 pipe 62-65 % active).
        # Measured at n = 33: the program form is SLOWER (1044 vs 944 ms per solve) -- a dispatch per step and no
        # scheduling across gates cost more than the instruction fetches saved -- so it stays an experiment.
        est = 400
        for st in jp.stages:
            est += 120
            for op in st.ops:
                est += {"g2": 140 if op.real else 280, "g1": 40 if op.real else 72}.get(op.kind, 110)
        self.share = os.environ.get("QLS_JIT_SHAPES", "0") != "0" and est > int(os.environ.get("QLS_JIT_SHARE_MIN", "2300"))
        self.shapes: Dict[tuple, Tuple[int, List[str]]] = {}
        self.steps: List[tuple] = []
        self.chunk: List[str] = []
        self.orders = [choose_thread_order(s.regq, self.lay) for s in jp.stages]

    # -- per-stage address pieces ----------------------------------------------------------
    def _tb(self, s: int) -> str:
        order = self.orders[s]
        return deposit_expr("tid", [(j, order[j]) for j in range(T - KR)])

    def _off(self, s: int, r: int) -> int:
        regq = self.jp.stages[s].regq
        return sum(((r >> j) & 1) << regq[j] for j in range(KR))

    def emit(self, line: str = "", ind: int = 1) -> None:
        self.lines.append("  " * ind + line)

    def _flush_chunk(self) -> None:
        if self.share and self.lines:
            self.steps.append(("chunk", list(self.lines)))
            del self.lines[:]

    def _landing(self, s: int, store: bool, ind: int) -> None:
        order = self.orders[s]
        self.emit("{", ind)
        self.emit(f"const u32 lb = {self.lay.land_expr_from('tid', [(j, order[j]) for j in range(T - KR)])};", ind + 1)
        for r in range(2**KR):
            c = self.lay.land(self._off(s, r))
            ref = f"sm[lb + {c}u]" if self.lay.additive else f"sm[lb ^ {c}u]"
            self.emit(f"{ref} = a{r};" if store else f"a{r} = {ref};", ind + 1)
        self.emit("}", ind)

    def _exchange(self, s: int, store: bool, ind: int) -> None:
        self.emit("{", ind)
        self.emit(f"const u32 tb = {self._tb(s)};", ind + 1)
        self.emit(f"const u32 sb = {self.lay.swz_expr('tb')};", ind + 1)
        for r in range(2**KR):
            c = self.lay.swz(self._off(s, r))
            self.emit(f"sm[sb ^ {c}u] = a{r};" if store else f"a{r} = sm[sb ^ {c}u];", ind + 1)
        self.emit("}", ind)

    # -- controls --------------------------------------------------------------------------
    def _pred(self, op: JOp) -> Optional[str]:
        if not op.xc_mask:
            return None
        key = (op.xc_mask, op.xc_val)
        if key not in self.preds:
            self.preds[key] = f"f{len(self.preds)}"
        return self.preds[key]

    def _split_ctl(self, s: int, op: JOp):
        """(thread mask, thread value, register mask, register value) of the op's controls inside the tile."""
        regq, order = self.jp.stages[s].regq, self.orders[s]
        tm = tv = rm = rv = 0
        for pos, val in op.tile_ctl.items():
            if pos in regq:
                j = regq.index(pos)
                rm |= 1 << j
                rv |= val << j
            else:
                j = order.index(pos)
                tm |= 1 << j
                tv |= val << j
        return tm, tv, rm, rv

    # -- ops -------------------------------------------------------------------------------
    def _table_index(self, s: int, op: JOp, ind: int):
        """Emit the per-thread table pointer; returns the per-register-combination index offsets."""
        regq, order = self.jp.stages[s].regq, self.orders[s]
        tpairs = [(order.index(pos), b) for pos, b in op.idx_t if pos not in regq]
        rbits = [(regq.index(pos), b) for pos, b in op.idx_t if pos in regq]
        parts = []
        if op.idx_x:
            parts.append("(u32)" + deposit_expr("gbase", op.idx_x, wide=True))
        if tpairs:
            parts.append(deposit_expr("tid", tpairs))
        idx = " | ".join(parts) if parts else "0u"
        self.emit(f"const V* const tab = reinterpret_cast<const V*>(pool + {op.pool_off}ull) + ({idx});", ind)
        return [sum(((r >> j) & 1) << b for j, b in rbits) for r in range(2**KR)]

    def _g_lines(self, k: int, real: bool, regs: Sequence[int], rm: int, rv: int, mexpr: str) -> List[str]:
        fn = f"qls_g{k}{'r' if real else 'c'}"
        rest = [j for j in range(KR) if j not in regs]
        out = []
        for g in range(2 ** (KR - k)):
            i0 = sum(((g >> n) & 1) << rest[n] for n in range(len(rest)))
            if (i0 & rm) != rv:
                continue
            args = [f"a{i0 | sum(((c >> j) & 1) << regs[j] for j in range(k))}" for c in range(2**k)]
            out.append(f"{fn}({', '.join(args)}, {mexpr});")
        return out

    def _emit_op(self, s: int, op: JOp, ind: int) -> None:
        regq = self.jp.stages[s].regq
        tm, tv, rm, rv = self._split_ctl(s, op)
        pred = self._pred(op)
        if op.kind == "g2" and self.share:
            # a step of the tile program: the code of a (targets, real/complex, register controls) shape exists once
            k = len(op.targets)
            moff = len(self.cm)
            self.cm += [complex(v) for v in op.matrix.reshape(-1)]
            regs = tuple(regq.index(t) for t in op.targets)
            key = (k, op.real, regs, rm, rv)
            if key not in self.shapes:
                self.shapes[key] = (len(self.shapes), self._g_lines(k, op.real, regs, rm, rv, "M"))
            self._flush_chunk()
            pi = 255 if pred is None else int(pred[1:])
            self.steps.append(("g", self.shapes[key][0], moff, pi, tm, tv))
            return
        conds = ([pred] if pred else []) + ([f"(tid & {tm:#x}u) == {tv:#x}u"] if tm else [])
        if conds:
            self.emit("if (" + " && ".join(conds) + ") {", ind)
        else:
            self.emit("{", ind)
        ind += 1
        if op.kind in ("g1", "g2"):
            k = len(op.targets)
            moff = len(self.cm)
            self.cm += [complex(v) for v in op.matrix.reshape(-1)]
            regs = [regq.index(t) for t in op.targets]
            for line in self._g_lines(k, op.real, regs, rm, rv, f"qls_cm + {moff}"):
                self.emit(line, ind)
        elif op.kind == "diag":
            rt = self._table_index(s, op, ind)
            todo = [r for r in range(2**KR) if (r & rm) == rv]
            by_idx: Dict[int, List[int]] = {}
            for r in todo:
                by_idx.setdefault(rt[r], []).append(r)
            keys = list(by_idx)
            nb_ = int(os.environ.get("QLS_JIT_DIAG_BATCH", "4"))
            for c in range(0, len(keys), nb_):  # four phases in flight (eight measured: see profiles/README.md)
                chunk = keys[c:c + nb_]
                for n, key in enumerate(chunk):
                    self.emit(f"const V p{c + n} = __ldg(tab + {key});", ind)
                for n, key in enumerate(chunk):
                    for r in by_idx[key]:
                        self.emit(f"a{r} = qls_cmul(a{r}, p{c + n});", ind)
        elif op.kind == "mux":
            rt = self._table_index(s, op, ind)
            q = regq.index(op.targets[0])
            pairs = []
            for g in range(2 ** (KR - 1)):
                i0 = ((g >> q) << (q + 1)) | (g & ((1 << q) - 1))
                if (i0 & rm) == rv:
                    pairs.append(i0)
            nb_ = int(os.environ.get("QLS_JIT_DIAG_BATCH", "4"))
            for c in range(0, len(pairs), nb_):
                chunk = pairs[c:c + nb_]
                for n, i0 in enumerate(chunk):
                    self.emit(f"const V p{c + n} = __ldg(tab + {rt[i0]});", ind)
                for n, i0 in enumerate(chunk):
                    self.emit(f"qls_ry(a{i0}, a{i0 | (1 << q)}, p{c + n});", ind)
        else:
            raise ValueError(op.kind)
        ind -= 1
        self.emit("}", ind)

    def tile_order(self) -> List[Tuple[int, int]]:
        """(bit of the tile ordinal, state-index position) pairs of ``qls_tile_base``.

        Default: the free index positions in ascending order.  QLS_JIT_TILE_ORDER=pattern gives the control qubits OUTSIDE
        the tile the top bits of the ordinal, so that a CTA works through all tiles of one control pattern (= one subset of
        the pass's code) before the next pattern starts.  Measured at n = 33 (profiles/r02_tile_order_ab_n33.txt): no
        difference (935 vs 940 ms per solve) -- the control qubits are high clock qubits, so the pattern of a CTA's
        consecutive tiles already changes only every ~50 tiles in index order."""
        pos = [dst + w for _, wd, dst in sorted(self.jp.fseg) for w in range(wd)]
        if os.environ.get("QLS_JIT_TILE_ORDER", "index") == "pattern":
            ctl = 0
            for st in self.jp.stages:
                for op in st.ops:
                    ctl |= op.xc_mask
            pos = [p for p in pos if not (ctl >> p) & 1] + [p for p in pos if (ctl >> p) & 1]
        return list(enumerate(pos))

    # -- whole unit ------------------------------------------------------------------------
    def source(self) -> str:
        jp = self.jp
        ns = len(jp.stages)
        always = [any(not op.xc_mask for op in st.ops) for st in jp.stages]
        body: List[str] = []
        self.lines = body
        # Every stage block: (enter) landing read if it is the first stage that fires on this tile, else read the
        # exchange image; (ops); (leave) if a later stage fires on this tile, write the exchange image, else write the
        # landing image.  Each piece of movement code exists once per stage; which one runs is decided by uniform
        # predicates on the tile's control qubits.
        fires = [None if always[s] else " || ".join(sorted({self._pred(op) for op in st.ops})) for s, st in enumerate(jp.stages)]
        for s, st in enumerate(jp.stages):
            fire = fires[s]
            ind = 1
            self.emit(f"// ---- stage {s}: register qubits at tile positions {list(st.regq)}, thread bits {self.orders[s]}")
            if fire:
                self.emit(f"if ({fire}) {{")
                ind = 2
            may_be_first = not any(always[:s])
            if s == 0:
                self._landing(0, False, ind)
                self.emit("landed = true;", ind)
            elif may_be_first:
                self.emit("if (!loaded) {", ind)
                self._landing(s, False, ind + 1)
                self.emit("landed = true;", ind + 1)
                self.emit("} else {", ind)
                self._exchange(s, False, ind + 1)
                self.emit("landed = false;", ind + 1)
                self.emit("}", ind)
            else:
                self._exchange(s, False, ind)
                self.emit("landed = false;", ind)
            self.emit("loaded = true;", ind)
            self.emit(f"QLS_TR({10 + 2 * s});", ind)
            if self.share and fire:  # the ops are steps of their own (each with its predicate): close the block here ...
                self.emit("}")
            for op in st.ops:
                self._emit_op(s, op, 1 if self.share else ind)
            if self.share and fire:  # ... and reopen it for the code that leaves the stage
                self.emit(f"if ({fire}) {{")
            self.emit(f"QLS_TR({11 + 2 * s});", ind)
            self.emit(f"tid = QLS_FRESH_TID({1 + s % 20});  // recompute the addresses below instead of keeping sixteen of them across the ops", ind)
            # leave
            if s == ns - 1:
                later = "false"
            elif any(always[s + 1:]):
                later = "true"
            else:
                later = " || ".join(sorted({f for f in fires[s + 1:] if f}))
                later = " || ".join(sorted(set(later.split(" || "))))
            if later != "false":
                if later != "true":
                    self.emit(f"if ({later}) {{", ind)
                    ind += 1
                self.emit("if (landed) QLS_BODY_BAR();  // the landing image has been read by everyone", ind)
                self._exchange(s, True, ind)
                self.emit("QLS_BODY_BAR();", ind)
                if later != "true":
                    ind -= 1
                    self.emit("} else {", ind)
            if later != "true":
                i2 = ind + (1 if later != "false" else 0)
                self.emit("if (!landed) QLS_BODY_BAR();  // the exchange image has been read by everyone", i2)
                self._landing(s, True, i2)
                if later != "false":
                    self.emit("}", ind)
            if fire:
                self.emit("}")

        out: List[str] = []
        out.append(f"// generated by qls_b200/jit.py (version {GEN_VERSION}); do not edit")
        lay = self.lay
        out.append(f"#define QLS_L {jp.L}")
        out.append(f"#define QLS_NRUNS {1 << (T - jp.L)}u")
        out.append(f"#define QLS_TILE_BITS {T}")
        out.append(f"#define QLS_BUF_UNITS {lay.land_units()}u")
        if lay.mode == "tma":
            out.append("#define QLS_TMA 1")
            out.append(f"#define QLS_TMA_OPS {lay.plan.n_ops}u")
        out.append('#include "rs2_runtime.cuh"')
        out.append("// landing unit (16 bytes) of tile-local amplitude i")
        out.append("QLS_DEV u32 qls_land_unit(const u32 i) {")
        out.append(f"  return {lay.land_expr_from('i', [(p, p) for p in range(T)])};")
        out.append("}")
        if lay.mode == "tma":
            # copy k of a tile: amplitude offset of its sub-box (the split groups' bits of the global index)
            pairs, bit = [], 0
            for _, gpos, bits in lay.plan.split:
                pairs += [(bit + w, gpos + w) for w in range(bits)]
                bit += bits
            out.append("QLS_DEV u64 qls_tma_split_offset(const u32 k) {")
            out.append(f"  return {deposit_expr('(u64)k', pairs, wide=True)};")
            out.append("}")
        cm = self.cm or [0j]
        out.append(f"QLS_CONST V qls_cm[{len(cm)}] = {{")
        for j in range(0, len(cm), 2):
            out.append("  " + " ".join(f"{{{_hexf(v.real)}, {_hexf(v.imag)}}}," for v in cm[j:j + 2]))
        out.append("};")
        out.append("QLS_DEV u64 qls_tile_base(const u64 t) {")
        out.append(f"  return {deposit_expr('t', self.tile_order(), wide=True)} | {jp.fix_val:#x}ull;")
        out.append("}")
        out.append("QLS_DEV u64 qls_run_offset(const u32 r) {")
        out.append(f"  return {deposit_expr('(u64)r', [(j, q) for j, q in enumerate(jp.high)], wide=True)};")
        out.append("}")
        pred_lines = [f"  const bool {name} = (gbase & {m:#x}ull) == {v:#x}ull;" for (m, v), name in self.preds.items()]
        out.append("QLS_DEV bool qls_tile_idle(const u64 gbase) {")
        if any(always) or not self.preds:
            out.append("  (void)gbase;")
            out.append("  return false;")
        else:
            out += pred_lines
            out.append("  return !(" + " || ".join(self.preds.values()) + ");")
        out.append("}")
        out.append("QLS_DEV void qls_tile_body(V* __restrict__ sm, const u64 gbase, const u32 tid_in, const u32 qls_group,")
        out.append("                           const unsigned char* __restrict__ pool, u64* qls_trace, const u32 qls_q,")
        out.append("                           volatile u32* qls_cursor) {")
        out.append("  (void)qls_group; (void)pool; (void)qls_trace; (void)qls_q; (void)qls_cursor;")
        out.append("  // opaque copy of the thread index: keeps the per-stage addresses from being hoisted out of the tile loop")
        out.append("  // (a dozen values per thread that would live -- spilled -- across the whole pass)")
        out.append("  u32 tid = QLS_FRESH_TID(0);")
        out.append("  V " + ", ".join(f"a{r}" for r in range(2**KR)) + ";")
        out.append("  bool loaded = false;  // a stage has fired on this tile: the amplitudes are in registers / the exchange image")
        out.append("  bool landed = false;  // the registers were filled from the landing image (no exchange since)")
        out += pred_lines
        if not self.share:
            out += body
            out.append("}")
        else:
            self._flush_chunk()
            npred = len(self.preds)
            out.append("  const u32 fmask = 0u" + "".join(f" | ({name} ? {1 << j}u : 0u)" for j, name in enumerate(self.preds.values())) + ";")
            out.append("  (void)fmask;")
            out.append("  // the tile program: word 0 = opcode | predicate << 8 | matrix offset << 16, word 1 = thread-control mask | value << 16")
            out.append("  // (the shuffles make the step words provably warp-uniform: ptxas then keeps the matrix elements in uniform")
            out.append("  // registers; the words of step k + 1 are fetched while step k runs)")
            out.append("  u32 n0 = qls_prog[0], n1 = qls_prog[1];")
            out.append("  for (u32 pc = 1;; ++pc) {")
            out.append("    const u32 w0 = QLS_UNIFORM(n0), w1 = QLS_UNIFORM(n1);")
            out.append("    n0 = qls_prog[2 * pc];")
            out.append("    n1 = qls_prog[2 * pc + 1];")
            out.append("    const u32 pi = (w0 >> 8) & 0xffu;")
            out.append("    if (pi != 255u && !((fmask >> pi) & 1u)) continue;")
            out.append("    switch (w0 & 0xffu) {")
            out.append("      case 0: return;")
            for key, (j, lines) in self.shapes.items():
                out.append(f"      case {1 + j}: {{  // dense {2**key[0]}x{2**key[0]} ({'real' if key[1] else 'complex'}) on register qubits {list(key[2])}, register controls {key[3]:#x}/{key[4]:#x}")
                out.append("        const V* const M = qls_cm + (w0 >> 16);")
                out.append("        if ((tid & (w1 & 0xffffu)) == (w1 >> 16)) {")
                out += ["          " + ln for ln in lines]
                out.append("        }")
                out.append("      } break;")
            prog: List[int] = []
            nchunk = 0
            for st in self.steps:
                if st[0] == "chunk":
                    out.append(f"      case {32 + nchunk}: {{")
                    out += ["      " + ln for ln in st[1]]
                    out.append("      } break;")
                    prog += [32 + nchunk | (255 << 8), 0]
                    nchunk += 1
                else:
                    _, shape, moff, pi, tm, tv = st
                    prog += [(1 + shape) | (pi << 8) | (moff << 16), tm | (tv << 16)]
            prog += [0 | (255 << 8), 0, 0 | (255 << 8), 0]  # (the fetch runs one step ahead)
            out.append("      default: break;")
            out.append("    }")
            out.append("  }")
            out.append("}")
            if len(self.shapes) > 31 or nchunk > 200 or len(self.cm) >= 65536 or npred > 31:
                raise _TooBig()
            # the program table must precede the function: insert it after the matrix constants
            table = [f"QLS_CONST u32 qls_prog[{len(prog)}] = {{"]
            for j in range(0, len(prog), 8):
                table.append("  " + " ".join(f"{v:#x}u," for v in prog[j:j + 8]))
            table.append("};")
            at = next(i for i, ln in enumerate(out) if ln.startswith("QLS_DEV u64 qls_tile_base"))
            out[at:at] = table
        out.append('#include "rs2_kernel.cuh"')
        return "\n".join(out) + "\n"


def generate_source(jp: JPass, mode: Optional[str] = None) -> str:
    try:
        return _Gen(jp, mode).source()
    except _TooBig:  # more shapes / steps than the tile program can name: straight-line form
        g = _Gen(jp, mode)
        g.share = False
        return g.source()


def smem_bytes(jp: JPass, mode: Optional[str] = None) -> int:
    buf = (Layout(jp.L, jp.high, mode).land_units() * 16 + 1023) & ~1023
    return 3 * buf + 2 * 3 * 8 + 64 + 1024  # + slack to align the buffers to 1024 bytes (tensor-map swizzle)


def tma_descriptor(jp: JPass, mode: Optional[str] = None):
    """(low_bits, [(global index bit, bits)] of dims 2..4) of the pass's tensor map, or None (bulk landing)."""
    lay = Layout(jp.L, jp.high, mode)
    if lay.mode != "tma":
        return None
    return lay.plan.low_bits, [(g, b) for _, g, b in lay.plan.dims]


# --------------------------------------------------------------------------------------
# compilation + cache
# --------------------------------------------------------------------------------------

_headers_cache: Optional[Dict[str, str]] = None


def _headers() -> Dict[str, str]:
    global _headers_cache
    if _headers_cache is None:
        _headers_cache = {name: open(os.path.join(CSRC, name)).read() for name in ("rs2_runtime.cuh", "rs2_kernel.cuh")}
    return _headers_cache


NVRTC_OPTIONS = ["--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-default-device"]
if os.environ.get("QLS_JIT_EXPERIMENT", ""):  # timing experiments (wrong results): -DQLS_EXP_<NAME>
    NVRTC_OPTIONS += ["-DQLS_EXP_" + x.upper() for x in os.environ["QLS_JIT_EXPERIMENT"].split(",")]
if os.environ.get("QLS_JIT_TRACE", "0") != "0":  # timeline of CTA 0 (tools/jit_trace.py); never set in production
    NVRTC_OPTIONS.append("-DQLS_TRACE")


def cache_key(source: str) -> str:
    h = hashlib.sha256()
    h.update(f"v{GEN_VERSION}".encode())
    for name, text in sorted(_headers().items()):
        h.update(name.encode())
        h.update(text.encode())
    h.update(" ".join(NVRTC_OPTIONS).encode())
    h.update(source.encode())
    return h.hexdigest()[:32]


def compile_source(source: str, use_cache: bool = True) -> bytes:
    """CUDA source -> sm_100a cubin (NVRTC inside libqls_b200.so; no GPU needed)."""
    key = cache_key(source)
    path = os.path.join(CACHE_DIR, key + ".cubin")
    if use_cache and os.path.exists(path):
        with open(path, "rb") as f:
            return f.read()
    lib = _abi.load()
    hdr = _headers()
    names = (C.c_char_p * len(hdr))(*[n.encode() for n in hdr])
    texts = (C.c_char_p * len(hdr))(*[t.encode() for t in hdr.values()])
    opts = (C.c_char_p * len(NVRTC_OPTIONS))(*[o.encode() for o in NVRTC_OPTIONS])
    out = C.c_void_p()
    size = C.c_uint64()
    log = C.create_string_buffer(1 << 16)
    src_path = os.path.join(CACHE_DIR, key + ".cu")  # kept beside the cubin: profilers map the line table to it
    if use_cache:
        os.makedirs(CACHE_DIR, exist_ok=True)
        with open(src_path, "w") as f:
            f.write(source)
    rc = lib.qls_jit_compile(source.encode(), src_path.encode(), names, texts, len(hdr), opts, len(NVRTC_OPTIONS), C.byref(out), C.byref(size),
                             log, len(log))
    if rc != 0:
        raise _abi.QlsError(f"NVRTC failed ({rc}): {lib.qls_last_error().decode()}\n{log.value.decode(errors='replace')[-4000:]}")
    cubin = C.string_at(out, size.value)
    lib.qls_jit_free(out)
    if use_cache:
        os.makedirs(CACHE_DIR, exist_ok=True)
        tmp = path + f".{os.getpid()}.tmp"
        with open(tmp, "wb") as f:
            f.write(cubin)
        os.replace(tmp, path)
    return cubin


def enabled() -> bool:
    return os.environ.get("QLS_JIT", "1") != "0"


def min_qubits() -> int:
    """States smaller than this keep the interpreter kernels (compiling is not worth it)."""
    return int(os.environ.get("QLS_JIT_MIN_QUBITS", "20"))


@dataclass
class CompiledPass:
    index: int
    cubin: bytes
    smem: int
    tma: Optional[Tuple[int, List[Tuple[int, int]]]]  # tensor-map description (tma_descriptor) or None: bulk landing
    ir: JPass


def compile_pass(jp: JPass, index: int, mode: Optional[str] = None) -> CompiledPass:
    return CompiledPass(index, compile_source(generate_source(jp, mode)), smem_bytes(jp, mode), tma_descriptor(jp, mode), jp)


def compile_program(lowered, threads: int = 0) -> List[CompiledPass]:
    """One CompiledPass for every pass of ``lowered`` that has a specialised form (NVRTC runs on a thread pool)."""
    from concurrent.futures import ThreadPoolExecutor

    jobs = [(i, jp) for i, jp in ((i, extract_pass(lowered, i)) for i in range(lowered.n_passes)) if jp is not None]
    if not jobs:
        return []
    workers = threads or min(len(jobs), os.cpu_count() or 1)
    with ThreadPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(lambda j: compile_pass(j[1], j[0]), jobs))


def attach(lib, handle, cp: CompiledPass, n_local: int) -> str:
    """Attach a compiled pass to an uploaded program; returns the landing mode in use.  If the driver rejects
    the tensor map, the pass is recompiled with per-run bulk copies (announced on stderr: slower, not wrong)."""
    import sys

    _abi.check(lib.qls_program_set_pass_kernel(handle, cp.index, cp.cubin, len(cp.cubin), KERNEL_NAME.encode(), THREADS, cp.smem))
    if cp.tma is None:
        return "bulk"
    low_bits, groups = cp.tma
    pos = (C.c_int * 3)(*([g for g, _ in groups] + [0] * (3 - len(groups))))
    bits = (C.c_int * 3)(*([b for _, b in groups] + [0] * (3 - len(groups))))
    rc = lib.qls_program_set_pass_tma(handle, cp.index, low_bits, len(groups), pos, bits, n_local)
    if rc == 0:
        return "tma"
    msg = lib.qls_last_error().decode()
    sys.stderr.write(f"qls_b200.jit: pass {cp.index}: tensor map rejected ({msg}); using per-run bulk copies\n")
    alt = compile_pass(cp.ir, cp.index, "bulk")
    _abi.check(lib.qls_program_set_pass_kernel(handle, alt.index, alt.cubin, len(alt.cubin), KERNEL_NAME.encode(), THREADS, alt.smem))
    return "bulk"
