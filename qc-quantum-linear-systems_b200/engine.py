"""Device execution of lowered programs.

PyTorch is used for exactly three things: allocating device buffers (the statevector, the GEMM
scratch, the reduction workspace), exposing the current CUDA stream, and (in ``shard.py``)
``torch.distributed``.  All arithmetic happens in libqls_b200.so through the C ABI.
"""
from __future__ import annotations

import os

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _abi
from .lower import LoweredProgram, Step


def _torch():
    import torch

    return torch


def require_cuda(device: Optional[int] = None) -> int:
    torch = _torch()
    if not torch.cuda.is_available():
        raise _abi.QlsError("no CUDA device: the B200 statevector engine has no CPU fallback")
    _abi.load()
    return torch.cuda.current_device() if device is None else int(device)


def _stream_ptr() -> int:
    return int(_torch().cuda.current_stream().cuda_stream)


class DeviceState:
    """A statevector shard of ``2**n_local`` amplitudes in HBM (interleaved re/im)."""

    def __init__(self, n_local: int, dtype: str = "fp64", device: Optional[int] = None, index_offset: int = 0):
        torch = _torch()
        self.device = require_cuda(device)
        self.lib = _abi.load()
        self.n_local = int(n_local)
        self.dtype = dtype
        self.abi_dtype = _abi.QLS_C128 if dtype == "fp64" else _abi.QLS_C64
        self.np_dtype = np.complex128 if dtype == "fp64" else np.complex64
        self.index_offset = int(index_offset)
        tdt = torch.complex128 if dtype == "fp64" else torch.complex64
        with torch.cuda.device(self.device):
            self.tensor = torch.empty(2**self.n_local, dtype=tdt, device=f"cuda:{self.device}")
            self.workspace = torch.empty(_abi.QLS_REDUCE_WORKSPACE_BYTES, dtype=torch.uint8, device=f"cuda:{self.device}")
        self._scratch = None

    @property
    def ptr(self) -> int:
        return self.tensor.data_ptr()

    @property
    def dim(self) -> int:
        return 2**self.n_local

    def scratch(self, n_amps: int):
        torch = _torch()
        if self._scratch is None or self._scratch.numel() < n_amps:
            with torch.cuda.device(self.device):
                self._scratch = torch.empty(n_amps, dtype=self.tensor.dtype, device=self.tensor.device)
        return self._scratch

    # -- initialisation / access --------------------------------------------------------
    def init_basis(self, index: Optional[int] = 0) -> None:
        with _torch().cuda.device(self.device):
            _abi.check(self.lib.qls_sv_init_basis(self.ptr, self.n_local, self.abi_dtype,
                                                  _abi.UINT64_MAX if index is None else int(index), _stream_ptr()))

    def write(self, offset: int, amps: np.ndarray) -> None:
        a = np.ascontiguousarray(amps, dtype=self.np_dtype)
        if offset < 0 or offset + a.size > self.dim:
            raise ValueError("write outside the state")
        with _torch().cuda.device(self.device):
            _abi.check(self.lib.qls_sv_write(self.ptr, self.abi_dtype, int(offset), a.ctypes.data, a.size, _stream_ptr()))

    def read(self, offset: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.dim - offset if count is None else int(count)
        if offset < 0 or count < 0 or offset + count > self.dim:
            raise ValueError("read outside the state")
        out = np.empty(count, dtype=self.np_dtype)
        with _torch().cuda.device(self.device):
            _abi.check(self.lib.qls_sv_read(self.ptr, self.abi_dtype, int(offset), out.ctypes.data, count, _stream_ptr()))
        return out

    def scale(self, factor: complex) -> None:
        with _torch().cuda.device(self.device):
            _abi.check(self.lib.qls_sv_scale(self.ptr, self.n_local, self.abi_dtype, float(np.real(factor)),
                                             float(np.imag(factor)), _stream_ptr()))

    # -- reductions ---------------------------------------------------------------------
    def _reduce(self, kind: int, a: int, b: int, other: Optional["DeviceState"] = None) -> complex:
        out = (C.c_double * 2)()
        with _torch().cuda.device(self.device):
            _abi.check(self.lib.qls_reduce(self.ptr, self.n_local, self.abi_dtype, kind, int(a), int(b), self.index_offset,
                                           other.ptr if other is not None else None, self.workspace.data_ptr(), out,
                                           _stream_ptr()))
        return complex(out[0], out[1])

    def norm_sq_range(self, offset: int, count: int) -> float:
        return self._reduce(_abi.QLS_RED_NORM_RANGE, offset, count).real

    def prob_mask(self, mask: int, value: int) -> float:
        return self._reduce(_abi.QLS_RED_PROB_MASK, mask, value).real

    def expectation_z(self, qubit: int) -> float:
        return self._reduce(_abi.QLS_RED_EXPECT_Z, qubit, 0).real

    def inner(self, other: "DeviceState") -> complex:
        """<other|self>"""
        return self._reduce(_abi.QLS_RED_INNER, 0, 0, other)


class DeviceProgram:
    """A lowered program uploaded to one GPU."""

    def __init__(self, lowered: LoweredProgram, device: Optional[int] = None):
        torch = _torch()
        self.device = require_cuda(device)
        self.lib = _abi.load()
        self.lowered = lowered
        self.abi_dtype = _abi.QLS_C128 if lowered.dtype == "fp64" else _abi.QLS_C64
        handle = C.c_void_p()
        pool = lowered.pool
        with torch.cuda.device(self.device):
            _abi.check(self.lib.qls_program_create(lowered.passes, lowered.n_passes, lowered.ops,
                                                   sum(p.n_ops for p in lowered.pass_info), lowered.rs_ops,
                                                   lowered.n_rs_ops, pool.ctypes.data, pool.size,
                                                   self.abi_dtype, self.device, C.byref(handle)))
        self.handle = handle
        # Specialised kernels: every pass with a register-stage form gets a kernel compiled for exactly that pass
        # (qls_b200/jit.py; NVRTC through the C ABI, cubins cached on disk).  A failure is an error, not a silent
        # fall-back to the interpreter kernels; QLS_JIT=0 selects those on purpose.
        self.jit_passes = 0
        self.jit_modes = {}
        self.jit_attached = set()  # pass indices that have a specialised kernel
        from . import jit

        if jit.enabled() and lowered.dtype == "fp64" and lowered.n_local >= jit.min_qubits():
            with torch.cuda.device(self.device):
                for cp in jit.compile_program(lowered):
                    mode = jit.attach(self.lib, self.handle, cp, lowered.n_local)
                    self.jit_passes += 1
                    self.jit_attached.add(cp.index)
                    self.jit_modes[mode] = self.jit_modes.get(mode, 0) + 1
        self._gemm_mats = {}
        cdt = torch.complex128 if lowered.dtype == "fp64" else torch.complex64
        for i, s in enumerate(lowered.steps):
            if s.kind == "gemm":
                with torch.cuda.device(self.device):
                    self._gemm_mats[i] = torch.from_numpy(np.ascontiguousarray(s.matrix)).to(device=f"cuda:{self.device}", dtype=cdt)
        # kernels launched per run (for bench.py's gpu_launches claim)
        self.launches_per_run = sum(
            (s.pass_end - s.pass_begin) if s.kind == "passes" else (2 if s.kind == "gemm" else 2) for s in lowered.steps
        )

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value:
                self.lib.qls_program_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def run(self, state: DeviceState, initialise: bool = True, first_pass: int = 0, last_pass: Optional[int] = None) -> None:
        """Evolve ``state`` (|0...0> unless ``initialise`` is False) through the whole program.  ``first_pass`` / ``last_pass``
        clip the tile passes to ``[first_pass, last_pass)`` (the sharded solver runs the passes next to an exchange itself,
        half a shard at a time: ``run_pass_part``)."""
        lw = self.lowered
        last_pass = lw.n_passes if last_pass is None else last_pass
        if state.n_local != lw.n_local:
            raise ValueError(f"state has {state.n_local} local qubits, program expects {lw.n_local}")
        steps = lw.steps
        torch = _torch()
        with torch.cuda.device(self.device):
            stream = _stream_ptr()
            first = 0
            if steps and steps[0].kind == "init":
                # initialise=False: the caller has already prepared the register (e.g. another b)
                if initialise:
                    self._run_init(state, steps[0])
                first = 1
            elif initialise:
                state.init_basis(0 if state.index_offset == 0 else None)
            elif lw.zero_start:
                raise _abi.QlsError("this program was lowered for a |0...0> input (zero-qubit tracking): run it with initialise=True "
                               "or lower the circuit without LoweringOptions.zero_start")
            for i in range(first, len(steps)):
                s = steps[i]
                if s.kind == "passes":
                    lo, hi = max(s.pass_begin, first_pass), min(s.pass_end, last_pass)
                    if lo < hi:
                        _abi.check(self.lib.qls_program_run(self.handle, state.ptr, state.n_local, state.index_offset, lo, hi, stream))
                elif s.kind == "gemm":
                    n_rows_sel = 2 ** (state.n_local - s.nb - bin(s.ctrl_mask & (2**state.n_local - 1)).count("1"))
                    # out of place through a bounded scratch (<= 2^14 rows) + copy-back; QLS_GEMM_INPLACE=1: no scratch at all
                    in_place = lw.dtype == "fp64" and s.nb <= 10 and os.environ.get("QLS_GEMM_INPLACE", "0") == "1"
                    scratch_ptr = None if in_place else state.scratch(min(n_rows_sel, _abi.QLS_GEMM_CHUNK_ROWS) * 2**s.nb).data_ptr()
                    _abi.check(self.lib.qls_ctrl_gemm(state.ptr, scratch_ptr, state.n_local, self.abi_dtype, s.nb,
                                                      s.ctrl_mask, s.ctrl_val, state.index_offset,
                                                      self._gemm_mats[i].data_ptr(), stream))
                elif s.kind == "init":
                    raise NotImplementedError("state preparation is only supported as the first operation")
                else:
                    raise ValueError(s.kind)

    def run_pass_part(self, state: DeviceState, pass_index: int, sel: Sequence[Tuple[int, int]]) -> None:
        """Pass ``pass_index`` on the part of the state whose index bits ``pos`` equal ``val`` for the one or two ``(pos, val)``
        of ``sel`` (current stream)."""
        (p1, v1), (p2, v2) = sel[0], (sel[1] if len(sel) > 1 else (-1, 0))
        with _torch().cuda.device(self.device):
            _abi.check(self.lib.qls_program_run_part(self.handle, state.ptr, state.n_local, state.index_offset, int(pass_index),
                                                     int(p1), int(v1), int(p2), int(v2), _stream_ptr()))

    def free_positions(self, pass_index: int) -> List[int]:
        """State-index positions a pass neither transforms nor fixes (it can run on one value of any of them)."""
        p = self.lowered.passes[pass_index]
        return [p.fseg[j].dst + w for j in range(p.n_fseg) for w in range(p.fseg[j].width)]

    def run_each_pass_timed(self, state: DeviceState):
        """Debug aid: per-pass milliseconds (one CUDA event pair per tile pass)."""
        torch = _torch()
        lw = self.lowered
        out = []
        with torch.cuda.device(self.device):
            stream = _stream_ptr()
            for s in lw.steps:
                if s.kind != "passes":
                    continue
                for j in range(s.pass_begin, s.pass_end):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    _abi.check(self.lib.qls_program_run(self.handle, state.ptr, state.n_local, state.index_offset, j, j + 1, stream))
                    b.record()
                    out.append((j, a, b))
            torch.cuda.synchronize()
        return [(j, a.elapsed_time(b)) for j, a, b in out]

    def run_passes_timed(self, state: DeviceState):
        """Run every step after the state preparation with CUDA events (on the launching stream)
        around each contiguous range of tile passes.  Returns (tile-pass ms, tile-pass launches,
        algorithmic bytes of those launches); other steps run untimed."""
        torch = _torch()
        lw = self.lowered
        pairs = []
        launches = 0
        nbytes = 0
        with torch.cuda.device(self.device):
            stream = _stream_ptr()
            for i, s in enumerate(lw.steps):
                if s.kind == "init":
                    continue
                if s.kind == "passes":
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    _abi.check(self.lib.qls_program_run(self.handle, state.ptr, state.n_local, state.index_offset,
                                                        s.pass_begin, s.pass_end, stream))
                    b.record()
                    pairs.append((a, b))
                    launches += s.pass_end - s.pass_begin
                    nbytes += sum(lw.pass_bytes(j, state.n_local) for j in range(s.pass_begin, s.pass_end))
                elif s.kind == "gemm":
                    n_rows_sel = 2 ** (state.n_local - s.nb - bin(s.ctrl_mask & (2**state.n_local - 1)).count("1"))
                    # out of place through a bounded scratch (<= 2^14 rows) + copy-back; QLS_GEMM_INPLACE=1: no scratch at all
                    in_place = lw.dtype == "fp64" and s.nb <= 10 and os.environ.get("QLS_GEMM_INPLACE", "0") == "1"
                    scratch_ptr = None if in_place else state.scratch(min(n_rows_sel, _abi.QLS_GEMM_CHUNK_ROWS) * 2**s.nb).data_ptr()
                    _abi.check(self.lib.qls_ctrl_gemm(state.ptr, scratch_ptr, state.n_local, self.abi_dtype, s.nb,
                                                      s.ctrl_mask, s.ctrl_val, state.index_offset,
                                                      self._gemm_mats[i].data_ptr(), stream))
            torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in pairs), launches, nbytes

    def _run_init(self, state: DeviceState, s: Step) -> None:
        k = len(s.qubits)
        if tuple(s.qubits) != tuple(range(k)):
            raise NotImplementedError("state preparation must act on the low qubits 0..k-1")
        state.init_basis(None)
        rank = state.index_offset >> state.n_local
        factor = (1.0 if rank == 0 else 0.0) if s.rank_factors is None else s.rank_factors[rank]
        if factor != 0:
            state.write(0, s.amps if factor == 1.0 else np.asarray(s.amps) * factor)
