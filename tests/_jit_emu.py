"""Host emulation of the generated (specialised) pass units -- test infrastructure.

The source ``qls_b200.jit.generate_source`` emits for NVRTC is compiled here with g++ and
``-DQLS_HOST_EMU``: ``csrc/rs2_host_shim.h`` supplies double2, the barrier and 256 real threads, and
``csrc/rs2_kernel.cuh`` a plain loop over tiles instead of the producer/consumer kernel.  That runs the
generated index arithmetic (stage layouts, landing / exchange images, table indices, controls) on the
CPU so that it can be compared with the oracle without a GPU."""
import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np

from qls_b200 import jit

_CACHE = os.path.join(tempfile.gettempdir(), "qls_jit_emu")


def build(source: str) -> C.CDLL:
    os.makedirs(_CACHE, exist_ok=True)
    key = hashlib.sha256((source + open(os.path.join(jit.CSRC, "rs2_runtime.cuh")).read()
                          + open(os.path.join(jit.CSRC, "rs2_kernel.cuh")).read()
                          + open(os.path.join(jit.CSRC, "rs2_host_shim.h")).read()).encode()).hexdigest()[:24]
    so = os.path.join(_CACHE, key + ".so")
    if not os.path.exists(so):
        src = os.path.join(_CACHE, key + ".cpp")
        with open(src, "w") as f:
            f.write(source)
        cmd = ["g++", "-std=c++17", "-O1", "-march=native", "-ffp-contract=off", "-DQLS_HOST_EMU", "-I", jit.CSRC, "-shared", "-fPIC",
               "-pthread", src, "-o", so + ".tmp"]
        subprocess.run(cmd, check=True)
        os.replace(so + ".tmp", so)
    lib = C.CDLL(so)
    lib.qls_emu_run_pass.restype = None
    lib.qls_emu_run_pass.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
    return lib


def run_pass(state: np.ndarray, lowered, index: int, n_local: int, index_offset: int = 0, mode=None) -> bool:
    """Apply pass ``index`` of ``lowered`` to ``state`` (complex128, in place) through the generated
    code.  Returns False if the pass has no specialised form."""
    jp = jit.extract_pass(lowered, index)
    if jp is None:
        return False
    lib = build(jit.generate_source(jp, mode or os.environ.get("QLS_JIT_LANDING")))
    p = lowered.passes[index]
    n_tiles = 2 ** (n_local - p.tile_bits - bin(p.fix_mask).count("1"))
    assert state.dtype == np.complex128 and state.flags.c_contiguous
    pool = np.ascontiguousarray(lowered.pool)
    lib.qls_emu_run_pass(state.ctypes.data, index_offset, n_tiles, pool.ctypes.data)
    return True
