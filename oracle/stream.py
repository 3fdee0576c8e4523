"""Turn a flattened leaf stream (duck-typed: ``kind``, ``targets``, ``payload``, ``controls``,
``ctrl_state``) into the plain ``(tag, matrix, qargs)`` stream of ``oracle/numpy_sv.py``, with
controls expanded into full little-endian matrices exactly as qiskit would hold them
(``qargs = controls + targets``).  TEST / BASELINE INFRASTRUCTURE ONLY."""
import math

import numpy as np


def controlled_matrix(base: np.ndarray, num_ctrl: int, ctrl_state: int) -> np.ndarray:
    k = int(round(math.log2(base.shape[0])))
    dim = 2 ** (k + num_ctrl)
    full = np.eye(dim, dtype=complex)
    sel = [i for i in range(dim) if (i & (2**num_ctrl - 1)) == ctrl_state]
    full[np.ix_(sel, sel)] = base
    return full


def ucry_matrix(angles: np.ndarray) -> np.ndarray:
    dim = 2 * len(angles)
    m = np.zeros((dim, dim), dtype=complex)
    for v, th in enumerate(angles):
        c, s = math.cos(th / 2), math.sin(th / 2)
        m[2 * v, 2 * v], m[2 * v, 2 * v + 1], m[2 * v + 1, 2 * v], m[2 * v + 1, 2 * v + 1] = c, -s, s, c
    return m


def leaves_to_stream(leaves, global_phase=0.0):
    stream = []
    if global_phase:
        stream.append(("phase", global_phase))
    for lf in leaves:
        if lf.kind == "init":
            stream.append(("init", np.asarray(lf.payload), list(lf.targets)))
            continue
        if lf.kind == "dense":
            base = np.asarray(lf.payload)
        elif lf.kind == "diag":
            base = np.diag(np.asarray(lf.payload))
        elif lf.kind == "ucry":
            base = ucry_matrix(np.asarray(lf.payload))
        else:
            raise ValueError(lf.kind)
        nc = len(lf.controls)
        mat = controlled_matrix(base, nc, lf.ctrl_state) if nc else base
        stream.append(("u", mat, list(lf.controls) + list(lf.targets)))
    return stream
