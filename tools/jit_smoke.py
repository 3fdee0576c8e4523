"""Smallest end-to-end check of the specialised kernels on a GPU: one HHL-shaped program, specialised vs
interpreter kernels on the same lowered program (max |delta| over the whole state).

    python tools/jit_smoke.py [--qubits 20]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=20)
    args = ap.parse_args()
    import numpy as np
    import torch

    from qls_b200 import _abi
    from qls_b200.engine import DeviceProgram, DeviceState
    from qls_b200.lower import LoweringOptions, lower_circuit
    from qls_b200.synthetic import hhl_shaped_circuit

    n = args.qubits
    t0 = time.time()
    low = lower_circuit(hhl_shaped_circuit(n), LoweringOptions())
    prog = DeviceProgram(low)
    print(f"lowered + compiled {prog.jit_passes}/{low.n_passes} passes in {time.time() - t0:.1f} s, landing modes {prog.jit_modes}", flush=True)
    lib = _abi.load()
    a, b = DeviceState(n), DeviceState(n)
    _abi.check(lib.qls_set_use_jit(0))
    prog.run(b)
    torch.cuda.synchronize()
    print("interpreter kernels done", flush=True)
    _abi.check(lib.qls_set_use_jit(1))
    prog.run(a)
    torch.cuda.synchronize()
    print("specialised kernels done", flush=True)
    x, y = a.read(), b.read()
    print(f"n={n} max|delta|={np.max(np.abs(x - y)):.3e} norm={np.vdot(x, x).real:.15f}", flush=True)
    if np.max(np.abs(x - y)) > 1e-12:
        sys.exit(1)


if __name__ == "__main__":
    main()
