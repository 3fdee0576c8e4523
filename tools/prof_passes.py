"""Run selected tile passes of the HHL-shaped program once each (for ncu captures).

    python tools/prof_passes.py --qubits 27 --passes 0,9,4
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=27)
    ap.add_argument("--passes", default="0")
    ap.add_argument("--precision", default="fp64")
    ap.add_argument("--block", default="chain2")
    args = ap.parse_args()
    import torch

    from qls_b200 import _abi
    from qls_b200.engine import _stream_ptr
    from qls_b200.solver_shaped import ShapedSolver
    from qls_b200.synthetic import hhl_shaped_circuit

    qc = hhl_shaped_circuit(args.qubits, block=args.block)
    solver = ShapedSolver(qc, qc.qregs[0].size, precision=args.precision)
    solver.run_device(None)  # warm-up: whole program once (also makes the state dense)
    torch.cuda.synchronize()
    low = solver.lowered
    for j in [int(x) for x in args.passes.split(",")]:
        pi = low.pass_info[j]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _abi.check(solver.program.lib.qls_program_run(solver.program.handle, solver.state.ptr, solver.state.n_local, 0, j, j + 1,
                                                      _stream_ptr()))
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print(f"pass {j}: {ms:.3f} ms {low.pass_bytes(j) / ms / 1e6:.1f} GB/s L={pi.low_bits} high={pi.high} recs={pi.n_ops}")


if __name__ == "__main__":
    main()
