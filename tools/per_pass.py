"""Per-pass times of the HHL-shaped program, specialised kernels vs interpreter kernels (A/B in one process).

    python tools/per_pass.py --qubits 30 [--reps 2]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

HBM = 6543.4e9
DFMA = 18.4e12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--block", default="chain2")
    ap.add_argument("--modes", default="jit,interp")
    args = ap.parse_args()
    import torch

    from qls_b200 import _abi
    from qls_b200.solver_shaped import ShapedSolver
    from qls_b200.synthetic import hhl_shaped_circuit

    qc = hhl_shaped_circuit(args.qubits, block=args.block)
    solver = ShapedSolver(qc, qc.qregs[0].size)
    low = solver.lowered
    lib = _abi.load()
    res = {}
    for mode in args.modes.split(","):
        _abi.check(lib.qls_set_use_jit(1 if mode == "jit" else 0))
        best = None
        for _ in range(args.reps + 1):
            solver.state.init_basis(None)
            solver.state.write(0, solver._init_step.amps)
            per = solver.program.run_each_pass_timed(solver.state)
            best = per if best is None else [(j, min(a, b)) for (j, a), (_, b) in zip(best, per)]
        res[mode] = best
        p_flag, norm_sq = solver.reduce()
        print(f"# {mode}: total {sum(ms for _, ms in best):.2f} ms  P(flag=1)={p_flag:.12f} |x|^2={norm_sq:.6e}", flush=True)
    _abi.check(lib.qls_set_use_jit(1))
    modes = list(res)
    fma = None
    for idx, (j, _) in enumerate(res[modes[0]]):
        pi = low.pass_info[j]
        nbytes = low.pass_bytes(j)
        cols = "  ".join(f"{m} {res[m][idx][1]:8.3f} ms {nbytes / res[m][idx][1] / 1e6:7.0f} GB/s" for m in modes)
        print(f"pass {j:2d} L={pi.low_bits:2d} fix={pi.fixed_bits} stages={pi.rs_stages} hbm-floor {nbytes / HBM * 1e3:7.3f} ms  {cols}  high={pi.high}")
    tot_bytes = sum(low.pass_bytes(j) for j, _ in res[modes[0]])
    for m in modes:
        t = sum(ms for _, ms in res[m])
        print(f"# {m}: {t:.2f} ms, {tot_bytes / t / 1e6:.0f} GB/s = {tot_bytes / t / 1e6 / (HBM / 1e9):.3f} of HBM peak; "
              f"fp64 {low.fp64_fma_count() / t / 1e9:.2f} TFMA/s = {low.fp64_fma_count() / t / 1e9 / (DFMA / 1e12):.3f} of FP64 peak")


if __name__ == "__main__":
    main()
