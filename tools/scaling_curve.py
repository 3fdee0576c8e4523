"""Size sweep of the headline workload on ONE GPU in one process: for each n, lower the HHL-shaped
circuit, warm up, then time K device-resident solves and the tile-pass kernels alone with CUDA events
(the same quantities as bench.py's `value` and `roofline`).  One JSON line per size on stdout.

    python tools/scaling_curve.py 26 28 29 30 31 32 [--steps 3] [--warmup 3] [--precision fp64]
"""
import argparse
import gc
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("qubits", type=int, nargs="+")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
    args = ap.parse_args()

    import numpy as np
    import torch

    import qls_b200  # noqa: F401
    from bench import _peaks
    from qls_b200.solver_shaped import ShapedSolver
    from qls_b200.synthetic import hhl_shaped_circuit

    peak, _src = _peaks()
    cdt = torch.complex128 if args.precision == "fp64" else torch.complex64
    for n in args.qubits:
        qc = hhl_shaped_circuit(n)
        nb = qc.qregs[0].size
        solver = ShapedSolver(qc, nb, precision=args.precision, device=0)
        low = solver.lowered
        rng = np.random.default_rng(99)
        b = rng.standard_normal(2**nb) + 0j
        b_dev = torch.from_numpy(b / np.linalg.norm(b)).to(cdt).to("cuda:0")
        for _ in range(args.warmup):
            solver.run_device(b_dev)
            solver.reduce()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            solver.run_device(b_dev)
            p_flag, norm_sq = solver.reduce()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        kern_ms, launches, kern_bytes = solver.time_passes(args.steps)
        achieved = kern_bytes / (kern_ms * 1e-3) / 1e9
        print(json.dumps({
            "workload": f"hhl_shaped_n{n}", "precision": args.precision, "state_bytes": low.amp_bytes() << n, "ms_per_step": ms,
            "amp_updates_per_s": low.amp_updates() / (ms * 1e-3), "passes": launches, "kernel_ms_per_step": kern_ms,
            "algorithmic_bytes_per_step": kern_bytes, "achieved_GBps": achieved, "hbm_peak_GBps": peak, "frac": achieved / peak,
            "real_tfma_per_s": low.fp64_fma_count() / (kern_ms * 1e-3) / 1e12, "p_flag": p_flag, "steps": args.steps,
            "warmup": args.warmup}), flush=True)
        del solver, b_dev
        gc.collect()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
