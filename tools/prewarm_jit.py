"""Fill the on-disk cubin cache of the specialised pass kernels (NVRTC needs no GPU): the programs bench.py runs at
N = 1, 2, 4, 8 GPUs (33 + log2 N qubits), the sharded-vs-unsharded checks that precede the timed N > 1 runs
(26 + log2 N qubits) and the sizes the GPU tests use.  The cache directory travels with the repository snapshot, so a
fresh box starts without compiling.

    python tools/prewarm_jit.py [--quick]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="bench sizes only")
    args = ap.parse_args()
    from qls_b200 import jit
    from qls_b200.lower import LoweringOptions, lower_circuit
    from qls_b200.shard import plan_sharded
    from qls_b200.synthetic import hhl_shaped_circuit

    t0 = time.time()
    n_units = 0

    def single(n, block="chain2"):
        nonlocal n_units
        low = lower_circuit(hhl_shaped_circuit(n, block=block), LoweringOptions(dtype="fp64", zero_start=True))
        n_units += len(jit.compile_program(low))

    def sharded(n, world):
        nonlocal n_units
        plan = plan_sharded(hhl_shaped_circuit(n), world, LoweringOptions(dtype="fp64"), defer_final_1q=True)
        for seg in plan.segments:
            if seg.kind == "local":
                n_units += len(jit.compile_program(seg.lowered))

    sizes_single = [33, 30] + ([] if args.quick else [20, 21, 22, 23, 24, 25, 26, 27, 28, 29])
    for n in sizes_single:
        single(n)
        print(f"n = {n}: {n_units} units, {time.time() - t0:.0f} s", flush=True)
    for world, g in ((2, 1), (4, 2), (8, 3)):
        for n in [33 + g, 26 + g] + ([] if args.quick else [22 + g]):
            sharded(n, world)
            print(f"world {world}, n = {n}: {n_units} units, {time.time() - t0:.0f} s", flush=True)
    if not args.quick:
        single(30, block="dense5")
    print(f"{n_units} units in {time.time() - t0:.0f} s; cache: {jit.CACHE_DIR}")


if __name__ == "__main__":
    main()
