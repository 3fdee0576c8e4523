"""Driver for profiling the GEMM-shaped paths alone (ncu -k regex:gemm ...):  python tools/prof_gemm.py 2|3 [repeats]
2 = BASELINE.json configs[1] (HHL on the tridiagonal N = 2^10 matrix, n = 23: 24 controlled dense products),
3 = configs[2] (two launches of 4096 Hadamard-test circuits at 10 qubits)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import run_configs  # noqa: E402

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "2"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    print(json.dumps(run_configs.config2(repeats=reps) if which == "2" else run_configs.config3(repeats=reps)))
