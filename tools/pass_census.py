"""Per-pass census of the lowered HHL-shaped program: micro-ops per stage, FP64 floor vs HBM floor.

    python tools/pass_census.py --qubits 33
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

DFMA_PER_S = 18.4e12  # measured: profiles/r01_microbench_fp64_pipes.txt
HBM_BPS = 6543.4e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=33)
    ap.add_argument("--block", default="chain2")
    args = ap.parse_args()
    from qls_b200 import _abi
    from qls_b200.lower import LoweringOptions, lower_circuit
    from qls_b200.synthetic import hhl_shaped_circuit

    n = args.qubits
    low = lower_circuit(hhl_shaped_circuit(n, block=args.block), LoweringOptions(dtype="fp64"))
    tot_f = tot_m = 0.0
    for j, pi in enumerate(low.pass_info):
        p = low.passes[j]
        desc, dfma = [], 0.0
        for o in range(p.rs_begin, p.rs_begin + p.rs_count):
            op = low.rs_ops[o]
            if op.kind == _abi.QLS_RS_STAGE:
                desc.append("|" + ("D" if op.flags & 0x10 else ""))
                continue
            real = bool(op.flags & _abi.QLS_OP_FLAG_REAL)
            nx, nt, nr = bin(op.xc_mask).count("1"), bin(op.ctl_tmask).count("1"), bin(op.ctl_rmask).count("1")
            k, c = {_abi.QLS_RS_G1: ("g1", 4 if real else 8), _abi.QLS_RS_G2: ("g2", 8 if real else 16),
                    _abi.QLS_RS_DIAG: ("dg", 4), _abi.QLS_RS_MUXRY: ("mx", 4)}[op.kind]
            dfma += c * 0.5 ** (nx + nt + nr) * 2**n
            desc.append(k + ("r" if real else "") + (f"x{nx}" if nx else "") + (f"t{nt}" if nt else "") + (f"r{nr}" if nr else ""))
        tf, tm = dfma / DFMA_PER_S * 1e3, low.pass_bytes(j) / HBM_BPS * 1e3
        tot_f += tf
        tot_m += tm
        print(f"pass {j:2d} L={pi.low_bits} high={pi.high} fp64 {tf:7.2f} ms  hbm {tm:7.2f} ms  " + " ".join(desc))
    print(f"total: fp64 floor {tot_f:.1f} ms, hbm floor {tot_m:.1f} ms")


if __name__ == "__main__":
    main()
