"""Timeline of CTA 0 of specialised pass kernels (debug build: QLS_JIT_TRACE=1 is set by this tool).

    python tools/jit_trace.py --qubits 28 --passes 11,8 [--dump]

Per pass: where the two consumer groups and the producer warp of one SM spend their time -- waiting for a
tile, in stage transitions, in the ops of a stage -- and how often both groups are outside an op phase at
the same time (the FP64 pipe idles)."""
import argparse
import os
import sys

os.environ["QLS_JIT_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CAP = 8192


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=28)
    ap.add_argument("--passes", default="")
    ap.add_argument("--dump", action="store_true")
    ap.add_argument("--block", default="chain2")
    args = ap.parse_args()
    import numpy as np
    import torch

    from qls_b200 import _abi
    from qls_b200.engine import _stream_ptr
    from qls_b200.solver_shaped import ShapedSolver
    from qls_b200.synthetic import hhl_shaped_circuit

    qc = hhl_shaped_circuit(args.qubits, block=args.block)
    solver = ShapedSolver(qc, qc.qregs[0].size)
    low = solver.lowered
    lib = _abi.load()
    solver.run_device(None)
    torch.cuda.synchronize()
    trace = torch.zeros(2 + 2 * CAP, dtype=torch.int64, device="cuda")
    _abi.check(lib.qls_set_trace_buffer(trace.data_ptr()))
    passes = [int(x) for x in args.passes.split(",")] if args.passes else list(range(low.n_passes))
    for j in passes:
        trace.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _abi.check(lib.qls_program_run(solver.program.handle, solver.state.ptr, solver.state.n_local, 0, j, j + 1, _stream_ptr()))
        b.record()
        torch.cuda.synchronize()
        t = trace.cpu().numpy().astype(np.uint64)
        ev = [(int((t[2 + 2 * k] >> np.uint64(48)) & np.uint64(0x7FFF)), int((t[2 + 2 * k] >> np.uint64(32)) & np.uint64(0xFFFF)),
               int(t[2 + 2 * k] & np.uint64(0xFFFFFFFF)), int(t[3 + 2 * k])) for k in range(CAP) if t[2 + 2 * k] >> np.uint64(63)]
        n = len(ev)
        # compare like with like: only the time window all three writers cover
        if ev:
            t_end = min(max(c for who, e, q, c in ev if who == w) for w in {who for who, _, _, _ in ev})
            ev = [x for x in ev if x[3] <= t_end]
        ev.sort(key=lambda e: e[3])
        pi = low.pass_info[j]
        print(f"== pass {j}: {a.elapsed_time(b):.3f} ms, L={pi.low_bits} high={pi.high} stages={pi.rs_stages}, {n} events")
        if not ev:
            continue
        t0 = ev[0][3]
        if args.dump:
            for who, e, q, c in ev[:400]:
                print(f"  {c - t0:9d} who={who} ev={e:2d} q={q}")
        # per consumer group: accumulate phase durations
        tot = {}
        per_group_intervals = {0: [], 1: []}
        for g in (0, 1):
            seq = [(e, q, c) for who, e, q, c in ev if who == g]
            for (e0, q0, c0), (e1, q1, c1) in zip(seq, seq[1:]):
                d = c1 - c0
                if e0 == 1 and e1 == 2:
                    key = "wait tile"
                elif e0 == 2:
                    key = "landing read / first stage entry"
                elif e0 >= 10 and e0 % 2 == 0 and e1 == e0 + 1:
                    key = "ops"
                    per_group_intervals[g].append((c0, c1))
                elif e0 >= 11 and e0 % 2 == 1 and e1 >= 10:
                    key = "transition"
                elif e0 >= 11 and e0 % 2 == 1 and e1 == 3:
                    key = "landing write"
                elif e0 == 3:
                    key = "final barrier + arrive"
                elif e0 == 4 and e1 == 1:
                    key = "ticket"
                else:
                    key = f"other {e0}->{e1}"
                tot.setdefault((g, key), [0, 0])
                tot[(g, key)][0] += d
                tot[(g, key)][1] += 1
        span = ev[-1][3] - t0
        tiles = len({q for who, e, q, c in ev if who < 2 and e == 4})
        print(f"  span {span} cycles, {tiles} tiles -> {span / max(tiles, 1):.0f} cycles per tile")
        for (g, key), (d, cnt) in sorted(tot.items()):
            print(f"  group {g} {key:34s} {100.0 * d / span:5.1f} % of span, {d / cnt:8.0f} cycles x {cnt}")
        # overlap of op phases
        marks = []
        for g in (0, 1):
            for c0, c1 in per_group_intervals[g]:
                marks += [(c0, 1), (c1, -1)]
        marks.sort()
        level, last, hist = 0, t0, {0: 0, 1: 0, 2: 0}
        for c, dl in marks:
            hist[level] += c - last
            level += dl
            last = c
        print(f"  groups in an op phase: none {100.0 * hist[0] / span:.1f} %, one {100.0 * hist[1] / span:.1f} %, both {100.0 * hist[2] / span:.1f} %")
        prod = [(e, q, c) for who, e, q, c in ev if who == 2]
        ptot = {}
        for (e0, q0, c0), (e1, q1, c1) in zip(prod, prod[1:]):
            key = {20: "wait done", 21: "store + read-out", 22: "issue load", 23: "loop"}.get(e0, str(e0))
            ptot.setdefault(key, [0, 0])
            ptot[key][0] += c1 - c0
            ptot[key][1] += 1
        for key, (d, cnt) in ptot.items():
            print(f"  producer {key:20s} {100.0 * d / span:5.1f} % of span, {d / cnt:8.0f} cycles x {cnt}")
    _abi.check(lib.qls_set_trace_buffer(None))


if __name__ == "__main__":
    main()
