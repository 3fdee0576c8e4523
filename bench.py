#!/usr/bin/env python
"""Benchmark of the hot path: HHL-shaped statevector solves on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--qubits n] [--impl reference]

One *step* = one full solve of the synthetic HHL-shaped circuit (BASELINE.json configs[3] at
N = 1: QPE + uniformly-controlled RY + inverse QPE, 33 qubits fp64 = 128 GiB state; configs[4] =
the same circuit at 33 + log2(N) qubits sharded over N GPUs): state preparation, every fused tile
pass, the reductions P(flag=1) and |x|^2.  The headline is ms per solve (``ms_per_step``); ``value`` restates it as
amplitude updates per second with the job defined on the SOURCE circuit -- the sum over its leaf gates of 2^(n-c),
c = number of control qubits of the gate -- so that it does not depend on how the lowering
fuses gates, and is the same job for the CPU arm.

Prints ONE JSON line (see the task contract): value (device-timed, inputs resident in HBM), e2e
(host buffers in and out through ShapedSolver.solve), roofline (tile-pass kernel vs measured HBM
peak), cpu_baseline (oracle/numpy_sv.py = the reference's qiskit Statevector algorithm restated,
on a bounded smaller instance of the same circuit), clocks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hhl_amp_updates_per_s"


def _traffic(algorithmic_bytes_per_launch: float):
    """DRAM bytes per launch of the dominant kernel: the read+write / algorithmic ratio of the ncu
    --set full capture summarised in profiles/traffic.json (tools/summarize_ncu.py), applied to this
    run's launch size.  None if no capture is on file."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as fh:
        t = json.load(fh)
    return t["dram_bytes_per_algorithmic_byte"] * algorithmic_bytes_per_launch, t["source"]

UNIT = "amp-updates/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def source_amp_updates(qc) -> tuple:
    """(sum over the leaf gates of the source circuit of 2^(n - controls), number of leaf gates)."""
    from qls_b200.circuit import flatten

    leaves, _ = flatten(qc)
    n = qc.num_qubits
    total = 0
    for lf in leaves:
        total += 2 ** len(lf.targets) if lf.kind == "init" else 2 ** (n - len(lf.controls))
    return total, len(leaves)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm / baseline: the oracle (restated qiskit Statevector path) on a bounded instance
# ------------------------------------------------------------------------------------------------


def _use_all_host_threads() -> int:
    """The BLAS behind NumPy reads its thread count when it is loaded: set it BEFORE the first ``import numpy``
    (``torch.distributed.run`` exports OMP_NUM_THREADS=1 to its workers)."""
    cores = os.cpu_count() or 1
    if "numpy" not in sys.modules:
        for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[var] = str(cores)
    return cores


def cpu_reference_sample(n_cpu: int, repeats: int = 1):
    """Time the unfused leaf stream of the HHL-shaped circuit at ``n_cpu`` qubits through
    ``oracle.numpy_sv`` (reshape/transpose/np.dot per gate).  Returns (source-gate amp updates, seconds, leaves)."""
    _use_all_host_threads()
    from oracle import numpy_sv
    from oracle.stream import leaves_to_stream
    from qls_b200.circuit import flatten
    from qls_b200.synthetic import hhl_shaped_circuit

    qc = hhl_shaped_circuit(n_cpu)
    leaves, phase = flatten(qc)
    stream = leaves_to_stream(leaves, phase)
    job, _ = source_amp_updates(qc)  # same job definition as the GPU arm
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        psi = numpy_sv.run(n_cpu, stream)
        half = psi.size // 2
        _ = numpy_sv.prob_mask(psi, half, half), float(abs(psi[half: half + 2 ** ((n_cpu - 1) // 2)] ** 2).sum())
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return job, best, len(leaves)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_cpu = args.cpu_qubits
    times = []
    job = leaves = 0
    for i in range(args.warmup + args.steps):
        job, dt, leaves = cpu_reference_sample(n_cpu)
        if i >= args.warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    value = job / sec
    cores = os.cpu_count() or 1
    n_gpu = args.qubits or 33
    sample = (f"HHL-shaped circuit at n={n_cpu} qubits fp64 ({leaves} unfused leaf gates, one reshape/transpose/np.dot "
              f"pass each = qiskit Statevector._evolve_operator restated in oracle/numpy_sv.py); BLAS threads = "
              f"{os.environ.get('OMP_NUM_THREADS')}; same_config: false (the GPU arm runs n={n_gpu}: the per-amplitude cost of "
              f"this algorithm does not depend on n, a solve at n={n_gpu} would take about {sec * 2.0 ** (n_gpu - n_cpu):.0f} s)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"hhl_shaped_n{n_cpu}_cpu_sample_of_n{n_gpu}", "n_qubits": n_cpu, "same_config": False,
                   "job": "source gates: sum over leaf gates of 2^(n - controls)", "source_gates": leaves,
                   "amp_updates_per_step": job, "ms_per_solve": sec * 1e3,
                   "ms_per_solve_extrapolated_to_gpu_arm_size": sec * 1e3 * 2.0 ** (n_gpu - n_cpu)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------


def sharded_parity_check(world: int, rank: int, local_rank: int, precision: str, n_local: int = 26,
                         chunk_amps: int = 1 << 20) -> dict:
    """Before anything is timed: the same generator at ``n_local + log2(world)`` qubits, sharded over all ranks and
    unsharded on rank 0, same right-hand side -- solution slice, P(flag = 1) and |x|^2 must agree to 1e-12 (fp64).
    Every rank raises if they do not."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from qls_b200.shard import ShardedShapedSolver
    from qls_b200.solver_shaped import ShapedSolver
    from qls_b200.synthetic import hhl_shaped_circuit

    n = n_local + int(np.log2(world))
    qc = hhl_shaped_circuit(n)
    nb = qc.qregs[0].size
    b = np.random.default_rng(7).standard_normal(2**nb) + 0j
    b /= np.linalg.norm(b)
    # (small chunks: at this size the exchanges then split into halves like the full-size ones do, so the passes
    # overlapped with them -- half a shard at a time -- are part of what is checked)
    sharded = ShardedShapedSolver(qc, nb, precision=precision, device=local_rank, chunk_amps=chunk_amps)
    xs, ps, ns = sharded.solve(b)
    overlapped = {str(j): {"split_positions": list(v[0]), "parts": 1 << len(v[0]), "passes_before": len(v[2]), "passes_after": len(v[3])}
                  for j, v in (getattr(sharded, "_overlap", None) or {}).items()}
    transport = "peer memory" if getattr(sharded.engine, "peers", None) else "nccl"
    out = torch.zeros(4, dtype=torch.float64, device=f"cuda:{local_rank}")
    if rank == 0:
        single = ShapedSolver(qc, nb, precision=precision, device=local_rank)
        x1, p1, n1 = single.solve(b)
        out[0] = float(np.max(np.abs(xs - x1)))
        out[1] = abs(ps - p1)
        out[2] = abs(ns - n1)
        out[3] = float(np.max(np.abs(x1)))
        del single
    del sharded
    torch.cuda.empty_cache()
    dist.broadcast(out, 0)
    dx, dp, dn, xmax = (float(v) for v in out.tolist())
    tol = 1e-12 if precision == "fp64" else 1e-5
    res = {"n_qubits": n, "ranks": world, "shard_parity_max_abs": max(dx, dp, dn), "x_slice_max_abs": dx, "p_flag1_abs": dp,
           "norm_sq_abs": dn, "x_slice_max_amplitude": xmax, "tolerance": tol, "transport": transport,
           "exchanges_overlapped_with_passes": overlapped}
    if not max(dx, dp, dn) <= tol:
        raise SystemExit(f"sharded / unsharded mismatch: {res}")
    return res


def state_checks(solver, world: int, precision: str) -> dict:
    """Size-independent properties of the evolved state (unitary circuit on a unit vector): total norm, P(flag=0) +
    P(flag=1), <Z_flag> = P0 - P1.  Raises (the run fails) outside 1e-11 in fp64."""
    import torch

    st = solver.state
    n_local = st.n_local
    total = st.norm_sq_range(0, 2**n_local)
    if world > 1:
        import torch.distributed as dist

        flag_mask = solver._masks()[0]
        p1, p0 = st.prob_mask(flag_mask, flag_mask), st.prob_mask(flag_mask, 0)
        t = torch.tensor([total, p0, p1], dtype=torch.float64, device=st.tensor.device)
        dist.all_reduce(t)
        total, p0, p1 = (float(v) for v in t.tolist())
        z = None
    else:
        half = 2 ** (n_local - 1)
        p1, p0 = st.prob_mask(half, half), st.prob_mask(half, 0)
        z = st.expectation_z(n_local - 1)
    tol = 1e-11 if precision == "fp64" else 1e-4
    out = {"state_norm_sq": total, "p_flag0_plus_p_flag1": p0 + p1, "z_flag": z, "p_flag0_minus_p_flag1": p0 - p1, "tolerance": tol}
    bad = abs(total - 1.0) > tol or abs(p0 + p1 - 1.0) > tol or (z is not None and abs(z - (p0 - p1)) > tol)
    out["passed"] = not bad
    if bad:
        raise SystemExit(f"state checks failed: {out}")
    return out


def pick_qubits(requested: int, free_bytes: int) -> int:
    if requested:
        return requested
    n = 33
    while n > 20 and (16 << n) + (6 << 30) > free_bytes:
        n -= 1
    return n


def run_gpu(args):
    # Libraries (NCCL's version banner) write to the C-level stdout: keep fd 1 for the ONE JSON line
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 through torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    from qls_b200.lower import LoweringOptions
    from qls_b200.solver_shaped import ShapedSolver
    from qls_b200.synthetic import hhl_shaped_circuit

    free_bytes, _total = torch.cuda.mem_get_info()
    shard_parity = None
    if world > 1:
        from qls_b200.shard import ShardedShapedSolver  # sharded states: 33 + log2(N) qubits

        shard_parity = sharded_parity_check(world, rank, local_rank, args.precision)

        n = args.qubits or (pick_qubits(0, free_bytes) + int(np.log2(world)))
        qc = hhl_shaped_circuit(n)
        nb = qc.qregs[0].size
        solver = ShardedShapedSolver(qc, nb, precision=args.precision, device=local_rank)
    else:
        n = pick_qubits(args.qubits, free_bytes)
        qc = hhl_shaped_circuit(n)
        nb = qc.qregs[0].size
        solver = ShapedSolver(qc, nb, precision=args.precision, device=local_rank)
    low = solver.lowered
    amp_updates, n_leaves = source_amp_updates(qc)
    fused_block_updates = solver.amp_updates_global() if world > 1 else low.amp_updates()
    S = low.amp_bytes()

    rng = np.random.default_rng(99)
    b = rng.standard_normal(2**nb) + 0j
    b /= np.linalg.norm(b)
    cdt = torch.complex128 if args.precision == "fp64" else torch.complex64
    b_pinned = torch.from_numpy(b).to(cdt).pin_memory()
    b_dev = b_pinned.to(f"cuda:{local_rank}")
    b_host = b_pinned.numpy()

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        solver.run_device(b_dev)
        return solver.reduce()

    # ---- warm-up ------------------------------------------------------------------------------
    for _ in range(args.warmup):
        device_step()
    barrier()

    # ---- timed: device-resident inputs ----------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    last = None
    for _ in range(args.steps):
        last = device_step()
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1) / args.steps
    checks = state_checks(solver, world, args.precision)  # on the state the last timed solve left in HBM

    # ---- timed: tile-pass kernels alone (roofline numerator), CUDA events on the launch stream --
    kern_ms, kern_launches, kern_bytes = solver.time_passes(args.steps)
    if args.profile_passes and rank == 0 and world == 1:
        solver.state.init_basis(None)
        solver.state.write(0, b_host)
        per = solver.program.run_each_pass_timed(solver.state)
        for j, ms in per:
            pi = low.pass_info[j]
            ps = low.passes[j]
            recs = [low.ops[t] for t in range(ps.op_begin, ps.op_begin + ps.op_count)]
            sweeps = [(r.kind, r.k) for r in recs if not (r.flags & 12)]
            gbs = low.pass_bytes(j) / (ms * 1e-3) / 1e9
            print(f"pass {j:3d} {ms:9.3f} ms {gbs:8.1f} GB/s L={pi.low_bits} high={pi.high} fix={pi.fixed_bits} "
                  f"recs={pi.n_ops} sweeps={sweeps}", file=sys.stderr)

    # ---- timed: end to end, host buffers in and out --------------------------------------------
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        x, p_flag, norm_sq = solver.solve(b_host)
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([dev_ms, e2e_ms, kern_ms or 0.0], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, kern_ms = (float(v) for v in t.tolist())

    if rank == 0:
        peak, peak_src = _peaks()
        st = solver.stats()
        roofline = None
        if kern_ms:
            achieved = kern_bytes / (kern_ms * 1e-3) / 1e9
            traffic, traffic_src = _traffic(kern_bytes / max(1, kern_launches))
            fma = low.fp64_fma_count() if world == 1 else 0.0
            jit_n = getattr(solver.program, "jit_passes", 0)
            # second view of the same measurement: every pass has an HBM floor AND an FP64 floor and cannot beat the larger
            # one (the data-register passes are FP64-bound), so the sum of the larger floors bounds the pass time from below
            two_roof = None
            if world == 1:
                peak_fma = 18.4e12
                floors = [max(low.pass_bytes(i) / (peak * 1e9), low.pass_fma_count(i) / peak_fma) for i in range(low.n_passes)]
                two_roof = {"floor_ms": sum(floors) * 1e3, "hbm_floor_ms": kern_bytes / (peak * 1e9) * 1e3,
                            "fp64_floor_ms": fma / peak_fma * 1e3, "frac": sum(floors) * 1e3 / kern_ms,
                            "definition": "sum over passes of max(pass bytes / HBM peak, pass FMAs / FP64 peak) / measured pass time"}
            roofline = {"bound": "hbm", "kernel": "qls_rs2_pass (specialised per pass, NVRTC)" if jit_n else "qls_rs_pass_kernel",
                        "specialised_passes": jit_n, "landing": getattr(solver.program, "jit_modes", None), "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                        "launches_per_step": kern_launches, "algorithmic_bytes_per_step": kern_bytes,
                        "algorithmic_bytes_per_launch": kern_bytes / max(1, kern_launches),
                        "avg_launch_ms": kern_ms / max(1, kern_launches), "kernel_ms_per_step": kern_ms,
                        "two_roof": two_roof,
                        # second ceiling of the same kernel: the FP64 pipe (DFMA/s measured by tools/microbench/fp64_pipes.cu)
                        "fp64": ({"fma_per_step": fma, "achieved_tfma_s": fma / (kern_ms * 1e-3) / 1e12, "peak_tfma_s": 18.4,
                                  "frac": fma / (kern_ms * 1e-3) / 18.4e12} if world == 1 else None)}
        # CPU baseline on a bounded sample of the same workload (N = 1 only: torchrun pins OMP threads to 1)
        cpu = None
        if world == 1:
            job_cpu, sec_cpu, leaves = cpu_reference_sample(args.cpu_qubits)
            cpu = {"value": job_cpu / sec_cpu, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": f"same circuit at n={args.cpu_qubits} ({leaves} unfused leaf gates through oracle/numpy_sv.py, "
                             f"{sec_cpu:.2f} s); per-amplitude cost is size independent"}
        launches_per_step = solver.program.launches_per_run + 2 + 4  # + init fill/copy + two 2-kernel reductions
        remap = None
        if world > 1:
            sent = st["remap_bytes_sent_per_gpu"]
            ms = getattr(solver, "remap_ms", 0.0)
            p2p = bool(getattr(solver.engine, "peers", None))
            push = os.environ.get("QLS_REMAP_PUSH", "1")
            if p2p and push == "0":  # one swap kernel per pairwise round
                launches_per_step += sum(2 ** len(sg.pairs) - 1 for sg in solver.plan.segments if sg.kind == "swap")
                how = "peer memory (CUDA IPC over NVLink): one in-place swap kernel per pairwise round"
            elif p2p:
                # exchanges of consecutive positions: copy engines only (no kernel of ours); others: one gather + one scatter
                # kernel per chunk.  Passes overlapped with an exchange are launched once per part of the shard.
                swaps = [sg for sg in solver.plan.segments if sg.kind == "swap"]
                ce = [solver.runner.ce_capable([l for _, l in sg.pairs]) for sg in swaps]
                for sg, is_ce in zip(swaps, ce):
                    if not is_ce:
                        vol = int((1 - 0.5 ** len(sg.pairs)) * 2 ** solver.plan.n_local)
                        launches_per_step += 2 * max(1, vol // solver.runner.chunk)
                ov = getattr(solver, "_overlap", None) or {}
                for hs, _, pa, pb in ov.values():
                    launches_per_step += ((1 << len(hs)) - 1) * (len(pa) + len(pb))
                how = ("peer memory (CUDA IPC over NVLink): " + (
                    "pitched copies by the copy engines into the partner's receive slot and back (no kernel), hand-shake per chunk"
                    if all(ce) else "gather kernel -> copy engines into the partner's receive slot -> scatter kernel")
                    + f", 2 GiB chunks, two slots; {len(ov)} of {len(swaps)} exchanges overlapped with the passes beside them"
                    if push != "sm" else "peer memory (CUDA IPC over NVLink): remote stores from the gather kernel, scatter kernel")
            else:  # one pack + one unpack per chunk on the wire
                launches_per_step += 2 * max(1, sent // (S * solver.runner.chunk))
                how = "pack + NCCL send/recv + unpack, chunked"
            fallback = getattr(solver, "transport_fallback", None)
            remap = {"remapped_qubits_per_step": st["remaps"], "exchanges_per_step": st["remap_exchanges"],
                     "bytes_sent_per_gpu": sent, "ms_per_step": ms,
                     "achieved_gbs_per_direction": (sent / (ms * 1e-3) / 1e9) if ms else None, "peak": 770.0,
                     "peak_source": "measured peer copy per direction (B200_PROFILING.md)",
                     "frac": (sent / (ms * 1e-3) / 1e9 / 770.0) if ms else None,
                     "transport": how, "transport_fallback_reason": fallback,
                     "deferred_final_gates": len(getattr(solver.plan, "deferred", {}) or {}),
                     "overlap": {str(j): {"split_positions": list(v[0]), "parts": 1 << len(v[0]), "passes_before": len(v[2]),
                                          "passes_after": len(v[3])} for j, v in (getattr(solver, "_overlap", None) or {}).items()},
                     "includes": "ms_per_step = the exchanges ALONE (measured apart from the passes, hand-shakes included); inside "
                                 "a solve they overlap the passes listed under overlap, so ms_per_solve < pass kernels + exchanges"}
        line = {
            "metric": METRIC, "value": amp_updates / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms, "ms_per_solve": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.precision == "fp64" else "f32", "data": "synthetic",
            "config": {"workload": f"hhl_shaped_n{n}" + (f"_sharded_{world}gpu" if world > 1 else ""),
                       "n_qubits": n, "nb": nb, "nl": n - 1 - nb, "state_bytes_per_gpu": S << low.n_local,
                       "fused_passes": st["fused_passes"], "fused_blocks": st["fused_blocks"],
                       "source_gates": n_leaves, "amp_updates_per_step": amp_updates,
                       "job": "source gates: sum over the leaf gates of the circuit of 2^(n - controls)",
                       "fused_block_amp_updates_per_step": fused_block_updates,
                       "hbm_bytes_per_step_per_gpu": st["pass_bytes_per_gpu"] + (S << low.n_local),
                       "l2_policy": "inputs_larger_than_L2" if (S << low.n_local) > (256 << 20) else "state_fits_L2_no_flush",
                       "program": "lowered once outside the timed region; b differs per solve only through init"},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": amp_updates / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(b_host.nbytes), "d2h_bytes_per_step": int(x.nbytes) + 2 * 16,
                    "api": "ShapedSolver.solve(b_host) -> (x_host, P(flag=1), |x|^2)"},
            "remap": remap,
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
            "result_check": dict({"p_flag1": last[0], "norm_sq": last[1], "e2e_p_flag1": p_flag,
                                  "e2e_matches_device_run": abs(last[0] - p_flag) <= 1e-12 and abs(last[1] - norm_sq) <= 1e-12},
                                 **checks, **({"sharded_vs_unsharded": shard_parity} if shard_parity else {})),
        }
        if args.all_configs and world == 1:
            # BASELINE.json configs[1] and configs[2] on the same box (secondary lines; not the headline)
            del solver
            torch.cuda.empty_cache()
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import run_configs

                line["other_configs"] = {"hhl_tridiagonal_n23": run_configs.config2(), "vqls_batch_4096x10q": run_configs.config3(),
                                         "e2e_plugin_hhl_tridiagonal_n23": run_configs.plugin_e2e(),
                                         "hhl_shaped_n30_dense5": run_configs.config4_dense5()}
            except Exception as exc:  # never lose the headline line because a secondary config failed
                line["other_configs"] = {"error": f"{type(exc).__name__}: {exc}"}
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--qubits", type=int, default=0, help="total qubits (default: 33 on 1 GPU if it fits)")
    ap.add_argument("--cpu-qubits", type=int, default=22, help="size of the bounded CPU sample")
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--profile-passes", action="store_true", help="print per-pass timings to stderr")
    ap.add_argument("--no-all-configs", dest="all_configs", action="store_false",
                    help="skip the secondary configs (HHL tridiagonal n=23, VQLS batch)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
