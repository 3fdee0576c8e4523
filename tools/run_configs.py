"""Run BASELINE.json configs[1] (HHL on tridiagonal N=2^10, n=23) and configs[2] (4096 Hadamard-test
circuits at 10 qubits) once and print timings + sanity checks.  Used by bench.py (--all-configs)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def config2(nb=10, precision="fp64", repeats=3):
    import numpy as np
    import torch

    from qls_b200.engine import DeviceProgram, DeviceState
    from qls_b200.hhl import HHL
    from qls_b200.lower import LoweringOptions, lower_circuit
    from qls_b200.toymodels import TridiagonalToeplitz

    model = TridiagonalToeplitz(nb)
    t0 = time.perf_counter()
    hhl = HHL(power_mode="eig", precision=precision)
    qc = hhl.construct_circuit(model.matrix_a, model.vector_b)
    t_build = time.perf_counter() - t0
    n = qc.num_qubits
    t0 = time.perf_counter()
    low = lower_circuit(qc, LoweringOptions(dtype=precision))
    prog = DeviceProgram(low)
    t_lower = time.perf_counter() - t0
    st = DeviceState(n, precision)
    prog.run(st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(repeats):
        prog.run(st)
        half = 2 ** (n - 1)
        norm2 = st.norm_sq_range(half, 2**nb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / repeats
    x = np.real(st.read(half, 2**nb))
    x = x / np.linalg.norm(x) * (np.sqrt(norm2) / hhl.scaling)
    cs = model.classical_solution
    rel = float(np.linalg.norm(cs - x) / np.linalg.norm(cs))
    n_gemm = sum(1 for s in low.steps if s.kind == "gemm")
    flops = n_gemm * 8 * 2**nb * 2 ** (n - 1)
    # one controlled product alone (the same call the program makes), CUDA events on the launch stream
    gemm_ms = None
    gi = next((i for i, s_ in enumerate(low.steps) if s_.kind == "gemm"), None)
    if gi is not None:
        from qls_b200 import _abi

        s_ = low.steps[gi]
        in_place = os.environ.get("QLS_GEMM_INPLACE", "0") == "1" and precision == "fp64" and s_.nb <= 10
        rows_sel = 2 ** (n - s_.nb - bin(s_.ctrl_mask).count("1"))
        scratch = None if in_place else st.scratch(min(rows_sel, _abi.QLS_GEMM_CHUNK_ROWS) * 2**s_.nb).data_ptr()
        stream = int(torch.cuda.current_stream().cuda_stream)

        def one():
            _abi.check(prog.lib.qls_ctrl_gemm(st.ptr, scratch, n, prog.abi_dtype, s_.nb, s_.ctrl_mask, s_.ctrl_val, 0,
                                              prog._gemm_mats[gi].data_ptr(), stream))
        one()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(10):
            one()
        g1.record()
        torch.cuda.synchronize()
        gemm_ms = g0.elapsed_time(g1) / 10
    peak_tf = 36.9  # FP64 tensor pipe, microbenchmarked on this pool's B200 (profiles/r01_microbench_fp64_pipes.txt)
    gemm_tf = flops / n_gemm / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None
    return {"workload": f"hhl_tridiagonal_N2^{nb}", "n_qubits": n, "nl": hhl.nl, "ms_per_solve": ms, "gemm_blocks": n_gemm,
            "gemm_flops_per_solve": flops, "fp64_tflops_if_all_gemm": flops / (ms * 1e-3) / 1e12,
            "gemm_form": "in place (thread-block clusters)" if os.environ.get("QLS_GEMM_INPLACE", "0") == "1" else "out of place, bounded scratch + copy-back",
            "roofline": {"bound": "tensor", "kernel": "qls_ctrl_gemm (DMMA m8n8k4, one controlled product incl. copy-back)",
                         "ms_per_launch": gemm_ms, "achieved": gemm_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": gemm_tf / peak_tf if gemm_tf else None,
                         "peak_source": "DMMA microbenchmark on this pool (profiles/r01_microbench_fp64_pipes.txt)"},
            "tile_passes": low.n_passes, "amp_updates_per_s": low.amp_updates() / (ms * 1e-3),
            "relative_distance_to_classical": rel, "host_build_s": t_build, "host_lower_upload_s": t_lower}


def config4_dense5(n=30, repeats=2):
    """SURVEY.md section 8(d) config 4 AS WRITTEN: the controlled block is ONE Haar-random 5-qubit unitary (256 flop per 32
    bytes: beyond the FP64 ridge of a B200, so the roof is the FP64 pipe).  Dense blocks wider than 2 qubits have no
    register-stage form: their passes run on the shared-memory sweep kernel (csrc/tile_pass.cu)."""
    import torch

    from qls_b200.solver_shaped import ShapedSolver
    from qls_b200.synthetic import hhl_shaped_circuit

    qc = hhl_shaped_circuit(n, block="dense5")
    nb = qc.qregs[0].size
    solver = ShapedSolver(qc, nb)
    low = solver.lowered
    solver.run_device(None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(repeats):
        solver.run_device(None)
        p_flag, norm_sq = solver.reduce()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / repeats
    # flops of the dense blocks: 8 * 2^k real flops per amplitude touched, halved per control (incl. qubits still |0>)
    flops = 0.0
    sweep_passes = 0
    for i, pi in enumerate(low.pass_info):
        sweep_passes += pi.rs_stages == 0
        p = low.passes[i]
        for o in range(p.op_begin, p.op_begin + p.op_count):
            op = low.ops[o]
            if op.kind == 1 and op.k >= 1:
                nc = bin(op.xc_mask | p.fix_mask).count("1") + bin(op.lc_mask).count("1")
                flops += 8.0 * 2**op.k * 2.0 ** (n - nc)
    nbytes = low.total_pass_bytes()
    return {"workload": f"hhl_shaped_n{n}_dense5", "n_qubits": n, "ms_per_solve": ms, "fused_passes": low.n_passes,
            "passes_on_the_sweep_kernel": int(sweep_passes), "p_flag1": p_flag, "norm_sq": norm_sq,
            "roofline": {"bound": "fp64", "achieved": flops / (ms * 1e-3) / 1e12, "peak": 36.8, "unit": "TFLOP/s",
                         "frac": flops / (ms * 1e-3) / 1e12 / 36.8, "hbm_gbs": nbytes / (ms * 1e-3) / 1e9,
                         "note": "5-qubit dense blocks on csrc/tile_pass.cu (matrix in __constant__ memory, 8 accumulators per thread; no register-stage / specialised form yet)"}}


def plugin_e2e(nb=10, repeats=2):
    """The call a user of the reference makes (quantum_linear_solver.py:82-89 -> hhl_qiskit_implementation.py:20-79), end to end
    on BASELINE.json configs[1]: wall time of ``QuantumLinearSolver().solve(A, b, method="hhl_qiskit")`` including circuit
    construction, lowering + upload, the simulation, decompose(reps=10), the QASM3 text and the file it leaves behind."""
    import contextlib
    import io
    import tempfile

    import numpy as np

    from qls_b200.quantum_linear_solver import QuantumLinearSolver
    from qls_b200.toymodels import TridiagonalToeplitz

    model = TridiagonalToeplitz(nb)
    best, qls = None, None
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            for _ in range(repeats):
                qls = QuantumLinearSolver()
                t0 = time.perf_counter()
                with contextlib.redirect_stdout(io.StringIO()):
                    qls.solve(matrix_a=model.matrix_a, vector_b=model.vector_b, method="hhl_qiskit", file_basename="cfg2")
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            file_bytes = os.path.getsize("hhl_qiskit_circuit.qasm3")
        finally:
            os.chdir(cwd)
    cs = model.classical_solution
    rel = float(np.linalg.norm(cs - qls.solution) / np.linalg.norm(cs))
    return {"api": "QuantumLinearSolver().solve(A, b, method='hhl_qiskit')", "workload": f"hhl_tridiagonal_N2^{nb}",
            "wall_s": best, "reported_run_time_s": qls.run_time, "circuit_width": qls.circuit_width,
            "circuit_depth": qls.circuit_depth, "qasm_bytes": file_bytes, "relative_distance_to_classical": rel,
            "note": "host work dominates: eigen-decomposition + 2 nl dense powers of exp(iAt), lowering + upload, and a QASM3 "
                    "text that carries the 2 nl dense 2^nb x 2^nb unitaries losslessly; the GPU part is hhl_tridiagonal_n23.ms_per_solve"}


def config3(n_data=9, n_circuits=4096, precision="fp64", repeats=5, seed=4321):
    import numpy as np
    import torch

    from qls_b200.circuit import QuantumCircuit
    from qls_b200.estimator import BatchProgram
    from qls_b200.library import RealAmplitudes
    from qls_b200.synthetic import haar_unitary
    from qls_b200.vqls import HadamardTest

    rng = np.random.default_rng(7)
    ansatz = RealAmplitudes(n_data, "full", reps=3)
    op = QuantumCircuit(n_data)
    op.unitary(haar_unitary(n_data, rng), list(range(n_data)))
    test = HadamardTest([op], apply_initial_state=ansatz)
    theta = np.random.default_rng(seed).uniform(-np.pi, np.pi, (n_circuits, ansatz.num_parameters))
    out = {}
    progs = [BatchProgram(c, precision) for c in test.circuits]
    tdev = torch.from_numpy(theta).cuda()
    for p in progs:
        p.expect_z(0, tdev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(repeats):
        vals = [p.expect_z(0, tdev) for p in progs]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / repeats
    k = n_circuits * len(progs)
    n = n_data + 1
    # useful flops: the controlled dense operator on the ancilla = 1 half of every circuit (the ansatz is O(n 2^n))
    flops = len(progs) * n_circuits * 8 * (2**n_data) ** 2
    vals = [v.cpu().numpy() for v in vals]
    return {"workload": f"vqls_hadamard_batch_{n_circuits}x{n}q", "circuits_per_launch": n_circuits, "launches": len(progs),
            "ms_per_batch": ms, "circuits_per_s": k / (ms * 1e-3),
            "amp_updates_per_s": k * sum(len(pi.op_controls) for pi in progs[0].lowered.pass_info) * 2**n / (ms * 1e-3),
            "mean_re": float(vals[0].mean()), "mean_im": float(vals[1].mean()),
            "param_bytes_in": int(theta.nbytes), "result_bytes_out": 8 * k,
            "roofline": {"bound": "tensor", "kernel": "qls_batch_segment_kernel + qls_gemm_rows_inplace (whole launch)",
                         "achieved": flops / (ms * 1e-3) / 1e12, "peak": 36.9, "unit": "TFLOP/s",
                         "frac": flops / (ms * 1e-3) / 1e12 / 36.9,
                         "note": "dense-operator flops / whole batch time; the product itself runs in 2 x 0.56 ms "
                                 "(profiles/r02_gemm_paths.md), the per-circuit ansatz segments take the rest"}}


if __name__ == "__main__":
    which = sys.argv[1:] or ["2", "3"]
    if "2" in which:
        print(json.dumps(config2()))
    if "3" in which:
        print(json.dumps(config3()))
    if "4" in which:
        print(json.dumps(config4_dense5()))
    if "e2e" in which:
        print(json.dumps(plugin_e2e()))
