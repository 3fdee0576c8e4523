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
    return {"workload": f"hhl_tridiagonal_N2^{nb}", "n_qubits": n, "nl": hhl.nl, "ms_per_solve": ms, "gemm_blocks": n_gemm,
            "gemm_flops_per_solve": flops, "fp64_tflops_if_all_gemm": flops / (ms * 1e-3) / 1e12,
            "tile_passes": low.n_passes, "amp_updates_per_s": low.amp_updates() / (ms * 1e-3),
            "relative_distance_to_classical": rel, "host_build_s": t_build, "host_lower_upload_s": t_lower}


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
    vals = [v.cpu().numpy() for v in vals]
    return {"workload": f"vqls_hadamard_batch_{n_circuits}x{n}q", "circuits_per_launch": n_circuits, "launches": len(progs),
            "ms_per_batch": ms, "circuits_per_s": k / (ms * 1e-3),
            "amp_updates_per_s": k * sum(len(pi.op_controls) for pi in progs[0].lowered.pass_info) * 2**n / (ms * 1e-3),
            "mean_re": float(vals[0].mean()), "mean_im": float(vals[1].mean()),
            "param_bytes_in": int(theta.nbytes), "result_bytes_out": 8 * k}


if __name__ == "__main__":
    which = sys.argv[1:] or ["2", "3"]
    if "2" in which:
        print(json.dumps(config2()))
    if "3" in which:
        print(json.dumps(config3()))
