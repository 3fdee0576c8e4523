"""Reusable solver for HHL-shaped circuits: lower once, then solve for many right-hand sides.

The circuit (QPE, eigenvalue inversion, inverse QPE) does not depend on ``b``; only the state
preparation does, and the engine short-cuts that to writing ``b/|b|`` into the low amplitudes.
So one lowered program serves every ``b``; ``solve(b)`` takes a HOST vector and returns HOST
results, with the statevector never leaving HBM:

    H2D   2**nb amplitudes (b)
    run   all tile passes
    reduce P(flag = 1) and |psi[2**(n-1) : 2**(n-1) + 2**nb]|^2   (HHL._calculate_norm)
    D2H   2**nb amplitudes (the flag = 1, clock = 0 slice the reference's
          extract_hhl_solution_vector_from_state_vector reads, utils.py:77-94)
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .circuit import QuantumCircuit
from .engine import DeviceProgram, DeviceState
from .lower import LoweredProgram, LoweringOptions, lower_circuit


class ShapedSolver:
    def __init__(self, circuit: QuantumCircuit, nb: int, precision: str = "fp64", device: Optional[int] = None,
                 options: Optional[LoweringOptions] = None):
        opts = options or LoweringOptions(dtype=precision)
        opts.dtype = precision
        self.circuit = circuit
        self.n = circuit.num_qubits
        self.nb = int(nb)
        self.lowered: LoweredProgram = lower_circuit(circuit, opts)
        self.program = DeviceProgram(self.lowered, device)
        self.state = DeviceState(self.n, precision, device)
        if not (self.lowered.steps and self.lowered.steps[0].kind == "init"):
            raise ValueError("circuit must start with a state preparation on the data register")
        self._init_step = self.lowered.steps[0]

    def run_device(self, b_device=None) -> None:
        """Evolve with the right-hand side already resident in HBM (``b_device``: complex tensor of
        ``2**nb`` amplitudes) or with the circuit's own state preparation."""
        self.state.init_basis(None)
        if b_device is not None:
            self.state.tensor[: 2**self.nb].copy_(b_device)
        else:
            self.state.write(0, self._init_step.amps)
        self._run_rest()

    def _run_rest(self) -> None:
        self.program.run(self.state, initialise=False)  # step 0 (state preparation) is done above

    def stats(self):
        low = self.lowered
        return {
            "fused_passes": low.n_passes,
            "fused_blocks": sum(len(p.op_controls) for p in low.pass_info),
            "source_gates": low.n_source_gates,
            "pass_bytes_per_gpu": low.total_pass_bytes(),
            "remaps": 0,
            "remap_bytes_sent_per_gpu": 0,
        }

    def time_passes(self, repeats: int = 1):
        """(ms, launches, algorithmic bytes) of the tile-pass kernels per solve, timed with CUDA
        events on the launching stream (the roofline numerator of bench.py)."""
        tot_ms, launches, nbytes = 0.0, 0, 0
        for _ in range(repeats):
            self.state.init_basis(None)
            self.state.write(0, self._init_step.amps)
            ms, launches, nbytes = self.program.run_passes_timed(self.state)
            tot_ms += ms
        return tot_ms / repeats, launches, nbytes

    def reduce(self) -> Tuple[float, float]:
        half = 2 ** (self.n - 1)
        # the flag is the top qubit: P(flag = 1) is the norm of the upper half -- a range reduction reads 64 GiB of a
        # 33-qubit state, the masked one (every amplitude, test the bit) all 128
        p_flag = self.state.norm_sq_range(half, half)
        norm_sq = self.state.norm_sq_range(half, 2**self.nb)
        return p_flag, norm_sq

    def solve(self, b_host: np.ndarray) -> Tuple[np.ndarray, float, float]:
        """Host in, host out: returns (solution slice, P(flag=1), |slice|^2)."""
        b = np.asarray(b_host).reshape(-1)
        if b.size != 2**self.nb:
            raise ValueError("right-hand side has the wrong length")
        self.state.init_basis(None)
        self.state.write(0, b)
        self._run_rest()
        p_flag, norm_sq = self.reduce()
        x = self.state.read(2 ** (self.n - 1), 2**self.nb)
        return x, p_flag, norm_sq
