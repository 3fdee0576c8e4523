"""Closed-form HHL final state (dense linear algebra in A's eigenbasis).

TEST INFRASTRUCTURE ONLY (see ``oracle/numpy_sv.py``).

Independent restatement of what ``linear_solvers.HHL.construct_circuit`` (quantum-linear-solvers
0.1.1, ``poetry.lock:3185-3186``; called from the reference at
``hhl_qiskit_implementation.py:35-38``) builds, without going through any gate stream:

* ``nb = log2 N``; ``nl = max(nb+1, ceil(log2(kappa+1))) + 1`` (``neg_vals=True``);
* ``delta = _get_delta(nl-1, lambda_min, lambda_max)``; ``t = 2*pi*delta/lambda_min/2``;
* QPE on the clock register = ``W_phi = R F^dagger diag(exp(2*pi*i*phi*x)) H^{(x)nl}`` for an
  eigenvector with eigenphase ``phi = lambda*t/(2*pi)`` (``R`` = bit reversal, ``F`` = DFT on the
  little-endian value), see SURVEY.md Appendix A;
* ``ExactReciprocal(nl, delta, neg_vals=True)`` rotates the flag by ``RY(theta(c))`` with ``c``
  the bit-reversed clock register value;
* final state ``sum_k beta_k |u_k> (x) (W_k^dagger (x) I_flag) Rot (W_k|0> (x) |0>_flag)``.

Used to check the gate-level builder + oracle (and through them the CUDA engine) on the toy
problems; the 2x2 case reproduces Appendix B of SURVEY.md.
"""
from __future__ import annotations

import math

import numpy as np


def get_delta(n_l: int, lambda_min: float, lambda_max: float) -> float:
    """``HHL._get_delta``: ``lambda_min`` scaled to ``n_l`` bits, truncated, as a binary fraction."""
    lam = abs(lambda_min * (2**n_l - 1) / lambda_max)
    if abs(lam - 1) < 1e-7:
        lam = 1
    bits = format(int(lam), "#0" + str(n_l + 2) + "b")[2:]
    return sum(int(ch) / 2 ** (i + 1) for i, ch in enumerate(bits))


def hhl_parameters(matrix: np.ndarray):
    n = matrix.shape[0]
    nb = int(np.log2(n))
    kappa = np.linalg.cond(matrix)
    nl = max(nb + 1, int(np.ceil(np.log2(kappa + 1)))) + 1
    ev = np.abs(np.linalg.eigvals(matrix))
    lmin, lmax = float(np.min(ev)), float(np.max(ev))
    delta = get_delta(nl - 1, lmin, lmax)
    t = 2 * np.pi * delta / lmin / 2
    return nb, nl, delta, t, lmin


def reciprocal_angle_by_value(nl: int, delta: float) -> np.ndarray:
    """Net flag rotation angle for every clock *value* ``c`` (sign bit = MSB of ``c``)."""
    m = 2 ** (nl - 1)
    out = np.zeros(2**nl)
    for c in range(2**nl):
        v, sign = c & (m - 1), c >> (nl - 1)
        if v == 0:
            continue
        # sign = 0: the value bits read lambda ~ v/m;  sign = 1 (two's complement): lambda ~ -(1 - v/m)
        lam = v / m if not sign else -(1.0 - v / m)
        r = delta / lam
        if math.isclose(abs(r), 1, abs_tol=1e-5):
            th = math.copysign(math.pi, r)
        elif abs(r) < 1:
            th = 2 * math.asin(r)
        else:
            th = 0.0
        out[c] = th
    return out


def _bitrev(x: int, nbits: int) -> int:
    return int(format(x, f"0{nbits}b")[::-1], 2) if nbits else 0


def _bitrev_perm(nbits: int) -> np.ndarray:
    idx = np.arange(2**nbits)
    rev = np.zeros_like(idx)
    for j in range(nbits):
        rev |= ((idx >> j) & 1) << (nbits - 1 - j)
    return rev


def _wht(a: np.ndarray) -> np.ndarray:
    """Unnormalised Walsh-Hadamard transform along axis 0 (natural order: H^{(x)m}[x][y] = (-1)^{popcount(x & y)})."""
    a = a.copy()
    n = a.shape[0]
    h = 1
    while h < n:
        v = a.reshape(n // (2 * h), 2, h, -1)
        lo, hi = v[:, 0].copy(), v[:, 1].copy()
        v[:, 0] = lo + hi
        v[:, 1] = lo - hi
        h *= 2
    return a


def eigenphase_table(w_eig: np.ndarray, t: float, dim_c: int) -> np.ndarray:
    """``exp(2*pi*i*phi_k*x)`` for every clock value ``x`` and eigenvalue ``k`` (``phi_k = lambda_k t / 2 pi``).
    ``phi_k * x`` reaches ``2**(nl-1)`` turns, which costs ``nl - 1`` bits of a double: the turn count is
    reduced modulo 1 in extended precision, so that the table is good to ~1e-16 at every ``x`` (the circuit
    the oracle describes applies ``U**(2**j)`` exactly)."""
    ld = np.longdouble
    two_pi = 2 * np.arccos(ld(-1))
    phi = w_eig.astype(ld) * ld(t) / two_pi
    turns = np.outer(np.arange(dim_c).astype(ld), phi)
    turns -= np.floor(turns)
    ang = (two_pi * turns).astype(np.float64)
    return np.cos(ang) + 1j * np.sin(ang)


def hhl_state(matrix: np.ndarray, vector: np.ndarray) -> np.ndarray:
    """Final ``2**(nb+nl+1)`` amplitude vector of the HHL circuit (register order b, clock, flag).

    Per eigenpair ``(lambda_k, u_k)`` the clock/flag register ends in ``(W_k^dagger (x) I) Rot (W_k|0> (x) |0>)`` with
    ``W_phi = R F^dagger diag(exp(2 pi i phi x)) H^{(x)nl}``; ``F`` is applied as an FFT and ``H^{(x)nl}`` as a
    Walsh-Hadamard transform, for all ``k`` at once, so that the N = 2**10, nl = 12 case of BASELINE.json configs[1]
    (2**23 amplitudes) is a few seconds of host time.  The eigen-decomposition is taken of the matrix AS GIVEN
    (``np.linalg.eigh`` on a real array for a real matrix): eigenvalue bits differ between LAPACK's real and complex
    drivers at the 1e-16 level, and ``U**(2**(nl-1))`` amplifies that by ``2**(nl-1)``."""
    a = np.asarray(matrix)
    nb, nl, delta, t, _ = hhl_parameters(a)
    dim_c = 2**nl
    w_eig, v_eig = np.linalg.eigh(a)
    b = np.asarray(vector, dtype=complex).reshape(-1)
    b = b / np.linalg.norm(b)
    beta = v_eig.conj().T @ b
    rev = _bitrev_perm(nl)
    phase = eigenphase_table(w_eig, t, dim_c)                    # [x, k]
    w0 = np.empty_like(phase)
    w0[rev] = np.fft.fft(phase, axis=0) / dim_c                   # W_k|0> = R F^dagger (phase / sqrt(D)), [register, k]
    theta_by_reg = reciprocal_angle_by_value(nl, delta)[rev]      # register index r holds value rev(r)
    out = []
    for rot in (np.cos(theta_by_reg / 2), np.sin(theta_by_reg / 2)):
        chi = rot[:, None] * w0
        z = chi[rev]                                              # R^T
        z = np.fft.ifft(z, axis=0) * math.sqrt(dim_c)             # F
        z *= phase.conj()
        z = _wht(z) / math.sqrt(dim_c)                            # H^{(x)nl}
        out.append(z @ (beta[:, None] * v_eig.T))                 # sum_k beta_k (W_k^dagger chi_k)[c] u_k[b]
    return np.stack(out).reshape(-1)


def hhl_solution(matrix: np.ndarray, vector: np.ndarray) -> np.ndarray:
    """The reference's post-processing (``hhl_qiskit_implementation.py:49-54``) on the closed-form
    state: slice, real part, normalise, multiply by ``euclidean_norm``."""
    psi = hhl_state(matrix, vector)
    n = matrix.shape[0]
    _, _, _, _, lmin = hhl_parameters(np.asarray(matrix))
    half = psi.size // 2
    sl = psi[half : half + n]
    v = np.real(sl)
    return np.linalg.norm(sl) / lmin * v / np.linalg.norm(v)
