"""NumPy interpreter of the binary program format of ``include/qls_abi.h`` (test infrastructure).

Executes ``QlsPass`` / ``QlsOp`` records with the semantics the CUDA kernels implement
(``csrc/tile_pass.cu``, ``csrc/tile_ops.cuh``, ``csrc/ctrl_gemm.cu``), so that the lowering pass
can be validated against the oracle in the GPU-less container.  It is NOT a product code path:
nothing in the package imports it.
"""
import numpy as np

from qls_b200 import _abi


def _gather(x, segs, n):
    r = np.zeros_like(x)
    for i in range(n):
        s = segs[i]
        r |= ((x >> s.src) & ((1 << s.width) - 1)) << s.dst
    return r


def _insert_zero(x, pos):
    low = x & ((1 << pos) - 1)
    return ((x >> pos) << (pos + 1)) | low


def _swz(i, sh):
    u = i >> sh
    return i ^ ((((u >> 3) ^ (u >> 6) ^ (u >> 9)) & 7) << sh)


def run_pass_rs(state, p, rs_ops, pool, cdtype, n_local, index_offset=0):
    """Register-stage form (csrc/rs_pass.cu): the tile lives in a swizzled shared-memory image and
    is re-read per stage as a[thread][register combo]."""
    T, L = p.tile_bits, p.low_bits
    sh = 0 if np.dtype(cdtype).itemsize == 16 else 1
    KR = 5 if (rs_ops[p.rs_begin].flags & 0x20) else 4  # STAGE flag: five register qubits
    A = 2**KR
    NT = 2 ** (T - KR)
    fixed = bin(p.fix_mask).count("1")
    tiles = np.arange(2 ** (n_local - T - fixed), dtype=np.int64)
    base = _gather(tiles, p.fseg, p.n_fseg) | p.fix_val
    i = np.arange(2**T, dtype=np.int64)
    spread = i & ((1 << L) - 1)
    rr = i >> L
    for j in range(p.n_high):
        spread |= ((rr >> j) & 1) << p.high[j]
    G = base[:, None] | spread[None, :]
    sm = np.empty((len(tiles), 2**T), dtype=cdtype)
    sm[:, _swz(i, sh)] = state[G]  # swizzled shared-memory image
    gbase = base | index_offset
    t = np.arange(NT, dtype=np.int64)
    r = np.arange(A, dtype=np.int64)
    a = None
    addr = None

    def tab_len(op):
        mask = 0
        for sg in list(op.tseg)[: op.n_tseg] + list(op.xseg)[: op.n_xseg]:
            mask |= ((1 << sg.width) - 1) << sg.dst
        for k in range(A):
            mask |= int(op.rt[k])
        return mask + 1

    for o in range(p.rs_begin, p.rs_begin + p.rs_count):
        op = rs_ops[o]
        if op.kind == _abi.QLS_RS_STAGE:
            if a is not None:
                sm[:, addr.reshape(-1)] = a.reshape(len(tiles), -1)
            tb = _swz(_gather(t, op.tseg, op.n_tseg), sh)
            rt = np.array([op.rt[k] for k in range(A)], dtype=np.int64)
            addr = tb[:, None] ^ rt[None, :]
            assert len(np.unique(addr)) == 2**T, "stage layout is not a permutation of the tile"
            a = sm[:, addr.reshape(-1)].reshape(len(tiles), NT, A)
            continue
        act = (gbase & op.xc_mask) == op.xc_val
        tok = (t & op.ctl_tmask) == op.ctl_tval
        rok = (r & op.ctl_rmask) == op.ctl_rval
        if op.kind in (_abi.QLS_RS_G1, _abi.QLS_RS_G2):
            k = 1 if op.kind == _abi.QLS_RS_G1 else 2
            M = np.frombuffer(pool, dtype=cdtype, count=4**k, offset=op.pool_off).reshape(2**k, 2**k)
            regs = [op.r[j] for j in range(k)]
            g = np.arange(A >> k, dtype=np.int64)
            i0 = g.copy()
            for q in regs:
                i0 = _insert_zero(i0, q)
            offs = np.array([sum(((c >> j) & 1) << regs[j] for j in range(k)) for c in range(2**k)])
            cols = i0[:, None] | offs[None, :]
            sub = a[:, :, cols]  # tiles x t x groups x 2^k
            res = np.einsum("rc,xtgc->xtgr", M, sub)
            ok = act[:, None, None, None] & tok[None, :, None, None] & rok[i0][None, None, :, None]
            a[:, :, cols] = np.where(ok, res, sub).astype(cdtype)
        elif op.kind == _abi.QLS_RS_DIAG:
            table = np.frombuffer(pool, dtype=cdtype, count=tab_len(op), offset=op.pool_off)
            rt = np.array([op.rt[k] for k in range(A)], dtype=np.int64)
            ti = (_gather(gbase, op.xseg, op.n_xseg)[:, None, None] | _gather(t, op.tseg, op.n_tseg)[None, :, None]
                  | rt[None, None, :])
            ok = act[:, None, None] & tok[None, :, None] & rok[None, None, :]
            a = np.where(ok, a * table[ti], a).astype(cdtype)
        elif op.kind == _abi.QLS_RS_MUXRY:
            table = np.frombuffer(pool, dtype=cdtype, count=tab_len(op), offset=op.pool_off)
            rt = np.array([op.rt[k] for k in range(A)], dtype=np.int64)
            q = op.r[0]
            i0 = _insert_zero(np.arange(A // 2, dtype=np.int64), q)
            i1 = i0 | (1 << q)
            ti = (_gather(gbase, op.xseg, op.n_xseg)[:, None, None] | _gather(t, op.tseg, op.n_tseg)[None, :, None]
                  | rt[i0][None, None, :])
            cs = table[ti]
            x, y = a[:, :, i0], a[:, :, i1]
            ok = act[:, None, None] & tok[None, :, None] & rok[i0][None, None, :]
            a[:, :, i0] = np.where(ok, cs.real * x - cs.imag * y, x).astype(cdtype)
            a[:, :, i1] = np.where(ok, cs.imag * x + cs.real * y, y).astype(cdtype)
        else:
            raise ValueError(op.kind)
    sm[:, addr.reshape(-1)] = a.reshape(len(tiles), -1)
    state[G] = sm[:, _swz(i, sh)]


def run_pass(state, p, ops, pool, cdtype, n_local, index_offset=0, params=None):
    T, L = p.tile_bits, p.low_bits
    fixed = bin(p.fix_mask).count("1")
    free_bits = n_local - T - fixed
    tiles = np.arange(2**free_bits, dtype=np.int64)
    base = _gather(tiles, p.fseg, p.n_fseg) | p.fix_val
    i = np.arange(2**T, dtype=np.int64)
    spread = i & ((1 << L) - 1)
    r = i >> L
    for j in range(p.n_high):
        spread |= ((r >> j) & 1) << p.high[j]
    G = base[:, None] | spread[None, :]
    sm = state[G]  # tiles x 2^T
    gbase = base | index_offset
    esz = np.dtype(cdtype).itemsize
    ATT = _abi.QLS_OP_FLAG_ATTACH_PRE | _abi.QLS_OP_FLAG_ATTACH_POST
    order = []  # execution order: PRE attachments, host, POST attachments
    o = p.op_begin
    end = p.op_begin + p.op_count
    while o < end:
        assert not (ops[o].flags & ATT), "attachment without a host"
        att = []
        while len(att) < _abi.QLS_MAX_ATTACH and o + 1 + len(att) < end and (ops[o + 1 + len(att)].flags & ATT):
            att.append(o + 1 + len(att))
        if att:
            host = ops[o]
            assert host.kind in (_abi.QLS_OP_DENSE, _abi.QLS_OP_DIAG) and host.lc_mask == 0
            assert all(ops[a].kind == _abi.QLS_OP_DIAG and ops[a].lc_mask == 0 and ops[a].xc_mask == 0 for a in att)
            if host.kind == _abi.QLS_OP_DIAG:
                assert all(ops[a].flags & _abi.QLS_OP_FLAG_ATTACH_POST for a in att)
        order += [a for a in att if ops[a].flags & _abi.QLS_OP_FLAG_ATTACH_PRE] + [o]
        order += [a for a in att if ops[a].flags & _abi.QLS_OP_FLAG_ATTACH_POST]
        o += 1 + len(att)
    for o in order:
        op = ops[o]
        act = (gbase & op.xc_mask) == op.xc_val
        if not act.any():
            continue
        if op.kind == _abi.QLS_OP_MATVEC:
            K, lo = op.k, op.tq[0]
            U = np.frombuffer(pool, dtype=cdtype, count=4**K, offset=op.pool_off).reshape(2**K, 2**K)
            outer = np.arange(2 ** (T - K), dtype=np.int64)
            base_o = ((outer >> lo) << (lo + K)) | (outer & ((1 << lo) - 1))
            base_o = base_o[(base_o & op.lc_mask) == op.lc_val]
            cols = base_o[:, None] | (np.arange(2**K, dtype=np.int64) << lo)[None, :]
            sub = sm[:, cols]  # tiles x outer x 2^K
            sm[:, cols] = np.einsum("rc,toc->tor", U, sub).astype(cdtype)
        elif op.kind == _abi.QLS_OP_DENSE:
            K = op.k
            if op.flags & _abi.QLS_OP_FLAG_PARAM:
                ang = float(np.float32(op.param_coeff)) * params[op.param_index] + float(np.float32(op.param_offset))
                c_, s_ = np.cos(ang / 2), np.sin(ang / 2)
                M = {0: np.array([[c_, -s_], [s_, c_]], dtype=complex),
                     1: np.array([[c_, -1j * s_], [-1j * s_, c_]]),
                     2: np.diag([np.exp(-0.5j * ang), np.exp(0.5j * ang)]),
                     3: np.diag([1, np.exp(1j * ang)])}[op.pad0].astype(cdtype)
            else:
                M = np.frombuffer(pool, dtype=cdtype, count=4**K, offset=op.pool_off).reshape(2**K, 2**K)
            g = np.arange(2 ** (T - K), dtype=np.int64)
            idx = g.copy()
            for j in range(K):
                idx = _insert_zero(idx, op.tq[j])
            ok = (idx & op.lc_mask) == op.lc_val
            idx = idx[ok]
            offs = np.zeros(2**K, dtype=np.int64)
            for c in range(2**K):
                for j in range(K):
                    if (c >> j) & 1:
                        offs[c] |= 1 << op.tq[j]
            cols = idx[:, None] | offs[None, :]  # groups x 2^K
            sub = sm[np.ix_(act.nonzero()[0], cols.reshape(-1))].reshape(-1, len(idx), 2**K)
            res = np.einsum("rc,tgc->tgr", M, sub)
            sm[np.ix_(act.nonzero()[0], cols.reshape(-1))] = res.reshape(res.shape[0], -1)
        elif op.kind == _abi.QLS_OP_DIAG:
            n_tab = 1
            for s in list(op.lseg)[: op.n_lseg] + list(op.xseg)[: op.n_xseg]:
                n_tab = max(n_tab, 1 << (s.dst + s.width))
            table = np.frombuffer(pool, dtype=cdtype, count=n_tab, offset=op.pool_off)
            ti = _gather(gbase, op.xseg, op.n_xseg)[:, None] | _gather(i, op.lseg, op.n_lseg)[None, :]
            ok = ((i & op.lc_mask) == op.lc_val)[None, :] & act[:, None]
            fac = np.where(ok, table[ti], 1.0)
            sm = (sm * fac).astype(cdtype)
        elif op.kind == _abi.QLS_OP_MUXRY:
            n_tab = 1
            for s in list(op.lseg)[: op.n_lseg] + list(op.xseg)[: op.n_xseg]:
                n_tab = max(n_tab, 1 << (s.dst + s.width))
            table = np.frombuffer(pool, dtype=cdtype, count=n_tab, offset=op.pool_off)
            tq = op.tq[0]
            g = np.arange(2 ** (T - 1), dtype=np.int64)
            i0 = _insert_zero(g, tq)
            ok = (i0 & op.lc_mask) == op.lc_val
            i0 = i0[ok]
            i1 = i0 | (1 << tq)
            ti = _gather(gbase, op.xseg, op.n_xseg)[:, None] | _gather(i0, op.lseg, op.n_lseg)[None, :]
            cs = table[ti]
            c, s_ = cs.real, cs.imag
            a0, a1 = sm[:, i0], sm[:, i1]
            b0 = c * a0 - s_ * a1
            b1 = s_ * a0 + c * a1
            m = act[:, None]
            sm[:, i0] = np.where(m, b0, a0)
            sm[:, i1] = np.where(m, b1, a1)
        else:
            raise ValueError(op.kind)
    state[G] = sm
    _ = esz


def run_program(lowered, psi0=None, index_offset=0, params=None, form="sweep"):
    """Evolve |0...0> (or ``psi0``) through a LoweredProgram; returns the final local state."""
    n = lowered.n_local
    cdtype = np.complex128 if lowered.dtype == "fp64" else np.complex64
    pool = lowered.pool.tobytes()
    state = np.zeros(2**n, dtype=cdtype)
    steps = list(lowered.steps)
    if psi0 is not None:
        state[:] = psi0
    elif steps and steps[0].kind == "init":
        s = steps.pop(0)
        assert tuple(s.qubits) == tuple(range(len(s.qubits)))
        rank = index_offset >> n
        f = (1.0 if rank == 0 else 0.0) if s.rank_factors is None else s.rank_factors[rank]
        state[: len(s.amps)] = np.asarray(s.amps) * f
    elif index_offset == 0:
        state[0] = 1
    for s in steps:
        if s.kind == "passes":
            for pi in range(s.pass_begin, s.pass_end):
                ps = lowered.passes[pi]
                if form == "jit" and ps.rs_count > 0:
                    # the generated (specialised) source, compiled for the host: tests/_jit_emu.py
                    import _jit_emu

                    if not _jit_emu.run_pass(state, lowered, pi, n, index_offset):
                        run_pass_rs(state, ps, lowered.rs_ops, pool, cdtype, n, index_offset)
                elif form in ("rs", "jit") and ps.rs_count > 0:
                    run_pass_rs(state, ps, lowered.rs_ops, pool, cdtype, n, index_offset)
                else:
                    run_pass(state, ps, lowered.ops, pool, cdtype, n, index_offset, params)
        elif s.kind == "gemm":
            nb = s.nb
            rows = np.arange(2 ** (n - nb), dtype=np.int64)
            gidx = (rows << nb) | index_offset
            sel = (gidx & s.ctrl_mask) == s.ctrl_val
            mat = state.reshape(-1, 2**nb)
            mat[sel] = (mat[sel] @ s.matrix.T).astype(cdtype)
        else:
            raise ValueError(s.kind)
    return state
