"""Sharded CUDA path on >= 2 GPUs (skipped on a single-GPU box): one process per GPU, NCCL
send/recv remaps, result compared with the oracle and with the single-GPU engine."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, out_dir, p2p):
    import torch
    import torch.distributed as dist

    os.environ["QLS_REMAP_P2P"] = "1" if p2p else "0"  # remaps through peer memory (CUDA IPC) or through NCCL send/recv

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device(f"cuda:{rank}"))
    try:
        from qls_b200.shard import ShardedShapedSolver
        from qls_b200.synthetic import hhl_shaped_circuit

        qc = hhl_shaped_circuit(n)
        nb = qc.qregs[0].size
        solver = ShardedShapedSolver(qc, nb, device=rank, chunk_amps=1 << 14)
        rng = np.random.default_rng(99)
        b = rng.standard_normal(2**nb) + 0j
        b /= np.linalg.norm(b)
        x, p_flag, norm_sq = solver.solve(b)
        shard = solver.state.read()
        np.save(os.path.join(out_dir, f"shard{rank}.npy"), shard)
        if rank == 0:
            np.save(os.path.join(out_dir, "x.npy"), x)
            np.save(os.path.join(out_dir, "scalars.npy"), np.array([p_flag, norm_sq]))
            np.save(os.path.join(out_dir, "layout.npy"), np.array(solver.plan.final_layout))
            dq = sorted(solver.plan.deferred)
            np.save(os.path.join(out_dir, "deferred_q.npy"), np.array(dq, dtype=np.int64))
            np.save(os.path.join(out_dir, "deferred_m.npy"), np.array([solver.plan.deferred[q] for q in dq], dtype=complex).reshape(-1, 2, 2))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("p2p", [False, True])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_hhl_shaped_matches_oracle(cuda_device, tmp_path, world, p2p):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from oracle import numpy_sv

    from qls_b200.synthetic import hhl_shaped_circuit

    from _util import circuit_to_stream

    n = 19
    port = 29600 + (os.getpid() + world + 7 * p2p) % 2000
    mp.spawn(_worker, args=(world, port, n, str(tmp_path), p2p), nprocs=world, join=True)
    qc = hhl_shaped_circuit(n)
    nb = qc.qregs[0].size
    want = numpy_sv.run(n, circuit_to_stream(qc))
    layout = list(np.load(tmp_path / "layout.npy"))
    phys = np.concatenate([np.load(tmp_path / f"shard{r}.npy") for r in range(world)])
    idx = np.arange(2**n, dtype=np.int64)
    p = np.zeros_like(idx)
    for q in range(n):
        p |= ((idx >> q) & 1) << layout[q]
    got = phys[p]
    # the solver path leaves the final gate of clock qubits that end on a global position to the slice reduction
    # (ShardedPlan.deferred): the shards hold the state BEFORE those gates -- apply them here
    dq, dm = np.load(tmp_path / "deferred_q.npy"), np.load(tmp_path / "deferred_m.npy")
    if world > 2:
        assert len(dq) > 0
    for q, m in zip(dq, dm):
        got = numpy_sv.evolve_operator(got, m, [int(q)])
    assert np.max(np.abs(got - want)) <= 1e-12
    half = 2 ** (n - 1)
    x = np.load(tmp_path / "x.npy")
    assert np.max(np.abs(x - want[half: half + 2**nb])) <= 1e-12
    p_flag, norm_sq = np.load(tmp_path / "scalars.npy")
    assert abs(p_flag - numpy_sv.prob_mask(want, half, half)) < 1e-12
    assert abs(norm_sq - np.sum(np.abs(want[half: half + 2**nb]) ** 2)) < 1e-13


def _overlap_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device(f"cuda:{rank}"))
    try:
        import bench

        res = bench.sharded_parity_check(world, rank, rank, "fp64", n_local=22, chunk_amps=1 << 15)
        if rank == 0:
            import json

            with open(os.path.join(out_dir, "parity.json"), "w") as fh:
                json.dump(res, fh)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_exchanges_overlapped_with_passes(cuda_device, tmp_path, world):
    """At 22 local qubits the passes run on the specialised kernels and the exchanges split into halves: the passes on both
    sides of an exchange run half a shard at a time while the other half is on the wire (ShardedShapedSolver._run_segments).
    Sharded vs single-GPU solver on the same right-hand side, 1e-12; at least one exchange must actually be overlapped."""
    import json

    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29700 + (os.getpid() + world) % 2000
    mp.spawn(_overlap_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = json.load(open(tmp_path / "parity.json"))
    assert res["shard_parity_max_abs"] <= 1e-12
    assert res["exchanges_overlapped_with_passes"], res
