"""world_size-2 gloo tests (CPU) of the N > 1 host logic: the reference's cyclic column deal
(scale/common/common_mpi_scale.f90:264-283, 1428-1440) used by bench.py to shard the analysis,
and the member<->grid all-to-all transposes (:1279-1423) of scale_letkf_b200.transpose."""
import socket

import numpy as np
import pytest

from scale_letkf_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("nmem", [2, 3, 5])   # 3, 5: last round has fewer members than ranks (ALLTOALLV)
def test_transposes_world2_gloo(nmem):
    import torch.multiprocessing as mp
    import mr_worker
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=mr_worker.run, args=(r, 2, port, nmem, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_transposes_with_fused_state_trans_world2_gloo():
    """EnsTranspose(thermo=...): restart variables in, LETKF variables on the dealt columns, restart variables back."""
    import torch.multiprocessing as mp
    import mr_worker
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=mr_worker.run_thermo, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


@pytest.mark.parametrize("world,nmem", [(2, 3), (4, 6), (4, 8)])   # (4, 6): ranks 2, 3 hold no member in the last round
def test_p2p_handle_exchange_is_rank_consistent_gloo(world, nmem):
    """EnsTransposeP2P exchanges CUDA IPC handles with a collective; every rank must take part in every exchange even
    when its own grid of the round is None (50 members on 4 or 8 ranks) -- otherwise the cycle leg of bench.py hangs."""
    import torch.multiprocessing as mp
    import mr_worker
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=mr_worker.run_p2p_bookkeeping, args=(r, world, port, nmem, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=150) for _ in procs]
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.terminate()
    assert sorted(res) == [(r, "ok") for r in range(world)], res


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_column_deal_partitions_the_plane(world):
    """every (ilon, ilat) column is analysed by exactly one rank; nij1 follows set_common_mpi_grid"""
    nlon, nlat = 13, 7
    seen = np.zeros((nlon, nlat), dtype=int)
    sizes = []
    for r in range(world):
        ilon, ilat = synth.column_deal(nlon, nlat, world, r)
        seen[ilon - 1, ilat - 1] += 1
        sizes.append(len(ilon))
    assert (seen == 1).all()
    tot = nlon * nlat
    i = tot % world
    mx = (tot - i) // world + 1
    assert sizes == [mx if r < i else mx - 1 for r in range(world)]


@pytest.mark.gpu
@pytest.mark.parametrize("k", [6, 5])   # 5: rank 1 holds no member in the last round
def test_cycle_two_gpus_nccl(k):
    """2 GPUs: NCCL all-to-all transposes + per-rank das_letkf == single-domain oracle (<= 1e-10)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    import mr_worker
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=mr_worker.run_gpu, args=(r, 2, port, q, k)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
