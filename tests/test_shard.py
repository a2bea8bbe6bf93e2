"""not gpu: the N>1 path -- read groups sharded by query-name range over ranks, no collective on
the data path, host-side merge in rank order -- exercised with 2 gloo processes on the CPU.  Each
rank scores its shard with the host simulation of the kernels (tests/hostsim; the GPU-less stand-in
for the device) and rank 0 checks the merged result against a single run over all groups."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_groups():
    from secphase_b200.shard import shard_range
    for n in (0, 1, 7, 8, 4096, 10001):
        for world in (1, 2, 3, 4, 8):
            rs = [shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import pyoracle
    from secphase_b200.shard import merge_results, shard_range
    from tests.conftest import make_case
    from tests.hostsim import pyhostsim
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s, b, codes, off = make_case("hifi", 41, locus_len=200000)
        hp = pyhostsim.params_from_oracle(pyoracle.preset_params("hifi"))
        g0, g1 = shard_range(b.n_groups, rank, world)
        mine = pyhostsim.run(b.group_slice(g0, g1), hp, codes, off)
        part = {"groups": mine["groups"], "scores": mine["scores"], "extents": mine["extents"],
                "markers_final": mine["markers_final"], "markers_final_off": mine["markers_final_off"],
                "hmm_cells": mine["cells"]}
        gathered = [None] * world
        dist.all_gather_object(gathered, part)  # plumbing only: results travel, the data path has no exchange
        if rank == 0:
            merged = merge_results(gathered)
            whole = pyhostsim.run(b, hp, codes, off)
            ok = (np.array_equal(merged["scores"].view(np.int64), whole["scores"].view(np.int64))
                  and np.array_equal(merged["groups"], whole["groups"])
                  and np.array_equal(merged["extents"], whole["extents"])
                  and np.array_equal(merged["markers_final"], whole["markers_final"])
                  and np.array_equal(merged["markers_final_off"], whole["markers_final_off"])
                  and merged["hmm_cells"] == whole["cells"])
            q.put(bool(ok))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_run_equals_single_run():
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) is True
