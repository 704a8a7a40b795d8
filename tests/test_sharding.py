"""Multi-GPU mode A host logic on CPU: deterministic length-bucketed dealing, and a real world_size-2 gloo run in
which every rank aligns its shard (the CPU oracle stands in for the device) and rank 0 reassembles pair order."""
import os
import socket

import numpy as np
import pytest

import oracle_lib as ol
from dpx_gpu_genomics_project_b200 import sharding, synth


def _pairs(n=400, seed=11):
    rng = synth.Rng(seed)
    pp = []
    for k in range(n):
        R = 1 + int(rng.below(1, 120)[0])
        r = synth.random_seq(rng, R)
        q = synth.mutate(rng, r, 0.05, 0.02, 0.02) if k % 2 else synth.random_seq(rng, 1 + int(rng.below(1, 120)[0]))
        pp.append((r, q))
    return ol.parse_image(synth.pairs_to_file_bytes(pp))


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shards_are_a_balanced_partition(world):
    blob, pairs = _pairs()
    shards = sharding.shard_indices(pairs, world)
    allidx = np.concatenate(shards)
    assert len(allidx) == len(pairs) and len(np.unique(allidx)) == len(pairs)
    assert all((np.diff(s) > 0).all() for s in shards if len(s) > 1)
    cells = sharding.shard_cells(pairs, shards)
    assert cells.max() - cells.min() <= 2 * int((pairs["referenceSize"].astype(np.int64) * pairs["querySize"]).max())
    again = sharding.shard_indices(pairs, world)
    assert all((a == b).all() for a, b in zip(shards, again))


def _oracle_compute(seqs, sub):
    s, e, _ = ol.align_batch(ol.params(ol.LSW), seqs, sub, strings=False)
    return s, e


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blob, pairs = _pairs()
    s, e = sharding.align_sharded(_oracle_compute, blob, pairs, rank, world, dist)
    if rank == 0:
        np.savez(out_path, s=s, e=e)
    else:
        assert s is None and e is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_run_reassembles_pair_order(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    blob, pairs = _pairs()
    s, e = _oracle_compute(blob, pairs)
    assert (got["s"] == s).all() and (got["e"] == e).all()


# ---- mode B host logic (column stripes of one long pair) ---------------------------------------------------------
def test_stripe_bounds_cover_the_reference_in_whole_warp_blocks():
    from dpx_gpu_genomics_project_b200.longpair import stripe_bounds
    for R in (1, 511, 512, 513, 1023, 1025, 40000, 1_000_000, 999_999):
        for n in (1, 2, 3, 4, 8):
            b = stripe_bounds(R, n)
            assert b[0] == 0 and b[-1] == R and len(b) == n + 1
            assert all(b[i] <= b[i + 1] for i in range(n))
            # every stripe that is followed by a non-empty stripe ends on a multiple of 1024 columns (whole warps at K <= 32)
            assert all(b[i + 1] % 1024 == 0 for i in range(n - 1) if b[i + 2] > b[i + 1])


def test_stripe_result_reduction_follows_the_reference_end_cell_rule():
    from dpx_gpu_genomics_project_b200.longpair import reduce_results
    # higher score wins; ties: smaller row, then smaller column (c++/LinearSmithWaterman.cpp:145-157); score 0 never wins
    assert reduce_results([(5, 10, 3), (7, 50, 900), None, (7, 40, 1200)]) == (7, 40, 1200)
    assert reduce_results([(7, 40, 1200), (7, 40, 900)]) == (7, 40, 900)
    assert reduce_results([(0, 0, 0), None]) == (0, 0, 0)
