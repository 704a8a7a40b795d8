"""Two-GPU runs (skipped on a single-GPU box): mode A (sharded batch, no collective on the data path) and mode B
(one long pair in column stripes, NVLink P2P boundary exchange through CUDA-IPC mapped inboxes)."""
import os
import socket

import numpy as np
import pytest

import oracle_lib as ol
from dpx_gpu_genomics_project_b200 import sharding, synth

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _pairs():
    rng = synth.Rng(123)
    pp = []
    for k in range(3000):
        R = 1 + int(rng.below(1, 200)[0])
        r = synth.random_seq(rng, R)
        q = synth.mutate(rng, r, 0.05, 0.02, 0.02) if k % 2 else synth.random_seq(rng, 1 + int(rng.below(1, 200)[0]))
        pp.append((r, q))
    return ol.parse_image(synth.pairs_to_file_bytes(pp))


def _long_inputs():
    img = synth.mutated_fixed_file_bytes(1, 30000, 26000, 0x5EED0005, 0.01, 0.001, 0.001)
    return img[2:2 + 30000].tobytes(), img[3 + 30000:3 + 30000 + 26000].tobytes()


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    from dpx_gpu_genomics_project_b200 import api, longpair
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    eng = api.Engine(rank)
    # mode A
    blob, pairs = _pairs()
    p = api.make_params(api.LSW, flags=api.OUT_SCORE | api.OUT_END_COORDS)

    def compute(seqs, sub):
        r = eng.align_batch(p, seqs, sub)
        return r.scores, r.end_row_col
    s, e = sharding.align_sharded(compute, blob, pairs, rank, world, dist)
    # mode B
    ref, qry = _long_inputs()
    job = longpair.StripedLongPair(eng, api.make_params(api.LSW), ref, qry, rank, world, dist)
    res1, _ = job.run()
    res2, ms = job.run()          # second run exercises reset()
    job.free()
    # every lane width the stripes can take (the exported right edge must be the stripe's last REAL column)
    wide = []
    for k in (32, 16, 8, 2):
        with eng.options(long_k=k):
            job = longpair.StripedLongPair(eng, api.make_params(api.LSW), ref, qry, rank, world, dist)
            wide.append(job.run()[0])
            job.free()
    if rank == 0:
        np.savez(out_path, s=s, e=e, long1=np.array(res1), long2=np.array(res2), wide=np.array(wide))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("world", sorted({2, _ngpu()} - {0, 1}) or [2])
def test_multi_gpu_modes_match_oracle(tmp_path, world):
    """Mode A (sharded batch) and mode B (striped long pair) on 2 GPUs and on every GPU of the box: identical bytes at every
    rank count (test-plan item 4)."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    blob, pairs = _pairs()
    s, e, _ = ol.align_batch(ol.params(ol.LSW), blob, pairs, strings=False, threads=8)
    assert (got["s"] == s).all() and (got["e"] == e).all()
    ref, qry = _long_inputs()
    want = ol.lsw_score_only(ol.params(ol.LSW), ref, qry)
    assert tuple(got["long1"]) == want and tuple(got["long2"]) == want
    assert all(tuple(w) == want for w in got["wide"])
