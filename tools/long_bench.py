"""Long-pair runs.  Single GPU:  python tools/long_bench.py R Q      (dpx_align_long_pair, checks small cases vs the oracle)
Multi GPU (mode B):            torchrun --nproc-per-node N tools/long_bench.py R Q   (column stripes over NVLink P2P)"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from dpx_gpu_genomics_project_b200 import api, synth, longpair

R = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else R
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
# config 5 input: query = reference mutated 1% / 0.1% / 0.1%, seed 0x5EED0005
img = synth.mutated_fixed_file_bytes(1, R, Q, 0x5EED0005, 0.01, 0.001, 0.001)
ref = img[2:2 + R].tobytes(); qry = img[3 + R:3 + R + Q].tobytes()
eng = api.Engine(local)
p = api.make_params(api.LSW)
out = {"R": R, "Q": Q, "world": world}
if world == 1:
    for _ in range(reps):
        t0 = time.perf_counter(); res = eng.align_long_pair(p, ref, qry); dt = time.perf_counter() - t0
    out.update(result=res, seconds=dt, gcups=R * Q / dt / 1e9, mode="single GPU, dpx_align_long_pair (incl. H2D)")
res2 = None
try:
    job = longpair.StripedLongPair(eng, p, ref, qry, rank, world, dist)
    for _ in range(reps):
        res2, ms = job.run()
    out.update(striped_result=res2, striped_kernel_ms=ms, striped_gcups=R * Q / (ms * 1e-3) / 1e9, bounds=job.bounds)
    job.free()
except api.DpxError as e:
    if world > 1:
        raise
    out["striped_error"] = str(e)
if rank == 0:
    if R * Q <= 4e9:
        import oracle_lib as ol
        out["oracle"] = ol.lsw_score_only(ol.params(ol.LSW), ref, qry)
        out["match"] = (res2 is None or tuple(out["oracle"]) == tuple(res2)) and (world > 1 or tuple(res) == tuple(out["oracle"]))
    print(json.dumps(out))
if dist is not None:
    dist.destroy_process_group()
