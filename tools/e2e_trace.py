"""Prints the host-side timeline of one dpx_align_batch call (option "trace") for the bench workload."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dpx_gpu_genomics_project_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
blob, pairs = synth.uniform_blob_pairs(n, 150, 150, 0x5EED0002)
pb = torch.from_numpy(blob).pin_memory(); pp = torch.from_numpy(pairs.view(np.int32)).pin_memory()
sc = torch.empty(n, dtype=torch.int32).pin_memory(); rc = torch.empty((n, 2), dtype=torch.int32).pin_memory()
eng = api.Engine(0); p = api.make_params(api.LSW, flags=3)
def once():
    st = eng.L.dpx_align_batch(eng.ctx, C.byref(p), pb.numpy().ctypes.data, pb.numel(), pp.numpy().ctypes.data, n, sc.numpy().ctypes.data, rc.numpy().ctypes.data, None, None)
    assert st == 0
for _ in range(3): once()
ts = []
for _ in range(5):
    t0 = time.perf_counter(); once(); ts.append((time.perf_counter() - t0) * 1e3)
print("ms per call:", [round(t, 2) for t in ts])
eng.set_option("trace", 1)
once()
