"""Host-side cost of dpx_multi_align_batch with W worker threads (all on GPU 0 when the box has fewer GPUs): the kernels of 1M pairs
take ~3.3 ms in total however they are split, so anything above that is the per-call host path (driver calls serialising between threads).
usage: python tools/multi_trace.py [workers] [pairs]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dpx_gpu_genomics_project_b200 import api, synth
W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
inp = api.parse_image_native(synth.uniform_file_bytes(n, 150, 150, 0x5EED0002))
sc = torch.empty(n, dtype=torch.int32).pin_memory(); rc = torch.empty((n, 2), dtype=torch.int32).pin_memory()
p = api.make_params(api.LSW, flags=3)
ng = torch.cuda.device_count()
m = api.MultiEngine(devices=[d % ng for d in range(W)])
def once():
    st = m.L.dpx_multi_align_batch(m.h, C.byref(p), inp.sequences.ctypes.data, inp.sequences.size, inp.pairs.ctypes.data, n, sc.numpy().ctypes.data, rc.numpy().ctypes.data, None, None)
    assert st == 0
for ch in (12, 6, 4, 3, 2, 1):
    m.set_option("chunks_packed", ch)
    for _ in range(2): once()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); once(); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{W} workers on {ng} GPU(s), {n} pairs, chunks_packed {ch}: ms per call", [round(t, 2) for t in ts])
