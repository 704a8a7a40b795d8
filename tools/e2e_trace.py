"""Host-side timeline of one dpx_align_batch call (option "trace") for the bench workload, raw-byte and packed-sidecar input,
and the call time as a function of the number of chunks.   usage: python tools/e2e_trace.py [pairs] [devices]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dpx_gpu_genomics_project_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ndev = int(sys.argv[2]) if len(sys.argv) > 2 else 1
inp = api.parse_image_native(synth.uniform_file_bytes(n, 150, 150, 0x5EED0002))
sc = torch.empty(n, dtype=torch.int32).pin_memory(); rc = torch.empty((n, 2), dtype=torch.int32).pin_memory()
eng = api.Engine(0); p = api.make_params(api.LSW, flags=3)
def once():
    st = eng.L.dpx_align_batch(eng.ctx, C.byref(p), inp.sequences.ctypes.data, inp.sequences.size, inp.pairs.ctypes.data, n, sc.numpy().ctypes.data, rc.numpy().ctypes.data, None, None)
    assert st == 0
def timed(k=5):
    for _ in range(2): once()
    ts = []
    for _ in range(k):
        t0 = time.perf_counter(); once(); ts.append((time.perf_counter() - t0) * 1e3)
    return [round(t, 2) for t in ts]
for ch in (2, 3, 4, 6, 8, 12):
    eng.set_option("chunks_packed", ch)
    print("sidecar, chunks_packed", ch, "ms per call:", timed())
eng.set_option("chunks_packed", 6)
eng.set_option("trace", 1); once(); eng.set_option("trace", 0)
eng.set_option("no_sidecar", 1)
print("raw bytes (pageable host memory) ms per call:", timed(3))
eng.set_option("no_sidecar", 0)
if ndev > 1 or len(sys.argv) > 2:
    m = api.MultiEngine(devices=[d % torch.cuda.device_count() for d in range(ndev)])
    def monce():
        st = m.L.dpx_multi_align_batch(m.h, C.byref(p), inp.sequences.ctypes.data, inp.sequences.size, inp.pairs.ctypes.data, n, sc.numpy().ctypes.data, rc.numpy().ctypes.data, None, None)
        assert st == 0
    for _ in range(2): monce()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); monce(); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"multi ABI, {ndev} workers:", [round(t, 2) for t in ts])
