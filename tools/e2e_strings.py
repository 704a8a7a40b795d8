"""One-call time of the string configs (BASELINE 3 and 4) as a function of the number of chunks of the pipelined route.
usage: python tools/e2e_strings.py 3|4 [pairs]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dpx_gpu_genomics_project_b200 import api, synth
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
if cfg == 3:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    img = synth.mutated_fixed_file_bytes(n, 1000, 1000, 0x5EED0003, 0.02, 0.005, 0.005)
    p = api.make_params(api.ANW, gap_open=-3, gap_extend=-1, flags=7)
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
    img = synth.mutated_fixed_file_bytes(n, 10000, 10000, 0x5EED0004, 0.05, 0.01, 0.01)
    p = api.make_params(api.BSW, gap_open=-2, band=64, flags=7)
inp = api.parse_image_native(img)
eng = api.Engine(0)
sc = torch.empty(n, dtype=torch.int32).pin_memory().numpy(); rc = torch.empty((n, 2), dtype=torch.int32).pin_memory().numpy()    # page-locked, like bench.py (pageable result buffers serialise the chunks)
def once():
    sb, so = C.c_void_p(), C.c_void_p()
    st = eng.L.dpx_align_batch(eng.ctx, C.byref(p), inp.sequences.ctypes.data, inp.sequences.size, inp.pairs.ctypes.data, n, sc.ctypes.data, rc.ctypes.data, C.byref(sb), C.byref(so))
    assert st == 0, eng.L.dpx_last_error(eng.ctx)
    eng.L.dpx_free(sb); eng.L.dpx_free(so)
for ch in (1, 2, 3, 4, 6, 8):
    eng.set_option("chunks_strings", ch)
    once(); once()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); once(); ts.append((time.perf_counter() - t0) * 1e3)
    print("config", cfg, "chunks", ch, "ms per call", [round(t, 2) for t in ts])
eng.set_option("serial_strings", 1)
once(); once(); t0 = time.perf_counter(); once(); print("serial route", round((time.perf_counter() - t0) * 1e3, 2))
