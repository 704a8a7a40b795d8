"""One launch of the short-read kernel at config 2's shape for an ncu capture with source counters.
usage: ncu --set full --import-source on --clock-control none -k regex:sr_lsw -c 1 -o gpurun_out/sr python tools/profile_cfg2.py [pairs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpx_gpu_genomics_project_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
eng = api.Engine(0)
inp = api.parse_image_native(synth.uniform_file_bytes(n, 150, 150, 0x5EED0002))
b = eng.upload(inp.sequences, inp.pairs)
b.run(api.make_params(api.LSW, flags=api.OUT_SCORE | api.OUT_END_COORDS)); b.sync()
print("cfg2", b.stats()); b.free()
