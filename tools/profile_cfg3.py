"""One Gotoh fill + walk at config 3's shape (one wave of warps) for an ncu capture with source counters.
usage: ncu --set full --import-source on --clock-control none -k regex:pw_nw -c 1 -o gpurun_out/pw python tools/profile_cfg3.py [pairs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpx_gpu_genomics_project_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 23680          # 2 pairs per warp, 20 warps per SM: four full waves
eng = api.Engine(0)
eng.set_option("serial_chunks", 1)
blob, pairs = synth.mutated_blob_pairs(n, 1000, 1000, 0x5EED0003, 0.02, 0.005, 0.005)
b = eng.upload(blob, pairs)
b.run(api.make_params(api.ANW, gap_open=-3, gap_extend=-1, flags=api.OUT_SCORE | api.OUT_END_COORDS | api.OUT_STRINGS)); b.sync()
print("cfg3", b.stats()); b.free()
