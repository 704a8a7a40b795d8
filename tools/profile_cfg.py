"""Small driver for ncu: one run of BASELINE config 3 (Gotoh + traceback) and config 4 (banded SW + traceback) at
profile-friendly sizes.  usage: python tools/profile_cfg.py [n3] [n4]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
from dpx_gpu_genomics_project_b200 import api, synth
n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
n4 = int(sys.argv[2]) if len(sys.argv) > 2 else 2400
eng = api.Engine(0)
ALL = api.OUT_SCORE | api.OUT_END_COORDS | api.OUT_STRINGS
if n3:
    blob, pairs = ol.parse_image(synth.mutated_fixed_file_bytes(n3, 1000, 1000, 0x5EED0003, 0.02, 0.005, 0.005))
    b = eng.upload(blob, pairs)
    for _ in range(2):
        b.run(api.make_params(api.ANW, gap_open=-3, gap_extend=-1, flags=ALL)); b.sync()
    print("cfg3", b.stats()); b.free()
if n4:
    blob, pairs = ol.parse_image(synth.mutated_fixed_file_bytes(n4, 10000, 10000, 0x5EED0004, 0.05, 0.01, 0.01))
    b = eng.upload(blob, pairs)
    for _ in range(2):
        b.run(api.make_params(api.BSW, gap_open=-2, band=64, flags=ALL)); b.sync()
    print("cfg4", b.stats()); b.free()
