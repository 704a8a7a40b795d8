"""Driver for the round's ncu captures: one run of each dominant kernel at profile-friendly sizes.
usage: python tools/profile_all.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpx_gpu_genomics_project_b200 import api, synth
eng = api.Engine(0)
ALL = api.OUT_SCORE | api.OUT_END_COORDS | api.OUT_STRINGS
eng.set_option("serial_chunks", 1)          # kernels one after the other: clean per-kernel captures
# config 2: short-read kernel, 1M pairs, uploaded from the parser's packed copy (as bench.py does)
inp = api.parse_image_native(synth.uniform_file_bytes(1_000_000, 150, 150, 0x5EED0002))
b = eng.upload(inp.sequences, inp.pairs)
for _ in range(2):
    b.run(api.make_params(api.LSW, flags=api.OUT_SCORE | api.OUT_END_COORDS)); b.sync()
print("cfg2", b.stats()); b.free()
# config 3: Gotoh + traceback, 6000 pairs (one wave of warps)
blob, pairs = synth.mutated_blob_pairs(6000, 1000, 1000, 0x5EED0003, 0.02, 0.005, 0.005)
b = eng.upload(blob, pairs)
for _ in range(2):
    b.run(api.make_params(api.ANW, gap_open=-3, gap_extend=-1, flags=ALL)); b.sync()
print("cfg3", b.stats())
for _ in range(2):
    b.run(api.make_params(api.LSW, gap_open=-2, flags=ALL)); b.sync()
print("lsw+tb", b.stats()); b.free()
# config 4: banded, 2400 pairs
blob, pairs = synth.mutated_blob_pairs(2400, 10000, 10000, 0x5EED0004, 0.05, 0.01, 0.01)
b = eng.upload(blob, pairs)
for _ in range(2):
    b.run(api.make_params(api.BSW, gap_open=-2, band=64, flags=ALL)); b.sync()
print("cfg4", b.stats()); b.free()
