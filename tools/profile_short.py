"""Small driver for ncu: uploads a 150x150 batch and runs the short-read kernel a few times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpx_gpu_genomics_project_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 3
blob, pairs = synth.uniform_blob_pairs(n, 150, 150, 0x5EED0002)
eng = api.Engine(0)
b = eng.upload(blob, pairs)
p = api.make_params(api.LSW, flags=flags)
for _ in range(4):
    b.run(p)
b.sync()
print(b.stats())
