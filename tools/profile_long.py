import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpx_gpu_genomics_project_b200 import api, synth
R = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000
rng = synth.Rng(5)
ref = synth.random_seq(rng, R); qry = synth.random_seq(rng, Q)
eng = api.Engine(0)
for _ in range(2):
    print(eng.align_long_pair(api.make_params(api.LSW), ref, qry))
