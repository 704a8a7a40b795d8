"""Workload for profiling the long-pair alignment path: one pair with its alignment strings.
usage: python tools/profile_longtrace.py [R] [Q]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpx_gpu_genomics_project_b200 import api, synth
R = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else R
rng = synth.Rng(0x5EED0005)
ref = synth.random_seq(rng, R); qry = (synth.mutate(rng, ref, 0.01, 0.001, 0.001) + synth.random_seq(rng, Q))[:Q]
eng = api.Engine(0)
end, start, lines, st = eng.align_long_pair_strings(api.make_params(api.LSW), ref, qry)
print(end, start, len(lines[0]), st)
