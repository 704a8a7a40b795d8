"""One run of the long-pair chain at the per-GPU shape of an 8-GPU stripe (125 000 columns), for ncu.
usage: python tools/profile_long2.py [K]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpx_gpu_genomics_project_b200 import api, synth, longpair
K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
eng = api.Engine(0)
eng.set_option("long_k", K)
rng = synth.Rng(5)
ref = synth.random_seq(rng, 125_000); qry = synth.random_seq(rng, 60_000)
job = longpair.StripedLongPair(eng, api.make_params(api.LSW), ref, qry, 0, 1, None)
print(job.run())
job.free()
