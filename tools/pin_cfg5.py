"""Pins BASELINE config 5 at its FULL size on the CPU oracle (SURVEY.md §8c last bullet): one 1 Mbp x 1 Mbp LinearSmithWaterman
pair, (score, end row, end col) by the rolling-row restatement oracle/dpx_oracle.c:orc_lsw_score_only (itself pinned
byte-for-byte on the compiled reference for <= 5 kbp pairs, tests/test_oracle.py).  ~1e12 cells, single thread, 15-20 minutes.
Writes tests/golden/cfg5_1m.json, which bench.py --config 5 and tests/test_gpu_longpair.py assert against.

  python tools/pin_cfg5.py [R [Q]]      (defaults 1 000 000 x 1 000 000; the generator, seed and weights are bench.py's LONG)
"""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
from dpx_gpu_genomics_project_b200 import synth

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else R
SEED, MUT = 0x5EED0005, (0.01, 0.001, 0.001)
W = dict(match=3, mismatch=-1, gap_open=-2)
img = synth.mutated_fixed_file_bytes(1, R, Q, SEED, *MUT)
ref = img[2:2 + R].tobytes(); qry = img[3 + R:3 + R + Q].tobytes()
t0 = time.perf_counter()
score, row, col = ol.lsw_score_only(ol.params(ol.LSW, **W), ref, qry)
dt = time.perf_counter() - t0
out = {"R": R, "Q": Q, "seed": f"{SEED:#x}", "generator": "synth.mutated_fixed_file_bytes(1, R, Q, seed, 0.01, 0.001, 0.001)",
       "weights": W, "score": score, "end_row": row, "end_col": col,
       "ref_sha256": hashlib.sha256(ref).hexdigest(), "qry_sha256": hashlib.sha256(qry).hexdigest(),
       "oracle": "oracle/dpx_oracle.c:orc_lsw_score_only (rolling row, 1 thread)", "oracle_seconds": round(dt, 1),
       "oracle_gcups": round(R * Q / dt / 1e9, 3)}
name = "cfg5_1m.json" if (R, Q) == (1_000_000, 1_000_000) else f"cfg5_{R}x{Q}.json"
with open(os.path.join(ROOT, "tests", "golden", name), "w") as f:
    json.dump(out, f, indent=1); f.write("\n")
print(json.dumps(out))
