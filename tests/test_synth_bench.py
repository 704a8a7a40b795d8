"""CPU checks of the host-side helpers the benchmark relies on: deterministic generators in parseInput form, the in-band cell
count used for GCUPS of the banded config, and the workload table of bench.py."""
import importlib.util
import os

import numpy as np

import oracle_lib as ol
from dpx_gpu_genomics_project_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def test_generators_are_deterministic_and_parse_back():
    for make in (lambda: synth.uniform_blob_pairs(50, 30, 20, 7), lambda: synth.mutated_blob_pairs(40, 60, 55, 9, 0.05, 0.02, 0.02),
                 lambda: synth.ragged_mutated_blob_pairs(60, 10, 90, 11, 0.05, 0.02, 0.02)):
        b1, p1 = make(); b2, p2 = make()
        assert (b1 == b2).all() and (p1 == p2).all()
        blob, pairs = ol.parse_image(synth.blob_to_file_bytes(b1))          # the reference parser's view of the same file
        assert (pairs == p1).all() and (blob == b1).all()
        seqs = b1[np.concatenate([np.arange(r, r + n) for r, n in zip(p1["referenceIdx"], p1["referenceSize"])])]
        assert set(seqs.tolist()) <= set(b"0123")


def test_ragged_queries_are_related_to_their_references():
    blob, pairs = synth.ragged_mutated_blob_pairs(30, 100, 300, 3, 0.05, 0.02, 0.02)
    s, _, _ = ol.align_batch(ol.params(ol.LNW), blob, pairs, strings=False, threads=4)
    assert (s > 1.5 * pairs["referenceSize"]).all()                         # ~9 % divergence: scores stay close to 3 * length


def test_banded_cell_count_matches_brute_force():
    bench = _bench()
    wl = dict(bench.WORKLOADS[4])
    for Q, R, W in ((50, 50, 8), (40, 70, 64), (300, 120, 17)):
        wl["weights"] = dict(wl["weights"], band=W)
        pairs = np.zeros(3, dtype=ol.PAIR_DTYPE); pairs["querySize"] = Q; pairs["referenceSize"] = R
        want = sum(1 for i in range(1, Q + 1) for j in range(1, R + 1) if abs(i - j) <= W) * 3
        assert bench.total_cells(wl, pairs) == want


def test_workload_table_names_the_baseline_configs():
    bench = _bench()
    assert sorted(bench.WORKLOADS) == [1, 2, 3, 4]
    assert bench.WORKLOADS[2]["pairs"] == 1_000_000 and (bench.WORKLOADS[2]["R"], bench.WORKLOADS[2]["Q"]) == (150, 150)
    assert bench.WORKLOADS[3]["pairs"] == 100_000 and bench.WORKLOADS[3]["algo"] == "ANW" and bench.WORKLOADS[3]["strings"]
    assert bench.WORKLOADS[4]["weights"]["band"] == 64 and bench.WORKLOADS[4]["pairs"] == 10_000
    for wl in bench.WORKLOADS.values():
        assert bench.cpu_sample_size(wl, wl["pairs"], 16) >= 1


def test_config5_reference_arm_prints_the_contract_line():
    """bench.py --config 5 --impl reference (CPU only): the rolling-row port on a bounded sample, same JSON shape as the other configs."""
    import json
    import subprocess
    import sys
    bench = _bench()
    assert (bench.LONG["R"], bench.LONG["Q"]) == (1_000_000, 1_000_000)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "5", "--impl", "reference", "--pairs", "3000", "--steps", "1"],
                         check=True, capture_output=True, timeout=300).stdout
    line = json.loads(out.decode().strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "GCUPS" and line["scaling"] == "strong" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["config"]["baseline_config"] == 5


def test_every_roofline_key_of_the_bench_is_in_the_committed_sass_counts():
    """bench.py takes a kernel's clocks per cell from profiles/sass_counts.json by key; a renamed template parameter or a missing
    instantiation would silently drop the roofline from the line.  Every key the default run can ask for must be there, with the
    three resources of the model, and the bound must be the largest of them."""
    bench = _bench()
    keys = []
    for no, wl in bench.WORKLOADS.items():
        for track in (True, False):
            k = wl["sass"].format(track=track)
            keys.append(k)
            if no != 2:
                k2 = k.replace("traceback=True", "traceback=False")
                keys.append(k2.replace("K=8", "K=16") if no == 3 else k2)
    keys.append(bench.LONG["sass"])
    for k in set(keys):
        m = bench.sass_counts(k)
        assert m is not None, k
        assert m["alu_clk_per_cell"] > 0 and m["issued_instr_per_cell"] > 0 and m["fma_clk_per_cell"] >= 0
        assert abs(bench.issue_roof_cells_per_clk_per_sm(m) - 128.0 / max(m["alu_clk_per_cell"], m["fma_clk_per_cell"], m["issued_instr_per_cell"])) < 1e-9
        assert 0 < m["half_rate_alu_instr_per_cell"] * 2 <= m["alu_clk_per_cell"] + 1e-9
