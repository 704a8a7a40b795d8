#!/usr/bin/env python
"""bench.py — benchmark of the alignment hot path (BASELINE.json metric: GCUPS).

Default workload (N=1, BASELINE.json configs[1]): batched LinearSmithWaterman, 1 000 000 synthetic DNA pairs of
150 x 150 bp, score + end coordinates, match 3 / mismatch -1 / gap -2.  `--config 1` runs the substitute for BASELINE configs[0] (LinearNeedlemanWunsch with
traceback on ragged 100-300 bp pairs; the bundled pairs are missing from the reference mount), `--config 3` / `--config 4` run BASELINE configs
[2] / [3] (AffineNeedlemanWunsch 1000 x 1000 with full traceback + alignment strings; BandedSmithWaterman band 64 on
10 kbp x 10 kbp with traceback + strings) through the same code.  A "step" is one pass of the hot path over the batch.
With N>1 (torchrun, one rank per GPU) every rank aligns its own shard of that size (independent pairs: no data-path
collective, weak scaling); value = cells of all ranks / max-over-ranks device time.

  value     GCUPS with the batch already resident (packed) in HBM: CUDA events around dpx_batch_run (fill kernel, and for
            traceback configs the GPU backtrack that leaves the alignment strings in HBM).
  e2e       GCUPS through the one-call C ABI (dpx_align_batch) with HOST buffers in and out: H2D of the parseInput blob +
            index, alphabet scan + 2-bit pack, kernels, D2H of scores / end cells (and of the strings for configs 3, 4).
  roofline  DPX-issue roofline of the dominant kernel (north_star / SURVEY.md §8d): cells/clk/SM = min(64 / ALU-pipe instr
            per cell, 128 / issued instr per cell) from the committed SASS counts (profiles/sass_counts.json) x SMs x SM
            clock sampled during the timed region; the HBM view of the same launch (algorithmic bytes / kernel time vs
            MEASURED_PEAKS.json) is reported beside it — for the traceback configs that is the packed traceback stream.
  cpu_baseline  the reference's own C++ classes (oracle/_ref/ref_align) on the host cores, bounded sample (config 4: the C
            port, because the reference's banded class is not executable).

`--config 5` is BASELINE configs[4]: ONE 1 Mbp x 1 Mbp LinearSmithWaterman pair, score + end cell; with N > 1 the pair is split into
column stripes (strong scaling, see main_long).

`--impl reference` times that CPU path alone (rank 0 only) and prints the same line shape.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "GCUPS"

# BASELINE.json configs (1-based numbering of SURVEY.md §8d): shapes, weights, outputs, the kernel that dominates the step
WORKLOADS = {
    1: dict(algo="LNW", R=200, Q=200, pairs=10_000, weights=dict(match=3, mismatch=-1, gap_open=-2), strings=True, seed=0x5EED0001,
            gen=("ragged", 100, 300, 0.05, 0.02, 0.02), dtype="int16x2", kernel="pw_nw_kernel<LNW,TB,K=8> (+ pw_bt_kernel)",
            sass="pairwf_s16x2:algo=0,traceback=True,K=8",
            title="LinearNeedlemanWunsch batch (substitute for the missing bundled pairs): {pairs} pairs, R ~ U[100,300], query = reference mutated "
                  "5 % / 2 % / 2 %, 2-bit traceback + alignment strings, match 3 / mismatch -1 / gap -2", cpu_per_core=0.05e9, tb_bytes_per_cell=0.25),
    2: dict(algo="LSW", R=150, Q=150, pairs=1_000_000, weights=dict(match=3, mismatch=-1, gap_open=-2), strings=False, seed=0x5EED0002,
            gen="uniform", dtype="int16x2", kernel="sr_lsw_kernel<G=8,K=19>", sass="shortread_s16x2:G=8,K=19,track={track},xormode=False",
            title="LinearSmithWaterman batch: {pairs} pairs x (150x150) bp per GPU, score{ends}, match 3 / mismatch -1 / gap -2",
            cpu_per_core=0.045e9),
    3: dict(algo="ANW", R=1000, Q=1000, pairs=100_000, weights=dict(match=3, mismatch=-1, gap_open=-3, gap_extend=-1), strings=True, seed=0x5EED0003,
            gen=(0.02, 0.005, 0.005), dtype="int16x2", kernel="pw_nw_kernel<ANW,TB,K=8> (+ pw_bt_kernel)", sass="pairwf_s16x2:algo=1,traceback=True,K=8",
            title="AffineNeedlemanWunsch (Gotoh H/E/F) batch: {pairs} pairs x (1000x1000) bp per GPU, full 4-bit traceback + alignment strings, "
                  "match 3 / mismatch -1 / open -3 / extend -1", cpu_per_core=0.024e9, tb_bytes_per_cell=0.5),
    4: dict(algo="BSW", R=10000, Q=10000, pairs=10_000, weights=dict(match=3, mismatch=-1, gap_open=-2, band=64), strings=True, seed=0x5EED0004,
            gen=(0.05, 0.01, 0.01), dtype="int32", kernel="band_sw_kernel<M=2,EXTRA,TB> (+ band_bt_kernel)", sass="band_s32:M=2,extra=True,traceback=True",
            title="BandedSmithWaterman band 64: {pairs} pairs x (10 kbp x 10 kbp) per GPU, 2-bit traceback + alignment strings, in-band cells, "
                  "match 3 / mismatch -1 / gap -2", cpu_per_core=0.05e9, tb_bytes_per_cell=0.25),
}


# BASELINE config 5: ONE pair of 1 Mbp x 1 Mbp, LinearSmithWaterman score + end cell (strong scaling: column stripes over NVLink)
LONG = dict(R=1_000_000, Q=1_000_000, seed=0x5EED0005, weights=dict(match=3, mismatch=-1, gap_open=-2), mutate=(0.01, 0.001, 0.001),
            cpu_sample=80_000, kernel="long_sw_kernel<K=32,PACK,TABLE>", sass="long_s32:K=32,pack=True,table=True,ck=False")


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def sass_counts(key):
    """ALU-pipe and issued instructions per cell of a kernel's hot loop, from the committed SASS counts."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "sass_counts.json")))
        k = d[key]
        return k["alu_per_cell"], k["issue_per_cell"]
    except Exception:
        return None, None


def make_inputs(wl, n_pairs, seed):
    """(blob, pairs) in parseInput's output form for `n_pairs` pairs of the workload."""
    from dpx_gpu_genomics_project_b200 import synth
    if wl["gen"] == "uniform":
        return synth.uniform_blob_pairs(n_pairs, wl["R"], wl["Q"], seed)
    if wl["gen"][0] == "ragged":
        _, lo, hi, sub, ins, dele = wl["gen"]
        return synth.ragged_mutated_blob_pairs(n_pairs, lo, hi, seed, sub, ins, dele)
    sub, ins, dele = wl["gen"]
    return synth.mutated_blob_pairs(n_pairs, wl["R"], wl["Q"], seed, sub, ins, dele)


def total_cells(wl, pairs):
    r = pairs["referenceSize"].astype(np.int64); q = pairs["querySize"].astype(np.int64)
    if wl["algo"] != "BSW":
        return float((r * q).sum())
    W = wl["weights"]["band"]                      # in-band cells: sum_i |{j in [1,R] : |i-j| <= W}|  (uniform lengths here)
    Q, R = int(q[0]), int(r[0])
    i = np.arange(1, Q + 1, dtype=np.int64)
    per_pair = float(np.maximum(0, np.minimum(R, i + W) - np.maximum(1, i - W) + 1).sum())
    return per_pair * len(pairs)


# ---- CPU reference arm ------------------------------------------------------------------------------------
def run_cpu_reference(wl, blob, pairs, n_sample, threads):
    """Reference C++ classes (oracle/_ref/ref_align, compiled from the unmodified sources) on `threads` host threads over the
    first n_sample pairs; the C port (oracle/liboracle.so) when the binary is absent or the algorithm is the banded one
    (c++/BandedSmithWaterman.cpp is not executable).  Returns (gcups, kind, seconds)."""
    import oracle_lib as ol
    from dpx_gpu_genomics_project_b200 import synth
    n_sample = int(max(1, min(n_sample, len(pairs))))
    cells = total_cells(wl, pairs[:n_sample])
    w = wl["weights"]
    if wl["algo"] != "BSW" and ol.have_ref_binary():
        end = int(pairs["queryIdx"][n_sample - 1] + pairs["querySize"][n_sample - 1] + 1)
        with tempfile.NamedTemporaryFile(prefix="dpx_cpu_", suffix=".txt", delete=False) as f:
            f.write(synth.blob_to_file_bytes(blob[:end]).tobytes())
            path = f.name
        try:
            cmd = [ol.REF_ALIGN, "-algo", wl["algo"], "-pairs", path, "-match", str(w["match"]), "-mismatch", str(w["mismatch"]),
                   "-open", str(w["gap_open"]), "-extend", str(w.get("gap_extend", -1)), "-threads", str(threads), "-noheader"]
            p = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, check=True)
            usec = float(p.stderr.decode().split("ALIGN_USEC")[1].split()[0])
        finally:
            os.unlink(path)
        return cells / (usec * 1e-6) / 1e9, "reference", usec * 1e-6
    algo = {"LNW": ol.LNW, "ANW": ol.ANW, "LSW": ol.LSW, "BSW": ol.BSW}[wl["algo"]]
    t0 = time.perf_counter()
    ol.align_batch(ol.params(algo, **w), blob, pairs[:n_sample], strings=wl["strings"], threads=threads, bandmem=(wl["algo"] == "BSW"))
    dt = time.perf_counter() - t0
    return cells / dt / 1e9, "port", dt


def cpu_sample_size(wl, n_pairs, cores, seconds=12.0):
    cells_per_pair = wl["R"] * wl["Q"] if wl["algo"] != "BSW" else wl["Q"] * (2 * wl["weights"]["band"] + 1)   # config 1: mean lengths
    return int(max(min(n_pairs, 2 * cores), min(n_pairs, wl["cpu_per_core"] * cores * seconds / cells_per_pair)))


def main_long(args):
    """--config 5: one step = one alignment of the 1 Mbp x 1 Mbp pair.  N > 1: the reference columns are split into one stripe per
    GPU (multi-GPU mode B, the right edge of a stripe streams to the next GPU by NVLink P2P stores): total work is fixed, so
    scaling is "strong".  value = R*Q / max-over-ranks kernel time (CUDA events inside the library, sequences resident);
    e2e = the host-buffer call (N = 1: dpx_align_long_pair, H2D of both sequences + kernel + D2H of the result;
    N > 1: reset + barrier + run + gather of the per-stripe results, sequences resident since stripe creation)."""
    import oracle_lib as ol
    from dpx_gpu_genomics_project_b200 import synth
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    cores = os.cpu_count() or 1
    R, Q, w = LONG["R"], LONG["Q"], LONG["weights"]
    if args.pairs:                                            # --pairs N shrinks the pair to N x N bases (smoke runs)
        R = Q = args.pairs

    def make_pair(r_len, q_len):
        img = synth.mutated_fixed_file_bytes(1, r_len, q_len, LONG["seed"], *LONG["mutate"])
        return img[2:2 + r_len].tobytes(), img[3 + r_len:3 + r_len + q_len].tobytes()

    config = {"workload": f"LinearSmithWaterman, ONE pair of {R} x {Q} bp (query = reference mutated 1 % / 0.1 % / 0.1 %), score + end cell, "
                          "match 3 / mismatch -1 / gap -2", "baseline_config": 5, "R": R, "Q": Q,
              "sharding": f"{max(args.gpus, world)} column stripe(s), right edges streamed GPU to GPU over NVLink P2P (no NCCL on the data path)",
              "l2": "256 MiB memset between timed steps", "seed": f"{LONG['seed']:#x}"}
    if args.impl == "reference":
        if rank != 0:
            return
        n = min(LONG["cpu_sample"], R)
        ref, qry = make_pair(n, n)
        secs = []
        for _ in range(max(1, min(args.steps, 2))):
            t0 = time.perf_counter(); ol.lsw_score_only(ol.params(ol.LSW, **w), ref, qry); secs.append(time.perf_counter() - t0)
        v = n * n / float(np.mean(secs)) / 1e9
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "GCUPS", "n_gpus": max(args.gpus, world), "steps": len(secs), "warmup": 0,
            "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": 1, "kind": "port",
                             "sample": f"a {n} x {n} bp pair of the same generator; rolling-row C port of LinearSmithWaterman (the reference's "
                                       "full-matrix class needs 8 B per cell: 8 TB at 1 Mbp x 1 Mbp), 1 thread"},
            "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return

    import torch
    from dpx_gpu_genomics_project_b200 import api, longpair
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libdpxalign has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    ref, qry = make_pair(R, Q)
    eng = api.Engine(local)
    params = api.make_params(api.LSW, **w)
    job = longpair.StripedLongPair(eng, params, ref, qry, rank, world, dist)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    steps = min(args.steps, 10)
    for _ in range(args.warmup):
        res, _ = job.run()
    sampler = ClockSampler(local)
    sampler.start()
    ms_steps, wall = [], []
    for _ in range(steps):
        flush.zero_(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        res, ms = job.run()                                   # barrier inside; ms = max over ranks of the stripes' kernel times
        wall.append(time.perf_counter() - t0); ms_steps.append(ms)
    sampler.stop_flag.set(); sampler.join(timeout=1.0)
    e2e_s = float(np.mean(wall))
    h2d = 0
    if world == 1:                                            # the host-buffer call of the ABI
        eng.align_long_pair(params, ref, qry)
        t0 = time.perf_counter()
        for _ in range(3):
            res1 = eng.align_long_pair(params, ref, qry)
        e2e_s = (time.perf_counter() - t0) / 3
        assert tuple(res1) == tuple(res), "one-call and striped paths disagree"
        h2d = R + Q
    job.free()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    ms_per_step = float(np.mean(ms_steps))
    cells = float(R) * float(Q)
    value = cells / (ms_per_step * 1e-3) / 1e9
    clocks = sampler.result()
    peaks, _ = measured_peaks()
    f_mhz = clocks["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)
    roofline = {"bound": "dpx_issue", "achieved": value, "unit": "GCUPS", "kernel": LONG["kernel"], "traffic": None, "kernel_ms": ms_per_step}
    alu_pc, issue_pc = sass_counts(LONG["sass"])
    if alu_pc and world == 1 and R >= 600_000:                # the lane width (hence the SASS loop) is K = 32 only for one wide stripe
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        peak = min(64.0 / alu_pc, 128.0 / issue_pc) * sms * f_mhz * 1e6 / 1e9
        roofline.update({"peak": peak, "frac": value / peak,
                         "model": {"alu_instr_per_cell": alu_pc, "issued_instr_per_cell": issue_pc, "sms": sms, "sm_mhz": f_mhz, "sass_key": LONG["sass"],
                                   "note": "a chain of 977 warps (1.65 per SM sub-partition), each row step a dependent chain: latency-bound below the issue roofline"}})
    out = {"metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
           "clocks": clocks, "result": list(res),
           "e2e": {"value": cells / e2e_s / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 20 * world, "ms_per_step": e2e_s * 1e3,
                   "api": "dpx_align_long_pair (C ABI), host sequences in, (score, row, col) out" if world == 1 else
                          "dpx_stripe_reset + barrier + dpx_stripe_run + gather of the per-stripe results; sequences resident since dpx_stripe_create"},
           "gpu_launches": steps * 1, "roofline": roofline}
    if not args.no_cpu_baseline and world == 1:
        n = min(LONG["cpu_sample"], R)
        r2, q2 = make_pair(n, n)
        t0 = time.perf_counter(); want = ol.lsw_score_only(ol.params(ol.LSW, **w), r2, q2); dt = time.perf_counter() - t0
        got = eng.align_long_pair(params, r2, q2)
        out["parity_spot_check"] = tuple(got) == tuple(want)
        out["cpu_baseline"] = {"value": n * n / dt / 1e9, "unit": "GCUPS", "cores": 1, "kind": "port", "seconds": dt,
                               "sample": f"a {n} x {n} bp pair of the same generator, rolling-row C port (the reference's full-matrix class cannot "
                                         "allocate this problem), 1 thread; also the parity check of the GPU result at that size"}
    print(json.dumps(out))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dpx", choices=["dpx", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(WORKLOADS) + [5], help="BASELINE.json config number (SURVEY.md §8d)")
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU (default: the config's own size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--score-only", action="store_true", help="config 2: omit end coordinates; configs 3/4: no traceback / strings")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "dpx":
        args.warmup = 3
    if args.config == 5:
        return main_long(args)
    wl = dict(WORKLOADS[args.config])
    n_pairs = args.pairs or wl["pairs"]
    want_strings = wl["strings"] and not args.score_only
    if args.config != 2 and args.steps > 10 and args.impl == "dpx":
        args.steps = 10                                       # traceback configs: ~10-60 ms per step plus the strings download in e2e

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    n_gpus = max(args.gpus, world)
    cores = os.cpu_count() or 1

    ends = "" if (args.config == 2 and args.score_only) else (" + end coords" if args.config == 2 else "")
    config = {"workload": wl["title"].format(pairs=n_pairs, ends=ends) + ("" if want_strings or not wl["strings"] else " [score only]"),
              "baseline_config": args.config, "pairs_per_gpu": n_pairs, "R": wl["R"], "Q": wl["Q"],
              "sharding": f"independent pairs x{n_gpus} (no collective)", "l2": "256 MiB memset between timed steps",
              "seed": f"{wl['seed']:#x} + rank"}

    # ---------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        n_s = cpu_sample_size(wl, n_pairs, cores, seconds=8.0)
        blob, pairs = make_inputs(wl, n_s, wl["seed"])
        vals, secs = [], []
        run_cpu_reference(wl, blob, pairs, max(2, n_s // 8), cores)       # warm-up
        kind = "reference"
        for _ in range(max(1, args.steps if args.steps <= 5 else 3)):
            v, kind, dt = run_cpu_reference(wl, blob, pairs, n_s, cores)
            vals.append(v); secs.append(dt)
        v = float(np.mean(vals))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "GCUPS", "n_gpus": n_gpus, "steps": len(vals),
            "warmup": 1, "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": cores, "kind": kind,
                             "sample": f"first {n_s} pairs of the workload per step, {cores} host threads"},
            "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    # ---------------------------------------------------------------------------------------------------
    import torch
    from dpx_gpu_genomics_project_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libdpxalign has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))

    blob, pairs = make_inputs(wl, n_pairs, wl["seed"] + rank)
    cells = total_cells(wl, pairs)
    algo = {"LNW": api.LNW, "ANW": api.ANW, "LSW": api.LSW, "BSW": api.BSW}[wl["algo"]]
    if args.config == 2:
        flags = api.OUT_SCORE | (0 if args.score_only else api.OUT_END_COORDS)
    else:
        flags = api.OUT_SCORE | api.OUT_END_COORDS | (api.OUT_STRINGS if want_strings else 0)
    params = api.make_params(algo, flags=flags, **wl["weights"])

    eng = api.Engine(local)
    stream = torch.cuda.Stream()                         # a real (non-default) stream shared by torch events and the library
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    batch = eng.upload(blob, pairs)                      # resident + packed in HBM before the timed region
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        batch.run(params)
    batch.sync()
    st0 = batch.stats()
    launches_per_step = int(st0["kernel_launches"])

    sampler = ClockSampler(local)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                                    # L2 flush, outside the per-step event bracket
        ev[k][0].record(stream)
        batch.run(params)
        ev[k][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    sampler.stop_flag.set(); sampler.join(timeout=1.0)
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    ms_total = float(sum(ms_steps))
    batch.sync()
    st_last = batch.stats()
    tb_bytes = float(st_last["traceback_bytes"])
    # The dominant kernel timed ALONE (library's own event pairs): traceback runs pipeline their chunks (fill of chunk c+1 next to
    # the walk of chunk c, on separate streams), so for the roofline the same step is repeated with the chunks serialised.
    if want_strings:
        os.environ["DPX_SERIAL_CHUNKS"] = "1"
        for _ in range(2):
            flush.zero_(); batch.run(params)
        batch.sync()
        st_last = batch.stats()
        del os.environ["DPX_SERIAL_CHUNKS"]
        batch.run(params); batch.sync()                  # back to the pipelined slab layout before the fetch
    fill_ms = float(st_last["fill_ms"])
    bt_ms = float(st_last["backtrack_ms"])

    # result check of the timed configuration on a sample (outside the timed region)
    import oracle_lib as ol
    n_chk = 2000 if args.config in (1, 2) else (64 if args.config == 3 else 8)
    res = batch.fetch() if not want_strings else None
    oalgo = {"LNW": ol.LNW, "ANW": ol.ANW, "LSW": ol.LSW, "BSW": ol.BSW}[wl["algo"]]

    # ---- e2e: one-call ABI, pinned host buffers, H2D + D2H inside ------------------------------------------
    e2e_steps = max(2, min(5 if args.config in (1, 2) else 3, args.steps))
    pin_blob = torch.from_numpy(np.ascontiguousarray(blob)).pin_memory()
    pin_pairs = torch.from_numpy(np.ascontiguousarray(pairs).view(np.int32)).pin_memory()
    nb, npairs_bytes = pin_blob.numel(), pin_pairs.numel() * 4
    blob_p = pin_blob.numpy()
    pairs_p = pin_pairs.numpy().view(api.PAIR_DTYPE)
    out_scores = torch.empty(n_pairs, dtype=torch.int32).pin_memory()
    out_rc = torch.empty((n_pairs, 2), dtype=torch.int32).pin_memory()
    import ctypes as C
    L = eng.L
    str_bytes = int(3 * (pairs["referenceSize"].astype(np.int64) + pairs["querySize"] + 1).sum()) if want_strings else 0
    e2e_first_strings = []

    def e2e_once(keep=False):
        sb, so = C.c_void_p(), C.c_void_p()
        st = L.dpx_align_batch(eng.ctx, C.byref(params), blob_p.ctypes.data, nb, pairs_p.ctypes.data, n_pairs,
                               out_scores.numpy().ctypes.data, out_rc.numpy().ctypes.data,
                               C.byref(sb) if want_strings else None, C.byref(so) if want_strings else None)
        if st != 0:
            raise RuntimeError(f"dpx_align_batch failed: {st} {L.dpx_last_error(eng.ctx)}")
        if want_strings:
            if keep:
                last = int(np.frombuffer(C.string_at(so.value + (3 * n_pairs - 1) * 8, 8), dtype=np.uint64)[0])
                e2e_first_strings.append(last + len(C.string_at(sb.value + last)) + 1)      # size of the compacted blob
                offs = np.frombuffer(C.string_at(so, 3 * n_chk * 8), dtype=np.uint64)
                for i in range(n_chk):
                    e2e_first_strings.append(tuple(C.string_at(sb.value + int(offs[3 * i + k])) for k in range(3)))
            L.dpx_free(sb); L.dpx_free(so)

    e2e_once(keep=True)
    if want_strings:
        str_bytes = e2e_first_strings.pop(0)              # what actually crosses PCIe: the compacted strings
    # context for the e2e number: what a bare pinned H2D copy of the same input bytes costs on this box
    dev_blob = torch.empty(nb, dtype=torch.uint8, device="cuda")
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev_blob.copy_(pin_blob, non_blocking=True)
    torch.cuda.synchronize()
    c0.record(stream); dev_blob.copy_(pin_blob, non_blocking=True); c1.record(stream)
    torch.cuda.synchronize()
    h2d_ms = c0.elapsed_time(c1)
    del dev_blob
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_once()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if res is not None:
        assert (out_scores.numpy() == res.scores).all(), "e2e and staged paths disagree"

    # ---- reduce over ranks -------------------------------------------------------------------------------
    if dist is not None:
        t = torch.tensor([ms_total, e2e_s, fill_ms, bt_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s, fill_ms, bt_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    ms_per_step = ms_total / args.steps
    value = cells * world / (ms_per_step * 1e-3) / 1e9
    e2e_value = cells * world / e2e_s / 1e9

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- parity spot check against the oracle (outside every timed region) --------------------------------
    s_ref, e_ref, t_ref = ol.align_batch(ol.params(oalgo, **wl["weights"]), blob, pairs[:n_chk], strings=want_strings, threads=min(cores, 16))
    parity_ok = bool((out_scores.numpy()[:n_chk] == s_ref).all())
    if wl["algo"] in ("LSW", "BSW") and not (args.config == 2 and args.score_only):
        parity_ok = parity_ok and bool((out_rc.numpy()[:n_chk] == e_ref).all())
    if want_strings:
        parity_ok = parity_ok and e2e_first_strings == t_ref

    # ---- roofline ---------------------------------------------------------------------------------------
    clocks = sampler.result()
    peaks, peaks_src = measured_peaks()
    f_mhz = clocks["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)
    sass_key = wl["sass"].format(track=not args.score_only)
    if args.config != 2 and not want_strings:
        sass_key = sass_key.replace("traceback=True", "traceback=False")
    alu_pc, issue_pc = sass_counts(sass_key)
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    kern_ms = fill_ms if fill_ms > 0 else ms_per_step          # the dominant kernel = the fill kernel(s) of one step
    cells_rank = cells
    kernel_gcups = cells_rank / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "dpx_issue", "achieved": kernel_gcups, "unit": "GCUPS", "kernel": wl["kernel"].split(" (+")[0], "traffic": None,
                "kernel_ms": kern_ms}
    if alu_pc:
        slot_eff = (2 * wl["weights"]["band"] + 1) / 160.0 if args.config == 4 else 1.0   # band 64: 129 diagonals on 5 slots x 32 lanes
        cells_clk_sm = min(64.0 / alu_pc, 128.0 / issue_pc) * slot_eff
        peak = cells_clk_sm * sms * f_mhz * 1e6 / 1e9
        roofline.update({"peak": peak, "frac": kernel_gcups / peak,
                         "model": {"alu_pipe_lanes_per_clk_per_sm": 64, "issue_lanes_per_clk_per_sm": 128,
                                   "alu_instr_per_cell": alu_pc, "issued_instr_per_cell": issue_pc, "slot_efficiency": slot_eff,
                                   "sms": sms, "sm_mhz": f_mhz, "sass_key": sass_key,
                                   "source": "profiles/sass_counts.json + profiles/r01_dpx_microbench*.json"}})
    if args.config == 2:
        # HBM view: algorithmic bytes = 2-bit bases + the index entry + 12 B of results per pair
        algo_bytes = n_pairs * ((wl["R"] + wl["Q"]) * 0.25 + 16 + 8 + 12)
    else:
        # HBM view: the packed traceback stream (0.5 B/cell Gotoh, 0.25 B/cell linear / banded) written once by the fill kernel
        algo_bytes = cells_rank * wl["tb_bytes_per_cell"] if want_strings else n_pairs * ((wl["R"] + wl["Q"]) * 0.25 + 28)
    if args.config == 1:
        roofline["note"] = "ragged duos: a warp sweeps max(Q) x max(R) of its two pairs in passes of 256 rows, so ~60 % of its cell slots hold real cells"
    hbm_ach = algo_bytes / (kern_ms * 1e-3) / 1e9
    roofline["hbm"] = {"achieved": hbm_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / peaks["hbm_gbs"],
                       "peak_source": f"MEASURED_PEAKS.json ({peaks_src})", "algorithmic_bytes_per_launch": algo_bytes,
                       "slab_bytes_written": tb_bytes if want_strings else None}

    try:                                                      # DRAM traffic of the dominant kernel, from the committed ncu --set full capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))[str(args.config)]
        if want_strings or args.config == 2:
            per_launch = tr["dram_bytes_per_pair"] * n_pairs if "dram_bytes_per_pair" in tr else tr["dram_bytes_per_cell"] * cells_rank
            roofline["traffic"] = per_launch
            roofline["traffic_unit"] = "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, scaled from the capture's size)"
            roofline["traffic_source"] = tr["source"]
    except Exception:
        pass
    d2h = n_pairs * 12 + str_bytes + (3 * 8 * n_pairs if want_strings else 0)
    out = {"metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": wl["dtype"], "data": "synthetic", "config": config, "clocks": clocks,
           "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": int(nb + npairs_bytes),
                   "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                   "api": "dpx_align_batch (C ABI), pinned host input buffers" + (", library-allocated string blob" if want_strings else ""),
                   "h2d_note": "bytes of the host buffers handed to the call; when the seqPair index of a chunk is an arithmetic progression "
                               "(fixed-length records) the library rebuilds it on the device instead of copying its 16 B per pair",
                   "bare_h2d_copy_ms": h2d_ms, "bare_h2d_gbs": nb / (h2d_ms * 1e-3) / 1e9},
           "gpu_launches": launches_per_step * args.steps, "roofline": roofline,
           "fill_ms_serialised": fill_ms, "backtrack_ms_serialised": bt_ms, "wall_s_timed_region": t_wall, "parity_spot_check": parity_ok}

    if not args.no_cpu_baseline and world == 1:
        n_s = cpu_sample_size(wl, n_pairs, cores)
        v, kind, dt = run_cpu_reference(wl, blob, pairs, n_s, cores)
        out["cpu_baseline"] = {"value": v, "unit": "GCUPS", "cores": cores, "kind": kind, "seconds": dt,
                               "sample": f"first {n_s} pairs of the same workload, {cores} host threads, align loop only"}
    print(json.dumps(out))
    batch.free(); eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
