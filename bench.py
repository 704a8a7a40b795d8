#!/usr/bin/env python
"""bench.py — headline benchmark of the alignment hot path (BASELINE.json metric: GCUPS).

Workload at N=1 (BASELINE.json configs[1]): batched LinearSmithWaterman, 1 000 000 synthetic DNA pairs of
150 x 150 bp, score + end coordinates, match 3 / mismatch -1 / gap -2.  A "step" is one pass of the hot path
over that batch.  With N>1 (torchrun, one rank per GPU) every rank aligns its own 1M-pair shard (independent
pairs: no data-path collective, weak scaling); value = cells of all ranks / max-over-ranks device time.

  value     GCUPS with the batch already resident (packed) in HBM: CUDA events around dpx_batch_run.
  e2e       GCUPS through the one-call C ABI (dpx_align_batch) with pinned HOST buffers in and out:
            H2D of the parseInput blob + index, alphabet scan + 2-bit pack, kernel, D2H of scores/end cells.
  roofline  DPX-issue roofline of the dominant kernel (north_star / SURVEY.md §8d):
            cells/clk/SM = min(64 / ALU-pipe instr per cell, 128 / issued instr per cell) from the committed SASS
            counts (profiles/sass_counts.json) x SMs x SM clock sampled during the timed region; the HBM view of
            the same launch (algorithmic bytes / time vs MEASURED_PEAKS.json) is reported beside it.
  cpu_baseline  the reference's own C++ classes (oracle/_ref/ref_align) on the host cores, bounded sample.

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
R_LEN, Q_LEN = 150, 150
WEIGHTS = dict(match=3, mismatch=-1, gap_open=-2)


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


def sass_counts(track=True):
    """ALU-pipe and issued instructions per cell of the short-read kernel, from the committed SASS counts."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "sass_counts.json")))
        k = d[f"shortread_s16x2:G=8,K=19,track={track},xormode=False"]
        return k["alu_per_cell"], k["issue_per_cell"]
    except Exception:
        return None, None


# ---- CPU reference arm ------------------------------------------------------------------------------------
def run_cpu_reference(blob, pairs, n_sample, threads):
    """Reference C++ classes (oracle/_ref/ref_align, compiled from the unmodified sources) on `threads` host
    threads over the first n_sample pairs; falls back to the C port (oracle/liboracle.so) if the binary is
    absent.  Returns (gcups, kind, seconds)."""
    import oracle_lib as ol
    n_sample = int(min(n_sample, len(pairs)))
    cells = float((pairs["referenceSize"][:n_sample].astype(np.int64) * pairs["querySize"][:n_sample]).sum())
    if ol.have_ref_binary():
        from dpx_gpu_genomics_project_b200 import synth
        end = int(pairs["queryIdx"][n_sample - 1] + pairs["querySize"][n_sample - 1] + 1)
        with tempfile.NamedTemporaryFile(prefix="dpx_cpu_", suffix=".txt", delete=False) as f:
            f.write(synth.blob_to_file_bytes(blob[:end]).tobytes())
            path = f.name
        try:
            cmd = [ol.REF_ALIGN, "-algo", "LSW", "-pairs", path, "-match", str(WEIGHTS["match"]), "-mismatch", str(WEIGHTS["mismatch"]),
                   "-open", str(WEIGHTS["gap_open"]), "-threads", str(threads), "-noheader"]
            p = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, check=True)
            usec = float(p.stderr.decode().split("ALIGN_USEC")[1].split()[0])
        finally:
            os.unlink(path)
        return cells / (usec * 1e-6) / 1e9, "reference", usec * 1e-6
    t0 = time.perf_counter()
    ol.align_batch(ol.params(ol.LSW, **WEIGHTS), blob, pairs[:n_sample], strings=False, threads=threads)
    dt = time.perf_counter() - t0
    return cells / dt / 1e9, "port", dt


def cpu_sample_size(n_pairs, cores, seconds=12.0):
    per_core = 0.045e9                       # reference LSW ~0.05 GCUPS / core
    return int(max(2000, min(n_pairs, per_core * cores * seconds / (R_LEN * Q_LEN))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dpx", choices=["dpx", "reference"])
    ap.add_argument("--pairs", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--score-only", action="store_true", help="omit end coordinates (no position tracking)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "dpx":
        args.warmup = 3

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    n_gpus = max(args.gpus, world)
    cores = os.cpu_count() or 1
    from dpx_gpu_genomics_project_b200 import synth

    config = {"workload": f"LinearSmithWaterman batch: {args.pairs} pairs x ({R_LEN}x{Q_LEN}) bp per GPU, score"
                          + ("" if args.score_only else " + end coords") + ", match 3 / mismatch -1 / gap -2",
              "pairs_per_gpu": args.pairs, "R": R_LEN, "Q": Q_LEN, "sharding": f"independent pairs x{n_gpus} (no collective)",
              "l2": "256 MiB memset between timed steps", "seed": "0x5EED0002 + rank"}

    # ---------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        blob, pairs = synth.uniform_blob_pairs(min(args.pairs, 400_000), R_LEN, Q_LEN, 0x5EED0002)
        n_s = cpu_sample_size(len(pairs), cores, seconds=8.0)
        vals, secs = [], []
        for _ in range(max(1, min(args.warmup, 1))):
            run_cpu_reference(blob, pairs, max(2000, n_s // 8), cores)
        kind = "reference"
        for _ in range(max(1, args.steps if args.steps <= 5 else 3)):
            v, kind, dt = run_cpu_reference(blob, pairs, n_s, cores)
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

    blob, pairs = synth.uniform_blob_pairs(args.pairs, R_LEN, Q_LEN, 0x5EED0002 + rank)
    cells = float(args.pairs) * R_LEN * Q_LEN
    flags = api.OUT_SCORE | (0 if args.score_only else api.OUT_END_COORDS)
    params = api.make_params(api.LSW, flags=flags, **WEIGHTS)

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
    kernel_ms = float(batch.stats()["fill_ms"])          # library's own event pair around the last fill kernel

    # result check of the timed configuration on a sample (outside the timed region)
    res = batch.fetch()

    # ---- e2e: one-call ABI, pinned host buffers, H2D + D2H inside ------------------------------------------
    e2e_steps = max(2, min(5, args.steps))
    pin_blob = torch.from_numpy(blob).pin_memory()
    pin_pairs = torch.from_numpy(pairs.view(np.int32)).pin_memory()
    nb, npairs_bytes = pin_blob.numel(), pin_pairs.numel() * 4
    blob_p = pin_blob.numpy()
    pairs_p = pin_pairs.numpy().view(api.PAIR_DTYPE)
    out_scores = torch.empty(args.pairs, dtype=torch.int32).pin_memory()
    out_rc = torch.empty((args.pairs, 2), dtype=torch.int32).pin_memory()
    import ctypes as C
    L = eng.L

    def e2e_once():
        st = L.dpx_align_batch(eng.ctx, C.byref(params), blob_p.ctypes.data, nb, pairs_p.ctypes.data, args.pairs,
                               out_scores.numpy().ctypes.data, out_rc.numpy().ctypes.data, None, None)
        if st != 0:
            raise RuntimeError(f"dpx_align_batch failed: {st} {L.dpx_last_error(eng.ctx)}")

    e2e_once()
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
    assert (out_scores.numpy() == res.scores).all(), "e2e and staged paths disagree"

    # ---- reduce over ranks -------------------------------------------------------------------------------
    if dist is not None:
        t = torch.tensor([ms_total, e2e_s, kernel_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s, kernel_ms = float(t[0]), float(t[1]), float(t[2])
    ms_per_step = ms_total / args.steps
    value = cells * world / (ms_per_step * 1e-3) / 1e9
    e2e_value = cells * world / e2e_s / 1e9

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- parity spot check against the oracle (outside every timed region) --------------------------------
    import oracle_lib as ol
    n_chk = 2000
    s_ref, e_ref, _ = ol.align_batch(ol.params(ol.LSW, **WEIGHTS), blob, pairs[:n_chk], strings=False, threads=min(cores, 16))
    parity_ok = bool((res.scores[:n_chk] == s_ref).all() and (args.score_only or (res.end_row_col[:n_chk] == e_ref).all()))

    # ---- roofline ---------------------------------------------------------------------------------------
    clocks = sampler.result()
    peaks, peaks_src = measured_peaks()
    f_mhz = clocks["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)
    alu_pc, issue_pc = sass_counts(track=not args.score_only)
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    kernel_gcups = cells / (ms_per_step * 1e-3) / 1e9         # per GPU, the step is one launch of the fill kernel
    roofline = {"bound": "dpx_issue", "achieved": kernel_gcups, "unit": "GCUPS", "kernel": "sr_lsw_kernel<G=8,K=19>",
                "traffic": None}
    if alu_pc:
        cells_clk_sm = min(64.0 / alu_pc, 128.0 / issue_pc)
        peak = cells_clk_sm * sms * f_mhz * 1e6 / 1e9
        roofline.update({"peak": peak, "frac": kernel_gcups / peak,
                         "model": {"alu_pipe_lanes_per_clk_per_sm": 64, "issue_lanes_per_clk_per_sm": 128,
                                   "alu_instr_per_cell": alu_pc, "issued_instr_per_cell": issue_pc,
                                   "sms": sms, "sm_mhz": f_mhz, "source": "profiles/sass_counts.json + profiles/r01_dpx_microbench*.json"}})
    # HBM view of the same launch: algorithmic bytes = 2-bit bases + the index entry + 12 B of results per pair
    algo_bytes = args.pairs * ((R_LEN + Q_LEN) * 0.25 + 16 + 8 + 12)
    hbm_ach = algo_bytes / (ms_per_step * 1e-3) / 1e9
    roofline["hbm"] = {"achieved": hbm_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / peaks["hbm_gbs"],
                       "peak_source": f"MEASURED_PEAKS.json ({peaks_src})", "algorithmic_bytes_per_launch": algo_bytes}

    out = {"metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "int16x2", "data": "synthetic", "config": config, "clocks": clocks,
           "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": int(nb + npairs_bytes),
                   "d2h_bytes_per_step": int(args.pairs * 12), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                   "api": "dpx_align_batch (C ABI), pinned host buffers",
                   "bare_h2d_copy_ms": h2d_ms, "bare_h2d_gbs": nb / (h2d_ms * 1e-3) / 1e9},
           "gpu_launches": launches_per_step * args.steps, "roofline": roofline,
           "kernel_ms_last_step": kernel_ms, "wall_s_timed_region": t_wall, "parity_spot_check": parity_ok}

    if not args.no_cpu_baseline and world == 1:
        n_s = cpu_sample_size(args.pairs, cores)
        v, kind, dt = run_cpu_reference(blob, pairs, n_s, cores)
        out["cpu_baseline"] = {"value": v, "unit": "GCUPS", "cores": cores, "kind": kind, "seconds": dt,
                               "sample": f"first {n_s} pairs of the same workload, {cores} host threads, align loop only"}
    print(json.dumps(out))
    batch.free(); eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
