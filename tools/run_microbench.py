"""Runs tools/dpx_microbench while sampling the SM clock through NVML; writes profiles/r02_dpx_microbench.json (the instruction
rates are in SM clocks read on the device; the NVML samples say which frequency those clocks ran at)."""
import json, os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_dpx_microbench.json")
samples, stop = [], threading.Event()
def sample():
    try:
        import pynvml as nv
        nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
        while not stop.is_set():
            samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)); time.sleep(0.02)
    except Exception as e:
        print("nvml:", e)
t = threading.Thread(target=sample, daemon=True); t.start()
log = subprocess.run([os.path.join(ROOT, "tools", "dpx_microbench"), out], capture_output=True, text=True)
stop.set(); t.join(timeout=1)
print(log.stdout); print(log.stderr, file=sys.stderr)
d = json.load(open(out))
s = sorted(samples)
d["nvml_sm_mhz"] = {"samples": len(s), "median": s[len(s) // 2] if s else None, "min": s[0] if s else None, "max": s[-1] if s else None}
json.dump(d, open(out, "w"), indent=1)
open(out.replace(".json", ".log"), "w").write(log.stdout)
print(d["nvml_sm_mhz"])
