"""Writes the SASS of every fill kernel's hot loop (innermost backward-branch loop that contains DPX instructions) from the
built libdpxalign.so to profiles/sass/<kernel>.sass, with a per-opcode histogram header.  These are the listings the design
notes cite for instruction counts per cell (tools/sass_counts.py condenses the same loops into profiles/sass_counts.json).
usage: python tools/sass_dump.py [libdpxalign.so] [outdir]"""
import collections
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_loops import kernels, loops, opcode  # noqa: E402

WANT = [r"sr_lsw_kernelILi8ELi19ELb1ELb0E", r"sr_lsw_kernelILi8ELi19ELb0ELb0E", r"pw_nw_kernelILi1ELb1ELi8ELb1ELb0ELb0E", r"pw_nw_kernelILi0ELb1ELi8ELb1ELb0ELb0E",
        r"pw_nw_kernelILi2ELb1ELi8ELb1ELb0ELb0E", r"pw_nw_kernelILi1ELb1ELi8ELb0ELb1ELb0E", r"band_sw_kernelILi2ELb1ELb1E", r"band_sw_kernelILi2ELb1ELb0E", r"long_sw_kernelILi16ELb1ELb1E",
        r"long_sw_kernelILi8ELb1ELb1E", r"wf_fill_kernelILi1ELb1ELi8E"]


def demangle(name):
    try:
        return subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else "dpx_gpu_genomics_project_b200/libdpxalign.so"
    out = sys.argv[2] if len(sys.argv) > 2 else "profiles/sass"
    os.makedirs(out, exist_ok=True)
    for name, ins in kernels(lib).items():
        if not any(re.search(w, name) for w in WANT):
            continue
        best = None
        for (s, e) in loops(ins):
            body = [(a, t) for a, t in ins if s <= a <= e]
            dpx = sum(1 for _, t in body if re.match(r"(@!?U?P\d+\s+)?VI(ADD)?MNMX", t))
            if dpx and (best is None or len(body) < len(best)):
                best = body
        if best is None:
            continue
        hist = collections.Counter(opcode(t) for _, t in best)
        short = re.sub(r"^_ZN3dpx\d+", "", name)
        short = re.sub(r"EEvNS_.*$", "", short).replace("ILi", "_").replace("ELi", "_").replace("ELb", "_")
        with open(os.path.join(out, short + ".sass"), "w") as f:
            f.write(f"// {demangle(name)}\n// hot loop: {len(best)} instructions\n// " +
                    ", ".join(f"{k} x{v}" for k, v in hist.most_common()) + "\n")
            for a, t in best:
                f.write(f"/*{a:04x}*/  {t}\n")
        print(short, len(best))


if __name__ == "__main__":
    main()
