"""Counts the instructions of each fill kernel's hot loop in the built libdpxalign.so and turns them into clocks of an SM
sub-partition per loop trip with the measured model of profiles/r02_dpx_microbench.json (r02, second version):
  * ALU pipe: 2 clocks per warp instruction -- DPX (VIMNMX3*, VIADDMNMX*), PRMT, SHF, LEA, LOP3, ISETP, SEL all run at 63.9
    lanes/clk/SM, alone and in any mix ("1 PRMT + 1 VIADDMNMX.S16x2", "1 VIADDMNMX.S16x2 + 1 LOP3 (volatile)": 63.9 for the pair
    = 4 clocks; the first LOP3 measurement had been folded by the compiler) -- except plain adds: "1 VIADDMNMX + 1 VIADD" = 85.3
    lanes/clk/SM = 3 clocks per pair, so VIADD / IADD3 count 1;
  * FMA pipe: IMAD* 2 clocks (63.9 lanes/clk/SM);
  * issue: 1 clock per warp instruction.
A loop trip cannot take less than the largest of the three; bench.py turns that into the roofline (I_cell counted
from the committed SASS inner loop).  The two pipes do not overlap perfectly: next to packed DPX instructions an IMAD costs 0.3 - 0.9 clocks
("1 VIADDMNMX.S16x2 + 1 IMAD" = 100.5 lanes/clk/SM = 2.55 clocks per pair); tools/sr_rowmix_bench.cu measures that for the
short-read kernel's exact mix (8.9 clocks per row of 7 + 1 ALU-pipe clocks and one IMAD).  The model leaves it out, so `frac`
is against an upper bound.
Output: profiles/sass_counts.json.

usage: python tools/sass_counts.py [libdpxalign.so] [out.json]"""
import collections
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_loops import kernels, loops, opcode  # noqa: E402

FMA_PIPE = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "HADD2", "HMUL2")
OTHER = ("LDS", "STS", "LDG", "STG", "LDC", "LDCU", "SHFL", "BRA", "ATOM", "RED", "BAR", "S2R", "S2UR", "CS2R", "NOP",
         "BSSY", "BSYNC", "EXIT", "MEMBAR", "ERRBAR", "CCTL", "WARPSYNC", "YIELD", "VOTE", "VOTEU", "R2UR", "REDUX", "CREDUX",
         "UIADD3", "UISETP", "UMOV", "ULEA", "ULOP3", "USHF", "UIMAD", "USEL", "UPRMT", "UFLO", "UPOPC", "LDSM", "MATCH", "CALL", "RET")


FULL_RATE_ALU = ("VIADD", "IADD3", "IADD")


def pipe(op):
    base = op.split(".")[0]
    if base in FMA_PIPE:
        return "fma"
    if base in OTHER:
        return "other"
    return "alu"


def alu_clocks(op):
    """Clocks of the ALU pipe one warp instruction takes (module docstring)."""
    base = op.split(".")[0]
    if pipe(op) != "alu":
        return 0
    return 1 if base in FULL_RATE_ALU else 2


# kernel-name regex -> (label, cells per inner-loop trip as a function of template ints)
SPECS = [
    (r"sr_lsw_kernelILi(\d+)ELi(\d+)ELb([01])ELb([01])E", "shortread_s16x2",
     # one loop trip = two column steps of K rows, two pairs per slot (WIDE: one pair)
     lambda g: dict(G=int(g[0]), K=int(g[1]), track=bool(int(g[2])), wide=bool(int(g[3])), cells=(1 if int(g[3]) else 2) * 2 * int(g[1]))),
    (r"pw_nw_kernelILi(\d+)ELb([01])ELi(\d+)ELb1ELb0ELb0E", "pairwf_s16x2",
     lambda g: dict(algo=int(g[0]), traceback=bool(int(g[1])), K=int(g[2]), cells=2 * 2 * int(g[2]))),
    (r"pw_nw_kernelILi(\d+)ELb([01])ELi(\d+)ELb0ELb0ELb0E", "pairwf_s32",
     lambda g: dict(algo=int(g[0]), traceback=bool(int(g[1])), K=int(g[2]), cells=2 * int(g[2]))),
    (r"band_sw_kernelILi(\d+)ELb([01])ELb([01])E", "band_s32",
     # one loop trip = one traceback word = SPW super-steps (8 / 4 / 2 for M = 1 / 2 / 3) of 2M (+1) slots
     lambda g: dict(M=int(g[0]), extra=bool(int(g[1])), traceback=bool(int(g[2])), cells={1: 8, 2: 4, 3: 2}[int(g[0])] * (2 * int(g[0]) + int(g[1])))),
    (r"wf_fill_kernelILi(\d+)ELb([01])ELi(\d+)E", "wavefront_s32",
     lambda g: dict(algo=int(g[0]), traceback=bool(int(g[1])), K=int(g[2]), cells=int(g[2]))),
    # long-pair chain: the steady row loop is unrolled 4 row steps of K cells (several overlapping back-edges: pick the densest loop)
    (r"long_sw_kernelILi(\d+)ELb([01])ELb([01])ELb([01])E", "long_s32",
     lambda g: dict(K=int(g[0]), pack=bool(int(g[1])), table=bool(int(g[2])), ck=bool(int(g[3])), cells=4 * int(g[0]), largest=True)),
]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else "dpx_gpu_genomics_project_b200/libdpxalign.so"
    out = sys.argv[2] if len(sys.argv) > 2 else "profiles/sass_counts.json"
    res = {}
    for name, ins in kernels(lib).items():
        for pat, label, fn in SPECS:
            m = re.search(pat, name)
            if not m:
                continue
            info = fn(m.groups())
            # hot loop = the innermost loop with the most DPX (VIMNMX*/VIADDMNMX*) or max instructions
            best = None
            cand = []
            for (s, e) in loops(ins):
                body = [t for a, t in ins if s <= a <= e]
                dpx = sum(1 for t in body if re.match(r"(@!?U?P\d+\s+)?VI(ADD)?MNMX", t))
                if dpx:
                    cand.append((s, e, body, dpx))
            if info.get("largest"):                           # the unrolled steady loop = the loop with the highest DPX density
                cand = [max(cand, key=lambda c: c[3] / len(c[2]))] if cand else []
            for (s, e, body, dpx) in cand:
                if best is None or (len(body) > len(best[0]) if info.get("largest") else len(body) < len(best[0])):
                    best = (body, dpx)
            if best is None:
                continue
            body, dpx = best
            hist = collections.Counter(opcode(t) for t in body)
            by_pipe = collections.Counter()
            for op, c in hist.items():
                by_pipe[pipe(op)] += c
            cells = info.pop("cells"); info.pop("largest", None)
            alu_clk = sum(alu_clocks(op) * c for op, c in hist.items())
            fma_clk = 2 * by_pipe["fma"]
            key = label + ":" + ",".join(f"{k}={v}" for k, v in info.items())
            res[key] = dict(mangled=name, loop_instructions=len(body), cells_per_trip=cells, dpx_instructions=dpx,
                            alu_pipe=by_pipe["alu"], fma_pipe=by_pipe["fma"], other=by_pipe["other"],
                            alu_per_cell=by_pipe["alu"] / cells, fma_per_cell=by_pipe["fma"] / cells,
                            issue_per_cell=len(body) / cells,
                            alu_clk_per_cell=alu_clk / cells, fma_clk_per_cell=fma_clk / cells,
                            half_rate_alu_per_cell=sum(c for op, c in hist.items() if alu_clocks(op) == 2) / cells,
                            bound_clk_per_cell=max(alu_clk, fma_clk, len(body)) / cells,
                            histogram=dict(hist.most_common()), **info)
    with open(out, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    for k, v in sorted(res.items()):
        print(f"{k:70s} clk/cell: issue {v['issue_per_cell']:.2f}  alu {v['alu_clk_per_cell']:.2f}  fma {v['fma_clk_per_cell']:.2f}")


if __name__ == "__main__":
    main()
