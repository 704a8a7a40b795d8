"""Per-kernel SASS opcode histogram (whole kernel and hottest backward-branch loop).
usage: python tools/sass_loops.py <binary-or-.so> [kernel-name-regex]"""
import collections
import re
import subprocess
import sys


def kernels(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    cur = None
    out = collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); out[cur] = []; continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            out[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return out


def opcode(ins):
    toks = ins.split()
    if toks[0].startswith("@"):
        toks = toks[1:]
    return toks[0]


def loops(ins):
    """(start, end) address ranges of backward branches."""
    res = []
    for addr, text in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?(\w+)\)?", text)
        m2 = re.search(r"BRA.*0x([0-9a-f]+)", text)
        if m2:
            tgt = int(m2.group(1), 16)
            if tgt < addr:
                res.append((tgt, addr))
    return res


def main():
    path = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    for name, ins in kernels(path).items():
        if pat and not pat.search(name):
            continue
        print("=" * 100)
        print(name, f"({len(ins)} instructions)")
        lp = loops(ins)
        for (s, e) in sorted(lp, key=lambda t: t[1] - t[0]):
            body = [t for a, t in ins if s <= a <= e]
            hist = collections.Counter(opcode(t) for t in body)
            print(f"  loop 0x{s:04x}-0x{e:04x}: {len(body)} instr: " + ", ".join(f"{k} x{v}" for k, v in hist.most_common()))


if __name__ == "__main__":
    main()
