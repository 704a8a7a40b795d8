"""The C++ drop-in (host/main and the shim classes) prints exactly the reference driver's bytes."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "dpx_gpu_genomics_project_b200", "host")
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.run(["make", "-s", "-C", HOST], check=True)


@pytest.mark.parametrize("algo,flags", [("LNW", ["-open", "-2"]), ("LSW", ["-open", "-2"]), ("ANW", ["-open", "-3", "-extend", "-1"])])
@pytest.mark.parametrize("name", ["adversarial", "cfg1_small", "mid"])
def test_dropin_main_stdout_equals_reference_driver(algo, flags, name):
    path = os.path.join(GOLD, f"{name}.in.txt")
    out = subprocess.run([os.path.join(HOST, "main"), "-pairs", path, "-match", "3", "-mismatch", "-1"] + flags + ["-algo", algo],
                         check=True, capture_output=True).stdout
    lines = out.split(b"\n")
    # reference c++/main.cpp:153,165 header and :257,260 footer around the blocks
    assert lines[0] == b"Parsing input file: " + path.encode()
    assert lines[1] == b"Pair # | Score"
    assert lines[-3].startswith(b"Elapsed time (usec): ") and lines[-2] == b"Cleaning up" and lines[-1] == b""
    body = b"\n".join(lines[2:-3]) + b"\n"
    assert body == open(os.path.join(GOLD, f"{name}.{algo}.out.txt"), "rb").read()


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("algo,flags", [("LNW", ["-open", "-2"]), ("LSW", ["-open", "-2"]), ("ANW", ["-open", "-3", "-extend", "-1"])])
@pytest.mark.parametrize("name", ["cfg1_small", "mid", "adversarial"])
def test_dropin_main_on_several_devices_prints_reference_bytes(algo, flags, name):
    """-devices / -gpus: dpx_create_multi + dpx_multi_align_batch_text (one process, one worker per device, contiguous shards).
    On a one-GPU box the three workers share device 0 (same sharding / stitching code); with more GPUs every GPU takes a shard."""
    path = os.path.join(GOLD, f"{name}.in.txt")
    n = _ngpu()
    for sel in (["-devices", "0,0,0"], ["-gpus", str(n)]) if n > 1 else (["-devices", "0,0,0"],):
        out = subprocess.run([os.path.join(HOST, "main"), "-pairs", path, "-match", "3", "-mismatch", "-1"] + flags + ["-algo", algo] + sel,
                             check=True, capture_output=True).stdout
        lines = out.split(b"\n")
        body = b"\n".join(lines[2:-3]) + b"\n"
        assert body == open(os.path.join(GOLD, f"{name}.{algo}.out.txt"), "rb").read(), sel


@pytest.mark.parametrize("algo", ["LNW", "LSW", "ANW", "BSW"])
def test_reference_per_pair_loop_compiles_against_shims_and_matches(algo):
    """c++/main.cpp:237-252's loop (construct aligner, align()) against the shim classes; BSW with a full band == LSW."""
    path = os.path.join(GOLD, "adversarial.in.txt")
    out = subprocess.run([os.path.join(HOST, "example_per_pair"), algo, path, "40"], check=True, capture_output=True).stdout
    want = open(os.path.join(GOLD, f"adversarial.{'LSW' if algo == 'BSW' else algo}.out.txt"), "rb").read()
    assert want.startswith(out) and out.count(b"\n") == 160


def test_dropin_main_reports_parse_errors_like_the_reference(tmp_path):
    bad = tmp_path / "bad.txt"
    bad.write_bytes(b"0\n0123\n")
    r = subprocess.run([os.path.join(HOST, "main"), "-pairs", str(bad)], capture_output=True)
    assert r.returncode == 1 and b"Number of lines not a multiple of 3" in r.stderr
    r = subprocess.run([os.path.join(HOST, "main"), "-pairs", str(tmp_path / "nope.txt")], capture_output=True)
    assert r.returncode == 1 and b"Could not open file" in r.stderr


@pytest.mark.parametrize("name", ["adversarial", "mid"])
def test_dropin_main_long_flag_prints_the_same_blocks(name):
    """-long sends every pair through the checkpointed long-pair path (dpx_align_long_pair_strings): same bytes as the batch path."""
    path = os.path.join(GOLD, f"{name}.in.txt")
    out = subprocess.run([os.path.join(HOST, "main"), "-pairs", path, "-match", "3", "-mismatch", "-1", "-open", "-2", "-algo", "LSW", "-long"],
                         check=True, capture_output=True).stdout
    lines = out.split(b"\n")
    body = b"\n".join(lines[2:-3]) + b"\n"
    assert body == open(os.path.join(GOLD, f"{name}.LSW.out.txt"), "rb").read()


def test_dropin_main_reads_fasta_pairs(tmp_path):
    """-fastx: the same pairs as FASTA records (multi-line) give the golden blocks of the 3-line file."""
    src = open(os.path.join(GOLD, "mid.in.txt"), "rb").read().split(b"\n")
    fa = tmp_path / "mid.fa"
    with open(fa, "wb") as f:
        for k in range(len(src) // 3):
            for tag, seq in ((b"r", src[3 * k + 1]), (b"q", src[3 * k + 2])):
                f.write(b">" + tag + b"%d\n" % k + b"\n".join(seq[i:i + 50] for i in range(0, len(seq), 50)) + b"\n")
    out = subprocess.run([os.path.join(HOST, "main"), "-fastx", str(fa), "-match", "3", "-mismatch", "-1", "-open", "-2", "-algo", "LNW"],
                         check=True, capture_output=True).stdout
    lines = out.split(b"\n")
    body = b"\n".join(lines[2:-3]) + b"\n"
    assert body == open(os.path.join(GOLD, "mid.LNW.out.txt"), "rb").read()


def test_dropin_main_all_flag_prints_every_maximum(tmp_path):
    """-all: the stdout of the reference built with -DBACKTRACK_ALL (checked here against the oracle restatement pinned on that build)."""
    import oracle_lib as ol
    from dpx_gpu_genomics_project_b200 import synth
    pp = [(b"0101", b"1010"), (b"0123012", b"0123"), (b"000", b"111"), (b"00100", b"00"), (b"0101010101", b"01010"), (b"012012012", b"12")]
    img = synth.pairs_to_file_bytes(pp)
    path = tmp_path / "ties.txt"
    path.write_bytes(bytes(img))
    out = subprocess.run([os.path.join(HOST, "main"), "-pairs", str(path), "-match", "3", "-mismatch", "-1", "-open", "-2", "-algo", "LSW", "-all"],
                         check=True, capture_output=True).stdout
    lines = out.split(b"\n")
    body = b"\n".join(lines[2:-3]) + b"\n"
    blob, pairs = ol.parse_image(img)
    want, n = ol.lsw_all_text(ol.params(ol.LSW), blob, pairs)
    assert n == 11 and body == want
