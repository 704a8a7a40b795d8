"""CPU-side checks of the drop-in boundary: libdpxalign.so loads without a GPU, exports every symbol
include/dpxalign.h declares, parses the reference's input format, and fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as ol
from dpx_gpu_genomics_project_b200 import _lib, api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def L():
    _lib.build()
    return _lib.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dpxalign.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dpx_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(L):
    names = declared_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/dpxalign.h but not exported"


def test_abi_version_and_strerror(L):
    assert L.dpx_abi_version() == 2
    assert b"CPU fallback" in L.dpx_strerror(-2)
    assert L.dpx_strerror(0) == b"ok"


def test_struct_layouts_match_reference_seqpair():
    assert api.PAIR_DTYPE.itemsize == 16 and C.sizeof(_lib.Params) == 28


def test_parse_input_matches_reference_parser_semantics(L):
    for name in ("adversarial", "shapes", "cfg1_small"):
        path = os.path.join(GOLD, f"{name}.in.txt")
        got = api.parse_input(path)
        blob, pairs = ol.parse_image(open(path, "rb").read())
        assert (got.sequences == blob).all()
        assert (got.pairs == pairs).all()
        assert got.info["numPairs"] == len(pairs)
        assert got.info["numCells"] == int((pairs["referenceSize"].astype(np.int64) * pairs["querySize"]).sum())
        assert got.info["maxReferenceLength"] == int(pairs["referenceSize"].max())
        assert got.info["minQueryLength"] == int(pairs["querySize"].min())


def test_parse_input_errors(tmp_path, L):
    with pytest.raises(api.DpxError) as e:
        api.parse_input(str(tmp_path / "missing.txt"))
    assert e.value.status == -5
    bad = tmp_path / "bad.txt"
    bad.write_bytes(b"0\n0123\n")            # 2 lines: not a multiple of 3 (c++/parseInput.cpp:38-41 exits)
    with pytest.raises(api.DpxError) as e:
        api.parse_input(str(bad))
    assert e.value.status == -6
    empty = tmp_path / "empty.txt"
    empty.write_bytes(b"")
    assert api.parse_input(str(empty)).info["numPairs"] == 0


@pytest.mark.skipif(_lib.lib().dpx_device_count() > 0 if os.path.exists(_lib.LIB_PATH) else False, reason="GPU present")
def test_no_cpu_fallback_without_device(L):
    with pytest.raises(api.DpxError) as e:
        api.Engine(0)
    assert e.value.status == -2


def test_synth_generators_are_deterministic_and_well_formed():
    a = synth.uniform_file_bytes(50, 150, 150, 0x5EED0002)
    b = synth.uniform_file_bytes(50, 150, 150, 0x5EED0002)
    assert (a == b).all()
    blob, pairs = ol.parse_image(a)
    assert len(pairs) == 50 and (pairs["referenceSize"] == 150).all() and (pairs["querySize"] == 150).all()
    assert set(np.unique(blob)) <= {0, 48, 49, 50, 51}
    m = synth.mutated_fixed_file_bytes(20, 300, 300, 7, 0.02, 0.005, 0.005)
    blob, pairs = ol.parse_image(m)
    assert len(pairs) == 20 and (pairs["querySize"] == 300).all()
    # mutated queries stay similar to their reference: SW score far above the random-pair level
    s, _, _ = ol.align_batch(ol.params(ol.LSW), blob, pairs, strings=False)
    assert s.min() > 300


def _pairs_of(p):
    blob = p.sequences.tobytes()
    return [(blob[int(r["referenceIdx"]): int(r["referenceIdx"]) + int(r["referenceSize"])],
             blob[int(r["queryIdx"]): int(r["queryIdx"]) + int(r["querySize"])]) for r in p.pairs]


def test_fasta_and_fastq_front_end(tmp_path, L):
    """dpx_parse_fastx: multi-line FASTA, CRLF, FASTQ with '@' / '>' inside the qualities, one interleaved file or two files."""
    from dpx_gpu_genomics_project_b200 import api, synth
    rng = synth.Rng(3)
    want = []
    for k in range(7):
        r = synth.random_seq(rng, 50 + 37 * k, b"ACGT")
        want.append((r, synth.mutate(rng, r, 0.1, 0.05, 0.05, b"ACGT")))
    want.append((b"", b"ACGT")); want.append((b"N", b""))

    def fasta(seq, name, width=60, eol=b"\n"):
        lines = [seq[i:i + width] for i in range(0, len(seq), width)] or [b""]
        return b">" + name + eol + eol.join(lines) + eol

    def fastq(seq, name):
        qual = (b"@>+I" * (len(seq) // 4 + 1))[:len(seq)]
        return b"@" + name + b"\n" + seq + b"\n+\n" + qual + b"\n"

    one = tmp_path / "pairs.fa"
    one.write_bytes(b"\n" + b"".join(fasta(r, b"ref%d some description" % k) + fasta(q, b"qry%d" % k, 13, b"\r\n") for k, (r, q) in enumerate(want)))
    p = api.parse_fastx(str(one))
    assert _pairs_of(p) == want
    assert p.info["numPairs"] == len(want) and p.info["maxReferenceLength"] == max(len(r) for r, _ in want)
    assert p.info["numCells"] == sum(len(r) * len(q) for r, q in want)

    a, b = tmp_path / "refs.fq", tmp_path / "reads.fq"
    a.write_bytes(b"".join(fastq(r, b"r%d" % k) for k, (r, _) in enumerate(want)))
    b.write_bytes(b"".join(fastq(q, b"q%d/1" % k) for k, (_, q) in enumerate(want)))
    assert _pairs_of(api.parse_fastx(str(a), str(b))) == want
    # mixed: FASTA references, FASTQ reads
    a2 = tmp_path / "refs.fa"
    a2.write_bytes(b"".join(fasta(r, b"r%d" % k, 70) for k, (r, _) in enumerate(want)))
    assert _pairs_of(api.parse_fastx(str(a2), str(b))) == want

    bad = tmp_path / "odd.fa"
    bad.write_bytes(fasta(b"ACGT", b"only"))
    with pytest.raises(api.DpxError):
        api.parse_fastx(str(bad))                          # a reference without its query
    with pytest.raises(api.DpxError):
        api.parse_fastx(str(a2), str(bad))                 # record counts differ
    bad.write_bytes(b"ACGT\n")
    with pytest.raises(api.DpxError):
        api.parse_fastx(str(bad))                          # neither '>' nor '@'
    bad.write_bytes(b"@r\nACGT\n+\nII\n")
    with pytest.raises(api.DpxError):
        api.parse_fastx(str(bad))                          # truncated qualities
    with pytest.raises(api.DpxError):
        api.parse_fastx(str(tmp_path / "missing.fa"))
