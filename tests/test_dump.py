"""Dump / replay formats (rovaca_b200/dump.py): the reference's 7-line kernel dump (pairhmm_internal.cpp:493-503), GKL-style
test-vector lines, and the binary region dump the C++ adapter writes.  CPU: parsing and round trips.  GPU: a dump of the golden
pairs replays to the golden bits, and what the adapter captured replays to what the adapter returned."""
import io
import os

import numpy as np
import pytest

import _golden
from oracle import bindings as ob
from rovaca_b200 import dump, workloads as wl


def _same(a, b):
    return all(np.array_equal(a[k], b[k]) for k in ("hap", "rs", "q", "ins", "dele", "gcp"))


def test_pair_text_round_trips():
    gp = _golden.pairs()[:40]
    bits = [int(p["ref_avx512"]) for p in gp]
    for style in ("line", "record"):
        back = dump.parse_pairs(dump.format_pairs(gp, style, bits if style == "line" else None))
        assert len(back) == len(gp) and all(_same(x, y) for x, y in zip(gp, back))
        if style == "line":
            assert [p["expected_bits"] for p in back] == bits
    # the reference's own print format, written out by hand: "<rslen> \t <haplen>", hap, read, q+33, ins+33, del+33, gcp+33
    text = "4 \t 6\nACGTAC\nACGT\n" + "".join(chr(q + 33) for q in (30, 6, 40, 93)) + "\n" + "N" * 4 + "\n" + "N" * 4 + "\n" + "+" * 4 + "\n"
    (p,) = dump.parse_pairs(text.replace("\\t", "\t"))
    assert bytes(p["hap"]) == b"ACGTAC" and list(p["q"]) == [30, 6, 40, 93] and list(p["ins"]) == [45] * 4 and list(p["gcp"]) == [10] * 4
    with pytest.raises(ValueError):
        dump.parse_pairs("ACGT ACG III NNN NNN")
    with pytest.raises(ValueError):
        dump.parse_pairs("ACGT ACG III NN NNN +++")
    assert dump.parse_pairs("ACGT ACG III NNN NNN +++ -3.25")[0]["expected"] == -3.25


def test_region_dump_round_trip():
    b = wl.cfg5(5, seed=9)
    buf = io.BytesIO()
    for g in range(b.n_regions):
        reads, haps = b.region(g)
        dump.write_region(buf, reads, haps, 18 + g, g % 4)
    buf.seek(0)
    back = dump.read_regions(buf)
    assert len(back) == b.n_regions
    for g, (reads, haps, thr, pcr) in enumerate(back):
        r0, h0 = b.region(g)
        assert (thr, pcr) == (18 + g, g % 4)
        assert all(np.array_equal(x["bases"], y["bases"]) and np.array_equal(x["qual"], y["qual"]) and x["mapq"] == y["mapq"] for x, y in zip(r0, reads))
        assert all(np.array_equal(x, y) for x, y in zip(h0, haps))
    with pytest.raises(ValueError):
        dump.read_regions(io.BytesIO(b"garbage!" + bytes(32)))


@pytest.mark.gpu
def test_replay_golden_pair_dump(engine):
    gp = _golden.pairs()
    text = dump.format_pairs(gp, "line", [int(p["ref_avx512"]) for p in gp])
    pairs = dump.parse_pairs(text)
    raw, lk = dump.replay_pairs(engine, pairs)
    assert np.array_equal(raw.view(np.uint32), np.array([p["expected_bits"] for p in pairs], np.uint32))
    orc = ob.oracle()
    for p, r, v in list(zip(gp, raw, lk))[:64]:
        r64 = orc.pair_f64(p["hap"], p["rs"], p["q"], p["ins"], p["dele"], p["gcp"]) if r < np.float32(1e-28) else 0.0
        assert v == orc.log10_likelihood(float(r), r64) or (np.isinf(v) and np.isinf(orc.log10_likelihood(float(r), r64)))


@pytest.mark.gpu
def test_adapter_dump_replays(engine, tmp_path):
    """ROVACA_B200_DUMP: the regions the adapter saw, replayed through the Python door, give the adapter's own results."""
    path = tmp_path / "regions.bin"
    os.environ["ROVACA_B200_DUMP"] = str(path)
    try:
        adapter = ob.adapter()
        if adapter is None:
            pytest.skip("oracle/_ref/libphmm_adapter_test.so not built (needs /root/reference)")
        adapter.init()
        rng = np.random.default_rng(41)
        want = []
        for _ in range(3):
            reads, haps = wl._one_region(rng, 20, 4, 200, 300, 80, 151, junk=0.1)
            want.append(adapter.region(reads, haps, 18, ob.PCR_CONSERVATIVE))
        adapter.shutdown()
    finally:
        del os.environ["ROVACA_B200_DUMP"]
    regs = dump.read_regions(open(path, "rb"))
    assert len(regs) == 3
    for (reads, haps, thr, pcr), (m0, k0) in zip(regs, want):
        m, k = engine.call_pairhmm(haps, reads, thr, pcr)
        assert np.array_equal(k, k0) and np.array_equal(m.view(np.uint64), m0.view(np.uint64))
