"""CPU: the Smith-Waterman restatement (oracle/sw_oracle.c) against the golden vectors of the reference's sw_avx.c and, where
oracle/_ref is present, against the reference itself (AVX-512 and AVX2 builds) on fresh random alignments."""
import numpy as np
import pytest

import _golden
from oracle import bindings as ob

B = np.frombuffer(b"ACGT", np.uint8)


def test_sw_oracle_matches_golden(orc):
    for k, c in enumerate(_golden.sw()):
        for name, prm in (("realign", ob.SW_REALIGN), ("assemble", ob.SW_ASSEMBLE)):
            cig, off, score = orc.sw_align(c["hap"], c["read"], prm)
            want_cig, want_off, want_score = c[name]
            assert np.array_equal(cig, want_cig) and off == want_off and score == want_score, (k, name, ob.cigar_string(cig), ob.cigar_string(want_cig))
            # a CIGAR consumes the whole read; its reference span starts at the offset and stays inside the haplotype
            q = sum(int(x) >> 4 for x in cig if (int(x) & 15) in (0, 1, 4))
            r = sum(int(x) >> 4 for x in cig if (int(x) & 15) in (0, 2))
            assert q == len(c["read"]) and 0 <= off and off + r <= len(c["hap"])


def test_sw_oracle_known_answers(orc):
    cig, off, score = orc.sw_align(b"TTTTACGTACGTACGTTTTT", b"ACGTACGTACGT", ob.SW_REALIGN)
    assert ob.cigar_string(cig) == "12M" and off == 4 and score == 120
    cig, off, _ = orc.sw_align(b"GGGGACGTACGTTTGACCGGGG", b"ACGTACGTGACC", ob.SW_ASSEMBLE)      # TT deleted from the read
    assert ob.cigar_string(cig) == "7M2D5M" and off == 4                                        # (the reference's own answer: ties go left)
    cig, off, _ = orc.sw_align(b"GGGGACGTACGTGACCGGGG", b"ACGTACGTAAGACC", ob.SW_ASSEMBLE)      # AA inserted into the read
    assert ob.cigar_string(cig) == "8M2I4M" and off == 4
    cig, off, _ = orc.sw_align(b"ACGTACGTAC", b"CCCCCCACGTACGTAC", ob.SW_REALIGN)                # read hangs over the start
    assert ob.cigar_string(cig) == "6S10M" and off == 0
    cig, off, score = orc.sw_align(B[np.zeros(2048, int)], b"ACGT", ob.SW_REALIGN)               # the reference gives up at 2048
    assert cig is None and off == 1


def test_sw_oracle_vs_reference(orc, ref):
    rng = np.random.default_rng(5)
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)      # the case generator is shared with the golden file
    spec.loader.exec_module(mg)
    for hap, rd in mg.sw_cases(rng, 400):
        for asm, prm in ((False, ob.SW_REALIGN), (True, ob.SW_ASSEMBLE)):
            want = ref.sw_align(hap, rd, asm)
            got = orc.sw_align(hap, rd, prm)
            assert np.array_equal(got[0], want[0]) and got[1:] == want[1:], (len(hap), len(rd), asm)
            avx2 = ref.sw_align(hap, rd, asm, isa=0)
            assert np.array_equal(avx2[0], want[0]) and avx2[1:] == want[1:]
