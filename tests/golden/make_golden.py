"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libphmm_ref.so, i.e. the
reference's own kernels, host wrapper and ReadRecord compiled in place from /root/reference).

Run in the build container (needs /root/reference and an AVX-512 CPU):
    make -C oracle && python tests/golden/make_golden.py
The reference ships no PairHMM vectors of its own, so these files are what pins parity:
  pairs.npz    per-pair raw FP32 sums from core/avx512_s.cpp, core/avx2_s.cpp and rovaca/avx_512_float.cc
  regions.npz  whole-region outputs of call_old_pairhmm / call_new_pairhmm002 (matrix bits, kept reads)
  pcr.npz      apply_pcr_error_model / find_tandem_repeat_units outputs
  sw.npz       score / alignment offset / CIGAR of common/base/lib/sw_avx/sw_avx.c (the step after the path)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as ob          # noqa: E402
from rovaca_b200 import workloads as wl    # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
B = np.frombuffer(b"ACGT", np.uint8)


def mkpair(rng, R, H, mut, generic, orc):
    hap = B[rng.integers(0, 4, H)]
    if H >= R:
        off = int(rng.integers(0, H - R + 1))
        rs = hap[off:off + R].copy()
    else:
        rs = B[rng.integers(0, 4, R)]
    m = rng.random(R) < mut
    rs[m] = B[rng.integers(0, 4, int(m.sum()))]
    if rng.random() < 0.3:
        rs[int(rng.integers(0, R))] = ord("N")
    if rng.random() < 0.2:
        hap = hap.copy()
        hap[int(rng.integers(0, H))] = ord("N")
    q = rng.integers(2, 41, R).astype(np.uint8)
    if generic:
        ins = rng.integers(10, 46, R).astype(np.uint8)
        dele = rng.integers(10, 46, R).astype(np.uint8)
        gcp = rng.integers(5, 30, R).astype(np.uint8)
    else:
        ins = orc.apply_pcr_error_model(rs, np.full(R, 45, np.uint8), ob.PCR_CONSERVATIVE)
        dele = ins
        gcp = np.full(R, 10, np.uint8)
    return hap, rs, q, ins, dele, gcp


def main():
    ref = ob.ref()
    orc = ob.oracle()
    assert ref is not None and ref.avx512, "needs oracle/_ref built and an AVX-512 host"
    rng = np.random.default_rng(20261018)

    # ---- pairs --------------------------------------------------------------------------------
    shapes = [(150, 300), (7, 40), (100, 200), (250, 1000), (33, 41), (16, 16), (17, 300), (1, 5), (151, 500), (41, 41),
              (31, 64), (32, 64), (159, 310), (160, 310), (255, 300), (1, 1), (64, 10), (383, 400)]
    blobs = {k: [] for k in ("hap", "rs", "q", "ins", "del", "gcp")}
    lens, f512, f2, f1xn = [], [], [], []
    for (R, H) in shapes:
        for it in range(8):
            p = mkpair(rng, R, H, [0.0, 0.01, 0.05, 0.3][it % 4], it % 2 == 0, orc)
            for k, a in zip(blobs, p):
                blobs[k].append(a)
            lens.append((R, H))
            f512.append(ref.pair_f32(1, *p))
            f2.append(ref.pair_f32(0, *p))
            f1xn.append(ref.one_by_n_f32(p[0], [p[1:]])[0])
    np.savez_compressed(os.path.join(HERE, "pairs.npz"), lens=np.array(lens, np.uint32),
                        **{k: np.concatenate(v) for k, v in blobs.items()},
                        ref_avx512=np.array(f512, np.float32).view(np.uint32), ref_avx2=np.array(f2, np.float32).view(np.uint32),
                        ref_1xn=np.array(f1xn, np.float32).view(np.uint32))

    # ---- regions ------------------------------------------------------------------------------
    specs = [dict(n_reads=40, n_haps=6, hap_lo=300, hap_hi=300, read_lo=150, read_hi=150, junk=0.05),
             dict(n_reads=20, n_haps=3, hap_lo=120, hap_hi=120, read_lo=30, read_hi=100, junk=0.2),
             dict(n_reads=70, n_haps=9, hap_lo=250, hap_hi=450, read_lo=100, read_hi=151, junk=0.05),
             dict(n_reads=33, n_haps=2, hap_lo=60, hap_hi=60, read_lo=10, read_hi=41, junk=0.1),
             dict(n_reads=100, n_haps=12, hap_lo=400, hap_hi=400, read_lo=140, read_hi=151, junk=0.1),
             dict(n_reads=5, n_haps=1, hap_lo=50, hap_hi=50, read_lo=20, read_hi=20, junk=0.0),
             dict(n_reads=48, n_haps=5, hap_lo=1000, hap_hi=1000, read_lo=250, read_hi=250, junk=0.05),
             dict(n_reads=64, n_haps=8, hap_lo=300, hap_hi=300, read_lo=150, read_hi=150, junk=0.0)]
    regions, meta, mats, kepts = [], [], [], []
    for si, sp in enumerate(specs):
        for model in (ob.PCR_CONSERVATIVE, ob.PCR_NONE, ob.PCR_HOSTILE):
            reads, haps = wl._one_region(rng, **sp)
            if model == ob.PCR_HOSTILE and si % 2 == 0:     # homopolymer / STR rich reads for the PCR model
                for r in reads[:: 3]:
                    n = len(r["bases"])
                    u = B[rng.integers(0, 4, int(rng.integers(1, 5)))]
                    s = int(rng.integers(0, max(1, n - 30)))
                    r["bases"][s:s + 30] = np.tile(u, 30)[: min(30, n - s)]
            m_old, k_old = ref.region(reads, haps, 18, model, 0)
            m_new, k_new = ref.region(reads, haps, 18, model, 1)
            assert np.array_equal(m_old.view(np.uint64), m_new.view(np.uint64)) and np.array_equal(k_old, k_new)
            regions.append((reads, haps))
            meta.append((18, model))
            mats.append(m_old.reshape(-1).view(np.uint64))
            kepts.append(k_old)
    batch = wl.RegionBatch.from_regions(regions)
    np.savez_compressed(os.path.join(HERE, "regions.npz"), meta=np.array(meta, np.int32),
                        reg_read_off=batch.reg_read_off, reg_hap_off=batch.reg_hap_off, read_len=batch.read_len,
                        read_mapq=batch.read_mapq, read_off=batch.read_off, read_bases=batch.read_bases, read_qual=batch.read_qual,
                        hap_len=batch.hap_len, hap_off=batch.hap_off, hap_bases=batch.hap_bases,
                        mats=np.concatenate(mats), kept=np.concatenate(kepts), n_kept=np.array([len(k) for k in kepts], np.uint32))

    # ---- PCR model ----------------------------------------------------------------------------
    seqs, slen, outs, reps = [], [], [], []
    for it in range(120):
        n = int(rng.integers(1, 200))
        kind = it % 3
        if kind == 0:
            b = B[rng.integers(0, 4, n)]
        elif kind == 1:
            u = B[rng.integers(0, 4, int(rng.integers(1, 9)))]
            b = np.tile(u, n // len(u) + 1)[:n].copy()
            m = rng.random(n) < 0.05
            b[m] = B[rng.integers(0, 4, int(m.sum()))]
        else:
            b = B[rng.integers(0, 2, n)]
        seqs.append(b)
        slen.append(n)
        for model in (1, 2, 3):
            outs.append(ref.apply_pcr_error_model(b, np.full(n, 45, np.uint8), model))
        reps.append(np.array([ref.find_tandem_repeat_units(b, o) for o in range(n)], np.int32))
    np.savez_compressed(os.path.join(HERE, "pcr.npz"), seq=np.concatenate(seqs), seq_len=np.array(slen, np.uint32),
                        capped=np.concatenate(outs), repeats=np.concatenate(reps))
    print("golden vectors written:", len(lens), "pairs,", len(regions), "regions,", len(seqs), "PCR sequences")


def sw_cases(rng, n=160):
    """(haplotype, read) pairs for the Smith-Waterman step after the path: true sub-sequences with SNPs / indels / overhangs,
    unrelated sequences, low-complexity repeats (score ties), and the edge lengths."""
    out = []
    for t in range(n):
        H, R = int(rng.integers(1, 420)), int(rng.integers(1, 200))
        hap = B[rng.integers(0, 4, H)]
        kind = t % 5
        if kind == 4 or H < 8:
            rd = B[rng.integers(0, 4, R)]
        elif kind == 3:
            hap = np.tile(B[rng.integers(0, 4, int(rng.integers(1, 4)))], 210)[:H]
            rd = hap[: max(1, min(R, H))].copy()
            if len(rd) > 4:
                rd = np.delete(rd, int(rng.integers(1, len(rd) - 1)))
        else:
            o = int(rng.integers(0, H - 4))
            rd = hap[o:o + R].copy()
            m = rng.random(len(rd)) < 0.03
            rd[m] = B[rng.integers(0, 4, int(m.sum()))]
            if kind >= 1 and len(rd) > 10:
                p, k = int(rng.integers(1, len(rd) - 1)), int(rng.integers(1, 12))
                rd = np.concatenate([rd[:p], B[rng.integers(0, 4, k)], rd[p:]]) if rng.random() < 0.5 else np.concatenate([rd[:p], rd[p + k:]])
            if kind == 2:
                rd = np.concatenate([B[rng.integers(0, 4, int(rng.integers(1, 10)))], rd, B[rng.integers(0, 4, int(rng.integers(0, 6)))]])
        out.append((hap, rd))
    return out


def make_sw():
    """tests/golden/sw.npz: score, alignment offset and CIGAR of the UNMODIFIED reference Smith-Waterman
    (common/base/lib/sw_avx/sw_avx.c, compiled in place) for both parameter sets."""
    ref = ob.ref()
    assert ref is not None, "needs oracle/_ref (make -C oracle with /root/reference present)"
    rng = np.random.default_rng(20261018)
    cases = sw_cases(rng)
    rec = dict(n=np.array([len(cases)]))
    for k, (hap, rd) in enumerate(cases):
        rec[f"hap{k}"], rec[f"read{k}"] = hap, rd
        for tag, asm in (("r", False), ("a", True)):
            cig, off, score = ref.sw_align(hap, rd, asm)
            rec[f"cig_{tag}{k}"], rec[f"meta_{tag}{k}"] = cig, np.array([off, score], np.int64)
    np.savez_compressed(os.path.join(HERE, "sw.npz"), **rec)
    print("sw.npz:", len(cases), "cases")


def make_alleles():
    """tests/golden/alleles.npz: what the UNMODIFIED reference's AlleleLikelihoods template (genotype/allele_likelihoods.hpp,
    instantiated in place by oracle/al_shim.cpp) answers for best_alleles_breaking_ties and marginalize on call_pairhmm's
    matrices: regions with duplicated haplotypes (ties), random priorities, alleles that map to nothing."""
    from rovaca_b200 import workloads as wl
    ref, orc = ob.ref(), ob.oracle()
    assert ref is not None, "needs oracle/_ref (make -C oracle with /root/reference present)"
    rng = np.random.default_rng(20261019)
    rec, n_cases = {}, 40
    for k in range(n_cases):
        reads, haps = wl._one_region(rng, int(rng.integers(1, 60)), int(rng.integers(1, 9)), 100, 150, 30, 90, junk=0.15)
        haps = list(haps)
        for _ in range(int(rng.integers(0, 4))):
            haps.insert(int(rng.integers(0, len(haps) + 1)), haps[int(rng.integers(0, len(haps)))].copy())
        m, kept = ref.region(reads, haps)                 # the reference's own call_pairhmm
        n = len(kept)
        pri = rng.integers(-3, 3, len(haps)).astype(np.float64)
        best, second, bl, sl, conf = ref.best_alleles(m, n, pri)
        n_new = int(rng.integers(1, 5))
        mp = rng.integers(-1, n_new, len(haps)).astype(np.int32)
        rec[f"m{k}"], rec[f"n{k}"], rec[f"pri{k}"] = m.view(np.uint64), np.array([n]), pri
        rec[f"best{k}"], rec[f"second{k}"], rec[f"bl{k}"], rec[f"sl{k}"], rec[f"conf{k}"] = best, second, bl.view(np.uint64), sl.view(np.uint64), conf.view(np.uint64)
        rec[f"map{k}"], rec[f"marg{k}"] = mp, ref.marginalize(m, n, mp, n_new + 1).view(np.uint64)
    rec["count"] = np.array([n_cases])
    np.savez_compressed(os.path.join(HERE, "alleles.npz"), **rec)
    print("alleles.npz:", n_cases, "cases")


if __name__ == "__main__":
    if "--alleles" in sys.argv:      # --alleles: only the best-allele / marginalisation vectors
        make_alleles()
    else:
        if "--sw" not in sys.argv:   # --sw: only the Smith-Waterman vectors
            main()
        make_sw()
        if "--sw" not in sys.argv:
            make_alleles()
