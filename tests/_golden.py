"""Loaders for tests/golden/*.npz (made by tests/golden/make_golden.py from the compiled reference)."""
import os

import numpy as np

from rovaca_b200.workloads import RegionBatch

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pairs():
    """-> list of dict(hap, rs, q, ins, del, gcp, ref_avx512, ref_avx2, ref_1xn) (raw sums as uint32 bit patterns)."""
    z = np.load(os.path.join(HERE, "pairs.npz"))
    out, ro, ho = [], 0, 0
    for k, (R, H) in enumerate(z["lens"]):
        R, H = int(R), int(H)
        d = dict(hap=z["hap"][ho:ho + H], rs=z["rs"][ro:ro + R], q=z["q"][ro:ro + R], ins=z["ins"][ro:ro + R],
                 dele=z["del"][ro:ro + R], gcp=z["gcp"][ro:ro + R], ref_avx512=z["ref_avx512"][k], ref_avx2=z["ref_avx2"][k],
                 ref_1xn=z["ref_1xn"][k])
        out.append(d)
        ro += R
        ho += H
    return out


def regions():
    """-> (RegionBatch, meta[n,2] (threshold, pcr model), list of matrix bit patterns, list of kept idx)."""
    z = np.load(os.path.join(HERE, "regions.npz"))
    b = RegionBatch(reg_read_off=z["reg_read_off"], reg_hap_off=z["reg_hap_off"], read_len=z["read_len"], read_mapq=z["read_mapq"],
                    read_off=z["read_off"], read_bases=z["read_bases"], read_qual=z["read_qual"], hap_len=z["hap_len"],
                    hap_off=z["hap_off"], hap_bases=z["hap_bases"], name="golden regions")
    lik = b.lik_off()
    mats, kept, ko = [], [], 0
    for g in range(b.n_regions):
        mats.append(z["mats"][int(lik[g]):int(lik[g + 1])])
        n = int(z["n_kept"][g])
        kept.append(z["kept"][ko:ko + n])
        ko += n
    return b, z["meta"], mats, kept


def pcr():
    z = np.load(os.path.join(HERE, "pcr.npz"))
    out, so, co = [], 0, 0
    for n in z["seq_len"]:
        n = int(n)
        seq = z["seq"][so:so + n]
        capped = [z["capped"][co + m * n: co + (m + 1) * n] for m in range(3)]
        out.append(dict(seq=seq, capped=capped, repeats=z["repeats"][so:so + n]))
        so += n
        co += 3 * n
    return out


def sw():
    """-> list of dict(hap, read, realign=(cigar, offset, score), assemble=(cigar, offset, score)) from the reference's sw_avx.c."""
    z = np.load(os.path.join(HERE, "sw.npz"))
    out = []
    for k in range(int(z["n"][0])):
        d = dict(hap=z[f"hap{k}"], read=z[f"read{k}"])
        for tag, name in (("r", "realign"), ("a", "assemble")):
            off, score = (int(x) for x in z[f"meta_{tag}{k}"])
            d[name] = (z[f"cig_{tag}{k}"], off, score)
        out.append(d)
    return out


def alleles():
    """-> list of dict(matrix, n_evidence, priorities, best, second, best_lik, second_lik, confidence, old_to_new, marginal) from
    the reference's AlleleLikelihoods (make_golden.py --alleles)."""
    z = np.load(os.path.join(HERE, "alleles.npz"))
    out = []
    for k in range(int(z["count"][0])):
        out.append(dict(matrix=z[f"m{k}"].view(np.float64), n_evidence=int(z[f"n{k}"][0]), priorities=z[f"pri{k}"], best=z[f"best{k}"],
                        second=z[f"second{k}"], best_lik=z[f"bl{k}"], second_lik=z[f"sl{k}"], confidence=z[f"conf{k}"],
                        old_to_new=z[f"map{k}"], marginal=z[f"marg{k}"]))
    return out
