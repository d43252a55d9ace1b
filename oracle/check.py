"""Checker helpers (test infrastructure, like everything under oracle/): used by tests/ and by bench.py's CPU-baseline leg only.

``sample_parity`` compares a sample of the regions of an engine result with the reference's own call_pairhmm (oracle/_ref,
compiled in place from /root/reference) or, where that cannot load, with the scalar port."""
from typing import Optional

import numpy as np

from rovaca_b200 import workloads


def sample_parity(batch, result, n_sample: int = 96, threads: Optional[int] = None, seed: int = 0) -> dict:
    """``n_sample`` regions of ``batch``, evenly spread, through the reference's
    own call_pairhmm (oracle/_ref, compiled in place) — or the scalar port where that cannot load — compared byte for byte
    with ``result`` (a FlatResult of the whole batch)."""
    from oracle import bindings as ob
    n = batch.n_regions
    pick = np.unique(np.linspace(0, n - 1, min(n_sample, n)).astype(np.int64))
    sub = workloads.take(batch, pick)
    th = workloads._threads(threads)
    ref = ob.ref()
    lo, slo = result.lik_off, sub.lik_off()
    if ref is not None:
        lik, kept, nk = ref.regions_flat(sub, th)
        kind = "reference"
    else:
        from concurrent.futures import ThreadPoolExecutor
        orc = ob.oracle()
        lik, kept, nk = np.empty(max(int(slo[-1]), 1)), np.zeros(max(sub.n_reads, 1), np.uint32), np.zeros(len(pick), np.uint32)

        def one(i):
            reads, haps = sub.region(i)
            m, k, *_ = orc.region(reads, haps)
            lik[int(slo[i]):int(slo[i + 1])] = m.reshape(-1)
            r0 = int(sub.reg_read_off[i])
            kept[r0:r0 + len(k)] = k
            nk[i] = len(k)
        with ThreadPoolExecutor(th) as ex:
            list(ex.map(one, range(len(pick))))
        kind = "port"
    bad = 0
    for i, g in enumerate(pick):
        a = result.likelihoods[int(lo[g]):int(lo[g + 1])].view(np.uint64)
        b = lik[int(slo[i]):int(slo[i + 1])].view(np.uint64)
        r0, s0 = int(batch.reg_read_off[g]), int(sub.reg_read_off[i])
        same = np.array_equal(a, b) and int(result.n_kept[g]) == int(nk[i]) and \
            np.array_equal(result.kept_idx[r0:r0 + int(nk[i])], kept[s0:s0 + int(nk[i])])
        bad += not same
    return {"regions_checked": int(len(pick)), "regions_differing": int(bad), "against": kind, "ok": bad == 0}
