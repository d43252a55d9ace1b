#!/usr/bin/env python
"""bench.py — PairHMM throughput (GCUPS) of the B200 engine on BASELINE.json's headline workload, plus the other configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--regions R]
                    [--no-workloads] [--no-cpu-baseline] [--sustain SECONDS] [--stream-regions M]

One "step" = one pass of the hot path (quality capping + PCR-indel model, forward recurrence, log10 / normalise / filter /
transpose epilogue, and in the end-to-end leg planning, H2D, D2H and the hand-over to the caller's arrays) over one batch of
synthetic active regions: BASELINE.json configs[1], 150 bp reads x 300 bp haplotypes, 8 haplotypes and 64 reads per region,
R regions per GPU (default 4096 = 94.4 G cells per GPU per step; raw inputs + packed rows + outputs ~ 270 MB > the 126 MB L2).
Multi-GPU (torchrun, one rank per GPU): `value` is weak scaling — independent replicas, every rank its own R regions through
its own engine, no data-path collective; torch.distributed carries the barrier, the max-over-ranks and, for the
`sharded_stream` record, one 64-bit digest per region.

Prints ONE JSON line (rank 0).  `value` is device-timed (CUDA events on the engine's stream) with inputs resident in HBM;
`e2e` is the same metric through the public host-buffer call (phmm_run_flat) with pack/H2D/D2H/epilogue in the timed region.
Beside the headline: `workloads` (N = 1: BASELINE.json configs[2..4] — cfg3 high depth, cfg4 long, cfg4 mix, the cfg5 region
stream at 0 / 50 / 100 % post-trim shapes — each with device GCUPS, end-to-end GCUPS, forward-kernel roofline fraction,
FP64-branch pairs and a parity sample against the reference's own CPU path), `sustained` (the cfg2 device step repeated for
`--sustain` seconds, default 3, with its own clock record) and `sharded_stream` (ONE common cfg5 stream sharded over the N
ranks, per-region digests gathered on rank 0 and compared with a single-GPU run of the whole stream).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "PairHMM GCUPS"
WORKLOAD = "cfg2: 150bp reads x 300bp haplotypes, 8 haps x 64 reads per region (BASELINE.json configs[1])"
FP32_INSTR_PER_CELL = 8            # 4 FMUL + 4 FFMA, SURVEY.md §8d
FLOP_PER_CELL = 12                 # 8 mul + 4 add


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def visible_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        ids = [x for x in vis.split(",") if x != ""]
        if local_rank < len(ids):
            return ids[local_rank]
    return str(local_rank)


def captured_traffic(regions):
    """DRAM bytes (read + write) of one forward-kernel launch from the committed ncu --set full capture, if it was taken at
    this workload size; None otherwise (traffic is per launch, like `achieved`)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "fwd_traffic.json")))
        if int(t["regions"]) == int(regions):
            return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"]), t["source"]
    except Exception:
        pass
    return None, None


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def cpu_reference_run(batch, n_threads, steps=1, warmup=0):
    """The reference's own CPU PairHMM (oracle/_ref: its kernels + call_pairhmm wrapper compiled in place) over `batch` with
    n_threads host threads; falls back to the scalar port (oracle/phmm_oracle.c) where oracle/_ref cannot run.
    -> (kind, [seconds per step], isa)"""
    from oracle import bindings as ob
    ref = ob.ref()
    times = []
    if ref is not None:
        isa = "avx512 (call_new_pairhmm002)" if ref.avx512 else "avx2 (call_old_pairhmm)"
        for it in range(warmup + steps):
            secs, _ = ref.bench_regions(batch, n_threads)
            if it >= warmup:
                times.append(secs)
        return "reference", times, isa
    from concurrent.futures import ThreadPoolExecutor
    orc = ob.oracle()
    regs = [batch.region(g) for g in range(batch.n_regions)]
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(n_threads) as ex:
            list(ex.map(lambda rh: orc.region(rh[0], rh[1]), regs))
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return "port", times, "scalar C restatement"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cfg2_config(regions, cells, pairs, world):
    """The `config` object of the headline workload: the same keys on both arms (the reference arm times a bounded sample of
    it and says so in `sample_of`)."""
    per_region_cells, per_region_pairs = 64 * 150 * 8 * 300, 64 * 8
    return {"workload": WORKLOAD, "regions_per_gpu": regions,
            "cells_per_gpu_per_step": int(cells if cells is not None else regions * per_region_cells),
            "pairs_per_gpu_per_step": int(pairs if pairs is not None else regions * per_region_pairs),
            "pcr_indel_model": "CONSERVATIVE", "base_quality_score_threshold": 18, "fma_mode": "avx512-contracted",
            "l2_policy": "inputs larger than L2: ~270 MB touched per step (raw reads, packed rows, outputs) vs 126 MB L2",
            "parallelism": f"independent replicas: {regions} regions per GPU on each of {world} GPU(s), no collective (weak scaling)"}


def workload_specs(quick):
    """BASELINE.json configs[2..4] at the shapes of SURVEY.md §8d, generated by the C stream generator (workloads.stream)."""
    from rovaca_b200 import workloads as wl
    import numpy as np

    def cfg4_mix(n):
        a, b = wl.stream("cfg4", n // 2, seed=4), wl.stream("cfg2", n // 2, seed=4)
        both = wl.concat([a, b])
        order = np.arange(n).reshape(2, n // 2).T.reshape(-1)          # long, short, long, short, ...
        return wl.take(both, order)
    q = 4 if quick else 1
    return [
        ("cfg3", "high depth: 500 reads x 32 haplotypes per region, R 100-151, H 250-450, 5 % junk reads (FP64 branch)",
         lambda: wl.stream("cfg3", 96 // q, seed=3)),
        ("cfg4", "long: 250 bp reads x 1000 bp haplotypes, 16 haplotypes x 64 reads per region", lambda: wl.stream("cfg4", 192 // q, seed=4)),
        ("cfg4_mix", "cfg4 regions alternating with 150x300 (cfg2) regions: length-binned load balance", lambda: cfg4_mix(384 // q)),
        ("cfg5_trim0", "region stream, Poisson(41) reads of 60-151 bp x ~8.8 haplotypes of 201-499 bp", lambda: wl.stream("cfg5", 16384 // q, seed=5)),
        ("cfg5_trim50", "region stream, 50 % of the regions with post-trim shapes (H 41-120, R <= H)",
         lambda: wl.stream("cfg5", 16384 // q, seed=5, post_trim_fraction=0.5)),
        ("cfg5_trim100", "region stream, post-trim shapes only", lambda: wl.stream("cfg5", 65536 // q, seed=5, post_trim_fraction=1.0)),
    ]


def run_workloads(eng, prop_sms, sm_max_mhz, quick, parity):
    """Every other BASELINE config through the same two legs as the headline (device-resident, end to end)."""
    import numpy as np
    from rovaca_b200.pairhmm import FlatResult
    rows = []
    for key, what, make in workload_specs(quick):
        b = make().pin()
        cells = b.cells()
        res = eng.upload(b)
        for _ in range(2):
            res.run()
        t = [res.run() for _ in range(3)]
        dev, fwd = sum(x[0] for x in t) / 3, sum(x[1] for x in t) / 3
        launches = res.launches
        res.free()
        out = FlatResult(b)
        for _ in range(2):                                   # the first calls size the pinned / device arenas of every worker
            eng.run_flat(b, out=out)
        s0 = eng.stats()
        t0 = time.perf_counter()
        for _ in range(3):
            eng.run_flat(b, out=out)
        e2e = (time.perf_counter() - t0) / 3
        s1 = eng.stats()
        row = {"workload": key, "what": what, "regions": b.n_regions, "gcells": round(cells / 1e9, 2), "pairs": b.pairs(),
               "device_ms": round(dev, 3), "fwd_ms": round(fwd, 3), "device_gcups": round(cells / dev / 1e6, 1),
               "e2e_ms": round(e2e * 1e3, 3), "e2e_gcups": round(cells / e2e / 1e9, 1), "e2e_over_device": round(dev * 1e-3 / e2e, 3),
               "fwd_frac": round(cells * FP32_INSTR_PER_CELL / (fwd * 1e-3) / (prop_sms * 128 * sm_max_mhz * 1e6), 4),
               "fp64_pairs": int((s1["fp64_pairs"] - s0["fp64_pairs"]) // 3), "slow_pairs": int((s1["slow_pairs"] - s0["slow_pairs"]) // 3),
               "launches_per_step": int(launches), "batches_per_e2e_step": int((s1["batches"] - s0["batches"]) // 3)}
        if parity:
            from oracle import check
            chk = check.sample_parity(b, out, n_sample=48 if key.startswith("cfg5") else 4)
            row["parity_sample_ok"] = chk["ok"]
            row["parity_sample"] = f"{chk['regions_checked']} regions byte-compared with the {chk['against']} CPU path, {chk['regions_differing']} differ"
        rows.append(row)
    return rows


def sustained_leg(resident, seconds, gpu_index, cells):
    """The device step back to back for `seconds` of wall time, with its own clock record: evidence for a seconds-long run
    (the K timed steps of the headline last about half a second)."""
    if seconds <= 0:
        return None
    sampler = ClockSampler(gpu_index)
    t0 = time.perf_counter()
    ms = []
    while time.perf_counter() - t0 < seconds:
        ms.append(resident.run()[0])
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    half = ms[len(ms) // 2:]
    return {"seconds": round(wall, 2), "steps": len(ms), "gcups": round(cells * len(ms) / (sum(ms) * 1e-3) / 1e9, 1),
            "gcups_second_half": round(cells * len(half) / (sum(half) * 1e-3) / 1e9, 1),
            "ms_per_step_min_max": [round(min(ms), 3), round(max(ms), 3)], "device_busy_fraction_of_wall": round(sum(ms) * 1e-3 / wall, 3), "clocks": clocks}


def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores (rank 0 only)."""
    if rank != 0:
        return
    from rovaca_b200 import workloads
    cores = host_threads()
    n_regions = min(args.regions, max(64, 96 * cores))
    batch = workloads.cfg2(n_regions, seed=2)
    kind, times, isa = cpu_reference_run(batch, cores, args.steps, args.warmup)
    t = max(statistics.mean(times), 1e-9)
    gcups = batch.cells() / t / 1e9
    sample = f"{n_regions} of {args.regions} cfg2 regions per step ({batch.cells() / 1e9:.1f} G cells), {isa}, whole regions per thread"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gcups, 3), "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(t * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": cfg2_config(args.regions, None, None, args.gpus) | {"sample_of": f"{n_regions} of the {args.regions} regions per step"},
        "cpu_baseline": {"value": round(gcups, 3), "unit": "GCUPS", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(gcups, 3), "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def ours_arm(args, rank, world, local_rank):
    import numpy as np
    import torch
    from rovaca_b200 import PairHmmEngine, workloads, sharding
    from rovaca_b200.pairhmm import FlatResult

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    K, W = args.steps, max(args.warmup, 0)
    batch = workloads.cfg2(args.regions, seed=2 + rank)       # every rank its own regions: weak scaling
    batch.pin()                                               # the step's inputs wait in pinned host memory (bench contract)
    cells = batch.cells()
    eng = PairHmmEngine(device=local_rank)
    sampler = ClockSampler(visible_gpu_index(local_rank)) if rank == 0 else None

    # ---- leg 1: device path only, inputs resident in HBM -----------------------------------------
    resident = eng.upload(batch)
    for _ in range(W):
        resident.run()
    st0 = eng.stats()
    sharding.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    all_ms, fwd_ms = [], []
    for _ in range(K):
        a, f = resident.run()
        all_ms.append(a)
        fwd_ms.append(f)
    torch.cuda.synchronize()
    sharding.barrier()
    wall_dev = (time.perf_counter() - t0) / K
    st1 = eng.stats()
    launches = int(st1["kernel_launches"] - st0["kernel_launches"])
    ms_step = sharding.max_over_ranks(sum(all_ms) / K)
    fwd_step = sharding.max_over_ranks(sum(fwd_ms) / K)
    total_cells = sharding.sum_over_ranks(float(cells))
    value = total_cells / (ms_step * 1e-3) / 1e9
    check = resident.download()
    fp64_step = int(check.n_fp64)
    sustained = sustained_leg(resident, args.sustain, visible_gpu_index(local_rank), cells) if rank == 0 else None
    sharding.barrier()
    resident.free()

    # ---- leg 2: end to end through the host-buffer call ------------------------------------------
    out = FlatResult(batch)
    for _ in range(max(1, min(W, 2))):
        eng.run_flat(batch, out=out)
    st2 = eng.stats()
    sharding.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        eng.run_flat(batch, out=out)
    torch.cuda.synchronize()
    e2e_s = sharding.max_over_ranks((time.perf_counter() - t0) / K)
    st3 = eng.stats()
    e2e = total_cells / e2e_s / 1e9
    same = bool(np.array_equal(out.likelihoods.view(np.uint64), check.likelihoods.view(np.uint64)))
    clocks = sampler.stop() if sampler else None

    # ---- ONE common stream sharded over the ranks, digests compared with a single-GPU run (BASELINE.json configs[4]) ----
    shard_rec = None
    if args.stream_regions > 0:
        from rovaca_b200 import streams
        shard_rec = streams.sharded_stream(eng, streams.StreamSpec("cfg5", args.stream_regions, 5, 0.5), rank, world)

    if rank != 0:
        eng.close()
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (forward wavefront) -------------------------------------
    prop = torch.cuda.get_device_properties(local_rank)
    peaks = measured_peaks()
    sm_max = float(peaks.get("sm_max_mhz", (clocks or {}).get("sm_max_mhz", 1965.0)))
    n_sm = prop.multi_processor_count
    peak_tflops = 2 * 128 * n_sm * sm_max * 1e6 / 1e12                       # FP32 FMA peak at the max SM clock
    achieved = cells * FP32_INSTR_PER_CELL * 2 / (fwd_step * 1e-3) / 1e12    # FMA-slot equivalents (8 FP32 instr/cell)
    _, ffma = eng.ubench(0)
    _, ffma2 = eng.ubench(1)
    measured_peak = max(ffma, ffma2) * 2 / 1e12
    ki = eng.kernel_info(150)
    regs, spill = ki["regs"], ki["spill_bytes"]
    traffic, traffic_src = captured_traffic(args.regions)
    roof = {"bound": "fp32", "achieved": round(achieved, 3), "peak": round(peak_tflops, 3), "unit": "TFLOP/s", "frac": round(achieved / peak_tflops, 4),
            "traffic": traffic, "traffic_unit": "bytes of DRAM read+write per forward launch", "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": int(4 * int(batch.read_len.sum()) + int(batch.hap_len.sum()) + 32 * batch.n_reads + 24 * batch.n_haps + 4 * batch.pairs()),
            "algorithmic_bytes_model": "4 B packed word per read base + 1 B per haplotype base + 32 B per read record + 24 B per haplotype record in, 4 B per pair out",
            "kernel": f"phmm_fwd_kernel<RPT={ki['rows_per_lane']}, LANES={ki['lanes_per_read']}> (forward wavefront, cfg2 reads)",
            "peak_source": f"2 x 128 lanes x {n_sm} SMs x {sm_max:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json); no FP32 figure in that file",
            "work_model": "8 FP32-pipe instructions per cell (4 FMUL + 4 FFMA) counted as FMA slots = 16 flop-equivalents; plain 12 flop/cell also given",
            "fwd_kernel_ms": round(fwd_step, 3), "tcups": round(cells / (fwd_step * 1e-3) / 1e12, 4),
            "plain_flop_tflops": round(cells * FLOP_PER_CELL / (fwd_step * 1e-3) / 1e12, 3),
            "measured_ffma_peak_tflops": round(measured_peak, 3), "frac_of_measured_peak": round(achieved / measured_peak, 4),
            "ubench_tflops": {"ffma": round(ffma * 2 / 1e12, 2), "ffma2": round(ffma2 * 2 / 1e12, 2)},
            "regs_per_thread": regs, "spill_bytes": spill}
    # the contract's usual bound for a data-parallel path is HBM; shown here to document that this path is nowhere near it
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_bytes = traffic if traffic else roof["algorithmic_bytes_per_launch"]
    roof["hbm_view"] = {"achieved_gbs": round(hbm_bytes / (fwd_step * 1e-3) / 1e9, 2), "peak_gbs": hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.65 TB/s",
                        "frac": round(hbm_bytes / (fwd_step * 1e-3) / 1e9 / hbm_peak, 5),
                        "note": "0.002 B per cell: the recurrence is bound by the FP32 issue / register-file budget, not by memory"}
    if clocks:
        roof["frac_at_observed_clock"] = round(achieved / (2 * 128 * n_sm * clocks["sm_mhz"] * 1e6 / 1e12), 4)

    # ---- CPU baseline beside it (N = 1 only) -----------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = host_threads()
        n_cpu = min(args.regions, max(64, 128 * cores))
        sample_batch = batch.subset(range(n_cpu)) if n_cpu < batch.n_regions else batch
        kind, times, isa = cpu_reference_run(sample_batch, cores, 1, 0)
        cpu = {"value": round(sample_batch.cells() / times[0] / 1e9, 3), "unit": "GCUPS", "cores": cores, "kind": kind,
               "sample": f"first {n_cpu} of {args.regions} regions ({sample_batch.cells() / 1e9:.1f} G cells), {isa}, {times[0]:.2f} s wall"}
    # ---- the other BASELINE configs (N = 1 only; the scaling runs time the headline and the sharded stream) ----
    work_rows = None
    if world == 1 and not args.no_workloads:
        work_rows = run_workloads(eng, n_sm, sm_max, args.quick, parity=not args.no_cpu_baseline)

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": "GCUPS", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": cfg2_config(args.regions, cells, batch.pairs(), world),
        "e2e": {"value": round(e2e, 2), "unit": "GCUPS", "h2d_bytes_per_step": int((st3["h2d_bytes"] - st2["h2d_bytes"]) / K),
                "d2h_bytes_per_step": int((st3["d2h_bytes"] - st2["d2h_bytes"]) / K), "ms_per_step": round(e2e_s * 1e3, 3),
                "api": "phmm_run_flat (pinned host buffers in, host likelihood matrices out)", "matches_resident_run": same,
                "host_plan_ms_per_step": round((st3["host_plan_ms"] - st2["host_plan_ms"]) / K, 2),
                "host_epilogue_ms_per_step": round((st3["host_epilogue_ms"] - st2["host_epilogue_ms"]) / K, 2),
                "gpu_ms_per_step": round((st3["gpu_ms"] - st2["gpu_ms"]) / K, 2), "batches_per_step": int((st3["batches"] - st2["batches"]) / K)},
        "gpu_launches": launches,
        "roofline": roof,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "device_step_wall_ms": round(wall_dev * 1e3, 3),
        "fp64_pairs_per_step": fp64_step,
        "gpu": prop.name,
        "sustained": sustained,
        "workloads": work_rows,
        "sharded_stream": shard_rec,
    }
    print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--regions", type=int, default=4096, help="cfg2 regions per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU reference legs (baseline and parity samples)")
    ap.add_argument("--no-workloads", action="store_true", help="skip the cfg3 / cfg4 / cfg5 table")
    ap.add_argument("--quick", action="store_true", help="quarter-size workloads table")
    ap.add_argument("--sustain", type=float, default=3.0, help="seconds of back-to-back device steps for the sustained record (0 = off)")
    ap.add_argument("--stream-regions", type=int, default=131072, help="regions of the common cfg5 stream sharded over the ranks (0 = off)")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        print(f"bench.py: --gpus {args.gpus} needs torchrun (one rank per GPU); running 1 GPU", file=sys.stderr)
    ours_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
