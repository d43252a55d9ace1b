"""CPU: the C-ABI library builds, loads, exports every symbol include/phmm_b200.h declares, refuses to run without a
GPU (no CPU fallback), carries the reference's tables, and its SASS has the expected instruction mix."""
import ctypes
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

import rovaca_b200
from rovaca_b200 import pairhmm as ph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_header_symbols():
    lib = ph.load_library()
    header = open(os.path.join(ROOT, "include", "phmm_b200.h")).read()
    declared = set(re.findall(r"\b(phmm_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(ph.EXPORTS), (declared ^ set(ph.EXPORTS))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/phmm_b200.h but not exported"
    assert b"sm_100a" in lib.phmm_version()


def test_header_cites_reference_interfaces():
    header = open(os.path.join(ROOT, "include", "phmm_b200.h")).read()
    for cite in ("pairhmm_engine.h:14-15", "pairhmm_engine.cpp:21", "haplotypecaller_engine.cpp:250", "read_record.h:67-103",
                 "pairhmm_internal.cpp:161-182"):
        assert cite in header


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_fails_loudly_without_gpu():
    with pytest.raises(ph.PhmmError) as ei:
        rovaca_b200.PairHmmEngine()
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_multi_device_engine_fails_loudly_without_gpu():
    """phmm_create_multi — one engine over several GPUs — has no CPU fallback either, for any device list."""
    for devices in ("all", [0], [0, 1], [0, 0]):
        with pytest.raises(ph.PhmmError) as ei:
            ph.PairHmmEngine(devices=devices)
        assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_tables_match_oracle(orc):
    p, mm, pd, mmd = ph.host_tables()
    assert np.array_equal(p.view(np.uint32), orc.ph2pr_f32().view(np.uint32))
    assert np.array_equal(mm.view(np.uint32), orc.mm_tri_f32().view(np.uint32))
    assert np.array_equal(pd.view(np.uint64), orc.ph2pr_f64().view(np.uint64))
    assert np.array_equal(mmd.view(np.uint64), orc.mm_tri_f64().view(np.uint64))


def test_restated_log10f_is_libm_or_is_not_used():
    """The host epilogue converts raw sums with the fastest restatement of log10f whose bits equal this host's libm — the
    fdlibm formula around libm's logf (mode 1), or that formula with glibc's table-driven logf restated in AVX-512 (mode 2).
    Whatever phmm_create picks must match libm on every sampled float; a mode that does not match must not be picked.  The
    restated logf only ever sees inputs in [0.5, 2): checked there exhaustively, all 2^24 floats."""
    L = ph.load_library()
    picked = ctypes.c_int(-1)
    assert L.phmm_debug_log10f_check(257, -1, ctypes.byref(picked)) == 0
    assert picked.value in (0, 1, 2)
    for mode in (1, 2):
        bad = L.phmm_debug_log10f_check(263, mode, None)
        if bad == 2**64 - 1:                       # mode 2 without AVX-512 on this host
            assert picked.value < mode
        elif bad:
            assert picked.value < mode, f"mode {mode} differs from libm on {bad} inputs but would be used"
    # mode 3 is the restatement one value at a time, in the device epilogue's operations: the device may run the epilogue only
    # on a host where it gives libm's bits
    assert bool(L.phmm_debug_device_epilogue_allowed()) == (L.phmm_debug_log10f_check(263, 3, None) == 0)
    exhaustive = L.phmm_debug_logf_check(1)
    if exhaustive != 2**64 - 1:
        assert exhaustive == 0 or picked.value < 2


def test_product_never_touches_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "rovaca_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.replace("the oracle", "").lower() or f == "workloads.py", f
    ldd = subprocess.run(["ldd", ph.library_path()], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "phmm_ref" not in ldd


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_sass_audit():
    """Per step the forward kernel must issue exactly 3 FMUL2 + 4 FFMA2 + 2 scalar FMUL (the prior multiply) per row (two
    cells each) and 2 FADD2, i.e. ptxas has not contracted any rounded product into a neighbouring add (it does for
    single-use packed products), and the code is sm_100a packed-FP32 (FFMA2/FMUL2)."""
    sass = subprocess.run(["cuobjdump", "-sass", ph.library_path()], capture_output=True, text=True).stdout
    cur, counts = None, {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = dict(FFMA2=0, FMUL2=0, FADD2=0, SHFL=0, FMUL=0, LDS=0)
            continue
        if cur:
            for op in counts[cur]:
                if re.search(r"\b%s\b" % op, line):
                    counts[cur][op] += 1
    seen = 0
    for name, c in counts.items():
        m = re.search(r"phmm_fwd_kernelILi(\d+)ELi([12])ELi\d+E", name)         # the two contracted variants
        if not m:
            continue
        rpt = int(m.group(1))
        seen += 1
        # copies of the step in the code: the loop is unrolled by two, or by four (+ one step for the remainder) at 10..13 rows
        steps = 5 if 10 <= rpt <= 13 else 2
        assert c["FFMA2"] == 4 * rpt * steps and c["FMUL2"] == 3 * rpt * steps, (name, c)
        assert c["FMUL"] >= 2 * rpt * steps, (name, c)
        assert c["FADD2"] == 2 * steps, (name, c)
        assert c["SHFL"] >= 6 * steps, (name, c)
    # 1..12 rows per lane for any gap qualities, 1..19 for the reference's own, + fixed 8 x 19, + paired reads at 12 rows
    assert seen == 12 + 19 + 1 + 1


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not on PATH")
def test_example_compiles_against_the_header_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/realign_chain.cpp (call_pairhmm + best haplotypes -> Smith-Waterman) must build against
    include/phmm_b200.h and link against the library as it stands; without an sm_100 device it must stop at phmm_create with a
    message, not compute anything."""
    exe = str(tmp_path / "realign_chain")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "realign_chain.cpp"), "-L" + os.path.join(ROOT, "rovaca_b200"), "-lphmm_b200",
                    "-Wl,-rpath," + os.path.join(ROOT, "rovaca_b200"), "-o", exe], check=True)
    if not _has_gpu():
        run = subprocess.run([exe], capture_output=True, text=True)
        assert run.returncode == 1 and "no CPU fallback" in run.stderr and run.stdout == ""


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not on PATH")
def test_header_is_plain_c(tmp_path):
    """include/phmm_b200.h is what a cgo / JNI / ctypes-style binding includes: it must compile as C11, pedantically."""
    src = tmp_path / "hdr.c"
    src.write_text('#include "phmm_b200.h"\nint main(void) { phmm_config c = {0}; phmm_sw_params p = PHMM_SW_PARAMS_REALIGN; '
                   'phmm_result r = {0}; (void)c; (void)p; (void)r; return 0; }\n')
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"), "-fsyntax-only",
                    str(src)], check=True)
