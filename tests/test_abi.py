"""CPU tests of the drop-in boundary: libnpe_cuda.so loads, exports every symbol include/npe_cuda.h declares, and
fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "npe_cuda.h")


@pytest.fixture(scope="module")
def lib():
    from dynamics_b200.build import build
    from dynamics_b200.lib import load_library
    return load_library(build())


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(npe_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from dynamics_b200.lib import SIGNATURES
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib.dll, s), f"{s} declared in npe_cuda.h but not exported"
        assert s in SIGNATURES, f"{s} has no ctypes signature"
    assert set(SIGNATURES) == set(syms)
    assert lib.npe_abi_version() == 1


def test_config_struct_matches_header(lib):
    from dynamics_b200.lib import NPEConfig
    cfg = lib.default_config()
    assert cfg.struct_size == ctypes.sizeof(NPEConfig) == 17 * 4
    assert (cfg.rnn_dim, cfg.layers, cfg.num_past, cfg.num_future, cfg.object_dim) == (50, 5, 2, 1, 22)
    assert cfg.nbrhd == 1 and abs(cfg.nbrhdsize - 3.5) < 1e-7 and cfg.max_ctx == 12
    assert cfg.vlambda == 1.0 and cfg.lambda_ == 1.0
    assert lib.npe_param_count(None) == 28222


def test_neighbourhood_threshold_on_squared_distances(lib):
    """The pair builder tests s <= s* instead of fl(800 * fl(sqrt(s))) <= thr (npe_nbr_threshold_sq): the two must agree for
    EVERY float; checked on the 2^16 floats either side of s*, on random magnitudes and at the extremes (host arithmetic)."""
    import numpy as np
    def reference(s, thr):                                 # one rounding per operation, float32 throughout
        return np.sqrt(s.astype(np.float32)) * np.float32(800.0) <= np.float32(thr)
    rng = np.random.default_rng(0)
    for thr in (210.0, 3.5 * 60, 60.0, 1.0, 799.99, 1e-3, 5e4):
        f = getattr(lib.dll, "npe_nbr_threshold_sq")
        sstar = np.float32(f(ctypes.c_float(thr)))
        assert sstar > 0
        bits = np.array([sstar], np.float32).view(np.uint32)[0]
        near = (np.arange(-65536, 65536, dtype=np.int64) + int(bits)).astype(np.uint32).view(np.float32)
        wide = np.abs(rng.standard_normal(200000).astype(np.float32)) * np.float32(10.0) ** rng.integers(-12, 6, 200000).astype(np.float32)
        edge = np.array([0.0, np.finfo(np.float32).tiny, 1e-45, np.finfo(np.float32).max, np.inf], np.float32)
        for s in (near, wide, edge):
            assert np.array_equal(reference(s, thr), s <= sstar)
    assert lib.dll.npe_nbr_threshold_sq(ctypes.c_float(-1.0)) < 0      # nothing is within a negative radius


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_fails_loudly_without_gpu(lib):
    from dynamics_b200 import NPEError, NPEModel
    with pytest.raises(NPEError) as ei:
        NPEModel()
    assert "no CPU path" in str(ei.value)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under dynamics_b200/ may import or execute it."""
    pkg = os.path.join(ROOT, "dynamics_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".lua", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def _run_c_client(tmp_path):
    import shutil
    import subprocess
    from dynamics_b200.build import build
    lib = build(force=False)
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    exe = tmp_path / "c_abi_check"
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c_abi_check.c")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-O1", src, "-o", str(exe), lib, "-lm",
                    "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    return r.stdout.strip()


def test_plain_c_client(tmp_path):
    """The header is C, the library links from C and reports errors as codes + messages (tests/c_abi_check.c).  On the
    CPU box the binary must see the loud no-GPU failure; the B200 run of the same binary is test_plain_c_client_gpu."""
    assert _run_c_client(tmp_path) in ("gpu", "nogpu")


@pytest.mark.gpu
def test_plain_c_client_gpu(tmp_path):
    """The same plain-C binary on the B200: a tiny batch through upload / forward / backward / Adam from C."""
    assert _run_c_client(tmp_path) == "gpu"


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under dynamics_b200/ (Python, CUDA, Lua) may import, link or execute it, and
    bench.py / __graft_entry__.py only in their checker legs (cpu_baseline / --impl reference / smoke)."""
    pkg = os.path.join(ROOT, "dynamics_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".lua")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "oracle/" not in text and "oracle." not in text.replace("the oracle.", ""), f
    bench = open(os.path.join(ROOT, "bench.py")).read()
    uses = [m.start() for m in re.finditer(r"from oracle|import oracle", bench)]
    assert uses, "bench.py times the CPU port in its cpu_baseline / reference legs"
    for pos in uses:                                            # every use sits inside one of the two checker functions
        head = bench[:pos]
        fn = re.findall(r"^def (\w+)\(", head, re.M)[-1]
        assert fn in ("cpu_baseline_leg", "cpu_rollout_leg", "run_reference"), fn
