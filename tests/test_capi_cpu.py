"""CPU-side checks of the product library: it builds for sm_100a, loads, exports every symbol
include/shodh_gpu.h declares, and FAILS LOUDLY without a GPU (no CPU fallback)."""
import os
import subprocess

import pytest


def test_library_exports_every_declared_symbol(sg):
    syms = sg.capi.declared_symbols()
    assert len(syms) >= 20
    L = sg.capi.lib()
    for s in syms:
        assert hasattr(L, s), s
    assert "sm_100a" in sg.version()


def test_sass_is_sm100_and_uses_tma(sg):
    out = subprocess.run(["cuobjdump", "-lelf", sg.capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", sg.capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTMALDG" in sass                # cp.async.bulk.tensor -> TMA load
    assert "SYNCS" in sass                  # mbarrier


def test_no_cpu_fallback_without_gpu(sg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sg.SgError) as e:
        sg.GpuVectorStore(384, "f16")
    assert e.value.code == sg.capi.SG_ERR_NO_DEVICE
    import numpy as np
    with pytest.raises(sg.SgError) as e:                       # every other entry point that owns device state, too
        sg.GpuProductQuantizer.from_centroids(np.zeros((48, 256, 8), np.float32))
    assert e.value.code == sg.capi.SG_ERR_NO_DEVICE
    with pytest.raises(sg.SgError) as e:
        sg.GpuSpannIndex.from_arrays(np.zeros((4, 384), np.float32), np.zeros((48, 256, 8), np.float32),
                                     np.zeros(5, np.uint64), np.zeros(0, np.uint32), np.zeros((0, 48), np.uint8))
    assert e.value.code == sg.capi.SG_ERR_NO_DEVICE
    with pytest.raises(sg.SgError) as e:
        sg.top_k_similar(np.ones(32, np.float32), np.ones((3, 32), np.float32), 2)
    assert e.value.code == sg.capi.SG_ERR_NO_DEVICE
    assert "UBLKCP" in subprocess.run(["cuobjdump", "-sass", sg.capi.LIB_PATH], capture_output=True, text=True).stdout


def test_product_package_does_not_import_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "shodh_memory_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_rust_binding_source_matches_the_header(sg):
    """rust/gpu_ffi.rs is not compiled here (no Rust toolchain): at least every `pub fn sg_*` it declares must be an
    entry point include/shodh_gpu.h declares (and the library exports), with the same number of parameters."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rs = open(os.path.join(root, "rust", "gpu_ffi.rs")).read()
    hdr = open(os.path.join(root, "include", "shodh_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    declared = set(sg.capi.declared_symbols())
    c_arity = {}
    for m in re.finditer(r"SG_API\s+[\w\s\*]+?\b(sg_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        c_arity[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    fns = re.findall(r"pub fn (sg_\w+)\s*\(([^;]*?)\)\s*(?:->\s*[\w\*\s:]+)?;", rs, flags=re.S)
    assert len(fns) >= 60
    for name, args in fns:
        assert name in declared, f"{name} is declared in rust/gpu_ffi.rs but not in include/shodh_gpu.h"
        n = 0 if not args.strip() else args.count(",") + 1
        assert n == c_arity[name], f"{name}: {n} parameters in rust/gpu_ffi.rs, {c_arity[name]} in the header"
    for f in ("gpu_store.rs", "gpu_embedder.rs", "gpu_pq.rs"):       # every call in the wrappers resolves to a declaration
        for name in set(re.findall(r"\b(sg_\w+)\s*\(", open(os.path.join(root, "rust", f)).read())):
            assert name in dict(fns) or name == "sg_last_error", (f, name)
