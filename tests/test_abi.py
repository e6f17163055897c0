"""CPU tests of the drop-in boundary: the library builds for sm_100a, loads, exports every symbol that
include/ipm_b200.h declares, and fails loudly (no fallback) without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ipm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ipm_[a-z_0-9A-Z]+)\s*\(", text)))


def test_header_symbols_are_all_exported(built_library):
    lib = ctypes.CDLL(built_library)
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "library does not export %s" % n


def test_ctypes_table_covers_the_header(built_library):
    from interiorpointmethod_b200 import _lib
    assert sorted(_lib.SYMBOLS) == _declared_symbols()
    _lib.load()


def test_library_contains_sm100a_dmma_code(built_library):
    import subprocess
    out = subprocess.run(["cuobjdump", "-sass", "-arch", "sm_100a", built_library], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "DMMA.8x8x4" in out.stdout


def test_no_cpu_fallback_without_gpu(built_library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from interiorpointmethod_b200 import NewtonStep, _lib
    with pytest.raises(_lib.IpmError, match="IPM_ERR_CUDA"):
        NewtonStep(np.eye(2), [1, 1], [1, 1])


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "interiorpointmethod_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|liboracle|oracle/|oracle_chol|ipm_oracle", text, flags=re.M), f


def test_shard_range_partitions_exactly():
    from interiorpointmethod_b200.batch import shard_range
    for B in (1, 7, 8192, 8191):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                first, count = shard_range(B, r, world)
                seen.extend(range(first, first + count))
            assert seen == list(range(B))
