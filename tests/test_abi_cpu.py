"""CPU tests of the C ABI: the library loads, exports every symbol include/mds_rdf.h declares, its
host-only pieces agree with the oracle, and compute entry points fail loudly without a GPU."""
import os
import re

import numpy as np
import pytest

from mdsuite_b200 import _cuda
from oracle import rdf_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "mds_rdf.h")).read()
    return sorted(set(re.findall(r"MDS_API\s+[\w\s\*]+?\b(mds_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _cuda.load()
    declared = header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mds_rdf.h but not exported"
    assert sorted(_cuda.EXPORTED_SYMBOLS) == declared
    assert lib.mds_abi_version() == _cuda.ABI_VERSION


def test_struct_sizes_match_the_header_layout():
    """sizeof() of the header's structs, as gcc lays them out, equals the ctypes mirrors."""
    import ctypes
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "sz.c")
        with open(src, "w") as fh:
            fh.write('#include <stdio.h>\n#include "mds_rdf.h"\nint main(void){printf("%zu %zu %d\\n",'
                     'sizeof(mds_rdf_config), sizeof(mds_rdf_stats), MDS_RDF_ABI_VERSION);return 0;}\n')
        exe = os.path.join(tmp, "sz")
        subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), src, "-o", exe])
        cfg, st, ver = map(int, subprocess.check_output([exe]).split())
    assert ctypes.sizeof(_cuda._Config) == cfg == 48
    assert ctypes.sizeof(_cuda._Stats) == st == 136
    assert ver == _cuda.ABI_VERSION


@pytest.mark.parametrize("cutoff,nbins", [(10.0, 500), (37.699999999999996, 3769), (3.0, 300),
                                          (8.0, 800), (49.9, 5000), (25.0, 100), (1.0, 1)])
def test_library_threshold_table_equals_oracle_derivation(cutoff, nbins):
    edges, s_cut = _cuda.build_threshold_table(cutoff, nbins)
    want_edges, want_cut = rdf_oracle.s_threshold_table(cutoff, nbins)
    assert s_cut == want_cut and np.array_equal(edges, want_edges)
    assert edges[0] == 0 and edges[nbins] == s_cut and np.all(np.diff(edges) >= 0)


def test_bad_arguments_are_rejected_with_a_message():
    with pytest.raises(_cuda.MdsCudaError, match="bad cutoff"):
        _cuda.build_threshold_table(-1.0, 10)
    with pytest.raises(_cuda.MdsCudaError):
        _cuda.build_threshold_table(1.0, 0)
    with pytest.raises(_cuda.MdsCudaError, match="nbins"):
        _cuda.RDFHandle([4, 4], 10 ** 7, 1.0, [5.0, 5.0, 5.0])
    with pytest.raises(_cuda.MdsCudaError, match="box"):
        _cuda.RDFHandle([4, 4], 10, 1.0, [5.0, 0.0, 5.0])
    with pytest.raises(_cuda.MdsCudaError, match="cutoff"):
        _cuda.RDFHandle([4, 4], 10, float("nan"), [5.0, 5.0, 5.0])


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is for CPU-only hosts")
    with pytest.raises(_cuda.MdsCudaError, match="no CUDA device|CPU path"):
        _cuda.RDFHandle([4, 4], 10, 1.0, [5.0, 5.0, 5.0])
    with pytest.raises(_cuda.MdsCudaError):
        _cuda.bench_shared_atomics(0)


def test_product_never_imports_the_oracle():
    for dirpath, _dirs, files in os.walk(os.path.join(ROOT, "mdsuite_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in text and "import oracle" not in text, f
                assert "librdf_oracle" not in text, f


@pytest.mark.parametrize("cell_path", [False, True])
def test_host_submit_batch_plan(cell_path):
    """mds_rdf_plan_host_batches (include/mds_rdf.h): how a host submit is cut into copy/compute
    batches.  Every frame exactly once, no batch above the handle's size; a ramp only when the
    handle is idle, growing no faster than x 1.5 (cell path) / x 4; equal batches otherwise, so that
    no short batch's kernels have to hide a full-size copy."""
    for bf in (1, 3, 16, 164, 296):
        for n in (0, 1, 7, 8, 100, bf, bf + 1, 1200, 10_000):
            busy = _cuda.plan_host_batches(n, bf, False, cell_path)
            assert sum(busy) == n and all(1 <= b <= bf for b in busy)
            assert len(busy) == -(-n // bf)                      # as few batches as possible ...
            assert all(b == busy[0] for b in busy[:-1])            # ... all equal but the last,
            assert not busy or busy[0] - len(busy) < busy[-1] <= busy[0]  # short by rounding only
            idle = _cuda.plan_host_batches(n, bf, True, cell_path)
            assert sum(idle) == n and all(1 <= b <= bf for b in idle)
            if n < 8 or bf < 2:
                assert idle == busy
                continue
            assert idle[0] == min(n, max(1, bf // 16))           # the exposed first copy is small
            growth = 1.5 if cell_path else 4.0
            full = [k for k, b in enumerate(idle) if b * growth >= bf] or [len(idle)]
            for a, b in zip(idle[:full[0]], idle[1:full[0] + 1]):
                assert b <= np.ceil(a * growth) + 1 or b <= -(-n // max(1, len(idle)))
    # the schedule quoted in DESIGN.md §7 for cfg3 (296-frame batches, 1,250 frames of a rank)
    assert _cuda.plan_host_batches(1250, 296, True, True) == [18, 27, 41, 62, 93, 140, 210, 220, 220, 219]
    assert _cuda.plan_host_batches(1200, 296, False, True) == [240] * 5
    assert _cuda.plan_host_batches(1000, 48, True, False)[:3] == [3, 12, 47]
    with pytest.raises(ValueError):
        _cuda.plan_host_batches(10, 0, True, True)
