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
