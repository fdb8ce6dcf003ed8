"""CPU-side checks of the ADF boundary (include/mds_adf.h): exported symbols, struct layout, and the
two host-built threshold tables against NumPy — the cosine table against np.histogram's own
binning of float32(arccos(c)), the squared-distance window against the float16 comparison."""
import ctypes
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from mdsuite_b200 import _cuda
from mdsuite_b200._cuda import adf as cuda_adf
from oracle import adf_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = cuda_adf.load()
    with open(os.path.join(ROOT, "include", "mds_adf.h")) as fh:
        declared = sorted(set(re.findall(r"MDS_ADF_API\s+[\w\s\*]+?\b(mds_adf_\w+)\s*\(", fh.read())))
    assert len(declared) >= 9
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mds_adf.h but not exported"
    assert sorted(cuda_adf.EXPORTED_SYMBOLS) == declared
    assert lib.mds_adf_abi_version() == cuda_adf.ABI_VERSION


def test_struct_sizes_match_the_header_layout():
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "sz.c")
        with open(src, "w") as fh:
            fh.write('#include <stdio.h>\n#include "mds_adf.h"\nint main(void){printf("%zu %zu %d %d\\n",'
                     'sizeof(mds_adf_config), sizeof(mds_adf_stats), MDS_ADF_ABI_VERSION,'
                     'MDS_ADF_MAX_NEIGHBOURS);return 0;}\n')
        exe = os.path.join(tmp, "sz")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        cfg, st, ver, max_nb = map(int, subprocess.check_output([exe]).split())
    assert ctypes.sizeof(cuda_adf._Config) == cfg
    assert ctypes.sizeof(cuda_adf._Stats) == st
    assert ver == cuda_adf.ABI_VERSION and max_nb == cuda_adf.MAX_NEIGHBOURS


def _numpy_bins(cosines, nbins):
    theta = np.arccos(cosines.astype(np.float64)).astype(np.float32)
    return np.array([int(np.argmax(np.histogram(theta[k:k + 1], bins=nbins,
                                                range=adf_oracle.BIN_RANGE)[0]))
                     for k in range(len(theta))])


@pytest.mark.parametrize("nbins", [1, 2, 7, 50, 500, 1000])
def test_cosine_table_equals_numpy_binning(nbins):
    edges = cuda_adf.build_threshold_table(nbins)
    assert edges.shape == (nbins + 1,) and np.isinf(edges[0]) and edges[nbins] == -2.0
    inner = edges[1:nbins]
    assert np.all(np.diff(inner) <= 0)  # non-increasing: larger angle, smaller cosine
    reachable = np.nonzero(inner > -2.0)[0] + 1
    # pi = 3.14159 < 3.15: the bins past pi are never hit
    assert abs(len(reachable) - np.pi / 3.15 * nbins) <= 1
    if len(reachable) == 0:
        return
    c_at = edges[reachable]
    c_above = np.nextafter(c_at, np.float32(np.inf), dtype=np.float32)
    assert np.all(_numpy_bins(c_at, nbins) >= reachable)
    ok = c_above <= 1.0
    assert np.all(_numpy_bins(c_above[ok], nbins) < reachable[ok])


def test_cosine_table_bins_random_cosines_like_numpy():
    nbins = 180
    edges = cuda_adf.build_threshold_table(nbins)
    rng = np.random.Generator(np.random.Philox(77))
    c = np.concatenate([rng.uniform(-1, 1, 20000), [-1.0, 1.0, 0.0, -0.0]]).astype(np.float32)
    theta = np.arccos(c.astype(np.float64)).astype(np.float32)
    want = np.histogram(theta, bins=nbins, range=adf_oracle.BIN_RANGE)[0]
    mine = (c[:, None] <= edges[None, 1:nbins]).sum(axis=1)
    assert np.array_equal(np.bincount(mine, minlength=nbins), want)


@pytest.mark.parametrize("cutoff", [3.2, 6.0, 2.5, 0.37, 11.999, 1e-3])
def test_neighbour_window_equals_the_float16_comparison(cutoff):
    s_lo, s_hi = cuda_adf.neighbour_window(cutoff)
    assert 0 < s_lo < s_hi
    for s_edge in (s_lo, s_hi):
        s = np.array([np.nextafter(s_edge, np.float32(0)), s_edge,
                      np.nextafter(s_edge, np.float32(np.inf))], dtype=np.float32)
        for _ in range(3):  # a few more floats on either side
            s = np.concatenate([[np.nextafter(s[0], np.float32(0))], s,
                                [np.nextafter(s[-1], np.float32(np.inf))]]).astype(np.float32)
        want = adf_oracle.neighbour_mask(np.sqrt(s), cutoff)
        mine = (s >= s_lo) & (s < s_hi)
        assert np.array_equal(mine, want), (cutoff, s_edge)
    rng = np.random.Generator(np.random.Philox(5))
    s = (rng.uniform(0, 1.5 * cutoff, 50000).astype(np.float32)) ** 2
    assert np.array_equal((s >= s_lo) & (s < s_hi), adf_oracle.neighbour_mask(np.sqrt(s), cutoff))


def test_bad_arguments_are_rejected_with_a_message():
    with pytest.raises(_cuda.MdsCudaError, match="cutoff"):
        cuda_adf.neighbour_window(-1.0)
    with pytest.raises(_cuda.MdsCudaError, match="cutoff"):
        cuda_adf.neighbour_window(float("nan"))
    with pytest.raises(_cuda.MdsCudaError, match="nbins"):
        cuda_adf.build_threshold_table(0)
    with pytest.raises(_cuda.MdsCudaError, match="n_species"):
        cuda_adf.ADFHandle([1] * 9, 10, 3.0, [5.0, 5.0, 5.0])
    with pytest.raises(_cuda.MdsCudaError, match="norm_power"):
        cuda_adf.ADFHandle([4, 4], 10, 3.0, [5.0, 5.0, 5.0], norm_power=-1)
    with pytest.raises(_cuda.MdsCudaError, match="box"):
        cuda_adf.ADFHandle([4, 4], 10, 3.0, [5.0, 0.0, 5.0])


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is for CPU-only hosts")
    with pytest.raises(_cuda.MdsCudaError, match="no CUDA device|CPU path"):
        cuda_adf.ADFHandle([4, 4], 10, 3.0, [5.0, 5.0, 5.0])
