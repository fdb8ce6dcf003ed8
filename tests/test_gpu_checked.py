"""The self-checking build of the CUDA library (-DMDS_CHECKED, libmds_rdf_checked.so): every CTA
verifies at exit that the histogram atomics it issued all landed inside its own histogram region
(rdf_common.cuh).  compute-sanitizer is closed on the GPU pool this was developed on; this is the
replacement for its memcheck on the one kind of access the parity tests cannot see — a stray
shared-memory atomic.  Runs a cross-section of the parity suites in a subprocess with
MDS_RDF_LIB pointing at the checked library."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKED = os.path.join(ROOT, "mdsuite_b200", "_cuda", "libmds_rdf_checked.so")

SELECTION = [
    "tests/test_gpu_parity.py::test_tiny_two_species_vs_literal_numpy_oracle",
    "tests/test_gpu_parity.py::test_three_species_noncubic_cutoff_beyond_half_box",
    "tests/test_gpu_parity.py::test_unwrapped_coordinates_use_general_minimum_image",
    "tests/test_gpu_parity.py::test_nan_and_empty_inputs",
    "tests/test_gpu_parity.py::test_every_kernel_variant",
    "tests/test_gpu_parity.py::test_cfg2_shape_nacl_13824_default_cutoff_and_bins",
    "tests/test_gpu_parity.py::test_many_species_multi_pass_and_large_bins",
    "tests/test_gpu_cells.py::test_lj_4096_vs_oracle",
    "tests/test_gpu_cells.py::test_water_three_pairs_vs_oracle",
    "tests/test_gpu_cells.py::test_noncubic_three_species_minimum_grid",
    "tests/test_gpu_cells.py::test_unwrapped_centered_and_far_frames",
    "tests/test_gpu_cells.py::test_nan_inf_coordinates",
    "tests/test_gpu_cells.py::test_clustered_atoms_row_passes_and_histogram_flush",
    "tests/test_gpu_cells.py::test_many_species_pair_groups_and_global_histogram",
    "tests/test_gpu_cells.py::test_cfg3_lj_100k_one_frame_vs_oracle_and_allpairs",
]


def test_parity_suites_pass_on_the_checked_library():
    if not os.path.exists(CHECKED):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "mdsuite_b200", "_cuda", "csrc"),
                               "checked"])
    env = dict(os.environ, MDS_RDF_LIB=CHECKED)
    proc = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", *SELECTION],
                          cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (proc.stdout + proc.stderr)[-3000:]
    assert proc.returncode == 0, tail
    assert "MDS_CHECKED" not in tail, tail
    assert " passed" in proc.stdout


def test_the_checked_library_is_what_the_subprocess_loaded():
    env = dict(os.environ, MDS_RDF_LIB=CHECKED)
    out = subprocess.check_output(
        [sys.executable, "-c", "from mdsuite_b200 import _cuda; _cuda.load(); print(_cuda.LIB_PATH)"],
        cwd=ROOT, env=env, text=True)
    assert out.strip().endswith("libmds_rdf_checked.so")


def test_the_self_check_fires_on_an_injected_stray_atomic():
    """MDS_CHECKED_INJECT_FAULT=1 aims the all-pairs kernel's histogram atomics at its column tiles:
    the conservation check must notice (device message + trap -> the run fails)."""
    code = ("import numpy as np\n"
            "from mdsuite_b200 import _cuda, synthetic\n"
            "s = synthetic.nacl_melt(3, 2, seed=1)\n"
            "h = _cuda.RDFHandle(s.counts, 100, 8.0, s.box, path='allpairs')\n"
            "h.submit(s.positions)\n"
            "print('sum', int(h.counts().sum()))\n")
    env = dict(os.environ, MDS_RDF_LIB=CHECKED, MDS_CHECKED_INJECT_FAULT="1")
    proc = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True,
                          text=True, timeout=300)
    out = proc.stdout + proc.stderr
    assert proc.returncode != 0, out[-2000:]
    assert "MDS_CHECKED rdf_allpairs_kernel" in out or "failed" in out, out[-2000:]
    env.pop("MDS_CHECKED_INJECT_FAULT")
    proc = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True,
                          text=True, timeout=300)
    assert proc.returncode == 0 and "sum" in proc.stdout, proc.stdout + proc.stderr
