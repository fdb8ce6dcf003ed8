"""The bit-exactness of the pair kernels rests on how ptxas lowers the packed float32 PTX
(rdf_common.cuh): `sub.rn.f32x2` / `mul.rn.f32x2` must stay FADD2 / FMUL2 and must NOT be
contracted into FFMA2 — a fused multiply-add would skip the separate rounding of the products
that the reference's `(x*x + y*y) + z*z` has (reference calculators/radial_distribution_function.py:
686, utils/linalg.py:99).  This test disassembles the in-tree library and checks exactly that, plus
the sm_100a-native instructions the design relies on (TMA bulk copies, shared-memory atomics that
merge lanes).  No GPU needed: cuobjdump reads the cubin."""
import os
import re
import shutil
import subprocess

import pytest

from mdsuite_b200 import _cuda

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


def _sass_by_kernel():
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(_cuda.LIB_PATH):
        pytest.fail(f"{_cuda.LIB_PATH} has not been built")
    out = subprocess.run([CUOBJDUMP, "-sass", _cuda.LIB_PATH], check=True, capture_output=True,
                         text=True).stdout
    kernels, name = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            kernels[name].append(line.split("*/", 1)[1].split("/*")[0].strip())
    return kernels


def _count(instrs, mnemonic):
    return sum(1 for i in instrs if re.search(r"(^|\s)%s(\.|\s|$)" % re.escape(mnemonic), i))


def test_pair_kernels_keep_separately_rounded_products():
    kernels = _sass_by_kernel()
    pair = {k: v for k, v in kernels.items() if "rdf_allpairs_kernel" in k or "rdf_cells_kernel" in k}
    assert len(pair) >= 4, sorted(kernels)
    arch = subprocess.run([CUOBJDUMP, "-lelf", _cuda.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in arch
    for name, instrs in pair.items():
        assert _count(instrs, "FFMA2") == 0, f"{name}: ptxas contracted packed mul + add into FFMA2"
        assert _count(instrs, "FADD2") > 0 and _count(instrs, "FMUL2") > 0, name
        # histograms in shared memory (the <.., true> instantiations) count with shared atomics
        if "Lb1EEE" in name:
            assert _count(instrs, "ATOMS") > 0, name


def test_design_relies_on_sm100_instructions():
    kernels = _sass_by_kernel()
    allpairs = [v for k, v in kernels.items() if "rdf_allpairs_kernel" in k]
    cells = [v for k, v in kernels.items() if "rdf_cells_kernel" in k]
    assert allpairs and cells
    assert all(_count(k, "UBLKCP") > 0 for k in allpairs)      # cp.async.bulk (TMA 1-D) column tiles
    assert all(_count(k, "LDGSTS") > 0 for k in cells)         # cp.async staging of the slabs
    smem_hist = [v for k, v in kernels.items()
                 if ("rdf_allpairs_kernel" in k and k.endswith("Lb1EEEvNS_14AllPairsParamsE"))
                 or "rdf_cells_kernelILb1" in k]
    assert smem_hist
    assert all(any("ATOMS.POPC.INC" in i for i in k) for k in smem_hist)
    assert not any(i.startswith(("HMMA", "IMMA", "UTC")) for k in allpairs + cells for i in k)
