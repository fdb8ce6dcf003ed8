"""Per-kernel SASS summary of the in-tree library (cuobjdump, no GPU needed): instruction mix of
every RDF kernel and the size of every innermost loop that bins (instructions per ATOMS), the
numbers DESIGN.md quotes and tests/test_sass_cpu.py guards.

    python scripts/sass_summary.py > profiles/r2_sass_instruction_mix.txt
"""
import collections
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdsuite_b200 import _cuda  # noqa: E402

cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
out = subprocess.run([cuobjdump, "-sass", _cuda.LIB_PATH], check=True, capture_output=True, text=True).stdout
demangle = shutil.which("c++filt")
kernels, name = {}, None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        kernels[name] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;", line)
    if name and m:
        kernels[name].append((int(m.group(1), 16), m.group(2)))


def mnemonic(text):
    parts = text.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    return op.split(".")[0]


print(f"library: {os.path.relpath(_cuda.LIB_PATH, ROOT)}  (cuobjdump -sass)\n")
for name in sorted(kernels):
    if "rdf_" not in name:
        continue
    ins = kernels[name]
    pretty = name
    if demangle:
        pretty = subprocess.run([demangle, name], capture_output=True, text=True).stdout.strip() or name
    mix = collections.Counter(mnemonic(t) for _, t in ins)
    print(f"{pretty}\n  {len(ins)} instructions; FFMA2 {mix.get('FFMA2', 0)}, FADD2 {mix.get('FADD2', 0)}, "
          f"FMUL2 {mix.get('FMUL2', 0)}, ATOMS {mix.get('ATOMS', 0)}, UBLKCP {mix.get('UBLKCP', 0)}, "
          f"LDGSTS {mix.get('LDGSTS', 0)}, MUFU {mix.get('MUFU', 0)}")
    addr = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\s+(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                body = ins[addr[tgt]:i + 1]
                n_atoms = sum("ATOMS.POPC" in x for _, x in body)
                if n_atoms and len(body) < 400:
                    loops.append((len(body), n_atoms, collections.Counter(mnemonic(x) for _, x in body)))
    for size, n_atoms, lm in loops:
        top = ", ".join(f"{k} {v}" for k, v in lm.most_common(14))
        print(f"  binning loop: {size} instructions per {n_atoms} pairs = {size / n_atoms:.2f} per pair  [{top}]")
    print()
