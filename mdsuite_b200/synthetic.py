"""
Seeded synthetic trajectories of the shapes BASELINE.json names.

All generators return float32 positions wrapped into [0, L), frame-major (F, N, 3), atoms in
species-contiguous order (the order ``_format_data`` produces,
reference calculators/radial_distribution_function.py:535-563), plus the species table and box.
RNG: ``np.random.Generator(np.random.Philox(seed))``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List

import numpy as np


@dataclass
class SyntheticSystem:
    name: str
    positions: np.ndarray  # (F, N, 3) float32, wrapped
    species: Dict[str, int]  # name -> atom count, in storage order
    box: np.ndarray  # (3,) float64

    @property
    def n_atoms(self) -> int:
        return int(self.positions.shape[1])

    @property
    def n_frames(self) -> int:
        return int(self.positions.shape[0])

    @property
    def counts(self) -> List[int]:
        return list(self.species.values())

    def pairs_per_frame(self, parity: bool = True) -> float:
        """The reference's (row, col) pair count per frame; ``parity`` applies quirk Q1."""
        n = [c - (1 if parity else 0) for c in self.counts]
        total = 0.0
        for a in range(len(n)):
            for b in range(a, len(n)):
                total += 0.5 * n[a] * (n[a] - 1) if a == b else n[a] * n[b]
        return total


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(seed))


def _wrap(x: np.ndarray, box: np.ndarray) -> np.ndarray:
    x = x - np.floor(x / box) * box
    x32 = x.astype(np.float32)
    b32 = box.astype(np.float32)
    # float32 rounding can land exactly on L; fold those back to 0
    return np.where(x32 >= b32, x32 - b32, x32).astype(np.float32)


def nacl_melt(n_cells: int, n_frames: int, seed: int, lattice: float = 6.3,
              sigma: float = 0.6) -> SyntheticSystem:
    """Rocksalt n_cells^3 x 8 ions (cfg1: n_cells=5 -> 1 000; cfg2: n_cells=12 -> 13 824), box
    L = n_cells * 6.3 A (rho = 0.0320 A^-3), per-frame independent Gaussian displacements."""
    rng = _rng(seed)
    box = np.full(3, n_cells * lattice)
    cell = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    g = np.stack(np.meshgrid(*[np.arange(n_cells)] * 3, indexing="ij"), -1).reshape(-1, 3)
    na = (g[:, None, :] + cell[None]).reshape(-1, 3) * lattice
    cl = (g[:, None, :] + cell[None] + np.array([0.5, 0, 0])).reshape(-1, 3) * lattice
    base = np.concatenate([na, cl])  # species-contiguous: all Na, then all Cl
    pos = np.empty((n_frames, base.shape[0], 3), dtype=np.float32)
    for f in range(n_frames):
        pos[f] = _wrap(base + rng.normal(0.0, sigma, base.shape), box)
    return SyntheticSystem(f"nacl_{base.shape[0]}", pos, {"Na": len(na), "Cl": len(cl)}, box)


def _lattice_sites(n_side: int, n_keep: int, rng) -> np.ndarray:
    """(n_keep, 3) integer sites of an n_side^3 simple-cubic lattice (random vacancies if fewer)."""
    g = np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), -1).reshape(-1, 3)
    if n_keep < len(g):
        g = g[np.sort(rng.permutation(len(g))[:n_keep])]
    return g


def lj_fluid(n_side: int, n_frames: int, seed: int, rho: float = 0.8442,
             sigma: float = 0.1, n_atoms: int = None) -> SyntheticSystem:
    """Atoms on a jittered simple-cubic lattice at reduced density rho (cfg3): n_side^3 of them, or
    exactly ``n_atoms`` (<= n_side^3, the rest of the sites stay vacant; cfg3 proper is
    lj_fluid(47, F, seed, n_atoms=100_000))."""
    rng = _rng(seed)
    n = n_side ** 3 if n_atoms is None else int(n_atoms)
    box = np.full(3, (n / rho) ** (1.0 / 3.0))
    g = _lattice_sites(n_side, n, rng)
    base = (g + 0.5) * (box[0] / n_side)
    pos = np.empty((n_frames, n, 3), dtype=np.float32)
    for f in range(n_frames):
        pos[f] = _wrap(base + rng.normal(0.0, sigma, base.shape), box)
    return SyntheticSystem(f"lj_{n}", pos, {"Ar": n}, box)


def spc_water(n_side: int, n_frames: int, seed: int, rho_mol: float = 0.0334,
              n_molecules: int = None) -> SyntheticSystem:
    """Rigid SPC molecules (r_OH = 1.0 A, 109.47 deg), random orientations, on a jittered lattice;
    species order O (n), H (2n): n_side^3 molecules, or exactly ``n_molecules`` (cfg4 proper is
    spc_water(70, F, seed, n_molecules=333_333): 999 999 atoms, L = 215.3 A)."""
    rng = _rng(seed)
    n = n_side ** 3 if n_molecules is None else int(n_molecules)
    box = np.full(3, (n / rho_mol) ** (1.0 / 3.0))
    g = _lattice_sites(n_side, n, rng)
    base = (g + 0.5) * (box[0] / n_side)
    half = np.deg2rad(109.47) / 2
    h_local = np.array([[np.sin(half), 0.0, np.cos(half)], [-np.sin(half), 0.0, np.cos(half)]])
    pos = np.empty((n_frames, 3 * n, 3), dtype=np.float32)
    for f in range(n_frames):
        o = base + rng.normal(0.0, 0.3, base.shape)
        q = rng.normal(size=(n, 4))
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        w, x, y, z = q.T
        rot = np.stack([
            np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
            np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
            np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1),
        ], 1)  # (n, 3, 3)
        h = o[:, None, :] + np.einsum("nij,kj->nki", rot, h_local)  # (n, 2, 3)
        pos[f, :n] = _wrap(o, box)
        pos[f, n:] = _wrap(h.reshape(-1, 3), box)
    return SyntheticSystem(f"spc_{3 * n}", pos, {"O": n, "H": 2 * n}, box)


def ideal_gas(n_atoms: int, n_frames: int, seed: int, box_length: float = 100.0) -> SyntheticSystem:
    """Uniform random points (cfg5): g(r) = 1 analytically."""
    rng = _rng(seed)
    box = np.full(3, box_length)
    pos = _wrap(rng.random((n_frames, n_atoms, 3)) * box_length, box)
    return SyntheticSystem(f"ideal_{n_atoms}", pos, {"X": n_atoms}, box)


def write_lammps_dump(system: SyntheticSystem, path: str, timestep_stride: int = 1) -> None:
    """LAMMPS text dump with ``id element x y z`` columns and 9 header lines per frame — what
    the reference's LAMMPSTrajectoryFile reads (file_io/lammps_trajectory_files.py:103-146)."""
    names = []
    for name, count in system.species.items():
        names += [name] * count
    with open(path, "w") as fh:
        for f in range(system.n_frames):
            fh.write("ITEM: TIMESTEP\n%d\n" % (f * timestep_stride))
            fh.write("ITEM: NUMBER OF ATOMS\n%d\n" % system.n_atoms)
            fh.write("ITEM: BOX BOUNDS pp pp pp\n")
            for k in range(3):
                fh.write("%.9g %.9g\n" % (0.0, system.box[k]))
            fh.write("ITEM: ATOMS id element x y z\n")
            p = system.positions[f]
            fh.write("".join("%d %s %.9g %.9g %.9g\n" % (i + 1, names[i], p[i, 0], p[i, 1], p[i, 2])
                             for i in range(system.n_atoms)))


def write_extxyz(system: SyntheticSystem, path: str, time_stride: float = 1.0,
                 interleave: bool = True) -> None:
    """Extended XYZ file with ``species pos`` columns — what the reference's EXTXYZFile reads
    (file_io/extxyz_files.py:93-181).  ``interleave`` mixes the species along the file (the reader
    has to gather them); positions are written with 9 significant digits (float32 round trip)."""
    names = []
    for name, count in system.species.items():
        names += [name] * count
    order = np.arange(system.n_atoms)
    if interleave:
        order = np.random.Generator(np.random.Philox(17)).permutation(system.n_atoms)
    lattice = "%.9g 0.0 0.0 0.0 %.9g 0.0 0.0 0.0 %.9g" % tuple(system.box)
    with open(path, "w") as fh:
        for f in range(system.n_frames):
            fh.write("%d\n" % system.n_atoms)
            fh.write('Lattice="%s" Properties=species:S:1:pos:R:3 time=%.9g pbc="T T T"\n'
                     % (lattice, f * time_stride))
            p = system.positions[f]
            fh.write("".join("%s %.9g %.9g %.9g\n" % (names[i], p[i, 0], p[i, 1], p[i, 2])
                             for i in order))
