"""Packed batch of protein fragments -- the input format of the drop-in boundary.

Host-side mirror of the reference's ``FragmentData`` (``/root/reference/src/AIMD/fragment.py:7-55``):
same field names, same slicing and split semantics, so callers written against the reference
(``DLBondedCalculator.calculate``, ``bonded.py:65-93``) work unchanged.  ASE is not a dependency
here; ``get_atoms`` returns a plain ``(numbers, positions)`` tuple instead of ``ase.Atoms``.
"""
from __future__ import annotations

import numpy as np


class FragmentData:
    """z[N] atomic numbers, pos[N,3] float32 Angstrom, start/end[G] atom ranges, batch[N] graph id
    (sorted, contiguous).  Even fragment slots are dipeptides, odd slots are ACE-NMEs
    (``distancefrag.py:250-284``); ``scalar_split``/``vector_split`` rely on that parity."""

    def __init__(self, z, pos, start, end, batch):
        self.z = z
        self.pos = pos
        self.start = start
        self.end = end
        self.batch = batch
        self._scalar_split = None
        self._vector_split = None

    def __len__(self):
        return len(self.start)

    def __getitem__(self, f_idx):
        # fragment.py:15-29 -- slice of fragments, re-based to start at atom 0 / graph 0
        if isinstance(f_idx, (int, np.integer)):
            f_idx = slice(int(f_idx), int(f_idx) + 1)
        lo, hi, step = f_idx.indices(len(self))
        if step != 1:
            raise IndexError("FragmentData only supports contiguous fragment slices")
        if hi <= lo:
            raise IndexError("empty fragment slice")
        a0, a1 = int(self.start[lo]), int(self.end[hi - 1])
        return FragmentData(
            self.z[a0:a1],
            self.pos[a0:a1],
            self.start[lo:hi] - self.start[lo],
            self.end[lo:hi] - self.start[lo],
            self.batch[a0:a1] - self.batch[a0],
        )

    def scalar_split(self):
        """Masks over the G per-fragment values: (dipeptide slots, ACE-NME slots); fragment.py:31-38."""
        if self._scalar_split is None:
            valid = np.flatnonzero(np.asarray(self.end) - np.asarray(self.start))
            split = np.zeros(len(self), dtype=int)
            split[0::2] = 1
            split = split[valid]
            self._scalar_split = (split == 1, split == 0)
        return self._scalar_split

    def vector_split(self):
        """Masks over the N per-atom rows: (atoms of dipeptides, atoms of ACE-NMEs); fragment.py:40-47."""
        if self._vector_split is None:
            n = int(self.end[-1])
            mark = np.zeros(n + 1, dtype=int)
            np.add.at(mark, np.asarray(self.start[0::2]), 1)
            np.add.at(mark, np.asarray(self.start[1::2]), -1)
            mark = np.cumsum(mark[:n])
            self._vector_split = (mark == 1, mark == 0)
        return self._vector_split

    def get_atoms(self, idx: int):
        s, e = int(self.start[idx]), int(self.end[idx])
        return self.z[s:e], self.pos[s:e]


class FragmentInfo:
    @classmethod
    def split(cls, total):
        # fragment.py:58-61
        return (total + 1) // 2, total // 2
