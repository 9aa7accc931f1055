"""Loaders for the committed fixture files under ``tests/golden`` (used by tests, bench.py and tools)."""
from __future__ import annotations

import os

import numpy as np

from .fragment_data import FragmentData
from .pdbfrag import FragmentRecipe, ProteinMap

GOLDEN = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
WEIGHTS = os.path.join(GOLDEN, "weights_2ef43f29.npz")


def load_fragments(name: str):
    g = np.load(os.path.join(GOLDEN, f"fragments_{name}.npz"))
    fd = FragmentData(g["z"], g["pos"], g["start"], g["end"], g["batch"])
    pm = ProteinMap(int(g["n_protein"]), g["src_atom"], g["dst_atom"], g["sign"], g["frag_sign"])
    return fd, pm


def load_protein(name: str):
    """(protein positions [n,3] f64, atomic numbers [n], FragmentRecipe) of an example protein."""
    g = np.load(os.path.join(GOLDEN, f"fragments_{name}.npz"))
    return g["prot_pos"], g["prot_z"], FragmentRecipe(g["rc_real"], g["rc_acc"], g["rc_rem"], g["rc_blen"])
