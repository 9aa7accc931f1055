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


def load_capped_protein(name: str):
    """The example protein as a :class:`ai2bmd_b200.pdbfrag.CappedProtein` (names, residues, elements, positions)."""
    from .pdbfrag import CappedProtein
    g = np.load(os.path.join(GOLDEN, f"fragments_{name}.npz"))
    el = {1: "H", 6: "C", 7: "N", 8: "O", 16: "S"}
    return CappedProtein([str(x) for x in g["prot_names"]], [str(x) for x in g["prot_resnames"]], g["prot_resnums"].astype(np.int64),
                         [el[int(z)] for z in g["prot_z"]], g["prot_pos"])


def load_caph_tables(name: str):
    """[(parsed prmtop table, atom names)] per dipeptide of an example protein, as stored next to the reference's own
    refinement output (tests/golden/reference_caph_batch*.npz), plus that file."""
    g = np.load(os.path.join(GOLDEN, "reference_caph_batch.npz" if name == "chig" else f"reference_caph_batch_{name}.npz"))
    tables = []
    for k in range(int(g["n_graphs"])):
        pre = f"g{k}_t_"
        t = {n[len(pre):]: (int(g[n]) if g[n].ndim == 0 else g[n]) for n in g.files if n.startswith(pre)}
        tables.append((t, [str(x) for x in g[f"g{k}_names"]]))
    return tables, g


def load_synthetic_cyx():
    """The WW-domain example with four SER turned into CYX (OG -> SG) -- the reference ships no CYX-containing example --
    exactly as tests/golden/make_golden.py builds it for the reference's own pairing function; returns (protein, golden)."""
    import json
    from .pdbfrag import CappedProtein
    prot = load_capped_protein("ww")
    gold = json.load(open(os.path.join(GOLDEN, "reference_cyx.json")))
    names, resn, elem = list(prot.names), list(prot.resnames), list(prot.elements)
    for i in range(len(prot)):
        if int(prot.resnums[i]) in gold["cyx_residues"]:
            resn[i] = "CYX"
            if names[i] == "OG":
                names[i], elem[i] = "SG", "S"
    return CappedProtein(names, resn, prot.resnums, elem, prot.positions), gold
