"""Host side of the per-step cap-hydrogen refinement (``csrc/k_caph.cuh`` behind ``vb_set_caph`` / ``vb_caph_relax``).

The reference relaxes, every MD step, the added hydrogens of all dipeptides with one LBFGS call on Amber terms taken from
per-residue ``.prmtop`` tables (``/root/reference/src/Fragmentation/hydrogen/ctable.py:58-240``, ``energies.py:9-60,
211-242``; called from ``distancefrag.py:56-92``).  This module turns those tables and the fragmentation into the flat term
arrays the device kernel consumes -- one-off setup, plain numpy:

* :func:`parse_prmtop` / :func:`prmtop_atom_names`   the fields ``CTable.from_prmtop`` reads (``ctable.py:58-170``);
* :func:`hydrogen_terms`                             the terms that involve an added hydrogen (``ctable.py:172-240``);
* :func:`table_layout`                               which protein atom / added hydrogen sits at every position of a
                                                     dipeptide's table, BY ATOM NAME (every table is ACE + residue + NME;
                                                     the reference reaches the same order through its per-residue index
                                                     tables, ``distancefrag.py:506-737``);
* :func:`build_problem`                              flat arrays over the PACKED FRAGMENT atoms + the ACE-NME mirror list.

Amber units throughout (kcal/mol, Angstrom, radians), as the tables hold them.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np

from .fragment_data import FragmentData
from .pdbfrag import CappedProtein, FragmentRecipe

SCNB, SCEE = 1.2, 2.0                     # HydrogenOptimizer defaults, energies.py:76-80
LBFGS = dict(max_iter=10, lr=0.1, tol_grad=0.1, tol_change=0.01)       # energies.py:211-242
# residue name -> prmtop file stem (src/Fragmentation/prmtop/<stem>.prmtop)
PRMTOP_STEM = {"ALA": "AA", "ARG": "RR", "ASP": "DD", "CYS": "CC", "GLN": "QQ", "GLU": "EE", "GLY": "GG", "LYS": "KK",
               "ASN": "NN", "LEU": "LL", "PRO": "PP", "SER": "SS", "THR": "TT", "VAL": "VV", "MET": "MM", "HIS": "HH",
               "HIE": "HH", "HID": "HID", "TRP": "WW", "TYR": "YY", "ILE": "II", "PHE": "FF"}


# ---------------------------------------------------------------------------------------------------------
# prmtop tables
# ---------------------------------------------------------------------------------------------------------
def _flags(text: str) -> Dict[str, List[str]]:
    out, lines, i = {}, text.splitlines(), 0
    while i < len(lines):
        if lines[i].startswith("%FLAG"):
            name, j, raw = lines[i].split()[1], i + 2, []
            while j < len(lines) and not lines[j].startswith("%"):
                raw.append(lines[j].rstrip("\n"))
                j += 1
            out[name] = raw
            i = j
        else:
            i += 1
    return out


def prmtop_atom_names(text: str) -> List[str]:
    """ATOM_NAME entries (fixed 4-character fields)."""
    names = []
    for ln in _flags(text)["ATOM_NAME"]:
        names.extend(ln[k:k + 4].strip() for k in range(0, len(ln), 4) if ln[k:k + 4].strip())
    return names


def parse_prmtop(text: str) -> dict:
    """The sections ``CTable.from_prmtop`` uses, as numpy arrays with 0-based indices."""
    fl = {k: " ".join(v).split() for k, v in _flags(text).items()}

    def f(name):
        return np.array([float(v) for v in fl[name]], dtype=np.float32)

    def n(name):
        return np.array([int(v) for v in fl[name]], dtype=np.int64)

    ptr = n("POINTERS")
    t = {"natom": int(ptr[0]), "ntypes": int(ptr[1]), "numbnd": int(ptr[15]), "numang": int(ptr[16]), "nptra": int(ptr[17]),
         "charge": f("CHARGE"), "atomic_number": n("ATOMIC_NUMBER"), "atom_type_idx": n("ATOM_TYPE_INDEX") - 1,
         "number_excluded_atoms": n("NUMBER_EXCLUDED_ATOMS"), "nonbonded_parm_index": n("NONBONDED_PARM_INDEX") - 1,
         "bond_force_constant": f("BOND_FORCE_CONSTANT"), "bond_equil_value": f("BOND_EQUIL_VALUE"),
         "angle_force_constant": f("ANGLE_FORCE_CONSTANT"), "angle_equil_value": f("ANGLE_EQUIL_VALUE"),
         "dihedral_force_constant": f("DIHEDRAL_FORCE_CONSTANT"), "dihedral_periodicity": f("DIHEDRAL_PERIODICITY"),
         "dihedral_phase": f("DIHEDRAL_PHASE"), "lennard_jones_acoef": f("LENNARD_JONES_ACOEF"),
         "lennard_jones_bcoef": f("LENNARD_JONES_BCOEF"), "excluded_atoms_list": n("EXCLUDED_ATOMS_LIST") - 1}
    # coordinate-array offsets (3 * atom index; dihedral sign flags stay negative under floor division), 1-based type last
    for key, flag, w in (("bonds_inc_hydrogen", "BONDS_INC_HYDROGEN", 3), ("angles_inc_hydrogen", "ANGLES_INC_HYDROGEN", 4),
                        ("dihedrals_inc_hydrogen", "DIHEDRALS_INC_HYDROGEN", 5)):
        a = n(flag).reshape(-1, w).copy() if fl.get(flag) else np.zeros((0, w), np.int64)
        a[:, :-1] = np.floor_divide(a[:, :-1], 3)
        a[:, -1] -= 1
        t[key] = a
    return t


def hydrogen_terms(t: dict, atom_idx) -> dict:
    """Bonds / angles / dihedrals that contain one of ``atom_idx`` and the non-excluded pairs between them and every
    other atom of the table (``ctable.py:172-240``); pairs in canonical order (i < j, ascending)."""
    atom_idx = np.asarray(atom_idx, dtype=np.int64)
    sel = np.zeros(t["natom"], dtype=bool)
    sel[atom_idx] = True
    b, a, d = t["bonds_inc_hydrogen"], t["angles_inc_hydrogen"], t["dihedrals_inc_hydrogen"]
    bm = sel[b[:, :2]].any(1) if len(b) else np.zeros(0, bool)
    am = sel[a[:, :3]].any(1) if len(a) else np.zeros(0, bool)
    dm = (np.isin(d[:, :4], atom_idx).any(1) & (d[:, 2:4] >= 0).all(1)) if len(d) else np.zeros(0, bool)
    ptr = np.concatenate([[0], np.cumsum(t["number_excluded_atoms"])])
    excl = np.zeros((t["natom"], t["natom"]), dtype=bool)
    for i in range(t["natom"]):
        for j in t["excluded_atoms_list"][ptr[i]:ptr[i + 1]]:
            if j >= 0:
                excl[i, int(j)] = True
    ii, jj = np.triu_indices(t["natom"], k=1)
    keep = (sel[ii] | sel[jj]) & ~excl[ii, jj]
    pairs = np.stack([ii[keep], jj[keep]], axis=1).astype(np.int64)
    lj_idx = t["nonbonded_parm_index"][t["ntypes"] * t["atom_type_idx"][pairs[:, 0]] + t["atom_type_idx"][pairs[:, 1]]]
    return {"bonds": b[bm], "angles": a[am], "dihedrals": d[dm], "pairs": pairs, "lj_idx": lj_idx}


# ---------------------------------------------------------------------------------------------------------
# which atom sits where in a dipeptide's table
# ---------------------------------------------------------------------------------------------------------
def table_layout(prot: CappedProtein, centre: int, names: Sequence[str]) -> List[Tuple]:
    """For the dipeptide centred on residue ``centre`` (1-based numbering of the capped chain): one entry per table atom,
    ``("real", protein_index)`` or ``("cap", acceptor_protein_index, removed_protein_index)`` for an added hydrogen.

    Every table is ACE ``[H1 CH3 H2 H3 C O]`` + the residue (its own atom names) + NME ``[N H CH3 HH31 HH32 HH33]``.  The ACE
    methyl of an inner dipeptide is the previous residue's CA with its HA atom(s) first and then the hydrogens added in
    place of N (and CB); the NME methyl is the next residue's CA, HA atom(s), then the hydrogens replacing C (and CB); a
    PRO neighbour has no amide H, its place is taken by the hydrogen added on the N->CD ray."""
    by_res = [i for i in range(len(prot)) if prot.resnums[i] == centre]

    def members(res):
        return [i for i in range(len(prot)) if prot.resnums[i] == res]

    def at(res, name):
        hit = [i for i in members(res) if prot.names[i] == name]
        return hit[0] if hit else None

    def resname(res):
        return prot.resnames[members(res)[0]]

    def methyl(res, removed):
        ca = at(res, "CA")
        hs = [("real", i) for i in members(res) if prot.names[i].startswith("HA")]
        hs.append(("cap", ca, at(res, removed)))
        if resname(res) != "GLY":
            hs.append(("cap", ca, at(res, "CB")))
        if len(hs) != 3:
            raise ValueError(f"residue {res} ({resname(res)}): cannot form a cap methyl")
        return hs

    lead, trail = centre - 1, centre + 1
    rows: List[Tuple] = []
    if resname(lead) == "ACE":
        hs = [("real", i) for i in members(lead) if prot.elements[i] == "H"]
        ch3 = [i for i in members(lead) if prot.elements[i] == "C" and prot.names[i] != "C"][0]
        rows += [hs[0], ("real", ch3), hs[1], hs[2], ("real", at(lead, "C")), ("real", at(lead, "O"))]
    else:
        hs = methyl(lead, "N")
        rows += [hs[0], ("real", at(lead, "CA")), hs[1], hs[2], ("real", at(lead, "C")), ("real", at(lead, "O"))]
    n_res = len(names) - 12
    for nm in names[6:6 + n_res]:
        hit = [i for i in by_res if prot.names[i] == nm]
        if not hit:
            raise ValueError(f"residue {centre} ({resname(centre)}) has no atom named {nm!r}")
        rows.append(("real", hit[0]))
    if resname(trail) == "NME":
        hs = [("real", i) for i in members(trail) if prot.elements[i] == "H" and prot.names[i] != "H"]
        ch3 = [i for i in members(trail) if prot.elements[i] == "C"][0]
        rows += [("real", at(trail, "N")), ("real", at(trail, "H")), ("real", ch3)] + hs
    else:
        hs = methyl(trail, "C")
        nh = ("real", at(trail, "H")) if at(trail, "H") is not None else ("cap", at(trail, "N"), at(trail, "CD"))
        rows += [("real", at(trail, "N")), nh, ("real", at(trail, "CA"))] + hs
    if len(rows) != len(names):
        raise ValueError(f"dipeptide {centre}: {len(rows)} atoms, table has {len(names)}")
    return rows


# ---------------------------------------------------------------------------------------------------------
# flat problem over the packed fragment atoms
# ---------------------------------------------------------------------------------------------------------
@dataclass
class CapHProblem:
    """Arguments of ``vb_set_caph`` (numpy arrays, indices into the packed fragment atoms) + bookkeeping for tests."""
    h_idx: np.ndarray
    bond_ij: np.ndarray; bond_k: np.ndarray; bond_r0: np.ndarray
    angle_ijk: np.ndarray; angle_k: np.ndarray; angle_t0: np.ndarray
    dih_ijkl: np.ndarray; dih_k: np.ndarray; dih_n: np.ndarray; dih_p: np.ndarray
    pair_ij: np.ndarray; pair_a: np.ndarray; pair_b: np.ndarray; pair_qq: np.ndarray
    mirror_dst: np.ndarray; mirror_src: np.ndarray
    table_to_frag: List[np.ndarray]         # per dipeptide: packed fragment atom of every table atom
    scnb: float = SCNB
    scee: float = SCEE
    max_iter: int = LBFGS["max_iter"]
    lr: float = LBFGS["lr"]
    tol_grad: float = LBFGS["tol_grad"]
    tol_change: float = LBFGS["tol_change"]


def _fragment_lookup(recipe: FragmentRecipe, lo: int, hi: int):
    real, cap = {}, {}
    for a in range(lo, hi):
        if recipe.real[a] >= 0:
            real[int(recipe.real[a])] = a
        else:
            cap[(int(recipe.acc[a]), int(recipe.rem[a]))] = a
    return real, cap


def build_problem(prot: CappedProtein, frags: FragmentData, recipe: FragmentRecipe,
                  tables: Sequence[Tuple[dict, Sequence[str]]]) -> CapHProblem:
    """``tables[k]`` = (parsed prmtop table, its atom names) of dipeptide k (centred on residue k + 2).  Fragments are
    packed dipeptide 0, ACE-NME 0, dipeptide 1, ... (``distancefrag.py:250-284``): dipeptide k is fragment 2k."""
    n_dip = (len(frags) + 1) // 2
    if len(tables) != n_dip:
        raise ValueError(f"{len(tables)} tables for {n_dip} dipeptides")
    acc = {k: [] for k in ("h_idx", "bond_ij", "bond_k", "bond_r0", "angle_ijk", "angle_k", "angle_t0", "dih_ijkl", "dih_k",
                           "dih_n", "dih_p", "pair_ij", "pair_a", "pair_b", "pair_qq")}
    t2f_all, cap_owner = [], {}
    for k, (t, names) in enumerate(tables):
        lo, hi = int(frags.start[2 * k]), int(frags.end[2 * k])
        real, cap = _fragment_lookup(recipe, lo, hi)
        layout = table_layout(prot, k + 2, names)
        t2f = np.empty(len(layout), dtype=np.int64)
        hs = []
        for pos_in_table, ent in enumerate(layout):
            if ent[0] == "real":
                t2f[pos_in_table] = real[ent[1]]
            else:
                t2f[pos_in_table] = cap[(ent[1], ent[2])]
                hs.append(pos_in_table)
                cap_owner[(ent[1], ent[2])] = int(t2f[pos_in_table])
        if len(set(t2f.tolist())) != hi - lo:
            raise ValueError(f"dipeptide {k}: the table does not cover the fragment's atoms one to one")
        t2f_all.append(t2f)
        terms = hydrogen_terms(t, np.asarray(hs, dtype=np.int64))
        b, a, d, pr, lj = terms["bonds"], terms["angles"], terms["dihedrals"], terms["pairs"], terms["lj_idx"]
        q = t["charge"].astype(np.float32)
        acc["h_idx"].append(t2f[hs])
        acc["bond_ij"].append(t2f[b[:, :2]]); acc["bond_k"].append(t["bond_force_constant"][b[:, 2]]); acc["bond_r0"].append(t["bond_equil_value"][b[:, 2]])
        acc["angle_ijk"].append(t2f[a[:, :3]]); acc["angle_k"].append(t["angle_force_constant"][a[:, 3]]); acc["angle_t0"].append(t["angle_equil_value"][a[:, 3]])
        acc["dih_ijkl"].append(t2f[d[:, :4]]); acc["dih_k"].append(t["dihedral_force_constant"][d[:, 4]])
        acc["dih_n"].append(t["dihedral_periodicity"][d[:, 4]]); acc["dih_p"].append(t["dihedral_phase"][d[:, 4]])
        acc["pair_ij"].append(t2f[pr]); acc["pair_a"].append(t["lennard_jones_acoef"][lj]); acc["pair_b"].append(t["lennard_jones_bcoef"][lj])
        acc["pair_qq"].append(q[pr[:, 0]] * q[pr[:, 1]])
    # ACE-NME fragments (odd slots): an added hydrogen there is the copy of the dipeptide hydrogen with the same
    # (acceptor, removed atom); real atoms are placed from the protein directly
    m_dst, m_src = [], []
    for g in range(1, len(frags), 2):
        for a in range(int(frags.start[g]), int(frags.end[g])):
            if recipe.real[a] < 0:
                m_dst.append(a)
                m_src.append(cap_owner[(int(recipe.acc[a]), int(recipe.rem[a]))])
    width = {"bond_ij": 2, "angle_ijk": 3, "dih_ijkl": 4, "pair_ij": 2}
    out = {}
    for key, parts in acc.items():
        is_idx = key in width or key == "h_idx"
        arr = np.concatenate(parts) if parts else np.zeros(0)
        if key in width:
            arr = arr.reshape(-1, width[key])
        out[key] = np.ascontiguousarray(arr, dtype=np.int32 if is_idx else np.float32)
    return CapHProblem(mirror_dst=np.asarray(m_dst, dtype=np.int32), mirror_src=np.asarray(m_src, dtype=np.int32),
                       table_to_frag=t2f_all, **out)
