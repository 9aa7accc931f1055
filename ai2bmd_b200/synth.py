"""Synthetic fragment batches of the shapes BASELINE.json names.

Templates are the 220 real fragments cut from the four example proteins (``tests/golden/fragments_*.npz``);
every instance is a template under a random rotation (uniform SO(3)) plus Gaussian jitter, so the batches
have realistic element mix, neighbour counts (mean degree ~17 incl. self-loop) and fragment sizes (12-36 atoms).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .fixtures import load_fragments
from .fragment_data import FragmentData
from .pdbfrag import ProteinMap


def _templates(dipeptides_only: bool) -> List[Tuple[np.ndarray, np.ndarray]]:
    out = []
    for name in ("chig", "trpcage", "ww", "abd"):
        fd, _ = load_fragments(name)
        for g in range(len(fd)):
            if dipeptides_only and g % 2 == 1:
                continue
            s, e = int(fd.start[g]), int(fd.end[g])
            p = fd.pos[s:e].astype(np.float64)
            out.append((fd.z[s:e].copy(), p - p.mean(0)))
    return out


def _random_rotation(rng) -> np.ndarray:
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    a, b, c, d = q
    return np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                     [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                     [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])


def _assemble(picks, rng, jitter) -> FragmentData:
    z, pos, batch, start, end = [], [], [], [], []
    off = 0
    for g, (tz, tp) in enumerate(picks):
        p = tp @ _random_rotation(rng).T + rng.normal(scale=jitter, size=tp.shape)
        z.append(tz)
        pos.append(p)
        batch.append(np.full(len(tz), g))
        start.append(off)
        off += len(tz)
        end.append(off)
    return FragmentData(np.concatenate(z).astype(np.int64), np.concatenate(pos).astype(np.float32),
                        np.asarray(start, dtype=np.int64), np.asarray(end, dtype=np.int64),
                        np.concatenate(batch).astype(np.int64))


def synthetic_batch(n_fragments: int = 512, seed: int = 0, jitter: float = 0.05, min_atoms: int = 0) -> FragmentData:
    """Config C4: ``n_fragments`` dipeptide templates drawn with ``default_rng(seed)`` (~14k atoms for 512);
    with ``min_atoms`` keep drawing until the batch holds at least that many atoms (the "~20k atoms" variant)."""
    rng = np.random.default_rng(seed)
    tmpl = _templates(dipeptides_only=True)
    picks = [tmpl[i] for i in rng.integers(0, len(tmpl), size=n_fragments)]
    while min_atoms and sum(len(t[0]) for t in picks) < min_atoms:
        picks.append(tmpl[int(rng.integers(0, len(tmpl)))])
    return _assemble(picks, rng, jitter)


def conformer_batch(n_conformers: int = 2048, seed: int = 1, jitter: float = 0.1) -> FragmentData:
    """Config C5: dipeptide templates cycled in order, jitter 0.1 A (inference-only throughput shape)."""
    rng = np.random.default_rng(seed)
    tmpl = _templates(dipeptides_only=True)
    picks = [tmpl[i % len(tmpl)] for i in range(n_conformers)]
    return _assemble(picks, rng, jitter)


def synthetic_protein_map(frags: FragmentData) -> ProteinMap:
    """A whole-"protein" reduction map for synthetic batches: N_prot = N/2 destinations, alternating signs."""
    n = len(frags.z)
    frag_sign = np.where(np.arange(len(frags)) % 2 == 0, 1.0, -1.0).astype(np.float32)
    src = np.arange(n, dtype=np.int32)
    dst = (src // 2).astype(np.int32)
    sign = frag_sign[np.asarray(frags.batch)]
    return ProteinMap((n + 1) // 2, src, dst, sign.astype(np.float32), frag_sign)
