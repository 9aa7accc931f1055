#!/usr/bin/env python
"""Tiny evaluation for compute-sanitizer (memcheck / racecheck / synccheck): a few fragments, chosen kernels."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from ai2bmd_b200.engine import Engine                    # noqa: E402
from ai2bmd_b200.fixtures import WEIGHTS, load_fragments  # noqa: E402
from ai2bmd_b200.weights import load_state_dict          # noqa: E402

edge_tc = int(sys.argv[1]) if len(sys.argv) > 1 else 3
nfrag = int(sys.argv[2]) if len(sys.argv) > 2 else 6
fd, _ = load_fragments("chig")
sub = fd[0:nfrag]
eng = Engine(load_state_dict(WEIGHTS), 0)
eng.set_topology(sub.z, sub.batch)
eng.set_option("use_graph", 0)
eng.set_option("edge_tc", edge_tc)
e, f = eng.forward_host(sub.pos)
e2, f2 = eng.forward_host(sub.pos)
print("edge_tc", edge_tc, "E0", e[0], "|F|max", np.abs(f).max(), "repeat diff", np.abs(f - f2).max())
