"""Per-stage parity (SURVEY section 8 rows a3-a18 one by one): the evaluation is run one launch at a time
(``vb_debug_run``) and every buffer a stage produces is compared with the fp64 hand-adjoint oracle
(``oracle/adjoint_ref.py``, equal to autograd at 1e-14).  Both launch plans are covered: the separate node / edge
stages and the fused per-layer kernels.  Tolerance: 2e-3 relative to the largest reference entry of the buffer
(fp32 + 3xTF32 against fp64; measured 1e-6 .. 3e-4, adjoint buffers deep in the reverse sweep being the largest)."""
import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("opts", ["fused=0", "fused=1", "fused=0,edge_tc=0", "node_tc=1",
                                  "node_nb=2", "node_nb=3", "node_nb=4", "node_nb=8"])     # default here: node_nb=1 (one wave)
@pytest.mark.parametrize("weights", ["real", "3"])
def test_every_stage_against_the_fp64_adjoint_oracle(opts, weights):
    from stage_check import stage_report
    lines, worst = stage_report("chig", weights, max_frags=4, opts=opts)
    bad = [(s, w, r) for s, w, r in worst if not r <= 2e-3]
    assert not bad, "\n".join(lines)
    stages = {s for s, _, _ in worst}
    assert "head" in stages and "embed_node_bwd" in stages and "finalize" in stages
    assert ("fwd3" in stages) == ("fused=1" in opts)
    assert ("proj3" in stages and "bwdB2" in stages) == ("node_tc=1" in opts)
