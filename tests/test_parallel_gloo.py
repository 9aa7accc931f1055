"""world_size-2 gloo test of the sharded evaluation + all-reduce (host logic of the multi-GPU path).

The per-rank evaluator is the CPU oracle (the checker); the engine itself needs a GPU and is covered by
the -m gpu tests.  What is exercised here: partitioning, shard-local protein maps, the single all-reduce
of the [3*N_prot+1] buffer, and equality with the unsharded result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ai2bmd_b200.fixtures import load_fragments
    from ai2bmd_b200.parallel import ShardedBondedCalculator
    from ai2bmd_b200.weights import load_state_dict
    from oracle import visnet_ref as O
    fd, pm = load_fragments("chig")
    sd = load_state_dict(os.path.join(ROOT, "tests", "golden", "weights_2ef43f29.npz"))
    oracle = O.OracleCalculatorModel({k: torch.from_numpy(v) for k, v in sd.items()})
    calc = ShardedBondedCalculator(fd, pm, rank, world)
    e, f = calc.evaluate_host(fd, oracle.dl_potential_loader)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), e=e, f=f, lo=calc.lo, hi=calc.hi)
    dist.destroy_process_group()


def test_two_rank_sharded_evaluation_equals_single(tmp_path, chig, real_weights):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert r0["hi"] == r1["lo"] and r0["lo"] == 0 and r1["hi"] == 19
    assert np.array_equal(r0["f"], r1["f"]) and r0["e"] == r1["e"]          # all-reduce leaves every rank with the sum
    from ai2bmd_b200.parallel import combine_local
    from oracle import visnet_ref as O
    fd, pm = chig
    oracle = O.OracleCalculatorModel({k: torch.from_numpy(v) for k, v in real_weights.items()})
    e, f = oracle.dl_potential_loader(fd)
    ef = combine_local(pm, e, f)
    assert np.abs(ef[:-1].reshape(-1, 3) - r0["f"]).max() < 5e-5
    assert abs(float(ef[-1]) - float(r0["e"])) < 5e-2        # fp32 sums of O(1e4) eV fragment energies
