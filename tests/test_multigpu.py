"""N > 1 on real GPUs (skipped on boxes with one GPU): fragments sharded over ranks, the whole-protein buffer combined by
the engine's own one-shot all-reduce over NVLink peer memory (csrc/k_comm.cuh) inside the step graph, checked against

* the reference-source golden vectors combined on the host (tests/golden/reference_outputs.npz), same tolerance as N = 1;
* the same step with ``torch.distributed.all_reduce`` (NCCL) instead of the peer-memory kernel;
* the single-GPU device-resident MD trajectory (every rank integrates the replicated state in lock-step).

Run by ``pytest -m gpu`` on a multi-GPU box, or directly:  torchrun-free, the test spawns its own ranks."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from ai2bmd_b200.fixtures import load_fragments, load_protein
    from ai2bmd_b200.md import DeviceLangevin
    from ai2bmd_b200.parallel import DeviceShard, combine_local
    from ai2bmd_b200.pdbfrag import FragmentRecipe
    from ai2bmd_b200.weights import load_state_dict
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sd = load_state_dict(os.path.join(ROOT, "tests", "golden", "weights_2ef43f29.npz"))
    res = {}
    for name in ("chig", "trpcage"):
        fd, pm = load_fragments(name)
        sh = DeviceShard(sd, fd, pm, rank, world, rank, native_comm=True)
        res[f"{name}_native"] = sh.native
        for _ in range(3):                                   # several steps: both window parities, graph replays
            ef = sh.step().clone()
        torch.cuda.synchronize()
        nccl = DeviceShard(sd, fd, pm, rank, world, rank, native_comm=False)
        ef2 = nccl.step().clone()
        torch.cuda.synchronize()
        gathered = [torch.empty_like(ef) for _ in range(world)]
        dist.all_gather(gathered, ef)                        # the fixed-order sum is bit-identical on every rank
        res[f"{name}_identical_on_all_ranks"] = all(bool((g == ef).all()) for g in gathered)
        res[f"{name}_ef"], res[f"{name}_ef_nccl"] = ef.cpu().numpy(), ef2.cpu().numpy()
        del sh, nccl
    # device-resident MD, replicated state, one graph replay per step including the all-reduce
    fd, pm = load_fragments("chig")
    prot_pos, prot_z, recipe = load_protein("chig")
    sh = DeviceShard(sd, fd, pm, rank, world, rank, native_comm=True)
    lo, hi = sh.plan.atom_lo, sh.plan.atom_hi
    rec = FragmentRecipe(recipe.real[lo:hi], recipe.acc[lo:hi], recipe.rem[lo:hi], recipe.blen[lo:hi])
    md = DeviceLangevin(None, None, pm, rec, prot_pos, prot_z, dt_fs=1.0, temperature_K=300.0, friction_per_fs=0.001, seed=0,
                        device=rank, group=dist.group.WORLD, engine=sh.engine)
    md.run(20)
    x, v, step, _ = md.state()
    res["md_x"], res["md_step"], res["md_one_graph"] = x, step, md._native_comm
    if rank == 0:
        np.savez(out, **res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
def test_sharded_evaluation_peer_memory_allreduce(tmp_path, real_weights):
    import torch.multiprocessing as mp
    from ai2bmd_b200.fixtures import load_fragments, load_protein
    from ai2bmd_b200.md import DeviceLangevin
    from ai2bmd_b200.parallel import combine_local
    world = min(torch.cuda.device_count(), 8)
    out = str(tmp_path / "ranks.npz")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r = np.load(out)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.npz"))
    for name in ("chig", "trpcage"):
        fd, pm = load_fragments(name)
        ref = combine_local(pm, gold[f"{name}_ref_e"].reshape(-1), gold[f"{name}_ref_f"])
        assert bool(r[f"{name}_native"]), "peer-memory all-reduce could not be set up on this box"
        assert bool(r[f"{name}_identical_on_all_ranks"])
        for key in (f"{name}_ef", f"{name}_ef_nccl"):
            ef = r[key]
            assert np.abs(ef[:-1] - ref[:-1]).max() <= 5e-5 + 2e-5 * np.abs(ref[:-1]).max(), key
            assert abs(ef[-1] - ref[-1]) <= 4e-3 * len(fd), key
        assert np.abs(r[f"{name}_ef"] - r[f"{name}_ef_nccl"]).max() <= 2e-5 * max(1.0, np.abs(ref[:-1]).max())
    # the sharded MD trajectory equals the single-GPU one up to fp32 force rounding amplified by 20 steps
    assert bool(r["md_one_graph"]) and int(r["md_step"]) == 20
    fd, pm = load_fragments("chig")
    prot_pos, prot_z, recipe = load_protein("chig")
    one = DeviceLangevin(real_weights, fd, pm, recipe, prot_pos, prot_z, dt_fs=1.0, temperature_K=300.0, friction_per_fs=0.001,
                         seed=0, device=0)
    one.run(20)
    x1, _, _, _ = one.state()
    assert np.abs(r["md_x"] - x1).max() <= 2e-5
