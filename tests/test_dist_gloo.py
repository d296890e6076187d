"""CPU: the N > 1 path -- sharding by global frame index + one counter all-reduce --
with world_size 2 and 3 over gloo.  The per-shard Monte-Carlo points are computed by
the CPU oracle here (same Philox frames the GPUs generate); on the GPU box the same
two functions wrap ldpc_experiment_run (bench.py, tests/test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from tests.helpers import load_rows

SEED = 239239239


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_path):
    import torch.distributed as dist
    from oracle.oracle import Oracle, dense_to_csr
    import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H = load_rows("optimalH")
    orc = Oracle()
    begin, end = sharding.shard_range(total, rank, world)
    mine = orc.experiment("qpadmm", dense_to_csr(H), H.shape[0], H.shape[1], -2.5, 300, SEED, begin, end - begin,
                          alpha=1.2, mu=0.55, eps_stop=1e-5)
    summed = sharding.allreduce_counters(mine)
    if rank == 0:
        np.save(out_path, np.array([summed[k] for k in sorted(summed)], np.int64))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_every_frame_once():
    import sharding
    for total in (0, 1, 7, 1000, 10 ** 9 + 7):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_counters_equal_single_rank(oracle, world, tmp_path):
    from oracle.oracle import dense_to_csr
    total = 41
    H = load_rows("optimalH")
    whole = oracle.experiment("qpadmm", dense_to_csr(H), H.shape[0], H.shape[1], -2.5, 300, SEED, 0, total, alpha=1.2,
                              mu=0.55, eps_stop=1e-5)
    out = str(tmp_path / "sum.npy")
    mp.spawn(_worker, args=(world, _free_port(), total, out), nprocs=world, join=True)
    got = np.load(out)
    assert list(got) == [whole[k] for k in sorted(whole)]
    assert whole["total"] == total
