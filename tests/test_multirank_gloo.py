"""World-size-2 checks of the host-side multi-rank logic on the gloo backend (CPU).

The loss path shards by sample with no data-path collective; the only loss-side
exchange is the `ready` gate of hiera_triplet_loss.py:193-200 / rmi_hiera_triplet_loss.py:530-536,
which the CUDA path replaces by one MIN all-reduce of a device flag (ops._world_ready), and the
whole-job throughput aggregation of bench.py (max over ranks of the device time)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, counts, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from seghiero_b200 import ops
        out = []
        for case in counts:
            # status[0] = local "found >= 1 triplet class" flag, status[1] = label error flag
            status = torch.tensor([1 if case[rank] > 0 else 0, 0], dtype=torch.int32)
            ops._world_ready(status)
            # the reference's own formulation: all_gather the class counts, ready = all > 0
            mine = torch.tensor([case[rank]], dtype=torch.int64)
            gathered = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(gathered, mine)
            ref_ready = int(all(int(g.item()) > 0 for g in gathered))
            out.append((int(status[0]), ref_ready, int(status[1])))
        # bench.py aggregation: value = world * px / max_r(ms)
        ms = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ret[rank] = (out, float(ms))
    finally:
        dist.destroy_process_group()


def test_ready_gate_matches_reference_all_gather():
    world = 2
    counts = [(3, 0), (0, 0), (2, 5), (0, 7)]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), counts, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        out, ms = ret[rank]
        assert ms == 11.0
        for (ready, ref_ready, err), case in zip(out, counts):
            assert ready == ref_ready == int(all(c > 0 for c in case))
            assert err == 0          # only element 0 is reduced


def test_world_ready_is_noop_without_process_group():
    from seghiero_b200 import ops
    status = torch.tensor([1, 0], dtype=torch.int32)
    ops._world_ready(status)
    assert status.tolist() == [1, 0]
