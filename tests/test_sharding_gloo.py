"""N > 1 path on CPU (gloo, world_size 2): streams are partitioned across ranks without overlap, nothing crosses ranks on
the data path, and the whole-job measurement is frames summed over ranks divided by the slowest rank's time."""
import os
import socket
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from libhm_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.assign_streams(8, world, rank)
    frames_local = 17 * len(mine)                       # configs[4]: 17 pictures per low-delay stream
    ms_local = 100.0 + 50.0 * rank                      # rank 1 is the slow one
    dist.barrier()
    frames, ms = sharding.reduce_measurement(frames_local, ms_local)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out.put((frames, ms, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_streams_partition_and_measurement_reduce_over_two_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    frames, ms, gathered = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(sum(gathered, [])) == list(range(8))              # every stream exactly once
    assert not set(gathered[0]) & set(gathered[1])
    assert frames == 8 * 17 and ms == 150.0                         # sum of frames, max of times
    assert abs(sharding.frames_per_second(frames, ms) - 8 * 17 / 0.150) < 1e-6


def test_assignment_shapes():
    for world in (1, 2, 4, 8):
        parts = [sharding.assign_streams(8, world, r) for r in range(world)]
        assert all(len(p) == 8 // world for p in parts)
    assert sharding.host_cores_of_rank(16, 2, 1) == list(range(8, 16))
    assert sharding.host_cores_of_rank(16, 1, 0) == list(range(16))
