"""Multi-GPU layout of the decoder: whole bitstreams are the unit of independence (DESIGN.md §e — a bitstream does not
shard: pictures depend on the DPB, CTUs on their neighbours, CABAC is serial).  Stream i goes to rank i mod world; there
is no data-path collective, torch.distributed is used only for the barrier and for reducing the measurements."""
import torch
import torch.distributed as dist


def assign_streams(n_streams, world, rank):
    """Indices of the bitstreams rank `rank` decodes (round robin, BASELINE.json configs[4]: 8 streams on 1/2/4/8 GPUs)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return [i for i in range(n_streams) if i % world == rank]


def host_cores_of_rank(n_cores, world, local_rank):
    """Contiguous block of host cores that feed the decoders of one GPU (parse threads are pinned next to their GPU)."""
    per = max(1, n_cores // max(1, world))
    first = (local_rank * per) % max(1, n_cores)
    return list(range(first, min(n_cores, first + per)))


def reduce_measurement(frames_local, ms_local, device="cpu"):
    """Whole-job numbers from per-rank ones: frames are summed, the timed region is the MAX over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(frames_local), float(ms_local)
    f = torch.tensor([float(frames_local)], dtype=torch.float64, device=device)
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    dist.all_reduce(f, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(round(f.item())), float(t.item())


def frames_per_second(frames_total, ms_max):
    return frames_total / (ms_max / 1000.0) if ms_max > 0 else 0.0
