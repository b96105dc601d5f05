"""reference lib/utils/distributed.py:31-50: the evaluation edge's only collective -- every rank gathers every rank's clip
predictions, labels and clip indices (tools/test_net.py:113) so that each can run the multi-view ensemble."""
import torch
import torch.distributed as dist


def all_gather(tensors):
    """[t0, t1, ...] (same shapes on every rank) -> [cat over ranks of t0, ...]; identity without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(tensors)
    world = dist.get_world_size()
    out = []
    for t in tensors:
        t = t.contiguous()
        buf = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
        if hasattr(dist, "all_gather_into_tensor") and t.is_cuda:
            dist.all_gather_into_tensor(buf, t)                    # one NCCL call into the concatenated buffer
        else:
            dist.all_gather(list(buf.chunk(world, dim=0)), t)
        out.append(buf)
    return out
