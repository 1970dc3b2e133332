"""Data-parallel plumbing of the forward/grounding path: one process per GPU, batch sharded across
ranks, no collective on the data path (samples are independent: SURVEY 8e).

Mirrors the three places the reference touches torch.distributed around `forward`:
  - per-rank batch = global batch / world        (reference pythia/utils/general.py:233-246)
  - loss / metric scalars reduced to rank 0       (pythia/utils/distributed_utils.py:91-110 `reduce_dict`)
  - prediction tensors all-gathered for reports   (distributed_utils.py:74-88 `gather_tensor`,
                                                   called from common/test_reporter.py:141-142)
torch.distributed is plumbing here: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def per_rank_batch(global_batch, world_size=None):
    """reference general.py:233-246 `get_batch_size`: the global batch must divide evenly."""
    w = world() if world_size is None else world_size
    if global_batch % w != 0:
        raise RuntimeError("Batch size {} must be divisible by number of GPUs {} used.".format(global_batch, w))
    return global_batch // w


def shard_sample_list(sample_list, rank_=None, world_size=None):
    """Rank r keeps samples [r*b, (r+1)*b) of every batched tensor / list field; scalars and strings
    are replicated.  Returns a new container of the same type (SampleList or dict)."""
    r = rank() if rank_ is None else rank_
    w = world() if world_size is None else world_size
    n = None
    for v in sample_list.values():
        if torch.is_tensor(v) and v.dim() > 0:
            n = v.shape[0]
            break
    if n is None:
        raise ValueError("no batched tensor field to shard")
    b = per_rank_batch(n, w)
    out = type(sample_list)()
    for k, v in sample_list.items():
        if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n:
            piece = v[r * b:(r + 1) * b]
        elif isinstance(v, (list, tuple)) and len(v) == n:
            piece = v[r * b:(r + 1) * b]
        else:
            piece = v
        if hasattr(out, "add_field"):
            out.add_field(k, piece)
        else:
            out[k] = piece
    return out


def reduce_dict(dictionary):
    """Stack the values, `dist.reduce` to rank 0 and divide by world there (distributed_utils.py:91-110);
    other ranks keep their partial values, exactly like the reference."""
    w = world()
    if w < 2 or len(dictionary) == 0:
        return dictionary
    with torch.no_grad():
        keys, values = zip(*sorted(dictionary.items()))
        values = torch.stack([v.reshape(()) if v.numel() == 1 else v for v in values], dim=0)
        dist.reduce(values, dst=0)
        if dist.get_rank() == 0:
            values = values / w
        return {k: v for k, v in zip(keys, values)}


def gather_tensor(tensor):
    """all_gather -> [world, ...] (distributed_utils.py:74-88); identity on one rank."""
    w = world()
    if w < 2:
        return tensor
    with torch.no_grad():
        pieces = [torch.zeros_like(tensor) for _ in range(w)]
        dist.all_gather(pieces, tensor.contiguous())
        return torch.stack(pieces, dim=0)


def gather_predictions(model_output, keys=("pos_scores", "ground_frame", "ground_box")):
    """What the reference's TestReporter gathers before formatting predictions, flattened back to the
    global batch order produced by `shard_sample_list`."""
    out = {}
    for k in keys:
        if k in model_output:
            g = gather_tensor(model_output[k])
            out[k] = g.reshape((-1,) + tuple(model_output[k].shape[1:])) if world() > 1 else g
    return out


def all_reduce_flat_(flat, mean=False):
    """Gradient all-reduce of one flat buffer (reference: DistributedDataParallel's bucketed all-reduce inside
    backward, pythia/trainers/base_trainer.py:134-137; DDP averages).  Sums in place and returns the factor that
    turns the sum into the mean (1 / world) so the optimizer kernel can fold it in; mean=True applies it here."""
    w = world()
    if w < 2:
        return 1.0
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if mean:
        flat.mul_(1.0 / w)
        return 1.0
    return 1.0 / w
