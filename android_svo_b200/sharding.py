"""Multi-GPU partitioning of the batched workload: independent sequences are block-sharded
over the ranks of one node, there is NO data-path collective, and one all-gather at the end carries a
fixed 64-byte record per sequence (8 float64: sequence id, tracked, matched, seeds updated, seeds
converged, Gauss-Newton iterations, pose error rot / trans).  torch.distributed is plumbing only:
backend "nccl" on the GPU box (NVLink), "gloo" in the CPU tests.
"""
import numpy as np

RECORD_FIELDS = ("seq_id", "n_tracked", "n_matched", "n_seeds_updated", "n_seeds_converged", "align_iters", "pose_err_rot", "pose_err_trans")
RECORD_BYTES = 8 * len(RECORD_FIELDS)


def shard(total_sequences, rank, world):
    """Contiguous block partition: rank r owns [r*per, (r+1)*per); the remainder (total % world) is dropped so
    that every rank carries the same load.  Returns (usable_total, per_rank, range)."""
    usable = total_sequences - total_sequences % world
    per = usable // world
    return usable, per, range(rank * per, (rank + 1) * per)


def make_records(seq_ids, stats, pose_err):
    """stats: structured array with the svob200_step_stats fields; pose_err: [n,2] (rot, trans)."""
    rec = np.zeros((len(seq_ids), len(RECORD_FIELDS)), np.float64)
    rec[:, 0] = np.asarray(seq_ids)
    for k, name in enumerate(RECORD_FIELDS[1:6], start=1):
        rec[:, k] = stats[name]
    rec[:, 6:8] = pose_err
    return rec


def gather_records(rec, world, device=None):
    """all_gather_into_tensor of the per-sequence records (rank order == sequence order for a block partition)."""
    if world == 1:
        return rec
    import torch
    import torch.distributed as dist
    mine = torch.from_numpy(np.ascontiguousarray(rec))
    if device is not None:
        mine = mine.to(device)
    out = torch.empty((world * rec.shape[0], rec.shape[1]), dtype=torch.float64, device=mine.device)
    dist.all_gather_into_tensor(out, mine)
    return out.cpu().numpy()


def summarize(rec):
    return {"n_sequences": int(len(rec)), "tracked_mean": float(rec[:, 1].mean()), "matched_mean": float(rec[:, 2].mean()),
            "seeds_updated_mean": float(rec[:, 3].mean()), "seeds_converged_mean": float(rec[:, 4].mean()),
            "align_iters_mean": float(rec[:, 5].mean()), "pose_err_rot_max": float(rec[:, 6].max()),
            "pose_err_trans_max": float(rec[:, 7].max())}


def max_over_ranks(x, world, device=None):
    """a timing is the MAX over ranks (never the wall clock of one rank)"""
    if world == 1:
        return float(x)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
