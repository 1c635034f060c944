"""Multi-GPU partitioning of the hot path: one process per GPU.

* Extraction shards over clips with NO data-path collective: rank r of W takes clips r, r+W, ...
  (clips are independent in all three reference loops: extract_vision_huggingface.py:104-107,
  extract_audio_huggingface.py:72, extract_text_huggingface.py:209) and writes its own .npy files.
* Fusion training is data-parallel: every rank holds a replica, computes the gradient of
  (sum of per-sample losses) / GLOBAL batch on its slice of the batch, one all-reduce(SUM) of the flat
  gradient buffer gives the reference's batch-mean gradient, then every rank applies the same Adam
  update.  At W=1 this is exactly the reference step; at W>1 it equals the reference step on the
  concatenated batch.
"""
from __future__ import annotations

import os


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_indices(n_items, rank, world):
    """Indices of the items rank `rank` owns (round-robin)."""
    assert 0 <= rank < world
    return list(range(rank, n_items, world))


def shard_list(items, rank, world):
    return [items[i] for i in shard_indices(len(items), rank, world)]


def device_index(gpu):
    """The GPU a process drives: under torchrun (one process per GPU) LOCAL_RANK, otherwise the script's --gpu."""
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))
    return gpu


def my_work(items, out_path, resume=None):
    """The extractor CLIs' work-list rule: sort for a rank-independent order, drop the items
    whose output file already exists (resume after a restart, as MER2024 extract_sun_videomae.py:344 does; env
    MER_RESUME=0 or resume=False recomputes everything), then this rank's round-robin share ``[rank::world]``.
    ``out_path(item)`` -> the .npy the item produces.  Each rank writes only its own files: no collective."""
    rank, world = env_rank_world()
    if resume is None:
        resume = os.environ.get("MER_RESUME", "1") != "0"
    todo = sorted(items)
    if resume:
        todo = [it for it in todo if not os.path.exists(out_path(it))]
    return shard_list(todo, rank, world), rank, world


def batch_slice(global_batch, rank, world):
    """Contiguous slice [lo, hi) of a global batch owned by `rank` (sizes differ by at most 1);
    the reference's sampler permutation is kept, each rank takes its contiguous part."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_grads_(flat_grads, world):
    """SUM all-reduce of the flat gradient buffer (NCCL on GPUs, gloo in the CPU tests)."""
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(flat_grads)
    return flat_grads
