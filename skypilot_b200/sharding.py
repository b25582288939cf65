"""Multi-GPU sharding of a batch of independent DAGs.

One process per GPU (torchrun); the catalog is replicated on every GPU, DAG
*i* belongs to rank `i % world`, and there is no data-path collective: every
rank optimises its own DAGs with its own device calls. `torch.distributed`
(NCCL on the GPUs, gloo in the CPU tests) is only used to take the maximum of
the per-rank times and to collect the plans on rank 0.
"""
from typing import Any, List, Optional, Sequence


def owner(index: int, world: int) -> int:
    """Rank that optimises DAG `index`."""
    return index % world


def shard(items: Sequence[Any], rank: int, world: int) -> List[Any]:
    """The items of `rank`: index i goes to rank i % world."""
    return list(items[rank::world])


def merge(shards: Sequence[Sequence[Any]], n_items: int) -> List[Any]:
    """Inverse of `shard` over all ranks: results back in submission order."""
    world = len(shards)
    out: List[Any] = [None] * n_items
    for rank, part in enumerate(shards):
        expected = len(range(rank, n_items, world))
        if len(part) != expected:
            raise ValueError(f'rank {rank} returned {len(part)} results, '
                             f'expected {expected}')
        out[rank::world] = list(part)
    return out


def max_over_ranks(value: float, dist: Optional[Any], device: str = 'cpu'
                  ) -> float:
    """Slowest rank's time (the number a multi-GPU measurement reports)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch  # pylint: disable=import-outside-toplevel
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_on_root(local: List[Any], n_items: int, dist: Optional[Any]
                  ) -> Optional[List[Any]]:
    """Collects every rank's per-DAG results on rank 0, in submission order
    (None on the other ranks). `local` must be picklable."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(local)
    world = dist.get_world_size()
    parts: List[Any] = [None] * world
    dist.all_gather_object(parts, list(local))
    if dist.get_rank() != 0:
        return None
    return merge(parts, n_items)


def optimize_shard(dags: Sequence[Any], rank: int, world: int, device: int,
                   **kwargs) -> List[Any]:
    """`Optimizer.optimize_batch` on this rank's share of `dags` (one GPU)."""
    from skypilot_b200 import optimizer  # pylint: disable=import-outside-toplevel
    mine = shard(dags, rank, world)
    if not mine:
        return []
    return optimizer.Optimizer.optimize_batch(mine, devices=[device], **kwargs)
