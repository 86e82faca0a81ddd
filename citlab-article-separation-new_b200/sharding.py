"""Page-wise sharding of an image list over GPU ranks.

The reference's only parallelism on this path is page-level data parallelism:
``run_net_post_processing.py:59-82`` cuts the image list into sub-lists of
``min(50, len // num_processes)`` pages and submits them to a ``ProcessPoolExecutor``; every worker
builds its own post-processor (its own ``load_graph``).  Pages are independent, so the B200 version
keeps exactly that shape - one process, one engine and one stream set per GPU - and needs no
collective: the "gather" is the file system (each rank writes its own PAGE-XML / maps).

Two helpers:
  * ``reference_sublists`` - the reference's cut, restated (same sub-list boundaries);
  * ``shard_for_rank``     - which pages rank r of world N processes: pages sorted by descending
    pixel count and dealt round-robin (pages differ in size, the net cost is linear in pixels),
    or plain round-robin when sizes are unknown.  Deterministic on every rank, no communication.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

MAX_SUBLIST_SIZE = 50  # run_net_post_processing.py:59


def reference_sublists(image_paths: Sequence[str], num_processes: int) -> List[List[str]]:
    """run_net_post_processing.py:62-69."""
    paths = list(image_paths)
    if not paths:
        return []
    size = len(paths) // max(1, num_processes)
    if size == 0:
        size = 1
    size = min(MAX_SUBLIST_SIZE, size)
    return [paths[i:i + size] for i in range(0, len(paths), size)]


def shard_indices(n_pages: int, world: int, rank: int, pixel_counts: Optional[Sequence[int]] = None) -> List[int]:
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} of world {world}")
    if pixel_counts is not None and len(pixel_counts) != n_pages:
        raise ValueError("pixel_counts must have one entry per page")
    order = list(range(n_pages))
    if pixel_counts is not None:
        order.sort(key=lambda i: (-int(pixel_counts[i]), i))   # stable: ties keep list order
        # serpentine deal (0..N-1, N-1..0, ...) keeps the per-rank pixel sums within one page of each other
        mine = []
        for k, i in enumerate(order):
            lap, slot = divmod(k, world)
            owner = slot if lap % 2 == 0 else world - 1 - slot
            if owner == rank:
                mine.append(i)
        return mine
    return order[rank::world]


def shard_for_rank(image_paths: Sequence[str], world: int, rank: int,
                   pixel_counts: Optional[Sequence[int]] = None) -> List[str]:
    return [image_paths[i] for i in shard_indices(len(image_paths), world, rank, pixel_counts)]


def rank_from_env() -> Tuple[int, int, int]:
    """(rank, world, local_rank) as torchrun exports them; a plain run is rank 0 of 1."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def group_by_shape(shapes: Sequence[Tuple[int, int]]) -> List[List[int]]:
    """Indices grouped by (h, w) in first-seen order: pages of one shape share a plan and can be batched."""
    groups, index = [], {}
    for i, s in enumerate(shapes):
        k = (int(s[0]), int(s[1]))
        if k not in index:
            index[k] = len(groups)
            groups.append([])
        groups[index[k]].append(i)
    return groups
