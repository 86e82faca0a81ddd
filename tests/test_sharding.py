"""Page sharding: the reference's sub-list cut restated, the per-rank shards, and the
N > 1 path under torch.distributed (gloo, world_size 2, CPU) - shards are disjoint and cover the list with no
data-path collective; the only collective is bench.py's max-over-ranks of the timing."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_sublists_match_the_reference_cut():
    from aru_b200.sharding import reference_sublists
    paths = [f"p{i}.jpg" for i in range(230)]
    subs = reference_sublists(paths, 8)                   # 230 // 8 = 28 pages per sub-list (run_net_post_processing.py:62-69)
    assert [len(s) for s in subs] == [28] * 8 + [6] and sum(subs, []) == paths
    assert [len(s) for s in reference_sublists(paths, 2)] == [50, 50, 50, 50, 30]   # capped at MAX_SUBLIST_SIZE
    assert reference_sublists(paths[:3], 8) == [["p0.jpg"], ["p1.jpg"], ["p2.jpg"]]
    assert reference_sublists([], 8) == []


@pytest.mark.parametrize("n,world", [(0, 1), (1, 4), (7, 2), (64, 8), (1024, 8), (13, 13)])
def test_shards_partition_the_list(n, world):
    from aru_b200.sharding import shard_indices
    import numpy as np
    sizes = list(np.random.default_rng(n).integers(1, 1000, size=n))
    for counts in (None, sizes):
        shards = [shard_indices(n, world, r, counts) for r in range(world)]
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(n))
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    if n >= world > 1:
        loads = [sum(sizes[i] for i in shard_indices(n, world, r, sizes)) for r in range(world)]
        assert max(loads) - min(loads) <= max(sizes)      # serpentine deal: within one page of each other


def test_shard_arguments_are_validated():
    from aru_b200.sharding import shard_indices
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)
    with pytest.raises(ValueError):
        shard_indices(4, 2, 0, [1, 2, 3])


def test_group_by_shape():
    from aru_b200.sharding import group_by_shape
    assert group_by_shape([(10, 8), (12, 8), (10, 8), (12, 8), (5, 5)]) == [[0, 2], [1, 3], [4]]


_WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from aru_b200.sharding import rank_from_env, shard_for_rank
rank, world, local = rank_from_env()
dist.init_process_group("gloo", rank=rank, world_size=world)
paths = [f"page_{{i:03d}}.png" for i in range(37)]
sizes = [(i * 7919) % 1000 + 1 for i in range(37)]
mine = shard_for_rank(paths, world, rank, sizes)
# the data path needs no collective; gathering here is only the test's way to look at all shards
allshards = [None] * world
dist.all_gather_object(allshards, mine)
# bench.py's timing rule: max over ranks
t = torch.tensor([float(10 + rank)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
dist.barrier()
if rank == 0:
    print(json.dumps({{"shards": allshards, "tmax": float(t.item())}}))
dist.destroy_process_group()
"""


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    a, b = res["shards"]
    assert not set(a) & set(b) and sorted(a + b) == [f"page_{i:03d}.png" for i in range(37)]
    assert abs(len(a) - len(b)) <= 1 and res["tmax"] == 11.0


def test_cli_parser_matches_the_reference_surface():
    from aru_b200.run_net_post_processing import build_parser, default_fixed_height
    args = build_parser().parse_args(["--path_to_image_list", "l", "--path_to_pb", "p", "--mode", "separator"])
    assert args.num_processes == 8 and args.scaling_factor == 1.0 and args.threshold == 0.05 and args.fixed_height is None
    assert default_fixed_height("separator") == 1500 and default_fixed_height("heading") == 900
    with pytest.raises(SystemExit):
        build_parser().parse_args(["--path_to_image_list", "l", "--path_to_pb", "p", "--mode", "textblock"])
