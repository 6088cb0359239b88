"""world_size-2 CPU test (gloo) of the multi-GPU host logic: shard assignment by top key nibble, the all-gather of
16-entry frontiers and the owner-rank merge that bench.py performs before b200_root_from_frontier.  The device
kernels are not involved (no GPU here): every rank fabricates recognisable frontier entries for the nibbles it
owns, and the merged frontier must contain, for every nibble, exactly the owner's entry."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def owner_of_nibble(nibble: int, world: int) -> int:
    """Inverse of bench.py's shard map (rank r owns nibbles [16r/world, 16(r+1)/world))."""
    for r in range(world):
        if r * 16 // world <= nibble < (r + 1) * 16 // world:
            return r
    raise AssertionError


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = rank * 16 // world, (rank + 1) * 16 // world
    front = torch.zeros(16, 68, dtype=torch.uint8)
    for b in range(lo, hi):
        front[b, 0] = 33                      # as_child_len
        front[b, 1] = 0xA0
        front[b, 2:34] = 16 * rank + b        # recognisable payload
        front[b, 34] = 32                     # as_root_len
    gathered = [torch.zeros(16 * 68, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(gathered, front.view(-1))
    merged = torch.stack(gathered).view(world, 16, 68)
    pick = torch.arange(16) * world // 16     # same expression as bench.py
    out = merged[pick, torch.arange(16)]
    np.save(os.path.join(out_dir, f"merged_{rank}.npy"), out.numpy())
    dist.destroy_process_group()


def test_two_rank_frontier_merge(tmp_path):
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    m0 = np.load(tmp_path / "merged_0.npy")
    m1 = np.load(tmp_path / "merged_1.npy")
    assert (m0 == m1).all()                   # every rank finishes the root from the same frontier
    for b in range(16):
        r = owner_of_nibble(b, world)
        assert m0[b, 0] == 33 and m0[b, 34] == 32
        assert (m0[b, 2:34] == 16 * r + b).all()


def test_owner_map_matches_bench_expression():
    for world in (1, 2, 4, 8, 16):
        pick = [(b * world) // 16 for b in range(16)]
        assert pick == [owner_of_nibble(b, world) for b in range(16)]
