"""One process per GPU: the dynamic resident state sharded by top key nibble.

Rank r of `world` owns the accounts whose hashed address starts with a nibble in [16r/world, 16(r+1)/world) together with
their storage tries.  A block's HashedPostState is filtered to the rank's buckets and committed to the local
`b200_dstate` shard; the ranks then all-gather their 16 frontier entries (16 x 68 bytes, torch.distributed: NCCL on
GPUs, gloo in the CPU tests) and every rank finishes the state root with b200_root_from_frontier.  No other data moves
between ranks: this is the collective-free partitioning of the reference's ParallelStateRoot fan-out
(crates/trie/parallel/src/root.rs:101-127) carried over to the live path."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .engine import Engine
from .hashed_state import HashedPostState, HashedPostStateSorted
from .trie import DynamicStateRoot, TrieUpdates


def owner_of(key: bytes, world: int) -> int:
    return (key[0] >> 4) * world // 16


class ShardedDynamicStateRoot:
    def __init__(self, engine: Engine, state: HashedPostState, rank: int, world: int, group=None, comm=None):
        """state: the full initial hashed state (every rank may pass the same object; only its buckets are kept) or
        already just this rank's part.  comm: a reth_b200.Comm — the frontier exchange then runs inside the library
        (b200_dstate_root_sharded, NCCL) instead of through torch.distributed."""
        if 16 % world and world > 16:
            raise ValueError("at most 16 ranks (one top-nibble bucket each)")
        self.engine, self.rank, self.world, self.group, self.comm = engine, rank, world, group, comm
        mine = HashedPostState({k: a for k, a in state.accounts.items() if owner_of(k, world) == rank},
                               {k: s for k, s in state.storages.items() if owner_of(k, world) == rank})
        self.local = DynamicStateRoot(engine, mine.into_sorted(), sharded=True)
        self._root = self._gather_root()

    def root(self) -> bytes:
        return self._root

    def _gather_root(self) -> bytes:
        if self.comm is not None:
            return self.comm.dstate_root_sharded(self.local.ds)
        import torch
        import torch.distributed as dist
        fr = self.local.ds.frontier()                                  # (16, 68) uint8, empty outside this rank's buckets
        if self.world == 1:
            return self.engine.root_from_frontier(fr)
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        mine = torch.from_numpy(fr.reshape(-1)).to(dev)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=self.group)
        allf = torch.stack(gathered).view(self.world, 16, 68).cpu().numpy()
        pick = np.arange(16) * self.world // 16                       # owner of every bucket (same map as bench.py)
        return self.engine.root_from_frontier(np.ascontiguousarray(allf[pick, np.arange(16)]))

    def commit(self, post: HashedPostState) -> Tuple[bytes, TrieUpdates]:
        """-> (state root, this rank's part of the block's TrieUpdates)."""
        part = HashedPostState({k: a for k, a in post.accounts.items() if owner_of(k, self.world) == self.rank},
                               {k: s for k, s in post.storages.items() if owner_of(k, self.world) == self.rank})
        _, updates = self.local.commit(part)
        self._root = self._gather_root()
        return self._root, updates

    def close(self):
        self.local.close()


def list_range_of(rank: int, world: int, n_lists: int) -> Tuple[int, int]:
    """Contiguous share of a batch of lists (blocks) for a rank."""
    return rank * n_lists // world, (rank + 1) * n_lists // world


def sharded_ordered_trie_roots(engine: Engine, lists, rank: int, world: int, group=None):
    """Ordered (transactions / receipts / withdrawals) roots of a batch of lists over `world` ranks: the lists are
    independent tries, so rank r folds lists [r·n/world, (r+1)·n/world) with b200_ordered_roots — no data-path
    collective — and only the 32-byte roots are all-gathered.  `lists` may be the full batch on every rank (only the
    rank's share is read).  -> all roots, in list order, on every rank."""
    from .ordered_root import ordered_trie_roots
    n = len(lists)
    lo, hi = list_range_of(rank, world, n)
    mine = ordered_trie_roots(engine, lists[lo:hi])
    if world == 1:
        return mine
    import torch
    import torch.distributed as dist
    width = (n + world - 1) // world                                   # every rank contributes a fixed-size block
    buf = np.zeros((width, 32), np.uint8)
    if mine:
        buf[:len(mine)] = np.frombuffer(b"".join(mine), np.uint8).reshape(-1, 32)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.from_numpy(buf.reshape(-1)).to(dev)
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t, group=group)
    out = []
    for r in range(world):
        r_lo, r_hi = list_range_of(r, world, n)
        rows = gathered[r].cpu().numpy().reshape(width, 32)
        out += [rows[i].tobytes() for i in range(r_hi - r_lo)]
    return out
