"""world_size-2 CPU test (gloo) of the sharded dynamic state's host path (reth_b200/sharded.py): every rank commits its
part of each block to its own shard, the 16-entry frontiers are all-gathered and both ranks must arrive at the oracle's
state root.  The shards run on tools/emu's CPU emulation of the CUDA sources (test-side redirection of the loader, as
`pytest --emu` does); on GPUs the same class runs over NCCL."""
import os
import socket
import subprocess
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tools", "emu")


def _blocks(seed):
    """deterministic initial state + blocks (identical on every rank)"""
    from reth_b200 import Account, HashedPostState, HashedStorage
    rng = np.random.default_rng(seed)
    rk = lambda: bytes(rng.integers(0, 256, 32, dtype=np.uint8))
    base = HashedPostState()
    for _ in range(300):
        k = rk()
        base.accounts[k] = Account(int(rng.integers(0, 9)), int(rng.integers(1, 2**60)))
        if rng.random() < 0.3:
            base.storages[k] = HashedStorage(False, {rk(): int(rng.integers(1, 2**60)) for _ in range(int(rng.integers(1, 12)))})
    blocks = []
    live = sorted(base.accounts)
    for b in range(3):
        post = HashedPostState()
        for i in rng.choice(len(live), 25, replace=False):
            post.accounts[live[i]] = Account(b + 10, int(rng.integers(1, 2**50)))
        for _ in range(6):
            k = rk()
            post.accounts[k] = Account(0, 1)
            post.storages[k] = HashedStorage(False, {rk(): 7})
        for i in rng.choice(len(live), 5, replace=False):
            if live[i] not in post.accounts:
                post.storages[live[i]] = HashedStorage(False, {rk(): int(rng.integers(1, 2**40))})
        victim = live[int(rng.integers(0, len(live)))]
        post.accounts[victim] = None
        post.storages[victim] = HashedStorage(True, {})
        blocks.append(post)
    return base, blocks


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), B200_EMU="1")
    from reth_b200 import _lib
    _lib.LIB_PATH = os.path.join(EMU, "build", "libb200trie_emu.so")   # test-side redirection only
    from reth_b200 import Engine, ShardedDynamicStateRoot
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base, blocks = _blocks(5)
    eng = Engine(0)
    sh = ShardedDynamicStateRoot(eng, base, rank, world)
    roots = [sh.root()]
    for post in blocks:
        roots.append(sh.commit(post)[0])
    sh.close()
    with open(os.path.join(out_dir, f"roots_{rank}.txt"), "w") as f:
        f.write("\n".join(r.hex() for r in roots))
    dist.destroy_process_group()


def test_two_rank_sharded_dynamic_state(tmp_path):
    subprocess.run(["make", "-j8", "-C", EMU], check=True, capture_output=True)
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = open(tmp_path / "roots_0.txt").read().split()
    r1 = open(tmp_path / "roots_1.txt").read().split()
    assert r0 == r1 and len(r0) == 4
    # the oracle over the merged state after every block
    import oracle
    from reth_b200 import HashedPostState, HashedStorage
    base, blocks = _blocks(5)
    merged = HashedPostState(dict(base.accounts), {k: HashedStorage(False, dict(v.storage)) for k, v in base.storages.items()})

    def oracle_root(st):
        keys, accts, skeys, svals, offs = st.into_sorted().to_flat()
        return oracle.state_root_full(keys, accts, skeys, svals, offs).hex()

    assert r0[0] == oracle_root(merged)
    for i, post in enumerate(blocks):
        for k, hs in post.storages.items():
            cur = {} if hs.wiped else dict(merged.storages.get(k, HashedStorage()).storage)
            for sk, v in hs.storage.items():
                if v == 0:
                    cur.pop(sk, None)
                else:
                    cur[sk] = v
            merged.storages[k] = HashedStorage(False, cur)
        for k, a in post.accounts.items():
            if a is None:
                merged.accounts.pop(k, None)
                merged.storages.pop(k, None)
            else:
                merged.accounts[k] = a
        assert r0[i + 1] == oracle_root(merged), i
