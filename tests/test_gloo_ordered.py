"""world_size-2 CPU test (gloo) of the sharded ordered roots (reth_b200/sharded.py: sharded_ordered_trie_roots): the lists
of a batch are independent tries, each rank folds its contiguous share and the 32-byte roots are all-gathered; both
ranks must hold the oracle's roots in list order.  The shards run on tools/emu's CPU emulation of the CUDA sources
(test-side redirection of the loader, as `pytest --emu` does); on GPUs the same function runs over NCCL."""
import os
import socket
import subprocess
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tools", "emu")


def _lists(seed):
    rng = np.random.default_rng(seed)
    sizes = [0, 3, 130, 1, 0, 47, 200, 9, 128]          # 9 lists over 2 ranks: shares of 4 and 5
    return [[rng.integers(0, 256, int(rng.integers(1, 400)), dtype=np.uint8).tobytes() for _ in range(n)] for n in sizes]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), B200_EMU="1")
    from reth_b200 import _lib
    _lib.LIB_PATH = os.path.join(EMU, "build", "libb200trie_emu.so")   # test-side redirection only
    from reth_b200 import Engine, sharded_ordered_trie_roots
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = Engine(0)
    roots = sharded_ordered_trie_roots(eng, _lists(8), rank, world)
    eng.close()
    with open(os.path.join(out_dir, f"roots_{rank}.txt"), "w") as f:
        f.write("\n".join(r.hex() for r in roots))
    dist.destroy_process_group()


def test_two_rank_sharded_ordered_roots(tmp_path):
    subprocess.run(["make", "-j8", "-C", EMU], check=True, capture_output=True)
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = open(tmp_path / "roots_0.txt").read().split()
    r1 = open(tmp_path / "roots_1.txt").read().split()
    import oracle
    want = [r.tobytes().hex() for r in oracle.ordered_roots(*oracle.pack_lists(_lists(8)))]
    assert r0 == want and r1 == want
