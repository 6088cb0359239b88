"""b200_comm_*: the exchange steps behind the C ABI on 2 GPUs of one box (one thread per GPU, each with its own context; the
128-byte NCCL id travels through a Python variable).  Skipped on boxes with a single GPU.
  * b200_state_root_sharded: shard by top key nibble -> frontier -> ncclAllGather -> root == the unsharded oracle root;
  * b200_hash_partition_dev: every rank hashes an arbitrary slice of the plain table, the all-to-all delivers each rank its
    buckets sorted by digest with the rows attached == keccak + sort of the whole table, cut at the bucket boundary."""
import threading

import numpy as np
import pytest

import oracle
from tests.util import sort_rows, synth_accounts, synth_storage

pytestmark = [pytest.mark.gpu]


def _two_gpus():
    from reth_b200 import _lib
    return _lib.load().b200_device_count() >= 2


def _run_ranks(world, fn):
    from reth_b200 import Comm, Engine
    uid = Comm.unique_id()
    out, errs = [None] * world, []

    def rank_main(r):
        try:
            eng = Engine(r)
            comm = Comm(eng, uid, world, r)
            out[r] = fn(r, eng, comm)
            comm.close()
            eng.close()
        except Exception as e:  # noqa: BLE001
            errs.append((r, repr(e)))

    ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not errs, errs
    return out


@pytest.mark.skipif(not _two_gpus(), reason="needs 2 GPUs")
def test_state_root_sharded_two_ranks():
    n = 20_000
    akeys, accs = synth_accounts(31, n)
    skeys, svals, offs = synth_storage(32, (np.arange(n) % 4 == 0) * 6)
    want = oracle.state_root_full(akeys, accs, skeys, svals, offs, threads=4)
    top = akeys[:, 0] >> 4

    def shard(r, eng, comm):
        sel = np.nonzero((top * 2 // 16) == r)[0]
        a0, a1 = int(sel[0]), int(sel[-1]) + 1                       # keys are sorted: a rank's accounts are contiguous
        s0, s1 = int(offs[a0]), int(offs[a1])
        return comm.state_root_sharded(akeys[a0:a1], accs[a0:a1], skeys[s0:s1], svals[s0:s1], (offs[a0:a1 + 1] - offs[a0]).astype(np.uint64))

    roots = _run_ranks(2, shard)
    assert roots[0] == want and roots[1] == want


@pytest.mark.skipif(not _two_gpus(), reason="needs 2 GPUs")
def test_hash_partition_two_ranks():
    import torch  # (before the communicator: the library then shares torch's NCCL instead of loading the system copy first)
    n = 50_000
    rng = np.random.default_rng(5)
    addrs = rng.integers(0, 256, (n, 20), dtype=np.uint8)
    rows = rng.integers(0, 256, (n, 72), dtype=np.uint8)              # the plain account rows riding along
    dig = oracle.keccak256_fixed(addrs)
    order = sort_rows(dig)
    owner = (dig[:, 0] >> 4).astype(np.int64) * 2 // 16

    def part(r, eng, comm):
        mine = slice(r * n // 2, (r + 1) * n // 2)                    # an arbitrary slice of the plain table
        dev = torch.device("cuda", r)
        with torch.cuda.device(dev):
            t_in = torch.from_numpy(addrs[mine].copy()).to(dev).view(-1)
            t_val = torch.from_numpy(rows[mine].copy()).to(dev).view(-1)
            cap = n
            t_k = torch.empty(cap * 32, dtype=torch.uint8, device=dev)
            t_v = torch.empty(cap * 72, dtype=torch.uint8, device=dev)
            got = comm.hash_partition_dev(t_in, 20, 20, n // 2, t_val, 72, cap, t_k, t_v)
            return t_k.view(cap, 32)[:got].cpu().numpy(), t_v.view(cap, 72)[:got].cpu().numpy()

    res = _run_ranks(2, part)
    for r in range(2):
        idx = order[owner[order] == r]                                # the whole table hashed + sorted, this rank's buckets
        assert (res[r][0] == dig[idx]).all()
        assert (res[r][1] == rows[idx]).all()


@pytest.mark.skipif(not _two_gpus(), reason="needs 2 GPUs")
def test_sharded_dynamic_state_two_ranks():
    """The live path at N = 2: every rank keeps its top-nibble buckets resident (b200_dstate_create_sharded), applies its part
    of each block in place and b200_dstate_root_sharded gathers the frontiers (NCCL inside the library): the root equals the
    oracle's from-scratch StateRoot over the merged state after every block."""
    from reth_b200 import Account, HashedPostState, HashedStorage, ShardedDynamicStateRoot
    rng = np.random.default_rng(44)
    rk = lambda: bytes(rng.integers(0, 256, 32, dtype=np.uint8))
    base = HashedPostState()
    for _ in range(1500):
        k = rk()
        base.accounts[k] = Account(int(rng.integers(0, 9)), int(rng.integers(0, 2**60)), None)
        if rng.random() < 0.3:
            base.storages[k] = HashedStorage(False, {rk(): int(rng.integers(1, 2**62)) for _ in range(int(rng.integers(1, 12)))})
    blocks = []
    live = sorted(base.accounts)
    for step in range(3):
        b = HashedPostState()
        for k in [live[int(i)] for i in rng.choice(len(live), 40, replace=False)]:
            r = rng.random()
            if r < 0.5:
                b.accounts[k] = Account(step + 10, int(rng.integers(0, 2**60)), None)
            elif r < 0.7:
                b.accounts[k] = None
                b.storages[k] = HashedStorage(True, {})
            else:
                b.storages[k] = HashedStorage(False, {rk(): int(rng.integers(1, 2**62)) for _ in range(4)})
        for _ in range(10):
            b.accounts[rk()] = Account(0, 7, None)
        blocks.append(b)
    # expected roots: from-scratch oracle StateRoot over the merged state after every block
    merged = HashedPostState(dict(base.accounts), {k: HashedStorage(False, dict(s.storage)) for k, s in base.storages.items()})
    want = []
    for b in blocks:
        merged.extend(b)
        post = HashedPostState({k: a for k, a in merged.accounts.items() if a is not None},
                               {k: HashedStorage(False, {s: v for s, v in st.storage.items() if v != 0})
                                for k, st in merged.storages.items() if merged.accounts.get(k) is not None})
        want.append(oracle.state_root_full(*post.into_sorted().to_flat(), threads=4))

    def rank_main(r, eng, comm):
        sh = ShardedDynamicStateRoot(eng, base, r, 2, comm=comm)
        roots = [sh.commit(b)[0] for b in blocks]
        sh.close()
        return roots

    res = _run_ranks(2, rank_main)
    assert res[0] == want and res[1] == want
