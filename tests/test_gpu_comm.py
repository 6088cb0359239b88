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
