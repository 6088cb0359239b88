"""CPU checks of bench.py's synthetic-data generators (they run on the GPU in the benchmark; torch CPU tensors here):
the splitmix64 stream equals the numpy one used by the parity tests, and the C3 / C4 shards satisfy the C-ABI's input
contract (keys strictly ascending inside every trie, non-zero values, monotone offsets, shard nibble range) — checked
by running the oracle over them."""
import numpy as np
import torch

import oracle
from bench import be_sort_key, effective_cpus, make_c3_shard, make_c4_shard, random_keys_torch, splitmix64_torch
from tests.util import random_keys, splitmix64_stream

CPU = torch.device("cpu")


def test_splitmix64_matches_numpy_reference():
    for seed in (2, 3, 12345):
        t = splitmix64_torch(seed, 1000, CPU).numpy().view(np.uint64)
        assert (t == splitmix64_stream(seed, 1000)).all()
    assert (random_keys_torch(2, 64, CPU).numpy().view(np.uint8).reshape(64, 32) == random_keys(2, 64)).all()


def test_be_sort_key_orders_like_bytes():
    k = random_keys_torch(7, 5000, CPU)
    order = torch.sort(be_sort_key(k), stable=True).indices.numpy()
    b = k.numpy().view(np.uint8).reshape(-1, 32)[order]
    assert all(b[i, :8].tobytes() <= b[i + 1, :8].tobytes() for i in range(len(b) - 1))


def _check_shard(sh, lo, hi):
    n, m = sh["n_accounts"], sh["n_slots"]
    akeys = sh["akeys"].numpy().reshape(n, 32)
    accts = sh["accts"].numpy().view(oracle.ACCOUNT_DTYPE).reshape(-1)
    skeys = sh["skeys"].numpy().reshape(m, 32)
    svals = sh["svals"].numpy().reshape(m, 32)
    offs = sh["offs"].numpy().astype(np.uint64)
    assert len(accts) == n and offs[0] == 0 and offs[-1] == m and (np.diff(offs.astype(np.int64)) >= 0).all()
    top = akeys[:, 0] >> 4
    assert top.min() >= lo and top.max() < hi
    assert svals.any(axis=1).all()                                   # no zero-valued slot
    # the oracle rejects unsorted / duplicate keys, so a root coming back means the ordering contract holds
    root = oracle.state_root_full(akeys, accts, skeys, svals, offs, threads=4)
    assert len(root) == 32
    return offs


def test_c3_shard_contract():
    sh = make_c3_shard(3, 3000, 16, 4, 8, CPU)
    offs = _check_shard(sh, 4, 8)
    assert (np.diff(offs.astype(np.int64)) == 16).all()


def test_c4_shard_contract_and_shape():
    sh = make_c4_shard(4, 60_000, 0, 16, CPU)
    offs = _check_shard(sh, 0, 16)
    counts = np.diff(offs.astype(np.int64))
    assert (counts == 0).mean() > 0.75                               # ~80 % EOAs
    assert counts.max() == sh["max_trie"] and counts.max() > 5_000   # one dominant contract
    assert sh["n_accounts"] + sh["n_slots"] == 60_000


def test_effective_cpus_is_sane():
    n = effective_cpus()
    assert 1 <= n <= 4096
