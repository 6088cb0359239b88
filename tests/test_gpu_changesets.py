"""b200_hash_changesets: a block range's account / storage changesets -> the hashed dirty set, against a host
restatement of HashedPostStateSorted::from_reverts (crates/trie/db/src/state.rs:289-347: first occurrence of every address
and (address, slot) pair wins, keys keccak-hashed and sorted — its test `from_reverts_keeps_first_occurrence`, state.rs:442)
and of load_prefix_sets_with_provider (crates/trie/db/src/prefix_set.rs:22-60)."""
import numpy as np
import pytest

import oracle

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def eng():
    from reth_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def restate(acct, sa, ss):
    """from_reverts + load_prefix_sets on the host (dicts keep the first occurrence)."""
    first_a = {}
    for i, a in enumerate(acct):
        first_a.setdefault(a.tobytes(), i)
    first_s = {}
    for i in range(len(sa)):
        first_s.setdefault((sa[i].tobytes(), ss[i].tobytes()), i)
    k = lambda b: oracle.keccak256(b)
    accounts = sorted((k(a), i) for a, i in first_a.items())
    storages = {}
    for (a, s), i in first_s.items():
        storages.setdefault(k(a), []).append((k(s), i))
    for v in storages.values():
        v.sort()
    prefix = sorted({h for h, _ in accounts} | set(storages))
    return accounts, storages, prefix


def check(eng, acct, sa, ss):
    got = eng.hash_changesets(acct, sa, ss)
    accounts, storages, prefix = restate(acct, sa, ss)
    assert [bytes(x) for x in got["account_keys"]] == [h for h, _ in accounts]
    assert list(got["account_first"]) == [i for _, i in accounts]
    assert [bytes(x) for x in got["storage_account_keys"]] == sorted(storages)
    offs = got["storage_seg_offsets"]
    assert int(offs[0]) == 0 and int(offs[-1]) == len(got["slot_keys"])
    for j, h in enumerate(sorted(storages)):
        seg = slice(int(offs[j]), int(offs[j + 1]))
        assert [bytes(x) for x in got["slot_keys"][seg]] == [s for s, _ in storages[h]]
        assert list(got["slot_first"][seg]) == [i for _, i in storages[h]]
    assert [bytes(x) for x in got["account_prefix_keys"]] == prefix


def test_keeps_first_occurrence(eng):
    """state.rs:442 from_reverts_keeps_first_occurrence: the same address / slot changed in several blocks."""
    a1, a2 = np.full(20, 1, np.uint8), np.full(20, 2, np.uint8)
    acct = np.stack([a1, a2, a1, a1, a2])                     # blocks 1,1,2,3,3
    sa = np.stack([a1, a1, a2, a1, a1])
    s1, s2 = np.zeros(32, np.uint8), np.zeros(32, np.uint8)
    s1[31], s2[31] = 1, 2
    ss = np.stack([s1, s2, s1, s1, s2])
    got = eng.hash_changesets(acct, sa, ss)
    assert list(got["account_first"]) == sorted([0, 1], key=lambda i: oracle.keccak256(acct[i].tobytes()))
    check(eng, acct, sa, ss)


@pytest.mark.parametrize("seed,n_addr,n_acct,n_stor", [(1, 50, 400, 3000), (2, 3, 10, 40), (3, 2000, 5000, 60000), (4, 1, 1, 1)])
def test_random_ranges(eng, seed, n_addr, n_acct, n_stor):
    rng = np.random.default_rng(seed)
    pool = rng.integers(0, 256, (n_addr, 20), dtype=np.uint8)
    slots = rng.integers(0, 256, (max(4, n_addr // 2), 32), dtype=np.uint8)
    acct = pool[rng.integers(0, n_addr, n_acct)]
    # storage rows come in runs of one address (an account's slots inside one block)
    rows_a, rows_s = [], []
    while len(rows_a) < n_stor:
        a = pool[rng.integers(0, n_addr)]
        for _ in range(int(rng.integers(1, 12))):
            rows_a.append(a)
            rows_s.append(slots[rng.integers(0, len(slots))])
    check(eng, acct, np.stack(rows_a[:n_stor]), np.stack(rows_s[:n_stor]))


def test_empty_sides(eng):
    rng = np.random.default_rng(9)
    acct = rng.integers(0, 256, (30, 20), dtype=np.uint8)
    z20, z32 = np.zeros((0, 20), np.uint8), np.zeros((0, 32), np.uint8)
    check(eng, acct, z20, z32)
    check(eng, z20, acct[:10], rng.integers(0, 256, (10, 32), dtype=np.uint8))
    check(eng, z20, z20, z32)
