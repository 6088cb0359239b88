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


def test_pipeline_incremental_from_changesets_equals_rebuild(eng):
    """The three stages on their changeset legs, as the pipeline runs a short block range (hashing_account.rs:240-262,
    hashing_storage.rs:180-206, merkle.rs:255-300 with load_prefix_sets_with_provider): the range's AccountChangeSets /
    StorageChangeSets rows -> b200_hash_changesets -> hashed tables updated, prefix sets, incremental root over the trie
    tables (b200_root_from_items) — hashed tables, root and trie tables equal a from-scratch run of the full-pass stages over
    the same plain state.  Nothing is resident on the device between the ranges."""
    from reth_b200 import Account, AccountHashingStage, MerkleStage, StorageHashingStage
    from reth_b200.stages import Tables
    rng = np.random.default_rng(61)
    addr = lambda: bytes(rng.integers(0, 256, 20, dtype=np.uint8))
    t = Tables()
    for _ in range(700):
        a = addr()
        t.plain_accounts[a] = Account(int(rng.integers(0, 9)), int(rng.integers(1, 2**60)), None)
        if rng.random() < 0.3:
            t.plain_storage[a] = {int(rng.integers(0, 2**40)): int(rng.integers(1, 2**62)) for _ in range(int(rng.integers(1, 30)))}
    AccountHashingStage(eng).execute(t)
    StorageHashingStage(eng).execute(t)
    ms = MerkleStage(eng)
    ms.execute(t)
    rows = lambda tu: {k: v.storage_nodes for k, v in tu.storage_tries.items() if v.storage_nodes}
    for _range in range(3):
        acct_cs, stor_cs = [], []
        for _block in range(4):
            live = sorted(t.plain_accounts)
            picks = sorted({live[int(i)] for i in rng.choice(len(live), 25, replace=False)} | {addr() for _ in range(5)})
            for a in picks:                                   # changeset order inside a block: by address
                r = rng.random()
                if a not in t.plain_accounts:                 # created (with storage, sometimes)
                    acct_cs.append(a)
                    t.plain_accounts[a] = Account(0, int(rng.integers(1, 2**60)), None)
                    if rng.random() < 0.5:
                        t.plain_storage[a] = {}
                        for _ in range(4):
                            sl = int(rng.integers(0, 2**40))
                            stor_cs.append((a, sl))
                            t.plain_storage[a][sl] = int(rng.integers(1, 2**62))
                elif r < 0.35:                                # balance / nonce change
                    acct_cs.append(a)
                    old = t.plain_accounts[a]
                    t.plain_accounts[a] = Account(old.nonce + 1, int(rng.integers(1, 2**60)), None)
                elif r < 0.5:                                 # destroyed: the changesets list the account and every slot it had
                    acct_cs.append(a)
                    for sl in sorted(t.plain_storage.get(a, {})):
                        stor_cs.append((a, sl))
                    t.plain_accounts.pop(a)
                    t.plain_storage.pop(a, None)
                else:                                         # storage writes: new slots, changed slots, cleared slots
                    st = t.plain_storage.setdefault(a, {})
                    for sl in list(st)[:3]:
                        stor_cs.append((a, sl))
                        if rng.random() < 0.5:
                            st.pop(sl)
                        else:
                            st[sl] = int(rng.integers(1, 2**62))
                    for _ in range(3):
                        sl = int(rng.integers(0, 2**40))
                        stor_cs.append((a, sl))
                        st[sl] = int(rng.integers(1, 2**62))
        assert len(acct_cs) > len(set(acct_cs)) or len(stor_cs) > len(set(stor_cs)) or _range    # repeats across blocks do occur
        AccountHashingStage(eng).execute_incremental(t, acct_cs)
        StorageHashingStage(eng).execute_incremental(t, stor_cs)
        root = ms.execute_incremental_from_changesets(t, acct_cs, stor_cs)
        ref = Tables(plain_accounts=dict(t.plain_accounts), plain_storage={a: dict(s) for a, s in t.plain_storage.items()})
        AccountHashingStage(eng).execute(ref)
        StorageHashingStage(eng).execute(ref)
        ref_root = MerkleStage(eng).execute(ref)
        assert t.hashed_accounts == ref.hashed_accounts
        assert t.hashed_storages == ref.hashed_storages
        assert root == ref_root
        assert t.trie_updates.account_nodes == ref.trie_updates.account_nodes
        assert rows(t.trie_updates) == rows(ref.trie_updates)
    with pytest.raises(Exception):
        ms.execute_incremental_from_changesets(t, acct_cs[:1], [], expected_state_root=bytes(32))


def test_from_reverts_mirror(eng):
    """crates/trie/db/src/state.rs:439-517 `from_reverts_keeps_first_occurrence_and_ordering` and :519-541
    `from_reverts_empty_range`, through HashedPostStateSorted.from_reverts (one b200_hash_changesets call)."""
    from reth_b200 import Account, HashedPostStateSorted
    a1, a2 = bytes(19) + b"\x01", bytes(19) + b"\x02"
    acct = [(a1, Account(1, 0, None)), (a1, Account(2, 0, None)), (a2, None)]        # blocks 1, 2, 3
    stor = [(a1, 22, 200), (a1, 11, 100), (a1, 11, 999)]                              # the last row must be ignored
    st = HashedPostStateSorted.from_reverts(eng, acct, stor)
    h1, h2 = oracle.keccak256(a1), oracle.keccak256(a2)
    assert len(st.accounts) == 2
    assert dict(st.accounts)[h1].nonce == 1 and dict(st.accounts)[h2] is None
    assert [k for k, _ in st.accounts] == sorted([h1, h2])
    slots = st.storages[h1].storage_slots
    assert slots == sorted([(oracle.keccak256((11).to_bytes(32, "big")), 100), (oracle.keccak256((22).to_bytes(32, "big")), 200)])
    assert st.storages[h1].wiped is False and set(st.storages) == {h1}
    empty = HashedPostStateSorted.from_reverts(eng, [], [])
    assert empty.accounts == [] and empty.storages == {}
    # random ranges against the host restatement above
    rng = np.random.default_rng(17)
    addrs = [bytes(rng.integers(0, 256, 20, dtype=np.uint8)) for _ in range(40)]
    acct = [(addrs[int(rng.integers(0, 40))], Account(int(i), 0, None) if rng.random() < 0.8 else None) for i in range(300)]
    stor = [(addrs[int(rng.integers(0, 40))], int(rng.integers(0, 25)), int(i) + 1) for i in range(900)]
    st = HashedPostStateSorted.from_reverts(eng, acct, stor)
    first_a, first_s = {}, {}
    for a, info in acct:
        first_a.setdefault(a, info)
    for a, sl, v in stor:
        first_s.setdefault((a, sl), v)
    assert st.accounts == sorted((oracle.keccak256(a), info) for a, info in first_a.items())
    want = {}
    for (a, sl), v in first_s.items():
        want.setdefault(oracle.keccak256(a), []).append((oracle.keccak256(sl.to_bytes(32, "big")), v))
    assert {k: v.storage_slots for k, v in st.storages.items()} == {k: sorted(v) for k, v in want.items()}
