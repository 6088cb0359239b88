"""Incremental root with the trie NOT resident: b200_root_from_items folds the element stream of reth's
TrieWalker + TrieNodeIter — leaves and stored hashes of unchanged subtrees.

Two criteria, both the reference's own:
  * the fold equals alloy-trie's HashBuilder fed the same stream (add_leaf / add_branch; the oracle restates it), root and
    updated branch nodes — crates/trie/trie/src/trie.rs:247-309;
  * incremental == full (crates/trie/db/tests/trie.rs:680-717 incremental vs full root, fuzz_in_memory_nodes.rs): after a
    random block of inserts / updates / deletes the root over (stored nodes + prefix sets) equals the from-scratch root, and
    the trie tables after applying the TrieUpdates (removed paths deleted, updated nodes upserted) equal the from-scratch
    tables — for storage tries inside a whole state as well."""
import numpy as np
import pytest

import oracle
from tests.util import sort_rows

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def eng():
    from reth_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def slot_tables(eng, keys, vals):
    """from-scratch storage trie: root + stored nodes {path: BranchNodeCompact}"""
    from reth_b200 import BranchNodeCompact
    if len(keys) == 0:
        return oracle.EMPTY_ROOT_HASH, {}
    roots, recs = eng.storage_roots(keys, vals, np.array([0, len(keys)], np.uint64), want_updates=True)
    return roots[0].tobytes(), {bytes(r[1]): BranchNodeCompact(r[2], r[3], r[4], tuple(r[5])) for r in recs}


def rand_vals(rng, n):
    v = np.zeros((n, 32), np.uint8)
    for i in range(n):
        ln = int(rng.integers(1, 33))
        v[i, 32 - ln:] = rng.integers(0, 256, ln, dtype=np.uint8)
        v[i, 32 - ln] |= 1
    return v


def clustered_keys(rng, n, prefixes):
    ks = []
    for _ in range(n):
        if prefixes and rng.random() < 0.5:
            p = prefixes[int(rng.integers(0, len(prefixes)))]
            ks.append(p + rng.integers(0, 256, 32 - len(p), dtype=np.uint8).tobytes())
        else:
            ks.append(rng.integers(0, 256, 32, dtype=np.uint8).tobytes())
    return ks


@pytest.mark.parametrize("seed,n,clustered", [(1, 400, False), (2, 3000, False), (3, 600, True), (4, 40, True), (5, 3, False)])
def test_storage_trie_incremental_equals_full_and_hashbuilder(eng, seed, n, clustered):
    from reth_b200 import PrefixSetMut, walk
    from reth_b200.walker import _items_arrays, unpack
    rng = np.random.default_rng(seed)
    prefixes = [rng.integers(0, 256, int(rng.integers(1, 20)), dtype=np.uint8).tobytes() for _ in range(4)] if clustered else []
    state = dict(zip(clustered_keys(rng, n, prefixes), [v.tobytes() for v in rand_vals(rng, n)]))
    for step in range(5):
        ks = sorted(state)
        keys = np.frombuffer(b"".join(ks), np.uint8).reshape(-1, 32)
        vals = np.frombuffer(b"".join(state[k] for k in ks), np.uint8).reshape(-1, 32)
        root0, tables = slot_tables(eng, keys, vals)
        # ---- a block of changes
        changed = PrefixSetMut()
        m = max(1, len(ks) // 20)
        for k in [ks[int(i)] for i in rng.choice(len(ks), min(m, len(ks)), replace=False)]:
            if rng.random() < 0.5:
                del state[k]
            else:
                state[k] = rand_vals(rng, 1)[0].tobytes()
            changed.insert(unpack(k))
        for k in clustered_keys(rng, m, prefixes):
            state[k] = rand_vals(rng, 1)[0].tobytes()
            changed.insert(unpack(k))
        ks2 = sorted(state)
        keys2 = np.frombuffer(b"".join(ks2), np.uint8).reshape(-1, 32) if ks2 else np.zeros((0, 32), np.uint8)
        vals2 = np.frombuffer(b"".join(state[k] for k in ks2), np.uint8).reshape(-1, 32) if ks2 else np.zeros((0, 32), np.uint8)
        full_root, full_tables = slot_tables(eng, keys2, vals2)
        assert full_root == (oracle.storage_roots(keys2, vals2, np.array([0, len(ks2)], np.uint64))[0].tobytes() if ks2 else oracle.EMPTY_ROOT_HASH)
        # ---- the walk and the fold
        elements, removed = walk(tables, changed.freeze(), ks2)
        n_leaves = sum(e.is_leaf for e in elements)
        if len(ks) > 200 and not clustered:
            assert n_leaves < len(ks2) // 2                   # most of the trie enters through stored hashes
        k, nb, fl, v = _items_arrays(elements, 32)
        for i, e in enumerate(elements):
            if e.is_leaf:
                v[i] = vals2[e.leaf_index]
        roots, recs = eng.root_from_items(k, nb, fl, v, None, None, account=False, want_updates=True)
        assert roots[0].tobytes() == full_root
        # HashBuilder over the same stream
        hb = oracle.HashBuilder(retain_updates=True)
        for e in elements:
            if e.is_leaf:
                hb.add_leaf(e.path, oracle.encode_u256(int.from_bytes(vals2[e.leaf_index].tobytes(), "big")))
            else:
                hb.add_branch(e.path, e.hash, e.children_are_in_trie)
        assert hb.root() == full_root
        want = {p: (u["state_mask"], u["tree_mask"], u["hash_mask"], u["hashes"]) for p, u in hb.updates().items() if p != b""}
        got = {bytes(r[1]): (r[2], r[3], r[4], list(r[5])) for r in recs}
        assert got == want
        # tables after the block == tables of the from-scratch build
        from reth_b200 import BranchNodeCompact
        new_tables = {p: nd for p, nd in tables.items() if p not in removed}
        new_tables.update({p: BranchNodeCompact(a, b, c, tuple(h)) for p, (a, b, c, h) in got.items()})
        assert new_tables == full_tables


def test_untouched_trie_is_one_lookup_per_top_node(eng):
    """Empty prefix set: nothing is re-hashed below the stored nodes; the fold sees only hashes (and the leaves no stored
    node covers) and reproduces the root."""
    from reth_b200 import PrefixSet, walk
    from reth_b200.walker import _items_arrays
    rng = np.random.default_rng(11)
    keys = rng.integers(0, 256, (5000, 32), dtype=np.uint8)
    keys = keys[sort_rows(keys)]
    vals = rand_vals(rng, len(keys))
    root, tables = slot_tables(eng, keys, vals)
    ks = [k.tobytes() for k in keys]
    elements, removed = walk(tables, PrefixSet([]), ks)
    assert sum(e.is_leaf for e in elements) == 0 and len(elements) <= 256
    k, nb, fl, v = _items_arrays(elements, 32)
    assert eng.root_from_items(k, nb, fl, v, None, None, account=False)[0].tobytes() == root


def test_state_incremental_equals_full(eng):
    """IncrementalStateRoot over (AccountsTrie / StoragesTrie rows + prefix sets) == StateRoot from scratch, root and tables,
    over blocks that change balances, create and destroy accounts, write / clear slots and wipe storages."""
    from reth_b200 import (Account, HashedPostState, HashedStorage, IncrementalStateRoot, MerkleStage, StateRoot, TrieUpdates)
    rng = np.random.default_rng(21)
    rk = lambda: bytes(rng.integers(0, 256, 32, dtype=np.uint8))
    state = HashedPostState()
    for _ in range(1500):
        k = rk()
        state.accounts[k] = Account(int(rng.integers(0, 50)), int(rng.integers(0, 2**60)), None)
        if rng.random() < 0.3:
            state.storages[k] = HashedStorage(False, {rk(): int(rng.integers(1, 2**62)) for _ in range(int(rng.integers(1, 40)))})
    root, upd = StateRoot(eng, state.into_sorted()).root_with_updates()
    tables = TrieUpdates()
    MerkleStage.write_trie_updates(tables, upd)
    for step in range(4):
        block = HashedPostState()
        live = sorted(k for k, a in state.accounts.items() if a is not None)
        for k in [live[int(i)] for i in rng.choice(len(live), 60, replace=False)]:
            r = rng.random()
            if r < 0.4:
                block.accounts[k] = Account(state.accounts[k].nonce + 1, int(rng.integers(0, 2**60)), None)
            elif r < 0.55:
                block.accounts[k] = None                                        # destroyed
                block.storages[k] = HashedStorage(True, {})
            elif r < 0.9:
                cur = state.storages.get(k)
                slots = {rk(): int(rng.integers(1, 2**62)) for _ in range(5)}
                if cur is not None and cur.storage:
                    for s in list(cur.storage)[:3]:
                        slots[s] = 0 if rng.random() < 0.5 else int(rng.integers(1, 2**62))
                block.storages[k] = HashedStorage(False, slots)
            else:
                block.storages[k] = HashedStorage(True, {rk(): 7})               # wiped, then one slot
        for _ in range(20):
            k = rk()
            block.accounts[k] = Account(0, 1, None)
            if rng.random() < 0.5:
                block.storages[k] = HashedStorage(False, {rk(): 5 for _ in range(3)})
        prefix_sets = block.construct_prefix_sets().freeze()
        state.extend(block)
        post = HashedPostState({k: a for k, a in state.accounts.items() if a is not None},
                               {k: HashedStorage(False, {s: v for s, v in st.storage.items() if v != 0})
                                for k, st in state.storages.items() if state.accounts.get(k) is not None})
        sorted_post = post.into_sorted()
        full_root, full_upd = StateRoot(eng, sorted_post).root_with_updates()
        full_tables = TrieUpdates()
        MerkleStage.write_trie_updates(full_tables, full_upd)
        inc = IncrementalStateRoot(eng, tables, sorted_post, prefix_sets)
        inc_root, inc_upd = inc.root_with_updates()
        assert inc_root == full_root
        assert inc.hashed_entries_walked < (len(post.accounts) + sum(len(s.storage) for s in post.storages.values())) // 3
        MerkleStage.write_trie_updates(tables, inc_upd)
        assert tables.account_nodes == full_tables.account_nodes
        rows = lambda t: {k: v.storage_nodes for k, v in t.storage_tries.items() if v.storage_nodes}
        assert rows(tables) == rows(full_tables)


def test_malformed_streams_are_rejected(eng):
    """A leaf below a hash item (the walker never yields that), descending keys, key_nibbles > 64: errors, not garbage."""
    from reth_b200 import B200Error, _lib
    h = oracle.keccak256(b"x")
    keys = np.zeros((2, 32), np.uint8)
    keys[0, 0] = 0xA0                      # hash item at path [a]
    keys[1, 0] = 0xA5                      # a leaf below it
    nib = np.array([1, 64], np.uint8)
    fl = np.zeros(2, np.uint8)
    vals = np.zeros((2, 32), np.uint8)
    vals[0] = np.frombuffer(h, np.uint8)
    vals[1, 31] = 7
    with pytest.raises(B200Error) as e:
        eng.root_from_items(keys, nib, fl, vals, None, None, account=False)
    assert e.value.status == _lib.ERR_UNSORTED
    with pytest.raises(B200Error) as e:
        eng.root_from_items(keys[::-1].copy(), nib[::-1].copy(), fl, vals[::-1].copy(), None, None, account=False)
    assert e.value.status == _lib.ERR_UNSORTED
    with pytest.raises(B200Error) as e:
        eng.root_from_items(keys, np.array([1, 65], np.uint8), fl, vals, None, None, account=False)
    assert e.value.status == _lib.ERR_INVALID_ARG
    # the context stays usable: a lone hash at the empty path is the root itself, a lone hash deeper hangs under an extension
    k1 = np.zeros((1, 32), np.uint8)
    assert eng.root_from_items(k1, np.array([0], np.uint8), fl[:1], vals[:1], None, None, account=False)[0].tobytes() == h
    hb = oracle.HashBuilder()
    hb.add_branch(bytes([0xA]), h, False)
    assert eng.root_from_items(keys[:1], nib[:1], fl[:1], vals[:1], None, None, account=False)[0].tobytes() == hb.root()
