"""GPU tests written against the host mirror of reth's interface (StateRoot / StorageRoot / ParallelStateRoot /
HashedPostState / the hashing + merkle stages) so they read like the reference's own tests."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from reth_b200 import (Account, AccountHashingStage, Engine, HashedPostState, HashedStorage, MerkleStage,
                       ParallelStateRoot, StageError, StateRoot, StorageHashingStage, StorageRoot, Tables,
                       EMPTY_ROOT_HASH)
from tests.util import alloc_to_flat

H = bytes.fromhex
ETHER = 10**18


@pytest.fixture(scope="module")
def eng():
    e = Engine(0)
    yield e
    e.close()


def test_from_bundle_state_known_root(eng):
    """from_bundle_state_with_rayon — crates/trie/db/src/state.rs:408-437."""
    a1, a2 = bytes(19) + b"\x01", bytes(19) + b"\x02"
    bundle = [
        (a1, {"info": Account(nonce=1), "storage": {1015: 10}}),
        (a2, {"info": Account(nonce=2), "storage": {2015: 20}}),
    ]
    post_state = HashedPostState.from_bundle_state(eng, bundle)
    assert len(post_state.accounts) == 2 and len(post_state.storages) == 2
    root = StateRoot(eng, post_state.into_sorted()).root()
    assert root.hex() == "b464525710cafcf5d4044ac85b72c08b1e76231b8d91f288fe438cc41d8eaafd"


def test_account_and_storage_trie_via_mirror(eng):
    """account_and_storage_trie — crates/trie/db/tests/trie.rs:357-477 (root, two stored account nodes)."""
    storage = {H("12" + "00" * 31): 0x42, H("14" + "00" * 31): 0x01,
               H("30" + "00" * 28 + "E00000"): 0x127A89, H("30" + "00" * 28 + "E00001"): 0x05}
    key3 = oracle.keccak256(H("16b07afd1c635f77172e842a000ead9a2a222459"))
    state = HashedPostState(
        accounts={
            H("b0" + "00" * 31): Account(0, 3 * ETHER),
            oracle.keccak256(H("7db3e81b72d2695e19764583f6d219dbee0f35ca")): Account(0, ETHER),
            key3: Account(0, 2 * ETHER, H("5be74cad16203c4905c068b012a2e9fb6d19d036c410f16fd177f337541440dd")),
            H("B1A0" + "00" * 30): Account(0, 4 * ETHER),
            H("B310" + "00" * 30): Account(0, 8 * ETHER),
            H("B340" + "00" * 30): Account(0, 1 * ETHER),
        },
        storages={key3: HashedStorage(False, storage)},
    ).into_sorted()
    root, updates = StateRoot(eng, state).root_with_updates()
    assert root.hex() == "72861041bc90cd2f93777956f058a545412b56de79af5eb6b8075fe2eabbe015"
    nodes = sorted(updates.account_nodes.items())
    assert [list(p) for p, _ in nodes] == [[0xB], [0xB, 0x0]]
    n1, n2 = nodes[0][1], nodes[1][1]
    assert (n1.state_mask, n1.tree_mask, n1.hash_mask, len(n1.hashes), n1.root_hash) == (0b1011, 0b0001, 0b1001, 2, None)
    assert (n2.state_mask, n2.tree_mask, n2.hash_mask, len(n2.hashes)) == (0b10001, 0, 0b10000, 1)
    # every account without storage carries StorageTrieUpdates::deleted() (trie.rs:622-629 + updates.rs:126-137)
    assert sum(u.is_deleted for u in updates.storage_tries.values()) == 5
    assert key3 not in updates.storage_tries  # its 4-leaf trie stores no branch node
    # ParallelStateRoot and the storage root on its own agree
    assert ParallelStateRoot(eng, state).incremental_root() == root
    sroot = StorageRoot(eng, key3, state.storages[key3]).root()
    o = oracle.storage_roots(np.frombuffer(b"".join(sorted(storage)), np.uint8).reshape(-1, 32),
                             np.frombuffer(b"".join(int(storage[k]).to_bytes(32, "big") for k in sorted(storage)),
                                           np.uint8).reshape(-1, 32), [0, 4])
    assert sroot == o[0].tobytes()


def test_storage_root_empty_and_zero_values(eng):
    from reth_b200 import HashedStorageSorted
    r, walked, upd = StorageRoot(eng, b"\x01" * 32, HashedStorageSorted([])).root_with_updates()
    assert r == EMPTY_ROOT_HASH and walked == 0 and upd.is_deleted
    # zero-valued slots are deletions: a storage holding only zeros is empty
    r2, _, _ = StorageRoot(eng, b"\x01" * 32, HashedStorageSorted([(b"\x05" * 32, 0)])).root_with_updates()
    assert r2 == EMPTY_ROOT_HASH


def test_hashing_and_merkle_stages_on_genesis(eng, golden_allocs):
    """The three stages of HashingStages (crates/stages/stages/src/sets.rs:427-442) over the holesky genesis alloc
    (contract code + storage); MerkleStage validates against the stateRoot of the genesis file and fails loudly on
    a wrong expectation (merkle.rs:437-453)."""
    g = golden_allocs["holesky"]
    t = Tables()
    for addr, e in g["alloc"].items():
        code = bytes.fromhex(e["code"][2:]) if e.get("code") else b""
        t.plain_accounts[bytes.fromhex(addr)] = Account(
            int(e.get("nonce", "0x0"), 16), int(e["balance"], 16), eng.keccak256(code) if code else None)
        if e.get("storage"):
            t.plain_storage[bytes.fromhex(addr)] = {int(k, 16): int(v, 16) for k, v in e["storage"].items()}
    assert AccountHashingStage(eng).execute(t) == len(g["alloc"])
    keys = [k for k, _ in t.hashed_accounts]
    assert keys == sorted(keys)
    StorageHashingStage(eng).execute(t)
    root = MerkleStage(eng).execute(t, expected_state_root=bytes.fromhex(g["state_root"]))
    assert root.hex() == g["state_root"]
    assert t.trie_updates is not None and len(t.trie_updates.account_nodes) > 0
    with pytest.raises(StageError):
        MerkleStage(eng).execute(t, expected_state_root=b"\x00" * 32)
    # same state through the flat path
    assert eng.state_root_full(*alloc_to_flat(g["alloc"], keccak_rows=eng.keccak256_fixed)).hex() == g["state_root"]


def test_destroyed_accounts_and_wiped_storage(eng):
    """A destroyed account (None) disappears from the trie; construct_prefix_sets reports it and finalize marks its
    storage trie deleted (updates.rs:153-157)."""
    k1, k2 = b"\x11" * 32, b"\x22" * 32
    st = HashedPostState(accounts={k1: Account(1, 1), k2: None},
                         storages={k2: HashedStorage(True, {})})
    sets = st.construct_prefix_sets().freeze()
    root, updates = StateRoot(eng, st.into_sorted()).with_prefix_sets(sets).root_with_updates()
    only = HashedPostState(accounts={k1: Account(1, 1)}).into_sorted()
    assert root == StateRoot(eng, only).root()
    assert updates.storage_tries[k2].is_deleted


@pytest.mark.parametrize("dynamic", [False, True])
def test_resident_state_root_commits_blocks(eng, dynamic):
    """Live path: fold a sequence of per-block HashedPostStates (balance changes, new accounts, destroyed accounts,
    storage writes / zeroing / wipes) into a resident state; after every block the root equals a from-scratch
    StateRoot over the merged state — the reference's incremental == full criterion
    (crates/trie/db/tests/trie.rs:680-717, crates/trie/parallel/src/root.rs:287-400)."""
    from reth_b200 import ResidentStateRoot
    rng = np.random.default_rng(77)
    rk = lambda: bytes(rng.integers(0, 256, 32, dtype=np.uint8))
    base = HashedPostState()
    for _ in range(800):
        k = rk()
        base.accounts[k] = Account(int(rng.integers(0, 50)), int(rng.integers(1, 2**62)))
        if rng.random() < 0.25:
            base.storages[k] = HashedStorage(False, {rk(): int(rng.integers(1, 2**62)) for _ in range(int(rng.integers(1, 30)))})
    merged = HashedPostState(dict(base.accounts), {k: HashedStorage(False, dict(v.storage)) for k, v in base.storages.items()})
    rs = ResidentStateRoot(eng, base.into_sorted(), dynamic=dynamic)
    assert rs.root() == StateRoot(eng, base.into_sorted()).root()
    for block in range(5):
        post = HashedPostState()
        live = [k for k, a in merged.accounts.items() if a is not None]
        for k in rng.choice(len(live), 40, replace=False):
            a = merged.accounts[live[k]]
            post.accounts[live[k]] = Account(a.nonce + 1, a.balance + 7, a.bytecode_hash)
        if block % 2 == 0:
            for _ in range(10):
                post.accounts[rk()] = Account(0, int(rng.integers(1, 10**18)))               # new accounts
            for k in rng.choice(len(live), 5, replace=False):
                post.accounts[live[k]] = None                                                 # destroyed
                post.storages[live[k]] = HashedStorage(True, {})
        with_storage = [k for k in merged.storages if merged.accounts.get(k) is not None and k not in post.accounts]
        for k in with_storage[:6]:
            st = merged.storages[k].storage
            slots = list(st)
            changes = {slots[0]: 0} if slots else {}                                          # zero = delete
            changes[rk()] = int(rng.integers(1, 2**60))
            post.storages[k] = HashedStorage(block == 3, changes)                             # one block wipes
        root, rebuilt = rs.commit(post)
        # model: merge like HashedPostState::extend + deletion rules
        for k, hs in post.storages.items():
            cur = {} if hs.wiped else dict(merged.storages.get(k, HashedStorage()).storage)
            for s, v in hs.storage.items():
                if v == 0:
                    cur.pop(s, None)
                else:
                    cur[s] = v
            merged.storages[k] = HashedStorage(False, cur)
        for k, a in post.accounts.items():
            if a is None:
                merged.accounts.pop(k, None)
                merged.storages.pop(k, None)
            else:
                merged.accounts[k] = a
        assert root == StateRoot(eng, merged.into_sorted()).root(), block
        assert rebuilt == (block % 2 == 0 and not dynamic)
    rs.close()


def test_verifier_reports_extra_wrong_and_missing_nodes(eng):
    """reth's Verifier (crates/trie/trie/src/verify.rs, tests :540-760): consistent tables give no output; a dropped node is
    Missing, a planted one Extra, a node with a flipped mask Wrong — for the accounts trie and for a storage trie."""
    from dataclasses import replace
    from reth_b200 import Verifier
    from reth_b200.verify import Extra, Missing, Wrong
    rng = np.random.default_rng(4)
    rk = lambda: bytes(rng.integers(0, 256, 32, dtype=np.uint8))
    st = HashedPostState()
    for i in range(600):
        k = rk()
        st.accounts[k] = Account(i % 5, 10 + i)
        if i % 50 == 0:
            st.storages[k] = HashedStorage(False, {rk(): int(rng.integers(1, 2**60)) for _ in range(400)})
    sorted_state = st.into_sorted()
    _, tables = StateRoot(eng, sorted_state).root_with_updates()
    v = Verifier(eng, sorted_state)
    assert v.verify(tables) == []
    apaths = sorted(tables.account_nodes)
    dropped = tables.account_nodes.pop(apaths[3])
    flipped_path = apaths[7]
    good = tables.account_nodes[flipped_path]
    tables.account_nodes[flipped_path] = replace(good, tree_mask=good.tree_mask ^ 1)
    planted_path = bytes([0xF, 0xF, 0xF, 0xF, 0xF])
    tables.account_nodes[planted_path] = good
    addr = next(a for a, s in tables.storage_tries.items() if len(s.storage_nodes) > 3)
    spaths = sorted(tables.storage_tries[addr].storage_nodes)
    sdropped = tables.storage_tries[addr].storage_nodes.pop(spaths[1])
    out = v.verify(tables)
    assert Missing(None, apaths[3], dropped) in out
    assert Wrong(None, flipped_path, good, tables.account_nodes[flipped_path]) in out
    assert Extra(None, planted_path, good) in out
    assert Missing(addr, spaths[1], sdropped) in out
    assert len(out) == 4
