"""Streamed / resumable state root (b200_root_stream_*): the state pushed in ascending account-key ranges must
give exactly the root and the stored nodes of one b200_state_root_full call over the whole state — the criterion of reth's
own threshold tests (crates/trie/db/tests/trie.rs `arbitrary_state_root_with_progress`: root_with_progress looped over
intermediate states == the one-shot root) — and a checkpoint taken in between must resume to the same result."""
import numpy as np
import pytest

import oracle
from tests.util import synth_accounts, synth_storage

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def eng():
    from reth_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def _state(seed, n, slots_of):
    akeys, accs = synth_accounts(seed, n)
    counts = np.array([slots_of(i) for i in range(n)], np.int64)
    skeys, svals, offs = synth_storage(seed + 1, counts)
    return akeys, accs, skeys, svals, offs


def _slice(state, a0, a1):
    akeys, accs, skeys, svals, offs = state
    s0, s1 = int(offs[a0]), int(offs[a1])
    return akeys[a0:a1], accs[a0:a1], skeys[s0:s1], svals[s0:s1], (offs[a0:a1 + 1] - offs[a0]).astype(np.uint64)


def _one_shot(eng, state):
    root, au, su = eng.state_root_full(*state, want_updates=True)
    acct = {r[1]: r[2:] for r in au}
    stor = {(state[0][r[0]].tobytes(), r[1]): r[2:] for r in su}
    return root, acct, stor


def _stream(eng, state, cuts, retain=True):
    from reth_b200 import RootStream
    s = RootStream(eng, retain_updates=retain)
    acct, stor = {}, {}
    n = len(state[0])
    bounds = [0] + list(cuts) + [n]
    k = 0
    while k < len(bounds) - 1:
        a0, a1 = bounds[k], bounds[k + 1]
        part = _slice(state, a0, a1)
        res = s.push(*part)
        if retain:
            prog, au, su = res
            for r in au:
                assert r[1] not in acct
                acct[r[1]] = r[2:]
            for r in su:
                stor[(part[0][r[0]].tobytes(), r[1])] = r[2:]
        else:
            prog = res
        assert prog["accounts"] == a1
        k += 1
    res = s.finish()
    if retain:
        root, au = res
        for r in au:
            assert r[1] not in acct
            acct[r[1]] = r[2:]
    else:
        root = res
    s.close()
    return root, acct, stor


@pytest.mark.parametrize("n,cuts", [(3000, [1, 700, 701, 1500, 2999]), (3000, []), (40, [10, 20, 30]), (5000, [2500])])
def test_pushes_equal_one_shot_root_and_updates(eng, n, cuts):
    state = _state(70 + n, n, lambda i: (i % 5 == 0) * (1 + i % 23))
    root, acct, stor = _one_shot(eng, state)
    assert root == oracle.state_root_full(*state, threads=2)
    s_root, s_acct, s_stor = _stream(eng, state, cuts)
    assert s_root == root
    assert s_acct == acct
    assert s_stor == stor
    assert _stream(eng, state, cuts, retain=False)[0] == root


def test_degenerate_shapes(eng):
    from reth_b200 import RootStream
    # nothing at all
    s = RootStream(eng)
    assert s.finish() == oracle.EMPTY_ROOT_HASH
    s.close()
    # a single account; all accounts in one bucket pushed one by one; two buckets with one account each
    akeys, accs = synth_accounts(9, 6)
    z = np.zeros((0, 32), np.uint8)
    for keyset in ([0], [0, 1, 2, 3, 4, 5]):
        k = akeys[keyset].copy()
        k[:, 0] = (k[:, 0] & 0x0F) | 0x70            # same top nibble
        order = np.lexsort(tuple(k[:, i] for i in range(31, -1, -1)))
        k, a = k[order], accs[keyset][order]
        s = RootStream(eng)
        for i in range(len(k)):
            s.push(k[i:i + 1], a[i:i + 1], z, z, np.array([0, 0], np.uint64))
        assert s.finish() == oracle.state_root(k, a)
        s.close()
    k2 = akeys[:2].copy()
    k2[0, 0], k2[1, 0] = 0x10, 0xF0
    s = RootStream(eng)
    s.push(k2[:1], accs[:1], z, z, np.array([0, 0], np.uint64))
    s.push(k2[1:], accs[1:2], z, z, np.array([0, 0], np.uint64))
    assert s.finish() == oracle.state_root(k2, accs[:2])
    s.close()


def test_out_of_order_push_is_rejected(eng):
    from reth_b200 import B200Error, RootStream, _lib
    akeys, accs = synth_accounts(10, 100)
    z = np.zeros((0, 32), np.uint8)
    offs = lambda m: np.zeros(m + 1, np.uint64)
    s = RootStream(eng)
    s.push(akeys[50:], accs[50:], z, z, offs(50))
    with pytest.raises(B200Error) as e:
        s.push(akeys[:50], accs[:50], z, z, offs(50))
    assert e.value.status == _lib.ERR_UNSORTED
    s.close()


@pytest.mark.parametrize("after", [0, 1, 2])
def test_checkpoint_resume(eng, after):
    """MerkleCheckpoint round trip: stop after a push, keep 1104 bytes, resume in a fresh stream (the open bucket is pushed
    again from its first key) — same root."""
    from reth_b200 import RootStream
    n = 4000
    state = _state(91, n, lambda i: (i % 7 == 0) * 3)
    root = oracle.state_root_full(*state, threads=2)
    cuts = [900, 1800, 3100]
    bounds = [0] + cuts + [n]
    s = RootStream(eng)
    for k in range(after + 1):
        s.push(*_slice(state, bounds[k], bounds[k + 1]))
    cp = s.checkpoint()
    s.close()
    assert len(cp) == 1104
    nib = RootStream.resume_nibble(cp)
    top = state[0][:, 0] >> 4
    restart = int(np.searchsorted(top, nib))
    assert restart <= bounds[after + 1]
    s = RootStream.resume(eng, cp)
    rest = [restart] + [b for b in bounds if b > restart]
    for k in range(len(rest) - 1):
        s.push(*_slice(state, rest[k], rest[k + 1]))
    assert s.finish() == root
    s.close()


def _hps(state):
    from reth_b200 import Account, HashedPostStateSorted, HashedStorageSorted
    akeys, accs, skeys, svals, offs = state
    accounts, storages = [], {}
    for i in range(len(akeys)):
        a = accs[i]
        ch = bytes(a["code_hash"])
        accounts.append((akeys[i].tobytes(), Account(int(a["nonce"]), int.from_bytes(bytes(a["balance"]), "big"),
                                                       None if ch == oracle.KECCAK_EMPTY else ch)))
        if offs[i + 1] > offs[i]:
            storages[akeys[i].tobytes()] = HashedStorageSorted(
                [(skeys[j].tobytes(), int.from_bytes(svals[j].tobytes(), "big")) for j in range(int(offs[i]), int(offs[i + 1]))])
    return HashedPostStateSorted(accounts, storages)


def test_state_root_with_threshold_mirror(eng):
    """StateRoot::with_threshold(..).root_with_progress() looped over with_intermediate_state == root_with_updates()
    (reth: crates/trie/db/tests/trie.rs arbitrary_state_root_with_progress)."""
    from reth_b200 import StateRoot
    state = _state(55, 1500, lambda i: (i % 4 == 0) * (2 + i % 9))
    hps = _hps(state)
    root, full = StateRoot(eng, hps).root_with_updates()
    assert root == oracle.state_root_full(*state, threads=2)
    for threshold in (1, 300, 10_000_000):
        inter, steps, walked = None, 0, 0
        acct_nodes, storage = {}, {}
        while True:
            p = StateRoot(eng, hps).with_threshold(threshold).with_intermediate_state(inter).root_with_progress()
            steps += 1
            walked += p.hashed_entries_walked
            acct_nodes.update(p.updates.account_nodes)
            storage.update(p.updates.storage_tries)
            if p.complete:
                assert p.state is None and p.root == root
                break
            assert p.root is None and p.state is not None
            inter = p.state
        assert walked == len(state[0]) + len(state[2])
        assert acct_nodes == full.account_nodes and storage == full.storage_tries
        assert (steps == 1) == (threshold >= walked)
        if threshold == 1:
            assert steps == len(state[0])        # one account (with its storage) per step at least


def test_merkle_stage_chunked_rebuild(eng):
    """MerkleStage with an inner threshold: execute() returns None (not done) after max_steps ranges, keeps its inner
    checkpoint, and continues; root and trie tables equal the one-shot rebuild (merkle.rs:210-310)."""
    from reth_b200 import MerkleStage, Tables
    state = _state(57, 1200, lambda i: (i % 3 == 0) * 4)
    hps = _hps(state)
    t1, t2 = Tables(), Tables()
    for t in (t1, t2):
        t.hashed_accounts = list(hps.accounts)
        t.hashed_storages = {k: list(v.storage_slots) for k, v in hps.storages.items()}
    st = MerkleStage(eng)
    root = st.execute(t1)
    calls = 0
    while True:
        r = st.execute(t2, expected_state_root=root, threshold=500, max_steps=2)
        calls += 1
        if r is not None:
            break
        assert t2.merkle_checkpoint is not None and len(t2.merkle_checkpoint[1]) == 1104
    assert r == root and calls > 2 and t2.merkle_checkpoint is None
    assert t2.trie_updates.account_nodes == t1.trie_updates.account_nodes
    # (the one-shot leg stores the TrieUpdates as they come — empty storages as is_deleted markers —, the chunked leg
    # applies every step to the tables: compare the rows)
    rows = lambda tu: {k: v.storage_nodes for k, v in tu.storage_tries.items() if v.storage_nodes}
    assert rows(t2.trie_updates) == rows(t1.trie_updates)


def test_stream_misuse_is_an_error(eng):
    from reth_b200 import B200Error, RootStream, _lib
    akeys, accs = synth_accounts(12, 10)
    z = np.zeros((0, 32), np.uint8)
    offs = np.zeros(11, np.uint64)
    s = RootStream(eng)
    s.push(akeys, accs, z, z, offs)
    root = s.finish()
    assert root == oracle.state_root(akeys, accs)
    for call in (lambda: s.finish(), lambda: s.push(akeys, accs, z, z, offs)):
        with pytest.raises(B200Error) as e:
            call()
        assert e.value.status == _lib.ERR_INVALID_ARG
    s.close()
    with pytest.raises(B200Error):                      # a checkpoint that claims a closed bucket at or after the resume point
        cp = bytearray(RootStream(eng).checkpoint())
        cp[1088:1092] = (1 << 5).to_bytes(4, "little")   # closed_mask: bucket 5
        cp[1092:1096] = (3).to_bytes(4, "little")        # resume_nibble 3
        RootStream.resume(eng, bytes(cp))
