"""Property tests of the oracle: the HashBuilder restatement against the independent recursive trie
(the reference does the same against `triehash`: crates/trie/db/tests/trie.rs:242-295,
crates/trie/parallel/src/root.rs:287-400)."""
import numpy as np
import pytest

import oracle
from tests.util import random_keys, sort_rows, synth_accounts, synth_storage


def _keys_with_shared_prefixes(rng, n, max_share):
    keys = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    for i in range(1, n):
        if rng.random() < 0.6:
            share = int(rng.integers(0, max_share + 1))  # nibbles copied from a previous key
            src = keys[int(rng.integers(0, i))]
            nb = share // 2
            keys[i, :nb] = src[:nb]
            if share & 1:
                keys[i, nb] = (src[nb] & 0xF0) | (keys[i, nb] & 0x0F)
    keys = np.unique(keys, axis=0)
    return keys[sort_rows(keys)]


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 17, 40, 300])
def test_hash_builder_vs_recursive(n):
    rng = np.random.default_rng(1000 + n)
    for case in range(40):
        keys = _keys_with_shared_prefixes(rng, n, max_share=63)
        vals = []
        for _ in range(len(keys)):
            ln = int(rng.integers(1, 34))
            v = rng.integers(0, 256, ln, dtype=np.uint8).tobytes()
            if rng.random() < 0.4:
                v = bytes([int(rng.integers(1, 0x80))])  # tiny value -> inline (<32 B) leaf nodes when suffix is short
            vals.append(v)
        hb = oracle.HashBuilder()
        for k, v in zip(keys, vals):
            hb.add_leaf(oracle.unpack_nibbles(k.tobytes()), v)
        assert hb.root() == oracle.trie_root_recursive(keys, vals), (n, case)


def test_storage_roots_vs_recursive_and_threads():
    counts = [0, 1, 2, 16, 0, 300, 5, 1, 0, 33]
    keys, vals, offs = synth_storage(11, counts, value_mode="mixed")
    roots = oracle.storage_roots(keys, vals, offs)
    roots_mt, upd_mt = oracle.storage_roots(keys, vals, offs, want_updates=True, threads=3)
    _, upd_st = oracle.storage_roots(keys, vals, offs, want_updates=True, threads=1)
    assert (roots == roots_mt).all() and upd_mt == upd_st
    for a, c in enumerate(counts):
        s, e = int(offs[a]), int(offs[a + 1])
        enc = [oracle.encode_u256(int.from_bytes(vals[i].tobytes(), "big")) for i in range(s, e)]
        assert roots[a].tobytes() == oracle.trie_root_recursive(keys[s:e], enc)
    assert roots[0].tobytes() == oracle.EMPTY_ROOT_HASH


def test_state_root_full_vs_recursive():
    n = 2000
    akeys, accs = synth_accounts(5, n)
    counts = (np.arange(n) % 7 == 0) * 16 + (np.arange(n) % 501 == 0) * 400
    skeys, svals, offs = synth_storage(6, counts)
    root = oracle.state_root_full(akeys, accs, skeys, svals, offs)
    sroots = oracle.storage_roots(skeys, svals, offs, threads=2)
    leaves = [oracle.encode_trie_account(int(accs[i]["nonce"]), int.from_bytes(accs[i]["balance"].tobytes(), "big"),
                                         sroots[i].tobytes(), accs[i]["code_hash"].tobytes()) for i in range(n)]
    assert root == oracle.trie_root_recursive(akeys, leaves)
    assert root == oracle.state_root_full(akeys, accs, skeys, svals, offs, threads=4)


def test_zero_value_slot_rejected():
    keys = random_keys(3, 2)
    keys = keys[sort_rows(keys)]
    vals = np.zeros((2, 32), np.uint8)
    vals[0, 31] = 1
    with pytest.raises(ValueError):
        oracle.storage_roots(keys, vals, [0, 2])


def test_unsorted_keys_rejected():
    keys = random_keys(4, 3)
    keys = keys[sort_rows(keys)][::-1].copy()
    vals = np.ones((3, 32), np.uint8)
    with pytest.raises(ValueError):
        oracle.storage_roots(keys, vals, [0, 3])


def test_structure_stats_match_survey_appendix_c():
    """~1.38 hashed nodes and ~1.54 Keccak-f per leaf for 10k random accounts."""
    akeys, accs = synth_accounts(9, 10_000, with_code=False)
    oracle.stats_reset()
    oracle.state_root(akeys, accs)
    s = oracle.stats()
    assert s["leaves"] == 10_000
    assert 1.30 < s["hashed_nodes"] / s["leaves"] < 1.45
    assert 1.40 < s["keccak_f"] / s["leaves"] < 1.65
