"""Table rows: b200_account_trie_rows / b200_storage_trie_rows against an independent restatement of
the reference codecs.  Host-only code in the C ABI, so these run without a GPU.

The restatement below follows crates/trie/common/src/nibbles.rs (StoredNibbles :44-66, StoredNibblesSubKey :111-135,
PackedStoredNibbles/-SubKey :190-213,:274-297), crates/trie/common/src/storage.rs:24-44,70-86 and alloy-trie 0.9.5's
`Compact for BranchNodeCompact` (external; masks as big-endian u16, then hashes).  The nibble-key vectors are the ones
the reference's own tests assert (nibbles.rs:321-356,425-441)."""
import json
import os

import numpy as np
import pytest

import oracle
from reth_b200 import tables
from tests.util import alloc_to_flat, synth_accounts, synth_storage

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "genesis_allocs.json")


# ---- restatement of the codecs -------------------------------------------------------------------------------------
def stored_nibbles(path: bytes) -> bytes:
    return bytes(path)


def stored_nibbles_subkey(path: bytes) -> bytes:
    return bytes(path) + bytes(64 - len(path)) + bytes([len(path)])


def packed_stored_nibbles(path: bytes) -> bytes:
    out = bytearray(33)
    for i, n in enumerate(path):
        out[i // 2] |= n << 4 if i % 2 == 0 else n
    out[32] = len(path)
    return bytes(out)


def branch_node_compact(state: int, tree: int, hmask: int, hashes) -> bytes:
    assert len(hashes) == bin(hmask).count("1")
    return state.to_bytes(2, "big") + tree.to_bytes(2, "big") + hmask.to_bytes(2, "big") + b"".join(hashes)


def decode_branch_node_compact(buf: bytes):
    assert len(buf) % 32 == 6  # alloy-trie asserts exactly this
    state, tree, hmask = (int.from_bytes(buf[i:i + 2], "big") for i in (0, 2, 4))
    hashes = [buf[6 + 32 * i:38 + 32 * i] for i in range((len(buf) - 6) // 32)]
    assert len(hashes) == bin(hmask).count("1")  # no root_hash on stored nodes
    return state, tree, hmask, hashes


def expected_account_rows(records, packed):
    key = packed_stored_nibbles if packed else stored_nibbles
    rows = [(key(r[1]), branch_node_compact(r[2], r[3], r[4], r[5])) for r in records]
    return sorted(rows, key=lambda kv: kv[0])


def expected_storage_rows(records, acct_keys, packed):
    sub = packed_stored_nibbles if packed else stored_nibbles_subkey
    rows = [(acct_keys[r[0]].tobytes(), sub(r[1]) + branch_node_compact(r[2], r[3], r[4], r[5])) for r in records]
    return sorted(rows, key=lambda kv: (kv[0], kv[1][:33 if packed else 65]))


# ---- the reference's own key vectors ---------------------------------------------------------------------------------
def test_reference_nibble_key_vectors():
    h = [bytes([i]) * 32 for i in range(1, 3)]
    rec = [(0, bytes([2, 4]), 0b101, 0, 0b101, h)]
    (k, v), = tables.account_trie_rows(rec, tables.KEYS_LEGACY)
    assert k == bytes([2, 4])                                    # nibbles.rs:321-328
    assert v == bytes([0, 5, 0, 0, 0, 5]) + h[0] + h[1]
    keys = np.arange(32, dtype=np.uint8).reshape(1, 32)
    (k, v), = tables.storage_trie_rows(rec, keys, tables.KEYS_LEGACY)
    assert k == keys.tobytes() and len(v) == 65 + 6 + 64
    assert v[:2] == bytes([2, 4]) and v[2:64] == bytes(62) and v[64] == 2   # nibbles.rs:338-346
    full = bytes(i % 16 for i in range(64))                      # nibbles.rs:433-441
    (k, _), = tables.account_trie_rows([(0, full, 1, 0, 0, [])], tables.KEYS_PACKED)
    assert len(k) == 33 and k[32] == 64 and k[:32] == bytes((2 * i % 16) << 4 | (2 * i + 1) % 16 for i in range(32))
    (k, _), = tables.account_trie_rows([(0, bytes([0xA, 0xB, 0xC]), 1, 0, 0, [])], tables.KEYS_PACKED)
    assert k == bytes([0xAB, 0xC0]) + bytes(30) + bytes([3])     # nibbles.rs:443-450 (odd length: low nibble zero)


def test_rejects_malformed_records():
    with pytest.raises(Exception):
        tables.account_trie_rows([(0, b"", 1, 0, 0, [])])         # the empty path is never stored (updates.rs:140-158)
    with pytest.raises(Exception):
        tables.account_trie_rows([(0, bytes([1]), 1, 0, 0b11, [bytes(32)])])   # popcount(hash_mask) != hashes
    with pytest.raises(Exception):
        tables.storage_trie_rows([(5, bytes([1]), 1, 0, 0, [])], np.zeros((2, 32), np.uint8))  # trie_id out of range
    assert tables.account_trie_rows([]) == []


# ---- whole tries -----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("packed", [False, True])
def test_mainnet_genesis_account_rows(packed):
    alloc = json.load(open(GOLDEN))["mainnet"]
    keys, accts, skeys, svals, offs = alloc_to_flat(alloc["alloc"])
    root, ua, _ = oracle.state_root_full(keys, accts, skeys, svals, offs, want_updates=True)
    assert root.hex() == alloc["state_root"].removeprefix("0x")
    fmt = tables.KEYS_PACKED if packed else tables.KEYS_LEGACY
    rows = tables.account_trie_rows(ua, fmt)
    assert len(rows) == len(ua) > 100
    assert rows == expected_account_rows(ua, packed)
    assert [k for k, _ in rows] == sorted(k for k, _ in rows)     # MDBX append order
    # decode round trip: the rows carry exactly the records
    back = sorted((decode_branch_node_compact(v) for _, v in rows), key=repr)
    assert back == sorted(((r[2], r[3], r[4], r[5]) for r in ua), key=repr)


@pytest.mark.parametrize("packed", [False, True])
def test_storage_rows_synthetic(packed):
    n = 300
    keys, accts = synth_accounts(11, n)
    counts = np.where(np.arange(n) % 3 == 0, 0, 40 + (np.arange(n) * 7) % 90)
    skeys, svals, offs = synth_storage(11, counts, "mixed")
    _, ua, us = oracle.state_root_full(keys, accts, skeys, svals, offs, want_updates=True)
    assert len(us) > n // 2
    fmt = tables.KEYS_PACKED if packed else tables.KEYS_LEGACY
    rows = tables.storage_trie_rows(us, keys, fmt)
    assert rows == expected_storage_rows(us, keys, packed)
    sub = 33 if packed else 65
    order = [(k, v[:sub]) for k, v in rows]
    assert order == sorted(order)                                 # key, then dup subkey
    assert tables.account_trie_rows(ua, fmt) == expected_account_rows(ua, packed)


def test_packed_and_legacy_orders_agree():
    # PackedStoredNibbles keeps the nibble order under memcmp (nibbles.rs:385-414): both formats list the same nodes in
    # the same order, including a path that is a strict prefix of the next one.
    paths = [bytes([1]), bytes([1, 0]), bytes([1, 0, 0]), bytes([1, 0, 1]), bytes([1, 1]), bytes([0xF]), bytes([0, 0xF])]
    rec = [(0, p, 1 << (i % 16), 0, 0, []) for i, p in enumerate(paths)]
    legacy = tables.account_trie_rows(rec, tables.KEYS_LEGACY)
    packed = tables.account_trie_rows(rec, tables.KEYS_PACKED)
    assert [v for _, v in legacy] == [v for _, v in packed]
    assert [k for k, _ in legacy] == sorted(paths)


def test_raw_updates_path():
    # rows_from_updates takes the C structs as filled by the library.  orc_updates has the layout of b200_updates minus
    # the trailing owner pointer (oracle.h / b200trie.h), so the oracle's raw output drives the same code here on CPU.
    import ctypes as C
    n = 120
    keys, accts = synth_accounts(12, n)
    skeys, svals, offs = synth_storage(12, np.full(n, 48), "u64")
    L = oracle.lib()
    root = np.empty(32, np.uint8)
    ua, us = oracle._Updates(), oracle._Updates()
    rc = L.orc_state_root_full(keys.ctypes.data, accts.ctypes.data, n, skeys.ctypes.data, svals.ctypes.data,
                               offs.ctypes.data, root.ctypes.data, C.byref(ua), C.byref(us), 1)
    assert rc == 0
    arows, srows = tables.rows_from_updates(ua, us, keys, tables.KEYS_PACKED)
    a_list, s_list = arows.to_list(), srows.to_list()
    arows.release(); srows.release()
    rec_a, rec_s = oracle._updates_to_py(ua), oracle._updates_to_py(us)   # frees the oracle buffers
    assert a_list == expected_account_rows(rec_a, True)
    assert s_list == expected_storage_rows(rec_s, keys, True) and len(s_list) > 0


def test_row_codec_round_trip_property():
    """hypothesis: arbitrary node sets encode to rows that decode back to the same nodes, in key order, for both key formats."""
    from hypothesis import given, settings, strategies as st

    node = st.tuples(st.lists(st.integers(0, 15), min_size=1, max_size=64).map(bytes), st.integers(1, 0xFFFF), st.integers(0, 0xFFFF),
                     st.integers(0, 0xFFFF))

    @settings(max_examples=60, deadline=None)
    @given(st.lists(node, max_size=40, unique_by=lambda t: t[0]), st.booleans())
    def check(nodes, packed):
        recs = [(0, p, sm, tm, hm, [bytes([i]) * 32 for i in range(bin(hm).count("1"))]) for p, sm, tm, hm in nodes]
        fmt = tables.KEYS_PACKED if packed else tables.KEYS_LEGACY
        rows = tables.account_trie_rows(recs, fmt)
        assert rows == expected_account_rows(recs, packed)
        for (k, v), (_, p, sm, tm, hm, hs) in zip(rows, sorted(recs, key=lambda r: r[1])):
            assert decode_branch_node_compact(v) == (sm, tm, hm, hs)
            assert (k[:len(p) // 2 + 1] if packed else k)[:1] is not None   # keys checked above against the restatement

    check()


def test_stored_subnode_codec_roundtrip():
    """crates/trie/common/src/subnode.rs:70-96 `subnode_roundtrip`, plus the byte layout of :16-47 spelled out, and the
    BranchNodeCompact codec against the C encoder's row values."""
    from reth_b200 import BranchNodeCompact
    from reth_b200.tables import StoredSubNode, branch_node_compact_from_bytes, branch_node_compact_to_bytes
    node = BranchNodeCompact(1, 0, 1, (bytes(32),), None)
    sub = StoredSubNode(b"", None, node)
    enc = sub.to_compact()
    assert enc == b"\x00\x00" + b"\x00" + b"\x01" + b"\x00\x01\x00\x00\x00\x01" + bytes(32)
    assert StoredSubNode.from_compact(enc) == sub
    rng = np.random.default_rng(3)
    for _ in range(200):
        hm = int(rng.integers(0, 1 << 16))
        hashes = tuple(bytes(rng.integers(0, 256, 32, dtype=np.uint8)) for _ in range(bin(hm).count("1")))
        root = bytes(rng.integers(0, 256, 32, dtype=np.uint8)) if rng.random() < 0.3 else None
        nd = BranchNodeCompact(int(rng.integers(1, 1 << 16)) | hm, int(rng.integers(0, 1 << 16)), hm, hashes, root)
        assert branch_node_compact_from_bytes(branch_node_compact_to_bytes(nd)) == nd
        s2 = StoredSubNode(bytes(rng.integers(0, 16, int(rng.integers(0, 65)), dtype=np.uint8)),
                           int(rng.integers(0, 16)) if rng.random() < 0.5 else None, nd if rng.random() < 0.8 else None)
        assert StoredSubNode.from_compact(s2.to_compact()) == s2
    # the row values the C encoder writes are this codec (no root hash in table rows)
    recs = [(0, bytes([1, 2, 3]), 0b1011, 0b0001, 0b1001, [bytes([7]) * 32, bytes([9]) * 32])]
    rows = tables.account_trie_rows(recs)
    assert rows[0][1] == branch_node_compact_to_bytes(BranchNodeCompact(0b1011, 0b0001, 0b1001, (bytes([7]) * 32, bytes([9]) * 32)))


def test_hash_builder_state_codec_roundtrip():
    """crates/trie/common/src/hash_builder/state.rs:149-170: `hash_builder_state_regression` (a default state with one
    default — empty — RlpNode on the stack) with its bytes spelled out, and `hash_builder_state_roundtrip` over random
    states.  (No byte vector of this codec exists in the reference tree; the element codecs are restated from reth-codecs.)"""
    from reth_b200.tables import HashBuilderState
    st = HashBuilderState(stack=[b""])
    enc = st.to_compact()
    assert enc == b"\x00" + b"\x00\x01" + b"\x00\x00" + b"\x01\x00" + bytes(6) + b"\x00"
    assert HashBuilderState.from_compact(enc) == st
    # a key of nibbles 0, 5, 15: count 3, then (len 0), (len 1, 5), (len 1, 15); a 32-byte hash as the pending value
    st = HashBuilderState(key=bytes([0, 5, 15]), value=("hash", bytes(range(32))), stack=[b"\xa0" + bytes(32), b"\xc2\x80\x80"],
                          groups=[0b101, 0], tree_masks=[1, 0], hash_masks=[0x8000, 0], stored_in_database=True)
    enc = st.to_compact()
    assert enc.startswith(b"\x03\x00\x01\x05\x01\x0f" + b"\x00\x02" + b"\x00\x21\xa0")
    assert enc.endswith(b"\x00\x02\x00\x05\x00\x00" + b"\x00\x02\x00\x01\x00\x00" + b"\x00\x02\x80\x00\x00\x00" + b"\x01")
    assert HashBuilderState.from_compact(enc) == st
    rng = np.random.default_rng(11)
    for _ in range(300):
        depth = int(rng.integers(0, 65))
        masks = lambda: [int(x) for x in rng.integers(0, 1 << 16, depth)]
        value = ("hash", bytes(rng.integers(0, 256, 32, dtype=np.uint8))) if rng.random() < 0.4 else \
            ("bytes", bytes(rng.integers(0, 256, int(rng.integers(0, 200)), dtype=np.uint8)))
        stack = [bytes(rng.integers(0, 256, int(rng.integers(0, 34)), dtype=np.uint8)) for _ in range(int(rng.integers(0, 20)))]
        s = HashBuilderState(bytes(rng.integers(0, 16, depth, dtype=np.uint8)), value, stack, masks(), masks(), masks(),
                             bool(rng.integers(0, 2)))
        assert HashBuilderState.from_compact(s.to_compact()) == s
    with pytest.raises(ValueError):
        HashBuilderState(key=bytes([16]))
