"""CPU tests of the host-side mirror (reth_b200/hashed_state.py, trie.py): the parts of reth's interface that do
not touch the device.  Expected behaviour is quoted from the reference's own unit tests."""
import numpy as np

from reth_b200.hashed_state import (Account, HashedPostState, HashedPostStateSorted, HashedStorage,
                                    HashedStorageSorted, PrefixSetMut, unpack_nibbles)
from reth_b200.engine import KECCAK_EMPTY
from reth_b200.trie import StorageTrieUpdates, TrieUpdates


def nib(*xs):
    return bytes(xs)


def test_prefix_set_contains_with_duplicates():
    """crates/trie/common/src/prefix_set.rs:292-305."""
    m = PrefixSetMut()
    for k in (nib(1, 2, 3), nib(1, 2, 4), nib(4, 5, 6), nib(1, 2, 3)):
        m.insert(k)
    ps = m.freeze()
    assert ps.contains(nib(1, 2)) and ps.contains(nib(4, 5)) and not ps.contains(nib(7, 8))
    assert len(ps) == 3


def test_prefix_set_cursor_moves_both_ways():
    ps = PrefixSetMut([nib(1), nib(3, 3), nib(5, 0, 1), nib(9)]).freeze()
    assert ps.contains(nib(9))
    assert ps.contains(nib(1))       # cursor must walk back
    assert not ps.contains(nib(2))
    assert ps.contains(nib(5, 0))
    assert not ps.contains(nib(5, 1))
    assert ps.contains(nib(3))


def test_prefix_set_all_extend():
    """prefix_set.rs:346-351."""
    m = PrefixSetMut()
    m.extend(PrefixSetMut.all_())
    assert m.all
    assert m.freeze().contains(nib(0xF, 0xF))


def test_construct_prefix_sets():
    """hashed_state.rs:105-126: accounts and storage owners enter the account prefix set, destroyed accounts are
    collected, a wiped storage yields PrefixSet::all."""
    a1, a2, a3 = bytes([1]) * 32, bytes([2]) * 32, bytes([3]) * 32
    st = HashedPostState(
        accounts={a1: Account(1, 5), a2: None},
        storages={a3: HashedStorage(False, {bytes([9]) * 32: 7}), a2: HashedStorage(True, {})},
    )
    ps = st.construct_prefix_sets()
    assert ps.destroyed_accounts == {a2}
    frozen = ps.freeze()
    for a in (a1, a2, a3):
        assert frozen.account_prefix_set.contains(unpack_nibbles(a))
    assert frozen.storage_prefix_sets[a2].all
    assert frozen.storage_prefix_sets[a3].contains(unpack_nibbles(bytes([9]) * 32))
    assert not frozen.storage_prefix_sets[a3].contains(unpack_nibbles(bytes([8]) * 32))


def test_into_sorted_and_flat_layout():
    """hashed_state.rs:329-340 + the deletion rules of hashed_cursor/post_state.rs:260-297: destroyed accounts
    and zero-valued slots never reach the trie."""
    k = [bytes([i]) * 32 for i in (7, 3, 5, 1)]
    st = HashedPostState(
        accounts={k[0]: Account(1, 10), k[1]: Account(0, 2**200, b"\xaa" * 32), k[2]: None, k[3]: Account()},
        storages={k[1]: HashedStorage(False, {b"\x02" * 32: 5, b"\x01" * 32: 0, b"\x03" * 32: 2**255}),
                  k[2]: HashedStorage(True, {b"\x01" * 32: 1}),
                  b"\xee" * 32: HashedStorage(False, {b"\x01" * 32: 1})},  # storage without an account entry
    )
    s = st.into_sorted()
    assert [a for a, _ in s.accounts] == sorted(k)
    keys, accts, skeys, svals, offs = s.to_flat()
    assert [r.tobytes() for r in keys] == [k[3], k[1], k[0]]           # key order, destroyed account dropped
    assert list(offs) == [0, 0, 2, 2]                                   # only k[1] keeps (non-zero) slots
    assert [r.tobytes() for r in skeys] == [b"\x02" * 32, b"\x03" * 32]
    assert int.from_bytes(svals[1].tobytes(), "big") == 2**255
    assert accts[0]["code_hash"].tobytes() == KECCAK_EMPTY and accts[1]["code_hash"].tobytes() == b"\xaa" * 32
    assert int(accts[2]["nonce"]) == 1 and int.from_bytes(accts[1]["balance"].tobytes(), "big") == 2**200


def test_hashed_storage_extend_wipe():
    a = HashedStorage(False, {b"\x01" * 32: 1, b"\x02" * 32: 2})
    a.extend(HashedStorage(True, {b"\x03" * 32: 3}))
    assert a.wiped and a.storage == {b"\x03" * 32: 3}
    assert HashedStorage().is_empty() and not HashedStorage(True).is_empty()


def test_storage_trie_updates_semantics():
    """updates.rs:249-290: deleted() is not empty; insert_storage_updates drops empty updates (:126-137)."""
    assert not StorageTrieUpdates.deleted().is_empty() and len(StorageTrieUpdates.deleted()) == 1
    u = TrieUpdates()
    u.insert_storage_updates(b"\x01" * 32, StorageTrieUpdates())
    assert u.is_empty()
    u.insert_storage_updates(b"\x01" * 32, StorageTrieUpdates.deleted())
    assert list(u.storage_tries) == [b"\x01" * 32]


def test_empty_state_flat():
    keys, accts, skeys, svals, offs = HashedPostStateSorted().to_flat()
    assert keys.shape == (0, 32) and skeys.shape == (0, 32) and list(offs) == [0]


# ---------------------------------------------------------------- ordered roots: host side (no device needed)
def test_ordered_root_builder_bookkeeping_and_errors():
    """crates/trie/common/src/ordered_root.rs:315-353 — the parts that never reach the device"""
    import pytest
    from reth_b200 import EMPTY_ROOT_HASH, OrderedRootError, OrderedTrieRootEncodedBuilder
    b = OrderedTrieRootEncodedBuilder.new(None, 0)
    assert b.is_complete() and b.finalize() == EMPTY_ROOT_HASH          # :241-243, no engine involved
    b = OrderedTrieRootEncodedBuilder.new(None, 3)
    b.push(2, b"c")
    b.push(0, b"a")
    assert (b.pushed_count(), b.expected_count(), b.is_complete()) == (2, 3, False)
    with pytest.raises(OrderedRootError) as e:
        b.finalize()
    assert e.value.is_incomplete() and e.value.index() is None
    assert str(e.value) == "incomplete: expected 3 items, received 2"     # Display impl :66-80
    with pytest.raises(OrderedRootError) as e:
        b.push(3, b"x")
    assert e.value.is_index_out_of_bounds() and str(e.value) == "index 3 out of bounds for length 3"
    with pytest.raises(OrderedRootError) as e:
        b.push(2, b"again")
    assert e.value.is_duplicate_index() and str(e.value) == "duplicate item at index 2" and e.value.index() == 2


def test_ordered_root_pack_lists_and_rank_shares():
    from reth_b200.ordered_root import pack_lists
    from reth_b200.sharded import list_range_of
    values, vo, so = pack_lists([[b"ab", b""], [], [b"cde"]])
    assert values.tobytes() == b"abcde" and list(vo) == [0, 2, 2, 5] and list(so) == [0, 2, 2, 3]
    values, vo, so = pack_lists([])
    assert len(values) == 0 and list(vo) == [0] and list(so) == [0]
    for n in (0, 1, 7, 16, 1000):
        for world in (1, 2, 3, 8):
            shares = [list_range_of(r, world, n) for r in range(world)]
            assert shares[0][0] == 0 and shares[-1][1] == n
            assert all(shares[r][1] == shares[r + 1][0] for r in range(world - 1))
            assert max(hi - lo for lo, hi in shares) - min(hi - lo for lo, hi in shares) <= 1
