"""GPU parity of b200_ordered_roots (transactions / receipts / withdrawals roots) through the C ABI:
the reference's golden roots (crates/ethereum/primitives/src/receipt.rs:180-245), the reference's own builder tests
(crates/trie/common/src/ordered_root.rs:263-353) through the host mirror, and the CPU oracle on random batches.
Bit-exact.  (Named to run last: first validated under tools/emu)"""
import numpy as np
import pytest

# first hardware run happens at round end: bound a hang (method=thread ends the process even inside a blocked CUDA sync)
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]

import oracle
from tests.test_ordered_root_oracle import golden_cases
from tests.util import read_device, to_device_ptrs


@pytest.fixture(scope="module")
def eng():
    from reth_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_reference_golden_roots(eng, case):
    items = [bytes.fromhex(x) for x in case["items"]]
    assert eng.ordered_root(items).hex() == case["root"]


def test_ordered_encoded_builder_equivalence(eng):
    """ordered_root.rs:264-283"""
    from reth_b200 import OrderedTrieRootEncodedBuilder
    for n in [0, 1, 2, 3, 10, 127, 128, 129, 130, 200]:
        items = [f"item_{i}_data".encode() for i in range(n)]
        expected = oracle.ordered_roots(*oracle.pack_lists([items]))[0].tobytes()
        b = OrderedTrieRootEncodedBuilder.new(eng, n)
        for i, it in enumerate(items):
            b.push(i, it)
        assert b.finalize() == expected, n


def test_ordered_builder_out_of_order(eng):
    """ordered_root.rs:286-312"""
    from reth_b200 import OrderedTrieRootEncodedBuilder
    for n in [2, 3, 5, 10, 50]:
        items = [f"item_{i}_data".encode() for i in range(n)]
        expected = oracle.ordered_roots(*oracle.pack_lists([items]))[0].tobytes()
        b = OrderedTrieRootEncodedBuilder.new(eng, n)
        for i in reversed(range(n)):
            b.push(i, items[i])
        assert b.finalize() == expected
        b = OrderedTrieRootEncodedBuilder.new(eng, n)
        for i in list(range(1, n, 2)) + list(range(0, n, 2)):
            b.push(i, items[i])
        assert b.finalize() == expected


def test_ordered_builder_empty_incomplete_and_index_errors(eng):
    """ordered_root.rs:315-353"""
    from reth_b200 import EMPTY_ROOT_HASH, OrderedRootError, OrderedTrieRootEncodedBuilder
    b = OrderedTrieRootEncodedBuilder.new(eng, 0)
    assert b.is_complete() and b.finalize() == EMPTY_ROOT_HASH
    b = OrderedTrieRootEncodedBuilder.new(eng, 3)
    b.push(0, b"item_0")
    b.push(1, b"item_1")
    assert not b.is_complete()
    with pytest.raises(OrderedRootError) as e:
        b.finalize()
    assert e.value == OrderedRootError("Incomplete", expected=3, received=2) and e.value.is_incomplete()
    b = OrderedTrieRootEncodedBuilder.new(eng, 2)
    with pytest.raises(OrderedRootError) as e:
        b.push(5, b"item")
    assert e.value == OrderedRootError("IndexOutOfBounds", index=5, len=2) and e.value.index() == 5
    b.push(0, b"item_0")
    with pytest.raises(OrderedRootError) as e:
        b.push(0, b"item_0_dup")
    assert e.value == OrderedRootError("DuplicateIndex", index=0)
    b.push(1, b"item_1")
    assert b.is_complete() and b.pushed_count() == 2 and b.expected_count() == 2


def _random_items(rng, n, shapes):
    return [rng.integers(0, 256, int(shapes[int(rng.integers(0, len(shapes)))]), dtype=np.uint8).tobytes()
            for _ in range(n)]


def test_item_shapes_and_block_boundaries(eng):
    """values around the RLP header switches (1, 55/56, 255/256 bytes), the inline/hashed switch (leaf RLP of 32 bytes)
    and the keccak block boundaries (135/136/137, 271/272/273), at every byte alignment inside the blob"""
    rng = np.random.default_rng(11)
    shapes = [0, 1, 2, 3, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 54, 55, 56, 57, 120, 128, 129, 130, 131, 132, 133, 134,
              135, 136, 137, 255, 256, 257, 262, 263, 264, 265, 271, 272, 273, 407, 408, 1000, 4099]
    lists = []
    for n in (1, 2, 3, 17, 127, 128, 129, 300):
        items = _random_items(rng, n, shapes)
        items[0] = b"\x05"
        if n > 1:
            items[1] = b"\x80"
        lists.append(items)
    lists.append([bytes([b]) for b in range(0x70, 0x90)])  # single bytes either side of 0x80
    packed = oracle.pack_lists(lists)
    assert (eng.ordered_roots(*packed) == oracle.ordered_roots(*packed)).all()


def test_batches_with_empty_lists_and_stats(eng):
    rng = np.random.default_rng(12)
    sizes = [0, 0, 1, 0, 5, 200, 0, 128, 1, 1, 0, 77, 129, 0]
    lists = [_random_items(rng, n, [8, 40, 110, 300, 600]) for n in sizes]
    packed = oracle.pack_lists(lists)
    roots, stats = eng.ordered_roots(*packed, want_stats=True)
    want = oracle.ordered_roots(*packed)
    assert (roots == want).all()
    assert roots[0].tobytes() == oracle.EMPTY_ROOT_HASH
    assert stats["leaves_added"] == sum(sizes)
    # nothing at all
    assert eng.ordered_roots(np.zeros(0, np.uint8), np.zeros(1, np.uint64), np.zeros(1, np.uint64)).shape == (0, 32)
    empty_only = oracle.pack_lists([[], []])
    assert (eng.ordered_roots(*empty_only) == oracle.ordered_roots(*empty_only)).all()


def test_three_byte_index_keys(eng):
    """more than 65535 items: keys 0x83 ‖ 3 bytes; tiny items keep every leaf inline below the top levels"""
    n = 66000
    items = [(i * 2654435761 % 251).to_bytes(1, "big") * (1 + i % 3) for i in range(n)]
    packed = oracle.pack_lists([items, items[:300]])
    assert (eng.ordered_roots(*packed) == oracle.ordered_roots(*packed)).all()


def test_receipt_and_transaction_shaped_batch(eng):
    """a batch of blocks: receipts (>= 261 bytes with the bloom) and transactions (110 B .. 20 KB of calldata)"""
    rng = np.random.default_rng(13)
    lists = []
    for blk in range(24):
        n = int(rng.integers(0, 260))
        receipts = [b"\x02" + rng.integers(0, 256, 261 + int(rng.integers(0, 400)), dtype=np.uint8).tobytes()
                    for _ in range(n)]
        txs = [rng.integers(0, 256, int(rng.choice([110, 115, 180, 700, 3000, 20000], p=[.4, .2, .2, .1, .07, .03])),
                            dtype=np.uint8).tobytes() for _ in range(n)]
        lists += [txs, receipts]
    packed = oracle.pack_lists(lists)
    assert (eng.ordered_roots(*packed) == oracle.ordered_roots(*packed)).all()


def test_long_items_take_the_warp_path(eng):
    """items of >= 32 rate blocks (≈4.3 KB) are hashed by a warp each: lengths either side of that switch, block
    multiples, and a maximum-size (128 KiB) calldata transaction, mixed with short items in the same lists"""
    rng = np.random.default_rng(15)
    lens = [4300, 4334, 4335, 4336, 4351, 4352, 4353, 8192, 13600 - 9, 13600, 40000, 131072]
    big = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]
    small = _random_items(rng, 150, [1, 40, 110, 300])
    lists = [big[:4] + small[:50], small[50:100], [big[11]], small[100:] + big[4:]]
    packed = oracle.pack_lists(lists)
    assert (eng.ordered_roots(*packed) == oracle.ordered_roots(*packed)).all()
    for skew in (1, 5):
        assert (_dev_roots(eng, *packed, skew=skew) == oracle.ordered_roots(*packed)).all()


def test_bad_offsets_are_rejected(eng):
    from reth_b200 import B200Error
    vals = np.zeros(10, np.uint8)
    with pytest.raises((B200Error, ValueError)):
        eng.ordered_roots(vals, np.array([0, 6, 4], np.uint64), np.array([0, 2], np.uint64))   # not monotone
    with pytest.raises((B200Error, ValueError)):
        eng.ordered_roots(vals, np.array([0, 4, 40], np.uint64), np.array([0, 2], np.uint64))  # past the blob
    with pytest.raises((B200Error, ValueError)):
        eng.ordered_roots(vals, np.array([0, 4, 8], np.uint64), np.array([0, 3], np.uint64))   # more items than offsets
    ok = oracle.pack_lists([[b"ab", b"cd"]])
    assert (eng.ordered_roots(*ok) == oracle.ordered_roots(*ok)).all()  # the context stays usable


def _dev_roots(eng, values, value_offsets, seg_offsets, skew=0):
    """b200_ordered_roots_dev with every buffer in device memory; `skew` shifts the items off 8-byte alignment"""
    m, n = len(seg_offsets) - 1, len(value_offsets) - 1
    blob = np.concatenate([np.zeros(skew, np.uint8), values, np.zeros(8, np.uint8)])
    roots = np.zeros((max(m, 1), 32), np.uint8)
    (p_blob, p_vo, p_so, p_roots), hold = to_device_ptrs([blob, value_offsets, seg_offsets, roots])
    eng._check(eng.lib.b200_ordered_roots_dev(eng.ctx, p_blob + skew, len(values), p_vo, p_so, m, n, p_roots))
    eng.sync()
    return read_device(hold[3]).reshape(-1, 32)[:m]


def test_device_resident_variant_any_alignment(eng):
    rng = np.random.default_rng(14)
    lists = [_random_items(rng, n, [1, 20, 33, 136, 137, 300, 1111]) for n in (0, 1, 40, 130, 3)]
    packed = oracle.pack_lists(lists)
    want = oracle.ordered_roots(*packed)
    for skew in (0, 1, 3, 4, 7):
        assert (_dev_roots(eng, *packed, skew=skew) == want).all(), skew


def test_device_side_offset_validation(eng):
    """the asynchronous variant reports violations through the sticky status (b200_sync), and the context recovers"""
    from reth_b200 import B200Error
    ok = oracle.pack_lists([[b"abcd" * 20, b"ef" * 50, b"g"]])
    values, vo, so = ok
    for bad_vo, bad_so in [(np.array([0, 90, 80, 181], np.uint64), so),          # value offsets not monotone
                           (np.array([0, 80, 180, 9999], np.uint64), so),         # past the blob
                           (vo, np.array([0, 5], np.uint64)),                     # list longer than the items
                           (vo, np.array([2, 3], np.uint64))]:                    # does not start at 0
        with pytest.raises(B200Error):
            _dev_roots(eng, values, bad_vo, bad_so)
    with pytest.raises(B200Error):   # items but no list: refused on the host, before any launch
        _dev_roots(eng, values, vo, np.array([0], np.uint64))
    assert (_dev_roots(eng, *ok) == oracle.ordered_roots(*ok)).all()
