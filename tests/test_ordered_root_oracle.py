"""CPU: the ordered-root oracle (oracle/ordered_root.c) against the reference's golden roots
(crates/ethereum/primitives/src/receipt.rs:180-245, committed as tests/golden/ordered_roots.json by
tests/golden/make_ordered_roots.py) and against an independent recursive trie for the list sizes the reference's
own equivalence test sweeps (crates/trie/common/src/ordered_root.rs:264-283)."""
import json
import os

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def golden_cases():
    with open(os.path.join(HERE, "golden", "ordered_roots.json")) as f:
        return json.load(f)["cases"]


# ---------------------------------------------------------------- an independent implementation (yellow paper, recursive)
def _rlp_bytes(x: bytes) -> bytes:
    if len(x) == 1 and x[0] < 0x80:
        return x
    if len(x) < 56:
        return bytes([0x80 + len(x)]) + x
    ll = (len(x).bit_length() + 7) // 8
    return bytes([0xB7 + ll]) + len(x).to_bytes(ll, "big") + x


def _rlp_list(body: bytes) -> bytes:
    if len(body) < 56:
        return bytes([0xC0 + len(body)]) + body
    ll = (len(body).bit_length() + 7) // 8
    return bytes([0xF7 + ll]) + len(body).to_bytes(ll, "big") + body


def _hp(nibs, leaf: bool) -> bytes:
    flag = 2 if leaf else 0
    if len(nibs) % 2:
        out = [((flag | 1) << 4) | nibs[0]]
        nibs = nibs[1:]
    else:
        out = [flag << 4]
    out += [(nibs[i] << 4) | nibs[i + 1] for i in range(0, len(nibs), 2)]
    return bytes(out)


def _node(pairs, depth) -> bytes:
    """RLP of the node over (nibble tuple, value) pairs that share their first `depth` nibbles."""
    if len(pairs) == 1:
        k, v = pairs[0]
        return _rlp_list(_rlp_bytes(_hp(list(k[depth:]), True)) + _rlp_bytes(v))
    first = pairs[0][0]
    lcp = min(len(k) for k, _ in pairs)
    for k, _ in pairs:
        j = depth
        while j < lcp and k[j] == first[j]:
            j += 1
        lcp = j
    if lcp > depth:
        return _rlp_list(_rlp_bytes(_hp(list(first[depth:lcp]), False)) + _ref(_node(pairs, lcp)))
    body = b""
    for nib in range(16):
        sub = [(k, v) for k, v in pairs if len(k) > depth and k[depth] == nib]
        body += _ref(_node(sub, depth + 1)) if sub else b"\x80"
    assert all(len(k) > depth for k, _ in pairs)  # rlp(index) keys are prefix-free: no value in a branch
    return _rlp_list(body + b"\x80")


def _ref(rlp: bytes) -> bytes:
    return rlp if len(rlp) < 32 else _rlp_bytes(oracle.keccak256(rlp))


def independent_ordered_root(items) -> bytes:
    if not items:
        return oracle.EMPTY_ROOT_HASH
    pairs = []
    for i, it in enumerate(items):
        key = _rlp_bytes(i.to_bytes((i.bit_length() + 7) // 8, "big"))
        pairs.append((tuple(x for b in key for x in (b >> 4, b & 15)), bytes(it)))
    pairs.sort()
    return oracle.keccak256(_node(pairs, 0))


def oracle_root(items) -> bytes:
    return oracle.ordered_roots(*oracle.pack_lists([items]))[0].tobytes()


# ---------------------------------------------------------------- tests
@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_oracle_matches_reference_golden_roots(case):
    items = [bytes.fromhex(x) for x in case["items"]]
    assert oracle_root(items).hex() == case["root"]
    assert independent_ordered_root(items).hex() == case["root"]


@pytest.mark.parametrize("n", [0, 1, 2, 3, 10, 127, 128, 129, 130, 200, 256, 257, 700])
def test_oracle_matches_independent_trie(n):
    """list sizes of test_ordered_encoded_builder_equivalence (:266), items as there, plus the 2-byte index range"""
    items = [f"item_{i}_data".encode() for i in range(n)]
    assert oracle_root(items) == independent_ordered_root(items)


def test_oracle_item_shapes():
    rng = np.random.default_rng(7)
    shapes = [0, 1, 2, 30, 31, 32, 33, 54, 55, 56, 57, 135, 136, 137, 255, 256, 271, 272, 273, 1000, 70000]
    for n in (1, 2, 5, 130):
        items = [rng.integers(0, 256, shapes[int(rng.integers(0, len(shapes)))], dtype=np.uint8).tobytes()
                 for _ in range(n)]
        items[0] = b"\x05"          # single byte below 0x80: its own RLP
        if n > 1:
            items[1] = b"\x80"      # single byte at 0x80: needs a string header
        assert oracle_root(items) == independent_ordered_root(items)


def test_oracle_batch_equals_single_lists():
    lists = [[f"l{l}_{i}".encode() * (1 + (i % 7)) for i in range(n)] for l, n in enumerate([3, 0, 1, 129, 0, 40])]
    roots = oracle.ordered_roots(*oracle.pack_lists(lists))
    for l, items in enumerate(lists):
        assert roots[l].tobytes() == oracle_root(items)
    assert roots[1].tobytes() == oracle.EMPTY_ROOT_HASH
