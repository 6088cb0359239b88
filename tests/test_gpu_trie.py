"""GPU parity: storage roots / state roots / TrieUpdates through the C ABI against the reference's golden
vectors and the CPU oracle.  Bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from tests.test_oracle_golden import EXT_KEYS, _check_ext_updates, _flat_from_hashed, _vector5, ETHER
from tests.util import alloc_to_flat, random_keys, sort_rows, synth_accounts, synth_storage, u256_be

H = bytes.fromhex


@pytest.fixture(scope="module")
def eng():
    from reth_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


# ---------------------------------------------------------------- golden vectors through the C ABI
@pytest.mark.parametrize("chain", ["mainnet", "sepolia", "holesky", "goerli"])
def test_genesis_state_roots(eng, golden_allocs, chain):
    """stateRoot of crates/chainspec/res/genesis/<chain>.json; addresses and slots hashed on the GPU too."""
    g = golden_allocs[chain]
    flat = alloc_to_flat(g["alloc"], keccak_rows=eng.keccak256_fixed)
    assert eng.state_root_full(*flat).hex() == g["state_root"]


def test_testspec_root(eng, golden_allocs):
    """crates/trie/db/tests/proof.rs:55 — root = keccak(extension node e48200a7a0...)."""
    flat = alloc_to_flat(golden_allocs["testspec"]["alloc"], keccak_rows=eng.keccak256_fixed)
    ext = H("e48200a7a040f916999be583c572cc4dd369ec53b0a99f7de95f13880cf203d98f935ed1b3")
    assert eng.state_root_full(*flat) == oracle.keccak256(ext)


def test_account_and_storage_trie(eng):
    """crates/trie/db/tests/trie.rs:357-477."""
    flat = _flat_from_hashed(_vector5())
    root, au, su = eng.state_root_full(*flat, want_updates=True)
    assert root.hex() == "72861041bc90cd2f93777956f058a545412b56de79af5eb6b8075fe2eabbe015"
    assert [(list(p), s, t, h, len(hs)) for (_, p, s, t, h, hs) in au] == [
        ([0xB], 0b1011, 0b0001, 0b1001, 2), ([0xB, 0], 0b10001, 0, 0b10000, 1)]
    o_root, o_au, o_su = oracle.state_root_full(*flat, want_updates=True)
    assert (root, au, su) == (o_root, o_au, o_su)


def test_account_and_storage_trie_after_insert(eng):
    """crates/trie/db/tests/trie.rs:479-522 (root of the full rebuild)."""
    accounts = _vector5()
    accounts.append((oracle.keccak256(H("4f61f2d5ebd991b85aa1677db97307caf5215c91")), 0, 5 * ETHER, None, []))
    root, au, _ = eng.state_root_full(*_flat_from_hashed(accounts), want_updates=True)
    assert root.hex() == "8e263cd4eefb0c3cbbb14e5541a66a755cad25bcfab1e10dd9d706263e811b28"
    assert [(list(p), s, t, h, len(hs)) for (_, p, s, t, h, hs) in au] == [
        ([0xB], 0b1011, 0b0001, 0b1011, 3), ([0xB, 0], 0b10001, 0, 0b10000, 1)]


def test_known_bundle_root(eng):
    """crates/trie/db/src/state.rs:408-437."""
    alloc = {
        "00" * 19 + "01": {"nonce": "0x1", "balance": "0x0", "storage": {"0x" + (1015).to_bytes(32, "big").hex(): hex(10)}},
        "00" * 19 + "02": {"nonce": "0x2", "balance": "0x0", "storage": {"0x" + (2015).to_bytes(32, "big").hex(): hex(20)}},
    }
    flat = alloc_to_flat(alloc, keccak_rows=eng.keccak256_fixed)
    assert eng.state_root_full(*flat).hex() == "b464525710cafcf5d4044ac85b72c08b1e76231b8d91f288fe438cc41d8eaafd"


def test_extension_node_tries(eng):
    """crates/trie/db/tests/trie.rs:719-805: stored nodes [3] and [3,0,A,F]."""
    keys = np.frombuffer(b"".join(H(k) for k in EXT_KEYS), np.uint8).reshape(-1, 32)
    roots, upd = eng.storage_roots(keys, u256_be([1] * 6), [0, 6], want_updates=True)
    _check_ext_updates(upd)
    o_roots, o_upd = oracle.storage_roots(keys, u256_be([1] * 6), [0, 6], want_updates=True)
    assert (roots == o_roots).all() and upd == o_upd
    accs = oracle.make_accounts([(0, 1, bytes(range(32)))] * 6)
    root, upd = eng.state_root(keys, accs, want_updates=True)
    _check_ext_updates(upd)
    assert (root, upd) == oracle.state_root(keys, accs, want_updates=True)


# ---------------------------------------------------------------- edge cases
def test_empty_and_single(eng):
    assert eng.state_root(np.zeros((0, 32), np.uint8), np.zeros(0, oracle.ACCOUNT_DTYPE)) == oracle.EMPTY_ROOT_HASH
    roots = eng.storage_roots(np.zeros((0, 32), np.uint8), np.zeros((0, 32), np.uint8), [0, 0, 0])
    assert all(r.tobytes() == oracle.EMPTY_ROOT_HASH for r in roots)
    k, a = synth_accounts(1, 1)
    assert eng.state_root(k, a) == oracle.state_root(k, a)
    sk, sv, so = synth_storage(2, [1])
    assert (eng.storage_roots(sk, sv, so) == oracle.storage_roots(sk, sv, so)).all()
    k, a = synth_accounts(3, 2)
    assert eng.state_root(k, a, want_updates=True) == oracle.state_root(k, a, want_updates=True)


def _shared_prefix_keys(rng, n, max_share=63):
    keys = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    for i in range(1, n):
        if rng.random() < 0.6:
            share = int(rng.integers(0, max_share + 1))
            src = keys[int(rng.integers(0, i))]
            nb = share // 2
            keys[i, :nb] = src[:nb]
            if share & 1:
                keys[i, nb] = (src[nb] & 0xF0) | (keys[i, nb] & 0x0F)
    keys = np.unique(keys, axis=0)
    return keys[sort_rows(keys)]


@pytest.mark.parametrize("seed", range(6))
def test_forest_with_deep_shared_prefixes(eng, seed):
    """Many small tries whose keys share up to 63 nibbles: every level 0..63 populated, extension nodes at all
    depths, inline (<32 byte) leaves and branches.  Roots must equal the oracle's.  (Updates are not retained
    here: an inline branch child under a hash bit makes alloy-trie itself panic.)"""
    rng = np.random.default_rng(100 + seed)
    segs, vals = [], []
    for _ in range(60):
        n = int(rng.choice([1, 2, 3, 5, 8, 17, 40]))
        k = _shared_prefix_keys(rng, n)
        segs.append(k)
        v = np.zeros((len(k), 32), np.uint8)
        for i in range(len(k)):
            if rng.random() < 0.5:
                v[i, 31] = rng.integers(1, 0x80)  # tiny values -> inline leaves when the suffix is short
            else:
                ln = int(rng.integers(1, 33))
                v[i, 32 - ln:] = rng.integers(0, 256, ln, dtype=np.uint8)
                v[i, 32 - ln] |= 1
        vals.append(v)
    offs = np.zeros(len(segs) + 1, np.uint64)
    offs[1:] = np.cumsum([len(s) for s in segs])
    keys, values = np.concatenate(segs), np.concatenate(vals)
    got = eng.storage_roots(keys, values, offs)
    exp = oracle.storage_roots(keys, values, offs, threads=2)
    assert (got == exp).all()


def test_node_heads_across_trie_boundaries_and_long_spans(eng):
    """Which gaps start a branch node (tk_structure.cuh): most learn it from the 32 gaps to their left (gap_keys_kernel), the
    others after the sort from the keys (head_fix_kernel) — which must not join two tries.  Trie A ends in a large top-nibble-7
    group, trie B starts with one: the previous depth-0 gap of B's first depth-0 gap lies in A, more than 32 gaps away, and
    the leaves between them all share nibble 7.  Then: tries whose spans straddle exactly 31 / 32 / 33 / 34 gaps; empty and
    one-leaf tries in between."""
    rng = np.random.default_rng(77)

    def keys_with_top(nibbles, n):
        k = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        k[:, 0] = (rng.choice(nibbles, n).astype(np.uint8) << 4) | (k[:, 0] & 0x0F)
        k = np.unique(k, axis=0)
        return k[sort_rows(k)]

    def span_trie(span):
        # two depth-0 gaps with `span` leaves of one top nibble between them (their gaps are all deeper)
        k = np.concatenate([keys_with_top([1], 1), keys_with_top([2], span), keys_with_top([3], 1)])
        return k[sort_rows(k)]

    segs = [keys_with_top(list(range(0, 8)), 3000), keys_with_top(list(range(7, 16)), 1500), np.zeros((0, 32), np.uint8),
            keys_with_top([4], 1)]
    segs += [span_trie(sp) for sp in (30, 31, 32, 33, 34, 35, 64, 65)]
    segs += [keys_with_top([7], 700), keys_with_top([7, 8], 900)]
    offs = np.zeros(len(segs) + 1, np.uint64)
    offs[1:] = np.cumsum([len(x) for x in segs])
    keys = np.concatenate(segs)
    vals = u256_be(rng.integers(1, 1 << 62, len(keys)))
    roots, upd = eng.storage_roots(keys, vals, offs, want_updates=True)
    o_roots, o_upd = oracle.storage_roots(keys, vals, offs, want_updates=True, threads=4)
    assert (roots == o_roots).all()
    assert upd == o_upd


@pytest.mark.parametrize("seed", range(3))
def test_node_heads_at_block_boundaries_of_the_gap_pass(eng, seed):
    """gap_keys_kernel works on 256 gaps per CTA behind a 32-gap window of the previous block: tries whose sizes put trie
    boundaries, first gaps and long same-nibble spans on and around multiples of 256 (and of 32), random and clustered keys."""
    rng = np.random.default_rng(500 + seed)
    sizes = [255, 1, 256, 257, 31, 33, 32, 512, 513, 2, 254, 0, 258, 1023, 1, 1025, 224, 288, 16, 240]
    rng.shuffle(sizes)
    segs = []
    for n in sizes:
        k = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        if n and rng.random() < 0.5:                      # clustered: few top nibbles, long runs of deeper gaps
            k[:, 0] = (rng.choice([3, 3, 3, 9], n).astype(np.uint8) << 4) | (k[:, 0] & 0x0F)
            k[: n // 2, 1] = 0x55
        k = np.unique(k, axis=0)
        segs.append(k[sort_rows(k)] if len(k) else k)
    offs = np.zeros(len(segs) + 1, np.uint64)
    offs[1:] = np.cumsum([len(x) for x in segs])
    keys = np.concatenate(segs)
    vals = u256_be(rng.integers(1, 1 << 62, len(keys)))
    roots, upd = eng.storage_roots(keys, vals, offs, want_updates=True)
    o_roots, o_upd = oracle.storage_roots(keys, vals, offs, want_updates=True, threads=4)
    assert (roots == o_roots).all()
    assert upd == o_upd


@pytest.mark.parametrize("mode", ["u64", "mixed"])
def test_storage_forest_random(eng, mode):
    rng = np.random.default_rng(4)
    counts = rng.choice([0, 1, 2, 3, 16, 17, 100, 1000], 3000, p=[.2, .2, .1, .1, .2, .1, .09, .01])
    keys, vals, offs = synth_storage(31, counts, value_mode=mode)
    roots, upd, stats = eng.storage_roots(keys, vals, offs, want_updates=True, want_stats=True)
    o_roots, o_upd = oracle.storage_roots(keys, vals, offs, want_updates=True, threads=4)
    assert (roots == o_roots).all()
    assert upd == o_upd
    assert stats["leaves_added"] == len(keys)


def test_account_trie_random_with_updates(eng):
    keys, accs = synth_accounts(8, 200_000)
    sroots = random_keys(99, len(keys))
    root, upd, stats = eng.state_root(keys, accs, sroots, want_updates=True, want_stats=True)
    o_root, o_upd = oracle.state_root(keys, accs, sroots, want_updates=True)
    assert root == o_root
    assert upd == o_upd
    oracle.stats_reset()
    oracle.state_root(keys, accs, sroots)
    s = oracle.stats()
    assert stats["branches_added"] == s["branch_nodes"]
    assert stats["extension_nodes"] == s["extension_nodes"]
    assert stats["hashed_nodes"] == s["hashed_nodes"]


def test_c1_config(eng):
    """BASELINE.json configs[0]: 10k synthetic accounts, no storage."""
    n = 10_000
    seed_msgs = np.zeros((n, 16), np.uint8)
    seed_msgs[:, 0] = 1
    seed_msgs[:, 8:16] = np.arange(n, dtype="<u8").view(np.uint8).reshape(n, 8)
    addrs = oracle.keccak256_fixed(seed_msgs)[:, :20].copy()
    hashed = eng.keccak256_fixed(addrs)
    assert (hashed == oracle.keccak256_fixed(addrs)).all()
    order = sort_rows(hashed)
    accs = oracle.make_accounts([(int(i) % 7, (int(i) + 1) * 10**15, None) for i in order])
    root = eng.state_root(hashed[order], accs)
    assert root == oracle.state_root(hashed[order], accs)


def test_stats_count_the_keccak_permutations(eng):
    """b200_stats.keccak_f (the numerator of bench.py's alu_frac) == the permutations the oracle's HashBuilder executes for the
    same state: digests plus the extra rate blocks of the 4..16-child branch nodes; hashed_nodes == the oracle's digests."""
    n = 20_000
    akeys, accs = synth_accounts(52, n)
    counts = np.where(np.arange(n) % 3 == 0, 16, 0) + np.where(np.arange(n) % 499 == 0, 700, 0)
    skeys, svals, offs = synth_storage(53, counts, value_mode="u64")
    oracle.stats_reset()
    o_root = oracle.state_root_full(akeys, accs, skeys, svals, offs)
    want = oracle.stats()
    root, st = eng.state_root_full(akeys, accs, skeys, svals, offs, want_stats=True)
    assert root == o_root
    assert st["hashed_nodes"] == want["hashed_nodes"]
    assert st["keccak_f"] == want["keccak_f"] > st["hashed_nodes"]
    # not counted where the child-count classes do not apply
    assert eng.ordered_roots(*pack_lists_for_stats(), want_stats=True)[1]["keccak_f"] == 0


def pack_lists_for_stats():
    items = [bytes([i]) * (40 + 7 * i) for i in range(20)]
    values = np.frombuffer(b"".join(items), np.uint8)
    vo = np.zeros(len(items) + 1, np.uint64)
    vo[1:] = np.cumsum([len(x) for x in items])
    return values, vo, np.array([0, len(items)], np.uint64)


def test_full_state_random(eng):
    n = 30_000
    akeys, accs = synth_accounts(12, n)
    counts = np.where(np.arange(n) % 5 == 0, 16, 0) + np.where(np.arange(n) % 997 == 0, 3000, 0)
    skeys, svals, offs = synth_storage(13, counts, value_mode="mixed")
    root, au, su = eng.state_root_full(akeys, accs, skeys, svals, offs, want_updates=True)
    o_root, o_au, o_su = oracle.state_root_full(akeys, accs, skeys, svals, offs, want_updates=True, threads=4)
    assert root == o_root
    assert au == o_au
    assert su == o_su


@pytest.mark.parametrize("packed", [0, 1])
def test_full_state_table_rows(eng, packed):
    """the stored nodes of a device build as AccountsTrie / StoragesTrie rows, against rows encoded
    from the oracle's TrieUpdates by the test-side restatement of the reference codecs."""
    from tests.test_table_rows import expected_account_rows, expected_storage_rows
    n = 20_000
    akeys, accs = synth_accounts(21, n)
    counts = np.where(np.arange(n) % 4 == 0, 24, 0) + np.where(np.arange(n) % 1999 == 0, 2500, 0)
    skeys, svals, offs = synth_storage(22, counts, value_mode="mixed")
    root, arows, srows = eng.state_root_full_rows(akeys, accs, skeys, svals, offs, key_format=packed, encode_on_host=True)
    o_root, o_au, o_su = oracle.state_root_full(akeys, accs, skeys, svals, offs, want_updates=True, threads=4)
    assert root == o_root
    assert arows.to_list() == expected_account_rows(o_au, bool(packed))
    assert srows.to_list() == expected_storage_rows(o_su, akeys, bool(packed))
    assert len(arows) > 1000 and len(srows) > 1000
    arows.release(); srows.release()


def test_update_records_leave_the_device_in_table_order(eng):
    """full builds: b200_updates records are already sorted by (trie id, path) — pre-order, the key order of the trie
    tables — with hashes following their records; the row encoder then has nothing to sort"""
    import ctypes as C
    from reth_b200._lib import Updates
    from reth_b200.engine import updates_to_records
    n = 6000
    akeys, accs = synth_accounts(31, n)
    counts = np.where(np.arange(n) % 3 == 0, 40, 0) + np.where(np.arange(n) % 997 == 0, 1500, 0)
    skeys, svals, offs = synth_storage(32, counts, value_mode="mixed")
    root = np.empty(32, np.uint8)
    au, su = Updates(), Updates()
    eng._check(eng.lib.b200_state_root_full(eng.ctx, akeys.ctypes.data, accs.ctypes.data, n, skeys.ctypes.data,
                                            svals.ctypes.data, offs.ctypes.data, root.ctypes.data, C.byref(au),
                                            C.byref(su), None))
    for u in (au, su):
        ho = np.ctypeslib.as_array(u.hash_offset, (int(u.n_nodes) + 1,)).copy()
        assert (np.diff(ho.astype(np.int64)) >= 0).all() and ho[0] == 0
        recs = updates_to_records(u, eng.lib, sort=False)
        assert len(recs) > 300
        assert recs == sorted(recs, key=lambda r: (r[0], r[1]))
    o_root, o_au, o_su = oracle.state_root_full(akeys, accs, skeys, svals, offs, want_updates=True, threads=4)
    assert root.tobytes() == o_root


def test_one_million_leaves(eng):
    """Size-independent check at scale: 1M-account trie root equals the oracle's."""
    keys, accs = synth_accounts(77, 1_000_000)
    root, stats = eng.state_root(keys, accs, want_stats=True)
    assert root == oracle.state_root(keys, accs)
    assert stats["leaves_added"] == 1_000_000


# ---------------------------------------------------------------- error behaviour
def test_rejects_unsorted_and_zero(eng):
    from reth_b200 import B200Error
    from reth_b200 import _lib
    keys, accs = synth_accounts(5, 100)
    bad = keys.copy()
    bad[[10, 11]] = bad[[11, 10]]
    with pytest.raises(B200Error) as e:
        eng.state_root(bad, accs)
    assert e.value.status == _lib.ERR_UNSORTED
    dup = keys.copy()
    dup[11] = dup[10]
    with pytest.raises(B200Error) as e:
        eng.state_root(dup, accs)
    assert e.value.status == _lib.ERR_UNSORTED
    sk, sv, so = synth_storage(6, [4, 4])
    sv[5] = 0
    with pytest.raises(B200Error) as e:
        eng.storage_roots(sk, sv, so)
    assert e.value.status == _lib.ERR_ZERO_VALUE
    with pytest.raises(B200Error) as e:
        eng.storage_roots(sk, np.ones_like(sv), np.array([0, 5, 3], np.uint64))
    assert e.value.status == _lib.ERR_INVALID_ARG
    # the context stays usable after an error
    assert eng.state_root(keys, accs) == oracle.state_root(keys, accs)


# ---------------------------------------------------------------- multi-GPU frontier (emulated shards on one GPU)
@pytest.mark.parametrize("world", [1, 2, 4, 8, 16])
def test_frontier_sharding_matches_full_root(eng, world):
    n = 20_000
    akeys, accs = synth_accounts(41, n)
    counts = np.where(np.arange(n) % 11 == 0, 5, 0)
    skeys, svals, offs = synth_storage(42, counts)
    full = oracle.state_root_full(akeys, accs, skeys, svals, offs, threads=4)
    merged = np.zeros((16, 68), np.uint8)
    top = akeys[:, 0] >> 4
    for rank in range(world):
        lo_n, hi_n = rank * 16 // world, (rank + 1) * 16 // world
        sel = np.nonzero((top >= lo_n) & (top < hi_n))[0]
        a0, a1 = (int(sel[0]), int(sel[-1]) + 1) if len(sel) else (0, 0)
        s0, s1 = int(offs[a0]), int(offs[a1])
        fr = eng.subtrie_frontier(akeys[a0:a1], accs[a0:a1], skeys[s0:s1], svals[s0:s1], offs[a0:a1 + 1] - offs[a0])
        for b in range(lo_n, hi_n):
            merged[b] = fr[b]
        others = [b for b in range(16) if not lo_n <= b < hi_n]
        assert not fr[others].any()
    assert eng.root_from_frontier(merged) == full


def test_frontier_degenerate_shapes(eng):
    # all accounts in one top-nibble bucket: the root is not a depth-0 branch
    keys, accs = synth_accounts(51, 500)
    keys[:, 0] = (keys[:, 0] & 0x0F) | 0x70
    keys = keys[sort_rows(keys)]
    z = np.zeros((0, 32), np.uint8)
    offs = np.zeros(len(keys) + 1, np.uint64)
    fr = eng.subtrie_frontier(keys, accs, z, z, offs)
    assert eng.root_from_frontier(fr) == oracle.state_root(keys, accs)
    # a single account; and nothing at all
    fr = eng.subtrie_frontier(keys[:1], accs[:1], z, z, offs[:2])
    assert eng.root_from_frontier(fr) == oracle.state_root(keys[:1], accs[:1])
    fr = eng.subtrie_frontier(keys[:0], accs[:0], z, z, offs[:1])
    assert eng.root_from_frontier(fr) == oracle.EMPTY_ROOT_HASH
    # two buckets with one account each
    k2 = np.zeros((2, 32), np.uint8)
    k2[0, 0], k2[1, 0] = 0x10, 0xF0
    fr = eng.subtrie_frontier(k2, accs[:2], z, z, offs[:3])
    assert eng.root_from_frontier(fr) == oracle.state_root(k2, accs[:2])


def test_device_resident_full_state(eng):
    import torch
    n = 50_000
    akeys, accs = synth_accounts(61, n)
    skeys, svals, offs = synth_storage(62, np.full(n, 4))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).cuda()
    d_root = torch.zeros(32, dtype=torch.uint8, device="cuda")
    eng.use_torch_stream()
    eng.state_root_full_dev(t(akeys), t(accs), n, t(skeys), t(svals), t(offs), len(skeys), d_root)
    eng.dev_status()
    eng.set_stream(None)
    assert d_root.cpu().numpy().tobytes() == oracle.state_root_full(akeys, accs, skeys, svals, offs, threads=4)


def test_async_error_is_sticky_across_back_to_back_dev_calls(eng):
    """include/b200trie.h: a violation of an async (*_dev) call is reported by the next b200_sync / b200_dev_status —
    also when another, well-formed *_dev build was enqueued after it (the later build must not erase it)."""
    import torch
    from reth_b200 import B200Error, _lib
    n = 2000
    keys, accs = synth_accounts(63, n)
    bad = keys.copy()
    bad[[100, 101]] = bad[[101, 100]]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).cuda()
    d_bad, d_good, d_accs = t(bad), t(keys), t(accs)
    r1 = torch.zeros(32, dtype=torch.uint8, device="cuda")
    r2 = torch.zeros(32, dtype=torch.uint8, device="cuda")
    eng.use_torch_stream()
    try:
        eng.state_root_dev(d_bad, d_accs, None, n, r1)    # unsorted: flags the error on the device
        eng.state_root_dev(d_good, d_accs, None, n, r2)   # fine on its own
        with pytest.raises(B200Error) as e:
            eng.dev_status()
        assert e.value.status == _lib.ERR_UNSORTED
        eng.dev_status()                                   # reported once; the context is clean again
        assert r2.cpu().numpy().tobytes() == oracle.state_root(keys, accs)  # the second build was not disturbed
        eng.state_root_dev(d_good, d_accs, None, n, r1)
        eng.dev_status()
        assert r1.cpu().numpy().tobytes() == oracle.state_root(keys, accs)
    finally:
        eng.set_stream(None)


# ---------------------------------------------------------------- resident trie / incremental root (BASELINE config 5)
def _mutate(accs, idx, seed):
    rng = np.random.default_rng(seed)
    new = accs[idx].copy()
    new["nonce"] = new["nonce"] + np.uint64(1)
    bal = rng.integers(0, 256, (len(idx), 32), dtype=np.uint8)
    bal[:, :20] = 0
    new["balance"] = bal
    return new


@pytest.mark.parametrize("n,m", [(1, 1), (2, 1), (17, 5), (5000, 300), (200_000, 10_000)])
def test_resident_trie_update_matches_full_rebuild(eng, n, m):
    """incremental == full rebuild is reth's own correctness criterion (crates/trie/db/tests/trie.rs:60-126,
    680-717): after updating m existing accounts the resident trie's root must equal the root of a from-scratch
    build of the modified state; the re-emitted stored nodes must equal the from-scratch records of the same
    paths."""
    from reth_b200 import ResidentTrie
    keys, accs = synth_accounts(1000 + n, n)
    sroots = random_keys(2000 + n, n)
    t = ResidentTrie.create(eng, keys, accs, sroots)
    assert t.root() == oracle.state_root(keys, accs, sroots) and len(t) == n
    rng = np.random.default_rng(n)
    for round_ in range(2):
        idx = np.sort(rng.choice(n, size=min(m, n), replace=False))
        new = _mutate(accs, idx, 7 + round_)
        new_sr = random_keys(3000 + round_, len(idx))
        root, upd, stats = t.update(keys[idx], new, new_sr, want_updates=True, want_stats=True)
        accs[idx] = new
        sroots[idx] = new_sr
        o_root, o_upd = oracle.state_root(keys, accs, sroots, want_updates=True)
        assert root == o_root
        full = {bytes(r[1]): r for r in o_upd}
        assert all(full[bytes(r[1])] == r for r in upd)          # every re-emitted node equals the from-scratch one
        if n >= 5000:
            assert 0 < len(upd) <= len(o_upd) and stats["branches_added"] < 8 * len(idx)
    # shuffled (unsorted) dirty keys are fine too
    idx = rng.choice(n, size=min(m, n), replace=False)
    new = _mutate(accs, idx, 99)
    root = t.update(keys[idx], new)
    accs[idx] = new
    assert root == oracle.state_root(keys, accs, sroots)
    t.close()


def test_resident_trie_rejects_unknown_key_and_stays_consistent(eng):
    from reth_b200 import B200Error, ResidentTrie, _lib
    keys, accs = synth_accounts(31, 3000)
    t = ResidentTrie.create(eng, keys, accs)
    before = t.root()
    bad = keys[[5, 6, 7]].copy()
    bad[1, 31] ^= 0xFF
    with pytest.raises(B200Error) as e:
        t.update(bad, _mutate(accs, np.array([5, 6, 7]), 1))
    assert e.value.status == _lib.ERR_NOT_FOUND
    assert t.root() == before
    new = _mutate(accs, np.array([5]), 2)
    accs[[5]] = new
    assert t.update(keys[[5]], new) == oracle.state_root(keys, accs)   # still usable, nothing half-applied
    # trie without storage roots cannot take them later
    with pytest.raises(B200Error):
        t.update(keys[[6]], accs[[6]], random_keys(1, 1))
    t.close()
    # other builds on the same context still work after buffers were handed to a resident trie
    assert eng.state_root(keys, accs) == oracle.state_root(keys, accs)


def test_pipelined_host_path_large_state(eng):
    """>= 2 Mi slots without retained updates takes the chunked H2D/compute pipeline of b200_state_root_full; the
    root must not depend on the chunking (includes accounts with empty storage and a skewed large trie)."""
    n = 140_000
    akeys, accs = synth_accounts(71, n)
    counts = np.full(n, 16)
    counts[::7] = 0
    counts[12345] = 300_000
    skeys, svals, offs = synth_storage(72, counts)
    assert len(skeys) >= 2 << 20
    root, stats = eng.state_root_full(akeys, accs, skeys, svals, offs, want_stats=True)
    assert root == oracle.state_root_full(akeys, accs, skeys, svals, offs, threads=8)
    assert stats["leaves_added"] == n + len(skeys)
    # the monolithic path (updates retained) agrees
    root2, _, _ = eng.state_root_full(akeys, accs, skeys, svals, offs, want_updates=True)
    assert root2 == root


def test_resident_trie_apply_inserts_deletes_updates(eng):
    """b200_trie_apply with HashedPostStateSorted semantics (upsert / delete, keys ascending): after every batch the
    resident root equals a from-scratch oracle build of the model state — fuzz_state_root_incremental
    (crates/trie/db/tests/trie.rs:680-717) restated.  Pure value batches must take the in-place path."""
    from reth_b200 import ResidentTrie
    rng = np.random.default_rng(5)
    keys0, accs0 = synth_accounts(400, 4000)
    model = {keys0[i].tobytes(): accs0[i].copy() for i in range(len(keys0))}
    t = ResidentTrie.create(eng, keys0, accs0)

    def model_root():
        ks = sorted(model)
        k = np.frombuffer(b"".join(ks), np.uint8).reshape(-1, 32) if ks else np.zeros((0, 32), np.uint8)
        a = np.array([model[x] for x in ks], oracle.ACCOUNT_DTYPE) if ks else np.zeros(0, oracle.ACCOUNT_DTYPE)
        return oracle.state_root(k, a), k, a

    for batch in range(6):
        existing = sorted(model)
        upd = [existing[i] for i in rng.choice(len(existing), 150, replace=False)]
        dele = [existing[i] for i in rng.choice(len(existing), 60 if batch % 2 == 0 else 0, replace=False)] if batch < 5 else []
        ins = [bytes(r) for r in random_keys(9000 + batch, 80 if batch % 3 != 2 else 0)]
        absent_del = [bytes(r) for r in random_keys(9900 + batch, 5)]          # deleting what is not there: no-op
        entries = {}
        for k in upd + ins:
            a = np.zeros((), oracle.ACCOUNT_DTYPE)
            a["nonce"] = rng.integers(0, 1000)
            a["balance"][20:] = rng.integers(0, 256, 12, dtype=np.uint8)
            a["code_hash"] = np.frombuffer(oracle.KECCAK_EMPTY, np.uint8)
            entries[k] = (True, a)
        for k in dele + (absent_del if batch % 2 == 0 else []):
            entries[k] = (False, np.zeros((), oracle.ACCOUNT_DTYPE))
        ks = sorted(entries)
        kk = np.frombuffer(b"".join(ks), np.uint8).reshape(-1, 32)
        aa = np.array([entries[k][1] for k in ks], oracle.ACCOUNT_DTYPE)
        pres = np.array([entries[k][0] for k in ks], np.uint8)
        root, rebuilt, updates = t.apply(kk, aa, pres, want_updates=True)
        for k in ks:
            if entries[k][0]:
                model[k] = entries[k][1]
            else:
                model.pop(k, None)
        exp_root, mk, ma = model_root()
        assert root == exp_root and len(t) == len(model)
        structural = bool(ins) or bool(dele) or (batch % 2 == 0)
        assert rebuilt == structural
        if rebuilt:
            assert updates == oracle.state_root(mk, ma, want_updates=True)[1]      # complete node set of the new trie
    # unsorted dirty keys are rejected, nothing changes
    from reth_b200 import B200Error, _lib
    before = t.root()
    with pytest.raises(B200Error) as e:
        t.apply(kk[::-1].copy(), aa[::-1].copy(), pres[::-1].copy())
    assert e.value.status == _lib.ERR_UNSORTED and t.root() == before
    # delete everything -> empty trie; then re-insert
    ks = sorted(model)
    kk = np.frombuffer(b"".join(ks), np.uint8).reshape(-1, 32)
    root, rebuilt = t.apply(kk, np.zeros(len(ks), oracle.ACCOUNT_DTYPE), np.zeros(len(ks), np.uint8))
    assert root == oracle.EMPTY_ROOT_HASH and rebuilt and len(t) == 0
    root, rebuilt = t.apply(keys0[:10], accs0[:10])
    assert root == oracle.state_root(keys0[:10], accs0[:10]) and len(t) == 10
    t.close()


def test_reentrancy_two_contexts_and_shared_context(eng):
    """ParallelStateRoot calls StorageRoot::calculate from many blocking threads (crates/trie/parallel/src/root.rs:
    111-125): the library must be re-entrant.  Two contexts run concurrently; one context shared by several threads
    serialises internally.  Every thread checks its own results against the oracle."""
    import threading
    from reth_b200 import Engine
    other = Engine(0)
    errors = []

    def worker(e, seed):
        try:
            for it in range(4):
                keys, accs = synth_accounts(seed * 10 + it, 20_000 + 1000 * seed)
                sk, sv, so = synth_storage(seed * 10 + it, np.full(50, 30))
                assert e.state_root(keys, accs) == oracle.state_root(keys, accs)
                assert (e.storage_roots(sk, sv, so) == oracle.storage_roots(sk, sv, so)).all()
                msgs = random_keys(seed * 100 + it, 50_000)
                assert (e.keccak256_fixed(msgs) == oracle.keccak256_fixed(msgs)).all()
        except Exception as ex:  # noqa: BLE001
            errors.append(repr(ex))

    threads = [threading.Thread(target=worker, args=(eng if i % 2 == 0 else other, i)) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    other.close()
    assert not errors, errors


def test_null_and_bad_arguments_are_rejected(eng):
    """The boundary never crashes on bad input: null pointers / inconsistent sizes give B200_ERR_INVALID_ARG."""
    import ctypes as C
    from reth_b200 import _lib
    L = eng.lib
    out = np.zeros(32, np.uint8)
    assert L.b200_keccak256_fixed(eng.ctx, None, 32, 32, 5, out.ctypes.data) == _lib.ERR_INVALID_ARG
    assert L.b200_keccak256_fixed(eng.ctx, out.ctypes.data, 32, 16, 1, out.ctypes.data) == _lib.ERR_INVALID_ARG  # stride < len
    assert L.b200_state_root(eng.ctx, None, None, None, 3, out.ctypes.data, None, None) == _lib.ERR_INVALID_ARG
    assert L.b200_storage_roots(eng.ctx, None, None, None, 1, out.ctypes.data, None, None) == _lib.ERR_INVALID_ARG
    assert L.b200_keccak256_fixed(None, out.ctypes.data, 32, 32, 1, out.ctypes.data) == _lib.ERR_INVALID_ARG
    assert b"bad argument" in L.b200_last_error(eng.ctx)
    # n == 0 is fine everywhere
    assert L.b200_keccak256_fixed(eng.ctx, None, 32, 32, 0, None) == 0
    assert L.b200_state_root(eng.ctx, None, None, None, 0, out.ctypes.data, None, None) == 0
    assert out.tobytes() == oracle.EMPTY_ROOT_HASH
