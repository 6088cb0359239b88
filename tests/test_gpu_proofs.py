"""Merkle proofs from the dynamic resident state (b200_dstate_account_proofs / _storage_proofs) against the
reference's own vectors — crates/trie/db/tests/proof.rs: testspec_proofs (:44-115), testspec_empty_storage_proof
(:117-140), mainnet_genesis_account_proof (:142-165) — and, for random states, against a verifier that walks the proof
from the state root exactly like `AccountProof::verify` (hash of node k+1 must sit in node k at the key's next nibble)."""
import os

import numpy as np
import pytest

import oracle
from tests.util import alloc_to_flat

pytestmark = [pytest.mark.gpu]

H = bytes.fromhex
TESTSPEC = {  # address -> proof, crates/trie/db/tests/proof.rs:50-98
    "2031f89b3ea8014eb51a78c316e42af3e0d7695f": [
        "e48200a7a040f916999be583c572cc4dd369ec53b0a99f7de95f13880cf203d98f935ed1b3",
        "f87180a04fb9bab4bb88c062f32452b7c94c8f64d07b5851d44a39f1e32ba4b1829fdbfb8080808080a0b61eeb2eb82808b73c4ad14140a2836689f4ab8445d69dd40554eaf1fce34bc080808080808080a0dea230ff2026e65de419288183a340125b04b8405cc61627b3b4137e2260a1e880",
        "f8719f31355ec1c8f7e26bb3ccbcb0b75d870d15846c0b98e5cc452db46c37faea40b84ff84d80890270801d946c940000a056e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421a0c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470",
    ],
    "33f0fc440b8477fcfbe9d0bf8649e7dea9baedb2": [
        "e48200a7a040f916999be583c572cc4dd369ec53b0a99f7de95f13880cf203d98f935ed1b3",
        "f87180a04fb9bab4bb88c062f32452b7c94c8f64d07b5851d44a39f1e32ba4b1829fdbfb8080808080a0b61eeb2eb82808b73c4ad14140a2836689f4ab8445d69dd40554eaf1fce34bc080808080808080a0dea230ff2026e65de419288183a340125b04b8405cc61627b3b4137e2260a1e880",
        "e48200d3a0ef957210bca5b9b402d614eb8408c88cfbf4913eb6ab83ca233c8b8f0e626b54",
        "f851808080a02743a5addaf4cf9b8c0c073e1eaa555deaaf8c41cb2b41958e88624fa45c2d908080808080a0bfbf6937911dfb88113fecdaa6bde822e4e99dae62489fcf61a91cb2f36793d680808080808080",
        "f8679e207781e762f3577784bab7491fcc43e291ce5a356b9bc517ac52eed3a37ab846f8448001a056e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421a0c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470",
    ],
    "62b0dd4aab2b1a0a04e279e2b828791a10755528": [
        "e48200a7a040f916999be583c572cc4dd369ec53b0a99f7de95f13880cf203d98f935ed1b3",
        "f87180a04fb9bab4bb88c062f32452b7c94c8f64d07b5851d44a39f1e32ba4b1829fdbfb8080808080a0b61eeb2eb82808b73c4ad14140a2836689f4ab8445d69dd40554eaf1fce34bc080808080808080a0dea230ff2026e65de419288183a340125b04b8405cc61627b3b4137e2260a1e880",
        "f8709f3936599f93b769acf90c7178fd2ddcac1b5b4bc9949ee5a04b7e0823c2446eb84ef84c80880f43fc2c04ee0000a056e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421a0c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470",
    ],
    "1ed9b1dd266b607ee278726d324b855a093394a6": [
        "e48200a7a040f916999be583c572cc4dd369ec53b0a99f7de95f13880cf203d98f935ed1b3",
        "f87180a04fb9bab4bb88c062f32452b7c94c8f64d07b5851d44a39f1e32ba4b1829fdbfb8080808080a0b61eeb2eb82808b73c4ad14140a2836689f4ab8445d69dd40554eaf1fce34bc080808080808080a0dea230ff2026e65de419288183a340125b04b8405cc61627b3b4137e2260a1e880",
        "e48200d3a0ef957210bca5b9b402d614eb8408c88cfbf4913eb6ab83ca233c8b8f0e626b54",
        "f851808080a02743a5addaf4cf9b8c0c073e1eaa555deaaf8c41cb2b41958e88624fa45c2d908080808080a0bfbf6937911dfb88113fecdaa6bde822e4e99dae62489fcf61a91cb2f36793d680808080808080",
        "f86f9e207a32b8ab5eb4b043c65b1f00c93f517bc8883c5cd31baf8e8a279475e3b84ef84c808801aa535d3d0c0000a056e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421a0c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470",
    ],
}
MAINNET_TARGET = "000d836201318ec6899a67540690382780743280"          # proof.rs:149-151
MAINNET_PROOF_HEADS = [  # proof.rs:154-160 (full RLPs there; lengths and both ends pinned here, hashes chain to the root)
    ("f90211a090dcaf88c40c7bbc95a912cbdde67c175767b31173df9ee4b0d733bfdd511c43", "a089d613f26159af43616fd9455bb461f4869bfede26f2130835ed067a8b967bfb80", 0x214),
    ("f90211a0dae48f5b47930c28bb116fbd55e52cd47242c71bf55373b55eb2805ee2e4a929", "a049bf6e8df0acafd0eff86defeeb305568e44d52d2235cf340ae15c6034e2b24180", 0x214),
    ("f901f1a0cf67e0f5d5f8d70e53a6278056a14ddca46846f5ef69c7bde6810d058d4a9eda80", "a0cd367d0679950e9c5f2aa4298fd4b081ade2ea429d71ff390c50f8520e16e30880", 0x1f4),
    ("f87180808080808080a0dbee8b33c73b86df839f309f7ac92eee19836e08b39302ffa33921b3c6a09f66", "a0fe7779c7d58c2fda43eba0a6644043c86ebb9ceb4836f89e30831f23eb059ece8080", 0x73),
    ("f8719f20b71c90b0d523dd5004cf206f325748da347685071b34812e21801f5270c4b84ff84d80890ad78ebc5ac6200000", "a0c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470", 0x73),
]


@pytest.fixture(scope="module")
def eng():
    from reth_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def state_from_alloc(eng, alloc):
    from reth_b200 import DynamicState
    keys, accs, skeys, svals, offs = alloc_to_flat(alloc)
    return DynamicState.create(eng, keys, accs, skeys, svals, offs), keys


# ---- a proof verifier (what AccountProof::verify / verify_proof of alloy-trie do) ---------------------------------------
def rlp_items(b):
    """decode one RLP list -> list of raw item payloads (strings) or nested raw lists (kept as bytes incl. header)"""
    def head(buf, i):
        x = buf[i]
        if x < 0x80:
            return i, i + 1, False
        if x < 0xb8:
            return i + 1, i + 1 + (x - 0x80), False
        if x < 0xc0:
            ll = x - 0xb7
            n = int.from_bytes(buf[i + 1:i + 1 + ll], "big")
            return i + 1 + ll, i + 1 + ll + n, False
        if x < 0xf8:
            return i + 1, i + 1 + (x - 0xc0), True
        ll = x - 0xf7
        n = int.from_bytes(buf[i + 1:i + 1 + ll], "big")
        return i + 1 + ll, i + 1 + ll + n, True
    s, e, is_list = head(b, 0)
    assert is_list and e == len(b)
    out, i = [], s
    while i < e:
        ps, pe, pl = head(b, i)
        out.append(b[i:pe] if pl else b[ps:pe])   # nested lists (inline nodes) keep their header
        i = pe
    return out


def nibbles(key):
    return [x for byte in key for x in (byte >> 4, byte & 15)]


def verify(root, key, proof, expect_value):
    """expect_value: leaf value bytes for inclusion, None for exclusion"""
    want = root
    path = nibbles(key)
    pos = 0
    for i, node in enumerate(proof):
        if len(node) >= 32 or i == 0:
            assert oracle.keccak256(node) == want, f"node {i} does not hash to the reference held by its parent"
        else:
            assert node == want, f"inline node {i} differs from the bytes embedded in its parent"
        items = rlp_items(node)
        if len(items) == 17:
            if pos == 64:
                return
            child = items[path[pos]]
            pos += 1
            if child == b"":
                assert expect_value is None and i == len(proof) - 1
                return
            want = child
        else:
            assert len(items) == 2
            hp = items[0]
            flag = hp[0] >> 4
            nib = ([hp[0] & 15] if flag & 1 else []) + nibbles(hp[1:])
            if flag & 2:   # leaf
                assert i == len(proof) - 1
                if path[pos:] == nib:
                    assert expect_value is not None and items[1] == expect_value
                else:
                    assert expect_value is None
                return
            if path[pos:pos + len(nib)] != nib:      # diverging extension: exclusion
                assert expect_value is None and i == len(proof) - 1
                return
            pos += len(nib)
            want = items[1]
    raise AssertionError("proof ended without a terminal node")


# ---- reference vectors ---------------------------------------------------------------------------------------------------
def test_testspec_account_proofs(eng, golden_allocs):
    ds, _ = state_from_alloc(eng, golden_allocs["testspec"]["alloc"])
    targets = list(TESTSPEC)
    hashed = np.frombuffer(b"".join(oracle.keccak256(H(a)) for a in targets), np.uint8).reshape(-1, 32)
    proofs = ds.account_proofs(hashed)
    for addr, proof in zip(targets, proofs):
        assert [p.hex() for p in proof] == TESTSPEC[addr], addr
    # testspec_empty_storage_proof: both slots of an account without storage prove against the empty trie
    slots = np.frombuffer(oracle.keccak256((1).to_bytes(32, "big")) + oracle.keccak256((3).to_bytes(32, "big")), np.uint8).reshape(2, 32)
    sroot, sproofs = ds.storage_proofs(oracle.keccak256(H("1ed9b1dd266b607ee278726d324b855a093394a6")), slots)
    assert sroot == oracle.EMPTY_ROOT_HASH and sproofs == [[b"\x80"], [b"\x80"]]
    ds.close()


def test_mainnet_genesis_account_proof(eng, golden_allocs):
    ds, _ = state_from_alloc(eng, golden_allocs["mainnet"]["alloc"])
    key = oracle.keccak256(H(MAINNET_TARGET))
    assert key.hex() == "cf67b71c90b0d523dd5004cf206f325748da347685071b34812e21801f5270c4"    # proof.rs:148
    (proof,) = ds.account_proofs(np.frombuffer(key, np.uint8).reshape(1, 32))
    assert len(proof) == len(MAINNET_PROOF_HEADS)
    for node, (head, tail, length) in zip(proof, MAINNET_PROOF_HEADS):
        assert node.hex().startswith(head) and node.hex().endswith(tail) and len(node) == length
    root = H(golden_allocs["mainnet"]["state_root"])
    value = rlp_items(proof[-1])[1]
    verify(root, key, proof, value)
    ds.close()


# ---- random states ---------------------------------------------------------------------------------------------------------
def test_random_state_inclusion_and_exclusion_proofs(eng):
    from reth_b200 import DynamicState
    from tests.util import synth_accounts, synth_storage
    n = 3000
    keys, accs = synth_accounts(61, n)
    counts = np.where(np.arange(n) % 7 == 0, 1 + (np.arange(n) % 60), 0)
    skeys, svals, offs = synth_storage(62, counts, "mixed")
    ds = DynamicState.create(eng, keys, accs, skeys, svals, offs)
    root = oracle.state_root_full(keys, accs, skeys, svals, offs)
    sroots = oracle.storage_roots(skeys, svals, offs)
    rng = np.random.default_rng(5)
    present = keys[rng.choice(n, 40, replace=False)]
    absent = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    near = present.copy()
    near[:, 31] ^= 1                                     # share 62+ nibbles with an existing key
    targets = np.concatenate([present, absent, near])
    proofs = ds.account_proofs(targets)
    index = {keys[i].tobytes(): i for i in range(n)}
    for t, proof in zip(targets, proofs):
        i = index.get(t.tobytes())
        if i is None:
            verify(root, t.tobytes(), proof, None)
        else:
            val = oracle.encode_trie_account(int(accs[i]["nonce"]), int.from_bytes(accs[i]["balance"].tobytes(), "big"),
                                             sroots[i].tobytes(), accs[i]["code_hash"].tobytes())
            verify(root, t.tobytes(), proof, val)
    # storage proofs of an account with many slots, one with a single slot, and one without storage
    for a in (int(np.argmax(counts)), int(np.nonzero(counts == 1)[0][0]), 1):
        lo, hi = int(offs[a]), int(offs[a + 1])
        have = skeys[lo:hi]
        tgt = np.concatenate([have[:10], rng.integers(0, 256, (5, 32), dtype=np.uint8)])
        sroot, sproofs = ds.storage_proofs(keys[a].tobytes(), tgt)
        assert sroot == sroots[a].tobytes()
        slot_val = {have[j].tobytes(): svals[lo + j] for j in range(hi - lo)}
        for t, proof in zip(tgt, sproofs):
            if hi == lo:
                assert proof == [b"\x80"]
                continue
            v = slot_val.get(t.tobytes())
            verify(sroot, t.tobytes(), proof, None if v is None else oracle.encode_u256(int.from_bytes(v.tobytes(), "big")))
    # an account that does not exist: empty storage root and one-node proofs
    sroot, sproofs = ds.storage_proofs(bytes(32), absent[:2])
    assert sroot == oracle.EMPTY_ROOT_HASH and sproofs == [[b"\x80"], [b"\x80"]]
    ds.close()


def test_proofs_follow_the_state_through_blocks(eng):
    from tests.test_gpu_dstate import EXISTS, UNCHANGED, Harness, acct, random_block, random_state, rkey
    rng = np.random.default_rng(12)
    h = Harness(eng, random_state(rng, 300))
    for step in range(3):
        root = h.commit(random_block(rng, h.state, 40, step + 1))
        live = sorted(h.state)
        tg = [live[i] for i in rng.choice(len(live), 10, replace=False)] + [rkey(rng) for _ in range(5)]
        proofs = h.ds.account_proofs(np.frombuffer(b"".join(tg), np.uint8).reshape(-1, 32))
        for k, proof in zip(tg, proofs):
            if k in h.state:
                assert rlp_items(proof[-1])[0][0] >> 4 in (2, 3)      # ends in a leaf ...
                verify(root, k, proof, rlp_items(proof[-1])[1])        # ... that chains to the new root
            else:
                verify(root, k, proof, None)
    h.ds.close()


def test_account_multiproof_is_the_union_of_the_proofs(eng, golden_allocs):
    """MultiProof::account_subtree: path -> node over all targets.  On the testspec state the four targets of proof.rs give
    exactly the eight distinct nodes of the trie, keyed by their positions (root extension at the empty path, ...)."""
    ds, keys = state_from_alloc(eng, golden_allocs["testspec"]["alloc"])
    hashed = np.frombuffer(b"".join(oracle.keccak256(H(a)) for a in TESTSPEC), np.uint8).reshape(-1, 32)
    mp = ds.account_multiproof(hashed)
    all_nodes = {n for proof in TESTSPEC.values() for n in proof}
    assert {v.hex() for v in mp.values()} == all_nodes and len(mp) == len(all_nodes) == 8
    assert mp[b""].hex() == TESTSPEC["2031f89b3ea8014eb51a78c316e42af3e0d7695f"][0]          # the root extension node
    for path, rlp in mp.items():                                                               # every path is a prefix of a target
        assert any(bytes(nibbles(k.tobytes()))[:len(path)] == path for k in hashed)
    # the extension 'a7' (2 nibbles) puts the first branch at path [a, 7]
    assert bytes([0xa, 0x7]) in mp and len(rlp_items(mp[bytes([0xa, 0x7])])) == 17
    ds.close()


def test_multiproof_batch_equals_the_single_proofs(eng):
    """b200_dstate_multiproof (Proof::multiproof over MultiProofTargets, proof/mod.rs:143-193): accounts with their slot targets
    in one call == the union of the per-target account proofs and per-account storage proofs, storage roots included; absent
    accounts and accounts without storage behave as in testspec_empty_storage_proof (proof.rs:117-140)."""
    from reth_b200 import DynamicState
    from tests.util import synth_accounts, synth_storage
    n = 2000
    keys, accs = synth_accounts(71, n)
    counts = np.where(np.arange(n) % 5 == 0, 1 + (np.arange(n) % 40), 0)
    skeys, svals, offs = synth_storage(72, counts, "mixed")
    ds = DynamicState.create(eng, keys, accs, skeys, svals, offs)
    rng = np.random.default_rng(8)
    targets = {}
    for a in rng.choice(n, 25, replace=False):
        lo, hi = int(offs[a]), int(offs[a + 1])
        have = [skeys[j].tobytes() for j in range(lo, min(hi, lo + 6))]
        targets[keys[a].tobytes()] = have + [bytes(rng.integers(0, 256, 32, dtype=np.uint8)) for _ in range(2)]
    targets[bytes(rng.integers(0, 256, 32, dtype=np.uint8))] = [bytes(32)]        # an account that does not exist
    targets[keys[1].tobytes()] = []                                               # an account target without slot targets
    mp = ds.multiproof(targets)
    addrs = sorted(targets)
    assert mp["account_subtree"] == ds.account_multiproof(np.frombuffer(b"".join(addrs), np.uint8).reshape(-1, 32))
    nib = lambda k: bytes(x for b in k for x in (b >> 4, b & 15))
    for a in addrs:
        sl = sorted(set(targets[a]))
        tgt = np.frombuffer(b"".join(sl), np.uint8).reshape(-1, 32) if sl else np.zeros((0, 32), np.uint8)
        sroot, sproofs = ds.storage_proofs(a, tgt)
        assert mp["storages"][a]["root"] == sroot
        want = {}
        for k, proof in zip(sl, sproofs):
            # node k of a proof sits at the path consumed so far: recompute it by walking the proof
            path = b""
            for node in proof:
                want[path] = node
                items = rlp_items(node) if node != b"\x80" else []
                if len(items) == 17:
                    path = path + nib(k)[len(path):len(path) + 1]
                elif len(items) == 2 and not (items[0][0] & 0x20):
                    enc = items[0]
                    ext = list(nib(enc))[1:] if enc[0] & 0x10 else list(nib(enc))[2:]
                    path = path + bytes(ext)
        assert mp["storages"][a]["subtree"] == want
    # branch_node_masks (Proof::with_branch_node_masks): exactly the branch nodes of the proof that the hash builder stores —
    # the oracle's updated_branch_nodes of a full build, restricted to the proof's paths (the root path is left aside:
    # TrieUpdates drops it, the proof may carry it)
    sroots = oracle.storage_roots(skeys, svals, offs)
    _, acct_upd = oracle.state_root(keys, accs, sroots, want_updates=True)
    stored = {bytes(r[1]): (r[4], r[3]) for r in acct_upd}                       # path -> (hash_mask, tree_mask)
    got = {pth: m for pth, m in mp["branch_node_masks"].items() if pth}
    on_proof = {pth for pth in mp["account_subtree"] if pth}
    assert got == {pth: m for pth, m in stored.items() if pth in on_proof} and got
    _, sto_upd = oracle.storage_roots(skeys, svals, offs, want_updates=True)
    index_of = {keys[i].tobytes(): i for i in range(n)}
    checked = 0
    for a in addrs:
        if a not in index_of:
            assert mp["storages"][a]["branch_node_masks"] == {}
            continue
        stored = {bytes(r[1]): (r[4], r[3]) for r in sto_upd if r[0] == index_of[a]}
        sub = mp["storages"][a]
        got = {pth: m for pth, m in sub["branch_node_masks"].items() if pth}
        assert got == {pth: m for pth, m in stored.items() if pth in sub["subtree"]}
        checked += len(got)
    assert checked > 0
    ds.close()
