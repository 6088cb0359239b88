"""Size-independent properties at BASELINE.json's full sizes (C2: 10M keys, C3: 1M accounts x 16 slots, C5: resident
trie updates), where the oracle is too slow to recompute everything: sampled digests against the oracle,
permutation equivariance, sharded == monolithic == pipelined roots, update/revert round trips."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle


@pytest.fixture(scope="module")
def eng():
    from reth_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def test_c2_ten_million_keys(eng):
    import torch
    from bench import random_keys_torch
    n = 10_000_000
    dev = torch.device("cuda", 0)
    keys = random_keys_torch(2, n, dev).view(torch.uint8).view(n, 32)
    out = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    eng.use_torch_stream()
    eng.keccak256_fixed_dev(keys, 32, 32, n, out)
    # permutation equivariance: hashing a permuted batch permutes the digests (no cross-talk between messages)
    perm = torch.randperm(n, device=dev)
    out2 = torch.empty_like(out)
    eng.keccak256_fixed_dev(keys[perm].contiguous(), 32, 32, n, out2)
    torch.cuda.synchronize()
    assert torch.equal(out[perm], out2)
    # a sample of 20k digests against the oracle
    idx = torch.randint(0, n, (20_000,), device=dev)
    assert (out[idx].cpu().numpy() == oracle.keccak256_fixed(keys[idx].cpu().numpy(), threads=4)).all()
    # the 20-byte address variant on the same scale (stride 32, first 20 bytes hashed)
    eng.keccak256_fixed_dev(keys, 20, 32, n, out2)
    torch.cuda.synchronize()
    assert (out2[idx].cpu().numpy() == oracle.keccak256_fixed(keys[idx].cpu().numpy(), 20, threads=4)).all()
    eng.set_stream(None)


def test_c3_sharded_monolithic_and_host_paths_agree(eng):
    import torch
    from bench import make_c3_shard
    dev = torch.device("cuda", 0)
    n_acc, slots = 1_000_000, 16
    sh = make_c3_shard(3, n_acc, slots, 0, 16, dev)
    eng.use_torch_stream()
    d_root = torch.zeros(32, dtype=torch.uint8, device=dev)
    eng.state_root_full_dev(sh["akeys"], sh["accts"], n_acc, sh["skeys"], sh["svals"], sh["offs"], sh["n_slots"], d_root)
    eng.dev_status()
    mono = bytes(d_root.cpu().numpy())
    stats = eng.last_stats()
    assert stats["leaves_added"] == n_acc * (slots + 1)
    assert 1.30 < stats["hashed_nodes"] / stats["leaves_added"] < 1.45      # SURVEY.md Appendix C structure factor
    # 4 emulated ranks by top nibble -> frontier merge -> same root
    akeys = sh["akeys"].view(n_acc, 32)
    top = (akeys[:, 0] >> 4).cpu().numpy()
    merged = torch.zeros(16 * 68, dtype=torch.uint8, device=dev)
    for rank in range(4):
        lo_n, hi_n = rank * 4, rank * 4 + 4
        sel = np.nonzero((top >= lo_n) & (top < hi_n))[0]
        a0, a1 = int(sel[0]), int(sel[-1]) + 1
        s0, s1 = a0 * slots, a1 * slots
        fr = torch.zeros(16 * 68, dtype=torch.uint8, device=dev)
        offs = (sh["offs"][a0:a1 + 1] - sh["offs"][a0]).contiguous()
        eng.subtrie_frontier_dev(akeys[a0:a1].contiguous().view(-1), sh["accts"].view(n_acc, 72)[a0:a1].contiguous().view(-1),
                                 a1 - a0, sh["skeys"].view(-1, 32)[s0:s1].contiguous().view(-1),
                                 sh["svals"].view(-1, 32)[s0:s1].contiguous().view(-1), offs, s1 - s0, fr)
        merged.view(16, 68)[lo_n:hi_n] = fr.view(16, 68)[lo_n:hi_n]
    eng.root_from_frontier_dev(merged, d_root)
    eng.dev_status()
    assert bytes(d_root.cpu().numpy()) == mono
    eng.set_stream(None)
    # host-pointer path (chunked H2D/compute pipeline) on the same data
    from reth_b200 import ACCOUNT_DTYPE
    h = lambda t: t.cpu().numpy()
    root = eng.state_root_full(h(sh["akeys"]).reshape(-1, 32), h(sh["accts"]).view(ACCOUNT_DTYPE).reshape(-1),
                               h(sh["skeys"]).reshape(-1, 32), h(sh["svals"]).reshape(-1, 32),
                               h(sh["offs"]).astype(np.uint64))
    assert root == mono
    # a 30k-account prefix of the same data against the oracle (exact)
    m = 30_000
    exp = oracle.state_root_full(h(akeys[:m]), h(sh["accts"].view(n_acc, 72)[:m]).view(ACCOUNT_DTYPE).reshape(-1),
                                 h(sh["skeys"].view(-1, 32)[:m * slots]), h(sh["svals"].view(-1, 32)[:m * slots]),
                                 h(sh["offs"][:m + 1]).astype(np.uint64), threads=8)
    got = eng.state_root_full(h(akeys[:m]), h(sh["accts"].view(n_acc, 72)[:m]).view(ACCOUNT_DTYPE).reshape(-1),
                              h(sh["skeys"].view(-1, 32)[:m * slots]), h(sh["svals"].view(-1, 32)[:m * slots]),
                              h(sh["offs"][:m + 1]).astype(np.uint64))
    assert got == exp


def test_c5_update_revert_round_trip(eng):
    """10M-leaf resident trie: update 10k accounts, then write the old values back -> the original root; the updated
    root equals a from-scratch build of the modified state on the device."""
    import torch
    from bench import be_sort_key, random_keys_torch, splitmix64_torch
    from reth_b200 import ResidentTrie
    dev = torch.device("cuda", 0)
    n, m = 10_000_000, 10_000
    keys = random_keys_torch(5, n, dev)
    keys = keys[torch.sort(be_sort_key(keys), stable=True).indices].contiguous()
    accts = torch.zeros((n, 72), dtype=torch.uint8, device=dev)
    accts[:, 32:40] = splitmix64_torch(9, n, dev).view(torch.uint8).view(n, 8)
    eng.use_torch_stream()
    root0 = torch.zeros(32, dtype=torch.uint8, device=dev)
    trie = ResidentTrie.create_dev(eng, keys.view(torch.uint8).view(-1), accts.view(-1), None, n, root0)
    idx = torch.unique(torch.randint(0, n, (m + 2000,), device=dev))[:m]
    mm = int(idx.numel())
    dk = keys[idx].contiguous().view(torch.uint8).view(-1)
    old = accts[idx].clone()
    new = old.clone()
    new[:, 0] = 7
    new[:, 33] ^= 0x5A
    root1 = torch.zeros(32, dtype=torch.uint8, device=dev)
    trie.update_dev(dk, new.view(-1), None, mm, root1)
    eng.dev_status()
    assert not torch.equal(root0, root1)
    # from-scratch build of the modified state
    accts2 = accts.clone()
    accts2[idx] = new
    root_full = torch.zeros(32, dtype=torch.uint8, device=dev)
    eng.state_root_dev(keys.view(torch.uint8).view(-1), accts2.view(-1), None, n, root_full)
    eng.dev_status()
    assert torch.equal(root1, root_full)
    # revert
    root2 = torch.zeros(32, dtype=torch.uint8, device=dev)
    trie.update_dev(dk, old.view(-1), None, mm, root2)
    eng.dev_status()
    assert torch.equal(root2, root0)
    trie.close()
    eng.set_stream(None)


def test_c4_mainnet_shape_small_vs_oracle(eng):
    """The mainnet-shaped generator of bench.py (80% EOAs, Zipf(1.2) storage sizes: one huge trie, a long tail of
    1-slot tries) at 300k leaves: root and TrieUpdates against the oracle."""
    import torch
    from bench import make_c4_shard
    from reth_b200 import ACCOUNT_DTYPE
    dev = torch.device("cuda", 0)
    sh = make_c4_shard(4, 300_000, 0, 16, dev)
    h = lambda t: t.cpu().numpy()
    args = (h(sh["akeys"]).reshape(-1, 32), h(sh["accts"]).view(ACCOUNT_DTYPE).reshape(-1), h(sh["skeys"]).reshape(-1, 32),
            h(sh["svals"]).reshape(-1, 32), h(sh["offs"]).astype(np.uint64))
    assert sh["max_trie"] > 20_000 and int((np.diff(args[4]) == 1).sum()) > 3_000
    root, au, su = eng.state_root_full(*args, want_updates=True)
    o_root, o_au, o_su = oracle.state_root_full(*args, want_updates=True, threads=8)
    assert root == o_root and au == o_au and su == o_su


def test_c3_streamed_pushes_equal_one_shot_with_updates(eng):
    """a14 at the C3 shape (1M accounts x 16 slots): the state pushed in 11 account-key ranges through b200_root_stream_* ==
    one b200_state_root_full call — the root, and the stored nodes (TrieUpdates) as multisets of records."""
    import hashlib

    import torch

    from bench import make_c3_shard
    from reth_b200 import ACCOUNT_DTYPE, RootStream
    dev = torch.device("cuda", 0)
    n_acc, slots = 1_000_000, 16
    sh = make_c3_shard(3, n_acc, slots, 0, 16, dev)
    akeys = sh["akeys"].view(n_acc, 32).cpu().numpy()
    accts = sh["accts"].view(n_acc, 72).cpu().numpy().view(ACCOUNT_DTYPE).reshape(-1)
    skeys = sh["skeys"].view(-1, 32).cpu().numpy()
    svals = sh["svals"].view(-1, 32).cpu().numpy()
    offs = sh["offs"].cpu().numpy().astype(np.uint64)
    del sh
    eng.set_stream(None)
    root, au, su = eng.state_root_full(akeys, accts, skeys, svals, offs, want_updates=True)

    def digest(acct_recs, stor_recs):
        # order-independent fingerprint of the record sets (storage records keyed by the owning account's key)
        ha, hs = 0, 0
        for r in acct_recs:
            ha ^= int.from_bytes(hashlib.blake2b(repr((bytes(r[1]), r[2], r[3], r[4], r[5])).encode(), digest_size=16).digest(), "big")
        for key, r in stor_recs:
            hs ^= int.from_bytes(hashlib.blake2b(repr((key, bytes(r[1]), r[2], r[3], r[4], r[5])).encode(), digest_size=16).digest(), "big")
        return ha, hs, len(acct_recs), len(stor_recs)

    want = digest(au, [(akeys[r[0]].tobytes(), r) for r in su])
    cuts = [0, 1, 90_000, 250_000, 250_001, 400_000, 555_555, 700_000, 812_345, 950_000, 999_999, n_acc]
    s = RootStream(eng, retain_updates=True)
    acct_recs, stor_recs = [], []
    for a0, a1 in zip(cuts[:-1], cuts[1:]):
        s0, s1 = int(offs[a0]), int(offs[a1])
        prog, a_r, s_r = s.push(akeys[a0:a1], accts[a0:a1], skeys[s0:s1], svals[s0:s1], (offs[a0:a1 + 1] - offs[a0]).astype(np.uint64))
        assert prog["accounts"] == a1 and prog["open_accounts"] <= n_acc // 8
        acct_recs += a_r
        stor_recs += [(akeys[a0 + r[0]].tobytes(), r) for r in s_r]
    s_root, a_r = s.finish()
    acct_recs += a_r
    s.close()
    assert s_root == root
    assert digest(acct_recs, stor_recs) == want
