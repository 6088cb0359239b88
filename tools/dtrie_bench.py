#!/usr/bin/env python3
"""Round-2 measurement aid for the dynamic resident trie (b200_dtrie_*): first B200 validation + latency.

    B200_DTRIE_ON_GPU=1 python -m pytest tests/test_gpu_dtrie.py -m gpu -q      # correctness first
    python tools/dtrie_bench.py --base 100000000 --dirty 10000 --mix 80,10,10   # then latency

Builds a base trie of --base accounts on the device, then applies blocks of --dirty keys with the given
update,insert,delete percentages through the host-pointer C ABI (the call a reth shim makes: H2D of the dirty set and D2H
of the root inside) and prints one JSON line: median wall µs, device ms of the build stats, re-hashed nodes, and the same
block through b200_trie_apply (merge + rebuild) for comparison.  Roots of the two paths are compared every block."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--base", type=int, default=10_000_000)
    ap.add_argument("--dirty", type=int, default=10_000)
    ap.add_argument("--mix", default="80,10,10", help="update,insert,delete percent")
    ap.add_argument("--blocks", type=int, default=12)
    ap.add_argument("--compare", action="store_true", help="also run every block through b200_trie_apply and compare roots")
    args = ap.parse_args()
    import torch
    from bench import be_sort_key, random_keys_torch, splitmix64_torch
    from reth_b200 import ACCOUNT_DTYPE, DynamicTrie, Engine, ResidentTrie
    dev = torch.device("cuda", 0)
    eng = Engine(0)
    eng.use_torch_stream()
    n, m = args.base, args.dirty
    pu, pi, pd = (int(x) for x in args.mix.split(","))
    keys = random_keys_torch(5, n, dev)
    keys = keys[torch.sort(be_sort_key(keys), stable=True).indices].contiguous()
    accts = torch.zeros((n, 72), dtype=torch.uint8, device=dev)
    accts[:, 32:40] = splitmix64_torch(5 ^ 0xACC0, n, dev).view(torch.uint8).view(n, 8)
    accts[:, 40:72] = torch.frombuffer(bytearray(bytes.fromhex(
        "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470")), dtype=torch.uint8).to(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    trie = DynamicTrie.create_dev(eng, keys.view(torch.uint8).view(-1), accts.view(-1), None, n)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    ref = ResidentTrie.create_dev(eng, keys.view(torch.uint8).view(-1), accts.view(-1), None, n) if args.compare else None
    assert ref is None or trie.root() == ref.root()
    eng.set_stream(None)
    rng = np.random.default_rng(55)
    live = keys.view(torch.uint8).view(n, 32).cpu().numpy()   # host copy of the key set, kept in step with the trie
    live_set = None
    lat, dev_ms, built, mismatches = [], [], [], 0
    for b in range(args.blocks):
        n_upd, n_ins, n_del = m * pu // 100, m * pi // 100, m * pd // 100
        pick = rng.choice(len(live), n_upd + n_del, replace=False)
        upd_keys, del_keys = live[pick[:n_upd]], live[pick[n_upd:]]
        ins_keys = rng.integers(0, 256, (n_ins, 32), dtype=np.uint8)
        dk = np.concatenate([upd_keys, del_keys, ins_keys])
        present = np.concatenate([np.ones(n_upd, np.uint8), np.zeros(n_del, np.uint8), np.ones(n_ins, np.uint8)])
        order = np.lexsort(tuple(dk[:, i] for i in range(31, -1, -1)))
        dk, present = np.ascontiguousarray(dk[order]), np.ascontiguousarray(present[order])
        da = np.zeros(len(dk), ACCOUNT_DTYPE)
        da["nonce"] = b + 1
        da["balance"][:, 24:] = rng.integers(0, 256, (len(dk), 8), dtype=np.uint8)
        da["code_hash"] = np.frombuffer(bytes.fromhex("c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"), np.uint8)
        t0 = time.perf_counter()
        root, stats = trie.apply(dk, da, present, want_stats=True)
        wall = time.perf_counter() - t0
        if b >= 2:
            lat.append(wall * 1e6)
            dev_ms.append(stats["device_ms"])
            built.append(stats["branches_added"])
        if ref is not None:
            r2 = ref.apply(dk, da, present)[0]
            mismatches += r2 != root
        mask = np.ones(len(live), bool)
        mask[pick[n_upd:]] = False
        live = np.concatenate([live[mask], ins_keys])
    print(json.dumps({"base_leaves": n, "dirty": m, "mix_update_insert_delete": [pu, pi, pd], "blocks": args.blocks,
                      "apply_wall_us_median": float(np.median(lat)), "apply_device_ms_median": float(np.median(dev_ms)),
                      "rehashed_nodes_median": float(np.median(built)), "base_build_ms": build_s * 1e3,
                      "leaves_after": len(trie), "resident_bytes": trie.device_bytes(),
                      "root_mismatches_vs_trie_apply": mismatches if ref is not None else None}))


if __name__ == "__main__":
    main()
