#!/usr/bin/env python3
"""Round-2 measurement aid for the dynamic resident state (b200_dstate_*).

    python -m pytest tests/test_gpu_dstate.py tests/test_gpu_dtrie.py -m gpu -q   # correctness first
    python tools/dstate_bench.py --accounts 1000000 --slots 16 --touch 2000 --slot-writes 10

Seeds a state of --accounts accounts x --slots slots (the C3 shape), then commits blocks that touch --touch accounts:
60 % storage-only (--slot-writes new / changed / zeroed slots each), 25 % balance changes, 10 % new accounts with storage,
5 % destroyed.  Every block goes through the host-pointer C ABI (what a reth shim calls: H2D of the block and D2H of the
root inside).  Prints one JSON line; with --check the root of every block is compared with b200_state_root_full over the
merged state kept on the host (slow: only for small sizes)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--accounts", type=int, default=1_000_000)
    ap.add_argument("--slots", type=int, default=16)
    ap.add_argument("--touch", type=int, default=2000)
    ap.add_argument("--slot-writes", type=int, default=10)
    ap.add_argument("--blocks", type=int, default=12)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="accounts of the oracle's from-scratch build timed as the CPU baseline (0 = skip)")
    ap.add_argument("--device-resident", action="store_true",
                    help="also apply every block from device memory (b200_dstate_apply_dev) on a twin state and compare roots")
    args = ap.parse_args()
    from reth_b200 import ACCOUNT_DTYPE, KECCAK_EMPTY, DynamicState, Engine
    eng = Engine(0)
    n = args.accounts
    g = np.random.default_rng(3)

    def sorted_keys(count, groups=None):
        k = g.integers(0, 256, (count, 32), dtype=np.uint8)
        cols = tuple(k[:, i] for i in range(31, -1, -1))
        order = np.lexsort(cols if groups is None else cols + (groups,))
        return k[order]

    keys = sorted_keys(n)
    accs = np.zeros(n, ACCOUNT_DTYPE)
    accs["nonce"] = g.integers(0, 1 << 16, n)
    accs["balance"][:, 24:] = g.integers(0, 256, (n, 8), dtype=np.uint8)
    accs["code_hash"] = np.frombuffer(KECCAK_EMPTY, np.uint8)
    total = n * args.slots
    skeys = sorted_keys(total, np.repeat(np.arange(n), args.slots))
    svals = np.zeros((total, 32), np.uint8)
    svals[:, 24:] = g.integers(0, 256, (total, 8), dtype=np.uint8)
    svals[:, 31] |= 1
    offs = (np.arange(n + 1, dtype=np.uint64) * np.uint64(args.slots))
    t0 = time.perf_counter()
    ds = DynamicState.create(eng, keys, accs, skeys, svals, offs)
    build_s = time.perf_counter() - t0
    rng = np.random.default_rng(77)
    state = None
    if args.check:
        state = {keys[i].tobytes(): (accs[i].copy(), {skeys[j].tobytes(): svals[j].tobytes() for j in range(int(offs[i]), int(offs[i + 1]))})
                 for i in range(n)}
    live = [keys[i].tobytes() for i in range(n)]
    live_base = {k: i for i, k in enumerate(live)}   # base account -> row
    slot_of = {}  # a few known slots per touched account, to zero / change later
    twin = DynamicState.create(eng, keys, accs, skeys, svals, offs) if args.device_resident else None
    lat, dev_ms, built, mism, lat_dev, mism_twin, launches = [], [], [], 0, [], 0, []
    seed_root = ds.root()
    touched_existing, inserted_accounts = set(), set()
    base_set = None
    EX, UN, WI = DynamicState.EXISTS, DynamicState.UNCHANGED, DynamicState.WIPED
    for b in range(args.blocks):
        block = {}
        pick = rng.choice(len(live), args.touch, replace=False)
        for q, pi in enumerate(pick):
            k = live[pi]
            r = q / args.touch
            if r < 0.60:
                slots = {rng.integers(0, 256, 32, dtype=np.uint8).tobytes(): int(rng.integers(1, 2**60)).to_bytes(32, "big")
                         for _ in range(args.slot_writes)}
                for s in slot_of.get(k, [])[:2]:
                    slots[s] = bytes(32) if rng.random() < 0.5 else int(rng.integers(1, 2**60)).to_bytes(32, "big")
                slot_of.setdefault(k, []).extend(list(slots)[:2])
                block[k] = (EX | UN, None, slots)
            elif r < 0.85:
                a = np.zeros((), ACCOUNT_DTYPE)
                a["nonce"] = b + 1
                a["balance"][24:] = rng.integers(0, 256, 8, dtype=np.uint8)
                a["code_hash"] = np.frombuffer(bytes.fromhex("c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"), np.uint8)
                block[k] = (EX, a, {})
            elif r < 0.95:
                nk = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
                a = np.zeros((), ACCOUNT_DTYPE)
                a["code_hash"] = np.frombuffer(bytes.fromhex("c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"), np.uint8)
                block[nk] = (EX, a, {rng.integers(0, 256, 32, dtype=np.uint8).tobytes(): int(rng.integers(1, 2**60)).to_bytes(32, "big")
                                     for _ in range(args.slot_writes)})
            else:
                block[k] = (0, None, {})
        ks = sorted(block)
        m = len(ks)
        bk = np.frombuffer(b"".join(ks), np.uint8).reshape(m, 32)
        ba = np.zeros(m, ACCOUNT_DTYPE)
        bf = np.zeros(m, np.uint8)
        sk, sv, so = [], [], [0]
        for i, k in enumerate(ks):
            fl, a, slots = block[k]
            bf[i] = fl
            if a is not None:
                ba[i] = a
            for s in sorted(slots):
                sk.append(s)
                sv.append(slots[s])
            so.append(len(sk))
        bsk = np.frombuffer(b"".join(sk), np.uint8).reshape(-1, 32) if sk else np.zeros((0, 32), np.uint8)
        bsv = np.frombuffer(b"".join(sv), np.uint8).reshape(-1, 32) if sv else np.zeros((0, 32), np.uint8)
        so_arr = np.array(so, np.uint64)
        if twin is not None:  # the same block from device memory on the twin state: "value" next to the e2e figure
            import torch
            pad = lambda x: x if len(x) else np.zeros((1, 32), np.uint8)
            dev_in = [torch.from_numpy(np.ascontiguousarray(x).view(np.uint8).reshape(-1)).cuda()
                      for x in (bk, ba, bf, pad(bsk), pad(bsv), so_arr)]
            d_root = torch.zeros(32, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            twin.apply_dev(dev_in[0].data_ptr(), dev_in[1].data_ptr(), dev_in[2].data_ptr(), m, dev_in[3].data_ptr(),
                           dev_in[4].data_ptr(), dev_in[5].data_ptr(), len(sk), d_root.data_ptr())
            dev_wall = time.perf_counter() - t0
            if b >= 2:
                lat_dev.append(dev_wall * 1e6)
        l0 = eng.launch_count()
        t0 = time.perf_counter()
        root = ds.apply(bk, ba, bf, bsk, bsv, so_arr)
        wall = time.perf_counter() - t0
        if b >= 2:
            launches.append(eng.launch_count() - l0)
        for k in ks:
            (touched_existing if k in live_base else inserted_accounts).add(k)
        if twin is not None:
            mism_twin += bytes(d_root.cpu().numpy()) != root
        st = eng.last_stats()
        if b >= 2:
            lat.append(wall * 1e6)
            dev_ms.append(st["device_ms"])
            built.append(st["branches_added"])
        dead = {k for k in ks if block[k][0] == 0}
        live = [k for k in live if k not in dead] + [k for k in ks if block[k][0] == EX and block[k][2] and k not in set(live[:0])]
        live = list(dict.fromkeys(live))
        if state is not None:
            for k in ks:
                fl, a, slots = block[k]
                if not fl & EX:
                    state.pop(k, None)
                    continue
                cur_a, cur_s = state.get(k, (a, {}))
                if not fl & UN:
                    cur_a = a
                cur_s = {} if fl & WI else dict(cur_s)
                for s, v in slots.items():
                    if v == bytes(32):
                        cur_s.pop(s, None)
                    else:
                        cur_s[s] = v
                state[k] = (cur_a, cur_s)
            sk2 = sorted(state)
            fk = np.frombuffer(b"".join(sk2), np.uint8).reshape(-1, 32)
            fa = np.zeros(len(sk2), ACCOUNT_DTYPE)
            fsk, fsv, fo = [], [], [0]
            for i, k in enumerate(sk2):
                fa[i] = state[k][0]
                for s in sorted(state[k][1]):
                    fsk.append(s)
                    fsv.append(state[k][1][s])
                fo.append(len(fsk))
            full = eng.state_root_full(fk, fa, np.frombuffer(b"".join(fsk), np.uint8).reshape(-1, 32),
                                       np.frombuffer(b"".join(fsv), np.uint8).reshape(-1, 32), np.array(fo, np.uint64))
            mism += full != root
    # ---- undo everything in one block: every touched base account gets its base value and (wipe +) its base slots back,
    # every account the blocks created is destroyed.  The root must return to the seed root — the root of the from-scratch
    # build the state was created from — without any model of the 17M-leaf state on the host.
    undo_keys = sorted(touched_existing | inserted_accounts)
    m = len(undo_keys)
    bk = np.frombuffer(b"".join(undo_keys), np.uint8).reshape(m, 32)
    ba = np.zeros(m, ACCOUNT_DTYPE)
    bf = np.zeros(m, np.uint8)
    rows, so = [], [0]
    for i, k in enumerate(undo_keys):
        r = live_base.get(k)
        if r is not None:
            ba[i] = accs[r]
            bf[i] = EX | WI
            rows.append(np.arange(int(offs[r]), int(offs[r + 1])))
            so.append(so[-1] + int(offs[r + 1] - offs[r]))
        else:
            so.append(so[-1])
    rr = np.concatenate(rows) if rows else np.zeros(0, np.int64)
    undo_root = ds.apply(bk, ba, bf, skeys[rr], svals[rr], np.array(so, np.uint64))
    undo_ok = bool(undo_root == seed_root and ds.accounts() == n and ds.slots() == total)
    cpu = None
    if args.cpu_sample:
        import oracle
        cs = args.cpu_sample
        threads = len(os.sched_getaffinity(0))
        t0 = time.perf_counter()
        oracle.state_root_full(keys[:cs], accs[:cs], skeys[:int(offs[cs])], svals[:int(offs[cs])], offs[:cs + 1], threads=threads)
        dt = time.perf_counter() - t0
        leaves = cs + int(offs[cs])
        cpu = {"value": leaves / dt, "unit": "leaves/s", "cores": threads, "kind": "port",
               "sample": f"from-scratch ParallelStateRoot-shaped build of {cs} accounts x {args.slots} slots "
                         "(the CPU restatement has no in-place update path)",
               "equivalent_rebuild_s_at_state_size": (n + total) / (leaves / dt)}
    print(json.dumps({"undo_block_restores_seed_root": undo_ok, "launches_per_block": float(np.median(launches)) if launches else None,
                      "cpu_baseline": cpu,
                      "accounts": n, "slots_per_account": args.slots, "touched_accounts_per_block": args.touch,
                      "slot_writes_per_storage_touch": args.slot_writes, "blocks": args.blocks,
                      "apply_wall_us_median": float(np.median(lat)), "apply_device_ms_median": float(np.median(dev_ms)),
                      "rehashed_nodes_median": float(np.median(built)), "seed_build_s": build_s,
                      "accounts_after": ds.accounts(), "slots_after": ds.slots(), "resident_bytes": ds.device_bytes(),
                      "root_mismatches_vs_full_rebuild": mism if state is not None else None,
                      "apply_dev_wall_us_median": float(np.median(lat_dev)) if lat_dev else None,
                      "root_mismatches_dev_vs_host": mism_twin if twin is not None else None}))


if __name__ == "__main__":
    main()
