#!/usr/bin/env python3
"""Round-2 measurement aid for b200_state_root_full_rows (table rows laid out on the device).

    python tools/rows_bench.py --accounts 1000000 --slots 16

Builds a synthetic state, then times (wall, median of --reps) the root + AccountsTrie / StoragesTrie rows through
(a) the device encoder and (b) records + the host encoder, checks that both yield the same bytes, and prints one JSON line."""
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
    ap.add_argument("--packed", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    from reth_b200 import Engine
    from tests.util import synth_accounts, synth_storage
    akeys, accs = synth_accounts(3, args.accounts)
    skeys, svals, offs = synth_storage(4, np.full(args.accounts, args.slots, np.int64))
    eng = Engine(0)
    # the caller's buffers page-locked (b200_host_alloc), as in bench.py's end-to-end legs: the 1.1 GB H2D is a plain DMA
    pinned = []
    for a in (akeys, accs, skeys, svals, offs):
        a = np.ascontiguousarray(a)
        b = eng.pinned_empty(a.shape, a.dtype)
        b[...] = a
        pinned.append(b)
    akeys, accs, skeys, svals, offs = pinned
    res = {"host_buffers": "page-locked"}
    same = True
    keep = None
    for name, host in (("device", False), ("host", True)):
        ts = []
        for _ in range(args.reps + 1):
            t0 = time.perf_counter()
            root, ar, sr = eng.state_root_full_rows(akeys, accs, skeys, svals, offs, key_format=args.packed, encode_on_host=host)
            ts.append((time.perf_counter() - t0) * 1e3)   # (the rows live in page-locked library memory until release())
            cur = (root, ar.bytes.copy(), sr.bytes.copy(), ar.row_offset.copy(), sr.row_offset.copy())
            rows, nbytes = len(ar) + len(sr), int(ar.row_offset[-1]) + int(sr.row_offset[-1])
            ar.release(); sr.release()
        res[name + "_ms"] = round(float(np.median(ts[1:])), 2)
        if keep is None:
            keep = cur
        else:
            same = cur[0] == keep[0] and all((a == b).all() for a, b in zip(cur[1:], keep[1:]))
    t0 = time.perf_counter()
    eng.state_root_full(akeys, accs, skeys, svals, offs)
    t0 = time.perf_counter()
    eng.state_root_full(akeys, accs, skeys, svals, offs)
    res["root_only_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
    print(json.dumps({"tool": "rows_bench", "accounts": args.accounts, "slots_per_account": args.slots, "rows": rows,
                      "row_bytes": nbytes, **res, "device_equals_host": bool(same)}))
    eng.close()
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
