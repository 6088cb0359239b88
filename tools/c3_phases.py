#!/usr/bin/env python
"""C3 build (1M accounts x 16 slots) timed end to end with CUDA events, and — with B200_PHASE_TIMING=1 — per phase
(the engine prints `[b200 phases] ...` on stderr at every sync).   python tools/c3_phases.py [--accounts N] [--reps R]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import bench
from reth_b200 import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--accounts", type=int, default=1_000_000)
ap.add_argument("--reps", type=int, default=6)
args = ap.parse_args()
dev = torch.device("cuda", 0)
eng = Engine(0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
eng.use_torch_stream()
sh = bench.make_c3_shard(3, args.accounts, 16, 0, 16, dev)
d_root = torch.zeros(32, dtype=torch.uint8, device=dev)
ms = []
for it in range(args.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.state_root_full_dev(sh["akeys"], sh["accts"], args.accounts, sh["skeys"], sh["svals"], sh["offs"], sh["n_slots"], d_root)
    e1.record()
    torch.cuda.synchronize()
    eng.dev_status()
    ms.append(e0.elapsed_time(e1))
print("ms per build:", [round(x, 3) for x in ms], "root", bytes(d_root.cpu().numpy()).hex(), "launches", eng.launch_count(), flush=True)
