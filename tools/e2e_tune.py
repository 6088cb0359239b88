"""Times b200_keccak256_fixed (host pointers) for the chunk size given by B200_KECCAK_CHUNK."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from reth_b200 import Engine
eng = Engine(0)
n = 10_000_000
h_in = eng.pinned_empty((n, 32)); h_out = eng.pinned_empty((n, 32))
h_in[:] = np.random.default_rng(1).integers(0, 256, (n, 32), dtype=np.uint8)
for _ in range(3): eng.keccak256_fixed(h_in, 32, out=h_out)
t0 = time.perf_counter()
for _ in range(10): eng.keccak256_fixed(h_in, 32, out=h_out)
dt = (time.perf_counter() - t0) / 10
print(os.environ.get("B200_KECCAK_CHUNK"), f"{dt*1e3:.3f} ms  {n/dt/1e9:.3f} G digests/s  {640e6/dt/1e9:.1f} GB/s PCIe aggregate")
