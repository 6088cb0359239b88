"""Kernel logic without a GPU: the CUDA sources of reth_b200/csrc, translated and run by tools/emu (fibers for the
threads of a block, yield-based barriers and warp collectives), pass the fast part of the `gpu` parity tests
bit-exact against the oracle and the golden vectors.

This is a check of the *sources* (indexing, masks, RLP assembly, barrier placement), not a product path: only this
test process points the loader at the emulated build (tests/conftest.py --emu); the B200 results come from the
`-m gpu` run of the very same tests on the real library."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAST = ("not one_million and not pipelined and not reentrancy and not frontier_sharding and not commits_blocks "
        "and not 200000")


def test_cuda_sources_pass_parity_under_cpu_emulation():
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_trie.py", "tests/test_gpu_keccak.py",
                        "tests/test_gpu_host_mirror.py", "tests/test_gpu_dtrie.py", "-m", "gpu", "--emu", "-q", "-x", "-k", FAST,
                        "-p", "no:cacheprovider"],
                       cwd=ROOT, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
