"""GPU parity: batched keccak256 through the C ABI (include/b200trie.h) against the oracle and the
reference's known-answer vectors.  Bit-exact (byte work)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from tests.util import random_keys, sort_rows

H = bytes.fromhex


@pytest.fixture(scope="module")
def eng():
    from reth_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def test_known_answers(eng):
    """KECCAK_EMPTY / EMPTY_ROOT_HASH / HASHED_ZERO_ADDRESS (hashing_storage.rs:34-35) and the address table of
    crates/trie/db/tests/proof.rs:23-29."""
    assert eng.keccak256(b"") == H("c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470")
    assert eng.keccak256(b"\x80") == H("56e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421")
    z = eng.keccak256_fixed(np.zeros((1, 20), np.uint8))
    assert z[0].tobytes() == H("5380c7b7ae81a58eb98d9c78de4a1fd7fd9535fc953ed2be602daaa41767312a")
    table = {
        "2031f89b3ea8014eb51a78c316e42af3e0d7695f": "a711355ec1c8f7e26bb3ccbcb0b75d870d15846c0b98e5cc452db46c37faea40",
        "33f0fc440b8477fcfbe9d0bf8649e7dea9baedb2": "a77d337781e762f3577784bab7491fcc43e291ce5a356b9bc517ac52eed3a37a",
        "62b0dd4aab2b1a0a04e279e2b828791a10755528": "a7f936599f93b769acf90c7178fd2ddcac1b5b4bc9949ee5a04b7e0823c2446e",
        "1ed9b1dd266b607ee278726d324b855a093394a6": "a77d397a32b8ab5eb4b043c65b1f00c93f517bc8883c5cd31baf8e8a279475e3",
        "000d836201318ec6899a67540690382780743280": "cf67b71c90b0d523dd5004cf206f325748da347685071b34812e21801f5270c4",
    }
    addrs = np.frombuffer(b"".join(H(a) for a in table), np.uint8).reshape(-1, 20)
    got = eng.keccak256_fixed(addrs)
    assert [g.tobytes().hex() for g in got] == list(table.values())


@pytest.mark.parametrize("msg_len,stride", [(32, 32), (20, 20), (20, 32), (32, 48), (32, 33), (7, 7), (64, 64),
                                            (135, 136), (136, 136), (137, 140), (300, 300)])
def test_fixed_matches_oracle(eng, msg_len, stride):
    n = 20_011
    msgs = random_keys(msg_len * 1000 + stride, (n * stride + 31) // 32)[:].reshape(-1)[: n * stride].reshape(n, stride)
    got = eng.keccak256_fixed(msgs, msg_len)
    exp = oracle.keccak256_fixed(msgs, msg_len, threads=4)
    assert (got == exp).all()


def test_fixed_large_multi_chunk(eng):
    """> 1 Mi messages exercises the chunked double-buffered host path."""
    n = (1 << 21) + 12345
    keys = random_keys(2, n)
    got = eng.keccak256_fixed(keys)
    idx = np.random.default_rng(0).integers(0, n, 5000)
    exp = oracle.keccak256_fixed(keys[idx], threads=4)
    assert (got[idx] == exp).all()
    assert (got[-1] == oracle.keccak256_fixed(keys[-1:])[0]).all()


def test_var_matches_oracle(eng):
    rng = np.random.default_rng(5)
    lens = np.concatenate([np.arange(0, 300), rng.integers(0, 1200, 700),
                           np.array([135, 136, 137, 271, 272, 273, 407, 408, 409, 543, 544])])
    offs = np.zeros(len(lens) + 1, np.uint64)
    offs[1:] = np.cumsum(lens)
    data = rng.integers(0, 256, int(offs[-1]), dtype=np.uint8)
    got = eng.keccak256_var(data, offs)
    exp = oracle.keccak256_var(data, offs, threads=2)
    assert (got == exp).all()


def test_empty_batches(eng):
    assert eng.keccak256_fixed(np.zeros((0, 32), np.uint8)).shape == (0, 32)
    assert eng.keccak256_var(np.zeros(0, np.uint8), np.zeros(1, np.uint64)).shape == (0, 32)


def test_hash_sort_keys(eng):
    """AccountHashing full pass: digests sorted ascending + permutation (hashing_account.rs:192-230)."""
    n = 300_007
    addrs = random_keys(77, n)[:, :20].copy()
    sorted_d, perm = eng.hash_sort_keys(addrs)
    exp = oracle.keccak256_fixed(addrs, threads=4)
    order = sort_rows(exp)
    assert (sorted_d == exp[order]).all()
    assert (perm.astype(np.int64) == order).all()


def test_sort_keys32_equal_prefix_fallback(eng):
    """Keys sharing their first 8..31 bytes force the full-key LSD fallback."""
    import torch
    rng = np.random.default_rng(9)
    n = 50_000
    keys = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    keys[:, :8] = keys[0, :8]            # identical 64-bit prefix everywhere
    keys[: n // 2, 8:24] = keys[1, 8:24]  # half of them also share the next 16 bytes
    keys = np.unique(keys, axis=0)
    rng.shuffle(keys)
    n = len(keys)
    t = torch.from_numpy(keys).cuda()
    out = torch.empty_like(t)
    perm = torch.empty(n, dtype=torch.int32, device="cuda")
    eng.sort_keys32_dev(t, n, out, perm)
    eng.sync()
    order = sort_rows(keys)
    assert (out.cpu().numpy() == keys[order]).all()
    assert (perm.cpu().numpy().astype(np.int64) == order).all()


class _HostAsDevice:
    """--emu only: the emulation treats host memory as device memory, so a numpy array can stand in for a CUDA tensor."""

    def __init__(self, a):
        self.a = a

    def data_ptr(self):
        return self.a.ctypes.data


def _sort_keys32(eng, keys):
    """b200_sort_keys32_dev over `keys`: CUDA tensors on a GPU, the arrays themselves under --emu."""
    import os
    n = len(keys)
    if os.environ.get("B200_EMU"):
        out, perm = np.empty_like(keys), np.empty(n, np.uint32)
        eng.sort_keys32_dev(_HostAsDevice(keys), n, _HostAsDevice(out), _HostAsDevice(perm))
        eng.sync()
        return out, perm
    import torch as th
    t = th.from_numpy(keys).cuda()
    d_out, d_perm = th.empty_like(t), th.empty(n, dtype=th.int32, device="cuda")
    eng.sort_keys32_dev(t, n, d_out, d_perm)
    eng.sync()
    return d_out.cpu().numpy(), d_perm.cpu().numpy()


def test_sort_keys32_runs_of_equal_leading_bytes(eng):
    """The sort orders by the top 32 bits, then the head of every run of equal tops orders its run by all 32 bytes
    (hash_sort.cu fix_runs_kernel): runs of 2 .. 16 rows (in place), 17 and more (left to the full four-word sort), runs at
    both ends of the batch, rows that differ only in their last byte, and a crowd of ordinary rows around them."""
    rng = np.random.default_rng(41)
    groups = []
    for ln in list(range(2, 20)) + [40, 2, 16, 17]:
        g = rng.integers(0, 256, (ln, 32), dtype=np.uint8)
        g[:, :4] = rng.integers(0, 256, 4, dtype=np.uint8)
        if ln % 3 == 0:
            g[:, 4:31] = g[0, 4:31]                      # equal up to the last byte
            g[:, 31] = rng.permutation(256)[:ln].astype(np.uint8)
        groups.append(g)
    lo = np.zeros((3, 32), np.uint8)
    lo[:, 31] = [3, 1, 2]                                 # a run at the very start of the sorted order
    hi = np.full((2, 32), 255, np.uint8)
    hi[:, 31] = [254, 9]                                  # and one at the very end
    keys = np.concatenate(groups + [lo, hi, rng.integers(0, 256, (20_000, 32), dtype=np.uint8)])
    keys = np.unique(keys, axis=0)
    rng.shuffle(keys)
    order = sort_rows(keys)
    out, perm = _sort_keys32(eng, keys)
    assert (out == keys[order]).all()
    assert (perm.astype(np.int64) == order).all()
    # without the long runs the in-place path alone must get it right (no fallback: the launch count tells)
    short = np.concatenate([g for g in groups if len(g) <= 16] + [lo, hi, rng.integers(0, 256, (5_000, 32), dtype=np.uint8)])
    short = np.unique(short, axis=0)
    rng.shuffle(short)
    l0 = eng.launch_count()
    out, perm = _sort_keys32(eng, short)
    assert eng.launch_count() - l0 == 5, "the four-word fallback ran"
    order = sort_rows(short)
    assert (out == short[order]).all() and (perm.astype(np.int64) == order).all()


def test_device_resident_path(eng):
    import torch
    n = 100_000
    keys = random_keys(21, n)
    t = torch.from_numpy(keys).cuda()
    out = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
    eng.use_torch_stream()
    eng.keccak256_fixed_dev(t, 32, 32, n, out)
    torch.cuda.synchronize()
    eng.set_stream(None)
    assert (out.cpu().numpy() == oracle.keccak256_fixed(keys, threads=4)).all()


def test_hash_sort_storage_composite(eng):
    """StorageHashing full pass (hashing_storage.rs:106-178): rows sorted by keccak(address) || keccak(slot)."""
    rng = np.random.default_rng(3)
    n_addr, n = 5000, 120_000
    addrs = random_keys(5, n_addr)[:, :20].copy()
    owner = rng.integers(0, n_addr, n).astype(np.uint32)
    owner[:40_000] = 7                      # one contract with a large storage: identical address prefix everywhere
    slots = random_keys(6, n)
    keys, perm = eng.hash_sort_storage(addrs, owner, slots)
    ha = oracle.keccak256_fixed(addrs, threads=4)
    hs = oracle.keccak256_fixed(slots, threads=4)
    comp = np.concatenate([ha[owner], hs], axis=1)
    v = comp.view(">u8")
    order = np.lexsort(tuple(v[:, i] for i in range(7, -1, -1)))
    assert (keys == comp[order]).all()
    assert (perm.astype(np.int64) == order).all()


def test_hash_sort_storage_large_storage_and_repeated_addresses(eng):
    """The composite sort orders by (dense rank of the address digest, top 32 bits of the slot digest) and then orders the rows
    that agree in both in place (hash_sort.cu fix_runs_composite_kernel): one contract with 400k slots has ~18 pairs of slot
    digests with equal top 32 bits; the address table lists one address twice (equal digests share a rank) and holds
    addresses no entry refers to."""
    rng = np.random.default_rng(12)
    n_addr, n = 300, 420_000
    addrs = random_keys(31, n_addr)[:, :20].copy()
    addrs[17] = addrs[5]                                    # the same address under two indices
    owner = rng.integers(100, 200, n).astype(np.uint32)     # indices 0..99 and 200..299 are never used
    owner[:400_000] = 150
    owner[400_000:400_500] = 5
    owner[400_500:401_000] = 17
    slots = random_keys(32, n)
    keys, perm = eng.hash_sort_storage(addrs, owner, slots)
    ha = oracle.keccak256_fixed(addrs, threads=4)
    hs = oracle.keccak256_fixed(slots, threads=4)
    comp = np.concatenate([ha[owner], hs], axis=1)
    v = comp.view(">u8")
    order = np.lexsort(tuple(v[:, i] for i in range(7, -1, -1)))
    top = comp[order][:, :36]
    assert (np.all(top[1:] == top[:-1], axis=1)).sum() >= 5, "the input was meant to hold rows with equal (address, slot top)"
    assert (keys == comp[order]).all()
    assert (perm.astype(np.int64) == order).all()


def test_hash_sort_storage_rejects_duplicates_and_bad_index(eng):
    from reth_b200 import B200Error, _lib
    addrs = random_keys(5, 4)[:, :20].copy()
    slots = random_keys(6, 8)
    slots[5] = slots[2]
    owner = np.array([0, 1, 2, 3, 0, 2, 1, 3], np.uint32)
    with pytest.raises(B200Error) as e:
        eng.hash_sort_storage(addrs, owner, slots)
    assert e.value.status == _lib.ERR_UNSORTED
    owner[0] = 9
    with pytest.raises(B200Error) as e:
        eng.hash_sort_storage(addrs, owner, random_keys(7, 8))
    assert e.value.status == _lib.ERR_INVALID_ARG


@pytest.mark.gpu
def test_hash_sort_host_paths_in_chunks():
    """The host-pointer hash+sort entry points copy and hash in chunks on the copy streams (eng_keccak.inl hash_from_host):
    several chunks incl. a ragged last one, results in the caller's page-locked arrays, against the oracle.  Own process:
    the chunk size is read once per process (B200_KECCAK_CHUNK)."""
    import subprocess
    import sys
    from reth_b200 import _lib
    code = "from reth_b200 import _lib\n_lib.LIB_PATH = %r   # (the emulation build under --emu)\n" % _lib.LIB_PATH + r"""
import numpy as np, oracle
from reth_b200 import Engine, B200Error
from tests.util import random_keys, sort_rows
eng = Engine(0)
n = 5 * 1024 + 77
addrs = random_keys(21, n)[:, :20].copy()
out, perm = eng.pinned_empty((n, 32)), eng.pinned_empty((n,), np.uint32)
r_out, r_perm = eng.hash_sort_keys(addrs, out=out, perm=perm)
assert r_out is out and r_perm is perm
dig = oracle.keccak256_fixed(addrs)
assert (out == dig[sort_rows(dig)]).all() and (dig[perm] == out).all()
# strided input: 20-byte messages at a 32-byte stride
wide = random_keys(22, n)
out2, perm2 = eng.hash_sort_keys(wide, 20)
dig2 = oracle.keccak256_fixed(np.ascontiguousarray(wide[:, :20]))
assert (out2 == dig2[sort_rows(dig2)]).all() and (dig2[perm2] == out2).all()
# storage: composite keys, slots hashed in chunks
n_addr = 9
a20 = random_keys(23, n_addr)[:, :20].copy()
owner = np.sort(np.random.default_rng(5).integers(0, n_addr, n)).astype(np.uint32)
slots = random_keys(24, n)
keys, sperm = eng.hash_sort_storage(a20, owner, slots)
comp = np.concatenate([oracle.keccak256_fixed(a20)[owner], oracle.keccak256_fixed(slots)], axis=1)
v = comp.view(">u8")
order = np.lexsort(tuple(v[:, i] for i in range(7, -1, -1)))
assert (keys == comp[order]).all() and (sperm == order).all()
try:
    eng.hash_sort_keys(addrs, out=np.empty((n, 31), np.uint8))
    raise SystemExit("a mis-shaped result array was accepted")
except B200Error:
    pass
eng.close()
print("ok")
"""
    import os
    env = dict(os.environ, B200_KECCAK_CHUNK="1024")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
