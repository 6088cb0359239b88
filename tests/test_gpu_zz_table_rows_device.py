"""GPU parity of b200_state_root_full_rows: AccountsTrie / StoragesTrie rows sized, ordered and laid out on
the device, against rows encoded from the oracle's TrieUpdates by the test-side restatement of the reference codecs
(tests/test_table_rows.py) and against the host encoder over the records (b200_account_trie_rows /
b200_storage_trie_rows).  Byte-exact.  (Named to run last: first validated under tools/emu)"""
import numpy as np
import pytest

# first hardware run happens at round end: bound a hang (method=thread ends the process even inside a blocked CUDA sync)
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]

import oracle
from tests.test_table_rows import expected_account_rows, expected_storage_rows
from tests.util import synth_accounts, synth_storage


@pytest.fixture(scope="module")
def eng():
    from reth_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("packed", [0, 1])
def test_device_rows_match_reference_codecs_and_host_encoder(eng, packed):
    n = 20_000
    akeys, accs = synth_accounts(21, n)
    counts = np.where(np.arange(n) % 4 == 0, 24, 0) + np.where(np.arange(n) % 1999 == 0, 2500, 0)
    skeys, svals, offs = synth_storage(22, counts, value_mode="mixed")
    root, arows, srows = eng.state_root_full_rows(akeys, accs, skeys, svals, offs, key_format=packed)
    o_root, o_au, o_su = oracle.state_root_full(akeys, accs, skeys, svals, offs, want_updates=True, threads=4)
    assert root == o_root
    assert arows.to_list() == expected_account_rows(o_au, bool(packed))
    assert srows.to_list() == expected_storage_rows(o_su, akeys, bool(packed))
    assert len(arows) > 1000 and len(srows) > 1000
    root2, arows2, srows2 = eng.state_root_full_rows(akeys, accs, skeys, svals, offs, key_format=packed, encode_on_host=True)
    assert root2 == root
    for dev_rows, host_rows in ((arows, arows2), (srows, srows2)):
        assert (dev_rows.row_offset == host_rows.row_offset).all() and (dev_rows.key_len == host_rows.key_len).all()
        assert (dev_rows.bytes == host_rows.bytes).all()
    arows.release(); srows.release(); arows2.release(); srows2.release()


@pytest.mark.parametrize("packed", [0, 1])
def test_device_rows_of_degenerate_states(eng, packed):
    """no stored nodes at all (tiny tries), no storage, a single account"""
    for n, per in ((0, 0), (1, 0), (3, 2), (40, 0), (300, 1)):
        akeys, accs = synth_accounts(41 + n, n)
        skeys, svals, offs = synth_storage(42 + n, np.full(n, per, np.int64))
        root, arows, srows = eng.state_root_full_rows(akeys, accs, skeys, svals, offs, key_format=packed)
        o_root, o_au, o_su = oracle.state_root_full(akeys, accs, skeys, svals, offs, want_updates=True)
        assert root == o_root
        assert arows.to_list() == expected_account_rows(o_au, bool(packed))
        assert srows.to_list() == expected_storage_rows(o_su, akeys, bool(packed))
        arows.release(); srows.release()


def test_device_rows_of_the_mainnet_genesis_trie(eng, golden_allocs):
    from tests.util import alloc_to_flat
    g = golden_allocs["mainnet"]
    flat = alloc_to_flat(g["alloc"])
    root, arows, srows = eng.state_root_full_rows(*flat, key_format=0)
    assert root.hex() == g["state_root"]
    o_root, o_au, o_su = oracle.state_root_full(*flat, want_updates=True)
    assert arows.to_list() == expected_account_rows(o_au, False) and len(arows) > 100
    assert srows.to_list() == expected_storage_rows(o_su, flat[0], False)
    arows.release(); srows.release()


def test_rejects_bad_key_format(eng):
    from reth_b200 import B200Error
    akeys, accs = synth_accounts(5, 10)
    skeys, svals, offs = synth_storage(6, np.zeros(10, np.int64))
    with pytest.raises(B200Error):
        eng.state_root_full_rows(akeys, accs, skeys, svals, offs, key_format=7)
