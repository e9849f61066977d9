"""Host-side containers of the C ABI (frame header, partition container, BytesListUtils) against byte strings minted
by the reference's own BytesListUtils (tests/golden/make_container_golden.py) and against the layouts written at
lossl_coord_int/model.py:447-463.  No GPU needed: these entry points touch host memory only."""
import hashlib
import io
import json
import os.path as osp

import pytest

from fastpcc_b200 import bitstream as B
from tests.golden.container_cases import bytes_lists

GOLDEN = json.load(open(osp.join(osp.dirname(__file__), 'golden', 'container_golden.json')))
CASES = bytes_lists()


@pytest.mark.parametrize('name', list(CASES))
def test_bytes_list_concat_equals_reference_and_splits_back(name):
    items, gold = CASES[name], GOLDEN[name]
    blob = B.BytesListUtils.concat_bytes_list(items)
    assert len(blob) == gold['len'] and blob[:64].hex() == gold['head_hex']
    assert hashlib.sha256(blob).hexdigest() == gold['sha256']
    assert B.BytesListUtils.split_bytes_list(blob, len(items)) == items
    # stream form: writes into / reads from a BytesIO and leaves the cursor after the last item
    bs = io.BytesIO()
    bs.write(b'HDR')
    assert B.BytesListUtils.concat_bytes_list(items, bs) is None
    bs.write(b'TAIL')
    bs.seek(3)
    assert B.BytesListUtils.split_bytes_list(None, len(items), bs) == items
    assert bs.read() == b'TAIL'


def test_bytes_list_errors():
    with pytest.raises(AssertionError):
        B.BytesListUtils.concat_bytes_list([b'only one'])
    blob = B.BytesListUtils.concat_bytes_list([b'abc', b'defg'])
    with pytest.raises(RuntimeError):
        B.BytesListUtils.split_bytes_list(blob[:-2], 2)   # truncated payload
    with pytest.raises(RuntimeError):
        B.BytesListUtils.split_bytes_list(b'\x00\x00', 2)  # no marker bit


def test_frame_header_layout():
    h = B.write_frame_header([5, 300, 65535], 513)
    assert h == (5).to_bytes(2, 'little') + (300).to_bytes(2, 'little') + (65535).to_bytes(2, 'little') + (513).to_bytes(2, 'little')
    off, n, payload = B.read_frame_header(h + b'payload')
    assert off == [5, 300, 65535] and n == 513 and payload == b'payload'
    with pytest.raises(RuntimeError):
        B.write_frame_header([0, 0, 65536], 1)
    with pytest.raises(RuntimeError):
        B.read_frame_header(b'short')


def test_partition_container_layout():
    parts = [b'', b'a' * 5, bytes(range(256)) * 300]
    blob = B.pack_partitions(parts)
    assert blob == b''.join(len(s).to_bytes(3, 'little') + s for s in parts)  # model.py:462
    assert B.split_partitions(blob) == parts
    assert B.split_partitions(b'') == [] and B.pack_partitions([]) == b''
    with pytest.raises(RuntimeError):
        B.split_partitions(blob[:-1])
    with pytest.raises(RuntimeError):
        B.split_partitions(blob + b'\x01')
