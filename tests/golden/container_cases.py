"""Byte lists for the container golden vectors: every length width (1..4 bytes), empty items, both head widths,
item counts around the head-byte boundaries."""
import numpy as np


def _blob(rng, n):
    return rng.integers(0, 256, n, dtype=np.uint8).tobytes()


def bytes_lists():
    rng = np.random.default_rng(77)
    cases = {
        'two_small': [b'ab', b'cde'],
        'with_empty': [b'', b'x', b''],
        'len_255_256': [_blob(rng, 255), _blob(rng, 256)],
        'len_65535_65536': [_blob(rng, 65535), _blob(rng, 65536), _blob(rng, 1)],
        'four_byte_length': [_blob(rng, 1 << 24), b'tail'],
    }
    for n in (6, 7, 8, 14, 15, 16, 33):
        cases[f'count_{n}_w1'] = [_blob(rng, int(rng.integers(0, 300))) for _ in range(n)]
    for n in (3, 4, 5, 7, 8, 19):
        cases[f'count_{n}_w2'] = [_blob(rng, int(rng.integers(0, 70000))) for _ in range(n)]
    return cases
