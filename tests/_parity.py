"""Block-by-block comparison of compressed output with the CPU checker (test infrastructure).

The checker is the unmodified reference (oracle/_ref, scalar CompressMulti<K>) when it was built,
else the oracle restatement; blocks are spread over the host cores (ctypes releases the GIL)."""
import hashlib
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np


def checker():
    """(name, compress(k, bytes-like) -> bytes)."""
    from _libs import Oracle, Ref, have_ref
    if have_ref():
        r = Ref()
        return "reference scalar CompressMulti (oracle/_ref)", lambda k, raw: r.compress(k, raw)
    o = Oracle()
    return "oracle restatement (oracle/huf_oracle.c)", lambda k, raw: o.compress(k, raw)


def compare_blocks(raw, comp, offsets, sizes, k, block_size, blocks=None, threads=None):
    """raw: uint8 numpy array of the whole input; comp/offsets/sizes: our compressed blocks (numpy).
    Compares every block in `blocks` (default: all) by SHA-256 with the checker's output.
    Returns (checker name, number of blocks compared, list of mismatching block indices)."""
    name, comp_fn = checker()
    n = raw.size
    nb = (n + block_size - 1) // block_size
    blocks = list(range(nb)) if blocks is None else [int(b) for b in blocks]

    def one(b):
        lo = b * block_size
        want = comp_fn(k, raw[lo: min(n, lo + block_size)])
        o, s = int(offsets[b]), int(sizes[b])
        got = comp[o: o + s]
        if len(want) != s:
            return b
        return None if hashlib.sha256(want).digest() == hashlib.sha256(got).digest() else b

    with ThreadPoolExecutor(threads or os.cpu_count() or 1) as ex:
        bad = [b for b in ex.map(one, blocks) if b is not None]
    return name, len(blocks), bad
