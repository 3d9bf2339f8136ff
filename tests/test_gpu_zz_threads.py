"""Threading model of the library on the GPU: one context per thread, contexts independent (include/lz4b200.h).
Kept in its own file, collected last, so that the rest of the GPU suite has run when this one starts."""
import hashlib
import os

import numpy as np
import pytest

import oracle
from lz4_flex_b200 import block, corpus

pytestmark = pytest.mark.gpu


def test_contexts_are_independent_across_threads():
    """One context per thread is the library's threading model (include/lz4b200.h): two threads compress their own batches
    on their own contexts at the same time while the main thread decompresses finished ones on a third — what bench.py's
    stream-of-batches e2e mode does.  Every compressed batch must equal the oracle's bytes and decode to its input."""
    import threading
    import queue
    nb = 1024
    srcs = [corpus.tiled("compression_66k_JSON.txt", nb * 65536), corpus.tiled("dickens.txt", nb * 65536)]
    offs = np.arange(nb, dtype=np.uint64) * 65536
    lens = np.full(nb, 65536, dtype=np.uint32)
    want = []
    for s in srcs:
        w = np.zeros(nb * 72112, dtype=np.uint8)
        wl, _ = oracle.compress_batch(s, offs, lens, w, np.arange(nb, dtype=np.uint64) * 72112, np.full(nb, 72112, dtype=np.uint32), os.cpu_count())
        want.append(hashlib.sha256(np.concatenate([w[b * 72112: b * 72112 + int(wl[b])] for b in range(nb)]).tobytes()).hexdigest())
    done, err = queue.Queue(), []

    def worker(i):
        c = block.Context(0)
        try:
            for _ in range(6):
                out, ooff, olen = block.compress_batch(srcs[i], offs, lens, ctx=c)
                done.put((i, out, ooff, olen))
        except Exception as e:                              # noqa: BLE001
            err.append(e)
            done.put(None)
        finally:
            c.close()

    ths = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in ths:
        t.start()
    dctx = block.Context(0)
    back = np.zeros(nb * 65536, dtype=np.uint8)
    for _ in range(12):
        item = done.get()
        assert item is not None, err
        i, out, ooff, olen = item
        used = int(ooff[-1]) + int(olen[-1])
        assert hashlib.sha256(out[:used].tobytes()).hexdigest() == want[i]
        block.decompress_batch(out, ooff, olen, back, offs, lens, ctx=dctx)
        assert np.array_equal(back, srcs[i])
    for t in ths:
        t.join()
    dctx.close()
    assert not err
