"""CPU model of K2's cooperative match copy (lz4b200_kernels.cuh::copy_match<G>): G lanes reproduce the byte-serial LZ77
copy `for i in 0..len: out[op+i] = out[op-dist+i]` (duplicate / duplicate_overlapping, reference
src/block/decompress.rs:11-82; offset 1 = run fill, decompress_safe.rs:311-313) in three regimes —
source older than the match, overlap with distance >= G (G-byte steps separated by a group barrier), overlap with
distance < G (every byte is a copy of the seed period).  Pins the lane arithmetic independently of CUDA."""
import numpy as np
import pytest


def serial(buf, op, dist, n):
    b = bytearray(buf)
    for i in range(n):
        b[op + i] = b[op - dist + i]
    return bytes(b)


def group_copy(buf, op, dist, n, G):
    b = bytearray(buf)
    frm = op - dist
    if dist >= n:                                    # all lanes load, then store: no lane reads a byte of this match
        vals = {i: b[frm + i] for i in range(n)}
        for i, v in vals.items():
            b[op + i] = v
    elif dist >= G:                                  # step k only needs steps < k: barrier between steps
        for base in range(0, n, G):
            vals = {base + sub: b[frm + base + sub] for sub in range(G) if base + sub < n}   # loads of one step
            for i, v in vals.items():
                b[op + i] = v
    else:                                            # period < G: byte i = seed[i mod dist], r advanced by G mod dist
        seed = bytes(b[frm: frm + dist])
        for sub in range(G):
            r, step = sub % dist, G % dist
            for i in range(sub, n, G):
                b[op + i] = seed[r]
                r += step
                if r >= dist:
                    r -= dist
    return bytes(b)


@pytest.mark.parametrize("G", [4, 8, 16, 32])
def test_group_copy_equals_byte_serial(G):
    rng = np.random.default_rng(G)
    base = rng.integers(0, 256, 400, dtype=np.uint8).tobytes()
    for dist in list(range(1, 41)) + [63, 64, 65, 100]:
        for n in list(range(1, 70)) + [100, 129, 200]:
            op = 120
            assert group_copy(base, op, dist, n, G) == serial(base, op, dist, n), (G, dist, n)
