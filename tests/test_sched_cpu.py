"""CPU tests of the split-KV tile schedule (DESIGN 3.2 "dynamic tail") through the host-only hook b200_debug_sched:
the chunks b200_attn_plan writes are a partition of the tile sequence into consecutive ranges, the closed-form maps
tile -> chunk and chunk -> first tile agree, and the chunk count stays inside what the workspace is sized for -- the
invariants the (chunk + request) partial slots and the merge's lookup rest on.  No device needed."""
import ctypes as C
import random

import pytest

from scratchpad_b200 import _lib


def sched(total, grid, min_base=0, probe=-1):
    out = (C.c_int32 * 8)()
    assert _lib.load().b200_debug_sched(total, grid, min_base, probe, out) == _lib.OK
    return list(out)


def chunks(total, grid, min_base):
    """(first tile, end tile) of every chunk, read back tile by tile at the chunk boundaries."""
    s_tiles, d_tiles, dyn0, nch, cap = sched(total, grid, min_base)[:5]
    out, t = [], 0
    while t < total:
        c, begin, nxt = sched(total, grid, min_base, probe=t)[5:8]
        assert begin == t and c == len(out), (total, grid, t, c, begin)
        out.append((begin, min(nxt, total)))
        t = min(nxt, total)
    assert len(out) == nch <= cap, (total, grid, len(out), nch, cap)
    return out, (s_tiles, d_tiles, dyn0, nch, cap)


@pytest.mark.parametrize("grid", [1, 37, 148, 296])
def test_chunks_partition_the_tiles_and_fit_the_workspace(grid):
    rng = random.Random(grid)
    totals = [0, 1, grid - 1, grid, grid + 1, 7 * grid + 3, 8 * grid, 255 * grid + 1, 256 * grid, 443 * grid + 17, 2000 * grid + 5]
    totals += [rng.randrange(1, 3000 * grid) for _ in range(8)]
    for total in totals:
        for min_base in (8, 0):  # forced onto small launches (tests) / the library's threshold
            if total > 40000 and min_base == 8:
                continue  # keep the walk short
            ch, (s_tiles, d_tiles, dyn0, nch, cap) = chunks(total, grid, min_base)
            if total == 0:
                assert nch == 0
                continue
            base = -(-total // grid)
            dynamic = nch > grid
            assert dynamic == (base >= (min_base or 256)) or not dynamic, (total, grid, min_base, nch)
            # static chunks are equal, dynamic ones equal and small, only the last chunk may be short
            for i, (b, e) in enumerate(ch):
                want = s_tiles if i < grid else d_tiles
                assert e - b == want or i == len(ch) - 1, (total, grid, i, b, e, want)
            if dynamic:
                assert dyn0 == grid * s_tiles and s_tiles == base - base // 8 and d_tiles == -(-base // 32)
                assert nch - grid <= 4 * grid + 1
            else:
                assert s_tiles == base and nch <= grid


def test_every_tile_maps_into_its_chunk():
    for total, grid, min_base in ((1000, 37, 8), (70000, 148, 0), (5000, 296, 8)):
        ch, _ = chunks(total, grid, min_base)
        rng = random.Random(total)
        for _ in range(300):
            t = rng.randrange(total)
            c, begin, nxt = sched(total, grid, min_base, probe=t)[5:8]
            assert ch[c][0] == begin <= t < min(nxt, total) == ch[c][1]


def test_threshold_and_argument_checks():
    assert sched(256 * 148, 148)[3] > 148        # 256 tiles per CTA: the tail is dynamic
    assert sched(255 * 148, 148)[3] == 148       # just below: one static range per CTA
    out = (C.c_int32 * 8)()
    lib = _lib.load()
    assert lib.b200_debug_sched(10, 0, 0, -1, out) == _lib.ERR_INVALID
    assert lib.b200_debug_sched(-1, 4, 0, -1, out) == _lib.ERR_INVALID
    assert lib.b200_debug_sched(10, 4, 0, -1, None) == _lib.ERR_INVALID
