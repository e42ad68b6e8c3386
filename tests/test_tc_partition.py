"""Host-side planner of the tensor-core engine's work partition (csrc/vsf_api.cu
plan_tc_partition, exported as vsf_debug_tc_plan; layout in csrc/vsf_device.cuh TcBatch).  No
GPU needed: the planner is integer arithmetic, and the walk the kernel's CTAs make over their
slot ranges (walk_begin / walk_unit / walk_next in csrc/knn2_tc_kernel.cu) is restated here to
check that every (query block, train tile) pair is computed exactly once and that the partial
slots the refine kernel merges are exactly the ones the distance kernel writes."""
import ctypes as C

import numpy as np
import pytest

import vision_slam_frontend_b200 as vsf


def plan(qblocks, tiles, sm, force=0, rows=0, cap=1 << 40):
    L = vsf.load_library()
    out = (C.c_int * 5)()
    rc = L.vsf_debug_tc_plan(qblocks, tiles, sm, force, rows, cap, out)
    assert rc == 0
    return tuple(out)          # pieces, tiles_per_piece, grid, slots, segments of block 0


def owner(x, T, G):            # tc_owner: the largest i with floor(i*T/G) <= x
    return ((x + 1) * G - 1) // T


def walk(qblocks, tiles, pieces, tpp, G):
    """(cta, block, slot, first tile, end tile) of every segment, in the kernel's order."""
    T = qblocks * pieces
    for i in range(G):
        x, end = i * T // G, (i + 1) * T // G
        first = True
        gqb = t0 = slot = 0
        while x < end:
            if first:
                gqb = x // pieces
                t0 = x - gqb * pieces
                slot = i - owner(gqb * pieces, T, G)
                first = False
            else:
                gqb, t0, slot = gqb + 1, 0, 0
            n = min(pieces - t0, end - x)
            yield i, gqb, slot, min(tiles, t0 * tpp), min(tiles, (t0 + n) * tpp)
            x += n


def check(qblocks, tiles, sm, force=0, rows=0, cap=1 << 40):
    pieces, tpp, G, slots, seg0 = plan(qblocks, tiles, sm, force, rows, cap)
    assert 1 <= pieces <= tiles and pieces * tpp >= tiles and (pieces - 1) * tpp < tiles
    T = qblocks * pieces
    assert 1 <= G <= min(T, max(sm, 1)) or (G == T and qblocks * pieces <= sm)
    cover = np.zeros((qblocks, tiles), np.int32)
    used = [set() for _ in range(qblocks)]
    busy = set()
    for i, b, slot, t0, t1 in walk(qblocks, tiles, pieces, tpp, G):
        assert 0 <= slot < slots, (slot, slots)
        assert slot not in used[b]
        used[b].add(slot)
        cover[b, t0:t1] += 1
        busy.add(i)
    assert (cover == 1).all(), "a (query block, train tile) pair is skipped or computed twice"
    assert len(busy) == G, "a CTA without work"
    for b in range(qblocks):
        nseg = owner(b * pieces + pieces - 1, T, G) - owner(b * pieces, T, G) + 1   # tc_block_segments
        assert used[b] == set(range(nseg)), (b, used[b], nseg)
    assert seg0 == len(used[0])
    if rows:
        assert pieces == 1 or rows * slots * 2 <= cap
    return pieces, tpp, G, slots


def test_plan_c4_is_one_tile_per_piece_on_every_sm():
    pieces, tpp, G, slots = check(200, 20, 148)           # C4: 10 frames x 20 blocks, 20 tiles
    assert (pieces, tpp, G) == (20, 1, 148) and slots == 2


def test_plan_small_launch_is_one_piece_per_cta():
    pieces, tpp, G, slots = check(8, 8, 148)              # C2 on the tensor engine
    assert G == 8 * pieces and slots == pieces
    pieces, tpp, G, slots = check(20, 20, 148)            # 5000 x 5000, one pair
    assert (pieces, tpp, G) == (5, 4, 100)                # about two thirds of the SMs, not all of them
    pieces, tpp, G, slots = check(12, 12, 148)            # 3000 x 3000
    assert (pieces, tpp, G) == (6, 2, 72)


def test_plan_very_large_launch_uses_long_pieces():
    pieces, tpp, G, slots = check(800, 80, 148)           # 20480 features x 10 frames
    assert G == 148 and tpp >= 16 and slots <= 2
    check(2528, 79, 148)                                  # C5: 32 frames x 79 blocks


@pytest.mark.parametrize("force", [1, 2, 3, 5, 8, 32])
def test_plan_forced_pieces(force):
    check(8, 8, 148, force)
    check(200, 20, 148, force)
    check(1, 1, 148, force)


def test_plan_respects_partial_capacity():
    # 40 blocks x 64 tiles on 148 SMs would want many segments per block; a small partial
    # buffer forces coarser pieces
    rows = 40 * 256
    loose = check(40, 64, 148, 0, rows, 1 << 40)
    tight = check(40, 64, 148, 0, rows, rows * 2 * 2)
    assert tight[3] <= 2 <= loose[3]


def test_plan_random_shapes():
    rng = np.random.default_rng(11)
    for _ in range(300):
        qblocks = int(rng.integers(1, 400))
        tiles = int(rng.integers(1, 90))
        sm = int(rng.choice([1, 2, 7, 108, 132, 148, 160]))
        force = int(rng.choice([0, 0, 0, 1, 2, 4, 9]))
        check(qblocks, tiles, sm, force)
