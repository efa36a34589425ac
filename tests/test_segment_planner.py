"""Host logic of the two-step sweep: the x-segment planner (csrc/engine_launch.inl: plan_segments), called through
the C ABI's host-only entry fdtd_plan_segments — no GPU needed.  Invariants the kernel relies on:
  * segments partition [0, nx) exactly;
  * a segment [a, b) is flagged `ops` iff a source / monitor plane lies in [a, b + 1] (the planes on which the sweep
    applies the intermediate step's ops for that segment) — a missed flag would silently drop a source;
  * at most 48 segments (FusedTiling::seg_lo / seg_hi), op mask fits 64 bits;
  * on a slab that reads ghost planes, the ghost-reading segment is never dispatched first (when there is a choice).
"""
import ctypes as C

import numpy as np
import pytest

from prismo_b200 import _lib

MAX_SEGS = 48


def plan(nx, op_planes=(), n_flags=None, tiles=1548, halo=False, fused_lx=0, zones=-1):
    lib = _lib.load()
    n_flags = (nx + 4 if n_flags is None else n_flags) if len(op_planes) else 0
    flags = np.zeros(max(n_flags, 1), dtype=np.uint8)
    for p in op_planes:
        flags[p] = 1
    lo = (C.c_int32 * MAX_SEGS)()
    hi = (C.c_int32 * MAX_SEGS)()
    ops = (C.c_int32 * MAX_SEGS)()
    n = lib.fdtd_plan_segments(nx, flags.ctypes.data_as(C.POINTER(C.c_uint8)), n_flags, tiles, int(halo), fused_lx, zones,
                               lo, hi, ops, MAX_SEGS)
    assert n > 0, _lib.last_error() if hasattr(_lib, "last_error") else n
    return [(lo[i], hi[i], bool(ops[i])) for i in range(n)], flags[:n_flags]


def check_invariants(nx, segs, flags):
    assert len(segs) <= MAX_SEGS
    cover = sorted((a, b) for a, b, _ in segs)
    assert cover[0][0] == 0 and cover[-1][1] == nx
    for (a0, b0), (a1, b1) in zip(cover, cover[1:]):
        assert b0 == a1 and a0 < b0
    for a, b, ops in segs:
        want = bool(flags[a:b + 2].any()) if len(flags) else False
        assert ops == want, (a, b, ops, want)


@pytest.mark.parametrize("nx", [1, 2, 5, 8, 17, 64, 128, 129, 256, 500, 1024, 1500])
@pytest.mark.parametrize("tiles", [1, 9, 148, 1548, 5000])
def test_partition_without_ops(nx, tiles):
    segs, flags = plan(nx, tiles=tiles)
    check_invariants(nx, segs, flags)
    assert not any(o for _, _, o in segs)


def test_random_op_planes_all_modes():
    rng = np.random.default_rng(0)
    for trial in range(300):
        nx = int(rng.integers(1, 1200))
        n_ops = int(rng.integers(1, 6)) if trial % 7 else int(rng.integers(20, 60))
        planes = sorted(set(int(p) for p in rng.integers(0, nx + 4, size=n_ops)))
        for zones in (-1, 0, 1):
            for halo in (False, True):
                lx = int(rng.choice([0, 0, 13, 64, 300]))
                segs, flags = plan(nx, planes, tiles=int(rng.choice([3, 200, 1548])), halo=halo, fused_lx=lx, zones=zones)
                check_invariants(nx, segs, flags)


def test_c4_shape_gets_narrow_zones_and_four_bulk_parts():
    # 1024 planes, source plane 256, monitor plane 768 (bench.py c4), 86 x 18 tiles
    segs, flags = plan(1024, (256, 768))
    check_invariants(1024, segs, flags)
    zones = [(a, b) for a, b, o in segs if o]
    bulk = [(a, b) for a, b, o in segs if not o]
    assert zones == [(254, 262), (766, 774)]
    assert len(bulk) == 4 and max(b - a for a, b in bulk) <= 256
    assert [o for _, _, o in segs] == [False] * 4 + [True] * 2          # short op zones fill the tail


def test_short_slab_keeps_two_plain_segments():
    # 128-plane slab (8-GPU share of 1024^3): zones would cost more prologue planes than the op path costs
    segs, flags = plan(128, (32,), halo=True)
    check_invariants(128, segs, flags)
    assert sorted((a, b) for a, b, _ in segs) == [(0, 64), (64, 128)]


def test_ghost_reader_is_not_dispatched_first():
    for nx, planes in ((128, ()), (128, (0,)), (128, (130,)), (64, ()), (16, ()), (512, (3,)), (1024, (256, 768))):
        segs, flags = plan(nx, planes, halo=True)
        check_invariants(nx, segs, flags)
        if len(segs) > 1:
            assert segs[0][1] + 3 < nx, (nx, planes, segs)
    # a single-segment slab is cut in two so that only half of the CTAs can spin on the neighbour's push
    segs, _ = plan(100, (), tiles=100000, halo=True)
    assert len(segs) >= 2


def test_ghost_plane_source_marks_last_segment():
    # the right neighbour's source on our ghost plane nx+1 must put the segment that ends at nx on the op path
    segs, flags = plan(128, (129,), halo=True)
    last = [s for s in segs if s[1] == 128][0]
    assert last[2]
    assert sum(o for _, _, o in segs) == 1


def test_forced_length_and_cap():
    segs, flags = plan(1024, (), fused_lx=8)                  # 128 parts wanted: capped to the segment table
    check_invariants(1024, segs, flags)
    segs, flags = plan(1024, tuple(range(0, 1024, 20)), zones=1)    # 52 op planes: zones widen until they fit
    check_invariants(1024, segs, flags)
