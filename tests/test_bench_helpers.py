"""bench.py's workload description on x-slabs (CPU): the ops every rank builds for its slab — equal or load-balanced —
must tile the global source / monitor planes exactly once, with ghost copies of the source only on the rank whose
three ghost planes contain it (the two-step sweep recomputes the intermediate step there)."""
import numpy as np
import pytest

import bench as B
from prismo_b200.multigpu import balanced_slab_ranges, plane_costs, slab_range


def _spans(nx, world, balanced):
    if not balanced:
        return [slab_range(nx, r, world) for r in range(world)]
    dt, sp = B.workload_timestep()
    s, m = B.workload_ops((nx, 64, 64), dt, sp)
    return balanced_slab_ranges(plane_costs(nx, 64 * 64, s, m), world)


@pytest.mark.parametrize("nx,world", [(1024, 8), (1024, 4), (1024, 2), (96, 8), (100, 3)])
@pytest.mark.parametrize("balanced", [False, True])
def test_slab_ops_tile_the_global_ops(nx, world, balanced):
    dims = (nx, 64, 64)
    dt, sp = B.workload_timestep()
    g_src, g_mon = B.workload_ops(dims, dt, sp)
    src_plane, mon_plane = g_src[0].lo[0], g_mon[0].lo[0]
    spans = _spans(nx, world, balanced)
    assert spans[0][0] == 0 and sum(n for _, n in spans) == nx
    owners_src, owners_mon, ghosts = [], [], []
    for r, (x0, n) in enumerate(spans):
        s, m = B.workload_ops(dims, dt, sp, x0, n)
        real = [o for o in s if not o.ghost]
        gh = [o for o in s if o.ghost]
        if real:
            owners_src.append(r)
            assert all(o.lo[0] + x0 == src_plane and o.hi[0] == o.lo[0] + 1 and 0 <= o.lo[0] < n for o in real)
            assert [o.component for o in real] == [o.component for o in g_src]
        if gh:
            ghosts.append(r)
            assert all(n <= o.lo[0] < n + 3 and o.lo[0] + x0 == src_plane for o in gh)
        if m:
            owners_mon.append(r)
            assert all(o.lo[0] + x0 == mon_plane and 0 <= o.lo[0] < n and o.n_freq == 5 for o in m)
    assert len(owners_src) == 1 and len(owners_mon) == 1
    x0s, ns = spans[owners_src[0]]
    assert ghosts == ([owners_src[0] - 1] if (src_plane - x0s < 3 and owners_src[0] > 0) else [])


def test_component_shapes_follow_the_staggering_on_every_slab():
    dims = (40, 12, 10)
    dt, sp = B.workload_timestep()
    for x0, n in ((0, 13), (13, 14), (27, 13)):
        s, m = B.workload_ops(dims, dt, sp, x0, n)
        for o in s + m:
            # Ey: (nx-1, ny, nz-1) ; Hz: (nx, ny, nz-1): full rows in y, one less in z
            assert o.hi[1] == 12 and o.hi[2] == 9 and o.lo[1:] == (0, 0)


def test_parse_workload_names():
    assert B.parse_workload("c4")[1] == (1024, 1024, 1024)
    assert B.parse_workload("128x1024x1024")[1] == (128, 1024, 1024)
    assert B.parse_workload("c2")[1] == (121, 121, 121)


@pytest.mark.parametrize("name,dims", [("c3", (64, 64, 48)), ("c5", (128, 96, 48))])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_medium_recursions_tile_over_slabs(name, dims, world):
    """The dispersive-medium boxes every rank builds for its x-slab must tile the single-GPU boxes exactly once."""
    dt, _ = B.workload_timestep()
    med = B.Medium(name, dims)
    whole = med.ade_ops(dt)
    assert len(whole) == (3 if name == "c3" else 9)
    cover = {(o.component, o.kind, o.lo[1:], o.hi[1:]): np.zeros(dims[0], dtype=int) for o in whole}
    for r in range(world):
        x0, n = slab_range(dims[0], r, world)
        for o in med.ade_ops(dt, x0, n):
            assert 0 <= o.lo[0] < o.hi[0] <= n
            cover[(o.component, o.kind, o.lo[1:], o.hi[1:])][x0 + o.lo[0]:x0 + o.hi[0]] += 1
    for o in whole:
        want = np.zeros(dims[0], dtype=int)
        want[o.lo[0]:o.hi[0]] = 1
        assert np.array_equal(cover[(o.component, o.kind, o.lo[1:], o.hi[1:])], want), o
    eps = med.eps()
    assert eps.shape == dims[1:] and set(np.unique(eps)) == {1.0, 2.07, 12.11}
    ca, cb, da, db = med.coefficients(dt, 5, (3, 20, 4, 30))
    assert cb.shape == (5, 17, 26) and np.array_equal(cb[0], dt / (8.854187817e-12 * eps[3:20, 4:30]))


def test_port_planes_are_owned_by_exactly_one_slab():
    dims = (128, 96, 48)
    for world in (1, 2, 8):
        seen = []
        for r in range(world):
            x0, n = slab_range(dims[0], r, world)
            seen += [(port, which, p) for port, which, p, ops in B.port_monitor_ops(dims, x0, n) if ops is not None]
        assert sorted(seen) == [(0, 0, 16), (0, 1, 24), (1, 0, 104), (1, 1, 112)]


@pytest.mark.parametrize("name,aniso", [("c3", False), ("c5", False), ("c5", True)])
def test_medium_shape_list_paints_the_same_coefficients_as_the_host_arrays(name, aniso):
    """bench.py paints the named media on the DEVICE from Medium.shapes(); the self-check's oracle uses Medium.coefficients():
    both must describe the same medium (checked here with the rasterisation oracle, no GPU)."""
    import bench as B
    from oracle import raster

    dims, spacing, dt = (20, 64, 48), (2e-8, 2e-8, 2e-8), 3e-17
    med = B.Medium(name, dims, aniso)
    ax = [np.arange(n) * d for n, d in zip(dims, spacing)]
    shapes = [dict(kind="box", center=tuple(s.center), size=tuple(s.size), eps_r=s.material.epsilon_r) for s in med.shapes(spacing)]
    got = raster.coefficient_arrays(shapes, ax[0], ax[1], ax[2], dt)
    want = med.coefficients(dt, dims[0])
    for k in (0, 2, 3):
        assert np.array_equal(got[k], want[k]), k
    if aniso:
        for c in range(3):
            assert np.array_equal(got[1][c], want[1][c]), c
        assert not np.array_equal(want[1][0], want[1][2])
    else:
        assert np.array_equal(got[1], want[1])
    # and on an x-slab: planes [x0, x0 + n)
    got = raster.coefficient_arrays(shapes, ax[0][7:13], ax[1], ax[2], dt)
    assert np.array_equal(got[0], want[0][7:13]) and np.array_equal(got[3], want[3][7:13])
