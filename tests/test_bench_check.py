"""bench_check.py (the bench's self-check) on the CPU: a full-grid oracle run stands in for the device, split over several
"ranks" to exercise the assembly of boxes and checksums.  The cropped oracle must reproduce the full-grid run exactly
(locality of the forward-differenced scheme), the DFT crop must match, the checksum digest must not depend on the
decomposition, and a corrupted cell must be caught."""
import numpy as np
import pytest

import bench as B
import bench_check as BC
from oracle import kernels

DIMS = (96, 72, 80)


class FullGridStandIn:
    def __init__(self, corrupt=None):
        self.dt, self.spacing = B.workload_timestep()
        self.src, self.mon = DIMS[0] // 4, (3 * DIMS[0]) // 4
        self.corrupt = corrupt
        self.F, self.dft = None, None

    def reseed(self, n, amp, ph):
        self.amp, self.ph = amp, ph
        planes = BC.seed_planes(DIMS)
        self.F = {c: BC.seed_box(planes, c, (0, 0, 0), BC.comp_shape(c, DIMS)).astype(np.float64) for c in BC.COMPONENTS}
        self.dft = {c: np.zeros((ph.shape[1],) + BC.comp_shape(c, DIMS)[1:], dtype=np.complex128) for c in ("Ey", "Hz")}
        return planes

    def run(self, n):
        coeffs = kernels.vacuum_coefficients(DIMS, self.dt)
        for s in range(n):
            kernels.step(self.F, coeffs, self.spacing, False)
            self.F["Ey"][self.src] += self.amp[s, 0]
            self.F["Hz"][self.src] += self.amp[s, 1]
            for c in self.dft:
                for f in range(self.ph.shape[1]):
                    self.dft[c][f] += self.F[c][self.mon] * self.ph[s, f] * self.dt
        if self.corrupt:
            c, idx = self.corrupt
            self.F[c][idx] *= 1.0 + 1e-3

    def box(self, c, lo, hi):
        return self.F[c][tuple(slice(l, h) for l, h in zip(lo, hi))].copy()

    def dft_crop(self, c, lo2, hi2):
        return self.dft[c][:, lo2[0]:hi2[0], lo2[1]:hi2[1]]

    def checksums(self, cuts=(0, DIMS[0])):
        """NumPy model of fdtd_field_checksum on fp64 arrays, assembled from slabs cut at `cuts`."""
        out = {}
        for c in BC.COMPONENTS:
            a = self.F[c]
            parts = []
            for x0, x1 in zip(cuts[:-1], cuts[1:]):
                sl = a[x0:min(x1, a.shape[0])]
                bits = sl.view(np.uint64).reshape(sl.shape[0], -1)
                w = (np.arange(bits.shape[1], dtype=np.uint64) + np.uint64(1))[None, :]
                with np.errstate(over="ignore"):
                    parts.append(np.stack([bits.sum(axis=1, dtype=np.uint64), (bits * w).sum(axis=1, dtype=np.uint64)], axis=1))
            out[c] = np.concatenate(parts, axis=0)
        return out


def _check(dev, cuts=(0, DIMS[0])):
    return BC.run_check(DIMS, dev.dt, dev.spacing, "float64", B.tables, dev.src, dev.mon, dev.reseed, dev.run, dev.box,
                        dev.dft_crop, lambda: dev.checksums(cuts))


def test_cropped_oracle_equals_full_grid_run():
    out = _check(FullGridStandIn())
    assert len(out["crops"]) == 2 and out["crops"][0]["has_source_plane"] and out["crops"][1]["at_monitor_plane"]
    assert out["crop_bit_exact"] and out["crop_rel_l2"] == 0.0
    assert out["dft_rel_l2"] is not None and out["dft_rel_l2"] < 1e-15
    assert out["ok"]


def test_checksum_digest_is_independent_of_the_slab_cut():
    a = _check(FullGridStandIn())
    b = _check(FullGridStandIn(), cuts=(0, 17, 40, 41, DIMS[0]))
    assert a["fields_sha"] == b["fields_sha"] and len(a["fields_sha"]) == 32


@pytest.mark.parametrize("where", [("Ey", (26, 38, 41)), ("Hx", (73, 25, 27))])
def test_a_wrong_cell_is_caught(where):
    good = _check(FullGridStandIn())
    bad = _check(FullGridStandIn(corrupt=where))
    assert not bad["ok"] and bad["crop_rel_l2"] > 1e-10
    assert bad["fields_sha"] != good["fields_sha"]


def test_seed_is_a_function_of_the_global_index():
    planes = BC.seed_planes(DIMS)
    whole = BC.seed_box(planes, "Ez", (0, 0, 0), BC.comp_shape("Ez", DIMS))
    part = BC.seed_box(planes, "Ez", (40, 3, 5), (61, 30, 44))
    assert np.array_equal(whole[40:61, 3:30, 5:44], part) and whole.dtype == np.float32
    assert np.abs(np.diff(whole[:, 7, 9])).max() > 0            # varies along x
