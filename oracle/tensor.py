"""Oracle restatement of the reference's anisotropic (tensor) field update.  Test infrastructure only.

Follows /root/reference/src/prismo/materials/tensor.py:482-536 (update_e_from_curl_h) and :538-588
(update_h_from_curl_e): co-located arrays supplied by the caller, NumPy evaluation order, NumPy dtype promotion
(Python-float tensor entries are weak scalars; the inverse tensor's entries are float64).  Row a23 of the scope table;
pinned against the real reference by tests/golden/aniso.npz (tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np

EPS0 = 8.854187817e-12
MU0 = 4 * np.pi * 1e-7


def update_e(E, curl_H, dt, diagonal=None, inverse=None):
    """diagonal = (xx, yy, zz) relative permittivities, or inverse = (..., 3, 3) inverse tensor."""
    s = dt / EPS0
    if inverse is None:
        return tuple(e + s * c / d for e, c, d in zip(E, curl_H, diagonal))
    cx, cy, cz = curl_H
    return tuple(e + s * (inverse[..., i, 0] * cx + inverse[..., i, 1] * cy + inverse[..., i, 2] * cz)
                 for i, e in enumerate(E))


def update_h(H, curl_E, dt, diagonal=None, inverse=None):
    s = dt / MU0
    if inverse is None:
        return tuple(h - s * c / d for h, c, d in zip(H, curl_E, diagonal))
    cx, cy, cz = curl_E
    return tuple(h + -s * (inverse[..., i, 0] * cx + inverse[..., i, 1] * cy + inverse[..., i, 2] * cz)
                 for i, h in enumerate(H))
