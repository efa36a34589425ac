"""Cases for the anisotropic tensor update (row a23): shared by the golden generator, the oracle test and the GPU test."""
import numpy as np

DT = 3.1e-17
SHAPE = (6, 5, 4)


def _rot():
    """A rotated uniaxial tensor (full, symmetric), built without SciPy: R diag(no^2, no^2, ne^2) R^T."""
    a, b = 0.4, 1.1
    rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1.0, 0], [-np.sin(b), 0, np.cos(b)]])
    r = rz @ ry
    return r @ np.diag([2.2 ** 2, 2.2 ** 2, 2.29 ** 2]) @ r.T


def inputs(dtype):
    rng = np.random.default_rng(77)
    f = [rng.standard_normal(SHAPE).astype(dtype) for _ in range(3)]
    c = [(rng.standard_normal(SHAPE) * 1e6).astype(dtype) for _ in range(3)]
    return f, c


def cases():
    """name -> dict(eps=TensorComponents kwargs, mu=kwargs or None).  Values are Python floats unless arrays."""
    rng = np.random.default_rng(3)
    t = _rot()
    var = 1.0 + rng.random(SHAPE)
    full_var = np.broadcast_to(t, SHAPE + (3, 3)).copy()
    full_var[..., 0, 0] += var
    full_var[..., 1, 2] += 0.1 * var
    full_var[..., 2, 1] += 0.1 * var

    def comps(m):
        return dict(xx=m[..., 0, 0], yy=m[..., 1, 1], zz=m[..., 2, 2], xy=m[..., 0, 1], xz=m[..., 0, 2], yz=m[..., 1, 2],
                    yx=m[..., 1, 0], zx=m[..., 2, 0], zy=m[..., 2, 1])

    return {
        "diag_scalar": dict(eps=dict(xx=2.25, yy=4.0, zz=12.11), mu=dict(xx=1.0, yy=1.5, zz=2.0)),
        "diag_array": dict(eps=dict(xx=var, yy=var * 2.0, zz=3.0), mu=None),
        "full_scalar": dict(eps={k: float(v) for k, v in comps(t).items()}, mu={k: float(v) for k, v in comps(t * 0.5).items()}),
        "full_array": dict(eps=comps(full_var), mu=None),
    }
