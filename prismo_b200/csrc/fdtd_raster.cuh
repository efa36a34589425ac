// fdtd_raster.cuh — geometry rasterisation on the device (first half).
//
// The reference paints structures on the host: Shape.rasterize (geometry/shapes.py:71-99) builds the full meshgrid of
// cell coordinates (three fp64 values per cell: 25 GB at 1024^3), evaluates contains() on it, the user writes the
// material into eps_rel / mu_rel / sigma arrays through the mask, and MaxwellUpdater turns those into the four update
// coefficient arrays (core/solver.py:113-133).  Here one kernel goes from the shape list to Ca, Cb, Da, Db in the
// engine's layout: each cell evaluates the shapes in list order at its own coordinate (x[i], y[j], z[k]) — later shapes
// paint over earlier ones — and forms its coefficients; nothing of full-grid size ever exists on the host, and the
// only HBM traffic is the 4 (6 with per-component Cb) coefficient stores: 16 B per cell in fp32.
//
// Bit-exactness with the reference: the contains() tests use the same fp64 operations in the same order —
//   Box       all(|p - c| <= size / 2)                                   shapes.py:125-132   (size / 2 formed by the caller)
//   Sphere    sqrt((dx^2 + dy^2) + dz^2) <= radius                       shapes.py:155-159   (np.linalg.norm(axis=1): add.reduce
//             over the three squares of a row accumulates left to right; pinned by tests/test_raster.py)
//   Cylinder  sqrt(da**2 + db**2) <= radius  &  |p_ax - c_ax| <= h / 2   shapes.py:194-214
//   Polygon   ray casting in xy, z_min <= z <= z_max                     shapes.py:249-283
//   GeometryGroup  union / intersection / difference of member masks     shapes.py:338-378
// — and the coefficient formulas are evaluated operation by operation as NumPy does (true divisions, no contraction).
#pragma once
#include "fdtd_kernels.cuh"

namespace fdtd {

struct RasterShape {
    int kind, axis;                    // 0 box, 1 sphere, 2 cylinder, 3 polygon; cylinder axis 0/1/2
    int combine, paint;                // combine: 0 new mask, 1 |=, 2 &=, 3 &= ~ ; paint: 1 = cells of the running mask take the material
    double c[3];                       // centre
    double a[3];                       // box: half sizes; sphere: a[0] = radius; cylinder: a[0] = radius, a[1] = height / 2; polygon: a[0] = z_min, a[1] = z_max
    int v0, nv;                        // polygon: first vertex / vertex count in the xy vertex list
    double eps[3], mu_r, sigma_e, sigma_m;   // eps[0..2] = eps_xx, eps_yy, eps_zz (equal for isotropic materials)
};

__device__ __forceinline__ bool raster_contains(const RasterShape& s, const double* __restrict__ verts, double px, double py, double pz)
{
    if (s.kind == 3) {
        if (!(pz >= s.a[0] && pz <= s.a[1])) return false;
        int count = 0;
        for (int q = 0; q < s.nv; ++q) {
            const int q2 = (q + 1 == s.nv) ? 0 : q + 1;
            const double v1x = verts[2 * (s.v0 + q)], v1y = verts[2 * (s.v0 + q) + 1];
            const double v2x = verts[2 * (s.v0 + q2)], v2y = verts[2 * (s.v0 + q2) + 1];
            if ((v1y > py) != (v2y > py)) {
                const double xc = __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(v2x, v1x), __dsub_rn(py, v1y)), __dsub_rn(v2y, v1y)), v1x);
                if (px < xc) ++count;
            }
        }
        return (count & 1) != 0;
    }
    const double dx = __dsub_rn(px, s.c[0]), dy = __dsub_rn(py, s.c[1]), dz = __dsub_rn(pz, s.c[2]);
    if (s.kind == 0) return fabs(dx) <= s.a[0] && fabs(dy) <= s.a[1] && fabs(dz) <= s.a[2];
    if (s.kind == 1)
        return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz))) <= s.a[0];
    const double d[3] = {dx, dy, dz};
    const int p0 = s.axis == 0 ? 1 : 0, p1 = s.axis == 2 ? 1 : 2;
    const double radial = __dsqrt_rn(__dadd_rn(__dmul_rn(d[p0], d[p0]), __dmul_rn(d[p1], d[p1])));
    return radial <= s.a[0] && fabs(d[s.axis]) <= s.a[1];
}

struct RasterBg { double eps[3], mu_r, sigma_e, sigma_m; };

// core/solver.py:119-130 for one material, evaluated left to right in fp64 and rounded once to T:
// out = Ca, Cb (eps_xx), Da, Db, Cb_y, Cb_z
template <typename T, bool ANISO>
__device__ __forceinline__ void raster_coefs(const double* eps3, double mr, double se, double sm, double dt, double eps0, double mu0, T* out)
{
    const double sdt = __dmul_rn(se, dt);
    const double eps = __dmul_rn(eps0, eps3[0]);
    const double s2e = __ddiv_rn(sdt, __dmul_rn(2.0, eps));
    out[0] = (T)__ddiv_rn(__dsub_rn(1.0, s2e), __dadd_rn(1.0, s2e));
    out[1] = (T)__ddiv_rn(__ddiv_rn(dt, eps), __dadd_rn(1.0, s2e));
    const double mu = __dmul_rn(mu0, mr);
    const double s2m = __ddiv_rn(__dmul_rn(sm, dt), __dmul_rn(2.0, mu));
    out[2] = (T)__ddiv_rn(__dsub_rn(1.0, s2m), __dadd_rn(1.0, s2m));
    out[3] = (T)__ddiv_rn(__ddiv_rn(dt, mu), __dadd_rn(1.0, s2m));
    if (ANISO) {
        // per-component Cb (Ca stays the eps_xx one: anisotropic materials are lossless here, the host checks it)
        const double epy = __dmul_rn(eps0, eps3[1]), epz = __dmul_rn(eps0, eps3[2]);
        const double s2y = __ddiv_rn(sdt, __dmul_rn(2.0, epy)), s2z = __ddiv_rn(sdt, __dmul_rn(2.0, epz));
        out[4] = (T)__ddiv_rn(__ddiv_rn(dt, epy), __dadd_rn(1.0, s2y));
        out[5] = (T)__ddiv_rn(__ddiv_rn(dt, epz), __dadd_rn(1.0, s2z));
    }
}

// Coefficient arrays of `planes` x c1 x c2 cells from the shape list.  cby / cbz: per-component Cb (diagonal anisotropy,
// only written when ANISO); cb always holds the eps_xx one.  The coefficients depend on the MATERIAL only, so a list of up
// to kRasterSmemShapes entries is staged in shared memory together with one coefficient row per entry (+ the background)
// computed once per CTA: a cell then costs its containment tests, one table row and 4 (6) stores — the kernel is bound by
// its 16 (24) B of fp32 stores per cell.  Longer lists are walked in global memory and evaluate the formulas per cell.
// One CTA works on whole rows of the contiguous axis (k in 3-D, j in 2-D): coalesced stores, one integer division per row.
constexpr int kRasterSmemShapes = 64;

template <typename T, bool ANISO>
__global__ void __launch_bounds__(256)
k_rasterize(T* __restrict__ ca, T* __restrict__ cb, T* __restrict__ da, T* __restrict__ db, T* __restrict__ cby, T* __restrict__ cbz,
            const double* __restrict__ xs, const double* __restrict__ ys, const double* __restrict__ zs,
            const RasterShape* __restrict__ shapes, int n_shapes, const double* __restrict__ verts, RasterBg bg, double dt,
            double eps0, double mu0, long long total, int c1, int c2, Strides3 st,
            unsigned char* __restrict__ mat, T* __restrict__ mat_tab)
{
    __shared__ RasterShape s_sh[kRasterSmemShapes];
    __shared__ T s_tab[kRasterSmemShapes + 1][6];
    const bool staged = n_shapes <= kRasterSmemShapes;
    if (staged) {
        for (int q = threadIdx.x; q < n_shapes; q += blockDim.x) s_sh[q] = shapes[q];
        for (int q = threadIdx.x; q <= n_shapes; q += blockDim.x) {
            if (q < n_shapes) raster_coefs<T, ANISO>(shapes[q].eps, shapes[q].mu_r, shapes[q].sigma_e, shapes[q].sigma_m, dt, eps0, mu0, s_tab[q]);
            else raster_coefs<T, ANISO>(bg.eps, bg.mu_r, bg.sigma_e, bg.sigma_m, dt, eps0, mu0, s_tab[q]);
        }
    }
    __syncthreads();
    // material-index coding (mat != null, staged lists only): index 0 = background, q + 1 = list entry q; the table goes
    // to global memory once, in that order, for the sweeps to stage (fdtd_het.cuh IDX)
    if (mat_tab && blockIdx.x == 0)
        for (int q = threadIdx.x; q <= n_shapes; q += blockDim.x) {
            const T* r = s_tab[q == 0 ? n_shapes : q - 1];
            T* o = mat_tab + 6 * q;
            o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = r[3];
            o[4] = ANISO ? r[4] : r[1]; o[5] = ANISO ? r[5] : r[1];
        }
    const RasterShape* __restrict__ sh = staged ? s_sh : shapes;
    const bool is3 = zs != nullptr;
    const int inner = is3 ? c2 : c1;                      // contiguous axis
    const long long rows = total / inner;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const long long i = is3 ? r / c1 : r;
        const int jr = is3 ? (int)(r - i * c1) : 0;
        const double px = xs[i];
        for (int q = threadIdx.x; q < inner; q += blockDim.x) {
            const int j = is3 ? jr : q, k = is3 ? q : 0;
            const double py = ys[j], pz = is3 ? zs[k] : 0.0;
            int win = n_shapes;                           // the background
            bool m = false;
            for (int n = 0; n < n_shapes; ++n) {
                const RasterShape& s = sh[n];
                const bool in = raster_contains(s, verts, px, py, pz);
                m = s.combine == 0 ? in : s.combine == 1 ? (m || in) : s.combine == 2 ? (m && in) : (m && !in);
                if (s.paint && m) win = n;
            }
            T row[6];
            if (staged) {
#pragma unroll
                for (int c = 0; c < (ANISO ? 6 : 4); ++c) row[c] = s_tab[win][c];
            } else if (win < n_shapes) {
                raster_coefs<T, ANISO>(sh[win].eps, sh[win].mu_r, sh[win].sigma_e, sh[win].sigma_m, dt, eps0, mu0, row);
            } else {
                raster_coefs<T, ANISO>(bg.eps, bg.mu_r, bg.sigma_e, bg.sigma_m, dt, eps0, mu0, row);
            }
            const long long o = i * st.s[0] + j * st.s[1] + k * st.s[2];
            ca[o] = row[0]; cb[o] = row[1]; da[o] = row[2]; db[o] = row[3];
            if (ANISO) { cby[o] = row[4]; cbz[o] = row[5]; }
            if (mat) mat[o] = (unsigned char)(win == n_shapes ? 0 : win + 1);
        }
    }
}

}  // namespace fdtd
