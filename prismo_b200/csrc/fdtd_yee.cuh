// fdtd_yee.cuh — opt-in "physics" mode: the stable Yee leap-frog on the reference's staggering, with CPML.
//
// NOT parity code: the reference's own update is forward/forward-differenced and unstable and its
// CPML is never applied (F6, boundaries/pml.py:259-328).  On the reference's staggering
//   Ex(i, j+1/2, k+1/2) Ey(i+1/2, j, k+1/2) Ez(i+1/2, j+1/2, k)   Hx(i+1/2, j, k) Hy(i, j+1/2, k) Hz(i, j, k+1/2)
// the consistent scheme keeps the reference's FORWARD differences in the E update and uses BACKWARD
// differences in the H update.  CPML (Roden & Gedney 2000): every derivative d/da inside the boundary slabs
// becomes  (1/kappa_a) d/da + psi,  psi <- b_a psi + a_a d/da ; the 12 psi arrays exist ONLY in the slabs and
// only slab threads touch them.  Graded profiles follow boundaries/pml.py:117-151 (PMLParams).
// Validated against oracle/yee.py (our own restatement) and by physics checks (tests/test_physics_mode.py):
// parity is UNPINNED for this mode — there are no reference numbers to match.
#pragma once
#include "fdtd_kernels.cuh"

namespace fdtd {

// per-axis CPML coefficient vectors in device memory: b, a, 1/kappa, at E-derivative (half) positions [0..2]
// and H-derivative (integer) positions [3..5]; identity (b=0,a=0,1/kappa=1) outside the layer
struct CpmlAxis { const double* c[6]; const float* f[6]; };    // fp64 tables (fp64 engines) and their fp32 copies (fp32 engines)
template <typename T> __device__ __forceinline__ const T* cpml_tab(const CpmlAxis& ax, int v);
template <> __device__ __forceinline__ const double* cpml_tab<double>(const CpmlAxis& ax, int v) { return ax.c[v]; }
template <> __device__ __forceinline__ const float* cpml_tab<float>(const CpmlAxis& ax, int v) { return ax.f[v]; }
// psi <- b psi + a d ;  d_eff = d / kappa + psi.  fp64: every operation rounded separately (what oracle/yee.py's NumPy
// expressions do); fp32: two fused multiply-adds in fp32 (the mode's fp32 bound is 1e-4 against the fp64 run).
template <typename T> struct CpmlMath;
template <> struct CpmlMath<double> {
    static __device__ __forceinline__ double psi(double b, double p, double a, double d) { return __dadd_rn(__dmul_rn(b, p), __dmul_rn(a, d)); }
    static __device__ __forceinline__ double eff(double ki, double d, double p) { return __dadd_rn(__dmul_rn(ki, d), p); }
};
template <> struct CpmlMath<float> {
    static __device__ __forceinline__ float psi(float b, float p, float a, float d) { return fmaf(b, p, a * d); }
    static __device__ __forceinline__ float eff(float ki, float d, float p) { return fmaf(ki, d, p); }
};
struct Cpml {
    int t;                    // layer thickness (0 = no CPML)
    int ns;                   // slab entries per axis = 2t + 1
    CpmlAxis ax[3];
    // psi arrays: [0..5] E update (Ex_y, Ex_z, Ey_z, Ey_x, Ez_x, Ez_y), [6..11] H update (Hx_y, Hx_z, Hy_z, Hy_x, Hz_x, Hz_y)
    void* psi[12];
};

__device__ __forceinline__ int slab_index(int n, int N, int t)      // -1 outside the two slabs
{
    if (n < t) return n;
    if (n >= N - t - 1) return n - (N - 2 * t - 1);
    return -1;
}

// storage offsets of the three slab families (same row pitch pz as the fields)
struct SlabGeom { long long x_sx, y_sx, z_sx; int z_pitch; };

template <typename T>
__device__ __forceinline__ T cpml_apply(T d, T* psi, long long o, const CpmlAxis& ax, int v, int n)
{
    const T p = CpmlMath<T>::psi(cpml_tab<T>(ax, v)[n], psi[o], cpml_tab<T>(ax, v + 1)[n], d);
    psi[o] = p;
    return CpmlMath<T>::eff(cpml_tab<T>(ax, v + 2)[n], d, p);
}

template <typename T, bool HET>
__global__ void __launch_bounds__(256)
k_h3d_yee(Fields<T> f, Coefs<T> c, Geom g, Cpml pm, SlabGeom sg)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int i = blockIdx.z;
    if (k >= g.nz || j >= g.ny) return;
    const long long o = (long long)i * g.sx + (long long)j * g.sy + k;
    const int sxi = pm.t ? slab_index(i, g.nx, pm.t) : -1;
    const int syj = pm.t ? slab_index(j, g.ny, pm.t) : -1;
    const int szk = pm.t ? slab_index(k, g.nz, pm.t) : -1;
    const long long ox = sxi * sg.x_sx + (long long)j * g.sy + k;
    const long long oy = (long long)i * sg.y_sx + (long long)syj * g.sy + k;
    const long long oz = ((long long)i * g.ny + j) * sg.z_pitch + szk;
    T da = c.uda, db = c.udb;
    // Hx(i+1/2, j, k): i < nx-1, 1 <= j <= ny-2, 1 <= k <= nz-2
    if (i < g.nx - 1 && j >= 1 && j <= g.ny - 2 && k >= 1 && k <= g.nz - 2) {
        if (HET) { da = mean2<T>(c.da[o], c.da[o + g.sx]); db = mean2<T>(c.db[o], c.db[o + g.sx]); }
        T dy = Ar<T>::diff(f.ez[o], f.ez[o - g.sy], g.dy, g.rdy), dz = Ar<T>::diff(f.ey[o], f.ey[o - 1], g.dz, g.rdz);
        if (syj >= 0) dy = cpml_apply<T>(dy, (T*)pm.psi[6], oy, pm.ax[1], 3, j);
        if (szk >= 0) dz = cpml_apply<T>(dz, (T*)pm.psi[7], oz, pm.ax[2], 3, k);
        f.hx[o] = upd_h<T>(da, f.hx[o], db, dy, dz);
    }
    // Hy(i, j+1/2, k): 1 <= i <= nx-2, j < ny-1, 1 <= k <= nz-2
    if (i >= 1 && i <= g.nx - 2 && j < g.ny - 1 && k >= 1 && k <= g.nz - 2) {
        if (HET) { da = mean2<T>(c.da[o], c.da[o + g.sy]); db = mean2<T>(c.db[o], c.db[o + g.sy]); }
        T dz = Ar<T>::diff(f.ex[o], f.ex[o - 1], g.dz, g.rdz), dx = Ar<T>::diff(f.ez[o], f.ez[o - g.sx], g.dx, g.rdx);
        if (szk >= 0) dz = cpml_apply<T>(dz, (T*)pm.psi[8], oz, pm.ax[2], 3, k);
        if (sxi >= 0) dx = cpml_apply<T>(dx, (T*)pm.psi[9], ox, pm.ax[0], 3, i);
        f.hy[o] = upd_h<T>(da, f.hy[o], db, dz, dx);
    }
    // Hz(i, j, k+1/2): 1 <= i <= nx-2, 1 <= j <= ny-2, k < nz-1
    if (i >= 1 && i <= g.nx - 2 && j >= 1 && j <= g.ny - 2 && k < g.nz - 1) {
        if (HET) { da = mean2<T>(c.da[o], c.da[o + 1]); db = mean2<T>(c.db[o], c.db[o + 1]); }
        T dx = Ar<T>::diff(f.ey[o], f.ey[o - g.sx], g.dx, g.rdx), dy = Ar<T>::diff(f.ex[o], f.ex[o - g.sy], g.dy, g.rdy);
        if (sxi >= 0) dx = cpml_apply<T>(dx, (T*)pm.psi[10], ox, pm.ax[0], 3, i);
        if (syj >= 0) dy = cpml_apply<T>(dy, (T*)pm.psi[11], oy, pm.ax[1], 3, j);
        f.hz[o] = upd_h<T>(da, f.hz[o], db, dx, dy);
    }
}

template <typename T, bool HET>
__global__ void __launch_bounds__(256)
k_e3d_yee(Fields<T> f, Coefs<T> c, Geom g, Cpml pm, SlabGeom sg)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int i = blockIdx.z;
    if (k >= g.nz || j >= g.ny) return;
    const long long o = (long long)i * g.sx + (long long)j * g.sy + k;
    const int sxi = pm.t ? slab_index(i, g.nx, pm.t) : -1;
    const int syj = pm.t ? slab_index(j, g.ny, pm.t) : -1;
    const int szk = pm.t ? slab_index(k, g.nz, pm.t) : -1;
    const long long ox = sxi * sg.x_sx + (long long)j * g.sy + k;
    const long long oy = (long long)i * sg.y_sx + (long long)syj * g.sy + k;
    const long long oz = ((long long)i * g.ny + j) * sg.z_pitch + szk;
    T ca = c.uca, cb = c.ucb;
    // Ex(i, j+1/2, k+1/2): j < ny-1, k < nz-1
    if (j < g.ny - 1 && k < g.nz - 1) {
        if (HET) {
            ca = mean4<T>(c.ca[o], c.ca[o + g.sy], c.ca[o + 1], c.ca[o + g.sy + 1]);
            cb = mean4<T>(c.cb[o], c.cb[o + g.sy], c.cb[o + 1], c.cb[o + g.sy + 1]);
        }
        T dy = Ar<T>::diff(f.hz[o + g.sy], f.hz[o], g.dy, g.rdy), dz = Ar<T>::diff(f.hy[o + 1], f.hy[o], g.dz, g.rdz);
        if (syj >= 0) dy = cpml_apply<T>(dy, (T*)pm.psi[0], oy, pm.ax[1], 0, j);
        if (szk >= 0) dz = cpml_apply<T>(dz, (T*)pm.psi[1], oz, pm.ax[2], 0, k);
        f.ex[o] = upd_e<T>(ca, f.ex[o], cb, dy, dz);
    }
    // Ey(i+1/2, j, k+1/2): i < nx-1, k < nz-1
    if (i < g.nx - 1 && k < g.nz - 1) {
        if (HET) {
            ca = mean4<T>(c.ca[o], c.ca[o + g.sx], c.ca[o + 1], c.ca[o + g.sx + 1]);
            cb = mean4<T>(c.cb[o], c.cb[o + g.sx], c.cb[o + 1], c.cb[o + g.sx + 1]);
        }
        T dz = Ar<T>::diff(f.hx[o + 1], f.hx[o], g.dz, g.rdz), dx = Ar<T>::diff(f.hz[o + g.sx], f.hz[o], g.dx, g.rdx);
        if (szk >= 0) dz = cpml_apply<T>(dz, (T*)pm.psi[2], oz, pm.ax[2], 0, k);
        if (sxi >= 0) dx = cpml_apply<T>(dx, (T*)pm.psi[3], ox, pm.ax[0], 0, i);
        f.ey[o] = upd_e<T>(ca, f.ey[o], cb, dz, dx);
    }
    // Ez(i+1/2, j+1/2, k): i < nx-1, j < ny-1
    if (i < g.nx - 1 && j < g.ny - 1) {
        if (HET) {
            ca = mean4<T>(c.ca[o], c.ca[o + g.sx], c.ca[o + g.sy], c.ca[o + g.sx + g.sy]);
            cb = mean4<T>(c.cb[o], c.cb[o + g.sx], c.cb[o + g.sy], c.cb[o + g.sx + g.sy]);
        }
        T dx = Ar<T>::diff(f.hy[o + g.sx], f.hy[o], g.dx, g.rdx), dy = Ar<T>::diff(f.hx[o + g.sy], f.hx[o], g.dy, g.rdy);
        if (sxi >= 0) dx = cpml_apply<T>(dx, (T*)pm.psi[4], ox, pm.ax[0], 0, i);
        if (syj >= 0) dy = cpml_apply<T>(dy, (T*)pm.psi[5], oy, pm.ax[1], 0, j);
        f.ez[o] = upd_e<T>(ca, f.ez[o], cb, dx, dy);
    }
}

}  // namespace fdtd
