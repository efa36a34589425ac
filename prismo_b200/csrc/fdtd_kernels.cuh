// fdtd_kernels.cuh — sm_100a device code of the B200 FDTD engine (two-pass kernels, 2-D kernels,
// source / monitor / layout kernels).  The fused single-sweep kernel lives in fdtd_fused.cuh.
//
// Reference semantics (bug-compatible): /root/reference/src/prismo/core/solver.py
//   3-D H pass :167-253, E pass :255-309, 2-D H pass :311-397, E pass :399-456, averaging :458-533.
//
// Layout: six SoA arrays with IDENTICAL strides.  3-D: element (i,j,k) at i*sx + j*sy + k with
// sy = pz = round_up(nz,32), sx = ny*pz; 2-D: (i,j) at i*sx + j with sx = round_up(ny,32).
// Each array owns nx+2 planes: [0,nx) data, plane nx = right-neighbour ghost (multi-GPU) or zeros,
// plane nx+1 = guard, so "+1" neighbour loads never leave the allocation.  Cells outside a
// component's staggered shape are padding: always zero, never stored with anything else.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fdtd {

template <typename T> struct Fields { T *ex, *ey, *ez, *hx, *hy, *hz; };
template <typename T> struct CFields { const T *ex, *ey, *ez, *hx, *hy, *hz; };

constexpr int kMatTabRows = 72;    // material-index coding: rows of the per-CTA material table (background + <= 64 list entries)

template <typename T> struct Coefs {
    const T *ca, *cb, *da, *db;   // cell-centred arrays (field layout) when het != 0
    const T *cby, *cbz;           // per-component Cb of Ey / Ez (cb is then Ex's); null = isotropic
    const unsigned char* mat;     // material-index coding (fdtd_rasterize): one byte per cell, field layout; null = off
    const T* mat_tab;             // [n_mat][6] = Ca, Cb, Da, Db, Cb_y, Cb_z of material m (0 = background)
    int n_mat;
    T uca, ucb, uda, udb;         // uniform values otherwise
};

// 1/d for the two arithmetic policies: fp32 multiplies by f; fp64 divides exactly through y = RN(1/d) (IEEE division on
// the host; fdtd_create accepts fp64 spacings in [2^-60, 2^60] only, so that a = q*d stays far from the exponent limits
// whenever the quotient passes the kernels' range test).
struct Rcp { float f; double y; };

struct Geom {
    int nx, ny, nz;        // local logical dims
    int nxg, x0;           // global nx, global index of local plane 0
    int pz;                // padded length of the contiguous axis
    long long sx, sy;      // strides (elements); 2-D: sy = 1
    double dx, dy, dz;     // spacings
    Rcp rdx, rdy, rdz;     // reciprocals: fp32, and the correctly rounded fp64 one for the exact division sequence
};

// ---- arithmetic policies ---------------------------------------------------------------------
// fp64 reproduces NumPy bit for bit: every operation rounded separately (no FMA contraction),
// "/ d" is a true division.  fp32 is free to contract and multiplies by 1/d.
template <typename T> struct Ar;
template <> struct Ar<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    // a / d, correctly rounded, for a launch-constant divisor d with y = RN(1/d) formed once on the host.
    //   q0 = RN(a*y)                                   within 1.5 ulp of a/d
    //   r0 = a - q0*d (exact, FMA);  q1 = RN(q0 + r0*y)  faithful (error 0.5 ulp + 2^-52 ulp)
    //   r1 = a - q1*d (exact, FMA);  q2 = RN(q1 + r1*y)  = RN(a/d)   (Markstein's theorem: y = RN(1/d), q1 faithful)
    // Five FMA-pipe operations instead of __ddiv_rn's ~20-instruction sequence with two branches.  The theorem needs the
    // absence of over/underflow: a quotient outside [2^-900, 2^901) (inf, nan, denormals) sets `bad` and the CALLER
    // recomputes its whole stage with true divisions (one rarely taken branch per stage instead of two per division:
    // the hot loop stays in large basic blocks).  a == 0 returns q0 = (+-0)*y, the signed zero a true division gives.
    static __device__ __forceinline__ double div_fast(double a, double d, double y, unsigned& bad) {
        const double q0 = __dmul_rn(a, y);
        const double r0 = __fma_rn(-q0, d, a);
        const double q1 = __fma_rn(r0, y, q0);
        const double r1 = __fma_rn(-q1, d, a);
        const double q2 = __fma_rn(r1, y, q1);
        const unsigned ex = ((unsigned)__double2hiint(q0) >> 20) & 0x7ffu;      // biased exponent of q0
        const bool zero = a == 0.0;
        bad |= (unsigned)((ex - 123u > 1800u) && !zero);
        return zero ? q0 : q2;
    }
    // the same with the fallback inline (kernels off the hot path: two-pass, 2-D, heterogeneous, physics mode)
    static __device__ __forceinline__ double div_rn(double a, double d, double y) {
        unsigned bad = 0;
        const double q = div_fast(a, d, y, bad);
        return bad ? __ddiv_rn(a, d) : q;
    }
    static __device__ __forceinline__ double diff(double a1, double a0, double d, Rcp r) {
        return div_rn(__dsub_rn(a1, a0), d, r.y);
    }
    static __device__ __forceinline__ double diff_fast(double a1, double a0, double d, Rcp r, unsigned& bad) {
        return div_fast(__dsub_rn(a1, a0), d, r.y, bad);
    }
    static __device__ __forceinline__ double diff_exact(double a1, double a0, double d, Rcp) {
        return __ddiv_rn(__dsub_rn(a1, a0), d);
    }
};
template <> struct Ar<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
    static __device__ __forceinline__ float diff(float a1, float a0, double, Rcp rd) {
        return (a1 - a0) * rd.f;
    }
    static __device__ __forceinline__ float diff_fast(float a1, float a0, double, Rcp rd, unsigned&) { return (a1 - a0) * rd.f; }
    static __device__ __forceinline__ float diff_exact(float a1, float a0, double, Rcp rd) { return (a1 - a0) * rd.f; }
};

// H: da*h - db*(c1 - c2)     E: ca*e + cb*(c1 - c2)      (solver.py:201-205, :275)
template <typename T> __device__ __forceinline__ T upd_h(T da, T h, T db, T c1, T c2) {
    return Ar<T>::sub(Ar<T>::mul(da, h), Ar<T>::mul(db, Ar<T>::sub(c1, c2)));
}
template <typename T> __device__ __forceinline__ T upd_e(T ca, T e, T cb, T c1, T c2) {
    return Ar<T>::add(Ar<T>::mul(ca, e), Ar<T>::mul(cb, Ar<T>::sub(c1, c2)));
}
// 0.5*(a+b) and 0.25*(((a+b)+c)+d) in the reference's summation order (solver.py:458-501)
template <typename T> __device__ __forceinline__ T mean2(T a, T b) {
    return Ar<T>::mul((T)0.5, Ar<T>::add(a, b));
}
template <typename T> __device__ __forceinline__ T mean4(T a, T b, T c, T d) {
    return Ar<T>::mul((T)0.25, Ar<T>::add(Ar<T>::add(Ar<T>::add(a, b), c), d));
}

// ---- vector helpers: V consecutive elements along the contiguous axis, 16 bytes ---------------
template <typename T> struct VecOf;
template <> struct VecOf<float> { typedef float4 type; static const int V = 4; };
template <> struct VecOf<double> { typedef double2 type; static const int V = 2; };

template <typename T, int V> struct Pack { T v[V]; };

template <typename T> __device__ __forceinline__ Pack<T, VecOf<T>::V> ldv(const T* p) {
    typedef typename VecOf<T>::type VT;
    union { VT q; Pack<T, VecOf<T>::V> r; } u;
    u.q = *reinterpret_cast<const VT*>(p);
    return u.r;
}
template <typename T> __device__ __forceinline__ void stv(T* p, const Pack<T, VecOf<T>::V>& r) {
    typedef typename VecOf<T>::type VT;
    union { VT q; Pack<T, VecOf<T>::V> r; } u;
    u.r = r;
    *reinterpret_cast<VT*>(p) = u.q;
}
// V+1 elements: the vector plus the first element of the next one (the k+1 neighbour of lane V-1)
template <typename T> struct PackP { T v[VecOf<T>::V + 1]; };
template <typename T> __device__ __forceinline__ PackP<T> ldvp(const T* p) {
    PackP<T> r;
    Pack<T, VecOf<T>::V> a = ldv<T>(p);
#pragma unroll
    for (int e = 0; e < VecOf<T>::V; ++e) r.v[e] = a.v[e];
    r.v[VecOf<T>::V] = p[VecOf<T>::V];
    return r;
}

// =================================================================================================
// 3-D two-pass kernels.  One thread owns V consecutive k of one (i,j) row.
// grid: x = ceil(pz/V / bx), y = ceil(ny / by), z = planes in [i_begin, i_end)
// =================================================================================================
template <typename T, bool HET>
__global__ void __launch_bounds__(256)
k_h3d(Fields<T> f, Coefs<T> c, Geom g, int i_begin)
{
    constexpr int V = VecOf<T>::V;
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) * V;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int i = i_begin + blockIdx.z;
    if (k >= g.pz || j >= g.ny) return;
    const long long o = (long long)i * g.sx + (long long)j * g.sy + k;

    const PackP<T> ex = ldvp<T>(f.ex + o), ey = ldvp<T>(f.ey + o);
    const Pack<T, V> ez = ldv<T>(f.ez + o);
    const Pack<T, V> ez_j = ldv<T>(f.ez + o + g.sy), ex_j = ldv<T>(f.ex + o + g.sy);
    const Pack<T, V> ez_i = ldv<T>(f.ez + o + g.sx), ey_i = ldv<T>(f.ey + o + g.sx);
    Pack<T, V> hx = ldv<T>(f.hx + o), hy = ldv<T>(f.hy + o), hz = ldv<T>(f.hz + o);

    PackP<T> da, db;
    Pack<T, V> da_i, db_i, da_j, db_j;
    if (HET) {
        da = ldvp<T>(c.da + o); db = ldvp<T>(c.db + o);
        da_i = ldv<T>(c.da + o + g.sx); db_i = ldv<T>(c.db + o + g.sx);
        da_j = ldv<T>(c.da + o + g.sy); db_j = ldv<T>(c.db + o + g.sy);
    }
    const int gi = g.x0 + i;
    const bool ix1 = gi < g.nxg - 1, ix2 = gi < g.nxg - 2;
    const bool jy1 = j < g.ny - 1, jy2 = j < g.ny - 2;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const bool kz1 = (k + e) < g.nz - 1, kz2 = (k + e) < g.nz - 2;
        T a, b;
        // Hx (solver.py:178-205): j < ny-2, k < nz-2, every plane of the (nx-1)-plane array
        if (HET) { a = mean2<T>(da.v[e], da_i.v[e]); b = mean2<T>(db.v[e], db_i.v[e]); }
        else { a = c.uda; b = c.udb; }
        T n = upd_h<T>(a, hx.v[e], b, Ar<T>::diff(ez_j.v[e], ez.v[e], g.dy, g.rdy),
                       Ar<T>::diff(ey.v[e + 1], ey.v[e], g.dz, g.rdz));
        if (ix1 && jy2 && kz2) hx.v[e] = n;
        // Hy (:212-229): i < nx-2, k < nz-2
        if (HET) { a = mean2<T>(da.v[e], da_j.v[e]); b = mean2<T>(db.v[e], db_j.v[e]); }
        n = upd_h<T>(a, hy.v[e], b, Ar<T>::diff(ex.v[e + 1], ex.v[e], g.dz, g.rdz),
                     Ar<T>::diff(ez_i.v[e], ez.v[e], g.dx, g.rdx));
        if (ix2 && jy1 && kz2) hy.v[e] = n;
        // Hz (:236-253): i < nx-2, j < ny-2
        if (HET) { a = mean2<T>(da.v[e], da.v[e + 1]); b = mean2<T>(db.v[e], db.v[e + 1]); }
        n = upd_h<T>(a, hz.v[e], b, Ar<T>::diff(ey_i.v[e], ey.v[e], g.dx, g.rdx),
                     Ar<T>::diff(ex_j.v[e], ex.v[e], g.dy, g.rdy));
        if (ix2 && jy2 && kz1) hz.v[e] = n;
    }
    stv<T>(f.hx + o, hx); stv<T>(f.hy + o, hy); stv<T>(f.hz + o, hz);
}

template <typename T, bool HET>
__global__ void __launch_bounds__(256)
k_e3d(Fields<T> f, Coefs<T> c, Geom g, int i_begin)
{
    constexpr int V = VecOf<T>::V;
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) * V;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int i = i_begin + blockIdx.z;
    if (k >= g.pz || j >= g.ny) return;
    const long long o = (long long)i * g.sx + (long long)j * g.sy + k;

    const PackP<T> hx = ldvp<T>(f.hx + o), hy = ldvp<T>(f.hy + o);
    const Pack<T, V> hz = ldv<T>(f.hz + o);
    const Pack<T, V> hz_j = ldv<T>(f.hz + o + g.sy), hx_j = ldv<T>(f.hx + o + g.sy);
    const Pack<T, V> hz_i = ldv<T>(f.hz + o + g.sx), hy_i = ldv<T>(f.hy + o + g.sx);
    Pack<T, V> ex = ldv<T>(f.ex + o), ey = ldv<T>(f.ey + o), ez = ldv<T>(f.ez + o);

    PackP<T> ca, cb, ca_j, cb_j, ca_i, cb_i;
    Pack<T, V> ca_ij, cb_ij;
    if (HET) {
        ca = ldvp<T>(c.ca + o); cb = ldvp<T>(c.cb + o);
        ca_j = ldvp<T>(c.ca + o + g.sy); cb_j = ldvp<T>(c.cb + o + g.sy);
        ca_i = ldvp<T>(c.ca + o + g.sx); cb_i = ldvp<T>(c.cb + o + g.sx);
        ca_ij = ldv<T>(c.ca + o + g.sx + g.sy); cb_ij = ldv<T>(c.cb + o + g.sx + g.sy);
    }
    // per-component Cb (opt-in diagonal anisotropy): Ey averages Cb_y over x and z, Ez averages Cb_z over x and y
    PackP<T> cby, cby_i;
    Pack<T, V> cbz, cbz_i, cbz_j, cbz_ij;
    if (HET) {
        if (c.cby) {
            cby = ldvp<T>(c.cby + o); cby_i = ldvp<T>(c.cby + o + g.sx);
            cbz = ldv<T>(c.cbz + o); cbz_i = ldv<T>(c.cbz + o + g.sx);
            cbz_j = ldv<T>(c.cbz + o + g.sy); cbz_ij = ldv<T>(c.cbz + o + g.sx + g.sy);
        } else {
            cby = cb; cby_i = cb_i; cbz_ij = cb_ij;
#pragma unroll
            for (int e = 0; e < V; ++e) { cbz.v[e] = cb.v[e]; cbz_i.v[e] = cb_i.v[e]; cbz_j.v[e] = cb_j.v[e]; }
        }
    }
    const int gi = g.x0 + i;
    const bool ix0 = gi < g.nxg, ix1 = gi < g.nxg - 1;
    const bool jy0 = j < g.ny, jy1 = j < g.ny - 1;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const bool kz0 = (k + e) < g.nz, kz1 = (k + e) < g.nz - 1;
        T a, b;
        // Ex (solver.py:265-275): averaged over y and z
        if (HET) {
            a = mean4<T>(ca.v[e], ca_j.v[e], ca.v[e + 1], ca_j.v[e + 1]);
            b = mean4<T>(cb.v[e], cb_j.v[e], cb.v[e + 1], cb_j.v[e + 1]);
        } else { a = c.uca; b = c.ucb; }
        T n = upd_e<T>(a, ex.v[e], b, Ar<T>::diff(hz_j.v[e], hz.v[e], g.dy, g.rdy),
                       Ar<T>::diff(hy.v[e + 1], hy.v[e], g.dz, g.rdz));
        if (ix0 && jy1 && kz1) ex.v[e] = n;
        // Ey (:282-292): averaged over x and z
        if (HET) {
            a = mean4<T>(ca.v[e], ca_i.v[e], ca.v[e + 1], ca_i.v[e + 1]);
            b = mean4<T>(cby.v[e], cby_i.v[e], cby.v[e + 1], cby_i.v[e + 1]);
        }
        n = upd_e<T>(a, ey.v[e], b, Ar<T>::diff(hx.v[e + 1], hx.v[e], g.dz, g.rdz),
                     Ar<T>::diff(hz_i.v[e], hz.v[e], g.dx, g.rdx));
        if (ix1 && jy0 && kz1) ey.v[e] = n;
        // Ez (:299-309): averaged over x and y
        if (HET) {
            a = mean4<T>(ca.v[e], ca_i.v[e], ca_j.v[e], ca_ij.v[e]);
            b = mean4<T>(cbz.v[e], cbz_i.v[e], cbz_j.v[e], cbz_ij.v[e]);
        }
        n = upd_e<T>(a, ez.v[e], b, Ar<T>::diff(hy_i.v[e], hy.v[e], g.dx, g.rdx),
                     Ar<T>::diff(hx_j.v[e], hx.v[e], g.dy, g.rdy));
        if (ix1 && jy1 && kz0) ez.v[e] = n;
    }
    stv<T>(f.ex + o, ex); stv<T>(f.ey + o, ey); stv<T>(f.ez + o, ez);
}

// =================================================================================================
// 2-D kernels (y contiguous).  One thread per (i,j).  Gates (solver.py:321,367) come from exact
// device-side counts of non-zero / NaN cells per E component: cnt[0..2] = #(|F|>0), cnt[3..5] = #NaN.
// =================================================================================================
__device__ __forceinline__ bool gate_on(const int* cnt, int comp) {
    // np.max(np.abs(F)) > 0  <=>  no NaN anywhere (max would be NaN) and some |F| > 0
    return cnt[3 + comp] == 0 && cnt[comp] > 0;
}

template <typename T, bool HET>
__global__ void __launch_bounds__(256)
k_h2d(Fields<T> f, Coefs<T> c, Geom g, const int* __restrict__ cnt_cur, int* __restrict__ cnt_next)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 6 && cnt_next) cnt_next[threadIdx.x] = 0;
    if (j >= g.ny) return;
    const bool g1 = gate_on(cnt_cur, 2);
    const bool g2 = gate_on(cnt_cur, 0) || gate_on(cnt_cur, 1);
    const long long o = (long long)i * g.sx + j;
    const bool ix1 = i < g.nx - 1, ix2 = i < g.nx - 2;
    const bool jy1 = j < g.ny - 1, jy2 = j < g.ny - 2;
    T da = c.uda, db = c.udb, da_i = c.uda, db_i = c.udb, da_j = c.uda, db_j = c.udb;
    if (HET) {
        da = c.da[o]; db = c.db[o];
        da_i = c.da[o + g.sx]; db_i = c.db[o + g.sx];
        da_j = c.da[o + 1]; db_j = c.db[o + 1];
    }
    if (g1) {
        const T ez = f.ez[o];
        if (ix1 && jy2) {   // Hx[:, :ny-2] = da*Hx + db*dEz/dy   (sign kept, solver.py:338-341)
            const T a = HET ? mean2<T>(da, da_i) : da, b = HET ? mean2<T>(db, db_i) : db;
            f.hx[o] = Ar<T>::add(Ar<T>::mul(a, f.hx[o]),
                                 Ar<T>::mul(b, Ar<T>::diff(f.ez[o + 1], ez, g.dy, g.rdy)));
        }
        if (ix2 && jy1) {   // Hy[:nx-2, :] = da*Hy - db*dEz/dx     (:358-361)
            const T a = HET ? mean2<T>(da, da_j) : da, b = HET ? mean2<T>(db, db_j) : db;
            f.hy[o] = Ar<T>::sub(Ar<T>::mul(a, f.hy[o]),
                                 Ar<T>::mul(b, Ar<T>::diff(f.ez[o + g.sx], ez, g.dx, g.rdx)));
        }
    }
    if (g2) {               // whole Hz array, curls zero-padded, coefficients not averaged (:371-397)
        const T cey = ix2 ? Ar<T>::diff(f.ey[o + g.sx], f.ey[o], g.dx, g.rdx) : (T)0;
        const T cex = jy2 ? Ar<T>::diff(f.ex[o + 1], f.ex[o], g.dy, g.rdy) : (T)0;
        f.hz[o] = upd_h<T>(da, f.hz[o], db, cey, cex);
    }
}

template <typename T> __device__ __forceinline__ void count_cell(T v, int& nz, int& nn) {
    nz += (fabs((double)v) > 0.0) ? 1 : 0;
    nn += (v != v) ? 1 : 0;
}

__device__ __forceinline__ void block_count_flush(int* cnt, int comp, int nz, int nn) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        nz += __shfl_xor_sync(0xffffffffu, nz, s);
        nn += __shfl_xor_sync(0xffffffffu, nn, s);
    }
    if ((threadIdx.x & 31) == 0) {
        if (nz) atomicAdd(cnt + comp, nz);
        if (nn) atomicAdd(cnt + 3 + comp, nn);
    }
}

template <typename T, bool HET>
__global__ void __launch_bounds__(256)
k_e2d(Fields<T> f, Coefs<T> c, Geom g, int* __restrict__ cnt_next)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    int nz0 = 0, nn0 = 0, nz1 = 0, nn1 = 0, nz2 = 0, nn2 = 0;
    if (j < g.ny) {
        const long long o = (long long)i * g.sx + j;
        const bool ix1 = i < g.nx - 1, jy1 = j < g.ny - 1;
        T ca = c.uca, cb = c.ucb, ca_i = c.uca, cb_i = c.ucb, ca_j = c.uca, cb_j = c.ucb,
          ca_ij = c.uca, cb_ij = c.ucb;
        if (HET) {
            ca = c.ca[o]; cb = c.cb[o];
            ca_i = c.ca[o + g.sx]; cb_i = c.cb[o + g.sx];
            ca_j = c.ca[o + 1]; cb_j = c.cb[o + 1];
            ca_ij = c.ca[o + g.sx + 1]; cb_ij = c.cb[o + g.sx + 1];
        }
        const T hz = f.hz[o];
        if (ix1 && jy1) {   // Ez (solver.py:409-423)
            const T a = HET ? mean4<T>(ca, ca_i, ca_j, ca_ij) : ca;
            const T b = HET ? mean4<T>(cb, cb_i, cb_j, cb_ij) : cb;
            const T hy = f.hy[o], hx = f.hx[o];
            const T n = upd_e<T>(a, f.ez[o], b, Ar<T>::diff(f.hy[o + g.sx], hy, g.dx, g.rdx),
                                 Ar<T>::diff(f.hx[o + 1], hx, g.dy, g.rdy));
            f.ez[o] = n; count_cell<T>(n, nz2, nn2);
        }
        if (jy1) {          // Ex = ca*Ex + cb*dHz/dy (:433-442)
            const T a = HET ? mean2<T>(ca, ca_j) : ca, b = HET ? mean2<T>(cb, cb_j) : cb;
            const T n = Ar<T>::add(Ar<T>::mul(a, f.ex[o]),
                                   Ar<T>::mul(b, Ar<T>::diff(f.hz[o + 1], hz, g.dy, g.rdy)));
            f.ex[o] = n; count_cell<T>(n, nz0, nn0);
        }
        if (ix1) {          // Ey = ca*Ey - cb*dHz/dx (:447-456)
            const T a = HET ? mean2<T>(ca, ca_i) : ca, b = HET ? mean2<T>(cb, cb_i) : cb;
            const T n = Ar<T>::sub(Ar<T>::mul(a, f.ey[o]),
                                   Ar<T>::mul(b, Ar<T>::diff(f.hz[o + g.sx], hz, g.dx, g.rdx)));
            f.ey[o] = n; count_cell<T>(n, nz1, nn1);
        }
    }
    if (cnt_next) {
        block_count_flush(cnt_next, 0, nz0, nn0);
        block_count_flush(cnt_next, 1, nz1, nn1);
        block_count_flush(cnt_next, 2, nz2, nn2);
    }
}

// count non-zero / NaN cells of the three E arrays (2-D gates), whole padded array (padding is 0)
template <typename T>
__global__ void k_count2d(CFields<T> f, long long n, int* __restrict__ cnt)
{
    int nz0 = 0, nn0 = 0, nz1 = 0, nn1 = 0, nz2 = 0, nn2 = 0;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x) {
        count_cell<T>(f.ex[t], nz0, nn0);
        count_cell<T>(f.ey[t], nz1, nn1);
        count_cell<T>(f.ez[t], nz2, nn2);
    }
    block_count_flush(cnt, 0, nz0, nn0);
    block_count_flush(cnt, 1, nz1, nn1);
    block_count_flush(cnt, 2, nz2, nn2);
}

// =================================================================================================
// sources and monitors
// =================================================================================================
struct SrcOp {
    int comp;
    int lo[3], n[3];          // box origin and extent (n[2] = 1 in 2-D)
    int table;
    long long prof_off;       // offset into the profile pool, -1 = uniform
    double divisor;
    long long first_thread;   // prefix of cells over the ops of one launch
};
struct MonOp {
    int comp;
    int lo[3], n[3];
    int record, n_freq, phasor_col;
    long long rec_off;        // offset into the record pool (elements of T) of step 0
    long long dft_off;        // offset into the dft pool (complex = 2 doubles)
    long long cells;
    long long first_thread;
};

struct Strides3 { long long s[3]; };   // 3-D: (sx, sy, 1); 2-D: (sx, 1, 0)

// F[box] += amp[step][table] (* profile / divisor).  Arithmetic in fp64, rounded once to T on store
// to T on store.  Keeps the 2-D gate counters exact.
template <typename T>
__global__ void k_sources(T* const* __restrict__ comp_ptr, const SrcOp* __restrict__ ops, int n_ops,
                          long long total, Strides3 st, const double* __restrict__ amp, int n_amp,
                          const int* __restrict__ step_ptr, int step_off,
                          const double* __restrict__ prof, int* __restrict__ cnt)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= total) return;
    int q = 0;
    while (q + 1 < n_ops && ops[q + 1].first_thread <= t) ++q;
    const SrcOp op = ops[q];
    const long long cell = t - op.first_thread;
    const int c2 = (int)(cell % op.n[2]);
    const long long r = cell / op.n[2];
    const int c1 = (int)(r % op.n[1]);
    const int c0 = (int)(r / op.n[1]);
    const long long o = (op.lo[0] + c0) * st.s[0] + (op.lo[1] + c1) * st.s[1] + (op.lo[2] + c2) * st.s[2];
    const int step = *step_ptr + step_off;
    double a = amp[(long long)step * n_amp + op.table];
    if (op.prof_off >= 0) {
        a = __dmul_rn(a, prof[op.prof_off + cell]);
        if (op.divisor != 1.0) a = __ddiv_rn(a, op.divisor);
    }
    T* p = comp_ptr[op.comp] + o;
    const T old = *p;
    const T nv = (T)__dadd_rn((double)old, a);
    *p = nv;
    if (cnt && op.comp < 3) {
        const int dz = ((fabs((double)nv) > 0.0) ? 1 : 0) - ((fabs((double)old) > 0.0) ? 1 : 0);
        const int dn = ((nv != nv) ? 1 : 0) - ((old != old) ? 1 : 0);
        if (dz) atomicAdd(cnt + op.comp, dz);
        if (dn) atomicAdd(cnt + 3 + op.comp, dn);
    }
}

// record the box and / or accumulate the running DFT  acc[f] += (F*ph)*dt  (monitors/field.py:139-143)
template <typename T>
__global__ void k_monitors(const T* const* __restrict__ comp_ptr, const MonOp* __restrict__ ops,
                           int n_ops, long long total, Strides3 st, const double* __restrict__ phasors,
                           int n_phasor, const int* __restrict__ step_ptr, int step_off, double dt,
                           T* __restrict__ rec, double2* __restrict__ dft)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= total) return;
    int q = 0;
    while (q + 1 < n_ops && ops[q + 1].first_thread <= t) ++q;
    const MonOp op = ops[q];
    const long long cell = t - op.first_thread;
    const int c2 = (int)(cell % op.n[2]);
    const long long r = cell / op.n[2];
    const int c1 = (int)(r % op.n[1]);
    const int c0 = (int)(r / op.n[1]);
    const long long o = (op.lo[0] + c0) * st.s[0] + (op.lo[1] + c1) * st.s[1] + (op.lo[2] + c2) * st.s[2];
    const int step = *step_ptr + step_off;
    const T v = comp_ptr[op.comp][o];
    if (op.record) rec[op.rec_off + (long long)step * op.cells + cell] = v;
    if (op.n_freq > 0) {
        const double d = (double)v;
        const double* ph = phasors + ((long long)step * n_phasor + op.phasor_col) * 2;
        for (int fq = 0; fq < op.n_freq; ++fq) {
            double2* a = dft + op.dft_off + (long long)fq * op.cells + cell;
            double2 acc = *a;
            acc.x = __dadd_rn(acc.x, __dmul_rn(__dmul_rn(d, ph[2 * fq]), dt));
            acc.y = __dadd_rn(acc.y, __dmul_rn(__dmul_rn(d, ph[2 * fq + 1]), dt));
            *a = acc;
        }
    }
}

// Auxiliary-differential-equation recursions of dispersive media, cell-local, driven by one E component
// (materials/ade.py:116-160; coefficients materials/dispersion.py:189-231, 267-286, 323-336):
//   Lorentz  P+ = C0*E + C1*E + C2*P + C3*P_prev      Drude  J+ = C0*E + C1*J      Debye  P+ = C0*E + C1*P
// evaluated left to right like NumPy; "E" is E*mask as ADEManager.update_all forms it (ade.py:291-308).
struct AdeOp {
    int comp, kind;              // kind: 0 Lorentz, 1 Drude, 2 Debye
    int lo[3], n[3];
    double c0, c1, c2, c3;
    long long cur_off, prev_off; // offsets (elements) into the aux pool; prev_off < 0 unless Lorentz
    long long mask_off;          // offset into the mask pool, -1 = no mask
    long long cells, first_thread;
};

// one recursion step of one cell; returns the new state and leaves the old one in `old`
template <typename T>
__device__ __forceinline__ double ade_update(const AdeOp& op, long long cell, double e, T* __restrict__ aux,
                                             const unsigned char* __restrict__ mask, double& old)
{
    if (op.mask_off >= 0) e = __dmul_rn(e, mask[op.mask_off + cell] ? 1.0 : 0.0);
    T* cur = aux + op.cur_off + cell;
    const double a = (double)*cur;
    double nv;
    if (op.kind == 0) {
        T* prev = aux + op.prev_off + cell;
        nv = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(op.c0, e), __dmul_rn(op.c1, e)), __dmul_rn(op.c2, a)),
                       __dmul_rn(op.c3, (double)*prev));
        *prev = (T)a;
    } else {
        nv = __dadd_rn(__dmul_rn(op.c0, e), __dmul_rn(op.c1, a));
    }
    *cur = (T)nv;
    old = a;
    return nv;
}

template <typename T>
__global__ void k_ade(const T* const* __restrict__ comp_ptr, const AdeOp* __restrict__ ops, int n_ops,
                      long long total, Strides3 st, T* __restrict__ aux, const unsigned char* __restrict__ mask)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= total) return;
    int q = 0;
    while (q + 1 < n_ops && ops[q + 1].first_thread <= t) ++q;
    const AdeOp op = ops[q];
    const long long cell = t - op.first_thread;
    const int c2 = (int)(cell % op.n[2]);
    const long long r = cell / op.n[2];
    const int c1 = (int)(r % op.n[1]);
    const int c0 = (int)(r / op.n[1]);
    const long long o = (op.lo[0] + c0) * st.s[0] + (op.lo[1] + c1) * st.s[1] + (op.lo[2] + c2) * st.s[2];
    double old;
    ade_update<T>(op, cell, (double)comp_ptr[op.comp][o], aux, mask, old);
}

// The same recursions INSIDE a fused sweep (north-star subsystem 4).  The sweep that computes step n+1 holds the final
// E of step n (sources included) of every owner cell in registers exactly once — as the input of its E stage — so the
// recursion "after step n" (ADEManager.update_all, materials/ade.py:291-308) is applied right there, one step late in
// wall-clock order but on the same values: bit-identical to k_ade, and the E array is not read a second time.  The
// step whose successor sweep is not known yet (last step of an fdtd_run call / of a graph) still uses k_ade.
//   coupled != 0 (opt-in, NOT the reference's behaviour, which never feeds P back): the polarisation current of the
//   recursion is subtracted in the same E update,  E+ -= Cb * eps0 * (P+ - P) / dt  (Lorentz, Debye)  or
//   Cb * eps0 * J+  (Drude) — the reference's recursion coefficients (dispersion.py:189-336) carry no eps0, so P and J
//   are in units of eps0 —, with the recursion applied at the BEGINNING of every step (oracle/ade.py: coupled_step).
// order: optional permutation of the sweep's work items (x-segment, tile) that dispatches the items which meet a recursion
// box FIRST — they run a few times longer than the others and would otherwise set the length of the launch's tail.
struct AdeIn { const AdeOp* ops; int n; void* aux; const unsigned char* mask; int coupled; double kp, kj; const int* order; };   // kp = eps0/dt, kj = eps0

constexpr int kAdeSmemOps = 24;          // recursion descriptors kept in shared memory per CTA (bits 0..23 of the thread mask)

// Which recursions can ever touch this thread's cells (row j, cells k .. k+V-1, planes [i0, i1))?  A thread's (j, k) is fixed
// for the whole sweep, so this is evaluated once per kernel: the plane loop of the 97 % of threads outside every box
// then pays one register test (without it the c3 workload ran 2.4x slower: profiles/r02_tuning.md).
template <int V>
__device__ __forceinline__ unsigned ade_thread_mask(const AdeIn& ad, int i0, int i1, int j, int k)
{
    unsigned m = 0;
    const int n = ad.n < kAdeSmemOps ? ad.n : kAdeSmemOps;
    for (int q = 0; q < n; ++q) {
        const AdeOp& op = ad.ops[q];
        if ((unsigned)(j - op.lo[1]) < (unsigned)op.n[1] && k + V > op.lo[2] && k < op.lo[2] + op.n[2] &&
            i1 > op.lo[0] && i0 < op.lo[0] + op.n[0])
            m |= 1u << q;
    }
    if (ad.n > kAdeSmemOps) m |= 0x80000000u;   // more recursions than the shared copy holds: the tail is tested op by op
    return m;
}

// CTA-level version of the same test: does any recursion touch the owner cells of a (j, k) tile on planes [i0, i1)?  CTAs
// that answer no run the sweep body compiled WITHOUT the recursion code (its presence alone costs registers and
// memory-level parallelism: the c3 sweep ran 2.3x slower with it in every CTA, profiles/r02_tuning.md).
__device__ __forceinline__ bool ade_tile_touched(const AdeIn& ad, int i0, int i1, int j0, int j1, int k0, int k1)
{
    for (int q = 0; q < ad.n; ++q) {
        const AdeOp& op = ad.ops[q];
        if (j1 > op.lo[1] && j0 < op.lo[1] + op.n[1] && k1 > op.lo[2] && k0 < op.lo[2] + op.n[2] &&
            i1 > op.lo[0] && i0 < op.lo[0] + op.n[0])
            return true;
    }
    return false;
}

// One plane of one thread.  `ops` is the CTA's shared-memory copy of the recursion descriptors (ade_stage_ops): the
// first version read them from global memory op by op and state by state, a chain of dependent L1 / DRAM latencies in
// the only warps of the CTA that do this work — every other warp waits for them at the per-plane barrier, and the few
// CTAs that meet a box set the length of the whole launch (c3: 2.0 ms against 0.87 ms, profiles/r02_tuning.md).  Now the
// (up to three) recursions that can touch the thread are resolved to static slots, ALL their state loads are issued
// first, then the arithmetic and the stores follow; further recursions (rare) take the sequential path.
__device__ __forceinline__ void ade_stage_ops(AdeOp* s_ops, const AdeIn& ad, int tid, int nthreads)
{
    const int words = min(ad.n, kAdeSmemOps) * (int)(sizeof(AdeOp) / 4);
    const unsigned* src = reinterpret_cast<const unsigned*>(ad.ops);
    unsigned* dst = reinterpret_cast<unsigned*>(s_ops);
    for (int w = tid; w < words; w += nthreads) dst[w] = src[w];
}

template <typename T, int V>
__device__ __forceinline__ void ade_one(const AdeIn& ad, const AdeOp& op, int i, int j, int k, const Pack<T, V>& ex,
                                        const Pack<T, V>& ey, const Pack<T, V>& ez, double (&jx)[V], double (&jy)[V], double (&jz)[V])
{
    const int a0 = i - op.lo[0], a1 = j - op.lo[1];
    if ((unsigned)a0 >= (unsigned)op.n[0] || (unsigned)a1 >= (unsigned)op.n[1]) return;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const int a2 = k + e - op.lo[2];
        if ((unsigned)a2 >= (unsigned)op.n[2]) continue;
        const long long cell = ((long long)a0 * op.n[1] + a1) * op.n[2] + a2;
        const double ev = (double)(op.comp == 0 ? ex.v[e] : (op.comp == 1 ? ey.v[e] : ez.v[e]));
        double old;
        const double nv = ade_update<T>(op, cell, ev, (T*)ad.aux, ad.mask, old);
        if (ad.coupled) {
            const double term = op.kind == 1 ? __dmul_rn(nv, ad.kj) : __dmul_rn(__dsub_rn(nv, old), ad.kp);
            if (op.comp == 0) jx[e] = __dadd_rn(jx[e], term);
            else if (op.comp == 1) jy[e] = __dadd_rn(jy[e], term);
            else jz[e] = __dadd_rn(jz[e], term);
        }
    }
}

template <typename T, int V>
__device__ __forceinline__ void ade_in_sweep(const AdeIn& ad, const AdeOp* __restrict__ ops, unsigned mask, int i, int j, int k,
                                             const Pack<T, V>& ex, const Pack<T, V>& ey, const Pack<T, V>& ez, double (&jx)[V],
                                             double (&jy)[V], double (&jz)[V])
{
    constexpr int K = 3;
    T* __restrict__ aux = (T*)ad.aux;
    unsigned mm = mask & 0x7fffffffu;
    int q[K];
#pragma unroll
    for (int m = 0; m < K; ++m) { q[m] = mm ? (int)__ffs(mm) - 1 : -1; mm &= mm - 1u; }
    // ---- phase 1: which cells, and every state value they need --------------------------------------------------------------
    unsigned cell[K];                     // a recursion box has fewer than 2^32 cells (array_elems < 2^32 on this path)
    unsigned em[K];                       // bit e: cell e is inside the box; bit 8+e: its mask byte is zero
    T cur[K][V], prv[K][V];
#pragma unroll
    for (int m = 0; m < K; ++m) {
        em[m] = 0;
        if (q[m] < 0) continue;
        const AdeOp& op = ops[q[m]];
        const int a0 = i - op.lo[0], a1 = j - op.lo[1];
        if ((unsigned)a0 >= (unsigned)op.n[0] || (unsigned)a1 >= (unsigned)op.n[1]) continue;
        cell[m] = ((unsigned)a0 * (unsigned)op.n[1] + (unsigned)a1) * (unsigned)op.n[2] + (unsigned)(k - op.lo[2]);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            if ((unsigned)(k + e - op.lo[2]) >= (unsigned)op.n[2]) continue;
            em[m] |= 1u << e;
            const unsigned ce = cell[m] + (unsigned)e;      // modulo 2^32 on purpose: k - lo may be -1 for the cell before the box
            cur[m][e] = aux[op.cur_off + ce];
            prv[m][e] = op.kind == 0 ? aux[op.prev_off + ce] : (T)0;
            if (op.mask_off >= 0 && !ad.mask[op.mask_off + ce]) em[m] |= 1u << (8 + e);
        }
    }
    // ---- phase 2: the recursions (same operation order as ade_update / k_ade), stores, feedback terms -----------------------
#pragma unroll
    for (int m = 0; m < K; ++m) {
        if (!em[m]) continue;
        const AdeOp& op = ops[q[m]];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            if (!((em[m] >> e) & 1u)) continue;
            double ev = (double)(op.comp == 0 ? ex.v[e] : (op.comp == 1 ? ey.v[e] : ez.v[e]));
            if (op.mask_off >= 0) ev = __dmul_rn(ev, ((em[m] >> (8 + e)) & 1u) ? 0.0 : 1.0);
            const double a = (double)cur[m][e];
            double nv;
            if (op.kind == 0) {
                nv = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(op.c0, ev), __dmul_rn(op.c1, ev)), __dmul_rn(op.c2, a)),
                               __dmul_rn(op.c3, (double)prv[m][e]));
                aux[op.prev_off + (cell[m] + (unsigned)e)] = (T)a;
            } else {
                nv = __dadd_rn(__dmul_rn(op.c0, ev), __dmul_rn(op.c1, a));
            }
            aux[op.cur_off + (cell[m] + (unsigned)e)] = (T)nv;
            if (ad.coupled) {
                const double term = op.kind == 1 ? __dmul_rn(nv, ad.kj) : __dmul_rn(__dsub_rn(nv, a), ad.kp);
                if (op.comp == 0) jx[e] = __dadd_rn(jx[e], term);
                else if (op.comp == 1) jy[e] = __dadd_rn(jy[e], term);
                else jz[e] = __dadd_rn(jz[e], term);
            }
        }
    }
    // ---- more than three recursions on this thread, or more descriptors than the shared copy holds: one by one ---------------
    while (mm) {
        const int qq = (int)__ffs(mm) - 1;
        mm &= mm - 1u;
        ade_one<T, V>(ad, ops[qq], i, j, k, ex, ey, ez, jx, jy, jz);
    }
    if (mask & 0x80000000u)
        for (int qq = kAdeSmemOps; qq < ad.n; ++qq) ade_one<T, V>(ad, ad.ops[qq], i, j, k, ex, ey, ez, jx, jy, jz);
}

// Region-correct flux (extension): instantaneous power through a box, P = sum (E x H)_n over the
// box's cells, fields taken co-located like the reference's FluxMonitor does on its patch (monitors/flux.py:146-175).
// Deterministic: FLUX_BLOCKS fixed partial sums per op (grid-stride, tree reduce), then one thread adds them in order.
struct FluxOp { int dir; int lo[3], n[3]; long long cells; long long out_off; };
constexpr int FLUX_BLOCKS = 64;

template <typename T>
__global__ void __launch_bounds__(256)
k_flux_partial(const T* const* __restrict__ comp_ptr, const FluxOp* __restrict__ ops, Strides3 st, double* __restrict__ partial)
{
    const FluxOp op = ops[blockIdx.y];
    const T* ex = comp_ptr[0]; const T* ey = comp_ptr[1]; const T* ez = comp_ptr[2];
    const T* hx = comp_ptr[3]; const T* hy = comp_ptr[4]; const T* hz = comp_ptr[5];
    double acc = 0.0;
    for (long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x; cell < op.cells;
         cell += (long long)gridDim.x * blockDim.x) {
        const int c2 = (int)(cell % op.n[2]);
        const long long r = cell / op.n[2];
        const int c1 = (int)(r % op.n[1]);
        const int c0 = (int)(r / op.n[1]);
        const long long o = (op.lo[0] + c0) * st.s[0] + (op.lo[1] + c1) * st.s[1] + (op.lo[2] + c2) * st.s[2];
        double s;
        if (op.dir == 0) s = (double)ey[o] * (double)hz[o] - (double)ez[o] * (double)hy[o];
        else if (op.dir == 1) s = (double)ez[o] * (double)hx[o] - (double)ex[o] * (double)hz[o];
        else s = (double)ex[o] * (double)hy[o] - (double)ey[o] * (double)hx[o];
        acc += s;
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.y * FLUX_BLOCKS + blockIdx.x] = sh[0];
}

__global__ void k_flux_final(const FluxOp* __restrict__ ops, int n_ops, const double* __restrict__ partial,
                             double* __restrict__ out, const int* __restrict__ step_ptr, int step_off, int n_steps)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_ops) return;
    double s = 0.0;
    for (int b = 0; b < FLUX_BLOCKS; ++b) s += partial[q * FLUX_BLOCKS + b];
    out[ops[q].out_off * n_steps + (*step_ptr + step_off)] = s;
}

// Mode-overlap numerator over a monitor plane, per frequency (utils/mode_matching.py:41-131):
//   0.5 * sum_cells [ (e1*conj(m_h2) - e2*conj(m_h1)) + (m_e1*conj(h2) - m_e2*conj(h1)) ]
// (e1, h2, e2, h1) = the four tangential DFT planes the direction selects, m_* the mode's fields on the same cells.
// Deterministic: FLUX_BLOCKS partial sums per frequency (grid-stride, tree reduce), then added in block order.
struct OverlapIn { long long off[4]; long long cells; };      // offsets of the e1, h2, e2, h1 DFT planes of frequency 0

__device__ __forceinline__ double2 cmul_conj(double2 a, double2 b)     // a * conj(b)
{
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

__global__ void __launch_bounds__(256)
k_overlap_partial(const double2* __restrict__ dft, OverlapIn in, const double2* __restrict__ mode, double2* __restrict__ partial)
{
    const int f = blockIdx.y;
    const double2* e1 = dft + in.off[0] + f * in.cells; const double2* h2 = dft + in.off[1] + f * in.cells;
    const double2* e2 = dft + in.off[2] + f * in.cells; const double2* h1 = dft + in.off[3] + f * in.cells;
    const double2* me1 = mode; const double2* mh2 = mode + in.cells;
    const double2* me2 = mode + 2 * in.cells; const double2* mh1 = mode + 3 * in.cells;
    double2 acc = make_double2(0.0, 0.0);
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < in.cells; c += (long long)gridDim.x * blockDim.x) {
        const double2 a = cmul_conj(e1[c], mh2[c]), b = cmul_conj(e2[c], mh1[c]);
        const double2 p = cmul_conj(me1[c], h2[c]), q = cmul_conj(me2[c], h1[c]);
        acc.x += (a.x - b.x) + (p.x - q.x);
        acc.y += (a.y - b.y) + (p.y - q.y);
    }
    __shared__ double2 sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) { sh[threadIdx.x].x += sh[threadIdx.x + w].x; sh[threadIdx.x].y += sh[threadIdx.x + w].y; }
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[f * FLUX_BLOCKS + blockIdx.x] = sh[0];
}

__global__ void k_overlap_final(const double2* __restrict__ partial, int n_freq, double2* __restrict__ out)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_freq) return;
    double2 s = make_double2(0.0, 0.0);
    for (int b = 0; b < FLUX_BLOCKS; ++b) { s.x += partial[f * FLUX_BLOCKS + b].x; s.y += partial[f * FLUX_BLOCKS + b].y; }
    out[f] = make_double2(0.5 * s.x, 0.5 * s.y);
}

__global__ void k_bump(int* step_ptr, int n) { *step_ptr += n; }

// =================================================================================================
// layout: compact host-order chunk <-> padded device array (with dtype conversion)
// =================================================================================================
template <typename TD, typename TH>
__global__ void k_scatter(TD* __restrict__ dst, const TH* __restrict__ src, long long first, long long count,
                          int c1, int c2, Strides3 st)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < count;
         t += (long long)gridDim.x * blockDim.x) {
        const long long lin = first + t;
        const int k = (int)(lin % c2);
        const long long r = lin / c2;
        const int j = (int)(r % c1);
        const long long i = r / c1;
        dst[i * st.s[0] + j * st.s[1] + k * st.s[2]] = (TD)src[t];
    }
}
template <typename TD, typename TH>
__global__ void k_gather(TH* __restrict__ dst, const TD* __restrict__ src, long long first, long long count,
                         int c1, int c2, Strides3 st)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < count;
         t += (long long)gridDim.x * blockDim.x) {
        const long long lin = first + t;
        const int k = (int)(lin % c2);
        const long long r = lin / c2;
        const int j = (int)(r % c1);
        const long long i = r / c1;
        dst[t] = (TH)src[i * st.s[0] + j * st.s[1] + k * st.s[2]];
    }
}
// box [lo, lo + n) of one array -> dense fp64 (bench self-check, cropped parity tests)
template <typename TD>
__global__ void k_gather_box(double* __restrict__ dst, const TD* __restrict__ src, long long count, int l0, int l1, int l2,
                             int n1, int n2, Strides3 st)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < count;
         t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t % n2);
        const long long r = t / n2;
        const int j = (int)(r % n1);
        const long long i = r / n1;
        dst[t] = (double)src[(l0 + i) * st.s[0] + (l1 + j) * st.s[1] + (l2 + k) * st.s[2]];
    }
}
// Order-independent checksum of the logical cells of each x-plane: out[2p] = sum of the value bit patterns,
// out[2p+1] = sum of bits * (1 + cell index in the plane), both modulo 2^64 (integer atomics: deterministic, and
// independent of how planes are distributed over GPUs).  grid = (blocks per plane, planes).
template <typename TD>
__global__ void k_plane_checksum(unsigned long long* __restrict__ out, const TD* __restrict__ src, int c1, int c2, Strides3 st)
{
    const long long plane = blockIdx.y;
    const long long cells = (long long)c1 * c2;
    unsigned long long s1 = 0, s2 = 0;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < cells; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t % c2);
        const int j = (int)(t / c2);
        const TD v = src[plane * st.s[0] + j * st.s[1] + k * st.s[2]];
        unsigned long long b;
        if (sizeof(TD) == 8) b = (unsigned long long)__double_as_longlong((double)v);
        else b = (unsigned long long)__float_as_uint((float)v);
        s1 += b;
        s2 += b * (unsigned long long)(t + 1);
    }
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out + 2 * plane, s1);
        atomicAdd(out + 2 * plane + 1, s2);
    }
}
template <typename TD, typename TH>
__global__ void k_convert(TH* __restrict__ dst, const TD* __restrict__ src, long long n)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x)
        dst[t] = (TH)src[t];
}

}  // namespace fdtd
