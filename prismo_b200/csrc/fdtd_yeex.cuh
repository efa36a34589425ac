// fdtd_yeex.cuh — physics mode (stable Yee leap-frog + CPML) as ONE TMA-fed sweep per step: the north-star kernel
// ("fused 3-D E/H + CPML step").  Same arithmetic, update ranges and psi layout as the two-pass pair in fdtd_yee.cuh
// (which oracle/yee.py pins) and as the first fused version in fdtd_yee_fused.cuh; what changes is the plumbing, which
// is that of fdtd_tb2x.cuh:
//
//   * the six input arrays of plane p arrive by cp.async.bulk.tensor (one producer thread, S-stage ring, full / empty
//     mbarriers).  The box carries the tile's low-side halo — box row 0 is j0 - 1, box lane 0 is k0 - 2 lanes, zero-
//     filled outside the grid — so the -j neighbours of E that the backward-differenced H update needs are plain
//     shared-memory reads of the staged tile and cost no warp: of R = 15 consumer rows 14 own cells, the 15th only
//     provides H+ of the next tile's first row;
//   * the +j neighbours of H+ (z, x components) that the E update needs travel through a D-slot ring with one mbarrier
//     per row: neighbour-only synchronisation, no CTA-wide barrier in the plane loop;
//   * CPML is applied by slab threads only, and decided per WARP and PLANE: a warp whose row lies outside the y slabs,
//     in a tile outside the z slabs and clear of the grid faces, on planes outside the x slabs, runs a lean body with
//     no masks and no psi code at all (about 85 % of the warp-iterations of a 1024^3 grid with 10-cell layers); the
//     others run the full body (masks + psi recursion), whose psi arrays are ping-ponged like the fields;
//   * the psi values a slab thread needs are LOADED ONE ITERATION AHEAD, all at once (x / y families as 8-byte vectors,
//     z family per cell), into registers that the stage which consumed the previous set has just freed: the first
//     version loaded each psi value right where the recursion used it, one exposed DRAM latency per derivative, and
//     spent 30 % of its stall samples there (profiles/r02_ncu_yeex.md).
//
//   H+[p] = f(H[p], E[p] (own, j-1, k-1), E[p-1])          E+[q] = g(E[q], H+[q] (own, j+1, k+1), H+[q+1])
// Iteration `it` (i = i0 - 1 + it) reads TMA stage it + 1 (E[i+1], H[i+1]), produces H+[i+1] and E+[i].  Of 32 lanes 30
// own cells (lane 0 and lane 31 are halo providers: -k comes by shuffle from the lane below, +k from the lane above).
//
// PARITY UNPINNED (there are no reference numbers for a working CPML): validated bitwise in fp64 against the two-pass
// kernels and through them against oracle/yee.py.
#pragma once
#include "fdtd_tb2x.cuh"
#include "fdtd_yee_fused.cuh"

namespace fdtd {

// consumer rows per CTA (+ 1 producer warp): 15 rows = 512 threads at 128 registers.  Measured alternative for fp32: 19 rows =
// 640 threads at 96 registers (5 warps per scheduler, 18 of 19 rows own cells) spills 120 B and runs at 78.5 instead of
// 90.2 Gcell/s (profiles/r02_tuning.md §3).
template <typename T> struct YeexRows { static constexpr int R = 15; };
template <typename T> constexpr int kYeexBoxRowsOf = YeexRows<T>::R + 1;
// TMA box: 272 B x (R + 1) rows.  The first cell a tile CONSUMES is one lane (8 B) left of its first owner lane, i.e. at
// byte 240 * tk - 8 of the row, but the box origin of cp.async.bulk.tensor must be 16-byte aligned in the innermost
// dimension (a misaligned origin raises "illegal instruction": measured, gpurun_out/e2_sanitize_memcheck_yee.log), so the
// box starts one more lane to the left and lane l reads its cells at byte (l + 1) * 8 of the staged row.
constexpr int kYeexBoxBytes = 272;
template <int R> __host__ __device__ constexpr size_t yeex_tile_bytes() { return ((size_t)(R + 1) * kYeexBoxBytes + 127) / 128 * 128; }
template <int R> __host__ __device__ constexpr size_t yeex_stage_bytes() { return 6 * yeex_tile_bytes<R>(); }

template <int R> static inline size_t yeex_smem_bytes(int S, int D)
{
    return 128 + (size_t)S * yeex_stage_bytes<R>() + (size_t)D * R * 2 * kTb2xRowBytes + (size_t)(2 * S + R * D) * 8 + 64;
}

// the producer thread has nothing else to do: let the hardware park it for up to `ns` per try instead of spinning
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity, uint32_t ns)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAITP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONEP;\n"
        "bra LAB_WAITP;\n"
        "DONEP:\n"
        "}" ::"r"(bar), "r"(parity), "r"(ns) : "memory");
}

// ---- lean bodies: interior cells, no masks, no CPML -------------------------------------------------------------------------
template <typename T, int V, int AM, bool SLOW = false>
__device__ __forceinline__ void yee_h_lean(const Coefs<T>& c, const Geom& g,
                                           const Pack<T, V>& hx, const Pack<T, V>& hy, const Pack<T, V>& hz,
                                           const Pack<T, V>& ex, const Pack<T, V>& ey, const Pack<T, V>& ez,
                                           const Pack<T, V>& ez_jm, const Pack<T, V>& ex_jm, T ey_km, T ex_km,
                                           const Pack<T, V>& ey_im, const Pack<T, V>& ez_im,
                                           Pack<T, V>& ox, Pack<T, V>& oy, Pack<T, V>& oz)
{
    unsigned bad = 0;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const T ey_k = e > 0 ? ey.v[(e + V - 1) % V] : ey_km;
        const T ex_k = e > 0 ? ex.v[(e + V - 1) % V] : ex_km;
        if (SLOW) {
            ox.v[e] = upd_h<T>(c.uda, hx.v[e], c.udb, Ar<T>::diff_exact(ez.v[e], ez_jm.v[e], g.dy, g.rdy), Ar<T>::diff_exact(ey.v[e], ey_k, g.dz, g.rdz));
            oy.v[e] = upd_h<T>(c.uda, hy.v[e], c.udb, Ar<T>::diff_exact(ex.v[e], ex_k, g.dz, g.rdz), Ar<T>::diff_exact(ez.v[e], ez_im.v[e], g.dx, g.rdx));
            oz.v[e] = upd_h<T>(c.uda, hz.v[e], c.udb, Ar<T>::diff_exact(ey.v[e], ey_im.v[e], g.dx, g.rdx), Ar<T>::diff_exact(ex.v[e], ex_jm.v[e], g.dy, g.rdy));
        } else {
            ox.v[e] = upd_h<T>(c.uda, hx.v[e], c.udb, Ar<T>::diff_fast(ez.v[e], ez_jm.v[e], g.dy, g.rdy, bad), Ar<T>::diff_fast(ey.v[e], ey_k, g.dz, g.rdz, bad));
            oy.v[e] = upd_h<T>(c.uda, hy.v[e], c.udb, Ar<T>::diff_fast(ex.v[e], ex_k, g.dz, g.rdz, bad), Ar<T>::diff_fast(ez.v[e], ez_im.v[e], g.dx, g.rdx, bad));
            oz.v[e] = upd_h<T>(c.uda, hz.v[e], c.udb, Ar<T>::diff_fast(ey.v[e], ey_im.v[e], g.dx, g.rdx, bad), Ar<T>::diff_fast(ex.v[e], ex_jm.v[e], g.dy, g.rdy, bad));
        }
    }
    if (sizeof(T) == 8 && !SLOW) {
        if (bad) yee_h_lean<T, V, AM, true>(c, g, hx, hy, hz, ex, ey, ez, ez_jm, ex_jm, ey_km, ex_km, ey_im, ez_im, ox, oy, oz);
    }
}

template <typename T, int V, int AM, bool SLOW = false>
__device__ __forceinline__ void yee_e_lean(const Coefs<T>& c, const Geom& g,
                                           const Pack<T, V>& ex, const Pack<T, V>& ey, const Pack<T, V>& ez,
                                           const Pack<T, V>& hx, const Pack<T, V>& hy, const Pack<T, V>& hz,
                                           const Pack<T, V>& hz_jp, const Pack<T, V>& hx_jp, T hy_kp, T hx_kp,
                                           const Pack<T, V>& hy_ip, const Pack<T, V>& hz_ip,
                                           Pack<T, V>& ox, Pack<T, V>& oy, Pack<T, V>& oz)
{
    unsigned bad = 0;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const T hy_k = (e + 1 < V) ? hy.v[(e + 1) % V] : hy_kp;
        const T hx_k = (e + 1 < V) ? hx.v[(e + 1) % V] : hx_kp;
        if (SLOW) {
            ox.v[e] = upd_e<T>(c.uca, ex.v[e], c.ucb, Ar<T>::diff_exact(hz_jp.v[e], hz.v[e], g.dy, g.rdy), Ar<T>::diff_exact(hy_k, hy.v[e], g.dz, g.rdz));
            oy.v[e] = upd_e<T>(c.uca, ey.v[e], c.ucb, Ar<T>::diff_exact(hx_k, hx.v[e], g.dz, g.rdz), Ar<T>::diff_exact(hz_ip.v[e], hz.v[e], g.dx, g.rdx));
            oz.v[e] = upd_e<T>(c.uca, ez.v[e], c.ucb, Ar<T>::diff_exact(hy_ip.v[e], hy.v[e], g.dx, g.rdx), Ar<T>::diff_exact(hx_jp.v[e], hx.v[e], g.dy, g.rdy));
        } else {
            ox.v[e] = upd_e<T>(c.uca, ex.v[e], c.ucb, Ar<T>::diff_fast(hz_jp.v[e], hz.v[e], g.dy, g.rdy, bad), Ar<T>::diff_fast(hy_k, hy.v[e], g.dz, g.rdz, bad));
            oy.v[e] = upd_e<T>(c.uca, ey.v[e], c.ucb, Ar<T>::diff_fast(hx_k, hx.v[e], g.dz, g.rdz, bad), Ar<T>::diff_fast(hz_ip.v[e], hz.v[e], g.dx, g.rdx, bad));
            oz.v[e] = upd_e<T>(c.uca, ez.v[e], c.ucb, Ar<T>::diff_fast(hy_ip.v[e], hy.v[e], g.dx, g.rdx, bad), Ar<T>::diff_fast(hx_jp.v[e], hx.v[e], g.dy, g.rdy, bad));
        }
    }
    if (sizeof(T) == 8 && !SLOW) {
        if (bad) yee_e_lean<T, V, AM, true>(c, g, ex, ey, ez, hx, hy, hz, hz_jp, hx_jp, hy_kp, hx_kp, hy_ip, hz_ip, ox, oy, oz);
    }
}

// ---- non-lean bodies: update-range masks + CPML recursion on slab cells (fdtd_yee.cuh:63-89, :106-138) ---------------------------
// The first version of these bodies spent 726 SASS instructions per H stage, 156 of them floating point: masks, slab
// indices and psi addresses were re-derived per cell with nested branches (profiles/r02_ncu_yeex.md).  Now
//   * everything that depends only on the thread's (j, k) is computed ONCE per kernel (YeexThread: mask bits, slab
//     indices, 32-bit psi offsets), what depends on the plane once per iteration (uniform);
//   * the three psi families are COMPILE-TIME flags <X, Y, Z>: a warp in a z-slab tile of an interior row on an
//     interior plane instantiates only the four z recursions, and so on (8 variants per stage, chosen by a warp-
//     uniform switch);
//   * cells are updated branch-free: compute, then select by mask bit (a cell outside the update range keeps its field
//     AND its psi), z coefficients default to the identity (b = a = 0, 1/kappa = 1) outside the slab.
//
// The six psi arrays of a stage, in the order of Cpml::psi from `base` (0: E stage, 6: H stage):
//   base+0 y0 (x-component, d/dy)   base+1 z0 (x-component, d/dz)   base+2 z1 (y-component, d/dz)
//   base+3 x0 (y-component, d/dx)   base+4 x1 (z-component, d/dx)   base+5 y1 (z-component, d/dy)
template <typename T, int V> struct YeexPsi { Pack<T, V> x0, x1, y0, y1, z0, z1; };

// (Evict-first field stores, st.global.cs, were measured and make no difference: 88.0 vs 88.1 Gcell/s, profiles/r02_tuning.md.)
template <typename T, int V> __device__ __forceinline__ Pack<T, V> ldnc8(const T* p)
{
    typedef typename Vec8<T>::type VT;
    union { VT q; Pack<T, V> r; } u;
    u.q = __ldg(reinterpret_cast<const VT*>(p));
    return u.r;
}

template <int V> struct YeexThread {
    int j, k, syj;
    int szk[V];                 // z-slab index of each cell (-1: none)
    unsigned off_xy;            // j * sy + k: x family (+ sxp * x_sx) and, with syj in place of j, ...
    unsigned off_y;             // syj * sy + k: y family (+ plane * y_sx)
    unsigned off_z[V];          // j * z_pitch + szk: z family (+ plane * ny * z_pitch)
    unsigned bits;              // bit e: 1 <= k+e <= nz-2 (km); 8+e: k+e < nz-1 (kz1); 16+e: k+e < nz (kz0);
                                // 24: 1 <= j <= ny-2 (jm); 25: j < ny-1 (jy1); 26: j < ny (jy0)
    bool ok;                    // inside the arrays: psi may be loaded
};

template <int V>
__device__ __forceinline__ YeexThread<V> yeex_thread(const Geom& g, const SlabGeom& sg, int tpm, int j, int k, int syj, bool in_grid)
{
    YeexThread<V> th;
    th.j = j; th.k = k; th.syj = syj; th.ok = in_grid;
    th.off_xy = (unsigned)j * (unsigned)g.sy + (unsigned)max(k, 0);
    th.off_y = (unsigned)max(syj, 0) * (unsigned)g.sy + (unsigned)max(k, 0);
    unsigned b = 0;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const int ke = k + e;
        th.szk[e] = (tpm && ke >= 0 && ke < g.nz) ? slab_index(ke, g.nz, tpm) : -1;
        th.off_z[e] = (unsigned)j * (unsigned)sg.z_pitch + (unsigned)max(th.szk[e], 0);
        if (ke >= 1 && ke <= g.nz - 2) b |= 1u << e;
        if (ke >= 0 && ke < g.nz - 1) b |= 1u << (8 + e);
        if (ke >= 0 && ke < g.nz) b |= 1u << (16 + e);
    }
    if (j >= 1 && j <= g.ny - 2) b |= 1u << 24;
    if (j < g.ny - 1) b |= 1u << 25;
    if (j < g.ny) b |= 1u << 26;
    th.bits = b;
    return th;
}

// Issue the loads of the psi values the thread will use at plane pl (runtime family tests: a handful of predicated loads).
template <typename T, int V>
__device__ __forceinline__ void yeex_psi_load(YeexPsi<T, V>& r, const T* const* __restrict__ psi_in, const int base, const Geom& g,
                                              const SlabGeom& sg, const int tpm, const int pl, const YeexThread<V>& th)
{
    if (!th.ok || !tpm || pl >= g.nx) return;
    const int sxp = slab_index(pl, g.nx, tpm);
    if (sxp >= 0) {
        const unsigned o = (unsigned)sxp * (unsigned)sg.x_sx + th.off_xy;
        r.x0 = ldnc8<T, V>(psi_in[base + 3] + o); r.x1 = ldnc8<T, V>(psi_in[base + 4] + o);
    }
    if (th.syj >= 0) {
        const unsigned o = (unsigned)pl * (unsigned)sg.y_sx + th.off_y;
        r.y0 = ldnc8<T, V>(psi_in[base + 0] + o); r.y1 = ldnc8<T, V>(psi_in[base + 5] + o);
    }
    const unsigned zp = (unsigned)pl * (unsigned)(g.ny * sg.z_pitch);
#pragma unroll
    for (int e = 0; e < V; ++e)
        if (th.szk[e] >= 0) { r.z0.v[e] = __ldg(psi_in[base + 1] + zp + th.off_z[e]); r.z1.v[e] = __ldg(psi_in[base + 2] + zp + th.off_z[e]); }
}

// the same addresses, written to the OTHER psi set (psi is ping-ponged like the fields: rim threads and segment prologues
// recompute the recursion without storing)
template <typename T, int V, bool X, bool Y, bool Z>
__device__ __forceinline__ void yeex_psi_store(const YeexPsi<T, V>& r, T* const* __restrict__ psi_out, const int base, const Geom& g,
                                               const SlabGeom& sg, const int sxp, const int pl, const YeexThread<V>& th)
{
    if (X) {
        const unsigned o = (unsigned)sxp * (unsigned)sg.x_sx + th.off_xy;
        st8<T, V>(psi_out[base + 3] + o, r.x0); st8<T, V>(psi_out[base + 4] + o, r.x1);
    }
    if (Y) {
        const unsigned o = (unsigned)pl * (unsigned)sg.y_sx + th.off_y;
        st8<T, V>(psi_out[base + 0] + o, r.y0); st8<T, V>(psi_out[base + 5] + o, r.y1);
    }
    if (Z) {
        const unsigned zp = (unsigned)pl * (unsigned)(g.ny * sg.z_pitch);
#pragma unroll
        for (int e = 0; e < V; ++e)
            if (th.szk[e] >= 0) { psi_out[base + 1][zp + th.off_z[e]] = r.z0.v[e]; psi_out[base + 2][zp + th.off_z[e]] = r.z1.v[e]; }
    }
}

// CPML coefficients of one stage at this thread's position (v0 = 0: E-derivative positions, 3: H): b, a, 1/kappa
template <typename T, int V, bool X, bool Y, bool Z> struct YeexCoef { T bx, ax, kx, by, ay, ky; T bz[V], az[V], kz[V]; };

template <typename T, int V, bool X, bool Y, bool Z>
__device__ __forceinline__ void yeex_coef_load(YeexCoef<T, V, X, Y, Z>& q, const Cpml& pm, const int v0, const int pl,
                                               const YeexThread<V>& th)
{
    if (X) { q.bx = __ldg(cpml_tab<T>(pm.ax[0], v0) + pl); q.ax = __ldg(cpml_tab<T>(pm.ax[0], v0 + 1) + pl); q.kx = __ldg(cpml_tab<T>(pm.ax[0], v0 + 2) + pl); }
    if (Y) { q.by = __ldg(cpml_tab<T>(pm.ax[1], v0) + th.j); q.ay = __ldg(cpml_tab<T>(pm.ax[1], v0 + 1) + th.j); q.ky = __ldg(cpml_tab<T>(pm.ax[1], v0 + 2) + th.j); }
    if (Z) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
            q.bz[e] = q.az[e] = (T)0; q.kz[e] = (T)1;
            if (th.szk[e] >= 0) {
                const int ke = th.k + e;
                q.bz[e] = __ldg(cpml_tab<T>(pm.ax[2], v0) + ke); q.az[e] = __ldg(cpml_tab<T>(pm.ax[2], v0 + 1) + ke); q.kz[e] = __ldg(cpml_tab<T>(pm.ax[2], v0 + 2) + ke);
            }
        }
    }
}

// one derivative through one recursion: psi' = b psi + a d, d' = d / kappa + psi'; `on` = the cell is in the update range
template <typename T>
__device__ __forceinline__ T yeex_cpml(T d, T& psi, T b, T a, T ki, bool on)
{
    const T p = CpmlMath<T>::psi(b, psi, a, d);
    psi = on ? p : psi;
    return CpmlMath<T>::eff(ki, d, p);
}

template <typename T, int V, bool X, bool Y, bool Z>
__device__ __forceinline__ void yee_h_full(const Coefs<T>& c, const Geom& g, const Cpml& pm, YeexPsi<T, V>& r, const YeexThread<V>& th,
                                           const int p,
                                           const Pack<T, V>& hx, const Pack<T, V>& hy, const Pack<T, V>& hz,
                                           const Pack<T, V>& ex, const Pack<T, V>& ey, const Pack<T, V>& ez,
                                           const Pack<T, V>& ez_jm, const Pack<T, V>& ex_jm, T ey_km, T ex_km,
                                           const Pack<T, V>& ey_im, const Pack<T, V>& ez_im,
                                           Pack<T, V>& ox, Pack<T, V>& oy, Pack<T, V>& oz)
{
    YeexCoef<T, V, X, Y, Z> q;
    yeex_coef_load<T, V, X, Y, Z>(q, pm, 3, p, th);
    const bool px1 = p < g.nx - 1, pxm = p >= 1 && p <= g.nx - 2;
    const bool jm = (th.bits >> 24) & 1u, jy1 = (th.bits >> 25) & 1u;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const bool km = (th.bits >> e) & 1u, kz1 = (th.bits >> (8 + e)) & 1u;
        const T ey_k = e > 0 ? ey.v[(e + V - 1) % V] : ey_km;
        const T ex_k = e > 0 ? ex.v[(e + V - 1) % V] : ex_km;
        {                                                  // Hx(p+1/2, j, k): p < nx-1, 1 <= j <= ny-2, 1 <= k <= nz-2
            const bool on = px1 && jm && km;
            T dy = Ar<T>::diff(ez.v[e], ez_jm.v[e], g.dy, g.rdy), dz = Ar<T>::diff(ey.v[e], ey_k, g.dz, g.rdz);
            if (Y) dy = yeex_cpml<T>(dy, r.y0.v[e], q.by, q.ay, q.ky, on);
            if (Z) dz = yeex_cpml<T>(dz, r.z0.v[e], q.bz[e], q.az[e], q.kz[e], on);
            const T n = upd_h<T>(c.uda, hx.v[e], c.udb, dy, dz);
            ox.v[e] = on ? n : hx.v[e];
        }
        {                                                  // Hy(p, j+1/2, k): 1 <= p <= nx-2, j < ny-1, 1 <= k <= nz-2
            const bool on = pxm && jy1 && km;
            T dz = Ar<T>::diff(ex.v[e], ex_k, g.dz, g.rdz), dx = Ar<T>::diff(ez.v[e], ez_im.v[e], g.dx, g.rdx);
            if (Z) dz = yeex_cpml<T>(dz, r.z1.v[e], q.bz[e], q.az[e], q.kz[e], on);
            if (X) dx = yeex_cpml<T>(dx, r.x0.v[e], q.bx, q.ax, q.kx, on);
            const T n = upd_h<T>(c.uda, hy.v[e], c.udb, dz, dx);
            oy.v[e] = on ? n : hy.v[e];
        }
        {                                                  // Hz(p, j, k+1/2): 1 <= p <= nx-2, 1 <= j <= ny-2, k < nz-1
            const bool on = pxm && jm && kz1;
            T dx = Ar<T>::diff(ey.v[e], ey_im.v[e], g.dx, g.rdx), dy = Ar<T>::diff(ex.v[e], ex_jm.v[e], g.dy, g.rdy);
            if (X) dx = yeex_cpml<T>(dx, r.x1.v[e], q.bx, q.ax, q.kx, on);
            if (Y) dy = yeex_cpml<T>(dy, r.y1.v[e], q.by, q.ay, q.ky, on);
            const T n = upd_h<T>(c.uda, hz.v[e], c.udb, dx, dy);
            oz.v[e] = on ? n : hz.v[e];
        }
    }
}

template <typename T, int V, bool X, bool Y, bool Z>
__device__ __forceinline__ void yee_e_full(const Coefs<T>& c, const Geom& g, const Cpml& pm, YeexPsi<T, V>& r, const YeexThread<V>& th,
                                           const int i,
                                           const Pack<T, V>& ex, const Pack<T, V>& ey, const Pack<T, V>& ez,
                                           const Pack<T, V>& hx, const Pack<T, V>& hy, const Pack<T, V>& hz,
                                           const Pack<T, V>& hz_jp, const Pack<T, V>& hx_jp, T hy_kp, T hx_kp,
                                           const Pack<T, V>& hy_ip, const Pack<T, V>& hz_ip,
                                           Pack<T, V>& ox, Pack<T, V>& oy, Pack<T, V>& oz)
{
    YeexCoef<T, V, X, Y, Z> q;
    yeex_coef_load<T, V, X, Y, Z>(q, pm, 0, i, th);
    const bool qx1 = i < g.nx - 1;
    const bool jy1 = (th.bits >> 25) & 1u, jy0 = (th.bits >> 26) & 1u;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const bool kz1 = (th.bits >> (8 + e)) & 1u, kz0 = (th.bits >> (16 + e)) & 1u;
        const T hy_k = (e + 1 < V) ? hy.v[(e + 1) % V] : hy_kp;
        const T hx_k = (e + 1 < V) ? hx.v[(e + 1) % V] : hx_kp;
        {                                                  // Ex(i, j+1/2, k+1/2): j < ny-1, k < nz-1
            const bool on = jy1 && kz1;
            T dy = Ar<T>::diff(hz_jp.v[e], hz.v[e], g.dy, g.rdy), dz = Ar<T>::diff(hy_k, hy.v[e], g.dz, g.rdz);
            if (Y) dy = yeex_cpml<T>(dy, r.y0.v[e], q.by, q.ay, q.ky, on);
            if (Z) dz = yeex_cpml<T>(dz, r.z0.v[e], q.bz[e], q.az[e], q.kz[e], on);
            const T n = upd_e<T>(c.uca, ex.v[e], c.ucb, dy, dz);
            ox.v[e] = on ? n : ex.v[e];
        }
        {                                                  // Ey(i+1/2, j, k+1/2): i < nx-1, k < nz-1
            const bool on = qx1 && kz1 && jy0;
            T dz = Ar<T>::diff(hx_k, hx.v[e], g.dz, g.rdz), dx = Ar<T>::diff(hz_ip.v[e], hz.v[e], g.dx, g.rdx);
            if (Z) dz = yeex_cpml<T>(dz, r.z1.v[e], q.bz[e], q.az[e], q.kz[e], on);
            if (X) dx = yeex_cpml<T>(dx, r.x0.v[e], q.bx, q.ax, q.kx, on);
            const T n = upd_e<T>(c.uca, ey.v[e], c.ucb, dz, dx);
            oy.v[e] = on ? n : ey.v[e];
        }
        {                                                  // Ez(i+1/2, j+1/2, k): i < nx-1, j < ny-1
            const bool on = qx1 && jy1 && kz0;
            T dx = Ar<T>::diff(hy_ip.v[e], hy.v[e], g.dx, g.rdx), dy = Ar<T>::diff(hx_jp.v[e], hx.v[e], g.dy, g.rdy);
            if (X) dx = yeex_cpml<T>(dx, r.x1.v[e], q.bx, q.ax, q.kx, on);
            if (Y) dy = yeex_cpml<T>(dy, r.y1.v[e], q.by, q.ay, q.ky, on);
            const T n = upd_e<T>(c.uca, ez.v[e], c.ucb, dx, dy);
            oz.v[e] = on ? n : ez.v[e];
        }
    }
}

// warp-uniform dispatch of a body templated on the psi families that are active: cls = X | Y << 1 | Z << 2
#define YEEX_DISPATCH(cls, CALL)                                                                  \
    switch (cls) {                                                                                \
    case 0: { constexpr bool X = false, Y = false, Z = false; CALL; } break;                      \
    case 1: { constexpr bool X = true, Y = false, Z = false; CALL; } break;                       \
    case 2: { constexpr bool X = false, Y = true, Z = false; CALL; } break;                       \
    case 3: { constexpr bool X = true, Y = true, Z = false; CALL; } break;                        \
    case 4: { constexpr bool X = false, Y = false, Z = true; CALL; } break;                       \
    case 5: { constexpr bool X = true, Y = false, Z = true; CALL; } break;                        \
    case 6: { constexpr bool X = false, Y = true, Z = true; CALL; } break;                        \
    default: { constexpr bool X = true, Y = true, Z = true; CALL; } break;                        \
    }

template <typename T, int R, int AM>
__global__ void __launch_bounds__(32 * (R + 1), 1)
k_fused3d_yeex(const __grid_constant__ Tb2xMaps maps, const __grid_constant__ Fields<T> out,
               const __grid_constant__ Coefs<T> c, const __grid_constant__ Geom g, const __grid_constant__ FusedTiling t,
               const __grid_constant__ Cpml pm, const __grid_constant__ PsiOut pout, const __grid_constant__ SlabGeom sg,
               const int S, const int D)
{
    constexpr int V = Vec8<T>::V;
    typedef Pack<T, V> P;
    constexpr uint32_t ROWS = (uint32_t)yeex_tile_bytes<R>();      // one staged array tile (128-byte multiple: TMA destination)
    constexpr uint32_t XROW = 2 * kTb2xRowBytes;           // one row of one exchange slot: Hz+, Hx+
    extern __shared__ __align__(128) unsigned char smem_raw_[];
    const uint32_t tiles = (smem_u32(smem_raw_) + 127u) & ~127u;
    const uint32_t xch = tiles + (uint32_t)S * (uint32_t)yeex_stage_bytes<R>();
    const uint32_t full = xch + (uint32_t)D * R * XROW;
    const uint32_t empty = full + (uint32_t)S * 8u;
    const uint32_t xfull = empty + (uint32_t)S * 8u;

    const int lane = threadIdx.x, row = threadIdx.y;
    const int ntiles = t.ntj * t.ntk;
    const int seg = blockIdx.x / ntiles, tile = blockIdx.x - seg * ntiles;
    const int tj = tile / t.ntk, tk = tile - tj * t.ntk;
    const int j0 = tj * (R - 1), k0 = (tk * t.own_lanes - 1) * V;      // row 0 / lane 0 of the CTA (lane 0 is the -k halo lane)
    const int jbox = j0 - 1, kbox = k0 - V;                // box origin: one halo row; 16-byte aligned (own_lanes is even)
    const int i0 = t.i_begin + seg * t.lx;
    const int i1 = min(i0 + t.lx, t.i_end);
    const int n_it = i1 - i0 + 1;                          // i = i0 - 1 .. i1 - 1

    if (row == R && lane == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full + 8u * s, 1); mbar_init(empty + 8u * s, R); }
        for (int q = 0; q < R * D; ++q) mbar_init(xfull + 8u * q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (row == R) {
        // ---- producer: stage q <- E[i0 - 1 + q] (x,y,z) and, for q > 0, H[i0 - 1 + q] (x,y,z) --------------------------------
        if (lane == 0) {
            int s = 0, ph = 0;
            for (int q = 0; q <= n_it; ++q) {
                mbar_wait_parked(empty + 8u * s, (uint32_t)(ph ^ 1), 2000u);
                const uint32_t dst = tiles + (uint32_t)s * (uint32_t)yeex_stage_bytes<R>();
                const uint32_t bar = full + 8u * s;
                mbar_arrive_expect_tx(bar, (uint32_t)((q > 0 ? 6 : 3) * (R + 1) * kYeexBoxBytes));
#pragma unroll
                for (int a = 0; a < 3; ++a) tma_load_3d(dst + a * ROWS, &maps.m[a], kbox, jbox, i0 - 1 + q, bar);
                if (q > 0) {
#pragma unroll
                    for (int a = 3; a < 6; ++a) tma_load_3d(dst + a * ROWS, &maps.m[a], kbox, jbox, i0 - 1 + q, bar);
                }
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
        return;
    }

    // ---- consumers ---------------------------------------------------------------------------------------------------------------------
    const int j = j0 + row, k = k0 + lane * V;
    const int rowp = min(row + 1, R - 1);
    const uint32_t own_off = (uint32_t)((row + 1) * kYeexBoxBytes + (lane + 1) * 8);
    const uint32_t dn_off = (uint32_t)(row * kYeexBoxBytes + (lane + 1) * 8);
    const bool in_grid = j < g.ny && k >= 0 && k < g.pz;
    const bool owner = in_grid && row <= R - 2 && lane >= 1 && lane <= t.own_lanes;
    const unsigned ofs = (unsigned)j * (unsigned)g.sy + (unsigned)max(k, 0);
    const int tpm = pm.t;
    const int syj = (tpm && j < g.ny) ? slab_index(j, g.ny, tpm) : -1;
    // lean test: every CONSUMED cell of the tile (rows 0..R-1, lanes 1..31 for H+; one less on the high side for E+) is
    // strictly inside the grid faces (1 <= j <= ny-2, 1 <= k <= nz-2) and outside the y / z slabs
    const int jlo = j0, jhi = j0 + R - 1, klo = k0 + V, khi = k0 + 32 * V - 1;
    const bool tile_in = jlo >= 1 && jhi <= g.ny - 2 && klo >= 1 && khi <= g.nz - 2;
    const bool tile_z = tpm && (klo < tpm || khi >= g.nz - tpm - 1);
    const bool lean_thread = tile_in && !tile_z && syj < 0;
    // planes on which the H stage (plane p) / the E stage (plane i) of a lean thread need neither masks nor x psi
    const int hl_lo = max(1, tpm), el_lo = tpm, xl_hi = g.nx - 2 - tpm;
    const T* const* psi_in = reinterpret_cast<const T* const*>(pm.psi);
    T* const* psi_out = reinterpret_cast<T* const*>(pout.p);
    const YeexThread<V> th = yeex_thread<V>(g, sg, tpm, j, k, syj, in_grid);
    // psi families of this warp that do not depend on the plane: y (row inside a y slab), z (tile touches a z slab)
    const int cls_yz = (syj >= 0 ? 2 : 0) | (tile_z ? 4 : 0);

    P z_;
#pragma unroll
    for (int e = 0; e < V; ++e) z_.v[e] = (T)0;
    {   // publish number 0: H+[i0 - 1] is not needed (E+[i0 - 1] is never produced)
        const uint32_t xme = xch + (uint32_t)row * XROW + (uint32_t)(lane * 8);
        sts8<T, V>(xme, z_); sts8<T, V>(xme + kTb2xRowBytes, z_);
        __syncwarp();
        if (lane == 0) mbar_arrive(xfull + (uint32_t)((row * D) * 8));
    }
    // psi of the first H stage (plane i0); the E stage's first set (plane i0) is requested at the end of iteration 0
    YeexPsi<T, V> hps, eps;
    hps.x0 = hps.x1 = hps.y0 = hps.y1 = hps.z0 = hps.z1 = z_;
    eps = hps;
    if (!(lean_thread && i0 >= hl_lo && i0 <= xl_hi))
        yeex_psi_load<T, V>(hps, psi_in, 6, g, sg, tpm, i0, th);
    // window at it = 0 (i = i0 - 1): E[i] = stage 0 (Ey, Ez feed the x-differences of H+[i0]); H+[i] unused
    mbar_wait(full, 0);
    P eAx = lds8<T, V>(tiles + 0 * ROWS + own_off), eAy = lds8<T, V>(tiles + 1 * ROWS + own_off), eAz = lds8<T, V>(tiles + 2 * ROWS + own_off);
    __syncwarp();
    if (lane == 0) mbar_arrive(empty);
    P hAx = z_, hAy = z_, hAz = z_;
    P eBx = z_, eBy = z_, eBz = z_, hBx = z_, hBy = z_, hBz = z_;
    int sq = 1 % S, sph = (1 / S) & 1, xd = 0, xph = 0;

    // one plane: (E[i], H+[i]) in, (E[p], H+[p]) out; the caller alternates the two register windows instead of moving them
    auto iter = [&](const int it, const P& e0x, const P& e0y, const P& e0z, const P& hpx, const P& hpy, const P& hpz,
                    P& e1x, P& e1y, P& e1z, P& hnx, P& hny, P& hnz) {
        const int i = i0 - 1 + it, p = i + 1;
        // ---- stage it + 1: E[p] (own, -j rows), H[p] own; hand the stage back as soon as it is in registers ------------------------
        const uint32_t st = tiles + (uint32_t)sq * (uint32_t)yeex_stage_bytes<R>();
        mbar_wait(full + (uint32_t)sq * 8u, (uint32_t)sph);
        e1x = lds8<T, V>(st + 0 * ROWS + own_off); e1y = lds8<T, V>(st + 1 * ROWS + own_off); e1z = lds8<T, V>(st + 2 * ROWS + own_off);
        const P h1x = lds8<T, V>(st + 3 * ROWS + own_off), h1y = lds8<T, V>(st + 4 * ROWS + own_off), h1z = lds8<T, V>(st + 5 * ROWS + own_off);
        const P ez_jm = lds8<T, V>(st + 2 * ROWS + dn_off), ex_jm = lds8<T, V>(st + 0 * ROWS + dn_off);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + (uint32_t)sq * 8u);
        if (++sq == S) { sq = 0; sph ^= 1; }
        // ---- +j neighbours of H+[i] from the row above (slot xd) -----------------------------------------------------------------------
        mbar_wait(xfull + (uint32_t)((rowp * D + xd) * 8), (uint32_t)xph);
        const uint32_t xup = xch + (uint32_t)xd * (R * XROW) + (uint32_t)rowp * XROW + (uint32_t)(lane * 8);
        const P hz_jp = lds8<T, V>(xup), hx_jp = lds8<T, V>(xup + kTb2xRowBytes);
        // ---- -k / +k neighbours from the neighbouring lanes ----------------------------------------------------------------------------
        const T ey_km = shfl_prev<T>(e1y.v[V - 1]), ex_km = shfl_prev<T>(e1x.v[V - 1]);
        const T hy_kp = shfl_next<T>(hpy.v[0]), hx_kp = shfl_next<T>(hpx.v[0]);

        // ---- H+[p] ------------------------------------------------------------------------------------------------------------------------
        const bool more = it + 1 < n_it;
        const bool st_h = owner && p < i1;
        if (lean_thread && p >= hl_lo && p <= xl_hi)
            yee_h_lean<T, V, AM>(c, g, h1x, h1y, h1z, e1x, e1y, e1z, ez_jm, ex_jm, ey_km, ex_km, e0y, e0z, hnx, hny, hnz);
        else {
            const int sxp = (tpm && p < g.nx) ? slab_index(p, g.nx, tpm) : -1;
            YEEX_DISPATCH(cls_yz | (sxp >= 0 ? 1 : 0),
                          (yee_h_full<T, V, X, Y, Z>(c, g, pm, hps, th, p, h1x, h1y, h1z, e1x, e1y, e1z, ez_jm, ex_jm, ey_km, ex_km,
                                                     e0y, e0z, hnx, hny, hnz),
                           st_h ? yeex_psi_store<T, V, X, Y, Z>(hps, psi_out, 6, g, sg, sxp, p, th) : (void)0))
        }
        // psi of the next H stage (plane p + 1): the registers are free now, the loads have a whole iteration to land
        if (more && !(lean_thread && p + 1 >= hl_lo && p + 1 <= xl_hi))
            yeex_psi_load<T, V>(hps, psi_in, 6, g, sg, tpm, p + 1, th);
        // ---- publish Hz+, Hx+ of plane p for the row below (its E+ of the next iteration) ----------------------------------------------
        int xd1 = xd + 1, xph1 = xph;
        if (xd1 == D) { xd1 = 0; xph1 ^= 1; }
        if (more) {
            if (row > 0 && it + 1 >= D) {
                int bd = xd1 + 1, bph = xph1 ^ 1;
                if (bd == D) { bd = 0; bph ^= 1; }
                mbar_wait(xfull + (uint32_t)(((row - 1) * D + bd) * 8), (uint32_t)bph);
            }
            const uint32_t xme = xch + (uint32_t)xd1 * (R * XROW) + (uint32_t)row * XROW + (uint32_t)(lane * 8);
            sts8<T, V>(xme, hnz); sts8<T, V>(xme + kTb2xRowBytes, hnx);
            __syncwarp();
            if (lane == 0) mbar_arrive(xfull + (uint32_t)((row * D + xd1) * 8));
        }
        xd = xd1; xph = xph1;
        if (st_h) {
            const unsigned oh = ofs + (unsigned)p * (unsigned)g.sx;
            st8<T, V>(out.hx + oh, hnx); st8<T, V>(out.hy + oh, hny); st8<T, V>(out.hz + oh, hnz);
        }
        // ---- E+[i]: owners only (nobody consumes the E+ of a rim thread) -------------------------------------------------------------------
        if (it > 0 && row <= R - 2) {                         // (the 15th row only provides H+: it has no E stage)
            P nx_ = e0x, ny_ = e0y, nz_ = e0z;
            if (lean_thread && i >= el_lo && i <= xl_hi)
                yee_e_lean<T, V, AM>(c, g, e0x, e0y, e0z, hpx, hpy, hpz, hz_jp, hx_jp, hy_kp, hx_kp, hny, hnz, nx_, ny_, nz_);
            else if (owner) {
                const int sxq = tpm ? slab_index(i, g.nx, tpm) : -1;
                YEEX_DISPATCH(cls_yz | (sxq >= 0 ? 1 : 0),
                              (yee_e_full<T, V, X, Y, Z>(c, g, pm, eps, th, i, e0x, e0y, e0z, hpx, hpy, hpz, hz_jp, hx_jp, hy_kp, hx_kp,
                                                         hny, hnz, nx_, ny_, nz_),
                               yeex_psi_store<T, V, X, Y, Z>(eps, psi_out, 0, g, sg, sxq, i, th)))
            }
            if (owner) {
                const unsigned oe = ofs + (unsigned)i * (unsigned)g.sx;
                st8<T, V>(out.ex + oe, nx_); st8<T, V>(out.ey + oe, ny_); st8<T, V>(out.ez + oe, nz_);
            }
        }
        // psi of the next E stage (plane i + 1 = p)
        if (more && owner && !(lean_thread && p >= el_lo && p <= xl_hi))
            yeex_psi_load<T, V>(eps, psi_in, 0, g, sg, tpm, p, th);
    };

    int it = 0;
    for (; it + 2 <= n_it; it += 2) {
        iter(it, eAx, eAy, eAz, hAx, hAy, hAz, eBx, eBy, eBz, hBx, hBy, hBz);
        iter(it + 1, eBx, eBy, eBz, hBx, hBy, hBz, eAx, eAy, eAz, hAx, hAy, hAz);
    }
    if (it < n_it) iter(it, eAx, eAy, eAz, hAx, hAy, hAz, eBx, eBy, eBz, hBx, hBy, hBz);
}

}  // namespace fdtd
