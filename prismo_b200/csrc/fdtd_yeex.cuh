// fdtd_yeex.cuh — physics mode (stable Yee leap-frog + CPML) as ONE TMA-fed sweep per step: the north-star kernel
// ("fused 3-D E/H + CPML step").  Same arithmetic, update ranges and psi layout as the two-pass pair in fdtd_yee.cuh
// (which oracle/yee.py pins) and as the first fused version in fdtd_yee_fused.cuh; what changes is the plumbing, which
// is that of fdtd_tb2x.cuh:
//
//   * the six input arrays of plane p arrive by cp.async.bulk.tensor (one producer thread, S-stage ring, full / empty
//     mbarriers).  The tile carries its own halo — row 0 is j0 - 1, lane 0 is k0 - 1, zero-filled outside the grid —
//     so the -j / -k neighbours of E that the backward-differenced H update needs are plain shared-memory reads;
//   * the +j neighbours of H+ (z, x components) that the E update needs travel through a D-slot ring with one mbarrier
//     per row: neighbour-only synchronisation, no CTA-wide barrier in the plane loop;
//   * CPML is applied by slab threads only, and decided per WARP and PLANE: a warp whose row lies outside the y slabs,
//     in a tile outside the z slabs and clear of the grid faces, on planes outside the x slabs, runs a lean body with
//     no masks and no psi code at all (about 85 % of the warp-iterations of a 1024^3 grid with 10-cell layers); the
//     others run the full body (masks + psi recursion), whose psi arrays are ping-ponged like the fields.
//
//   H+[p] = f(H[p], E[p] (own, j-1, k-1), E[p-1])          E+[q] = g(E[q], H+[q] (own, j+1, k+1), H+[q+1])
// Iteration `it` (i = i0 - 1 + it) reads TMA stage it + 1 (E[i+1], H[i+1]), produces H+[i+1] and E+[i].  Of R = 15 rows
// 13 own cells (row 0 and row 14 are halo providers), of 32 lanes 30 (lane 0 and lane 31).
//
// PARITY UNPINNED (there are no reference numbers for a working CPML): validated bitwise in fp64 against the two-pass
// kernels and through them against oracle/yee.py.
#pragma once
#include "fdtd_tb2x.cuh"
#include "fdtd_yee_fused.cuh"

namespace fdtd {

constexpr int kYeexRows = kTb2xRows;       // 15 consumer rows + 1 producer warp, as in the two-step sweep
// TMA box: 272 B x R rows.  The first cell a tile CONSUMES is one lane (8 B) left of its first owner lane, i.e. at byte
// 240 * tk - 8 of the row, but the box origin of cp.async.bulk.tensor must be 16-byte aligned in the innermost
// dimension (a misaligned origin raises "illegal instruction": measured, gpurun_out/e2_sanitize_memcheck_yee.log), so the
// box starts one more lane to the left and lane l reads its cells at byte (l + 1) * 8 of the staged row.
constexpr int kYeexBoxBytes = 272;
template <int R> __host__ __device__ constexpr size_t yeex_tile_bytes() { return ((size_t)R * kYeexBoxBytes + 127) / 128 * 128; }
template <int R> __host__ __device__ constexpr size_t yeex_stage_bytes() { return 6 * yeex_tile_bytes<R>(); }

template <int R> static inline size_t yeex_smem_bytes(int S, int D)
{
    return 128 + (size_t)S * yeex_stage_bytes<R>() + (size_t)D * R * 2 * kTb2xRowBytes + (size_t)(2 * S + R * D) * 8 + 64;
}

template <typename T> struct YeexCtx {
    int j, k, row, lane;
    bool owner, interior_tile, tile_z, row_y;
    int syj;                         // y-slab index of this row (-1 outside)
    unsigned ofs;                    // j * sy + k (owner threads only use it)
};

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

// ---- full bodies: update-range masks + CPML recursion on slab cells (fdtd_yee.cuh:63-89, :106-138) --------------------------------
template <typename T, int V>
__device__ __forceinline__ void yee_h_full(const Coefs<T>& c, const Geom& g, const Cpml& pm, const SlabGeom& sg,
                                           const T* const* psi_in, T* const* psi_out, bool store, int p, int j, int k, int syj,
                                           const Pack<T, V>& hx, const Pack<T, V>& hy, const Pack<T, V>& hz,
                                           const Pack<T, V>& ex, const Pack<T, V>& ey, const Pack<T, V>& ez,
                                           const Pack<T, V>& ez_jm, const Pack<T, V>& ex_jm, T ey_km, T ex_km,
                                           const Pack<T, V>& ey_im, const Pack<T, V>& ez_im,
                                           Pack<T, V>& ox, Pack<T, V>& oy, Pack<T, V>& oz)
{
    const int tpm = pm.t;
    const int sxp = tpm ? slab_index(p, g.nx, tpm) : -1;
    const bool px1 = p < g.nx - 1, pxm = p >= 1 && p <= g.nx - 2;
    const bool jm = j >= 1 && j <= g.ny - 2, jy1 = j >= 0 && j < g.ny - 1;
    const long long o = (long long)j * g.sy + k;
    ox = hx; oy = hy; oz = hz;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const int ke = k + e;
        const bool km = ke >= 1 && ke <= g.nz - 2, kz1 = ke >= 0 && ke < g.nz - 1;
        const int szk = (tpm && ke >= 0) ? slab_index(ke, g.nz, tpm) : -1;
        const long long oxs = (long long)sxp * sg.x_sx + o + e;
        const long long oys = (long long)p * sg.y_sx + (long long)syj * g.sy + ke;
        const long long ozs = ((long long)p * g.ny + j) * sg.z_pitch + szk;
        const T ey_k = e > 0 ? ey.v[(e + V - 1) % V] : ey_km;
        const T ex_k = e > 0 ? ex.v[(e + V - 1) % V] : ex_km;
        if (px1 && jm && km) {                             // Hx(p+1/2, j, k)
            T dy = Ar<T>::diff(ez.v[e], ez_jm.v[e], g.dy, g.rdy), dz = Ar<T>::diff(ey.v[e], ey_k, g.dz, g.rdz);
            if (syj >= 0) dy = cpml_step<T>(dy, psi_in[6], psi_out[6], store, oys, pm.ax[1].c[3], pm.ax[1].c[4], pm.ax[1].c[5], j);
            if (szk >= 0) dz = cpml_step<T>(dz, psi_in[7], psi_out[7], store, ozs, pm.ax[2].c[3], pm.ax[2].c[4], pm.ax[2].c[5], ke);
            ox.v[e] = upd_h<T>(c.uda, hx.v[e], c.udb, dy, dz);
        }
        if (pxm && jy1 && km) {                            // Hy(p, j+1/2, k)
            T dz = Ar<T>::diff(ex.v[e], ex_k, g.dz, g.rdz), dx = Ar<T>::diff(ez.v[e], ez_im.v[e], g.dx, g.rdx);
            if (szk >= 0) dz = cpml_step<T>(dz, psi_in[8], psi_out[8], store, ozs, pm.ax[2].c[3], pm.ax[2].c[4], pm.ax[2].c[5], ke);
            if (sxp >= 0) dx = cpml_step<T>(dx, psi_in[9], psi_out[9], store, oxs, pm.ax[0].c[3], pm.ax[0].c[4], pm.ax[0].c[5], p);
            oy.v[e] = upd_h<T>(c.uda, hy.v[e], c.udb, dz, dx);
        }
        if (pxm && jm && kz1) {                            // Hz(p, j, k+1/2)
            T dx = Ar<T>::diff(ey.v[e], ey_im.v[e], g.dx, g.rdx), dy = Ar<T>::diff(ex.v[e], ex_jm.v[e], g.dy, g.rdy);
            if (sxp >= 0) dx = cpml_step<T>(dx, psi_in[10], psi_out[10], store, oxs, pm.ax[0].c[3], pm.ax[0].c[4], pm.ax[0].c[5], p);
            if (syj >= 0) dy = cpml_step<T>(dy, psi_in[11], psi_out[11], store, oys, pm.ax[1].c[3], pm.ax[1].c[4], pm.ax[1].c[5], j);
            oz.v[e] = upd_h<T>(c.uda, hz.v[e], c.udb, dx, dy);
        }
    }
}

template <typename T, int V>
__device__ __forceinline__ void yee_e_full(const Coefs<T>& c, const Geom& g, const Cpml& pm, const SlabGeom& sg,
                                           const T* const* psi_in, T* const* psi_out, bool store, bool can, int i, int j, int k, int syj,
                                           const Pack<T, V>& ex, const Pack<T, V>& ey, const Pack<T, V>& ez,
                                           const Pack<T, V>& hx, const Pack<T, V>& hy, const Pack<T, V>& hz,
                                           const Pack<T, V>& hz_jp, const Pack<T, V>& hx_jp, T hy_kp, T hx_kp,
                                           const Pack<T, V>& hy_ip, const Pack<T, V>& hz_ip,
                                           Pack<T, V>& ox, Pack<T, V>& oy, Pack<T, V>& oz)
{
    const int tpm = pm.t;
    const int sxq = tpm ? slab_index(i, g.nx, tpm) : -1;
    const bool qx1 = i < g.nx - 1, jy1 = j >= 0 && j < g.ny - 1, jy0 = j >= 0 && j < g.ny;
    const long long o = (long long)j * g.sy + k;
    ox = ex; oy = ey; oz = ez;
    if (!can) return;                                      // last row / lane of the tile: no +j / +k neighbour at hand
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const int ke = k + e;
        const bool kz0 = ke >= 0 && ke < g.nz, kz1 = ke >= 0 && ke < g.nz - 1;
        const int szk = (tpm && ke >= 0) ? slab_index(ke, g.nz, tpm) : -1;
        const long long oxs = (long long)sxq * sg.x_sx + o + e;
        const long long oys = (long long)i * sg.y_sx + (long long)syj * g.sy + ke;
        const long long ozs = ((long long)i * g.ny + j) * sg.z_pitch + szk;
        const T hy_k = (e + 1 < V) ? hy.v[(e + 1) % V] : hy_kp;
        const T hx_k = (e + 1 < V) ? hx.v[(e + 1) % V] : hx_kp;
        if (jy1 && kz1) {                                  // Ex(i, j+1/2, k+1/2)
            T dy = Ar<T>::diff(hz_jp.v[e], hz.v[e], g.dy, g.rdy), dz = Ar<T>::diff(hy_k, hy.v[e], g.dz, g.rdz);
            if (syj >= 0) dy = cpml_step<T>(dy, psi_in[0], psi_out[0], store, oys, pm.ax[1].c[0], pm.ax[1].c[1], pm.ax[1].c[2], j);
            if (szk >= 0) dz = cpml_step<T>(dz, psi_in[1], psi_out[1], store, ozs, pm.ax[2].c[0], pm.ax[2].c[1], pm.ax[2].c[2], ke);
            ox.v[e] = upd_e<T>(c.uca, ex.v[e], c.ucb, dy, dz);
        }
        if (qx1 && kz1 && jy0) {                           // Ey(i+1/2, j, k+1/2)
            T dz = Ar<T>::diff(hx_k, hx.v[e], g.dz, g.rdz), dx = Ar<T>::diff(hz_ip.v[e], hz.v[e], g.dx, g.rdx);
            if (szk >= 0) dz = cpml_step<T>(dz, psi_in[2], psi_out[2], store, ozs, pm.ax[2].c[0], pm.ax[2].c[1], pm.ax[2].c[2], ke);
            if (sxq >= 0) dx = cpml_step<T>(dx, psi_in[3], psi_out[3], store, oxs, pm.ax[0].c[0], pm.ax[0].c[1], pm.ax[0].c[2], i);
            oy.v[e] = upd_e<T>(c.uca, ey.v[e], c.ucb, dz, dx);
        }
        if (qx1 && jy1 && kz0) {                           // Ez(i+1/2, j+1/2, k)
            T dx = Ar<T>::diff(hy_ip.v[e], hy.v[e], g.dx, g.rdx), dy = Ar<T>::diff(hx_jp.v[e], hx.v[e], g.dy, g.rdy);
            if (sxq >= 0) dx = cpml_step<T>(dx, psi_in[4], psi_out[4], store, oxs, pm.ax[0].c[0], pm.ax[0].c[1], pm.ax[0].c[2], i);
            if (syj >= 0) dy = cpml_step<T>(dy, psi_in[5], psi_out[5], store, oys, pm.ax[1].c[0], pm.ax[1].c[1], pm.ax[1].c[2], j);
            oz.v[e] = upd_e<T>(c.uca, ez.v[e], c.ucb, dx, dy);
        }
    }
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
    const int j0 = tj * (R - 2) - 1, k0 = (tk * t.own_lanes - 1) * V;      // tile origin: one halo row / lane on the low side
    const int kbox = k0 - V;                               // box origin: 16-byte aligned (own_lanes is even)
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
                mbar_wait(empty + 8u * s, (uint32_t)(ph ^ 1));
                const uint32_t dst = tiles + (uint32_t)s * (uint32_t)yeex_stage_bytes<R>();
                const uint32_t bar = full + 8u * s;
                mbar_arrive_expect_tx(bar, (uint32_t)((q > 0 ? 6 : 3) * R * kYeexBoxBytes));
#pragma unroll
                for (int a = 0; a < 3; ++a) tma_load_3d(dst + a * ROWS, &maps.m[a], kbox, j0, i0 - 1 + q, bar);
                if (q > 0) {
#pragma unroll
                    for (int a = 3; a < 6; ++a) tma_load_3d(dst + a * ROWS, &maps.m[a], kbox, j0, i0 - 1 + q, bar);
                }
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
        return;
    }

    // ---- consumers ---------------------------------------------------------------------------------------------------------------------
    const int j = j0 + row, k = k0 + lane * V;
    const int rowm = max(row - 1, 0), rowp = min(row + 1, R - 1);
    const uint32_t own_off = (uint32_t)(row * kYeexBoxBytes + (lane + 1) * 8);
    const uint32_t dn_off = (uint32_t)(rowm * kYeexBoxBytes + (lane + 1) * 8);
    const bool in_grid = j >= 0 && j < g.ny && k >= 0 && k < g.pz;
    const bool owner = in_grid && row >= 1 && row <= R - 2 && lane >= 1 && lane <= t.own_lanes;
    const bool can_e = row <= R - 2 && lane <= 30;         // has its +j row and +k lane inside the tile
    const unsigned ofs = (unsigned)max(j, 0) * (unsigned)g.sy + (unsigned)max(k, 0);
    const int tpm = pm.t;
    const int syj = (tpm && j >= 0 && j < g.ny) ? slab_index(j, g.ny, tpm) : -1;
    // lean test: every CONSUMED cell of the tile (rows 1..R-1, lanes 1..31 for H+; one less for E+) is strictly inside the
    // grid faces (1 <= j <= ny-2, 1 <= k <= nz-2) and outside the y / z slabs
    const int jlo = j0 + 1, jhi = j0 + R - 1, klo = k0 + V, khi = k0 + 32 * V - 1;
    const bool tile_in = jlo >= 1 && jhi <= g.ny - 2 && klo >= 1 && khi <= g.nz - 2;
    const bool tile_z = tpm && (klo < tpm || khi >= g.nz - tpm - 1);
    const bool lean_thread = tile_in && !tile_z && syj < 0;
    const T* const* psi_in = reinterpret_cast<const T* const*>(pm.psi);
    T* const* psi_out = reinterpret_cast<T* const*>(pout.p);

    P z_;
#pragma unroll
    for (int e = 0; e < V; ++e) z_.v[e] = (T)0;
    {   // publish number 0: H+[i0 - 1] is not needed (E+[i0 - 1] is never produced)
        const uint32_t xme = xch + (uint32_t)row * XROW + (uint32_t)(lane * 8);
        sts8<T, V>(xme, z_); sts8<T, V>(xme + kTb2xRowBytes, z_);
        __syncwarp();
        if (lane == 0) mbar_arrive(xfull + (uint32_t)((row * D) * 8));
    }
    // window at it = 0 (i = i0 - 1): E[i] = stage 0 (Ey, Ez feed the x-differences of H+[i0]); H+[i] unused
    mbar_wait(full, 0);
    P e0x = lds8<T, V>(tiles + 0 * ROWS + own_off), e0y = lds8<T, V>(tiles + 1 * ROWS + own_off), e0z = lds8<T, V>(tiles + 2 * ROWS + own_off);
    __syncwarp();
    if (lane == 0) mbar_arrive(empty);
    P hpx = z_, hpy = z_, hpz = z_;
    int sq = 1 % S, sph = (1 / S) & 1, xd = 0, xph = 0;

    for (int it = 0; it < n_it; ++it) {
        const int i = i0 - 1 + it, p = i + 1;
        // ---- stage it + 1: E[p] (own, -j rows), H[p] own; hand the stage back as soon as it is in registers ------------------------
        const uint32_t st = tiles + (uint32_t)sq * (uint32_t)yeex_stage_bytes<R>();
        mbar_wait(full + (uint32_t)sq * 8u, (uint32_t)sph);
        const P e1x = lds8<T, V>(st + 0 * ROWS + own_off), e1y = lds8<T, V>(st + 1 * ROWS + own_off), e1z = lds8<T, V>(st + 2 * ROWS + own_off);
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
        const bool px_lean = p >= 1 && p <= g.nx - 2 && !(tpm && (p < tpm || p >= g.nx - tpm - 1));
        const bool st_h = owner && p < i1;
        P hnx, hny, hnz;
        if (lean_thread && px_lean)
            yee_h_lean<T, V, AM>(c, g, h1x, h1y, h1z, e1x, e1y, e1z, ez_jm, ex_jm, ey_km, ex_km, e0y, e0z, hnx, hny, hnz);
        else
            yee_h_full<T, V>(c, g, pm, sg, psi_in, psi_out, st_h, p, j, k, syj, h1x, h1y, h1z, e1x, e1y, e1z, ez_jm, ex_jm,
                             ey_km, ex_km, e0y, e0z, hnx, hny, hnz);
        // ---- publish Hz+, Hx+ of plane p for the row below (its E+ of the next iteration) ----------------------------------------------
        int xd1 = xd + 1, xph1 = xph;
        if (xd1 == D) { xd1 = 0; xph1 ^= 1; }
        if (it + 1 < n_it) {
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
        // ---- E+[i] ------------------------------------------------------------------------------------------------------------------------
        if (it > 0) {
            const bool qx_lean = i <= g.nx - 2 && !(tpm && (i < tpm || i >= g.nx - tpm - 1));
            P nx_, ny_, nz_;
            if (lean_thread && qx_lean)
                yee_e_lean<T, V, AM>(c, g, e0x, e0y, e0z, hpx, hpy, hpz, hz_jp, hx_jp, hy_kp, hx_kp, hny, hnz, nx_, ny_, nz_);
            else
                yee_e_full<T, V>(c, g, pm, sg, psi_in, psi_out, owner, can_e, i, j, k, syj, e0x, e0y, e0z, hpx, hpy, hpz, hz_jp, hx_jp,
                                 hy_kp, hx_kp, hny, hnz, nx_, ny_, nz_);
            if (owner) {
                const unsigned oe = ofs + (unsigned)i * (unsigned)g.sx;
                st8<T, V>(out.ex + oe, nx_); st8<T, V>(out.ey + oe, ny_); st8<T, V>(out.ez + oe, nz_);
            }
        }
        e0x = e1x; e0y = e1y; e0z = e1z;
        hpx = hnx; hpy = hny; hpz = hnz;
    }
}

}  // namespace fdtd
