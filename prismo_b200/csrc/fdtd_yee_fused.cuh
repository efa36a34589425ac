// fdtd_yee_fused.cuh — physics mode (stable Yee leap-frog + CPML) as ONE fused sweep per step.
//
// OPT-IN and NOT YET MEASURED: selected with fdtd_set_option(e, "yee_fused", 1) / FDTD_B200_YEE_FUSED=1; the default
// physics path is the two-pass pair in fdtd_yee.cuh, which this kernel must reproduce bit for bit in fp64
// (tools/check_yee_fused.py).  Same arithmetic (Ar<T>::diff, separately rounded CPML recursion, upd_h / upd_e), same
// update ranges; what changes is the data movement: every array is read once and written once per step (48 B per
// cell-update instead of 72), the psi arrays are touched by slab threads only, inside the same sweep.
//
// Dependences on the reference's staggering with backward differences in H and forward differences in E:
//   H+[p]   needs H[p], E[p] (own, j-1, k-1) and E[p-1]
//   E+[q]   needs E[q], H+[q] (own, j+1, k+1) and H+[q+1]
// A CTA owns a (j,k) tile and marches ascending in x with the register window E[i], E[i+1], H+[i]; iteration i produces
// H+[i+1] and E+[i].  j-1 / j+1 neighbours travel through double-buffered shared memory (one barrier per plane), k-1 /
// k+1 through warp shuffles.  Row 0 and lane 0 fetch their j-1 / k-1 inputs straight from global memory, so only the
// +j side needs a rim row (it recomputes H+ of the next tile's first row) and the +k side rim lanes: 15 owner rows of
// 16, 30 owner lanes of 32, exactly the tiling of k_fused3d.  Rim threads and the segment prologue recompute H+ (and
// its psi recursion) without storing, so psi is ping-ponged like the fields: read from the input set, written to the
// output set by the owner only.
#pragma once
#include "fdtd_fused.cuh"
#include "fdtd_yee.cuh"

namespace fdtd {

struct PsiOut { void* p[12]; };

// psi <- b psi + a d ;  d_eff = d / kappa + psi        (same rounding sequence as cpml_apply in fdtd_yee.cuh)
template <typename T>
__device__ __forceinline__ T cpml_step(T d, const T* psi_in, T* psi_out, bool store, long long o, const CpmlAxis& ax, int v, int n)
{
    const T p = CpmlMath<T>::psi(cpml_tab<T>(ax, v)[n], psi_in[o], cpml_tab<T>(ax, v + 1)[n], d);
    if (store) psi_out[o] = p;
    return CpmlMath<T>::eff(cpml_tab<T>(ax, v + 2)[n], d, p);
}

template <typename T> __device__ __forceinline__ T shfl_prev(T v) { return __shfl_up_sync(0xffffffffu, v, 1); }

template <typename T, int TJ>
__global__ void __launch_bounds__(32 * (TJ + 1), 1)
k_fused3d_yee(CFields<T> in, Fields<T> out, Coefs<T> c, Geom g, FusedTiling t, Cpml pm, PsiOut pout, SlabGeom sg)
{
    constexpr int V = VecOf<T>::V;
    typedef Pack<T, V> P;
    typedef typename VecOf<T>::type VT;
    extern __shared__ __align__(16) unsigned char smem_[];
    // [parity][row][Ez, Ex of plane i+1 | Hz+, Hx+ of plane i][lane]
    VT (*s_x)[TJ + 1][4][32] = reinterpret_cast<VT (*)[TJ + 1][4][32]>(smem_);

    const int lane = threadIdx.x, row = threadIdx.y;
    const int ntiles = t.ntj * t.ntk;
    const int seg = blockIdx.x / ntiles, tile = blockIdx.x - seg * ntiles;
    const int tj = tile / t.ntk, tk = tile - tj * t.ntk;
    const int j = tj * TJ + row;
    const int k = (tk * t.own_lanes + lane) * V;
    const int i0 = t.i_begin + seg * t.lx;
    const int i1 = min(i0 + t.lx, t.i_end);
    const bool ld_ok = (j < g.ny) && (k < g.pz);
    const bool owner = ld_ok && row < TJ && lane < t.own_lanes;
    const bool rim_row = row == TJ;
    const long long o = (long long)j * g.sy + k;

    const T* pex = in.ex + o; const T* pey = in.ey + o; const T* pez = in.ez + o;
    const T* phx = in.hx + o; const T* phy = in.hy + o; const T* phz = in.hz + o;

    // CPML bookkeeping that does not depend on the plane
    const int tpm = pm.t;
    const int syj = tpm ? slab_index(j, g.ny, tpm) : -1;
    const T* const* psi_in = reinterpret_cast<const T* const*>(pm.psi);
    T* const* psi_out = reinterpret_cast<T* const*>(pout.p);

    // window: E[i] (only Ey, Ez are used), E[i+1], H+[i]; entry state is i = i0 - 1
    P e0y = zero_pack<T>(), e0z = e0y, e0x = e0y, hpx = e0y, hpy = e0y, hpz = e0y;
    long long po = (long long)i0 * g.sx;                       // plane i + 1
    if (i0 > 0) {                                              // E[i0-1] feeds the x-differences of H+[i0]
        e0y = ldv_if<T>(pey + po - g.sx, ld_ok); e0z = ldv_if<T>(pez + po - g.sx, ld_ok);
    }
    P e1x = ldv_if<T>(pex + po, ld_ok), e1y = ldv_if<T>(pey + po, ld_ok), e1z = ldv_if<T>(pez + po, ld_ok);
    P h1x = ldv_if<T>(phx + po, ld_ok), h1y = ldv_if<T>(phy + po, ld_ok), h1z = ldv_if<T>(phz + po, ld_ok);

    for (int i = i0 - 1; i < i1; ++i) {
        const int par = (i - i0 + 1) & 1;
        const int p = i + 1;                                   // plane of the H stage
        po = (long long)p * g.sx;
        // ---- prefetch for the next iteration: E[i+2], H[i+2] ---------------------------------------------------------
        const bool more = (i + 1 < i1) && ld_ok;
        const P n_ex = ldv_if<T>(pex + po + g.sx, more), n_ey = ldv_if<T>(pey + po + g.sx, more),
                n_ez = ldv_if<T>(pez + po + g.sx, more);
        const P n_hx = ldv_if<T>(phx + po + g.sx, more), n_hy = ldv_if<T>(phy + po + g.sx, more),
                n_hz = ldv_if<T>(phz + po + g.sx, more);
        // ---- publish: the row above needs our Ez, Ex of plane p (its j-1); the row below our Hz+, Hx+ of plane i (its j+1)
        {
            union { VT q; P r; } u;
            u.r = e1z; s_x[par][row][0][lane] = u.q;
            u.r = e1x; s_x[par][row][1][lane] = u.q;
            u.r = hpz; s_x[par][row][2][lane] = u.q;
            u.r = hpx; s_x[par][row][3][lane] = u.q;
        }
        __syncthreads();
        P ez_jm, ex_jm, hz_jp = zero_pack<T>(), hx_jp = hz_jp;
        {
            union { VT q; P r; } u;
            if (row > 0) {
                u.q = s_x[par][row - 1][0][lane]; ez_jm = u.r;
                u.q = s_x[par][row - 1][1][lane]; ex_jm = u.r;
            } else {                                           // first row of the tile: j-1 belongs to another CTA
                const bool ok = ld_ok && j >= 1;
                ez_jm = ldv_if<T>(pez + po - g.sy, ok);
                ex_jm = ldv_if<T>(pex + po - g.sy, ok);
            }
            if (!rim_row) {
                u.q = s_x[par][row + 1][2][lane]; hz_jp = u.r;
                u.q = s_x[par][row + 1][3][lane]; hx_jp = u.r;
            }
        }
        // ---- k-1 from the previous lane (lane 0: global), k+1 from the next lane ----------------------------------------
        T ey_km = shfl_prev<T>(e1y.v[V - 1]), ex_km = shfl_prev<T>(e1x.v[V - 1]);
        if (lane == 0) {
            const bool ok = ld_ok && k >= 1;
            ey_km = ok ? pey[po - 1] : (T)0;
            ex_km = ok ? pex[po - 1] : (T)0;
        }
        const T hy_kp = shfl_next<T>(hpy.v[0]), hx_kp = shfl_next<T>(hpx.v[0]);

        // ---- H+[p] ----------------------------------------------------------------------------------------------------
        const int sxp = tpm ? slab_index(p, g.nx, tpm) : -1;
        const bool st_h = owner && p < i1;
        const bool px1 = p < g.nx - 1, pxm = p >= 1 && p <= g.nx - 2;
        const bool jm = j >= 1 && j <= g.ny - 2, jy1 = j < g.ny - 1;
        P hnx = h1x, hny = h1y, hnz = h1z;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const int ke = k + e;
            const bool km = ke >= 1 && ke <= g.nz - 2, kz1 = ke < g.nz - 1;
            const int szk = tpm ? slab_index(ke, g.nz, tpm) : -1;
            const long long ox = (long long)sxp * sg.x_sx + o + e;
            const long long oy = (long long)p * sg.y_sx + (long long)syj * g.sy + ke;
            const long long oz = ((long long)p * g.ny + j) * sg.z_pitch + szk;
            const T ey_k = e > 0 ? e1y.v[(e + V - 1) % V] : ey_km;
            const T ex_k = e > 0 ? e1x.v[(e + V - 1) % V] : ex_km;
            if (px1 && jm && km) {                             // Hx(p+1/2, j, k)
                T dy = Ar<T>::diff(e1z.v[e], ez_jm.v[e], g.dy, g.rdy), dz = Ar<T>::diff(e1y.v[e], ey_k, g.dz, g.rdz);
                if (syj >= 0) dy = cpml_step<T>(dy, psi_in[6], psi_out[6], st_h, oy, pm.ax[1], 3, j);
                if (szk >= 0) dz = cpml_step<T>(dz, psi_in[7], psi_out[7], st_h, oz, pm.ax[2], 3, ke);
                hnx.v[e] = upd_h<T>(c.uda, h1x.v[e], c.udb, dy, dz);
            }
            if (pxm && jy1 && km) {                            // Hy(p, j+1/2, k)
                T dz = Ar<T>::diff(e1x.v[e], ex_k, g.dz, g.rdz), dx = Ar<T>::diff(e1z.v[e], e0z.v[e], g.dx, g.rdx);
                if (szk >= 0) dz = cpml_step<T>(dz, psi_in[8], psi_out[8], st_h, oz, pm.ax[2], 3, ke);
                if (sxp >= 0) dx = cpml_step<T>(dx, psi_in[9], psi_out[9], st_h, ox, pm.ax[0], 3, p);
                hny.v[e] = upd_h<T>(c.uda, h1y.v[e], c.udb, dz, dx);
            }
            if (pxm && jm && kz1) {                            // Hz(p, j, k+1/2)
                T dx = Ar<T>::diff(e1y.v[e], e0y.v[e], g.dx, g.rdx), dy = Ar<T>::diff(e1x.v[e], ex_jm.v[e], g.dy, g.rdy);
                if (sxp >= 0) dx = cpml_step<T>(dx, psi_in[10], psi_out[10], st_h, ox, pm.ax[0], 3, p);
                if (syj >= 0) dy = cpml_step<T>(dy, psi_in[11], psi_out[11], st_h, oy, pm.ax[1], 3, j);
                hnz.v[e] = upd_h<T>(c.uda, h1z.v[e], c.udb, dx, dy);
            }
        }
        if (st_h) { stv_pol<T, 0>(out.hx + o + po, hnx); stv_pol<T, 0>(out.hy + o + po, hny); stv_pol<T, 0>(out.hz + o + po, hnz); }

        // ---- E+[i] ----------------------------------------------------------------------------------------------------
        if (i >= i0) {
            const int sxq = tpm ? slab_index(i, g.nx, tpm) : -1;
            const bool qx1 = i < g.nx - 1;
            const bool st_e = owner;
            const long long pe = (long long)i * g.sx;
            P nx_ = e0x, ny_ = e0y, nz_ = e0z;
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const int ke = k + e;
                const bool kz0 = ke < g.nz, kz1 = ke < g.nz - 1;
                const int szk = tpm ? slab_index(ke, g.nz, tpm) : -1;
                const long long ox = (long long)sxq * sg.x_sx + o + e;
                const long long oy = (long long)i * sg.y_sx + (long long)syj * g.sy + ke;
                const long long oz = ((long long)i * g.ny + j) * sg.z_pitch + szk;
                const T hy_k = (e + 1 < V) ? hpy.v[(e + 1) % V] : hy_kp;
                const T hx_k = (e + 1 < V) ? hpx.v[(e + 1) % V] : hx_kp;
                if (jy1 && kz1 && !rim_row) {                  // Ex(i, j+1/2, k+1/2)
                    T dy = Ar<T>::diff(hz_jp.v[e], hpz.v[e], g.dy, g.rdy), dz = Ar<T>::diff(hy_k, hpy.v[e], g.dz, g.rdz);
                    if (syj >= 0) dy = cpml_step<T>(dy, psi_in[0], psi_out[0], st_e, oy, pm.ax[1], 0, j);
                    if (szk >= 0) dz = cpml_step<T>(dz, psi_in[1], psi_out[1], st_e, oz, pm.ax[2], 0, ke);
                    nx_.v[e] = upd_e<T>(c.uca, e0x.v[e], c.ucb, dy, dz);
                }
                if (qx1 && kz1 && ld_ok && !rim_row) {         // Ey(i+1/2, j, k+1/2)
                    T dz = Ar<T>::diff(hx_k, hpx.v[e], g.dz, g.rdz), dx = Ar<T>::diff(hnz.v[e], hpz.v[e], g.dx, g.rdx);
                    if (szk >= 0) dz = cpml_step<T>(dz, psi_in[2], psi_out[2], st_e, oz, pm.ax[2], 0, ke);
                    if (sxq >= 0) dx = cpml_step<T>(dx, psi_in[3], psi_out[3], st_e, ox, pm.ax[0], 0, i);
                    ny_.v[e] = upd_e<T>(c.uca, e0y.v[e], c.ucb, dz, dx);
                }
                if (qx1 && jy1 && kz0 && !rim_row) {           // Ez(i+1/2, j+1/2, k)
                    T dx = Ar<T>::diff(hny.v[e], hpy.v[e], g.dx, g.rdx), dy = Ar<T>::diff(hx_jp.v[e], hpx.v[e], g.dy, g.rdy);
                    if (sxq >= 0) dx = cpml_step<T>(dx, psi_in[4], psi_out[4], st_e, ox, pm.ax[0], 0, i);
                    if (syj >= 0) dy = cpml_step<T>(dy, psi_in[5], psi_out[5], st_e, oy, pm.ax[1], 0, j);
                    nz_.v[e] = upd_e<T>(c.uca, e0z.v[e], c.ucb, dx, dy);
                }
            }
            if (st_e) { stv_pol<T, 0>(out.ex + o + pe, nx_); stv_pol<T, 0>(out.ey + o + pe, ny_); stv_pol<T, 0>(out.ez + o + pe, nz_); }
        }
        // ---- rotate the window ----------------------------------------------------------------------------------------
        e0x = e1x; e0y = e1y; e0z = e1z;
        e1x = n_ex; e1y = n_ey; e1z = n_ez;
        hpx = hnx; hpy = hny; hpz = hnz;
        h1x = n_hx; h1y = n_hy; h1z = n_hz;
    }
}

}  // namespace fdtd
