// fdtd_het.cuh — one-step fused sweep for HETEROGENEOUS media (cell-centred Ca, Cb, Da, Db arrays).
//
// Same idea as fdtd_fused.cuh (march ascending in x with a register window, k+1 from the next lane, j+1 from the
// next warp-row through shared memory, ping-pong output) but the window also carries the four coefficient arrays,
// so one step reads 6 + 4 arrays and writes 6: 64 B per cell-update in fp32 instead of the two-pass
// kernels' 104 B.  The 2- / 4-point averaging that the reference recomputes into 12 full-grid temporaries every
// step (core/solver.py:458-501) happens in registers, in the reference's summation order (fp64 stays bit-exact):
//   Hx: mean2(D[i+1], D[i+2])           Hy: mean2(D[j], D[j+1])               Hz: mean2(D[k], D[k+1])      at plane i+1
//   Ex: mean4(C, C[j+1], C[k+1], C[j+1,k+1])   Ey: mean4(C[i], C[i+1], C[i,k+1], C[i+1,k+1])
//   Ez: mean4(C[i], C[i+1], C[i,j+1], C[i+1,j+1])                                                         at plane i
// 8-byte vectors per thread (2 fp32 / 1 fp64 cells) keep the window in registers.  Of R warp rows the first R-2
// own cells (row R-2 recomputes H+ on the rim, row R-1 only provides raw neighbours); of 32 lanes the first 30.
#pragma once
#include <type_traits>

#include "fdtd_tb2.cuh"

namespace fdtd {

// fp64: the divisions by the launch-constant spacings go through the correctly rounded FMA sequence (Ar<double>::div_fast)
// with ONE range check per stage; a stage whose check fails (inf / nan / denormal quotients) is recomputed with true
// divisions.  The first version used the per-division fallback (two branches per division): 10 Gcell/s on c3.
template <typename T, bool SLOW> __device__ __forceinline__ T het_diff(T a1, T a0, double d, Rcp r, unsigned& bad)
{
    if (SLOW) return Ar<T>::diff_exact(a1, a0, d, r);
    return Ar<T>::diff_fast(a1, a0, d, r, bad);
}

constexpr int kHetRows = 16;
constexpr int kHetOwnLanes = 30;
template <typename T, int R, bool ANISO = false> constexpr size_t het_smem_bytes() { return 2 * (size_t)R * (ANISO ? 9 : 8) * 32 * 8; }

// ADE: apply the dispersive-medium recursions of the PREVIOUS step on the E stage's input values (ade_in_sweep).
// ANISO (opt-in extension): per-component Cb — a diagonal permittivity tensor, the case the reference's
// AnisotropicUpdater handles with one division per component (materials/tensor.py:508-514) — inside the same E stage:
// Ex averages c.cb (the eps_xx array) over j and k exactly as before, Ey averages c.cby over i and k, Ez averages c.cbz
// over i and j.  Two more arrays are streamed (72 B per cell-update in fp32), one more smem slot carries Cb_z to row - 1.
// IDX (material-index coding): media painted from a shape list hold a handful of distinct materials, so the sweep reads ONE
// byte per cell (c.mat) instead of 4 (6) coefficient values and looks the values up in a per-CTA shared-memory copy of the
// material table when an index plane enters the register window: 49 B per cell-update in fp32 instead of 64 (72).  The
// looked-up values are the very numbers the coefficient arrays hold, so results are bit-identical to the array path.
template <typename T, int R, bool ADE, bool ANISO, bool IDX>
__device__ __forceinline__ void
het_sweep(const CFields<T>& in, const Fields<T>& out, const Coefs<T>& c, const Geom& g, const FusedTiling& t, const int planes_alloc,
          const AdeIn& ad, const int item)
{
    constexpr int V = Vec8<T>::V;
    typedef Pack<T, V> P;
    typedef typename Vec8<T>::type VT;
    extern __shared__ __align__(16) unsigned char smem_[];
    // [parity][row][q][lane], q: 0 Ez,1 Ex of plane i+1; 2 Hz+,3 Hx+ of plane i; 4 Da,5 Db of plane i+1; 6 Ca,7 Cb of plane i+1
    // (8: Cb_z of plane i+1 when ANISO)
    constexpr int NQ = ANISO ? 9 : 8;
    VT (*s_x)[R][NQ][32] = reinterpret_cast<VT (*)[R][NQ][32]>(smem_);

    const int lane = threadIdx.x, row = threadIdx.y;
    const int ntiles = t.ntj * t.ntk;
    const int seg = item / ntiles, tile = item - seg * ntiles;
    const int tj = tile / t.ntk, tk = tile - tj * t.ntk;
    const int j = tj * (R - 2) + row;
    const int k = (tk * t.own_lanes + lane) * V;
    const int i0 = t.i_begin + seg * t.lx;
    const int i1 = min(i0 + t.lx, t.i_end);
    if (t.halo_flag && i1 + 1 >= g.nx) {          // x-slabs: this segment reads E / H up to plane i1+1, ghost planes start at nx
        if (threadIdx.x == 0 && threadIdx.y == 0) wait_flag_ge(t.halo_flag, t.halo_need, t.error_word, t.timeout_ns);
        __syncthreads();
    }
    const bool ld_ok = (j < g.ny) && (k < g.pz);
    const bool owner = ld_ok && row < R - 2 && lane < t.own_lanes;
    const int rown = min(row + 1, R - 1);
    const unsigned ofs = (unsigned)j * (unsigned)g.sy + (unsigned)k;
    const bool jy1 = j < g.ny - 1, jy2 = j < g.ny - 2;

    const T* pex = in.ex + ofs; const T* pey = in.ey + ofs; const T* pez = in.ez + ofs;
    const T* phx = in.hx + ofs; const T* phy = in.hy + ofs; const T* phz = in.hz + ofs;
    const T* pca = IDX ? nullptr : c.ca + ofs; const T* pcb = IDX ? nullptr : c.cb + ofs;
    const T* pda = IDX ? nullptr : c.da + ofs; const T* pdb = IDX ? nullptr : c.db + ofs;
    P z_;
#pragma unroll
    for (int e = 0; e < V; ++e) z_.v[e] = (T)0;
    // IDX: material table in shared memory; V indices per thread and plane travel packed in one register
    __shared__ T s_tc[IDX ? kMatTabRows : 1][4];            // Ca, Cb (x), Cb_y, Cb_z
    __shared__ T s_td[IDX ? kMatTabRows : 1][2];            // Da, Db
    const unsigned char* pm = IDX ? c.mat + ofs : nullptr;
    if (IDX) {
        for (int q = threadIdx.y * 32 + threadIdx.x; q < c.n_mat; q += 32 * R) {
            const T* r = c.mat_tab + 6 * q;
            s_tc[q][0] = r[0]; s_tc[q][1] = r[1]; s_tc[q][2] = r[4]; s_tc[q][3] = r[5];
            s_td[q][0] = r[2]; s_td[q][1] = r[3];
        }
        __syncthreads();
    }
    auto ldm = [&](unsigned off, bool ok) -> unsigned {
        if (!ok) return 0u;
        if (V == 2) return *reinterpret_cast<const unsigned short*>(pm + off);
        return pm[off];
    };
    auto look_c = [&](unsigned m, P& a, P& b, P& by, P& bz) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const unsigned q = (m >> (8 * e)) & 255u;
            a.v[e] = s_tc[q][0]; b.v[e] = s_tc[q][1];
            if (ANISO) { by.v[e] = s_tc[q][2]; bz.v[e] = s_tc[q][3]; }
        }
    };
    auto look_d = [&](unsigned m, P& a, P& b) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const unsigned q = (m >> (8 * e)) & 255u;
            a.v[e] = s_td[q][0]; b.v[e] = s_td[q][1];
        }
    };

    // window at i = i0 - 1
    unsigned po = (unsigned)i0 * (unsigned)g.sx;
    P e0x = z_, e0y = z_, e0z = z_, hpx = z_, hpy = z_, hpz = z_;                        // E[i], H+[i]
    P e1x = ld8<T, V>(pex + po, ld_ok), e1y = ld8<T, V>(pey + po, ld_ok), e1z = ld8<T, V>(pez + po, ld_ok);            // E[i+1]
    P e2x = ld8<T, V>(pex + po + g.sx, ld_ok), e2y = ld8<T, V>(pey + po + g.sx, ld_ok), e2z = ld8<T, V>(pez + po + g.sx, ld_ok);
    P h1x = ld8<T, V>(phx + po, ld_ok), h1y = ld8<T, V>(phy + po, ld_ok), h1z = ld8<T, V>(phz + po, ld_ok);            // H[i+1]
    const bool pm_ok = ld_ok && i0 > 0;
    P ca0 = z_, cb0 = z_, ca1 = z_, cb1 = z_, da1 = z_, db1 = z_, da2 = z_, db2 = z_;
    P ca0_j = z_, cb0_j = z_;                                                           // C[i] at j+1
    // ANISO: Cb_y at planes i, i+1 (k+1 from the next lane), Cb_z at planes i, i+1 and their j+1 neighbours
    const T* pcby = (ANISO && !IDX) ? c.cby + ofs : nullptr; const T* pcbz = (ANISO && !IDX) ? c.cbz + ofs : nullptr;
    P cby0 = z_, cby1 = z_, cbz0 = z_, cbz1 = z_, cbz0_j = z_;
    unsigned m_a = 0, m_b = 0;                                                           // IDX: indices of planes i+2, i+3
    if (IDX) {
        look_c(ldm(po - (unsigned)g.sx, pm_ok), ca0, cb0, cby0, cbz0);                   // C[i]   (unused while i < i0)
        const unsigned m1 = ldm(po, ld_ok);
        look_c(m1, ca1, cb1, cby1, cbz1);                                                 // C[i+1]
        look_d(m1, da1, db1);                                                             // D[i+1]
        m_a = ldm(po + (unsigned)g.sx, ld_ok);
        look_d(m_a, da2, db2);                                                            // D[i+2]
        m_b = ldm(po + 2u * (unsigned)g.sx, ld_ok && (i0 + 2 < planes_alloc));
    } else {
        ca0 = ld8<T, V>(pca + po - g.sx, pm_ok); cb0 = ld8<T, V>(pcb + po - g.sx, pm_ok);     // C[i]   (unused while i < i0)
        ca1 = ld8<T, V>(pca + po, ld_ok); cb1 = ld8<T, V>(pcb + po, ld_ok);                 // C[i+1]
        da1 = ld8<T, V>(pda + po, ld_ok); db1 = ld8<T, V>(pdb + po, ld_ok);                 // D[i+1]
        da2 = ld8<T, V>(pda + po + g.sx, ld_ok); db2 = ld8<T, V>(pdb + po + g.sx, ld_ok);   // D[i+2]
        if (ANISO) {
            cby0 = ld8<T, V>(pcby + po - g.sx, pm_ok); cbz0 = ld8<T, V>(pcbz + po - g.sx, pm_ok);
            cby1 = ld8<T, V>(pcby + po, ld_ok); cbz1 = ld8<T, V>(pcbz + po, ld_ok);
        }
    }
    unsigned ade_mask = 0;
    __shared__ AdeOp s_ade[ADE ? kAdeSmemOps : 1];
    if (ADE) {
        ade_stage_ops(s_ade, ad, threadIdx.y * 32 + threadIdx.x, 32 * R);
        if (owner) ade_mask = ade_thread_mask<V>(ad, i0, i1, j, k);
        __syncthreads();
    }

    for (int i = i0 - 1; i < i1; ++i) {
        const int par = (i - i0 + 1) & 1;
        // ---- prefetch for the next iteration: E[i+3], H[i+2], C[i+2], D[i+3] ---------------------------------------
        const bool more = ld_ok && (i + 1 < i1);
        const bool p3 = more && (i + 3 < planes_alloc), p2 = more && (i + 2 < planes_alloc);
        const unsigned pp = (unsigned)(i + 2) * (unsigned)g.sx;
        const P n_ex = ld8<T, V>(pex + pp + g.sx, p3), n_ey = ld8<T, V>(pey + pp + g.sx, p3), n_ez = ld8<T, V>(pez + pp + g.sx, p3);
        const P n_hx = ld8<T, V>(phx + pp, p2), n_hy = ld8<T, V>(phy + pp, p2), n_hz = ld8<T, V>(phz + pp, p2);
        P n_ca = z_, n_cb = z_, n_da = z_, n_db = z_, n_cby = z_, n_cbz = z_;
        unsigned n_m = 0;
        if (IDX) {
            // the index planes i+2 / i+3 arrived one / two iterations ago: their table rows cost no DRAM latency
            look_c(m_a, n_ca, n_cb, n_cby, n_cbz);
            look_d(m_b, n_da, n_db);
            n_m = ldm(pp + 2u * (unsigned)g.sx, more && (i + 4 < planes_alloc));
        } else {
            n_ca = ld8<T, V>(pca + pp, p2); n_cb = ld8<T, V>(pcb + pp, p2);
            n_da = ld8<T, V>(pda + pp + g.sx, p3); n_db = ld8<T, V>(pdb + pp + g.sx, p3);
            if (ANISO) { n_cby = ld8<T, V>(pcby + pp, p2); n_cbz = ld8<T, V>(pcbz + pp, p2); }
        }
        // ---- publish the j+1 inputs ------------------------------------------------------------------------------------
        {
            union { VT q; P r; } u;
            u.r = e1z; s_x[par][row][0][lane] = u.q;  u.r = e1x; s_x[par][row][1][lane] = u.q;
            u.r = hpz; s_x[par][row][2][lane] = u.q;  u.r = hpx; s_x[par][row][3][lane] = u.q;
            u.r = da1; s_x[par][row][4][lane] = u.q;  u.r = db1; s_x[par][row][5][lane] = u.q;
            u.r = ca1; s_x[par][row][6][lane] = u.q;  u.r = cb1; s_x[par][row][7][lane] = u.q;
            if (ANISO) { u.r = cbz1; s_x[par][row][NQ - 1][lane] = u.q; }
        }
        __syncthreads();
        P ez_j, ex_j, hz_j, hx_j, da1_j, db1_j, ca1_j, cb1_j;
        {
            union { VT q; P r; } u;
            u.q = s_x[par][rown][0][lane]; ez_j = u.r;   u.q = s_x[par][rown][1][lane]; ex_j = u.r;
            u.q = s_x[par][rown][2][lane]; hz_j = u.r;   u.q = s_x[par][rown][3][lane]; hx_j = u.r;
            u.q = s_x[par][rown][4][lane]; da1_j = u.r;  u.q = s_x[par][rown][5][lane]; db1_j = u.r;
            u.q = s_x[par][rown][6][lane]; ca1_j = u.r;  u.q = s_x[par][rown][7][lane]; cb1_j = u.r;
        }
        P cbz1_j = z_;
        if (ANISO) { union { VT q; P r; } u; u.q = s_x[par][rown][NQ - 1][lane]; cbz1_j = u.r; }
        // ---- k+1 neighbours from the next lane --------------------------------------------------------------------------------
        const T ey1_n = shfl_next<T>(e1y.v[0]), ex1_n = shfl_next<T>(e1x.v[0]);
        const T hpy_n = shfl_next<T>(hpy.v[0]), hpx_n = shfl_next<T>(hpx.v[0]);
        const T da1_n = shfl_next<T>(da1.v[0]), db1_n = shfl_next<T>(db1.v[0]);
        const T ca0_n = shfl_next<T>(ca0.v[0]), cb0_n = shfl_next<T>(cb0.v[0]);
        const T ca1_n = shfl_next<T>(ca1.v[0]), cb1_n = shfl_next<T>(cb1.v[0]);
        const T ca0j_n = shfl_next<T>(ca0_j.v[0]), cb0j_n = shfl_next<T>(cb0_j.v[0]);
        T cby0_n = (T)0, cby1_n = (T)0;
        if (ANISO) { cby0_n = shfl_next<T>(cby0.v[0]); cby1_n = shfl_next<T>(cby1.v[0]); }

        // ---- H+[i+1] ----------------------------------------------------------------------------------------------------------
        const int gi1 = g.x0 + i + 1;
        const bool ix1 = gi1 < g.nxg - 1, ix2 = gi1 < g.nxg - 2;
        P hnx = h1x, hny = h1y, hnz = h1z;
        auto h_stage = [&](auto slow) -> unsigned {
            constexpr bool SLOW = decltype(slow)::value;
            unsigned bad = 0;
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const bool kz1 = (k + e) < g.nz - 1, kz2 = (k + e) < g.nz - 2;
                const T ey_k = (e + 1 < V) ? e1y.v[(e + 1) % V] : ey1_n;
                const T ex_k = (e + 1 < V) ? e1x.v[(e + 1) % V] : ex1_n;
                const T da_k = (e + 1 < V) ? da1.v[(e + 1) % V] : da1_n;
                const T db_k = (e + 1 < V) ? db1.v[(e + 1) % V] : db1_n;
                T n = upd_h<T>(mean2<T>(da1.v[e], da2.v[e]), h1x.v[e], mean2<T>(db1.v[e], db2.v[e]),
                               het_diff<T, SLOW>(ez_j.v[e], e1z.v[e], g.dy, g.rdy, bad), het_diff<T, SLOW>(ey_k, e1y.v[e], g.dz, g.rdz, bad));
                if (ix1 && jy2 && kz2) hnx.v[e] = n;
                n = upd_h<T>(mean2<T>(da1.v[e], da1_j.v[e]), h1y.v[e], mean2<T>(db1.v[e], db1_j.v[e]),
                             het_diff<T, SLOW>(ex_k, e1x.v[e], g.dz, g.rdz, bad), het_diff<T, SLOW>(e2z.v[e], e1z.v[e], g.dx, g.rdx, bad));
                if (ix2 && jy1 && kz2) hny.v[e] = n;
                n = upd_h<T>(mean2<T>(da1.v[e], da_k), h1z.v[e], mean2<T>(db1.v[e], db_k),
                             het_diff<T, SLOW>(e2y.v[e], e1y.v[e], g.dx, g.rdx, bad), het_diff<T, SLOW>(ex_j.v[e], e1x.v[e], g.dy, g.rdy, bad));
                if (ix2 && jy2 && kz1) hnz.v[e] = n;
            }
            return bad;
        };
        if (h_stage(std::false_type{})) {
            if (sizeof(T) == 8) { hnx = h1x; hny = h1y; hnz = h1z; h_stage(std::true_type{}); }
        }
        if (owner && i + 1 < i1) {
            const unsigned q = ofs + (unsigned)(i + 1) * (unsigned)g.sx;
            st8<T, V>(out.hx + q, hnx); st8<T, V>(out.hy + q, hny); st8<T, V>(out.hz + q, hnz);
        }
        // ---- E+[i] --------------------------------------------------------------------------------------------------------------
        if (i >= i0) {
            const int gi = g.x0 + i;
            const bool ex0 = gi < g.nxg, ex1 = gi < g.nxg - 1;
            P ox = e0x, oy = e0y, oz = e0z;
            double jx[V], jy[V], jz[V];
#pragma unroll
            for (int e = 0; e < V; ++e) jx[e] = jy[e] = jz[e] = 0.0;
            if (ADE) { if (ade_mask) ade_in_sweep<T, V>(ad, s_ade, ade_mask, i, j, k, e0x, e0y, e0z, jx, jy, jz); }
            auto e_stage = [&](auto slow) -> unsigned {
                constexpr bool SLOW = decltype(slow)::value;
                unsigned bad = 0;
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const bool kz0 = (k + e) < g.nz, kz1 = (k + e) < g.nz - 1;
                const T hy_k = (e + 1 < V) ? hpy.v[(e + 1) % V] : hpy_n;
                const T hx_k = (e + 1 < V) ? hpx.v[(e + 1) % V] : hpx_n;
                const T ca0_k = (e + 1 < V) ? ca0.v[(e + 1) % V] : ca0_n, cb0_k = (e + 1 < V) ? cb0.v[(e + 1) % V] : cb0_n;
                const T ca1_k = (e + 1 < V) ? ca1.v[(e + 1) % V] : ca1_n, cb1_k = (e + 1 < V) ? cb1.v[(e + 1) % V] : cb1_n;
                const T ca0_jk = (e + 1 < V) ? ca0_j.v[(e + 1) % V] : ca0j_n, cb0_jk = (e + 1 < V) ? cb0_j.v[(e + 1) % V] : cb0j_n;
                // Cb of Ey (mean over i, k) and of Ez (mean over i, j): the component's own array when ANISO
                T cbm_y, cbm_z;
                if (ANISO) {
                    const T cby0_k = (e + 1 < V) ? cby0.v[(e + 1) % V] : cby0_n, cby1_k = (e + 1 < V) ? cby1.v[(e + 1) % V] : cby1_n;
                    cbm_y = mean4<T>(cby0.v[e], cby1.v[e], cby0_k, cby1_k);
                    cbm_z = mean4<T>(cbz0.v[e], cbz1.v[e], cbz0_j.v[e], cbz1_j.v[e]);
                } else {
                    cbm_y = mean4<T>(cb0.v[e], cb1.v[e], cb0_k, cb1_k);
                    cbm_z = mean4<T>(cb0.v[e], cb1.v[e], cb0_j.v[e], cb1_j.v[e]);
                }
                T n = upd_e<T>(mean4<T>(ca0.v[e], ca0_j.v[e], ca0_k, ca0_jk), e0x.v[e],
                               mean4<T>(cb0.v[e], cb0_j.v[e], cb0_k, cb0_jk),
                               het_diff<T, SLOW>(hz_j.v[e], hpz.v[e], g.dy, g.rdy, bad), het_diff<T, SLOW>(hy_k, hpy.v[e], g.dz, g.rdz, bad));
                if (ADE) { if (ad.coupled) n = Ar<T>::sub(n, Ar<T>::mul(mean4<T>(cb0.v[e], cb0_j.v[e], cb0_k, cb0_jk), (T)jx[e])); }
                if (ex0 && jy1 && kz1) ox.v[e] = n;
                n = upd_e<T>(mean4<T>(ca0.v[e], ca1.v[e], ca0_k, ca1_k), e0y.v[e], cbm_y,
                             het_diff<T, SLOW>(hx_k, hpx.v[e], g.dz, g.rdz, bad), het_diff<T, SLOW>(hnz.v[e], hpz.v[e], g.dx, g.rdx, bad));
                if (ADE) { if (ad.coupled) n = Ar<T>::sub(n, Ar<T>::mul(cbm_y, (T)jy[e])); }
                if (ex1 && kz1) oy.v[e] = n;
                n = upd_e<T>(mean4<T>(ca0.v[e], ca1.v[e], ca0_j.v[e], ca1_j.v[e]), e0z.v[e], cbm_z,
                             het_diff<T, SLOW>(hny.v[e], hpy.v[e], g.dx, g.rdx, bad), het_diff<T, SLOW>(hx_j.v[e], hpx.v[e], g.dy, g.rdy, bad));
                if (ADE) { if (ad.coupled) n = Ar<T>::sub(n, Ar<T>::mul(cbm_z, (T)jz[e])); }
                if (ex1 && jy1 && kz0) oz.v[e] = n;
            }
                return bad;
            };
            if (e_stage(std::false_type{})) {
                if (sizeof(T) == 8) { ox = e0x; oy = e0y; oz = e0z; e_stage(std::true_type{}); }
            }
            if (owner) {
                const unsigned q = ofs + (unsigned)i * (unsigned)g.sx;
                st8<T, V>(out.ex + q, ox); st8<T, V>(out.ey + q, oy); st8<T, V>(out.ez + q, oz);
            }
        }
        // ---- rotate -----------------------------------------------------------------------------------------------------------------
        e0x = e1x; e0y = e1y; e0z = e1z;
        e1x = e2x; e1y = e2y; e1z = e2z;
        e2x = n_ex; e2y = n_ey; e2z = n_ez;
        hpx = hnx; hpy = hny; hpz = hnz;
        h1x = n_hx; h1y = n_hy; h1z = n_hz;
        ca0 = ca1; cb0 = cb1; ca1 = n_ca; cb1 = n_cb;
        ca0_j = ca1_j; cb0_j = cb1_j;
        if (ANISO) { cby0 = cby1; cby1 = n_cby; cbz0 = cbz1; cbz1 = n_cbz; cbz0_j = cbz1_j; }
        if (IDX) { m_a = m_b; m_b = n_m; }
        da1 = da2; db1 = db2; da2 = n_da; db2 = n_db;
    }
}

template <typename T, int R, bool ADE, bool ANISO, bool IDX>
__global__ void __launch_bounds__(32 * R, 1)
k_fused3d_het(const __grid_constant__ CFields<T> in, const __grid_constant__ Fields<T> out, const __grid_constant__ Coefs<T> c,
              const __grid_constant__ Geom g, const __grid_constant__ FusedTiling t, const int planes_alloc,
              const __grid_constant__ AdeIn ad)
{
    if (ADE) {
        // only the CTAs whose tile and x-segment meet a recursion box run the body that carries the recursion code
        constexpr int V = Vec8<T>::V;
        const int item = ad.order ? ad.order[blockIdx.x] : (int)blockIdx.x;
        const int ntiles = t.ntj * t.ntk;
        const int seg = item / ntiles, tile = item - seg * ntiles;
        const int tj = tile / t.ntk, tk = tile - tj * t.ntk;
        const int i0 = t.i_begin + seg * t.lx, i1 = min(i0 + t.lx, t.i_end);
        if (ade_tile_touched(ad, i0, i1, tj * (R - 2), tj * (R - 2) + R - 2, tk * t.own_lanes * V, (tk + 1) * t.own_lanes * V)) {
            het_sweep<T, R, true, ANISO, IDX>(in, out, c, g, t, planes_alloc, ad, item);
            return;
        }
        het_sweep<T, R, false, ANISO, IDX>(in, out, c, g, t, planes_alloc, ad, item);
        return;
    }
    het_sweep<T, R, false, ANISO, IDX>(in, out, c, g, t, planes_alloc, ad, (int)blockIdx.x);
}

}  // namespace fdtd
