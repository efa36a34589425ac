// fdtd_tb2.cuh — temporally blocked fused sweep: TWO full time steps per pass over HBM.
//
// The single-step fused sweep (fdtd_fused.cuh) already moves the algorithmic minimum of one step, 48 B per
// cell-update in fp32 (read 6 arrays + write 6 arrays).  The only way below that is to not write step n+1 to HBM
// at all: this kernel carries a FOUR-stage pipeline through the register window while marching ascending in x
//     A: H1[i+3] = f(H0[i+3], E0[i+3], E0[i+4])        B: E1[i+2] = g(E0[i+2], H1[i+2], H1[i+3])
//     C: H2[i+1] = f(H1[i+1], E1[i+1], E1[i+2])        D: E2[i]   = g(E1[i],   H2[i],   H2[i+1])
// (time levels 0 = input, 1 = intermediate, 2 = output), so every array is read once and written once per TWO
// steps: 24 B per cell-update.  Each stage needs the +1 neighbours (j+1 through shared memory, k+1 through the
// next lane) of the previous stage's result, so validity shrinks by one row and one lane per stage: of R warp
// rows the first R-4 own cells, of 32 lanes the first 28; the rest are rim providers that store nothing.
//
// What the reference does BETWEEN the two steps — sources added to E1/H1, monitors sampling them
// (core/simulation.py:158-164) — happens on the register window: E sources right after stage B (stage C must
// see them, stage B of the neighbouring plane must not see H sources: they are added one iteration later, just
// before stage C consumes the own-cell H1), monitors sample the post-source values exactly once per cell (owner
// thread, own segment).  Arithmetic per cell is the same sequence of rounded operations as k_sources /
// k_monitors / the single-step kernels, so results stay bit-identical in fp64.
#pragma once
#include "fdtd_fused.cuh"

namespace fdtd {

struct MidOps {
    const SrcOp* src; int n_src;
    const SrcOp* gsrc; int n_gsrc;              // x-slabs: the right neighbour's source ops on our ghost planes
    int n_planes;                               // length of plane_flags (nx, or nx+4 with ghost planes)
    const double* amp; int n_amp; const double* prof;
    const MonOp* mon; int n_mon;
    const double* phasors; int n_phasor;
    void* rec; double2* dft; double dt;
    const int* step_ptr; int step_off;          // table row of the intermediate step = *step_ptr + step_off
    const unsigned char* plane_flags;           // per local plane: bit0 a source op covers it, bit1 a monitor op
    int op_lo, op_span;                         // flagged planes lie in [op_lo, op_lo + op_span]: a register compare
                                                // keeps the flag load out of every other iteration
};

template <typename T> struct Vec8;
template <> struct Vec8<float> { typedef float2 type; static const int V = 2; };
template <> struct Vec8<double> { typedef double type; static const int V = 1; };

template <typename T, int V> __device__ __forceinline__ Pack<T, V> ld8(const T* p, bool ok)
{
    Pack<T, V> r;
    if (!ok) {
#pragma unroll
        for (int e = 0; e < V; ++e) r.v[e] = (T)0;
        return r;
    }
    typedef typename Vec8<T>::type VT;
    union { VT q; Pack<T, V> r; } u;
    u.q = *reinterpret_cast<const VT*>(p);
    return u.r;
}
template <typename T, int V> __device__ __forceinline__ void st8(T* p, const Pack<T, V>& r)
{
    typedef typename Vec8<T>::type VT;
    union { VT q; Pack<T, V> r; } u;
    u.r = r;
    *reinterpret_cast<VT*>(p) = u.q;
}

// sources of the intermediate step on a register value (same arithmetic as k_sources)
template <typename T, int V>
__device__ __forceinline__ void mid_sources(const MidOps& m, int comp, int p, int j, int k0, int row, Pack<T, V>& v)
{
    const int n_all = m.n_src + m.n_gsrc;
    for (int q = 0; q < n_all; ++q) {
        const SrcOp& op = q < m.n_src ? m.src[q] : m.gsrc[q - m.n_src];
        if (op.comp != comp) continue;
        const unsigned dp = (unsigned)(p - op.lo[0]), dj = (unsigned)(j - op.lo[1]);
        if (dp >= (unsigned)op.n[0] || dj >= (unsigned)op.n[1]) continue;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const unsigned dk = (unsigned)(k0 + e - op.lo[2]);
            if (dk >= (unsigned)op.n[2]) continue;
            double a = m.amp[(long long)row * m.n_amp + op.table];
            if (op.prof_off >= 0) {
                const long long cell = ((long long)dp * op.n[1] + dj) * op.n[2] + dk;
                a = __dmul_rn(a, m.prof[op.prof_off + cell]);
                if (op.divisor != 1.0) a = __ddiv_rn(a, op.divisor);
            }
            v.v[e] = (T)__dadd_rn((double)v.v[e], a);
        }
    }
}

// monitors of the intermediate step (same arithmetic as k_monitors); call from the owner thread only
template <typename T, int V>
__device__ __forceinline__ void mid_monitors(const MidOps& m, int comp, int p, int j, int k0, int row, const Pack<T, V>& v)
{
    for (int q = 0; q < m.n_mon; ++q) {
        const MonOp& op = m.mon[q];
        if (op.comp != comp) continue;
        const unsigned dp = (unsigned)(p - op.lo[0]), dj = (unsigned)(j - op.lo[1]);
        if (dp >= (unsigned)op.n[0] || dj >= (unsigned)op.n[1]) continue;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const unsigned dk = (unsigned)(k0 + e - op.lo[2]);
            if (dk >= (unsigned)op.n[2]) continue;
            const long long cell = ((long long)dp * op.n[1] + dj) * op.n[2] + dk;
            if (op.record) ((T*)m.rec)[op.rec_off + (long long)row * op.cells + cell] = v.v[e];
            if (op.n_freq > 0) {
                const double d = (double)v.v[e];
                const double* ph = m.phasors + ((long long)row * m.n_phasor + op.phasor_col) * 2;
                // The update below is a chain of dependent DRAM round trips (the store of one frequency may alias the load
                // of the next as far as the compiler knows): request all accumulators of the cell first, so that the chain
                // runs on L2 hits.  On the rank that owns the DFT plane the chain was 0.2 ms of a 1.6 ms pair of steps
                // (8 GPUs, profiles/r02_tuning.md §6).
                for (int fq = 0; fq < op.n_freq; ++fq)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(m.dft + op.dft_off + (long long)fq * op.cells + cell));
                for (int fq = 0; fq < op.n_freq; ++fq) {
                    double2* a = m.dft + op.dft_off + (long long)fq * op.cells + cell;
                    double2 acc = *a;
                    acc.x = __dadd_rn(acc.x, __dmul_rn(__dmul_rn(d, ph[2 * fq]), m.dt));
                    acc.y = __dadd_rn(acc.y, __dmul_rn(__dmul_rn(d, ph[2 * fq + 1]), m.dt));
                    *a = acc;
                }
            }
        }
    }
}

constexpr int kTb2Rows = 16;           // warp rows per CTA: 12 owners + 4 rim
// Owner lanes per row.  Each of the four stages consumes one more element in +k.  With one element per lane (fp64)
// that costs a lane per stage: 28 owners.  With two elements per lane (fp32) validity shrinks by half a lane per
// stage (lane 31 keeps a valid v[0] after stage A, lane 30 stays whole after stage B, ...): 30 owners.
template <typename T> constexpr int tb2_own_lanes() { return Vec8<T>::V >= 2 ? 30 : 28; }

template <typename T, int R> constexpr size_t tb2_smem_bytes() { return 2 * (size_t)R * 8 * 32 * 8; }

template <typename T, int R, bool OPS, int AM>
__device__ __forceinline__ void
tb2_sweep(const CFields<T>& in, const Fields<T>& out, const Coefs<T>& c, const Geom& g, const FusedTiling& t,
          const MidOps& m, const int planes_alloc, const Fold& fo)
{
    constexpr int V = Vec8<T>::V;
    typedef Pack<T, V> P;
    typedef typename Vec8<T>::type VT;
    extern __shared__ __align__(16) unsigned char smem_[];
    // [parity][row][quantity 0..7][lane]: E0(z,x) of plane i+3, H1(z,x) of i+2, E1(z,x) of i+1, H2(z,x) of i
    VT (*s_x)[R][8][32] = reinterpret_cast<VT (*)[R][8][32]>(smem_);

    const int lane = threadIdx.x, row = threadIdx.y;
    const int ntiles = t.ntj * t.ntk;
    const int slot = blockIdx.x / ntiles, tile = blockIdx.x - slot * ntiles;
    const int tj = tile / t.ntk, tk = tile - tj * t.ntk;
    const int j = tj * (R - 4) + row;
    const int k = (tk * t.own_lanes + lane) * V;
    const int i0 = t.seg_lo[slot];
    const int i1 = t.seg_hi[slot];
    if (t.halo_flag && i1 + 3 >= g.nx) {          // this segment reads E0 up to plane i1+3: ghost planes start at nx
        if (threadIdx.x == 0 && threadIdx.y == 0) wait_flag_ge(t.halo_flag, t.halo_need, t.error_word, t.timeout_ns);
        __syncthreads();
    }
    const bool ld_ok = (j < g.ny) && (k < g.pz);
    const bool owner = ld_ok && row < R - 4 && lane < t.own_lanes;
    const int rown = min(row + 1, R - 1);
    // element offsets fit 32 bits (host checks planes_alloc * sx < 2^32): one IMAD.WIDE per access
    const unsigned ofs = (unsigned)j * (unsigned)g.sy + (unsigned)k;
    const bool jy1 = j < g.ny - 1, jy2 = j < g.ny - 2;
    const int step_row = m.step_ptr ? (*m.step_ptr + m.step_off) : 0;
    // interior tile: every thread of the CTA (rim rows / lanes included) is clear of the +j / +k boundary ranges
    const bool interior = (tj * (R - 4) + R - 1 < g.ny - 2) && ((tk * t.own_lanes + 31) * V + V - 1 < g.nz - 2);

    const T* pex = in.ex + ofs; const T* pey = in.ey + ofs; const T* pez = in.ez + ofs;
    const T* phx = in.hx + ofs; const T* phy = in.hy + ofs; const T* phz = in.hz + ofs;
    P z_;
#pragma unroll
    for (int e = 0; e < V; ++e) z_.v[e] = (T)0;
    // window at i = i0-3
    P e0ax = z_, e0ay = z_, e0az = z_;                                   // E0[i+2]
    unsigned po = (unsigned)i0 * (unsigned)g.sx;
    P e0bx = ld8<T, V>(pex + po, ld_ok), e0by = ld8<T, V>(pey + po, ld_ok), e0bz = ld8<T, V>(pez + po, ld_ok);   // E0[i+3]
    P ne0x = ld8<T, V>(pex + po + g.sx, ld_ok), ne0y = ld8<T, V>(pey + po + g.sx, ld_ok),
      ne0z = ld8<T, V>(pez + po + g.sx, ld_ok);                          // E0[i+4]
    P nh0x = ld8<T, V>(phx + po, ld_ok), nh0y = ld8<T, V>(phy + po, ld_ok), nh0z = ld8<T, V>(phz + po, ld_ok);   // H0[i+3]
    P h1ax = z_, h1ay = z_, h1az = z_, h1bx = z_, h1by = z_, h1bz = z_;  // H1[i+1], H1[i+2]
    P e1ax = z_, e1ay = z_, e1az = z_, e1bx = z_, e1by = z_, e1bz = z_;  // E1[i], E1[i+1]
    P h2ax = z_, h2ay = z_, h2az = z_;                                   // H2[i]

    for (int i = i0 - 3; i < i1; ++i) {
        const int par = (i - i0 + 3) & 1;
        // ---- prefetch for the next iteration: E0[i+5], H0[i+4] ----------------------------------------------
        const bool more = ld_ok && (i + 1 < i1);
        const bool pe_ok = more && (i + 5 < planes_alloc), ph_ok = more && (i + 4 < planes_alloc);
        const unsigned pp = (unsigned)(i + 4) * (unsigned)g.sx;
        const P pe0x = ld8<T, V>(pex + pp + g.sx, pe_ok), pe0y = ld8<T, V>(pey + pp + g.sx, pe_ok),
                pe0z = ld8<T, V>(pez + pp + g.sx, pe_ok);
        const P ph0x = ld8<T, V>(phx + pp, ph_ok), ph0y = ld8<T, V>(phy + pp, ph_ok), ph0z = ld8<T, V>(phz + pp, ph_ok);

        // ---- intermediate-step H sources / monitors on H1[i+1] (all of its pre-source uses are done) --------------
        if (OPS && (unsigned)(i + 1 - m.op_lo) <= (unsigned)m.op_span && i + 1 >= i0) {
            const unsigned char fl = m.plane_flags[i + 1];
            if (fl & 1) {
                mid_sources<T, V>(m, 3, i + 1, j, k, step_row, h1ax);
                mid_sources<T, V>(m, 4, i + 1, j, k, step_row, h1ay);
                mid_sources<T, V>(m, 5, i + 1, j, k, step_row, h1az);
            }
            if ((fl & 2) && owner && i + 1 < i1) {
                mid_monitors<T, V>(m, 3, i + 1, j, k, step_row, h1ax);
                mid_monitors<T, V>(m, 4, i + 1, j, k, step_row, h1ay);
                mid_monitors<T, V>(m, 5, i + 1, j, k, step_row, h1az);
            }
        }
        // ---- publish the j+1 inputs of all four stages ------------------------------------------------------------------
        {
            union { VT q; P r; } u;
            u.r = e0bz; s_x[par][row][0][lane] = u.q;  u.r = e0bx; s_x[par][row][1][lane] = u.q;
            u.r = h1bz; s_x[par][row][2][lane] = u.q;  u.r = h1bx; s_x[par][row][3][lane] = u.q;
            u.r = e1bz; s_x[par][row][4][lane] = u.q;  u.r = e1bx; s_x[par][row][5][lane] = u.q;
            u.r = h2az; s_x[par][row][6][lane] = u.q;  u.r = h2ax; s_x[par][row][7][lane] = u.q;
        }
        __syncthreads();
        P e0z_j, e0x_j, h1z_j, h1x_j, e1z_j, e1x_j, h2z_j, h2x_j;
        {
            union { VT q; P r; } u;
            u.q = s_x[par][rown][0][lane]; e0z_j = u.r;  u.q = s_x[par][rown][1][lane]; e0x_j = u.r;
            u.q = s_x[par][rown][2][lane]; h1z_j = u.r;  u.q = s_x[par][rown][3][lane]; h1x_j = u.r;
            u.q = s_x[par][rown][4][lane]; e1z_j = u.r;  u.q = s_x[par][rown][5][lane]; e1x_j = u.r;
            u.q = s_x[par][rown][6][lane]; h2z_j = u.r;  u.q = s_x[par][rown][7][lane]; h2x_j = u.r;
        }
        const T e0y_n = shfl_next<T>(e0by.v[0]), e0x_n = shfl_next<T>(e0bx.v[0]);
        const T h1y_n = shfl_next<T>(h1by.v[0]), h1x_n = shfl_next<T>(h1bx.v[0]);
        const T e1y_n = shfl_next<T>(e1by.v[0]), e1x_n = shfl_next<T>(e1bx.v[0]);
        const T h2y_n = shfl_next<T>(h2ay.v[0]), h2x_n = shfl_next<T>(h2ax.v[0]);

#define TB2_STAGES(MASKED, STEADY)                                                                                      \
        stage_h<T, V, MASKED, AM>(c, g, fo, g.x0 + i + 3, jy1, jy2, k, nh0x, nh0y, nh0z, e0bx, e0by, e0bz, e0z_j, e0x_j, e0y_n,  \
                              e0x_n, ne0y, ne0z, h1cx, h1cy, h1cz);                                                  \
        stage_e<T, V, MASKED, AM>(c, g, fo, g.x0 + i + 2, jy1, k, e0ax, e0ay, e0az, h1bx, h1by, h1bz, h1z_j, h1x_j, h1y_n,        \
                              h1x_n, h1cy, h1cz, e1cx, e1cy, e1cz);                                                  \
        if (OPS && (unsigned)(i + 2 - m.op_lo) <= (unsigned)m.op_span && (STEADY || i + 2 >= i0)) {                 \
            const unsigned char fl = m.plane_flags[i + 2];                                                         \
            if (fl & 1) {                                                                                          \
                mid_sources<T, V>(m, 0, i + 2, j, k, step_row, e1cx);                                              \
                mid_sources<T, V>(m, 1, i + 2, j, k, step_row, e1cy);                                              \
                mid_sources<T, V>(m, 2, i + 2, j, k, step_row, e1cz);                                              \
            }                                                                                                      \
            if ((fl & 2) && owner && (STEADY || i + 2 < i1)) {                                                     \
                mid_monitors<T, V>(m, 0, i + 2, j, k, step_row, e1cx);                                             \
                mid_monitors<T, V>(m, 1, i + 2, j, k, step_row, e1cy);                                             \
                mid_monitors<T, V>(m, 2, i + 2, j, k, step_row, e1cz);                                             \
            }                                                                                                      \
        }                                                                                                          \
        stage_h<T, V, MASKED, AM>(c, g, fo, g.x0 + i + 1, jy1, jy2, k, h1ax, h1ay, h1az, e1bx, e1by, e1bz, e1z_j, e1x_j, e1y_n,  \
                              e1x_n, e1cy, e1cz, h2bx, h2by, h2bz);                                                  \
        if (owner && (STEADY || (i + 1 >= i0 && i + 1 < i1))) {                                                    \
            st8<T, V>(out.hx + ost1, h2bx); st8<T, V>(out.hy + ost1, h2by);                                        \
            st8<T, V>(out.hz + ost1, h2bz);                                                                        \
        }                                                                                                          \
        if (STEADY || i >= i0) {                                                                                   \
            P e2x, e2y, e2z;                                                                                       \
            stage_e<T, V, MASKED, AM>(c, g, fo, g.x0 + i, jy1, k, e1ax, e1ay, e1az, h2ax, h2ay, h2az, h2z_j, h2x_j, h2y_n,     \
                                  h2x_n, h2by, h2bz, e2x, e2y, e2z);                                               \
            if (owner) { st8<T, V>(out.ex + ost, e2x); st8<T, V>(out.ey + ost, e2y); st8<T, V>(out.ez + ost, e2z); } \
        }

        // ---- A: H1[i+3], B: E1[i+2] (+ intermediate-step E sources / monitors), C: H2[i+1], D: E2[i] -------------------
        // Interior CTAs on interior planes skip every boundary mask (the reference's "never updated" ranges and the
        // staggered array ends only touch the last two rows / columns / planes).
        P h1cx, h1cy, h1cz, e1cx, e1cy, e1cz, h2bx, h2by, h2bz;
        const unsigned ost = ofs + (unsigned)i * (unsigned)g.sx;            // plane i   (used only when i   >= i0 >= 0)
        const unsigned ost1 = ofs + (unsigned)(i + 1) * (unsigned)g.sx;     // plane i+1 (used only when i+1 >= i0 >= 0)
        if (!(interior && g.x0 + i + 3 < g.nxg - 2)) { TB2_STAGES(true, false) }
        else if (i >= i0 && i + 2 < i1) { TB2_STAGES(false, true) }
        else { TB2_STAGES(false, false) }
#undef TB2_STAGES
        // ---- rotate ---------------------------------------------------------------------------------------------------------------------
        e0ax = e0bx; e0ay = e0by; e0az = e0bz;
        e0bx = ne0x; e0by = ne0y; e0bz = ne0z;
        ne0x = pe0x; ne0y = pe0y; ne0z = pe0z;
        nh0x = ph0x; nh0y = ph0y; nh0z = ph0z;
        h1ax = h1bx; h1ay = h1by; h1az = h1bz;
        h1bx = h1cx; h1by = h1cy; h1bz = h1cz;
        e1ax = e1bx; e1ay = e1by; e1az = e1bz;
        e1bx = e1cx; e1by = e1cy; e1bz = e1cz;
        h2ax = h2bx; h2ay = h2by; h2az = h2bz;
    }
}

// One launch covers every x-segment of the pair of steps.  Whether a segment carries sources / monitors is uniform per
// CTA, so op-free segments run a loop without any op code (measured +10 %) and nothing serialises between the two kinds.
template <typename T, int R, int AM>
__global__ void __launch_bounds__(32 * R, 1)
k_fused3d_tb2(const __grid_constant__ CFields<T> in, const __grid_constant__ Fields<T> out,
              const __grid_constant__ Coefs<T> c, const __grid_constant__ Geom g, const __grid_constant__ FusedTiling t,
              const __grid_constant__ MidOps m, const int planes_alloc, const __grid_constant__ Fold fo)
{
    const int slot = blockIdx.x / (t.ntj * t.ntk);
    if ((t.seg_ops >> slot) & 1ull) tb2_sweep<T, R, true, AM>(in, out, c, g, t, m, planes_alloc, fo);
    else tb2_sweep<T, R, false, AM>(in, out, c, g, t, m, planes_alloc, fo);
}

}  // namespace fdtd
