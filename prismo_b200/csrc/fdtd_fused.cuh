// fdtd_fused.cuh — single-sweep fused H+E step for 3-D grids with uniform coefficients.
//
// One launch = one full time step: every field array is read once and written once (48 B per cell
// in fp32, the algorithmic minimum), instead of the two-pass kernels' 72 B.
//
// Dependency shape (both curls are forward differences, /root/reference/src/prismo/core/solver.py:178-300):
//     H+[i] = f(H[i], E[i], E[i+1])            E+[i] = g(E[i], H+[i], H+[i+1])
// so a CTA that owns a (j,k) tile can march ASCENDING in x with a register window: at iteration i it
// holds E[i], E[i+1], H+[i], receives the prefetched E[i+2], H[i+1], computes H+[i+1] and then E+[i].
// +1 neighbours: k+1 comes from the next lane (warp shuffle), j+1 from the next warp-row (shared
// memory, double buffered: one __syncthreads per plane).  The last warp-row and the last two lanes of
// every row are halo providers: they recompute H+ on the tile's +j / +k rim and store nothing.
// Output goes to the OTHER buffer set (ping-pong): a neighbouring CTA re-reads this tile's rim from
// the input set, so in-place stores would race.  Cells the reference never updates, and padding, are
// copied through unchanged.
//
// Work item = (x-segment, tile); items are enumerated segment-major so that CTAs resident at the
// same time work on neighbouring tiles of the same planes and share their rims in L2.
#pragma once
#include "fdtd_kernels.cuh"

namespace fdtd {

constexpr int kMaxSegs = 48;

struct FusedTiling {
    int i_begin, i_end;     // planes [i_begin, i_end) are produced by this launch
    int lx;                 // planes per segment
    int nseg, ntj, ntk;     // items = nseg * ntj * ntk
    int own_lanes;          // lanes per row that own (store) cells; lanes >= own_lanes are rim providers
    // two-step sweep only: explicit x-segments [seg_lo, seg_hi) in dispatch order (bulk first, narrow op-carrying
    // ones next, the one that reads the ghost planes as late as possible); bit s of seg_ops = segment s has sources /
    // monitors on its planes and runs the op-carrying code path
    int seg_lo[kMaxSegs], seg_hi[kMaxSegs];
    unsigned long long seg_ops;
    // x-slabs: items that read the ghost planes (their segment ends at plane nx) first wait until the right
    // neighbour has pushed them:  *halo_flag >= halo_need  (system-scope acquire; null = no wait)
    const int* halo_flag;
    int halo_need;
    int* error_word;        // set to 1 if the wait times out (peer died): never hang the GPU
    unsigned long long timeout_ns;
};

__device__ __forceinline__ int ld_acquire_sys(const int* p)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v)
{
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// spin until *flag >= need (or ~20 s pass: then flag the error and carry on with whatever is there)
__device__ __forceinline__ void wait_flag_ge(const int* flag, int need, int* error_word, unsigned long long timeout_ns)
{
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flag) < need) {
        __nanosleep(256);
        if (global_ns() - t0 > timeout_ns) { if (error_word) atomicExch(error_word, 1); break; }
    }
}
__global__ void k_signal(int* flag, int v) { __threadfence_system(); st_release_sys(flag, v); }
__global__ void k_signal_wait(int* flag, int v, const int* other, int need, int* error_word, unsigned long long timeout_ns)
{
    __threadfence_system();
    st_release_sys(flag, v);
    wait_flag_ge(other, need, error_word, timeout_ns);
}
__global__ void k_wait(const int* flag, int need, int* error_word, unsigned long long timeout_ns)
{
    wait_flag_ge(flag, need, error_word, timeout_ns);
}

struct FusedPlan { bool attr_set[4] = {false, false, false, false}; };
static inline void fused_release(FusedPlan&) {}
constexpr int kFusedTJ = 15;        // 15 owner rows + 1 rim row = 16 warps, one CTA per SM

template <typename T, int TJ> constexpr size_t fused_smem_bytes()
{
    return 2 * 2 * (size_t)(TJ + 1) * 2 * 32 * sizeof(typename VecOf<T>::type);
}

// Folded arithmetic: coefficient * 1/d products computed once on the host, so an update is 2 ADD + 1 MUL + 2 FMA
// instead of 7 separately rounded operations with two divisions.  fp32 always uses it.  fp64 uses it only in the
// opt-in FAST mode (FDTD_FLAG_FAST_F64: results within ~1e-15 relative per step of the exact mode, far inside the
// 1e-10 north-star tolerance); the default fp64 mode keeps the reference's exact operation sequence and is
// bit-identical to NumPy.
struct Fold {
    float f[6];             // db/dx, db/dy, db/dz, cb/dx, cb/dy, cb/dz
    double d[6];
};
enum { FHX = 0, FHY = 1, FHZ = 2, FEX = 3, FEY = 4, FEZ = 5 };

// AM = arithmetic mode of the fp64 instantiations: 0 exact (reference operation sequence, bit-identical to NumPy),
// 1 folded (FDTD_FLAG_FAST_F64).  fp32 is always folded.  SLOW = the exact mode's fallback with true divisions.
template <typename T, int AM, bool SLOW>
__device__ __forceinline__ T upd_h2(const Coefs<T>& c, const Geom& g, const Fold& f, T h,
                                    T a1, T a0, double da_, Rcp ra, int ia,
                                    T b1, T b0, double db_, Rcp rb, int ib, unsigned& bad)
{
    if (sizeof(T) == 4)
        return fmaf(f.f[ib], (float)(b1 - b0), fmaf(-f.f[ia], (float)(a1 - a0), (float)c.uda * (float)h));
    if (AM == 1)
        return fma(f.d[ib], (double)(b1 - b0), fma(-f.d[ia], (double)(a1 - a0), (double)c.uda * (double)h));
    if (SLOW) return upd_h<T>(c.uda, h, c.udb, Ar<T>::diff_exact(a1, a0, da_, ra), Ar<T>::diff_exact(b1, b0, db_, rb));
    return upd_h<T>(c.uda, h, c.udb, Ar<T>::diff_fast(a1, a0, da_, ra, bad), Ar<T>::diff_fast(b1, b0, db_, rb, bad));
}
template <typename T, int AM, bool SLOW>
__device__ __forceinline__ T upd_e2(const Coefs<T>& c, const Geom& g, const Fold& f, T e,
                                    T a1, T a0, double da_, Rcp ra, int ia,
                                    T b1, T b0, double db_, Rcp rb, int ib, unsigned& bad)
{
    if (sizeof(T) == 4)
        return fmaf(-f.f[ib], (float)(b1 - b0), fmaf(f.f[ia], (float)(a1 - a0), (float)c.uca * (float)e));
    if (AM == 1)
        return fma(-f.d[ib], (double)(b1 - b0), fma(f.d[ia], (double)(a1 - a0), (double)c.uca * (double)e));
    if (SLOW) return upd_e<T>(c.uca, e, c.ucb, Ar<T>::diff_exact(a1, a0, da_, ra), Ar<T>::diff_exact(b1, b0, db_, rb));
    return upd_e<T>(c.uca, e, c.ucb, Ar<T>::diff_fast(a1, a0, da_, ra, bad), Ar<T>::diff_fast(b1, b0, db_, rb, bad));
}

// H stage: (ox,oy,oz) <- f(h, e (own, j+1: ez_j/ex_j, k+1: ey_n/ex_n), e of the next plane (ey_p, ez_p own)); plane gi.
// Shared by the one-step sweep (H+[i+1]) and the two-step sweeps (stages A and C).  Outputs must not alias inputs: the
// exact fp64 mode recomputes the stage with true divisions when a quotient left the fast sequence's range.
template <typename T, int V, bool MASKED, int AM, bool SLOW = false>
__device__ __forceinline__ void stage_h(const Coefs<T>& c, const Geom& g, const Fold& fo, int gi, bool jy1, bool jy2, int k,
                                        const Pack<T, V>& hx, const Pack<T, V>& hy, const Pack<T, V>& hz,
                                        const Pack<T, V>& ex, const Pack<T, V>& ey, const Pack<T, V>& ez,
                                        const Pack<T, V>& ez_j, const Pack<T, V>& ex_j, T ey_n, T ex_n,
                                        const Pack<T, V>& ey_p, const Pack<T, V>& ez_p,
                                        Pack<T, V>& ox, Pack<T, V>& oy, Pack<T, V>& oz)
{
    const bool ix1 = gi < g.nxg - 1, ix2 = gi < g.nxg - 2;
    unsigned bad = 0;
    ox = hx; oy = hy; oz = hz;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const bool kz1 = (k + e) < g.nz - 1, kz2 = (k + e) < g.nz - 2;
        const T ey_k = (e + 1 < V) ? ey.v[(e + 1) % V] : ey_n;
        const T ex_k = (e + 1 < V) ? ex.v[(e + 1) % V] : ex_n;
        T n = upd_h2<T, AM, SLOW>(c, g, fo, hx.v[e], ez_j.v[e], ez.v[e], g.dy, g.rdy, FHY, ey_k, ey.v[e], g.dz, g.rdz, FHZ, bad);
        if (!MASKED || (ix1 && jy2 && kz2)) ox.v[e] = n;
        n = upd_h2<T, AM, SLOW>(c, g, fo, hy.v[e], ex_k, ex.v[e], g.dz, g.rdz, FHZ, ez_p.v[e], ez.v[e], g.dx, g.rdx, FHX, bad);
        if (!MASKED || (ix2 && jy1 && kz2)) oy.v[e] = n;
        n = upd_h2<T, AM, SLOW>(c, g, fo, hz.v[e], ey_p.v[e], ey.v[e], g.dx, g.rdx, FHX, ex_j.v[e], ex.v[e], g.dy, g.rdy, FHY, bad);
        if (!MASKED || (ix2 && jy2 && kz1)) oz.v[e] = n;
    }
    if (sizeof(T) == 8 && AM == 0 && !SLOW) {
        if (bad) stage_h<T, V, MASKED, AM, true>(c, g, fo, gi, jy1, jy2, k, hx, hy, hz, ex, ey, ez, ez_j, ex_j, ey_n, ex_n, ey_p, ez_p, ox, oy, oz);
    }
}

// E stage: (ox,oy,oz) <- g(e, h (own, j+1: hz_j/hx_j, k+1: hy_n/hx_n), h of the next plane (hy_p, hz_p own)); plane gi
template <typename T, int V, bool MASKED, int AM, bool SLOW = false>
__device__ __forceinline__ void stage_e(const Coefs<T>& c, const Geom& g, const Fold& fo, int gi, bool jy1, int k,
                                        const Pack<T, V>& ex, const Pack<T, V>& ey, const Pack<T, V>& ez,
                                        const Pack<T, V>& hx, const Pack<T, V>& hy, const Pack<T, V>& hz,
                                        const Pack<T, V>& hz_j, const Pack<T, V>& hx_j, T hy_n, T hx_n,
                                        const Pack<T, V>& hy_p, const Pack<T, V>& hz_p,
                                        Pack<T, V>& ox, Pack<T, V>& oy, Pack<T, V>& oz)
{
    const bool ex0 = gi < g.nxg, ex1 = gi < g.nxg - 1;
    unsigned bad = 0;
    ox = ex; oy = ey; oz = ez;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const bool kz0 = (k + e) < g.nz, kz1 = (k + e) < g.nz - 1;
        const T hy_k = (e + 1 < V) ? hy.v[(e + 1) % V] : hy_n;
        const T hx_k = (e + 1 < V) ? hx.v[(e + 1) % V] : hx_n;
        T n = upd_e2<T, AM, SLOW>(c, g, fo, ex.v[e], hz_j.v[e], hz.v[e], g.dy, g.rdy, FEY, hy_k, hy.v[e], g.dz, g.rdz, FEZ, bad);
        if (!MASKED || (ex0 && jy1 && kz1)) ox.v[e] = n;
        n = upd_e2<T, AM, SLOW>(c, g, fo, ey.v[e], hx_k, hx.v[e], g.dz, g.rdz, FEZ, hz_p.v[e], hz.v[e], g.dx, g.rdx, FEX, bad);
        if (!MASKED || (ex1 && kz1)) oy.v[e] = n;
        n = upd_e2<T, AM, SLOW>(c, g, fo, ez.v[e], hy_p.v[e], hy.v[e], g.dx, g.rdx, FEX, hx_j.v[e], hx.v[e], g.dy, g.rdy, FEY, bad);
        if (!MASKED || (ex1 && jy1 && kz0)) oz.v[e] = n;
    }
    if (sizeof(T) == 8 && AM == 0 && !SLOW) {
        if (bad) stage_e<T, V, MASKED, AM, true>(c, g, fo, gi, jy1, k, ex, ey, ez, hx, hy, hz, hz_j, hx_j, hy_n, hx_n, hy_p, hz_p, ox, oy, oz);
    }
}

template <typename T> __device__ __forceinline__ Pack<T, VecOf<T>::V> zero_pack()
{
    Pack<T, VecOf<T>::V> r;
#pragma unroll
    for (int e = 0; e < VecOf<T>::V; ++e) r.v[e] = (T)0;
    return r;
}
template <typename T> __device__ __forceinline__ Pack<T, VecOf<T>::V> ldv_if(const T* p, bool ok)
{
    return ok ? ldv<T>(p) : zero_pack<T>();
}
// streaming variant: ld.global.cg (L2 only) — every byte is used once per CTA, L1 allocation buys nothing
template <typename T, int POL> __device__ __forceinline__ Pack<T, VecOf<T>::V> ldv_pol(const T* p, bool ok)
{
    typedef typename VecOf<T>::type VT;
    if (!ok) return zero_pack<T>();
    if (POL & 2) {
        union { VT q; Pack<T, VecOf<T>::V> r; } u;
        u.q = __ldcg(reinterpret_cast<const VT*>(p));
        return u.r;
    }
    return ldv<T>(p);
}
template <typename T> __device__ __forceinline__ T shfl_next(T v)
{
    return __shfl_down_sync(0xffffffffu, v, 1);
}

// store with an L2 evict-first hint: the output set is not read again until the next step (>> L2 away),
// so it should not push the rim lines that neighbouring CTAs are about to re-read out of L2
template <typename T, int POL> __device__ __forceinline__ void stv_pol(T* p, const Pack<T, VecOf<T>::V>& r)
{
    typedef typename VecOf<T>::type VT;
    union { VT q; Pack<T, VecOf<T>::V> r; } u;
    u.r = r;
    if (POL & 1) __stcs(reinterpret_cast<VT*>(p), u.q);
    else *reinterpret_cast<VT*>(p) = u.q;
}

// TJ owner rows per CTA; blockDim = (32, TJ + 1);  AM = fp64 arithmetic mode (0 exact, 1 folded).
// (Rejected by measurement, profiles/r01_tuning.md, and removed: evict-first stores, ld.global.cg loads, 2 CTAs x 8 warps.)
// ADE: apply the dispersive-medium recursions of the PREVIOUS step on the E stage's input values (ade_in_sweep).
template <typename T, int TJ, int AM, bool ADE>
__device__ __forceinline__ void
fused_sweep(const CFields<T>& in, const Fields<T>& out, const Coefs<T>& c, const Geom& g, const FusedTiling& t, const Fold& fo,
            const AdeIn& ad, const int item)
{
    constexpr int POL = 0;
    constexpr int V = VecOf<T>::V;
    typedef Pack<T, V> P;
    typedef typename VecOf<T>::type VT;
    extern __shared__ __align__(16) unsigned char smem_[];
    // [parity][row][Ez, Ex of plane i+1][lane]  and  [parity][row][Hz+, Hx+ of plane i][lane]
    VT (*s_raw)[TJ + 1][2][32] = reinterpret_cast<VT (*)[TJ + 1][2][32]>(smem_);
    VT (*s_h)[TJ + 1][2][32] = reinterpret_cast<VT (*)[TJ + 1][2][32]>(smem_ + 2 * (TJ + 1) * 2 * 32 * sizeof(VT));

    const int lane = threadIdx.x, row = threadIdx.y;
    const int ntiles = t.ntj * t.ntk;
    const int seg = item / ntiles, tile = item - seg * ntiles;
    const int tj = tile / t.ntk, tk = tile - tj * t.ntk;
    const int j = tj * TJ + row;
    const int k = (tk * t.own_lanes + lane) * V;
    const int i0 = t.i_begin + seg * t.lx;
    const int i1 = min(i0 + t.lx, t.i_end);
    if (t.halo_flag && i1 + 1 >= g.nx) {          // this segment reads E up to plane i1+1: ghost planes start at nx
        if (threadIdx.x == 0 && threadIdx.y == 0) wait_flag_ge(t.halo_flag, t.halo_need, t.error_word, t.timeout_ns);
        __syncthreads();
    }
    const bool ld_ok = (j < g.ny) && (k < g.pz);
    const bool owner = ld_ok && row < TJ && lane < t.own_lanes;
    const bool halo_row = row == TJ;
    const long long o = (long long)j * g.sy + k;

    const bool jy1 = j < g.ny - 1, jy2 = j < g.ny - 2;
    bool kz0[V], kz1[V], kz2[V];
#pragma unroll
    for (int e = 0; e < V; ++e) { kz0[e] = (k + e) < g.nz; kz1[e] = (k + e) < g.nz - 1; kz2[e] = (k + e) < g.nz - 2; }

    // window: E[i], E[i+1], H+[i]  (i = i0 - 1 at entry; E[i0-1] and H+[i0-1] are never used)
    P e0x = zero_pack<T>(), e0y = e0x, e0z = e0x, hpx = e0x, hpy = e0x, hpz = e0x;
    const T* pex = in.ex + o; const T* pey = in.ey + o; const T* pez = in.ez + o;
    const T* phx = in.hx + o; const T* phy = in.hy + o; const T* phz = in.hz + o;
    long long po = (long long)i0 * g.sx;                 // plane offset of E[i+1] for i = i0-1
    P e1x = ldv_pol<T, POL>(pex + po, ld_ok), e1y = ldv_pol<T, POL>(pey + po, ld_ok), e1z = ldv_pol<T, POL>(pez + po, ld_ok);
    // prefetched: E[i+2], H[i+1]
    P e2x = ldv_pol<T, POL>(pex + po + g.sx, ld_ok), e2y = ldv_pol<T, POL>(pey + po + g.sx, ld_ok),
      e2z = ldv_pol<T, POL>(pez + po + g.sx, ld_ok);
    P h1x = ldv_pol<T, POL>(phx + po, ld_ok), h1y = ldv_pol<T, POL>(phy + po, ld_ok), h1z = ldv_pol<T, POL>(phz + po, ld_ok);

    unsigned ade_mask = 0;
    __shared__ AdeOp s_ade[ADE ? kAdeSmemOps : 1];
    if (ADE) {
        ade_stage_ops(s_ade, ad, threadIdx.y * 32 + threadIdx.x, 32 * (TJ + 1));
        if (owner) ade_mask = ade_thread_mask<V>(ad, i0, i1, j, k);
        __syncthreads();
    }

    for (int i = i0 - 1; i < i1; ++i) {
        const int par = (i - i0 + 1) & 1;
        po = (long long)(i + 1) * g.sx;                  // plane i+1
        // ---- issue next iteration's loads: E[i+3], H[i+2] -----------------------------------------
        const bool more = (i + 1 < i1) && ld_ok;
        const P n_ex = ldv_pol<T, POL>(pex + po + 2 * g.sx, more), n_ey = ldv_pol<T, POL>(pey + po + 2 * g.sx, more),
                n_ez = ldv_pol<T, POL>(pez + po + 2 * g.sx, more);
        const P n_hx = ldv_pol<T, POL>(phx + po + g.sx, more), n_hy = ldv_pol<T, POL>(phy + po + g.sx, more),
                n_hz = ldv_pol<T, POL>(phz + po + g.sx, more);
        // ---- publish what the row below (j-1) needs from us -------------------------------------------
        {
            union { VT q; P r; } u;
            u.r = e1z; s_raw[par][row][0][lane] = u.q;
            u.r = e1x; s_raw[par][row][1][lane] = u.q;
            u.r = hpz; s_h[par][row][0][lane] = u.q;
            u.r = hpx; s_h[par][row][1][lane] = u.q;
        }
        __syncthreads();
        P ez_j, ex_j, hz_j, hx_j;
        if (!halo_row) {
            union { VT q; P r; } u;
            u.q = s_raw[par][row + 1][0][lane]; ez_j = u.r;
            u.q = s_raw[par][row + 1][1][lane]; ex_j = u.r;
            u.q = s_h[par][row + 1][0][lane]; hz_j = u.r;
            u.q = s_h[par][row + 1][1][lane]; hx_j = u.r;
        } else {
            const bool ok = ld_ok && (j + 1 < g.ny);
            ez_j = ldv_pol<T, POL>(pez + po + g.sy, ok);
            ex_j = ldv_pol<T, POL>(pex + po + g.sy, ok);
            hz_j = zero_pack<T>(); hx_j = hz_j;          // the rim row produces no E+
        }
        // ---- k+1 neighbours from the next lane -----------------------------------------------------------
        const T ey1_n = shfl_next<T>(e1y.v[0]), ex1_n = shfl_next<T>(e1x.v[0]);
        const T hpy_n = shfl_next<T>(hpy.v[0]), hpx_n = shfl_next<T>(hpx.v[0]);

        // ---- H+[i+1] ---------------------------------------------------------------------------------------
        P hnx, hny, hnz;
        stage_h<T, V, true, AM>(c, g, fo, g.x0 + i + 1, jy1, jy2, k, h1x, h1y, h1z, e1x, e1y, e1z, ez_j, ex_j, ey1_n, ex1_n,
                                e2y, e2z, hnx, hny, hnz);
        if (owner && i + 1 < i1) {
            stv_pol<T, POL>(out.hx + o + po, hnx); stv_pol<T, POL>(out.hy + o + po, hny); stv_pol<T, POL>(out.hz + o + po, hnz);
        }
        // ---- E+[i] ---------------------------------------------------------------------------------------------
        if (i >= i0) {
            P nx_, ny_, nz_;
            stage_e<T, V, true, AM>(c, g, fo, g.x0 + i, jy1, k, e0x, e0y, e0z, hpx, hpy, hpz, hz_j, hx_j, hpy_n, hpx_n,
                                    hny, hnz, nx_, ny_, nz_);
            if (ADE) {
                if (ade_mask) {
                    double jx[V], jy[V], jz[V];
#pragma unroll
                    for (int e = 0; e < V; ++e) jx[e] = jy[e] = jz[e] = 0.0;
                    ade_in_sweep<T, V>(ad, s_ade, ade_mask, i, j, k, e0x, e0y, e0z, jx, jy, jz);
                    if (ad.coupled) {
#pragma unroll
                        for (int e = 0; e < V; ++e) {                  // op boxes lie inside the updated range of their component
                            if (jx[e] != 0.0) nx_.v[e] = Ar<T>::sub(nx_.v[e], Ar<T>::mul(c.ucb, (T)jx[e]));
                            if (jy[e] != 0.0) ny_.v[e] = Ar<T>::sub(ny_.v[e], Ar<T>::mul(c.ucb, (T)jy[e]));
                            if (jz[e] != 0.0) nz_.v[e] = Ar<T>::sub(nz_.v[e], Ar<T>::mul(c.ucb, (T)jz[e]));
                        }
                    }
                }
            }
            if (owner) {
                const long long pe = (long long)i * g.sx;
                stv_pol<T, POL>(out.ex + o + pe, nx_); stv_pol<T, POL>(out.ey + o + pe, ny_); stv_pol<T, POL>(out.ez + o + pe, nz_);
            }
        }
        // ---- rotate the window --------------------------------------------------------------------------------------
        e0x = e1x; e0y = e1y; e0z = e1z;
        e1x = e2x; e1y = e2y; e1z = e2z;
        e2x = n_ex; e2y = n_ey; e2z = n_ez;
        hpx = hnx; hpy = hny; hpz = hnz;
        h1x = n_hx; h1y = n_hy; h1z = n_hz;
    }
}

template <typename T, int TJ, int AM, bool ADE = false>
__global__ void __launch_bounds__(32 * (TJ + 1), 1)
k_fused3d(const __grid_constant__ CFields<T> in, const __grid_constant__ Fields<T> out, const __grid_constant__ Coefs<T> c,
          const __grid_constant__ Geom g, const __grid_constant__ FusedTiling t, const __grid_constant__ Fold fo,
          const __grid_constant__ AdeIn ad)
{
    if (ADE) {
        constexpr int V = VecOf<T>::V;
        const int item = ad.order ? ad.order[blockIdx.x] : (int)blockIdx.x;
        const int ntiles = t.ntj * t.ntk;
        const int seg = item / ntiles, tile = item - seg * ntiles;
        const int tj = tile / t.ntk, tk = tile - tj * t.ntk;
        const int i0 = t.i_begin + seg * t.lx, i1 = min(i0 + t.lx, t.i_end);
        if (ade_tile_touched(ad, i0, i1, tj * TJ, tj * TJ + TJ, tk * t.own_lanes * V, (tk + 1) * t.own_lanes * V)) {
            fused_sweep<T, TJ, AM, true>(in, out, c, g, t, fo, ad, item);
            return;
        }
        fused_sweep<T, TJ, AM, false>(in, out, c, g, t, fo, ad, item);
        return;
    }
    fused_sweep<T, TJ, AM, false>(in, out, c, g, t, fo, ad, (int)blockIdx.x);
}

}  // namespace fdtd
