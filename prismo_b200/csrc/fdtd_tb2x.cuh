// fdtd_tb2x.cuh — the two-step temporally blocked sweep (fdtd_tb2.cuh) re-plumbed for Blackwell:
//
//   * INPUT by TMA.  One producer thread issues cp.async.bulk.tensor.3d loads of the CTA's (j,k) tile — rim rows and
//     rim lanes included, out-of-range rows / columns / planes zero-filled by the TMA unit — for E0[p] (x,y,z) and
//     H0[p-1] (x,y,z) into an S-stage shared-memory ring guarded by full / empty mbarriers.  The consumer warps no
//     longer carry prefetch registers, load predicates or global addresses for the inputs, and the j+1 neighbours of
//     E0 come straight out of the staged tile (row + 1) instead of a publish-and-barrier round trip.
//   * NEIGHBOUR-ONLY SYNCHRONISATION.  Row r needs the z,x components of H1, E1, H2 of row r+1 (its +j neighbours) and
//     nothing else from the CTA: every row publishes them into a D-slot ring and arrives on ITS OWN mbarrier; row r
//     waits on row r+1's barrier (data) and on row r-1's (back-pressure before a slot is overwritten).  There is no
//     CTA-wide barrier in the plane loop, so the warps drift apart and the load / shared-memory / FP phases of
//     different rows overlap (the per-plane __syncthreads of fdtd_tb2.cuh held 2.3 of 4 warps per scheduler at the
//     barrier: ncu, profiles/r01_ncu_summary.md).
//   * NO REGISTER ROTATION in the steady state: the plane loop is unrolled three times with the three-deep windows
//     (E0, H1, E1) renamed instead of moved.
//
// The arithmetic, the window contents, the masks, the mid-step sources / monitors and therefore every result bit are
// those of fdtd_tb2.cuh (same stage_h / stage_e / mid_* functions); tests compare the two kernels bitwise.
//
// Pipeline at iteration it (plane i = i0 - 3 + it), TMA stage q holds E0[i0 + q] and H0[i0 + q - 1]:
//     A: H1[i+3] = f(H0[i+3] (stage it+1), E0[i+3] (registers; +j from stage it, +k by shuffle), E0[i+4] (stage it+1))
//     B: E1[i+2] = g(E0[i+2], H1[i+2], H1[i+3])      C: H2[i+1] = f(H1[i+1], E1[i+1], E1[i+2])
//     D: E2[i]   = g(E1[i],   H2[i],   H2[i+1])      -> st.global E2[i], H2[i+1]
#pragma once
#include <cuda.h>

#include "fdtd_tb2.cuh"

namespace fdtd {

struct Tb2xMaps { CUtensorMap m[6]; };      // Ex Ey Ez Hx Hy Hz of the INPUT set: dims (pz, ny, planes), box (256 B, R, 1)

// ---- mbarrier / TMA primitives (PTX ISA 8.x, sm_90+; SASS: SYNCS.*, UTMALDG) -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

constexpr int kTb2xRows = 15;            // consumer warp rows per CTA: 11 owners + 4 rim; a 16th warp produces (17 warps would cap
                                         // registers at 96 per thread: five warps on one scheduler share its 16 K registers)
constexpr int kTb2xRowBytes = 256;       // 32 lanes x 8 B: one row of one tile / one quantity of one exchange slot

// shared memory: [S stages][6 arrays][R rows][256 B]  |  [D slots][R rows][6 quantities][256 B]  |  barriers
template <int R> __host__ __device__ constexpr size_t tb2x_stage_bytes() { return (size_t)6 * R * kTb2xRowBytes; }
template <int R> __host__ __device__ constexpr size_t tb2x_slot_bytes() { return (size_t)R * 6 * kTb2xRowBytes; }
template <int R> static inline size_t tb2x_smem_bytes(int S, int D)
{
    return 128 + (size_t)S * tb2x_stage_bytes<R>() + (size_t)D * tb2x_slot_bytes<R>() + (size_t)(2 * S + R * D) * 8 + 64;
}

// everything a consumer thread keeps across iterations besides the window
template <typename T> struct Tb2xCtx {
    uint32_t tiles, xch, full, empty, xfull;     // shared-memory addresses (bytes)
    uint32_t own_off;                            // row * 256 + lane * 8
    uint32_t up_off;                             // rown * 256 + lane * 8
    int S, D, row, lane, j, k, i0, i1, step_row;
    bool owner, jy1, jy2, interior;
    bool arrive;                                 // this lane arrives on the barriers (lane 0; every lane in the sanitizer mode)
    unsigned ofs;                                // j * sy + k
};

template <typename T, int V> __device__ __forceinline__ Pack<T, V> lds8(uint32_t addr)
{
    typedef typename Vec8<T>::type VT;
    union { VT q; Pack<T, V> r; } u;
    if (sizeof(VT) == 8) {
        unsigned long long w;
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(w) : "r"(addr));
        union { unsigned long long w; VT q; } c;
        c.w = w;
        u.q = c.q;
    }
    return u.r;
}
template <typename T, int V> __device__ __forceinline__ void sts8(uint32_t addr, const Pack<T, V>& r)
{
    typedef typename Vec8<T>::type VT;
    union { VT q; Pack<T, V> r; unsigned long long w; } u;
    u.r = r;
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(u.w) : "memory");
}

// One plane iteration of one consumer thread.  Window arguments are passed in ROTATED order by the unrolled caller:
//   e0a = E0[i+2], e0b = E0[i+3], e0n <- E0[i+4];   h1a = H1[i+1], h1b = H1[i+2], h1c <- H1[i+3];
//   e1a = E1[i],   e1b = E1[i+1], e1c <- E1[i+2];   h2a = H2[i],   h2b <- H2[i+1]
template <typename T, int R, int AM, bool OPS, bool MASKED, bool STEADY>
__device__ __forceinline__ void
tb2x_iter(const Tb2xCtx<T>& cx, const int it, const Fields<T>& out, const Coefs<T>& c, const Geom& g, const MidOps& m,
          const Fold& fo,
          Pack<T, Vec8<T>::V>& e0ax, Pack<T, Vec8<T>::V>& e0ay, Pack<T, Vec8<T>::V>& e0az,
          Pack<T, Vec8<T>::V>& e0bx, Pack<T, Vec8<T>::V>& e0by, Pack<T, Vec8<T>::V>& e0bz,
          Pack<T, Vec8<T>::V>& e0nx, Pack<T, Vec8<T>::V>& e0ny, Pack<T, Vec8<T>::V>& e0nz,
          Pack<T, Vec8<T>::V>& h1ax, Pack<T, Vec8<T>::V>& h1ay, Pack<T, Vec8<T>::V>& h1az,
          Pack<T, Vec8<T>::V>& h1bx, Pack<T, Vec8<T>::V>& h1by, Pack<T, Vec8<T>::V>& h1bz,
          Pack<T, Vec8<T>::V>& h1cx, Pack<T, Vec8<T>::V>& h1cy, Pack<T, Vec8<T>::V>& h1cz,
          Pack<T, Vec8<T>::V>& e1ax, Pack<T, Vec8<T>::V>& e1ay, Pack<T, Vec8<T>::V>& e1az,
          Pack<T, Vec8<T>::V>& e1bx, Pack<T, Vec8<T>::V>& e1by, Pack<T, Vec8<T>::V>& e1bz,
          Pack<T, Vec8<T>::V>& e1cx, Pack<T, Vec8<T>::V>& e1cy, Pack<T, Vec8<T>::V>& e1cz,
          Pack<T, Vec8<T>::V>& h2ax, Pack<T, Vec8<T>::V>& h2ay, Pack<T, Vec8<T>::V>& h2az,
          Pack<T, Vec8<T>::V>& h2bx, Pack<T, Vec8<T>::V>& h2by, Pack<T, Vec8<T>::V>& h2bz,
          int& sq, int& sph, int& xd, int& xph)
{
    constexpr int V = Vec8<T>::V;
    typedef Pack<T, V> P;
    constexpr uint32_t ROWS = R * kTb2xRowBytes;         // bytes of one array tile
    const int i = cx.i0 - 3 + it;
    const int k = cx.k, j = cx.j;
    (void)j;

    // ---- +j neighbours of H1[i+2], E1[i+1], H2[i] (z, x components): slot `xd` of the row above -------------------------
    const uint32_t xrow = cx.xch + (uint32_t)xd * (uint32_t)tb2x_slot_bytes<R>() + cx.up_off * 6u - (uint32_t)(cx.lane * 8) * 5u;
    // (row stride inside a slot is 6 x 256 B; up_off = rown * 256 + lane * 8)
    mbar_wait(cx.xfull + (uint32_t)(((cx.up_off >> 8) * cx.D + xd) * 8), (uint32_t)xph);
    const P h1z_j = lds8<T, V>(xrow + 0 * kTb2xRowBytes), h1x_j = lds8<T, V>(xrow + 1 * kTb2xRowBytes);
    const P e1z_j = lds8<T, V>(xrow + 2 * kTb2xRowBytes), e1x_j = lds8<T, V>(xrow + 3 * kTb2xRowBytes);
    const P h2z_j = lds8<T, V>(xrow + 4 * kTb2xRowBytes), h2x_j = lds8<T, V>(xrow + 5 * kTb2xRowBytes);

    // ---- TMA stages: q0 = it (E0[i+3]: +j rows), q1 = it + 1 (E0[i+4], H0[i+3] own cells) ------------------------------------
    const uint32_t st0 = cx.tiles + (uint32_t)sq * (uint32_t)tb2x_stage_bytes<R>();
    int sq1 = sq + 1, sph1 = sph;
    if (sq1 == cx.S) { sq1 = 0; sph1 ^= 1; }
    const uint32_t st1 = cx.tiles + (uint32_t)sq1 * (uint32_t)tb2x_stage_bytes<R>();
    mbar_wait(cx.full + (uint32_t)sq1 * 8u, (uint32_t)sph1);
    e0nx = lds8<T, V>(st1 + 0 * ROWS + cx.own_off);
    e0ny = lds8<T, V>(st1 + 1 * ROWS + cx.own_off);
    e0nz = lds8<T, V>(st1 + 2 * ROWS + cx.own_off);
    const P nh0x = lds8<T, V>(st1 + 3 * ROWS + cx.own_off), nh0y = lds8<T, V>(st1 + 4 * ROWS + cx.own_off),
            nh0z = lds8<T, V>(st1 + 5 * ROWS + cx.own_off);
    const P e0z_j = lds8<T, V>(st0 + 2 * ROWS + cx.up_off), e0x_j = lds8<T, V>(st0 + 0 * ROWS + cx.up_off);
    // stage `it` of the input ring is no longer needed by this warp: hand it back to the producer right away (its +j
    // rows were the last thing read from it), so the ring prefetches S - 2 planes beyond the one being consumed
    __syncwarp();
    if (cx.arrive) mbar_arrive(cx.empty + (uint32_t)sq * 8u);

    // ---- intermediate-step H sources / monitors on H1[i+1] (all of its pre-source uses are done) ---------------------------
    if (OPS && (unsigned)(i + 1 - m.op_lo) <= (unsigned)m.op_span && i + 1 >= cx.i0) {
        const unsigned char fl = m.plane_flags[i + 1];
        if (fl & 1) {
            mid_sources<T, V>(m, 3, i + 1, cx.j, k, cx.step_row, h1ax);
            mid_sources<T, V>(m, 4, i + 1, cx.j, k, cx.step_row, h1ay);
            mid_sources<T, V>(m, 5, i + 1, cx.j, k, cx.step_row, h1az);
        }
        if ((fl & 2) && cx.owner && i + 1 < cx.i1) {
            mid_monitors<T, V>(m, 3, i + 1, cx.j, k, cx.step_row, h1ax);
            mid_monitors<T, V>(m, 4, i + 1, cx.j, k, cx.step_row, h1ay);
            mid_monitors<T, V>(m, 5, i + 1, cx.j, k, cx.step_row, h1az);
        }
    }
    // ---- +k neighbours from the next lane -----------------------------------------------------------------------------------------
    const T e0y_n = shfl_next<T>(e0by.v[0]), e0x_n = shfl_next<T>(e0bx.v[0]);
    const T h1y_n = shfl_next<T>(h1by.v[0]), h1x_n = shfl_next<T>(h1bx.v[0]);
    const T e1y_n = shfl_next<T>(e1by.v[0]), e1x_n = shfl_next<T>(e1bx.v[0]);
    const T h2y_n = shfl_next<T>(h2ay.v[0]), h2x_n = shfl_next<T>(h2ax.v[0]);

    // ---- A: H1[i+3], B: E1[i+2] (+ intermediate-step E sources / monitors), C: H2[i+1], D: E2[i] --------------------------------
    stage_h<T, V, MASKED, AM>(c, g, fo, g.x0 + i + 3, cx.jy1, cx.jy2, k, nh0x, nh0y, nh0z, e0bx, e0by, e0bz, e0z_j, e0x_j, e0y_n,
                          e0x_n, e0ny, e0nz, h1cx, h1cy, h1cz);
    stage_e<T, V, MASKED, AM>(c, g, fo, g.x0 + i + 2, cx.jy1, k, e0ax, e0ay, e0az, h1bx, h1by, h1bz, h1z_j, h1x_j, h1y_n,
                          h1x_n, h1cy, h1cz, e1cx, e1cy, e1cz);
    if (OPS && (unsigned)(i + 2 - m.op_lo) <= (unsigned)m.op_span && (STEADY || i + 2 >= cx.i0)) {
        const unsigned char fl = m.plane_flags[i + 2];
        if (fl & 1) {
            mid_sources<T, V>(m, 0, i + 2, cx.j, k, cx.step_row, e1cx);
            mid_sources<T, V>(m, 1, i + 2, cx.j, k, cx.step_row, e1cy);
            mid_sources<T, V>(m, 2, i + 2, cx.j, k, cx.step_row, e1cz);
        }
        if ((fl & 2) && cx.owner && (STEADY || i + 2 < cx.i1)) {
            mid_monitors<T, V>(m, 0, i + 2, cx.j, k, cx.step_row, e1cx);
            mid_monitors<T, V>(m, 1, i + 2, cx.j, k, cx.step_row, e1cy);
            mid_monitors<T, V>(m, 2, i + 2, cx.j, k, cx.step_row, e1cz);
        }
    }
    stage_h<T, V, MASKED, AM>(c, g, fo, g.x0 + i + 1, cx.jy1, cx.jy2, k, h1ax, h1ay, h1az, e1bx, e1by, e1bz, e1z_j, e1x_j, e1y_n,
                          e1x_n, e1cy, e1cz, h2bx, h2by, h2bz);

    // ---- publish the next iteration's +j inputs (H1[i+3], E1[i+2], H2[i+1]) as soon as they exist ------------------------------
    int xd1 = xd + 1, xph1 = xph;
    if (xd1 == cx.D) { xd1 = 0; xph1 ^= 1; }
    if (STEADY || it + 1 < cx.i1 - cx.i0 + 3) {
        // back-pressure: slot xd1 was last read by the row below during its iteration it+1-D; it has published
        // iteration it+2-D since then (slot (it+2) % D, phase (it+2)/D - 1)
        if (cx.row > 0 && it + 1 >= cx.D) {
            int bd = xd1 + 1, bph = xph1 ^ 1;
            if (bd == cx.D) { bd = 0; bph ^= 1; }
            mbar_wait(cx.xfull + (uint32_t)(((cx.row - 1) * cx.D + bd) * 8), (uint32_t)bph);
        }
        const uint32_t xme = cx.xch + (uint32_t)xd1 * (uint32_t)tb2x_slot_bytes<R>() + cx.own_off * 6u - (uint32_t)(cx.lane * 8) * 5u;
        sts8<T, V>(xme + 0 * kTb2xRowBytes, h1cz); sts8<T, V>(xme + 1 * kTb2xRowBytes, h1cx);
        sts8<T, V>(xme + 2 * kTb2xRowBytes, e1cz); sts8<T, V>(xme + 3 * kTb2xRowBytes, e1cx);
        sts8<T, V>(xme + 4 * kTb2xRowBytes, h2bz); sts8<T, V>(xme + 5 * kTb2xRowBytes, h2bx);
        __syncwarp();
        if (cx.arrive) mbar_arrive(cx.xfull + (uint32_t)((cx.row * cx.D + xd1) * 8));
    }

    const unsigned ost = cx.ofs + (unsigned)i * (unsigned)g.sx;            // plane i   (used only when i   >= i0 >= 0)
    const unsigned ost1 = cx.ofs + (unsigned)(i + 1) * (unsigned)g.sx;     // plane i+1 (used only when i+1 >= i0 >= 0)
    if (cx.owner && (STEADY || (i + 1 >= cx.i0 && i + 1 < cx.i1))) {
        st8<T, V>(out.hx + ost1, h2bx); st8<T, V>(out.hy + ost1, h2by); st8<T, V>(out.hz + ost1, h2bz);
    }
    if (STEADY || i >= cx.i0) {
        P e2x, e2y, e2z;
        stage_e<T, V, MASKED, AM>(c, g, fo, g.x0 + i, cx.jy1, k, e1ax, e1ay, e1az, h2ax, h2ay, h2az, h2z_j, h2x_j, h2y_n,
                              h2x_n, h2by, h2bz, e2x, e2y, e2z);
        if (cx.owner) { st8<T, V>(out.ex + ost, e2x); st8<T, V>(out.ey + ost, e2y); st8<T, V>(out.ez + ost, e2z); }
    }
    sq = sq1; sph = sph1; xd = xd1; xph = xph1;
}

template <typename T, int R, bool OPS, int AM>
__device__ __forceinline__ void
tb2x_sweep(const Tb2xMaps& maps, const Fields<T>& out, const Coefs<T>& c, const Geom& g, const FusedTiling& t,
           const MidOps& m, const Fold& fo, const int S, const int D, const int all_arrive)
{
    constexpr int V = Vec8<T>::V;
    typedef Pack<T, V> P;
    extern __shared__ __align__(128) unsigned char smem_raw_[];
    const uint32_t base = (smem_u32(smem_raw_) + 127u) & ~127u;
    const uint32_t tiles = base;
    const uint32_t xch = tiles + (uint32_t)S * (uint32_t)tb2x_stage_bytes<R>();
    const uint32_t full = xch + (uint32_t)D * (uint32_t)tb2x_slot_bytes<R>();
    const uint32_t empty = full + (uint32_t)S * 8u;
    const uint32_t xfull = empty + (uint32_t)S * 8u;

    const int lane = threadIdx.x, row = threadIdx.y;
    const int ntiles = t.ntj * t.ntk;
    const int slot = blockIdx.x / ntiles, tile = blockIdx.x - slot * ntiles;
    const int tj = tile / t.ntk, tk = tile - tj * t.ntk;
    const int j0 = tj * (R - 4), k0 = tk * t.own_lanes * V;
    const int i0 = t.seg_lo[slot], i1 = t.seg_hi[slot];
    const int n_it = i1 - i0 + 3;

    if (t.halo_flag && i1 + 3 >= g.nx) {          // this segment reads E0 up to plane i1+3: ghost planes start at nx
        if (lane == 0 && row == 0) wait_flag_ge(t.halo_flag, t.halo_need, t.error_word, t.timeout_ns);
    }
    if (row == R && lane == 0) {
        // arrivals per phase: one elected lane per warp after __syncwarp(), or (all_arrive: the mode compute-sanitizer's
        // racecheck can follow) every lane for itself
        const uint32_t per_warp = all_arrive ? 32u : 1u;
        for (int s = 0; s < S; ++s) { mbar_init(full + 8u * s, 1); mbar_init(empty + 8u * s, R * per_warp); }
        for (int q = 0; q < R * D; ++q) mbar_init(xfull + 8u * q, per_warp);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (row == R) {
        // ---- producer: stage q <- E0[i0 + q] (x,y,z) and, for q > 0, H0[i0 + q - 1] (x,y,z) ------------------------------------
        if (lane == 0) {
            int s = 0, ph = 0;
            for (int q = 0; q <= n_it; ++q) {
                mbar_wait(empty + 8u * s, (uint32_t)(ph ^ 1));
                const uint32_t dst = tiles + (uint32_t)s * (uint32_t)tb2x_stage_bytes<R>();
                const uint32_t bar = full + 8u * s;
                mbar_arrive_expect_tx(bar, (uint32_t)((q > 0 ? 6 : 3) * R * kTb2xRowBytes));
#pragma unroll
                for (int a = 0; a < 3; ++a) tma_load_3d(dst + a * R * kTb2xRowBytes, &maps.m[a], k0, j0, i0 + q, bar);
                if (q > 0) {
#pragma unroll
                    for (int a = 3; a < 6; ++a) tma_load_3d(dst + a * R * kTb2xRowBytes, &maps.m[a], k0, j0, i0 + q - 1, bar);
                }
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
        return;
    }

    // ---- consumers ---------------------------------------------------------------------------------------------------------------------
    Tb2xCtx<T> cx;
    cx.tiles = tiles; cx.xch = xch; cx.full = full; cx.empty = empty; cx.xfull = xfull;
    cx.S = S; cx.D = D; cx.row = row; cx.lane = lane; cx.i0 = i0; cx.i1 = i1;
    cx.arrive = all_arrive || lane == 0;
    cx.j = j0 + row; cx.k = k0 + lane * V;
    const int rown = min(row + 1, R - 1);
    cx.own_off = (uint32_t)(row * kTb2xRowBytes + lane * 8);
    cx.up_off = (uint32_t)(rown * kTb2xRowBytes + lane * 8);
    const bool ld_ok = (cx.j < g.ny) && (cx.k < g.pz);
    cx.owner = ld_ok && row < R - 4 && lane < t.own_lanes;
    cx.ofs = (unsigned)cx.j * (unsigned)g.sy + (unsigned)cx.k;
    cx.jy1 = cx.j < g.ny - 1; cx.jy2 = cx.j < g.ny - 2;
    cx.step_row = m.step_ptr ? (*m.step_ptr + m.step_off) : 0;
    cx.interior = (tj * (R - 4) + R - 1 < g.ny - 2) && ((tk * t.own_lanes + 31) * V + V - 1 < g.nz - 2);

    P z_;
#pragma unroll
    for (int e = 0; e < V; ++e) z_.v[e] = (T)0;
    // publish number 0: the +j inputs of iteration 0 are all zero (planes below i0 carry nothing)
    {
        const uint32_t xme = xch + cx.own_off * 6u - (uint32_t)(lane * 8) * 5u;
#pragma unroll
        for (int qn = 0; qn < 6; ++qn) sts8<T, V>(xme + qn * kTb2xRowBytes, z_);
        __syncwarp();
        if (cx.arrive) mbar_arrive(xfull + (uint32_t)((row * D) * 8));
    }
    // window at it = 0 (i = i0 - 3): E0[i+2] = 0 (never used), E0[i+3] = stage 0 own cells
    P e0ax = z_, e0ay = z_, e0az = z_, e0cx = z_, e0cy = z_, e0cz = z_;
    mbar_wait(full, 0);
    P e0bx = lds8<T, V>(tiles + 0 * R * kTb2xRowBytes + cx.own_off), e0by = lds8<T, V>(tiles + 1 * R * kTb2xRowBytes + cx.own_off),
      e0bz = lds8<T, V>(tiles + 2 * R * kTb2xRowBytes + cx.own_off);
    P h1ax = z_, h1ay = z_, h1az = z_, h1bx = z_, h1by = z_, h1bz = z_, h1cx = z_, h1cy = z_, h1cz = z_;
    P e1ax = z_, e1ay = z_, e1az = z_, e1bx = z_, e1by = z_, e1bz = z_, e1cx = z_, e1cy = z_, e1cz = z_;
    P h2ax = z_, h2ay = z_, h2az = z_, h2bx = z_, h2by = z_, h2bz = z_;
    int sq = 0, sph = 0, xd = 0, xph = 0;

#define TB2X_ARGS_0 e0ax, e0ay, e0az, e0bx, e0by, e0bz, e0cx, e0cy, e0cz, h1ax, h1ay, h1az, h1bx, h1by, h1bz, h1cx, h1cy, h1cz, \
                    e1ax, e1ay, e1az, e1bx, e1by, e1bz, e1cx, e1cy, e1cz
#define TB2X_ARGS_1 e0bx, e0by, e0bz, e0cx, e0cy, e0cz, e0ax, e0ay, e0az, h1bx, h1by, h1bz, h1cx, h1cy, h1cz, h1ax, h1ay, h1az, \
                    e1bx, e1by, e1bz, e1cx, e1cy, e1cz, e1ax, e1ay, e1az
#define TB2X_ARGS_2 e0cx, e0cy, e0cz, e0ax, e0ay, e0az, e0bx, e0by, e0bz, h1cx, h1cy, h1cz, h1ax, h1ay, h1az, h1bx, h1by, h1bz, \
                    e1cx, e1cy, e1cz, e1ax, e1ay, e1az, e1bx, e1by, e1bz
#define TB2X_H2_0 h2ax, h2ay, h2az, h2bx, h2by, h2bz
#define TB2X_H2_1 h2bx, h2by, h2bz, h2ax, h2ay, h2az
#define TB2X_ROTATE()                                                                                                   \
    do {                                                                                                                \
        e0ax = e0bx; e0ay = e0by; e0az = e0bz; e0bx = e0cx; e0by = e0cy; e0bz = e0cz;                                   \
        h1ax = h1bx; h1ay = h1by; h1az = h1bz; h1bx = h1cx; h1by = h1cy; h1bz = h1cz;                                   \
        e1ax = e1bx; e1ay = e1by; e1az = e1bz; e1bx = e1cx; e1by = e1cy; e1bz = e1cz;                                   \
        h2ax = h2bx; h2ay = h2by; h2az = h2bz;                                                                          \
    } while (0)

    int it = 0;
    // general iterations (prologue, boundary tiles, op-carrying segments): one at a time, window moved
    auto general = [&](int upto) {
        for (; it < upto; ++it) {
            const int i = i0 - 3 + it;
            if (!(cx.interior && g.x0 + i + 3 < g.nxg - 2))
                tb2x_iter<T, R, AM, OPS, true, false>(cx, it, out, c, g, m, fo, TB2X_ARGS_0, TB2X_H2_0, sq, sph, xd, xph);
            else
                tb2x_iter<T, R, AM, OPS, false, false>(cx, it, out, c, g, m, fo, TB2X_ARGS_0, TB2X_H2_0, sq, sph, xd, xph);
            TB2X_ROTATE();
        }
    };
    if (!OPS && cx.interior) {
        // steady range: i >= i0, i + 2 < i1, global plane i + 3 < nxg - 2   <=>   it in [3, hi)
        const int hi = min(n_it - 2, g.nxg - 2 - g.x0 - i0);
        general(min(3, n_it));
        for (; it + 6 <= hi; it += 6) {
            tb2x_iter<T, R, AM, false, false, true>(cx, it + 0, out, c, g, m, fo, TB2X_ARGS_0, TB2X_H2_0, sq, sph, xd, xph);
            tb2x_iter<T, R, AM, false, false, true>(cx, it + 1, out, c, g, m, fo, TB2X_ARGS_1, TB2X_H2_1, sq, sph, xd, xph);
            tb2x_iter<T, R, AM, false, false, true>(cx, it + 2, out, c, g, m, fo, TB2X_ARGS_2, TB2X_H2_0, sq, sph, xd, xph);
            tb2x_iter<T, R, AM, false, false, true>(cx, it + 3, out, c, g, m, fo, TB2X_ARGS_0, TB2X_H2_1, sq, sph, xd, xph);
            tb2x_iter<T, R, AM, false, false, true>(cx, it + 4, out, c, g, m, fo, TB2X_ARGS_1, TB2X_H2_0, sq, sph, xd, xph);
            tb2x_iter<T, R, AM, false, false, true>(cx, it + 5, out, c, g, m, fo, TB2X_ARGS_2, TB2X_H2_1, sq, sph, xd, xph);
        }
    }
    general(n_it);
#undef TB2X_ARGS_0
#undef TB2X_ARGS_1
#undef TB2X_ARGS_2
#undef TB2X_H2_0
#undef TB2X_H2_1
#undef TB2X_ROTATE
}

template <typename T, int R, int AM>
__global__ void __launch_bounds__(32 * (R + 1), 1)
k_fused3d_tb2x(const __grid_constant__ Tb2xMaps maps, const __grid_constant__ Fields<T> out,
               const __grid_constant__ Coefs<T> c, const __grid_constant__ Geom g, const __grid_constant__ FusedTiling t,
               const __grid_constant__ MidOps m, const __grid_constant__ Fold fo, const int S, const int D,
               const int all_arrive)
{
    const int slot = blockIdx.x / (t.ntj * t.ntk);
    if ((t.seg_ops >> slot) & 1ull) tb2x_sweep<T, R, true, AM>(maps, out, c, g, t, m, fo, S, D, all_arrive);
    else tb2x_sweep<T, R, false, AM>(maps, out, c, g, t, m, fo, S, D, all_arrive);
}

}  // namespace fdtd
