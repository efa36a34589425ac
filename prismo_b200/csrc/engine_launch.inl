// kernel launch helpers: two-pass, fused, heterogeneous, two-step sweep and its segment planner
// (textual part of fdtd_engine.cu — one translation unit; not compiled on its own)
// ---- kernels launch helpers --------------------------------------------------------------------------------
template <typename T> static int launch_pass3d(fdtd_engine* e, int phase, int i_begin, int i_end, cudaStream_t s)
{
    if (i_end <= i_begin) return 0;
    constexpr int V = VecOf<T>::V;
    const Geom& g = e->g;
    const int vec_per_row = g.pz / V;
    dim3 block(std::min(vec_per_row, 64), 1, 1);
    block.y = std::max(1, 256 / (int)block.x);
    dim3 grid((vec_per_row + block.x - 1) / block.x, (g.ny + block.y - 1) / block.y, i_end - i_begin);
    Fields<T> f = fields_of<T>(cur_fields(e));
    Coefs<T> c = coefs_of<T>(e);
    if (phase == 0) {
        if (e->het) k_h3d<T, true><<<grid, block, 0, s>>>(f, c, g, i_begin);
        else k_h3d<T, false><<<grid, block, 0, s>>>(f, c, g, i_begin);
    } else {
        if (e->het) k_e3d<T, true><<<grid, block, 0, s>>>(f, c, g, i_begin);
        else k_e3d<T, false><<<grid, block, 0, s>>>(f, c, g, i_begin);
    }
    e->launches++;
    CU(cudaGetLastError());
    return 0;
}

template <typename T> static int launch_pass2d(fdtd_engine* e, int phase, int parity, cudaStream_t s)
{
    const Geom& g = e->g;
    dim3 block(128, 1, 1), grid((g.ny + 127) / 128, g.nx, 1);
    Fields<T> f = fields_of<T>(cur_fields(e));
    Coefs<T> c = coefs_of<T>(e);
    int* cur = e->d_cnt + 6 * (parity & 1);
    int* nxt = e->d_cnt + 6 * ((parity + 1) & 1);
    if (phase == 0) {
        if (e->het) k_h2d<T, true><<<grid, block, 0, s>>>(f, c, g, cur, nxt);
        else k_h2d<T, false><<<grid, block, 0, s>>>(f, c, g, cur, nxt);
    } else {
        if (e->het) k_e2d<T, true><<<grid, block, 0, s>>>(f, c, g, nxt);
        else k_e2d<T, false><<<grid, block, 0, s>>>(f, c, g, nxt);
    }
    e->launches++;
    CU(cudaGetLastError());
    return 0;
}

template <typename T> static int launch_count2d(fdtd_engine* e, int parity, cudaStream_t s)
{
    int* cur = e->d_cnt + 6 * (parity & 1);
    CU(cudaMemsetAsync(cur, 0, 6 * sizeof(int), s));
    void** p = cur_fields(e);
    CFields<T> f; f.ex = (const T*)p[0]; f.ey = (const T*)p[1]; f.ez = (const T*)p[2];
    f.hx = f.hy = f.hz = nullptr;
    const long long n = e->plane_elems * e->g.nx;
    k_count2d<T><<<(int)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, s>>>(f, n, cur);
    e->launches++;
    CU(cudaGetLastError());
    return 0;
}

// sources (group by group, list order) then monitors, for table row (*d_step + step_off)
static AdeIn ade_in_of(fdtd_engine* e)
{
    AdeIn ad{};
    ad.ops = e->d_ade; ad.n = (int)e->ade.size(); ad.aux = e->d_aux; ad.mask = e->d_ade_mask;
    ad.coupled = e->ade_coupled; ad.kj = 8.854187817e-12; ad.kp = 8.854187817e-12 / e->cfg.dt;
    return ad;
}
// does the sweep being launched now carry the dispersive-medium recursions (of the previous step)?
static bool ade_in_this_sweep(const fdtd_engine* e) { return !e->ade.empty() && (e->ade_coupled || e->ade_deferred); }

template <typename T> static int launch_post(fdtd_engine* e, int step_off, int parity, cudaStream_t s, bool defer_ade = false)
{
    void** comp = e->d_comp_ptr[e->cur];
    int* cnt_next = e->cfg.ndim == 2 ? e->d_cnt + 6 * ((parity + 1) & 1) : nullptr;
    for (size_t gidx = 0; gidx < e->grp_first.size(); ++gidx) {
        const long long total = e->grp_threads[gidx];
        if (total == 0) continue;
        k_sources<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
            (T* const*)comp, e->d_src + e->grp_first[gidx], e->grp_count[gidx], total, e->st, e->d_amp, e->n_amp,
            e->d_step, step_off, e->d_prof, cnt_next);
        e->launches++;
        CU(cudaGetLastError());
    }
    if (e->mon_threads > 0) {
        k_monitors<T><<<(unsigned)((e->mon_threads + 255) / 256), 256, 0, s>>>(
            (const T* const*)comp, e->d_mon, (int)e->mon.size(), e->mon_threads, e->st, e->d_phasor, e->n_phasor,
            e->d_step, step_off, e->cfg.dt, (T*)e->d_rec, e->d_dft);
        e->launches++;
        CU(cudaGetLastError());
    }
    if (!e->flux.empty() && e->d_flux_out) {
        dim3 grid(FLUX_BLOCKS, (unsigned)e->flux.size());
        k_flux_partial<T><<<grid, 256, 0, s>>>((const T* const*)comp, e->d_flux, e->st, e->d_flux_partial);
        k_flux_final<<<1, 64, 0, s>>>(e->d_flux, (int)e->flux.size(), e->d_flux_partial, e->d_flux_out, e->d_step, step_off,
                                      std::max(e->n_steps_tab, 1));
        e->launches += 2;
        CU(cudaGetLastError());
    }
    if (e->ade_threads > 0 && e->ade_coupled) {
        // coupled mode: the recursion runs at the beginning of every step, inside the sweep, never here
    } else if (e->ade_threads > 0 && defer_ade) {
        e->ade_deferred = true;                              // the next fused sweep applies it on its input planes
    } else if (e->ade_threads > 0) {
        k_ade<T><<<(unsigned)((e->ade_threads + 255) / 256), 256, 0, s>>>(
            (const T* const*)comp, e->d_ade, (int)e->ade.size(), e->ade_threads, e->st, (T*)e->d_aux, e->d_ade_mask);
        e->launches++;
        CU(cudaGetLastError());
    }
    return 0;
}

// physics mode (opt-in): stable Yee leap-frog + CPML, see fdtd_yee.cuh
template <typename T> static int launch_yee(fdtd_engine* e, int phase, cudaStream_t s)
{
    const Geom& g = e->g;
    dim3 block(64, 4, 1), grid((g.nz + 63) / 64, (g.ny + 3) / 4, g.nx);
    Fields<T> f = fields_of<T>(cur_fields(e));
    Coefs<T> c = coefs_of<T>(e);
    if (phase == 0) {
        if (e->het) k_h3d_yee<T, true><<<grid, block, 0, s>>>(f, c, g, e->cpml, e->slabg);
        else k_h3d_yee<T, false><<<grid, block, 0, s>>>(f, c, g, e->cpml, e->slabg);
    } else {
        if (e->het) k_e3d_yee<T, true><<<grid, block, 0, s>>>(f, c, g, e->cpml, e->slabg);
        else k_e3d_yee<T, false><<<grid, block, 0, s>>>(f, c, g, e->cpml, e->slabg);
    }
    e->launches++;
    CU(cudaGetLastError());
    return 0;
}

// coef: host fp64 [3 axes][6 vectors][N_axis] concatenated axis by axis (x: 6*nx, y: 6*ny, z: 6*nz):
// b, a, 1/kappa at E-derivative (half) positions, then at H-derivative (integer) positions
extern "C" int fdtd_set_cpml(fdtd_engine* e, int32_t thickness, const double* coef)
{
    if (!e) return fail(FDTD_EINVAL, "null engine");
    if (!(e->cfg.flags & FDTD_FLAG_YEE) || e->cfg.ndim != 3)
        return fail(FDTD_ESTATE, "CPML belongs to the opt-in physics mode (FDTD_FLAG_YEE, 3-D)");
    const Geom& g = e->g;
    if (thickness < 0 || 2 * thickness + 1 > std::min(g.nx, std::min(g.ny, g.nz)))
        return fail(FDTD_EINVAL, "CPML thickness %d does not fit the grid", thickness);
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    cudaFree(e->d_cpml_coef); e->d_cpml_coef = nullptr;
    cudaFree(e->d_cpml_coef_f); e->d_cpml_coef_f = nullptr;
    for (int q = 0; q < 12; ++q) { cudaFree(e->cpml.psi[q]); e->cpml.psi[q] = nullptr; cudaFree(e->psiB[q]); e->psiB[q] = nullptr; }
    e->cpml = Cpml{};
    drop_graph(e);
    if (thickness == 0) return 0;
    if (!coef) return fail(FDTD_EINVAL, "fdtd_set_cpml: null coefficients");
    const int N[3] = {g.nx, g.ny, g.nz};
    const size_t total = 6 * ((size_t)g.nx + g.ny + g.nz);
    CU(cudaMalloc(&e->d_cpml_coef, total * sizeof(double)));
    CU(cudaMemcpy(e->d_cpml_coef, coef, total * sizeof(double), cudaMemcpyHostToDevice));
    {
        std::vector<float> cf(total);
        for (size_t q = 0; q < total; ++q) cf[q] = (float)coef[q];
        CU(cudaMalloc(&e->d_cpml_coef_f, total * sizeof(float)));
        CU(cudaMemcpy(e->d_cpml_coef_f, cf.data(), total * sizeof(float), cudaMemcpyHostToDevice));
    }
    size_t off = 0;
    for (int a = 0; a < 3; ++a) {
        for (int v = 0; v < 6; ++v) {
            e->cpml.ax[a].c[v] = e->d_cpml_coef + off + (size_t)v * N[a];
            e->cpml.ax[a].f[v] = e->d_cpml_coef_f + off + (size_t)v * N[a];
        }
        off += (size_t)6 * N[a];
    }
    e->cpml.t = thickness; e->cpml.ns = 2 * thickness + 1;
    const long long ns = e->cpml.ns;
    e->slabg.x_sx = g.sx;                          // x family: (ns, ny, pz)
    e->slabg.y_sx = ns * g.sy;                     // y family: (nx, ns, pz)
    e->slabg.z_pitch = (int)round_up(ns, 4);       // z family: (nx, ny, z_pitch)
    const size_t bx = (size_t)ns * g.sx * e->esz, by = (size_t)g.nx * ns * g.sy * e->esz,
                 bz = (size_t)g.nx * g.ny * e->slabg.z_pitch * e->esz;
    static const int family[12] = {1, 2, 2, 0, 0, 1, 1, 2, 2, 0, 0, 1};     // axis of each psi array
    for (int q = 0; q < 12; ++q) {
        const size_t b = family[q] == 0 ? bx : (family[q] == 1 ? by : bz);
        CU(cudaMalloc(&e->cpml.psi[q], b));
        CU(cudaMemset(e->cpml.psi[q], 0, b));
        CU(cudaMalloc(&e->psiB[q], b));
        CU(cudaMemset(e->psiB[q], 0, b));
        e->psi_bytes[q] = b;
    }
    return 0;
}

static bool use_fused(const fdtd_engine* e)
{
    // slabs (nxg != nx) use the fused sweep too, but through fdtd_sweep: the host interleaves the halo exchange
    return e->cfg.ndim == 3 && !e->het && !(e->cfg.flags & (FDTD_FLAG_TWO_PASS | FDTD_FLAG_YEE));
}

static int ensure_set_b(fdtd_engine* e)
{
    if (e->fldB[0]) return 0;
    for (int c = 0; c < 6; ++c) {
        CU(cudaMalloc(&e->fldB[c], e->array_elems * e->esz));
        CU(cudaMemsetAsync(e->fldB[c], 0, e->array_elems * e->esz, e->stream));
    }
    CU(cudaMemcpyAsync(e->d_comp_ptr[1], e->fldB, 6 * sizeof(void*), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

// Work-item order of a sweep that carries dispersive-medium recursions: the (x-segment, tile) items that meet a recursion
// box first (they run a few times longer: dispatched last they would be the launch's tail).  Rebuilt when the tiling or
// the recursion list changes; must be called OUTSIDE stream capture (allocation + synchronous copy).
static int ensure_ade_order(fdtd_engine* e, const FusedTiling& t, int own_rows, int V)
{
    const long long items = (long long)t.nseg * t.ntj * t.ntk;
    const std::vector<long long> sig = {e->ade_epoch, items, t.nseg, t.ntj, t.ntk, t.lx, t.own_lanes, own_rows, V, t.i_begin, t.i_end};
    if (e->d_ade_order && sig == e->ade_order_sig) return 0;
    std::vector<int> first, rest;
    const int ntiles = t.ntj * t.ntk;
    for (long long it = 0; it < items; ++it) {
        const int seg = (int)(it / ntiles), tile = (int)(it - (long long)seg * ntiles);
        const int tj = tile / t.ntk, tk = tile - tj * t.ntk;
        const int i0 = t.i_begin + seg * t.lx, i1 = std::min(i0 + t.lx, t.i_end);
        const int j0 = tj * own_rows, j1 = j0 + own_rows, k0 = tk * t.own_lanes * V, k1 = (tk + 1) * t.own_lanes * V;
        bool touched = false;
        for (const AdeOp& op : e->ade)
            if (j1 > op.lo[1] && j0 < op.lo[1] + op.n[1] && k1 > op.lo[2] && k0 < op.lo[2] + op.n[2] && i1 > op.lo[0] &&
                i0 < op.lo[0] + op.n[0]) { touched = true; break; }
        (touched ? first : rest).push_back((int)it);
    }
    first.insert(first.end(), rest.begin(), rest.end());
    if (e->ade_order_items < items) {
        cudaFree(e->d_ade_order); e->d_ade_order = nullptr; e->ade_order_items = 0;
        CU(cudaMalloc(&e->d_ade_order, (size_t)items * sizeof(int)));
        e->ade_order_items = items;
    }
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemcpy(e->d_ade_order, first.data(), (size_t)items * sizeof(int), cudaMemcpyHostToDevice));
    e->ade_order_sig = sig;
    return 0;
}

template <typename T, int TJ> static FusedTiling fused_tiling(const fdtd_engine* e, int i_begin, int i_end)
{
    constexpr int V = VecOf<T>::V;
    const Geom& g = e->g;
    FusedTiling t{};
    t.i_begin = i_begin; t.i_end = i_end;
    const int vec_per_row = g.pz / V;
    const int ncols = (vec_per_row + 29) / 30;
    int own = (vec_per_row + ncols - 1) / ncols;
    own += own & 1;                                    // even: tiles start on 32-byte sectors
    t.own_lanes = std::min(own, 30);
    t.ntk = (vec_per_row + t.own_lanes - 1) / t.own_lanes;
    t.ntj = (g.ny + TJ - 1) / TJ;
    const int planes = i_end - i_begin;
    int lx = e->fused_lx;
    if (lx <= 0) {
        // >= ~40 waves of 148 CTAs so the ragged last wave costs ~1%, with segments of >= 32 planes
        // (each segment re-reads one plane of H and two of E as its prologue); short segments also keep
        // co-resident CTAs on nearby planes, so tile rims are re-read from L2 instead of DRAM
        const long long tiles = (long long)t.ntj * t.ntk;
        long long want = (148ll * 40 + tiles - 1) / tiles;
        lx = (int)std::max<long long>(32, (planes + want - 1) / std::max<long long>(want, 1));
    }
    t.lx = std::min(lx, planes);
    t.nseg = (planes + t.lx - 1) / t.lx;
    return t;
}

template <typename T> static FusedTiling het_tiling(const fdtd_engine* e)
{
    constexpr int R = kHetRows, V = Vec8<T>::V;
    const Geom& g = e->g;
    FusedTiling t{};
    t.i_begin = 0; t.i_end = g.nx;
    t.own_lanes = kHetOwnLanes;
    const int vec_per_row = g.pz / V;
    t.ntk = (vec_per_row + t.own_lanes - 1) / t.own_lanes;
    t.ntj = (g.ny + (R - 2) - 1) / (R - 2);
    int lx = e->fused_lx;
    if (lx <= 0) {
        const long long tiles = (long long)t.ntj * t.ntk;
        long long want = (148ll * 40 + tiles - 1) / tiles;
        lx = (int)std::max<long long>(32, (g.nx + want - 1) / std::max<long long>(want, 1));
        while (lx > 8 && tiles * ((g.nx + lx - 1) / lx) < 148 * 2) lx = (lx + 1) / 2;
    }
    t.lx = std::min(lx, g.nx);
    t.nseg = (g.nx + t.lx - 1) / t.lx;
    return t;
}

// one fused sweep over planes [i_begin, i_end): reads the current set, writes the other one
template <typename T, int TJ> static int launch_fused_tj(fdtd_engine* e, int i_begin, int i_end, cudaStream_t s)
{
    const Geom& g = e->g;
    void** src = cur_fields(e);
    void** dst = e->cur ? e->fld : e->fldB;
    CFields<T> in;
    in.ex = (const T*)src[0]; in.ey = (const T*)src[1]; in.ez = (const T*)src[2];
    in.hx = (const T*)src[3]; in.hy = (const T*)src[4]; in.hz = (const T*)src[5];
    Fields<T> out = fields_of<T>(dst);
    FusedTiling t = fused_tiling<T, TJ>(e, i_begin, i_end);
    t.halo_flag = nullptr; t.halo_need = 0; t.error_word = nullptr; t.timeout_ns = e->slab.timeout_ns;
    if (e->slab.connected && e->slab.has_right && i_end == g.nx) {
        t.halo_flag = e->slab.flags; t.halo_need = (int)e->slab.step + 1; t.error_word = e->slab.flags + 2;
    }
    const size_t smem = fused_smem_bytes<T, TJ>();
    const bool ade = ade_in_this_sweep(e) && e->g.nxg == e->g.nx;
    auto kern = k_fused3d<T, TJ, 0, false>;
    if (sizeof(T) == 8 && fold64(e)) kern = k_fused3d<T, TJ, (sizeof(T) == 8 ? 1 : 0), false>;
    if (ade) {
        kern = k_fused3d<T, TJ, 0, true>;
        if (sizeof(T) == 8 && fold64(e)) kern = k_fused3d<T, TJ, (sizeof(T) == 8 ? 1 : 0), true>;
    }
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 block(32, TJ + 1, 1);
    const unsigned items = (unsigned)t.nseg * t.ntj * t.ntk;
    AdeIn ad{};
    if (ade) {
        ad = ade_in_of(e);
        if (e->ade_order_sig.size() && e->ade_order_sig[0] == e->ade_epoch && e->ade_order_sig[1] == (long long)items)
            ad.order = e->d_ade_order;                   // prepared by prepare_ade_order for this tiling
    }
    kern<<<items, block, smem, s>>>(in, out, coefs_of<T>(e), g, t, fold_of(e), ad);
    if (ade && i_end == g.nx) e->ade_deferred = false;
    e->launches++;
    CU(cudaGetLastError());
    return 0;
}

// physics mode, opt-in: Yee leap-frog + CPML as one fused sweep; fields AND psi go from the current set to the other one
static bool use_yee_fused(const fdtd_engine* e)
{
    return e->cfg.ndim == 3 && (e->cfg.flags & FDTD_FLAG_YEE) && e->yee_fused && !e->het && e->g.nxg == e->g.nx;
}

template <typename T> static int launch_yee_fused(fdtd_engine* e, cudaStream_t s)
{
    constexpr int V = VecOf<T>::V, TJ = kFusedTJ;
    const Geom& g = e->g;
    void** src = cur_fields(e);
    void** dst = e->cur ? e->fld : e->fldB;
    CFields<T> in;
    in.ex = (const T*)src[0]; in.ey = (const T*)src[1]; in.ez = (const T*)src[2];
    in.hx = (const T*)src[3]; in.hy = (const T*)src[4]; in.hz = (const T*)src[5];
    Fields<T> out = fields_of<T>(dst);
    Cpml pm = e->cpml;                                   // psi of the CURRENT set in, the other set out
    PsiOut pout;
    for (int q = 0; q < 12; ++q) {
        pm.psi[q] = e->cur ? e->psiB[q] : e->cpml.psi[q];
        pout.p[q] = e->cur ? e->cpml.psi[q] : e->psiB[q];
    }
    FusedTiling t{};
    t.i_begin = 0; t.i_end = g.nx;
    const int vec_per_row = g.pz / V;
    const int ncols = (vec_per_row + 29) / 30;
    int own = (vec_per_row + ncols - 1) / ncols;
    own += own & 1;
    t.own_lanes = std::min(own, 30);
    t.ntk = (vec_per_row + t.own_lanes - 1) / t.own_lanes;
    t.ntj = (g.ny + TJ - 1) / TJ;
    int lx = e->fused_lx;
    if (lx <= 0) {
        const long long tiles = (long long)t.ntj * t.ntk;
        long long want = (148ll * 40 + tiles - 1) / tiles;
        lx = (int)std::max<long long>(32, (g.nx + want - 1) / std::max<long long>(want, 1));
    }
    t.lx = std::min(lx, g.nx);
    t.nseg = (g.nx + t.lx - 1) / t.lx;
    const size_t smem = fused_smem_bytes<T, TJ>();
    auto kern = k_fused3d_yee<T, TJ>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 block(32, TJ + 1, 1);
    kern<<<(unsigned)t.nseg * t.ntj * t.ntk, block, smem, s>>>(in, out, coefs_of<T>(e), g, t, pm, pout, e->slabg);
    e->launches++;
    CU(cudaGetLastError());
    e->cur ^= 1;
    return 0;
}

// physics mode, default where it applies: the TMA-fed fused sweep (fdtd_yeex.cuh)
static bool use_yeex(const fdtd_engine* e)
{
    return use_yee_fused(e) && e->yee_fused >= 2 && (long long)e->g.pz * (long long)e->esz >= kYeexBoxBytes &&
           e->array_elems < (1ll << 32);
}
static int ensure_ymaps(fdtd_engine* e);

template <typename T> static int launch_yeex(fdtd_engine* e, cudaStream_t s)
{
    constexpr int R = YeexRows<T>::R, V = Vec8<T>::V;
    if (int rc = ensure_ymaps(e)) return rc;
    const Geom& g = e->g;
    void** dst = e->cur ? e->fld : e->fldB;
    Fields<T> out = fields_of<T>(dst);
    Cpml pm = e->cpml;                                   // psi of the CURRENT set in, the other set out
    PsiOut pout;
    for (int q = 0; q < 12; ++q) {
        pm.psi[q] = e->cur ? e->psiB[q] : e->cpml.psi[q];
        pout.p[q] = e->cur ? e->cpml.psi[q] : e->psiB[q];
    }
    FusedTiling t{};
    t.i_begin = 0; t.i_end = g.nx;
    t.own_lanes = 30;
    const int vec_per_row = g.pz / V;
    t.ntk = (vec_per_row + t.own_lanes - 1) / t.own_lanes;
    t.ntj = (g.ny + (R - 1) - 1) / (R - 1);
    int lx = e->fused_lx;
    if (lx <= 0) {
        // Short segments: measured on 1024^3 (profiles/r02_tuning.md) 64 planes 92.3, 128: 91.4, 205: 87.7, 342: 74.4,
        // 1024: 62.5 Gcell/s — many short-lived CTAs keep neighbouring tiles on nearby planes, so the rim rows / lanes a
        // tile re-reads are still in L2, and the slow boundary tiles do not leave a long tail.  Each segment pays one
        // prologue plane (1.6 % at 64).
        lx = 64;
    }
    t.lx = std::min(lx, g.nx);
    t.nseg = (g.nx + t.lx - 1) / t.lx;
    const int S = e->yeex_stages, D = e->yeex_slots;
    const size_t smem = yeex_smem_bytes<R>(S, D);
    if (smem > 227 * 1024) return fail(FDTD_EINVAL, "physics sweep rings (%d stages, %d slots) need %zu B of shared memory", S, D, smem);
    auto kern = k_fused3d_yeex<T, R, 0>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 block(32, R + 1, 1);
    kern<<<(unsigned)t.nseg * t.ntj * t.ntk, block, smem, s>>>(e->ymaps[e->cur], out, coefs_of<T>(e), g, t, pm, pout, e->slabg, S, D);
    e->launches++;
    CU(cudaGetLastError());
    e->cur ^= 1;
    return 0;
}

// one fused sweep over planes [i_begin, i_end): reads the current set, writes the other one
template <typename T> static int launch_fused(fdtd_engine* e, int i_begin, int i_end, cudaStream_t s)
{
    return launch_fused_tj<T, kFusedTJ>(e, i_begin, i_end, s);
}

// heterogeneous media, one GPU: fused one-step sweep that also streams the four coefficient arrays
static bool het_sweep_ok(const fdtd_engine* e)          // the kernel itself also runs on an x-slab (fdtd_slab_run)
{
    return e->cfg.ndim == 3 && e->het && !(e->cfg.flags & (FDTD_FLAG_TWO_PASS | FDTD_FLAG_YEE)) &&
           e->array_elems < (1ll << 32) && e->het_fused;
}
static bool use_het_fused(const fdtd_engine* e) { return het_sweep_ok(e) && e->g.nxg == e->g.nx; }

template <typename T> static int launch_het(fdtd_engine* e, cudaStream_t s)
{
    constexpr int R = kHetRows;
    const Geom& g = e->g;
    void** src = cur_fields(e);
    void** dst = e->cur ? e->fld : e->fldB;
    CFields<T> in;
    in.ex = (const T*)src[0]; in.ey = (const T*)src[1]; in.ez = (const T*)src[2];
    in.hx = (const T*)src[3]; in.hy = (const T*)src[4]; in.hz = (const T*)src[5];
    Fields<T> out = fields_of<T>(dst);
    FusedTiling t = het_tiling<T>(e);
    t.timeout_ns = e->slab.timeout_ns;
    if (e->slab.connected && e->slab.has_right) {
        // the last x-segment reads the right neighbour's planes 0 and 1 (same 7 halo planes as the uniform one-step sweep)
        // and one ghost plane of the coefficient arrays, which is static (fdtd_set_coeffs with nx + 1 planes)
        if (e->coef_planes < g.nx + 1)
            return fail(FDTD_ESTATE, "x-slab with a right neighbour: fdtd_set_coeffs needs nx + 1 = %d planes", g.nx + 1);
        t.halo_flag = e->slab.flags; t.halo_need = (int)e->slab.step + 1; t.error_word = e->slab.flags + 2;
    }
    const size_t smem = e->aniso ? het_smem_bytes<T, R, true>() : het_smem_bytes<T, R, false>();
    const bool ade = ade_in_this_sweep(e);
    const Coefs<T> cf = coefs_of<T>(e);
    const bool idx = cf.mat != nullptr;         // material-index coding (fdtd_rasterize + option "het_indexed")
    auto pick = [&](auto aniso_c, auto idx_c) {
        constexpr bool A = decltype(aniso_c)::value, I = decltype(idx_c)::value;
        return ade ? k_fused3d_het<T, R, true, A, I> : k_fused3d_het<T, R, false, A, I>;
    };
    auto kern = e->aniso ? (idx ? pick(std::true_type{}, std::true_type{}) : pick(std::true_type{}, std::false_type{}))
                         : (idx ? pick(std::false_type{}, std::true_type{}) : pick(std::false_type{}, std::false_type{}));
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 block(32, R, 1);
    const long long items = (long long)t.nseg * t.ntj * t.ntk;
    AdeIn ad{};
    if (ade) {
        ad = ade_in_of(e);
        if (e->ade_order_sig.size() && e->ade_order_sig[0] == e->ade_epoch && e->ade_order_sig[1] == items)
            ad.order = e->d_ade_order;                   // prepared by prepare_ade_order for this tiling
    }
    kern<<<(unsigned)items, block, smem, s>>>(in, out, cf, g, t, (int)e->planes_alloc, ad);
    e->ade_deferred = false;
    e->launches++;
    CU(cudaGetLastError());
    e->cur ^= 1;
    return 0;
}

static bool tb2_ok(const fdtd_engine* e)
{
    return e->tb2 && use_fused(e) && e->ade.empty() && e->flux.empty() && e->array_elems < (1ll << 32);
}
static bool use_tb2(const fdtd_engine* e) { return tb2_ok(e) && e->g.nxg == e->g.nx; }

// x-segments of one two-step sweep (FusedTiling::seg_lo/seg_hi/seg_ops), in dispatch order.
//  * A segment [a, b) applies the intermediate step's sources / monitors on planes [a, b+1]: planes that carry ops
//    get NARROW zones of their own ([p-2, p+2) widened to >= 8 planes), so that the op-carrying code path (10 % slower)
//    runs on a few planes only and everything else takes the op-free path.
//  * Every segment pays 3 prologue planes; CTAs are dispatched in waves of 148: the number of bulk parts minimises
//    ceil(tiles * n / 148) * (nx / n + 3).
//  * Dispatch order: bulk parts first, zones (short items) last to fill the tail; within each kind the segment that
//    reads the ghost planes (slabs: it spins until the right neighbour's push has landed) goes last.
struct SegIv { int lo, hi; bool ops; };
static std::vector<SegIv> plan_segments(int nx, const unsigned char* flags, int nflag, long long tiles, bool halo,
                                        int fused_lx, int zones_mode, int* lx_out)
{
    typedef SegIv Iv;
    // target length of a bulk part
    int lxt = fused_lx;
    if (lxt <= 0) {
        double best = 1e300;
        int best_n = 1;
        // measured on 1024^3: parts of 64..256 planes within 0.5 % of each other, 512 planes 1.5 % slower (ragged tail)
        for (int n = (nx + 255) / 256; n <= kMaxSegs / 2 && (n == 1 || nx / n >= 8); ++n) {
            const double waves = std::ceil((double)tiles * n / 148.0);
            const double cost = waves * ((double)(nx + n - 1) / n + 3.0);
            if (cost < best * 0.999) { best = cost; best_n = n; }
        }
        lxt = (nx + best_n - 1) / best_n;
    }
    lxt = std::max(lxt, (nx + kMaxSegs / 2 - 1) / (kMaxSegs / 2));
    std::vector<Iv> zones;
    // narrow zones cost two more segments (6 prologue planes + CTA start-up): worth it only against long bulk parts
    const bool want_zones = zones_mode < 0 ? lxt >= 112 : zones_mode != 0;
    for (int W = 8; want_zones; W *= 2) {
        zones.clear();
        for (int p = 0; p < nflag; ++p) {
            if (!flags[p]) continue;
            int lo = std::max(0, std::min(p - 2, nx - W));
            int hi = std::min(nx, std::max(p + 2, lo + W));
            if (lo >= nx) continue;
            if (!zones.empty() && lo <= zones.back().hi) zones.back().hi = std::max(zones.back().hi, hi);
            else zones.push_back({lo, hi, true});
        }
        if ((int)zones.size() <= kMaxSegs / 4 || W >= nx) break;
    }
    // intervals in x order: zones and the gaps between them, each cut into equal parts of about lxt planes
    std::vector<Iv> ivs;
    int at = 0;
    for (size_t z = 0; z <= zones.size(); ++z) {
        const int lo = z < zones.size() ? zones[z].lo : nx;
        if (lo > at) ivs.push_back({at, lo, false});
        if (z < zones.size()) { ivs.push_back(zones[z]); at = zones[z].hi; }
    }
    std::vector<Iv> parts;
    for (const Iv& iv : ivs) {
        const int len = iv.hi - iv.lo;
        int n = std::max(1, (len + lxt / 2) / lxt);
        for (int q = 0; q < n; ++q) parts.push_back({iv.lo + (int)((long long)len * q / n), iv.lo + (int)((long long)len * (q + 1) / n), iv.ops});
    }
    // a slab whose only segment reads the ghost planes would make every CTA spin for the neighbour's push: cut it
    if (halo && parts.size() == 1 && nx >= 16) {
        const Iv p = parts[0];
        parts = {{p.lo, (p.lo + p.hi) / 2, p.ops}, {(p.lo + p.hi) / 2, p.hi, p.ops}};
    }
    while ((int)parts.size() > kMaxSegs) {                  // cannot happen with the caps above; stay safe: merge neighbours
        size_t k = 0;
        for (size_t q = 0; q + 1 < parts.size(); ++q)
            if (parts[q + 1].hi - parts[q].lo < parts[k + 1].hi - parts[k].lo) k = q;
        parts[k].hi = parts[k + 1].hi; parts[k].ops |= parts[k + 1].ops;
        parts.erase(parts.begin() + k + 1);
    }
    for (Iv& pt : parts) {                                  // the rule the kernel needs: ops on planes [lo, hi + 1]
        pt.ops = false;
        for (int p = pt.lo; p <= pt.hi + 1 && p < nflag; ++p) pt.ops |= flags[p] != 0;
    }
    // dispatch order
    std::stable_sort(parts.begin(), parts.end(), [&](const Iv& a, const Iv& b) {
        if (a.ops != b.ops) return !a.ops;
        const bool ha = halo && a.hi + 3 >= nx, hb = halo && b.hi + 3 >= nx;
        if (ha != hb) return !ha;
        return a.lo < b.lo;
    });
    if (halo && parts.size() > 1 && parts[0].hi + 3 >= nx) std::rotate(parts.begin(), parts.begin() + 1, parts.end());
    if (lx_out) *lx_out = lxt;
    return parts;
}

static void plan_tb2_segments(const fdtd_engine* e, long long tiles, bool any_ops, bool halo, FusedTiling& t)
{
    const int nflag = any_ops ? (int)e->plane_flags_host.size() : 0;
    const std::vector<SegIv> parts = plan_segments(e->g.nx, e->plane_flags_host.data(), nflag, tiles, halo, e->fused_lx,
                                                   e->tb2_zones, &t.lx);
    t.nseg = (int)parts.size();
    t.seg_ops = 0;
    for (int q = 0; q < t.nseg; ++q) {
        t.seg_lo[q] = parts[q].lo; t.seg_hi[q] = parts[q].hi;
        if (parts[q].ops) t.seg_ops |= 1ull << q;
    }
}

// host-only: the segment plan for a hypothetical slab (unit tests of the planner run without a GPU)
extern "C" int fdtd_plan_segments(int32_t nx, const uint8_t* plane_flags, int32_t n_flags, int64_t tiles, int32_t halo,
                                  int32_t fused_lx, int32_t zones_mode, int32_t* seg_lo, int32_t* seg_hi, int32_t* seg_ops,
                                  int32_t max_segs)
{
    if (nx <= 0 || tiles <= 0 || n_flags < 0 || (n_flags && !plane_flags) || !seg_lo || !seg_hi || !seg_ops)
        return fail(FDTD_EINVAL, "fdtd_plan_segments: bad argument");
    const std::vector<SegIv> parts = plan_segments(nx, plane_flags, n_flags, tiles, halo != 0, fused_lx, zones_mode, nullptr);
    if ((int)parts.size() > max_segs) return fail(FDTD_EINVAL, "fdtd_plan_segments: %d segments > max_segs", (int)parts.size());
    for (size_t q = 0; q < parts.size(); ++q) { seg_lo[q] = parts[q].lo; seg_hi[q] = parts[q].hi; seg_ops[q] = parts[q].ops; }
    return (int)parts.size();
}

// ---- TMA-fed variant of the two-step sweep (fdtd_tb2x.cuh) ---------------------------------------------------------------
// cuTensorMapEncodeTiled comes from the driver; the library links only the runtime, so fetch the entry point at run time.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// descriptors of the six arrays of both buffer sets: tensor (pz, ny, planes_alloc), box (256 bytes, R rows, 1 plane);
// rows / columns / planes outside the tensor read as zero (the padding and guard planes of the layout, for free)
static int encode_maps(fdtd_engine* e, Tb2xMaps* maps, int box_bytes, int box_rows, int promo = 3)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(FDTD_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const Geom& g = e->g;
    const bool d64 = e->cfg.dtype == FDTD_F64;
    const cuuint64_t dims[3] = {(cuuint64_t)g.pz, (cuuint64_t)g.ny, (cuuint64_t)e->planes_alloc};
    const cuuint64_t strides[2] = {(cuuint64_t)g.sy * e->esz, (cuuint64_t)g.sx * e->esz};
    const cuuint32_t box[3] = {(cuuint32_t)(box_bytes / e->esz), (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    for (int s = 0; s < 2; ++s)
        for (int c = 0; c < 6; ++c) {
            void* base = s ? e->fldB[c] : e->fld[c];
            CUresult r = enc(&maps[s].m[c], d64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base,
                             dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(FDTD_ECUDA, "cuTensorMapEncodeTiled failed (%d) for set %d array %d", (int)r, s, c);
        }
    return 0;
}
static int ensure_tmaps(fdtd_engine* e)
{
    if (e->tmaps_ok) return 0;
    if (int rc = encode_maps(e, e->tmaps, kTb2xRowBytes, kTb2xRows)) return rc;
    e->tmaps_ok = true;
    return 0;
}
// the physics sweep's boxes are 16 B wider (fdtd_yeex.cuh: the box origin must stay 16-byte aligned)
static int ensure_ymaps(fdtd_engine* e)
{
    if (e->ymaps_ok) return 0;
    if (int rc = encode_maps(e, e->ymaps, kYeexBoxBytes, e->esz == 4 ? kYeexBoxRowsOf<float> : kYeexBoxRowsOf<double>,
                             getenv("FDTD_B200_YEEX_L2PROMO") ? atoi(getenv("FDTD_B200_YEEX_L2PROMO")) : 3)) return rc;
    e->ymaps_ok = true;
    return 0;
}

static bool use_tb2x(const fdtd_engine* e)
{
    // the TMA box is 256 B wide and R rows high: rows of the arrays must be at least that long
    return e->tb2x && (long long)e->g.pz * (long long)e->esz >= kTb2xRowBytes;
}

template <typename T> static int launch_tb2x(fdtd_engine* e, const Fields<T>& out, const FusedTiling& t, const MidOps& m,
                                             cudaStream_t s)
{
    constexpr int R = kTb2xRows;
    if (int rc = ensure_tmaps(e)) return rc;
    const int S = e->tb2x_stages, D = e->tb2x_slots;
    static const size_t pad = getenv("FDTD_B200_TB2X_PAD") ? (size_t)atoll(getenv("FDTD_B200_TB2X_PAD")) : 0;   // tuning experiment
    const size_t smem = tb2x_smem_bytes<R>(S, D) + pad;
    if (smem > 227 * 1024) return fail(FDTD_EINVAL, "two-step sweep rings (%d stages, %d slots) need %zu B of shared memory", S, D, smem);
    auto kern = k_fused3d_tb2x<T, R, 0>;
    if (sizeof(T) == 8 && fold64(e)) kern = k_fused3d_tb2x<T, R, (sizeof(T) == 8 ? 1 : 0)>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 block(32, R + 1, 1);
    const unsigned items = (unsigned)t.nseg * t.ntj * t.ntk;
    static const int all_arrive = getenv("FDTD_B200_TB2X_ARRIVE_ALL") ? atoi(getenv("FDTD_B200_TB2X_ARRIVE_ALL")) : 0;
    kern<<<items, block, smem, s>>>(e->tmaps[e->cur], out, coefs_of<T>(e), e->g, t, m, fold_of(e), S, D, all_arrive);
    return 0;
}

// TWO steps in one pass over planes [0, nx): reads the current set, writes the other one; the intermediate
// step's sources / monitors (table row *d_step + step_off) are applied inside the kernel
template <typename T> static int launch_tb2(fdtd_engine* e, int step_off, cudaStream_t s)
{
    constexpr int R = kTb2Rows, V = Vec8<T>::V;
    const Geom& g = e->g;
    void** src = cur_fields(e);
    void** dst = e->cur ? e->fld : e->fldB;
    CFields<T> in;
    in.ex = (const T*)src[0]; in.ey = (const T*)src[1]; in.ez = (const T*)src[2];
    in.hx = (const T*)src[3]; in.hy = (const T*)src[4]; in.hz = (const T*)src[5];
    Fields<T> out = fields_of<T>(dst);
    FusedTiling t;
    t.i_begin = 0; t.i_end = g.nx;
    t.halo_flag = nullptr; t.halo_need = 0; t.error_word = nullptr; t.timeout_ns = e->slab.timeout_ns;
    if (e->slab.connected && e->slab.has_right) {
        t.halo_flag = e->slab.flags; t.halo_need = (int)e->slab.step + 1; t.error_word = e->slab.flags + 2;
    }
    const int vec_per_row = g.pz / V;
    t.own_lanes = tb2_own_lanes<T>();
    if (const char* ow = getenv("FDTD_B200_TB2_OWN")) t.own_lanes = std::min(t.own_lanes, std::max(2, atoi(ow)));   // tuning experiment
    t.ntk = (vec_per_row + t.own_lanes - 1) / t.own_lanes;
    const int own_rows = (use_tb2x(e) ? kTb2xRows : R) - 4;       // the four stages shrink validity by one row each
    t.ntj = (g.ny + own_rows - 1) / own_rows;
    MidOps m{};
    m.src = e->d_src; m.n_src = 0;
    for (int c : e->grp_count) m.n_src += c;
    m.amp = e->d_amp; m.n_amp = e->n_amp; m.prof = e->d_prof;
    m.mon = e->d_mon; m.n_mon = (int)e->mon.size();
    m.phasors = e->d_phasor; m.n_phasor = e->n_phasor;
    m.rec = e->d_rec; m.dft = e->d_dft; m.dt = e->cfg.dt;
    m.step_ptr = e->d_step; m.step_off = step_off;
    m.gsrc = e->d_src_ghost; m.n_gsrc = (int)e->src_ghost.size();
    m.n_planes = g.nx + 4;
    const bool any_ops = m.n_src || m.n_mon || m.n_gsrc;
    m.plane_flags = any_ops ? e->d_plane_flags : nullptr;
    m.op_lo = 1 << 30; m.op_span = 0;                    // no plane passes the range test
    if (any_ops) {
        int lo = -1, hi = -1;
        for (int p = 0; p < (int)e->plane_flags_host.size() && p < m.n_planes; ++p)
            if (e->plane_flags_host[p]) { if (lo < 0) lo = p; hi = p; }
        if (lo >= 0) { m.op_lo = lo; m.op_span = hi - lo; }
    }
    const size_t smem = tb2_smem_bytes<T, R>();
    dim3 block(32, R, 1);
    const Coefs<T> cf = coefs_of<T>(e);
    const Fold fo = fold_of(e);
    plan_tb2_segments(e, (long long)t.ntj * t.ntk, any_ops, t.halo_flag != nullptr, t);
    if (use_tb2x(e)) {
        if (int rc = launch_tb2x<T>(e, out, t, m, s)) return rc;
    } else {
        auto kern = k_fused3d_tb2<T, R, 0>;
        if (sizeof(T) == 8 && fold64(e)) kern = k_fused3d_tb2<T, R, (sizeof(T) == 8 ? 1 : 0)>;
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned items = (unsigned)t.nseg * t.ntj * t.ntk;
        kern<<<items, block, smem, s>>>(in, out, cf, g, t, m, (int)e->planes_alloc, fo);
    }
    e->launches++;
    CU(cudaGetLastError());
    e->cur ^= 1;
    return 0;
}

// two full steps: temporally blocked sweep (step A's sources/monitors inside), then step B's sources/monitors
template <typename T> static int two_steps(fdtd_engine* e, int step_off, cudaStream_t s)
{
    if (int rc = launch_tb2<T>(e, step_off, s)) return rc;
    return launch_post<T>(e, step_off + 1, 0, s);
}

// 3-D field update of one step, in two halves: half 0 = H pass (or the whole fused sweep), half 1 = E pass
template <typename T> static int step_fields3d(fdtd_engine* e, int half, cudaStream_t s)
{
    if (use_fused(e)) {
        if (half == 1) return 0;
        if (int rc = ensure_set_b(e)) return rc;
        if (int rc = launch_fused<T>(e, 0, e->g.nx, s)) return rc;
        e->cur ^= 1;
        return 0;
    }
    if (use_yee_fused(e)) {
        if (half == 1) return 0;
        if (int rc = ensure_set_b(e)) return rc;
        if (use_yeex(e)) return launch_yeex<T>(e, s);
        return launch_yee_fused<T>(e, s);
    }
    if (e->cfg.flags & FDTD_FLAG_YEE) return launch_yee<T>(e, half, s);
    if (use_het_fused(e)) {
        if (half == 1) return 0;
        if (int rc = ensure_set_b(e)) return rc;
        return launch_het<T>(e, s);
    }
    return launch_pass3d<T>(e, half, 0, e->g.nx, s);
}

// can the step after this one apply this step's dispersive-medium recursions inside its sweep?
static bool ade_sweep_capable(const fdtd_engine* e)
{
    return e->ade_fused && !e->ade.empty() && e->cfg.ndim == 3 && e->g.nxg == e->g.nx && (use_fused(e) || use_het_fused(e));
}

// before a run (never inside stream capture): the box-first item order of the sweeps that will carry recursions
template <typename T> static int prepare_ade_order(fdtd_engine* e)
{
    if (e->ade.empty() || !(e->ade_coupled || ade_sweep_capable(e))) return 0;
    if (use_het_fused(e)) return ensure_ade_order(e, het_tiling<T>(e), kHetRows - 2, Vec8<T>::V);
    if (use_fused(e)) return ensure_ade_order(e, fused_tiling<T, kFusedTJ>(e, 0, e->g.nx), kFusedTJ, VecOf<T>::V);
    return 0;
}

template <typename T> static int one_step(fdtd_engine* e, int step_off, int parity, cudaStream_t s, bool defer_ade = false)
{
    if (e->cfg.ndim == 3) {
        if (int rc = step_fields3d<T>(e, 0, s)) return rc;
        if (int rc = step_fields3d<T>(e, 1, s)) return rc;
    } else {
        if (int rc = launch_pass2d<T>(e, 0, parity, s)) return rc;
        if (int rc = launch_pass2d<T>(e, 1, parity, s)) return rc;
    }
    return launch_post<T>(e, step_off, parity, s, defer_ade);
}

static bool has_tables(const fdtd_engine* e) { return !e->src.empty() || !e->mon.empty() || !e->src_ghost.empty() || !e->flux.empty(); }
static bool has_post(const fdtd_engine* e) { return has_tables(e) || !e->ade.empty(); }

template <typename T> static int run_steps(fdtd_engine* e, int n)
{
    cudaStream_t s = e->stream;
    if (e->ade_coupled && !e->ade.empty() && !(e->cfg.ndim == 3 && e->g.nxg == e->g.nx && (use_fused(e) || use_het_fused(e))))
        return fail(FDTD_ESTATE, "coupled dispersive media need the fused one-step sweeps (3-D, one GPU, no two-pass / physics flag)");
    if (int rc = prepare_ade_order<T>(e)) return rc;
    if (use_fused(e) || use_het_fused(e) || use_yee_fused(e)) if (int rc = ensure_set_b(e)) return rc;
    if (e->cfg.ndim == 2) if (int rc = launch_count2d<T>(e, 0, s)) return rc;
    const bool use_graph = !(e->cfg.flags & FDTD_FLAG_NO_GRAPH) && n >= 32;
    int done = 0;
    if (use_graph) {
        const int G = 16;                  // even: the ping-pong set and the 2-D counter parity come back
        const int c0 = e->cur;
        if (!e->gexec[c0]) {
            cudaGraph_t graph = nullptr;
            const long long l0 = e->launches;
            CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            int rc = 0;
            if (use_tb2(e)) for (int q = 0; q < G && !rc; q += 2) rc = two_steps<T>(e, q, s);
            else for (int q = 0; q < G && !rc; ++q) rc = one_step<T>(e, q, q, s, ade_sweep_capable(e) && q + 1 < G);
            if (!rc) { k_bump<<<1, 1, 0, s>>>(e->d_step, G); e->launches++; }
            cudaError_t ce = cudaStreamEndCapture(s, &graph);
            e->cur = c0;
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) return fail(FDTD_ECUDA, "graph capture: %s", cudaGetErrorString(ce));
            ce = cudaGraphInstantiate(&e->gexec[c0], graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) { e->gexec[c0] = nullptr; return fail(FDTD_ECUDA, "graph instantiate: %s", cudaGetErrorString(ce)); }
            e->graph_steps = G;
            e->graph_kernels[c0] = (int)(e->launches - l0);
            e->launches = l0;
        }
        while (n - done >= e->graph_steps) {
            CU(cudaGraphLaunch(e->gexec[c0], s));
            e->launches += e->graph_kernels[c0];
            done += e->graph_steps;
        }
    }
    const int rest = n - done;
    int q = 0;
    if (use_tb2(e))
        for (; q + 2 <= rest; q += 2)
            if (int rc = two_steps<T>(e, q, s)) return rc;
    for (; q < rest; ++q)
        if (int rc = one_step<T>(e, q, done + q, s, ade_sweep_capable(e) && q + 1 < rest)) return rc;
    if (rest > 0) { k_bump<<<1, 1, 0, s>>>(e->d_step, rest); e->launches++; CU(cudaGetLastError()); }
    return 0;
}
