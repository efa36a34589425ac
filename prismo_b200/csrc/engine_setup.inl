// ABI version, create / destroy, coefficients, field upload / download
// (textual part of fdtd_engine.cu — one translation unit; not compiled on its own)
extern "C" int fdtd_abi_version(void) { return FDTD_B200_ABI_VERSION; }
extern "C" const char* fdtd_last_error(void) { return g_err.c_str(); }
extern "C" int fdtd_struct_size(int32_t which)
{
    switch (which) {
    case 0: return (int)sizeof(fdtd_config);
    case 1: return (int)sizeof(fdtd_source_op);
    case 2: return (int)sizeof(fdtd_monitor_op);
    case 3: return (int)sizeof(fdtd_ade_op);
    case 4: return (int)sizeof(fdtd_shape);
    default: return -1;
    }
}

extern "C" int fdtd_create(const fdtd_config* cfg, fdtd_engine** out)
{
    if (!cfg || !out) return fail(FDTD_EINVAL, "fdtd_create: null argument");
    if (cfg->ndim != 2 && cfg->ndim != 3) return fail(FDTD_EINVAL, "ndim must be 2 or 3, got %d", cfg->ndim);
    if (cfg->nx < 3 || cfg->ny < 3 || (cfg->ndim == 3 && cfg->nz < 3))
        return fail(FDTD_EINVAL, "grid %dx%dx%d too small (need >= 3 cells per axis)", cfg->nx, cfg->ny, cfg->nz);
    if (cfg->dtype != FDTD_F32 && cfg->dtype != FDTD_F64) return fail(FDTD_EINVAL, "bad dtype %d", cfg->dtype);
    if (!(cfg->dx > 0) || !(cfg->dy > 0) || (cfg->ndim == 3 && !(cfg->dz > 0)) || !(cfg->dt > 0))
        return fail(FDTD_EINVAL, "spacings and dt must be positive");
    const int nxg = cfg->nx_global > 0 ? cfg->nx_global : cfg->nx;
    if (cfg->x_offset < 0 || cfg->x_offset + cfg->nx > nxg)
        return fail(FDTD_EINVAL, "slab [%d,%d) outside global nx=%d", cfg->x_offset, cfg->x_offset + cfg->nx, nxg);
    if (cfg->ndim == 2 && nxg != cfg->nx) return fail(FDTD_EINVAL, "2-D grids are not slab-decomposed");
    if (cfg->dtype == FDTD_F64) {
        const double dd[3] = {cfg->dx, cfg->dy, cfg->ndim == 3 ? cfg->dz : cfg->dx};
        for (double d : dd)
            if (!(d >= 0x1p-60 && d <= 0x1p60))
                return fail(FDTD_EINVAL, "fp64 engines take spacings in [2^-60, 2^60] (exact division by a constant), got %g", d);
    }
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(FDTD_EINVAL, "device %d not in [0,%d)", cfg->device, ndev);
    CU(cudaSetDevice(cfg->device));

    fdtd_engine* e = new fdtd_engine();
    e->cfg = *cfg;
    e->cfg.nx_global = nxg;
    Geom& g = e->g;
    g.nx = cfg->nx; g.ny = cfg->ny; g.nz = cfg->ndim == 3 ? cfg->nz : 1;
    g.nxg = nxg; g.x0 = cfg->x_offset;
    g.dx = cfg->dx; g.dy = cfg->dy; g.dz = cfg->ndim == 3 ? cfg->dz : 0.0;
    g.rdx = make_rcp(g.dx); g.rdy = make_rcp(g.dy); g.rdz = cfg->ndim == 3 ? make_rcp(g.dz) : Rcp{0.f, 0.0};
    if (cfg->ndim == 3) {
        g.pz = (int)round_up(g.nz, 32);
        g.sy = g.pz; g.sx = (long long)g.ny * g.pz;
        e->st.s[0] = g.sx; e->st.s[1] = g.sy; e->st.s[2] = 1;
    } else {
        g.pz = (int)round_up(g.ny, 32);
        g.sy = 1; g.sx = g.pz;
        e->st.s[0] = g.sx; e->st.s[1] = 1; e->st.s[2] = 0;
    }
    e->esz = cfg->dtype == FDTD_F64 ? 8 : 4;
    e->plane_elems = g.sx;
    e->planes_alloc = g.nx + 4;         // data + up to 4 ghost/guard planes (two-step sweep reads E0 up to plane nx+3)
    e->array_elems = e->plane_elems * e->planes_alloc;

    cudaError_t ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) { delete e; return fail(FDTD_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(ce)); }
    for (int c = 0; c < 6; ++c) {
        ce = cudaMalloc(&e->fld[c], e->array_elems * e->esz);
        if (ce == cudaSuccess) ce = cudaMemsetAsync(e->fld[c], 0, e->array_elems * e->esz, e->stream);
        if (ce != cudaSuccess) {
            fdtd_destroy(e);
            return fail(ce == cudaErrorMemoryAllocation ? FDTD_ENOMEM : FDTD_ECUDA,
                        "allocating field %d (%lld bytes): %s", c, (long long)(e->array_elems * e->esz),
                        cudaGetErrorString(ce));
        }
    }
    ce = cudaMalloc(&e->d_step, sizeof(int));
    if (ce == cudaSuccess) ce = cudaMemsetAsync(e->d_step, 0, sizeof(int), e->stream);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_cnt, 12 * sizeof(int));
    if (ce == cudaSuccess) ce = cudaMemsetAsync(e->d_cnt, 0, 12 * sizeof(int), e->stream);
    for (int s = 0; s < 2 && ce == cudaSuccess; ++s) ce = cudaMalloc((void**)&e->d_comp_ptr[s], 6 * sizeof(void*));
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(e->d_comp_ptr[0], e->fld, 6 * sizeof(void*), cudaMemcpyHostToDevice, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) { fdtd_destroy(e); return fail(FDTD_ECUDA, "engine setup: %s", cudaGetErrorString(ce)); }
    // vacuum defaults (solver.py:84-97, :113-133): Ca = Da = 1, Cb = dt/eps0, Db = dt/mu0
    const double eps0 = 8.854187817e-12, mu0 = 4 * M_PI * 1e-7;
    e->uni[0] = 1.0; e->uni[1] = cfg->dt / eps0; e->uni[2] = 1.0; e->uni[3] = cfg->dt / mu0;
    if (const char* lx = getenv("FDTD_B200_FUSED_LX")) e->fused_lx = atoi(lx);    // tuning / tests
    if (const char* tb = getenv("FDTD_B200_TB2")) e->tb2 = atoi(tb);
    if (const char* z = getenv("FDTD_B200_TB2_ZONES")) e->tb2_zones = atoi(z);
    if (const char* x = getenv("FDTD_B200_TB2X")) e->tb2x = atoi(x);
    if (const char* x = getenv("FDTD_B200_TB2X_STAGES")) e->tb2x_stages = std::max(3, atoi(x));
    if (const char* x = getenv("FDTD_B200_TB2X_SLOTS")) e->tb2x_slots = std::max(2, atoi(x));
    if (const char* yf = getenv("FDTD_B200_YEE_FUSED")) e->yee_fused = atoi(yf);
    if (const char* x = getenv("FDTD_B200_YEEX_STAGES")) e->yeex_stages = std::max(2, atoi(x));
    if (const char* x = getenv("FDTD_B200_YEEX_SLOTS")) e->yeex_slots = std::max(2, atoi(x));
    if (const char* hf = getenv("FDTD_B200_HET_FUSED")) e->het_fused = atoi(hf);
    if (const char* hi = getenv("FDTD_B200_HET_INDEXED")) e->het_indexed = atoi(hi);
    if (const char* af = getenv("FDTD_B200_ADE_FUSED")) e->ade_fused = atoi(af);
    *out = e;
    return 0;
}

extern "C" int fdtd_destroy(fdtd_engine* e)
{
    if (!e) return 0;
    cudaSetDevice(e->cfg.device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    drop_graph(e);
    for (int c = 0; c < 6; ++c) { cudaFree(e->fld[c]); cudaFree(e->fldB[c]); }
    for (int c = 0; c < 6; ++c) cudaFree(e->coef[c]);
    cudaFree(e->mat); cudaFree(e->mat_tab);
    cudaFree(e->d_src); cudaFree(e->d_mon); cudaFree(e->d_prof); cudaFree(e->d_src_ghost);
    cudaFree(e->d_ade); cudaFree(e->d_aux); cudaFree(e->d_ade_mask);
    cudaFree(e->d_flux); cudaFree(e->d_flux_partial); cudaFree(e->d_flux_out);
    cudaFree(e->d_cpml_coef); cudaFree(e->d_cpml_coef_f); cudaFree(e->d_plane_flags); cudaFree(e->d_ade_order);
    for (int q = 0; q < 12; ++q) { cudaFree(e->cpml.psi[q]); cudaFree(e->psiB[q]); }
    cudaFree(e->d_comp_ptr[0]); cudaFree(e->d_comp_ptr[1]);
    cudaFree(e->d_amp); cudaFree(e->d_phasor); cudaFree(e->d_rec); cudaFree(e->d_dft);
    cudaFree(e->d_step); cudaFree(e->d_cnt); cudaFree(e->d_stage);
    fused_release(e->fused);
    if (e->t0) { cudaEventDestroy(e->t0); cudaEventDestroy(e->t1); }
    if (e->slab.comm) { cudaStreamSynchronize(e->slab.comm); cudaStreamDestroy(e->slab.comm); }
    if (e->slab.seq) cudaFree(e->slab.seq);
    if (e->slab.post_done) { cudaEventDestroy(e->slab.post_done); cudaEventDestroy(e->slab.push_done); }
    for (void* b : e->slab.left_base) if (b) cudaIpcCloseMemHandle(b);
    cudaFree(e->slab.flags);
    if (e->xfer2) { cudaStreamDestroy(e->xfer2); cudaEventDestroy(e->xfer_ev); }
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return 0;
}

// ---- coefficients -----------------------------------------------------------------------------------
extern "C" int fdtd_set_uniform_coeffs(fdtd_engine* e, double ca, double cb, double da, double db)
{
    if (!e) return fail(FDTD_EINVAL, "null engine");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    for (int c = 0; c < 6; ++c) { cudaFree(e->coef[c]); e->coef[c] = nullptr; }
    e->aniso = false;
    e->n_mat = 0;
    e->het = false;
    e->uni[0] = ca; e->uni[1] = cb; e->uni[2] = da; e->uni[3] = db;
    drop_graph(e);
    return 0;
}

// same dtype: one strided DMA between the caller's (compact) buffer and the padded device array
template <typename T> static int copy_staged(fdtd_engine* e, T* dev, T* host, long long c0, int c1, int c2, bool to_device);
static int copy_strided(fdtd_engine* e, void* dev, void* host, long long c0, int c1, int c2, bool to_device)
{
    if (c0 * c1 * c2 == 0) return 0;
    const size_t esz = e->esz;
    static const bool staged = !(getenv("FDTD_B200_STAGED_COPY") && atoi(getenv("FDTD_B200_STAGED_COPY")) == 0);
    if (staged && e->cfg.ndim == 3 && (size_t)(c0 * c1 * c2) * esz >= (64u << 20)) {
        if (esz == 8) return copy_staged<double>(e, (double*)dev, (double*)host, c0, c1, c2, to_device);
        return copy_staged<float>(e, (float*)dev, (float*)host, c0, c1, c2, to_device);
    }
    if (e->cfg.ndim == 3) {
        cudaMemcpy3DParms p = {};
        cudaPitchedPtr h = make_cudaPitchedPtr(host, (size_t)c2 * esz, (size_t)c2 * esz, (size_t)c1);
        cudaPitchedPtr d = make_cudaPitchedPtr(dev, (size_t)e->g.pz * esz, (size_t)e->g.pz * esz, (size_t)e->g.ny);
        p.srcPtr = to_device ? h : d;
        p.dstPtr = to_device ? d : h;
        p.extent = make_cudaExtent((size_t)c2 * esz, (size_t)c1, (size_t)c0);
        p.kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
        CU(cudaMemcpy3DAsync(&p, e->stream));
    } else {
        if (to_device)
            CU(cudaMemcpy2DAsync(dev, (size_t)e->g.sx * esz, host, (size_t)c1 * esz, (size_t)c1 * esz, (size_t)c0,
                                 cudaMemcpyHostToDevice, e->stream));
        else
            CU(cudaMemcpy2DAsync(host, (size_t)c1 * esz, dev, (size_t)e->g.sx * esz, (size_t)c1 * esz, (size_t)c0,
                                 cudaMemcpyDeviceToHost, e->stream));
    }
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

// same dtype, large 3-D arrays: the compact host array travels as contiguous 32 MB chunks through two staging buffers on
// two streams (chunk q+1 is on the copy engine while chunk q is scattered into the padded array by the SMs), instead of
// one strided DMA with a descriptor per 4 KB row (measured: 32 GB/s for cudaMemcpy3D on 1024^3 fp32 against a PCIe 5
// x16 link).  Needs page-locked host memory to overlap; pageable memory degrades to synchronous chunks, still correct.
template <typename T> static int copy_staged(fdtd_engine* e, T* dev, T* host, long long c0, int c1, int c2, bool to_device)
{
    const long long total = c0 * c1 * c2;
    const long long chunk = (32ll << 20) / (long long)sizeof(T);
    if (int rc = ensure_stage(e, 2 * (size_t)chunk * sizeof(T))) return rc;
    if (!e->xfer2) {
        CU(cudaStreamCreateWithFlags(&e->xfer2, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&e->xfer_ev, cudaEventDisableTiming));
    }
    CU(cudaEventRecord(e->xfer_ev, e->stream));              // everything queued so far (memset of the target, last step)
    CU(cudaStreamWaitEvent(e->xfer2, e->xfer_ev, 0));
    int q = 0;
    for (long long first = 0; first < total; first += chunk, ++q) {
        const long long n = std::min(chunk, total - first);
        cudaStream_t s = (q & 1) ? e->xfer2 : e->stream;     // a staging half is only ever used on one stream: ordered
        T* st = (T*)e->d_stage + (size_t)(q & 1) * chunk;
        const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 16);
        if (to_device) {
            CU(cudaMemcpyAsync(st, host + first, (size_t)n * sizeof(T), cudaMemcpyHostToDevice, s));
            k_scatter<T, T><<<blocks, 256, 0, s>>>(dev, st, first, n, c1, c2, e->st);
        } else {
            k_gather<T, T><<<blocks, 256, 0, s>>>(st, dev, first, n, c1, c2, e->st);
            CU(cudaMemcpyAsync(host + first, st, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, s));
        }
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(e->xfer2));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

template <typename TD, typename TH>
static int scatter_host(fdtd_engine* e, TD* dst, const TH* host, long long c0, int c1, int c2)
{
    const long long total = c0 * c1 * c2;
    const long long chunk = std::min<long long>(total, (64ll << 20) / sizeof(TH));
    if (total == 0) return 0;
    if (int rc = ensure_stage(e, chunk * sizeof(TH))) return rc;
    for (long long first = 0; first < total; first += chunk) {
        const long long n = std::min(chunk, total - first);
        CU(cudaMemcpyAsync(e->d_stage, host + first, n * sizeof(TH), cudaMemcpyHostToDevice, e->stream));
        const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 16);
        k_scatter<TD, TH><<<blocks, 256, 0, e->stream>>>(dst, (const TH*)e->d_stage, first, n, c1, c2, e->st);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(e->stream));   // staging buffer is reused
    }
    return 0;
}
template <typename TD, typename TH>
static int gather_host(fdtd_engine* e, TH* host, const TD* src, long long c0, int c1, int c2)
{
    const long long total = c0 * c1 * c2;
    const long long chunk = std::min<long long>(total, (64ll << 20) / sizeof(TH));
    if (total == 0) return 0;
    if (int rc = ensure_stage(e, chunk * sizeof(TH))) return rc;
    for (long long first = 0; first < total; first += chunk) {
        const long long n = std::min(chunk, total - first);
        const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 16);
        k_gather<TD, TH><<<blocks, 256, 0, e->stream>>>((TH*)e->d_stage, src, first, n, c1, c2, e->st);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(host + first, e->d_stage, n * sizeof(TH), cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    return 0;
}

static int set_coef_arrays(fdtd_engine* e, const double* const* src, int n_arrays, int32_t planes, const char* who)
{
    if (planes != e->g.nx && planes != e->g.nx + 1)
        return fail(FDTD_EINVAL, "%s: coefficient arrays must have nx=%d (or nx+1) planes, got %d", who, e->g.nx, planes);
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    for (int c = 0; c < n_arrays; ++c) {
        if (!e->coef[c]) CU(cudaMalloc(&e->coef[c], e->array_elems * e->esz));
        CU(cudaMemsetAsync(e->coef[c], 0, e->array_elems * e->esz, e->stream));
        int rc;
        const int c1 = e->g.ny, c2 = e->cfg.ndim == 3 ? e->g.nz : 1;
        if (e->cfg.dtype == FDTD_F64) rc = scatter_host<double, double>(e, (double*)e->coef[c], src[c], planes, c1, c2);
        else rc = scatter_host<float, double>(e, (float*)e->coef[c], src[c], planes, c1, c2);
        if (rc) return rc;
    }
    for (int c = n_arrays; c < 6; ++c) { cudaFree(e->coef[c]); e->coef[c] = nullptr; }
    e->het = true;
    e->aniso = n_arrays == 6;
    e->n_mat = 0;                   // host arrays: no index coding
    e->coef_planes = planes;
    drop_graph(e);
    return 0;
}

extern "C" int fdtd_set_coeffs(fdtd_engine* e, const double* ca, const double* cb, const double* da,
                               const double* db, int32_t planes)
{
    if (!e || !ca || !cb || !da || !db) return fail(FDTD_EINVAL, "fdtd_set_coeffs: null argument");
    const double* src[4] = {ca, cb, da, db};
    return set_coef_arrays(e, src, 4, planes, "fdtd_set_coeffs");
}

// per-component Cb: the parity sweeps of 3-D grids only (the physics-mode and 2-D kernels have one Cb per cell)
static int aniso_supported(const fdtd_engine* e, const char* who)
{
    if (e->cfg.ndim != 3) return fail(FDTD_EINVAL, "%s: per-component Cb needs a 3-D grid", who);
    if (e->cfg.flags & FDTD_FLAG_YEE) return fail(FDTD_EINVAL, "%s: per-component Cb is not available in physics mode", who);
    return 0;
}

extern "C" int fdtd_set_coeffs_aniso(fdtd_engine* e, const double* ca, const double* cbx, const double* cby,
                                     const double* cbz, const double* da, const double* db, int32_t planes)
{
    if (!e || !ca || !cbx || !cby || !cbz || !da || !db) return fail(FDTD_EINVAL, "fdtd_set_coeffs_aniso: null argument");
    if (int rc = aniso_supported(e, "fdtd_set_coeffs_aniso")) return rc;
    const double* src[6] = {ca, cbx, da, db, cby, cbz};
    return set_coef_arrays(e, src, 6, planes, "fdtd_set_coeffs_aniso");
}

// ---- fields ---------------------------------------------------------------------------------------------
extern "C" int fdtd_upload_field(fdtd_engine* e, int32_t comp, const void* host, int32_t host_dtype)
{
    if (!e || !host || comp < 0 || comp > 5) return fail(FDTD_EINVAL, "fdtd_upload_field: bad argument");
    CU(cudaSetDevice(e->cfg.device));
    int s[3]; comp_shape(e, comp, s);
    void* dst = cur_fields(e)[comp];
    // On an x-slab the planes from nx on are the right neighbour's: it pushes into them as soon as its own
    // fdtd_slab_run starts, with no handshake against this upload.  Clear the owned planes only and leave the ghost
    // planes to the halo protocol (callers still order "all uploads" before "any run" with a barrier, see the header).
    const long long clear_planes = (e->g.nxg != e->g.nx || e->slab.connected) ? e->g.nx : e->planes_alloc;
    CU(cudaMemsetAsync(dst, 0, (size_t)clear_planes * e->plane_elems * e->esz, e->stream));
    const bool d64 = e->cfg.dtype == FDTD_F64, h64 = host_dtype == FDTD_F64;
    if (host_dtype != FDTD_F32 && host_dtype != FDTD_F64) return fail(FDTD_EINVAL, "bad host dtype %d", host_dtype);
    if (d64 == h64) return copy_strided(e, dst, const_cast<void*>(host), s[0], s[1], s[2], true);
    if (d64 && !h64) return scatter_host<double, float>(e, (double*)dst, (const float*)host, s[0], s[1], s[2]);
    if (!d64 && h64) return scatter_host<float, double>(e, (float*)dst, (const double*)host, s[0], s[1], s[2]);
    return scatter_host<float, float>(e, (float*)dst, (const float*)host, s[0], s[1], s[2]);
}

extern "C" int fdtd_download_field(fdtd_engine* e, int32_t comp, void* host, int32_t host_dtype)
{
    if (!e || !host || comp < 0 || comp > 5) return fail(FDTD_EINVAL, "fdtd_download_field: bad argument");
    CU(cudaSetDevice(e->cfg.device));
    int s[3]; comp_shape(e, comp, s);
    const void* src = cur_fields(e)[comp];
    const bool d64 = e->cfg.dtype == FDTD_F64, h64 = host_dtype == FDTD_F64;
    if (host_dtype != FDTD_F32 && host_dtype != FDTD_F64) return fail(FDTD_EINVAL, "bad host dtype %d", host_dtype);
    if (d64 == h64) return copy_strided(e, const_cast<void*>(src), host, s[0], s[1], s[2], false);
    if (d64 && !h64) return gather_host<double, float>(e, (float*)host, (const double*)src, s[0], s[1], s[2]);
    if (!d64 && h64) return gather_host<float, double>(e, (double*)host, (const float*)src, s[0], s[1], s[2]);
    return gather_host<float, float>(e, (float*)host, (const float*)src, s[0], s[1], s[2]);
}

extern "C" int fdtd_zero_fields(fdtd_engine* e)
{
    if (!e) return fail(FDTD_EINVAL, "null engine");
    CU(cudaSetDevice(e->cfg.device));
    for (int c = 0; c < 6; ++c) CU(cudaMemsetAsync(cur_fields(e)[c], 0, e->array_elems * e->esz, e->stream));
    for (int q = 0; q < 12; ++q) {
        if (e->cpml.psi[q]) CU(cudaMemsetAsync(e->cpml.psi[q], 0, e->psi_bytes[q], e->stream));
        if (e->psiB[q]) CU(cudaMemsetAsync(e->psiB[q], 0, e->psi_bytes[q], e->stream));
    }
    return 0;
}

extern "C" int fdtd_field_device_ptr(fdtd_engine* e, int32_t comp, void** ptr, int64_t* plane_stride,
                                     int64_t* row_stride, int64_t* planes_allocated)
{
    if (!e || comp < 0 || comp > 5) return fail(FDTD_EINVAL, "fdtd_field_device_ptr: bad argument");
    if (ptr) *ptr = cur_fields(e)[comp];
    if (plane_stride) *plane_stride = e->g.sx;
    if (row_stride) *row_stride = e->cfg.ndim == 3 ? e->g.sy : 1;
    if (planes_allocated) *planes_allocated = e->planes_alloc;
    return 0;
}
