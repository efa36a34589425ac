// monitor read-out and introspection
// (textual part of fdtd_engine.cu — one translation unit; not compiled on its own)
// ---- monitor read-out ---------------------------------------------------------------------------------------------
extern "C" int fdtd_download_records(fdtd_engine* e, int32_t id, double* host, int32_t max_steps)
{
    if (!e || !host || id < 0 || id >= (int)e->mon.size()) return fail(FDTD_EINVAL, "fdtd_download_records: bad argument");
    const MonOp& m = e->mon[id];
    if (!m.record) return fail(FDTD_EINVAL, "monitor op %d does not record", id);
    CU(cudaSetDevice(e->cfg.device));
    const int steps = std::min<int>(max_steps, e->cursor);
    const long long n = (long long)steps * m.cells;
    if (n == 0) return 0;
    const long long off = m.rec_off * (long long)std::max(e->n_steps_tab, 1);
    CU(cudaStreamSynchronize(e->stream));
    if (e->cfg.dtype == FDTD_F64) {
        CU(cudaMemcpy(host, (const double*)e->d_rec + off, n * sizeof(double), cudaMemcpyDeviceToHost));
    } else {
        const long long chunk = std::min<long long>(n, (64ll << 20) / sizeof(double));
        if (int rc = ensure_stage(e, chunk * sizeof(double))) return rc;
        for (long long first = 0; first < n; first += chunk) {
            const long long c = std::min(chunk, n - first);
            k_convert<float, double><<<(int)std::min<long long>((c + 255) / 256, 148 * 16), 256, 0, e->stream>>>(
                (double*)e->d_stage, (const float*)e->d_rec + off + first, c);
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(host + first, e->d_stage, c * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
            CU(cudaStreamSynchronize(e->stream));
        }
    }
    return 0;
}

extern "C" int fdtd_download_dft(fdtd_engine* e, int32_t id, double* host)
{
    if (!e || !host || id < 0 || id >= (int)e->mon.size()) return fail(FDTD_EINVAL, "fdtd_download_dft: bad argument");
    CU(cudaSetDevice(e->cfg.device));
    if (int rc = finalize_ops(e)) return rc;
    const MonOp& m = e->mon[id];
    const long long n = (long long)m.n_freq * m.cells;
    if (n == 0) return 0;
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemcpy(host, e->d_dft + m.dft_off, n * sizeof(double2), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int fdtd_upload_dft(fdtd_engine* e, int32_t id, const double* host)
{
    if (!e || !host || id < 0 || id >= (int)e->mon.size()) return fail(FDTD_EINVAL, "fdtd_upload_dft: bad argument");
    CU(cudaSetDevice(e->cfg.device));
    if (int rc = finalize_ops(e)) return rc;
    const MonOp& m = e->mon[id];
    const long long n = (long long)m.n_freq * m.cells;
    if (n == 0) return 0;
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemcpy(e->d_dft + m.dft_off, host, n * sizeof(double2), cudaMemcpyHostToDevice));
    return 0;
}

// ---- cropped read-out and plane checksums (self-checking bench, parity tests at sizes the host cannot mirror) ---------
extern "C" int fdtd_download_box(fdtd_engine* e, int32_t comp, const int32_t* lo, const int32_t* hi, double* host)
{
    if (!e || !host || !lo || !hi || comp < 0 || comp > 5) return fail(FDTD_EINVAL, "fdtd_download_box: bad argument");
    CU(cudaSetDevice(e->cfg.device));
    int s[3]; comp_shape(e, comp, s);
    const int nd = e->cfg.ndim;
    int l[3] = {lo[0], lo[1], nd == 3 ? lo[2] : 0}, h[3] = {hi[0], hi[1], nd == 3 ? hi[2] : 1};
    for (int a = 0; a < 3; ++a)
        if (l[a] < 0 || h[a] > s[a] || h[a] < l[a])
            return fail(FDTD_EINVAL, "fdtd_download_box: box [%d,%d) outside axis %d of component %d (extent %d)", l[a], h[a], a, comp, s[a]);
    const long long n = (long long)(h[0] - l[0]) * (h[1] - l[1]) * (h[2] - l[2]);
    if (n == 0) return 0;
    if (n > (1ll << 27)) return fail(FDTD_EINVAL, "fdtd_download_box: box of %lld cells is too large (max 2^27)", n);
    if (int rc = ensure_stage(e, (size_t)n * sizeof(double))) return rc;
    const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 16);
    const void* src = cur_fields(e)[comp];
    if (e->cfg.dtype == FDTD_F64)
        k_gather_box<double><<<blocks, 256, 0, e->stream>>>((double*)e->d_stage, (const double*)src, n, l[0], l[1], l[2], h[1] - l[1], h[2] - l[2], e->st);
    else
        k_gather_box<float><<<blocks, 256, 0, e->stream>>>((double*)e->d_stage, (const float*)src, n, l[0], l[1], l[2], h[1] - l[1], h[2] - l[2], e->st);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host, e->d_stage, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

extern "C" int fdtd_field_checksum(fdtd_engine* e, int32_t comp, uint64_t* out, int32_t planes)
{
    if (!e || !out || comp < 0 || comp > 5) return fail(FDTD_EINVAL, "fdtd_field_checksum: bad argument");
    CU(cudaSetDevice(e->cfg.device));
    int s[3]; comp_shape(e, comp, s);
    if (planes != s[0]) return fail(FDTD_EINVAL, "fdtd_field_checksum: component %d has %d planes here, got room for %d", comp, s[0], planes);
    if (planes == 0) return 0;
    if (int rc = ensure_stage(e, (size_t)planes * 2 * sizeof(uint64_t))) return rc;
    CU(cudaMemsetAsync(e->d_stage, 0, (size_t)planes * 2 * sizeof(uint64_t), e->stream));
    const long long cells = (long long)s[1] * s[2];
    dim3 grid((unsigned)std::max<long long>(1, std::min<long long>((cells + 255) / 256, 64)), (unsigned)planes);
    const void* src = cur_fields(e)[comp];
    if (e->cfg.dtype == FDTD_F64)
        k_plane_checksum<double><<<grid, 256, 0, e->stream>>>((unsigned long long*)e->d_stage, (const double*)src, s[1], s[2], e->st);
    else
        k_plane_checksum<float><<<grid, 256, 0, e->stream>>>((unsigned long long*)e->d_stage, (const float*)src, s[1], s[2], e->st);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, e->d_stage, (size_t)planes * 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

// ---- mode overlap on the resident DFT planes --------------------------------------------------------------
extern "C" int fdtd_mode_overlap(fdtd_engine* e, const int32_t* monitor_ids, int32_t direction, const double* mode, double* out)
{
    if (!e || !monitor_ids || !mode || !out || direction < 0 || direction > 2)
        return fail(FDTD_EINVAL, "fdtd_mode_overlap: bad argument");
    CU(cudaSetDevice(e->cfg.device));
    if (int rc = finalize_ops(e)) return rc;
    // tangential components (e1, h2, e2, h1) of  S_n = e1 x conj(h2) - e2 x conj(h1)   (mode_matching.py:104-115)
    static const int pick[3][4] = {{FDTD_EY, FDTD_HZ, FDTD_EZ, FDTD_HY}, {FDTD_EZ, FDTD_HX, FDTD_EX, FDTD_HZ},
                                   {FDTD_EX, FDTD_HY, FDTD_EY, FDTD_HX}};
    OverlapIn in{};
    int nf = -1;
    for (int q = 0; q < 4; ++q) {
        const int id = monitor_ids[pick[direction][q]];
        if (id < 0 || id >= (int)e->mon.size()) return fail(FDTD_EINVAL, "fdtd_mode_overlap: monitor id %d out of range", id);
        const MonOp& m = e->mon[id];
        if (m.comp != pick[direction][q]) return fail(FDTD_EINVAL, "fdtd_mode_overlap: monitor %d samples component %d, expected %d", id, m.comp, pick[direction][q]);
        if (m.n_freq <= 0) return fail(FDTD_EINVAL, "fdtd_mode_overlap: monitor %d has no running DFT", id);
        if (q == 0) { in.cells = m.cells; nf = m.n_freq; }
        if (m.cells != in.cells || m.n_freq != nf) return fail(FDTD_EINVAL, "fdtd_mode_overlap: the monitors do not share one box / frequency list");
        in.off[q] = m.dft_off;
    }
    const size_t mode_bytes = (size_t)4 * in.cells * sizeof(double2);
    const size_t part_bytes = (size_t)nf * FLUX_BLOCKS * sizeof(double2), out_bytes = (size_t)nf * sizeof(double2);
    if (int rc = ensure_stage(e, mode_bytes + part_bytes + out_bytes)) return rc;
    char* base = (char*)e->d_stage;
    // mode: host (6, cells) complex128 in component order Ex..Hz; only the four tangential planes travel
    for (int q = 0; q < 4; ++q)
        CU(cudaMemcpyAsync(base + (size_t)q * in.cells * sizeof(double2), mode + (size_t)pick[direction][q] * in.cells * 2,
                           (size_t)in.cells * sizeof(double2), cudaMemcpyHostToDevice, e->stream));
    double2* partial = (double2*)(base + mode_bytes);
    double2* d_out = (double2*)(base + mode_bytes + part_bytes);
    dim3 grid(FLUX_BLOCKS, (unsigned)nf);
    k_overlap_partial<<<grid, 256, 0, e->stream>>>(e->d_dft, in, (const double2*)base, partial);
    k_overlap_final<<<(nf + 63) / 64, 64, 0, e->stream>>>(partial, nf, d_out);
    e->launches += 2;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

// ---- introspection ---------------------------------------------------------------------------------------------------
extern "C" int fdtd_steps_done(fdtd_engine* e, int64_t* steps)
{
    if (!e || !steps) return fail(FDTD_EINVAL, "bad argument");
    *steps = e->steps_done;
    return 0;
}
extern "C" int fdtd_kernel_launches(fdtd_engine* e, int64_t* launches)
{
    if (!e || !launches) return fail(FDTD_EINVAL, "bad argument");
    *launches = e->launches;
    return 0;
}
extern "C" int fdtd_mem_info(fdtd_engine* e, int64_t* free_bytes, int64_t* total_bytes)
{
    if (!e) return fail(FDTD_EINVAL, "null engine");
    CU(cudaSetDevice(e->cfg.device));
    size_t f = 0, t = 0;
    CU(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = (int64_t)f;
    if (total_bytes) *total_bytes = (int64_t)t;
    return 0;
}

// ---- device-level services for the backend object (backends/base.py:188-219: synchronize, get_memory_info) and pinned
// host mirrors (array factories hand out page-locked NumPy memory so uploads are plain DMA) ------------------
extern "C" int fdtd_device_mem_info(int32_t device, int64_t* free_bytes, int64_t* total_bytes)
{
    CU(cudaSetDevice(device));
    size_t f = 0, t = 0;
    CU(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = (int64_t)f;
    if (total_bytes) *total_bytes = (int64_t)t;
    return 0;
}
extern "C" int fdtd_device_sync(int32_t device)
{
    CU(cudaSetDevice(device));
    CU(cudaDeviceSynchronize());
    return 0;
}
extern "C" int fdtd_host_alloc(int64_t bytes, void** ptr)
{
    if (!ptr || bytes <= 0) return fail(FDTD_EINVAL, "fdtd_host_alloc: bad argument");
    CU(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocPortable));
    return 0;
}
extern "C" int fdtd_host_free(void* ptr)
{
    if (ptr) CU(cudaFreeHost(ptr));
    return 0;
}
