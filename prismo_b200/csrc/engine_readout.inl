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
