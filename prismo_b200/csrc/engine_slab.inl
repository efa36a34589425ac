// peer-to-peer x-slabs: CUDA-IPC export / connect, fdtd_slab_run
// (textual part of fdtd_engine.cu — one translation unit; not compiled on its own)
// ---- peer-to-peer slabs ------------------------------------------------------------------------------------------
constexpr int kSeqLen = 1 << 20;          // exchanges per connect that publish their flag by DMA (then: a kernel)

struct IpcBlob {
    cudaIpcMemHandle_t fld[2][6];
    cudaIpcMemHandle_t flags;
    int32_t nx, ny, nz, dtype;
    int64_t plane_elems;
};

extern "C" int fdtd_ipc_export(fdtd_engine* e, void* blob, int32_t* nbytes)
{
    if (!e || !nbytes) return fail(FDTD_EINVAL, "fdtd_ipc_export: bad argument");
    if (!blob) { *nbytes = (int32_t)sizeof(IpcBlob); return 0; }
    if (*nbytes < (int32_t)sizeof(IpcBlob)) return fail(FDTD_EINVAL, "blob too small (%d < %d)", *nbytes, (int)sizeof(IpcBlob));
    if (!use_fused(e) && !het_sweep_ok(e))
        return fail(FDTD_ESTATE, "peer-to-peer slabs need a fused one-sweep path (3-D, no two-pass / physics flag)");
    CU(cudaSetDevice(e->cfg.device));
    if (int rc = ensure_set_b(e)) return rc;
    if (!e->slab.flags) {
        CU(cudaMalloc(&e->slab.flags, 64));
        CU(cudaMemset(e->slab.flags, 0, 64));
    }
    IpcBlob b;
    memset(&b, 0, sizeof b);
    for (int c = 0; c < 6; ++c) {
        CU(cudaIpcGetMemHandle(&b.fld[0][c], e->fld[c]));
        CU(cudaIpcGetMemHandle(&b.fld[1][c], e->fldB[c]));
    }
    CU(cudaIpcGetMemHandle(&b.flags, e->slab.flags));
    b.nx = e->g.nx; b.ny = e->g.ny; b.nz = e->g.nz; b.dtype = e->cfg.dtype; b.plane_elems = e->plane_elems;
    memcpy(blob, &b, sizeof b);
    *nbytes = (int32_t)sizeof b;
    return 0;
}

extern "C" int fdtd_ipc_connect(fdtd_engine* e, const void* left_blob, int32_t has_right)
{
    if (!e) return fail(FDTD_EINVAL, "null engine");
    CU(cudaSetDevice(e->cfg.device));
    auto& sl = e->slab;
    if (!sl.flags) return fail(FDTD_ESTATE, "call fdtd_ipc_export first");
    if (!sl.comm) {
        // The comm stream carries DMA only (plane copies + a 4-byte copy of the exchange number into the neighbour's
        // halo_ready word): nothing on it needs an SM slot while the sweep occupies every SM.
        CU(cudaStreamCreateWithFlags(&sl.comm, cudaStreamNonBlocking));
        std::vector<int> seq(kSeqLen);
        for (int i = 0; i < kSeqLen; ++i) seq[i] = i;
        CU(cudaMalloc(&sl.seq, sizeof(int) * kSeqLen));
        CU(cudaMemcpy(sl.seq, seq.data(), sizeof(int) * kSeqLen, cudaMemcpyHostToDevice));
        CU(cudaEventCreateWithFlags(&sl.post_done, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&sl.push_done, cudaEventDisableTiming));
    }
    sl.has_left = left_blob != nullptr;
    sl.has_right = has_right != 0;
    if (e->stream) CU(cudaStreamSynchronize(e->stream));
    CU(cudaStreamSynchronize(sl.comm));
    for (void*& m : sl.left_base)                    // re-connect: drop the previous neighbour's mappings first
        if (m) { cudaIpcCloseMemHandle(m); m = nullptr; }
    sl.left_flags = nullptr;
    if (left_blob) {
        IpcBlob b;
        memcpy(&b, left_blob, sizeof b);
        if (b.ny != e->g.ny || b.nz != e->g.nz || b.dtype != e->cfg.dtype || b.plane_elems != e->plane_elems)
            return fail(FDTD_EINVAL, "left neighbour has a different plane geometry / dtype");
        int n = 0;
        for (int s = 0; s < 2; ++s)
            for (int c = 0; c < 6; ++c) {
                CU(cudaIpcOpenMemHandle(&sl.left_fld[s][c], b.fld[s][c], cudaIpcMemLazyEnablePeerAccess));
                sl.left_base[n++] = sl.left_fld[s][c];
            }
        void* f = nullptr;
        CU(cudaIpcOpenMemHandle(&f, b.flags, cudaIpcMemLazyEnablePeerAccess));
        sl.left_flags = (int*)f;
        sl.left_base[n++] = f;
        sl.left_nx = b.nx;
    }
    if (const char* t = getenv("FDTD_B200_HALO_TIMEOUT_MS")) sl.timeout_ns = 1000000ull * (unsigned long long)atoll(t);
    if (e->cur != 0) return fail(FDTD_ESTATE, "connect slabs before stepping (buffer-set parity must agree across ranks)");
    sl.connected = true;
    sl.step = 0;
    CU(cudaMemset(sl.flags, 0, 64));
    return 0;
}

// n full steps of this slab, everything enqueued asynchronously (no host synchronisation inside):
//   comm stream    : DMA only — our first planes (7 per step, 21 per pair of steps) into the left neighbour's ghost
//                    planes over NVLink, then a 4-byte copy of the exchange number into its halo_ready word
//   compute stream : ONE fused sweep over all planes — only the CTAs of the ghost-reading x-segment wait (in-kernel)
//                    for our own halo_ready word —, publish ghost_consumed and wait until the left neighbour has
//                    consumed the ghosts our next push overwrites (1 thread, SMs idle), then sources + monitors
template <typename T> static int slab_run(fdtd_engine* e, int n)
{
    auto& sl = e->slab;
    // FDTD_B200_SLAB_DEBUG=1: print this rank's mean sweep duration and pair period to stderr (synchronises)
    static const bool dbg = getenv("FDTD_B200_SLAB_DEBUG") && atoi(getenv("FDTD_B200_SLAB_DEBUG"));
    std::vector<cudaEvent_t> dbg_ev;
    cudaStream_t cs = e->stream, ms = sl.comm;
    // single-step sweep: plane 0 of Ex Ey Ez Hy Hz + plane 1 of Ey Ez; two-step sweep: planes 0..3 of E, 0..2 of H
    static const int planes1[6] = {1, 2, 2, 0, 1, 1};
    static const int planes2[6] = {4, 4, 4, 3, 3, 3};
    const size_t pbytes = (size_t)e->plane_elems * e->esz;
    CU(cudaEventRecord(sl.post_done, cs));
    // The planes we push (0..3) are final as soon as the sweep has written them unless a source op of ours touches them:
    // then the push need not wait for this rank's sources / monitors (on the rank that owns a DFT plane that is 0.1-0.25 ms
    // per pair, which the left neighbour's ghost-reading segment would otherwise spend spinning: measured at 8 GPUs,
    // profiles/r02_tuning.md §6).
    bool early_push = true;
    for (const HostSrc& h : e->src) if (h.op.lo[0] < 4) early_push = false;
    std::vector<cudaEvent_t> push_ev;
    int q = 0;
    while (q < n) {
        const bool pair = tb2_ok(e) && q + 2 <= n;        // never for heterogeneous media (tb2_ok needs use_fused)
        const int* planes = pair ? planes2 : planes1;
        const long long st = sl.step;                    // exchange counter, identical on every rank
        if (sl.has_left) {
            CU(cudaStreamWaitEvent(ms, sl.post_done, 0));        // our first planes of the current set are final
            if (dbg) { push_ev.emplace_back(); cudaEventCreate(&push_ev.back()); cudaEventRecord(push_ev.back(), ms); }
            void** mine = cur_fields(e);
            void** theirs = sl.left_fld[e->cur];
            for (int c = 0; c < 6; ++c)
                if (planes[c])
                    CU(cudaMemcpyAsync((char*)theirs[c] + (size_t)sl.left_nx * pbytes, mine[c], planes[c] * pbytes,
                                       cudaMemcpyDefault, ms));
            if (st + 1 < kSeqLen)
                CU(cudaMemcpyAsync(sl.left_flags, sl.seq + (st + 1), sizeof(int), cudaMemcpyDefault, ms));
            else { k_signal<<<1, 1, 0, ms>>>(sl.left_flags, (int)(st + 1)); e->launches++; }
            CU(cudaEventRecord(sl.push_done, ms));
            if (dbg) { push_ev.emplace_back(); cudaEventCreate(&push_ev.back()); cudaEventRecord(push_ev.back(), ms); }
        }
        if (pair) {
            if (dbg) { dbg_ev.emplace_back(); cudaEventCreate(&dbg_ev.back()); cudaEventRecord(dbg_ev.back(), cs); }
            if (int rc = launch_tb2<T>(e, q, cs)) return rc;             // flips the sets itself
            if (dbg) { dbg_ev.emplace_back(); cudaEventCreate(&dbg_ev.back()); cudaEventRecord(dbg_ev.back(), cs); }
        } else if (e->het) {
            if (int rc = launch_het<T>(e, cs)) return rc;                 // heterogeneous media: flips the sets itself
        } else {
            if (int rc = launch_fused<T>(e, 0, e->g.nx, cs)) return rc;   // reads the current set (+ ghosts)
            e->cur ^= 1;
        }
        // ghosts consumed; and (for our NEXT push) wait until the left neighbour has finished the sweep that read the
        // ghost planes that push will overwrite — in order on the compute stream, when the SMs are idle anyway
        if (sl.has_left && st >= 1)
            k_signal_wait<<<1, 1, 0, cs>>>(sl.flags + 1, (int)(st + 1), sl.left_flags + 1, (int)st, sl.flags + 2, sl.timeout_ns);
        else
            k_signal<<<1, 1, 0, cs>>>(sl.flags + 1, (int)(st + 1));
        e->launches++;
        // the push read the set that is now the output set of the NEXT sweep: it must finish before that sweep
        if (sl.has_left) CU(cudaStreamWaitEvent(cs, sl.push_done, 0));
        q += pair ? 2 : 1;
        if (early_push) CU(cudaEventRecord(sl.post_done, cs));          // (the event the next push waits for)
        if (has_post(e)) if (int rc = launch_post<T>(e, q - 1, 0, cs)) return rc;
        if (!early_push) CU(cudaEventRecord(sl.post_done, cs));
        sl.step++;
    }
    k_bump<<<1, 1, 0, cs>>>(e->d_step, n); e->launches++;
    CU(cudaGetLastError());
    if (dbg_ev.size() >= 4) {
        cudaStreamSynchronize(cs);
        double kern = 0, period = 0;
        const size_t np = dbg_ev.size() / 2;
        for (size_t p = 0; p < np; ++p) {
            float ms = 0;
            cudaEventElapsedTime(&ms, dbg_ev[2 * p], dbg_ev[2 * p + 1]); kern += ms;
            if (p + 1 < np) { cudaEventElapsedTime(&ms, dbg_ev[2 * p], dbg_ev[2 * p + 2]); period += ms; }
        }
        double push = 0;
        cudaStreamSynchronize(ms);
        for (size_t p = 0; p + 1 < push_ev.size(); p += 2) { float t = 0; cudaEventElapsedTime(&t, push_ev[p], push_ev[p + 1]); push += t; }
        fprintf(stderr, "[fdtd dbg] dev %d pairs %zu sweep %.4f ms period %.4f ms push %.4f ms\n", e->cfg.device, np, kern / np,
                period / (np - 1), push_ev.empty() ? 0.0 : push / (push_ev.size() / 2));
        for (auto ev : dbg_ev) cudaEventDestroy(ev);
        for (auto ev : push_ev) cudaEventDestroy(ev);
    }
    return 0;
}

extern "C" int fdtd_slab_run(fdtd_engine* e, int32_t n_steps)
{
    if (!e || n_steps < 0) return fail(FDTD_EINVAL, "fdtd_slab_run: bad argument");
    if (!e->slab.connected) return fail(FDTD_ESTATE, "fdtd_slab_run: call fdtd_ipc_export / fdtd_ipc_connect first");
    if (n_steps == 0) return 0;
    CU(cudaSetDevice(e->cfg.device));
    if (int rc = finalize_ops(e)) return rc;
    if (has_tables(e)) {
        if (e->cursor + n_steps > e->n_steps_tab)
            return fail(FDTD_ESTATE, "fdtd_slab_run(%d): only %d tabled steps left", n_steps, e->n_steps_tab - e->cursor);
        if (!e->d_mon && !e->mon.empty()) if (int rc = upload_mon_ops(e)) return rc;
    }
    int rc = e->cfg.dtype == FDTD_F64 ? slab_run<double>(e, n_steps) : slab_run<float>(e, n_steps);
    if (rc) return rc;
    e->cursor += n_steps;
    e->steps_done += n_steps;
    return 0;
}

// blocks until everything enqueued by fdtd_slab_run is done; reports a halo time-out (dead peer)
extern "C" int fdtd_slab_sync(fdtd_engine* e)
{
    if (!e) return fail(FDTD_EINVAL, "null engine");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    if (e->slab.comm) CU(cudaStreamSynchronize(e->slab.comm));
    if (e->slab.flags) {
        int err = 0;
        CU(cudaMemcpy(&err, e->slab.flags + 2, sizeof(int), cudaMemcpyDeviceToHost));
        if (err) return fail(FDTD_ECUDA, "halo wait timed out: a neighbouring rank stopped making progress");
    }
    return 0;
}

extern "C" int fdtd_post_step(fdtd_engine* e, void* stream)
{
    if (!e) return fail(FDTD_EINVAL, "null engine");
    CU(cudaSetDevice(e->cfg.device));
    if (int rc = finalize_ops(e)) return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    if (has_tables(e)) {
        if (e->cursor + 1 > e->n_steps_tab) return fail(FDTD_ESTATE, "fdtd_post_step: no tabled steps left");
        if (!e->d_mon && !e->mon.empty()) if (int rc = upload_mon_ops(e)) return rc;
        int rc = e->cfg.dtype == FDTD_F64 ? launch_post<double>(e, 0, 0, s) : launch_post<float>(e, 0, 0, s);
        if (rc) return rc;
    }
    k_bump<<<1, 1, 0, s>>>(e->d_step, 1); e->launches++;
    CU(cudaGetLastError());
    e->cursor += 1; e->steps_done += 1;
    return 0;
}

extern "C" int fdtd_halo_ptrs(fdtd_engine* e, int32_t comp, void** first_plane, void** ghost_plane, int64_t* plane_bytes)
{
    if (!e || comp < 0 || comp > 5) return fail(FDTD_EINVAL, "fdtd_halo_ptrs: bad argument");
    char* base = (char*)cur_fields(e)[comp];
    if (first_plane) *first_plane = base;
    if (ghost_plane) *ghost_plane = base + (size_t)e->g.nx * e->plane_elems * e->esz;
    if (plane_bytes) *plane_bytes = (int64_t)(e->plane_elems * e->esz);
    return 0;
}
