// step loops: fdtd_run (CUDA graph), half steps, profiled run, options, split entry points
// (textual part of fdtd_engine.cu — one translation unit; not compiled on its own)
extern "C" int fdtd_run(fdtd_engine* e, int32_t n_steps)
{
    if (!e || n_steps < 0) return fail(FDTD_EINVAL, "fdtd_run: bad argument");
    if (n_steps == 0) return 0;
    if (e->g.nxg != e->g.nx)
        return fail(FDTD_ESTATE, "fdtd_run on an x-slab: drive slabs with fdtd_sweep / fdtd_pass + fdtd_post_step "
                                 "and exchange the halo planes in between");
    CU(cudaSetDevice(e->cfg.device));
    if (int rc = finalize_ops(e)) return rc;
    if (has_tables(e)) {
        if (e->cursor + n_steps > e->n_steps_tab)
            return fail(FDTD_ESTATE, "fdtd_run(%d): only %d tabled steps left (call fdtd_set_tables)", n_steps,
                        e->n_steps_tab - e->cursor);
        if (!e->d_mon && !e->mon.empty()) if (int rc = upload_mon_ops(e)) return rc;
    }
    int rc = e->cfg.dtype == FDTD_F64 ? run_steps<double>(e, n_steps) : run_steps<float>(e, n_steps);
    if (rc) return rc;
    e->cursor += n_steps;
    e->steps_done += n_steps;
    return 0;
}

static int single_pass(fdtd_engine* e, int phase)
{
    CU(cudaSetDevice(e->cfg.device));
    cudaStream_t s = e->stream;
    const bool d64 = e->cfg.dtype == FDTD_F64;
    if (e->cfg.ndim == 3)
        return d64 ? launch_pass3d<double>(e, phase, 0, e->g.nx, s) : launch_pass3d<float>(e, phase, 0, e->g.nx, s);
    if (phase == 0) {
        if (int rc = d64 ? launch_count2d<double>(e, 0, s) : launch_count2d<float>(e, 0, s)) return rc;
    }
    return d64 ? launch_pass2d<double>(e, phase, 0, s) : launch_pass2d<float>(e, phase, 0, s);
}
extern "C" int fdtd_update_h(fdtd_engine* e) { return e ? single_pass(e, 0) : fail(FDTD_EINVAL, "null engine"); }
extern "C" int fdtd_update_e(fdtd_engine* e) { return e ? single_pass(e, 1) : fail(FDTD_EINVAL, "null engine"); }

// K steps with CUDA events between the kernels of every step, on the engine's stream (no graph).
// out_ms[0] = sum of H-pass (or fused-step) kernel time, [1] = E-pass, [2] = sources+monitors, [3] = total
template <typename T> static int run_profiled(fdtd_engine* e, int n, double* out_ms)
{
    cudaStream_t s = e->stream;
    if (int rc = prepare_ade_order<T>(e)) return rc;
    std::vector<cudaEvent_t> ev((size_t)n * 3 + 1);
    for (auto& x : ev) CU(cudaEventCreate(&x));
    if (e->cfg.ndim == 2) if (int rc = launch_count2d<T>(e, 0, s)) return rc;
    CU(cudaEventRecord(ev[0], s));
    const bool tb2 = use_tb2(e);
    const int n_pairs = tb2 ? n / 2 * 2 : 0;              // an odd last step runs the one-step sweep
    for (int q = 0; q < n_pairs; q += 2) {
        // one temporally blocked sweep = two steps: its time goes to slot 0, step B's sources/monitors to slot 2
        int rc = launch_tb2<T>(e, q, s);
        if (rc) return rc;
        CU(cudaEventRecord(ev[3 * q + 1], s));
        CU(cudaEventRecord(ev[3 * q + 2], s));
        if ((rc = launch_post<T>(e, q + 1, 0, s))) return rc;
        for (int k = 3; k <= 6; ++k) CU(cudaEventRecord(ev[3 * q + k], s));
    }
    for (int q = n_pairs; q < n; ++q) {
        int rc;
        if (e->cfg.ndim == 3) rc = step_fields3d<T>(e, 0, s); else rc = launch_pass2d<T>(e, 0, q, s);
        if (rc) return rc;
        CU(cudaEventRecord(ev[3 * q + 1], s));
        if (e->cfg.ndim == 3) rc = step_fields3d<T>(e, 1, s); else rc = launch_pass2d<T>(e, 1, q, s);
        if (rc) return rc;
        CU(cudaEventRecord(ev[3 * q + 2], s));
        if ((rc = launch_post<T>(e, q, q, s, ade_sweep_capable(e) && q + 1 < n))) return rc;
        CU(cudaEventRecord(ev[3 * q + 3], s));
    }
    k_bump<<<1, 1, 0, s>>>(e->d_step, n); e->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s));
    out_ms[0] = out_ms[1] = out_ms[2] = 0;
    for (int q = 0; q < n; ++q)
        for (int k = 0; k < 3; ++k) {
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, ev[3 * q + k], ev[3 * q + k + 1]));
            out_ms[k] += ms;
        }
    float tot = 0;
    CU(cudaEventElapsedTime(&tot, ev[0], ev[(size_t)n * 3]));
    out_ms[3] = tot;
    for (auto& x : ev) cudaEventDestroy(x);
    return 0;
}

extern "C" int fdtd_run_profiled(fdtd_engine* e, int32_t n_steps, double* out_ms)
{
    if (!e || n_steps <= 0 || !out_ms) return fail(FDTD_EINVAL, "fdtd_run_profiled: bad argument");
    CU(cudaSetDevice(e->cfg.device));
    if (int rc = finalize_ops(e)) return rc;
    if (has_tables(e)) {
        if (e->cursor + n_steps > e->n_steps_tab)
            return fail(FDTD_ESTATE, "fdtd_run_profiled(%d): only %d tabled steps left", n_steps, e->n_steps_tab - e->cursor);
        if (!e->d_mon && !e->mon.empty()) if (int rc = upload_mon_ops(e)) return rc;
    }
    int rc = e->cfg.dtype == FDTD_F64 ? run_profiled<double>(e, n_steps, out_ms) : run_profiled<float>(e, n_steps, out_ms);
    if (rc) return rc;
    e->cursor += n_steps;
    e->steps_done += n_steps;
    return 0;
}

extern "C" int fdtd_timer_start(fdtd_engine* e)
{
    if (!e) return fail(FDTD_EINVAL, "null engine");
    CU(cudaSetDevice(e->cfg.device));
    if (!e->t0) { CU(cudaEventCreate(&e->t0)); CU(cudaEventCreate(&e->t1)); }
    CU(cudaEventRecord(e->t0, e->stream));
    return 0;
}
extern "C" int fdtd_timer_stop(fdtd_engine* e, double* ms)
{
    if (!e || !ms || !e->t0) return fail(FDTD_EINVAL, "fdtd_timer_stop: bad argument / timer not started");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaEventRecord(e->t1, e->stream));
    CU(cudaEventSynchronize(e->t1));
    float f = 0;
    CU(cudaEventElapsedTime(&f, e->t0, e->t1));
    *ms = f;
    return 0;
}

// tuning knobs, by name: "tb2" (0/1 two-step sweep), "fused_lx" (planes per x-segment, 0 = auto)
extern "C" int fdtd_set_option(fdtd_engine* e, const char* key, int32_t value)
{
    if (!e || !key) return fail(FDTD_EINVAL, "fdtd_set_option: null argument");
    if (!strcmp(key, "tb2")) e->tb2 = value ? 1 : 0;
    else if (!strcmp(key, "tb2x")) e->tb2x = value ? 1 : 0;
    else if (!strcmp(key, "tb2x_stages")) e->tb2x_stages = std::max(3, (int)value);
    else if (!strcmp(key, "tb2x_slots")) e->tb2x_slots = std::max(2, (int)value);
    else if (!strcmp(key, "het_fused")) e->het_fused = value ? 1 : 0;
    else if (!strcmp(key, "het_indexed")) e->het_indexed = value ? 1 : 0;
    else if (!strcmp(key, "fused_lx")) e->fused_lx = value;
    else if (!strcmp(key, "ade_fused")) e->ade_fused = value ? 1 : 0;
    else if (!strcmp(key, "ade_coupled")) e->ade_coupled = value ? 1 : 0;
    else if (!strcmp(key, "yee_fused")) {
        // the fused physics sweep ping-pongs fields and psi, the two-pass kernels update the current set in place:
        // switch only between runs that start from freshly uploaded / zeroed state
        e->yee_fused = value < 0 ? 0 : (value > 2 ? 2 : value);
    }
    else return fail(FDTD_EINVAL, "unknown option '%s'", key);
    drop_graph(e);
    return 0;
}

extern "C" int fdtd_sync(fdtd_engine* e)
{
    if (!e) return fail(FDTD_EINVAL, "null engine");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

// ---- multi-GPU split entry points ------------------------------------------------------------------------------
extern "C" int fdtd_pass(fdtd_engine* e, int32_t phase, int32_t part, void* stream)
{
    if (!e || phase < 0 || phase > 1 || part < 0 || part > 2) return fail(FDTD_EINVAL, "fdtd_pass: bad argument");
    if (e->cfg.ndim != 3) return fail(FDTD_EINVAL, "fdtd_pass is 3-D only");
    CU(cudaSetDevice(e->cfg.device));
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    const int nx = e->g.nx;
    int b = 0, t = nx;
    if (part == 0) t = nx - 1;
    if (part == 1) b = nx - 1;
    return e->cfg.dtype == FDTD_F64 ? launch_pass3d<double>(e, phase, b, t, s) : launch_pass3d<float>(e, phase, b, t, s);
}

// Fused sweep over local planes [i_begin, i_end) of the CURRENT set into the other set; flip != 0 makes the
// other set current afterwards (pass it on the last piece of a step).  For x-slabs: planes nx and nx+1 of the
// current set must hold the right neighbour's planes 0 and 1 (Ex,Ey,Ez,Hy,Hz / Ey,Ez) before the piece that
// contains plane nx-1 runs; H+ of the ghost plane is recomputed locally (fused-sweep variant).
extern "C" int fdtd_sweep(fdtd_engine* e, int32_t i_begin, int32_t i_end, int32_t flip, void* stream)
{
    if (!e || i_begin < 0 || i_end > e->g.nx || i_end < i_begin) return fail(FDTD_EINVAL, "fdtd_sweep: bad plane range");
    if (!use_fused(e)) return fail(FDTD_ESTATE, "fdtd_sweep needs a 3-D engine with uniform coefficients (fused path)");
    CU(cudaSetDevice(e->cfg.device));
    if (int rc = ensure_set_b(e)) return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    if (i_end > i_begin) {
        int rc = e->cfg.dtype == FDTD_F64 ? launch_fused<double>(e, i_begin, i_end, s) : launch_fused<float>(e, i_begin, i_end, s);
        if (rc) return rc;
    }
    if (flip) e->cur ^= 1;
    return 0;
}
