// source / monitor / flux / ADE ops, tables
// (textual part of fdtd_engine.cu — one translation unit; not compiled on its own)
// ---- ops ------------------------------------------------------------------------------------------------
static int check_box(const fdtd_engine* e, int comp, const int32_t* lo, const int32_t* hi, int n[3])
{
    if (comp < 0 || comp > 5) return fail(FDTD_EINVAL, "component %d out of range", comp);
    int s[3]; comp_shape(e, comp, s);
    for (int a = 0; a < 3; ++a) {
        if (lo[a] < 0 || hi[a] > s[a] || hi[a] < lo[a])
            return fail(FDTD_EINVAL, "box [%d,%d) outside axis %d extent %d of component %d", lo[a], hi[a], a, s[a], comp);
        n[a] = hi[a] - lo[a];
    }
    return 0;
}

extern "C" int fdtd_clear_ops(fdtd_engine* e)
{
    if (!e) return fail(FDTD_EINVAL, "null engine");
    e->src.clear(); e->mon.clear(); e->prof_host.clear(); e->src_ghost.clear();
    e->ade.clear(); e->ade_mask_host.clear(); e->flux.clear();
    e->ade_epoch++;
    e->ops_dirty = true;
    drop_graph(e);
    return 0;
}

extern "C" int fdtd_add_source_op(fdtd_engine* e, const fdtd_source_op* op)
{
    if (!e || !op) return fail(FDTD_EINVAL, "fdtd_add_source_op: null argument");
    HostSrc h{};
    if (op->reserved & 1) {
        // ghost op (x-slabs, two-step sweep): the right neighbour's injection on our ghost planes [nx, nx+3),
        // applied only to the intermediate step inside the sweep; never by the post-step kernel
        if (op->component < 0 || op->component > 5) return fail(FDTD_EINVAL, "component %d out of range", op->component);
        if (op->lo[0] < e->g.nx || op->hi[0] > e->g.nx + 3 || op->hi[0] < op->lo[0])
            return fail(FDTD_EINVAL, "ghost source op must lie in planes [nx, nx+3)");
        if (op->profile) return fail(FDTD_EINVAL, "ghost source ops are uniform (no profile)");
        SrcOp g{};
        g.comp = op->component; g.table = op->table; g.divisor = 1.0; g.prof_off = -1;
        for (int a = 0; a < 3; ++a) { g.lo[a] = op->lo[a]; g.n[a] = op->hi[a] - op->lo[a]; }
        if (g.n[0] > 0 && g.n[1] > 0 && g.n[2] > 0) e->src_ghost.push_back(g);
        e->ops_dirty = true;
        drop_graph(e);
        return 0;
    }
    if (int rc = check_box(e, op->component, op->lo, op->hi, h.op.n)) return rc;
    if (op->table < 0) return fail(FDTD_EINVAL, "negative table index");
    h.op.comp = op->component;
    for (int a = 0; a < 3; ++a) h.op.lo[a] = op->lo[a];
    h.op.table = op->table;
    h.op.divisor = op->divisor;
    h.op.prof_off = -1;
    const long long cells = (long long)h.op.n[0] * h.op.n[1] * h.op.n[2];
    if (op->profile && cells > 0) {
        h.op.prof_off = (long long)e->prof_host.size();
        e->prof_host.insert(e->prof_host.end(), op->profile, op->profile + cells);
    }
    h.group = op->group;
    if (cells > 0) e->src.push_back(h);
    e->ops_dirty = true;
    drop_graph(e);
    return 0;
}

extern "C" int fdtd_add_monitor_op(fdtd_engine* e, const fdtd_monitor_op* op, int32_t* id)
{
    if (!e || !op) return fail(FDTD_EINVAL, "fdtd_add_monitor_op: null argument");
    MonOp m{};
    if (int rc = check_box(e, op->component, op->lo, op->hi, m.n)) return rc;
    if (op->n_freq < 0 || (op->n_freq > 0 && op->phasor_col < 0)) return fail(FDTD_EINVAL, "bad n_freq/phasor_col");
    m.comp = op->component;
    for (int a = 0; a < 3; ++a) m.lo[a] = op->lo[a];
    m.record = op->record; m.n_freq = op->n_freq; m.phasor_col = op->phasor_col;
    m.cells = (long long)m.n[0] * m.n[1] * m.n[2];
    if (id) *id = (int32_t)e->mon.size();
    e->mon.push_back(m);
    e->ops_dirty = true;
    drop_graph(e);
    return 0;
}

// region-correct flux (extension): power through [lo,hi) (a box valid for all six components), normal = direction
extern "C" int fdtd_add_flux_op(fdtd_engine* e, int32_t direction, const int32_t* lo, const int32_t* hi, int32_t* id)
{
    if (!e || !lo || !hi || direction < 0 || direction > 2) return fail(FDTD_EINVAL, "fdtd_add_flux_op: bad argument");
    FluxOp f{};
    for (int c = 0; c < 6; ++c)
        if (int rc = check_box(e, c, lo, hi, f.n)) return rc;
    f.dir = direction;
    for (int a = 0; a < 3; ++a) f.lo[a] = lo[a];
    f.cells = (long long)f.n[0] * f.n[1] * f.n[2];
    f.out_off = (long long)e->flux.size();
    if (id) *id = (int32_t)e->flux.size();
    e->flux.push_back(f);
    e->ops_dirty = true;
    drop_graph(e);
    return 0;
}

// instantaneous power samples of flux op id: host fp64 [steps_run] (sum of (E x H)_n over the box; multiply by dA)
extern "C" int fdtd_download_flux(fdtd_engine* e, int32_t id, double* host, int32_t max_steps)
{
    if (!e || !host || id < 0 || id >= (int)e->flux.size()) return fail(FDTD_EINVAL, "fdtd_download_flux: bad argument");
    CU(cudaSetDevice(e->cfg.device));
    const int steps = std::min<int>(max_steps, e->cursor);
    if (steps <= 0 || !e->d_flux_out) return 0;
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemcpy(host, e->d_flux_out + (size_t)id * std::max(e->n_steps_tab, 1), steps * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int fdtd_add_ade_op(fdtd_engine* e, const fdtd_ade_op* op, int32_t* id)
{
    if (!e || !op) return fail(FDTD_EINVAL, "fdtd_add_ade_op: null argument");
    if (op->component < 0 || op->component > 2) return fail(FDTD_EINVAL, "ADE ops are driven by an E component (0..2)");
    if (op->kind < 0 || op->kind > 2) return fail(FDTD_EINVAL, "ADE kind must be 0 (Lorentz), 1 (Drude) or 2 (Debye)");
    AdeOp a{};
    if (int rc = check_box(e, op->component, op->lo, op->hi, a.n)) return rc;
    a.comp = op->component; a.kind = op->kind;
    for (int k = 0; k < 3; ++k) a.lo[k] = op->lo[k];
    a.c0 = op->c0; a.c1 = op->c1; a.c2 = op->c2; a.c3 = op->c3;
    a.cells = (long long)a.n[0] * a.n[1] * a.n[2];
    a.mask_off = -1;
    if (op->mask && a.cells > 0) {
        a.mask_off = (long long)e->ade_mask_host.size();
        e->ade_mask_host.insert(e->ade_mask_host.end(), op->mask, op->mask + a.cells);
    }
    if (id) *id = (int32_t)e->ade.size();
    e->ade.push_back(a);
    e->ade_epoch++;
    e->ops_dirty = true;
    drop_graph(e);
    return 0;
}

// upload op tables, (re)allocate the dft pool; keeps existing DFT sums when the layout is unchanged
static int finalize_ops(fdtd_engine* e)
{
    if (!e->ops_dirty) return 0;
    CU(cudaStreamSynchronize(e->stream));
    std::stable_sort(e->src.begin(), e->src.end(), [](const HostSrc& a, const HostSrc& b) { return a.group < b.group; });
    e->grp_first.clear(); e->grp_count.clear(); e->grp_threads.clear();
    std::vector<SrcOp> flat;
    for (size_t i = 0; i < e->src.size();) {
        size_t j = i; long long t = 0;
        while (j < e->src.size() && e->src[j].group == e->src[i].group) {
            e->src[j].op.first_thread = t;
            t += (long long)e->src[j].op.n[0] * e->src[j].op.n[1] * e->src[j].op.n[2];
            flat.push_back(e->src[j].op);
            ++j;
        }
        e->grp_first.push_back((int)i); e->grp_count.push_back((int)(j - i)); e->grp_threads.push_back(t);
        i = j;
    }
    cudaFree(e->d_src); e->d_src = nullptr;
    if (!flat.empty()) {
        CU(cudaMalloc(&e->d_src, flat.size() * sizeof(SrcOp)));
        CU(cudaMemcpy(e->d_src, flat.data(), flat.size() * sizeof(SrcOp), cudaMemcpyHostToDevice));
    }
    cudaFree(e->d_prof); e->d_prof = nullptr;
    if (!e->prof_host.empty()) {
        CU(cudaMalloc(&e->d_prof, e->prof_host.size() * sizeof(double)));
        CU(cudaMemcpy(e->d_prof, e->prof_host.data(), e->prof_host.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    long long t = 0, rec = 0, dft = 0;
    for (auto& m : e->mon) {
        m.first_thread = t; t += m.cells;
        m.rec_off = rec;      // per-step offset; the kernel adds step * cells
        if (m.record) rec += m.cells;
        m.dft_off = dft; dft += (long long)m.n_freq * m.cells;
    }
    e->mon_threads = t;
    e->rec_elems_per_step = rec;
    cudaFree(e->d_mon); e->d_mon = nullptr;
    // The running sums survive a re-registration of the SAME ops (the Session re-lowers on every advance); any other
    // op list starts from zero, even when the pool happens to have the same size.
    std::vector<long long> dsig;
    for (auto& m : e->mon)
        if (m.n_freq > 0)
            dsig.insert(dsig.end(), {(long long)m.comp, m.lo[0], m.lo[1], m.lo[2], m.n[0], m.n[1], m.n[2], m.n_freq, m.dft_off});
    if (dft != e->dft_elems || !e->d_dft) {
        cudaFree(e->d_dft); e->d_dft = nullptr;
        if (dft > 0) {
            CU(cudaMalloc(&e->d_dft, dft * sizeof(double2)));
            CU(cudaMemset(e->d_dft, 0, dft * sizeof(double2)));
        }
        e->dft_elems = dft;
    } else if (dft > 0 && dsig != e->dft_sig) {
        CU(cudaMemset(e->d_dft, 0, dft * sizeof(double2)));
    }
    e->dft_sig = dsig;
    cudaFree(e->d_flux); e->d_flux = nullptr;
    cudaFree(e->d_flux_partial); e->d_flux_partial = nullptr;
    if (!e->flux.empty()) {
        CU(cudaMalloc(&e->d_flux, e->flux.size() * sizeof(FluxOp)));
        CU(cudaMemcpy(e->d_flux, e->flux.data(), e->flux.size() * sizeof(FluxOp), cudaMemcpyHostToDevice));
        CU(cudaMalloc(&e->d_flux_partial, e->flux.size() * FLUX_BLOCKS * sizeof(double)));
    }
    // ADE ops: aux pool (cur [+ prev] per op), zero-initialised when the layout changes
    {
        long long th = 0, aux = 0;
        for (auto& a : e->ade) {
            a.first_thread = th; th += a.cells;
            a.cur_off = aux; aux += a.cells;
            a.prev_off = -1;
            if (a.kind == 0) { a.prev_off = aux; aux += a.cells; }
        }
        e->ade_threads = th;
        cudaFree(e->d_ade); e->d_ade = nullptr;
        cudaFree(e->d_ade_mask); e->d_ade_mask = nullptr;
        if (!e->ade.empty()) {
            CU(cudaMalloc(&e->d_ade, e->ade.size() * sizeof(AdeOp)));
            CU(cudaMemcpy(e->d_ade, e->ade.data(), e->ade.size() * sizeof(AdeOp), cudaMemcpyHostToDevice));
            if (!e->ade_mask_host.empty()) {
                CU(cudaMalloc(&e->d_ade_mask, e->ade_mask_host.size()));
                CU(cudaMemcpy(e->d_ade_mask, e->ade_mask_host.data(), e->ade_mask_host.size(), cudaMemcpyHostToDevice));
            }
        }
        std::vector<long long> asig;
        for (auto& a : e->ade)
            asig.insert(asig.end(), {(long long)a.comp, (long long)a.kind, a.lo[0], a.lo[1], a.lo[2], a.n[0], a.n[1], a.n[2], a.cur_off});
        if (aux != e->aux_elems || (!e->d_aux && aux > 0)) {
            cudaFree(e->d_aux); e->d_aux = nullptr;
            if (aux > 0) {
                CU(cudaMalloc(&e->d_aux, aux * e->esz));
                CU(cudaMemset(e->d_aux, 0, aux * e->esz));
            }
            e->aux_elems = aux;
        } else if (aux > 0 && asig != e->aux_sig) {
            CU(cudaMemset(e->d_aux, 0, aux * e->esz));
        }
        e->aux_sig = asig;
    }
    // per-plane op flags for the temporally blocked sweep (bit0: a source op covers the plane, bit1: a monitor op)
    {
        const int npl = e->g.nx + 4;
        std::vector<unsigned char> fl(npl, 0);
        for (auto& h : e->src)
            for (int p = h.op.lo[0]; p < h.op.lo[0] + h.op.n[0] && p < npl; ++p) fl[p] |= 1;
        for (auto& g : e->src_ghost)
            for (int p = g.lo[0]; p < g.lo[0] + g.n[0] && p < npl; ++p) fl[p] |= 1;
        for (auto& m : e->mon)
            for (int p = m.lo[0]; p < m.lo[0] + m.n[0] && p < npl; ++p) fl[p] |= 2;
        if (!e->d_plane_flags) CU(cudaMalloc(&e->d_plane_flags, npl));
        CU(cudaMemcpy(e->d_plane_flags, fl.data(), npl, cudaMemcpyHostToDevice));
        e->plane_flags_host = fl;
        cudaFree(e->d_src_ghost); e->d_src_ghost = nullptr;
        if (!e->src_ghost.empty()) {
            CU(cudaMalloc(&e->d_src_ghost, e->src_ghost.size() * sizeof(SrcOp)));
            CU(cudaMemcpy(e->d_src_ghost, e->src_ghost.data(), e->src_ghost.size() * sizeof(SrcOp), cudaMemcpyHostToDevice));
        }
    }
    e->ops_dirty = false;
    return 0;
}

// state of an ADE op: which = 0 current (P or J), 1 previous (Lorentz only); host fp64 [cells]
static int ade_state_copy(fdtd_engine* e, int32_t id, int32_t which, double* host, bool to_device)
{
    if (!e || !host || id < 0 || id >= (int)e->ade.size() || which < 0 || which > 1)
        return fail(FDTD_EINVAL, "ADE state: bad argument");
    CU(cudaSetDevice(e->cfg.device));
    if (int rc = finalize_ops(e)) return rc;
    const AdeOp& a = e->ade[id];
    const long long off = which == 0 ? a.cur_off : a.prev_off;
    if (off < 0) return fail(FDTD_EINVAL, "ADE op %d has no previous-step state", id);
    if (a.cells == 0) return 0;
    CU(cudaStreamSynchronize(e->stream));
    if (e->cfg.dtype == FDTD_F64) {
        if (to_device) CU(cudaMemcpy((double*)e->d_aux + off, host, a.cells * sizeof(double), cudaMemcpyHostToDevice));
        else CU(cudaMemcpy(host, (double*)e->d_aux + off, a.cells * sizeof(double), cudaMemcpyDeviceToHost));
        return 0;
    }
    std::vector<float> tmp(a.cells);
    if (to_device) {
        for (long long i = 0; i < a.cells; ++i) tmp[i] = (float)host[i];
        CU(cudaMemcpy((float*)e->d_aux + off, tmp.data(), a.cells * sizeof(float), cudaMemcpyHostToDevice));
    } else {
        CU(cudaMemcpy(tmp.data(), (float*)e->d_aux + off, a.cells * sizeof(float), cudaMemcpyDeviceToHost));
        for (long long i = 0; i < a.cells; ++i) host[i] = tmp[i];
    }
    return 0;
}
extern "C" int fdtd_download_ade(fdtd_engine* e, int32_t id, int32_t which, double* host) { return ade_state_copy(e, id, which, host, false); }
extern "C" int fdtd_upload_ade(fdtd_engine* e, int32_t id, int32_t which, const double* host) { return ade_state_copy(e, id, which, const_cast<double*>(host), true); }

// record offsets depend on the number of tabled steps: op.rec_off = base(op) * n_steps
static int upload_mon_ops(fdtd_engine* e)
{
    cudaFree(e->d_mon); e->d_mon = nullptr;
    if (e->mon.empty()) return 0;
    std::vector<MonOp> ops = e->mon;
    for (auto& m : ops) m.rec_off = m.rec_off * (long long)std::max(e->n_steps_tab, 1);
    CU(cudaMalloc(&e->d_mon, ops.size() * sizeof(MonOp)));
    CU(cudaMemcpy(e->d_mon, ops.data(), ops.size() * sizeof(MonOp), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int fdtd_set_tables(fdtd_engine* e, int32_t n_steps, int32_t n_amp, const double* amp,
                               int32_t n_phasor, const double* phasors)
{
    if (!e || n_steps < 0 || n_amp < 0 || n_phasor < 0) return fail(FDTD_EINVAL, "fdtd_set_tables: bad argument");
    if ((n_amp > 0 && n_steps > 0 && !amp) || (n_phasor > 0 && n_steps > 0 && !phasors))
        return fail(FDTD_EINVAL, "fdtd_set_tables: null table");
    CU(cudaSetDevice(e->cfg.device));
    if (int rc = finalize_ops(e)) return rc;
    for (auto& h : e->src)
        if (h.op.table >= n_amp) return fail(FDTD_EINVAL, "source op uses table %d but n_amp = %d", h.op.table, n_amp);
    for (auto& m : e->mon)
        if (m.n_freq > 0 && m.phasor_col + m.n_freq > n_phasor)
            return fail(FDTD_EINVAL, "monitor op uses phasors [%d,%d) but n_phasor = %d", m.phasor_col, m.phasor_col + m.n_freq, n_phasor);
    CU(cudaStreamSynchronize(e->stream));
    cudaFree(e->d_amp); cudaFree(e->d_phasor); cudaFree(e->d_rec);
    e->d_amp = e->d_phasor = nullptr; e->d_rec = nullptr;
    e->n_steps_tab = n_steps; e->n_amp = n_amp; e->n_phasor = n_phasor;
    if (n_steps > 0 && n_amp > 0) {
        CU(cudaMalloc(&e->d_amp, (size_t)n_steps * n_amp * sizeof(double)));
        CU(cudaMemcpy(e->d_amp, amp, (size_t)n_steps * n_amp * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (n_steps > 0 && n_phasor > 0) {
        CU(cudaMalloc(&e->d_phasor, (size_t)n_steps * n_phasor * 2 * sizeof(double)));
        CU(cudaMemcpy(e->d_phasor, phasors, (size_t)n_steps * n_phasor * 2 * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (n_steps > 0 && e->rec_elems_per_step > 0)
        CU(cudaMalloc(&e->d_rec, (size_t)n_steps * e->rec_elems_per_step * e->esz));
    cudaFree(e->d_flux_out); e->d_flux_out = nullptr;
    if (n_steps > 0 && !e->flux.empty()) {
        CU(cudaMalloc(&e->d_flux_out, (size_t)n_steps * e->flux.size() * sizeof(double)));
        CU(cudaMemset(e->d_flux_out, 0, (size_t)n_steps * e->flux.size() * sizeof(double)));
    }
    if (int rc = upload_mon_ops(e)) return rc;
    e->cursor = 0;
    CU(cudaMemset(e->d_step, 0, sizeof(int)));
    drop_graph(e);
    return 0;
}
