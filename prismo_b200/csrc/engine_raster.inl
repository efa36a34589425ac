// geometry rasterisation on the device: shape list -> coefficient arrays, coefficient read-back
// (textual part of fdtd_engine.cu — one translation unit; not compiled on its own)

extern "C" int fdtd_rasterize(fdtd_engine* e, const fdtd_shape* shapes, int32_t n_shapes, const double* verts_xy,
                              int32_t n_verts, const double* x, const double* y, const double* z, int32_t planes,
                              const double* background)
{
    if (!e || !x || !y || !background || n_shapes < 0 || (n_shapes && !shapes))
        return fail(FDTD_EINVAL, "fdtd_rasterize: null argument");
    const bool is3 = e->cfg.ndim == 3;
    if (is3 && !z) return fail(FDTD_EINVAL, "fdtd_rasterize: a 3-D grid needs z coordinates");
    if (planes != e->g.nx && planes != e->g.nx + 1)
        return fail(FDTD_EINVAL, "fdtd_rasterize: x must have nx=%d (or nx+1) entries, got %d", e->g.nx, planes);
    RasterBg bg;
    bg.eps[0] = background[0]; bg.eps[1] = background[1]; bg.eps[2] = background[2];
    bg.mu_r = background[3]; bg.sigma_e = background[4]; bg.sigma_m = background[5];
    bool aniso = !(bg.eps[0] == bg.eps[1] && bg.eps[1] == bg.eps[2]);
    if (aniso && bg.sigma_e != 0.0) return fail(FDTD_EINVAL, "fdtd_rasterize: an anisotropic background needs sigma_e == 0");
    std::vector<RasterShape> hs((size_t)n_shapes);
    for (int q = 0; q < n_shapes; ++q) {
        const fdtd_shape& s = shapes[q];
        RasterShape& r = hs[q];
        if (s.kind < 0 || s.kind > 3 || s.axis < 0 || s.axis > 2 || s.combine < 0 || s.combine > 3)
            return fail(FDTD_EINVAL, "fdtd_rasterize: shape %d: bad kind / axis / combine", q);
        if (s.kind == 3 && (s.vert_first < 0 || s.vert_count < 1 || s.vert_first + s.vert_count > n_verts || !verts_xy))
            return fail(FDTD_EINVAL, "fdtd_rasterize: shape %d: polygon vertices [%d, %d) outside the vertex list of %d",
                        q, s.vert_first, s.vert_first + s.vert_count, n_verts);
        r.kind = s.kind; r.axis = s.axis; r.combine = s.combine; r.paint = s.paint != 0;
        for (int d = 0; d < 3; ++d) { r.c[d] = s.center[d]; r.a[d] = s.a[d]; r.eps[d] = s.eps_r[d]; }
        r.v0 = s.vert_first; r.nv = s.vert_count;
        r.mu_r = s.mu_r; r.sigma_e = s.sigma_e; r.sigma_m = s.sigma_m;
        if (r.paint && !(r.eps[0] == r.eps[1] && r.eps[1] == r.eps[2])) {
            if (r.sigma_e != 0.0) return fail(FDTD_EINVAL, "fdtd_rasterize: shape %d: an anisotropic material needs sigma_e == 0", q);
            aniso = true;
        }
    }
    if (aniso) if (int rc = aniso_supported(e, "fdtd_rasterize")) return rc;
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    const int n_arrays = aniso ? 6 : 4;
    for (int c = 0; c < n_arrays; ++c) {
        if (!e->coef[c]) CU(cudaMalloc(&e->coef[c], e->array_elems * e->esz));
        CU(cudaMemsetAsync(e->coef[c], 0, e->array_elems * e->esz, e->stream));
    }
    for (int c = n_arrays; c < 6; ++c) { cudaFree(e->coef[c]); e->coef[c] = nullptr; }
    // material-index coding rides along when the list is short enough for the kernel's shared-memory table
    const bool indexed = n_shapes <= kRasterSmemShapes && n_shapes + 1 <= kMatTabRows && e->cfg.ndim == 3;
    e->n_mat = 0;
    if (indexed) {
        if (!e->mat) CU(cudaMalloc(&e->mat, e->array_elems));
        if (!e->mat_tab) CU(cudaMalloc(&e->mat_tab, (size_t)kMatTabRows * 6 * e->esz));
        CU(cudaMemsetAsync(e->mat, 0, e->array_elems, e->stream));
        CU(cudaMemsetAsync(e->mat_tab, 0, (size_t)kMatTabRows * 6 * e->esz, e->stream));
    }
    // one small device buffer: coordinates, vertices, shapes
    const int c1 = e->g.ny, c2 = is3 ? e->g.nz : 1;
    const size_t n_coord = (size_t)planes + c1 + (is3 ? c2 : 0);
    const size_t bytes_d = (n_coord + 2 * (size_t)std::max(n_verts, 0)) * sizeof(double);
    const size_t bytes_s = hs.size() * sizeof(RasterShape);
    std::vector<double> hd(n_coord + 2 * (size_t)std::max(n_verts, 0));
    std::copy(x, x + planes, hd.begin());
    std::copy(y, y + c1, hd.begin() + planes);
    if (is3) std::copy(z, z + c2, hd.begin() + planes + c1);
    if (n_verts > 0) std::copy(verts_xy, verts_xy + 2 * (size_t)n_verts, hd.begin() + n_coord);
    unsigned char* d_buf = nullptr;
    CU(cudaMalloc(&d_buf, bytes_d + bytes_s + 16));
    cudaError_t err = cudaMemcpyAsync(d_buf, hd.data(), bytes_d, cudaMemcpyHostToDevice, e->stream);
    if (err == cudaSuccess && bytes_s)
        err = cudaMemcpyAsync(d_buf + bytes_d, hs.data(), bytes_s, cudaMemcpyHostToDevice, e->stream);
    if (err == cudaSuccess) {
        const double* d_x = (const double*)d_buf;
        const double* d_y = d_x + planes;
        const double* d_z = is3 ? d_y + c1 : nullptr;
        const double* d_v = d_x + n_coord;
        const RasterShape* d_s = (const RasterShape*)(d_buf + bytes_d);
        const long long total = (long long)planes * c1 * c2;
        const long long rows = total / (is3 ? c2 : c1);                      // one CTA per row of the contiguous axis, grid-stride
        const int blocks = (int)std::max<long long>(1, std::min<long long>(rows, 148 * 8));
        const double eps0 = 8.854187817e-12, mu0 = 4 * 3.141592653589793 * 1e-7;        // core/solver.py:62-63
        if (total > 0) {
#define RASTER_LAUNCH(T, A)                                                                                                   \
    k_rasterize<T, A><<<blocks, 256, 0, e->stream>>>((T*)e->coef[0], (T*)e->coef[1], (T*)e->coef[2], (T*)e->coef[3],          \
                                                     (T*)e->coef[4], (T*)e->coef[5], d_x, d_y, d_z, d_s, n_shapes, d_v, bg,   \
                                                     e->cfg.dt, eps0, mu0, total, c1, c2, e->st,                             \
                                                     indexed ? e->mat : nullptr, indexed ? (T*)e->mat_tab : nullptr)
            if (e->cfg.dtype == FDTD_F64) { if (aniso) RASTER_LAUNCH(double, true); else RASTER_LAUNCH(double, false); }
            else { if (aniso) RASTER_LAUNCH(float, true); else RASTER_LAUNCH(float, false); }
#undef RASTER_LAUNCH
            e->launches++;
            err = cudaGetLastError();
        }
    }
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);     // hd / hs / d_buf go away
    cudaFree(d_buf);
    if (err != cudaSuccess) return fail(FDTD_ECUDA, "fdtd_rasterize: %s", cudaGetErrorString(err));
    e->het = true;
    e->aniso = aniso;
    e->n_mat = indexed ? n_shapes + 1 : 0;
    e->coef_planes = planes;
    drop_graph(e);
    return 0;
}

extern "C" int fdtd_download_coeffs(fdtd_engine* e, int32_t which, double* host, int32_t planes)
{
    if (!e || !host || which < 0 || which > 5) return fail(FDTD_EINVAL, "fdtd_download_coeffs: bad argument");
    if (!e->het || !e->coef[which]) return fail(FDTD_ESTATE, "fdtd_download_coeffs: coefficient array %d is not set", which);
    if (planes < 1 || planes > e->coef_planes)
        return fail(FDTD_EINVAL, "fdtd_download_coeffs: %d planes requested, %d held", planes, e->coef_planes);
    CU(cudaSetDevice(e->cfg.device));
    const int c1 = e->g.ny, c2 = e->cfg.ndim == 3 ? e->g.nz : 1;
    if (e->cfg.dtype == FDTD_F64) return gather_host<double, double>(e, host, (const double*)e->coef[which], planes, c1, c2);
    return gather_host<float, double>(e, host, (const float*)e->coef[which], planes, c1, c2);
}
