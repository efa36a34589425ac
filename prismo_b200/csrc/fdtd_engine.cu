// fdtd_engine.cu — host side of libfdtd_b200.so: the C ABI declared in include/fdtd_b200.h.
//
// Owns the device state of one FDTD engine (one GPU / one x-slab): six padded SoA field arrays,
// optional cell-centred coefficient arrays, source / monitor op tables, per-step amplitude and
// phasor tables, record and DFT pools, and the stream / CUDA graph that replays the step loop.
// No torch, no Python: plain CUDA runtime.
#include "../../include/fdtd_b200.h"
#include "fdtd_kernels.cuh"
#include "fdtd_fused.cuh"
#include "fdtd_yee.cuh"
#include "fdtd_tb2.cuh"
#include "fdtd_het.cuh"
#include "fdtd_tensor.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace fdtd;

static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(e_ == cudaErrorMemoryAllocation ? FDTD_ENOMEM : FDTD_ECUDA, "%s: %s",     \
                        #call, cudaGetErrorString(e_));                                           \
    } while (0)

static inline long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

struct HostSrc { SrcOp op; int group; };

struct fdtd_engine {
    fdtd_config cfg{};
    Geom g{};
    Strides3 st{};
    size_t esz = 4;                 // element size of T
    long long plane_elems = 0;      // sx
    long long planes_alloc = 0;     // nx + 2
    long long array_elems = 0;
    void* fld[6] = {};              // set A
    void* fldB[6] = {};             // set B (fused ping-pong), allocated lazily
    int cur = 0;                    // which set holds the current fields (fused path)
    void* coef[4] = {};             // Ca Cb Da Db arrays (T) or null
    bool het = false;
    double uni[4] = {1, 0, 1, 0};
    cudaStream_t stream = nullptr;
    // ops
    std::vector<HostSrc> src;
    std::vector<SrcOp> src_ghost; SrcOp* d_src_ghost = nullptr;   // slabs: neighbour's sources on our ghost planes
    std::vector<MonOp> mon;
    std::vector<double> prof_host;
    std::vector<FluxOp> flux; FluxOp* d_flux = nullptr; double* d_flux_partial = nullptr; double* d_flux_out = nullptr;
    std::vector<AdeOp> ade; std::vector<unsigned char> ade_mask_host;
    AdeOp* d_ade = nullptr; void* d_aux = nullptr; unsigned char* d_ade_mask = nullptr;
    long long aux_elems = 0, ade_threads = 0;
    bool ops_dirty = true;
    Cpml cpml{}; SlabGeom slabg{}; double* d_cpml_coef = nullptr; size_t psi_bytes[12] = {};
    SrcOp* d_src = nullptr;         // all source ops, ordered by group
    std::vector<int> grp_first, grp_count; std::vector<long long> grp_threads;
    MonOp* d_mon = nullptr; long long mon_threads = 0;
    double* d_prof = nullptr;
    void** d_comp_ptr[2] = {nullptr, nullptr};   // device arrays of 6 component pointers (set A / B)
    // tables
    int n_steps_tab = 0, n_amp = 0, n_phasor = 0;
    double *d_amp = nullptr, *d_phasor = nullptr;
    void* d_rec = nullptr; long long rec_elems_per_step = 0;
    double2* d_dft = nullptr; long long dft_elems = 0;
    int* d_step = nullptr;          // table cursor (device)
    int cursor = 0;                 // host mirror of the cursor
    int* d_cnt = nullptr;           // 2 x 6 gate counters (2-D)
    long long steps_done = 0, launches = 0;
    // graph
    cudaGraphExec_t gexec[2] = {nullptr, nullptr}; int graph_steps = 0; int graph_kernels[2] = {0, 0};
    int fused_lx = 0;               // planes per fused segment (0 = auto)
    int het_fused = 1;              // heterogeneous media: fused one-step sweep (0: two-pass kernels)
    int tb2_zones = -1;             // two-step sweep: narrow x-segments around op planes (-1 auto, 0 never, 1 always)
    int tb2 = 1;                    // 1: temporally blocked sweep (two steps per pass) where applicable
    unsigned char* d_plane_flags = nullptr; std::vector<unsigned char> plane_flags_host;
    int fused_tj = 15;              // owner rows per CTA (15: one 16-warp CTA/SM; 7: two 8-warp CTAs/SM)
    int fused_pol = 0;              // bit0: streaming (evict-first) stores (measured 1.4% slower: off)
    // staging
    void* d_stage = nullptr; size_t stage_bytes = 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    // x-slab peer-to-peer halo (one process per GPU, CUDA IPC over NVLink)
    struct Slab {
        bool connected = false, has_left = false, has_right = false;
        int* flags = nullptr;            // own device words: [0] halo_ready (written by the right neighbour),
                                         // [1] ghost_consumed (written by us, read by the right neighbour), [2] error
        void* left_fld[2][6] = {};       // left neighbour's arrays (IPC-mapped): we push into its ghost planes
        int* left_flags = nullptr;       // left neighbour's flags (we write [0], read [1])
        int* seq = nullptr;              // seq[i] = i: source of the 4-byte DMA that publishes halo_ready = i
        void* left_base[13] = {};        // mapped bases to close
        cudaStream_t comm = nullptr;
        cudaEvent_t post_done = nullptr, push_done = nullptr;
        long long step = 0;              // steps run through fdtd_slab_run (same on every rank)
        int left_nx = 0;
        unsigned long long timeout_ns = 20000000000ull;
    } slab;
    FusedPlan fused{};
};

// ---------------------------------------------------------------------------------------------------
static void comp_shape(const fdtd_engine* e, int comp, int shp[3])
{
    // staggered shapes, core/grid.py:157-168 (LOCAL nx; the global trim of the last plane is applied
    // by the caller through x_offset/nx_global)
    const int nx = e->g.nx, ny = e->g.ny, nz = e->g.nz;
    const bool last = (e->g.x0 + nx == e->g.nxg);
    const int nxm = last ? nx - 1 : nx;
    static const int shortx[6] = {0, 1, 1, 1, 0, 0}, shorty[6] = {1, 0, 1, 0, 1, 0}, shortz[6] = {1, 1, 0, 0, 0, 1};
    shp[0] = shortx[comp] ? nxm : nx;
    shp[1] = shorty[comp] ? ny - 1 : ny;
    if (e->cfg.ndim == 3) shp[2] = shortz[comp] ? nz - 1 : nz;
    else shp[2] = 1;
}

template <typename T> static Fields<T> fields_of(void* const* p)
{
    Fields<T> f;
    f.ex = (T*)p[0]; f.ey = (T*)p[1]; f.ez = (T*)p[2]; f.hx = (T*)p[3]; f.hy = (T*)p[4]; f.hz = (T*)p[5];
    return f;
}
template <typename T> static Coefs<T> coefs_of(const fdtd_engine* e)
{
    Coefs<T> c;
    c.ca = (const T*)e->coef[0]; c.cb = (const T*)e->coef[1];
    c.da = (const T*)e->coef[2]; c.db = (const T*)e->coef[3];
    c.uca = (T)e->uni[0]; c.ucb = (T)e->uni[1]; c.uda = (T)e->uni[2]; c.udb = (T)e->uni[3];
    return c;
}
// fp32 fused kernels: db/d and cb/d folded once (both the one-step and the two-step sweep use the SAME folded
// arithmetic, so fp32 results do not depend on how steps are paired)
static Fold fold_of(const fdtd_engine* e)
{
    Fold fo;
    const Geom& g = e->g;
    const double d[3] = {g.dx, g.dy, g.dz};
    for (int a = 0; a < 3; ++a) {
        fo.d[a] = e->uni[3] / d[a];          // db / d
        fo.d[3 + a] = e->uni[1] / d[a];      // cb / d
    }
    for (int a = 0; a < 6; ++a) fo.f[a] = (float)fo.d[a];
    fo.fast64 = (e->cfg.flags & FDTD_FLAG_FAST_F64) ? 1 : 0;
    return fo;
}
static void** cur_fields(fdtd_engine* e) { return e->cur ? e->fldB : e->fld; }

static int ensure_stage(fdtd_engine* e, size_t bytes)
{
    if (e->stage_bytes >= bytes) return 0;
    if (e->d_stage) cudaFree(e->d_stage);
    e->d_stage = nullptr; e->stage_bytes = 0;
    CU(cudaMalloc(&e->d_stage, bytes));
    e->stage_bytes = bytes;
    return 0;
}

static void drop_graph(fdtd_engine* e)
{
    for (int q = 0; q < 2; ++q)
        if (e->gexec[q]) { cudaGraphExecDestroy(e->gexec[q]); e->gexec[q] = nullptr; }
    e->graph_steps = 0;
}

// ---------------------------------------------------------------------------------------------------
extern "C" int fdtd_abi_version(void) { return FDTD_B200_ABI_VERSION; }
extern "C" const char* fdtd_last_error(void) { return g_err.c_str(); }
extern "C" int fdtd_struct_size(int32_t which)
{
    switch (which) {
    case 0: return (int)sizeof(fdtd_config);
    case 1: return (int)sizeof(fdtd_source_op);
    case 2: return (int)sizeof(fdtd_monitor_op);
    case 3: return (int)sizeof(fdtd_ade_op);
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
    g.rdx = (float)(1.0 / g.dx); g.rdy = (float)(1.0 / g.dy); g.rdz = cfg->ndim == 3 ? (float)(1.0 / g.dz) : 0.f;
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
    if (const char* pol = getenv("FDTD_B200_FUSED_POL")) e->fused_pol = atoi(pol) & 3;
    if (const char* tj = getenv("FDTD_B200_FUSED_TJ")) e->fused_tj = atoi(tj);
    if (const char* tb = getenv("FDTD_B200_TB2")) e->tb2 = atoi(tb);
    if (const char* z = getenv("FDTD_B200_TB2_ZONES")) e->tb2_zones = atoi(z);
    if (const char* hf = getenv("FDTD_B200_HET_FUSED")) e->het_fused = atoi(hf);
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
    for (int c = 0; c < 4; ++c) cudaFree(e->coef[c]);
    cudaFree(e->d_src); cudaFree(e->d_mon); cudaFree(e->d_prof); cudaFree(e->d_src_ghost);
    cudaFree(e->d_ade); cudaFree(e->d_aux); cudaFree(e->d_ade_mask);
    cudaFree(e->d_flux); cudaFree(e->d_flux_partial); cudaFree(e->d_flux_out);
    cudaFree(e->d_cpml_coef); cudaFree(e->d_plane_flags);
    for (int q = 0; q < 12; ++q) cudaFree(e->cpml.psi[q]);
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
    for (int c = 0; c < 4; ++c) { cudaFree(e->coef[c]); e->coef[c] = nullptr; }
    e->het = false;
    e->uni[0] = ca; e->uni[1] = cb; e->uni[2] = da; e->uni[3] = db;
    drop_graph(e);
    return 0;
}

// same dtype: one strided DMA between the caller's (compact) buffer and the padded device array
static int copy_strided(fdtd_engine* e, void* dev, void* host, long long c0, int c1, int c2, bool to_device)
{
    if (c0 * c1 * c2 == 0) return 0;
    const size_t esz = e->esz;
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

extern "C" int fdtd_set_coeffs(fdtd_engine* e, const double* ca, const double* cb, const double* da,
                               const double* db, int32_t planes)
{
    if (!e || !ca || !cb || !da || !db) return fail(FDTD_EINVAL, "fdtd_set_coeffs: null argument");
    if (planes != e->g.nx && planes != e->g.nx + 1)
        return fail(FDTD_EINVAL, "coefficient arrays must have nx=%d (or nx+1) planes, got %d", e->g.nx, planes);
    CU(cudaSetDevice(e->cfg.device));
    const double* src[4] = {ca, cb, da, db};
    for (int c = 0; c < 4; ++c) {
        if (!e->coef[c]) CU(cudaMalloc(&e->coef[c], e->array_elems * e->esz));
        CU(cudaMemsetAsync(e->coef[c], 0, e->array_elems * e->esz, e->stream));
        int rc;
        const int c1 = e->g.ny, c2 = e->cfg.ndim == 3 ? e->g.nz : 1;
        if (e->cfg.dtype == FDTD_F64) rc = scatter_host<double, double>(e, (double*)e->coef[c], src[c], planes, c1, c2);
        else rc = scatter_host<float, double>(e, (float*)e->coef[c], src[c], planes, c1, c2);
        if (rc) return rc;
    }
    e->het = true;
    drop_graph(e);
    return 0;
}

// ---- fields ---------------------------------------------------------------------------------------------
extern "C" int fdtd_upload_field(fdtd_engine* e, int32_t comp, const void* host, int32_t host_dtype)
{
    if (!e || !host || comp < 0 || comp > 5) return fail(FDTD_EINVAL, "fdtd_upload_field: bad argument");
    CU(cudaSetDevice(e->cfg.device));
    int s[3]; comp_shape(e, comp, s);
    void* dst = cur_fields(e)[comp];
    CU(cudaMemsetAsync(dst, 0, e->array_elems * e->esz, e->stream));
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
    for (int q = 0; q < 12; ++q)
        if (e->cpml.psi[q]) CU(cudaMemsetAsync(e->cpml.psi[q], 0, e->psi_bytes[q], e->stream));
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
    if (dft != e->dft_elems || !e->d_dft) {
        cudaFree(e->d_dft); e->d_dft = nullptr;
        if (dft > 0) {
            CU(cudaMalloc(&e->d_dft, dft * sizeof(double2)));
            CU(cudaMemset(e->d_dft, 0, dft * sizeof(double2)));
        }
        e->dft_elems = dft;
    }
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
        if (aux != e->aux_elems || (!e->d_aux && aux > 0)) {
            cudaFree(e->d_aux); e->d_aux = nullptr;
            if (aux > 0) {
                CU(cudaMalloc(&e->d_aux, aux * e->esz));
                CU(cudaMemset(e->d_aux, 0, aux * e->esz));
            }
            e->aux_elems = aux;
        }
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
template <typename T> static int launch_post(fdtd_engine* e, int step_off, int parity, cudaStream_t s)
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
    if (e->ade_threads > 0) {
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
    for (int q = 0; q < 12; ++q) { cudaFree(e->cpml.psi[q]); e->cpml.psi[q] = nullptr; }
    e->cpml = Cpml{};
    drop_graph(e);
    if (thickness == 0) return 0;
    if (!coef) return fail(FDTD_EINVAL, "fdtd_set_cpml: null coefficients");
    const int N[3] = {g.nx, g.ny, g.nz};
    const size_t total = 6 * ((size_t)g.nx + g.ny + g.nz);
    CU(cudaMalloc(&e->d_cpml_coef, total * sizeof(double)));
    CU(cudaMemcpy(e->d_cpml_coef, coef, total * sizeof(double), cudaMemcpyHostToDevice));
    size_t off = 0;
    for (int a = 0; a < 3; ++a) {
        for (int v = 0; v < 6; ++v) e->cpml.ax[a].c[v] = e->d_cpml_coef + off + (size_t)v * N[a];
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

// one fused sweep over planes [i_begin, i_end): reads the current set, writes the other one
template <typename T, int TJ> static int launch_fused_tj(fdtd_engine* e, int i_begin, int i_end, cudaStream_t s)
{
    constexpr int V = VecOf<T>::V;
    const Geom& g = e->g;
    void** src = cur_fields(e);
    void** dst = e->cur ? e->fld : e->fldB;
    CFields<T> in;
    in.ex = (const T*)src[0]; in.ey = (const T*)src[1]; in.ez = (const T*)src[2];
    in.hx = (const T*)src[3]; in.hy = (const T*)src[4]; in.hz = (const T*)src[5];
    Fields<T> out = fields_of<T>(dst);
    FusedTiling t;
    t.i_begin = i_begin; t.i_end = i_end;
    t.halo_flag = nullptr; t.halo_need = 0; t.error_word = nullptr; t.timeout_ns = e->slab.timeout_ns;
    if (e->slab.connected && e->slab.has_right && i_end == g.nx) {
        t.halo_flag = e->slab.flags; t.halo_need = (int)e->slab.step + 1; t.error_word = e->slab.flags + 2;
    }
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
    const size_t smem = fused_smem_bytes<T, TJ>();
    auto kern = k_fused3d<T, TJ, 0>;
    switch (e->fused_pol & 3) {
    case 1: kern = k_fused3d<T, TJ, 1>; break;
    case 2: kern = k_fused3d<T, TJ, 2>; break;
    case 3: kern = k_fused3d<T, TJ, 3>; break;
    default: break;
    }
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 block(32, TJ + 1, 1);
    const unsigned items = (unsigned)t.nseg * t.ntj * t.ntk;
    kern<<<items, block, smem, s>>>(in, out, coefs_of<T>(e), g, t, fold_of(e));
    e->launches++;
    CU(cudaGetLastError());
    return 0;
}

// one fused sweep over planes [i_begin, i_end): reads the current set, writes the other one
template <typename T> static int launch_fused(fdtd_engine* e, int i_begin, int i_end, cudaStream_t s)
{
    if (e->fused_tj == 7) return launch_fused_tj<T, 7>(e, i_begin, i_end, s);
    return launch_fused_tj<T, kFusedTJ>(e, i_begin, i_end, s);
}

// heterogeneous media, one GPU: fused one-step sweep that also streams the four coefficient arrays
static bool use_het_fused(const fdtd_engine* e)
{
    return e->cfg.ndim == 3 && e->het && !(e->cfg.flags & (FDTD_FLAG_TWO_PASS | FDTD_FLAG_YEE)) && e->g.nxg == e->g.nx &&
           e->array_elems < (1ll << 32) && e->het_fused;
}

template <typename T> static int launch_het(fdtd_engine* e, cudaStream_t s)
{
    constexpr int R = kHetRows, V = Vec8<T>::V;
    const Geom& g = e->g;
    void** src = cur_fields(e);
    void** dst = e->cur ? e->fld : e->fldB;
    CFields<T> in;
    in.ex = (const T*)src[0]; in.ey = (const T*)src[1]; in.ez = (const T*)src[2];
    in.hx = (const T*)src[3]; in.hy = (const T*)src[4]; in.hz = (const T*)src[5];
    Fields<T> out = fields_of<T>(dst);
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
    const size_t smem = het_smem_bytes<T, R>();
    auto kern = k_fused3d_het<T, R>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 block(32, R, 1);
    kern<<<(unsigned)t.nseg * t.ntj * t.ntk, block, smem, s>>>(in, out, coefs_of<T>(e), g, t, (int)e->planes_alloc);
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
    t.ntk = (vec_per_row + t.own_lanes - 1) / t.own_lanes;
    t.ntj = (g.ny + (R - 4) - 1) / (R - 4);
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
    auto kern = k_fused3d_tb2<T, R>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned items = (unsigned)t.nseg * t.ntj * t.ntk;
    kern<<<items, block, smem, s>>>(in, out, cf, g, t, m, (int)e->planes_alloc, fo);
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
    if (e->cfg.flags & FDTD_FLAG_YEE) return launch_yee<T>(e, half, s);
    if (use_het_fused(e)) {
        if (half == 1) return 0;
        if (int rc = ensure_set_b(e)) return rc;
        return launch_het<T>(e, s);
    }
    return launch_pass3d<T>(e, half, 0, e->g.nx, s);
}

template <typename T> static int one_step(fdtd_engine* e, int step_off, int parity, cudaStream_t s)
{
    if (e->cfg.ndim == 3) {
        if (int rc = step_fields3d<T>(e, 0, s)) return rc;
        if (int rc = step_fields3d<T>(e, 1, s)) return rc;
    } else {
        if (int rc = launch_pass2d<T>(e, 0, parity, s)) return rc;
        if (int rc = launch_pass2d<T>(e, 1, parity, s)) return rc;
    }
    return launch_post<T>(e, step_off, parity, s);
}

static bool has_tables(const fdtd_engine* e) { return !e->src.empty() || !e->mon.empty() || !e->src_ghost.empty() || !e->flux.empty(); }
static bool has_post(const fdtd_engine* e) { return has_tables(e) || !e->ade.empty(); }

template <typename T> static int run_steps(fdtd_engine* e, int n)
{
    cudaStream_t s = e->stream;
    if (use_fused(e) || use_het_fused(e)) if (int rc = ensure_set_b(e)) return rc;
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
            else for (int q = 0; q < G && !rc; ++q) rc = one_step<T>(e, q, q, s);
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
        if (int rc = one_step<T>(e, q, done + q, s)) return rc;
    if (rest > 0) { k_bump<<<1, 1, 0, s>>>(e->d_step, rest); e->launches++; CU(cudaGetLastError()); }
    return 0;
}

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
        if ((rc = launch_post<T>(e, q, q, s))) return rc;
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
    else if (!strcmp(key, "het_fused")) e->het_fused = value ? 1 : 0;
    else if (!strcmp(key, "fused_lx")) e->fused_lx = value;
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

// ---- anisotropic tensor update on caller-supplied arrays (row a23; materials/tensor.py:482-588) -----------------------
template <typename T>
static int tensor_update(int device, long long n, const void* const* f, const void* const* curl, void* const* out,
                         double scale, int negative, int mode, const double* coef, const void* const* coef_arrays)
{
    const int full = mode & 1;
    CU(cudaSetDevice(device));
    const int n_coef = full ? 9 : 3;
    // device staging: f, curl, out (3 each) + per-cell coefficient arrays, processed in chunks
    long long chunk_max = 1ll << 24;
    if (const char* ce = getenv("FDTD_B200_TENSOR_CHUNK")) chunk_max = std::max<long long>(1, atoll(ce));
    const long long chunk = std::min<long long>(n, chunk_max);
    int n_arr = 9;
    for (int q = 0; q < n_coef; ++q) if (coef_arrays && coef_arrays[q]) ++n_arr;
    T* pool = nullptr;
    CU(cudaMalloc(&pool, sizeof(T) * (size_t)chunk * n_arr));
    cudaStream_t st;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { cudaFree(pool); return fail(FDTD_ECUDA, "stream"); }
    int rc = 0;
    for (long long off = 0; off < n && rc == 0; off += chunk) {
        const long long m = std::min(chunk, n - off);
        TensorArgs<T> a;
        memset(&a, 0, sizeof a);
        T* next = pool;
        auto stage = [&](const void* host) -> T* {
            T* d = next; next += chunk;
            if (cudaMemcpyAsync(d, (const T*)host + off, sizeof(T) * (size_t)m, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = FDTD_ECUDA;
            return d;
        };
        for (int c = 0; c < 3; ++c) {
            const bool need_curl = curl[c] && (full || f[c]);
            if (f[c]) { a.f[c] = stage(f[c]); a.out[c] = next; next += chunk; }
            if (need_curl) a.curl[c] = stage(curl[c]);
        }
        for (int q = 0; q < n_coef; ++q) {
            a.coef[q] = (T)coef[q];
            if (coef_arrays && coef_arrays[q]) a.coef_arr[q] = stage(coef_arrays[q]);
        }
        a.s = (T)scale; a.negative = negative; a.full = full;
        a.mul_f32 = (mode >> 1) & 1; a.div_f32 = (mode >> 2) & 1;
        if (rc) break;
        const int block = 256;
        const int grid = (int)std::min<long long>((m + block - 1) / block, 148 * 16);
        k_tensor_update<T><<<grid, block, 0, st>>>(a, m);
        if (cudaGetLastError() != cudaSuccess) { rc = FDTD_ECUDA; break; }
        for (int c = 0; c < 3; ++c)
            if (f[c] && cudaMemcpyAsync((T*)out[c] + off, a.out[c], sizeof(T) * (size_t)m, cudaMemcpyDeviceToHost, st) != cudaSuccess)
                rc = FDTD_ECUDA;
        if (cudaStreamSynchronize(st) != cudaSuccess) rc = FDTD_ECUDA;
    }
    cudaStreamDestroy(st);
    cudaFree(pool);
    if (rc) return fail(rc, "fdtd_tensor_update: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

extern "C" int fdtd_tensor_update(int32_t device, int32_t dtype, int64_t n, const void* const* f, const void* const* curl,
                                  void* const* out, double scale, int32_t negative, int32_t mode, const double* coef,
                                  const void* const* coef_arrays)
{
    const int full = mode & 1;
    if (n < 0 || !f || !curl || !out || !coef) return fail(FDTD_EINVAL, "fdtd_tensor_update: bad argument");
    if (dtype != FDTD_F32 && dtype != FDTD_F64) return fail(FDTD_EINVAL, "fdtd_tensor_update: dtype %d", dtype);
    const bool any = f[0] || f[1] || f[2];
    for (int c = 0; c < 3; ++c) {
        if (f[c] && !out[c]) return fail(FDTD_EINVAL, "fdtd_tensor_update: component %d has no output array", c);
        const bool need_curl = full ? any : f[c] != nullptr;          // the full tensor mixes all three curls
        if (need_curl && !curl[c]) return fail(FDTD_EINVAL, "fdtd_tensor_update: curl component %d missing", c);
    }
    if (n == 0 || !(f[0] || f[1] || f[2])) return 0;
    return dtype == FDTD_F64 ? tensor_update<double>(device, n, f, curl, out, scale, negative, mode, coef, coef_arrays)
                             : tensor_update<float>(device, n, f, curl, out, scale, negative, mode, coef, coef_arrays);
}

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
    if (!use_fused(e)) return fail(FDTD_ESTATE, "peer-to-peer slabs need the fused path (3-D, uniform coefficients)");
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
    int q = 0;
    while (q < n) {
        const bool pair = tb2_ok(e) && q + 2 <= n;
        const int* planes = pair ? planes2 : planes1;
        const long long st = sl.step;                    // exchange counter, identical on every rank
        if (sl.has_left) {
            CU(cudaStreamWaitEvent(ms, sl.post_done, 0));        // our first planes of the current set are final
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
        }
        if (pair) {
            if (dbg) { dbg_ev.emplace_back(); cudaEventCreate(&dbg_ev.back()); cudaEventRecord(dbg_ev.back(), cs); }
            if (int rc = launch_tb2<T>(e, q, cs)) return rc;             // flips the sets itself
            if (dbg) { dbg_ev.emplace_back(); cudaEventCreate(&dbg_ev.back()); cudaEventRecord(dbg_ev.back(), cs); }
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
        if (has_post(e)) if (int rc = launch_post<T>(e, q - 1, 0, cs)) return rc;
        CU(cudaEventRecord(sl.post_done, cs));
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
        fprintf(stderr, "[fdtd dbg] dev %d pairs %zu sweep %.4f ms period %.4f ms\n", e->cfg.device, np, kern / np, period / (np - 1));
        for (auto ev : dbg_ev) cudaEventDestroy(ev);
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
