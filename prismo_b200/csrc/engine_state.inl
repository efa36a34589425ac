// engine state, error reporting, small helpers
// (textual part of fdtd_engine.cu — one translation unit; not compiled on its own)

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

// Reciprocals of a spacing.  y = RN(1/d) (IEEE division on the host) drives the exact FMA division sequence of
// Ar<double>::div_fast; fdtd_create accepts fp64 spacings in [2^-60, 2^60] only, so that a = q*d stays far from the
// exponent limits whenever the quotient passes the kernels' range test.
static Rcp make_rcp(double d)
{
    Rcp r;
    r.f = (float)(1.0 / d);
    r.y = 1.0 / d;
    return r;
}
static bool fold64(const fdtd_engine* e);

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
    void* coef[6] = {};             // Ca Cb Da Db arrays (T) or null; [4], [5] = Cb of Ey, Ez when aniso (then Cb is Ex's)
    bool het = false;
    unsigned char* mat = nullptr;   // material-index coding (fdtd_rasterize, lists of <= 64 entries): one byte per cell, field layout
    void* mat_tab = nullptr;        // [kMatTabRows][6] T
    int n_mat = 0;                  // > 0: mat / mat_tab describe the same medium as coef[]
    int het_indexed = 1;            // the fused heterogeneous sweep reads mat (1 B per cell) instead of the coefficient arrays where mat exists
    bool aniso = false;             // per-component Cb (diagonal permittivity tensor) in the E stage: OPT-IN extension, 3-D parity sweeps only
    int coef_planes = 0;            // planes of Ca..Db supplied by the caller (nx, or nx + 1 with the right neighbour's first)
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
    int ade_fused = 1;              // apply the recursions inside the next fused sweep where one follows (0: always k_ade)
    int ade_coupled = 0;            // OPT-IN, not the reference's behaviour: feed the polarisation current back into E
    bool ade_deferred = false;      // the recursion of the last enqueued step is still to be applied by the next sweep
    long long ade_epoch = 0;        // bumps whenever the recursion list changes
    int* d_ade_order = nullptr; long long ade_order_items = 0; std::vector<long long> ade_order_sig;   // box items first
    bool ops_dirty = true;
    Cpml cpml{}; SlabGeom slabg{}; double* d_cpml_coef = nullptr; float* d_cpml_coef_f = nullptr; size_t psi_bytes[12] = {};
    void* psiB[12] = {};            // second psi set: the fused physics sweep ping-pongs psi like the fields
    int yee_fused = 2;              // physics mode: 2 = TMA-fed fused one-sweep step (fdtd_yeex.cuh, default where it applies),
                                    // 1 = register-prefetch fused sweep (fdtd_yee_fused.cuh), 0 = two-pass kernels (fdtd_yee.cuh)
    int yeex_stages = 4, yeex_slots = 3;
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
    std::vector<long long> dft_sig, aux_sig;   // layout of the DFT / ADE pools: sums are kept only across identical op lists
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
    int tb2x = 1;                   // two-step sweep variant: 1 = TMA-fed, mbarrier-synchronised (fdtd_tb2x.cuh), 0 = fdtd_tb2.cuh
    int tb2x_stages = 4, tb2x_slots = 3;   // depth of its input ring (TMA stages) and of its row-exchange ring
    Tb2xMaps tmaps[2];              // TMA descriptors of the six arrays of set A / set B
    bool tmaps_ok = false;
    Tb2xMaps ymaps[2];              // the same arrays with the physics sweep's 272-byte boxes
    bool ymaps_ok = false;
    unsigned char* d_plane_flags = nullptr; std::vector<unsigned char> plane_flags_host;
    // staging
    void* d_stage = nullptr; size_t stage_bytes = 0;
    cudaStream_t xfer2 = nullptr; cudaEvent_t xfer_ev = nullptr;     // second lane of the staged host <-> device copies
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
    c.cby = e->aniso ? (const T*)e->coef[4] : nullptr; c.cbz = e->aniso ? (const T*)e->coef[5] : nullptr;
    const bool idx = e->het && e->het_indexed && e->n_mat > 0;
    c.mat = idx ? e->mat : nullptr; c.mat_tab = idx ? (const T*)e->mat_tab : nullptr; c.n_mat = idx ? e->n_mat : 0;
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

static bool fold64(const fdtd_engine* e) { return (e->cfg.flags & FDTD_FLAG_FAST_F64) != 0; }
