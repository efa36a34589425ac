/*
 * fdtd_b200.h — C ABI of the B200-native FDTD time-stepping engine (libfdtd_b200.so).
 *
 * This is the drop-in boundary for ONE path of rithulkamesh/prismo: the FDTD time step
 *   Simulation.step            /root/reference/src/prismo/core/simulation.py:147-164
 *   -> FDTDSolver.step         core/solver.py:664-678
 *   -> MaxwellUpdater.step     core/solver.py:535-552  (H pass :167-253/:311-397, E pass :255-309/:399-456)
 *   -> Source.update_fields    sources/{point,plane_wave,tfsf,gaussian,mode}.py
 *   -> Monitor.update          monitors/{field,dft,flux,mode_monitor}.py
 * The reference is pure Python, so its "FFI" for this path is ctypes: INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain C types only; all host pointers are caller-owned; no callbacks.
 *   - every entry point returns 0 on success, a negative FDTD_E* code on failure;
 *     fdtd_last_error() returns a message for the last failure on this thread.
 *   - component ids: 0 Ex, 1 Ey, 2 Ez, 3 Hx, 4 Hy, 5 Hz   (core/fields.py:61-70 order)
 *   - host field arrays are C-order with the reference's staggered shapes (core/grid.py:157-168),
 *     e.g. Ex is (nx, ny-1, nz-1); 2-D arrays drop the z axis.
 *   - index boxes are half-open [lo, hi) in the component's own array index space
 *     (what np.ix_ of core/grid.py:383-513 enumerates).
 *   - a handle is not thread-safe; one handle drives one GPU (one x-slab in multi-GPU runs).
 */
#ifndef FDTD_B200_H
#define FDTD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDTD_B200_ABI_VERSION 1

enum { FDTD_OK = 0, FDTD_EINVAL = -1, FDTD_ECUDA = -2, FDTD_ENOMEM = -3, FDTD_ESTATE = -4 };
enum { FDTD_F32 = 0, FDTD_F64 = 1 };
enum { FDTD_EX = 0, FDTD_EY = 1, FDTD_EZ = 2, FDTD_HX = 3, FDTD_HY = 4, FDTD_HZ = 5 };

/* flags for fdtd_config.flags */
enum {
    FDTD_FLAG_NO_GRAPH = 1,     /* never replay the step loop from a CUDA graph            */
    FDTD_FLAG_TWO_PASS = 2,     /* force the two-pass (H kernel, E kernel) 3-D step        */
    FDTD_FLAG_FAST_F64 = 8,     /* fp64 fused sweeps with folded FMA arithmetic (like fp32): ~1e-15 relative per
                                   step from the exact mode, which stays the default (bit-identical to NumPy)  */
    FDTD_FLAG_YEE = 4           /* OPT-IN physics mode, not parity: stable Yee leap-frog (backward
                                   differences in the H update) + CPML; 3-D, two-pass kernels      */
};

typedef struct fdtd_engine fdtd_engine; /* opaque */

/* Replaces YeeGrid + MaxwellUpdater construction (core/grid.py:78-120, core/solver.py:41-77). */
typedef struct fdtd_config {
    int32_t ndim;          /* 2 or 3 (2 <=> reference Lz == 0, core/grid.py:92)                   */
    int32_t nx, ny, nz;    /* LOCAL total grid incl. PML cells (YeeGrid.dimensions); nz = 1 in 2-D */
    double  dx, dy, dz;    /* spacing (dz ignored in 2-D)                                          */
    double  dt;            /* time step                                                            */
    int32_t dtype;         /* FDTD_F32 | FDTD_F64: storage + arithmetic type on the device         */
    int32_t device;        /* CUDA device ordinal                                                  */
    int32_t nx_global;     /* x-slab decomposition: global nx (== nx on one GPU)                   */
    int32_t x_offset;      /* global index of local plane 0                                        */
    int32_t flags;         /* FDTD_FLAG_*                                                          */
    int32_t reserved;
} fdtd_config;

/* One additive injection  F[box] += amp(step) [* profile(cell)] [/ divisor]
 * (sources/point.py:64-73, plane_wave.py:202-221, tfsf.py:314-411, gaussian.py:256-298, mode.py:255-361) */
typedef struct fdtd_source_op {
    int32_t component;
    int32_t lo[3], hi[3];   /* box in the component's index space (LOCAL x); z = [0,1) in 2-D */
    int32_t table;          /* column of the amplitude table (see fdtd_set_tables)             */
    const double* profile;  /* NULL, or box-shaped C-order host array (copied)                 */
    double  divisor;        /* applied only when profile != NULL; 1.0 = none                   */
    int32_t group;          /* ops sharing a group touch disjoint cells and may run concurrently;
                               groups run in ascending order (source list order, simulation.py:159) */
    int32_t reserved;       /* bit 0: GHOST op (x-slabs): the right neighbour's uniform injection on our ghost
                               planes [nx, nx+3); only the two-step sweep applies it (intermediate step)   */
} fdtd_source_op;

/* One sampled box (monitors/field.py:124-143, dft.py:121-160, flux.py:194-215). */
typedef struct fdtd_monitor_op {
    int32_t component;
    int32_t lo[3], hi[3];
    int32_t record;         /* !=0: keep the box for every step (FieldMonitor time_domain)     */
    int32_t n_freq;         /* >0: running DFT  acc[f] += F * phasor[step][f] * dt  (complex128) */
    int32_t phasor_col;     /* first column of this op's phasors in the phasor table            */
    int32_t reserved;
} fdtd_monitor_op;

/* One auxiliary-differential-equation recursion of a dispersive medium, cell-local, driven by E[component] after
 * every step (inside the fused sweeps: on the E values the NEXT step's sweep reads, i.e. the same numbers) (materials/ade.py:116-160 with the coefficients of materials/dispersion.py:189-336):
 *   kind 0 Lorentz  P+ = c0*E + c1*E + c2*P + c3*P_prev     kind 1 Drude  J+ = c0*E + c1*J     kind 2 Debye  P+ = c0*E + c1*P
 * mask (optional, box-shaped bytes) reproduces ADEManager.update_all's "E * mask" (ade.py:291-308).              */
typedef struct fdtd_ade_op {
    int32_t component;          /* 0..2 */
    int32_t kind;
    int32_t lo[3], hi[3];
    double  c0, c1, c2, c3;
    const uint8_t* mask;
} fdtd_ade_op;

/* ---- lifetime ------------------------------------------------------------------------------- */
int  fdtd_abi_version(void);
int  fdtd_struct_size(int32_t which);   /* sizeof: 0 fdtd_config, 1 fdtd_source_op, 2 fdtd_monitor_op, 3 fdtd_ade_op, 4 fdtd_shape */
const char* fdtd_last_error(void);
int  fdtd_create(const fdtd_config* cfg, fdtd_engine** out);
int  fdtd_destroy(fdtd_engine* e);

/* ---- coefficients (core/solver.py:113-133): cell-centred Ca,Cb,Da,Db ---------------------- */
int  fdtd_set_uniform_coeffs(fdtd_engine* e, double ca, double cb, double da, double db);
/* host fp64 arrays of shape (nx[+1], ny, nz) C-order; planes = nx, or nx+1 when a right
 * neighbour slab exists (the extra plane is the neighbour's first plane).                      */
int  fdtd_set_coeffs(fdtd_engine* e, const double* ca, const double* cb, const double* da,
                     const double* db, int32_t planes);

/* OPT-IN extension (no counterpart in the reference's solver, which applies one Cb per cell, core/solver.py:458-533):
 * per-component Cb for a DIAGONAL permittivity tensor — the case AnisotropicUpdater.update_e_from_curl_h handles at
 * function level (materials/tensor.py:508-514) — applied inside the E stage of the 3-D parity sweeps: Ex averages
 * cbx over (j, k), Ey averages cby over (i, k), Ez averages cbz over (i, j), exactly as the scalar Cb is averaged.
 * With cbx == cby == cbz the result is bit-identical to fdtd_set_coeffs.  3-D, not in physics mode.               */
int  fdtd_set_coeffs_aniso(fdtd_engine* e, const double* ca, const double* cbx, const double* cby,
                           const double* cbz, const double* da, const double* db, int32_t planes);

/* Geometry rasterisation on the device: shape list -> Ca, Cb, Da, Db in the engine's layout, replacing
 *   mask = Shape.rasterize(x, y, z)              geometry/shapes.py:71-99 (a 3 x cells fp64 meshgrid on the host)
 *   eps_rel[mask] = material.epsilon_r ...       user code, in list order (later shapes paint over earlier ones)
 *   MaxwellUpdater._compute_update_coefficients  core/solver.py:113-133
 * with the same fp64 operations in the same order (bit-identical masks and coefficients).                         */
typedef struct fdtd_shape {
    int32_t kind;           /* 0 Box :125-132, 1 Sphere :155-159, 2 Cylinder :194-214, 3 Polygon :249-283          */
    int32_t axis;           /* cylinder axis 0/1/2                                                                  */
    int32_t combine;        /* GeometryGroup :338-378: 0 start a new mask, 1 union, 2 intersection, 3 difference
                               of the running mask with this shape                                                  */
    int32_t paint;          /* != 0: cells of the running mask take this entry's material                           */
    double  center[3];
    double  a[3];           /* box: size/2; sphere: radius; cylinder: radius, height/2; polygon: z_min, z_max       */
    int32_t vert_first, vert_count;   /* polygon: its vertices in the (x, y) vertex list                            */
    double  eps_r[3];       /* eps_xx, eps_yy, eps_zz (all equal for an isotropic material)                         */
    double  mu_r, sigma_e, sigma_m;
} fdtd_shape;
/* x, y, z: host fp64 cell coordinates of length planes (nx, or nx+1 on a slab with a right neighbour), ny, nz
 * (z == NULL in 2-D: the reference rasterises 2-D grids at z = 0).  background = eps_xx, eps_yy, eps_zz, mu_r,
 * sigma_e, sigma_m of unpainted cells.  Any entry with unequal eps_r components switches the engine to per-component
 * Cb (see fdtd_set_coeffs_aniso; such entries need sigma_e == 0).                                                 */
int  fdtd_rasterize(fdtd_engine* e, const fdtd_shape* shapes, int32_t n_shapes, const double* verts_xy,
                    int32_t n_verts, const double* x, const double* y, const double* z, int32_t planes,
                    const double* background);
/* read-back of a device coefficient array as host fp64 (planes, ny, nz): which = 0 Ca, 1 Cb (Cb_x), 2 Da, 3 Db,
 * 4 Cb_y, 5 Cb_z                                                                                                   */
int  fdtd_download_coeffs(fdtd_engine* e, int32_t which, double* host, int32_t planes);

/* Physics mode only: working CPML (the reference's is a stub that is never applied, boundaries/pml.py:259-328).
 * coef = host fp64, axis by axis (x, y, z), 6 vectors of length N_axis each: b, a, 1/kappa at the E-update
 * (half-cell) positions, then at the H-update (integer) positions; identity outside the layer.
 * thickness = 0 removes the layer.  The 12 psi arrays are allocated only over the boundary slabs.          */
int  fdtd_set_cpml(fdtd_engine* e, int32_t thickness, const double* coef);

/* ---- fields (core/fields.py:64-114) --------------------------------------------------------- */
/* Upload replaces the OWNED planes [0, nx) of the current set.  On an x-slab (nx_global != nx) the ghost planes
 * nx.. are written by the right neighbour's halo push and are left alone; every rank must finish its uploads
 * before any rank calls fdtd_slab_run (one barrier between "upload" and "run").                             */
int  fdtd_upload_field(fdtd_engine* e, int32_t component, const void* host, int32_t host_dtype);
int  fdtd_download_field(fdtd_engine* e, int32_t component, void* host, int32_t host_dtype);
int  fdtd_zero_fields(fdtd_engine* e);
/* device address / geometry of a component's padded array, for zero-copy interop (torch, NCCL) */
int  fdtd_field_device_ptr(fdtd_engine* e, int32_t component, void** ptr, int64_t* plane_stride,
                           int64_t* row_stride, int64_t* planes_allocated);

/* ---- sources / monitors ---------------------------------------------------------------------- */
int  fdtd_clear_ops(fdtd_engine* e);
int  fdtd_add_source_op(fdtd_engine* e, const fdtd_source_op* op);
int  fdtd_add_monitor_op(fdtd_engine* e, const fdtd_monitor_op* op, int32_t* id);
int  fdtd_add_ade_op(fdtd_engine* e, const fdtd_ade_op* op, int32_t* id);
/* Extension (no reference counterpart: its FluxMonitor samples a corner patch and raises in 3-D): region-correct
 * instantaneous power sum((E x H)_normal) over a box valid for all six components, one fp64 sample per step.     */
int  fdtd_add_flux_op(fdtd_engine* e, int32_t direction, const int32_t* lo, const int32_t* hi, int32_t* id);
int  fdtd_download_flux(fdtd_engine* e, int32_t id, double* host, int32_t max_steps);
/* ADE state, host fp64 [cells of the box]: which = 0 current (P or J), 1 previous (Lorentz only) */
int  fdtd_download_ade(fdtd_engine* e, int32_t id, int32_t which, double* host);
int  fdtd_upload_ade(fdtd_engine* e, int32_t id, int32_t which, const double* host);
/* Per-step host-evaluated tables for steps [0, n_steps) of the NEXT fdtd_run calls:
 *   amp     [n_steps][n_amp]        fp64  source amplitudes  (waveform(t_n), sign, /377 baked in)
 *   phasors [n_steps][n_phasor][2]  fp64  exp(-j*2*pi*f*t_n) as (re, im)
 * Resets the engine's table cursor to 0 and sizes the record buffers for n_steps.             */
int  fdtd_set_tables(fdtd_engine* e, int32_t n_steps, int32_t n_amp, const double* amp,
                     int32_t n_phasor, const double* phasors);

/* ---- stepping ---------------------------------------------------------------------------------- */
int  fdtd_run(fdtd_engine* e, int32_t n_steps);   /* H pass, E pass, sources, monitors, n times     */
int  fdtd_update_h(fdtd_engine* e);               /* MaxwellUpdater.update_magnetic_fields :135-149 */
int  fdtd_update_e(fdtd_engine* e);               /* MaxwellUpdater.update_electric_fields :151-165 */
int  fdtd_sync(fdtd_engine* e);
/* switches: "tb2" 0/1 (two-step sweep), "fused_lx" planes per x-segment (0 = auto), "het_fused" 0/1,
 * "het_indexed" 0/1 (default 1: after fdtd_rasterize with a list of <= 64 entries the fused heterogeneous sweep reads ONE
 * material-index byte per cell and looks Ca..Db up in a shared-memory copy of the material table instead of streaming
 * the 4 (6) coefficient arrays — the same numbers, bit-identical results, 49 instead of 64 (72) B per fp32 cell-update;
 * host-supplied coefficient arrays always take the array path),
 * "yee_fused" 0/1/2 (physics mode: two-pass kernels / fused sweep / TMA-fed fused sweep, the default; set it before the
 * first step of a run), "ade_fused" 0/1 (dispersive-medium recursions applied inside the next fused sweep instead of
 * a kernel of their own; default 1, results identical), "ade_coupled" 0/1 (OPT-IN extension, not the reference's
 * behaviour: the recursions run at the beginning of each step and their polarisation current is subtracted in the same
 * E update, E+ -= Cb*eps0*(P+ - P)/dt or Cb*eps0*J+; needs a fused one-step sweep: 3-D, one GPU)                */
int  fdtd_set_option(fdtd_engine* e, const char* key, int32_t value);
/* measurement: CUDA events on the engine's own stream (torch.cuda.Event cannot see it).
 * fdtd_run_profiled runs n real steps without a graph and returns summed kernel times in ms:
 * out_ms[0] H pass (or the fused sweep), [1] E pass, [2] sources+monitors, [3] first-to-last event. */
int  fdtd_timer_start(fdtd_engine* e);
int  fdtd_timer_stop(fdtd_engine* e, double* elapsed_ms);
int  fdtd_run_profiled(fdtd_engine* e, int32_t n_steps, double* out_ms);

/* multi-GPU x-slabs: split entry points so the host can interleave the halo exchange.
 * phase 0 = H pass, 1 = E pass; part 0 = all planes but the last local one, 1 = last plane
 * (the only one that reads the right neighbour's ghost plane), 2 = everything.
 * stream = a cudaStream_t (0 = the engine's own stream).                                        */
int  fdtd_pass(fdtd_engine* e, int32_t phase, int32_t part, void* stream);
/* fused single-sweep step over local planes [i_begin, i_end): current set -> other set; flip != 0 makes the
 * other set current (last piece of a step).  Slabs: planes nx, nx+1 of the current set are ghosts that must
 * hold the right neighbour's planes 0, 1 before the piece containing plane nx-1 is launched.                */
int  fdtd_sweep(fdtd_engine* e, int32_t i_begin, int32_t i_end, int32_t flip, void* stream);
int  fdtd_post_step(fdtd_engine* e, void* stream);
/* Peer-to-peer slabs (one process per GPU).  Each rank exports a blob (CUDA IPC handles of its arrays and
 * flag words; call with blob == NULL to get the size), the host plumbing (torch.distributed) hands every rank
 * its LEFT neighbour's blob, and fdtd_slab_run then runs n steps with the 7 halo planes pushed by DMA into the
 * left neighbour's ghost planes over NVLink and release/acquire flags in peer memory — no host round trips,
 * one sweep launch per step (only the CTAs of the last x-segment wait for the halo, inside the kernel).      */
int  fdtd_ipc_export(fdtd_engine* e, void* blob, int32_t* nbytes);
int  fdtd_ipc_connect(fdtd_engine* e, const void* left_blob /* NULL on rank 0 */, int32_t has_right);
int  fdtd_slab_run(fdtd_engine* e, int32_t n_steps);
int  fdtd_slab_sync(fdtd_engine* e);   /* sources + monitors + cursor advance       */
/* first local plane (send side) / first ghost plane (receive side) of a component IN THE CURRENT SET;
 * plane 1 / the second ghost plane follow contiguously at +plane_bytes                                    */
int  fdtd_halo_ptrs(fdtd_engine* e, int32_t component, void** first_plane, void** ghost_plane,
                    int64_t* plane_bytes);

/* ---- monitor read-out ---------------------------------------------------------------------------- */
/* records: host fp64 [steps_run][cells]; dft: host complex128 [n_freq][cells] */
int  fdtd_download_records(fdtd_engine* e, int32_t monitor_id, double* host, int32_t max_steps);
int  fdtd_download_dft(fdtd_engine* e, int32_t monitor_id, double* host);
int  fdtd_upload_dft(fdtd_engine* e, int32_t monitor_id, const double* host);

/* Cropped read-out: box [lo, hi) of a component's CURRENT array as dense C-order host fp64 (at most 2^27 cells).  The
 * reference has no counterpart (its arrays are host NumPy: fields[c][box]); used by bench.py's self-check and by parity
 * tests at sizes whose full arrays the host would not want to mirror.                                           */
int  fdtd_download_box(fdtd_engine* e, int32_t component, const int32_t* lo, const int32_t* hi, double* host);
/* Per-plane checksums of the logical cells of a component (planes = its local x extent): out[2p] = sum of the value
 * bit patterns, out[2p+1] = sum of bits * (1 + j*n2 + k), modulo 2^64.  Independent of the summation order and of the
 * x-slab decomposition, so the concatenation over ranks is identical at 1/2/4/8 GPUs iff the fields are.          */
int  fdtd_field_checksum(fdtd_engine* e, int32_t component, uint64_t* out, int32_t planes);

/* Mode-overlap numerator on the RESIDENT DFT planes of six monitor ops that share one plane box and one frequency list
 * (monitor_ids in component order Ex..Hz; the two normal components are not read):
 *   out[f] = 0.5 * sum_cells [ (E_sim x conj(H_mode))_n + (E_mode x conj(H_sim))_n ]     (re, im), f = 0..n_freq-1
 * i.e. the sum of utils/mode_matching.py:104-118 before "* dx * dy / mode_power", which the caller applies (the mode's own
 * power is a host-side constant).  direction 0/1/2 = plane normal x/y/z; mode = host complex128 [6][cells] (Ex..Hz on the
 * cells of the box, C order).  Replaces downloading 6 x n_freq planes to evaluate the sum in NumPy.              */
int  fdtd_mode_overlap(fdtd_engine* e, const int32_t* monitor_ids, int32_t direction, const double* mode, double* out);

/* ---- introspection ----------------------------------------------------------------------------------- */
int  fdtd_steps_done(fdtd_engine* e, int64_t* steps);
int  fdtd_kernel_launches(fdtd_engine* e, int64_t* launches); /* kernels launched by this handle  */
int  fdtd_mem_info(fdtd_engine* e, int64_t* free_bytes, int64_t* total_bytes);
/* device-level services of the backend object (reference: Backend.synchronize / get_memory_info, backends/base.py:188-219)
 * and page-locked host memory for the NumPy mirrors that Backend.zeros/ones/empty hand out (fields.py:70)            */
int  fdtd_device_mem_info(int32_t device, int64_t* free_bytes, int64_t* total_bytes);
int  fdtd_device_sync(int32_t device);
int  fdtd_host_alloc(int64_t bytes, void** ptr);
int  fdtd_host_free(void* ptr);

/* Anisotropic tensor update on caller-supplied HOST arrays of n elements each (function level, not a stage of the
 * step — the reference never couples it either): replaces AnisotropicUpdater.update_e_from_curl_h and
 * update_h_from_curl_e, /root/reference/src/prismo/materials/tensor.py:482-536 and :538-588.
 *   mode & 1 == 0 : out[c] = f[c] (+|-) ((scale * curl[c]) / d_c),        negative != 0 selects "-" (H update)
 *   mode & 1 == 1 : out[c] = f[c] + scale * ((r_c0*curl[0] + r_c1*curl[1]) + r_c2*curl[2])   (pass -dt/mu0 for H)
 *   mode & 2 / mode & 4 (diagonal, dtype f64): round scale*curl / the division in float32 first — what NumPy's
 *   promotion does when a float32 curl meets a float64 field or tensor array (values staged as exact f64 copies).
 * coef = d_x,d_y,d_z (diagonal) or the row-major INVERSE tensor (9, inverted by the caller like the reference does);
 * coef_arrays (may be NULL) holds optional per-cell arrays of the same dtype overriding single entries.  f[c] == NULL
 * skips component c.  Every operation is rounded separately in the reference's order: bit-exact in fp64 and fp32. */
int  fdtd_tensor_update(int32_t device, int32_t dtype, int64_t n, const void* const* f, const void* const* curl,
                        void* const* out, double scale, int32_t negative, int32_t mode, const double* coef,
                        const void* const* coef_arrays);

/* Host-only (no device needed): the x-segment plan of one two-step sweep over nx planes.  plane_flags[p] != 0 marks
 * planes that carry source / monitor ops (n_flags may exceed nx by the ghost planes of a slab); tiles = (j,k) tiles
 * per segment; halo != 0: the slab reads ghost planes; fused_lx > 0 forces the bulk part length; zones_mode -1 auto,
 * 0 off, 1 on.  Writes segments in dispatch order ([lo, hi), ops flag) and returns their count (< 0: error).
 * No reference counterpart (the reference has no tiling); exported for the planner's unit tests. */
int  fdtd_plan_segments(int32_t nx, const uint8_t* plane_flags, int32_t n_flags, int64_t tiles, int32_t halo,
                        int32_t fused_lx, int32_t zones_mode, int32_t* seg_lo, int32_t* seg_hi, int32_t* seg_ops,
                        int32_t max_segs);

#ifdef __cplusplus
}
#endif
#endif /* FDTD_B200_H */
