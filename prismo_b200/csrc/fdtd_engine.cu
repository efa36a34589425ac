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
#include "fdtd_yee_fused.cuh"
#include "fdtd_tb2.cuh"
#include "fdtd_tb2x.cuh"
#include "fdtd_yeex.cuh"
#include "fdtd_het.cuh"
#include "fdtd_tensor.cuh"
#include "fdtd_raster.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "engine_state.inl"   // engine state, error reporting, small helpers
#include "engine_setup.inl"   // ABI version, create / destroy, coefficients, field upload / download
#include "engine_ops.inl"   // source / monitor / flux / ADE ops, tables
#include "engine_launch.inl"   // kernel launch helpers: two-pass, fused, heterogeneous, two-step sweep and its segment planner
#include "engine_run.inl"   // step loops: fdtd_run (CUDA graph), half steps, profiled run, options, split entry points
#include "engine_tensor.inl"   // anisotropic tensor update on caller-supplied arrays (row a23)
#include "engine_slab.inl"   // peer-to-peer x-slabs: CUDA-IPC export / connect, fdtd_slab_run
#include "engine_raster.inl"   // geometry rasterisation on the device: shape list -> coefficient arrays (row f4)
#include "engine_readout.inl"   // monitor read-out and introspection
