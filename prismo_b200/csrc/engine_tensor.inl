// anisotropic tensor update on caller-supplied arrays (row a23)
// (textual part of fdtd_engine.cu — one translation unit; not compiled on its own)
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
