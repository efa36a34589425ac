// Anisotropic (tensor) material update at FUNCTION level — row a23 of the scope table:
//   AnisotropicUpdater.update_e_from_curl_h / update_h_from_curl_e   /root/reference/src/prismo/materials/tensor.py:482-588
// The reference never calls these inside the step (co-located arrays supplied by the caller), so this is an operator on
// caller-supplied arrays, not a stage of the sweep.  Every operation is rounded separately in NumPy's order:
//   diagonal : out_c = f_c  (+|-)  ((s * curl_c) / d_c)                                  (tensor.py:513-517, :561-564)
//   full     : out_c = f_c  +  ((+|-)s) * ((r_c0*curl_0 + r_c1*curl_1) + r_c2*curl_2)    (tensor.py:519-541, :566-586)
// with s = dt/eps0 (dt/mu0), d = diagonal tensor entries, r = rows of the inverse tensor (inverted on the host by NumPy,
// exactly as the reference does).  Entries are uniform scalars or per-cell arrays.
#pragma once
#include <cuda_runtime.h>

namespace fdtd {

template <typename T> struct Rn;
template <> struct Rn<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};
template <> struct Rn<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};

template <typename T> struct TensorArgs {
    const T* f[3];          // field triple (null: component skipped)
    const T* curl[3];
    T* out[3];
    const T* coef_arr[9];   // per-cell entries (null: uniform coef[] applies); diagonal uses [0..2]
    T coef[9];              // diagonal: d_x, d_y, d_z; full: inverse tensor, row-major
    T s;                    // dt/eps0 or dt/mu0 (full + negative: already negated, like the reference's -(dt/mu0))
    int negative;           // diagonal only: 1 = subtract (H update)
    int full;
    // NumPy's promotion of mixed inputs (float32 curl with a float64 field or tensor array): the reference then rounds
    // s*curl — and, with weak Python-float entries, the division too — in float32 before widening.  Diagonal, T = double.
    int mul_f32, div_f32;
};

template <typename T> __device__ __forceinline__ T mul_maybe_f32(T s, T c, int) { return Rn<T>::mul(s, c); }
template <> __device__ __forceinline__ double mul_maybe_f32<double>(double s, double c, int f32)
{
    return f32 ? (double)__fmul_rn((float)s, (float)c) : __dmul_rn(s, c);
}
template <typename T> __device__ __forceinline__ T div_maybe_f32(T t, T d, int) { return Rn<T>::div(t, d); }
template <> __device__ __forceinline__ double div_maybe_f32<double>(double t, double d, int f32)
{
    return f32 ? (double)__fdiv_rn((float)t, (float)d) : __ddiv_rn(t, d);
}

template <typename T>
__global__ void k_tensor_update(TensorArgs<T> a, long long n)
{
    typedef Rn<T> R;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
        if (!a.full) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (!a.f[c]) continue;
                const T d = a.coef_arr[c] ? a.coef_arr[c][q] : a.coef[c];
                const T t = div_maybe_f32<T>(mul_maybe_f32<T>(a.s, a.curl[c][q], a.mul_f32), d, a.div_f32);
                a.out[c][q] = a.negative ? R::sub(a.f[c][q], t) : R::add(a.f[c][q], t);
            }
        } else {
            const T c0 = a.curl[0][q], c1 = a.curl[1][q], c2 = a.curl[2][q];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (!a.f[c]) continue;
                const T r0 = a.coef_arr[3 * c] ? a.coef_arr[3 * c][q] : a.coef[3 * c];
                const T r1 = a.coef_arr[3 * c + 1] ? a.coef_arr[3 * c + 1][q] : a.coef[3 * c + 1];
                const T r2 = a.coef_arr[3 * c + 2] ? a.coef_arr[3 * c + 2][q] : a.coef[3 * c + 2];
                const T sum = R::add(R::add(R::mul(r0, c0), R::mul(r1, c1)), R::mul(r2, c2));
                a.out[c][q] = R::add(a.f[c][q], R::mul(a.s, sum));
            }
        }
    }
}

}  // namespace fdtd
