// dp_rate.cu — measures the FP64 FMA issue rate of the GPU (lanes per clock per SM), to bound the exact-fp64 kernels.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dp_rate tools/dp_rate.cu && ./dp_rate
#include <cstdio>
#include <cuda_runtime.h>
template <typename T> __global__ void k(T* out, int n, T a, T b)
{
    T x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < n; ++i) {
        x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
        x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
template <typename T> static void run(const char* name)
{
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    T* out; cudaMalloc(&out, sizeof(T) * sms * 8 * 512);
    const int n = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<T><<<sms * 8, 512>>>(out, 100, (T)1.0000001, (T)1e-9);
    cudaEventRecord(e0);
    k<T><<<sms * 8, 512>>>(out, n, (T)1.0000001, (T)1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)sms * 8 * 512 * 8.0 * n;
    printf("%s: %.2f T FMA/s = %.1f lanes/clk/SM at the nominal %d MHz (%d SMs)\n", name, fma / ms / 1e9, fma / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000, sms);
    cudaFree(out);
}
int main() { run<float>("fp32"); run<double>("fp64"); return 0; }
