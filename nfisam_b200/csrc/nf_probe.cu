// Pipe-peak probes used by bench.py for the roofline denominators of the flow kernels, which are
// bound by the FP32 FMA and MUFU (SFU) pipes rather than by HBM or the tensor cores.
#include "nf_internal.h"

namespace {

__global__ void __launch_bounds__(256) nf_probe_fma_kernel(float* out, int iters, float a, float b) {
    float v0 = threadIdx.x, v1 = v0 + 1.f, v2 = v0 + 2.f, v3 = v0 + 3.f, v4 = v0 + 4.f, v5 = v0 + 5.f, v6 = v0 + 6.f,
          v7 = v0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            v0 = fmaf(v0, a, b); v1 = fmaf(v1, a, b); v2 = fmaf(v2, a, b); v3 = fmaf(v3, a, b);
            v4 = fmaf(v4, a, b); v5 = fmaf(v5, a, b); v6 = fmaf(v6, a, b); v7 = fmaf(v7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
}

__global__ void __launch_bounds__(256) nf_probe_mufu_kernel(float* out, int iters) {
    float v0 = threadIdx.x * 1e-3f, v1 = v0 + .1f, v2 = v0 + .2f, v3 = v0 + .3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            v0 = exp2f(-fabsf(v0)); v1 = exp2f(-fabsf(v1)); v2 = exp2f(-fabsf(v2)); v3 = exp2f(-fabsf(v3));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = v0 + v1 + v2 + v3;
}

}  // namespace

extern "C" int nfisam_probe_pipe_peaks(int device, double* fp32_tflops, double* mufu_gops) {
    if (!fp32_tflops || !mufu_gops) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    int prev = 0;
    cudaGetDevice(&prev);
    NF_CUDA(cudaSetDevice(device));
    const int blocks = nf_sm_count(device) * 8, threads = 256;
    float* buf = nullptr;
    NF_CUDA(cudaMalloc(&buf, sizeof(float) * (size_t)blocks * threads));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0.f;
    const int iters = 2048;
    double best_f = 0.0, best_m = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        nf_probe_fma_kernel<<<blocks, threads>>>(buf, iters, 0.999f, 0.001f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        if (rep > 0 && flops / (ms * 1e-3) > best_f) best_f = flops / (ms * 1e-3);
        cudaEventRecord(e0);
        nf_probe_mufu_kernel<<<blocks, threads>>>(buf, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = 4.0 * 16 * (double)iters * blocks * threads;
        if (rep > 0 && ops / (ms * 1e-3) > best_m) best_m = ops / (ms * 1e-3);
    }
    nf_count_launch(8);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    cudaSetDevice(prev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return nf_cuda_fail(e, "pipe probe");
    *fp32_tflops = best_f * 1e-12;
    *mufu_gops = best_m * 1e-9;
    return NF_OK;
}
