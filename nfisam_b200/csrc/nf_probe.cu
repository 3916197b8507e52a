// Pipe-peak probes used by bench.py for the roofline denominators of the flow kernels, which are
// bound by the FP32 FMA and MUFU (SFU) pipes rather than by HBM or the tensor cores.
#include "nf_internal.h"

namespace {

// FFMA with two constant-bank operands (the cheapest encoding)
__global__ void __launch_bounds__(256) nf_probe_fma_const_kernel(float* out, int iters, float a, float b) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = threadIdx.x + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA with three register operands (what an MLP inner product issues)
__global__ void __launch_bounds__(256) nf_probe_fma_reg_kernel(float* out, int iters, const float* __restrict__ src) {
    float v[8], a[4], b[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = threadIdx.x + j;
#pragma unroll
    for (int j = 0; j < 4; ++j) { a[j] = src[j]; b[j] = src[4 + j]; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], a[(j + u) & 3], b[(j + 2 * u) & 3]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed FFMA2 (fma.rn.f32x2, new on sm_100)
__global__ void __launch_bounds__(256) nf_probe_fma2_kernel(float* out, int iters, const float* __restrict__ src) {
    float2 v[8], a[4], b[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = make_float2(threadIdx.x + j, threadIdx.x - j);
#pragma unroll
    for (int j = 0; j < 4; ++j) { a[j] = make_float2(src[j], src[j + 1]); b[j] = make_float2(src[4 + j], src[5 + j]); }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ffma2_rn(v[j], a[(j + u) & 3], b[(j + 2 * u) & 3]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j].x + v[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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

extern "C" int nfisam_probe_pipe_peaks(int device, double* peaks4) {
    if (!peaks4) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    int prev = 0;
    cudaGetDevice(&prev);
    NF_CUDA(cudaSetDevice(device));
    const int blocks = nf_sm_count(device) * 8, threads = 256;
    float* buf = nullptr;
    float* src = nullptr;
    NF_CUDA(cudaMalloc(&buf, sizeof(float) * (size_t)blocks * threads));
    NF_CUDA(cudaMalloc(&src, sizeof(float) * 16));
    const float hsrc[16] = {0.999f, 0.998f, 0.997f, 0.996f, 1e-3f, 2e-3f, 3e-3f, 4e-3f, 5e-3f, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpy(src, hsrc, sizeof(hsrc), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 2048;
    double best[4] = {0, 0, 0, 0};
    for (int rep = 0; rep < 4; ++rep) {
        for (int which = 0; which < 4; ++which) {
            float ms = 0.f;
            cudaEventRecord(e0);
            double ops = 0.0;
            switch (which) {
                case 0: nf_probe_fma_reg_kernel<<<blocks, threads>>>(buf, iters, src); ops = 2.0 * 8 * 16; break;
                case 1: nf_probe_fma2_kernel<<<blocks, threads>>>(buf, iters, src); ops = 4.0 * 8 * 16; break;
                case 2: nf_probe_fma_const_kernel<<<blocks, threads>>>(buf, iters, 0.999f, 0.001f); ops = 2.0 * 8 * 16; break;
                default: nf_probe_mufu_kernel<<<blocks, threads>>>(buf, iters); ops = 4.0 * 16; break;
            }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            const double rate = ops * (double)iters * blocks * threads / (ms * 1e-3);
            if (rep > 0 && rate > best[which]) best[which] = rate;
        }
    }
    nf_count_launch(16);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    cudaFree(src);
    cudaSetDevice(prev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return nf_cuda_fail(e, "pipe probe");
    peaks4[0] = best[0] * 1e-12;
    peaks4[1] = best[1] * 1e-12;
    peaks4[2] = best[2] * 1e-12;
    peaks4[3] = best[3] * 1e-9;
    return NF_OK;
}
