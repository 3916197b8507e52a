// Probe: mma.sync throughput per SM on sm_100a for tf32 m16n8k8, bf16 m16n8k16, fp16 m16n8k16 (legacy tensor path).
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
template <int KIND>
__global__ void k(float* out, int iters, long long* cyc) {
    float c[4][4] = {};
    uint32_t a[4] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u}, b0 = 0x3f800000u, b1 = 0x3f800000u;
    if (KIND != 0) { a[0] = a[1] = a[2] = a[3] = 0x3f803f80u; b0 = b1 = 0x3f803f80u; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(b0));
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < 4; ++j) for (int e = 0; e < 4; ++e) s += c[j][e];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int KIND>
void run(const char* name, int warps) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    int iters = 20000;
    k<KIND><<<148, warps * 32>>>(out, 100, cyc); cudaDeviceSynchronize();
    k<KIND><<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_warp = (double)h / (iters * 4.0);
    printf("%s warps/SM=%d: %.2f cycles per mma per warp; per SMSP issue interval %.2f cycles\n", name, warps, per_warp, per_warp / (warps / 4.0 > 1 ? warps / 4.0 : 1));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {1, 4, 8, 16}) { run<0>("tf32 m16n8k8", w); }
    for (int w : {1, 4, 8, 16}) { run<1>("bf16 m16n8k16", w); }
    for (int w : {1, 4, 8, 16}) { run<2>("tf32 m16n8k4", w); }
    return 0;
}
