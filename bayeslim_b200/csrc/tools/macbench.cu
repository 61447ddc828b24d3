// Stand-alone probe of the 8x8 complex multiply-accumulate register tile (ant_kernels.cu):
// how close do W warps per SM sub-partition get to the packed-FP32 peak, with and without the
// shared-memory operand loads?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I.. macbench.cu
#include <cstdio>
#include "../rime_math.cuh"
using namespace b200rime;

template <bool CONJ>
__device__ __forceinline__ void mac8(P2 (&aR)[8][4], P2 (&aI)[8][4], const float4 (&x)[4],
                                     const float4 (&yr)[2], const float4 (&yi)[2]) {
    const P2 YR[4] = {p2(yr[0].x, yr[0].y), p2(yr[0].z, yr[0].w), p2(yr[1].x, yr[1].y), p2(yr[1].z, yr[1].w)};
    const P2 YI[4] = {p2(yi[0].x, yi[0].y), p2(yi[0].z, yi[0].w), p2(yi[1].x, yi[1].y), p2(yi[1].z, yi[1].w)};
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const float xr[2] = {x[h].x, x[h].z}, xi[2] = {x[h].y, x[h].w};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = 2 * h + e;
            const P2 XR = p2(xr[e], xr[e]), XI = p2(xi[e], xi[e]), NXI = p2(-xi[e], -xi[e]);
#pragma unroll
            for (int j = 0; j < 4; ++j) { p2_mac(aR[i][j], XR, YR[j]); p2_mac(aI[i][j], XR, YI[j]); }
#pragma unroll
            for (int j = 0; j < 4; ++j) { p2_mac(aR[i][j], CONJ ? XI : NXI, YI[j]); p2_mac(aI[i][j], CONJ ? NXI : XI, YR[j]); }
        }
    }
}

// MODE 0: operands from shared memory every step (product loop); 1: operands loaded once per
// stage (1/8 of the LDS traffic); 2: operands loaded once per kernel
template <int MODE, int REGS>
__global__ void __launch_bounds__(384, 1) mac_kernel(float* out, int nstage, int nwarps) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 4 * 512; i += blockDim.x)
        sm[i] = make_float4(1e-3f * (i & 7), 1e-3f, -1e-3f, 2e-3f * (i & 3));
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= 8) { setmaxnreg_dec<72>(); return; }
    setmaxnreg_inc<REGS>();
    if (warp >= nwarps) return;
    P2 aR[8][4], aI[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) aR[i][j] = aI[i][j] = p2(0.f, 0.f);
    const int ti = lane & 7, yq = (warp >> 2) * 8 + (lane >> 3) * 2;
    float4 x[4], yr[2], yi[2];
    for (int it = 0; it < nstage; ++it) {
        const float4* X4 = sm + (it & 3) * 512;
        const float4* YR4 = X4 + 256;
        const float4* YI4 = X4 + 384;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (MODE == 0 || (MODE == 1 && r == 0) || (MODE == 2 && r == 0 && it == 0)) {
#pragma unroll
                for (int h = 0; h < 4; ++h) x[h] = X4[r * 32 + h * 8 + ti];
                yr[0] = YR4[r * 16 + yq]; yr[1] = YR4[r * 16 + yq + 1];
                yi[0] = YI4[r * 16 + yq]; yi[1] = YI4[r * 16 + yq + 1];
            }
            mac8<true>(aR, aI, x, yr, yi);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { float a, b; p2_get(aR[i][j], a, b); s += a + b; p2_get(aI[i][j], a, b); s += a + b; }
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int MODE, int REGS> void run(const char* name, int nwarps, float* out) {
    const int nstage = 20000;
    cudaFuncSetAttribute(mac_kernel<MODE, REGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 8192);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mac_kernel<MODE, REGS><<<148, 384, 4 * 8192>>>(out, 100, nwarps);
    cudaEventRecord(e0);
    mac_kernel<MODE, REGS><<<148, 384, 4 * 8192>>>(out, nstage, nwarps);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 148.0 * nwarps * 32 * (double)nstage * 1024 * 4;
    printf("{\"probe\": \"%s\", \"warps\": %d, \"ms\": %.3f, \"tflops\": %.2f, \"err\": \"%s\"}\n", name, nwarps, ms,
           flops / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float* out; cudaMalloc(&out, 148 * 256 * 4);
    run<0, 216>("lds_every_step", 8, out);
    run<0, 216>("lds_every_step", 4, out);
    run<1, 216>("lds_once_per_stage", 8, out);
    run<1, 216>("lds_once_per_stage", 4, out);
    run<2, 216>("no_lds", 8, out);
    run<2, 216>("no_lds", 4, out);
    return 0;
}
