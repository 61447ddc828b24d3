// Shared definitions for the B200 RIME kernels: layout constants, seed/phase math
// (host+device so the numerics can be exercised on a CPU, see csrc/emulate.cu),
// mbarrier / bulk-copy (TMA) wrappers.
#pragma once
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

namespace b200rime {

constexpr double C_LIGHT = 2.99792458e8;   // reference telescope_model.py:355
constexpr int SRC_TILE = 64;               // sources per shared-memory stage (fwd, bwd_bl)
constexpr int SRC_PAD = 128;               // per-time padding of the packed source axis
constexpr int BL_TILE = 64;                // baselines per shared-memory stage (bwd_sky)
constexpr int FWD_THREADS = 128;           // thread <-> baseline
constexpr int SKY_THREADS = 128;           // thread <-> source
constexpr int BL_SEGMENT = 4096;           // baselines per fp32 accumulation segment (bwd_sky)

template <typename T> struct Cfg;
#ifndef B200_KC_F32
#define B200_KC_F32 64
#endif
#ifndef B200_MINB_F32
#define B200_MINB_F32 2
#endif
#ifndef B200_MINB_SKY_F32
#define B200_MINB_SKY_F32 3
#endif
template <> struct Cfg<float> {
    static constexpr int KC = B200_KC_F32;
    typedef float2 cplx;
};
template <> struct Cfg<double> {
    static constexpr int KC = 32;
    typedef double2 cplx;
};

// ---------------------------------------------------------------------------------
// Phase seeds.  phase (cycles) of source s on baseline b at frequency nu is
//   p = sgn * (b . shat) * nu / c.
// u = b . shat is formed in float64 (|p| reaches ~10^3 cycles; float32 would lose the
// fraction).  The fraction of a cycle is extracted in float64 with the 1.5*2^52
// round-to-nearest trick fused into two FMAs, then handed to float32 trigonometry.
// ---------------------------------------------------------------------------------
#define B200_MAGIC 6755399441055744.0

__host__ __device__ __forceinline__ double frac_cycles(double u, double k) {
    // u*k - rint(u*k), with a single rounding of the product in each FMA
#ifdef __CUDA_ARCH__
    double r = __dadd_rn(__fma_rn(u, k, B200_MAGIC), -B200_MAGIC);
    return __fma_rn(u, k, -r);
#else
    double r = std::fma(u, k, B200_MAGIC) - B200_MAGIC;
    return std::fma(u, k, -r);
#endif
}

// exp(2 pi i p) for |p| <= 0.5 cycles.
// "fast": MUFU.SIN/COS (abs err ~4e-7) -- used for chunk seeds, whose error is not amplified.
// "accurate": ~1 ulp -- used for the per-channel step w, whose angle error is multiplied by
// up to KC/2 along the recurrence.
__host__ __device__ __forceinline__ void cis_fast(float p, float& c, float& s) {
#ifdef __CUDA_ARCH__
    __sincosf(p * 6.283185307179586f, &s, &c);
#else
    s = sinf(p * 6.283185307179586f);
    c = cosf(p * 6.283185307179586f);
#endif
}
__host__ __device__ __forceinline__ void cis_accurate(float p, float& c, float& s) {
#ifdef __CUDA_ARCH__
    sincospif(2.0f * p, &s, &c);
#else
    s = (float)sin(2.0 * M_PI * (double)p);
    c = (float)cos(2.0 * M_PI * (double)p);
#endif
}
__host__ __device__ __forceinline__ void cis_accurate(double p, double& c, double& s) {
#ifdef __CUDA_ARCH__
    sincospi(2.0 * p, &s, &c);
#else
    s = sin(2.0 * M_PI * p);
    c = cos(2.0 * M_PI * p);
#endif
}

// seeds for one (baseline, source, chunk): z = exp(2 pi i u k_mid), w = exp(2 pi i u k_step)
__host__ __device__ __forceinline__ void chunk_seed(double u, double k_mid, double k_step,
                                                    float& zr, float& zi, float& wr, float& wi) {
    float pf = (float)frac_cycles(u, k_mid);
    float qf = (float)frac_cycles(u, k_step);
    cis_fast(pf, zr, zi);
    cis_accurate(qf, wr, wi);
}
__host__ __device__ __forceinline__ void chunk_seed(double u, double k_mid, double k_step,
                                                    double& zr, double& zi, double& wr, double& wi) {
    cis_accurate(frac_cycles(u, k_mid), zr, zi);
    cis_accurate(frac_cycles(u, k_step), wr, wi);
}
// direct phase of one channel (non-uniform frequency path)
__host__ __device__ __forceinline__ void channel_cis(double u, double k, float& zr, float& zi) {
    cis_fast((float)frac_cycles(u, k), zr, zi);
}
__host__ __device__ __forceinline__ void channel_cis(double u, double k, double& zr, double& zi) {
    cis_accurate(frac_cycles(u, k), zr, zi);
}

template <typename T>
__host__ __device__ __forceinline__ void rot(T& zr, T& zi, T wr, T wi) {   // z *= w
    T t1 = zi * wi;
    T t2 = zi * wr;
    T nr = zr * wr - t1;
    T ni = zr * wi + t2;
    zr = nr;
    zi = ni;
}
template <typename T>
__host__ __device__ __forceinline__ void rotc(T& zr, T& zi, T wr, T wi) {  // z *= conj(w)
    T t1 = zi * wi;
    T t2 = zr * wi;
    T nr = zr * wr + t1;
    T ni = zi * wr - t2;
    zr = nr;
    zi = ni;
}

// per-chunk frequency constants: k_mid = sgn*nu(chunk centre)/c, k_step = sgn*dnu/c [cycles/m]
struct ChunkFreq {
    double k_mid, k_step;
};
__host__ __device__ __forceinline__ ChunkFreq chunk_freq(const double* freqs, int nfreq, int chunk,
                                                         int KC, double sgn_over_c) {
    // channel spacing from the chunk's end points (halves the rounding noise of a
    // neighbour difference); centre frequency read directly when the chunk is full enough
    const int k0 = chunk * KC;
    const int kl = (k0 + KC - 1 < nfreq - 1) ? (k0 + KC - 1) : (nfreq - 1);
    double df = 0.0;
    if (kl > k0) df = (freqs[kl] - freqs[k0]) / (double)(kl - k0);
    else if (nfreq > 1) df = (freqs[nfreq - 1] - freqs[0]) / (double)(nfreq - 1);
    const double fmid = (k0 + KC / 2 <= kl) ? freqs[k0 + KC / 2] : freqs[k0] + (KC / 2) * df;
    ChunkFreq c;
    c.k_mid = sgn_over_c * fmid;
    c.k_step = sgn_over_c * df;
    return c;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) wrappers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// wait that yields the issue slots: try_wait with a suspend-time hint, then a short sleep.  For
// warps that run far ahead of their partners (producers waiting for a free stage): a bare
// try_wait loop spins at full rate and takes issue cycles from the consumer warps of the same
// SM sub-partition.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(2000u)
            : "memory");
        if (done) break;
        __nanosleep(200);
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// warpgroup register reallocation (producer / consumer warp specialisation)
template <int N> __device__ __forceinline__ void setmaxnreg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N> __device__ __forceinline__ void setmaxnreg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
// barrier among the first `nthreads` threads of the CTA only (id 1..15)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
#endif  // __CUDACC__

}  // namespace b200rime
