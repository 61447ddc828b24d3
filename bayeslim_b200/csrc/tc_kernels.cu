// Tensor-core (tcgen05 / TMEM) fringe-sum kernels for sm_100a: the antenna-factorised source sum
// as a batched complex GEMM on the 5th-generation tensor cores.
//
// The fringe of baseline (i, j) is conj(E_i) E_j with E_a = exp(2 pi i sgn r_a.shat nu / c)
// (ant_kernels.cu), so per channel and time V = E^H diag(A) E is a Hermitian rank-Ns update:
// M = N = antennas, K = sources.  This file runs it on tcgen05.mma with FP32 accumulators in
// tensor memory:
//
//   * operands are generated on the fly inside the CTA (float64 r_a.shat, the phase fraction
//     taken straight from the low mantissa word of one DFMA, MUFU sine / cosine) and written to
//     shared memory in the UMMA canonical K-major layout -- nothing fringe-shaped, per baseline
//     or per antenna, reaches HBM;
//   * every operand is split into two float16 numbers hi + lo (|E| <= 1 and A E scaled by a power
//     of two into float16 range, so hi + lo carries 22 mantissa bits) and a real product costs
//     three MMAs (hi.hi + hi.lo + lo.hi; the dropped lo.lo term is 2^-24 relative): float32-grade
//     results at one third of the float16 tensor rate, which is still ~7x the FP32 FMA pipes;
//   * the complex product needs four real ones; the minus sign of Im = Er.Yi - Ei.Yr is the
//     negate-A bit of the instruction descriptor, so X and Y are stored once.
//
// Forward item = (block of 128 first antennas i0.., range of N <= 256 second antennas j0..,
// channel k, unit of sources).  Warps 0..11 are producers (thread <-> operand row = antenna,
// 16 sources per stage = one UMMA K step), warp 12 issues the MMAs (one elected lane) and owns
// the TMEM allocation; full / empty mbarriers per stage, tcgen05.commit releases stages and
// signals the epilogue, in which the producer warps read the accumulators back with tcgen05.ld
// and scatter them through the antenna-pair table into Vpart[unit][baseline][channel].
//
// Replaces telescope_model.py:310-358 (gen_fringe) + rime_model.py:426-429 (multiply, sum).
#include <cuda_fp16.h>
#include "rime_math.cuh"
#include "internal.h"

namespace b200rime {

constexpr int TC_M = 128;            // X rows (first antennas) per item = UMMA M
constexpr int TC_NMAX = 256;         // Y rows (second antennas) per item <= UMMA N max
constexpr int TC_KS = 16;            // sources per stage = one kind::f16 UMMA K step
constexpr int TC_NSTAGE = 4;
constexpr int TC_PROD_WARPS = (TC_M + TC_NMAX) / 32;         // 12
constexpr int TC_THREADS = (TC_PROD_WARPS + 1) * 32;         // + the MMA warp
constexpr int TC_KC = B200_KC_F32;
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_IM_COL = 256;       // column offset of the imaginary accumulator

struct TcSmem {
    // one operand array = rows x 16 float16 in the canonical no-swizzle K-major layout:
    //   byte offset(row, kgroup of 8) = (row / 8) * 256 + kgroup * 128 + (row % 8) * 16
    // i.e. 8 x 16-byte core matrices, LBO (K direction) = 128 B, SBO (row direction) = 256 B
    static constexpr int X_ARR = TC_M * TC_KS * 2;            // 4 KB
    static constexpr int Y_ARR = TC_NMAX * TC_KS * 2;         // 8 KB
    static constexpr int XR_H = 0, XR_L = X_ARR, XI_H = 2 * X_ARR, XI_L = 3 * X_ARR;
    static constexpr int YR_H = 4 * X_ARR, YR_L = YR_H + Y_ARR, YI_H = YR_H + 2 * Y_ARR,
                         YI_L = YR_H + 3 * Y_ARR;
    static constexpr int STAGE = 4 * X_ARR + 4 * Y_ARR;       // 48 KB
    static constexpr int BAR_OFF = TC_NSTAGE * STAGE;         // full[NSTAGE], empty[NSTAGE], done
    static constexpr int TMEM_OFF = BAR_OFF + (2 * TC_NSTAGE + 1) * 8;
    static constexpr int TOTAL = TMEM_OFF + 16;
};

// ---------------------------------------------------------------------------------------
// tcgen05 wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr) {
    // shared-memory matrix descriptor, no swizzle, K-major: start address, LBO = 128 B,
    // SBO = 256 B (all >> 4), descriptor version 1 (Blackwell) in bits 46..47
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(128 >> 4) << 16) |
           ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t umma_idesc_f16(int n, bool neg_a) {
    // D = F32 (bits 4..5 = 1), A = B = F16 (0), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
    return (1u << 4) | (neg_a ? (1u << 13) : 0u) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(TC_M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// mbarrier wait with a bound: a protocol error traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(1000u)
            : "memory");
        if (done) return;
        if (spin > (1u << 21)) __trap();
    }
}

// exp(2 pi i frac(u kappa)): the phase fraction is read as a 32-bit fixed-point number from the
// low mantissa word of fma(u, kappa, 1.5 * 2^20) (ulp 2^-32 cycles, |u kappa| < 2^19), turned
// into a float in [-2^22, 2^22) with an exponent trick, scaled to radians for MUFU sin / cos.
__device__ __forceinline__ void antenna_cis(double u, double kappa, float& c, float& s) {
    const double t = __fma_rn(u, kappa, 1572864.0);
    const uint32_t lo = (uint32_t)__double2loint(t);
    const float fb = __uint_as_float((lo >> 9) ^ 0x4B400000u) - 12582912.0f;
    const float ang = fb * 7.4901405e-07f;          // 2 pi / 2^23
    c = __cosf(ang);
    s = __sinf(ang);
}

// hi / lo float16 split of 8 values -> two 16-byte rows
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __half2 hh = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
        const float2 back = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(v[2 * q] - back.x, v[2 * q + 1] - back.y);
        h[q] = *reinterpret_cast<const uint32_t*>(&hh);
        l[q] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// -------------------------------------------------------------------------------------
// forward.  grid = (nitems * nfreq, nunits), block = 416.
//   items    int32 [nitems][4]  {i0, j0, N, 0}: X rows = antennas i0 .. i0 + 127, Y rows =
//                                antennas j0 .. j0 + N - 1 (N a multiple of 32, <= 256)
//   pair_bl  int32 [ldp][ldp]   (baseline << 1 | conj) of V_ij = sum conj(E_i) A E_j, or -1
//   ascale   float [1]          power of two that brings max |A| into [2^14, 2^15)
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_fringe_fwd_kernel(const float* __restrict__ A, const float* __restrict__ ascale,
                     const double* __restrict__ shat, const double* __restrict__ antv,
                     const double* __restrict__ freqs, const int4* __restrict__ units,
                     const int4* __restrict__ items, const int* __restrict__ pair_bl, int ldp,
                     int na, int nbl, int nfreq, int nfp, long long S, double sgn_over_c,
                     float* __restrict__ vpart) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int item = blockIdx.x / nfreq, k = blockIdx.x % nfreq;
    const int4 itm = items[item];
    const int i0 = itm.x, j0 = itm.y, N = itm.z;
    const int4 un = units[blockIdx.y];
    const int nst = (un.z - un.y) / TC_KS;
    if (nst <= 0) return;

    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TcSmem::BAR_OFF);
    uint64_t* empty = full + TC_NSTAGE;
    uint64_t* done = empty + TC_NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + TcSmem::TMEM_OFF);

    // producer warps with at least one real antenna row take part in the pipeline
    int nactive = 0;
#pragma unroll
    for (int w = 0; w < TC_PROD_WARPS; ++w) {
        const int r0 = 32 * w;
        const bool act = (r0 < TC_M) ? (i0 + r0 < na) : (r0 - TC_M < N && j0 + r0 - TC_M < na);
        nactive += act ? 1 : 0;
    }
    // rows nobody writes must still hold finite numbers (their products land in rows / columns
    // of the accumulator that the epilogue skips)
    for (int o = tid * 16; o < TC_NSTAGE * TcSmem::STAGE; o += TC_THREADS * 16)
        *reinterpret_cast<uint4*>(smem + o) = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int st = 0; st < TC_NSTAGE; ++st) {
            mbar_init(&full[st], nactive);
            mbar_init(&empty[st], 1);
        }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == TC_PROD_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "r"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();         // the zero fill above must be visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == TC_PROD_WARPS) {
        // ---------------- MMA issuer
        if (lane == 0) {
            const uint32_t id_pos = umma_idesc_f16(N, false), id_neg = umma_idesc_f16(N, true);
            const uint32_t d_re = tmem, d_im = tmem + TC_IM_COL;
            for (int it = 0; it < nst; ++it) {
                const int stage = it % TC_NSTAGE;
                mbar_wait_bounded(&full[stage], (uint32_t)((it / TC_NSTAGE) & 1));
                tc_fence_after();
                const uint32_t b = smem_u32(smem + stage * TcSmem::STAGE);
                const uint64_t xrh = umma_desc_kmajor(b + TcSmem::XR_H),
                               xrl = umma_desc_kmajor(b + TcSmem::XR_L),
                               xih = umma_desc_kmajor(b + TcSmem::XI_H),
                               xil = umma_desc_kmajor(b + TcSmem::XI_L),
                               yrh = umma_desc_kmajor(b + TcSmem::YR_H),
                               yrl = umma_desc_kmajor(b + TcSmem::YR_L),
                               yih = umma_desc_kmajor(b + TcSmem::YI_H),
                               yil = umma_desc_kmajor(b + TcSmem::YI_L);
                const uint32_t acc = it > 0 ? 1u : 0u;
                // Re V = Er.Yr + Ei.Yi
                umma_f16(d_re, xrh, yrh, id_pos, acc);
                umma_f16(d_re, xrh, yrl, id_pos, 1u);
                umma_f16(d_re, xrl, yrh, id_pos, 1u);
                umma_f16(d_re, xih, yih, id_pos, 1u);
                umma_f16(d_re, xih, yil, id_pos, 1u);
                umma_f16(d_re, xil, yih, id_pos, 1u);
                // Im V = Er.Yi - Ei.Yr
                umma_f16(d_im, xrh, yih, id_pos, acc);
                umma_f16(d_im, xrh, yil, id_pos, 1u);
                umma_f16(d_im, xrl, yih, id_pos, 1u);
                umma_f16(d_im, xih, yrh, id_neg, 1u);
                umma_f16(d_im, xih, yrl, id_neg, 1u);
                umma_f16(d_im, xil, yrh, id_neg, 1u);
                umma_commit(&empty[stage]);       // stage free once these MMAs have read it
            }
            umma_commit(done);
        }
        __syncwarp();
    } else {
        // ---------------- producers: thread <-> operand row
        const bool isY = tid >= TC_M;
        const int rloc = isY ? tid - TC_M : tid;
        const int ant = isY ? j0 + rloc : i0 + rloc;
        const bool wact = isY ? (32 * (rloc >> 5) < N && j0 + 32 * (rloc >> 5) < na)
                              : (i0 + 32 * (rloc >> 5) < na);
        if (wact) {
            const bool live = ant < na && (!isY || rloc < N);
            double px = 0.0, py = 0.0, pz = 0.0;
            if (live) {
                px = antv[4 * (size_t)ant];
                py = antv[4 * (size_t)ant + 1];
                pz = antv[4 * (size_t)ant + 2];
            }
            const double kappa = sgn_over_c * freqs[k];
            const float sc = isY ? __ldg(ascale) : 1.f;
            const float* Ak = A + (size_t)(k / TC_KC) * (size_t)S * TC_KC + (k % TC_KC);
            const int roff = (isY ? TcSmem::YR_H : TcSmem::XR_H) + (rloc >> 3) * 256 + (rloc & 7) * 16;
            const int arr = isY ? TcSmem::Y_ARR : TcSmem::X_ARR;
            for (int it = 0; it < nst; ++it) {
                const int stage = it % TC_NSTAGE;
                if (it >= TC_NSTAGE)
                    mbar_wait_bounded(&empty[stage], (uint32_t)(((it / TC_NSTAGE) - 1) & 1));
                unsigned char* row = smem + stage * TcSmem::STAGE + roff;
                const long long sbase = (long long)un.y + (long long)it * TC_KS;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float c[8], s[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const long long src = sbase + h * 8 + e;
                        const double2 s01 = __ldg(reinterpret_cast<const double2*>(shat + 4 * src));
                        const double s2 = __ldg(shat + 4 * src + 2);
                        const double u = __fma_rn(px, s01.x, __fma_rn(py, s01.y, pz * s2));
                        antenna_cis(u, kappa, c[e], s[e]);
                        if (isY) {
                            const float a = __ldg(Ak + src * TC_KC) * sc;
                            c[e] *= a;
                            s[e] *= a;
                        }
                    }
                    uint4 hi, lo;
                    split8(c, hi, lo);
                    *reinterpret_cast<uint4*>(row + h * 128) = hi;
                    *reinterpret_cast<uint4*>(row + arr + h * 128) = lo;
                    split8(s, hi, lo);
                    *reinterpret_cast<uint4*>(row + 2 * arr + h * 128) = hi;
                    *reinterpret_cast<uint4*>(row + 3 * arr + h * 128) = lo;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[stage]);
            }
        }
        // ---------------- epilogue: TMEM -> registers -> Vpart through the pair table
        mbar_wait_bounded(done, 0u);
        tc_fence_after();
        const int q = warp & 3, cg = warp >> 2;
        const int ai = i0 + 32 * q + lane;
        const float inv = 1.f / __ldg(ascale);
        float2* vp = reinterpret_cast<float2*>(vpart) + (size_t)blockIdx.y * (size_t)nbl * nfp + k;
        for (int cb = cg; cb < N / 32; cb += TC_PROD_WARPS / 4) {
            float re[32], im[32];
            const uint32_t ta = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(cb * 32);
            tmem_ld32(ta, re);
            tmem_ld32(ta + TC_IM_COL, im);
            if (ai < na) {
                const int* pb = pair_bl + (size_t)ai * ldp + j0 + cb * 32;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const int4 e4 = __ldg(reinterpret_cast<const int4*>(pb) + j4);
                    const int e[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        if (e[jj] < 0) continue;
                        const int j = 4 * j4 + jj;
                        vp[(size_t)(e[jj] >> 1) * nfp] =
                            make_float2(re[j] * inv, (e[jj] & 1) ? -im[j] * inv : im[j] * inv);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_PROD_WARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                     "r"(TC_TMEM_COLS)
                     : "memory");
    }
}


int launch_tc_fwd(const float* A, const float* ascale, const double* shat, const double* antv,
                  const double* freqs, const int* units, int nunits, const int* items, int nitems,
                  const int* pair_bl, int ldp, int na, int nbl, int nfreq, long long S, int conj,
                  float* vpart, cudaStream_t st) {
    if (nunits <= 0 || nitems <= 0 || nfreq <= 0) return 0;
    if (S % SRC_PAD) return set_error("tcfringe_fwd: S must be a multiple of 128");
    if (ldp % 32 || ldp < na) return set_error("tcfringe_fwd: pair table pitch must be na padded to 32");
    const int nfp = ((nfreq + TC_KC - 1) / TC_KC) * TC_KC;
    const long long gx = (long long)nitems * nfreq;
    if (gx > 2147483647LL || nunits > 65535) return set_error("tcfringe_fwd: grid too large");
    static DeviceOnce attr_once;
    if (attr_once.first() &&
        cudaFuncSetAttribute(tc_fringe_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             TcSmem::TOTAL) != cudaSuccess)
        return set_error("tcfringe_fwd: cannot reserve shared memory");
    dim3 grid((unsigned)gx, nunits);
    tc_fringe_fwd_kernel<<<grid, TC_THREADS, TcSmem::TOTAL, st>>>(
        A, ascale, shat, antv, freqs, reinterpret_cast<const int4*>(units),
        reinterpret_cast<const int4*>(items), pair_bl, ldp, na, nbl, nfreq, nfp, S,
        (conj ? -1.0 : 1.0) / C_LIGHT, vpart);
    return check_launch("tcfringe_fwd");
}

}  // namespace b200rime

extern "C" {

int b200rime_tcfringe_fwd_f32(const float* A, const float* ascale, const double* shat,
                              const double* antv, const double* freqs, const int* units, int nunits,
                              const int* items, int nitems, const int* pair_bl, int ldp, int na,
                              int nbl, int nfreq, long long S, int conj, float* Vpart, void* stream) {
    return b200rime::launch_tc_fwd(A, ascale, shat, antv, freqs, units, nunits, items, nitems,
                                   pair_bl, ldp, na, nbl, nfreq, S, conj, Vpart,
                                   (cudaStream_t)stream);
}
int b200rime_tc_rows(void) { return b200rime::TC_M; }
int b200rime_tc_cols_max(void) { return b200rime::TC_NMAX; }

}  // extern "C"
