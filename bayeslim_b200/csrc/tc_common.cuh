// tcgen05 / TMEM building blocks shared by the tensor-core translation units of libb200rime.so
// (tc_kernels.cu: fringe sums; alm_kernels.cu: spherical-harmonic complex GEMM): shared-memory
// stage layout, UMMA descriptors, MMA issue of one split-float16 complex stage, accumulator
// read-out of one TMEM chain.  See the header of tc_kernels.cu for the numerical design.
#pragma once
#include <cuda_fp16.h>
#include "rime_math.cuh"
#include "internal.h"

namespace b200rime {

#ifndef B200_TC_PROBE
#define B200_TC_PROBE 0            // timing probes (wrong results): 1 no sine / cosine, 2 no MMAs, 4 no generation
#endif
#ifndef B200_TC_UNROLL_FWD
#define B200_TC_UNROLL_FWD 1       // unroll the two 8-source halves of a stage in the generating loop
#endif
#ifndef B200_TC_UNROLL_BWD
#define B200_TC_UNROLL_BWD 0
#endif
#ifndef B200_TC_FLUSH
#define B200_TC_FLUSH 4            // stages (of 16 sources) per TMEM accumulation chain
#endif
constexpr int TC_M = 128;            // X rows (first antennas) per item = UMMA M
constexpr int TC_NMAX = 128;         // Y rows (second antennas) per item, at most (UMMA N = 2 x this)
constexpr int TC_KS = 16;            // sources per stage = one kind::f16 UMMA K step
constexpr int TC_NSTAGE = 5;         // operand stages (40 KB each)
constexpr int TC_NSRC = 8;           // source-data slots (staged TC_NSRC stages ahead)
constexpr int TC_FLUSH = B200_TC_FLUSH;
constexpr int TC_WARPS = 16;         // every warp holds a (lane quarter x 32 columns) accumulator slice
constexpr int TC_THREADS = TC_WARPS * 32;         // 512 threads x 128 registers = the register file
constexpr int TC_CTRL_WARP = 0;      // stages the sources (TMA) and issues the MMAs
constexpr int TC_GROUPS = 3;         // warps 4..15: group g generates stages it = g (mod 3)
constexpr int TC_KC = B200_KC_F32;
constexpr int TC_TMEM_COLS = 512;    // two accumulator sets of (re | im) x 128 columns
constexpr int TC_SET_COLS = 256;
constexpr int TC_CG = 32;            // accumulator columns per warp (4 column groups)

struct TcSmem {
    // one operand array = 128 rows x 16 float16 in the canonical no-swizzle K-major layout:
    //   byte offset(row, kgroup of 8) = (row / 8) * 256 + kgroup * 128 + (row % 8) * 16
    // i.e. 8 x 16-byte core matrices, LBO (K direction) = 128 B, SBO (row direction) = 256 B
    static constexpr int ARR = TC_M * TC_KS * 2;              // 4 KB
    static constexpr int XR_H = 0, XR_L = ARR, XI_H = 2 * ARR, XI_L = 3 * ARR;
    // B operands are 2 N rows tall, (re ; im) stacked.  P = (Yr ; Yi) pairs with Er, M = (Yi ; -Yr)
    // with Ei: both are windows of ONE buffer of three halves (Yr ; Yi ; -Yr), M starting N rows =
    // (N / 8) * 256 bytes after P (the backward kernel stores (-Hi ; Hr ; Hi): M first, then P).
    static constexpr int BBUF = 3 * ARR;                      // 12 KB per three-half buffer
    static constexpr int B_H = 4 * ARR, B_L = B_H + BBUF;
    static constexpr int STAGE = 4 * ARR + 2 * BBUF;          // 40 KB
    // source slot: 16 x (x, y, z, 0) float64 unit vectors, then 16 float32 sky values
    static constexpr int SRC_SHAT = TC_KS * 32, SRC_A = TC_KS * 4, SRC_SLOT = SRC_SHAT + SRC_A;
    static constexpr int SRC_OFF = TC_NSTAGE * STAGE;
    // full[NSTAGE], empty[NSTAGE], sfull[NSRC], tfull[2], tempty[2]
    static constexpr int BAR_OFF = SRC_OFF + TC_NSRC * SRC_SLOT;
    static constexpr int TMEM_OFF = BAR_OFF + (2 * TC_NSTAGE + TC_NSRC + 4) * 8;
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
__host__ __device__ constexpr uint32_t umma_idesc_f16(int n) {
    // D = F32 (bits 4..5 = 1), A = B = F16 (0), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                         uint32_t accumulate) {
#if B200_TC_PROBE & 2
    return;
#endif
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
// 8 consecutive columns of the warp's 32 TMEM lanes -> 8 registers per thread (issue only)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}"
        : "=r"(pred));
    return pred != 0;
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

// exp(2 pi i frac(p)) for t = p + 1.5 * 2^20, p the phase in cycles (|p| < 2^19) summed onto the
// offset with FMAs: the offset makes the ulp of t 2^-32 cycles, so the low mantissa word of t is
// the phase fraction as a signed 32-bit fixed-point number in [-1/2, 1/2): one integer
// conversion, one scaling to radians, MUFU sine / cosine.
__device__ __forceinline__ void antenna_cis(double t, float& c, float& s) {
#if B200_TC_PROBE & 1
    c = __uint_as_float(__double2loint(t)) * 1e-30f + 0.5f;
    s = 0.25f;
    return;
#endif
    const float ang = __int2float_rn(__double2loint(t)) * 1.4629180792671596e-09f;   // 2 pi / 2^32
    c = __cosf(ang);
    s = __sinf(ang);
}

// hi / lo float16 split of 8 values -> two 16-byte core-matrix rows.  lo = v - float(hi) is one
// mixed-precision FMA (FHFMA: float16 x float16 + float32), exact before its final rounding.
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
    const unsigned short m1 = 0xBC00;               // -1.0 in float16
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __half2 hh = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
        const uint32_t hu = *reinterpret_cast<const uint32_t*>(&hh);
        float l0, l1;
        asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(l0) : "h"((unsigned short)(hu & 0xffffu)), "h"(m1), "f"(v[2 * q]));
        asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(l1) : "h"((unsigned short)(hu >> 16)), "h"(m1), "f"(v[2 * q + 1]));
        const __half2 ll = __floats2half2_rn(l0, l1);
        h[q] = hu;
        l[q] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ uint4 neg_half8(const uint4& v) {      // sign flip of 8 float16
    return make_uint4(v.x ^ 0x80008000u, v.y ^ 0x80008000u, v.z ^ 0x80008000u, v.w ^ 0x80008000u);
}
// (cos, sin) of 8 sources -> the four arrays (re_hi, re_lo, im_hi, im_lo) of an A operand row
__device__ __forceinline__ void store_split8(unsigned char* dst, const float (&c)[8], const float (&s)[8]) {
    uint4 hi, lo;
    split8(c, hi, lo);
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + TcSmem::ARR) = lo;
    split8(s, hi, lo);
    *reinterpret_cast<uint4*>(dst + 2 * TcSmem::ARR) = hi;
    *reinterpret_cast<uint4*>(dst + 3 * TcSmem::ARR) = lo;
}
// phase of 8 sources (float64 unit vectors at sh, 32 bytes apart) for the antenna position p
__device__ __forceinline__ void cis8(const double (&p)[3], const unsigned char* sh, float (&c)[8],
                                     float (&s)[8]) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const double2 xy = *reinterpret_cast<const double2*>(sh + 32 * e);
        const double z = *reinterpret_cast<const double*>(sh + 32 * e + 16);
        antenna_cis(__fma_rn(p[0], xy.x, __fma_rn(p[1], xy.y, __fma_rn(p[2], z, 1572864.0))), c[e], s[e]);
    }
}

struct TcIssue {                 // state of the issuing warp (warp-uniform values)
    uint32_t tmem, idesc, smem_base;
    uint32_t poff, moff;         // byte offsets of the P and M windows inside a B buffer
};

// the six MMAs of one stage: (Re | Im) += A_re (P) + A_im (M), three float16 split terms each
__device__ __forceinline__ void tc_issue_stage(const TcIssue& q, int stage, int set, bool first) {
    const uint32_t d = q.tmem + set * TC_SET_COLS;            // Re columns [0, N), Im [N, 2 N)
    const uint32_t b = q.smem_base + stage * TcSmem::STAGE;
    const uint64_t arh = umma_desc_kmajor(b + TcSmem::XR_H), arl = umma_desc_kmajor(b + TcSmem::XR_L),
                   aih = umma_desc_kmajor(b + TcSmem::XI_H), ail = umma_desc_kmajor(b + TcSmem::XI_L),
                   bph = umma_desc_kmajor(b + TcSmem::B_H + q.poff),
                   bpl = umma_desc_kmajor(b + TcSmem::B_L + q.poff),
                   bmh = umma_desc_kmajor(b + TcSmem::B_H + q.moff),
                   bml = umma_desc_kmajor(b + TcSmem::B_L + q.moff);
    umma_f16(d, arh, bph, q.idesc, first ? 0u : 1u);
    umma_f16(d, arh, bpl, q.idesc, 1u);
    umma_f16(d, arl, bph, q.idesc, 1u);
    umma_f16(d, aih, bmh, q.idesc, 1u);
    umma_f16(d, aih, bml, q.idesc, 1u);
    umma_f16(d, ail, bmh, q.idesc, 1u);
}

// one warp adds accumulator set rc & 1 (its 32 TMEM lanes x its 32 columns, re at column 0 and im
// at column im_col of the set) to its register accumulators and hands the set back
__device__ __forceinline__ void tc_read_chain(uint64_t* tfull, uint64_t* tempty, int rc,
                                              uint32_t ta0, uint32_t im_col, float (&aR)[TC_CG],
                                              float (&aI)[TC_CG], int lane) {
    const int set = rc & 1;
    mbar_wait_bounded(&tfull[set], (uint32_t)((rc >> 1) & 1));
    tc_fence_after();
    const uint32_t ta = ta0 + (uint32_t)(set * TC_SET_COLS);
#pragma unroll
    for (int g = 0; g < TC_CG / 8; ++g) {
        uint32_t vr[8], vi[8];
        tmem_ld8(ta + 8 * g, vr);
        tmem_ld8(ta + im_col + 8 * g, vi);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            aR[8 * g + c] += __uint_as_float(vr[c]);
            aI[8 * g + c] += __uint_as_float(vi[c]);
        }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tempty[set]);
}

}  // namespace b200rime
