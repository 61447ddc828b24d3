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
//     of two into float16 range, so hi + lo carries 22 mantissa bits; lo = v - hi is one
//     mixed-precision FHFMA) and a real product costs three MMAs (hi.hi + hi.lo + lo.hi; the
//     dropped lo.lo term is 2^-24 relative): float32-grade results at a third of the float16
//     tensor rate;
//   * the B operand stacks the real and the imaginary part: (Re | Im) += Er (Yr | Yi) +
//     Ei (Yi | -Yr) is six MMAs of width 2 N per 16 sources instead of twelve of width N.
//     Measured on B200 (scripts/gpu_tc_variants.sh, operands resident, no generation): a
//     128 x 128 x 16 MMA takes ~116 cycles (55 % of the tensor pipe), 128 x 256 x 16 ~166 (77 %):
//     there is a fixed cost of ~65 cycles per instruction.
//
// Accumulation.  The tensor core adds into its FP32 accumulator with truncation (measured on
// B200: a chain of n MMAs loses ~0.75 * 2^-24 * n of a coherent sum -- 6e-5 after 8192 sources,
// profiles/r02_tc_v1_truncation_study.jsonl), so a TMEM chain is kept short (TC_FLUSH stages =
// 64 sources by default) and then added, with ordinary round-to-nearest FADDs, to float32
// accumulators held in registers; two TMEM accumulator sets alternate so that the MMAs of chain
// c + 1 overlap the read-out of chain c.
//
// CTA = 16 warps x 128 registers.  Every warp holds a slice of the register accumulators (TMEM
// lane quarter warp % 4 x 32 columns warp / 4, real and imaginary) and reads its slice of every
// chain.  Beyond that the roles are:
//   warp 0        control: stages the source data (unit vectors, channel row of the perceived
//                 sky) by 1-D TMA bulk copies and issues the MMAs.  Its code is warp-uniform, so
//                 descriptors live in uniform registers and a stage costs ~50 instructions.
//   warps 1..3    read-out only (the backward kernel also gives them the fixed-order sum over
//                 column groups).
//   warps 4..15   three generating groups of four warps; group g generates the whole stages
//                 it = g (mod 3): thread <-> operand row (32 (warp % 4) + lane) in both roles, all
//                 16 sources.  With one stage per group in flight, a stage may take three stage
//                 periods: the latency of the per-stage handshakes (mbarrier waits, proxy fence)
//                 is hidden, which a design where every warp touches every stage cannot do
//                 (measured: 2380 -> 1850 cycles per stage before the stacked operands).
// full / empty mbarriers per operand stage, sfull per source slot, tfull / tempty per TMEM set;
// tcgen05.commit releases stages and hands chains over; finally the registers are scattered
// through the antenna-pair table into Vpart[unit][baseline][channel].
//
// Replaces telescope_model.py:310-358 (gen_fringe) + rime_model.py:426-429 (multiply, sum).
#include "tc_common.cuh"

namespace b200rime {

// -------------------------------------------------------------------------------------
// forward.  grid = (nitems * nfreq, nunits), block = 512.
//   Acm      float [Nfp][S]     perceived sky, channel-major (row k = channel k over the packed
//                                source axis): the 16 values of a stage are one 64-byte bulk copy
//   items    int32 [nitems][4]  {i0, j0, N, 0}: X rows = antennas i0 .. i0 + 127, Y rows =
//                                antennas j0 .. j0 + N - 1 (N a multiple of 32, <= 128)
//   pair_bl  int32 [ldp][ldp]   (baseline << 1 | conj) of V_ij = sum conj(E_i) A E_j, or -1
//   ascale   float [1]          power of two that brings max |A| into [2^14, 2^15)
// A operand X = E_i (re, im; hi, lo), B windows P = (Yr ; Yi), M = (Yi ; -Yr) of the buffer
// (Yr ; Yi ; -Yr) with Y = A E_j:
//   Re V = Er.Yr + Ei.Yi,   Im V = Er.Yi - Ei.Yr.
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_fringe_fwd_kernel(const float* __restrict__ Acm, const float* __restrict__ ascale,
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
    const int nchain = (nst + TC_FLUSH - 1) / TC_FLUSH;

    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TcSmem::BAR_OFF);
    uint64_t* empty = full + TC_NSTAGE;
    uint64_t* sfull = empty + TC_NSTAGE;
    uint64_t* tfull = sfull + TC_NSRC;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + TcSmem::TMEM_OFF);

    // rows nobody writes (antennas beyond the array) must still hold finite numbers: their
    // products land in rows / columns of the accumulator that are never stored
    for (int o = tid * 16; o < TC_NSTAGE * TcSmem::STAGE; o += TC_THREADS * 16)
        *reinterpret_cast<uint4*>(smem + o) = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int st = 0; st < TC_NSTAGE; ++st) {
            mbar_init(&full[st], 4);                  // the four warps of the generating group
            mbar_init(&empty[st], 1);
        }
        for (int st = 0; st < TC_NSRC; ++st) mbar_init(&sfull[st], 1);
        for (int q = 0; q < 2; ++q) {
            mbar_init(&tfull[q], 1);
            mbar_init(&tempty[q], TC_WARPS);
        }
        mbar_fence_init();
    }
    if (warp == TC_CTRL_WARP) {
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
    const float* Ak = Acm + (size_t)k * (size_t)S;

    // q = TMEM lane quarter (and operand rows 32 q ..), cg = accumulator column group
    const int q = warp & 3, cg = warp >> 2;
    const int role = __shfl_sync(0xffffffffu, warp, 0);              // warp-uniform warp index
    const uint32_t ta0 = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(TC_CG * cg);
    const uint32_t im_col = (uint32_t)N;
    const float sc = __ldg(ascale);
    float aR[TC_CG], aI[TC_CG];
#pragma unroll
    for (int c = 0; c < TC_CG; ++c) aR[c] = aI[c] = 0.f;
    int next_read = 0;

    if (role == TC_CTRL_WARP) {
        // ---------------- control warp: source staging by TMA, MMA issue (uniform datapath)
        TcIssue iq;
        iq.tmem = tmem;
        iq.idesc = umma_idesc_f16(2 * N);
        iq.smem_base = smem_u32(smem);
        iq.poff = 0;
        iq.moff = (uint32_t)((N >> 3) * 256);
        auto stage_sources = [&](int st) {
            const int slot = st % TC_NSRC;
            unsigned char* dst = smem + TcSmem::SRC_OFF + slot * TcSmem::SRC_SLOT;
            const long long s0 = (long long)un.y + (long long)st * TC_KS;
            mbar_expect_tx(&sfull[slot], TcSmem::SRC_SLOT);
            bulk_g2s(dst, shat + 4 * s0, TcSmem::SRC_SHAT, &sfull[slot]);
            bulk_g2s(dst + TcSmem::SRC_SHAT, Ak + s0, TcSmem::SRC_A, &sfull[slot]);
        };
        if (elect_one())
            for (int st = 0; st < min(nst, TC_NSRC); ++st) stage_sources(st);
        __syncwarp();
        for (int it = 0; it < nst; ++it) {
            const int stage = it % TC_NSTAGE, chain = it / TC_FLUSH, set = chain & 1;
            const bool first = (it % TC_FLUSH) == 0;
            if (first) {
                // this warp's own share of the read-out, then: the set's previous chain has
                // been read out by everybody
                while (next_read <= chain - 2)
                    tc_read_chain(tfull, tempty, next_read++, ta0, im_col, aR, aI, lane);
                if (chain >= 2)
                    mbar_wait_bounded(&tempty[set], (uint32_t)(((chain >> 1) - 1) & 1));
            }
            mbar_wait_bounded(&full[stage], (uint32_t)((it / TC_NSTAGE) & 1));
            tc_fence_after();
            if (elect_one()) {
                tc_issue_stage(iq, stage, set, first);
                umma_commit(&empty[stage]);       // stage free once these MMAs have read it
                if ((it % TC_FLUSH) == TC_FLUSH - 1 || it == nst - 1) umma_commit(&tfull[set]);
                // the owning group has consumed the source slot of this stage: refill it
                if (it + TC_NSRC < nst) stage_sources(it + TC_NSRC);
            }
            __syncwarp();
        }
    } else if (cg >= 1) {
        // ---------------- generating groups: thread <-> operand row 32 q + lane in both roles,
        // all 16 sources of the stages it = cg - 1 (mod 3)
        const int row = 32 * q + lane;
        const int nx = min(TC_M, na - i0), ny = min(N, na - j0);      // live X / Y rows
        const bool xlive = row < nx, ylive = row < ny, diag = (i0 == j0);
        const double kappa = sgn_over_c * freqs[k];
        // antenna positions in cycles per unit direction cosine: phase = r' . shat
        double xp[3] = {0.0, 0.0, 0.0}, yp[3] = {0.0, 0.0, 0.0};
        if (xlive) {
            const double* a = antv + 4 * (size_t)(i0 + row);
            xp[0] = kappa * a[0], xp[1] = kappa * a[1], xp[2] = kappa * a[2];
        }
        if (ylive) {
            const double* a = antv + 4 * (size_t)(j0 + row);
            yp[0] = kappa * a[0], yp[1] = kappa * a[1], yp[2] = kappa * a[2];
        }
        const int roff = (row >> 3) * 256 + (row & 7) * 16;
        const int half2 = (N >> 3) * 256;              // second half of a stacked operand
        for (int it = cg - 1; it < nst; it += TC_GROUPS) {
            while (next_read <= it / TC_FLUSH - 2)
                tc_read_chain(tfull, tempty, next_read++, ta0, im_col, aR, aI, lane);
            const int stage = it % TC_NSTAGE, slot = it % TC_NSRC;
            mbar_wait_bounded(&sfull[slot], (uint32_t)((it / TC_NSRC) & 1));
            if (it >= TC_NSTAGE)
                mbar_wait_bounded(&empty[stage], (uint32_t)(((it / TC_NSTAGE) - 1) & 1));
            const unsigned char* src = smem + TcSmem::SRC_OFF + slot * TcSmem::SRC_SLOT;
            unsigned char* dst = smem + stage * TcSmem::STAGE + roff;
#if B200_TC_UNROLL_FWD      // both halves of a stage in one body: -2.3 % forward (measured); the backward loses 3 %
#pragma unroll
#else
#pragma unroll 1
#endif
            for (int kg = 0; kg < ((B200_TC_PROBE & 4) ? 0 : 2); ++kg) {
                float c[8], s[8];
                const unsigned char* sh = src + kg * 8 * 32;
                if (xlive || (diag && ylive)) cis8(xp, sh, c, s);
                if (xlive) store_split8(dst + kg * 128 + TcSmem::XR_H, c, s);
                if (ylive) {
                    if (!diag) cis8(yp, sh, c, s);
                    const float4 a0 = reinterpret_cast<const float4*>(src + TcSmem::SRC_SHAT)[2 * kg];
                    const float4 a1 = reinterpret_cast<const float4*>(src + TcSmem::SRC_SHAT)[2 * kg + 1];
                    const float a[8] = {a0.x * sc, a0.y * sc, a0.z * sc, a0.w * sc,
                                        a1.x * sc, a1.y * sc, a1.z * sc, a1.w * sc};
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        c[e] *= a[e];
                        s[e] *= a[e];
                    }
                    // (Yr ; Yi ; -Yr), hi and lo parts
                    unsigned char* d = dst + kg * 128;
                    uint4 rh, rl, ih, il;
                    split8(c, rh, rl);
                    split8(s, ih, il);
                    *reinterpret_cast<uint4*>(d + TcSmem::B_H) = rh;
                    *reinterpret_cast<uint4*>(d + TcSmem::B_H + half2) = ih;
                    *reinterpret_cast<uint4*>(d + TcSmem::B_H + 2 * half2) = neg_half8(rh);
                    *reinterpret_cast<uint4*>(d + TcSmem::B_L) = rl;
                    *reinterpret_cast<uint4*>(d + TcSmem::B_L + half2) = il;
                    *reinterpret_cast<uint4*>(d + TcSmem::B_L + 2 * half2) = neg_half8(rl);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);
        }
    }
    while (next_read < nchain) tc_read_chain(tfull, tempty, next_read++, ta0, im_col, aR, aI, lane);

    // ---- scatter through the pair table
    {
        const int ai = i0 + 32 * q + lane;
        if (ai < na && TC_CG * cg < N) {
            const float inv = 1.f / sc;
            float2* vp = reinterpret_cast<float2*>(vpart) + (size_t)blockIdx.y * (size_t)nbl * nfp + k;
            const int* pb = pair_bl + (size_t)ai * ldp + j0 + TC_CG * cg;
#pragma unroll
            for (int c4 = 0; c4 < TC_CG / 4; ++c4) {
                const int4 e4 = __ldg(reinterpret_cast<const int4*>(pb) + c4);
                const int e[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    if (e[cc] < 0) continue;
                    const int c = 4 * c4 + cc;
                    vp[(size_t)(e[cc] >> 1) * nfp] =
                        make_float2(aR[c] * inv, (e[cc] & 1) ? -aI[c] * inv : aI[c] * inv);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_CTRL_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                     "r"(TC_TMEM_COLS)
                     : "memory");
    }
}

// -------------------------------------------------------------------------------------
// backward.  grid = (nunits, nitem * nfreq), block = 512 (units fastest: CTAs that run together
// share the cotangent operand of their (antenna block, channel) through L2).
//
// With the Hermitian cotangent matrix H[a][m] (ant_kernels.cu) the adjoints are
//   y_a[s] = sum_m H[a][m] E_m[s],  p = conj(E_a[s]) y_a[s],
//   dL/dA[s] = 1/2 sum_a Re p,      dL/dr_a = sum_s shat_s A_s (2 pi sgn nu / c) Im p.
// y is a GEMM with M = sources (tiles of 128), N = antennas a (items of 128), K = partner antennas
// m: the A operand E_m[s] is generated in shared memory (generating thread <-> source row, all 16
// antennas of a stage), the B operands come by TMA from a copy the host has scaled, split into
// float16 hi / lo parts, stacked (re ; im) and laid out in the UMMA canonical order
//   Hq[t][k][item][stage of 16 m][hi | lo][(-Hi ; Hr ; Hi): 384 rows x 16 m, canonical],
//   M = (-Hi ; Hr) (rows 0..255) pairs with Ei, P = (Hr ; Hi) (rows 128..383) with Er:
//   Re y = Er.Hr - Ei.Hi, Im y = Er.Hi + Ei.Hr
// Chains of TC_FLUSH stages are added to register accumulators as in the forward kernel; after
// the last chain of a source tile every warp regenerates E_a for its source and its 32 antennas,
// forms p and reduces: dL/dA over its columns, then across the four column groups in a fixed
// order through shared memory (summed by the read-out warps 1..3; one float per source and
// channel, written channel-major); dL/dr over the 32 sources of the warp with a transposed
// shuffle reduction (lane <-> antenna), accumulated over the tiles of the unit.
//   mrange [nitem][2]: stages of 16 partner antennas [lo, hi) that hold cotangent entries for the
//   item.  When only dL/dA is wanted H is the doubled lower triangle (a > m) and item ib stops
//   after its own antennas; a baseline group that covers only some antenna blocks skips the
//   stages it leaves empty.
//   dAcm   [nitem][Nfp][S]                     partial dL/dA, channel-major (sum over axis 0)
//   drpart [nunits][Nfp][4][nitem * 128][4]    partial dL/dr (float32; sum over the first three)
// -------------------------------------------------------------------------------------
struct TcBwdSmem {
    static constexpr int POS_MAX = 512;                        // antennas the position table holds
    static constexpr int POS_OFF = TC_NSTAGE * TcSmem::STAGE;  // [POS_MAX][4] float64, kappa-scaled
    static constexpr int RED_OFF = POS_OFF + POS_MAX * 32;     // [2][4 column groups][128] float
    // full[NSTAGE], empty[NSTAGE], tfull[2], tempty[2], rbar[2], rdone[2]
    static constexpr int BAR_OFF = RED_OFF + 2 * 4 * TC_M * 4;
    static constexpr int TMEM_OFF = BAR_OFF + (2 * TC_NSTAGE + 8) * 8;
    static constexpr int TOTAL = TMEM_OFF + 16;
    static constexpr int H_BYTES = 2 * TcSmem::BBUF;           // one stage of the cotangent operands (24 KB)
};

// v[i] summed over the 32 lanes, result for index i = lane left in v[0] (31 shuffles)
__device__ __forceinline__ float warp_reduce_scatter32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? v[i] : v[i + off];
            const float keep = up ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_fringe_bwd_kernel(const unsigned char* __restrict__ Hq, const float* __restrict__ hscale,
                     const float* __restrict__ Acm, const double* __restrict__ shat,
                     const double* __restrict__ antv, const double* __restrict__ freqs,
                     const int4* __restrict__ units, int nitem, int na, int nm_pad,
                     const int2* __restrict__ mrange, int nfreq, int nfp, long long S,
                     double sgn_over_c, int need_a, int need_r, float* __restrict__ dAcm,
                     float* __restrict__ drpart) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ai = blockIdx.y / nfreq, k = blockIdx.y % nfreq;
    const int4 un = units[blockIdx.x];
    const int a0 = ai * TC_M;
    // all 128 columns are computed: the operand rows of antennas >= na are zero, so are their
    // sums, and the epilogue needs no column mask
    const int nmst_all = nm_pad / TC_KS;
    const int2 mr = mrange[ai];
    const int mlo = max(0, mr.x), nmst = min(nmst_all, mr.y) - mlo;
    const int ntile = (un.z - un.y + TC_M - 1) / TC_M;
    if (ntile <= 0 || nmst <= 0) return;        // nothing to add: dAcm / drpart were zeroed
    const int nct = (nmst + TC_FLUSH - 1) / TC_FLUSH;                 // chains per source tile
    const int nchain = ntile * nct;
    const int nst = ntile * nmst;                                     // stages of this CTA

    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TcBwdSmem::BAR_OFF);
    uint64_t* empty = full + TC_NSTAGE;
    uint64_t* tfull = empty + TC_NSTAGE;
    uint64_t* tempty = tfull + 2;
    uint64_t* rbar = tempty + 2;             // column-group partials of a source tile are in `red`
    uint64_t* rdone = rbar + 2;              // ... and have been summed: the buffer is free again
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + TcBwdSmem::TMEM_OFF);
    double4* pos = reinterpret_cast<double4*>(smem + TcBwdSmem::POS_OFF);
    float* red = reinterpret_cast<float*>(smem + TcBwdSmem::RED_OFF);
    const double kappa = sgn_over_c * freqs[k];

    for (int o = tid * 16; o < TC_NSTAGE * TcSmem::STAGE; o += TC_THREADS * 16)
        *reinterpret_cast<uint4*>(smem + o) = make_uint4(0, 0, 0, 0);
    for (int a = tid; a < TcBwdSmem::POS_MAX; a += TC_THREADS) {
        double4 p = make_double4(0.0, 0.0, 0.0, 0.0);
        if (a < na) {
            p.x = kappa * antv[4 * (size_t)a];
            p.y = kappa * antv[4 * (size_t)a + 1];
            p.z = kappa * antv[4 * (size_t)a + 2];
        }
        pos[a] = p;
    }
    if (tid == 0) {
        for (int st = 0; st < TC_NSTAGE; ++st) {
            mbar_init(&full[st], 4 + 1);         // the generating group's warps + the TMA's expect_tx
            mbar_init(&empty[st], 1);
        }
        for (int q = 0; q < 2; ++q) {
            mbar_init(&tfull[q], 1);
            mbar_init(&tempty[q], TC_WARPS);
            mbar_init(&rbar[q], TC_WARPS);
            mbar_init(&rdone[q], 3);
        }
        mbar_fence_init();
    }
    if (warp == TC_CTRL_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "r"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // warp roles as in the forward kernel: q = TMEM lane quarter = source rows 32 q .. of the tile,
    // cg = column group = antennas a0 + 32 cg ..
    const int q = warp & 3, cg = warp >> 2;
    const int role = __shfl_sync(0xffffffffu, warp, 0);
    const uint32_t ta0 = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(TC_CG * cg);
    const float inv_h = 1.f / __ldg(hscale);
    const float kf = (float)(kappa * 6.283185307179586476925286766559);
    float aR[TC_CG], aI[TC_CG];
#pragma unroll
    for (int c = 0; c < TC_CG; ++c) aR[c] = aI[c] = 0.f;
    float gr[3] = {0.f, 0.f, 0.f};
    int next_read = 0;

    // chain rc -> registers; after the last chain of a source tile: the tile's epilogue
    auto read_chain = [&](int rc) {
        tc_read_chain(tfull, tempty, rc, ta0, (uint32_t)TC_NMAX, aR, aI, lane);
        if (rc % nct != nct - 1) return;
        const int tile = rc / nct;
        const long long s = (long long)un.y + (long long)tile * TC_M + 32 * q + lane;
        const bool valid = s < un.z;
        double sv[3] = {0.0, 0.0, 0.0};
        float wk = 0.f;
        if (valid) {
            const double2 s01 = __ldg(reinterpret_cast<const double2*>(shat + 4 * s));
            sv[0] = s01.x, sv[1] = s01.y, sv[2] = __ldg(shat + 4 * s + 2);
            if (need_r) wk = __ldg(Acm + (size_t)k * (size_t)S + s) * kf * inv_h;
        }
        float dacc = 0.f;
        const double4* pa = pos + a0 + TC_CG * cg;
#pragma unroll
        for (int c = 0; c < TC_CG; ++c) {
            const double4 p = pa[c];
            float cs, sn;
            antenna_cis(__fma_rn(p.x, sv[0], __fma_rn(p.y, sv[1], __fma_rn(p.z, sv[2], 1572864.0))),
                        cs, sn);
            const float yr = aR[c], yi = aI[c];
            dacc = fmaf(cs, yr, fmaf(sn, yi, dacc));            // Re(conj(E) y)
            aR[c] = wk * fmaf(cs, yi, -sn * yr);                // A kappa 2 pi Im(conj(E) y)
        }
        if (need_r) {
            const float fx = (float)sv[0], fy = (float)sv[1], fz = (float)sv[2];
#pragma unroll
            for (int c = 0; c < TC_CG; ++c) aI[c] = aR[c] * fx;
            gr[0] += warp_reduce_scatter32(aI, lane);
#pragma unroll
            for (int c = 0; c < TC_CG; ++c) aI[c] = aR[c] * fy;
            gr[1] += warp_reduce_scatter32(aI, lane);
#pragma unroll
            for (int c = 0; c < TC_CG; ++c) aI[c] = aR[c] * fz;
            gr[2] += warp_reduce_scatter32(aI, lane);
        }
#pragma unroll
        for (int c = 0; c < TC_CG; ++c) aR[c] = aI[c] = 0.f;
        if (need_a) {
            // sum over the four column groups in a fixed order (deterministic): partials through
            // shared memory, summed by the read-out warps 1..3 (warp 1 also covers quarter 0)
            const int buf = tile & 1;
            float* rb = red + buf * 4 * TC_M;
            if (tile >= 2) mbar_wait_bounded(&rdone[buf], (uint32_t)(((tile >> 1) - 1) & 1));
            rb[cg * TC_M + 32 * q + lane] = dacc;
            __syncwarp();
            if (lane == 0) mbar_arrive(&rbar[buf]);
            if (role >= 1 && role <= 3) {
                mbar_wait_bounded(&rbar[buf], (uint32_t)((tile >> 1) & 1));
                for (int qq = (role == 1 ? 0 : role); qq <= role; ++qq) {
                    const long long s2 = (long long)un.y + (long long)tile * TC_M + 32 * qq + lane;
                    const float* r4 = rb + 32 * qq + lane;
                    const float tot = ((r4[0] + r4[TC_M]) + r4[2 * TC_M]) + r4[3 * TC_M];
                    if (s2 < un.z) dAcm[((size_t)ai * nfp + k) * (size_t)S + s2] = 0.5f * tot * inv_h;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&rdone[buf]);
            }
        }
    };

    if (role == TC_CTRL_WARP) {
        // ---------------- control warp: MMA issue (uniform datapath)
        TcIssue iq;
        iq.tmem = tmem;
        iq.idesc = umma_idesc_f16(2 * TC_NMAX);
        iq.smem_base = smem_u32(smem);
        iq.moff = 0;                         // (-Hi ; Hr ; Hi): M = first two halves, P = last two
        iq.poff = TcSmem::ARR;
        int g = 0;
        for (int tile = 0; tile < ntile; ++tile) {
            for (int ms = 0; ms < nmst; ++ms, ++g) {
                const int stage = g % TC_NSTAGE;
                const int chain = tile * nct + ms / TC_FLUSH, set = chain & 1;
                const bool first = (ms % TC_FLUSH) == 0;
                if (first) {
                    while (next_read <= chain - 2) read_chain(next_read++);
                    if (chain >= 2)
                        mbar_wait_bounded(&tempty[set], (uint32_t)(((chain >> 1) - 1) & 1));
                }
                mbar_wait_bounded(&full[stage], (uint32_t)((g / TC_NSTAGE) & 1));
                tc_fence_after();
                if (elect_one()) {
                    tc_issue_stage(iq, stage, set, first);
                    umma_commit(&empty[stage]);
                    if ((ms % TC_FLUSH) == TC_FLUSH - 1 || ms == nmst - 1) umma_commit(&tfull[set]);
                }
                __syncwarp();
            }
        }
    } else if (cg >= 1) {
        // ---------------- generating groups: thread <-> source row 32 q + lane of the tile, all
        // 16 partner antennas of the stages g = cg - 1 (mod 3)
        const int row = 32 * q + lane;
        const int roff = (row >> 3) * 256 + (row & 7) * 16;
        const unsigned char* Hbase = Hq + ((((size_t)un.x * nfp + k) * nitem + ai) * (size_t)nmst_all + mlo) *
                                              TcBwdSmem::H_BYTES;
        int cur_tile = -1;
        bool valid = false;
        double sv[3] = {0.0, 0.0, 0.0};
        for (int g = cg - 1; g < nst; g += TC_GROUPS) {
            const int tile = g / nmst, ms = g - tile * nmst;
            {
                const int chain = tile * nct + ms / TC_FLUSH;
                while (next_read <= chain - 2) read_chain(next_read++);
            }
            if (tile != cur_tile) {
                cur_tile = tile;
                const long long s = (long long)un.y + (long long)tile * TC_M + row;
                valid = s < un.z;
                sv[0] = sv[1] = sv[2] = 0.0;
                if (valid) {
                    const double2 s01 = __ldg(reinterpret_cast<const double2*>(shat + 4 * s));
                    sv[0] = s01.x, sv[1] = s01.y, sv[2] = __ldg(shat + 4 * s + 2);
                }
            }
            const int stage = g % TC_NSTAGE;
            unsigned char* sbase = smem + stage * TcSmem::STAGE;
            if (g >= TC_NSTAGE)
                mbar_wait_bounded(&empty[stage], (uint32_t)(((g / TC_NSTAGE) - 1) & 1));
            if (q == 0 && lane == 0) {
                // this thread also fetches the cotangent operands of the stage
                mbar_expect_tx(&full[stage], TcBwdSmem::H_BYTES);
                bulk_g2s(sbase + TcSmem::B_H, Hbase + (size_t)ms * TcBwdSmem::H_BYTES,
                         TcBwdSmem::H_BYTES, &full[stage]);
            }
#if B200_TC_UNROLL_BWD
#pragma unroll
#else
#pragma unroll 1
#endif
            for (int kg = 0; kg < 2; ++kg) {
                float c[8], sn[8];
                const double4* pm = pos + (mlo + ms) * TC_KS + 8 * kg;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const double4 p = pm[e];
                    antenna_cis(__fma_rn(p.x, sv[0], __fma_rn(p.y, sv[1], __fma_rn(p.z, sv[2], 1572864.0))),
                                c[e], sn[e]);
                    if (!valid) c[e] = sn[e] = 0.f;
                }
                store_split8(sbase + roff + kg * 128 + TcSmem::XR_H, c, sn);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);
        }
    }
    while (next_read < nchain) read_chain(next_read++);
    if (need_r) {
        float4* dst = reinterpret_cast<float4*>(drpart) +
                      (((size_t)blockIdx.x * nfp + k) * 4 + q) * (size_t)(nitem * TC_M) + a0 + TC_CG * cg;
        dst[lane] = make_float4(gr[0], gr[1], gr[2], 0.f);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_CTRL_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                     "r"(TC_TMEM_COLS)
                     : "memory");
    }
}

// -------------------------------------------------------------------------------------
// Cotangent operand of the backward kernel, straight from the autograd cotangent:
//   H[a][m] = G_b for b = (m, a), conj(G_b) for b = (a, m), 2 Re G_b for an auto-correlation
//   (lower_only: the doubled lower triangle a > m; the upper one is zero)
// scaled by hscale, split into float16 hi / lo and written as the stacked three-half buffers
// (-Hi ; Hr ; Hi) in UMMA canonical order (layout in the header of tc_fringe_bwd_kernel).
// Thread <-> (channel, group of 8 partner antennas, antenna row, stage, item, time), channel
// fastest: the eight gathers of a warp read 256 contiguous bytes of G each, the pair-table
// lookups are warp-uniform.  Replaces a chain of index_put / scale / split / stack / permute
// passes over a (Nt, Nfp, Na, Na) complex matrix (1.1 GB per time at HERA-350 x 1024 channels).
//   G        float2 [nbl][..][nf], element (b, t, k) at G[b * ldb + t * nf + k]
//   pair_bl  int32 [ldp][ldp]  (baseline << 1 | swapped) of the pair (x <= y), -1: none
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tc_pack_cotangent_kernel(const float2* __restrict__ G, long long ldb, const int* __restrict__ pair_bl,
                         int ldp, int nt, int nf, int nfp, int na, int nitem, int nstage,
                         int lower_only, const float* __restrict__ hscale,
                         unsigned char* __restrict__ Hq) {
    const long long total = (long long)nt * nitem * nstage * TC_M * 2 * nfp;
    const float sc = __ldg(hscale);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        long long r = idx;
        const int k = (int)(r % nfp); r /= nfp;
        const int kg = (int)(r % 2); r /= 2;
        const int row = (int)(r % TC_M); r /= TC_M;
        const int stage = (int)(r % nstage); r /= nstage;
        const int item = (int)(r % nitem);
        const int t = (int)(r / nitem);
        const int a = item * TC_M + row;
        float hr[8], hi[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int m = stage * TC_KS + kg * 8 + e;
            hr[e] = hi[e] = 0.f;
            if (k >= nf || a >= na || m >= na || (lower_only && m > a)) continue;
            const int x = min(a, m), y = max(a, m);
            const int ent = __ldg(pair_bl + (size_t)x * ldp + y);
            if (ent < 0) continue;
            const float2 g = __ldg(G + (size_t)(ent >> 1) * ldb + (size_t)t * nf + k);
            if (a == m) {
                hr[e] = 2.f * g.x * sc;
            } else {
                // first antenna of the listed baseline: x, or y when the listing was swapped
                const int first = (ent & 1) ? y : x;
                const float w = lower_only ? 2.f * sc : sc;
                hr[e] = g.x * w;
                hi[e] = (first == m) ? g.y * w : -g.y * w;
            }
        }
        uint4 rh, rl, ih, il;
        split8(hr, rh, rl);
        split8(hi, ih, il);
        unsigned char* d = Hq + ((((size_t)t * nfp + k) * nitem + item) * nstage + stage) * (2 * TcSmem::BBUF)
                           + (row >> 3) * 256 + kg * 128 + (row & 7) * 16;
        *reinterpret_cast<uint4*>(d) = neg_half8(ih);
        *reinterpret_cast<uint4*>(d + TcSmem::ARR) = rh;
        *reinterpret_cast<uint4*>(d + 2 * TcSmem::ARR) = ih;
        *reinterpret_cast<uint4*>(d + TcSmem::BBUF) = neg_half8(il);
        *reinterpret_cast<uint4*>(d + TcSmem::BBUF + TcSmem::ARR) = rl;
        *reinterpret_cast<uint4*>(d + TcSmem::BBUF + 2 * TcSmem::ARR) = il;
    }
}

int launch_tc_fwd(const float* Acm, const float* ascale, const double* shat, const double* antv,
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
    if (attr_once.first()) {
        if (cudaFuncSetAttribute(tc_fringe_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 TcSmem::TOTAL) != cudaSuccess)
            return set_error("tcfringe_fwd: cannot reserve shared memory");
    }
    dim3 grid((unsigned)gx, nunits);
    tc_fringe_fwd_kernel<<<grid, TC_THREADS, TcSmem::TOTAL, st>>>(
        Acm, ascale, shat, antv, freqs, reinterpret_cast<const int4*>(units),
        reinterpret_cast<const int4*>(items), pair_bl, ldp, na, nbl, nfreq, nfp, S,
        (conj ? -1.0 : 1.0) / C_LIGHT, vpart);
    return check_launch("tcfringe_fwd");
}

int launch_tc_bwd(const void* Hq, const float* hscale, const float* Acm, const double* shat,
                  const double* antv, const double* freqs, const int* units, int nunits, int nitem,
                  int na, int nm_pad, const int* mrange, int nfreq, long long S, int conj,
                  float* dAcm, float* drpart, cudaStream_t st) {
    if (nunits <= 0 || nitem <= 0 || nfreq <= 0) return 0;
    if (S % SRC_PAD) return set_error("tcfringe_bwd: S must be a multiple of 128");
    if (nm_pad % TC_KS || nm_pad < na || nm_pad > TcBwdSmem::POS_MAX || nitem * TC_M > TcBwdSmem::POS_MAX ||
        nitem != (na + TC_M - 1) / TC_M)
        return set_error("tcfringe_bwd: antenna count / padding not supported (na <= 512, nm_pad % 16)");
    if (drpart != nullptr && Acm == nullptr)
        return set_error("tcfringe_bwd: the antenna gradient needs the perceived sky");
    const int nfp = ((nfreq + TC_KC - 1) / TC_KC) * TC_KC;
    const long long gy = (long long)nitem * nfreq;
    if (gy > 65535) return set_error("tcfringe_bwd: grid too large");
    static DeviceOnce attr_once;
    if (attr_once.first()) {
        if (cudaFuncSetAttribute(tc_fringe_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 TcBwdSmem::TOTAL) != cudaSuccess)
            return set_error("tcfringe_bwd: cannot reserve shared memory");
    }
    dim3 grid(nunits, (unsigned)gy);
    tc_fringe_bwd_kernel<<<grid, TC_THREADS, TcBwdSmem::TOTAL, st>>>(
        static_cast<const unsigned char*>(Hq), hscale, Acm, shat, antv, freqs,
        reinterpret_cast<const int4*>(units), nitem, na, nm_pad,
        reinterpret_cast<const int2*>(mrange), nfreq, nfp, S, (conj ? -1.0 : 1.0) / C_LIGHT,
        dAcm != nullptr, drpart != nullptr, dAcm, drpart);
    return check_launch("tcfringe_bwd");
}

}  // namespace b200rime

extern "C" {

int b200rime_tcfringe_bwd_f32(const void* Hq, const float* hscale, const float* Acm,
                              const double* shat, const double* antv, const double* freqs,
                              const int* units, int nunits, int nitem, int na, int nm_pad,
                              const int* mrange, int nfreq, long long S, int conj, float* dAcm,
                              float* drpart, void* stream) {
    return b200rime::launch_tc_bwd(Hq, hscale, Acm, shat, antv, freqs, units, nunits, nitem, na,
                                   nm_pad, mrange, nfreq, S, conj, dAcm, drpart,
                                   (cudaStream_t)stream);
}
int b200rime_tcfringe_fwd_f32(const float* Acm, const float* ascale, const double* shat,
                              const double* antv, const double* freqs, const int* units, int nunits,
                              const int* items, int nitems, const int* pair_bl, int ldp, int na,
                              int nbl, int nfreq, long long S, int conj, float* Vpart, void* stream) {
    return b200rime::launch_tc_fwd(Acm, ascale, shat, antv, freqs, units, nunits, items, nitems,
                                   pair_bl, ldp, na, nbl, nfreq, S, conj, Vpart,
                                   (cudaStream_t)stream);
}
int b200rime_tc_pack_cotangent_f32(const float* G, long long ldb, const int* pair_bl, int ldp,
                                   int nt, int nf, int na, int nm_pad, int lower_only,
                                   const float* hscale, void* Hq, void* stream) {
    using namespace b200rime;
    if (nt <= 0 || nf <= 0 || na <= 0) return 0;
    if (nm_pad % TC_KS || nm_pad < na) return set_error("tc_pack_cotangent: nm_pad must be na rounded up to 16");
    if (G == nullptr || pair_bl == nullptr || hscale == nullptr || Hq == nullptr)
        return set_error("tc_pack_cotangent: null operand");
    const int nfp = ((nf + TC_KC - 1) / TC_KC) * TC_KC;
    const int nitem = (na + TC_M - 1) / TC_M, nstage = nm_pad / TC_KS;
    const long long total = (long long)nt * nitem * nstage * TC_M * 2 * nfp;
    const long long want = (total + 255) / 256;
    const int grid = (int)(want < 148LL * 64 ? want : 148LL * 64);
    tc_pack_cotangent_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(G), ldb, pair_bl, ldp, nt, nf, nfp, na, nitem, nstage,
        lower_only, hscale, static_cast<unsigned char*>(Hq));
    return check_launch("tc_pack_cotangent");
}
int b200rime_tc_rows(void) { return b200rime::TC_M; }
int b200rime_tc_cols_max(void) { return b200rime::TC_NMAX; }

}  // extern "C"
