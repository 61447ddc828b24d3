// Antenna-factorised fringe-sum kernels for sm_100a (float32, all-pairs style baseline sets).
//
// The fringe of baseline (i, j) is a product of two antenna terms,
//     exp(2 pi i sgn (r_j - r_i).shat nu / c) = conj(E_i) E_j,   E_a = exp(2 pi i sgn r_a.shat nu / c),
// so for a dense set of antenna pairs the source sum is, per channel, V = E^H diag(A) E and one
// source.baseline.channel evaluation is a single complex multiply-accumulate = 2 packed FFMA2
// (4 FP32-pipe lane-cycles, every product a fused multiply-add) instead of the 3 packed
// instructions (6 lane-cycles) of the rotation-recurrence kernels in fringe_kernels.cu.  The
// antenna terms are generated on the fly inside the CTA, straight from float64 phases (r_a.shat
// and the fraction of a cycle in float64, MUFU sine / cosine), so no fringe -- per baseline or per
// antenna -- ever reaches HBM.  FP32 FMA pipes only: the tensor cores are not used.
//
// Execution model (both kernels).  A CTA has 384 threads: warps 0..7 are consumers, warps 8..11
// producers, with setmaxnreg moving registers from the producer warpgroup to the two consumer
// warpgroups.  It runs four independent pipelines ("slots"), one per SM sub-partition: slot q owns
// one (tile, channel) work item, a 4-stage shared-memory ring (8 KB per stage) and full / empty
// mbarriers; its producer warp fills stages, its two consumer warps drain them.  A consumer thread
// keeps an 8 x 8 block of complex outputs in 128 accumulator registers and reads, per reduction
// index, 8 x-values and 8 y-values (8 LDS.128 per 128 FFMA2; a 4 x 4 tile would need exactly the
// shared-memory bandwidth the SM has).
//
// Operand layout of a stage (ANT_ST reduction indices r):
//   X[r][64]      complex (re, im) pairs in `xpos` order: a thread's 8 values are four 16-byte
//                 quarters, each quarter contiguous over the 8 thread-rows of a quarter-warp, so
//                 every LDS.128 phase reads 128 contiguous bytes.  Used as scalar-broadcast FFMA2
//                 operands with the negation folded into the operand.
//   YR/YI[r][64]  split real / imaginary rows, read as packed pairs: accumulators pair two
//                 neighbouring outputs (re_j0, re_j1), (im_j0, im_j1), so the complex product
//                 needs no swap, no negate and no extra instruction.
// Forward:  X = E_i (conjugated in the product), Y = A_s E_j, reduction over sources.
// Backward: X = H[a, m] (Hermitian cotangent matrix, TMA bulk copies), Y = E_m over 64 sources,
//           reduction over partner antennas m:  y_a = sum_m H[a, m] E_m, then with
//           p = conj(E_a) y_a:  dL/dA[s, k] = 1/2 sum_a Re p  and
//           dL/dr_a = sum_{s,k} shat_s A[s,k] (2 pi sgn nu_k / c) Im p.
//           (H[a][m] = G_b for b = (m, a), conj(G_b) for b = (a, m): sum_b Re(conj(F_b) G_b)
//           = 1/2 e^H H e, and d/dr_a of it is the imaginary part of the same products.)
//
// Replaces, like fringe_kernels.cu, telescope_model.py:310-358 + rime_model.py:426-429 and their
// autograd backward.  Every output has one owner and a fixed summation order.
#include "rime_math.cuh"
#include "internal.h"

namespace b200rime {

#ifndef B200_ANT_NSTAGE
#define B200_ANT_NSTAGE 4
#endif
#ifndef B200_ANT_ST
#define B200_ANT_ST 8
#endif
#ifndef B200_ANT_DECORR      // consumer warp `half 1` of slot q runs on the SMSP of slot q - 1
#define B200_ANT_DECORR 1
#endif
#ifndef B200_ANT_PROBE       // tuning probes (wrong results): 1 = producers only signal,
#define B200_ANT_PROBE 0     // 2 = consumers only signal, 3 = no backward epilogue, 4 = no TMA
#endif
constexpr int ANT_TILE = 64;       // antennas per tile side
constexpr int ANT_ST = B200_ANT_ST;   // reduction indices per shared-memory stage (8 or 16)
constexpr int ANT_NSTAGE = B200_ANT_NSTAGE;   // stages in flight between producer and consumers
constexpr int ANT_SLOTS = 4;       // independent (tile, channel) pipelines per CTA, one per SMSP
constexpr int ANT_CONSUMERS = 256; // warps 0..7: multiply-accumulate; warp = half * 4 + slot
constexpr int ANT_PRODUCERS = 128; // warps 8..11: operand generation; warp = 8 + slot
constexpr int ANT_THREADS = ANT_CONSUMERS + ANT_PRODUCERS;
// registers per thread after setmaxnreg; 128 * producer + 256 * consumer = 64512 = the CTA's
// pool (384 threads * 168): more cannot be granted and setmaxnreg.inc would spin for ever
constexpr int ANT_PROD_REGS = 72, ANT_CONS_REGS = 216;          // forward
constexpr int ANT_PROD_REGS_BWD = 72, ANT_CONS_REGS_BWD = 216;  // backward (224 / 56 measured 3% slower)
constexpr int ANT_KC = B200_KC_F32;

struct AntSmem {
    static constexpr int X_BYTES = ANT_ST * ANT_TILE * 8;          // 4 KB complex operand rows
    static constexpr int Y_BYTES = ANT_ST * ANT_TILE * 4;          // 2 KB each (re / im rows)
    static constexpr int STAGE_BYTES = X_BYTES + 2 * Y_BYTES;      // 8 KB
    static constexpr int SLOT_BYTES = ANT_NSTAGE * STAGE_BYTES;    // 32 KB
    static constexpr int FLAG_OFF = ANT_SLOTS * SLOT_BYTES;        // consumer-warp activity flags
    static constexpr int BAR_OFF = FLAG_OFF + 64;                  // full / empty [slot][stage]
    static constexpr int TOTAL = BAR_OFF + 2 * ANT_SLOTS * ANT_NSTAGE * 8;
};

// position of antenna slot a (0..63) inside an X row of 64 complex numbers: the eight slots of
// thread-row ti = a / 8 are split into four 16-byte quarters, each quarter contiguous over ti,
// so that the 8 lanes of a quarter-warp read 128 contiguous bytes per LDS.128
__device__ __forceinline__ int xpos(int a) {
    return (((a >> 1) & 3) << 4) | ((a >> 3) << 1) | (a & 1);
}
__device__ __forceinline__ int xpos_inv(int p) {      // antenna slot stored at position p
    return (((p >> 1) & 7) << 3) | ((p >> 4) << 1) | (p & 1);
}

// acc += conj(x) * y (CONJ) or x * y for 8 x-values (scalar-broadcast operands, negation folded
// into the operand) times 8 y-values (four packed pairs); aR / aI hold the real / imaginary
// parts of the output pairs [i][jp]
template <bool CONJ>
__device__ __forceinline__ void ant_mac(P2 (&aR)[8][4], P2 (&aI)[8][4], const float4 (&x)[4],
                                        const float4 (&yr)[2], const float4 (&yi)[2]) {
    const P2 YR[4] = {p2(yr[0].x, yr[0].y), p2(yr[0].z, yr[0].w), p2(yr[1].x, yr[1].y),
                      p2(yr[1].z, yr[1].w)};
    const P2 YI[4] = {p2(yi[0].x, yi[0].y), p2(yi[0].z, yi[0].w), p2(yi[1].x, yi[1].y),
                      p2(yi[1].z, yi[1].w)};
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const float xr[2] = {x[h].x, x[h].z}, xi[2] = {x[h].y, x[h].w};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = 2 * h + e;
            const P2 XR = p2(xr[e], xr[e]), XI = p2(xi[e], xi[e]), NXI = p2(-xi[e], -xi[e]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                p2_mac(aR[i][j], XR, YR[j]);
                p2_mac(aI[i][j], XR, YI[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                p2_mac(aR[i][j], CONJ ? XI : NXI, YI[j]);
                p2_mac(aI[i][j], CONJ ? NXI : XI, YR[j]);
            }
        }
    }
}

// one shared-memory stage: ANT_ST reduction indices; the thread's 8 x-values are antenna slots
// 8 ti .. 8 ti + 7, its 8 y-values are row entries 32 half + 8 tj .. + 7
template <bool CONJ>
__device__ __forceinline__ void ant_mac_stage(P2 (&aR)[8][4], P2 (&aI)[8][4],
                                              const unsigned char* buf, int ti, int yq) {
    const float4* X4 = reinterpret_cast<const float4*>(buf);
    const float4* YR4 = reinterpret_cast<const float4*>(buf + AntSmem::X_BYTES);
    const float4* YI4 = reinterpret_cast<const float4*>(buf + AntSmem::X_BYTES + AntSmem::Y_BYTES);
#pragma unroll
    for (int r = 0; r < ANT_ST; ++r) {
        float4 x[4], yr[2], yi[2];
#pragma unroll
        for (int h = 0; h < 4; ++h) x[h] = X4[r * 32 + h * 8 + ti];
        yr[0] = YR4[r * 16 + yq];
        yr[1] = YR4[r * 16 + yq + 1];
        yi[0] = YI4[r * 16 + yq];
        yi[1] = YI4[r * 16 + yq + 1];
        ant_mac<CONJ>(aR, aI, x, yr, yi);
    }
}

// antenna term E = exp(2 pi i frac(u kappa)), u = r_a . shat [m], kappa = sgn nu / c [cycles/m]
__device__ __forceinline__ void antenna_term(double u, double kappa, float& c, float& s) {
    cis_fast((float)frac_cycles(u, kappa), c, s);
}
__device__ __forceinline__ double dot3(const double (&a)[3], const double2 s01, const double s2) {
    return __fma_rn(a[0], s01.x, __fma_rn(a[1], s01.y, a[2] * s2));
}

// warp -> (slot, half).  Producers: warp 8 + slot.  Consumers: warps 0..3 are `half 0` of slots
// 0..3; warps 4..7 are `half 1`, by default of the NEXT slot, so that the two consumer warps that
// share an SM sub-partition belong to different pipelines and do not stall in lockstep.
__device__ __forceinline__ int ant_slot_of(int warp) {
    if (warp >= 2 * ANT_SLOTS || warp < ANT_SLOTS || !B200_ANT_DECORR) return warp & (ANT_SLOTS - 1);
    return (warp + 1) & (ANT_SLOTS - 1);
}

// one consumer pass over the stages [it0, it1) of a slot
template <bool CONJ>
__device__ __forceinline__ void ant_consume(P2 (&aR)[8][4], P2 (&aI)[8][4], const unsigned char* sbuf,
                                            uint64_t* full, uint64_t* empty, long long g0,
                                            int nstages, int ti, int yq, int lane) {
    // nstages is even (forward: units are multiples of 64 sources; backward: the partner axis is
    // padded to 16), so the loop body holds two stages and ptxas amortises its accumulator
    // fix-up moves over 2048 FFMA2
    long long g = g0;
    if (ANT_ST >= 16) {                 // one 16-deep stage per iteration
        for (int n = nstages; n > 0; --n, ++g) {
            const int sa = (int)(g % ANT_NSTAGE);
            mbar_wait(&full[sa], (uint32_t)((g / ANT_NSTAGE) & 1));
            if (B200_ANT_PROBE != 2) ant_mac_stage<CONJ>(aR, aI, sbuf + sa * AntSmem::STAGE_BYTES, ti, yq);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[sa]);
        }
        return;
    }
    for (int n = nstages; n > 0; n -= 2, g += 2) {
        const int sa = (int)(g % ANT_NSTAGE), sb = (int)((g + 1) % ANT_NSTAGE);
        mbar_wait(&full[sa], (uint32_t)((g / ANT_NSTAGE) & 1));
        if (B200_ANT_PROBE != 2) ant_mac_stage<CONJ>(aR, aI, sbuf + sa * AntSmem::STAGE_BYTES, ti, yq);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[sa]);
        mbar_wait(&full[sb], (uint32_t)(((g + 1) / ANT_NSTAGE) & 1));
        if (B200_ANT_PROBE != 2) ant_mac_stage<CONJ>(aR, aI, sbuf + sb * AntSmem::STAGE_BYTES, ti, yq);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[sb]);
    }
}

// shared prologue: barriers of the four slot pipelines.  flags[w] = consumer warp w takes part.
__device__ __forceinline__ void ant_init_barriers(unsigned char* smem, int tid, int lane, int warp,
                                                  bool active, int full_count) {
    int* flags = reinterpret_cast<int*>(smem + AntSmem::FLAG_OFF);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + AntSmem::BAR_OFF);
    uint64_t* empty = full + ANT_SLOTS * ANT_NSTAGE;
    if (warp < ANT_CONSUMERS / 32 && lane == 0)
        flags[(warp >> 2) * ANT_SLOTS + ant_slot_of(warp)] = active ? 1 : 0;
    __syncthreads();
    if (tid < ANT_SLOTS) {
        const int n = flags[tid] + flags[tid + ANT_SLOTS];
#pragma unroll
        for (int st = 0; st < ANT_NSTAGE; ++st) {
            mbar_init(&full[tid * ANT_NSTAGE + st], full_count);
            mbar_init(&empty[tid * ANT_NSTAGE + st], n > 0 ? n : 1);
        }
        mbar_fence_init();
    }
    __syncthreads();
}

// -------------------------------------------------------------------------------------
// forward.  grid = (ntile * Nfp / 4, nunits), block = 384.
// A CTA runs four independent pipelines ("slots", one per SM sub-partition): slot q works on
// item 4 blockIdx.x + q = (tile, channel), channel fastest, tiles in `tile_order` (tiles that
// need both consumer warps first).  Per slot: producer warp 8 + q generates, per stage of
// ANT_ST sources, X = E of the tile's 64 first antennas and Y = A_s E of its 64 second antennas;
// consumer warps q (second antennas 0..31) and 4 + q (32..63) accumulate conj(X) Y with 8 x 8
// register tiles.  tile_bl[tile][x][y] = (baseline << 1 | conj) or -1.
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ANT_THREADS, 1)
ant_fringe_fwd_kernel(const float* __restrict__ A, const double* __restrict__ shat,
                      const double* __restrict__ antv, const double* __restrict__ freqs,
                      const int4* __restrict__ units, const int* __restrict__ tile_ant,
                      const int* __restrict__ tile_bl, const int* __restrict__ tile_order,
                      int nitems, int nk, int nbl, int nfreq, long long S, double sgn_over_c,
                      float* __restrict__ vpart) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = ant_slot_of(warp);
    const int half = (warp >> 2) & 1;
    const int item = blockIdx.x * ANT_SLOTS + slot;
    const bool valid = item < nitems;
    const int tile = valid ? __ldg(tile_order + item / nk) : 0;
    const int k = valid ? item % nk : 0;
    const int4 un = units[blockIdx.y];
    const int nst = (un.z - un.y) / ANT_ST;

    // consumer role: which outputs are wanted
    const int ti = lane & 7, tj = lane >> 3;
    const int* tb = tile_bl + (size_t)tile * (ANT_TILE * ANT_TILE) + (8 * ti) * ANT_TILE +
                    32 * half + 8 * tj;
    bool mine = false;
    if (valid && warp < ANT_CONSUMERS / 32) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int4 e0 = __ldg(reinterpret_cast<const int4*>(tb + i * ANT_TILE));
            const int4 e1 = __ldg(reinterpret_cast<const int4*>(tb + i * ANT_TILE + 4));
            mine |= (e0.x >= 0) | (e0.y >= 0) | (e0.z >= 0) | (e0.w >= 0) | (e1.x >= 0) |
                    (e1.y >= 0) | (e1.z >= 0) | (e1.w >= 0);
        }
    }
    const bool active = __any_sync(0xffffffffu, mine) && nst > 0;
    ant_init_barriers(smem, tid, lane, warp, active, 1);
    const int* flags = reinterpret_cast<const int*>(smem + AntSmem::FLAG_OFF);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + AntSmem::BAR_OFF) + slot * ANT_NSTAGE;
    uint64_t* empty = full + ANT_SLOTS * ANT_NSTAGE;
    unsigned char* sbuf = smem + slot * AntSmem::SLOT_BYTES;
    const bool slot_live = (flags[slot] + flags[slot + ANT_SLOTS]) > 0;

    if (warp >= ANT_CONSUMERS / 32) {
        // ---------------- producer warp of this slot
        setmaxnreg_dec<ANT_PROD_REGS>();
        if (!slot_live) return;
        // lane owns X positions lane, lane + 32 (conflict-free stores) and Y slots lane, lane + 32
        double pos[4][3];
        int xp[2];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int a;
            if (q < 2) {
                xp[q] = lane + 32 * q;
                a = xpos_inv(xp[q]);
            } else {
                a = ANT_TILE + lane + 32 * (q - 2);
            }
            const int ant = __ldg(tile_ant + tile * (2 * ANT_TILE) + a);
            pos[q][0] = pos[q][1] = pos[q][2] = 0.0;
            if (ant >= 0) {
                pos[q][0] = antv[4 * (size_t)ant];
                pos[q][1] = antv[4 * (size_t)ant + 1];
                pos[q][2] = antv[4 * (size_t)ant + 2];
            }
        }
        const double kappa = (k < nfreq) ? sgn_over_c * freqs[k] : 0.0;
        const float* Ak = A + (size_t)(k / ANT_KC) * (size_t)S * ANT_KC + (k % ANT_KC);
        for (int it = 0; it < nst; ++it) {
            const int stage = it % ANT_NSTAGE;
            if (it >= ANT_NSTAGE) mbar_wait_relaxed(&empty[stage], ((it / ANT_NSTAGE) - 1) & 1);
            unsigned char* buf = sbuf + stage * AntSmem::STAGE_BYTES;
            float2* X2 = reinterpret_cast<float2*>(buf);
            float* YR = reinterpret_cast<float*>(buf + AntSmem::X_BYTES);
            float* YI = reinterpret_cast<float*>(buf + AntSmem::X_BYTES + AntSmem::Y_BYTES);
            const long long sbase = (long long)un.y + (long long)it * ANT_ST;
#pragma unroll 2
            for (int sl = 0; sl < (B200_ANT_PROBE == 1 ? 0 : ANT_ST); ++sl) {
                const long long s = sbase + sl;
                const double2 s01 = __ldg(reinterpret_cast<const double2*>(shat + 4 * s));
                const double s2 = __ldg(shat + 4 * s + 2);
                const float a = __ldg(Ak + s * ANT_KC);
                float c[4], sn[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) antenna_term(dot3(pos[q], s01, s2), kappa, c[q], sn[q]);
                X2[sl * ANT_TILE + xp[0]] = make_float2(c[0], sn[0]);
                X2[sl * ANT_TILE + xp[1]] = make_float2(c[1], sn[1]);
                YR[sl * ANT_TILE + lane] = a * c[2];
                YI[sl * ANT_TILE + lane] = a * sn[2];
                YR[sl * ANT_TILE + lane + 32] = a * c[3];
                YI[sl * ANT_TILE + lane + 32] = a * sn[3];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);
        }
        return;
    }

    // ---------------- consumer warps
    setmaxnreg_inc<ANT_CONS_REGS>();
    if (!active) return;
    P2 aR[8][4], aI[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) aR[i][j] = aI[i][j] = p2(0.f, 0.f);
    const int yq = half * 8 + tj * 2;
    ant_consume<true>(aR, aI, sbuf, full, empty, 0, nst, ti, yq, lane);

    if (!mine) return;
    float* vp = vpart + ((size_t)blockIdx.y * (size_t)nbl * nk + k) * 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int4 e0 = __ldg(reinterpret_cast<const int4*>(tb + i * ANT_TILE));
        const int4 e1 = __ldg(reinterpret_cast<const int4*>(tb + i * ANT_TILE + 4));
        const int e[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (e[j] < 0) continue;
            float r0, r1, i0, i1;
            p2_get(aR[i][j >> 1], r0, r1);
            p2_get(aI[i][j >> 1], i0, i1);
            const float re = (j & 1) ? r1 : r0;
            const float im = (j & 1) ? i1 : i0;
            *reinterpret_cast<float2*>(vp + (size_t)(e[j] >> 1) * nk * 2) =
                make_float2(re, (e[j] & 1) ? -im : im);
        }
    }
}

// -------------------------------------------------------------------------------------
// backward.  grid = (nunits, nblk * Nfp / 4), block = 384.
// Slot item = (antenna block ib, channel k), channel fastest.  The slot walks the unit in tiles
// of 64 sources and, per tile, reduces over all partner antennas m in stages of ANT_ST:
//   X = H[a, m] for the 64 antennas a of block ib (TMA-staged from the pre-arranged Hermitian
//       cotangent), Y = E_m over the 64 sources (generated by the producer warp);
// consumer warp `half` owns sources 32 half .. + 31, thread (ti, tj) antennas 8 ti .. + 7 and
// sources 8 tj .. + 7 of those.  After the last stage y_a = sum_m H[a, m] E_m is complete and
// with p = conj(E_a) y_a the thread adds Re p to dL/dA and shat A kappa Im p to dL/dr_a.
//   Hp[t][k][ib][mstage][r][64 a (xpos order)] complex64, mstage < nm_pad / 8
//   dApart[ib][chunk][S][KC]                  (summed over ib by the caller)
//   drpart[unit][k][half][ib * 64 + a][4]     float64 (summed over unit, k, half by the caller)
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ANT_THREADS, 1)
ant_fringe_bwd_kernel(const float* __restrict__ Hp, const float* __restrict__ A,
                      const double* __restrict__ shat, const double* __restrict__ antv,
                      const double* __restrict__ freqs, const int4* __restrict__ units, int nitems,
                      int nk, int na, int na_pad, int nm_pad, int nfreq, long long S,
                      double sgn_over_c, int need_a, int need_r, float* __restrict__ dApart,
                      double* __restrict__ drpart) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = ant_slot_of(warp);
    const int half = (warp >> 2) & 1;
    const int item = blockIdx.y * ANT_SLOTS + slot;
    const bool valid = item < nitems;
    const int ib = valid ? item / nk : 0;
    const int k = valid ? item % nk : 0;
    const int nblk = na_pad / ANT_TILE;
    const int4 un = units[blockIdx.x];
    // partner-antenna stages per source tile.  Without the antenna gradient only dL/dA is wanted,
    // Re(e^H H e) = 2 Re sum_{a > m} conj(E_a) H[a,m] E_m: the caller passes the lower triangle
    // of H (doubled) and block ib stops after its own antennas -- 21 instead of 36 block products
    // for six blocks
    const int nmst_all = nm_pad / ANT_ST;
    const int nmst = need_r ? nmst_all : min(nmst_all, (ib + 1) * (ANT_TILE / ANT_ST));
    const int nsrc_tiles = (un.z - un.y) / ANT_TILE;
    // a last block with at most 32 antennas is "narrow": one consumer warp covers its 32 antennas
    // x all 64 sources (thread: 8 antennas x 8 sources, 4 x 8 threads), the other one retires
    const bool narrow = na - ib * ANT_TILE <= ANT_TILE / 2;
    const bool active = valid && nsrc_tiles > 0 &&
                        !(narrow && half == 1 && warp < ANT_CONSUMERS / 32);

    ant_init_barriers(smem, tid, lane, warp, active, 2);     // producer warp + TMA expect_tx
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + AntSmem::BAR_OFF) + slot * ANT_NSTAGE;
    uint64_t* empty = full + ANT_SLOTS * ANT_NSTAGE;
    unsigned char* sbuf = smem + slot * AntSmem::SLOT_BYTES;
    const double kappa = (k < nfreq) ? sgn_over_c * freqs[k] : 0.0;

    if (warp >= ANT_CONSUMERS / 32) {
        // ---------------- producer warp: lane owns sources lane and lane + 32 of the tile
        setmaxnreg_dec<ANT_PROD_REGS_BWD>();
        if (!active) return;
        const float* Hbase = Hp + ((((size_t)un.x * nk + k) * nblk + ib) * (size_t)nmst_all) *
                                      (AntSmem::X_BYTES / 4);
        long long g = 0;
        for (int st = 0; st < nsrc_tiles; ++st) {
            const long long s = (long long)un.y + (long long)st * ANT_TILE + lane;
            const double2 sa01 = __ldg(reinterpret_cast<const double2*>(shat + 4 * s));
            const double sa2 = __ldg(shat + 4 * s + 2);
            const double2 sb01 = __ldg(reinterpret_cast<const double2*>(shat + 4 * (s + 32)));
            const double sb2 = __ldg(shat + 4 * (s + 32) + 2);
            for (int ms = 0; ms < nmst; ++ms, ++g) {
                const int stage = (int)(g % ANT_NSTAGE);
                if (g >= ANT_NSTAGE)
                    mbar_wait_relaxed(&empty[stage], (uint32_t)(((g / ANT_NSTAGE) - 1) & 1));
                unsigned char* buf = sbuf + stage * AntSmem::STAGE_BYTES;
                if (lane == 0) {
                    if (B200_ANT_PROBE == 4 && g > 0) {
                        mbar_arrive(&full[stage]);
                    } else {
                        mbar_expect_tx(&full[stage], AntSmem::X_BYTES);
                        bulk_g2s(buf, Hbase + (size_t)ms * (AntSmem::X_BYTES / 4),
                                 AntSmem::X_BYTES, &full[stage]);
                    }
                }
                float* YR = reinterpret_cast<float*>(buf + AntSmem::X_BYTES);
                float* YI = reinterpret_cast<float*>(buf + AntSmem::X_BYTES + AntSmem::Y_BYTES);
#pragma unroll 2
                for (int r = 0; r < (B200_ANT_PROBE == 1 ? 0 : ANT_ST); ++r) {
                    const double* ap = antv + 4 * (size_t)(ms * ANT_ST + r);
                    const double2 a01 = __ldg(reinterpret_cast<const double2*>(ap));
                    const double a2 = __ldg(ap + 2);
                    const double am[3] = {a01.x, a01.y, a2};
                    float c0, s0, c1, s1;
                    antenna_term(dot3(am, sa01, sa2), kappa, c0, s0);
                    antenna_term(dot3(am, sb01, sb2), kappa, c1, s1);
                    YR[r * ANT_TILE + lane] = c0;
                    YI[r * ANT_TILE + lane] = s0;
                    YR[r * ANT_TILE + lane + 32] = c1;
                    YI[r * ANT_TILE + lane + 32] = s1;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[stage]);
            }
        }
        return;
    }

    // ---------------- consumer warps
    setmaxnreg_inc<ANT_CONS_REGS_BWD>();
    if (!active) return;
    const int ti = narrow ? (lane & 3) : (lane & 7);
    const int tj = narrow ? (lane >> 2) : (lane >> 3);
    const int yq = narrow ? tj * 2 : half * 8 + tj * 2;
    float gx[8], gy[8], gz[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) gx[i] = gy[i] = gz[i] = 0.f;

    P2 aR[8][4], aI[8][4];
    long long g = 0;
    for (int st = 0; st < nsrc_tiles; ++st) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) aR[i][j] = aI[i][j] = p2(0.f, 0.f);
        ant_consume<false>(aR, aI, sbuf, full, empty, g, nmst, ti, yq, lane);
        g += nmst;
        // ---- epilogue of this source tile: p = conj(E_a) y_a
        if (B200_ANT_PROBE == 3 && st > 0) continue;
        const long long sb = (long long)un.y + (long long)st * ANT_TILE + (narrow ? 0 : 32 * half) +
                             8 * tj;
        const float kf = (float)kappa;
        const float* Ak = A + (size_t)(k / ANT_KC) * (size_t)S * ANT_KC + (k % ANT_KC);
        float* dAk = dApart + ((size_t)ib * (nk / ANT_KC) + (k / ANT_KC)) * (size_t)S * ANT_KC +
                     (k % ANT_KC);
        const double* apos = antv + 4 * (size_t)(ib * ANT_TILE + 8 * ti);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const long long s = sb + j;
            const double2 s01 = __ldg(reinterpret_cast<const double2*>(shat + 4 * s));
            const double s2 = __ldg(shat + 4 * s + 2);
            const float wk = need_r ? __ldg(Ak + s * ANT_KC) * kf : 0.f;   // A may be NULL
            const float sx = (float)s01.x, sy = (float)s01.y, sz = (float)s2;
            float dacc = 0.f;       // sum over own 8 antennas of Re p
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double2 p01 = __ldg(reinterpret_cast<const double2*>(apos + 4 * i));
                const double pz = __ldg(apos + 4 * i + 2);
                const double am[3] = {p01.x, p01.y, pz};
                float c, sn, r0, r1, i0, i1;
                antenna_term(dot3(am, s01, s2), kappa, c, sn);
                p2_get(aR[i][j >> 1], r0, r1);
                p2_get(aI[i][j >> 1], i0, i1);
                const float yr = (j & 1) ? r1 : r0, yi = (j & 1) ? i1 : i0;
                dacc = fmaf(c, yr, fmaf(sn, yi, dacc));              // Re(conj(E) y)
                const float w = wk * fmaf(c, yi, -sn * yr);          // A kappa Im(conj(E) y)
                gx[i] = fmaf(w, sx, gx[i]);
                gy[i] = fmaf(w, sy, gy[i]);
                gz[i] = fmaf(w, sz, gz[i]);
            }
            if (need_a) {
                // sum over the thread-rows ti (lane bits 0..2, or 0..1 in a narrow block), fixed
                // order
                dacc += __shfl_xor_sync(0xffffffffu, dacc, 1);
                dacc += __shfl_xor_sync(0xffffffffu, dacc, 2);
                const float d4 = __shfl_xor_sync(0xffffffffu, dacc, 4);
                if (!narrow) dacc += d4;
                if (ti == 0) dAk[s * ANT_KC] = 0.5f * dacc;
            }
        }
    }

    if (need_r) {
        // sum the antenna gradients over the thread-columns tj (lane bits 3, 4; 2..4 when narrow);
        // rows the kernel does not own (second half, antennas 32.. of a narrow block) are zeroed
        // by the caller
        const double twopi = 6.283185307179586476925286766559;
        double* dst = drpart + ((((size_t)blockIdx.x * nk + k) * 2 + half) * (size_t)na_pad +
                                ib * ANT_TILE + 8 * ti) * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v[3] = {gx[i], gy[i], gz[i]};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v4 = __shfl_xor_sync(0xffffffffu, v[c], 4);
                if (narrow) v[c] += v4;
                v[c] += __shfl_xor_sync(0xffffffffu, v[c], 8);
                v[c] += __shfl_xor_sync(0xffffffffu, v[c], 16);
            }
            if (tj == 0) {
                dst[4 * i] = twopi * (double)v[0];
                dst[4 * i + 1] = twopi * (double)v[1];
                dst[4 * i + 2] = twopi * (double)v[2];
                dst[4 * i + 3] = 0.0;
            }
        }
    }
}

// -------------------------------------------------------------------------------------
// launchers
// -------------------------------------------------------------------------------------
int launch_ant_fwd(const float* A, const double* shat, const double* antv, const double* freqs,
                   const int* units, int nunits, const int* tile_ant, const int* tile_bl,
                   const int* tile_order, int ntile, int nbl, int nfreq, long long S, int conj,
                   float* vpart, cudaStream_t st) {
    if (nunits <= 0 || ntile <= 0 || nfreq <= 0) return 0;
    if (S % SRC_PAD) return set_error("antfringe_fwd: S must be a multiple of 128");
    const int nfp = ((nfreq + ANT_KC - 1) / ANT_KC) * ANT_KC;
    const long long nitems = (long long)ntile * nfp;
    if (nitems / ANT_SLOTS > 2147483647LL || nunits > 65535)
        return set_error("antfringe_fwd: grid too large");
    static DeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(ant_fringe_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             AntSmem::TOTAL);
    }
    dim3 grid((unsigned)((nitems + ANT_SLOTS - 1) / ANT_SLOTS), nunits);
    ant_fringe_fwd_kernel<<<grid, ANT_THREADS, AntSmem::TOTAL, st>>>(
        A, shat, antv, freqs, reinterpret_cast<const int4*>(units), tile_ant, tile_bl, tile_order,
        (int)nitems, nfp, nbl, nfreq, S, (conj ? -1.0 : 1.0) / C_LIGHT, vpart);
    return check_launch("antfringe_fwd");
}

int launch_ant_bwd(const float* Hp, const float* A, const double* shat, const double* antv,
                   const double* freqs, const int* units, int nunits, int na, int na_pad, int nm_pad,
                   int nfreq, long long S, int conj, float* dApart, double* drpart,
                   cudaStream_t st) {
    if (nunits <= 0 || na_pad <= 0 || nfreq <= 0) return 0;
    if (S % SRC_PAD) return set_error("antfringe_bwd: S must be a multiple of 128");
    if (na_pad % ANT_TILE || na > na_pad || na <= na_pad - ANT_TILE)
        return set_error("antfringe_bwd: na_pad must be the antenna count rounded up to 64");
    if (nm_pad % (ANT_ST >= 16 ? ANT_ST : 2 * ANT_ST) || nm_pad > na_pad || nm_pad <= 0)
        return set_error("antfringe_bwd: partner axis must be padded to 16 and fit na_pad");
    const int nfp = ((nfreq + ANT_KC - 1) / ANT_KC) * ANT_KC;
    const long long nitems = (long long)(na_pad / ANT_TILE) * nfp;
    if ((nitems + ANT_SLOTS - 1) / ANT_SLOTS > 65535)
        return set_error("antfringe_bwd: grid too large");
    static DeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(ant_fringe_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             AntSmem::TOTAL);
    }
    // units fastest: CTAs that run together share their (antenna block, channel) items, so a
    // cotangent tile is fetched from HBM once and then served to the other units from L2
    dim3 grid(nunits, (unsigned)((nitems + ANT_SLOTS - 1) / ANT_SLOTS));
    ant_fringe_bwd_kernel<<<grid, ANT_THREADS, AntSmem::TOTAL, st>>>(
        Hp, A, shat, antv, freqs, reinterpret_cast<const int4*>(units), (int)nitems, nfp, na, na_pad,
        nm_pad, nfreq, S, (conj ? -1.0 : 1.0) / C_LIGHT, dApart != nullptr, drpart != nullptr, dApart,
        drpart);
    return check_launch("antfringe_bwd");
}

}  // namespace b200rime

extern "C" {

int b200rime_antfringe_fwd_f32(const float* A, const double* shat, const double* antv,
                               const double* freqs, const int* units, int nunits,
                               const int* tile_ant, const int* tile_bl, const int* tile_order,
                               int ntile, int nbl, int nfreq, long long S, int conj, float* Vpart,
                               void* stream) {
    return b200rime::launch_ant_fwd(A, shat, antv, freqs, units, nunits, tile_ant, tile_bl,
                                    tile_order, ntile, nbl, nfreq, S, conj, Vpart,
                                    (cudaStream_t)stream);
}
int b200rime_antfringe_bwd_f32(const float* Hp, const float* A, const double* shat,
                               const double* antv, const double* freqs, const int* units,
                               int nunits, int na, int na_pad, int nm_pad, int nfreq, long long S,
                               int conj, float* dApart, double* drpart, void* stream) {
    return b200rime::launch_ant_bwd(Hp, A, shat, antv, freqs, units, nunits, na, na_pad, nm_pad,
                                    nfreq, S, conj, dApart, drpart, (cudaStream_t)stream);
}
int b200rime_ant_tile(void) { return b200rime::ANT_TILE; }
int b200rime_ant_stage(void) { return b200rime::ANT_ST; }

}  // extern "C"
