// Antenna-factorised fringe-sum kernels for sm_100a (float32, all-pairs style baseline sets).
//
// The fringe of baseline (i, j) is a product of two antenna terms,
//     exp(2 pi i sgn (r_j - r_i).shat nu / c) = conj(E_i) E_j,   E_a = exp(2 pi i sgn r_a.shat nu / c),
// so for a dense set of antenna pairs the source sum is, per channel, V = E^H diag(A) E.  A CTA
// owns a 64 x 64 block of antenna pairs and 4 channels; every thread keeps a 4 x 4 block of
// complex visibilities per channel in registers (128 accumulator registers) and the antenna
// terms are generated on the fly by the CTA itself, straight from float64 phases (r_a.shat in
// registers, fraction of a cycle in float64, MUFU sine/cosine), into a double-buffered shared
// memory stage.  One source.baseline.channel evaluation is then a single complex
// multiply-accumulate = 2 packed FFMA2 (4 FP32-pipe lane-cycles) instead of the 3 packed
// instructions (6 lane-cycles) of the rotation-recurrence kernels in fringe_kernels.cu, and no
// fringe -- per baseline or per antenna -- ever reaches HBM.  Still FP32 FMA pipes only: the
// tensor cores are not used.
//
// Operand layout in shared memory (per stage, per channel k and reduction index r):
//   X[k][r][64]  complex (re, im) pairs, read as two warp-conflict-free LDS.128 per thread and
//                used as scalar-broadcast FFMA2 operands (negation folded into the operand),
//   YR/YI[k][r][64]  split real / imaginary rows, read as one LDS.128 each and used as packed
//                pairs: accumulators pair two neighbouring outputs (re_j0, re_j1), (im_j0, im_j1).
// Forward:  X = E_i (conjugated in the product), Y = A_s E_j, reduction over sources.
// Backward: X = H[a, m] (Hermitian cotangent matrix), Y = E_m over 64 sources, reduction over
//           partner antennas m:  y_a = sum_m H[a, m] E_m, then with p = conj(E_a) y_a
//           dA[s, k] = 1/2 sum_a Re p   and   dr_a = sum_{s,k} shat_s A[s,k] (2 pi sgn nu_k / c) Im p.
//
// Replaces, like fringe_kernels.cu, telescope_model.py:310-358 + rime_model.py:426-429 and their
// autograd backward.  Every output has one owner and a fixed summation order.
#include "rime_math.cuh"
#include "internal.h"

namespace b200rime {

constexpr int ANT_TILE = 64;       // antennas per tile side
constexpr int ANT_KG = 4;          // channels per pass
constexpr int ANT_ST = 8;          // reduction indices per shared-memory stage
constexpr int ANT_THREADS = 256;
constexpr int ANT_KC = B200_KC_F32;

struct AntSmem {
    static constexpr int X_BYTES = ANT_KG * ANT_ST * ANT_TILE * 8;
    static constexpr int Y_BYTES = ANT_KG * ANT_ST * ANT_TILE * 4;
    static constexpr int STAGE_BYTES = X_BYTES + 2 * Y_BYTES;
    static constexpr int TOTAL = 2 * STAGE_BYTES;
};

// position of antenna slot a (0..63) inside an X row of 64 complex numbers: the four slots of
// thread-row ti = a / 4 are split into two 16-byte halves, each half contiguous over ti, so that
// the 8 distinct ti of a warp read 128 contiguous bytes
__device__ __forceinline__ int xpos(int a) {
    return (((a >> 1) & 1) << 5) | ((a >> 2) << 1) | (a & 1);
}

// acc += conj(x) * y (CONJ) or x * y, for 4 x-values (scalar broadcast) times 4 y-values
// (two packed pairs); aR / aI hold (re, re) / (im, im) of the pairs [i][jp]
template <bool CONJ>
__device__ __forceinline__ void ant_mac(P2 (&aR)[8], P2 (&aI)[8], const float4 x01,
                                        const float4 x23, const float4 yr, const float4 yi) {
    const P2 yr0 = p2(yr.x, yr.y), yr1 = p2(yr.z, yr.w);
    const P2 yi0 = p2(yi.x, yi.y), yi1 = p2(yi.z, yi.w);
    const float xr[4] = {x01.x, x01.z, x23.x, x23.z};
    const float xi[4] = {x01.y, x01.w, x23.y, x23.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const P2 XR = p2(xr[i], xr[i]);
        const P2 XI = p2(xi[i], xi[i]);
        const P2 NXI = p2(-xi[i], -xi[i]);
        p2_mac(aR[2 * i], XR, yr0);
        p2_mac(aR[2 * i + 1], XR, yr1);
        p2_mac(aI[2 * i], XR, yi0);
        p2_mac(aI[2 * i + 1], XR, yi1);
        if (CONJ) {
            p2_mac(aR[2 * i], XI, yi0);
            p2_mac(aR[2 * i + 1], XI, yi1);
            p2_mac(aI[2 * i], NXI, yr0);
            p2_mac(aI[2 * i + 1], NXI, yr1);
        } else {
            p2_mac(aR[2 * i], NXI, yi0);
            p2_mac(aR[2 * i + 1], NXI, yi1);
            p2_mac(aI[2 * i], XI, yr0);
            p2_mac(aI[2 * i + 1], XI, yr1);
        }
    }
}

// one shared-memory stage of multiply-accumulates: ANT_ST reduction indices x ANT_KG channels
template <bool CONJ>
__device__ __forceinline__ void ant_mac_stage(P2 (&aR)[ANT_KG][8], P2 (&aI)[ANT_KG][8],
                                              const unsigned char* buf, int ti, int tj) {
    const float4* X4 = reinterpret_cast<const float4*>(buf);
    const float4* YR4 = reinterpret_cast<const float4*>(buf + AntSmem::X_BYTES);
    const float4* YI4 = reinterpret_cast<const float4*>(buf + AntSmem::X_BYTES + AntSmem::Y_BYTES);
#pragma unroll
    for (int r = 0; r < ANT_ST; ++r) {
#pragma unroll
        for (int k = 0; k < ANT_KG; ++k) {
            const int row = k * ANT_ST + r;
            const float4 x01 = X4[row * 32 + ti];
            const float4 x23 = X4[row * 32 + 16 + ti];
            const float4 yr = YR4[row * 16 + tj];
            const float4 yi = YI4[row * 16 + tj];
            ant_mac<CONJ>(aR[k], aI[k], x01, x23, yr, yi);
        }
    }
}

__device__ __forceinline__ double dot3(double ax, double ay, double az, const double* __restrict__ s) {
    const double2 s01 = __ldg(reinterpret_cast<const double2*>(s));
    const double s2 = __ldg(s + 2);
    return __fma_rn(ax, s01.x, __fma_rn(ay, s01.y, az * s2));
}

// -------------------------------------------------------------------------------------
// forward.  grid = (ntile, Nfp / 4, nunits), block = 256.
// Thread (ti, tj) of the 16 x 16 thread grid owns antenna slots 4 ti .. 4 ti + 3 of the tile's
// X set and 4 tj .. 4 tj + 3 of its Y set.  tile_bl[tile][x][y] = (baseline << 1 | conj) or -1.
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ANT_THREADS, 1)
ant_fringe_fwd_kernel(const float* __restrict__ A, const double* __restrict__ shat,
                      const double* __restrict__ antv, const double* __restrict__ freqs,
                      const int4* __restrict__ units, const int* __restrict__ tile_ant,
                      const int* __restrict__ tile_bl, int nbl, int nfreq, long long S,
                      double sgn_over_c, float* __restrict__ vpart) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int k0 = blockIdx.y * ANT_KG;
    const int nfp = gridDim.y * ANT_KG;
    const int4 un = units[blockIdx.z];
    const int nst = (un.z - un.y) / ANT_ST;

    // multiply-accumulate role
    const int ti = ((warp >> 2) << 3) | (lane & 7);
    const int tj = ((warp & 3) << 2) | (lane >> 3);
    const int* tb = tile_bl + (size_t)tile * (ANT_TILE * ANT_TILE) + (4 * ti) * ANT_TILE + 4 * tj;
    bool mine = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int4 e = __ldg(reinterpret_cast<const int4*>(tb + i * ANT_TILE));
        mine |= (e.x >= 0) | (e.y >= 0) | (e.z >= 0) | (e.w >= 0);
    }
    const bool active = __any_sync(0xffffffffu, mine);   // warps with no wanted pair only generate

    // generation role: one antenna slot (X: 0..63, Y: 64..127), sources (tid >> 7) + 2 r
    const int slot = tid & 127;
    const bool is_y = slot >= ANT_TILE;
    const int ant = __ldg(tile_ant + tile * (2 * ANT_TILE) + slot);
    double ax = 0.0, ay = 0.0, az = 0.0;
    if (ant >= 0) {
        ax = antv[4 * (size_t)ant];
        ay = antv[4 * (size_t)ant + 1];
        az = antv[4 * (size_t)ant + 2];
    }
    double kf[ANT_KG];
#pragma unroll
    for (int k = 0; k < ANT_KG; ++k) kf[k] = (k0 + k < nfreq) ? sgn_over_c * freqs[k0 + k] : 0.0;
    const float* Ak = A + (size_t)(k0 / ANT_KC) * (size_t)S * ANT_KC + (k0 % ANT_KC);
    const int sl0 = tid >> 7;
    const int xw = is_y ? (slot - ANT_TILE) : xpos(slot);

    auto generate = [&](int it, unsigned char* buf) {
        float2* X2 = reinterpret_cast<float2*>(buf);
        float* YR = reinterpret_cast<float*>(buf + AntSmem::X_BYTES);
        float* YI = reinterpret_cast<float*>(buf + AntSmem::X_BYTES + AntSmem::Y_BYTES);
        const long long sbase = (long long)un.y + (long long)it * ANT_ST;
#pragma unroll
        for (int r = 0; r < ANT_ST / 2; ++r) {
            const int sl = sl0 + 2 * r;
            const long long s = sbase + sl;
            const double u = dot3(ax, ay, az, shat + 4 * s);
            float a4[4] = {1.f, 1.f, 1.f, 1.f};
            if (is_y) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(Ak + s * ANT_KC));
                a4[0] = t.x, a4[1] = t.y, a4[2] = t.z, a4[3] = t.w;
            }
#pragma unroll
            for (int k = 0; k < ANT_KG; ++k) {
                float c, sn;
                cis_fast((float)frac_cycles(u, kf[k]), c, sn);
                const int row = (k * ANT_ST + sl) * ANT_TILE;
                if (is_y) {
                    YR[row + xw] = a4[k] * c;
                    YI[row + xw] = a4[k] * sn;
                } else {
                    X2[row + xw] = make_float2(c, sn);
                }
            }
        }
    };

    P2 aR[ANT_KG][8], aI[ANT_KG][8];
#pragma unroll
    for (int k = 0; k < ANT_KG; ++k)
#pragma unroll
        for (int q = 0; q < 8; ++q) aR[k][q] = aI[k][q] = p2(0.f, 0.f);

    if (nst > 0) generate(0, smem);
    __syncthreads();
    for (int it = 0; it < nst; ++it) {
        unsigned char* cur = smem + (it & 1) * AntSmem::STAGE_BYTES;
        unsigned char* nxt = smem + ((it + 1) & 1) * AntSmem::STAGE_BYTES;
        if (it + 1 < nst) generate(it + 1, nxt);
        if (active) ant_mac_stage<true>(aR, aI, cur, ti, tj);
        __syncthreads();
    }

    if (!mine) return;
    float* vp = vpart + (size_t)blockIdx.z * (size_t)nbl * nfp * 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int4 e4 = __ldg(reinterpret_cast<const int4*>(tb + i * ANT_TILE));
        const int e[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (e[j] < 0) continue;
            const float sg = (e[j] & 1) ? -1.f : 1.f;
            float re[ANT_KG], im[ANT_KG];
#pragma unroll
            for (int k = 0; k < ANT_KG; ++k) {
                float r0, r1, i0, i1;
                p2_get(aR[k][2 * i + (j >> 1)], r0, r1);
                p2_get(aI[k][2 * i + (j >> 1)], i0, i1);
                re[k] = (j & 1) ? r1 : r0;
                im[k] = sg * ((j & 1) ? i1 : i0);
            }
            float4* dst = reinterpret_cast<float4*>(vp + ((size_t)(e[j] >> 1) * nfp + k0) * 2);
            dst[0] = make_float4(re[0], im[0], re[1], im[1]);
            dst[1] = make_float4(re[2], im[2], re[3], im[3]);
        }
    }
}

// -------------------------------------------------------------------------------------
// backward.  grid = (nblk, Nfp / 4, nunits), block = 256.
// The CTA owns antenna block ib (64 output antennas a), 4 channels and a unit of sources; it
// walks the unit in tiles of 64 sources and, per tile, reduces over all partner antennas m in
// stages of ANT_ST.  Thread (ti, tj): antennas 4 ti .. +3 (X = H[a, m], TMA-staged from the
// pre-arranged Hermitian cotangent), sources 4 tj .. +3 (Y = E_m, generated).
//   Hp[t][kg][ib][mstage][k][r][64 a (xpos order)] complex64
//   dApart[ib][chunk][S][KC]            (summed over ib by the caller)
//   drpart[unit][kg][ib * 64 + a][4]    float64 (summed by the caller)
// -------------------------------------------------------------------------------------
struct AntBwdSmem {
    static constexpr int STAGE_BYTES = AntSmem::STAGE_BYTES;
    static constexpr int RED_OFF = 2 * STAGE_BYTES;            // cross-warp reduction scratch
    static constexpr int RED_BYTES = 8 * 32 * 16 * 4;          // 16 KB
    static constexpr int BAR_OFF = RED_OFF + RED_BYTES;
    static constexpr int TOTAL = BAR_OFF + 16;
};

__global__ void __launch_bounds__(ANT_THREADS, 1)
ant_fringe_bwd_kernel(const float* __restrict__ Hp, const float* __restrict__ A,
                      const double* __restrict__ shat, const double* __restrict__ antv,
                      const double* __restrict__ freqs, const int4* __restrict__ units, int na_pad,
                      int nfreq, long long S, double sgn_over_c, int need_a, int need_r,
                      float* __restrict__ dApart, double* __restrict__ drpart) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AntBwdSmem::BAR_OFF);
    float* red = reinterpret_cast<float*>(smem + AntBwdSmem::RED_OFF);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ib = blockIdx.x, nblk = gridDim.x;
    const int kg = blockIdx.y, nkg = gridDim.y;
    const int k0 = kg * ANT_KG;
    const int4 un = units[blockIdx.z];
    const int nmst = na_pad / ANT_ST;                  // partner-antenna stages per source tile
    const int nsrc_tiles = (un.z - un.y) / ANT_TILE;
    const long long total = (long long)nsrc_tiles * nmst;

    const int ti = ((warp >> 2) << 3) | (lane & 7);    // antennas 4 ti ..
    const int tj = ((warp & 3) << 2) | (lane >> 3);    // sources  4 tj ..

    double kf[ANT_KG];
#pragma unroll
    for (int k = 0; k < ANT_KG; ++k) kf[k] = (k0 + k < nfreq) ? sgn_over_c * freqs[k0 + k] : 0.0;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // H tiles of this (time, channel group, antenna block): nmst consecutive 16 KB blocks
    const float* Hbase = Hp + ((((size_t)un.x * nkg + kg) * nblk + ib) * (size_t)nmst) *
                                  (AntSmem::X_BYTES / 4);
    auto issue = [&](long long g, int stage) {
        const int ms = (int)(g % nmst);
        mbar_expect_tx(&bars[stage], AntSmem::X_BYTES);
        bulk_g2s(smem + stage * AntSmem::STAGE_BYTES, Hbase + (size_t)ms * (AntSmem::X_BYTES / 4),
                 AntSmem::X_BYTES, &bars[stage]);
    };

    // generation role: partner antenna m = ms * ANT_ST + (tid >> 6) + 4 r, source tid & 63
    const int gs = tid & 63;
    const int gm0 = tid >> 6;
    auto generate = [&](long long g, unsigned char* buf) {
        float* YR = reinterpret_cast<float*>(buf + AntSmem::X_BYTES);
        float* YI = reinterpret_cast<float*>(buf + AntSmem::X_BYTES + AntSmem::Y_BYTES);
        const int ms = (int)(g % nmst);
        const long long s = (long long)un.y + (g / nmst) * ANT_TILE + gs;
        const double2 s01 = __ldg(reinterpret_cast<const double2*>(shat + 4 * s));
        const double s2 = __ldg(shat + 4 * s + 2);
#pragma unroll
        for (int r = 0; r < ANT_ST / 4; ++r) {
            const int ml = gm0 + 4 * r;
            const double* ap = antv + 4 * (size_t)(ms * ANT_ST + ml);
            const double2 a01 = __ldg(reinterpret_cast<const double2*>(ap));
            const double a2 = __ldg(ap + 2);
            const double u = __fma_rn(a01.x, s01.x, __fma_rn(a01.y, s01.y, a2 * s2));
#pragma unroll
            for (int k = 0; k < ANT_KG; ++k) {
                float c, sn;
                cis_fast((float)frac_cycles(u, kf[k]), c, sn);
                const int row = (k * ANT_ST + ml) * ANT_TILE;
                YR[row + gs] = c;
                YI[row + gs] = sn;
            }
        }
    };

    // own antennas: positions, and gradient accumulators over the whole unit
    double pax[4], pay[4], paz[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double* ap = antv + 4 * (size_t)(ib * ANT_TILE + 4 * ti + i);
        pax[i] = ap[0], pay[i] = ap[1], paz[i] = ap[2];
    }
    float gx[4], gy[4], gz[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) gx[i] = gy[i] = gz[i] = 0.f;
    const float* Ak = A + (size_t)(k0 / ANT_KC) * (size_t)S * ANT_KC + (k0 % ANT_KC);
    float* dAk = dApart + ((size_t)ib * gridDim.y * ANT_KG / ANT_KC + (k0 / ANT_KC)) * (size_t)S * ANT_KC +
                 (k0 % ANT_KC);

    P2 aR[ANT_KG][8], aI[ANT_KG][8];

    if (total > 0) {
        if (tid == 0) issue(0, 0);
        generate(0, smem);
    }
    __syncthreads();
    long long g = 0;
    for (int st = 0; st < nsrc_tiles; ++st) {
#pragma unroll
        for (int k = 0; k < ANT_KG; ++k)
#pragma unroll
            for (int q = 0; q < 8; ++q) aR[k][q] = aI[k][q] = p2(0.f, 0.f);
        for (int ms = 0; ms < nmst; ++ms, ++g) {
            const int stage = (int)(g & 1);
            unsigned char* cur = smem + stage * AntSmem::STAGE_BYTES;
            if (g + 1 < total) {
                if (tid == 0) issue(g + 1, stage ^ 1);
                generate(g + 1, smem + (stage ^ 1) * AntSmem::STAGE_BYTES);
            }
            mbar_wait(&bars[stage], (uint32_t)((g >> 1) & 1));
            ant_mac_stage<false>(aR, aI, cur, ti, tj);
            __syncthreads();
        }
        // ---- epilogue of this source tile: p = conj(E_a) y_a
        const long long sb = (long long)un.y + (long long)st * ANT_TILE + 4 * tj;
        float dAacc[4][ANT_KG];        // [source][channel], sum over own 4 antennas of Re p
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < ANT_KG; ++k) dAacc[j][k] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long s = sb + j;
            const double2 s01 = __ldg(reinterpret_cast<const double2*>(shat + 4 * s));
            const double s2 = __ldg(shat + 4 * s + 2);
            const float4 a4v = __ldg(reinterpret_cast<const float4*>(Ak + s * ANT_KC));
            const float a4[4] = {a4v.x, a4v.y, a4v.z, a4v.w};
            const float sx = (float)s01.x, sy = (float)s01.y, sz = (float)s2;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double u = __fma_rn(pax[i], s01.x, __fma_rn(pay[i], s01.y, paz[i] * s2));
                float w = 0.f;      // sum_k A kappa_k Im p
#pragma unroll
                for (int k = 0; k < ANT_KG; ++k) {
                    float c, sn, r0, r1, i0, i1;
                    cis_fast((float)frac_cycles(u, kf[k]), c, sn);
                    p2_get(aR[k][2 * i + (j >> 1)], r0, r1);
                    p2_get(aI[k][2 * i + (j >> 1)], i0, i1);
                    const float yr = (j & 1) ? r1 : r0, yi = (j & 1) ? i1 : i0;
                    dAacc[j][k] += c * yr + sn * yi;            // Re(conj(E) y)
                    const float pim = c * yi - sn * yr;         // Im(conj(E) y)
                    w = fmaf(a4[k] * (float)kf[k], pim, w);
                }
                gx[i] = fmaf(w, sx, gx[i]);
                gy[i] = fmaf(w, sy, gy[i]);
                gz[i] = fmaf(w, sz, gz[i]);
            }
        }
        if (need_a) {
            // sum over the 16 thread-rows ti: lanes (bits 0..2) by shuffle, then the two warp
            // halves (warp >> 2) through shared memory, fixed order
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int k = 0; k < ANT_KG; ++k) {
                    float v = dAacc[j][k];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    dAacc[j][k] = v;
                }
            if ((warp >> 2) == 1 && (lane & 7) == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int k = 0; k < ANT_KG; ++k) red[(tj * 4 + j) * ANT_KG + k] = dAacc[j][k];
            }
            __syncthreads();
            if ((warp >> 2) == 0 && (lane & 7) == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float o[ANT_KG];
#pragma unroll
                    for (int k = 0; k < ANT_KG; ++k)
                        o[k] = 0.5f * (dAacc[j][k] + red[(tj * 4 + j) * ANT_KG + k]);
                    *reinterpret_cast<float4*>(dAk + (sb + j) * ANT_KC) =
                        make_float4(o[0], o[1], o[2], o[3]);
                }
            }
            __syncthreads();
        }
    }

    if (need_r) {
        // sum the antenna gradients over the 16 thread-columns tj: lanes (bits 3, 4) by shuffle,
        // then the four warps of a half through shared memory, in float64
        double* redd = reinterpret_cast<double*>(red);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v[3] = {gx[i], gy[i], gz[i]};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                v[c] += __shfl_xor_sync(0xffffffffu, v[c], 8);
                v[c] += __shfl_xor_sync(0xffffffffu, v[c], 16);
            }
            if ((lane >> 3) == 0) {
                double* dst = redd + (((warp & 3) * 64 + 4 * ti + i) * 4);
                dst[0] = v[0], dst[1] = v[1], dst[2] = v[2];
            }
        }
        __syncthreads();
        if (tid < ANT_TILE) {
            double o[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int c = 0; c < 3; ++c) o[c] += redd[(q * 64 + tid) * 4 + c];
            const double twopi = 6.283185307179586476925286766559;
            double* dst = drpart + ((((size_t)blockIdx.z * nkg + kg) * nblk + ib) * ANT_TILE + tid) * 4;
            dst[0] = twopi * o[0], dst[1] = twopi * o[1], dst[2] = twopi * o[2], dst[3] = 0.0;
        }
    }
}

// -------------------------------------------------------------------------------------
// launchers
// -------------------------------------------------------------------------------------
int launch_ant_fwd(const float* A, const double* shat, const double* antv, const double* freqs,
                   const int* units, int nunits, const int* tile_ant, const int* tile_bl, int ntile,
                   int nbl, int nfreq, long long S, int conj, float* vpart, cudaStream_t st) {
    if (nunits <= 0 || ntile <= 0 || nfreq <= 0) return 0;
    if (S % SRC_PAD) return set_error("antfringe_fwd: S must be a multiple of 128");
    const int nfp = ((nfreq + ANT_KC - 1) / ANT_KC) * ANT_KC;
    if (nfp / ANT_KG > 65535 || nunits > 65535) return set_error("antfringe_fwd: grid too large");
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(ant_fringe_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             AntSmem::TOTAL);
        attr_set = true;
    }
    dim3 grid(ntile, nfp / ANT_KG, nunits);
    ant_fringe_fwd_kernel<<<grid, ANT_THREADS, AntSmem::TOTAL, st>>>(
        A, shat, antv, freqs, reinterpret_cast<const int4*>(units), tile_ant, tile_bl, nbl, nfreq, S,
        (conj ? -1.0 : 1.0) / C_LIGHT, vpart);
    return check_launch("antfringe_fwd");
}

int launch_ant_bwd(const float* Hp, const float* A, const double* shat, const double* antv,
                   const double* freqs, const int* units, int nunits, int na_pad, int nfreq,
                   long long S, int conj, float* dApart, double* drpart, cudaStream_t st) {
    if (nunits <= 0 || na_pad <= 0 || nfreq <= 0) return 0;
    if (S % SRC_PAD) return set_error("antfringe_bwd: S must be a multiple of 128");
    if (na_pad % ANT_TILE) return set_error("antfringe_bwd: antenna count must be padded to 64");
    const int nfp = ((nfreq + ANT_KC - 1) / ANT_KC) * ANT_KC;
    if (nfp / ANT_KG > 65535 || nunits > 65535) return set_error("antfringe_bwd: grid too large");
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(ant_fringe_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             AntBwdSmem::TOTAL);
        attr_set = true;
    }
    dim3 grid(na_pad / ANT_TILE, nfp / ANT_KG, nunits);
    ant_fringe_bwd_kernel<<<grid, ANT_THREADS, AntBwdSmem::TOTAL, st>>>(
        Hp, A, shat, antv, freqs, reinterpret_cast<const int4*>(units), na_pad, nfreq, S,
        (conj ? -1.0 : 1.0) / C_LIGHT, dApart != nullptr, drpart != nullptr, dApart, drpart);
    return check_launch("antfringe_bwd");
}

}  // namespace b200rime

extern "C" {

int b200rime_antfringe_fwd_f32(const float* A, const double* shat, const double* antv,
                               const double* freqs, const int* units, int nunits,
                               const int* tile_ant, const int* tile_bl, int ntile, int nbl,
                               int nfreq, long long S, int conj, float* Vpart, void* stream) {
    return b200rime::launch_ant_fwd(A, shat, antv, freqs, units, nunits, tile_ant, tile_bl, ntile,
                                    nbl, nfreq, S, conj, Vpart, (cudaStream_t)stream);
}
int b200rime_antfringe_bwd_f32(const float* Hp, const float* A, const double* shat,
                               const double* antv, const double* freqs, const int* units,
                               int nunits, int na_pad, int nfreq, long long S, int conj,
                               float* dApart, double* drpart, void* stream) {
    return b200rime::launch_ant_bwd(Hp, A, shat, antv, freqs, units, nunits, na_pad, nfreq, S, conj,
                                    dApart, drpart, (cudaStream_t)stream);
}
int b200rime_ant_tile(void) { return b200rime::ANT_TILE; }
int b200rime_ant_kg(void) { return b200rime::ANT_KG; }
int b200rime_ant_stage(void) { return b200rime::ANT_ST; }

}  // extern "C"
