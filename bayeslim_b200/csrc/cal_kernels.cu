// Gain application V_out = g_1 V g_2^H per baseline (reference calibration._apply_cal,
// calibration.py:2412-2487, the step that follows the RIME in a BayesLIM Sequential; SURVEY
// section 8(f) row f3) and its adjoint to the gains.  HBM bound: one read of the visibilities and
// one write per element, the (Nant, Nt, Nf) gain table stays in L2.
//
//   vis / out  [npol][npol][nbl][nt][nf] complex          (nf contiguous)
//   gains      [npol][npol][nant][ntg][nfg] complex, ntg in {1, nt}, nfg in {1, nf} (broadcast)
//   g1, g2     [nbl] row of the first / second antenna of every baseline in the gain table
//   mode 0 ("diag": 1pol, and 2pol = 4pol data with cal_2pol): out[p][p] = g1[p][p] v[p][p]
//           conj(g2[p][p]), off-diagonal outputs zero (linalg.diag_matmul, linalg.py:116-149)
//   mode 1 ("full", npol = 2): out[a][d] = sum_{b,c} g1[a][b] v[b][c] conj(g2[d][c])
//           (calibration.py:2485)
//   cov / cov_out (diag mode, optional): cov_out[p][p] = |g1 conj(g2)|^2 cov[p][p] (:2470-2476)
#include "common.cuh"
#include "internal.h"

namespace b200rime {

constexpr int CAL_THREADS = 128;

template <typename T> struct CalC { T re, im; };
template <typename T> __device__ __forceinline__ CalC<T> cmul(CalC<T> a, CalC<T> b) {
    return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
template <typename T> __device__ __forceinline__ CalC<T> cmulc(CalC<T> a, CalC<T> b) {   // a conj(b)
    return {a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im};
}
template <typename T> __device__ __forceinline__ CalC<T> cconj(CalC<T> a) { return {a.re, -a.im}; }
template <typename T> __device__ __forceinline__ void cacc(CalC<T>& s, CalC<T> a) {
    s.re += a.re;
    s.im += a.im;
}

template <typename T, int NPOL, bool FULL>
__global__ void __launch_bounds__(CAL_THREADS)
apply_cal_kernel(const CalC<T>* __restrict__ vis, const CalC<T>* __restrict__ gains,
                 const int* __restrict__ g1, const int* __restrict__ g2, int nbl, int nt, int nf,
                 int nant, int ntg, int nfg, const T* __restrict__ cov, CalC<T>* __restrict__ out,
                 T* __restrict__ cov_out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const int fg = nfg == 1 ? 0 : f;
    const size_t vplane = (size_t)nbl * nt * nf;
    const size_t gplane = (size_t)nant * ntg * nfg;
    for (long long row = blockIdx.y; row < (long long)nbl * nt; row += gridDim.y) {
        const int b = (int)(row / nt), t = (int)(row % nt);
        const int tg = ntg == 1 ? 0 : t;
        const size_t vo = ((size_t)b * nt + t) * nf + f;
        const size_t o1 = ((size_t)g1[b] * ntg + tg) * nfg + fg;
        const size_t o2 = ((size_t)g2[b] * ntg + tg) * nfg + fg;
        if (!FULL) {
#pragma unroll
            for (int p = 0; p < NPOL; ++p) {
                const size_t pp = (size_t)(p * NPOL + p);
                const CalC<T> G = cmulc(gains[pp * gplane + o1], gains[pp * gplane + o2]);
                out[pp * vplane + vo] = cmul(G, vis[pp * vplane + vo]);
                if (cov != nullptr)
                    cov_out[pp * vplane + vo] = (G.re * G.re + G.im * G.im) * cov[pp * vplane + vo];
            }
            if (NPOL == 2) {
                out[1 * vplane + vo] = {0, 0};
                out[2 * vplane + vo] = {0, 0};
                if (cov != nullptr) cov_out[1 * vplane + vo] = cov_out[2 * vplane + vo] = 0;
            }
        } else {
            CalC<T> v[2][2], a[2][2], c[2][2];
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    v[p][q] = vis[(size_t)(p * 2 + q) * vplane + vo];
                    a[p][q] = gains[(size_t)(p * 2 + q) * gplane + o1];
                    c[p][q] = gains[(size_t)(p * 2 + q) * gplane + o2];
                }
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                CalC<T> tmp[2];                      // (g1 v)[x][c]
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    tmp[cc] = cmul(a[x][0], v[0][cc]);
                    cacc(tmp[cc], cmul(a[x][1], v[1][cc]));
                }
#pragma unroll
                for (int d = 0; d < 2; ++d) {
                    CalC<T> s = cmulc(tmp[0], c[d][0]);
                    cacc(s, cmulc(tmp[1], c[d][1]));
                    out[(size_t)(x * 2 + d) * vplane + vo] = s;
                }
            }
        }
    }
}

// diagonal modes, two channels per thread (16-byte accesses for complex64) and two rows in flight:
// a streaming kernel needs tens of KB of loads outstanding per SM to approach the HBM rate
template <typename T> struct __align__(2 * sizeof(CalC<T>)) CalC2 { CalC<T> a, b; };
template <typename T> struct __align__(2 * sizeof(T)) CalR2 { T a, b; };

template <typename T, int NPOL>
__global__ void __launch_bounds__(CAL_THREADS)
apply_cal_diag2_kernel(const CalC<T>* __restrict__ vis, const CalC<T>* __restrict__ gains,
                       const int* __restrict__ g1, const int* __restrict__ g2, int nbl, int nt,
                       int nf, int nant, int ntg, int nfg, const T* __restrict__ cov,
                       CalC<T>* __restrict__ out, T* __restrict__ cov_out) {
    const int f = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (f >= nf) return;
    const size_t vplane = (size_t)nbl * nt * nf;
    const size_t gplane = (size_t)nant * ntg * nfg;
    const long long rows = (long long)nbl * nt;
#pragma unroll 2
    for (long long row = blockIdx.y; row < rows; row += gridDim.y) {
        const int b = (int)(row / nt), t = (int)(row % nt);
        const int tg = ntg == 1 ? 0 : t;
        const size_t vo = ((size_t)b * nt + t) * nf + f;
        const size_t r1 = ((size_t)g1[b] * ntg + tg) * nfg, r2 = ((size_t)g2[b] * ntg + tg) * nfg;
#pragma unroll
        for (int p = 0; p < NPOL; ++p) {
            const size_t pp = (size_t)(p * NPOL + p);
            CalC<T> ga, gb, ha, hb;
            if (nfg == 1) {
                ga = gb = gains[pp * gplane + r1];
                ha = hb = gains[pp * gplane + r2];
            } else {
                const CalC2<T> x = *reinterpret_cast<const CalC2<T>*>(gains + pp * gplane + r1 + f);
                const CalC2<T> y = *reinterpret_cast<const CalC2<T>*>(gains + pp * gplane + r2 + f);
                ga = x.a, gb = x.b, ha = y.a, hb = y.b;
            }
            const CalC<T> Ga = cmulc(ga, ha), Gb = cmulc(gb, hb);
            const CalC2<T> v = *reinterpret_cast<const CalC2<T>*>(vis + pp * vplane + vo);
            CalC2<T> o;
            o.a = cmul(Ga, v.a);
            o.b = cmul(Gb, v.b);
            *reinterpret_cast<CalC2<T>*>(out + pp * vplane + vo) = o;
            if (cov != nullptr) {
                const CalR2<T> c = *reinterpret_cast<const CalR2<T>*>(cov + pp * vplane + vo);
                CalR2<T> co;
                co.a = (Ga.re * Ga.re + Ga.im * Ga.im) * c.a;
                co.b = (Gb.re * Gb.re + Gb.im * Gb.im) * c.b;
                *reinterpret_cast<CalR2<T>*>(cov_out + pp * vplane + vo) = co;
            }
        }
        if (NPOL == 2) {
            CalC2<T> z;
            z.a = z.b = {0, 0};
            *reinterpret_cast<CalC2<T>*>(out + 1 * vplane + vo) = z;
            *reinterpret_cast<CalC2<T>*>(out + 2 * vplane + vo) = z;
            if (cov != nullptr) {
                CalR2<T> zr;
                zr.a = zr.b = 0;
                *reinterpret_cast<CalR2<T>*>(cov_out + 1 * vplane + vo) = zr;
                *reinterpret_cast<CalR2<T>*>(cov_out + 2 * vplane + vo) = zr;
            }
        }
    }
}

// adjoint to the gains: one thread per (antenna, t, f) walks the antenna's baselines in a fixed
// order (CSR lists built by the caller: baselines where it is the first / the second antenna).
//   dg[p][q][ant][t][f] (full time / frequency axes; the caller sums broadcast axes)
template <typename T, int NPOL, bool FULL>
__global__ void __launch_bounds__(CAL_THREADS)
apply_cal_bwd_gains_kernel(const CalC<T>* __restrict__ vis, const CalC<T>* __restrict__ gains,
                           const CalC<T>* __restrict__ gout, const int* __restrict__ g1,
                           const int* __restrict__ g2, const int* __restrict__ p1,
                           const int* __restrict__ b1, const int* __restrict__ p2,
                           const int* __restrict__ b2, int nbl, int nt, int nf, int nant, int ntg,
                           int nfg, CalC<T>* __restrict__ dg) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const int fg = nfg == 1 ? 0 : f;
    const size_t vplane = (size_t)nbl * nt * nf;
    const size_t gplane = (size_t)nant * ntg * nfg;
    const size_t dplane = (size_t)nant * nt * nf;
    for (long long row = blockIdx.y; row < (long long)nant * nt; row += gridDim.y) {
        const int ant = (int)(row / nt), t = (int)(row % nt);
        const int tg = ntg == 1 ? 0 : t;
        CalC<T> acc[NPOL][NPOL];
#pragma unroll
        for (int p = 0; p < NPOL; ++p)
#pragma unroll
            for (int q = 0; q < NPOL; ++q) acc[p][q] = {0, 0};
        // baselines with `ant` first: out = g_ant (v g_o^H)  =>  dg_ant += G (v g_o^H)^H
        for (int e = p1[ant]; e < p1[ant + 1]; ++e) {
            const int b = b1[e];
            const size_t vo = ((size_t)b * nt + t) * nf + f;
            const size_t oo = ((size_t)g2[b] * ntg + tg) * nfg + fg;
            if (!FULL) {
#pragma unroll
                for (int p = 0; p < NPOL; ++p) {
                    const size_t pp = (size_t)(p * NPOL + p);
                    // G conj(v conj(g_o)) = G conj(v) g_o
                    cacc(acc[p][p], cmul(cmulc(gout[pp * vplane + vo], vis[pp * vplane + vo]),
                                         gains[pp * gplane + oo]));
                }
            } else {
#pragma unroll
                for (int y = 0; y < 2; ++y) {
                    // w[d] = sum_c conj(v[y][c]) g_o[d][c]
                    CalC<T> w[2];
#pragma unroll
                    for (int d = 0; d < 2; ++d) {
                        w[d] = cmulc(gains[(size_t)(d * 2 + 0) * gplane + oo],
                                     vis[(size_t)(y * 2 + 0) * vplane + vo]);
                        cacc(w[d], cmulc(gains[(size_t)(d * 2 + 1) * gplane + oo],
                                         vis[(size_t)(y * 2 + 1) * vplane + vo]));
                    }
#pragma unroll
                    for (int x = 0; x < 2; ++x) {
                        cacc(acc[x][y], cmul(gout[(size_t)(x * 2 + 0) * vplane + vo], w[0]));
                        cacc(acc[x][y], cmul(gout[(size_t)(x * 2 + 1) * vplane + vo], w[1]));
                    }
                }
            }
        }
        // baselines with `ant` second: out = (g_o v) g_ant^H  =>  dg_ant += G^H (g_o v)
        for (int e = p2[ant]; e < p2[ant + 1]; ++e) {
            const int b = b2[e];
            const size_t vo = ((size_t)b * nt + t) * nf + f;
            const size_t oo = ((size_t)g1[b] * ntg + tg) * nfg + fg;
            if (!FULL) {
#pragma unroll
                for (int p = 0; p < NPOL; ++p) {
                    const size_t pp = (size_t)(p * NPOL + p);
                    // conj(G) g_o v
                    const CalC<T> u = cmul(gains[pp * gplane + oo], vis[pp * vplane + vo]);
                    cacc(acc[p][p], cmulc(u, gout[pp * vplane + vo]));
                }
            } else {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    // u[x] = sum_y g_o[x][y] v[y][c]
                    CalC<T> u[2];
#pragma unroll
                    for (int x = 0; x < 2; ++x) {
                        u[x] = cmul(gains[(size_t)(x * 2 + 0) * gplane + oo],
                                    vis[(size_t)(0 * 2 + c) * vplane + vo]);
                        cacc(u[x], cmul(gains[(size_t)(x * 2 + 1) * gplane + oo],
                                        vis[(size_t)(1 * 2 + c) * vplane + vo]));
                    }
#pragma unroll
                    for (int d = 0; d < 2; ++d) {
                        cacc(acc[d][c], cmulc(u[0], gout[(size_t)(0 * 2 + d) * vplane + vo]));
                        cacc(acc[d][c], cmulc(u[1], gout[(size_t)(1 * 2 + d) * vplane + vo]));
                    }
                }
            }
        }
        const size_t go = ((size_t)ant * nt + t) * nf + f;
#pragma unroll
        for (int p = 0; p < NPOL; ++p)
#pragma unroll
            for (int q = 0; q < NPOL; ++q) dg[(size_t)(p * NPOL + q) * dplane + go] = acc[p][q];
    }
}

template <typename T>
int launch_apply_cal(const T* vis, const T* gains, const int* g1, const int* g2, int npol, int full,
                     int nbl, int nt, int nf, int nant, int ntg, int nfg, const T* cov, T* out,
                     T* cov_out, cudaStream_t st) {
    if (nbl <= 0 || nt <= 0 || nf <= 0) return 0;
    if (npol != 1 && npol != 2) return set_error("apply_cal: npol must be 1 or 2");
    if (full && npol != 2) return set_error("apply_cal: full mode needs npol = 2");
    if ((ntg != 1 && ntg != nt) || (nfg != 1 && nfg != nf))
        return set_error("apply_cal: gains must match or broadcast the time / frequency axes");
    if (full && cov != nullptr) return set_error("apply_cal: covariance update is diagonal-only");
    const long long rows = (long long)nbl * nt;
    dim3 grid((nf + CAL_THREADS - 1) / CAL_THREADS, (unsigned)(rows < 65535 ? rows : 65535));
    auto V = reinterpret_cast<const CalC<T>*>(vis);
    auto G = reinterpret_cast<const CalC<T>*>(gains);
    auto O = reinterpret_cast<CalC<T>*>(out);
    auto aligned = [](const void* q, size_t a) { return q == nullptr || ((uintptr_t)q % a) == 0; };
    const size_t ca = 2 * sizeof(CalC<T>), ra = 2 * sizeof(T);
    if (!full && nf % 2 == 0 && aligned(vis, ca) && aligned(gains, ca) && aligned(out, ca) &&
        aligned(cov, ra) && aligned(cov_out, ra)) {
        // even channel count and aligned bases: every row starts aligned, two channels per thread
        dim3 grid2((nf / 2 + CAL_THREADS - 1) / CAL_THREADS, grid.y);
        if (npol == 1)
            apply_cal_diag2_kernel<T, 1><<<grid2, CAL_THREADS, 0, st>>>(V, G, g1, g2, nbl, nt, nf, nant,
                                                                        ntg, nfg, cov, O, cov_out);
        else
            apply_cal_diag2_kernel<T, 2><<<grid2, CAL_THREADS, 0, st>>>(V, G, g1, g2, nbl, nt, nf, nant,
                                                                        ntg, nfg, cov, O, cov_out);
    } else if (npol == 1)
        apply_cal_kernel<T, 1, false><<<grid, CAL_THREADS, 0, st>>>(V, G, g1, g2, nbl, nt, nf, nant,
                                                                    ntg, nfg, cov, O, cov_out);
    else if (!full)
        apply_cal_kernel<T, 2, false><<<grid, CAL_THREADS, 0, st>>>(V, G, g1, g2, nbl, nt, nf, nant,
                                                                    ntg, nfg, cov, O, cov_out);
    else
        apply_cal_kernel<T, 2, true><<<grid, CAL_THREADS, 0, st>>>(V, G, g1, g2, nbl, nt, nf, nant,
                                                                   ntg, nfg, cov, O, cov_out);
    return check_launch("apply_cal");
}

template <typename T>
int launch_apply_cal_bwd_gains(const T* vis, const T* gains, const T* gout, const int* g1,
                               const int* g2, const int* p1, const int* b1, const int* p2,
                               const int* b2, int npol, int full, int nbl, int nt, int nf, int nant,
                               int ntg, int nfg, T* dg, cudaStream_t st) {
    if (nant <= 0 || nt <= 0 || nf <= 0) return 0;
    if (npol != 1 && npol != 2) return set_error("apply_cal_bwd_gains: npol must be 1 or 2");
    if (full && npol != 2) return set_error("apply_cal_bwd_gains: full mode needs npol = 2");
    const long long rows = (long long)nant * nt;
    dim3 grid((nf + CAL_THREADS - 1) / CAL_THREADS, (unsigned)(rows < 65535 ? rows : 65535));
    auto V = reinterpret_cast<const CalC<T>*>(vis);
    auto G = reinterpret_cast<const CalC<T>*>(gains);
    auto O = reinterpret_cast<const CalC<T>*>(gout);
    auto D = reinterpret_cast<CalC<T>*>(dg);
    if (npol == 1)
        apply_cal_bwd_gains_kernel<T, 1, false><<<grid, CAL_THREADS, 0, st>>>(
            V, G, O, g1, g2, p1, b1, p2, b2, nbl, nt, nf, nant, ntg, nfg, D);
    else if (!full)
        apply_cal_bwd_gains_kernel<T, 2, false><<<grid, CAL_THREADS, 0, st>>>(
            V, G, O, g1, g2, p1, b1, p2, b2, nbl, nt, nf, nant, ntg, nfg, D);
    else
        apply_cal_bwd_gains_kernel<T, 2, true><<<grid, CAL_THREADS, 0, st>>>(
            V, G, O, g1, g2, p1, b1, p2, b2, nbl, nt, nf, nant, ntg, nfg, D);
    return check_launch("apply_cal_bwd_gains");
}

}  // namespace b200rime

using namespace b200rime;

extern "C" {

int b200rime_apply_cal_f32(const float* vis, const float* gains, const int* g1, const int* g2,
                           int npol, int full, int nbl, int nt, int nf, int nant, int ntg, int nfg,
                           const float* cov, float* out, float* cov_out, void* stream) {
    return launch_apply_cal<float>(vis, gains, g1, g2, npol, full, nbl, nt, nf, nant, ntg, nfg, cov,
                                   out, cov_out, (cudaStream_t)stream);
}
int b200rime_apply_cal_f64(const double* vis, const double* gains, const int* g1, const int* g2,
                           int npol, int full, int nbl, int nt, int nf, int nant, int ntg, int nfg,
                           const double* cov, double* out, double* cov_out, void* stream) {
    return launch_apply_cal<double>(vis, gains, g1, g2, npol, full, nbl, nt, nf, nant, ntg, nfg, cov,
                                    out, cov_out, (cudaStream_t)stream);
}
int b200rime_apply_cal_bwd_gains_f32(const float* vis, const float* gains, const float* gout,
                                     const int* g1, const int* g2, const int* p1, const int* b1,
                                     const int* p2, const int* b2, int npol, int full, int nbl,
                                     int nt, int nf, int nant, int ntg, int nfg, float* dg,
                                     void* stream) {
    return launch_apply_cal_bwd_gains<float>(vis, gains, gout, g1, g2, p1, b1, p2, b2, npol, full,
                                             nbl, nt, nf, nant, ntg, nfg, dg, (cudaStream_t)stream);
}
int b200rime_apply_cal_bwd_gains_f64(const double* vis, const double* gains, const double* gout,
                                     const int* g1, const int* g2, const int* p1, const int* b1,
                                     const int* p2, const int* b2, int npol, int full, int nbl,
                                     int nt, int nf, int nant, int ntg, int nfg, double* dg,
                                     void* stream) {
    return launch_apply_cal_bwd_gains<double>(vis, gains, gout, g1, g2, p1, b1, p2, b2, npol, full,
                                              nbl, nt, nf, nant, ntg, nfg, dg, (cudaStream_t)stream);
}

}  // extern "C"
