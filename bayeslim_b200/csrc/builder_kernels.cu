// Perceived-sky builders (HBM-bound): layout conversion between row-major (Nf, Ns) planes and the
// tiled source layout A[chunk][S][KC], fused with beam evaluation (bilinear / bipolynomial pixel
// interpolation, Airy disk), the FOV gather of the sky (cut_sky_fov) and the beam*sky product.
//
// All kernels use one 32-source x KC-channel shared-memory tile per CTA as a transpose buffer:
// global reads run along the source axis (the contiguous axis of the sky / beam maps), global
// writes of A run along the channel axis (the contiguous axis of the tiled layout), so both
// sides are coalesced.  Every output element has a single owner; the adjoint of the
// interpolation goes through a CSR transpose instead of atomics (bitwise reproducible).
#include "common.cuh"
#include "internal.h"

namespace b200rime {

constexpr int TS = 32;            // sources per builder tile
constexpr int BUILD_THREADS = 256;

// ---- Bessel J1 as the reference evaluates it ---------------------------------------------
// The reference's Airy beam calls torch.special.bessel_j1 (special.py:535).  torch 2.11's
// kernel (ATen/native/Math.h, bessel_j1_forward) uses the Cephes j1 coefficient sets but
// evaluates the monic denominators RQ and QQ by plain Horner recursion, without the implied
// leading 1, so its float64 result differs from the true J1 by up to 4.7e-7 (5 < x < 8).  Parity
// is with the reference, so this routine restates THAT evaluation (same coefficients, same
// grouping) in float64 for both kernel precisions.  The optional analytic-derivative path
// (full_grad) uses CUDA libm's accurate j0().
__device__ __forceinline__ double horner(const double* c, int n, double z) {
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i) r = r * z + c[i];
    return r;
}
__device__ double j1_ref(double x) {
    const double PP[7] = {7.62125616208173112003e-04, 7.31397056940917570436e-02,
                          1.12719608129684925192e+00, 5.11207951146807644818e+00,
                          8.42404590141772420927e+00, 5.21451598682361504063e+00,
                          1.00000000000000000254e+00};
    const double PQ[7] = {5.71323128072548699714e-04, 6.88455908754495404082e-02,
                          1.10514232634061696926e+00, 5.07386386128601488557e+00,
                          8.39985554327604159757e+00, 5.20982848682361821619e+00,
                          9.99999999999999997461e-01};
    const double QP[8] = {5.10862594750176621635e-02, 4.98213872951233449420e+00,
                          7.58238284132545283818e+01, 3.66779609360150777800e+02,
                          7.10856304998926107277e+02, 5.97489612400613639965e+02,
                          2.11688757100572135698e+02, 2.52070205858023719784e+01};
    const double QQ[7] = {7.42373277035675149943e+01, 1.05644886038262816351e+03,
                          4.98641058337653607651e+03, 9.56231892404756170795e+03,
                          7.99704160447350683650e+03, 2.82619278517639096600e+03,
                          3.36093607810698293419e+02};
    const double RP[4] = {-8.99971225705559398224e+08, 4.52228297998194034323e+11,
                          -7.27494245221818276015e+13, 3.68295732863852883286e+15};
    const double RQ[8] = {6.20836478118054335476e+02, 2.56987256757748830383e+05,
                          8.35146791431949253037e+07, 2.21511595479792499675e+10,
                          4.74914122079991414898e+12, 7.84369607876235854894e+14,
                          8.95222336184627338078e+16, 5.32278620332680085395e+18};
    const double sgn = x < 0.0 ? -1.0 : 1.0;
    x = fabs(x);
    if (x <= 5.0) {
        const double z = x * x;
        const double rp = horner(RP, 4, z), rq = horner(RQ, 8, z);
        return sgn * (rp / rq * x * (z - 1.46819706421238932572e+01) *
                      (z - 4.92184563216946036703e+01));
    }
    const double w = 5.0 / x;
    const double z = w * w;
    const double pp = horner(PP, 7, z), pq = horner(PQ, 7, z);
    const double qp = horner(QP, 8, z), qq = horner(QQ, 7, z);
    const double xn = x - 2.356194490192344928846982537459627163;
    double sn, cs;
    sincos(xn, &sn, &cs);
    return sgn * ((pp / pq * cs - w * (qp / qq) * sn) * 0.797884560802865355879892119868763737 /
                  sqrt(x));
}

// write a [KC][TS+1] smem tile to A[chunk][soff + s0 + srow][k], zero beyond nvalid rows handled
// by the producer (tile already holds zeros there)
template <typename T, int KC>
__device__ __forceinline__ void store_tile(const T (*tile)[TS + 1], T* __restrict__ A, long long S,
                                           long long srow0, int chunk, int rows) {
    T* dst = A + ((size_t)chunk * (size_t)S + (size_t)srow0) * KC;
    for (int i = threadIdx.x; i < rows * KC; i += BUILD_THREADS) {
        const int r = i / KC, k = i - r * KC;
        dst[i] = tile[k][r];
    }
}
template <typename T, int KC>
__device__ __forceinline__ void load_tile(T (*tile)[TS + 1], const T* __restrict__ A, long long S,
                                          long long srow0, int chunk, int rows) {
    const T* src = A + ((size_t)chunk * (size_t)S + (size_t)srow0) * KC;
    for (int i = threadIdx.x; i < rows * KC; i += BUILD_THREADS) {
        const int r = i / KC, k = i - r * KC;
        tile[k][r] = src[i];
    }
}

// ---- pack / unpack -------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BUILD_THREADS)
pack_kernel(const T* __restrict__ X, long long ldx, int nfreq, int ns, long long soff, long long S,
            T* __restrict__ A) {
    constexpr int KC = Cfg<T>::KC;
    __shared__ T tile[KC][TS + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int chunk = blockIdx.y;
    const int s = blockIdx.x * TS + tx;
    for (int k = ty; k < KC; k += BUILD_THREADS / 32) {
        const int f = chunk * KC + k;
        T v = 0;
        if (f < nfreq && s < ns) v = X[(size_t)f * ldx + s];
        tile[k][tx] = v;
    }
    __syncthreads();
    store_tile<T, KC>(tile, A, S, soff + (long long)blockIdx.x * TS, chunk, TS);
}

template <typename T>
__global__ void __launch_bounds__(BUILD_THREADS)
unpack_kernel(const T* __restrict__ A, long long ldx, int nfreq, int ns, long long soff,
              long long S, T* __restrict__ X) {
    constexpr int KC = Cfg<T>::KC;
    __shared__ T tile[KC][TS + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int chunk = blockIdx.y;
    load_tile<T, KC>(tile, A, S, soff + (long long)blockIdx.x * TS, chunk, TS);
    __syncthreads();
    const int s = blockIdx.x * TS + tx;
    for (int k = ty; k < KC; k += BUILD_THREADS / 32) {
        const int f = chunk * KC + k;
        if (f < nfreq && s < ns) X[(size_t)f * ldx + s] = tile[k][tx];
    }
}

// ---- interpolated pixel beam x sky -----------------------------------------------------
template <typename T>
__device__ __forceinline__ T interp_at(const T* __restrict__ brow, const int* __restrict__ inds,
                                       const T* __restrict__ wgts, int nnn, int s) {
    T b = 0;
    const int* ip = inds + (size_t)s * nnn;
    const T* wp = wgts + (size_t)s * nnn;
    for (int i = 0; i < nnn; ++i) b += brow[ip[i]] * wp[i];
    return b;
}

// Neighbour tables of one source held in registers (NNN = 1 or 4: nearest / bilinear, the
// common cases) or re-read through L1 (NNN = 0: any count, e.g. 9 / 16 for bi-quadratic / cubic).
template <typename T, int NNN> struct Nbr {
    int idx[NNN];
    T w[NNN];
    __device__ __forceinline__ void load(const int* __restrict__ inds, const T* __restrict__ wgts,
                                         int nnn, int s) {
#pragma unroll
        for (int i = 0; i < NNN; ++i) {
            idx[i] = inds[(size_t)s * NNN + i];
            w[i] = wgts[(size_t)s * NNN + i];
        }
    }
    __device__ __forceinline__ T at(const T* __restrict__ brow, const int*, const T*, int, int) const {
        T b = 0;
#pragma unroll
        for (int i = 0; i < NNN; ++i) b += brow[idx[i]] * w[i];
        return b;
    }
};
template <typename T> struct Nbr<T, 0> {
    int idx[1];
    T w[1];
    __device__ __forceinline__ void load(const int*, const T*, int, int) {}
    __device__ __forceinline__ T at(const T* __restrict__ brow, const int* __restrict__ inds,
                                    const T* __restrict__ wgts, int nnn, int s) const {
        return interp_at<T>(brow, inds, wgts, nnn, s);
    }
};

template <typename T, int NNN>
__global__ void __launch_bounds__(BUILD_THREADS)
build_interp_kernel(const T* __restrict__ bmap, long long ldb, const int* __restrict__ inds,
                    const T* __restrict__ wgts, int nnn, const T* __restrict__ sky, long long lds,
                    const int* __restrict__ cut, int nfreq, int ns, long long soff, long long S,
                    T* __restrict__ A) {
    constexpr int KC = Cfg<T>::KC;
    constexpr int ROWS = BUILD_THREADS / 32;
    __shared__ T tile[KC][TS + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int chunk = blockIdx.y;
    const int s = blockIdx.x * TS + tx;
    const int pix = (s < ns) ? cut[s] : -1;       // cut < 0 marks a padding entry
    const bool live = pix >= 0;
    Nbr<T, NNN> nb;
    if (live && bmap) nb.load(inds, wgts, nnn, s);
    // all gathers of this thread's KC/ROWS channels are issued before they are consumed
    T v[KC / ROWS];
#pragma unroll
    for (int i = 0; i < KC / ROWS; ++i) {
        const int f = chunk * KC + ty + i * ROWS;
        v[i] = 0;
        if (live && f < nfreq) {
            // bmap == NULL: pure FOV gather of the sky; sky == NULL: pure beam interpolation
            const T b = bmap ? nb.at(bmap + (size_t)f * ldb, inds, wgts, nnn, s) : (T)1;
            const T I = sky ? sky[(size_t)f * lds + pix] : (T)1;
            v[i] = b * I;
        }
    }
#pragma unroll
    for (int i = 0; i < KC / ROWS; ++i) tile[ty + i * ROWS][tx] = v[i];
    __syncthreads();
    store_tile<T, KC>(tile, A, S, soff + (long long)blockIdx.x * TS, chunk, TS);
}

// The same product from a CHANNEL-MAJOR beam map bmapT[Npb][ldt] (ldt >= nchunk * KC, channels
// beyond nfreq zero).  In the pixel-major form above a warp gathers 32 scattered pixels of one
// channel row per load (ncu: 13.7 sectors per request, L1TEX pipe at 87 % of peak, HBM at 30 %);
// here a warp owns one source and its lanes run over channels, so every neighbour read is one
// fully used 128-byte line and the result is written to A without a transpose.  Only the sky tile
// (read coalesced along sources) goes through shared memory.
template <typename T, int NNN>
__global__ void __launch_bounds__(BUILD_THREADS)
build_interp_t_kernel(const T* __restrict__ bmapT, long long ldt, const int* __restrict__ inds,
                      const T* __restrict__ wgts, int nnn, const T* __restrict__ sky, long long lds,
                      const int* __restrict__ cut, int nfreq, int ns, long long soff, long long S,
                      T* __restrict__ A) {
    constexpr int KC = Cfg<T>::KC;
    constexpr int ROWS = BUILD_THREADS / 32;          // warps
    constexpr int SPW = TS / ROWS;                    // sources per warp
    __shared__ T tile[KC][TS + 1];
    __shared__ int spix[TS];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int chunk = blockIdx.y;
    {
        const int s = blockIdx.x * TS + tx;
        const int pix = (s < ns) ? cut[s] : -1;       // cut < 0 marks a padding entry
        if (ty == 0) spix[tx] = pix;
#pragma unroll
        for (int i = 0; i < KC / ROWS; ++i) {
            const int k = ty + i * ROWS, f = chunk * KC + k;
            T I = 0;
            if (pix >= 0 && f < nfreq) I = sky ? sky[(size_t)f * lds + pix] : (T)1;
            tile[k][tx] = I;
        }
    }
    __syncthreads();
    T* dst = A + ((size_t)chunk * (size_t)S + (size_t)soff + (size_t)blockIdx.x * TS) * KC;
    const T* bcol = bmapT + (size_t)chunk * KC;
    if constexpr (sizeof(T) == 4 && NNN == 4 && KC == 64) {
        // float32, bilinear: a HALF-warp owns a source, a lane four channels -- every neighbour
        // read is one 16-byte load (four requests of 256 bytes per source instead of eight of
        // 128), all eight loads of a thread's two sources in flight before the first use
        const int half = tx >> 4, c4 = (tx & 15) * 4;
        float4 v[SPW / 2][4];
        float w[SPW / 2][4];
        bool lv[SPW / 2];
#pragma unroll
        for (int j = 0; j < SPW / 2; ++j) {
            const int sl = ty * SPW + 2 * j + half;
            const int s = blockIdx.x * TS + sl;
            lv[j] = spix[sl] >= 0;
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                v[j][n] = make_float4(0.f, 0.f, 0.f, 0.f);
                w[j][n] = 0.f;
                if (lv[j]) {
                    const int ix = __ldg(inds + (size_t)s * 4 + n);
                    w[j][n] = (float)__ldg(wgts + (size_t)s * 4 + n);
                    v[j][n] = __ldg(reinterpret_cast<const float4*>(
                        reinterpret_cast<const float*>(bcol) + (size_t)ix * ldt + c4));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < SPW / 2; ++j) {
            const int sl = ty * SPW + 2 * j + half;
            float4 b;
            b.x = v[j][0].x * w[j][0] + v[j][1].x * w[j][1] + v[j][2].x * w[j][2] + v[j][3].x * w[j][3];
            b.y = v[j][0].y * w[j][0] + v[j][1].y * w[j][1] + v[j][2].y * w[j][2] + v[j][3].y * w[j][3];
            b.z = v[j][0].z * w[j][0] + v[j][1].z * w[j][1] + v[j][2].z * w[j][2] + v[j][3].z * w[j][3];
            b.w = v[j][0].w * w[j][0] + v[j][1].w * w[j][1] + v[j][2].w * w[j][2] + v[j][3].w * w[j][3];
            b.x *= (float)tile[c4][sl], b.y *= (float)tile[c4 + 1][sl];
            b.z *= (float)tile[c4 + 2][sl], b.w *= (float)tile[c4 + 3][sl];
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + (size_t)sl * KC + c4) = b;
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < SPW; ++j) {
        const int sl = ty * SPW + j;
        const int s = blockIdx.x * TS + sl;
        const bool live = spix[sl] >= 0;
        Nbr<T, NNN> nb;
        if (live) nb.load(inds, wgts, nnn, s);
#pragma unroll
        for (int c = tx; c < KC; c += 32) {
            T b = 0;
            if (live) {
                if (NNN > 0) {
#pragma unroll
                    for (int n = 0; n < (NNN > 0 ? NNN : 1); ++n)
                        b += bcol[(size_t)nb.idx[n] * ldt + c] * nb.w[n];
                } else {
                    for (int n = 0; n < nnn; ++n)
                        b += bcol[(size_t)inds[(size_t)s * nnn + n] * ldt + c] *
                             wgts[(size_t)s * nnn + n];
                }
            }
            dst[(size_t)sl * KC + c] = b * tile[c][sl];
        }
    }
}

template <typename T, int NNN>
__global__ void __launch_bounds__(BUILD_THREADS)
build_interp_bwd_kernel(const T* __restrict__ dA, const T* __restrict__ bmap, long long ldb,
                        const int* __restrict__ inds, const T* __restrict__ wgts, int nnn,
                        const T* __restrict__ sky, long long lds, const int* __restrict__ cut,
                        int nfreq, int ns, long long soff, long long S, T* __restrict__ dsky,
                        T* __restrict__ dBI, long long ldd, T* __restrict__ dIs) {
    constexpr int KC = Cfg<T>::KC;
    constexpr int ROWS = BUILD_THREADS / 32;
    __shared__ T tile[KC][TS + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int chunk = blockIdx.y;
    load_tile<T, KC>(tile, dA, S, soff + (long long)blockIdx.x * TS, chunk, TS);
    __syncthreads();
    const int s = blockIdx.x * TS + tx;
    if (s >= ns) return;
    const int pix = cut[s];
    if (pix < 0) {                                // padding entry: contributes nothing
        for (int k = ty; k < KC; k += ROWS) {
            const int f = chunk * KC + k;
            if (f >= nfreq) break;
            if (dBI) dBI[(size_t)f * ldd + s] = 0;
            if (dIs) dIs[(size_t)f * ldd + s] = 0;
        }
        return;
    }
    Nbr<T, NNN> nb;
    if (bmap) nb.load(inds, wgts, nnn, s);
    T b[KC / ROWS], I[KC / ROWS];
#pragma unroll
    for (int i = 0; i < KC / ROWS; ++i) {
        const int f = chunk * KC + ty + i * ROWS;
        b[i] = 0;
        I[i] = 0;
        if (f < nfreq) {
            b[i] = bmap ? nb.at(bmap + (size_t)f * ldb, inds, wgts, nnn, s) : (T)1;
            I[i] = sky ? sky[(size_t)f * lds + pix] : (T)1;
        }
    }
#pragma unroll
    for (int i = 0; i < KC / ROWS; ++i) {
        const int k = ty + i * ROWS;
        const int f = chunk * KC + k;
        if (f >= nfreq) break;
        const T g = tile[k][tx];
        if (dsky) dsky[(size_t)f * lds + pix] += b[i] * g;
        if (dIs) dIs[(size_t)f * ldd + s] = b[i] * g;
        if (dBI) dBI[(size_t)f * ldd + s] = I[i] * g;
    }
}

// Backward of build_interp_t for float32 bilinear tables from the CHANNEL-MAJOR beam map: like the
// forward, a half-warp owns a source and a lane four channels, so the neighbour reads and the
// cotangent read are 16-byte loads of fully used lines; the sky tile comes in and the two outputs
// (row-major over sources) go out through shared-memory transposes.
//   dIs[f*ldd + s] = B[f][s] dA[f][s]      dBI[f*ldd + s] = sky[f][cut[s]] dA[f][s]
__global__ void __launch_bounds__(BUILD_THREADS)
build_interp_bwd_t_kernel(const float* __restrict__ dA, const float* __restrict__ bmapT,
                          long long ldt, const int* __restrict__ inds,
                          const float* __restrict__ wgts, const float* __restrict__ sky,
                          long long lds, const int* __restrict__ cut, int nfreq, int ns,
                          long long soff, long long S, float* __restrict__ dBI, long long ldd,
                          float* __restrict__ dIs) {
    constexpr int KC = Cfg<float>::KC;
    constexpr int ROWS = BUILD_THREADS / 32, SPW = TS / ROWS;
    static_assert(KC == 64, "a half-warp covers the 64 channels of a chunk");
    __shared__ float tI[KC][TS + 1], o1[KC][TS + 1], o2[KC][TS + 1];
    __shared__ int spix[TS];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int chunk = blockIdx.y;
    {
        const int s = blockIdx.x * TS + tx;
        const int pix = (s < ns) ? cut[s] : -1;
        if (ty == 0) spix[tx] = pix;
#pragma unroll
        for (int i = 0; i < KC / ROWS; ++i) {
            const int k = ty + i * ROWS, f = chunk * KC + k;
            tI[k][tx] = (pix >= 0 && f < nfreq) ? sky[(size_t)f * lds + pix] : 0.f;
        }
    }
    __syncthreads();
    const float* g0 = dA + ((size_t)chunk * (size_t)S + (size_t)soff + (size_t)blockIdx.x * TS) * KC;
    const float* bcol = bmapT + (size_t)chunk * KC;
    const int half = tx >> 4, c4 = (tx & 15) * 4;
#pragma unroll
    for (int j = 0; j < SPW / 2; ++j) {
        const int sl = ty * SPW + 2 * j + half;
        const int s = blockIdx.x * TS + sl;
        const bool live = spix[sl] >= 0;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f), g = b;
        if (live) {
            g = __ldg(reinterpret_cast<const float4*>(g0 + (size_t)sl * KC + c4));
            if (dIs != nullptr) {
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    const int ix = __ldg(inds + (size_t)s * 4 + n);
                    const float w = __ldg(wgts + (size_t)s * 4 + n);
                    const float4 v = __ldg(reinterpret_cast<const float4*>(bcol + (size_t)ix * ldt + c4));
                    b.x += v.x * w, b.y += v.y * w, b.z += v.z * w, b.w += v.w * w;
                }
            }
        }
        o1[c4][sl] = b.x * g.x, o1[c4 + 1][sl] = b.y * g.y;
        o1[c4 + 2][sl] = b.z * g.z, o1[c4 + 3][sl] = b.w * g.w;
        o2[c4][sl] = tI[c4][sl] * g.x, o2[c4 + 1][sl] = tI[c4 + 1][sl] * g.y;
        o2[c4 + 2][sl] = tI[c4 + 2][sl] * g.z, o2[c4 + 3][sl] = tI[c4 + 3][sl] * g.w;
    }
    __syncthreads();
    const int s = blockIdx.x * TS + tx;
    if (s >= ns) return;
#pragma unroll
    for (int i = 0; i < KC / ROWS; ++i) {
        const int k = ty + i * ROWS, f = chunk * KC + k;
        if (f >= nfreq) break;
        if (dIs != nullptr) dIs[(size_t)f * ldd + s] = o1[k][tx];
        if (dBI != nullptr) dBI[(size_t)f * ldd + s] = o2[k][tx];
    }
}

// dbmap[f][p] += sum_j val[j] * dBI[f][col[j]] over the CSR row of pixel p
template <typename T>
__global__ void __launch_bounds__(BUILD_THREADS)
interp_transpose_kernel(const T* __restrict__ dBI, long long ldd, const int* __restrict__ rowptr,
                        const int* __restrict__ col, const T* __restrict__ val, int npix, int nfreq,
                        T* __restrict__ dbmap, long long ldb) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const int j0 = rowptr[p], j1 = rowptr[p + 1];
    if (j0 == j1) return;
    for (int f = blockIdx.y; f < nfreq; f += gridDim.y) {
        const T* row = dBI + (size_t)f * ldd;
        T acc = 0;
        for (int j = j0; j < j1; ++j) acc += val[j] * row[col[j]];
        dbmap[(size_t)f * ldb + p] += acc;
    }
}

// ---- Airy beam x sky -------------------------------------------------------------------
template <typename T> struct AiryArgs {
    double Dew, Dns, kfac;   // kfac = pi * freq_ratio / c
    const double* diam;      // device (Dew, Dns) overriding the two host values, or nullptr
    int square, asym;
    const T* sinzen;
    const T* sin2az;
    const double* freqs;
};

template <typename T>
__device__ __forceinline__ double airy_x(const AiryArgs<T>& a, int s, double nu, double& dxdDew,
                                         double& dxdDns) {
    const double sz = (double)a.sinzen[s];
    const double e = a.asym ? (double)a.sin2az[s] : 1.0;
    const double D = a.asym ? (a.Dns + e * (a.Dew - a.Dns)) : a.Dew;
    const double g = sz * a.kfac * nu;
    dxdDew = e * g;
    dxdDns = a.asym ? (1.0 - e) * g : 0.0;
    return D * g;
}

template <typename T>
__global__ void __launch_bounds__(BUILD_THREADS)
build_airy_kernel(AiryArgs<T> a, const T* __restrict__ sky, long long lds,
                  const int* __restrict__ cut, int nfreq, int ns, long long soff, long long S,
                  T* __restrict__ A, T* __restrict__ Bout, long long ldo) {
    constexpr int KC = Cfg<T>::KC;
    __shared__ T tile[KC][TS + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int chunk = blockIdx.y;
    if (a.diam != nullptr) {
        a.Dew = a.diam[0];
        a.Dns = a.asym ? a.diam[1] : a.Dew;
    }
    const int s = blockIdx.x * TS + tx;
    const int pix = (s < ns) ? cut[s] : -1;
    const bool live = pix >= 0;
    for (int k = ty; k < KC; k += BUILD_THREADS / 32) {
        const int f = chunk * KC + k;
        T v = 0;
        if (live && f < nfreq) {
            double d1, d2;
            double x = airy_x<T>(a, s, a.freqs[f], d1, d2);
            x = fmax(x, 1e-10);
            double hd = 2.0 * j1_ref(x) / x;
            if (a.square) hd = hd * hd;
            const T h = (T)hd;
            if (Bout) Bout[(size_t)f * ldo + s] = h;
            v = h * sky[(size_t)f * lds + pix];
        }
        tile[k][tx] = v;
    }
    __syncthreads();
    if (A) store_tile<T, KC>(tile, A, S, soff + (long long)blockIdx.x * TS, chunk, TS);
}

template <typename T>
__global__ void __launch_bounds__(BUILD_THREADS)
build_airy_bwd_kernel(const T* __restrict__ dA, AiryArgs<T> a, int full_grad,
                      const T* __restrict__ sky, long long lds, const int* __restrict__ cut,
                      int nfreq, int ns, long long soff, long long S, T* __restrict__ dsky,
                      double* __restrict__ dD, T* __restrict__ dIs, long long ldd) {
    constexpr int KC = Cfg<T>::KC;
    __shared__ T tile[KC][TS + 1];
    __shared__ double red[2][BUILD_THREADS / 32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int chunk = blockIdx.y;
    if (a.diam != nullptr) {
        a.Dew = a.diam[0];
        a.Dns = a.asym ? a.diam[1] : a.Dew;
    }
    load_tile<T, KC>(tile, dA, S, soff + (long long)blockIdx.x * TS, chunk, TS);
    __syncthreads();
    const int s = blockIdx.x * TS + tx;
    double gew = 0.0, gns = 0.0;
    const int pix = (s < ns) ? cut[s] : -1;
    if (s < ns && pix < 0 && dIs) {
        for (int k = ty; k < KC; k += BUILD_THREADS / 32) {
            const int f = chunk * KC + k;
            if (f < nfreq) dIs[(size_t)f * ldd + s] = 0;
        }
    }
    if (pix >= 0) {
        for (int k = ty; k < KC; k += BUILD_THREADS / 32) {
            const int f = chunk * KC + k;
            if (f >= nfreq) break;
            const T g = tile[k][tx];
            double d1, d2;
            const double xr = airy_x<T>(a, s, a.freqs[f], d1, d2);
            const bool clipped = xr < 1e-10;
            const double xt = fmax(xr, 1e-10);
            const double J1 = j1_ref(xt);
            const double h = 2.0 * J1 / xt;
            const T B = (T)(a.square ? h * h : h);
            const T I = sky[(size_t)f * lds + pix];
            if (dsky) dsky[(size_t)f * lds + pix] += B * g;
            if (dIs) dIs[(size_t)f * ldd + s] = B * g;
            if (dD && !clipped) {
                // dh/dx: analytic (full) or with J1 held constant (what reference autograd sees)
                const double hp = full_grad ? (2.0 * j0(xt) / xt - 4.0 * J1 / (xt * xt)) : (-h / xt);
                const double dBdx = a.square ? 2.0 * h * hp : hp;
                const double w = (double)I * (double)g * dBdx;
                gew += w * d1;
                gns += w * d2;
            }
        }
    }
    if (dD) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            gew += __shfl_xor_sync(0xffffffffu, gew, o);
            gns += __shfl_xor_sync(0xffffffffu, gns, o);
        }
        if (tx == 0) {
            red[0][ty] = gew;
            red[1][ty] = gns;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a0 = 0.0, a1 = 0.0;
            for (int w = 0; w < BUILD_THREADS / 32; ++w) {
                a0 += red[0][w];
                a1 += red[1][w];
            }
            const size_t blk = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
            dD[2 * blk + 0] = a0;
            dD[2 * blk + 1] = a1;
        }
    }
}

// dsky[f][p] += sum_t dIs[f][pos[t][p]]  (pos < 0: pixel p is outside the FOV at time t).
// One owner per (f, p), times added in index order: the deterministic replacement of the
// per-time index_add of cut_sky_fov's backward when all times are built in one launch.
template <typename T>
__global__ void __launch_bounds__(BUILD_THREADS)
gather_times_kernel(const T* __restrict__ dIs, long long ldd, const int* __restrict__ pos, int nt,
                    int npix, int nfreq, T* __restrict__ dsky, long long lds) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    for (int f = blockIdx.y; f < nfreq; f += gridDim.y) {
        const T* row = dIs + (size_t)f * ldd;
        T acc = 0;
        for (int t = 0; t < nt; ++t) {
            const int j = pos[(size_t)t * npix + p];
            if (j >= 0) acc += row[j];
        }
        dsky[(size_t)f * lds + p] += acc;
    }
}

// ---- launchers -----------------------------------------------------------------------
template <typename T> static inline int nchunks(int nfreq) {
    return (nfreq + Cfg<T>::KC - 1) / Cfg<T>::KC;
}

template <typename T>
int launch_pack(const T* X, long long ldx, int nfreq, int ns, int ns_pad, long long soff,
                long long S, T* A, cudaStream_t st) {
    if (ns_pad <= 0 || nfreq <= 0) return 0;
    if (ns_pad % TS || soff % TS || soff + ns_pad > S) return set_error("pack: bad padding/offset");
    dim3 grid(ns_pad / TS, nchunks<T>(nfreq));
    pack_kernel<T><<<grid, BUILD_THREADS, 0, st>>>(X, ldx, nfreq, ns, soff, S, A);
    return check_launch("pack");
}
template <typename T>
int launch_unpack(const T* A, long long ldx, int nfreq, int ns, long long soff, long long S, T* X,
                  cudaStream_t st) {
    if (ns <= 0 || nfreq <= 0) return 0;
    if (soff % TS) return set_error("unpack: bad offset");
    dim3 grid((ns + TS - 1) / TS, nchunks<T>(nfreq));
    unpack_kernel<T><<<grid, BUILD_THREADS, 0, st>>>(A, ldx, nfreq, ns, soff, S, X);
    return check_launch("unpack");
}
template <typename T>
int launch_build_interp(const T* bmap, long long ldb, const int* inds, const T* wgts, int nnn,
                        const T* sky, long long lds, const int* cut, int nfreq, int ns, int ns_pad,
                        long long soff, long long S, T* A, cudaStream_t st) {
    if (ns_pad <= 0 || nfreq <= 0) return 0;
    if (ns_pad % TS || soff % TS || soff + ns_pad > S)
        return set_error("build_interp: bad padding/offset");
    dim3 grid(ns_pad / TS, nchunks<T>(nfreq));
    if (nnn == 4)
        build_interp_kernel<T, 4><<<grid, BUILD_THREADS, 0, st>>>(bmap, ldb, inds, wgts, nnn, sky,
                                                                  lds, cut, nfreq, ns, soff, S, A);
    else if (nnn == 1)
        build_interp_kernel<T, 1><<<grid, BUILD_THREADS, 0, st>>>(bmap, ldb, inds, wgts, nnn, sky,
                                                                  lds, cut, nfreq, ns, soff, S, A);
    else
        build_interp_kernel<T, 0><<<grid, BUILD_THREADS, 0, st>>>(bmap, ldb, inds, wgts, nnn, sky,
                                                                  lds, cut, nfreq, ns, soff, S, A);
    return check_launch("build_interp");
}
template <typename T>
int launch_build_interp_t(const T* bmapT, long long ldt, const int* inds, const T* wgts, int nnn,
                          const T* sky, long long lds, const int* cut, int nfreq, int ns, int ns_pad,
                          long long soff, long long S, T* A, cudaStream_t st) {
    if (ns_pad <= 0 || nfreq <= 0) return 0;
    if (ns_pad % TS || soff % TS || soff + ns_pad > S)
        return set_error("build_interp_t: bad padding/offset");
    if (bmapT == nullptr || ldt < (long long)nchunks<T>(nfreq) * Cfg<T>::KC)
        return set_error("build_interp_t: needs a channel-major beam map padded to whole chunks");
    dim3 grid(ns_pad / TS, nchunks<T>(nfreq));
    if (nnn == 4)
        build_interp_t_kernel<T, 4><<<grid, BUILD_THREADS, 0, st>>>(bmapT, ldt, inds, wgts, nnn, sky,
                                                                    lds, cut, nfreq, ns, soff, S, A);
    else if (nnn == 1)
        build_interp_t_kernel<T, 1><<<grid, BUILD_THREADS, 0, st>>>(bmapT, ldt, inds, wgts, nnn, sky,
                                                                    lds, cut, nfreq, ns, soff, S, A);
    else
        build_interp_t_kernel<T, 0><<<grid, BUILD_THREADS, 0, st>>>(bmapT, ldt, inds, wgts, nnn, sky,
                                                                    lds, cut, nfreq, ns, soff, S, A);
    return check_launch("build_interp_t");
}
template <typename T>
int launch_build_interp_bwd(const T* dA, const T* bmap, long long ldb, const int* inds,
                            const T* wgts, int nnn, const T* sky, long long lds, const int* cut,
                            int nfreq, int ns, long long soff, long long S, T* dsky, T* dBI,
                            long long ldd, T* dIs, cudaStream_t st) {
    if (ns <= 0 || nfreq <= 0) return 0;
    if (soff % TS) return set_error("build_interp_bwd: bad offset");
    dim3 grid((ns + TS - 1) / TS, nchunks<T>(nfreq));
    if (nnn == 4)
        build_interp_bwd_kernel<T, 4><<<grid, BUILD_THREADS, 0, st>>>(
            dA, bmap, ldb, inds, wgts, nnn, sky, lds, cut, nfreq, ns, soff, S, dsky, dBI, ldd, dIs);
    else if (nnn == 1)
        build_interp_bwd_kernel<T, 1><<<grid, BUILD_THREADS, 0, st>>>(
            dA, bmap, ldb, inds, wgts, nnn, sky, lds, cut, nfreq, ns, soff, S, dsky, dBI, ldd, dIs);
    else
        build_interp_bwd_kernel<T, 0><<<grid, BUILD_THREADS, 0, st>>>(
            dA, bmap, ldb, inds, wgts, nnn, sky, lds, cut, nfreq, ns, soff, S, dsky, dBI, ldd, dIs);
    return check_launch("build_interp_bwd");
}
template <typename T>
int launch_interp_transpose(const T* dBI, long long ldd, const int* rowptr, const int* col,
                            const T* val, int npix, int nfreq, T* dbmap, long long ldb,
                            cudaStream_t st) {
    if (npix <= 0 || nfreq <= 0) return 0;
    dim3 grid((npix + BUILD_THREADS - 1) / BUILD_THREADS, nfreq < 65535 ? nfreq : 65535);
    interp_transpose_kernel<T><<<grid, BUILD_THREADS, 0, st>>>(dBI, ldd, rowptr, col, val, npix,
                                                                nfreq, dbmap, ldb);
    return check_launch("interp_transpose");
}
template <typename T>
AiryArgs<T> make_airy(double Dew, double Dns, const double* diam_dev, double freq_ratio, int square, const T* sinzen,
                      const T* sin2az, const double* freqs) {
    AiryArgs<T> a;
    a.Dew = Dew;
    a.Dns = Dns;
    a.diam = diam_dev;
    a.kfac = 3.14159265358979323846 * freq_ratio / C_LIGHT;
    a.square = square;
    a.asym = sin2az != nullptr;
    a.sinzen = sinzen;
    a.sin2az = sin2az;
    a.freqs = freqs;
    return a;
}
template <typename T>
int launch_build_airy(double Dew, double Dns, const double* diam_dev, double freq_ratio, int square, const T* sinzen,
                      const T* sin2az, const double* freqs, const T* sky, long long lds,
                      const int* cut, int nfreq, int ns, int ns_pad, long long soff, long long S,
                      T* A, T* Bout, long long ldo, cudaStream_t st) {
    if (ns_pad <= 0 || nfreq <= 0) return 0;
    if (ns_pad % TS || soff % TS || (A && soff + ns_pad > S))
        return set_error("build_airy: bad padding/offset");
    dim3 grid(ns_pad / TS, nchunks<T>(nfreq));
    build_airy_kernel<T><<<grid, BUILD_THREADS, 0, st>>>(
        make_airy<T>(Dew, Dns, diam_dev, freq_ratio, square, sinzen, sin2az, freqs), sky, lds, cut, nfreq, ns,
        soff, S, A, Bout, ldo);
    return check_launch("build_airy");
}
template <typename T>
int launch_build_airy_bwd(const T* dA, double Dew, double Dns, const double* diam_dev, double freq_ratio, int square,
                          int full_grad, const T* sinzen, const T* sin2az, const double* freqs,
                          const T* sky, long long lds, const int* cut, int nfreq, int ns,
                          long long soff, long long S, T* dsky, double* dD, T* dIs, long long ldd,
                          cudaStream_t st) {
    if (ns <= 0 || nfreq <= 0) return 0;
    if (soff % TS) return set_error("build_airy_bwd: bad offset");
    dim3 grid((ns + TS - 1) / TS, nchunks<T>(nfreq));
    build_airy_bwd_kernel<T><<<grid, BUILD_THREADS, 0, st>>>(
        dA, make_airy<T>(Dew, Dns, diam_dev, freq_ratio, square, sinzen, sin2az, freqs), full_grad, sky, lds,
        cut, nfreq, ns, soff, S, dsky, dD, dIs, ldd);
    return check_launch("build_airy_bwd");
}
template <typename T>
int launch_gather_times(const T* dIs, long long ldd, const int* pos, int nt, int npix, int nfreq,
                        T* dsky, long long lds, cudaStream_t st) {
    if (npix <= 0 || nfreq <= 0 || nt <= 0) return 0;
    dim3 grid((npix + BUILD_THREADS - 1) / BUILD_THREADS, nfreq < 65535 ? nfreq : 65535);
    gather_times_kernel<T><<<grid, BUILD_THREADS, 0, st>>>(dIs, ldd, pos, nt, npix, nfreq, dsky, lds);
    return check_launch("gather_times");
}

// -------------------------------------------------------------------------------------
// Jones sandwich of the polarised beam modes (beam_model.py:347, :363; the reference runs it as
// einsum("ab...,bc...,dc...->ad...")): P[a][d] = sum_{b,c} J1[a][b] C[b][c] J2[d][c] for REAL
// 2 x 2 Jones matrices and coherencies, element-wise over planes in the tiled layout (HBM
// bound: 12 plane reads, 4 plane writes).  Planes are addressed through a small table so that
// the operands need not be stacked.  Backward (one pass): dJ1 = dP (J2 C^T), dJ2 = dP^T (J1 C),
// dC = J1^T dP J2; when both antennas share one beam model (J2 == J1, `same`) the two Jones
// gradients are summed into dJ1.
// -------------------------------------------------------------------------------------
template <typename T> struct Planes4 {
    const T* p[4];
};
template <typename T> struct PlanesOut4 {
    T* p[4];
};
template <typename T> struct Vec4 {
    T v[4];
};
template <typename T> __device__ __forceinline__ Vec4<T> ld4(const T* p, long long i) {
    Vec4<T> r;
    if constexpr (sizeof(T) == 4) {
        const float4 x = *reinterpret_cast<const float4*>(p + i);
        r.v[0] = x.x, r.v[1] = x.y, r.v[2] = x.z, r.v[3] = x.w;
    } else {
        const double2 x = *reinterpret_cast<const double2*>(p + i);
        const double2 y = *reinterpret_cast<const double2*>(p + i + 2);
        r.v[0] = x.x, r.v[1] = x.y, r.v[2] = y.x, r.v[3] = y.y;
    }
    return r;
}
template <typename T> __device__ __forceinline__ void st4(T* p, long long i, const Vec4<T>& r) {
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(p + i) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    } else {
        *reinterpret_cast<double2*>(p + i) = make_double2(r.v[0], r.v[1]);
        *reinterpret_cast<double2*>(p + i + 2) = make_double2(r.v[2], r.v[3]);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
jones_sandwich_kernel(Planes4<T> J1, Planes4<T> J2, Planes4<T> C, long long n, PlanesOut4<T> P) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        Vec4<T> j1[4], j2[4], c[4], o[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            j1[m] = ld4(J1.p[m], i);
            j2[m] = ld4(J2.p[m], i);
            c[m] = ld4(C.p[m], i);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            // M = J1 C (index [a][c]), P = M J2^T
            T M[4];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc)
                    M[2 * a + cc] = j1[2 * a].v[e] * c[cc].v[e] + j1[2 * a + 1].v[e] * c[2 + cc].v[e];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int d = 0; d < 2; ++d)
                    o[2 * a + d].v[e] = M[2 * a] * j2[2 * d].v[e] + M[2 * a + 1] * j2[2 * d + 1].v[e];
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) st4(P.p[m], i, o[m]);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
jones_sandwich_bwd_kernel(Planes4<T> dP, Planes4<T> J1, Planes4<T> J2, Planes4<T> C, long long n,
                          int same, PlanesOut4<T> dJ1, PlanesOut4<T> dJ2, PlanesOut4<T> dC) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        Vec4<T> g[4], j1[4], j2[4], c[4], o1[4], o2[4], oc[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            g[m] = ld4(dP.p[m], i);
            j1[m] = ld4(J1.p[m], i);
            j2[m] = ld4(J2.p[m], i);
            c[m] = ld4(C.p[m], i);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            T G[4], A[4], B[4], K[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) G[m] = g[m].v[e], A[m] = j1[m].v[e], B[m] = j2[m].v[e], K[m] = c[m].v[e];
            // N2 = J2 C^T ([d][b]), N1 = J1 C ([a][c])
            T N2[4], N1[4];
#pragma unroll
            for (int d = 0; d < 2; ++d)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    N2[2 * d + b] = B[2 * d] * K[2 * b] + B[2 * d + 1] * K[2 * b + 1];
                    N1[2 * d + b] = A[2 * d] * K[b] + A[2 * d + 1] * K[2 + b];
                }
            T D1[4], D2[4], DC[4];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    // dJ1[a][b] = sum_d G[a][d] N2[d][b];  dJ2[d=a][c=b] = sum_a' G[a'][d] N1[a'][c]
                    D1[2 * a + b] = G[2 * a] * N2[b] + G[2 * a + 1] * N2[2 + b];
                    D2[2 * a + b] = G[a] * N1[b] + G[2 + a] * N1[2 + b];
                }
            // dC[b][c] = sum_{a,d} J1[a][b] G[a][d] J2[d][c]: T1 = J1^T G ([b][d]), dC = T1 J2
            T T1[4];
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int d = 0; d < 2; ++d) T1[2 * b + d] = A[b] * G[d] + A[2 + b] * G[2 + d];
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) DC[2 * b + cc] = T1[2 * b] * B[cc] + T1[2 * b + 1] * B[2 + cc];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                o1[m].v[e] = same ? D1[m] + D2[m] : D1[m];
                o2[m].v[e] = D2[m];
                oc[m].v[e] = DC[m];
            }
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            if (dJ1.p[m] != nullptr) st4(dJ1.p[m], i, o1[m]);
            if (!same && dJ2.p[m] != nullptr) st4(dJ2.p[m], i, o2[m]);
            if (dC.p[m] != nullptr) st4(dC.p[m], i, oc[m]);
        }
    }
}

template <typename T> static int sandwich_grid(long long n) {
    const long long want = (n / 4 + 255) / 256;
    const long long cap = 148LL * 16;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}
template <typename T>
int launch_jones_sandwich(const T* const* J1, const T* const* J2, const T* const* C, long long n,
                          T* const* P, cudaStream_t st) {
    if (n <= 0) return 0;
    if (n % 4) return set_error("jones_sandwich: plane length must be a multiple of 4");
    Planes4<T> a, b, c;
    PlanesOut4<T> o;
    for (int m = 0; m < 4; ++m) a.p[m] = J1[m], b.p[m] = J2[m], c.p[m] = C[m], o.p[m] = P[m];
    jones_sandwich_kernel<T><<<sandwich_grid<T>(n), 256, 0, st>>>(a, b, c, n, o);
    return check_launch("jones_sandwich");
}
template <typename T>
int launch_jones_sandwich_bwd(const T* const* dP, const T* const* J1, const T* const* J2,
                              const T* const* C, long long n, int same, T* const* dJ1, T* const* dJ2,
                              T* const* dC, cudaStream_t st) {
    if (n <= 0) return 0;
    if (n % 4) return set_error("jones_sandwich_bwd: plane length must be a multiple of 4");
    Planes4<T> g, a, b, c;
    PlanesOut4<T> o1, o2, oc;
    for (int m = 0; m < 4; ++m) {
        g.p[m] = dP[m], a.p[m] = J1[m], b.p[m] = J2[m], c.p[m] = C[m];
        o1.p[m] = dJ1 ? dJ1[m] : nullptr;
        o2.p[m] = dJ2 ? dJ2[m] : nullptr;
        oc.p[m] = dC ? dC[m] : nullptr;
    }
    jones_sandwich_bwd_kernel<T><<<sandwich_grid<T>(n), 256, 0, st>>>(g, a, b, c, n, same, o1, o2, oc);
    return check_launch("jones_sandwich_bwd");
}

}  // namespace b200rime

using namespace b200rime;
#define ST(s) ((cudaStream_t)(s))

extern "C" {

int b200rime_pack_f32(const float* X, long long ldx, int nfreq, int ns, int ns_pad, long long soff,
                      long long S, float* A, void* stream) {
    return launch_pack<float>(X, ldx, nfreq, ns, ns_pad, soff, S, A, ST(stream));
}
int b200rime_pack_f64(const double* X, long long ldx, int nfreq, int ns, int ns_pad,
                      long long soff, long long S, double* A, void* stream) {
    return launch_pack<double>(X, ldx, nfreq, ns, ns_pad, soff, S, A, ST(stream));
}
int b200rime_unpack_f32(const float* A, long long ldx, int nfreq, int ns, long long soff,
                        long long S, float* X, void* stream) {
    return launch_unpack<float>(A, ldx, nfreq, ns, soff, S, X, ST(stream));
}
int b200rime_unpack_f64(const double* A, long long ldx, int nfreq, int ns, long long soff,
                        long long S, double* X, void* stream) {
    return launch_unpack<double>(A, ldx, nfreq, ns, soff, S, X, ST(stream));
}
int b200rime_build_interp_t_f32(const float* bmapT, long long ldt, const int* inds,
                                const float* wgts, int nnn, const float* sky, long long lds,
                                const int* cut, int nfreq, int ns, int ns_pad, long long soff,
                                long long S, float* A, void* stream) {
    return launch_build_interp_t<float>(bmapT, ldt, inds, wgts, nnn, sky, lds, cut, nfreq, ns, ns_pad,
                                        soff, S, A, ST(stream));
}
int b200rime_build_interp_t_f64(const double* bmapT, long long ldt, const int* inds,
                                const double* wgts, int nnn, const double* sky, long long lds,
                                const int* cut, int nfreq, int ns, int ns_pad, long long soff,
                                long long S, double* A, void* stream) {
    return launch_build_interp_t<double>(bmapT, ldt, inds, wgts, nnn, sky, lds, cut, nfreq, ns,
                                         ns_pad, soff, S, A, ST(stream));
}
int b200rime_build_interp_f32(const float* bmap, long long ldb, const int* inds, const float* wgts,
                              int nnn, const float* sky, long long lds, const int* cut, int nfreq,
                              int ns, int ns_pad, long long soff, long long S, float* A,
                              void* stream) {
    return launch_build_interp<float>(bmap, ldb, inds, wgts, nnn, sky, lds, cut, nfreq, ns, ns_pad,
                                      soff, S, A, ST(stream));
}
int b200rime_build_interp_f64(const double* bmap, long long ldb, const int* inds,
                              const double* wgts, int nnn, const double* sky, long long lds,
                              const int* cut, int nfreq, int ns, int ns_pad, long long soff,
                              long long S, double* A, void* stream) {
    return launch_build_interp<double>(bmap, ldb, inds, wgts, nnn, sky, lds, cut, nfreq, ns, ns_pad,
                                       soff, S, A, ST(stream));
}
int b200rime_build_interp_bwd_f32(const float* dA, const float* bmap, long long ldb,
                                  const int* inds, const float* wgts, int nnn, const float* sky,
                                  long long lds, const int* cut, int nfreq, int ns, long long soff,
                                  long long S, float* dsky, float* dBI, long long ldd, float* dIs,
                                  void* stream) {
    return launch_build_interp_bwd<float>(dA, bmap, ldb, inds, wgts, nnn, sky, lds, cut, nfreq, ns,
                                          soff, S, dsky, dBI, ldd, dIs, ST(stream));
}
int b200rime_build_interp_bwd_f64(const double* dA, const double* bmap, long long ldb,
                                  const int* inds, const double* wgts, int nnn, const double* sky,
                                  long long lds, const int* cut, int nfreq, int ns, long long soff,
                                  long long S, double* dsky, double* dBI, long long ldd, double* dIs,
                                  void* stream) {
    return launch_build_interp_bwd<double>(dA, bmap, ldb, inds, wgts, nnn, sky, lds, cut, nfreq, ns,
                                           soff, S, dsky, dBI, ldd, dIs, ST(stream));
}
int b200rime_build_interp_bwd_t_f32(const float* dA, const float* bmapT, long long ldt,
                                    const int* inds, const float* wgts, const float* sky,
                                    long long lds, const int* cut, int nfreq, int ns, long long soff,
                                    long long S, float* dBI, long long ldd, float* dIs,
                                    void* stream) {
    using namespace b200rime;
    if (ns <= 0 || nfreq <= 0) return 0;
    if (soff % TS) return set_error("build_interp_bwd_t: bad offset");
    if (dA == nullptr || bmapT == nullptr || sky == nullptr || inds == nullptr || wgts == nullptr)
        return set_error("build_interp_bwd_t: null operand");
    if (ldt < (long long)nchunks<float>(nfreq) * Cfg<float>::KC)
        return set_error("build_interp_bwd_t: needs a channel-major beam map padded to whole chunks");
    dim3 grid((ns + TS - 1) / TS, nchunks<float>(nfreq));
    build_interp_bwd_t_kernel<<<grid, BUILD_THREADS, 0, ST(stream)>>>(
        dA, bmapT, ldt, inds, wgts, sky, lds, cut, nfreq, ns, soff, S, dBI, ldd, dIs);
    return check_launch("build_interp_bwd_t");
}
int b200rime_interp_transpose_f32(const float* dBI, long long ldd, const int* rowptr,
                                  const int* col, const float* val, int npix, int nfreq,
                                  float* dbmap, long long ldb, void* stream) {
    return launch_interp_transpose<float>(dBI, ldd, rowptr, col, val, npix, nfreq, dbmap, ldb,
                                          ST(stream));
}
int b200rime_interp_transpose_f64(const double* dBI, long long ldd, const int* rowptr,
                                  const int* col, const double* val, int npix, int nfreq,
                                  double* dbmap, long long ldb, void* stream) {
    return launch_interp_transpose<double>(dBI, ldd, rowptr, col, val, npix, nfreq, dbmap, ldb,
                                           ST(stream));
}
int b200rime_build_airy_f32(double Dew, double Dns, const double* diam_dev, double freq_ratio, int square,
                            const float* sinzen, const float* sin2az, const double* freqs,
                            const float* sky, long long lds, const int* cut, int nfreq, int ns,
                            int ns_pad, long long soff, long long S, float* A, float* Bout,
                            long long ldo, void* stream) {
    return launch_build_airy<float>(Dew, Dns, diam_dev, freq_ratio, square, sinzen, sin2az, freqs, sky, lds,
                                    cut, nfreq, ns, ns_pad, soff, S, A, Bout, ldo, ST(stream));
}
int b200rime_build_airy_f64(double Dew, double Dns, const double* diam_dev, double freq_ratio, int square,
                            const double* sinzen, const double* sin2az, const double* freqs,
                            const double* sky, long long lds, const int* cut, int nfreq, int ns,
                            int ns_pad, long long soff, long long S, double* A, double* Bout,
                            long long ldo, void* stream) {
    return launch_build_airy<double>(Dew, Dns, diam_dev, freq_ratio, square, sinzen, sin2az, freqs, sky, lds,
                                     cut, nfreq, ns, ns_pad, soff, S, A, Bout, ldo, ST(stream));
}
int b200rime_airy_bwd_blocks(int nfreq, int ns) {
    // same chunking for f32 and f64 callers is not assumed: the caller passes nfreq already
    // divided into its own KC; use the finer (f64) chunking as the upper bound
    return ((ns + TS - 1) / TS) * ((nfreq + Cfg<double>::KC - 1) / Cfg<double>::KC);
}
int b200rime_build_airy_bwd_f32(const float* dA, double Dew, double Dns, const double* diam_dev, double freq_ratio,
                                int square, int full_grad, const float* sinzen,
                                const float* sin2az, const double* freqs, const float* sky,
                                long long lds, const int* cut, int nfreq, int ns, long long soff,
                                long long S, float* dsky, double* dD, float* dIs, long long ldd,
                                void* stream) {
    return launch_build_airy_bwd<float>(dA, Dew, Dns, diam_dev, freq_ratio, square, full_grad, sinzen, sin2az,
                                        freqs, sky, lds, cut, nfreq, ns, soff, S, dsky, dD, dIs, ldd,
                                        ST(stream));
}
int b200rime_build_airy_bwd_f64(const double* dA, double Dew, double Dns, const double* diam_dev, double freq_ratio,
                                int square, int full_grad, const double* sinzen,
                                const double* sin2az, const double* freqs, const double* sky,
                                long long lds, const int* cut, int nfreq, int ns, long long soff,
                                long long S, double* dsky, double* dD, double* dIs, long long ldd,
                                void* stream) {
    return launch_build_airy_bwd<double>(dA, Dew, Dns, diam_dev, freq_ratio, square, full_grad, sinzen,
                                         sin2az, freqs, sky, lds, cut, nfreq, ns, soff, S, dsky, dD,
                                         dIs, ldd, ST(stream));
}
int b200rime_gather_times_f32(const float* dIs, long long ldd, const int* pos, int nt, int npix,
                              int nfreq, float* dsky, long long lds, void* stream) {
    return launch_gather_times<float>(dIs, ldd, pos, nt, npix, nfreq, dsky, lds, ST(stream));
}
int b200rime_gather_times_f64(const double* dIs, long long ldd, const int* pos, int nt, int npix,
                              int nfreq, double* dsky, long long lds, void* stream) {
    return launch_gather_times<double>(dIs, ldd, pos, nt, npix, nfreq, dsky, lds, ST(stream));
}

int b200rime_jones_sandwich_f32(const float* const* J1, const float* const* J2,
                                const float* const* C, long long n, float* const* P, void* stream) {
    return launch_jones_sandwich<float>(J1, J2, C, n, P, ST(stream));
}
int b200rime_jones_sandwich_f64(const double* const* J1, const double* const* J2,
                                const double* const* C, long long n, double* const* P,
                                void* stream) {
    return launch_jones_sandwich<double>(J1, J2, C, n, P, ST(stream));
}
int b200rime_jones_sandwich_bwd_f32(const float* const* dP, const float* const* J1,
                                    const float* const* J2, const float* const* C, long long n,
                                    int same, float* const* dJ1, float* const* dJ2,
                                    float* const* dC, void* stream) {
    return launch_jones_sandwich_bwd<float>(dP, J1, J2, C, n, same, dJ1, dJ2, dC, ST(stream));
}
int b200rime_jones_sandwich_bwd_f64(const double* const* dP, const double* const* J1,
                                    const double* const* J2, const double* const* C, long long n,
                                    int same, double* const* dJ1, double* const* dJ2,
                                    double* const* dC, void* stream) {
    return launch_jones_sandwich_bwd<double>(dP, J1, J2, C, n, same, dJ1, dJ2, dC, ST(stream));
}
}  // extern "C"
