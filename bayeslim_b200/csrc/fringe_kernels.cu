// Fringe-sum kernels for sm_100a: forward (K1), per-unit reduction, backward to the perceived
// sky (K2) and backward to the baseline vectors (K3).
//
// Design: the (Nbl, Nf, Ns) fringe tensor of the reference
// (telescope_model.py:356) is never formed.  A thread owns one baseline (K1, K3) or one source
// (K2) and KC frequency channels; the reduction axis is streamed through shared memory by the
// TMA engine (1-D cp.async.bulk + mbarrier, double buffered) and read with warp-broadcast
// LDS.128.  The fringe phase is reduced mod one cycle in float64 once per (baseline, source,
// chunk) and advanced across the chunk by a complex rotation recurrence seeded at the chunk
// centre, so the steady state is 4 FP32 ops (rotation) + 2 (multiply-accumulate) per
// source.baseline.channel with no transcendental.  Every output has exactly one owner thread
// and a fixed summation order: results are bitwise reproducible.
#include <type_traits>
#include "rime_math.cuh"
#include "internal.h"

namespace b200rime {

// -------------------------------------------------------------------------------------
// shared-memory plans
// -------------------------------------------------------------------------------------
template <typename T> struct FwdSmem {
    static constexpr int KC = Cfg<T>::KC;
    static constexpr int A_BYTES = SRC_TILE * KC * (int)sizeof(T);
    static constexpr int S_BYTES = SRC_TILE * 4 * (int)sizeof(double);
    static constexpr int STAGE_BYTES = A_BYTES + S_BYTES;
    static constexpr int KF_OFF = 2 * STAGE_BYTES;
    static constexpr int BAR_OFF = KF_OFF + KC * 8;
    static constexpr int TOTAL = BAR_OFF + 16;
};
template <typename T> struct SkySmem {
    static constexpr int KC = Cfg<T>::KC;
    static constexpr int ROW_BYTES = KC * 2 * (int)sizeof(T);
    static constexpr int G_BYTES = BL_TILE * ROW_BYTES;
    static constexpr int B_BYTES = BL_TILE * 4 * (int)sizeof(double);
    static constexpr int STAGE_BYTES = G_BYTES + B_BYTES;
    static constexpr int KF_OFF = 2 * STAGE_BYTES;
    static constexpr int BAR_OFF = KF_OFF + KC * 8;
    static constexpr int TOTAL = BAR_OFF + 16;
};

// -------------------------------------------------------------------------------------
// K1: forward.  grid = (ceil(Nbl/128), nchunk, nunits), block = 128 (thread <-> baseline)
// -------------------------------------------------------------------------------------
template <typename T, bool UNIFORM>
__global__ void __launch_bounds__(FWD_THREADS, sizeof(T) == 4 ? B200_MINB_F32 : 2)
fringe_sum_fwd_kernel(const T* __restrict__ A, const double* __restrict__ shat,
                      const double* __restrict__ blv, const double* __restrict__ freqs,
                      const int4* __restrict__ units, int nbl, int nfreq, long long S,
                      double sgn_over_c, T* __restrict__ vpart) {
    constexpr int KC = Cfg<T>::KC;
    using SM = FwdSmem<T>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
    double* kf = reinterpret_cast<double*>(smem + SM::KF_OFF);

    const int tid = threadIdx.x;
    const int chunk = blockIdx.y;
    const int nfp = gridDim.y * KC;
    const int4 un = units[blockIdx.z];
    const int ntiles = (un.z - un.y) / SRC_TILE;
    const int b = blockIdx.x * blockDim.x + tid;      // block = 128, 64 or 32 baselines
    const bool valid = b < nbl;
    double bx = 0.0, by = 0.0, bz = 0.0;
    if (valid) {
        const double* p = blv + 4 * (size_t)b;
        bx = p[0];
        by = p[1];
        bz = p[2];
    }
    const ChunkFreq cf = chunk_freq(freqs, nfreq, chunk, KC, sgn_over_c);
    if (!UNIFORM) {
        for (int k = tid; k < KC; k += blockDim.x) {
            int f = min(chunk * KC + k, nfreq - 1);
            kf[k] = sgn_over_c * freqs[f];
        }
    }
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const T* Abase = A + ((size_t)chunk * (size_t)S + (size_t)un.y) * KC;
    const double* Sbase = shat + (size_t)un.y * 4;
    auto issue = [&](int tile, int stage) {
        unsigned char* dst = smem + stage * SM::STAGE_BYTES;
        mbar_expect_tx(&bars[stage], SM::STAGE_BYTES);
        bulk_g2s(dst, Abase + (size_t)tile * SRC_TILE * KC, SM::A_BYTES, &bars[stage]);
        bulk_g2s(dst + SM::A_BYTES, Sbase + (size_t)tile * SRC_TILE * 4, SM::S_BYTES, &bars[stage]);
    };
    if (tid == 0 && ntiles > 0) issue(0, 0);

    FwdTile<T, KC> acc;
    acc.zero();

    for (int it = 0; it < ntiles; ++it) {
        const int stage = it & 1;
        if (tid == 0 && it + 1 < ntiles) issue(it + 1, stage ^ 1);
        mbar_wait(&bars[stage], (it >> 1) & 1);
        const T* As = reinterpret_cast<const T*>(smem + stage * SM::STAGE_BYTES);
        const double4* Ss =
            reinterpret_cast<const double4*>(smem + stage * SM::STAGE_BYTES + SM::A_BYTES);
        if (UNIFORM) {
            // software pipeline: the seeds of source s+1 (float64 phase reduction + trigonometry,
            // a long dependent chain that uses no FP32-pipe throughput) are issued inside the
            // FFMA2 stream of source s, so a warp never leaves the FP32 pipe idle between sources
            T zr, zi, wr, wi;
            {
                const double4 sh = Ss[0];
                chunk_seed(fma(bx, sh.x, fma(by, sh.y, bz * sh.z)), cf.k_mid, cf.k_step, zr, zi, wr,
                           wi);
            }
#pragma unroll 1
            for (int s = 0; s < SRC_TILE; ++s) {
                T nzr, nzi, nwr, nwi;
                const double4 sh = Ss[(s + 1) & (SRC_TILE - 1)];
                chunk_seed(fma(bx, sh.x, fma(by, sh.y, bz * sh.z)), cf.k_mid, cf.k_step, nzr, nzi,
                           nwr, nwi);
                acc.accumulate(As + s * KC, zr, zi, wr, wi);
                zr = nzr;
                zi = nzi;
                wr = nwr;
                wi = nwi;
            }
        } else {
#pragma unroll 1
            for (int s = 0; s < SRC_TILE; ++s) {
                const double4 sh = Ss[s];
                const double u = fma(bx, sh.x, fma(by, sh.y, bz * sh.z));
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    T zr, zi;
                    channel_cis(u, kf[k], zr, zi);
                    acc.mac(k, As[s * KC + k], zr, zi);
                }
            }
        }
        __syncthreads();
    }

    if (valid) {
        T* out = vpart + (((size_t)blockIdx.z * nbl + b) * nfp + (size_t)chunk * KC) * 2;
        constexpr int N = Vec16<T>::N;
        constexpr int NC = N / 2;
#pragma unroll
        for (int k = 0; k < KC; k += NC) {
            T r0, i0, r1, i1;
            acc.get(k, r0, i0);
            acc.get(k + NC - 1, r1, i1);
            if (NC == 2) {
                *reinterpret_cast<float4*>(out + 2 * k) =
                    make_float4((float)r0, (float)i0, (float)r1, (float)i1);
            } else {
                *reinterpret_cast<double2*>(out + 2 * k) = make_double2((double)r0, (double)i0);
            }
        }
    }
}

// -------------------------------------------------------------------------------------
// per-time reduction of unit partials into the (strided) visibility tensor
// grid = (ceil(Nbl*Nf/256), Nt)
// -------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
reduce_units_kernel(const T* __restrict__ vpart, const int* __restrict__ ubeg, int nbl, int nfreq,
                    int nfp, T* __restrict__ V, long long sb, long long st, long long sf,
                    double are, double aim, int accumulate) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nbl * nfreq) return;
    const int t = blockIdx.y;
    const int b = (int)(idx / nfreq);
    const int f = (int)(idx - (long long)b * nfreq);
    double sr = 0.0, si = 0.0;
    const int u0 = ubeg[t], u1 = ubeg[t + 1];
    for (int u = u0; u < u1; ++u) {
        const T* p = vpart + (((size_t)u * nbl + b) * nfp + f) * 2;
        sr += (double)p[0];
        si += (double)p[1];
    }
    double orr = are * sr - aim * si;
    double oi = are * si + aim * sr;
    T* o = V + ((long long)b * sb + (long long)t * st + (long long)f * sf) * 2;
    if (accumulate) {
        orr += (double)o[0];
        oi += (double)o[1];
    }
    o[0] = (T)orr;
    o[1] = (T)oi;
}

// -------------------------------------------------------------------------------------
// Likelihood epilogue: the unit reduction fused with the Gaussian
// chi-square of optim.py:1012-1024 (res = V - D, chisq = sum conj(res) res icov, apply_icov
// :1836 with cov_axis None).  V is never written: its slot receives the cotangent
// G = dchisq/dV = 2 icov (V - D) (PyTorch's convention for a real loss), which is what the
// backward kernels consume; chisq goes out as one float64 partial per block, summed in block
// order by the caller (bitwise reproducible).  grid = (ceil(nbl * nfreq / 256), nt).
// -------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
reduce_units_chisq_kernel(const T* __restrict__ vpart, const int* __restrict__ ubeg, int nbl,
                          int nfreq, int nfp, T* __restrict__ V, const T* __restrict__ D,
                          const T* __restrict__ W, long long sb, long long st, long long sf,
                          int accumulate, double* __restrict__ chi_part) {
    __shared__ double red[8];
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    double chi = 0.0;
    if (idx < (long long)nbl * nfreq) {
        const int b = (int)(idx / nfreq);
        const int f = (int)(idx - (long long)b * nfreq);
        double sr = 0.0, si = 0.0;
        const int u0 = ubeg[t], u1 = ubeg[t + 1];
        for (int u = u0; u < u1; ++u) {
            const T* p = vpart + (((size_t)u * nbl + b) * nfp + f) * 2;
            sr += (double)p[0];
            si += (double)p[1];
        }
        const long long e = (long long)b * sb + (long long)t * st + (long long)f * sf;
        T* o = V + e * 2;
        if (accumulate) {
            sr += (double)o[0];
            si += (double)o[1];
        }
        const double rr = sr - (double)D[e * 2], ri = si - (double)D[e * 2 + 1];
        const double w = W != nullptr ? (double)W[e] : 1.0;
        chi = w * (rr * rr + ri * ri);
        o[0] = (T)(2.0 * w * rr);
        o[1] = (T)(2.0 * w * ri);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) chi += __shfl_xor_sync(0xffffffffu, chi, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = chi;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) tot += red[w8];
        chi_part[(size_t)t * gridDim.x + blockIdx.x] = tot;
    }
}

// -------------------------------------------------------------------------------------
// K2: backward to the perceived sky.  grid = (S/128, nchunk), block = 128 (thread <-> source)
// -------------------------------------------------------------------------------------
template <typename T, bool UNIFORM>
__global__ void __launch_bounds__(SKY_THREADS, sizeof(T) == 4 ? B200_MINB_SKY_F32 : 3)
fringe_sum_bwd_sky_kernel(const T* __restrict__ Gp, const double* __restrict__ shat,
                          const double* __restrict__ blv, const double* __restrict__ freqs,
                          const int* __restrict__ tile_time, int nbl, int nt, int nfreq,
                          long long S, double sgn_over_c, T* __restrict__ dA) {
    constexpr int KC = Cfg<T>::KC;
    using SM = SkySmem<T>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
    double* kf = reinterpret_cast<double*>(smem + SM::KF_OFF);

    const int tid = threadIdx.x;
    const int chunk = blockIdx.y;
    const int nfp = gridDim.y * KC;
    const int t = tile_time[blockIdx.x];
    const size_t s = (size_t)blockIdx.x * SKY_THREADS + tid;
    const double sx = shat[4 * s + 0], sy = shat[4 * s + 1], sz = shat[4 * s + 2];
    const ChunkFreq cf = chunk_freq(freqs, nfreq, chunk, KC, sgn_over_c);
    if (!UNIFORM) {
        for (int k = tid; k < KC; k += SKY_THREADS) {
            int f = min(chunk * KC + k, nfreq - 1);
            kf[k] = sgn_over_c * freqs[f];
        }
    }
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int ntiles = (nbl + BL_TILE - 1) / BL_TILE;
    // the cotangent is laid out [t][chunk][baseline][KC]: a tile of rows is one contiguous block
    auto issue = [&](int tile, int stage) {
        const int b0 = tile * BL_TILE;
        const int rows = min(BL_TILE, nbl - b0);
        unsigned char* dst = smem + stage * SM::STAGE_BYTES;
        mbar_expect_tx(&bars[stage], rows * (SM::ROW_BYTES + 32));
        const T* src = Gp + ((((size_t)t * gridDim.y + chunk) * nbl + b0) * KC) * 2;
        bulk_g2s(dst, src, rows * SM::ROW_BYTES, &bars[stage]);
        bulk_g2s(dst + SM::G_BYTES, blv + (size_t)b0 * 4, rows * 32, &bars[stage]);
    };
    if (tid == 0) issue(0, 0);

    SkyTile<T, KC> acc;
    acc.zero();
    T* out = dA + ((size_t)chunk * (size_t)S + s) * KC;
    bool first_flush = true;
    constexpr int SEG_TILES = BL_SEGMENT / BL_TILE;

    for (int it = 0; it < ntiles; ++it) {
        const int stage = it & 1;
        if (tid == 0 && it + 1 < ntiles) issue(it + 1, stage ^ 1);
        mbar_wait(&bars[stage], (it >> 1) & 1);
        const T* Gs = reinterpret_cast<const T*>(smem + stage * SM::STAGE_BYTES);
        const double4* Bs =
            reinterpret_cast<const double4*>(smem + stage * SM::STAGE_BYTES + SM::G_BYTES);
        const int rows = min(BL_TILE, nbl - it * BL_TILE);
        if (UNIFORM) {
            // seeds of baseline j+1 are issued inside the FFMA2 stream of baseline j (see K1)
            T zr, zi, wr, wi;
            {
                const double4 bv = Bs[0];
                chunk_seed(fma(bv.x, sx, fma(bv.y, sy, bv.z * sz)), cf.k_mid, cf.k_step, zr, zi, wr,
                           wi);
            }
#pragma unroll 1
            for (int j = 0; j < rows; ++j) {
                T nzr, nzi, nwr, nwi;
                const double4 bv = Bs[min(j + 1, rows - 1)];
                chunk_seed(fma(bv.x, sx, fma(bv.y, sy, bv.z * sz)), cf.k_mid, cf.k_step, nzr, nzi,
                           nwr, nwi);
                acc.accumulate(Gs + j * 2 * KC, zr, zi, wr, wi);
                zr = nzr;
                zi = nzi;
                wr = nwr;
                wi = nwi;
            }
        } else {
#pragma unroll 1
            for (int j = 0; j < rows; ++j) {
                const double4 bv = Bs[j];
                const double u = fma(bv.x, sx, fma(bv.y, sy, bv.z * sz));
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    T zr, zi;
                    channel_cis(u, kf[k], zr, zi);
                    acc.mac(k, zr, zi, Gs[j * 2 * KC + 2 * k], Gs[j * 2 * KC + 2 * k + 1]);
                }
            }
        }
        __syncthreads();
        // bounded-length fp accumulation: spill the segment sum to the (thread-owned) output
        if ((it + 1) % SEG_TILES == 0 || it + 1 == ntiles) {
            constexpr int N = Vec16<T>::N;
#pragma unroll
            for (int k = 0; k < KC; k += N) {
                Vec16<T>* o = reinterpret_cast<Vec16<T>*>(out + k);
                if (N == 4) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!first_flush) v = *reinterpret_cast<float4*>(o);
                    v.x += (float)acc.value(k);
                    v.y += (float)acc.value(k + 1);
                    v.z += (float)acc.value(k + N - 2);
                    v.w += (float)acc.value(k + N - 1);
                    *reinterpret_cast<float4*>(o) = v;
                } else {
                    double2 v = make_double2(0.0, 0.0);
                    if (!first_flush) v = *reinterpret_cast<double2*>(o);
                    v.x += (double)acc.value(k);
                    v.y += (double)acc.value(k + 1);
                    *reinterpret_cast<double2*>(o) = v;
                }
            }
            acc.zero();
            first_flush = false;
        }
    }
    if (ntiles == 0) {
#pragma unroll
        for (int k = 0; k < KC; ++k) out[k] = 0;
    }
}

// -------------------------------------------------------------------------------------
// K3: backward to the baseline vectors.  grid/block as K1.
// -------------------------------------------------------------------------------------
template <typename T, bool UNIFORM>
__global__ void __launch_bounds__(FWD_THREADS, sizeof(T) == 4 ? B200_MINB_F32 : 2)
fringe_sum_bwd_bl_kernel(const T* __restrict__ Gp, const T* __restrict__ A,
                         const double* __restrict__ shat, const double* __restrict__ blv,
                         const double* __restrict__ freqs, const int4* __restrict__ units, int nbl,
                         int nt, int nfreq, long long S, double sgn_over_c,
                         double* __restrict__ dblpart) {
    constexpr int KC = Cfg<T>::KC;
    using SM = FwdSmem<T>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
    double* kf = reinterpret_cast<double*>(smem + SM::KF_OFF);

    const int tid = threadIdx.x;
    const int chunk = blockIdx.y;
    const int nchunk = gridDim.y;
    const int nfp = nchunk * KC;
    const int4 un = units[blockIdx.z];
    const int ntiles = (un.z - un.y) / SRC_TILE;
    const int b = blockIdx.x * blockDim.x + tid;      // block = 128, 64 or 32 baselines
    const bool valid = b < nbl;
    double bx = 0.0, by = 0.0, bz = 0.0;
    constexpr bool PACKED = std::is_same<T, float>::value && UNIFORM;
    // pre-scaled cotangent G'_k = nu_k G_k of this thread's baseline: scalar arrays, or (re, im)
    // register pairs for the packed Horner form
    T gr[PACKED ? 1 : KC], gi[PACKED ? 1 : KC];
    P2 g2[PACKED ? KC : 1];
    if (PACKED) {
#pragma unroll
        for (int k = 0; k < (PACKED ? KC : 1); ++k) g2[k] = p2(0.f, 0.f);
    } else {
#pragma unroll
        for (int k = 0; k < (PACKED ? 1 : KC); ++k) {
            gr[k] = 0;
            gi[k] = 0;
        }
    }
    if (valid) {
        const double* p = blv + 4 * (size_t)b;
        bx = p[0];
        by = p[1];
        bz = p[2];
        const T* g = Gp + ((((size_t)un.x * nchunk + chunk) * nbl + b) * KC) * 2;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const int f = min(chunk * KC + k, nfreq - 1);
            const T nu = (T)freqs[f];
            if (PACKED) {
                g2[PACKED ? k : 0] = p2((float)(g[2 * k] * nu), (float)(g[2 * k + 1] * nu));
            } else {
                gr[PACKED ? 0 : k] = g[2 * k] * nu;
                gi[PACKED ? 0 : k] = g[2 * k + 1] * nu;
            }
        }
    }
    const ChunkFreq cf = chunk_freq(freqs, nfreq, chunk, KC, sgn_over_c);
    if (!UNIFORM) {
        for (int k = tid; k < KC; k += blockDim.x) {
            int f = min(chunk * KC + k, nfreq - 1);
            kf[k] = sgn_over_c * freqs[f];
        }
    }
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const T* Abase = A + ((size_t)chunk * (size_t)S + (size_t)un.y) * KC;
    const double* Sbase = shat + (size_t)un.y * 4;
    auto issue = [&](int tile, int stage) {
        unsigned char* dst = smem + stage * SM::STAGE_BYTES;
        mbar_expect_tx(&bars[stage], SM::STAGE_BYTES);
        bulk_g2s(dst, Abase + (size_t)tile * SRC_TILE * KC, SM::A_BYTES, &bars[stage]);
        bulk_g2s(dst + SM::A_BYTES, Sbase + (size_t)tile * SRC_TILE * 4, SM::S_BYTES, &bars[stage]);
    };
    if (tid == 0 && ntiles > 0) issue(0, 0);

    double dbx = 0.0, dby = 0.0, dbz = 0.0;
    for (int it = 0; it < ntiles; ++it) {
        const int stage = it & 1;
        if (tid == 0 && it + 1 < ntiles) issue(it + 1, stage ^ 1);
        mbar_wait(&bars[stage], (it >> 1) & 1);
        const T* As = reinterpret_cast<const T*>(smem + stage * SM::STAGE_BYTES);
        const double4* Ss =
            reinterpret_cast<const double4*>(smem + stage * SM::STAGE_BYTES + SM::A_BYTES);
        if (UNIFORM) {
            T zr, zi, wr, wi;
            {
                const double4 sh = Ss[0];
                chunk_seed(fma(bx, sh.x, fma(by, sh.y, bz * sh.z)), cf.k_mid, cf.k_step, zr, zi, wr,
                           wi);
            }
#pragma unroll 1
            for (int s = 0; s < SRC_TILE; ++s) {
                T nzr, nzi, nwr, nwi;
                const double4 shn = Ss[(s + 1) & (SRC_TILE - 1)];
                chunk_seed(fma(bx, shn.x, fma(by, shn.y, bz * shn.z)), cf.k_mid, cf.k_step, nzr,
                           nzi, nwr, nwi);
                T du;
                if constexpr (PACKED)
                    du = bl_accumulate_f32<KC>(As + s * KC, zr, zi, wr, wi, g2);
                else
                    du = bl_accumulate<T, KC>(As + s * KC, zr, zi, wr, wi, gr, gi);
                const double4 sh = Ss[s];
                const double dud = (double)du;
                dbx = fma(dud, sh.x, dbx);
                dby = fma(dud, sh.y, dby);
                dbz = fma(dud, sh.z, dbz);
                zr = nzr;
                zi = nzi;
                wr = nwr;
                wi = nwi;
            }
        } else {
#pragma unroll 1
            for (int s = 0; s < SRC_TILE; ++s) {
                const double4 sh = Ss[s];
                const double u = fma(bx, sh.x, fma(by, sh.y, bz * sh.z));
                T du = 0;
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    T zr, zi;
                    channel_cis(u, kf[k], zr, zi);
                    du += As[s * KC + k] * (zr * gi[k] - zi * gr[k]);
                }
                const double dud = (double)du;
                dbx = fma(dud, sh.x, dbx);
                dby = fma(dud, sh.y, dby);
                dbz = fma(dud, sh.z, dbz);
            }
        }
        __syncthreads();
    }
    if (valid) {
        const double scale = sgn_over_c * 6.283185307179586476925;
        double* o = dblpart + (((size_t)blockIdx.z * nchunk + chunk) * nbl + b) * 4;
        o[0] = scale * dbx;
        o[1] = scale * dby;
        o[2] = scale * dbz;
        o[3] = 0.0;
    }
}

// -------------------------------------------------------------------------------------
// launchers
// -------------------------------------------------------------------------------------
// baselines per CTA: 128 unless a ragged baseline count would idle more than ~4% of the lanes
// (e.g. the 63 unique baselines of HERA-37), then 64 or 32
static inline int pick_bl_block(int nbl) {
    auto padded = [&](int n) { return ((nbl + n - 1) / n) * n; };
    const double best = (double)padded(32);
    if (padded(128) <= 1.04 * best) return 128;
    if (padded(64) <= 1.04 * best) return 64;
    return 32;
}

template <typename T>
int launch_fwd(const T* A, const double* shat, const double* blv, const double* freqs,
               const int* units, int nunits, int nbl, int nfreq, long long S, int conj, int uniform,
               T* vpart, cudaStream_t st) {
    if (nunits <= 0 || nbl <= 0 || nfreq <= 0) return 0;
    if (S % SRC_PAD) return set_error("fringe_sum_fwd: S must be a multiple of 128");
    constexpr int KC = Cfg<T>::KC;
    const int nchunk = (nfreq + KC - 1) / KC;
    if (nchunk > 65535 || nunits > 65535) return set_error("fringe_sum_fwd: grid too large");
    const int nthr = pick_bl_block(nbl);
    dim3 grid((nbl + nthr - 1) / nthr, nchunk, nunits);
    const double sgn_over_c = (conj ? -1.0 : 1.0) / C_LIGHT;
    const int smem = FwdSmem<T>::TOTAL;
    static DeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(fringe_sum_fwd_kernel<T, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(fringe_sum_fwd_kernel<T, false>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    if (uniform)
        fringe_sum_fwd_kernel<T, true><<<grid, nthr, smem, st>>>(
            A, shat, blv, freqs, reinterpret_cast<const int4*>(units), nbl, nfreq, S, sgn_over_c,
            vpart);
    else
        fringe_sum_fwd_kernel<T, false><<<grid, nthr, smem, st>>>(
            A, shat, blv, freqs, reinterpret_cast<const int4*>(units), nbl, nfreq, S, sgn_over_c,
            vpart);
    return check_launch("fringe_sum_fwd");
}

template <typename T>
int launch_reduce(const T* vpart, const int* ubeg, int nt, int nbl, int nfreq, T* V, long long sb,
                  long long stt, long long sf, double are, double aim, int accumulate,
                  cudaStream_t st) {
    if (nt <= 0 || nbl <= 0 || nfreq <= 0) return 0;
    constexpr int KC = Cfg<T>::KC;
    const int nfp = ((nfreq + KC - 1) / KC) * KC;
    const long long n = (long long)nbl * nfreq;
    dim3 grid((unsigned)((n + 255) / 256), nt);
    reduce_units_kernel<T><<<grid, 256, 0, st>>>(vpart, ubeg, nbl, nfreq, nfp, V, sb, stt, sf, are,
                                                  aim, accumulate);
    return check_launch("reduce_units");
}

template <typename T>
int launch_reduce_chisq(const T* vpart, const int* ubeg, int nt, int nbl, int nfreq, T* V, const T* D,
                        const T* W, long long sb, long long stt, long long sf, int accumulate,
                        double* chi_part, cudaStream_t st) {
    if (nt <= 0 || nbl <= 0 || nfreq <= 0) return 0;
    if (D == nullptr || chi_part == nullptr) return set_error("reduce_units_chisq: data and partials are required");
    constexpr int KC = Cfg<T>::KC;
    const int nfp = ((nfreq + KC - 1) / KC) * KC;
    const long long n = (long long)nbl * nfreq;
    dim3 grid((unsigned)((n + 255) / 256), nt);
    reduce_units_chisq_kernel<T><<<grid, 256, 0, st>>>(vpart, ubeg, nbl, nfreq, nfp, V, D, W, sb, stt,
                                                        sf, accumulate, chi_part);
    return check_launch("reduce_units_chisq");
}

template <typename T>
int launch_bwd_sky(const T* Gp, const double* shat, const double* blv, const double* freqs,
                   const int* tile_time, int nbl, int nt, int nfreq, long long S, int conj,
                   int uniform, T* dA, cudaStream_t st) {
    if (S <= 0 || nfreq <= 0) return 0;
    if (S % SRC_PAD) return set_error("fringe_sum_bwd_sky: S must be a multiple of 128");
    constexpr int KC = Cfg<T>::KC;
    const int nchunk = (nfreq + KC - 1) / KC;
    if (nchunk > 65535) return set_error("fringe_sum_bwd_sky: grid too large");
    dim3 grid((unsigned)(S / SKY_THREADS), nchunk);
    const double sgn_over_c = (conj ? -1.0 : 1.0) / C_LIGHT;
    const int smem = SkySmem<T>::TOTAL;
    static DeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(fringe_sum_bwd_sky_kernel<T, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(fringe_sum_bwd_sky_kernel<T, false>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    if (uniform)
        fringe_sum_bwd_sky_kernel<T, true><<<grid, SKY_THREADS, smem, st>>>(
            Gp, shat, blv, freqs, tile_time, nbl, nt, nfreq, S, sgn_over_c, dA);
    else
        fringe_sum_bwd_sky_kernel<T, false><<<grid, SKY_THREADS, smem, st>>>(
            Gp, shat, blv, freqs, tile_time, nbl, nt, nfreq, S, sgn_over_c, dA);
    return check_launch("fringe_sum_bwd_sky");
}

template <typename T>
int launch_bwd_bl(const T* Gp, const T* A, const double* shat, const double* blv,
                  const double* freqs, const int* units, int nunits, int nbl, int nt, int nfreq,
                  long long S, int conj, int uniform, double* dblpart, cudaStream_t st) {
    if (nunits <= 0 || nbl <= 0 || nfreq <= 0) return 0;
    if (S % SRC_PAD) return set_error("fringe_sum_bwd_bl: S must be a multiple of 128");
    constexpr int KC = Cfg<T>::KC;
    const int nchunk = (nfreq + KC - 1) / KC;
    if (nchunk > 65535 || nunits > 65535) return set_error("fringe_sum_bwd_bl: grid too large");
    const int nthr = pick_bl_block(nbl);
    dim3 grid((nbl + nthr - 1) / nthr, nchunk, nunits);
    const double sgn_over_c = (conj ? -1.0 : 1.0) / C_LIGHT;
    const int smem = FwdSmem<T>::TOTAL;
    static DeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(fringe_sum_bwd_bl_kernel<T, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(fringe_sum_bwd_bl_kernel<T, false>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    if (uniform)
        fringe_sum_bwd_bl_kernel<T, true><<<grid, nthr, smem, st>>>(
            Gp, A, shat, blv, freqs, reinterpret_cast<const int4*>(units), nbl, nt, nfreq, S,
            sgn_over_c, dblpart);
    else
        fringe_sum_bwd_bl_kernel<T, false><<<grid, nthr, smem, st>>>(
            Gp, A, shat, blv, freqs, reinterpret_cast<const int4*>(units), nbl, nt, nfreq, S,
            sgn_over_c, dblpart);
    return check_launch("fringe_sum_bwd_bl");
}

}  // namespace b200rime

using namespace b200rime;

extern "C" {

int b200rime_fringe_sum_fwd_f32(const float* A, const double* shat, const double* blv,
                                const double* freqs, const int* units, int nunits, int nbl,
                                int nfreq, long long S, int conj, int uniform, float* Vpart,
                                void* stream) {
    return launch_fwd<float>(A, shat, blv, freqs, units, nunits, nbl, nfreq, S, conj, uniform, Vpart,
                             (cudaStream_t)stream);
}
int b200rime_fringe_sum_fwd_f64(const double* A, const double* shat, const double* blv,
                                const double* freqs, const int* units, int nunits, int nbl,
                                int nfreq, long long S, int conj, int uniform, double* Vpart,
                                void* stream) {
    return launch_fwd<double>(A, shat, blv, freqs, units, nunits, nbl, nfreq, S, conj, uniform,
                              Vpart, (cudaStream_t)stream);
}
int b200rime_reduce_units_f32(const float* Vpart, const int* ubeg, int nt, int nbl, int nfreq,
                              float* V, long long sb, long long st, long long sf, double are,
                              double aim, int accumulate, void* stream) {
    return launch_reduce<float>(Vpart, ubeg, nt, nbl, nfreq, V, sb, st, sf, are, aim, accumulate,
                                (cudaStream_t)stream);
}
int b200rime_chisq_blocks(int nbl, int nfreq) {
    return (int)(((long long)nbl * nfreq + 255) / 256);
}
int b200rime_reduce_units_chisq_f32(const float* Vpart, const int* ubeg, int nt, int nbl, int nfreq,
                                    float* V, const float* D, const float* W, long long sb,
                                    long long st, long long sf, int accumulate, double* chi_part,
                                    void* stream) {
    return launch_reduce_chisq<float>(Vpart, ubeg, nt, nbl, nfreq, V, D, W, sb, st, sf, accumulate,
                                      chi_part, (cudaStream_t)stream);
}
int b200rime_reduce_units_chisq_f64(const double* Vpart, const int* ubeg, int nt, int nbl,
                                    int nfreq, double* V, const double* D, const double* W,
                                    long long sb, long long st, long long sf, int accumulate,
                                    double* chi_part, void* stream) {
    return launch_reduce_chisq<double>(Vpart, ubeg, nt, nbl, nfreq, V, D, W, sb, st, sf, accumulate,
                                       chi_part, (cudaStream_t)stream);
}
int b200rime_reduce_units_f64(const double* Vpart, const int* ubeg, int nt, int nbl, int nfreq,
                              double* V, long long sb, long long st, long long sf, double are,
                              double aim, int accumulate, void* stream) {
    return launch_reduce<double>(Vpart, ubeg, nt, nbl, nfreq, V, sb, st, sf, are, aim, accumulate,
                                 (cudaStream_t)stream);
}
int b200rime_fringe_sum_bwd_sky_f32(const float* Gp, const double* shat, const double* blv,
                                    const double* freqs, const int* tile_time, int nbl, int nt,
                                    int nfreq, long long S, int conj, int uniform, float* dA,
                                    void* stream) {
    return launch_bwd_sky<float>(Gp, shat, blv, freqs, tile_time, nbl, nt, nfreq, S, conj, uniform,
                                 dA, (cudaStream_t)stream);
}
int b200rime_fringe_sum_bwd_sky_f64(const double* Gp, const double* shat, const double* blv,
                                    const double* freqs, const int* tile_time, int nbl, int nt,
                                    int nfreq, long long S, int conj, int uniform, double* dA,
                                    void* stream) {
    return launch_bwd_sky<double>(Gp, shat, blv, freqs, tile_time, nbl, nt, nfreq, S, conj, uniform,
                                  dA, (cudaStream_t)stream);
}
int b200rime_fringe_sum_bwd_bl_f32(const float* Gp, const float* A, const double* shat,
                                   const double* blv, const double* freqs, const int* units,
                                   int nunits, int nbl, int nt, int nfreq, long long S, int conj,
                                   int uniform, double* dblpart, void* stream) {
    return launch_bwd_bl<float>(Gp, A, shat, blv, freqs, units, nunits, nbl, nt, nfreq, S, conj,
                                uniform, dblpart, (cudaStream_t)stream);
}
int b200rime_fringe_sum_bwd_bl_f64(const double* Gp, const double* A, const double* shat,
                                   const double* blv, const double* freqs, const int* units,
                                   int nunits, int nbl, int nt, int nfreq, long long S, int conj,
                                   int uniform, double* dblpart, void* stream) {
    return launch_bwd_bl<double>(Gp, A, shat, blv, freqs, units, nunits, nbl, nt, nfreq, S, conj,
                                 uniform, dblpart, (cudaStream_t)stream);
}

}  // extern "C"
