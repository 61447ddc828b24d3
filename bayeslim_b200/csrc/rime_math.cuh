// Per-thread inner loops of the fringe-sum kernels, written __host__ __device__ so that the
// exact same arithmetic (fp64 phase reduction, fp32 seeds, centre-seeded two-sided rotation
// recurrence) can be exercised on a CPU by csrc/tools/emulate.cu.
//
// float32 path: Blackwell's packed FP32 pipe (PTX fma/mul.rn.f32x2 -> SASS FFMA2/FMUL2).
// A complex number lives in one 64-bit register pair (re, im); ptxas folds the half swap, the
// +- sign pattern and scalar broadcasts into FFMA2/FMUL2 operand modifiers
// (.F32x2.LO_HI.NP, Rn.F32), so one source.baseline.channel evaluation is
//     t = swap(z) * (-wi, +wi)      FMUL2
//     z = z * (wr, wr) + t          FFMA2        (complex rotation z *= w)
//     acc += z * (a, a)             FFMA2        (multiply-accumulate)
// i.e. 3 issue slots for the same 6 FP32-pipe lane-cycles as the scalar form, which frees
// issue slots and register-file ports for the seed / shared-memory work.
#pragma once
#include <cstring>
#include "common.cuh"

namespace b200rime {

// ------------------------------------------------------------------------------------
// packed pair of floats in one 64-bit register
// ------------------------------------------------------------------------------------
struct P2 {
    unsigned long long v;
};
__host__ __device__ __forceinline__ P2 p2(float x, float y) {
    P2 r;
#ifdef __CUDA_ARCH__
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(x), "f"(y));
#else
    float t[2] = {x, y};
    std::memcpy(&r.v, t, 8);
#endif
    return r;
}
__host__ __device__ __forceinline__ void p2_get(P2 a, float& x, float& y) {
#ifdef __CUDA_ARCH__
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
#else
    float t[2];
    std::memcpy(t, &a.v, 8);
    x = t[0];
    y = t[1];
#endif
}
__host__ __device__ __forceinline__ P2 p2_swap(P2 a) {
    float x, y;
    p2_get(a, x, y);
    return p2(y, x);
}
__host__ __device__ __forceinline__ P2 p2_mul(P2 a, P2 b) {
#ifdef __CUDA_ARCH__
    P2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
#else
    float ax, ay, bx, by;
    p2_get(a, ax, ay);
    p2_get(b, bx, by);
    return p2(ax * bx, ay * by);
#endif
}
__host__ __device__ __forceinline__ P2 p2_fma(P2 a, P2 b, P2 c) {
#ifdef __CUDA_ARCH__
    P2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
#else
    float ax, ay, bx, by, cx, cy;
    p2_get(a, ax, ay);
    p2_get(b, bx, by);
    p2_get(c, cx, cy);
    return p2(fmaf(ax, bx, cx), fmaf(ay, by, cy));
#endif
}
// acc += a * b, accumulator updated in place (keeps ptxas from renaming accumulator pairs)
__host__ __device__ __forceinline__ void p2_mac(P2& acc, P2 a, P2 b) {
#ifdef __CUDA_ARCH__
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc.v) : "l"(a.v), "l"(b.v));
#else
    acc = p2_fma(a, b, acc);
#endif
}
// z *= w  /  z *= conj(w)   with W1 = (wr, wr), Wup = (-wi, +wi), Wdn = (+wi, -wi)
__host__ __device__ __forceinline__ P2 p2_rot(P2 z, P2 W1, P2 Wsgn) {
    return p2_fma(z, W1, p2_mul(p2_swap(z), Wsgn));
}

// 16-byte vector of T with element access
template <typename T> struct Vec16;
template <> struct Vec16<float> {
    static constexpr int N = 4;
    float4 v;
    __host__ __device__ __forceinline__ float get(int i) const {
        return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
    }
};
template <> struct Vec16<double> {
    static constexpr int N = 2;
    double2 v;
    __host__ __device__ __forceinline__ double get(int i) const { return i == 0 ? v.x : v.y; }
};

// ------------------------------------------------------------------------------------
// forward: acc[k] += a[k] * z_k, z_k = z_mid * w^(k-MID).  Two chains leave the chunk centre
// in opposite directions so that the error of w is amplified by at most KC/2 steps.
// ------------------------------------------------------------------------------------
template <typename T, int KC> struct FwdTile {          // generic (float64, host reference)
    T r[KC], i[KC];
    __host__ __device__ __forceinline__ void zero() {
#pragma unroll
        for (int k = 0; k < KC; ++k) r[k] = i[k] = 0;
    }
    __host__ __device__ __forceinline__ void get(int k, T& re, T& im) const {
        re = r[k];
        im = i[k];
    }
    __host__ __device__ __forceinline__ void mac(int k, T a, T zr, T zi) {
        r[k] += a * zr;
        i[k] += a * zi;
    }
    __host__ __device__ __forceinline__ void accumulate(const T* __restrict__ a, T zr, T zi, T wr,
                                                        T wi) {
        constexpr int MID = KC / 2;
        constexpr int N = Vec16<T>::N;
        T yr = zr, yi = zi;
        rotc(yr, yi, wr, wi);  // channel MID-1
        const Vec16<T>* av = reinterpret_cast<const Vec16<T>*>(a);
#pragma unroll
        for (int j = 0; j < MID; j += N) {
            Vec16<T> up = av[(MID + j) / N];
            Vec16<T> dn = av[(MID - N - j) / N];
#pragma unroll
            for (int q = 0; q < N; ++q) {
                T au = up.get(q);
                T ad = dn.get(N - 1 - q);
                r[MID + j + q] += au * zr;
                i[MID + j + q] += au * zi;
                rot(zr, zi, wr, wi);
                r[MID - 1 - j - q] += ad * yr;
                i[MID - 1 - j - q] += ad * yi;
                rotc(yr, yi, wr, wi);
            }
        }
    }
};
template <int KC> struct FwdTile<float, KC> {           // packed FP32x2
    P2 v[KC];
    __host__ __device__ __forceinline__ void zero() {
#pragma unroll
        for (int k = 0; k < KC; ++k) v[k] = p2(0.f, 0.f);
    }
    __host__ __device__ __forceinline__ void get(int k, float& re, float& im) const {
        p2_get(v[k], re, im);
    }
    __host__ __device__ __forceinline__ void mac(int k, float a, float zr, float zi) {
        v[k] = p2_fma(p2(zr, zi), p2(a, a), v[k]);
    }
    __host__ __device__ __forceinline__ void accumulate(const float* __restrict__ a, float zr,
                                                        float zi, float wr, float wi) {
        constexpr int MID = KC / 2;
        const P2 W1 = p2(wr, wr), Wup = p2(-wi, wi), Wdn = p2(wi, -wi);
        P2 z = p2(zr, zi);
        P2 y = p2_rot(z, W1, Wdn);  // channel MID-1
        const Vec16<float>* av = reinterpret_cast<const Vec16<float>*>(a);
#pragma unroll
        for (int j = 0; j < MID; j += 4) {
            Vec16<float> up = av[(MID + j) / 4];
            Vec16<float> dn = av[(MID - 4 - j) / 4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float au = up.get(q);
                const float ad = dn.get(3 - q);
                p2_mac(v[MID + j + q], z, p2(au, au));
                z = p2_rot(z, W1, Wup);
                p2_mac(v[MID - 1 - j - q], y, p2(ad, ad));
                y = p2_rot(y, W1, Wdn);
            }
        }
    }
};

// ------------------------------------------------------------------------------------
// backward to sky: acc[k] += Re(conj(z_k) * G_k) = zr*Gr + zi*Gi
// g: KC interleaved complex values (re, im) of this baseline's cotangent row.
// ------------------------------------------------------------------------------------
template <typename T, int KC> struct SkyTile {
    T acc[KC];
    __host__ __device__ __forceinline__ void zero() {
#pragma unroll
        for (int k = 0; k < KC; ++k) acc[k] = 0;
    }
    __host__ __device__ __forceinline__ T value(int k) const { return acc[k]; }
    __host__ __device__ __forceinline__ void mac(int k, T zr, T zi, T gr, T gi) {
        acc[k] += zr * gr + zi * gi;
    }
    __host__ __device__ __forceinline__ void accumulate(const T* __restrict__ g, T zr, T zi, T wr,
                                                        T wi) {
        constexpr int MID = KC / 2;
        constexpr int N = Vec16<T>::N;
        constexpr int NC = N / 2;
        T yr = zr, yi = zi;
        rotc(yr, yi, wr, wi);
        const Vec16<T>* gv = reinterpret_cast<const Vec16<T>*>(g);
#pragma unroll
        for (int j = 0; j < MID; j += NC) {
            Vec16<T> up = gv[(MID + j) / NC];
            Vec16<T> dn = gv[(MID - NC - j) / NC];
#pragma unroll
            for (int q = 0; q < NC; ++q) {
                T ur = up.get(2 * q), ui = up.get(2 * q + 1);
                T dr = dn.get(2 * (NC - 1 - q)), di = dn.get(2 * (NC - 1 - q) + 1);
                acc[MID + j + q] += zr * ur + zi * ui;
                rot(zr, zi, wr, wi);
                acc[MID - 1 - j - q] += yr * dr + yi * di;
                rotc(yr, yi, wr, wi);
            }
        }
    }
};
template <int KC> struct SkyTile<float, KC> {   // packed rotation, scalar real accumulators
    float acc[KC];
    __host__ __device__ __forceinline__ void zero() {
#pragma unroll
        for (int k = 0; k < KC; ++k) acc[k] = 0.f;
    }
    __host__ __device__ __forceinline__ float value(int k) const { return acc[k]; }
    __host__ __device__ __forceinline__ void mac(int k, float zr, float zi, float gr, float gi) {
        acc[k] = fmaf(zi, gi, fmaf(zr, gr, acc[k]));
    }
    __host__ __device__ __forceinline__ void accumulate(const float* __restrict__ g, float zr,
                                                        float zi, float wr, float wi) {
        constexpr int MID = KC / 2;
        const P2 W1 = p2(wr, wr), Wup = p2(-wi, wi), Wdn = p2(wi, -wi);
        P2 z = p2(zr, zi);
        P2 y = p2_rot(z, W1, Wdn);
        const Vec16<float>* gv = reinterpret_cast<const Vec16<float>*>(g);
#pragma unroll
        for (int j = 0; j < MID; j += 2) {
            Vec16<float> up = gv[(MID + j) / 2];
            Vec16<float> dn = gv[(MID - 2 - j) / 2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                float ar, ai, br, bi;
                p2_get(z, ar, ai);
                p2_get(y, br, bi);
                mac(MID + j + q, ar, ai, up.get(2 * q), up.get(2 * q + 1));
                z = p2_rot(z, W1, Wup);
                mac(MID - 1 - j - q, br, bi, dn.get(2 * (1 - q)), dn.get(2 * (1 - q) + 1));
                y = p2_rot(y, W1, Wdn);
            }
        }
    }
};

// ------------------------------------------------------------------------------------
// backward to baseline vectors: du = sum_k a[k] * Im(conj(z_k) * G'_k)
// gr/gi: this thread's pre-scaled cotangent G'_k = nu_k * G_k (registers).
// ------------------------------------------------------------------------------------
template <typename T, int KC>
__host__ __device__ __forceinline__ T bl_accumulate(const T* __restrict__ a, T zr, T zi, T wr, T wi,
                                                    const T* __restrict__ gr,
                                                    const T* __restrict__ gi) {
    constexpr int MID = KC / 2;
    constexpr int N = Vec16<T>::N;
    T yr = zr, yi = zi;
    rotc(yr, yi, wr, wi);
    T du_up = 0, du_dn = 0;
    const Vec16<T>* av = reinterpret_cast<const Vec16<T>*>(a);
#pragma unroll
    for (int j = 0; j < MID; j += N) {
        Vec16<T> up = av[(MID + j) / N];
        Vec16<T> dn = av[(MID - N - j) / N];
#pragma unroll
        for (int q = 0; q < N; ++q) {
            T au = up.get(q);
            T ad = dn.get(N - 1 - q);
            du_up += au * (zr * gi[MID + j + q] - zi * gr[MID + j + q]);
            rot(zr, zi, wr, wi);
            du_dn += ad * (yr * gi[MID - 1 - j - q] - yi * gr[MID - 1 - j - q]);
            rotc(yr, yi, wr, wi);
        }
    }
    return du_up + du_dn;
}
// float32: Horner form.  For one (baseline, source) the K3 output is a single scalar,
//   du = Im( conj(z_mid) * sum_k c_k v^(k-MID) ),  c_k = a_k G'_k,  v = conj(w),
// i.e. a polynomial in the unit complex v.  Horner evaluation folds the accumulation into the
// complex multiply, H <- H*v + c_k = FFMA2(swap(H), (-vi, vi), c_k) then FFMA2(H, (vr, vr), .),
// so an evaluation costs 3 packed ops (c_k: FMUL2 with a scalar-broadcast operand) instead of
// rotation (2) + projection (3 scalar).  Two chains (channels >= MID in powers of v, channels
// < MID in powers of conj(v)) keep the amplification of v's angle error at KC/2 steps.
// g2[k] = (Gr'_k, Gi'_k) register pairs.
template <int KC>
__host__ __device__ __forceinline__ float bl_accumulate_f32(const float* __restrict__ a, float zr,
                                                            float zi, float wr, float wi,
                                                            const P2* __restrict__ g2) {
    constexpr int MID = KC / 2;
    const P2 V1 = p2(wr, wr);
    const P2 Vup = p2(wi, -wi);      // multiply by v = conj(w) = (wr, -wi)
    const P2 Vdn = p2(-wi, wi);      // multiply by conj(v) = w
    const Vec16<float>* av = reinterpret_cast<const Vec16<float>*>(a);
    P2 Hu = p2(0.f, 0.f), Hd = p2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < MID; j += 4) {
        // up chain walks k = KC-1 ... MID ; down chain walks k = 0 ... MID-1
        Vec16<float> up = av[(KC - 4 - j) / 4];
        Vec16<float> dn = av[j / 4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int ku = KC - 1 - j - q;
            const int kd = j + q;
            const float au = up.get(3 - q);
            const float ad = dn.get(q);
            const P2 cu = p2_mul(g2[ku], p2(au, au));
            const P2 cd = p2_mul(g2[kd], p2(ad, ad));
            Hu = p2_fma(Hu, V1, p2_fma(p2_swap(Hu), Vup, cu));   // Hu = Hu*v + c_ku
            Hd = p2_fma(Hd, V1, p2_fma(p2_swap(Hd), Vdn, cd));   // Hd = Hd*w + c_kd
        }
    }
    // Hu = sum_{k>=MID} c_k v^(k-MID);  Hd = sum_{k<MID} c_k w^(MID-1-k)  ->  times w once more
    Hd = p2_fma(Hd, V1, p2_mul(p2_swap(Hd), Vdn));
    float sr, si, dr, di;
    p2_get(Hu, sr, si);
    p2_get(Hd, dr, di);
    sr += dr;
    si += di;
    return zr * si - zi * sr;        // Im(conj(z_mid) * S)
}

}  // namespace b200rime
