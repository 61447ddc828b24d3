// Per-thread inner loops of the fringe-sum kernels, written __host__ __device__ so that the
// exact same arithmetic (fp64 phase reduction, fp32 seeds, centre-seeded two-sided rotation
// recurrence) can be exercised on a CPU by csrc/emulate.cu.
#pragma once
#include "common.cuh"

namespace b200rime {

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

// ---- forward: acc[k] += a[k] * z_k, z_k = z_mid * w^(k-MID) ---------------------------------
// a: KC real values of this source (16-byte aligned).  Two chains leave the chunk centre in
// opposite directions so that the error of w is amplified by at most KC/2 steps.
template <typename T, int KC>
__host__ __device__ __forceinline__ void fwd_accumulate(const T* __restrict__ a, T zr, T zi, T wr,
                                                        T wi, T* __restrict__ accr,
                                                        T* __restrict__ acci) {
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
        for (int i = 0; i < N; ++i) {
            T au = up.get(i);
            T ad = dn.get(N - 1 - i);
            accr[MID + j + i] += au * zr;
            acci[MID + j + i] += au * zi;
            rot(zr, zi, wr, wi);
            accr[MID - 1 - j - i] += ad * yr;
            acci[MID - 1 - j - i] += ad * yi;
            rotc(yr, yi, wr, wi);
        }
    }
}

// ---- backward to sky: acc[k] += Re(conj(z_k) * G_k) = zr*Gr + zi*Gi ------------------------
// g: KC interleaved complex values (re, im) of this baseline's cotangent row.
template <typename T, int KC>
__host__ __device__ __forceinline__ void sky_accumulate(const T* __restrict__ g, T zr, T zi, T wr,
                                                        T wi, T* __restrict__ acc) {
    constexpr int MID = KC / 2;
    constexpr int N = Vec16<T>::N;      // reals per vector
    constexpr int NC = N / 2;           // complex per vector (2 for float, 1 for double)
    T yr = zr, yi = zi;
    rotc(yr, yi, wr, wi);
    const Vec16<T>* gv = reinterpret_cast<const Vec16<T>*>(g);
#pragma unroll
    for (int j = 0; j < MID; j += NC) {
        Vec16<T> up = gv[(MID + j) / NC];
        Vec16<T> dn = gv[(MID - NC - j) / NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            T ur = up.get(2 * i), ui = up.get(2 * i + 1);
            T dr = dn.get(2 * (NC - 1 - i)), di = dn.get(2 * (NC - 1 - i) + 1);
            acc[MID + j + i] += zr * ur + zi * ui;
            rot(zr, zi, wr, wi);
            acc[MID - 1 - j - i] += yr * dr + yi * di;
            rotc(yr, yi, wr, wi);
        }
    }
}

// ---- backward to baseline vectors: du = sum_k a[k] * Im(conj(z_k) * G'_k) ------------------
// gr/gi: this thread's pre-scaled cotangent G'_k = nu_k * G_k (registers).
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
        for (int i = 0; i < N; ++i) {
            T au = up.get(i);
            T ad = dn.get(N - 1 - i);
            du_up += au * (zr * gi[MID + j + i] - zi * gr[MID + j + i]);
            rot(zr, zi, wr, wi);
            du_dn += ad * (yr * gi[MID - 1 - j - i] - yi * gr[MID - 1 - j - i]);
            rotc(yr, yi, wr, wi);
        }
    }
    return du_up + du_dn;
}

}  // namespace b200rime
