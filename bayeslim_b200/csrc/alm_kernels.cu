// Spherical-harmonic forward model a_lm -> beam map on the tensor cores (SURVEY section 8(f) row
// f2): AlmModel.forward_alm (sph_harm.py:1289-1373, "...i,ij->...j") behind YlmResponse.forward
// (beam_model.py:1166-1233), and its adjoint to the coefficients.
//
//   out[m][n] = sum_k X[m][k] Y[n][k]        complex, float32-grade
//
// forward:  m = (pol, vec, model, channel) rows of the coefficient tensor, k = (l, m) mode,
//           n = pixel of the beam map, Y = Ylm^T (constant: packed once per set of angles);
// adjoint:  m as before, k = pixel, n = mode, X = dL/d(map) (real for a real beam), Y = conj(Ylm).
//
// Same arithmetic as the tensor-core fringe kernels (tc_common.cuh): every real operand is a
// float16 pair hi + lo, a real product is three tcgen05.mma (hi.hi + hi.lo + lo.hi), the real and
// the imaginary part of Y are stacked so that one stage of 16 k is six MMAs of shape
// 128 x 256 x 16, TMEM chains of TC_FLUSH stages are added to register accumulators with
// round-to-nearest.  The difference is where the operands come from: here they exist in HBM, so a
// pack pass (elementwise, HBM-bound; once per set of angles for Ylm, once per call for the small
// coefficient / cotangent operand) writes them split, scaled by a power of two, and already in
// the UMMA canonical K-major order, one CONTIGUOUS block per (row block, stage):
//
//   Aq[mblk][kstage][ Xr_hi | Xr_lo | Xi_hi | Xi_lo ]            4 x 4 KB   (TcSmem::XR_H ...)
//   Bq[nblk][kstage][ (Yr ; Yi ; -Yr)_hi | (Yr ; Yi ; -Yr)_lo ]  2 x 12 KB  (TcSmem::B_H, B_L)
//
// so that a 40 KB pipeline stage is two 1-D TMA bulk copies (no tensor map, no swizzle) issued by
// the control warp, which also issues the MMAs; all 16 warps read the accumulators out.  The MMA
// pairing of tc_issue_stage computes conj(X) Y; the pack pass stores -Im X to get X Y.
//
// grid = (mblk * nblk, ksplit), row blocks fastest: CTAs that run together share their Y tile
// (the big operand) through L2.  ksplit > 1 writes per-split partial tiles that
// alm_reduce_kernel sums in a fixed order (deterministic; no atomics).
#include "tc_common.cuh"

namespace b200rime {

constexpr int ALM_A_STAGE = 4 * TcSmem::ARR;          // 16 KB
constexpr int ALM_B_STAGE = 2 * TcSmem::BBUF;         // 24 KB
static_assert(ALM_A_STAGE + ALM_B_STAGE == TcSmem::STAGE, "stage = A block + B block");
static_assert(TcSmem::B_H == ALM_A_STAGE, "B block follows the A block");

// ---------------------------------------------------------------------------------------
// pack: thread <-> (row, group of 8 k).  Source element (row r, k) at re[r * sr + k * sk]
// (and im[...]; im == nullptr: real operand).  Rows >= nrows and k >= K are zero.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void alm_load8(const float* __restrict__ p, long long sk, int k0, int K,
                                          float sc, float (&v)[8]) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (k0 + e < K) ? __ldg(p + (long long)(k0 + e) * sk) * sc : 0.f;
}

// which = 0: A operand (Xr, Xi split into four arrays); 1: B operand (three halves, hi and lo)
__global__ void __launch_bounds__(256)
alm_pack_kernel(const float* __restrict__ re, const float* __restrict__ im, long long sr,
                long long sk, int nrows, int K, const float* __restrict__ scale, int negate_im,
                int which, unsigned char* __restrict__ out) {
    const int nkg = ((K + TC_KS - 1) / TC_KS) * 2;                 // groups of 8 k, padded
    const int nkst = nkg / 2;
    const int rows_pad = ((nrows + TC_M - 1) / TC_M) * TC_M;
    const long long total = (long long)rows_pad * nkg;
    const float sc = __ldg(scale);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        // rows fastest: eight neighbouring threads write one 128-byte core matrix
        const int r = (int)(idx % rows_pad), kg = (int)(idx / rows_pad);
        const int blk = r / TC_M, row = r % TC_M, kst = kg >> 1;
        float vr[8], vi[8];
        if (r < nrows) {
            alm_load8(re + (long long)r * sr, sk, 8 * kg, K, sc, vr);
            if (im != nullptr) alm_load8(im + (long long)r * sr, sk, 8 * kg, K, negate_im ? -sc : sc, vi);
        }
        if (r >= nrows) {
#pragma unroll
            for (int e = 0; e < 8; ++e) vr[e] = 0.f;
        }
        if (r >= nrows || im == nullptr) {
#pragma unroll
            for (int e = 0; e < 8; ++e) vi[e] = 0.f;
        }
        uint4 rh, rl, ih, il;
        split8(vr, rh, rl);
        split8(vi, ih, il);
        const int roff = (row >> 3) * 256 + (kg & 1) * 128 + (row & 7) * 16;
        if (which == 0) {
            unsigned char* d = out + ((size_t)blk * nkst + kst) * ALM_A_STAGE + roff;
            *reinterpret_cast<uint4*>(d + TcSmem::XR_H) = rh;
            *reinterpret_cast<uint4*>(d + TcSmem::XR_L) = rl;
            *reinterpret_cast<uint4*>(d + TcSmem::XI_H) = ih;
            *reinterpret_cast<uint4*>(d + TcSmem::XI_L) = il;
        } else {
            unsigned char* d = out + ((size_t)blk * nkst + kst) * ALM_B_STAGE + roff;
            *reinterpret_cast<uint4*>(d) = rh;
            *reinterpret_cast<uint4*>(d + TcSmem::ARR) = ih;
            *reinterpret_cast<uint4*>(d + 2 * TcSmem::ARR) = neg_half8(rh);
            *reinterpret_cast<uint4*>(d + TcSmem::BBUF) = rl;
            *reinterpret_cast<uint4*>(d + TcSmem::BBUF + TcSmem::ARR) = il;
            *reinterpret_cast<uint4*>(d + TcSmem::BBUF + 2 * TcSmem::ARR) = neg_half8(rl);
        }
    }
}

// the MMAs of one stage; a_real: the imaginary part of X is zero (three MMAs instead of six)
__device__ __forceinline__ void alm_issue_stage(const TcIssue& q, int stage, int set, bool first,
                                                bool a_real) {
    if (!a_real) {
        tc_issue_stage(q, stage, set, first);
        return;
    }
    const uint32_t d = q.tmem + set * TC_SET_COLS;
    const uint32_t b = q.smem_base + stage * TcSmem::STAGE;
    const uint64_t arh = umma_desc_kmajor(b + TcSmem::XR_H), arl = umma_desc_kmajor(b + TcSmem::XR_L),
                   bph = umma_desc_kmajor(b + TcSmem::B_H + q.poff),
                   bpl = umma_desc_kmajor(b + TcSmem::B_L + q.poff);
    umma_f16(d, arh, bph, q.idesc, first ? 0u : 1u);
    umma_f16(d, arh, bpl, q.idesc, 1u);
    umma_f16(d, arl, bph, q.idesc, 1u);
}

// -------------------------------------------------------------------------------------
// GEMM.  grid = (mblk * nblk, ksplit), block = 512.
//   out   float [ksplit or 1][M][ldo] (real_out) or float2 [..][M][ldo]; split y writes plane y
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
alm_cgemm_kernel(const unsigned char* __restrict__ Aq, const unsigned char* __restrict__ Bq, int M,
                 int N, int nkst, int mblk, int a_real, int real_out,
                 const float* __restrict__ scale_a, const float* __restrict__ scale_b,
                 float* __restrict__ out, long long ldo) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bm = blockIdx.x % mblk, bn = blockIdx.x / mblk;
    const int ks0 = (int)((long long)nkst * blockIdx.y / gridDim.y);
    const int ks1 = (int)((long long)nkst * (blockIdx.y + 1) / gridDim.y);
    const int nst = ks1 - ks0;
    const int nchain = (nst + TC_FLUSH - 1) / TC_FLUSH;

    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TcSmem::BAR_OFF);
    uint64_t* empty = full + TC_NSTAGE;
    uint64_t* tfull = empty + TC_NSTAGE + TC_NSRC;          // same slots as the fringe kernels
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + TcSmem::TMEM_OFF);

    if (tid == 0) {
        for (int st = 0; st < TC_NSTAGE; ++st) {
            mbar_init(&full[st], 1);                  // one expect_tx arrival + the copied bytes
            mbar_init(&empty[st], 1);
        }
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
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int q = warp & 3, cg = warp >> 2;
    const int role = __shfl_sync(0xffffffffu, warp, 0);
    const uint32_t ta0 = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(TC_CG * cg);
    const uint32_t im_col = (uint32_t)TC_NMAX;
    float aR[TC_CG], aI[TC_CG];
#pragma unroll
    for (int c = 0; c < TC_CG; ++c) aR[c] = aI[c] = 0.f;
    int next_read = 0;

    if (role == TC_CTRL_WARP && nst > 0) {
        TcIssue iq;
        iq.tmem = tmem;
        iq.idesc = umma_idesc_f16(2 * TC_NMAX);
        iq.smem_base = smem_u32(smem);
        iq.poff = 0;
        iq.moff = (uint32_t)((TC_NMAX >> 3) * 256);
        const unsigned char* a_src = Aq + ((size_t)bm * nkst + ks0) * ALM_A_STAGE;
        const unsigned char* b_src = Bq + ((size_t)bn * nkst + ks0) * ALM_B_STAGE;
        auto load_stage = [&](int it) {
            const int stage = it % TC_NSTAGE;
            unsigned char* dst = smem + stage * TcSmem::STAGE;
            mbar_expect_tx(&full[stage], TcSmem::STAGE);
            bulk_g2s(dst, a_src + (size_t)it * ALM_A_STAGE, ALM_A_STAGE, &full[stage]);
            bulk_g2s(dst + TcSmem::B_H, b_src + (size_t)it * ALM_B_STAGE, ALM_B_STAGE, &full[stage]);
        };
        if (elect_one())
            for (int it = 0; it < min(nst, TC_NSTAGE); ++it) load_stage(it);
        __syncwarp();
        for (int it = 0; it < nst; ++it) {
            const int stage = it % TC_NSTAGE, chain = it / TC_FLUSH, set = chain & 1;
            const bool first = (it % TC_FLUSH) == 0;
            if (first) {
                while (next_read <= chain - 2)
                    tc_read_chain(tfull, tempty, next_read++, ta0, im_col, aR, aI, lane);
                if (chain >= 2)
                    mbar_wait_bounded(&tempty[set], (uint32_t)(((chain >> 1) - 1) & 1));
            }
            mbar_wait_bounded(&full[stage], (uint32_t)((it / TC_NSTAGE) & 1));
            tc_fence_after();
            if (elect_one()) {
                alm_issue_stage(iq, stage, set, first, a_real != 0);
                umma_commit(&empty[stage]);
                if ((it % TC_FLUSH) == TC_FLUSH - 1 || it == nst - 1) umma_commit(&tfull[set]);
            }
            __syncwarp();
            // refill the stage of the PREVIOUS iteration: its MMAs finish while the ones just
            // issued keep the tensor pipe busy, so this wait does not drain the pipe
            if (it >= 1 && it - 1 + TC_NSTAGE < nst) {
                const int pst = (it - 1) % TC_NSTAGE;
                mbar_wait_bounded(&empty[pst], (uint32_t)(((it - 1) / TC_NSTAGE) & 1));
                if (elect_one()) load_stage(it - 1 + TC_NSTAGE);
                __syncwarp();
            }
        }
    }
    while (next_read < nchain) tc_read_chain(tfull, tempty, next_read++, ta0, im_col, aR, aI, lane);

    // ---- write the tile: thread <-> row, 32 consecutive columns
    {
        const int row = bm * TC_M + 32 * q + lane;
        const int col0 = bn * TC_NMAX + TC_CG * cg;
        if (row < M && col0 < N) {
            const float inv = 1.f / (__ldg(scale_a) * __ldg(scale_b));
            const size_t plane = (size_t)blockIdx.y * (size_t)M * (size_t)ldo;
            const int nc = min(TC_CG, N - col0);
            if (real_out) {
                float* o = out + plane + (size_t)row * ldo + col0;
#pragma unroll
                for (int c = 0; c < TC_CG; ++c)
                    if (c < nc) o[c] = aR[c] * inv;
            } else {
                float2* o = reinterpret_cast<float2*>(out) + plane + (size_t)row * ldo + col0;
#pragma unroll
                for (int c = 0; c < TC_CG; ++c)
                    if (c < nc) o[c] = make_float2(aR[c] * inv, aI[c] * inv);
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

// out[r][c] = sum_y part[y][r][c], y in index order; rows of ncols floats, ld floats apart
__global__ void __launch_bounds__(256)
alm_reduce_kernel(const float* __restrict__ part, int nrows, long long ncols, long long ld,
                  int ksplit, float* __restrict__ out) {
    const long long n = (long long)nrows * ncols;
    const size_t plane = (size_t)nrows * (size_t)ld;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const size_t o = (size_t)(i / ncols) * (size_t)ld + (size_t)(i % ncols);
        float acc = part[o];
        for (int y = 1; y < ksplit; ++y) acc += part[(size_t)y * plane + o];
        out[o] = acc;
    }
}


// -------------------------------------------------------------------------------------
// float64 sessions (complex128 parity, 1e-10): the same product on the FP64 pipes, straight from
// the strided operands.  64 x 64 output tile per CTA, 16 x 16 threads x (4 x 4) complex
// accumulators, k in steps of 16 through shared memory.  DFMA-bound; no tensor cores (FP64).
//   X element (m, k) at xr[m * sxm + k * sxk] (xi likewise, nullptr = real), Y element (n, k) at
//   yr[n * syn + k * syk]; conj_x / conj_y conjugate the operand.
// -------------------------------------------------------------------------------------
constexpr int F64_T = 64, F64_K = 16;
__global__ void __launch_bounds__(256)
alm_cgemm_f64_kernel(const double* __restrict__ xr, const double* __restrict__ xi, long long sxm,
                     long long sxk, const double* __restrict__ yr, const double* __restrict__ yi,
                     long long syn, long long syk, int M, int N, int K, double sgn_x, double sgn_y,
                     int real_out, double* __restrict__ out, long long ldo) {
    __shared__ double Xs[2][F64_K][F64_T + 1], Ys[2][F64_K][F64_T + 1];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int m0 = blockIdx.y * F64_T, n0 = blockIdx.x * F64_T;
    double aR[4][4], aI[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) aR[i][j] = aI[i][j] = 0.0;
    for (int k0 = 0; k0 < K; k0 += F64_K) {
        for (int e = threadIdx.x; e < F64_T * F64_K; e += 256) {
            // k fastest when the k stride is the short one, rows fastest otherwise
            const int kk = (sxk <= sxm) ? e % F64_K : e / F64_T, r = (sxk <= sxm) ? e / F64_K : e % F64_T;
            const bool ok = (m0 + r < M) && (k0 + kk < K);
            const long long o = (long long)(m0 + r) * sxm + (long long)(k0 + kk) * sxk;
            Xs[0][kk][r] = ok ? xr[o] : 0.0;
            Xs[1][kk][r] = (ok && xi != nullptr) ? sgn_x * xi[o] : 0.0;
        }
        for (int e = threadIdx.x; e < F64_T * F64_K; e += 256) {
            const int kk = (syk <= syn) ? e % F64_K : e / F64_T, r = (syk <= syn) ? e / F64_K : e % F64_T;
            const bool ok = (n0 + r < N) && (k0 + kk < K);
            const long long o = (long long)(n0 + r) * syn + (long long)(k0 + kk) * syk;
            Ys[0][kk][r] = ok ? yr[o] : 0.0;
            Ys[1][kk][r] = (ok && yi != nullptr) ? sgn_y * yi[o] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < F64_K; ++kk) {
            double xre[4], xim[4], yre[4], yim[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                xre[i] = Xs[0][kk][ty + 16 * i], xim[i] = Xs[1][kk][ty + 16 * i];
                yre[i] = Ys[0][kk][tx + 16 * i], yim[i] = Ys[1][kk][tx + 16 * i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    aR[i][j] = fma(xre[i], yre[j], fma(-xim[i], yim[j], aR[i][j]));
                    aI[i][j] = fma(xre[i], yim[j], fma(xim[i], yre[j], aI[i][j]));
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
            if (m >= M || n >= N) continue;
            if (real_out) {
                out[(size_t)m * ldo + n] = aR[i][j];
            } else {
                out[2 * ((size_t)m * ldo + n)] = aR[i][j];
                out[2 * ((size_t)m * ldo + n) + 1] = aI[i][j];
            }
        }
}

static int alm_pack(const float* re, const float* im, long long sr, long long sk, int nrows, int K,
                    const float* scale, int negate_im, int which, void* out, cudaStream_t st) {
    if (nrows <= 0 || K <= 0) return 0;
    if (re == nullptr || scale == nullptr || out == nullptr)
        return set_error("cgemm_pack: null operand");
    const long long rows_pad = ((nrows + TC_M - 1) / TC_M) * (long long)TC_M;
    const long long total = rows_pad * (((K + TC_KS - 1) / TC_KS) * 2);
    const long long want = (total + 255) / 256;
    const int grid = (int)(want < 148LL * 16 ? want : 148LL * 16);
    alm_pack_kernel<<<grid, 256, 0, st>>>(re, im, sr, sk, nrows, K, scale, negate_im, which,
                                          static_cast<unsigned char*>(out));
    return check_launch("cgemm_pack");
}

}  // namespace b200rime

extern "C" {

long long b200rime_cgemm_a_bytes(int M, int K) {
    using namespace b200rime;
    return (long long)((M + TC_M - 1) / TC_M) * ((K + TC_KS - 1) / TC_KS) * ALM_A_STAGE;
}
long long b200rime_cgemm_b_bytes(int N, int K) {
    using namespace b200rime;
    return (long long)((N + TC_NMAX - 1) / TC_NMAX) * ((K + TC_KS - 1) / TC_KS) * ALM_B_STAGE;
}
int b200rime_cgemm_pack_a_f32(const float* re, const float* im, long long stride_row,
                              long long stride_k, int M, int K, const float* scale, int negate_im,
                              void* Aq, void* stream) {
    return b200rime::alm_pack(re, im, stride_row, stride_k, M, K, scale, negate_im, 0, Aq,
                              (cudaStream_t)stream);
}
int b200rime_cgemm_pack_b_f32(const float* re, const float* im, long long stride_row,
                              long long stride_k, int N, int K, const float* scale, int negate_im,
                              void* Bq, void* stream) {
    return b200rime::alm_pack(re, im, stride_row, stride_k, N, K, scale, negate_im, 1, Bq,
                              (cudaStream_t)stream);
}
int b200rime_cgemm_f32(const void* Aq, const void* Bq, int M, int N, int K, int ksplit, int a_real,
                       int real_out, const float* scale_a, const float* scale_b, float* out,
                       long long ldo, float* part, void* stream) {
    using namespace b200rime;
    cudaStream_t st = (cudaStream_t)stream;
    if (M <= 0 || N <= 0) return 0;
    if (K <= 0) return set_error("cgemm: K must be positive");
    if (Aq == nullptr || Bq == nullptr || out == nullptr || scale_a == nullptr || scale_b == nullptr)
        return set_error("cgemm: null operand");
    const int nkst = (K + TC_KS - 1) / TC_KS;
    const int mblk = (M + TC_M - 1) / TC_M, nblk = (N + TC_NMAX - 1) / TC_NMAX;
    if (ksplit < 1 || ksplit > nkst || ksplit > 65535) return set_error("cgemm: 1 <= ksplit <= K / 16");
    if (ksplit > 1 && part == nullptr) return set_error("cgemm: ksplit > 1 needs the workspace");
    if (ldo < N) return set_error("cgemm: ldo < N");
    if ((long long)mblk * nblk > 2147483647LL) return set_error("cgemm: grid too large");
    static DeviceOnce attr_once;
    if (attr_once.first()) {
        if (cudaFuncSetAttribute(alm_cgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 TcSmem::TOTAL) != cudaSuccess)
            return set_error("cgemm: cannot reserve shared memory");
    }
    dim3 grid((unsigned)(mblk * nblk), (unsigned)ksplit);
    alm_cgemm_kernel<<<grid, TC_THREADS, TcSmem::TOTAL, st>>>(
        static_cast<const unsigned char*>(Aq), static_cast<const unsigned char*>(Bq), M, N, nkst,
        mblk, a_real, real_out, scale_a, scale_b, ksplit > 1 ? part : out, ldo);
    int rc = check_launch("cgemm");
    if (rc != 0 || ksplit == 1) return rc;
    const int w = real_out ? 1 : 2;
    const long long want = ((long long)M * N * w + 255) / 256;
    alm_reduce_kernel<<<(int)(want < 148LL * 16 ? want : 148LL * 16), 256, 0, st>>>(
        part, M, (long long)N * w, ldo * w, ksplit, out);
    return check_launch("cgemm reduce");
}

int b200rime_cgemm_f64(const double* xr, const double* xi, long long stride_xm, long long stride_xk,
                       const double* yr, const double* yi, long long stride_yn, long long stride_yk,
                       int M, int N, int K, int conj_x, int conj_y, int real_out, double* out,
                       long long ldo, void* stream) {
    using namespace b200rime;
    if (M <= 0 || N <= 0) return 0;
    if (K <= 0) return set_error("cgemm: K must be positive");
    if (xr == nullptr || yr == nullptr || out == nullptr) return set_error("cgemm: null operand");
    if (ldo < N) return set_error("cgemm: ldo < N");
    const int gy = (M + F64_T - 1) / F64_T;
    if (gy > 65535) return set_error("cgemm: grid too large");
    dim3 grid((N + F64_T - 1) / F64_T, gy);
    alm_cgemm_f64_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        xr, xi, stride_xm, stride_xk, yr, yi, stride_yn, stride_yk, M, N, K, conj_x ? -1.0 : 1.0,
        conj_y ? -1.0 : 1.0, real_out, out, ldo);
    return check_launch("cgemm_f64");
}

}  // extern "C"
