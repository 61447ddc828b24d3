/*
 * b200rime.h -- C ABI of the B200-native RIME hot path.
 *
 * The reference (BayesLIM) is pure Python/PyTorch and has no FFI of its own; the
 * boundary these entry points replace is the chain of ATen calls made by
 *   rime_model.RIME._prod_and_sum            (bayeslim/rime_model.py:391-440)
 *   telescope_model.ArrayModel.gen_fringe    (bayeslim/telescope_model.py:310-358)
 *   beam_model.PixelBeam.apply_beam          (bayeslim/beam_model.py:273-372)
 *   utils.PixInterp.interp                   (bayeslim/utils.py:815-861)
 *   beam_model.airy_disk                     (bayeslim/beam_model.py:1418-1482)
 *   beam_model.cut_sky_fov                   (bayeslim/beam_model.py:1681-1698)
 * and their autograd backward.  Each function below names the reference lines it
 * stands in for.  INTEGRATION.md shows the ctypes binding a BayesLIM maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; no allocation happens
 *     inside the library (workspaces are passed in); no torch types cross the boundary
 *   - `stream` is a cudaStream_t (pass the caller's current stream, 0 = legacy default)
 *   - all functions return 0 on success; nonzero -> b200rime_last_error() has the text
 *   - suffix _f32 / _f64 selects the arithmetic of the kernel (complex64 / complex128
 *     visibilities); geometry (shat, blvecs, freqs) is always float64
 *   - complex arrays are interleaved (re, im) pairs of the real type
 *
 * Data layout ("tiled source layout")
 *   Sources of all times of a time-group are packed along one axis: time t owns the
 *   range [toff[t], toff[t+1]) whose length is Ns_t (sources inside the beam FOV at
 *   that time) rounded up to a multiple of b200rime_src_pad() (=128); padding entries
 *   carry zero intensity.  S = toff[Nt] is the packed length.  Frequencies are split
 *   into chunks of KC = b200rime_kc(is_f64) channels (64 for f32, 32 for f64);
 *   nchunk = ceil(Nf / KC), Nfp = nchunk*KC.
 *     A     : real  [nchunk][S][KC]   perceived sky  B_p conj(B_q) I  (one real plane)
 *     shat  : f64   [S][4]            unit vectors (x, y, z, 0), ENU
 *     blv   : f64   [Nbl][4]          baseline vectors (x, y, z, 0) in metres
 *     units : int32 [nunits][4]       {time index, s_begin, s_end, 0}; s_* are packed
 *                                     indices, multiples of 64, inside one time's range
 *     Vpart : cplx  [nunits][Nbl][Nfp] per-unit partial visibilities
 *     Gp    : cplx  [Nt][nchunk][Nbl][KC]  cotangent dL/dV, zero padded in frequency
 */
#ifndef B200RIME_H
#define B200RIME_H

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b200rime_stream_t;

/* library identification / diagnostics */
const char* b200rime_version(void);
const char* b200rime_last_error(void);
int b200rime_src_pad(void);
int b200rime_src_tile(void);
int b200rime_kc(int is_f64);
/* host query: fills sm_count, sm_clock_khz (max), compute capability major/minor */
int b200rime_device_info(int device, int* sm_count, int* clock_khz, int* cc_major, int* cc_minor);

/* ---- fringe sum forward --------------------------------------------------------
 * Vpart[u][b][f] = sum_{s in unit u} A[f][s] * exp(sgn * 2 pi i (blv[b] . shat[s]) freqs[f] / c)
 * sgn = +1 (conj == 0, RIME) or -1 (conj != 0, imaging).  Replaces gen_fringe
 * (telescope_model.py:351-356) + `fringe * psky` + torch.sum (rime_model.py:426-429);
 * the (Nbl, Nf, Ns) fringe tensor is never formed.  uniform != 0 asserts that freqs is
 * equally spaced inside every chunk (rotation recurrence); uniform == 0 evaluates every
 * channel's phase directly.  freqs: f64 [Nf] device. */
int b200rime_fringe_sum_fwd_f32(const float* A, const double* shat, const double* blv,
                                const double* freqs, const int* units, int nunits, int nbl,
                                int nfreq, long long S, int conj, int uniform, float* Vpart,
                                b200rime_stream_t stream);
int b200rime_fringe_sum_fwd_f64(const double* A, const double* shat, const double* blv,
                                const double* freqs, const int* units, int nunits, int nbl,
                                int nfreq, long long S, int conj, int uniform, double* Vpart,
                                b200rime_stream_t stream);

/* V[b*sb + t*st + f*sf] (+)= alpha * sum_{u in [ubeg[t], ubeg[t+1])} Vpart[u][b][f]
 * (fixed order, float64 accumulation -> bitwise reproducible).  Strides in complex
 * elements; alpha = (alpha_re, alpha_im); accumulate != 0 adds to V.  Stands in for
 * torch.stack(skyvis, dim=3) (rime_model.py:368). ubeg: int32 [Nt+1] device. */
int b200rime_reduce_units_f32(const float* Vpart, const int* ubeg, int nt, int nbl, int nfreq,
                              float* V, long long sb, long long st, long long sf,
                              double alpha_re, double alpha_im, int accumulate,
                              b200rime_stream_t stream);
int b200rime_reduce_units_f64(const double* Vpart, const int* ubeg, int nt, int nbl, int nfreq,
                              double* V, long long sb, long long st, long long sf,
                              double alpha_re, double alpha_im, int accumulate,
                              b200rime_stream_t stream);

/* ---- likelihood epilogue ----------------------------------------------------------
 * reduce_units fused with the Gaussian chi-square of optim.LogProb.forward_chisq
 * (optim.py:1012-1024: res = prediction - data, chisq = sum conj(res) res icov; apply_icov
 * optim.py:1836 with cov_axis None).  With V = (V if accumulate) + sum_u Vpart[u]:
 *   chi_part[t * chisq_blocks(nbl, nfreq) + block] = sum over the block's 256 (b, f) elements of
 *                                                    W |V - D|^2        (float64, fixed order)
 *   V[b*sb + t*st + f*sf] <- 2 W (V - D)    the cotangent dchisq/dV the backward kernels consume
 * so the visibilities themselves never reach HBM when only the likelihood is wanted.  D (complex)
 * and W (real, NULL = 1) use V's strides. */
int b200rime_chisq_blocks(int nbl, int nfreq);
int b200rime_reduce_units_chisq_f32(const float* Vpart, const int* ubeg, int nt, int nbl, int nfreq,
                                    float* V, const float* D, const float* W, long long sb,
                                    long long st, long long sf, int accumulate, double* chi_part,
                                    b200rime_stream_t stream);
int b200rime_reduce_units_chisq_f64(const double* Vpart, const int* ubeg, int nt, int nbl,
                                    int nfreq, double* V, const double* D, const double* W,
                                    long long sb, long long st, long long sf, int accumulate,
                                    double* chi_part, b200rime_stream_t stream);

/* ---- fringe sum backward to the perceived sky ----------------------------------
 * dA[f][s] = sum_b Re( conj(F[b][f][s]) * Gp[t(s)][f][b] )   (autograd of rime_model.py:429
 * w.r.t. psky, for a real plane).  tile_time: int32 [S/128] time index of each 128-source
 * tile.  Baselines are summed in index order by the thread that owns (s, chunk): no atomics. */
int b200rime_fringe_sum_bwd_sky_f32(const float* Gp, const double* shat, const double* blv,
                                    const double* freqs, const int* tile_time, int nbl, int nt,
                                    int nfreq, long long S, int conj, int uniform, float* dA,
                                    b200rime_stream_t stream);
int b200rime_fringe_sum_bwd_sky_f64(const double* Gp, const double* shat, const double* blv,
                                    const double* freqs, const int* tile_time, int nbl, int nt,
                                    int nfreq, long long S, int conj, int uniform, double* dA,
                                    b200rime_stream_t stream);

/* ---- fringe sum backward to the baseline vectors --------------------------------
 * dblpart[u][chunk][b][0..2] = sum_{s in u} shat[s] * sgn*(2 pi / c) *
 *                              sum_{f in chunk} freqs[f] A[f][s] Im( conj(F) Gp[t][f][b] )
 * (autograd of telescope_model.py:356 w.r.t. blvecs); float64 [nunits][nchunk][Nbl][4]. */
int b200rime_fringe_sum_bwd_bl_f32(const float* Gp, const float* A, const double* shat,
                                   const double* blv, const double* freqs, const int* units,
                                   int nunits, int nbl, int nt, int nfreq, long long S, int conj,
                                   int uniform, double* dblpart, b200rime_stream_t stream);
int b200rime_fringe_sum_bwd_bl_f64(const double* Gp, const double* A, const double* shat,
                                   const double* blv, const double* freqs, const int* units,
                                   int nunits, int nbl, int nt, int nfreq, long long S, int conj,
                                   int uniform, double* dblpart, b200rime_stream_t stream);

/* ---- layout conversion (generic beam responses) ---------------------------------
 * pack:   A[chunk][soff + s][k] = X[(chunk*KC + k)*ldx + s]   for s < ns, f < Nf; 0 elsewhere
 *         over the padded range [soff, soff + ns_pad).
 * unpack: X[f*ldx + s] = A[...]                                (the adjoint / inverse)
 * X is a row-major (Nf, ns) real plane, e.g. the perceived sky of beam_model.py:341. */
int b200rime_pack_f32(const float* X, long long ldx, int nfreq, int ns, int ns_pad,
                      long long soff, long long S, float* A, b200rime_stream_t stream);
int b200rime_pack_f64(const double* X, long long ldx, int nfreq, int ns, int ns_pad,
                      long long soff, long long S, double* A, b200rime_stream_t stream);
int b200rime_unpack_f32(const float* A, long long ldx, int nfreq, int ns, long long soff,
                        long long S, float* X, b200rime_stream_t stream);
int b200rime_unpack_f64(const double* A, long long ldx, int nfreq, int ns, long long soff,
                        long long S, double* X, b200rime_stream_t stream);

/* ---- fused perceived-sky builders (1-pol power beam) ------------------------------
 * Interpolated pixel beam (PixInterp.interp utils.py:833-841 + cut_sky_fov beam_model.py:1696
 * + beam*sky beam_model.py:341):
 *   A[f][soff+s] = ( sum_{i<nnn} bmap[f*ldb + inds[s][i]] * wgts[s][i] ) * sky[f*lds + cut[s]]
 * bmap == NULL gives the pure FOV gather A = sky[f][cut[s]], sky == NULL the pure interpolation
 * (factor 1 in place of the missing operand; used to build the Jones / coherency planes of the
 * polarised modes, beam_model.py:343-363, which torch then combines element-wise in the tiled
 * layout).  inds int32 [ns][nnn], wgts real [ns][nnn], cut int32 [ns]; cut[s] < 0 marks a padding entry
 * (A = 0), so one call with ns = ns_pad = S, soff = 0 builds every time of a group at once. */
/* The same product from a channel-major beam map bmapT[npix_beam][ldt], ldt >= nchunk * KC, the
 * channels beyond nfreq zero (bmapT must not be NULL; sky may be): every neighbour read is a full
 * 128-byte line instead of 32 scattered pixels per load.  Same outputs as build_interp. */
int b200rime_build_interp_t_f32(const float* bmapT, long long ldt, const int* inds,
                                const float* wgts, int nnn, const float* sky, long long lds,
                                const int* cut, int nfreq, int ns, int ns_pad, long long soff,
                                long long S, float* A, b200rime_stream_t stream);
int b200rime_build_interp_t_f64(const double* bmapT, long long ldt, const int* inds,
                                const double* wgts, int nnn, const double* sky, long long lds,
                                const int* cut, int nfreq, int ns, int ns_pad, long long soff,
                                long long S, double* A, b200rime_stream_t stream);
int b200rime_build_interp_f32(const float* bmap, long long ldb, const int* inds,
                              const float* wgts, int nnn, const float* sky, long long lds,
                              const int* cut, int nfreq, int ns, int ns_pad, long long soff,
                              long long S, float* A, b200rime_stream_t stream);
int b200rime_build_interp_f64(const double* bmap, long long ldb, const int* inds,
                              const double* wgts, int nnn, const double* sky, long long lds,
                              const int* cut, int nfreq, int ns, int ns_pad, long long soff,
                              long long S, double* A, b200rime_stream_t stream);
/* backward of the above given dA: writes (each output optional, NULL to skip)
 *   dsky[f*lds + cut[s]] += B[f][s] * dA[f][s]   read-modify-write; legal only when the call
 *                                                covers ONE time (cut indices unique)
 *   dIs[f*ldd + s]        = B[f][s] * dA[f][s]   row-major (Nf, ns) -- the multi-time form:
 *                                                follow with b200rime_gather_times_*
 *   dBI[f*ldd + s]        = sky[f][cut[s]] * dA[f][s]
 * b200rime_interp_transpose_* then gathers dBI into the beam map through the CSR transpose of
 * (inds, wgts):  dbmap[f*ldb + p] += sum_{j in [rowptr[p], rowptr[p+1])} val[j] *
 * dBI[f*ldd + col[j]]   -- one owner per (f, p), no atomics. */
/* float32, bilinear (4 neighbours) form of the backward from the channel-major beam map of
 * build_interp_t (16-byte neighbour and cotangent reads; dIs / dBI as above, each optional). */
int b200rime_build_interp_bwd_t_f32(const float* dA, const float* bmapT, long long ldt,
                                    const int* inds, const float* wgts, const float* sky,
                                    long long lds, const int* cut, int nfreq, int ns, long long soff,
                                    long long S, float* dBI, long long ldd, float* dIs,
                                    b200rime_stream_t stream);
int b200rime_build_interp_bwd_f32(const float* dA, const float* bmap, long long ldb,
                                  const int* inds, const float* wgts, int nnn, const float* sky,
                                  long long lds, const int* cut, int nfreq, int ns, long long soff,
                                  long long S, float* dsky, float* dBI, long long ldd, float* dIs,
                                  b200rime_stream_t stream);
int b200rime_build_interp_bwd_f64(const double* dA, const double* bmap, long long ldb,
                                  const int* inds, const double* wgts, int nnn, const double* sky,
                                  long long lds, const int* cut, int nfreq, int ns, long long soff,
                                  long long S, double* dsky, double* dBI, long long ldd, double* dIs,
                                  b200rime_stream_t stream);
/* dsky[f*lds + p] += sum_t dIs[f*ldd + pos[t*npix + p]]  over the times with pos >= 0
 * (pos: int32 [nt][npix], packed source index of sky pixel p at time t, or -1 when p is outside
 * the FOV).  One owner per (f, p), times added in index order: deterministic adjoint of
 * cut_sky_fov (beam_model.py:1696) for a whole time group. */
int b200rime_gather_times_f32(const float* dIs, long long ldd, const int* pos, int nt, int npix,
                              int nfreq, float* dsky, long long lds, b200rime_stream_t stream);
int b200rime_gather_times_f64(const double* dIs, long long ldd, const int* pos, int nt, int npix,
                              int nfreq, double* dsky, long long lds, b200rime_stream_t stream);
int b200rime_interp_transpose_f32(const float* dBI, long long ldd, const int* rowptr,
                                  const int* col, const float* val, int npix, int nfreq,
                                  float* dbmap, long long ldb, b200rime_stream_t stream);
int b200rime_interp_transpose_f64(const double* dBI, long long ldd, const int* rowptr,
                                  const int* col, const double* val, int npix, int nfreq,
                                  double* dbmap, long long ldb, b200rime_stream_t stream);

/* Airy power/voltage beam (beam_model.py:1464-1480, special.py:535):
 *   x = max(D(s) * sinzen[s] * pi * freqs[f] * freq_ratio / c, 1e-10),
 *   D(s) = Dns + sin2az[s] * (Dew - Dns),  B = (2 J1(x)/x)^(square ? 2 : 1)
 *   A[f][soff+s] = B * sky[f*lds + cut[s]]
 * sinzen = sin(min(zen, 90 deg)), sin2az = sin(az)^2: real [ns], precomputed per time.
 * diam_dev (optional): DEVICE float64 [2] = (Dew, Dns) that overrides the two host arguments, so
 * that a caller holding the diameters on the device needs no host synchronisation (and the launch
 * can be captured into a CUDA graph).
 * Bout (optional, may be NULL): row-major (Nf, ns) copy of B with row stride ldo. */
int b200rime_build_airy_f32(double Dew, double Dns, const double* diam_dev, double freq_ratio, int square,
                            const float* sinzen, const float* sin2az, const double* freqs,
                            const float* sky, long long lds, const int* cut, int nfreq, int ns,
                            int ns_pad, long long soff, long long S, float* A, float* Bout,
                            long long ldo, b200rime_stream_t stream);
int b200rime_build_airy_f64(double Dew, double Dns, const double* diam_dev, double freq_ratio, int square,
                            const double* sinzen, const double* sin2az, const double* freqs,
                            const double* sky, long long lds, const int* cut, int nfreq, int ns,
                            int ns_pad, long long soff, long long S, double* A, double* Bout,
                            long long ldo, b200rime_stream_t stream);
/* backward: dsky[f*lds + cut[s]] += B * dA (single-time calls) or dIs[f*ldd + s] = B * dA
 * (multi-time calls, then b200rime_gather_times_*); dD[block][0..1] partial sums of
 * dL/dDew, dL/dDns (float64 [nblocks][2], zero-initialised by the caller,
 * nblocks = b200rime_airy_bwd_blocks(nfreq, ns)), summed by the caller in index order.  full_grad != 0 uses the analytic derivative
 * d(2J1/x)/dx = 2 J0/x - 4 J1/x^2; full_grad == 0 reproduces the reference's autograd,
 * which treats J1(x) as a constant (torch.special.bessel_j1 has no derivative formula). */
int b200rime_airy_bwd_blocks(int nfreq, int ns);
int b200rime_build_airy_bwd_f32(const float* dA, double Dew, double Dns, const double* diam_dev, double freq_ratio,
                                int square, int full_grad, const float* sinzen,
                                const float* sin2az, const double* freqs, const float* sky,
                                long long lds, const int* cut, int nfreq, int ns, long long soff,
                                long long S, float* dsky, double* dD, float* dIs, long long ldd,
                                b200rime_stream_t stream);
int b200rime_build_airy_bwd_f64(const double* dA, double Dew, double Dns, const double* diam_dev, double freq_ratio,
                                int square, int full_grad, const double* sinzen,
                                const double* sin2az, const double* freqs, const double* sky,
                                long long lds, const int* cut, int nfreq, int ns, long long soff,
                                long long S, double* dsky, double* dD, double* dIs, long long ldd,
                                b200rime_stream_t stream);

/* ---- antenna-factorised fringe sum (float32; dense antenna-pair baseline sets) ------------
 * Same operator as b200rime_fringe_sum_fwd_f32 / _bwd_sky_f32 / _bwd_bl_f32 (reference
 * telescope_model.py:310-358 gen_fringe + rime_model.py:426-429 product and sum, and their
 * autograd backward), for baseline lists that cover most pairs of an antenna set.  The fringe
 * of baseline (i, j) with vector r_j - r_i (telescope_model.py:221-239) factorises as
 * conj(E_i) E_j, E_a = exp(+-2 pi i r_a . shat nu / c), so one evaluation is one complex
 * multiply-accumulate on the FP32 pipes; antenna terms are generated inside the kernels.
 *
 * antv       [na_pad][4] float64 antenna positions (ENU metres), rows >= na zero
 * tile_ant   [ntile][128] antenna row of X slots 0..63 (first antenna of a pair, conjugated)
 *            and Y slots 64..127 (second antenna); -1 = empty slot
 * tile_bl    [ntile][64][64] (baseline index << 1 | c) or -1; c = 1: the tile holds the pair
 *            swapped and the result is conjugated on output
 * tile_order [ntile] processing order of the tiles (tiles with second antennas only in slots
 *            0..31 last)
 * Vpart      [nunits][nbl][Nfp][2]   (same as fringe_sum_fwd; reduce with reduce_units)
 * ant_tile() = 64 antennas per tile side, ant_stage() = 8 reduction indices per stage. */
int b200rime_ant_tile(void);
int b200rime_ant_stage(void);
int b200rime_antfringe_fwd_f32(const float* A, const double* shat, const double* antv,
                               const double* freqs, const int* units, int nunits,
                               const int* tile_ant, const int* tile_bl, const int* tile_order,
                               int ntile, int nbl, int nfreq, long long S, int conj, float* Vpart,
                               b200rime_stream_t stream);
/* Backward of the same sum to the perceived sky and to the antenna positions in one pass.
 * Hp  [nt][Nfp][na_pad/64][nm_pad/8][8][64][2]: Hermitian cotangent matrix (nm_pad = number of
 *     antennas rounded up to 16: the partner axis m is not padded to a whole block)
 *     H[a][m] = G_b for b = (m, a), conj(G_b) for b = (a, m), 2 Re G_b for a = m (autos),
 *     indexed [time][channel][block of a][stage of m][m in stage][a in block], a stored at
 *     position ((a >> 1) & 3) * 16 + (a >> 3) * 2 + (a & 1) of its block.
 *     When drpart is NULL (no antenna gradient) only the part of H with m < 64 (block of a + 1)
 *     is read: pass the lower triangle H[a][m] = 2 G_b (b = (m, a)) / 2 conj(G_b) (b = (a, m)),
 *     a > m, and H[a][a] = 2 Re G_b for autos, and the kernel does half the work.
 *     A is only read for the antenna gradient and may be NULL together with drpart.
 * dApart [na_pad/64][nchunk][S][KC]    partial dL/dA per antenna block (sum over blocks), or NULL
 * drpart [nunits][Nfp][2][na_pad][4]   float64 partial dL/d(antenna position) (sum over the
 *                                       first three axes), or NULL; ZERO it before the call: a last
 *                                       block of at most 32 antennas (na - 64 (na_pad/64 - 1))
 *                                       is processed by one warp that writes its own rows only */
int b200rime_antfringe_bwd_f32(const float* Hp, const float* A, const double* shat,
                               const double* antv, const double* freqs, const int* units,
                               int nunits, int na, int na_pad, int nm_pad, int nfreq,
                               long long S, int conj, float* dApart, double* drpart,
                               b200rime_stream_t stream);

/* ---- tensor-core fringe sum (tcgen05 / TMEM) ---------------------------------------------
 * The same antenna-factorised sum as antfringe_fwd (reference telescope_model.py:310-358 +
 * rime_model.py:426-429) as a batched complex GEMM V = E^H diag(A) E on the 5th-generation
 * tensor cores: operands generated in shared memory (float64 phases, MUFU sine / cosine), split
 * into float16 hi + lo pairs (three MMAs per real product, float32-grade results), FP32
 * accumulators in tensor memory, read out every 64 sources into register accumulators (the
 * tensor core's accumulate truncates; short chains keep the bias below 1e-6).
 *   Acm      float [Nfp][S]   the perceived sky A, channel-major (row = channel over the packed
 *            source axis; the tiled A transposed), so that a stage's sources are contiguous
 *   ascale   float [1]   power of two s with max|A| s in [2^14, 2^15) (float16 range of A E)
 *   antv     f64 [>= na][4]   antenna positions (ENU metres)
 *   items    int32 [nitems][4]   {i0, j0, N, 0}: first antennas i0 .. i0 + 127 (tc_rows()) against
 *            second antennas j0 .. j0 + N - 1, N a multiple of 32, <= tc_cols_max() = 128;
 *            i0 + 128 and j0 + N may exceed na (rows beyond na are skipped)
 *   pair_bl  int32 [ldp][ldp]    (baseline << 1 | c) of the pair (first i, second j), or -1;
 *            c = 1: the baseline is (j, i) and the result is conjugated on output; ldp = na
 *            rounded up to 32.  Every wanted baseline must be covered by exactly one item.
 *   Vpart    [nunits][nbl][Nfp][2]   (as fringe_sum_fwd; reduce with reduce_units) */
int b200rime_tc_rows(void);
int b200rime_tc_cols_max(void);
int b200rime_tcfringe_fwd_f32(const float* Acm, const float* ascale, const double* shat,
                              const double* antv, const double* freqs, const int* units,
                              int nunits, const int* items, int nitems, const int* pair_bl,
                              int ldp, int na, int nbl, int nfreq, long long S, int conj,
                              float* Vpart, b200rime_stream_t stream);

/* Backward of the same sum on the tensor cores: y_a[s] = sum_m H[a][m] E_m[s] as a GEMM
 * (M = sources, N = antennas a, K = partner antennas m), then p = conj(E_a) y_a,
 * dL/dA = 1/2 sum_a Re p and dL/dr_a = sum_s shat_s A_s (2 pi sgn nu / c) Im p (autograd of
 * rime_model.py:429 w.r.t. psky and of telescope_model.py:356 w.r.t. the antenna positions).
 *   Hq      float16 [nt][Nfp][nitem][nm_pad/16][2][3][16][2][8][8]: the Hermitian cotangent matrix
 *           (as for antfringe_bwd: H[a][m] = G_b for b = (m, a), conj(G_b) for b = (a, m),
 *           2 Re G_b for autos; when only dL/dA is wanted the doubled lower triangle a > m is
 *           enough), times hscale, split into float16 hi + lo, as STACKED UMMA K-major B operands
 *           (one MMA of width 256 feeds the real and the imaginary accumulator):
 *           [item of 128 antennas a][stage of 16 m][hi | lo][-Im H ; Re H ; Im H][a / 8][m / 8]
 *           [a % 8][m % 8]; rows 0..255 of a buffer are the operand M = (-Im H ; Re H), rows
 *           128..383 the operand P = (Re H ; Im H)
 *   hscale  float [1]   power of two that brings max |H| into [2^14, 2^15)
 *   Acm     float [Nfp][S] channel-major perceived sky (only for drpart, else NULL)
 *   nitem = ceil(na / 128), nm_pad = na rounded up to 16, na <= 512
 *   mrange  int32 [nitem][2]   stages of 16 partner antennas [lo, hi) that hold entries of H for
 *           the item (all stages: {0, nm_pad / 16}; lower triangle: item ib ends at 8 (ib + 1))
 *   dAcm    float [nitem][Nfp][S]       partial dL/dA, channel-major; ZERO before the call
 *           (padding sources and channels are not written); sum over the first axis; or NULL
 *   drpart  float [nunits][Nfp][4][nitem * 128][4]   partial dL/dr; sum over the first three
 *           axes; or NULL */
int b200rime_tcfringe_bwd_f32(const void* Hq, const float* hscale, const float* Acm,
                              const double* shat, const double* antv, const double* freqs,
                              const int* units, int nunits, int nitem, int na, int nm_pad,
                              const int* mrange, int nfreq, long long S, int conj, float* dAcm,
                              float* drpart, b200rime_stream_t stream);

/* Cotangent operand of tcfringe_bwd straight from the autograd cotangent G (complex64, element
 * (baseline b, time t, channel k) at G[2 * (b * ldb + t * nf + k)]; ldb in complex elements):
 * builds H through the antenna-pair table of tcfringe_fwd (lower_only: the doubled lower
 * triangle), multiplies by hscale (device float [1], a power of two with max|H| hscale < 2^15),
 * splits into float16 hi + lo and writes Hq in the layout above.  One pass; replaces the
 * index_put / scale / split / stack / permute chain over a (Nt, Nfp, Na, Na) complex matrix. */
int b200rime_tc_pack_cotangent_f32(const float* G, long long ldb, const int* pair_bl, int ldp,
                                   int nt, int nf, int na, int nm_pad, int lower_only,
                                   const float* hscale, void* Hq, b200rime_stream_t stream);

/* ---- spherical-harmonic forward model on the tensor cores -----
 * AlmModel.forward_alm (sph_harm.py:1289-1373: einsum "...i,ij->...j" of the coefficients with the
 * Ylm matrix) behind YlmResponse.forward / set_beam_cache (beam_model.py:1166-1250), and its
 * adjoint to the coefficients (autograd of the same einsum in the reference):
 *     out[m][n] = sum_k X[m][k] Y[n][k]      complex, float32-grade (three float16 split MMAs per
 *                                            real product, FP32 accumulation in tensor memory)
 * Operands are first PACKED (split into float16 hi + lo, multiplied by a power of two `scale`
 * that brings max|.| into [2^14, 2^15), laid out in the UMMA canonical K-major order, one
 * contiguous block per (block of 128 rows, stage of 16 k)):
 *   cgemm_pack_a: X, element (row, k) at re[row * stride_row + k * stride_k] (+ im[...], NULL for a
 *                 real operand; strides in floats, so a complex64 tensor passes re = ptr,
 *                 im = ptr + 1 and doubled strides) -> Aq, cgemm_a_bytes(M, K) bytes
 *   cgemm_pack_b: Y likewise (N rows) -> Bq, cgemm_b_bytes(N, K) bytes
 *   negate_im = 1 packs the complex conjugate.
 * cgemm: a_real = 1 promises Im X = 0 (half the MMAs); real_out = 1 stores Re out only.
 *   out   float [M][ldo] (real_out) or complex64 [M][ldo]  (ldo in elements, >= N)
 *   ksplit > 1 splits the k axis over grid.y (1 <= ksplit <= ceil(K / 16)); `part`
 *   [ksplit][M][ldo] elements of workspace is then required and summed in index order. */
long long b200rime_cgemm_a_bytes(int M, int K);
long long b200rime_cgemm_b_bytes(int N, int K);
int b200rime_cgemm_pack_a_f32(const float* re, const float* im, long long stride_row,
                              long long stride_k, int M, int K, const float* scale, int negate_im,
                              void* Aq, b200rime_stream_t stream);
int b200rime_cgemm_pack_b_f32(const float* re, const float* im, long long stride_row,
                              long long stride_k, int N, int K, const float* scale, int negate_im,
                              void* Bq, b200rime_stream_t stream);
int b200rime_cgemm_f32(const void* Aq, const void* Bq, int M, int N, int K, int ksplit, int a_real,
                       int real_out, const float* scale_a, const float* scale_b, float* out,
                       long long ldo, float* part, b200rime_stream_t stream);
/* float64 sessions: the same product on the FP64 pipes straight from strided operands (no pack
 * pass, no tensor cores): X element (m, k) at xr[m * stride_xm + k * stride_xk] (xi likewise or
 * NULL), Y element (n, k) at yr[n * stride_yn + k * stride_yk]; conj_x / conj_y conjugate. */
int b200rime_cgemm_f64(const double* xr, const double* xi, long long stride_xm, long long stride_xk,
                       const double* yr, const double* yi, long long stride_yn, long long stride_yk,
                       int M, int N, int K, int conj_x, int conj_y, int real_out, double* out,
                       long long ldo, b200rime_stream_t stream);

/* ---- equatorial -> topocentric angles on the device ----------
 * The per-source part of TelescopeModel.eq2top (telescope_model.py:469-502, astropy ICRS -> AltAz
 * on the host in the reference): p = unit(ra, dec); p += v3 (annual aberration, first order),
 * renormalise; (E, N, U) = m9 p; zen = acos U, az = atan2(E, N) in [0, 360) degrees.  m9 (row-major
 * 3 x 3, ICRS -> East / North / Up at the observation time) and v3 (observer velocity / c, ICRS;
 * NULL = no aberration) are HOST arrays (telescope_model.icrs_to_enu builds them: IAU 1976
 * precession, IAU 1980 nutation leading terms, apparent sidereal time, latitude).  ra, dec, zen,
 * az: device float64 [n].  No refraction. */
int b200rime_eq2top_f64(const double* ra_deg, const double* dec_deg, long long n, const double* m9,
                        const double* v3, double* zen_deg, double* az_deg,
                        b200rime_stream_t stream);

/* ---- Jones sandwich of the polarised beam modes -------------------------------------------
 * P[a][d] = sum_{b,c} J1[a][b] C[b][c] J2[d][c] for real 2 x 2 Jones planes and coherency planes
 * (beam_model.py:347, :363 einsum "ab...,bc...,dc...->ad..."), element-wise over n elements per
 * plane (n % 4 == 0; the tiled layout).  J1 / J2 / C / P: HOST arrays of 4 DEVICE plane pointers,
 * index 2 * row + col.  Backward: dJ1 = dP (J2 C^T), dJ2 = dP^T (J1 C), dC = J1^T dP J2; `same`
 * != 0 (both antennas use one beam model, J2 == J1): the sum dJ1 + dJ2 is written to dJ1 and dJ2
 * is not touched.  Output tables (or single entries) may be NULL. */
int b200rime_jones_sandwich_f32(const float* const* J1, const float* const* J2,
                                const float* const* C, long long n, float* const* P,
                                b200rime_stream_t stream);
int b200rime_jones_sandwich_f64(const double* const* J1, const double* const* J2,
                                const double* const* C, long long n, double* const* P,
                                b200rime_stream_t stream);
int b200rime_jones_sandwich_bwd_f32(const float* const* dP, const float* const* J1,
                                    const float* const* J2, const float* const* C, long long n,
                                    int same, float* const* dJ1, float* const* dJ2,
                                    float* const* dC, b200rime_stream_t stream);
int b200rime_jones_sandwich_bwd_f64(const double* const* dP, const double* const* J1,
                                    const double* const* J2, const double* const* C, long long n,
                                    int same, double* const* dJ1, double* const* dJ2,
                                    double* const* dC, b200rime_stream_t stream);

/* ---- gain application --------------------------------------
 * V_out = g_1 V g_2^H per baseline: reference calibration._apply_cal (calibration.py:2412-2487),
 * the step that follows the RIME in a BayesLIM Sequential.
 *   vis / out [npol][npol][nbl][nt][nf] complex; gains [npol][npol][nant][ntg][nfg] complex with
 *   ntg in {1, nt}, nfg in {1, nf}; g1 / g2 [nbl] gain-table row of the first / second antenna.
 *   full = 0: 1pol, or 2pol (4pol data, diagonal gains; off-diagonal outputs are zeroed as
 *   linalg.diag_matmul does, linalg.py:116-149); full = 1 (npol = 2): 2x2 Jones products.
 *   cov / cov_out: optional real variance of the data's shape, scaled by |g_1 conj(g_2)|^2
 *   (diagonal modes only), else NULL. */
int b200rime_apply_cal_f32(const float* vis, const float* gains, const int* g1, const int* g2,
                           int npol, int full, int nbl, int nt, int nf, int nant, int ntg, int nfg,
                           const float* cov, float* out, float* cov_out,
                           b200rime_stream_t stream);
int b200rime_apply_cal_f64(const double* vis, const double* gains, const int* g1, const int* g2,
                           int npol, int full, int nbl, int nt, int nf, int nant, int ntg, int nfg,
                           const double* cov, double* out, double* cov_out,
                           b200rime_stream_t stream);
/* Adjoint of the same product to the gains for a cotangent gout of the output's shape (the
 * adjoint to vis is apply_cal itself with conjugate-transposed gains).  p1/b1 (p2/b2): CSR lists
 * of the baselines in which an antenna is the first (second) one: p* [nant + 1], b* [nbl].
 * dg [npol][npol][nant][nt][nf] complex: full time / frequency axes, fixed summation order; the
 * caller sums the axes along which the gains broadcast. */
int b200rime_apply_cal_bwd_gains_f32(const float* vis, const float* gains, const float* gout,
                                     const int* g1, const int* g2, const int* p1, const int* b1,
                                     const int* p2, const int* b2, int npol, int full, int nbl,
                                     int nt, int nf, int nant, int ntg, int nfg, float* dg,
                                     b200rime_stream_t stream);
int b200rime_apply_cal_bwd_gains_f64(const double* vis, const double* gains, const double* gout,
                                     const int* g1, const int* g2, const int* p1, const int* b1,
                                     const int* p2, const int* b2, int npol, int full, int nbl,
                                     int nt, int nf, int nant, int ntg, int nfg, double* dg,
                                     b200rime_stream_t stream);

/* ---- on-device peak measurements used as roofline denominators ---------------------
 * kind: 0 = FP32 FFMA chains, 1 = FP64 DFMA chains, 2 = MUFU sin+cos, 3 = packed FP32x2
 * FFMA2 chains.  Runs `iters`
 * dependent-chain iterations on every SM and returns achieved Gop/s (FMA counted as 2 flop;
 * MUFU as 1 op) in *gops and the kernel time in *ms.  Synchronises the device. */
int b200rime_microbench(int kind, int iters, double* gops, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* B200RIME_H */
