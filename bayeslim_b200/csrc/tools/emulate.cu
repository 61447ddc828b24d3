// CPU emulation of the per-thread arithmetic of the fringe-sum kernels (rime_math.cuh).
// Build:  nvcc -O2 -std=c++17 -I.. tools/emulate.cu -o /tmp/emulate   (runs on the host, no GPU)
// Reports the worst single-source relative error of the float32 recurrence against a direct
// float64 evaluation -- the error budget behind the 1e-5 parity tolerance.
#include <cstdio>
#include <cstdlib>
#include <complex>
#include <random>
#include <vector>
#include "../rime_math.cuh"

using namespace b200rime;

template <typename T>
double run(int nfreq, double blmax, int nsrc, int nbl, unsigned seed, double* sum_err) {
    constexpr int KC = Cfg<T>::KC;
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    std::vector<double> freqs(nfreq);
    for (int f = 0; f < nfreq; ++f) freqs[f] = 100e6 + 100e6 * f / (nfreq - 1);
    const int nchunk = (nfreq + KC - 1) / KC;
    double worst_single = 0.0, worst_sum = 0.0;
    for (int b = 0; b < nbl; ++b) {
        double bx = blmax * U(rng), by = blmax * U(rng), bz = 0.01 * blmax * U(rng);
        std::vector<std::complex<double>> ref(nfreq, 0.0), got(nfreq, 0.0);
        double amax = 0;
        for (int s = 0; s < nsrc; ++s) {
            double z = U(rng) * 0.5 + 0.5, ph = 3.14159265358979 * U(rng);
            double r = std::sqrt(1 - z * z);
            double sx = r * std::sin(ph), sy = r * std::cos(ph), sz = z;
            double u = std::fma(bx, sx, std::fma(by, sy, bz * sz));
            double a_amp = std::abs(U(rng)) + 0.1;
            for (int c = 0; c < nchunk; ++c) {
                std::vector<double> fpad(freqs);
                ChunkFreq cf = chunk_freq(freqs.data(), nfreq, c, KC, 1.0 / C_LIGHT);
                alignas(16) T a[KC];
                FwdTile<T, KC> acc;
                acc.zero();
                for (int k = 0; k < KC; ++k) a[k] = (c * KC + k < nfreq) ? (T)a_amp : (T)0;
                T zr, zi, wr, wi;
                chunk_seed(u, cf.k_mid, cf.k_step, zr, zi, wr, wi);
                acc.accumulate(a, zr, zi, wr, wi);
                for (int k = 0; k < KC && c * KC + k < nfreq; ++k) {
                    int f = c * KC + k;
                    double p = 2 * M_PI * std::fmod(u * freqs[f] / C_LIGHT, 1.0);
                    std::complex<double> e = (double)(T)a_amp * std::complex<double>(std::cos(p), std::sin(p));
                    T gr_, gi_;
                    acc.get(k, gr_, gi_);
                    std::complex<double> g((double)gr_, (double)gi_);
                    worst_single = std::max(worst_single, std::abs(g - e) / (double)a_amp);
                    ref[f] += e;
                    got[f] += g;
                }
            }
            amax += a_amp;
        }
        double vmax = 0, emax = 0;
        for (int f = 0; f < nfreq; ++f) {
            vmax = std::max(vmax, std::abs(ref[f]));
            emax = std::max(emax, std::abs(ref[f] - got[f]));
        }
        worst_sum = std::max(worst_sum, emax / vmax);
    }
    *sum_err = worst_sum;
    return worst_single;
}

int main() {
    double se;
    double e32 = run<float>(1024, 1000.0, 64, 64, 1, &se);
    std::printf("f32 recurrence: worst single-source rel err %.3e ; worst summed rel err (64 src) %.3e\n", e32, se);
    double e32b = run<float>(1000, 300.0, 256, 32, 2, &se);
    std::printf("f32 (Nf=1000, ragged last chunk): single %.3e ; summed %.3e\n", e32b, se);
    double e64 = run<double>(1024, 1000.0, 64, 64, 3, &se);
    std::printf("f64 recurrence: worst single-source rel err %.3e ; summed %.3e\n", e64, se);
    int bad = (e32 > 5e-6) || (e64 > 3e-12);
    std::printf(bad ? "FAIL\n" : "OK\n");
    return bad;
}
