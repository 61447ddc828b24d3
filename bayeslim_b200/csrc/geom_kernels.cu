// On-device equatorial -> topocentric conversion: the geometry
// stage the reference runs on the host through astropy, ICRS -> AltAz per time and per source
// (telescope_model.py:469-502), 7.9e5 sources per time at nside 256.
//
// The rotation ICRS -> local (East, North, Up) at the observation time (frame bias omitted,
// IAU 1976 precession, IAU 1980 nutation leading terms, apparent sidereal time, latitude) is a
// 3 x 3 matrix built once per time on the host in float64 (telescope_model.icrs_to_enu); annual
// aberration is a per-source shift by the observer's velocity v / c (first order, then
// renormalised).  This kernel is the per-source part: unit vector, aberration, rotation,
// (zenith angle, azimuth East of North) in degrees.  HBM-bound: 16 bytes in, 16 bytes out per
// source.  No refraction (astropy's AltAz default, pressure 0).
#include "common.cuh"
#include "internal.h"

namespace b200rime {

struct Eq2TopArgs {
    double m[9];      // row-major ICRS -> ENU
    double v[3];      // observer velocity / c in ICRS
};

__global__ void __launch_bounds__(256)
eq2top_kernel(const double* __restrict__ ra, const double* __restrict__ dec, long long n,
              Eq2TopArgs a, double* __restrict__ zen, double* __restrict__ az) {
    const double d2r = 0.017453292519943295769, r2d = 57.295779513082320877;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double sr, cr, sd, cd;
        sincos(ra[i] * d2r, &sr, &cr);
        sincos(dec[i] * d2r, &sd, &cd);
        double px = cd * cr + a.v[0], py = cd * sr + a.v[1], pz = sd + a.v[2];
        const double inv = rsqrt(px * px + py * py + pz * pz);
        px *= inv, py *= inv, pz *= inv;
        const double e = a.m[0] * px + a.m[1] * py + a.m[2] * pz;
        const double nn = a.m[3] * px + a.m[4] * py + a.m[5] * pz;
        const double u = a.m[6] * px + a.m[7] * py + a.m[8] * pz;
        zen[i] = atan2(sqrt(e * e + nn * nn), u) * r2d;      // well conditioned at the zenith
        double azd = atan2(e, nn) * r2d;
        if (azd < 0.0) azd += 360.0;
        az[i] = azd;
    }
}

}  // namespace b200rime

extern "C" int b200rime_eq2top_f64(const double* ra_deg, const double* dec_deg, long long n,
                                   const double* m9, const double* v3, double* zen_deg,
                                   double* az_deg, void* stream) {
    using namespace b200rime;
    if (n <= 0) return 0;
    if (m9 == nullptr) return set_error("eq2top: the rotation matrix is required");
    Eq2TopArgs a;
    for (int i = 0; i < 9; ++i) a.m[i] = m9[i];
    for (int i = 0; i < 3; ++i) a.v[i] = v3 != nullptr ? v3[i] : 0.0;
    const long long want = (n + 255) / 256;
    const int grid = (int)(want < 148LL * 8 ? want : 148LL * 8);
    eq2top_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ra_deg, dec_deg, n, a, zen_deg, az_deg);
    return check_launch("eq2top");
}
