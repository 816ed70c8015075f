// R4: default prototype low-pass of the channelizer.
// dsp.Channelizer(M) (matlab/create_pdws_channelized.m:33, one-argument constructor) uses the toolbox
// defaults NumTapsPerBand = 12 and StopbandAttenuation = 80 dB.  The toolbox designer is closed
// source, so this is the standard Kaiser-window design those defaults describe: a windowed sinc with
// cutoff fs/(2M) (a Nyquist-M filter), beta from Kaiser's formula, normalised to unity DC gain.
#include <cmath>
#include <vector>

#include "channelizer.h"

namespace {
double i0(double x) {   // modified Bessel function of the first kind, order 0 (power series)
  const double q = 0.25 * x * x;
  double term = 1.0, sum = 1.0;
  for (int k = 1; k < 1000; k++) {
    term *= q / ((double)k * (double)k);
    sum += term;
    if (term < sum * 1e-17) break;
  }
  return sum;
}
double kaiser_beta(double atten_db) {
  if (atten_db > 50.0) return 0.1102 * (atten_db - 8.7);
  if (atten_db >= 21.0) return 0.5842 * std::pow(atten_db - 21.0, 0.4) + 0.07886 * (atten_db - 21.0);
  return 0.0;
}
}  // namespace

extern "C" int chz_design_prototype(uint32_t M, uint32_t taps_per_band, double stopband_atten_db, float* taps) {
  if (!taps || M == 0 || taps_per_band == 0) return CHZ_EINVAL;
  const size_t L = (size_t)M * taps_per_band;
  const double pi = 3.14159265358979323846264338327950288;
  const double beta = kaiser_beta(stopband_atten_db), denom = i0(beta), centre = 0.5 * (double)(L - 1);
  std::vector<double> h(L);
  double dc = 0.0;
  for (size_t n = 0; n < L; n++) {
    const double t = (double)n - centre;
    const double arg = pi * t / (double)M;
    const double ideal = t == 0.0 ? 1.0 : std::sin(arg) / arg;
    const double r = L > 1 ? t / centre : 0.0;
    h[n] = ideal * i0(beta * std::sqrt(std::fmax(0.0, 1.0 - r * r))) / denom;
    dc += h[n];
  }
  for (size_t n = 0; n < L; n++) taps[n] = (float)(h[n] / dc);
  return CHZ_OK;
}
