// Event prediction from PDWs: the host-side analysis that follows the PDW extractor in
// matlab/predict_event.m:125-138 and cpp/usrp_predict_event.cpp:28-52,348-373.  A few dozen doubles per
// recording: plain host code, no GPU work.
#include <algorithm>
#include <cmath>
#include <vector>

#include "channelizer.h"

extern "C" {

// p = polyfit(pdw.toa, pdw.snr, 2) (predict_event.m:125) / T.householderQr().solve(V) with T = [1 t t^2]
// (usrp_predict_event.cpp:31-49): Householder QR of the n x 3 Vandermonde matrix, back substitution.
int chz_event_peak_time(const double* t, const double* v, uint64_t n, double* t_peak, double* v_peak, double* coef) {
  if (!t || !v || n < 3) return CHZ_EINVAL;
  const size_t N = (size_t)n;
  std::vector<double> A(N * 3), b(v, v + N);
  for (size_t i = 0; i < N; i++) { A[i * 3] = 1.0; A[i * 3 + 1] = t[i]; A[i * 3 + 2] = t[i] * t[i]; }
  double R[3][3] = {{0}};
  for (int k = 0; k < 3; k++) {
    double norm = 0.0;
    for (size_t i = k; i < N; i++) norm += A[i * 3 + k] * A[i * 3 + k];
    norm = std::sqrt(norm);
    if (norm == 0.0) return CHZ_EINVAL;                       // rank deficient (all times equal, ...)
    const double alpha = A[(size_t)k * 3 + k] > 0 ? -norm : norm;
    // Householder vector w = x - alpha e1 (stored in place), H = I - 2 w w^T / (w^T w)
    A[(size_t)k * 3 + k] -= alpha;
    double ww = 0.0;
    for (size_t i = k; i < N; i++) ww += A[i * 3 + k] * A[i * 3 + k];
    if (ww == 0.0) return CHZ_EINVAL;
    for (int j = k + 1; j < 3; j++) {
      double dot = 0.0;
      for (size_t i = k; i < N; i++) dot += A[i * 3 + k] * A[i * 3 + j];
      const double f = 2.0 * dot / ww;
      for (size_t i = k; i < N; i++) A[i * 3 + j] -= f * A[i * 3 + k];
    }
    double dot = 0.0;
    for (size_t i = k; i < N; i++) dot += A[i * 3 + k] * b[i];
    const double f = 2.0 * dot / ww;
    for (size_t i = k; i < N; i++) b[i] -= f * A[i * 3 + k];
    R[k][k] = alpha;
    for (int j = k + 1; j < 3; j++) R[k][j] = A[(size_t)k * 3 + j];
  }
  double p[3];
  for (int k = 2; k >= 0; k--) {
    double s = b[k];
    for (int j = k + 1; j < 3; j++) s -= R[k][j] * p[j];
    if (R[k][k] == 0.0) return CHZ_EINVAL;
    p[k] = s / R[k][k];
  }
  if (coef) { coef[0] = p[0]; coef[1] = p[1]; coef[2] = p[2]; }
  if (p[2] == 0.0) return CHZ_EINVAL;                         // a straight line has no peak
  const double tm = -p[1] / (2.0 * p[2]);                     // predict_event.m:128, usrp_predict_event.cpp:51
  if (t_peak) *t_peak = tm;
  if (v_peak) *v_peak = p[2] * tm * tm + p[1] * tm + p[0];    // :129
  return CHZ_OK;
}

// nextEvent = median(diff(event)) + t_max (predict_event.m:133-135); fewer than two events: t_max + a fixed
// interval (:137).  upper_median: the element [size/2] of the sorted differences, as the C++ tool takes it
// (usrp_predict_event.cpp:364-368) instead of MATLAB's mean of the two middle values.
int chz_next_event_time(const double* events, uint64_t n, double fallback_interval, int upper_median, double* next) {
  if (!events || !next || n == 0) return CHZ_EINVAL;
  const double last = events[n - 1];
  if (n < 2) { *next = last + fallback_interval; return CHZ_OK; }
  std::vector<double> d((size_t)n - 1);
  for (size_t i = 1; i < (size_t)n; i++) d[i - 1] = events[i] - events[i - 1];
  std::sort(d.begin(), d.end());
  const size_t m = d.size();
  const double med = (upper_median || (m & 1)) ? d[m / 2] : 0.5 * (d[m / 2 - 1] + d[m / 2]);
  *next = last + med;
  return CHZ_OK;
}

}  // extern "C"
