// mpcv_c2d.cuh — exact zero-order hold of xdot = Ac x + Bc u:  [A B; 0 I] = expm([Ac Bc; 0 0] dt), what
// mpc.util.c2d computes on the host for every scenario / every step of the LTV scripts
// (Inverted_pendulum/inverted_pendulum_single_shooting_mpctools.py:24, Trjectory_tracking_le_LTV.py:126-133,
// Trajectory_tracking_dynamic_model.py:134).  One thread per system; scaling and squaring around a degree-18 Taylor
// polynomial evaluated by Horner's rule (||M / 2^s||_1 <= 1/4, so truncation is far below one ulp; the stiff
// dynamic bicycle needs s = 9 squarings).
#pragma once

namespace mpcv {

// E = expm(M) for an S x S matrix in registers (M is overwritten by its scaled copy)
template <int S>
__device__ __forceinline__ void expm_small(double* M, double* E) {
  double T[S * S];
  double nrm = 0.0;
#pragma unroll
  for (int j = 0; j < S; ++j) {
    double c = 0.0;
#pragma unroll
    for (int i = 0; i < S; ++i) c += fabs(M[i * S + j]);
    nrm = fmax(nrm, c);
  }
  int s = 0;
  while (nrm > 0.25 && s < 60) { nrm *= 0.5; ++s; }
  const double sc = ldexp(1.0, -s);
#pragma unroll
  for (int i = 0; i < S * S; ++i) M[i] *= sc;
  // Horner: E = I + M (I + M/2 (I + M/3 (... (I + M/18))))
#pragma unroll
  for (int i = 0; i < S * S; ++i) E[i] = 0.0;
#pragma unroll
  for (int i = 0; i < S; ++i) E[i * S + i] = 1.0;
  for (int k = 18; k >= 1; --k) {
    const double rk = 1.0 / k;
#pragma unroll
    for (int i = 0; i < S; ++i) {
#pragma unroll
      for (int j = 0; j < S; ++j) {
        double v = 0.0;
#pragma unroll
        for (int l = 0; l < S; ++l) v += M[i * S + l] * E[l * S + j];
        T[i * S + j] = v * rk;
      }
    }
#pragma unroll
    for (int i = 0; i < S * S; ++i) E[i] = T[i];
#pragma unroll
    for (int i = 0; i < S; ++i) E[i * S + i] += 1.0;
  }
  for (int q = 0; q < s; ++q) {
#pragma unroll
    for (int i = 0; i < S; ++i) {
#pragma unroll
      for (int j = 0; j < S; ++j) {
        double v = 0.0;
#pragma unroll
        for (int l = 0; l < S; ++l) v += E[i * S + l] * E[l * S + j];
        T[i * S + j] = v;
      }
    }
#pragma unroll
    for (int i = 0; i < S * S; ++i) E[i] = T[i];
  }
}

}  // namespace mpcv
