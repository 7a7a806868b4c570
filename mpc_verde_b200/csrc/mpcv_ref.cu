// mpcv_ref.cu — the reference-trajectory pipeline on the device: everything the scripts
// compute in Python loops right before the solve, as batched kernels whose outputs are read in place by
// mpcv_closed_loop_ex.
//
//   reference (host, one scenario, nested Python loops)                          here (device, B scenarios)
//   ---------------------------------------------------------------------------  ---------------------------
//   par[:, k, t] of the lateral-error trackers                                     mpcv_ref_lateral
//       Trajectory Tracking/Trajectory_tracking_lateral_error.py:94-116, Phiref.py:124-155
//   p[k, :] of the Frenet bicycle                                                  mpcv_ref_frenet
//       Trajectory Tracking/test2.py:79-100
//   circle reference of the unicycle tracker                                       mpcv_ref_circle
//       Trajectory Tracking/Trajectory_tracking.py:84-97
//   (x, y, theta, v, omega) references cut from a lane-change path  mpcv_ref_unicycle_path
//   lane-change path extension (arcs and straights)                                mpcv_path_lane_change_ext
//       Trajectory Tracking/lane_change.py:5-79
//   per-step c2d of the LTV models                                                 mpcv_ltv_lateral, mpcv_ltv_dynbike
//       Trjectory_tracking_le_LTV.py:126-133, Trajectory_tracking_dynamic_model.py:119-134
//
// The scripts fill their tables sequentially and read entries of the previous step's slice back (`par[1, k, t-1]`).
// Every such entry is itself a closed-form function of the path, so here each output (scenario, step t, stage k) is
// computed independently by one thread from the path resident in HBM.  The scripts' quirks are kept and named
// where they enter (they are pinned by dados2.csv / out.csv in tests/): the still-zero slice read at t = 0, the
// wrap of stage k-1 to the last stage, the swapped p[2] / p[3] of test2.py.
#include <cmath>
#include <vector>

#include "mpcv_host.h"
#include "mpcv_c2d.cuh"

namespace {

// a scenario is the base path stretched by (sx, sy): SURVEY 8d scales the CSV path laterally and in speed
struct PathView {
  const double* x; const double* y; int n; double sx, sy;
  __device__ double px(int i) const { return x[i] * sx; }
  __device__ double py(int i) const { return y[i] * sy; }
  // heading of the chord arriving at sample i
  __device__ double chord(int i) const { return atan2(py(i) - py(i - 1), px(i) - px(i - 1)); }
};
__device__ PathView path_of(const double* x, const double* y, int n, const double* scale, long b) {
  return PathView{x, y, n, scale ? scale[2 * b] : 1.0, scale ? scale[2 * b + 1] : 1.0};
}

// phi_ref of stage k at MPC step t, as the lateral-error scripts fill it (Phiref.py:127-136).  t < 0 is the slice the
// script reads at t = 0 through `par[1, k, t-1]`: the LAST slice of the table, still zero at that time.
__device__ double lat_phi(const PathView& p, int k, int t) {
  if (t < 0) return 0.0;
  const int i = t + k, last = p.n - 1;
  if (i > last) return p.chord(last);
  if (i == 0) return 0.0;
  return p.chord(i);
}

__global__ void ref_lateral_kernel(const double* x, const double* y, int nsim, const double* scale, int Nt, double Delta,
                                   double ar, double br, double* pwin, long B) {
  const long items = B * nsim * (long)Nt;
  for (long it = (long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long)gridDim.x * blockDim.x) {
    const int k = (int)(it % Nt), t = (int)((it / Nt) % nsim);
    const long b = it / ((long)Nt * nsim);
    const PathView p = path_of(x, y, nsim, scale, b);
    const int i = t + k, last = nsim - 1;
    const double yref = p.py(i > last ? last : i);
    const double phi = lat_phi(p, k, t);
    double r, delta;
    if (i < 2) {
      // start of the path: forward differences of the chord heading
      const double plus = p.chord(i + 1), plus2 = p.chord(i + 2);
      r = (plus - phi) / Delta;
      delta = (((plus2 - 2 * plus + phi) / (Delta * Delta)) - ar * r) / br;
    } else if (i > nsim - 3) {
      // end of the path: backward differences against the previous step's slice; stage k-1 of k = 0 wraps to the
      // last stage of that slice (Phiref.py:147-150)
      const double prev = lat_phi(p, k, t - 1), prev_km1 = lat_phi(p, k > 0 ? k - 1 : Nt - 1, t - 1);
      r = (phi - prev) / Delta;
      delta = (((phi - 2 * prev + prev_km1) / (Delta * Delta)) - ar * r) / br;
    } else {
      const double plus = p.chord(i + 1), prev = lat_phi(p, k, t - 1);
      r = (plus - prev) / (2 * Delta);
      delta = (((plus - 2 * phi + prev) / (Delta * Delta)) - ar * r) / br;
    }
    double* o = pwin + it * 4;
    o[0] = yref; o[1] = phi; o[2] = r; o[3] = delta;
  }
}

// test2.py:79-100.  p[k] = (y_ref, phi_ref, p2, p3): p2 = vdes, p3 = |(xdd, ydd)| by central differences — which the
// cost and the model then unpack as (kappat, vdes) = (p[2], p[3]): swapped in the script, and kept that way.  Beyond
// the end of the path p3 repeats its last central-difference value (the script copies p[k-1, 3] stage by stage and,
// at k = 0, from the last stage of the previous step: the same number).
__global__ void ref_frenet_kernel(const double* x, const double* y, const double* vdes, int nsim, const double* scale,
                                  int Nt, double Delta, int n_steps, double* pwin, long B) {
  const long items = B * n_steps * (long)Nt;
  for (long it = (long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long)gridDim.x * blockDim.x) {
    const int k = (int)(it % Nt), t = (int)((it / Nt) % n_steps);
    const long b = it / ((long)Nt * n_steps);
    const PathView p = path_of(x, y, nsim, scale, b);
    const int i = t + k, last = nsim - 1;
    double* o = pwin + it * 4;
    o[0] = p.py(i > last ? last : i);
    o[1] = i > last ? p.chord(last) : (i == 0 ? 0.0 : p.chord(i));
    if (i < 2) {
      o[3] = 1.0;
      o[2] = vdes[i];
    } else {
      const int c = i > nsim - 2 ? nsim - 2 : i;      // the last sample with both neighbours
      const double ddx = (p.px(c - 1) - 2 * p.px(c) + p.px(c + 1)) / (Delta * Delta);
      const double ddy = (p.py(c - 1) - 2 * p.py(c) + p.py(c + 1)) / (Delta * Delta);
      o[3] = sqrt(ddx * ddx + ddy * ddy);
      o[2] = vdes[i > nsim - 2 ? last : i];
    }
  }
}

// (x, y, theta, v, omega) references of the unicycle tracker from a path sampled every dt: heading and speed of the
// path by central differences (one-sided at the ends), turn rate by differences of the heading; v and omega clipped to
// the control box
__device__ double diff_c(double m, double c, double p, int i, int n) { return i == 0 ? p - c : (i == n - 1 ? c - m : 0.5 * (p - m)); }
__global__ void ref_unicycle_kernel(const double* x, const double* y, int T, const double* scale, double dt, double vmax,
                                    double wmax, double* ptraj, long B) {
  const long items = B * T;
  for (long it = (long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long)gridDim.x * blockDim.x) {
    const int i = (int)(it % T);
    const long b = it / T;
    const PathView p = path_of(x, y, T, scale, b);
    auto heading = [&](int j) {
      const int jm = j > 0 ? j - 1 : j, jp = j < T - 1 ? j + 1 : j;
      return atan2(diff_c(p.py(jm), p.py(j), p.py(jp), j, T), diff_c(p.px(jm), p.px(j), p.px(jp), j, T));
    };
    const int im = i > 0 ? i - 1 : i, ip = i < T - 1 ? i + 1 : i;
    const double dx = diff_c(p.px(im), p.px(i), p.px(ip), i, T), dy = diff_c(p.py(im), p.py(i), p.py(ip), i, T);
    const double th = atan2(dy, dx);
    const double v = sqrt(dx * dx + dy * dy) / dt;
    const double w = diff_c(heading(im), th, heading(ip), i, T) / dt;
    double* o = ptraj + it * 5;
    o[0] = p.px(i); o[1] = p.py(i); o[2] = th;
    o[3] = fmin(fmax(v, -vmax), vmax);
    o[4] = fmin(fmax(w, -wmax), wmax);
  }
}

// Trajectory_tracking.py:84-97: unit circle, x = cos 0.1 t, y = sin 0.1 t, theta = pi/2 + 0.1 t, v_ref = omega_ref = 1.
// The script's window at step t, stage k is the sample at time (t + k) Delta: one trajectory, sliding window.
__global__ void ref_circle_kernel(int T, double Delta, double* ptraj) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  const double tt = i * Delta;
  double s, c;
  sincos(0.1 * tt, &s, &c);
  double* o = ptraj + (long)i * 5;
  o[0] = c; o[1] = s; o[2] = M_PI / 2 + 0.1 * tt; o[3] = 1.0; o[4] = 1.0;
}

// ---- lane_change.py:5-79 --------------------------------------------------------------------------------------
// The extension of the 500-sample lane change is six geometric pieces: arcs (centre, radius, angle range) and
// straights (start, length) sampled uniformly; every piece starts where the previous one ends and drops its first
// sample.  The plan (pieces, sample counts, end points) is a handful of scalars computed on the host; the samples are
// evaluated by one thread each.
struct Piece {
  int kind;          // 0 arc, 1 straight along -x
  int n;             // samples of the piece (its first one is dropped in the output)
  int out0;          // index of its second sample in the output
  double cx, cy, r;  // arc: centre, radius; straight: start point (cx, cy), speed r
  double t0, t1;     // parameter range (angle, or time)
};
struct Plan { Piece p[6]; int total; };

__global__ void path_ext_kernel(const Plan plan, const double* a, const double* b, const double* c, int n0, double v,
                                double* xt, double* yt, double* c2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= plan.total) return;
  if (i < n0) { xt[i] = a[i]; yt[i] = b[i]; c2[i] = c[i]; return; }
  int s = 0;
  while (s < 5 && i >= plan.p[s + 1].out0) ++s;
  const Piece& q = plan.p[s];
  const int j = i - q.out0 + 1;                               // sample of the piece (0 is dropped)
  // numpy.linspace: start + j * step, the last sample exactly `stop`
  const double step = (q.t1 - q.t0) / (q.n - 1);
  const double t = j == q.n - 1 ? q.t1 : q.t0 + j * step;
  if (q.kind == 0) {
    double sn, cs;
    sincos(t, &sn, &cs);
    xt[i] = q.cx + q.r * cs;
    yt[i] = q.cy + q.r * sn;
  } else {
    xt[i] = q.cx - q.r * t;
    yt[i] = q.cy;
  }
  c2[i] = v;
}

// Ac(u_ref) of the lateral-error bicycle, discretised per step (Trjectory_tracking_le_LTV.py:126-133):
// pglob = [A row-major (9), B (3)]
__global__ void ltv_lateral_kernel(const double* c, int T, const double* spd, double ar, double br, double dt, int n_steps,
                                   double* out, long B) {
  const long items = B * n_steps;
  for (long it = (long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long)gridDim.x * blockDim.x) {
    const int t = (int)(it % n_steps);
    const long b = it / n_steps;
    const double u = c[t < T ? t : T - 1] * (spd ? spd[b] : 1.0);
    double M[16] = {0, u, 0, 0, 0, 0, 1, 0, 0, 0, ar, br, 0, 0, 0, 0}, E[16];
    for (int i = 0; i < 16; ++i) M[i] *= dt;
    mpcv::expm_small<4>(M, E);
    double* o = out + it * 12;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) o[i * 3 + j] = E[i * 4 + j];
    for (int i = 0; i < 3; ++i) o[9 + i] = E[i * 4 + 3];
  }
}

// dynamic bicycle, LTV in v = vref[t] (Trajectory_tracking_dynamic_model.py:37-43,119-134), formulas exactly as
// written — including the operator precedence of A34 at :120 — then c2d: pglob = [A row-major (16), B (4)]
__global__ void ltv_dynbike_kernel(const double* v, int T, int per_scenario, double m, double a, double bb, double Ca,
                                   double Jz, double dt, int n_steps, double* out, long B) {
  const long items = B * n_steps;
  for (long it = (long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long)gridDim.x * blockDim.x) {
    const int t = (int)(it % n_steps);
    const long b = it / n_steps;
    const int tt = t < T ? t : T - 1;
    const double vv = per_scenario ? v[b * T + tt] : v[tt];
    const double A33 = -4 * Ca / (m * vv);
    const double A34 = (2 * Ca * (bb - a) / m * vv) - vv;
    const double A43 = 2 * Ca * ((bb - a) / (Jz * vv));
    const double A44 = -2 * Ca * (a * a + bb * bb) / (Jz * vv);
    const double B31 = 2 * Ca / m, B41 = 2 * Ca * a / Jz;
    double M[25] = {0, vv, 1, 0, 0,  0, 0, 0, 1, 0,  0, 0, A33, A34, B31,  0, 0, A43, A44, B41,  0, 0, 0, 0, 0}, E[25];
    for (int i = 0; i < 25; ++i) M[i] *= dt;
    mpcv::expm_small<5>(M, E);
    double* o = out + it * 20;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) o[i * 4 + j] = E[i * 5 + j];
    for (int i = 0; i < 4; ++i) o[16 + i] = E[i * 5 + 4];
  }
}

unsigned grid_for(long items) {
  const long g = (items + 127) / 128;
  return (unsigned)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}
int have_device() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return mpcv_set_error(-ENODEV, "no CUDA device: mpc_verde_b200 has no CPU path");
  return 0;
}

}  // namespace

extern "C" {

int mpcv_ref_lateral(const double* x, const double* y, int32_t nsim, const double* scale, int32_t Nt, double Delta,
                     double ar, double br, double* pwin, int64_t B, void* stream) {
  if (!x || !y || !pwin || nsim < 5 || Nt < 1) return mpcv_set_error(-EINVAL, "mpcv_ref_lateral: bad argument");
  if (int rc = have_device()) return rc;
  if (B <= 0) return 0;
  ref_lateral_kernel<<<grid_for(B * nsim * (long)Nt), 128, 0, (cudaStream_t)stream>>>(x, y, nsim, scale, Nt, Delta, ar, br, pwin, (long)B);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int mpcv_ref_frenet(const double* x, const double* y, const double* vdes, int32_t nsim, const double* scale, int32_t Nt,
                    double Delta, int32_t n_steps, double* pwin, int64_t B, void* stream) {
  if (!x || !y || !vdes || !pwin || nsim < 5 || Nt < 1 || n_steps < 1 || n_steps > nsim)
    return mpcv_set_error(-EINVAL, "mpcv_ref_frenet: bad argument");
  if (int rc = have_device()) return rc;
  if (B <= 0) return 0;
  ref_frenet_kernel<<<grid_for(B * n_steps * (long)Nt), 128, 0, (cudaStream_t)stream>>>(x, y, vdes, nsim, scale, Nt, Delta, n_steps, pwin, (long)B);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int mpcv_ref_unicycle_path(const double* x, const double* y, int32_t T, const double* scale, double dt, double vmax,
                           double wmax, double* ptraj, int64_t B, void* stream) {
  if (!x || !y || !ptraj || T < 3) return mpcv_set_error(-EINVAL, "mpcv_ref_unicycle_path: bad argument");
  if (int rc = have_device()) return rc;
  if (B <= 0) return 0;
  ref_unicycle_kernel<<<grid_for(B * (long)T), 128, 0, (cudaStream_t)stream>>>(x, y, T, scale, dt, vmax, wmax, ptraj, (long)B);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int mpcv_ref_circle(int32_t T, double Delta, double* ptraj, void* stream) {
  if (!ptraj || T < 1) return mpcv_set_error(-EINVAL, "mpcv_ref_circle: bad argument");
  if (int rc = have_device()) return rc;
  ref_circle_kernel<<<(T + 127) / 128, 128, 0, (cudaStream_t)stream>>>(T, Delta, ptraj);
  CUDA_OK(cudaGetLastError());
  return 0;
}

// a_end, b_end: the last sample of the base path (host values).  xt, yt, c2: device, `capacity` samples each.
// Returns the number of samples of the extended path in *n_out (call with xt = NULL to size the buffers).
int mpcv_path_lane_change_ext(const double* a, const double* b, const double* c, int32_t n0, double a_end, double b_end,
                              double v, double dt, double* xt, double* yt, double* c2, int32_t capacity, int32_t* n_out,
                              void* stream) {
  if (!n_out || n0 < 1 || !(v > 0) || !(dt > 0)) return mpcv_set_error(-EINVAL, "mpcv_path_lane_change_ext: bad argument");
  Plan plan;
  int out = n0;
  double ex = a_end, ey = b_end;      // end of the path so far
  auto arc = [&](int s, double r, double cy_off, double t0, double t1, int n) {
    Piece& q = plan.p[s];
    q.kind = 0; q.n = n; q.out0 = out; q.cx = ex; q.cy = ey + cy_off; q.r = r; q.t0 = t0; q.t1 = t1;
    out += n - 1;
    ex = q.cx + r * std::cos(t1);
    ey = q.cy + r * std::sin(t1);
  };
  auto straight = [&](int s, double t1, int n) {
    Piece& q = plan.p[s];
    q.kind = 1; q.n = n; q.out0 = out; q.cx = ex; q.cy = ey; q.r = v; q.t0 = 0.0; q.t1 = t1;
    out += n - 1;
    ex = q.cx - v * t1;
  };
  // half turn of 500 samples at speed v, a 10 m straight, two half turns of half the radius (an S), the straight
  // back to x = 0 and the closing half turn whose diameter is the remaining height
  const int k0 = 500;
  const double r = v / (M_PI / (k0 * dt));
  arc(0, r, r, 1.5 * M_PI, 2.5 * M_PI, k0);
  const double ds = 10.0;
  straight(1, ds, (int)(ds / (v * dt)));
  const int k4 = (int)(M_PI / ((v / (r / 2)) * dt));
  arc(2, r / 2, -r / 2, 0.5 * M_PI, 1.5 * M_PI, k4);
  arc(3, r / 2, -0.5 * r, 0.5 * M_PI, -0.5 * M_PI, k4);
  const double k6 = ex / (v * dt);
  straight(4, k6 * dt, (int)k6);
  const double r7 = ey / 2;
  arc(5, r7, -r7, 0.5 * M_PI, 1.5 * M_PI, (int)(M_PI / ((v / r7) * dt)));
  plan.total = out;
  *n_out = out;
  if (!xt) return 0;
  if (!a || !b || !c || !yt || !c2) return mpcv_set_error(-EINVAL, "mpcv_path_lane_change_ext: null argument");
  if (capacity < out) return mpcv_set_error(-EINVAL, "mpcv_path_lane_change_ext: output buffers too small");
  if (int rc = have_device()) return rc;
  path_ext_kernel<<<(out + 127) / 128, 128, 0, (cudaStream_t)stream>>>(plan, a, b, c, n0, v, xt, yt, c2);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int mpcv_ltv_lateral(const double* c, int32_t T, const double* spd, double ar, double br, double dt, int32_t n_steps,
                     double* pglob_traj, int64_t B, void* stream) {
  if (!c || !pglob_traj || T < 1 || n_steps < 1) return mpcv_set_error(-EINVAL, "mpcv_ltv_lateral: bad argument");
  if (int rc = have_device()) return rc;
  if (B <= 0) return 0;
  ltv_lateral_kernel<<<grid_for(B * (long)n_steps), 128, 0, (cudaStream_t)stream>>>(c, T, spd, ar, br, dt, n_steps, pglob_traj, (long)B);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int mpcv_ltv_dynbike(const double* v, int32_t T, int32_t per_scenario, const double* params, double dt, int32_t n_steps,
                     double* pglob_traj, int64_t B, void* stream) {
  if (!v || !pglob_traj || T < 1 || n_steps < 1) return mpcv_set_error(-EINVAL, "mpcv_ltv_dynbike: bad argument");
  if (int rc = have_device()) return rc;
  if (B <= 0) return 0;
  // m, a, b, Ca, Jz of Trajectory_tracking_dynamic_model.py:37-43 unless given
  const double m = params ? params[0] : 1200.0, a = params ? params[1] : 1.5, b = params ? params[2] : 2.0,
               Ca = params ? params[3] : 55000.0, Jz = params ? params[4] : 1350.0;
  ltv_dynbike_kernel<<<grid_for(B * (long)n_steps), 128, 0, (cudaStream_t)stream>>>(v, T, per_scenario, m, a, b, Ca, Jz, dt, n_steps, pglob_traj, (long)B);
  CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
