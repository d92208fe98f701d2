// Leaderboard metrics of the Map-free benchmark for a batch of poses (one warp per pose):
//   trans_err  = |t_est - t_gt|                                   benchmark/metrics.py:49-50
//   rot_err    = 2 asin(|vec(q_gt q_est^-1)|) in degrees          benchmark/metrics.py:54-55, utils.py:95-129 (sine variant)
//   reproj_err = mean over a fixed 7 x 4 x 7 grid of virtual 3-D points of the pixel distance between their projection
//                and the projection after the residual transform inv(T_est) T_gt, clamped to the image ("VCRE")
//                                                                   benchmark/reprojection.py:7-86
// fp64 like the reference's numpy code. The aggregation (medians, precision, AUC) lives in mfr_b200/metrics.py.
#include "common.cuh"

namespace mfr {

namespace {

__device__ __forceinline__ void quat2mat_t3d(const double* q, double* R) {   // transforms3d.quaternions.quat2mat
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double Nq = w * w + x * x + y * y + z * z;
  if (Nq < 2.220446049250313e-16) {
    for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  const double s = 2.0 / Nq;
  const double X = x * s, Y = y * s, Z = z * s;
  const double wX = w * X, wY = w * Y, wZ = w * Z, xX = x * X, xY = x * Y, xZ = x * Z, yY = y * Y, yZ = y * Z, zZ = z * Z;
  R[0] = 1.0 - (yY + zZ); R[1] = xY - wZ; R[2] = xZ + wY;
  R[3] = xY + wZ; R[4] = 1.0 - (xX + zZ); R[5] = yZ - wX;
  R[6] = xZ - wY; R[7] = yZ + wX; R[8] = 1.0 - (xX + yY);
}

__global__ void __launch_bounds__(128) pose_metrics_kernel(const double* __restrict__ q_gt, const double* __restrict__ t_gt,
                                                           const double* __restrict__ q_est, const double* __restrict__ t_est,
                                                           const double* __restrict__ K, int W, int H, int n,
                                                           double* __restrict__ trans_err, double* __restrict__ rot_err,
                                                           double* __restrict__ reproj_err) {
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const double *qg = q_gt + 4 * i, *qe = q_est + 4 * i, *tg = t_gt + 3 * i, *te = t_est + 3 * i, *Kp = K + 9 * i;
  if (lane == 0) {
    const double dx = te[0] - tg[0], dy = te[1] - tg[1], dz = te[2] - tg[2];
    trans_err[i] = sqrt(dx * dx + dy * dy + dz * dz);
    // utils.py:113-125 with label = q_est, pred = q_gt: q1 = pred / |pred|, q2 = label / |label|, sine = q1 * q2^-1
    const double n1 = sqrt(qg[0] * qg[0] + qg[1] * qg[1] + qg[2] * qg[2] + qg[3] * qg[3]);
    const double n2 = sqrt(qe[0] * qe[0] + qe[1] * qe[1] + qe[2] * qe[2] + qe[3] * qe[3]);
    const double a[4] = {qg[0] / n1, qg[1] / n1, qg[2] / n1, qg[3] / n1};
    double b[4] = {qe[0] / n2, qe[1] / n2, qe[2] / n2, qe[3] / n2};
    const double nb = b[0] * b[0] + b[1] * b[1] + b[2] * b[2] + b[3] * b[3];     // qinverse = conjugate / |q|^2
    b[0] /= nb; b[1] /= -nb; b[2] /= -nb; b[3] /= -nb;
    const double vx = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    const double vy = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    const double vz = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    rot_err[i] = asin(fmin(sqrt(vx * vx + vy * vy + vz * vz), 1.0)) * 114.59155902616465;
  }
  double Re[9], Rg[9], Rr[9], tr[3];
  quat2mat_t3d(qe, Re);
  quat2mat_t3d(qg, Rg);
  // residual = inv([Re | te]) [Rg | tg] = [Re^T Rg | Re^T (tg - te)]   (general inverse of a rigid 4x4 with rotation block Re;
  // like np.linalg.inv this holds for the matrices quat2mat returns for non-unit quaternions only up to their orthogonality,
  // which quat2mat guarantees by its 2 / Nq scaling)
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) Rr[3 * r + c] = Re[r] * Rg[c] + Re[3 + r] * Rg[3 + c] + Re[6 + r] * Rg[6 + c];
    tr[r] = Re[r] * (tg[0] - te[0]) + Re[3 + r] * (tg[1] - te[1]) + Re[6 + r] * (tg[2] - te[2]);
  }
  auto project = [&](double X, double Y, double Z, double& u, double& v) {
    const double a0 = Kp[0] * X + Kp[1] * Y + Kp[2] * Z, a1 = Kp[3] * X + Kp[4] * Y + Kp[5] * Z, a2 = Kp[6] * X + Kp[7] * Y + Kp[8] * Z;
    u = fmin(fmax(a0 / a2, 0.0), static_cast<double>(W));
    v = fmin(fmax(a1 / a2, 0.0), static_cast<double>(H));
  };
  double acc = 0.0;
  for (int k = lane; k < 196; k += 32) {          // reprojection.py:33-55: 7 (x) x 4 (y) x 7 (z) points, step 0.3 m, z from 1.8 m
    const int iz = k % 7, ix = (k / 7) % 7, iy = k / 49;
    const double X = (ix - 3.0) * 0.3, Y = (iy - 1.5) * 0.3, Z = iz * 0.3 + 1.8;
    double u0, v0, u1, v1;
    project(X, Y, Z, u0, v0);
    project(Rr[0] * X + Rr[1] * Y + Rr[2] * Z + tr[0], Rr[3] * X + Rr[4] * Y + Rr[5] * Z + tr[1], Rr[6] * X + Rr[7] * Y + Rr[8] * Z + tr[2], u1, v1);
    acc += sqrt((u0 - u1) * (u0 - u1) + (v0 - v1) * (v0 - v1));
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) reproj_err[i] = acc / 196.0;
}

}  // namespace

int pose_metrics(const double* q_gt, const double* t_gt, const double* q_est, const double* t_est, const double* K, int W, int H,
                 int n, double* trans_err, double* rot_err, double* reproj_err, cudaStream_t st) {
  if (n <= 0) return MFR_OK;
  pose_metrics_kernel<<<(n + 3) / 4, 128, 0, st>>>(q_gt, t_gt, q_est, t_est, K, W, H, n, trans_err, rot_err, reproj_err);
  MFR_CUDA_CHECK(cudaGetLastError());
  return MFR_OK;
}

}  // namespace mfr
