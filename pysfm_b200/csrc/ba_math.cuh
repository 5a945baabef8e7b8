// Device-side per-observation model of the bundle-adjustment path (FP64 throughout).
//
// What each function stands in for in the reference (paths relative to the pysfm tree):
//   observe()        bundle.py:243-277 (predict / reproj_error / residual / Jresidual),
//                    bundle.py:8-11 (Jpr), algebra.py:5-8 (pr), lie.py:38-40 (J_expm_x)
//   sensor_apply()   sensor_model.py:23-29 (Gaussian), :48-69 (Cauchy)
//   so3_exp()        lie.py:21-34
//   sym3_pinv()      numpy.linalg.pinv(HPP, rcond) at bundle_adjuster.py:252-256
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace ba {

struct ModelParams {
  int kind;        // BA_MODEL_GAUSSIAN / BA_MODEL_CAUCHY
  double p[4];     // Gaussian: row-major L; Cauchy: sigma, sigma^2, linear window, -
};

struct Intrinsics {
  double K[9];     // full 3x3, not assumed upper triangular (the reference fixture has K[1][0] != 0)
};

// r = model(e), Jr = d r / d e  (2x2 row-major)
__device__ __forceinline__ void sensor_apply(const ModelParams& m, double e0, double e1,
                                             double r[2], double Jr[4]) {
  if (m.kind == 0) {
    Jr[0] = m.p[0]; Jr[1] = m.p[1]; Jr[2] = m.p[2]; Jr[3] = m.p[3];
    r[0] = Jr[0] * e0 + Jr[1] * e1;
    r[1] = Jr[2] * e0 + Jr[3] * e1;
  } else {
    const double sigma = m.p[0], sig2 = m.p[1], window = m.p[2];
    const double rho2 = e0 * e0 + e1 * e1;
    const double rho = sqrt(rho2);
    if (rho < window) {
      const double is = 1.0 / sigma;
      Jr[0] = is; Jr[1] = 0.0; Jr[2] = 0.0; Jr[3] = is;
      r[0] = e0 * is; r[1] = e1 * is;
    } else {
      const double s = sqrt(log(1.0 + rho2 / sig2));
      const double g = s / rho;
      r[0] = e0 * g; r[1] = e1 * g;
      // J = e e^T / (rho s (rho^2 + sigma^2)) + (rho I - e e^T / rho) s / rho^2
      const double a = 1.0 / (rho * s * (rho2 + sig2));
      const double b = s / rho2;
      const double e00 = e0 * e0, e01 = e0 * e1, e11 = e1 * e1;
      Jr[0] = e00 * a + (rho - e00 / rho) * b;
      Jr[1] = e01 * a + (-e01 / rho) * b;
      Jr[2] = Jr[1];
      Jr[3] = e11 * a + (rho - e11 / rho) * b;
    }
  }
}

// Residual only (cost evaluation of a candidate).
__device__ __forceinline__ void residual_only(const Intrinsics& in, const ModelParams& m,
                                              const double* __restrict__ R,
                                              const double* __restrict__ t, const double x[3],
                                              double u, double v, double r[2]) {
  const double y0 = R[0] * x[0] + R[1] * x[1] + R[2] * x[2] + t[0];
  const double y1 = R[3] * x[0] + R[4] * x[1] + R[5] * x[2] + t[1];
  const double y2 = R[6] * x[0] + R[7] * x[1] + R[8] * x[2] + t[2];
  const double* K = in.K;
  const double p0 = K[0] * y0 + K[1] * y1 + K[2] * y2;
  const double p1 = K[3] * y0 + K[4] * y1 + K[5] * y2;
  const double p2 = K[6] * y0 + K[7] * y1 + K[8] * y2;
  double Jr[4];
  const double ip2 = 1.0 / p2;
  sensor_apply(m, p0 * ip2 - u, p1 * ip2 - v, r, Jr);
}

// Full per-observation linearisation.
//   r[2]; Jc[12] = 2x6 row-major, columns [rotation(3) | translation(3)]; Jp[6] = 2x3 row-major.
__device__ __forceinline__ void observe(const Intrinsics& in, const ModelParams& m,
                                        const double* __restrict__ R,
                                        const double* __restrict__ t, const double x[3],
                                        double u, double v, double r[2], double Jc[12],
                                        double Jp[6]) {
  const double y0 = R[0] * x[0] + R[1] * x[1] + R[2] * x[2] + t[0];
  const double y1 = R[3] * x[0] + R[4] * x[1] + R[5] * x[2] + t[1];
  const double y2 = R[6] * x[0] + R[7] * x[1] + R[8] * x[2] + t[2];
  const double* K = in.K;
  const double p0 = K[0] * y0 + K[1] * y1 + K[2] * y2;
  const double p1 = K[3] * y0 + K[4] * y1 + K[5] * y2;
  const double p2 = K[6] * y0 + K[7] * y1 + K[8] * y2;
  const double ip2 = 1.0 / p2;
  // Jpr (2x3) = [[1/p2, 0, -p0/p2^2], [0, 1/p2, -p1/p2^2]]  (one division, two multiplies: the two
  // extra FP64 divisions were ~60 instructions on the per-observation dependency chain)
  const double ip22 = ip2 * ip2;
  const double a02 = -p0 * ip22, a12 = -p1 * ip22;
  // Jt = Jpr K
  double Jt[6];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    Jt[c] = ip2 * K[c] + a02 * K[6 + c];
    Jt[3 + c] = ip2 * K[3 + c] + a12 * K[6 + c];
  }
  // Jx = Jt R
  double Jx[6];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      Jx[a * 3 + c] = Jt[a * 3] * R[c] + Jt[a * 3 + 1] * R[3 + c] + Jt[a * 3 + 2] * R[6 + c];
  // JR = Jx hat(-x),  hat(-x) = [[0, x2, -x1], [-x2, 0, x0], [x1, -x0, 0]]
  double JR[6];
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    JR[a * 3 + 0] = -Jx[a * 3 + 1] * x[2] + Jx[a * 3 + 2] * x[1];
    JR[a * 3 + 1] = Jx[a * 3 + 0] * x[2] - Jx[a * 3 + 2] * x[0];
    JR[a * 3 + 2] = -Jx[a * 3 + 0] * x[1] + Jx[a * 3 + 1] * x[0];
  }
  double Jr[4];
  sensor_apply(m, p0 * ip2 - u, p1 * ip2 - v, r, Jr);
  // chain rule: rows of [JR | Jt | Jx] mixed by Jr
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    Jc[c] = Jr[0] * JR[c] + Jr[1] * JR[3 + c];
    Jc[6 + c] = Jr[2] * JR[c] + Jr[3] * JR[3 + c];
    Jc[3 + c] = Jr[0] * Jt[c] + Jr[1] * Jt[3 + c];
    Jc[9 + c] = Jr[2] * Jt[c] + Jr[3] * Jt[3 + c];
    Jp[c] = Jr[0] * Jx[c] + Jr[1] * Jx[3 + c];
    Jp[3 + c] = Jr[2] * Jx[c] + Jr[3] * Jx[3 + c];
  }
}

// Rodrigues; identity below 1e-8 rad like the reference.
__device__ __forceinline__ void so3_exp(const double m[3], double E[9]) {
  const double th2 = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
  const double th = sqrt(th2);
  E[0] = 1.0; E[1] = 0.0; E[2] = 0.0;
  E[3] = 0.0; E[4] = 1.0; E[5] = 0.0;
  E[6] = 0.0; E[7] = 0.0; E[8] = 1.0;
  if (th < 1e-8) return;
  const double A = sin(th) / th;
  const double B = (1.0 - cos(th)) / th2;
  const double W[9] = {0.0, -m[2], m[1], m[2], 0.0, -m[0], -m[1], m[0], 0.0};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double w2 = W[i * 3] * W[j] + W[i * 3 + 1] * W[3 + j] + W[i * 3 + 2] * W[6 + j];
      E[i * 3 + j] += A * W[i * 3 + j] + B * w2;
    }
}

// cand = R * exp(d[0:3]),  tc = t + d[3:6]   (Camera.perturb, bundle.py:76-80)
__device__ __forceinline__ void camera_retract(const double* __restrict__ R,
                                               const double* __restrict__ t, const double d[6],
                                               double* __restrict__ Rc, double* __restrict__ tc) {
  double E[9];
  so3_exp(d, E);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      Rc[i * 3 + j] = R[i * 3] * E[j] + R[i * 3 + 1] * E[3 + j] + R[i * 3 + 2] * E[6 + j];
  tc[0] = t[0] + d[3]; tc[1] = t[1] + d[4]; tc[2] = t[2] + d[5];
}

// One Jacobi rotation zeroing A[P][Q_] of the symmetric 3x3 A, accumulating the eigenvector
// matrix Q (columns = eigenvectors).
template <int P, int Q_>
__device__ __forceinline__ void jacobi_rotate(double A[3][3], double Q[3][3]) {
  const double apq = A[P][Q_];
  if (apq == 0.0) return;
  const double tau = (A[Q_][Q_] - A[P][P]) / (2.0 * apq);
  const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
  const double c = 1.0 / sqrt(1.0 + tt * tt);
  const double s = tt * c;
  constexpr int Rr = 3 - P - Q_;  // the remaining index
  const double app = A[P][P], aqq = A[Q_][Q_];
  A[P][P] = app - tt * apq;
  A[Q_][Q_] = aqq + tt * apq;
  A[P][Q_] = 0.0; A[Q_][P] = 0.0;
  const double arp = A[Rr][P], arq = A[Rr][Q_];
  A[Rr][P] = c * arp - s * arq; A[P][Rr] = A[Rr][P];
  A[Rr][Q_] = s * arp + c * arq; A[Q_][Rr] = A[Rr][Q_];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double qp = Q[i][P], qq = Q[i][Q_];
    Q[i][P] = c * qp - s * qq;
    Q[i][Q_] = s * qp + c * qq;
  }
}

// Inverse of a symmetric 3x3 by cofactors (only the upper triangle of V is read); the result
// is exactly symmetric.
__device__ __forceinline__ void sym3_inv(const double V[9], double det, double out[9]) {
  const double id = 1.0 / det;
  const double c00 = V[4] * V[8] - V[5] * V[5];
  const double c01 = V[2] * V[5] - V[1] * V[8];
  const double c02 = V[1] * V[5] - V[2] * V[4];
  const double c11 = V[0] * V[8] - V[2] * V[2];
  const double c12 = V[1] * V[2] - V[0] * V[5];
  const double c22 = V[0] * V[4] - V[1] * V[1];
  out[0] = c00 * id; out[1] = c01 * id; out[2] = c02 * id;
  out[3] = out[1];   out[4] = c11 * id; out[5] = c12 * id;
  out[6] = out[2];   out[7] = out[5];   out[8] = c22 * id;
}

// Pseudo-inverse of a symmetric positive semi-definite 3x3 (V[9] row-major, symmetric) with
// numpy.linalg.pinv's rule: eigen-directions whose |eigenvalue| <= rcond * max|eigenvalue|
// are dropped.  rcond < 0: plain inverse (numpy.linalg.inv branch of the reference).
//
// Fast path: for eigenvalues l1 <= l2 <= l3 of a PSD matrix, l1 = det/(l2 l3) >= 4 det/tr^2
// and l3 <= tr, so det > rcond tr^3 proves l1 > rcond l3: nothing is truncated and the
// pseudo-inverse IS the inverse.  Only blocks that fail this (conservative) test pay for the
// Jacobi eigen-decomposition.
__device__ __forceinline__ void sym3_pinv(const double V[9], double rcond, double out[9]) {
  const double det = V[0] * (V[4] * V[8] - V[5] * V[5]) + V[1] * (V[2] * V[5] - V[1] * V[8]) +
                     V[2] * (V[1] * V[5] - V[2] * V[4]);
  if (rcond < 0.0) {
    sym3_inv(V, det, out);
    return;
  }
  const double tr = V[0] + V[4] + V[8];
  if (det > rcond * tr * tr * tr) {
    sym3_inv(V, det, out);
    return;
  }
  double A[3][3] = {{V[0], V[1], V[2]}, {V[1], V[4], V[5]}, {V[2], V[5], V[8]}};
  double Q[3][3] = {{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 0.0, 1.0}};
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    const double dg = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off <= 1e-36 * dg || off == 0.0) break;
    jacobi_rotate<0, 1>(A, Q);
    jacobi_rotate<0, 2>(A, Q);
    jacobi_rotate<1, 2>(A, Q);
  }
  const double w0 = A[0][0], w1 = A[1][1], w2 = A[2][2];
  const double wmax = fmax(fabs(w0), fmax(fabs(w1), fabs(w2)));
  const double cut = rcond * wmax;
  const double i0 = fabs(w0) > cut ? 1.0 / w0 : 0.0;
  const double i1 = fabs(w1) > cut ? 1.0 / w1 : 0.0;
  const double i2 = fabs(w2) > cut ? 1.0 / w2 : 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j) {
      const double v = Q[i][0] * i0 * Q[j][0] + Q[i][1] * i1 * Q[j][1] + Q[i][2] * i2 * Q[j][2];
      out[i * 3 + j] = v;
      out[j * 3 + i] = v;
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace ba
