// 3x3 SVD-based rotation fit shared by the ego-motion head (ego.cu) and the ICP refinement (nn_grid.cu).
#pragma once

// one-sided Jacobi SVD of a 3x3 (double): M = U diag(S) V^T.  Returns R = V diag(1,1,det(V^T U^T)) U^T.
__device__ inline void kabsch_rotation(const double C[3][3], double R[3][3]) {
  double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = C[i][j];
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double a = 0, b = 0, g = 0;
        for (int i = 0; i < 3; ++i) a += A[i][p] * A[i][p], b += A[i][q] * A[i][q], g += A[i][p] * A[i][q];
        off += g * g;
        if (fabs(g) < 1e-300 || fabs(g) <= 1e-17 * sqrt(a * b)) continue;
        double zeta = (b - a) / (2.0 * g);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
        for (int i = 0; i < 3; ++i) {
          double x = A[i][p], y = A[i][q];
          A[i][p] = cs * x - sn * y, A[i][q] = sn * x + cs * y;
          x = V[i][p], y = V[i][q];
          V[i][p] = cs * x - sn * y, V[i][q] = sn * x + cs * y;
        }
      }
    if (off < 1e-60) break;
  }
  // column norms = singular values; order descending
  double s[3];
  int idx[3] = {0, 1, 2};
  for (int j = 0; j < 3; ++j) s[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
  for (int a = 0; a < 2; ++a)
    for (int b = a + 1; b < 3; ++b)
      if (s[idx[b]] > s[idx[a]]) {
        int t = idx[a];
        idx[a] = idx[b], idx[b] = t;
      }
  double U[3][3], Vs[3][3];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) Vs[i][j] = V[i][idx[j]];
  for (int j = 0; j < 2; ++j) {
    double nrm = s[idx[j]];
    for (int i = 0; i < 3; ++i) U[i][j] = nrm > 0 ? A[i][idx[j]] / nrm : (i == j ? 1.0 : 0.0);
  }
  if (s[idx[2]] > 1e-12 * s[idx[0]] && s[idx[2]] > 0) {
    for (int i = 0; i < 3; ++i) U[i][2] = A[i][idx[2]] / s[idx[2]];
  } else {  // rank deficient: complete the basis (the determinant factor below removes the sign choice)
    U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
    U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
    U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
  }
  auto det3 = [](const double M[3][3]) {
    return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
           M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
  };
  double d = det3(Vs) * det3(U);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i][j] = Vs[i][0] * U[j][0] + Vs[i][1] * U[j][1] + d * Vs[i][2] * U[j][2];
}
