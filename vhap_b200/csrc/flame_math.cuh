// Per-frame FLAME pose math: axis-angle -> rotation (with the reference's 1e-8-inside-the-norm quirk), the 5-joint
// kinematic chain and their analytic adjoints.  Restates vhap/model/lbs.py: batch_rodrigues :25-57,
// batch_rigid_transform :254-304 (parents = [-1,0,1,1,1], flame.py:119-121).  Oracle: oracle/lbs.py.
#pragma once
#include "common.cuh"

#define VH_NJ 5
VH_HD int vh_parent(int j) { return j == 0 ? -1 : (j == 1 ? 0 : 1); }

struct PoseFwd {
  float R[VH_NJ][9];      // per-joint rotation
  float J[VH_NJ][3];      // rest joints
  float GR[VH_NJ][9];     // chained rotation
  float Gt[VH_NJ][3];     // chained translation
  float A[VH_NJ][12];     // rel transforms, row-major 3x4
  float pf[36];           // pose feature (R_j - I), j = 1..4
};

VH_HD void mat3_mul(const float* a, const float* b, float* c) {
  VH_UNROLL for (int i = 0; i < 3; ++i) VH_UNROLL for (int j = 0; j < 3; ++j)
    c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}
VH_HD void mat3_mul_bt(const float* a, const float* b, float* c) {   // a * b^T
  VH_UNROLL for (int i = 0; i < 3; ++i) VH_UNROLL for (int j = 0; j < 3; ++j)
    c[i * 3 + j] = a[i * 3] * b[j * 3] + a[i * 3 + 1] * b[j * 3 + 1] + a[i * 3 + 2] * b[j * 3 + 2];
}
VH_HD void mat3_mul_at(const float* a, const float* b, float* c) {   // a^T * b
  VH_UNROLL for (int i = 0; i < 3; ++i) VH_UNROLL for (int j = 0; j < 3; ++j)
    c[i * 3 + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}

VH_HD void rodrigues(const float* r, float* R) {
  float ex = r[0] + 1e-8f, ey = r[1] + 1e-8f, ez = r[2] + 1e-8f;       // lbs.py:40
  float th = sqrtf(ex * ex + ey * ey + ez * ez);
  float nx = r[0] / th, ny = r[1] / th, nz = r[2] / th;
  float s = sinf(th), c = cosf(th);
  float K[9] = {0, -nz, ny, nz, 0, -nx, -ny, nx, 0}, M[9];
  mat3_mul(K, K, M);
  VH_UNROLL for (int i = 0; i < 9; ++i) R[i] = s * K[i] + (1.f - c) * M[i];
  R[0] += 1.f; R[4] += 1.f; R[8] += 1.f;
}

VH_HD void rodrigues_bwd(const float* r, const float* gR, float* g_r) {
  float ex = r[0] + 1e-8f, ey = r[1] + 1e-8f, ez = r[2] + 1e-8f;
  float th = sqrtf(ex * ex + ey * ey + ez * ez);
  float nx = r[0] / th, ny = r[1] / th, nz = r[2] / th;
  float s = sinf(th), c = cosf(th);
  float K[9] = {0, -nz, ny, nz, 0, -nx, -ny, nx, 0}, M[9];
  mat3_mul(K, K, M);
  float g_s = 0, g_c = 0;
  VH_UNROLL for (int i = 0; i < 9; ++i) { g_s += gR[i] * K[i]; g_c -= gR[i] * M[i]; }
  float t1[9], t2[9], gK[9];
  mat3_mul_bt(gR, K, t1);      // gR K^T
  mat3_mul_at(K, gR, t2);      // K^T gR
  VH_UNROLL for (int i = 0; i < 9; ++i) gK[i] = s * gR[i] + (1.f - c) * (t1[i] + t2[i]);
  float g_th = g_s * c - g_c * s;
  float gnx = gK[7] - gK[5], gny = gK[2] - gK[6], gnz = gK[3] - gK[1];
  float ith = 1.f / th;
  g_th -= (gnx * r[0] + gny * r[1] + gnz * r[2]) * ith * ith;
  g_r[0] = gnx * ith + g_th * ex * ith;
  g_r[1] = gny * ith + g_th * ey * ith;
  g_r[2] = gnz * ith + g_th * ez * ith;
}

// pose[15] = (rotation, neck, jaw, eye_l, eye_r); J = rest joints
VH_HD void pose_forward(const float* pose, PoseFwd& f) {
  VH_UNROLL for (int j = 0; j < VH_NJ; ++j) rodrigues(pose + 3 * j, f.R[j]);
  VH_UNROLL for (int j = 0; j < VH_NJ; ++j) {
    int p = vh_parent(j);
    float rel[3];
    VH_UNROLL for (int c = 0; c < 3; ++c) rel[c] = f.J[j][c] - (p >= 0 ? f.J[p][c] : 0.f);
    if (p < 0) {
      VH_UNROLL for (int i = 0; i < 9; ++i) f.GR[j][i] = f.R[j][i];
      VH_UNROLL for (int c = 0; c < 3; ++c) f.Gt[j][c] = rel[c];
    } else {
      mat3_mul(f.GR[p], f.R[j], f.GR[j]);
      VH_UNROLL for (int c = 0; c < 3; ++c)
        f.Gt[j][c] = f.GR[p][c * 3] * rel[0] + f.GR[p][c * 3 + 1] * rel[1] + f.GR[p][c * 3 + 2] * rel[2] + f.Gt[p][c];
    }
  }
  VH_UNROLL for (int j = 0; j < VH_NJ; ++j)
    VH_UNROLL for (int r = 0; r < 3; ++r) {
      const float* g = f.GR[j] + r * 3;
      f.A[j][r * 4 + 0] = g[0]; f.A[j][r * 4 + 1] = g[1]; f.A[j][r * 4 + 2] = g[2];
      f.A[j][r * 4 + 3] = f.Gt[j][r] - (g[0] * f.J[j][0] + g[1] * f.J[j][1] + g[2] * f.J[j][2]);
    }
  VH_UNROLL for (int j = 1; j < VH_NJ; ++j)
    VH_UNROLL for (int i = 0; i < 9; ++i) f.pf[(j - 1) * 9 + i] = f.R[j][i] - ((i == 0 || i == 4 || i == 8) ? 1.f : 0.f);
}

// gA[5][12], gpf[36] -> g_pose[15], g_J[5][3]
VH_HD void pose_backward(const float* pose, const PoseFwd& f, const float gA[VH_NJ][12], const float* gpf,
                         float* g_pose, float g_J[VH_NJ][3]) {
  float gGR[VH_NJ][9], gGt[VH_NJ][3], gR[VH_NJ][9];
  VH_UNROLL for (int j = 0; j < VH_NJ; ++j) {
    VH_UNROLL for (int c = 0; c < 3; ++c) g_J[j][c] = 0.f;
    VH_UNROLL for (int i = 0; i < 9; ++i) gR[j][i] = 0.f;
  }
  VH_UNROLL for (int j = 0; j < VH_NJ; ++j)
    VH_UNROLL for (int r = 0; r < 3; ++r) {
      float gt = gA[j][r * 4 + 3];
      gGt[j][r] = gt;
      VH_UNROLL for (int c = 0; c < 3; ++c) {
        gGR[j][r * 3 + c] = gA[j][r * 4 + c] - gt * f.J[j][c];
        g_J[j][c] -= f.GR[j][r * 3 + c] * gt;
      }
    }
  VH_UNROLL for (int j = VH_NJ - 1; j >= 0; --j) {
    int p = vh_parent(j);
    float rel[3], g_rel[3];
    VH_UNROLL for (int c = 0; c < 3; ++c) rel[c] = f.J[j][c] - (p >= 0 ? f.J[p][c] : 0.f);
    if (p < 0) {
      VH_UNROLL for (int i = 0; i < 9; ++i) gR[j][i] += gGR[j][i];
      VH_UNROLL for (int c = 0; c < 3; ++c) g_rel[c] = gGt[j][c];
    } else {
      float t[9];
      mat3_mul_bt(gGR[j], f.R[j], t);                 // g_GpR += gGR_j R_j^T
      VH_UNROLL for (int i = 0; i < 9; ++i) gGR[p][i] += t[i];
      mat3_mul_at(f.GR[p], gGR[j], t);                // g_R_j += GpR^T gGR_j
      VH_UNROLL for (int i = 0; i < 9; ++i) gR[j][i] += t[i];
      VH_UNROLL for (int r = 0; r < 3; ++r) VH_UNROLL for (int c = 0; c < 3; ++c) gGR[p][r * 3 + c] += gGt[j][r] * rel[c];
      VH_UNROLL for (int c = 0; c < 3; ++c) {
        g_rel[c] = f.GR[p][c] * gGt[j][0] + f.GR[p][3 + c] * gGt[j][1] + f.GR[p][6 + c] * gGt[j][2];
        gGt[p][c] += gGt[j][c];
      }
    }
    VH_UNROLL for (int c = 0; c < 3; ++c) { g_J[j][c] += g_rel[c]; if (p >= 0) g_J[p][c] -= g_rel[c]; }
  }
  VH_UNROLL for (int j = 1; j < VH_NJ; ++j) VH_UNROLL for (int i = 0; i < 9; ++i) gR[j][i] += gpf[(j - 1) * 9 + i];
  VH_UNROLL for (int j = 0; j < VH_NJ; ++j) rodrigues_bwd(pose + 3 * j, gR[j], g_pose + 3 * j);
}
