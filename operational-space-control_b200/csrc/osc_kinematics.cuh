// osc_kinematics.cuh -- the step BEFORE the hot path on the device:
// what the reference reads from MuJoCo in update_mj_data / update_osc_data
// (walter_sr/operational_space_controller.h:394-513) for a floating-base tree of hinge joints,
//   M = mj_fullM (:436-438), C = qfrc_bias (:441-442), J = [jacp; jacr] per site (:459-487),
//   bias = Jdot qvel (:491-492),
// from qpos / qvel and a host-supplied model (osc_kin_model), written straight into the
// controller's input buffers in the OSCData layouts -- 13.4 kB per environment that then never
// cross PCIe.  MuJoCo's conventions: qvel = [world-frame base linear velocity, body-frame base
// angular velocity, hinge rates]; hinge anchors at the child body's origin.
// The robots' MJCF models are external to the reference and absent here, so the numbers of a
// real robot have to come from the user; the arithmetic is pinned by oracle/osc_kinematics.py,
// itself pinned on finite differences (tests/test_kinematics_oracle.py).
//
// One warp per environment: lanes = bodies for the frames / velocity recursions (one tree
// level per round), then lanes stride over the OUTPUT ENTRIES in memory order, so every store
// is coalesced and nothing is staged: an entry of J is one component of axis x (x - anchor),
// an entry of M a sum over the bodies of two small dot products.  HBM-write bound.
#pragma once

#include "../../include/osc_b200.h"

namespace osc {

template <class D>
struct alignas(16) KinWorkspace {
  static constexpr int NB = D::NV - 5, NV = D::NV, NS = D::NS;
  double R[NB][9], p[NB][3], w[NB][3], al[NB][3], a[NB][3], com[NB][3], Iw[NB][6];
  double F[NB][3], N[NB][3];
  double axis[NV][3], anchor[NV][3];
  double site[NS][3];
  double JC[NB][NV][3];   // column d of the point Jacobian at body b's centre of mass (0 if d does not move b)
  double IA[NB][NV][3];   // I_b axis_d for the rotational dofs that move b (else 0)
  unsigned aff[NB];       // bit d: dof d moves body b
  int depth[NB];
};

__device__ __forceinline__ void kin_cross(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ void kin_quat_mat(const double* q, double* R) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
__device__ __forceinline__ void kin_mm(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
__device__ __forceinline__ void kin_mv(const double* A, const double* x, double* y) {
#pragma unroll
  for (int i = 0; i < 3; ++i) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
}
// symmetric 3x3 (xx yy zz xy xz yz) times vector
__device__ __forceinline__ void kin_sv(const double* S, const double* x, double* y) {
  y[0] = S[0] * x[0] + S[3] * x[1] + S[4] * x[2];
  y[1] = S[3] * x[0] + S[1] * x[1] + S[5] * x[2];
  y[2] = S[4] * x[0] + S[5] * x[1] + S[2] * x[2];
}

constexpr int kKinWarps = 8;

template <class D>
__global__ void __launch_bounds__(kKinWarps * 32)
kinematics_kernel(const __grid_constant__ osc_kin_model m, const double* __restrict__ qpos,
                  const double* __restrict__ qvel, double* __restrict__ Mout,
                  double* __restrict__ Cout, double* __restrict__ Jout,
                  double* __restrict__ bias_out, int n_envs) {
  using WS = KinWorkspace<D>;
  constexpr int NB = WS::NB, NV = D::NV, NS = D::NS, NQ = NV + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WS& w = reinterpret_cast<WS*>(smem_raw)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  // model-only quantities of this warp's workspace: tree depth and dof masks
  if (lane < NB) {
    int d = 0;
    unsigned aff = 0x3fu;
    for (int b = lane; b > 0; b = m.parent[b]) {
      ++d;
      aff |= 1u << (5 + b);
    }
    w.depth[lane] = d;
    w.aff[lane] = aff;
  }
  __syncwarp();
  int maxd = 0;
  for (int b = 0; b < NB; ++b) maxd = w.depth[b] > maxd ? w.depth[b] : maxd;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; env < n_envs; env += warps) {
    const double* q = qpos + (size_t)env * NQ;
    const double* v = qvel + (size_t)env * NV;
    __syncwarp();
    // ---- frames, velocities, accelerations at qacc = 0: one tree level per round
    for (int lev = 0; lev <= maxd; ++lev) {
      if (lane < NB && w.depth[lane] == lev) {
        const int b = lane;
        if (b == 0) {
          double qq[4] = {q[3], q[4], q[5], q[6]};
          const double inv = rsqrt(qq[0] * qq[0] + qq[1] * qq[1] + qq[2] * qq[2] + qq[3] * qq[3]);
          for (int k = 0; k < 4; ++k) qq[k] *= inv;
          kin_quat_mat(qq, w.R[0]);
          const double wl[3] = {v[3], v[4], v[5]};
          kin_mv(w.R[0], wl, w.w[0]);
          for (int k = 0; k < 3; ++k) {
            w.p[0][k] = q[k];
            w.al[0][k] = 0.0;
            w.a[0][k] = 0.0;
          }
        } else {
          const int pa = m.parent[b];
          double Rq[9], Rj[9], T[9];
          kin_quat_mat(m.bquat[b], Rq);
          const double th = q[6 + b], c = cos(th), s = sin(th);
          const double x = m.jaxis[b][0], y = m.jaxis[b][1], z = m.jaxis[b][2], t = 1 - c;
          Rj[0] = c + t * x * x; Rj[1] = t * x * y - s * z; Rj[2] = t * x * z + s * y;
          Rj[3] = t * x * y + s * z; Rj[4] = c + t * y * y; Rj[5] = t * y * z - s * x;
          Rj[6] = t * x * z - s * y; Rj[7] = t * y * z + s * x; Rj[8] = c + t * z * z;
          kin_mm(w.R[pa], Rq, T);
          kin_mm(T, Rj, w.R[b]);
          double r[3], ax[3], wj[3], t1[3], t2[3];
          kin_mv(w.R[pa], m.bpos[b], r);
          kin_mv(w.R[b], m.jaxis[b], ax);
          for (int k = 0; k < 3; ++k) {
            w.p[b][k] = w.p[pa][k] + r[k];
            w.axis[5 + b][k] = ax[k];
            wj[k] = ax[k] * v[5 + b];
            w.w[b][k] = w.w[pa][k] + wj[k];
          }
          kin_cross(w.w[pa], wj, t1);
          for (int k = 0; k < 3; ++k) w.al[b][k] = w.al[pa][k] + t1[k];
          kin_cross(w.al[pa], r, t1);
          kin_cross(w.w[pa], r, t2);
          double t3[3];
          kin_cross(w.w[pa], t2, t3);
          for (int k = 0; k < 3; ++k) w.a[b][k] = w.a[pa][k] + t1[k] + t3[k];
        }
      }
      __syncwarp();
    }
    // ---- per body: centre of mass, world inertia, Newton-Euler force / moment; per dof: the
    //      free joint's axes; per site: world position
    if (lane < NB) {
      const int b = lane;
      double rho[3], t1[3], t2[3], t3[3];
      kin_mv(w.R[b], m.ipos[b], rho);
      for (int k = 0; k < 3; ++k) {
        w.com[b][k] = w.p[b][k] + rho[k];
        w.anchor[5 + b][k] = w.p[b][k];
      }
      // I_w = R I R'
      const double* I = m.inertia[b];
      const double If[9] = {I[0], I[3], I[4], I[3], I[1], I[5], I[4], I[5], I[2]};
      double T[9], Rt[9], Iw[9];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Rt[3 * i + j] = w.R[b][3 * j + i];
      kin_mm(w.R[b], If, T);
      kin_mm(T, Rt, Iw);
      w.Iw[b][0] = Iw[0]; w.Iw[b][1] = Iw[4]; w.Iw[b][2] = Iw[8];
      w.Iw[b][3] = Iw[1]; w.Iw[b][4] = Iw[2]; w.Iw[b][5] = Iw[5];
      kin_cross(w.al[b], rho, t1);
      kin_cross(w.w[b], rho, t2);
      kin_cross(w.w[b], t2, t3);
      double Ial[3], Iom[3];
      kin_sv(w.Iw[b], w.al[b], Ial);
      kin_sv(w.Iw[b], w.w[b], Iom);
      kin_cross(w.w[b], Iom, t2);
      for (int k = 0; k < 3; ++k) {
        w.F[b][k] = m.mass[b] * (w.a[b][k] + t1[k] + t3[k] - m.gravity[k]);
        w.N[b][k] = Ial[k] + t2[k];
      }
    }
    if (lane < 6) {
      const int d = lane;
      for (int k = 0; k < 3; ++k) {
        w.axis[d][k] = d < 3 ? (k == d ? 1.0 : 0.0) : w.R[0][3 * k + (d - 3)];
        w.anchor[d][k] = d < 3 ? 0.0 : w.p[0][k];
      }
    }
    for (int s = lane; s < NS; s += 32) {
      const int b = m.site_body[s];
      double r[3];
      kin_mv(w.R[b], m.site_pos[s], r);
      for (int k = 0; k < 3; ++k) w.site[s][k] = w.p[b][k] + r[k];
    }
    __syncwarp();
    // ---- centre-of-mass Jacobian columns and I_b axis_d per (body, dof)
    for (int e = lane; e < NB * NV; e += 32) {
      const int b = e / NV, d = e - b * NV;
      const bool on = (w.aff[b] >> d) & 1u;
      double jc[3] = {0, 0, 0}, ia[3] = {0, 0, 0};
      if (on) {
        if (d < 3) {
          jc[d] = 1.0;
        } else {
          const double r[3] = {w.com[b][0] - w.anchor[d][0], w.com[b][1] - w.anchor[d][1],
                               w.com[b][2] - w.anchor[d][2]};
          kin_cross(w.axis[d], r, jc);
          kin_sv(w.Iw[b], w.axis[d], ia);
        }
      }
      for (int k = 0; k < 3; ++k) {
        w.JC[b][d][k] = jc[k];
        w.IA[b][d][k] = ia[k];
      }
    }
    __syncwarp();
    // ---- outputs, entries in memory order (coalesced stores)
    double* Mo = Mout + (size_t)env * NV * NV;
    for (int e = lane; e < NV * NV; e += 32) {
      const int i = e / NV, j = e - i * NV;
      double acc = 0.0;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const double* ci = w.JC[b][i];
        const double* cj = w.JC[b][j];
        acc += m.mass[b] * (ci[0] * cj[0] + ci[1] * cj[1] + ci[2] * cj[2]);
        if (i >= 3) {  // IA is zero for dofs that do not turn body b
          const bool on = (w.aff[b] >> i) & 1u;
          const double* ia = w.IA[b][j];
          acc += on ? w.axis[i][0] * ia[0] + w.axis[i][1] * ia[1] + w.axis[i][2] * ia[2] : 0.0;
        }
      }
      Mo[e] = acc;
    }
    if (lane < NV) {
      const int d = lane;
      double acc = 0.0;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const double* c = w.JC[b][d];
        acc += c[0] * w.F[b][0] + c[1] * w.F[b][1] + c[2] * w.F[b][2];
        if (d >= 3 && ((w.aff[b] >> d) & 1u))
          acc += w.axis[d][0] * w.N[b][0] + w.axis[d][1] * w.N[b][1] + w.axis[d][2] * w.N[b][2];
      }
      Cout[(size_t)env * NV + d] = acc;
    }
    double* Jo = Jout + (size_t)env * 6 * NS * NV;
    for (int e = lane; e < 6 * NS * NV; e += 32) {
      const int row = e / NV, d = e - row * NV;
      const bool rot = row >= 3 * NS;
      const int rr = rot ? row - 3 * NS : row;
      const int s = rr / 3, k = rr - 3 * s;
      const int b = m.site_body[s];
      double val = 0.0;
      if ((w.aff[b] >> d) & 1u) {
        if (d < 3) {
          val = (!rot && k == d) ? 1.0 : 0.0;
        } else if (rot) {
          val = w.axis[d][k];
        } else {
          const int k1 = k == 2 ? 0 : k + 1, k2 = k == 0 ? 2 : k - 1;
          const double r1 = w.site[s][k1] - w.anchor[d][k1], r2 = w.site[s][k2] - w.anchor[d][k2];
          val = w.axis[d][k1] * r2 - w.axis[d][k2] * r1;  // (axis x r)_k
        }
      }
      Jo[e] = val;
    }
    double* bo = bias_out + (size_t)env * 6 * NS;
    for (int e = lane; e < 6 * NS; e += 32) {
      const bool rot = e >= 3 * NS;
      const int rr = rot ? e - 3 * NS : e;
      const int s = rr / 3, k = rr - 3 * s;
      const int b = m.site_body[s];
      double val;
      if (rot) {
        val = w.al[b][k];
      } else {
        const double rho[3] = {w.site[s][0] - w.p[b][0], w.site[s][1] - w.p[b][1], w.site[s][2] - w.p[b][2]};
        double t1[3], t2[3], t3[3];
        kin_cross(w.al[b], rho, t1);
        kin_cross(w.w[b], rho, t2);
        kin_cross(w.w[b], t2, t3);
        val = w.a[b][k] + t1[k] + t3[k];
      }
      bo[e] = val;
    }
  }
}

}  // namespace osc
