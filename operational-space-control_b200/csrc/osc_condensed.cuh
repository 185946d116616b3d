// osc_condensed.cuh -- the CONDENSED fast mode of the controller: one warp per environment.
//
// The reference keeps the dynamics  M dv + C = B u + Jc z  as equality rows of its QP
// (walter_sr/autogen/autogen.py:82-93, stacked at walter_sr/operational_space_controller.h:543-555)
// and never factors M.  Because dv is unbounded (dv_lb/ub = -/+inf, :286-287) it can be eliminated
// exactly -- the operational-space form BASELINE.json's north_star names:
//   M = L L'                       (Cholesky, one warp, factor in shared memory)
//   G = M^-1 [B  Jc]               (nv x n', n' = nu + 3 nc; one lane per column)
//   d0 = -M^-1 C ,  dv = G w + d0 ,  w = [u; z]
//   P' = G' Hd G + R ,  q' = G'(Hd d0 + fd)        Hd = 2 J'WJ + 2 w_reg I (build kernel)
//   min 1/2 w'P'w + q'w   s.t.  F z <= 0 (friction pyramid),  u, z boxes (contact mask)
// n' = 32 / 24 variables and m' = 64 / 40 rows instead of 46 / 42 and 92 / 76: the SAME unique
// optimum, a DIFFERENT ADMM iterate sequence.  So this is a separately reported mode
// (osc_step_condensed), never the path that is gated against the reference's OSQP iterates;
// its own oracle is oracle/osc_condensed.py (numpy condensation + the OSQP restatement's
// generic QP entry), against which it is held to the same iterate-level gates.
//
// Control-step protocol (mirrors the reference's re-Init branch, :571-584): every step is
// osqp_setup on the new condensed data (Ruiz scaling from scratch, rho carried over from the
// previous step), osqp_warm_start(previous w, previous dual), osqp_solve.
//
// Mapping: lane 4c + r owns friction row r of contact c and, for r < 3, force component r of
// that contact; the u variables sit in the r == 3 lanes (Walter: nu == nc) or after the contact
// lanes (Go2).  The lane of variable j holds ROW j of P' -- later of K^-1,
// K = P + sigma I + A' diag(rho) A (n' x n', dense, SPD) -- in registers: an ADMM iteration is
// ONE shared-memory exchange (r1) and n' DFMAs per lane against broadcast loads, the friction
// rows talk to their variables through 4-lane shuffles.
// Written against osc_warp.cuh like osc_core3.cuh: tests/host_core runs this source on the CPU.
#pragma once

#include "osc_core3.cuh"

namespace osc {

template <class D>
struct alignas(16) WorkspaceC {
  static constexpr int NV = D::NV, NU = D::NU, NC = D::NC, NZ = D::NZ, NF = D::NF;
  static constexpr int NP = NU + NZ;        // condensed variables
  static constexpr int MP = NF + NP;        // condensed rows: friction, then boxes
  static constexpr int STATE = NP + MP + 2; // w, dual (unscaled), rho, flag
  static_assert(STATE % 2 == 0 && NP % 2 == 0, "16-byte records");
  struct alignas(16) Stage {  // landing stage of the bulk copies
    // mass_matrix (overwritten by its Cholesky factor L, lower) followed by the contact rows
    // of J (= contact_jacobian'): both are dead once G is known, and the one array then serves
    // as the lane-private scratch of p_row (SCR doubles)
    double MJ[NV * NV + NZ * NV];
    double H[NV * NV];   // Hd (build kernel)
    double Cv[NV], fv[NV];
    double maskv[NC];
    double st[STATE];
  };
  Stage in;
  static constexpr int SCR = 32 * (NP / 2);
  static_assert(SCR <= NV * NV + NZ * NV, "p_row scratch fits over M and Jc");
  OSC_HD double* m() { return in.MJ; }
  OSC_HD const double* m() const { return in.MJ; }
  OSC_HD double* jc() { return in.MJ + NV * NV; }
  OSC_HD const double* jc() const { return in.MJ + NV * NV; }
  double Gt[NP * NV];      // G transposed: column j of G contiguous
  double d0[NV], vv[NV];   // -M^-1 C ;  Hd d0 + fd
  double linv[NV];         // 1 / L_ii
  double ds[2][NP], efs[2][32];  // Ruiz exchange, double buffered
  double r1[2][NP];        // iteration exchange, double buffered
  double xs[NP], av[NV], bv[NV];  // residual / output exchange
  double pub[2][2 * NP];   // Gauss-Jordan publish buffers (two pivot rows, double buffered)
  double Dv[NP], Eb[NP], Ef[32];
  double red[16];
};

template <class D>
struct CoreC {
  using WS = WorkspaceC<D>;
  using C3 = Core3<D>;
  static constexpr int NV = D::NV, NU = D::NU, NC = D::NC, NZ = D::NZ, NF = D::NF, NB = D::NB;
  static constexpr int NP = WS::NP, MP = WS::MP, STATE = WS::STATE;
  static_assert(NP <= 32 && NF <= 32, "one lane per condensed variable / friction row");

  // condensed (w-order) index of the lane's variable: u_k -> k, z_k -> NU + k; -1 if none
  static OSC_HD int wvar(int l) {
    const int u = C3::uk(l), z = C3::zk(l);
    return u >= 0 ? u : (z >= 0 ? NU + z : -1);
  }
  static OSC_HD Pair ld2(const double* p) { return C3::ld2(p); }
  static OSC_HD double pmax(double a, double v) { return v > a ? v : a; }

  struct Regs {
    Var<double> KI[NP];                 // row of P', then of K^-1
    Var<double> x, zu, yu, lu, uu, ru, riu, ibu, q;   // the lane's variable + its box row
    Var<double> zf, yf, rf, rif;        // the lane's friction row (upper bound 0)
    Var<double> fr[3], fc[4];           // its coefficients / the variable's column in the four rows
    Var<double> pd;                     // scaled diagonal contribution R_jj (for P x)
  };

  // ---- M = L L' (lower triangle, written back over M), 1 / L_ii on the side.  Lane i holds
  //      row i in registers; per column k the pivot comes by shuffle from lane k and the
  //      scaled column goes through a double-buffered shared-memory vector: one barrier per
  //      column, no read-modify-write of shared memory.
  static OSC_HD void cholesky(WS& w, const int lane0) {
    Var<double> a[NV];
    OSC_LANES(l) {
#pragma unroll
      for (int j = 0; j < NV; ++j) a[j][l] = l < NV ? w.m()[l * NV + j] : 0.0;
    }
    Warp::sync();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double s = C3::inv_sqrt(Warp::bcast(a[k], k));
      double* col = (k & 1) ? w.bv : w.av;
      OSC_LANES(l) {
        a[k][l] = a[k][l] * s;  // L_lk (lane k: sqrt(A_kk)); rows above k carry don't-cares
        if (l < NV) col[l] = a[k][l];
        if (l == k) w.linv[k] = s;
      }
      Warp::sync();
      OSC_LANES(l) {
#pragma unroll
        for (int j = k + 1; j < NV; ++j) a[j][l] -= a[k][l] * col[j];  // (entries j > l: unused)
      }
    }
    OSC_LANES(l) {
      if (l < NV) {
#pragma unroll
        for (int j = 0; j < NV; ++j)
          if (j <= l) w.m()[l * NV + j] = a[j][l];
      }
    }
    Warp::sync();
  }

  // x = M^-1 b for two right-hand sides per lane at once (the lane's column of [B Jc] and -C),
  // L read by broadcast loads
  static OSC_HD void solve2(const WS& w, double (&g)[NV], double (&d)[NV]) {
    const double* L = w.m();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double s = g[i], t = d[i];
#pragma unroll
      for (int j = 0; j < i; ++j) {
        const double lij = L[i * NV + j];
        s -= lij * g[j];
        t -= lij * d[j];
      }
      g[i] = s * w.linv[i];
      d[i] = t * w.linv[i];
      OSC_COMPILER_BARRIER();
    }
#pragma unroll
    for (int i = NV - 1; i >= 0; --i) {
      double s = g[i], t = d[i];
#pragma unroll
      for (int j = i + 1; j < NV; ++j) {
        const double lji = L[j * NV + i];
        s -= lji * g[j];
        t -= lji * d[j];
      }
      g[i] = s * w.linv[i];
      d[i] = t * w.linv[i];
      OSC_COMPILER_BARRIER();
    }
  }

  // Row j of P' = G' Hd G + R from G (shared memory) and Hd; q'_j on the side.  The loop over
  // the NP entries of the row is ROLLED (a register array cannot be indexed dynamically, so
  // the entries go through lane-private scratch slots, half a row at a time): unrolled it was
  // 1 k instructions of straight-line code per call, and this kernel is instruction-fetch
  // bound when its warps stream through long unrolled phases (ncu: stall_no_inst).
  static OSC_HD void p_row(WS& w, const Params& p, Regs& L, const int lane0, Var<double>* qout) {
    const double hu = 2.0 * (p.w_reg + p.w_torque), hz = 2.0 * p.w_reg;
    double* scr = w.m();
    constexpr int HALF = NP / 2;
    OSC_LANES(l) {
      const int j = wvar(l);
      const double* gj = &w.Gt[(j >= 0 ? j : 0) * NV];
      double t[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int k = 0; k < NV; k += 2) {
          const Pair h = ld2(&w.in.H[i * NV + k]), g = ld2(&gj[k]);
          a0 += h.x * g.x;
          a1 += h.y * g.y;
        }
        t[i] = a0 + a1;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll 1
        for (int ii = 0; ii < HALF; ++ii) {
          const int i = h * HALF + ii;
          double a0 = 0.0, a1 = 0.0;
#pragma unroll
          for (int k = 0; k < NV; k += 2) {
            const Pair g = ld2(&w.Gt[i * NV + k]);
            a0 += g.x * t[k];
            a1 += g.y * t[k + 1];
          }
          double v = a0 + a1;
          if (i == j) v += (j < NU ? hu : hz);
          scr[ii * 32 + l] = j >= 0 ? v : 0.0;
        }
#pragma unroll
        for (int ii = 0; ii < HALF; ++ii) L.KI[h * HALF + ii][l] = scr[ii * 32 + l];
      }
      if (qout) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int k = 0; k < NV; k += 2) {
          const Pair g = ld2(&gj[k]), v = ld2(&w.vv[k]);
          a0 += g.x * v.x;
          a1 += g.y * v.y;
        }
        (*qout)[l] = j >= 0 ? a0 + a1 : 0.0;
      }
    }
  }

  // ---- condensation: Cholesky, G, d0, v (P' rows / q' follow in p_row)
  static OSC_HD void condense(WS& w, const Params& p, const int lane0) {
    cholesky(w, lane0);
    OSC_LANES(l) {
      const int j = wvar(l);
      double g[NV], d[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        double b = 0.0;
        if (j >= 0 && j < NU) b = (i == NB + j) ? 1.0 : 0.0;          // B = [0; I] (autogen.py:54-60)
        if (j >= NU) b = w.jc()[(j - NU) * NV + i];                  // column of Jc (:497-503)
        g[i] = b;
        d[i] = -w.in.Cv[i];
      }
      solve2(w, g, d);
      if (j >= 0) {
#pragma unroll
        for (int i = 0; i < NV; i += 2) C3::st2(&w.Gt[j * NV + i], g[i], g[i + 1]);
      }
      if (l == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) w.d0[i] = d[i];
      }
      if (l < NV) {  // v = Hd d0 + fd
        double a = w.in.fv[l];
#pragma unroll
        for (int k = 0; k < NV; ++k) a += w.in.H[l * NV + k] * d[k];
        w.vv[l] = a;
      }
    }
    Warp::sync();
  }

  // ---- OSQP scale_data on [P' A''; A' 0] (cumulative form, like Core3::ruiz): D, E_box,
  //      E_friction, c.  P' rows stay unscaled in registers.
  static OSC_HD double ruiz(WS& w, const Params& p, const Regs& L, const Var<double>& qv,
                            Var<double>& Dj, Var<double>& Ebj, Var<double>& Efl, const int lane0) {
    OSC_LANES(l) {
      for (int b = 0; b < 2; ++b) {
        if (l < NP) w.ds[b][l] = 1.0;
        w.efs[b][l] = 1.0;
      }
      Dj[l] = Ebj[l] = Efl[l] = 1.0;
    }
    Warp::sync();
    Var<double> mH;
    auto sweep = [&](int b) {
      const double* ds = w.ds[b];
      OSC_LANES(l) {
        double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;
#pragma unroll
        for (int t = 0; t < NP; t += 4) {
          const Pair d0 = ld2(&ds[t]), d1 = ld2(&ds[t + 2]);
          m0 = pmax(m0, d0.x * fabs(L.KI[t][l]));
          m1 = pmax(m1, d0.y * fabs(L.KI[t + 1][l]));
          m2 = pmax(m2, d1.x * fabs(L.KI[t + 2][l]));
          m3 = pmax(m3, d1.y * fabs(L.KI[t + 3][l]));
        }
        mH[l] = pmax(pmax(m0, m1), pmax(m2, m3));
      }
    };
    static_assert(NP % 4 == 0, "sweep unrolls by four");
    sweep(0);
    double c = 1.0;
    for (int it = 0; it < p.scaling; ++it) {
      const int b = it & 1, nb = b ^ 1;
      const double* ds = w.ds[b];
      const double* efs = w.efs[b];
      double* dsn = w.ds[nb];
      double* efsn = w.efs[nb];
      OSC_LANES(l) {
        const int j = wvar(l);
        if (j >= 0) {
          const bool isz = j >= NU;
          const double dj = Dj[l];
          const double fm = !isz ? 0.0 : ((l & 3) < 2 ? 1.0 : p.mu);
          const Pair e01 = ld2(&efs[l & 28]), e23 = ld2(&efs[(l & 28) + 2]);
          const double bb = pmax(Ebj[l], pmax(pmax(e01.x, e01.y), pmax(e23.x, e23.y)) * fm);
          const double dt = C3::inv_sqrt(C3::limit_scaling(pmax((c * dj) * mH[l], dj * bb)));
          const double et = C3::inv_sqrt(C3::limit_scaling(Ebj[l] * dj));
          Dj[l] *= dt;
          Ebj[l] *= et;
          dsn[j] = Dj[l];
        }
        if (l < NF) {
          const double* dz = &ds[NU + 3 * (l >> 2)];
          const double e = pmax(pmax(dz[0], dz[1]), p.mu * dz[2]);
          Efl[l] *= C3::inv_sqrt(C3::limit_scaling(Efl[l] * e));
          efsn[l] = Efl[l];
        }
      }
      Warp::sync();
      sweep(nb);
      Var<double> sv, qq;
      OSC_LANES(l) {
        const bool has = wvar(l) >= 0;
        sv[l] = has ? (c * Dj[l]) * mH[l] : 0.0;
        qq[l] = has ? (c * Dj[l]) * fabs(qv[l]) : 0.0;
      }
      const double sum = Warp::sum(sv);
      const double qmax = Warp::max(qq);
      double ct = sum * (1.0 / (double)NP);
      ct = pmax(ct, C3::limit_scaling(qmax));
      ct = C3::limit_scaling(ct);
      c *= C3::rcp(ct);
    }
    OSC_LANES(l) {
      const int j = wvar(l);
      if (j >= 0) {
        w.Dv[j] = Dj[l];
        w.Eb[j] = Ebj[l];
      }
      w.Ef[l] = l < NF ? Efl[l] : 0.0;
    }
    Warp::sync();
    return c;
  }

  // ---- scaled bounds, linear cost, friction coefficients of the lane's rows / variable
  static OSC_HD void assemble(WS& w, const Params& p, Regs& L, const Var<double>& qv, double c,
                              const int lane0) {
    const double hu = 2.0 * (p.w_reg + p.w_torque), hz = 2.0 * p.w_reg;
    OSC_LANES(l) {
      const int j = wvar(l);
      L.ibu[l] = L.lu[l] = L.uu[l] = L.q[l] = L.pd[l] = 0.0;
#pragma unroll
      for (int r = 0; r < 4; ++r) L.fc[r][l] = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) L.fr[k][l] = 0.0;
      if (j >= 0) {
        const double dj = w.Dv[j], eb = w.Eb[j];
        L.ibu[l] = eb * dj;
        L.q[l] = (dj * qv[l]) * c;
        L.pd[l] = (c * dj) * dj * (j < NU ? hu : hz);
        double lo, hi;
        if (j < NU) {
          lo = p.u_lb[j];
          hi = p.u_ub[j];
        } else {
          // z bounds times the contact mask; OSQP_INFTY is finite so inf * 0 == 0 (:546-555)
          const int cc = l >> 2, kk = l & 3;
          const double mk = w.in.maskv[cc];
          lo = (kk < 2 ? -kInfty : 0.0) * mk;
          hi = (kk < 2 ? kInfty : p.fz_max) * mk;
          const double fm = kk < 2 ? 0.0 : -p.mu;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            double f = fm;
            if (kk == 0) f = (r & 1) ? -1.0 : 1.0;
            if (kk == 1) f = (r & 2) ? -1.0 : 1.0;
            L.fc[r][l] = (w.Ef[4 * cc + r] * f) * dj;  // pyramid rows (autogen.py:116-121)
          }
        }
        L.lu[l] = eb * fmax(lo, -kInfty);
        L.uu[l] = eb * fmin(hi, kInfty);
      }
      if (l < NF) {
        const int cc = l >> 2, r = l & 3;
        const double ef = w.Ef[l];
        const double f0 = (r & 1) ? -1.0 : 1.0, f1 = (r & 2) ? -1.0 : 1.0;
        L.fr[0][l] = (ef * f0) * w.Dv[NU + 3 * cc];
        L.fr[1][l] = (ef * f1) * w.Dv[NU + 3 * cc + 1];
        L.fr[2][l] = (ef * -p.mu) * w.Dv[NU + 3 * cc + 2];
      }
    }
  }

  static OSC_HD void set_rho(const WS& w, Regs& L, double rho, const int lane0) {
    OSC_LANES(l) {
      L.ru[l] = L.riu[l] = L.rf[l] = L.rif[l] = 0.0;
      if (wvar(l) >= 0) {
        L.ru[l] = C3::rho_of(L.lu[l], L.uu[l], rho);
        L.riu[l] = C3::rcp(L.ru[l]);
      }
      if (l < NF) {
        const double ef = w.Ef[l];
        L.rf[l] = C3::rho_of(ef * -kInfty, ef * 0.0, rho);
        L.rif[l] = C3::rcp(L.rf[l]);
      }
    }
  }

  // K = P + sigma I + A' diag(rho) A of the scaled problem from the UNSCALED P' row in L.KI,
  // then K^-1 by the two-pivot sweep (Core3::gj_sweep's algorithm on whole rows); the lane
  // ends up with row j of K^-1 in L.KI.
  static OSC_HD void factor(WS& w, const Params& p, Regs& L, double c, const int lane0) {
    // friction part of K seen from the lane's z variable (component kk of contact cc): entry
    // (kk, k2) of the contact's 3 x 3 block  sum_r rho_r A_r,kk A_r,k2
    Var<double> sk[3];
    {
      Var<double> rr[4], t;
#pragma unroll
      for (int r = 0; r < 4; ++r) Warp::group4(rr[r], L.rf, r);
#pragma unroll
      for (int k2 = 0; k2 < 3; ++k2) {
        OSC_LANES(l) { sk[k2][l] = 0.0; }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          Warp::group4(t, L.fc[r], k2);
          OSC_LANES(l) { sk[k2][l] += (rr[r][l] * L.fc[r][l]) * t[l]; }
        }
      }
    }
    OSC_LANES(l) {
      const int j = wvar(l);
      const double cdj = j >= 0 ? c * w.Dv[j] : 0.0;
#pragma unroll
      for (int t = 0; t < NP; t += 2) {
        const Pair d = ld2(&w.Dv[t]);
        L.KI[t][l] = (cdj * L.KI[t][l]) * d.x;
        L.KI[t + 1][l] = (cdj * L.KI[t + 1][l]) * d.y;
      }
      if (j >= 0) {
        const double dg = p.sigma + (L.ibu[l] * L.ibu[l]) * L.ru[l];
        const int cj = j >= NU ? (j - NU) / 3 : -1;
#pragma unroll
        for (int t = 0; t < NP; ++t) {
          double add = (t == j) ? dg : 0.0;
          if (t >= NU && (t - NU) / 3 == cj) add += sk[t >= NU ? (t - NU) % 3 : 0][l];
          L.KI[t][l] += add;
        }
      }
    }
    // two-pivot symmetric sweep (Core3::gj_sweep's algorithm on whole rows), the loop over the
    // pivot pairs ROLLED: instead of indexing the register row by the (dynamic) pivot, every
    // round rotates the columns left by two while it updates them, so that the pivots are
    // always columns 0 and 1 and land in columns NP-2, NP-1; after NP/2 rounds the order is
    // the original one again.  Register t holds column (t + 2 b) mod NP in round b.
#pragma unroll 1
    for (int b = 0; b < NP / 2; ++b) {
      const int off = 2 * b;
      double* buf = w.pub[b & 1];
      OSC_LANES(l) {
        const int j = wvar(l);
        if (j == off || j == off + 1) {
          double* d = buf + (j - off) * NP;
#pragma unroll
          for (int t = 0; t < NP; t += 2) C3::st2(d + t, L.KI[t][l], L.KI[t + 1][l]);
        }
      }
      Warp::sync();
      OSC_LANES(l) {
        const int j = wvar(l);
        const double* rp = buf;
        const double* rq = buf + NP;
        const Pair bp = ld2(rp);  // A_pp, A_pq
        const double aqq = rq[1];
        const double dinv = C3::rcp(bp.x * aqq - bp.y * bp.y);
        const double b11 = aqq * dinv, b12 = -(bp.y * dinv), b22 = bp.x * dinv;
        // A_ip == A_pi, A_iq == A_qi up to rounding: taken from the published rows at the
        // lane's own column (position j - off in the rotated order)
        int pos = j - off;
        pos = pos < 0 ? pos + NP : pos;
        const double cp = j >= 0 ? rp[pos] : 0.0, cq = j >= 0 ? rq[pos] : 0.0;
        double g1 = cp * b11 + cq * b12, g2 = cp * b12 + cq * b22;
        if (j == off) {
          g1 = -b11;
          g2 = -b12;
        }
        if (j == off + 1) {
          g1 = -b12;
          g2 = -b22;
        }
        const double keep = (j == off || j == off + 1) ? 0.0 : 1.0;
#pragma unroll
        for (int t = 0; t < NP - 2; t += 2) {
          const Pair r1 = ld2(rp + t + 2), r2 = ld2(rq + t + 2);
          L.KI[t][l] = (L.KI[t + 2][l] * keep - g1 * r1.x) - g2 * r2.x;
          L.KI[t + 1][l] = (L.KI[t + 3][l] * keep - g1 * r1.y) - g2 * r2.y;
        }
        L.KI[NP - 2][l] = g1;
        L.KI[NP - 1][l] = g2;
      }
    }
    OSC_LANES(l) {
      const bool has = wvar(l) >= 0;
#pragma unroll
      for (int t = 0; t < NP; ++t) L.KI[t][l] = has ? -L.KI[t][l] : 0.0;
    }
    Warp::sync();
  }

  // ---- one ADMM iteration (osqp.c: update_xz_tilde, update_x, update_z, update_y)
  static OSC_HD void iterate(WS& w, const Params& p, Regs& L, double* r1s, const int lane0) {
    Var<double> wf, w0, w1, w2, w3, xt;
    OSC_LANES(l) { wf[l] = L.rf[l] * L.zf[l] - L.yf[l]; }
    Warp::group4(w0, wf, 0);
    Warp::group4(w1, wf, 1);
    Warp::group4(w2, wf, 2);
    Warp::group4(w3, wf, 3);
    OSC_LANES(l) {
      double v = (p.sigma * L.x[l] - L.q[l]) + L.ibu[l] * (L.ru[l] * L.zu[l] - L.yu[l]);
      v += (L.fc[0][l] * w0[l] + L.fc[1][l] * w1[l]) + (L.fc[2][l] * w2[l] + L.fc[3][l] * w3[l]);
      const int j = wvar(l);
      if (j >= 0) r1s[j] = v;
    }
    Warp::sync();
    OSC_LANES(l) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int t = 0; t < NP; t += 4) {
        const Pair u0 = ld2(&r1s[t]), u1 = ld2(&r1s[t + 2]);
        a0 += L.KI[t][l] * u0.x;
        a1 += L.KI[t + 1][l] * u0.y;
        a2 += L.KI[t + 2][l] * u1.x;
        a3 += L.KI[t + 3][l] * u1.y;
      }
      xt[l] = (a0 + a1) + (a2 + a3);
    }
    Var<double> x0, x1, x2;
    Warp::group4(x0, xt, 0);
    Warp::group4(x1, xt, 1);
    Warp::group4(x2, xt, 2);
    const double al = p.alpha, be = 1.0 - p.alpha;
    OSC_LANES(l) {
      {
        const double zr = al * (L.ibu[l] * xt[l]) + be * L.zu[l];
        const double zn = C3::clip(zr + L.riu[l] * L.yu[l], L.lu[l], L.uu[l]);
        L.yu[l] += L.ru[l] * (zr - zn);
        L.zu[l] = zn;
        L.x[l] = al * xt[l] + be * L.x[l];
      }
      {
        const double zt = L.fr[0][l] * x0[l] + L.fr[1][l] * x1[l] + L.fr[2][l] * x2[l];
        const double zr = al * zt + be * L.zf[l];
        double zn = zr + L.rif[l] * L.yf[l];
        zn = zn > 0.0 ? 0.0 : zn;
        L.yf[l] += L.rf[l] * (zr - zn);
        L.zf[l] = zn;
      }
    }
  }

  struct Residuals {
    double pri_res, dua_res, eps_pri_norm, eps_dua_norm, rho_pri, rho_dua;
  };

  // P x of the scaled problem without a stored P: P = c D (G' Hd G + R) D
  static OSC_HD void px(WS& w, const Regs& L, double c, const int lane0, Var<double>& out) {
    Warp::sync();
    OSC_LANES(l) {
      const int j = wvar(l);
      if (j >= 0) w.xs[j] = w.Dv[j] * L.x[l];
    }
    Warp::sync();
    OSC_LANES(l) {
      if (l < NV) {  // a = G (D x)
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int t = 0; t < NP; t += 2) {
          a0 += w.Gt[t * NV + l] * w.xs[t];
          a1 += w.Gt[(t + 1) * NV + l] * w.xs[t + 1];
        }
        w.av[l] = a0 + a1;
      }
    }
    Warp::sync();
    OSC_LANES(l) {
      if (l < NV) {  // b = Hd a
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int k = 0; k < NV; k += 2) {
          const Pair h = ld2(&w.in.H[l * NV + k]), v = ld2(&w.av[k]);
          a0 += h.x * v.x;
          a1 += h.y * v.y;
        }
        w.bv[l] = a0 + a1;
      }
    }
    Warp::sync();
    OSC_LANES(l) {
      const int j = wvar(l);
      double v = 0.0;
      if (j >= 0) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int k = 0; k < NV; k += 2) {
          const Pair g = ld2(&w.Gt[j * NV + k]), b = ld2(&w.bv[k]);
          a0 += g.x * b.x;
          a1 += g.y * b.y;
        }
        v = (c * w.Dv[j]) * (a0 + a1) + L.pd[l] * L.x[l];
      }
      out[l] = v;
    }
  }

  static OSC_HD Residuals residuals(WS& w, const Regs& L, double c, const int lane0) {
    Var<double> pxv;
    px(w, L, c, lane0, pxv);
    Var<double> y0, y1, y2, y3, x0, x1, x2;
    Warp::group4(y0, L.yf, 0);
    Warp::group4(y1, L.yf, 1);
    Warp::group4(y2, L.yf, 2);
    Warp::group4(y3, L.yf, 3);
    Warp::group4(x0, L.x, 0);
    Warp::group4(x1, L.x, 1);
    Warp::group4(x2, L.x, 2);
    Var<double> m[8];
    OSC_LANES(l) {
      double pr_u = 0, pr_s = 0, np_u = 0, np_s = 0, du_u = 0, du_s = 0, nd_u = 0, nd_s = 0;
      auto prim = [&](double axv, double zi, double ei) {
        const double d = axv - zi;
        pr_s = pmax(pr_s, fabs(d));
        pr_u = pmax(pr_u, fabs(ei * d));
        np_s = pmax(pmax(np_s, fabs(zi)), fabs(axv));
        np_u = pmax(pmax(np_u, fabs(ei * zi)), fabs(ei * axv));
      };
      const int j = wvar(l);
      if (j >= 0) {
        prim(L.ibu[l] * L.x[l], L.zu[l], C3::rcp(w.Eb[j]));
        const double aty = ((L.fc[0][l] * y0[l] + L.fc[1][l] * y1[l]) +
                            (L.fc[2][l] * y2[l] + L.fc[3][l] * y3[l])) + L.ibu[l] * L.yu[l];
        const double di = C3::rcp(w.Dv[j]);
        const double d = L.q[l] + pxv[l] + aty;
        du_s = fabs(d);
        du_u = fabs(di * d);
        nd_s = pmax(pmax(fabs(L.q[l]), fabs(pxv[l])), fabs(aty));
        nd_u = pmax(pmax(fabs(di * L.q[l]), fabs(di * pxv[l])), fabs(di * aty));
      }
      if (l < NF) {
        const double axv = L.fr[0][l] * x0[l] + L.fr[1][l] * x1[l] + L.fr[2][l] * x2[l];
        prim(axv, L.zf[l], C3::rcp(w.Ef[l]));
      }
      m[0][l] = pr_u; m[1][l] = np_u; m[2][l] = du_u; m[3][l] = nd_u;
      m[4][l] = pr_s; m[5][l] = np_s; m[6][l] = du_s; m[7][l] = nd_s;
    }
    double r8[8];
    Warp::maxn<8>(m, r8, w.red, lane0);
    Warp::sync();
    const double cinv = 1.0 / c;
    Residuals r;
    r.pri_res = r8[0];
    r.dua_res = cinv * r8[2];
    r.eps_pri_norm = r8[1];
    r.eps_dua_norm = cinv * r8[3];
    r.rho_pri = r8[4] / (r8[5] + 1e-10);
    r.rho_dua = r8[6] / (r8[7] + 1e-10);
    return r;
  }

  // osqp_solve: same loop structure as Core3::admm (stretches that end with an event
  // iteration).  The condensed QP is always feasible and strictly convex, so the
  // infeasibility certificates of check_termination can never fire and are not evaluated.
  // Resumable: returns false when rho changed and the caller has to rebuild K^-1 (p_row +
  // factor have ONE call site each, in step(): inlined twice they were 4 k instructions of
  // a kernel whose warps run through different phases of it at the same time).
  struct Loop {
    int iter, to_check, to_adapt, interval;
    bool adaptive;
    double rho;
    Result res;
    Residuals r;
  };
  static OSC_HD void admm_begin(const Params& p, Loop& s, double rho) {
    s.res.iter = 0;
    s.res.status = kUnsolved;
    s.res.rho_updates = 0;
    s.res.reinit = 0;
    s.interval = p.adaptive_rho_interval;
    if (p.adaptive_rho && !s.interval)
      s.interval = p.check_termination ? 4 * p.check_termination : 100;
    s.adaptive = p.adaptive_rho && s.interval;
    s.r.pri_res = s.r.dua_res = s.r.eps_pri_norm = s.r.eps_dua_norm = s.r.rho_pri = s.r.rho_dua = 0.0;
    s.iter = 0;
    s.to_check = p.check_termination;
    s.to_adapt = s.interval;
    s.rho = rho;
  }
  static OSC_HD bool admm_run(WS& w, const Params& p, Regs& L, double c, Loop& s, const int lane0) {
    for (;;) {
      int n = p.max_iter - s.iter;
      if (p.check_termination && s.to_check < n) n = s.to_check;
      if (s.adaptive && s.to_adapt < n) n = s.to_adapt;
      OSC_TICK(27);
#pragma unroll 1
      for (int k = 0; k < n; ++k) iterate(w, p, L, w.r1[(s.iter + k) & 1], lane0);
      OSC_TICK(28);
      s.iter += n;
      s.to_check -= n;
      s.to_adapt -= n;
      const bool last = s.iter >= p.max_iter;
      const bool check = p.check_termination && s.to_check == 0;
      const bool adapt = s.adaptive && s.to_adapt == 0;
      if (check) s.to_check = p.check_termination;
      if (adapt) s.to_adapt = s.interval;
      s.r = residuals(w, L, c, lane0);
      bool ended_at_check = false;
      if (check || last) {
        if (s.r.pri_res > kInfty || s.r.dua_res > kInfty) {
          s.res.status = kNonCvx;
        } else {
          const bool ok = s.r.pri_res < p.eps_abs + p.eps_rel * s.r.eps_pri_norm &&
                          s.r.dua_res < p.eps_abs + p.eps_rel * s.r.eps_dua_norm;
          if (ok) s.res.status = kSolved;
          else if (last && s.r.pri_res < 10 * (p.eps_abs + p.eps_rel * s.r.eps_pri_norm) &&
                   s.r.dua_res < 10 * (p.eps_abs + p.eps_rel * s.r.eps_dua_norm))
            s.res.status = kSolvedInaccurate;
        }
        ended_at_check = check && s.res.status != kUnsolved;
      }
      bool refactor = false;
      if (adapt && !ended_at_check) {
        double rho_new = s.rho * sqrt(s.r.rho_pri / (s.r.rho_dua + 1e-10));
        rho_new = fmin(fmax(rho_new, kRhoMin), kRhoMax);
        if (rho_new > s.rho * p.rho_tol || rho_new < s.rho / p.rho_tol) {
          s.rho = rho_new;
          s.res.rho_updates++;
          refactor = !last;
        }
      }
      if (s.res.status != kUnsolved) return true;
      if (last) {
        s.res.status = kMaxIterReached;
        return true;
      }
      Warp::sync();
      if (refactor) return false;
    }
  }

  // One control step of one environment.  w.in holds the landed record; out_x [N] gets the
  // reference-shaped solution [dv; u; z], out_y [M] the dual (friction and box rows; the
  // eliminated dynamics rows and the free dv rows report 0), torque [NU], state_out [STATE].
  static OSC_HD Result step(WS& w, const Params& p, const int lane0, double* out_x, double* out_y,
                            double* torque, double* state_out) {
    Regs L;
    Var<double> qv, Dj, Ebj, Efl;
    OSC_TICK(17);
    condense(w, p, lane0);
    OSC_TICK(18);
    double c = 1.0;
    Loop lp;
    bool first = true;
    for (;;) {
      // P' row of the lane (the ONE call site: first pass, and again after every rho change,
      // when K^-1 has overwritten it)
      OSC_TICK(19);
      p_row(w, p, L, lane0, first ? &qv : nullptr);
      OSC_TICK(20);
      if (first) {
        first = false;
        OSC_TICK(21);
        if (p.scaling > 0) {
          c = ruiz(w, p, L, qv, Dj, Ebj, Efl, lane0);
        } else {
          OSC_LANES(l) {
            const int j = wvar(l);
            if (j >= 0) w.Dv[j] = w.Eb[j] = 1.0;
            w.Ef[l] = l < NF ? 1.0 : 0.0;
          }
          Warp::sync();
        }
        assemble(w, p, L, qv, c, lane0);
        const bool have = w.in.st[NP + MP + 1] != 0.0;
        double rho = have ? w.in.st[NP + MP] : p.rho0;
        rho = fmin(fmax(rho, kRhoMin), kRhoMax);
        // osqp_warm_start(x, y): x <- Dinv x, y <- c Einv y, z <- A x  (cold: zeros)
        const bool warm = have && p.warm_start;
        OSC_LANES(l) {
          const int j = wvar(l);
          L.x[l] = L.yu[l] = L.yf[l] = 0.0;
          if (warm && j >= 0) {
            L.x[l] = C3::rcp(w.Dv[j]) * w.in.st[j];
            L.yu[l] = (C3::rcp(w.Eb[j]) * w.in.st[NP + NF + j]) * c;
          }
          if (warm && l < NF) L.yf[l] = (C3::rcp(w.Ef[l]) * w.in.st[NP + l]) * c;
        }
        Var<double> x0, x1, x2;
        Warp::group4(x0, L.x, 0);
        Warp::group4(x1, L.x, 1);
        Warp::group4(x2, L.x, 2);
        OSC_LANES(l) {
          L.zu[l] = L.ibu[l] * L.x[l];
          L.zf[l] = L.fr[0][l] * x0[l] + L.fr[1][l] * x1[l] + L.fr[2][l] * x2[l];
        }
        admm_begin(p, lp, rho);
        OSC_TICK(22);
      }
      OSC_TICK(23);
      set_rho(w, L, lp.rho, lane0);
      factor(w, p, L, c, lane0);
      OSC_TICK(24);
      const bool done = admm_run(w, p, L, c, lp, lane0);
      OSC_TICK(25);
      if (done) break;
    }
    Result res = lp.res;
    res.iter = lp.iter;
    res.pri_res = lp.r.pri_res;
    res.dua_res = lp.r.dua_res;
    res.rho = lp.rho;
    // ---- un-scale, recover dv = G w + d0, outputs
    const double cinv = 1.0 / c;
    Warp::sync();
    OSC_LANES(l) {
      const int j = wvar(l);
      if (j >= 0) {
        const double wj = w.Dv[j] * L.x[l];
        const double yb = (w.Eb[j] * L.yu[l]) * cinv;
        w.xs[j] = wj;
        out_x[NV + j] = wj;
        if (j < NU) torque[j] = wj;  // torque_command = solution[nv : nv+nu] (:631)
        out_y[D::RB + NV + j] = yb;
        state_out[j] = wj;
        state_out[NP + NF + j] = yb;
      }
      if (l < NF) {
        const double yfv = (w.Ef[l] * L.yf[l]) * cinv;
        out_y[D::RF + l] = yfv;
        state_out[NP + l] = yfv;
      }
      if (l < NV) {
        out_y[l] = 0.0;
        out_y[D::RB + l] = 0.0;
      }
      if (l == 0) {
        state_out[NP + MP] = res.rho;
        state_out[NP + MP + 1] = 1.0;
      }
    }
    Warp::sync();
    OSC_LANES(l) {
      if (l < NV) {
        double a0 = w.d0[l], a1 = 0.0;
#pragma unroll
        for (int t = 0; t < NP; t += 2) {
          a0 += w.Gt[t * NV + l] * w.xs[t];
          a1 += w.Gt[(t + 1) * NV + l] * w.xs[t + 1];
        }
        out_x[l] = a0 + a1;
      }
    }
    Warp::sync();
    OSC_TICK(26);
    return res;
  }
};

}  // namespace osc
