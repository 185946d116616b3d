// tests/host_core/host_core.cpp -- TEST HARNESS ONLY.
// Instantiates the product's per-environment algorithm (csrc/osc_core3.cuh) on the host with
// an emulated 32-lane warp (csrc/osc_warp.cuh) so that `pytest -m "not gpu"` can compare the
// exact code the GPU runs against the oracle.  Never linked into libosc_b200.so.
#include <cstring>
#include <memory>
#include <type_traits>

#include "../../operational-space-control_b200/csrc/osc_params.h"
#include "../../operational-space-control_b200/csrc/osc_condensed.cuh"

namespace {
// added to the linear cost after the objective build: lets a test hand the solver a cost the
// controller API cannot produce (f outside the range of H -> dual infeasible QP)
double g_f_offset[32] = {0};

template <class D>
int run(const osc::Params& p, const double* M, const double* C, const double* J,
        const double* bias, const double* targets, const double* mask, double* state, double* x,
        double* y, double* torque, int* info_i, double* info_d, double* Hdv_out, double* f_out) {
  using Core = osc::Core3<D>;
  using WSpace = osc::Workspace3<D>;
  using B = osc::BuildQP<D>;
  auto ws = std::make_unique<WSpace>();
  std::memset(ws.get(), 0, sizeof(*ws));
  // objective build (the build kernel's per-item functions)
  double H[D::NV * D::NV], f[D::NV];
  for (int a = 0; a < D::NV; ++a) {
    for (int b = 0; b <= a; ++b) {
      const double v = B::h_entry(J, p.w_row, p.w_reg, a, b);
      H[a * D::NV + b] = v;
      H[b * D::NV + a] = v;
    }
    f[a] = B::f_entry(J, bias, targets, p.w_row, a) + g_f_offset[a];
  }
  if (Hdv_out) std::memcpy(Hdv_out, H, sizeof(H));
  if (f_out) std::memcpy(f_out, f, sizeof(f));
  if (!state) return 0;
  double *dM = ws->in.M, *dH = ws->in.H, *dJc = ws->in.Jc, *dLand = ws->in.land;
  double *dC = ws->in.Cv, *dF = ws->in.fv, *dMask = ws->in.maskv;
  std::memcpy(dM, M, sizeof(double) * D::NV * D::NV);
  std::memcpy(dH, H, sizeof(H));
  std::memcpy(dJc, J + D::JC0 * D::NV, sizeof(double) * D::NZ * D::NV);
  if (!x) {
    // osc_setup (init_state_kernel): cold iterates, previous linear cost = f, rho0,
    // initialised flag, sparsity signature of the set-up data
    std::memset(state, 0, sizeof(double) * D::STATE);
    std::memcpy(state + D::N + 2 * D::M, f, sizeof(f));
    state[D::N + 2 * D::M + D::NV] = p.rho0;
    state[D::N + 2 * D::M + D::NV + 1] = 1.0;
    for (int q = 0; q < D::SIG; ++q)
      state[D::SIG0 + q] = Core::as_f64(Core::sig_word(dH, dM, dJc, q, 0));
    return 0;
  }
  std::memcpy(dLand, state, sizeof(double) * D::STATE);
  std::memcpy(dC, C, sizeof(double) * D::NV);
  std::memcpy(dF, f, sizeof(f));
  std::memcpy(dMask, mask, sizeof(double) * D::NC);
  {
    // the equilibration kernel's part (scale_kernel3): Ruiz passes + path decision
    auto rw = std::make_unique<osc::RuizWorkspace<D>>();
    std::memset(rw.get(), 0, sizeof(*rw));
    std::memcpy(rw->in.M, M, sizeof(rw->in.M));
    std::memcpy(rw->in.H, H, sizeof(rw->in.H));
    std::memcpy(rw->in.Jc, J + D::JC0 * D::NV, sizeof(rw->in.Jc));
    std::memcpy(rw->in.tail, state + D::N + 2 * D::M, sizeof(rw->in.tail));
    std::memcpy(rw->in.fv, f, sizeof(rw->in.fv));
    Core::ruiz(*rw, p, 0, ws->in.scal, state + D::SIG0, [] {});
    // (the solve kernel lands the state record after the equilibration kernel ran)
    std::memcpy(dLand, state, sizeof(double) * D::STATE);
  }
  osc::Result r = Core::step(*ws, p, 0, x, y, torque, state);
  info_i[0] = r.iter;
  info_i[1] = r.status;
  info_i[2] = r.rho_updates;
  info_i[3] = r.reinit;
  info_d[0] = r.pri_res;
  info_d[1] = r.dua_res;
  info_d[2] = r.rho;
  return 0;
}
// the condensed fast mode (csrc/osc_condensed.cuh): one control step of one environment;
// `state` is the WorkspaceC::STATE-double record (all zeros = cold start), updated in place
template <class D>
int run_condensed(const osc::Params& p, const double* M, const double* C, const double* J,
                  const double* bias, const double* targets, const double* mask, double* state,
                  double* x, double* y, double* torque, int* info_i, double* info_d) {
  using Core = osc::CoreC<D>;
  using WSpace = osc::WorkspaceC<D>;
  using B = osc::BuildQP<D>;
  auto ws = std::make_unique<WSpace>();
  std::memset(ws.get(), 0, sizeof(*ws));
  for (int a = 0; a < D::NV; ++a) {
    for (int b = 0; b <= a; ++b) {
      const double v = B::h_entry(J, p.w_row, p.w_reg, a, b);
      ws->in.H[a * D::NV + b] = v;
      ws->in.H[b * D::NV + a] = v;
    }
    ws->in.fv[a] = B::f_entry(J, bias, targets, p.w_row, a);
  }
  std::memcpy(ws->m(), M, sizeof(double) * D::NV * D::NV);
  std::memcpy(ws->jc(), J + D::JC0 * D::NV, sizeof(double) * D::NZ * D::NV);
  std::memcpy(ws->in.Cv, C, sizeof(ws->in.Cv));
  std::memcpy(ws->in.maskv, mask, sizeof(ws->in.maskv));
  std::memcpy(ws->in.st, state, sizeof(ws->in.st));
  osc::Result r = Core::step(*ws, p, 0, x, y, torque, state);
  info_i[0] = r.iter;
  info_i[1] = r.status;
  info_i[2] = r.rho_updates;
  info_i[3] = 0;
  info_d[0] = r.pri_res;
  info_d[1] = r.dua_res;
  info_d[2] = r.rho;
  return 0;
}
}  // namespace

extern "C" int osc_condensed_host_state_size(const osc_robot_spec* spec) {
  switch (osc::shape_of(*spec)) {
    case osc::Shape::kWalter: return osc::WorkspaceC<osc::WalterDims>::STATE;
    case osc::Shape::kGo2: return osc::WorkspaceC<osc::Go2Dims>::STATE;
    default: return -1;
  }
}

extern "C" int osc_condensed_host_step(const osc_robot_spec* spec, const osc_settings* settings,
                                       const double* M, const double* C, const double* J,
                                       const double* bias, const double* targets,
                                       const double* mask, double* state, double* x, double* y,
                                       double* torque, int* info_i, double* info_d) {
  const osc::Params p = osc::make_params(*spec, *settings);
  switch (osc::shape_of(*spec)) {
    case osc::Shape::kWalter:
      return run_condensed<osc::WalterDims>(p, M, C, J, bias, targets, mask, state, x, y, torque,
                                            info_i, info_d);
    case osc::Shape::kGo2:
      return run_condensed<osc::Go2Dims>(p, M, C, J, bias, targets, mask, state, x, y, torque,
                                         info_i, info_d);
    default:
      return -1;
  }
}

extern "C" int osc_core_host_state_size(const osc_robot_spec* spec) {
  switch (osc::shape_of(*spec)) {
    case osc::Shape::kWalter: return osc::WalterDims::STATE;
    case osc::Shape::kGo2: return osc::Go2Dims::STATE;
    default: return -1;
  }
}

extern "C" void osc_core_host_set_f_offset(const double* df, int n) {
  for (int a = 0; a < 32; ++a) g_f_offset[a] = (df && a < n) ? df[a] : 0.0;
}

// state == NULL: only build H (nv*nv) and f (nv).
// state != NULL, x == NULL: osc_setup -- fill the STATE-double record for this data.
// otherwise: one control step; state is updated in place; x / y must hold the previous
// step's solution on entry (read on the sparsity-change path) and receive the new one.
extern "C" int osc_core_host_step(const osc_robot_spec* spec, const osc_settings* settings,
                                  const double* M, const double* C, const double* J,
                                  const double* bias, const double* targets, const double* mask,
                                  double* state, double* x, double* y, double* torque,
                                  int* info_i, double* info_d, double* Hdv_out, double* f_out) {
  const osc::Params p = osc::make_params(*spec, *settings);
  switch (osc::shape_of(*spec)) {
    case osc::Shape::kWalter:
      return run<osc::WalterDims>(p, M, C, J, bias, targets, mask, state, x, y, torque, info_i,
                                  info_d, Hdv_out, f_out);
    case osc::Shape::kGo2:
      return run<osc::Go2Dims>(p, M, C, J, bias, targets, mask, state, x, y, torque, info_i,
                               info_d, Hdv_out, f_out);
    default:
      return -1;
  }
}
