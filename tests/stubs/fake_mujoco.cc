// tests/stubs/fake_mujoco.cc -- a SCRIPTED stand-in for the MuJoCo calls of
// update_mj_data / update_osc_data (reference walter_sr/operational_space_controller.h:394-513),
// so that the OSC_B200_HAVE_MUJOCO branch of the drop-in classes is compiled AND run by the
// tests.  It is not physics: "loading a model" reads a binary blob
//   int32 nv, nu, ns, nc ; double M[nv*nv], C[nv], J[6 ns nv], bias[6 ns]
// and the "dynamics" hand those numbers back through MuJoCo's interfaces in the way the real
// library would deliver them:
//   mj_fullM      -> M                      qfrc_bias -> C
//   site_xpos[id] -> (1000 id, 0, 0)        so that mj_jac / mj_jacDot can tell which site a
//                                           `point` belongs to (the controller passes points,
//                                           not ids, :459-482)
//   mj_jac        -> rows 3i..3i+2 of Jp, Jr for the site of `point`
//   mj_jacDot     -> Jdot with bias/qvel[0] in column 0, so that Jdot * qvel == bias when the
//                    test sets linear_body_velocity = (1, 0, 0) and everything else to zero
// name -> id: ids are handed out in order of first request per object type, i.e. in the
// order of the controller's site_list / body_list.
// The last qpos / qvel seen by mj_fwdPosition are kept for the test to inspect
// (fake_mj_last_qpos / fake_mj_last_qvel).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "mujoco/mujoco.h"

namespace {
struct FakeModel {
  int nv = 0, nu = 0, ns = 0, nc = 0;
  std::vector<double> M, C, J, bias;
  std::map<std::string, int> ids[8];
};
struct FakeData {
  std::vector<double> qpos, qvel, ctrl, qfrc_actuator, qfrc_bias, qM, site_xpos;
};
std::vector<double> g_last_qpos, g_last_qvel;
int g_fwd_calls = 0;
}  // namespace

extern "C" {

const double* fake_mj_last_qpos(int* n) { if (n) *n = (int)g_last_qpos.size(); return g_last_qpos.data(); }
const double* fake_mj_last_qvel(int* n) { if (n) *n = (int)g_last_qvel.size(); return g_last_qvel.data(); }
int fake_mj_forward_calls() { return g_fwd_calls; }

mjModel* mj_loadXML(const char* filename, const void*, char* error, int error_sz) {
  FILE* f = std::fopen(filename, "rb");
  if (!f) {
    if (error) std::snprintf(error, (size_t)error_sz, "fake_mujoco: cannot open %s", filename);
    return nullptr;
  }
  auto* fm = new FakeModel();
  int32_t hdr[4];
  bool ok = std::fread(hdr, sizeof(int32_t), 4, f) == 4;
  if (ok) {
    fm->nv = hdr[0]; fm->nu = hdr[1]; fm->ns = hdr[2]; fm->nc = hdr[3];
    const size_t nv = fm->nv, s = 6 * (size_t)fm->ns;
    fm->M.resize(nv * nv); fm->C.resize(nv); fm->J.resize(s * nv); fm->bias.resize(s);
    ok = std::fread(fm->M.data(), 8, fm->M.size(), f) == fm->M.size() &&
         std::fread(fm->C.data(), 8, fm->C.size(), f) == fm->C.size() &&
         std::fread(fm->J.data(), 8, fm->J.size(), f) == fm->J.size() &&
         std::fread(fm->bias.data(), 8, fm->bias.size(), f) == fm->bias.size();
  }
  std::fclose(f);
  if (!ok) {
    if (error) std::snprintf(error, (size_t)error_sz, "fake_mujoco: short blob %s", filename);
    delete fm;
    return nullptr;
  }
  auto* m = new mjModel();
  std::memset(m, 0, sizeof(*m));
  m->nv = fm->nv; m->nq = fm->nv + 1; m->nu = fm->nu; m->nsite = fm->ns; m->nbody = fm->ns + 1;
  m->opt.timestep = 0.001;
  m->fake = fm;
  return m;
}

mjData* mj_makeData(const mjModel* m) {
  auto* fm = static_cast<FakeModel*>(m->fake);
  auto* fd = new FakeData();
  fd->qpos.assign(m->nq, 0.0); fd->qvel.assign(m->nv, 0.0); fd->ctrl.assign(m->nu, 0.0);
  fd->qfrc_actuator.assign(m->nv, 0.0); fd->qfrc_bias = fm->C;
  fd->qM = fm->M;  // (the real qM is a sparse packing; only mj_fullM reads it)
  fd->site_xpos.assign(3 * (size_t)(m->nsite + 8), 0.0);
  auto* d = new mjData();
  std::memset(d, 0, sizeof(*d));
  d->qpos = fd->qpos.data(); d->qvel = fd->qvel.data(); d->ctrl = fd->ctrl.data();
  d->qfrc_actuator = fd->qfrc_actuator.data(); d->qfrc_bias = fd->qfrc_bias.data();
  d->qM = fd->qM.data(); d->site_xpos = fd->site_xpos.data();
  d->fake = fd;
  return d;
}

void mj_deleteData(mjData* d) {
  if (!d) return;
  delete static_cast<FakeData*>(d->fake);
  delete d;
}
void mj_deleteModel(mjModel* m) {
  if (!m) return;
  delete static_cast<FakeModel*>(m->fake);
  delete m;
}

int mj_name2id(const mjModel* m, int type, const char* name) {
  auto* fm = static_cast<FakeModel*>(m->fake);
  auto& tab = fm->ids[type & 7];
  auto it = tab.find(name);
  if (it != tab.end()) return it->second;
  const int id = (int)tab.size();
  tab[name] = id;
  return id;
}

void mj_fwdPosition(const mjModel* m, mjData* d) {
  g_last_qpos.assign(d->qpos, d->qpos + m->nq);
  g_last_qvel.assign(d->qvel, d->qvel + m->nv);
  ++g_fwd_calls;
  for (int i = 0; i < m->nsite; ++i) {
    d->site_xpos[3 * i + 0] = 1000.0 * i;
    d->site_xpos[3 * i + 1] = 0.0;
    d->site_xpos[3 * i + 2] = 0.0;
  }
}
void mj_fwdVelocity(const mjModel*, mjData*) {}

void mj_fullM(const mjModel* m, mjtNum* dst, const mjtNum*) {
  auto* fm = static_cast<FakeModel*>(m->fake);
  std::memcpy(dst, fm->M.data(), sizeof(double) * fm->M.size());
}

static int site_of_point(const mjtNum point[3]) { return (int)(point[0] / 1000.0 + 0.5); }

void mj_jac(const mjModel* m, const mjData*, mjtNum* jacp, mjtNum* jacr, const mjtNum point[3], int) {
  auto* fm = static_cast<FakeModel*>(m->fake);
  const int i = site_of_point(point), nv = fm->nv, ns = fm->ns;
  for (int k = 0; k < 3; ++k)
    for (int c = 0; c < nv; ++c) {
      if (jacp) jacp[k * nv + c] = fm->J[(size_t)(3 * i + k) * nv + c];
      if (jacr) jacr[k * nv + c] = fm->J[(size_t)(3 * ns + 3 * i + k) * nv + c];
    }
}

void mj_jacDot(const mjModel* m, const mjData* d, mjtNum* jacp, mjtNum* jacr, const mjtNum point[3], int) {
  auto* fm = static_cast<FakeModel*>(m->fake);
  const int i = site_of_point(point), nv = fm->nv, ns = fm->ns;
  const double v0 = d->qvel[0];
  for (int k = 0; k < 3; ++k)
    for (int c = 0; c < nv; ++c) {
      if (jacp) jacp[k * nv + c] = c == 0 ? fm->bias[3 * i + k] / v0 : 0.0;
      if (jacr) jacr[k * nv + c] = c == 0 ? fm->bias[3 * ns + 3 * i + k] / v0 : 0.0;
    }
}

}  // extern "C"
