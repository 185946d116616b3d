#pragma once
// walter_sr/operational_space_controller.h -- drop-in for the reference header of the same
// path: same class name (global namespace), constructor and public methods
// (reference walter_sr/operational_space_controller.h), backed by the B200 library.
// Like the reference, one robot per translation unit (all three define the same names).
#include "operational-space-control/walter_sr/containers.h"
#include "operational-space-control/compat/controller_impl.h"

using namespace operational_space_controller::constants;
using namespace operational_space_controller::containers;
using namespace operational_space_controller::aliases;
using namespace osqp;

namespace osc_b200 {
struct walter_sr_traits {
    static constexpr int nv = model::nv_size, nu = model::nu_size,
                         nc = model::contact_site_ids_size, ns = model::site_ids_size;
    using State = operational_space_controller::containers::State;
    using OSCData = operational_space_controller::containers::OSCData;
    using TaskspaceTargets = operational_space_controller::aliases::TaskspaceTargets;
    using TorqueVector = Vector<model::nu_size>;
    using SolutionVector = Vector<optimization::design_vector_size>;
    using Points = Matrix<model::site_ids_size, 3>;

    static void fill_spec(osc_robot_spec& spec) {
        spec = osc_robot_spec{};
        spec.nv = nv; spec.nu = nu; spec.nc = nc; spec.ns = ns;
        for (int i = 0; i < ns; ++i) {
            spec.w_trans[i] = weights::translational[i];
            spec.w_rot[i] = weights::rotational[i];
        }
        spec.w_torque = weights::torque;
        spec.w_reg = weights::regularization;
        spec.mu = weights::friction_coefficient;
        for (int j = 0; j < nu; ++j) { spec.u_lb[j] = -1000.0; spec.u_ub[j] = 1000.0; }  // reference :309-320
        spec.fz_max = 1e4;  // big_number
    }

#if OSC_B200_HAVE_MUJOCO
    // name -> id look-ups of initialize() (reference :127-158)
    static absl::Status resolve_ids(const mjModel* m, std::vector<int>& site_ids, std::vector<int>& body_ids) {
        site_ids.clear(); body_ids.clear();
        for (auto name : model::site_list) {
            int id = mj_name2id(m, mjOBJ_SITE, std::string(name).c_str());
            if (id < 0) return absl::InternalError("site not found in model");
            site_ids.push_back(id);
        }
        for (auto name : model::body_list) {
            int id = mj_name2id(m, mjOBJ_BODY, std::string(name).c_str());
            if (id < 0) return absl::InternalError("body not found in model");
            body_ids.push_back(id);
        }
        return absl::OkStatus();
    }
    // update_mj_data (reference :394-432): floating base at the origin, FK + velocity pass.
    // The state is copied INTO mj_data (the reference re-points qpos/qvel at stack locals).
    static void update_mj_data(const mjModel* m, mjData* d, const State& s, const std::vector<int>& site_ids, Points& points) {
        for (int i = 0; i < 3; ++i) d->qpos[i] = 0.0;
        for (int i = 0; i < 4; ++i) d->qpos[3 + i] = s.body_rotation(i);
        for (int i = 0; i < nu; ++i) d->qpos[7 + i] = s.motor_position(i);
        for (int i = 0; i < 3; ++i) { d->qvel[i] = s.linear_body_velocity(i); d->qvel[3 + i] = s.angular_body_velocity(i); }
        for (int i = 0; i < nu; ++i) d->qvel[6 + i] = s.motor_velocity(i);
        mj_fwdPosition(m, d);
        mj_fwdVelocity(m, d);
        for (int i = 0; i < ns; ++i) {
            const int sid = true ? site_ids[i] : i;
            for (int k = 0; k < 3; ++k) points(i, k) = d->site_xpos[3 * sid + k];
        }
    }
    // update_osc_data (reference :434-513): M, bias forces, stacked [Jp;Jr], Jdot*qvel
    static void update_osc_data(const mjModel* m, mjData* d, const std::vector<int>& body_ids, const Points& points, OSCData& out) {
        mj_fullM(m, out.mass_matrix.data(), d->qM);
        for (int i = 0; i < nv; ++i) out.coriolis_matrix(i) = d->qfrc_bias[i];
        double jp[3 * nv], jr[3 * nv], jpd[3 * nv], jrd[3 * nv];
        for (int r = 0; r < 6 * ns; ++r) out.taskspace_bias(r) = 0.0;
        for (int i = 0; i < ns; ++i) {
            double pt[3] = {points(i, 0), points(i, 1), points(i, 2)};
            mj_jac(m, d, jp, jr, pt, body_ids[i]);
            mj_jacDot(m, d, jpd, jrd, pt, body_ids[i]);
            for (int k = 0; k < 3; ++k) {
                double bp = 0.0, br = 0.0;
                for (int c = 0; c < nv; ++c) {
                    out.taskspace_jacobian(3 * i + k, c) = jp[k * nv + c];
                    out.taskspace_jacobian(3 * ns + 3 * i + k, c) = jr[k * nv + c];
                    bp += jpd[k * nv + c] * d->qvel[c];
                    br += jrd[k * nv + c] * d->qvel[c];
                }
                out.taskspace_bias(3 * i + k) = bp;
                out.taskspace_bias(3 * ns + 3 * i + k) = br;
            }
        }
        for (int c = 0; c < nv; ++c)
            for (int k = 0; k < 3 * nc; ++k)
                out.contact_jacobian(c, k) = out.taskspace_jacobian(3 * ns - 3 * nc + k, c);
    }
#endif
};
}  // namespace osc_b200

using OperationalSpaceController = osc_b200::Controller<osc_b200::walter_sr_traits>;
using BatchedOperationalSpaceController = osc_b200::BatchedController<osc_b200::walter_sr_traits>;
