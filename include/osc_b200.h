/*
 * include/osc_b200.h -- C-ABI of the B200 batched operational-space controller.
 *
 * Drop-in boundary for ONE hot path of vannem95/operational-space-control: the
 * per-step control law  OSCData + TaskspaceTargets + contact mask -> QP -> torques.
 * The reference has no FFI of its own; its "operator API" is the C++ class
 * OperationalSpaceController (one per robot directory).  Each entry point below
 * names the reference code it replaces (paths relative to the reference's
 * operational-space-control/ directory, walter_sr lines; the wheels and go2
 * headers hold the same code at the offsets listed in SURVEY.md 8a).
 *
 * Plain C: pointers and sizes only, no torch / Eigen / abseil types.
 * All matrices are FP64, row-major and contiguous per environment, exactly
 * the memory layout of the reference's OSCData fields (containers.h:13-21):
 *   M       [n_envs][nv*nv]      osc_data.mass_matrix        (:436-438)
 *   C       [n_envs][nv]         osc_data.coriolis_matrix    (:441-442)
 *   J       [n_envs][6*ns*nv]    osc_data.taskspace_jacobian (:485-487)  rows [Jp(all sites); Jr(all sites)]
 *   bias    [n_envs][6*ns]       osc_data.taskspace_bias     (:491-492)
 *   targets [n_envs][ns*6]       TaskspaceTargets            (aliases.h:21)
 *   mask    [n_envs][nc]         State.contact_mask          (containers.h:41), 0.0 / 1.0
 * osc_data.contact_jacobian is not an input: it is rows [3ns-3nc, 3ns) of J
 * transposed (:497-503) and is read from J on the device.
 *
 * Thread-safety: one caller per handle.  Multi-GPU: one handle per device.
 * There is NO CPU fallback: every call fails with OSC_ERR_CUDA when the
 * device or the kernels are unavailable.
 */
#ifndef OSC_B200_H
#define OSC_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSC_MAX_SITES 32
#define OSC_MAX_NU 16

enum {
  OSC_OK = 0,
  OSC_ERR_INVALID = -1,     /* bad argument / unsupported robot shape */
  OSC_ERR_CUDA = -2,        /* CUDA runtime error (no device, launch failure, ...) */
  OSC_ERR_STATE = -3,       /* call order violated (step before setup, ...) */
  OSC_ERR_ALLOC = -4
};

/* per-environment solver status, OSQP's values (osqp/include/constants.h) */
enum {
  OSC_DUAL_INFEASIBLE_INACCURATE = 4,
  OSC_PRIMAL_INFEASIBLE_INACCURATE = 3,
  OSC_SOLVED_INACCURATE = 2,
  OSC_SOLVED = 1,
  OSC_MAX_ITER_REACHED = -2,
  OSC_PRIMAL_INFEASIBLE = -3, /* as in OSQP: solution / torque of that environment are NaN */
  OSC_DUAL_INFEASIBLE = -4,   /* and its iterates restart from zero at the next step      */
  OSC_NON_CVX = -7,
  OSC_UNSOLVED = -10
};

/* What the reference bakes in per robot at build time:
 * autogen_defines.h sizes (<robot>/autogen/autogen.py:469-513), the YAML's
 * weights_config / friction_coefficient (config/<robot>/ *.yaml, consumed by
 * autogen.py:20-41,331-336) and the hard-coded bounds
 * (operational_space_controller.h:282,309-353; go2 :285-308). */
typedef struct {
  int nv, nu, nc, ns;             /* model::nv_size, nu_size, contact_site_ids_size, site_ids_size */
  double w_trans[OSC_MAX_SITES];  /* <site>_translational_tracking, site order = noncontact then contact */
  double w_rot[OSC_MAX_SITES];    /* <site>_rotational_tracking */
  double w_torque, w_reg;         /* torque, regularization */
  double mu;                      /* friction_coefficient */
  double u_lb[OSC_MAX_NU], u_ub[OSC_MAX_NU];
  double fz_max;                  /* big_number */
} osc_robot_spec;

/* The subset of osqp::OsqpSettings that changes the iterates; same names and
 * OSQP 0.6.3 defaults (the reference passes OsqpSettings() untouched,
 * operational_space_controller.h:110).  adaptive_rho_interval = 0 resolves to
 * 4 x check_termination (OSQP's rule without PROFILING); a PROFILING build
 * picks it from wall-clock time, so pass it explicitly to mimic one. */
typedef struct {
  double rho, sigma, alpha;
  double eps_abs, eps_rel;
  double adaptive_rho_tolerance;
  int scaling;
  int adaptive_rho;
  int adaptive_rho_interval;
  int max_iter;
  int check_termination;
  int warm_start;
  double eps_prim_inf, eps_dual_inf; /* infeasibility certificates (util.c is_*_infeasible) */
} osc_settings;

typedef struct osc_handle osc_handle;

typedef struct {
  /* device pointers of the handle's input buffers (layouts above) for callers
   * whose data is already in HBM; write them on the stream passed to osc_step */
  double *M, *C, *J, *bias, *targets, *mask;
  /* device pointers of the outputs */
  double *torque;   /* [n_envs][nu]   torque_command  (:631) */
  double *solution; /* [n_envs][n]    solution        (:592) */
  double *dual;     /* [n_envs][m]    dual_solution   (:593) */
  int *iters;       /* [n_envs]       OSQP info->iter */
  int *status;      /* [n_envs]       OSC_SOLVED ...  (exit_code, :591) */
  double *pri_res, *dua_res, *rho; /* [n_envs] */
} osc_device_buffers;

/* OsqpSettings() defaults. */
int osc_default_settings(osc_settings *s);

/* Replaces the controller constructor + initialize()'s allocation
 * (operational_space_controller.h:110-163) for n_envs environments on CUDA
 * device `device`.  Only the reference's shapes are compiled:
 * (nv,nu,nc,ns) = (14,8,8,17) [walter_sr, walter_sr_wheels], (18,12,4,5) [unitree_go2]. */
int osc_create(const osc_robot_spec *spec, const osc_settings *settings, int n_envs, int device,
               osc_handle **out);
int osc_destroy(osc_handle *h);
/* message of the last error on this handle (or of a failed osc_create when h == NULL) */
const char *osc_last_error(const osc_handle *h);

int osc_num_envs(const osc_handle *h);
int osc_get_device_buffers(osc_handle *h, osc_device_buffers *out);

/* Host AoS -> device: cudaMemcpyAsync on `stream` straight from the caller's arrays (no
 * staging copy).  The copies run at full PCIe rate and asynchronously only from page-locked
 * memory (osc_host_alloc); from pageable memory every copy blocks the host.  Replaces the
 * six row->column-major copies of update_optimization_data (:517-522): no
 * transposes are needed, the kernels read the reference's row-major layout.
 * Any pointer may be NULL to leave that field unchanged. `stream` is a cudaStream_t. */
int osc_upload(osc_handle *h, const double *M, const double *C, const double *J,
               const double *bias, const double *targets, const double *mask, void *stream);

/* set_up_optimization() for every environment (:355-392): first QP build and
 * solver Init (cold iterates, rho = settings.rho).  Must precede osc_step. */
int osc_setup(osc_handle *h, void *stream);

/* One control_loop body after update_osc_data() for every environment:
 * update_optimization_data (:515-539, CasADi H,f,Aeq,beq,Aineq,bineq),
 * update_optimization (:541-587, bounds with contact mask, OSQP data update that
 * keeps the warm start), solve_optimization (:589-594) and the torque slice (:631). */
int osc_step(osc_handle *h, void *stream);

/* osc_step runs two kernels by default: the objective build (CasADi H, f, :529-530) inside the
 * equilibration kernel, then the solve.  on = 0 selects the three-kernel form (separate
 * build_qp_kernel) -- same results, kept for measurements of the build kernel alone. */
int osc_set_fused_build(osc_handle *h, int on);

/* reset_optimization() (:596-601): zero primal/dual warm start. */
int osc_reset_warm_start(osc_handle *h, void *stream);

/* Device -> host copies of the results of the last step (async on `stream`;
 * call osc_sync before reading).  NULL pointers are skipped. */
int osc_download(osc_handle *h, double *torque, double *solution, double *dual, int *iters,
                 int *status, double *pri_res, double *dua_res, double *rho, void *stream);
int osc_sync(osc_handle *h, void *stream);
/* The objective the last osc_setup/osc_step built: H's dv block [n_envs][nv*nv] and f's dv
 * part [n_envs][nv] (OptimizationData.H / .f, containers.h:23-25; all other entries of H and f
 * are the constants 2 w_reg (+2 w_torque on u) and 0).  Inspection / tests. */
int osc_download_objective(osc_handle *h, double *H_dv, double *f_dv, void *stream);

/* update_state/update_taskspace_targets + control step + get_torque_command for a
 * whole batch with HOST buffers: upload, step, download torques, synchronise.
 * The upload is pipelined in chunks of environments against the kernels, and skips the rows
 * of the task Jacobian nothing reads: rows of a (site, translational|rotational) block whose
 * objective weight is zero and that are not contact rows (H = 2 J'WJ gives them weight 0 and
 * contact_jacobian is the contact rows only).  Everything else goes up in full.
 * A handle of a few robots (inputs <= 128 KiB: the reference's own one-robot 1 kHz loop) takes a
 * low-latency route instead: the inputs are staged in one pinned slab, uploaded with a single
 * copy on `stream`, and the torques return through pinned memory.
 * The chunked pipeline copies straight from / to the caller's buffers: the overlap of the H2D
 * copy of chunk c+1 with the kernels of chunk c needs PAGE-LOCKED host buffers (osc_host_alloc);
 * with pageable memory the call is still correct but every copy blocks.  Fails with
 * OSC_ERR_STATE while any input is bound to caller-owned device memory. */
int osc_step_host(osc_handle *h, const double *M, const double *C, const double *J,
                  const double *bias, const double *targets, const double *mask, double *torque,
                  void *stream);

/* osc_step_host with the task Jacobian handed over in FP32 (opt-in: a producer that already
 * holds J in single precision -- MuJoCo built with mjUSESINGLE, an FP32 simulator -- halves the
 * 77 % of the host-link bytes that are J).  The rows are widened to FP64 on the device
 * (widen_rows_kernel) and EVERYTHING downstream stays FP64: with J values that are exactly
 * representable in FP32 the results are those of osc_step_host bit for bit; otherwise the QP
 * data differ by the rounding of J (relative 6e-8), measured against the parity gate in
 * tests/test_gpu_parity.py and reported separately by bench.py, never the headline path. */
int osc_step_host_j32(osc_handle *h, const double *M, const double *C, const float *J32,
                      const double *bias, const double *targets, const double *mask,
                      double *torque, void *stream);

/* Self-test of the warp primitives the register-resident solver core is written against
 * (csrc/osc_warp.cuh): runs the shuffles, the 16-way transposing max, the warp sum and one FP64
 * tensor-core tile product on `in` ([18][32] doubles: rows 0-15 non-negative values, row 16 an
 * mma.m8n8k4 A fragment, row 17 a B fragment) and returns, in `out` (154 doubles): max16[16],
 * sum(row 0), 0, xchg16(row 0)[32], group4(row 0, 2)[32], and the D fragments d0[32], d1[32]
 * of D = A B + D0 with D0 = (+1, -1), and the 8-way max of rows 0-7.  tests/ compare it with
 * the host emulation of the same primitives (what lets the CPU suite run the device solver
 * source). */
int osc_selftest_warp(int device, const double *in, double *out);

/* Bytes the last osc_step_host moved over PCIe: host -> device (inputs) and device -> host
 * (torques). */
int osc_host_traffic(const osc_handle *h, size_t *h2d_bytes, size_t *d2h_bytes);

/* ---- CONDENSED fast mode (BASELINE.json north_star subsystems (1) and (2); reported
 * separately, NEVER the path that is gated against the reference's OSQP iterates).  The
 * reference keeps  M dv + C = B u + Jc z  as equality rows (walter_sr/autogen/autogen.py:82-93);
 * dv is unbounded (operational_space_controller.h:286-287), so it can be eliminated exactly:
 * Cholesky of M, G = M^-1 [B Jc], QP in (u, z) only with P' = G' Hd G + R (n' = nu + 3 nc
 * variables, 4 nc + n' rows).  Same unique optimum, different ADMM iterates: its torques are
 * a solution of the same QP to the same OSQP tolerances, not the reference's iterate.  Its
 * oracle is oracle/osc_condensed.py.  Every step is osqp_setup on the new condensed data with
 * rho carried over, osqp_warm_start(previous solution), osqp_solve (the shape of the
 * reference's re-Init branch, :571-584).  No osc_setup needed.  Outputs go to the same
 * buffers as osc_step: torque, solution = [G w + d0; u; z], dual (friction and box rows; the
 * eliminated dynamics rows and the free dv rows report 0), iters, status, residuals, rho. */
int osc_step_condensed(osc_handle *h, void *stream);
/* cold start (zero warm start, rho = settings.rho) for the next osc_step_condensed */
int osc_reset_condensed(osc_handle *h, void *stream);

/* Use caller-owned DEVICE memory as the inputs of subsequent osc_setup/osc_step calls
 * (same layouts; NULL keeps the handle's own buffer for that field).  Lets a rollout
 * loop or a benchmark keep several batches resident in HBM without copies. */
int osc_bind_device_inputs(osc_handle *h, const double *M, const double *C, const double *J,
                           const double *bias, const double *targets, const double *mask);

/* Page-locked host memory for the caller's OSCData arrays, so osc_upload /
 * osc_step_host copy at full PCIe rate without a staging pass. */
int osc_host_alloc(size_t bytes, void **out);
int osc_host_free(void *p);

/* Per-kernel device timing: when enabled, every osc_step brackets its kernels with CUDA
 * events on the launch stream; osc_timing_read synchronises, returns the average
 * milliseconds per step of each kernel since the last read, and resets.
 * scale = the fused objective-build + equilibration kernel (H, f, OSQP scale_data); solve =
 * assembly, factorisation, ADMM, un-scaling; build = the gap between the step's first event
 * and the fused kernel (a few microseconds) -- or, after osc_set_fused_build(h, 0) and in
 * osc_step_condensed, the stand-alone objective-build kernel, with scale = equilibration alone. */
typedef struct {
  float build_ms, solve_ms;
  int steps;
  float scale_ms;
} osc_kernel_times;
int osc_timing_enable(osc_handle *h, int on);
int osc_timing_read(osc_handle *h, osc_kernel_times *out);

/* ---- the step BEFORE the hot path: per-environment task-space PD
 * targets and the contact mask, so a roll-out loop can keep both on the device. ---- */

/* DEVICE pointers, [n_envs][ns][3] (pos, vel, angvel) / [n_envs][ns][4] (quat, w x y z).
 * Desired velocities may be NULL (= 0); everything else is required. */
typedef struct {
  const double *pos, *quat, *vel, *angvel;
  const double *pos_des, *quat_des, *vel_des, *angvel_des;
} osc_site_state;

/* examples/standing.cc:146-155 generalised to every task site (the Walter drivers' per-site
 * laws are the same PD form, e.g. walter_sr_true_tumbling_mjjoint.cc:873-973):
 *   targets[e][i][0:3] = kp_lin[i] (p_des - p) + kd_lin[i] (v_des - v)
 *   targets[e][i][3:6] = kp_ang[i] vec(q_des (x) conj(q)) + kd_ang[i] (w_des - w)
 * Gains are host arrays of ns entries.  Writes the handle's own `targets` input buffer. */
int osc_targets_pd(osc_handle *h, const osc_site_state *s, const double *kp_lin,
                   const double *kd_lin, const double *kp_ang, const double *kd_ang,
                   void *stream);

/* The Walter tumbling driver's own target laws (examples/walter_sr_true_tumbling_mjjoint.cc,
 * BASELINE.json configs[2]) for every environment, into the handle's `targets` input:
 *   rows 1-4, shins (:695-802):  alpha_y = shin_kp ((th0 + shin_rate t) - th) + shin_kv (shin_rate - (th - th_prev)/dt)
 *   rows 5-8, thighs (:873-973): a_z = thigh_kp ((z0 + thigh_height_offset) - z) + thigh_kv (thigh_rate - (z - z_prev)/dt)
 *   row 0, torso (:1001-1019): every gain is zero in that driver -> zeros; contact rows: zeros.
 * shin_* / thigh_*: DEVICE arrays [n_envs][4] in the driver's leg order (tl, tr, hl, hr):
 * joint angle (qpos[jnt_qposadr[2,4,6,8]]) / thigh site height (site_xpos z), their values at the
 * previous control step and at t = 0.  Only the Walter shapes (ns = 17). */
typedef struct {
  double shin_kp, shin_kv, shin_rate;       /* 2400, 2400, 4 rad/s */
  double thigh_kp, thigh_kv, thigh_rate;    /* 2000, 300, 0 */
  double thigh_height_offset;               /* -0.025 m */
} osc_walter_tumbling_gains;
int osc_walter_tumbling_default_gains(osc_walter_tumbling_gains *g);
int osc_targets_walter_tumbling(osc_handle *h, const osc_walter_tumbling_gains *g,
                                const double *shin_angle, const double *shin_angle_prev,
                                const double *shin_angle0, const double *thigh_z,
                                const double *thigh_z_prev, const double *thigh_z0, double time,
                                double dt, void *stream);

/* Contact mask from MuJoCo contact pairs (walter_sr_true_tumbling_mjjoint.cc:523-558):
 * every contact k < ncon[e] whose geom[0] or geom[1] is a listed id (`wheel_sites_mujoco`,
 * :436) contributes the first site on that geom's body (getSiteIdsOnSameBodyAsGeom, :106);
 * mask[e][c] = 1.0 iff list[c] is among the contributed sites
 * (getBinaryRepresentation_std_find, :152).
 * geom_pairs: DEVICE int [n_envs][max_con][2]; ncon: DEVICE int [n_envs];
 * contact_geom_ids: host int [nc] (the list); site_of_geom: host int [nc], the site each
 * listed geom maps to, or NULL when site and geom ids coincide (the Walter models).
 * Writes the handle's own `mask` input buffer. */
int osc_contact_mask_from_contacts(osc_handle *h, const int *geom_pairs, const int *ncon,
                                   int max_con, const int *contact_geom_ids,
                                   const int *site_of_geom, void *stream);

/* ---- the step before THAT: the MuJoCo-derived record on the device.
 * update_mj_data / update_osc_data (operational_space_controller.h:394-513) for a floating-base
 * tree of hinge joints, from qpos / qvel and a host-supplied model: M (mj_fullM, :436-438),
 * C (qfrc_bias, :441-442), J = [jacp; jacr] per task site (mj_jac, :459-487) and
 * bias = Jdot qvel (mj_jacDot, :491-492) are written into the handle's OWN M, C, J, bias input
 * buffers in the OSCData layouts, so that 13.4 kB per environment and step never cross PCIe.
 * MuJoCo conventions: qpos = [base xyz, base quat (w x y z), hinge angles], qvel = [world-frame
 * base linear velocity, BODY-frame base angular velocity, hinge rates]; body b >= 1 hangs off
 * parent[b] < b through one hinge about jaxis[b] anchored at its own origin.  The robots' MJCF
 * files are external to the reference (not in this repository): the model numbers are the
 * caller's; oracle/osc_kinematics.py is the CPU restatement this is tested against. */
#define OSC_KIN_MAX_BODIES 16
typedef struct {
  int nb, ns;                            /* bodies (nv = 6 + nb - 1), task sites (= spec.ns) */
  int parent[OSC_KIN_MAX_BODIES];        /* parent[0] = -1 */
  double bpos[OSC_KIN_MAX_BODIES][3];    /* body origin in the parent frame */
  double bquat[OSC_KIN_MAX_BODIES][4];   /* body orientation in the parent frame at zero angle */
  double jaxis[OSC_KIN_MAX_BODIES][3];   /* hinge axis, body frame, unit length */
  double mass[OSC_KIN_MAX_BODIES];
  double ipos[OSC_KIN_MAX_BODIES][3];    /* centre of mass, body frame */
  double inertia[OSC_KIN_MAX_BODIES][6]; /* about the centre of mass, body frame: xx yy zz xy xz yz */
  int site_body[OSC_MAX_SITES];
  double site_pos[OSC_MAX_SITES][3];     /* body frame; site order = the controller's site_list */
  double gravity[3];
} osc_kin_model;
/* qpos [n_envs][nv + 1], qvel [n_envs][nv]: DEVICE pointers. */
int osc_kinematics(osc_handle *h, const osc_kin_model *model, const double *qpos,
                   const double *qvel, void *stream);

/* How many environment-steps so far took the reference's sparsity-change path
 * (update_optimization :571-584: UpdateObjectiveAndConstraintMatrices rejected the new
 * pattern -> solver re-Init with rho reset + SetWarmStart(solution, dual_solution)).
 * Synchronises `stream`. */
int osc_reinit_count(osc_handle *h, int *count, void *stream);

/* ---- multi-GPU: the all-gather of torques and statistics after a step (the
 * reference has no counterpart -- one controller, one robot, one thread).  Environments are
 * sharded, one handle (and normally one process) per GPU; there is no collective on the hot
 * path.  The gather is done with PEER STORES instead of a collective library call: every rank
 * owns a gathered slab
 *     torque_all [world][n_envs][nu]   followed by   stats_all [world][OSC_GATHER_STATS]
 * that the other ranks map (CUDA IPC between processes, plain device pointers inside one
 * process); osc_gather_torques launches one kernel that writes the rank's slice and its
 * statistics into every mapped slab over NVLink / NVSwitch.  All ranks must use the same
 * n_envs.  A rank's view is complete once every rank's kernel has finished: synchronise the
 * ranks (barrier / events) before reading.
 *   stats_all[r] = { n_envs, #solved, sum iters, max iters, max pri_res, max dua_res,
 *                    #re-Inits so far, gather sequence number } of rank r.
 * Call order: osc_gather_create on every rank -> exchange the 64-byte handles by any means
 * (MPI, torch.distributed, a file) -> osc_gather_attach -> osc_gather_torques after steps. */
#define OSC_IPC_HANDLE_BYTES 64
#define OSC_GATHER_STATS 8
/* allocates the slab; ipc_handle_out (64 bytes, may be NULL inside one process) names it */
int osc_gather_create(osc_handle *h, int rank, int world, void *ipc_handle_out);
/* ipc_handles: world x 64 bytes in rank order (other processes), OR peer_slabs: world device
 * pointers from osc_gather_buffers of handles in THIS process (entry `rank` is ignored) */
int osc_gather_attach(osc_handle *h, const void *ipc_handles, double *const *peer_slabs);
int osc_gather_torques(osc_handle *h, void *stream);
/* device pointers of this rank's gathered copies */
int osc_gather_buffers(osc_handle *h, double **torque_all, double **stats_all);

/* number of kernels launched by this handle so far (bench bookkeeping) */
long long osc_kernel_launches(const osc_handle *h);

/* FP64 FMA micro-benchmark used as the roofline denominator for the solver
 * kernel: returns achieved TFLOP/s (2 flop per DFMA) on `device`. */
int osc_measure_dfma_tflops(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
