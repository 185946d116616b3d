/* tests/stubs/mujoco/mujoco.h -- DECLARATIONS ONLY: the slice of MuJoCo's C API that the
 * drop-in controller classes (operational-space-control_b200/<robot>/operational_space_controller.h,
 * compat/controller_impl.h) and the reference's example drivers (examples/walter_sr_standing.cc)
 * name.  Test infrastructure: MuJoCo is not in this image, so without this header the
 * OSC_B200_HAVE_MUJOCO branch of the classes would never meet a compiler.  Field and function
 * names, argument orders and types follow MuJoCo 3.2 (mjmodel.h / mjdata.h / mujoco.h /
 * mjvisualize.h / mjrender.h); only members somebody here reads are present.
 * tests/stubs/fake_mujoco.cc is a scripted implementation of the physics subset for tests. */
#ifndef OSC_B200_TESTS_STUB_MUJOCO_H
#define OSC_B200_TESTS_STUB_MUJOCO_H

#ifdef __cplusplus
extern "C" {
#endif

typedef double mjtNum;

typedef enum { mjOBJ_UNKNOWN = 0, mjOBJ_BODY = 1, mjOBJ_GEOM = 5, mjOBJ_SITE = 6 } mjtObj;
typedef enum { mjCAT_ALL = 7 } mjtCatBit;
typedef enum { mjFONTSCALE_150 = 150 } mjtFontScale;

typedef struct { double timestep; } mjOption;

typedef struct mjModel_ {
  int nq, nv, nu, nbody, nsite, ngeom, nkey;
  mjOption opt;
  mjtNum *key_qpos, *key_qvel, *key_ctrl;
  void* fake;  /* owned by the fake backend */
} mjModel;

typedef struct { int geom[2]; int geom1, geom2; } mjContact;

typedef struct mjData_ {
  double time;
  int ncon;
  mjtNum *qpos, *qvel, *ctrl, *qfrc_actuator, *qfrc_bias, *qM;
  mjtNum *site_xpos, *site_xmat, *xpos, *xmat;
  mjContact* contact;
  void* fake;
} mjData;

mjModel* mj_loadXML(const char* filename, const void* vfs, char* error, int error_sz);
mjData* mj_makeData(const mjModel* m);
void mj_deleteData(mjData* d);
void mj_deleteModel(mjModel* m);
int mj_name2id(const mjModel* m, int type, const char* name);
void mj_resetDataKeyframe(const mjModel* m, mjData* d, int key);
void mj_forward(const mjModel* m, mjData* d);
void mj_step(const mjModel* m, mjData* d);
void mj_fwdPosition(const mjModel* m, mjData* d);
void mj_fwdVelocity(const mjModel* m, mjData* d);
void mj_fullM(const mjModel* m, mjtNum* dst, const mjtNum* M);
void mj_jac(const mjModel* m, const mjData* d, mjtNum* jacp, mjtNum* jacr, const mjtNum point[3], int body);
void mj_jacDot(const mjModel* m, const mjData* d, mjtNum* jacp, mjtNum* jacr, const mjtNum point[3], int body);

/* visualisation (examples only; never implemented here: compile-only) */
typedef struct { int type; double lookat[3], distance, azimuth, elevation; } mjvCamera;
typedef struct { int select; } mjvPerturb;
typedef struct { int label; } mjvOption;
typedef struct { int maxgeom; } mjvScene;
typedef struct { int fontScale; } mjrContext;
typedef struct { int left, bottom, width, height; } mjrRect;
void mjv_defaultCamera(mjvCamera* cam);
void mjv_defaultPerturb(mjvPerturb* pert);
void mjv_defaultOption(mjvOption* opt);
void mjv_defaultScene(mjvScene* scn);
void mjr_defaultContext(mjrContext* con);
void mjv_makeScene(const mjModel* m, mjvScene* scn, int maxgeom);
void mjr_makeContext(const mjModel* m, mjrContext* con, int fontscale);
void mjv_updateScene(const mjModel* m, mjData* d, const mjvOption* opt, const mjvPerturb* pert,
                     mjvCamera* cam, int catmask, mjvScene* scn);
void mjr_render(mjrRect viewport, mjvScene* scn, const mjrContext* con);
void mjv_freeScene(mjvScene* scn);
void mjr_freeContext(mjrContext* con);

#ifdef __cplusplus
}
#endif
#endif
