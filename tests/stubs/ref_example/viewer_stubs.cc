// tests/stubs/ref_example/viewer_stubs.cc -- empty bodies for the viewer / simulation calls of the example
// drivers (GLFW, mjv_*, mjr_*, mj_step ...), so that the unmodified example LINKS against the
// drop-in controller + tests/stubs/fake_mujoco.cc.  Nothing here is ever measured or shipped.
#include "GLFW/glfw3.h"
#include "mujoco/mujoco.h"
extern "C" {
int glfwInit(void) { return 1; }
GLFWwindow* glfwCreateWindow(int, int, const char*, GLFWmonitor*, GLFWwindow*) { return nullptr; }
void glfwMakeContextCurrent(GLFWwindow*) {}
void glfwSwapInterval(int) {}
void glfwGetFramebufferSize(GLFWwindow*, int* w, int* h) { *w = 800; *h = 600; }
void glfwSwapBuffers(GLFWwindow*) {}
void glfwPollEvents(void) {}
void glfwTerminate(void) {}
void mjv_defaultCamera(mjvCamera*) {}
void mjv_defaultPerturb(mjvPerturb*) {}
void mjv_defaultOption(mjvOption*) {}
void mjv_defaultScene(mjvScene*) {}
void mjr_defaultContext(mjrContext*) {}
void mjv_makeScene(const mjModel*, mjvScene*, int) {}
void mjr_makeContext(const mjModel*, mjrContext*, int) {}
void mjv_updateScene(const mjModel*, mjData*, const mjvOption*, const mjvPerturb*, mjvCamera*, int, mjvScene*) {}
void mjr_render(mjrRect, mjvScene*, const mjrContext*) {}
void mjv_freeScene(mjvScene*) {}
void mjr_freeContext(mjrContext*) {}
void mj_resetDataKeyframe(const mjModel*, mjData*, int) {}
void mj_forward(const mjModel* m, mjData* d) { mj_fwdPosition(m, d); }
void mj_step(const mjModel* m, mjData* d) { d->time += m->opt.timestep; }
}
