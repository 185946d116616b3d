/* tests/stubs/ref_example/GLFW/glfw3.h -- the six GLFW calls of the example drivers (declarations). */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
typedef struct GLFWwindow GLFWwindow;
typedef struct GLFWmonitor GLFWmonitor;
int glfwInit(void);
GLFWwindow* glfwCreateWindow(int width, int height, const char* title, GLFWmonitor* monitor, GLFWwindow* share);
void glfwMakeContextCurrent(GLFWwindow* window);
void glfwSwapInterval(int interval);
void glfwGetFramebufferSize(GLFWwindow* window, int* width, int* height);
void glfwSwapBuffers(GLFWwindow* window);
void glfwPollEvents(void);
void glfwTerminate(void);
#ifdef __cplusplus
}
#endif
