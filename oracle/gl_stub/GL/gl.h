/* oracle/gl_stub/GL/gl.h -- minimal stand-in so the reference's cuda_helper/helper_cuda_gl.h:24
 * (#include <GL/gl.h>) resolves on a box without OpenGL headers.  TEST INFRASTRUCTURE ONLY.
 * Only the scalar typedefs are provided; no GL function is declared or called headless. */
#ifndef VH_GL_STUB_H
#define VH_GL_STUB_H
typedef unsigned int GLenum;
typedef unsigned int GLuint;
typedef int GLint;
typedef int GLsizei;
typedef unsigned char GLboolean;
typedef unsigned int GLbitfield;
typedef float GLfloat;
typedef double GLdouble;
typedef void GLvoid;
typedef unsigned char GLubyte;
#define GL_NO_ERROR 0
/* referenced (never called) by the inline sdkCheckErrorGL of helper_cuda_gl.h:134-156 */
static inline GLenum glGetError(void) { return GL_NO_ERROR; }
#endif
