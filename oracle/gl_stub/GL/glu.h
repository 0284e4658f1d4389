/* oracle/gl_stub/GL/glu.h -- empty stand-in for cuda_helper/helper_cuda_gl.h:25. TEST INFRASTRUCTURE ONLY. */
#ifndef VH_GLU_STUB_H
#define VH_GLU_STUB_H
static inline const unsigned char* gluErrorString(unsigned int) { return (const unsigned char*)""; }
#endif
