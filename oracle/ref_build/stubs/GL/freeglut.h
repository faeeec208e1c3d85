// No-op OpenGL/GLUT stand-in so the reference's engine host class (CudaKernel.cpp:313-388,
// GL blit in render_end) compiles and runs headless. Test infrastructure only.
#ifndef SOLR_B200_GL_STUB_H
#define SOLR_B200_GL_STUB_H
typedef unsigned int GLenum; typedef int GLint; typedef int GLsizei; typedef float GLfloat; typedef void GLvoid;
#define GL_TEXTURE_2D 0x0DE1
#define GL_TEXTURE_MIN_FILTER 0x2801
#define GL_TEXTURE_MAG_FILTER 0x2800
#define GL_NEAREST 0x2600
#define GL_RGB 0x1907
#define GL_UNSIGNED_BYTE 0x1401
#define GL_QUADS 0x0007
static inline void glEnable(GLenum) {}
static inline void glDisable(GLenum) {}
static inline void glTexParameterf(GLenum, GLenum, GLfloat) {}
static inline void glTexImage2D(GLenum, GLint, GLint, GLsizei, GLsizei, GLint, GLenum, GLenum, const GLvoid*) {}
static inline void glBegin(GLenum) {}
static inline void glEnd() {}
static inline void glTexCoord2f(GLfloat, GLfloat) {}
static inline void glVertex3f(GLfloat, GLfloat, GLfloat) {}
#endif
