// 2D instantiations of the register-blocked stiffness kernel (see gl_kernels_bdb.cu)
#define GL_BDB_SD 2
#include "gl_kernels_bdb.cu"
