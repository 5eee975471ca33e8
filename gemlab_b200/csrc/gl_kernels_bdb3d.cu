// 3D instantiations of the register-blocked stiffness kernel (see gl_kernels_bdb.cu) and the dimension dispatch
#define GL_BDB_SD 3
#include "gl_kernels_bdb.cu"

namespace gl {
int launch_bdb_2d(const IntegParams &p, void *stream);
int launch_bdb(const IntegParams &p, int sd, void *stream) { return sd == 2 ? launch_bdb_2d(p, stream) : launch_bdb_3d(p, stream); }
} // namespace gl
