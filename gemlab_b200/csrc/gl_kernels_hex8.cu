// gl_kernels_hex8.cu -- register-blocked Hex8 stiffness kernel (placeholder: not yet specialised; the generic kernel runs)
#include <cuda_runtime.h>

#include "gl_internal.hpp"

namespace gl {

int launch_hex8_bdb(const IntegParams &p, void *stream) {
    (void)p;
    (void)stream;
    return 0;
}

} // namespace gl
