// decode_kernels.cu -- decoder-step kernels (filled in with the decoder milestone)
#include "ptx.cuh"
#include "wb_kernels.hpp"
