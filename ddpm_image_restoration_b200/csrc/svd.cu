// svd_structure_preservation (0409_method.ipynb#c0:L321-346) -- placeholder, implemented next.
#include "common.cuh"
extern "C" int ddpmir_svd_lowrank(const float* x, int planes, int H, int W, int k, float* out, float* ws, int sweeps, ddpmir_stream_t stream) {
    ddpmir_set_error("svd_lowrank: not implemented yet");
    return DDPMIR_ERR_UNSUPPORTED;
}
