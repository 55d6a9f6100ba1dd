// phase_consistency (webp_inference.py:531-550) -- placeholder, implemented next.
#include "common.cuh"
extern "C" int ddpmir_phase_reference(const float* ref, int planes, int H, int W, float* phasor, float* ws, ddpmir_stream_t stream) {
    ddpmir_set_error("phase_reference: not implemented yet");
    return DDPMIR_ERR_UNSUPPORTED;
}
extern "C" int ddpmir_phase_consistency(const float* x, const float* phasor, float alpha, int planes, int H, int W, float* out, float* ws, ddpmir_stream_t stream) {
    ddpmir_set_error("phase_consistency: not implemented yet");
    return DDPMIR_ERR_UNSUPPORTED;
}
