// placeholder until the encoder lands (keeps the ABI complete)
#include "m2s_common.cuh"
using namespace m2s;
struct m2s_acoustic { int dummy; };
extern "C" int m2s_acoustic_create(const m2s_acoustic_config*, const m2s_tensor*, int32_t, m2s_acoustic**) { return fail(M2S_ERR_UNSUPPORTED, "acoustic model not built yet"); }
extern "C" void m2s_acoustic_destroy(m2s_acoustic*) {}
extern "C" size_t m2s_acoustic_workspace_bytes(const m2s_acoustic*, int32_t, int32_t) { return 0; }
extern "C" int m2s_acoustic_forward(m2s_acoustic*, const float*, int32_t, int32_t, const int32_t*, const int32_t*, float*, void*, size_t, m2s_stream_t) { return fail(M2S_ERR_UNSUPPORTED, "acoustic model not built yet"); }
extern "C" int m2s_acoustic_encode(m2s_acoustic*, const float*, int32_t, float*, void*, size_t, m2s_stream_t) { return fail(M2S_ERR_UNSUPPORTED, "acoustic model not built yet"); }
extern "C" int m2s_acoustic_rnn_head(m2s_acoustic*, const float*, int32_t, int32_t, const int32_t*, const int32_t*, float*, void*, size_t, m2s_stream_t) { return fail(M2S_ERR_UNSUPPORTED, "acoustic model not built yet"); }
extern "C" int m2s_acoustic_launches(const m2s_acoustic*) { return 0; }
