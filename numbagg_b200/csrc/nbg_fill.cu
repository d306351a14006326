#include "nbg_common.cuh"
extern "C" int nbg_fill(int, int, const void *, void *, int64_t, int64_t, int64_t, int64_t, const int64_t *, int64_t *,
                        void *, size_t, void *) {
    return nbg::fail(NBG_ERR_UNSUPPORTED, "nbg_fill: not built yet");
}
extern "C" size_t nbg_fill_workspace_bytes(int, int64_t, int64_t, int64_t) { return 0; }
