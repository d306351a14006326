#include "nbg_common.cuh"
extern "C" size_t nbg_group_workspace_bytes(int, int, int64_t, int64_t) { return 0; }
extern "C" int nbg_group_init(int, int, void *, int64_t, int64_t, void *) {
    return nbg::fail(NBG_ERR_UNSUPPORTED, "nbg_group: not built yet");
}
extern "C" int nbg_group_accumulate(int, int, int, const void *, const void *, int, void *, int64_t, int64_t, int64_t,
                                    int64_t, void *) {
    return nbg::fail(NBG_ERR_UNSUPPORTED, "nbg_group: not built yet");
}
extern "C" int nbg_group_combine(int, int, void *, const void *, int64_t, int64_t, void *) {
    return nbg::fail(NBG_ERR_UNSUPPORTED, "nbg_group: not built yet");
}
extern "C" int nbg_group_finalize(int, int, const void *, void *, int64_t, int64_t, int64_t, void *) {
    return nbg::fail(NBG_ERR_UNSUPPORTED, "nbg_group: not built yet");
}
extern "C" int nbg_group(int, int, int, const void *, const void *, int, void *, int64_t, int64_t, int64_t, int64_t,
                         void *, size_t, void *) {
    return nbg::fail(NBG_ERR_UNSUPPORTED, "nbg_group: not built yet");
}
