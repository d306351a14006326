#include "nbg_common.cuh"
extern "C" int nbg_move_exp(int, int, const void *, const void *, const void *, int, double, double, void *, int64_t,
                            int64_t, int64_t, const double *, double *, void *, size_t, void *) {
    return nbg::fail(NBG_ERR_UNSUPPORTED, "nbg_move_exp: not built yet");
}
extern "C" size_t nbg_move_exp_workspace_bytes(int, int, int64_t, int64_t, int64_t) { return 0; }
