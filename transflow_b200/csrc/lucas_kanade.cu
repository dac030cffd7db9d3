// placeholder replaced below in this round: pyramidal Lucas-Kanade
#include "common.cuh"
using namespace tf;
struct tf_lucas_kanade { int H, W; };
extern "C" int tf_lk_create(tf_lucas_kanade** out, int, int, int, int, int) { (void)out; return fail(TF_ERR_INVALID_ARG, "lucas-kanade: not built yet"); }
extern "C" int tf_lk_destroy(tf_lucas_kanade*) { return TF_OK; }
extern "C" int tf_lk_run(tf_lucas_kanade*, const uint8_t*, const uint8_t*, float*, int, void*) { return fail(TF_ERR_INVALID_ARG, "lucas-kanade: not built yet"); }
