// build.cu -- placeholder until the batched insert pipeline lands (next commit)
#include "index.h"
namespace hb {
int64_t build_insert(hb_index *, const void *, int64_t, const int64_t *)
{
    set_error("build path not implemented yet");
    return HB_ESTATE;
}
}
