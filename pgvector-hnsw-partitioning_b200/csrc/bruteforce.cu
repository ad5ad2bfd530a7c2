// bruteforce.cu -- placeholder until the tcgen05 scan lands
#include "index.h"
extern "C" int hb_bruteforce(hb_index *, const void *, int64_t, int, int32_t *, float *)
{
    hb::set_error("brute-force scan not implemented yet");
    return HB_ESTATE;
}
