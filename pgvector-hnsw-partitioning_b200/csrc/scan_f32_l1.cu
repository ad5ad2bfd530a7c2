#define HB_CAT_(a, b) a##b
#define HB_CAT(a, b) HB_CAT_(a, b)
#define HBI_T float
#define HBI_IP 2
#define HBI_NAME f32_l1
#include "inst_scan.cuh"
