#define HB_CAT_(a, b) a##b
#define HB_CAT(a, b) HB_CAT_(a, b)
#define HBI_T float
#define HBI_IP 1
#define HBI_NAME f32_ip
#include "inst_scan.cuh"
