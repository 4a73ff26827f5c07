// Library identity for the C ABI (include/npgp.h).
#include "common.cuh"

extern "C" int npgp_version(void) { return 100; }  // 0.1.0

long npgp_launch_counter = 0;
extern "C" long npgp_launch_count(void) { return npgp_launch_counter; }
