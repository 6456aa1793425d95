// api.cu — library info entry points of libpatchaug_b200.so.
#include "common.cuh"

int g_pab_launches = 0;

PAB_API int pab_version(void) { return 1; }
PAB_API int pab_num_launches(void) { return g_pab_launches; }
PAB_API void pab_reset_launch_counter(void) { g_pab_launches = 0; }
