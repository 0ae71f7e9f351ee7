// Library-level entry points of libpixart_sm100.so (version, error string, launch counter).
#include "host_common.cuh"

extern "C" int pxa_version(void) { return 100; }  // 0.1.0
extern "C" const char* pxa_last_error(void) { return pxa::last_error_buf(); }
extern "C" uint64_t pxa_launch_count(void) { return pxa::launch_counter().load(); }
