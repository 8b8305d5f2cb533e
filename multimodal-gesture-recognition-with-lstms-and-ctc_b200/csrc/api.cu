// Library-level entry points of include/gr_b200.h.
#include "common.cuh"

namespace gr {
thread_local char g_last_error[256] = {0};
}

extern "C" int gr_version(void) { return 100; }
extern "C" const char* gr_last_error(void) { return gr::g_last_error; }
