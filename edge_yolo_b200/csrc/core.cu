// Library-level entry points of libedgeline_b200.so.
#include <stdlib.h>

#include "el_common.cuh"

namespace el {
thread_local int g_last_cuda_error = 0;
unsigned long long g_launches = 0;
bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("EL_PDL");
        on = !(e && e[0] == '0');
    }
    return on != 0;
}
}

extern "C" const char* el_version(void) { return "edgeline_b200 0.1.0 (sm_100a)"; }

extern "C" const char* el_status_string(int s) {
    switch (s) {
        case EL_OK: return "ok";
        case EL_ERR_ARG: return "invalid argument";
        case EL_ERR_UNSUPPORTED: return "unsupported shape";
        case EL_ERR_WORKSPACE: return "workspace too small";
        case EL_ERR_CUDA: return "CUDA launch failed";
        default: return "unknown status";
    }
}

extern "C" int el_last_cuda_error(void) { return el::g_last_cuda_error; }

extern "C" unsigned long long el_launch_count(void) { return __atomic_load_n(&el::g_launches, __ATOMIC_RELAXED); }
