// K2 placeholder (implemented next).
#include "ssn_launch.h"
namespace ssn {
int launch_ift_gradient(const ssn_solver &, int, int, int, const float *, const ssn_jds &, const float *, int,
                        const float *, const float *, double, double *, float *, int *, int *, int *, cudaStream_t) {
    set_error("ift gradient: not built yet");
    return -1;
}
}
