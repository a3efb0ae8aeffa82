// K3/K4 placeholder (implemented next).
#include "ssn_launch.h"
namespace ssn {
int launch_euler_forward(const ssn_solver &, int, int, int, const float *, const ssn_jds &, const float *, int, int,
                         int, double, float *, double *, float *, float *, int *, cudaStream_t) {
    set_error("euler forward: not built yet");
    return -1;
}
int launch_euler_backward(const ssn_solver &, int, int, int, const float *, const ssn_jds &, int, int, double,
                          const float *, double, double, const float *, const float *, float *, double *, int *,
                          cudaStream_t) {
    set_error("euler backward: not built yet");
    return -1;
}
}
