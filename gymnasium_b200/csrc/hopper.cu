// hopper.cu -- Hopper-v5 instance of the planar MuJoCo kernels (mjc_planar.cuh).
#define MJC_ROBOT_HOPPER 1
#include "mjc_planar.cuh"
