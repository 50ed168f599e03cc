// walker2d.cu -- Walker2d-v5 instance of the planar MuJoCo kernels (mjc_planar.cuh).
#define MJC_ROBOT_WALKER2D 1
#include "mjc_planar.cuh"
