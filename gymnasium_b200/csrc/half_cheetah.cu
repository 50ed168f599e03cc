// half_cheetah.cu -- HalfCheetah-v5 instance of the planar MuJoCo kernels (mjc_planar.cuh).
#define MJC_ROBOT_HALFCHEETAH 1
#include "mjc_planar.cuh"
