// inverted_pendulum.cu -- InvertedPendulum-v5 instance of the planar MuJoCo kernels (mjc_planar.cuh).
#define MJC_ROBOT_INVPEND 1
#include "mjc_planar.cuh"
