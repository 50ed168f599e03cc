/* oracle/half_cheetah.c -- HalfCheetah-v5 instance of the planar MuJoCo oracle core (see mjc_planar.h; test infrastructure
 * only). */
#define ROBOT_HALFCHEETAH 1
#include "mjc_planar.h"
