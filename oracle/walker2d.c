/* oracle/walker2d.c -- Walker2d-v5 instance of the planar MuJoCo oracle core (see mjc_planar.h; test infrastructure only). */
#define ROBOT_WALKER2D 1
#include "mjc_planar.h"
