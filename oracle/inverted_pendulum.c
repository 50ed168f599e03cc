/* oracle/inverted_pendulum.c -- InvertedPendulum-v5 instance of the planar MuJoCo oracle core (see mjc_planar.h; test
 * infrastructure only). */
#define ROBOT_INVPEND 1
#include "mjc_planar.h"
