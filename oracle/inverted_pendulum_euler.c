/* oracle/inverted_pendulum_euler.c -- TEST-ONLY instance of the planar core: inverted_pendulum.xml integrated with mj_Euler
 * (implicit joint damping) instead of the RK4 its XML asks for.  No reference env uses it; it exists so that the Euler path,
 * which HalfCheetah-v5 runs, can be checked against the cart-pole equations of motion (tests/test_oracle_hopper.py). */
#define ROBOT_INVPEND 1
#define OPT_EULER 1
#include "mjc_planar.h"
