/* Forced into every reference translation unit by solaris_b200/host/build_dropin.sh (-include; same content as oracle/absfix.h, kept apart so that the product build does not read oracle/).
 * The reference was written for MSVC, where an unqualified abs(double) is fabs.
 * With g++/libstdc++ it silently binds to C's int abs(int), which
 * changes Ephemeris.cpp:30,66,200 and Acceleration.cpp:770.  This restores the
 * MSVC meaning without touching the reference sources. */
#ifndef SOLB200_ABSFIX_H
#define SOLB200_ABSFIX_H
#ifdef __cplusplus
#include <cmath>
#include <cstdlib>
using std::abs;
#endif
#endif
