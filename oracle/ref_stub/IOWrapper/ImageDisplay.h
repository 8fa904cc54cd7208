// ORACLE — TEST INFRASTRUCTURE ONLY.  util/globalFuncs.h includes the display wrapper (OpenCV windows) only for the Vec3b
// colour type of its rainbow helpers; this stand-in provides that type and nothing else.
#pragma once
#include "util/NumType.h"
namespace dso {
typedef Eigen::Matrix<unsigned char, 3, 1> Vec3b;
}
