// ORACLE — TEST INFRASTRUCTURE ONLY.  util/NumType.h only names these types in typedefs; the units compiled against this
// stub (accumulators, samplers, settings) never touch a pose.
#pragma once
#include "Eigen/Core"
namespace Sophus {
struct SO3d {};
struct SE3d {};
struct Sim3d {};
}  // namespace Sophus
