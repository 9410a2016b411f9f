// TEST INFRASTRUCTURE ONLY — stand-in for InteractiveComputerGraphics/Discregrid
// (@0b69062ff9c56fbb6dcecd296652028bedbacf0e, CMake/SetUpExternalProjects.cmake:13-20; not vendored).  Only the type
// names are needed: the density/volume-map boundary models (Koschier2017, Bender2019) that use it are never
// instantiated on the DiffDFSPH path (scenes use boundaryHandlingMethod 0 = Akinci2012).
#pragma once
#include <Eigen/Dense>
#include <array>
#include <functional>
#include <cstdio>
#include <cstdlib>
#include <string>
namespace Discregrid {
class DiscreteGrid {
 public:
  using CoefficientVector = Eigen::Matrix<double, 32, 1>;
  using ContinuousFunction = std::function<double(Eigen::Vector3d const &)>;
  virtual ~DiscreteGrid() {}
  // declared so that the density/volume-map code paths of TimeStep.cpp compile; never reached with Akinci2012
  double interpolate(unsigned int, Eigen::Vector3d const &, Eigen::Vector3d * = nullptr) const { unreachable(); return 0.0; }
  bool determineShapeFunctions(unsigned int, Eigen::Vector3d const &, std::array<unsigned int, 32> &, Eigen::Vector3d &,
                               Eigen::Matrix<double, 32, 1> &, Eigen::Matrix<double, 32, 3> * = nullptr) const { unreachable(); return false; }
  double interpolate(unsigned int, Eigen::Vector3d const &, const std::array<unsigned int, 32> &, const Eigen::Vector3d &,
                     const Eigen::Matrix<double, 32, 1> &, Eigen::Vector3d * = nullptr, Eigen::Matrix<double, 32, 3> * = nullptr) const { unreachable(); return 0.0; }
 private:
  static void unreachable() { std::fprintf(stderr, "oracle/_ref: Discregrid shim called (boundaryHandlingMethod != Akinci2012)\n"); std::abort(); }
};
class CubicLagrangeDiscreteGrid : public DiscreteGrid {
 public:
  CubicLagrangeDiscreteGrid() {}
  CubicLagrangeDiscreteGrid(std::string const &) {}
  CubicLagrangeDiscreteGrid(Eigen::AlignedBox3d const &, std::array<unsigned int, 3> const &) {}
};
}  // namespace Discregrid
