// TEST INFRASTRUCTURE ONLY — stand-in for InteractiveComputerGraphics/Discregrid
// (@0b69062ff9c56fbb6dcecd296652028bedbacf0e, CMake/SetUpExternalProjects.cmake:13-20; not vendored).  Only the type
// names are needed: the density/volume-map boundary models (Koschier2017, Bender2019) that use it are never
// instantiated on the DiffDFSPH path (scenes use boundaryHandlingMethod 0 = Akinci2012).
#pragma once
#include <Eigen/Dense>
#include <array>
#include <functional>
#include <string>
namespace Discregrid {
class DiscreteGrid {
 public:
  using CoefficientVector = Eigen::Matrix<double, 32, 1>;
  using ContinuousFunction = std::function<double(Eigen::Vector3d const &)>;
  virtual ~DiscreteGrid() {}
};
class CubicLagrangeDiscreteGrid : public DiscreteGrid {
 public:
  CubicLagrangeDiscreteGrid() {}
  CubicLagrangeDiscreteGrid(std::string const &) {}
  CubicLagrangeDiscreteGrid(Eigen::AlignedBox3d const &, std::array<unsigned int, 3> const &) {}
};
}  // namespace Discregrid
