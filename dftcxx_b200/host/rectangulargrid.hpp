// RectangularGrid: the reference's density-dump grid (src/rectangulargrid.h:28-67, src/rectangulargrid.cpp:24-95) as a view of
// the B200 engine.  build_grid fixes the box, set_density evaluates density and density gradient on the device
// (dftgrid_rectangular_density), write_gradient writes the reference's text format: x y z grad_x grad_y grad_z per point.
#pragma once
#include <string>
#include <vector>

#include "linalg.hpp"
#include "moleculargrid.hpp"

namespace dftcxx {

class RectangularGrid {
public:
    explicit RectangularGrid(MolecularGrid& engine) : engine(engine) {}
    void build_grid(double size, unsigned int dp);  // box edge (same units as the atomic positions), points per direction
    void set_density(const Mat& P);
    void write_gradient(const std::string& filename) const;

    size_t size() const { return rho.size(); }
    const std::vector<double>& positions() const { return pos; }   // [n][3]
    const std::vector<double>& densities() const { return rho; }   // [n]
    const std::vector<double>& gradients() const { return grad; }  // [n][3]

private:
    MolecularGrid& engine;
    double box = 0.0;
    unsigned int gridsize = 0;
    std::vector<double> pos, rho, grad;
};

}  // namespace dftcxx
