#include "rectangulargrid.hpp"

#include <cstdio>
#include <stdexcept>

namespace dftcxx {

void RectangularGrid::build_grid(double size_, unsigned int dp) {
    if (dp < 2 || !(size_ > 0.0)) throw std::runtime_error("RectangularGrid: need size > 0 and at least 2 points per direction");
    box = size_;
    gridsize = dp;
    const size_t n = (size_t)dp * dp * dp;
    pos.assign(3 * n, 0.0);
    rho.assign(n, 0.0);
    grad.assign(3 * n, 0.0);
}

void RectangularGrid::set_density(const Mat& P) {
    if (gridsize == 0) throw std::runtime_error("RectangularGrid: build_grid has not been called");
    engine.rectangular_density(box, gridsize, P, pos.data(), rho.data(), grad.data());
}

// "%12.8f  %12.8f  %12.8f  %12.8f  %12.8f  %12.8f\n" per point (src/rectangulargrid.cpp:82-95)
void RectangularGrid::write_gradient(const std::string& filename) const {
    std::FILE* f = std::fopen(filename.c_str(), "w");
    if (!f) throw std::runtime_error("Cannot open " + filename + " for writing");
    for (size_t i = 0; i < rho.size(); i++)
        std::fprintf(f, "%12.8f  %12.8f  %12.8f  %12.8f  %12.8f  %12.8f\n", pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], grad[3 * i], grad[3 * i + 1],
                     grad[3 * i + 2]);
    if (std::fclose(f) != 0) throw std::runtime_error("error while writing " + filename);
}

}  // namespace dftcxx
