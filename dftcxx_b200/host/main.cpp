// dftcxx -i <molecule.in> : drop-in command line of the reference (src/dftcxx.cpp) on the B200 grid engine.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>

#include "dft.hpp"

static const char* kVersion = "1.1.2-b200";

int main(int argc, char** argv) {
    std::string input;
    int device = 0, ngpus = -1, scf_mode = -1;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if ((a == "-i" || a == "--input") && i + 1 < argc) {
            input = argv[++i];
        } else if (a == "--device" && i + 1 < argc) {
            device = std::atoi(argv[++i]);
        } else if (a == "--gpus" && i + 1 < argc) {
            ngpus = std::atoi(argv[++i]);  // devices device .. device+N-1 of this box, one process (overrides `gpus =` in the input)
            if (ngpus < 1) {
                std::cerr << "error: --gpus needs a positive integer" << std::endl;
                return -1;
            }
        } else if (a == "--scf" && i + 1 < argc) {
            const std::string v = argv[++i];  // device (default) | host | host-separate
            scf_mode = v == "host" ? 1 : (v == "host-separate" ? 2 : 0);
        } else if (a == "--version") {
            std::cout << argv[0] << "  version: " << kVersion << std::endl;
            return 0;
        } else if (a == "-h" || a == "--help") {
            std::cout << "USAGE: " << argv[0] << " -i <filename> [--device N] [--gpus N] [--scf device|host|host-separate]\n\nPerform DFT calculation.\n";
            return 0;
        } else {
            std::cerr << "error: Couldn't find match for argument for arg " << a << std::endl;
            return -1;
        }
    }
    if (input.empty()) {
        std::cerr << "error: Required argument missing for arg input" << std::endl;
        return -1;
    }
    try {
        std::cout << "--------------------------------------------------------------" << std::endl << std::endl;
        std::cout << "Executing DFTCXX v." << kVersion << std::endl;
        std::cout << "Author: Ivo Filot <ivo@ivofilot.nl> (reference program); B200 grid engine: this repository" << std::endl << std::endl;
        std::cout << "--------------------------------------------------------------" << std::endl << std::endl;
        const auto t0 = std::chrono::system_clock::now();
        dftcxx::DFT dft(input, device, true, ngpus, scf_mode);
        dft.scf();
        const auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now() - t0).count();
        std::printf("Total elapsed time: %ld ms\n", (long)ms);
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return -1;
    }
}
