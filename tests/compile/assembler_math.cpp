// Prints VelodyneAssembler::multiply(A, B) and rigidInverse(A) for the two column-major 4x4 float matrices given as 32
// hexadecimal bit patterns on the command line; tests/test_filters.py compares the bits with the oracle's restatement.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "laser_slam/velodyne_assembler.hpp"

int main(int argc, char** argv) {
  if (argc != 33) return 2;
  laser_slam::VelodyneAssembler::Matrix4 A, B;
  for (int i = 0; i < 32; ++i) {
    const unsigned int u = (unsigned int)std::strtoul(argv[1 + i], NULL, 16);
    float f;
    std::memcpy(&f, &u, 4);
    (i < 16 ? A : B).data()[i % 16] = f;
  }
  const laser_slam::VelodyneAssembler::Matrix4 C = laser_slam::VelodyneAssembler::multiply(A, B);
  const laser_slam::VelodyneAssembler::Matrix4 I = laser_slam::VelodyneAssembler::rigidInverse(A);
  for (int pass = 0; pass < 2; ++pass) {
    const float* d = pass == 0 ? C.data() : I.data();
    for (int i = 0; i < 16; ++i) {
      unsigned int u;
      std::memcpy(&u, d + i, 4);
      std::printf("%08x%c", u, i == 15 ? '\n' : ' ');
    }
  }
  return 0;
}
