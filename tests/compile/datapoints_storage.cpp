// DataPoints storage semantics of include/laser_slam_compat/compat.hpp: copies share storage until written
// (copy-on-write), views borrow caller memory and copy on the first write.  Compiled and run by tests/test_abi.py;
// needs no GPU (nothing here reaches the device library).
#include <cstdio>
#include <vector>

#include "laser_slam_compat/compat.hpp"

typedef PointMatcher<float> PM;

#define EXPECT(c) do { if (!(c)) { std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
  const size_t n = 5;
  std::vector<float> f(4 * n), nr(3 * n);
  for (size_t i = 0; i < f.size(); ++i) f[i] = (float)i;
  for (size_t i = 0; i < nr.size(); ++i) nr[i] = 100.f + (float)i;

  PM::DataPoints a = PM::DataPoints::fromArrays(f.data(), nr.data(), n);
  EXPECT(a.features.rows() == 4 && a.features.cols() == n && a.descriptors.rows() == 3);
  const PM::DataPoints& ca = a;
  EXPECT(ca.features.data() != f.data());  // an owning copy
  PM::DataPoints b = a;                    // shares
  const PM::DataPoints& cb = b;
  EXPECT(cb.features.data() == ca.features.data());
  b.features(0, 0) = -1.f;                 // first write detaches b
  EXPECT(cb.features.data() != ca.features.data());
  EXPECT(ca.features(0, 0) == 0.f && cb.features(0, 0) == -1.f && cb.features(1, 0) == 1.f);

  PM::DataPoints v = PM::DataPoints::viewOfArrays(f.data(), nr.data(), n);
  const PM::DataPoints& cv = v;
  EXPECT(cv.features.data() == f.data() && cv.descriptors.data() == nr.data());  // borrowed
  EXPECT(cv.getNbPoints() == n && cv.descriptorExists("normals") && cv.descriptors(2, 4) == nr[14]);
  PM::DataPoints w = v;
  w.descriptors(0, 0) = 7.f;  // copies, the caller's array is untouched
  EXPECT(nr[0] == 100.f && static_cast<const PM::DataPoints&>(w).descriptors(0, 0) == 7.f);
  EXPECT(cv.descriptors.data() == nr.data());

  PM::DataPoints c = v;
  c.concatenate(a);
  EXPECT(c.getNbPoints() == 2 * n && static_cast<const PM::DataPoints&>(c).features(3, n + 1) == f[7]);
  EXPECT(cv.getNbPoints() == n);
  std::printf("ok\n");
  return 0;
}
