// ORACLE - test infrastructure only.  A C entry point in front of the REFERENCE's own Permutohedral class
// (utils/bilateralfilter/permutohedral.hpp, compiled from where it lies: oracle/Makefile `ref`), so that a lattice of
// ANY dimension can be driven from Python: the dense-CRF inference of utils/seg_helper.py:961-996 filters over a 2-D
// (x, y) lattice and a 5-D (x, y, R, G, B) one.  Exists only where /root/reference is mounted; its outputs are
// committed as golden vectors (tests/golden/crf_inference.npz).
#include "permutohedral.hpp"

extern "C" int ref_lattice_filter(const float *features, int d, int n, const float *in, float *out, int K) {
  Permutohedral lattice;
  lattice.init(features, d, n);            // features [n][d], point-major (bilateralfilter.cpp:7-17 builds them so)
  for (int k = 0; k < K; ++k) lattice.compute(out + (size_t)k * n, in + (size_t)k * n, 1);
  return 0;
}
