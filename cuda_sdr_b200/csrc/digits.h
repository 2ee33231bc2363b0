// 24-bit fixed point as three balanced int8 digits: the representation of the tap matrices B of the int8 tensor-core routes
// (chainKernel CONV = imma, toepKernel, channelKernel).  v * scale = d2 * 65536 + d1 * 256 + d0 with every digit in
// [-128, 127], so  X * B = sum_d (X * D_d) * digitScale[d]  is three exact int8 x int8 -> int32 MMAs per k-step.
#pragma once

#include <cmath>

namespace b200sdr {

struct FixedPoint24 {
  double scale;         // multiply a matrix entry by this before rounding to an integer
  float digitScale[3];  // weight of digit d in the reconstructed value: 256^d / scale
};

// `largest` = max |entry| of the matrix: it maps to 127 * 65536 + 127 * 256 + 127 = 8 355 711, the largest value three
// balanced digits reach on the positive side
inline FixedPoint24 fixedPoint24For(double largest) {
  FixedPoint24 f;
  f.scale = largest > 0.0 ? 8355711.0 / largest : 1.0;
  f.digitScale[0] = static_cast<float>(1.0 / f.scale);
  f.digitScale[1] = static_cast<float>(256.0 / f.scale);
  f.digitScale[2] = static_cast<float>(65536.0 / f.scale);
  return f;
}

inline void balancedDigits(double value, double scale, int out[3]) {
  long long q = std::llround(value * scale);
  for (int d = 0; d < 3; d++) {
    long long r = ((q % 256) + 256) % 256;
    if (r >= 128) r -= 256;
    out[d] = static_cast<int>(r);
    q = (q - r) / 256;
  }
}

}  // namespace b200sdr
