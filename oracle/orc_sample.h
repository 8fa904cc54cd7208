// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).
// Restatement of the bilinear samplers of src/util/globalFuncs.h on the Vector3f image {I, dx, dy} (3 floats per texel):
//   getInterpolatedElement33       :68-82     getInterpolatedElement31   :122-136    getInterpolatedElement33BiLin :161-182
// PINNED: tests/test_ref_pin.py compares them bit for bit with the reference functions themselves, compiled unmodified from
// /root/reference/src/util/globalFuncs.h (oracle/ref_build.sh -> oracle/_ref/libref_units.so).
#pragma once

namespace orc {

inline void interp33(const float *mat, float x, float y, int width, float out[3]) {   // globalFuncs.h:68-82
  int ix = (int)x, iy = (int)y;
  float dx = x - ix, dy = y - iy;
  float dxdy = dx * dy;
  const float *bp = mat + 3 * (ix + iy * width);
  const float w11 = dxdy, w01 = dy - dxdy, w10 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
  for (int c = 0; c < 3; c++)
    out[c] = w11 * bp[3 * (1 + width) + c] + w01 * bp[3 * width + c] + w10 * bp[3 + c] + w00 * bp[c];
}
inline float interp31(const float *mat, float x, float y, int width) {   // globalFuncs.h:122-136 (channel 0 of Vector3f)
  int ix = (int)x, iy = (int)y;
  float dx = x - ix, dy = y - iy;
  float dxdy = dx * dy;
  const float *bp = mat + 3 * (ix + iy * width);
  return dxdy * bp[3 * (1 + width)] + (dy - dxdy) * bp[3 * width] + (dx - dxdy) * bp[3] + (1 - dx - dy + dxdy) * bp[0];
}
inline void interp33BiLin(const float *mat, float x, float y, int width, float out[3]) {   // globalFuncs.h:161-182
  int ix = (int)x, iy = (int)y;
  const float *bp = mat + 3 * (ix + iy * width);
  float tl = bp[0], tr = bp[3], bl = bp[3 * width], br = bp[3 * (width + 1)];
  float dx = x - ix, dy = y - iy;
  float topInt = dx * tr + (1 - dx) * tl;
  float botInt = dx * br + (1 - dx) * bl;
  float leftInt = dy * bl + (1 - dy) * tl;
  float rightInt = dy * br + (1 - dy) * tr;
  out[0] = dx * rightInt + (1 - dx) * leftInt; out[1] = rightInt - leftInt; out[2] = botInt - topInt;
}

}  // namespace orc
