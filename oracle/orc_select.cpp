// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).  PARITY UNPINNED.
//
// orc_select.cpp: CPU restatement of the pixel selector
//   computeHistQuantil          src/FullSystem/PixelSelector2.cpp:59-67
//   PixelSelector::makeHists    src/FullSystem/PixelSelector2.cpp:69-145
//   PixelSelector::makeMaps     src/FullSystem/PixelSelector2.cpp:146-282   (FAST branch and plotting are dead / debug code)
//   PixelSelector::select       src/FullSystem/PixelSelector2.cpp:284-422
// ths / thsSmoothed carry the reference's 100 slack entries, zero-filled here (the reference leaves them uninitialised and
// reads them for pixels right of / below the last full 32x32 block when w or h is not a multiple of 32).
#include <cmath>
#include <cstring>

#include "orc_core.h"
#include "orc_host.h"

namespace orc {

static const float s_minGradHistCut = 0.5f, s_minGradHistAdd = 7, s_gradDownweightPerLevel = 0.75f;   // settings.cpp:122-124

static int computeHistQuantil(const int *hist, float below) {
  int th = hist[0] * below + 0.5f;
  for (int i = 0; i < 90; i++) {
    th -= hist[i + 1];
    if (th < 0) return i;
  }
  return 90;
}

static void makeHists(Oracle &o, int slot) {
  Oracle::Selector &S = o.sel;
  const float *mapmax0 = o.slots[slot].lvl[0].absg.data();
  const int w = o.wl[0], h = o.hl[0], w32 = w / 32, h32 = h / 32;
  S.thsStep = w32;
  S.ths.assign((size_t)w32 * h32 + 100, 0.f);
  S.thsSmoothed.assign((size_t)w32 * h32 + 100, 0.f);
  for (int y = 0; y < h32; y++)
    for (int x = 0; x < w32; x++) {
      const float *map0 = mapmax0 + 32 * x + 32 * y * w;
      int hist0[100];
      memset(hist0, 0, sizeof(hist0));   // the reference clears 50 entries of a larger heap block
      for (int j = 0; j < 32; j++)
        for (int i = 0; i < 32; i++) {
          int it = i + 32 * x, jt = j + 32 * y;
          if (it > w - 2 || jt > h - 2 || it < 1 || jt < 1) continue;
          int g = sqrtf(map0[i + j * w]);
          if (g > 48) g = 48;
          hist0[g + 1]++;
          hist0[0]++;
        }
      S.ths[x + y * w32] = computeHistQuantil(hist0, s_minGradHistCut) + s_minGradHistAdd;
    }
  for (int y = 0; y < h32; y++)
    for (int x = 0; x < w32; x++) {
      float sum = 0, num = 0;
      if (x > 0) {
        if (y > 0) { num++; sum += S.ths[x - 1 + (y - 1) * w32]; }
        if (y < h32 - 1) { num++; sum += S.ths[x - 1 + (y + 1) * w32]; }
        num++; sum += S.ths[x - 1 + y * w32];
      }
      if (x < w32 - 1) {
        if (y > 0) { num++; sum += S.ths[x + 1 + (y - 1) * w32]; }
        if (y < h32 - 1) { num++; sum += S.ths[x + 1 + (y + 1) * w32]; }
        num++; sum += S.ths[x + 1 + y * w32];
      }
      if (y > 0) { num++; sum += S.ths[x + (y - 1) * w32]; }
      if (y < h32 - 1) { num++; sum += S.ths[x + (y + 1) * w32]; }
      num++; sum += S.ths[x + y * w32];
      S.thsSmoothed[x + y * w32] = (sum / num) * (sum / num);
    }
}

static void select(Oracle &o, int slot, float *map_out, int pot, float thFactor, int n[3]) {
  const Oracle::Selector &S = o.sel;
  const Pyramid &P = o.slots[slot];
  const float *map0 = P.lvl[0].dI.data();
  const float *mapmax0 = P.lvl[0].absg.data(), *mapmax1 = P.lvl[1].absg.data(), *mapmax2 = P.lvl[2].absg.data();
  const int w = o.wl[0], w1 = o.wl[1], w2 = o.wl[2], h = o.hl[0];
  static const float directions[16][2] = {{0, 1.0000f},       {0.3827f, 0.9239f},  {0.1951f, 0.9808f},  {0.9239f, 0.3827f},
                                          {0.7071f, 0.7071f}, {0.3827f, -0.9239f}, {0.8315f, 0.5556f},  {0.8315f, -0.5556f},
                                          {0.5556f, -0.8315f}, {0.9808f, 0.1951f}, {0.9239f, -0.3827f}, {0.7071f, -0.7071f},
                                          {0.5556f, 0.8315f}, {0.9808f, -0.1951f}, {1.0000f, 0.0000f},  {0.1951f, -0.9808f}};
  memset(map_out, 0, sizeof(float) * w * h);
  const float dw1 = s_gradDownweightPerLevel, dw2 = dw1 * dw1;
  const uint8_t *rp = S.randomPattern.data();
  int n3 = 0, n2 = 0, n4 = 0;
  for (int y4 = 0; y4 < h; y4 += 4 * pot)
    for (int x4 = 0; x4 < w; x4 += 4 * pot) {
      const int my3 = std::min(4 * pot, h - y4), mx3 = std::min(4 * pot, w - x4);
      int bestIdx4 = -1;
      float bestVal4 = 0;
      const float *dir4 = directions[rp[n2] & 0xF];
      for (int y3 = 0; y3 < my3; y3 += 2 * pot)
        for (int x3 = 0; x3 < mx3; x3 += 2 * pot) {
          const int x34 = x3 + x4, y34 = y3 + y4;
          const int my2 = std::min(2 * pot, h - y34), mx2 = std::min(2 * pot, w - x34);
          int bestIdx3 = -1;
          float bestVal3 = 0;
          const float *dir3 = directions[rp[n2] & 0xF];
          for (int y2 = 0; y2 < my2; y2 += pot)
            for (int x2 = 0; x2 < mx2; x2 += pot) {
              const int x234 = x2 + x34, y234 = y2 + y34;
              const int my1 = std::min(pot, h - y234), mx1 = std::min(pot, w - x234);
              int bestIdx2 = -1;
              float bestVal2 = 0;
              const float *dir2 = directions[rp[n2] & 0xF];
              for (int y1 = 0; y1 < my1; y1++)
                for (int x1 = 0; x1 < mx1; x1++) {
                  const int xf = x1 + x234, yf = y1 + y234, idx = xf + w * yf;
                  if (xf < 4 || xf >= w - 5 || yf < 4 || yf > h - 4) continue;
                  const float pixelTH0 = S.thsSmoothed[(xf >> 5) + (yf >> 5) * S.thsStep];
                  const float pixelTH1 = pixelTH0 * dw1, pixelTH2 = pixelTH1 * dw2;
                  const float gx = map0[3 * idx + 1], gy = map0[3 * idx + 2];
                  const float ag0 = mapmax0[idx];
                  if (ag0 > pixelTH0 * thFactor) {
                    const float dirNorm = fabsf((float)(gx * dir2[0] + gy * dir2[1]));
                    if (dirNorm > bestVal2) { bestVal2 = dirNorm; bestIdx2 = idx; bestIdx3 = -2; bestIdx4 = -2; }
                  }
                  if (bestIdx3 == -2) continue;
                  const float ag1 = mapmax1[(int)(xf * 0.5f + 0.25f) + (int)(yf * 0.5f + 0.25f) * w1];
                  if (ag1 > pixelTH1 * thFactor) {
                    const float dirNorm = fabsf((float)(gx * dir3[0] + gy * dir3[1]));
                    if (dirNorm > bestVal3) { bestVal3 = dirNorm; bestIdx3 = idx; bestIdx4 = -2; }
                  }
                  if (bestIdx4 == -2) continue;
                  const float ag2 = mapmax2[(int)(xf * 0.25f + 0.125) + (int)(yf * 0.25f + 0.125) * w2];
                  if (ag2 > pixelTH2 * thFactor) {
                    const float dirNorm = fabsf((float)(gx * dir4[0] + gy * dir4[1]));
                    if (dirNorm > bestVal4) { bestVal4 = dirNorm; bestIdx4 = idx; }
                  }
                }
              if (bestIdx2 > 0) { map_out[bestIdx2] = 1; bestVal3 = 1e10; n2++; }
            }
          if (bestIdx3 > 0) { map_out[bestIdx3] = 2; bestVal4 = 1e10; n3++; }
        }
      if (bestIdx4 > 0) { map_out[bestIdx4] = 4; n4++; }
    }
  n[0] = n2; n[1] = n3; n[2] = n4;
}

// PixelSelector::makeMaps (PixelSelector2.cpp:146-282)
static int makeMaps(Oracle &o, int slot, float *map_out, float density, int recursionsLeft, float thFactor) {
  Oracle::Selector &S = o.sel;
  float numHave = 0, numWant = density, quotia;
  int idealPotential = S.currentPotential;
  {
    int n[3];
    select(o, slot, map_out, S.currentPotential, thFactor, n);
    numHave = n[0] + n[1] + n[2];
    quotia = numWant / numHave;
    float K = numHave * (S.currentPotential + 1) * (S.currentPotential + 1);
    idealPotential = sqrtf(K / numWant) - 1;
    if (idealPotential < 1) idealPotential = 1;
    if (recursionsLeft > 0 && quotia > 1.25 && S.currentPotential > 1) {
      if (idealPotential >= S.currentPotential) idealPotential = S.currentPotential - 1;
      S.currentPotential = idealPotential;
      return makeMaps(o, slot, map_out, density, recursionsLeft - 1, thFactor);
    } else if (recursionsLeft > 0 && quotia < 0.25) {
      if (idealPotential <= S.currentPotential) idealPotential = S.currentPotential + 1;
      S.currentPotential = idealPotential;
      return makeMaps(o, slot, map_out, density, recursionsLeft - 1, thFactor);
    }
  }
  int numHaveSub = numHave;
  if (quotia < 0.95) {
    const int wh = o.wl[0] * o.hl[0];
    int rn = 0;
    unsigned char charTH = 255 * quotia;
    for (int i = 0; i < wh; i++)
      if (map_out[i] != 0) {
        if (S.randomPattern[rn] > charTH) { map_out[i] = 0; numHaveSub--; }
        rn++;
      }
  }
  S.currentPotential = idealPotential;
  return numHaveSub;
}

int pixel_select(Oracle &o, int slot, float density, int recursionsLeft, float thFactor, float *map_out) {
  makeHists(o, slot);   // `if (fh != gradHistFrame) makeHists(fh)`: once per frame, reused by the recursion
  return makeMaps(o, slot, map_out, density, recursionsLeft, thFactor);
}

}  // namespace orc
