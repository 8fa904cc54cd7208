// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).  PARITY UNPINNED.
//
// orc_trace.cpp: CPU restatement of the immature-point row:
//   ImmaturePoint::ImmaturePoint       src/FullSystem/ImmaturePoint.cpp:28-60
//   ImmaturePoint::traceOn             src/FullSystem/ImmaturePoint.cpp:70-415
//   FullSystem::traceNewCoarse         src/FullSystem/FullSystem.cpp:311-361
//   getInterpolatedElement31 / 33 / 33BiLin   src/util/globalFuncs.h:122-136, 68-82, 161-182
// Eigen expressions are written out in the coefficient order Eigen uses for fixed-size products
// ((a0 x + a1 y) + a2 z; v^T G evaluated before (v^T G) v).
#include <cmath>
#include <cstring>
#include <vector>

#include "orc_core.h"
#include "orc_sample.h"

namespace orc {

static const float s_maxPixSearch = 0.027f, s_trace_stepsize = 1.0f, s_trace_GNThreshold = 0.1f, s_trace_extraSlackOnTH = 1.2f,
                   s_trace_slackInterval = 1.5f, s_trace_minImprovementFactor = 2.f, s_outlierTH = 12 * 12;   // settings.cpp:82,128-143
static const int s_trace_GNIterations = 3, s_minTraceTestRadius = 2;

// ImmaturePoint::ImmaturePoint (ImmaturePoint.cpp:28-60)
void immature_init(Oracle &o, int slot, int n, const int32_t *u, const int32_t *v, float *color, float *weights, float *gradH, float *energyTH) {
  const float *dI = o.slots[slot].lvl[0].dI.data();
  const int w = o.wl[0];
  const float c = o.cfg.outlier_th_sum_component;
  for (int p = 0; p < n; p++) {
    float G[4] = {0, 0, 0, 0};
    bool bad = false;
    for (int idx = 0; idx < 8 && !bad; idx++) {
      float ptc[3];
      interp33BiLin(dI, (float)(u[p] + patternP[idx][0]), (float)(v[p] + patternP[idx][1]), w, ptc);
      color[8 * p + idx] = ptc[0];
      if (!std::isfinite(ptc[0])) { bad = true; break; }
      G[0] += ptc[1] * ptc[1]; G[1] += ptc[1] * ptc[2]; G[2] += ptc[2] * ptc[1]; G[3] += ptc[2] * ptc[2];
      weights[8 * p + idx] = sqrtf(c / (c + (ptc[1] * ptc[1] + ptc[2] * ptc[2])));
    }
    for (int i = 0; i < 4; i++) gradH[4 * p + i] = G[i];
    if (bad) { energyTH[p] = NAN; continue; }
    float th = 8 * s_outlierTH;
    th *= o.cfg.overall_energy_th_weight * o.cfg.overall_energy_th_weight;
    energyTH[p] = th;
  }
}

struct IP {   // one point's view into the SoA
  float u, v, idepth_min, idepth_max, quality, energyTH;
  const float *color, *weights, *gradH;
  int status;
  float uv[2], pixint;
};

// ImmaturePoint::traceOn (ImmaturePoint.cpp:70-415)
static int traceOn(const Oracle &o, IP &p, const float *dI, const float *KRKi, const float *Kt, const float *aff) {
  if (p.status == SOSBA_IPS_OOB) return p.status;
  const int wG = o.wl[0], hG = o.hl[0];
  const float huberTH = o.cfg.huber_th;
  float maxPixSearch = (wG + hG) * s_maxPixSearch;
  float pr[3];
  for (int i = 0; i < 3; i++) pr[i] = (KRKi[3 * i] * p.u + KRKi[3 * i + 1] * p.v) + KRKi[3 * i + 2] * 1.0f;
  float ptpMin[3];
  for (int i = 0; i < 3; i++) ptpMin[i] = pr[i] + Kt[i] * p.idepth_min;
  float uMin = ptpMin[0] / ptpMin[2], vMin = ptpMin[1] / ptpMin[2];
  auto oob = [&](int st) { p.uv[0] = p.uv[1] = -1; p.pixint = 0; return p.status = st; };
  if (!(uMin > 4 && vMin > 4 && uMin < wG - 5 && vMin < hG - 5)) return oob(SOSBA_IPS_OOB);

  float dist, uMax, vMax, ptpMax[3];
  if (std::isfinite(p.idepth_max)) {
    for (int i = 0; i < 3; i++) ptpMax[i] = pr[i] + Kt[i] * p.idepth_max;
    uMax = ptpMax[0] / ptpMax[2]; vMax = ptpMax[1] / ptpMax[2];
    if (!(uMax > 4 && vMax > 4 && uMax < wG - 5 && vMax < hG - 5)) return oob(SOSBA_IPS_OOB);
    dist = (uMin - uMax) * (uMin - uMax) + (vMin - vMax) * (vMin - vMax);
    dist = sqrtf(dist);
    if (dist < s_trace_slackInterval) {
      p.uv[0] = (uMax + uMin) * 0.5f; p.uv[1] = (vMax + vMin) * 0.5f;
      p.pixint = dist;
      return p.status = SOSBA_IPS_SKIPPED;
    }
  } else {
    dist = maxPixSearch;
    for (int i = 0; i < 3; i++) ptpMax[i] = pr[i] + Kt[i] * 0.01f;
    uMax = ptpMax[0] / ptpMax[2]; vMax = ptpMax[1] / ptpMax[2];
    float dx = uMax - uMin, dy = vMax - vMin;
    float d = 1.0f / sqrtf(dx * dx + dy * dy);
    uMax = uMin + dist * dx * d;
    vMax = vMin + dist * dy * d;
    if (!(uMax > 4 && vMax > 4 && uMax < wG - 5 && vMax < hG - 5)) return oob(SOSBA_IPS_OOB);
  }
  if (!(p.idepth_min < 0 || (ptpMin[2] > 0.75f && ptpMin[2] < 1.5f))) return oob(SOSBA_IPS_OOB);

  float dx = s_trace_stepsize * (uMax - uMin), dy = s_trace_stepsize * (vMax - vMin);
  const float *G = p.gradH;
  float a = (dx * G[0] + dy * G[2]) * dx + (dx * G[1] + dy * G[3]) * dy;            // (v^T G) v, v = (dx, dy)
  float b = (dy * G[0] + (-dx) * G[2]) * dy + (dy * G[1] + (-dx) * G[3]) * (-dx);   // v = (dy, -dx)
  float errorInPixel = 0.2f + 0.2f * (a + b) / a;
  if (errorInPixel * s_trace_minImprovementFactor > dist && std::isfinite(p.idepth_max)) {
    p.uv[0] = (uMax + uMin) * 0.5f; p.uv[1] = (vMax + vMin) * 0.5f;
    p.pixint = dist;
    return p.status = SOSBA_IPS_BADCONDITION;
  }
  if (errorInPixel > 10) errorInPixel = 10;

  dx /= dist; dy /= dist;
  if (dist > maxPixSearch) { uMax = uMin + maxPixSearch * dx; vMax = vMin + maxPixSearch * dy; dist = maxPixSearch; }
  int numSteps = 1.9999f + dist / s_trace_stepsize;
  const float Rp[4] = {KRKi[0], KRKi[1], KRKi[3], KRKi[4]};
  float randShift = uMin * 1000 - floorf(uMin * 1000);
  float ptx = uMin - randShift * dx, pty = vMin - randShift * dy;
  float rot[8][2];
  for (int idx = 0; idx < 8; idx++) {
    rot[idx][0] = Rp[0] * patternP[idx][0] + Rp[1] * patternP[idx][1];
    rot[idx][1] = Rp[2] * patternP[idx][0] + Rp[3] * patternP[idx][1];
  }
  if (!std::isfinite(dx) || !std::isfinite(dy)) return oob(SOSBA_IPS_OOB);

  float errors[100];
  float bestU = 0, bestV = 0, bestEnergy = 1e10;
  int bestIdx = -1;
  if (numSteps >= 100) numSteps = 99;
  for (int i = 0; i < numSteps; i++) {
    float energy = 0;
    for (int idx = 0; idx < 8; idx++) {
      float hit = interp31(dI, (float)(ptx + rot[idx][0]), (float)(pty + rot[idx][1]), wG);
      if (!std::isfinite(hit)) { energy += 1e5; continue; }
      float residual = hit - (float)(aff[0] * p.color[idx] + aff[1]);
      float hw = fabsf(residual) < huberTH ? 1 : huberTH / fabsf(residual);
      energy += hw * residual * residual * (2 - hw);
    }
    errors[i] = energy;
    if (energy < bestEnergy) { bestU = ptx; bestV = pty; bestEnergy = energy; bestIdx = i; }
    ptx += dx; pty += dy;
  }
  float secondBest = 1e10;
  for (int i = 0; i < numSteps; i++)
    if ((i < bestIdx - s_minTraceTestRadius || i > bestIdx + s_minTraceTestRadius) && errors[i] < secondBest) secondBest = errors[i];
  float newQuality = secondBest / bestEnergy;
  if (newQuality < p.quality || numSteps > 10) p.quality = newQuality;

  float uBak = bestU, vBak = bestV, gnstepsize = 1, stepBack = 0;
  if (s_trace_GNIterations > 0) bestEnergy = 1e5;
  for (int it = 0; it < s_trace_GNIterations; it++) {
    float H = 1, bb = 0, energy = 0;
    for (int idx = 0; idx < 8; idx++) {
      float hit[3];
      interp33(dI, (float)(bestU + rot[idx][0]), (float)(bestV + rot[idx][1]), wG, hit);
      if (!std::isfinite(hit[0])) { energy += 1e5; continue; }
      float residual = hit[0] - (aff[0] * p.color[idx] + aff[1]);
      float dResdDist = dx * hit[1] + dy * hit[2];
      float hw = fabsf(residual) < huberTH ? 1 : huberTH / fabsf(residual);
      H += hw * dResdDist * dResdDist;
      bb += hw * residual * dResdDist;
      energy += p.weights[idx] * p.weights[idx] * hw * residual * residual * (2 - hw);
    }
    if (energy > bestEnergy) {
      stepBack *= 0.5f;
      bestU = uBak + stepBack * dx;
      bestV = vBak + stepBack * dy;
    } else {
      float step = -gnstepsize * bb / H;
      if (step < -0.5f) step = -0.5f;
      else if (step > 0.5f) step = 0.5f;
      if (!std::isfinite(step)) step = 0;
      uBak = bestU; vBak = bestV; stepBack = step;
      bestU += step * dx; bestV += step * dy;
      bestEnergy = energy;
    }
    if (fabsf(stepBack) < s_trace_GNThreshold) break;
  }

  if (!(bestEnergy < p.energyTH * s_trace_extraSlackOnTH)) {
    p.pixint = 0; p.uv[0] = p.uv[1] = -1;
    if (p.status == SOSBA_IPS_OUTLIER) return p.status = SOSBA_IPS_OOB;
    return p.status = SOSBA_IPS_OUTLIER;
  }

  if (dx * dx > dy * dy) {
    p.idepth_min = (pr[2] * (bestU - errorInPixel * dx) - pr[0]) / (Kt[0] - Kt[2] * (bestU - errorInPixel * dx));
    p.idepth_max = (pr[2] * (bestU + errorInPixel * dx) - pr[0]) / (Kt[0] - Kt[2] * (bestU + errorInPixel * dx));
  } else {
    p.idepth_min = (pr[2] * (bestV - errorInPixel * dy) - pr[1]) / (Kt[1] - Kt[2] * (bestV - errorInPixel * dy));
    p.idepth_max = (pr[2] * (bestV + errorInPixel * dy) - pr[1]) / (Kt[1] - Kt[2] * (bestV + errorInPixel * dy));
  }
  if (p.idepth_min > p.idepth_max) { float t = p.idepth_min; p.idepth_min = p.idepth_max; p.idepth_max = t; }
  if (!std::isfinite(p.idepth_min) || !std::isfinite(p.idepth_max) || (p.idepth_max < 0)) {
    p.pixint = 0; p.uv[0] = p.uv[1] = -1;
    return p.status = SOSBA_IPS_OUTLIER;
  }
  p.pixint = 2 * errorInPixel;
  p.uv[0] = bestU; p.uv[1] = bestV;
  return p.status = SOSBA_IPS_GOOD;
}

// FullSystem::traceNewCoarse (FullSystem.cpp:311-361)
void trace_immature(Oracle &o, int frame_slot, int nhosts, const float *KRKi, const float *Kt, const float *aff, sosba_immature *pts, int32_t counts[6]) {
  const float *dI = o.slots[frame_slot].lvl[0].dI.data();
  for (int i = 0; i < 6; i++) counts[i] = 0;
  for (int k = 0; k < pts->n; k++) {
    IP p;
    p.u = pts->u[k]; p.v = pts->v[k]; p.idepth_min = pts->idepth_min[k]; p.idepth_max = pts->idepth_max[k]; p.quality = pts->quality[k];
    p.energyTH = pts->energy_th[k]; p.color = pts->color + 8 * k; p.weights = pts->weights + 8 * k; p.gradH = pts->gradH + 4 * k;
    p.status = pts->last_trace_status[k];
    p.uv[0] = pts->last_trace_uv[2 * k]; p.uv[1] = pts->last_trace_uv[2 * k + 1]; p.pixint = pts->last_trace_pixel_interval[k];
    const int h = pts->host[k];
    (void)nhosts;
    traceOn(o, p, dI, KRKi + 9 * h, Kt + 3 * h, aff + 2 * h);
    pts->idepth_min[k] = p.idepth_min; pts->idepth_max[k] = p.idepth_max; pts->quality[k] = p.quality;
    pts->last_trace_status[k] = (uint8_t)p.status;
    pts->last_trace_uv[2 * k] = p.uv[0]; pts->last_trace_uv[2 * k + 1] = p.uv[1]; pts->last_trace_pixel_interval[k] = p.pixint;
    counts[p.status]++;
  }
}


// ---- activation: ImmaturePoint::linearizeResidual (ImmaturePoint.cpp:475-545) + optimizeImmaturePoint (FullSystemOptPoint.cpp:47-192)
struct TmpRes {   // ImmaturePointTemporaryResidual (ImmaturePoint.h:31-38)
  int state_state, state_NewState;
  double state_energy, state_NewEnergy;
  int target;
};
struct ActPoint {
  float u, v, energyTH;
  const float *color, *weights;
  int host;
};

static double linearizeResidual(const Oracle &o, const sosba_activation_window *win, const ActPoint &p, float outlierTHSlack, TmpRes &tr, float &Hdd, float &bd,
                                float idepth) {
  if (tr.state_state == SOSBA_RES_OOB) { tr.state_NewState = SOSBA_RES_OOB; return tr.state_energy; }
  const size_t pair = (size_t)p.host * win->nf + tr.target;
  const float *R = win->RTll + 9 * pair, *t = win->tTll + 3 * pair, *affLL = win->aff + 2 * pair;
  const float fxl = win->calib[0], fyl = win->calib[1], cxl = win->calib[2], cyl = win->calib[3];
  const float fxli = 1.0f / fxl, fyli = 1.0f / fyl;   // HessianBlocks.h:494-495
  const float *dIl = o.slots[win->frame_slot[tr.target]].lvl[0].dI.data();
  const int wG = o.wl[0];
  const float wM3G = o.wl[0] - 3, hM3G = o.hl[0] - 3, huberTH = o.cfg.huber_th;
  float energyLeft = 0;
  for (int idx = 0; idx < 8; idx++) {
    const int dx = patternP[idx][0], dy = patternP[idx][1];
    // projectPoint (ResidualProjections.h:52-73)
    float KliP[3] = {(p.u + dx - cxl) * fxli, (p.v + dy - cyl) * fyli, 1};
    float ptp[3];
    for (int i = 0; i < 3; i++) ptp[i] = ((R[3 * i] * KliP[0] + R[3 * i + 1] * KliP[1]) + R[3 * i + 2] * KliP[2]) + t[i] * idepth;
    float drescale = 1.0f / ptp[2];
    bool ok = drescale > 0;
    float u = 0, v = 0, Ku = 0, Kv = 0;
    if (ok) {
      u = ptp[0] * drescale; v = ptp[1] * drescale;
      Ku = u * fxl + cxl; Kv = v * fyl + cyl;
      ok = Ku > 1.1f && Kv > 1.1f && Ku < wM3G && Kv < hM3G;
    }
    if (!ok) { tr.state_NewState = SOSBA_RES_OOB; return tr.state_energy; }
    float hit[3];
    interp33(dIl, Ku, Kv, wG, hit);
    if (!std::isfinite(hit[0])) { tr.state_NewState = SOSBA_RES_OOB; return tr.state_energy; }
    float residual = hit[0] - (affLL[0] * p.color[idx] + affLL[1]);
    float hw = fabsf(residual) < huberTH ? 1 : huberTH / fabsf(residual);
    energyLeft += p.weights[idx] * p.weights[idx] * hw * residual * residual * (2 - hw);
    float dxInterp = hit[1] * fxl, dyInterp = hit[2] * fyl;
    float d_idepth = (dxInterp * drescale * (t[0] - t[2] * u) + dyInterp * drescale * (t[1] - t[2] * v)) * SCALE_IDEPTH;   // derive_idepth (:32-40)
    hw *= p.weights[idx] * p.weights[idx];
    Hdd += (hw * d_idepth) * d_idepth;
    bd += (hw * residual) * d_idepth;
  }
  if (energyLeft > p.energyTH * outlierTHSlack) {
    energyLeft = p.energyTH * outlierTHSlack;
    tr.state_NewState = SOSBA_RES_OUTLIER;
  } else {
    tr.state_NewState = SOSBA_RES_IN;
  }
  tr.state_NewEnergy = energyLeft;
  return energyLeft;
}

// -> SOSBA_ACT_*; idepth_out = currentIdepth; states[nf] (255 at the host)
static int optimizeImmaturePoint(const Oracle &o, const sosba_activation_window *win, const ActPoint &p, float idepth_min, float idepth_max, TmpRes *residuals,
                                 float *idepth_out, uint8_t *states) {
  const float minIdepthH_act = 100;      // settings.cpp:61
  const int GNItsOnPointActivation = 3;  // settings.cpp:133
  int nres = 0;
  for (int f = 0; f < win->nf; f++) {
    states[f] = 255;
    if (f != p.host) {
      residuals[nres].state_NewEnergy = residuals[nres].state_energy = 0;
      residuals[nres].state_NewState = SOSBA_RES_OUTLIER;
      residuals[nres].state_state = SOSBA_RES_IN;
      residuals[nres].target = f;
      nres++;
    }
  }
  float lastEnergy = 0, lastHdd = 0, lastbd = 0;
  float currentIdepth = (idepth_max + idepth_min) * 0.5f;
  *idepth_out = currentIdepth;
  for (int i = 0; i < nres; i++) {
    lastEnergy += linearizeResidual(o, win, p, 1000, residuals[i], lastHdd, lastbd, currentIdepth);   // float += double
    residuals[i].state_state = residuals[i].state_NewState;
    residuals[i].state_energy = residuals[i].state_NewEnergy;
  }
  auto publish = [&]() { for (int i = 0; i < nres; i++) states[residuals[i].target] = (uint8_t)residuals[i].state_state; };
  if (!std::isfinite(lastEnergy) || lastHdd < minIdepthH_act) { publish(); return SOSBA_ACT_SKIP; }
  float lambda = 0.1;
  for (int iteration = 0; iteration < GNItsOnPointActivation; iteration++) {
    float H = lastHdd;
    H *= 1 + lambda;
    float step = (1.0 / H) * lastbd;     // evaluated in double, rounded once
    float newIdepth = currentIdepth - step;
    float newHdd = 0, newbd = 0, newEnergy = 0;
    for (int i = 0; i < nres; i++) newEnergy += linearizeResidual(o, win, p, 1, residuals[i], newHdd, newbd, newIdepth);
    if (!std::isfinite(lastEnergy) || newHdd < minIdepthH_act) { *idepth_out = currentIdepth; publish(); return SOSBA_ACT_SKIP; }
    if (newEnergy < lastEnergy) {
      currentIdepth = newIdepth; lastHdd = newHdd; lastbd = newbd; lastEnergy = newEnergy;
      for (int i = 0; i < nres; i++) {
        residuals[i].state_state = residuals[i].state_NewState;
        residuals[i].state_energy = residuals[i].state_NewEnergy;
      }
      lambda *= 0.5;
    } else {
      lambda *= 5;
    }
    if (fabsf(step) < 0.0001 * currentIdepth) break;   // double comparison
  }
  *idepth_out = currentIdepth;
  publish();
  if (!std::isfinite(currentIdepth)) return SOSBA_ACT_DELETE;
  int numGoodRes = 0;
  for (int i = 0; i < nres; i++)
    if (residuals[i].state_state == SOSBA_RES_IN) numGoodRes++;
  if (numGoodRes < win->min_obs) return SOSBA_ACT_DELETE;
  if (!std::isfinite(p.energyTH)) return SOSBA_ACT_DELETE;   // PointHessian ctor copies energyTH (HessianBlocks.cpp:55)
  return SOSBA_ACT_ACTIVATED;
}

// FullSystem::activatePointsMT_Reductor (FullSystem.cpp:363-373)
void optimize_immature(Oracle &o, const sosba_activation_window *win, const sosba_immature *pts, int8_t *result, float *idepth, uint8_t *res_state) {
  std::vector<TmpRes> tr(win->nf);
  for (int k = 0; k < pts->n; k++) {
    ActPoint p;
    p.u = pts->u[k]; p.v = pts->v[k]; p.energyTH = pts->energy_th[k]; p.color = pts->color + 8 * (size_t)k; p.weights = pts->weights + 8 * (size_t)k;
    p.host = pts->host[k];
    result[k] = (int8_t)optimizeImmaturePoint(o, win, p, pts->idepth_min[k], pts->idepth_max[k], tr.data(), idepth + k, res_state + (size_t)k * win->nf);
  }
}


// ---- pre-pyramid image path: PhotometricUndistorter::processFrame (util/Undistort.cpp:194-227) + Undistort::undistort (:361-458)
void undistort_raw(const Oracle &o, const void *raw, int raw_bits, float factor, float *out) {
  const Oracle::Undist &U = o.und;
  const int wOrg = U.wOrg, hOrg = U.hOrg, w = o.wl[0], h = o.hl[0];
  std::vector<float> data((size_t)wOrg * hOrg);
  const uint8_t *r8 = (const uint8_t *)raw;
  const uint16_t *r16 = (const uint16_t *)raw;
  for (int i = 0; i < wOrg * hOrg; i++) {
    const int v = raw_bits == 8 ? r8[i] : r16[i];
    if (!U.haveG) data[i] = factor * v;
    else {
      data[i] = U.G[v];
      if (U.haveV) data[i] *= U.vignetteInv[i];
    }
  }
  if (U.passthrough) { memcpy(out, data.data(), sizeof(float) * w * h); return; }
  for (int idx = w * h - 1; idx >= 0; idx--) {
    float xx = U.remapX[idx], yy = U.remapY[idx];
    if (xx < 0) out[idx] = 0;
    else {
      int xxi = xx, yyi = yy;
      xx -= xxi; yy -= yyi;
      float xxyy = xx * yy;
      const float *src = data.data() + xxi + yyi * wOrg;
      out[idx] = xxyy * src[1 + wOrg] + (yy - xxyy) * src[wOrg] + (xx - xxyy) * src[1] + (1 - xx - yy + xxyy) * src[0];
    }
  }
}

}  // namespace orc
