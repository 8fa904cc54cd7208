// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).  PARITY UNPINNED.
//
// orc_capi.cpp: the oracle behind the same C signatures as include/sosba.h with an `orc_` prefix, so
// the parity tests drive product and oracle through one Python binding.
#include <cstdio>
#include <cstring>
#include <new>

#include "orc_core.h"
#include <map>

#include "orc_host.h"

using namespace orc;

struct orc_handle {
  Oracle o;
  BAState ba;
  bool ba_loaded = false;
  std::vector<double> HM, bM;
};

#define ORC_API extern "C" __attribute__((visibility("default")))

ORC_API void orc_config_default(sosba_config *c, int32_t w, int32_t h) {
  memset(c, 0, sizeof(*c));
  c->w = w; c->h = h; c->pyr_levels = 0; c->max_frames = 16; c->num_threads = 1;
  c->gamma_weights_pixel_select = 1; c->min_opt_iterations = 1;
  c->huber_th = 9; c->outlier_th_sum_component = 50 * 50;
  c->affine_opt_mode_a = 0; c->affine_opt_mode_b = 0;  // settingsDefault(mode=1), main.cpp:75-80
  c->coarse_cutoff_th = 20; c->idepth_fix_prior = 50 * 50; c->idepth_fix_prior_marg_fac = 600 * 600;
  c->frame_energy_th_const_weight = 0.5f; c->frame_energy_th_n = 0.7f; c->frame_energy_th_fac_median = 1.5f;
  c->overall_energy_th_weight = 1; c->initial_calib_hessian = 5e9f;
  c->initial_rot_prior = 1e11f; c->initial_trans_prior = 1e10f; c->initial_aff_a_prior = 1e14f; c->initial_aff_b_prior = 1e14f;
  c->marg_weight_fac = 0.5f * 0.5f; c->th_opt_iterations = 1.2f;
}

ORC_API int orc_create(const sosba_config *cfg, int32_t /*device*/, orc_handle **out) {
  if (!cfg || !out || cfg->w <= 0 || cfg->h <= 0) return SOSBA_E_ARG;
  orc_handle *h = new (std::nothrow) orc_handle();
  if (!h) return SOSBA_E_ARG;
  Oracle &o = h->o;
  o.cfg = *cfg;
  // setGlobalCalib (globalCalib.cpp:39-49)
  int wlvl = cfg->w, hlvl = cfg->h, lv = 1;
  while (wlvl % 2 == 0 && hlvl % 2 == 0 && wlvl * hlvl > 5000 && lv < SOSBA_MAX_LEVELS) { wlvl /= 2; hlvl /= 2; lv++; }
  o.levels = cfg->pyr_levels > 0 ? cfg->pyr_levels : lv;
  for (int l = 0; l < o.levels; l++) { o.wl[l] = cfg->w >> l; o.hl[l] = cfg->h >> l; }
  o.wM3G = cfg->w - 3; o.hM3G = cfg->h - 3;
  o.slots.resize(cfg->max_frames > 0 ? cfg->max_frames : 16);
  o.T = cfg->num_threads > 1 ? cfg->num_threads : 1;
  o.MT = cfg->num_threads > 1;
  if (o.MT) o.red.reset(new ThreadReduce(o.T));
  o.accA.resize(o.T); o.accL.resize(o.T); o.nresA.assign(o.T, 0); o.nresL.assign(o.T, 0);
  o.accE.resize(o.T); o.accEB.resize(o.T); o.accD.resize(o.T); o.accHcc.resize(o.T); o.accbc.resize(o.T);
  *out = h;
  return SOSBA_OK;
}
ORC_API void orc_destroy(orc_handle *h) { delete h; }
ORC_API const char *orc_last_error(void) { return ""; }
ORC_API int32_t orc_pyr_levels(const orc_handle *h) { return h->o.levels; }

ORC_API int orc_frame_make_images(orc_handle *h, int32_t slot, const float *color, const float *B) {
  if (slot < 0 || slot >= (int)h->o.slots.size() || !color) return SOSBA_E_ARG;
  make_images(h->o, slot, color, B);
  return SOSBA_OK;
}
ORC_API int orc_frame_get_level(orc_handle *h, int32_t slot, int32_t lvl, float *dI3, float *absg) {
  if (slot < 0 || slot >= (int)h->o.slots.size() || lvl < 0 || lvl >= h->o.levels || !h->o.slots[slot].valid) return SOSBA_E_ARG;
  const Level &L = h->o.slots[slot].lvl[lvl];
  if (dI3) memcpy(dI3, L.dI.data(), L.dI.size() * sizeof(float));
  if (absg) memcpy(absg, L.absg.data(), L.absg.size() * sizeof(float));
  return SOSBA_OK;
}

static int window_apply(Oracle &o, const sosba_window *w, bool full) {
  const int nf = w->nf;
  if (nf <= 0) return SOSBA_E_ARG;
  if (!full && nf != o.nf) return SOSBA_E_STATE;
  o.nf = nf;
  if (full) {
    o.frame_slot.assign(w->frame_slot, w->frame_slot + nf);
    o.adHost.assign(w->adHost, w->adHost + (size_t)nf * nf * 64);
    o.adTarget.assign(w->adTarget, w->adTarget + (size_t)nf * nf * 64);
    o.adHostF.resize(o.adHost.size()); o.adTargetF.resize(o.adTarget.size());
    for (size_t i = 0; i < o.adHost.size(); i++) { o.adHostF[i] = (float)o.adHost[i]; o.adTargetF[i] = (float)o.adTarget[i]; }
    for (int i = 0; i < 4; i++) o.cPrior[i] = w->cPrior[i];
    o.fprior.assign(w->frame_prior, w->frame_prior + (size_t)nf * 8);
  }
  o.pre.resize((size_t)nf * nf);
  for (int i = 0; i < nf * nf; i++) {
    const float *p = w->precalc + (size_t)i * SOSBA_PRECALC_FLOATS;
    Precalc &pc = o.pre[i];
    memcpy(pc.RTll_0, p + SOSBA_PC_RTLL0, 9 * sizeof(float)); memcpy(pc.tTll_0, p + SOSBA_PC_TTLL0, 3 * sizeof(float));
    memcpy(pc.KRKi, p + SOSBA_PC_KRKI, 9 * sizeof(float)); memcpy(pc.Kt, p + SOSBA_PC_KT, 3 * sizeof(float));
    pc.aff[0] = p[SOSBA_PC_AFF]; pc.aff[1] = p[SOSBA_PC_AFF + 1]; pc.b0 = p[SOSBA_PC_B0]; pc.dist = p[SOSBA_PC_DIST];
  }
  o.adHTdeltaF.assign(w->adHTdeltaF, w->adHTdeltaF + (size_t)nf * nf * 8);
  o.frameEnergyTH.assign(w->frame_energy_th, w->frame_energy_th + nf);
  o.fxl = w->calib[0]; o.fyl = w->calib[1]; o.cxl = w->calib[2]; o.cyl = w->calib[3];
  o.fxli = 1.0f / o.fxl; o.fyli = 1.0f / o.fyl;  // CalibHessian::setValue, HessianBlocks.h:495-496
  for (int i = 0; i < 4; i++) o.cDeltaF[i] = w->cDeltaF[i];
  o.fdelta_prior.assign(w->frame_delta_prior, w->frame_delta_prior + (size_t)nf * 8);
  o.fdelta.assign(w->frame_delta, w->frame_delta + (size_t)nf * 8);
  return SOSBA_OK;
}
ORC_API int orc_window_set(orc_handle *h, const sosba_window *w) { return window_apply(h->o, w, true); }
ORC_API int orc_window_update(orc_handle *h, const sosba_window *w) { return window_apply(h->o, w, false); }

ORC_API int orc_points_set(orc_handle *h, const sosba_points *p) {
  Oracle &o = h->o;
  o.pts.assign(p->n, Pt());
  for (int i = 0; i < p->n; i++) {
    Pt &q = o.pts[i];
    memset(&q, 0, sizeof(q));
    q.u = p->u[i]; q.v = p->v[i];
    q.idepth = p->idepth[i]; q.idepth_scaled = SCALE_IDEPTH * q.idepth;
    q.idepth_zero = p->idepth_zero[i]; q.idepth_zero_scaled = SCALE_IDEPTH * q.idepth_zero;
    memcpy(q.color, p->color + 8 * (size_t)i, 8 * sizeof(float)); memcpy(q.weights, p->weights + 8 * (size_t)i, 8 * sizeof(float));
    q.host = p->host[i]; q.priorF = p->priorF ? p->priorF[i] : 0; q.deltaF = p->deltaF ? p->deltaF[i] : 0;
    q.res_begin = q.res_end = 0;
  }
  return SOSBA_OK;
}
ORC_API int orc_points_update(orc_handle *h, const float *idepth, const float *idepth_zero, const float *deltaF) {
  Oracle &o = h->o;
  for (size_t i = 0; i < o.pts.size(); i++) {
    Pt &q = o.pts[i];
    if (idepth) { q.idepth = idepth[i]; q.idepth_scaled = SCALE_IDEPTH * q.idepth; }
    if (idepth_zero) { q.idepth_zero = idepth_zero[i]; q.idepth_zero_scaled = SCALE_IDEPTH * q.idepth_zero; }
    if (deltaF) q.deltaF = deltaF[i];
  }
  return SOSBA_OK;
}
ORC_API int orc_residuals_set(orc_handle *h, const sosba_residuals *r) {
  Oracle &o = h->o;
  if (!r || r->n < 0 || (r->n > 0 && (!r->point || !r->target))) return SOSBA_E_ARG;
  o.res.assign(r->n, Res());
  int prev = -1;
  for (auto &p : o.pts) p.res_begin = p.res_end = 0;
  for (int i = 0; i < r->n; i++) {
    Res &q = o.res[i];
    memset(&q, 0, sizeof(q));
    q.point = r->point[i];
    if (q.point < 0 || q.point < prev || q.point >= (int)o.pts.size() || r->target[i] < 0 || r->target[i] >= o.nf) return SOSBA_E_ARG;
    if (q.point != prev) { o.pts[q.point].res_begin = i; prev = q.point; }
    o.pts[q.point].res_end = i + 1;
    q.host = o.pts[q.point].host; q.target = r->target[i];
    q.state_state = r->state ? r->state[i] : SOSBA_RES_IN;
    q.state_NewState = SOSBA_RES_OUTLIER;
    q.state_energy = r->state_energy ? r->state_energy[i] : 0; q.state_NewEnergy = q.state_energy; q.state_NewEnergyWithOutlier = -1;
    q.isLinearized = r->is_linearized ? r->is_linearized[i] != 0 : false;
    q.isActive = r->is_active ? r->is_active[i] != 0 : false;
    q.isNew = r->is_new ? r->is_new[i] != 0 : true;
    q.dropped = false; q.sel = 0;
  }
  // empty points: res_begin == res_end
  o.activeResiduals.clear();
  for (int i = 0; i < r->n; i++) if (!o.res[i].isLinearized) o.activeResiduals.push_back(i);
  return SOSBA_OK;
}

ORC_API int orc_reset_oob(orc_handle *h) {
  Oracle &o = h->o;
  o.activeResiduals.clear();
  for (int i = 0; i < (int)o.res.size(); i++) {
    Res &r = o.res[i];
    if (r.dropped || r.isLinearized) continue;
    o.activeResiduals.push_back(i);
    r.state_NewEnergy = r.state_energy = 0; r.state_NewState = SOSBA_RES_OUTLIER; r.state_state = SOSBA_RES_IN;  // Residuals.h:81-86
  }
  return SOSBA_OK;
}
ORC_API int orc_linearize_all(orc_handle *h, int32_t fix, sosba_linearize_out *out) {
  linearizeAll(h->o, fix != 0, out);
  // setNewFrameEnergyTH writes newFrame->frameEnergyTH on the FrameHessian itself (FullSystemOptimize.cpp:116): keep the frame record
  // of a resident window (orc_ba_upload) in step, so that the next setPrecalcValues does not bring the old threshold back
  if (h->ba_loaded && !h->ba.frames.empty() && (int)h->ba.frames.size() == h->o.nf) h->ba.frames.back().frameEnergyTH = h->o.frameEnergyTH[h->o.nf - 1];
  return SOSBA_OK;
}
ORC_API int orc_apply_res(orc_handle *h) {
  Oracle &o = h->o;
  auto fn = [&](int a, int b) { for (int k = a; k < b; k++) applyRes(o.res[o.activeResiduals[k]], true); };
  if (o.MT) o.red->reduce([&](int a, int b, Stats10 *, int) { fn(a, b); }, 0, (int)o.activeResiduals.size(), 50);
  else fn(0, (int)o.activeResiduals.size());
  return SOSBA_OK;
}
ORC_API int orc_fix_linearization(orc_handle *h, const int32_t *ids, int32_t n) {
  for (int i = 0; i < n; i++) { if (ids[i] < 0 || ids[i] >= (int)h->o.res.size()) return SOSBA_E_ARG; fixLinearizationF(h->o, h->o.res[ids[i]]); }
  return SOSBA_OK;
}

ORC_API int orc_residuals_get_state(orc_handle *h, uint8_t *state, uint8_t *new_state, float *energy, float *new_energy, float *new_energy_wo,
                                    uint8_t *is_active, uint8_t *is_linearized) {
  Oracle &o = h->o;
  for (size_t i = 0; i < o.res.size(); i++) {
    const Res &r = o.res[i];
    if (state) state[i] = (uint8_t)r.state_state;
    if (new_state) new_state[i] = (uint8_t)r.state_NewState;
    if (energy) energy[i] = (float)r.state_energy;
    if (new_energy) new_energy[i] = (float)r.state_NewEnergy;
    if (new_energy_wo) new_energy_wo[i] = (float)r.state_NewEnergyWithOutlier;
    if (is_active) is_active[i] = r.isActive;
    if (is_linearized) is_linearized[i] = r.isLinearized;
  }
  return SOSBA_OK;
}
static void dumpJ(const RawJ &J, float *o) {
  int k = 0;
  for (int i = 0; i < 8; i++) o[k++] = J.resF[i];
  for (int a = 0; a < 2; a++) for (int i = 0; i < 6; i++) o[k++] = J.Jpdxi[a][i];
  for (int a = 0; a < 2; a++) for (int i = 0; i < 4; i++) o[k++] = J.Jpdc[a][i];
  o[k++] = J.Jpdd[0]; o[k++] = J.Jpdd[1];
  for (int a = 0; a < 2; a++) for (int i = 0; i < 8; i++) o[k++] = J.JIdx[a][i];
  for (int a = 0; a < 2; a++) for (int i = 0; i < 8; i++) o[k++] = J.JabF[a][i];
  for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) o[k++] = J.JIdx2[a][b];
  for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) o[k++] = J.JabJIdx[a][b];
  for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) o[k++] = J.Jab2[a][b];
}
ORC_API int orc_residuals_get_jacobians(orc_handle *h, int32_t committed, float *J) {
  Oracle &o = h->o;
  for (size_t i = 0; i < o.res.size(); i++) dumpJ(committed ? o.res[i].Jef() : o.res[i].Jdata(), J + SOSBA_J_FLOATS * i);
  return SOSBA_OK;
}
ORC_API int orc_residuals_get_aux(orc_handle *h, float *JpJdF, float *rtz, float *proj, float *center) {
  Oracle &o = h->o;
  for (size_t i = 0; i < o.res.size(); i++) {
    const Res &r = o.res[i];
    if (JpJdF) memcpy(JpJdF + 8 * i, r.JpJdF, 8 * sizeof(float));
    if (rtz) memcpy(rtz + 8 * i, r.res_toZeroF, 8 * sizeof(float));
    if (proj) memcpy(proj + 16 * i, r.projectedTo, 16 * sizeof(float));
    if (center) memcpy(center + 3 * i, r.centerProjectedTo, 3 * sizeof(float));
  }
  return SOSBA_OK;
}
ORC_API int orc_points_get_stats(orc_handle *h, float *mrb, int32_t *ngr) {
  Oracle &o = h->o;
  for (size_t i = 0; i < o.pts.size(); i++) { if (mrb) mrb[i] = o.pts[i].maxRelBaseline; if (ngr) ngr[i] = o.pts[i].numGoodResiduals; }
  return SOSBA_OK;
}

ORC_API int orc_accumulate(orc_handle *h, double *HA, double *bA, double *HL, double *bL, double *Hsc, double *bsc, int32_t *resInA, int32_t *resInL) {
  Oracle &o = h->o;
  const int D = CPARS + 8 * o.nf;
  std::vector<double> H((size_t)D * D), b(D);
  accumulateAF(o, H.data(), b.data());
  if (HA) memcpy(HA, H.data(), sizeof(double) * D * D);
  if (bA) memcpy(bA, b.data(), sizeof(double) * D);
  accumulateLF(o, H.data(), b.data());
  if (HL) memcpy(HL, H.data(), sizeof(double) * D * D);
  if (bL) memcpy(bL, b.data(), sizeof(double) * D);
  accumulateSCF(o, H.data(), b.data());
  if (Hsc) memcpy(Hsc, H.data(), sizeof(double) * D * D);
  if (bsc) memcpy(bsc, b.data(), sizeof(double) * D);
  if (resInA) *resInA = o.resInA;
  if (resInL) *resInL = o.resInL;
  return SOSBA_OK;
}
ORC_API int orc_points_get_acc(orc_handle *h, float *HddA, float *bdA, float *HcdA, float *HddL, float *bdL, float *HcdL, float *HdiF, float *bdSumF) {
  Oracle &o = h->o;
  for (size_t i = 0; i < o.pts.size(); i++) {
    const Pt &p = o.pts[i];
    if (HddA) HddA[i] = p.Hdd_accAF;
    if (bdA) bdA[i] = p.bd_accAF;
    if (HcdA) memcpy(HcdA + 4 * i, p.Hcd_accAF, 4 * sizeof(float));
    if (HddL) HddL[i] = p.Hdd_accLF;
    if (bdL) bdL[i] = p.bd_accLF;
    if (HcdL) memcpy(HcdL + 4 * i, p.Hcd_accLF, 4 * sizeof(float));
    if (HdiF) HdiF[i] = p.HdiF;
    if (bdSumF) bdSumF[i] = p.bdSumF;
  }
  return SOSBA_OK;
}
ORC_API int orc_solve_system(orc_handle *h, const double *HM, const double *bM, double *x, double *Hf, double *bf) {
  solveSystemF(h->o, HM, bM, x, Hf, bf);
  return SOSBA_OK;
}
ORC_API int orc_resubstitute(orc_handle *h, const double *x, float *step) {
  resubstituteF(h->o, x);
  if (step) for (size_t i = 0; i < h->o.pts.size(); i++) step[i] = h->o.pts[i].step;
  return SOSBA_OK;
}
ORC_API int orc_marginalize_points(orc_handle *h, const int32_t *ids, int32_t n, double *H, double *b, int32_t *resInM) {
  int r = 0;
  marginalizePoints(h->o, ids, n, H, b, &r);
  if (resInM) *resInM = r;
  return SOSBA_OK;
}

ORC_API int orc_tracker_make_k(orc_handle *h, const float calib[4]) { tracker_makeK(h->o, calib); return SOSBA_OK; }
ORC_API int orc_tracker_set_ref(orc_handle *h, int32_t lvl, int32_t n, const float *u, const float *v, const float *id, const float *c) {
  if (lvl < 0 || lvl >= h->o.levels) return SOSBA_E_ARG;
  h->o.pc_u[lvl].assign(u, u + n); h->o.pc_v[lvl].assign(v, v + n); h->o.pc_idepth[lvl].assign(id, id + n); h->o.pc_color[lvl].assign(c, c + n);
  return SOSBA_OK;
}
ORC_API int orc_tracker_calc_res_pose(orc_handle *h, int32_t lvl, int32_t slot, const double refToNew[12], const float affLL[2], float cutoff,
                                      double out6[6], int32_t counts[3]) {
  tracker_calcResPose(h->o, lvl, slot, refToNew, affLL, cutoff, out6, counts);
  return SOSBA_OK;
}
ORC_API int orc_tracker_calc_gs_pose(orc_handle *h, int32_t lvl, float a, float b0, double H[64], double b[8]) {
  tracker_calcGSSSEPose(h->o, lvl, a, b0, H, b);
  return SOSBA_OK;
}
// ---- a16: control loops of the direct alignment (orc_lm.cpp) --------------------------------------------------
ORC_API int orc_tracker_make_coarse_depth(orc_handle *h, int32_t ref_slot, int32_t n, const float *cpt, const float *HdiF, int32_t *pc_n_out) {
  Oracle &o = h->o;
  if (ref_slot < 0 || ref_slot >= (int)o.slots.size() || !o.slots[ref_slot].valid || n < 0 || (n > 0 && (!cpt || !HdiF))) return SOSBA_E_ARG;
  for (int i = 0; i < n; i++) {
    const int u = cpt[3 * i] + 0.5f, v = cpt[3 * i + 1] + 0.5f;
    if (!(cpt[3 * i] >= 0) || !(cpt[3 * i + 1] >= 0) || u >= o.wl[0] || v >= o.hl[0]) return SOSBA_E_ARG;
  }
  tracker_makeCoarseDepth(o, ref_slot, n, cpt, HdiF);
  if (pc_n_out) for (int l = 0; l < o.levels; l++) pc_n_out[l] = (int)o.pc_u[l].size();
  return SOSBA_OK;
}
ORC_API int orc_tracker_get_ref(orc_handle *h, int32_t lvl, int32_t *n, float *u, float *v, float *id, float *c) {
  Oracle &o = h->o;
  if (lvl < 0 || lvl >= o.levels) return SOSBA_E_ARG;
  const size_t m = o.pc_u[lvl].size();
  if (n) *n = (int)m;
  if (u) memcpy(u, o.pc_u[lvl].data(), 4 * m);
  if (v) memcpy(v, o.pc_v[lvl].data(), 4 * m);
  if (id) memcpy(id, o.pc_idepth[lvl].data(), 4 * m);
  if (c) memcpy(c, o.pc_color[lvl].data(), 4 * m);
  return SOSBA_OK;
}
ORC_API int orc_tracker_scale_coarse_depth(orc_handle *h, float scale) { tracker_scaleCoarseDepth(h->o, scale); return SOSBA_OK; }
ORC_API int orc_tracker_track(orc_handle *h, int32_t new_slot, float ref_exp, float new_exp, const double ref_aff[2], int32_t coarsest, int32_t n_hyp,
                              sosba_track_hypothesis *hyps) {
  Oracle &o = h->o;
  if (new_slot < 0 || new_slot >= (int)o.slots.size() || !o.slots[new_slot].valid || coarsest < 0 || coarsest >= o.levels || coarsest >= 5 || n_hyp < 0 ||
      (n_hyp > 0 && !hyps) || !ref_aff)
    return SOSBA_E_ARG;
  for (int i = 0; i < n_hyp; i++) tracker_track(o, new_slot, ref_exp, new_exp, ref_aff, coarsest, hyps + i);
  return SOSBA_OK;
}
ORC_API int orc_scale_optimize(orc_handle *h, int32_t stereo_slot, int32_t coarsest, int32_t n_hyp, sosba_scale_hypothesis *hyps) {
  Oracle &o = h->o;
  if (stereo_slot < 0 || stereo_slot >= (int)o.slots.size() || !o.slots[stereo_slot].valid || coarsest < 0 || coarsest >= o.levels || coarsest >= 5 ||
      n_hyp < 0 || (n_hyp > 0 && !hyps))
    return SOSBA_E_ARG;
  for (int i = 0; i < n_hyp; i++) scale_optimize(o, stereo_slot, coarsest, hyps + i);
  return SOSBA_OK;
}
ORC_API int orc_distance_map(orc_handle *h, int32_t nhosts, const float *KRKi, const float *Kt, int32_t n, const int32_t *host, const float *u, const float *v,
                             const float *idepth, float *dist_out) {
  if (h->o.levels < 2 || nhosts < 0 || n < 0 || !dist_out || (n > 0 && (!KRKi || !Kt || !host || !u || !v || !idepth))) return SOSBA_E_ARG;
  for (int i = 0; i < n; i++) if (host[i] < 0 || host[i] >= nhosts) return SOSBA_E_ARG;
  distance_map(h->o, nhosts, KRKi, Kt, n, host, u, v, idepth, dist_out);
  return SOSBA_OK;
}
ORC_API int orc_scale_set_stereo(orc_handle *h, const double T10[12], const float K1[4]) {
  Oracle &o = h->o;
  o.tfmF0ToF1 = SE3::from_rowmajor34(T10);
  o.fx1[0] = K1[0]; o.fy1[0] = K1[1]; o.cx1[0] = K1[2]; o.cy1[0] = K1[3];
  for (int level = 1; level < o.levels; ++level) {  // ScaleOptimizer.cpp:72-77
    o.fx1[level] = o.fx1[level - 1] * 0.5; o.fy1[level] = o.fy1[level - 1] * 0.5;
    o.cx1[level] = (o.cx1[0] + 0.5) / ((int)1 << level) - 0.5; o.cy1[level] = (o.cy1[0] + 0.5) / ((int)1 << level) - 0.5;
  }
  return SOSBA_OK;
}
ORC_API int orc_scale_calc_res(orc_handle *h, int32_t lvl, int32_t slot, float scale, float cutoff, double out6[6], int32_t counts[3]) {
  scale_calcRes(h->o, lvl, slot, scale, cutoff, out6, counts);
  return SOSBA_OK;
}
ORC_API int orc_scale_calc_gs(orc_handle *h, int32_t lvl, float scale, float *H, float *b) { scale_calcGSSSE(h->o, lvl, scale, H, b); return SOSBA_OK; }

// ---- composed GN loop ---------------------------------------------------------------------------
ORC_API int orc_ba_upload(orc_handle *h, const sosba_ba_problem *prob) {
  Oracle &o = h->o;
  o.nf = prob->nf;
  int rc = orc_points_set(h, &prob->points);
  if (rc) return rc;
  rc = orc_residuals_set(h, &prob->residuals);
  if (rc) return rc;
  o.frame_slot.resize(prob->nf);
  for (int i = 0; i < prob->nf; i++) o.frame_slot[i] = prob->frames[i].slot;
  h->ba.load(o, prob);
  const int D = CPARS + 8 * prob->nf;
  if (prob->HM && prob->bM) { h->HM.assign(prob->HM, prob->HM + (size_t)D * D); h->bM.assign(prob->bM, prob->bM + D); }
  else { h->HM.clear(); h->bM.clear(); }
  h->ba.setAdjoints(o);
  h->ba.setPrecalcValues(o);
  h->ba_loaded = true;
  return SOSBA_OK;
}
ORC_API int orc_optimize(orc_handle *h, sosba_ba_problem *prob, int32_t max_it, sosba_optimize_out *out) {
  int rc = orc_ba_upload(h, prob);
  if (rc) return rc;
  h->ba.optimize(h->o, h->HM.empty() ? nullptr : h->HM.data(), h->bM.empty() ? nullptr : h->bM.data(), max_it, out);
  h->ba.store(h->o, prob);
  return SOSBA_OK;
}
ORC_API int orc_ba_optimize(orc_handle *h, int32_t max_it, sosba_optimize_out *out) {
  if (!h->ba_loaded) return SOSBA_E_STATE;
  h->ba.optimize(h->o, h->HM.empty() ? nullptr : h->HM.data(), h->bM.empty() ? nullptr : h->bM.data(), max_it, out);
  return SOSBA_OK;
}
ORC_API int orc_ba_iterate(orc_handle *h, int32_t n, int32_t *n_res) {
  if (!h->ba_loaded) return SOSBA_E_STATE;
  sosba_linearize_out lo;
  for (int i = 0; i < n; i++) h->ba.iterate(h->o, h->HM.empty() ? nullptr : h->HM.data(), h->bM.empty() ? nullptr : h->bM.data(), &lo);
  if (n_res) *n_res = (int)h->o.activeResiduals.size();
  return SOSBA_OK;
}
ORC_API int orc_ba_download(orc_handle *h, sosba_ba_problem *prob) {
  if (!h->ba_loaded) return SOSBA_E_STATE;
  h->ba.store(h->o, prob);
  return SOSBA_OK;
}

// the loop body split around a caller-side solve (IMU configurations): see include/sosba.h
ORC_API int orc_ba_system(orc_handle *h, double *H_top, double *b_top, double *H_sc, double *b_sc, int32_t *resInA, int32_t *resInL) {
  if (!h->ba_loaded) return SOSBA_E_STATE;
  Oracle &o = h->o;
  const int D = CPARS + 8 * o.nf;
  std::vector<double> HA((size_t)D * D), bA(D), HL((size_t)D * D), bL(D), Hs((size_t)D * D), bs(D);
  accumulateAF(o, HA.data(), bA.data());
  accumulateLF(o, HL.data(), bL.data());
  accumulateSCF(o, Hs.data(), bs.data());
  for (size_t i = 0; i < (size_t)D * D; i++) { if (H_top) H_top[i] = HL[i] + HA[i]; if (H_sc) H_sc[i] = Hs[i]; }   // EnergyFunctional.cpp:1046-1047
  for (int i = 0; i < D; i++) { if (b_top) b_top[i] = bL[i] + bA[i]; if (b_sc) b_sc[i] = bs[i]; }
  if (resInA) *resInA = o.resInA;
  if (resInL) *resInL = o.resInL;
  return SOSBA_OK;
}
ORC_API int orc_ba_step(orc_handle *h, const double *x, sosba_step_out *out) {
  if (!h->ba_loaded || !x || !out) return SOSBA_E_STATE;
  Oracle &o = h->o;
  sosba_linearize_out lo;
  double sums[7];
  h->ba.step_with_x(o, x, &lo, sums);
  out->energy = lo.energy; out->new_frame_energy_th = lo.new_frame_energy_th; out->n_in = lo.n_in; out->n_oob = lo.n_oob; out->n_outlier = lo.n_outlier;
  out->sum_a = sums[0]; out->sum_b = sums[1]; out->sum_t = sums[2]; out->sum_r = sums[3]; out->sum_id = sums[4]; out->sum_nid = sums[5]; out->num_id = sums[6];
  return SOSBA_OK;
}

// host-side table helpers exposed for the tests (restated FrameFramePrecalc::set / setAdjointsF / setDeltaF)
ORC_API int orc_host_tables(orc_handle *h, const sosba_ba_problem *prob, float *precalc, double *adHost, double *adTarget, float *adHTdeltaF) {
  Oracle &o = h->o;
  BAState ba;
  ba.load(o, prob);
  const int nf = prob->nf;
  std::vector<double> aH, aT;
  set_adjoints(ba.frames, aH, aT);
  std::vector<float> aHF(aH.size()), aTF(aT.size()), dF;
  for (size_t i = 0; i < aH.size(); i++) { aHF[i] = (float)aH[i]; aTF[i] = (float)aT[i]; }
  set_delta(ba.frames, aHF, aTF, dF);
  if (adHost) memcpy(adHost, aH.data(), aH.size() * sizeof(double));
  if (adTarget) memcpy(adTarget, aT.data(), aT.size() * sizeof(double));
  if (adHTdeltaF) memcpy(adHTdeltaF, dF.data(), dF.size() * sizeof(float));
  if (precalc) {
    memset(precalc, 0, sizeof(float) * nf * nf * SOSBA_PRECALC_FLOATS);
    for (int hh = 0; hh < nf; hh++)
      for (int t = 0; t < nf; t++) {
        Precalc pc;
        precalc_set(ba.frames[hh], ba.frames[t], ba.calib, pc);
        float *p = precalc + (size_t)(hh * nf + t) * SOSBA_PRECALC_FLOATS;
        memcpy(p + SOSBA_PC_RTLL0, pc.RTll_0, 9 * sizeof(float)); memcpy(p + SOSBA_PC_TTLL0, pc.tTll_0, 3 * sizeof(float));
        memcpy(p + SOSBA_PC_KRKI, pc.KRKi, 9 * sizeof(float)); memcpy(p + SOSBA_PC_KT, pc.Kt, 3 * sizeof(float));
        p[SOSBA_PC_AFF] = pc.aff[0]; p[SOSBA_PC_AFF + 1] = pc.aff[1]; p[SOSBA_PC_B0] = pc.b0; p[SOSBA_PC_DIST] = pc.dist;
      }
  }
  return SOSBA_OK;
}

// SE3 helpers for the Sophus-property tests (tests/test_oracle_se3.py)
ORC_API void orc_se3_exp(const double a[6], double T[12]) { se3_exp(a).to_rowmajor34(T); }
ORC_API void orc_se3_log(const double T[12], double a[6]) { se3_log(SE3::from_rowmajor34(T), a); }
ORC_API void orc_se3_adj(const double T[12], double A[36]) { se3_adj(SE3::from_rowmajor34(T), A); }
ORC_API void orc_ldlt_solve(const double *A, const double *b, double *x, int32_t n) { ldlt_solve(A, b, x, n); }

// ---- next row: immature points ----------------------------------------------------
ORC_API int orc_immature_init(orc_handle *h, int32_t slot, int32_t n, const int32_t *u, const int32_t *v, float *color, float *weights, float *gradH,
                              float *energy_th) {
  if (slot < 0 || slot >= (int)h->o.slots.size() || n < 0) return SOSBA_E_ARG;
  immature_init(h->o, slot, n, u, v, color, weights, gradH, energy_th);
  return SOSBA_OK;
}
ORC_API int orc_trace_immature(orc_handle *h, int32_t frame_slot, int32_t nhosts, const float *KRKi, const float *Kt, const float *aff,
                               sosba_immature *pts, int32_t counts[6]) {
  if (frame_slot < 0 || frame_slot >= (int)h->o.slots.size() || !pts) return SOSBA_E_ARG;
  trace_immature(h->o, frame_slot, nhosts, KRKi, Kt, aff, pts, counts);
  return SOSBA_OK;
}
ORC_API int orc_optimize_immature(orc_handle *h, const sosba_activation_window *win, const sosba_immature *pts, int8_t *result, float *idepth,
                                  uint8_t *res_state) {
  if (!win || !pts || win->nf < 1) return SOSBA_E_ARG;
  for (int f = 0; f < win->nf; f++)
    if (win->frame_slot[f] < 0 || win->frame_slot[f] >= (int)h->o.slots.size()) return SOSBA_E_ARG;
  for (int k = 0; k < pts->n; k++)
    if (pts->host[k] < 0 || pts->host[k] >= win->nf) return SOSBA_E_ARG;
  optimize_immature(h->o, win, pts, result, idepth, res_state);
  return SOSBA_OK;
}

// ---- next row: pre-pyramid image path ---------------------------------------------
ORC_API int orc_undistort_set(orc_handle *h, int32_t w_org, int32_t h_org, const float *remapX, const float *remapY, const float *G, int32_t g_depth,
                              const float *vignette_inv) {
  Oracle &o = h->o;
  const int w = o.wl[0], hh = o.hl[0];
  if (w_org < 2 || h_org < 2 || (!remapX) != (!remapY) || (G && g_depth != 256 && g_depth != 65536) || (vignette_inv && !G)) return SOSBA_E_ARG;
  if (!remapX && (w_org != w || h_org != hh)) return SOSBA_E_ARG;
  Oracle::Undist &U = o.und;
  U = Oracle::Undist();
  U.wOrg = w_org; U.hOrg = h_org; U.passthrough = !remapX;
  if (remapX) {
    U.remapX.assign(remapX, remapX + (size_t)w * hh); U.remapY.assign(remapY, remapY + (size_t)w * hh);
    for (size_t i = 0; i < U.remapX.size(); i++)   // the bilinear tap reads (x+1, y+1): the reference's map construction guarantees this margin
      if (U.remapX[i] >= 0 && !(U.remapX[i] < w_org - 1 && U.remapY[i] >= 0 && U.remapY[i] < h_org - 1)) return SOSBA_E_ARG;
  }
  if (G) { U.haveG = true; U.gDepth = g_depth; U.G.assign(G, G + g_depth); }
  if (vignette_inv) { U.haveV = true; U.vignetteInv.assign(vignette_inv, vignette_inv + (size_t)w_org * h_org); }
  U.set = true;
  return SOSBA_OK;
}
ORC_API int orc_frame_make_images_raw(orc_handle *h, int32_t slot, const void *raw, int32_t raw_bits, float factor, const float *B, float *image_out) {
  Oracle &o = h->o;
  if (!o.und.set) return SOSBA_E_STATE;
  if (slot < 0 || slot >= (int)o.slots.size() || !raw || (raw_bits != 8 && raw_bits != 16)) return SOSBA_E_ARG;
  if (raw_bits == 8 && o.und.haveG && o.und.gDepth < 256) return SOSBA_E_ARG;
  if (raw_bits == 16 && o.und.haveG && o.und.gDepth != 65536) return SOSBA_E_ARG;
  std::vector<float> img((size_t)o.wl[0] * o.hl[0]);
  undistort_raw(o, raw, raw_bits, factor, img.data());
  if (image_out) memcpy(image_out, img.data(), img.size() * sizeof(float));
  make_images(o, slot, img.data(), B);
  return SOSBA_OK;
}

// ---- next row: loop-closure direct alignment ----------------------------------------
ORC_API int orc_loop_set_points(orc_handle *h, int32_t n, const double *xyz, const float *color) {
  if (n < 0 || (n > 0 && (!xyz || !color))) return SOSBA_E_ARG;
  Oracle &o = h->o;
  o.loop_xyz.resize((size_t)3 * n);
  for (size_t i = 0; i < (size_t)3 * n; i++) o.loop_xyz[i] = (float)xyz[i];
  o.loop_color.assign(color, color + (size_t)n * o.levels);
  return SOSBA_OK;
}
ORC_API int orc_loop_calc_res(orc_handle *h, int32_t lvl, int32_t slot, const double refToNew[12], const float affLL[2], float cutoff, double out6[6],
                              int32_t counts[3]) {
  if (lvl < 0 || lvl >= h->o.levels || slot < 0 || slot >= (int)h->o.slots.size() || !h->o.slots[slot].valid || !refToNew || !affLL) return SOSBA_E_ARG;
  int32_t c[3];
  double o6[6];
  loop_calcRes(h->o, lvl, slot, refToNew, affLL, cutoff, out6 ? out6 : o6, counts ? counts : c);
  return SOSBA_OK;
}
ORC_API int orc_loop_calc_gs(orc_handle *h, int32_t lvl, float a, float b0, double H[64], double b[8]) {
  if (lvl < 0 || lvl >= h->o.levels) return SOSBA_E_ARG;
  tracker_calcGSSSEPose(h->o, lvl, a, b0, H, b);
  return SOSBA_OK;
}

// ---- next row: pixel selection ----------------------------------------------------
ORC_API int orc_pixel_selector_set(orc_handle *h, const uint8_t *random_pattern, int32_t current_potential) {
  if (!random_pattern || current_potential < 1) return SOSBA_E_ARG;
  Oracle &o = h->o;
  o.sel.randomPattern.assign(random_pattern, random_pattern + (size_t)o.wl[0] * o.hl[0]);
  o.sel.currentPotential = current_potential;
  return SOSBA_OK;
}
ORC_API int orc_pixel_select(orc_handle *h, int32_t slot, float density, int32_t recursions_left, float th_factor, int32_t cap, int32_t *n_selected,
                             int32_t *u, int32_t *v, float *type, float *map_out, int32_t *current_potential) {
  Oracle &o = h->o;
  if (o.sel.randomPattern.empty()) return SOSBA_E_STATE;
  if (slot < 0 || slot >= (int)o.slots.size() || !o.slots[slot].valid || o.levels < 3 || !(density > 0)) return SOSBA_E_ARG;
  const int w = o.wl[0], hh = o.hl[0];
  std::vector<float> map((size_t)w * hh);
  const int n = pixel_select(o, slot, density, recursions_left, th_factor, map.data());
  if (n_selected) *n_selected = n;
  if (current_potential) *current_potential = o.sel.currentPotential;
  if (map_out) memcpy(map_out, map.data(), map.size() * sizeof(float));
  if (u || v || type) {
    if (n > cap) return SOSBA_E_ARG;
    int k = 0;
    for (int i = 0; i < w * hh; i++)
      if (map[i] != 0) {
        if (u) u[k] = i % w;
        if (v) v[k] = i / w;
        if (type) type[k] = map[i];
        k++;
      }
  }
  return SOSBA_OK;
}

ORC_API int orc_init_calc_res_and_gs(orc_handle *h, int32_t lvl, int32_t ref_slot, int32_t new_slot, const double refToNew[12], const float aff[2],
                                     const float tlog[3], float alphaW, float alphaK, float couplingWeight, sosba_init_points *pts, float H[64], float b[8],
                                     float Hsc[64], float bsc[8], float res3[3]) {
  Oracle &o = h->o;
  auto bad = [&](int s) { return s < 0 || s >= (int)o.slots.size() || !o.slots[s].valid; };
  if (lvl < 0 || lvl >= o.levels || bad(ref_slot) || bad(new_slot) || !refToNew || !aff || !tlog || !pts || pts->n < 0) return SOSBA_E_ARG;
  init_calcResAndGS(o, lvl, ref_slot, new_slot, refToNew, aff, tlog, alphaW, alphaK, couplingWeight, pts, H, b, Hsc, bsc, res3);
  return SOSBA_OK;
}

// resident pool: the CPU side simply keeps a copy of the arrays
struct OrcPool {
  std::vector<int32_t> host;
  std::vector<float> u, v, color, weights, gradH, eth, idmin, idmax, quality, uv, pixint;
  std::vector<uint8_t> status;
};
static std::map<orc_handle *, OrcPool> g_pools;
static sosba_immature pool_view(OrcPool &p) {
  sosba_immature m = {};
  m.n = (int32_t)p.host.size(); m.host = p.host.data(); m.u = p.u.data(); m.v = p.v.data(); m.color = p.color.data(); m.weights = p.weights.data();
  m.gradH = p.gradH.data(); m.energy_th = p.eth.data(); m.idepth_min = p.idmin.data(); m.idepth_max = p.idmax.data(); m.quality = p.quality.data();
  m.last_trace_status = p.status.data(); m.last_trace_uv = p.uv.data(); m.last_trace_pixel_interval = p.pixint.data();
  return m;
}
ORC_API int orc_immature_pool_set(orc_handle *h, const sosba_immature *pts) {
  if (!pts || pts->n < 0) return SOSBA_E_ARG;
  OrcPool &p = g_pools[h];
  const size_t n = pts->n;
  p.host.assign(pts->host, pts->host + n); p.u.assign(pts->u, pts->u + n); p.v.assign(pts->v, pts->v + n);
  p.color.assign(pts->color, pts->color + 8 * n); p.weights.assign(pts->weights, pts->weights + 8 * n); p.gradH.assign(pts->gradH, pts->gradH + 4 * n);
  p.eth.assign(pts->energy_th, pts->energy_th + n); p.idmin.assign(pts->idepth_min, pts->idepth_min + n); p.idmax.assign(pts->idepth_max, pts->idepth_max + n);
  p.quality.assign(pts->quality, pts->quality + n); p.status.assign(pts->last_trace_status, pts->last_trace_status + n);
  p.uv.assign(pts->last_trace_uv, pts->last_trace_uv + 2 * n); p.pixint.assign(pts->last_trace_pixel_interval, pts->last_trace_pixel_interval + n);
  return SOSBA_OK;
}
ORC_API int orc_immature_pool_trace(orc_handle *h, int32_t frame_slot, int32_t nhosts, const float *KRKi, const float *Kt, const float *aff, int32_t counts[6]) {
  auto it = g_pools.find(h);
  if (it == g_pools.end()) return SOSBA_E_STATE;
  sosba_immature m = pool_view(it->second);
  return orc_trace_immature(h, frame_slot, nhosts, KRKi, Kt, aff, &m, counts);
}
ORC_API int orc_immature_pool_get(orc_handle *h, sosba_immature *pts) {
  auto it = g_pools.find(h);
  if (it == g_pools.end()) return SOSBA_E_STATE;
  OrcPool &p = it->second;
  const size_t n = p.host.size();
  if (!pts || (size_t)pts->n != n) return SOSBA_E_ARG;
  memcpy(pts->idepth_min, p.idmin.data(), 4 * n); memcpy(pts->idepth_max, p.idmax.data(), 4 * n); memcpy(pts->quality, p.quality.data(), 4 * n);
  memcpy(pts->last_trace_status, p.status.data(), n); memcpy(pts->last_trace_uv, p.uv.data(), 8 * n);
  memcpy(pts->last_trace_pixel_interval, p.pixint.data(), 4 * n);
  return SOSBA_OK;
}
