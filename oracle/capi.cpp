// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.hpp).  C entry points for ctypes.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include <thread>

#include "../include/arc_rad.h"
#include "oracle.hpp"

namespace orc {
int oracle_swrad(const ArcDims &d, const ArcSwIn &in, ArcSwOut &out, ArcDebug *dbg, std::string &err);
int oracle_lwrad(const ArcDims &d, const ArcLwIn &in, ArcLwOut &out, ArcDebug *dbg, std::string &err);
}

static std::string g_err;

extern "C" {

int arc_oracle_init(const ArcConfig *cfg, const char *sw_path, const char *lw_path) {
  if (!cfg || !cfg->inline_tables || !sw_path || !lw_path) { g_err = "arc_oracle_init: null argument"; return ARC_ERR_BAD_ARG; }
  return orc::init_tables(cfg->inline_tables, sw_path, lw_path, cfg->cp, cfg->p_top, cfg->kme, g_err);
}

const char *arc_oracle_last_error(void) { return g_err.c_str(); }
int arc_oracle_lw_nlayers(void) { return orc::tables().lw_nlayers; }

int arc_oracle_sw(const ArcDims *d, const ArcSwIn *in, ArcSwOut *out, ArcDebug *dbg) {
  return orc::oracle_swrad(*d, *in, *out, dbg, g_err);
}
int arc_oracle_lw(const ArcDims *d, const ArcLwIn *in, ArcLwOut *out, ArcDebug *dbg) {
  return orc::oracle_lwrad(*d, *in, *out, dbg, g_err);
}

// Threads over j-rows, mimicking radiation_driver's "!$OMP PARALLEL DO" over tiles
// (module_radiation_driver.F:975-978) with a static schedule.  nthreads <= 0: all hardware threads.
} // extern C
template <class F>
static int run_tiles(const ArcDims *d, int nthreads, F fn) {
  int nj = d->jte - d->jts + 1;
  int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > nj) nt = nj;
  std::vector<std::thread> th;
  std::vector<int> rcs(nt, 0);
  std::vector<std::string> errs(nt);
  for (int t = 0; t < nt; t++) {
    th.emplace_back([&, t]() {
      int j0 = d->jts + (int)((long long)nj * t / nt), j1 = d->jts + (int)((long long)nj * (t + 1) / nt) - 1;
      if (j1 < j0) return;
      ArcDims tile = *d; tile.jts = j0; tile.jte = j1;
      rcs[t] = fn(tile, errs[t]);
    });
  }
  for (auto &x : th) x.join();
  for (int t = 0; t < nt; t++) if (rcs[t]) { g_err = errs[t]; return rcs[t]; }
  return 0;
}
extern "C" {
int arc_oracle_sw_omp(const ArcDims *d, const ArcSwIn *in, ArcSwOut *out, int nthreads) {
  // debug column index is tile-relative, so no taps here
  return run_tiles(d, nthreads, [&](const ArcDims &t, std::string &e) { return orc::oracle_swrad(t, *in, *out, nullptr, e); });
}
int arc_oracle_lw_omp(const ArcDims *d, const ArcLwIn *in, ArcLwOut *out, int nthreads) {
  return run_tiles(d, nthreads, [&](const ArcDims &t, std::string &e) { return orc::oracle_lwrad(t, *in, *out, nullptr, e); });
}
int arc_oracle_max_threads(void) { int n = (int)std::thread::hardware_concurrency(); return n > 0 ? n : 1; }

// jp | jt << 8 | jt1 << 12 of setcoef_sw (SW:2854-2887) for n (p hPa, T K) pairs, with glibc logf like the oracle's setcoef
int arc_oracle_pt(const float *p, const float *t, int n, int *packed) {
  const orc::Tables &T = orc::tables();
  if (!T.ready) return ARC_ERR_NOT_INIT;
  const orc::FArr &preflog = T.in.get("sw_preflog"), &tref = T.in.get("sw_tref");
  for (int q = 0; q < n; q++) {
    float plog = logf(p[q]);
    int jp = (int)(36.f - 5 * (plog + 0.04f));
    if (jp < 1) jp = 1; else if (jp > 58) jp = 58;
    int jt = (int)(3.f + (t[q] - tref(jp)) / 15.f);
    if (jt < 1) jt = 1; else if (jt > 4) jt = 4;
    int jt1 = (int)(3.f + (t[q] - tref(jp + 1)) / 15.f);
    if (jt1 < 1) jt1 = 1; else if (jt1 > 4) jt1 = 4;
    packed[q] = jp | (jt << 8) | (jt1 << 12);
  }
  return 0;
}

// cal_cldfra1, module_radiation_driver.F:2886-3122 (scalar restatement; f_q*: 1 true, 0 false, < 0 not PRESENT)
int arc_oracle_cal_cldfra1(const ArcDims *d, const float *QV, const float *QC, const float *QI, const float *QS, int f_qv, int f_qc, int f_qi,
                           int f_qs, const float *t_phy, const float *p_phy, const float *F_ICE_PHY, int mp_physics, float *CLDFRA, int *flag) {
  (void)f_qv;
  const float ALPHA0 = 100.f, GAMMA = 0.49f, QCLDMIN = 1.E-12f, PEXP = 0.25f, RHGRID = 1.0f;
  const float SVP1 = 0.61078f, SVP2 = 17.2693882f, SVPI2 = 21.8745584f, SVP3 = 35.86f, SVPI3 = 7.66f, SVPT0 = 273.15f;
  const float r_d = 287.f, r_v = 461.6f, ep_2 = r_d / r_v;
  const int ni = d->ime - d->ims + 1, nk = d->kme - d->kms + 1;
  for (int j = d->jts; j <= d->jte; j++)
    for (int k = d->kts; k <= d->kte; k++)
      for (int i = d->its; i <= d->ite; i++) {
        const size_t q = (size_t)(i - d->ims) + (size_t)ni * ((size_t)(k - d->kms) + (size_t)nk * (size_t)(j - d->jms));
        float tc = t_phy[q] - SVPT0;
        float esw = 1000.0f * SVP1 * expf(SVP2 * tc / (t_phy[q] - SVP3));
        float esi = 1000.0f * SVP1 * expf(SVPI2 * tc / (t_phy[q] - SVPI3));
        float QVSW = ep_2 * esw / (p_phy[q] - esw);
        float QVSI = ep_2 * esi / (p_phy[q] - esi);
        float weight = 0.f, QCLD = 0.f;
        const bool present = f_qi >= 0 && f_qc >= 0 && f_qs >= 0;
        if (present) {
          const bool F_QI = f_qi > 0, F_QC = f_qc > 0, F_QS = f_qs > 0;
          const float qi = QI ? QI[q] : 0.f, qc = QC ? QC[q] : 0.f, qs = QS ? QS[q] : 0.f;
          if (F_QI && F_QC && F_QS) { QCLD = qi + qc + qs; if (QCLD < QCLDMIN) weight = 0.f; else weight = (qi + qs) / QCLD; }
          if (F_QI && F_QC && !F_QS) { QCLD = qi + qc; if (QCLD < QCLDMIN) weight = 0.f; else weight = qi / QCLD; }
          if (F_QC && !F_QI && !F_QS) {
            QCLD = qc;
            if (QCLD < QCLDMIN) weight = 0.f; else { if (t_phy[q] > 273.15f) weight = 0.f; if (t_phy[q] <= 273.15f) weight = 1.f; }
          }
          if (F_QC && !F_QI && F_QS && F_ICE_PHY) {
            float QIMID = qs, QWMID = qc;
            QCLD = QWMID + QIMID;
            if (QCLD < QCLDMIN) weight = 0.f; else weight = F_ICE_PHY[q];
          }
          if (mp_physics == 5 || mp_physics == 15) {
            float QIMID = qi, QWMID = qc;
            QCLD = QWMID + QIMID;
            if (QCLD < QCLDMIN) weight = 0.f; else { weight = QIMID / QCLD; if (tc < -40.f) weight = 1.f; }
          }
        }
        float QVS_WEIGHT = (1 - weight) * QVSW + weight * QVSI;
        float RHUM = QV[q] / QVS_WEIGHT;
        int fl;
        if (!present || QCLD < QCLDMIN) { CLDFRA[q] = 0.f; fl = 1; }
        else if (RHUM >= RHGRID) { CLDFRA[q] = 1.f; fl = 2; }
        else {
          fl = 3;
          float SUBSAT = std::max(1.E-10f, RHGRID * QVS_WEIGHT - QV[q]);
          float DENOM = powf(SUBSAT, GAMMA);
          float ARG = std::max(-6.9f, -ALPHA0 * QCLD / DENOM);
          RHUM = std::max(1.E-10f, RHUM);
          CLDFRA[q] = powf(RHUM / RHGRID, PEXP) * (1.f - expf(ARG));
          if (CLDFRA[q] < .01f) CLDFRA[q] = 0.f;
        }
        if (flag) flag[q] = fl;
      }
  return 0;
}

// cal_cldfra2, module_radiation_driver.F:2801-2874
int arc_oracle_cal_cldfra2(const ArcDims *d, const float *QC, const float *QI, int F_QC, int F_QI, float *CLDFRA) {
  const float thresh = 1.0e-6f;
  const int ni = d->ime - d->ims + 1, nk = d->kme - d->kms + 1;
  for (int j = d->jts; j <= d->jte; j++)
    for (int k = d->kts; k <= d->kte; k++)
      for (int i = d->its; i <= d->ite; i++) {
        const size_t q = (size_t)(i - d->ims) + (size_t)ni * ((size_t)(k - d->kms) + (size_t)nk * (size_t)(j - d->jms));
        if (F_QI && F_QC) { if (QC[q] + QI[q] > thresh) CLDFRA[q] = 1.f; else CLDFRA[q] = 0.f; }
        else if (F_QC) { if (QC[q] > thresh) CLDFRA[q] = 1.f; else CLDFRA[q] = 0.f; }
        else CLDFRA[q] = 0.f;
      }
  return 0;
}

// ozn_time_int, module_radiation_driver.F:3993-4098 (line by line; ozncyc = .true.)
int arc_oracle_ozn_time_int(const ArcDims *d, int julday, float JULIAN, int levsiz, int num_months, const float *ozmixm, float *ozmixt) {
  (void)julday; (void)num_months;
  static const int date_oz[12] = {16, 45, 75, 105, 136, 166, 197, 228, 258, 289, 319, 350};
  const float daysperyear = 365.f;
  volatile float intJULIAN = JULIAN + 1.0f;
  int IJUL = (int)intJULIAN;
  intJULIAN = intJULIAN - (float)IJUL;
  IJUL = IJUL % 365;
  if (IJUL == 0) IJUL = 365;
  intJULIAN = intJULIAN + IJUL;
  int np1 = 1; bool finddate = false;
  for (int m = 1; m <= 12; m++)
    if (date_oz[m - 1] > intJULIAN && !finddate) { np1 = m; finddate = true; }
  float cdayozp = date_oz[np1 - 1], cdayozm;
  int np, nm;
  if (np1 > 1) { cdayozm = date_oz[np1 - 2]; np = np1; nm = np - 1; }
  else { cdayozm = date_oz[11]; np = np1; nm = 12; }
  volatile float deltat, fact1, fact2;
  if (np1 == 1) {
    deltat = cdayozp + daysperyear - cdayozm;
    if (intJULIAN > cdayozp) { fact1 = (cdayozp + daysperyear - intJULIAN) / deltat; fact2 = (intJULIAN - cdayozm) / deltat; }
    else { fact1 = (cdayozp - intJULIAN) / deltat; fact2 = (intJULIAN + daysperyear - cdayozm) / deltat; }
  } else {
    deltat = cdayozp - cdayozm;
    fact1 = (cdayozp - intJULIAN) / deltat;
    fact2 = (intJULIAN - cdayozm) / deltat;
  }
  const int ni = d->ime - d->ims + 1, nj = d->jme - d->jms + 1;
  const size_t nlev = (size_t)ni * levsiz * nj;
  for (int j = d->jts; j <= d->jte; j++)
    for (int k = 1; k <= levsiz; k++)
      for (int i = d->its; i <= d->ite; i++) {
        const size_t q = (size_t)(i - d->ims) + (size_t)ni * ((size_t)(k - 1) + (size_t)levsiz * (size_t)(j - d->jms));
        const float a = ozmixm[q + nlev * (nm - 1)] * fact1, b = ozmixm[q + nlev * (np - 1)] * fact2;
        ozmixt[q] = a + b;
      }
  return 0;
}

// ozn_p_int, module_radiation_driver.F:4100-4234 (line by line, including the row-wide kkstart / kount bookkeeping)
int arc_oracle_ozn_p_int(const ArcDims *d, const float *p, const float *pin, int levsiz, const float *ozmixt, float *o3vmr) {
  const int its = d->its, ite = d->ite, kts = d->kts, kte = d->kte;
  const int ni = d->ime - d->ims + 1, nk = d->kme - d->kms + 1;
  const int ncol = ite - its + 1, pver = kte - kts + 1;
  auto P3 = [&](int i, int k, int j) { return (size_t)(i - d->ims) + (size_t)ni * ((size_t)(k - d->kms) + (size_t)nk * (size_t)(j - d->jms)); };
  auto OZ = [&](int i, int k, int j) { return ozmixt[(size_t)(i - d->ims) + (size_t)ni * ((size_t)(k - 1) + (size_t)levsiz * (size_t)(j - d->jms))]; };
  auto PIN = [&](int k) { return pin[k - 1]; };
  std::vector<float> pmid((size_t)ncol * (pver + 1));
  std::vector<int> kupper(ncol);
  auto PM = [&](int i, int k) -> float & { return pmid[(size_t)(i - its) + (size_t)ncol * k]; };   // k = 1..pver
  for (int j = d->jts; j <= d->jte; j++) {
    for (int i = its; i <= ite; i++) kupper[i - its] = 1;
    for (int k = kts; k <= kte; k++) {
      const int kk = kte - k + kts;
      for (int i = its; i <= ite; i++) PM(i, kk) = p[P3(i, k, j)];
    }
    for (int k = 1; k <= pver; k++) {
      const int kout = pver - k + 1;
      int kkstart = levsiz;
      for (int i = its; i <= ite; i++) kkstart = std::min(kkstart, kupper[i - its]);
      int kount = 0;
      bool done = false;
      for (int kk = kkstart; kk <= levsiz - 1 && !done; kk++) {
        for (int i = its; i <= ite; i++)
          if (PIN(kk) < PM(i, k) && PM(i, k) <= PIN(kk + 1)) { kupper[i - its] = kk; kount = kount + 1; }
        if (kount == ncol) {
          for (int i = its; i <= ite; i++) {
            const int ku = kupper[i - its];
            const float dpu = PM(i, k) - PIN(ku), dpl = PIN(ku + 1) - PM(i, k);
            const float a = OZ(i, ku, j) * dpl, b = OZ(i, ku + 1, j) * dpu;
            o3vmr[P3(i, kout, j)] = (a + b) / (dpl + dpu);
          }
          done = true;
        }
      }
      if (done) continue;
      for (int i = its; i <= ite; i++) {
        const int ku = kupper[i - its];
        if (PM(i, k) < PIN(1)) { const float a = OZ(i, 1, j) * PM(i, k); o3vmr[P3(i, kout, j)] = a / PIN(1); }
        else if (PM(i, k) > PIN(levsiz)) o3vmr[P3(i, kout, j)] = OZ(i, levsiz, j);
        else {
          const float dpu = PM(i, k) - PIN(ku), dpl = PIN(ku + 1) - PM(i, k);
          const float a = OZ(i, ku, j) * dpl, b = OZ(i, ku + 1, j) * dpu;
          o3vmr[P3(i, kout, j)] = (a + b) / (dpl + dpu);
        }
      }
      if (kount > ncol) return -1;       // 'OZN_P_INT: Bad ozone data: non-monotonicity suspected'
    }
  }
  return 0;
}

// The C library's own logf / expf / powf (what gfortran's LOG / EXP / ** call): which = 0 logf(x), 1 expf(x), 2 powf(x, y).
// Checker for the product's glibc-compatible device functions (csrc/glibc_math.cuh).
int arc_oracle_libm(int which, const float *x, const float *y, int n, float *out) {
  for (int q = 0; q < n; q++) out[q] = which == 0 ? logf(x[q]) : which == 1 ? expf(x[q]) : powf(x[q], y[q]);
  return 0;
}

// radconst + calc_coszen, module_radiation_driver.F:2595-2666 (scalar restatement)
void arc_oracle_radconst(float xtime, float julian, float degrad, float dpd, float *declin, float *solcon) {
  (void)xtime;
  float obecl = 23.5f * degrad, sinob = sinf(obecl), sxlong;
  if (julian >= 80.f) sxlong = dpd * (julian - 80.f); else sxlong = dpd * (julian + 285.f);
  sxlong = sxlong * degrad;
  float arg = sinob * sinf(sxlong);
  *declin = asinf(arg);
  float djul = julian * 360.f / 365.f, rjul = djul * degrad;
  float eccfac = 1.000110f + 0.034221f * cosf(rjul) + 0.001280f * sinf(rjul) + 0.000719f * cosf(2 * rjul) + 0.000077f * sinf(2 * rjul);
  *solcon = 1370.f * eccfac;
}
int arc_oracle_calc_coszen(const ArcDims *d, float julian, float xtime, float gmt, float declin, float degrad, const float *xlon,
                           const float *xlat, float *coszen, float *hrang) {
  const int ni = d->ime - d->ims + 1;
  float da = 6.2831853071795862f * (julian - 1) / 365.f;
  float eot = (0.000075f + 0.001868f * cosf(da) - 0.032077f * sinf(da) - 0.014615f * cosf(2 * da) - 0.04089f * sinf(2 * da)) * (229.18f);
  float xt24 = fmodf(xtime, 1440.f) + eot;
  for (int j = d->jts; j <= d->jte; j++)
    for (int i = d->its; i <= d->ite; i++) {
      size_t q = (size_t)(i - d->ims) + (size_t)ni * (size_t)(j - d->jms);
      float tloctm = gmt + xt24 / 60.f + xlon[q] / 15.f;
      hrang[q] = 15.f * (tloctm - 12.f) * degrad;
      float xxlat = xlat[q] * degrad;
      coszen[q] = sinf(xxlat) * sinf(declin) + cosf(xxlat) * cosf(declin) * cosf(hrang[q]);
    }
  return 0;
}

// Reduced-table taps so tests can compare the product's init against this restatement.
// kind: 0 = SW, 1 = LW; name: "absa","absb","selfref","forref","sfluxref"/"fracrefa","fracrefb", ...
int arc_oracle_table(int kind, int band /*1-based within kind*/, const char *name, float *buf, int cap) {
  const orc::Tables &T = orc::tables();
  const orc::FArr *a = nullptr;
  std::string n(name);
  if (kind == 0) {
    const orc::SwBand &B = T.sw[band - 1];
    if (n == "absa") a = &B.absa; else if (n == "absb") a = &B.absb; else if (n == "selfref") a = &B.selfref;
    else if (n == "forref") a = &B.forref; else if (n == "sfluxref") a = &B.sfluxref; else if (n == "raylg") a = &B.raylg;
    else if (n == "rayla") a = &B.rayla; else if (n == "raylb") a = &B.raylb; else if (n == "abso3a") a = &B.abso3a;
    else if (n == "abso3b") a = &B.abso3b; else if (n == "absch4") a = &B.absch4; else if (n == "absh2o") a = &B.absh2o;
    else if (n == "absco2") a = &B.absco2;
  } else {
    const orc::LwBand &B = T.lw[band - 1];
    if (n == "absa") a = &B.absa; else if (n == "absb") a = &B.absb; else if (n == "selfref") a = &B.selfref;
    else if (n == "forref") a = &B.forref; else if (n == "fracrefa") a = &B.fracrefa; else if (n == "fracrefb") a = &B.fracrefb;
    else if (n == "ka_mn2") a = &B.ka_mn2; else if (n == "kb_mn2") a = &B.kb_mn2; else if (n == "ka_mn2o") a = &B.ka_mn2o;
    else if (n == "kb_mn2o") a = &B.kb_mn2o; else if (n == "ka_mo3") a = &B.ka_mo3; else if (n == "kb_mo3") a = &B.kb_mo3;
    else if (n == "ka_mco2") a = &B.ka_mco2; else if (n == "kb_mco2") a = &B.kb_mco2; else if (n == "ka_mco") a = &B.ka_mco;
    else if (n == "ka_mo2") a = &B.ka_mo2; else if (n == "kb_mo2") a = &B.kb_mo2; else if (n == "ccl4") a = &B.ccl4;
    else if (n == "cfc11adj") a = &B.cfc11adj; else if (n == "cfc12") a = &B.cfc12; else if (n == "cfc22adj") a = &B.cfc22adj;
  }
  if (!a) return -1;
  int cnt = (int)a->size();
  if (buf && cap >= cnt) memcpy(buf, a->v.data(), (size_t)cnt * 4);
  return cnt;
}

}  // extern "C"
