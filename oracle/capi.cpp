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

// rslf / rsif of module_mp_thompson (WRF v3.9.1 phys/module_mp_thompson.F) - NOT in the reference repository: the published
// Flatau et al. (1992) 8th-order polynomials of the saturation vapour pressure over water / ice as the Thompson scheme
// codes them (X = MAX(-80, T - 273.16); e capped at 15 % of p; r = 0.622 e / (p - e)).  PARITY UNPINNED for these two
// functions; cal_cldfra3 below is a line-by-line restatement of what IS in the reference.
static float thompson_rslf(float P, float T) {
  const float C0 = .611583699E03f, C1 = .444606896E02f, C2 = .143177157E01f, C3 = .264224321E-1f, C4 = .299291081E-3f, C5 = .203154182E-5f,
              C6 = .702620698E-8f, C7 = .379534310E-11f, C8 = -.321582393E-13f;
  const float X = std::max(-80.f, T - 273.16f);
  float ESL = C0 + X * (C1 + X * (C2 + X * (C3 + X * (C4 + X * (C5 + X * (C6 + X * (C7 + X * C8)))))));
  ESL = std::min(ESL, P * 0.15f);
  return .622f * ESL / (P - ESL);
}
static float thompson_rsif(float P, float T) {
  const float C0 = .609868993E03f, C1 = .499320233E02f, C2 = .184672631E01f, C3 = .402737184E-1f, C4 = .565392987E-3f, C5 = .521693933E-5f,
              C6 = .307839583E-7f, C7 = .105785160E-9f, C8 = .161444444E-12f;
  const float X = std::max(-80.f, T - 273.16f);
  float ESI = C0 + X * (C1 + X * (C2 + X * (C3 + X * (C4 + X * (C5 + X * (C6 + X * (C7 + X * C8)))))));
  ESI = std::min(ESI, P * 0.15f);
  return .622f * ESI / (P - ESI);
}

namespace {
// 1-based views (kts..kte) as the Fortran declares its 1-D work arrays
struct Col1D {
  std::vector<float> v; int kts;
  Col1D(int kts_, int kte) : v(kte - kts_ + 1, 0.f), kts(kts_) {}
  float &operator()(int k) { return v[k - kts]; }
};
// adjust_cloudIce, module_radiation_driver.F:3472-3512 (kmid and iwp_exists are computed there and never used)
void adjust_cloudIce(Col1D &cfr, Col1D &qi, Col1D &qs, Col1D &qvs, Col1D &T, Col1D &Rho, Col1D &dz, float entr, int k1, int k2) {
  (void)qs; (void)Rho;
  float tdz = 0.f;
  for (int k = k1; k <= k2; k++) tdz = tdz + dz(k);
  const float max_iwc = std::fabs(qvs(k2 - 1) - qvs(k1));
  float this_dz = 0.0f;
  for (int k = k1; k <= k2; k++) {
    if (k == k1) this_dz = this_dz + 0.5f * dz(k); else this_dz = this_dz + dz(k);
    const float this_iwc = max_iwc * this_dz / tdz;
    const float iwc = std::max(1.E-6f, this_iwc * (1.f - entr));
    if (cfr(k) > 0.01f && cfr(k) < 0.99f && T(k) >= 203.16f) qi(k) = qi(k) + 0.1f * cfr(k) * iwc;
    else if (qi(k) < 1.E-5f && cfr(k) >= 0.99f && T(k) >= 203.16f) qi(k) = qi(k) + 0.01f * iwc;
  }
}
// adjust_cloudH2O, module_radiation_driver.F:3516-3555
void adjust_cloudH2O(Col1D &cfr, Col1D &qc, Col1D &qvs, Col1D &T, Col1D &dz, float entr, int k1, int k2) {
  float tdz = 0.f;
  for (int k = k1; k <= k2; k++) tdz = tdz + dz(k);
  const float max_lwc = std::fabs(qvs(k2 - 1) - qvs(k1));
  float this_dz = 0.0f;
  for (int k = k1; k <= k2; k++) {
    if (k == k1) this_dz = this_dz + 0.5f * dz(k); else this_dz = this_dz + dz(k);
    const float this_lwc = max_lwc * this_dz / tdz;
    const float lwc = std::max(1.E-6f, this_lwc * (1.f - entr));
    if (cfr(k) > 0.01f && cfr(k) < 0.99f && T(k) < 298.16f && T(k) >= 253.16f) qc(k) = qc(k) + cfr(k) * cfr(k) * lwc;
    else if (cfr(k) >= 0.99f && qc(k) < 1.E-5f && T(k) < 298.16f && T(k) >= 253.16f) qc(k) = qc(k) + 0.1f * lwc;
  }
}
// adjust_cloudFinal, module_radiation_driver.F:3562-3599
void adjust_cloudFinal(Col1D &cfr, Col1D &qc, Col1D &qi, Col1D &Rho, Col1D &dz, int kts, int k_tropo) {
  float lwp = 0.f, iwp = 0.f;
  for (int k = kts; k <= k_tropo; k++)
    if (cfr(k) > 0.0f) { lwp = lwp + qc(k) * Rho(k) * dz(k); iwp = iwp + qi(k) * Rho(k) * dz(k); }
  if (lwp > 1.5f) { const float xfac = 1.f / lwp; for (int k = kts; k <= k_tropo; k++) if (cfr(k) > 0.01f && cfr(k) < 0.99f) qc(k) = qc(k) * xfac; }
  if (iwp > 1.5f) { const float xfac = 1.f / iwp; for (int k = kts; k <= k_tropo; k++) if (cfr(k) > 0.01f && cfr(k) < 0.99f) qi(k) = qi(k) * xfac; }
}
// find_cloudLayers, module_radiation_driver.F:3281-3468.  The second search starts at k_m12C + 2, which the Fortran does not
// bound by kte; a level above kte is read as cloud-free here (the product does the same).
void find_cloudLayers(Col1D &qvs1d, Col1D &cfr1d, Col1D &T1d, Col1D &P1d, Col1D &R1d, float entrmnt, Col1D &qc1d, Col1D &qi1d, Col1D &qs1d,
                      int kts, int kte) {
  Col1D theta(kts, kte), dz(kts, kte);
  int k, k2, k_tropo, k_m12C = 0, k_m40C = 0, k_cldb, k_cldt, kbot;
  bool in_cloud;
  for (k = kte; k >= kts; k--) {
    theta(k) = T1d(k) * powf(100000.0f / P1d(k), 287.05f / 1004.f);
    if (T1d(k) - 273.16f > -40.0f && P1d(k) > 7000.0f) k_m40C = std::max(k_m40C, k);
    if (T1d(k) - 273.16f > -12.0f && P1d(k) > 10000.0f) k_m12C = std::max(k_m12C, k);
  }
  if (k_m40C <= kts) k_m40C = kts;
  if (k_m12C <= kts) k_m12C = kts;
  float Z2 = 44307.692f * (1.0f - powf(P1d(kte) / 101325.f, 0.190f));
  for (k = kte - 1; k >= kts; k--) {
    const float Z1 = 44307.692f * (1.0f - powf(P1d(k) / 101325.f, 0.190f));
    dz(k + 1) = Z2 - Z1;
    Z2 = Z1;
  }
  dz(kts) = dz(kts + 1);
  for (k = kte - 3; k >= kts; k--) {
    const float theta1 = theta(k), theta2 = theta(k + 2);
    const float ht1 = 44307.692f * (1.0f - powf(P1d(k) / 101325.f, 0.190f));
    const float ht2 = 44307.692f * (1.0f - powf(P1d(k + 2) / 101325.f, 0.190f));
    if ((((theta2 - theta1) / (ht2 - ht1)) < 10.f / 1500.f) && (ht1 < 19000.f) && (ht1 > 4000.f)) break;
  }
  k_tropo = std::max(kts + 2, k + 2);
  for (k = k_tropo + 1; k <= kte; k++)
    if (cfr1d(k) > 0.0f && cfr1d(k) < 0.999f) cfr1d(k) = 0.f;
  kbot = kts + 2;
  for (k = kbot; k <= k_m12C; k++)
    if ((theta(k) - theta(k - 1)) > 0.05E-3f * dz(k)) break;
  kbot = std::max(kts + 1, k - 2);
  for (k = kts; k <= kbot; k++)
    if (cfr1d(k) > 0.0f && cfr1d(k) < 0.999f) cfr1d(k) = 0.f;
  auto CFR = [&](int kk) { return kk <= kte ? cfr1d(kk) : 0.f; };
  k_cldb = k_tropo;
  in_cloud = false;
  k = k_tropo;
  while (!in_cloud && k > k_m12C) {
    k_cldt = 0;
    if (CFR(k) >= 0.01f) { in_cloud = true; k_cldt = std::max(k_cldt, k); }
    if (in_cloud) {
      for (k2 = k_cldt - 1; k2 >= k_m12C; k2--)
        if (cfr1d(k2) < 0.01f || k2 == k_m12C) { k_cldb = k2 + 1; break; }
      in_cloud = false;
    }
    if ((k_cldt - k_cldb + 1) >= 2) {
      adjust_cloudIce(cfr1d, qi1d, qs1d, qvs1d, T1d, R1d, dz, entrmnt, k_cldb, k_cldt);
      k = k_cldb;
    } else {
      if (cfr1d(k_cldb) > 0.f && qi1d(k_cldb) < 1.E-6f) qi1d(k_cldb) = 1.E-5f * cfr1d(k_cldb);
    }
    k = k - 1;
  }
  k_cldb = k_tropo;
  in_cloud = false;
  k = k_m12C + 2;
  while (!in_cloud && k > kbot) {
    k_cldt = 0;
    if (CFR(k) >= 0.01f) { in_cloud = true; k_cldt = std::max(k_cldt, k); }
    if (in_cloud) {
      for (k2 = k_cldt - 1; k2 >= kbot; k2--)
        if (cfr1d(k2) < 0.01f || k2 == kbot) { k_cldb = k2 + 1; break; }
      in_cloud = false;
    }
    if ((k_cldt - k_cldb + 1) >= 2) {
      adjust_cloudH2O(cfr1d, qc1d, qvs1d, T1d, dz, entrmnt, k_cldb, k_cldt);
      k = k_cldb;
    } else {
      if (cfr1d(k_cldb) > 0.f && qc1d(k_cldb) < 1.E-6f) qc1d(k_cldb) = 1.E-5f * cfr1d(k_cldb);
    }
    k = k - 1;
  }
  adjust_cloudFinal(cfr1d, qc1d, qi1d, R1d, dz, kts, k_tropo);
}
}  // namespace

// cal_cldfra3, module_radiation_driver.F:3140-3274 (icloud = 3; G. Thompson's Sundqvist-type scheme); qc and qi are INOUT
int arc_oracle_cal_cldfra3(const ArcDims *d, float *CLDFRA, const float *qv, float *qc, float *qi, const float *qs, const float *p, const float *t,
                           const float *rho, const float *XLAND, float gridkm) {
  const int ni = d->ime - d->ims + 1, nk = d->kme - d->kms + 1;
  const int kts = d->kts, kte = d->kte;
  if (kte - kts < 4) return -1;
  auto Q3 = [&](int i, int k, int j) { return (size_t)(i - d->ims) + (size_t)ni * ((size_t)(k - d->kms) + (size_t)nk * (size_t)(j - d->jms)); };
  auto Q2 = [&](int i, int j) { return (size_t)(i - d->ims) + (size_t)ni * (size_t)(j - d->jms); };
  const float RH_00L = 0.7f + sqrtf(1.f / (25.0f + gridkm * gridkm * gridkm));
  const float RH_00O = 0.81f + sqrtf(1.f / (50.0f + gridkm * gridkm * gridkm));
  std::vector<float> qvsat((size_t)ni * nk * (d->jme - d->jms + 1), 0.f);
  for (int j = d->jts; j <= d->jte; j++)
    for (int k = kts; k <= kte; k++)
      for (int i = d->its; i <= d->ite; i++) {
        const size_t q = Q3(i, k, j);
        CLDFRA[q] = 0.0f;
        if (qc[q] > 1.E-6f || qi[q] >= 1.E-7f || qs[q] > 1.E-5f) {
          CLDFRA[q] = 1.0f;
          qvsat[q] = qv[q];
        } else {
          const float TK = t[q], TC = TK - 273.16f;
          const float qvsw = thompson_rslf(p[q], TK), qvsi = thompson_rsif(p[q], TK);
          if (TC >= -12.0f) qvsat[q] = qvsw;
          else if (TC < -20.0f) qvsat[q] = qvsi;
          else qvsat[q] = qvsw - (qvsw - qvsi) * (-12.0f - TC) / (-12.0f + 20.f);
          float RHUM = std::max(0.01f, std::min(qv[q] / qvsat[q], 0.9999f));
          const float RH_00 = (XLAND[Q2(i, j)] - 1.5f) > 0.f ? RH_00O : RH_00L;
          if (TC >= -12.0f) {
            RHUM = std::min(0.999f, RHUM);
            CLDFRA[q] = std::max(0.0f, 1.0f - sqrtf((1.0f - RHUM) / (1.f - RH_00)));
          } else if (TC < -12.f && TC > -70.f && RHUM > RH_00L) {
            RHUM = std::max(0.01f, std::min(qv[q] / qvsat[q], 1.0f - 1.E-6f));
            CLDFRA[q] = std::max(0.f, 1.0f - sqrtf((1.0f - RHUM) / (1.0f - RH_00L)));
          }
          CLDFRA[q] = std::min(0.90f, CLDFRA[q]);
        }
      }
  Col1D qvs1d(kts, kte), cfr1d(kts, kte), T1d(kts, kte), P1d(kts, kte), R1d(kts, kte), qc1d(kts, kte), qi1d(kts, kte), qs1d(kts, kte);
  for (int j = d->jts; j <= d->jte; j++)
    for (int i = d->its; i <= d->ite; i++) {
      const float entrmnt = 0.5f;
      for (int k = kts; k <= kte; k++) {
        const size_t q = Q3(i, k, j);
        qvs1d(k) = qvsat[q]; cfr1d(k) = CLDFRA[q]; T1d(k) = t[q]; P1d(k) = p[q]; R1d(k) = rho[q];
        qc1d(k) = qc[q]; qi1d(k) = qi[q]; qs1d(k) = qs[q];
      }
      find_cloudLayers(qvs1d, cfr1d, T1d, P1d, R1d, entrmnt, qc1d, qi1d, qs1d, kts, kte);
      for (int k = kts; k <= kte; k++) {
        const size_t q = Q3(i, k, j);
        CLDFRA[q] = cfr1d(k); qc[q] = qc1d(k); qi[q] = qi1d(k);
      }
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

// aer_time_int, module_radiation_driver.F:4236-4343: ozn_time_int per aerosol type
int arc_oracle_aer_time_int(const ArcDims *d, int julday, float JULIAN, int levsiz, int num_months, int no_src, const float *aerodm, float *aerodt) {
  const int ni = d->ime - d->ims + 1, nj = d->jme - d->jms + 1;
  const size_t nlev = (size_t)ni * levsiz * nj;
  for (int s = 0; s < no_src; s++) {
    int rc = arc_oracle_ozn_time_int(d, julday, JULIAN, levsiz, num_months, aerodm + nlev * (size_t)num_months * s, aerodt + nlev * (size_t)s);
    if (rc) return rc;
  }
  return 0;
}

// aer_p_int, module_radiation_driver.F:4345-4506 (line by line)
int arc_oracle_aer_p_int(const ArcDims *d, const float *p, const float *pin, int levsiz, const float *aerodt, float *aerod, int no_src,
                         const float *pf, float *totaod) {
  const int its = d->its, ite = d->ite, kts = d->kts, kte = d->kte;
  const int ni = d->ime - d->ims + 1, nk = d->kme - d->kms + 1, nj = d->jme - d->jms + 1;
  const int ncol = ite - its + 1, pver = kte - kts + 1;
  const size_t n3 = (size_t)ni * nk * nj, nlev = (size_t)ni * levsiz * nj;
  auto P3 = [&](int i, int k, int j) { return (size_t)(i - d->ims) + (size_t)ni * ((size_t)(k - d->kms) + (size_t)nk * (size_t)(j - d->jms)); };
  auto PIN = [&](int k) { return pin[k - 1]; };
  std::vector<float> pmid((size_t)ncol * (pver + 1));
  std::vector<int> kupper(ncol);
  auto PM = [&](int i, int k) -> float & { return pmid[(size_t)(i - its) + (size_t)ncol * k]; };
  for (int s = 1; s <= no_src; s++) {
    auto AT = [&](int i, int k, int j) { return aerodt[(size_t)(i - d->ims) + (size_t)ni * ((size_t)(k - 1) + (size_t)levsiz * (size_t)(j - d->jms)) + nlev * (size_t)(s - 1)]; };
    float *ao = aerod + n3 * (size_t)(s - 1);
    for (int j = d->jts; j <= d->jte; j++) {
      for (int i = its; i <= ite; i++) kupper[i - its] = 1;
      for (int k = kts; k <= kte; k++) {
        const int kk = kte - k + kts;
        for (int i = its; i <= ite; i++) PM(i, kk) = p[P3(i, k, j)] * 0.01f;
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
              const float dpm = pf[P3(i, kout, j)] - pf[P3(i, kout + 1, j)];
              const float a = AT(i, ku, j) * dpl, b = AT(i, ku + 1, j) * dpu;
              float v = (a + b) / (dpl + dpu);
              v = v * dpm;
              ao[P3(i, kout, j)] = v;
            }
            done = true;
          }
        }
        if (done) continue;
        for (int i = its; i <= ite; i++) {
          const int ku = kupper[i - its];
          const float dpm = pf[P3(i, kout, j)] - pf[P3(i, kout + 1, j)];
          float v;
          if (PM(i, k) < PIN(1)) { const float a = AT(i, 1, j) * PM(i, k); v = a / PIN(1); }
          else if (PM(i, k) > PIN(levsiz)) v = AT(i, levsiz, j);
          else {
            const float dpu = PM(i, k) - PIN(ku), dpl = PIN(ku + 1) - PM(i, k);
            const float a = AT(i, ku, j) * dpl, b = AT(i, ku + 1, j) * dpu;
            v = (a + b) / (dpl + dpu);
          }
          v = v * dpm;
          ao[P3(i, kout, j)] = v;
        }
        if (kount > ncol) return -1;       // 'AER_P_INT: Bad aerosol data: non-monotonicity suspected'
      }
    }
  }
  for (int j = d->jts; j <= d->jte; j++)
    for (int i = its; i <= ite; i++) totaod[(size_t)(i - d->ims) + (size_t)ni * (size_t)(j - d->jms)] = 0.f;
  for (int s = 1; s <= no_src; s++)
    for (int j = d->jts; j <= d->jte; j++)
      for (int k = 1; k <= pver; k++)
        for (int i = its; i <= ite; i++) {
          float &t = totaod[(size_t)(i - d->ims) + (size_t)ni * (size_t)(j - d->jms)];
          t = t + aerod[P3(i, k, j) + n3 * (size_t)(s - 1)];
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
