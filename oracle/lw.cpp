// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// Longwave: restates, in the reference's evaluation order and one column at a time,
//   cldprmc  module_ra_rrtmg_lw.F:2653-2914     rtrnmc  2974-3410
//   setcoef  3444-3809                          taumol  4712-7828 (taugb1..16)
//   rrtmg_lw 10578-11064    inatm 11067-11403   RRTMG_LWRAD 11451-12700
#include <cmath>
#include <cstdio>
#include <cstring>

#include "adapter_common.hpp"
#include "oracle.hpp"

namespace orc {

namespace {

struct LwCoef {
  int laytrop, jp[MXLAY], jt[MXLAY], jt1[MXLAY], indself[MXLAY], indfor[MXLAY], indminor[MXLAY];
  float planklay[NBLW + 1][MXLAY], planklev[NBLW + 1][MXLAY + 1], plankbnd[NBLW + 1];
  float colh2o[MXLAY], colco2[MXLAY], colo3[MXLAY], coln2o[MXLAY], colco[MXLAY], colch4[MXLAY], colo2[MXLAY], colbrd[MXLAY];
  float fac00[MXLAY], fac01[MXLAY], fac10[MXLAY], fac11[MXLAY];
  float rat_h2oco2[MXLAY], rat_h2oco2_1[MXLAY], rat_h2oo3[MXLAY], rat_h2oo3_1[MXLAY], rat_h2on2o[MXLAY], rat_h2on2o_1[MXLAY],
      rat_h2och4[MXLAY], rat_h2och4_1[MXLAY], rat_n2oco2[MXLAY], rat_n2oco2_1[MXLAY], rat_o3co2[MXLAY], rat_o3co2_1[MXLAY];
  float selffac[MXLAY], selffrac[MXLAY], forfac[MXLAY], forfrac[MXLAY], minorfrac[MXLAY], scaleminor[MXLAY], scaleminorn2[MXLAY];
};

// ---- setcoef LW:3444-3809 (istart = 1) -------------------------------------------------------------
void setcoef_lw(const Tables &T, int nlayers, const float *pavel, const float *tavel, const float *tz, float tbound,
                const float *semiss, const float *coldry, const float (*wkl)[MXLAY], const float *wbroad, LwCoef &c) {
  const FArr &preflog = T.in.get("lw_preflog"), &tref = T.in.get("lw_tref"), &chi_mls = T.in.get("lw_chi_mls"),
             &totplnk = T.in.get("lw_totplnk");
  const float stpfac = 296.f / 1013.f;
  int indbound = (int)(tbound - 159.f);
  if (indbound < 1) indbound = 1; else if (indbound > 180) indbound = 180;
  float tbndfrac = tbound - 159.f - (float)indbound;
  int indlev0 = (int)(tz[0] - 159.f);
  if (indlev0 < 1) indlev0 = 1; else if (indlev0 > 180) indlev0 = 180;
  float t0frac = tz[0] - 159.f - (float)indlev0;
  c.laytrop = 0;
  for (int lay = 1; lay <= nlayers; lay++) {
    int indlay = (int)(tavel[lay] - 159.f);
    if (indlay < 1) indlay = 1; else if (indlay > 180) indlay = 180;
    float tlayfrac = tavel[lay] - 159.f - (float)indlay;
    int indlev = (int)(tz[lay] - 159.f);
    if (indlev < 1) indlev = 1; else if (indlev > 180) indlev = 180;
    float tlevfrac = tz[lay] - 159.f - (float)indlev;
    for (int iband = 1; iband <= 16; iband++) {   // bands 1-15 and the istart != 16 branch for band 16 are identical
      float dbdtlev, dbdtlay;
      if (lay == 1) {
        dbdtlev = totplnk(indbound + 1, iband) - totplnk(indbound, iband);
        c.plankbnd[iband] = semiss[iband] * (totplnk(indbound, iband) + tbndfrac * dbdtlev);
        dbdtlev = totplnk(indlev0 + 1, iband) - totplnk(indlev0, iband);
        c.planklev[iband][0] = totplnk(indlev0, iband) + t0frac * dbdtlev;
      }
      dbdtlev = totplnk(indlev + 1, iband) - totplnk(indlev, iband);
      dbdtlay = totplnk(indlay + 1, iband) - totplnk(indlay, iband);
      c.planklay[iband][lay] = totplnk(indlay, iband) + tlayfrac * dbdtlay;
      c.planklev[iband][lay] = totplnk(indlev, iband) + tlevfrac * dbdtlev;
    }
    float plog = logf(pavel[lay]);
    c.jp[lay] = (int)(36.f - 5 * (plog + 0.04f));
    if (c.jp[lay] < 1) c.jp[lay] = 1; else if (c.jp[lay] > 58) c.jp[lay] = 58;
    int jp1 = c.jp[lay] + 1;
    float fp = 5.f * (preflog(c.jp[lay]) - plog);
    c.jt[lay] = (int)(3.f + (tavel[lay] - tref(c.jp[lay])) / 15.f);
    if (c.jt[lay] < 1) c.jt[lay] = 1; else if (c.jt[lay] > 4) c.jt[lay] = 4;
    float ft = ((tavel[lay] - tref(c.jp[lay])) / 15.f) - (float)(c.jt[lay] - 3);
    c.jt1[lay] = (int)(3.f + (tavel[lay] - tref(jp1)) / 15.f);
    if (c.jt1[lay] < 1) c.jt1[lay] = 1; else if (c.jt1[lay] > 4) c.jt1[lay] = 4;
    float ft1 = ((tavel[lay] - tref(jp1)) / 15.f) - (float)(c.jt1[lay] - 3);
    float water = wkl[1][lay] / coldry[lay];
    float scalefac = pavel[lay] * stpfac / tavel[lay];
    float factor;
    const int jpl = c.jp[lay];
    if (!(plog <= 4.56f)) {
      c.laytrop = c.laytrop + 1;
      c.forfac[lay] = scalefac / (1.f + water);
      factor = (332.0f - tavel[lay]) / 36.0f;
      c.indfor[lay] = std::min(2, std::max(1, (int)factor));
      c.forfrac[lay] = factor - (float)c.indfor[lay];
      c.selffac[lay] = water * c.forfac[lay];
      factor = (tavel[lay] - 188.0f) / 7.2f;
      c.indself[lay] = std::min(9, std::max(1, (int)factor - 7));
      c.selffrac[lay] = factor - (float)(c.indself[lay] + 7);
      c.scaleminor[lay] = pavel[lay] / tavel[lay];
      c.scaleminorn2[lay] = (pavel[lay] / tavel[lay]) * (wbroad[lay] / (coldry[lay] + wkl[1][lay]));
      factor = (tavel[lay] - 180.8f) / 7.2f;
      c.indminor[lay] = std::min(18, std::max(1, (int)factor));
      c.minorfrac[lay] = factor - (float)c.indminor[lay];
      c.rat_h2oco2[lay] = chi_mls(1, jpl) / chi_mls(2, jpl);
      c.rat_h2oco2_1[lay] = chi_mls(1, jpl + 1) / chi_mls(2, jpl + 1);
      c.rat_h2oo3[lay] = chi_mls(1, jpl) / chi_mls(3, jpl);
      c.rat_h2oo3_1[lay] = chi_mls(1, jpl + 1) / chi_mls(3, jpl + 1);
      c.rat_h2on2o[lay] = chi_mls(1, jpl) / chi_mls(4, jpl);
      c.rat_h2on2o_1[lay] = chi_mls(1, jpl + 1) / chi_mls(4, jpl + 1);
      c.rat_h2och4[lay] = chi_mls(1, jpl) / chi_mls(6, jpl);
      c.rat_h2och4_1[lay] = chi_mls(1, jpl + 1) / chi_mls(6, jpl + 1);
      c.rat_n2oco2[lay] = chi_mls(4, jpl) / chi_mls(2, jpl);
      c.rat_n2oco2_1[lay] = chi_mls(4, jpl + 1) / chi_mls(2, jpl + 1);
      c.rat_o3co2[lay] = 0.f; c.rat_o3co2_1[lay] = 0.f;
    } else {
      c.forfac[lay] = scalefac / (1.f + water);
      factor = (tavel[lay] - 188.0f) / 36.0f;
      c.indfor[lay] = 3;
      c.forfrac[lay] = factor - 1.0f;
      c.selffac[lay] = water * c.forfac[lay];
      c.indself[lay] = 0; c.selffrac[lay] = 0.f;   // not set by the reference aloft (unused)
      c.scaleminor[lay] = pavel[lay] / tavel[lay];
      c.scaleminorn2[lay] = (pavel[lay] / tavel[lay]) * (wbroad[lay] / (coldry[lay] + wkl[1][lay]));
      factor = (tavel[lay] - 180.8f) / 7.2f;
      c.indminor[lay] = std::min(18, std::max(1, (int)factor));
      c.minorfrac[lay] = factor - (float)c.indminor[lay];
      c.rat_h2oco2[lay] = chi_mls(1, jpl) / chi_mls(2, jpl);
      c.rat_h2oco2_1[lay] = chi_mls(1, jpl + 1) / chi_mls(2, jpl + 1);
      c.rat_o3co2[lay] = chi_mls(3, jpl) / chi_mls(2, jpl);
      c.rat_o3co2_1[lay] = chi_mls(3, jpl + 1) / chi_mls(2, jpl + 1);
    }
    c.colh2o[lay] = 1.e-20f * wkl[1][lay];
    c.colco2[lay] = 1.e-20f * wkl[2][lay];
    c.colo3[lay] = 1.e-20f * wkl[3][lay];
    c.coln2o[lay] = 1.e-20f * wkl[4][lay];
    c.colco[lay] = 1.e-20f * wkl[5][lay];
    c.colch4[lay] = 1.e-20f * wkl[6][lay];
    c.colo2[lay] = 1.e-20f * wkl[7][lay];
    if (c.colco2[lay] == 0.f) c.colco2[lay] = 1.e-32f * coldry[lay];
    if (c.colo3[lay] == 0.f) c.colo3[lay] = 1.e-32f * coldry[lay];
    if (c.coln2o[lay] == 0.f) c.coln2o[lay] = 1.e-32f * coldry[lay];
    if (c.colco[lay] == 0.f) c.colco[lay] = 1.e-32f * coldry[lay];
    if (c.colch4[lay] == 0.f) c.colch4[lay] = 1.e-32f * coldry[lay];
    c.colbrd[lay] = 1.e-20f * wbroad[lay];
    float compfp = 1.f - fp;
    c.fac10[lay] = compfp * ft;
    c.fac00[lay] = compfp * (1.f - ft);
    c.fac11[lay] = fp * ft1;
    c.fac01[lay] = fp * (1.f - ft1);
    c.selffac[lay] = c.colh2o[lay] * c.selffac[lay];
    c.forfac[lay] = c.colh2o[lay] * c.forfac[lay];
  }
}

struct LwTau { float taug[NGLW + 1][MXLAY], fracs[NGLW + 1][MXLAY]; };

inline float fmod1(float x) { return x - (float)(int)x; }
inline float pow4(float p) { float p2 = p * p; return p2 * p2; }  // p**4 (gfortran expands integer powers by repeated squaring)

// ---- taumol LW:4712-7828 ---------------------------------------------------------------------------
void taumol_lw(const Tables &T, int nlayers, const float *pavel, const float (*wx)[MXLAY], const float *coldry, const LwCoef &c,
               LwTau &o) {
  const int laytrop = c.laytrop;
  const float oneminus = T.oneminus;
  const FArr &chi_mls = T.in.get("lw_chi_mls");
  auto selfk = [&](const LwBand &B, int lay, int ig) {
    int inds = c.indself[lay];
    return c.selffac[lay] * (B.selfref(inds, ig) + c.selffrac[lay] * (B.selfref(inds + 1, ig) - B.selfref(inds, ig)));
  };
  auto fork = [&](const LwBand &B, int lay, int ig) {
    int indf = c.indfor[lay];
    return c.forfac[lay] * (B.forref(indf, ig) + c.forfrac[lay] * (B.forref(indf + 1, ig) - B.forref(indf, ig)));
  };
  auto minor1 = [&](const FArr &km, int lay, int ig) {  // T-interpolated minor (19,ng)
    int indm = c.indminor[lay];
    return km(indm, ig) + c.minorfrac[lay] * (km(indm + 1, ig) - km(indm, ig));
  };
  auto minor2 = [&](const FArr &km, int lay, int jm, float fm, int ig) {  // (eta,T)-interpolated minor (n,19,ng)
    int indm = c.indminor[lay];
    float m1 = km(jm, indm, ig) + fm * (km(jm + 1, indm, ig) - km(jm, indm, ig));
    float m2 = km(jm, indm + 1, ig) + fm * (km(jm + 1, indm + 1, ig) - km(jm, indm + 1, ig));
    return m1 + c.minorfrac[lay] * (m2 - m1);
  };
  auto k4 = [&](const FArr &ab, int lay, int ind0, int ind1, int ig) {
    return c.fac00[lay] * ab(ind0, ig) + c.fac10[lay] * ab(ind0 + 1, ig) + c.fac01[lay] * ab(ind1, ig) + c.fac11[lay] * ab(ind1 + 1, ig);
  };
  auto i0a = [&](int lay, int nsp) { return ((c.jp[lay] - 1) * 5 + (c.jt[lay] - 1)) * nsp + 1; };
  auto i1a = [&](int lay, int nsp) { return (c.jp[lay] * 5 + (c.jt1[lay] - 1)) * nsp + 1; };
  auto i0b = [&](int lay, int nsp) { return ((c.jp[lay] - 13) * 5 + (c.jt[lay] - 1)) * nsp + 1; };
  auto i1b = [&](int lay, int nsp) { return ((c.jp[lay] - 12) * 5 + (c.jt1[lay] - 1)) * nsp + 1; };
  // eta index/fraction for a binary species pair
  struct Eta { float speccomb, specparm, f; int j; };
  auto eta = [&](float cola, float ratio, float colb, float mult) {
    Eta e;
    e.speccomb = cola + ratio * colb;
    e.specparm = cola / e.speccomb;
    if (e.specparm >= oneminus) e.specparm = oneminus;
    float specmult = mult * e.specparm;
    e.j = 1 + (int)specmult;
    e.f = fmod1(specmult);
    return e;
  };
  // Lower-atmosphere major-species term of the binary bands (e.g. LW:5219-5349): one pressure level
  // (fa = fac00|fac01, fb = fac10|fac11), 3 stencil regimes in eta.
  auto major_lower = [&](const FArr &absa, const Eta &e, int ind, float fa, float fb, int ig) {
    if (e.specparm < 0.125f) {
      float p = e.f - 1;
      float p4 = pow4(p);
      float fk0 = p4, fk1 = 1 - p - 2.0f * p4, fk2 = p + p4;
      float f000 = fk0 * fa, f100 = fk1 * fa, f200 = fk2 * fa, f010 = fk0 * fb, f110 = fk1 * fb, f210 = fk2 * fb;
      return e.speccomb * (f000 * absa(ind, ig) + f100 * absa(ind + 1, ig) + f200 * absa(ind + 2, ig) + f010 * absa(ind + 9, ig) +
                           f110 * absa(ind + 10, ig) + f210 * absa(ind + 11, ig));
    } else if (e.specparm > 0.875f) {
      float p = -e.f;
      float p4 = pow4(p);
      float fk0 = p4, fk1 = 1 - p - 2.0f * p4, fk2 = p + p4;
      float f000 = fk0 * fa, f100 = fk1 * fa, f200 = fk2 * fa, f010 = fk0 * fb, f110 = fk1 * fb, f210 = fk2 * fb;
      return e.speccomb * (f200 * absa(ind - 1, ig) + f100 * absa(ind, ig) + f000 * absa(ind + 1, ig) + f210 * absa(ind + 8, ig) +
                           f110 * absa(ind + 9, ig) + f010 * absa(ind + 10, ig));
    } else {
      float f000 = (1.f - e.f) * fa, f010 = (1.f - e.f) * fb, f100 = e.f * fa, f110 = e.f * fb;
      return e.speccomb * (f000 * absa(ind, ig) + f100 * absa(ind + 1, ig) + f010 * absa(ind + 9, ig) + f110 * absa(ind + 10, ig));
    }
  };
  auto major_upper = [&](const FArr &absb, const Eta &e, const Eta &e1, int lay, int ind0, int ind1, int ig) {
    float f000 = (1.f - e.f) * c.fac00[lay], f010 = (1.f - e.f) * c.fac10[lay], f100 = e.f * c.fac00[lay], f110 = e.f * c.fac10[lay];
    float f001 = (1.f - e1.f) * c.fac01[lay], f011 = (1.f - e1.f) * c.fac11[lay], f101 = e1.f * c.fac01[lay], f111 = e1.f * c.fac11[lay];
    return e.speccomb * (f000 * absb(ind0, ig) + f100 * absb(ind0 + 1, ig) + f010 * absb(ind0 + 5, ig) + f110 * absb(ind0 + 6, ig)) +
           e1.speccomb * (f001 * absb(ind1, ig) + f101 * absb(ind1 + 1, ig) + f011 * absb(ind1 + 5, ig) + f111 * absb(ind1 + 6, ig));
  };
  auto frac_eta = [&](const FArr &fr, int ig, int jpl, float fpl) { return fr(ig, jpl) + fpl * (fr(ig, jpl + 1) - fr(ig, jpl)); };
  // empirical column rescaling of a minor gas, e.g. LW:5208-5216
  auto adjcol = [&](float col, int lay, int imol, float thresh, float base, float expo) {
    float chi = col / coldry[lay];
    float rat = 1.e20f * chi / chi_mls(imol, c.jp[lay] + 1);
    if (rat > thresh) {
      float adjfac = base + powf(rat - base, expo);
      return adjfac * chi_mls(imol, c.jp[lay] + 1) * coldry[lay] * 1.e-20f;
    }
    return col;
  };
  auto lowbin_idx = [&](int lay, int nsp, const Eta &e, const Eta &e1, int &ind0, int &ind1) {
    ind0 = ((c.jp[lay] - 1) * 5 + (c.jt[lay] - 1)) * nsp + e.j;
    ind1 = (c.jp[lay] * 5 + (c.jt1[lay] - 1)) * nsp + e1.j;
  };
  auto upbin_idx = [&](int lay, int nsp, const Eta &e, const Eta &e1, int &ind0, int &ind1) {
    ind0 = ((c.jp[lay] - 13) * 5 + (c.jt[lay] - 1)) * nsp + e.j;
    ind1 = ((c.jp[lay] - 12) * 5 + (c.jt1[lay] - 1)) * nsp + e1.j;
  };

  int ngs = 0;
  // ---- band 1 (LW:4961-5055)
  {
    const LwBand &B = T.lw[0];
    for (int lay = 1; lay <= nlayers; lay++) {
      bool low = lay <= laytrop;
      int ind0 = low ? i0a(lay, 1) : i0b(lay, 1), ind1 = low ? i1a(lay, 1) : i1b(lay, 1);
      float pp = pavel[lay];
      float corradj;
      if (low) { corradj = 1.f; if (pp < 250.f) corradj = 1.f - 0.15f * (250.f - pp) / 154.4f; }
      else corradj = 1.f - 0.15f * (pp / 95.6f);
      float scalen2 = c.colbrd[lay] * c.scaleminorn2[lay];
      for (int ig = 1; ig <= B.ng; ig++) {
        float taufor = fork(B, lay, ig);
        if (low) {
          float tauself = selfk(B, lay, ig);
          float taun2 = scalen2 * minor1(B.ka_mn2, lay, ig);
          o.taug[ngs + ig][lay] = corradj * (c.colh2o[lay] * k4(B.absa, lay, ind0, ind1, ig) + tauself + taufor + taun2);
          o.fracs[ngs + ig][lay] = B.fracrefa(ig, 1);
        } else {
          float taun2 = scalen2 * minor1(B.kb_mn2, lay, ig);
          o.taug[ngs + ig][lay] = corradj * (c.colh2o[lay] * k4(B.absb, lay, ind0, ind1, ig) + taufor + taun2);
          o.fracs[ngs + ig][lay] = B.fracrefb(ig, 1);
        }
      }
    }
    ngs += B.ng;
  }
  // ---- band 2 (LW:5057-5127)
  {
    const LwBand &B = T.lw[1];
    for (int lay = 1; lay <= nlayers; lay++) {
      bool low = lay <= laytrop;
      int ind0 = low ? i0a(lay, 1) : i0b(lay, 1), ind1 = low ? i1a(lay, 1) : i1b(lay, 1);
      float pp = pavel[lay];
      float corradj = 1.f - .05f * (pp - 100.f) / 900.f;
      for (int ig = 1; ig <= B.ng; ig++) {
        float taufor = fork(B, lay, ig);
        if (low) {
          float tauself = selfk(B, lay, ig);
          o.taug[ngs + ig][lay] = corradj * (c.colh2o[lay] * k4(B.absa, lay, ind0, ind1, ig) + tauself + taufor);
          o.fracs[ngs + ig][lay] = B.fracrefa(ig, 1);
        } else {
          o.taug[ngs + ig][lay] = c.colh2o[lay] * k4(B.absb, lay, ind0, ind1, ig) + taufor;
          o.fracs[ngs + ig][lay] = B.fracrefb(ig, 1);
        }
      }
    }
    ngs += B.ng;
  }
  // ---- band 3 (LW:5129-5442): h2o,co2 both; n2o minor
  {
    const LwBand &B = T.lw[2];
    float refrat_planck_a = chi_mls(1, 9) / chi_mls(2, 9), refrat_planck_b = chi_mls(1, 13) / chi_mls(2, 13);
    float refrat_m_a = chi_mls(1, 3) / chi_mls(2, 3), refrat_m_b = chi_mls(1, 13) / chi_mls(2, 13);
    for (int lay = 1; lay <= nlayers; lay++) {
      bool low = lay <= laytrop;
      float mult = low ? 8.f : 4.f;
      Eta e = eta(c.colh2o[lay], c.rat_h2oco2[lay], c.colco2[lay], mult);
      Eta e1 = eta(c.colh2o[lay], c.rat_h2oco2_1[lay], c.colco2[lay], mult);
      Eta em = eta(c.colh2o[lay], low ? refrat_m_a : refrat_m_b, c.colco2[lay], mult);
      float adjcoln2o = adjcol(c.coln2o[lay], lay, 4, 1.5f, 0.5f, 0.65f);
      Eta ep = eta(c.colh2o[lay], low ? refrat_planck_a : refrat_planck_b, c.colco2[lay], mult);
      int ind0, ind1;
      if (low) lowbin_idx(lay, 9, e, e1, ind0, ind1); else upbin_idx(lay, 5, e, e1, ind0, ind1);
      for (int ig = 1; ig <= B.ng; ig++) {
        float taufor = fork(B, lay, ig);
        if (low) {
          float tauself = selfk(B, lay, ig);
          float absn2o = minor2(B.ka_mn2o, lay, em.j, em.f, ig);
          float tau_major = major_lower(B.absa, e, ind0, c.fac00[lay], c.fac10[lay], ig);
          float tau_major1 = major_lower(B.absa, e1, ind1, c.fac01[lay], c.fac11[lay], ig);
          o.taug[ngs + ig][lay] = tau_major + tau_major1 + tauself + taufor + adjcoln2o * absn2o;
          o.fracs[ngs + ig][lay] = frac_eta(B.fracrefa, ig, ep.j, ep.f);
        } else {
          float absn2o = minor2(B.kb_mn2o, lay, em.j, em.f, ig);
          o.taug[ngs + ig][lay] = major_upper(B.absb, e, e1, lay, ind0, ind1, ig) + taufor + adjcoln2o * absn2o;
          o.fracs[ngs + ig][lay] = frac_eta(B.fracrefb, ig, ep.j, ep.f);
        }
      }
    }
    ngs += B.ng;
  }
  // ---- band 4 (LW:5444-5701): low h2o,co2; high o3,co2
  {
    const LwBand &B = T.lw[3];
    float refrat_planck_a = chi_mls(1, 11) / chi_mls(2, 11), refrat_planck_b = chi_mls(3, 13) / chi_mls(2, 13);
    for (int lay = 1; lay <= nlayers; lay++) {
      bool low = lay <= laytrop;
      int ind0, ind1;
      if (low) {
        Eta e = eta(c.colh2o[lay], c.rat_h2oco2[lay], c.colco2[lay], 8.f);
        Eta e1 = eta(c.colh2o[lay], c.rat_h2oco2_1[lay], c.colco2[lay], 8.f);
        Eta ep = eta(c.colh2o[lay], refrat_planck_a, c.colco2[lay], 8.f);
        lowbin_idx(lay, 9, e, e1, ind0, ind1);
        for (int ig = 1; ig <= B.ng; ig++) {
          float tauself = selfk(B, lay, ig), taufor = fork(B, lay, ig);
          float tau_major = major_lower(B.absa, e, ind0, c.fac00[lay], c.fac10[lay], ig);
          float tau_major1 = major_lower(B.absa, e1, ind1, c.fac01[lay], c.fac11[lay], ig);
          o.taug[ngs + ig][lay] = tau_major + tau_major1 + tauself + taufor;
          o.fracs[ngs + ig][lay] = frac_eta(B.fracrefa, ig, ep.j, ep.f);
        }
      } else {
        Eta e = eta(c.colo3[lay], c.rat_o3co2[lay], c.colco2[lay], 4.f);
        Eta e1 = eta(c.colo3[lay], c.rat_o3co2_1[lay], c.colco2[lay], 4.f);
        Eta ep = eta(c.colo3[lay], refrat_planck_b, c.colco2[lay], 4.f);
        upbin_idx(lay, 5, e, e1, ind0, ind1);
        for (int ig = 1; ig <= B.ng; ig++) {
          o.taug[ngs + ig][lay] = major_upper(B.absb, e, e1, lay, ind0, ind1, ig);
          o.fracs[ngs + ig][lay] = frac_eta(B.fracrefb, ig, ep.j, ep.f);
        }
        o.taug[ngs + 8][lay] = o.taug[ngs + 8][lay] * 0.92f;
        o.taug[ngs + 9][lay] = o.taug[ngs + 9][lay] * 0.88f;
        o.taug[ngs + 10][lay] = o.taug[ngs + 10][lay] * 1.07f;
        o.taug[ngs + 11][lay] = o.taug[ngs + 11][lay] * 1.1f;
        o.taug[ngs + 12][lay] = o.taug[ngs + 12][lay] * 0.99f;
        o.taug[ngs + 13][lay] = o.taug[ngs + 13][lay] * 0.88f;
        o.taug[ngs + 14][lay] = o.taug[ngs + 14][lay] * 0.943f;
      }
    }
    ngs += B.ng;
  }
  // ---- band 5 (LW:5703-5976): low h2o,co2 (+o3 minor); high o3,co2; ccl4
  {
    const LwBand &B = T.lw[4];
    float refrat_planck_a = chi_mls(1, 5) / chi_mls(2, 5), refrat_planck_b = chi_mls(3, 43) / chi_mls(2, 43);
    float refrat_m_a = chi_mls(1, 7) / chi_mls(2, 7);
    for (int lay = 1; lay <= nlayers; lay++) {
      bool low = lay <= laytrop;
      int ind0, ind1;
      if (low) {
        Eta e = eta(c.colh2o[lay], c.rat_h2oco2[lay], c.colco2[lay], 8.f);
        Eta e1 = eta(c.colh2o[lay], c.rat_h2oco2_1[lay], c.colco2[lay], 8.f);
        Eta em = eta(c.colh2o[lay], refrat_m_a, c.colco2[lay], 8.f);
        Eta ep = eta(c.colh2o[lay], refrat_planck_a, c.colco2[lay], 8.f);
        lowbin_idx(lay, 9, e, e1, ind0, ind1);
        for (int ig = 1; ig <= B.ng; ig++) {
          float tauself = selfk(B, lay, ig), taufor = fork(B, lay, ig);
          float abso3 = minor2(B.ka_mo3, lay, em.j, em.f, ig);
          float tau_major = major_lower(B.absa, e, ind0, c.fac00[lay], c.fac10[lay], ig);
          float tau_major1 = major_lower(B.absa, e1, ind1, c.fac01[lay], c.fac11[lay], ig);
          o.taug[ngs + ig][lay] = tau_major + tau_major1 + tauself + taufor + abso3 * c.colo3[lay] + wx[1][lay] * B.ccl4(ig);
          o.fracs[ngs + ig][lay] = frac_eta(B.fracrefa, ig, ep.j, ep.f);
        }
      } else {
        Eta e = eta(c.colo3[lay], c.rat_o3co2[lay], c.colco2[lay], 4.f);
        Eta e1 = eta(c.colo3[lay], c.rat_o3co2_1[lay], c.colco2[lay], 4.f);
        Eta ep = eta(c.colo3[lay], refrat_planck_b, c.colco2[lay], 4.f);
        upbin_idx(lay, 5, e, e1, ind0, ind1);
        for (int ig = 1; ig <= B.ng; ig++) {
          o.taug[ngs + ig][lay] = major_upper(B.absb, e, e1, lay, ind0, ind1, ig) + wx[1][lay] * B.ccl4(ig);
          o.fracs[ngs + ig][lay] = frac_eta(B.fracrefb, ig, ep.j, ep.f);
        }
      }
    }
    ngs += B.ng;
  }
  // ---- band 6 (LW:5978-6062): low h2o (+co2 minor); cfc11, cfc12
  {
    const LwBand &B = T.lw[5];
    for (int lay = 1; lay <= nlayers; lay++) {
      if (lay <= laytrop) {
        float adjcolco2 = adjcol(c.colco2[lay], lay, 2, 3.0f, 2.0f, 0.77f);
        int ind0 = i0a(lay, 1), ind1 = i1a(lay, 1);
        for (int ig = 1; ig <= B.ng; ig++) {
          float tauself = selfk(B, lay, ig), taufor = fork(B, lay, ig);
          float absco2 = minor1(B.ka_mco2, lay, ig);
          o.taug[ngs + ig][lay] = c.colh2o[lay] * k4(B.absa, lay, ind0, ind1, ig) + tauself + taufor + adjcolco2 * absco2 +
                                  wx[2][lay] * B.cfc11adj(ig) + wx[3][lay] * B.cfc12(ig);
          o.fracs[ngs + ig][lay] = B.fracrefa(ig, 1);
        }
      } else {
        for (int ig = 1; ig <= B.ng; ig++) {
          o.taug[ngs + ig][lay] = 0.0f + wx[2][lay] * B.cfc11adj(ig) + wx[3][lay] * B.cfc12(ig);
          o.fracs[ngs + ig][lay] = B.fracrefa(ig, 1);
        }
      }
    }
    ngs += B.ng;
  }
  // ---- band 7 (LW:6064-6336): low h2o,o3; high o3; co2 minor
  {
    const LwBand &B = T.lw[6];
    float refrat_planck_a = chi_mls(1, 3) / chi_mls(3, 3), refrat_m_a = chi_mls(1, 3) / chi_mls(3, 3);
    for (int lay = 1; lay <= nlayers; lay++) {
      if (lay <= laytrop) {
        Eta e = eta(c.colh2o[lay], c.rat_h2oo3[lay], c.colo3[lay], 8.f);
        Eta e1 = eta(c.colh2o[lay], c.rat_h2oo3_1[lay], c.colo3[lay], 8.f);
        Eta em = eta(c.colh2o[lay], refrat_m_a, c.colo3[lay], 8.f);
        float adjcolco2 = adjcol(c.colco2[lay], lay, 2, 3.0f, 3.0f, 0.79f);
        Eta ep = eta(c.colh2o[lay], refrat_planck_a, c.colo3[lay], 8.f);
        int ind0, ind1; lowbin_idx(lay, 9, e, e1, ind0, ind1);
        for (int ig = 1; ig <= B.ng; ig++) {
          float tauself = selfk(B, lay, ig), taufor = fork(B, lay, ig);
          float absco2 = minor2(B.ka_mco2, lay, em.j, em.f, ig);
          float tau_major = major_lower(B.absa, e, ind0, c.fac00[lay], c.fac10[lay], ig);
          float tau_major1 = major_lower(B.absa, e1, ind1, c.fac01[lay], c.fac11[lay], ig);
          o.taug[ngs + ig][lay] = tau_major + tau_major1 + tauself + taufor + adjcolco2 * absco2;
          o.fracs[ngs + ig][lay] = frac_eta(B.fracrefa, ig, ep.j, ep.f);
        }
      } else {
        float adjcolco2 = adjcol(c.colco2[lay], lay, 2, 3.0f, 2.0f, 0.79f);
        int ind0 = i0b(lay, 1), ind1 = i1b(lay, 1);
        for (int ig = 1; ig <= B.ng; ig++) {
          float absco2 = minor1(B.kb_mco2, lay, ig);
          o.taug[ngs + ig][lay] = c.colo3[lay] * k4(B.absb, lay, ind0, ind1, ig) + adjcolco2 * absco2;
          o.fracs[ngs + ig][lay] = B.fracrefb(ig, 1);
        }
        o.taug[ngs + 6][lay] = o.taug[ngs + 6][lay] * 0.92f;
        o.taug[ngs + 7][lay] = o.taug[ngs + 7][lay] * 0.88f;
        o.taug[ngs + 8][lay] = o.taug[ngs + 8][lay] * 1.07f;
        o.taug[ngs + 9][lay] = o.taug[ngs + 9][lay] * 1.1f;
        o.taug[ngs + 10][lay] = o.taug[ngs + 10][lay] * 0.99f;
        o.taug[ngs + 11][lay] = o.taug[ngs + 11][lay] * 0.855f;
      }
    }
    ngs += B.ng;
  }
  // ---- band 8 (LW:6338-6459): low h2o; high o3; minors co2, o3, n2o; cfc12, cfc22
  {
    const LwBand &B = T.lw[7];
    for (int lay = 1; lay <= nlayers; lay++) {
      float adjcolco2 = adjcol(c.colco2[lay], lay, 2, 3.0f, 2.0f, 0.65f);
      if (lay <= laytrop) {
        int ind0 = i0a(lay, 1), ind1 = i1a(lay, 1);
        for (int ig = 1; ig <= B.ng; ig++) {
          float tauself = selfk(B, lay, ig), taufor = fork(B, lay, ig);
          float absco2 = minor1(B.ka_mco2, lay, ig), abso3 = minor1(B.ka_mo3, lay, ig), absn2o = minor1(B.ka_mn2o, lay, ig);
          o.taug[ngs + ig][lay] = c.colh2o[lay] * k4(B.absa, lay, ind0, ind1, ig) + tauself + taufor + adjcolco2 * absco2 +
                                  c.colo3[lay] * abso3 + c.coln2o[lay] * absn2o + wx[3][lay] * B.cfc12(ig) + wx[4][lay] * B.cfc22adj(ig);
          o.fracs[ngs + ig][lay] = B.fracrefa(ig, 1);
        }
      } else {
        int ind0 = i0b(lay, 1), ind1 = i1b(lay, 1);
        for (int ig = 1; ig <= B.ng; ig++) {
          float absco2 = minor1(B.kb_mco2, lay, ig), absn2o = minor1(B.kb_mn2o, lay, ig);
          o.taug[ngs + ig][lay] = c.colo3[lay] * k4(B.absb, lay, ind0, ind1, ig) + adjcolco2 * absco2 + c.coln2o[lay] * absn2o +
                                  wx[3][lay] * B.cfc12(ig) + wx[4][lay] * B.cfc22adj(ig);
          o.fracs[ngs + ig][lay] = B.fracrefb(ig, 1);
        }
      }
    }
    ngs += B.ng;
  }
  // ---- band 9 (LW:6461-6722): low h2o,ch4; high ch4; n2o minor
  {
    const LwBand &B = T.lw[8];
    float refrat_planck_a = chi_mls(1, 9) / chi_mls(6, 9), refrat_m_a = chi_mls(1, 3) / chi_mls(6, 3);
    for (int lay = 1; lay <= nlayers; lay++) {
      float adjcoln2o = adjcol(c.coln2o[lay], lay, 4, 1.5f, 0.5f, 0.65f);
      if (lay <= laytrop) {
        Eta e = eta(c.colh2o[lay], c.rat_h2och4[lay], c.colch4[lay], 8.f);
        Eta e1 = eta(c.colh2o[lay], c.rat_h2och4_1[lay], c.colch4[lay], 8.f);
        Eta em = eta(c.colh2o[lay], refrat_m_a, c.colch4[lay], 8.f);
        Eta ep = eta(c.colh2o[lay], refrat_planck_a, c.colch4[lay], 8.f);
        int ind0, ind1; lowbin_idx(lay, 9, e, e1, ind0, ind1);
        for (int ig = 1; ig <= B.ng; ig++) {
          float tauself = selfk(B, lay, ig), taufor = fork(B, lay, ig);
          float absn2o = minor2(B.ka_mn2o, lay, em.j, em.f, ig);
          float tau_major = major_lower(B.absa, e, ind0, c.fac00[lay], c.fac10[lay], ig);
          float tau_major1 = major_lower(B.absa, e1, ind1, c.fac01[lay], c.fac11[lay], ig);
          o.taug[ngs + ig][lay] = tau_major + tau_major1 + tauself + taufor + adjcoln2o * absn2o;
          o.fracs[ngs + ig][lay] = frac_eta(B.fracrefa, ig, ep.j, ep.f);
        }
      } else {
        int ind0 = i0b(lay, 1), ind1 = i1b(lay, 1);
        for (int ig = 1; ig <= B.ng; ig++) {
          float absn2o = minor1(B.kb_mn2o, lay, ig);
          o.taug[ngs + ig][lay] = c.colch4[lay] * k4(B.absb, lay, ind0, ind1, ig) + adjcoln2o * absn2o;
          o.fracs[ngs + ig][lay] = B.fracrefb(ig, 1);
        }
      }
    }
    ngs += B.ng;
  }
  // ---- bands 10 (LW:6724-6789) and 11 (LW:6791-6869, + o2 minor): h2o both
  for (int q = 0; q < 2; q++) {
    const LwBand &B = T.lw[9 + q];
    for (int lay = 1; lay <= nlayers; lay++) {
      bool low = lay <= laytrop;
      int ind0 = low ? i0a(lay, 1) : i0b(lay, 1), ind1 = low ? i1a(lay, 1) : i1b(lay, 1);
      float scaleo2 = c.colo2[lay] * c.scaleminor[lay];
      for (int ig = 1; ig <= B.ng; ig++) {
        float taufor = fork(B, lay, ig);
        if (low) {
          float tauself = selfk(B, lay, ig);
          float t = c.colh2o[lay] * k4(B.absa, lay, ind0, ind1, ig) + tauself + taufor;
          if (q == 1) t = t + scaleo2 * minor1(B.ka_mo2, lay, ig);
          o.taug[ngs + ig][lay] = t;
          o.fracs[ngs + ig][lay] = B.fracrefa(ig, 1);
        } else {
          float t = c.colh2o[lay] * k4(B.absb, lay, ind0, ind1, ig) + taufor;
          if (q == 1) t = t + scaleo2 * minor1(B.kb_mo2, lay, ig);
          o.taug[ngs + ig][lay] = t;
          o.fracs[ngs + ig][lay] = B.fracrefb(ig, 1);
        }
      }
    }
    ngs += B.ng;
  }
  // ---- band 12 (LW:6871-7073): low h2o,co2; high nothing
  {
    const LwBand &B = T.lw[11];
    float refrat_planck_a = chi_mls(1, 10) / chi_mls(2, 10);
    for (int lay = 1; lay <= nlayers; lay++) {
      if (lay <= laytrop) {
        Eta e = eta(c.colh2o[lay], c.rat_h2oco2[lay], c.colco2[lay], 8.f);
        Eta e1 = eta(c.colh2o[lay], c.rat_h2oco2_1[lay], c.colco2[lay], 8.f);
        Eta ep = eta(c.colh2o[lay], refrat_planck_a, c.colco2[lay], 8.f);
        int ind0, ind1; lowbin_idx(lay, 9, e, e1, ind0, ind1);
        for (int ig = 1; ig <= B.ng; ig++) {
          float tauself = selfk(B, lay, ig), taufor = fork(B, lay, ig);
          float tau_major = major_lower(B.absa, e, ind0, c.fac00[lay], c.fac10[lay], ig);
          float tau_major1 = major_lower(B.absa, e1, ind1, c.fac01[lay], c.fac11[lay], ig);
          o.taug[ngs + ig][lay] = tau_major + tau_major1 + tauself + taufor;
          o.fracs[ngs + ig][lay] = frac_eta(B.fracrefa, ig, ep.j, ep.f);
        }
      } else {
        for (int ig = 1; ig <= B.ng; ig++) { o.taug[ngs + ig][lay] = 0.f; o.fracs[ngs + ig][lay] = 0.f; }
      }
    }
    ngs += B.ng;
  }
  // ---- band 13 (LW:7075-7332): low h2o,n2o (+co2, co minors); high o3 minor
  {
    const LwBand &B = T.lw[12];
    float refrat_planck_a = chi_mls(1, 5) / chi_mls(4, 5), refrat_m_a = chi_mls(1, 1) / chi_mls(4, 1),
          refrat_m_a3 = chi_mls(1, 3) / chi_mls(4, 3);
    for (int lay = 1; lay <= nlayers; lay++) {
      if (lay <= laytrop) {
        Eta e = eta(c.colh2o[lay], c.rat_h2on2o[lay], c.coln2o[lay], 8.f);
        Eta e1 = eta(c.colh2o[lay], c.rat_h2on2o_1[lay], c.coln2o[lay], 8.f);
        Eta em = eta(c.colh2o[lay], refrat_m_a, c.coln2o[lay], 8.f);
        float chi_co2 = c.colco2[lay] / (coldry[lay]);
        float ratco2 = 1.e20f * chi_co2 / 3.55e-4f;
        float adjcolco2;
        if (ratco2 > 3.0f) {
          float adjfac = 2.0f + powf(ratco2 - 2.0f, 0.68f);
          adjcolco2 = adjfac * 3.55e-4f * coldry[lay] * 1.e-20f;
        } else adjcolco2 = c.colco2[lay];
        Eta eco = eta(c.colh2o[lay], refrat_m_a3, c.coln2o[lay], 8.f);
        Eta ep = eta(c.colh2o[lay], refrat_planck_a, c.coln2o[lay], 8.f);
        int ind0, ind1; lowbin_idx(lay, 9, e, e1, ind0, ind1);
        for (int ig = 1; ig <= B.ng; ig++) {
          float tauself = selfk(B, lay, ig), taufor = fork(B, lay, ig);
          float absco2 = minor2(B.ka_mco2, lay, em.j, em.f, ig);
          float absco = minor2(B.ka_mco, lay, eco.j, eco.f, ig);
          float tau_major = major_lower(B.absa, e, ind0, c.fac00[lay], c.fac10[lay], ig);
          float tau_major1 = major_lower(B.absa, e1, ind1, c.fac01[lay], c.fac11[lay], ig);
          o.taug[ngs + ig][lay] = tau_major + tau_major1 + tauself + taufor + adjcolco2 * absco2 + c.colco[lay] * absco;
          o.fracs[ngs + ig][lay] = frac_eta(B.fracrefa, ig, ep.j, ep.f);
        }
      } else {
        for (int ig = 1; ig <= B.ng; ig++) {
          float abso3 = minor1(B.kb_mo3, lay, ig);
          o.taug[ngs + ig][lay] = c.colo3[lay] * abso3;
          o.fracs[ngs + ig][lay] = B.fracrefb(ig, 1);
        }
      }
    }
    ngs += B.ng;
  }
  // ---- band 14 (LW:7334-7393): co2 both
  {
    const LwBand &B = T.lw[13];
    for (int lay = 1; lay <= nlayers; lay++) {
      if (lay <= laytrop) {
        int ind0 = i0a(lay, 1), ind1 = i1a(lay, 1);
        for (int ig = 1; ig <= B.ng; ig++) {
          float tauself = selfk(B, lay, ig), taufor = fork(B, lay, ig);
          o.taug[ngs + ig][lay] = c.colco2[lay] * k4(B.absa, lay, ind0, ind1, ig) + tauself + taufor;
          o.fracs[ngs + ig][lay] = B.fracrefa(ig, 1);
        }
      } else {
        int ind0 = i0b(lay, 1), ind1 = i1b(lay, 1);
        for (int ig = 1; ig <= B.ng; ig++) {
          o.taug[ngs + ig][lay] = c.colco2[lay] * k4(B.absb, lay, ind0, ind1, ig);
          o.fracs[ngs + ig][lay] = B.fracrefb(ig, 1);
        }
      }
    }
    ngs += B.ng;
  }
  // ---- band 15 (LW:7395-7618): low n2o,co2 (+n2 minor); high nothing
  {
    const LwBand &B = T.lw[14];
    float refrat_planck_a = chi_mls(4, 1) / chi_mls(2, 1), refrat_m_a = chi_mls(4, 1) / chi_mls(2, 1);
    for (int lay = 1; lay <= nlayers; lay++) {
      if (lay <= laytrop) {
        Eta e = eta(c.coln2o[lay], c.rat_n2oco2[lay], c.colco2[lay], 8.f);
        Eta e1 = eta(c.coln2o[lay], c.rat_n2oco2_1[lay], c.colco2[lay], 8.f);
        Eta em = eta(c.coln2o[lay], refrat_m_a, c.colco2[lay], 8.f);
        Eta ep = eta(c.coln2o[lay], refrat_planck_a, c.colco2[lay], 8.f);
        int ind0, ind1; lowbin_idx(lay, 9, e, e1, ind0, ind1);
        float scalen2 = c.colbrd[lay] * c.scaleminor[lay];
        for (int ig = 1; ig <= B.ng; ig++) {
          float tauself = selfk(B, lay, ig), taufor = fork(B, lay, ig);
          float taun2 = scalen2 * minor2(B.ka_mn2, lay, em.j, em.f, ig);
          float tau_major = major_lower(B.absa, e, ind0, c.fac00[lay], c.fac10[lay], ig);
          float tau_major1 = major_lower(B.absa, e1, ind1, c.fac01[lay], c.fac11[lay], ig);
          o.taug[ngs + ig][lay] = tau_major + tau_major1 + tauself + taufor + taun2;
          o.fracs[ngs + ig][lay] = frac_eta(B.fracrefa, ig, ep.j, ep.f);
        }
      } else {
        for (int ig = 1; ig <= B.ng; ig++) { o.taug[ngs + ig][lay] = 0.f; o.fracs[ngs + ig][lay] = 0.f; }
      }
    }
    ngs += B.ng;
  }
  // ---- band 16 (LW:7620-7826): low h2o,ch4; high ch4
  {
    const LwBand &B = T.lw[15];
    float refrat_planck_a = chi_mls(1, 6) / chi_mls(6, 6);
    for (int lay = 1; lay <= nlayers; lay++) {
      if (lay <= laytrop) {
        Eta e = eta(c.colh2o[lay], c.rat_h2och4[lay], c.colch4[lay], 8.f);
        Eta e1 = eta(c.colh2o[lay], c.rat_h2och4_1[lay], c.colch4[lay], 8.f);
        Eta ep = eta(c.colh2o[lay], refrat_planck_a, c.colch4[lay], 8.f);
        int ind0, ind1; lowbin_idx(lay, 9, e, e1, ind0, ind1);
        for (int ig = 1; ig <= B.ng; ig++) {
          float tauself = selfk(B, lay, ig), taufor = fork(B, lay, ig);
          float tau_major = major_lower(B.absa, e, ind0, c.fac00[lay], c.fac10[lay], ig);
          float tau_major1 = major_lower(B.absa, e1, ind1, c.fac01[lay], c.fac11[lay], ig);
          o.taug[ngs + ig][lay] = tau_major + tau_major1 + tauself + taufor;
          o.fracs[ngs + ig][lay] = frac_eta(B.fracrefa, ig, ep.j, ep.f);
        }
      } else {
        int ind0 = i0b(lay, 1), ind1 = i1b(lay, 1);
        for (int ig = 1; ig <= B.ng; ig++) {
          o.taug[ngs + ig][lay] = c.colch4[lay] * k4(B.absb, lay, ind0, ind1, ig);
          o.fracs[ngs + ig][lay] = B.fracrefb(ig, 1);
        }
      }
    }
    ngs += B.ng;
  }
}

struct LwCld {
  float cldfmc[NGLW + 1][MXLAY], ciwpmc[NGLW + 1][MXLAY], clwpmc[NGLW + 1][MXLAY], cswpmc[NGLW + 1][MXLAY], taucmc[NGLW + 1][MXLAY];
  float reicmc[MXLAY], relqmc[MXLAY], resnmc[MXLAY];
};

// ---- cldprmc LW:2653-2914 (inflag >= 2, iceflag >= 3, liqflag 1 as set by RRTMG_LWRAD) -----------------
int cldprmc_lw(const Tables &T, int nlayers, int inflag, int iceflag, int liqflag, LwCld &s, std::string &err) {
  const float cldmin = 1.e-20f;
  const FArr &absliq1 = T.in.get("lw_absliq1"), &absice3 = T.in.get("lw_absice3");
  static thread_local float abscoice[NGLW + 1], abscoliq[NGLW + 1], abscosno[NGLW + 1];
  for (int lay = 1; lay <= nlayers; lay++) {
    for (int ig = 1; ig <= NGLW; ig++) {
      float cwp = s.ciwpmc[ig][lay] + s.clwpmc[ig][lay] + s.cswpmc[ig][lay];
      if (s.cldfmc[ig][lay] >= cldmin && (cwp >= cldmin || s.taucmc[ig][lay] >= cldmin)) {
        if (inflag == 0) return 0;
        if (inflag == 1) { err = "INFLAG = 1 OPTION NOT AVAILABLE WITH MCICA"; return ARC_ERR_UNSUPPORTED; }
        float radice = s.reicmc[lay];
        int ib = T.lw_ngb[ig - 1];
        if ((s.ciwpmc[ig][lay] + s.cswpmc[ig][lay]) == 0.0f) { abscoice[ig] = 0.f; abscosno[ig] = 0.f; }
        else if (iceflag >= 3) {
          if (radice < 5.0f || radice > 140.0f) { err = "ERROR: ICE GENERALIZED EFFECTIVE SIZE OUT OF BOUNDS"; return ARC_ERR_RADIUS; }
          float factor = (radice - 2.f) / 3.f;
          int index = (int)factor;
          if (index == 46) index = 45;
          float fint = factor - (float)index;
          abscoice[ig] = absice3(index, ib) + fint * (absice3(index + 1, ib) - (absice3(index, ib)));
          abscosno[ig] = 0.f;
        } else { err = "oracle: iceflag < 3 not restated (WRF always uses 3,4,5; LW:12047-12120)"; return ARC_ERR_UNSUPPORTED; }
        if (s.cswpmc[ig][lay] > 0.0f && iceflag == 5) {
          float radsno = s.resnmc[lay];
          if (radsno < 5.0f || radsno > 140.0f) { err = "ERROR: SNOW GENERALIZED EFFECTIVE SIZE OUT OF BOUNDS"; return ARC_ERR_RADIUS; }
          float factor = (radsno - 2.f) / 3.f;
          int index = (int)factor;
          if (index == 46) index = 45;
          float fint = factor - (float)index;
          abscosno[ig] = absice3(index, ib) + fint * (absice3(index + 1, ib) - (absice3(index, ib)));
        }
        if (s.clwpmc[ig][lay] == 0.0f) abscoliq[ig] = 0.f;
        else if (liqflag == 1) {
          float radliq = s.relqmc[lay];
          if (radliq < 2.5f || radliq > 60.f) { err = "LIQUID EFFECTIVE RADIUS OUT OF BOUNDS"; return ARC_ERR_RADIUS; }
          int index = (int)(radliq - 1.5f);
          if (index == 0) index = 1;
          if (index == 58) index = 57;
          float fint = radliq - 1.5f - (float)index;
          abscoliq[ig] = absliq1(index, ib) + fint * (absliq1(index + 1, ib) - (absliq1(index, ib)));
        }
        s.taucmc[ig][lay] = s.ciwpmc[ig][lay] * abscoice[ig] + s.clwpmc[ig][lay] * abscoliq[ig] + s.cswpmc[ig][lay] * abscosno[ig];
      }
    }
  }
  return 0;
}

struct LwFlux { float totuflux[MXLAY + 1], totdflux[MXLAY + 1], fnet[MXLAY + 1], htr[MXLAY + 1], totuclfl[MXLAY + 1], totdclfl[MXLAY + 1], fnetc[MXLAY + 1], htrc[MXLAY + 1]; };

// ---- rtrnmc LW:2974-3410 ----------------------------------------------------------------------------------
void rtrnmc(const Tables &T, int nlayers, const float *pz, const float *semiss, const LwCld &cl, const LwCoef &c, float pwvcm,
            const float (*fracs)[MXLAY], const float (*taut)[MXLAY], LwFlux &F) {
  const float wtdiff = 0.5f, rec_6 = 0.166667f, tblint = 10000.0f;
  const float bpade = T.lw_bpade;
  const float *tau_tbl = T.lw_tau_tbl.data(), *exp_tbl = T.lw_exp_tbl.data(), *tfn_tbl = T.lw_tfn_tbl.data();
  const FArr &a0 = T.in.get("lw_a0"), &a1 = T.in.get("lw_a1"), &a2 = T.in.get("lw_a2");
  float secdiff[NBLW + 1];
  static thread_local float odcld[NGLW + 1][MXLAY], abscld[NGLW + 1][MXLAY], efclfrac[NGLW + 1][MXLAY];
  float urad[MXLAY + 1], drad[MXLAY + 1], clrurad[MXLAY + 1], clrdrad[MXLAY + 1], atrans[MXLAY], atot[MXLAY], bbugas[MXLAY], bbutot[MXLAY];
  int icldlyr[MXLAY];
  for (int ibnd = 1; ibnd <= NBLW; ibnd++) {
    if (ibnd == 1 || ibnd == 4 || ibnd >= 10) secdiff[ibnd] = 1.66f;
    else {
      secdiff[ibnd] = a0(ibnd) + a1(ibnd) * expf(a2(ibnd) * pwvcm);
      if (secdiff[ibnd] > 1.80f) secdiff[ibnd] = 1.80f;
      if (secdiff[ibnd] < 1.50f) secdiff[ibnd] = 1.50f;
    }
  }
  urad[0] = 0.f; drad[0] = 0.f; F.totuflux[0] = 0.f; F.totdflux[0] = 0.f; clrurad[0] = 0.f; clrdrad[0] = 0.f; F.totuclfl[0] = 0.f; F.totdclfl[0] = 0.f;
  for (int lay = 1; lay <= nlayers; lay++) {
    urad[lay] = 0.f; drad[lay] = 0.f; F.totuflux[lay] = 0.f; F.totdflux[lay] = 0.f;
    clrurad[lay] = 0.f; clrdrad[lay] = 0.f; F.totuclfl[lay] = 0.f; F.totdclfl[lay] = 0.f;
    icldlyr[lay] = 0;
    for (int ig = 1; ig <= NGLW; ig++) {
      if (cl.cldfmc[ig][lay] == 1.f) {
        int ib = T.lw_ngb[ig - 1];
        odcld[ig][lay] = secdiff[ib] * cl.taucmc[ig][lay];
        float transcld = expf(-odcld[ig][lay]);
        abscld[ig][lay] = 1.f - transcld;
        efclfrac[ig][lay] = abscld[ig][lay] * cl.cldfmc[ig][lay];
        icldlyr[lay] = 1;
      } else { odcld[ig][lay] = 0.f; abscld[ig][lay] = 0.f; efclfrac[ig][lay] = 0.f; }
    }
  }
  int igc = 1;
  for (int iband = 1; iband <= 16; iband++) {
    do {
      float radld = 0.f, radclrd = 0.f;
      int iclddn = 0;
      for (int lev = nlayers; lev >= 1; lev--) {
        float plfrac = fracs[igc][lev];
        float blay = c.planklay[iband][lev];
        float dplankup = c.planklev[iband][lev] - blay;
        float dplankdn = c.planklev[iband][lev - 1] - blay;
        float odepth = secdiff[iband] * taut[igc][lev];
        if (odepth < 0.0f) odepth = 0.0f;
        float bbd;
        if (icldlyr[lev] == 1) {
          iclddn = 1;
          float odtot = odepth + odcld[igc][lev];
          float gassrc, bbdtot;
          if (odtot < 0.06f) {
            atrans[lev] = odepth - 0.5f * odepth * odepth;
            float odepth_rec = rec_6 * odepth;
            gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans[lev];
            atot[lev] = odtot - 0.5f * odtot * odtot;
            float odtot_rec = rec_6 * odtot;
            bbdtot = plfrac * (blay + dplankdn * odtot_rec);
            bbd = plfrac * (blay + dplankdn * odepth_rec);
            radld = radld - radld * (atrans[lev] + efclfrac[igc][lev] * (1.f - atrans[lev])) + gassrc +
                    cl.cldfmc[igc][lev] * (bbdtot * atot[lev] - gassrc);
            drad[lev - 1] = drad[lev - 1] + radld;
            bbugas[lev] = plfrac * (blay + dplankup * odepth_rec);
            bbutot[lev] = plfrac * (blay + dplankup * odtot_rec);
          } else if (odepth <= 0.06f) {
            atrans[lev] = odepth - 0.5f * odepth * odepth;
            float odepth_rec = rec_6 * odepth;
            gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans[lev];
            odtot = odepth + odcld[igc][lev];
            float tblind = odtot / (bpade + odtot);
            int ittot = (int)(tblint * tblind + 0.5f);
            float tfactot = tfn_tbl[ittot];
            bbdtot = plfrac * (blay + tfactot * dplankdn);
            bbd = plfrac * (blay + dplankdn * odepth_rec);
            atot[lev] = 1.f - exp_tbl[ittot];
            radld = radld - radld * (atrans[lev] + efclfrac[igc][lev] * (1.f - atrans[lev])) + gassrc +
                    cl.cldfmc[igc][lev] * (bbdtot * atot[lev] - gassrc);
            drad[lev - 1] = drad[lev - 1] + radld;
            bbugas[lev] = plfrac * (blay + dplankup * odepth_rec);
            bbutot[lev] = plfrac * (blay + tfactot * dplankup);
          } else {
            float tblind = odepth / (bpade + odepth);
            int itgas = (int)(tblint * tblind + 0.5f);
            odepth = tau_tbl[itgas];
            atrans[lev] = 1.f - exp_tbl[itgas];
            float tfacgas = tfn_tbl[itgas];
            gassrc = atrans[lev] * plfrac * (blay + tfacgas * dplankdn);
            odtot = odepth + odcld[igc][lev];
            tblind = odtot / (bpade + odtot);
            int ittot = (int)(tblint * tblind + 0.5f);
            float tfactot = tfn_tbl[ittot];
            bbdtot = plfrac * (blay + tfactot * dplankdn);
            bbd = plfrac * (blay + tfacgas * dplankdn);
            atot[lev] = 1.f - exp_tbl[ittot];
            radld = radld - radld * (atrans[lev] + efclfrac[igc][lev] * (1.f - atrans[lev])) + gassrc +
                    cl.cldfmc[igc][lev] * (bbdtot * atot[lev] - gassrc);
            drad[lev - 1] = drad[lev - 1] + radld;
            bbugas[lev] = plfrac * (blay + tfacgas * dplankup);
            bbutot[lev] = plfrac * (blay + tfactot * dplankup);
          }
        } else {
          if (odepth <= 0.06f) {
            atrans[lev] = odepth - 0.5f * odepth * odepth;
            odepth = rec_6 * odepth;
            bbd = plfrac * (blay + dplankdn * odepth);
            bbugas[lev] = plfrac * (blay + dplankup * odepth);
          } else {
            float tblind = odepth / (bpade + odepth);
            int itr = (int)(tblint * tblind + 0.5f);
            float transc = exp_tbl[itr];
            atrans[lev] = 1.f - transc;
            float tausfac = tfn_tbl[itr];
            bbd = plfrac * (blay + tausfac * dplankdn);
            bbugas[lev] = plfrac * (blay + tausfac * dplankup);
          }
          radld = radld + (bbd - radld) * atrans[lev];
          drad[lev - 1] = drad[lev - 1] + radld;
        }
        if (iclddn == 1) {
          radclrd = radclrd + (bbd - radclrd) * atrans[lev];
          clrdrad[lev - 1] = clrdrad[lev - 1] + radclrd;
        } else {
          radclrd = radld;
          clrdrad[lev - 1] = drad[lev - 1];
        }
      }
      float rad0 = fracs[igc][1] * c.plankbnd[iband];
      float reflect = 1.f - semiss[iband];
      float radlu = rad0 + reflect * radld;
      float radclru = rad0 + reflect * radclrd;
      urad[0] = urad[0] + radlu;
      clrurad[0] = clrurad[0] + radclru;
      for (int lev = 1; lev <= nlayers; lev++) {
        if (icldlyr[lev] == 1) {
          float gassrc = bbugas[lev] * atrans[lev];
          radlu = radlu - radlu * (atrans[lev] + efclfrac[igc][lev] * (1.f - atrans[lev])) + gassrc +
                  cl.cldfmc[igc][lev] * (bbutot[lev] * atot[lev] - gassrc);
          urad[lev] = urad[lev] + radlu;
        } else {
          radlu = radlu + (bbugas[lev] - radlu) * atrans[lev];
          urad[lev] = urad[lev] + radlu;
        }
        if (iclddn == 1) {
          radclru = radclru + (bbugas[lev] - radclru) * atrans[lev];
          clrurad[lev] = clrurad[lev] + radclru;
        } else {
          radclru = radlu;
          clrurad[lev] = urad[lev];
        }
      }
      igc = igc + 1;
    } while (igc <= T.lw_ngs[iband - 1]);
    for (int lev = nlayers; lev >= 0; lev--) {
      float uflux = urad[lev] * wtdiff, dflux = drad[lev] * wtdiff;
      urad[lev] = 0.f; drad[lev] = 0.f;
      F.totuflux[lev] = F.totuflux[lev] + uflux * T.lw_delwave[iband - 1];
      F.totdflux[lev] = F.totdflux[lev] + dflux * T.lw_delwave[iband - 1];
      float uclfl = clrurad[lev] * wtdiff, dclfl = clrdrad[lev] * wtdiff;
      clrurad[lev] = 0.f; clrdrad[lev] = 0.f;
      F.totuclfl[lev] = F.totuclfl[lev] + uclfl * T.lw_delwave[iband - 1];
      F.totdclfl[lev] = F.totdclfl[lev] + dclfl * T.lw_delwave[iband - 1];
    }
  }
  F.totuflux[0] = F.totuflux[0] * T.fluxfac; F.totdflux[0] = F.totdflux[0] * T.fluxfac;
  F.fnet[0] = F.totuflux[0] - F.totdflux[0];
  F.totuclfl[0] = F.totuclfl[0] * T.fluxfac; F.totdclfl[0] = F.totdclfl[0] * T.fluxfac;
  F.fnetc[0] = F.totuclfl[0] - F.totdclfl[0];
  for (int lev = 1; lev <= nlayers; lev++) {
    F.totuflux[lev] = F.totuflux[lev] * T.fluxfac; F.totdflux[lev] = F.totdflux[lev] * T.fluxfac;
    F.fnet[lev] = F.totuflux[lev] - F.totdflux[lev];
    F.totuclfl[lev] = F.totuclfl[lev] * T.fluxfac; F.totdclfl[lev] = F.totdclfl[lev] * T.fluxfac;
    F.fnetc[lev] = F.totuclfl[lev] - F.totdclfl[lev];
    int l = lev - 1;
    F.htr[l] = T.heatfac * (F.fnet[l] - F.fnet[lev]) / (pz[l] - pz[lev]);
    F.htrc[l] = T.heatfac * (F.fnetc[l] - F.fnetc[lev]) / (pz[l] - pz[lev]);
  }
  F.htr[nlayers] = 0.f; F.htrc[nlayers] = 0.f;
}

struct LwWork {
  LwCld cld; LwTau tau; LwCoef coef; LwFlux F, Fcln;
  float taut[NGLW + 1][MXLAY], taua[NBLW + 1][MXLAY], wkl[8][MXLAY], wx[5][MXLAY];
};

}  // namespace

// ---- RRTMG_LWRAD LW:11451-12700 with rrtmg_lw 10578-11064 and inatm 11067-11403 inlined ------------------
int oracle_lwrad(const ArcDims &d, const ArcLwIn &in, ArcLwOut &out, ArcDebug *dbg, std::string &err) {
  const Tables &T = tables();
  if (!T.ready) { err = "oracle not initialised"; return ARC_ERR_NOT_INIT; }
  if (in.aer_ra_feedback == 1) {
    for (int b = 0; b < 16; b++)
      if (!in.tauaerlw[b]) { err = "Warning: missing fields required for aerosol radiation"; return ARC_ERR_MISSING_FIELD; }
  }
  const int clean = in.clean_atm_diag;
  const Idx ix(d);
  const int kts = d.kts, kte = d.kte, nz = kte - kts + 1;
  const int nlayers = T.lw_nlayers - kts + 1;  // arrays are dimensioned kts:nlayers; kts = 1 in WRF
  const int nlay = nlayers;
  if (nlay + 2 >= MXLAY || nlay < nz + 1) { err = "bad LW layer count"; return ARC_ERR_BAD_ARG; }
  const FArr &retab = T.in.get("lw_retab"), &pprof = T.in.get("lw_pprof"), &tprof = T.in.get("lw_tprof");
  const float co2 = 379.e-6f, ch4 = 1774.e-9f, n2o = 319.e-9f, cfc11 = 0.251e-9f, cfc12 = 0.538e-9f, cfc22 = 0.169e-9f,
              ccl4 = 0.093e-9f, o2 = 0.209488f;
  const float amdw = 1.607793f, amdo = 0.603461f, deltap = 4.f, thresh = 1.e-9f;
  const float amd = 28.9660f, amw = 18.0160f;
  const int nproflevs = 60;
  const int nci = d.ite - d.its + 1;
  static thread_local LwWork *Wp = nullptr;
  if (!Wp) Wp = new LwWork;
  LwWork &W = *Wp;
  CloudIn ci{in.icloud, in.warm_rain, in.is_cammgmp_used, in.has_reqc, in.has_reqi, in.has_reqs, in.progn,
             in.f_qv, in.f_qc, in.f_qr, in.f_qi, in.f_qs, in.f_qg, in.f_qndrop, in.g,
             in.t3d, in.cldfra3d, in.lradius, in.iradius, in.qv3d, in.qc3d, in.qr3d, in.qi3d, in.qs3d, in.qg3d, in.qndrop3d,
             in.re_cloud, in.re_ice, in.re_snow, in.f_ice_phy, in.xland, in.xice, in.snow};
  for (int j = d.jts; j <= d.jte; j++) {
    for (int i = d.its; i <= d.ite; i++) {
      const size_t ij = ix.at2(i, j);
      const size_t cidx = (size_t)(j - d.jts) * nci + (i - d.its);
      float pw1d[MXLAY], tw1d[MXLAY], t1d[MXLAY], p1d[MXLAY], o31d[MXLAY];
      for (int k = 1; k <= nz + 1; k++) { pw1d[k] = in.p8w[ix.at3(i, kts + k - 1, j)] / 100.f; tw1d[k] = in.t8w[ix.at3(i, kts + k - 1, j)]; }
      for (int k = 1; k <= nz; k++) {
        size_t q = ix.at3(i, kts + k - 1, j);
        t1d[k] = in.t3d[q]; p1d[k] = in.p3d[q] / 100.f;
        o31d[k] = in.o33d ? in.o33d[q] : 0.f;
      }
      Col1D col;
      gather_hydrometeors(ci, ix, i, j, kts, kte, t1d, col);
      effective_radius_inputs(ci, ix, i, j, kts, kte, retab, col);
      const int inflglw = col.inflg, iceflglw = col.iceflg, liqflglw = col.liqflg;
      float plev[MXLAY + 2], tlev[MXLAY + 2], play[MXLAY], tlay[MXLAY], pdel[MXLAY], h2ovmr[MXLAY], o3vmr[MXLAY], o3mmr[MXLAY], varint[MXLAY + 2];
      plev[1] = pw1d[1]; tlev[1] = tw1d[1];
      float tsfc = in.tsk[ij];
      for (int k = 1; k <= nz; k++) {
        play[k] = p1d[k]; plev[k + 1] = pw1d[k + 1]; pdel[k] = plev[k] - plev[k + 1];
        tlay[k] = t1d[k]; tlev[k + 1] = tw1d[k + 1];
        h2ovmr[k] = col.qv[k] * amdw;
      }
      // buffer layers above the model top, LW:12207-12252
      for (int L = nz + 1; L <= nlayers; L++) { plev[L + 1] = plev[L] - deltap; play[L] = 0.5f * (plev[L] + plev[L + 1]); }
      plev[nlayers + 1] = 0.00f;
      play[nlayers] = 0.5f * (plev[nlayers] + plev[nlayers + 1]);
      for (int L = 1; L <= nlayers + 1; L++) {
        int klev = nproflevs;
        if (pprof(nproflevs) < plev[L]) {
          for (int LL = 2; LL <= nproflevs; LL++) { if (pprof(LL) < plev[L]) { klev = LL - 1; break; } }
        } else klev = nproflevs;
        float vark, vark1, wght;
        if (klev != nproflevs) { vark = tprof(klev); vark1 = tprof(klev + 1); wght = (plev[L] - pprof(klev)) / (pprof(klev + 1) - pprof(klev)); }
        else { vark = tprof(klev); vark1 = tprof(klev); wght = 0.0f; }
        varint[L] = wght * (vark1 - vark) + vark;
      }
      for (int L = nz + 1; L <= nlayers + 1; L++) {
        tlev[L] = varint[L] + (tlev[nz] - varint[nz]);
        tlay[L - 1] = 0.5f * (tlev[L] + tlev[L - 1]);
      }
      for (int L = nz + 1; L <= nlayers; L++) h2ovmr[L] = h2ovmr[nz];
      o3data(T.in, plev, nlayers, o3mmr);
      for (int k = 1; k <= nlayers; k++) {
        o3vmr[k] = o3mmr[k] * amdo;
        if (in.o33d && in.o3input == 2) {
          if (k <= nz) o3vmr[k] = o31d[k];
          else {
            o3vmr[k] = o31d[nz] - o3mmr[nz] * amdo + o3mmr[k] * amdo;
            if (o3vmr[k] <= 0.f) o3vmr[k] = o3mmr[k] * amdo;
          }
        }
      }
      float semiss[NBLW + 1];
      for (int nb = 1; nb <= NBLW; nb++) semiss[nb] = in.emiss[ij];
      CloudPaths cp;
      cloud_paths(ci, ix, i, j, kts, kte, col, pdel, tlay, retab, cp);
      for (int k = nz + 1; k <= nlayers; k++) {
        cp.clwp[k] = 0.f; cp.ciwp[k] = 0.f; cp.cswp[k] = 0.f; cp.rel[k] = 10.f; cp.rei[k] = 10.f; cp.res[k] = 10.f; cp.cldfrac[k] = 0.f;
      }
      // mcica_subcol_lw (permuteseed = 150, irng = 0), LW:2089-2204
      float pmid[MXLAY];
      for (int l = 1; l <= nlay; l++) pmid[l] = play[l] * 1.e2f;
      std::vector<float> cdf; std::vector<unsigned char> cloudy;
      mcica_mask(nlay, NGLW, pmid, cp.cldfrac, 150, cdf, cloudy);
      for (int l = 1; l <= nlay; l++) {
        for (int ig = 1; ig <= NGLW; ig++) {
          bool clf = cloudy[(size_t)(ig - 1) * (nlay + 1) + l] != 0;
          W.cld.cldfmc[ig][l] = clf ? 1.f : 0.f;
          W.cld.clwpmc[ig][l] = clf ? cp.clwp[l] : 0.f;
          W.cld.ciwpmc[ig][l] = clf ? cp.ciwp[l] : 0.f;
          W.cld.cswpmc[ig][l] = clf ? cp.cswp[l] : 0.f;
          W.cld.taucmc[ig][l] = 0.f;
        }
        W.cld.reicmc[l] = cp.rei[l]; W.cld.relqmc[l] = cp.rel[l]; W.cld.resnmc[l] = cp.res[l];
      }
      // aerosol, LW:12576-12629
      for (int nb = 1; nb <= NBLW; nb++) for (int k = 1; k <= nlayers; k++) W.taua[nb][k] = 0.f;
      if (in.aer_ra_feedback == 1) {
        for (int k = 1; k <= nz; k++) {
          size_t q = ix.at3(i, kts + k - 1, j);
          if (in.tauaerlw[0][q] > thresh && in.tauaerlw[15][q] > thresh)
            for (int nb = 1; nb <= NBLW; nb++) W.taua[nb][k] = in.tauaerlw[nb - 1][q];
        }
        for (int nb = 1; nb <= NBLW; nb++) {
          float slope = 0.f;
          for (int k = 1; k <= nz; k++) slope = slope + W.taua[nb][k];
          if (slope < 0.f) { err = "ERROR: Negative total lw optical depth"; return ARC_ERR_NEG_AOD; }
        }
      }
      // ---- inatm LW:11067-11403
      float pavel[MXLAY], tavel[MXLAY], pz[MXLAY + 1], tz[MXLAY + 1], coldry[MXLAY], wbrodl[MXLAY];
      for (int m = 0; m < 8; m++) for (int l = 0; l < MXLAY; l++) W.wkl[m][l] = 0.f;
      for (int m = 0; m < 5; m++) for (int l = 0; l < MXLAY; l++) W.wx[m][l] = 0.f;
      float amttl = 0.f, wvttl = 0.f;
      float tbound = tsfc;
      pz[0] = plev[1]; tz[0] = tlev[1];
      for (int l = 1; l <= nlayers; l++) {
        pavel[l] = play[l]; tavel[l] = tlay[l]; pz[l] = plev[l + 1]; tz[l] = tlev[l + 1];
        W.wkl[1][l] = h2ovmr[l]; W.wkl[2][l] = co2; W.wkl[3][l] = o3vmr[l]; W.wkl[4][l] = n2o; W.wkl[6][l] = ch4; W.wkl[7][l] = o2;
        float amm = (1.f - W.wkl[1][l]) * amd + W.wkl[1][l] * amw;
        coldry[l] = (pz[l - 1] - pz[l]) * 1.e3f * T.avogad / (1.e2f * T.grav * amm * (1.f + W.wkl[1][l]));
      }
      for (int l = 1; l <= nlayers; l++) { W.wx[1][l] = ccl4; W.wx[2][l] = cfc11; W.wx[3][l] = cfc12; W.wx[4][l] = cfc22; }
      for (int l = 1; l <= nlayers; l++) {
        float summol = 0.f;
        for (int imol = 2; imol <= 7; imol++) summol = summol + W.wkl[imol][l];
        wbrodl[l] = coldry[l] * (1.f - summol);
        for (int imol = 1; imol <= 7; imol++) W.wkl[imol][l] = coldry[l] * W.wkl[imol][l];
        amttl = amttl + coldry[l] + W.wkl[1][l];
        wvttl = wvttl + W.wkl[1][l];
        for (int ixs = 1; ixs <= 4; ixs++) W.wx[ixs][l] = coldry[l] * W.wx[ixs][l] * 1.e-20f;
      }
      float wvsh = (amw * wvttl) / (amd * amttl);
      float pwvcm = wvsh * (1.e3f * pz[0]) / (1.e2f * T.grav);
      int rc = cldprmc_lw(T, nlayers, inflglw, iceflglw, liqflglw, W.cld, err);
      if (rc) return rc;
      setcoef_lw(T, nlayers, pavel, tavel, tz, tbound, semiss, coldry, W.wkl, wbrodl, W.coef);
      taumol_lw(T, nlayers, pavel, W.wx, coldry, W.coef, W.tau);
      for (int k = 1; k <= nlayers; k++)
        for (int ig = 1; ig <= NGLW; ig++) W.taut[ig][k] = W.tau.taug[ig][k] + W.taua[T.lw_ngb[ig - 1]][k];
      if (clean > 0) rtrnmc(T, nlayers, pz, semiss, W.cld, W.coef, pwvcm, W.tau.fracs, W.tau.taug, W.Fcln);
      rtrnmc(T, nlayers, pz, semiss, W.cld, W.coef, pwvcm, W.tau.fracs, W.taut, W.F);
      // uflx(k+1) = totuflux(k)
      auto UF = [&](int k) { return W.F.totuflux[k - 1]; };
      auto DF = [&](int k) { return W.F.totdflux[k - 1]; };
      auto UFC = [&](int k) { return W.F.totuclfl[k - 1]; };
      auto DFC = [&](int k) { return W.F.totdclfl[k - 1]; };
      auto UFN = [&](int k) { return clean > 0 ? W.Fcln.totuflux[k - 1] : 0.f; };
      auto DFN = [&](int k) { return clean > 0 ? W.Fcln.totdflux[k - 1] : 0.f; };
      out.glw[ij] = DF(1);
      out.olr[ij] = UF(nlayers + 1);
      out.lwcf[ij] = UFC(nlayers + 1) - UF(nlayers + 1);
      if (out.lwupt) {
        out.lwupt[ij] = UF(nlayers + 1); out.lwuptc[ij] = UFC(nlayers + 1); out.lwdnt[ij] = DF(nlayers + 1); out.lwdntc[ij] = DFC(nlayers + 1);
        out.lwupb[ij] = UF(1); out.lwupbc[ij] = UFC(1); out.lwdnb[ij] = DF(1); out.lwdnbc[ij] = DFC(1);
        out.lwuptcln[ij] = UFN(nlayers + 1); out.lwdntcln[ij] = DFN(nlayers + 1); out.lwupbcln[ij] = UFN(1); out.lwdnbcln[ij] = DFN(1);
      }
      if (out.lwuptclnc) {
        out.lwuptclnc[ij] = clean > 0 ? W.Fcln.totuclfl[nlayers] : 0.f; out.lwdntclnc[ij] = clean > 0 ? W.Fcln.totdclfl[nlayers] : 0.f;
        out.lwupbclnc[ij] = clean > 0 ? W.Fcln.totuclfl[0] : 0.f; out.lwdnbclnc[ij] = clean > 0 ? W.Fcln.totdclfl[0] : 0.f;
      }
      if (out.lwupflx) {
        for (int k = 1; k <= nz + 2; k++) {
          size_t q = ix.atp(i, kts + k - 1, j);
          out.lwupflx[q] = UF(k); out.lwupflxc[q] = UFC(k); out.lwdnflx[q] = DF(k); out.lwdnflxc[q] = DFC(k);
          out.lwupflxcln[q] = UFN(k); out.lwdnflxcln[q] = DFN(k);
        }
      }
      for (int k = 1; k <= nz; k++) {
        float tten = W.F.htr[k - 1] / 86400.f;
        out.rthratenlw[ix.at3(i, kts + k - 1, j)] = tten / in.pi3d[ix.at3(i, kts + k - 1, j)];
      }
      if (dbg) {
        if (dbg->laytrop) dbg->laytrop[cidx] = W.coef.laytrop;
        for (int l = 1; l <= nlay; l++) {
          size_t q = cidx * nlay + (l - 1);
          if (dbg->jp) dbg->jp[q] = W.coef.jp[l];
          if (dbg->jt) dbg->jt[q] = W.coef.jt[l];
          if (dbg->jt1) dbg->jt1[q] = W.coef.jt1[l];
          if (dbg->indfor) dbg->indfor[q] = W.coef.indfor[l];
          if (dbg->indself) dbg->indself[q] = W.coef.indself[l];
          if (dbg->indminor) dbg->indminor[q] = W.coef.indminor[l];
          if (dbg->fac00) dbg->fac00[q] = W.coef.fac00[l];
          if (dbg->fac01) dbg->fac01[q] = W.coef.fac01[l];
          if (dbg->fac10) dbg->fac10[q] = W.coef.fac10[l];
          if (dbg->fac11) dbg->fac11[q] = W.coef.fac11[l];
          if (dbg->hr) dbg->hr[q] = W.F.htr[l - 1];
          for (int ig = 1; ig <= NGLW; ig++) {
            size_t qq = q * NGLW + (ig - 1);
            if (dbg->cldmask) dbg->cldmask[qq] = W.cld.cldfmc[ig][l] != 0.f;
            if (dbg->taug) dbg->taug[qq] = W.tau.taug[ig][l];
            if (dbg->taur) dbg->taur[qq] = W.tau.fracs[ig][l];
            if (dbg->taucmc) dbg->taucmc[qq] = W.cld.taucmc[ig][l];
          }
        }
      }
    }
  }
  return 0;
}

}  // namespace orc
