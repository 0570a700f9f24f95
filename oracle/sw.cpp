// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// Shortwave: restates, in the reference's evaluation order and one column at a time,
//   cldprmc_sw   module_ra_rrtmg_sw.F:1969-2390      reftra_sw  2422-2701
//   setcoef_sw   2734-2990                            taumol_sw  3081-4540 (taumol16..29)
//   vrtqdr_sw    7922-8046                            spcvmc_sw  8083-8658
//   rrtmg_sw     8740-9490   inatm_sw 9520-9873       RRTMG_SWRAD 9901-11207
#include <cmath>
#include <cstdio>
#include <cstring>

#include "adapter_common.hpp"
#include "oracle.hpp"

namespace orc {

namespace {

struct SwCoef {
  int laytrop, jp[MXLAY], jt[MXLAY], jt1[MXLAY], indself[MXLAY], indfor[MXLAY];
  float colh2o[MXLAY], colco2[MXLAY], colo3[MXLAY], coln2o[MXLAY], colch4[MXLAY], colo2[MXLAY], colmol[MXLAY], co2mult[MXLAY];
  float fac00[MXLAY], fac01[MXLAY], fac10[MXLAY], fac11[MXLAY];
  float selffac[MXLAY], selffrac[MXLAY], forfac[MXLAY], forfrac[MXLAY];
};

// ---- setcoef_sw SW:2734-2990 -------------------------------------------------------------------
void setcoef_sw(const Tables &T, int nlayers, const float *pavel, const float *tavel, const float *coldry,
                const float (*wkl)[MXLAY], SwCoef &c) {
  const FArr &preflog = T.in.get("sw_preflog"), &tref = T.in.get("sw_tref");
  const float stpfac = 296.f / 1013.f;
  c.laytrop = 0;
  for (int lay = 1; lay <= nlayers; lay++) {
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
    } else {
      c.forfac[lay] = scalefac / (1.f + water);
      factor = (tavel[lay] - 188.0f) / 36.0f;
      c.indfor[lay] = 3;
      c.forfrac[lay] = factor - 1.0f;
      c.selffac[lay] = 0.f; c.selffrac[lay] = 0.f; c.indself[lay] = 0;
    }
    c.colh2o[lay] = 1.e-20f * wkl[1][lay];
    c.colco2[lay] = 1.e-20f * wkl[2][lay];
    c.colo3[lay] = 1.e-20f * wkl[3][lay];
    c.coln2o[lay] = 1.e-20f * wkl[4][lay];
    c.colch4[lay] = 1.e-20f * wkl[6][lay];
    c.colo2[lay] = 1.e-20f * wkl[7][lay];
    c.colmol[lay] = 1.e-20f * coldry[lay] + c.colh2o[lay];
    if (c.colco2[lay] == 0.f) c.colco2[lay] = 1.e-32f * coldry[lay];
    if (c.coln2o[lay] == 0.f) c.coln2o[lay] = 1.e-32f * coldry[lay];
    if (c.colch4[lay] == 0.f) c.colch4[lay] = 1.e-32f * coldry[lay];
    if (c.colo2[lay] == 0.f) c.colo2[lay] = 1.e-32f * coldry[lay];
    float co2reg = 3.55e-24f * coldry[lay];
    c.co2mult[lay] = (c.colco2[lay] - co2reg) * 272.63f * expf(-1919.4f / tavel[lay]) / (8.7604e-4f * tavel[lay]);
    float compfp = 1.f - fp;
    c.fac10[lay] = compfp * ft;
    c.fac00[lay] = compfp * (1.f - ft);
    c.fac11[lay] = fp * ft1;
    c.fac01[lay] = fp * (1.f - ft1);
  }
}

// ---- taumol_sw SW:3081-4540 ---------------------------------------------------------------------
// taug/taur (lay, ig) as [ig][lay], 1-based.
struct SwTau { float taug[NGSW + 1][MXLAY], taur[NGSW + 1][MXLAY], sfluxzen[NGSW + 1]; };

inline float fmod1(float x) { return x - (float)(int)x; }  // Fortran MOD(x,1.) for x >= 0

void taumol_sw(const Tables &T, int nlayers, const SwCoef &c, SwTau &o) {
  const int laytrop = c.laytrop;
  const float oneminus = T.oneminus;
  auto selfk = [&](const SwBand &B, int lay, int ig) {
    int inds = c.indself[lay];
    return c.selffac[lay] * (B.selfref(inds, ig) + c.selffrac[lay] * (B.selfref(inds + 1, ig) - B.selfref(inds, ig)));
  };
  auto fork = [&](const SwBand &B, int lay, int ig) {
    int indf = c.indfor[lay];
    return c.forfac[lay] * (B.forref(indf, ig) + c.forfrac[lay] * (B.forref(indf + 1, ig) - B.forref(indf, ig)));
  };
  // 8-point (eta, T, p) interpolation; stride = 9 (lower) or 5 (upper)
  struct Bin { float speccomb, fs; int js, ind0, ind1; float f000, f010, f100, f110, f001, f011, f101, f111; };
  auto binary = [&](int lay, float cola, float colb_scaled, float mult, int nsp, bool lower) {
    Bin b;
    b.speccomb = cola + colb_scaled;
    float specparm = cola / b.speccomb;
    if (specparm >= oneminus) specparm = oneminus;
    float specmult = mult * specparm;
    b.js = 1 + (int)specmult;
    b.fs = fmod1(specmult);
    b.f000 = (1.f - b.fs) * c.fac00[lay]; b.f010 = (1.f - b.fs) * c.fac10[lay];
    b.f100 = b.fs * c.fac00[lay];         b.f110 = b.fs * c.fac10[lay];
    b.f001 = (1.f - b.fs) * c.fac01[lay]; b.f011 = (1.f - b.fs) * c.fac11[lay];
    b.f101 = b.fs * c.fac01[lay];         b.f111 = b.fs * c.fac11[lay];
    if (lower) {
      b.ind0 = ((c.jp[lay] - 1) * 5 + (c.jt[lay] - 1)) * nsp + b.js;
      b.ind1 = (c.jp[lay] * 5 + (c.jt1[lay] - 1)) * nsp + b.js;
    } else {
      b.ind0 = ((c.jp[lay] - 13) * 5 + (c.jt[lay] - 1)) * nsp + b.js;
      b.ind1 = ((c.jp[lay] - 12) * 5 + (c.jt1[lay] - 1)) * nsp + b.js;
    }
    return b;
  };
  auto k8 = [&](const FArr &ab, const Bin &b, int ig, int st) {
    return b.f000 * ab(b.ind0, ig) + b.f100 * ab(b.ind0 + 1, ig) + b.f010 * ab(b.ind0 + st, ig) + b.f110 * ab(b.ind0 + st + 1, ig) +
           b.f001 * ab(b.ind1, ig) + b.f101 * ab(b.ind1 + 1, ig) + b.f011 * ab(b.ind1 + st, ig) + b.f111 * ab(b.ind1 + st + 1, ig);
  };
  auto k4 = [&](const FArr &ab, int lay, int ind0, int ind1, int ig) {
    return c.fac00[lay] * ab(ind0, ig) + c.fac10[lay] * ab(ind0 + 1, ig) + c.fac01[lay] * ab(ind1, ig) + c.fac11[lay] * ab(ind1 + 1, ig);
  };
  auto i0a = [&](int lay, int nsp) { return ((c.jp[lay] - 1) * 5 + (c.jt[lay] - 1)) * nsp + 1; };
  auto i1a = [&](int lay, int nsp) { return (c.jp[lay] * 5 + (c.jt1[lay] - 1)) * nsp + 1; };
  auto i0b = [&](int lay, int nsp) { return ((c.jp[lay] - 13) * 5 + (c.jt[lay] - 1)) * nsp + 1; };
  auto i1b = [&](int lay, int nsp) { return ((c.jp[lay] - 12) * 5 + (c.jt1[lay] - 1)) * nsp + 1; };
  auto sflx_eta = [&](const SwBand &B, int ig, int js, float fs) {
    return B.sfluxref(ig, js) + fs * (B.sfluxref(ig, js + 1) - B.sfluxref(ig, js));
  };

  int ngs = 0, laysolfr;
  // ---- band 16 (SW:3293-3385): low h2o,ch4; high ch4
  {
    const SwBand &B = T.sw[0];
    for (int lay = 1; lay <= laytrop; lay++) {
      Bin b = binary(lay, c.colh2o[lay], B.strrat * c.colch4[lay], 8.f, 9, true);
      float tauray = c.colmol[lay] * B.rayl;
      for (int ig = 1; ig <= B.ng; ig++) {
        o.taug[ngs + ig][lay] = b.speccomb * k8(B.absa, b, ig, 9) + c.colh2o[lay] * (selfk(B, lay, ig) + fork(B, lay, ig));
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    laysolfr = nlayers;
    for (int lay = laytrop + 1; lay <= nlayers; lay++) {
      if (c.jp[lay - 1] < B.layreffr && c.jp[lay] >= B.layreffr) laysolfr = lay;
      int ind0 = i0b(lay, 1), ind1 = i1b(lay, 1);
      float tauray = c.colmol[lay] * B.rayl;
      for (int ig = 1; ig <= B.ng; ig++) {
        o.taug[ngs + ig][lay] = c.colch4[lay] * k4(B.absb, lay, ind0, ind1, ig);
        if (lay == laysolfr) o.sfluxzen[ngs + ig] = B.sfluxref(ig, 1);
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    ngs += B.ng;
  }
  // ---- band 17 (SW:3388-3504): low h2o,co2; high h2o,co2
  {
    const SwBand &B = T.sw[1];
    for (int lay = 1; lay <= laytrop; lay++) {
      Bin b = binary(lay, c.colh2o[lay], B.strrat * c.colco2[lay], 8.f, 9, true);
      float tauray = c.colmol[lay] * B.rayl;
      for (int ig = 1; ig <= B.ng; ig++) {
        o.taug[ngs + ig][lay] = b.speccomb * k8(B.absa, b, ig, 9) + c.colh2o[lay] * (selfk(B, lay, ig) + fork(B, lay, ig));
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    laysolfr = nlayers;
    for (int lay = laytrop + 1; lay <= nlayers; lay++) {
      if (c.jp[lay - 1] < B.layreffr && c.jp[lay] >= B.layreffr) laysolfr = lay;
      Bin b = binary(lay, c.colh2o[lay], B.strrat * c.colco2[lay], 4.f, 5, false);
      float tauray = c.colmol[lay] * B.rayl;
      for (int ig = 1; ig <= B.ng; ig++) {
        o.taug[ngs + ig][lay] = b.speccomb * k8(B.absb, b, ig, 5) + c.colh2o[lay] * fork(B, lay, ig);
        if (lay == laysolfr) o.sfluxzen[ngs + ig] = sflx_eta(B, ig, b.js, b.fs);
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    ngs += B.ng;
  }
  // ---- bands 18, 19 (SW:3507-3696): low h2o,(ch4|co2); high (ch4|co2)
  for (int q = 0; q < 2; q++) {
    const SwBand &B = T.sw[2 + q];
    const float *colb = q == 0 ? c.colch4 : c.colco2;
    laysolfr = laytrop;
    for (int lay = 1; lay <= laytrop; lay++) {
      if (c.jp[lay] < B.layreffr && c.jp[lay + 1] >= B.layreffr) laysolfr = std::min(lay + 1, laytrop);
      Bin b = binary(lay, c.colh2o[lay], B.strrat * colb[lay], 8.f, 9, true);
      float tauray = c.colmol[lay] * B.rayl;
      for (int ig = 1; ig <= B.ng; ig++) {
        o.taug[ngs + ig][lay] = b.speccomb * k8(B.absa, b, ig, 9) + c.colh2o[lay] * (selfk(B, lay, ig) + fork(B, lay, ig));
        if (lay == laysolfr) o.sfluxzen[ngs + ig] = sflx_eta(B, ig, b.js, b.fs);
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; lay++) {
      int ind0 = i0b(lay, 1), ind1 = i1b(lay, 1);
      float tauray = c.colmol[lay] * B.rayl;
      for (int ig = 1; ig <= B.ng; ig++) {
        o.taug[ngs + ig][lay] = colb[lay] * k4(B.absb, lay, ind0, ind1, ig);
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    ngs += B.ng;
  }
  // ---- band 20 (SW:3699-3780): h2o both, + ch4 continuum
  {
    const SwBand &B = T.sw[4];
    laysolfr = laytrop;
    for (int lay = 1; lay <= laytrop; lay++) {
      if (c.jp[lay] < B.layreffr && c.jp[lay + 1] >= B.layreffr) laysolfr = std::min(lay + 1, laytrop);
      int ind0 = i0a(lay, 1), ind1 = i1a(lay, 1);
      float tauray = c.colmol[lay] * B.rayl;
      for (int ig = 1; ig <= B.ng; ig++) {
        o.taug[ngs + ig][lay] = c.colh2o[lay] * ((k4(B.absa, lay, ind0, ind1, ig)) + selfk(B, lay, ig) + fork(B, lay, ig)) +
                                c.colch4[lay] * B.absch4(ig);
        o.taur[ngs + ig][lay] = tauray;
        if (lay == laysolfr) o.sfluxzen[ngs + ig] = B.sfluxref(ig, 1);
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; lay++) {
      int ind0 = i0b(lay, 1), ind1 = i1b(lay, 1);
      float tauray = c.colmol[lay] * B.rayl;
      for (int ig = 1; ig <= B.ng; ig++) {
        o.taug[ngs + ig][lay] = c.colh2o[lay] * (k4(B.absb, lay, ind0, ind1, ig) + fork(B, lay, ig)) + c.colch4[lay] * B.absch4(ig);
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    ngs += B.ng;
  }
  // ---- band 21 (SW:3783-3899): h2o,co2 both
  {
    const SwBand &B = T.sw[5];
    laysolfr = laytrop;
    for (int lay = 1; lay <= laytrop; lay++) {
      if (c.jp[lay] < B.layreffr && c.jp[lay + 1] >= B.layreffr) laysolfr = std::min(lay + 1, laytrop);
      Bin b = binary(lay, c.colh2o[lay], B.strrat * c.colco2[lay], 8.f, 9, true);
      float tauray = c.colmol[lay] * B.rayl;
      for (int ig = 1; ig <= B.ng; ig++) {
        o.taug[ngs + ig][lay] = b.speccomb * k8(B.absa, b, ig, 9) + c.colh2o[lay] * (selfk(B, lay, ig) + fork(B, lay, ig));
        if (lay == laysolfr) o.sfluxzen[ngs + ig] = sflx_eta(B, ig, b.js, b.fs);
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; lay++) {
      Bin b = binary(lay, c.colh2o[lay], B.strrat * c.colco2[lay], 4.f, 5, false);
      float tauray = c.colmol[lay] * B.rayl;
      for (int ig = 1; ig <= B.ng; ig++) {
        o.taug[ngs + ig][lay] = b.speccomb * k8(B.absb, b, ig, 5) + c.colh2o[lay] * fork(B, lay, ig);
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    ngs += B.ng;
  }
  // ---- band 22 (SW:3902-4005): low h2o,o2; high o2
  {
    const SwBand &B = T.sw[6];
    const float o2adj = 1.6f;
    laysolfr = laytrop;
    for (int lay = 1; lay <= laytrop; lay++) {
      if (c.jp[lay] < B.layreffr && c.jp[lay + 1] >= B.layreffr) laysolfr = std::min(lay + 1, laytrop);
      float o2cont = 4.35e-4f * c.colo2[lay] / (350.0f * 2.0f);
      Bin b = binary(lay, c.colh2o[lay], o2adj * B.strrat * c.colo2[lay], 8.f, 9, true);
      float tauray = c.colmol[lay] * B.rayl;
      for (int ig = 1; ig <= B.ng; ig++) {
        o.taug[ngs + ig][lay] = b.speccomb * k8(B.absa, b, ig, 9) + c.colh2o[lay] * (selfk(B, lay, ig) + fork(B, lay, ig)) + o2cont;
        if (lay == laysolfr) o.sfluxzen[ngs + ig] = sflx_eta(B, ig, b.js, b.fs);
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; lay++) {
      float o2cont = 4.35e-4f * c.colo2[lay] / (350.0f * 2.0f);
      int ind0 = i0b(lay, 1), ind1 = i1b(lay, 1);
      float tauray = c.colmol[lay] * B.rayl;
      for (int ig = 1; ig <= B.ng; ig++) {
        o.taug[ngs + ig][lay] = c.colo2[lay] * o2adj * k4(B.absb, lay, ind0, ind1, ig) + o2cont;
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    ngs += B.ng;
  }
  // ---- band 23 (SW:4008-4074): low h2o; high nothing
  {
    const SwBand &B = T.sw[7];
    laysolfr = laytrop;
    for (int lay = 1; lay <= laytrop; lay++) {
      if (c.jp[lay] < B.layreffr && c.jp[lay + 1] >= B.layreffr) laysolfr = std::min(lay + 1, laytrop);
      int ind0 = i0a(lay, 1), ind1 = i1a(lay, 1);
      for (int ig = 1; ig <= B.ng; ig++) {
        float tauray = c.colmol[lay] * B.raylg(ig);
        o.taug[ngs + ig][lay] = c.colh2o[lay] * (B.givfac * k4(B.absa, lay, ind0, ind1, ig) + selfk(B, lay, ig) + fork(B, lay, ig));
        if (lay == laysolfr) o.sfluxzen[ngs + ig] = B.sfluxref(ig, 1);
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; lay++)
      for (int ig = 1; ig <= B.ng; ig++) {
        o.taug[ngs + ig][lay] = 0.f;
        o.taur[ngs + ig][lay] = c.colmol[lay] * B.raylg(ig);
      }
    ngs += B.ng;
  }
  // ---- band 24 (SW:4077-4174): low h2o,o2; high o2; o3 continuum; rayleigh(eta)
  {
    const SwBand &B = T.sw[8];
    laysolfr = laytrop;
    for (int lay = 1; lay <= laytrop; lay++) {
      if (c.jp[lay] < B.layreffr && c.jp[lay + 1] >= B.layreffr) laysolfr = std::min(lay + 1, laytrop);
      Bin b = binary(lay, c.colh2o[lay], B.strrat * c.colo2[lay], 8.f, 9, true);
      for (int ig = 1; ig <= B.ng; ig++) {
        float tauray = c.colmol[lay] * (B.rayla(ig, b.js) + b.fs * (B.rayla(ig, b.js + 1) - B.rayla(ig, b.js)));
        o.taug[ngs + ig][lay] = b.speccomb * k8(B.absa, b, ig, 9) + c.colo3[lay] * B.abso3a(ig) +
                                c.colh2o[lay] * (selfk(B, lay, ig) + fork(B, lay, ig));
        if (lay == laysolfr) o.sfluxzen[ngs + ig] = sflx_eta(B, ig, b.js, b.fs);
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; lay++) {
      int ind0 = i0b(lay, 1), ind1 = i1b(lay, 1);
      for (int ig = 1; ig <= B.ng; ig++) {
        float tauray = c.colmol[lay] * B.raylb(ig);
        o.taug[ngs + ig][lay] = c.colo2[lay] * k4(B.absb, lay, ind0, ind1, ig) + c.colo3[lay] * B.abso3b(ig);
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    ngs += B.ng;
  }
  // ---- band 25 (SW:4177-4236): low h2o; o3 both
  {
    const SwBand &B = T.sw[9];
    laysolfr = laytrop;
    for (int lay = 1; lay <= laytrop; lay++) {
      if (c.jp[lay] < B.layreffr && c.jp[lay + 1] >= B.layreffr) laysolfr = std::min(lay + 1, laytrop);
      int ind0 = i0a(lay, 1), ind1 = i1a(lay, 1);
      for (int ig = 1; ig <= B.ng; ig++) {
        float tauray = c.colmol[lay] * B.raylg(ig);
        o.taug[ngs + ig][lay] = c.colh2o[lay] * k4(B.absa, lay, ind0, ind1, ig) + c.colo3[lay] * B.abso3a(ig);
        if (lay == laysolfr) o.sfluxzen[ngs + ig] = B.sfluxref(ig, 1);
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; lay++)
      for (int ig = 1; ig <= B.ng; ig++) {
        float tauray = c.colmol[lay] * B.raylg(ig);
        o.taug[ngs + ig][lay] = c.colo3[lay] * B.abso3b(ig);
        o.taur[ngs + ig][lay] = tauray;
      }
    ngs += B.ng;
  }
  // ---- band 26 (SW:4239-4287): rayleigh only
  {
    const SwBand &B = T.sw[10];
    laysolfr = laytrop;
    for (int lay = 1; lay <= nlayers; lay++)
      for (int ig = 1; ig <= B.ng; ig++) {
        if (lay <= laytrop && lay == laysolfr) o.sfluxzen[ngs + ig] = B.sfluxref(ig, 1);
        o.taug[ngs + ig][lay] = 0.f;
        o.taur[ngs + ig][lay] = c.colmol[lay] * B.raylg(ig);
      }
    ngs += B.ng;
  }
  // ---- band 27 (SW:4290-4355): o3 both
  {
    const SwBand &B = T.sw[11];
    for (int lay = 1; lay <= laytrop; lay++) {
      int ind0 = i0a(lay, 1), ind1 = i1a(lay, 1);
      for (int ig = 1; ig <= B.ng; ig++) {
        float tauray = c.colmol[lay] * B.raylg(ig);
        o.taug[ngs + ig][lay] = c.colo3[lay] * k4(B.absa, lay, ind0, ind1, ig);
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    laysolfr = nlayers;
    for (int lay = laytrop + 1; lay <= nlayers; lay++) {
      if (c.jp[lay - 1] < B.layreffr && c.jp[lay] >= B.layreffr) laysolfr = lay;
      int ind0 = i0b(lay, 1), ind1 = i1b(lay, 1);
      for (int ig = 1; ig <= B.ng; ig++) {
        float tauray = c.colmol[lay] * B.raylg(ig);
        o.taug[ngs + ig][lay] = c.colo3[lay] * k4(B.absb, lay, ind0, ind1, ig);
        if (lay == laysolfr) o.sfluxzen[ngs + ig] = B.scalekur * B.sfluxref(ig, 1);
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    ngs += B.ng;
  }
  // ---- band 28 (SW:4358-4460): o3,o2 both
  {
    const SwBand &B = T.sw[12];
    for (int lay = 1; lay <= laytrop; lay++) {
      Bin b = binary(lay, c.colo3[lay], B.strrat * c.colo2[lay], 8.f, 9, true);
      float tauray = c.colmol[lay] * B.rayl;
      for (int ig = 1; ig <= B.ng; ig++) {
        o.taug[ngs + ig][lay] = b.speccomb * k8(B.absa, b, ig, 9);
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    laysolfr = nlayers;
    for (int lay = laytrop + 1; lay <= nlayers; lay++) {
      if (c.jp[lay - 1] < B.layreffr && c.jp[lay] >= B.layreffr) laysolfr = lay;
      Bin b = binary(lay, c.colo3[lay], B.strrat * c.colo2[lay], 4.f, 5, false);
      float tauray = c.colmol[lay] * B.rayl;
      for (int ig = 1; ig <= B.ng; ig++) {
        o.taug[ngs + ig][lay] = b.speccomb * k8(B.absb, b, ig, 5);
        if (lay == laysolfr) o.sfluxzen[ngs + ig] = sflx_eta(B, ig, b.js, b.fs);
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    ngs += B.ng;
  }
  // ---- band 29 (SW:4463-4538): low h2o (+co2 cont); high co2 (+h2o cont)
  {
    const SwBand &B = T.sw[13];
    for (int lay = 1; lay <= laytrop; lay++) {
      int ind0 = i0a(lay, 1), ind1 = i1a(lay, 1);
      float tauray = c.colmol[lay] * B.rayl;
      for (int ig = 1; ig <= B.ng; ig++) {
        o.taug[ngs + ig][lay] = c.colh2o[lay] * ((k4(B.absa, lay, ind0, ind1, ig)) + selfk(B, lay, ig) + fork(B, lay, ig)) +
                                c.colco2[lay] * B.absco2(ig);
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    laysolfr = nlayers;
    for (int lay = laytrop + 1; lay <= nlayers; lay++) {
      if (c.jp[lay - 1] < B.layreffr && c.jp[lay] >= B.layreffr) laysolfr = lay;
      int ind0 = i0b(lay, 1), ind1 = i1b(lay, 1);
      float tauray = c.colmol[lay] * B.rayl;
      for (int ig = 1; ig <= B.ng; ig++) {
        o.taug[ngs + ig][lay] = c.colco2[lay] * k4(B.absb, lay, ind0, ind1, ig) + c.colh2o[lay] * B.absh2o(ig);
        if (lay == laysolfr) o.sfluxzen[ngs + ig] = B.sfluxref(ig, 1);
        o.taur[ngs + ig][lay] = tauray;
      }
    }
    ngs += B.ng;
  }
}

// ---- cldprmc_sw SW:1969-2390 (inflag >= 2 path; inflag 0 passthrough) -------------------------------
struct SwCld {
  // [ig][lay], 1-based
  float cldfmc[NGSW + 1][MXLAY], ciwpmc[NGSW + 1][MXLAY], clwpmc[NGSW + 1][MXLAY], cswpmc[NGSW + 1][MXLAY];
  float taucmc[NGSW + 1][MXLAY], ssacmc[NGSW + 1][MXLAY], asmcmc[NGSW + 1][MXLAY], fsfcmc[NGSW + 1][MXLAY], taormc[NGSW + 1][MXLAY];
  float reicmc[MXLAY], relqmc[MXLAY], resnmc[MXLAY];
};

int cldprmc_sw(const Tables &T, int nlayers, int inflag, int iceflag, int liqflag, SwCld &s, std::string &err) {
  const float eps = 1.e-06f, cldmin = 1.e-20f;
  (void)eps;
  const FArr &extliq1 = T.in.get("sw_extliq1"), &ssaliq1 = T.in.get("sw_ssaliq1"), &asyliq1 = T.in.get("sw_asyliq1");
  const FArr &extice3 = T.in.get("sw_extice3"), &ssaice3 = T.in.get("sw_ssaice3"), &asyice3 = T.in.get("sw_asyice3"),
             &fdlice3 = T.in.get("sw_fdlice3");
  static thread_local float extcoice[NGSW + 1], gice[NGSW + 1], ssacoice[NGSW + 1], forwice[NGSW + 1];
  static thread_local float extcoliq[NGSW + 1], gliq[NGSW + 1], ssacoliq[NGSW + 1], forwliq[NGSW + 1];
  static thread_local float extcosno[NGSW + 1], gsno[NGSW + 1], ssacosno[NGSW + 1], forwsno[NGSW + 1];
  for (int lay = 1; lay <= nlayers; lay++)
    for (int ig = 1; ig <= NGSW; ig++) s.taormc[ig][lay] = s.taucmc[ig][lay];
  for (int lay = 1; lay <= nlayers; lay++) {
    for (int ig = 1; ig <= NGSW; ig++) {
      float cwp = s.ciwpmc[ig][lay] + s.clwpmc[ig][lay] + s.cswpmc[ig][lay];
      if (s.cldfmc[ig][lay] >= cldmin && (cwp >= cldmin || s.taucmc[ig][lay] >= cldmin)) {
        if (inflag == 0) {
          float taucldorig_a = s.taucmc[ig][lay];
          float ffp = s.fsfcmc[ig][lay];
          float ffp1 = 1.0f - ffp;
          float ffpssa = 1.0f - ffp * s.ssacmc[ig][lay];
          float ssacloud_a = ffp1 * s.ssacmc[ig][lay] / ffpssa;
          float taucloud_a = ffpssa * taucldorig_a;
          s.taormc[ig][lay] = taucldorig_a;
          s.ssacmc[ig][lay] = ssacloud_a;
          s.taucmc[ig][lay] = taucloud_a;
          s.asmcmc[ig][lay] = (s.asmcmc[ig][lay] - ffp) / (ffp1);
        } else if (inflag == 1) {
          err = "INFLAG = 1 OPTION NOT AVAILABLE WITH MCICA"; return ARC_ERR_UNSUPPORTED;
        } else {
          float radice = s.reicmc[lay];
          int ib = T.sw_ngb[ig - 1];
          if ((s.ciwpmc[ig][lay] + s.cswpmc[ig][lay]) == 0.0f) {
            extcoice[ig] = 0.f; ssacoice[ig] = 0.f; gice[ig] = 0.f; forwice[ig] = 0.f;
            extcosno[ig] = 0.f; ssacosno[ig] = 0.f; gsno[ig] = 0.f; forwsno[ig] = 0.f;
          } else if (iceflag >= 3) {
            if (radice < 5.0f || radice > 140.0f) { err = "ERROR: ICE GENERALIZED EFFECTIVE SIZE OUT OF BOUNDS"; return ARC_ERR_RADIUS; }
            float factor = (radice - 2.f) / 3.f;
            int index = (int)factor;
            if (index == 46) index = 45;
            float fint = factor - (float)index;
            extcoice[ig] = extice3(index, ib) + fint * (extice3(index + 1, ib) - extice3(index, ib));
            ssacoice[ig] = ssaice3(index, ib) + fint * (ssaice3(index + 1, ib) - ssaice3(index, ib));
            gice[ig] = asyice3(index, ib) + fint * (asyice3(index + 1, ib) - asyice3(index, ib));
            float fdelta = fdlice3(index, ib) + fint * (fdlice3(index + 1, ib) - fdlice3(index, ib));
            if (fdelta < 0.0f) { err = "FDELTA LESS THAN 0.0"; return ARC_ERR_RADIUS; }
            if (fdelta > 1.0f) { err = "FDELTA GT THAN 1.0"; return ARC_ERR_RADIUS; }
            forwice[ig] = fdelta + 0.5f / ssacoice[ig];
            if (forwice[ig] > gice[ig]) forwice[ig] = gice[ig];
          } else {
            err = "oracle: iceflag < 3 not restated (WRF always uses 3,4,5; SW:10516-10576)"; return ARC_ERR_UNSUPPORTED;
          }
          if (s.cswpmc[ig][lay] > 0.0f && iceflag == 5) {
            float radsno = s.resnmc[lay];
            if (radsno < 5.0f || radsno > 140.0f) { err = "ERROR: SNOW GENERALIZED EFFECTIVE SIZE OUT OF BOUNDS"; return ARC_ERR_RADIUS; }
            float factor = (radsno - 2.f) / 3.f;
            int index = (int)factor;
            if (index == 46) index = 45;
            float fint = factor - (float)index;
            extcosno[ig] = extice3(index, ib) + fint * (extice3(index + 1, ib) - extice3(index, ib));
            ssacosno[ig] = ssaice3(index, ib) + fint * (ssaice3(index + 1, ib) - ssaice3(index, ib));
            gsno[ig] = asyice3(index, ib) + fint * (asyice3(index + 1, ib) - asyice3(index, ib));
            float fdelta = fdlice3(index, ib) + fint * (fdlice3(index + 1, ib) - fdlice3(index, ib));
            if (fdelta < 0.0f) { err = "FDELTA LESS THAN 0.0"; return ARC_ERR_RADIUS; }
            if (fdelta > 1.0f) { err = "FDELTA GT THAN 1.0"; return ARC_ERR_RADIUS; }
            forwsno[ig] = fdelta + 0.5f / ssacosno[ig];
            if (forwsno[ig] > gsno[ig]) forwsno[ig] = gsno[ig];
          } else {
            extcosno[ig] = 0.f; ssacosno[ig] = 0.f; gsno[ig] = 0.f; forwsno[ig] = 0.f;
          }
          if (s.clwpmc[ig][lay] == 0.0f) {
            extcoliq[ig] = 0.f; ssacoliq[ig] = 0.f; gliq[ig] = 0.f; forwliq[ig] = 0.f;
          } else if (liqflag == 1) {
            float radliq = s.relqmc[lay];
            if (radliq < 1.5f || radliq > 60.f) { err = "liquid effective radius out of bounds"; return ARC_ERR_RADIUS; }
            int index = (int)(radliq - 1.5f);
            if (index == 0) index = 1;
            if (index == 58) index = 57;
            float fint = radliq - 1.5f - (float)index;
            extcoliq[ig] = extliq1(index, ib) + fint * (extliq1(index + 1, ib) - extliq1(index, ib));
            ssacoliq[ig] = ssaliq1(index, ib) + fint * (ssaliq1(index + 1, ib) - ssaliq1(index, ib));
            if (fint < 0.f && ssacoliq[ig] > 1.f) ssacoliq[ig] = ssaliq1(index, ib);
            gliq[ig] = asyliq1(index, ib) + fint * (asyliq1(index + 1, ib) - asyliq1(index, ib));
            forwliq[ig] = gliq[ig] * gliq[ig];
          }
          float tauliqorig = s.clwpmc[ig][lay] * extcoliq[ig];
          float tauiceorig = s.ciwpmc[ig][lay] * extcoice[ig];
          float ssaliq = ssacoliq[ig] * (1.f - forwliq[ig]) / (1.f - forwliq[ig] * ssacoliq[ig]);
          float tauliq = (1.f - forwliq[ig] * ssacoliq[ig]) * tauliqorig;
          float ssaice = ssacoice[ig] * (1.f - forwice[ig]) / (1.f - forwice[ig] * ssacoice[ig]);
          float tauice = (1.f - forwice[ig] * ssacoice[ig]) * tauiceorig;
          float scatliq = ssaliq * tauliq, scatice = ssaice * tauice, scatsno;
          if (iceflag < 5) {
            s.taormc[ig][lay] = tauliqorig + tauiceorig;
            scatsno = 0.0f;
            s.taucmc[ig][lay] = tauliq + tauice;
          } else {
            float tausnoorig = s.cswpmc[ig][lay] * extcosno[ig];
            s.taormc[ig][lay] = tauliqorig + tauiceorig + tausnoorig;
            float ssasno = ssacosno[ig] * (1.f - forwsno[ig]) / (1.f - forwsno[ig] * ssacosno[ig]);
            float tausno = (1.f - forwsno[ig] * ssacosno[ig]) * tausnoorig;
            scatsno = ssasno * tausno;
            s.taucmc[ig][lay] = tauliq + tauice + tausno;
          }
          if (s.taucmc[ig][lay] == 0.f) s.taucmc[ig][lay] = cldmin;
          if (scatice == 0.f) scatice = cldmin;
          if (scatsno == 0.f) scatsno = cldmin;
          if (iceflag < 5) s.ssacmc[ig][lay] = (scatliq + scatice) / s.taucmc[ig][lay];
          else s.ssacmc[ig][lay] = (scatliq + scatice + scatsno) / s.taucmc[ig][lay];
          if (iceflag == 3 || iceflag == 4) {
            s.asmcmc[ig][lay] = (1.0f / (scatliq + scatice)) *
                                (scatliq * (gliq[ig] - forwliq[ig]) / (1.0f - forwliq[ig]) +
                                 scatice * ((gice[ig] - forwice[ig]) / (1.0f - forwice[ig])));
          } else if (iceflag == 5) {
            s.asmcmc[ig][lay] = (1.0f / (scatliq + scatice + scatsno)) *
                                (scatliq * (gliq[ig] - forwliq[ig]) / (1.0f - forwliq[ig]) +
                                 scatice * ((gice[ig] - forwice[ig]) / (1.0f - forwice[ig])) +
                                 scatsno * ((gsno[ig] - forwsno[ig]) / (1.0f - forwsno[ig])));
          }
        }
      }
    }
  }
  return 0;
}

static thread_local float g_min_cond = 1.f;   // conditioning diagnostic, see ArcDebug.sw_cond

// ---- reftra_sw SW:2422-2701 (kmodts = 2) -------------------------------------------------------------
void reftra_sw(const Tables &T, int nlayers, const bool *lrtchk, const float *pgg, float prmuz, const float *ptau,
               const float *pw, float *pref, float *prefd, float *ptra, float *ptrad) {
  const float eps = 1.e-08f, zwcrit = 0.9999995f, od_lo = 0.06f, tblint = 10000.0f;
  const float bpade = T.sw_bpade;
  const float *exp_tbl = T.sw_exp_tbl.data();
  for (int jk = 1; jk <= nlayers; jk++) {
    if (!lrtchk[jk]) {
      pref[jk] = 0.f; ptra[jk] = 1.f; prefd[jk] = 0.f; ptrad[jk] = 1.f;
    } else {
      float zto1 = ptau[jk], zw = pw[jk], zg = pgg[jk];
      float zg3 = 3.f * zg;
      float zgamma1 = (8.f - zw * (5.f + zg3)) * 0.25f;
      float zgamma2 = 3.f * (zw * (1.f - zg)) * 0.25f;
      float zgamma3 = (2.f - zg3 * prmuz) * 0.25f;
      float zgamma4 = 1.f - zgamma3;
      float q = zg / (1.f - zg);
      float denom = std::max((1.f - (1.f - zw) * (q * q)), 1.0E-30f);
      float zwo = zw / denom;
      if (zwo >= zwcrit) {
        float za = zgamma1 * prmuz;
        float za1 = za - zgamma3;
        float zgt = zgamma1 * zto1;
        float ze1 = std::min(zto1 / prmuz, 500.f), ze2;
        if (ze1 <= od_lo) ze2 = 1.f - ze1 + 0.5f * ze1 * ze1;
        else { float tblind = ze1 / (bpade + ze1); int itind = (int)(tblint * tblind + 0.5f); ze2 = exp_tbl[itind]; }
        pref[jk] = (zgt - za1 * (1.f - ze2)) / (1.f + zgt);
        ptra[jk] = 1.f - pref[jk];
        prefd[jk] = zgt / (1.f + zgt);
        ptrad[jk] = 1.f - prefd[jk];
        if (ze2 == 1.0f) { pref[jk] = 0.f; ptra[jk] = 1.f; prefd[jk] = 0.f; ptrad[jk] = 1.f; }
      } else {
        float za1 = zgamma1 * zgamma4 + zgamma2 * zgamma3;
        float za2 = zgamma1 * zgamma3 + zgamma2 * zgamma4;
        float zrk = sqrtf(zgamma1 * zgamma1 - zgamma2 * zgamma2);
        float zrp = zrk * prmuz;
        float zrp1 = 1.f + zrp, zrm1 = 1.f - zrp, zrk2 = 2.f * zrk;
        float zrpp = 1.f - zrp * zrp;
        if (fabsf(zrpp) < g_min_cond) g_min_cond = fabsf(zrpp);
        float zrkg = zrk + zgamma1;
        float zr1 = zrm1 * (za2 + zrk * zgamma3);
        float zr2 = zrp1 * (za2 - zrk * zgamma3);
        float zr3 = zrk2 * (zgamma3 - za2 * prmuz);
        float zr4 = zrpp * zrkg;
        float zr5 = zrpp * (zrk - zgamma1);
        float zt1 = zrp1 * (za1 + zrk * zgamma4);
        float zt2 = zrm1 * (za1 - zrk * zgamma4);
        float zt3 = zrk2 * (zgamma4 + za1 * prmuz);
        float zt4 = zr4, zt5 = zr5;
        float zbeta = (zgamma1 - zrk) / zrkg;
        float ze1 = std::min(zrk * zto1, 500.f);
        float ze2 = std::min(zto1 / prmuz, 500.f);
        float zem1, zep1, zem2, zep2;
        if (ze1 <= od_lo) { zem1 = 1.f - ze1 + 0.5f * ze1 * ze1; zep1 = 1.f / zem1; }
        else { float tblind = ze1 / (bpade + ze1); int itind = (int)(tblint * tblind + 0.5f); zem1 = exp_tbl[itind]; zep1 = 1.f / zem1; }
        if (ze2 <= od_lo) { zem2 = 1.f - ze2 + 0.5f * ze2 * ze2; zep2 = 1.f / zem2; }
        else { float tblind = ze2 / (bpade + ze2); int itind = (int)(tblint * tblind + 0.5f); zem2 = exp_tbl[itind]; zep2 = 1.f / zem2; }
        float zdenr = zr4 * zep1 + zr5 * zem1;
        float zdent = zt4 * zep1 + zt5 * zem1;
        if (zdenr >= -eps && zdenr <= eps) { pref[jk] = eps; ptra[jk] = zem2; }
        else {
          pref[jk] = zw * (zr1 * zep1 - zr2 * zem1 - zr3 * zem2) / zdenr;
          ptra[jk] = zem2 - zem2 * zw * (zt1 * zep1 - zt2 * zem1 - zt3 * zep2) / zdent;
        }
        float zemm = zem1 * zem1;
        float zdend = 1.f / ((1.f - zbeta * zemm) * zrkg);
        prefd[jk] = zgamma2 * (1.f - zemm) * zdend;
        ptrad[jk] = zrk2 * zem1 * zdend;
      }
    }
  }
}

// ---- vrtqdr_sw SW:7922-8046 -----------------------------------------------------------------------------
void vrtqdr_sw(int klev, const float *pref, const float *prefd, const float *ptra, const float *ptrad, const float *pdbt,
               float *prdnd, float *prup, float *prupd, const float *ptdbt, float *pfd, float *pfu) {
  float ztdn[MXLAY + 2];
  float zreflect = 1.f / (1.f - prefd[klev + 1] * prefd[klev]);
  prup[klev] = pref[klev] + (ptrad[klev] * ((ptra[klev] - pdbt[klev]) * prefd[klev + 1] + pdbt[klev] * pref[klev + 1])) * zreflect;
  prupd[klev] = prefd[klev] + ptrad[klev] * ptrad[klev] * prefd[klev + 1] * zreflect;
  for (int jk = 1; jk <= klev - 1; jk++) {
    int ikp = klev + 1 - jk, ikx = ikp - 1;
    zreflect = 1.f / (1.f - prupd[ikp] * prefd[ikx]);
    prup[ikx] = pref[ikx] + (ptrad[ikx] * ((ptra[ikx] - pdbt[ikx]) * prupd[ikp] + pdbt[ikx] * prup[ikp])) * zreflect;
    prupd[ikx] = prefd[ikx] + ptrad[ikx] * ptrad[ikx] * prupd[ikp] * zreflect;
  }
  ztdn[1] = 1.f; prdnd[1] = 0.f; ztdn[2] = ptra[1]; prdnd[2] = prefd[1];
  for (int jk = 2; jk <= klev; jk++) {
    int ikp = jk + 1;
    zreflect = 1.f / (1.f - prefd[jk] * prdnd[jk]);
    ztdn[ikp] = ptdbt[jk] * ptra[jk] + (ptrad[jk] * ((ztdn[jk] - ptdbt[jk]) + ptdbt[jk] * pref[jk] * prdnd[jk])) * zreflect;
    prdnd[ikp] = prefd[jk] + ptrad[jk] * ptrad[jk] * prdnd[jk] * zreflect;
  }
  for (int jk = 1; jk <= klev + 1; jk++) {
    zreflect = 1.f / (1.f - prdnd[jk] * prupd[jk]);
    pfu[jk] = (ptdbt[jk] * prup[jk] + (ztdn[jk] - ptdbt[jk]) * prupd[jk]) * zreflect;
    pfd[jk] = ptdbt[jk] + (ztdn[jk] - ptdbt[jk] + ptdbt[jk] * prup[jk] * prdnd[jk]) * zreflect;
  }
}

struct SwFlux14 {
  float bbfd[MXLAY + 2], bbfu[MXLAY + 2], bbcd[MXLAY + 2], bbcu[MXLAY + 2], uvfd[MXLAY + 2], uvcd[MXLAY + 2], nifd[MXLAY + 2],
      nicd[MXLAY + 2], bbfddir[MXLAY + 2], bbcddir[MXLAY + 2], uvfddir[MXLAY + 2], uvcddir[MXLAY + 2], nifddir[MXLAY + 2],
      nicddir[MXLAY + 2];
};

// ---- spcvmc_sw SW:8083-8658 (icpr = 1, iout = 0) --------------------------------------------------------
void spcvmc_sw(const Tables &T, int nlayers, const float *palbd, const float *palbp, const SwCld &cl,
               const float (*ptaua)[MXLAY], const float (*pasya)[MXLAY], const float (*pomga)[MXLAY], float prmu0,
               const float *adjflux, const SwCoef &c, SwFlux14 &F, SwTau &tau) {
  const float od_lo = 0.06f, tblint = 10000.0f, repclc = 1.e-12f;
  const float bpade = T.sw_bpade;
  const float *exp_tbl = T.sw_exp_tbl.data();
  const int klev = nlayers;
  bool lrtchkclr[MXLAY], lrtchkcld[MXLAY];
  float zdbt[MXLAY + 2], zdbt_nodel[MXLAY + 2], zgcc[MXLAY], zgco[MXLAY], zomcc[MXLAY], zomco[MXLAY];
  float zrdnd[MXLAY + 2], zrdndc[MXLAY + 2], zref[MXLAY + 2], zrefc[MXLAY + 2], zrefo[MXLAY + 2], zrefd[MXLAY + 2],
      zrefdc[MXLAY + 2], zrefdo[MXLAY + 2];
  float zrup[MXLAY + 2], zrupd[MXLAY + 2], zrupc[MXLAY + 2], zrupdc[MXLAY + 2], ztauc[MXLAY], ztauo[MXLAY];
  float ztdbt[MXLAY + 2], ztra[MXLAY + 2], ztrac[MXLAY + 2], ztrao[MXLAY + 2], ztrad[MXLAY + 2], ztradc[MXLAY + 2],
      ztrado[MXLAY + 2];
  float zdbtc[MXLAY + 2], ztdbtc[MXLAY + 2], zdbtc_nodel[MXLAY + 2], ztdbt_nodel[MXLAY + 2], ztdbtc_nodel[MXLAY + 2];
  float zcd[MXLAY + 2], zcu[MXLAY + 2], zfd[MXLAY + 2], zfu[MXLAY + 2];
  for (int jk = 1; jk <= klev + 1; jk++) {
    F.bbcd[jk] = 0.f; F.bbcu[jk] = 0.f; F.bbfd[jk] = 0.f; F.bbfu[jk] = 0.f; F.bbcddir[jk] = 0.f; F.bbfddir[jk] = 0.f;
    F.uvcd[jk] = 0.f; F.uvfd[jk] = 0.f; F.uvcddir[jk] = 0.f; F.uvfddir[jk] = 0.f;
    F.nicd[jk] = 0.f; F.nifd[jk] = 0.f; F.nicddir[jk] = 0.f; F.nifddir[jk] = 0.f;
  }
  taumol_sw(T, klev, c, tau);
  auto expt = [&](float ze1) {
    if (ze1 <= od_lo) return 1.f - ze1 + 0.5f * ze1 * ze1;
    float tblind = ze1 / (bpade + ze1);
    int itind = (int)(tblint * tblind + 0.5f);
    return exp_tbl[itind];
  };
  int iw = 0;
  for (int jb = 16; jb <= 29; jb++) {
    int ibm = jb - 15;
    int igt = T.sw_ngc[ibm - 1];
    for (int jg = 1; jg <= igt; jg++) {
      iw = iw + 1;
      float zincflx = adjflux[jb] * tau.sfluxzen[iw] * prmu0;
      ztdbtc[1] = 1.0f; ztdbtc_nodel[1] = 1.0f;
      zdbtc[klev + 1] = 0.0f; ztrac[klev + 1] = 0.0f; ztradc[klev + 1] = 0.0f;
      zrefc[klev + 1] = palbp[ibm]; zrefdc[klev + 1] = palbd[ibm]; zrupc[klev + 1] = palbp[ibm]; zrupdc[klev + 1] = palbd[ibm];
      ztdbt[1] = 1.0f; ztdbt_nodel[1] = 1.0f;
      zdbt[klev + 1] = 0.0f; ztra[klev + 1] = 0.0f; ztrad[klev + 1] = 0.0f;
      zref[klev + 1] = palbp[ibm]; zrefd[klev + 1] = palbd[ibm]; zrup[klev + 1] = palbp[ibm]; zrupd[klev + 1] = palbd[ibm];
      for (int jk = 1; jk <= klev; jk++) {
        int ikl = klev + 1 - jk;
        lrtchkclr[jk] = true;
        lrtchkcld[jk] = (cl.cldfmc[iw][ikl] > repclc);
        ztauc[jk] = tau.taur[iw][ikl] + tau.taug[iw][ikl] + ptaua[ibm][ikl];
        zomcc[jk] = tau.taur[iw][ikl] * 1.0f + ptaua[ibm][ikl] * pomga[ibm][ikl];
        zgcc[jk] = pasya[ibm][ikl] * pomga[ibm][ikl] * ptaua[ibm][ikl] / zomcc[jk];
        zomcc[jk] = zomcc[jk] / ztauc[jk];
        float zclear = 1.0f - cl.cldfmc[iw][ikl];
        float zcloud = cl.cldfmc[iw][ikl];
        float zdbtmc = expt(ztauc[jk] / prmu0);
        zdbtc_nodel[jk] = zdbtmc;
        ztdbtc_nodel[jk + 1] = zdbtc_nodel[jk] * ztdbtc_nodel[jk];
        float tauorig = ztauc[jk] + cl.taormc[iw][ikl];
        float zdbtmo = expt(tauorig / prmu0);
        zdbt_nodel[jk] = zclear * zdbtmc + zcloud * zdbtmo;
        ztdbt_nodel[jk + 1] = zdbt_nodel[jk] * ztdbt_nodel[jk];
        float zf = zgcc[jk] * zgcc[jk];
        float zwf = zomcc[jk] * zf;
        ztauc[jk] = (1.0f - zwf) * ztauc[jk];
        float denom = std::max((1.0f - zwf), 1.0E-30f);
        zomcc[jk] = (zomcc[jk] - zwf) / denom;
        denom = std::max((1.0f - zf), 1.0E-30f);
        zgcc[jk] = (zgcc[jk] - zf) / denom;
        // icpr >= 1
        ztauo[jk] = ztauc[jk] + cl.taucmc[iw][ikl];
        zomco[jk] = ztauc[jk] * zomcc[jk] + cl.taucmc[iw][ikl] * cl.ssacmc[iw][ikl];
        zgco[jk] = (cl.taucmc[iw][ikl] * cl.ssacmc[iw][ikl] * cl.asmcmc[iw][ikl] + ztauc[jk] * zomcc[jk] * zgcc[jk]) / zomco[jk];
        zomco[jk] = zomco[jk] / ztauo[jk];
      }
      reftra_sw(T, klev, lrtchkclr, zgcc, prmu0, ztauc, zomcc, zrefc, zrefdc, ztrac, ztradc);
      reftra_sw(T, klev, lrtchkcld, zgco, prmu0, ztauo, zomco, zrefo, zrefdo, ztrao, ztrado);
      for (int jk = 1; jk <= klev; jk++) {
        int ikl = klev + 1 - jk;
        float zclear = 1.0f - cl.cldfmc[iw][ikl];
        float zcloud = cl.cldfmc[iw][ikl];
        zref[jk] = zclear * zrefc[jk] + zcloud * zrefo[jk];
        zrefd[jk] = zclear * zrefdc[jk] + zcloud * zrefdo[jk];
        ztra[jk] = zclear * ztrac[jk] + zcloud * ztrao[jk];
        ztrad[jk] = zclear * ztradc[jk] + zcloud * ztrado[jk];
        float zdbtmc = expt(ztauc[jk] / prmu0);
        zdbtc[jk] = zdbtmc;
        ztdbtc[jk + 1] = zdbtc[jk] * ztdbtc[jk];
        float zdbtmo = expt(ztauo[jk] / prmu0);
        zdbt[jk] = zclear * zdbtmc + zcloud * zdbtmo;
        ztdbt[jk + 1] = zdbt[jk] * ztdbt[jk];
      }
      vrtqdr_sw(klev, zrefc, zrefdc, ztrac, ztradc, zdbtc, zrdndc, zrupc, zrupdc, ztdbtc, zcd, zcu);
      vrtqdr_sw(klev, zref, zrefd, ztra, ztrad, zdbt, zrdnd, zrup, zrupd, ztdbt, zfd, zfu);
      for (int jk = 1; jk <= klev + 1; jk++) {
        int ikl = klev + 2 - jk;
        F.bbfu[ikl] = F.bbfu[ikl] + zincflx * zfu[jk];
        F.bbfd[ikl] = F.bbfd[ikl] + zincflx * zfd[jk];
        F.bbcu[ikl] = F.bbcu[ikl] + zincflx * zcu[jk];
        F.bbcd[ikl] = F.bbcd[ikl] + zincflx * zcd[jk];
        F.bbfddir[ikl] = F.bbfddir[ikl] + zincflx * ztdbt_nodel[jk];
        F.bbcddir[ikl] = F.bbcddir[ikl] + zincflx * ztdbtc_nodel[jk];
        if (ibm >= 10 && ibm <= 13) {
          F.uvcd[ikl] = F.uvcd[ikl] + zincflx * zcd[jk];
          F.uvfd[ikl] = F.uvfd[ikl] + zincflx * zfd[jk];
          F.uvcddir[ikl] = F.uvcddir[ikl] + zincflx * ztdbtc_nodel[jk];
          F.uvfddir[ikl] = F.uvfddir[ikl] + zincflx * ztdbt_nodel[jk];
        } else if (ibm == 14 || ibm <= 9) {
          F.nicd[ikl] = F.nicd[ikl] + zincflx * zcd[jk];
          F.nifd[ikl] = F.nifd[ikl] + zincflx * zfd[jk];
          F.nicddir[ikl] = F.nicddir[ikl] + zincflx * ztdbtc_nodel[jk];
          F.nifddir[ikl] = F.nifddir[ikl] + zincflx * ztdbt_nodel[jk];
        }
      }
    }
  }
}

}  // namespace

// Column-level result of rrtmg_sw for ncol = 1 (levels 1..nlay+1, 1 = surface)
struct SwColumnOut {
  float swuflx[MXLAY + 2], swdflx[MXLAY + 2], swhr[MXLAY + 2], swuflxc[MXLAY + 2], swdflxc[MXLAY + 2], swhrc[MXLAY + 2];
  float swuflxcln[MXLAY + 2], swdflxcln[MXLAY + 2], swuflxclnc[MXLAY + 2], swdflxclnc[MXLAY + 2];
  float sibvisdir[MXLAY + 2], sibvisdif[MXLAY + 2], sibnirdir[MXLAY + 2], sibnirdif[MXLAY + 2], swdkdir[MXLAY + 2], swdkdif[MXLAY + 2];
};

struct SwWork {
  SwCld cld; SwTau tau; SwCoef coef; SwFlux14 F, Fcln;
  float taua[NBSW + 1][MXLAY], ssaa[NBSW + 1][MXLAY], asma[NBSW + 1][MXLAY], tauacln[NBSW + 1][MXLAY];
  float wkl[8][MXLAY];
};

// rrtmg_sw SW:8740-9490 with inatm_sw SW:9520-9873 inlined (ncol = 1, iaer = 10, icld = 2)
static int rrtmg_sw_column(const Tables &T, int nlay, const float *play, const float *plev, const float *tlay,
                           const float *h2ovmr, const float *o3vmr, float co2vmr, float ch4vmr, float n2ovmr, float o2vmr,
                           float asdir, float asdif, float aldir, float aldif, float coszen, float adjes, int dyofyr, float scon,
                           int inflgsw, int iceflgsw, int liqflgsw, int clean_atm_diag, SwWork &W, SwColumnOut &out,
                           std::string &err) {
  const float zepzen = 1.e-10f;
  const int nlayers = nlay;
  const float amd = 28.9660f, amw = 18.0160f;
  float pavel[MXLAY], tavel[MXLAY], pz[MXLAY + 1], pdp[MXLAY], coldry[MXLAY], adjflux[30];
  // inatm_sw
  float adjflx = adjes;
  if (dyofyr > 0) {
    float gamma = 2.f * T.pi * (dyofyr - 1) / 365.f;
    adjflx = 1.000110f + .034221f * cosf(gamma) + .001289f * sinf(gamma) + .000719f * cosf(2.f * gamma) + .000077f * sinf(2.f * gamma);
  }
  for (int ib = 16; ib <= 29; ib++) { float solvar = scon / 1.36822e+03f; adjflux[ib] = adjflx * solvar; }
  pz[0] = plev[1];
  for (int m = 0; m < 8; m++) for (int l = 0; l < MXLAY; l++) W.wkl[m][l] = 0.f;
  for (int l = 1; l <= nlayers; l++) {
    pavel[l] = play[l]; tavel[l] = tlay[l]; pz[l] = plev[l + 1];
    pdp[l] = pz[l - 1] - pz[l];
    W.wkl[1][l] = h2ovmr[l]; W.wkl[2][l] = co2vmr; W.wkl[3][l] = o3vmr[l]; W.wkl[4][l] = n2ovmr; W.wkl[6][l] = ch4vmr; W.wkl[7][l] = o2vmr;
    float amm = (1.f - W.wkl[1][l]) * amd + W.wkl[1][l] * amw;
    coldry[l] = (pz[l - 1] - pz[l]) * 1.e3f * T.avogad / (1.e2f * T.grav * amm * (1.f + W.wkl[1][l]));
  }
  for (int l = 1; l <= nlayers; l++)
    for (int imol = 1; imol <= 7; imol++) W.wkl[imol][l] = coldry[l] * W.wkl[imol][l];
  int rc = cldprmc_sw(T, nlayers, inflgsw, iceflgsw, liqflgsw, W.cld, err);
  if (rc) return rc;
  setcoef_sw(T, nlayers, pavel, tavel, coldry, W.wkl, W.coef);
  float cossza = coszen;
  if (cossza <= zepzen) cossza = zepzen;
  float albdir[NBSW + 1], albdif[NBSW + 1];
  for (int ib = 1; ib <= 9; ib++) { albdir[ib] = aldir; albdif[ib] = aldif; }
  albdir[NBSW] = aldir; albdif[NBSW] = aldif;
  for (int ib = 10; ib <= 13; ib++) { albdir[ib] = asdir; albdif[ib] = asdif; }
  for (int ib = 1; ib <= NBSW; ib++) for (int l = 1; l <= nlayers; l++) W.tauacln[ib][l] = 0.0f;
  spcvmc_sw(T, nlayers, albdif, albdir, W.cld, W.taua, W.asma, W.ssaa, cossza, adjflux, W.coef, W.F, W.tau);
  float swnflxc[MXLAY + 2], swnflx[MXLAY + 2];
  for (int i = 1; i <= nlayers + 1; i++) {
    out.swuflxc[i] = W.F.bbcu[i]; out.swdflxc[i] = W.F.bbcd[i]; out.swuflx[i] = W.F.bbfu[i]; out.swdflx[i] = W.F.bbfd[i];
    float dirdflux = W.F.bbfddir[i];
    float difdflux = out.swdflx[i] - dirdflux;
    out.swdkdir[i] = dirdflux; out.swdkdif[i] = difdflux;
    float dirdnuv = W.F.uvfddir[i];
    float difdnuv = W.F.uvfd[i] - dirdnuv;
    out.sibvisdir[i] = dirdnuv; out.sibvisdif[i] = difdnuv;
    float dirdnir = W.F.nifddir[i];
    float difdnir = W.F.nifd[i] - dirdnir;
    out.sibnirdir[i] = dirdnir; out.sibnirdif[i] = difdnir;
  }
  for (int i = 1; i <= nlayers + 1; i++) { swnflxc[i] = out.swdflxc[i] - out.swuflxc[i]; swnflx[i] = out.swdflx[i] - out.swuflx[i]; }
  for (int i = 1; i <= nlayers; i++) {
    float zdpgcp = T.heatfac / pdp[i];
    out.swhrc[i] = (swnflxc[i + 1] - swnflxc[i]) * zdpgcp;
    out.swhr[i] = (swnflx[i + 1] - swnflx[i]) * zdpgcp;
  }
  out.swhrc[nlayers] = 0.f; out.swhr[nlayers] = 0.f;
  if (clean_atm_diag > 0) {
    // second call with ztauacln = 0 (SW:9436-9470); its clear-sky by-product is the clean-clear stream
    SwTau &tau2 = W.tau;
    spcvmc_sw(T, nlayers, albdif, albdir, W.cld, W.tauacln, W.asma, W.ssaa, cossza, adjflux, W.coef, W.Fcln, tau2);
    for (int i = 1; i <= nlayers + 1; i++) {
      out.swuflxcln[i] = W.Fcln.bbfu[i]; out.swdflxcln[i] = W.Fcln.bbfd[i];
      out.swuflxclnc[i] = W.Fcln.bbcu[i]; out.swdflxclnc[i] = W.Fcln.bbcd[i];
    }
  } else {
    for (int i = 1; i <= nlayers + 1; i++) { out.swuflxcln[i] = 0.f; out.swdflxcln[i] = 0.f; out.swuflxclnc[i] = 0.f; out.swdflxclnc[i] = 0.f; }
  }
  return 0;
}

// ---- RRTMG_SWRAD SW:9901-11207 -------------------------------------------------------------------------
int oracle_swrad(const ArcDims &d, const ArcSwIn &in, ArcSwOut &out, ArcDebug *dbg, std::string &err) {
  const Tables &T = tables();
  if (!T.ready) { err = "oracle not initialised"; return ARC_ERR_NOT_INIT; }
  if (in.aer_ra_feedback == 1) {
    if (!(in.tauaer300 && in.tauaer400 && in.tauaer600 && in.tauaer999 && in.gaer300 && in.gaer400 && in.gaer600 && in.gaer999 &&
          in.waer300 && in.waer400 && in.waer600 && in.waer999)) {
      err = "Warning: missing fields required for aerosol radiation"; return ARC_ERR_MISSING_FIELD;
    }
  }
  if (in.aer_opt == 1) {
    // iaer = 6 (SW:9201-9205): the six ECMWF aerosol types with the optical depths AEROD(i,k,j,1:6).  The reference's second
    // ("clean") spcvmc_sw call reads ztauacln, which only the iaer = 10 branch defines (SW:9343-9352): undefined there, refused here.
    if (!in.aerod || in.no_src < 6) { err = "aer_opt=1 needs aerod(i,k,j,1:6) (no_src >= 6)"; return ARC_ERR_MISSING_FIELD; }
    if (in.clean_atm_diag > 0) { err = "aer_opt=1 with clean_atm_diag: the reference leaves the clean call's aerosol optical depth undefined"; return ARC_ERR_UNSUPPORTED; }
  }
  const int clean = in.clean_atm_diag;
  const Idx ix(d);
  const int kts = d.kts, kte = d.kte, nz = kte - kts + 1;
  const int nlay = nz + 1;
  if (nlay + 2 >= MXLAY) { err = "too many layers for oracle"; return ARC_ERR_BAD_ARG; }
  const FArr &retab = T.in.get("lw_retab");
  const FArr &wavemin = T.in.get("sw_wavemin"), &wavemax = T.in.get("sw_wavemax");
  const float co2 = 379.e-6f, ch4 = 1774.e-9f, n2o = 319.e-9f, o2 = 0.209488f;
  const float amdw = 1.607793f, amdo = 0.603461f;
  const float thresh = 1.e-9f;
  const int nci = d.ite - d.its + 1;
  static thread_local SwWork *Wp = nullptr;
  if (!Wp) Wp = new SwWork;
  SwWork &W = *Wp;
  CloudIn ci{in.icloud, in.warm_rain, in.is_cammgmp_used, in.has_reqc, in.has_reqi, in.has_reqs, in.progn,
             in.f_qv, in.f_qc, in.f_qr, in.f_qi, in.f_qs, in.f_qg, in.f_qndrop, in.g,
             in.t3d, in.cldfra3d, in.lradius, in.iradius, in.qv3d, in.qc3d, in.qr3d, in.qi3d, in.qs3d, in.qg3d, in.qndrop3d,
             in.re_cloud, in.re_ice, in.re_snow, in.f_ice_phy, in.xland, in.xice, in.snow};
  for (int j = d.jts; j <= d.jte; j++) {
    for (int i = d.its; i <= d.ite; i++) {
      const size_t ij = ix.at2(i, j);
      const size_t c = (size_t)(j - d.jts) * nci + (i - d.its);
      out.coszr[ij] = in.xcoszen[ij];
      float coszrs = in.xcoszen[ij];
      bool dorrsw = !(coszrs <= 0.0f);
      if (dbg && dbg->laytrop && !dorrsw) dbg->laytrop[c] = -1;
      if (dorrsw) {
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
        const int inflgsw = col.inflg, iceflgsw = col.iceflg, liqflgsw = col.liqflg;
        float plev[MXLAY + 2], tlev[MXLAY + 2], play[MXLAY], tlay[MXLAY], pdel[MXLAY], h2ovmr[MXLAY], o3vmr[MXLAY], o3mmr[MXLAY];
        plev[1] = pw1d[1]; tlev[1] = tw1d[1];
        float tsfc = in.tsk[ij]; (void)tsfc;
        for (int k = 1; k <= nz; k++) {
          play[k] = p1d[k]; plev[k + 1] = pw1d[k + 1]; pdel[k] = plev[k] - plev[k + 1];
          tlay[k] = t1d[k]; tlev[k + 1] = tw1d[k + 1];
          h2ovmr[k] = col.qv[k] * amdw;
        }
        play[nz + 1] = 0.5f * plev[nz + 1];
        tlay[nz + 1] = tlev[nz + 1] + 0.0f;
        plev[nz + 2] = 1.0e-5f;
        tlev[nz + 2] = tlev[nz + 1] + 0.0f;
        h2ovmr[nz + 1] = h2ovmr[nz];
        o3data(T.in, plev, nz + 1, o3mmr);
        for (int k = 1; k <= nz + 1; k++) {
          o3vmr[k] = o3mmr[k] * amdo;
          if (in.o33d && in.o3input == 2) {
            if (k <= nz) o3vmr[k] = o31d[k];
            else {
              o3vmr[k] = o31d[nz] - o3mmr[nz] * amdo + o3mmr[k] * amdo;
              if (o3vmr[k] <= 0.f) o3vmr[k] = o3mmr[k] * amdo;
            }
          }
        }
        float asdir, asdif, aldir, aldif;
        if (in.sf_surface_physics == 8 && in.xland[ij] < 1.5f) {
          asdir = in.alswvisdir[ij]; asdif = in.alswvisdif[ij]; aldir = in.alswnirdir[ij]; aldif = in.alswnirdif[ij];
        } else { asdir = asdif = aldir = aldif = in.albedo[ij]; }
        CloudPaths cp;
        cloud_paths(ci, ix, i, j, kts, kte, col, pdel, tlay, retab, cp);
        cp.clwp[nz + 1] = 0.f; cp.ciwp[nz + 1] = 0.f; cp.cswp[nz + 1] = 0.f;
        cp.rel[nz + 1] = 10.f; cp.rei[nz + 1] = 10.f; cp.res[nz + 1] = 10.f; cp.cldfrac[nz + 1] = 0.f;
        // mcica_subcol_sw (permuteseed = 1, irng = 0), SW:1392-1513
        float pmid[MXLAY];
        for (int l = 1; l <= nlay; l++) pmid[l] = play[l] * 1.e2f;
        std::vector<float> cdf; std::vector<unsigned char> cloudy;
        mcica_mask(nlay, NGSW, pmid, cp.cldfrac, 1, cdf, cloudy);
        for (int l = 1; l <= nlay; l++) {
          for (int ig = 1; ig <= NGSW; ig++) {
            bool cl = cloudy[(size_t)(ig - 1) * (nlay + 1) + l] != 0;
            W.cld.cldfmc[ig][l] = cl ? 1.f : 0.f;
            W.cld.clwpmc[ig][l] = cl ? cp.clwp[l] : 0.f;
            W.cld.ciwpmc[ig][l] = cl ? cp.ciwp[l] : 0.f;
            W.cld.cswpmc[ig][l] = (cl && iceflgsw == 5) ? cp.cswp[l] : 0.f;   // inatm_sw copies cswp only for iceflag 5 (SW:9852)
            W.cld.taucmc[ig][l] = 0.f;     // taucld = 0 (SW:10889-10896) -> tauc_stoch = 0
            W.cld.ssacmc[ig][l] = 1.f;
            W.cld.asmcmc[ig][l] = 0.f;
            W.cld.fsfcmc[ig][l] = 0.f;
          }
          W.cld.reicmc[l] = cp.rei[l]; W.cld.relqmc[l] = cp.rel[l];
          W.cld.resnmc[l] = iceflgsw == 5 ? cp.res[l] : 0.f;
        }
        // aerosol: 14-band optics from the four chem wavelengths, SW:10967-11071
        for (int nb = 1; nb <= NBSW; nb++)
          for (int k = 1; k <= nz + 1; k++) { W.taua[nb][k] = 0.f; W.ssaa[nb][k] = 1.f; W.asma[nb][k] = 0.f; }
        if (in.tauaer3d_sw) {
          for (int nb = 1; nb <= NBSW; nb++)
            for (int k = 1; k <= nz; k++) {
              size_t q = ix.at4(i, kts + k - 1, j, nb - 1);
              W.taua[nb][k] = in.tauaer3d_sw[q]; W.ssaa[nb][k] = in.ssaaer3d_sw[q]; W.asma[nb][k] = in.asyaer3d_sw[q];
            }
        }
        if (in.aer_ra_feedback == 1) {
          for (int nb = 1; nb <= NBSW; nb++) {
            float wavemid = 0.5f * (wavemin(nb) + wavemax(nb));
            for (int k = 1; k <= nz; k++) {
              size_t q = ix.at3(i, kts + k - 1, j);
              if (in.tauaer300[q] > thresh && in.tauaer999[q] > thresh) {
                float ang = logf(in.tauaer300[q] / in.tauaer999[q]) / logf(999.f / 300.f);
                W.taua[nb][k] = in.tauaer400[q] * powf(0.4f / wavemid, ang);
                float slope = (in.waer600[q] - in.waer400[q]) / .2f;
                W.ssaa[nb][k] = slope * (wavemid - .6f) + in.waer600[q];
                if (W.ssaa[nb][k] < 0.4f) W.ssaa[nb][k] = 0.4f;
                if (W.ssaa[nb][k] >= 1.0f) W.ssaa[nb][k] = 1.0f;
                slope = (in.gaer600[q] - in.gaer400[q]) / .2f;
                W.asma[nb][k] = slope * (wavemid - .6f) + in.gaer600[q];
                if (W.asma[nb][k] < 0.5f) W.asma[nb][k] = 0.5f;
                if (W.asma[nb][k] >= 1.0f) W.asma[nb][k] = 1.0f;
              }
            }
          }
          for (int nb = 1; nb <= NBSW; nb++) {
            float slope = 0.f;
            for (int k = 1; k <= nz; k++) slope = slope + W.taua[nb][k];
            if (slope < 0.f) { err = "ERROR: Negative total optical depth"; return ARC_ERR_NEG_AOD; }
            else if (slope > 6.f) { for (int k = 1; k <= nz; k++) W.taua[nb][k] = W.taua[nb][k] * 6.0f / slope; }
          }
        }
        if (in.aer_opt == 1) {
          // ecaer of the adapter (SW:11083-11100: model layers from AEROD, 0 in the extra top layer) and the iaer = 6 mixing of
          // rrtmg_sw (SW:9313-9341), which replaces taua / ssaa / asma
          const FArr &rsrtaua = T.in.get("sw_rsrtaua"), &rsrpiza = T.in.get("sw_rsrpiza"), &rsrasya = T.in.get("sw_rsrasya");
          const size_t n3 = (size_t)(d.ime - d.ims + 1) * (d.kme - d.kms + 1) * (d.jme - d.jms + 1);
          for (int k = 1; k <= nz + 1; k++) {
            float ecaer[7];
            for (int na = 1; na <= 6; na++) ecaer[na] = k <= nz ? in.aerod[ix.at3(i, kts + k - 1, j) + n3 * (size_t)(na - 1)] : 0.f;
            for (int ib = 1; ib <= NBSW; ib++) {
              float ztaua = 0.f, zasya = 0.f, zomga = 0.f;
              for (int ia = 1; ia <= 6; ia++) {
                ztaua = ztaua + rsrtaua(ib, ia) * ecaer[ia];
                zomga = zomga + rsrtaua(ib, ia) * ecaer[ia] * rsrpiza(ib, ia);
                zasya = zasya + rsrtaua(ib, ia) * ecaer[ia] * rsrpiza(ib, ia) * rsrasya(ib, ia);
              }
              if (zomga != 0.f) zasya = zasya / zomga;
              if (ztaua != 0.f) zomga = zomga / ztaua;
              W.taua[ib][k] = ztaua; W.ssaa[ib][k] = zomga; W.asma[ib][k] = zasya;
            }
          }
        }
        SwColumnOut co;
        g_min_cond = 1.f;
        int rc = rrtmg_sw_column(T, nlay, play, plev, tlay, h2ovmr, o3vmr, co2, ch4, n2o, o2, asdir, asdif, aldir, aldif, coszrs,
                                 1.0f, 0, in.solcon, inflgsw, iceflgsw, liqflgsw, clean, W, co, err);
        if (rc) return rc;
        out.gsw[ij] = co.swdflx[1] - co.swuflx[1];
        out.swcf[ij] = (co.swdflx[nz + 2] - co.swuflx[nz + 2]) - (co.swdflxc[nz + 2] - co.swuflxc[nz + 2]);
        if (out.swupt) {
          out.swupt[ij] = co.swuflx[nz + 2]; out.swuptc[ij] = co.swuflxc[nz + 2]; out.swuptcln[ij] = co.swuflxcln[nz + 2];
          out.swdnt[ij] = co.swdflx[nz + 2]; out.swdntc[ij] = co.swdflxc[nz + 2]; out.swdntcln[ij] = co.swdflxcln[nz + 2];
          out.swupb[ij] = co.swuflx[1]; out.swupbc[ij] = co.swuflxc[1]; out.swupbcln[ij] = co.swuflxcln[1];
          out.swdnb[ij] = co.swdflx[1]; out.swdnbc[ij] = co.swdflxc[1]; out.swdnbcln[ij] = co.swdflxcln[1];
          out.swvisdir[ij] = co.sibvisdir[1]; out.swvisdif[ij] = co.sibvisdif[1];
          out.swnirdir[ij] = co.sibnirdir[1]; out.swnirdif[ij] = co.sibnirdif[1];
        }
        if (out.swuptclnc) {
          out.swuptclnc[ij] = co.swuflxclnc[nz + 2]; out.swdntclnc[ij] = co.swdflxclnc[nz + 2];
          out.swupbclnc[ij] = co.swuflxclnc[1]; out.swdnbclnc[ij] = co.swdflxclnc[1];
        }
        out.swddir[ij] = co.swdkdir[1];
        out.swddni[ij] = out.swddir[ij] / coszrs;
        out.swddif[ij] = co.swdkdif[1];
        if (out.swupflx) {
          for (int k = 1; k <= nz + 2; k++) {
            size_t q = ix.atp(i, kts + k - 1, j);
            out.swupflx[q] = co.swuflx[k]; out.swupflxc[q] = co.swuflxc[k]; out.swupflxcln[q] = co.swuflxcln[k];
            out.swdnflx[q] = co.swdflx[k]; out.swdnflxc[q] = co.swdflxc[k]; out.swdnflxcln[q] = co.swdflxcln[k];
          }
        }
        for (int k = 1; k <= nz; k++) {
          float tten = co.swhr[k] / 86400.f;
          out.rthratensw[ix.at3(i, kts + k - 1, j)] = tten / in.pi3d[ix.at3(i, kts + k - 1, j)];
        }
        if (dbg) {
          if (dbg->laytrop) dbg->laytrop[c] = W.coef.laytrop;
          if (dbg->sw_cond) dbg->sw_cond[c] = g_min_cond;
          for (int l = 1; l <= nlay; l++) {
            size_t q = c * nlay + (l - 1);
            if (dbg->jp) dbg->jp[q] = W.coef.jp[l];
            if (dbg->jt) dbg->jt[q] = W.coef.jt[l];
            if (dbg->jt1) dbg->jt1[q] = W.coef.jt1[l];
            if (dbg->indfor) dbg->indfor[q] = W.coef.indfor[l];
            if (dbg->indself) dbg->indself[q] = W.coef.indself[l];
            if (dbg->fac00) dbg->fac00[q] = W.coef.fac00[l];
            if (dbg->fac01) dbg->fac01[q] = W.coef.fac01[l];
            if (dbg->fac10) dbg->fac10[q] = W.coef.fac10[l];
            if (dbg->fac11) dbg->fac11[q] = W.coef.fac11[l];
            if (dbg->hr) dbg->hr[q] = co.swhr[l];
            for (int ig = 1; ig <= NGSW; ig++) {
              size_t qq = q * NGSW + (ig - 1);
              if (dbg->cldmask) dbg->cldmask[qq] = W.cld.cldfmc[ig][l] != 0.f;
              if (dbg->taug) dbg->taug[qq] = W.tau.taug[ig][l];
              if (dbg->taur) dbg->taur[qq] = W.tau.taur[ig][l];
              if (dbg->taucmc) dbg->taucmc[qq] = W.cld.taucmc[ig][l];
            }
          }
          if (dbg->sfluxzen) for (int ig = 1; ig <= NGSW; ig++) dbg->sfluxzen[c * NGSW + ig - 1] = W.tau.sfluxzen[ig];
        }
      } else {
        if (out.swupt) {
          out.swupt[ij] = 0.f; out.swuptc[ij] = 0.f; out.swuptcln[ij] = 0.f; out.swdnt[ij] = 0.f; out.swdntc[ij] = 0.f; out.swdntcln[ij] = 0.f;
          out.swupb[ij] = 0.f; out.swupbc[ij] = 0.f; out.swupbcln[ij] = 0.f; out.swdnb[ij] = 0.f; out.swdnbc[ij] = 0.f; out.swdnbcln[ij] = 0.f;
          out.swvisdir[ij] = 0.f; out.swvisdif[ij] = 0.f; out.swnirdir[ij] = 0.f; out.swnirdif[ij] = 0.f;
        }
        if (out.swuptclnc) { out.swuptclnc[ij] = 0.f; out.swdntclnc[ij] = 0.f; out.swupbclnc[ij] = 0.f; out.swdnbclnc[ij] = 0.f; }
        out.swddir[ij] = 0.f; out.swddni[ij] = 0.f; out.swddif[ij] = 0.f; out.swcf[ij] = 0.f;
      }
    }
  }
  return 0;
}

}  // namespace orc
