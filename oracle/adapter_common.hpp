// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// Pieces of the WRF<->RRTMG adapters that the reference duplicates textually in RRTMG_SWRAD
// (module_ra_rrtmg_sw.F:10320-10906) and RRTMG_LWRAD (module_ra_rrtmg_lw.F:11877-12470),
// plus inirad/o3data (LW:12704-12840), relcalc (LW:14403-14462), reicalc (LW:14464-14491).
#pragma once
#include <algorithm>
#include <cmath>

#include "../include/arc_rad.h"
#include "oracle.hpp"

namespace orc {

struct Idx {
  int ims, ime, kms, kme, jms, jme;
  size_t ni, nk;
  explicit Idx(const ArcDims &d)
      : ims(d.ims), ime(d.ime), kms(d.kms), kme(d.kme), jms(d.jms), jme(d.jme),
        ni((size_t)(d.ime - d.ims + 1)), nk((size_t)(d.kme - d.kms + 1)) {}
  size_t at3(int i, int k, int j) const { return (size_t)(i - ims) + ni * ((size_t)(k - kms) + nk * (size_t)(j - jms)); }
  size_t at2(int i, int j) const { return (size_t)(i - ims) + ni * (size_t)(j - jms); }
  size_t atp(int i, int k, int j) const { return (size_t)(i - ims) + ni * ((size_t)(k - kms) + (nk + 2) * (size_t)(j - jms)); }
  size_t at4(int i, int k, int j, int n) const { return at3(i, k, j) + ni * nk * (size_t)(jme - jms + 1) * (size_t)n; }
};

// Inputs common to both adapters (a view on ArcSwIn / ArcLwIn)
struct CloudIn {
  int icloud, warm_rain, is_cammgmp_used, has_reqc, has_reqi, has_reqs, progn;
  int f_qv, f_qc, f_qr, f_qi, f_qs, f_qg, f_qndrop;
  float g;
  const float *t3d, *cldfra3d, *lradius, *iradius, *qv3d, *qc3d, *qr3d, *qi3d, *qs3d, *qg3d, *qndrop3d;
  const float *re_cloud, *re_ice, *re_snow, *f_ice_phy, *xland, *xice, *snow;
};

// 1-D column state after the gather + hydrometeor logic (k index 1..nz == kts..kte)
struct Col1D {
  float qv[MXLAY], qc[MXLAY], qr[MXLAY], qi[MXLAY], qs[MXLAY], qg[MXLAY], cldfra[MXLAY], qndrop[MXLAY];
  float recloud[MXLAY], reice1[MXLAY], resnow[MXLAY];
  int inflg, iceflg, liqflg;
};

// SW:10351-10500 == LW:11897-12044: gather water species with the F_Qx / warm_rain logic
inline void gather_hydrometeors(const CloudIn &c, const Idx &ix, int i, int j, int kts, int kte, const float *t1d /*1-based*/,
                                Col1D &o) {
  int nz = kte - kts + 1;
  for (int k = 1; k <= nz; k++) {
    o.qv[k] = 0.f; o.qc[k] = 0.f; o.qr[k] = 0.f; o.qi[k] = 0.f; o.qs[k] = 0.f; o.qg[k] = 0.f; o.cldfra[k] = 0.f; o.qndrop[k] = 0.f;
  }
  for (int k = 1; k <= nz; k++) { o.qv[k] = c.qv3d[ix.at3(i, kts + k - 1, j)]; o.qv[k] = std::max(0.f, o.qv[k]); }
  if (c.icloud != 0) {
    if (c.cldfra3d) for (int k = 1; k <= nz; k++) o.cldfra[k] = c.cldfra3d[ix.at3(i, kts + k - 1, j)];
    if (c.f_qc >= 0 && c.qc3d && c.f_qc) for (int k = 1; k <= nz; k++) o.qc[k] = std::max(0.f, c.qc3d[ix.at3(i, kts + k - 1, j)]);
    if (c.f_qr >= 0 && c.qr3d && c.f_qr) for (int k = 1; k <= nz; k++) o.qr[k] = std::max(0.f, c.qr3d[ix.at3(i, kts + k - 1, j)]);
    if (c.f_qndrop >= 0 && c.qndrop3d && c.f_qndrop) for (int k = 1; k <= nz; k++) o.qndrop[k] = c.qndrop3d[ix.at3(i, kts + k - 1, j)];
    bool predicate = (c.f_qi >= 0) ? (c.f_qi != 0) : false;
    if (!predicate && !c.warm_rain) {
      for (int k = 1; k <= nz; k++) {
        if (t1d[k] < 273.15f) { o.qi[k] = o.qc[k]; o.qs[k] = o.qr[k]; o.qc[k] = 0.f; o.qr[k] = 0.f; }
      }
    }
    if (c.f_qi >= 0 && c.qi3d && c.f_qi) for (int k = 1; k <= nz; k++) o.qi[k] = std::max(0.f, c.qi3d[ix.at3(i, kts + k - 1, j)]);
    if (c.f_qs >= 0 && c.qs3d && c.f_qs) for (int k = 1; k <= nz; k++) o.qs[k] = std::max(0.f, c.qs3d[ix.at3(i, kts + k - 1, j)]);
    if (c.f_qg >= 0 && c.qg3d && c.f_qg) for (int k = 1; k <= nz; k++) o.qg[k] = std::max(0.f, c.qg3d[ix.at3(i, kts + k - 1, j)]);
    if (c.f_qi >= 0 && c.f_qc >= 0 && c.f_qs >= 0 && c.f_ice_phy) {
      if (c.f_qc && !c.f_qi && c.f_qs) {
        for (int k = 1; k <= nz; k++) {
          float qs3 = c.qs3d[ix.at3(i, kts + k - 1, j)];
          o.qi[k] = 0.1f * qs3; o.qs[k] = 0.9f * qs3; o.qc[k] = c.qc3d[ix.at3(i, kts + k - 1, j)];
          o.qi[k] = std::max(0.f, o.qi[k]); o.qc[k] = std::max(0.f, o.qc[k]);
        }
      }
    }
  }
  for (int k = 1; k <= nz; k++) o.qv[k] = std::max(o.qv[k], 1.e-12f);
}

// SW:10514-10608 == LW:12044-12145 (EM_CORE==1 branches): flags + re_* handling
inline void effective_radius_inputs(const CloudIn &c, const Idx &ix, int i, int j, int kts, int kte, const FArr &retab,
                                    Col1D &o) {
  int nz = kte - kts + 1;
  o.inflg = 2; o.iceflg = 3; o.liqflg = 1;
  if (c.icloud != 0) {
    float xl = c.xland[ix.at2(i, j)];
    if (c.has_reqc != 0) {
      o.inflg = 3;
      for (int k = 1; k <= nz; k++) {
        size_t q = ix.at3(i, kts + k - 1, j);
        o.recloud[k] = std::max(2.5f, c.re_cloud[q] * 1.e6f);
        if (o.recloud[k] <= 2.5f && c.cldfra3d[q] > 0.f && (xl - 1.5f) > 0.f) o.recloud[k] = 10.5f;
        else if (o.recloud[k] <= 2.5f && c.cldfra3d[q] > 0.f && (xl - 1.5f) < 0.f) o.recloud[k] = 7.5f;
      }
    } else {
      for (int k = 1; k <= nz; k++) o.recloud[k] = 5.0f;
    }
    if (c.has_reqi != 0) {
      o.inflg = 4; o.iceflg = 4;
      for (int k = 1; k <= nz; k++) {
        size_t q = ix.at3(i, kts + k - 1, j);
        o.reice1[k] = std::max(5.f, c.re_ice[q] * 1.e6f);
        if (o.reice1[k] <= 5.f && c.cldfra3d[q] > 0.f) {
          float t = c.t3d[q];
          int idx_rei = (int)(t - 179.f);
          idx_rei = std::min(std::max(idx_rei, 1), 75);
          float corr = t - (float)(int)t;
          o.reice1[k] = retab(idx_rei) * (1.f - corr) + retab(idx_rei + 1) * corr;
          o.reice1[k] = std::max(o.reice1[k], 5.0f);
        }
      }
    } else {
      for (int k = 1; k <= nz; k++) o.reice1[k] = 10.f;
    }
    if (c.has_reqs != 0) {
      o.inflg = 5; o.iceflg = 5;
      for (int k = 1; k <= nz; k++) o.resnow[k] = std::max(10.f, c.re_snow[ix.at3(i, kts + k - 1, j)] * 1.e6f);
    } else {
      for (int k = 1; k <= nz; k++) o.resnow[k] = 10.0f;
    }
    if (c.has_reqs == 0 && c.has_reqi != 0 && c.has_reqc != 0) {
      o.inflg = 5; o.iceflg = 5;
      for (int k = 1; k <= nz; k++) {
        size_t q = ix.at3(i, kts + k - 1, j);
        o.resnow[k] = std::max(10.f, c.re_ice[q] * 1.e6f);
        o.qs[k] = c.qi3d[q];
        o.qi[k] = 0.f;
        o.reice1[k] = 10.f;
      }
    }
  }
}

// LW:14403-14462
inline void relcalc(int pver, const float *t /*1-based*/, float landfrac, float landm, float icefrac, float snowh, float *rel) {
  const float tmelt = 273.16f, rliqocean = 14.0f, rliqice = 14.0f, rliqland = 8.0f;
  (void)landfrac;
  for (int k = 1; k <= pver; k++) {
    rel[k] = rliqland + (rliqocean - rliqland) * std::min(1.0f, std::max(0.0f, (tmelt - t[k]) * 0.05f));
    rel[k] = rel[k] + (rliqocean - rel[k]) * std::min(1.0f, std::max(0.0f, snowh * 10.f));
    rel[k] = rel[k] + (rliqocean - rel[k]) * std::min(1.0f, std::max(0.0f, 1.0f - landm));
    rel[k] = rel[k] + (rliqice - rel[k]) * std::min(1.0f, std::max(0.0f, icefrac));
  }
}

// LW:14464-14491
inline void reicalc(int pver, const float *t, const FArr &retab, float *re) {
  for (int k = 1; k <= pver; k++) {
    int index = (int)(t[k] - 179.f);
    index = std::min(std::max(index, 1), 94);
    float corr = t[k] - (float)(int)t[k];
    re[k] = retab(index) * (1.f - corr) + retab(index + 1) * corr;
  }
}

// SW:10758-10906 == LW:12331-12452: in-cloud water paths and effective radii for k = 1..nz
struct CloudPaths { float clwp[MXLAY], ciwp[MXLAY], cswp[MXLAY], rel[MXLAY], rei[MXLAY], res[MXLAY], cldfrac[MXLAY]; };

inline void cloud_paths(const CloudIn &c, const Idx &ix, int i, int j, int kts, int kte, Col1D &o, const float *pdel,
                        const float *tlay, const FArr &retab, CloudPaths &cp) {
  int nz = kte - kts + 1;
  float cicewp[MXLAY], cliqwp[MXLAY], csnowp[MXLAY], reliq[MXLAY], reice[MXLAY];
  for (int k = 1; k <= nz; k++) cp.cldfrac[k] = o.cldfra[k];
  float gravmks = c.g;
  float landfrac = 2.f - c.xland[ix.at2(i, j)];
  float landm = landfrac;
  float snowh = 0.001f * c.snow[ix.at2(i, j)];
  float icefrac = c.xice[ix.at2(i, j)];
  for (int k = 1; k <= nz; k++) {
    float gicewp = (o.qi[k] + o.qs[k]) * pdel[k] * 100.0f / gravmks * 1000.0f;
    float gliqwp = o.qc[k] * pdel[k] * 100.0f / gravmks * 1000.0f;
    cicewp[k] = gicewp / std::max(0.01f, cp.cldfrac[k]);
    cliqwp[k] = gliqwp / std::max(0.01f, cp.cldfrac[k]);
  }
  if (o.iceflg >= 4) {
    for (int k = 1; k <= nz; k++) {
      float gicewp = o.qi[k] * pdel[k] * 100.0f / gravmks * 1000.0f;
      cicewp[k] = gicewp / std::max(0.01f, cp.cldfrac[k]);
    }
  }
  if (o.iceflg == 5) {
    for (int k = 1; k <= nz; k++) {
      float snow_mass_factor = 1.0f;
      if (o.resnow[k] > 130.f) {
        snow_mass_factor = (130.0f / o.resnow[k]) * (130.0f / o.resnow[k]);
        o.resnow[k] = 130.0f;
      }
      float gsnowp = o.qs[k] * snow_mass_factor * pdel[k] * 100.0f / gravmks * 1000.0f;
      csnowp[k] = gsnowp / std::max(0.01f, cp.cldfrac[k]);
    }
  }
  if (c.progn == 1) {
    float pi = 4.f * atanf(1.0f);
    float third = 1.f / 3.f;
    float rhoh2o = 1.e3f;
    float relconst = 3 / (4.f * pi * rhoh2o);
    float lwpmin = 3.e-5f;
    for (int k = 1; k <= nz; k++) {
      reliq[k] = 10.f;
      if (c.f_qndrop >= 0 && c.f_qndrop) {
        if (o.qc[k] * pdel[k] > lwpmin && o.qndrop[k] > 1000.f) {
          reliq[k] = powf(relconst * o.qc[k] / o.qndrop[k], third);
          reliq[k] = 1.1f * reliq[k];
          reliq[k] = reliq[k] * 1.e6f;
          reliq[k] = std::max(reliq[k], 4.f);
          reliq[k] = std::min(reliq[k], 20.f);
        }
      }
    }
  } else {
    relcalc(nz, tlay, landfrac, landm, icefrac, snowh, reliq);
  }
  reicalc(nz, tlay, retab, reice);
  if (o.inflg >= 3) for (int k = 1; k <= nz; k++) reliq[k] = o.recloud[k];
  if (o.iceflg >= 4) for (int k = 1; k <= nz; k++) reice[k] = o.reice1[k];
  if (o.iceflg == 3) {
    for (int k = 1; k <= nz; k++) { reice[k] = reice[k] * 1.0315f; reice[k] = std::min(140.0f, reice[k]); }
  }
  if (c.is_cammgmp_used) {
    for (int k = 1; k <= nz; k++) {
      size_t q = ix.at3(i, kts + k - 1, j);
      if (o.qi[k] > 1.e-20f || o.qs[k] > 1.e-20f) reice[k] = c.iradius[q]; else reice[k] = 25.f;
      reice[k] = std::max(5.f, std::min(140.0f, reice[k]));
      if (o.qc[k] > 1.e-20f) reliq[k] = c.lradius[q]; else reliq[k] = 10.f;
      reliq[k] = std::max(2.5f, std::min(60.0f, reliq[k]));
    }
  }
  for (int k = 1; k <= nz; k++) { cp.clwp[k] = cliqwp[k]; cp.ciwp[k] = cicewp[k]; cp.rel[k] = reliq[k]; cp.rei[k] = reice[k]; }
  if (o.inflg == 5) {
    for (int k = 1; k <= nz; k++) { cp.cswp[k] = csnowp[k]; cp.res[k] = o.resnow[k]; }
  } else {
    for (int k = 1; k <= nz; k++) { cp.cswp[k] = 0.f; cp.res[k] = 10.f; }
  }
}

// inirad + o3data, LW:12704-12840.  plev index 1..nl+1 (1 = surface, hPa); o3prof 1..nl
inline void o3data(const InlineTables &in, const float *plev, int nl, float *o3prof) {
  const FArr &o3sum = in.get("lw_o3sum"), &ppsum = in.get("lw_ppsum"), &o3win = in.get("lw_o3win"), &ppwin = in.get("lw_ppwin");
  float o3ann[32], ppann[32], o3wrk[32], ppwrk[32], ppwrkh[33];
  for (int k = 1; k <= 31; k++) ppann[k] = ppsum(k);
  o3ann[1] = 0.5f * (o3sum(1) + o3win(1));
  for (int k = 2; k <= 31; k++)
    o3ann[k] = o3win(k - 1) + (o3win(k) - o3win(k - 1)) / (ppwin(k) - ppwin(k - 1)) * (ppsum(k) - ppwin(k - 1));
  for (int k = 2; k <= 31; k++) o3ann[k] = 0.5f * (o3ann[k] + o3sum(k));
  for (int k = 1; k <= 31; k++) { o3wrk[k] = o3ann[k]; ppwrk[k] = ppann[k]; }
  ppwrkh[1] = 1100.f;
  for (int k = 2; k <= 31; k++) ppwrkh[k] = (ppwrk[k] + ppwrk[k - 1]) / 2.f;
  ppwrkh[32] = 0.f;
  for (int k = 1; k <= nl; k++) {
    o3prof[k] = 0.f;
    for (int jj = 1; jj <= 31; jj++) {
      float pb1, pb2, pt1, pt2;
      if ((-(plev[k] - ppwrkh[jj])) >= 0.f) pb1 = 0.f; else pb1 = plev[k] - ppwrkh[jj];
      if ((-(plev[k] - ppwrkh[jj + 1])) >= 0.f) pb2 = 0.f; else pb2 = plev[k] - ppwrkh[jj + 1];
      if ((-(plev[k + 1] - ppwrkh[jj])) >= 0.f) pt1 = 0.f; else pt1 = plev[k + 1] - ppwrkh[jj];
      if ((-(plev[k + 1] - ppwrkh[jj + 1])) >= 0.f) pt2 = 0.f; else pt2 = plev[k + 1] - ppwrkh[jj + 1];
      o3prof[k] = o3prof[k] + (pb2 - pb1 - pt2 + pt1) * o3wrk[jj];
    }
    o3prof[k] = o3prof[k] / (plev[k] - plev[k + 1]);
  }
}

// kissvec SW:1900-1932 == LW:2586-2618 (ncol = 1)
struct Kiss {
  int32_t s1, s2, s3, s4;
  static inline int32_t ishft(int32_t k, int n) {
    uint32_t u = (uint32_t)k;
    return n >= 0 ? (int32_t)(u << n) : (int32_t)(u >> (-n));
  }
  static inline int32_t m(int32_t k, int n) { return k ^ ishft(k, n); }
  inline float next() {
    s1 = (int32_t)(69069u * (uint32_t)s1 + 1327217885u);
    s2 = m(m(m(s2, 13), -17), 5);
    s3 = (int32_t)(18000u * (uint32_t)(s3 & 65535) + (uint32_t)ishft(s3, -16));
    s4 = (int32_t)(30903u * (uint32_t)(s4 & 65535) + (uint32_t)ishft(s4, -16));
    int32_t kiss = (int32_t)((uint32_t)s1 + (uint32_t)s2 + (uint32_t)ishft(s3, 16) + (uint32_t)s4);
    volatile float prod = (float)kiss * 2.328306e-10f;
    return prod + 0.5f;
  }
};

// generate_stochastic_clouds(_sw) SW:1517-1896 == LW:2208-2578 for icld=2, irng=0, ncol=1.
// pmid in Pa (1-based, 1 = lowest layer); cdf[isub][lev]; returns mask in iscloudy
inline void mcica_mask(int nlay, int nsubcol, const float *pmid, const float *cld, int changeSeed,
                       std::vector<float> &cdf, std::vector<unsigned char> &iscloudy) {
  const float cldmin = 1.0e-20f;
  float cldf[MXLAY];
  for (int l = 1; l <= nlay; l++) { cldf[l] = cld[l]; if (cldf[l] < cldmin) cldf[l] = 0.f; }
  Kiss K;
  K.s1 = (int32_t)((pmid[1] - (float)(int)pmid[1]) * 1000000000.f);
  K.s2 = (int32_t)((pmid[2] - (float)(int)pmid[2]) * 1000000000.f);
  K.s3 = (int32_t)((pmid[3] - (float)(int)pmid[3]) * 1000000000.f);
  K.s4 = (int32_t)((pmid[4] - (float)(int)pmid[4]) * 1000000000.f);
  for (int i = 1; i <= changeSeed; i++) (void)K.next();
  cdf.assign((size_t)nsubcol * (nlay + 1), 0.f);
  auto CDF = [&](int isub, int l) -> float & { return cdf[(size_t)(isub - 1) * (nlay + 1) + l]; };
  for (int isub = 1; isub <= nsubcol; isub++)
    for (int l = 1; l <= nlay; l++) CDF(isub, l) = K.next();
  for (int l = 2; l <= nlay; l++) {
    for (int isub = 1; isub <= nsubcol; isub++) {
      if (CDF(isub, l - 1) > 1.0f - cldf[l - 1]) CDF(isub, l) = CDF(isub, l - 1);
      else CDF(isub, l) = CDF(isub, l) * (1.0f - cldf[l - 1]);
    }
  }
  iscloudy.assign((size_t)nsubcol * (nlay + 1), 0);
  for (int l = 1; l <= nlay; l++)
    for (int isub = 1; isub <= nsubcol; isub++)
      iscloudy[(size_t)(isub - 1) * (nlay + 1) + l] = (CDF(isub, l) >= 1.0f - cldf[l]) ? 1 : 0;
}

}  // namespace orc
