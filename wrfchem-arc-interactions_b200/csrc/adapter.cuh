// Per-layer pieces of the WRF<->RRTMG adapters and of setcoef, written for the lane-per-layer
// mapping: every function works on ONE (column, layer) pair.
//
// Reference: RRTMG_SWRAD module_ra_rrtmg_sw.F:10320-10906, RRTMG_LWRAD module_ra_rrtmg_lw.F:11877-12470,
// inirad/o3data LW:12704-12840, relcalc LW:14403-14462, reicalc LW:14464-14491,
// setcoef_sw SW:2734-2990, setcoef LW:3444-3809.
#pragma once
#include "args.h"
#include "glibc_math.cuh"

namespace arc {

struct LayerCloud {
  float clwp, ciwp, cswp, rel, rei, res, cldfrac;
};

// flags that depend only on the call (SW:10514-10608, LW:12044-12145; EM_CORE=1)
__host__ __device__ inline void cloud_flags(const CloudFields &cf, int &inflg, int &iceflg, int &liqflg) {
  inflg = 2; iceflg = 3; liqflg = 1;
  if (cf.icloud != 0) {
    if (cf.has_reqc != 0) inflg = 3;
    if (cf.has_reqi != 0) { inflg = 4; iceflg = 4; }
    if (cf.has_reqs != 0) { inflg = 5; iceflg = 5; }
    if (cf.has_reqs == 0 && cf.has_reqi != 0 && cf.has_reqc != 0) { inflg = 5; iceflg = 5; }
  }
}

// One model layer k (kts..kte): hydrometeor gather, effective radii, in-cloud water paths.
// t1d = t3d(i,k,j); tlay = layer temperature as seen by relcalc/reicalc (LW modifies it at kte);
// pdel = layer pressure thickness in hPa.
__device__ inline void layer_cloud(const CloudFields &cf, const Geo &G, const DevTables &tb, int i, int j, int k,
                                   float t1d, float tlay, float pdel, int inflg, int iceflg, LayerCloud &o) {
  const size_t q = G.at3(i, k, j);
  float qc = 0.f, qr = 0.f, qi = 0.f, qs = 0.f, cld = 0.f, qnd = 0.f;
  if (cf.icloud != 0) {
    if (cf.cldfra3d) cld = cf.cldfra3d[q];
    if (cf.f_qc > 0 && cf.qc3d) qc = fmaxf(0.f, cf.qc3d[q]);
    if (cf.f_qr > 0 && cf.qr3d) qr = fmaxf(0.f, cf.qr3d[q]);
    if (cf.f_qndrop > 0 && cf.qndrop3d) qnd = cf.qndrop3d[q];
    const bool predicate = cf.f_qi > 0;
    if (!predicate && !cf.warm_rain) {
      if (t1d < 273.15f) { qi = qc; qs = qr; qc = 0.f; qr = 0.f; }
    }
    if (cf.f_qi > 0 && cf.qi3d) qi = fmaxf(0.f, cf.qi3d[q]);
    if (cf.f_qs > 0 && cf.qs3d) qs = fmaxf(0.f, cf.qs3d[q]);
    if (cf.f_qi >= 0 && cf.f_qc >= 0 && cf.f_qs >= 0 && cf.f_ice_phy) {
      if (cf.f_qc > 0 && cf.f_qi == 0 && cf.f_qs > 0) {
        const float qs3 = cf.qs3d[q];
        qi = fmaxf(0.f, 0.1f * qs3); qs = 0.9f * qs3; qc = fmaxf(0.f, cf.qc3d[q]);
      }
    }
  }
  // re_* handling
  float recloud = 5.0f, reice1 = 10.f, resnow = 10.0f;
  if (cf.icloud != 0) {
    const float xl = cf.xland[G.at2(i, j)];
    if (cf.has_reqc != 0) {
      recloud = fmaxf(2.5f, cf.re_cloud[q] * 1.e6f);
      const float cf3 = cf.cldfra3d[q];
      if (recloud <= 2.5f && cf3 > 0.f && (xl - 1.5f) > 0.f) recloud = 10.5f;
      else if (recloud <= 2.5f && cf3 > 0.f && (xl - 1.5f) < 0.f) recloud = 7.5f;
    }
    if (cf.has_reqi != 0) {
      reice1 = fmaxf(5.f, cf.re_ice[q] * 1.e6f);
      if (reice1 <= 5.f && cf.cldfra3d[q] > 0.f) {
        int idx = (int)(t1d - 179.f);
        idx = min(max(idx, 1), 75);
        const float corr = t1d - (float)(int)t1d;
        reice1 = tb.retab[idx - 1] * (1.f - corr) + tb.retab[idx] * corr;
        reice1 = fmaxf(reice1, 5.0f);
      }
    }
    if (cf.has_reqs != 0) resnow = fmaxf(10.f, cf.re_snow[q] * 1.e6f);
    if (cf.has_reqs == 0 && cf.has_reqi != 0 && cf.has_reqc != 0) {
      resnow = fmaxf(10.f, cf.re_ice[q] * 1.e6f);
      qs = cf.qi3d[q];
      qi = 0.f;
      reice1 = 10.f;
    }
  }
  // water paths (SW:10758-10794)
  o.cldfrac = cld;
  const float gravmks = cf.g;
  const float cfd = fmaxf(0.01f, cld);
  float cicewp = ((qi + qs) * pdel * 100.0f / gravmks * 1000.0f) / cfd;
  const float cliqwp = (qc * pdel * 100.0f / gravmks * 1000.0f) / cfd;
  if (iceflg >= 4) cicewp = (qi * pdel * 100.0f / gravmks * 1000.0f) / cfd;
  float csnowp = 0.f;
  if (iceflg == 5) {
    float smf = 1.0f;
    if (resnow > 130.f) { smf = (130.0f / resnow) * (130.0f / resnow); resnow = 130.0f; }
    csnowp = (qs * smf * pdel * 100.0f / gravmks * 1000.0f) / cfd;
  }
  // effective radii
  float reliq;
  const size_t ij = G.at2(i, j);
  if (cf.progn == 1) {
    const float pi = 4.f * atanf(1.0f);
    const float relconst = 3 / (4.f * pi * 1.e3f);
    reliq = 10.f;
    if (cf.f_qndrop > 0) {
      if (qc * pdel > 3.e-5f && qnd > 1000.f) {
        reliq = glm::powf_(relconst * qc / qnd, 1.f / 3.f);
        reliq = 1.1f * reliq;
        reliq = reliq * 1.e6f;
        reliq = fminf(fmaxf(reliq, 4.f), 20.f);
      }
    }
  } else {  // relcalc
    const float landm = 2.f - cf.xland[ij], snowh = 0.001f * cf.snow[ij], icefrac = cf.xice[ij];
    reliq = 8.0f + (14.0f - 8.0f) * fminf(1.0f, fmaxf(0.0f, (273.16f - tlay) * 0.05f));
    reliq = reliq + (14.0f - reliq) * fminf(1.0f, fmaxf(0.0f, snowh * 10.f));
    reliq = reliq + (14.0f - reliq) * fminf(1.0f, fmaxf(0.0f, 1.0f - landm));
    reliq = reliq + (14.0f - reliq) * fminf(1.0f, fmaxf(0.0f, icefrac));
  }
  float reice;
  {  // reicalc
    int index = (int)(tlay - 179.f);
    index = min(max(index, 1), 94);
    const float corr = tlay - (float)(int)tlay;
    reice = tb.retab[index - 1] * (1.f - corr) + tb.retab[index] * corr;
  }
  if (inflg >= 3) reliq = recloud;
  if (iceflg >= 4) reice = reice1;
  if (iceflg == 3) { reice = reice * 1.0315f; reice = fminf(140.0f, reice); }
  if (cf.is_cammgmp_used) {
    reice = (qi > 1.e-20f || qs > 1.e-20f) ? cf.iradius[q] : 25.f;
    reice = fmaxf(5.f, fminf(140.0f, reice));
    reliq = (qc > 1.e-20f) ? cf.lradius[q] : 10.f;
    reliq = fmaxf(2.5f, fminf(60.0f, reliq));
  }
  o.clwp = cliqwp; o.ciwp = cicewp; o.rel = reliq; o.rei = reice;
  if (inflg == 5) { o.cswp = csnowp; o.res = resnow; } else { o.cswp = 0.f; o.res = 10.f; }
}

// water vapour of a model layer after the adapter's floors (SW:10364-10367, 10508-10510)
__device__ inline float layer_qv(const CloudFields &cf, size_t q) { return fmaxf(fmaxf(0.f, cf.qv3d[q]), 1.e-12f); }

// ozone climatology mass mixing ratio of one layer between interface pressures pb (bottom) > pt (top), hPa
__device__ inline float o3_clim(const DevTables &tb, float pb, float pt) {
  float acc = 0.f;
  for (int jj = 0; jj < 31; jj++) {
    const float h0 = tb.ppwrkh[jj], h1 = tb.ppwrkh[jj + 1];
    const float pb1 = ((-(pb - h0)) >= 0.f) ? 0.f : pb - h0;
    const float pb2 = ((-(pb - h1)) >= 0.f) ? 0.f : pb - h1;
    const float pt1 = ((-(pt - h0)) >= 0.f) ? 0.f : pt - h0;
    const float pt2 = ((-(pt - h1)) >= 0.f) ? 0.f : pt - h1;
    acc = acc + (pb2 - pb1 - pt2 + pt1) * tb.o3wrk[jj];
  }
  return acc / (pb - pt);
}

// Pressure / temperature interpolation indices and weights shared by setcoef_sw and setcoef
// (SW:2851-2887, 2979-2983; LW:3649-3685, 3800-3804).
struct PTCoef { int jp, jt, jt1; float fac00, fac01, fac10, fac11, plog; };

__device__ inline void pt_coef(const float *__restrict__ preflog, const float *__restrict__ tref, float pavel, float tavel, PTCoef &c) {
  // glibc's logf bit for bit (glibc_math.cuh): plog defines jp and, through fp, every interpolation weight
  const float plog = glm::logf_(pavel);
  int jp = (int)(36.f - 5 * (plog + 0.04f));
  jp = min(max(jp, 1), 58);
  const int jp1 = jp + 1;
  const float fp = 5.f * (preflog[jp - 1] - plog);
  const float d0 = (tavel - tref[jp - 1]) / 15.f;
  int jt = (int)(3.f + d0);
  jt = min(max(jt, 1), 4);
  const float ft = d0 - (float)(jt - 3);
  const float d1 = (tavel - tref[jp1 - 1]) / 15.f;
  int jt1 = (int)(3.f + d1);
  jt1 = min(max(jt1, 1), 4);
  const float ft1 = d1 - (float)(jt1 - 3);
  const float compfp = 1.f - fp;
  c.jp = jp; c.jt = jt; c.jt1 = jt1; c.plog = plog;
  c.fac10 = compfp * ft;
  c.fac00 = compfp * (1.f - ft);
  c.fac11 = fp * ft1;
  c.fac01 = fp * (1.f - ft1);
}

// dry-air column (molecules/cm2), inatm_sw SW:9790-9799 == inatm LW:11306-11311
__device__ inline float coldry_of(float pbot, float ptop, float h2ovmr) {
  const float amd = 28.9660f, amw = 18.0160f, avogad = 6.02214199e+23f, grav = 9.8066f;
  const float amm = (1.f - h2ovmr) * amd + h2ovmr * amw;
  return (pbot - ptop) * 1.e3f * avogad / (1.e2f * grav * amm * (1.f + h2ovmr));
}

// kissvec (SW:1900-1932 == LW:2586-2618): one draw of the KISS generator, wrap-around int32
struct Kiss {
  uint32_t s1, s2, s3, s4;
  __device__ inline float next() {
    s1 = 69069u * s1 + 1327217885u;
    s2 ^= s2 << 13; s2 ^= s2 >> 17; s2 ^= s2 << 5;
    s3 = 18000u * (s3 & 65535u) + (s3 >> 16);
    s4 = 30903u * (s4 & 65535u) + (s4 >> 16);
    const int32_t kiss = (int32_t)(s1 + s2 + (s3 << 16) + s4);
    return (float)kiss * 2.328306e-10f + 0.5f;     // compiled with -fmad=false: mul then add
  }
};

}  // namespace arc
