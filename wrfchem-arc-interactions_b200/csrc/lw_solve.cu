// Longwave spectral solver: k_lw_solve = taumol (taugb1..16) + downward sweep of rtrnmc for the full and the clean
// (aerosol-free) call, k_lw_sweep = upward sweep + ordered sum over the g-points of a band, k_lw_reduce = band sum,
// heating rates and scatter.
//
// Reference (module_ra_rrtmg_lw.F v3.9.1): taumol 4712-7828, rtrnmc 2974-3410, rrtmg_lw 10984-11044
// (taut = taug + taua; clean rtrnmc(taug) then rtrnmc(taut)), RRTMG_LWRAD output scatter 12646-12692.
#include "args.h"
#include "../../include/arc_rad.h"

namespace arc {

static __constant__ LwBandDesc c_lw[16];
static __constant__ int c_lw_ngb[NGLW];    // band index 0..15 of each LW g-point
#ifndef LW_GMAX
#define LW_GMAX 16
#endif
static SweepGroups h_lw_grp;               // sweep groups, see sw_solve.cu
static __constant__ int c_lw_grp_band[SWEEP_MAXGRP];
void upload_band_descs_lw(const HostTables &T) {
  cudaMemcpyToSymbol(c_lw, T.lw, sizeof(LwBandDesc) * 16);
  int ngs[16], g0s[16];
  for (int b = 0; b < 16; b++) { ngs[b] = T.lw[b].ng; g0s[b] = T.lw[b].g0; }
  h_lw_grp = make_sweep_groups(ngs, g0s, 16, LW_GMAX);
  cudaMemcpyToSymbol(c_lw_grp_band, h_lw_grp.band, sizeof(int) * SWEEP_MAXGRP);
  int ngb[NGLW];
  for (int i = 0; i < NGLW; i++) ngb[i] = T.lw_ngb[i] - 1;
  cudaMemcpyToSymbol(c_lw_ngb, ngb, sizeof(int) * NGLW);
}

struct LwEta { float speccomb, specparm, f; int j; };

__device__ __forceinline__ LwEta lw_eta(float cola, float ratio, float colb, float mult, float oneminus) {
  LwEta e;
  e.speccomb = mul_add_rn(ratio, colb, cola);
  e.specparm = div_rn(cola, e.speccomb);
  if (e.specparm >= oneminus) e.specparm = oneminus;
  const float specmult = __fmul_rn(mult, e.specparm);
  e.j = 1 + (int)specmult;
  e.f = fmod1(specmult);
  return e;
}
__device__ __forceinline__ float pow4f(float p) { const float p2 = p * p; return p2 * p2; }

// lower-atmosphere major-species term of the binary bands for one pressure level (LW:5219-5349)
__device__ __forceinline__ float lw_major_lower(const float *__restrict__ A, const LwEta &e, int ind, float fa, float fb) {
  const float *p = A + ind - 1;
  if (e.specparm < 0.125f) {
    const float q = e.f - 1;
    const float p4 = pow4f(q);
    const float fk0 = p4, fk1 = 1 - q - 2.0f * p4, fk2 = q + p4;
    return e.speccomb * ((fk0 * fa) * p[0] + (fk1 * fa) * p[1] + (fk2 * fa) * p[2] + (fk0 * fb) * p[9] + (fk1 * fb) * p[10] + (fk2 * fb) * p[11]);
  } else if (e.specparm > 0.875f) {
    const float q = -e.f;
    const float p4 = pow4f(q);
    const float fk0 = p4, fk1 = 1 - q - 2.0f * p4, fk2 = q + p4;
    return e.speccomb * ((fk2 * fa) * p[-1] + (fk1 * fa) * p[0] + (fk0 * fa) * p[1] + (fk2 * fb) * p[8] + (fk1 * fb) * p[9] + (fk0 * fb) * p[10]);
  } else {
    const float f1 = 1.f - e.f;
    return e.speccomb * ((f1 * fa) * p[0] + (e.f * fa) * p[1] + (f1 * fb) * p[9] + (e.f * fb) * p[10]);
  }
}

#ifndef LW_BLOCK_SZ
#define LW_BLOCK_SZ 512
#endif
constexpr int LW_BLOCK = LW_BLOCK_SZ;   // 91 KB of staged tables per block: two blocks per SM
struct LwSmem { const float2 *et; const float *S, *plk, *rat, *chi; };

// One (column, g-point) of band BAND: taumol + both rtrnmc calls.  BAND is a template parameter so that each band's
// instantiation only contains (and requests up front) the workspace loads and the table arithmetic of that band.
template <int NL, int BAND>
__device__ __forceinline__ void lw_solve_band(const LwArgs &a, const LwSmem &sm, const LwBandDesc &D, int g, int c) {
  const float2 *s_et = sm.et;
  const float *S = sm.S, *s_plk = sm.plk, *s_rat = sm.rat, *s_chi = sm.chi;
  const DevTables &tb = a.tb;
  constexpr int band = BAND;
  constexpr int b = BAND - 1;
  const LwWs &ws = a.ws;
  const int nlay = ws.nlay;
  const size_t cap = ws.cap;
  const float bpade = tb.bpade, oneminus = tb.oneminus;
  const bool do_clean = (a.variants & ARC_VAR_CLEAN) != 0;
  const int laytrop = ws.laytrop[c];
  const float secdiff = ws.secdiff[(size_t)b * cap + c];
  auto CHI = [&](int imol, int jp) { return s_chi[(imol - 1) + 7 * (jp - 1)]; };   // 1-based like the reference

  // band constants (reference ratios at fixed pressure levels)
  float rp_a = 0.f, rp_b = 0.f, rm_a = 0.f, rm_b = 0.f, rm_a3 = 0.f;
  switch (band) {
    case 3: rp_a = div_rn(CHI(1, 9), CHI(2, 9)); rp_b = div_rn(CHI(1, 13), CHI(2, 13));
            rm_a = div_rn(CHI(1, 3), CHI(2, 3)); rm_b = div_rn(CHI(1, 13), CHI(2, 13)); break;
    case 4: rp_a = div_rn(CHI(1, 11), CHI(2, 11)); rp_b = div_rn(CHI(3, 13), CHI(2, 13)); break;
    case 5: rp_a = div_rn(CHI(1, 5), CHI(2, 5)); rp_b = div_rn(CHI(3, 43), CHI(2, 43)); rm_a = div_rn(CHI(1, 7), CHI(2, 7)); break;
    case 7: rp_a = div_rn(CHI(1, 3), CHI(3, 3)); rm_a = div_rn(CHI(1, 3), CHI(3, 3)); break;
    case 9: rp_a = div_rn(CHI(1, 9), CHI(6, 9)); rm_a = div_rn(CHI(1, 3), CHI(6, 3)); break;
    case 12: rp_a = div_rn(CHI(1, 10), CHI(2, 10)); break;
    case 13: rp_a = div_rn(CHI(1, 5), CHI(4, 5)); rm_a = div_rn(CHI(1, 1), CHI(4, 1)); rm_a3 = div_rn(CHI(1, 3), CHI(4, 3)); break;
    case 15: rp_a = div_rn(CHI(4, 1), CHI(2, 1)); rm_a = div_rn(CHI(4, 1), CHI(2, 1)); break;
    case 16: rp_a = div_rn(CHI(1, 6), CHI(6, 6)); break;
    default: break;
  }

  uint32_t mw[NL / 32], aw[NL / 32];
#pragma unroll
  for (int w = 0; w < NL / 32; w++) {
    mw[w] = w < ws.W ? ws.mask[((size_t)g * ws.W + w) * cap + c] : 0u;
    aw[w] = w < ws.W ? ws.anyc[(size_t)w * cap + c] : 0u;
  }

  // Per-layer results of the downward sweep that the upward sweep (k_lw_sweep) needs go to level-indexed scratch
  // records (layout in args.h): level = layer + 1 for the layer quantities, level = the layer's lower interface for the
  // downward fluxes.  v = 0 full (taug + taua), 1 clean (taug).
  float radld[2] = {0.f, 0.f}, radclrd[2] = {0.f, 0.f};
  int iclddn = 0;
  float fracs_bot = 0.f;
  const size_t pcap = ws.pcap;
  const int nv = do_clean ? 2 : 1;
  const unsigned lvstride = (unsigned)nv * NGLW;             // records per (tile, level)
  const size_t r0 = (size_t)(c / REC_TILE) * (nlay + 1) * nv * NGLW + g;      // record (this tile, level 0, stream 0, g)
  const int lane = c % REC_TILE;
  float *rec = ws.rec + r0 * LW_REC;                          // U | D records, see args.h
  float *recC = ws.recC + r0 * LW_REC_D;                      // cloudy-layer records
  // record lay + 1 holds, per stream, U of layer lay and the downward radiances D at the layer's LOWER interface (level lay),
  // so a layer writes one record and the sweep reads one record per step (the TOA downward radiance is zero: not stored)

  // Planck function at the top interface of the current layer; carried downwards
  auto planck_at = [&](float t) {
    int ind = (int)(t - 159.f);
    ind = min(max(ind, 1), 180);
    const float frac = t - 159.f - (float)ind;
    const float p0 = s_plk[ind - 1], p1 = s_plk[ind];
    return p0 + frac * (p1 - p0);
  };

  // Every workspace field this band reads is requested at the top of the layer iteration (one wait per layer instead of
  // one per field); BAND is a compile-time constant, so the other loads do not exist in this instantiation.
  constexpr unsigned COMMON = (1u << LWC_FAC00) | (1u << LWC_FAC01) | (1u << LWC_FAC10) | (1u << LWC_FAC11) | (1u << LWC_SELFFAC) |
                              (1u << LWC_SELFFRAC) | (1u << LWC_FORFAC) | (1u << LWC_FORFRAC) | (1u << LWC_TAVEL) | (1u << LWC_IDX);
  constexpr unsigned H2O = 1u << LWC_H2O, CO2 = 1u << LWC_CO2, O3 = 1u << LWC_O3, N2O = 1u << LWC_N2O, CO = 1u << LWC_CO, CH4 = 1u << LWC_CH4,
                     O2 = 1u << LWC_O2, BRD = 1u << LWC_BRD, MF = 1u << LWC_MINORFRAC, SM = 1u << LWC_SCALEMINOR, SN2 = 1u << LWC_SCALEMINORN2,
                     PAV = 1u << LWC_PAVEL, DRY = 1u << LWC_COLDRY;
  constexpr unsigned per_band[16] = {H2O | BRD | SN2 | MF | PAV, H2O | PAV, H2O | CO2 | N2O | MF | DRY, H2O | CO2 | O3, H2O | CO2 | O3 | MF | DRY,
                                     H2O | CO2 | MF | DRY, H2O | O3 | CO2 | MF | DRY, H2O | O3 | CO2 | N2O | MF | DRY, H2O | CH4 | N2O | MF | DRY,
                                     H2O, H2O | O2 | SM | MF, H2O | CO2, H2O | N2O | CO2 | CO | O3 | MF | DRY, CO2, N2O | CO2 | BRD | SM | MF, H2O | CH4};
  constexpr unsigned need = COMMON | per_band[BAND - 1];
  const unsigned ucap = (unsigned)cap, ustf = (unsigned)nlay * (unsigned)cap;     // 32-bit offsets: LWC_N*nlay*cap < 2^31
  const float *coefc = ws.coef + coef_index(0, 0, c, cap, LWC_N), *aerc = ws.aer + c + (unsigned)b * ustf;
  const unsigned lstride = (unsigned)cap * LWC_N;                                  // coefficient words per layer
  float tz_up = coefc[(size_t)((unsigned)(nlay - 1) * lstride) + LWC_TZ * 32];      // temperature of the interface above the layer
  float plev_up = planck_at(tz_up);
  // Software pipeline: the workspace words of layer lay-1 are requested right after the gas optics of layer lay have
  // consumed theirs (the registers are free then) and land while the radiative-transfer step of layer lay executes.
  float fv[LWC_N], taua_nx, tz_nx;
  auto load_layer = [&](int lay) {
    const float *p = coefc + (size_t)((unsigned)lay * lstride);      // fields at immediate offsets of 128 bytes
#pragma unroll
    for (int f = 0; f < LWC_N; f++) fv[f] = ((need >> f) & 1u) ? p[f * 32] : 0.f;
    taua_nx = aerc[(unsigned)lay * ucap];
    tz_nx = lay > 0 ? *(p + LWC_TZ * 32 - (ptrdiff_t)lstride) : ws.colf[(size_t)LWF_TZ0 * cap + c];
  };
  load_layer(nlay - 1);
  for (int lay = nlay - 1; lay >= 0; lay--) {
    const float taua = taua_nx, tz_dn = tz_nx;
    auto F = [&](int f) { return fv[f]; };
    const int pk = __float_as_int(F(LWC_IDX));
    const int jp = IDX_JP(pk), jt = IDX_JT(pk), jt1 = IDX_JT1(pk), indself = IDX_SELF(pk), indfor = IDX_FOR(pk), indminor = IDX_MINOR(pk);
    const float fac00 = F(LWC_FAC00), fac01 = F(LWC_FAC01), fac10 = F(LWC_FAC10), fac11 = F(LWC_FAC11);
    const bool low = lay < laytrop;
    const float *A = S + D.oA, *B = S + D.oB;
    auto selfk = [&]() { const float *r = S + D.oSelf + indself - 1; return F(LWC_SELFFAC) * (r[0] + F(LWC_SELFFRAC) * (r[1] - r[0])); };
    auto fork = [&]() { const float *r = S + D.oFor + indfor - 1; return F(LWC_FORFAC) * (r[0] + F(LWC_FORFRAC) * (r[1] - r[0])); };
    auto minor1 = [&](int off) { const float *r = S + off + indminor - 1; return r[0] + F(LWC_MINORFRAC) * (r[1] - r[0]); };
    auto minor2 = [&](int off, int ne, int jm, float fm) {
      const float *r = S + off + (jm - 1) + ne * (indminor - 1);
      const float m1 = r[0] + fm * (r[1] - r[0]);
      const float m2 = r[ne] + fm * (r[ne + 1] - r[ne]);
      return m1 + F(LWC_MINORFRAC) * (m2 - m1);
    };
    auto k4 = [&](const float *ab, bool lower) {
      int ind0, ind1;
      if (lower) { ind0 = ((jp - 1) * 5 + (jt - 1)); ind1 = (jp * 5 + (jt1 - 1)); }
      else { ind0 = ((jp - 13) * 5 + (jt - 1)); ind1 = ((jp - 12) * 5 + (jt1 - 1)); }
      return fac00 * ab[ind0] + fac10 * ab[ind0 + 1] + fac01 * ab[ind1] + fac11 * ab[ind1 + 1];
    };
    auto RAT = [&](int r, int jpp) { return s_rat[r * 60 + jpp - 1]; };
    auto major_lower2 = [&](const LwEta &e, const LwEta &e1) {
      const int ind0 = ((jp - 1) * 5 + (jt - 1)) * 9 + e.j, ind1 = (jp * 5 + (jt1 - 1)) * 9 + e1.j;
      return lw_major_lower(A, e, ind0, fac00, fac10) + lw_major_lower(A, e1, ind1, fac01, fac11);
    };
    auto major_upper = [&](const LwEta &e, const LwEta &e1) {
      const int ind0 = ((jp - 13) * 5 + (jt - 1)) * 5 + e.j, ind1 = ((jp - 12) * 5 + (jt1 - 1)) * 5 + e1.j;
      const float *q0 = B + ind0 - 1, *q1 = B + ind1 - 1;
      const float f0 = 1.f - e.f, f1 = 1.f - e1.f;
      return e.speccomb * ((f0 * fac00) * q0[0] + (e.f * fac00) * q0[1] + (f0 * fac10) * q0[5] + (e.f * fac10) * q0[6]) +
             e1.speccomb * ((f1 * fac01) * q1[0] + (e1.f * fac01) * q1[1] + (f1 * fac11) * q1[5] + (e1.f * fac11) * q1[6]);
    };
    auto frac_eta = [&](int off, const LwEta &ep) { const float *r = S + off + ep.j - 1; return r[0] + ep.f * (r[1] - r[0]); };
    // empirical column rescaling of a minor gas (e.g. LW:5208-5216)
    auto adjcol = [&](float col, int imol, float thresh, float base, float expo) {
      const float coldry = F(LWC_COLDRY);
      const float chim = CHI(imol, jp + 1);
      const float chi = col / coldry;
      const float rat = 1.e20f * chi / chim;
      if (rat > thresh) {
        const float adjfac = base + powf(rat - base, expo);
        return adjfac * chim * coldry * 1.e-20f;
      }
      return col;
    };
    auto WX = [&](float vmr) { return F(LWC_COLDRY) * vmr * 1.e-20f; };
    const float vccl4 = 0.093e-9f, vcfc11 = 0.251e-9f, vcfc12 = 0.538e-9f, vcfc22 = 0.169e-9f;

    float taug = 0.f, fracs = 0.f;
    switch (band) {
      case 1: {
        const float pp = F(LWC_PAVEL);
        const float scalen2 = F(LWC_BRD) * F(LWC_SCALEMINORN2);
        if (low) {
          float corradj = 1.f; if (pp < 250.f) corradj = 1.f - 0.15f * (250.f - pp) / 154.4f;
          taug = corradj * (F(LWC_H2O) * k4(A, true) + selfk() + fork() + scalen2 * minor1(D.oMinA[M_N2]));
          fracs = S[D.oFracA];
        } else {
          const float corradj = 1.f - 0.15f * (pp / 95.6f);
          taug = corradj * (F(LWC_H2O) * k4(B, false) + fork() + scalen2 * minor1(D.oMinB[M_N2]));
          fracs = S[D.oFracB];
        }
        break; }
      case 2: {
        if (low) {
          const float pp = F(LWC_PAVEL);
          const float corradj = 1.f - .05f * (pp - 100.f) / 900.f;
          taug = corradj * (F(LWC_H2O) * k4(A, true) + selfk() + fork());
          fracs = S[D.oFracA];
        } else { taug = F(LWC_H2O) * k4(B, false) + fork(); fracs = S[D.oFracB]; }
        break; }
      case 3: {
        const float h2o = F(LWC_H2O), co2 = F(LWC_CO2);
        const float mult = low ? 8.f : 4.f;
        const LwEta e = lw_eta(h2o, RAT(0, jp), co2, mult, oneminus), e1 = lw_eta(h2o, RAT(0, jp + 1), co2, mult, oneminus);
        const LwEta em = lw_eta(h2o, low ? rm_a : rm_b, co2, mult, oneminus);
        const LwEta ep = lw_eta(h2o, low ? rp_a : rp_b, co2, mult, oneminus);
        const float adjcoln2o = adjcol(F(LWC_N2O), 4, 1.5f, 0.5f, 0.65f);
        if (low) {
          taug = major_lower2(e, e1) + selfk() + fork() + adjcoln2o * minor2(D.oMinA[M_N2O], 9, em.j, em.f);
          fracs = frac_eta(D.oFracA, ep);
        } else {
          taug = major_upper(e, e1) + fork() + adjcoln2o * minor2(D.oMinB[M_N2O], 5, em.j, em.f);
          fracs = frac_eta(D.oFracB, ep);
        }
        break; }
      case 4: {
        const float co2 = F(LWC_CO2);
        if (low) {
          const float h2o = F(LWC_H2O);
          const LwEta e = lw_eta(h2o, RAT(0, jp), co2, 8.f, oneminus), e1 = lw_eta(h2o, RAT(0, jp + 1), co2, 8.f, oneminus);
          const LwEta ep = lw_eta(h2o, rp_a, co2, 8.f, oneminus);
          taug = major_lower2(e, e1) + selfk() + fork();
          fracs = frac_eta(D.oFracA, ep);
        } else {
          const float o3 = F(LWC_O3);
          const LwEta e = lw_eta(o3, RAT(5, jp), co2, 4.f, oneminus), e1 = lw_eta(o3, RAT(5, jp + 1), co2, 4.f, oneminus);
          const LwEta ep = lw_eta(o3, rp_b, co2, 4.f, oneminus);
          taug = major_upper(e, e1);
          fracs = frac_eta(D.oFracB, ep);
          const int ig = g - D.g0 + 1;
          if (ig == 8) taug = taug * 0.92f; else if (ig == 9) taug = taug * 0.88f; else if (ig == 10) taug = taug * 1.07f;
          else if (ig == 11) taug = taug * 1.1f; else if (ig == 12) taug = taug * 0.99f; else if (ig == 13) taug = taug * 0.88f;
          else if (ig == 14) taug = taug * 0.943f;
        }
        break; }
      case 5: {
        const float co2 = F(LWC_CO2);
        const float wx1 = WX(vccl4);
        if (low) {
          const float h2o = F(LWC_H2O);
          const LwEta e = lw_eta(h2o, RAT(0, jp), co2, 8.f, oneminus), e1 = lw_eta(h2o, RAT(0, jp + 1), co2, 8.f, oneminus);
          const LwEta em = lw_eta(h2o, rm_a, co2, 8.f, oneminus), ep = lw_eta(h2o, rp_a, co2, 8.f, oneminus);
          taug = major_lower2(e, e1) + selfk() + fork() + minor2(D.oMinA[M_O3], 9, em.j, em.f) * F(LWC_O3) + wx1 * S[D.oCfc + 0];
          fracs = frac_eta(D.oFracA, ep);
        } else {
          const float o3 = F(LWC_O3);
          const LwEta e = lw_eta(o3, RAT(5, jp), co2, 4.f, oneminus), e1 = lw_eta(o3, RAT(5, jp + 1), co2, 4.f, oneminus);
          const LwEta ep = lw_eta(o3, rp_b, co2, 4.f, oneminus);
          taug = major_upper(e, e1) + wx1 * S[D.oCfc + 0];
          fracs = frac_eta(D.oFracB, ep);
        }
        break; }
      case 6: {
        const float wx2 = WX(vcfc11), wx3 = WX(vcfc12);
        if (low) {
          const float adjcolco2 = adjcol(F(LWC_CO2), 2, 3.0f, 2.0f, 0.77f);
          taug = F(LWC_H2O) * k4(A, true) + selfk() + fork() + adjcolco2 * minor1(D.oMinA[M_CO2]) + wx2 * S[D.oCfc + 1] + wx3 * S[D.oCfc + 2];
        } else taug = 0.0f + wx2 * S[D.oCfc + 1] + wx3 * S[D.oCfc + 2];
        fracs = S[D.oFracA];
        break; }
      case 7: {
        if (low) {
          const float h2o = F(LWC_H2O), o3 = F(LWC_O3);
          const LwEta e = lw_eta(h2o, RAT(1, jp), o3, 8.f, oneminus), e1 = lw_eta(h2o, RAT(1, jp + 1), o3, 8.f, oneminus);
          const LwEta em = lw_eta(h2o, rm_a, o3, 8.f, oneminus), ep = lw_eta(h2o, rp_a, o3, 8.f, oneminus);
          const float adjcolco2 = adjcol(F(LWC_CO2), 2, 3.0f, 3.0f, 0.79f);
          taug = major_lower2(e, e1) + selfk() + fork() + adjcolco2 * minor2(D.oMinA[M_CO2], 9, em.j, em.f);
          fracs = frac_eta(D.oFracA, ep);
        } else {
          const float adjcolco2 = adjcol(F(LWC_CO2), 2, 3.0f, 2.0f, 0.79f);
          taug = F(LWC_O3) * k4(B, false) + adjcolco2 * minor1(D.oMinB[M_CO2]);
          fracs = S[D.oFracB];
          const int ig = g - D.g0 + 1;
          if (ig == 6) taug = taug * 0.92f; else if (ig == 7) taug = taug * 0.88f; else if (ig == 8) taug = taug * 1.07f;
          else if (ig == 9) taug = taug * 1.1f; else if (ig == 10) taug = taug * 0.99f; else if (ig == 11) taug = taug * 0.855f;
        }
        break; }
      case 8: {
        const float adjcolco2 = adjcol(F(LWC_CO2), 2, 3.0f, 2.0f, 0.65f);
        const float wx3 = WX(vcfc12), wx4 = WX(vcfc22);
        if (low) {
          taug = F(LWC_H2O) * k4(A, true) + selfk() + fork() + adjcolco2 * minor1(D.oMinA[M_CO2]) + F(LWC_O3) * minor1(D.oMinA[M_O3]) +
                 F(LWC_N2O) * minor1(D.oMinA[M_N2O]) + wx3 * S[D.oCfc + 2] + wx4 * S[D.oCfc + 3];
          fracs = S[D.oFracA];
        } else {
          taug = F(LWC_O3) * k4(B, false) + adjcolco2 * minor1(D.oMinB[M_CO2]) + F(LWC_N2O) * minor1(D.oMinB[M_N2O]) +
                 wx3 * S[D.oCfc + 2] + wx4 * S[D.oCfc + 3];
          fracs = S[D.oFracB];
        }
        break; }
      case 9: {
        const float adjcoln2o = adjcol(F(LWC_N2O), 4, 1.5f, 0.5f, 0.65f);
        if (low) {
          const float h2o = F(LWC_H2O), ch4 = F(LWC_CH4);
          const LwEta e = lw_eta(h2o, RAT(3, jp), ch4, 8.f, oneminus), e1 = lw_eta(h2o, RAT(3, jp + 1), ch4, 8.f, oneminus);
          const LwEta em = lw_eta(h2o, rm_a, ch4, 8.f, oneminus), ep = lw_eta(h2o, rp_a, ch4, 8.f, oneminus);
          taug = major_lower2(e, e1) + selfk() + fork() + adjcoln2o * minor2(D.oMinA[M_N2O], 9, em.j, em.f);
          fracs = frac_eta(D.oFracA, ep);
        } else {
          taug = F(LWC_CH4) * k4(B, false) + adjcoln2o * minor1(D.oMinB[M_N2O]);
          fracs = S[D.oFracB];
        }
        break; }
      case 10: case 11: {
        float t;
        if (low) { t = F(LWC_H2O) * k4(A, true) + selfk() + fork(); fracs = S[D.oFracA]; }
        else { t = F(LWC_H2O) * k4(B, false) + fork(); fracs = S[D.oFracB]; }
        if (band == 11) { const float scaleo2 = F(LWC_O2) * F(LWC_SCALEMINOR); t = t + scaleo2 * minor1(low ? D.oMinA[M_O2] : D.oMinB[M_O2]); }
        taug = t;
        break; }
      case 12: {
        if (low) {
          const float h2o = F(LWC_H2O), co2 = F(LWC_CO2);
          const LwEta e = lw_eta(h2o, RAT(0, jp), co2, 8.f, oneminus), e1 = lw_eta(h2o, RAT(0, jp + 1), co2, 8.f, oneminus);
          const LwEta ep = lw_eta(h2o, rp_a, co2, 8.f, oneminus);
          taug = major_lower2(e, e1) + selfk() + fork();
          fracs = frac_eta(D.oFracA, ep);
        }
        break; }
      case 13: {
        if (low) {
          const float h2o = F(LWC_H2O), n2o = F(LWC_N2O), co2 = F(LWC_CO2), coldry = F(LWC_COLDRY);
          const LwEta e = lw_eta(h2o, RAT(2, jp), n2o, 8.f, oneminus), e1 = lw_eta(h2o, RAT(2, jp + 1), n2o, 8.f, oneminus);
          const LwEta em = lw_eta(h2o, rm_a, n2o, 8.f, oneminus), eco = lw_eta(h2o, rm_a3, n2o, 8.f, oneminus);
          const LwEta ep = lw_eta(h2o, rp_a, n2o, 8.f, oneminus);
          const float chi_co2 = co2 / coldry;
          const float ratco2 = 1.e20f * chi_co2 / 3.55e-4f;
          float adjcolco2;
          if (ratco2 > 3.0f) { const float adjfac = 2.0f + powf(ratco2 - 2.0f, 0.68f); adjcolco2 = adjfac * 3.55e-4f * coldry * 1.e-20f; }
          else adjcolco2 = co2;
          taug = major_lower2(e, e1) + selfk() + fork() + adjcolco2 * minor2(D.oMinA[M_CO2], 9, em.j, em.f) +
                 F(LWC_CO) * minor2(D.oMinA[M_CO], 9, eco.j, eco.f);
          fracs = frac_eta(D.oFracA, ep);
        } else { taug = F(LWC_O3) * minor1(D.oMinB[M_O3]); fracs = S[D.oFracB]; }
        break; }
      case 14: {
        if (low) { taug = F(LWC_CO2) * k4(A, true) + selfk() + fork(); fracs = S[D.oFracA]; }
        else { taug = F(LWC_CO2) * k4(B, false); fracs = S[D.oFracB]; }
        break; }
      case 15: {
        if (low) {
          const float n2o = F(LWC_N2O), co2 = F(LWC_CO2);
          const LwEta e = lw_eta(n2o, RAT(4, jp), co2, 8.f, oneminus), e1 = lw_eta(n2o, RAT(4, jp + 1), co2, 8.f, oneminus);
          const LwEta em = lw_eta(n2o, rm_a, co2, 8.f, oneminus), ep = lw_eta(n2o, rp_a, co2, 8.f, oneminus);
          const float scalen2 = F(LWC_BRD) * F(LWC_SCALEMINOR);
          taug = major_lower2(e, e1) + selfk() + fork() + scalen2 * minor2(D.oMinA[M_N2], 9, em.j, em.f);
          fracs = frac_eta(D.oFracA, ep);
        }
        break; }
      default: {  // 16
        if (low) {
          const float h2o = F(LWC_H2O), ch4 = F(LWC_CH4);
          const LwEta e = lw_eta(h2o, RAT(3, jp), ch4, 8.f, oneminus), e1 = lw_eta(h2o, RAT(3, jp + 1), ch4, 8.f, oneminus);
          const LwEta ep = lw_eta(h2o, rp_a, ch4, 8.f, oneminus);
          taug = major_lower2(e, e1) + selfk() + fork();
          fracs = frac_eta(D.oFracA, ep);
        } else { taug = F(LWC_CH4) * k4(B, false); fracs = S[D.oFracB]; }
        break; }
    }
    if (a.dbg.taug) {
      const size_t q = ((size_t)(a.col0 + c) * nlay + lay) * NGLW + g;
      a.dbg.taug[q] = taug; a.dbg.taur[q] = fracs;
    }
    if (lay == 0) fracs_bot = fracs;

    // ---- rtrnmc downward step for this layer (LW:3207-3300)
    const float blay = planck_at(F(LWC_TAVEL));
    if (lay > 0) load_layer(lay - 1);          // prefetch (fv is dead from here on)
    const float plev_dn = planck_at(tz_dn);
    const float dplankup = plev_up - blay, dplankdn = plev_dn - blay;
    const float plfrac = fracs;
    const bool icldlyr = (aw[lay >> 5] >> (lay & 31)) & 1u;
    const bool cloudy = (mw[lay >> 5] >> (lay & 31)) & 1u;
    float odcld = 0.f, efclfrac = 0.f;
    const float cldfmc = cloudy ? 1.f : 0.f;
    if (cloudy) {
      const float taucmc = ws.cld[(size_t)((unsigned)b * ustf + (unsigned)lay * ucap) + c];
      if (a.dbg.taucmc) a.dbg.taucmc[((size_t)(a.col0 + c) * nlay + lay) * NGLW + g] = taucmc;
      odcld = secdiff * taucmc;
      const float transcld = expf(-odcld);
      const float abscld = 1.f - transcld;
      efclfrac = abscld * cldfmc;
    }
    if (icldlyr) iclddn = 1;
    float *qrec = rec + (size_t)((unsigned)(lay + 1) * lvstride) * LW_REC;
    float *qrecC = recC + (size_t)((unsigned)(lay + 1) * lvstride) * LW_REC_D;
#pragma unroll
    for (int v = 0; v < 2; v++) {
      if (v == 1 && !do_clean) break;
      const float taut = v == 0 ? taug + taua : taug;
      float odepth = secdiff * taut;
      if (odepth < 0.0f) odepth = 0.0f;
      float atrans, bbd, bbugas;
      if (icldlyr) {
        float odtot = odepth + odcld;
        float gassrc, bbdtot, atot, bbutot;
        if (odtot < 0.06f) {
          atrans = odepth - 0.5f * odepth * odepth;
          const float odepth_rec = 0.166667f * odepth;
          gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans;
          atot = odtot - 0.5f * odtot * odtot;
          const float odtot_rec = 0.166667f * odtot;
          bbdtot = plfrac * (blay + dplankdn * odtot_rec);
          bbd = plfrac * (blay + dplankdn * odepth_rec);
          bbugas = plfrac * (blay + dplankup * odepth_rec);
          bbutot = plfrac * (blay + dplankup * odtot_rec);
        } else if (odepth <= 0.06f) {
          atrans = odepth - 0.5f * odepth * odepth;
          const float odepth_rec = 0.166667f * odepth;
          gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans;
          const float tblind = div_rn(odtot, __fadd_rn(bpade, odtot));
          const int ittot = (int)__fadd_rn(__fmul_rn(10000.0f, tblind), 0.5f);
          const float2 et = s_et[ittot];
          const float tfactot = et.y;
          bbdtot = plfrac * (blay + tfactot * dplankdn);
          bbd = plfrac * (blay + dplankdn * odepth_rec);
          atot = 1.f - et.x;
          bbugas = plfrac * (blay + dplankup * odepth_rec);
          bbutot = plfrac * (blay + tfactot * dplankup);
        } else {
          float tblind = div_rn(odepth, __fadd_rn(bpade, odepth));
          const int itgas = (int)__fadd_rn(__fmul_rn(10000.0f, tblind), 0.5f);
          // tau_tbl(itgas) recomputed with the table generator's arithmetic (LW:7944-7950)
          if (itgas >= 10000) odepth = 1.e10f;
          else { const float tfn = div_rn((float)itgas, 10000.0f); odepth = div_rn(__fmul_rn(bpade, tfn), __fsub_rn(1.0f, tfn)); }
          const float2 eg = s_et[itgas];
          atrans = 1.f - eg.x;
          const float tfacgas = eg.y;
          gassrc = atrans * plfrac * (blay + tfacgas * dplankdn);
          odtot = odepth + odcld;
          tblind = div_rn(odtot, __fadd_rn(bpade, odtot));
          const int ittot = (int)__fadd_rn(__fmul_rn(10000.0f, tblind), 0.5f);
          const float2 et = s_et[ittot];
          const float tfactot = et.y;
          bbdtot = plfrac * (blay + tfactot * dplankdn);
          bbd = plfrac * (blay + tfacgas * dplankdn);
          atot = 1.f - et.x;
          bbugas = plfrac * (blay + tfacgas * dplankup);
          bbutot = plfrac * (blay + tfactot * dplankup);
        }
        radld[v] = radld[v] - radld[v] * (atrans + efclfrac * (1.f - atrans)) + gassrc + cldfmc * (bbdtot * atot - gassrc);
        // the same step for the upward radiance is radlu - radlu * X + Y (LW:3334-3338): hand over X and Y
        const float gassrcu = bbugas * atrans;
        reinterpret_cast<float2 *>(qrecC + v * (NGLW * LW_REC_D))[lane] =
            make_float2(atrans + efclfrac * (1.f - atrans), gassrcu + cldfmc * (bbutot * atot - gassrcu));
      } else {
        if (odepth <= 0.06f) {
          atrans = odepth - 0.5f * odepth * odepth;
          odepth = 0.166667f * odepth;
          bbd = plfrac * (blay + dplankdn * odepth);
          bbugas = plfrac * (blay + dplankup * odepth);
        } else {
          const float tblind = div_rn(odepth, __fadd_rn(bpade, odepth));
          const int itr = (int)__fadd_rn(__fmul_rn(10000.0f, tblind), 0.5f);
          const float2 et = s_et[itr];
          atrans = 1.f - et.x;
          const float tausfac = et.y;
          bbd = plfrac * (blay + tausfac * dplankdn);
          bbugas = plfrac * (blay + tausfac * dplankup);
        }
        radld[v] = radld[v] + (bbd - radld[v]) * atrans;
      }
      if (iclddn == 1) radclrd[v] = radclrd[v] + (bbd - radclrd[v]) * atrans;
      else radclrd[v] = radld[v];
      reinterpret_cast<float4 *>(qrec + v * (NGLW * LW_REC))[lane] = make_float4(atrans, bbugas, radld[v], radclrd[v]);
    }
    plev_up = plev_dn;
  }
  // ---- surface (LW:3303-3320)
  const float emis = ws.colf[(size_t)LWF_EMISS * cap + c];
  float plankbnd;
  {
    const float tbound = ws.colf[(size_t)LWF_TBOUND * cap + c];
    int ind = (int)(tbound - 159.f);
    ind = min(max(ind, 1), 180);
    const float frac = tbound - 159.f - (float)ind;
    const float dbdtlev = s_plk[ind] - s_plk[ind - 1];
    plankbnd = emis * (s_plk[ind - 1] + frac * dbdtlev);
  }
  const float rad0 = fracs_bot * plankbnd;
  const float reflect = 1.f - emis;
  // upward radiances leaving the surface; the upward sweep itself runs in k_lw_sweep
  ws.scrS[(size_t)g * pcap + c] = make_float2(rad0 + reflect * radld[0], rad0 + reflect * radclrd[0]);
  if (do_clean) ws.scrS[(size_t)(NGLW + g) * pcap + c] = make_float2(rad0 + reflect * radld[1], rad0 + reflect * radclrd[1]);
}

template <int NL>
__global__ void __launch_bounds__(LW_BLOCK, 2) k_lw_solve(LwArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *s_et = reinterpret_cast<float2 *>(smem_raw);                // 10002 x (exp_tbl, tfn_tbl)
  float *S = reinterpret_cast<float *>(s_et + 10002);                 // slice
  float *s_plk = S + SLICE_MAX;                                       // totplnk(1:181, band), padded to 184
  float *s_rat = s_plk + 184;                                         // [6][60] chi_mls ratios by jp
  float *s_chi = s_rat + 6 * 60;                                      // chi_mls(7,59)
  uint64_t *bar = reinterpret_cast<uint64_t *>(s_chi + 416);

  // block order: band-major, then column tile, then g-point within the band (see k_sw_solve)
  const int ntiles = (a.ncols + LW_BLOCK - 1) / LW_BLOCK;
  int b = 0;
  while (b < NBLW - 1 && (int)blockIdx.x >= c_lw[b + 1].g0 * ntiles) b++;
  const LwBandDesc &D = c_lw[b];
  const int rblk = blockIdx.x - D.g0 * ntiles;
  const int tile = rblk / D.ng;
  const int g = D.g0 + rblk % D.ng;
  const DevTables &tb = a.tb;
  {
    StageReq req[3] = {{s_et, tb.lw_exptfn, 10002 * 8},
                       {S, tb.lw_tab + D.slice_base + (size_t)D.slice_floats * (g - D.g0), (uint32_t)D.slice_floats * 4},
                       {s_plk, tb.totplnk + 184 * b, 184 * 4}};
    stage_tables(bar, req, 3);
  }
  // chi_mls(7,59) ratios: 0 h2o/co2, 1 h2o/o3, 2 h2o/n2o, 3 h2o/ch4, 4 n2o/co2, 5 o3/co2   (setcoef LW:3700-3760)
  for (int t = threadIdx.x; t < 6 * 59; t += blockDim.x) {
    const int r = t / 59, jp = t % 59;          // jp 0-based
    const float *chi = tb.chi_mls + 7 * jp;
    const int num[6] = {0, 0, 0, 0, 3, 2}, den[6] = {1, 2, 3, 5, 1, 1};
    s_rat[r * 60 + jp] = div_rn(chi[num[r]], chi[den[r]]);
  }
  for (int t = threadIdx.x; t < 7 * 59; t += blockDim.x) s_chi[t] = tb.chi_mls[t];
  __syncthreads();
  const int c = tile * LW_BLOCK + threadIdx.x;
  if (c >= a.ncols) return;
  const LwSmem sm{s_et, S, s_plk, s_rat, s_chi};
  switch (b) {
    case 0: lw_solve_band<NL, 1>(a, sm, D, g, c); break;
    case 1: lw_solve_band<NL, 2>(a, sm, D, g, c); break;
    case 2: lw_solve_band<NL, 3>(a, sm, D, g, c); break;
    case 3: lw_solve_band<NL, 4>(a, sm, D, g, c); break;
    case 4: lw_solve_band<NL, 5>(a, sm, D, g, c); break;
    case 5: lw_solve_band<NL, 6>(a, sm, D, g, c); break;
    case 6: lw_solve_band<NL, 7>(a, sm, D, g, c); break;
    case 7: lw_solve_band<NL, 8>(a, sm, D, g, c); break;
    case 8: lw_solve_band<NL, 9>(a, sm, D, g, c); break;
    case 9: lw_solve_band<NL, 10>(a, sm, D, g, c); break;
    case 10: lw_solve_band<NL, 11>(a, sm, D, g, c); break;
    case 11: lw_solve_band<NL, 12>(a, sm, D, g, c); break;
    case 12: lw_solve_band<NL, 13>(a, sm, D, g, c); break;
    case 13: lw_solve_band<NL, 14>(a, sm, D, g, c); break;
    case 14: lw_solve_band<NL, 15>(a, sm, D, g, c); break;
    default: lw_solve_band<NL, 16>(a, sm, D, g, c); break;
  }
}

static int lw_solve_smem() { return 10002 * 8 + (SLICE_MAX + 184 + 6 * 60 + 416) * 4 + 16; }

void launch_lw_solve(const LwArgs &a, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_lw_solve<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, lw_solve_smem());
    cudaFuncSetAttribute(k_lw_solve<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, lw_solve_smem());
    cudaFuncSetAttribute(k_lw_solve<160>, cudaFuncAttributeMaxDynamicSharedMemorySize, lw_solve_smem());
    attr = true;
  }
  dim3 grid(NGLW * ((a.ncols + LW_BLOCK - 1) / LW_BLOCK));
  if (a.ws.nlay <= 64) k_lw_solve<64><<<grid, LW_BLOCK, lw_solve_smem(), s>>>(a);
  else if (a.ws.nlay <= 128) k_lw_solve<128><<<grid, LW_BLOCK, lw_solve_smem(), s>>>(a);
  else k_lw_solve<160><<<grid, LW_BLOCK, lw_solve_smem(), s>>>(a);
  count_launch();
}

// ------------------------------------------------------------------------------------------------------
// Upward sweep of rtrnmc (LW:3322-3356) + ordered sum over the g-points of a band (LW:3365-3395).
// One thread per (column, band, stream), block = one 128-column record tile: the NG upward radiances of the band's
// g-points are its register state; per step it requests the NG float4 records of that level together (written by
// k_lw_solve; the addresses do not depend on the recurrence, so the memory system sees NG independent 2 KB loads per
// block), advances the NG two-term recurrences and adds the NG up / down radiances in g order.  It writes ONE band partial
// [band][level][kind][c] per kind; no per-g flux is ever stored.  No shared memory, no barriers, no atomics.
// HBM-bound: 16 B per (column, g, level, stream).
template <int NG>
__global__ void __launch_bounds__(128) k_lw_sweep(LwArgs a, int grp, int g0) {
  const LwWs &ws = a.ws;
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= a.ncols) return;
  const int v = blockIdx.y;                    // 0 full (+ clear), 1 clean (+ clean-clear)
  const int nlay = ws.nlay, nk = ws.nk;
  const size_t pcap = ws.pcap, cap = ws.cap;
  const int nv = gridDim.y, lane = threadIdx.x;                // block = one record tile
  const unsigned lvstride = (unsigned)nv * NGLW;               // records per (tile, level)
  const size_t r0 = ((size_t)blockIdx.x * (nlay + 1) * nv + v) * NGLW + g0;
  const float *__restrict__ rec = ws.rec + r0 * LW_REC;
  const float *__restrict__ recC = ws.recC + r0 * LW_REC_D;

  bool iclddn = false;                         // the flag the downward sweep leaves behind (LW:3218): any cloud in the column
  for (int w = 0; w < ws.W; w++) iclddn = iclddn || ws.anyc[(size_t)w * cap + c] != 0u;

  float rl[NG], rc[NG];
  float *__restrict__ bpart = ws.bpart + (size_t)grp * (nlay + 1) * nk * pcap + c;
  const int kU = ws.kslot[v == 0 ? K_FU : K_NU], kD = ws.kslot[v == 0 ? K_FD : K_ND];
  const int kCU = ws.kslot[v == 0 ? K_CU : K_XU], kCD = ws.kslot[v == 0 ? K_CD : K_XD];
  const bool clr = v == 0 || (a.variants & ARC_VAR_CLEANCLEAR) != 0;
  {   // level 0: the upward radiances leaving the surface
    float sU = 0.f, sCU = 0.f;
#pragma unroll
    for (int i = 0; i < NG; i++) {
      const float2 s0 = ws.scrS[((size_t)v * NGLW + g0 + i) * pcap + c];
      rl[i] = s0.x; rc[i] = s0.y;
    }
#pragma unroll
    for (int i = 0; i < NG; i++) { sU = sU + rl[i]; sCU = sCU + rc[i]; }
    __stcs(bpart + (size_t)kU * pcap, sU);
    if (clr) __stcs(bpart + (size_t)kCU * pcap, sCU);
  }
  uint32_t aw = 0u;
  for (int lev = 1; lev <= nlay; lev++) {
    // record lev: U of layer lev-1 (-> upward radiances at level lev) and the downward radiances at level lev-1
    const int lay = lev - 1;
    if ((lay & 31) == 0) aw = ws.anyc[(size_t)(lay >> 5) * cap + c];
    const bool icldlyr = (aw >> (lay & 31)) & 1u;
    float2 u[NG], d[NG];
    const float *__restrict__ q = rec + (size_t)((unsigned)lev * lvstride) * LW_REC;       // the group's records: immediate offsets
#pragma unroll
    for (int i = 0; i < NG; i++) {
      const float4 r = __ldcs(reinterpret_cast<const float4 *>(q + i * LW_REC) + lane);
      u[i] = make_float2(r.x, r.y); d[i] = make_float2(r.z, r.w);
    }
    float sU = 0.f, sCU = 0.f, sD = 0.f, sCD = 0.f;
    if (icldlyr) {
#pragma unroll
      for (int i = 0; i < NG; i++) {
        const float2 xy = __ldcs(reinterpret_cast<const float2 *>(recC + ((size_t)((unsigned)lev * lvstride) + i) * LW_REC_D) + lane);
        rl[i] = rl[i] - rl[i] * xy.x + xy.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NG; i++) rl[i] = rl[i] + (u[i].y - rl[i]) * u[i].x;
    }
#pragma unroll
    for (int i = 0; i < NG; i++) {
      if (iclddn) rc[i] = rc[i] + (u[i].y - rc[i]) * u[i].x;
      else rc[i] = rl[i];
      sU = sU + rl[i]; sCU = sCU + rc[i]; sD = sD + d[i].x; sCD = sCD + d[i].y;
    }
    float *bp = bpart + (size_t)lev * nk * pcap;
    __stcs(bp + (size_t)kU * pcap, sU); __stcs(bp - (size_t)nk * pcap + (size_t)kD * pcap, sD);
    if (clr) { __stcs(bp + (size_t)kCU * pcap, sCU); __stcs(bp - (size_t)nk * pcap + (size_t)kCD * pcap, sCD); }
  }
  {   // TOA: no downward radiance
    float *bp = bpart + (size_t)nlay * nk * pcap;
    __stcs(bp + (size_t)kD * pcap, 0.f);
    if (clr) __stcs(bp + (size_t)kCD * pcap, 0.f);
  }
}

int lw_sweep_groups() { return h_lw_grp.n; }
void launch_lw_sweep(const LwArgs &a, cudaStream_t s) {
  const dim3 grid((a.ncols + 127) / 128, (a.variants & ARC_VAR_CLEAN) ? 2 : 1);
  for (int q = 0; q < h_lw_grp.n; q++) {
    const int g0 = h_lw_grp.g0[q];
    switch (h_lw_grp.ng[q]) {
#define SWEEP_CASE(N) case N: k_lw_sweep<N><<<grid, 128, 0, s>>>(a, q, g0); break;
      SWEEP_CASE(1) SWEEP_CASE(2) SWEEP_CASE(3) SWEEP_CASE(4) SWEEP_CASE(5) SWEEP_CASE(6) SWEEP_CASE(7) SWEEP_CASE(8)
      SWEEP_CASE(9) SWEEP_CASE(10) SWEEP_CASE(11) SWEEP_CASE(12) SWEEP_CASE(13) SWEEP_CASE(14) SWEEP_CASE(15) SWEEP_CASE(16)
#undef SWEEP_CASE
      default: break;
    }
  }
  count_launch(h_lw_grp.n);
}

// ------------------------------------------------------------------------------------------------------
// Reduction: per band the sum of its sweep-group partials (its g-points were added in index order by k_lw_sweep),
// x wtdiff x delwave, sum over bands, x fluxfac (LW:3365-3395); heating rates (LW:3397-3408); scatter (LW:12646-12692).
// Block = 64 columns x 4 level-lanes.
constexpr int RED_CX = 64, RED_LY = 4;
__global__ void __launch_bounds__(RED_CX * RED_LY, 4) k_lw_reduce(LwArgs a) {
  __shared__ float s_net[161][RED_CX];
  const int cx = threadIdx.x, ly = threadIdx.y;
  const int c = blockIdx.x * RED_CX + cx;
  const Geo &G = a.geo;
  const LwWs &ws = a.ws;
  const int nlay = ws.nlay, nz = G.kte - G.kts + 1;
  const size_t cap = ws.pcap;      // the reduce only touches the partial buffer
  const bool active = c < a.ncols;
  const int tc = a.col0 + c;
  int i = 0, j = 0; size_t ij = 0;
  if (active) { G.ij(tc, i, j); ij = G.at2(i, j); }
  const bool do_clean = (a.variants & ARC_VAR_CLEAN) != 0;
  const float wtdiff = 0.5f;
  const int nk = ws.nk;
  const bool do_clnc = (a.variants & ARC_VAR_CLEANCLEAR) != 0;
  const size_t gstride = (size_t)(nlay + 1) * nk * cap;
  const unsigned ucap = (unsigned)cap;       // 32-bit kind offsets: nk * pcap < 2^31
  const unsigned oFU = ws.kslot[K_FU] * ucap, oFD = ws.kslot[K_FD] * ucap, oCU = ws.kslot[K_CU] * ucap, oCD = ws.kslot[K_CD] * ucap,
                 oNU = ws.kslot[K_NU] * ucap, oND = ws.kslot[K_ND] * ucap, oXU = ws.kslot[K_XU] * ucap, oXD = ws.kslot[K_XD] * ucap;
  for (int lev = ly; lev <= nlay && active; lev += RED_LY) {
    float tot[NKIND];
#pragma unroll
    for (int k = 0; k < NKIND; k++) tot[k] = 0.f;
    const float *p = ws.bpart + ((size_t)lev * nk) * cap + c;      // band sums from k_lw_sweep
    float r[NKIND];
#pragma unroll
    for (int k = 0; k < NKIND; k++) r[k] = 0.f;
#pragma unroll 4
    for (int q = 0; q < a.ngroups; q++, p += gstride) {       // sweep groups in g order; a band's groups are consecutive
      r[K_FU] = r[K_FU] + p[oFU]; r[K_FD] = r[K_FD] + p[oFD]; r[K_CU] = r[K_CU] + p[oCU]; r[K_CD] = r[K_CD] + p[oCD];
      if (do_clean) { r[K_NU] = r[K_NU] + p[oNU]; r[K_ND] = r[K_ND] + p[oND]; }
      if (do_clnc) { r[K_XU] = r[K_XU] + p[oXU]; r[K_XD] = r[K_XD] + p[oXD]; }
      const int b = c_lw_grp_band[q];
      if (q + 1 == a.ngroups || c_lw_grp_band[q + 1] != b) {
        const float dw = a.tb.delwave[b];
#pragma unroll
        for (int k = 0; k < NKIND; k++) { tot[k] = tot[k] + (r[k] * wtdiff) * dw; r[k] = 0.f; }
      }
    }
#pragma unroll
    for (int k = 0; k < NKIND; k++) tot[k] = tot[k] * a.tb.fluxfac;
    s_net[lev][cx] = tot[K_FU] - tot[K_FD];
    if (lev <= nz + 1 && a.lwupflx) {
      const size_t q = G.atp(i, G.kts + lev, j);
      a.lwupflx[q] = tot[K_FU]; a.lwupflxc[q] = tot[K_CU]; a.lwdnflx[q] = tot[K_FD]; a.lwdnflxc[q] = tot[K_CD];
      a.lwupflxcln[q] = tot[K_NU]; a.lwdnflxcln[q] = tot[K_ND];
    }
    if (lev == 0) {
      a.glw[ij] = tot[K_FD];
      if (a.lwupt) { a.lwupb[ij] = tot[K_FU]; a.lwupbc[ij] = tot[K_CU]; a.lwdnb[ij] = tot[K_FD]; a.lwdnbc[ij] = tot[K_CD];
                     a.lwupbcln[ij] = tot[K_NU]; a.lwdnbcln[ij] = tot[K_ND]; }
      if (a.lwuptclnc) { a.lwupbclnc[ij] = tot[K_XU]; a.lwdnbclnc[ij] = tot[K_XD]; }
    }
    if (lev == nlay) {
      a.olr[ij] = tot[K_FU];
      a.lwcf[ij] = tot[K_CU] - tot[K_FU];
      if (a.lwupt) { a.lwupt[ij] = tot[K_FU]; a.lwuptc[ij] = tot[K_CU]; a.lwdnt[ij] = tot[K_FD]; a.lwdntc[ij] = tot[K_CD];
                     a.lwuptcln[ij] = tot[K_NU]; a.lwdntcln[ij] = tot[K_ND]; }
      if (a.lwuptclnc) { a.lwuptclnc[ij] = tot[K_XU]; a.lwdntclnc[ij] = tot[K_XD]; }
    }
  }
  __syncthreads();
  for (int L = 1 + ly; L <= nz && active; L += RED_LY) {
    const int k = G.kts + L - 1;
    const float pz0 = a.p8w[G.at3(i, k, j)] / 100.f, pz1 = a.p8w[G.at3(i, k + 1, j)] / 100.f;
    const float htr = a.tb.heatfac * (s_net[L - 1][cx] - s_net[L][cx]) / (pz0 - pz1);
    const float tten = htr / 86400.f;
    a.rthratenlw[G.at3(i, k, j)] = tten / a.pi3d[G.at3(i, k, j)];
    if (a.dbg.hr) a.dbg.hr[(size_t)tc * nlay + L - 1] = htr;
  }
}
void launch_lw_reduce(const LwArgs &a, cudaStream_t s) {
  k_lw_reduce<<<(a.ncols + RED_CX - 1) / RED_CX, dim3(RED_CX, RED_LY), 0, s>>>(a);
  count_launch();
}

}  // namespace arc
