// Longwave spectral solver.  k_lw_band = taumol (taugb1..16) + both sweeps of rtrnmc for the full and the clean (aerosol-free)
// call, one thread per (column, band group); k_lw_reduce = band sum, heating rates and scatter.
//
// Reference (module_ra_rrtmg_lw.F v3.9.1): taumol 4712-7828, rtrnmc 2974-3410, rrtmg_lw 10984-11044
// (taut = taug + taua; clean rtrnmc(taug) then rtrnmc(taut)), RRTMG_LWRAD output scatter 12646-12692.
//
// Mapping.  The reference evaluates, per layer, everything that depends on the band only (the binary-species parameter eta,
// its table indices js / ind0 / ind1, the interpolation weights, minor-gas columns, Planck-fraction abscissa) once and then
// loops over the band's g-points (LW:5185-5349).  So does this kernel: a thread owns one column and the NG <= 8 consecutive
// g-points of one band ("band group"), walks the layers, computes the band-level values once per layer and then the NG gas
// optical depths / Planck fractions with table addresses that differ between g-points by compile-time constants only
// (the slice layout of every band is a compile-time constant, LWK below).  Lanes = 32 neighbouring columns.
//
// No level records.  rtrnmc's downward and upward sweeps each need the layer transmittance and source of every layer; the
// radiances themselves are two-term recurrences.  Instead of handing the layer quantities of the first sweep to the second
// through memory (the round-1 design: 32 B per column x g-point x level written and read back, 73 GB per C2 step), the
// thread makes two passes over the layers and recomputes the gas optics in the second one: top-down (downward radiances,
// kept in registers: NG x streams x {all-sky, clear-sky}) then, from the surface, bottom-up.  The recomputation is ~35
// instructions per (g-point, layer); the record traffic, the 19 GB of record buffers and the separate sweep kernels are gone.
// Sums over the group's g-points are formed in index order at every level (the reference's accumulation order,
// LW:3365-3395) and written as one partial per (group, level, kind); k_lw_reduce adds the groups in band order.
#include "args.h"
#include "glibc_math.cuh"
#include "../../include/arc_rad.h"

namespace arc {

static __constant__ LwBandDesc c_lw[16];
static __constant__ int c_lw_ngb[NGLW];    // band index 0..15 of each LW g-point
#ifndef LW_GMAX
#define LW_GMAX 8
#endif
static SweepGroups h_lw_grp;               // band groups: consecutive g-points of one band, at most LW_GMAX
static __constant__ int c_lw_grp_band[SWEEP_MAXGRP];
static __constant__ int c_lw_grp_g0[SWEEP_MAXGRP];
static bool h_lw_layout_ok = true;

// Compile-time slice layout of each band (floats inside one g-point's slice, tables.cpp; oA = 0).  The layout follows from
// RRTMG's fixed table shapes (nspa / nspb and the minor-gas tables of each band); upload_band_descs_lw checks it against the
// descriptors the host table builder produced.
struct LwK { int sf, oB, oSelf, oFor, oFracA, oFracB, oCfc, mA[M_COUNT], mB[M_COUNT]; };
#define NONE6 {-1, -1, -1, -1, -1, -1}
static constexpr LwK LWK[16] = {
    {408, 80, 328, 340, 344, 356, 404, {364, -1, -1, -1, -1, -1}, {384, -1, -1, -1, -1, -1}},
    {368, 80, 328, 340, 344, 356, 364, NONE6, NONE6},
    {2096, 600, 1788, 1800, 1804, 1816, 2092, {-1, 1824, -1, -1, -1, -1}, {-1, 1996, -1, -1, -1, -1}},
    {1828, 600, 1788, 1800, 1804, 1816, 1824, NONE6, NONE6},
    {2000, 600, 1788, 1800, 1804, 1816, 1996, {-1, -1, 1824, -1, -1, -1}, NONE6},
    {152, 80, 92, 104, 108, 120, 148, {-1, -1, -1, 128, -1, -1}, NONE6},
    {1080, 600, 848, 860, 864, 876, 1076, {-1, -1, -1, 884, -1, -1}, {-1, -1, -1, 1056, -1, -1}},
    {468, 80, 328, 340, 344, 356, 464, {-1, 364, 404, 424, -1, -1}, {-1, 384, -1, 444, -1, -1}},
    {1080, 600, 848, 860, 864, 876, 1076, {-1, 884, -1, -1, -1, -1}, {-1, 1056, -1, -1, -1, -1}},
    {368, 80, 328, 340, 344, 356, 364, NONE6, NONE6},
    {408, 80, 328, 340, 344, 356, 404, {-1, -1, -1, -1, -1, 364}, {-1, -1, -1, -1, -1, 384}},
    {652, 600, 612, 624, 628, 640, 648, NONE6, NONE6},
    {1016, 600, 612, 624, 628, 640, 1012, {-1, -1, -1, 668, 840, -1}, {-1, -1, 648, -1, -1, -1}},
    {368, 80, 328, 340, 344, 356, 364, NONE6, NONE6},
    {824, 600, 612, 624, 628, 640, 820, {648, -1, -1, -1, -1, -1}, NONE6},
    {888, 600, 848, 860, 864, 876, 884, NONE6, NONE6}};
#undef NONE6
constexpr int LW_SLICE_MAX = 2096;
#ifndef LW_PF
#define LW_PF 3
#endif

void upload_band_descs_lw(const HostTables &T) {
  cudaMemcpyToSymbol(c_lw, T.lw, sizeof(LwBandDesc) * 16);
  int ngs[16], g0s[16];
  for (int b = 0; b < 16; b++) { ngs[b] = T.lw[b].ng; g0s[b] = T.lw[b].g0; }
  h_lw_grp = make_sweep_groups(ngs, g0s, 16, LW_GMAX);
  cudaMemcpyToSymbol(c_lw_grp_band, h_lw_grp.band, sizeof(int) * SWEEP_MAXGRP);
  cudaMemcpyToSymbol(c_lw_grp_g0, h_lw_grp.g0, sizeof(int) * SWEEP_MAXGRP);
  int ngb[NGLW];
  for (int i = 0; i < NGLW; i++) ngb[i] = T.lw_ngb[i] - 1;
  cudaMemcpyToSymbol(c_lw_ngb, ngb, sizeof(int) * NGLW);
  h_lw_layout_ok = true;
  const int ngc[16] = {10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2};
  for (int b = 0; b < 16; b++) {
    const LwBandDesc &D = T.lw[b];
    const LwK &K = LWK[b];
    bool ok = D.slice_floats == K.sf && D.oA == 0 && D.oB == K.oB && D.oSelf == K.oSelf && D.oFor == K.oFor && D.oFracA == K.oFracA &&
              D.oFracB == K.oFracB && D.oCfc == K.oCfc && D.ng == ngc[b];
    for (int m = 0; m < M_COUNT; m++) ok = ok && D.oMinA[m] == K.mA[m] && D.oMinB[m] == K.mB[m];
    if (!ok) h_lw_layout_ok = false;
  }
}
bool lw_layout_ok() { return h_lw_layout_ok; }

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

// Stencil of the lower-atmosphere major-species interpolation in eta at one pressure level (LW:5219-5349): three table
// entries from offset `o` (relative to the (p, T) row) with weights w[0..2]; the usual 2-point interpolation has w[2] = 0.
struct LwStencil { int o; float w0, w1, w2; bool three; };
__device__ __forceinline__ LwStencil lw_stencil(const LwEta &e) {
  LwStencil s;
  if (e.specparm < 0.125f) {
    const float q = e.f - 1;
    const float p4 = pow4f(q);
    s.o = e.j - 1; s.w0 = p4; s.w1 = 1 - q - 2.0f * p4; s.w2 = q + p4; s.three = true;
  } else if (e.specparm > 0.875f) {
    const float q = -e.f;
    const float p4 = pow4f(q);
    s.o = e.j - 2; s.w0 = q + p4; s.w1 = 1 - q - 2.0f * p4; s.w2 = p4; s.three = true;
  } else {
    s.o = e.j - 1; s.w0 = 1.f - e.f; s.w1 = e.f; s.w2 = 0.f; s.three = false;
  }
  return s;
}

#ifndef LW_BLOCK_SZ
#define LW_BLOCK_SZ 512
#endif
constexpr int LW_BLOCK = LW_BLOCK_SZ;   // one block per SM: 80 KB exp / tfn table + up to 67 KB of g-point slices
struct LwSmem { const float2 *et; const float *S, *plk, *rat, *chi; };

// Band-level values of one layer: everything taumol computes before its loop over g-points (LW:5185-5230 and the
// corresponding lines of every taugbN).  Pointers address g-point 0 of the group; g-point i is `+ i * slice floats`.
struct LwBL {
  bool low, three, any3;       // three: this lane uses the 3-point eta stencil; any3: some lane of the warp does
  float fac00, fac01, fac10, fac11;
  const float *pA0, *pA1;                  // major-species rows at (jp, jt) and (jp+1, jt1), eta stencil offset included
  float a0, a1, a2, b0, b1, b2, c0, c1, c2, d0, d1, d2;   // eta stencil weights x fac00 / fac10 (level jp) and x fac01 / fac11 (level jp+1)
  float sc0, sc1;                          // speccomb at the two levels (binary bands) / key-species column (others)
  const float *pSelf, *pFor, *pM0, *pM1, *pM2, *pFrac, *pCfc;
  float selffac, selffrac, forfac, forfrac, minorfrac, fm0, fm1, fFrac;
  float x0, x1, x2, x3, x4;                // band-specific scalars (rescaled minor columns, CFC columns, pressure correction)
};

// Gas optical depth and Planck fraction of g-point i of the group (off = i * slice floats, ig = 1-based g-point inside the
// band).  Expressions and operation order are those of taugb1..16 (LW:4961-7826).
template <int BAND>
__device__ __forceinline__ void lw_gas(const LwBL &L, const int off, const int ig, float &taug, float &fracs) {
  auto selfk = [&]() { const float *r = L.pSelf + off; return L.selffac * (r[0] + L.selffrac * (r[1] - r[0])); };
  auto fork = [&]() { const float *r = L.pFor + off; return L.forfac * (r[0] + L.forfrac * (r[1] - r[0])); };
  auto minor1 = [&](const float *p) { const float *r = p + off; return r[0] + L.minorfrac * (r[1] - r[0]); };
  auto minor2 = [&](const float *p, int ne, float fm) {
    const float *r = p + off;
    const float m1 = r[0] + fm * (r[1] - r[0]);
    const float m2 = r[ne] + fm * (r[ne + 1] - r[ne]);
    return m1 + L.minorfrac * (m2 - m1);
  };
  auto k4 = [&]() { const float *q0 = L.pA0 + off, *q1 = L.pA1 + off; return L.fac00 * q0[0] + L.fac10 * q0[1] + L.fac01 * q1[0] + L.fac11 * q1[1]; };
  // binary bands, lower atmosphere: eta stencil at both pressure levels, T stride 9 (LW:5219-5349)
  auto major_lower2 = [&]() {
    const float *p = L.pA0 + off, *q = L.pA1 + off;
    if (L.any3)     // warp-uniform: some lane sits in the 3-point region (eta < 0.125 or > 0.875); the others carry a zero weight
      return L.sc0 * (L.a0 * p[0] + L.a1 * p[1] + L.a2 * p[2] + L.b0 * p[9] + L.b1 * p[10] + L.b2 * p[11]) +
             L.sc1 * (L.c0 * q[0] + L.c1 * q[1] + L.c2 * q[2] + L.d0 * q[9] + L.d1 * q[10] + L.d2 * q[11]);
    return L.sc0 * (L.a0 * p[0] + L.a1 * p[1] + L.b0 * p[9] + L.b1 * p[10]) + L.sc1 * (L.c0 * q[0] + L.c1 * q[1] + L.d0 * q[9] + L.d1 * q[10]);
  };
  // binary bands, upper atmosphere: 2-point eta interpolation, T stride 5
  auto major_upper = [&]() {
    const float *p = L.pA0 + off, *q = L.pA1 + off;
    return L.sc0 * (L.a0 * p[0] + L.a1 * p[1] + L.b0 * p[5] + L.b1 * p[6]) + L.sc1 * (L.c0 * q[0] + L.c1 * q[1] + L.d0 * q[5] + L.d1 * q[6]);
  };
  auto frac_eta = [&]() { const float *r = L.pFrac + off; return r[0] + L.fFrac * (r[1] - r[0]); };
  auto frac1 = [&]() { return L.pFrac[off]; };
  auto cfc = [&](int k) { return L.pCfc[off + k]; };

  taug = 0.f; fracs = 0.f;
  if (BAND == 1) {                           // x0 = colbrd * scaleminorn2, x4 = corradj
    if (L.low) taug = L.x4 * (L.sc0 * k4() + selfk() + fork() + L.x0 * minor1(L.pM0));
    else taug = L.x4 * (L.sc0 * k4() + fork() + L.x0 * minor1(L.pM0));
    fracs = frac1();
  } else if (BAND == 2) {
    if (L.low) taug = L.x4 * (L.sc0 * k4() + selfk() + fork());
    else taug = L.sc0 * k4() + fork();
    fracs = frac1();
  } else if (BAND == 3) {                    // x0 = adjcoln2o
    if (L.low) taug = major_lower2() + selfk() + fork() + L.x0 * minor2(L.pM0, 9, L.fm0);
    else taug = major_upper() + fork() + L.x0 * minor2(L.pM0, 5, L.fm0);
    fracs = frac_eta();
  } else if (BAND == 4) {
    if (L.low) taug = major_lower2() + selfk() + fork();
    else {
      taug = major_upper();
      if (ig == 8) taug = taug * 0.92f; else if (ig == 9) taug = taug * 0.88f; else if (ig == 10) taug = taug * 1.07f;
      else if (ig == 11) taug = taug * 1.1f; else if (ig == 12) taug = taug * 0.99f; else if (ig == 13) taug = taug * 0.88f;
      else if (ig == 14) taug = taug * 0.943f;
    }
    fracs = frac_eta();
  } else if (BAND == 5) {                    // x0 = wx(ccl4), x1 = colo3
    if (L.low) taug = major_lower2() + selfk() + fork() + minor2(L.pM0, 9, L.fm0) * L.x1 + L.x0 * cfc(0);
    else taug = major_upper() + L.x0 * cfc(0);
    fracs = frac_eta();
  } else if (BAND == 6) {                    // x0 = wx(cfc11), x1 = wx(cfc12), x2 = adjcolco2
    if (L.low) taug = L.sc0 * k4() + selfk() + fork() + L.x2 * minor1(L.pM0) + L.x0 * cfc(1) + L.x1 * cfc(2);
    else taug = 0.0f + L.x0 * cfc(1) + L.x1 * cfc(2);
    fracs = frac1();
  } else if (BAND == 7) {                    // x0 = adjcolco2
    if (L.low) { taug = major_lower2() + selfk() + fork() + L.x0 * minor2(L.pM0, 9, L.fm0); fracs = frac_eta(); }
    else {
      taug = L.sc0 * k4() + L.x0 * minor1(L.pM0);
      fracs = frac1();
      if (ig == 6) taug = taug * 0.92f; else if (ig == 7) taug = taug * 0.88f; else if (ig == 8) taug = taug * 1.07f;
      else if (ig == 9) taug = taug * 1.1f; else if (ig == 10) taug = taug * 0.99f; else if (ig == 11) taug = taug * 0.855f;
    }
  } else if (BAND == 8) {                    // x0 = adjcolco2, x1 = colo3, x2 = coln2o, x3 = wx(cfc12), x4 = wx(cfc22); pM0 CO2, pM1 O3, pM2 N2O
    if (L.low) taug = L.sc0 * k4() + selfk() + fork() + L.x0 * minor1(L.pM0) + L.x1 * minor1(L.pM1) + L.x2 * minor1(L.pM2) + L.x3 * cfc(2) + L.x4 * cfc(3);
    else taug = L.sc0 * k4() + L.x0 * minor1(L.pM0) + L.x2 * minor1(L.pM2) + L.x3 * cfc(2) + L.x4 * cfc(3);
    fracs = frac1();
  } else if (BAND == 9) {                    // x0 = adjcoln2o
    if (L.low) { taug = major_lower2() + selfk() + fork() + L.x0 * minor2(L.pM0, 9, L.fm0); fracs = frac_eta(); }
    else { taug = L.sc0 * k4() + L.x0 * minor1(L.pM0); fracs = frac1(); }
  } else if (BAND == 10) {
    if (L.low) taug = L.sc0 * k4() + selfk() + fork();
    else taug = L.sc0 * k4() + fork();
    fracs = frac1();
  } else if (BAND == 11) {                   // x0 = colo2 * scaleminor
    float t;
    if (L.low) t = L.sc0 * k4() + selfk() + fork();
    else t = L.sc0 * k4() + fork();
    taug = t + L.x0 * minor1(L.pM0);
    fracs = frac1();
  } else if (BAND == 12) {
    if (L.low) { taug = major_lower2() + selfk() + fork(); fracs = frac_eta(); }
  } else if (BAND == 13) {                   // x0 = adjcolco2 (lower) / colo3 (upper), x1 = colco; pM0 CO2 / O3, pM1 CO
    if (L.low) {
      taug = major_lower2() + selfk() + fork() + L.x0 * minor2(L.pM0, 9, L.fm0) + L.x1 * minor2(L.pM1, 9, L.fm1);
      fracs = frac_eta();
    } else { taug = L.x0 * minor1(L.pM0); fracs = frac1(); }
  } else if (BAND == 14) {
    if (L.low) taug = L.sc0 * k4() + selfk() + fork();
    else taug = L.sc0 * k4();
    fracs = frac1();
  } else if (BAND == 15) {                   // x0 = colbrd * scaleminor
    if (L.low) { taug = major_lower2() + selfk() + fork() + L.x0 * minor2(L.pM0, 9, L.fm0); fracs = frac_eta(); }
  } else {                                   // 16
    if (L.low) { taug = major_lower2() + selfk() + fork(); fracs = frac_eta(); }
    else { taug = L.sc0 * k4(); fracs = frac1(); }
  }
}

// Band-level setup of one layer from the layer's workspace fields fv[] (prep.cu).  S0 = slice of the group's first g-point.
template <int BAND>
__device__ __forceinline__ void lw_setup(const float (&fv)[LWC_N], const LwSmem &sm, const float *__restrict__ S0, const bool low, const float oneminus,
                                         const float (&rc)[5], LwBL &L) {
  constexpr LwK K = LWK[BAND - 1];
  const float *s_rat = sm.rat, *s_chi = sm.chi;
  auto CHI = [&](int imol, int jp) { return s_chi[(imol - 1) + 7 * (jp - 1)]; };   // 1-based like the reference
  auto RAT = [&](int r, int jpp) { return s_rat[r * 60 + jpp - 1]; };
  const int pk = __float_as_int(fv[LWC_IDX]);
  const int jp = IDX_JP(pk), jt = IDX_JT(pk), jt1 = IDX_JT1(pk), indself = IDX_SELF(pk), indfor = IDX_FOR(pk), indminor = IDX_MINOR(pk);
  L.low = low; L.three = false; L.any3 = false;
  L.fac00 = fv[LWC_FAC00]; L.fac01 = fv[LWC_FAC01]; L.fac10 = fv[LWC_FAC10]; L.fac11 = fv[LWC_FAC11];
  L.selffac = fv[LWC_SELFFAC]; L.selffrac = fv[LWC_SELFFRAC]; L.forfac = fv[LWC_FORFAC]; L.forfrac = fv[LWC_FORFRAC];
  L.minorfrac = fv[LWC_MINORFRAC];
  L.pSelf = S0 + K.oSelf + indself - 1;
  L.pFor = S0 + K.oFor + indfor - 1;
  L.pCfc = S0 + K.oCfc;
  L.pM0 = L.pM1 = L.pM2 = S0; L.fm0 = L.fm1 = 0.f; L.fFrac = 0.f;
  L.x0 = L.x1 = L.x2 = L.x3 = L.x4 = 0.f; L.sc0 = L.sc1 = 0.f;
  // single-key-species rows (k4): 4 points, the two T rows are neighbours
  auto rows4 = [&]() {
    if (low) { L.pA0 = S0 + ((jp - 1) * 5 + (jt - 1)); L.pA1 = S0 + (jp * 5 + (jt1 - 1)); }
    else { L.pA0 = S0 + K.oB + ((jp - 13) * 5 + (jt - 1)); L.pA1 = S0 + K.oB + ((jp - 12) * 5 + (jt1 - 1)); }
  };
  // binary-species rows: eta at the two pressure levels
  auto rows_bin = [&](const LwEta &e, const LwEta &e1) {
    L.sc0 = e.speccomb; L.sc1 = e1.speccomb;
    if (low) {
      const LwStencil s0 = lw_stencil(e), s1 = lw_stencil(e1);
      L.pA0 = S0 + ((jp - 1) * 5 + (jt - 1)) * 9 + s0.o;
      L.pA1 = S0 + (jp * 5 + (jt1 - 1)) * 9 + s1.o;
      L.a0 = s0.w0 * L.fac00; L.a1 = s0.w1 * L.fac00; L.a2 = s0.w2 * L.fac00;
      L.b0 = s0.w0 * L.fac10; L.b1 = s0.w1 * L.fac10; L.b2 = s0.w2 * L.fac10;
      L.c0 = s1.w0 * L.fac01; L.c1 = s1.w1 * L.fac01; L.c2 = s1.w2 * L.fac01;
      L.d0 = s1.w0 * L.fac11; L.d1 = s1.w1 * L.fac11; L.d2 = s1.w2 * L.fac11;
      L.three = s0.three || s1.three;
    } else {
      const float f0 = 1.f - e.f, f1 = 1.f - e1.f;
      L.pA0 = S0 + K.oB + ((jp - 13) * 5 + (jt - 1)) * 5 + e.j - 1;
      L.pA1 = S0 + K.oB + ((jp - 12) * 5 + (jt1 - 1)) * 5 + e1.j - 1;
      L.a0 = f0 * L.fac00; L.a1 = e.f * L.fac00; L.b0 = f0 * L.fac10; L.b1 = e.f * L.fac10;
      L.c0 = f1 * L.fac01; L.c1 = e1.f * L.fac01; L.d0 = f1 * L.fac11; L.d1 = e1.f * L.fac11;
    }
  };
  auto set_frac_eta = [&](int o, const LwEta &ep) { L.pFrac = S0 + o + ep.j - 1; L.fFrac = ep.f; };
  auto set_minor1 = [&](const float *&p, int o) { p = S0 + o + indminor - 1; };
  auto set_minor2 = [&](const float *&p, float &fm, int o, int ne, const LwEta &em) { p = S0 + o + (em.j - 1) + ne * (indminor - 1); fm = em.f; };
  // empirical column rescaling of a minor gas (e.g. LW:5208-5216)
  auto adjcol = [&](float col, int imol, float thresh, float base, float expo) {
    const float coldry = fv[LWC_COLDRY];
    const float chim = CHI(imol, jp + 1);
    const float chi = col / coldry;
    const float rat = 1.e20f * chi / chim;
    if (rat > thresh) {
      const float adjfac = base + glm::powf_(rat - base, expo);
      return adjfac * chim * coldry * 1.e-20f;
    }
    return col;
  };
  auto WX = [&](float vmr) { return fv[LWC_COLDRY] * vmr * 1.e-20f; };
  const float vccl4 = 0.093e-9f, vcfc11 = 0.251e-9f, vcfc12 = 0.538e-9f, vcfc22 = 0.169e-9f;
  // rc[] = reference ratios at fixed pressure levels: 0 rp_a, 1 rp_b, 2 rm_a, 3 rm_b, 4 rm_a3
  L.pFrac = S0 + (low ? K.oFracA : K.oFracB);

  if (BAND == 1) {
    const float pp = fv[LWC_PAVEL];
    L.x0 = fv[LWC_BRD] * fv[LWC_SCALEMINORN2];
    L.sc0 = fv[LWC_H2O]; rows4();
    if (low) { float corradj = 1.f; if (pp < 250.f) corradj = 1.f - 0.15f * (250.f - pp) / 154.4f; L.x4 = corradj; set_minor1(L.pM0, K.mA[M_N2]); }
    else { L.x4 = 1.f - 0.15f * (pp / 95.6f); set_minor1(L.pM0, K.mB[M_N2]); }
  } else if (BAND == 2) {
    L.sc0 = fv[LWC_H2O]; rows4();
    if (low) { const float pp = fv[LWC_PAVEL]; L.x4 = 1.f - .05f * (pp - 100.f) / 900.f; }
  } else if (BAND == 3) {
    const float h2o = fv[LWC_H2O], co2 = fv[LWC_CO2];
    const float mult = low ? 8.f : 4.f;
    const LwEta e = lw_eta(h2o, RAT(0, jp), co2, mult, oneminus), e1 = lw_eta(h2o, RAT(0, jp + 1), co2, mult, oneminus);
    const LwEta em = lw_eta(h2o, low ? rc[2] : rc[3], co2, mult, oneminus);
    const LwEta ep = lw_eta(h2o, low ? rc[0] : rc[1], co2, mult, oneminus);
    L.x0 = adjcol(fv[LWC_N2O], 4, 1.5f, 0.5f, 0.65f);
    rows_bin(e, e1);
    if (low) { set_minor2(L.pM0, L.fm0, K.mA[M_N2O], 9, em); set_frac_eta(K.oFracA, ep); }
    else { set_minor2(L.pM0, L.fm0, K.mB[M_N2O], 5, em); set_frac_eta(K.oFracB, ep); }
  } else if (BAND == 4) {
    const float co2 = fv[LWC_CO2];
    if (low) {
      const float h2o = fv[LWC_H2O];
      const LwEta e = lw_eta(h2o, RAT(0, jp), co2, 8.f, oneminus), e1 = lw_eta(h2o, RAT(0, jp + 1), co2, 8.f, oneminus);
      const LwEta ep = lw_eta(h2o, rc[0], co2, 8.f, oneminus);
      rows_bin(e, e1); set_frac_eta(K.oFracA, ep);
    } else {
      const float o3 = fv[LWC_O3];
      const LwEta e = lw_eta(o3, RAT(5, jp), co2, 4.f, oneminus), e1 = lw_eta(o3, RAT(5, jp + 1), co2, 4.f, oneminus);
      const LwEta ep = lw_eta(o3, rc[1], co2, 4.f, oneminus);
      rows_bin(e, e1); set_frac_eta(K.oFracB, ep);
    }
  } else if (BAND == 5) {
    const float co2 = fv[LWC_CO2];
    L.x0 = WX(vccl4);
    if (low) {
      const float h2o = fv[LWC_H2O];
      const LwEta e = lw_eta(h2o, RAT(0, jp), co2, 8.f, oneminus), e1 = lw_eta(h2o, RAT(0, jp + 1), co2, 8.f, oneminus);
      const LwEta em = lw_eta(h2o, rc[2], co2, 8.f, oneminus), ep = lw_eta(h2o, rc[0], co2, 8.f, oneminus);
      L.x1 = fv[LWC_O3];
      rows_bin(e, e1); set_minor2(L.pM0, L.fm0, K.mA[M_O3], 9, em); set_frac_eta(K.oFracA, ep);
    } else {
      const float o3 = fv[LWC_O3];
      const LwEta e = lw_eta(o3, RAT(5, jp), co2, 4.f, oneminus), e1 = lw_eta(o3, RAT(5, jp + 1), co2, 4.f, oneminus);
      const LwEta ep = lw_eta(o3, rc[1], co2, 4.f, oneminus);
      rows_bin(e, e1); set_frac_eta(K.oFracB, ep);
    }
  } else if (BAND == 6) {
    L.x0 = WX(vcfc11); L.x1 = WX(vcfc12);
    L.pFrac = S0 + K.oFracA;
    if (low) { L.x2 = adjcol(fv[LWC_CO2], 2, 3.0f, 2.0f, 0.77f); L.sc0 = fv[LWC_H2O]; rows4(); set_minor1(L.pM0, K.mA[M_CO2]); }
    else { L.pA0 = L.pA1 = S0; }
  } else if (BAND == 7) {
    if (low) {
      const float h2o = fv[LWC_H2O], o3 = fv[LWC_O3];
      const LwEta e = lw_eta(h2o, RAT(1, jp), o3, 8.f, oneminus), e1 = lw_eta(h2o, RAT(1, jp + 1), o3, 8.f, oneminus);
      const LwEta em = lw_eta(h2o, rc[2], o3, 8.f, oneminus), ep = lw_eta(h2o, rc[0], o3, 8.f, oneminus);
      L.x0 = adjcol(fv[LWC_CO2], 2, 3.0f, 3.0f, 0.79f);
      rows_bin(e, e1); set_minor2(L.pM0, L.fm0, K.mA[M_CO2], 9, em); set_frac_eta(K.oFracA, ep);
    } else {
      L.x0 = adjcol(fv[LWC_CO2], 2, 3.0f, 2.0f, 0.79f);
      L.sc0 = fv[LWC_O3]; rows4(); set_minor1(L.pM0, K.mB[M_CO2]);
    }
  } else if (BAND == 8) {
    L.x0 = adjcol(fv[LWC_CO2], 2, 3.0f, 2.0f, 0.65f);
    L.x1 = fv[LWC_O3]; L.x2 = fv[LWC_N2O]; L.x3 = WX(vcfc12); L.x4 = WX(vcfc22);
    rows4();
    if (low) { L.sc0 = fv[LWC_H2O]; set_minor1(L.pM0, K.mA[M_CO2]); set_minor1(L.pM1, K.mA[M_O3]); set_minor1(L.pM2, K.mA[M_N2O]); }
    else { L.sc0 = fv[LWC_O3]; set_minor1(L.pM0, K.mB[M_CO2]); set_minor1(L.pM2, K.mB[M_N2O]); }
  } else if (BAND == 9) {
    L.x0 = adjcol(fv[LWC_N2O], 4, 1.5f, 0.5f, 0.65f);
    if (low) {
      const float h2o = fv[LWC_H2O], ch4 = fv[LWC_CH4];
      const LwEta e = lw_eta(h2o, RAT(3, jp), ch4, 8.f, oneminus), e1 = lw_eta(h2o, RAT(3, jp + 1), ch4, 8.f, oneminus);
      const LwEta em = lw_eta(h2o, rc[2], ch4, 8.f, oneminus), ep = lw_eta(h2o, rc[0], ch4, 8.f, oneminus);
      rows_bin(e, e1); set_minor2(L.pM0, L.fm0, K.mA[M_N2O], 9, em); set_frac_eta(K.oFracA, ep);
    } else { L.sc0 = fv[LWC_CH4]; rows4(); set_minor1(L.pM0, K.mB[M_N2O]); }
  } else if (BAND == 10) {
    L.sc0 = fv[LWC_H2O]; rows4();
  } else if (BAND == 11) {
    L.sc0 = fv[LWC_H2O]; rows4();
    L.x0 = fv[LWC_O2] * fv[LWC_SCALEMINOR];
    set_minor1(L.pM0, low ? K.mA[M_O2] : K.mB[M_O2]);
  } else if (BAND == 12) {
    if (low) {
      const float h2o = fv[LWC_H2O], co2 = fv[LWC_CO2];
      const LwEta e = lw_eta(h2o, RAT(0, jp), co2, 8.f, oneminus), e1 = lw_eta(h2o, RAT(0, jp + 1), co2, 8.f, oneminus);
      const LwEta ep = lw_eta(h2o, rc[0], co2, 8.f, oneminus);
      rows_bin(e, e1); set_frac_eta(K.oFracA, ep);
    } else { L.pA0 = L.pA1 = S0; }
  } else if (BAND == 13) {
    if (low) {
      const float h2o = fv[LWC_H2O], n2o = fv[LWC_N2O], co2 = fv[LWC_CO2], coldry = fv[LWC_COLDRY];
      const LwEta e = lw_eta(h2o, RAT(2, jp), n2o, 8.f, oneminus), e1 = lw_eta(h2o, RAT(2, jp + 1), n2o, 8.f, oneminus);
      const LwEta em = lw_eta(h2o, rc[2], n2o, 8.f, oneminus), eco = lw_eta(h2o, rc[4], n2o, 8.f, oneminus);
      const LwEta ep = lw_eta(h2o, rc[0], n2o, 8.f, oneminus);
      const float chi_co2 = co2 / coldry;
      const float ratco2 = 1.e20f * chi_co2 / 3.55e-4f;
      if (ratco2 > 3.0f) { const float adjfac = 2.0f + glm::powf_(ratco2 - 2.0f, 0.68f); L.x0 = adjfac * 3.55e-4f * coldry * 1.e-20f; }
      else L.x0 = co2;
      L.x1 = fv[LWC_CO];
      rows_bin(e, e1);
      set_minor2(L.pM0, L.fm0, K.mA[M_CO2], 9, em); set_minor2(L.pM1, L.fm1, K.mA[M_CO], 9, eco); set_frac_eta(K.oFracA, ep);
    } else { L.x0 = fv[LWC_O3]; L.pA0 = L.pA1 = S0; set_minor1(L.pM0, K.mB[M_O3]); }
  } else if (BAND == 14) {
    L.sc0 = fv[LWC_CO2]; rows4();
  } else if (BAND == 15) {
    if (low) {
      const float n2o = fv[LWC_N2O], co2 = fv[LWC_CO2];
      const LwEta e = lw_eta(n2o, RAT(4, jp), co2, 8.f, oneminus), e1 = lw_eta(n2o, RAT(4, jp + 1), co2, 8.f, oneminus);
      const LwEta em = lw_eta(n2o, rc[2], co2, 8.f, oneminus), ep = lw_eta(n2o, rc[0], co2, 8.f, oneminus);
      L.x0 = fv[LWC_BRD] * fv[LWC_SCALEMINOR];
      rows_bin(e, e1); set_minor2(L.pM0, L.fm0, K.mA[M_N2], 9, em); set_frac_eta(K.oFracA, ep);
    } else { L.pA0 = L.pA1 = S0; }
  } else {  // 16
    if (low) {
      const float h2o = fv[LWC_H2O], ch4 = fv[LWC_CH4];
      const LwEta e = lw_eta(h2o, RAT(3, jp), ch4, 8.f, oneminus), e1 = lw_eta(h2o, RAT(3, jp + 1), ch4, 8.f, oneminus);
      const LwEta ep = lw_eta(h2o, rc[0], ch4, 8.f, oneminus);
      rows_bin(e, e1); set_frac_eta(K.oFracA, ep);
    } else { L.sc0 = fv[LWC_CH4]; rows4(); }
  }
}

// One layer of rtrnmc for one g-point and stream, both directions at once (LW:3207-3300 downward, 3322-3356 upward): the
// gas-only transmittance `atrans`, the source functions towards the lower (bbd) and upper (bbu) interface - they share the
// table look-up - and, where the column has cloud in the layer, the terms of  rad' = rad - rad * X + src + Z  (downward:
// srcd, Zd; upward: Yu = srcu + Zu).  The exponentials are table look-ups as in the reference.
struct LwRT { float atrans, bbd, bbu, X, srcd, Zd, Yu; };
// clear layer (no sub-column of the column is cloudy in it): series below 0.06, table above; both are evaluated and selected
// (the table index of a small optical depth is valid), no divergence
// The per-stream arithmetic below spells out every rounding (explicit fmaf / __fmul_rn / __fadd_rn): the full and the clean
// stream are two inlined copies of the same source, and left to itself the compiler contracts them differently, so that a
// clean stream with zero aerosol would no longer equal the full one bit for bit (the reference runs the same code twice).
__device__ __forceinline__ void lw_rt_clear(const float2 *__restrict__ s_et, const float bpade, const float odepth, const float plfrac,
                                            const float blay, const float dplankdn, const float dplankup, float &atrans, float &bbd, float &bbu) {
  const float tblind = div_rn(odepth, __fadd_rn(bpade, odepth));
  const int itr = (int)__fadd_rn(__fmul_rn(10000.0f, tblind), 0.5f);
  const float2 et = s_et[itr];
  const bool ser = odepth <= 0.06f;
  atrans = ser ? fmaf(-__fmul_rn(0.5f, odepth), odepth, odepth) : __fsub_rn(1.f, et.x);
  const float tf = ser ? __fmul_rn(0.166667f, odepth) : et.y;
  bbd = __fmul_rn(plfrac, fmaf(tf, dplankdn, blay));
  bbu = __fmul_rn(plfrac, fmaf(tf, dplankup, blay));
}
// layer in which the column has cloud in some sub-column (icldlyr, LW:3218-3290): three optical-depth regimes, gas-only and
// gas + cloud quantities
__device__ __forceinline__ void lw_rt_cloudy(const float2 *__restrict__ s_et, const float bpade, float odepth, const float odcld,
                                             const float efclfrac, const float cldfmc, const float plfrac, const float blay, const float dplankdn,
                                             const float dplankup, LwRT &o) {
  float odtot = __fadd_rn(odepth, odcld);
  float tf, tftot, atot;
  bool table_gas = false;
  if (odtot < 0.06f) {
    o.atrans = fmaf(-__fmul_rn(0.5f, odepth), odepth, odepth);
    tf = __fmul_rn(0.166667f, odepth);
    atot = fmaf(-__fmul_rn(0.5f, odtot), odtot, odtot);
    tftot = __fmul_rn(0.166667f, odtot);
  } else if (odepth <= 0.06f) {
    o.atrans = fmaf(-__fmul_rn(0.5f, odepth), odepth, odepth);
    tf = __fmul_rn(0.166667f, odepth);
    const float tblind = div_rn(odtot, __fadd_rn(bpade, odtot));
    const int ittot = (int)__fadd_rn(__fmul_rn(10000.0f, tblind), 0.5f);
    const float2 et = s_et[ittot];
    tftot = et.y; atot = __fsub_rn(1.f, et.x);
  } else {
    float tblind = div_rn(odepth, __fadd_rn(bpade, odepth));
    const int itgas = (int)__fadd_rn(__fmul_rn(10000.0f, tblind), 0.5f);
    // tau_tbl(itgas) recomputed with the table generator's arithmetic (LW:7944-7950)
    if (itgas >= 10000) odepth = 1.e10f;
    else { const float tfn = div_rn((float)itgas, 10000.0f); odepth = div_rn(__fmul_rn(bpade, tfn), __fsub_rn(1.0f, tfn)); }
    const float2 eg = s_et[itgas];
    o.atrans = __fsub_rn(1.f, eg.x);
    tf = eg.y;
    table_gas = true;
    odtot = __fadd_rn(odepth, odcld);
    tblind = div_rn(odtot, __fadd_rn(bpade, odtot));
    const int ittot = (int)__fadd_rn(__fmul_rn(10000.0f, tblind), 0.5f);
    const float2 et = s_et[ittot];
    tftot = et.y; atot = __fsub_rn(1.f, et.x);
  }
  o.bbd = __fmul_rn(plfrac, fmaf(tf, dplankdn, blay));
  o.bbu = __fmul_rn(plfrac, fmaf(tf, dplankup, blay));
  const float bbdtot = __fmul_rn(plfrac, fmaf(tftot, dplankdn, blay)), bbutot = __fmul_rn(plfrac, fmaf(tftot, dplankup, blay));
  o.srcd = table_gas ? __fmul_rn(__fmul_rn(o.atrans, plfrac), fmaf(tf, dplankdn, blay)) : __fmul_rn(o.bbd, o.atrans);
  const float srcu = __fmul_rn(o.bbu, o.atrans);
  o.X = fmaf(efclfrac, __fsub_rn(1.f, o.atrans), o.atrans);
  o.Zd = __fmul_rn(cldfmc, fmaf(bbdtot, atot, -o.srcd));
  o.Yu = fmaf(cldfmc, fmaf(bbutot, atot, -srcu), srcu);
}

// One (column, band group): the ng consecutive g-points of band BAND that start at g-point g0 (absolute index).
//
// Pass 1 walks the layers top-down: band-level setup, then a real (LW_UNROLL_G-fold unrolled) loop over the group's g-points -
// gas optics, layer transmittance and sources, downward radiances, whose running values live in shared memory
// [g-point][4][thread].  (A fully unrolled g-point loop with the radiances in registers was measured twice: 77 KB of code per
// pass with everything inline was instruction-fetch-bound - the SM's instruction cache holds 32 KB - and a compact version
// that called the cloudy-layer routine out of line spilled around every call.)  Pass 1 adds the downward radiances of the
// group in g order into the partial buffer and leaves, per (g-point, layer), ONE 16-byte record (atrans, bbu) x (full, clean)
// for the way back up - half of round 1's record, and the thread that wrote it reads it.  Pass 2 walks bottom-up over the
// records only: two fused multiply-adds per stream and g-point.
#ifndef LW_UNROLL_G
#define LW_UNROLL_G 1
#endif
template <int BAND>
__device__ __forceinline__ void lw_band_body(const LwArgs &a, const LwSmem &sm, float *__restrict__ st, const int grp, const int g0, const int ng,
                                             const int c, const bool live) {
  constexpr LwK K = LWK[BAND - 1];
  constexpr int SF = K.sf;
  constexpr int b = BAND - 1;
  constexpr int UG = LW_UNROLL_G;
  const float2 *s_et = sm.et;
  const float *S0 = sm.S, *s_plk = sm.plk, *s_chi = sm.chi;
  const DevTables &tb = a.tb;
  const LwWs &ws = a.ws;
  const int nlay = ws.nlay, nk = ws.nk;
  const size_t cap = ws.cap, pcap = ws.pcap;
  const float bpade = tb.bpade, oneminus = tb.oneminus;
  const bool do_clean = (a.variants & ARC_VAR_CLEAN) != 0;
  const bool do_clnc = (a.variants & ARC_VAR_CLEANCLEAR) != 0;
  const int laytrop = ws.laytrop[c];
  const float secdiff = ws.secdiff[(size_t)b * cap + c];
  const int ig0 = g0 - c_lw[b].g0 + 1;                                              // 1-based g-point of the group's first member inside the band
  auto CHI = [&](int imol, int jp) { return s_chi[(imol - 1) + 7 * (jp - 1)]; };
  // state of g-point i: 0 all-sky full, 1 clear-sky full, 2 all-sky clean, 3 clear-sky clean
  auto ST = [&](int i, int f) -> float & { return st[(i * 4 + f) * LW_BLOCK]; };

  // band constants: reference ratios at fixed pressure levels (0 rp_a, 1 rp_b, 2 rm_a, 3 rm_b, 4 rm_a3)
  float rc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  switch (BAND) {
    case 3: rc[0] = div_rn(CHI(1, 9), CHI(2, 9)); rc[1] = div_rn(CHI(1, 13), CHI(2, 13));
            rc[2] = div_rn(CHI(1, 3), CHI(2, 3)); rc[3] = div_rn(CHI(1, 13), CHI(2, 13)); break;
    case 4: rc[0] = div_rn(CHI(1, 11), CHI(2, 11)); rc[1] = div_rn(CHI(3, 13), CHI(2, 13)); break;
    case 5: rc[0] = div_rn(CHI(1, 5), CHI(2, 5)); rc[1] = div_rn(CHI(3, 43), CHI(2, 43)); rc[2] = div_rn(CHI(1, 7), CHI(2, 7)); break;
    case 7: rc[0] = div_rn(CHI(1, 3), CHI(3, 3)); rc[2] = div_rn(CHI(1, 3), CHI(3, 3)); break;
    case 9: rc[0] = div_rn(CHI(1, 9), CHI(6, 9)); rc[2] = div_rn(CHI(1, 3), CHI(6, 3)); break;
    case 12: rc[0] = div_rn(CHI(1, 10), CHI(2, 10)); break;
    case 13: rc[0] = div_rn(CHI(1, 5), CHI(4, 5)); rc[2] = div_rn(CHI(1, 1), CHI(4, 1)); rc[4] = div_rn(CHI(1, 3), CHI(4, 3)); break;
    case 15: rc[0] = div_rn(CHI(4, 1), CHI(2, 1)); rc[2] = div_rn(CHI(4, 1), CHI(2, 1)); break;
    case 16: rc[0] = div_rn(CHI(1, 6), CHI(6, 6)); break;
    default: break;
  }

  auto planck_at = [&](float t) {
    int ind = (int)(t - 159.f);
    ind = min(max(ind, 1), 180);
    const float frac = t - 159.f - (float)ind;
    const float p0 = s_plk[ind - 1], p1 = s_plk[ind];
    return p0 + frac * (p1 - p0);
  };

  // workspace fields this band reads (BAND is a compile-time constant: the other loads do not exist in this instantiation)
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
  const float *cldc = ws.cld + c + (unsigned)b * ustf;
  const uint32_t *maskc = ws.mask + (size_t)g0 * ws.W * cap + c;
  const unsigned mstride = (unsigned)ws.W * ucap;                                  // mask words between consecutive g-points
  const unsigned lstride = (unsigned)cap * LWC_N;                                  // coefficient words per layer
  // the layer's workspace words are requested one layer ahead (fvn, taua_n, tz_n) and land while the g-point loop of the
  // current layer runs: with 16 warps per SM an exposed L2 / HBM round trip per layer was 25 % of all stall samples
  float fv[LWC_N], fvn[LWC_N], taua_n = 0.f, tz_n = 0.f;
  auto load_layer = [&](int lay) {
    const float *p = coefc + (size_t)((unsigned)lay * lstride);      // fields at immediate offsets of 128 bytes
#pragma unroll
    for (int f = 0; f < LWC_N; f++) fvn[f] = ((need >> f) & 1u) ? p[f * 32] : 0.f;
    taua_n = aerc[(unsigned)lay * ucap];
    tz_n = lay > 0 ? *(p + LWC_TZ * 32 - (ptrdiff_t)lstride) : ws.colf[(size_t)LWF_TZ0 * cap + c];     // interface below the layer
  };
  auto tz_at = [&](int lev) {     // interface temperature: level 0 = surface
    return lev > 0 ? coefc[(size_t)((unsigned)(lev - 1) * lstride) + LWC_TZ * 32] : ws.colf[(size_t)LWF_TZ0 * cap + c];
  };
  float *__restrict__ bpart = ws.bpart + (size_t)grp * (nlay + 1) * nk * pcap + c;
  const unsigned oFU = ws.kslot[K_FU] * (unsigned)pcap, oFD = ws.kslot[K_FD] * (unsigned)pcap, oCU = ws.kslot[K_CU] * (unsigned)pcap,
                 oCD = ws.kslot[K_CD] * (unsigned)pcap, oNU = ws.kslot[K_NU] * (unsigned)pcap, oND = ws.kslot[K_ND] * (unsigned)pcap,
                 oXU = ws.kslot[K_XU] * (unsigned)pcap, oXD = ws.kslot[K_XD] * (unsigned)pcap;
  const size_t lvs = (size_t)nk * pcap;                                            // partial-buffer words per level
  // records [layer][g-point][column]: float4 per lane, a warp writes / reads 512 contiguous bytes
  float4 *__restrict__ rec = ws.rec + (size_t)g0 * pcap + c;
  float4 *__restrict__ recC = ws.recC + (size_t)g0 * pcap + c;
  const size_t rls = (size_t)NGLW * pcap;                                          // records per layer

  for (int i = 0; i < ng; i++) { ST(i, 0) = 0.f; ST(i, 1) = 0.f; ST(i, 2) = 0.f; ST(i, 3) = 0.f; }
  uint32_t awc = 0u;
  uint32_t mw[LW_GMAX];                      // McICA bits of the group's g-points for the current 32 layers
  int iclddn = 0;
  LwBL L;

  float plankbnd, reflect;                   // surface (LW:3303-3320)
  {
    const float emis = ws.colf[(size_t)LWF_EMISS * cap + c];
    const float tbound = ws.colf[(size_t)LWF_TBOUND * cap + c];
    int ind = (int)(tbound - 159.f);
    ind = min(max(ind, 1), 180);
    const float frac = tbound - 159.f - (float)ind;
    const float dbdtlev = s_plk[ind] - s_plk[ind - 1];
    plankbnd = emis * (s_plk[ind - 1] + frac * dbdtlev);
    reflect = 1.f - emis;
  }

  // ---------------- pass 1: downward (LW:3207-3300), top layer first
  if (live) { bpart[(size_t)nlay * lvs + oFD] = 0.f; bpart[(size_t)nlay * lvs + oCD] = 0.f;
              if (do_clean) bpart[(size_t)nlay * lvs + oND] = 0.f;
              if (do_clnc) bpart[(size_t)nlay * lvs + oXD] = 0.f; }
  float plev_up = planck_at(tz_at(nlay));
  load_layer(nlay - 1);
  for (int lay = nlay - 1; lay >= 0; lay--) {
    if ((lay & 31) == 31 || lay == nlay - 1) {
      awc = ws.anyc[(size_t)(lay >> 5) * cap + c];
#pragma unroll
      for (int i = 0; i < LW_GMAX; i++) mw[i] = i < ng ? maskc[(unsigned)i * mstride + (unsigned)(lay >> 5) * ucap] : 0u;
    }
#pragma unroll
    for (int f = 0; f < LWC_N; f++) fv[f] = fvn[f];
    const float taua = taua_n, tz_dn = tz_n;
    if (lay > 0) load_layer(lay - 1);
    const bool low = lay < laytrop;
    lw_setup<BAND>(fv, sm, S0, low, oneminus, rc, L);
    L.any3 = __any_sync(0xffffffffu, L.three);      // all 32 lanes are here: the layer loop has no early exit
    const float blay = planck_at(fv[LWC_TAVEL]);
    const float plev_dn = planck_at(tz_dn);
    const float dplankdn = plev_dn - blay, dplankup = plev_up - blay;
    plev_up = plev_dn;
    const bool icldlyr = (awc >> (lay & 31)) & 1u;
    float odcld = 0.f, abscld = 0.f, taucmc = 0.f;
    unsigned gbits = 0u;                             // bit i: g-point i of the group is cloudy in this layer
    if (icldlyr) {
      taucmc = cldc[(unsigned)lay * ucap];
      odcld = secdiff * taucmc;
      abscld = 1.f - glm::expf_(-odcld);
      iclddn = 1;
#pragma unroll
      for (int i = 0; i < LW_GMAX; i++) gbits |= ((mw[i] >> (lay & 31)) & 1u) << i;
    }
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    float z0 = 0.f, z1 = 0.f, z2 = 0.f, z3 = 0.f;          // upward radiances leaving the surface (lay == 0 only)
    float4 *rl = rec + (size_t)lay * rls, *rlC = recC + (size_t)lay * rls;        // running record pointers: + pcap per g-point
#pragma unroll UG
    for (int i = 0; i < ng; i++) {
      float taug, fracs;
      lw_gas<BAND>(L, i * SF, ig0 + i, taug, fracs);
      const float cldfmc = ((gbits >> i) & 1u) ? 1.f : 0.f;
      if (a.dbg.taug && live) {
        const size_t q = ((size_t)(ws.cols ? ws.cols[c] : a.col0 + c) * nlay + lay) * NGLW + g0 + i;
        a.dbg.taug[q] = taug; a.dbg.taur[q] = fracs;
        if (a.dbg.taucmc && cldfmc != 0.f) a.dbg.taucmc[q] = taucmc;
      }
      const float efclfrac = abscld * cldfmc;
      float r[2] = {ST(i, 0), ST(i, 2)}, rcl[2] = {ST(i, 1), ST(i, 3)};
      float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
      float od[2];
      od[0] = fmaxf(__fmul_rn(secdiff, __fadd_rn(taug, taua)), 0.f); od[1] = fmaxf(__fmul_rn(secdiff, taug), 0.f);
      if (!icldlyr) {
#pragma unroll
        for (int v = 0; v < 2; v++) {
          if (v == 1 && !do_clean) break;
          float atrans, bbd, bbu;
          lw_rt_clear(s_et, bpade, od[v], fracs, blay, dplankdn, dplankup, atrans, bbd, bbu);
          r[v] = fmaf(__fsub_rn(bbd, r[v]), atrans, r[v]);
          if (iclddn == 1) rcl[v] = fmaf(__fsub_rn(bbd, rcl[v]), atrans, rcl[v]);
          else rcl[v] = r[v];
          if (v == 0) { q4.x = atrans; q4.y = bbu; } else { q4.z = atrans; q4.w = bbu; }
        }
      } else {
        float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int v = 0; v < 2; v++) {
          if (v == 1 && !do_clean) break;
          LwRT t;
          lw_rt_cloudy(s_et, bpade, od[v], odcld, efclfrac, cldfmc, fracs, blay, dplankdn, dplankup, t);
          r[v] = __fadd_rn(__fadd_rn(fmaf(-r[v], t.X, r[v]), t.srcd), t.Zd);
          rcl[v] = fmaf(__fsub_rn(t.bbd, rcl[v]), t.atrans, rcl[v]);          // iclddn == 1 here
          if (v == 0) { q4.x = t.atrans; q4.y = t.bbu; c4.x = t.X; c4.y = t.Yu; }
          else { q4.z = t.atrans; q4.w = t.bbu; c4.z = t.X; c4.w = t.Yu; }
        }
        __stcs(rlC, c4);
      }
      __stcs(rl, q4);
      rl += pcap; rlC += pcap;
      // downward radiances at the lower interface of the layer, summed over the group's g-points in index order
      s0 = s0 + r[0]; s1 = s1 + rcl[0]; s2 = s2 + r[1]; s3 = s3 + rcl[1];
      if (lay == 0) {
        const float rad0 = fracs * plankbnd;            // fracs of the lowest layer (LW:3305)
#pragma unroll
        for (int v = 0; v < 2; v++) { r[v] = fmaf(reflect, r[v], rad0); rcl[v] = fmaf(reflect, rcl[v], rad0); }
        z0 = z0 + r[0]; z1 = z1 + rcl[0]; z2 = z2 + r[1]; z3 = z3 + rcl[1];
      }
      ST(i, 0) = r[0]; ST(i, 1) = rcl[0]; ST(i, 2) = r[1]; ST(i, 3) = rcl[1];
    }
    if (live) {
      float *bp = bpart + (size_t)lay * lvs;
      __stcs(bp + oFD, s0); __stcs(bp + oCD, s1);
      if (do_clean) __stcs(bp + oND, s2);
      if (do_clnc) __stcs(bp + oXD, s3);
      if (lay == 0) {
        __stcs(bp + oFU, z0); __stcs(bp + oCU, z1);
        if (do_clean) __stcs(bp + oNU, z2);
        if (do_clnc) __stcs(bp + oXU, z3);
      }
    }
  }
  // ---------------- pass 2: upward (LW:3322-3356), bottom layer first, from the records of pass 1 (requested one layer ahead)
  float ru[LW_GMAX][2], rcu[LW_GMAX][2];
  float4 qn[LW_GMAX];
  {
    const float4 *p = rec;
#pragma unroll
    for (int i = 0; i < LW_GMAX; i++) {
      ru[i][0] = i < ng ? ST(i, 0) : 0.f; rcu[i][0] = i < ng ? ST(i, 1) : 0.f;
      ru[i][1] = i < ng ? ST(i, 2) : 0.f; rcu[i][1] = i < ng ? ST(i, 3) : 0.f;
      qn[i] = i < ng ? __ldcs(p) : make_float4(0.f, 0.f, 0.f, 0.f);
      p += pcap;
    }
  }
  const float4 *rl = rec, *rlC = recC;
  float *bp = bpart;
  for (int lay = 0; lay < nlay; lay++) {
    if ((lay & 31) == 0) awc = ws.anyc[(size_t)(lay >> 5) * cap + c];
    const bool icldlyr = (awc >> (lay & 31)) & 1u;
    float4 q4[LW_GMAX];
#pragma unroll
    for (int i = 0; i < LW_GMAX; i++) q4[i] = qn[i];
    if (lay + 1 < nlay) {
      const float4 *p = rl + rls;
#pragma unroll
      for (int i = 0; i < LW_GMAX; i++) { if (i < ng) qn[i] = __ldcs(p); p += pcap; }
    }
#if LW_PF > 0
    // the records were written a whole downward sweep ago and come from HBM: lines LW_PF layers ahead are pulled into L2 so
    // that the register prefetch above (one layer ahead) finds them there
    if (lay + LW_PF < nlay) {
      const float4 *p = rl + (size_t)LW_PF * rls;
#pragma unroll
      for (int i = 0; i < LW_GMAX; i++) { if (i < ng) asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); p += pcap; }
    }
    if (((lay + 1) & 31) != 0 && ((awc >> ((lay + 1) & 31)) & 1u)) {
      const float4 *p = rlC + rls;
#pragma unroll
      for (int i = 0; i < LW_GMAX; i++) { if (i < ng) asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); p += pcap; }
    }
#endif
    if (icldlyr) {          // rad' = rad - rad X + Y
      const float4 *p = rlC;
#pragma unroll
      for (int i = 0; i < LW_GMAX; i++) {
        if (i < ng) { const float4 c4 = __ldcs(p); ru[i][0] = __fadd_rn(fmaf(-ru[i][0], c4.x, ru[i][0]), c4.y); ru[i][1] = __fadd_rn(fmaf(-ru[i][1], c4.z, ru[i][1]), c4.w); }
        p += pcap;
      }
    } else {
#pragma unroll
      for (int i = 0; i < LW_GMAX; i++) { ru[i][0] = fmaf(__fsub_rn(q4[i].y, ru[i][0]), q4[i].x, ru[i][0]); ru[i][1] = fmaf(__fsub_rn(q4[i].w, ru[i][1]), q4[i].z, ru[i][1]); }
    }
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;          // upward radiances at the top of the layer; g-points beyond ng hold zeros
    if (iclddn == 1) {      // iclddn as the downward sweep left it: any cloud in the column
#pragma unroll
      for (int i = 0; i < LW_GMAX; i++) { rcu[i][0] = fmaf(__fsub_rn(q4[i].y, rcu[i][0]), q4[i].x, rcu[i][0]); rcu[i][1] = fmaf(__fsub_rn(q4[i].w, rcu[i][1]), q4[i].z, rcu[i][1]); }
    } else {
#pragma unroll
      for (int i = 0; i < LW_GMAX; i++) { rcu[i][0] = ru[i][0]; rcu[i][1] = ru[i][1]; }
    }
#pragma unroll
    for (int i = 0; i < LW_GMAX; i++) { s0 = s0 + ru[i][0]; s1 = s1 + rcu[i][0]; s2 = s2 + ru[i][1]; s3 = s3 + rcu[i][1]; }
    rl += rls; rlC += rls; bp += lvs;
    if (live) {
      __stcs(bp + oFU, s0); __stcs(bp + oCU, s1);
      if (do_clean) __stcs(bp + oNU, s2);
      if (do_clnc) __stcs(bp + oXU, s3);
    }
  }
}

// Block = LW_BLOCK columns x one band group.  Blocks of one column tile are neighbours in launch order (its coefficient lines
// are re-read from L2 by the 23 groups); thread 0 stages the exp / tfn table, the group's table slices and the band's Planck
// column with TMA bulk copies.
__global__ void __launch_bounds__(LW_BLOCK, 1) k_lw_band(LwArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *s_et = reinterpret_cast<float2 *>(smem_raw);                // 10002 x (exp_tbl, tfn_tbl)
  float *S = reinterpret_cast<float *>(s_et + 10002);                 // the group's slices
  float *s_plk = S + LW_GMAX * LW_SLICE_MAX;                          // totplnk(1:181, band), padded to 184
  float *s_rat = s_plk + 184;                                         // [6][60] chi_mls ratios by jp
  float *s_chi = s_rat + 6 * 60;                                      // chi_mls(7,59)
  float *s_state = s_chi + 416;                                       // [LW_GMAX][4][LW_BLOCK] radiances
  uint64_t *bar = reinterpret_cast<uint64_t *>(s_state + LW_GMAX * 4 * LW_BLOCK);

  const int ngrp = a.ngroups;
  const int tile = blockIdx.x / ngrp, grp = blockIdx.x % ngrp;
  const int b = c_lw_grp_band[grp];
  const int g0 = c_lw_grp_g0[grp];
  const LwBandDesc &D = c_lw[b];
  // number of g-points of this group = distance to the next group's first g-point (or the band's end)
  const int gend = (grp + 1 < ngrp && c_lw_grp_band[grp + 1] == b) ? c_lw_grp_g0[grp + 1] : D.g0 + D.ng;
  const int ng = gend - g0;
  const DevTables &tb = a.tb;
  {
    StageReq req[3] = {{s_et, tb.lw_exptfn, 10002 * 8},
                       {S, tb.lw_tab + D.slice_base + (size_t)D.slice_floats * (g0 - D.g0), (uint32_t)(D.slice_floats * ng) * 4},
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
  int c = tile * LW_BLOCK + threadIdx.x;
  const bool live = c < a.ncols;              // every lane stays: the band bodies use warp votes
  if (!live) c = a.ncols - 1;
  const LwSmem sm{s_et, S, s_plk, s_rat, s_chi};
  float *st = s_state + threadIdx.x;
#define LWB(B_) case B_ - 1: lw_band_body<B_>(a, sm, st, grp, g0, ng, c, live); break;
  switch (b) {
    LWB(1) LWB(2) LWB(3) LWB(4) LWB(5) LWB(6) LWB(7) LWB(8) LWB(9) LWB(10) LWB(11) LWB(12) LWB(13) LWB(14) LWB(15) LWB(16)
    default: break;
  }
#undef LWB
}

static int lw_band_smem() { return 10002 * 8 + (LW_GMAX * LW_SLICE_MAX + 184 + 6 * 60 + 416 + LW_GMAX * 4 * LW_BLOCK) * 4 + 16; }

int lw_sweep_groups() { return h_lw_grp.n; }
void launch_lw_band(const LwArgs &a, cudaStream_t s) {
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_lw_band, cudaFuncAttributeMaxDynamicSharedMemorySize, lw_band_smem()); attr = true; }
  const int ntiles = (a.ncols + LW_BLOCK - 1) / LW_BLOCK;
  k_lw_band<<<ntiles * h_lw_grp.n, LW_BLOCK, lw_band_smem(), s>>>(a);
  count_launch();
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
  const int tc = active ? (ws.cols ? ws.cols[c] : a.col0 + c) : 0;
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
